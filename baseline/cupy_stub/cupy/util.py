"""``cupy.util.memoize`` (CuPy 7): cache a function's results per argument tuple (and per current device)."""
import functools


def memoize(for_each_device=False):
    def decorator(fn):
        cache = {}

        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            dev = -1
            if for_each_device:
                try:
                    import torch
                    dev = torch.cuda.current_device() if torch.cuda.is_available() else -1
                except Exception:
                    dev = -1
            key = (dev, args, tuple(sorted(kwargs.items())))
            if key not in cache:
                cache[key] = fn(*args, **kwargs)
            return cache[key]
        return wrapper
    return decorator
