"""pivlfn: B200-native (sm_100a) PIV-LiteFlowNet forward pass behind the reference's Python API.

Host code is Python/PyTorch (device memory, streams, torch.distributed); every arithmetic step of
the forward pass runs in hand-written CUDA reached through the C ABI in ``include/pivlfn.h``
(``libpivlfn.so``).  There is no CPU fallback: CUDA tensors are required and a missing library is
an ImportError/RuntimeError, never a silent eager path.
"""
from .arch import CFGS  # noqa: F401
