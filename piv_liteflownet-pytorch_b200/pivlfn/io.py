"""Host-side input / output pipeline of the batch driver, built for several hundred
pairs per second per GPU instead of the reference's one-pair-per-step DataLoader loop (run.py:137-168):

  PairIndex     which (image 1, image 2, output stem) triples a directory holds -- the naming rules of the reference's
                ``Run`` dataset (src/datasets.py:438-487): ``*_img1.ext`` + ``*_img2.ext`` files, or n sequential frames
                giving n-1 pairs -- as a plain list that can be sliced per rank
  BatchReader   decoder threads fill PINNED uint8 staging buffers [B,H,W,3] a few batches ahead of the GPU (PIL releases the
                GIL while decoding); a batch is a run of consecutive pairs of one frame size.  Frames travel to the device as
                uint8 (a quarter of the fp32 bytes) and are unpacked to the model's [B,3,H,W] fp32 in [0,1] there.
                Optional brightness / contrast factors (run.py:86-95) are applied by PIL's own ImageEnhance in the decoder
                threads, i.e. with exactly the integer arithmetic torchvision's adjust_brightness / adjust_contrast use.
  FloWriter     Middlebury .flo files written from pinned host buffers by worker threads once the device-to-host copy
                they wait for has finished; the GPU never waits for the disk.
  read_flo / write_flo   the wire format: float32 202021.25, int32 width, int32 height, H*W*bands float32 (HWC).
"""
from __future__ import annotations

import os
import queue
import struct
import threading
from concurrent.futures import Future, ThreadPoolExecutor
from dataclasses import dataclass
from typing import Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

IMAGE_SUFFIXES = ("jpg", "jpeg", "png", "bmp", "tif", "ppm")
FLO_MAGIC = 202021.25          # the four bytes "PIEH" read as a little-endian float32


# ---------------------------------------------------------------------------------------------------------------- .flo
def write_flo(path: str, field: np.ndarray) -> None:
    """field: [H, W, 2] (u, v) or [H, W, 3] (stereo u, v, w)."""
    if not isinstance(path, str) or not path.endswith(".flo"):
        raise AssertionError(f"a .flo path is required, got {path!r}")
    arr = np.asarray(field)
    if arr.ndim != 3 or arr.shape[2] not in (2, 3):
        raise AssertionError(f"flow fields are [H, W, 2] or [H, W, 3], got {arr.shape}")
    h, w = arr.shape[:2]
    with open(path, "wb") as fh:
        fh.write(struct.pack("<fii", FLO_MAGIC, w, h))
        fh.write(np.ascontiguousarray(arr, dtype="<f4").tobytes())


def read_flo(source, bands: int = 2) -> np.ndarray:
    """source: a path or an open binary file.  Returns [H, W, bands] float32."""
    own = isinstance(source, str)
    if own:
        if not source.endswith(".flo") or not os.path.isfile(source):
            raise AssertionError(f"not a readable .flo file: {source!r}")
        fh = open(source, "rb")
    elif hasattr(source, "read"):
        fh = source
    else:
        raise AssertionError(f"a path or a binary file object is required, got {type(source).__name__}")
    try:
        head = fh.read(12)
        if len(head) != 12:
            raise AssertionError("truncated .flo header")
        magic, w, h = struct.unpack("<fii", head)
        if magic != FLO_MAGIC:
            raise AssertionError(f"bad .flo magic {magic!r}")
        if not (0 < w < 100000 and 0 < h < 100000):
            raise AssertionError(f"implausible .flo size {w} x {h}")
        n = h * w * bands
        data = np.frombuffer(fh.read(4 * n), dtype="<f4")
        if data.size != n:
            raise AssertionError("truncated .flo payload")
        return data.reshape(h, w, bands).astype(np.float32)
    finally:
        if own:
            fh.close()


# ---------------------------------------------------------------------------------------------------------- pair index
@dataclass(frozen=True)
class Pair:
    first: str
    second: str
    stem: str            # name of the output file without directory and suffix


def list_images(folder: str, only_first_of_pair: bool, both_cases: bool) -> List[str]:
    """Image files of a folder grouped by suffix in the order of IMAGE_SUFFIXES, sorted by name inside a group (the order
    the reference's glob-per-extension listing produces, src/utils_data.py:12-33)."""
    names = os.listdir(folder)
    out: List[str] = []
    for suf in IMAGE_SUFFIXES:
        variants = (suf, suf.upper()) if both_cases else (suf,)
        for v in variants:
            tail = ("_img1." if only_first_of_pair else ".") + v
            out += [os.path.join(folder, n) for n in sorted(names) if n.endswith(tail)]
    return out


class PairIndex(Sequence):
    def __init__(self, folder: str, paired_files: bool, count: int = -1, offset: int = 0):
        if not os.path.isdir(folder):
            raise ValueError(f"Input image directory is NOT found! '{folder}'")
        files = list_images(folder, paired_files, both_cases=False)
        files = files[offset:] if count < 0 else files[offset:offset + count]
        pairs: List[Pair] = []
        if paired_files:
            for f in files:
                base, suffix = os.path.splitext(os.path.basename(f))
                stem = base[:base.rfind("_")] if "_" in base else base
                pairs.append(Pair(f, os.path.join(folder, stem + "_img2" + suffix), stem))
        else:
            for a, b in zip(files, files[1:]):                     # n frames -> n - 1 pairs
                pairs.append(Pair(a, b, os.path.splitext(os.path.basename(a))[0]))
        self._pairs = [p for p in pairs if os.path.isfile(p.first) and os.path.isfile(p.second)]

    def __len__(self) -> int:
        return len(self._pairs)

    def __getitem__(self, i):
        return self._pairs[i]


# -------------------------------------------------------------------------------------------------------------- decode
def decode_rgb(path: str, brightness: float = 1.0, contrast: float = 1.0) -> np.ndarray:
    """[H, W, 3] uint8.  Factors != 1 go through PIL.ImageEnhance (what torchvision's PIL backend calls, run.py:86-95)."""
    import PIL.Image
    im = PIL.Image.open(path).convert("RGB")
    if brightness != 1.0 or contrast != 1.0:
        import PIL.ImageEnhance
        im = PIL.ImageEnhance.Brightness(im).enhance(brightness)
        im = PIL.ImageEnhance.Contrast(im).enhance(contrast)
    return np.asarray(im, dtype=np.uint8)


def unpack_u8(batch_u8: torch.Tensor) -> torch.Tensor:
    """[B, H, W, 3] uint8 on the device -> [B, 3, H, W] float32 in [0, 1] (torchvision.transforms.ToTensor semantics)."""
    return batch_u8.permute(0, 3, 1, 2).to(torch.float32).div_(255.0).contiguous()


@dataclass
class Batch:
    first: torch.Tensor          # pinned uint8 [B, H, W, 3]
    second: torch.Tensor
    stems: List[str]
    _slot: int = -1


class BatchReader:
    """Iterates Batches of up to ``batch`` consecutive same-size pairs of ``pairs[lo:hi]``; ``depth`` batches are decoded
    ahead by ``workers`` threads into a ring of pinned buffers.  ``release(batch)`` returns a batch's buffers to the ring
    (call it once the host-to-device copies that read them have been enqueued and synchronised)."""

    def __init__(self, pairs: Sequence[Pair], lo: int, hi: int, batch: int, depth: int = 3, workers: int = 8,
                 brightness: float = 1.0, contrast: float = 1.0, pin: bool = True):
        self.pairs, self.lo, self.hi, self.batch = pairs, lo, hi, max(1, batch)
        self.factors = (float(brightness), float(contrast))
        self.pin = pin and torch.cuda.is_available()
        self._pool = ThreadPoolExecutor(max_workers=max(1, workers))
        self._ready: "queue.Queue[Optional[Batch]]" = queue.Queue(maxsize=max(1, depth))
        self._free: "queue.Queue[int]" = queue.Queue()
        self._buffers: dict = {}
        self._nslots = max(1, depth) + 1
        for s in range(self._nslots):
            self._free.put(s)
        self._error: Optional[BaseException] = None
        self._thread = threading.Thread(target=self._produce, daemon=True)
        self._thread.start()

    def _staging(self, slot: int, shape: Tuple[int, ...]) -> Tuple[torch.Tensor, torch.Tensor]:
        buf = self._buffers.get(slot)
        if buf is None or buf[0].shape != shape:
            mk = lambda: torch.empty(shape, dtype=torch.uint8).pin_memory() if self.pin else torch.empty(shape, dtype=torch.uint8)
            buf = (mk(), mk())
            self._buffers[slot] = buf
        return buf

    def _produce(self):
        try:
            b, c = self.factors
            window = 2 * self.batch                          # pairs being decoded ahead of the batch that is being assembled
            futs = {}
            nxt = self.lo

            def item(k):
                nonlocal nxt
                while nxt < self.hi and nxt < k + window:
                    p = self.pairs[nxt]
                    futs[nxt] = (self._pool.submit(decode_rgb, p.first, b, c), self._pool.submit(decode_rgb, p.second, b, c), p.stem)
                    nxt += 1
                fa, fb, stem = futs.pop(k)
                x, y = fa.result(), fb.result()
                if x.shape != y.shape:
                    raise ValueError(f"the two frames of pair '{stem}' differ in size")
                return x, y, stem

            group = []
            for k in range(self.lo, self.hi):
                it = item(k)
                if group and (len(group) == self.batch or it[0].shape != group[0][0].shape):
                    self._emit(group)                        # a batch is a run of consecutive pairs of one frame size
                    group = []
                group.append(it)
            if group:
                self._emit(group)
        except BaseException as ex:      # surfaced in the consumer thread
            self._error = ex
        finally:
            self._ready.put(None)

    def _emit(self, group):
        slot = self._free.get()
        h, w, _ = group[0][0].shape
        a, b = self._staging(slot, (len(group), h, w, 3))
        for k, (x, y, _) in enumerate(group):
            a[k].numpy()[...] = x
            b[k].numpy()[...] = y
        self._ready.put(Batch(a, b, [g[2] for g in group], slot))

    def __iter__(self) -> Iterator[Batch]:
        while True:
            item = self._ready.get()
            if item is None:
                if self._error is not None:
                    raise self._error
                return
            yield item

    def release(self, batch: Batch):
        if batch._slot >= 0:
            self._free.put(batch._slot)
            batch._slot = -1

    def close(self):
        self._pool.shutdown(wait=False, cancel_futures=True)


# -------------------------------------------------------------------------------------------------------------- writer
class FloWriter:
    """``submit(host, stems, event)``: host is a pinned [B, H, W, bands] float32 tensor that a device-to-host copy is filling;
    a worker waits for ``event`` and writes one ``<stem><suffix>`` file per sample into ``folder``."""

    def __init__(self, folder: str, suffix: str = "_out.flo", workers: int = 4):
        os.makedirs(folder, exist_ok=True)
        self.folder, self.suffix = folder, suffix
        self._pool = ThreadPoolExecutor(max_workers=max(1, workers))
        self._pending: List[Future] = []
        self.written: List[str] = []
        self._lock = threading.Lock()

    def path_for(self, stem: str) -> str:
        return os.path.join(self.folder, stem + self.suffix)

    def _write(self, host: torch.Tensor, stems: List[str], event, done) -> None:
        if event is not None:
            event.synchronize()
        arr = host.numpy()
        for k, stem in enumerate(stems):
            path = self.path_for(stem)
            write_flo(path, arr[k])
            with self._lock:
                self.written.append(path)
        if done is not None:
            done(host)                   # e.g. Feeder.recycle: the pinned buffer may be reused

    def submit(self, host: torch.Tensor, stems: List[str], event=None, done=None) -> None:
        self._pending.append(self._pool.submit(self._write, host, list(stems), event, done))

    def drain(self) -> List[str]:
        for f in self._pending:
            f.result()
        self._pending.clear()
        return list(self.written)

    def close(self) -> List[str]:
        out = self.drain()
        self._pool.shutdown()
        return out
