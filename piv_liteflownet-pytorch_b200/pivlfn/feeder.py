"""Keep the GPU busy between host batches: uploads of batch k+1 and the download of batch k-1's flow run on two side streams
while batch k's forward runs on the caller's stream (three-stage software pipeline over double-buffered device inputs and
pinned host outputs; push(k) enqueues upload k and then runs forward k-1).  Used by ``run.py`` and by the end-to-end leg of ``bench.py``; every step still pays its own
host-to-device copy of both images and its device-to-host copy of the flow -- they just overlap the neighbours' compute.

    feeder = Feeder(net, device)
    for a_host, b_host in batches:            # pinned host tensors, [B,3,H,W] float32 (or uint8 [B,H,W,3] with unpack=...)
        for tag, flow_host, event in feeder.push(a_host, b_host, tag):     # items older than `depth` batches, oldest first
            event.synchronize(); consume(flow_host); feeder.recycle(flow_host)
    ... same for feeder.drain()
"""
from __future__ import annotations

import threading
from collections import deque
from typing import Callable, List, Optional, Tuple

import torch


class Feeder:
    def __init__(self, net, device, unpack: Optional[Callable[[torch.Tensor], torch.Tensor]] = None, depth: int = 2,
                 to_host: Optional[Callable[[torch.Tensor], torch.Tensor]] = None):
        self.net, self.device, self.unpack, self.depth = net, torch.device(device), unpack, max(2, depth)
        self.to_host = to_host                       # device flow -> device tensor laid out like the host buffer (default: as is)
        self.up = torch.cuda.Stream(self.device)
        self.down = torch.cuda.Stream(self.device)
        self._dev: List[Optional[Tuple[torch.Tensor, torch.Tensor]]] = [None] * self.depth
        self._free = [None] * self.depth             # event: the slot's device inputs may be overwritten
        self._inflight: deque = deque()
        self._pending = None                         # the batch whose upload is enqueued and whose forward has not run yet
        self._k = 0
        self._pool: List[torch.Tensor] = []
        self._lock = threading.Lock()

    def _slot_tensors(self, s: int, a_host: torch.Tensor, b_host: torch.Tensor):
        slot = self._dev[s]
        if slot is None or slot[0].shape != a_host.shape or slot[0].dtype != a_host.dtype:
            # The caching allocator hands out memory that is free in the order of the CURRENT stream; the upload stream must not
            # touch it before everything already queued there (which may still use a recycled block) has finished, and the
            # allocator must know that the upload stream uses these tensors too.
            cur = torch.cuda.current_stream(self.device)
            self._dev[s] = (torch.empty(a_host.shape, dtype=a_host.dtype, device=self.device),
                            torch.empty(b_host.shape, dtype=b_host.dtype, device=self.device))
            self.up.wait_stream(cur)
            for t in self._dev[s]:
                t.record_stream(self.up)
        return self._dev[s]

    def push(self, a_host: torch.Tensor, b_host: torch.Tensor, tag=None):
        """Enqueue one batch; returns the (tag, flow_host, event) items that are older than ``depth`` batches -- wait for ``event``
        before reading ``flow_host`` (a pinned tensor that now belongs to the caller).

        The upload of THIS batch is enqueued first, then the forward of the PREVIOUS one runs: the model's forward may block the
        host (the synchronous fp16 range check of large batches reads a device flag), and an upload that is only enqueued after
        that returns cannot overlap the forward it was meant to hide behind."""
        s = self._k % self.depth
        da, db = self._slot_tensors(s, a_host, b_host)
        with torch.cuda.stream(self.up):
            if self._free[s] is not None:
                self.up.wait_event(self._free[s])     # the forward that read (and mean-subtracted) this slot has finished
            da.copy_(a_host, non_blocking=True)
            db.copy_(b_host, non_blocking=True)
            uploaded = torch.cuda.Event()
            uploaded.record(self.up)
        prev, self._pending = self._pending, (s, da, db, uploaded, tag)
        self._k += 1
        if prev is not None:
            self._forward(prev)
        out = []
        while len(self._inflight) >= self.depth:       # its host buffer is the next one to be reused: hand it out now
            out.append(self._inflight.popleft())
        return out

    def _forward(self, item) -> None:
        s, da, db, uploaded, tag = item
        cur = torch.cuda.current_stream(self.device)
        cur.wait_event(uploaded)
        with torch.no_grad():
            x1, x2 = (self.unpack(da), self.unpack(db)) if self.unpack is not None else (da, db)
            flow = self.net(x1, x2)
            if self.to_host is not None:
                flow = self.to_host(flow)
        done = torch.cuda.Event()
        done.record(cur)
        self._free[s] = done
        host = self._take_host(flow)
        with torch.cuda.stream(self.down):
            self.down.wait_event(done)
            host.copy_(flow, non_blocking=True)
            flow.record_stream(self.down)
            landed = torch.cuda.Event()
            landed.record(self.down)
        self._inflight.append((tag, host, landed))

    def _take_host(self, like: torch.Tensor) -> torch.Tensor:
        """A pinned host tensor for one batch's flow: from the pool of recycled ones if the shape fits, else newly pinned
        (cudaHostAlloc of tens of MB blocks for milliseconds, so steady state must not allocate)."""
        with self._lock:
            for i, t in enumerate(self._pool):
                if t.shape == like.shape and t.dtype == like.dtype:
                    return self._pool.pop(i)
        return torch.empty(like.shape, dtype=like.dtype, pin_memory=True)

    def recycle(self, host: torch.Tensor) -> None:
        """Hand a flow buffer received from push() / drain() back once it has been consumed (thread-safe)."""
        with self._lock:
            if len(self._pool) < 8:
                self._pool.append(host)

    def drain(self):
        if self._pending is not None:
            prev, self._pending = self._pending, None
            self._forward(prev)
        out = list(self._inflight)
        self._inflight.clear()
        return out

    def join(self):
        """Make the caller's stream wait for everything enqueued on the side streams."""
        cur = torch.cuda.current_stream(self.device)
        for st in (self.up, self.down):
            ev = torch.cuda.Event()
            ev.record(st)
            cur.wait_event(ev)
