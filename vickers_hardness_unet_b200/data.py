"""Host -> device input prefetch for the training loop (drop-in around the reference's DataLoader iteration,
/root/reference/train.py:428-434: `for imgs, masks in loader: imgs = imgs.to(device, non_blocking=True) ...`).

`DevicePrefetcher(loader, device)` yields the same (imgs, masks) pairs already resident on the GPU: batch k+1 is uploaded
from pinned host memory on a side stream while batch k trains, into one of two device buffer pairs; the compute stream
waits on the copy's event only.  Every batch is still copied host -> device exactly once per step."""
from __future__ import annotations

import torch


class DevicePrefetcher:
    def __init__(self, loader, device, depth: int = 2):
        self.loader, self.device = loader, torch.device(device)
        self.stream = torch.cuda.Stream(self.device)
        self.depth = max(2, int(depth))
        self._slots = [None] * self.depth      # device tensors per slot (allocated on first use, reused afterwards)
        self._done = [None] * self.depth       # the compute stream's last use of the slot

    def _upload(self, k, batch):
        cur = torch.cuda.current_stream(self.device)
        s = k % self.depth
        with torch.cuda.stream(self.stream):
            if self._done[s] is not None:
                self.stream.wait_event(self._done[s])   # the step that read this slot has finished
            if self._slots[s] is None or any(d.shape != h.shape or d.dtype != h.dtype
                                             for d, h in zip(self._slots[s], batch)):
                self._slots[s] = [torch.empty(h.shape, dtype=h.dtype, device=self.device) for h in batch]
            for d, h in zip(self._slots[s], batch):
                d.copy_(h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        del cur
        return s, ev

    def __iter__(self):
        it = iter(self.loader)
        k = 0
        try:
            nxt = self._upload(k, next(it))
        except StopIteration:
            return
        while nxt is not None:
            s, ev = nxt
            try:
                k += 1
                nxt = self._upload(k, next(it))     # batch k+1 uploads while batch k trains
            except StopIteration:
                nxt = None
            cur = torch.cuda.current_stream(self.device)
            cur.wait_event(ev)
            yield tuple(self._slots[s])
            done = torch.cuda.Event()
            done.record(cur)
            self._done[s] = done

    def __len__(self):
        return len(self.loader)
