"""Frame ingest for host-resident video (upstream anchor: README.md:42,
"provide the input data (e.g., video feed or images)").

``Tracker.step(frame)`` accepts a CPU frame and uploads it on the compute stream, i.e. the copy sits
in front of the step.  ``FrameIngest`` takes it off the critical path: frames are staged in pinned
host memory and uploaded on a side stream into one of ``depth`` device buffers while the previous
frame is still being tracked; the compute stream only waits on the upload's event.

    ingest = FrameIngest((1080, 1920, 3), device="cuda")
    for state in track(tracker, frames, ingest):      # frames: any iterable of uint8 [H,W,3]
        ...

Host code only (streams, events, pinned memory); no kernels of its own.
"""
from __future__ import annotations

import numpy as np
import torch


class FrameIngest:
    def __init__(self, frame_shape, device="cuda", depth: int = 2):
        if not torch.cuda.is_available():
            raise RuntimeError("FrameIngest needs a CUDA device")
        if len(frame_shape) != 3 or frame_shape[2] != 3:
            raise ValueError("frame_shape must be (H, W, 3)")
        if depth < 2:
            raise ValueError("depth must be >= 2 (one buffer in use, one being filled)")
        self.shape = tuple(int(v) for v in frame_shape)
        self.device = torch.device(device)
        self.depth = int(depth)
        self._host = [torch.empty(self.shape, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        self._dev = [torch.empty(self.shape, dtype=torch.uint8, device=self.device) for _ in range(depth)]
        self._ready = [torch.cuda.Event() for _ in range(depth)]
        self._released = [None] * depth        # event after which the compute stream is done with a slot
        self._stream = torch.cuda.Stream(device=self.device)
        self._next = 0
        self.bytes_per_frame = int(np.prod(self.shape))

    def submit(self, frame) -> int:
        """Stages `frame` (numpy / CPU torch uint8 [H,W,3]) and starts its upload.  Returns the slot."""
        if isinstance(frame, np.ndarray):
            frame = torch.from_numpy(frame)
        if frame.dtype != torch.uint8 or tuple(frame.shape) != self.shape:
            raise ValueError(f"frame must be uint8 {self.shape}, got {frame.dtype} {tuple(frame.shape)}")
        k = self._next
        self._next = (k + 1) % self.depth
        if self._released[k] is not None:
            self._released[k].synchronize()     # the step that consumed this slot has finished reading it
            self._released[k] = None
        if frame.is_cuda or frame.is_pinned():   # device frame / caller-pinned host frame: copy as is
            src = frame
        else:
            self._host[k].copy_(frame)           # pageable -> pinned (host memcpy)
            src = self._host[k]
        with torch.cuda.stream(self._stream):
            self._dev[k].copy_(src, non_blocking=True)
            self._ready[k].record(self._stream)
        return k

    def get(self, slot: int) -> torch.Tensor:
        """Device frame of `slot`; the current stream waits for its upload (no host sync)."""
        torch.cuda.current_stream(self.device).wait_event(self._ready[slot])
        return self._dev[slot]

    def release(self, slot: int) -> None:
        """Marks everything enqueued so far on the current stream as the last reader of `slot`."""
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))
        self._released[slot] = ev


def track(tracker, frames, ingest: FrameIngest | None = None):
    """Generator: ``tracker.step`` over an iterable of host frames with the upload of frame t+1
    overlapped with the tracking of frame t.  Yields one State per frame."""
    it = iter(frames)
    try:
        first = next(it)
    except StopIteration:
        return
    if ingest is None:
        ingest = FrameIngest(tuple(first.shape), device=tracker.device)
    slot = ingest.submit(first)
    for nxt in it:
        tracker.step_begin(ingest.get(slot))     # frame t: enqueued, the GPU is busy from here on
        ingest.release(slot)
        nslot = ingest.submit(nxt)               # frame t+1: host staging + upload beside it
        slot = nslot
        yield tracker.step_end()
    tracker.step_begin(ingest.get(slot))
    ingest.release(slot)
    yield tracker.step_end()
