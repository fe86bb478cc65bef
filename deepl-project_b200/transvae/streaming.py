"""Host-to-host reconstruction with the copies off the compute stream.

The reference's inference loop (inference_example.py:56-73) does ``images.to(device)`` -> ``encode`` -> ``decode`` ->
``.cpu()`` one batch after the other on one stream.  ``StreamedReconstructor`` keeps the same per-batch call but puts the
host->device copy of the NEXT batch and the device->host copy of the PREVIOUS result on two side streams (B200 has
separate copy engines), so a steady-state step costs the kernels only:

    pipe = StreamedReconstructor(model)
    for i, x in enumerate(batches):                      # pinned fp32 [B, 3, H, W] host tensors
        pipe.reconstruct(x, outs[i], next_x_host=batches[i + 1] if i + 1 < len(batches) else None)
    pipe.synchronize()                                   # outs[*] (pinned) are complete

Everything is ordered with CUDA events; the caching allocator is told about the cross-stream uses (``record_stream``).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

Tensor = torch.Tensor


class StreamedReconstructor:
    def __init__(self, model: torch.nn.Module, device: Optional[torch.device] = None):
        self.model = model
        self.device = torch.device(device) if device is not None else next(model.parameters()).device
        if self.device.type != "cuda":
            raise RuntimeError("StreamedReconstructor needs a CUDA model (there is no CPU fallback)")
        self.h2d = torch.cuda.Stream(self.device)
        self.d2h = torch.cuda.Stream(self.device)
        self._staged: Optional[Tuple[Tensor, Tensor, torch.cuda.Event]] = None    # (host source, device copy, ready)

    def _stage(self, x_host: Tensor) -> None:
        if not x_host.is_pinned():
            raise ValueError("host batches must be pinned (torch.Tensor.pin_memory) for asynchronous copies")
        with torch.cuda.stream(self.h2d):
            xd = x_host.to(self.device, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.h2d)
        self._staged = (x_host, xd, ev)

    @torch.no_grad()
    def reconstruct(self, x_host: Tensor, out_host: Tensor, next_x_host: Optional[Tensor] = None,
                    sample: bool = False) -> None:
        """decode(encode(x_host).mu) into the pinned ``out_host`` (asynchronously: call ``synchronize`` before reading
        it).  ``next_x_host``: the batch of the following call, whose upload starts now.  ``sample=True`` decodes a
        sample z = mu + eps * sigma instead of mu (TransVAE.forward semantics)."""
        cur = torch.cuda.current_stream(self.device)
        if self._staged is None or self._staged[0] is not x_host:
            self._stage(x_host)                      # cold start (or an unannounced batch): upload now
        _, xd, ready = self._staged
        self._staged = None
        if next_x_host is not None:
            self._stage(next_x_host)                 # runs on the copy engine while the kernels below execute
        cur.wait_event(ready)
        if sample:
            rec = self.model(xd)[0]
        else:
            mu, _ = self.model.encode(xd)
            rec = self.model.decode(mu)
        xd.record_stream(cur)
        done = torch.cuda.Event()
        done.record(cur)
        if not out_host.is_pinned():
            raise ValueError("out_host must be pinned")
        self.d2h.wait_event(done)
        with torch.cuda.stream(self.d2h):
            out_host.copy_(rec, non_blocking=True)
        rec.record_stream(self.d2h)

    def join(self) -> None:
        """Make the current stream wait for the outstanding copies (for device-side timing)."""
        cur = torch.cuda.current_stream(self.device)
        cur.wait_stream(self.d2h)
        cur.wait_stream(self.h2d)

    def synchronize(self) -> None:
        self.d2h.synchronize()
        self.h2d.synchronize()
