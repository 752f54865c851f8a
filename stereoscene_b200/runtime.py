"""Serving-style runtime around the volumetric forward: static device buffers, the whole forward
captured in ONE CUDA graph (~300 kernel launches per stereo pair collapse into a single graph launch),
pinned host staging and two copy streams so the host->device transfer of pair i+1 and the device->host
read of pair i-1 overlap the compute of pair i.

    eng = VolumetricEngine(model, left_calib, right_calib, calib, occ_size)
    labels = eng.infer(x_left_host, x_right_host)            # uint8 [B, X, Y, Z] on the host
    for labels in eng.stream(pairs): ...                      # pipelined

The reference has no counterpart (it launches eagerly from Python, detectors/bevdepth_occupancy.py:83-128);
this is the B200-side replacement for that orchestration on the inference path.
"""
from __future__ import annotations

from typing import Iterable, Iterator, Tuple

import torch


class VolumetricEngine:
    def __init__(self, model, left: dict, right: dict, calib: torch.Tensor, occ_size, feature_shape: Tuple[int, ...],
                 device=None, warmup: int = 2):
        self.model = model
        self.device = device or next(model.parameters()).device
        self.occ_size = tuple(int(s) for s in occ_size)
        dev = self.device
        self.left = {k: v.to(dev) for k, v in left.items()}
        self.right = {k: v.to(dev) for k, v in right.items()}
        self.calib = calib.to(dev)
        B = feature_shape[0]
        # two input / output slots: slot i%2 is being filled while slot (i-1)%2 is being consumed
        self.xl = [torch.empty(feature_shape, dtype=torch.float32, device=dev) for _ in range(2)]
        self.xr = [torch.empty(feature_shape, dtype=torch.float32, device=dev) for _ in range(2)]
        self.labels_host = [torch.empty((B, *self.occ_size), dtype=torch.uint8).pin_memory() for _ in range(2)]
        # separate streams for the two copy directions: a download waits for ITS compute, and on a shared stream
        # that wait would also hold back the next pair's upload (which only needs the slot's previous compute)
        self.copy_stream = torch.cuda.Stream(device=dev)          # host -> device
        self.d2h_stream = torch.cuda.Stream(device=dev)           # device -> host
        self.compute_stream = torch.cuda.Stream(device=dev)
        self.h2d_done = [torch.cuda.Event() for _ in range(2)]
        self.compute_done = [torch.cuda.Event() for _ in range(2)]
        self.d2h_done = [torch.cuda.Event() for _ in range(2)]
        self.graphs, self.outs = [], []
        with torch.no_grad():
            for s in range(2):
                self.xl[s].normal_(); self.xr[s].normal_()
            torch.cuda.synchronize(dev)
            with torch.cuda.stream(self.compute_stream):
                for _ in range(max(1, warmup)):                       # fills the packed-weight / splat-index caches
                    self._forward(0)
            torch.cuda.synchronize(dev)
            for s in range(2):
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, stream=self.compute_stream):
                    out = self._forward(s)
                self.graphs.append(g)
                self.outs.append(out)
            # The graphs bake in the device addresses of everything the forward read from the op / calibration caches
            # (packed weights, BatchNorm affines, disparity taps, the splat index).  Those caches are shared by every
            # engine / eager call on the process and evict; the engine therefore OWNS references to what its graphs
            # read, so an eviction can never hand that memory to another tensor while a graph still points at it.
            from . import ops
            self._graph_reads = ops.cached_state() + model.img_view_transformer.cached_state() + \
                [getattr(m, a) for m in model.modules() for a in ("_split", "_affine", "_gconvs") if getattr(m, a, None) is not None]
        torch.cuda.synchronize(dev)

    def _forward(self, slot: int):
        return self.model.forward_features(self.xl[slot], self.xr[slot], self.left, self.right, self.calib,
                                           occ_size=self.occ_size, want_labels=True)

    # ---- pipelined pieces (all asynchronous) ----------------------------------------------------------
    def _upload(self, slot: int, xl_host: torch.Tensor, xr_host: torch.Tensor):
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(self.compute_done[slot])      # slot's previous graph run has consumed its inputs
            self.xl[slot].copy_(xl_host, non_blocking=True)
            self.xr[slot].copy_(xr_host, non_blocking=True)
            self.h2d_done[slot].record(self.copy_stream)

    def _compute(self, slot: int):
        with torch.cuda.stream(self.compute_stream):
            self.compute_stream.wait_event(self.h2d_done[slot])
            self.compute_stream.wait_event(self.d2h_done[slot])       # previous labels of this slot have left the device
            self.graphs[slot].replay()
            self.compute_done[slot].record(self.compute_stream)

    def _download(self, slot: int):
        with torch.cuda.stream(self.d2h_stream):
            self.d2h_stream.wait_event(self.compute_done[slot])
            self.labels_host[slot].copy_(self.outs[slot]["labels"], non_blocking=True)
            self.d2h_done[slot].record(self.d2h_stream)

    # ---- public API --------------------------------------------------------------------------------------
    def infer(self, xl_host: torch.Tensor, xr_host: torch.Tensor) -> torch.Tensor:
        """One stereo pair (batch) from host memory to host labels, synchronous."""
        self._upload(0, xl_host, xr_host)
        self._compute(0)
        self._download(0)
        self.d2h_done[0].synchronize()
        return self.labels_host[0]

    def logits(self, slot: int = 0) -> torch.Tensor:
        """Device logits [B, classes, X, Y, Z] of the last run in `slot` (valid until the slot is reused)."""
        return self.outs[slot]["output_voxels"]

    def stream(self, pairs: Iterable[Tuple[torch.Tensor, torch.Tensor]]) -> Iterator[torch.Tensor]:
        """Pipelined inference: yields the host label volume of every pair, in order.  The yielded tensor is one of two
        reused pinned buffers and is valid only until the NEXT ``next()`` on the iterator (advancing enqueues the download
        of pair i+2 into the slot of pair i): consume or copy it before advancing."""
        pending = []
        for i, (xl_host, xr_host) in enumerate(pairs):
            slot = i & 1
            self._upload(slot, xl_host, xr_host)
            self._compute(slot)
            self._download(slot)
            pending.append(slot)
            if len(pending) == 2:
                done = pending.pop(0)
                self.d2h_done[done].synchronize()
                yield self.labels_host[done]
        for done in pending:
            self.d2h_done[done].synchronize()
            yield self.labels_host[done]
