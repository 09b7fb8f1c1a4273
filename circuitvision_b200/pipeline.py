"""Batched host-facing pipeline: pinned host crops -> SAM 2.1 wire masks -> node analysis -> host results.

This is the call a batch user makes (bench.py's `e2e` number goes through it).  The two hot calls of the reference's
pipeline (`/root/reference/src/analysis_pipeline.py:206` segment_with_sam2, `:234` get_node_connections) are chained on
the device: the uint8 mask never visits the host between them.  Host<->device traffic is software-pipelined over four
CUDA streams with `depth` slots in flight:

    copy-in stream : H2D of crop batch i+1         (pinned host memory -> slot buffer)
    compute stream : cv_sam2_forward of batch i
    nodes stream   : cv_nodes_analyze of batch i-1 (small grids, latency-bound border tracers: hides under the forward)
    copy-out stream: D2H of batch i-1's node tables, emptied masks and enhanced images into pinned host buffers

Nothing here computes on the CPU; without libcv_b200.so and an sm_100 device construction raises CvError.
"""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import CONTOUR_DTYPE, PAIR_DTYPE, RESULT_DTYPE, CvError
from .nodes import NodeAnalyzer, NodeBatchResult


class _Slot:
    def __init__(self, dev, B, S, caps, points_prefix, want_images):
        self.na = NodeAnalyzer(dev, caps)
        self.d_rgb = torch.empty((B, S, S, 3), dtype=torch.uint8, device=dev)
        c = self.na.caps
        pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
        self.h_results = pin((B, RESULT_DTYPE.itemsize), torch.uint8)
        self.h_contours = pin((B, c["max_contours"], CONTOUR_DTYPE.itemsize), torch.uint8)
        self.h_pairs = pin((B, c["max_pairs"], PAIR_DTYPE.itemsize), torch.uint8)
        self.h_points = pin((B, points_prefix, 2), torch.int32)
        self.h_extents = pin((B, 4), torch.int32)
        self.h_masks = pin((B, S, S), torch.uint8) if want_images else None
        self.h_emptied = pin((B, S, S), torch.uint8) if want_images else None
        self.h_enhanced = None  # allocated on first use (width depends on the aspect ratio)
        self.ev_in, self.ev_mask, self.ev_done, self.ev_out = (torch.cuda.Event() for _ in range(4))
        self.busy = False
        self.result = None
        self.rboxes = None
        self.launches = 0


class BatchResult:
    """Host-side results of one submitted batch."""

    def __init__(self, nodes: NodeBatchResult, masks, emptied, enhanced, extents):
        self.nodes_table = nodes          # NodeBatchResult with host tables: .nodes(b) -> reference's new_nodes_list
        self.masks = masks                # [B,S,S] uint8 numpy (pinned view) or None
        self.emptied = emptied            # [B,S,S] uint8 or None
        self.enhanced = enhanced          # [B,600,w'] uint8 or None
        self.extents = extents            # [B,4] int32: min x, min y, max x, max y of each mask's foreground

    def nodes(self, b: int):
        return self.nodes_table.nodes(b)


class CropPipeline:
    """`submit()` enqueues a batch (non-blocking), `collect()` returns the oldest batch's host results.

    model: `sam2_infer.SAM2ImageWrapper` on a CUDA device.  Crops are uint8 [B,1024,1024,3] in PINNED host memory
    (channel order as the reference's pipeline passes it to segment_with_sam2, i.e. the :343 swap is applied)."""

    def __init__(self, model, batch: int, depth: int = 2, caps: Optional[dict] = None, points_prefix: int = 16384,
                 want_images: bool = True):
        eng = model.engine()
        self.model, self.dev = model, torch.device("cuda", eng.dev)
        _lib.require_device(eng.dev)
        self.B, self.S = int(batch), 1024
        model.set_max_batch(max(model.max_batch, self.B))
        self.points_prefix, self.want_images = points_prefix, want_images
        with torch.cuda.device(self.dev):
            self.s_in, self.s_compute, self.s_nodes, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(4))
            self.slots = [_Slot(self.dev, self.B, self.S, caps, points_prefix, want_images) for _ in range(depth)]
        self._next, self._oldest, self._inflight = 0, 0, 0
        self.h2d_bytes = self.d2h_bytes = 0

    def submit(self, host_crops: torch.Tensor, boxes_list: List[list]):
        if self._inflight == len(self.slots):
            raise CvError("pipeline full: collect() a batch before submitting another")
        if host_crops.dtype != torch.uint8 or tuple(host_crops.shape) != (self.B, self.S, self.S, 3):
            raise CvError(f"crops must be uint8 [{self.B},1024,1024,3]")
        if not host_crops.is_pinned():
            raise CvError("crops must live in pinned host memory (torch.Tensor.pin_memory())")
        s = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        self._inflight += 1
        with torch.cuda.device(self.dev):
            with torch.cuda.stream(self.s_in):
                s.d_rgb.copy_(host_crops, non_blocking=True)
                s.ev_in.record(self.s_in)
            with torch.cuda.stream(self.s_compute):
                self.s_compute.wait_event(s.ev_in)
                masks = self.model.segment_batch_u8(s.d_rgb)
                ext = self.model.last_extents
                s.ev_mask.record(self.s_compute)
            with torch.cuda.stream(self.s_nodes):
                d_rec, d_off, rboxes, max_per = s.na.upload_boxes(boxes_list, self.S, self.S)
                self.s_nodes.wait_event(s.ev_mask)
                r = s.na.run(masks, d_rec, d_off, max_per, rboxes)
                masks.record_stream(self.s_nodes)
                s.launches = self.model.last_launches + r.launches
                s.ev_done.record(self.s_nodes)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(s.ev_done)
                s.h_results.copy_(r.results, non_blocking=True)
                s.h_contours.copy_(r.contours, non_blocking=True)
                s.h_pairs.copy_(r.pairs, non_blocking=True)
                s.h_points.copy_(r.points[:, :self.points_prefix], non_blocking=True)
                s.h_extents.copy_(ext, non_blocking=True)
                nbytes = s.h_results.numel() + s.h_contours.numel() + s.h_pairs.numel() + s.h_points.numel() * 4 + 16 * self.B
                if self.want_images:
                    if s.h_enhanced is None or s.h_enhanced.shape != r.enhanced.shape:
                        s.h_enhanced = torch.empty(tuple(r.enhanced.shape), dtype=torch.uint8, pin_memory=True)
                    s.h_masks.copy_(masks, non_blocking=True)
                    s.h_emptied.copy_(r.emptied, non_blocking=True)
                    s.h_enhanced.copy_(r.enhanced, non_blocking=True)
                    nbytes += s.h_masks.numel() + s.h_emptied.numel() + s.h_enhanced.numel()
                # keep `masks` alive until the copy-out stream has read it
                masks.record_stream(self.s_out)
                ext.record_stream(self.s_out)
                s.ev_out.record(self.s_out)
            s.result, s.rboxes = r, rboxes
        self.h2d_bytes = host_crops.numel() + d_rec.numel() + d_off.numel() * 4
        self.d2h_bytes = nbytes
        self.last_launches = s.launches

    def collect(self) -> BatchResult:
        if self._inflight == 0:
            raise CvError("nothing in flight")
        s = self.slots[self._oldest]
        self._oldest = (self._oldest + 1) % len(self.slots)
        self._inflight -= 1
        s.ev_out.synchronize()
        r = s.result
        res = s.h_results.numpy().view(RESULT_DTYPE).reshape(self.B)
        npts = int(res["n_points"].max()) if self.B else 0
        if npts > self.points_prefix:  # rare: fetch the long point pools synchronously
            pts = r.points[:, :npts].cpu().numpy()
        else:
            pts = s.h_points.numpy()
        # views into the slot's pinned buffers: valid until this slot is reused (`depth` submits later)
        host = dict(results=res, contours=s.h_contours.numpy().view(CONTOUR_DTYPE).reshape(self.B, -1),
                    pairs=s.h_pairs.numpy().view(PAIR_DTYPE).reshape(self.B, -1), points=pts)
        nb = NodeBatchResult(r.B, r.H, r.W, r.new_w, r.caps, None, None, None, None, None, None, None, s.rboxes, r.launches)
        nb._host = host
        img = (lambda t: None if t is None else t.numpy())
        return BatchResult(nb, img(s.h_masks), img(s.h_emptied), img(s.h_enhanced), s.h_extents.numpy().copy())

    def run(self, host_crops: torch.Tensor, boxes_list: List[list]) -> BatchResult:
        """Synchronous convenience: one batch in, its results out."""
        self.submit(host_crops, boxes_list)
        return self.collect()
