"""Batched host-facing pipelines: pinned host inputs -> device hot path -> host results, software-pipelined over CUDA streams.

`CropPipeline` is the call a batch user makes for the whole path (bench.py's `e2e` number goes through it): the two hot
calls of the reference's pipeline (`/root/reference/src/analysis_pipeline.py:206` segment_with_sam2, `:234`
get_node_connections) are chained on the device, the uint8 mask never visits the host between them.

    copy-in stream : H2D of crop batch i+1         (pinned host memory -> slot buffer)
    compute stream : cv_sam2_forward of batch i
    nodes stream   : cv_nodes_analyze of batch i-1 (small grids, latency-bound border tracers: hides under the forward)
    copy-out stream: D2H of batch i-1's results into pinned host buffers

`MaskPipeline` is the same pattern for the nodes-only path (BASELINE cfg 4: wire masks in, node tables + emptied masks out).

Result tables are fixed-capacity on the device (mostly empty); `cv_nodes_pack` compacts their used prefixes into one blob
and only that blob (plus a 32-byte header per image) crosses PCIe.

Nothing here computes on the CPU; without libcv_b200.so and an sm_100 device construction raises CvError.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional

import numpy as np
import torch

from . import _lib
from ._lib import CONTOUR_DTYPE, PAIR_DTYPE, RESULT_DTYPE, CvError, cv_nodes_caps
from .nodes import RESIZED_HEIGHT, NodeAnalyzer, NodeBatchResult, resized_width

_BLOB_BYTES_PER_IMAGE = 192 * 1024  # pinned staging for the compacted tables; larger batches fall back to the dense copy


class _TableStager:
    """Device + pinned-host staging of one batch's compacted node tables (cv_nodes_pack)."""

    def __init__(self, dev, B, caps):
        self.B, self.dev = B, dev
        self.caps = cv_nodes_caps(**caps)
        self.cap = B * _BLOB_BYTES_PER_IMAGE
        self.d_header = torch.empty((B + 1, 4), dtype=torch.int64, device=dev)
        self.d_blob = torch.empty(self.cap, dtype=torch.uint8, device=dev)
        self.h_header = torch.empty((B + 1, 4), dtype=torch.int64, pin_memory=True)
        self.h_blob = torch.empty(self.cap, dtype=torch.uint8, pin_memory=True)
        self.h_results = torch.empty((B, RESULT_DTYPE.itemsize), dtype=torch.uint8, pin_memory=True)
        self.ev_header = torch.cuda.Event()
        # the blob is fetched on a stream of its own: the copy-out stream already holds the NEXT batch's image copies, which
        # wait for that batch's compute — synchronising on it would serialise the pipeline
        self.s_fetch = torch.cuda.Stream(dev)
        self.lib = _lib.load()

    def pack(self, r: NodeBatchResult, stream):
        """On `stream` (after the analysis): compact the tables, start the D2H of the header and the result rows."""
        rc = self.lib.cv_nodes_pack(r.contours.data_ptr(), r.points.data_ptr(), r.pairs.data_ptr(), r.results.data_ptr(),
                                    self.B, C.byref(self.caps), self.d_header.data_ptr(), self.d_blob.data_ptr(), self.cap,
                                    stream.cuda_stream)
        _lib.check(rc, "cv_nodes_pack")
        self.h_header.copy_(self.d_header, non_blocking=True)
        self.h_results.copy_(r.results, non_blocking=True)
        self.ev_header.record(stream)

    def fetch(self, r: NodeBatchResult, stream=None):
        """Host side: wait for the header, copy exactly the used bytes of the blob; returns the host-table dict."""
        stream = self.s_fetch
        self.ev_header.synchronize()
        total = int(self.h_header[self.B, 0])
        res = self.h_results.numpy().view(RESULT_DTYPE).reshape(self.B)
        if total > self.cap:  # rare: dense copy of the fixed-capacity tables
            r._host = None
            host = r.tables_to_host()
            return host, sum(v.nbytes for v in host.values())
        with torch.cuda.stream(stream):
            if total:
                self.h_blob[:total].copy_(self.d_blob[:total], non_blocking=True)
            stream.synchronize()
        host = dict(results=res, header=self.h_header.numpy(), blob=self.h_blob.numpy())
        return host, total + self.h_header.numel() * 8 + self.h_results.numel()


class _Slot:
    def __init__(self, dev, B, S, caps, want_images):
        self.na = NodeAnalyzer(dev, caps)
        self.d_rgb = torch.empty((B, S, S, 3), dtype=torch.uint8, device=dev)
        pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
        self.tables = _TableStager(dev, B, self.na.caps)
        self.h_extents = pin((B, 4), torch.int32)
        self.h_boxes = (pin((B * 256, 48), torch.uint8), pin((B + 1,), torch.int32))  # async box upload staging
        self.h_masks = pin((B, S, S), torch.uint8) if want_images else None
        self.h_emptied = pin((B, S, S), torch.uint8) if want_images else None
        self.h_enhanced = None  # allocated on first use (width depends on the aspect ratio)
        self.ev_in, self.ev_mask, self.ev_done, self.ev_out = (torch.cuda.Event() for _ in range(4))
        self.result = None
        self.rboxes = None
        self.launches = 0
        self.image_bytes = 0


class BatchResult:
    """Host-side results of one submitted batch."""

    def __init__(self, nodes: NodeBatchResult, masks, emptied, enhanced, extents):
        self.nodes_table = nodes          # NodeBatchResult with host tables: .nodes(b) -> reference's new_nodes_list
        self.masks = masks                # [B,S,S] uint8 numpy (pinned view) or None
        self.emptied = emptied            # [B,S,S] uint8 or None
        self.enhanced = enhanced          # [B,600,w'] uint8 or None
        self.extents = extents            # [B,4] int32: min x, min y, max x, max y of each mask's foreground

    @property
    def n_nodes(self):
        return self.nodes_table.tables_to_host()["results"]["n_nodes"]

    def nodes(self, b: int):
        return self.nodes_table.nodes(b)


class CropPipeline:
    """`submit()` enqueues a batch (non-blocking), `collect()` returns the oldest batch's host results.

    model: `sam2_infer.SAM2ImageWrapper` on a CUDA device.  Crops are uint8 [B,1024,1024,3] in PINNED host memory
    (channel order as the reference's pipeline passes it to segment_with_sam2, i.e. the :343 swap is applied).

    The image arrays of a BatchResult (`masks`, `emptied`, `enhanced`) and its node tables are views into the slot's pinned
    buffers: they stay valid until `depth` further batches have been submitted; copy what must outlive that."""

    def __init__(self, model, batch: int, depth: int = 2, caps: Optional[dict] = None, want_images: bool = True):
        eng = model.engine()
        self.model, self.dev = model, torch.device("cuda", eng.dev)
        _lib.require_device(eng.dev)
        self.B, self.S = int(batch), 1024
        model.set_max_batch(max(model.max_batch, min(self.B, 64)))
        self.want_images = want_images
        with torch.cuda.device(self.dev):
            self.s_in, self.s_compute, self.s_nodes, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(4))
            self.slots = [_Slot(self.dev, self.B, self.S, caps, want_images) for _ in range(depth)]
        self._next, self._oldest, self._inflight = 0, 0, 0
        self.h2d_bytes = self.d2h_bytes = 0
        self.last_launches = 0

    @property
    def depth(self):
        return len(self.slots)

    @property
    def inflight(self):
        return self._inflight

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
                d_rec, d_off, rboxes, max_per = s.na.upload_boxes(boxes_list, self.S, self.S, pinned=s.h_boxes)
                s.ev_in.record(self.s_in)
            with torch.cuda.stream(self.s_compute):
                self.s_compute.wait_event(s.ev_in)
                masks = self.model.segment_batch_u8(s.d_rgb)
                ext = self.model.last_extents
                s.ev_mask.record(self.s_compute)
            with torch.cuda.stream(self.s_nodes):
                self.s_nodes.wait_event(s.ev_mask)  # (follows ev_in on the compute stream: the boxes have landed too)
                r = s.na.run(masks, d_rec, d_off, max_per, rboxes)
                masks.record_stream(self.s_nodes)
                d_rec.record_stream(self.s_nodes)
                d_off.record_stream(self.s_nodes)
                s.tables.pack(r, self.s_nodes)
                s.launches = self.model.last_launches + r.launches + 2
                s.ev_done.record(self.s_nodes)
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(s.ev_done)
                s.h_extents.copy_(ext, non_blocking=True)
                nbytes = 16 * self.B
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
            s.result, s.rboxes, s.image_bytes = r, rboxes, nbytes
        self.h2d_bytes = host_crops.numel() + d_rec.numel() + d_off.numel() * 4
        self.last_launches = s.launches

    def collect(self) -> BatchResult:
        if self._inflight == 0:
            raise CvError("nothing in flight")
        s = self.slots[self._oldest]
        self._oldest = (self._oldest + 1) % len(self.slots)
        self._inflight -= 1
        r = s.result
        with torch.cuda.device(self.dev):
            host, table_bytes = s.tables.fetch(r, self.s_out)
        s.ev_out.synchronize()
        self.d2h_bytes = s.image_bytes + table_bytes
        nb = NodeBatchResult(r.B, r.H, r.W, r.new_w, r.caps, None, None, None, None, None, None, None, s.rboxes, r.launches)
        nb._host = host
        img = (lambda t: None if t is None else t.numpy())
        return BatchResult(nb, img(s.h_masks), img(s.h_emptied), img(s.h_enhanced), s.h_extents.numpy().copy())

    def run(self, host_crops: torch.Tensor, boxes_list: List[list]) -> BatchResult:
        """Synchronous convenience: one batch in, its results out."""
        self.submit(host_crops, boxes_list)
        return self.collect()


class _MaskSlot:
    def __init__(self, dev, B, H, W, caps, want_images):
        self.na = NodeAnalyzer(dev, caps)
        self.d_masks = torch.empty((B, H, W), dtype=torch.uint8, device=dev)
        self.tables = _TableStager(dev, B, self.na.caps)
        pin = lambda shape, dt: torch.empty(shape, dtype=dt, pin_memory=True)
        self.h_boxes = (pin((B * 1024, 48), torch.uint8), pin((B + 1,), torch.int32))  # async box upload staging
        self.h_emptied = pin((B, H, W), torch.uint8) if want_images else None
        self.h_enhanced = pin((B, RESIZED_HEIGHT, resized_width(H, W)), torch.uint8) if want_images else None
        self.ev_in, self.ev_done, self.ev_out = (torch.cuda.Event() for _ in range(3))
        self.result = None
        self.rboxes = None
        self.image_bytes = 0


class MaskPipeline:
    """Nodes-only batch API (BASELINE cfg 4): pinned host wire masks [B,H,W] uint8 + box lists in, per-image node tables
    (+ the emptied masks and enhanced images `get_node_connections` returns) out.  H2D of batch i+1, analysis of batch i and
    D2H of batch i-1 run on three streams; the tables cross PCIe compacted (cv_nodes_pack)."""

    def __init__(self, device: int, batch: int, H: int, W: int, depth: int = 2, caps: Optional[dict] = None,
                 want_images: bool = True):
        self.dev = torch.device("cuda", device)
        _lib.require_device(device)
        self.B, self.H, self.W = int(batch), int(H), int(W)
        self.want_images = want_images
        with torch.cuda.device(self.dev):
            self.s_in, self.s_compute, self.s_out = (torch.cuda.Stream(self.dev) for _ in range(3))
            self.slots = [_MaskSlot(self.dev, self.B, H, W, caps, want_images) for _ in range(depth)]
        self._next, self._oldest, self._inflight = 0, 0, 0
        self.h2d_bytes = self.d2h_bytes = 0
        self.last_launches = 0

    @property
    def depth(self):
        return len(self.slots)

    @property
    def inflight(self):
        return self._inflight

    def submit(self, host_masks: torch.Tensor, boxes_list: List[list]):
        if self._inflight == len(self.slots):
            raise CvError("pipeline full: collect() a batch before submitting another")
        if host_masks.dtype != torch.uint8 or tuple(host_masks.shape) != (self.B, self.H, self.W):
            raise CvError(f"masks must be uint8 [{self.B},{self.H},{self.W}]")
        if not host_masks.is_pinned():
            raise CvError("masks must live in pinned host memory (torch.Tensor.pin_memory())")
        s = self.slots[self._next]
        self._next = (self._next + 1) % len(self.slots)
        self._inflight += 1
        with torch.cuda.device(self.dev):
            with torch.cuda.stream(self.s_in):
                s.d_masks.copy_(host_masks, non_blocking=True)
                d_rec, d_off, rboxes, max_per = s.na.upload_boxes(boxes_list, self.H, self.W, pinned=s.h_boxes)
                s.ev_in.record(self.s_in)
            with torch.cuda.stream(self.s_compute):
                self.s_compute.wait_event(s.ev_in)
                r = s.na.run(s.d_masks, d_rec, d_off, max_per, rboxes)
                d_rec.record_stream(self.s_compute)
                d_off.record_stream(self.s_compute)
                s.tables.pack(r, self.s_compute)
                s.ev_done.record(self.s_compute)
            nbytes = 0
            with torch.cuda.stream(self.s_out):
                self.s_out.wait_event(s.ev_done)
                if self.want_images:
                    s.h_emptied.copy_(r.emptied, non_blocking=True)
                    s.h_enhanced.copy_(r.enhanced, non_blocking=True)
                    nbytes = s.h_emptied.numel() + s.h_enhanced.numel()
                s.ev_out.record(self.s_out)
            s.result, s.rboxes, s.image_bytes = r, rboxes, nbytes
        self.h2d_bytes = host_masks.numel() + d_rec.numel() + d_off.numel() * 4
        self.last_launches = r.launches + 2

    def collect(self) -> BatchResult:
        if self._inflight == 0:
            raise CvError("nothing in flight")
        s = self.slots[self._oldest]
        self._oldest = (self._oldest + 1) % len(self.slots)
        self._inflight -= 1
        r = s.result
        with torch.cuda.device(self.dev):
            host, table_bytes = s.tables.fetch(r, self.s_out)
        s.ev_out.synchronize()
        self.d2h_bytes = s.image_bytes + table_bytes
        nb = NodeBatchResult(r.B, r.H, r.W, r.new_w, r.caps, None, None, None, None, None, None, None, s.rboxes, r.launches)
        nb._host = host
        img = (lambda t: None if t is None else t.numpy())
        return BatchResult(nb, None, img(s.h_emptied), img(s.h_enhanced), None)

    def run(self, host_masks: torch.Tensor, boxes_list: List[list]) -> BatchResult:
        self.submit(host_masks, boxes_list)
        return self.collect()


class PagePipeline:
    """The reference's per-page sequence (`/root/reference/src/analysis_pipeline.py:177-246`: crop_image_and_adjust_bboxes ->
    segment_with_sam2 on the crop -> get_node_connections on the crop's mask) for a BATCH of whole pages of any size.

    Only the raw uint8 pages cross PCIe (<= 3 bytes per page pixel instead of 12 MB of fp32 per crop): the crop window comes
    from the host box geometry (crop.py, the reference's rule), the crop + ToTensor + antialiased Resize((1024,1024)) +
    Normalize run in one batched device kernel pair (cv_sam2_preprocess_pages), SAM 2.1 runs on the whole batch, the logits go
    back to each crop's own resolution (postprocess_masks), and the node analysis runs per crop (their sizes differ)."""

    def __init__(self, model, padding: int = 80):
        from . import sam2_infer as _s
        self._s = _s
        self.model = model
        self.eng = model.engine()
        self.dev = torch.device("cuda", self.eng.dev)
        self.padding = padding
        self.na = NodeAnalyzer(self.dev)
        self.lib = _lib.load()
        self.lib.cv_sam2_preprocess_pages.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                                      C.c_void_p]
        self.h2d_bytes = 0

    def run(self, pages: List[np.ndarray], boxes_list: List[list]):
        """pages: (H,W,3) uint8 arrays in the channel order the reference's pipeline passes to segment_with_sam2.
        -> list of dicts: mask [Hc,Wc] u8, extent, nodes, emptied, enhanced, boxes (moved into the crop), crop_info."""
        from . import crop as _crop
        from .nodes import NON_COMPONENTS
        B = len(pages)
        if B == 0 or B != len(boxes_list):
            raise CvError("one box list per page is required")
        geom = np.zeros(B, np.dtype([("off", "<i8"), ("pw", "<i4"), ("x0", "<i4"), ("y0", "<i4"), ("x1", "<i4"), ("y1", "<i4"),
                                     ("pad", "<i4")]))
        moved, infos, off = [], [], 0
        for b, (pg, boxes) in enumerate(zip(pages, boxes_list)):
            if pg.ndim != 3 or pg.shape[2] != 3 or pg.dtype != np.uint8:
                raise CvError("pages must be (H,W,3) uint8 arrays")
            _, mv, info = _crop.crop_image_and_adjust_bboxes(pg, boxes, self.padding, NON_COMPONENTS)  # host box geometry only
            win = info["final_crop_window_abs"] if info.get("crop_applied") else (0, 0, pg.shape[1], pg.shape[0])
            geom[b] = (off, pg.shape[1], win[0], win[1], win[2], win[3], 0)
            off += pg.size
            moved.append(mv)
            infos.append(info)
        host = torch.empty(off, dtype=torch.uint8, pin_memory=True)
        hv = host.numpy()
        for b, pg in enumerate(pages):
            hv[int(geom["off"][b]):int(geom["off"][b]) + pg.size] = np.ascontiguousarray(pg).reshape(-1)
        self.h2d_bytes = off + geom.nbytes
        hc = (geom["y1"] - geom["y0"]).astype(int)
        wc = (geom["x1"] - geom["x0"]).astype(int)
        max_hc = int(hc.max())
        out = []
        with torch.cuda.device(self.dev):
            st = torch.cuda.current_stream().cuda_stream
            d_pages = host.to(self.dev, non_blocking=True)
            d_geom = torch.from_numpy(geom.view(np.uint8).reshape(B, 32).copy()).to(self.dev)
            x = torch.empty((B, 3, 1024, 1024), dtype=torch.float32, device=self.dev)
            tmp = torch.empty((B, max_hc, 1024, 3), dtype=torch.float32, device=self.dev)
            _lib.check(self.lib.cv_sam2_preprocess_pages(d_pages.data_ptr(), d_geom.data_ptr(), B, max_hc, 1, tmp.data_ptr(),
                                                         x.data_ptr(), st), "cv_sam2_preprocess_pages")
            self.last_input = x
            r = self.eng.forward(x, 1, False, want_high=True, want_low=False)
            high = r["high"]
            libs = self._s._libsam()
            for b in range(B):
                H, W = int(hc[b]), int(wc[b])
                mask = torch.empty((H, W), dtype=torch.uint8, device=self.dev)
                ext = torch.empty((4,), dtype=torch.int32, device=self.dev)
                _lib.check(libs.cv_sam2_resize_logits(high[b].data_ptr(), 1, 1024, H, W, None, mask.data_ptr(), ext.data_ptr(), st),
                           "cv_sam2_resize_logits")
                nr = self.na.analyze(mask[None], [moved[b]])
                e = ext.cpu().tolist()
                out.append(dict(mask=mask.cpu().numpy(), extent=(e[0], e[1], e[2] + 1, e[3] + 1) if e[2] >= 0 else None,
                                nodes=nr.nodes(0), emptied=nr.emptied[0].cpu().numpy(), enhanced=nr.enhanced[0].cpu().numpy(),
                                boxes=moved[b], crop_info=infos[b]))
        return out
