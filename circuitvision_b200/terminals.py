"""Host side of the preliminary terminal reclassification.

Mirrors `/root/reference/src/circuit_analyzer.py:2217-2310` (`reclassify_terminals_based_on_connectivity`, called from
`src/analysis_pipeline.py:127` on the full RGB page with the NMS'd YOLO boxes): the page is segmented with the
reference's grey + 31x31 adaptive threshold (:313-319), every non-preserved box is zeroed (:2241-2249), the external
contours of that mask are extracted at native resolution with the 1e-4 area filter (:2252), and every 'terminal' box
that has a contour vertex "near" it (threshold 10, :2276) for two or more distinct contours becomes a 'voltage.dc'
(:2291-2307).  All pixel work runs in `cv_terminals_analyze` (include/cv_b200.h); this file packs the boxes and applies
the relabelling to the caller's dicts in place.  No CPU implementation exists behind it.
"""
from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib
from ._lib import (BOX_DTYPE, CONTOUR_DTYPE, RESULT_DTYPE, CV_BOX_IS_TERMINAL, CV_BOX_ZERO_IN_MASK, CvError,
                   cv_nodes_caps)
from .nodes import PRESERVE_IN_MASK

RECLASS_MIN_CONTOURS = 2  # :2291


def pack_page_boxes(boxes_list):
    """-> (cv_box records [sum n], offsets int32 [B+1], max boxes per page).  Page coordinates, int()-truncated (:2243)."""
    total = sum(len(b) for b in boxes_list)
    rec = np.zeros(max(total, 1), BOX_DTYPE)
    offs = np.zeros(len(boxes_list) + 1, np.int32)
    k = 0
    for bi, boxes in enumerate(boxes_list):
        for b in boxes:
            r = rec[k]
            # the near test (:811-846) compares against the dict values themselves; YOLO boxes are round()ed ints (:276-283)
            for key in ("xmin", "ymin", "xmax", "ymax"):
                if int(b[key]) != b[key]:
                    raise CvError("terminal reclassification expects integer page coordinates (circuit_analyzer.py:280-283)")
                r[key] = int(b[key])
            cls = b.get("class")
            r["flags"] = (CV_BOX_ZERO_IN_MASK if cls not in PRESERVE_IN_MASK else 0) | \
                         (CV_BOX_IS_TERMINAL if cls == "terminal" else 0)
            k += 1
        offs[bi + 1] = k
    return rec[:total], offs, max((len(b) for b in boxes_list), default=0)


class TerminalBatchResult:
    def __init__(self, B, H, W, caps, wire_mask, counts, offs, contours, points, results, launches):
        self.B, self.H, self.W, self.caps = B, H, W, caps
        self.wire_mask, self.d_counts, self.offsets = wire_mask, counts, offs
        self.contours, self.points, self.results = contours, points, results
        self.launches = launches
        self._host = None

    def to_host(self):
        if self._host is None:
            res = self.results.cpu().numpy().view(RESULT_DTYPE).reshape(self.B)
            self._host = dict(results=res, counts=self.d_counts.cpu().numpy())
        return self._host

    def status(self):
        return self.to_host()["results"]["status"]

    def counts(self, b: int):
        """Per box of page b: distinct contours in contact (terminals), -1 for every other class."""
        h = self.to_host()
        return h["counts"][int(self.offsets[b]):int(self.offsets[b + 1])]

    def n_contours(self, b: int) -> int:
        return int(self.to_host()["results"][b]["n_contours"])

    def page_contours(self, b: int):
        """The page's filtered external contours (get_contours' list, id order) as (N,1,2) int32 arrays."""
        res = self.to_host()["results"][b]
        nK = int(res["n_contours"])
        con = self.contours[b].cpu().numpy().view(CONTOUR_DTYPE).reshape(-1)[:nK]
        pts = self.points[b, :max(int(res["n_points"]), 1)].cpu().numpy()
        return [np.ascontiguousarray(pts[int(c["offset"]):int(c["offset"]) + int(c["nverts"])]).reshape(-1, 1, 2)
                for c in con]


class TerminalAnalyzer:
    """Owns the device workspace of `cv_terminals_analyze`; re-entrant per instance + stream."""

    def __init__(self, device: int | torch.device = 0, caps: dict | None = None):
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        _lib.require_device(self.device.index or 0)
        self.lib = _lib.load()
        # native-resolution pages carry far more vertices than the 600-row node masks
        self.caps = dict(_lib.DEFAULT_CAPS, max_points=1 << 20, max_external=1 << 17)
        if caps:
            self.caps.update(caps)
        self._bufs = {}
        self._ws = None

    def _buffers(self, B, H, W, n_boxes):
        key = (B, H, W, n_boxes, tuple(sorted(self.caps.items())))
        if self._bufs.get("key") != key:
            dev, c, u8 = self.device, self.caps, torch.uint8
            self._bufs = dict(
                key=key,
                wire_mask=torch.empty((B, H, W), dtype=u8, device=dev),
                counts=torch.empty((max(n_boxes, 1),), dtype=torch.int32, device=dev),
                contours=torch.empty((B, c["max_contours"], CONTOUR_DTYPE.itemsize), dtype=u8, device=dev),
                points=torch.empty((B, c["max_points"], 2), dtype=torch.int32, device=dev),
                results=torch.empty((B, RESULT_DTYPE.itemsize), dtype=u8, device=dev),
            )
            cc = cv_nodes_caps(**self.caps)
            self._ws = torch.empty(self.lib.cv_terminals_workspace_bytes(B, H, W, C.byref(cc)), dtype=u8, device=dev)
        return self._bufs

    def run(self, d_pages: torch.Tensor, rec, offs, max_per) -> TerminalBatchResult:
        """d_pages: contiguous uint8 [B,H,W,3] RGB on this device.  Launches on torch's current stream."""
        if d_pages.dtype != torch.uint8 or d_pages.dim() != 4 or d_pages.shape[3] != 3 or not d_pages.is_contiguous():
            raise CvError("pages must be a contiguous uint8 [B,H,W,3] device tensor")
        if d_pages.device != self.device:
            raise CvError("pages live on another device")
        B, H, W, _ = d_pages.shape
        n_boxes = len(rec)
        bufs = self._buffers(B, H, W, n_boxes)
        dev = self.device
        if n_boxes:
            d_rec = torch.from_numpy(rec.view(np.uint8).reshape(-1, BOX_DTYPE.itemsize).copy()).to(dev)
        else:
            d_rec = torch.zeros((1, BOX_DTYPE.itemsize), dtype=torch.uint8, device=dev)
        d_off = torch.from_numpy(offs).to(dev)
        cc = cv_nodes_caps(**self.caps)
        st = torch.cuda.current_stream(dev).cuda_stream
        rc = self.lib.cv_terminals_analyze(
            d_pages.data_ptr(), B, H, W, d_rec.data_ptr(), d_off.data_ptr(), int(max_per), int(n_boxes),
            bufs["wire_mask"].data_ptr(), bufs["counts"].data_ptr(), bufs["contours"].data_ptr(),
            bufs["points"].data_ptr(), bufs["results"].data_ptr(), C.byref(cc), self._ws.data_ptr(), self._ws.numel(), st)
        _lib.check(rc, "cv_terminals_analyze")
        return TerminalBatchResult(B, H, W, dict(self.caps), bufs["wire_mask"], bufs["counts"], offs, bufs["contours"],
                                   bufs["points"], bufs["results"], self.lib.cv_last_launch_count())

    def analyze(self, pages_rgb, boxes_list, grow: bool = True) -> TerminalBatchResult:
        """pages_rgb: [B,H,W,3] uint8 (numpy => copied to the device; cuda tensor => used in place)."""
        if isinstance(pages_rgb, np.ndarray):
            pages_rgb = torch.from_numpy(np.ascontiguousarray(pages_rgb)).to(self.device)
        if len(boxes_list) != pages_rgb.shape[0]:
            raise CvError("one box list per page is required")
        rec, offs, max_per = pack_page_boxes(boxes_list)
        with torch.cuda.device(self.device):
            for _ in range(6):
                r = self.run(pages_rgb, rec, offs, max_per)
                st = int(np.bitwise_or.reduce(r.status())) if r.B else 0
                if not st or not grow:
                    return r
                if st & 1:
                    self.caps["max_external"] *= 4
                if st & 2:
                    self.caps["max_contours"] *= 4
                if st & 4:
                    self.caps["max_points"] *= 4
            raise CvError("terminal analysis capacities could not be satisfied")


def apply_reclassification(boxes, counts, class_names=None):
    """:2258-2307 — relabel in place; `class_names` plays `self.yolo.model.names` ({numeric id: name})."""
    vdc_id = None
    for num_id, name in (class_names or {}).items():
        if name == "voltage.dc":
            vdc_id = num_id
            break
    for b, n in zip(boxes, counts):
        if b.get("class") == "terminal" and int(n) >= RECLASS_MIN_CONTOURS:
            b["original_yolo_class_if_reclassified"] = b["class"]
            b["class"] = "voltage.dc"
            if vdc_id is not None:
                b["_yolo_class_id_temp"] = vdc_id
            b["was_reclassified_from_terminal"] = True
