"""Host side of the batched node / connection analysis.

Packs the reference's bbox dicts (`/root/reference/src/circuit_analyzer.py:1286-1605`) into `cv_box` records,
runs `cv_nodes_analyze` (include/cv_b200.h) on a batch of masks that stay resident on the device, and rebuilds
the reference's Python return structures from the small device tables.  torch is used for device memory and
streams only.  There is no CPU implementation behind this module: without libcv_b200.so and an sm_100 device
every entry point raises `CvError`.
"""
from __future__ import annotations

import ctypes as C
from collections import defaultdict
from copy import deepcopy
from operator import itemgetter

import numpy as np
import torch

from . import _lib
from ._lib import (BOX_DTYPE, CONTOUR_DTYPE, PAIR_DTYPE, RESULT_DTYPE, CV_BOX_IS_COMPONENT, CV_BOX_IS_SOURCE,
                   CV_BOX_ZERO_IN_MASK, CvError, cv_nodes_caps)

# class tables of the reference (circuit_analyzer.py:51-52, :1326, :1407-1415)
NON_COMPONENTS = frozenset(["text", "junction", "crossover", "vss", "explanatory", "circuit"])
SOURCE_COMPONENTS = frozenset(["voltage.ac", "voltage.dc", "voltage.dependent", "current.dc", "current.dependent"])
PRESERVE_IN_MASK = frozenset(["crossover", "junction", "circuit", "vss"])
THRESH_8 = frozenset(["diode", "diode.light_emitting", "diode.zener", "transistor.bjt", "transistor.fet"])
RESIZED_HEIGHT = 600  # circuit_analyzer.py:1361 new_height


def resized_width(H: int, W: int) -> int:
    """int(600 * (W / H)) — circuit_analyzer.py:800-803."""
    return int(RESIZED_HEIGHT * (W / H))


def resize_bboxes(boxes, width_scale: float, height_scale: float):
    """circuit_analyzer.py:461-477 (host-side box geometry; microseconds)."""
    out = []
    for b in boxes:
        r = b.copy()
        r["xmin"] = int(b["xmin"] * width_scale)
        r["ymin"] = int(b["ymin"] * height_scale)
        r["xmax"] = int(b["xmax"] * width_scale)
        r["ymax"] = int(b["ymax"] * height_scale)
        out.append(r)
    return out


def _box_ref(b):
    """identity used by the reference's de-duplication (:1424-1436)."""
    u = b.get("persistent_uid")
    if u is None:
        u = (b["class"], b["xmin"], b["ymin"], b["xmax"], b["ymax"])
    return u


_CLASS_CACHE: dict = {}


def _class_flags(cls):
    """(cv_box flags, contact threshold) of a class name (:1326, :51-52, :1404-1415) — cached per distinct name."""
    v = _CLASS_CACHE.get(cls)
    if v is None:
        flags = 0
        if cls not in PRESERVE_IN_MASK:
            flags |= CV_BOX_ZERO_IN_MASK
        if cls not in NON_COMPONENTS:
            flags |= CV_BOX_IS_COMPONENT
        if cls in SOURCE_COMPONENTS:
            flags |= CV_BOX_IS_SOURCE
        v = _CLASS_CACHE[cls] = (flags, 20 if cls in SOURCE_COMPONENTS else (8 if cls in THRESH_8 else 6))
    return v


_ATOMIC = (str, int, float, bool, type(None), np.integer, np.floating)
_XYXY = itemgetter("xmin", "ymin", "xmax", "ymax")


class ResizedBoxes:
    """The reference's `processing_bboxes_resized` (:807, resize_bboxes :461-477) of one image, materialised lazily: the
    integer coordinates of every box are computed vectorised at pack time, the dict copy of a box is only built when a
    node actually references it (a handful per image)."""

    __slots__ = ("boxes", "xyxy", "_cache", "_flat")

    def __init__(self, boxes, xyxy):
        self.boxes, self.xyxy, self._cache, self._flat = boxes, xyxy, {}, {}

    def __len__(self):
        return len(self.boxes)

    def __getitem__(self, j):
        r = self._cache.get(j)
        if r is None:
            r = self.boxes[j].copy()
            r["xmin"], r["ymin"], r["xmax"], r["ymax"] = self.xyxy[j].tolist()
            self._cache[j] = r
            # deepcopy(bbox) of :1422 degenerates to dict.copy() for the flat scalar dicts YOLO post-processing builds
            self._flat[j] = all(isinstance(v, _ATOMIC) for v in r.values())
        return r

    def copy_of(self, j):
        """An independent copy of resized box j (the reference attaches deepcopy(bbox) per (node, uid), :1422)."""
        r = self[j]
        return r.copy() if self._flat[j] else deepcopy(r)

    def __iter__(self):
        return (self[j] for j in range(len(self.boxes)))


def pack_boxes(boxes_list, H: int, W: int):
    """-> (cv_box records [sum n], offsets int32 [B+1], per-image ResizedBoxes, max boxes per image).
    One NumPy pass over all boxes of the batch: `int(v)` (:1339-1340) and `int(v * scale)` (:466-469) are truncations
    toward zero of the same float64 products the reference forms."""
    new_w = resized_width(H, W)
    sx, sy = new_w / W, RESIZED_HEIGHT / H  # :807
    counts = [len(b) for b in boxes_list]
    total = sum(counts)
    offs = np.zeros(len(boxes_list) + 1, np.int32)
    np.cumsum(counts, out=offs[1:])
    rec = np.zeros(max(total, 1), BOX_DTYPE)
    resized_all = []
    if total:
        flat = [b for boxes in boxes_list for b in boxes]
        xy = np.array([_XYXY(b) for b in flat], dtype=np.float64).reshape(total, 4)
        ixy = np.trunc(xy).astype(np.int64)
        rxy = np.trunc(xy * np.array([sx, sy, sx, sy])).astype(np.int64)
        for k, f in enumerate(("xmin", "ymin", "xmax", "ymax")):
            rec[f][:total] = ixy[:, k]
            rec["r" + f][:total] = rxy[:, k]
        cache = _CLASS_CACHE
        try:
            ft = [cache[b["class"]] for b in flat]
        except KeyError:
            ft = [_class_flags(b["class"]) for b in flat]
        ft = np.array(ft, dtype=np.int32).reshape(total, 2)
        rec["flags"][:total], rec["thresh"][:total] = ft[:, 0], ft[:, 1]
        groups = []
        k = 0
        for boxes, n in zip(boxes_list, counts):
            r = rxy[k:k + n]
            first = {}
            sd = first.setdefault
            uids = [b.get("persistent_uid") for b in boxes]
            if None in uids:  # identity used by the reference's de-duplication (:1424-1436), on RESIZED coordinates
                uids = [u if u is not None else (b["class"],) + tuple(r[j].tolist()) for j, (u, b) in enumerate(zip(uids, boxes))]
            groups.extend([sd(u, j) for j, u in enumerate(uids)])
            resized_all.append(ResizedBoxes(boxes, r))
            k += n
        rec["uid_group"][:total] = groups
    else:
        resized_all = [ResizedBoxes(b, np.zeros((0, 4), np.int64)) for b in boxes_list]
    max_per = max(counts, default=0)
    return rec[:total] if total else rec[:0], offs, resized_all, max_per


class NodeBatchResult:
    """Device-resident outputs of one `cv_nodes_analyze` call plus lazy host views."""

    def __init__(self, B, H, W, new_w, caps, emptied, resized, enhanced, contours, points, pairs, results, rboxes,
                 launches):
        self.B, self.H, self.W, self.new_w, self.caps = B, H, W, new_w, caps
        self.emptied, self.resized, self.enhanced = emptied, resized, enhanced
        self.contours, self.points, self.pairs, self.results = contours, points, pairs, results
        self.resized_boxes = rboxes
        self.launches = launches
        self._host = None

    def tables_to_host(self):
        """Copies the small per-image tables (not the images) to the host; returns dict of numpy arrays."""
        if self._host is None:
            res = self.results.cpu().numpy().view(RESULT_DTYPE).reshape(self.B)
            con = self.contours.cpu().numpy().view(CONTOUR_DTYPE).reshape(self.B, self.caps["max_contours"])
            prs = self.pairs.cpu().numpy().view(PAIR_DTYPE).reshape(self.B, self.caps["max_pairs"])
            # only the used prefix of each point pool travels
            npts = int(res["n_points"].max()) if self.B else 0
            pts = self.points[:, :max(npts, 1)].cpu().numpy()
            self._host = dict(results=res, contours=con, pairs=prs, points=pts)
        return self._host

    def status(self):
        return self.tables_to_host()["results"]["status"]

    def _image_tables(self, b: int):
        """(result row, contours[nK], pairs[nP], points[nPts,2]) of image b from the dense tables or from the packed blob
        of cv_nodes_pack (pipeline.py)."""
        h = self.tables_to_host()
        res = h["results"][b]
        if "blob" in h:
            off, nK, nP, nPt = (int(v) for v in h["header"][b])
            blob = h["blob"]
            con = np.frombuffer(blob, CONTOUR_DTYPE, nK, off)
            off += nK * CONTOUR_DTYPE.itemsize
            prs = np.frombuffer(blob, PAIR_DTYPE, nP, off)
            off += nP * PAIR_DTYPE.itemsize
            pts = np.frombuffer(blob, np.int32, 2 * nPt, off).reshape(-1, 2)
            return res, con, prs, pts
        nK, nP = int(res["n_contours"]), int(res["n_pairs"])
        return res, h["contours"][b, :nK], h["pairs"][b, :nP], h["points"][b]

    def nodes(self, b: int):
        """new_nodes_list of image b in the reference's format (:1547-1568)."""
        res, con, prs, pts = self._image_tables(b)
        if res["status"]:
            raise CvError(f"node analysis of image {b} overflowed a capacity (status {int(res['status'])})")
        rb = self.resized_boxes[b]
        comps = defaultdict(list)
        cache, flat = rb._cache, rb._flat
        for ci, bi in zip(prs["contour"].tolist(), prs["box"].tolist()):
            r = cache.get(bi)
            if r is None:
                r = rb[bi]
            comps[ci].append(r.copy() if flat[bi] else deepcopy(r))  # :1422
        out = []
        new_id = con["new_id"]
        keep = np.nonzero(new_id >= 0)[0]
        order = keep[np.argsort(new_id[keep], kind="stable")]
        offs, nv = con["offset"][order].tolist(), con["nverts"][order].tolist()
        pts3 = pts.reshape(-1, 1, 2)
        for k, nid, o, n in zip(order.tolist(), new_id[order].tolist(), offs, nv):
            out.append({"id": nid, "components": comps.get(k, []), "contour": pts3[o:o + n].astype(np.int32, copy=True)})
        return out

    def connection_points(self, b: int):
        _, _, prs, _ = self._image_tables(b)
        return list(zip(prs["px"].tolist(), prs["py"].tolist()))

    def all_contours(self, b: int):
        """Every contour that passed the area filter (get_contours' list, :412), id order."""
        _, con, _, pts = self._image_tables(b)
        out = []
        for k in range(len(con)):
            c = con[k]
            poly = np.array(pts[int(c["offset"]):int(c["offset"]) + int(c["nverts"])], dtype=np.int32)
            out.append({"id": k, "contour": poly.reshape(-1, 1, 2),
                        "area": abs(int(c["a00"])) * 0.5 / (RESIZED_HEIGHT * self.new_w),
                        "rectangle": (int(c["xmin"]), int(c["ymin"]), int(c["xmax"] - c["xmin"] + 1),
                                      int(c["ymax"] - c["ymin"] + 1))})
        return out


class NodeAnalyzer:
    """Owns the device workspace for `cv_nodes_analyze`; re-entrant per instance + stream."""

    def __init__(self, device: int | torch.device = 0, caps: dict | None = None):
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        _lib.require_device(self.device.index or 0)
        self.lib = _lib.load()
        self.caps = dict(_lib.DEFAULT_CAPS)
        if caps:
            self.caps.update(caps)
        self._ws = None
        self._bufs = {}

    def _ccaps(self):
        return cv_nodes_caps(**self.caps)

    def _buffers(self, B, H, W, new_w):
        key = (B, H, W, tuple(sorted(self.caps.items())))
        if self._bufs.get("key") != key:
            dev, c = self.device, self.caps
            u8 = torch.uint8
            self._bufs = dict(
                key=key,
                emptied=torch.empty((B, H, W), dtype=u8, device=dev),
                resized=torch.empty((B, RESIZED_HEIGHT, new_w), dtype=u8, device=dev),
                enhanced=torch.empty((B, RESIZED_HEIGHT, new_w), dtype=u8, device=dev),
                contours=torch.empty((B, c["max_contours"], CONTOUR_DTYPE.itemsize), dtype=u8, device=dev),
                points=torch.empty((B, c["max_points"], 2), dtype=torch.int32, device=dev),
                pairs=torch.empty((B, c["max_pairs"], PAIR_DTYPE.itemsize), dtype=u8, device=dev),
                results=torch.empty((B, RESULT_DTYPE.itemsize), dtype=u8, device=dev),
            )
            cc = self._ccaps()
            need = self.lib.cv_nodes_workspace_bytes(B, H, W, C.byref(cc))
            self._ws = torch.empty(need, dtype=u8, device=dev)
        return self._bufs

    def upload_boxes(self, boxes_list, H, W, pinned=None):
        """Pack + upload the box records on torch's current stream.  `pinned` = (uint8 [cap,48], int32 [B+1]) pinned staging
        tensors makes the copy asynchronous (pipelines); without it the pageable copy blocks the host until it is done."""
        rec, offs, rboxes, max_per = pack_boxes(boxes_list, H, W)
        dev = self.device
        raw = rec.view(np.uint8).reshape(-1, BOX_DTYPE.itemsize)
        if pinned is not None and len(rec) <= pinned[0].shape[0] and len(offs) <= pinned[1].shape[0]:
            h_rec, h_off = pinned
            n = max(len(rec), 1)
            h_rec[:len(rec)].numpy()[...] = raw
            h_off[:len(offs)].numpy()[...] = offs
            d_rec = torch.empty((n, BOX_DTYPE.itemsize), dtype=torch.uint8, device=dev)
            d_rec.copy_(h_rec[:n], non_blocking=True)
            d_off = torch.empty((len(offs),), dtype=torch.int32, device=dev)
            d_off.copy_(h_off[:len(offs)], non_blocking=True)
            return d_rec, d_off, rboxes, max_per
        if len(rec):
            d_rec = torch.from_numpy(raw.copy()).to(dev)
        else:
            d_rec = torch.zeros((1, BOX_DTYPE.itemsize), dtype=torch.uint8, device=dev)
        d_off = torch.from_numpy(offs).to(dev)
        return d_rec, d_off, rboxes, max_per

    def run(self, d_masks: torch.Tensor, d_boxes, d_offs, max_per: int, rboxes, fresh_outputs: bool = False):
        """Launch on torch's current stream.  d_masks: [B,H,W] uint8 on this device."""
        if d_masks.dtype != torch.uint8 or d_masks.dim() != 3 or not d_masks.is_contiguous():
            raise CvError("masks must be a contiguous uint8 [B,H,W] device tensor")
        if d_masks.device != self.device:
            raise CvError("masks live on another device")
        B, H, W = d_masks.shape
        new_w = resized_width(H, W)
        if new_w <= 0:
            raise CvError("resized width is 0 (extreme aspect ratio)")
        if fresh_outputs:
            self._bufs = {}
        bufs = self._buffers(B, H, W, new_w)
        cc = self._ccaps()
        st = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.cv_nodes_analyze(
            d_masks.data_ptr(), B, H, W, d_boxes.data_ptr(), d_offs.data_ptr(), int(max_per),
            bufs["emptied"].data_ptr(), bufs["resized"].data_ptr(), bufs["enhanced"].data_ptr(),
            bufs["contours"].data_ptr(), bufs["points"].data_ptr(), bufs["pairs"].data_ptr(),
            bufs["results"].data_ptr(), C.byref(cc), self._ws.data_ptr(), self._ws.numel(), st)
        _lib.check(rc, "cv_nodes_analyze")
        return NodeBatchResult(B, H, W, new_w, dict(self.caps), bufs["emptied"], bufs["resized"], bufs["enhanced"],
                               bufs["contours"], bufs["points"], bufs["pairs"], bufs["results"], rboxes,
                               self.lib.cv_last_launch_count())

    def analyze(self, masks, boxes_list, grow: bool = True) -> NodeBatchResult:
        """masks: [B,H,W] uint8 (numpy => copied to the device; torch cuda tensor => used in place)."""
        if isinstance(masks, np.ndarray):
            masks = torch.from_numpy(np.ascontiguousarray(masks)).to(self.device, non_blocking=False)
        if len(boxes_list) != masks.shape[0]:
            raise CvError("one box list per mask is required")
        with torch.cuda.device(self.device):
            d_rec, d_off, rboxes, max_per = self.upload_boxes(boxes_list, masks.shape[1], masks.shape[2])
            for _ in range(6):
                r = self.run(masks, d_rec, d_off, max_per, rboxes)
                st = int(np.bitwise_or.reduce(r.status())) if r.B else 0
                if not st or not grow:
                    return r
                # capacity overflow: grow the table that overflowed and run again (results are otherwise incomplete)
                if st & 1:
                    self.caps["max_external"] *= 4
                if st & 2:
                    self.caps["max_contours"] *= 4
                if st & 4:
                    self.caps["max_points"] *= 4
                if st & 8:
                    self.caps["max_pairs"] *= 4
                if self.caps["max_pairs"] > (1 << 20) or self.caps["max_points"] > (1 << 24):
                    break
            raise CvError("node analysis capacities could not be satisfied")


def ccl_label(d_masks: torch.Tensor, connectivity: int = 8, want_counts: bool = True):
    """Native-resolution CCL (BASELINE cfg 4): labels[p] = 1 + min linear index of p's component, 0 = background."""
    lib = _lib.load()
    _lib.require_device(d_masks.device.index or 0)
    if d_masks.dtype != torch.uint8 or d_masks.dim() != 3 or not d_masks.is_contiguous():
        raise CvError("masks must be a contiguous uint8 [B,H,W] device tensor")
    B, H, W = d_masks.shape
    labels = torch.empty((B, H, W), dtype=torch.int32, device=d_masks.device)
    counts = torch.empty((B,), dtype=torch.int32, device=d_masks.device) if want_counts else None
    with torch.cuda.device(d_masks.device):
        st = torch.cuda.current_stream(d_masks.device).cuda_stream
        ws_bytes = lib.cv_ccl_workspace_bytes(B, H, W)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=d_masks.device)
        rc = lib.cv_ccl_label(d_masks.data_ptr(), B, H, W, connectivity, labels.data_ptr(),
                              counts.data_ptr() if want_counts else None, ws.data_ptr(), ws_bytes, st)
    _lib.check(rc, "cv_ccl_label")
    return labels, counts
