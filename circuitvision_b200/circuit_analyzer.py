"""Drop-in for the two hot methods of the reference's `CircuitAnalyzer`
(`/root/reference/src/circuit_analyzer.py`): `segment_with_sam2` (:321-386) and `get_node_connections`
(:1286-1605).  Same names, argument meaning, return structures and error behaviour, so
`src/analysis_pipeline.py:206,234` can call this class unchanged (INTEGRATION.md shows the two-line patch).

Everything numeric runs in libcv_b200.so on the GPU; this file is box bookkeeping plus the UI-only debug
drawings (cv2 drawing calls on the 600-row canvases, exactly the reference's `drawContours/putText/circle`
calls fed with device-produced contours — SURVEY §8 row a18; byte-compared with the reference's three images in
tests/test_nodes_gpu.py).
"""
from __future__ import annotations

import threading
import traceback

import numpy as np
import torch

from . import nodes as _nodes
from ._lib import CvError

_PALETTE = [(255, 0, 0), (0, 255, 0), (0, 0, 255), (255, 255, 0), (0, 255, 255), (255, 0, 255), (255, 128, 0),
            (128, 0, 255), (0, 255, 128), (255, 192, 203), (173, 216, 230), (255, 165, 0), (127, 255, 212),
            (240, 230, 140), (255, 105, 180)]  # circuit_analyzer.py:415-431


def _gray2bgr(img):
    return np.repeat(img[:, :, None], 3, axis=2)


class CircuitAnalyzer:
    """Hot-path subset of the reference class.  `sam2_model` / `sam2_transforms` are the objects from
    `circuitvision_b200.sam2_infer` (same attribute names as the reference, :203,:245)."""

    non_components = set(_nodes.NON_COMPONENTS)
    source_components = set(_nodes.SOURCE_COMPONENTS)

    def __init__(self, sam2_model=None, sam2_transforms=None, use_sam2=None, debug=False, device=0,
                 render_debug_images=True, class_names=None):
        self.debug = debug
        self.device = torch.device("cuda", device) if isinstance(device, int) else torch.device(device)
        self.sam2_model = sam2_model
        self.sam2_transforms = sam2_transforms
        self.sam2_device = self.device
        self.use_sam2 = (sam2_model is not None) if use_sam2 is None else use_sam2
        self.last_sam2_output = None
        self.render_debug_images = render_debug_images
        self._lock = threading.Lock()  # the app shares one analyzer across session threads (app.py:134)
        self._node_analyzer = None
        self._terminal_analyzer = None
        # {numeric id: class name}: what the reference reads from self.yolo.model.names (:2259); YOLO itself is out of scope
        self.class_names = dict(class_names or {})

    # ------------------------------------------------------------------ terminal reclassification
    def _ta(self):
        if self._terminal_analyzer is None:
            from . import terminals as _terminals
            self._terminal_analyzer = _terminals.TerminalAnalyzer(self.device)
        return self._terminal_analyzer

    def reclassify_terminals_based_on_connectivity(self, image_rgb_original, bboxes_list_to_modify):
        """Reference :2217.  Modifies `bboxes_list_to_modify` in place: a 'terminal' whose box is near two or more
        distinct wire contours of the page becomes 'voltage.dc' (keys set exactly as :2295-2307).  Returns None."""
        from . import terminals as _terminals
        page = np.ascontiguousarray(np.asarray(image_rgb_original))
        if page.ndim != 3 or page.shape[2] != 3 or page.dtype != np.uint8:
            raise CvError("image_rgb_original must be an (H, W, 3) uint8 array")
        with self._lock:
            r = self._ta().analyze(page[None], [list(bboxes_list_to_modify)])
            if int(r.status()[0]):
                raise CvError(f"terminal analysis overflowed a capacity (status {int(r.status()[0])})")
            counts = r.counts(0)
        _terminals.apply_reclassification(bboxes_list_to_modify, counts, self.class_names)

    def reclassify_terminals_batch(self, pages_rgb, boxes_list):
        """Batched form: pages [B,H,W,3] uint8 (numpy or cuda tensor); relabels every list in place and returns the
        device-resident `TerminalBatchResult`."""
        from . import terminals as _terminals
        with self._lock:
            r = self._ta().analyze(pages_rgb, boxes_list)
            r.to_host()  # the analyzer's device buffers are reused by the next call: read them while holding the lock
        for b, boxes in enumerate(boxes_list):
            _terminals.apply_reclassification(boxes, r.counts(b), self.class_names)
        return r

    # ------------------------------------------------------------------ YOLO-cluster crop (host geometry)
    def crop_image_and_adjust_bboxes(self, image_to_crop, all_yolo_bboxes_input, padding=20):
        """Reference :937 — (crop view, boxes shifted into it, crop_debug_info); circuitvision_b200/crop.py."""
        from . import crop as _crop
        return _crop.crop_image_and_adjust_bboxes(image_to_crop, all_yolo_bboxes_input, padding, self.non_components)

    # ------------------------------------------------------------------ netlist lines
    def generate_netlist_from_nodes(self, node_list):
        """Reference :1607 — host bookkeeping on the node table (circuitvision_b200/netlist.py)."""
        from . import netlist as _netlist
        return _netlist.generate_netlist_from_nodes(node_list)

    def stringify_line(self, netlist_line):
        """Reference :1909."""
        from . import netlist as _netlist
        return _netlist.stringify_line(netlist_line)

    # ------------------------------------------------------------------ node analysis
    def _na(self):
        if self._node_analyzer is None:
            self._node_analyzer = _nodes.NodeAnalyzer(self.device)
        return self._node_analyzer

    def get_node_connections(self, _image_for_context, processing_wire_mask, bboxes_relative_to_mask):
        """Reference :1286.  Returns (new_nodes_list, emptied_mask, enhanced, contour_image_viz,
        final_node_viz_image, connection_points_visualization)."""
        if processing_wire_mask is None:  # :1291-1305
            h, w = (100, 100)
            if _image_for_context is not None:
                h, w = _image_for_context.shape[:2]
            blank = np.zeros((h, w, 3), np.uint8)
            return [], blank, blank, blank, blank, blank
        mask = np.asarray(processing_wire_mask)
        if mask.ndim != 2 or mask.dtype != np.uint8:
            raise CvError("processing_wire_mask must be a 2-D uint8 array")
        with self._lock:
            r = self._na().analyze(mask[None], [list(bboxes_relative_to_mask)])
            nodes = r.nodes(0)
            emptied = r.emptied[0].cpu().numpy()
            enhanced = r.enhanced[0].cpu().numpy()
            resized = r.resized[0].cpu().numpy()
            contours = r.all_contours(0) if self.render_debug_images else []
            points = r.connection_points(0)
        contour_viz, final_viz, conn_viz = self._render(resized, contours, nodes, points)
        return nodes, emptied, enhanced, contour_viz, final_viz, conn_viz

    def get_node_connections_batch(self, masks, boxes_list):
        """Batched form used by bench.py / multi-image callers: masks [B,H,W] uint8 (numpy or cuda tensor).
        Returns the device-resident `NodeBatchResult`; `.nodes(b)` gives the reference structure of image b."""
        with self._lock:
            r = self._na().analyze(masks, boxes_list)
            r.tables_to_host()  # the analyzer's device tables are reused by the next call: read them under the lock
            return r

    def _render(self, resized, contours, nodes, points):
        """UI-only drawings (:414-458, :1585-1603) on the 600-row canvases."""
        h, w = resized.shape
        contour_viz = np.zeros((h, w, 3), np.uint8)
        final_viz = _gray2bgr(resized).copy()
        if not self.render_debug_images:
            return contour_viz, final_viz, contour_viz.copy()
        import cv2  # drawing only
        for c in contours:
            cv2.drawContours(contour_viz, [c["contour"]], -1, _PALETTE[c["id"] % len(_PALETTE)], 2)
            M = cv2.moments(c["contour"])
            cx, cy = (int(M["m10"] / M["m00"]), int(M["m01"] / M["m00"])) if M["m00"] != 0 else (0, 0)
            cv2.putText(contour_viz, str(c["id"]), (cx + 10, cy + 10), cv2.FONT_HERSHEY_SIMPLEX, 0.5, (255, 0, 0), 2)
        for n in nodes:
            M = cv2.moments(n["contour"])
            if M["m00"] != 0:
                cx, cy = int(M["m10"] / M["m00"]), int(M["m01"] / M["m00"])
                cv2.drawContours(final_viz, [n["contour"]], -1, (0, 255, 0), 2)
                cv2.putText(final_viz, str(n["id"]), (cx - 10, cy + 10), cv2.FONT_HERSHEY_SIMPLEX, 0.9, (0, 0, 255), 2)
        conn_viz = contour_viz.copy()
        for p in points:  # without valid nodes there are no points either (:1443 appends only with an attachment)
            cv2.circle(conn_viz, p, radius=5, color=(0, 255, 255), thickness=-1)
        return contour_viz, final_viz, conn_viz

    # ------------------------------------------------------------------ SAM 2 segmentation
    def segment_with_sam2(self, image_np_bgr):
        """Reference :321.  Never raises: any failure prints a traceback and returns (None, None, None)."""
        if not self.use_sam2 or self.sam2_model is None or self.sam2_transforms is None:
            print("SAM 2 is not available or not initialized. Cannot segment.")
            self.last_sam2_output = None
            return None, None, None
        try:
            from . import sam2_infer
            with self._lock:
                mask, bbox = sam2_infer.segment_to_mask(self.sam2_model, self.sam2_transforms, image_np_bgr)
            colored = np.zeros(mask.shape + (3,), np.uint8)
            colored[:, :, 1] = mask  # :359-361 only the G channel keeps the mask
            self.last_sam2_output = colored
            return mask, colored, bbox
        except Exception as e:  # noqa: BLE001 — reference behaviour (:381-386)
            print(f"Error during SAM 2 segmentation: {e}")
            print(traceback.format_exc())
            self.last_sam2_output = None
            return None, None, None
