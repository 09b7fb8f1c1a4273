"""circuitvision_b200 — B200-native (sm_100a) implementation of CircuitVision's data-parallel hot path:
SAM 2.1 crop segmentation (`sam2_infer`) + pixel-level node/connection analysis (`circuit_analyzer`).
See DESIGN.md.  The product path has no CPU fallback: libcv_b200.so + an sm_100 device are required."""
from ._lib import CvError, LIB_PATH, load  # noqa: F401

__version__ = "0.1.0"
