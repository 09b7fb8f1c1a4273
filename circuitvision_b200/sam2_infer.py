"""Drop-in for the reference's `src/sam2_infer.py` (`/root/reference/src/sam2_infer.py`): same module-level names
(`device`, `SAM2Transforms` :29-128, `MultiKernelRefinement` :130-189, `SAM2ImageWrapper` :191-275,
`get_modified_sam2` :277-410), same argument meaning and return shapes, so `circuit_analyzer.py:17-21,203,245,
347-354` keep working unchanged.  All arithmetic runs in libcv_b200.so (hand-written sm_100a kernels behind the
C ABI of include/cv_b200.h); torch is used for parameters, device memory and streams.  There is no CPU fallback:
without the library and an sm_100 device every compute entry point raises `CvError`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np
import torch
import torch.nn as nn

from . import _lib
from . import sam2_weights as _w
from ._lib import CvError

# select the device for computation (sam2_infer.py:19-25)
device = torch.device("cuda") if torch.cuda.is_available() else torch.device("cpu")

IMAGE_SIZE = 1024
SAM2Params = _w.SAM2Params


class cv_sam2_cfg(C.Structure):
    _fields_ = [("embed_dim", C.c_int32), ("num_heads", C.c_int32), ("stages", C.c_int32 * 4),
                ("window_spec", C.c_int32 * 4), ("global_blocks", C.c_int32 * 8), ("n_global", C.c_int32),
                ("use_refinement", C.c_int32), ("max_batch", C.c_int32), ("operand_fp16", C.c_int32)]


def _declare(lib):
    vp, i32, ll = C.c_void_p, C.c_int, C.c_longlong
    lib.cv_sam2_create.argtypes = [C.POINTER(cv_sam2_cfg), i32, C.POINTER(vp)]
    lib.cv_sam2_destroy.argtypes = [vp]
    lib.cv_sam2_set_tensor.argtypes = [vp, C.c_char_p, vp, i32, ll]
    lib.cv_sam2_finalize.argtypes = [vp]
    lib.cv_sam2_set_max_batch.argtypes = [vp, i32]
    lib.cv_sam2_forward.argtypes = [vp, vp, i32, i32, i32, vp, vp, vp, vp, i32, i32, vp, vp, vp]
    lib.cv_sam2_last_launches.argtypes = [vp]
    lib.cv_sam2_set_debug.argtypes = [vp, i32]
    lib.cv_sam2_read_buffer.argtypes = [vp, C.c_char_p, vp, ll, vp]
    lib.cv_sam2_preprocess.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    lib.cv_sam2_resize_logits.argtypes = [vp, i32, i32, i32, i32, vp, vp, vp, vp]
    lib.cv_sam2_refine.argtypes = [vp, i32, C.POINTER(vp), C.POINTER(vp), vp, C.c_float, vp, vp]


_declared = False


def _libsam():
    global _declared
    lib = _lib.load()
    if not _declared:
        _declare(lib)
        _declared = True
    return lib


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _cuda_index(dev) -> int:
    dev = torch.device(dev)
    if dev.type != "cuda":
        raise CvError(f"circuitvision_b200 runs on sm_100a CUDA devices only (got '{dev}'); there is no CPU fallback")
    return torch.cuda.current_device() if dev.index is None else dev.index


def _as_u8_hwc(x) -> np.ndarray:
    """PIL image or ndarray -> contiguous (H,W,3) uint8 (what torchvision's ToTensor scales by 1/255)."""
    if not isinstance(x, np.ndarray):
        x = np.asarray(x)  # PIL.Image
    if x.ndim != 3 or x.shape[2] != 3 or x.dtype != np.uint8:
        raise CvError(f"expected an (H,W,3) uint8 image, got shape {x.shape} dtype {x.dtype}")
    return np.ascontiguousarray(x)


# ------------------------------------------------------------------------------------------ transforms
class SAM2Transforms(nn.Module):
    """sam2_infer.py:29-128.  `__call__` = ToTensor -> Resize((R,R)) (bilinear, antialias) -> Normalize on the GPU;
    returns a CUDA float tensor [3,R,R] (the reference returns the same values on the CPU and the caller does
    `.unsqueeze(0).to(device)`, circuit_analyzer.py:347, which is a no-op here)."""

    def __init__(self, resolution, mask_threshold, max_hole_area=0.0, max_sprinkle_area=0.0):
        super().__init__()
        if resolution != IMAGE_SIZE:
            raise CvError("SAM2Transforms: the B200 kernels are built for resolution 1024 (sam2_infer.py:200-204)")
        if max_hole_area > 0 or max_sprinkle_area > 0:
            # dead code in the reference (circuit_analyzer.py:245-250 passes 0 for both)
            raise CvError("SAM2Transforms: hole / sprinkle filtering is not on the reference's path (both areas are 0)")
        self.resolution = resolution
        self.mask_threshold = mask_threshold
        self.max_hole_area = max_hole_area
        self.max_sprinkle_area = max_sprinkle_area
        self.mean = [0.485, 0.456, 0.406]
        self.std = [0.229, 0.224, 0.225]
        self.device = device

    def _one(self, x, swap_rb=False) -> torch.Tensor:
        img = _as_u8_hwc(x)
        dev = _cuda_index(self.device)
        _lib.require_device(dev)
        H, W = img.shape[:2]
        with torch.cuda.device(dev):
            d = torch.from_numpy(img).cuda(non_blocking=False)
            tmp = torch.empty((H, IMAGE_SIZE, 3), device="cuda", dtype=torch.float32)
            out = torch.empty((3, IMAGE_SIZE, IMAGE_SIZE), device="cuda", dtype=torch.float32)
            _lib.check(_libsam().cv_sam2_preprocess(d.data_ptr(), H, W, int(swap_rb), tmp.data_ptr(), out.data_ptr(),
                                                    _stream()), "cv_sam2_preprocess")
        return out

    def __call__(self, x):
        return self._one(x)

    def forward_batch(self, img_list):
        return torch.stack([self._one(img) for img in img_list], dim=0)

    def transform_coords(self, coords: torch.Tensor, normalize=False, orig_hw=None) -> torch.Tensor:
        if normalize:
            assert orig_hw is not None
            h, w = orig_hw
            coords = coords.clone()
            coords[..., 0] = coords[..., 0] / w
            coords[..., 1] = coords[..., 1] / h
        return coords * self.resolution

    def transform_boxes(self, boxes: torch.Tensor, normalize=False, orig_hw=None) -> torch.Tensor:
        return self.transform_coords(boxes.reshape(-1, 2, 2), normalize, orig_hw)

    def postprocess_masks(self, masks: torch.Tensor, orig_hw) -> torch.Tensor:
        """:88-128 with the (disabled) hole / sprinkle filters: bilinear resize of the logits to `orig_hw`."""
        if masks.dim() != 4 or masks.shape[2] != masks.shape[3]:
            raise CvError("postprocess_masks expects [B,C,S,S] logits")
        dev = _cuda_index(masks.device if masks.is_cuda else self.device)
        _lib.require_device(dev)
        H, W = int(orig_hw[0]), int(orig_hw[1])
        with torch.cuda.device(dev):
            m = masks.float().cuda().contiguous()
            B, Cc, S, _ = m.shape
            out = torch.empty((B, Cc, H, W), device="cuda", dtype=torch.float32)
            _lib.check(_libsam().cv_sam2_resize_logits(m.data_ptr(), B * Cc, S, H, W, out.data_ptr(), None, None, _stream()),
                       "cv_sam2_resize_logits")
        return out


# ------------------------------------------------------------------------------------------ refinement head
class MultiKernelRefinement(nn.Module):
    """sam2_infer.py:130-189: parallel Conv2d(1->C,k,'same') + exact GELU, concat, 1x1 combine.  The fused kernel is
    built for the configuration the reference instantiates (kernels 3,5,7,11, 4 channels: sam2_infer.py:211-216)."""

    def __init__(self, in_channels=1, out_channels=1, kernel_sizes=[3, 5, 7, 9, 11], intermediate_channels=8):
        super().__init__()
        self.kernel_sizes = list(kernel_sizes)
        self.intermediate_channels = intermediate_channels
        self.conv_branches = nn.ModuleList(
            nn.Conv2d(in_channels, intermediate_channels, k, padding="same") for k in kernel_sizes)
        self.activation = nn.GELU()
        self.combiner_conv = nn.Conv2d(len(kernel_sizes) * intermediate_channels, out_channels, 1)

    def supported(self) -> bool:
        return tuple(self.kernel_sizes) == _w.REFINE_KERNELS and self.intermediate_channels == 4

    def forward(self, x):
        if not self.supported():
            raise CvError(f"MultiKernelRefinement: the B200 kernel covers kernel sizes {_w.REFINE_KERNELS} with 4 "
                          f"channels per branch (the reference's configuration); got {self.kernel_sizes} x "
                          f"{self.intermediate_channels}")
        if x.dim() != 4 or x.shape[1] != 1 or x.shape[2:] != (IMAGE_SIZE, IMAGE_SIZE):
            raise CvError("MultiKernelRefinement: expected [B,1,1024,1024] logits")
        dev = _cuda_index(x.device)
        _lib.require_device(dev)
        with torch.cuda.device(dev):
            xin = x.float().contiguous()
            ws = [b.weight.detach().float().cuda().contiguous() for b in self.conv_branches]
            bs = [b.bias.detach().float().cuda().contiguous() for b in self.conv_branches]
            cw = self.combiner_conv.weight.detach().float().cuda().reshape(16).contiguous()
            cb = float(self.combiner_conv.bias.detach().float().cpu()[0])
            out = torch.empty_like(xin)
            wp = (C.c_void_p * 4)(*[t.data_ptr() for t in ws])
            bp = (C.c_void_p * 4)(*[t.data_ptr() for t in bs])
            _lib.check(_libsam().cv_sam2_refine(xin.data_ptr(), xin.shape[0], wp, bp, cw.data_ptr(), cb, out.data_ptr(),
                                                _stream()), "cv_sam2_refine")
        return out


# ------------------------------------------------------------------------------------------ engine
class _Engine:
    """Owns one cv_sam2 handle (device weights + workspace) for one wrapper on one device."""

    def __init__(self, folded: dict, variant: dict, use_refinement: bool, dev: int, max_batch: int, operand_fp16: bool):
        lib = _libsam()
        _lib.require_device(dev)
        cfg = cv_sam2_cfg()
        cfg.embed_dim, cfg.num_heads = variant["embed"], variant["heads"]
        for i in range(4):
            cfg.stages[i] = variant["stages"][i]
            cfg.window_spec[i] = variant["window_spec"][i]
        gb = list(variant["global_blocks"])
        if len(gb) > 8:
            raise CvError("at most 8 global-attention blocks")
        for i, b in enumerate(gb):
            cfg.global_blocks[i] = b
        cfg.n_global = len(gb)
        cfg.use_refinement = int(use_refinement)
        cfg.max_batch = max_batch
        cfg.operand_fp16 = int(operand_fp16)
        self.dev, self.max_batch, self.lib = dev, max_batch, lib
        h = C.c_void_p()
        with torch.cuda.device(dev):
            _lib.check(lib.cv_sam2_create(C.byref(cfg), dev, C.byref(h)), "cv_sam2_create")
            self.h = h
            for name, t in folded.items():
                if t.dtype in (torch.bfloat16, torch.float16):
                    arr, dt = t.view(torch.int16).numpy(), 1
                else:
                    arr, dt = t.numpy(), 0
                _lib.check(lib.cv_sam2_set_tensor(h, name.encode(), arr.ctypes.data, dt, arr.size), f"cv_sam2_set_tensor({name})")
            _lib.check(lib.cv_sam2_finalize(h), "cv_sam2_finalize")

    def set_max_batch(self, n: int):
        if n != self.max_batch:
            with torch.cuda.device(self.dev):
                _lib.check(self.lib.cv_sam2_set_max_batch(self.h, n), "cv_sam2_set_max_batch")
            self.max_batch = n

    def forward(self, images: torch.Tensor, input_kind: int, swap_rb: bool, want_high=True, want_low=True, out_hw=None,
                want_logits=False, want_mask=False):
        """images: cuda uint8 [B,1024,1024,3] (kind 0) or float32 [B,3,1024,1024] (kind 1)."""
        B = images.shape[0]
        opts = dict(device=images.device)
        low = torch.empty((B, 1, 256, 256), dtype=torch.float32, **opts) if want_low else None
        iou = torch.empty((B, 1), dtype=torch.float32, **opts)
        high = torch.empty((B, 1, IMAGE_SIZE, IMAGE_SIZE), dtype=torch.float32, **opts) if want_high else None
        oh, ow = out_hw if out_hw is not None else (IMAGE_SIZE, IMAGE_SIZE)
        mask = torch.empty((B, oh, ow), dtype=torch.uint8, **opts) if want_mask else None
        logits = torch.empty((B, 1, oh, ow), dtype=torch.float32, **opts) if want_logits else None
        ext = torch.empty((B, 4), dtype=torch.int32, **opts) if want_mask else None
        p = lambda t: None if t is None else t.data_ptr()
        self.launches = 0
        with torch.cuda.device(self.dev):
            for s in range(0, B, self.max_batch):
                e = min(B, s + self.max_batch)
                sl = lambda t: None if t is None else t[s:e].data_ptr()
                _lib.check(self.lib.cv_sam2_forward(self.h, images[s:e].data_ptr(), input_kind, int(swap_rb), e - s, sl(low),
                                                    sl(iou), sl(high), sl(mask), oh, ow, sl(logits), sl(ext), _stream()),
                           "cv_sam2_forward")
                self.launches += self.lib.cv_sam2_last_launches(self.h)
        return dict(high=high, low=low, iou=iou, mask=mask, logits=logits, extents=ext)

    def set_debug(self, count_saturation: bool):
        """Debug tap: count fp32 -> fp16 conversions that saturated (+-65504) in the 16-bit activation buffers."""
        _lib.check(self.lib.cv_sam2_set_debug(self.h, int(bool(count_saturation))), "cv_sam2_set_debug")

    def saturation_count(self) -> int:
        """Saturated fp16 conversions of the last forward (needs set_debug(True) before it; 0 in bf16 mode)."""
        return int(self.read_buffer("satcount", (1,), torch.int32).cpu()[0])

    def read_buffer(self, name: str, shape, dtype=torch.float32) -> torch.Tensor:
        with torch.cuda.device(self.dev):
            t = torch.empty(shape, dtype=dtype, device="cuda")
            _lib.check(self.lib.cv_sam2_read_buffer(self.h, name.encode(), t.data_ptr(), t.numel() * t.element_size(),
                                                    _stream()), "cv_sam2_read_buffer")
        return t

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.cv_sam2_destroy(self.h)
                self.h = None
        except Exception:
            pass


# ------------------------------------------------------------------------------------------ wrapper
class SAM2ImageWrapper(nn.Module):
    """sam2_infer.py:191-275.  `modified_sam2_model` is a `SAM2Params` tree (what `get_modified_sam2` builds here in
    place of sam2's SAM2Base); parameter names match the reference's so fine-tuned checkpoints load
    (`load_state_dict` also accepts the PEFT-wrapped key names and merges the LoRA factors)."""

    def __init__(self, modified_sam2_model, embedding_r=4, use_refinement=False, refinement_kernel_sizes=[3, 5, 7, 9, 11]):
        super().__init__()
        self.sam2_model = modified_sam2_model
        self.use_refinement = use_refinement
        self._bb_feat_sizes = [(256, 256), (128, 128), (64, 64)]
        self.embedding_r = embedding_r
        self.dense_embedding1 = nn.Parameter(torch.randn(1, 256, self.embedding_r))
        self.dense_embedding2 = nn.Parameter(torch.randn(1, self.embedding_r, 64 * 64))
        self.sparse_embedding = nn.Parameter(torch.randn(1, 32, 256))
        if self.use_refinement:
            self.refinement_layer = MultiKernelRefinement(in_channels=1, out_channels=1,
                                                          kernel_sizes=refinement_kernel_sizes, intermediate_channels=4)
        else:
            self.refinement_layer = None
        self.max_batch = 8
        # 16-bit tensor-core operand format.  IEEE fp16 (11-bit significand, conversions saturate) is the default: with
        # random-init weights bf16 operands alone leave the IoU >= 0.99 gate no margin (CPU emulation of the roundings,
        # scripts/error_budget.py: 0.993 +- 0.004), fp16 runs at the same tensor-core rate with 8x smaller rounding.
        self.operand_dtype = torch.float16
        self.lora_alpha = 16.0
        self._engine = None
        self._engine_lock = threading.Lock()

    # ---- engine lifetime: any parameter movement / reload drops the folded device copy
    def _apply(self, fn, *a, **k):
        self._engine = None
        return super()._apply(fn, *a, **k)

    def load_state_dict(self, state_dict, strict=True, assign=False):
        wanted = set(self.state_dict().keys())
        sd, report = _w.normalize_state_dict(dict(state_dict), lora_alpha=self.lora_alpha, wanted=wanted)
        self._engine = None
        self.last_load_report = report
        return super().load_state_dict(sd, strict=strict, assign=assign)

    def refresh(self):
        """Call after mutating parameters in place."""
        self._engine = None

    def set_operand_dtype(self, dtype):
        """torch.float16 (default) or torch.bfloat16; rebuilds the device copy of the weights on the next call."""
        if dtype not in (torch.float16, torch.bfloat16):
            raise CvError("operand dtype must be torch.float16 or torch.bfloat16")
        self.operand_dtype = dtype
        self._engine = None

    def set_max_batch(self, n: int):
        self.max_batch = int(n)
        if self._engine is not None:
            self._engine.set_max_batch(self.max_batch)

    def engine(self) -> _Engine:
        with self._engine_lock:
            if self._engine is None:
                p = next(self.parameters())
                dev = _cuda_index(p.device)
                if self.refinement_layer is not None and not self.refinement_layer.supported():
                    raise CvError(f"refinement kernels {self.refinement_layer.kernel_sizes}: the fused tail kernel covers "
                                  f"{_w.REFINE_KERNELS} (the reference's configuration, circuit_analyzer.py:218)")
                folded = _w.fold_state_dict(self.state_dict(), self.sam2_model.variant, self.refinement_layer is not None,
                                            self.operand_dtype)
                self._engine = _Engine(folded, self.sam2_model.variant, self.refinement_layer is not None, dev, self.max_batch,
                                       self.operand_dtype == torch.float16)
            return self._engine

    def segment_batch_u8(self, images_u8: torch.Tensor, swap_rb: bool = True) -> torch.Tensor:
        """Batched, device-resident form of segment_with_sam2's numeric part for crops that are already 1024x1024:
        images_u8 cuda uint8 [B,1024,1024,3] -> wire masks cuda uint8 [B,1024,1024] {0,255}.  Per-image extents are
        left in `self.last_extents` ([B,4] int32, device)."""
        eng = self.engine()
        r = eng.forward(images_u8, 0, swap_rb, want_high=False, want_low=False, want_mask=True)
        self.last_launches = eng.launches
        self.last_extents = r["extents"]
        return r["mask"]

    def forward(self, images, points=None, point_labels=None, masks_prompt=None, multimask_output=False):
        """-> (high_res_masks [B,1,1024,1024], low_res_masks [B,1,256,256], iou_predictions [B,1]); the prompt
        arguments are accepted and ignored exactly as in the reference (:220, they are never read)."""
        if images.dim() != 4 or tuple(images.shape[1:]) != (3, IMAGE_SIZE, IMAGE_SIZE):
            raise CvError(f"SAM2ImageWrapper.forward expects [B,3,1024,1024] (sam2_infer.py:200-204), got {tuple(images.shape)}")
        eng = self.engine()
        with torch.cuda.device(eng.dev):
            x = images.detach().to(device=f"cuda:{eng.dev}", dtype=torch.float32).contiguous()
            r = eng.forward(x, 1, False)
        return r["high"], r["low"], r["iou"]


# ------------------------------------------------------------------------------------------ factory
def get_modified_sam2(model_cfg_path: str, checkpoint_path: str, device: str = "cuda:0" if torch.cuda.is_available() else "cpu",
                      use_high_res_features: bool = True, use_peft: bool = True, lora_rank: int = 12, lora_alpha: int = 16,
                      lora_dropout: float = 0.2, lora_target_modules: list = None, use_wrapper: bool = True,
                      trainable_embedding_r: int = 4, use_refinement_layer: bool = False,
                      refinement_kernels: list = [3, 5, 7, 11], kernel_channels: int = 4,
                      weight_dice=0.5, weight_focal=0.4, weight_iou=0.1, weight_tversky: float = 0.0, weight_tv: float = 0.0,
                      weight_freq: float = 0.0, dice_smooth=1e-5, focal_alpha=0.25, focal_gamma=2.0, iou_smooth=1e-5,
                      iou_threshold=0.5, tversky_alpha=0.2, tversky_beta=0.8, apply_sigmoid=True, lr=1e-3):
    """sam2_infer.py:277-410.  `model_cfg_path`: an upstream hydra yaml (only the trunk hyper-parameters are read) or
    one of 'tiny' / 'small' / 'base_plus' / 'large'.  `checkpoint_path`: upstream `sam2.1_hiera_*.pt` (its image-path
    tensors are loaded) or None / missing file -> PyTorch-default random init.  LoRA is an inference-time rank-r
    additive update: it is merged into the base weights when a fine-tuned state dict is loaded (`load_state_dict`),
    never run as separate GEMMs, so `use_peft` only records alpha for that merge.  Loss / optimizer arguments are
    accepted and unused, as in the reference."""
    if not use_high_res_features:
        raise CvError("use_high_res_features=False is not on the reference's path (circuit_analyzer.py:203-223)")
    if model_cfg_path in _w.VARIANTS:
        variant = _w.VARIANTS[model_cfg_path]
    else:
        variant = _w.variant_from_yaml(model_cfg_path)
    params = SAM2Params(variant)
    if checkpoint_path and os.path.exists(checkpoint_path):
        raw = torch.load(checkpoint_path, map_location="cpu", weights_only=True)
        sd, _ = _w.normalize_state_dict(raw, wanted={"sam2_model." + k for k in params.state_dict().keys()})
        params.load_state_dict({k[len("sam2_model."):]: v for k, v in sd.items()}, strict=True)
    if not use_wrapper:
        return params.to(torch.device(device))
    model = SAM2ImageWrapper(params, embedding_r=trainable_embedding_r, use_refinement=use_refinement_layer,
                             refinement_kernel_sizes=refinement_kernels)
    model.lora_alpha = float(lora_alpha) if use_peft else 0.0
    for p in model.parameters():
        p.requires_grad_(False)
    return model.to(torch.device(device))



def build_random_init(variant: str = "tiny", device="cuda:0", seed: int = 0, max_batch: int = 8, use_refinement: bool = True,
                      operand_dtype=torch.float16):
    """Random-init wrapper of the named architecture (PyTorch-default init; there are no checkpoints offline)."""
    torch.manual_seed(seed)
    m = get_modified_sam2(variant, None, device=str(device), use_refinement_layer=use_refinement)
    m.set_max_batch(max_batch)
    m.set_operand_dtype(operand_dtype)
    return m


# ------------------------------------------------------------------------------------------ segment driver
def segment_to_mask(model: SAM2ImageWrapper, transforms: SAM2Transforms, image_np_bgr: np.ndarray):
    """The numeric part of CircuitAnalyzer.segment_with_sam2 (circuit_analyzer.py:343-370) fused on the device:
    channel swap (:343), ToTensor/Resize/Normalize (:347), forward (:351), postprocess_masks (:354), `> 0` (:356) and
    the extent box of the foreground (:366-370).  Only the uint8 mask (H·W bytes) and 4 ints return to the host.
    -> (mask [H,W] uint8 {0,255}, (xmin, ymin, xmax, ymax) or None)."""
    img = _as_u8_hwc(image_np_bgr)
    H, W = img.shape[:2]
    eng = model.engine()
    with torch.cuda.device(eng.dev):
        d = torch.from_numpy(img).cuda()
        if (H, W) == (IMAGE_SIZE, IMAGE_SIZE):
            r = eng.forward(d[None], 0, True, want_high=False, want_low=False, out_hw=(H, W), want_mask=True)
        else:
            tmp = torch.empty((H, IMAGE_SIZE, 3), device="cuda", dtype=torch.float32)
            x = torch.empty((1, 3, IMAGE_SIZE, IMAGE_SIZE), device="cuda", dtype=torch.float32)
            _lib.check(_libsam().cv_sam2_preprocess(d.data_ptr(), H, W, 1, tmp.data_ptr(), x.data_ptr(), _stream()),
                       "cv_sam2_preprocess")
            r = eng.forward(x, 1, False, want_high=False, want_low=False, out_hw=(H, W), want_mask=True)
        mask = r["mask"][0].cpu().numpy()
        e = r["extents"][0].cpu().tolist()
    bbox = (e[0], e[1], e[2] + 1, e[3] + 1) if e[2] >= 0 else None
    return mask, bbox
