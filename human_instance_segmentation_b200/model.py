"""Host-side mirror of the reference model API, executing on ``libhis_b200.so``.

Mirrors (same names, kwargs, state-dict keys, return structure):
  * ``create_rgb_hierarchical_model``  hed/advanced/hierarchical_segmentation_rgb.py:925-1026
  * ``HierarchicalRGBSegmentationModelWithFullImagePretrainedUNet``  rgb.py:564-774
  * ``PreTrainedPeopleSegmentationUNetWrapper`` / ``PreTrainedPeopleSegmentationUNet``
    hed/advanced/hierarchical_segmentation_unet.py:1708-1992
  * ``RGBHierarchicalWrapper`` outputs (``infer``)  hed/export_onnx_advanced.py:353-420

``forward(images, rois) -> (logits [N,3,mh,mw] fp32, aux dict)``.  Inference only (eval semantics:
BatchNorm uses running statistics, dropout is identity).  Arithmetic: fp16 operands / activations,
fp32 accumulation and fp32 tails (logits, softmax, combine) -- see DESIGN.md for the tolerance.
There is no torch-op fallback: without a CUDA device or the built library the calls raise.
"""
from __future__ import annotations

import collections
import ctypes
import os
import warnings
from typing import Dict, Optional, Tuple, Union

import torch
import torch.nn as nn

from . import lib as _lib
from . import param_tree as pt
from .engine import ACT, RES_ADD, RES_MUL, RES_NONE, Act, Plan, fold_bn, pack_direct_weight, pack_gemm_weight, pad_vec, round_up
from .roi_align import DynamicRoIAlign


def _on_model_device(fn):
    """Runs a model method with the CUDA current device set to the model's device (function attributes, occupancy queries,
    graph capture and stream handles are resolved against the current device; a model on cuda:1 is otherwise unusable while
    the current device is 0)."""
    import functools

    @functools.wraps(fn)
    def wrapped(self, *a, **k):
        dev = next(self.parameters()).device
        if dev.type == "cuda" and dev.index != torch.cuda.current_device():
            with torch.cuda.device(dev):
                return fn(self, *a, **k)
        return fn(self, *a, **k)
    return wrapped


def _bucket_rois(n: int) -> int:
    """ROI capacity of the launch plan that serves n ROIs (exported contract: dynamic `num_rois`, export_onnx_advanced.py:427-457):
    exact up to 16, then steps of 8 / 32 / 64 -- at most ~12-25 % padded head work, a handful of plans for any stream of counts."""
    if n <= 16:
        return n
    step = 8 if n <= 64 else 32 if n <= 256 else 64
    return (n + step - 1) // step * step


def _bucket_images(b: int) -> int:
    """Image capacity (dynamic `batch_size`): exact up to 16, then multiples of 8."""
    return b if b <= 16 else (b + 7) // 8 * 8


def _pair(v) -> Tuple[int, int]:
    return (int(v), int(v)) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _conv_out(n: int, k: int, s: int) -> int:
    p = ((s - 1) + (k - 1)) // 2
    return (n + 2 * p - k) // s + 1


class PreTrainedPeopleSegmentationUNet(nn.Module):
    """Parameter holder + input normalisation constants (..._unet.py:1708-1916)."""

    def __init__(self, in_channels=3, classes=1, pretrained_weights_path="ext_extractor/2020-09-23a.pth", mean=None, std=None,
                 freeze_weights=False, encoder_name="timm-efficientnet-b3"):
        super().__init__()
        if in_channels != 3 or classes != 1:
            raise NotImplementedError("the B200 path implements the reference's in_channels=3, classes=1 configuration")
        self.encoder_name = encoder_name
        if mean is not None and std is not None:
            self.mean, self.std = list(mean), list(std)
        elif any(v in pretrained_weights_path.lower() for v in ("b0", "b1", "b7")):   # ..._unet.py:1744-1758
            self.mean, self.std = [0.485, 0.456, 0.406], [0.229, 0.224, 0.225]
        else:
            self.mean, self.std = [0.5, 0.5, 0.5], [0.5, 0.5, 0.5]
        self.model = pt.SmpUnetParams(encoder_name)
        self.load_report = None
        if pretrained_weights_path and os.path.exists(pretrained_weights_path):
            self.load_report = self._load_pretrained(pretrained_weights_path)
        else:       # the reference prints the same warning and keeps its random initialisation (..._unet.py:1866)
            warnings.warn(f"Pre-trained weights not found at {pretrained_weights_path!r}: the full-image UNet keeps its initial "
                          "parameters until a checkpoint is loaded with load_state_dict", stacklevel=3)
        self._freeze_bn = bool(freeze_weights)
        if freeze_weights:
            for p in self.model.parameters():
                p.requires_grad = False
        self.register_buffer("norm_mean", torch.tensor(self.mean).view(1, 3, 1, 1))
        self.register_buffer("norm_std", torch.tensor(self.std).view(1, 3, 1, 1))


    def _load_pretrained(self, path: str):
        """The reference constructor's weight loading (..._unet.py:1776-1864): a plain state dict or a checkpoint holding it under
        'state_dict' / 'model_state_dict'; a leading 'model.' or 'unet.' prefix (decided from the first key) is dropped; loaded
        with strict=False.  Returns (missing_keys, unexpected_keys)."""
        ckpt = torch.load(path, map_location="cpu", weights_only=False)
        sd = ckpt
        if isinstance(ckpt, dict):
            sd = ckpt.get("state_dict", ckpt.get("model_state_dict", ckpt))
        first = next(iter(sd), "")
        prefix = "model." if first.startswith("model.") else "unet." if first.startswith("unet.") else ""
        sd = {(k[len(prefix):] if prefix and k.startswith(prefix) else k): v for k, v in sd.items()}
        n_enc = sum("encoder" in k for k in sd)           # B0 ~358, B1 ~506, B3 ~572, B7 ~1198 (..._unet.py:1815-1828)
        detected = "b0" if n_enc < 400 else "b1" if n_enc < 540 else "b3" if n_enc < 700 else "b7"
        expected = next((v for v in ("b0", "b1", "b3", "b7") if v in self.encoder_name.lower()), None)
        if expected is not None and detected != expected:
            warnings.warn(f"{path}: {n_enc} encoder keys look like EfficientNet-{detected.upper()} weights, the encoder is {self.encoder_name}")
        res = self.model.load_state_dict(sd, strict=False)
        if res.missing_keys or res.unexpected_keys:
            warnings.warn(f"{path}: loaded with {len(res.missing_keys)} missing and {len(res.unexpected_keys)} unexpected keys "
                          f"(first missing: {res.missing_keys[:5]}, first unexpected: {res.unexpected_keys[:5]})")
        return list(res.missing_keys), list(res.unexpected_keys)


class PreTrainedPeopleSegmentationUNetWrapper(nn.Module):
    """..._unet.py:1919-1992; ``output_conv`` weights pinned to [+1,-1], bias 0 (:1963-1971)."""

    def __init__(self, in_channels=3, pretrained_weights_path="ext_extractor/2020-09-23a.pth", freeze_weights=False,
                 encoder_name="timm-efficientnet-b3", **_ignored):
        super().__init__()
        self.model = PreTrainedPeopleSegmentationUNet(in_channels, 1, pretrained_weights_path, None, None, freeze_weights, encoder_name)
        self.output_conv = nn.Conv2d(1, 2, kernel_size=1)
        with torch.no_grad():
            self.output_conv.weight.data[0, 0, 0, 0] = 1.0
            self.output_conv.weight.data[1, 0, 0, 0] = -1.0
            self.output_conv.bias.data.zero_()

    def forward(self, x):
        """Standalone use (the exporter calls ``model.pretrained_unet(images)``, export_onnx_advanced.py:374-377).
        Returns ``(two_channel_logits [B,2,H,W], [])`` like the reference."""
        owner = getattr(self, "_owner", None)
        if owner is None:
            raise _lib.HisError("PreTrainedPeopleSegmentationUNetWrapper must be owned by a B200 model to run")
        return owner()._run_unet_only(x), []


class _PlannedModel(nn.Module):
    """Shared execution plumbing: a launch plan per input geometry, replayed on the caller's stream (no torch-op fallback)."""

    def _init_exec_state(self):
        self.aux_outputs = "full"        # "full" (reference dict) | "light" (no 256-channel tensors) | "none"
        # "fast": fp16 operands / activations (11 significant bits, the tensor-core operand precision), documented bounds in
        # DESIGN section 4.  "strict": split-fp16 operands (hi + lo planes, three MMA passes, ~21 bits) -- meets north_star's
        # <= 1e-3 / >= 99.9 % against the reference's fp32 evaluation on any weights, at ~2.5x the time.
        self.precision = "fast"
        self.copy_outputs = True         # return fresh tensors (False: views of the plan's static buffers)
        self.use_cuda_graph = False
        self.max_rois_per_pass = None    # None: derived from the ROI size (bounds the activation footprint)
        self.max_images_per_pass = None  # None: derived from the image size / encoder width
        # Launch plans are keyed on CAPACITY buckets of (batch, ROI count): a request with fewer images / ROIs runs on the bucket's
        # plan with a masked tail (zero images, zero-area ROIs; outputs sliced), bit-identical to an exact-size plan because no
        # kernel mixes data across images or ROIs.  Plans are evicted least-recently-used once their buffers exceed the budget.
        self.plan_buckets = True
        self.max_plan_bytes = None       # None: 70 % of the device's memory
        self._plans: "collections.OrderedDict[tuple, _BuiltPlan]" = collections.OrderedDict()
        self.eval()

    def train(self, mode: bool = True):
        if mode:
            raise NotImplementedError("the B200 path is inference-only (eval semantics); train with the reference")
        return super().train(False)

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        r = super().load_state_dict(state_dict, strict=strict, **kw)
        self.invalidate()
        return r

    def invalidate(self):
        """Drop packed weights / plans (call after mutating parameters in place)."""
        self._plans.clear()
        if getattr(self, "_pipe", None) is not None:
            torch.cuda.synchronize()
            self._pipe = None

    def _apply(self, fn, *a, **k):
        r = super()._apply(fn, *a, **k)
        self._plans.clear()
        if getattr(self, "_pipe", None) is not None:        # streams / events of the pipelined API belong to the old device
            torch.cuda.synchronize()
            self._pipe = None
        return r

    @torch.no_grad()
    @_on_model_device
    def forward(self, images: torch.Tensor, rois: torch.Tensor):
        bp = self._get_plan(images, rois)
        bp.load_inputs(images, rois)
        bp.plan.replay()
        return bp.outputs(self.copy_outputs)

    def plan_bytes(self) -> int:
        """HBM held by the cached launch plans."""
        return sum(bp.nbytes for bp in self._plans.values())

    def _aligners(self):
        if hasattr(self, "roi_aligns"):
            return list(self.roi_aligns.values())
        return [self.roi_align_mask, self.roi_align_rgb] if hasattr(self, "roi_align_mask") else [self.roi_align]

    def _get_plan(self, images: torch.Tensor, rois: torch.Tensor, slot: int = 0) -> "_BuiltPlan":
        if not images.is_cuda and not torch.cuda.is_available():
            raise _lib.HisError("human_instance_segmentation_b200 needs a CUDA (B200) device; there is no CPU fallback")
        if images.dim() != 4 or images.shape[1] != 3:
            raise ValueError(f"images must be [B,3,H,W], got {tuple(images.shape)}")
        if rois.dim() != 2 or rois.shape[1] != 5:
            raise ValueError(f"rois must be [N,5] = [batch_idx,x1,y1,x2,y2], got {tuple(rois.shape)}")
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise _lib.HisError("move the model to a CUDA device first (model.to('cuda')); there is no CPU fallback")
        B, _, H, W = images.shape
        if self.precision not in ("fast", "strict"):
            raise ValueError(f"precision must be 'fast' or 'strict', got {self.precision!r}")
        N = rois.shape[0]
        # the boundary refiner normalises its edge map over the WHOLE [N,3,mh,mw] tensor (..._refinement.py:94-129): padded ROIs would
        # take part in that min / max, so such models keep exact-size plans
        bucket = self.plan_buckets and not getattr(self, "use_boundary_refinement", False)
        Bc, Nc = (_bucket_images(B), _bucket_rois(N)) if bucket else (B, N)
        key = (Bc, H, W, Nc, self.aux_outputs, self.precision, self.max_rois_per_pass, self.max_images_per_pass,
               tuple((_scale_hw(ra), ra.aligned) for ra in self._aligners()), dev.index, slot, self.use_cuda_graph)
        bp = self._plans.get(key)
        if bp is None:
            self.plans_built = getattr(self, "plans_built", 0) + 1
            budget = self.max_plan_bytes if self.max_plan_bytes is not None else int(0.7 * torch.cuda.get_device_properties(dev).total_memory)
            self._evict(budget)
            bp = _BuiltPlan(self, dev, Bc, H, W, Nc)
            if self.use_cuda_graph:
                bp.plan.capture()
            self._plans[key] = bp
            self.plan_log = getattr(self, "plan_log", [])[-63:] + [("build", Bc, Nc, self.aux_outputs, self.precision, slot, round(bp.nbytes / 1e9, 2))]
            self._evict(budget, protect=key)
        else:
            self._plans.move_to_end(key)
        bp.set_active(B, N)
        return bp

    def _evict(self, keep_bytes: int, protect=None):
        """Drops least-recently-used plans until the cache fits the byte budget (the plan in use is never dropped; plans still
        referenced by the pipelined API's in-flight batches are kept alive by those references until they finish)."""
        while self._plans and self.plan_bytes() > keep_bytes:
            victim = next((k for k in self._plans if k != protect), None)
            if victim is None:
                break
            torch.cuda.synchronize()           # its kernels may still be running
            self.plan_log = getattr(self, "plan_log", [])[-63:] + [("evict", victim[0], victim[3], round(self._plans[victim].nbytes / 1e9, 2))]
            del self._plans[victim]


class HierarchicalRGBSegmentationModelWithFullImagePretrainedUNet(_PlannedModel):
    """rgb.py:564-774 on the B200 kernels."""

    def __init__(self, roi_size: Union[int, Tuple[int, int]] = 28, mask_size: Union[int, Tuple[int, int]] = 56,
                 pretrained_weights_path: str = "ext_extractor/2020-09-23a.pth", use_attention_module: bool = False,
                 freeze_pretrained_weights: bool = False, use_boundary_refinement: bool = False,
                 use_progressive_upsampling: bool = False, use_subpixel_conv: bool = False, use_contour_detection: bool = False,
                 use_distance_transform: bool = False, normalization_type: str = "layernorm2d", normalization_groups: int = 8,
                 activation_function: str = "relu", activation_beta: float = 1.0, **kwargs):
        super().__init__()
        base = kwargs.get("hierarchical_base_channels", 96)
        depth = kwargs.get("hierarchical_depth", 3)
        self.roi_size = _pair(roi_size)
        self.mask_size = _pair(mask_size)
        self.activation_function = pt.check_activation(activation_function)
        self.activation_beta = float(activation_beta)
        self.normalization_type = normalization_type
        self.use_attention_module = bool(use_attention_module)
        self.use_contour_detection = bool(use_contour_detection)
        self.use_distance_transform = bool(use_distance_transform)
        self.use_progressive_upsampling = bool(use_progressive_upsampling)
        self.use_boundary_refinement = bool(use_boundary_refinement)
        self.use_subpixel_conv = bool(use_subpixel_conv)
        self.use_refinement = bool(use_contour_detection or use_distance_transform or use_boundary_refinement or use_subpixel_conv
                                   or use_progressive_upsampling)   # rgb.py:683-689
        self.pretrained_unet = PreTrainedPeopleSegmentationUNetWrapper(
            in_channels=3, pretrained_weights_path=pretrained_weights_path, freeze_weights=freeze_pretrained_weights,
            encoder_name=kwargs.get("encoder_name", "timm-efficientnet-b3"))
        import weakref
        object.__setattr__(self.pretrained_unet, "_owner", weakref.ref(self))
        self.roi_align_mask = DynamicRoIAlign(spatial_scale=640.0, sampling_ratio=2, aligned=True)
        self.roi_align_rgb = DynamicRoIAlign(spatial_scale=640.0, sampling_ratio=2, aligned=True)
        n = pt.norm_spec(normalization_type, normalization_groups)
        self.rgb_feature_extractor = nn.Sequential(
            nn.Conv2d(3, 64, 3, padding=1), pt.norm_params(n, 64), pt.Slot(), pt.ResidualBlockParams(64, n),
            nn.Conv2d(64, 128, 3, padding=1), pt.norm_params(n, 128), pt.Slot(), pt.ResidualBlockParams(128, n),
            nn.Conv2d(128, 256, 3, padding=1), pt.norm_params(n, 256), pt.Slot(), pt.ResidualBlockParams(256, n),
            nn.Conv2d(256, 256, 1), pt.norm_params(n, 256), pt.Slot())
        if self.use_refinement:
            self.feature_combiner = nn.Conv2d(258, 256, 1)
            self.segmentation_head = pt.RefinedHeadParams(256, 256, n, self.use_attention_module, self.use_contour_detection,
                                                          self.use_distance_transform, base, depth, self.use_boundary_refinement,
                                                          self.use_subpixel_conv, self.use_progressive_upsampling)
        else:                                  # rgb.py:715-727: the UNet-guided head takes (features, roi masks) directly
            self.segmentation_head = pt.GuidedHeadParams(256, 256, n, self.use_attention_module)
        self.hierarchical_depth = depth
        self._init_exec_state()

    # ------------------------------------------------------------------ public API
    @torch.no_grad()
    @_on_model_device
    def infer(self, images: torch.Tensor, rois: torch.Tensor, raw_logits: bool = False):
        """Exported-ONNX contract (hed/export_onnx_advanced.py:353-457): returns
        ``(instance_masks [N,1,mh,mw] in {0,1}, binary_masks [B,1,H,W])``; ``raw_logits=True`` gives the older
        released flavour ``(masks [N,3,mh,mw], binary_masks)`` (README.md:537)."""
        from . import postprocess
        bp = self._masks_only_plan(images, rois)
        bp.load_inputs(images, rois)
        bp.plan.replay()
        logits, binary = bp.out_logits, bp.out_binary
        if raw_logits:
            return (logits.clone(), binary.clone()) if self.copy_outputs else (logits, binary)
        inst = postprocess.instance_masks(logits)
        return inst, (binary.clone() if self.copy_outputs else binary)

    def _masks_only_plan(self, images, rois, slot: int = 0):
        """Launch plan of the exported-ONNX contract: only what `masks` / `binary_masks` depend on (no aux tensors, no contour /
        distance branches -- the ONNX export prunes them as dead code, export_onnx_advanced.py:353-457)."""
        saved = self.aux_outputs
        self.aux_outputs = "none"
        try:
            return self._get_plan(images, rois, slot=slot)
        finally:
            self.aux_outputs = saved

    def release_plans(self):
        """Frees every launch plan (and its HBM buffers); the next call rebuilds what it needs."""
        self.invalidate()
        torch.cuda.empty_cache()

    @torch.no_grad()
    @_on_model_device
    def infer_pipelined(self, images: torch.Tensor, rois: torch.Tensor, out_instance_masks: torch.Tensor, out_binary_masks: torch.Tensor):
        """``infer`` for a stream of batches held in (pinned) HOST memory: the call enqueues H2D -> forward -> argmax -> D2H into
        the caller's host tensors on three streams and returns at once; consecutive calls alternate between two launch plans,
        so batch i+1 is uploaded while batch i computes and batch i-1 is downloaded.  Results are valid after
        ``pipeline_sync()`` (or an event wait).  Same arithmetic as ``infer`` (bit-identical outputs)."""
        from . import lib as L_
        dev = next(self.parameters()).device
        st = getattr(self, "_pipe", None)
        if st is None:
            st = self._pipe = {"h2d": torch.cuda.Stream(dev), "compute": torch.cuda.Stream(dev), "d2h": torch.cuda.Stream(dev), "i": 0,
                               "done": [None, None], "fetched": [None, None], "inst": [None, None]}
        slot = st["i"] & 1
        st["i"] += 1
        cur = torch.cuda.current_stream(dev)
        with torch.cuda.stream(st["compute"]):
            bp = self._masks_only_plan(images, rois, slot=slot)     # built (and graph-captured) on first use of the slot
        mh, mw = self.mask_size
        if st["inst"][slot] is None or st["inst"][slot].shape[0] < rois.shape[0]:
            old = st["inst"][slot]
            if old is not None:      # still read by the slot's last download / written by its last argmax on the side streams
                old.record_stream(st["compute"]); old.record_stream(st["d2h"])
            st["inst"][slot] = torch.empty((_bucket_rois(rois.shape[0]), 1, mh, mw), dtype=torch.float32, device=dev)
        inst = st["inst"][slot][: rois.shape[0]]
        # upload: the slot's input buffers are free once the forward that last read them has finished
        st["h2d"].wait_stream(cur)
        if st["done"][slot] is not None:
            st["h2d"].wait_event(st["done"][slot])
        with torch.cuda.stream(st["h2d"]):
            bp.load_inputs(images, rois)
            up = torch.cuda.Event(); up.record()
        # forward + argmax: the slot's output buffers are free once their last download has finished
        st["compute"].wait_event(up)
        if st["fetched"][slot] is not None:
            st["compute"].wait_event(st["fetched"][slot])
        with torch.cuda.stream(st["compute"]):
            bp.plan.replay()
            n = rois.shape[0]
            if n:
                L_.check(L_.load().his_post_instance_mask(bp.out_logits.data_ptr(), n, mh, mw, 0.0, inst.data_ptr(), None,
                                                          ctypes.c_void_p(st["compute"].cuda_stream)), "his_post_instance_mask")
            done = torch.cuda.Event(); done.record()
        st["done"][slot] = done
        # download
        st["d2h"].wait_event(done)
        with torch.cuda.stream(st["d2h"]):
            out_instance_masks.copy_(inst, non_blocking=True)
            out_binary_masks.copy_(bp.out_binary, non_blocking=True)
            fetched = torch.cuda.Event(); fetched.record()
        st["fetched"][slot] = fetched
        return fetched

    def pipeline_sync(self):
        """Waits until every batch enqueued by ``infer_pipelined`` has landed in its host tensors."""
        st = getattr(self, "_pipe", None)
        if st is not None:
            for ev in st["fetched"]:
                if ev is not None:
                    ev.synchronize()
            torch.cuda.current_stream(next(self.parameters()).device).wait_stream(st["d2h"])

    @_on_model_device
    def _run_unet_only(self, images: torch.Tensor) -> torch.Tensor:
        rois = torch.zeros((0, 5), dtype=torch.float32, device=images.device)
        bp = self._get_plan(images, rois)
        bp.load_inputs(images, rois)
        bp.plan.replay()
        return bp.two[: bp.nB].clone()


class HierarchicalRGBSegmentationModel(_PlannedModel):
    """rgb.py:298-439: the model the factory returns without a pre-trained UNet -- one DynamicRoIAlign (aligned=False) on the
    RGB image, ``RGBFeatureExtractor`` (rgb.py:221-295; 3->64->128->192->256, ReLU), then ``HierarchicalSegmentationHeadUNetV2``
    (..._unet.py:670-845; LayerNorm2d and ReLU hard-coded, EnhancedUNet 96/3) or, with a refinement flag, the refined head."""

    def __init__(self, roi_size: Union[int, Tuple[int, int]] = 28, mask_size: Union[int, Tuple[int, int]] = 56, feature_channels: int = 256,
                 num_classes: int = 3, use_attention_module: bool = False, use_boundary_refinement: bool = False,
                 use_progressive_upsampling: bool = False, use_subpixel_conv: bool = False, use_contour_detection: bool = False,
                 use_distance_transform: bool = False, **kwargs):
        super().__init__()
        if num_classes != 3:
            raise AssertionError("Hierarchical model designed for 3 classes")           # ..._unet.py:703
        if feature_channels != 256:
            raise NotImplementedError("feature_channels != 256 is not implemented on B200 (the factory never passes it)")
        self.use_progressive_upsampling = bool(use_progressive_upsampling)
        self.use_boundary_refinement = bool(use_boundary_refinement)
        self.use_subpixel_conv = bool(use_subpixel_conv)
        self.roi_size = _pair(roi_size)
        self.mask_size = _pair(mask_size)
        n = kwargs.get("normalization_type", "layernorm2d")
        self.use_attention_module = bool(use_attention_module)
        self.use_contour_detection = bool(use_contour_detection)
        self.use_distance_transform = bool(use_distance_transform)
        self.use_refinement = bool(use_contour_detection or use_distance_transform or use_boundary_refinement or use_subpixel_conv
                                   or use_progressive_upsampling)
        # the reference passes no activation to any sub-module here: everything is ReLU
        self.activation_function, self.activation_beta = "relu", 1.0
        self.normalization_type = n if self.use_refinement else "layernorm2d"       # head norm (V2 head: LayerNorm2d hard-coded)
        self.extractor_normalization_type = n
        n = pt.norm_spec(n, kwargs.get("normalization_groups", 8))
        self.rgb_extractor = pt.RGBFeatureExtractorParams(n)
        if self.use_refinement:
            self.segmentation_head = pt.RefinedHeadParams(256, 256, n, self.use_attention_module, self.use_contour_detection,
                                                          self.use_distance_transform, 96, 3, self.use_boundary_refinement, self.use_subpixel_conv,
                                                          self.use_progressive_upsampling)
        else:
            self.segmentation_head = pt.BaseHeadParams(256, 256, "layernorm2d", self.use_attention_module, 96, 3)
        self.roi_align = DynamicRoIAlign(spatial_scale=640.0, sampling_ratio=2, aligned=False)      # rgb.py:404-408
        self._init_exec_state()

class MultiScaleRGBSegmentationModel(_PlannedModel):
    """rgb.py:777-922: one RGBFeatureExtractor per ROI scale (DynamicRoIAlign aligned=False), features resized to 28x28, fused
    ('concat' | 'sum' | 'adaptive'), projected 1x1 to 256 channels, HierarchicalSegmentationHeadUNetV2 (LayerNorm2d + ReLU)."""

    def __init__(self, roi_sizes=None, mask_size: Union[int, Tuple[int, int]] = 56, feature_channels: int = 256, fusion_method: str = "concat",
                 num_classes: int = 3, use_attention_module: bool = False, normalization_type: str = "layernorm2d", normalization_groups: int = 8,
                 activation_function: str = "relu", activation_beta: float = 1.0):
        super().__init__()
        if num_classes != 3:
            raise AssertionError("Hierarchical model designed for 3 classes")
        if feature_channels != 256:
            raise NotImplementedError("feature_channels != 256 is not implemented on B200 (the factory never passes it)")
        roi_sizes = dict(roi_sizes) if roi_sizes is not None else {"scale1": 56, "scale2": 42, "scale3": 28}
        self.roi_sizes, self.fusion_method, self.scales = roi_sizes, fusion_method, list(roi_sizes.keys())
        if fusion_method not in ("concat", "sum", "adaptive"):
            self._bad_fusion = fusion_method          # the reference raises at forward time (rgb.py:905-906)
        self.mask_size = _pair(mask_size)
        big = max(int(v) for v in roi_sizes.values())
        self.roi_size = (big, big)                    # bounds the per-pass footprint
        self.feature_size = (28, 28)                  # rgb.py:887: every scale is resized to 28x28
        self.activation_function = pt.check_activation(activation_function)
        self.activation_beta = float(activation_beta)
        self.extractor_normalization_type = normalization_type
        self.normalization_type = "layernorm2d"       # the V2 head hard-codes LayerNorm2d
        self.use_attention_module = bool(use_attention_module)
        self.use_contour_detection = self.use_distance_transform = self.use_refinement = False
        nspec = pt.norm_spec(normalization_type, normalization_groups)
        self.rgb_extractors = nn.ModuleDict({s_: pt.RGBFeatureExtractorParams(nspec) for s_ in self.scales})
        self.roi_aligns = nn.ModuleDict({s_: DynamicRoIAlign(spatial_scale=640.0, sampling_ratio=2, aligned=False) for s_ in self.scales})
        if fusion_method == "adaptive":
            self.fusion_weights = nn.Parameter(torch.ones(len(roi_sizes)))
        fused = 256 * len(roi_sizes) if fusion_method == "concat" else 256
        self.fusion_proj = pt.FusionProjParams(fused, 256, nspec)
        self.segmentation_head = pt.BaseHeadParams(256, 256, "layernorm2d", self.use_attention_module, 96, 3)
        self._init_exec_state()


def _scale_hw(ra: DynamicRoIAlign):
    return float(ra.spatial_scale_h), float(ra.spatial_scale_w)


# ======================================================================================= plan construction
class _CompositePlan:
    """bench/graph-facing view of the (UNet sub-plan x image chunks) + (head sub-plan x ROI chunks) schedule."""

    def __init__(self, bp: "_BuiltPlan"):
        import weakref
        self._bp = weakref.ref(bp)           # the built plan owns this object: no reference cycle, buffers are freed on eviction
        self.graph: Optional[torch.cuda.CUDAGraph] = None

    @property
    def bp(self) -> "_BuiltPlan":
        bp = self._bp()
        if bp is None:
            raise _lib.HisError("launch plan used after it was released")
        return bp

    def _parts(self):
        bp = self.bp
        parts = [(bp.unet_plan, bp.n_unet_chunks)]
        if bp.head_plan is not None:
            parts.append((bp.head_plan, bp.n_head_chunks))
        if bp.post_plan is not None:
            parts.append((bp.post_plan, 1))
        return parts

    @property
    def launches(self) -> int:
        extra = self.bp.n_unet_chunks if self.bp.chunked_unet else 0
        return sum(pl.launches * reps for pl, reps in self._parts()) + extra

    @property
    def flops(self) -> int:
        return sum(pl.flops * reps for pl, reps in self._parts())

    @property
    def tensor_flops(self) -> int:
        return sum(pl.tensor_flops * reps for pl, reps in self._parts())

    @property
    def op_desc(self):
        return [d for pl, reps in self._parts() for _ in range(reps) for d in pl.op_desc]

    def run_timed(self):
        """Per-op CUDA-event times of one full step (chunk copies excluded)."""
        return self.bp.run(timed=True)

    def replay(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self.bp.run()

    def capture(self):
        dev = self.bp.dev
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(dev)
        with torch.cuda.stream(side):
            self.bp.run()                       # warm-up outside capture (module load, attribute sets)
        side.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            self.bp.run()
        self.graph = g


def _rois_per_pass(roi_hw) -> int:
    """ROIs per head pass: bounds the activation footprint (~40 B0-equivalent bytes/ROI-pixel x 2e6 ROI-pixels)."""
    return max(1, int(2.0e6 // (roi_hw[0] * roi_hw[1])) // 32 * 32 or 1)


def _images_per_pass(h: int, w: int, stem_c: int) -> int:
    return max(1, int(64 * 480 * 640 * 32 // (h * w * stem_c)))


class _BuiltPlan:
    """Launch schedule for a fixed (B,H,W,N): static input/output buffers, one UNet sub-plan replayed per image chunk
    and one head sub-plan replayed per ROI chunk (chunks bound the HBM footprint; a single chunk aliases the full
    buffers, so the common case has no extra copies)."""

    def __init__(self, m: HierarchicalRGBSegmentationModelWithFullImagePretrainedUNet, dev, B, H, W, N):
        self.m, self.dev, self.B, self.H, self.W, self.N = m, dev, B, H, W, N      # B, N: CAPACITY of the plan
        self.nB, self.nN = B, N                                                    # images / ROIs of the current request
        self._filled_B, self._filled_N = B, 0          # buffer rows that may hold stale inputs (images: allocated uninitialised)
        self.split = getattr(m, "precision", "fast") == "strict"
        self.S = 1 if self.split else 0          # the `split` argument of the C entry points
        self.act_rgb = {"relu": ACT["relu"], "swish": ACT["silu"], "silu": ACT["silu"], "gelu": ACT["gelu"]}[m.activation_function]
        self.act_ref = {"relu": ACT["relu"], "swish": ACT["swish"], "silu": ACT["silu"], "gelu": ACT["gelu"]}[m.activation_function]
        self.beta = m.activation_beta
        self.has_unet = hasattr(m, "pretrained_unet")
        stem_c = m.pretrained_unet.model.model.encoder.out_channels[1] if self.has_unet else 32
        # strict mode stores two fp16 planes per activation: half the images / ROIs per pass keep the footprint of a plan
        div = 2 if self.split else 1
        bc_max = max(1, m.max_images_per_pass or max(1, _images_per_pass(H, W, stem_c) // div))
        nc_max = max(1, m.max_rois_per_pass or max(1, _rois_per_pass(m.roi_size) // div))

        def balanced(total, cap):
            """Chunk size for `total` items in passes of at most `cap`: every pass runs the full chunk plan, so equal chunks
            (170 ROIs at cap 160 -> 2 x 88, not 160 + a 160-wide pass holding 10)."""
            if total <= cap:
                return total
            n = (total + cap - 1) // cap
            c = (total + n - 1) // n
            return min(cap, (c + 7) // 8 * 8 if c > 16 else c)

        self.Bc = max(1, balanced(B, bc_max)) if self.has_unet else B
        self.Nc = balanced(N, nc_max)
        if B == 0 and N:
            raise ValueError("rois given for an empty image batch")
        self.chunked_unet, self.chunked_head = self.Bc < B, self.Nc < N
        self.n_unet_chunks = (B + self.Bc - 1) // self.Bc if self.has_unet else 0
        self.n_head_chunks = (N + self.Nc - 1) // self.Nc if N else 0
        # ---- UNet sub-plan
        self.unet_plan = self.plan = Plan(dev, self.split)
        p = self.plan
        p.tag = "unet"
        self.images = p.f32(B, 3, H, W)
        self.rois = p.f32(max(N, 1), 5, zero=True)
        self.two = p.f32(B, 2, H, W) if self.has_unet else None
        self.binary = p.f32(B, 1, H, W) if self.has_unet else None
        self.u_images = p.f32(self.Bc, 3, H, W) if self.chunked_unet else self.images
        if self.has_unet:
            self._build_unet()
        # ---- head sub-plan
        self.aux: Dict[str, torch.Tensor] = {}
        self.h_aux: Dict[str, torch.Tensor] = {}
        self.head_plan = None
        mh, mw = m.mask_size
        self.logits = self.unet_plan.f32(N, 3, mh, mw)
        if m.aux_outputs != "none" and self.has_unet:
            self.aux["full_image_logits"] = self.two
        if N:
            self.head_plan = self.plan = Plan(dev, self.split)
            self.plan.tag = "head"
            self.h_rois = self.plan.f32(self.Nc, 5, zero=True) if self.chunked_head else self.rois
            self.h_logits = self.plan.f32(self.Nc, 3, mh, mw) if self.chunked_head else self.logits
            if self.has_unet:
                self._build_head()
            elif hasattr(m, "roi_aligns"):
                self._build_head_multiscale()
            else:
                self._build_head_standard()
            for k, v in self.h_aux.items():
                self.aux[k] = (self.plan.f32(N, *v.shape[1:]) if self.chunked_head else v) if v is not None else None
        # ---- post sub-plan over ALL ROIs (the boundary refiner normalises its edge map over the whole batch tensor)
        self.post_plan = None
        if N and getattr(m, "use_boundary_refinement", False):
            self.post_plan = self.plan = Plan(dev, self.split)
            self.plan.tag = "post"
            self._build_boundary_refiner()
        self.plan = _CompositePlan(self)

    # -------------------------------------------------------------- I/O + schedule
    def set_active(self, n_images: int, n_rois: int):
        assert 0 <= n_images <= self.B and 0 <= n_rois <= self.N
        self.nB, self.nN = n_images, n_rois

    @property
    def nbytes(self) -> int:
        if getattr(self, "_nbytes", None) is None:
            seen, total = set(), 0
            for pl in (self.unet_plan, self.head_plan, self.post_plan):
                for t in (pl.keep if pl is not None else ()):
                    if isinstance(t, torch.Tensor) and t.data_ptr() not in seen:
                        seen.add(t.data_ptr()); total += t.numel() * t.element_size()
            self._nbytes = total
        return self._nbytes

    def load_inputs(self, images: torch.Tensor, rois: torch.Tensor):
        """Copies the request into the plan's static buffers; rows beyond the request (capacity plans) are zeroed when they
        hold an earlier request: a stale 0..255 image would flip the whole-batch `x.max() > 1` decision, a stale ROI costs nothing
        but keeps results reproducible."""
        nb, nn = self.nB, self.nN
        self.images[:nb].copy_(images, non_blocking=True)
        if self._filled_B > nb:
            self.images[nb:self._filled_B].zero_()
        self._filled_B = nb
        if self.N:
            self.rois[:nn].copy_(rois, non_blocking=True)
            if self._filled_N > nn:
                self.rois[nn:self._filled_N].zero_()
            self._filled_N = nn

    @property
    def out_logits(self) -> torch.Tensor:
        return self.logits[: self.nN]

    @property
    def out_binary(self) -> torch.Tensor:
        return self.binary[: self.nB]

    def outputs(self, copy: bool):
        f = (lambda t: t.clone()) if copy else (lambda t: t)
        image_level = ("full_image_logits",)
        aux = {k: (f(v[: self.nB] if k in image_level else v[: self.nN]) if v is not None else None) for k, v in self.aux.items()}
        return f(self.out_logits), aux

    def run(self, timed: bool = False):
        L = self.unet_plan.lib
        H, W = self.H, self.W
        out = []
        for i0 in range(0, self.B if self.has_unet else 0, self.Bc):
            n = min(self.Bc, self.B - i0)
            if self.chunked_unet:
                self.u_images[:n].copy_(self.images[i0:i0 + n])
            if timed:
                out += self.unet_plan.run_timed()
            else:
                self.unet_plan.replay()
            if self.chunked_unet:
                st = torch.cuda.current_stream(self.dev).cuda_stream
                _lib.check(L.his_unet_outputs(self.u_one.data_ptr(), n, H, W, *self._oc, self.two.data_ptr() + i0 * 2 * H * W * 4,
                                              self.binary.data_ptr() + i0 * H * W * 4, ctypes.c_void_p(st)), "his_unet_outputs")
        for j0 in range(0, self.N, max(self.Nc, 1)):
            n = min(self.Nc, self.N - j0)
            if self.chunked_head:
                self.h_rois.zero_()
                self.h_rois[:n].copy_(self.rois[j0:j0 + n])
            if timed:
                out += self.head_plan.run_timed()
            else:
                self.head_plan.replay()
            if self.chunked_head:
                self.logits[j0:j0 + n].copy_(self.h_logits[:n])
                for k, v in self.h_aux.items():
                    if v is not None:
                        self.aux[k][j0:j0 + n].copy_(v[:n])
        if self.post_plan is not None:
            if timed:
                out += self.post_plan.run_timed()
            else:
                self.post_plan.replay()
        return out

    # -------------------------------------------------------------- helpers
    def _is_bn(self, norm) -> bool:
        """BatchNorm2d, or MixedNormalization whose eval forward is its batch_norm alone (normalization_comparison.py:143-147)."""
        return isinstance(norm, (nn.BatchNorm2d, pt.MixedNormParams))

    def conv(self, x: Act, conv: nn.Module, norm, act: int, res: Optional[Act] = None, res_mode: int = RES_NONE,
             out: Optional[Act] = None, out_f32: Optional[torch.Tensor] = None, tail=None, aux_f32: Optional[torch.Tensor] = None,
             in_gate=None, row_scale: Optional[torch.Tensor] = None, stats_out: Optional[torch.Tensor] = None,
             up_input: Optional[Act] = None, res_scale: Optional[torch.Tensor] = None, alg_flops: Optional[int] = None,
             ln_stats: bool = False) -> Optional[Act]:
        """conv (+ folded BatchNorm) (+res) + activation.  Dense shapes -> tcgen05 GEMM, odd shapes -> direct kernel.
        tail = (conv1x1 module with <=2 outputs, sigmoid?, out_f32 NCHW): the following 1x1 conv, fused into the GEMM
        epilogue; the wide activation itself is then not written (its only consumer is the tail).
        aux_f32: fp32 NCHW copy of the output from the epilogue; in_gate: see Plan.conv_gemm (GEMM shapes only)."""
        p = self.plan
        if isinstance(norm, pt.MixedNormParams):
            norm = norm.batch_norm
        if norm is not None and not self._is_bn(norm):
            # LayerNorm2d (hed/model.py:18-38) and the group / instance family: per-sample statistics cannot fold into the conv ->
            # conv(+bias) to fp16, then the statistics + normalise kernels apply norm + residual + activation.
            assert tail is None and out_f32 is None and aux_f32 is None and stats_out is None
            # LayerNorm2d: the statistics of the conv output come from the GEMM epilogue (no separate pass over the tensor)
            raw = self.conv(x, conv, None, ACT["none"], in_gate=in_gate, row_scale=row_scale,
                            ln_stats=isinstance(norm, pt.LayerNorm2dParams) and os.environ.get("HIS_LN_EPILOGUE", "1") != "0")
            return self.layernorm(raw, norm, act, res, res_mode, out)
        transposed = isinstance(conv, nn.ConvTranspose2d)
        w = conv.weight
        cout = w.shape[1] if transposed else w.shape[0]
        cin = w.shape[0] if transposed else w.shape[1]
        k = 1 if transposed else w.shape[2]
        assert cin == x.C, (cin, x.C)
        scale, shift = fold_bn(conv.bias, norm, cout)
        oh, ow = (2 * x.H, 2 * x.W) if transposed else (x.H, x.W)
        use_gemm = cin >= 16 and cout >= 16 and (transposed or k in (1, 3)) and out_f32 is None
        if (not use_gemm and not transposed and k == 3 and cin < 8 and cout >= 16 and out_f32 is None and res is None
                and getattr(x, "zero_tail", False) and x.cs >= 8 and x.c_off == 0):
            # Cin = 3 (RGB patches): the buffer's channel tail is zero by construction, so the layer runs on the tensor cores as a
            # Cin = 8 halo-mode GEMM with zero-padded weights (K = 16 per tap) instead of the CUDA-core direct kernel
            w = torch.cat([w.detach().float().cpu(), torch.zeros(cout, 8 - cin, k, k)], 1)
            x = x.widen(8)
            cin, use_gemm = 8, True
        if out is None and out_f32 is None:
            out = p.null_act(x.N, oh, ow, cout) if (tail is not None and use_gemm) else p.act(x.N, oh, ow, cout)
        if use_gemm:
            nt, bn = ctypes.c_int(), ctypes.c_int()
            p.lib.his_conv_gemm_tile_n(cout, ctypes.byref(nt), ctypes.byref(bn))
            slab = nt.value * bn.value
            wp, cin_pad = pack_gemm_weight(w, slab, transposed, scale, split=self.split)
            tl = None
            if tail is not None:
                tconv, tsig, tout = tail
                tc = tconv.weight.shape[0]
                tw = torch.zeros(tc, slab)
                tw[:, :cout] = tconv.weight.detach().float().cpu().reshape(tc, cout)
                tb = tconv.bias.detach().float().cpu().tolist() + [0.0]
                tl = (p.const(tw), (tb[0], tb[1]), tc, tsig, tout, False)
            p.conv_gemm(x, p.const(wp, torch.float16), cin_pad, p.const(pad_vec(shift, slab)), out,
                        k, act, self.beta, res, res_mode, transposed, tail=tl, aux_f32=aux_f32, in_gate=in_gate,
                        row_scale=row_scale, stats_out=stats_out, up_input=up_input, res_scale=res_scale, alg_flops=alg_flops,
                        ln_stats=ln_stats)
        else:
            if transposed:
                raise NotImplementedError("direct transposed convolution")
            assert tail is None and aux_f32 is None and in_gate is None and row_scale is None and stats_out is None and res_scale is None
            p.conv_direct(x, 0, x.N, x.H, x.W, cin, x.cs, self.direct_w(w), p.const(scale), p.const(shift),
                          cout, k, 1, k // 2, act, self.beta, None, res, res_mode, out, out_f32)
        return out

    def direct_w(self, w: torch.Tensor) -> torch.Tensor:
        """Weights of the CUDA-core convolutions: fp16 [kh][kw][Cin][Cout], fp32 in strict mode."""
        return self.plan.const(pack_direct_weight(w, f32=self.split), torch.float32 if self.split else torch.float16)

    def layernorm(self, x: Act, norm, act: int, res: Optional[Act] = None, res_mode: int = RES_NONE,
                  out: Optional[Act] = None) -> Act:
        """Per-sample statistic norms: LayerNorm2d (one group over C,H,W) or GroupNorm / SpatialGroupNorm (G groups) + affine
        (+ residual) + activation."""
        p, L = self.plan, self.plan.lib
        if out is None:
            out = p.act(x.N, x.H, x.W, x.C)
        if isinstance(norm, pt.INSTANCE_NORMS) and not self.split:
            raise NotImplementedError("normalization_type 'instance' / 'adaptive_instance' / 'foreground_aware' needs model.precision = 'strict': "
                                      "per-channel statistics of a single-fp16 pre-normalisation tensor are 1.5e-2 .. 4e-2 off the reference "
                                      "(DESIGN section 4)")
        if isinstance(norm, pt.ForegroundAwareNormParams):
            # fg detector on the un-normalised input: conv1x1 C -> C/4 + ReLU (GEMM / direct kernel), conv1x1 C/4 -> 1 + sigmoid to fp32
            hid = self.conv(x, norm.fg_detector[0], None, ACT["relu"])
            prob = p.f32(x.N, 1, x.H, x.W)
            self.conv(hid, norm.fg_detector[2], None, ACT["sigmoid"], out_f32=prob)
            parts = L.his_groupnorm_parts(x.N, x.H * x.W, x.C)
            ws = torch.empty((x.N, parts + 1, x.C, 2), dtype=torch.float32, device=self.dev)
            p.keep.append(ws)
            p.add("fgaware_norm", L.his_fgaware_norm_act, x.ptr, x.N, x.H * x.W, x.C, x.cs, p.const(norm.fg_scale).data_ptr(),
                  p.const(norm.fg_bias).data_ptr(), p.const(norm.bg_scale).data_ptr(), p.const(norm.bg_bias).data_ptr(), prob.data_ptr(),
                  float(norm.eps), act, self.beta, res_mode, res.ptr if res is not None else None, res.cs if res is not None else 0,
                  ws.data_ptr(), out.ptr, out.cs, self.S)
            return out
        gn = pt.group_norm_args(norm)
        if gn is not None:
            groups, gamma, beta, eps = gn
            parts = L.his_groupnorm_parts(x.N, x.H * x.W, x.C)
            ws = torch.empty((x.N, parts + 1, x.C, 2), dtype=torch.float32, device=self.dev)
            p.keep.append(ws)
            p.add("groupnorm", L.his_groupnorm_act, x.ptr, x.N, x.H * x.W, x.C, x.cs, int(groups), p.const(gamma.reshape(-1)).data_ptr(),
                  p.const(beta.reshape(-1)).data_ptr(), float(eps), act, self.beta, res_mode, res.ptr if res is not None else None,
                  res.cs if res is not None else 0, ws.data_ptr(), out.ptr, out.cs, self.S)
            return out
        given = 0
        if getattr(x, "ln_ws", None) is not None:         # partial statistics written by the producing GEMM's epilogue
            ws, given = x.ln_ws
        else:
            parts = L.his_layernorm2d_parts(x.N, x.H * x.W, x.C)
            ws = torch.empty((x.N, parts, 2), dtype=torch.float64, device=self.dev)
            p.keep.append(ws)
        p.add("layernorm2d" if not given else "layernorm2d_apply", L.his_layernorm2d_act, x.ptr, x.N, x.H * x.W, x.C, x.cs,
              p.const(norm.weight.reshape(-1)).data_ptr(), p.const(norm.bias.reshape(-1)).data_ptr(), float(norm.eps), act, self.beta, res_mode,
              res.ptr if res is not None else None, res.cs if res is not None else 0, ws.data_ptr(), given, out.ptr, out.cs, self.S)
        return out

    def residual_block(self, x: Act, rb: pt.ResidualBlockParams, act: int, out: Optional[Act] = None, tail=None, aux_f32=None,
                       stats_out=None) -> Act:
        t = self.conv(x, rb.conv1, rb.norm1, act)
        return self.conv(t, rb.conv2, rb.norm2, act, res=x, res_mode=RES_ADD, out=out, tail=tail, aux_f32=aux_f32, stats_out=stats_out)

    def conv_transpose4(self, x: Act, convt: nn.ConvTranspose2d, norm, act: int) -> Act:
        """ConvTranspose2d(k4, s2, p1) + norm + activation (ProgressiveUpsamplingDecoder, ..._refinement.py:161-181).  Output pixel
        (2y+py, 2x+px) sees a 2x2 input neighbourhood through the kernel taps ky = py+1-2dy, kx = px+1-2dx: the four phases run as ONE
        3x3 tensor-core conv with 4*Cout phase-major output channels (unused taps zero), then depth-to-space interleaves them."""
        p, L = self.plan, self.plan.lib
        norm = getattr(norm, "batch_norm", norm)                 # MixedNormalization: eval = its BatchNorm
        wt = convt.weight.detach().float().cpu()                 # [cin, cout, 4, 4]
        cin, cout = wt.shape[:2]
        conv4 = nn.Conv2d(cin, 4 * cout, 3, padding=1)
        tap = {(0, -1): 3, (0, 0): 1, (1, 0): 2, (1, 1): 0}      # (phase, input offset) -> kernel index
        with torch.no_grad():
            conv4.weight.zero_()
            for (py, dy), ky in tap.items():
                for (px, dx), kx in tap.items():
                    ph = py * 2 + px
                    conv4.weight[ph * cout:(ph + 1) * cout, :, dy + 1, dx + 1] = wt[:, :, ky, kx].t()
            conv4.bias.copy_(convt.bias.detach().float().cpu().repeat(4))
        bn4 = None
        if self._is_bn(norm):
            bn4 = nn.BatchNorm2d(4 * cout, eps=norm.eps)
            with torch.no_grad():
                for name in ("weight", "bias", "running_mean", "running_var"):
                    getattr(bn4, name).copy_(getattr(norm, name).detach().float().cpu().repeat(4))
        wide = self.conv(x, conv4, bn4, act if bn4 is not None else ACT["none"])
        out = p.act(x.N, 2 * x.H, 2 * x.W, cout)
        p.add("depth_to_space", L.his_depth_to_space2_half, wide.ptr, x.N, x.H, x.W, cout, wide.cs, out.ptr, out.cs, self.S)
        if bn4 is None:
            out = self.layernorm(out, norm, act)
        return out

    def _aux_fusable(self, cin: int) -> bool:
        """The epilogue export exists for BatchNorm-folded layers whose K block is 64 wide (Cin >= 52)."""
        return self._is_bn_mode() and cin >= 52

    def _is_bn_mode(self) -> bool:
        return self.m.normalization_type.lower() in ("batch", "batchnorm", "batchnorm2d")

    def _tail_ok(self, act: int) -> bool:
        """The fused tail exists for BatchNorm + none/relu epilogues (every preset); otherwise the separate 1x1 kernel runs."""
        return act in (ACT["none"], ACT["relu"]) and self.m.normalization_type.lower() in ("batch", "batchnorm", "batchnorm2d")

    def to_mask_size(self, t: torch.Tensor) -> torch.Tensor:
        """F.interpolate(size=mask, bilinear, align_corners=False) when sizes differ (..._refinement.py:561-566)."""
        mh, mw = self.m.mask_size
        n, c, h, w = t.shape
        if (h, w) == (mh, mw):
            return t
        out = self.plan.f32(n, c, mh, mw)
        self.plan.add("resize_bilinear", self.plan.lib.his_resize_bilinear_f32, t.data_ptr(), n * c, h, w, mh, mw, out.data_ptr())
        return out

    def export_nchw(self, x: Act) -> torch.Tensor:
        out = self.plan.f32(x.N, x.C, x.H, x.W)
        self.plan.add("nhwc2nchw", self.plan.lib.his_nhwc_half_to_nchw_float, x.ptr, x.N, x.H * x.W, x.C, x.cs, out.data_ptr(), self.S)
        return out

    # -------------------------------------------------------------- EfficientNet-UNet (full image)
    def _build_unet(self):
        m, p, L = self.m, self.plan, self.plan.lib
        B, H, W = self.Bc, self.H, self.W
        holder = m.pretrained_unet.model
        net = holder.model
        enc, dec = net.encoder, net.decoder
        # input normalisation folded into the stem loader (device-side max check, no host sync)
        flag = torch.zeros(1, dtype=torch.int32, device=self.dev); p.keep.append(flag)
        affine = p.f32(6)
        mean = (ctypes.c_float * 3)(*[float(v) for v in holder.norm_mean.flatten().tolist()])
        std = (ctypes.c_float * 3)(*[float(v) for v in holder.norm_std.flatten().tolist()])
        p.keep += [mean, std]
        # the reference tests x.max() over the WHOLE batch tensor (..._unet.py:1888), so the flag always comes from self.images
        p.add("input_affine", L.his_unet_input_affine, self.images.data_ptr(), self.B * 3 * H * W, mean, std, flag.data_ptr(), affine.data_ptr())

        # geometry of the five feature levels and the decoder concat buffers
        sizes = [(H, W)]
        h, w = _conv_out(H, 3, 2), _conv_out(W, 3, 2)
        sizes.append((h, w))
        stage_last = {1: None, 2: None, 4: None, 6: None}
        for si, stage in enumerate(enc.blocks):
            for blk in stage:
                h, w = _conv_out(h, blk.k, blk.s), _conv_out(w, blk.k, blk.s)
            if si in stage_last:
                sizes.append((h, w))
        oc = enc.out_channels                     # (3, c1, c2, c3, c4, c5)
        cats = []
        for i, blk in enumerate(dec.blocks):
            th, tw = sizes[4 - i]                 # block i upsamples to feature level (4-i); level 0 = the image
            # (the last block has no skip: its buffer is only needed when the block does not run in phase-packed form, see below)
            cats.append(p.act(B, th, tw, blk.cin + blk.cskip) if blk.cskip else None)
        skip_slot = {1: cats[3].slice(dec.blocks[3].cin, oc[1]), 2: cats[2].slice(dec.blocks[2].cin, oc[2]),
                     3: cats[1].slice(dec.blocks[1].cin, oc[3]), 4: cats[0].slice(dec.blocks[0].cin, oc[4])}

        # stem: conv3x3 s2 (no bias) + BN + SiLU, reading the NCHW fp32 images directly
        stem_c = oc[1]
        x = skip_slot[1]
        scale, shift = fold_bn(None, enc.bn1, stem_c)
        if H % 2 == 0 and W % 2 == 0 and stem_c >= 16:
            # stem on the tensor cores: normalise + space-to-depth the image (one pass), then the 3x3 s2 conv is a 2x2 s1 conv over
            # [B, H/2, W/2, 16], run by the halo-mode GEMM as a 3x3 whose other five taps are zero
            s2d = p.act(B, H // 2, W // 2, 16)
            p.add("s2d_input", L.his_s2d_input, self.u_images.data_ptr(), B, H, W, affine.data_ptr(), s2d.ptr, self.S)
            w = enc.conv_stem.weight.detach().float().cpu()                       # [stem_c, 3, 3, 3]
            w2 = torch.zeros(stem_c, 16, 3, 3)
            for a in (-1, 0):                   # s2d row offset
                for sy in (0, 1):
                    ky = 2 * a + sy + 1
                    if not 0 <= ky <= 2:
                        continue
                    for b_ in (-1, 0):
                        for sx in (0, 1):
                            kx = 2 * b_ + sx + 1
                            if 0 <= kx <= 2:
                                w2[:, (sy * 2 + sx) * 3:(sy * 2 + sx) * 3 + 3, a + 1, b_ + 1] = w[:, :, ky, kx]
            stem = nn.Conv2d(16, stem_c, 3, padding=1, bias=False)
            with torch.no_grad():
                stem.weight.copy_(w2)
            self.conv(s2d, stem, enc.bn1, ACT["silu"], out=x)
        else:
            p.conv_direct(self.u_images, 1, B, H, W, 3, 0, self.direct_w(enc.conv_stem.weight), p.const(scale),
                          p.const(shift), stem_c, 3, 2, 1, ACT["silu"], 1.0, in_affine=affine, out=x)
        level_of_stage = {1: 2, 2: 3, 4: 4}
        # scratch for the per-image gated projection weights, sized for the widest block (the blocks run back to back)
        need = 0
        for stage in enc.blocks:
            for blk in stage:
                proj = blk.conv_pwl if blk.kind == "ir" else blk.conv_pw
                nt, bn = ctypes.c_int(), ctypes.c_int()
                L.his_conv_gemm_tile_n(proj.weight.shape[0], ctypes.byref(nt), ctypes.byref(bn))
                need = max(need, B * nt.value * bn.value * round_up(blk.mid, 64) * (2 if self.split else 1))
        self._gated_w = torch.empty(max(need, 8), dtype=torch.float16, device=self.dev)
        p.keep.append(self._gated_w)
        for si, stage in enumerate(enc.blocks):
            for bi, blk in enumerate(stage):
                last_of_stage = bi == len(stage) - 1
                dst = skip_slot[level_of_stage[si]] if (last_of_stage and si in level_of_stage) else None
                x = self._mbconv(x, blk, dst)
        # decoder (smp UnetDecoder, nearest resize + concat + 2x conv3x3-BN-ReLU)
        head_fmt = 0
        for i, blk in enumerate(dec.blocks):
            cat = cats[i]
            if self._phase_packed_tail(i, blk, x, sizes[4 - i]):
                # Last decoder block (no skip): nearest-2x + conv3x3 + conv3x3 at 480x640 with 16 channels = 153 600 pixel tiles per
                # 64 images whose cost is the per-tile latency of the GEMM pipeline, not their 16-wide MMAs.  The same arithmetic at
                # the LOW resolution in phase-packed form -- channel block (py*2+px) of low pixel (y, x) = full pixel (2y+py, 2x+px):
                #   conv1(up(x)) = a 3x3 conv Cin -> 4*C1 on x (taps that land on the same low pixel summed),
                #   conv2        = a 3x3 conv 4*C1 -> 4*C2 whose weights are conv2's taps scattered by (input phase, low offset)
                # -- a quarter of the tiles at N = 64; the 16 -> 1 head reads the phase-packed tensor directly.
                x = self._decoder_tail_phase_packed(x, blk)
                head_fmt = 2
                continue
            if cat is None:
                cat = p.act(B, sizes[4 - i][0], sizes[4 - i][1], blk.cin + blk.cskip)
            up = cat.slice(0, blk.cin)
            full = cat.widen(blk.cin + blk.cskip)
            fuse_up = (cat.H, cat.W) == (2 * x.H, 2 * x.W) and x.C == blk.cin and \
                L.his_conv_gemm_can_fuse_upsample(cat.H, cat.W, blk.cin + blk.cskip, blk.conv1[0].weight.shape[0], blk.cin) == 1
            if fuse_up:   # nearest 2x + concat happen inside the conv's window loads: the upsampled tensor is never materialised
                y = self.conv(full, blk.conv1[0], blk.conv1[1], ACT["relu"], up_input=x)
            else:
                p.add("resize_nearest", L.his_resize_nearest, x.ptr, B, x.H, x.W, x.C, x.cs, cat.H, cat.W, up.ptr, up.cs, self.S)
                y = self.conv(full, blk.conv1[0], blk.conv1[1], ACT["relu"])
            x = self.conv(y, blk.conv2[0], blk.conv2[1], ACT["relu"])
        # segmentation head conv3x3 16->1 (+bias) -> fp32 logits; output_conv 1->2; export-wrapper binary mask
        head = net.segmentation_head[0]
        self.u_one = p.f32(B, 1, H, W)
        p.conv_direct(x, head_fmt, B, H, W, head.weight.shape[1], x.cs, self.direct_w(head.weight), p.const(torch.ones(1)),
                      p.const(head.bias.detach().float()), 1, 3, 1, 1, ACT["none"], out_f32=self.u_one)
        oc_w = m.pretrained_unet.output_conv.weight.detach().float().flatten().tolist()
        oc_b = m.pretrained_unet.output_conv.bias.detach().float().flatten().tolist()
        self._oc = (oc_w[0], oc_w[1], oc_b[0], oc_b[1])
        if not self.chunked_unet:          # chunked: run() writes each chunk's slice of the full-batch outputs
            p.add("unet_outputs", L.his_unet_outputs, self.u_one.data_ptr(), B, H, W, *self._oc, self.two.data_ptr(), self.binary.data_ptr())

    def _phase_packed_tail(self, i: int, blk, x: Act, out_hw) -> bool:
        if os.environ.get("HIS_UNET_PHASE_TAIL", "1") == "0":
            return False
        dec = self.m.pretrained_unet.model.model.decoder
        c1, c2 = blk.conv1[0].weight.shape[0], blk.conv2[0].weight.shape[0]
        return (i == len(dec.blocks) - 1 and blk.cskip == 0 and tuple(out_hw) == (2 * x.H, 2 * x.W) and x.C == blk.cin and x.C % 8 == 0
                and c1 == 16 and c2 == 16 and self._is_bn(blk.conv1[1]) and self._is_bn(blk.conv2[1]))

    def _decoder_tail_phase_packed(self, x: Act, blk) -> Act:
        """smp UnetDecoderBlock without skip, at the input's resolution: returns the block's output in phase-packed form
        [B, h, w, 4*C2] (see _build_unet)."""
        w1 = blk.conv1[0].weight.detach().float().cpu()           # [C1, Cin, 3, 3]
        w2 = blk.conv2[0].weight.detach().float().cpu()           # [C2, C1, 3, 3]
        c1, cin = w1.shape[:2]
        c2 = w2.shape[0]
        W1 = torch.zeros(4 * c1, cin, 3, 3)
        W2 = torch.zeros(4 * c2, 4 * c1, 3, 3)
        for py in (0, 1):
            for px in (0, 1):
                ph = py * 2 + px
                for dy in (-1, 0, 1):
                    oy, qy = (py + dy) // 2, (py + dy) % 2      # low-resolution row offset and the input phase of full row 2y+py+dy
                    for dx in (-1, 0, 1):
                        ox, qx = (px + dx) // 2, (px + dx) % 2
                        W1[ph * c1:(ph + 1) * c1, :, oy + 1, ox + 1] += w1[:, :, dy + 1, dx + 1]       # up(x) rows 2y+py+dy share low row y+oy
                        W2[ph * c2:(ph + 1) * c2, (qy * 2 + qx) * c1:(qy * 2 + qx + 1) * c1, oy + 1, ox + 1] = w2[:, :, dy + 1, dx + 1]

        def packed(conv_w, bn):
            conv = nn.Conv2d(conv_w.shape[1], conv_w.shape[0], 3, padding=1, bias=False)
            bn4 = nn.BatchNorm2d(conv_w.shape[0], eps=bn.eps)
            with torch.no_grad():
                conv.weight.copy_(conv_w)
                for name in ("weight", "bias", "running_mean", "running_var"):
                    getattr(bn4, name).copy_(getattr(bn, name).detach().float().cpu().repeat(4))
            return conv, bn4

        full_px = x.N * (2 * x.H) * (2 * x.W)
        cv1, bn1 = packed(W1, blk.conv1[1])
        cv2, bn2 = packed(W2, blk.conv2[1])
        y = self.conv(x, cv1, bn1, ACT["relu"], alg_flops=2 * full_px * cin * c1 * 9)
        return self.conv(y, cv2, bn2, ACT["relu"], alg_flops=2 * full_px * c1 * c2 * 9)

    def _mbconv(self, x: Act, blk, dst: Optional[Act]) -> Act:
        """timm DepthwiseSeparableConv / InvertedResidual (oracle/effunet.py _DS/_IR restates them)."""
        p, L = self.plan, self.plan.lib
        B = x.N
        has_skip = blk.s == 1 and blk.cin == blk.cout
        if blk.kind == "ir":
            t = self.conv(x, blk.conv_pw, blk.bn1, ACT["silu"])
            dw_bn, proj, proj_bn = blk.bn2, blk.conv_pwl, blk.bn3
        else:
            t = x
            dw_bn, proj, proj_bn = blk.bn1, blk.conv_pw, blk.bn2
        ho, wo = _conv_out(t.H, blk.k, blk.s), _conv_out(t.W, blk.k, blk.s)
        d = p.act(B, ho, wo, blk.mid)
        parts = L.his_depthwise_pool_parts(B, t.H, t.W, blk.mid, blk.k, blk.s)
        pool = p.f32(B, parts, blk.mid)
        wdw = blk.conv_dw.weight.detach().float().cpu().reshape(blk.mid, blk.k * blk.k).t().contiguous()
        scale, shift = fold_bn(None, dw_bn, blk.mid)
        p.add("depthwise", L.his_depthwise_conv, t.ptr, B, t.H, t.W, blk.mid, t.cs,
              p.const(wdw, torch.float32 if self.split else torch.float16).data_ptr(),          # strict mode: fp32 taps
              p.const(scale).data_ptr(), p.const(shift).data_ptr(), blk.k, blk.s, ACT["silu"], d.ptr, d.cs, pool.data_ptr(), self.S,
              desc=f"N{B} {t.H}x{t.W} C{blk.mid} k{blk.k} s{blk.s}")
        p._add_flops(2 * B * ho * wo * blk.mid * blk.k * blk.k, False)
        se = blk.se
        r = se.conv_reduce.weight.shape[0]
        gate = p.f32(B, blk.mid)
        p.add("se_gate", L.his_se_gate, pool.data_ptr(), parts, B, ho * wo, blk.mid, r,
              p.const(se.conv_reduce.weight.reshape(r, blk.mid)).data_ptr(), p.const(se.conv_reduce.bias).data_ptr(),
              p.const(se.conv_expand.weight.reshape(blk.mid, r)).data_ptr(), p.const(se.conv_expand.bias).data_ptr(),
              ACT["silu"], 1.0, p.f32(B, r).data_ptr(), gate.data_ptr())
        # the SE product x*gate is folded into per-image projection weights (conv(x*g) == conv_{w*g}(x)): no pass over d
        return self.conv(d, proj, proj_bn, ACT["none"], res=x if has_skip else None, res_mode=RES_ADD if has_skip else RES_NONE, out=dst,
                         in_gate=(gate, self._gated_w))

    # -------------------------------------------------------------- per-ROI head
    def _build_head(self):
        m, p, L = self.m, self.plan, self.plan.lib
        N, B, H, W = self.Nc, self.B, self.H, self.W
        rh, rw = m.roi_size
        mh, mw = m.mask_size
        aux_level = m.aux_outputs
        A_rgb, A_ref = self.act_rgb, self.act_ref

        # --- Dynamic RoI Align (rgb.py:751-755): UNet logits -> channels 256..257 of the combiner input, RGB -> patches
        # (guided head: channel 256 of its 257-channel input_adjust input holds sigmoid(fg logit) instead)
        comb_in = p.act(N, rh, rw, 258 if m.use_refinement else 257)
        patches = p.act_zeroed(N, rh, rw, 3)
        roi_feat = p.f32(N, 2, rh, rw)
        roi_patch = p.f32(N, 3, rh, rw) if aux_level != "none" else None
        ram, rar = m.roi_align_mask, m.roi_align_rgb
        msk = comb_in.slice(256, 2) if m.use_refinement else None
        # one launch for both aligners: warp per (ROI, output row), source rows staged in shared memory, outputs written straight
        # into the consumers' NHWC slices (+ the fp32 aux tensors)
        p.add("roi_align", L.his_roi_align_fused,
              self.two.data_ptr(), 2, float(ram.spatial_scale_h), float(ram.spatial_scale_w), 1 if ram.aligned else 0,
              msk.ptr if msk else None, msk.cs if msk else 0, roi_feat.data_ptr(),
              self.images.data_ptr(), 3, float(rar.spatial_scale_h), float(rar.spatial_scale_w), 1 if rar.aligned else 0,
              patches.ptr, patches.cs, roi_patch.data_ptr() if roi_patch is not None else None,
              B, H, W, self.h_rois.data_ptr(), N, rh, rw, self.S)

        # --- rgb_feature_extractor (rgb.py:657-673)
        fe = m.rgb_feature_extractor
        x = patches
        for i in (0, 4, 8):
            x = self.conv(x, fe[i], fe[i + 1], A_rgb)
            x = self.residual_block(x, fe[i + 3], A_ref)
        self.conv(x, fe[12], fe[13], A_rgb, out=comb_in.slice(0, 256))
        if not m.use_refinement:
            self._build_guided_head(comb_in, roi_feat, roi_patch)
            return
        # --- feature_combiner 1x1 258->256 (rgb.py:695,758-762); the concat is the buffer layout itself
        feats = self.conv(comb_in.widen(258), m.feature_combiner, None, ACT["none"])
        self._hier_head(feats, m.segmentation_head.base_head, m.segmentation_head)
        if aux_level != "none":
            self.h_aux["roi_features"] = roi_feat
            self.h_aux["roi_patches"] = roi_patch

    def _build_head_standard(self):
        """HierarchicalRGBSegmentationModel.forward (rgb.py:410-439): RoIAlign(aligned=False) -> RGBFeatureExtractor -> head."""
        m, p, L = self.m, self.plan, self.plan.lib
        N, B, H, W = self.Nc, self.B, self.H, self.W
        rh, rw = m.roi_size
        A = ACT["relu"]
        patches = p.act_zeroed(N, rh, rw, 3)
        roi_patch = p.f32(N, 3, rh, rw) if m.aux_outputs != "none" else None
        ra = m.roi_align
        p.add("roi_align_rgb", L.his_roi_align, self.images.data_ptr(), 0, 3 * H * W, H * W, W, 1, B, 3, H, W, self.h_rois.data_ptr(), N, rh, rw,
              float(ra.spatial_scale_h), float(ra.spatial_scale_w), 1 if ra.aligned else 0, patches.ptr, patches.cs,
              roi_patch.data_ptr() if roi_patch is not None else None, self.S)
        # RGBFeatureExtractor (rgb.py:221-295): [conv3x3, norm, ReLU] (+ ResidualBlock after every stage but the first)
        x = self._rgb_extractor(patches, m.rgb_extractor.features, m.extractor_normalization_type, A, A)
        if m.use_refinement:
            self._hier_head(x, m.segmentation_head.base_head, m.segmentation_head)
        else:
            self._hier_head(x, m.segmentation_head, None)
            self.h_aux.pop("shared_features", None)                  # the V2 head returns 4 aux tensors (..._unet.py:836-841)
        if m.aux_outputs != "none":
            self.h_aux["roi_patches"] = roi_patch

    def _build_boundary_refiner(self):
        """BoundaryRefinementModule (..._refinement.py:58-149) on the full [N,3,mh,mw] logits, in place."""
        m, p, L = self.m, self.plan, self.plan.lib
        N = self.N
        mh, mw = m.mask_size
        br = m.segmentation_head.boundary_refiner
        A = self.act_ref
        edges = p.f32(N, mh, mw)
        minmax = torch.zeros(2, dtype=torch.int32, device=self.dev)
        p.keep.append(minmax)
        p.add("boundary_edges", L.his_boundary_edges, self.logits.data_ptr(), N, mh, mw, edges.data_ptr(), minmax.data_ptr())
        ec = br.edge_conv
        c = ec[0].weight.shape[0]
        x = p.act(N, mh, mw, c)
        if self._is_bn(ec[1]):
            sc, sh = fold_bn(ec[0].bias, ec[1], c)
            p.conv_direct(self.logits, 1, N, mh, mw, 3, 0, self.direct_w(ec[0].weight), p.const(sc), p.const(sh), c, 3, 1, 1,
                          A, self.beta, out=x)
        else:
            sc, sh = fold_bn(ec[0].bias, None, c)
            p.conv_direct(self.logits, 1, N, mh, mw, 3, 0, self.direct_w(ec[0].weight), p.const(sc), p.const(sh), c, 3, 1, 1,
                          ACT["none"], self.beta, out=x)
            x = self.layernorm(x, ec[1], A)
        x = self.conv(x, ec[3], ec[4], A)
        corr = p.f32(N, 3, mh, mw)
        self.conv(x, ec[6], None, ACT["none"], out_f32=corr)
        p.add("boundary_blend", L.his_boundary_blend, self.logits.data_ptr(), corr.data_ptr(), edges.data_ptr(), minmax.data_ptr(),
              p.const(br.blend_weight.detach().reshape(1)).data_ptr(), N, mh, mw, self.logits.data_ptr())

    def _rgb_extractor(self, x: Act, features, norm_kind: str, act_stage: int, act_rb: int) -> Act:
        """RGBFeatureExtractor (rgb.py:221-295): [conv3x3, norm, act] per stage, a ResidualBlock after every stage but the first."""
        m = self.m
        saved = m.normalization_type
        m.normalization_type = norm_kind             # _tail_ok/_aux_fusable look at the current norm kind
        conv = None
        for mod in features:
            if isinstance(mod, nn.Conv2d):
                conv = mod
            elif isinstance(mod, pt.NORM_MODULES):
                x = self.conv(x, conv, mod, act_stage)
            elif isinstance(mod, pt.ResidualBlockParams):
                x = self.residual_block(x, mod, act_rb)
        m.normalization_type = saved
        return x

    def _build_head_multiscale(self):
        """MultiScaleRGBSegmentationModel.forward (rgb.py:866-922)."""
        m, p, L = self.m, self.plan, self.plan.lib
        if hasattr(m, "_bad_fusion"):
            raise ValueError(f"Unknown fusion method: {m._bad_fusion}")
        N, B, H, W = self.Nc, self.B, self.H, self.W
        fh, fw = m.feature_size
        A_stage = self.act_rgb
        bn = m.extractor_normalization_type.lower() in ("batch", "batchnorm", "batchnorm2d")
        A_rb = self.act_rgb if bn else ACT["relu"]       # rgb.py:267-276: only the batchnorm branch forwards the activation
        concat = m.fusion_method == "concat"
        ns = len(m.scales)
        fused_in = p.act(N, fh, fw, 256 * ns)
        roi_patch = None
        for i, sc in enumerate(m.scales):
            rs = int(m.roi_sizes[sc])
            ra = m.roi_aligns[sc]
            patches = p.act_zeroed(N, rs, rs, 3)
            want_patch = i == 0 and m.aux_outputs != "none"
            if want_patch:
                roi_patch = p.f32(N, 3, rs, rs)
            p.add("roi_align_rgb", L.his_roi_align, self.images.data_ptr(), 0, 3 * H * W, H * W, W, 1, B, 3, H, W, self.h_rois.data_ptr(), N, rs, rs,
                  float(ra.spatial_scale_h), float(ra.spatial_scale_w), 1 if ra.aligned else 0, patches.ptr, patches.cs,
                  roi_patch.data_ptr() if want_patch else None, self.S)
            slot = fused_in.slice(256 * i, 256)
            f = self._rgb_extractor(patches, m.rgb_extractors[sc].features, m.extractor_normalization_type, A_stage, A_rb)
            # F.interpolate to 28x28 (identity when the scale already is 28) straight into the scale's concat slot
            p.add("resize_bilinear_half", L.his_resize_bilinear_half, f.ptr, N, rs, rs, 256, f.cs, fh, fw, slot.ptr, slot.cs, self.S)
        # fusion + projection: 'sum' / 'adaptive' are the 1x1 projection applied to the weighted sum of the scales, i.e. a 1x1 conv
        # over the concat buffer whose weight is [w_0*W | w_1*W | ...] -- no extra pass over the features
        proj = m.fusion_proj[0]
        if concat:
            conv_mod = proj
        else:
            w = torch.ones(ns) if m.fusion_method == "sum" else torch.softmax(m.fusion_weights.detach().float().cpu(), 0)
            conv_mod = nn.Conv2d(256 * ns, 256, 1)
            with torch.no_grad():
                conv_mod.weight.copy_(torch.cat([proj.weight.detach().float().cpu() * w[i] for i in range(ns)], 1))
                conv_mod.bias.copy_(proj.bias.detach().float().cpu())
        saved = m.normalization_type
        m.normalization_type = m.extractor_normalization_type
        x = self.conv(fused_in.widen(256 * ns), conv_mod, m.fusion_proj[1], A_stage)
        m.normalization_type = saved
        # the head works at 28x28 whatever the ROI scales are, and with ReLU whatever the extractors use (the reference builds
        # HierarchicalSegmentationHeadUNetV2 without forwarding the activation, rgb.py:854-860)
        saved_roi, saved_acts = m.roi_size, (self.act_rgb, self.act_ref)
        m.roi_size = m.feature_size
        self.act_rgb = self.act_ref = ACT["relu"]
        self._hier_head(x, m.segmentation_head, None)
        m.roi_size = saved_roi
        self.act_rgb, self.act_ref = saved_acts
        self.h_aux.pop("shared_features", None)
        if m.aux_outputs != "none":
            self.h_aux["roi_patches"] = roi_patch

    def _hier_head(self, feats: Act, bh, head):
        """ExtendedHierarchicalSegmentationHeadUNetV2 / HierarchicalSegmentationHeadUNetV2 (``bh``) and, when ``head`` is the
        refined wrapper, its contour / distance branches (..._refinement.py:734-804)."""
        m, p, L = self.m, self.plan, self.plan.lib
        N = self.Nc
        rh, rw = m.roi_size
        mh, mw = m.mask_size
        aux_level = m.aux_outputs
        A_rgb, A_ref = self.act_rgb, self.act_ref
        # --- shared trunk (..._refinement.py:479-487)
        sf = bh.shared_features
        shared = self.conv(feats, sf[0], sf[1], A_ref)
        shared = self.residual_block(shared, sf[4], A_ref)
        shared_nchw = p.f32(N, 256, rh, rw) if (aux_level == "full" and self._aux_fusable(256)) else None
        shared = self.residual_block(shared, sf[6], A_ref, aux_f32=shared_nchw)
        # --- bg/fg EnhancedUNet -> low-res logits (fp32 NCHW)
        low = p.f32(N, 2, rh, rw)
        self._enhanced_unet(shared, bh.bg_vs_fg_unet, low)
        # --- upsample_bg_fg (fused) + resize + softmax happens inside the combine
        up = bh.upsample_bg_fg
        bgfg_nat = p.f32(N, 2, 2 * rh, 2 * rw)
        if self._is_bn(up[1]):
            s32, t32 = fold_bn(up[0].bias, up[1], 32)
            p.add("upsample_bgfg", L.his_upsample_bgfg, low.data_ptr(), N, rh, rw, p.const(up[0].weight).data_ptr(), p.const(s32).data_ptr(),
                  p.const(t32).data_ptr(), p.const(up[3].weight.reshape(2, 32)).data_ptr(), p.const(up[3].bias).data_ptr(), A_ref, self.beta,
                  bgfg_nat.data_ptr())
        else:   # LayerNorm2d over (32, 2rh, 2rw) needs the whole sample: ConvT -> LN+act -> 1x1
            u = p.act(N, 2 * rh, 2 * rw, 32)
            p.add("convT2x2_small", L.his_convT2x2_small, low.data_ptr(), N, 2, rh, rw, p.const(up[0].weight).data_ptr(),
                  p.const(up[0].bias).data_ptr(), 32, u.ptr, u.cs, self.S)
            u = self.layernorm(u, up[1], A_ref)
            self.conv(u, up[3], None, ACT["none"], out_f32=bgfg_nat)
        bgfg = self.to_mask_size(bgfg_nat)
        # --- fg_gate (..._refinement.py:537-545,569-570): 1x1 2->64 (from the fp32 logits), 64->128, 128->256, sigmoid, * shared
        fg = bh.fg_gate
        g1 = p.act(N, rh, rw, 64)
        sc, sh = fold_bn(fg[0].bias, None, 64)
        p.conv_direct(low, 1, N, rh, rw, 2, 0, self.direct_w(fg[0].weight), p.const(sc), p.const(sh), 64, 1, 1, 0,
                      A_ref, self.beta, out=g1)
        g2 = self.conv(g1, fg[3], None, A_ref)
        gate_nchw = None
        if aux_level == "full":
            if self._aux_fusable(g2.C):
                gate_nchw = p.f32(N, 256, rh, rw)
                self.h_aux["fg_attention"] = gate_nchw
            else:
                gate = self.conv(g2, fg[5], None, ACT["sigmoid"])
                self.h_aux["fg_attention"] = self.export_nchw(gate)
        gated = self.conv(g2, fg[5], None, ACT["sigmoid"], res=shared, res_mode=RES_MUL, aux_f32=gate_nchw)
        # --- target vs non-target branch
        tb = bh.target_vs_nontarget_branch
        fuse_sa = m.use_attention_module and self._is_bn_mode()
        stats = p.f32(N, rh, rw, 2) if m.use_attention_module else None
        tn_nat = p.f32(N, 2, 2 * rh, 2 * rw)
        x = self.residual_block(gated, tb[0], A_ref, stats_out=stats if fuse_sa else None)
        if m.use_attention_module:
            kk = tb[1].conv.weight.shape[-1]
            wsa = p.const(tb[1].conv.weight.reshape(2, kk, kk))
            if fuse_sa:
                # SpatialAttentionModule without a pass over the 256-channel tensor: channel mean / max come from the residual
                # block's epilogue, the gate sigmoid(conv7x7(stats)) scales the rows of the ConvT that consumes x
                # (conv(g*x) = g*conv(x) for a per-pixel g and a 1x1-per-pixel transposed conv)
                sgate = p.f32(N, rh, rw)
                p.add("spatial_gate", L.his_spatial_gate, stats.data_ptr(), N, rh, rw, wsa.data_ptr(), kk, sgate.data_ptr())
                x = self.conv(x, tb[3], tb[4], A_ref, row_scale=sgate)       # ConvT 256->128 k2s2 + norm + act
            else:
                sa = p.act(N, rh, rw, 256)
                p.add("spatial_attention", L.his_spatial_attention, x.ptr, N, rh, rw, 256, x.cs, wsa.data_ptr(), kk, stats.data_ptr(), sa.ptr, sa.cs, self.S)
                x = self.conv(sa, tb[3], tb[4], A_ref)                       # ConvT 256->128 k2s2 + norm + act
            ca = tb[6]
            r = ca.fc1.weight.shape[0]
            parts = L.his_pool_sum_parts(N, x.H * x.W, 128)
            pool = p.f32(N, parts, 128); gate_c = p.f32(N, 128)
            p.add("pool_sum", L.his_pool_sum, x.ptr, N, x.H * x.W, 128, x.cs, pool.data_ptr(), self.S)
            p.add("se_gate", L.his_se_gate, pool.data_ptr(), parts, N, x.H * x.W, 128, r, p.const(ca.fc1.weight.reshape(r, 128)).data_ptr(), None,
                  p.const(ca.fc2.weight.reshape(128, r)).data_ptr(), None, A_ref, self.beta, p.f32(N, r).data_ptr(), gate_c.data_ptr())
            last_rb, tail = tb[8], tb[9]
            if self._is_bn_mode() and self._tail_ok(A_ref):
                # ChannelAttention without a pass over the mask-resolution tensor: x*g feeds the residual block twice -- as conv1's
                # input (gate folded into per-ROI weights: conv(x*g) = conv_{w*g}(x)) and as the residual (scaled when it is read)
                c_mid = last_rb.conv1.weight.shape[0]
                nt, bnn = ctypes.c_int(), ctypes.c_int()
                L.his_conv_gemm_tile_n(c_mid, ctypes.byref(nt), ctypes.byref(bnn))
                wscr = torch.empty(N * 9 * nt.value * bnn.value * round_up(c_mid, 64) * (2 if self.split else 1), dtype=torch.float16, device=self.dev)
                p.keep.append(wscr)
                t = self.conv(x, last_rb.conv1, last_rb.norm1, A_ref, in_gate=(gate_c, wscr))
                self.conv(t, last_rb.conv2, last_rb.norm2, A_ref, res=x, res_mode=RES_ADD, res_scale=gate_c, tail=(tail, False, tn_nat))
                ca_fused = True
            else:
                p.add("scale_channels", L.his_scale_channels, x.ptr, x.cs, gate_c.data_ptr(), N, x.H * x.W, 128, x.ptr, x.cs, self.S)
                ca_fused = False
        else:
            x = self.conv(x, tb[2], tb[3], A_ref)
            last_rb, tail = tb[6], tb[7]
            ca_fused = False
        if ca_fused:
            pass
        elif self._tail_ok(A_ref):
            self.residual_block(x, last_rb, A_ref, tail=(tail, False, tn_nat))
        else:
            x = self.residual_block(x, last_rb, A_ref)
            self.conv(x, tail, None, ACT["none"], out_f32=tn_nat)
        tn = self.to_mask_size(tn_nat)
        p.add("head_combine", L.his_head_combine, bgfg.data_ptr(), tn.data_ptr(), N, mh, mw, self.h_logits.data_ptr())
        if head is not None and getattr(m, "use_progressive_upsampling", False):
            # ProgressiveUpsamplingDecoder (..._refinement.py:152-215) re-decodes the shared features at 4x the ROI resolution and
            # REPLACES the hierarchical logits (:753-756; it wins over the sub-pixel decoder)
            st = head.progressive_decoder.stages
            x = shared
            for i in (0, 1):
                x = self.conv_transpose4(x, st[i][0], st[i][1], A_ref)
                x = self.residual_block(x, st[i][3], A_ref)
            nat = self.h_logits if (4 * rh, 4 * rw) == (mh, mw) else p.f32(N, 3, 4 * rh, 4 * rw)
            self.conv(x, st[2], None, ACT["none"], out_f32=nat)
            if nat is not self.h_logits:
                p.add("resize_bilinear", L.his_resize_bilinear_f32, nat.data_ptr(), N * 3, 4 * rh, 4 * rw, mh, mw, self.h_logits.data_ptr())
        elif head is not None and getattr(m, "use_subpixel_conv", False):
            # SubPixelDecoder (..._refinement.py:218-252) re-decodes the shared features and REPLACES the hierarchical logits (:753-763):
            # conv3x3 256 -> 12 with the fp32 NCHW copy written by the GEMM epilogue, PixelShuffle(2), bilinear to the mask size
            sp = head.subpixel_decoder
            conv16 = nn.Conv2d(256, 16, 3, padding=1)              # 12 real output channels, padded to the GEMM's minimum N tile
            with torch.no_grad():
                conv16.weight.zero_(); conv16.bias.zero_()
                conv16.weight[:12].copy_(sp.conv.weight.detach().float().cpu()); conv16.bias[:12].copy_(sp.conv.bias.detach().float().cpu())
            sub16 = p.f32(N, 16, rh, rw)
            self.conv(shared, conv16, None, ACT["none"], aux_f32=sub16)
            shuf = self.h_logits if (2 * rh, 2 * rw) == (mh, mw) else p.f32(N, 3, 2 * rh, 2 * rw)
            p.add("pixel_shuffle", L.his_pixel_shuffle2_f32, sub16.data_ptr(), N, 3, 16, rh, rw, shuf.data_ptr())
            if shuf is not self.h_logits:
                p.add("resize_bilinear", L.his_resize_bilinear_f32, shuf.data_ptr(), N * 3, 2 * rh, 2 * rw, mh, mw, self.h_logits.data_ptr())

        if aux_level != "none":
            self.h_aux.update({"bg_fg_logits": bgfg, "bg_fg_logits_low": low, "target_nontarget_logits": tn})
            if aux_level == "full":
                self.h_aux["shared_features"] = shared_nchw if shared_nchw is not None else self.export_nchw(shared)
        # --- auxiliary branches (..._refinement.py:772-802); computed like the reference forward does
        if head is None or aux_level == "none":
            return          # contour / distance branches only feed aux outputs: dead code when none is returned (the exported ONNX
                            # graph, whose outputs are the masks alone, prunes them the same way)
        merged = None
        if m.use_contour_detection and m.use_distance_transform and self._is_bn_mode():
            # both aux branches open with a 3x3 conv + BatchNorm + activation of the shared features (256->64 and 256->128): one
            # N = 192 GEMM reads `shared` once and runs in the tensor-bound per-tap mode instead of two narrow layers
            c0, d0 = head.contour_branch.contour_branch, head.distance_decoder.distance_head
            both = nn.Conv2d(256, 192, 3, padding=1)
            bn = nn.BatchNorm2d(192, eps=c0[1].eps)
            with torch.no_grad():
                both.weight.copy_(torch.cat([c0[0].weight.detach().float().cpu(), d0[0].weight.detach().float().cpu()]))
                both.bias.copy_(torch.cat([c0[0].bias.detach().float().cpu(), d0[0].bias.detach().float().cpu()]))
                for name in ("weight", "bias", "running_mean", "running_var"):
                    getattr(bn, name).copy_(torch.cat([getattr(c0[1], name).detach().float().cpu(), getattr(d0[1], name).detach().float().cpu()]))
            if c0[1].eps == d0[1].eps:
                merged = self.conv(shared, both, bn, A_ref)
        if m.use_contour_detection:
            cb = head.contour_branch.contour_branch
            c = merged.slice(0, 64) if merged is not None else self.conv(shared, cb[0], cb[1], A_ref)
            c_low = p.f32(N, 1, rh, rw)
            if self._tail_ok(A_ref):
                self.conv(c, cb[3], cb[4], A_ref, tail=(cb[6], True, c_low))
            else:
                c = self.conv(c, cb[3], cb[4], A_ref)
                self.conv(c, cb[6], None, ACT["sigmoid"], out_f32=c_low)
            contours = self.to_mask_size(c_low)
            if aux_level != "none":
                self.h_aux["contours"] = contours
        if m.use_distance_transform:
            dd = head.distance_decoder
            dh = dd.distance_head
            d = merged.slice(64, 128) if merged is not None else self.conv(shared, dh[0], dh[1], A_ref)
            d_low = p.f32(N, 1, rh, rw)
            if self._tail_ok(A_ref):
                self.residual_block(d, dh[3], A_ref, tail=(dh[4], False, d_low))
            else:
                d = self.residual_block(d, dh[3], A_ref)
                self.conv(d, dh[4], None, ACT["none"], out_f32=d_low)
            m_low = p.f32(N, 1, rh, rw)
            thr = p.const(dd.threshold.detach().reshape(1))
            p.add("distance_mask", L.his_map_f32, d_low.data_ptr(), d_low.numel(), 1, thr.data_ptr(), m_low.data_ptr())
            dmask, dmap = self.to_mask_size(m_low), self.to_mask_size(d_low)
            if aux_level != "none":
                self.h_aux["distance_mask"], self.h_aux["distance_map"] = dmask, dmap

    def _build_guided_head(self, comb_in: Act, roi_feat: torch.Tensor, roi_patch: Optional[torch.Tensor]):
        """PretrainedUNetGuidedSegmentationHead.forward (rgb.py:125-218): the UNet's foreground probability is the 257th input
        channel and modulates the optional attention; a direct 3-class classifier; aux log-probabilities from the resized mask."""
        m, p, L = self.m, self.plan, self.plan.lib
        N = self.Nc
        rh, rw = m.roi_size
        mh, mw = m.mask_size
        A_rgb, A_ref = self.act_rgb, self.act_ref
        hd = m.segmentation_head
        fg_slot = comb_in.slice(256, 1)
        fg_low = p.f32(N, 1, rh, rw)
        p.add("sigmoid_channel", L.his_sigmoid_channel, roi_feat.data_ptr(), N, 2, rh * rw, 1, fg_slot.ptr, fg_slot.cs, fg_low.data_ptr(), self.S)
        x = self.conv(comb_in.widen(257), hd.input_adjust, None, ACT["none"])
        fp = hd.feature_processor
        x = self.conv(x, fp[0], fp[1], A_rgb)
        x = self.residual_block(x, fp[4], A_ref)
        x = self.residual_block(x, fp[6], A_ref)
        attention = None
        if m.use_attention_module:
            am = hd.attention_module
            a = self.conv(x, am[0], None, A_rgb)
            attention = p.f32(N, 1, rh, rw)
            self.conv(a, am[2], None, ACT["sigmoid"], out_f32=attention)
            p.add("scale_pixels", L.his_scale_pixels, x.ptr, x.cs, attention.data_ptr(), fg_low.data_ptr(), N * rh * rw, 256, x.ptr, x.cs, self.S)
        fc = hd.final_classifier
        y = self.conv(x, fc[0], fc[1], A_rgb)
        same = (rh, rw) == (mh, mw)
        low = self.h_logits if same else p.f32(N, 3, rh, rw)
        self.conv(y, fc[3], None, ACT["none"], out_f32=low)
        if not same:
            p.add("resize_bilinear", L.his_resize_bilinear_f32, low.data_ptr(), N * 3, rh, rw, mh, mw, self.h_logits.data_ptr())
        if m.aux_outputs != "none":
            mask_up, fg_up, bgfg = p.f32(N, 1, mh, mw), p.f32(N, 1, mh, mw), p.f32(N, 2, mh, mw)
            p.add("guided_aux", L.his_guided_aux, roi_feat.data_ptr(), N, 2, 1, rh, rw, mh, mw, mask_up.data_ptr(), fg_up.data_ptr(), bgfg.data_ptr())
            self.h_aux.update({"bg_fg_logits": bgfg, "target_nontarget_logits": self.h_logits[:, 1:3], "fg_prob": fg_up,
                               "pretrained_bg_fg_mask": mask_up, "attention": attention, "roi_features": roi_feat, "roi_patches": roi_patch})

    def _enhanced_unet(self, x: Act, u: pt.EnhancedUNetParams, low_out: torch.Tensor):
        """EnhancedUNet.forward (..._unet.py:375-417).  Skip concats are buffer layouts: decoder level i reads
        [ConvT output | encoder skip] from one [N,h,w,2c] buffer."""
        p, L = self.plan, self.plan.lib
        A = self.act_rgb
        d, ch, N = u.depth, u.channels, x.N
        # spatial size per level
        hw = [(x.H, x.W)]
        for i in range(1, d):
            hw.append((hw[-1][0] // 2, hw[-1][1] // 2))
        cats = {}
        for lvl in range(d - 1):                                     # level lvl has ch[lvl+1] channels
            c = ch[lvl + 1]
            if hw[lvl] != (hw[lvl + 1][0] * 2, hw[lvl + 1][1] * 2):
                raise NotImplementedError("EnhancedUNet with odd intermediate sizes (bilinear re-alignment of the skip, "
                                          "..._unet.py:408) is not implemented; use roi sizes divisible by 2**(depth-1)")
            cats[lvl] = p.act(N, hw[lvl][0], hw[lvl][1], 2 * c)
        for i in range(d):
            e = u.encoders[i]
            dst = cats[i].slice(ch[i + 1], ch[i + 1]) if i < d - 1 else None
            if i == 0:
                x = self.conv(x, e[0], e[1], A)
                x = self.residual_block(x, e[3], A)
                x = self.residual_block(x, e[4], A, out=dst)
            else:
                x = self.residual_block(x, e[0], A)
                x = self.residual_block(x, e[1], A)
                x = self.conv(x, e[2], e[3], A, out=dst)
            if i < d - 1:
                pooled = p.act(N, hw[i + 1][0], hw[i + 1][1], ch[i + 1])
                p.add("maxpool2", L.his_maxpool2, x.ptr, N, x.H, x.W, x.C, x.cs, pooled.ptr, pooled.cs, self.S)
                x = pooled
        b = u.bottleneck
        a = self.residual_block(x, b[0], A)
        a = self.residual_block(a, b[1], A)
        a = self.conv(a, b[2], b[3], A)
        a = self.conv(a, b[5], None, ACT["sigmoid"])
        x = self.conv(x, u.bottleneck_conv, None, ACT["none"], res=a, res_mode=RES_MUL)
        for i in range(d - 1):
            lvl = d - 2 - i
            c = ch[lvl + 1]
            self.conv(x, u.upconvs[i], None, ACT["none"], out=cats[lvl].slice(0, c))
            dec = u.decoders[i]
            x = self.conv(cats[lvl].widen(2 * c), dec[0], dec[1], A)
            x = self.residual_block(x, dec[3], A)
            x = self.residual_block(x, dec[4], A)
        f = u.final
        if self._tail_ok(A):
            self.conv(x, f[0], f[1], A, tail=(f[3], False, low_out))
        else:
            x = self.conv(x, f[0], f[1], A)
            self.conv(x, f[3], None, ACT["none"], out_f32=low_out)


# ======================================================================================= factory
def create_rgb_hierarchical_model(roi_size: Union[int, Tuple[int, int]] = 28, mask_size: Union[int, Tuple[int, int]] = 56,
                                  multi_scale: bool = False, activation_function: str = "relu", activation_beta: float = 1.0,
                                  normalization_type: str = "layernorm2d", normalization_groups: int = 8, **kwargs) -> nn.Module:
    """Same signature and kwarg handling as the reference factory (rgb.py:925-1026)."""
    use_attention_module = kwargs.pop("use_attention_module", False)
    use_boundary_refinement = kwargs.pop("use_boundary_refinement", False)
    use_progressive_upsampling = kwargs.pop("use_progressive_upsampling", False)
    use_subpixel_conv = kwargs.pop("use_subpixel_conv", False)
    use_contour_detection = kwargs.pop("use_contour_detection", False)
    use_distance_transform = kwargs.pop("use_distance_transform", False)
    use_pretrained_unet = kwargs.pop("use_pretrained_unet", False)
    pretrained_weights_path = kwargs.pop("pretrained_weights_path", "")
    freeze_pretrained_weights = kwargs.pop("freeze_pretrained_weights", False)
    use_full_image_unet = kwargs.pop("use_full_image_unet", False)
    if multi_scale:          # rgb.py:961-975
        return MultiScaleRGBSegmentationModel(
            roi_sizes=kwargs.get("roi_sizes", {"scale1": 56, "scale2": 42, "scale3": 28}), mask_size=mask_size,
            fusion_method=kwargs.get("fusion_method", "concat"), use_attention_module=use_attention_module,
            normalization_type=normalization_type, normalization_groups=normalization_groups,
            activation_function=activation_function, activation_beta=activation_beta)
    kwargs.pop("roi_sizes", None)
    kwargs.pop("fusion_method", None)
    if use_pretrained_unet and not use_full_image_unet:
        # rgb.py:442-561: the reference's own constructor dies on an undefined name (`kwargs`, :497) -> nothing to mirror
        raise NotImplementedError("HierarchicalRGBSegmentationModelWithPretrainedUNet (ROI-level UNet) cannot be constructed in the "
                                  "reference either (NameError at hierarchical_segmentation_rgb.py:497); not implemented")
    if not use_pretrained_unet:       # rgb.py:1011-1026: activation kwargs are NOT forwarded to this model
        pt.check_activation(activation_function)
        return HierarchicalRGBSegmentationModel(
            roi_size=roi_size, mask_size=mask_size, use_attention_module=use_attention_module,
            use_boundary_refinement=use_boundary_refinement, use_progressive_upsampling=use_progressive_upsampling,
            use_subpixel_conv=use_subpixel_conv, use_contour_detection=use_contour_detection,
            use_distance_transform=use_distance_transform, normalization_type=normalization_type,
            normalization_groups=normalization_groups)
    return HierarchicalRGBSegmentationModelWithFullImagePretrainedUNet(
        roi_size=roi_size, mask_size=mask_size, pretrained_weights_path=pretrained_weights_path,
        use_attention_module=use_attention_module, freeze_pretrained_weights=freeze_pretrained_weights,
        use_boundary_refinement=use_boundary_refinement, use_progressive_upsampling=use_progressive_upsampling,
        use_subpixel_conv=use_subpixel_conv, use_contour_detection=use_contour_detection,
        use_distance_transform=use_distance_transform, normalization_type=normalization_type,
        normalization_groups=normalization_groups, activation_function=activation_function, activation_beta=activation_beta, **kwargs)
