"""ctypes binding of ``libhis_b200.so`` (include/his_b200.h) and its in-tree nvcc build.

There is no fallback: if the library cannot be built/loaded, or a call fails, a
``HisError`` is raised.  Nothing in this package routes compute through torch ops.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from ctypes import POINTER, c_char_p, c_float, c_int, c_longlong, c_uint, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(_HERE, "csrc")
LIB_PATH = os.path.join(_HERE, "libhis_b200.so")
SOURCES = ["kernels.cu", "conv_gemm_sm100.cu", "conv_gemm_sm100_split.cu", "post.cu", "post_stencil.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC"]


class HisError(RuntimeError):
    pass


ABI_VERSION = 200          # his_version() of the library this binding was written against (csrc/kernels.cu)
HEADER = os.path.join(os.path.dirname(_HERE), "include", "his_b200.h")


HASH_PATH = LIB_PATH + ".srchash"


def _source_hash() -> str:
    """Content hash of everything the library is built from (mtimes do not survive a snapshot copy to the GPU box)."""
    import hashlib
    h = hashlib.sha256()
    srcs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h")))
    if os.path.exists(HEADER):
        srcs.append(HEADER)
    for f in srcs:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _stale() -> bool:
    if not os.path.exists(LIB_PATH) or not os.path.exists(HASH_PATH):
        return True
    with open(HASH_PATH) as fh:
        return fh.read().strip() != _source_hash()


def build(force: bool = False, verbose: bool = False) -> str:
    """Compiles csrc/*.cu for sm_100a into libhis_b200.so next to this file (nvcc cross-compiles without a GPU)."""
    if not force and not _stale():
        return LIB_PATH
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    if not os.path.exists(nvcc):
        nvcc = "nvcc"
    objs = []
    procs = []
    os.makedirs(os.path.join(_HERE, "build"), exist_ok=True)
    for src in SOURCES:
        path = os.path.join(CSRC, src)
        if not os.path.exists(path):
            continue
        obj = os.path.join(_HERE, "build", src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, "-c", path, "-o", obj]
        if verbose:
            print(" ".join(cmd))
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0:
            raise HisError(f"nvcc failed on {src}:\n{out.decode(errors='replace')}")
    cmd = [nvcc, "-shared", "-o", LIB_PATH, *objs, "-lcudart"]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)
    if r.returncode != 0:
        raise HisError(f"link failed:\n{r.stdout.decode(errors='replace')}")
    with open(HASH_PATH, "w") as fh:
        fh.write(_source_hash())
    return LIB_PATH


_P = c_void_p
_F = POINTER(c_float)
_LL = c_longlong

# name -> argtypes (restype is int unless listed in _RESTYPES); mirrors include/his_b200.h
SIGNATURES = {
    "his_last_error": [],
    "his_version": [],
    "his_roi_align": [_P, c_int, _LL, _LL, _LL, _LL, c_int, c_int, c_int, c_int, _P, c_int, c_int, c_int, c_float, c_float,
                      c_int, _P, c_int, _P, c_int, _P],
    "his_roi_align_fused": [_P, c_int, c_float, c_float, c_int, _P, c_int, _P, _P, c_int, c_float, c_float, c_int, _P, c_int, _P,
                            c_int, c_int, c_int, _P, c_int, c_int, c_int, c_int, _P],
    "his_conv_gemm_tile_n": [c_int, POINTER(c_int), POINTER(c_int)],
    "his_conv_gemm_create": [POINTER(c_void_p), _P, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, c_int, c_int, _P, c_int,
                             _P, c_int, c_int, c_int, c_float, c_int, c_int],
    "his_conv_gemm_set_tail": [_P, _P, c_float, c_float, c_int, c_int, _P, c_int],
    "his_conv_gemm_set_aux": [_P, _P],
    "his_conv_gemm_set_ln_partials": [_P, _P, POINTER(c_int)],
    "his_conv_gemm_work_items": [_P],
    "his_conv_gemm_set_row_ops": [_P, _P, _P],
    "his_conv_gemm_set_res_scale": [_P, _P],
    "his_conv_gemm_set_upsampled_input": [_P, _P, c_int, c_int],
    "his_conv_gemm_can_fuse_upsample": [c_int, c_int, c_int, c_int, c_int],
    "his_conv_gemm_set_image_weights": [_P, _P],
    "his_conv_gemm_run": [_P, _P],
    "his_conv_gemm_destroy": [_P],
    "his_conv_gemm_issued_macs": [_P],
    "his_conv_direct": [_P, c_int, _P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, c_int, c_int, c_int, c_int, c_int, c_int,
                        c_float, c_int, _P, c_int, _P, c_int, _P, c_int, _P],
    "his_depthwise_conv": [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, c_int, c_int, c_int, _P, c_int, _P, c_int, _P],
    "his_pool_sum": [_P, c_int, c_int, c_int, c_int, _P, c_int, _P],
    "his_se_gate": [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P, c_int, c_float, _P, _P, _P],
    "his_depthwise_pool_parts": [c_int, c_int, c_int, c_int, c_int, c_int],
    "his_pool_sum_parts": [c_int, c_int, c_int],
    "his_scale_weights": [_P, _P, c_int, _LL, c_int, c_int, _P, c_int, _P],
    "his_scale_channels": [_P, c_int, _P, c_int, c_int, c_int, _P, c_int, c_int, _P],
    "his_layernorm2d_parts": [c_int, c_int, c_int],
    "his_layernorm2d_act": [_P, c_int, c_int, c_int, c_int, _P, _P, c_float, c_int, c_float, c_int, _P, c_int, _P, c_int, _P, c_int, c_int, _P],
    "his_groupnorm_parts": [c_int, c_int, c_int],
    "his_groupnorm_act": [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_float, c_int, c_float, c_int, _P, c_int, _P, _P, c_int, c_int, _P],
    "his_fgaware_norm_act": [_P, c_int, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_float, c_int, c_float, c_int, _P, c_int, _P, _P, c_int,
                             c_int, _P],
    "his_convT2x2_small": [_P, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, c_int, c_int, _P],
    "his_spatial_attention": [_P, c_int, c_int, c_int, c_int, c_int, _P, c_int, _P, _P, c_int, c_int, _P],
    "his_spatial_gate": [_P, c_int, c_int, c_int, _P, c_int, _P, _P],
    "his_maxpool2": [_P, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P],
    "his_resize_nearest": [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P],
    "his_resize_bilinear_half": [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P],
    "his_resize_bilinear_f32": [_P, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "his_upsample_bgfg": [_P, c_int, c_int, c_int, _P, _P, _P, _P, _P, c_int, c_float, _P, _P],
    "his_head_combine": [_P, _P, c_int, c_int, c_int, _P, _P],
    "his_map_f32": [_P, _LL, c_int, _P, _P, _P],
    "his_depth_to_space2_half": [_P, c_int, c_int, c_int, c_int, c_int, _P, c_int, c_int, _P],
    "his_pixel_shuffle2_f32": [_P, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "his_boundary_edges": [_P, c_int, c_int, c_int, _P, _P, _P],
    "his_boundary_blend": [_P, _P, _P, _P, _P, c_int, c_int, c_int, _P, _P],
    "his_nhwc_half_to_nchw_float": [_P, c_int, c_int, c_int, c_int, _P, c_int, _P],
    "his_sigmoid_channel": [_P, c_int, c_int, c_int, c_int, _P, c_int, _P, c_int, _P],
    "his_scale_pixels": [_P, c_int, _P, _P, _LL, c_int, _P, c_int, c_int, _P],
    "his_guided_aux": [_P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P, _P, _P],
    "his_unet_input_affine": [_P, _LL, _F, _F, _P, _P, _P],
    "his_s2d_input": [_P, c_int, c_int, c_int, _P, _P, c_int, _P],
    "his_unet_outputs": [_P, c_int, c_int, c_int, c_float, c_float, c_float, c_float, _P, _P, _P],
    "his_memset_async": [_P, c_int, _LL, _P],
    # post-processing (csrc/post.cu)
    "his_post_instance_mask": [_P, c_int, c_int, c_int, c_float, _P, _P, _P],
    "his_post_dilate_logits": [_P, c_int, c_int, c_int, c_int, _P, _P],
    "his_post_dilate_instance_mask": [_P, c_int, c_int, c_int, c_int, c_float, _P, _P, _P],
    "his_post_edge_smooth": [_P, c_int, c_int, c_int, c_float, c_float, _P, _P],
    "his_post_binary_bilateral": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_float, _P, _P, _P, _P],
    "his_post_morph_bilateral": [_P, c_int, c_int, c_int, _P, c_int, c_int, _P, _P, _P, _P],
    "his_post_paste": [_P, c_int, c_int, c_int, _P, _P, c_int, c_int, c_int, _P],
    "his_preprocess_u8": [_P, c_int, c_int, c_int, c_int, c_int, _P, _P, c_int, _P, _P],
    "his_eval_confusion": [_P, _P, c_int, c_int, c_int, c_int, _P, _P],
    # shared-memory tiled stencils (csrc/post_stencil.cu)
    "his_post_edge_smooth_tiled": [_P, c_int, c_int, c_int, c_float, c_float, _P, _P],
    "his_post_edge_directional": [_P, c_int, c_int, c_int, _P, _P],
    "his_post_edge_adaptive": [_P, c_int, c_int, c_int, _P, _P, _P, _P, _P],
    "his_post_edge_optimized": [_P, c_int, c_int, c_int, c_int, _P, _P],
    "his_post_class_masks": [_P, c_int, c_int, c_int, c_int, c_int, c_int, _P, _P],
    "his_post_bilateral_exact": [_P, c_int, c_int, c_int, _P, c_int, c_float, _P, _P],
    "his_post_bilateral_fast": [_P, c_int, c_int, c_int, _P, c_int, c_float, c_int, _P, _P, _P],
    "his_post_guided_filter": [_P, _P, c_int, c_int, c_int, c_int, c_float, _P, _P, _P, _P],
    "his_post_binary_bilateral_tiled": [_P, c_int, c_int, c_int, _P, c_int, c_int, c_float, _P, _P],
    "his_post_mask_cleanup_fused": [_P, c_int, c_int, c_int, c_float, c_float, _P, c_int, c_int, c_float, _P, _P],
    "his_post_mask_cleanup_fused_u8": [_P, c_int, c_int, c_int, c_float, c_float, _P, c_int, c_int, c_float, _P, _P],
}
_RESTYPES = {"his_last_error": c_char_p, "his_conv_gemm_issued_macs": _LL}

_lib = None


def load():
    """Loads libhis_b200.so (building it if it is missing/stale).  Raises HisError -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    try:
        path = build()
    except HisError:
        # A failed rebuild may only be papered over when the tree cannot be written at all (a read-only deployment that ships
        # the built library); a stale library next to newer sources would be called with the wrong ABI.
        if os.path.exists(LIB_PATH) and not os.access(_HERE, os.W_OK):
            path = LIB_PATH
        else:
            raise
    try:
        lib = ctypes.CDLL(path)
    except OSError as e:
        raise HisError(f"cannot load {path}: {e}") from e
    try:
        ver = int(lib.his_version())
    except AttributeError as e:
        raise HisError(f"{path} does not export his_version") from e
    if ver != ABI_VERSION:
        raise HisError(f"{path} reports ABI version {ver}, this binding needs {ABI_VERSION}: rebuild (human_instance_segmentation_b200.build(force=True))")
    for name, argtypes in SIGNATURES.items():
        try:
            fn = getattr(lib, name)
        except AttributeError as e:
            raise HisError(f"{path} does not export {name}") from e
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def on_tensor_device(fn):
    """Runs ``fn`` with the CUDA *current device* set to the device of its first CUDA tensor argument.  Kernel launches go to
    the stream the caller passes, but function attributes, occupancy queries and stream handles are resolved against the current
    device: a tensor on cuda:1 while the current device is 0 would fail with "invalid resource handle"."""
    import functools

    @functools.wraps(fn)
    def wrapped(*args, **kw):
        import torch
        for a in list(args) + list(kw.values()):
            if isinstance(a, torch.Tensor) and a.is_cuda:
                if a.device.index == torch.cuda.current_device():
                    break
                with torch.cuda.device(a.device):
                    return fn(*args, **kw)
        return fn(*args, **kw)
    return wrapped


def check(rc: int, what: str = ""):
    if rc != 0:
        msg = load().his_last_error()
        raise HisError(f"{what or 'his call'} failed with code {rc}: {msg.decode() if msg else ''}")
