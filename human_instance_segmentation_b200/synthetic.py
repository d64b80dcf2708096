"""Synthetic inputs and procedural ("random-init") weights for benchmarks, smoke runs and parity tests.

No checkpoints exist offline (reference ``.MISSING_LARGE_BLOBS``) and 20-130 M floats are too large to commit, so
weights are generated: every tensor of a state dict is filled from a generator seeded by ``crc32(key) ^ seed``.  The
same call therefore gives bit-identical parameters to the real reference model (authoring container), the CPU oracle
and this package, on any machine.  Input generators follow SURVEY §8d.
"""
from __future__ import annotations

import zlib
from typing import Dict

import torch

# parameters the reference pins in code (hierarchical_segmentation_unet.py:1963-1971)
# and :1881-1883 (input normalisation buffers)
_PINNED = ("pretrained_unet.output_conv.", "pretrained_unet.model.norm_mean", "pretrained_unet.model.norm_std")


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator()
    g.manual_seed((zlib.crc32(key.encode()) ^ (seed * 0x9E3779B1)) & 0x7FFFFFFF)
    return g


def fill_state_dict(sd: Dict[str, torch.Tensor], seed: int = 0, gain: float = 1.1, mode: str = "stress") -> Dict[str, torch.Tensor]:
    """Returns a new state dict with the same keys/shapes/dtypes, procedurally filled.

    mode="stress" (default): randomised BatchNorm statistics/affine and gain-1.1 weights -- every folded
    constant is exercised and activations stay O(1) through ~40 layers (a deliberately hard parity case).
    mode="torch_default": the statistics of PyTorch's own random initialisation, i.e. the "random-init weights"
    of BASELINE.json: conv/linear weight and bias U(+-1/sqrt(fan_in)) (kaiming_uniform a=sqrt(5)), norm gamma=1,
    beta=0, running_mean=0, running_var=1.

    Kind is inferred from the key + shape only (never from module types), so the
    same call works on the reference model, the oracle and the product model:
      * ``*.running_var``  U[0.5,1.5];  ``*.running_mean``  N(0,0.2)
      * 1-D ``*.weight`` (BatchNorm gamma) and ``[1,C,1,1]`` LayerNorm2d gamma (told apart
        from a 1x1 conv to one channel by its ``[1,C,1,1]`` sibling bias): U[0.7,1.3]
      * 1-D ``*.bias``  N(0,0.1)
      * conv / conv-transpose / fc weights: U[-b,b], b = gain*sqrt(3/fan_in)
      * scalars (``threshold``): kept.
    """
    out = {}
    for key in sd:
        t = sd[key]
        if key.endswith("num_batches_tracked") or t.dim() == 0 or any(p in key for p in _PINNED):
            out[key] = t.clone()
            continue
        g = _gen(key, seed)
        if mode == "torch_default":
            out[key] = _torch_default(key, t, sd, g)
            continue
        if key.endswith("running_var"):
            v = torch.rand(t.shape, generator=g) + 0.5
        elif key.endswith("running_mean"):
            v = torch.randn(t.shape, generator=g) * 0.2
        elif key.endswith(".bias") or key == "bias":  # conv bias [C], BN beta [C], LayerNorm2d beta [1,C,1,1]
            v = torch.randn(t.shape, generator=g) * 0.1
        elif t.dim() == 1:
            v = torch.rand(t.shape, generator=g) * 0.6 + 0.7
        elif _is_ln2d_gamma(key, sd):
            v = torch.rand(t.shape, generator=g) * 0.6 + 0.7
        else:
            fan_in = t[0].numel() if t.dim() > 1 else t.numel()
            if _is_transposed(key, sd):
                fan_in = t.shape[0]  # ConvTranspose2d weight is [Cin, Cout, kh, kw]; k2s2 taps do not overlap
            b = gain * (3.0 / max(fan_in, 1)) ** 0.5
            v = (torch.rand(t.shape, generator=g) * 2 - 1) * b
        out[key] = v.to(t.dtype)
    return out


def _torch_default(key, t, sd, g):
    if key.endswith("running_var"):
        return torch.ones_like(t)
    if key.endswith("running_mean"):
        return torch.zeros_like(t)
    is_norm_w = (t.dim() == 1 and key.endswith(".weight")) or _is_ln2d_gamma(key, sd)
    if is_norm_w:
        return torch.ones_like(t)
    wkey = key[:-len("bias")] + "weight"
    if key.endswith(".bias"):
        w = sd.get(wkey)
        if w is None or w.dim() == 1 or _is_ln2d_gamma(wkey, sd):
            return torch.zeros_like(t)                       # norm beta
        fan_in = w[0].numel() if w.dim() > 1 else w.numel()  # torch: weight.size(1) * receptive field
        b = 1.0 / max(fan_in, 1) ** 0.5
        return ((torch.rand(t.shape, generator=g) * 2 - 1) * b).to(t.dtype)
    fan_in = t[0].numel() if t.dim() > 1 else t.numel()
    b = 1.0 / max(fan_in, 1) ** 0.5
    return ((torch.rand(t.shape, generator=g) * 2 - 1) * b).to(t.dtype)


def _is_ln2d_gamma(key: str, sd) -> bool:
    """reference model.py:18-38 LayerNorm2d keeps gamma AND beta as [1,C,1,1]; a 1x1 conv to one
    channel has the same weight shape but a [1] bias -> tell them apart by the sibling bias."""
    t = sd[key]
    if not (t.dim() == 4 and t.shape[0] == 1 and tuple(t.shape[2:]) == (1, 1) and key.endswith(".weight")):
        return False
    b = sd.get(key[:-len("weight")] + "bias")
    return b is not None and b.shape == t.shape and t.shape[1] > 1


_TRANSPOSED_HINTS = ("upconvs.", "upsample_bg_fg.0.", "upsample.0.")


def _is_transposed(key: str, sd) -> bool:
    if any(h in key for h in _TRANSPOSED_HINTS):
        return True
    # target_vs_nontarget_branch.3 (attention variant) / .2 (plain) is the ConvTranspose2d
    if "target_vs_nontarget_branch." in key:
        t = sd[key]
        return t.dim() == 4 and t.shape[2:] == (2, 2)
    return False


# ----------------------------------------------------------------------------- seeded inputs
def synth_images(seed: int, b: int, h: int, w: int) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    return torch.rand(b, 3, h, w, generator=g)


def synth_rois(seed: int, n_images: int, per_image: int) -> torch.Tensor:
    """x1,y1~U[0,.5), w,h~U[.2,.5), x2=min(x1+w,1), y2=min(y1+h,1); ``per_image`` boxes per image, grouped by image."""
    g = torch.Generator().manual_seed(seed + 7919)
    n = n_images * per_image
    xy = torch.rand(n, 2, generator=g) * 0.5
    wh = torch.rand(n, 2, generator=g) * 0.3 + 0.2
    b = torch.arange(n_images, dtype=torch.float32).repeat_interleave(per_image)[:, None]
    return torch.cat([b, xy, (xy + wh).clamp(max=1.0)], 1)


def edge_rois(n_images: int) -> torch.Tensor:
    """Edge cases: x2=1/y2=1, zero-area ROI, ROI outside [0,1], reversed box."""
    r = [[0, 0.0, 0.0, 1.0, 1.0], [n_images - 1, 0.25, 0.5, 0.25, 0.5], [0, -0.2, -0.1, 0.4, 0.6],
         [n_images - 1, 0.6, 0.7, 1.3, 1.2], [0, 0.7, 0.6, 0.3, 0.2], [0, 0.1, 0.2, 0.9, 0.95]]
    return torch.tensor(r, dtype=torch.float32)
