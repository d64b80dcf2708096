"""Oracle (test infrastructure): deterministic, key-seeded parameter filler.

The hot path's weights are ~20-130 M floats, too large to commit as fixtures, and
the reference ships no checkpoints (``.MISSING_LARGE_BLOBS``).  Goldens therefore
use *procedural* weights: every tensor of a state dict is filled from a generator
seeded by ``crc32(key) ^ seed``, so the real reference model (in this container),
the oracle port and the CUDA model (anywhere) hold bit-identical parameters
without shipping them.  BatchNorm running statistics and affine terms are
randomised too, so that BN folding is actually exercised.
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


def fill_state_dict(sd: Dict[str, torch.Tensor], seed: int = 0, gain: float = 1.1) -> Dict[str, torch.Tensor]:
    """Returns a new state dict with the same keys/shapes/dtypes, procedurally filled.

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
