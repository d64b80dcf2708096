"""CPU: pins for the restated smp.Unet/timm EfficientNet (third-party code absent from the image)."""
import pytest
import torch
import torchvision

from oracle import effunet
from tests import common


@pytest.mark.parametrize("variant,n_enc", [("b0", 358), ("b1", 506), ("b3", 572), ("b7", 1198)])
def test_encoder_key_counts_match_reference_evidence(variant, n_enc):
    # reference hierarchical_segmentation_unet.py:1815-1828: "B0: ~358, B1: ~506, B3: ~572, B7: ~1198"
    net = effunet.Unet(f"timm-efficientnet-{variant}")
    keys = list(net.state_dict())
    assert sum("encoder" in k for k in keys) == n_enc
    # export_peopleseg_onnx.py:111-136: decoder key pattern and default channels
    for i, c in enumerate(effunet.DECODER_CHANNELS):
        assert net.state_dict()[f"decoder.blocks.{i}.conv1.0.weight"].shape[0] == c
    assert net.state_dict()["segmentation_head.0.weight"].shape == (1, 16, 3, 3)


@pytest.mark.parametrize("variant", ["b0", "b1", "b3", "b7"])
def test_encoder_equals_torchvision(variant):
    """torchvision implements the same MBConv topology independently (width / depth scaling, SE squeeze = block input / 4, SiLU,
    ceil depth rounding); with weights remapped by position the two encoders must agree to float rounding.  B7 (width 2.0, depth
    3.1: 55 blocks, 1198 encoder keys) is the preset of BASELINE configs[2].  torchvision builds b5-b7 with BatchNorm eps 1e-3 (the
    TF checkpoints' value); timm's non-tf `efficientnet_b7` that smp wraps keeps PyTorch's 1e-5, so the comparison sets 1e-5."""
    torch.manual_seed(0)
    ours = effunet.Unet(f"timm-efficientnet-{variant}").eval()
    tv = getattr(torchvision.models, f"efficientnet_{variant}")(weights=None).eval()
    for mod in tv.modules():
        if isinstance(mod, torch.nn.BatchNorm2d):
            mod.eps = 1e-5
    osd = {k: v for k, v in ours.state_dict().items() if k.startswith("encoder.")}
    tsd = {k: v for k, v in tv.state_dict().items() if k.startswith("features.")}
    assert len(osd) == len(tsd)
    filled = common.paramfill.fill_state_dict(osd, seed=3)
    remap = {}
    for (ko, vo), (kt, vt) in zip(filled.items(), tsd.items()):
        assert vo.shape == vt.shape, (ko, kt)
        remap[kt] = vo
    ours.load_state_dict({**ours.state_dict(), **filled})
    tv.load_state_dict({**tv.state_dict(), **remap})
    x = torch.rand(1, 3, 64, 96)
    with torch.no_grad():
        feats = ours.encoder(x)
        y = tv.features[:8](x)
    assert common.rel_err(feats[-1], y) < 1e-5
    assert [f.shape[1] for f in feats] == list(ours.encoder.out_channels)


def _functional_unet(sd, x):
    """A second, independent statement of smp 0.5.0's Unet forward over a state dict (functional ops, explicit loops), written
    from the package's published forward: encoder features [x, stem, stage1, stage2, stage4, stage6]; decoder walks them deepest
    first, each block = nearest-interpolate to the NEXT feature's spatial size, concat [upsampled, skip], conv3x3-BN-ReLU twice
    (the last block has no skip and targets the input size); head conv3x3 with bias."""
    import torch.nn.functional as F

    def bn(p, t, act=None):
        t = F.batch_norm(t, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"], False, 0.0, 1e-5)
        return F.silu(t) if act == "silu" else F.relu(t) if act == "relu" else t

    feats = [x]
    t = bn("encoder.bn1", F.conv2d(x, sd["encoder.conv_stem.weight"], None, 2, 1), "silu")
    feats.append(t)
    si = 0
    while f"encoder.blocks.{si}.0.conv_dw.weight" in sd:
        bi = 0
        while f"encoder.blocks.{si}.{bi}.conv_dw.weight" in sd:
            p = f"encoder.blocks.{si}.{bi}."
            inp = t
            wdw = sd[p + "conv_dw.weight"]
            k = wdw.shape[-1]
            if p + "conv_pwl.weight" in sd:                     # InvertedResidual
                t = bn(p + "bn1", F.conv2d(t, sd[p + "conv_pw.weight"]), "silu")
                dwbn, proj, projbn = "bn2", "conv_pwl", "bn3"
            else:                                               # DepthwiseSeparableConv
                dwbn, proj, projbn = "bn1", "conv_pw", "bn2"
            # stride: the first block of stages 1, 2, 3, 5 halves the resolution
            s = 2 if (bi == 0 and si in (1, 2, 3, 5)) else 1
            t = bn(p + dwbn, F.conv2d(t, wdw, None, s, ((s - 1) + (k - 1)) // 2, 1, wdw.shape[0]), "silu")
            g = t.mean((2, 3), keepdim=True)
            g = F.conv2d(F.silu(F.conv2d(g, sd[p + "se.conv_reduce.weight"], sd[p + "se.conv_reduce.bias"])),
                         sd[p + "se.conv_expand.weight"], sd[p + "se.conv_expand.bias"])
            t = bn(p + projbn, F.conv2d(t * torch.sigmoid(g), sd[p + proj + ".weight"]))
            if s == 1 and inp.shape[1] == t.shape[1]:
                t = t + inp
            bi += 1
        if si in (1, 2, 4, 6):
            feats.append(t)
        si += 1
    sizes = [f.shape[2:] for f in feats][::-1]                  # deepest first
    rev = feats[1:][::-1]
    t, skips = rev[0], rev[1:]
    for i in range(5):
        t = F.interpolate(t, size=tuple(sizes[i + 1]), mode="nearest")
        if i < len(skips):
            t = torch.cat([t, skips[i]], 1)
        for c in ("conv1", "conv2"):
            p = f"decoder.blocks.{i}.{c}"
            t = bn(p + ".1", F.conv2d(t, sd[p + ".0.weight"], None, 1, 1), "relu")
    return F.conv2d(t, sd["segmentation_head.0.weight"], sd["segmentation_head.0.bias"], 1, 1)


@pytest.mark.parametrize("variant,hw", [("b0", (64, 96)), ("b7", (64, 64)), ("b1", (72, 100)), ("b0", (120, 160))])
def test_unet_equals_an_independent_functional_statement(variant, hw):
    """Decoder (nearest resize to the skip's size, concat order, conv-BN-ReLU pairs, head) and whole-network wiring of
    oracle/effunet.py against a second statement written separately over the raw state dict -- including inputs whose size is not
    a multiple of 32 (120x160 of BASELINE configs[3]: the deepest maps have odd sizes and the resize is not a plain 2x)."""
    net = effunet.Unet(f"timm-efficientnet-{variant}").eval()
    sd = common.paramfill.fill_state_dict(net.state_dict(), seed=5)
    net.load_state_dict(sd)
    x = torch.rand(1, 3, *hw)
    with torch.no_grad():
        a = net(x)
        b = _functional_unet(sd, x)
    assert a.shape == (1, 1, *hw)
    assert common.rel_err(a, b) < 1e-5


def test_export_outputs_binary_mask_is_sigmoid_2x():
    from oracle import headport
    x = torch.randn(2, 1, 8, 8)
    two = torch.cat([x, -x], 1)
    inst, binary = headport.export_outputs(torch.randn(3, 3, 4, 4), two)
    assert torch.allclose(binary, torch.sigmoid(2 * x), atol=1e-6)
    assert set(inst.unique().tolist()) <= {0.0, 1.0}
