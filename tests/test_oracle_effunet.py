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


@pytest.mark.parametrize("variant", ["b0", "b1"])
def test_encoder_equals_torchvision(variant):
    """torchvision implements the same MBConv topology independently; with weights remapped by
    position the two encoders must agree to float rounding."""
    torch.manual_seed(0)
    ours = effunet.Unet(f"timm-efficientnet-{variant}").eval()
    tv = getattr(torchvision.models, f"efficientnet_{variant}")(weights=None).eval()
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


def test_export_outputs_binary_mask_is_sigmoid_2x():
    from oracle import headport
    x = torch.randn(2, 1, 8, 8)
    two = torch.cat([x, -x], 1)
    inst, binary = headport.export_outputs(torch.randn(3, 3, 4, 4), two)
    assert torch.allclose(binary, torch.sigmoid(2 * x), atol=1e-6)
    assert set(inst.unique().tolist()) <= {0.0, 1.0}
