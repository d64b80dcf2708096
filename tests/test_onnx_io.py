"""f3: released ``.onnx`` files as a weight source, without the ``onnx`` package -- protobuf wire-format reader
for ModelProto.graph.initializer, mapping onto the reference's state-dict keys, and the exporter's metadata JSON
(export_hierarchical_instance_peopleseg_onnx.py:423-441, 511-528; hed/export_onnx_advanced.py:427-457)."""
import json
import struct

import numpy as np
import pytest
import torch

import human_instance_segmentation_b200 as his
from human_instance_segmentation_b200 import onnx_io
from tests import common

KW = dict(roi_size=(16, 12), mask_size=(32, 24), use_pretrained_unet=True, use_full_image_unet=True, encoder_name="timm-efficientnet-b0",
          pretrained_weights_path="ext_extractor/best_model_b0_0.8741.pth", use_attention_module=True, use_contour_detection=True,
          use_distance_transform=True, normalization_type="batchnorm", hierarchical_base_channels=64, hierarchical_depth=3)


def _varint(v):
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(field, payload):
    return _varint((field << 3) | 2) + _varint(len(payload)) + payload


def _independent_writer(path, tensors, inputs, outputs, packed_dims=False, typed_data=False):
    """A second, independent encoder of the same wire format (what the exporter's protobuf library emits): dims as packed or
    repeated varints, payload as raw_data or as the typed repeated field."""
    chunks = []
    for name, arr in tensors.items():
        t = b""
        if packed_dims:
            t += _ld(1, b"".join(_varint(d) for d in arr.shape))
        else:
            t += b"".join(_varint(1 << 3) + _varint(d) for d in arr.shape)
        dt = {np.dtype("float32"): 1, np.dtype("int64"): 7, np.dtype("float16"): 10}[arr.dtype]
        t += _varint(2 << 3) + _varint(dt)
        if typed_data and arr.dtype == np.float32:
            t += _ld(4, arr.astype("<f4").tobytes())                     # float_data, packed
        elif typed_data and arr.dtype == np.int64:
            t += _ld(7, b"".join(_varint(int(v) & ((1 << 64) - 1)) for v in arr.flatten()))
        else:
            t += _ld(9, arr.tobytes())
        t += _ld(8, name.encode())
        chunks.append(_ld(5, t))
    # a node, like a real graph has (skipped by the reader except for its op type / name)
    chunks.append(_ld(1, _ld(1, b"images") + _ld(2, b"y") + _ld(3, b"/model/Conv") + _ld(4, b"Conv")))
    for n in inputs:
        chunks.append(_ld(11, _ld(1, n.encode()) + _ld(2, b"\x0a\x02\x08\x01")))    # + a TypeProto the reader must skip
    for n in outputs:
        chunks.append(_ld(12, _ld(1, n.encode())))
    graph = b"".join(chunks)
    model = _varint(1 << 3) + _varint(8) + _ld(2, b"pytorch") + _ld(3, b"2.7.1") + _ld(7, graph) + _ld(8, _varint(2 << 3) + _varint(16))
    with open(path, "wb") as fh:
        fh.write(model)


def test_round_trip_through_the_weights_container(tmp_path):
    donor = his.create_rgb_hierarchical_model(**KW)
    sd = common.procedural_state({k: list(v.shape) for k, v in donor.state_dict().items()})
    path = str(tmp_path / "weights.onnx")
    onnx_io.save_initializers(path, sd)
    info = onnx_io.read_model(path)
    assert info["inputs"] == ["images", "rois"] and info["outputs"] == ["instance_masks", "binary_masks"] and info["opset"] == 16
    m = his.create_rgb_hierarchical_model(**KW)
    rep = onnx_io.load_onnx_initializers(m, path)
    assert rep["missing"] == [] and rep["unused"] == [] and rep["anonymous"] == []
    got = m.state_dict()
    for k, v in sd.items():
        if not k.endswith("num_batches_tracked"):
            assert torch.equal(got[k], v), k


@pytest.mark.parametrize("packed_dims,typed_data,prefix", [(False, False, "model."), (True, True, "base_model.model."), (True, False, "")])
def test_reader_against_an_independent_encoder(tmp_path, packed_dims, typed_data, prefix):
    """Exporter-style files: wrapper prefixes (RGBHierarchicalWrapper.model, ModelWithDilation.base_model), dims packed or not,
    raw_data or typed payloads, a scalar, an int64 tensor, a fp16 tensor and an exporter-generated (BatchNorm-folded) initializer."""
    donor = his.create_rgb_hierarchical_model(**KW)
    sd = common.procedural_state({k: list(v.shape) for k, v in donor.state_dict().items()})
    tensors = {prefix + k: v.numpy() for k, v in sd.items() if not k.endswith("num_batches_tracked")}
    tensors["onnx::Conv_1234"] = np.arange(24, dtype=np.float32).reshape(2, 3, 2, 2)
    tensors["shape_const"] = np.array([-1, 3, 48, 64], dtype=np.int64)
    tensors["half_const"] = np.array([0.5, -2.0], dtype=np.float16)
    path = str(tmp_path / "exported.onnx")
    _independent_writer(path, tensors, ["images", "rois"], ["masks", "binary_masks"], packed_dims, typed_data)
    info = onnx_io.read_model(path)
    assert info["outputs"] == ["masks", "binary_masks"] and ("Conv", "/model/Conv") in info["nodes"] and info["producer"] == "pytorch"
    assert np.array_equal(info["initializers"]["shape_const"], tensors["shape_const"])
    assert np.array_equal(info["initializers"]["half_const"], tensors["half_const"])
    assert np.array_equal(info["initializers"]["onnx::Conv_1234"], tensors["onnx::Conv_1234"])
    m = his.create_rgb_hierarchical_model(**KW)
    rep = onnx_io.load_onnx_initializers(m, path)
    assert rep["missing"] == [] and rep["anonymous"] == ["onnx::Conv_1234"] and sorted(rep["unused"]) == ["half_const", "shape_const"]
    for k, v in m.state_dict().items():
        if not k.endswith("num_batches_tracked"):
            assert torch.equal(v, sd[k]), k
    # a file that lacks parameters (BatchNorm folded away by constant folding) is refused unless strict=False
    part = {k: v for k, v in tensors.items() if "running_var" not in k}
    _independent_writer(path, part, ["images", "rois"], ["instance_masks", "binary_masks"])
    with pytest.raises(onnx_io.OnnxFormatError, match="no initializer"):
        onnx_io.load_onnx_initializers(his.create_rgb_hierarchical_model(**KW), path)
    rep = onnx_io.load_onnx_initializers(his.create_rgb_hierarchical_model(**KW), path, strict=False)
    assert rep["missing"] and all("running_var" in k for k in rep["missing"])


def test_malformed_files_are_rejected(tmp_path):
    p = tmp_path / "bad.onnx"
    p.write_bytes(b"\x0a\x05hello")                     # a ModelProto without a graph
    with pytest.raises(onnx_io.OnnxFormatError, match="GraphProto"):
        onnx_io.read_model(str(p))
    p.write_bytes(_ld(7, _ld(5, _ld(8, b"w") + _varint(1 << 3) + _varint(4) + _ld(9, b"\x00" * 8))))     # 4 floats declared, 2 stored
    with pytest.raises(onnx_io.OnnxFormatError, match="elements"):
        onnx_io.read_model(str(p))
    p.write_bytes(_ld(7, b"\x2a\x7f\x00"))             # length runs past the end
    with pytest.raises(onnx_io.OnnxFormatError):
        onnx_io.read_model(str(p))


def test_metadata_json_matches_the_exporter(tmp_path):
    m = his.create_rgb_hierarchical_model(**KW)
    meta = onnx_io.export_metadata(m, checkpoint_path="experiments/x/checkpoints/best_model.pth", experiment_config="cfg_b0", dilation_pixels=1,
                                   image_size=(480, 640), batch_size=1, checkpoint={"epoch": 12, "best_miou": torch.tensor(0.8545)})
    out = onnx_io.write_metadata(str(tmp_path / "model_b0.onnx"), meta)
    assert out.endswith("model_b0.json")
    got = json.load(open(out))
    # key set and value formats of export_hierarchical_instance_peopleseg_onnx.py:511-535
    assert list(got) == ["checkpoint_path", "architecture", "roi_size", "mask_size", "experiment_config", "dilation_pixels", "image_size",
                         "input_format", "output_format", "epoch", "best_miou"]
    assert got["architecture"] == "B0" and got["roi_size"] == [16, 12] and got["mask_size"] == [32, 24] and got["image_size"] == [480, 640]
    assert got["input_format"] == {"images": "[1, 3, 480, 640] - RGB input images", "rois": "[N, 5] - ROIs in format [batch_idx, x1, y1, x2, y2]"}
    assert got["output_format"]["instance_masks"] == "[N, 1, 32, 24] - Binary class-1 mask per ROI (0.0 or 1.0)"
    assert got["output_format"]["binary_masks"] == "[B, 1, 480, 640] - Binary foreground/background masks from pretrained UNet"
    assert got["epoch"] == 12 and abs(got["best_miou"] - 0.8545) < 1e-6
