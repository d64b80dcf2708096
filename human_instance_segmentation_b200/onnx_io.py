"""ONNX-side tooling of the exported contract without the ``onnx`` package (it is not installable offline).

The reference ships its models as ``.onnx`` files written by ``export_hierarchical_instance_peopleseg_onnx.py:331-551`` through
``AdvancedONNXExporter._export_hierarchical`` (hed/export_onnx_advanced.py:306-473): inputs ``images`` / ``rois``, outputs
``instance_masks`` / ``binary_masks`` (older releases: ``masks``), parameters stored as graph initializers named after the
wrapped module's state-dict keys (``model.<key>``, or ``base_model.<key>`` under ``ModelWithDilation``), and a metadata JSON
next to the file (:511-528).  This module

  * reads the initializers of such a file with a minimal protobuf wire-format parser (``read_model``),
  * maps them onto a model's state-dict keys and loads them (``load_onnx_initializers``) -- checkpoints are then not needed,
  * writes the same metadata JSON (``export_metadata`` / ``write_metadata``), and
  * writes a weights-only ONNX container with the exported I/O names (``save_initializers``), the inverse of the reader.

Wire format (protobuf encoding; onnx.proto field numbers):
  ModelProto   { 1 ir_version, 2 producer_name, 7 graph, 8 opset_import }
  GraphProto   { 1 node, 2 name, 5 initializer, 11 input, 12 output }
  TensorProto  { 1 dims, 2 data_type, 4 float_data, 5 int32_data, 7 int64_data, 8 name, 9 raw_data, 10 double_data }
  ValueInfoProto { 1 name },  NodeProto { 1 input, 2 output, 3 name, 4 op_type }
"""
from __future__ import annotations

import json
import os
import struct
from typing import Dict, Iterator, List, Optional, Sequence, Tuple

import numpy as np
import torch

# TensorProto.DataType -> numpy dtype
_DTYPES = {1: np.float32, 2: np.uint8, 3: np.int8, 5: np.int16, 6: np.int32, 7: np.int64, 9: np.bool_, 10: np.float16, 11: np.float64,
           12: np.uint32, 13: np.uint64}
_ONNX_OF = {np.dtype(v): k for k, v in _DTYPES.items()}


class OnnxFormatError(ValueError):
    pass


# ----------------------------------------------------------------------------- wire-format reader
def _varint(buf: memoryview, pos: int) -> Tuple[int, int]:
    result, shift = 0, 0
    while True:
        if pos >= len(buf):
            raise OnnxFormatError("truncated varint")
        b = buf[pos]
        pos += 1
        result |= (b & 0x7F) << shift
        if not b & 0x80:
            return result, pos
        shift += 7
        if shift > 70:
            raise OnnxFormatError("varint too long")


def _fields(buf: memoryview) -> Iterator[Tuple[int, int, object]]:
    """Yields (field number, wire type, value) of one message: varint -> int, 64-bit / 32-bit -> bytes, length-delimited -> memoryview."""
    pos, n = 0, len(buf)
    while pos < n:
        key, pos = _varint(buf, pos)
        field, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _varint(buf, pos)
        elif wt == 1:
            v, pos = bytes(buf[pos:pos + 8]), pos + 8
        elif wt == 2:
            ln, pos = _varint(buf, pos)
            if pos + ln > n:
                raise OnnxFormatError("length-delimited field runs past the end of its message")
            v, pos = buf[pos:pos + ln], pos + ln
        elif wt == 5:
            v, pos = bytes(buf[pos:pos + 4]), pos + 4
        else:
            raise OnnxFormatError(f"unsupported wire type {wt}")
        yield field, wt, v


def _signed64(v: int) -> int:
    return v - (1 << 64) if v >= (1 << 63) else v


def _packed_varints(v, wt) -> List[int]:
    if wt == 0:
        return [_signed64(v)]
    out, pos = [], 0
    while pos < len(v):
        x, pos = _varint(v, pos)
        out.append(_signed64(x))
    return out


def _tensor(buf: memoryview) -> Tuple[str, np.ndarray]:
    dims: List[int] = []
    dtype, name, raw = 1, "", None
    floats: List[np.ndarray] = []
    ints: List[int] = []
    doubles: List[np.ndarray] = []
    for f, wt, v in _fields(buf):
        if f == 1:
            dims += _packed_varints(v, wt)
        elif f == 2:
            dtype = int(v)
        elif f == 4:
            floats.append(np.frombuffer(v, dtype="<f4") if wt == 2 else np.frombuffer(v, dtype="<f4"))
        elif f in (5, 7):
            ints += _packed_varints(v, wt)
        elif f == 8:
            name = bytes(v).decode("utf-8")
        elif f == 9:
            raw = bytes(v)
        elif f == 10:
            doubles.append(np.frombuffer(v, dtype="<f8"))
        elif f in (13, 14) and (f == 13 or int(v) == 1):
            raise OnnxFormatError(f"initializer {name!r} keeps its data in an external file; only in-file tensors are supported")
    if dtype not in _DTYPES:
        raise OnnxFormatError(f"initializer {name!r}: unsupported data_type {dtype}")
    np_dt = np.dtype(_DTYPES[dtype])
    count = int(np.prod(dims)) if dims else 1
    if raw is not None:
        arr = np.frombuffer(raw, dtype=np_dt.newbyteorder("<"))
    elif floats:
        arr = np.concatenate(floats).astype(np_dt)
    elif doubles:
        arr = np.concatenate(doubles).astype(np_dt)
    elif ints:
        if dtype == 10:      # FLOAT16 in int32_data: the bit patterns
            arr = np.array(ints, dtype=np.uint16).view(np.float16)
        else:
            arr = np.array(ints).astype(np_dt)
    else:
        arr = np.zeros(0, dtype=np_dt)
    if arr.size != count:
        raise OnnxFormatError(f"initializer {name!r}: {arr.size} elements for dims {dims}")
    return name, arr.reshape(dims).copy()


def _value_name(buf: memoryview) -> str:
    for f, wt, v in _fields(buf):
        if f == 1 and wt == 2:
            return bytes(v).decode("utf-8")
    return ""


def read_model(path: str) -> Dict[str, object]:
    """Parses an ``.onnx`` file: {'initializers': {name: ndarray}, 'inputs': [...], 'outputs': [...], 'nodes': [(op_type, name)],
    'producer': str, 'opset': int | None}.  Graph inputs that are initializers (old exporters list them) are dropped from 'inputs'."""
    with open(path, "rb") as fh:
        data = memoryview(fh.read())
    graph, producer, opset = None, "", None
    for f, wt, v in _fields(data):
        if f == 7 and wt == 2:
            graph = v
        elif f == 2 and wt == 2:
            producer = bytes(v).decode("utf-8", "replace")
        elif f == 8 and wt == 2:
            for f2, wt2, v2 in _fields(v):
                if f2 == 2 and wt2 == 0:
                    opset = int(v2)
    if graph is None:
        raise OnnxFormatError(f"{path}: no GraphProto (field 7) in the ModelProto -- not an ONNX model file")
    inits: Dict[str, np.ndarray] = {}
    inputs, outputs, nodes = [], [], []
    for f, wt, v in _fields(graph):
        if wt != 2:
            continue
        if f == 5:
            name, arr = _tensor(v)
            inits[name] = arr
        elif f == 11:
            inputs.append(_value_name(v))
        elif f == 12:
            outputs.append(_value_name(v))
        elif f == 1:
            op, nm = "", ""
            for f2, wt2, v2 in _fields(v):
                if f2 == 4 and wt2 == 2:
                    op = bytes(v2).decode("utf-8")
                elif f2 == 3 and wt2 == 2:
                    nm = bytes(v2).decode("utf-8")
            nodes.append((op, nm))
    return {"initializers": inits, "inputs": [n for n in inputs if n not in inits], "outputs": outputs, "nodes": nodes,
            "producer": producer, "opset": opset}


def read_initializers(path: str) -> Dict[str, torch.Tensor]:
    return {k: torch.from_numpy(v) for k, v in read_model(path)["initializers"].items()}


# ----------------------------------------------------------------------------- initializers -> state dict
def map_initializers(initializers: Dict[str, torch.Tensor], state_dict: Dict[str, torch.Tensor]):
    """Maps exported initializer names onto state-dict keys.  The exporter wraps the model (``RGBHierarchicalWrapper.model``,
    optionally ``ModelWithDilation.base_model``; export_onnx_advanced.py:353-420, export_hier...py:144-181), so an initializer is
    named ``<wrapper prefixes>.<state-dict key>``: a key matches the initializer whose name equals it or ends with ``.<key>`` and
    whose shape agrees (a scalar initializer also matches a 0-d parameter).  Returns (matched state dict, report) with
    report = {'missing': keys without initializer, 'unused': initializers without key, 'anonymous': exporter-generated names such
    as ``onnx::Conv_123`` -- BatchNorm folded into the convolution by constant folding; those cannot be un-folded by name}."""
    by_suffix: Dict[str, List[str]] = {}
    for name in initializers:
        parts = name.split(".")
        for i in range(len(parts)):
            by_suffix.setdefault(".".join(parts[i:]), []).append(name)
    out, used = {}, set()
    for key, ref in state_dict.items():
        cands = [n for n in by_suffix.get(key, []) if tuple(initializers[n].shape) == tuple(ref.shape)]
        if not cands:
            continue
        name = min(cands, key=len)         # the shortest wrapper prefix
        out[key] = initializers[name].to(ref.dtype) if initializers[name].dtype != ref.dtype else initializers[name]
        used.add(name)
    unused = [n for n in initializers if n not in used]
    report = {"missing": [k for k in state_dict if k not in out],
              "unused": [n for n in unused if "::" not in n], "anonymous": [n for n in unused if "::" in n]}
    return out, report


def load_onnx_initializers(model: torch.nn.Module, path: str, strict: bool = True) -> Dict[str, List[str]]:
    """Loads the parameters stored in a released ``.onnx`` file into ``model`` (the reference's checkpoint path is
    ``load_state_dict(ckpt['model_state_dict'])``, export_hier...py:423-441).  BatchNorm's ``num_batches_tracked`` is never exported
    and is not required.  strict: raise when a parameter / buffer of the model has no initializer."""
    info = read_model(path)
    sd = model.state_dict()
    matched, report = map_initializers({k: torch.from_numpy(v) for k, v in info["initializers"].items()}, sd)
    report["missing"] = [k for k in report["missing"] if not k.endswith("num_batches_tracked")]
    report["inputs"], report["outputs"] = info["inputs"], info["outputs"]
    if strict and report["missing"]:
        hint = " (the file holds BatchNorm-folded convolutions: export with do_constant_folding=False or load the checkpoint)" if report["anonymous"] else ""
        raise OnnxFormatError(f"{path}: {len(report['missing'])} model parameters have no initializer, e.g. {report['missing'][:5]}{hint}")
    full = dict(sd)
    full.update(matched)
    model.load_state_dict(full, strict=True)
    return report


# ----------------------------------------------------------------------------- wire-format writer (weights-only container)
def _enc_varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        out.append(b | (0x80 if v else 0))
        if not v:
            return bytes(out)


def _ld(field: int, payload: bytes) -> bytes:
    return _enc_varint((field << 3) | 2) + _enc_varint(len(payload)) + payload


def _vi(field: int, v: int) -> bytes:
    return _enc_varint(field << 3) + _enc_varint(v)


def _tensor_proto(name: str, arr: np.ndarray) -> bytes:
    arr = np.asarray(arr)                  # (ascontiguousarray would turn a 0-d tensor into shape (1,))
    if arr.dtype not in _ONNX_OF:
        raise OnnxFormatError(f"{name}: dtype {arr.dtype} has no ONNX tensor type")
    msg = b"".join(_vi(1, int(d)) for d in arr.shape) + _vi(2, _ONNX_OF[arr.dtype]) + _ld(8, name.encode()) + \
        _ld(9, arr.astype(arr.dtype.newbyteorder("<")).tobytes())
    return msg


def save_initializers(path: str, state_dict: Dict[str, torch.Tensor], prefix: str = "model.", inputs: Sequence[str] = ("images", "rois"),
                      outputs: Sequence[str] = ("instance_masks", "binary_masks"), opset: int = 16) -> None:
    """Writes a weights-only ONNX container: a ModelProto whose graph has no nodes, the exported contract's input / output names
    and one initializer per state-dict entry named ``prefix + key`` (``num_batches_tracked`` is skipped like the exporter does)."""
    parts = [_ld(2, b"his_b200_weights")]
    for k, v in state_dict.items():
        if k.endswith("num_batches_tracked"):
            continue
        parts.append(_ld(5, _tensor_proto(prefix + k, v.detach().cpu().numpy())))
    parts += [_ld(11, _ld(1, n.encode())) for n in inputs]
    parts += [_ld(12, _ld(1, n.encode())) for n in outputs]
    graph = b"".join(parts)
    model = _vi(1, 8) + _ld(2, b"human_instance_segmentation_b200") + _ld(7, graph) + _ld(8, _vi(2, opset))
    with open(path, "wb") as fh:
        fh.write(model)


# ----------------------------------------------------------------------------- metadata JSON (export_hier...py:511-528)
def export_metadata(model, checkpoint_path: str = "", experiment_config: str = "", dilation_pixels: int = 0, image_size: Tuple[int, int] = (480, 640),
                    batch_size: int = 1, checkpoint: Optional[dict] = None) -> dict:
    """The metadata dictionary the reference writes next to an exported model."""
    roi_size, mask_size = tuple(model.roi_size), tuple(model.mask_size)
    h, w = image_size
    enc = getattr(getattr(model, "pretrained_unet", None), "model", None)
    arch = getattr(enc, "encoder_name", "").replace("timm-efficientnet-", "").upper() or "unknown"
    meta = {
        "checkpoint_path": str(checkpoint_path),
        "architecture": arch,
        "roi_size": list(roi_size),
        "mask_size": list(mask_size),
        "experiment_config": experiment_config,
        "dilation_pixels": dilation_pixels,
        "image_size": [h, w],
        "input_format": {"images": f"[{batch_size}, 3, {h}, {w}] - RGB input images", "rois": "[N, 5] - ROIs in format [batch_idx, x1, y1, x2, y2]"},
        "output_format": {"instance_masks": f"[N, 1, {mask_size[0]}, {mask_size[1]}] - Binary class-1 mask per ROI (0.0 or 1.0)",
                          "binary_masks": f"[B, 1, {h}, {w}] - Binary foreground/background masks from pretrained UNet"},
    }
    if isinstance(checkpoint, dict):
        if "epoch" in checkpoint:
            meta["epoch"] = checkpoint["epoch"]
        if "best_miou" in checkpoint:
            meta["best_miou"] = float(checkpoint["best_miou"])
    return meta


def write_metadata(path: str, meta: dict) -> str:
    """Writes ``meta`` as ``<path without suffix>.json`` (the reference: ``output_path.with_suffix('.json')``)."""
    out = os.path.splitext(path)[0] + ".json"
    with open(out, "w") as fh:
        json.dump(meta, fh, indent=2)
    return out
