"""Oracle (test infrastructure): import the REAL reference modules (this container only).

``/root/reference`` is read-only and does not exist on the GPU box, so nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` goes through this file.  It is used by
``oracle/make_golden.py`` (to generate tests/golden/*) and by the CPU tests that pin
``oracle/headport.py`` / ``oracle/postport.py`` against the reference itself.

``segmentation_models_pytorch`` is not installed; the reference imports it lazily at
``hierarchical_segmentation_unet.py:1761-1767``.  We register a stub module whose
``Unet`` is the restatement in ``oracle/effunet.py`` -- everything else that runs is
reference code.
"""
from __future__ import annotations

import contextlib
import importlib
import io
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("HIS_REFERENCE_ROOT", "/root/reference")


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "src", "human_edge_detection"))


def _install_stub():
    if "segmentation_models_pytorch" not in sys.modules:
        from . import effunet
        stub = types.ModuleType("segmentation_models_pytorch")
        stub.Unet = effunet.Unet
        stub.__version__ = "0.5.0-oracle-restatement"
        sys.modules["segmentation_models_pytorch"] = stub


def ref_import(modname: str):
    """Imports ``src.human_edge_detection.<modname>`` from the reference tree."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the tree is read-only
    return importlib.import_module(f"src.human_edge_detection.{modname}")


def ref_root_import(modname: str):
    """Imports a top-level reference script module (e.g. export_edge_smoothing_onnx)."""
    if not available():
        raise RuntimeError(f"reference tree not found at {REFERENCE_ROOT}")
    _install_stub()
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    sys.dont_write_bytecode = True
    return importlib.import_module(modname)


def ref_class_from_script(script: str, class_name: str):
    """Executes ONE class definition of a top-level reference script (e.g. ``MaskDilationModule``
    in export_hierarchical_instance_peopleseg_onnx.py:85-141) without running the script's
    imports (``onnx``/``onnxsim``/``train_advanced`` are not importable here).  The class body
    that runs is the reference's own source, read from the read-only tree."""
    import ast
    import numpy as np
    import torch
    import torch.nn as nn
    import torch.nn.functional as F
    path = os.path.join(REFERENCE_ROOT, script)
    tree = ast.parse(open(path).read(), filename=path)
    for node in tree.body:
        if isinstance(node, ast.ClassDef) and node.name == class_name:
            mod = ast.Module(body=[node], type_ignores=[])
            ns = {"torch": torch, "nn": nn, "F": F, "np": np, "__name__": "ref_" + class_name}
            import typing
            ns.update({k: getattr(typing, k) for k in ("Optional", "Union", "Tuple", "List", "Dict")})
            exec(compile(mod, path, "exec"), ns)
            return ns[class_name]
    raise KeyError(f"{class_name} not found in {script}")


def ref_functions_from_script(script: str, names):
    """Like ``ref_class_from_script`` for top-level functions (e.g. ``denormalize_bbox`` / ``process_mask_output`` of
    test_hierarchical_instance_peopleseg_onnx.py, whose module-level imports need onnxruntime/pycocotools)."""
    import ast
    import typing
    import cv2
    import numpy as np
    import torch
    import torch.nn.functional as F
    path = os.path.join(REFERENCE_ROOT, script)
    tree = ast.parse(open(path).read(), filename=path)
    ns = {"torch": torch, "F": F, "np": np, "cv2": cv2, "__name__": "ref_funcs"}
    ns.update({k: getattr(typing, k) for k in ("Optional", "Union", "Tuple", "List", "Dict", "Any")})
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in names]
    exec(compile(ast.Module(body=body, type_ignores=[]), path, "exec"), ns)
    return {n: ns[n] for n in names}


def build_reference_model(**kwargs):
    """``create_rgb_hierarchical_model(**kwargs)`` of the reference, stdout silenced, eval mode."""
    mod = ref_import("advanced.hierarchical_segmentation_rgb")
    with contextlib.redirect_stdout(io.StringIO()):
        model = mod.create_rgb_hierarchical_model(**kwargs)
    return model.eval()
