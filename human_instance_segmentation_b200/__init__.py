"""B200-native (sm_100a) implementation of the RGB hierarchical instance-segmentation inference path of
PINTO0309/human-instance-segmentation, behind the reference's ``src/human_edge_detection`` model/ROI API.

    from human_instance_segmentation_b200 import create_rgb_hierarchical_model
    model = create_rgb_hierarchical_model(**same_kwargs_as_reference).to("cuda")
    logits, aux = model(images, rois)

Compute runs in ``libhis_b200.so`` (hand-written CUDA: tcgen05/TMEM implicit-GEMM convolutions fed by TMA, fused
depthwise/SE/attention/stencil kernels).  No CPU or torch-op fallback: calls raise ``HisError`` without the library/GPU.
"""
from .lib import HisError, build, load  # noqa: F401
from .model import (HierarchicalRGBSegmentationModelWithFullImagePretrainedUNet,  # noqa: F401
                    PreTrainedPeopleSegmentationUNet, PreTrainedPeopleSegmentationUNetWrapper, create_rgb_hierarchical_model)
from .roi_align import DynamicRoIAlign  # noqa: F401
from . import metrics, postprocess, preprocess  # noqa: F401

__all__ = ["create_rgb_hierarchical_model", "HierarchicalRGBSegmentationModelWithFullImagePretrainedUNet",
           "PreTrainedPeopleSegmentationUNetWrapper", "PreTrainedPeopleSegmentationUNet", "DynamicRoIAlign", "postprocess", "metrics", "preprocess",
           "HisError", "build", "load"]
