"""The three production presets as ``create_rgb_hierarchical_model`` kwargs, restated from the reference's
``hed/experiments/config_manager.py`` (B0 std :2558-2589 with the ModelConfig defaults :189-190, B1 enhanced
:3643-3676, B7 ultra :3851-3884) the way ``train_advanced.build_model`` passes them (train_advanced.py:132-160)."""


def _kw(roi, mask, enc, path, base, depth):
    return dict(roi_size=roi, mask_size=mask, multi_scale=False, use_attention_module=True, use_boundary_refinement=False,
                use_progressive_upsampling=False, use_subpixel_conv=False, use_contour_detection=True, use_distance_transform=True,
                normalization_type="batchnorm", normalization_groups=8, activation_function="relu", activation_beta=1.0,
                use_pretrained_unet=True, pretrained_weights_path=path, freeze_pretrained_weights=True, use_full_image_unet=True,
                encoder_name=enc, hierarchical_base_channels=base, hierarchical_depth=depth)


PRESETS = {
    "b0": _kw((64, 48), (128, 96), "timm-efficientnet-b0", "ext_extractor/best_model_b0_0.8741.pth", 64, 3),
    "b1_enhanced": _kw((80, 60), (160, 120), "timm-efficientnet-b1", "ext_extractor/best_model_b1_0.8833.pth", 72, 3),
    "b7_ultra": _kw((128, 96), (256, 192), "timm-efficientnet-b7", "ext_extractor/best_model_b7_0.9009.pth", 96, 4),
}
