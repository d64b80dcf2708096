"""First-contact script for a GPU box: runs every conv-GEMM case and the model parity checks with verbose numbers.
Usage (on the box): timeout 600 python tools/gpu_bringup.py [gemm|model|all]"""
import sys
import time
import traceback

import torch

sys.path.insert(0, ".")


def gemm():
    from tests.test_gpu_conv_gemm import CASES, run_case
    for c in CASES:
        t = time.time()
        try:
            err, ref = run_case(**c)
            print(f"GEMM {c}: err={err:.3e} ref={ref:.3e} {'OK' if err <= 2e-3 * max(ref, 1) else 'FAIL'} ({time.time() - t:.2f}s)", flush=True)
        except Exception:
            print(f"GEMM {c}: EXCEPTION", flush=True)
            traceback.print_exc()


def model():
    import human_instance_segmentation_b200 as his
    from tests import common
    for name in ["small_b0_bn_relu", "small_b0_bn_relu_exportscale", "small_b1_bc72", "small_b7_bc96_d4", "small_b0_bn_swish_beta",
                 "small_b0_resize", "cfg1_b0"]:
        try:
            if name == "cfg1_b0":
                cfg, images, rois = common.cfg1_inputs()
                shapes = common.golden_keys()["preset_b0"]
            else:
                cfg, images, rois = common.small_case_inputs(name)
                shapes = common.shapes_for_case(name)
            g = common.golden(name)
            m = his.create_rgb_hierarchical_model(**cfg.factory_kwargs())
            m.load_state_dict(common.procedural_state(shapes, weights_path=cfg.pretrained_weights_path))
            m = m.to("cuda")
            for ra in (m.roi_align_mask, m.roi_align_rgb):
                ra.spatial_scale = cfg.spatial_scale
                ra.spatial_scale_h, ra.spatial_scale_w = cfg.spatial_scale
            t = time.time()
            logits, aux = m(images.cuda(), rois.cuda())
            torch.cuda.synchronize()
            dt = time.time() - t
            logits = logits.cpu()
            print(f"MODEL {name}: logits rel_err={common.rel_err(logits, g['logits']):.3e} argmax={common.argmax_agreement(logits, g['logits']):.5f} "
                  f"({dt:.2f}s first call)", flush=True)
            for k, v in aux.items():
                v = v.cpu()
                if k in g:
                    print(f"    {k}: rel_err={common.rel_err(v, g[k]):.3e}")
                elif k == "full_image_logits" and "full_image_logits_ch0" in g:
                    print(f"    {k}: rel_err={common.rel_err(v[:, 0], g['full_image_logits_ch0']):.3e}")
                elif k == "full_image_logits" and "full_image_logits_ch0_s2" in g:
                    print(f"    {k}: rel_err={common.rel_err(v[:, 0, ::2, ::2], g['full_image_logits_ch0_s2']):.3e}")
                elif k + "_sub" in g:
                    sub = v[:, ::8] if name != "cfg1_b0" else v[:, ::8, ::4, ::4]
                    print(f"    {k}: rel_err={common.rel_err(sub, g[k + '_sub']):.3e}")
        except Exception:
            print(f"MODEL {name}: EXCEPTION", flush=True)
            traceback.print_exc()


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    print(torch.cuda.get_device_name(0), flush=True)
    if what in ("gemm", "all"):
        gemm()
    if what in ("model", "all"):
        model()
