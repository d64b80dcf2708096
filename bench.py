#!/usr/bin/env python
"""Benchmark of the hot path: RGB hierarchical instance-segmentation forward, ROI-masks/s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload b0|b1|b7|b0_160x120] [--impl reference]

A step = one pass of ``model(images, rois)`` (full-image EfficientNet-UNet, Dynamic RoI Align, ROI feature extractor,
hierarchical head with contour/distance branches -> logits [N,3,mh,mw] + binary masks) over one synthetic batch.
  * value : ROI-masks/s with the batch already resident in HBM (CUDA events, max over ranks, barrier+sync both sides)
  * e2e   : the same through the public API with HOST (pinned) inputs: H2D of images+rois and D2H of
            instance_masks+binary_masks inside the timed region
  * roofline : tcgen05 implicit-GEMM kernel, algorithmic FLOPs / event-timed launch durations vs the measured peak
  * cpu_baseline : the reference path on the box's host cores (oracle port: the reference tree cannot travel)
N > 1 (torchrun): every rank runs the same per-GPU batch on its own images (weak scaling, no collective on the path).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (preset, images per GPU, H, W, rois per image)
    "b0": ("b0", 64, 480, 640, 10),            # the metric's model (B0, 640x480) on configs[1]'s 1xB200 batch geometry
    "b1": ("b1_enhanced", 64, 480, 640, 10),   # BASELINE.json configs[1]
    "b7": ("b7_ultra", 4, 480, 640, 10),       # configs[2] per-GPU share at 8 GPUs (32 images / 8)
    "b0_160x120": ("b0", 512, 120, 160, 10),   # configs[3] per-GPU share at 8 GPUs (4096 / 8)
}
HEAD_GFLOP_PER_ROI = {"b0": 57.84, "b1_enhanced": 93.01, "b7_ultra": 278.44}      # SURVEY §8d / BASELINE.md §2
UNET_GFLOP_PER_IMG = {"b0": 27.72, "b1_enhanced": 29.97, "b7_ultra": 90.21}


def measured_traffic(kernel: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r1_traffic.json: dram__bytes_read.sum
    + dram__bytes_write.sum summed over the kernel's launches of one step / number of launches).  None if absent."""
    path = os.path.join(ROOT, "profiles", "r1_traffic.json")
    try:
        return json.load(open(path)).get(kernel)
    except Exception:
        return None


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        p = json.load(open(path))
        return {"tflops": p.get("bf16_tflops_sustained", 1367.9), "tflops_burst": p.get("bf16_tflops", 1660.9),
                "hbm_gbs": p.get("hbm_gbs", 6555.2), "source": "measured (MEASURED_PEAKS.json)"}
    return {"tflops": 1400.0, "tflops_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index: int):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1])); power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(power) if power else None, "samples": len(sm), "reasons": sorted(reasons)}


def synth_batch(seed, n_images, h, w, per_image):
    from human_instance_segmentation_b200.synthetic import synth_images, synth_rois
    return synth_images(seed, n_images, h, w), synth_rois(seed, n_images, per_image)


def build_model(preset):
    """The product arm: nothing under oracle/ is imported here."""
    import human_instance_segmentation_b200 as his
    from human_instance_segmentation_b200 import presets, synthetic
    kw = presets.PRESETS[preset]
    model = his.create_rgb_hierarchical_model(**kw)
    # random-init weights of that architecture (no checkpoints exist offline); BatchNorm statistics randomised so that
    # nothing folds to a no-op
    model.load_state_dict(synthetic.fill_state_dict(model.state_dict(), seed=0))
    return kw, model


def cpu_reference_rate(preset, h, w, n_images, per_image, iters, threads):
    """ROI-masks/s of the reference path on host cores (oracle port of the reference modules, fp32, all threads)."""
    import torch
    from oracle import headport, paramfill
    import human_instance_segmentation_b200 as his
    torch.set_num_threads(threads)
    cfg = headport.PRESETS[preset]
    holder = his.create_rgb_hierarchical_model(**cfg.factory_kwargs())           # parameter container only (key names/shapes)
    sd = paramfill.fill_state_dict(holder.state_dict(), seed=0)
    images, rois = synth_batch(1, n_images, h, w, per_image)
    headport.forward(sd, images, rois, cfg)                                      # warm-up
    times = []
    for _ in range(iters):
        t = time.perf_counter()
        headport.forward(sd, images, rois, cfg)
        times.append(time.perf_counter() - t)
    return rois.shape[0] / min(times), min(times)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    preset, n_img, h, w, per_image = WORKLOADS[args.workload]
    threads = os.cpu_count() or 1
    sample_images = 2
    import torch
    from oracle import headport, paramfill
    import human_instance_segmentation_b200 as his
    torch.set_num_threads(threads)
    cfg = headport.PRESETS[preset]
    holder = his.create_rgb_hierarchical_model(**cfg.factory_kwargs())
    sd = paramfill.fill_state_dict(holder.state_dict(), seed=0)
    images, rois = synth_batch(1, sample_images, h, w, per_image)
    for _ in range(max(1, min(args.warmup, 2))):
        headport.forward(sd, images, rois, cfg)
    steps = max(1, min(args.steps, 8))
    t0 = time.perf_counter()
    for _ in range(steps):
        headport.forward(sd, images, rois, cfg)
    dt = (time.perf_counter() - t0) / steps
    value = rois.shape[0] / dt
    sample = f"{sample_images} images x {per_image} ROIs of the {args.workload} workload per step, fp32, torch CPU ({threads} threads)"
    print(json.dumps({
        "impl": "reference", "metric": "roi_masks_per_sec", "value": value, "unit": "ROI-masks/s", "n_gpus": args.gpus, "steps": steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.workload, sample_images),
        "cpu_baseline": {"value": value, "unit": "ROI-masks/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "ROI-masks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "reference path = oracle port of the reference modules (the reference tree and smp/timm cannot travel to the GPU box)"}))


def run_post(args):
    """BASELINE configs[4]: post-processing only.  A step = (a) the ROI chain on 10 ROIs per image -- MaskDilationModule on the
    [N,3,128,96] logits, argmax -> instance mask (u8), NEAREST paste-back onto the 480x640 label canvases -- and (b) the
    full-image mask clean-up -- BinaryMaskEdgeSmoothing + BinaryMaskBilateralFilter fused in one shared-memory pass over
    [B,1,480,640] masks.  value = masks (ROI masks + full-image masks) per second with inputs resident in HBM; the roofline is
    the fused stencil kernel against the measured HBM copy bandwidth (algorithmic bytes: read + write once, 2*H*W*4 per mask)."""
    import torch
    import torch.distributed as dist
    from human_instance_segmentation_b200 import postprocess as pp
    from human_instance_segmentation_b200.synthetic import synth_rois
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, H, W, per_image, mh, mw = 512, 480, 640, 10, 128, 96
    g = torch.Generator().manual_seed(5 + rank)
    full_h = (torch.nn.functional.avg_pool2d((torch.rand(B, 1, H // 4, W // 4, generator=g) > 0.5).float(), 5, 1, 2) > 0.5).float()
    full_h = torch.nn.functional.interpolate(full_h, size=(H, W), mode="nearest").contiguous().pin_memory()     # blob-like masks
    rois_h = synth_rois(5 + rank, B, per_image).pin_memory()
    N = rois_h.shape[0]
    logits_h = (torch.randn(N, 3, mh // 8, mw // 8, generator=g) * 2)
    logits_h = torch.nn.functional.interpolate(logits_h, size=(mh, mw), mode="bilinear").contiguous().pin_memory()
    full, rois, logits = full_h.to(dev), rois_h.to(dev), logits_h.to(dev)
    cleanup, dil = pp.MaskCleanup().to(dev), pp.MaskDilationModule(1)
    out_full = torch.empty_like(full)

    def step(f, r, lg):
        masks = pp.instance_masks(dil(lg), as_uint8=True)
        canvas = pp.paste_masks(masks[:65535], r[:65535], B, H, W)
        return cleanup(f, out=out_full), canvas

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        step(full, rois, logits)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step(full, rois, logits)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    # fused stencil alone (the dominant kernel), event-timed
    e0.record()
    for _ in range(args.steps):
        cleanup(full, out=out_full)
    e1.record()
    barrier()
    ms_fused = e0.elapsed_time(e1) / args.steps
    # e2e: pinned host in, cleaned masks + canvases back to pinned host
    res_h = torch.empty_like(full_h).pin_memory()
    canvas_h = torch.empty((B, H, W), dtype=torch.int32).pin_memory()

    def e2e_step():
        o, c = step(full_h.to(dev, non_blocking=True), rois_h.to(dev, non_blocking=True), logits_h.to(dev, non_blocking=True))
        res_h.copy_(o, non_blocking=True); canvas_h.copy_(c, non_blocking=True)

    e2e_step(); barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    t_all = torch.tensor([ms, ms_e2e, ms_fused], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
        dist.destroy_process_group()
    if rank != 0:
        return
    ms, ms_e2e, ms_fused = [float(v) for v in t_all]
    pk = peaks()
    units = B + N
    alg_bytes = 2.0 * B * H * W * 4
    achieved = alg_bytes / (ms_fused * 1e-3) / 1e9
    out = {"metric": "roi_masks_per_sec", "value": world * units / (ms * 1e-3), "unit": "masks/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (masks), u8/int32 (paste-back)",
           "data": "synthetic",
           "config": {"workload": f"post-processing only (BASELINE configs[4]): {B} full-image 640x480 masks (edge smoothing + binary bilateral, fused) + "
                                  f"{N} ROI logits 128x96 (dilation, argmax, NEAREST paste-back) per GPU per step", "masks_per_gpu": units,
                      "cache": f"inputs larger than L2 ({int(alg_bytes / 2e6)} MB of masks, {int(N * 3 * mh * mw * 4 / 1e6)} MB of logits per step)"},
           "clocks": clocks,
           "e2e": {"value": world * units / (ms_e2e * 1e-3), "unit": "masks/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": full_h.numel() * 4 + logits_h.numel() * 4 + rois_h.numel() * 4,
                   "d2h_bytes_per_step": res_h.numel() * 4 + canvas_h.numel() * 4,
                   "api": "postprocess.MaskDilationModule / instance_masks / paste_masks / MaskCleanup on pinned host tensors"},
           "gpu_launches": 5 * args.steps, "launches_per_step": 5,
           "roofline": {"bound": "hbm", "kernel": "mask_cleanup_fused_kernel", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": achieved / pk["hbm_gbs"], "traffic": measured_traffic("mask_cleanup_fused_kernel:post"), "peak_source": pk["source"],
                        "avg_launch_ms": ms_fused,
                        "algorithmic_bytes_per_launch": alg_bytes, "share_of_step": ms_fused / ms,
                        "how": "2*H*W*4 bytes per mask (read once + write once) x masks per launch / CUDA-event time of the launch"}}
    if not args.no_cpu_baseline and world == 1:
        from oracle import postport
        threads = os.cpu_count() or 1
        torch.set_num_threads(threads)
        n_s = 8
        t0 = time.perf_counter()
        postport.binary_bilateral(postport.edge_smooth(full_h[:n_s]))
        lg = postport.instance_mask(__import__("oracle.headport", fromlist=["x"]).mask_dilation(logits_h[: n_s * per_image], 1))
        postport.paste_back(lg[:, 0].numpy().astype("uint8"), rois_h[: n_s * per_image].numpy(), B, H, W)
        dt = time.perf_counter() - t0
        out["cpu_baseline"] = {"value": (n_s + n_s * per_image) / dt, "unit": "masks/s", "cores": threads, "kind": "port",
                               "sample": f"{n_s} full-image masks + {n_s * per_image} ROI masks of the same workload, one pass ({dt:.2f} s), torch CPU"}
    print(json.dumps(out))


def workload_config(name, n_img=None):
    preset, b, h, w, per_image = WORKLOADS[name]
    b = n_img or b
    from human_instance_segmentation_b200 import presets
    kw = presets.PRESETS[preset]
    rs, msz = kw["roi_size"], kw["mask_size"]
    return {"workload": f"{preset} hierarchical RGB model, {w}x{h} input, batch {b} per GPU, {per_image} ROIs/image, ROI {rs[0]}x{rs[1]} -> mask "
                        f"{msz[0]}x{msz[1]}, random-init weights", "images_per_gpu": b, "rois_per_gpu": b * per_image,
            "encoder": kw["encoder_name"], "cache": "inputs larger than L2 (236 MB images, multi-GB activations per step)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="b0", choices=list(WORKLOADS) + ["post"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--breakdown", action="store_true", help="print the per-op time table of one instrumented step to stderr")
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--no-pipeline", action="store_true", help="e2e through the blocking model.infer instead of infer_pipelined")
    ap.add_argument("--no-graph", action="store_true", help="replay the launch plan kernel by kernel instead of as one CUDA graph")
    args = ap.parse_args()
    if args.workload == "post":
        return run_post(args)
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    warmup = max(args.warmup, 3)
    preset, n_img, h, w, per_image = WORKLOADS[args.workload]
    cfg, model = build_model(preset)
    model = model.to(dev)
    for ra in (model.roi_align_mask, model.roi_align_rgb):       # exporter convention spatial_scale=(H,W)
        ra.spatial_scale = (float(h), float(w)); ra.spatial_scale_h, ra.spatial_scale_w = float(h), float(w)
    model.copy_outputs = False
    model.use_cuda_graph = not args.no_graph          # launch-bound tail of ~200 small kernels -> one graph submission per step
    images_h, rois_h = synth_batch(100 + rank, n_img, h, w, per_image)
    images_h, rois_h = images_h.pin_memory(), rois_h.pin_memory()
    images_d, rois_d = images_h.to(dev), rois_h.to(dev)
    n_rois = rois_h.shape[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # ---------------- value: inputs resident in HBM
    for _ in range(warmup):
        model(images_d, rois_d)
    bp = model._get_plan(images_d, rois_d)
    plan = bp.plan
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        model(images_d, rois_d)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1) / args.steps
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- per-kernel timing for the roofline (instrumented replay right after the timed region)
    timed = plan.run_timed()
    timed = plan.run_timed()
    n_launches, step_flops, op_desc = plan.launches, plan.flops, list(plan.op_desc)
    del plan, bp
    model.release_plans()          # the e2e arm builds its own (masks-only) plans; keep the HBM footprint to two of them

    # ---------------- e2e: host (pinned) buffers in, instance/binary masks back to host, through model.infer()
    inst_h = torch.empty((n_rois, 1) + tuple(cfg["mask_size"]), dtype=torch.float32).pin_memory()
    bin_h = torch.empty((n_img, 1, h, w), dtype=torch.float32).pin_memory()

    # two sets of host output buffers: the pipelined API downloads batch i-1 while batch i computes
    inst_h2, bin_h2 = torch.empty_like(inst_h).pin_memory(), torch.empty_like(bin_h).pin_memory()
    host_out = [(inst_h, bin_h), (inst_h2, bin_h2)]
    e2e_i = [0]

    def e2e_step():
        if args.no_pipeline:
            inst, binary = model.infer(images_h, rois_h)
            inst_h.copy_(inst, non_blocking=True)
            bin_h.copy_(binary, non_blocking=True)
        else:
            o = host_out[e2e_i[0] & 1]; e2e_i[0] += 1
            model.infer_pipelined(images_h, rois_h, o[0], o[1])

    for _ in range(2):
        e2e_step()
    model.pipeline_sync()
    barrier()
    e0.record()
    for _ in range(args.steps):
        e2e_step()
    model.pipeline_sync()          # every batch of the timed region has landed in host memory (current stream waits for the downloads)
    e1.record()
    barrier()
    ms_e2e = e0.elapsed_time(e1) / args.steps
    # the same through the blocking call (upload, forward, download strictly one after the other), for comparison
    ms_e2e_blocking = None
    if not args.no_pipeline:
        args.no_pipeline = True
        e2e_step(); barrier()
        e0.record()
        for _ in range(args.steps):
            e2e_step()
        e1.record()
        barrier()
        ms_e2e_blocking = e0.elapsed_time(e1) / args.steps
        args.no_pipeline = False

    gemm = [(t, f) for name, t, f in timed if name == "conv_gemm"]
    gemm_ms, gemm_flops = sum(t for t, _ in gemm), sum(f for _, f in gemm)
    step_ms_instr = sum(t for _, t, _ in timed)
    by_name = {}
    for name, t, f in timed:
        by_name.setdefault(name, [0, 0.0]); by_name[name][0] += 1; by_name[name][1] += t
    if args.breakdown and rank == 0:
        for name, (cnt, t) in sorted(by_name.items(), key=lambda kv: -kv[1][1]):
            print(f"  {name:22s} x{cnt:4d} {t:9.3f} ms {100 * t / step_ms_instr:5.1f}%", file=sys.stderr)
        rows = sorted(((t, op_desc[i], f) for i, (n_, t, f) in enumerate(timed)), reverse=True)
        for t, desc, f in rows[:args.top]:
            print(f"  {t:7.3f} ms {f / t / 1e9 if f else 0:8.1f} TFLOP/s  {desc}", file=sys.stderr)

    t_all = torch.tensor([ms, ms_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t_all, op=dist.ReduceOp.MAX)
    ms, ms_e2e = float(t_all[0]), float(t_all[1])
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    out = {
        "metric": "roi_masks_per_sec", "value": world * n_rois / (ms * 1e-3), "unit": "ROI-masks/s", "n_gpus": world, "steps": args.steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands, f32 accumulate", "data": "synthetic", "config": workload_config(args.workload),
        "clocks": clocks,
        "e2e": {"value": world * n_rois / (ms_e2e * 1e-3), "unit": "ROI-masks/s", "ms_per_step": ms_e2e,
                "blocking_value": (world * n_rois / (ms_e2e_blocking * 1e-3)) if ms_e2e_blocking else None,
                "h2d_bytes_per_step": images_h.numel() * 4 + rois_h.numel() * 4, "d2h_bytes_per_step": inst_h.numel() * 4 + bin_h.numel() * 4,
                "api": ("model.infer(images_host_pinned, rois_host) -> (instance_masks, binary_masks) copied to pinned host" if args.no_pipeline else
                        "model.infer_pipelined(images_host_pinned, rois_host, instance_masks_host, binary_masks_host): upload / forward / "
                        "download of consecutive batches overlap on three streams, two launch plans; every step's H2D and D2H are inside the timed region; "
                        "exported-ONNX contract (masks + binary masks only: aux-only contour / distance branches are not computed, as in the exported graph)")},
        "gpu_launches": n_launches * args.steps,
        "launches_per_step": n_launches,
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_sm100_kernel", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                     "frac": achieved / pk["tflops"], "traffic": measured_traffic("conv_gemm_sm100_kernel:" + args.workload),
                     "traffic_note": "DRAM bytes per launch (read+write), averaged over the kernel's launches of one step, ncu capture in profiles/",
                     "peak_source": pk["source"] + ", sustained bf16",
                     "launches_per_step": len(gemm), "avg_launch_ms": gemm_ms / max(len(gemm), 1),
                     "algorithmic_flop_per_launch": gemm_flops / max(len(gemm), 1), "share_of_step": gemm_ms / step_ms_instr,
                     "how": "sum of algorithmic FLOPs of all conv_gemm launches of one step / sum of their CUDA-event durations "
                            "(instrumented replay immediately after the timed region)"},
        "step_algorithmic_tflop": step_flops / 1e12,
        "step_tflops": step_flops / (ms * 1e-3) / 1e12,
        "time_share_by_op": {k: round(v[1] / step_ms_instr, 4) for k, v in sorted(by_name.items(), key=lambda kv: -kv[1][1])},
    }
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        rate, t_iter = cpu_reference_rate(preset, h, w, 2, per_image, 2, threads)
        out["cpu_baseline"] = {"value": rate, "unit": "ROI-masks/s", "cores": threads, "kind": "port",
                               "sample": f"2 images x {per_image} ROIs of the same workload, min of 2 iterations ({t_iter:.2f} s each), fp32 torch CPU; "
                                         "oracle port of the reference modules (reference tree cannot travel)"}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
