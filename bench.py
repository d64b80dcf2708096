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
    "b0_ln": ("b0", 64, 480, 640, 10),         # the b0 workload with the factory-default normalisation (LayerNorm2d, hed/model.py:18-38)
}
OVERRIDES = {"b0_ln": {"normalization_type": "layernorm2d"}}
HEAD_GFLOP_PER_ROI = {"b0": 57.84, "b1_enhanced": 93.01, "b7_ultra": 278.44}      # SURVEY §8d / BASELINE.md §2
UNET_GFLOP_PER_IMG = {"b0": 27.72, "b1_enhanced": 29.97, "b7_ultra": 90.21}


def measured_traffic(kernel: str):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (profiles/r2_traffic.json, else r1: dram__bytes_read.sum
    + dram__bytes_write.sum summed over the kernel's launches of one step / number of launches).  None if absent."""
    for name in ("r2_traffic.json", "r1_traffic.json"):
        try:
            v = json.load(open(os.path.join(ROOT, "profiles", name))).get(kernel)
            if v is not None:
                return v
        except Exception:
            pass
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
        self.t0 = self.t1 = None

    def mark_begin(self):
        """Start of the timed region (the sampler itself is started before the warm-up so that nvidia-smi is already polling)."""
        self.t0 = time.time()

    def mark_end(self):
        self.t1 = time.time()

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
            self.lines.append((time.time(), line.strip()))

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        if self.t1 is None:
            self.t1 = time.time()
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons, power = [], [], set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        # samples taken DURING the timed region; a region shorter than the polling period falls back to the samples under load since
        # the start of the warm-up (same kernels, same clocks) and says so
        window = "timed region"
        lines = [ln for t, ln in self.lines if self.t0 is None or (self.t0 - 0.02 <= t <= self.t1 + 0.12)]
        if not lines:
            lines, window = [ln for _, ln in self.lines], "warm-up + timed region (timed region shorter than the 100 ms polling period)"
        for ln in lines:
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
                "power_w_max": max(power) if power else None, "samples": len(sm), "window": window, "reasons": sorted(reasons)}


def synth_batch(seed, n_images, h, w, per_image):
    from human_instance_segmentation_b200.synthetic import synth_images, synth_rois
    return synth_images(seed, n_images, h, w), synth_rois(seed, n_images, per_image)


def build_model(preset, overrides=None):
    """The product arm: nothing under oracle/ is imported here."""
    import human_instance_segmentation_b200 as his
    from human_instance_segmentation_b200 import presets, synthetic
    kw = dict(presets.PRESETS[preset], **(overrides or {}))
    model = his.create_rgb_hierarchical_model(**kw)
    # random-init weights of that architecture (no checkpoints exist offline); BatchNorm statistics randomised so that
    # nothing folds to a no-op
    model.load_state_dict(synthetic.fill_state_dict(model.state_dict(), seed=0))
    return kw, model


def cpu_reference_rate(preset, h, w, n_images, per_image, iters, threads, masks_only=True):
    """ROI-masks/s of the reference path on host cores (oracle port of the reference modules, fp32, all threads).  masks_only:
    the exported-ONNX contract -- forward without the aux-only contour / distance branches + instance / binary masks -- i.e.
    the same work as the GPU arm's `e2e`."""
    import torch
    from oracle import headport, paramfill
    import human_instance_segmentation_b200 as his
    torch.set_num_threads(threads)
    cfg = headport.PRESETS[preset]
    holder = his.create_rgb_hierarchical_model(**cfg.factory_kwargs())           # parameter container only (key names/shapes)
    sd = paramfill.fill_state_dict(holder.state_dict(), seed=0)
    images, rois = synth_batch(1, n_images, h, w, per_image)

    def once():
        logits, aux = headport.forward(sd, images, rois, cfg, masks_only=masks_only)
        if masks_only:
            headport.export_outputs(logits, aux["full_image_logits"])

    once()                                                                       # warm-up
    times = []
    for _ in range(iters):
        t = time.perf_counter()
        once()
        times.append(time.perf_counter() - t)
    return rois.shape[0] / min(times), min(times)


def run_reference(args, out_stream=None):
    """Reference arm: the reference's CPU implementation of the path (oracle port -- the reference tree and smp/timm cannot travel
    to the GPU box) on all host cores, same contract as the GPU arm's `e2e`: images, rois -> instance_masks, binary_masks (the
    exported-ONNX outputs, export_onnx_advanced.py:353-457; aux-only branches not computed, as in the exported graph)."""
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

    def once():
        logits, aux = headport.forward(sd, images, rois, cfg, masks_only=True)
        return headport.export_outputs(logits, aux["full_image_logits"])

    for _ in range(max(1, min(args.warmup, 2))):
        once()
    steps = max(1, args.steps)
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        once()
        done += 1
        if time.perf_counter() - t0 > 150.0:          # bounded: the whole run ends within a few minutes whatever K is
            break
    dt = (time.perf_counter() - t0) / done
    value = rois.shape[0] / dt
    sample = f"{sample_images} images x {per_image} ROIs of the {args.workload} workload per step, fp32, torch CPU ({threads} threads)"
    (out_stream or sys.stdout).write(json.dumps({
        "impl": "reference", "metric": "roi_masks_per_sec", "value": value, "unit": "ROI-masks/s", "n_gpus": args.gpus, "steps": done,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic", "config": workload_config(args.workload, sample_images),
        "cpu_baseline": {"value": value, "unit": "ROI-masks/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "ROI-masks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0,
                "api": "oracle headport.forward(masks_only=True) + export_outputs: images, rois -> instance_masks, binary_masks "
                       "(the exported-ONNX contract, same work as the GPU arm's e2e)"},
        "note": "reference path = oracle port of the reference modules (the reference tree and smp/timm cannot travel to the GPU box); "
                "per-ROI throughput on a bounded sample of the workload"}) + "\n")


class Ctx:
    """Process-group plumbing shared by every measurement of one bench.py invocation (one process per GPU under torchrun)."""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize(self.dev)

    def max_over_ranks(self, values):
        t = self.torch.tensor(values, dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return [float(v) for v in t]

    def close(self):
        if self.world > 1:
            self.dist.destroy_process_group()


def run_post(ctx, steps, warmup, cpu_baseline=True):
    """BASELINE configs[4]: post-processing only.  A step = (a) the ROI chain on 10 ROIs per image -- MaskDilationModule on the
    [N,3,128,96] logits + argmax -> instance mask (u8) in one kernel, NEAREST paste-back onto the 480x640 label canvases -- and (b) the
    full-image mask clean-up -- BinaryMaskEdgeSmoothing + BinaryMaskBilateralFilter fused in one shared-memory pass over
    [B,1,480,640] masks.  value = masks (ROI masks + full-image masks) per second with inputs resident in HBM; the roofline is
    the fused stencil kernel against the measured HBM copy bandwidth (algorithmic bytes: read + write once, 2*H*W*4 per mask)."""
    torch = ctx.torch
    from human_instance_segmentation_b200 import postprocess as pp
    from human_instance_segmentation_b200.synthetic import synth_rois
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    B, H, W, per_image, mh, mw = 512, 480, 640, 10, 128, 96
    g = torch.Generator().manual_seed(5 + rank)
    full_h = (torch.nn.functional.avg_pool2d((torch.rand(B, 1, H // 4, W // 4, generator=g) > 0.5).float(), 5, 1, 2) > 0.5).float()
    full_h = torch.nn.functional.interpolate(full_h, size=(H, W), mode="nearest").contiguous().pin_memory()     # blob-like masks
    rois_h = synth_rois(5 + rank, B, per_image).pin_memory()
    N = rois_h.shape[0]
    logits_h = (torch.randn(N, 3, mh // 8, mw // 8, generator=g) * 2)
    logits_h = torch.nn.functional.interpolate(logits_h, size=(mh, mw), mode="bilinear").contiguous().pin_memory()
    full, rois, logits = full_h.to(dev), rois_h.to(dev), logits_h.to(dev)
    cleanup = pp.MaskCleanup().to(dev)
    out_full = torch.empty_like(full)

    def step(f, r, lg, out=out_full):
        masks = pp.instance_masks(lg, as_uint8=True, dilation_pixels=1)       # MaskDilationModule(1) + argmax, one pass
        canvas = pp.paste_masks(masks, r, B, H, W)
        return cleanup(f, out=out), canvas

    warmup = max(warmup, 3)
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    for _ in range(warmup):
        step(full, rois, logits)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record()
    for _ in range(steps):
        step(full, rois, logits)
    e1.record()
    ctx.barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1) / steps
    clocks = sampler.stop() if rank == 0 else None
    # fused stencil alone (the dominant kernel), event-timed
    e0.record()
    for _ in range(steps):
        cleanup(full, out=out_full)
    e1.record()
    ctx.barrier()
    ms_fused = e0.elapsed_time(e1) / steps
    # e2e: pinned host in, cleaned masks + canvases back to pinned host
    res_h = torch.empty_like(full_h).pin_memory()
    canvas_h = torch.empty((B, H, W), dtype=torch.int32).pin_memory()

    def e2e_step():
        o, c = step(full_h.to(dev, non_blocking=True), rois_h.to(dev, non_blocking=True), logits_h.to(dev, non_blocking=True))
        res_h.copy_(o, non_blocking=True); canvas_h.copy_(c, non_blocking=True)

    e2e_step(); ctx.barrier()
    e0.record()
    for _ in range(steps):
        e2e_step()
    e1.record()
    ctx.barrier()
    ms_e2e = e0.elapsed_time(e1) / steps
    # the same step with the full-image masks as bytes on both sides (MaskCleanup's uint8 variant: a quarter of the mask bytes over PCIe)
    full8_h = full_h.to(torch.uint8).pin_memory()
    res8_h = torch.empty_like(full8_h).pin_memory()

    def e2e_step_u8():
        o, c = step(full8_h.to(dev, non_blocking=True), rois_h.to(dev, non_blocking=True), logits_h.to(dev, non_blocking=True), out=None)
        res8_h.copy_(o, non_blocking=True); canvas_h.copy_(c, non_blocking=True)

    e2e_step_u8(); ctx.barrier()
    same_u8 = bool(torch.equal(res8_h, res_h.to(torch.uint8)))
    e0.record()
    for _ in range(steps):
        e2e_step_u8()
    e1.record()
    ctx.barrier()
    ms_e2e_u8 = e0.elapsed_time(e1) / steps
    ms, ms_e2e, ms_fused, ms_e2e_u8 = ctx.max_over_ranks([ms, ms_e2e, ms_fused, ms_e2e_u8])
    if rank != 0:
        return None
    pk = peaks()
    units = B + N
    alg_bytes = 2.0 * B * H * W * 4
    achieved = alg_bytes / (ms_fused * 1e-3) / 1e9
    out = {"metric": "roi_masks_per_sec", "value": world * units / (ms * 1e-3), "unit": "masks/s", "n_gpus": world, "steps": steps, "warmup": warmup,
           "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32 (masks), u8/int32 (paste-back)",
           "data": "synthetic",
           "config": {"workload": f"post-processing only (BASELINE configs[4]): {B} full-image 640x480 masks (edge smoothing + binary bilateral, fused) + "
                                  f"{N} ROI logits 128x96 (dilation, argmax, NEAREST paste-back) per GPU per step", "masks_per_gpu": units,
                      "cache": f"inputs larger than L2 ({int(alg_bytes / 2e6)} MB of masks, {int(N * 3 * mh * mw * 4 / 1e6)} MB of logits per step)"},
           "clocks": clocks,
           "e2e": {"value": world * units / (ms_e2e * 1e-3), "unit": "masks/s", "ms_per_step": ms_e2e,
                   "h2d_bytes_per_step": full_h.numel() * 4 + logits_h.numel() * 4 + rois_h.numel() * 4,
                   "d2h_bytes_per_step": res_h.numel() * 4 + canvas_h.numel() * 4,
                   "api": "postprocess.instance_masks(dilation_pixels=1) / paste_masks / MaskCleanup on pinned host tensors",
                   "u8_masks": {"value": world * units / (ms_e2e_u8 * 1e-3), "ms_per_step": ms_e2e_u8,
                                "h2d_bytes_per_step": full8_h.numel() + logits_h.numel() * 4 + rois_h.numel() * 4,
                                "d2h_bytes_per_step": res8_h.numel() + canvas_h.numel() * 4, "equal_to_fp32_result": same_u8,
                                "what": "the same step with the full-image masks as uint8 on the host and on the device (MaskCleanup's byte variant)"}},
           "gpu_launches": 3 * steps, "launches_per_step": 4,      # dilate+argmax, paste, clean-up (ours) + the canvas memset (torch)
           "roofline": {"bound": "hbm", "kernel": "mask_cleanup_wide_kernel<7,2,float>", "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s",
                        "frac": achieved / pk["hbm_gbs"], "traffic": measured_traffic("mask_cleanup_wide_kernel:post"), "peak_source": pk["source"],
                        "avg_launch_ms": ms_fused,
                        "algorithmic_bytes_per_launch": alg_bytes, "share_of_step": ms_fused / ms,
                        "how": "2*H*W*4 bytes per mask (read once + write once) x masks per launch / CUDA-event time of the launch"}}
    if cpu_baseline and world == 1:
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
    return out


def workload_config(name, n_img=None, per_image=None):
    preset, b, h, w, ppi = WORKLOADS[name]
    b = n_img or b
    per_image = per_image or ppi
    from human_instance_segmentation_b200 import presets
    kw = dict(presets.PRESETS[preset], **OVERRIDES.get(name, {}))
    rs, msz = kw["roi_size"], kw["mask_size"]
    extra = "".join(f", {k}={v}" for k, v in OVERRIDES.get(name, {}).items())
    return {"workload": f"{preset} hierarchical RGB model{extra}, {w}x{h} input, batch {b} per GPU, {per_image} ROIs/image, ROI {rs[0]}x{rs[1]} -> mask "
                        f"{msz[0]}x{msz[1]}, random-init weights", "images_per_gpu": b, "rois_per_gpu": b * per_image,
            "encoder": kw["encoder_name"], "cache": "inputs larger than L2 (images + multi-GB activations per step)"}


def _set_export_scale(model, h, w):
    for ra in (model.roi_align_mask, model.roi_align_rgb):       # exporter convention spatial_scale=(H,W)
        ra.spatial_scale = (float(h), float(w)); ra.spatial_scale_h, ra.spatial_scale_w = float(h), float(w)


def measure_model(ctx, workload, steps, warmup, precision="fast", e2e="pipelined+blocking", breakdown=False, top=45, use_graph=True,
                  n_img=None, per_image=None, model=None):
    """One workload on this rank's GPU.  Returns the record (rank 0) or None.  e2e: "pipelined+blocking" | "pipelined" | "blocking" | None."""
    torch = ctx.torch
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    warmup = max(warmup, 3)
    preset, b_def, h, w, ppi_def = WORKLOADS[workload]
    n_img, per_image = n_img or b_def, per_image or ppi_def
    if model is None:
        cfg, model = build_model(preset, OVERRIDES.get(workload))
        model = model.to(dev)
    else:
        from human_instance_segmentation_b200 import presets
        cfg = presets.PRESETS[preset]
    _set_export_scale(model, h, w)
    model.precision = precision
    model.aux_outputs = "full"
    model.copy_outputs = False
    model.use_cuda_graph = use_graph          # launch-bound tail of ~200 small kernels -> one graph submission per step
    images_h, rois_h = synth_batch(100 + rank, n_img, h, w, per_image)
    images_h, rois_h = images_h.pin_memory(), rois_h.pin_memory()
    images_d, rois_d = images_h.to(dev), rois_h.to(dev)
    n_rois = rois_h.shape[0]

    # ---------------- value: inputs resident in HBM
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    for _ in range(warmup):
        model(images_d, rois_d)
    bp = model._get_plan(images_d, rois_d)
    plan = bp.plan
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record()
    for _ in range(steps):
        model(images_d, rois_d)
    e1.record()
    ctx.barrier()
    sampler.mark_end()
    ms = e0.elapsed_time(e1) / steps
    clocks = sampler.stop() if rank == 0 else None

    # ---------------- per-kernel timing for the roofline (instrumented replay right after the timed region)
    timed = plan.run_timed()
    timed = plan.run_timed()
    n_launches, step_flops, op_desc = plan.launches, plan.flops, list(plan.op_desc)
    del plan, bp
    model.release_plans()          # the e2e arm builds its own (masks-only) plans; keep the HBM footprint to two of them

    # ---------------- e2e: host (pinned) buffers in, instance/binary masks back to host (the exported-ONNX contract)
    ms_e2e = ms_e2e_blocking = ms_e2e_full = None
    inst_h = torch.empty((n_rois, 1) + tuple(cfg["mask_size"]), dtype=torch.float32).pin_memory()
    bin_h = torch.empty((n_img, 1, h, w), dtype=torch.float32).pin_memory()
    if e2e and "pipelined" in e2e:
        # two sets of host output buffers: the pipelined API downloads batch i-1 while batch i computes
        host_out = [(inst_h, bin_h), (torch.empty_like(inst_h).pin_memory(), torch.empty_like(bin_h).pin_memory())]
        for i in range(2):
            model.infer_pipelined(images_h, rois_h, *host_out[i & 1])
        model.pipeline_sync()
        ctx.barrier()
        e0.record()
        for i in range(steps):
            model.infer_pipelined(images_h, rois_h, *host_out[i & 1])
        model.pipeline_sync()      # every batch of the timed region has landed in host memory (current stream waits for the downloads)
        e1.record()
        ctx.barrier()
        ms_e2e = e0.elapsed_time(e1) / steps
    if e2e and "blocking" in e2e:
        # the same through the blocking call (upload, forward, download strictly one after the other)
        def blocking_step():
            inst, binary = model.infer(images_h, rois_h)
            inst_h.copy_(inst, non_blocking=True)
            bin_h.copy_(binary, non_blocking=True)
        blocking_step(); ctx.barrier()
        e0.record()
        for _ in range(steps):
            blocking_step()
        e1.record()
        ctx.barrier()
        ms_e2e_blocking = e0.elapsed_time(e1) / steps
        if ms_e2e is None:
            ms_e2e = ms_e2e_blocking
    if e2e and "full" in e2e:
        # the FULL reference forward (11 aux outputs computed, like `value`) from host memory, logits + binary masks back to host
        model.release_plans()
        logits_h = torch.empty((n_rois, 3) + tuple(cfg["mask_size"]), dtype=torch.float32).pin_memory()

        def full_step():
            logits, aux = model(images_h.to(dev, non_blocking=True), rois_h.to(dev, non_blocking=True))
            logits_h.copy_(logits, non_blocking=True)
            bin_h.copy_(model._get_plan(images_d, rois_d).out_binary, non_blocking=True)
        full_step(); ctx.barrier()
        e0.record()
        for _ in range(steps):
            full_step()
        e1.record()
        ctx.barrier()
        ms_e2e_full = e0.elapsed_time(e1) / steps
    plans_built, plan_gb, plan_log = getattr(model, "plans_built", 0), model.plan_bytes() / 1e9, list(getattr(model, "plan_log", []))
    model.release_plans()

    gemm = [(t, f) for name, t, f in timed if name == "conv_gemm"]
    gemm_ms, gemm_flops = sum(t for t, _ in gemm), sum(f for _, f in gemm)
    step_ms_instr = sum(t for _, t, _ in timed)
    by_name, by_tag = {}, {}
    for i, (name, t, f) in enumerate(timed):
        by_name.setdefault(name, [0, 0.0]); by_name[name][0] += 1; by_name[name][1] += t
        tag = op_desc[i].split(":", 1)[0]
        by_tag[tag] = by_tag.get(tag, 0.0) + t
    if breakdown and rank == 0:
        for name, (cnt, t) in sorted(by_name.items(), key=lambda kv: -kv[1][1]):
            print(f"  {name:22s} x{cnt:4d} {t:9.3f} ms {100 * t / step_ms_instr:5.1f}%", file=sys.stderr)
        for tag, t in sorted(by_tag.items(), key=lambda kv: -kv[1]):
            print(f"  [{tag}] {t:9.3f} ms", file=sys.stderr)
        rows = sorted(((t, op_desc[i], f) for i, (n_, t, f) in enumerate(timed)), reverse=True)
        for t, desc, f in rows[:top]:
            print(f"  {t:7.3f} ms {f / t / 1e9 if f else 0:8.1f} TFLOP/s  {desc}", file=sys.stderr)

    vals = ctx.max_over_ranks([ms, ms_e2e or 0.0, ms_e2e_blocking or 0.0, ms_e2e_full or 0.0])
    ms, ms_e2e, ms_e2e_blocking, ms_e2e_full = vals[0], vals[1] or None, vals[2] or None, vals[3] or None
    if rank != 0:
        return None
    pk = peaks()
    achieved = gemm_flops / (gemm_ms * 1e-3) / 1e12 if gemm_ms > 0 else 0.0
    mma_passes = 3 if precision == "strict" else 1
    out = {
        "metric": "roi_masks_per_sec", "value": world * n_rois / (ms * 1e-3), "unit": "ROI-masks/s", "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f16 operands, f32 accumulate" if precision == "fast" else "split f16 (hi+lo) operands, 3 MMA passes, f32 accumulate",
        "precision": precision, "data": "synthetic", "config": workload_config(workload, n_img, per_image), "clocks": clocks,
        "gpu_launches": n_launches * steps, "launches_per_step": n_launches,
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_sm100_kernel", "achieved": achieved, "peak": pk["tflops"], "unit": "TFLOP/s",
                     "frac": achieved / pk["tflops"], "traffic": measured_traffic("conv_gemm_sm100_kernel:" + workload) if precision == "fast" else None,
                     "traffic_note": "DRAM bytes per launch (read+write), averaged over the kernel's launches of one step, ncu capture in profiles/",
                     "peak_source": pk["source"] + ", sustained bf16",
                     "launches_per_step": len(gemm), "avg_launch_ms": gemm_ms / max(len(gemm), 1),
                     "algorithmic_flop_per_launch": gemm_flops / max(len(gemm), 1), "share_of_step": gemm_ms / step_ms_instr,
                     "mma_passes": mma_passes, "issued_frac": mma_passes * achieved / pk["tflops"],
                     "how": "sum of ALGORITHMIC FLOPs (2*MAC of the reference conv, counted once whatever the number of MMA passes) of all conv_gemm "
                            "launches of one step / sum of their CUDA-event durations (instrumented replay immediately after the timed region)"},
        "step_algorithmic_tflop": step_flops / 1e12,
        "step_tflops": step_flops / (ms * 1e-3) / 1e12,
        "time_share_by_op": {k: round(v[1] / step_ms_instr, 4) for k, v in sorted(by_name.items(), key=lambda kv: -kv[1][1])},
        "ms_by_subplan": {k: round(v, 3) for k, v in by_tag.items()},
    }
    if ms_e2e is not None:
        out["e2e"] = {
            "value": world * n_rois / (ms_e2e * 1e-3), "unit": "ROI-masks/s", "ms_per_step": ms_e2e,
            "blocking_value": (world * n_rois / (ms_e2e_blocking * 1e-3)) if ms_e2e_blocking else None,
            "full_forward_value": (world * n_rois / (ms_e2e_full * 1e-3)) if ms_e2e_full else None,
            "h2d_bytes_per_step": images_h.numel() * 4 + rois_h.numel() * 4, "d2h_bytes_per_step": inst_h.numel() * 4 + bin_h.numel() * 4,
            "launch_plans_built_in_run": plans_built, "launch_plan_gb_at_end": round(plan_gb, 2), "launch_plan_log": plan_log,
            "api": ("model.infer_pipelined(images_host_pinned, rois_host, instance_masks_host, binary_masks_host): upload / forward / download of "
                    "consecutive batches overlap on three streams, two launch plans; every step's H2D and D2H are inside the timed region" if "pipelined" in e2e
                    else "model.infer(images_host_pinned, rois_host) -> (instance_masks, binary_masks) copied to pinned host") +
                   "; work = the exported-ONNX contract (masks + binary masks; the aux-only contour / distance branches are not computed, as in the "
                   "exported graph) -- the SAME contract the reference arm (--impl reference) and cpu_baseline time; `value` and "
                   "e2e.full_forward_value compute the full reference forward with its 11 aux outputs"}
    return out


def measure_latency_cfg0(ctx, steps, warmup):
    """BASELINE configs[0] literally: B0, batch 2, 8 ROIs -- a latency point (one small batch per call)."""
    rec = measure_model(ctx, "b0", steps, warmup, e2e="blocking", n_img=2, per_image=4)
    if rec is None:
        return None
    rec["latency_ms"] = rec["ms_per_step"]
    rec["e2e_latency_ms"] = rec["e2e"]["ms_per_step"]
    rec["config"]["workload"] = "BASELINE configs[0]: " + rec["config"]["workload"] + " (latency of one call)"
    return rec


def measure_strong_b7(ctx, steps, warmup):
    """BASELINE configs[2]: B7 ultra, ONE global batch of 32 images (uneven ROI counts, 320 ROIs) sharded over the ranks with
    sharding.partition_images / shard_batch (no collective on the path).  After the timed region the per-rank logits are gathered
    over NCCL (sharding.gather_logits) and compared bit for bit with rank 0 running the whole batch alone."""
    torch = ctx.torch
    from human_instance_segmentation_b200 import sharding
    from human_instance_segmentation_b200.synthetic import synth_images, synth_rois
    world, rank, dev = ctx.world, ctx.rank, ctx.dev
    n_img, h, w = 32, 480, 640
    counts = [4 + (i * 7) % 13 for i in range(n_img)]          # 4..16 ROIs per image
    counts[-1] += 320 - sum(counts)
    images = synth_images(300, n_img, h, w)
    pool = synth_rois(300, n_img, max(counts))
    per = max(counts)
    rois = torch.cat([pool[i * per:i * per + c] for i, c in enumerate(counts)], 0)
    bounds = sharding.partition_images(rois[:, 0], n_img, world)
    im_s, rois_s, sel = sharding.shard_batch(images, rois, world, rank, bounds)
    cfg, model = build_model("b7_ultra")
    model = model.to(dev)
    _set_export_scale(model, h, w)
    model.copy_outputs, model.use_cuda_graph, model.aux_outputs = False, True, "full"
    im_d, rois_d = im_s.to(dev), rois_s.to(dev)
    sampler = ClockSampler(ctx.local)
    if rank == 0:
        sampler.start()
    for _ in range(max(warmup, 3)):
        model(im_d, rois_d)
    ctx.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.mark_begin()
    e0.record()
    for _ in range(steps):
        logits_local, _ = model(im_d, rois_d)
    e1.record()
    ctx.barrier()
    sampler.mark_end()
    ms = ctx.max_over_ranks([e0.elapsed_time(e1) / steps])[0]
    clocks = sampler.stop() if rank == 0 else None
    check = None
    if world > 1:
        gathered = sharding.gather_logits(logits_local.clone(), sel.to(dev), rois.shape[0])
        model.release_plans()
        if rank == 0:
            whole, _ = model(images.to(dev), rois.to(dev))
            diff = (gathered - whole).abs().max().item() if rois.shape[0] else 0.0
            check = {"bit_identical": bool(torch.equal(gathered, whole)), "max_abs_diff": diff, "rois": int(rois.shape[0]),
                     "how": "sharding.gather_logits (NCCL all_gather) of the per-rank logits vs rank 0 running the whole 32-image batch alone"}
    model.release_plans()
    del model
    torch.cuda.empty_cache()
    if rank != 0:
        return None
    return {"metric": "roi_masks_per_sec", "value": rois.shape[0] / (ms * 1e-3), "unit": "ROI-masks/s", "n_gpus": world, "steps": steps,
            "ms_per_step": ms, "scaling": "strong", "clocks": clocks,
            "config": {"workload": "BASELINE configs[2]: b7_ultra, ONE global batch of 32 images 640x480 with 4..16 ROIs each (320 ROIs), ROI 128x96 -> mask "
                                   "256x192, sharded over the ranks by sharding.partition_images (balanced on ROI count)",
                       "images_per_rank": [hi - lo for lo, hi in bounds],
                       "rois_per_rank": [int(((rois[:, 0] >= lo) & (rois[:, 0] < hi)).sum()) for lo, hi in bounds]},
            "gather_check": check}


def main():
    # exactly ONE line on stdout: libraries that print to fd 1 (NCCL's version banner ...) are sent to stderr, the JSON line is
    # written to the real stdout at the end
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = sys.stderr
    try:
        _main(real_stdout)
    finally:
        real_stdout.flush()


def _main(real_stdout):
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="b0", choices=list(WORKLOADS) + ["post", "strong_b7", "cfg0"])
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--precision", default="fast", choices=["fast", "strict"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--quick", action="store_true", help="headline line only: no strict-mode / per-config sub-records")
    ap.add_argument("--breakdown", action="store_true", help="print the per-op time table of one instrumented step to stderr")
    ap.add_argument("--top", type=int, default=45)
    ap.add_argument("--no-pipeline", action="store_true", help="e2e through the blocking model.infer instead of infer_pipelined")
    ap.add_argument("--no-graph", action="store_true", help="replay the launch plan kernel by kernel instead of as one CUDA graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args, real_stdout)
    ctx = Ctx()
    try:
        if args.workload == "post":
            out = run_post(ctx, args.steps, args.warmup, not args.no_cpu_baseline)
        elif args.workload == "strong_b7":
            out = measure_strong_b7(ctx, args.steps, args.warmup)
        elif args.workload == "cfg0":
            out = measure_latency_cfg0(ctx, args.steps, args.warmup)
        else:
            out = measure_model(ctx, args.workload, args.steps, args.warmup, args.precision,
                                e2e="blocking" if args.no_pipeline else "pipelined+blocking+full", breakdown=args.breakdown, top=args.top,
                                use_graph=not args.no_graph)
            headline = args.workload == "b0" and args.precision == "fast" and not args.quick
            if headline:
                # ---- the other precision mode and the other BASELINE configs as short sub-records of the same JSON line
                sub_steps = 3
                strict = measure_model(ctx, "b0", sub_steps, 3, "strict", e2e="pipelined")
                if ctx.rank == 0:
                    out["precision_modes"] = {
                        "fast": {"value": out["value"], "ms_per_step": out["ms_per_step"], "e2e": out["e2e"]["value"],
                                 "tolerance": "stress weights: L2 <= 2.5e-3, max <= 6e-3, argmax >= 99.8 % (tests/test_gpu_model.py)"},
                        "strict": {"value": strict["value"], "ms_per_step": strict["ms_per_step"], "e2e": strict["e2e"]["value"],
                                   "roofline": strict["roofline"], "clocks": strict["clocks"],
                                   "tolerance": "north_star: max-rel <= 1e-3, argmax >= 99.9 % on every golden (tests/test_gpu_model.py)"}}
                if ctx.world == 1:
                    subs = {"cfg0_b0_batch2_8rois": measure_latency_cfg0(ctx, sub_steps, 3)}
                    for wl in ("b1", "b7", "b0_160x120", "b0_ln"):
                        subs[wl] = measure_model(ctx, wl, sub_steps, 3, "fast", e2e="pipelined")
                    subs["post"] = run_post(ctx, sub_steps, 3, cpu_baseline=False)
                    out["configs"] = subs
                out_strong = measure_strong_b7(ctx, sub_steps, 3)
                if ctx.rank == 0:
                    out["strong_b7"] = out_strong
            if ctx.rank == 0 and not args.no_cpu_baseline and ctx.world == 1:
                preset, _, h, w, per_image = WORKLOADS[args.workload]
                threads = os.cpu_count() or 1
                rate, t_iter = cpu_reference_rate(preset, h, w, 2, per_image, 2, threads)
                out["cpu_baseline"] = {"value": rate, "unit": "ROI-masks/s", "cores": threads, "kind": "port",
                                       "sample": f"2 images x {per_image} ROIs of the same workload, min of 2 iterations ({t_iter:.2f} s each), fp32 torch CPU; "
                                                 "oracle port of the reference modules (reference tree cannot travel), exported-ONNX contract "
                                                 "(same work as e2e)"}
        if ctx.rank == 0 and out is not None:
            real_stdout.write(json.dumps(out) + "\n")
    finally:
        ctx.close()


if __name__ == "__main__":
    main()
