"""Runs a few conv-GEMM shapes back to back (for ncu): python tools/prof_gemm.py [reps]"""
import sys
import torch
sys.path.insert(0, ".")
from tests.test_gpu_conv_gemm import run_case
from human_instance_segmentation_b200.engine import RES_MUL, RES_ADD

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
SHAPES = [
    dict(n=148, h=64, w=48, cin=256, cout=256, k=3),                            # hot 3x3
    dict(n=148, h=64, w=48, cin=256, cout=256, k=3, res_mode=RES_ADD),          # hot 3x3 + residual
    dict(n=148, h=64, w=48, cin=128, cout=256, k=1, act=3, res_mode=RES_MUL),   # 1x1 gate
    dict(n=148, h=128, w=96, cin=128, cout=128, k=3),                           # mask-res 3x3
    dict(n=8, h=480, w=640, cin=16, cout=16, k=3),                              # UNet decoder tail
]
for _ in range(reps):
    for s in SHAPES:
        err, ref = run_case(**s)
        print(s, err, ref, flush=True)
