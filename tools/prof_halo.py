"""Runs three halo-mode conv-GEMM shapes once each (for ncu): python tools/prof_halo.py"""
import sys
sys.path.insert(0, ".")
from tests.test_gpu_conv_gemm import run_case

SHAPES = [
    dict(n=8, h=480, w=640, cin=16, cout=16, k=3),        # UNet decoder tail (HBM-bound)
    dict(n=148, h=128, w=96, cin=128, cout=128, k=3),     # mask-res 3x3
    dict(n=148, h=80, w=60, cin=72, cout=72, k=3),        # B1 EnhancedUNet
]
for s in SHAPES:
    err, ref = run_case(**s)
    print(s, err, ref, flush=True)
