import ctypes, sys
import torch
sys.path.insert(0, ".")
from tests import common
from tests.test_gpu_model import build

cfg, images, rois = common.small_case_inputs("small_b0_bn_relu")
m = build(cfg, common.shapes_for_case("small_b0_bn_relu"))
im, r = images.cuda(), rois.cuda()
m(im, r)
bp = m._get_plan(im, r)
plan = bp.plan
tensors = [t for t in plan.keep if isinstance(t, torch.Tensor)]
s = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

def run():
    sums = []
    for name, fn, args in plan.ops:
        fn(*args, s)
        torch.cuda.synchronize()
        sums.append(torch.stack([torch.nan_to_num(t.double(), nan=0.0, posinf=0.0, neginf=0.0).abs().sum() for t in tensors]).cpu())
    return sums

a = run(); b = run(); c = run()
found = False
for i in range(len(plan.ops)):
    prev_b = b[i - 1] if i else a[-1]
    prev_c = c[i - 1] if i else b[-1]
    written = ((b[i] != prev_b) | (c[i] != prev_c)).nonzero().flatten().tolist()
    bad = [j for j in written if b[i][j] != c[i][j]]
    if bad:
        name = plan.ops[i][0]
        print(f"op {i} ({name}) wrote different values in two runs; buffers {bad[:5]} shapes {[tuple(tensors[j].shape) for j in bad[:5]]} rel diff {[float(abs(b[i][j]-c[i][j])/abs(b[i][j])) for j in bad[:5]]}")
        if name == "conv_gemm":
            gi = [k for k, o in enumerate(plan.ops[:i + 1]) if o[0] == "conv_gemm"]
            print("   gemm shape", plan.gemm_shapes[len(gi) - 1])
        print("   prev ops:", [o[0] for o in plan.ops[max(0, i - 6):i + 1]])
        found = True
        break
if not found:
    print("deterministic")
