import sys
import torch
sys.path.insert(0, ".")
import human_instance_segmentation_b200 as his
from tests import common
from tests.test_gpu_model import build, l2_rel

cfg, images, rois = common.small_case_inputs("small_b0_bn_relu")
m = build(cfg, common.shapes_for_case("small_b0_bn_relu"))
im, r = images.cuda(), rois.cuda()
outs = []
for i in range(3):
    lg, aux = m(im, r)
    outs.append((lg.clone(), {k: v.clone() for k, v in aux.items()}))
torch.cuda.synchronize()
for i in (1, 2):
    print("same plan run0 vs run%d:" % i, l2_rel(outs[i][0].cpu(), outs[0][0].cpu()))
    for k in outs[0][1]:
        d = l2_rel(outs[i][1][k].cpu(), outs[0][1][k].cpu())
        if d > 0:
            print("    ", k, d)
m.invalidate()
lg2, aux2 = m(im, r)
print("fresh plan vs run0:", l2_rel(lg2.cpu(), outs[0][0].cpu()))
for k in aux2:
    d = l2_rel(aux2[k].cpu(), outs[0][1][k].cpu())
    if d > 0:
        print("    ", k, d)
m.invalidate(); m.use_cuda_graph = True
lg3, aux3 = m(im, r)
print("graph plan vs run0:", l2_rel(lg3.cpu(), outs[0][0].cpu()))
for k in aux3:
    d = l2_rel(aux3[k].cpu(), outs[0][1][k].cpu())
    if d > 0:
        print("    ", k, d)
