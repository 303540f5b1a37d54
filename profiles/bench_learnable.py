"""Learnable-eta A-FAN iteration time on ResNet-56, batch 128, 9 perturbation layers, PGD-3 + clip:
reference pass order (9 sequential ascents, eager) vs batched ascents (eager) vs batched + CUDA graph."""
import importlib
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cv_a-fan_b200")
dev = torch.device("cuda:0")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.benchmark = True
g = torch.Generator().manual_seed(3)
x = torch.rand(128, 3, 32, 32, generator=g).to(dev)
y = torch.randint(0, 10, (128,), generator=g).to(dev)
res = {}
for name, kw in (("reference_order_eager", dict(batched=False, use_cuda_graph=False)),
                 ("batched_eager", dict(batched=True, use_cuda_graph=False)),
                 ("batched_graph", dict(batched=True, use_cuda_graph=True))):
    torch.manual_seed(3)
    model = pkg.resnet_s.resnet56(init_weight_eta=1 / 9).to(dev)
    tr = pkg.trainer_learnable.LearnableEtaTrainer(model, steps=3, gamma=1.0, eps=2.0, randinit=True, clip=True, **kw)
    for _ in range(4):
        tr.step(x, y)
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(10):
        tr.step(x, y)
    e.record()
    e.synchronize()
    res[name] = {"ms_per_iter": s.elapsed_time(e) / 10, "img_per_s": 128 * 10 / (s.elapsed_time(e) * 1e-3),
                 "afan_kernels_per_iter": tr.kernel_launches_per_iter}
    tr.close()
print(json.dumps(res))
