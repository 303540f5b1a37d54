"""Kernel-time table of one config-4 Detection iteration (torch.profiler, CUDA activities), tf32 or fp32 convolutions."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
PKG = importlib.import_module("cv_a-fan_b200")
tf32 = "--tf32" in sys.argv
torch.backends.cudnn.allow_tf32 = tf32
torch.backends.cuda.matmul.allow_tf32 = tf32
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
B, H, W, NC, G = 8, 600, 1000, 21, 8
g = torch.Generator().manual_seed(3)
images = torch.rand(B, 3, H, W, generator=g).to(dev)
x0, y0 = torch.rand(B, G, generator=g) * (W - 320), torch.rand(B, G, generator=g) * (H - 320)
bw, bh = 50 + torch.rand(B, G, generator=g) * 250, 50 + torch.rand(B, G, generator=g) * 250
boxes = torch.stack((x0, y0, x0 + bw, y0 + bh), dim=2).to(dev)
labels = torch.randint(1, NC, (B, G), generator=g).to(dev)
torch.manual_seed(3)
m = PKG.faster_rcnn.FasterRCNN(NC, sampler="device", fuse_frozen_bn="--nofuse" not in sys.argv).to(dev)
if "--cl" in sys.argv:
    m = m.to(memory_format=torch.channels_last)
    images = images.contiguous(memory_format=torch.channels_last)
for n_, p in m.named_parameters():
    if ("_anchor_" in n_ or "_proposal_" in n_) and n_.endswith("weight"):
        p.data.mul_(0.01)
    if n_.endswith("bn3.weight"):
        p.data.fill_(0.25)
tr = PKG.trainer_det.DetAfanTrainer(m, pertub_idx_se=3, randinit=True, mix_layer="0101", mix_sd=True, lr=1e-4, rng="philox", seed=3)
for _ in range(3):
    tr.step(images, boxes, labels)
torch.cuda.synchronize()
t0 = time.time()
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA, torch.profiler.ProfilerActivity.CPU]) as prof:
    tr.step(images, boxes, labels)
    torch.cuda.synchronize()
print("wall ms", (time.time() - t0) * 1e3)
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=45, max_name_column_width=70))
