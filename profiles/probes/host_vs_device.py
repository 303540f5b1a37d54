"""How long does the host take to ENQUEUE one Seg / Det iteration, against the device time of the same iteration?
(decides whether graph capture of those steps can pay)."""
import importlib, os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
PKG = importlib.import_module("cv_a-fan_b200")
torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")


def measure(name, step, iters=6):
    for _ in range(3):
        step()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    host = 0.0
    a.record()
    for _ in range(iters):
        t0 = time.perf_counter()
        step()
        host += time.perf_counter() - t0
    b.record(); b.synchronize()
    print(name, {"host_enqueue_ms": round(host / iters * 1e3, 1), "device_ms": round(a.elapsed_time(b) / iters, 1)}, flush=True)


g = torch.Generator().manual_seed(3)
B, S, NC = 4, 513, 19
images = torch.rand(B, 3, S, S, generator=g).to(dev)
labels = torch.randint(0, NC, (B, S, S), generator=g).to(dev)
torch.manual_seed(3)
m = PKG.deeplab.deeplabv3plus_resnet101(num_classes=NC, output_stride=16).to(dev).train()
tr = PKG.trainer_seg.SegAfanTrainer(m, pertub_idx_se=4, pertub_idx_sd="aspp", steps=1, eps=2.0, gamma_se=0.5, gamma_sd=0.5, randinit=True,
                                    clip=False, mix_sd=True, noise_sd=0.0, mix_layer="01", head_cache=True, dual_bn=True)
measure("seg cfg5", lambda: tr.step(images, labels))
del m, tr
B, H, W, NC, G = 8, 600, 1000, 21, 8
images = torch.rand(B, 3, H, W, generator=g).to(dev)
x0, y0 = torch.rand(B, G, generator=g) * (W - 320), torch.rand(B, G, generator=g) * (H - 320)
bw, bh = 50 + torch.rand(B, G, generator=g) * 250, 50 + torch.rand(B, G, generator=g) * 250
boxes = torch.stack((x0, y0, x0 + bw, y0 + bh), dim=2).to(dev)
labels = torch.randint(1, NC, (B, G), generator=g).to(dev)
torch.manual_seed(3)
m = PKG.faster_rcnn.FasterRCNN(NC, sampler="device").to(dev)
for n_, p in m.named_parameters():
    if ("_anchor_" in n_ or "_proposal_" in n_) and n_.endswith("weight"):
        p.data.mul_(0.01)
    if n_.endswith("bn3.weight"):
        p.data.fill_(0.25)
tr = PKG.trainer_det.DetAfanTrainer(m, pertub_idx_se=3, randinit=True, mix_layer="0101", mix_sd=True, lr=1e-4, rng="philox", seed=3)
measure("det cfg4", lambda: tr.step(images, boxes, labels))
