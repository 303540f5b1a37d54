"""Config 4 (BASELINE.json configs[3]): Faster R-CNN ResNet-101 C4, VOC-shaped synthetic 600x1000 images, feature PGD on the
backbone layer3 output (B x 1024 x 38 x 63) + ROI-side PGD, one A-FAN iteration of Detection/train_aug_final.py:78-163
(PGD-1, 5 SAT points, masked mix_feature, six training forwards, SGD).  Times (CUDA events, eager launches):
  afan_literal       DetAfanTrainer(head_cache=False): the reference's schedule (three backbone sweeps, two RPN + NMS passes
                     on the clean images) on this package's model and kernels
  afan               head_cache=True, CPU-stream sampling ('reference') and on-device sampling ('device')
Usage: python profiles/bench_det.py [--tf32] [--batch B]"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
PKG = importlib.import_module("cv_a-fan_b200")
tf32 = "--tf32" in sys.argv
B = int(sys.argv[sys.argv.index("--batch") + 1]) if "--batch" in sys.argv else 8
torch.backends.cudnn.allow_tf32 = tf32
torch.backends.cuda.matmul.allow_tf32 = tf32
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
H, W, NC, G = 600, 1000, 21, 8
g = torch.Generator().manual_seed(3)
images = torch.rand(B, 3, H, W, generator=g).to(dev)
x0, y0 = torch.rand(B, G, generator=g) * (W - 320), torch.rand(B, G, generator=g) * (H - 320)
bw, bh = 50 + torch.rand(B, G, generator=g) * 250, 50 + torch.rand(B, G, generator=g) * 250
boxes = torch.stack((x0, y0, x0 + bw, y0 + bh), dim=2).to(dev)
labels = torch.randint(1, NC, (B, G), generator=g).to(dev)


def timed(fn, warm=3, iters=6):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / iters


if "--only-fast" in sys.argv:
    VARIANTS = (("afan_head_cache_device_rng", True, "device", "philox"),)
else:
    VARIANTS = (("afan_literal", False, "reference", "reference"), ("afan_head_cache", True, "reference", "reference"),
                ("afan_head_cache_device_rng", True, "device", "philox"))
res = {"config": f"Faster R-CNN R101-C4, {B}x3x{H}x{W}, se=3 (layer3 output {B}x1024x38x63), sd=roi, PGD-1, randinit, mix_sd, "
                 f"mix_layer 0101, {G} boxes / image, pre/post NMS 12000/2000", "conv_math": "tf32" if tf32 else "fp32"}
for key, hc, sampler, rng in VARIANTS:
    torch.manual_seed(3)
    m = PKG.faster_rcnn.FasterRCNN(NC, sampler=sampler).to(dev)
    for n_, p in m.named_parameters():                       # small prediction layers / residual branches: activations stay O(1)
        if ("_anchor_" in n_ or "_proposal_" in n_) and n_.endswith("weight"):
            p.data.mul_(0.01)
        if n_.endswith("bn3.weight"):
            p.data.fill_(0.25)
    if "--cl" in sys.argv:
        m = m.to(memory_format=torch.channels_last)
        images = images.contiguous(memory_format=torch.channels_last)
    tr = PKG.trainer_det.DetAfanTrainer(m, pertub_idx_se=3, gamma_se=0.5, gamma_sd=0.1, randinit=True, clip=False, mix_layer="0101",
                                        mix_sd=True, noise_sd=0.0, lr=1e-4, head_cache=hc, rng=rng, seed=3)
    res[key + "_ms"] = timed(lambda: tr.step(images, boxes, labels))
    res[key + "_loss"] = float(tr.step(images, boxes, labels)["loss"])
    res[key + "_peak_gb"] = round(torch.cuda.max_memory_allocated() / 2 ** 30, 2)
    del m, tr
    torch.cuda.empty_cache(); torch.cuda.reset_peak_memory_stats()
res["img_per_s"] = {k.replace("_ms", ""): round(1e3 * B / v, 2) for k, v in res.items() if k.endswith("_ms")}
print(json.dumps(res))
