"""Config 5 (BASELINE.json configs[4]): DeepLabv3+ ResNet-101, Cityscapes-shaped synthetic 513x513 crops, batch 4,
feature PGD on the ASPP input (backbone stage 4) + decoder PGD on the ASPP output, one A-FAN iteration of
Segmentation/main_aug_final.py:160-232.  Times (CUDA events, eager launches):
  reference_on_gpu   the iteration restated in plain PyTorch (oracle/seg_ref_step.py: un-fused ATen PGD, 3 head sweeps, torch SGD)
  afan               SegAfanTrainer, head_cache off / on
Usage: python profiles/bench_seg.py [--tf32]"""
import importlib, json, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import seg_ref_step as ref
PKG = importlib.import_module("cv_a-fan_b200")
tf32 = "--tf32" in sys.argv
torch.backends.cudnn.allow_tf32 = tf32
torch.backends.cuda.matmul.allow_tf32 = tf32
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
B, S, NC = 4, 513, 19
c = dict(se=4, sd="aspp", steps=1, eps=2.0, gamma_se=0.5, gamma_sd=0.5, randinit=True, clip=False, mix_sd=True, noise_sd=0.0, mix_layer="01")
g = torch.Generator().manual_seed(3)
images = torch.rand(B, 3, S, S, generator=g).to(dev)
labels = torch.randint(0, NC, (B, S, S), generator=g).to(dev)


def model_():
    torch.manual_seed(3)
    m = PKG.deeplab.deeplabv3plus_resnet101(num_classes=NC, output_stride=16).to(dev)
    return m.train()


def timed(fn, warm=3, iters=8):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) / iters


res = {"config": "DeepLabv3+ R101 os16, 4x3x513x513, se=4 (ASPP input 4x2048x33x33), sd=aspp, PGD-1, randinit, mix_sd, mix_layer 01",
       "conv_math": "tf32" if tf32 else "fp32"}
m = model_()
opt = torch.optim.SGD(params=[{"params": m.backbone.parameters(), "lr": 0.001}, {"params": m.classifier.parameters(), "lr": 0.01}],
                      lr=0.01, momentum=0.9, weight_decay=1e-4)
crit = torch.nn.CrossEntropyLoss(ignore_index=255, reduction="mean")
algo = ref.TorchAttackAlgo(lambda shape: torch.rand(shape))          # CPU torch.rand + H2D like the reference
res["reference_on_gpu_ms"] = timed(lambda: ref.reference_iteration(m, algo, images, labels, c, crit, opt,
                                                                   rand=lambda shape: torch.rand(shape).to(dev)))
del m, opt
for hc, dbn in ((False, False), (True, False), (True, True)):
    m = model_()
    tr = PKG.trainer_seg.SegAfanTrainer(m, pertub_idx_se=c["se"], pertub_idx_sd=c["sd"], steps=c["steps"], eps=c["eps"],
                                        gamma_se=c["gamma_se"], gamma_sd=c["gamma_sd"], randinit=True, clip=False, mix_sd=True,
                                        noise_sd=0.0, mix_layer="01", head_cache=hc, dual_bn=dbn)
    res[f"afan_head_cache_{int(hc)}_dual_bn_{int(dbn)}_ms"] = timed(lambda: tr.step(images, labels))
    del m, tr
m = model_()
tr = PKG.trainer_seg.SegAfanTrainer(m, pertub_idx_se=c["se"], pertub_idx_sd=c["sd"], steps=c["steps"], eps=c["eps"], gamma_se=c["gamma_se"],
                                    gamma_sd=c["gamma_sd"], randinit=True, clip=False, mix_sd=True, noise_sd=0.0, mix_layer="01",
                                    head_cache=True, dual_bn=True, use_cuda_graph=True)
res["afan_head_cache_1_dual_bn_1_graph_ms"] = timed(lambda: tr.step(images, labels))
del m, tr
res["img_per_s"] = {k.replace("_ms", ""): round(1e3 * B / v, 2) for k, v in res.items() if k.endswith("_ms")}
print(json.dumps(res))
