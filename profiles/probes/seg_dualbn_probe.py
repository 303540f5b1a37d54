"""Diagnostic: DeepLab tail with DualBatchNorm2d vs nn.BatchNorm2d -- forward outputs, input gradients, parameter gradients."""
import copy, importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import seg_ref_step as ref
PKG = importlib.import_module("cv_a-fan_b200")
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
dev = torch.device("cuda:0")
m0 = PKG.deeplab.deeplabv3plus_resnet50(num_classes=ref.NUM_CLASSES, output_stride=16)
for m in m0.modules():
    if isinstance(m, torch.nn.Dropout): m.p = 0.0
ref.procedural_init(m0, seed=7)
m0.to(dev).train()
m1 = copy.deepcopy(m0)
from torchvision.models.segmentation.deeplabv3 import ASPPPooling
PKG.dual_bn.convert_batchnorm(m1.backbone.layer4); PKG.dual_bn.convert_batchnorm(m1.classifier, pooled=(ASPPPooling,))
images, labels = ref.make_batches(seed=21)
x = images[0].to(dev); y = labels[0].to(dev)
crit = torch.nn.CrossEntropyLoss(ignore_index=255)
res = []
for m in (m0, m1):
    head = m({"x": x, "adv": None, "out_idx": 3, "flag": "head"})
    feat = head["out"].detach().clone().requires_grad_(True)
    out = m({"x": x, "adv": feat, "out_idx": 3, "flag": "tail", "low_level_feat": head["low_level"].detach()})
    loss = crit(out, y)
    m.zero_grad()
    loss.backward()
    grads = {k: p.grad.detach().clone() for k, p in m.named_parameters() if p.grad is not None}
    res.append((out.detach(), float(loss), feat.grad.clone(), grads))
(o0, l0, g0, p0), (o1, l1, g1, p1) = res
rel = lambda a, b: float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))
print("loss", l0, l1, "out rel", rel(o1, o0), "dfeat rel", rel(g1, g0))
worst = sorted(((rel(p1[k], p0[k]), k) for k in p0), reverse=True)[:8]
for w in worst: print("param grad rel", w)
sd0, sd1 = m0.state_dict(), m1.state_dict()
worst = sorted(((rel(sd1[k].float(), sd0[k].float()), k) for k in sd0 if "running" in k or "tracked" in k), reverse=True)[:6]
for w in worst: print("buffer rel", w)
# per-BN check on the real activations: hook inputs of each tail BN in m0, run both implementations on them
import torch.nn as nn
acts = {}
hooks = [mod.register_forward_hook(lambda mod, inp, out, name=name: acts.__setitem__(name, inp[0].detach().clone()))
         for name, mod in m0.named_modules() if isinstance(mod, nn.BatchNorm2d) and (name.startswith("backbone.layer4") or name.startswith("classifier"))]
head = m0({"x": x, "adv": None, "out_idx": 3, "flag": "head"})
m0({"x": x, "adv": head["out"].detach(), "out_idx": 3, "flag": "tail", "low_level_feat": head["low_level"].detach()})
for h in hooks: h.remove()
mods0, mods1 = dict(m0.named_modules()), dict(m1.named_modules())
bad = []
for name, a in acts.items():
    b0, b1 = copy.deepcopy(mods0[name]), copy.deepcopy(mods1[name])
    b1.running_mean.copy_(b0.running_mean); b1.running_var.copy_(b0.running_var)
    a0 = a.clone().requires_grad_(True); a1 = a.clone().requires_grad_(True)
    y0 = b0(a0); y1 = b1(a1)
    dy = torch.randn_like(y0)
    (d0,) = torch.autograd.grad(y0, a0, dy); (d1,) = torch.autograd.grad(y1, a1, dy)
    rv = rel(b1.running_var, b0.running_var) if hasattr(b1, "running_var") else 0.0
    bad.append((max(rel(y1, y0), rel(d1, d0), rv), name, tuple(a.shape), rel(y1, y0), rel(d1, d0), rv))
for b in sorted(bad, reverse=True)[:8]: print("bn", b)
