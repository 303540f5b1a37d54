import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import test_gpu_seg as T
from oracle import seg_ref_step as ref
PKG = T.PKG
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
np.set_printoptions(precision=7, suppress=True, linewidth=200)
dev = torch.device("cuda:0")
def run_restatement(name):
    c = ref.CASES[name]
    model = PKG.deeplab.deeplabv3plus_resnet50(num_classes=ref.NUM_CLASSES, output_stride=16)
    for m in model.modules():
        if isinstance(m, torch.nn.Dropout): m.p = 0.0
    ref.procedural_init(model, seed=7); model.to(dev).train()
    images, labels = ref.make_batches(seed=21)
    opt = torch.optim.SGD(params=[{"params": model.backbone.parameters(), "lr": 0.1 * ref.LR}, {"params": model.classifier.parameters(), "lr": ref.LR}], lr=ref.LR, momentum=0.9, weight_decay=ref.WD)
    crit = torch.nn.CrossEntropyLoss(ignore_index=255, reduction="mean")
    out = []
    for it in range(ref.ITERS):
        draws = []
        if c["randinit"]: draws += [torch.from_numpy(T.G[f"{name}/noise_se{it}"]).to(dev), torch.from_numpy(T.G[f"{name}/noise_sd{it}"]).to(dev)]
        if c["noise_sd"] != 0: draws.append(torch.from_numpy(T.G[f"{name}/noise_n{it}"]).to(dev))
        q = iter(draws); rand = lambda shape: next(q)
        out.append(ref.reference_iteration(model, ref.TorchAttackAlgo(rand), images[it].to(dev), labels[it].to(dev), c, crit, opt, rand=rand))
    return np.array(out), {k: v.detach().float().cpu() for k, v in model.state_dict().items()}
for name in ("A", "B"):
    lr_, sdr = run_restatement(name)
    for hc in (False, True):
        lt, sdt = T._run(name, hc)
        print(name, hc, "restatement-on-GPU vs trainer: loss rel", np.abs(lt / lr_ - 1).max())
        errs = sorted(((float((sdt[k] - sdr[k]).abs().max()), k) for k in sdr if not k.endswith("tracked")), reverse=True)
        print("   worst abs:", errs[:4])
    print(name, "restatement-on-GPU vs CPU golden: rel\n", lr_ / T.G[f"{name}/losses"] - 1)
