import json, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from oracle import afan_ref_torch as ref_t
from oracle.full_case import full_case_inputs
from tests.util import GOLDEN, PKG, dev
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
z = np.load(os.path.join(GOLDEN, "cls_train_full.npz"))
r = json.loads(str(z["recipe"]))
for mode in ("afan", "tc3"):
    PKG.conv.MODE = mode
    torch.manual_seed(r["weight_seed"])
    model = PKG.resnet_s.ResNet(num_blocks=tuple(r["num_blocks"]), num_classes=r["num_classes"])
    port = ref_t.CifarResNetRef(tuple(r["num_blocks"]), r["num_classes"]); port.load_state_dict(model.state_dict()); port.train()
    opt, crit = ref_t.make_sgd(port), torch.nn.CrossEntropyLoss()
    model.to(dev())
    images, targets, noises = full_case_inputs(r)
    kw = dict(steps=r["steps"], gamma=r["gamma"], eps=r["eps"], perturb_idx=r["perturb_idx"], randinit=True, clip=True)
    tr = PKG.trainer.AfanTrainer(model, lr=0.1, use_cuda_graph=False, **kw)
    for i in range(r["iters"]):
        out = tr.step(images[i].to(dev()), targets[i].to(dev()), noises[i].to(dev()))
        loss, l2, linf = float(out["loss"]), out["l2"].cpu().numpy(), out["linf"].cpu().numpy()
        loss_p, _, l2_p, linf_p, _ = ref_t.afan_train_iteration(port, opt, crit, images[i], targets[i], noise=noises[i], **kw)
        ce = z["ce_values"][i][-2:]
        print(mode, i, "loss", loss, float(loss_p), (ce[0]+ce[1])/2, "linf rel", np.abs(linf/linf_p.numpy()-1).max(), "l2 rel", np.abs(l2/l2_p.numpy()-1).max(), "l2 mean rel", abs(l2.mean()/float(l2_p.mean())-1))
    sd = {k: v.detach().cpu().numpy() for k, v in model.state_dict().items()}
    worst = max(((np.abs(sd[k[10:]].reshape(-1)[::r["sub"]]-z[k]).max(), k) for k in z.files if k.startswith("final_sub/")))
    worst2 = max(((np.abs(sd[k[6:]].astype(np.float64)-z[k]).max(), k) for k in z.files if k.startswith("final/")))
    print(mode, "worst weight err", worst, worst2)
    tr.close()
