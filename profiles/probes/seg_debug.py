import importlib, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tests import test_gpu_seg as T
torch.backends.cudnn.allow_tf32 = False; torch.backends.cuda.matmul.allow_tf32 = False
for name in ("A", "B"):
    for hc in (False, True):
        losses, sd = T._run(name, hc)
        print(name, hc, "loss rel err", np.abs(losses / T.G[f"{name}/losses"] - 1).max())
        keys = [str(k) for k in T.G["keys"]]
        errs = []
        for k, want in zip(keys, T.G[f"{name}/norms"]):
            if k.endswith("num_batches_tracked"): continue
            got = float(sd[k].double().norm())
            errs.append((abs(got - want) / max(want, 1e-3), k))
        errs.sort(reverse=True)
        print("   worst:", [(round(e, 5), k) for e, k in errs[:8]])
        rs = [e for e, k in errs if k.endswith(("running_mean", "running_var"))]
        ps = [e for e, k in errs if not k.endswith(("running_mean", "running_var"))]
        print("   max stat err", max(rs), "max param err", max(ps))
np.set_printoptions(precision=6, suppress=True, linewidth=200)
for name in ("A", "B"):
    losses, sd = T._run(name, False)
    print(name, "got\n", losses, "\nwant\n", T.G[f"{name}/losses"], "\nrel\n", losses / T.G[f"{name}/losses"] - 1)
