import importlib, os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
PKG = importlib.import_module("cv_a-fan_b200")
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
if os.environ.get("AFAN_DET") == "1":
    torch.backends.cudnn.deterministic = True
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(4)
imgs = [torch.rand(8, 3, 32, 32, generator=g) for _ in range(3)]
tgts = [torch.randint(0, 10, (8,), generator=g) for _ in range(3)]
res = []
for graph in (False, True, False, True):
    torch.manual_seed(3)
    model = PKG.resnet_s.ResNet(num_blocks=(1, 1, 1)).to(dev)
    tr = PKG.trainer.AfanTrainer(model, perturb_idx=5, steps=2, gamma=1.0, eps=2.0, randinit=True, clip=True,
                                 rng="philox", seed=9, use_cuda_graph=graph)
    losses = [float(tr.step(i.to(dev), t.to(dev))["loss"]) for i, t in zip(imgs, tgts)]
    sd = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
    res.append((losses, sd))
    print("graph" if graph else "eager", losses)
for a, b, name in ((0, 2, "eager-eager"), (1, 3, "graph-graph"), (0, 1, "eager-graph")):
    worst = max(((res[a][1][k].float() - res[b][1][k].float()).abs().max().item(), k) for k in res[a][1])
    print(name, worst)
