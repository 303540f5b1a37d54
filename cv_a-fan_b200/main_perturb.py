"""Drop-in for the reference training driver, Classification/main_perturb.py, on the B200 path.

Identical hot-path flags and defaults (main_perturb.py:28-49): --steps 5 --perturb_idx 13 --gamma 1.5
--eps 2 --randinit --clip, plus the base / optimiser flags (--batch_size --lr --momentum --weight_decay
--epochs --decreasing_lr --seed --gpu --print_freq --save_dir --resume).  Extras: --norm {linf,l2},
--rng {philox,reference}, --no_graph, --no_head_cache, --num_classes, --arch, --synthetic N (no dataset on
this box: N synthetic CIFAR-shaped batches per epoch), --bench.

    python -m torch.distributed.run --nproc-per-node 8 -m afan_b200.main_perturb --synthetic 50 ...   # data parallel
"""
import argparse
import os
import pickle
import random
import time

import numpy as np
import torch
import torch.nn as nn

from . import conv, resnet_s
from .trainer import AfanTrainer


def build_parser():
    p = argparse.ArgumentParser(description="A-FAN CIFAR training on B200 (afan_b200)")
    # base setting (main_perturb.py:28-33)
    p.add_argument("--data", type=str, default="../data", help="location of the data corpus")
    p.add_argument("--print_freq", default=50, type=int)
    p.add_argument("--seed", default=None, type=int)
    p.add_argument("--gpu", type=int, default=0)
    p.add_argument("--resume", action="store_true")
    p.add_argument("--save_dir", default="res56s_adv_aug", type=str)
    # optimizer setting (:36-41)
    p.add_argument("--batch_size", type=int, default=128)
    p.add_argument("--lr", default=0.1, type=float)
    p.add_argument("--momentum", default=0.9, type=float)
    p.add_argument("--weight_decay", default=5e-4, type=float)
    p.add_argument("--epochs", default=200, type=int)
    p.add_argument("--decreasing_lr", default="50,150")
    # A-FAN setting (:44-49)
    p.add_argument("--steps", default=5, type=int, help="PGD-steps")
    p.add_argument("--perturb_idx", default=13, type=int, help="index of perturb layers")
    p.add_argument("--gamma", default=1.5, type=float, help="PGD step size (in 1/255)")
    p.add_argument("--eps", default=2, type=float, help="ball radius (in 1/255)")
    p.add_argument("--randinit", action="store_true", help="whether using randinit")
    p.add_argument("--clip", action="store_true", help="whether using clip")
    # extras of this implementation
    p.add_argument("--norm", default="linf", choices=["linf", "l2"])
    p.add_argument("--rng", default="philox", choices=["philox", "reference"])
    p.add_argument("--conv_math", default="fp32", choices=["fp32", "tf32", "3xtf32", "cudnn"],
                   help="3x3 convolutions: hand-written strict-fp32 kernels (default), their TF32 / 3xTF32 tensor-core twins, "
                        "or the cuDNN library path")
    p.add_argument("--no_graph", action="store_true")
    p.add_argument("--no_head_cache", action="store_true")
    p.add_argument("--no_sync_bn", action="store_true")
    p.add_argument("--arch", default="resnet56", choices=["resnet56", "resnet20"])
    p.add_argument("--num_classes", default=10, type=int)
    p.add_argument("--synthetic", default=0, type=int, help="use N synthetic batches per epoch instead of CIFAR")
    return p


def setup_seed(seed):
    torch.manual_seed(seed)
    torch.cuda.manual_seed_all(seed)
    np.random.seed(seed)
    random.seed(seed)
    torch.backends.cudnn.deterministic = True


def warmup_lr(step, warm_up_steps, max_lr):
    """main_perturb.py:288-293"""
    return min(step * max_lr / max(warm_up_steps - 1, 1), max_lr)


def multistep_lr(epoch, base_lr, milestones, gamma=0.1):
    return base_lr * gamma ** sum(epoch >= m for m in milestones)


def accuracy(output, target):
    return (output.argmax(1) == target).float().mean() * 100.0


def synthetic_loader(n_batches, batch, num_classes, seed, device):
    g = torch.Generator().manual_seed(seed)
    for _ in range(n_batches):
        yield (torch.rand(batch, 3, 32, 32, generator=g).to(device, non_blocking=True),
               torch.randint(0, num_classes, (batch,), generator=g).to(device, non_blocking=True))


def cifar_loader(args, train, device):
    import torchvision
    import torchvision.transforms as T
    tf = T.Compose([T.RandomCrop(32, padding=4), T.RandomHorizontalFlip(), T.ToTensor()]) if train else T.ToTensor()
    ds = torchvision.datasets.CIFAR10(args.data, train=train, transform=tf, download=False)
    dl = torch.utils.data.DataLoader(ds, batch_size=args.batch_size, shuffle=train, num_workers=2, pin_memory=True,
                                     drop_last=train)
    for x, y in dl:
        yield x.to(device, non_blocking=True), y.to(device, non_blocking=True)


def main(argv=None):
    args = build_parser().parse_args(argv)
    print(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(args.gpu)))
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    pg = None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=device)
        pg = torch.distributed.group.WORLD
    if args.seed:
        setup_seed(args.seed)
    conv.MODE = {"fp32": "afan", "tf32": "tf32", "3xtf32": "3xtf32", "cudnn": "cudnn"}[args.conv_math]
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = args.conv_math == "tf32"
    model = (resnet_s.resnet56(num_classes=args.num_classes) if args.arch == "resnet56"
             else resnet_s.resnet20(num_classes=args.num_classes)).to(device)
    trainer = AfanTrainer(model, perturb_idx=args.perturb_idx, steps=args.steps, gamma=args.gamma, eps=args.eps,
                          randinit=args.randinit, clip=args.clip, lr=args.lr, momentum=args.momentum,
                          weight_decay=args.weight_decay, norm=args.norm, rng=args.rng, seed=(args.seed or 0) + rank,
                          process_group=pg, sync_bn=not args.no_sync_bn, head_cache=not args.no_head_cache,
                          use_cuda_graph=not args.no_graph)
    milestones = list(map(int, args.decreasing_lr.split(",")))
    os.makedirs(args.save_dir, exist_ok=True)
    start_epoch, best_prec1 = 0, 0.0
    ckpt_path = os.path.join(args.save_dir, "checkpoint.pt")
    if args.resume and os.path.exists(ckpt_path):
        ck = torch.load(ckpt_path, map_location=device)
        model.load_state_dict(ck["state_dict"])
        start_epoch, best_prec1 = ck["epoch"], ck["best_prec1"]
        if "momentum_buffer" in ck and trainer._arena_built:
            trainer.flat_buf.copy_(ck["momentum_buffer"])
    all_norm = {"l2": {}, "linf": {}}
    for epoch in range(start_epoch, args.epochs):
        base_lr = multistep_lr(epoch, args.lr, milestones)
        n_batches = args.synthetic if args.synthetic else None
        loader = (synthetic_loader(args.synthetic, args.batch_size, args.num_classes, 1000 * epoch + rank, device)
                  if args.synthetic else cifar_loader(args, True, device))
        wp_steps = args.synthetic if args.synthetic else 50000 // (args.batch_size * world)
        t0, seen = time.time(), 0
        l2s, linfs, loss_sum, acc_sum, count = [], [], 0.0, 0.0, 0
        for i, (x, y) in enumerate(loader):
            trainer.set_lr(warmup_lr(i, wp_steps, args.lr) if epoch == 0 else base_lr)     # :167-168
            out = trainer.step(x, y)
            l2s.append(out["l2"].clone()); linfs.append(out["linf"].clone())
            seen += x.shape[0] * world
            if i % args.print_freq == 0:                                                   # the only host sync
                loss, prec = float(out["loss"]), float(accuracy(out["output_clean"], y))
                loss_sum += loss; acc_sum += prec; count += 1
                if rank == 0:
                    print(f"Epoch: [{epoch}][{i}/{n_batches}]\tLoss {loss:.4f}\tAccuracy {prec:.3f}\t"
                          f"{seen / (time.time() - t0):.0f} img/s")
        all_norm["l2"][epoch + 1] = float(torch.cat(l2s).mean()) if l2s else 0.0
        all_norm["linf"][epoch + 1] = float(torch.cat(linfs).mean()) if linfs else 0.0
        if rank == 0:
            print(f"l2 mean = {all_norm['l2'][epoch + 1]}\nlinf mean = {all_norm['linf'][epoch + 1]}")
            state = {"epoch": epoch + 1, "state_dict": model.state_dict(), "best_prec1": best_prec1}
            if trainer._arena_built:
                state["momentum_buffer"] = trainer.flat_buf.clone()
            torch.save(state, ckpt_path)
            pickle.dump(all_norm, open(os.path.join(args.save_dir, "result_norm.pkl"), "wb"))
    trainer.close()
    return trainer


if __name__ == "__main__":
    main()
