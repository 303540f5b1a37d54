"""Drop-in for the reference training driver, Classification/main_perturb.py, on the B200 path.

Identical hot-path flags and defaults (main_perturb.py:28-49): --steps 5 --perturb_idx 13 --gamma 1.5
--eps 2 --randinit --clip, plus the base / optimiser flags (--batch_size --lr --momentum --weight_decay
--epochs --decreasing_lr --seed --gpu --print_freq --save_dir --resume).  Extras: --norm {linf,l2},
--rng {reference,philox}, --no_graph, --no_head_cache, --num_classes, --arch, --synthetic N (no dataset on
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
    p.add_argument("--rng", default="reference", choices=["philox", "reference"],
                   help="random start (--randinit): 'reference' = torch.rand on the CPU generator + H2D like attack_algo.py:44 "
                        "(default, same stream as the reference under the same seed); 'philox' = on-device generator")
    p.add_argument("--conv_math", default="tc3", choices=["fp32", "tf32", "3xtf32", "tc3", "cudnn"],
                   help="3x3 convolutions: tcgen05 3xTF32 implicit GEMM (tc3, default: fp32-grade accuracy on the tensor cores), "
                        "hand-written strict-fp32 FFMA kernels (fp32), their mma.sync TF32 / 3xTF32 twins, or the cuDNN library path")
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


class CifarLoaders:
    """The reference's splits (Classification/dataset.py:34-55): train = first 45000 training images (augmented, shuffled,
    drop_last), val = training images 45000..49999, test = the 10000 test images.  Data parallel: the TRAIN split is
    sharded with a DistributedSampler (every rank sees 1/world of each epoch, reshuffled per epoch); val / test are
    evaluated in full on every rank (identical weights -> identical numbers, no collective)."""

    def __init__(self, args, rank, world):
        import torchvision
        import torchvision.transforms as T
        from torch.utils.data import DataLoader, Subset
        from torch.utils.data.distributed import DistributedSampler
        aug = T.Compose([T.RandomCrop(32, padding=4), T.RandomHorizontalFlip(), T.ToTensor()])
        cifar = torchvision.datasets.CIFAR100 if args.num_classes == 100 else torchvision.datasets.CIFAR10
        train = Subset(cifar(args.data, train=True, transform=aug, download=False), list(range(45000)))
        val = Subset(cifar(args.data, train=True, transform=T.ToTensor(), download=False), list(range(45000, 50000)))
        test = cifar(args.data, train=False, transform=T.ToTensor(), download=False)
        self.sampler = DistributedSampler(train, num_replicas=world, rank=rank, shuffle=True, drop_last=True,
                                          seed=args.seed or 0) if world > 1 else None
        self.train = DataLoader(train, batch_size=args.batch_size, shuffle=self.sampler is None, sampler=self.sampler,
                                num_workers=2, pin_memory=True, drop_last=True)
        self.val = DataLoader(val, batch_size=args.batch_size, shuffle=False, num_workers=2, pin_memory=True)
        self.test = DataLoader(test, batch_size=args.batch_size, shuffle=False, num_workers=2, pin_memory=True)

    def set_epoch(self, epoch):
        if self.sampler is not None:
            self.sampler.set_epoch(epoch)


def to_device(loader, device):
    """Batches on the device, copied one step ahead on a copy stream (prefetch.DevicePrefetcher)."""
    from .prefetch import DevicePrefetcher
    return iter(DevicePrefetcher(loader, device))


def validate(trainer, loader, print_freq, tag, rank):
    """main_perturb.py:227-262: eval-mode top-1 / mean loss, sample-weighted; one host sync per print_freq batches."""
    loss_sum = torch.zeros((), device=trainer.device)
    correct = torch.zeros((), device=trainer.device)
    seen = 0
    n = None
    try:
        n = len(loader)
    except TypeError:
        pass
    for i, (x, y) in enumerate(loader):
        loss, out = trainer.evaluate(x, y)
        loss_sum += loss * x.shape[0]
        correct += (out.argmax(1) == y).sum()
        seen += x.shape[0]
        if i % print_freq == 0 and rank == 0:
            print(f"{tag}: [{i}/{n}]\tLoss {float(loss):.4f} ({float(loss_sum) / seen:.4f})\t"
                  f"Accuracy ({100.0 * float(correct) / seen:.3f})")
    acc = 100.0 * float(correct) / max(seen, 1)
    if rank == 0:
        print(f"valid_accuracy {acc:.3f}")
    return acc, float(loss_sum) / max(seen, 1)


def scheduler_state(milestones, gamma, base_lr, epoch_done, lr_now):
    """A state dict torch.optim.lr_scheduler.MultiStepLR.load_state_dict accepts (main_perturb.py:86,133)."""
    from collections import Counter
    return {"milestones": Counter(milestones), "gamma": gamma, "base_lrs": [base_lr], "last_epoch": epoch_done,
            "_step_count": epoch_done + 1, "_get_lr_called_within_step": False, "_last_lr": [lr_now]}


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
        setup_seed(args.seed)            # same seed on every rank: identical initial weights; data differs via the sampler
    conv.MODE = {"fp32": "afan", "tf32": "tf32", "3xtf32": "3xtf32", "tc3": "tc3", "cudnn": "cudnn"}[args.conv_math]
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = args.conv_math == "tf32"
    model = (resnet_s.resnet56(num_classes=args.num_classes) if args.arch == "resnet56"
             else resnet_s.resnet20(num_classes=args.num_classes)).to(device)
    trainer = AfanTrainer(model, perturb_idx=args.perturb_idx, steps=args.steps, gamma=args.gamma, eps=args.eps,
                          randinit=args.randinit, clip=args.clip, lr=args.lr, momentum=args.momentum,
                          weight_decay=args.weight_decay, norm=args.norm, rng="philox", seed=(args.seed or 0) + rank,
                          process_group=pg, sync_bn=not args.no_sync_bn, head_cache=not args.no_head_cache,
                          use_cuda_graph=not args.no_graph)
    milestones = list(map(int, args.decreasing_lr.split(",")))
    os.makedirs(args.save_dir, exist_ok=True)
    start_epoch, best_prec1 = 0, 0.0
    ckpt_path = os.path.join(args.save_dir, "checkpoint.pt")
    if args.resume:
        if rank == 0:
            print("resume from checkpoint")
        ck = torch.load(ckpt_path, map_location=device, weights_only=False)
        best_prec1, start_epoch = ck["best_prec1"], ck["epoch"]
        model.load_state_dict(ck["state_dict"])
        trainer.load_optimizer_state_dict(ck["optimizer"])       # momentum buffers: applied once the arena exists
    # the reference draws the random start on the CPU generator and copies it H2D (attack_algo.py:44): that is the
    # default here too (--rng reference); --rng philox is the on-device fast path (distributionally equal only)
    noise_shape = None
    if args.randinit and args.rng == "reference":
        with torch.no_grad():
            was = model.training
            model.eval()
            noise_shape = tuple(model(torch.zeros(1, 3, 32, 32, device=device), end_point=args.perturb_idx).shape[1:])
            model.train(was)
    loaders = None if args.synthetic else CifarLoaders(args, rank, world)
    all_norm, all_result = {"l2": {}, "linf": {}}, {"train": [], "ta": [], "test_ta": []}
    for epoch in range(start_epoch, args.epochs):
        base_lr = multistep_lr(epoch, args.lr, milestones)
        if rank == 0:
            print(base_lr)
        if loaders is None:
            loader = synthetic_loader(args.synthetic, args.batch_size, args.num_classes, 1000 * epoch + rank, device)
            n_batches = args.synthetic
        else:
            loaders.set_epoch(epoch)
            loader, n_batches = to_device(loaders.train, device), len(loaders.train)
        wp_steps = n_batches                                             # main_perturb.py:160 `len(train_loader)`
        t0, seen = time.time(), 0
        l2s, linfs = [], []
        loss_sum = torch.zeros((), device=device)
        correct = torch.zeros((), device=device)
        for i, (x, y) in enumerate(loader):
            trainer.set_lr(warmup_lr(i, wp_steps, args.lr) if epoch == 0 else base_lr)     # :167-168
            noise = torch.rand((x.shape[0],) + noise_shape).pin_memory().to(device, non_blocking=True) if noise_shape else None
            out = trainer.step(x, y, noise)
            l2s.append(out["l2"].clone()); linfs.append(out["linf"].clone())
            loss_sum += out["loss"] * x.shape[0]
            correct += (out["output_clean"].argmax(1) == y).sum()
            seen += x.shape[0]
            if i % args.print_freq == 0:                                                   # the only host sync
                trainer.check()                                                            # lost peer in the BN exchange -> raise
                if rank == 0:
                    print(f"Epoch: [{epoch}][{i}/{n_batches}]\tLoss {float(out['loss']):.4f} ({float(loss_sum) / seen:.4f})\t"
                          f"Accuracy {float(accuracy(out['output_clean'], y)):.3f} ({100.0 * float(correct) / seen:.3f})\t"
                          f"{seen * world / (time.time() - t0):.0f} img/s")
        train_acc = 100.0 * float(correct) / max(seen, 1)
        all_norm["l2"][epoch + 1] = float(torch.cat(l2s).mean()) if l2s else 0.0
        all_norm["linf"][epoch + 1] = float(torch.cat(linfs).mean()) if linfs else 0.0
        if rank == 0:
            print(f"train_accuracy {train_acc:.3f}\nl2 mean = {all_norm['l2'][epoch + 1]}\nlinf mean = {all_norm['linf'][epoch + 1]}")
        if loaders is None:                                               # synthetic data: a held-out synthetic batch set
            val = list(synthetic_loader(max(1, args.synthetic // 8), args.batch_size, args.num_classes, 7, device))
            test = list(synthetic_loader(max(1, args.synthetic // 8), args.batch_size, args.num_classes, 8, device))
        else:
            val, test = to_device(loaders.val, device), to_device(loaders.test, device)
        tacc, _ = validate(trainer, val, args.print_freq, "Test", rank)                   # :106
        test_tacc, _ = validate(trainer, test, args.print_freq, "Test", rank)             # :109
        all_result["train"].append(train_acc); all_result["ta"].append(tacc); all_result["test_ta"].append(test_tacc)
        is_best = tacc > best_prec1
        best_prec1 = max(tacc, best_prec1)
        if rank == 0:
            lr_next = multistep_lr(epoch + 1, args.lr, milestones)
            state = {"epoch": epoch + 1, "state_dict": model.state_dict(), "best_prec1": best_prec1,
                     "optimizer": trainer.optimizer_state_dict(lr=lr_next),
                     "scheduler": scheduler_state(milestones, 0.1, args.lr, epoch + 1, lr_next)}
            if is_best:
                torch.save(state, os.path.join(args.save_dir, "best_model.pt"))           # :122-129
            torch.save(state, ckpt_path)                                                  # :131-137
            pickle.dump(all_result, open(os.path.join(args.save_dir, "result.pkl"), "wb"))
            pickle.dump(all_norm, open(os.path.join(args.save_dir, "result_norm.pkl"), "wb"))
    trainer.close()
    return trainer


if __name__ == "__main__":
    main()
