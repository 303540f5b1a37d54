"""A-FAN training step (Classification/main_perturb.py:173-201), re-designed for B200.

Reference iteration (per batch):  head fwd -> PGD (steps x [tail fwd, dgrad, 13 elementwise launches,
4 host syncs]) -> D2H of the whole perturbation for its norms -> adv tail fwd -> FULL clean fwd (head
again) -> backward -> SGD.  Here:

  * head cache      the head runs ONCE per batch; its output is reused (detached) as the PGD anchor and
                    (with its graph) as the clean half of the final pass.  The head's BatchNorm running
                    statistics are advanced twice (`replay=2`) because the reference forwards the head
                    twice in train() mode.
  * dual-BN tail    the final adversarial and clean tail passes run as ONE pass over [adv; clean] with
                    per-half batch statistics (groups=2) -- same maths as the reference's two passes.
  * fused PGD       one kernel per step (ascent + projection), per-sample ||delta|| norms fused into
                    the last step (no D2H), tail parameters frozen during the ascent (dgrad only) -- which
                    also lets every identity block run conv1 -> [bn1 + relu folded into conv2] -> conv2
                    (resnet_s.FUSE_BN1: statistics in conv1's tcgen05 epilogue, normalise-on-load).
  * flat arena      parameters / gradients / momentum live in three flat buffers: ONE fused SGD kernel,
                    ONE NCCL all-reduce of the gradient arena per iteration (weak-scaling data parallel).
  * CUDA graph      the whole iteration (forward, ascent loop, backward, all-reduce, SGD) is captured
                    once and replayed: ~3000 launches per iteration cost no Python/launch latency.
"""
import contextlib
from typing import Optional

import torch
import torch.nn as nn

from . import _lib, attack_algo, conv, ops, sync
from ._lib import AfanError
from .dual_bn import DualBatchNorm2d


class AfanTrainer:
    def __init__(self, model: nn.Module, *, perturb_idx: int = 13, steps: int = 5, gamma: float = 1.5,
                 eps: float = 2.0, randinit: bool = False, clip: bool = False, lr: float = 0.1,
                 momentum: float = 0.9, weight_decay: float = 5e-4, norm: str = "linf", rng: str = "philox",
                 seed: int = 0, criterion: Optional[nn.Module] = None, process_group=None, sync_bn: bool = True,
                 head_cache: bool = True, use_cuda_graph: bool = True, bn_exchange: str = "p2p"):
        """gamma / eps are in 1/255 units like the reference flags (main_perturb.py:180,183)."""
        self.model, self.k, self.L = model, int(perturb_idx), len(model.sequential_model)
        self.steps, self.gamma, self.eps = int(steps), gamma / 255.0, eps / 255.0
        self.randinit, self.clip, self.norm, self.rng, self.seed = randinit, clip, norm, rng, int(seed)
        self.momentum, self.weight_decay = momentum, weight_decay
        self.criterion = criterion if criterion is not None else nn.CrossEntropyLoss()
        self.pg = process_group
        self.world = torch.distributed.get_world_size(process_group) if process_group is not None else 1
        self.head_cache, self.use_graph = head_cache, use_cuda_graph
        self.device = next(model.parameters()).device
        if self.device.type != "cuda":
            raise AfanError("AfanTrainer needs the model on a CUDA device: there is no CPU path")
        self.mailbox = None
        self.bn_exchange_used = None                 # "p2p" | "nccl" | None (single process / per-replica statistics)
        if sync_bn and self.world > 1:
            self.bn_exchange_used = bn_exchange
            if bn_exchange == "p2p":                 # fused exchange over NVLink peer memory inside the BN kernels
                from .p2p import PeerMailbox
                cmax = max(m.num_features for m in model.modules() if isinstance(m, DualBatchNorm2d))
                ok = torch.ones(1, device=self.device)
                try:
                    self.mailbox = PeerMailbox(process_group, self.device, cmax=cmax)
                except AfanError as e:               # e.g. no peer access between the GPUs: all ranks must agree
                    ok.zero_()
                    err = e
                torch.distributed.all_reduce(ok, op=torch.distributed.ReduceOp.MIN, group=process_group)
                if not bool(ok.item()):
                    import warnings
                    warnings.warn("afan_b200: peer-mapped mailboxes unavailable on some rank; dual-BN statistics fall "
                                  "back to NCCL all-reduce (stats -> all-reduce -> finalize -> apply)")
                    self.mailbox = None
                    self.bn_exchange_used = "nccl"
            elif bn_exchange != "nccl":
                raise AfanError(f"bn_exchange must be 'p2p' or 'nccl', got {bn_exchange!r}")
            for m in model.modules():
                if isinstance(m, DualBatchNorm2d):
                    m.process_group = process_group
                    m.mailbox = self.mailbox
        self.lr_dev = torch.full((1,), float(lr), dtype=torch.float32, device=self.device)
        self._lr = float(lr)
        self.rng_offset = torch.zeros(1, dtype=torch.int64, device=self.device)
        self._arena_built = False
        self._pending_opt_state = None
        self._graph = None
        self._static = {}
        self._bn = [m for m in model.modules() if isinstance(m, DualBatchNorm2d)]
        self._bn_per_iter = None
        self._conv_pack = conv.PackPlan(model)     # one-launch weight repack for the hand-written 3x3 convolutions
        self.iterations = 0
        self.kernel_launches_per_iter = None       # afan kernels per iteration (counted at trace time)

    # ---- optimizer state in torch.optim.SGD's format (main_perturb.py:85,119-136 checkpoints) ----------------
    def optimizer_state_dict(self, lr: Optional[float] = None):
        """What `torch.optim.SGD(model.parameters(), ...).state_dict()` holds: momentum buffers keyed by the parameter's
        index in model.parameters() (parameters that never received a gradient, e.g. resnet_s.py:113 `w`, have no state,
        as under torch.optim.SGD) + one param group.  The reference can load_state_dict() it and vice versa."""
        params = list(self.model.parameters())
        state = {}
        if self._arena_built:
            index = {id(p): i for i, p in enumerate(params)}
            off = 0
            for p in self._params:
                k = p.numel()
                state[index[id(p)]] = {"momentum_buffer": self.flat_buf[off:off + k].view_as(p.data).clone()}
                off += k
        elif self._pending_opt_state is not None:
            state = self._pending_opt_state["state"]
        group = {"lr": self._lr if lr is None else float(lr), "momentum": self.momentum, "dampening": 0,
                 "weight_decay": self.weight_decay, "nesterov": False, "maximize": False, "foreach": None,
                 "differentiable": False, "fused": None, "params": list(range(len(params)))}
        return {"state": state, "param_groups": [group]}

    def load_optimizer_state_dict(self, sd):
        """Accepts the reference's `optimizer.state_dict()`.  The arena is built on the first step(), so the buffers are
        stashed until then (a resume must not silently reset momentum)."""
        if "param_groups" in sd and sd["param_groups"]:
            self.set_lr(float(sd["param_groups"][0].get("lr", self._lr)))
        if self._arena_built:
            self._apply_optimizer_state(sd)
        else:
            self._pending_opt_state = sd

    def _apply_optimizer_state(self, sd):
        params = list(self.model.parameters())
        index = {id(p): i for i, p in enumerate(params)}
        off = 0
        for p in self._params:
            k = p.numel()
            st = sd["state"].get(index[id(p)])
            if st is not None and st.get("momentum_buffer") is not None:
                self.flat_buf[off:off + k].copy_(st["momentum_buffer"].reshape(-1))
            off += k

    # ---- learning rate lives on the device so a captured graph follows the schedule --------------
    def set_lr(self, lr: float):
        if lr != self._lr:
            self.lr_dev.fill_(float(lr))
            self._lr = float(lr)

    # ---- flat parameter arena -------------------------------------------------------------------
    def _build_arena(self, used):
        n = sum(p.numel() for p in used)
        pad = (-n) % 4
        self.flat_param = torch.zeros(n + pad, dtype=torch.float32, device=self.device)
        self.flat_grad = torch.zeros_like(self.flat_param)
        self.flat_buf = torch.zeros_like(self.flat_param)
        off = 0
        for p in used:
            k = p.numel()
            self.flat_param[off:off + k].copy_(p.data.reshape(-1))
            p.data = self.flat_param[off:off + k].view_as(p.data)
            p.grad = self.flat_grad[off:off + k].view_as(p.data)
            off += k
        self._params = used
        self._arena_built = True
        if self._pending_opt_state is not None:        # --resume: momentum buffers loaded before the arena existed
            self._apply_optimizer_state(self._pending_opt_state)
            self._pending_opt_state = None
        # the arena is zeroed before every backward: gradient kernels may write straight into it.  Convolution weight
        # gradients ADD (safe for any number of uses); BatchNorm d(weight)/d(bias) are STORED, which needs the layer to
        # run exactly once per differentiated pass -- true for the head-cached [adv; clean] schedule of this trainer.
        for m in self._conv_pack.mods:
            m.grad_direct = True
        if self.head_cache and type(self) is AfanTrainer:
            for m in self._bn:
                m.grad_direct = True

    @contextlib.contextmanager
    def _params_frozen(self):
        """During the ascent only d(loss)/d(x_adv) is needed (attack_algo.py:52 `only_inputs=True`):
        freezing the parameters makes every tail layer skip its weight gradient."""
        ps = [p for p in self.model.parameters() if p.requires_grad]
        for p in ps:
            p.requires_grad_(False)
        try:
            yield
        finally:
            for p in ps:
                p.requires_grad_(True)

    # ---- one iteration (eager; also the body that gets captured) ----------------------------------
    def _iteration(self, images, target, noise, norms_out, ws):
        model, k, L, n = self.model, self.k, self.L, images.shape[0]
        if self.head_cache and k > 0:
            feat = model(images, end_point=k, start_point=0, replay=2)
        elif k > 0:
            with torch.no_grad():
                feat = model(images, end_point=k, start_point=0)                    # main_perturb.py:173
        else:
            feat = images
        anchor = feat.detach()
        extras = dict(norms_out=norms_out, workspace=ws, norm=self.norm)
        if self.randinit:
            if noise is not None:
                extras["noise"] = noise
            else:
                extras.update(rng=self.rng, seed=self.seed, offset_device=self.rng_offset)
        with self._params_frozen():
            x_adv = attack_algo.PGD(anchor, self.criterion, y=target, model=model, steps=self.steps,
                                    gamma=self.gamma, start_idx=k, layer_number=L, eps=self.eps,
                                    randinit=self.randinit, clip=self.clip, **extras)        # :176-185
        if self.randinit and noise is None and self.rng == "philox":
            self.rng_offset += (anchor.numel() + 3) // 4
        if self.head_cache:
            both = torch.cat([x_adv.detach(), feat], dim=0)
            logits = model(both, end_point=L, start_point=k, groups=2)                  # :195 + :196 in one pass
            out_adv, out_clean = logits[:n], logits[n:]
        else:
            out_adv = model(x_adv.detach(), end_point=L, start_point=k)                 # :195
            out_clean = model(images, end_point=L, start_point=0)                       # :196
        loss = (self.criterion(out_adv, target) + self.criterion(out_clean, target)) / 2     # :197
        return loss, out_clean

    def _optimize(self, loss):
        self.flat_grad.zero_()                                                       # :199
        conv.wgrad_overlap_begin(self.device)      # weight gradients on a side stream, beside the dgrad / BN chain
        try:
            loss.backward()                                                          # :200
        finally:
            conv.wgrad_overlap_join()
        scale = sync.allreduce_grad_arena_(self.flat_grad, self.pg)                  # one NCCL message / iteration
        ops.sgd_momentum_(self.flat_param, self.flat_grad, self.flat_buf, self.lr_dev, momentum=self.momentum,
                          weight_decay=self.weight_decay, grad_scale=scale)                   # :201
        self._conv_pack.pack()                     # the SGD kernel wrote the weights behind autograd's back

    def _discover_arena(self, images, target, noise, norms_out, ws):
        """First iteration: find the parameters that actually receive gradients (torch.optim.SGD skips
        parameters whose .grad is None, e.g. the unused `w` of resnet_s.py:113) and build the arena from
        them, without touching any state (BN statistics are restored)."""
        snap = self._snapshot()
        for p in self.model.parameters():
            p.grad = None
        self._conv_pack.pack()
        loss, _ = self._iteration(images, target, noise, norms_out, ws)
        loss.backward()
        used = [p for p in self.model.parameters() if p.grad is not None]
        self._restore(snap)
        self._build_arena(used)
        self._conv_pack.pack()

    # ---- state snapshot (graph warm-up must not leak into training state) -------------------------
    def _snapshot(self):
        s = {"bn": [(m.running_mean.clone(), m.running_var.clone(), m.num_batches_tracked.clone(), m._pending_batches)
                    for m in self._bn], "rng": self.rng_offset.clone()}
        if self._arena_built:
            s["param"], s["buf"] = self.flat_param.clone(), self.flat_buf.clone()
        return s

    def _restore(self, s):
        for m, (rm, rv, nbt, pend) in zip(self._bn, s["bn"]):
            m.running_mean.copy_(rm); m.running_var.copy_(rv); m.num_batches_tracked.copy_(nbt)
            m._pending_batches = pend
        self.rng_offset.copy_(s["rng"])
        if "param" in s:
            self.flat_param.copy_(s["param"]); self.flat_buf.copy_(s["buf"])

    # ---- public step --------------------------------------------------------------------------------
    def step(self, images: torch.Tensor, target: torch.Tensor, noise: Optional[torch.Tensor] = None):
        """One A-FAN training iteration on device tensors.  Returns a dict of DEVICE tensors
        {loss, output_clean, l2, linf}: nothing is synchronised or copied to the host."""
        self.model.train()
        n = images.shape[0]
        if not self._static:
            self._static = {"norms": torch.zeros(2, n, dtype=torch.float32, device=self.device),
                            "ws": ops.norms_workspace(n, self.device)}
        st = self._static
        if not self._arena_built:
            self._discover_arena(images, target, noise, st["norms"], st["ws"])
        if not self.use_graph:
            l0 = _lib.launch_count
            self._conv_pack.pack()                 # weights may have been written from outside (load_state_dict)
            loss, out_clean = self._iteration(images, target, noise, st["norms"], st["ws"])
            self._optimize(loss)
            self.kernel_launches_per_iter = _lib.launch_count - l0
            self.iterations += 1
            return {"loss": loss.detach(), "output_clean": out_clean.detach(), "l2": st["norms"][0], "linf": st["norms"][1]}
        if self._graph is None:
            self._capture(images, target, noise)
        st["images"].copy_(images, non_blocking=True)
        st["target"].copy_(target, non_blocking=True)
        if noise is not None:
            st["noise"].copy_(noise, non_blocking=True)
        self._graph.replay()
        for m, inc in zip(self._bn, self._bn_per_iter):
            m._pending_batches += inc
        self.iterations += 1
        return {"loss": st["loss"], "output_clean": st["out_clean"], "l2": st["norms"][0], "linf": st["norms"][1]}

    def _capture(self, images, target, noise):
        st = self._static
        st["images"], st["target"] = images.clone(), target.clone()
        st["noise"] = noise.clone() if noise is not None else None
        snap = self._snapshot()
        # warm-up and capture run on ONE private stream so that autograd's per-parameter AccumulateGrad
        # nodes live on the stream that is later captured
        side = self._stream = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):                      # warm-up: cuDNN autotune, NCCL communicator, lazy module load
                self._conv_pack.pack()
                loss, _ = self._iteration(st["images"], st["target"], st["noise"], st["norms"], st["ws"])
                self._optimize(loss)
            del loss
        torch.cuda.current_stream().wait_stream(side)
        self._restore(snap)
        pend0 = [m._pending_batches for m in self._bn]
        self._graph = torch.cuda.CUDAGraph()
        l0 = _lib.launch_count
        with torch.cuda.graph(self._graph, stream=side):
            self._conv_pack.pack()
            loss, out_clean = self._iteration(st["images"], st["target"], st["noise"], st["norms"], st["ws"])
            self._optimize(loss)
            st["loss"], st["out_clean"] = loss.detach(), out_clean.detach()
        del loss, out_clean
        self.kernel_launches_per_iter = _lib.launch_count - l0      # afan kernels replayed per iteration
        self._bn_per_iter = [m._pending_batches - p0 for m, p0 in zip(self._bn, pend0)]
        for m, p0 in zip(self._bn, pend0):
            m._pending_batches = p0                 # capture records launches, it does not run them

    def check(self):
        """Raise if a fused BN statistics exchange timed out on this rank (synchronises the device).  The kernels already
        poison the statistics with NaN on a timeout; this turns the condition into an exception.  Call it wherever the
        host synchronises anyway (main_perturb does at every print_freq; bench.py after each timed region)."""
        if self.mailbox is not None:
            self.mailbox.check()

    def close(self):
        """Drop the captured graph (it pins NCCL communicator resources: destroy_process_group() blocks
        while a graph holding captured collectives is alive)."""
        torch.cuda.synchronize(self.device)
        self._graph = None
        self._static = {}
        # hand the model back in stand-alone mode: modules repack their own weights again and gradients go through autograd
        self._conv_pack.release()
        for m in list(self._conv_pack.mods) + list(self._bn):
            m.grad_direct = False
        if self.mailbox is not None:
            self.mailbox.check()

    # ---- evaluation (main_perturb.py:227-262) ---------------------------------------------------------
    @torch.no_grad()
    def evaluate(self, images, target):
        self.model.eval()
        self._conv_pack.pack()
        out = self.model(images, end_point=self.L, start_point=0)
        return self.criterion(out, target), out
