"""3x3 convolution module of the splittable ResNet, backed by the hand-written sm_100a kernels.

`Conv3x3` is an `nn.Conv2d` (same parameter name / shape / init, so reference checkpoints load,
Classification/resnet_s.py:53,55) whose forward + backward run `afan_conv3x3_f32` (forward and, with the other weight
packing, the input gradient) and `afan_conv3x3_wgrad_f32` for the shapes those kernels cover -- stride 1, C_in == C_out
in {16, 32, 64}, square 8/16/32 maps, i.e. every BasicBlock convolution of the CIFAR ResNets except the two
stride-2 stage transitions and the stem.  Anything else is the library convolution (cuDNN), which the north star leaves
outside the hand-written scope.

Weight packing.  The kernels read W as [reduction channel][tap][output channel].  A module repacks its own weight when
`weight._version` moved (optimizer.step(), load_state_dict, ...).  A trainer that updates the weights behind autograd's
back (the arena SGD kernel) calls `pack_all(model)` once per iteration instead: ONE launch for all layers.
"""
import os

import torch
import torch.nn as nn

from . import ops

# Module-level switch (env AFAN_CONV), read at call time:
#   "afan"   hand-written kernels, strict fp32 FFMA accumulation
#   "tf32"   hand-written tensor-core kernels, one TF32 pass (PyTorch's default conv math on Ampere+; opt-in here)
#   "3xtf32" hand-written tensor-core kernels, hi/lo split (fp32-level accuracy; measured no faster than "afan")
#   "tc3"    (default) tcgen05 implicit GEMM (Blackwell tensor cores, TMEM accumulators, bulk-copy fed), 3xTF32 split:
#            fp32-grade accuracy, deterministic; covers the tail shapes (C, H) in {(32, 16), (64, 8)} that every PGD step
#            re-executes; the C = 16 layers, the stride-2 transitions and every weight gradient stay on the FFMA kernels
#   "cudnn"  the library convolution (benchmark comparisons)
MODE = os.environ.get("AFAN_CONV", "tc3")
STRIDE2 = os.environ.get("AFAN_S2", "1") != "0"      # hand-written stride-2 transitions (0: library convolution, for A/B timing)
_MATH = {"afan": "fp32", "tf32": "tf32", "3xtf32": "3xtf32", "tc3": "umma"}      # MODE -> weight packing


def _call_math(c: int, x) -> str:
    """Kernel family for one call (None: library convolution).  In "tc3" mode the packing is per layer (by C), the kernel
    per call: a C in {32, 64} layer on a map the tcgen05 kernel does not cover falls back to the library."""
    if MODE == "tc3":
        if c in (32, 64):
            return "umma" if ops.conv3x3_umma_supported(x.shape[0], c, x.shape[2]) else None
        return "fp32"
    return _MATH.get(MODE)


# ---- weight gradients on a second stream --------------------------------------------------------------------------
# The input-gradient chain of a backward pass (dgrad conv -> BatchNorm backward -> dgrad conv ...) is sequential and made of
# launch-latency-bound kernels, while the weight gradients feed nothing but the gradient arena.  A trainer that owns that
# arena brackets loss.backward() with wgrad_overlap_begin / wgrad_overlap_join: every weight-gradient launch then goes to a
# side stream (forked from the backward stream at the point its dy exists) and fills otherwise idle SMs; the join orders
# the arena before the all-reduce / SGD.  Works under CUDA-graph capture (fork / join become graph edges).
WGRAD_OVERLAP = os.environ.get("AFAN_WGRAD_OVERLAP", "0") == "1"      # opt-in: measured 10.58 vs 10.63 ms/step, i.e. nothing
_overlap = {"stream": None, "keep": [], "active": False}


def wgrad_overlap_begin(device):
    if not WGRAD_OVERLAP:
        return
    if _overlap["stream"] is None or _overlap["stream"].device != device:
        _overlap["stream"] = torch.cuda.Stream(device=device)
    _overlap["active"] = True


def wgrad_overlap_join():
    if _overlap["active"]:
        torch.cuda.current_stream().wait_stream(_overlap["stream"])
        _overlap["keep"].clear()              # x / dy stayed alive until the side stream's work was ordered before ours
        _overlap["active"] = False


WGRAD_UMMA = os.environ.get("AFAN_WGRAD_UMMA", "1") == "1"      # tcgen05 weight gradient in "tc3" mode (tail shapes)


def _wgrad_math(mod, x) -> str:
    # measured on B200 (L2-cold, kernel + fold): 128x32x16x16 21.1 vs 22.5 us (FFMA), 256x32x16x16 29.2 vs 38.7 us,
    # 256x64x8x8 50.8 vs 36.8 us -> the tcgen05 weight gradient is used for the C = 32 layers only
    if MODE == "tc3" and WGRAD_UMMA and mod.stride == (1, 1) and mod.out_channels == 32 and \
            ops.conv3x3_wgrad_umma_supported(x.shape[0], mod.out_channels, x.shape[2]):
        return "umma"
    return "fp32"


def _wgrad_s1(x, dy, ws, accumulate_into=None, mod=None):
    return ops.conv3x3_wgrad(x, dy, ws, accumulate_into=accumulate_into, math=_wgrad_math(mod, x))


def _arena_wgrad(fn, x, dy, mod):
    """Weight gradient added straight into mod.weight.grad (a slice of the trainer's gradient arena)."""
    if fn is ops.conv3x3_wgrad:
        fn = lambda a, b, ws, accumulate_into=None: _wgrad_s1(a, b, ws, accumulate_into, mod)
    if not _overlap["active"]:
        fn(x, dy, mod.wgrad_workspace(), accumulate_into=mod.weight.grad)
        return
    side = _overlap["stream"]
    side.wait_stream(torch.cuda.current_stream())                 # dy (and x) are complete on the backward stream
    with torch.cuda.stream(side):
        fn(x, dy, mod.wgrad_workspace(), accumulate_into=mod.weight.grad)
    _overlap["keep"].append((x, dy))


class _Conv3x3Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, mod):
        x = x.contiguous()
        wf, wd = mod.packed()
        ctx.save_for_backward(x)
        ctx.mod, ctx.wd, ctx.math = mod, wd, _call_math(mod.out_channels, x)
        return ops.conv3x3(x, wf, math=ctx.math)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = ops.conv3x3(dy, ctx.wd, math=ctx.math) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            mod = ctx.mod
            if mod.grad_direct and mod.weight.grad is not None:
                _arena_wgrad(ops.conv3x3_wgrad, x, dy, mod)                                        # no temporary, no add launch
            else:
                dw = _wgrad_s1(x, dy, mod.wgrad_workspace(), mod=mod)
        return dx, dw, None


class _Conv3x3TapFn(torch.autograd.Function):
    """(conv(x), x) -- the second output is x itself, handed to the block's identity shortcut (resnet_s.py:75).  Both
    gradients meet in ONE backward call, so the shortcut's gradient is added in the dgrad kernel's epilogue instead of
    by an autograd accumulation launch."""

    @staticmethod
    def forward(ctx, x, weight, mod):
        x = x.contiguous()
        wf, wd = mod.packed()
        ctx.save_for_backward(x)
        ctx.mod, ctx.wd, ctx.math = mod, wd, _call_math(mod.out_channels, x)
        return ops.conv3x3(x, wf, math=ctx.math), x.view_as(x)

    @staticmethod
    def backward(ctx, dy, dtap):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = ops.conv3x3(dy, ctx.wd, math=ctx.math, addend=dtap.contiguous() if dtap is not None else None)
        dw = None
        if ctx.needs_input_grad[1]:
            mod = ctx.mod
            if mod.grad_direct and mod.weight.grad is not None:
                _arena_wgrad(ops.conv3x3_wgrad, x, dy, mod)
            else:
                dw = _wgrad_s1(x, dy, mod.wgrad_workspace(), mod=mod)
        return dx, dw, None


class _Conv3x3S2Fn(torch.autograd.Function):
    """Stride-2 stage transition (resnet_s.py:98): hand-written forward, input gradient and weight gradient (strict fp32
    in every MODE, deterministic)."""

    @staticmethod
    def forward(ctx, x, weight, mod):
        x = x.contiguous()
        wf, wd = mod.packed()
        ctx.save_for_backward(x)
        ctx.wd, ctx.mod = wd, mod
        return ops.conv3x3s2(x, wf)

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = ops.conv3x3s2(dy, ctx.wd, dgrad=True) if ctx.needs_input_grad[0] else None
        dw = None
        if ctx.needs_input_grad[1]:
            mod = ctx.mod
            if mod.grad_direct and mod.weight.grad is not None:
                _arena_wgrad(ops.conv3x3s2_wgrad, x, dy, mod)
            else:
                dw = ops.conv3x3s2_wgrad(x, dy, mod.wgrad_workspace())
        return dx, dw, None


class Conv3x3(nn.Conv2d):
    def __init__(self, in_planes: int, planes: int, stride: int = 1):
        super().__init__(in_planes, planes, 3, stride, 1, bias=False)
        self._packed = None            # float32 [2, 2*C*9*C]: forward packing, dgrad packing (sized for the hi/lo split)
        self._packed_key = None        # (weight data_ptr, weight version) the packing was made from
        self._managed = False          # True while a trainer repacks all layers itself (pack_all)
        self.grad_direct = False       # trainer-owned gradient arena: wgrad adds straight into weight.grad
        self._ws = None

    def _buffers_for(self, device):
        if self._packed is None or self._packed.device != device:
            c = self.out_channels
            self._packed = torch.empty((2, 2 * c * 9 * c), dtype=torch.float32, device=device)
            self._packed_key = None
        return self._packed

    @property
    def is_transition(self) -> bool:
        return self.stride == (2, 2) and self.out_channels == 2 * self.in_channels

    def desc_row(self):
        p = self._buffers_for(self.weight.device)
        c = -self.in_channels if self.is_transition else self.out_channels     # c < 0: stride-2 packing
        return [self.weight.data_ptr(), p[0].data_ptr(), p[1].data_ptr(), c]

    def packed(self):
        p = self._buffers_for(self.weight.device)
        if not self._managed:
            key = (self.weight.data_ptr(), self.weight._version, MODE)
            if key != self._packed_key:
                descs = torch.tensor([self.desc_row()], dtype=torch.int64, device=self.weight.device)
                ops.conv3x3_pack(descs, self.out_channels, _MATH[MODE])
                self._packed_key = key
        return p[0], p[1]

    def wgrad_workspace(self):
        if self._ws is None or self._ws.device != self.weight.device:
            self._ws = ops.conv3x3_wgrad_workspace(self.out_channels, self.weight.device)
        return self._ws

    def hand_written(self, x) -> bool:
        return (MODE in _MATH and self.stride == (1, 1) and self.in_channels == self.out_channels
                and ops.conv3x3_supported(x, self.weight) and _call_math(self.out_channels, x) is not None)

    def forward_with_tap(self, x):
        """(conv(x), x') where x' aliases x for the identity shortcut; see _Conv3x3TapFn."""
        if self.hand_written(x) and torch.is_grad_enabled() and x.requires_grad:
            return _Conv3x3TapFn.apply(x, self.weight, self)
        return self.forward(x), x

    def forward(self, x):
        if MODE in _MATH and STRIDE2 and self.is_transition and ops.conv3x3s2_supported(x, self.weight):
            return _Conv3x3S2Fn.apply(x, self.weight, self)
        if self.hand_written(x):
            return _Conv3x3Fn.apply(x, self.weight, self)
        return super().forward(x)


class PackPlan:
    """One-launch repack of every eligible Conv3x3 of a model (for trainers whose optimiser writes the weights with a
    raw kernel).  The descriptor table is rebuilt when a weight moved (e.g. into the flat parameter arena)."""

    def __init__(self, model: nn.Module):
        self.mods = [m for m in model.modules() if isinstance(m, Conv3x3)
                     and ((m.stride == (1, 1) and m.in_channels == m.out_channels and m.out_channels in ops.CONV3X3_CHANNELS)
                          or (m.is_transition and m.in_channels in (16, 32)))]
        self._ptrs, self._descs = None, None
        for m in self.mods:
            m._managed = True

    def pack(self):
        if not self.mods or MODE not in _MATH:
            return
        ptrs = tuple(m.weight.data_ptr() for m in self.mods)
        if ptrs != self._ptrs:
            dev = self.mods[0].weight.device
            self._descs = torch.tensor([m.desc_row() for m in self.mods], dtype=torch.int64, device=dev)
            self._ptrs = ptrs
        ops.conv3x3_pack(self._descs, max(m.out_channels for m in self.mods), _MATH[MODE])

    def release(self):
        """Back to stand-alone mode: every module repacks its own weight on its next forward."""
        for m in self.mods:
            m._managed, m._packed_key = False, None
