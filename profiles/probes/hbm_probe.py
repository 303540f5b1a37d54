"""What can a read-only / write-only / copy kernel stream on this B200?  (afan_hbm_probe; 1 GiB buffers, CUDA events, graph replay)"""
import importlib, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
PKG = importlib.import_module("cv_a-fan_b200")
L = PKG._lib
dev = torch.device("cuda:0")
n = 256 * 1024 * 1024                       # floats = 1 GiB
src = torch.randn(n, device=dev)
dst = torch.empty(n, device=dev)
sink = torch.zeros(148 * 8 * 2, device=dev)
for mode, name, bytes_ in ((0, "read", 4 * n), (1, "write", 4 * n), (2, "copy", 8 * n)):
    def run():
        L.check(L.lib().afan_hbm_probe(mode, L.f32(src), L.f32(dst), n, L.f32(sink), L.stream()), "afan_hbm_probe")
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10):
        run()
    b.record(); b.synchronize()
    ms = a.elapsed_time(b) / 10
    print(name, "GB/s", round(bytes_ / ms / 1e6, 1), flush=True)
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for _ in range(10):
    dst.copy_(src)
b.record(); b.synchronize()
print("torch copy GB/s", round(8 * n / (a.elapsed_time(b) / 10) / 1e6, 1))
a.record()
for _ in range(10):
    src.sum()
b.record(); b.synchronize()
print("torch sum (read) GB/s", round(4 * n / (a.elapsed_time(b) / 10) / 1e6, 1))
