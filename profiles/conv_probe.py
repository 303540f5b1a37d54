import torch, time, json
torch.backends.cudnn.benchmark = True
dev = torch.device("cuda:0")
res = {}
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    for fmt_name, fmt in (("nchw", torch.contiguous_format), ("nhwc", torch.channels_last)):
        for (n, c, h) in ((128, 16, 32), (128, 32, 16), (128, 64, 8), (256, 32, 16)):
            conv = torch.nn.Conv2d(c, c, 3, 1, 1, bias=False).to(dev).to(memory_format=fmt)
            x = torch.randn(n, c, h, h, device=dev).to(memory_format=fmt).requires_grad_(True)
            def step():
                y = conv(x)
                gx, gw = torch.autograd.grad(y, (x, conv.weight), torch.ones_like(y))
            for _ in range(5): step()
            g = torch.cuda.CUDAGraph()
            s = torch.cuda.Stream()
            with torch.cuda.stream(s):
                step()
            torch.cuda.synchronize()
            with torch.cuda.graph(g):
                for _ in range(20): step()
            g.replay(); torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5): g.replay()
            b.record(); b.synchronize()
            res[f"{'tf32' if tf32 else 'fp32'}_{fmt_name}_{n}x{c}x{h}"] = round(a.elapsed_time(b) * 1e3 / 100, 1)
print(json.dumps(res, indent=0))
