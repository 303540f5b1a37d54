import importlib, os, sys, torch, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
pkg = importlib.import_module("cv_a-fan_b200"); ops = pkg.ops
dev = torch.device("cuda:0"); g = torch.Generator(device=dev).manual_seed(3)
def t(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); b.synchronize()
    return a.elapsed_time(b) * 1e3 / n
res = {}
cl = torch.relu(torch.randn(4, 256, 128, 128, device=dev, generator=g)); ad = cl + 0.01 * torch.randn(cl.shape, device=dev, generator=g)
res["sat_3pts_2mixed_4x256x128x128_us"] = t(lambda: pkg.segmentation.sat_sample_points(cl, ad, 3, (True, True)))
cl2 = torch.relu(torch.randn(8, 1024, 38, 63, device=dev, generator=g)); ad2 = cl2 + 0.01 * torch.randn(cl2.shape, device=dev, generator=g)
res["sat_5pts_4mixed_8x1024x38x63_us"] = t(lambda: pkg.segmentation.sat_sample_points(cl2, ad2, 5, (True, True, True, True)))
res["unfused_lerp3+mix4_8x1024x38x63_us"] = t(lambda: [ops.mix_feature(cl2, p) for p in pkg.segmentation.get_sample_points(cl2, ad2, 5)[1:]])
boxes = torch.rand(12000, 4, device=dev, generator=g) * 300; boxes[:, 2:] += boxes[:, :2] + 4
scores = torch.rand(12000, device=dev, generator=g)
res["nms_12000_us"] = t(lambda: ops.nms_flags(boxes, scores, 0.7))
import torchvision
res["torchvision_nms_12000_us"] = t(lambda: torchvision.ops.nms(boxes, scores, 0.7))
print(json.dumps(res))
