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
# ROIAlign at the Faster R-CNN pooler shape: 256 ROIs x 1024 channels, 14 x 14 bins, 38 x 63 map (8 images)
feat = torch.randn(8, 1024, 38, 63, device=dev, generator=g)
r = 256
x1 = torch.rand(r, device=dev, generator=g) * 800; y1 = torch.rand(r, device=dev, generator=g) * 480
rois = torch.stack([torch.randint(0, 8, (r,), device=dev, generator=g).float(), x1, y1,
                    x1 + 16 + torch.rand(r, device=dev, generator=g) * 400, y1 + 16 + torch.rand(r, device=dev, generator=g) * 300], 1).contiguous()
dout = torch.randn(r, 1024, 14, 14, device=dev, generator=g)
L = pkg._lib.lib(); st = pkg._lib.stream
out = torch.empty(r, 1024, 14, 14, device=dev); dfeat = torch.empty_like(feat)
roi_res = {}
roi_res["roi_align_fwd_256x1024_us"] = t(lambda: pkg._lib.check(L.afan_roi_align_fwd_f32(feat.data_ptr(), rois.data_ptr(), out.data_ptr(), 8, 1024, 38, 63, r, 14, 14, 1 / 16, 0, st()), "afan_roi_align_fwd_f32"))
roi_res["roi_align_bwd_256x1024_us"] = t(lambda: pkg._lib.check(L.afan_roi_align_bwd_f32(dout.data_ptr(), rois.data_ptr(), dfeat.data_ptr(), 8, 1024, 38, 63, r, 14, 14, 1 / 16, 0, st()), "afan_roi_align_bwd_f32"))
roi_res["roi_path"] = ("forward: " + ("one thread per element" if os.environ.get("AFAN_ROI_LEGACY") == "1" else "separable tap tables in smem")
                       + "; backward: " + ("plane-resident smem scatter" if os.environ.get("AFAN_ROI_PLANE_BWD") == "1" else "global atomics"))
print(json.dumps(roi_res))
