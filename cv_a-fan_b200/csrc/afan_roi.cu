// afan_roi.cu -- ROIAlign forward / backward for sm_100a.
//
// Replaces Detection/support/src/cuda/ROIAlign_cuda.cu:64-122 (forward) and :177-254 (backward): bilinear-sampled
// average pooling of a [N,C,H,W] feature map into [R,C,PH,PW] with the legacy (non-"aligned") box coordinates and an
// adaptive sampling grid (sampling_ratio <= 0 -> ceil(roi_size / pooled_size) points per bin), as used by the
// Faster R-CNN pooler (Detection/roi/pooler.py:35).  One thread per pooled element, pw fastest so that a warp walks
// neighbouring bins of one channel plane (its 4-neighbour gathers stay within a few L1 lines); the per-ROI geometry is
// hoisted out of the sampling loops.  Backward scatters with red.global.add.f32 (summation order is not deterministic,
// exactly like the reference's atomicAdd; parity is tolerance-based).
#include "afan_common.cuh"

namespace afan {

struct RoiGeom {
    float start_w, start_h, bin_w, bin_h;
    int grid_w, grid_h, batch;
};

__device__ __forceinline__ RoiGeom roi_geometry(const float* __restrict__ roi, float scale, int ph, int pw, int sampling_ratio) {
    RoiGeom g;
    g.batch = static_cast<int>(roi[0]);
    g.start_w = roi[1] * scale;
    g.start_h = roi[2] * scale;
    // explicit round-to-nearest ops: an FMA-contracted coordinate can land on the other side of an integer / of the
    // "outside the map" test and change which taps a sample uses
    const float roi_w = fmaxf(__fsub_rn(__fmul_rn(roi[3], scale), g.start_w), 1.f);      // malformed ROIs are forced to 1x1
    const float roi_h = fmaxf(__fsub_rn(__fmul_rn(roi[4], scale), g.start_h), 1.f);
    g.bin_w = roi_w / static_cast<float>(pw);
    g.bin_h = roi_h / static_cast<float>(ph);
    g.grid_w = sampling_ratio > 0 ? sampling_ratio : static_cast<int>(ceilf(roi_w / pw));
    g.grid_h = sampling_ratio > 0 ? sampling_ratio : static_cast<int>(ceilf(roi_h / ph));
    return g;
}

// start + p * bin + (i + .5) * bin / grid, evaluated left to right with one rounding per operation
__device__ __forceinline__ float sample_coord(float start, int p, float bin, int i, int grid) {
    return __fadd_rn(__fadd_rn(start, __fmul_rn(static_cast<float>(p), bin)),
                     __fdiv_rn(__fmul_rn(static_cast<float>(i) + .5f, bin), static_cast<float>(grid)));
}

struct Tap {                 // the four neighbours of one sample point and their bilinear weights
    int lo, hi_x, hi_y, hi_xy;   // element offsets inside the channel plane; lo < 0 -> sample outside the map
    float w1, w2, w3, w4;
};

__device__ __forceinline__ Tap make_tap(float y, float x, int height, int width) {
    Tap t;
    if (y < -1.0f || y > height || x < -1.0f || x > width) {
        t.lo = -1; t.hi_x = t.hi_y = t.hi_xy = 0; t.w1 = t.w2 = t.w3 = t.w4 = 0.f;
        return t;
    }
    y = fmaxf(y, 0.f);
    x = fmaxf(x, 0.f);
    int y0 = static_cast<int>(y), x0 = static_cast<int>(x), y1, x1;
    if (y0 >= height - 1) { y1 = y0 = height - 1; y = static_cast<float>(y0); } else { y1 = y0 + 1; }
    if (x0 >= width - 1) { x1 = x0 = width - 1; x = static_cast<float>(x0); } else { x1 = x0 + 1; }
    const float ly = y - y0, lx = x - x0, hy = 1.f - ly, hx = 1.f - lx;
    t.lo = y0 * width + x0; t.hi_x = y0 * width + x1; t.hi_y = y1 * width + x0; t.hi_xy = y1 * width + x1;
    t.w1 = hy * hx; t.w2 = hy * lx; t.w3 = ly * hx; t.w4 = ly * lx;
    return t;
}

__global__ void __launch_bounds__(kThreads)
roi_align_fwd_kernel(const float* __restrict__ feat, const float* __restrict__ rois, float* __restrict__ out, long long total,
                     int channels, int height, int width, int ph, int pw, float scale, int sampling_ratio) {
    const long long stride = static_cast<long long>(gridDim.x) * kThreads;
    for (long long idx = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; idx < total; idx += stride) {
        const int ow = static_cast<int>(idx % pw), oh = static_cast<int>((idx / pw) % ph);
        const int c = static_cast<int>((idx / pw / ph) % channels), r = static_cast<int>(idx / pw / ph / channels);
        const RoiGeom g = roi_geometry(rois + 5 * r, scale, ph, pw, sampling_ratio);
        const float* plane = feat + (static_cast<size_t>(g.batch) * channels + c) * height * width;
        const float count = static_cast<float>(g.grid_h * g.grid_w);
        float acc = 0.f;
        for (int iy = 0; iy < g.grid_h; ++iy) {
            const float y = sample_coord(g.start_h, oh, g.bin_h, iy, g.grid_h);
            for (int ix = 0; ix < g.grid_w; ++ix) {
                const float x = sample_coord(g.start_w, ow, g.bin_w, ix, g.grid_w);
                const Tap t = make_tap(y, x, height, width);
                if (t.lo >= 0)
                    acc += t.w1 * __ldg(plane + t.lo) + t.w2 * __ldg(plane + t.hi_x) + t.w3 * __ldg(plane + t.hi_y) +
                           t.w4 * __ldg(plane + t.hi_xy);
            }
        }
        out[idx] = acc / count;
    }
}

__global__ void __launch_bounds__(kThreads)
roi_align_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ rois, float* __restrict__ dfeat, long long total,
                     int channels, int height, int width, int ph, int pw, float scale, int sampling_ratio) {
    const long long stride = static_cast<long long>(gridDim.x) * kThreads;
    for (long long idx = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; idx < total; idx += stride) {
        const int ow = static_cast<int>(idx % pw), oh = static_cast<int>((idx / pw) % ph);
        const int c = static_cast<int>((idx / pw / ph) % channels), r = static_cast<int>(idx / pw / ph / channels);
        const RoiGeom g = roi_geometry(rois + 5 * r, scale, ph, pw, sampling_ratio);
        float* plane = dfeat + (static_cast<size_t>(g.batch) * channels + c) * height * width;
        const float count = static_cast<float>(g.grid_h * g.grid_w);
        const float d = __ldcs(dout + idx);
        for (int iy = 0; iy < g.grid_h; ++iy) {
            const float y = sample_coord(g.start_h, oh, g.bin_h, iy, g.grid_h);
            for (int ix = 0; ix < g.grid_w; ++ix) {
                const float x = sample_coord(g.start_w, ow, g.bin_w, ix, g.grid_w);
                const Tap t = make_tap(y, x, height, width);
                if (t.lo >= 0) {
                    atomicAdd(plane + t.lo, d * t.w1 / count);
                    atomicAdd(plane + t.hi_x, d * t.w2 / count);
                    atomicAdd(plane + t.hi_y, d * t.w3 / count);
                    atomicAdd(plane + t.hi_xy, d * t.w4 / count);
                }
            }
        }
    }
}

inline int roi_grid(long long total) {
    const long long want = (total + kThreads - 1) / kThreads, cap = static_cast<long long>(sm_count()) * kCtasPerSm * 4;
    return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

}  // namespace afan

using namespace afan;

static int roi_check(int64_t n, int64_t c, int64_t h, int64_t w, int64_t r, int64_t ph, int64_t pw) {
    if (n < 0 || c < 0 || h < 0 || w < 0 || r < 0 || ph <= 0 || pw <= 0) return AFAN_ERR_SIZE;
    if (h * w >= (int64_t(1) << 31) || c >= (int64_t(1) << 31)) return AFAN_ERR_UNSUPPORTED;
    return AFAN_OK;
}

AFAN_EXPORT int afan_roi_align_fwd_f32(const float* feat, const float* rois, float* out, int64_t n, int64_t c, int64_t h,
                                       int64_t w, int64_t r, int64_t ph, int64_t pw, float spatial_scale,
                                       int sampling_ratio, afan_stream_t stream) {
    int rc = roi_check(n, c, h, w, r, ph, pw);
    if (rc != AFAN_OK) return rc;
    const long long total = static_cast<long long>(r) * c * ph * pw;
    if (total == 0) return AFAN_OK;
    if (!feat || !rois || !out) return AFAN_ERR_NULL;
    roi_align_fwd_kernel<<<roi_grid(total), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        feat, rois, out, total, static_cast<int>(c), static_cast<int>(h), static_cast<int>(w), static_cast<int>(ph),
        static_cast<int>(pw), spatial_scale, sampling_ratio);
    return launch_status();
}

AFAN_EXPORT int afan_roi_align_bwd_f32(const float* dout, const float* rois, float* dfeat, int64_t n, int64_t c, int64_t h,
                                       int64_t w, int64_t r, int64_t ph, int64_t pw, float spatial_scale,
                                       int sampling_ratio, afan_stream_t stream) {
    int rc = roi_check(n, c, h, w, r, ph, pw);
    if (rc != AFAN_OK) return rc;
    if (n * c * h * w == 0) return AFAN_OK;
    if (!dfeat) return AFAN_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (cudaMemsetAsync(dfeat, 0, static_cast<size_t>(n) * c * h * w * sizeof(float), st) != cudaSuccess) {
        cudaGetLastError();
        return AFAN_ERR_LAUNCH;
    }
    const long long total = static_cast<long long>(r) * c * ph * pw;
    if (total == 0) return AFAN_OK;
    if (!dout || !rois) return AFAN_ERR_NULL;
    roi_align_bwd_kernel<<<roi_grid(total), kThreads, 0, st>>>(dout, rois, dfeat, total, static_cast<int>(c), static_cast<int>(h),
                                                              static_cast<int>(w), static_cast<int>(ph), static_cast<int>(pw),
                                                              spatial_scale, sampling_ratio);
    return launch_status();
}
