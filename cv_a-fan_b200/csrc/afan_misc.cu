// afan_misc.cu -- library identification, error strings, and the fused SGD step.
#include "afan_common.cuh"

namespace afan {

// torch.optim.SGD(momentum, weight_decay) over a flat arena, Classification/main_perturb.py:72-74,201.
// lr comes from device memory so that a captured CUDA graph follows the LR schedule.
template <int VEC>
__global__ void __launch_bounds__(kThreads)
sgd_momentum_kernel(float* __restrict__ param, const float* __restrict__ grad, float* __restrict__ buf, long long nv,
                    const float* __restrict__ lr_device, float momentum, float wd, float grad_scale) {
    const float lr = __ldg(lr_device);
    const long long stride = static_cast<long long>(gridDim.x) * kThreads;
    auto one = [&](float& p, float g, float& b) {
        const float d = fmaf(wd, p, g * grad_scale);
        b = fmaf(momentum, b, d);
        p = fmaf(-lr, b, p);
    };
    for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < nv; i += stride) {
        if constexpr (VEC == 4) {
            float4 p = reinterpret_cast<float4*>(param)[i], b = reinterpret_cast<float4*>(buf)[i];
            const float4 g = ld_stream(reinterpret_cast<const float4*>(grad) + i);
            one(p.x, g.x, b.x); one(p.y, g.y, b.y); one(p.z, g.z, b.z); one(p.w, g.w, b.w);
            reinterpret_cast<float4*>(param)[i] = p;
            reinterpret_cast<float4*>(buf)[i] = b;
        } else {
            float p = param[i], b = buf[i];
            one(p, ld_stream(grad + i), b);
            param[i] = p;
            buf[i] = b;
        }
    }
}

// FFMA pipe probe: 64 independent accumulators per thread, 8x8 outer-product updates (the inner-loop shape of the direct
// convolutions).  bench.py times it to obtain the fp32 FFMA peak of THIS device at its current clocks, the roofline
// denominator of the convolution kernels.
__global__ void __launch_bounds__(kThreads) ffma_probe_kernel(float* __restrict__ out, int iters) {
    float acc[8][8], a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = 1.0f + 1e-3f * (threadIdx.x + i); b[i] = 1.0f - 1e-3f * (threadIdx.x + 2 * i); }
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
#pragma unroll
        for (int i = 0; i < 8; ++i) a[i] += 1e-9f;
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * kThreads + threadIdx.x] = s;
}


// Option-A shortcut of a stage transition (Classification/resnet_s.py:60-63: F.pad(x[:, :, ::2, ::2], (0,0,0,0,p,p))): the
// library runs it as a strided copy + a zero fill + a padded copy (and three more launches backwards).  One pass each way:
// forward writes every element of y [n][c+2p][h/2][w/2] (zeros in the padded channels), backward writes every element of
// dx [n][c][h][w] (dy at the even positions, zero elsewhere).
__global__ void __launch_bounds__(kThreads)
shortcut_a_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, long long total, int c, int h, int w, int pad) {
    const int ho = (h + 1) >> 1, wo = (w + 1) >> 1, co = c + 2 * pad;
    const long long stride = static_cast<long long>(gridDim.x) * kThreads;
    for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < total; i += stride) {
        const int j = static_cast<int>(i % wo);
        long long t = i / wo;
        const int r = static_cast<int>(t % ho);
        t /= ho;
        const int ch = static_cast<int>(t % co) - pad;
        const long long n = t / co;
        y[i] = (ch >= 0 && ch < c) ? __ldg(x + ((n * c + ch) * h + 2 * r) * w + 2 * j) : 0.f;
    }
}

__global__ void __launch_bounds__(kThreads)
shortcut_a_bwd_kernel(const float* __restrict__ dy, float* __restrict__ dx, long long total, int c, int h, int w, int pad) {
    const int ho = (h + 1) >> 1, wo = (w + 1) >> 1, co = c + 2 * pad;
    const long long stride = static_cast<long long>(gridDim.x) * kThreads;
    for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < total; i += stride) {
        const int j = static_cast<int>(i % w);
        long long t = i / w;
        const int r = static_cast<int>(t % h);
        t /= h;
        const int ch = static_cast<int>(t % c);
        const long long n = t / c;
        dx[i] = ((r | j) & 1) ? 0.f : __ldg(dy + ((n * co + ch + pad) * ho + (r >> 1)) * wo + (j >> 1));
    }
}

// Weight / bias gradient of the classifier (nn.Linear at the end of resnet_s.py:93-95): dW[o][i] = sum_b dy[b][o] * x[b][i],
// db[o] = sum_b dy[b][o].  A 100 x 64 x 256 problem: the library's SIMT sgemm takes 91 us for it (ncu launch list); one CTA
// per output row, one thread per input feature, batch walked in order (deterministic), dy staged through shared memory.
__global__ void __launch_bounds__(256)
linear_wgrad_kernel(const float* __restrict__ dy, const float* __restrict__ x, float* __restrict__ dw, float* __restrict__ db,
                    int batch, int out_f, int in_f) {
    __shared__ float s_dy[1024];
    const int o = blockIdx.x;
    float bsum = 0.f;
    for (int i0 = 0; i0 < in_f; i0 += blockDim.x) {
        const int i = i0 + threadIdx.x;
        float acc = 0.f;
        for (int b0 = 0; b0 < batch; b0 += 1024) {
            const int nb = min(1024, batch - b0);
            __syncthreads();
            for (int b = threadIdx.x; b < nb; b += blockDim.x) s_dy[b] = __ldg(dy + static_cast<size_t>(b0 + b) * out_f + o);
            __syncthreads();
            if (i < in_f) {
                float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;                // four chains, combined in a fixed order
                int b = 0;
                for (; b + 4 <= nb; b += 4) {
                    a0 = fmaf(s_dy[b], __ldg(x + static_cast<size_t>(b0 + b) * in_f + i), a0);
                    a1 = fmaf(s_dy[b + 1], __ldg(x + static_cast<size_t>(b0 + b + 1) * in_f + i), a1);
                    a2 = fmaf(s_dy[b + 2], __ldg(x + static_cast<size_t>(b0 + b + 2) * in_f + i), a2);
                    a3 = fmaf(s_dy[b + 3], __ldg(x + static_cast<size_t>(b0 + b + 3) * in_f + i), a3);
                }
                for (; b < nb; ++b) a0 = fmaf(s_dy[b], __ldg(x + static_cast<size_t>(b0 + b) * in_f + i), a0);
                acc += (a0 + a1) + (a2 + a3);
            }
            if (i0 == 0 && threadIdx.x == 0 && db)
                for (int b = 0; b < nb; ++b) bsum += s_dy[b];
        }
        if (i < in_f) dw[static_cast<size_t>(o) * in_f + i] = acc;
    }
    if (threadIdx.x == 0 && db) db[o] = bsum;
}


// HBM stream probes: what a kernel that ONLY reads, ONLY writes, or copies can move on this device (bench.py reports them
// beside MEASURED_PEAKS.json's copy bandwidth: the roofline of a read-dominated kernel is the read figure, not the copy one).
// mode 0: read (sum into one float per CTA), 1: write, 2: copy; 128-bit accesses, 8 independent accesses per thread in flight.
template <int MODE>
__global__ void __launch_bounds__(kThreads) hbm_probe_kernel(const float4* __restrict__ src, float4* __restrict__ dst, long long nv,
                                                             float* __restrict__ sink) {
    const long long stride = static_cast<long long>(gridDim.x) * kThreads;
    float acc = 0.f;
    for (long long i0 = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i0 < nv; i0 += 8 * stride) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long i = i0 + u * stride;
            if (MODE != 1) v[u] = i < nv ? ld_stream(src + i) : make_float4(0.f, 0.f, 0.f, 0.f);
            else v[u] = make_float4(1.f, 2.f, 3.f, 4.f);
        }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const long long i = i0 + u * stride;
            if (MODE == 0) acc += (v[u].x + v[u].y) + (v[u].z + v[u].w);
            else if (i < nv) st_stream(dst + i, v[u]);
        }
    }
    if (MODE == 0 && acc == 123.456f) sink[blockIdx.x] = acc;          // keeps the loads alive
}

}  // namespace afan

using namespace afan;

// Launches the probe on 2 CTAs per SM; `out` holds 2 * sm_count * 256 floats.  Returns the FLOP count of the launch
// (2 * 64 * iters per thread) through *flops_out so the caller can divide by its own CUDA-event time.
AFAN_EXPORT int afan_ffma_probe(float* out, int64_t out_elems, int64_t iters, double* flops_out, afan_stream_t stream) {
    if (!out || !flops_out) return AFAN_ERR_NULL;
    const int grid = 2 * sm_count();
    if (iters <= 0 || out_elems < static_cast<int64_t>(grid) * kThreads) return AFAN_ERR_SIZE;
    ffma_probe_kernel<<<grid, kThreads, 0, static_cast<cudaStream_t>(stream)>>>(out, static_cast<int>(iters));
    *flops_out = 2.0 * 64.0 * static_cast<double>(iters) * kThreads * grid;
    return launch_status();
}

AFAN_EXPORT const char* afan_version(void) { return "afan_b200 0.1.0 (sm_100a)"; }

AFAN_EXPORT const char* afan_strerror(int code) {
    switch (code) {
        case AFAN_OK: return "ok";
        case AFAN_ERR_NULL: return "required pointer is NULL";
        case AFAN_ERR_SIZE: return "negative or inconsistent size";
        case AFAN_ERR_WORKSPACE: return "workspace missing, misaligned or too small";
        case AFAN_ERR_LAUNCH: return "CUDA kernel launch failed";
        case AFAN_ERR_UNSUPPORTED: return "shape not supported by the sm_100a kernels";
        default: return "unknown afan error code";
    }
}

AFAN_EXPORT int afan_device_info(int* sm_count_out, int* cc_major, int* cc_minor) {
    int dev = 0, sms = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) {
        cudaGetLastError();
        return AFAN_ERR_LAUNCH;
    }
    if (sm_count_out) *sm_count_out = sms;
    if (cc_major) *cc_major = major;
    if (cc_minor) *cc_minor = minor;
    return major == 10 ? AFAN_OK : AFAN_ERR_UNSUPPORTED;
}

AFAN_EXPORT int afan_sgd_momentum_f32(float* param, const float* grad, float* momentum_buf, int64_t n_elem,
                                      const float* lr_device, float momentum, float weight_decay, float grad_scale,
                                      afan_stream_t stream) {
    if (n_elem < 0) return AFAN_ERR_SIZE;
    if (n_elem == 0) return AFAN_OK;
    if (!param || !grad || !momentum_buf || !lr_device) return AFAN_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long cap = static_cast<long long>(sm_count()) * kCtasPerSm;
    if (n_elem % 4 == 0 && aligned16(param) && aligned16(grad) && aligned16(momentum_buf)) {
        const long long nv = n_elem / 4, want = (nv + kThreads - 1) / kThreads;
        sgd_momentum_kernel<4><<<static_cast<unsigned int>(want < cap ? want : cap), kThreads, 0, st>>>(
            param, grad, momentum_buf, nv, lr_device, momentum, weight_decay, grad_scale);
    } else {
        const long long want = (n_elem + kThreads - 1) / kThreads;
        sgd_momentum_kernel<1><<<static_cast<unsigned int>(want < cap ? want : cap), kThreads, 0, st>>>(
            param, grad, momentum_buf, n_elem, lr_device, momentum, weight_decay, grad_scale);
    }
    return launch_status();
}

AFAN_EXPORT int afan_shortcut_a_fwd_f32(const float* x, float* y, int64_t n, int64_t c, int64_t h, int64_t w, int64_t pad,
                                        afan_stream_t stream) {
    if (n < 0 || c < 0 || h < 0 || w < 0 || pad < 0) return AFAN_ERR_SIZE;
    const int64_t total = n * (c + 2 * pad) * ((h + 1) / 2) * ((w + 1) / 2);
    if (total == 0) return AFAN_OK;
    if (!x || !y) return AFAN_ERR_NULL;
    if (c + 2 * pad >= (int64_t(1) << 31) || h >= (int64_t(1) << 30) || w >= (int64_t(1) << 30)) return AFAN_ERR_UNSUPPORTED;
    const int64_t want = (total + kThreads - 1) / kThreads, cap = static_cast<int64_t>(sm_count()) * kCtasPerSm;
    shortcut_a_fwd_kernel<<<static_cast<unsigned int>(want < cap ? want : cap), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        x, y, total, static_cast<int>(c), static_cast<int>(h), static_cast<int>(w), static_cast<int>(pad));
    return launch_status();
}

AFAN_EXPORT int afan_shortcut_a_bwd_f32(const float* dy, float* dx, int64_t n, int64_t c, int64_t h, int64_t w, int64_t pad,
                                        afan_stream_t stream) {
    if (n < 0 || c < 0 || h < 0 || w < 0 || pad < 0) return AFAN_ERR_SIZE;
    const int64_t total = n * c * h * w;
    if (total == 0) return AFAN_OK;
    if (!dy || !dx) return AFAN_ERR_NULL;
    if (c + 2 * pad >= (int64_t(1) << 31) || h >= (int64_t(1) << 30) || w >= (int64_t(1) << 30)) return AFAN_ERR_UNSUPPORTED;
    const int64_t want = (total + kThreads - 1) / kThreads, cap = static_cast<int64_t>(sm_count()) * kCtasPerSm;
    shortcut_a_bwd_kernel<<<static_cast<unsigned int>(want < cap ? want : cap), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        dy, dx, total, static_cast<int>(c), static_cast<int>(h), static_cast<int>(w), static_cast<int>(pad));
    return launch_status();
}

AFAN_EXPORT int afan_linear_wgrad_f32(const float* dy, const float* x, float* dweight, float* dbias, int64_t batch,
                                      int64_t out_features, int64_t in_features, afan_stream_t stream) {
    if (batch < 0 || out_features < 0 || in_features < 0) return AFAN_ERR_SIZE;
    if (out_features == 0) return AFAN_OK;
    if (!dweight || (batch > 0 && (!dy || !x))) return AFAN_ERR_NULL;
    if (batch >= (int64_t(1) << 31) || out_features >= (int64_t(1) << 31) || in_features >= (int64_t(1) << 31)) return AFAN_ERR_UNSUPPORTED;
    const int threads = in_features >= 256 ? 256 : (in_features > 32 ? static_cast<int>((in_features + 31) / 32 * 32) : 32);
    linear_wgrad_kernel<<<static_cast<unsigned int>(out_features), threads, 0, static_cast<cudaStream_t>(stream)>>>(
        dy, x, dweight, dbias, static_cast<int>(batch), static_cast<int>(out_features), static_cast<int>(in_features));
    return launch_status();
}

AFAN_EXPORT int afan_hbm_probe(int mode, const float* src, float* dst, int64_t n_elem, float* sink, afan_stream_t stream) {
    if (n_elem < 0 || mode < 0 || mode > 2) return AFAN_ERR_SIZE;
    if ((mode != 1 && !src) || (mode != 0 && !dst) || !sink) return AFAN_ERR_NULL;
    if ((mode != 1 && !aligned16(src)) || (mode != 0 && !aligned16(dst))) return AFAN_ERR_UNSUPPORTED;
    const long long nv = n_elem / 4;
    const int grid = sm_count() * kCtasPerSm;                         // sink: at least `grid` floats
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const float4* s4 = reinterpret_cast<const float4*>(src);
    float4* d4 = reinterpret_cast<float4*>(dst);
    if (mode == 0) hbm_probe_kernel<0><<<grid, kThreads, 0, st>>>(s4, d4, nv, sink);
    else if (mode == 1) hbm_probe_kernel<1><<<grid, kThreads, 0, st>>>(s4, d4, nv, sink);
    else hbm_probe_kernel<2><<<grid, kThreads, 0, st>>>(s4, d4, nv, sink);
    return launch_status();
}
