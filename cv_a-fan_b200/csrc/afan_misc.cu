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
