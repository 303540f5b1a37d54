// afan_common.cuh -- shared device/host helpers for the sm_100a A-FAN kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>

#include "../../include/afan_b200.h"

#define AFAN_EXPORT extern "C" __attribute__((visibility("default")))

namespace afan {

constexpr int kThreads = 256;          // CTA size of every streaming kernel
constexpr int kSmCountB200 = 148;      // grids are sized in multiples of this
constexpr int kCtasPerSm = 8;          // 8 x 256 threads = 2048 resident threads / SM

__host__ inline int sm_count() {
    static int cached = 0;             // benign race: every thread computes the same value
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) == cudaSuccess &&
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0)
            cached = n;
        else
            return kSmCountB200;
    }
    return cached;
}

__host__ inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

__host__ inline int launch_status() {
    return cudaPeekAtLastError() == cudaSuccess ? AFAN_OK : (cudaGetLastError(), AFAN_ERR_LAUNCH);
}

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------
// Kernels that begin with pdl_wait() may be launched with cudaLaunchAttributeProgrammaticStreamSerialization: their
// CTAs are scheduled while the previous kernel of the stream drains, and block in pdl_wait() until that kernel's
// memory is visible -- the ~2 us launch ramp of every small kernel overlaps its predecessor's tail.  Both
// instructions are no-ops for a normal launch.  AFAN_PDL=0 in the environment disables the launch attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

__host__ inline bool pdl_enabled() {
    static int cached = -1;            // benign race
    if (cached < 0) {
        const char* e = getenv("AFAN_PDL");
        cached = (e && e[0] == '0') ? 0 : 1;
    }
    return cached == 1;
}

// <<<grid, block, smem, stream>>> with the PDL attribute
template <typename K, typename... Args>
__host__ inline int launch_pdl(K kernel, dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr{};
    attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr.val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = &attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    if (cudaLaunchKernelEx(&cfg, kernel, args...) != cudaSuccess) { cudaGetLastError(); return AFAN_ERR_LAUNCH; }
    return cudaPeekAtLastError() == cudaSuccess ? AFAN_OK : (cudaGetLastError(), AFAN_ERR_LAUNCH);
}

// ---- 128-bit global access with cache hints --------------------------------------------
// .cs (cache-streaming / evict-first): data touched once (gradients, write-only outputs).
__device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }
__device__ __forceinline__ float ld_stream(const float* p) { return __ldcs(p); }
__device__ __forceinline__ void st_stream(float4* p, float4 v) { __stcs(p, v); }
__device__ __forceinline__ void st_stream(float* p, float v) { __stcs(p, v); }
// read-only path for tensors that stay hot in L2 across PGD steps (the cached clean feature)
__device__ __forceinline__ float4 ld_ro(const float4* p) { return __ldg(p); }
__device__ __forceinline__ float ld_ro(const float* p) { return __ldg(p); }

// ---- reductions: warp shuffle, then a shared-memory tree across the CTA's warps ---------
template <typename T>
__device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Sum of (a, b) over the CTA, result valid in thread 0.  `scratch` holds 2*32 doubles.
__device__ __forceinline__ void block_sum2(double& a, double& b, double* scratch) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = (blockDim.x + 31) >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    if (lane == 0) { scratch[warp] = a; scratch[32 + warp] = b; }
    __syncthreads();
    if (warp == 0) {
        a = lane < nwarp ? scratch[lane] : 0.0;
        b = lane < nwarp ? scratch[32 + lane] : 0.0;
        a = warp_sum(a);
        b = warp_sum(b);
    }
    __syncthreads();
}

// "last CTA done" ticket: returns true in ALL threads of the CTA that arrives last of `expected`
// at `counter`; that CTA sees every other CTA's prior global writes.  The counter is reset to 0
// by the last CTA so the workspace can be reused by the next launch without a memset.
__device__ __forceinline__ bool last_cta_arrives(unsigned int* counter, unsigned int expected, int* smem_flag) {
    __threadfence();                   // publish this CTA's partials
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned int t = atomicAdd(counter, 1u);
        *smem_flag = (t == expected - 1u);
        if (t == expected - 1u) *counter = 0u;
    }
    __syncthreads();
    bool last = *smem_flag != 0;
    if (last) __threadfence();         // acquire side
    return last;
}

}  // namespace afan
