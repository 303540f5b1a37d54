// Throughput probe (B200): FFMA outer-product loop vs mma.sync m16n8k8 TF32 -- decides the math pipe for the
// hand-written 3x3 convolutions.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_probe pipe_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__global__ void __launch_bounds__(256) ffma_kernel(float* out, const float* in, int iters) {
    float acc[8][8];
    float a[8], b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a[i] = in[threadIdx.x + i * 32]; b[i] = in[threadIdx.x + 256 + i * 32]; }
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
        for (int i = 0; i < 8; ++i) { a[i] += 1e-9f; }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) s += acc[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void __launch_bounds__(256) mma_tf32_kernel(float* out, const float* in, int iters) {
    float c[NT][4];
    uint32_t a[4], b[NT][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(in[threadIdx.x + i * 32]) & 0xffffe000u;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        b[t][0] = __float_as_uint(in[threadIdx.x + 128 + t * 32]) & 0xffffe000u;
        b[t][1] = __float_as_uint(in[threadIdx.x + 512 + t * 32]) & 0xffffe000u;
#pragma unroll
        for (int i = 0; i < 4; ++i) c[t][i] = 0.f;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
            asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[t][0]), "r"(b[t][1]));
    }
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += c[t][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NT>
__global__ void __launch_bounds__(256) mma_bf16_kernel(float* out, const float* in, int iters) {
    float c[NT][4];
    uint32_t a[4], b[NT][2];
#pragma unroll
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(in[threadIdx.x + i * 32]);
#pragma unroll
    for (int t = 0; t < NT; ++t) {
        b[t][0] = __float_as_uint(in[threadIdx.x + 128 + t * 32]);
        b[t][1] = __float_as_uint(in[threadIdx.x + 512 + t * 32]);
#pragma unroll
        for (int i = 0; i < 4; ++i) c[t][i] = 0.f;
    }
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int t = 0; t < NT; ++t)
            asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                         : "+f"(c[t][0]), "+f"(c[t][1]), "+f"(c[t][2]), "+f"(c[t][3])
                         : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[t][0]), "r"(b[t][1]));
    }
    float s = 0.f;
#pragma unroll
    for (int t = 0; t < NT; ++t)
#pragma unroll
        for (int i = 0; i < 4; ++i) s += c[t][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    return ms;
}

int main() {
    float *in, *out;
    cudaMalloc(&in, 1 << 20); cudaMalloc(&out, 148 * 8 * 256 * 4 * 4);
    cudaMemset(in, 0, 1 << 20);
    const int iters = 20000;
    for (int cps = 1; cps <= 4; cps *= 2) {
        int grid = 148 * cps;
        float ms = time_ms([&] { ffma_kernel<<<grid, 256>>>(out, in, iters); });
        double fl = 2.0 * 64 * iters * 256.0 * grid;
        printf("ffma 8x8 outer  %d CTA/SM x256thr: %.3f ms  %.1f TFLOP/s\n", cps, ms, fl / ms / 1e9);
    }
    for (int cps = 1; cps <= 4; cps *= 2) {
        int grid = 148 * cps;
        float ms = time_ms([&] { mma_tf32_kernel<8><<<grid, 256>>>(out, in, iters); });
        double fl = 2.0 * 16 * 8 * 8 * 8 * iters * 8.0 * grid;
        printf("mma.sync tf32 m16n8k8 x8 tiles %d CTA/SM x8 warps: %.3f ms  %.1f TFLOP/s\n", cps, ms, fl / ms / 1e9);
        ms = time_ms([&] { mma_tf32_kernel<4><<<grid, 256>>>(out, in, iters); });
        fl = 2.0 * 16 * 8 * 8 * 4 * iters * 8.0 * grid;
        printf("mma.sync tf32 m16n8k8 x4 tiles %d CTA/SM x8 warps: %.3f ms  %.1f TFLOP/s\n", cps, ms, fl / ms / 1e9);
        ms = time_ms([&] { mma_bf16_kernel<8><<<grid, 256>>>(out, in, iters); });
        fl = 2.0 * 16 * 8 * 16 * 8 * iters * 8.0 * grid;
        printf("mma.sync bf16 m16n8k16 x8 tiles %d CTA/SM x8 warps: %.3f ms  %.1f TFLOP/s\n", cps, ms, fl / ms / 1e9);
    }
    printf("err=%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
