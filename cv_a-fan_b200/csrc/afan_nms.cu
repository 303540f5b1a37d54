// afan_nms.cu -- greedy non-maximum suppression for sm_100a, fully on the device.
//
// Replaces Detection/support/src/cuda/nms.cu:23-131 (the only native code of the reference besides ROIAlign):
// a 64x64-tile bitmask kernel followed by a D2H copy of the whole N x N/64 mask and a serial CPU sweep per image
// per forward.  Semantics kept: boxes sorted by score (descending), legacy "+1" areas (devIoU :13-21), a box is
// suppressed when IoU > threshold with an earlier kept box.  Here the sweep stays on the GPU:
//   nms_mask_kernel   only the upper-triangular tiles (row block <= col block); one 64-bit word per (box, col block)
//   nms_sweep_kernel  ONE CTA walks the 64-box blocks in order: the block's 64 diagonal words are staged in shared
//                     memory, a single lane resolves the block's internal dependencies on them (<= 64 steps), then all threads OR the mask rows
//                     of the newly kept boxes into the running `removed` bit vector in shared memory.
// Outputs live on the device: keep flags indexed by ORIGINAL box index + the count; no host synchronisation.
#include "afan_common.cuh"

namespace afan {

constexpr int kNmsTile = 64;
constexpr int kNmsSweepThreads = 256;

__device__ __forceinline__ float iou_plus_one(const float4 a, const float4 b) {
    const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
    const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
    const float width = fmaxf(__fadd_rn(__fsub_rn(right, left), 1.f), 0.f);
    const float height = fmaxf(__fadd_rn(__fsub_rn(bottom, top), 1.f), 0.f);
    const float inter = __fmul_rn(width, height);
    const float sa = __fmul_rn(__fadd_rn(__fsub_rn(a.z, a.x), 1.f), __fadd_rn(__fsub_rn(a.w, a.y), 1.f));
    const float sb = __fmul_rn(__fadd_rn(__fsub_rn(b.z, b.x), 1.f), __fadd_rn(__fsub_rn(b.w, b.y), 1.f));
    return __fdiv_rn(inter, __fsub_rn(__fadd_rn(sa, sb), inter));
}

__global__ void __launch_bounds__(kNmsTile)
nms_mask_kernel(const float4* __restrict__ boxes, unsigned long long* __restrict__ mask, int n, int col_blocks, float thr) {
    const int row_blk = blockIdx.y, col_blk = blockIdx.x;
    if (row_blk > col_blk) return;                                   // lower triangle is never read by the sweep
    __shared__ float4 tile[kNmsTile];
    const int col_size = min(n - col_blk * kNmsTile, kNmsTile), row_size = min(n - row_blk * kNmsTile, kNmsTile);
    if (threadIdx.x < col_size) tile[threadIdx.x] = boxes[col_blk * kNmsTile + threadIdx.x];
    __syncthreads();
    if (threadIdx.x < row_size) {
        const int cur = row_blk * kNmsTile + threadIdx.x;
        const float4 me = boxes[cur];
        unsigned long long bits = 0ULL;
        const int start = (row_blk == col_blk) ? threadIdx.x + 1 : 0;
        for (int i = start; i < col_size; ++i)
            if (iou_plus_one(me, tile[i]) > thr) bits |= 1ULL << i;
        mask[static_cast<size_t>(cur) * col_blocks + col_blk] = bits;
    }
}

__global__ void __launch_bounds__(kNmsSweepThreads)
nms_sweep_kernel(const unsigned long long* __restrict__ mask, const long long* __restrict__ order,
                 unsigned char* __restrict__ keep_flags, int* __restrict__ count_out, int n, int col_blocks) {
    extern __shared__ unsigned long long removed[];                  // col_blocks words
    __shared__ unsigned long long s_keep, s_diag[kNmsTile];
    for (int j = threadIdx.x; j < col_blocks; j += kNmsSweepThreads) removed[j] = 0ULL;
    for (int i = threadIdx.x; i < n; i += kNmsSweepThreads) keep_flags[i] = 0;
    __syncthreads();
    int kept = 0;
    for (int b = 0; b < col_blocks; ++b) {
        const int size = min(n - b * kNmsTile, kNmsTile);
        if (threadIdx.x < size)                                      // the block's 64 diagonal words, fetched in parallel
            s_diag[threadIdx.x] = mask[static_cast<size_t>(b * kNmsTile + threadIdx.x) * col_blocks + b];
        __syncthreads();
        if (threadIdx.x == 0) {                                      // resolve the block's internal chain out of shared memory
            unsigned long long rem = removed[b], keep = 0ULL;
            for (int i = 0; i < size; ++i)
                if (!((rem >> i) & 1ULL)) {
                    keep |= 1ULL << i;
                    rem |= s_diag[i];
                }
            s_keep = keep;
        }
        __syncthreads();
        const unsigned long long keep = s_keep;
        kept += __popcll(keep);
        for (int j = b + 1 + threadIdx.x; j < col_blocks; j += kNmsSweepThreads) {   // OR the kept rows into `removed`
            unsigned long long acc = removed[j], k = keep;
            while (k) {
                const int i = __ffsll(static_cast<long long>(k)) - 1;
                k &= k - 1;
                acc |= mask[static_cast<size_t>(b * kNmsTile + i) * col_blocks + j];
            }
            removed[j] = acc;
        }
        for (int i = threadIdx.x; i < size; i += kNmsSweepThreads)
            if ((keep >> i) & 1ULL) keep_flags[order[b * kNmsTile + i]] = 1;        // flag by ORIGINAL index
        __syncthreads();
    }
    if (threadIdx.x == 0 && count_out) *count_out = kept;
}

}  // namespace afan

using namespace afan;

AFAN_EXPORT int64_t afan_nms_workspace_bytes(int64_t n) {
    if (n < 0) return AFAN_ERR_SIZE;
    const int64_t col_blocks = (n + kNmsTile - 1) / kNmsTile;
    return n * col_blocks * 8 + 256;
}

AFAN_EXPORT int afan_nms_f32(const float* boxes_sorted, const int64_t* order, float threshold, uint8_t* keep_flags,
                             int32_t* count_out, void* workspace, int64_t workspace_bytes, int64_t n, afan_stream_t stream) {
    if (n < 0) return AFAN_ERR_SIZE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        if (count_out) cudaMemsetAsync(count_out, 0, sizeof(int32_t), st);
        return AFAN_OK;
    }
    if (!boxes_sorted || !order || !keep_flags) return AFAN_ERR_NULL;
    if (!aligned16(boxes_sorted)) return AFAN_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < afan_nms_workspace_bytes(n)) return AFAN_ERR_WORKSPACE;
    const int64_t col_blocks = (n + kNmsTile - 1) / kNmsTile;
    if (col_blocks > 65535 || col_blocks * 8 > 200 * 1024) return AFAN_ERR_UNSUPPORTED;       // n <= 1.6 M boxes
    unsigned long long* mask = static_cast<unsigned long long*>(workspace);
    dim3 grid(static_cast<unsigned int>(col_blocks), static_cast<unsigned int>(col_blocks));
    nms_mask_kernel<<<grid, kNmsTile, 0, st>>>(reinterpret_cast<const float4*>(boxes_sorted), mask, static_cast<int>(n),
                                              static_cast<int>(col_blocks), threshold);
    const size_t smem = static_cast<size_t>(col_blocks) * 8;
    if (smem > 48 * 1024)
        cudaFuncSetAttribute(nms_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    nms_sweep_kernel<<<1, kNmsSweepThreads, smem, st>>>(mask, reinterpret_cast<const long long*>(order), keep_flags, count_out,
                                                       static_cast<int>(n), static_cast<int>(col_blocks));
    return launch_status();
}
