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


// Batched form for the proposal layer (Detection/rpn/region_proposal_network.py:244-256: per image NMS at 0.7 over the
// ranked boxes, then `[:post_nms_top_n]`): one launch for the whole batch, ONE CTA PER IMAGE sweeping in parallel, the
// sweep stops as soon as `max_keep` boxes are kept (what the reference's slice throws away is never computed), and the
// kept boxes are written compacted, in rank order, straight into the zero-padded proposal tensor.  The row ORs of a
// block's kept boxes are issued four loads at a time (the single-image sweep above waits for each L2 round trip).
__global__ void __launch_bounds__(kNmsTile)
nms_mask_batched_kernel(const float4* __restrict__ boxes_all, unsigned long long* __restrict__ mask_all, int n, int col_blocks, float thr) {
    const int row_blk = blockIdx.y, col_blk = blockIdx.x;
    if (row_blk > col_blk) return;
    const float4* boxes = boxes_all + static_cast<size_t>(blockIdx.z) * n;
    unsigned long long* mask = mask_all + static_cast<size_t>(blockIdx.z) * n * col_blocks;
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
nms_sweep_batched_kernel(const unsigned long long* __restrict__ mask_all, const float4* __restrict__ boxes_all,
                         float4* __restrict__ kept_boxes, unsigned char* __restrict__ keep_flags, int* __restrict__ counts,
                         int n, int col_blocks, int max_keep) {
    extern __shared__ unsigned long long removed[];                  // col_blocks words
    __shared__ unsigned long long s_diag[kNmsTile];
    __shared__ int s_list[kNmsTile], s_cnt;
    const int img = blockIdx.x, tid = threadIdx.x;
    const unsigned long long* mask = mask_all + static_cast<size_t>(img) * n * col_blocks;
    const float4* boxes = boxes_all + static_cast<size_t>(img) * n;
    for (int j = tid; j < col_blocks; j += kNmsSweepThreads) removed[j] = 0ULL;
    if (keep_flags)
        for (int i = tid; i < n; i += kNmsSweepThreads) keep_flags[static_cast<size_t>(img) * n + i] = 0;
    if (kept_boxes)
        for (int i = tid; i < max_keep; i += kNmsSweepThreads) kept_boxes[static_cast<size_t>(img) * max_keep + i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();
    int kept = 0;
    for (int b = 0; b < col_blocks && kept < max_keep; ++b) {        // `kept` is replicated: the exit is uniform
        const int size = min(n - b * kNmsTile, kNmsTile);
        if (tid < size) s_diag[tid] = mask[static_cast<size_t>(b * kNmsTile + tid) * col_blocks + b];
        __syncthreads();
        if (tid == 0) {
            unsigned long long rem = removed[b];
            const int room = max_keep - kept;
            int cnt = 0;
            for (int i = 0; i < size && cnt < room; ++i)
                if (!((rem >> i) & 1ULL)) {
                    s_list[cnt++] = i;
                    rem |= s_diag[i];
                }
            s_cnt = cnt;
        }
        __syncthreads();
        const int cnt = s_cnt;
        for (int j = b + 1 + tid; j < col_blocks; j += kNmsSweepThreads) {
            unsigned long long acc = removed[j];
            const unsigned long long* col = mask + static_cast<size_t>(b) * kNmsTile * col_blocks + j;
            int t = 0;
            for (; t + 4 <= cnt; t += 4) {
                const unsigned long long m0 = col[static_cast<size_t>(s_list[t]) * col_blocks], m1 = col[static_cast<size_t>(s_list[t + 1]) * col_blocks];
                const unsigned long long m2 = col[static_cast<size_t>(s_list[t + 2]) * col_blocks], m3 = col[static_cast<size_t>(s_list[t + 3]) * col_blocks];
                acc |= (m0 | m1) | (m2 | m3);
            }
            for (; t < cnt; ++t) acc |= col[static_cast<size_t>(s_list[t]) * col_blocks];
            removed[j] = acc;
        }
        if (tid < cnt) {
            const int pos = b * kNmsTile + s_list[tid];
            if (keep_flags) keep_flags[static_cast<size_t>(img) * n + pos] = 1;       // flag by RANK (position in the sorted input)
            if (kept_boxes) kept_boxes[static_cast<size_t>(img) * max_keep + kept + tid] = boxes[pos];
        }
        kept += cnt;
        __syncthreads();
    }
    if (tid == 0) counts[img] = kept;
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

AFAN_EXPORT int64_t afan_nms_batched_workspace_bytes(int64_t images, int64_t n) {
    if (n < 0 || images < 0) return AFAN_ERR_SIZE;
    const int64_t col_blocks = (n + kNmsTile - 1) / kNmsTile;
    return images * n * col_blocks * 8 + 256;
}

// boxes_sorted [images, n, 4] ranked by descending score per image; kept_boxes [images, max_keep, 4] (nullable) receives
// the first max_keep surviving boxes of every image in rank order, zero-padded; keep_flags [images, n] (nullable) is by
// rank; counts [images] = min(survivors, max_keep).  No host synchronisation.
AFAN_EXPORT int afan_nms_batched_f32(const float* boxes_sorted, float threshold, int64_t max_keep, float* kept_boxes,
                                     uint8_t* keep_flags, int32_t* counts, void* workspace, int64_t workspace_bytes,
                                     int64_t images, int64_t n, afan_stream_t stream) {
    if (n < 0 || images < 0 || max_keep < 0) return AFAN_ERR_SIZE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (images == 0) return AFAN_OK;
    if (!counts) return AFAN_ERR_NULL;
    if (n == 0 || max_keep == 0) {
        cudaMemsetAsync(counts, 0, sizeof(int32_t) * images, st);
        if (kept_boxes && max_keep) cudaMemsetAsync(kept_boxes, 0, sizeof(float) * 4 * images * max_keep, st);
        if (keep_flags && n) cudaMemsetAsync(keep_flags, 0, images * n, st);
        return launch_status();
    }
    if (!boxes_sorted) return AFAN_ERR_NULL;
    if (!aligned16(boxes_sorted) || (kept_boxes && !aligned16(kept_boxes))) return AFAN_ERR_UNSUPPORTED;
    if (!workspace || workspace_bytes < afan_nms_batched_workspace_bytes(images, n)) return AFAN_ERR_WORKSPACE;
    const int64_t col_blocks = (n + kNmsTile - 1) / kNmsTile;
    if (col_blocks > 65535 || col_blocks * 8 > 200 * 1024 || images > 65535 || max_keep >= (int64_t(1) << 31)) return AFAN_ERR_UNSUPPORTED;
    unsigned long long* mask = static_cast<unsigned long long*>(workspace);
    dim3 grid(static_cast<unsigned int>(col_blocks), static_cast<unsigned int>(col_blocks), static_cast<unsigned int>(images));
    nms_mask_batched_kernel<<<grid, kNmsTile, 0, st>>>(reinterpret_cast<const float4*>(boxes_sorted), mask, static_cast<int>(n),
                                                      static_cast<int>(col_blocks), threshold);
    const size_t smem = static_cast<size_t>(col_blocks) * 8;
    if (smem > 40 * 1024)
        cudaFuncSetAttribute(nms_sweep_batched_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    nms_sweep_batched_kernel<<<static_cast<unsigned int>(images), kNmsSweepThreads, smem, st>>>(
        mask, reinterpret_cast<const float4*>(boxes_sorted), reinterpret_cast<float4*>(kept_boxes), keep_flags, counts,
        static_cast<int>(n), static_cast<int>(col_blocks), static_cast<int>(max_keep));
    return launch_status();
}
