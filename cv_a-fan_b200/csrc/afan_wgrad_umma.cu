// afan_wgrad_umma.cu -- weight gradient of the tail's 3x3 / stride 1 / pad 1 convolutions on tcgen05 (sm_100a).
//
// Replaces the dW part of `loss.backward()` (Classification/main_perturb.py:200) for nn.Conv2d `conv1` / `conv2` of
// BasicBlock (resnet_s.py:53,55) at (C, H) in {(32, 16), (64, 8)}:
//     dW[co][ci][ky][kx] = sum_{n, y, x} dy[n][co][y][x] * X[n][ci][y + ky - 1][x + kx - 1]          (zero padding)
// as a GEMM whose REDUCTION dimension is the pixels: per ky,
//     D_ky[(kx, ci)][co] += A[(kx, ci)][pixel] * B[co][pixel],   A = X shifted by (ky, kx),  B = dy
// with kind::tf32 and the 3xTF32 split (a_hi*b_hi + a_lo*b_hi + a_hi*b_lo), fp32 accumulators in TMEM, like the forward
// kernel (afan_conv_umma.cu).  The strict-FFMA kernel (afan_conv.cu) stays the fallback for the other shapes.
//
// Operand layout.  K-major operands want 4 consecutive K elements (pixels) per 16 bytes: NCHW gives exactly that along x.
// A row shift (ky) moves the start address by whole rows -- aligned; a column shift (kx) of one pixel is not, so the three
// kx variants are materialised while staging and become the M dimension: M = 128 rows = (kx slot, ci) for C = 32 (slot 3 is
// don't-care rows), two M = 128 instructions for C = 64.  Shared-memory tile, per staged input row r and pixel quad xq:
//     A: [r][xq][slot 0..2][ci][4 px]    (slot stride C*16 B, quad stride 3*C*16 B = LBO, SBO = 128 B = 8 channels)
//     B: [y][xq][hi | lo][co][4 px]      (quad stride 2*C*16 B)
// One K step = 8 pixels = two quads of one image row; the don't-care rows of slot 3 read the next quad's slot 0.
//
// Persistent CTAs (one per SM) walk images round-robin; 8 staging warps load float4 quads straight from global memory
// (lane = 8 channels x 4 quads: full 32-byte sectors; the kx neighbours come from warp shuffles), split to {hi, lo} TF32 and
// store the tiles (128-bit, conflict-free); one elected thread issues the MMAs; two staging stages are recycled through
// tcgen05.commit.  Every CTA leaves ONE partial dW (9*C*C floats); a fixed-order fold kernel adds them into dW: deterministic.
#include "afan_umma.cuh"

namespace afan {
namespace umma {

constexpr int kWgThreads = 256 + 32;             // 8 staging / epilogue warps + the MMA warp
constexpr uint32_t kWgIdesc64 = idesc_tf32(64), kWgIdesc32 = idesc_tf32(32);

template <int C, int H>
struct WgCfg {
    static_assert((C == 32 && H == 16) || (C == 64 && H == 8), "tail shapes");
    static constexpr int QW = H / 4;                         // pixel quads per image row
    static constexpr int RB = 4;                             // output rows per stage
    static constexpr int AR = RB + 2;                        // staged input rows (with the ky halo)
    static constexpr int NSTAGE_IMG = H / RB;                // stages per image
    static constexpr uint32_t QA = 3 * C * 16;               // bytes per (row, quad) of A: three kx slots
    static constexpr uint32_t QB = 2 * C * 16;               // bytes per (row, quad) of B: hi | lo
    static constexpr uint32_t A_TILE = AR * QW * QA + C * 16;   // + one slot: the don't-care rows of the last quad
    static constexpr uint32_t B_TILE = RB * QW * QB;
    static constexpr uint32_t STAGE = 2 * A_TILE + B_TILE;   // A_hi, A_lo, B
    static constexpr uint32_t OFF_BAR = 2 * STAGE;
    static constexpr uint32_t SMEM = OFF_BAR + 64;
    static constexpr int MH = C / 32;                        // M = 128 instructions per (ky, pass): 1 (C = 32) or 2 (C = 64)
    static constexpr int DCOLS = C == 32 ? 64 : 64;          // TMEM columns per accumulator: C = 32: [hi*hi + lo*hi | hi*lo]
    static constexpr uint32_t TMEM_COLS = C == 32 ? 256 : 512;   // 3 ky x MH accumulators x 64 columns
    static_assert(A_TILE % 16 == 0 && B_TILE % 16 == 0 && SMEM <= 227 * 1024, "shared memory");
};

template <int C, int H>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad_umma_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ partial, const int n_images) {
    using K = WgCfg<C, H>;
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t sbase = smem_u32(smem);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t bar0 = sbase + K::OFF_BAR;
    auto a_full = [&](int i) { return bar0 + 8u * i; };
    auto a_empty = [&](int i) { return bar0 + 8u * (2 + i); };
    const uint32_t acc_full = bar0 + 8u * 4;
    volatile uint32_t* tmem_slot = reinterpret_cast<volatile uint32_t*>(smem + K::OFF_BAR + 48);

    if (warp == 8) {
        if (elect_one()) {
            for (int i = 0; i < 2; ++i) { mbar_init(a_full(i), 8); mbar_init(a_empty(i), 1); }
            mbar_init(acc_full, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(const_cast<uint32_t*>(tmem_slot))), "r"(K::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
        // the spare slot behind the last quad of every A tile is read (don't-care rows) but never written: keep it finite
        for (int i = tid; i < 2 * 2 * (C * 16 / 16); i += 256) {
            const int st = i / (2 * C), rest = i - st * 2 * C, hl = rest / C, u = rest - hl * C;
            *reinterpret_cast<uint4*>(smem + st * K::STAGE + hl * K::A_TILE + K::AR * K::QW * K::QA + u * 16) = make_uint4(0u, 0u, 0u, 0u);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = *tmem_slot;
    int my_images = 0;
    for (int n = blockIdx.x; n < n_images; n += gridDim.x) ++my_images;
    const int total_stages = my_images * K::NSTAGE_IMG;

    if (warp == 8) {
        // ================= MMA issuer =================
        if (elect_one()) {
            for (int sg = 0; sg < total_stages; ++sg) {
                const int s = sg & 1;
                mbar_wait(a_full(s), (sg >> 1) & 1);
                tc_fence_after();
                const uint32_t st = sbase + s * K::STAGE;
                const uint64_t dA_hi = smem_desc(st, K::QA, 128), dA_lo = dA_hi + (K::A_TILE >> 4);
                const uint64_t dB = smem_desc(st + 2 * K::A_TILE, K::QB, 128);
#pragma unroll
                for (int r = 0; r < K::RB; ++r) {
#pragma unroll
                    for (int ks = 0; ks < K::QW / 2; ++ks) {
                        const uint32_t first = (sg | r | ks) == 0 ? 0u : 1u;       // the CTA's very first K step initialises TMEM
                        const uint64_t b_hi = dB + (((r * K::QW + 2 * ks) * K::QB) >> 4);
#pragma unroll
                        for (int ky = 0; ky < 3; ++ky) {
                            const uint32_t aoff = (((r + ky) * K::QW + 2 * ks) * K::QA) >> 4;
#pragma unroll
                            for (int mh = 0; mh < K::MH; ++mh) {
                                const uint32_t moff = aoff + ((mh * 2 * C * 16) >> 4);   // C = 64: kx slots {0, 1} / {2, don't care}
                                const uint32_t dcol = tmem + (ky * K::MH + mh) * 64;
                                if constexpr (C == 32) {
                                    // cols [0, 32): A_hi*B_hi + A_lo*B_hi;  cols [32, 64): A_hi*B_lo  ([B_hi ; B_lo] is one N = 64 operand)
                                    tc_mma_tf32(dcol, dA_hi + moff, b_hi, kWgIdesc64, first);
                                    tc_mma_tf32(dcol, dA_lo + moff, b_hi, kWgIdesc32, 1u);
                                } else {
                                    tc_mma_tf32(dcol, dA_lo + moff, b_hi, kWgIdesc64, first);
                                    tc_mma_tf32(dcol, dA_hi + moff, b_hi + ((C * 16) >> 4), kWgIdesc64, 1u);
                                    tc_mma_tf32(dcol, dA_hi + moff, b_hi, kWgIdesc64, 1u);
                                }
                            }
                        }
                    }
                }
                tc_commit(a_empty(s));
            }
            tc_commit(acc_full);
        }
        __syncwarp();
    } else {
        // ================= staging: global float4 quads -> {hi, lo} tiles =================
        // lane = (channel % 8) + 8 * (quad or row bit): 8 channels x 4 lanes of 16 B -> full 32-byte sectors, conflict-free stores
        constexpr int XQ = K::QW;                            // quads per row: 4 (H = 16) or 2 (H = 8)
        constexpr int LPR = 8 * XQ;                          // lanes per (row, 8-channel group): 32 or 16
        constexpr int RPW = 32 / LPR;                        // rows handled by one warp-load: 1 or 2
        const int cl = lane & 7, xq = (lane >> 3) % XQ, rsub = (lane >> 3) / XQ;
        constexpr int A_ITEMS = K::AR * (C / 8) / RPW;       // warp-level load items for A per stage
        constexpr int B_ITEMS = K::RB * (C / 8) / RPW;
        constexpr int A_PER_WARP = A_ITEMS / 8, B_PER_WARP = B_ITEMS / 8;
        static_assert(A_ITEMS % 8 == 0 && B_ITEMS % 8 == 0, "items divide over the 8 staging warps");
        const size_t plane = static_cast<size_t>(H) * H;
        // register double buffer: the quads of stages sg + 1 and sg + 2 are in flight while stage sg is split and stored
        float4 qa[2][A_PER_WARP], qb[2][B_PER_WARP];
        auto a_item = [&](int it, int& r, int& cg) { const int i = warp * A_PER_WARP + it; cg = i % (C / 8); r = (i / (C / 8)) * RPW + rsub; };
        auto b_item = [&](int it, int& r, int& cg) { const int i = warp * B_PER_WARP + it; cg = i % (C / 8); r = (i / (C / 8)) * RPW + rsub; };
        auto load_stage = [&](int sg, float4 (&da)[A_PER_WARP], float4 (&db)[B_PER_WARP]) {
            const int img = blockIdx.x + (sg / K::NSTAGE_IMG) * gridDim.x, y0 = (sg % K::NSTAGE_IMG) * K::RB;
            const float* xi = x + static_cast<size_t>(img) * C * plane;
            const float* di = dy + static_cast<size_t>(img) * C * plane;
#pragma unroll
            for (int it = 0; it < A_PER_WARP; ++it) {
                int r, cg;
                a_item(it, r, cg);
                const int ys = y0 - 1 + r;
                da[it] = (ys >= 0 && ys < H) ? __ldg(reinterpret_cast<const float4*>(xi + (cg * 8 + cl) * plane + ys * H + 4 * xq))
                                             : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int it = 0; it < B_PER_WARP; ++it) {
                int r, cg;
                b_item(it, r, cg);
                db[it] = __ldg(reinterpret_cast<const float4*>(di + (cg * 8 + cl) * plane + (y0 + r) * H + 4 * xq));
            }
        };
        auto split4 = [&](const float4& v, uint4& hi, uint4& lo) {
            split_tf32(v.x, hi.x, lo.x); split_tf32(v.y, hi.y, lo.y); split_tf32(v.z, hi.z, lo.z); split_tf32(v.w, hi.w, lo.w);
        };
        auto store_stage = [&](int sg, const float4 (&da)[A_PER_WARP], const float4 (&db)[B_PER_WARP]) {
            const int s = sg & 1;
            if (sg >= 2) mbar_wait(a_empty(s), ((sg >> 1) - 1) & 1);
            unsigned char* st = smem + s * K::STAGE;
#pragma unroll
            for (int it = 0; it < A_PER_WARP; ++it) {
                int r, cg;
                a_item(it, r, cg);
                const float4 q = da[it];
                // neighbours along x from the lanes holding the adjacent quads of the same (row, channel)
                const float lw = __shfl_up_sync(0xffffffffu, q.w, 8), rx = __shfl_down_sync(0xffffffffu, q.x, 8);
                const float left = xq > 0 ? lw : 0.f, right = xq < XQ - 1 ? rx : 0.f;
                const float4 s0 = make_float4(left, q.x, q.y, q.z), s2 = make_float4(q.y, q.z, q.w, right);
                const uint32_t base = (r * K::QW + xq) * K::QA + (cg * 8 + cl) * 16;
                uint4 hi, lo;
                split4(s0, hi, lo);
                *reinterpret_cast<uint4*>(st + base) = hi; *reinterpret_cast<uint4*>(st + K::A_TILE + base) = lo;
                split4(q, hi, lo);
                *reinterpret_cast<uint4*>(st + base + C * 16) = hi; *reinterpret_cast<uint4*>(st + K::A_TILE + base + C * 16) = lo;
                split4(s2, hi, lo);
                *reinterpret_cast<uint4*>(st + base + 2 * C * 16) = hi; *reinterpret_cast<uint4*>(st + K::A_TILE + base + 2 * C * 16) = lo;
            }
#pragma unroll
            for (int it = 0; it < B_PER_WARP; ++it) {
                int r, cg;
                b_item(it, r, cg);
                uint4 hi, lo;
                split4(db[it], hi, lo);
                const uint32_t base = 2 * K::A_TILE + (r * K::QW + xq) * K::QB + (cg * 8 + cl) * 16;
                *reinterpret_cast<uint4*>(st + base) = hi;
                *reinterpret_cast<uint4*>(st + base + C * 16) = lo;
            }
        };
        auto publish = [&](int sg) {
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(a_full(sg & 1));
        };
        if (total_stages > 0) load_stage(0, qa[0], qb[0]);
        if (total_stages > 1) load_stage(1, qa[1], qb[1]);
        for (int sg = 0; sg < total_stages; sg += 2) {                  // unrolled by two so that the register sets are static
            store_stage(sg, qa[0], qb[0]);
            if (sg + 2 < total_stages) load_stage(sg + 2, qa[0], qb[0]);
            publish(sg);
            if (sg + 1 < total_stages) {
                store_stage(sg + 1, qa[1], qb[1]);
                if (sg + 3 < total_stages) load_stage(sg + 3, qa[1], qb[1]);
                publish(sg + 1);
            }
        }
        // ================= epilogue: this CTA's partial dW =================
        if (total_stages > 0) {
            mbar_wait(acc_full, 0);
            tc_fence_after();
        }
        // TMEM lane m = (kx slot, ci) within an M = 128 accumulator; warps 0-3 read accumulators of ky 0 / 1, warps 4-7 of ky 2 + ...
        const int wq = warp & 3;
        const uint32_t lane_base = static_cast<uint32_t>(32 * wq) << 16;
        float* out = partial + static_cast<size_t>(blockIdx.x) * 9 * C * C;
        constexpr int NACC = 3 * K::MH;
        for (int acc = warp >> 2; acc < NACC; acc += 2) {
            const int ky = acc / K::MH, mh = acc % K::MH;
            const int m = 32 * wq + lane;                                       // row of the accumulator
            const int slot = (C == 32) ? m / 32 : 2 * mh + m / 64, ci = (C == 32) ? m % 32 : m % 64;
            if (slot < 3) {
                float* dst = out + ((static_cast<size_t>(ky) * 3 + slot) * C + ci) * C;          // [ky][kx][ci][co]
                if (total_stages == 0) {
                    for (int co = 0; co < C; ++co) dst[co] = 0.f;
                } else if constexpr (C == 32) {
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        float v[16], v2[16];
                        tc_ld16(tmem + lane_base + acc * 64 + half * 16, v);
                        tc_ld16(tmem + lane_base + acc * 64 + 32 + half * 16, v2);
#pragma unroll
                        for (int i = 0; i < 16; ++i) dst[half * 16 + i] = __fadd_rn(v[i], v2[i]);
                    }
                } else {
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        float v[16];
                        tc_ld16(tmem + lane_base + acc * 64 + q4 * 16, v);
#pragma unroll
                        for (int i = 0; i < 16; ++i) dst[q4 * 16 + i] = v[i];
                    }
                }
            } else if (total_stages > 0) {                    // tcgen05.ld is warp-collective: slot-3 lanes take part in it
                float v[16];
                if constexpr (C == 32) {
                    for (int half = 0; half < 2; ++half) { tc_ld16(tmem + lane_base + acc * 64 + half * 16, v); tc_ld16(tmem + lane_base + acc * 64 + 32 + half * 16, v); }
                } else {
                    for (int q4 = 0; q4 < 4; ++q4) tc_ld16(tmem + lane_base + acc * 64 + q4 * 16, v);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(K::TMEM_COLS) : "memory");
    }
}

// dW[co][ci][ky][kx] (+)= sum over CTAs of partial[cta][ky][kx][ci][co], CTA order fixed: deterministic
template <int C>
__global__ void __launch_bounds__(256) wgrad_umma_fold_kernel(const float* __restrict__ partial, float* __restrict__ dw, int n_ctas,
                                                              int accumulate) {
    __shared__ float red[8][32];
    const int o = blockIdx.x * 32 + threadIdx.x;            // index into [ky][kx][ci][co] (co fastest): coalesced partial reads
    float s = 0.f;
#pragma unroll 4
    for (int b = threadIdx.y; b < n_ctas; b += 8) s += partial[static_cast<size_t>(b) * 9 * C * C + o];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0) {
#pragma unroll
        for (int i = 1; i < 8; ++i) s += red[i][threadIdx.x];
        const int co = o % C, ci = (o / C) % C, t = o / (C * C);                 // t = ky * 3 + kx
        float* dst = dw + (static_cast<size_t>(co) * C + ci) * 9 + t;
        *dst = accumulate ? *dst + s : s;
    }
}

template <int C, int H>
static int launch_wgrad_umma(const float* x, const float* dy, float* dw, float* ws, int64_t ws_bytes, int n, int accumulate, cudaStream_t st) {
    using K = WgCfg<C, H>;
    static bool configured = false;      // benign race: idempotent
    if (!configured) {
        if (cudaFuncSetAttribute(wgrad_umma_kernel<C, H>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM) != cudaSuccess)
            return (cudaGetLastError(), AFAN_ERR_LAUNCH);
        configured = true;
    }
    int grid = sm_count();
    if (grid > n) grid = n;
    if (static_cast<int64_t>(grid) * 9 * C * C * 4 > ws_bytes) return AFAN_ERR_WORKSPACE;
    wgrad_umma_kernel<C, H><<<grid, kWgThreads, K::SMEM, st>>>(x, dy, ws, n);
    if (launch_status() != AFAN_OK) return AFAN_ERR_LAUNCH;
    wgrad_umma_fold_kernel<C><<<9 * C * C / 32, dim3(32, 8), 0, st>>>(ws, dw, grid, accumulate);
    return launch_status();
}

}  // namespace umma
}  // namespace afan

using namespace afan;

AFAN_EXPORT int afan_conv3x3_wgrad_umma_supported(int64_t n, int64_t c, int64_t hw) {
    return (n > 0 && n < (1 << 30) && ((c == 32 && hw == 16) || (c == 64 && hw == 8))) ? 1 : 0;
}

AFAN_EXPORT int afan_conv3x3_wgrad_umma_f32(const float* x, const float* dy, float* dw, void* workspace, int64_t workspace_bytes,
                                            int64_t n, int64_t c, int64_t hw, int accumulate, afan_stream_t stream) {
    if (n <= 0 || c <= 0 || hw <= 0) return AFAN_ERR_SIZE;
    if (!x || !dy || !dw) return AFAN_ERR_NULL;
    if (!workspace) return AFAN_ERR_WORKSPACE;
    if (!aligned16(x) || !aligned16(dy) || !aligned16(workspace) || !afan_conv3x3_wgrad_umma_supported(n, c, hw)) return AFAN_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* ws = static_cast<float*>(workspace);
    if (c == 32) return umma::launch_wgrad_umma<32, 16>(x, dy, dw, ws, workspace_bytes, static_cast<int>(n), accumulate, st);
    return umma::launch_wgrad_umma<64, 8>(x, dy, dw, ws, workspace_bytes, static_cast<int>(n), accumulate, st);
}
