// afan_conv.cu -- strict-fp32 3x3 / stride 1 / pad 1 convolutions of the tail sub-network that the PGD ascent
// re-executes `steps` times per batch (Classification/resnet_s.py:53,55 `conv1` / `conv2` of BasicBlock, driven by
// attack_algo.py:49-52 `model(x_adv, start_point=k)` + `autograd.grad`).  cuDNN serves these tiny-channel shapes
// (16/32/64 channels on 32x32 / 16x16 / 8x8 maps) with its generic `implicit_convolve_sgemm` / `dgrad_engine` pair at
// ~12-27 TFLOP/s plus a zero-fill launch per dgrad; a direct convolution that keeps a whole image (or half of one)
// in shared memory runs the same FFMA work at the pipe rate.
//
//   forward  y[n,co,h,w] = sum_{ci,kh,kw} x[n,ci,h+kh-1,w+kw-1] * W[co,ci,kh,kw]
//   dgrad    dx[n,ci,h,w] = sum_{co,kh,kw} dy[n,co,h-kh+1,w-kw+1] * W[co,ci,kh,kw]  == forward with the packed
//            weights W'[k][t'][j] = W[k][j][8-t']  (same kernel, other packing)
//   wgrad    dW[co,ci,kh,kw] = sum_{n,h,w} dy[n,co,h,w] * x[n,ci,h+kh-1,w+kw-1]
//
// Data layout: NCHW fp32 (the reference's).  Weights are repacked once per optimiser step (one launch for all layers)
// into [reduction channel][tap][output channel] so that a K-chunk of weights is one contiguous cp.async stream and a
// thread's TQ output channels are one 128-bit shared-memory load.
//
// Kernel shape (forward / dgrad): CTA = R rows x full width of ONE image x all output channels.  The reduction
// channels arrive in chunks of CK through a 2-stage cp.async pipeline (input rows incl. halo + that chunk's weights).
// Thread tile = 2 rows x 4 columns x TQ output channels: per reduction channel 4 x (LDS.32 + LDS.128 + LDS.32) input
// loads and 9 LDS.128 weight loads feed 72*TQ FFMAs.  Lanes 0-7 of a warp walk 8 neighbouring pixel blocks, lanes / 8
// walk 4 channel groups: input loads hit 8 distinct conflict-free 16-byte words (row pitch chosen per width), weight
// loads 4.  Accumulation order is fixed (ci ascending, taps row-major): results are deterministic.
#include "afan_common.cuh"

namespace afan {

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int C_, int HW_, int R_, int TQ_, int CK_>
struct ConvCfg {
    static constexpr int C = C_, H = HW_, W = HW_, R = R_, TQ = TQ_, CK = CK_;
    // shared-memory row pitch: interior starts at column 4 (16-byte aligned), halo columns 3 and W+4.  The pitch makes
    // the 8 pixel blocks of a quarter-warp (W/4 per row, rows 2 apart) land on distinct banks.
    static constexpr int S = (W == 32) ? 40 : (W == 16 ? 24 : 20);
    static constexpr int PBW = W / 4, PBH = R / 2, NPB = PBW * PBH, NCG = C / TQ;
    static constexpr int THREADS = NPB * NCG;
    static constexpr int ROWS = R + 2;
    static constexpr int IN_ELEMS = CK * ROWS * S;
    static constexpr int W_ELEMS = CK * 9 * C;
    static constexpr int STAGE = IN_ELEMS + W_ELEMS;
    static constexpr int SMEM_BYTES = 2 * STAGE * 4;
    static constexpr int NCHUNK = C / CK;
    static constexpr int TPI = H / R;      // CTAs (row tiles) per image
    static_assert(NPB % 8 == 0 && NCG % 4 == 0, "warp = 8 pixel blocks x 4 channel groups");
    static_assert(THREADS <= 1024 && THREADS % 32 == 0, "CTA size");
    static_assert(C % CK == 0 && H % R == 0 && R % 2 == 0 && (TQ == 2 || TQ == 4 || TQ == 8), "tiling");
};

template <class K>
__global__ void __launch_bounds__(K::THREADS)
conv3x3_kernel(const float* __restrict__ x, const float* __restrict__ wp, float* __restrict__ y, const float* __restrict__ addend) {
    constexpr int C = K::C, H = K::H, W = K::W, R = K::R, TQ = K::TQ, CK = K::CK, S = K::S;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x / K::TPI, r0 = (blockIdx.x % K::TPI) * R;
    constexpr int PG = K::NPB / 8;
    const int pix = (warp % PG) * 8 + (lane & 7);
    const int cg = (warp / PG) * 4 + (lane >> 3);
    const int pby = pix / K::PBW, pbx = pix % K::PBW;

    // zero both stages once: rows / columns outside the image are never written by the pipeline
    for (int i = tid; i < 2 * K::STAGE / 4; i += K::THREADS)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    auto load_chunk = [&](int chunk, int buf) {
        float* in_s = smem + buf * K::STAGE;
        float* w_s = in_s + K::IN_ELEMS;
        const int ci0 = chunk * CK;
        constexpr int SEGS = W / 4, NIN = CK * K::ROWS * SEGS;
        for (int i = tid; i < NIN; i += K::THREADS) {
            const int seg = i % SEGS, row = (i / SEGS) % K::ROWS, cil = i / (SEGS * K::ROWS);
            const int gr = r0 - 1 + row;
            if (gr >= 0 && gr < H)
                cp_async16(in_s + (cil * K::ROWS + row) * S + 4 + 4 * seg,
                           x + (static_cast<size_t>(n * C + ci0 + cil) * H + gr) * W + 4 * seg);
        }
        const float* wsrc = wp + static_cast<size_t>(ci0) * 9 * C;
        for (int i = tid; i < K::W_ELEMS / 4; i += K::THREADS) cp_async16(w_s + 4 * i, wsrc + 4 * i);
        cp_async_commit();
    };

    float acc[TQ][2][4];
#pragma unroll
    for (int q = 0; q < TQ; ++q)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[q][r][c] = 0.f;

    load_chunk(0, 0);
#pragma unroll 1
    for (int ch = 0; ch < K::NCHUNK; ++ch) {
        cp_async_wait_all();
        __syncthreads();               // chunk ch has landed for everyone; everyone is done with chunk ch-1's buffer
        if (ch + 1 < K::NCHUNK) load_chunk(ch + 1, (ch + 1) & 1);
        const float* in_s = smem + (ch & 1) * K::STAGE + (2 * pby) * S + 4 * pbx + 3;
        const float* w_s = smem + (ch & 1) * K::STAGE + K::IN_ELEMS + cg * TQ;
#pragma unroll
        for (int cil = 0; cil < CK; ++cil) {
            float v[4][6];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const float* p = in_s + (cil * K::ROWS + rr) * S;
                v[rr][0] = p[0];
                const float4 m = *reinterpret_cast<const float4*>(p + 1);
                v[rr][1] = m.x; v[rr][2] = m.y; v[rr][3] = m.z; v[rr][4] = m.w;
                v[rr][5] = p[5];
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int kh = t / 3, kw = t % 3;
                float wv[TQ];
                const float* wq = w_s + (cil * 9 + t) * C;
                if constexpr (TQ == 2) {
                    const float2 a = *reinterpret_cast<const float2*>(wq);
                    wv[0] = a.x; wv[1] = a.y;
                } else {
#pragma unroll
                    for (int q4 = 0; q4 < TQ / 4; ++q4) {
                        const float4 a = *reinterpret_cast<const float4*>(wq + 4 * q4);
                        wv[4 * q4] = a.x; wv[4 * q4 + 1] = a.y; wv[4 * q4 + 2] = a.z; wv[4 * q4 + 3] = a.w;
                    }
                }
#pragma unroll
                for (int q = 0; q < TQ; ++q)
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[q][r][c] = fmaf(wv[q], v[r + kh][c + kw], acc[q][r][c]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < TQ; ++q)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            const size_t off = (static_cast<size_t>(n * C + cg * TQ + q) * H + r0 + 2 * pby + r) * W + 4 * pbx;
            float4 o = make_float4(acc[q][r][0], acc[q][r][1], acc[q][r][2], acc[q][r][3]);
            if (addend != nullptr) {       // y = conv(x) + addend: the residual branch's gradient joins the dgrad here
                const float4 a = *reinterpret_cast<const float4*>(addend + off);
                o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w;
            }
            *reinterpret_cast<float4*>(y + off) = o;
        }
}

// W[co][ci][3][3] -> forward packing wf[ci][t][co] and dgrad packing wd[co][8-t][ci]; blockIdx.y = layer.
struct ConvPackDesc {
    const float* w;
    float* wf;
    float* wd;
    long long c;
};

// c < 0 marks a stride-2 stage transition with CIN = -c, COUT = 2 * CIN (W[COUT][CIN][3][3]): forward packing
// wf[ci][tap][co], input-gradient packing wd[co][tap][ci] (taps NOT flipped: the stride-2 dgrad kernel indexes them itself).
__device__ __forceinline__ void pack_stride2(const ConvPackDesc& d, int start, int stride) {
    const int CIN = static_cast<int>(-d.c), COUT = 2 * CIN, total = COUT * CIN * 9;
    for (int idx = start; idx < total; idx += stride) {
        const int co = idx / (CIN * 9), ci = (idx / 9) % CIN, t = idx % 9;
        const float v = d.w[idx];
        d.wf[(ci * 9 + t) * COUT + co] = v;
        d.wd[(co * 9 + t) * CIN + ci] = v;
    }
}

__global__ void __launch_bounds__(kThreads) conv3x3_pack_kernel(const ConvPackDesc* __restrict__ descs) {
    const ConvPackDesc d = descs[blockIdx.y];
    if (d.c < 0) { pack_stride2(d, blockIdx.x * kThreads + threadIdx.x, gridDim.x * kThreads); return; }
    const int C = static_cast<int>(d.c), total = C * C * 9;
    for (int idx = blockIdx.x * kThreads + threadIdx.x; idx < total; idx += gridDim.x * kThreads) {
        const int co = idx / (C * 9), ci = (idx / 9) % C, t = idx % 9;
        const float v = d.w[idx];
        d.wf[(ci * 9 + t) * C + co] = v;
        d.wd[(co * 9 + (8 - t)) * C + ci] = v;
    }
}

template <class K>
static int launch_conv(const float* x, const float* wp, float* y, const float* addend, int n, cudaStream_t st) {
    static bool attr_set = false;      // benign race: idempotent
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv3x3_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES) != cudaSuccess) {
            cudaGetLastError();
            return AFAN_ERR_LAUNCH;
        }
        attr_set = true;
    }
    // plain launches: with the PDL attribute the early-scheduled CTAs of these shared-memory-heavy kernels serialise
    // against their neighbours inside the captured step (measured: 27.4 ms/step with, 15.8 ms without)
    conv3x3_kernel<K><<<static_cast<unsigned>(n * K::TPI), K::THREADS, K::SMEM_BYTES, st>>>(x, wp, y, addend);
    return launch_status();
}


// ---- stride-2 stage transitions (resnet_s.py:98 `stride=2` first block of stages 2 and 3): CIN -> 2*CIN, H -> H/2 --------------
// Forward: same CTA shape as the stride-1 kernel (one image, 2x4 output pixels x TQ channels per thread); the thread's
// window is 5 input rows x 9 input columns per reduction channel (LDS.32 + 2 LDS.128 per row).
template <int CIN_, int HO_, int TQ_, int CK_, int LP_ = 8>
struct S2Cfg {
    static constexpr int CIN = CIN_, COUT = 2 * CIN_, HO = HO_, WO = HO_, HI = 2 * HO_, WI = 2 * HO_, TQ = TQ_, CK = CK_;
    static constexpr int LP = LP_;                         // lanes of a warp that walk pixel blocks (the other 32/LP walk channel groups)
    // interior from column 4, left halo column 3; for 16-wide inputs the pitch 28 puts the 2x2 pixel blocks of a
    // quarter-warp (8 floats apart in x, 4 rows apart in y) on distinct banks
    static constexpr int S = (WI == 16) ? 28 : WI + 8;
    static constexpr int ROWS = HI + 1;                    // row 0 = zero halo (global row -1)
    static constexpr int PBW = WO / 4, PBH = HO / 2, NPB = PBW * PBH, NCG = COUT / TQ, THREADS = NPB * NCG;
    static constexpr int IN_ELEMS = CK * ROWS * S, W_ELEMS = CK * 9 * COUT, STAGE = IN_ELEMS + W_ELEMS;
    static constexpr int SMEM_BYTES = 2 * STAGE * 4, NCHUNK = CIN / CK;
    static_assert(NPB % LP == 0 && NCG % (32 / LP) == 0 && THREADS <= 1024 && CIN % CK == 0 && IN_ELEMS % 4 == 0, "tiling");
};

template <class K>
__global__ void __launch_bounds__(K::THREADS)
conv3x3s2_kernel(const float* __restrict__ x, const float* __restrict__ wp, float* __restrict__ y) {
    constexpr int CIN = K::CIN, COUT = K::COUT, HO = K::HO, WO = K::WO, HI = K::HI, WI = K::WI, TQ = K::TQ, CK = K::CK, S = K::S;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x;
    constexpr int PG = K::NPB / K::LP;
    const int pix = (warp % PG) * K::LP + (lane % K::LP);
    const int cg = (warp / PG) * (32 / K::LP) + (lane / K::LP);
    const int pby = pix / K::PBW, pbx = pix % K::PBW;

    for (int i = tid; i < 2 * K::STAGE / 4; i += K::THREADS)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    auto load_chunk = [&](int chunk, int buf) {
        float* in_s = smem + buf * K::STAGE;
        float* w_s = in_s + K::IN_ELEMS;
        const int ci0 = chunk * CK;
        constexpr int SEGS = WI / 4, NIN = CK * HI * SEGS;
        for (int i = tid; i < NIN; i += K::THREADS) {
            const int seg = i % SEGS, row = (i / SEGS) % HI, cil = i / (SEGS * HI);
            cp_async16(in_s + (cil * K::ROWS + row + 1) * S + 4 + 4 * seg,
                       x + (static_cast<size_t>(n * CIN + ci0 + cil) * HI + row) * WI + 4 * seg);
        }
        const float* wsrc = wp + static_cast<size_t>(ci0) * 9 * COUT;
        for (int i = tid; i < K::W_ELEMS / 4; i += K::THREADS) cp_async16(w_s + 4 * i, wsrc + 4 * i);
        cp_async_commit();
    };

    float acc[TQ][2][4];
#pragma unroll
    for (int q = 0; q < TQ; ++q)
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) acc[q][r][c] = 0.f;

    load_chunk(0, 0);
#pragma unroll 1
    for (int ch = 0; ch < K::NCHUNK; ++ch) {
        cp_async_wait_all();
        __syncthreads();
        if (ch + 1 < K::NCHUNK) load_chunk(ch + 1, (ch + 1) & 1);
        const float* in_s = smem + (ch & 1) * K::STAGE + (4 * pby) * S + 8 * pbx + 3;
        const float* w_s = smem + (ch & 1) * K::STAGE + K::IN_ELEMS + cg * TQ;
#pragma unroll
        for (int cil = 0; cil < CK; ++cil) {
            float v[5][9];
#pragma unroll
            for (int rr = 0; rr < 5; ++rr) {
                const float* p = in_s + (cil * K::ROWS + rr) * S;
                v[rr][0] = p[0];
                const float4 a = *reinterpret_cast<const float4*>(p + 1);
                const float4 b = *reinterpret_cast<const float4*>(p + 5);
                v[rr][1] = a.x; v[rr][2] = a.y; v[rr][3] = a.z; v[rr][4] = a.w;
                v[rr][5] = b.x; v[rr][6] = b.y; v[rr][7] = b.z; v[rr][8] = b.w;
            }
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const int kh = t / 3, kw = t % 3;
                float wv[TQ];
                const float* wq = w_s + (cil * 9 + t) * COUT;
                if constexpr (TQ == 2) {
                    const float2 a = *reinterpret_cast<const float2*>(wq);
                    wv[0] = a.x; wv[1] = a.y;
                } else {
#pragma unroll
                    for (int q4 = 0; q4 < TQ / 4; ++q4) {
                        const float4 a = *reinterpret_cast<const float4*>(wq + 4 * q4);
                        wv[4 * q4] = a.x; wv[4 * q4 + 1] = a.y; wv[4 * q4 + 2] = a.z; wv[4 * q4 + 3] = a.w;
                    }
                }
#pragma unroll
                for (int q = 0; q < TQ; ++q)
#pragma unroll
                    for (int r = 0; r < 2; ++r)
#pragma unroll
                        for (int c = 0; c < 4; ++c) acc[q][r][c] = fmaf(wv[q], v[2 * r + kh][2 * c + kw], acc[q][r][c]);
            }
        }
    }
#pragma unroll
    for (int q = 0; q < TQ; ++q)
#pragma unroll
        for (int r = 0; r < 2; ++r) {
            float* dst = y + (static_cast<size_t>(n * COUT + cg * TQ + q) * HO + 2 * pby + r) * WO + 4 * pbx;
            *reinterpret_cast<float4*>(dst) = make_float4(acc[q][r][0], acc[q][r][1], acc[q][r][2], acc[q][r][3]);
        }
}

// Input gradient of the stride-2 convolution: dx[ci][hi][wi] = sum_{co, kh, kw : parity} dy[co][(hi+1-kh)/2][(wi+1-kw)/2] * W.
// A thread owns 2x4 "cells" (one cell = a dy pixel and the 2x2 dx pixels it is centred on) x TQ input channels:
// per reduction channel co it reads a 3x5 dy window and the 9 taps and issues 9 FMAs per cell and channel
// (1 + 2 + 2 + 4 taps for the four pixel parities).
template <int CIN_, int HO_, int TQ_, int CK_>
struct S2DgradCfg {
    static constexpr int CIN = CIN_, COUT = 2 * CIN_, HO = HO_, WO = HO_, HI = 2 * HO_, WI = 2 * HO_, TQ = TQ_, CK = CK_;
    static constexpr int S = WO + 4;                       // interior from column 0, zero halo column WO
    static constexpr int ROWS = HO + 1;                    // zero halo row HO
    static constexpr int PBW = WO / 4, PBH = HO / 2, NPB = PBW * PBH, NCG = CIN / TQ, THREADS = NPB * NCG;
    static constexpr int IN_ELEMS = CK * ROWS * S, W_ELEMS = CK * 9 * CIN, STAGE = IN_ELEMS + W_ELEMS;
    static constexpr int SMEM_BYTES = 2 * STAGE * 4, NCHUNK = COUT / CK;
    static_assert(NPB % 8 == 0 && NCG % 4 == 0 && THREADS <= 1024 && COUT % CK == 0 && IN_ELEMS % 4 == 0 && TQ == 2, "tiling");
};

template <class K>
__global__ void __launch_bounds__(K::THREADS)
conv3x3s2_dgrad_kernel(const float* __restrict__ dy, const float* __restrict__ wp, float* __restrict__ dx) {
    constexpr int CIN = K::CIN, COUT = K::COUT, HO = K::HO, WO = K::WO, HI = K::HI, WI = K::WI, TQ = K::TQ, CK = K::CK, S = K::S;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int n = blockIdx.x;
    constexpr int PG = K::NPB / 8;
    const int pix = (warp % PG) * 8 + (lane & 7);
    const int cg = (warp / PG) * 4 + (lane >> 3);
    const int pby = pix / K::PBW, pbx = pix % K::PBW;

    for (int i = tid; i < 2 * K::STAGE / 4; i += K::THREADS)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    auto load_chunk = [&](int chunk, int buf) {
        float* in_s = smem + buf * K::STAGE;
        float* w_s = in_s + K::IN_ELEMS;
        const int co0 = chunk * CK;
        constexpr int SEGS = WO / 4, NIN = CK * HO * SEGS;
        for (int i = tid; i < NIN; i += K::THREADS) {
            const int seg = i % SEGS, row = (i / SEGS) % HO, col = i / (SEGS * HO);
            cp_async16(in_s + (col * K::ROWS + row) * S + 4 * seg,
                       dy + (static_cast<size_t>(n * COUT + co0 + col) * HO + row) * WO + 4 * seg);
        }
        const float* wsrc = wp + static_cast<size_t>(co0) * 9 * CIN;
        for (int i = tid; i < K::W_ELEMS / 4; i += K::THREADS) cp_async16(w_s + 4 * i, wsrc + 4 * i);
        cp_async_commit();
    };

    float acc[TQ][4][8];                                   // dx rows 4*pby + 0..3, columns 8*pbx + 0..7
#pragma unroll
    for (int q = 0; q < TQ; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 8; ++c) acc[q][r][c] = 0.f;

    load_chunk(0, 0);
#pragma unroll 1
    for (int ch = 0; ch < K::NCHUNK; ++ch) {
        cp_async_wait_all();
        __syncthreads();
        if (ch + 1 < K::NCHUNK) load_chunk(ch + 1, (ch + 1) & 1);
        const float* in_s = smem + (ch & 1) * K::STAGE + (2 * pby) * S + 4 * pbx;
        const float* w_s = smem + (ch & 1) * K::STAGE + K::IN_ELEMS + cg * TQ;
#pragma unroll
        for (int col = 0; col < CK; ++col) {
            float d[3][5];
#pragma unroll
            for (int rr = 0; rr < 3; ++rr) {
                const float* p = in_s + (col * K::ROWS + rr) * S;
                const float4 a = *reinterpret_cast<const float4*>(p);
                d[rr][0] = a.x; d[rr][1] = a.y; d[rr][2] = a.z; d[rr][3] = a.w;
                d[rr][4] = p[4];
            }
            float w[9][TQ];
#pragma unroll
            for (int t = 0; t < 9; ++t) {
                const float2 a = *reinterpret_cast<const float2*>(w_s + (col * 9 + t) * CIN);
                w[t][0] = a.x; w[t][1] = a.y;
            }
#pragma unroll
            for (int q = 0; q < TQ; ++q)
#pragma unroll
                for (int a = 0; a < 2; ++a)
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const float d00 = d[a][b], d01 = d[a][b + 1], d10 = d[a + 1][b], d11 = d[a + 1][b + 1];
                        float& o00 = acc[q][2 * a][2 * b];
                        float& o01 = acc[q][2 * a][2 * b + 1];
                        float& o10 = acc[q][2 * a + 1][2 * b];
                        float& o11 = acc[q][2 * a + 1][2 * b + 1];
                        o00 = fmaf(d00, w[4][q], o00);                              // (kh, kw) = (1, 1)
                        o01 = fmaf(d01, w[3][q], o01);                              // (1, 0)
                        o01 = fmaf(d00, w[5][q], o01);                              // (1, 2)
                        o10 = fmaf(d10, w[1][q], o10);                              // (0, 1)
                        o10 = fmaf(d00, w[7][q], o10);                              // (2, 1)
                        o11 = fmaf(d11, w[0][q], o11);                              // (0, 0)
                        o11 = fmaf(d10, w[2][q], o11);                              // (0, 2)
                        o11 = fmaf(d01, w[6][q], o11);                              // (2, 0)
                        o11 = fmaf(d00, w[8][q], o11);                              // (2, 2)
                    }
        }
    }
#pragma unroll
    for (int q = 0; q < TQ; ++q)
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            float* dst = dx + (static_cast<size_t>(n * CIN + cg * TQ + q) * HI + 4 * pby + r) * WI + 8 * pbx;
            reinterpret_cast<float4*>(dst)[0] = make_float4(acc[q][r][0], acc[q][r][1], acc[q][r][2], acc[q][r][3]);
            reinterpret_cast<float4*>(dst)[1] = make_float4(acc[q][r][4], acc[q][r][5], acc[q][r][6], acc[q][r][7]);
        }
}

template <class K, class KernelT>
static int launch_s2(KernelT kernel, const float* in, const float* wp, float* out, int n, cudaStream_t st) {
    static bool attr_set = false;      // per (K, kernel) instantiation; benign race: idempotent
    if (!attr_set) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES) != cudaSuccess) {
            cudaGetLastError();
            return AFAN_ERR_LAUNCH;
        }
        attr_set = true;
    }
    kernel<<<static_cast<unsigned>(n), K::THREADS, K::SMEM_BYTES, st>>>(in, wp, out);
    return launch_status();
}

// Weight gradient of the stride-2 transition: dW[co][ci][kh][kw] = sum_{n,h,w} dy[n,co,h,w] * x[n,ci,2h+kh-1,2w+kw-1].
// Same scheme as conv3x3_wgrad_kernel (persistent CTAs over (image, RB dy-row band) units, 4 co x 2 ci x 9 taps per
// thread, 4 dy pixels per step); the x window of a step is 9 columns wide (LDS.32 + 2 LDS.128 per channel and tap row).
template <int CIN_, int HO_, int RB_, int PS_>
struct S2WgradCfg {
    static constexpr int CIN = CIN_, COUT = 2 * CIN_, HO = HO_, WO = HO_, HI = 2 * HO_, WI = 2 * HO_, RB = RB_, PS = PS_;
    static constexpr int S = WI + 8;                       // interior from column 4, left halo column 3
    static constexpr int NCG = COUT / 4, NCP = CIN / 2, TC = NCG * NCP, THREADS = TC * PS;
    static constexpr int XROWS = 2 * RB + 1;
    static constexpr int PLANE = ((XROWS * S + 23) / 32) * 32 + 8;
    static constexpr int X_ELEMS = CIN * PLANE, DY_ELEMS = RB * WO * COUT, STAGE = X_ELEMS + DY_ELEMS;
    static constexpr int RED_ELEMS = (PS - 1) * 72 * TC;
    static constexpr int SMEM_ELEMS = (2 * STAGE > RED_ELEMS) ? 2 * STAGE : RED_ELEMS;
    static constexpr int SMEM_BYTES = SMEM_ELEMS * 4;
    static constexpr int UPI = HO / RB;
    static_assert(PLANE >= XROWS * S && RB % PS == 0 && HO % RB == 0 && THREADS % 32 == 0 && THREADS <= 512, "tiling");
};

template <class K>
__global__ void __launch_bounds__(K::THREADS)
conv3x3s2_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ partial, int n_images) {
    constexpr int CIN = K::CIN, COUT = K::COUT, HO = K::HO, WO = K::WO, HI = K::HI, WI = K::WI, RB = K::RB, S = K::S;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int cg = tid % K::NCG, cp = (tid / K::NCG) % K::NCP, slice = tid / K::TC;
    const int units = n_images * K::UPI;

    for (int i = tid; i < 2 * K::STAGE / 4; i += K::THREADS)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    auto load_unit = [&](int u, int buf) {
        float* x_s = smem + buf * K::STAGE;
        float* dy_s = x_s + K::X_ELEMS;
        const int n = u / K::UPI, r0 = (u % K::UPI) * RB;
        constexpr int XSEGS = WI / 4, NX = CIN * K::XROWS * XSEGS, DSEGS = WO / 4, ND = COUT * RB * DSEGS;
        for (int i = tid; i < NX; i += K::THREADS) {
            const int seg = i % XSEGS, row = (i / XSEGS) % K::XROWS, ci = i / (XSEGS * K::XROWS);
            const int gr = 2 * r0 - 1 + row;
            float* dst = x_s + ci * K::PLANE + row * S + 4 + 4 * seg;
            if (gr >= 0)                                   // gr <= HI - 1 always
                cp_async16(dst, x + (static_cast<size_t>(n * CIN + ci) * HI + gr) * WI + 4 * seg);
            else
                *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
        for (int i = tid; i < ND; i += K::THREADS) {
            const int seg = i % DSEGS, row = (i / DSEGS) % RB, co = i / (DSEGS * RB);
            cp_async16(dy_s + ((row * DSEGS + seg) * COUT + co) * 4,
                       dy + (static_cast<size_t>(n * COUT + co) * HO + r0 + row) * WO + 4 * seg);
        }
        cp_async_commit();
    };

    float acc[4][2][9];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int t = 0; t < 9; ++t) acc[q][j][t] = 0.f;

    int u = blockIdx.x, buf = 0;
    if (u < units) load_unit(u, 0);
#pragma unroll 1
    for (; u < units; u += gridDim.x, buf ^= 1) {
        cp_async_wait_all();
        __syncthreads();
        if (u + static_cast<int>(gridDim.x) < units) load_unit(u + gridDim.x, buf ^ 1);
        const float* x_s = smem + buf * K::STAGE + (2 * cp) * K::PLANE + 3;
        const float* dy_s = smem + buf * K::STAGE + K::X_ELEMS + cg * 4;
#pragma unroll 1
        for (int r = slice; r < RB; r += K::PS) {
#pragma unroll 1
            for (int wq = 0; wq < WO / 4; ++wq) {
                float4 d[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    d[q] = *reinterpret_cast<const float4*>(dy_s + ((r * (WO / 4) + wq) * COUT + K::NCG * q) * 4);
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const float* p = x_s + j * K::PLANE + (2 * r + kh) * S + 8 * wq;
                        float xv[9];
                        xv[0] = p[0];
                        const float4 a = *reinterpret_cast<const float4*>(p + 1);
                        const float4 b = *reinterpret_cast<const float4*>(p + 5);
                        xv[1] = a.x; xv[2] = a.y; xv[3] = a.z; xv[4] = a.w;
                        xv[5] = b.x; xv[6] = b.y; xv[7] = b.z; xv[8] = b.w;
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float v = acc[q][j][kh * 3 + kw];
                                v = fmaf(d[q].x, xv[kw], v);
                                v = fmaf(d[q].y, xv[kw + 2], v);
                                v = fmaf(d[q].z, xv[kw + 4], v);
                                v = fmaf(d[q].w, xv[kw + 6], v);
                                acc[q][j][kh * 3 + kw] = v;
                            }
                    }
            }
        }
    }
    if constexpr (K::PS > 1) {
        __syncthreads();
        if (slice > 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int t = 0; t < 9; ++t)
                        smem[((slice - 1) * 72 + (q * 2 + j) * 9 + t) * K::TC + (tid % K::TC)] = acc[q][j][t];
        }
        __syncthreads();
        if (slice == 0) {
            for (int sl = 0; sl < K::PS - 1; ++sl)
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int t = 0; t < 9; ++t) acc[q][j][t] += smem[(sl * 72 + (q * 2 + j) * 9 + t) * K::TC + tid];
        }
    }
    if (slice == 0) {
        float* dst = partial + static_cast<size_t>(blockIdx.x) * 72 * K::TC + tid;
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int t = 0; t < 9; ++t) dst[((q * 2 + j) * 9 + t) * K::TC] = acc[q][j][t];
    }
}

template <class K>
__global__ void __launch_bounds__(256) conv3x3s2_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int nbx,
                                                                     int accumulate) {
    __shared__ float red[8][32];
    const int idx = blockIdx.x * 32 + threadIdx.x;          // < 72 * TC = COUT * CIN * 9 (a multiple of 32)
    const int tc = idx % K::TC, k = idx / K::TC;
    float s = 0.f;
#pragma unroll 4
    for (int b = threadIdx.y; b < nbx; b += 8) s += partial[(static_cast<size_t>(b) * 72 + k) * K::TC + tc];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0) {
#pragma unroll
        for (int i = 1; i < 8; ++i) s += red[i][threadIdx.x];
        const int q = k / 18, j = (k / 9) % 2, t = k % 9;
        const int cg = tc % K::NCG, cp = tc / K::NCG;
        const int co = cg + K::NCG * q, ci = 2 * cp + j;
        float* dst = dw + (co * K::CIN + ci) * 9 + t;
        *dst = accumulate ? *dst + s : s;
    }
}

template <class K>
static int launch_s2_wgrad(const float* x, const float* dy, float* dw, float* ws, long long ws_bytes, int n, int accumulate,
                           cudaStream_t st) {
    static bool attr_set = false;      // benign race: idempotent
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv3x3s2_wgrad_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES) != cudaSuccess) {
            cudaGetLastError();
            return AFAN_ERR_LAUNCH;
        }
        attr_set = true;
    }
    const int units = n * K::UPI;
    int gx = sm_count();
    if (gx > units) gx = units;
    if (static_cast<long long>(gx) * K::COUT * K::CIN * 9 * 4 > ws_bytes) return AFAN_ERR_WORKSPACE;
    conv3x3s2_wgrad_kernel<K><<<gx, K::THREADS, K::SMEM_BYTES, st>>>(x, dy, ws, n);
    if (launch_status() != AFAN_OK) return AFAN_ERR_LAUNCH;
    conv3x3s2_wgrad_reduce_kernel<K><<<K::COUT * K::CIN * 9 / 32, dim3(32, 8), 0, st>>>(ws, dw, gx, accumulate);
    return launch_status();
}

// ---- tensor-core forward / dgrad (mma.sync m16n8k8 TF32, fp32 accumulate) -----------------------------------------------
// Same CTA tile and cp.async pipeline as the FFMA kernel, but the tile is staged channel-innermost -- in_s[pixel][8 ci],
// w_s[tap][co][8 ci][hi, lo] -- so that one LDS.64 yields the (k, k+4) pair of an A fragment and one LDS.128 the B pair
// with its TF32 split.  NPASS = 3 is the "3xTF32" scheme: x = x_hi + x_lo (x_hi = x rounded to TF32), products
// a_lo*b_hi + a_hi*b_lo + a_hi*b_hi accumulated in fp32 recover fp32-level accuracy (the dropped a_lo*b_lo term is 2^-22
// relative) on the tensor pipe; NPASS = 1 is plain TF32 (opt-in, PyTorch's default cuDNN conv math on Ampere+).
// MMA k index <-> channel: k = tig <-> ci 2*tig, k = tig + 4 <-> ci 2*tig + 1 (any bijection works if A and B agree).
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
    const unsigned s = static_cast<unsigned>(__cvta_generic_to_shared(smem_dst));
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(s), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// round-to-nearest (ties away) to TF32's 10-bit mantissa with integer ops; the remainder x - hi is exact in fp32
__device__ __forceinline__ uint32_t tf32_hi(float x) { return (__float_as_uint(x) + 0x1000u) & 0xffffe000u; }

template <int C_, int HW_, int R_, int MT_, int NT_, int NPASS_>
struct MmaCfg {
    static constexpr int C = C_, H = HW_, W = HW_, R = R_, MT = MT_, NT = NT_, NPASS = NPASS_;
    static constexpr int MTILES = R * W / 16, NTILES = C / 8;
    static constexpr int MG = MTILES / MT, NG = NTILES / NT, WARPS = MG * NG, THREADS = WARPS * 32;
    static constexpr int ROWS = R + 2, PW = W + 2;                 // tile incl. halo, in pixels
    static constexpr int IN_ELEMS = ROWS * PW * 8;
    static constexpr int WPK = NPASS == 3 ? 16 : 8;                // floats per (tap, co): [8 ci][hi, lo] or [8 ci]
    static constexpr int W_ELEMS = 9 * C * WPK;
    static constexpr int STAGE = IN_ELEMS + W_ELEMS;
    static constexpr int SMEM_BYTES = 2 * STAGE * 4;
    static constexpr int NCHUNK = C / 8;
    static constexpr int TPI = H / R;
    static_assert(MTILES % MT == 0 && NTILES % NT == 0 && THREADS <= 1024, "warp tiling");
    static_assert(W == 8 ? R % 2 == 0 : W % 16 == 0, "an m-tile is 16 pixels of a row, or two rows of an 8-wide map");
};

template <class K>
__global__ void __launch_bounds__(K::THREADS)
conv3x3_mma_kernel(const float* __restrict__ x, const float* __restrict__ wm, float* __restrict__ y, const float* __restrict__ addend) {
    constexpr int C = K::C, H = K::H, W = K::W, R = K::R, MT = K::MT, NT = K::NT, PW = K::PW;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int g = lane >> 2, tig = lane & 3;
    const int n = blockIdx.x / K::TPI, r0 = (blockIdx.x % K::TPI) * R;
    const int mg = warp % K::MG, ng = warp / K::MG;

    for (int i = tid; i < 2 * K::STAGE / 4; i += K::THREADS)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    auto load_chunk = [&](int chunk, int buf) {
        float* in_s = smem + buf * K::STAGE;
        float* w_s = in_s + K::IN_ELEMS;
        const int ci0 = chunk * 8;
        constexpr int NIN = 8 * K::ROWS * W;
        for (int i = tid; i < NIN; i += K::THREADS) {
            const int col = i % W, row = (i / W) % K::ROWS, c = i / (W * K::ROWS);
            const int gr = r0 - 1 + row;
            if (gr >= 0 && gr < H)
                cp_async4(in_s + (row * PW + col + 1) * 8 + c, x + (static_cast<size_t>(n * C + ci0 + c) * H + gr) * W + col);
        }
        const float* wsrc = wm + static_cast<size_t>(chunk) * K::W_ELEMS;
        for (int i = tid; i < K::W_ELEMS / 4; i += K::THREADS) cp_async16(w_s + 4 * i, wsrc + 4 * i);
        cp_async_commit();
    };

    // pixel offsets (in floats) of the two row halves of each m-tile this warp owns, tap (0, 0)
    int pa[MT][2];
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int mt = mg * MT + i;
        int row0, col0, row1, col1;
        if constexpr (W == 8) { row0 = 2 * mt; col0 = g; row1 = 2 * mt + 1; col1 = g; }
        else { row0 = row1 = mt / (W / 16); col0 = (mt % (W / 16)) * 16 + g; col1 = col0 + 8; }
        pa[i][0] = (row0 * PW + col0) * 8 + 2 * tig;
        pa[i][1] = (row1 * PW + col1) * 8 + 2 * tig;
    }

    float acc[MT][NT][4];
#pragma unroll
    for (int i = 0; i < MT; ++i)
#pragma unroll
        for (int j = 0; j < NT; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) acc[i][j][e] = 0.f;

    load_chunk(0, 0);
#pragma unroll 1
    for (int ch = 0; ch < K::NCHUNK; ++ch) {
        cp_async_wait_all();
        __syncthreads();
        if (ch + 1 < K::NCHUNK) load_chunk(ch + 1, (ch + 1) & 1);
        const float* in_s = smem + (ch & 1) * K::STAGE;
        const float* w_s = in_s + K::IN_ELEMS + (K::NPASS == 3 ? ((ng * NT) * 8 + g) * 16 + 4 * tig : ((ng * NT) * 8 + g) * 8 + 2 * tig);
#pragma unroll
        for (int t = 0; t < 9; ++t) {
            const int toff = ((t / 3) * PW + (t % 3)) * 8;
            uint32_t ah[MT][4], al[MT][4];
#pragma unroll
            for (int i = 0; i < MT; ++i) {
                const float2 p0 = *reinterpret_cast<const float2*>(in_s + pa[i][0] + toff);   // (g,   k) , (g,   k+4)
                const float2 p1 = *reinterpret_cast<const float2*>(in_s + pa[i][1] + toff);   // (g+8, k) , (g+8, k+4)
                const float v[4] = {p0.x, p1.x, p0.y, p1.y};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    ah[i][e] = tf32_hi(v[e]);
                    if constexpr (K::NPASS == 3) al[i][e] = tf32_hi(v[e] - __uint_as_float(ah[i][e]));
                }
            }
            uint32_t bh[NT][2], bl[NT][2];
#pragma unroll
            for (int j = 0; j < NT; ++j) {
                if constexpr (K::NPASS == 3) {
                    const float4 b = *reinterpret_cast<const float4*>(w_s + (t * C + j * 8) * 16);   // hi(k), lo(k), hi(k+4), lo(k+4)
                    bh[j][0] = __float_as_uint(b.x); bl[j][0] = __float_as_uint(b.y);
                    bh[j][1] = __float_as_uint(b.z); bl[j][1] = __float_as_uint(b.w);
                } else {
                    const float2 b = *reinterpret_cast<const float2*>(w_s + (t * C + j * 8) * 8);    // hi(k), hi(k+4)
                    bh[j][0] = __float_as_uint(b.x); bh[j][1] = __float_as_uint(b.y);
                }
            }
            // the three passes of a tile are issued MT*NT MMAs apart, so consecutive MMAs never chain on one accumulator
            if constexpr (K::NPASS == 3) {
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int i = 0; i < MT; ++i) mma_tf32(acc[i][j], al[i], bh[j][0], bh[j][1]);
#pragma unroll
                for (int j = 0; j < NT; ++j)
#pragma unroll
                    for (int i = 0; i < MT; ++i) mma_tf32(acc[i][j], ah[i], bl[j][0], bl[j][1]);
            }
#pragma unroll
            for (int j = 0; j < NT; ++j)
#pragma unroll
                for (int i = 0; i < MT; ++i) mma_tf32(acc[i][j], ah[i], bh[j][0], bh[j][1]);
        }
    }
#pragma unroll
    for (int i = 0; i < MT; ++i) {
        const int mt = mg * MT + i;
        int row0, col0, row1, col1;
        if constexpr (W == 8) { row0 = 2 * mt; col0 = g; row1 = 2 * mt + 1; col1 = g; }
        else { row0 = row1 = mt / (W / 16); col0 = (mt % (W / 16)) * 16 + g; col1 = col0 + 8; }
#pragma unroll
        for (int j = 0; j < NT; ++j) {
            const int co = (ng * NT + j) * 8 + 2 * tig;
            const size_t o0 = (static_cast<size_t>(n * C + co) * H + r0) * W, o1 = o0 + static_cast<size_t>(H) * W;
            const size_t e0 = o0 + row0 * W + col0, e1 = o1 + row0 * W + col0, e2 = o0 + row1 * W + col1, e3 = o1 + row1 * W + col1;
            if (addend != nullptr) {
                acc[i][j][0] += addend[e0]; acc[i][j][1] += addend[e1]; acc[i][j][2] += addend[e2]; acc[i][j][3] += addend[e3];
            }
            y[e0] = acc[i][j][0];
            y[e1] = acc[i][j][1];
            y[e2] = acc[i][j][2];
            y[e3] = acc[i][j][3];
        }
    }
}

// W[co][ci][3][3] -> tensor-core packings [chunk = k / 8][tap][out][k % 8][hi, lo]: forward (k = ci, out = co, tap t) and
// dgrad (k = co, out = ci, tap 8 - t).  hi = TF32 rounding of w, lo = TF32 rounding of (w - hi).
__global__ void __launch_bounds__(kThreads) conv3x3_pack_mma_kernel(const ConvPackDesc* __restrict__ descs, int passes) {
    const ConvPackDesc d = descs[blockIdx.y];
    if (d.c < 0) { pack_stride2(d, blockIdx.x * kThreads + threadIdx.x, gridDim.x * kThreads); return; }   // fp32 kernels in every mode
    const int C = static_cast<int>(d.c), total = C * C * 9;
    for (int idx = blockIdx.x * kThreads + threadIdx.x; idx < total; idx += gridDim.x * kThreads) {
        const int co = idx / (C * 9), ci = (idx / 9) % C, t = idx % 9;
        const float v = d.w[idx];
        const float hi = __uint_as_float(tf32_hi(v));
        const float lo = __uint_as_float(tf32_hi(v - hi));
        const int f = (((ci >> 3) * 9 + t) * C + co) * 8 + (ci & 7);
        const int b = (((co >> 3) * 9 + (8 - t)) * C + ci) * 8 + (co & 7);
        if (passes == 3) {
            d.wf[2 * f] = hi; d.wf[2 * f + 1] = lo;
            d.wd[2 * b] = hi; d.wd[2 * b + 1] = lo;
        } else {
            d.wf[f] = hi;
            d.wd[b] = hi;
        }
    }
}

template <class K>
static int launch_conv_mma(const float* x, const float* wm, float* y, const float* addend, int n, cudaStream_t st) {
    static bool attr_set = false;      // benign race: idempotent
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv3x3_mma_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES) != cudaSuccess) {
            cudaGetLastError();
            return AFAN_ERR_LAUNCH;
        }
        attr_set = true;
    }
    conv3x3_mma_kernel<K><<<static_cast<unsigned>(n * K::TPI), K::THREADS, K::SMEM_BYTES, st>>>(x, wm, y, addend);
    return launch_status();
}

// ---- wgrad ----------------------------------------------------------------------------------------------------
// dW[co,ci,kh,kw] = sum over images and pixels of dy[n,co,h,w] * x[n,ci,h+kh-1,w+kw-1].  Persistent CTAs walk
// "units" (an RB-row band of one image) through a 2-stage cp.async pipeline; a thread owns 4 output channels x 2
// input channels x 9 taps (72 accumulators) and consumes 4 pixels per step: 4 LDS.128 of dy + 2x3 x (LDS.32, LDS.128,
// LDS.32) of x feed 288 FFMAs.  Lanes walk the output-channel groups (dy is staged as [pixel quad][co][4] so those
// loads are one contiguous 16-byte word per lane); the x words are shared by the whole group (broadcast).  PS thread
// slices take alternate rows of the band and are folded in shared memory; every CTA then writes its partial dW and a
// second kernel sums the partials in CTA order: no atomics, deterministic.
template <int C_, int HW_, int RB_, int PS_, int NY_>
struct WgradCfg {
    static constexpr int C = C_, H = HW_, W = HW_, RB = RB_, PS = PS_, NY = NY_;
    static constexpr int S = (W == 32) ? 40 : (W == 16 ? 24 : 20);
    static constexpr int CIB = C / NY;                     // input channels per CTA (blockIdx.y selects the slice)
    static constexpr int NCG = C / 4, NCP = CIB / 2, TC = NCG * NCP, THREADS = TC * PS;
    static constexpr int XROWS = RB + 2;
    static constexpr int PLANE = ((XROWS * S + 23) / 32) * 32 + 8;   // == 8 (mod 32): neighbouring channels on distinct banks
    static constexpr int X_ELEMS = CIB * PLANE;
    static constexpr int DY_ELEMS = RB * W * C;
    static constexpr int STAGE = X_ELEMS + DY_ELEMS;
    static constexpr int RED_ELEMS = (PS - 1) * 72 * TC;
    static constexpr int SMEM_ELEMS = (2 * STAGE > RED_ELEMS) ? 2 * STAGE : RED_ELEMS;
    static constexpr int SMEM_BYTES = SMEM_ELEMS * 4;
    static constexpr int UPI = H / RB;                     // units per image
    static_assert(PLANE >= XROWS * S && PLANE % 4 == 0, "plane pitch");
    static_assert(RB % PS == 0 && H % RB == 0 && THREADS % 32 == 0 && THREADS <= 512, "tiling");
};

template <class K>
__global__ void __launch_bounds__(K::THREADS)
conv3x3_wgrad_kernel(const float* __restrict__ x, const float* __restrict__ dy, float* __restrict__ partial, int n_images) {
    constexpr int C = K::C, H = K::H, W = K::W, RB = K::RB, S = K::S;
    extern __shared__ __align__(16) float smem[];
    const int tid = threadIdx.x;
    const int cg = tid % K::NCG, cp = (tid / K::NCG) % K::NCP, slice = tid / K::TC;
    const int ci_base = blockIdx.y * K::CIB;
    const int units = n_images * K::UPI;

    for (int i = tid; i < 2 * K::STAGE / 4; i += K::THREADS)
        reinterpret_cast<float4*>(smem)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    __syncthreads();

    auto load_unit = [&](int u, int buf) {
        float* x_s = smem + buf * K::STAGE;
        float* dy_s = x_s + K::X_ELEMS;
        const int n = u / K::UPI, r0 = (u % K::UPI) * RB;
        constexpr int SEGS = W / 4, NX = K::CIB * K::XROWS * SEGS, ND = C * RB * SEGS;
        for (int i = tid; i < NX; i += K::THREADS) {
            const int seg = i % SEGS, row = (i / SEGS) % K::XROWS, cil = i / (SEGS * K::XROWS);
            const int gr = r0 - 1 + row;
            float* dst = x_s + cil * K::PLANE + row * S + 4 + 4 * seg;
            if (gr >= 0 && gr < H)
                cp_async16(dst, x + (static_cast<size_t>(n * C + ci_base + cil) * H + gr) * W + 4 * seg);
            else
                *reinterpret_cast<float4*>(dst) = make_float4(0.f, 0.f, 0.f, 0.f);   // band touches the image border
        }
        for (int i = tid; i < ND; i += K::THREADS) {
            const int seg = i % SEGS, row = (i / SEGS) % RB, co = i / (SEGS * RB);
            cp_async16(dy_s + ((row * SEGS + seg) * C + co) * 4,
                       dy + (static_cast<size_t>(n * C + co) * H + r0 + row) * W + 4 * seg);
        }
        cp_async_commit();
    };

    float acc[4][2][9];
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
        for (int j = 0; j < 2; ++j)
#pragma unroll
            for (int t = 0; t < 9; ++t) acc[q][j][t] = 0.f;

    int u = blockIdx.x, buf = 0;
    if (u < units) load_unit(u, 0);
#pragma unroll 1
    for (; u < units; u += gridDim.x, buf ^= 1) {
        cp_async_wait_all();
        __syncthreads();
        if (u + static_cast<int>(gridDim.x) < units) load_unit(u + gridDim.x, buf ^ 1);
        const float* x_s = smem + buf * K::STAGE + (2 * cp) * K::PLANE + 3;
        const float* dy_s = smem + buf * K::STAGE + K::X_ELEMS + cg * 4;
#pragma unroll 1
        for (int r = slice; r < RB; r += K::PS) {
#pragma unroll 2
            for (int wq = 0; wq < W / 4; ++wq) {
                float4 d[4];
#pragma unroll
                for (int q = 0; q < 4; ++q)
                    d[q] = *reinterpret_cast<const float4*>(dy_s + ((r * (W / 4) + wq) * C + K::NCG * q) * 4);
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int kh = 0; kh < 3; ++kh) {
                        const float* p = x_s + j * K::PLANE + (r + kh) * S + 4 * wq;
                        float xv[6];
                        xv[0] = p[0];
                        const float4 m = *reinterpret_cast<const float4*>(p + 1);
                        xv[1] = m.x; xv[2] = m.y; xv[3] = m.z; xv[4] = m.w;
                        xv[5] = p[5];
#pragma unroll
                        for (int kw = 0; kw < 3; ++kw)
#pragma unroll
                            for (int q = 0; q < 4; ++q) {
                                float a = acc[q][j][kh * 3 + kw];
                                a = fmaf(d[q].x, xv[kw], a);
                                a = fmaf(d[q].y, xv[kw + 1], a);
                                a = fmaf(d[q].z, xv[kw + 2], a);
                                a = fmaf(d[q].w, xv[kw + 3], a);
                                acc[q][j][kh * 3 + kw] = a;
                            }
                    }
            }
        }
    }
    // fold the PS row slices (fixed order), then publish this CTA's partial as [k = (q, j, t)][thread of slice 0]
    if constexpr (K::PS > 1) {
        __syncthreads();               // everyone is done reading the tiles; reuse the shared memory
        if (slice > 0) {
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int j = 0; j < 2; ++j)
#pragma unroll
                    for (int t = 0; t < 9; ++t)
                        smem[((slice - 1) * 72 + (q * 2 + j) * 9 + t) * K::TC + (tid % K::TC)] = acc[q][j][t];
        }
        __syncthreads();
        if (slice == 0) {
            for (int s = 0; s < K::PS - 1; ++s)
#pragma unroll
                for (int q = 0; q < 4; ++q)
#pragma unroll
                    for (int j = 0; j < 2; ++j)
#pragma unroll
                        for (int t = 0; t < 9; ++t) acc[q][j][t] += smem[(s * 72 + (q * 2 + j) * 9 + t) * K::TC + tid];
        }
    }
    if (slice == 0) {
        float* dst = partial + static_cast<size_t>(blockIdx.x * K::NY + blockIdx.y) * 72 * K::TC + tid;
#pragma unroll
        for (int q = 0; q < 4; ++q)
#pragma unroll
            for (int j = 0; j < 2; ++j)
#pragma unroll
                for (int t = 0; t < 9; ++t) dst[((q * 2 + j) * 9 + t) * K::TC] = acc[q][j][t];
    }
}

// dW[o] (+)= sum over CTAs of their partials.  Block = 32 consecutive outputs x 8 slices of the CTA range: every thread
// folds its slice in ascending CTA order, the 8 slice sums are folded in slice order (fixed order: deterministic).
template <class K>
__global__ void __launch_bounds__(256) conv3x3_wgrad_reduce_kernel(const float* __restrict__ partial, float* __restrict__ dw, int nbx,
                                                                   int accumulate) {
    __shared__ float red[8][32];
    const int idx = blockIdx.x * 32 + threadIdx.x;          // < NY * 72 * TC (a multiple of 32)
    const int tc = idx % K::TC, k = (idx / K::TC) % 72, y = idx / (K::TC * 72);
    float s = 0.f;
#pragma unroll 4
    for (int b = threadIdx.y; b < nbx; b += 8) s += partial[(static_cast<size_t>(b * K::NY + y) * 72 + k) * K::TC + tc];
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.y == 0) {
#pragma unroll
        for (int i = 1; i < 8; ++i) s += red[i][threadIdx.x];
        const int q = k / 18, j = (k / 9) % 2, t = k % 9;
        const int cg = tc % K::NCG, cp = tc / K::NCG;
        const int co = cg + K::NCG * q, ci = y * K::CIB + 2 * cp + j;
        float* dst = dw + (co * K::C + ci) * 9 + t;
        *dst = accumulate ? *dst + s : s;
    }
}

template <class K>
static int launch_wgrad(const float* x, const float* dy, float* dw, float* ws, long long ws_bytes, int n, int accumulate,
                        cudaStream_t st) {
    static bool attr_set = false;      // benign race: idempotent
    if (!attr_set) {
        if (cudaFuncSetAttribute(conv3x3_wgrad_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::SMEM_BYTES) != cudaSuccess) {
            cudaGetLastError();
            return AFAN_ERR_LAUNCH;
        }
        attr_set = true;
    }
    const int units = n * K::UPI;
    int gx = sm_count() / K::NY;
    if (gx > units) gx = units;
    if (static_cast<long long>(gx) * K::NY * K::C * K::C * 9 * 4 > ws_bytes) return AFAN_ERR_WORKSPACE;
    conv3x3_wgrad_kernel<K><<<dim3(gx, K::NY), K::THREADS, K::SMEM_BYTES, st>>>(x, dy, ws, n);
    if (launch_status() != AFAN_OK) return AFAN_ERR_LAUNCH;
    const int total = K::C * K::C * 9;
    conv3x3_wgrad_reduce_kernel<K><<<total / 32, dim3(32, 8), 0, st>>>(ws, dw, gx, accumulate);
    return launch_status();
}

}  // namespace afan

using namespace afan;

AFAN_EXPORT int afan_conv3x3_pack_f32(const void* descs_device, int64_t n_layers, int64_t c_max, afan_stream_t stream) {
    if (n_layers < 0 || c_max <= 0) return AFAN_ERR_SIZE;
    if (n_layers == 0) return AFAN_OK;
    if (!descs_device) return AFAN_ERR_NULL;
    const long long total = c_max * c_max * 9;
    const unsigned gx = static_cast<unsigned>((total + kThreads - 1) / kThreads);
    conv3x3_pack_kernel<<<dim3(gx < 16u ? gx : 16u, static_cast<unsigned>(n_layers)), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const ConvPackDesc*>(descs_device));
    return launch_status();
}

// variant: 0 = default tiling for the shape; other values select tuning candidates (bench only)
AFAN_EXPORT int afan_conv3x3_f32(const float* x, const float* w_packed, float* y, const float* addend, int64_t n, int64_t c,
                                 int64_t hw, int variant, afan_stream_t stream) {
    if (n < 0 || c <= 0 || hw <= 0) return AFAN_ERR_SIZE;
    if (n == 0) return AFAN_OK;
    if (!x || !w_packed || !y) return AFAN_ERR_NULL;
    if (!aligned16(x) || !aligned16(w_packed) || !aligned16(y) || (addend && !aligned16(addend)) || n * c * hw > (1ll << 30) / hw)
        return AFAN_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ni = static_cast<int>(n);
#define AFAN_CONV(C, HW, R, TQ, CK) return launch_conv<ConvCfg<C, HW, R, TQ, CK>>(x, w_packed, y, addend, ni, st)
    if (c == 16 && hw == 32) {
        if (variant == 1) AFAN_CONV(16, 32, 16, 4, 8);
        if (variant == 2) AFAN_CONV(16, 32, 32, 4, 4);
        if (variant == 3) AFAN_CONV(16, 32, 8, 4, 4);
        AFAN_CONV(16, 32, 16, 4, 4);
    }
    if (c == 32 && hw == 16) {
        if (variant == 1) AFAN_CONV(32, 16, 16, 8, 8);
        if (variant == 2) AFAN_CONV(32, 16, 8, 4, 8);
        if (variant == 3) AFAN_CONV(32, 16, 16, 4, 4);
        AFAN_CONV(32, 16, 16, 4, 8);
    }
    if (c == 64 && hw == 8) {
        if (variant == 1) AFAN_CONV(64, 8, 8, 4, 8);
        if (variant == 2) AFAN_CONV(64, 8, 8, 8, 8);
        if (variant == 3) AFAN_CONV(64, 8, 8, 2, 4);
        AFAN_CONV(64, 8, 8, 2, 8);
    }
    if (c == 16 && hw == 16) AFAN_CONV(16, 16, 16, 4, 4);
    if (c == 16 && hw == 8) AFAN_CONV(16, 8, 8, 4, 4);
    if (c == 32 && hw == 32) AFAN_CONV(32, 32, 16, 4, 8);
    if (c == 32 && hw == 8) AFAN_CONV(32, 8, 8, 4, 8);
    if (c == 64 && hw == 16) AFAN_CONV(64, 16, 16, 4, 8);
    if (c == 64 && hw == 32) AFAN_CONV(64, 32, 16, 8, 8);
#undef AFAN_CONV
    return AFAN_ERR_UNSUPPORTED;
}

AFAN_EXPORT int64_t afan_conv3x3_wgrad_workspace_bytes(int64_t c) {
    return c <= 0 ? 0 : static_cast<int64_t>(sm_count()) * c * c * 9 * 4;
}

AFAN_EXPORT int afan_conv3x3_wgrad_f32(const float* x, const float* dy, float* dw, void* workspace, int64_t workspace_bytes,
                                       int64_t n, int64_t c, int64_t hw, int accumulate, afan_stream_t stream) {
    if (n <= 0 || c <= 0 || hw <= 0) return AFAN_ERR_SIZE;
    if (!x || !dy || !dw) return AFAN_ERR_NULL;
    if (!workspace) return AFAN_ERR_WORKSPACE;
    if (!aligned16(x) || !aligned16(dy) || !aligned16(workspace) || n * c * hw > (1ll << 30) / hw) return AFAN_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ni = static_cast<int>(n);
    float* ws = static_cast<float*>(workspace);
#define AFAN_WGRAD(C, HW, RB, PS, NY) return launch_wgrad<WgradCfg<C, HW, RB, PS, NY>>(x, dy, dw, ws, workspace_bytes, ni, accumulate, st)
    if (c == 16 && hw == 32) AFAN_WGRAD(16, 32, 8, 8, 1);
    if (c == 16 && hw == 16) AFAN_WGRAD(16, 16, 8, 8, 1);
    if (c == 16 && hw == 8) AFAN_WGRAD(16, 8, 8, 8, 1);
    if (c == 32 && hw == 32) AFAN_WGRAD(32, 32, 2, 2, 1);
    if (c == 32 && hw == 16) AFAN_WGRAD(32, 16, 2, 2, 1);
    if (c == 32 && hw == 8) AFAN_WGRAD(32, 8, 2, 2, 1);
    if (c == 64 && hw == 32) AFAN_WGRAD(64, 32, 2, 1, 2);
    if (c == 64 && hw == 16) AFAN_WGRAD(64, 16, 2, 1, 2);
    if (c == 64 && hw == 8) AFAN_WGRAD(64, 8, 2, 1, 2);
#undef AFAN_WGRAD
    return AFAN_ERR_UNSUPPORTED;
}

/* Tensor-core twins.  Packed weights are 2*c*9*c floats per direction (hi/lo TF32 split). */
AFAN_EXPORT int afan_conv3x3_pack_tc_f32(const void* descs_device, int64_t n_layers, int64_t c_max, int passes,
                                         afan_stream_t stream) {
    if (n_layers < 0 || c_max <= 0 || (passes != 1 && passes != 3)) return AFAN_ERR_SIZE;
    if (n_layers == 0) return AFAN_OK;
    if (!descs_device) return AFAN_ERR_NULL;
    const long long total = c_max * c_max * 9;
    const unsigned gx = static_cast<unsigned>((total + kThreads - 1) / kThreads);
    conv3x3_pack_mma_kernel<<<dim3(gx < 16u ? gx : 16u, static_cast<unsigned>(n_layers)), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const ConvPackDesc*>(descs_device), passes);
    return launch_status();
}

AFAN_EXPORT int afan_conv3x3_tc_f32(const float* x, const float* w_packed, float* y, const float* addend, int64_t n, int64_t c,
                                    int64_t hw, int passes, int variant, afan_stream_t stream) {
    if (n < 0 || c <= 0 || hw <= 0 || (passes != 1 && passes != 3)) return AFAN_ERR_SIZE;
    if (n == 0) return AFAN_OK;
    if (!x || !w_packed || !y) return AFAN_ERR_NULL;
    if (!aligned16(x) || !aligned16(w_packed) || !aligned16(y) || n * c * hw > (1ll << 30) / hw) return AFAN_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ni = static_cast<int>(n);
#define AFAN_MMA(C, HW, R, MT, NT)                                                                   \
    do {                                                                                             \
        if (passes == 3) return launch_conv_mma<MmaCfg<C, HW, R, MT, NT, 3>>(x, w_packed, y, addend, ni, st); \
        return launch_conv_mma<MmaCfg<C, HW, R, MT, NT, 1>>(x, w_packed, y, addend, ni, st);                  \
    } while (0)
    if (c == 16 && hw == 32) { if (variant == 1) AFAN_MMA(16, 32, 8, 2, 2); AFAN_MMA(16, 32, 16, 4, 2); }
    if (c == 32 && hw == 16) { if (variant == 1) AFAN_MMA(32, 16, 16, 4, 4); AFAN_MMA(32, 16, 16, 2, 4); }
    if (c == 64 && hw == 8) { if (variant == 1) AFAN_MMA(64, 8, 8, 4, 2); AFAN_MMA(64, 8, 8, 2, 2); }
    if (c == 16 && hw == 16) AFAN_MMA(16, 16, 16, 2, 2);
    if (c == 16 && hw == 8) AFAN_MMA(16, 8, 8, 2, 2);
    if (c == 32 && hw == 32) AFAN_MMA(32, 32, 16, 4, 4);
    if (c == 32 && hw == 8) AFAN_MMA(32, 8, 8, 2, 2);
    if (c == 64 && hw == 16) AFAN_MMA(64, 16, 16, 2, 4);
    if (c == 64 && hw == 32) AFAN_MMA(64, 32, 16, 4, 4);
#undef AFAN_MMA
    return AFAN_ERR_UNSUPPORTED;
}

/* Stride-2 stage transitions: x [n][cin][2*ho][2*ho] -> y [n][2*cin][ho][ho] (forward), or dy -> dx (dgrad != 0), with the
 * packings afan_conv3x3_pack_f32 writes for a descriptor whose c field is -cin.  (cin, ho) in {(16, 16), (32, 8)}. */
AFAN_EXPORT int afan_conv3x3s2_f32(const float* in, const float* w_packed, float* out, int64_t n, int64_t cin, int64_t ho,
                                   int dgrad, afan_stream_t stream) {
    if (n < 0 || cin <= 0 || ho <= 0) return AFAN_ERR_SIZE;
    if (n == 0) return AFAN_OK;
    if (!in || !w_packed || !out) return AFAN_ERR_NULL;
    if (!aligned16(in) || !aligned16(w_packed) || !aligned16(out) || n > (1 << 20)) return AFAN_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int ni = static_cast<int>(n);
    if (cin == 16 && ho == 16) {
        if (dgrad) { using K = S2DgradCfg<16, 16, 2, 8>; return launch_s2<K>(conv3x3s2_dgrad_kernel<K>, in, w_packed, out, ni, st); }
        using K = S2Cfg<16, 16, 4, 4>;
        return launch_s2<K>(conv3x3s2_kernel<K>, in, w_packed, out, ni, st);
    }
    if (cin == 32 && ho == 8) {
        if (dgrad) { using K = S2DgradCfg<32, 8, 2, 8>; return launch_s2<K>(conv3x3s2_dgrad_kernel<K>, in, w_packed, out, ni, st); }
        using K = S2Cfg<32, 8, 2, 8, 4>;                   // measured: TQ=2 11.2 us, TQ=4 19.1 us, TQ=8 34.6 us at batch 128
        return launch_s2<K>(conv3x3s2_kernel<K>, in, w_packed, out, ni, st);
    }
    return AFAN_ERR_UNSUPPORTED;
}

/* Weight gradient of the stride-2 transition: x [n][cin][2*ho][2*ho], dy [n][2*cin][ho][ho] -> dw [2*cin][cin][3][3]
 * (stored, or added when accumulate != 0).  workspace >= afan_conv3x3_wgrad_workspace_bytes(2*cin).  Deterministic. */
AFAN_EXPORT int afan_conv3x3s2_wgrad_f32(const float* x, const float* dy, float* dw, void* workspace, int64_t workspace_bytes,
                                         int64_t n, int64_t cin, int64_t ho, int accumulate, afan_stream_t stream) {
    if (n <= 0 || cin <= 0 || ho <= 0) return AFAN_ERR_SIZE;
    if (!x || !dy || !dw) return AFAN_ERR_NULL;
    if (!workspace) return AFAN_ERR_WORKSPACE;
    if (!aligned16(x) || !aligned16(dy) || !aligned16(workspace) || n > (1 << 20)) return AFAN_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    float* ws = static_cast<float*>(workspace);
    const int ni = static_cast<int>(n);
    if (cin == 16 && ho == 16) return launch_s2_wgrad<S2WgradCfg<16, 16, 4, 4>>(x, dy, dw, ws, workspace_bytes, ni, accumulate, st);
    if (cin == 32 && ho == 8) return launch_s2_wgrad<S2WgradCfg<32, 8, 2, 1>>(x, dy, dw, ws, workspace_bytes, ni, accumulate, st);
    return AFAN_ERR_UNSUPPORTED;
}
