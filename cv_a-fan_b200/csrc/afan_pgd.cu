// afan_pgd.cu -- feature-space PGD kernels for sm_100a (B200).
//
//   a2  random start            Classification/attack_algo.py:41-44
//   a3  sign-gradient ascent    Classification/attack_algo.py:53
//   a4  L-inf projection        Classification/attack_algo.py:55-56 -> :35-36 -> :9-19
//   a11 perturbation norms      Classification/main_perturb.py:188-192
//
// All of it is HBM-bound streaming work: the step kernel reads grad, x_adv and the cached clean
// feature ONCE with 128-bit coalesced loads and writes x_adv (and delta) ONCE -- 20 B/elem (clip +
// delta), 16 B/elem (clip), 12 B/elem (no clip) -- where the reference moves ~28-60 B/elem through
// 13 un-fused launches and 4 host syncs.  Arithmetic is ordered exactly like the reference's op
// sequence so x_adv is bit-identical: every fp32 op below is an explicit
// round-to-nearest intrinsic, so nvcc cannot contract or reassociate it.
#include <cuda_bf16.h>

#include "afan_common.cuh"

namespace afan {

// torch.sign for floats: (0 < a) - (a < 0)  ->  sign(NaN) = sign(+-0) = +0
__device__ __forceinline__ float torch_sign(float a) {
    return static_cast<float>(static_cast<int>(0.0f < a) - static_cast<int>(a < 0.0f));
}

// one element of attack_algo.py:53-56 (tensor_clamp's compare-and-assign keeps NaN untouched)
template <bool STEP, bool CLIP>
__device__ __forceinline__ float linf_update(float g, float xa, float xc, float gamma, float eps) {
    float t = STEP ? __fadd_rn(xa, __fmul_rn(gamma, torch_sign(g))) : xa;
    if (CLIP) {
        const float lo = __fsub_rn(xc, eps), hi = __fadd_rn(xc, eps);
        t = (t < lo) ? lo : t;
        t = (t > hi) ? hi : t;
    }
    return t;
}

struct NormAcc {
    float sumsq = 0.f, maxabs = 0.f;
    int nan = 0;
    __device__ __forceinline__ void add(float d) {
        sumsq = fmaf(d, d, sumsq);
        const float a = fabsf(d);
        nan |= (a != a);
        maxabs = fmaxf(maxabs, a);
    }
};

// Per-sample reduction epilogue shared by the step kernel (a11) and the L2 kernels: warp shuffle, then
// a shared-memory tree (sum of squares in double, max in float) -> one partial per CTA; the CTA that
// arrives last for the sample folds the partials in a FIXED order (deterministic, no float atomics).
__device__ __forceinline__ void finish_sample_norms(const NormAcc& acc, float* partials, unsigned int* counters,
                                                    int s, int k, int chunks, float* l2_out, float* linf_out) {
    __shared__ double scratch[64];
    __shared__ float smax[32];
    __shared__ int snan, sflag;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) snan = 0;
    __syncthreads();
    const float wmx = warp_max(acc.maxabs);
    if (__any_sync(0xffffffffu, acc.nan) && lane == 0) atomicOr(&snan, 1);
    if (lane == 0) smax[warp] = wmx;
    double sq = acc.sumsq, unused = 0.0;
    block_sum2(sq, unused, scratch);                   // __syncthreads inside: smax / snan now visible
    if (chunks == 1) {                                 // the CTA owns the whole sample: publish directly
        if (threadIdx.x == 0) {
            float m = 0.f;
            for (int w = 0; w < kThreads / 32; ++w) m = fmaxf(m, smax[w]);
            if (l2_out) l2_out[s] = static_cast<float>(sqrt(sq));
            if (linf_out) linf_out[s] = snan ? __int_as_float(0x7fc00000) : m;
        }
        return;
    }
    if (threadIdx.x == 0) {
        float m = 0.f;
        for (int w = 0; w < kThreads / 32; ++w) m = fmaxf(m, smax[w]);
        float* p = partials + (static_cast<long long>(s) * chunks + k) * 4;
        p[0] = static_cast<float>(sq);
        p[1] = m;
        p[2] = snan ? 1.f : 0.f;
    }
    if (!last_cta_arrives(counters + s, chunks, &sflag)) return;
    const volatile float* p = partials + static_cast<long long>(s) * chunks * 4;
    double tot = 0.0;
    float m = 0.f, nn = 0.f;
    for (int j = threadIdx.x; j < chunks; j += kThreads) {
        tot += static_cast<double>(p[j * 4]);
        m = fmaxf(m, p[j * 4 + 1]);
        nn = fmaxf(nn, p[j * 4 + 2]);
    }
    m = warp_max(m);
    nn = warp_max(nn);
    if (lane == 0) smax[warp] = m;
    if (lane == 0 && nn > 0.f) atomicOr(&snan, 1);
    unused = 0.0;
    block_sum2(tot, unused, scratch);
    if (threadIdx.x == 0) {
        float mm = 0.f;
        for (int w = 0; w < kThreads / 32; ++w) mm = fmaxf(mm, smax[w]);
        if (l2_out) l2_out[s] = static_cast<float>(sqrt(tot));
        if (linf_out) linf_out[s] = snan ? __int_as_float(0x7fc00000) : mm;
    }
}

template <int VEC> struct Vec;
template <> struct Vec<4> { using type = float4; };
template <> struct Vec<1> { using type = float; };

template <bool STEP, bool CLIP, bool DELTA, bool NORMS>
__device__ __forceinline__ void step_elem(float g, float& xa, float xc, float& d, float gamma, float eps,
                                          NormAcc& acc) {
    xa = linf_update<STEP, CLIP>(g, xa, xc, gamma, eps);
    if (DELTA || NORMS) {
        d = __fsub_rn(xa, xc);
        if (NORMS) acc.add(d);
    }
}

constexpr int kUnroll = 4;             // independent 128-bit loads in flight per thread per tensor

// grid = (chunks, samples).  CTA (k, s) owns vector indices [k*pv/chunks, (k+1)*pv/chunks) of sample s
// (pv = per_sample / VEC).  Without norms the whole tensor is one "sample".
template <int VEC, bool STEP, bool CLIP, bool DELTA, bool NORMS>
__global__ void __launch_bounds__(kThreads)
pgd_linf_step_kernel(const float* __restrict__ grad, const float* __restrict__ x_clean, float* x_adv,
                     float* __restrict__ delta_out, float* __restrict__ norms_out, float* partials,
                     unsigned int* counters, long long pv, int n_samples, float gamma, float eps) {
    using V = typename Vec<VEC>::type;
    constexpr bool NEED_CLEAN = CLIP || DELTA || NORMS;
    const int chunks = gridDim.x, k = blockIdx.x, s = blockIdx.y;
    const long long lo = pv * k / chunks, hi = pv * (k + 1) / chunks;
    const long long base = static_cast<long long>(s) * pv;
    pdl_wait();
    pdl_launch_dependents();
    const V* g_v = STEP ? reinterpret_cast<const V*>(grad) + base : nullptr;
    const V* c_v = NEED_CLEAN ? reinterpret_cast<const V*>(x_clean) + base : nullptr;
    V* a_v = reinterpret_cast<V*>(x_adv) + base;
    V* d_v = DELTA ? reinterpret_cast<V*>(delta_out) + base : nullptr;
    NormAcc acc;

    for (long long i0 = lo + threadIdx.x; i0 < hi; i0 += static_cast<long long>(kThreads) * kUnroll) {
        V g[kUnroll] = {}, a[kUnroll] = {}, c[kUnroll] = {};
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {            // all loads first: 3*kUnroll requests in flight
            const long long i = i0 + static_cast<long long>(u) * kThreads;
            if (i < hi) {
                if (STEP) g[u] = ld_stream(g_v + i);   // gradient: read once, evict first
                a[u] = a_v[i];
                if (NEED_CLEAN) c[u] = ld_ro(c_v + i); // clean feature: re-read every step, keep in L2
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + static_cast<long long>(u) * kThreads;
            if (i < hi) {
                V d;
                if constexpr (VEC == 4) {
                    step_elem<STEP, CLIP, DELTA, NORMS>(g[u].x, a[u].x, c[u].x, d.x, gamma, eps, acc);
                    step_elem<STEP, CLIP, DELTA, NORMS>(g[u].y, a[u].y, c[u].y, d.y, gamma, eps, acc);
                    step_elem<STEP, CLIP, DELTA, NORMS>(g[u].z, a[u].z, c[u].z, d.z, gamma, eps, acc);
                    step_elem<STEP, CLIP, DELTA, NORMS>(g[u].w, a[u].w, c[u].w, d.w, gamma, eps, acc);
                } else {
                    step_elem<STEP, CLIP, DELTA, NORMS>(g[u], a[u], c[u], d, gamma, eps, acc);
                }
                a_v[i] = a[u];                         // next consumer is the tail's first conv: keep in L2
                if (DELTA) st_stream(d_v + i, d);
            }
        }
    }

    if constexpr (NORMS) finish_sample_norms(acc, partials, counters, s, k, chunks, norms_out, norms_out + n_samples);
}

// ---- L2 mode (a5 l2ball_proj, a5b L2-normalised step) ---------------------------------------------
// per-sample ||a - b||_2 (b nullable -> ||a||_2); same grid / epilogue as the step kernel
template <int VEC, bool DIFF>
__global__ void __launch_bounds__(kThreads)
sample_l2norm_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out,
                     float* partials, unsigned int* counters, long long pv) {
    using V = typename Vec<VEC>::type;
    const int chunks = gridDim.x, k = blockIdx.x, s = blockIdx.y;
    const long long lo = pv * k / chunks, hi = pv * (k + 1) / chunks, base = static_cast<long long>(s) * pv;
    const V* a_v = reinterpret_cast<const V*>(a) + base;
    const V* b_v = DIFF ? reinterpret_cast<const V*>(b) + base : nullptr;
    NormAcc acc;
    for (long long i0 = lo + threadIdx.x; i0 < hi; i0 += static_cast<long long>(kThreads) * kUnroll) {
        V x[kUnroll] = {}, y[kUnroll] = {};
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + static_cast<long long>(u) * kThreads;
            if (i < hi) { x[u] = a_v[i]; if (DIFF) y[u] = b_v[i]; }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + static_cast<long long>(u) * kThreads;
            if (i < hi) {
                if constexpr (VEC == 4) {
                    acc.add(DIFF ? __fsub_rn(x[u].x, y[u].x) : x[u].x); acc.add(DIFF ? __fsub_rn(x[u].y, y[u].y) : x[u].y);
                    acc.add(DIFF ? __fsub_rn(x[u].z, y[u].z) : x[u].z); acc.add(DIFF ? __fsub_rn(x[u].w, y[u].w) : x[u].w);
                } else {
                    acc.add(DIFF ? __fsub_rn(x[u], y[u]) : x[u]);
                }
            }
        }
    }
    finish_sample_norms(acc, partials, counters, s, k, chunks, out, nullptr);
}

// MODE 0: x_adv += (gamma / max(||g||, tiny)) * g            (a5b; norm[s] = ||g_s||_2)
// MODE 1: d = t - c; d /= dist; d *= min(dist, radius); t = c + d   (a5, attack_algo.py:27-32; norm[s] = dist)
template <int VEC, int MODE, bool DELTA>
__global__ void __launch_bounds__(kThreads)
l2_apply_kernel(const float* __restrict__ src, const float* __restrict__ norm, float* t, float* __restrict__ delta_out,
                long long pv, float scalar, float tiny) {
    using V = typename Vec<VEC>::type;
    const int chunks = gridDim.x, k = blockIdx.x, s = blockIdx.y;
    const long long lo = pv * k / chunks, hi = pv * (k + 1) / chunks, base = static_cast<long long>(s) * pv;
    const V* s_v = reinterpret_cast<const V*>(src) + base;
    V* t_v = reinterpret_cast<V*>(t) + base;
    V* d_v = DELTA ? reinterpret_cast<V*>(delta_out) + base : nullptr;
    float nrm = __ldg(norm + s), coef;
    if (MODE == 0) {
        if (!(nrm > tiny)) nrm = tiny;
        coef = __fdiv_rn(scalar, nrm);                 // gamma / ||g||
    } else {
        coef = (nrm > scalar) ? scalar : nrm;          // dist[dist > radius] = radius (NaN stays)
    }
    auto one = [&](float sv, float tv, float& dlt) {
        if (MODE == 0) return __fadd_rn(tv, __fmul_rn(coef, sv));
        float d = __fsub_rn(tv, sv);
        d = __fdiv_rn(d, nrm);
        d = __fmul_rn(d, coef);
        const float o = __fadd_rn(sv, d);
        dlt = __fsub_rn(o, sv);                        // delta as the caller sees it: fl(t_new - center)
        return o;
    };
    for (long long i0 = lo + threadIdx.x; i0 < hi; i0 += static_cast<long long>(kThreads) * kUnroll) {
        V a[kUnroll] = {}, b[kUnroll] = {};
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + static_cast<long long>(u) * kThreads;
            if (i < hi) { a[u] = (MODE == 0) ? ld_stream(s_v + i) : ld_ro(s_v + i); b[u] = t_v[i]; }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + static_cast<long long>(u) * kThreads;
            if (i < hi) {
                V o, d = {};
                if constexpr (VEC == 4) {
                    o.x = one(a[u].x, b[u].x, d.x); o.y = one(a[u].y, b[u].y, d.y);
                    o.z = one(a[u].z, b[u].z, d.z); o.w = one(a[u].w, b[u].w, d.w);
                } else {
                    o = one(a[u], b[u], d);
                }
                t_v[i] = o;
                if (DELTA) st_stream(d_v + i, d);
            }
        }
    }
}

// ---- random start ------------------------------------------------------------------------
__device__ __forceinline__ float init_elem(float x, float u, float eps) {
    // x + fl(fl(fl(2u) - 1) * eps), attack_algo.py:44
    return __fadd_rn(x, __fmul_rn(__fsub_rn(__fmul_rn(2.0f, u), 1.0f), eps));
}

template <int VEC>
__global__ void __launch_bounds__(kThreads)
pgd_init_noise_kernel(const float* __restrict__ x, const float* __restrict__ u, float* __restrict__ x_adv,
                      long long nv, float eps) {
    using V = typename Vec<VEC>::type;
    const V* x_v = reinterpret_cast<const V*>(x);
    const V* u_v = reinterpret_cast<const V*>(u);
    V* o_v = reinterpret_cast<V*>(x_adv);
    const long long stride = static_cast<long long>(gridDim.x) * kThreads;
    for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < nv; i += stride) {
        V a = ld_ro(x_v + i), r = ld_stream(u_v + i), o;
        if constexpr (VEC == 4) {
            o.x = init_elem(a.x, r.x, eps); o.y = init_elem(a.y, r.y, eps);
            o.z = init_elem(a.z, r.z, eps); o.w = init_elem(a.w, r.w, eps);
        } else {
            o = init_elem(a, r, eps);
        }
        o_v[i] = o;
    }
}

struct Philox {
    static constexpr unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
    __device__ __forceinline__ static uint4 draw(unsigned long long ctr, unsigned long long seed) {
        unsigned int c0 = static_cast<unsigned int>(ctr), c1 = static_cast<unsigned int>(ctr >> 32), c2 = 0, c3 = 0;
        unsigned int k0 = static_cast<unsigned int>(seed), k1 = static_cast<unsigned int>(seed >> 32);
#pragma unroll
        for (int r = 0; r < 10; ++r) {
            if (r) { k0 += W0; k1 += W1; }
            const unsigned int h0 = __umulhi(M0, c0), l0 = M0 * c0, h1 = __umulhi(M1, c2), l1 = M1 * c2;
            c0 = h1 ^ c1 ^ k0; c1 = l1; c2 = h0 ^ c3 ^ k1; c3 = l0;
        }
        return make_uint4(c0, c1, c2, c3);
    }
    __device__ __forceinline__ static float uniform(unsigned int bits) {
        return static_cast<float>(bits >> 8) * 0x1p-24f;   // [0,1) on the 2^-24 grid, like torch.rand
    }
};

// Each thread owns one Philox counter = 4 consecutive elements (one float4 when aligned).
template <int VEC>
__global__ void __launch_bounds__(kThreads)
pgd_init_philox_kernel(const float* __restrict__ x, float* __restrict__ x_adv, long long n, float eps,
                       unsigned long long seed, unsigned long long offset,
                       const unsigned long long* __restrict__ offset_device) {
    if (offset_device) offset += __ldg(offset_device);
    const long long n4 = (n + 3) / 4;
    const long long stride = static_cast<long long>(gridDim.x) * kThreads;
    for (long long q = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; q < n4; q += stride) {
        const uint4 r = Philox::draw(static_cast<unsigned long long>(q) + offset, seed);
        const float u[4] = {Philox::uniform(r.x), Philox::uniform(r.y), Philox::uniform(r.z), Philox::uniform(r.w)};
        if (VEC == 4 && q * 4 + 3 < n) {
            const float4 a = ld_ro(reinterpret_cast<const float4*>(x) + q);
            float4 o;
            o.x = init_elem(a.x, u[0], eps); o.y = init_elem(a.y, u[1], eps);
            o.z = init_elem(a.z, u[2], eps); o.w = init_elem(a.w, u[3], eps);
            reinterpret_cast<float4*>(x_adv)[q] = o;
        } else {
            for (int j = 0; j < 4 && q * 4 + j < n; ++j) x_adv[q * 4 + j] = init_elem(x[q * 4 + j], u[j], eps);
        }
    }
}


// ---- bf16 I/O twins (BASELINE config 3: bf16 feature maps; fp32 math in registers) -----------------------------
// Storage is bf16 (8 elements per 128-bit access: half the bytes of the fp32 kernels), arithmetic is the SAME fp32
// sequence as above on the widened values, results rounded to nearest-even on store.  delta / norms are formed from
// the ROUNDED x_adv, i.e. what the caller would see by subtracting the stored tensors.  No reference exists for
// this dtype ("parity unpinned"); the oracle twin is orc_pgd_linf_step_bf16 and the contract is bit-equality with it.
struct Bf16x8 { uint4 raw; };
__device__ __forceinline__ void unpack8(const uint4& r, float (&f)[8]) {
    const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) { const float2 v = __bfloat1622float2(h[i]); f[2 * i] = v.x; f[2 * i + 1] = v.y; }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    uint4 r;
    __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&r);
#pragma unroll
    for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return r;
}
__device__ __forceinline__ float bf16_round(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

template <bool VEC8, bool STEP, bool CLIP, bool DELTA, bool NORMS>
__global__ void __launch_bounds__(kThreads)
pgd_linf_step_bf16_kernel(const __nv_bfloat16* __restrict__ grad, const __nv_bfloat16* __restrict__ x_clean,
                          __nv_bfloat16* x_adv, __nv_bfloat16* __restrict__ delta_out, float* __restrict__ norms_out,
                          float* partials, unsigned int* counters, long long pv, int n_samples, float gamma, float eps) {
    constexpr int NE = VEC8 ? 8 : 1;
    constexpr bool NEED_CLEAN = CLIP || DELTA || NORMS;
    const int chunks = gridDim.x, k = blockIdx.x, s = blockIdx.y;
    const long long lo = pv * k / chunks, hi = pv * (k + 1) / chunks, base = static_cast<long long>(s) * pv * NE;
    NormAcc acc;
    for (long long i0 = lo + threadIdx.x; i0 < hi; i0 += static_cast<long long>(kThreads) * kUnroll) {
        float g[kUnroll][NE] = {}, a[kUnroll][NE] = {}, c[kUnroll][NE] = {};
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + static_cast<long long>(u) * kThreads;
            if (i < hi) {
                if constexpr (VEC8) {
                    if (STEP) unpack8(__ldcs(reinterpret_cast<const uint4*>(grad + base) + i), g[u]);
                    unpack8(reinterpret_cast<const uint4*>(x_adv + base)[i], a[u]);
                    if (NEED_CLEAN) unpack8(__ldg(reinterpret_cast<const uint4*>(x_clean + base) + i), c[u]);
                } else {
                    if (STEP) g[u][0] = __bfloat162float(grad[base + i]);
                    a[u][0] = __bfloat162float(x_adv[base + i]);
                    if (NEED_CLEAN) c[u][0] = __bfloat162float(x_clean[base + i]);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < kUnroll; ++u) {
            const long long i = i0 + static_cast<long long>(u) * kThreads;
            if (i < hi) {
                float d[NE];
#pragma unroll
                for (int e = 0; e < NE; ++e) {
                    a[u][e] = bf16_round(linf_update<STEP, CLIP>(g[u][e], a[u][e], c[u][e], gamma, eps));
                    if (DELTA || NORMS) {
                        d[e] = __fsub_rn(a[u][e], c[u][e]);
                        if (NORMS) acc.add(d[e]);
                    }
                }
                if constexpr (VEC8) {
                    reinterpret_cast<uint4*>(x_adv + base)[i] = pack8(a[u]);
                    if (DELTA) __stcs(reinterpret_cast<uint4*>(delta_out + base) + i, pack8(d));
                } else {
                    x_adv[base + i] = __float2bfloat16_rn(a[u][0]);
                    if (DELTA) delta_out[base + i] = __float2bfloat16_rn(d[0]);
                }
            }
        }
    }
    if constexpr (NORMS) finish_sample_norms(acc, partials, counters, s, k, chunks, norms_out, norms_out + n_samples);
}

// random start, bf16 storage: u is either the caller's fp32 torch.rand draw or Philox (same stream as the fp32 kernel)
template <bool PHILOX>
__global__ void __launch_bounds__(kThreads)
pgd_init_bf16_kernel(const __nv_bfloat16* __restrict__ x, const float* __restrict__ u, __nv_bfloat16* __restrict__ x_adv,
                     long long n, float eps, unsigned long long seed, unsigned long long offset,
                     const unsigned long long* __restrict__ offset_device) {
    if (PHILOX && offset_device) offset += __ldg(offset_device);
    const long long n4 = (n + 3) / 4, stride = static_cast<long long>(gridDim.x) * kThreads;
    for (long long q = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; q < n4; q += stride) {
        float r[4];
        if (PHILOX) {
            const uint4 b = Philox::draw(static_cast<unsigned long long>(q) + offset, seed);
            r[0] = Philox::uniform(b.x); r[1] = Philox::uniform(b.y); r[2] = Philox::uniform(b.z); r[3] = Philox::uniform(b.w);
        }
        for (int j = 0; j < 4 && q * 4 + j < n; ++j) {
            const float uu = PHILOX ? r[j] : __ldcs(u + q * 4 + j);
            x_adv[q * 4 + j] = __float2bfloat16_rn(init_elem(__bfloat162float(x[q * 4 + j]), uu, eps));
        }
    }
}

// tensor_clamp with free-form tensor bounds, Classification/attack_algo.py:9-19:
//   idx = t < min; t[idx] = min[idx]; idx = t > max; t[idx] = max[idx]     (NaN compares false -> left alone)
template <int VEC>
__global__ void __launch_bounds__(kThreads)
tensor_clamp_kernel(float* t, const float* __restrict__ lo, const float* __restrict__ hi, long long nv) {
    using V = typename Vec<VEC>::type;
    const long long stride = static_cast<long long>(gridDim.x) * kThreads;
    auto one = [](float v, float l, float h) { v = (v < l) ? l : v; return (v > h) ? h : v; };
    for (long long i = static_cast<long long>(blockIdx.x) * kThreads + threadIdx.x; i < nv; i += stride) {
        V v = reinterpret_cast<V*>(t)[i];
        const V l = ld_stream(reinterpret_cast<const V*>(lo) + i), h = ld_stream(reinterpret_cast<const V*>(hi) + i);
        if constexpr (VEC == 4) { v.x = one(v.x, l.x, h.x); v.y = one(v.y, l.y, h.y); v.z = one(v.z, l.z, h.z); v.w = one(v.w, l.w, h.w); }
        else { v = one(v, l, h); }
        reinterpret_cast<V*>(t)[i] = v;
    }
}

// ---- host side -----------------------------------------------------------------------------
inline int flat_grid(long long n_vec, int per_cta) {
    const long long want = (n_vec + per_cta - 1) / per_cta;
    const long long cap = static_cast<long long>(sm_count()) * kCtasPerSm;
    return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

template <int VEC, bool STEP, bool CLIP, bool DELTA, bool NORMS>
int launch_step(const float* grad, const float* x_clean, float* x_adv, float* delta_out, float* norms_out,
                float* partials, unsigned int* counters, long long n_samples, long long per_sample, int chunks,
                float gamma, float eps, cudaStream_t st) {
    dim3 grid(chunks, static_cast<unsigned int>(n_samples));
    return launch_pdl(pgd_linf_step_kernel<VEC, STEP, CLIP, DELTA, NORMS>, grid, dim3(kThreads), 0, st, grad, x_clean, x_adv,
                      delta_out, norms_out, partials, counters, static_cast<long long>(per_sample / VEC),
                      static_cast<int>(n_samples), gamma, eps);
}

template <int VEC>
int dispatch_step(bool step, bool clip, bool delta, bool norms, const float* grad, const float* x_clean, float* x_adv,
                  float* delta_out, float* norms_out, float* partials, unsigned int* counters, long long ns,
                  long long per, int chunks, float gamma, float eps, cudaStream_t st) {
#define AFAN_CASE(C, D, N)                                                                                 \
    if (clip == C && delta == D && norms == N)                                                             \
        return step ? launch_step<VEC, true, C, D, N>(grad, x_clean, x_adv, delta_out, norms_out, partials, counters, \
                                                      ns, per, chunks, gamma, eps, st)                     \
                    : launch_step<VEC, false, C, D, N>(grad, x_clean, x_adv, delta_out, norms_out, partials, counters, \
                                                       ns, per, chunks, gamma, eps, st);
    AFAN_CASE(false, false, false) AFAN_CASE(false, false, true) AFAN_CASE(false, true, false)
    AFAN_CASE(false, true, true) AFAN_CASE(true, false, false) AFAN_CASE(true, false, true)
    AFAN_CASE(true, true, false) AFAN_CASE(true, true, true)
#undef AFAN_CASE
    return AFAN_ERR_UNSUPPORTED;
}

constexpr int kMaxNormChunks = 1184;   // 148 SMs x 8 CTAs: upper bound of chunks per sample

}  // namespace afan

using namespace afan;

AFAN_EXPORT int64_t afan_pgd_norms_workspace_bytes(int64_t n_samples) {
    if (n_samples < 0) return AFAN_ERR_SIZE;
    // counters [n_samples] (uint32, padded to 16 B) + partials [n_samples][chunks][4] floats;
    // chunks * n_samples <= kMaxNormChunks + n_samples by construction (see afan_pgd_linf_step_f32)
    const int64_t counters = ((n_samples * 4 + 15) / 16) * 16;
    return counters + (kMaxNormChunks + n_samples) * 16;
}

AFAN_EXPORT int afan_pgd_linf_step_f32(const float* grad, const float* x_clean, float* x_adv, float* delta_out,
                                       float* norms_out, void* workspace, int64_t workspace_bytes,
                                       int64_t n_samples, int64_t per_sample, float gamma, float eps, int clip,
                                       afan_stream_t stream) {
    if (n_samples < 0 || per_sample < 0) return AFAN_ERR_SIZE;
    if (n_samples == 0 || per_sample == 0) return AFAN_OK;
    const bool delta = delta_out != nullptr, norms = norms_out != nullptr;
    if (!x_adv || ((clip || delta || norms) && !x_clean)) return AFAN_ERR_NULL;
    const bool step = grad != nullptr;                 // grad == NULL: projection (+delta/norms) only
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    long long ns = n_samples, per = per_sample;
    float* partials = nullptr;
    unsigned int* counters = nullptr;
    if (norms) {
        if (ns > 65535) return AFAN_ERR_UNSUPPORTED;
        if (!workspace || workspace_bytes < afan_pgd_norms_workspace_bytes(n_samples)) return AFAN_ERR_WORKSPACE;
        counters = static_cast<unsigned int*>(workspace);
        partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + ((ns * 4 + 15) / 16) * 16);
    } else {
        per = ns * per;                // purely elementwise: fold into one "sample"
        ns = 1;
    }
    const bool vec = (per % 4 == 0) && (!grad || aligned16(grad)) && aligned16(x_adv) && (!x_clean || aligned16(x_clean)) &&
                     (!delta_out || aligned16(delta_out));
    const long long pv = vec ? per / 4 : per;
    // chunks per sample: enough CTAs to fill the chip (148 x 8), at most one CTA per 256*kUnroll vectors
    long long chunks = (pv + kThreads * kUnroll - 1) / (kThreads * kUnroll);
    const long long sms = sm_count();
    const long long cap = (sms * kCtasPerSm + ns - 1) / ns;
    if (ns == 1 && chunks > sms) chunks = (chunks + sms - 1) / sms * sms;   // whole waves: every SM gets the same CTA count
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    if (norms && chunks * ns > kMaxNormChunks + ns) chunks = (kMaxNormChunks + ns) / ns;
    // many small samples (config 2: 128 x 16 Ki elements): ONE CTA per sample already puts a CTA on most SMs, and its
    // norms then need no cross-CTA fold at all (no fence, no ticket, no partials round trip: 11.0 -> ~7 us at 128x16x32x32)
    if (norms && ns * 2 >= sms && pv <= static_cast<long long>(kThreads) * kUnroll * 8) chunks = 1;
    return vec ? dispatch_step<4>(step, clip != 0, delta, norms, grad, x_clean, x_adv, delta_out, norms_out, partials,
                                  counters, ns, per, static_cast<int>(chunks), gamma, eps, st)
               : dispatch_step<1>(step, clip != 0, delta, norms, grad, x_clean, x_adv, delta_out, norms_out, partials,
                                  counters, ns, per, static_cast<int>(chunks), gamma, eps, st);
}

AFAN_EXPORT int afan_pgd_init_noise_f32(const float* x, const float* u, float* x_adv, int64_t n_elem, float eps,
                                        afan_stream_t stream) {
    if (n_elem < 0) return AFAN_ERR_SIZE;
    if (n_elem == 0) return AFAN_OK;
    if (!x || !u || !x_adv) return AFAN_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_elem % 4 == 0 && aligned16(x) && aligned16(u) && aligned16(x_adv)) {
        const long long nv = n_elem / 4;
        pgd_init_noise_kernel<4><<<flat_grid(nv, kThreads * 2), kThreads, 0, st>>>(x, u, x_adv, nv, eps);
    } else {
        pgd_init_noise_kernel<1><<<flat_grid(n_elem, kThreads * 2), kThreads, 0, st>>>(x, u, x_adv, n_elem, eps);
    }
    return launch_status();
}

AFAN_EXPORT int afan_pgd_init_philox_f32(const float* x, float* x_adv, int64_t n_elem, float eps, uint64_t seed,
                                         uint64_t offset, const uint64_t* offset_device, afan_stream_t stream) {
    if (n_elem < 0) return AFAN_ERR_SIZE;
    if (n_elem == 0) return AFAN_OK;
    if (!x || !x_adv) return AFAN_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const long long n4 = (n_elem + 3) / 4;
    if (aligned16(x) && aligned16(x_adv))
        pgd_init_philox_kernel<4><<<flat_grid(n4, kThreads * 2), kThreads, 0, st>>>(
            x, x_adv, n_elem, eps, seed, offset, reinterpret_cast<const unsigned long long*>(offset_device));
    else
        pgd_init_philox_kernel<1><<<flat_grid(n4, kThreads * 2), kThreads, 0, st>>>(
            x, x_adv, n_elem, eps, seed, offset, reinterpret_cast<const unsigned long long*>(offset_device));
    return launch_status();
}

// ---- L2 entry points --------------------------------------------------------------------------
namespace afan {
struct SampleGrid { long long ns, per, pv; int chunks; bool vec; float* partials; unsigned int* counters; };
inline int sample_grid(SampleGrid& g, int64_t n_samples, int64_t per_sample, bool aligned, void* ws, int64_t ws_bytes,
                       bool need_ws) {
    if (n_samples > 65535) return AFAN_ERR_UNSUPPORTED;
    g.ns = n_samples; g.per = per_sample;
    g.vec = aligned && (per_sample % 4 == 0);
    g.pv = g.vec ? per_sample / 4 : per_sample;
    long long chunks = (g.pv + kThreads * kUnroll - 1) / (kThreads * kUnroll);
    const long long cap = (static_cast<long long>(sm_count()) * kCtasPerSm + g.ns - 1) / g.ns;
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    g.chunks = static_cast<int>(chunks);
    g.partials = nullptr; g.counters = nullptr;
    if (need_ws) {
        if (!ws || ws_bytes < afan_pgd_norms_workspace_bytes(n_samples)) return AFAN_ERR_WORKSPACE;
        g.counters = static_cast<unsigned int*>(ws);
        g.partials = reinterpret_cast<float*>(static_cast<char*>(ws) + ((g.ns * 4 + 15) / 16) * 16);
    }
    return AFAN_OK;
}
}  // namespace afan

AFAN_EXPORT int afan_sample_l2norm_f32(const float* a, const float* b, float* out_norm, void* workspace,
                                       int64_t workspace_bytes, int64_t n_samples, int64_t per_sample,
                                       afan_stream_t stream) {
    if (n_samples < 0 || per_sample < 0) return AFAN_ERR_SIZE;
    if (n_samples == 0) return AFAN_OK;
    if (!a || !out_norm) return AFAN_ERR_NULL;
    SampleGrid g;
    int rc = sample_grid(g, n_samples, per_sample, aligned16(a) && (!b || aligned16(b)), workspace, workspace_bytes, true);
    if (rc != AFAN_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(g.chunks, static_cast<unsigned int>(g.ns));
    if (g.vec) { if (b) sample_l2norm_kernel<4, true><<<grid, kThreads, 0, st>>>(a, b, out_norm, g.partials, g.counters, g.pv);
                 else   sample_l2norm_kernel<4, false><<<grid, kThreads, 0, st>>>(a, b, out_norm, g.partials, g.counters, g.pv); }
    else       { if (b) sample_l2norm_kernel<1, true><<<grid, kThreads, 0, st>>>(a, b, out_norm, g.partials, g.counters, g.pv);
                 else   sample_l2norm_kernel<1, false><<<grid, kThreads, 0, st>>>(a, b, out_norm, g.partials, g.counters, g.pv); }
    return launch_status();
}

AFAN_EXPORT int afan_pgd_l2_step_f32(const float* grad, const float* grad_norm, float* x_adv, int64_t n_samples,
                                     int64_t per_sample, float gamma, float tiny, afan_stream_t stream) {
    if (n_samples < 0 || per_sample < 0) return AFAN_ERR_SIZE;
    if (n_samples == 0 || per_sample == 0) return AFAN_OK;
    if (!grad || !grad_norm || !x_adv) return AFAN_ERR_NULL;
    SampleGrid g;
    int rc = sample_grid(g, n_samples, per_sample, aligned16(grad) && aligned16(x_adv), nullptr, 0, false);
    if (rc != AFAN_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(g.chunks, static_cast<unsigned int>(g.ns));
    if (g.vec) l2_apply_kernel<4, 0, false><<<grid, kThreads, 0, st>>>(grad, grad_norm, x_adv, nullptr, g.pv, gamma, tiny);
    else       l2_apply_kernel<1, 0, false><<<grid, kThreads, 0, st>>>(grad, grad_norm, x_adv, nullptr, g.pv, gamma, tiny);
    return launch_status();
}

AFAN_EXPORT int afan_l2ball_proj_f32(const float* center, const float* dist, float* t, float* delta_out,
                                     int64_t n_samples, int64_t per_sample, float radius, afan_stream_t stream) {
    if (n_samples < 0 || per_sample < 0) return AFAN_ERR_SIZE;
    if (n_samples == 0 || per_sample == 0) return AFAN_OK;
    if (!center || !dist || !t) return AFAN_ERR_NULL;
    SampleGrid g;
    int rc = sample_grid(g, n_samples, per_sample, aligned16(center) && aligned16(t) && (!delta_out || aligned16(delta_out)),
                         nullptr, 0, false);
    if (rc != AFAN_OK) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    dim3 grid(g.chunks, static_cast<unsigned int>(g.ns));
    if (g.vec) { if (delta_out) l2_apply_kernel<4, 1, true><<<grid, kThreads, 0, st>>>(center, dist, t, delta_out, g.pv, radius, 0.f);
                 else           l2_apply_kernel<4, 1, false><<<grid, kThreads, 0, st>>>(center, dist, t, nullptr, g.pv, radius, 0.f); }
    else       { if (delta_out) l2_apply_kernel<1, 1, true><<<grid, kThreads, 0, st>>>(center, dist, t, delta_out, g.pv, radius, 0.f);
                 else           l2_apply_kernel<1, 1, false><<<grid, kThreads, 0, st>>>(center, dist, t, nullptr, g.pv, radius, 0.f); }
    return launch_status();
}

// ---- bf16 entry points -------------------------------------------------------------------------------------------
namespace afan {
template <bool VEC8>
int dispatch_step_bf16(bool step, bool clip, bool delta, bool norms, const __nv_bfloat16* g, const __nv_bfloat16* xc,
                       __nv_bfloat16* xa, __nv_bfloat16* d, float* norms_out, float* partials, unsigned int* counters,
                       long long ns, long long pv, int chunks, float gamma, float eps, cudaStream_t st) {
    dim3 grid(chunks, static_cast<unsigned int>(ns));
#define AFAN_B16(S, C, D, N)                                                                                          \
    if (step == S && clip == C && delta == D && norms == N) {                                                         \
        pgd_linf_step_bf16_kernel<VEC8, S, C, D, N><<<grid, kThreads, 0, st>>>(g, xc, xa, d, norms_out, partials, counters, \
                                                                              pv, static_cast<int>(ns), gamma, eps); \
        return launch_status();                                                                                       \
    }
    AFAN_B16(true, false, false, false) AFAN_B16(true, false, false, true) AFAN_B16(true, false, true, false)
    AFAN_B16(true, false, true, true) AFAN_B16(true, true, false, false) AFAN_B16(true, true, false, true)
    AFAN_B16(true, true, true, false) AFAN_B16(true, true, true, true) AFAN_B16(false, false, false, false)
    AFAN_B16(false, false, false, true) AFAN_B16(false, false, true, false) AFAN_B16(false, false, true, true)
    AFAN_B16(false, true, false, false) AFAN_B16(false, true, false, true) AFAN_B16(false, true, true, false)
    AFAN_B16(false, true, true, true)
#undef AFAN_B16
    return AFAN_ERR_UNSUPPORTED;
}
}  // namespace afan

AFAN_EXPORT int afan_pgd_linf_step_bf16(const void* grad, const void* x_clean, void* x_adv, void* delta_out,
                                        float* norms_out, void* workspace, int64_t workspace_bytes, int64_t n_samples,
                                        int64_t per_sample, float gamma, float eps, int clip, afan_stream_t stream) {
    if (n_samples < 0 || per_sample < 0) return AFAN_ERR_SIZE;
    if (n_samples == 0 || per_sample == 0) return AFAN_OK;
    const bool delta = delta_out != nullptr, norms = norms_out != nullptr, step = grad != nullptr;
    if (!x_adv || ((clip || delta || norms) && !x_clean)) return AFAN_ERR_NULL;
    long long ns = n_samples, per = per_sample;
    float* partials = nullptr;
    unsigned int* counters = nullptr;
    if (norms) {
        if (ns > 65535) return AFAN_ERR_UNSUPPORTED;
        if (!workspace || workspace_bytes < afan_pgd_norms_workspace_bytes(n_samples)) return AFAN_ERR_WORKSPACE;
        counters = static_cast<unsigned int*>(workspace);
        partials = reinterpret_cast<float*>(static_cast<char*>(workspace) + ((ns * 4 + 15) / 16) * 16);
    } else {
        per = ns * per;
        ns = 1;
    }
    const bool vec = (per % 8 == 0) && (!grad || aligned16(grad)) && aligned16(x_adv) && (!x_clean || aligned16(x_clean)) &&
                     (!delta_out || aligned16(delta_out));
    const long long pv = vec ? per / 8 : per;
    long long chunks = (pv + kThreads * kUnroll - 1) / (kThreads * kUnroll);
    const long long sms = sm_count(), cap = (sms * kCtasPerSm + ns - 1) / ns;
    if (ns == 1 && chunks > sms) chunks = (chunks + sms - 1) / sms * sms;
    if (chunks > cap) chunks = cap;
    if (chunks < 1) chunks = 1;
    if (norms && chunks * ns > kMaxNormChunks + ns) chunks = (kMaxNormChunks + ns) / ns;
    const __nv_bfloat16* g = static_cast<const __nv_bfloat16*>(grad);
    const __nv_bfloat16* xc = static_cast<const __nv_bfloat16*>(x_clean);
    __nv_bfloat16* xa = static_cast<__nv_bfloat16*>(x_adv);
    __nv_bfloat16* d = static_cast<__nv_bfloat16*>(delta_out);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return vec ? dispatch_step_bf16<true>(step, clip != 0, delta, norms, g, xc, xa, d, norms_out, partials, counters, ns, pv,
                                          static_cast<int>(chunks), gamma, eps, st)
               : dispatch_step_bf16<false>(step, clip != 0, delta, norms, g, xc, xa, d, norms_out, partials, counters, ns, pv,
                                           static_cast<int>(chunks), gamma, eps, st);
}

AFAN_EXPORT int afan_pgd_init_noise_bf16(const void* x, const float* u, void* x_adv, int64_t n_elem, float eps,
                                         afan_stream_t stream) {
    if (n_elem < 0) return AFAN_ERR_SIZE;
    if (n_elem == 0) return AFAN_OK;
    if (!x || !u || !x_adv) return AFAN_ERR_NULL;
    pgd_init_bf16_kernel<false><<<flat_grid((n_elem + 3) / 4, kThreads * 2), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), u, static_cast<__nv_bfloat16*>(x_adv), n_elem, eps, 0ULL, 0ULL, nullptr);
    return launch_status();
}

AFAN_EXPORT int afan_pgd_init_philox_bf16(const void* x, void* x_adv, int64_t n_elem, float eps, uint64_t seed,
                                          uint64_t offset, const uint64_t* offset_device, afan_stream_t stream) {
    if (n_elem < 0) return AFAN_ERR_SIZE;
    if (n_elem == 0) return AFAN_OK;
    if (!x || !x_adv) return AFAN_ERR_NULL;
    pgd_init_bf16_kernel<true><<<flat_grid((n_elem + 3) / 4, kThreads * 2), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(
        static_cast<const __nv_bfloat16*>(x), nullptr, static_cast<__nv_bfloat16*>(x_adv), n_elem, eps, seed, offset,
        reinterpret_cast<const unsigned long long*>(offset_device));
    return launch_status();
}

AFAN_EXPORT int afan_tensor_clamp_f32(float* t, const float* min, const float* max, int64_t n_elem, afan_stream_t stream) {
    if (n_elem < 0) return AFAN_ERR_SIZE;
    if (n_elem == 0) return AFAN_OK;
    if (!t || !min || !max) return AFAN_ERR_NULL;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n_elem % 4 == 0 && aligned16(t) && aligned16(min) && aligned16(max))
        tensor_clamp_kernel<4><<<flat_grid(n_elem / 4, kThreads * 2), kThreads, 0, st>>>(t, min, max, n_elem / 4);
    else
        tensor_clamp_kernel<1><<<flat_grid(n_elem, kThreads * 2), kThreads, 0, st>>>(t, min, max, n_elem);
    return launch_status();
}
