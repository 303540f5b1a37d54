// afan_bn.cu -- dual (grouped-statistics) train-mode BatchNorm2d for sm_100a, forward + backward,
// with normalise + affine + residual-add + ReLU fused.
//
// Replaces nn.BatchNorm2d / F.relu / `out += shortcut` of the reference tail,
// Classification/resnet_s.py:54,56,70-76,89, as driven by main_perturb.py:195-196 where the SAME
// module sees the adversarial and then the clean batch: here both halves of a [adv; clean] batch
// are swept together, each with its own statistics ("groups"), shared affine and running averages.
//
// Layout: x [G*N][C][HW] contiguous fp32 (NCHW).  Per (group, channel) the statistic domain is N
// planes of HW contiguous elements.
//   pass 1 (bn_reduce_kernel): grid (S, G*C).  CTA (s, gc) streams slice s of the domain with 128-bit
//     loads, reduces with warp shuffles + a shared-memory tree, writes one double2 partial.  The CTA
//     that arrives last for a channel folds the partials of all its groups in a fixed order
//     (deterministic; no float atomics) and finalises: mean/invstd, running-stat update in pass
//     order, and the per-(g,c) coefficient table used by pass 2.
//   pass 2 (bn_apply_kernel): flat streaming sweep, y = relu(x*scale + shift (+res)).
// HBM bytes per element: fwd 8 algorithmic (read x, write y); the second read of x in pass 2 hits
// the 126 MB L2 for every shape in BASELINE.json's configs.  bwd: 12 algorithmic (+4 for y if relu).
#include <algorithm>
#include <cooperative_groups.h>

#include <cstdlib>
#include <type_traits>

#include "afan_common.cuh"
#include "afan_p2p.cuh"

namespace cg = cooperative_groups;

namespace afan {

constexpr int kBnUnroll = 4;

constexpr int kSplitWaves = 4;
struct BnLayout {            // workspace carve-up (bytes), shared by every BN entry point
    int64_t counters, table, coef, partials, total;
};
__host__ inline int64_t align256(int64_t v) { return (v + 255) / 256 * 256; }
__host__ inline BnLayout bn_layout(int64_t groups, int64_t c) {
    BnLayout l;
    l.counters = 0;
    l.table = align256(c * 4);
    l.coef = l.table + align256(groups * c * 8);
    l.partials = l.coef + align256(groups * c * 16);
    l.total = l.partials + (kSplitWaves * static_cast<int64_t>(sm_count()) * kCtasPerSm + groups * c) * 16;   // room for kSplitWaves waves of CTAs
    return l;
}

struct ReduceParams {
    const float* a;          // fwd: x            bwd: dy
    const float* b;          // fwd: unused       bwd: x
    const float* y;          // bwd + relu: forward output (mask = y > 0)
    const float* save_mean;  // bwd
    const float* save_invstd;
    double2* partials;
    unsigned int* counters;
    double* sums_out;        // nullable: [G][C][2] local sums for the NCCL path
    // finalisation (do_finalize != 0)
    int do_finalize;
    const float* weight;
    const float* bias;
    float* running_mean;
    float* running_var;
    float* out_mean;         // fwd: save_mean
    float* out_invstd;       // fwd: save_invstd
    float2* table;           // fwd: (scale, shift) per (g,c)
    float4* coef;            // bwd: (w*invstd, mean(dy), mean(dy*xhat)*invstd, mean) per (g,c)
    float* dweight;          // bwd
    float* dbias;
    double count;            // elements per (g,c) statistic (global count under NCCL sync)
    float eps, momentum;
    int replay;
    unsigned int groups, n, c, hwv, splits;
};

// ---- shifted accumulation --------------------------------------------------------------------------------------
// sum / sum-of-squares in fp32 partials cancels catastrophically when |mean| >> std (found on the DeepLab ASPP pooling
// branch: 2 values per channel, relu-positive, nearly equal -> 2.5e-3 error against the two-pass library BatchNorm).
// Every forward-statistics kernel therefore accumulates (x - K) with K = the first element of the (group, channel) domain
// (one broadcast load), and un-shifts its per-CTA partial in DOUBLE before it is folded / exchanged:
//   sum x = S1 + cnt K,   sum x^2 = S2 + 2 K S1 + cnt K^2          (cnt = elements behind the partial)
// so everything downstream (cluster fold, NVLink / NCCL exchange, finalisers) still sees raw sums.
__device__ __forceinline__ double2 unshift_sums(double s1, double s2, double cnt, double k) {
    return make_double2(s1 + cnt * k, s2 + 2.0 * k * s1 + cnt * k * k);
}

// ---- finalisation math, shared by the in-kernel path and the stand-alone finalize kernels ----
__device__ __forceinline__ void fwd_finalize_channel(const ReduceParams& p, unsigned int ch, const double* sum,
                                                     const double* sumsq) {
    float rm = p.running_mean ? p.running_mean[ch] : 0.f, rv = p.running_var ? p.running_var[ch] : 0.f;
    for (unsigned int g = 0; g < p.groups; ++g) {
        const double mean = sum[g] / p.count;
        double var = sumsq[g] / p.count - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        const double invstd = rsqrt(var + static_cast<double>(p.eps));
        const double unbiased = p.count > 1.0 ? var * (p.count / (p.count - 1.0)) : var;
        for (int r = 0; r < p.replay; ++r) {          // pass order: group 0 (adv) first, then group 1 (clean)
            rm = static_cast<float>((1.0 - p.momentum) * rm + p.momentum * mean);
            rv = static_cast<float>((1.0 - p.momentum) * rv + p.momentum * unbiased);
        }
        const unsigned int gc = g * p.c + ch;
        const float w = p.weight ? p.weight[ch] : 1.f, b = p.bias ? p.bias[ch] : 0.f;
        const float scale = static_cast<float>(w * invstd);
        p.out_mean[gc] = static_cast<float>(mean);
        p.out_invstd[gc] = static_cast<float>(invstd);
        p.table[gc] = make_float2(scale, static_cast<float>(b - mean * w * invstd));
    }
    if (p.running_mean) p.running_mean[ch] = rm;
    if (p.running_var) p.running_var[ch] = rv;
}

__device__ __forceinline__ void bwd_finalize_group(const ReduceParams& p, unsigned int gc, unsigned int ch,
                                                   double s_dy, double s_dyxh) {
    const float w = p.weight ? p.weight[ch] : 1.f;
    const float invstd = p.save_invstd[gc];
    p.coef[gc] = make_float4(w * invstd, static_cast<float>(s_dy / p.count),
                             static_cast<float>(s_dyxh / p.count * invstd), p.save_mean[gc]);
}

// ---- pass 1 -------------------------------------------------------------------------------------
// the last CTA of a channel (all groups, all splits) folds the partials in a fixed order and finalises
template <bool BWD>
__device__ __forceinline__ void reduce_finalize(const ReduceParams& p, unsigned int ch) {
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    constexpr int kMaxGroups = 16;                    // learnable-eta A-FAN batches 9 adversarial groups + clean
    double t0[kMaxGroups], t1[kMaxGroups];
    for (unsigned int gg = 0; gg < p.groups; ++gg) {
        const volatile double* pp = reinterpret_cast<const volatile double*>(p.partials + (static_cast<size_t>(gg) * p.c + ch) * p.splits);
        double x0 = 0.0, x1 = 0.0;
        for (unsigned int k = lane; k < p.splits; k += 32) { x0 += pp[2 * k]; x1 += pp[2 * k + 1]; }
        t0[gg] = warp_sum(x0);
        t1[gg] = warp_sum(x1);
    }
    if (lane != 0) return;
    if (p.sums_out)
        for (unsigned int gg = 0; gg < p.groups; ++gg) {
            p.sums_out[(static_cast<size_t>(gg) * p.c + ch) * 2] = t0[gg];
            p.sums_out[(static_cast<size_t>(gg) * p.c + ch) * 2 + 1] = t1[gg];
        }
    if (BWD) {
        double dw = 0.0, db = 0.0;
        for (unsigned int gg = 0; gg < p.groups; ++gg) { db += t0[gg]; dw += t1[gg]; }
        if (p.dweight) p.dweight[ch] = static_cast<float>(dw);
        if (p.dbias) p.dbias[ch] = static_cast<float>(db);
        if (p.do_finalize)
            for (unsigned int gg = 0; gg < p.groups; ++gg) bwd_finalize_group(p, gg * p.c + ch, ch, t0[gg], t1[gg]);
    } else if (p.do_finalize) {
        fwd_finalize_channel(p, ch, t0, t1);
    }
}

template <bool BWD, bool RELU>
__device__ __forceinline__ void peel_add(float av, float bv, float yv, float mean, float kshift, float& acc0, float& acc1) {
    if (BWD) {
        const float d = (RELU && !(yv > 0.f)) ? 0.f : av;
        acc0 += d;
        acc1 = fmaf(d, bv - mean, acc1);
    } else {
        const float t = av - kshift;
        acc0 += t;
        acc1 = fmaf(t, t, acc1);
    }
}

// Pass 1 for H*W % 4 != 0 (BASELINE config 5's 129 x 129 and 33 x 33 maps): the scalar form of bn_reduce_kernel spends ~40
// instructions per element on index arithmetic and is issue-bound at 0.3 of the HBM roofline (ncu, 8x256x129x129).  Here a
// CTA takes the same fraction [plo, phi) of EVERY plane of its (group, channel) domain; a plane segment starts at an
// arbitrary 4-byte alignment, so it is peeled into a scalar head, a 16-byte-aligned body moved with 128-bit loads and a
// scalar tail -- no division per element, four planes in flight per thread.  Same partials / finalisation as the kernel below.
template <bool BWD, bool RELU>
__global__ void __launch_bounds__(kThreads) bn_reduce_peel_kernel(const ReduceParams p) {
    const unsigned int s = blockIdx.x, gc = blockIdx.y;
    const unsigned int g = gc / p.c, ch = gc - g * p.c;
    const unsigned int hw = p.hwv;                                          // scalar geometry: hwv == H*W
    const unsigned int plo = static_cast<unsigned int>(static_cast<unsigned long long>(hw) * s / p.splits);
    const unsigned int phi = static_cast<unsigned int>(static_cast<unsigned long long>(hw) * (s + 1) / p.splits);
    const unsigned int len = phi - plo;
    const size_t plane0 = (static_cast<size_t>(g) * p.n * p.c + ch) * hw;
    const size_t plane_stride = static_cast<size_t>(p.c) * hw;
    const float mean = BWD ? p.save_mean[gc] : 0.f;
    const float kshift = BWD ? 0.f : __ldg(p.a + plane0);
    float acc0 = 0.f, acc1 = 0.f;
#define add(av, bv, yv) peel_add<BWD, RELU>((av), (bv), (yv), mean, kshift, acc0, acc1)
    constexpr int PL = 4;                                                    // planes in flight
    for (unsigned int n0 = 0; n0 < p.n; n0 += PL) {
        size_t base[PL];
        unsigned int head[PL], nbody[PL], most = 0u;
#pragma unroll
        for (int k = 0; k < PL; ++k) {
            const bool on = n0 + k < p.n;
            base[k] = plane0 + (on ? n0 + k : n0) * plane_stride + plo;
            const unsigned int mis = static_cast<unsigned int>(base[k] & 3u);      // tensor bases are 16-byte aligned
            head[k] = on ? min((4u - mis) & 3u, len) : 0u;
            nbody[k] = on ? (len - head[k]) >> 2 : 0u;
            most = max(most, nbody[k]);
        }
        for (unsigned int v = threadIdx.x; v < most; v += kThreads) {
            float4 a[PL] = {}, b[PL] = {}, y[PL] = {};
#pragma unroll
            for (int k = 0; k < PL; ++k)
                if (v < nbody[k]) {
                    const size_t idx = base[k] + head[k] + 4u * v;
                    a[k] = *reinterpret_cast<const float4*>(p.a + idx);
                    if (BWD) b[k] = *reinterpret_cast<const float4*>(p.b + idx);
                    if (BWD && RELU) y[k] = *reinterpret_cast<const float4*>(p.y + idx);
                }
#pragma unroll
            for (int k = 0; k < PL; ++k)
                if (v < nbody[k]) {
                    add(a[k].x, b[k].x, y[k].x); add(a[k].y, b[k].y, y[k].y);
                    add(a[k].z, b[k].z, y[k].z); add(a[k].w, b[k].w, y[k].w);
                }
        }
        // scalar heads (<= 3 elements) and tails (<= 3): lanes 0..7 of warp k take plane k's ends
        const unsigned int w = threadIdx.x >> 5, lane = threadIdx.x & 31u;
#pragma unroll
        for (int k = 0; k < PL; ++k)
            if (w == static_cast<unsigned int>(k) && n0 + k < p.n) {
                const unsigned int tail0 = head[k] + 4u * nbody[k];
                unsigned int e = len;                                        // none
                if (lane < head[k]) e = lane;
                else if (lane >= 4u && lane - 4u < len - tail0) e = tail0 + lane - 4u;
                if (e < len) {
                    const size_t idx = base[k] + e;
                    add(p.a[idx], BWD ? p.b[idx] : 0.f, (BWD && RELU) ? p.y[idx] : 0.f);
                }
            }
    }
    __shared__ double scratch[64];
    __shared__ int sflag;
    double d0 = acc0, d1 = acc1;
    block_sum2(d0, d1, scratch);
    if (threadIdx.x == 0) {
        if (BWD) d1 *= static_cast<double>(p.save_invstd[gc]);
        p.partials[static_cast<size_t>(gc) * p.splits + s] =
            BWD ? make_double2(d0, d1) : unshift_sums(d0, d1, static_cast<double>(len) * p.n, static_cast<double>(kshift));
    }
    if (!last_cta_arrives(p.counters + ch, p.groups * p.splits, &sflag)) return;
    reduce_finalize<BWD>(p, ch);
#undef add
}

template <int VEC, bool BWD, bool RELU>
__global__ void __launch_bounds__(kThreads) bn_reduce_kernel(const ReduceParams p) {
    using V = typename std::conditional<VEC == 4, float4, float>::type;
    const unsigned int s = blockIdx.x, gc = blockIdx.y;
    const unsigned int g = gc / p.c, ch = gc - g * p.c;
    const unsigned int J = p.n * p.hwv;                                   // vectors in this (g,c) domain
    const unsigned int lo = static_cast<unsigned int>(static_cast<unsigned long long>(J) * s / p.splits);
    const unsigned int hi = static_cast<unsigned int>(static_cast<unsigned long long>(J) * (s + 1) / p.splits);
    const V* a_v = reinterpret_cast<const V*>(p.a);
    const V* b_v = reinterpret_cast<const V*>(p.b);
    const V* y_v = reinterpret_cast<const V*>(p.y);
    const size_t plane0 = (static_cast<size_t>(g) * p.n * p.c + ch) * p.hwv;   // vector offset of plane (g*N+0, ch)
    const size_t plane_stride = static_cast<size_t>(p.c) * p.hwv;
    const float mean = BWD ? p.save_mean[gc] : 0.f;
    const float kshift = BWD ? 0.f : __ldg(p.a + plane0 * VEC);          // forward: accumulate x - K (see unshift_sums)

    float acc0 = 0.f, acc1 = 0.f;
    for (unsigned int j0 = lo + threadIdx.x; j0 < hi; j0 += kThreads * kBnUnroll) {
        V a[kBnUnroll] = {}, b[kBnUnroll] = {}, y[kBnUnroll] = {};
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const unsigned int j = j0 + u * kThreads;
            if (j < hi) {
                const unsigned int nn = j / p.hwv, off = j - nn * p.hwv;
                const size_t idx = plane0 + nn * plane_stride + off;
                a[u] = a_v[idx];                                         // everything is re-read by pass 2 -> keep in L2
                if (BWD) b[u] = b_v[idx];
                if (BWD && RELU) y[u] = y_v[idx];
            }
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const unsigned int j = j0 + u * kThreads;
            if (j < hi) {
                if constexpr (VEC == 4) {
                    const float av[4] = {a[u].x, a[u].y, a[u].z, a[u].w};
                    const float bv[4] = {b[u].x, b[u].y, b[u].z, b[u].w};
                    const float yv[4] = {y[u].x, y[u].y, y[u].z, y[u].w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (BWD) {
                            const float d = (RELU && !(yv[e] > 0.f)) ? 0.f : av[e];
                            acc0 += d;
                            acc1 = fmaf(d, bv[e] - mean, acc1);
                        } else {
                            const float t = av[e] - kshift;
                            acc0 += t;
                            acc1 = fmaf(t, t, acc1);
                        }
                    }
                } else {
                    if (BWD) {
                        const float d = (RELU && !(y[u] > 0.f)) ? 0.f : a[u];
                        acc0 += d;
                        acc1 = fmaf(d, b[u] - mean, acc1);
                    } else {
                        const float t = a[u] - kshift;
                        acc0 += t;
                        acc1 = fmaf(t, t, acc1);
                    }
                }
            }
        }
    }

    __shared__ double scratch[64];
    __shared__ int sflag;
    double d0 = acc0, d1 = acc1;
    block_sum2(d0, d1, scratch);
    if (threadIdx.x == 0) {
        if (BWD) d1 *= static_cast<double>(p.save_invstd[gc]);            // sum dy*(x-mean) -> sum dy*xhat
        p.partials[static_cast<size_t>(gc) * p.splits + s] =
            BWD ? make_double2(d0, d1) : unshift_sums(d0, d1, static_cast<double>(hi - lo) * VEC, static_cast<double>(kshift));
    }
    // last CTA of this CHANNEL (all groups, all splits) folds and finalises
    if (!last_cta_arrives(p.counters + ch, p.groups * p.splits, &sflag)) return;
    reduce_finalize<BWD>(p, ch);
}

// stand-alone finalisers for the NCCL path (sums already all-reduced): one thread per channel
__global__ void bn_fwd_finalize_kernel(const ReduceParams p, const double* sums) {
    const unsigned int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.c) return;
    double t0[16], t1[16];
    for (unsigned int g = 0; g < p.groups; ++g) {
        t0[g] = sums[(static_cast<size_t>(g) * p.c + ch) * 2];
        t1[g] = sums[(static_cast<size_t>(g) * p.c + ch) * 2 + 1];
    }
    fwd_finalize_channel(p, ch, t0, t1);
}
__global__ void bn_bwd_finalize_kernel(const ReduceParams p, const double* sums) {
    const unsigned int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= p.c) return;
    for (unsigned int g = 0; g < p.groups; ++g) {
        const size_t gc = static_cast<size_t>(g) * p.c + ch;
        bwd_finalize_group(p, static_cast<unsigned int>(gc), ch, sums[gc * 2], sums[gc * 2 + 1]);
    }
}

// ---- pass 2 -------------------------------------------------------------------------------------
struct ApplyParams {
    const float* x;          // fwd: x     bwd: dy
    const float* b;          // fwd: residual (nullable)   bwd: x
    const float* y;          // bwd + relu: forward output
    float* out;              // fwd: y     bwd: dx
    float* out2;             // bwd: dresidual (nullable)
    const void* table;       // fwd: float2 [G*C]   bwd: float4 [G*C]
    unsigned int total_v;    // vectors in the tensor
    unsigned int hwv, c, n;  // n = samples per group
    unsigned int hw_flat;    // != 0: H*W % 4 != 0 but the tensor is moved as 16-byte vectors of the FLAT index space; a
                             // vector may then straddle two planes (one in H*W/4 does): its first `cut` lanes belong to
                             // plane e0 / hw, the rest to the next plane
};

// table slot(s) of the vector at flat vector index i: (index of lanes [0, cut), index of lanes [cut, 4), cut)
template <bool FLAT>
__device__ __forceinline__ void apply_slots(const ApplyParams& p, unsigned int i, bool with_groups, unsigned int& s0,
                                            unsigned int& s1, unsigned int& cut) {
    auto slot = [&](unsigned int plane) {
        const unsigned int smp = plane / p.c, ch = plane - smp * p.c;
        return with_groups ? (smp / p.n) * p.c + ch : ch;
    };
    if constexpr (!FLAT) {
        s0 = s1 = slot(i / p.hwv);
        cut = 4u;
    } else {
        const unsigned int e0 = i * 4u, plane = e0 / p.hw_flat, left = p.hw_flat - (e0 - plane * p.hw_flat);
        s0 = slot(plane);
        cut = left < 4u ? left : 4u;
        s1 = left < 4u ? slot(plane + 1u) : s0;
    }
}

template <int VEC, bool RELU, bool RES, bool FLAT = false>
__global__ void __launch_bounds__(kThreads) bn_fwd_apply_kernel(const ApplyParams p) {
    using V = typename std::conditional<VEC == 4, float4, float>::type;
    const V* x_v = reinterpret_cast<const V*>(p.x);
    const V* r_v = reinterpret_cast<const V*>(p.b);
    V* y_v = reinterpret_cast<V*>(p.out);
    const float2* table = static_cast<const float2*>(p.table);
    const unsigned int span = gridDim.x * kThreads * kBnUnroll;
    for (unsigned int i0 = blockIdx.x * kThreads * kBnUnroll + threadIdx.x; i0 < p.total_v; i0 += span) {
        V x[kBnUnroll] = {}, r[kBnUnroll] = {};
        float2 ss[kBnUnroll] = {}, st[kBnUnroll] = {};
        unsigned int cut[kBnUnroll] = {};
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const unsigned int i = i0 + u * kThreads;
            if (i < p.total_v) {
                x[u] = ld_stream(x_v + i);                 // last use of x in the forward pass
                if (RES) r[u] = ld_stream(r_v + i);
                unsigned int s0, s1;
                apply_slots<FLAT>(p, i, true, s0, s1, cut[u]);
                ss[u] = __ldg(table + s0);
                st[u] = (FLAT && s1 != s0) ? __ldg(table + s1) : ss[u];
            }
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const unsigned int i = i0 + u * kThreads;
            if (i < p.total_v) {
                V o;
                if constexpr (VEC == 4) {
                    const float2 t1 = (!FLAT || cut[u] > 1u) ? ss[u] : st[u], t2 = (!FLAT || cut[u] > 2u) ? ss[u] : st[u], t3 = (!FLAT || cut[u] > 3u) ? ss[u] : st[u];
                    o.x = fmaf(x[u].x, ss[u].x, ss[u].y); o.y = fmaf(x[u].y, t1.x, t1.y);
                    o.z = fmaf(x[u].z, t2.x, t2.y); o.w = fmaf(x[u].w, t3.x, t3.y);
                    if (RES) { o.x += r[u].x; o.y += r[u].y; o.z += r[u].z; o.w += r[u].w; }
                    if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                } else {
                    o = fmaf(x[u], ss[u].x, ss[u].y);
                    if (RES) o += r[u];
                    if (RELU) o = fmaxf(o, 0.f);
                }
                y_v[i] = o;                                // consumed by the next conv: default policy (stay in L2)
            }
        }
    }
}

template <int VEC, bool RELU, bool DRES, bool FLAT = false>
__global__ void __launch_bounds__(kThreads) bn_bwd_apply_kernel(const ApplyParams p) {
    using V = typename std::conditional<VEC == 4, float4, float>::type;
    const V* dy_v = reinterpret_cast<const V*>(p.x);
    const V* x_v = reinterpret_cast<const V*>(p.b);
    const V* y_v = reinterpret_cast<const V*>(p.y);
    V* dx_v = reinterpret_cast<V*>(p.out);
    V* dr_v = reinterpret_cast<V*>(p.out2);
    const float4* coef = static_cast<const float4*>(p.table);
    const unsigned int span = gridDim.x * kThreads * kBnUnroll;
    for (unsigned int i0 = blockIdx.x * kThreads * kBnUnroll + threadIdx.x; i0 < p.total_v; i0 += span) {
        V dy[kBnUnroll] = {}, x[kBnUnroll] = {}, y[kBnUnroll] = {};
        float4 cf[kBnUnroll] = {}, cg[kBnUnroll] = {};
        unsigned int cut[kBnUnroll] = {};
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const unsigned int i = i0 + u * kThreads;
            if (i < p.total_v) {
                dy[u] = ld_stream(dy_v + i);
                x[u] = ld_stream(x_v + i);
                if (RELU) y[u] = ld_stream(y_v + i);
                unsigned int s0, s1;
                apply_slots<FLAT>(p, i, true, s0, s1, cut[u]);
                cf[u] = __ldg(coef + s0);
                cg[u] = (FLAT && s1 != s0) ? __ldg(coef + s1) : cf[u];
            }
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const unsigned int i = i0 + u * kThreads;
            if (i < p.total_v) {
                V dx, dr;
                // dx = w*invstd * (dy_eff - mean(dy) - (x - mean) * invstd * mean(dy*xhat))
                auto one = [&](const float4 c4, float d, float xv, float yv, float& o, float& r) {
                    const float de = (RELU && !(yv > 0.f)) ? 0.f : d;
                    r = de;
                    o = c4.x * (de - c4.y - (xv - c4.w) * c4.z);
                };
                if constexpr (VEC == 4) {
                    one(cf[u], dy[u].x, x[u].x, y[u].x, dx.x, dr.x);
                    one((!FLAT || cut[u] > 1u) ? cf[u] : cg[u], dy[u].y, x[u].y, y[u].y, dx.y, dr.y);
                    one((!FLAT || cut[u] > 2u) ? cf[u] : cg[u], dy[u].z, x[u].z, y[u].z, dx.z, dr.z);
                    one((!FLAT || cut[u] > 3u) ? cf[u] : cg[u], dy[u].w, x[u].w, y[u].w, dx.w, dr.w);
                } else {
                    one(cf[u], dy[u], x[u], y[u], dx, dr);
                }
                dx_v[i] = dx;
                if (DRES) dr_v[i] = dr;
            }
        }
    }
}

// Backward of the frozen-statistics affine (afan_bn_affine_f32): dx = scale[c] * dy_eff, dresidual = dy_eff, with
// dy_eff = dy where the forward output was positive (ReLU) -- one pass: reads dy (+ y), writes dx (+ dresidual).
template <int VEC, bool RELU, bool DRES, bool FLAT = false>
__global__ void __launch_bounds__(kThreads) bn_affine_bwd_kernel(const ApplyParams p) {
    using V = typename std::conditional<VEC == 4, float4, float>::type;
    const V* dy_v = reinterpret_cast<const V*>(p.x);
    const V* y_v = reinterpret_cast<const V*>(p.y);
    V* dx_v = reinterpret_cast<V*>(p.out);
    V* dr_v = reinterpret_cast<V*>(p.out2);
    const float2* table = static_cast<const float2*>(p.table);
    const unsigned int span = gridDim.x * kThreads * kBnUnroll;
    for (unsigned int i0 = blockIdx.x * kThreads * kBnUnroll + threadIdx.x; i0 < p.total_v; i0 += span) {
        V dy[kBnUnroll] = {}, y[kBnUnroll] = {};
        float sc[kBnUnroll] = {}, sd[kBnUnroll] = {};
        unsigned int cut[kBnUnroll] = {};
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const unsigned int i = i0 + u * kThreads;
            if (i < p.total_v) {
                dy[u] = ld_stream(dy_v + i);
                if (RELU) y[u] = __ldg(y_v + i);               // y stays live: it is the next layer's saved input
                unsigned int s0, s1;
                apply_slots<FLAT>(p, i, false, s0, s1, cut[u]);
                sc[u] = __ldg(table + s0).x;
                sd[u] = (FLAT && s1 != s0) ? __ldg(table + s1).x : sc[u];
            }
        }
#pragma unroll
        for (int u = 0; u < kBnUnroll; ++u) {
            const unsigned int i = i0 + u * kThreads;
            if (i < p.total_v) {
                V dx, dr;
                auto one = [&](float scale, float d, float yv, float& o, float& r) {
                    const float de = (RELU && !(yv > 0.f)) ? 0.f : d;
                    r = de;
                    o = scale * de;
                };
                if constexpr (VEC == 4) {
                    one(sc[u], dy[u].x, y[u].x, dx.x, dr.x);
                    one((!FLAT || cut[u] > 1u) ? sc[u] : sd[u], dy[u].y, y[u].y, dx.y, dr.y);
                    one((!FLAT || cut[u] > 2u) ? sc[u] : sd[u], dy[u].z, y[u].z, dx.z, dr.z);
                    one((!FLAT || cut[u] > 3u) ? sc[u] : sd[u], dy[u].w, y[u].w, dx.w, dr.w);
                } else {
                    one(sc[u], dy[u], y[u], dx, dr);
                }
                dx_v[i] = dx;
                if (DRES) dr_v[i] = dr;
            }
        }
    }
}


// =====================================================================================================
// Cluster path (the one every BASELINE config takes): ONE launch per BatchNorm direction.
//
// A thread-block cluster of CS CTAs owns one channel (all statistic groups).  Each CTA streams its slice
// of every group's domain once from HBM (128-bit loads), reduces with warp shuffles + a shared-memory
// tree, and publishes one double2 partial per group in its shared memory.  After a cluster barrier every
// CTA pulls the CS partials of each group over DISTRIBUTED SHARED MEMORY in rank order (deterministic),
// finalises mean / invstd (rank 0 also updates the running statistics in pass order), and sweeps its slice
// a second time -- now an L2 hit, the slice was touched microseconds ago -- to normalise + affine
// (+ residual) + ReLU.  HBM traffic is the algorithmic 8 B/elem (fwd) / 16 B/elem (bwd with ReLU mask);
// no global-memory partials, no atomics, no second launch.
// =====================================================================================================
constexpr int kClusterThreads = 512;
constexpr int kMaxGroups = 16;                    // learnable-eta A-FAN batches 9 adversarial groups + clean
constexpr int kMaxCluster = 8;                     // portable cluster size on sm_100a

struct ClusterParams {
    const float* a;          // fwd: x            bwd: dy
    const float* b;          // fwd: residual     bwd: x
    const float* y;          // bwd + relu: forward output
    float* out;              // fwd: y            bwd: dx
    float* out2;             // bwd: dresidual
    const float* weight;
    const float* bias;
    float* running_mean;
    float* running_var;
    float* save_mean;        // fwd: written      bwd: read
    float* save_invstd;
    float* dweight;
    float* dbias;
    double count;
    float eps, momentum;
    int replay;
    unsigned int groups, n, c, hwv;
    const float2* mask_table;   // bwd + relu, y == nullptr: ReLU mask recomputed as x * scale + shift > 0 from the forward's
                                // per-(group, channel) table (the fused convolution never materialised y)
};

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory"); }

// Gather the per-rank partials of every group over DSMEM into local shared memory, then fold them in
// rank order: tot[g] valid in threads g < groups after the call.  Contains the cluster barrier.
__device__ __forceinline__ double2 cluster_fold(cg::cluster_group& cluster, double2* s_part, double2 (*s_all)[kMaxCluster],
                                                unsigned int groups) {
    const unsigned int cs = cluster.num_blocks();
    cluster.sync();                                                 // every CTA's s_part is published
    if (threadIdx.x < groups * cs) {
        const unsigned int g = threadIdx.x / cs, r = threadIdx.x - g * cs;
        s_all[g][r] = *cluster.map_shared_rank(&s_part[g], r);      // parallel DSMEM reads: one latency, not G*CS
    }
    __syncthreads();
    cluster_arrive();                                               // remote reads done: peers may exit later
    double2 tot = make_double2(0.0, 0.0);
    if (threadIdx.x < groups)
        for (unsigned int r = 0; r < cs; ++r) { tot.x += s_all[threadIdx.x][r].x; tot.y += s_all[threadIdx.x][r].y; }
    return tot;
}

template <int VEC, bool RELU, bool RES>
__global__ void __launch_bounds__(kClusterThreads) bn_fwd_cluster_kernel(const ClusterParams p) {
    using V = typename std::conditional<VEC == 4, float4, float>::type;
    constexpr int U = VEC == 4 ? kBnUnroll : 4 * kBnUnroll;       // scalar path (H*W % 4 != 0): same bytes in flight per thread
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int cs = cluster.num_blocks(), rank = cluster.block_rank(), ch = blockIdx.x / cs;
    pdl_wait();                                                     // PDL: scheduled during the previous kernel's tail
    pdl_launch_dependents();
    __shared__ double2 s_part[kMaxGroups];
    __shared__ double2 s_all[kMaxGroups][kMaxCluster];
    __shared__ double2 s_stat[kMaxGroups];                          // (mean, unbiased var) for the running update
    __shared__ float2 s_ss[kMaxGroups];
    __shared__ double scratch[64];
    const unsigned int J = p.n * p.hwv;
    const unsigned int lo = static_cast<unsigned int>(static_cast<unsigned long long>(J) * rank / cs);
    const unsigned int hi = static_cast<unsigned int>(static_cast<unsigned long long>(J) * (rank + 1) / cs);
    const V* x_v = reinterpret_cast<const V*>(p.a);
    const V* r_v = reinterpret_cast<const V*>(p.b);
    V* y_v = reinterpret_cast<V*>(p.out);
    const size_t plane_stride = static_cast<size_t>(p.c) * p.hwv;

    // ---- sweep 1: per-group sum / sum of squares of this CTA's slice ----
    for (unsigned int g = 0; g < p.groups; ++g) {
        const size_t plane0 = (static_cast<size_t>(g) * p.n * p.c + ch) * p.hwv;
        const float kshift = __ldg(p.a + plane0 * VEC);               // accumulate x - K (see unshift_sums)
        float acc0 = 0.f, acc1 = 0.f;
        for (unsigned int j0 = lo + threadIdx.x; j0 < hi; j0 += kClusterThreads * U) {
            V a[U];
            bool ok[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned int j = j0 + u * kClusterThreads;
                ok[u] = j < hi;
                if constexpr (VEC == 4) a[u] = make_float4(kshift, kshift, kshift, kshift); else a[u] = kshift;
                if (ok[u]) {
                    const unsigned int nn = j / p.hwv, off = j - nn * p.hwv;
                    a[u] = x_v[plane0 + nn * plane_stride + off];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if constexpr (VEC == 4) {
                    const float tx = a[u].x - kshift, ty = a[u].y - kshift, tz = a[u].z - kshift, tw = a[u].w - kshift;
                    acc0 += (tx + ty) + (tz + tw);
                    acc1 = fmaf(tx, tx, fmaf(ty, ty, fmaf(tz, tz, fmaf(tw, tw, acc1))));
                } else {
                    const float t = a[u] - kshift;
                    acc0 += t;
                    acc1 = fmaf(t, t, acc1);
                }
            }
        }
        double d0 = acc0, d1 = acc1;
        block_sum2(d0, d1, scratch);
        if (threadIdx.x == 0) s_part[g] = unshift_sums(d0, d1, static_cast<double>(hi - lo) * VEC, static_cast<double>(kshift));
    }

    // ---- cluster-wide fold over DSMEM, finalise ----
    const double2 tot = cluster_fold(cluster, s_part, s_all, p.groups);
    if (threadIdx.x < p.groups) {
        const unsigned int g = threadIdx.x, gc = g * p.c + ch;
        const double mean = tot.x / p.count;
        double var = tot.y / p.count - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        const double invstd = rsqrt(var + static_cast<double>(p.eps));
        const float w = p.weight ? p.weight[ch] : 1.f, b = p.bias ? p.bias[ch] : 0.f;
        s_ss[g] = make_float2(static_cast<float>(w * invstd), static_cast<float>(b - mean * w * invstd));
        s_stat[g] = make_double2(mean, p.count > 1.0 ? var * (p.count / (p.count - 1.0)) : var);
        if (rank == 0) {
            p.save_mean[gc] = static_cast<float>(mean);
            p.save_invstd[gc] = static_cast<float>(invstd);
        }
    }
    __syncthreads();
    if (rank == 0 && threadIdx.x == 0 && p.running_mean && p.running_var) {
        float rm = p.running_mean[ch], rv = p.running_var[ch];
        for (unsigned int g = 0; g < p.groups; ++g)                  // pass order: adv group first, then clean
            for (int r = 0; r < p.replay; ++r) {
                rm = static_cast<float>((1.0 - p.momentum) * rm + p.momentum * s_stat[g].x);
                rv = static_cast<float>((1.0 - p.momentum) * rv + p.momentum * s_stat[g].y);
            }
        p.running_mean[ch] = rm;
        p.running_var[ch] = rv;
    }

    // ---- sweep 2 (L2-hot): normalise + affine (+ residual) + ReLU ----
    for (unsigned int g = 0; g < p.groups; ++g) {
        const size_t plane0 = (static_cast<size_t>(g) * p.n * p.c + ch) * p.hwv;
        const float2 ss = s_ss[g];
        for (unsigned int j0 = lo + threadIdx.x; j0 < hi; j0 += kClusterThreads * U) {
            V a[U] = {}, r[U] = {};
            size_t idx[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned int j = j0 + u * kClusterThreads;
                if (j < hi) {
                    const unsigned int nn = j / p.hwv, off = j - nn * p.hwv;
                    idx[u] = plane0 + nn * plane_stride + off;
                    a[u] = ld_stream(x_v + idx[u]);
                    if (RES) r[u] = ld_stream(r_v + idx[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned int j = j0 + u * kClusterThreads;
                if (j < hi) {
                    V o;
                    if constexpr (VEC == 4) {
                        o.x = fmaf(a[u].x, ss.x, ss.y); o.y = fmaf(a[u].y, ss.x, ss.y);
                        o.z = fmaf(a[u].z, ss.x, ss.y); o.w = fmaf(a[u].w, ss.x, ss.y);
                        if (RES) { o.x += r[u].x; o.y += r[u].y; o.z += r[u].z; o.w += r[u].w; }
                        if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                    } else {
                        o = fmaf(a[u], ss.x, ss.y);
                        if (RES) o += r[u];
                        if (RELU) o = fmaxf(o, 0.f);
                    }
                    y_v[idx[u]] = o;
                }
            }
        }
    }
    cluster_wait();                                                 // nobody exits while a peer may still read its smem
}

template <int VEC, bool RELU, bool DRES>
__global__ void __launch_bounds__(kClusterThreads) bn_bwd_cluster_kernel(const ClusterParams p) {
    using V = typename std::conditional<VEC == 4, float4, float>::type;
    constexpr int U = VEC == 4 ? kBnUnroll : 4 * kBnUnroll;       // scalar path (H*W % 4 != 0): same bytes in flight per thread
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int cs = cluster.num_blocks(), rank = cluster.block_rank(), ch = blockIdx.x / cs;
    pdl_wait();                                                     // PDL: scheduled during the previous kernel's tail
    pdl_launch_dependents();
    __shared__ double2 s_part[kMaxGroups];
    __shared__ double2 s_all[kMaxGroups][kMaxCluster];
    __shared__ double2 s_sum[kMaxGroups];
    __shared__ float4 s_cf[kMaxGroups];
    __shared__ double scratch[64];
    const unsigned int J = p.n * p.hwv;
    const unsigned int lo = static_cast<unsigned int>(static_cast<unsigned long long>(J) * rank / cs);
    const unsigned int hi = static_cast<unsigned int>(static_cast<unsigned long long>(J) * (rank + 1) / cs);
    const V* dy_v = reinterpret_cast<const V*>(p.a);
    const V* x_v = reinterpret_cast<const V*>(p.b);
    const V* y_v = reinterpret_cast<const V*>(p.y);
    V* dx_v = reinterpret_cast<V*>(p.out);
    V* dr_v = reinterpret_cast<V*>(p.out2);
    const size_t plane_stride = static_cast<size_t>(p.c) * p.hwv;

    // ---- sweep 1: sum dy_eff, sum dy_eff * (x - mean) ----
    for (unsigned int g = 0; g < p.groups; ++g) {
        const size_t plane0 = (static_cast<size_t>(g) * p.n * p.c + ch) * p.hwv;
        const float mean = p.save_mean[g * p.c + ch];
        float acc0 = 0.f, acc1 = 0.f;
        for (unsigned int j0 = lo + threadIdx.x; j0 < hi; j0 += kClusterThreads * U) {
            V d[U] = {}, x[U] = {}, y[U] = {};
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned int j = j0 + u * kClusterThreads;
                if (j < hi) {
                    const unsigned int nn = j / p.hwv, off = j - nn * p.hwv;
                    const size_t idx = plane0 + nn * plane_stride + off;
                    d[u] = dy_v[idx];
                    x[u] = x_v[idx];
                    if (RELU) y[u] = y_v[idx];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned int j = j0 + u * kClusterThreads;
                if (j < hi) {
                    auto one = [&](float dv, float xv, float yv) {
                        const float de = (RELU && !(yv > 0.f)) ? 0.f : dv;
                        acc0 += de;
                        acc1 = fmaf(de, xv - mean, acc1);
                    };
                    if constexpr (VEC == 4) {
                        one(d[u].x, x[u].x, y[u].x); one(d[u].y, x[u].y, y[u].y);
                        one(d[u].z, x[u].z, y[u].z); one(d[u].w, x[u].w, y[u].w);
                    } else {
                        one(d[u], x[u], y[u]);
                    }
                }
            }
        }
        double d0 = acc0, d1 = acc1;
        block_sum2(d0, d1, scratch);
        if (threadIdx.x == 0) s_part[g] = make_double2(d0, d1);
    }

    const double2 tot = cluster_fold(cluster, s_part, s_all, p.groups);
    if (threadIdx.x < p.groups) {
        const unsigned int g = threadIdx.x, gc = g * p.c + ch;
        const float invstd = p.save_invstd[gc], w = p.weight ? p.weight[ch] : 1.f;
        const double s_dy = tot.x, s_dyxh = tot.y * static_cast<double>(invstd);
        s_sum[g] = make_double2(s_dy, s_dyxh);
        s_cf[g] = make_float4(w * invstd, static_cast<float>(s_dy / p.count),
                              static_cast<float>(s_dyxh / p.count * invstd), p.save_mean[gc]);
    }
    __syncthreads();
    if (rank == 0 && threadIdx.x == 0) {
        double dw = 0.0, db = 0.0;
        for (unsigned int g = 0; g < p.groups; ++g) { db += s_sum[g].x; dw += s_sum[g].y; }
        if (p.dweight) p.dweight[ch] = static_cast<float>(dw);
        if (p.dbias) p.dbias[ch] = static_cast<float>(db);
    }

    // ---- sweep 2 (L2-hot): dx = w*invstd * (dy_eff - mean(dy_eff) - xhat * mean(dy_eff*xhat)) ----
    for (unsigned int g = 0; g < p.groups; ++g) {
        const size_t plane0 = (static_cast<size_t>(g) * p.n * p.c + ch) * p.hwv;
        const float4 cf = s_cf[g];
        for (unsigned int j0 = lo + threadIdx.x; j0 < hi; j0 += kClusterThreads * U) {
            V d[U] = {}, x[U] = {}, y[U] = {};
            size_t idx[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned int j = j0 + u * kClusterThreads;
                if (j < hi) {
                    const unsigned int nn = j / p.hwv, off = j - nn * p.hwv;
                    idx[u] = plane0 + nn * plane_stride + off;
                    d[u] = ld_stream(dy_v + idx[u]);
                    x[u] = ld_stream(x_v + idx[u]);
                    if (RELU) y[u] = ld_stream(y_v + idx[u]);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const unsigned int j = j0 + u * kClusterThreads;
                if (j < hi) {
                    V dx, dr;
                    auto one = [&](float dv, float xv, float yv, float& o, float& r) {
                        const float de = (RELU && !(yv > 0.f)) ? 0.f : dv;
                        r = de;
                        o = cf.x * (de - cf.y - (xv - cf.w) * cf.z);
                    };
                    if constexpr (VEC == 4) {
                        one(d[u].x, x[u].x, y[u].x, dx.x, dr.x); one(d[u].y, x[u].y, y[u].y, dx.y, dr.y);
                        one(d[u].z, x[u].z, y[u].z, dx.z, dr.z); one(d[u].w, x[u].w, y[u].w, dx.w, dr.w);
                    } else {
                        one(d[u], x[u], y[u], dx, dr);
                    }
                    dx_v[idx[u]] = dx;
                    if (DRES) dr_v[idx[u]] = dr;
                }
            }
        }
    }
    cluster_wait();
}



// -----------------------------------------------------------------------------------------------------
// Fused statistics exchange over NVLink peer memory (multi-GPU, one process per GPU).
//
// Instead of  [stats kernel] -> NCCL all-reduce -> [finalize kernel] -> [apply kernel]  the register-
// resident cluster kernel itself exchanges the per-(group, channel) sums with the other GPUs: the rank-0
// CTA of each cluster stores its LOCAL sums into every peer's mailbox (P2P stores over NVLink/NVSwitch)
// as self-validating words with in-band tags, spins (bounded) until the `world` contributions of its
// channel carry this call's tag, and folds them in rank order -- every GPU adds the same numbers in the same order, so the global
// statistics are bit-identical on all ranks.  One launch per BatchNorm direction, no NCCL call, clean and
// adversarial statistics travel together.
//
// Mailbox (one per GPU, peer-mapped by cudaIpc): RING slots x [world][cmax] x [2 groups][2 sums] 16-byte words.
// `state` (local device memory): {seq, ticket, error}.  seq counts BN calls; slot = seq % RING; the last CTA
// of every launch increments seq.  RING >= 2 suffices: a GPU can run ahead of a peer by at most one exchange
// (it cannot pass exchange k+1 before the peer has published k+1, i.e. finished reading k).
// -----------------------------------------------------------------------------------------------------
// Cluster fold + cross-GPU fold.  Returns the GLOBAL totals in threads g < groups of EVERY CTA; the rank-0
// CTA additionally keeps the LOCAL totals in s_loc (the backward needs them for dweight / dbias).
__device__ __forceinline__ double2 cluster_fold_p2p(cg::cluster_group& cluster, double2* s_part,
                                                    double2 (*s_all)[kMaxCluster], double2* s_loc, double2* s_glob,
                                                    unsigned int groups, unsigned int ch, const P2PParams& q,
                                                    unsigned long long seq) {
    // Every CTA of the cluster gathers the cluster's partials over DSMEM (as in cluster_fold) and then collects the ranks'
    // sums from this GPU's mailbox ITSELF; only cluster rank 0 publishes.  Round 1 had rank 0 do the exchange and hand the
    // result to its cluster mates through a second cluster barrier + a DSMEM read: that hop sat in the critical chain of every
    // exchange (390 per step).  The mailbox words are local memory; polling them from cs CTAs costs nothing on the wire.
    __shared__ double s_recv[kP2PMaxWorld][2][2];                   // [src rank][group][k]
    const unsigned int cs = cluster.num_blocks(), rank = cluster.block_rank(), t = threadIdx.x;
    cluster.sync();                                                 // every CTA's s_part is published
    if (t < groups * cs) {
        const unsigned int g = t / cs, r = t - g * cs;
        s_all[g][r] = *cluster.map_shared_rank(&s_part[g], r);
    }
    __syncthreads();
    cluster_arrive();                                               // remote reads done: peers may exit later
    if (t < groups) {
        double2 loc = make_double2(0.0, 0.0);
        for (unsigned int r = 0; r < cs; ++r) { loc.x += s_all[t][r].x; loc.y += s_all[t][r].y; }
        s_loc[t] = loc;
    }
    __syncthreads();
    const unsigned int slot = static_cast<unsigned int>(seq % kP2PRing);
    const unsigned int tag = p2p_tag(seq);
    const unsigned int words = static_cast<unsigned int>(q.world) * groups * 2;
    if (t < words) {
        const unsigned int peer = t / (groups * 2), rem = t - peer * groups * 2, g = rem >> 1, k = rem & 1;
        if (rank == 0) p2p_publish(q, static_cast<int>(peer), slot, ch, g, k, tag, k ? s_loc[g].y : s_loc[g].x);
        s_recv[peer][g][k] = p2p_collect(q, static_cast<int>(peer), slot, ch, g, k, tag);
    }
    __syncthreads();
    double2 tot = make_double2(0.0, 0.0);
    if (t < groups)
        for (int r = 0; r < q.world; ++r) { tot.x += s_recv[r][t][0]; tot.y += s_recv[r][t][1]; }   // rank order everywhere
    (void)s_glob;
    return tot;
}

// the last CTA of the launch advances the call sequence (stream order makes it visible to the next BN launch)
__device__ __forceinline__ void p2p_advance_seq(const P2PParams& q) {
    __shared__ int s_last;
    if (last_cta_arrives(reinterpret_cast<unsigned int*>(q.state + 1), gridDim.x, &s_last) && threadIdx.x == 0)
        q.state[0] = q.state[0] + 1ULL;
}

// -----------------------------------------------------------------------------------------------------
// Register-resident cluster kernels: when a CTA's slice fits in registers (<= 8 float4 per thread per
// group, G <= 2) every element is read from HBM exactly once, ALL loads are issued up front (one DRAM
// round trip), the statistics are folded over DSMEM, and the second sweep runs out of registers.
// These are the kernels the ResNet-56 tail (BASELINE config 2) runs.
// -----------------------------------------------------------------------------------------------------
template <int K>
__device__ __forceinline__ void block_reduce_k(const float (&v)[K], double* s_out, double (*s_warp)[K]) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float w[K];
#pragma unroll
    for (int k = 0; k < K; ++k) w[k] = warp_sum(v[k]);
    if (lane == 0) {
#pragma unroll
        for (int k = 0; k < K; ++k) s_warp[warp][k] = static_cast<double>(w[k]);
    }
    __syncthreads();
    if (threadIdx.x < K) {
        double t = 0.0;
        for (int ww = 0; ww < kClusterThreads / 32; ++ww) t += s_warp[ww][threadIdx.x];   // fixed order
        s_out[threadIdx.x] = t;
    }
}

// minBlocks = 2 (<= 64 registers) whenever the data fits: clusters of 8 only pack 16-per-chip when two CTAs
// can share an SM (ncu: launch__cluster_max_active = 15 at one CTA per SM -> a second wave).
template <int G, int NV, bool RELU, bool RES, bool P2P = false>
__global__ void __launch_bounds__(kClusterThreads, (G * NV <= 8) ? 2 : 1)
bn_fwd_cluster_reg_kernel(const ClusterParams p, const P2PParams q) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int cs = cluster.num_blocks(), rank = cluster.block_rank(), ch = blockIdx.x / cs;
    pdl_wait();                                                     // PDL: scheduled during the previous kernel's tail
    pdl_launch_dependents();
    __shared__ double2 s_part[kMaxGroups];
    __shared__ double2 s_all[kMaxGroups][kMaxCluster];
    __shared__ double2 s_stat[kMaxGroups];
    __shared__ float2 s_ss[kMaxGroups];
    __shared__ double s_warp[kClusterThreads / 32][2 * G];
    constexpr bool RES_EARLY = RES && (G * NV <= 4);   // residual prefetched with x only while it fits in 64 registers
    const unsigned int J = p.n * p.hwv;
    const unsigned int lo = static_cast<unsigned int>(static_cast<unsigned long long>(J) * rank / cs);
    const unsigned int hi = static_cast<unsigned int>(static_cast<unsigned long long>(J) * (rank + 1) / cs);
    const float4* x_v = reinterpret_cast<const float4*>(p.a);
    const float4* r_v = reinterpret_cast<const float4*>(p.b);
    float4* y_v = reinterpret_cast<float4*>(p.out);
    const unsigned int plane_stride = p.c * p.hwv;
    const unsigned int group_stride = p.n * plane_stride;            // total vectors < 2^32 (checked on the host)
    // call sequence number of the exchange: stable for the whole launch, fetched early so its latency hides under the loads
    const unsigned long long seq = P2P ? *reinterpret_cast<const volatile unsigned long long*>(q.state) : 0ULL;

    unsigned int rel[NV];
    float4 a[G][NV], r[RES_EARLY ? G : 1][RES_EARLY ? NV : 1];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const unsigned int j = lo + threadIdx.x + i * kClusterThreads;
        if (j < hi) {
            const unsigned int nn = j / p.hwv;
            rel[i] = ch * p.hwv + nn * plane_stride + (j - nn * p.hwv);
        } else {
            rel[i] = 0xffffffffu;
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            a[g][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rel[i] != 0xffffffffu) {
                a[g][i] = ld_stream(x_v + (g * group_stride + rel[i]));       // read exactly once
                if (RES_EARLY) r[g][i] = ld_stream(r_v + (g * group_stride + rel[i]));
            }
        }
    float acc[2 * G], kshift[G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        // accumulate x - K, K = first element of the (group, channel) domain (see unshift_sums); padding lanes hold 0 and
        // are excluded by the mask
        kshift[g] = __ldg(reinterpret_cast<const float*>(x_v + (g * group_stride + ch * p.hwv)));
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (rel[i] != 0xffffffffu) {
                const float tx = a[g][i].x - kshift[g], ty = a[g][i].y - kshift[g], tz = a[g][i].z - kshift[g], tw = a[g][i].w - kshift[g];
                s0 += (tx + ty) + (tz + tw);
                s1 = fmaf(tx, tx, fmaf(ty, ty, fmaf(tz, tz, fmaf(tw, tw, s1))));
            }
        }
        acc[2 * g] = s0;
        acc[2 * g + 1] = s1;
    }
    block_reduce_k<2 * G>(acc, reinterpret_cast<double*>(s_part), s_warp);
    __syncthreads();
    if (threadIdx.x < G) {                                            // raw sums from here on
        const double cnt = 4.0 * static_cast<double>(hi - lo);
        s_part[threadIdx.x] = unshift_sums(s_part[threadIdx.x].x, s_part[threadIdx.x].y, cnt, static_cast<double>(kshift[threadIdx.x]));
    }

    __shared__ double2 s_loc[kMaxGroups], s_glob[kMaxGroups];
    const double2 tot = P2P ? cluster_fold_p2p(cluster, s_part, s_all, s_loc, s_glob, G, ch, q, seq)
                            : cluster_fold(cluster, s_part, s_all, G);
    if (threadIdx.x < G) {
        const unsigned int g = threadIdx.x, gc = g * p.c + ch;
        const double mean = tot.x / p.count;
        double var = tot.y / p.count - mean * mean;
        var = var < 0.0 ? 0.0 : var;
        const double invstd = rsqrt(var + static_cast<double>(p.eps));
        const float w = p.weight ? p.weight[ch] : 1.f, b = p.bias ? p.bias[ch] : 0.f;
        s_ss[g] = make_float2(static_cast<float>(w * invstd), static_cast<float>(b - mean * w * invstd));
        s_stat[g] = make_double2(mean, p.count > 1.0 ? var * (p.count / (p.count - 1.0)) : var);
        if (rank == 0) {
            p.save_mean[gc] = static_cast<float>(mean);
            p.save_invstd[gc] = static_cast<float>(invstd);
        }
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const float2 ss = s_ss[g];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (rel[i] != 0xffffffffu) {
                float4 o;
                o.x = fmaf(a[g][i].x, ss.x, ss.y); o.y = fmaf(a[g][i].y, ss.x, ss.y);
                o.z = fmaf(a[g][i].z, ss.x, ss.y); o.w = fmaf(a[g][i].w, ss.x, ss.y);
                if (RES) {
                    const float4 rr = RES_EARLY ? r[RES_EARLY ? g : 0][RES_EARLY ? i : 0] : ld_stream(r_v + (g * group_stride + rel[i]));
                    o.x += rr.x; o.y += rr.y; o.z += rr.z; o.w += rr.w;
                }
                if (RELU) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
                y_v[g * group_stride + rel[i]] = o;
            }
        }
    }
    if (rank == 0 && threadIdx.x == 0 && p.running_mean && p.running_var) {
        float rm = p.running_mean[ch], rv = p.running_var[ch];
        for (int g = 0; g < G; ++g)
            for (int rr = 0; rr < p.replay; ++rr) {
                rm = static_cast<float>((1.0 - p.momentum) * rm + p.momentum * s_stat[g].x);
                rv = static_cast<float>((1.0 - p.momentum) * rv + p.momentum * s_stat[g].y);
            }
        p.running_mean[ch] = rm;
        p.running_var[ch] = rv;
    }
    cluster_wait();
    if (P2P) p2p_advance_seq(q);
}

template <int G, int NV, bool RELU, bool DRES, bool P2P = false>
__global__ void __launch_bounds__(kClusterThreads, (G * NV <= 4) ? 2 : 1)
bn_bwd_cluster_reg_kernel(const ClusterParams p, const P2PParams q) {
    cg::cluster_group cluster = cg::this_cluster();
    const unsigned int cs = cluster.num_blocks(), rank = cluster.block_rank(), ch = blockIdx.x / cs;
    pdl_wait();                                                     // PDL: scheduled during the previous kernel's tail
    pdl_launch_dependents();
    __shared__ double2 s_part[kMaxGroups];
    __shared__ double2 s_all[kMaxGroups][kMaxCluster];
    __shared__ double2 s_sum[kMaxGroups];
    __shared__ float4 s_cf[kMaxGroups];
    __shared__ double s_warp[kClusterThreads / 32][2 * G];
    const unsigned int J = p.n * p.hwv;
    const unsigned int lo = static_cast<unsigned int>(static_cast<unsigned long long>(J) * rank / cs);
    const unsigned int hi = static_cast<unsigned int>(static_cast<unsigned long long>(J) * (rank + 1) / cs);
    const float4* dy_v = reinterpret_cast<const float4*>(p.a);
    const float4* x_v = reinterpret_cast<const float4*>(p.b);
    const float4* y_v = reinterpret_cast<const float4*>(p.y);
    float4* dx_v = reinterpret_cast<float4*>(p.out);
    float4* dr_v = reinterpret_cast<float4*>(p.out2);
    const unsigned int plane_stride = p.c * p.hwv;
    const unsigned int group_stride = p.n * plane_stride;
    const unsigned long long seq = P2P ? *reinterpret_cast<const volatile unsigned long long*>(q.state) : 0ULL;

    unsigned int rel[NV];
    float4 d[G][NV], x[G][NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
        const unsigned int j = lo + threadIdx.x + i * kClusterThreads;
        if (j < hi) {
            const unsigned int nn = j / p.hwv;
            rel[i] = ch * p.hwv + nn * plane_stride + (j - nn * p.hwv);
        } else {
            rel[i] = 0xffffffffu;
        }
    }
#pragma unroll
    for (int g = 0; g < G; ++g)
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            d[g][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            x[g][i] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (rel[i] != 0xffffffffu) {
                d[g][i] = ld_stream(dy_v + (g * group_stride + rel[i]));
                x[g][i] = ld_stream(x_v + (g * group_stride + rel[i]));
                if (RELU) {                                         // fold the ReLU mask into dy right away
                    float4 yy;
                    if (p.mask_table) {                             // y = relu(x * scale + shift) was never stored: same fma, same mask
                        const float2 t = __ldg(p.mask_table + g * p.c + ch);
                        yy = make_float4(fmaf(x[g][i].x, t.x, t.y), fmaf(x[g][i].y, t.x, t.y), fmaf(x[g][i].z, t.x, t.y),
                                         fmaf(x[g][i].w, t.x, t.y));
                    } else {
                        yy = ld_stream(y_v + (g * group_stride + rel[i]));
                    }
                    if (!(yy.x > 0.f)) d[g][i].x = 0.f;
                    if (!(yy.y > 0.f)) d[g][i].y = 0.f;
                    if (!(yy.z > 0.f)) d[g][i].z = 0.f;
                    if (!(yy.w > 0.f)) d[g][i].w = 0.f;
                }
            }
        }
    float acc[2 * G];
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const float mean = p.save_mean[g * p.c + ch];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            s0 += (d[g][i].x + d[g][i].y) + (d[g][i].z + d[g][i].w);
            s1 = fmaf(d[g][i].x, x[g][i].x - mean, fmaf(d[g][i].y, x[g][i].y - mean,
                 fmaf(d[g][i].z, x[g][i].z - mean, fmaf(d[g][i].w, x[g][i].w - mean, s1))));
        }
        acc[2 * g] = s0;
        acc[2 * g + 1] = s1;
    }
    block_reduce_k<2 * G>(acc, reinterpret_cast<double*>(s_part), s_warp);

    __shared__ double2 s_loc[kMaxGroups], s_glob[kMaxGroups];
    const double2 tot = P2P ? cluster_fold_p2p(cluster, s_part, s_all, s_loc, s_glob, G, ch, q, seq)
                            : cluster_fold(cluster, s_part, s_all, G);
    if (threadIdx.x < G) {
        const unsigned int g = threadIdx.x, gc = g * p.c + ch;
        const float invstd = p.save_invstd[gc], w = p.weight ? p.weight[ch] : 1.f;
        const double s_dy = tot.x, s_dyxh = tot.y * static_cast<double>(invstd);
        // dweight / dbias come from the LOCAL sums (the gradient all-reduce averages them with the other parameters)
        s_sum[g] = (P2P && rank == 0) ? make_double2(s_loc[g].x, s_loc[g].y * static_cast<double>(invstd))
                                      : make_double2(s_dy, s_dyxh);
        s_cf[g] = make_float4(w * invstd, static_cast<float>(s_dy / p.count),
                              static_cast<float>(s_dyxh / p.count * invstd), p.save_mean[gc]);
    }
    __syncthreads();
#pragma unroll
    for (int g = 0; g < G; ++g) {
        const float4 cf = s_cf[g];
#pragma unroll
        for (int i = 0; i < NV; ++i) {
            if (rel[i] != 0xffffffffu) {
                float4 o;
                o.x = cf.x * (d[g][i].x - cf.y - (x[g][i].x - cf.w) * cf.z);
                o.y = cf.x * (d[g][i].y - cf.y - (x[g][i].y - cf.w) * cf.z);
                o.z = cf.x * (d[g][i].z - cf.y - (x[g][i].z - cf.w) * cf.z);
                o.w = cf.x * (d[g][i].w - cf.y - (x[g][i].w - cf.w) * cf.z);
                dx_v[g * group_stride + rel[i]] = o;
                if (DRES) dr_v[g * group_stride + rel[i]] = d[g][i];
            }
        }
    }
    if (rank == 0 && threadIdx.x == 0) {
        double dw = 0.0, db = 0.0;
        for (int g = 0; g < G; ++g) { db += s_sum[g].x; dw += s_sum[g].y; }
        if (p.dweight) p.dweight[ch] = static_cast<float>(dw);
        if (p.dbias) p.dbias[ch] = static_cast<float>(db);
    }
    cluster_wait();
    if (P2P) p2p_advance_seq(q);
}

// (cluster size, vectors per thread) for the register-resident kernels; nv == 0 -> not applicable
struct RegPlan { int cs, nv; };
__host__ inline RegPlan pick_reg_plan(int64_t groups, int64_t n, int64_t c, int64_t hw, bool vec, int max_gnv) {
    RegPlan r{0, 0};
    if (!vec || groups > 2) return r;
    const int64_t J = n * (hw / 4);
    int cs = 1;
    while (cs < kMaxCluster && c * cs * 2 <= sm_count()) cs *= 2;    // whole chip in ONE wave of 1 CTA / SM
    for (int nv = 1; nv <= 8; nv *= 2) {
        if (static_cast<int64_t>(cs) * kClusterThreads * nv >= J) {
            if (groups * nv > max_gnv) return r;
            if (cs == kMaxCluster && groups * nv > max_gnv / 2) return r;   // 8-CTA clusters need 2 CTAs/SM (<= 64 regs) to pack
            // do not spread a tiny domain over more CTAs than it can feed with >= 1 vector per thread
            while (cs > 1 && static_cast<int64_t>(cs / 2) * kClusterThreads * nv >= J) cs /= 2;
            r.cs = cs; r.nv = nv;
            return r;
        }
    }
    return r;
}

// cluster size: enough CTAs to cover the chip (C * CS >= ~148), power of two, <= 8, and every CTA keeps
// >= 512 vectors per group; 0 -> shape not suited (per-channel domain too large to stay L2-resident)
__host__ inline int pick_cluster(int64_t groups, int64_t n, int64_t c, int64_t hw, bool vec) {
    const int64_t v = vec ? 4 : 1;
    const int64_t J = n * (hw / v);                                   // vectors per (group, channel)
    const int64_t domain_bytes = groups * n * hw * 4;
    int cs = 1;
    while (cs < kMaxCluster && c * cs < sm_count() && J / (cs * 2) >= 512) cs *= 2;
    // concurrently resident footprint must fit L2 for sweep 2 to hit: (CTAs resident) * slice <= ~48 MB
    const int64_t resident = static_cast<int64_t>(sm_count()) * 4;   // 4 x 512 threads per SM
    const int64_t slice = domain_bytes / cs;
    const int64_t ctas = c * cs < resident ? c * cs : resident;
    if (ctas * slice > (int64_t(48) << 20)) return 0;
    return cs;
}

template <typename K, typename... Extra>
int launch_cluster(K kernel, const ClusterParams& p, int cs, cudaStream_t st, const Extra&... extra) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.c * cs);
    cfg.blockDim = dim3(kClusterThreads);
    cfg.stream = st;
    cudaLaunchAttribute attr[2]{};
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;        // every cluster kernel starts with pdl_wait()
    attr[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    if (cudaLaunchKernelEx(&cfg, kernel, p, extra...) != cudaSuccess) { cudaGetLastError(); return AFAN_ERR_LAUNCH; }
    return launch_status();
}

// -----------------------------------------------------------------------------------------------------
// Plane-resident kernels (round 2): small per-channel domains with ANY H*W -- the DeepLab tail at 33 x 33 = 1089
// (BASELINE config 5), where H*W % 4 != 0 sent every launch down scalar two-sweep paths (0.17-0.32 of the HBM roofline).
// One CTA per channel holds the channel's whole domain (groups x n planes) in shared memory: every plane is read from
// HBM exactly once with 128-bit loads -- a plane starts at an arbitrary 4-byte alignment, so it is peeled into a scalar
// head, a 16-byte-aligned body and a scalar tail, and stored in shared memory at the same alignment class -- the
// statistics and the normalisation both run out of shared memory, and the result leaves with the same peeled 128-bit
// pattern.  No cluster, no second HBM read, 2-6 CTAs per SM keep > 64 KB of loads in flight per SM.
// -----------------------------------------------------------------------------------------------------
constexpr int kPlaneThreads = 128;

__device__ __forceinline__ void cp_async16(float* smem_dst, const float* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<unsigned int>(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async4(float* smem_dst, const float* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<unsigned int>(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

struct PlaneParams {
    const float* a;          // fwd: x            bwd: dy
    const float* b;          // fwd: residual     bwd: x
    const float* y;          // bwd + relu: forward output
    float* out;              // fwd: y            bwd: dx
    float* out2;             // bwd: dresidual
    const float* weight;
    const float* bias;
    float* running_mean;
    float* running_var;
    float* save_mean;
    float* save_invstd;
    float* dweight;
    float* dbias;
    double count;
    float eps, momentum;
    int replay;
    unsigned int groups, n, c, hw, pitch;      // pitch: floats per plane slot in shared memory (hw + 8, multiple of 4)
};

// visit every element of one plane: f(global index inside the plane, shared index) with 128-bit accesses on the aligned body
template <typename FV, typename FS>
__device__ __forceinline__ void plane_sweep(const float* gplane, unsigned int hw, unsigned int slot, FV&& vec4, FS&& scalar) {
    const unsigned int mis = static_cast<unsigned int>((reinterpret_cast<uintptr_t>(gplane) >> 2) & 3u);   // floats past a 16-byte line
    const unsigned int head = (4u - mis) & 3u, nbody = hw > head ? (hw - head) >> 2 : 0u, tail0 = head + 4u * nbody;
    const unsigned int s0 = slot + mis;                               // shared index of element 0: same alignment class
    for (unsigned int v = threadIdx.x; v < nbody; v += kPlaneThreads) vec4(head + 4u * v, s0 + head + 4u * v);
    if (threadIdx.x < head && threadIdx.x < hw) scalar(threadIdx.x, s0 + threadIdx.x);
    if (threadIdx.x >= 32 && threadIdx.x - 32 < hw - tail0 && hw > head) scalar(tail0 + threadIdx.x - 32, s0 + tail0 + threadIdx.x - 32);
}

template <bool RELU, bool RES>
__global__ void __launch_bounds__(kPlaneThreads) bn_fwd_plane_kernel(const PlaneParams p) {
    extern __shared__ __align__(16) float sh[];
    __shared__ double scratch[64];
    __shared__ float2 s_tab[16];
    const unsigned int ch = blockIdx.x, planes = p.groups * p.n;
    pdl_wait();
    pdl_launch_dependents();
    // ---- one HBM read of the whole domain into shared memory: asynchronous copies (no register per byte in flight, no
    //      load -> store dependency), so every plane of the channel is requested before the first one has arrived ----
    for (unsigned int pl = 0; pl < planes; ++pl) {
        const float* gp = p.a + (static_cast<size_t>(pl) * p.c + ch) * p.hw;
        plane_sweep(gp, p.hw, pl * p.pitch,
                    [&](unsigned int gi, unsigned int si) { cp_async16(sh + si, gp + gi); },
                    [&](unsigned int gi, unsigned int si) { cp_async4(sh + si, gp + gi); });
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- statistics per group out of shared memory (fixed order: deterministic) ----
    __shared__ double2 s_raw[16];
    for (unsigned int g = 0; g < p.groups; ++g) {
        const float kshift = __ldg(p.a + (static_cast<size_t>(g) * p.n * p.c + ch) * p.hw);     // accumulate x - K (see unshift_sums)
        float a0 = 0.f, a1 = 0.f;
        for (unsigned int pl = g * p.n; pl < (g + 1) * p.n; ++pl) {
            const float* gp = p.a + (static_cast<size_t>(pl) * p.c + ch) * p.hw;
            plane_sweep(gp, p.hw, pl * p.pitch,
                        [&](unsigned int, unsigned int si) {
                            const float4 v = *reinterpret_cast<const float4*>(sh + si);
                            const float tx = v.x - kshift, ty = v.y - kshift, tz = v.z - kshift, tw = v.w - kshift;
                            a0 += (tx + ty) + (tz + tw);
                            a1 = fmaf(tx, tx, fmaf(ty, ty, fmaf(tz, tz, fmaf(tw, tw, a1))));
                        },
                        [&](unsigned int, unsigned int si) { const float t = sh[si] - kshift; a0 += t; a1 = fmaf(t, t, a1); });
        }
        double d0 = a0, d1 = a1;
        block_sum2(d0, d1, scratch);
        if (threadIdx.x == 0) s_raw[g] = unshift_sums(d0, d1, p.count, static_cast<double>(kshift));
        __syncthreads();
    }
    if (threadIdx.x == 0) {                                           // same maths / order as fwd_finalize_channel
        float rm = p.running_mean ? p.running_mean[ch] : 0.f, rv = p.running_var ? p.running_var[ch] : 0.f;
        const float w = p.weight ? p.weight[ch] : 1.f, b = p.bias ? p.bias[ch] : 0.f;
        for (unsigned int g = 0; g < p.groups; ++g) {
            const double mean = s_raw[g].x / p.count;
            double var = s_raw[g].y / p.count - mean * mean;
            var = var < 0.0 ? 0.0 : var;
            const double invstd = rsqrt(var + static_cast<double>(p.eps));
            const double unbiased = p.count > 1.0 ? var * (p.count / (p.count - 1.0)) : var;
            for (int r = 0; r < p.replay; ++r) {
                rm = static_cast<float>((1.0 - p.momentum) * rm + p.momentum * mean);
                rv = static_cast<float>((1.0 - p.momentum) * rv + p.momentum * unbiased);
            }
            p.save_mean[g * p.c + ch] = static_cast<float>(mean);
            p.save_invstd[g * p.c + ch] = static_cast<float>(invstd);
            s_tab[g] = make_float2(static_cast<float>(w * invstd), static_cast<float>(b - mean * w * invstd));
        }
        if (p.running_mean) p.running_mean[ch] = rm;
        if (p.running_var) p.running_var[ch] = rv;
    }
    __syncthreads();
    // ---- normalise (+ residual, + ReLU) out of shared memory, one HBM write ----
    for (unsigned int pl = 0; pl < planes; ++pl) {
        const float2 t = s_tab[pl / p.n];
        const size_t goff = (static_cast<size_t>(pl) * p.c + ch) * p.hw;
        const float* gp = p.a + goff;
        const float* rp = RES ? p.b + goff : nullptr;
        float* op = p.out + goff;
        auto one = [&](float v, float r) { float o = fmaf(v, t.x, t.y); if (RES) o += r; if (RELU) o = fmaxf(o, 0.f); return o; };
        plane_sweep(gp, p.hw, pl * p.pitch,
                    [&](unsigned int gi, unsigned int si) {
                        const float4 v = *reinterpret_cast<const float4*>(sh + si);
                        float4 r = make_float4(0.f, 0.f, 0.f, 0.f);
                        if (RES) r = ld_stream(reinterpret_cast<const float4*>(rp + gi));
                        *reinterpret_cast<float4*>(op + gi) = make_float4(one(v.x, r.x), one(v.y, r.y), one(v.z, r.z), one(v.w, r.w));
                    },
                    [&](unsigned int gi, unsigned int si) { op[gi] = one(sh[si], RES ? ld_stream(rp + gi) : 0.f); });
    }
}

// backward: dy (ReLU-masked on load) and x live in shared memory; sums, coefficients and dx out of shared memory
template <bool RELU, bool DRES>
__global__ void __launch_bounds__(kPlaneThreads) bn_bwd_plane_kernel(const PlaneParams p) {
    extern __shared__ __align__(16) float sh[];
    __shared__ double scratch[64];
    __shared__ float4 s_cf[16];
    __shared__ double2 s_sum[16];
    const unsigned int ch = blockIdx.x, planes = p.groups * p.n;
    float* sd = sh;                                                    // dy (raw: the ReLU mask is applied on use)
    float* sx = sh + static_cast<size_t>(planes) * p.pitch;            // x
    float* sy = sx + static_cast<size_t>(planes) * p.pitch;            // y (RELU only)
    pdl_wait();
    pdl_launch_dependents();
    for (unsigned int pl = 0; pl < planes; ++pl) {                     // asynchronous copies: the whole domain in flight at once
        const size_t goff = (static_cast<size_t>(pl) * p.c + ch) * p.hw;
        const float* dp = p.a + goff;
        const float* xp = p.b + goff;
        const float* yp = RELU ? p.y + goff : nullptr;
        plane_sweep(dp, p.hw, pl * p.pitch,
                    [&](unsigned int gi, unsigned int si) {
                        cp_async16(sd + si, dp + gi);
                        cp_async16(sx + si, xp + gi);
                        if (RELU) cp_async16(sy + si, yp + gi);
                    },
                    [&](unsigned int gi, unsigned int si) {
                        cp_async4(sd + si, dp + gi);
                        cp_async4(sx + si, xp + gi);
                        if (RELU) cp_async4(sy + si, yp + gi);
                    });
    }
    cp_async_wait_all();
    __syncthreads();
    auto m1 = [&](float d, unsigned int si) { return (RELU && !(sy[si] > 0.f)) ? 0.f : d; };
    auto m4 = [&](unsigned int si) {
        float4 d = *reinterpret_cast<const float4*>(sd + si);
        if (RELU) {
            const float4 yy = *reinterpret_cast<const float4*>(sy + si);
            if (!(yy.x > 0.f)) d.x = 0.f;
            if (!(yy.y > 0.f)) d.y = 0.f;
            if (!(yy.z > 0.f)) d.z = 0.f;
            if (!(yy.w > 0.f)) d.w = 0.f;
        }
        return d;
    };
    for (unsigned int g = 0; g < p.groups; ++g) {
        const float mean = p.save_mean[g * p.c + ch];
        float a0 = 0.f, a1 = 0.f;
        for (unsigned int pl = g * p.n; pl < (g + 1) * p.n; ++pl) {
            const float* dp = p.a + (static_cast<size_t>(pl) * p.c + ch) * p.hw;
            plane_sweep(dp, p.hw, pl * p.pitch,
                        [&](unsigned int, unsigned int si) {
                            const float4 d = m4(si), xv = *reinterpret_cast<const float4*>(sx + si);
                            a0 += (d.x + d.y) + (d.z + d.w);
                            a1 = fmaf(d.x, xv.x - mean, fmaf(d.y, xv.y - mean, fmaf(d.z, xv.z - mean, fmaf(d.w, xv.w - mean, a1))));
                        },
                        [&](unsigned int, unsigned int si) { const float d = m1(sd[si], si); a0 += d; a1 = fmaf(d, sx[si] - mean, a1); });
        }
        double d0 = a0, d1 = a1;
        block_sum2(d0, d1, scratch);
        if (threadIdx.x == 0) {
            const unsigned int gc = g * p.c + ch;
            const float invstd = p.save_invstd[gc], w = p.weight ? p.weight[ch] : 1.f;
            const double s_dyxh = d1 * static_cast<double>(invstd);
            s_sum[g] = make_double2(d0, s_dyxh);
            s_cf[g] = make_float4(w * invstd, static_cast<float>(d0 / p.count), static_cast<float>(s_dyxh / p.count * invstd), mean);
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double dw = 0.0, db = 0.0;
        for (unsigned int g = 0; g < p.groups; ++g) { db += s_sum[g].x; dw += s_sum[g].y; }
        if (p.dweight) p.dweight[ch] = static_cast<float>(dw);
        if (p.dbias) p.dbias[ch] = static_cast<float>(db);
    }
    for (unsigned int pl = 0; pl < planes; ++pl) {
        const float4 cf = s_cf[pl / p.n];
        const size_t goff = (static_cast<size_t>(pl) * p.c + ch) * p.hw;
        const float* dp = p.a + goff;
        float* op = p.out + goff;
        float* rp = DRES ? p.out2 + goff : nullptr;
        auto one = [&](float d, float xv) { return cf.x * (d - cf.y - (xv - cf.w) * cf.z); };
        plane_sweep(dp, p.hw, pl * p.pitch,
                    [&](unsigned int gi, unsigned int si) {
                        const float4 d = m4(si), xv = *reinterpret_cast<const float4*>(sx + si);
                        *reinterpret_cast<float4*>(op + gi) = make_float4(one(d.x, xv.x), one(d.y, xv.y), one(d.z, xv.z), one(d.w, xv.w));
                        if (DRES) *reinterpret_cast<float4*>(rp + gi) = d;
                    },
                    [&](unsigned int gi, unsigned int si) { const float d = m1(sd[si], si); op[gi] = one(d, sx[si]); if (DRES) rp[gi] = d; });
    }
}

// applicable when H*W is not a multiple of 4 (the vector paths cover the rest), every tensor is 4-byte aligned in the same
// class per plane (true for same-shaped contiguous tensors whose bases are 16-byte aligned) and the domain fits
__host__ inline size_t plane_smem_bytes(int64_t groups, int64_t n, int64_t hw, int tensors) {
    const int64_t pitch = ((hw + 3) / 4) * 4 + 8;
    return static_cast<size_t>(groups * n * pitch * tensors) * sizeof(float);
}
__host__ inline bool plane_path_ok(int64_t groups, int64_t n, int64_t c, int64_t hw, int tensors, bool bases_aligned) {
    // measured on B200 (graph-replayed launches, L2-cold): 4 x 2048 x 33 x 33 fwd 19.9 us / bwd 40.2 us; 2 x 4 x 256 x 33 x 33 fwd 11.1 us
    // against 23.7 us on the scalar cluster path -> whenever there is at least one CTA (= channel) per SM
    static const int64_t min_c = [] { const char* e = getenv("AFAN_PLANE_MIN_C"); return e ? static_cast<int64_t>(atoll(e)) : int64_t(-1); }();
    return bases_aligned && hw % 4 != 0 && groups <= 16 && c >= (min_c >= 0 ? min_c : static_cast<int64_t>(sm_count())) &&
           plane_smem_bytes(groups, n, hw, tensors) <= 200 * 1024;
}
template <typename K>
__host__ inline int launch_plane(K kernel, const PlaneParams& p, size_t smem, cudaStream_t st) {
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) != cudaSuccess)
        return (cudaGetLastError(), AFAN_ERR_LAUNCH);
    return launch_pdl(kernel, dim3(p.c), dim3(kPlaneThreads), smem, st, p);
}

// ---- host helpers -------------------------------------------------------------------------------
struct BnShape {
    bool ok, vec, aligned, flat; // flat: H*W % 4 != 0, but bases are 16-byte aligned and the element count is a multiple of 4
    unsigned int groups, n, c, hwv, splits, total_v, hw, total;
    int err;
};

__host__ inline BnShape bn_shape(int64_t groups, int64_t n, int64_t c, int64_t hw, bool all_aligned) {
    BnShape s{};
    s.err = AFAN_OK;
    if (groups < 1 || n < 0 || c < 0 || hw < 0) { s.err = AFAN_ERR_SIZE; return s; }
    if (groups > 16 || groups * c > 65535) { s.err = AFAN_ERR_UNSUPPORTED; return s; }
    const int64_t total = groups * n * c * hw;
    if (total >= (int64_t(1) << 32) || n * hw >= (int64_t(1) << 32)) { s.err = AFAN_ERR_UNSUPPORTED; return s; }
    s.vec = all_aligned && (hw % 4 == 0);
    s.aligned = all_aligned;
    s.flat = all_aligned && !s.vec && (total % 4 == 0);
    s.hw = static_cast<unsigned int>(hw);
    s.total = static_cast<unsigned int>(total);
    const int64_t v = s.vec ? 4 : 1;
    s.groups = static_cast<unsigned int>(groups);
    s.n = static_cast<unsigned int>(n);
    s.c = static_cast<unsigned int>(c);
    s.hwv = static_cast<unsigned int>(hw / v);
    s.total_v = static_cast<unsigned int>(total / v);
    // splits of each (g,c) domain: fill 148 x 8 CTAs, but keep >= 256*unroll vectors per CTA
    const int64_t J = n * (hw / v);
    int64_t target = (static_cast<int64_t>(sm_count()) * kCtasPerSm + groups * c - 1) / (groups * c > 0 ? groups * c : 1);
    int64_t by_work = (J + kThreads * kBnUnroll - 1) / (kThreads * kBnUnroll);
    int64_t sp = target < by_work ? target : by_work;
    s.splits = static_cast<unsigned int>(sp < 1 ? 1 : sp);
    s.ok = total > 0;
    return s;
}

__host__ inline int apply_grid(unsigned int total_v) {
    const int64_t want = (static_cast<int64_t>(total_v) + kThreads * kBnUnroll - 1) / (kThreads * kBnUnroll);
    const int64_t cap = static_cast<int64_t>(sm_count()) * kCtasPerSm;
    return static_cast<int>(want < 1 ? 1 : (want < cap ? want : cap));
}

// Split count of the two-launch reduce: the CTAs of one launch should fill WHOLE waves of the kernel's real residency.  Round 1
// took ceil(8 CTAs/SM * SMs / (groups * C)), which lands just above one wave for most shapes (1280 CTAs on 888 resident slots at
// 2x256x64x56x56: two rounds for 1.44 waves of work).  cost(s) = rounds(s) / s; the smallest s within 3 % of the best is taken.
template <auto Kernel>
int reduce_ctas_per_sm() {
    static const int v = [] {
        int n = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, Kernel, kThreads, 0) != cudaSuccess || n < 1) { cudaGetLastError(); n = 1; }
        return n;
    }();
    return v;
}
__host__ inline unsigned int pick_splits(unsigned int domains, unsigned int by_work, int ctas_per_sm) {
    const long long cap = static_cast<long long>(ctas_per_sm) * sm_count();
    long long smax = (static_cast<long long>(kSplitWaves) * sm_count() * kCtasPerSm) / (domains ? domains : 1);
    if (smax > by_work) smax = by_work;
    if (smax < 1) smax = 1;
    double best = 1e30;
    for (long long sp = 1; sp <= smax; ++sp) {
        const double cost = static_cast<double>((domains * sp + cap - 1) / cap) / static_cast<double>(sp);
        if (cost < best) best = cost;
    }
    for (long long sp = 1; sp <= smax; ++sp)
        if (static_cast<double>((domains * sp + cap - 1) / cap) / static_cast<double>(sp) <= best * 1.03) return static_cast<unsigned int>(sp);
    return 1u;
}

template <bool BWD>
int launch_reduce(const ReduceParams& p_in, bool vec, bool relu, cudaStream_t st, bool peel = false) {
    ReduceParams p = p_in;
    const unsigned int domains = p.groups * p.c;
    const unsigned long long J = static_cast<unsigned long long>(p.n) * p.hwv;
    const unsigned int by_work = static_cast<unsigned int>(std::max<unsigned long long>(1ULL, (J + kThreads * kBnUnroll - 1) / (kThreads * kBnUnroll)));
#define AFAN_RED(KERNEL)                                                                  \
    do {                                                                                   \
        p.splits = pick_splits(domains, by_work, reduce_ctas_per_sm<KERNEL>());            \
        KERNEL<<<dim3(p.splits, domains), kThreads, 0, st>>>(p);                           \
    } while (0)
    if (!vec && peel) {
        if (BWD && relu) AFAN_RED((bn_reduce_peel_kernel<BWD, true>));
        else AFAN_RED((bn_reduce_peel_kernel<BWD, false>));
        return launch_status();
    }
    if (vec) {
        if (BWD && relu) AFAN_RED((bn_reduce_kernel<4, BWD, true>));
        else AFAN_RED((bn_reduce_kernel<4, BWD, false>));
    } else {
        if (BWD && relu) AFAN_RED((bn_reduce_kernel<1, BWD, true>));
        else AFAN_RED((bn_reduce_kernel<1, BWD, false>));
    }
#undef AFAN_RED
    return launch_status();
}

// element-wise kernels: 16-byte vectors whenever the bases allow it -- over planes (H*W % 4 == 0) or over the flat index
// space with plane-straddling vectors (apply_slots); returns the launcher's `vec`
__host__ inline bool apply_mode(ApplyParams& p, const BnShape& s) {
    p.c = s.c; p.n = s.n;
    if (s.vec || !s.flat) { p.total_v = s.total_v; p.hwv = s.hwv; p.hw_flat = 0u; return s.vec; }
    p.total_v = s.total / 4u; p.hwv = 0u; p.hw_flat = s.hw;
    return true;
}

int launch_fwd_apply(const ApplyParams& p, bool vec, bool relu, bool res, cudaStream_t st) {
    const int grid = apply_grid(p.total_v);
#define AFAN_FA(V, R, S) bn_fwd_apply_kernel<V, R, S><<<grid, kThreads, 0, st>>>(p)
    if (p.hw_flat) {
        if (relu) { if (res) bn_fwd_apply_kernel<4, true, true, true><<<grid, kThreads, 0, st>>>(p); else bn_fwd_apply_kernel<4, true, false, true><<<grid, kThreads, 0, st>>>(p); }
        else      { if (res) bn_fwd_apply_kernel<4, false, true, true><<<grid, kThreads, 0, st>>>(p); else bn_fwd_apply_kernel<4, false, false, true><<<grid, kThreads, 0, st>>>(p); }
    } else if (vec) { if (relu) { if (res) AFAN_FA(4, true, true); else AFAN_FA(4, true, false); }
               else      { if (res) AFAN_FA(4, false, true); else AFAN_FA(4, false, false); } }
    else     { if (relu) { if (res) AFAN_FA(1, true, true); else AFAN_FA(1, true, false); }
               else      { if (res) AFAN_FA(1, false, true); else AFAN_FA(1, false, false); } }
#undef AFAN_FA
    return launch_status();
}

int launch_bwd_apply(const ApplyParams& p, bool vec, bool relu, bool dres, cudaStream_t st) {
    const int grid = apply_grid(p.total_v);
#define AFAN_BA(V, R, S) bn_bwd_apply_kernel<V, R, S><<<grid, kThreads, 0, st>>>(p)
    if (p.hw_flat) {
        if (relu) { if (dres) bn_bwd_apply_kernel<4, true, true, true><<<grid, kThreads, 0, st>>>(p); else bn_bwd_apply_kernel<4, true, false, true><<<grid, kThreads, 0, st>>>(p); }
        else      { if (dres) bn_bwd_apply_kernel<4, false, true, true><<<grid, kThreads, 0, st>>>(p); else bn_bwd_apply_kernel<4, false, false, true><<<grid, kThreads, 0, st>>>(p); }
    } else if (vec) { if (relu) { if (dres) AFAN_BA(4, true, true); else AFAN_BA(4, true, false); }
               else      { if (dres) AFAN_BA(4, false, true); else AFAN_BA(4, false, false); } }
    else     { if (relu) { if (dres) AFAN_BA(1, true, true); else AFAN_BA(1, true, false); }
               else      { if (dres) AFAN_BA(1, false, true); else AFAN_BA(1, false, false); } }
#undef AFAN_BA
    return launch_status();
}

int launch_affine_bwd(const ApplyParams& p, bool vec, bool relu, bool dres, cudaStream_t st) {
    const int grid = apply_grid(p.total_v);
#define AFAN_AB(V, R, S) bn_affine_bwd_kernel<V, R, S><<<grid, kThreads, 0, st>>>(p)
    if (p.hw_flat) {
        if (relu) { if (dres) bn_affine_bwd_kernel<4, true, true, true><<<grid, kThreads, 0, st>>>(p); else bn_affine_bwd_kernel<4, true, false, true><<<grid, kThreads, 0, st>>>(p); }
        else      { if (dres) bn_affine_bwd_kernel<4, false, true, true><<<grid, kThreads, 0, st>>>(p); else bn_affine_bwd_kernel<4, false, false, true><<<grid, kThreads, 0, st>>>(p); }
    } else if (vec) { if (relu) { if (dres) AFAN_AB(4, true, true); else AFAN_AB(4, true, false); }
               else      { if (dres) AFAN_AB(4, false, true); else AFAN_AB(4, false, false); } }
    else     { if (relu) { if (dres) AFAN_AB(1, true, true); else AFAN_AB(1, true, false); }
               else      { if (dres) AFAN_AB(1, false, true); else AFAN_AB(1, false, false); } }
#undef AFAN_AB
    return launch_status();
}

__host__ inline bool ws_ok(void* ws, int64_t bytes, int64_t groups, int64_t c) {
    return ws && aligned16(ws) && bytes >= bn_layout(groups, c).total;
}
__host__ inline char* wsp(const void* ws, int64_t off) { return static_cast<char*>(const_cast<void*>(ws)) + off; }

}  // namespace afan

using namespace afan;

AFAN_EXPORT int64_t afan_bn_workspace_bytes(int64_t groups, int64_t channels) {
    if (groups < 1 || channels < 0) return AFAN_ERR_SIZE;
    return bn_layout(groups, channels).total;
}

// ---- forward ------------------------------------------------------------------------------------
static int bn_fwd_reduce_impl(const float* x, double* sums_out, bool finalize, const float* weight, const float* bias,
                              float* running_mean, float* running_var, float* save_mean, float* save_invstd,
                              void* ws, int64_t ws_bytes, int64_t groups, int64_t n, int64_t c, int64_t hw, float eps,
                              float momentum, int replay, cudaStream_t st, BnShape* shape_out) {
    BnShape s = bn_shape(groups, n, c, hw, aligned16(x));
    if (s.err != AFAN_OK) return s.err;
    *shape_out = s;
    if (!s.ok) return AFAN_OK;
    if (!x || (finalize && (!save_mean || !save_invstd))) return AFAN_ERR_NULL;
    if (!ws_ok(ws, ws_bytes, groups, c)) return AFAN_ERR_WORKSPACE;
    const BnLayout l = bn_layout(groups, c);
    ReduceParams p{};
    p.a = x;
    p.partials = reinterpret_cast<double2*>(wsp(ws, l.partials));
    p.counters = reinterpret_cast<unsigned int*>(wsp(ws, l.counters));
    p.sums_out = sums_out;
    p.do_finalize = finalize ? 1 : 0;
    p.weight = weight; p.bias = bias; p.running_mean = running_mean; p.running_var = running_var;
    p.out_mean = save_mean; p.out_invstd = save_invstd;
    p.table = reinterpret_cast<float2*>(wsp(ws, l.table));
    p.count = static_cast<double>(n) * static_cast<double>(hw);
    p.eps = eps; p.momentum = momentum; p.replay = replay;
    p.groups = s.groups; p.n = s.n; p.c = s.c; p.hwv = s.hwv; p.splits = s.splits;
    return launch_reduce<false>(p, s.vec, false, st, s.aligned && hw >= 256);
}

AFAN_EXPORT int afan_bn_fwd_apply_f32(const float* x, const float* residual, float* y, const void* workspace,
                                      int64_t workspace_bytes, int64_t groups, int64_t n, int64_t c, int64_t hw,
                                      int relu, afan_stream_t stream) {
    BnShape s = bn_shape(groups, n, c, hw, aligned16(x) && aligned16(y) && (!residual || aligned16(residual)));
    if (s.err != AFAN_OK) return s.err;
    if (!s.ok) return AFAN_OK;
    if (!x || !y) return AFAN_ERR_NULL;
    if (!ws_ok(const_cast<void*>(workspace), workspace_bytes, groups, c)) return AFAN_ERR_WORKSPACE;
    ApplyParams p{};
    p.x = x; p.b = residual; p.out = y;
    p.table = wsp(workspace, bn_layout(groups, c).table);
    const bool vec = apply_mode(p, s);
    return launch_fwd_apply(p, vec, relu != 0, residual != nullptr, static_cast<cudaStream_t>(stream));
}

AFAN_EXPORT int afan_bn_fwd_f32(const float* x, const float* residual, const float* weight, const float* bias,
                                float* running_mean, float* running_var, float* y, float* save_mean,
                                float* save_invstd, void* workspace, int64_t workspace_bytes, int64_t groups,
                                int64_t n, int64_t c, int64_t hw, float eps, float momentum, int relu, int replay,
                                afan_stream_t stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool al = aligned16(x) && aligned16(y) && (!residual || aligned16(residual));
    BnShape s = bn_shape(groups, n, c, hw, al);
    if (s.err != AFAN_OK) return s.err;
    if (!s.ok) return AFAN_OK;
    if (!x || !y || !save_mean || !save_invstd) return AFAN_ERR_NULL;
    if (!ws_ok(workspace, workspace_bytes, groups, c)) return AFAN_ERR_WORKSPACE;   // uniform contract on both paths
    if (plane_path_ok(groups, n, c, hw, 1, al)) {                    // odd H*W, small domain: plane-resident single pass
        PlaneParams pp{};
        pp.a = x; pp.b = residual; pp.out = y; pp.weight = weight; pp.bias = bias;
        pp.running_mean = running_mean; pp.running_var = running_var; pp.save_mean = save_mean; pp.save_invstd = save_invstd;
        pp.count = static_cast<double>(n) * static_cast<double>(hw);
        pp.eps = eps; pp.momentum = momentum; pp.replay = replay;
        pp.groups = static_cast<unsigned int>(groups); pp.n = static_cast<unsigned int>(n); pp.c = static_cast<unsigned int>(c);
        pp.hw = static_cast<unsigned int>(hw); pp.pitch = static_cast<unsigned int>(((hw + 3) / 4) * 4 + 8);
        const size_t smem = plane_smem_bytes(groups, n, hw, 1);
        const bool r = relu != 0, rs = residual != nullptr;
        if (r) return rs ? launch_plane(bn_fwd_plane_kernel<true, true>, pp, smem, st) : launch_plane(bn_fwd_plane_kernel<true, false>, pp, smem, st);
        return rs ? launch_plane(bn_fwd_plane_kernel<false, true>, pp, smem, st) : launch_plane(bn_fwd_plane_kernel<false, false>, pp, smem, st);
    }
    const int cs = pick_cluster(groups, n, c, hw, s.vec);
    const RegPlan rp = pick_reg_plan(groups, n, c, hw, s.vec, 16);
    if (cs > 0 || rp.nv > 0) {                                       // single-launch cluster paths
        ClusterParams p{};
        p.a = x; p.b = residual; p.out = y; p.weight = weight; p.bias = bias;
        p.running_mean = running_mean; p.running_var = running_var; p.save_mean = save_mean; p.save_invstd = save_invstd;
        p.count = static_cast<double>(n) * static_cast<double>(hw);
        p.eps = eps; p.momentum = momentum; p.replay = replay;
        p.groups = s.groups; p.n = s.n; p.c = s.c; p.hwv = s.hwv;
        const bool r = relu != 0, rs = residual != nullptr;
        if (rp.nv > 0) {                                             // register-resident: x read from HBM exactly once
            const P2PParams noq{};
#define AFAN_RF4(G_, NV_) { if (r) { if (rs) return launch_cluster(bn_fwd_cluster_reg_kernel<G_, NV_, true, true, false>, p, rp.cs, st, noq);   \
                                     return launch_cluster(bn_fwd_cluster_reg_kernel<G_, NV_, true, false, false>, p, rp.cs, st, noq); }        \
                            if (rs) return launch_cluster(bn_fwd_cluster_reg_kernel<G_, NV_, false, true, false>, p, rp.cs, st, noq);           \
                            return launch_cluster(bn_fwd_cluster_reg_kernel<G_, NV_, false, false, false>, p, rp.cs, st, noq); }
#define AFAN_RFN(G_) { switch (rp.nv) { case 1: AFAN_RF4(G_, 1) case 2: AFAN_RF4(G_, 2) case 4: AFAN_RF4(G_, 4) default: AFAN_RF4(G_, 8) } }
            if (groups == 1) AFAN_RFN(1) else AFAN_RFN(2)
#undef AFAN_RFN
#undef AFAN_RF4
        }
#define AFAN_CF(V, R, S) return launch_cluster(bn_fwd_cluster_kernel<V, R, S>, p, cs, st)
        if (cs <= 0) goto two_launch_fwd;
        if (s.vec) { if (r) { if (rs) AFAN_CF(4, true, true); else AFAN_CF(4, true, false); }
                     else   { if (rs) AFAN_CF(4, false, true); else AFAN_CF(4, false, false); } }
        else       { if (r) { if (rs) AFAN_CF(1, true, true); else AFAN_CF(1, true, false); }
                     else   { if (rs) AFAN_CF(1, false, true); else AFAN_CF(1, false, false); } }
#undef AFAN_CF
    }
two_launch_fwd:
    // large per-channel domains: two-launch path (global partials + last-CTA finalise, then apply)
    int rc = bn_fwd_reduce_impl(x, nullptr, true, weight, bias, running_mean, running_var, save_mean, save_invstd,
                                workspace, workspace_bytes, groups, n, c, hw, eps, momentum, replay, st, &s);
    if (rc != AFAN_OK || !s.ok) return rc;
    return afan_bn_fwd_apply_f32(x, residual, y, workspace, workspace_bytes, groups, n, c, hw, relu, stream);
}

AFAN_EXPORT int afan_bn_fwd_stats_f32(const float* x, double* sums, void* workspace, int64_t workspace_bytes,
                                      int64_t groups, int64_t n, int64_t c, int64_t hw, afan_stream_t stream) {
    if (!sums) return AFAN_ERR_NULL;
    BnShape s{};
    return bn_fwd_reduce_impl(x, sums, false, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, workspace,
                              workspace_bytes, groups, n, c, hw, 0.f, 0.f, 0, static_cast<cudaStream_t>(stream), &s);
}

AFAN_EXPORT int afan_bn_fwd_finalize_f32(const double* sums, double count, const float* weight, const float* bias,
                                         float* running_mean, float* running_var, float* save_mean,
                                         float* save_invstd, void* workspace, int64_t workspace_bytes, int64_t groups,
                                         int64_t c, float eps, float momentum, int replay, afan_stream_t stream) {
    if (groups < 1 || c < 0 || !(count > 0)) return AFAN_ERR_SIZE;
    if (groups > 16) return AFAN_ERR_UNSUPPORTED;
    if (c == 0) return AFAN_OK;
    if (!sums || !save_mean || !save_invstd) return AFAN_ERR_NULL;
    if (!ws_ok(workspace, workspace_bytes, groups, c)) return AFAN_ERR_WORKSPACE;
    ReduceParams p{};
    p.weight = weight; p.bias = bias; p.running_mean = running_mean; p.running_var = running_var;
    p.out_mean = save_mean; p.out_invstd = save_invstd;
    p.table = reinterpret_cast<float2*>(wsp(workspace, bn_layout(groups, c).table));
    p.count = count; p.eps = eps; p.momentum = momentum; p.replay = replay;
    p.groups = static_cast<unsigned int>(groups); p.c = static_cast<unsigned int>(c);
    bn_fwd_finalize_kernel<<<static_cast<unsigned int>((c + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p, sums);
    return launch_status();
}

AFAN_EXPORT int afan_bn_affine_f32(const float* x, const float* residual, const float* scale_shift, float* y,
                                   int64_t n, int64_t c, int64_t hw, int relu, afan_stream_t stream) {
    BnShape s = bn_shape(1, n, c, hw, aligned16(x) && aligned16(y) && (!residual || aligned16(residual)));
    if (s.err != AFAN_OK) return s.err;
    if (!s.ok) return AFAN_OK;
    if (!x || !y || !scale_shift) return AFAN_ERR_NULL;
    ApplyParams p{};
    p.x = x; p.b = residual; p.out = y; p.table = scale_shift;
    const bool vec = apply_mode(p, s);
    return launch_fwd_apply(p, vec, relu != 0, residual != nullptr, static_cast<cudaStream_t>(stream));
}

AFAN_EXPORT int afan_bn_affine_bwd_f32(const float* dy, const float* y, const float* scale_shift, float* dx, float* dresidual,
                                       int64_t n, int64_t c, int64_t hw, int relu, afan_stream_t stream) {
    BnShape s = bn_shape(1, n, c, hw, aligned16(dy) && aligned16(dx) && (!relu || aligned16(y)) && (!dresidual || aligned16(dresidual)));
    if (s.err != AFAN_OK) return s.err;
    if (!s.ok) return AFAN_OK;
    if (!dy || !dx || !scale_shift || (relu && !y)) return AFAN_ERR_NULL;
    ApplyParams p{};
    p.x = dy; p.y = y; p.out = dx; p.out2 = dresidual; p.table = scale_shift;
    const bool vec = apply_mode(p, s);
    return launch_affine_bwd(p, vec, relu != 0, dresidual != nullptr, static_cast<cudaStream_t>(stream));
}

// ---- backward -----------------------------------------------------------------------------------
static int bn_bwd_reduce_impl(const float* dy, const float* x, const float* y, const float* weight,
                              const float* save_mean, const float* save_invstd, double* sums_out, bool finalize,
                              float* dweight, float* dbias, void* ws, int64_t ws_bytes, int64_t groups, int64_t n,
                              int64_t c, int64_t hw, int relu, bool aligned_all, cudaStream_t st, BnShape* shape_out) {
    BnShape s = bn_shape(groups, n, c, hw, aligned_all);
    if (s.err != AFAN_OK) return s.err;
    *shape_out = s;
    if (!s.ok) return AFAN_OK;
    if (!dy || !x || !save_mean || !save_invstd || (relu && !y)) return AFAN_ERR_NULL;
    if (!ws_ok(ws, ws_bytes, groups, c)) return AFAN_ERR_WORKSPACE;
    const BnLayout l = bn_layout(groups, c);
    ReduceParams p{};
    p.a = dy; p.b = x; p.y = y; p.save_mean = save_mean; p.save_invstd = save_invstd;
    p.partials = reinterpret_cast<double2*>(wsp(ws, l.partials));
    p.counters = reinterpret_cast<unsigned int*>(wsp(ws, l.counters));
    p.sums_out = sums_out;
    p.do_finalize = finalize ? 1 : 0;
    p.weight = weight;
    p.coef = reinterpret_cast<float4*>(wsp(ws, l.coef));
    p.dweight = dweight; p.dbias = dbias;
    p.count = static_cast<double>(n) * static_cast<double>(hw);
    p.groups = s.groups; p.n = s.n; p.c = s.c; p.hwv = s.hwv; p.splits = s.splits;
    return launch_reduce<true>(p, s.vec, relu != 0, st, s.aligned && hw >= 256);
}

AFAN_EXPORT int afan_bn_bwd_apply_f32(const float* dy, const float* x, const float* y, float* dx, float* dresidual,
                                      const void* workspace, int64_t workspace_bytes, int64_t groups, int64_t n,
                                      int64_t c, int64_t hw, int relu, afan_stream_t stream) {
    const bool al = aligned16(dy) && aligned16(x) && aligned16(dx) && (!y || aligned16(y)) &&
                    (!dresidual || aligned16(dresidual));
    BnShape s = bn_shape(groups, n, c, hw, al);
    if (s.err != AFAN_OK) return s.err;
    if (!s.ok) return AFAN_OK;
    if (!dy || !x || !dx || (relu && !y)) return AFAN_ERR_NULL;
    if (!ws_ok(const_cast<void*>(workspace), workspace_bytes, groups, c)) return AFAN_ERR_WORKSPACE;
    ApplyParams p{};
    p.x = dy; p.b = x; p.y = y; p.out = dx; p.out2 = dresidual;
    p.table = wsp(workspace, bn_layout(groups, c).coef);
    const bool vec = apply_mode(p, s);
    return launch_bwd_apply(p, vec, relu != 0, dresidual != nullptr, static_cast<cudaStream_t>(stream));
}

AFAN_EXPORT int afan_bn_bwd_f32(const float* dy, const float* x, const float* y, const float* weight,
                                const float* save_mean, const float* save_invstd, float* dx, float* dresidual,
                                float* dweight, float* dbias, void* workspace, int64_t workspace_bytes,
                                int64_t groups, int64_t n, int64_t c, int64_t hw, int relu, afan_stream_t stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool al = aligned16(dy) && aligned16(x) && aligned16(dx) && (!y || aligned16(y)) &&
                    (!dresidual || aligned16(dresidual));
    BnShape s = bn_shape(groups, n, c, hw, al);
    if (s.err != AFAN_OK) return s.err;
    if (!s.ok) return AFAN_OK;
    if (!dy || !x || !dx || !save_mean || !save_invstd || (relu && !y)) return AFAN_ERR_NULL;
    if (!ws_ok(workspace, workspace_bytes, groups, c)) return AFAN_ERR_WORKSPACE;
    if (plane_path_ok(groups, n, c, hw, relu ? 3 : 2, al)) {         // odd H*W, small domain: plane-resident single pass
        PlaneParams pp{};
        pp.a = dy; pp.b = x; pp.y = y; pp.out = dx; pp.out2 = dresidual; pp.weight = weight;
        pp.save_mean = const_cast<float*>(save_mean); pp.save_invstd = const_cast<float*>(save_invstd);
        pp.dweight = dweight; pp.dbias = dbias;
        pp.count = static_cast<double>(n) * static_cast<double>(hw);
        pp.groups = static_cast<unsigned int>(groups); pp.n = static_cast<unsigned int>(n); pp.c = static_cast<unsigned int>(c);
        pp.hw = static_cast<unsigned int>(hw); pp.pitch = static_cast<unsigned int>(((hw + 3) / 4) * 4 + 8);
        const size_t smem = plane_smem_bytes(groups, n, hw, relu ? 3 : 2);
        const bool r = relu != 0, dr = dresidual != nullptr;
        if (r) return dr ? launch_plane(bn_bwd_plane_kernel<true, true>, pp, smem, st) : launch_plane(bn_bwd_plane_kernel<true, false>, pp, smem, st);
        return dr ? launch_plane(bn_bwd_plane_kernel<false, true>, pp, smem, st) : launch_plane(bn_bwd_plane_kernel<false, false>, pp, smem, st);
    }
    const int cs = pick_cluster(groups * 3, n, c, hw, s.vec);        // three tensors are swept twice
    const RegPlan rp = pick_reg_plan(groups, n, c, hw, s.vec, 8);    // dy and x are both held: G*NV <= 8
    if (cs > 0 || rp.nv > 0) {
        ClusterParams p{};
        p.a = dy; p.b = x; p.y = y; p.out = dx; p.out2 = dresidual; p.weight = weight;
        p.save_mean = const_cast<float*>(save_mean); p.save_invstd = const_cast<float*>(save_invstd);
        p.dweight = dweight; p.dbias = dbias;
        p.count = static_cast<double>(n) * static_cast<double>(hw);
        p.groups = s.groups; p.n = s.n; p.c = s.c; p.hwv = s.hwv;
        const bool r = relu != 0, dr = dresidual != nullptr;
        if (rp.nv > 0) {
            const P2PParams noq{};
#define AFAN_RB4(G_, NV_) { if (r) { if (dr) return launch_cluster(bn_bwd_cluster_reg_kernel<G_, NV_, true, true, false>, p, rp.cs, st, noq);   \
                                     return launch_cluster(bn_bwd_cluster_reg_kernel<G_, NV_, true, false, false>, p, rp.cs, st, noq); }        \
                            if (dr) return launch_cluster(bn_bwd_cluster_reg_kernel<G_, NV_, false, true, false>, p, rp.cs, st, noq);           \
                            return launch_cluster(bn_bwd_cluster_reg_kernel<G_, NV_, false, false, false>, p, rp.cs, st, noq); }
            if (groups == 1) { switch (rp.nv) { case 1: AFAN_RB4(1, 1) case 2: AFAN_RB4(1, 2) case 4: AFAN_RB4(1, 4) default: AFAN_RB4(1, 8) } }
            else             { switch (rp.nv) { case 1: AFAN_RB4(2, 1) case 2: AFAN_RB4(2, 2) default: AFAN_RB4(2, 4) } }
#undef AFAN_RB4
        }
#define AFAN_CB(V, R, S) return launch_cluster(bn_bwd_cluster_kernel<V, R, S>, p, cs, st)
        if (cs <= 0) goto two_launch_bwd;
        if (s.vec) { if (r) { if (dr) AFAN_CB(4, true, true); else AFAN_CB(4, true, false); }
                     else   { if (dr) AFAN_CB(4, false, true); else AFAN_CB(4, false, false); } }
        else       { if (r) { if (dr) AFAN_CB(1, true, true); else AFAN_CB(1, true, false); }
                     else   { if (dr) AFAN_CB(1, false, true); else AFAN_CB(1, false, false); } }
#undef AFAN_CB
    }
two_launch_bwd:
    const bool al1 = aligned16(dy) && aligned16(x) && (!y || aligned16(y));
    int rc = bn_bwd_reduce_impl(dy, x, y, weight, save_mean, save_invstd, nullptr, true, dweight, dbias, workspace,
                                workspace_bytes, groups, n, c, hw, relu, al1, st, &s);
    if (rc != AFAN_OK || !s.ok) return rc;
    return afan_bn_bwd_apply_f32(dy, x, y, dx, dresidual, workspace, workspace_bytes, groups, n, c, hw, relu, stream);
}

/* Backward of BatchNorm + ReLU whose forward output was never materialised (the convolution that consumed it applied the
 * normalisation while loading, afan_conv3x3_umma_bn_f32): the ReLU mask is recomputed from x and the forward's
 * (scale, shift) table.  Register-resident cluster plan only (every tail shape of the CIFAR ResNets). */
AFAN_EXPORT int afan_bn_bwd_xmask_f32(const float* dy, const float* x, const float* mask_table, const float* weight,
                                      const float* save_mean, const float* save_invstd, float* dx, float* dweight,
                                      float* dbias, int64_t groups, int64_t n, int64_t c, int64_t hw, afan_stream_t stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool al = aligned16(dy) && aligned16(x) && aligned16(dx);
    BnShape s = bn_shape(groups, n, c, hw, al);
    if (s.err != AFAN_OK) return s.err;
    if (!s.ok) return AFAN_OK;
    if (!dy || !x || !dx || !save_mean || !save_invstd || !mask_table) return AFAN_ERR_NULL;
    const RegPlan rp = pick_reg_plan(groups, n, c, hw, s.vec, 8);
    if (rp.nv == 0) return AFAN_ERR_UNSUPPORTED;
    ClusterParams p{};
    p.a = dy; p.b = x; p.y = nullptr; p.out = dx; p.out2 = nullptr; p.weight = weight;
    p.save_mean = const_cast<float*>(save_mean); p.save_invstd = const_cast<float*>(save_invstd);
    p.dweight = dweight; p.dbias = dbias;
    p.count = static_cast<double>(n) * static_cast<double>(hw);
    p.groups = s.groups; p.n = s.n; p.c = s.c; p.hwv = s.hwv;
    p.mask_table = reinterpret_cast<const float2*>(mask_table);
    const P2PParams noq{};
#define AFAN_XB(G_, NV_) return launch_cluster(bn_bwd_cluster_reg_kernel<G_, NV_, true, false, false>, p, rp.cs, st, noq);
    if (groups == 1) { switch (rp.nv) { case 1: AFAN_XB(1, 1) case 2: AFAN_XB(1, 2) case 4: AFAN_XB(1, 4) default: AFAN_XB(1, 8) } }
    else             { switch (rp.nv) { case 1: AFAN_XB(2, 1) case 2: AFAN_XB(2, 2) default: AFAN_XB(2, 4) } }
#undef AFAN_XB
}

/* afan_bn_bwd_xmask_f32 with the statistics of the GLOBAL batch (fused NVLink exchange, like afan_bn_bwd_p2p_f32). */
AFAN_EXPORT int afan_bn_bwd_xmask_p2p_f32(const float* dy, const float* x, const float* mask_table, const float* weight,
                                          const float* save_mean, const float* save_invstd, float* dx, float* dweight,
                                          float* dbias, int64_t groups, int64_t n, int64_t c, int64_t hw, int world, int rank,
                                          void* const* peer_mailboxes, int64_t cmax, void* state, afan_stream_t stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool al = aligned16(dy) && aligned16(x) && aligned16(dx);
    BnShape s = bn_shape(groups, n, c, hw, al);
    if (s.err != AFAN_OK) return s.err;
    if (!s.ok) return AFAN_ERR_UNSUPPORTED;
    if (!dy || !x || !dx || !save_mean || !save_invstd || !mask_table) return AFAN_ERR_NULL;
    const RegPlan rp = pick_reg_plan(groups, n, c, hw, s.vec, 8);
    if (rp.nv == 0) return AFAN_ERR_UNSUPPORTED;
    P2PParams q{};
    const int rc = fill_p2p(q, world, rank, peer_mailboxes, cmax, state, c);
    if (rc != AFAN_OK) return rc;
    ClusterParams p{};
    p.a = dy; p.b = x; p.y = nullptr; p.out = dx; p.out2 = nullptr; p.weight = weight;
    p.save_mean = const_cast<float*>(save_mean); p.save_invstd = const_cast<float*>(save_invstd);
    p.dweight = dweight; p.dbias = dbias;
    p.count = static_cast<double>(n) * static_cast<double>(hw) * world;
    p.groups = s.groups; p.n = s.n; p.c = s.c; p.hwv = s.hwv;
    p.mask_table = reinterpret_cast<const float2*>(mask_table);
#define AFAN_XBP(G_, NV_) return launch_cluster(bn_bwd_cluster_reg_kernel<G_, NV_, true, false, true>, p, rp.cs, st, q);
    if (groups == 1) { switch (rp.nv) { case 1: AFAN_XBP(1, 1) case 2: AFAN_XBP(1, 2) case 4: AFAN_XBP(1, 4) default: AFAN_XBP(1, 8) } }
    else             { switch (rp.nv) { case 1: AFAN_XBP(2, 1) case 2: AFAN_XBP(2, 2) default: AFAN_XBP(2, 4) } }
#undef AFAN_XBP
}

AFAN_EXPORT int afan_bn_bwd_reduce_f32(const float* dy, const float* x, const float* y, const float* save_mean,
                                       const float* save_invstd, double* sums, float* dweight, float* dbias,
                                       void* workspace, int64_t workspace_bytes, int64_t groups, int64_t n, int64_t c,
                                       int64_t hw, int relu, afan_stream_t stream) {
    if (!sums) return AFAN_ERR_NULL;
    BnShape s{};
    const bool al = aligned16(dy) && aligned16(x) && (!y || aligned16(y));
    return bn_bwd_reduce_impl(dy, x, y, nullptr, save_mean, save_invstd, sums, false, dweight, dbias, workspace,
                              workspace_bytes, groups, n, c, hw, relu, al, static_cast<cudaStream_t>(stream), &s);
}

AFAN_EXPORT int afan_bn_bwd_finalize_f32(const double* sums, double count, const float* weight, const float* save_mean,
                                         const float* save_invstd, void* workspace, int64_t workspace_bytes,
                                         int64_t groups, int64_t c, afan_stream_t stream) {
    if (groups < 1 || c < 0 || !(count > 0)) return AFAN_ERR_SIZE;
    if (groups > 16) return AFAN_ERR_UNSUPPORTED;
    if (c == 0) return AFAN_OK;
    if (!sums || !save_mean || !save_invstd) return AFAN_ERR_NULL;
    if (!ws_ok(workspace, workspace_bytes, groups, c)) return AFAN_ERR_WORKSPACE;
    ReduceParams p{};
    p.weight = weight; p.save_mean = save_mean; p.save_invstd = save_invstd;
    p.coef = reinterpret_cast<float4*>(wsp(workspace, bn_layout(groups, c).coef));
    p.count = count;
    p.groups = static_cast<unsigned int>(groups); p.c = static_cast<unsigned int>(c);
    bn_bwd_finalize_kernel<<<static_cast<unsigned int>((c + 127) / 128), 128, 0, static_cast<cudaStream_t>(stream)>>>(p, sums);
    return launch_status();
}

// ---- fused multi-GPU entry points (statistics exchanged over NVLink peer memory inside the kernel) -------------
AFAN_EXPORT int64_t afan_bn_mailbox_bytes(int world, int64_t cmax) {
    if (world < 1 || world > kP2PMaxWorld || cmax < 1) return AFAN_ERR_SIZE;
    return p2p_mailbox_bytes(world, cmax);
}

AFAN_EXPORT int afan_bn_fwd_p2p_f32(const float* x, const float* residual, const float* weight, const float* bias,
                                    float* running_mean, float* running_var, float* y, float* save_mean,
                                    float* save_invstd, int64_t groups, int64_t n, int64_t c, int64_t hw, float eps,
                                    float momentum, int relu, int replay, int world, int rank,
                                    void* const* peer_mailboxes, int64_t cmax, void* state, afan_stream_t stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool al = aligned16(x) && aligned16(y) && (!residual || aligned16(residual));
    BnShape s = bn_shape(groups, n, c, hw, al);
    if (s.err != AFAN_OK) return s.err;
    if (!s.ok) return AFAN_ERR_UNSUPPORTED;                          // every rank must take part in every exchange
    if (!x || !y || !save_mean || !save_invstd) return AFAN_ERR_NULL;
    const RegPlan rp = pick_reg_plan(groups, n, c, hw, s.vec, 16);
    if (rp.nv == 0) return AFAN_ERR_UNSUPPORTED;                     // caller falls back to stats -> NCCL -> finalize -> apply
    P2PParams q{};
    int rc = fill_p2p(q, world, rank, peer_mailboxes, cmax, state, c);
    if (rc != AFAN_OK) return rc;
    ClusterParams p{};
    p.a = x; p.b = residual; p.out = y; p.weight = weight; p.bias = bias;
    p.running_mean = running_mean; p.running_var = running_var; p.save_mean = save_mean; p.save_invstd = save_invstd;
    p.count = static_cast<double>(n) * static_cast<double>(hw) * world;   // GLOBAL count
    p.eps = eps; p.momentum = momentum; p.replay = replay;
    p.groups = s.groups; p.n = s.n; p.c = s.c; p.hwv = s.hwv;
    const bool r = relu != 0, rs = residual != nullptr;
#define AFAN_PF4(G_, NV_) { if (r) { if (rs) return launch_cluster(bn_fwd_cluster_reg_kernel<G_, NV_, true, true, true>, p, rp.cs, st, q);   \
                                     return launch_cluster(bn_fwd_cluster_reg_kernel<G_, NV_, true, false, true>, p, rp.cs, st, q); }        \
                            if (rs) return launch_cluster(bn_fwd_cluster_reg_kernel<G_, NV_, false, true, true>, p, rp.cs, st, q);           \
                            return launch_cluster(bn_fwd_cluster_reg_kernel<G_, NV_, false, false, true>, p, rp.cs, st, q); }
#define AFAN_PFN(G_) { switch (rp.nv) { case 1: AFAN_PF4(G_, 1) case 2: AFAN_PF4(G_, 2) case 4: AFAN_PF4(G_, 4) default: AFAN_PF4(G_, 8) } }
    if (groups == 1) AFAN_PFN(1) else AFAN_PFN(2)
#undef AFAN_PFN
#undef AFAN_PF4
}

AFAN_EXPORT int afan_bn_bwd_p2p_f32(const float* dy, const float* x, const float* y, const float* weight,
                                    const float* save_mean, const float* save_invstd, float* dx, float* dresidual,
                                    float* dweight, float* dbias, int64_t groups, int64_t n, int64_t c, int64_t hw,
                                    int relu, int world, int rank, void* const* peer_mailboxes, int64_t cmax,
                                    void* state, afan_stream_t stream) {
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool al = aligned16(dy) && aligned16(x) && aligned16(dx) && (!y || aligned16(y)) &&
                    (!dresidual || aligned16(dresidual));
    BnShape s = bn_shape(groups, n, c, hw, al);
    if (s.err != AFAN_OK) return s.err;
    if (!s.ok) return AFAN_ERR_UNSUPPORTED;
    if (!dy || !x || !dx || !save_mean || !save_invstd || (relu && !y)) return AFAN_ERR_NULL;
    const RegPlan rp = pick_reg_plan(groups, n, c, hw, s.vec, 8);
    if (rp.nv == 0) return AFAN_ERR_UNSUPPORTED;
    P2PParams q{};
    int rc = fill_p2p(q, world, rank, peer_mailboxes, cmax, state, c);
    if (rc != AFAN_OK) return rc;
    ClusterParams p{};
    p.a = dy; p.b = x; p.y = y; p.out = dx; p.out2 = dresidual; p.weight = weight;
    p.save_mean = const_cast<float*>(save_mean); p.save_invstd = const_cast<float*>(save_invstd);
    p.dweight = dweight; p.dbias = dbias;
    p.count = static_cast<double>(n) * static_cast<double>(hw) * world;
    p.groups = s.groups; p.n = s.n; p.c = s.c; p.hwv = s.hwv;
    const bool r = relu != 0, dr = dresidual != nullptr;
#define AFAN_PB4(G_, NV_) { if (r) { if (dr) return launch_cluster(bn_bwd_cluster_reg_kernel<G_, NV_, true, true, true>, p, rp.cs, st, q);   \
                                     return launch_cluster(bn_bwd_cluster_reg_kernel<G_, NV_, true, false, true>, p, rp.cs, st, q); }        \
                            if (dr) return launch_cluster(bn_bwd_cluster_reg_kernel<G_, NV_, false, true, true>, p, rp.cs, st, q);           \
                            return launch_cluster(bn_bwd_cluster_reg_kernel<G_, NV_, false, false, true>, p, rp.cs, st, q); }
    if (groups == 1) { switch (rp.nv) { case 1: AFAN_PB4(1, 1) case 2: AFAN_PB4(1, 2) case 4: AFAN_PB4(1, 4) default: AFAN_PB4(1, 8) } }
    else             { switch (rp.nv) { case 1: AFAN_PB4(2, 1) case 2: AFAN_PB4(2, 2) default: AFAN_PB4(2, 4) } }
#undef AFAN_PB4
}
