// afan_mix.cu -- mix_feature for sm_100a: the reference's adversarial feature NORMALISATION.
//
// Replaces Segmentation/attack_algo.py:121-130 == Detection/attack_algo.py:254-265 (12 un-fused ATen
// launches, ~10 passes over memory):  per pixel (n,h,w), mean and sqrt(unbiased var + 1e-5) over the
// CHANNEL dim of the clean feature are swapped for those of the adversarial feature.
//
// NCHW makes the reduction dim (C) the strided one (stride H*W) while coalescing wants threads along
// W, so a CTA owns a tile of 32*VEC consecutive pixels of one sample: lanes run along pixels (128-bit
// loads when H*W % 4 == 0), the 16-32 warps split the channels.  Sweep 1 keeps one Welford state per
// (pixel, tensor) in registers -- clean and adversarial statistics in the SAME sweep -- then the warps'
// states are merged through shared memory in a fixed order (Chan).  Sweep 2 re-reads the tile (L1/L2
// hot: it was touched microseconds ago) and writes the mixed feature: 12 B/elem of HBM traffic.  The elementwise
// part is (x - m_c) * (s_a / s_c) + m_a: the reference expression with its division folded into a per-pixel ratio.
#include <cstdlib>
#include <type_traits>

#include "afan_common.cuh"

namespace afan {

// scalar path (H*W % 4 != 0, e.g. 33x33): 1024 threads x 8 channels x 2 tensors x 4 B = 64 KB in flight per SM
// vector path: 512 threads x 4 channels x 2 tensors x 16 B = 64 KB in flight per SM

struct Welford {
    float mean = 0.f, m2 = 0.f;
    __device__ __forceinline__ void push(float x, float rcp_count) {
        const float d = x - mean;
        mean = fmaf(d, rcp_count, mean);
        m2 = fmaf(d, x - mean, m2);
    }
};

// NP = pixel groups per thread (group q sits 32*VEC pixels after group q-1): the scalar path takes NP = 2 so that
// every thread owns 4 independent Welford chains (2 pixels x 2 tensors) and twice the loads in flight.
template <int VEC, int NP, int kMixThreads, int kMixUnroll, int kMinBlocks = (kMixThreads <= 512 ? 2 : 1)>
__global__ void __launch_bounds__(kMixThreads, kMinBlocks)
mix_feature_kernel(const float* __restrict__ clean, const float* __restrict__ adv, float* __restrict__ out,
                   unsigned int c, unsigned int hw, unsigned int tiles_per_sample) {
    using V = typename std::conditional<VEC == 4, float4, float>::type;
    constexpr int PT = 32 * VEC * NP;                                // pixels per CTA tile
    constexpr int S = VEC * NP;                                      // pixel slots per thread
    constexpr int kMixWarps = kMixThreads / 32;
    __shared__ float s_mean[2][kMixWarps][PT], s_m2[2][kMixWarps][PT];
    __shared__ float s_stat[4][PT];                                  // mean_cl, std_cl, mean_ad, std_ad

    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int n = blockIdx.x / tiles_per_sample, tile = blockIdx.x - n * tiles_per_sample;
    const unsigned int p0 = tile * PT + lane * VEC;                  // first pixel of group 0 of this lane
    bool active[NP];
#pragma unroll
    for (int q = 0; q < NP; ++q) active[q] = p0 + q * 32 * VEC < hw; // VEC==4 implies hw % 4 == 0: whole vector valid
    const size_t base = static_cast<size_t>(n) * c * hw + p0;

    auto slot_px = [&](int q, int v) { return lane * VEC + q * 32 * VEC + v; };
    auto unpack = [](const V& x, float (&f)[VEC]) {
        if constexpr (VEC == 4) { f[0] = x.x; f[1] = x.y; f[2] = x.z; f[3] = x.w; } else { f[0] = x; }
    };

    // ---- sweep 1: Welford over this warp's channels (warp, warp+W, ...), clean and adv together ----
    // pointer-stepping loops: full groups of kMixUnroll channels run without bounds checks, the tail is peeled
    Welford wc[S], wa[S];
    unsigned int count = 0;
    const size_t cstride = static_cast<size_t>(kMixWarps) * hw;      // elements between two channels of this warp
    const float* pc = clean + base + static_cast<size_t>(warp) * hw;
    const float* pa = adv + base + static_cast<size_t>(warp) * hw;
    const unsigned int my_channels = warp < c ? (c - warp + kMixWarps - 1) / kMixWarps : 0u;
    auto push_all = [&](const V (&xc)[NP], const V (&xa)[NP]) {
        const float r = 1.0f / static_cast<float>(++count);
#pragma unroll
        for (int q = 0; q < NP; ++q) {
            float fc[VEC], fa[VEC];
            unpack(xc[q], fc);
            unpack(xa[q], fa);
#pragma unroll
            for (int v = 0; v < VEC; ++v) { wc[q * VEC + v].push(fc[v], r); wa[q * VEC + v].push(fa[v], r); }
        }
    };
    unsigned int done = 0;
    for (; done + kMixUnroll <= my_channels; done += kMixUnroll) {
        V xc[kMixUnroll][NP] = {}, xa[kMixUnroll][NP] = {};
#pragma unroll
        for (int u = 0; u < kMixUnroll; ++u)
#pragma unroll
            for (int q = 0; q < NP; ++q)
                if (active[q]) {
                    xc[u][q] = *reinterpret_cast<const V*>(pc + u * cstride + q * 32 * VEC);
                    xa[u][q] = *reinterpret_cast<const V*>(pa + u * cstride + q * 32 * VEC);
                }
#pragma unroll
        for (int u = 0; u < kMixUnroll; ++u) push_all(xc[u], xa[u]);
        pc += kMixUnroll * cstride;
        pa += kMixUnroll * cstride;
    }
    for (; done < my_channels; ++done) {
        V xc[NP] = {}, xa[NP] = {};
#pragma unroll
        for (int q = 0; q < NP; ++q)
            if (active[q]) {
                xc[q] = *reinterpret_cast<const V*>(pc + q * 32 * VEC);
                xa[q] = *reinterpret_cast<const V*>(pa + q * 32 * VEC);
            }
        push_all(xc, xa);
        pc += cstride;
        pa += cstride;
    }
#pragma unroll
    for (int q = 0; q < NP; ++q)
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int px = slot_px(q, v), i = q * VEC + v;
            s_mean[0][warp][px] = wc[i].mean; s_m2[0][warp][px] = wc[i].m2;
            s_mean[1][warp][px] = wa[i].mean; s_m2[1][warp][px] = wa[i].m2;
        }
    __syncthreads();

    // ---- merge the warps' states in warp order (Chan et al.), one thread per (pixel, tensor) ----
    for (unsigned int idx = threadIdx.x; idx < 2 * PT; idx += kMixThreads) {
        const unsigned int t = idx / PT, px = idx - t * PT;
        float mean = 0.f, m2 = 0.f, cnt = 0.f;
        for (unsigned int w = 0; w < kMixWarps; ++w) {
            const float nb = w < c ? static_cast<float>((c - w + kMixWarps - 1) / kMixWarps) : 0.f;   // channels warp w saw
            if (nb == 0.f) continue;
            const float mb = s_mean[t][w][px], qb = s_m2[t][w][px];
            const float tot = cnt + nb, d = mb - mean;
            mean = fmaf(d, nb / tot, mean);
            m2 = m2 + qb + d * d * (cnt * nb / tot);
            cnt = tot;
        }
        const float var = m2 / (static_cast<float>(c) - 1.0f);        // torch.var default: unbiased; C == 1 -> NaN
        s_stat[2 * t][px] = mean;
        s_stat[2 * t + 1][px] = sqrtf(var + 1e-5f);
    }
    __syncthreads();

    // ---- sweep 2 (cache-hot): out = (clean - mean_cl) / std_cl * std_adv + mean_adv ----
    // per-pixel coefficients hoisted into registers; (x - m_c) * (s_a / s_c) + m_a is the reference expression with the
    // division folded into one per-pixel ratio (<= 2 ulp from the reference's op order, far inside the 2e-5 statistics tolerance)
    float m_c[S], ratio[S], m_a[S];
#pragma unroll
    for (int q = 0; q < NP; ++q)
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const int px = slot_px(q, v), i = q * VEC + v;
            m_c[i] = s_stat[0][px];
            ratio[i] = __fdiv_rn(s_stat[3][px], s_stat[1][px]);
            m_a[i] = s_stat[2][px];
        }
    pc = clean + base + static_cast<size_t>(warp) * hw;
    float* po = out + base + static_cast<size_t>(warp) * hw;
    auto emit = [&](const V (&xc)[NP], float* dst) {
#pragma unroll
        for (int q = 0; q < NP; ++q)
            if (active[q]) {
                float f[VEC], o[VEC];
                unpack(xc[q], f);
#pragma unroll
                for (int v = 0; v < VEC; ++v) o[v] = fmaf(f[v] - m_c[q * VEC + v], ratio[q * VEC + v], m_a[q * VEC + v]);
                V ov;
                if constexpr (VEC == 4) { ov.x = o[0]; ov.y = o[1]; ov.z = o[2]; ov.w = o[3]; } else { ov = o[0]; }
                st_stream(reinterpret_cast<V*>(dst + q * 32 * VEC), ov);
            }
    };
    done = 0;
    for (; done + kMixUnroll <= my_channels; done += kMixUnroll) {
        V xc[kMixUnroll][NP] = {};
#pragma unroll
        for (int u = 0; u < kMixUnroll; ++u)
#pragma unroll
            for (int q = 0; q < NP; ++q)
                if (active[q]) xc[u][q] = *reinterpret_cast<const V*>(pc + u * cstride + q * 32 * VEC);
#pragma unroll
        for (int u = 0; u < kMixUnroll; ++u) emit(xc[u], po + u * cstride);
        pc += kMixUnroll * cstride;
        po += kMixUnroll * cstride;
    }
    for (; done < my_channels; ++done) {
        V xc[NP] = {};
#pragma unroll
        for (int q = 0; q < NP; ++q)
            if (active[q]) xc[q] = *reinterpret_cast<const V*>(pc + q * 32 * VEC);
        emit(xc, po);
        pc += cstride;
        po += cstride;
    }
}

// ---- a9 + a8 fused: SAT sample points on the clean->adv segment with optional mix_feature per point -------------
// Replaces get_sample_points + the per-point `adv_list[i] = mix_feature(clean, adv_list[i])` lines,
// Segmentation/attack_algo.py:108-118 + main_aug_final.py:206-210 (n = 3), Detection/attack_algo.py:236-245 +
// train_aug_final.py:117-126 (n = 5): up to 4 points p_j = lerp(clean, adv, w_j); for flagged points the output is
// mix_feature(clean, p_j).  clean and adv are read once per sweep for ALL points (8 + 4m B/elem instead of
// (m-1) lerp launches + ~10 passes per mix): the lerps are recomputed in registers, never materialised.
constexpr int kSatMax = 4;
struct SatParams {
    const float* clean;
    const float* adv;
    float* out[kSatMax];
    float w[kSatMax];
    unsigned int mix_mask, m, c, hw, tiles_per_sample;
};

__device__ __forceinline__ float torch_lerp(float x, float y, float w) {
    const float diff = __fsub_rn(y, x);                                  // ATen lerp: w < 0.5 ? x + w*diff : y - diff*(1-w)
    return (fabsf(w) < 0.5f) ? fmaf(w, diff, x) : fmaf(-diff, __fsub_rn(1.0f, w), y);
}

template <int VEC, int kMixThreads, int kMixUnroll>
__global__ void __launch_bounds__(kMixThreads) sat_mix_kernel(const SatParams p) {
    using V = typename std::conditional<VEC == 4, float4, float>::type;
    constexpr int PT = 32 * VEC;
    constexpr int kMixWarps = kMixThreads / 32;
    __shared__ float s_mean[kMixWarps][PT], s_m2[kMixWarps][PT];
    __shared__ float s_stat[1 + kSatMax][2][PT];                         // [clean, p_0..p_3][mean, std][pixel]
    const unsigned int c = p.c, hw = p.hw;
    const unsigned int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const unsigned int n = blockIdx.x / p.tiles_per_sample, tile = blockIdx.x - n * p.tiles_per_sample;
    const unsigned int p0 = tile * PT + lane * VEC;
    const bool active = p0 < hw;
    const size_t base = static_cast<size_t>(n) * c * hw + p0;

    // ---- sweep 1: Welford of clean and of every flagged point, lerps formed in registers ----
    Welford wc[VEC], wp[kSatMax][VEC];
    unsigned int count = 0;
    if (active && p.mix_mask) {
        for (unsigned int k0 = warp; k0 < c; k0 += kMixWarps * kMixUnroll) {
            V xc[kMixUnroll] = {}, xa[kMixUnroll] = {};
#pragma unroll
            for (int u = 0; u < kMixUnroll; ++u) {
                const unsigned int k = k0 + u * kMixWarps;
                if (k < c) {
                    xc[u] = *reinterpret_cast<const V*>(p.clean + base + static_cast<size_t>(k) * hw);
                    xa[u] = *reinterpret_cast<const V*>(p.adv + base + static_cast<size_t>(k) * hw);
                }
            }
#pragma unroll
            for (int u = 0; u < kMixUnroll; ++u) {
                const unsigned int k = k0 + u * kMixWarps;
                if (k < c) {
                    const float r = 1.0f / static_cast<float>(++count);
                    float cv[VEC], av[VEC];
                    if constexpr (VEC == 4) {
                        cv[0] = xc[u].x; cv[1] = xc[u].y; cv[2] = xc[u].z; cv[3] = xc[u].w;
                        av[0] = xa[u].x; av[1] = xa[u].y; av[2] = xa[u].z; av[3] = xa[u].w;
                    } else {
                        cv[0] = xc[u]; av[0] = xa[u];
                    }
#pragma unroll
                    for (int v = 0; v < VEC; ++v) {
                        wc[v].push(cv[v], r);
#pragma unroll
                        for (int j = 0; j < kSatMax; ++j)
                            if (p.mix_mask & (1u << j)) wp[j][v].push(torch_lerp(cv[v], av[v], p.w[j]), r);
                    }
                }
            }
        }
    }
    // ---- merge the warps' states, one tensor at a time through the same shared buffers ----
    if (p.mix_mask) {
        for (int tsr = 0; tsr <= kSatMax; ++tsr) {
            if (tsr > 0 && !(p.mix_mask & (1u << (tsr - 1)))) continue;   // uniform across the CTA
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const Welford& w = tsr == 0 ? wc[v] : wp[tsr - 1][v];
                s_mean[warp][lane * VEC + v] = w.mean;
                s_m2[warp][lane * VEC + v] = w.m2;
            }
            __syncthreads();
            if (threadIdx.x < PT) {
                const unsigned int px = threadIdx.x;
                float mean = 0.f, m2 = 0.f, cnt = 0.f;
                for (unsigned int w = 0; w < kMixWarps; ++w) {
                    const float nb = w < c ? static_cast<float>((c - w + kMixWarps - 1) / kMixWarps) : 0.f;
                    if (nb == 0.f) continue;
                    const float mb = s_mean[w][px], qb = s_m2[w][px];
                    const float tot = cnt + nb, d = mb - mean;
                    mean = fmaf(d, nb / tot, mean);
                    m2 = m2 + qb + d * d * (cnt * nb / tot);
                    cnt = tot;
                }
                s_stat[tsr][0][px] = mean;
                s_stat[tsr][1][px] = sqrtf(m2 / (static_cast<float>(c) - 1.0f) + 1e-5f);
            }
            __syncthreads();
        }
    }
    if (!active) return;

    // ---- sweep 2 (cache-hot): write every point; per-pixel coefficients hoisted into registers ----
    float m_c[VEC], ratio[kSatMax][VEC], m_p[kSatMax][VEC];
#pragma unroll
    for (int v = 0; v < VEC; ++v) {
        const unsigned int px = lane * VEC + v;
        m_c[v] = s_stat[0][0][px];
#pragma unroll
        for (int j = 0; j < kSatMax; ++j) {
            ratio[j][v] = (p.mix_mask & (1u << j)) ? __fdiv_rn(s_stat[1 + j][1][px], s_stat[0][1][px]) : 0.f;
            m_p[j][v] = (p.mix_mask & (1u << j)) ? s_stat[1 + j][0][px] : 0.f;
        }
    }
    const size_t cstride = static_cast<size_t>(kMixWarps) * hw;
    const float* pc = p.clean + base + static_cast<size_t>(warp) * hw;
    const float* pa = p.adv + base + static_cast<size_t>(warp) * hw;
    size_t ooff = base + static_cast<size_t>(warp) * hw;
    const unsigned int my_channels = warp < c ? (c - warp + kMixWarps - 1) / kMixWarps : 0u;
    for (unsigned int done = 0; done < my_channels; done += kMixUnroll) {
        V xc[kMixUnroll] = {}, xa[kMixUnroll] = {};
#pragma unroll
        for (int u = 0; u < kMixUnroll; ++u)
            if (done + u < my_channels) {
                xc[u] = *reinterpret_cast<const V*>(pc + u * cstride);
                xa[u] = *reinterpret_cast<const V*>(pa + u * cstride);
            }
#pragma unroll
        for (int u = 0; u < kMixUnroll; ++u)
            if (done + u < my_channels) {
                float cv[VEC], av[VEC];
                if constexpr (VEC == 4) {
                    cv[0] = xc[u].x; cv[1] = xc[u].y; cv[2] = xc[u].z; cv[3] = xc[u].w;
                    av[0] = xa[u].x; av[1] = xa[u].y; av[2] = xa[u].z; av[3] = xa[u].w;
                } else {
                    cv[0] = xc[u]; av[0] = xa[u];
                }
#pragma unroll
                for (int j = 0; j < kSatMax; ++j) {
                    if (j < static_cast<int>(p.m)) {
                        float o[VEC];
#pragma unroll
                        for (int v = 0; v < VEC; ++v)
                            o[v] = (p.mix_mask & (1u << j)) ? fmaf(cv[v] - m_c[v], ratio[j][v], m_p[j][v])
                                                            : torch_lerp(cv[v], av[v], p.w[j]);
                        V ov;
                        if constexpr (VEC == 4) { ov.x = o[0]; ov.y = o[1]; ov.z = o[2]; ov.w = o[3]; } else { ov = o[0]; }
                        st_stream(reinterpret_cast<V*>(p.out[j] + ooff + u * cstride), ov);
                    }
                }
            }
        pc += kMixUnroll * cstride;
        pa += kMixUnroll * cstride;
        ooff += kMixUnroll * cstride;
    }
}

}  // namespace afan

using namespace afan;

AFAN_EXPORT int afan_mix_feature_f32(const float* clean, const float* adv, float* out, int64_t n, int64_t c,
                                     int64_t hw, afan_stream_t stream) {
    if (n < 0 || c < 0 || hw < 0) return AFAN_ERR_SIZE;
    if (n == 0 || c == 0 || hw == 0) return AFAN_OK;
    if (!clean || !adv || !out) return AFAN_ERR_NULL;
    if (c >= (int64_t(1) << 31) || hw >= (int64_t(1) << 31)) return AFAN_ERR_UNSUPPORTED;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const bool vec = (hw % 4 == 0) && aligned16(clean) && aligned16(adv) && aligned16(out);
    // pixel groups per thread: 2 only when the grid still covers the chip twice over (more ILP per thread), else 1
    const bool np2 = !vec && n * ((hw + 63) / 64) >= 4 * static_cast<int64_t>(sm_count());
    const int64_t pt = vec ? 128 : (np2 ? 64 : 32);
    const int64_t tiles = (hw + pt - 1) / pt;
    if (n * tiles >= (int64_t(1) << 31)) return AFAN_ERR_UNSUPPORTED;
    const unsigned int grid = static_cast<unsigned int>(n * tiles);
    const unsigned int uc = static_cast<unsigned int>(c), uhw = static_cast<unsigned int>(hw), ut = static_cast<unsigned int>(tiles);
    // vector path with many tiles: 256-thread CTAs holding 8 channels x 2 tensors x 16 B per thread in flight (115 registers,
    // one CTA per launch-bound slot): 45.2 -> 41.6 us at 4 x 256 x 128 x 128 (0.68 -> 0.74 of the roofline); the same depth on the
    // scalar path measured slower at every shape (profiles/r2_mix_ncu.md)
    if (vec && grid >= 2u * static_cast<unsigned int>(sm_count()))
        mix_feature_kernel<4, 1, 256, 8, 1><<<grid, 256, 0, st>>>(clean, adv, out, uc, uhw, ut);
    else if (vec)
        mix_feature_kernel<4, 1, 512, 4><<<grid, 512, 0, st>>>(clean, adv, out, uc, uhw, ut);
    else if (grid < 2u * static_cast<unsigned int>(sm_count()))        // few tiles: one fat CTA per SM keeps 64 KB in flight
        mix_feature_kernel<1, 1, 1024, 8><<<grid, 1024, 0, st>>>(clean, adv, out, uc, uhw, ut);
    else if (np2)                                                    // many tiles: 2 pixel groups per thread, 2-3 CTAs per SM
        mix_feature_kernel<1, 2, 512, 4><<<grid, 512, 0, st>>>(clean, adv, out, uc, uhw, ut);
    else
        mix_feature_kernel<1, 1, 512, 8><<<grid, 512, 0, st>>>(clean, adv, out, uc, uhw, ut);
    return launch_status();
}

AFAN_EXPORT int afan_sat_mix_f32(const float* clean, const float* adv, float* const* outs, const float* weights,
                                 const int* mix_flags, int m, int64_t n, int64_t c, int64_t hw, afan_stream_t stream) {
    if (n < 0 || c < 0 || hw < 0 || m < 0) return AFAN_ERR_SIZE;
    if (m > kSatMax) return AFAN_ERR_UNSUPPORTED;
    if (n == 0 || c == 0 || hw == 0 || m == 0) return AFAN_OK;
    if (!clean || !adv || !outs || !weights || !mix_flags) return AFAN_ERR_NULL;
    if (c >= (int64_t(1) << 31) || hw >= (int64_t(1) << 31)) return AFAN_ERR_UNSUPPORTED;
    SatParams p{};
    p.clean = clean; p.adv = adv; p.m = static_cast<unsigned int>(m);
    bool vec = (hw % 4 == 0) && aligned16(clean) && aligned16(adv);
    for (int j = 0; j < m; ++j) {
        if (!outs[j]) return AFAN_ERR_NULL;
        p.out[j] = outs[j];
        p.w[j] = weights[j];
        if (mix_flags[j]) p.mix_mask |= 1u << j;
        vec = vec && aligned16(outs[j]);
    }
    const int64_t pt = vec ? 128 : 32;
    const int64_t tiles = (hw + pt - 1) / pt;
    if (n * tiles >= (int64_t(1) << 31)) return AFAN_ERR_UNSUPPORTED;
    p.c = static_cast<unsigned int>(c); p.hw = static_cast<unsigned int>(hw); p.tiles_per_sample = static_cast<unsigned int>(tiles);
    const unsigned int grid = static_cast<unsigned int>(n * tiles);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (vec) sat_mix_kernel<4, 512, 2><<<grid, 512, 0, st>>>(p);
    else     sat_mix_kernel<1, 512, 8><<<grid, 512, 0, st>>>(p);
    return launch_status();
}
