/* afan_oracle.c -- CPU restatement of the A-FAN adversarial-feature inner loop.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is imported, linked or executed
 * by the product path (cv_a-fan_b200/); only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs use it, as the checker.
 *
 * Each function restates, scalar by scalar and in the reference's own order of
 * floating-point operations, what the reference (VITA-Group/CV_A-FAN, Python/PyTorch)
 * computes; citations are file:line relative to the reference root.  Parity of this
 * file with the reference is PINNED by tests/golden/ (vectors produced by executing
 * the unmodified reference functions, see oracle/gen_golden.py) and, when
 * /root/reference is mounted, by live cross-checks in tests/test_oracle_vs_reference.py.
 *
 * Build: make -C oracle   (gcc -O2 -ffp-contract=off: no FMA contraction, so every
 * fp32 operation below rounds exactly once, as the un-fused ATen ops do).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_API __attribute__((visibility("default")))

/* torch.sign on floats: (0 < a) - (a < 0); sign(NaN) = sign(-0) = +0
 * (probe on torch 2.11 CPU; used at Classification/attack_algo.py:53). */
static inline float orc_sign(float a) { return (float)((0.0f < a) - (a < 0.0f)); }

/* ---- a2: random start, Classification/attack_algo.py:42-44 ---------------------
 * x_adv = x.clone(); x_adv += (2.0 * torch.rand(shape) - 1.0) * eps
 * `u` is the torch.rand draw (CPU generator in the reference). */
ORC_API void orc_pgd_init_noise_f32(const float *x, const float *u, float *x_adv, int64_t n, float eps) {
    for (int64_t i = 0; i < n; ++i) {
        float t = 2.0f * u[i];
        t = t - 1.0f;
        t = t * eps;
        x_adv[i] = x[i] + t;
    }
}

/* ---- a3 + a4: one PGD update, Classification/attack_algo.py:53-56 ---------------
 * x_adv.data.add_(gamma * sign(grad));  if clip: linfball_proj(x, eps, x_adv)
 * linfball_proj -> tensor_clamp (attack_algo.py:9-19,35-36):
 *   min = x - eps; max = x + eps; idx = t < min; t[idx] = min[idx]; idx = t > max; t[idx] = max[idx]
 * (comparisons with NaN are false -> element left alone).  x_clean may be NULL if !clip. */
ORC_API void orc_pgd_linf_step_f32(const float *grad, const float *x_clean, float *x_adv, int64_t n,
                                   float gamma, float eps, int clip) {
    for (int64_t i = 0; i < n; ++i) {
        float v = gamma * orc_sign(grad[i]);
        float t = x_adv[i] + v;
        if (clip) {
            float lo = x_clean[i] - eps;
            float hi = x_clean[i] + eps;
            if (t < lo) t = lo;
            if (t > hi) t = hi;
        }
        x_adv[i] = t;
    }
}

/* ---- bf16-storage twin of the update (BASELINE config 3; NO reference exists for this dtype: unpinned) ----
 * tensors are bf16 bit patterns (uint16); the arithmetic is the fp32 sequence above on the widened values; the
 * result is rounded to nearest-even; delta = bf16(fl(widen(x_adv_new_bf16) - widen(x))). */
static inline float orc_bf16_to_f32(uint16_t h) { uint32_t b = (uint32_t)h << 16; float f; memcpy(&f, &b, 4); return f; }
static inline uint16_t orc_f32_to_bf16(float f) {
    uint32_t b; memcpy(&b, &f, 4);
    if ((b & 0x7fffffffu) > 0x7f800000u) return 0x7fff;                 /* NaN -> canonical (CUDA __float2bfloat16_rn) */
    b += 0x7fffu + ((b >> 16) & 1u);
    return (uint16_t)(b >> 16);
}
ORC_API void orc_pgd_linf_step_bf16(const uint16_t *grad, const uint16_t *x_clean, uint16_t *x_adv,
                                    uint16_t *delta /*nullable*/, int64_t n, float gamma, float eps, int clip) {
    for (int64_t i = 0; i < n; ++i) {
        float v = gamma * orc_sign(orc_bf16_to_f32(grad[i]));
        float t = orc_bf16_to_f32(x_adv[i]) + v;
        float xc = x_clean ? orc_bf16_to_f32(x_clean[i]) : 0.0f;
        if (clip) {
            float lo = xc - eps, hi = xc + eps;
            if (t < lo) t = lo;
            if (t > hi) t = hi;
        }
        x_adv[i] = orc_f32_to_bf16(t);
        if (delta) delta[i] = orc_f32_to_bf16(orc_bf16_to_f32(x_adv[i]) - xc);
    }
}
ORC_API void orc_pgd_init_noise_bf16(const uint16_t *x, const float *u, uint16_t *x_adv, int64_t n, float eps) {
    for (int64_t i = 0; i < n; ++i) {
        float t = 2.0f * u[i];
        t = t - 1.0f;
        t = t * eps;
        x_adv[i] = orc_f32_to_bf16(orc_bf16_to_f32(x[i]) + t);
    }
}

/* ---- a11: perturbation norms, Classification/main_perturb.py:188-192 -------------
 * perturbation = (adv - clean); per-sample torch.norm(p=2) and torch.norm(p=inf).
 * L2 accumulates in double and rounds once (torch's own reduction order is not
 * reproducible; tests allow 2 ulp); Linf is exact.  torch.norm(inf) propagates NaN. */
ORC_API void orc_delta_norms_f32(const float *x_adv, const float *x_clean, float *delta /*nullable*/,
                                 float *l2, float *linf, int64_t n_samples, int64_t per_sample) {
    for (int64_t s = 0; s < n_samples; ++s) {
        double acc = 0.0;
        float mx = 0.0f;
        int has_nan = 0;
        for (int64_t j = 0; j < per_sample; ++j) {
            int64_t i = s * per_sample + j;
            float d = x_adv[i] - x_clean[i];
            if (delta) delta[i] = d;
            acc += (double)d * (double)d;
            float a = fabsf(d);
            if (a != a) has_nan = 1;
            if (a > mx) mx = a;
        }
        l2[s] = (float)sqrt(acc);
        linf[s] = has_nan ? NAN : mx;
    }
}

/* ---- a5: l2ball_proj, Classification/attack_algo.py:21-33 ------------------------
 * direction = t - center; dist = ||direction||_2 per sample; direction /= dist
 * (0/0 -> NaN when t == center, kept); dist[dist > radius] = radius; direction *= dist;
 * t = center + direction. */
ORC_API void orc_l2ball_proj_f32(const float *center, float radius, float *t, int64_t n_samples,
                                 int64_t per_sample) {
    for (int64_t s = 0; s < n_samples; ++s) {
        const float *c = center + s * per_sample;
        float *p = t + s * per_sample;
        double acc = 0.0;
        for (int64_t j = 0; j < per_sample; ++j) {
            float d = p[j] - c[j];
            acc += (double)d * (double)d;
        }
        float dist = (float)sqrt(acc);
        float r = dist;
        if (r > radius) r = radius;
        for (int64_t j = 0; j < per_sample; ++j) {
            float d = p[j] - c[j];
            d = d / dist;
            d = d * r;
            p[j] = c[j] + d;
        }
    }
}

/* ---- a5b: L2-normalised ascent step (north-star item; NO reference implementation:
 * "parity unpinned" by the reference).  Defined by this build as
 *   x_adv += gamma * g / max(||g||_2 per sample, tiny);  if clip: l2ball_proj(x, eps, x_adv)
 * nearest relative: the commented-out untarget_PGD, Detection/attack_algo.py:199. */
ORC_API void orc_pgd_l2_step_f32(const float *grad, const float *x_clean, float *x_adv, int64_t n_samples,
                                 int64_t per_sample, float gamma, float eps, int clip, float tiny) {
    for (int64_t s = 0; s < n_samples; ++s) {
        const float *g = grad + s * per_sample;
        float *p = x_adv + s * per_sample;
        double acc = 0.0;
        for (int64_t j = 0; j < per_sample; ++j) acc += (double)g[j] * (double)g[j];
        float nrm = (float)sqrt(acc);
        if (!(nrm > tiny)) nrm = tiny;
        float scale = gamma / nrm;
        for (int64_t j = 0; j < per_sample; ++j) {
            float v = scale * g[j];
            p[j] = p[j] + v;
        }
    }
    if (clip) orc_l2ball_proj_f32(x_clean, eps, x_adv, n_samples, per_sample);
}

/* ---- a8: mix_feature, Segmentation/attack_algo.py:121-130, Detection/attack_algo.py:254-265
 * per (n, h, w): mean / sqrt(unbiased var + 1e-5) over the CHANNEL dim of clean and adv;
 * out = (clean - mean_cl) / std_cl * std_adv + mean_adv.  Statistics in double (torch's
 * reduction order is not reproducible; tests use 1e-5 rel), elementwise part in fp32 in
 * the reference's order. layout: NCHW contiguous, hw = H*W. */
ORC_API void orc_mix_feature_f32(const float *clean, const float *adv, float *out, int64_t n, int64_t c,
                                 int64_t hw) {
    const double eps = 1e-5;
    for (int64_t b = 0; b < n; ++b)
        for (int64_t p = 0; p < hw; ++p) {
            const float *cl = clean + b * c * hw + p;
            const float *ad = adv + b * c * hw + p;
            double s_cl = 0, s_ad = 0;
            for (int64_t k = 0; k < c; ++k) { s_cl += cl[k * hw]; s_ad += ad[k * hw]; }
            double m_cl = s_cl / (double)c, m_ad = s_ad / (double)c;
            double q_cl = 0, q_ad = 0;
            for (int64_t k = 0; k < c; ++k) {
                double a = cl[k * hw] - m_cl, d = ad[k * hw] - m_ad;
                q_cl += a * a; q_ad += d * d;
            }
            /* torch.var default: unbiased (c-1); c == 1 -> NaN like torch */
            double v_cl = q_cl / (double)(c - 1), v_ad = q_ad / (double)(c - 1);
            float mean_cl = (float)m_cl, mean_ad = (float)m_ad;
            float std_cl = sqrtf((float)v_cl + (float)eps), std_ad = sqrtf((float)v_ad + (float)eps);
            float *o = out + b * c * hw + p;
            for (int64_t k = 0; k < c; ++k) {
                float z = cl[k * hw] - mean_cl;
                z = z / std_cl;
                z = z * std_ad;
                o[k * hw] = z + mean_ad;
            }
        }
}

/* ---- a9: get_sample_points -> torch.lerp(x, y, w), Segmentation/attack_algo.py:108-118
 * ATen lerp: w < 0.5 ? x + w*(y-x) : y - (y-x)*(1-w). */
ORC_API void orc_lerp_f32(const float *x, const float *y, float w, float *out, int64_t n) {
    for (int64_t i = 0; i < n; ++i) {
        float diff = y[i] - x[i];
        out[i] = (fabsf(w) < 0.5f) ? x[i] + w * diff : y[i] - diff * (1.0f - w);
    }
}

/* ---- a10: train-mode BatchNorm2d over G statistic groups (G=1: one batch; G=2: the
 * [adv; clean] pair that the reference feeds through the SAME module in two passes,
 * Classification/main_perturb.py:195-196 + resnet_s.py:54,56,89).  x: [G*N, C, HW].
 * y = relu?( gamma*(x-mean)/sqrt(var_biased+eps) + beta (+ residual) ); running stats
 * updated group after group (pass order), `replay` times each (head-cache replays the
 * reference's double head forward, main_perturb.py:173,196), unbiased var for running. */
ORC_API void orc_bn_fwd_f32(const float *x, const float *residual /*nullable*/, const float *weight,
                            const float *bias, float *running_mean /*nullable*/,
                            float *running_var /*nullable*/, float *y, float *save_mean,
                            float *save_invstd, int64_t groups, int64_t n, int64_t c, int64_t hw,
                            float eps, float momentum, int relu, int replay) {
    for (int64_t g = 0; g < groups; ++g)
        for (int64_t k = 0; k < c; ++k) {
            double s = 0, q = 0;
            double cnt = (double)(n * hw);
            for (int64_t b = g * n; b < (g + 1) * n; ++b) {
                const float *p = x + (b * c + k) * hw;
                for (int64_t j = 0; j < hw; ++j) s += p[j];
            }
            double mean = s / cnt;
            for (int64_t b = g * n; b < (g + 1) * n; ++b) {
                const float *p = x + (b * c + k) * hw;
                for (int64_t j = 0; j < hw; ++j) { double d = p[j] - mean; q += d * d; }
            }
            double var = q / cnt;
            double invstd = 1.0 / sqrt(var + (double)eps);
            save_mean[g * c + k] = (float)mean;
            save_invstd[g * c + k] = (float)invstd;
            if (running_mean && running_var) {
                double unb = (cnt > 1) ? q / (cnt - 1) : var;
                for (int r = 0; r < replay; ++r) {
                    running_mean[k] = (float)((1.0 - momentum) * running_mean[k] + momentum * mean);
                    running_var[k] = (float)((1.0 - momentum) * running_var[k] + momentum * unb);
                }
            }
            float w = weight ? weight[k] : 1.0f, bb = bias ? bias[k] : 0.0f;
            for (int64_t b = g * n; b < (g + 1) * n; ++b) {
                const float *p = x + (b * c + k) * hw;
                const float *rs = residual ? residual + (b * c + k) * hw : NULL;
                float *o = y + (b * c + k) * hw;
                for (int64_t j = 0; j < hw; ++j) {
                    double v = ((double)p[j] - mean) * invstd * w + bb;
                    if (rs) v += rs[j];
                    if (relu && v < 0) v = 0;
                    o[j] = (float)v;
                }
            }
        }
}

/* Backward of the above.  dy_eff = dy * (y > 0) when relu; d_residual = dy_eff;
 * dx = w*invstd*(dy_eff - mean(dy_eff) - xhat*mean(dy_eff*xhat)) per group;
 * dweight = sum_g sum(dy_eff*xhat); dbias = sum_g sum(dy_eff). */
ORC_API void orc_bn_bwd_f32(const float *dy, const float *x, const float *y /*needed if relu*/,
                            const float *weight, const float *save_mean, const float *save_invstd,
                            float *dx, float *dresidual /*nullable*/, float *dweight, float *dbias,
                            int64_t groups, int64_t n, int64_t c, int64_t hw, int relu) {
    for (int64_t k = 0; k < c; ++k) {
        double dw = 0, db = 0;
        for (int64_t g = 0; g < groups; ++g) {
            double mean = save_mean[g * c + k], invstd = save_invstd[g * c + k];
            double s1 = 0, s2 = 0, cnt = (double)(n * hw);
            for (int64_t b = g * n; b < (g + 1) * n; ++b)
                for (int64_t j = 0; j < hw; ++j) {
                    int64_t i = (b * c + k) * hw + j;
                    double d = dy[i];
                    if (relu && !(y[i] > 0)) d = 0;
                    s1 += d;
                    s2 += d * ((double)x[i] - mean) * invstd;
                }
            dw += s2; db += s1;
            double w = weight ? weight[k] : 1.0;
            for (int64_t b = g * n; b < (g + 1) * n; ++b)
                for (int64_t j = 0; j < hw; ++j) {
                    int64_t i = (b * c + k) * hw + j;
                    double d = dy[i];
                    if (relu && !(y[i] > 0)) d = 0;
                    double xh = ((double)x[i] - mean) * invstd;
                    dx[i] = (float)(w * invstd * (d - s1 / cnt - xh * s2 / cnt));
                    if (dresidual) dresidual[i] = (float)d;
                }
        }
        if (dweight) dweight[k] = (float)dw;
        if (dbias) dbias[k] = (float)db;
    }
}

/* ---- a7 tail: torch.optim.SGD(momentum, weight_decay) step, main_perturb.py:72-74,201
 * g += wd*p; buf = mom*buf + g (buf starts at 0 == torch's first-step clone); p -= lr*buf */
ORC_API void orc_sgd_momentum_f32(float *p, const float *g, float *buf, int64_t n, float lr, float momentum,
                                  float weight_decay) {
    for (int64_t i = 0; i < n; ++i) {
        float d = g[i] + weight_decay * p[i];
        float b = momentum * buf[i] + d;
        buf[i] = b;
        p[i] = p[i] - lr * b;
    }
}

/* ---- Philox4x32-10 uniform stream for the on-device random start (fast path of a2;
 * distributionally, not bitwise, equal to the reference's CPU torch.rand).
 * Element i uses counter (i/4 + offset, 0, 0, 0), key (seed_lo, seed_hi), lane i%4;
 * u = (bits >> 8) * 2^-24 in [0,1), the same grid torch.rand uses for float32. */
static inline void philox_round(uint32_t *c, uint32_t *k) {
    const uint64_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
    uint64_t p0 = M0 * c[0], p1 = M1 * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k[0];
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k[1];
    uint32_t n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
}
static inline void philox4x32_10(uint64_t ctr, uint64_t seed, uint32_t out[4]) {
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), 0u, 0u};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    for (int r = 0; r < 10; ++r) {
        if (r) { k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u; }
        philox_round(c, k);
    }
    memcpy(out, c, sizeof(uint32_t) * 4);
}
ORC_API void orc_philox_uniform_f32(float *u, int64_t n, uint64_t seed, uint64_t offset) {
    for (int64_t i = 0; i < n; i += 4) {
        uint32_t r[4];
        philox4x32_10((uint64_t)(i / 4) + offset, seed, r);
        for (int j = 0; j < 4 && i + j < n; ++j) u[i + j] = (float)(r[j] >> 8) * 0x1p-24f;
    }
}

/* ---- f4: greedy NMS, Detection/support/src/cuda/nms.cu:13-21,99-123 (GPU flavour: suppress when IoU > thr;
 * the CPU flavour nms_cpu.cpp:60 uses >=) with the legacy "+1" areas.  order[] = indices sorted by score
 * descending (computed by the caller); keep[i] = 1 for kept ORIGINAL indices.  Returns the number kept. */
ORC_API int64_t orc_nms_f32(const float *boxes, const int64_t *order, int64_t n, float thr, int strict_gt, uint8_t *keep) {
    uint8_t *supp = (uint8_t *)calloc((size_t)(n > 0 ? n : 1), 1);
    int64_t kept = 0;
    for (int64_t i = 0; i < n; ++i) keep[i] = 0;
    for (int64_t _i = 0; _i < n; ++_i) {
        int64_t i = order[_i];
        if (supp[i]) continue;
        keep[i] = 1; ++kept;
        const float *a = boxes + 4 * i;
        float sa = (a[2] - a[0] + 1) * (a[3] - a[1] + 1);
        for (int64_t _j = _i + 1; _j < n; ++_j) {
            int64_t j = order[_j];
            if (supp[j]) continue;
            const float *b = boxes + 4 * j;
            float left = a[0] > b[0] ? a[0] : b[0], right = a[2] < b[2] ? a[2] : b[2];
            float top = a[1] > b[1] ? a[1] : b[1], bottom = a[3] < b[3] ? a[3] : b[3];
            float w = right - left + 1, h = bottom - top + 1;
            if (w < 0) w = 0;
            if (h < 0) h = 0;
            float inter = w * h, sb = (b[2] - b[0] + 1) * (b[3] - b[1] + 1);
            float ovr = inter / (sa + sb - inter);
            if (strict_gt ? (ovr > thr) : (ovr >= thr)) supp[j] = 1;
        }
    }
    free(supp);
    return kept;
}

/* ---- f4: ROIAlign forward / backward, Detection/support/src/cuda/ROIAlign_cuda.cu:15-122,124-254
 * (== Detection/support/src/cpu/ROIAlign_cpu.cpp; the maskrcnn-benchmark kernel that torchvision.ops.roi_align
 * (aligned=False) descends from).  feat [n][c][h][w]; rois [r][5]; out [r][c][ph][pw]. */
static int orc_tap(float y, float x, int h, int w, int off[4], float wt[4]) {
    if (y < -1.0f || y > h || x < -1.0f || x > w) return 0;
    if (y <= 0) y = 0;
    if (x <= 0) x = 0;
    int y0 = (int)y, x0 = (int)x, y1, x1;
    if (y0 >= h - 1) { y1 = y0 = h - 1; y = (float)y0; } else y1 = y0 + 1;
    if (x0 >= w - 1) { x1 = x0 = w - 1; x = (float)x0; } else x1 = x0 + 1;
    float ly = y - y0, lx = x - x0, hy = 1.f - ly, hx = 1.f - lx;
    off[0] = y0 * w + x0; off[1] = y0 * w + x1; off[2] = y1 * w + x0; off[3] = y1 * w + x1;
    wt[0] = hy * hx; wt[1] = hy * lx; wt[2] = ly * hx; wt[3] = ly * lx;
    return 1;
}
ORC_API void orc_roi_align_f32(const float *feat, const float *rois, float *out /*fwd: written*/, const float *dout /*bwd*/,
                               float *dfeat /*bwd: accumulated into (caller zeroes)*/, int64_t c, int64_t h, int64_t w,
                               int64_t r, int64_t ph, int64_t pw, float scale, int sampling_ratio) {
    for (int64_t ri = 0; ri < r; ++ri) {
        const float *roi = rois + 5 * ri;
        int b = (int)roi[0];
        float sw = roi[1] * scale, sh = roi[2] * scale, ew = roi[3] * scale, eh = roi[4] * scale;
        float rw = ew - sw > 1.f ? ew - sw : 1.f, rh = eh - sh > 1.f ? eh - sh : 1.f;
        float bh = rh / (float)ph, bw = rw / (float)pw;
        int gh = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rh / ph);
        int gw = sampling_ratio > 0 ? sampling_ratio : (int)ceilf(rw / pw);
        float count = (float)(gh * gw);
        for (int64_t ci = 0; ci < c; ++ci) {
            const float *plane = feat ? feat + (b * c + ci) * h * w : NULL;
            float *dplane = dfeat ? dfeat + (b * c + ci) * h * w : NULL;
            for (int64_t oh = 0; oh < ph; ++oh)
                for (int64_t ow = 0; ow < pw; ++ow) {
                    int64_t oi = ((ri * c + ci) * ph + oh) * pw + ow;
                    float acc = 0.f;
                    for (int iy = 0; iy < gh; ++iy) {
                        float y = sh + oh * bh + (iy + .5f) * bh / (float)gh;
                        for (int ix = 0; ix < gw; ++ix) {
                            float x = sw + ow * bw + (ix + .5f) * bw / (float)gw;
                            int off[4]; float wt[4];
                            if (!orc_tap(y, x, (int)h, (int)w, off, wt)) continue;
                            if (out) acc += wt[0] * plane[off[0]] + wt[1] * plane[off[1]] + wt[2] * plane[off[2]] + wt[3] * plane[off[3]];
                            if (dplane) for (int k = 0; k < 4; ++k) dplane[off[k]] += dout[oi] * wt[k] / count;
                        }
                    }
                    if (out) out[oi] = acc / count;
                }
        }
    }
}
