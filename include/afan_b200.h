/* afan_b200.h -- C ABI of libafan_b200.so: hand-written sm_100a CUDA kernels for A-FAN's
 * adversarial-feature inner loop (feature-space PGD + clean/adversarial normalisation).
 *
 * The reference (VITA-Group/CV_A-FAN) is pure Python/PyTorch and has NO FFI for this path; the
 * "reference interface each entry point replaces" is therefore a span of un-fused ATen calls in the
 * reference's Python, cited per function as file:line relative to the reference root.  The binding a
 * maintainer would add to the reference (a ctypes stub) is shown in INTEGRATION.md.
 *
 * Conventions (all entry points)
 *   - plain pointers + sizes; no torch / C++ types.  All data pointers are DEVICE pointers to
 *     caller-owned, contiguous NCHW fp32 storage unless stated; nothing is allocated or retained.
 *   - `stream` is a cudaStream_t (0 = legacy default stream).  Every call is asynchronous: it only
 *     enqueues kernels on `stream`, never synchronises, and is CUDA-graph capturable.
 *   - returns AFAN_OK (0) or a negative AFAN_ERR_* code; never throws, exits or prints.
 *   - re-entrant: no global mutable state.  16-byte aligned pointers with sizes that are multiples
 *     of 4 elements take the 128-bit vector path; anything else takes a scalar path (same results).
 *   - `workspace`: device scratch owned by the caller, at least *_workspace_bytes() bytes, zeroed
 *     ONCE before first use (kernels leave it zeroed); one workspace per concurrently used stream.
 */
#ifndef AFAN_B200_H_
#define AFAN_B200_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* afan_stream_t; /* cudaStream_t */

enum {
    AFAN_OK = 0,
    AFAN_ERR_NULL = -1,        /* a required pointer is NULL */
    AFAN_ERR_SIZE = -2,        /* negative / inconsistent size */
    AFAN_ERR_WORKSPACE = -3,   /* workspace missing or too small */
    AFAN_ERR_LAUNCH = -4,      /* CUDA launch error (cudaPeekAtLastError) */
    AFAN_ERR_UNSUPPORTED = -5  /* shape outside what the kernels handle */
};

/* Library identification / diagnostics. */
const char* afan_version(void);
const char* afan_strerror(int code);
/* Fills SM count and compute capability of the current device; AFAN_ERR_UNSUPPORTED if not sm_100. */
int afan_device_info(int* sm_count, int* cc_major, int* cc_minor);
/* Measurement aid: launches an FFMA-only kernel (64 independent accumulators per thread, 2 CTAs x 256 threads per SM,
 * `iters` 8x8 outer-product updates) and returns its FLOP count in *flops_out (host pointer); the caller times it with
 * CUDA events to get the fp32 FFMA peak of the device at its current clocks.  out: >= 2 * sm_count * 256 floats. */
int afan_ffma_probe(float* out, int64_t out_elems, int64_t iters, double* flops_out, afan_stream_t stream);
/* HBM stream probes for the same purpose: mode 0 = read-only (sum), 1 = write-only, 2 = copy; n_elem floats, 16-byte aligned;
 * sink: >= 8 * SM-count floats of scratch.  A read-dominated kernel is bounded by the read figure, not by the copy bandwidth. */
int afan_hbm_probe(int mode, const float* src, float* dst, int64_t n_elem, float* sink, afan_stream_t stream);

/* ---- a2: random start ------------------------------------------------------------------------
 * Replaces Classification/attack_algo.py:41-44 (== Segmentation/attack_algo.py:43-45,
 * Detection/attack_algo.py:51-53):  x_adv = x + fl(fl(fl(2u) - 1) * eps).
 * _noise: `u` is the caller's torch.rand draw (bitwise parity with the reference's CPU generator).
 * _philox: u generated on the fly, Philox4x32-10, element i <- counter i/4 + offset, lane i%4,
 *          u = (bits >> 8) * 2^-24 (no noise tensor in HBM: 8 B/elem instead of 12).
 *          offset_device (nullable, device uint64): added to `offset` at run time, so a captured
 *          CUDA graph draws fresh noise on every replay. */
int afan_pgd_init_noise_f32(const float* x, const float* u, float* x_adv, int64_t n_elem, float eps,
                            afan_stream_t stream);
int afan_pgd_init_philox_f32(const float* x, float* x_adv, int64_t n_elem, float eps, uint64_t seed,
                             uint64_t offset, const uint64_t* offset_device, afan_stream_t stream);

/* ---- a3 + a4 (+ a11): one fused L-inf PGD update ------------------------------------------------
 * Replaces Classification/attack_algo.py:53-56 incl. linfball_proj/tensor_clamp (:9-19,35-36)
 * (== Segmentation/attack_algo.py:54-57, Detection/attack_algo.py:69-72), and optionally the
 * perturbation-norm logging of Classification/main_perturb.py:188-192:
 *     t = x_adv + fl(gamma) * sign(grad);  if clip: if (t < x-eps) t = x-eps; if (t > x+eps) t = x+eps
 * x_adv is updated IN PLACE.  x_clean may be NULL iff !clip && !delta_out && !norms_out.
 * grad == NULL: no ascent, projection (+ delta / norms) only == linfball_proj(x_clean, eps, x_adv).
 * delta_out (nullable): receives fl(x_adv_new - x_clean).
 * norms_out (nullable): [2][n_samples] -> per-sample ||delta||_2 then ||delta||_inf; needs workspace.
 * Bitwise identical to the reference op sequence (sign(NaN) = sign(-0) = +0; NaN compares false). */
int64_t afan_pgd_norms_workspace_bytes(int64_t n_samples);
int afan_pgd_linf_step_f32(const float* grad, const float* x_clean, float* x_adv, float* delta_out,
                           float* norms_out, void* workspace, int64_t workspace_bytes,
                           int64_t n_samples, int64_t per_sample, float gamma, float eps, int clip,
                           afan_stream_t stream);

/* tensor_clamp with free-form tensor bounds (Classification/attack_algo.py:9-19), in place on t:
 * t < min -> min; then t > max -> max; NaN in t is left alone.  (linfball_proj is the grad == NULL form above.) */
int afan_tensor_clamp_f32(float* t, const float* min, const float* max, int64_t n_elem, afan_stream_t stream);

/* bf16-storage twins (BASELINE config 3): tensors are bf16 (void* = __nv_bfloat16*), arithmetic is the same fp32
 * sequence on the widened values, stores round to nearest even; delta / norms are formed from the ROUNDED x_adv.
 * `u` stays fp32.  No reference exists for this dtype: the contract is bit-equality with oracle/afan_oracle.c's
 * orc_pgd_linf_step_bf16 and <= 1e-2 relative distance of delta from the fp32 path (north star). */
int afan_pgd_linf_step_bf16(const void* grad, const void* x_clean, void* x_adv, void* delta_out,
                            float* norms_out, void* workspace, int64_t workspace_bytes, int64_t n_samples,
                            int64_t per_sample, float gamma, float eps, int clip, afan_stream_t stream);
int afan_pgd_init_noise_bf16(const void* x, const float* u, void* x_adv, int64_t n_elem, float eps,
                             afan_stream_t stream);
int afan_pgd_init_philox_bf16(const void* x, void* x_adv, int64_t n_elem, float eps, uint64_t seed,
                              uint64_t offset, const uint64_t* offset_device, afan_stream_t stream);

/* ---- a5 / a5b: L2 mode ---------------------------------------------------------------------------
 * afan_sample_l2norm_f32: out_norm[s] = ||a_s - b_s||_2 (b nullable -> ||a_s||_2), deterministic
 *   two-level reduction (replaces `.view(N,-1).norm(p=2, dim=1)`, Classification/attack_algo.py:28).
 * afan_pgd_l2_step_f32:   x_adv += gamma / max(grad_norm[s], tiny) * grad   (L2-normalised ascent;
 *   north-star item with NO reference implementation -- semantics defined in oracle/afan_oracle.c).
 * afan_l2ball_proj_f32:   replaces l2ball_proj, Classification/attack_algo.py:21-33, given
 *   dist[s] = ||t_s - center_s||_2:  d = t - c; d /= dist; d *= min(dist, radius); t = c + d
 *   (t == center gives 0/0 = NaN exactly like the reference).  delta_out (nullable) receives
 *   fl(t_new - center). */
int afan_sample_l2norm_f32(const float* a, const float* b, float* out_norm, void* workspace,
                           int64_t workspace_bytes, int64_t n_samples, int64_t per_sample,
                           afan_stream_t stream);
int afan_pgd_l2_step_f32(const float* grad, const float* grad_norm, float* x_adv, int64_t n_samples,
                         int64_t per_sample, float gamma, float tiny, afan_stream_t stream);
int afan_l2ball_proj_f32(const float* center, const float* dist, float* t, float* delta_out,
                         int64_t n_samples, int64_t per_sample, float radius, afan_stream_t stream);

/* ---- a8: mix_feature --------------------------------------------------------------------------
 * Replaces Segmentation/attack_algo.py:121-130 == Detection/attack_algo.py:254-265: per (n,h,w),
 * channel-dim mean / sqrt(unbiased var + 1e-5) of clean swapped for those of adv.
 * clean/adv/out: [n][c][hw].  One sweep computes both statistics (Welford), a second (cache-hot)
 * sweep writes out: 12 B/elem of HBM traffic. */
int afan_mix_feature_f32(const float* clean, const float* adv, float* out, int64_t n, int64_t c,
                         int64_t hw, afan_stream_t stream);

/* ---- a9 (+ a8): SAT sample points with optional per-point mix_feature, fused ------------------------------
 * Replaces get_sample_points (Segmentation/attack_algo.py:108-118, Detection/attack_algo.py:236-245) followed by
 * `adv_list[i] = mix_feature(clean, adv_list[i])` (Segmentation/main_aug_final.py:206-210, Detection/
 * train_aug_final.py:117-126):  for j < m (m <= 4):  p_j = lerp(clean, adv, weights[j]);
 * outs[j] = mix_flags[j] ? mix_feature(clean, p_j) : p_j.   outs / weights / mix_flags are HOST arrays of length m
 * (outs holds device pointers).  clean and adv are read once per sweep for all points: 8 + 4m B/elem. */
int afan_sat_mix_f32(const float* clean, const float* adv, float* const* outs, const float* weights,
                     const int* mix_flags, int m, int64_t n, int64_t c, int64_t hw, afan_stream_t stream);

/* ---- a10: dual (grouped-statistics) train-mode BatchNorm2d ---------------------------------------
 * Replaces nn.BatchNorm2d (+ the F.relu / residual add that follow it) in the tail as seen by the
 * adversarial and the clean batch, Classification/resnet_s.py:54,56,70-76,89 driven by
 * main_perturb.py:195-196.  x: [groups*n][c][hw]; group g = samples [g*n, (g+1)*n) gets its OWN batch
 * statistics, all groups share weight/bias and the running averages (updated group after group,
 * `replay` times each: the head cache replays the reference's two head forwards, main_perturb.py:173,196).
 *     y = relu?( weight * (x - mean_g) * invstd_g + bias  (+ residual) )
 * Single-GPU entry points (stats + finalise in one kernel, apply in a second):                     */
int64_t afan_bn_workspace_bytes(int64_t groups, int64_t channels);
int afan_bn_fwd_f32(const float* x, const float* residual /*nullable*/, const float* weight,
                    const float* bias, float* running_mean /*nullable*/, float* running_var /*nullable*/,
                    float* y, float* save_mean /*[g][c]*/, float* save_invstd /*[g][c]*/,
                    void* workspace, int64_t workspace_bytes, int64_t groups, int64_t n, int64_t c,
                    int64_t hw, float eps, float momentum, int relu, int replay, afan_stream_t stream);
int afan_bn_bwd_f32(const float* dy, const float* x, const float* y /*required iff relu*/,
                    const float* weight, const float* save_mean, const float* save_invstd, float* dx,
                    float* dresidual /*nullable*/, float* dweight /*[c]*/, float* dbias /*[c]*/,
                    void* workspace, int64_t workspace_bytes, int64_t groups, int64_t n, int64_t c,
                    int64_t hw, int relu, afan_stream_t stream);
/* Multi-GPU (sync) split: the caller all-reduces `sums` (double [g][c][2]) over NCCL between the
 * two halves -- ONE message carries the clean and the adversarial statistics.
 *   fwd: sums = {sum x, sum x^2};  count = global elements per (group, channel) = n*hw*world
 *   bwd: sums = {sum dy, sum dy*xhat}; dweight/dbias are written from the LOCAL sums by _reduce. */
int afan_bn_fwd_stats_f32(const float* x, double* sums, void* workspace, int64_t workspace_bytes,
                          int64_t groups, int64_t n, int64_t c, int64_t hw, afan_stream_t stream);
int afan_bn_fwd_finalize_f32(const double* sums, double count, const float* weight, const float* bias,
                             float* running_mean, float* running_var, float* save_mean,
                             float* save_invstd, void* workspace, int64_t workspace_bytes,
                             int64_t groups, int64_t c, float eps, float momentum, int replay,
                             afan_stream_t stream);
int afan_bn_fwd_apply_f32(const float* x, const float* residual, float* y, const void* workspace,
                          int64_t workspace_bytes, int64_t groups, int64_t n, int64_t c, int64_t hw,
                          int relu, afan_stream_t stream);
int afan_bn_bwd_reduce_f32(const float* dy, const float* x, const float* y, const float* save_mean,
                           const float* save_invstd, double* sums, float* dweight, float* dbias,
                           void* workspace, int64_t workspace_bytes, int64_t groups, int64_t n,
                           int64_t c, int64_t hw, int relu, afan_stream_t stream);
int afan_bn_bwd_finalize_f32(const double* sums, double count, const float* weight,
                             const float* save_mean, const float* save_invstd, void* workspace,
                             int64_t workspace_bytes, int64_t groups, int64_t c, afan_stream_t stream);
int afan_bn_bwd_apply_f32(const float* dy, const float* x, const float* y, float* dx, float* dresidual,
                          const void* workspace, int64_t workspace_bytes, int64_t groups, int64_t n,
                          int64_t c, int64_t hw, int relu, afan_stream_t stream);
/* Fused multi-GPU form (one process per GPU): the kernel itself exchanges the per-(group, channel) sums with the
 * other GPUs over NVLink peer memory -- P2P stores of self-validating {lo, tag, hi, tag} words into every peer's
 * mailbox (no fence, one one-way NVLink latency), bounded spin on the in-band tags, fold in rank order
 * (bit-identical statistics on all ranks) -- and then normalises.  ONE
 * launch per BatchNorm direction and no NCCL call; replaces stats -> all-reduce -> finalize -> apply.
 *   peer_mailboxes: HOST array of `world` device pointers, [i] = rank i's mailbox mapped into this process
 *                   (afan_p2p_* below); each mailbox is afan_bn_mailbox_bytes(world, cmax) bytes, zeroed.
 *   state:          LOCAL device memory, 3 x uint64 {call sequence, ticket, error}, zeroed once.  error != 0
 *                   after a launch means a peer did not arrive within AFAN_P2P_TIMEOUT_S seconds (environment,
 *                   default 60: the tolerated inter-rank skew).  The kernel never hangs; on a timeout it also folds
 *                   NaN into the statistics so the failure cannot pass silently.
 * Every rank must issue the same sequence of calls.  AFAN_ERR_UNSUPPORTED (shape does not fit the
 * register-resident cluster kernel) is returned identically on all ranks: fall back to the split form. */
int64_t afan_bn_mailbox_bytes(int world, int64_t cmax);
int afan_bn_fwd_p2p_f32(const float* x, const float* residual, const float* weight, const float* bias,
                        float* running_mean, float* running_var, float* y, float* save_mean,
                        float* save_invstd, int64_t groups, int64_t n, int64_t c, int64_t hw, float eps,
                        float momentum, int relu, int replay, int world, int rank,
                        void* const* peer_mailboxes, int64_t cmax, void* state, afan_stream_t stream);
int afan_bn_bwd_p2p_f32(const float* dy, const float* x, const float* y, const float* weight,
                        const float* save_mean, const float* save_invstd, float* dx, float* dresidual,
                        float* dweight, float* dbias, int64_t groups, int64_t n, int64_t c, int64_t hw,
                        int relu, int world, int rank, void* const* peer_mailboxes, int64_t cmax,
                        void* state, afan_stream_t stream);
/* Peer-mapped device memory (cudaMalloc + cudaIpc): alloc zeroes the buffer; handles are 64 opaque bytes that the
 * host exchanges between the processes of one node; open maps a peer's buffer (NVLink P2P enabled lazily). */
int afan_p2p_alloc(void** ptr, int64_t bytes);
int afan_p2p_free(void* ptr);
int afan_p2p_get_handle(void* ptr, void* handle_out_64);
int afan_p2p_open_handle(const void* handle_64, void** ptr_out);
int afan_p2p_close_handle(void* ptr);

/* Inference-mode affine (+residual, +relu) with a caller-supplied per-channel scale/shift table
 * (float [c][2]); used for model.eval() (main_perturb.py:232-246). */
int afan_bn_affine_f32(const float* x, const float* residual, const float* scale_shift, float* y,
                       int64_t n, int64_t c, int64_t hw, int relu, afan_stream_t stream);
/* Its backward, for BatchNorm layers that stay FROZEN during training (Detection/model.py:27-35,47-48: eval mode,
 * requires_grad False): dx = scale[c] * dy_eff, dresidual (nullable) = dy_eff, dy_eff = dy where y > 0 when relu. */
int afan_bn_affine_bwd_f32(const float* dy, const float* y, const float* scale_shift, float* dx, float* dresidual,
                           int64_t n, int64_t c, int64_t hw, int relu, afan_stream_t stream);

/* ---- small launches of the Classification tail (round 2: the library ran them as 3-6 launches / a 91 us SIMT sgemm) ----
 * Option-A shortcut of a stage transition, Classification/resnet_s.py:60-63 `F.pad(x[:, :, ::2, ::2], (0,0,0,0,pad,pad))`:
 * x [n][c][h][w] -> y [n][c+2*pad][ceil(h/2)][ceil(w/2)] and its backward dy -> dx (every output element written once). */
int afan_shortcut_a_fwd_f32(const float* x, float* y, int64_t n, int64_t c, int64_t h, int64_t w, int64_t pad,
                            afan_stream_t stream);
int afan_shortcut_a_bwd_f32(const float* dy, float* dx, int64_t n, int64_t c, int64_t h, int64_t w, int64_t pad,
                            afan_stream_t stream);
/* Weight / bias gradient of the classifier `nn.Linear` (resnet_s.py:93-95): dweight [out][in] = dy^T x, dbias [out]
 * (nullable) = column sums of dy [batch][out]; x [batch][in].  Batch walked in a fixed order (deterministic). */
int afan_linear_wgrad_f32(const float* dy, const float* x, float* dweight, float* dbias, int64_t batch,
                          int64_t out_features, int64_t in_features, afan_stream_t stream);

/* ---- f4: greedy NMS, fully on the device ---------------------------------------------------------------------
 * Replaces Detection/support/src/cuda/nms.cu:23-131 (+ its D2H mask copy and serial CPU sweep, :99-123).
 * boxes_sorted: [n][4] (x1,y1,x2,y2) ALREADY sorted by score descending (16-byte aligned); order[i] = original index
 * of sorted position i.  Legacy "+1" areas; a box is dropped when IoU > threshold with an earlier kept box.
 * keep_flags [n] (uint8, indexed by ORIGINAL box index) and count_out (device int32, nullable) are written on the
 * stream; nothing is copied to the host. */
int64_t afan_nms_workspace_bytes(int64_t n);
int afan_nms_f32(const float* boxes_sorted, const int64_t* order, float threshold, uint8_t* keep_flags,
                 int32_t* count_out, void* workspace, int64_t workspace_bytes, int64_t n, afan_stream_t stream);

/* Batched form for the proposal layer (Detection/rpn/region_proposal_network.py:244-256: per-image NMS over the ranked
 * boxes followed by `[:post_nms_top_n]` and zero padding): one launch pair for the whole batch, one CTA per image, the
 * sweep stops once max_keep boxes of an image are kept.  boxes_sorted [images][n][4] ranked by descending score;
 * kept_boxes [images][max_keep][4] (nullable, 16-byte aligned) = the surviving boxes in rank order, zero-padded;
 * keep_flags [images][n] (nullable) indexed by RANK; counts [images] = min(survivors, max_keep). */
int64_t afan_nms_batched_workspace_bytes(int64_t images, int64_t n);
int afan_nms_batched_f32(const float* boxes_sorted, float threshold, int64_t max_keep, float* kept_boxes,
                         uint8_t* keep_flags, int32_t* counts, void* workspace, int64_t workspace_bytes,
                         int64_t images, int64_t n, afan_stream_t stream);

/* ---- f4: ROIAlign ------------------------------------------------------------------------------------------------
 * Replaces Detection/support/src/cuda/ROIAlign_cuda.cu:64-122 (forward) and :177-254 (backward), bound by
 * Detection/support/layer/roi_align.py.  feat [n][c][h][w]; rois [r][5] = (batch index, x1, y1, x2, y2) in image
 * coordinates; out / dout [r][c][ph][pw].  Legacy (non-"aligned") coordinates; sampling_ratio <= 0 -> adaptive grid
 * ceil(roi_size / pooled_size).  _bwd zero-fills dfeat itself, then scatters with atomic adds. */
int afan_roi_align_fwd_f32(const float* feat, const float* rois, float* out, int64_t n, int64_t c, int64_t h,
                           int64_t w, int64_t r, int64_t ph, int64_t pw, float spatial_scale,
                           int sampling_ratio, afan_stream_t stream);
int afan_roi_align_bwd_f32(const float* dout, const float* rois, float* dfeat, int64_t n, int64_t c, int64_t h,
                           int64_t w, int64_t r, int64_t ph, int64_t pw, float spatial_scale,
                           int sampling_ratio, afan_stream_t stream);

/* ---- a1/a6 tail: 3x3 / stride 1 / pad 1 convolutions re-executed by every PGD step ---------------
 * Replace nn.Conv2d `conv1` / `conv2` of BasicBlock (Classification/resnet_s.py:53,55) as called by the
 * ascent's tail passes (attack_algo.py:49-52) and the final adversarial / clean passes
 * (main_perturb.py:195-200), for C_in == C_out == c in {16, 32, 64} on square hw in {8, 16, 32} maps,
 * fp32 NCHW, strict fp32 FFMA accumulation in a fixed order (deterministic).  Other shapes return
 * AFAN_ERR_UNSUPPORTED (the caller keeps the library convolution for them).
 *
 * afan_conv3x3_pack_f32: descs_device = device array of n_layers records {const float* w; float* wf;
 *     float* wd; int64 c} (c < 0: stride-2 transition layer with cin = -c, see afan_conv3x3s2_f32); repacks W[co][ci][3][3] of every layer in ONE launch into the forward packing
 *     wf[ci][tap][co] and the input-gradient packing wd[co][8-tap][ci] (c*9*c floats each).
 * afan_conv3x3_f32: y = conv(x, W) when given wf; dx = conv_transpose(dy, W) when given dy and wd.
 *     addend (nullable, shape of y): y = conv(...) + addend in the epilogue -- the gradient of the identity shortcut
 *     (resnet_s.py:75 `out += self.shortcut(x)`) joins the input gradient without an accumulation launch.
 *     variant 0 = tuned default; other values select alternative tilings (benchmarking only).
 * afan_conv3x3_wgrad_f32: dW[co][ci][3][3] = sum_{n,h,w} dy * shifted x.  Two launches: per-CTA partials
 *     into `workspace` (>= afan_conv3x3_wgrad_workspace_bytes(c), need not be zeroed), then a fixed-order fold
 *     that stores (accumulate = 0) or adds into dw (accumulate = 1: writes straight into a gradient arena
 *     instead of a temporary + autograd's accumulation launch). */
int afan_conv3x3_pack_f32(const void* descs_device, int64_t n_layers, int64_t c_max, afan_stream_t stream);
int afan_conv3x3_f32(const float* x, const float* w_packed, float* y, const float* addend, int64_t n, int64_t c,
                     int64_t hw, int variant, afan_stream_t stream);
/* Stride-2 stage transitions (first block of stages 2 and 3, resnet_s.py:98): W[2*cin][cin][3][3], x [n][cin][2*ho][2*ho]
 * -> y [n][2*cin][ho][ho]; dgrad != 0 computes dx from dy instead.  (cin, ho) in {(16, 16), (32, 8)}; the packings come
 * from afan_conv3x3_pack_f32 with a descriptor whose c field is -cin (18*cin*cin floats each). */
int afan_conv3x3s2_f32(const float* in, const float* w_packed, float* out, int64_t n, int64_t cin, int64_t ho, int dgrad,
                       afan_stream_t stream);
/* Weight gradient of the same transition: dw [2*cin][cin][3][3] stored (accumulate = 0) or added to (accumulate = 1);
 * workspace >= afan_conv3x3_wgrad_workspace_bytes(2 * cin); two launches, fixed-order fold (deterministic). */
int afan_conv3x3s2_wgrad_f32(const float* x, const float* dy, float* dw, void* workspace, int64_t workspace_bytes,
                             int64_t n, int64_t cin, int64_t ho, int accumulate, afan_stream_t stream);
/* Tensor-core twins of the two calls above (mma.sync m16n8k8 TF32, fp32 accumulate).  passes = 3: "3xTF32" split
 * (x = hi + lo; a_lo*b_hi + a_hi*b_lo + a_hi*b_hi) -- fp32-level accuracy on the tensor pipe; passes = 1: plain TF32.
 * The packings are [k/8][tap][out][k%8][hi, lo] (passes = 3: 2*c*9*c floats per direction) or [k/8][tap][out][k%8]
 * (passes = 1: c*9*c floats), hi = W rounded to TF32, lo = TF32 rounding of (W - hi). */
int afan_conv3x3_pack_tc_f32(const void* descs_device, int64_t n_layers, int64_t c_max, int passes, afan_stream_t stream);
int afan_conv3x3_tc_f32(const float* x, const float* w_packed, float* y, const float* addend, int64_t n, int64_t c,
                        int64_t hw, int passes, int variant, afan_stream_t stream);
/* tcgen05 implicit GEMM of the same convolution (Blackwell 5th-generation tensor cores, accumulators in TMEM, operands
 * staged by cp.async.bulk): kind::tf32 with the 3xTF32 hi/lo split, i.e. fp32-grade accuracy (1-3e-5 of fp64 like the
 * FFMA kernel's fixed-order fp32 sum), deterministic.  Replaces the same nn.Conv2d spans (resnet_s.py:53,55) for
 * (c, hw) in {(32, 16), (64, 8)} -- the tail shapes every PGD step re-executes; (64, 8) needs an even n.
 * afan_conv3x3_pack_umma_f32 takes the descriptor table of afan_conv3x3_pack_f32 and repacks the layers with c in
 * {32, 64} (2*c*9*c floats per direction: [k/8][n/32][tap][hi, lo][n/8][k slice][n%8][k%4]); other layers are left alone. */
int afan_conv3x3_umma_supported(int64_t n, int64_t c, int64_t hw);
int afan_conv3x3_pack_umma_f32(const void* descs_device, int64_t n_layers, int64_t c_max, afan_stream_t stream);
int afan_conv3x3_umma_f32(const float* x, const float* w_packed, float* y, const float* addend, int64_t n, int64_t c,
                          int64_t hw, afan_stream_t stream);
/* BatchNorm folded into the tcgen05 convolution (resnet_s.py:70-72 `relu(bn1(conv1(x)))` -> `conv2`, train mode):
 *   out_partials != NULL (producer, conv1): the kernel also writes per-CTA per-channel {sum, sum of squares} of its
 *     OUTPUT into out_partials (>= afan_conv3x3_umma_bn_workspace_bytes(n, c) bytes; need not be zeroed).
 *   in_partials != NULL (consumer, conv2): x is the producer's raw output and in_partials its statistics.  Every CTA
 *     folds them in a fixed order, finalises the train-mode statistics of `groups` statistic groups along the batch
 *     exactly like afan_bn_fwd_f32 and applies BatchNorm + ReLU while the input is staged (v = max(0, x*scale + shift)):
 *     the normalised activation is never written to memory.  One CTA also writes save_mean / save_invstd [groups][c],
 *     the {scale, shift} table [groups][c][2] (for afan_bn_bwd_xmask_f32) and advances the running statistics in
 *     group order, `replay` times.
 * afan_bn_bwd_xmask_f32 is the matching BatchNorm + ReLU backward: the ReLU mask is recomputed from x and the table
 * (there is no stored forward output).  Same shape support as afan_conv3x3_umma_f32; groups in {1, 2}. */
int64_t afan_conv3x3_umma_bn_workspace_bytes(int64_t n, int64_t c);
int afan_conv3x3_umma_bn_f32(const float* x, const float* w_packed, float* y, const void* in_partials,
                             const float* bn_weight, const float* bn_bias, float* running_mean, float* running_var,
                             float* save_mean, float* save_invstd, float* table_out, void* out_partials,
                             int64_t groups, int64_t n, int64_t c, int64_t hw, float eps, float momentum, int replay,
                             afan_stream_t stream);
int afan_bn_bwd_xmask_f32(const float* dy, const float* x, const float* mask_table, const float* weight,
                          const float* save_mean, const float* save_invstd, float* dx, float* dweight, float* dbias,
                          int64_t groups, int64_t n, int64_t c, int64_t hw, afan_stream_t stream);
/* Multi-GPU forms (one process per GPU, mailboxes / state as for afan_bn_fwd_p2p_f32): the consumer convolution folds its
 * LOCAL producer statistics, ONE of its CTAs publishes them to every peer's mailbox and every CTA collects all ranks' sums
 * (rank order: bit-identical statistics on all GPUs) while the operand loads are in flight -- global-batch BatchNorm without
 * a BatchNorm launch.  n <= SM count (the grid must be co-resident).  Every rank must issue the same sequence of
 * exchanging calls (these and the afan_bn_*_p2p_f32 ones share the mailbox sequence). */
int afan_conv3x3_umma_bn_p2p_f32(const float* x, const float* w_packed, float* y, const void* in_partials,
                                 const float* bn_weight, const float* bn_bias, float* running_mean, float* running_var,
                                 float* save_mean, float* save_invstd, float* table_out, int64_t groups, int64_t n,
                                 int64_t c, int64_t hw, float eps, float momentum, int replay, int world, int rank,
                                 void* const* peer_mailboxes, int64_t cmax, void* state, afan_stream_t stream);
int afan_bn_bwd_xmask_p2p_f32(const float* dy, const float* x, const float* mask_table, const float* weight,
                              const float* save_mean, const float* save_invstd, float* dx, float* dweight, float* dbias,
                              int64_t groups, int64_t n, int64_t c, int64_t hw, int world, int rank,
                              void* const* peer_mailboxes, int64_t cmax, void* state, afan_stream_t stream);
/* tcgen05 twin of afan_conv3x3_wgrad_f32 for (c, hw) in {(32, 16), (64, 8)}: the weight gradient as a GEMM whose
 * reduction dimension is the pixels (kind::tf32, 3xTF32 split, TMEM accumulators; the three kx shifts are materialised
 * while staging and form the M dimension, ky shifts are descriptor offsets), persistent CTAs, one partial dW per CTA and
 * a fixed-order fold (deterministic).  Same workspace (afan_conv3x3_wgrad_workspace_bytes(c)) and accumulate semantics. */
int afan_conv3x3_wgrad_umma_supported(int64_t n, int64_t c, int64_t hw);
int afan_conv3x3_wgrad_umma_f32(const float* x, const float* dy, float* dw, void* workspace, int64_t workspace_bytes,
                                int64_t n, int64_t c, int64_t hw, int accumulate, afan_stream_t stream);
int64_t afan_conv3x3_wgrad_workspace_bytes(int64_t c);
int afan_conv3x3_wgrad_f32(const float* x, const float* dy, float* dw, void* workspace, int64_t workspace_bytes,
                           int64_t n, int64_t c, int64_t hw, int accumulate, afan_stream_t stream);

/* ---- a7 tail: fused SGD(momentum, weight decay) over a flat parameter arena ---------------------
 * Replaces optimizer.step() of torch.optim.SGD, main_perturb.py:72-74,201:
 *     g = grad*grad_scale + wd*p;  buf = momentum*buf + g;  p -= lr*buf      (buf starts at 0)
 * lr is read from DEVICE memory so a captured CUDA graph follows the LR schedule (warm-up,
 * main_perturb.py:167-168,288-293). */
int afan_sgd_momentum_f32(float* param, const float* grad, float* momentum_buf, int64_t n_elem,
                          const float* lr_device, float momentum, float weight_decay, float grad_scale,
                          afan_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* AFAN_B200_H_ */
