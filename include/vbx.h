/* vbx.h - C ABI of libvbx_b200.so: the sm_100a kernels behind the EBEN
 * bandwidth-extension training step of jhauret/vibravox.
 *
 * The reference has NO native layer: its hot path is nn.Module.forward
 * calls that dispatch into ATen / cuDNN / cuFFT.  Each entry point below therefore
 * cites the reference *Python call site* whose ATen dispatch it replaces.
 *
 * Conventions
 *   - every tensor is fp32, contiguous, (B, C, T) time-contiguous, device memory
 *     owned by the caller; the library never allocates device memory, never
 *     synchronises, never changes the current device;
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream);
 *   - return 0 on success, a negative vbx_status on a bad argument, a positive
 *     cudaError_t if the launch failed; vbx_last_error() gives the text;
 *   - all entry points are re-entrant; state is limited to the launch counter.
 */
#ifndef VBX_H
#define VBX_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define VBX_ABI_VERSION 7
#if defined(__GNUC__)
#define VBX_API __attribute__((visibility("default")))
#else
#define VBX_API
#endif

enum vbx_status { VBX_OK = 0, VBX_BAD_SHAPE = -1, VBX_BAD_POINTER = -2, VBX_UNSUPPORTED = -3 };

/* Geometry of one Conv1d layer pass.  y[b,co,t] = sum W[co,ci,k]*x[b,ci,map(t*stride+k*dil-pad)],
 * where map() mirrors positions up to `refl` samples outside [0,Tin) (PyTorch 'reflect',
 * no edge repeat) and yields zero beyond that.  Tout must equal
 * (Tin + 2*pad - dil*(K-1) - 1)/stride + 1.  refl <= pad, refl <= Tin-1. */
typedef struct {
  int32_t B, Cin, Cout, Tin, Tout, K, stride, dil, pad, refl, groups;
} vbx_conv_desc;

/* Fused output stage shared by the conv entry points:
 *   v = acc + bias[ch];  mask[idx] = v > 0;  v = v > 0 ? v : slope*v;  v += res[idx];
 *   if gate:  y = gate[idx];
 *             if fm_other: v += fm_coef[0]*sign(y - fm_other[idx]) - fm_coef[1]*sign(y);
 *             v *= (y > 0 ? 1 : gate_slope);
 *   out[idx] = v + beta*out[idx]            (each step skipped when its pointer is NULL / slope==1 / beta==0)
 * The gate stage is the backward of the PRODUCER of this kernel's output, folded into an input-gradient epilogue:
 * `gate` is the post-LeakyReLU activation y the gradient belongs to (melgan_discriminator.py:89-156 /
 * eben_discriminator.py:66-157 keep every stage output for the feature-matching loss), gate_slope its negative slope,
 * and fm_other / fm_coef add the feature-matching gradient of that activation (feature_loss.py:37-50; fm_coef from
 * vbx_fm_coef) before the gate - replacing an aten::add, an aten::leaky_relu_backward and the L1-pair backward pass. */
typedef struct {
  const float* bias;
  const float* res;
  uint8_t* mask;
  float slope;
  float beta;
  const float* gate;
  const float* fm_other;
  const float* fm_coef;
  float gate_slope;
  float* gate_dbias;   /* with gate, input-gradient entries only: gate_dbias[c] += sum over (b, t) of out[b, c, t] - the
                        * bias gradient of the stage that produced `gate` (its dy IS this output), reduced in the same
                        * in-place pass where the gate runs as one, else in one read-only pass */
} vbx_epilogue;

VBX_API int vbx_abi_version(void);
VBX_API const char* vbx_last_error(void);
/* number of kernels this library has launched since load (for bench.py's gpu_launches) */
VBX_API uint64_t vbx_launch_count(void);
/* opt-in tensor-core path for dense layers (0 = fp32 FMA everywhere) */
VBX_API int vbx_set_tensor_core_mode(int mode);
/* Deterministic reductions (returns the previous setting).  Off (default): weight gradients split their (b,t)
 * reduction over several CTAs and bias gradients / loss sums over several blocks, combined with fp32 / fp64
 * atomics - fastest, but the summation order varies from launch to launch (~1e-7 relative).  On: every such
 * reduction is walked by ONE CTA per output tile in a fixed order, so repeated runs - eager launches or a CUDA-graph
 * replay - are bit-identical.  Meant for tests and debugging (large layers lose their split-K parallelism). */
VBX_API int vbx_set_deterministic(int on);
/* out[i] = scale * *p_i for up to 8 device scalars (NULL entries beyond n are ignored): packs the step's logged
 * losses into the tail of a gradient bucket so that ONE all-reduce carries them (the reference logs with
 * sync_dist=True, eben.py:103-124), and unpacks the rank mean afterwards. */
VBX_API int vbx_gather_scalars(const float* p0, const float* p1, const float* p2, const float* p3, const float* p4,
                       const float* p5, const float* p6, const float* p7, int32_t n, float scale, float* out,
                       void* stream);

/* ---- Conv1d family ------------------------------------------------------------------
 * replaces aten::conv1d / aten::reflection_pad1d / aten::leaky_relu / aten::add issued by
 * eben_generator.py:112-166,241-249,272-280,295-316, eben_discriminator.py:66-157,
 * melgan_discriminator.py:89-156, and (STFT-as-conv) auraloss STFTLoss.stft. */
VBX_API int vbx_conv1d_fwd(const vbx_conv_desc* d, const float* x, const float* w, const vbx_epilogue* e,
                   float* y, void* stream);
/* dx = epi(conv_transpose(dy)).  wt is the group-transposed weight Wt[g][ci][co_g][k]
 * (vbx_weight_norm_fwd / vbx_transpose_weight produce it).  Replaces aten::convolution_backward
 * (input gradient) AND the forward of nn.ConvTranspose1d (eben_generator.py:241-249). */
VBX_API int vbx_conv1d_dgrad(const vbx_conv_desc* d, const float* dy, const float* wt, const vbx_epilogue* e,
                     float* dx, void* stream);
/* dw[co][ci][k] += sum_{b,t} dy*x  (accumulates: the caller zeroes its flat gradient bucket once).
 * Replaces aten::convolution_backward (weight gradient). */
VBX_API int vbx_conv1d_wgrad(const vbx_conv_desc* d, const float* x, const float* dy, float* dw, void* stream);
/* col2im form of dgrad: dx += scatter(Wk^T dy), wk = Wk[g][(ci,k)][co_g]; dx must be pre-zeroed
 * (or hold the value to accumulate onto).  Used for the STFT-as-conv backward. */
VBX_API int vbx_conv1d_dgrad_scatter(const vbx_conv_desc* d, const float* dy, const float* wk, float* dx,
                             void* stream);
/* ---- tensor-core (tcgen05 / TMEM) variants of the same contractions, bf16x3 split operands, fp32
 * accumulate.  Weights are pre-packed once per weight update into K-major bf16 hi/lo tiles
 * (vbx_tc_pack_bytes gives the buffer size, -1 on a bad descriptor); activations stay fp32 (B,C,T).
 * mode 0 = forward pack (columns = output channels), 1 = dgrad pack (columns = input channels, one
 * tile set per stride phase).  w is the plain W[co][ci_g][k].
 * nsplit = bf16 components per operand: 2 -> hi+lo, 3 MMAs per product ("bf16x3", 16 mantissa bits);
 * 3 -> hi+mid+lo, 6 MMAs ("bf16x6", 24 bits = fp32-grade; used where a log / division amplifies rounding). */
VBX_API int64_t vbx_tc_pack_bytes(const vbx_conv_desc* d, int32_t mode, int32_t nsplit);
VBX_API int vbx_tc_pack(const vbx_conv_desc* d, int32_t mode, int32_t nsplit, const float* w, void* packed,
                void* stream);
VBX_API int vbx_tc_conv1d_fwd(const vbx_conv_desc* d, const float* x, const void* packed, const vbx_epilogue* e,
                      float* y, int32_t nsplit, void* stream);
VBX_API int vbx_tc_conv1d_dgrad(const vbx_conv_desc* d, const float* dy, const void* packed, const vbx_epilogue* e,
                        float* dx, int32_t nsplit, void* stream);
/* dw += sum dy*x on tensor cores (both operands gathered from the fp32 activations; nothing packed) */
VBX_API int vbx_tc_conv1d_wgrad(const vbx_conv_desc* d, const float* x, const float* dy, float* dw, void* stream);
/* ---- fused ResidualUnit forward (eben_generator.py:287-316: x + LeakyReLU(pointwise_conv(dilated_conv(x)))) ----
 * ONE persistent kernel per unit: x (B,C,T) is read once through a 3-D TMA tensor map, both convs run on tcgen05
 * (bf16x3, the dilated conv's accumulator is re-split in place as the pointwise conv's operand), LeakyReLU + the
 * residual (taken from the fp32 tile in shared memory) ride in the epilogue, `out` is written once.
 * Supported: C in {16, 32, 48, 64}, T % 4 == 0, 1 <= dil <= 16, T > dil, k = 3 / k = 1, no bias (the generator's C = 32
 * and C = 64 stages; vbx_ru_supported says so; other shapes use the two-launch form above).
 * w_dil (C,C,3) and w_pw (C,C,1) are the effective (weight-normed) weights; vbx_ru_pack turns them into the
 * resident shared-memory image (vbx_ru_pack_bytes bytes).  h / mask (optional, may be NULL) are what the existing
 * backward kernels consume: h = dilated_conv(x) fp32, mask = pre-activation > 0 (1 byte). */
VBX_API int vbx_ru_supported(int32_t B, int32_t C, int32_t T, int32_t dil);
VBX_API int64_t vbx_ru_pack_bytes(int32_t C);
VBX_API int vbx_ru_pack(int32_t C, const float* w_dil, const float* w_pw, void* packed, void* stream);
/* bring-up / profiling hook: when buf != NULL (device memory, >= 16 * tiles-per-CTA int64) CTA 0 of the next
 * vbx_ru_fwd launches records clock64() per (tile, pipeline event); NULL switches it off (tools/ru_timeline.py) */
VBX_API int vbx_ru_set_profile_buffer(void* buf);
VBX_API int vbx_ru_fwd(int32_t B, int32_t C, int32_t T, int32_t dil, float slope, const float* x, const void* packed,
               float* out, float* h, uint8_t* mask, void* stream);
/* Weight gradient of the residual-unit convs (stride 1, C -> C with C in {32, 64}; K = 3 dilated with reflect halo
 * `dil`, or K = 1) on the same TMA skeleton: dw[co][ci][k] = beta*dw + sum_{b,t} dy[b,co,t] * x[b,ci,mirror(t+(k-1)*dil)].
 * Persistent CTAs keep the accumulators in TMEM, write per-CTA partials to `workspace` (vbx_ru_wgrad_workspace bytes,
 * -1 = unsupported shape) and a second launch sums them in a fixed order: deterministic, no atomics.
 * Replaces aten::convolution_backward (weight gradient) of eben_generator.py:295-312. */
VBX_API int64_t vbx_ru_wgrad_workspace(int32_t B, int32_t C, int32_t T, int32_t dil, int32_t K);
VBX_API int vbx_ru_wgrad(int32_t B, int32_t C, int32_t T, int32_t dil, int32_t K, const float* x, const float* dy,
                 float* dw, float beta, void* workspace, void* stream);
/* W[co][ci_g][k] -> Wt[g][ci_g][co_g][k]  (layout for vbx_conv1d_dgrad) */
VBX_API int vbx_transpose_weight(const float* w, float* wt, int32_t Cout, int32_t Cin_g, int32_t K,
                         int32_t groups, void* stream);

/* ---- weight norm (torch_modules/utils.py:4-9; aten::_weight_norm_interface, dim=0) -----
 * w[r,:] = g[r] * v[r,:] / ||v[r,:]||  for R rows of length `row` (= C1*K).
 * inv_norm[r] = 1/||v[r,:]|| is saved for the backward.  When wt != NULL the group-transposed
 * copy is written as well (Cout=R, Cin_g, K, groups describe the conv the weight belongs to). */
VBX_API int vbx_weight_norm_fwd(const float* g, const float* v, float* w, float* wt, float* inv_norm,
                        int32_t R, int32_t Cin_g, int32_t K, int32_t groups, void* stream);
/* dg[r] (+)= <dw,v>/||v|| ;  dv (+)= g/||v|| * dw - g*<dw,v>/||v||^3 * v   (accumulate when beta=1) */
VBX_API int vbx_weight_norm_bwd(const float* g, const float* v, const float* inv_norm, const float* dw,
                        float* dg, float* dv, int32_t R, int32_t row, float beta, void* stream);

/* ---- PQMF (dsp/pqmf.py:194-213), one polyphase kernel each way ---------------------------
 * analysis: y[b,c,t] = sum_k w[c,k] * x[b, m*t + k - (n-1)], c < bands, zero halo.
 *   (F.conv1d(x, W[:bands], stride=m, padding=n-1), pqmf.py:196-202)
 * synthesis: per band z[b,c,u] = sum_{k: (u+n-1-k) % m == 0} w[c,k] * x[b,c,(u+n-1-k)/m];
 *   sum_bands=1 writes y[b,0,u] = sum_c z (the generator's `.sum(1)`, eben_generator.py:209-211),
 *   sum_bands=0 writes the (B,m,L) tensor F.conv_transpose1d returns (pqmf.py:204-213).
 * T = band-rate length, L = full-rate length (the reference has T = (L+n-2)/m + 1 and
 * L = m*T - n; any T, L are accepted and out-of-range taps read zero, which also makes each
 * kernel the other's backward: x_per_band=1 lets analysis read a (B,bands,L) input, `bands`
 * in synthesis is the number of input channels (<= m) that are filtered/summed). */
VBX_API int vbx_pqmf_analysis(const float* x, const float* w, float* y, int32_t B, int32_t L, int32_t T,
                      int32_t m, int32_t n, int32_t bands, int32_t x_per_band, void* stream);
VBX_API int vbx_pqmf_synthesis(const float* x, const float* w, float* y, int32_t B, int32_t T, int32_t L,
                       int32_t m, int32_t n, int32_t bands, int32_t sum_bands, void* stream);

/* ---- element-wise -------------------------------------------------------------------- */
/* y = x > 0 ? x : slope*x                       (nn.LeakyReLU, eben_generator.py:110,187-189) */
VBX_API int vbx_leaky_relu_fwd(const float* x, float* y, int64_t n, float slope, void* stream);
/* dx = dy * (ref > 0 ? 1 : slope) (+ beta*dx); ref = input or output of the activation.
 * If dbias != NULL also dbias[c] += sum_{b,t} dx[b,c,t]  (C, T give the layout; bias gradient
 * of the discriminator convs).  mask != NULL replaces the sign test on ref. */
VBX_API int vbx_leaky_relu_bwd(const float* dy, const float* ref, const uint8_t* mask, float* dx, float* dbias,
                       int32_t B, int32_t C, int32_t T, float slope, float beta, void* stream);
/* y[b,c,t] = tanh(x[b,c,t] + (c < p ? first[b,c,t] : 0))      (eben_generator.py:203-208) */
VBX_API int vbx_tanh_recompose_fwd(const float* x, const float* first, float* y, int32_t B, int32_t m,
                           int32_t p, int32_t T, void* stream);
/* dx = dy * (1 - y^2) */
VBX_API int vbx_tanh_bwd(const float* dy, const float* y, float* dx, int64_t n, void* stream);
/* y = a + b */
VBX_API int vbx_add(const float* a, const float* b, float* y, int64_t n, void* stream);
/* y = alpha[0] * x (+ beta*y), alpha a device scalar (the loss-balancing lambdas never visit the host) */
VBX_API int vbx_axpby_dev(const float* x, float* y, int64_t n, const float* alpha, float beta, void* stream);
/* y = alpha * x (+ beta*y) */
VBX_API int vbx_axpby(const float* x, float* y, int64_t n, float alpha, float beta, void* stream);

/* ---- losses -------------------------------------------------------------------------- */
/* Feature matching (losses/feature_loss.py:37-50).  For one embedding pair accumulate
 * sums[0] += sum|a-b|, sums[1] += sum|a| (double). */
VBX_API int vbx_l1_pair_sums(const float* a, const float* b, int64_t n, double* sums, void* stream);
/* loss = scale * sum_i sums[2i]/sums[2i+1] over npairs   (scale = 1/(scales*layers)) */
VBX_API int vbx_fm_finalize(const double* sums, int32_t npairs, float scale, float* loss, void* stream);
/* da = go*scale*( sign(a-b)/S_a - S_ab/S_a^2 * sign(a) ), db = -go*scale*sign(a-b)/S_a
 * (either may be NULL); go is a device scalar. */
VBX_API int vbx_l1_pair_bwd(const float* a, const float* b, int64_t n, const double* sums, const float* go,
                    float scale, float* da, float* db, void* stream);
/* ---- ResidualUnit backward through the composed conv ----
 * z = w2 * (w1 (*) x) is ONE k-tap conv with wf[co][ci][k] = sum_m w2[co][m] w1[m][ci][k] (vbx_unit_combine; w1 is
 * (C, C, K), w2 is (C, C)), so the unit's backward (eben_generator.py:314-316 under autograd) needs a single input
 * gradient and a single weight gradient - dx = dgrad(dz, wf) + g and dwf = wgrad(x, dz) - from which
 * vbx_unit_split_grads recovers  dw1 = w2^T dwf  and  dw2[co][m] = <dwf[co], w1[m]>  (either may be NULL; beta scales
 * what the outputs hold).  The intermediate activation h of the forward pass is not needed by the backward pass. */
/* dx += the mirror terms of the input gradient of a k = 3, stride-1, groups-1 conv (C -> C) whose reflect halo equals its
 * dilation: with them, the ZERO-halo input gradient (the fast slab-form kernels) becomes the reflect-halo one
 * (eben_generator.py:295-312 padding_mode="reflect").  w is W[co][ci][3]; touches 2*dil positions per row. */
VBX_API int vbx_reflect_fold_k3(const float* dy, const float* w, float* dx, int32_t B, int32_t C, int32_t T, int32_t dil,
                        void* stream);
VBX_API int vbx_unit_combine(const float* w1, const float* w2, int32_t C, int32_t K, float* wf, void* stream);
VBX_API int vbx_unit_split_grads(const float* dwf, const float* w1, const float* w2, int32_t C, int32_t K, float* dw1,
                         float* dw2, float beta, void* stream);
/* coef[2i] = go*scale/S_a_i, coef[2i+1] = go*scale*S_ab_i/S_a_i^2 for npairs layers (sums as vbx_l1_pair_sums left
 * them): the two scalars of the line above, for the conv epilogue's fm_coef and for vbx_fm_gate_bwd. */
VBX_API int vbx_fm_coef(const double* sums, int32_t npairs, const float* go, float scale, float* coef, void* stream);
/* out = ((g ? g : 0) + coef[0]*sign(y-other) - coef[1]*sign(y)) * (y > 0 ? 1 : gate_slope): the epilogue's gate stage
 * as a stand-alone pass, for a feature whose consumer's input gradient is not part of the backward pass being run
 * (other / coef may be NULL: plain LeakyReLU backward from the activation). */
VBX_API int vbx_fm_gate_bwd(const float* y, const float* other, const float* coef, float gate_slope, const float* g,
                    int64_t n, float* out, void* stream);
/* Hinge (losses/hinge_loss.py:35-43): acc[0] += scale * sum relu(1 - target*c) ; acc is a double. */
VBX_API int vbx_hinge_fwd(const float* c, int64_t n, float target, float scale, double* acc, void* stream);
/* dc = go * scale * (1 - target*c > 0 ? -target : 0) */
VBX_API int vbx_hinge_bwd(const float* c, int64_t n, float target, float scale, const float* go, float* dc,
                  void* stream);
/* double -> float scalar copy(s) with optional scaling */
VBX_API int vbx_d2f(const double* src, float* dst, int32_t n, float scale, void* stream);
/* STFT framing (torch.stft center/reflect semantics restricted to the non-zero window support):
 * U[b,k,f] = x[b, mirror(f*hop + k - pad)], F = (L + 2*pad - K)/hop + 1; and the adjoint
 * dx[b,p] = beta*dx[b,p] + sum dU[b,k,f] over the (f,k) that read p.  The DFT itself is then a
 * pointwise (K=1) conv over the frame axis and runs on the conv kernels. */
VBX_API int vbx_unfold_frames(const float* x, float* U, int32_t B, int32_t L, int32_t K, int32_t hop, int32_t pad,
                      void* stream);
VBX_API int vbx_fold_frames(const float* dU, float* dx, int32_t B, int32_t L, int32_t K, int32_t hop, int32_t pad,
                    float beta, void* stream);
/* STFT loss statistics (auraloss.freq.STFTLoss as configured by multi_stft.yaml:1-18).
 * X, Y: (B, 2*bins, F) outputs of the STFT-as-conv (rows [0,bins) real, [bins,2bins) imaginary).
 * stats[0] += sum (ym-xm)^2, stats[1] += sum ym^2, stats[2] += sum |log xm - log ym| (double),
 * with xm = sqrt(max(re^2+im^2, eps)). */
VBX_API int vbx_stft_stats(const float* X, const float* Y, int32_t B, int32_t bins, int32_t F, float eps,
                   double* stats, void* stream);
/* loss += w * ( sqrt(s0)/sqrt(s1) + s2/count ) for each of nres statistic triples */
VBX_API int vbx_stft_finalize(const double* stats, const double* counts, int32_t nres, float w, float* loss,
                      void* stream);
/* dX = d loss / d X for one resolution (go: device scalar, w: 1/nres) */
VBX_API int vbx_stft_bwd(const float* X, const float* Y, int32_t B, int32_t bins, int32_t F, float eps,
                 const double* stats, double count, const float* go, float w, float* dX, void* stream);

/* total[0] = sum_i lam[i]*x_i[0], terms[i] = lam[i]*x_i[0] for up to 4 device scalars x_i
 * (lam == NULL -> 1).  The `sum(atomic_losses.values())` of lightning_modules/eben.py:106,239. */
VBX_API int vbx_weighted_sum(const float* x0, const float* x1, const float* x2, const float* x3, int32_t n,
                     const float* lam, float* terms, float* total, void* stream);
/* out[i] = go[0] * (lam ? lam[i] : 1) */
VBX_API int vbx_scalar_mul(const float* go, const float* lam, float* out, int32_t n, void* stream);

/* ---- reductions / optimiser ------------------------------------------------------------ */
/* out[0] = sqrt(sum x^2) (float); scratch is a double the kernel zeroes itself is NOT assumed:
 * pass a zeroed double. */
VBX_API int vbx_sumsq(const float* x, int64_t n, double* acc, void* stream);
/* dynamically_balance_losses (lightning_modules/eben.py:222-240) on device scalars:
 * norm_i = sqrt(sumsq[i]); ema: old = first ? norm : old; old = beta*old + (1-beta)*norm;
 * lambda_i = clamp(1/(old_i + 1e-4), 0, 1e4).  mode 0 = "simple", 1 = "ema". */
VBX_API int vbx_balance(const double* sumsq, float* norms_old, int32_t* initialised, float* lambdas,
                float* norms_out, int32_t n, float beta_ema, int32_t mode, void* stream);
/* torch.optim.Adam (optimizer/adam.yaml:1-9), single flat bucket:
 * g = grad_scale*grad; m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
 * p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps);  t = ++step[0] (device int, bumped by
 * the kernel's first block AFTER all reads: launch vbx_adam_tick first). */
VBX_API int vbx_adam_tick(int32_t* step, void* stream);
VBX_API int vbx_adam_step(float* p, const float* grad, float* m, float* v, int64_t n, const int32_t* step,
                  float lr, float b1, float b2, float eps, float grad_scale, void* stream);
VBX_API int vbx_fill(float* p, int64_t n, float value, void* stream);

/* ---- noisy-BWE collate arithmetic (vibravox/utils.py:195-254, 50-81) ------------------- */
/* out_body[b,:] = speech_body[b, off[b] : off[b]+len] + noise[b, start[b] + off[b] : ...];
 * out_air[b,:] = speech_air[b, off[b] : off[b]+len].  start/off are device int arrays. */
VBX_API int vbx_noise_mix_crop(const float* body, const float* air, const float* noise, const int32_t* start,
                       const int32_t* off, float* out_body, float* out_air, int32_t B, int32_t Ls,
                       int32_t Ln, int32_t len, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VBX_H */
