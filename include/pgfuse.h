/* pgfuse.h -- C ABI of libpgfuse.so: the B200 (sm_100a) kernels behind the privatised fusion head
 * of Rachfu/EEG-multimodal.
 *
 * The reference exposes this path only as Python nn.Module code (there is no native layer to
 * bind), so every entry point below replaces a span of torch ops; the span is cited per function
 * (paths relative to the reference checkout).  A reference-side binding is a ctypes stub, shown
 * in INTEGRATION.md; eeg_multimodal_b200/_lib.py is that stub for this repo.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on error; pgf_last_error() gives the
 *     thread-local message.  Nothing here allocates caller-visible memory or synchronises:
 *     all work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - all pointers are DEVICE pointers unless the name says host; row-major; `ld*` are row
 *     strides in ELEMENTS; fp32 unless `*_dtype` says otherwise (PGF_DT_F32=0, PGF_DT_BF16=1).
 *   - "grouped" functions take `n_models` and per-operand model strides `s*` (elements between
 *     consecutive models; 0 = shared) so an eps x seed sweep is one launch.
 *   - scratch space is caller-owned: query the size with the matching *_workspace() call.
 */
#ifndef PGFUSE_H_
#define PGFUSE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PGF_DT_F32 0
#define PGF_DT_BF16 1

#define PGF_NOISE_INJECTED 0 /* caller supplies Laplace [B,D] and Gumbel [2,B,D] tensors (parity) */
#define PGF_NOISE_PHILOX 1   /* counter-based Philox4x32-10 in-kernel (production)               */
#define PGF_NOISE_NONE 2     /* non-private ConcatModel: normalise + concat only (model.py:53-64) */

#define PGF_ACT_NONE 0
#define PGF_ACT_RELU 1
#define PGF_ACT_TANH 2

#define PGF_EPI_STORE_BF16 0
#define PGF_EPI_BIAS_RELU_BF16 1
#define PGF_EPI_BIAS_TANH_BF16 2
/* 3 is not an epilogue any more (a bf16 mask-source tile, superseded by the sign-bit mask 9): rejected as a bad argument */
#define PGF_EPI_ATOMIC_F32 4
#define PGF_EPI_STORE_F32 5
#define PGF_EPI_BIAS_F32 6
#define PGF_EPI_BIAS_TANH_F32 7
#define PGF_EPI_BITMASK_BF16 9 /* C = acc * bit(aux): aux = the uint32 [M, N/32] ReLU sign bits that
                                  PGF_EPI_BIAS_RELU_BF16 writes when its aux is non-NULL (ld_aux in words) */

int pgf_version(void);
const char* pgf_last_error(void);
int pgf_num_sms(void);

/* ---- (a3,a4) per-column privacy coefficients -------------------------------------------------
 * replaces: w = F.sigmoid(self.DP); eps_hat = 1/(((eps.exp()-w)/(1-w)).log())
 *           python/src/custom_models/models.py:73,75 (== past_acc.py:130,132); the unfixed form
 *           (fixed_formula=0) is model.py:57.  Grouped: DP is [n_models, D] and `exp_eps` a DEVICE
 *           array [n_models] holding e^eps of each model, already rounded to fp32 by the host --
 *           exactly the scalar that enters `(eps.exp() - w)`.
 * outputs (each [n_models, D], any may be NULL): w, eps_hat, deps_dDP = d eps_hat / d DP.       */
int pgf_dp_coeffs(const float* DP, const float* exp_eps, int fixed_formula, int D, int n_models, float* w,
                  float* eps_hat, float* deps_dDP, void* stream);

/* ---- (a1,a2,a5,a6,a7) fused concat + row min-max normalise + Laplace perturbation + gate -----
 * replaces: models.py:69-79 (== past_acc.py:120-136): torch.cat, torch.min/max, (x-min)/(max-min),
 *           Laplace.sample on the host + .to(device), feature + noise*eps_hat, F.gumbel_softmax over
 *           the stacked (w, 1-w) planes, (feature*mask).sum(0).
 * x0/x1/x2: up to three feature blocks [B,d_i] (d_i % 4 == 0, d2 may be 0 with x2 NULL); D = sum d_i.
 * noise_mode INJECTED: lap [B,D], gum [2,B,D] (gum may be NULL -> gate skipped);
 *            PHILOX:   seed/offset/row0 key the counter (row0 = global index of row 0, so the
 *                      noise does not depend on how a batch is split across GPUs);
 *            NONE:     out = normalised features.
 * want_gate: evaluate the Gumbel gate (mask applied faithfully in INJECTED mode; in PHILOX mode the
 *            two mask planes sum to one so only the gate index is a real output).
 * out: [B,D] fp32 or bf16 (ld_out); gate_idx [B,D] uint8, row_min/row_max [B]: optional.
 * grouped: n_models launches' worth in one grid; sx* / s_coef / s_out are the model strides of the
 *          blocks (0 = one batch shared by the whole sweep), of w/eps_hat and of out; lap, gum,
 *          gate_idx, row_min/max are contiguous per model; model m uses seed + m*seed_step, or
 *          model_seeds[m] when that DEVICE array [n_models] is given (arbitrary eps x seed grids).   */
int pgf_perturb_gate_fwd(const float* x0, int d0, long long ld0, const float* x1, int d1, long long ld1,
                         const float* x2, int d2, long long ld2, const float* w, const float* eps_hat, int B,
                         int noise_mode, const float* lap, const float* gum, unsigned long long seed,
                         unsigned int offset, unsigned long long row0, float tau, int hard, int want_gate,
                         void* out, int out_dtype, long long ld_out, unsigned char* gate_idx, float* row_min,
                         float* row_max, int n_models, long long sx0, long long sx1, long long sx2,
                         long long s_coef, long long s_out, unsigned long long seed_step,
                         const unsigned long long* model_seeds, void* stream);

/* The same kernel with the three extensions the training / evaluation drivers use (all optional, NULL / 0 / 1 = off):
 *   step_state (pgf_step_state, device): offset += state.noise_offset;
 *   gather != 0: batch row b is row src_rows[cursor + b] (src_rows NULL: cursor + b) of the RESIDENT blocks x0..x2,
 *                cursor = state.cursor (0 without a state) -- a shuffled epoch without host work (data.py:37-45);
 *   n_rep > 1:   the batch is perturbed n_rep times with Philox offsets offset .. offset+n_rep-1, i.e. the n_eval
 *                repeated stochastic evaluations of train.py:126-131 in ONE launch; out (and row_min/row_max,
 *                gate_idx, lap, gum) then have n_rep*B rows per model, repetition-major.                            */
int pgf_perturb_gate_fwd_ex(const float* x0, int d0, long long ld0, const float* x1, int d1, long long ld1,
                            const float* x2, int d2, long long ld2, const float* w, const float* eps_hat, int B,
                            int noise_mode, const float* lap, const float* gum, unsigned long long seed,
                            unsigned int offset, unsigned long long row0, float tau, int hard, int want_gate,
                            void* out, int out_dtype, long long ld_out, unsigned char* gate_idx, float* row_min,
                            float* row_max, int n_models, long long sx0, long long sx1, long long sx2,
                            long long s_coef, long long s_out, unsigned long long seed_step,
                            const unsigned long long* model_seeds, const void* step_state, const long long* src_rows,
                            int gather, int n_rep, void* stream);

/* ---- (a11) dL/dDP through the perturbation ----------------------------------------------------
 * replaces: autograd through models.py:75-76: dDP[d] = deps_dDP[d] * sum_b dF[b,d] * noise[b,d]
 *           (the gate's own contribution is zero in exact arithmetic).
 * dF: gradient wrt the gated feature [B,D] fp32/bf16.  accumulate!=0 adds into dDP.             */
size_t pgf_perturb_gate_bwd_dp_workspace(int B, int D, int n_models);
int pgf_perturb_gate_bwd_dp(const void* dF, int dF_dtype, long long ld, long long s_dF, int B, int D, int n_models,
                            int noise_mode, const float* lap, unsigned long long seed,
                            unsigned long long seed_step, const unsigned long long* model_seeds,
                            unsigned int offset, unsigned long long row0,
                            const float* deps_dDP, long long s_coef, float* workspace, size_t workspace_bytes,
                            float* dDP, long long s_dDP, int accumulate, void* stream);

/* ---- (a11) gradient wrt the raw feature blocks through the min-max normalisation -------------
 * replaces: autograd through models.py:70-72 (only needed when the encoders are trained).
 * dn: gradient wrt the normalised (== perturbed == gated) feature [B,D].                        */
int pgf_minmax_norm_bwd(const float* x0, int d0, long long ld0, const float* x1, int d1, long long ld1,
                        const float* x2, int d2, long long ld2, const void* dn, int dn_dtype, long long ld_dn,
                        int B, float* dx0, long long ldd0, float* dx1, long long ldd1, float* dx2,
                        long long ldd2, void* stream);

/* ---- (a8) fusion MLP, fp32 CUDA-core path (small batch, weight-streaming bound), grouped -----
 * replaces: nn.Linear(+ReLU/Tanh) of fc_layers and autograd, models.py:46-51,80.
 * W is torch layout [N,K] (out,in).  fwd: Y = act(X W^T + bias).
 * dx:  dX = dY W, optionally times the derivative of the activation that produced mask_src
 *      (mask_mode PGF_ACT_RELU: (mask_src > 0); PGF_ACT_TANH: (1 - mask_src^2)).
 * dw:  dW = dY^T X (+= if accumulate), db = colsum(dY) (db may be NULL).                        */
int pgf_linear_fwd(const float* X, long long ldx, long long sX, const float* W, long long sW, const float* bias,
                   long long sb, float* Y, long long ldy, long long sY, int B, int N, int K, int act,
                   int n_models, void* stream);
size_t pgf_linear_bwd_dx_workspace(int B, int N, int K, int n_models);
int pgf_linear_bwd_dx(const float* dY, long long ldy, long long sdY, const float* W, long long sW,
                      const float* mask_src, int mask_mode, long long ld_mask, long long s_mask, float* dX, long long ldx,
                      long long sdX, int B, int N, int K, int n_models, float* workspace,
                      size_t workspace_bytes, void* stream);
int pgf_linear_bwd_dw(const float* dY, long long ldy, long long sdY, const float* X, long long ldx, long long sX,
                      float* dW, long long sdW, float* db, long long sdb, int B, int N, int K, int accumulate,
                      int n_models, void* stream);
/* The same gradient with caller scratch (pgf_linear_bwd_dw_workspace bytes; 0 = not needed): a narrow layer (N <= 8: the
 * 768 -> 2 classifier, models.py:81) at a large batch (B >= 512) is computed by batch slabs in parallel, the slab partials
 * summed in slab order; every other shape takes pgf_linear_bwd_dw's kernel.                                          */
size_t pgf_linear_bwd_dw_workspace(int B, int N, int K, int n_models);
int pgf_linear_bwd_dw_ex(const float* dY, long long ldy, long long sdY, const float* X, long long ldx, long long sX,
                         float* dW, long long sdW, float* db, long long sdb, int B, int N, int K, int accumulate,
                         int n_models, float* workspace, size_t workspace_bytes, void* stream);

/* ---- (a8) fusion MLP, tcgen05 tensor-core path (large batch) ----------------------------------
 * replaces: the same nn.Linear GEMMs when the batch is a real dense contraction.
 * C[M,N] = A[M,K] . B[N,K]^T, bf16 operands, fp32 accumulation in TMEM, fused epilogue `epi`.
 * a_mn / b_mn != 0: that operand is stored transposed ([K,M] / [K,N] row-major), as the
 * weight-gradient GEMMs need.  stream_k != 0 (with PGF_EPI_ATOMIC_F32, C zeroed by the caller)
 * splits K across CTAs.  bias [N] fp32; aux: uint32 [M, ld_aux >= N/32] ReLU sign bits (N % 128 == 0):
 * OUT (optional) for PGF_EPI_BIAS_RELU_BF16, IN for PGF_EPI_BITMASK_BF16; col_partial: below.     */
int pgf_gemm_bf16(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, void* C,
                  long long ldc, int M, int N, int K, int epi, const float* bias, void* aux,
                  long long ld_aux, int stream_k, float* col_partial, void* stream);

/* ---- (a8, a11) the same GEMMs held to the reference's fp32 arithmetic (north_star's 1e-5 bar) at large batch ----------
 * replaces: fp32 nn.Linear of fc_layers and its autograd, models.py:46-51,80, when the batch is a real dense contraction
 *           and bf16 operand rounding (2e-2 bar) is not acceptable.
 * pgf_split3: the elementwise stage between two such GEMMs.  v = act(src + bias) (act PGF_ACT_*; tanhf, bias optional),
 *           then v *= (mask_plane > 0) when mask_plane (the hi plane of a ReLU output, bf16 [R, ld_mask]) is given -- the
 *           ReLU backward of autograd; writes v to out_f32 (optional; may alias src) and/or as three bf16 planes
 *           hi/mid/lo with hi + mid + lo == v exactly (planes: bf16, plane p at planes + p*plane_stride, row stride ldp).
 * pgf_gemm_bf16x3: C[M,N] (fp32) = A . B^T with A and B given as such plane triples; six plane-pair products down to
 *           2^-24 of the result, one fp32 accumulation in TMEM, same kernel and operand layouts (a_mn / b_mn) as
 *           pgf_gemm_bf16.  epi: PGF_EPI_STORE_F32, PGF_EPI_BIAS_F32, or PGF_EPI_ATOMIC_F32 with k_slabs >= 1 (C zeroed by
 *           the caller; 1 = the launcher picks the number of K slabs, > 1 = exactly that many: tensor-core accumulation
 *           truncates, so long contractions are cut into slabs combined by round-to-nearest fp32 adds).            */
int pgf_split3(const float* src, long long ld, int R, int C, const float* bias, int act, const void* mask_plane,
               long long ld_mask, float* out_f32, long long ld_out, void* planes, long long ldp, long long plane_stride,
               void* stream);
int pgf_gemm_bf16x3(const void* A3, long long lda, long long a_plane, int a_mn, const void* B3, long long ldb,
                    long long b_plane, int b_mn, float* C, long long ldc, int M, int N, int K, int epi, const float* bias,
                    int k_slabs, void* stream);

/* Fused bias gradient: with a bf16-output epilogue, col_partial (optional, fp32
 * [pgf_gemm_partial_rows(M)][N]) receives the column sums of the epilogue values of every 32-row
 * accumulator slab (one per epilogue warp); pgf_reduce_partials() sums the slabs in a fixed order:
 *   out[n] = (coef ? coef[n] : 1) * sum_r partial[r][n]      (+= if accumulate)
 * replaces: the `.sum(0)` of autograd's nn.Linear bias gradient (fc_layers.0.bias).              */
int pgf_gemm_partial_rows(int M);
/* SMs the persistent GEMM grids launched by THIS thread leave free from now on (0 = none; default).  The data-parallel
 * mode sets it while a gradient bucket is being all-reduced on a side stream, so that the collective's kernels and the
 * one-CTA-per-SM GEMM grid are co-resident instead of queueing behind each other. */
int pgf_set_sm_reserve(int n_sms);
int pgf_reduce_partials(const float* partial, int rows, int N, const float* coef, float* out, int accumulate,
                        void* stream);

/* ---- (a11) dL/dDP fused into the input-gradient GEMM -------------------------------------------
 * replaces: dX = dZ1 . W1 (autograd of fc_layers.0, models.py:80) followed by the reduction of
 *           pgf_perturb_gate_bwd_dp: dDP[n] = deps_dDP[n] * sum_m (A . B^T)[m,n] * Laplace(row0+m, n),
 *           the noise regenerated from (seed, offset) exactly as pgf_perturb_gate_fwd drew it.
 *           The [M,N] product only ever exists in TMEM; workspace: pgf_gemm_partial_rows(M)*N floats. */
int pgf_gemm_bf16_ddp(const void* A, long long lda, const void* B, long long ldb, int b_mn, int M, int N, int K,
                      unsigned long long seed, unsigned int offset, unsigned long long row0,
                      const float* deps_dDP, float* workspace, size_t workspace_bytes, float* dDP,
                      int accumulate, void* stream);

/* ---- (a9,a10,a11) classifier + softmax cross-entropy + accuracy, fwd (+bwd), grouped ---------
 * replaces: self.classifier (models.py:81) and cal_loss (base_train.py:59-65 == past_acc.py:71-77)
 *           and their autograd, including the gradient through the Tanh in front of the classifier.
 * h [B,H] (fp32/bf16) is the Tanh output; labels int64 [B] (slabels = 0 shares them);
 * outputs: logits [B,2], pred int64 [B] (optional); stats[model*4 + {0..3}] =
 *          {loss_sum*loss_scale, n_correct, n_correct*loss_scale, B};
 * backward!=0: dz [B,H] = (dlogits . Wc) * (1-h^2 if through_tanh), with
 *          dlogits = (softmax - onehot) * grad_scale; optional (NULL = not formed, as in pass 1 of
 *          the reference step where zero_grad discards them, past_acc.py:206): dWc [2,H], dbc [2],
 *          dz_colsum [H] = column sums of dz = the bias gradient of fc_layers.2.                  */
size_t pgf_cls_ce_workspace(int B, int H, int n_models);
int pgf_cls_ce(const void* h, int h_dtype, long long ldh, long long sh, const float* Wc, long long sWc,
               const float* bc, long long sbc, const long long* labels, long long slabels, int B, int H,
               int n_models, float loss_scale, float grad_scale, int backward, int through_tanh, float* logits,
               long long slogits, long long* pred, long long spred, float* stats, void* dz, int dz_dtype,
               long long lddz, long long sdz, float* dWc, long long sdWc, float* dbc, long long sdbc,
               float* dz_colsum, long long sdz_colsum, float* workspace, size_t workspace_bytes, void* stream);

/* ---- (a12,f1) Adam over a flat fp32 buffer ----------------------------------------------------
 * replaces: torch.optim.Adam(...).step() for either parameter group (past_acc.py:155-160,203,212),
 *           torch defaults; `step` is the 1-based step count.  `bf16_shadow` (optional) receives a
 *           bf16 copy of the updated parameters for the tensor-core path.                        */
int pgf_adam_step(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n, int step,
                  float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* The same update over the segment [0,n) of each of n_models flat buffers `model_stride` elements apart
 * (one launch for, e.g., the classifier parameters of a whole sweep).                             */
int pgf_adam_step_strided(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n,
                          long long model_stride, int n_models, int step, float lr, float beta1, float beta2,
                          float eps, float grad_scale, void* stream);

/* ---- (a11,a12,f1) weight gradient fused into Adam, small batch (B <= 8), grouped -----------------
 * replaces: autograd's dW = dY^T X, db = colsum(dY) of one nn.Linear (models.py:46-51) followed by
 *           model_optimizer.step() on that layer (past_acc.py:212).  At the reference batch size the
 *           gradient is a rank-8 outer product: it is recomputed per element inside the optimiser, so
 *           dW is never written to / read from HBM (24 instead of 32 bytes per parameter and step).
 * W, mW, vW: [N,K] weight and Adam moments; bias, mb, vb: [N] (optional); all six live in per-model
 * flat buffers with the common model stride sP.  Same update arithmetic as pgf_adam_step.         */
int pgf_linear_adam_step(const float* dY, long long ldy, long long sdY, const float* X, long long ldx,
                         long long sX, int B, int N, int K, float* W, float* mW, float* vW, float* bias,
                         float* mb, float* vb, long long sP, int step, float lr, float beta1, float beta2,
                         float eps, float grad_scale, int n_models, void* stream);

/* ---- helpers for the tensor-core path ---------------------------------------------------------*/
int pgf_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream);
/* zero-fill (the split-K weight-gradient GEMMs accumulate into a zeroed buffer; replaces grad.zero_()) */
int pgf_fill_zero(void* p, size_t nbytes, void* stream);
/* Device-to-device copy between two GPUs of the box on `stream` (a stream of the CURRENT device), executed by the copy
 * engines over NVLink: cudaMemcpyPeerAsync with direct peer access enabled on first use.  dst / src may be mappings of
 * another process's allocation (CUDA IPC).  Used to fan a shared sweep batch out to the peers' input buffers without SMs
 * (replaces: one full DataLoader batch `.cuda()` per process, past_acc.py:192-196 run once per GPU). */
int pgf_memcpy_peer_async(void* dst, int dst_device, const void* src, int src_device, size_t nbytes, void* stream);
size_t pgf_colsum_workspace(int B, int N);
int pgf_colsum(const void* x, int dtype, long long ld, int B, int N, float* out, float* workspace,
               size_t workspace_bytes, void* stream);

/* ---- (a-alt) the OLDER PriGumbel head tail: Gumbel dropout + row Laplace + w-loss, fwd and bwd ----
 * replaces: gumbel_dropout (train_val.py:95-101), Lap_noise (train_val.py:114-123), the w term of
 *           loss_function (train_val.py:80-93: max_j((1-w_j) e^eps + w_j)) and their autograd, i.e. what
 *           sits between fc2 and the classifier in train_val.py:151-157.
 * coef:  per-forward column coefficients from w [H] and ONE [H,2] Gumbel draw (NULL = Philox:
 *        counter (j/4, 0, Gumbel plane, offset)); hard = eval mode (straight-through composite),
 *        soft = train mode (train_val.py:108-111).  coef [4,H] = {mask, 1-w, d mask/d w, 1/(1-w)};
 *        wloss [2] = {max_j((1-w_j) e^eps + w_j), its arg-max}.
 * fwd:   out[b,:] = minmax_row((z[b,:] * mask) / (1-w)) + n_b, n_b = lap[b] (injected Laplace(0,1/eps)
 *        draws) or Philox: counter (0, row0+b, Laplace, offset), scaled by 1/eps.
 * bwd:   dz [B,H] (gradient of fc2's output, full min-max backward with torch's first-occurrence
 *        arg-min/arg-max routing) and, when dw != NULL, dw [H] (+= if accumulate) = gradient through
 *        the gate, through 1/(1-w), and wloss_scale * d wloss/dw (1 - e^eps at the arg-max).
 *        workspace: pgf_prigumbel_bwd_workspace(B,H) bytes, fixed-order reduction, no atomics.      */
int pgf_prigumbel_coef(const float* w, const float* gumbel, int H, float exp_eps, float tau, int hard,
                       unsigned long long seed, unsigned int offset, float* coef, float* wloss, void* stream);
int pgf_prigumbel_fwd(const float* z, long long ldz, const float* coef, const float* lap, float eps,
                      unsigned long long seed, unsigned int offset, unsigned long long row0, float* out,
                      long long ld_out, float* row_min, float* row_max, int B, int H, void* stream);
size_t pgf_prigumbel_bwd_workspace(int B, int H);
int pgf_prigumbel_bwd(const float* z, long long ldz, const float* coef, const float* dout, long long ld_dout,
                      const float* wloss, float exp_eps, float wloss_scale, float* dz, long long ld_dz, float* dw,
                      int accumulate, int B, int H, float* workspace, size_t workspace_bytes, void* stream);

/* ---- (a12, f1, f3) the whole reference step at the reference batch size, one constant launch sequence ------------
 * replaces: one iteration of the training loop past_acc.py:198-212 (== base_train.py:183-210) for EVERY model of a
 *           sweep -- pass 1 (hard=False) forward / cal_loss / backward / DP_optimizer.step(), pass 2 (hard=True)
 *           forward / cal_loss / backward / model_optimizer.step() -- plus the shuffled DataLoader fetch in front of it
 *           (data.py:37-45): the batch is rows src_rows[cursor .. cursor+B) of feature blocks resident in HBM.
 * Everything that changes between steps lives in a 64-byte DEVICE struct (pgf_step_state: Philox offset, Adam step
 * counts and bias corrections, batch cursor) that the kernels add to their constant arguments, so the 13-launch chain
 * is captured once into a CUDA graph and replayed; kernels are chained with programmatic dependent launch.  Same
 * kernels and arithmetic as the per-kernel entry points above: bit-identical results.
 *
 * pgf_step_state_set: (re)initialise the device state.  noise_offset / t_dp / t_model: Philox offset of the next
 *           forward and Adam steps already taken by the DP / weight optimisers; cursor: first row of the next batch.
 * pgf_sweep_desc: every pointer is a DEVICE pointer owned by the caller and must stay valid for the plan's life.
 *           params / adam_m / adam_v / grads: per-model flat fp32 buffers `P` elements apart, segments at off_*;
 *           coef: [3][n_models][D] = (w, eps_hat, d eps_hat/d DP), current on entry (pgf_dp_coeffs) and kept current;
 *           stats_*: [n_models][4] as pgf_cls_ce writes them, of pass 1 / pass 2; logits / pred (optional): pass 2.
 *           n_rows > 0: the cursor wraps to 0 when the next batch would run past row n_rows.
 * pgf_sweep_plan_reset: zero the plan's internal counters (once, before the first run, on the run stream).
 * pgf_sweep_plan_capture: record `steps_per_graph` consecutive steps into one CUDA graph on `stream`.
 * pgf_sweep_plan_run: enqueue n_steps steps (graph replays while n_steps allows, direct launches for the rest).     */
typedef struct pgf_sweep_desc {
  int n_models, B, d0, d1, d2, H;
  int dp_pass;        /* 1: two-pass step (past_acc.py); 0: pass 2 only (train.py:100-105 has pass 1 commented out) */
  int fixed_formula;  /* eps_hat = 1/log(..) (past_acc.py:132) or log(..) (model.py:57) */
  int use_pdl;        /* chain the kernels with programmatic dependent launch */
  int reserved_;
  float tau, lr, beta1, beta2, adam_eps, reserved_f_;
  const float* x0; long long ld0;
  const float* x1; long long ld1;
  const float* x2; long long ld2;
  const long long* labels;      /* [rows of the resident dataset] */
  const long long* src_rows;    /* permutation of an epoch, indexed by the cursor; NULL = identity */
  long long n_rows;
  float* params; float* adam_m; float* adam_v; float* grads; long long P;
  long long off_W1, off_b1, off_W2, off_b2, off_Wc, off_bc;
  float* DP; float* DP_m; float* DP_v; float* dDP;
  float* coef;
  const float* exp_eps;
  const unsigned long long* seeds;
  unsigned long long row0;
  float* stats_dp; float* stats_model;
  float* logits; long long* pred;
  void* state;
  void* workspace; size_t workspace_bytes;
} pgf_sweep_desc;

int pgf_step_state_set(void* state, long long noise_offset, long long t_dp, long long t_model, long long cursor, float lr,
                       float beta1, float beta2, void* stream);
size_t pgf_sweep_plan_workspace(int n_models, int B, int D, int H);
int pgf_sweep_plan_create(const pgf_sweep_desc* desc, void** plan_out);
int pgf_sweep_plan_reset(void* plan, void* stream);
int pgf_sweep_plan_capture(void* plan, void* stream, int steps_per_graph);
int pgf_sweep_plan_run(void* plan, void* stream, int n_steps);
int pgf_sweep_plan_launches_per_step(void* plan);
int pgf_sweep_plan_destroy(void* plan);

#ifdef __cplusplus
}
#endif
#endif /* PGFUSE_H_ */
