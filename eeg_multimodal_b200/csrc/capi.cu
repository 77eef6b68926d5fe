// extern "C" layer of libpgfuse.so: argument validation + dispatch to the kernels.
// Contract: include/pgfuse.h.
#include <map>
#include <mutex>
#include <stdarg.h>
#include <string.h>

#include "../../include/pgfuse.h"
#include <stdlib.h>

#include "pgf_kernels.cuh"

namespace pgf {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int num_sms() {
  static thread_local int cached_dev = -1, cached = 0;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  if (dev != cached_dev) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached = n;
    cached_dev = dev;
  }
  return cached;
}

namespace {
struct KernelKey {
  int dev;
  const void* k;
  int threads;
  size_t smem;
  bool operator<(const KernelKey& o) const {
    if (dev != o.dev) return dev < o.dev;
    if (k != o.k) return k < o.k;
    if (threads != o.threads) return threads < o.threads;
    return smem < o.smem;
  }
};
std::mutex g_cache_mutex;
std::map<KernelKey, size_t> g_smem_set;   // (dev, kernel) -> largest dynamic smem size configured so far
std::map<KernelKey, int> g_occupancy;     // (dev, kernel, threads, smem) -> resident CTAs per SM
}  // namespace

void ensure_dynamic_smem(const void* kernel, size_t smem) {
  if (smem <= 48 * 1024) return;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  size_t& cur = g_smem_set[KernelKey{dev, kernel, 0, 0}];
  if (smem > cur) {
    cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    cur = smem;
  }
}

int cached_occupancy(const void* kernel, int threads, size_t smem, int fallback) {
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(g_cache_mutex);
  auto it = g_occupancy.find(KernelKey{dev, kernel, threads, smem});
  if (it != g_occupancy.end()) return it->second;
  int occ = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem) != cudaSuccess || occ < 1) occ = fallback;
  g_occupancy[KernelKey{dev, kernel, threads, smem}] = occ;
  return occ;
}

static thread_local bool g_pdl = false;
bool pdl_enabled() { return g_pdl; }
PdlScope::PdlScope(bool on) : prev(g_pdl) { g_pdl = on; }
PdlScope::~PdlScope() { g_pdl = prev; }

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

}  // namespace pgf

using namespace pgf;

extern "C" {

int pgf_version(void) { return 100; }
const char* pgf_last_error(void) { return g_err; }
int pgf_num_sms(void) { return num_sms(); }

int pgf_dp_coeffs(const float* DP, const float* exp_eps, int fixed_formula, int D, int n_models, float* w, float* eps_hat,
                  float* deps_dDP, void* stream) {
  PGF_CHECK_ARG(DP && exp_eps && D > 0 && n_models > 0, "pgf_dp_coeffs: DP / exp_eps is NULL or D, n_models <= 0");
  return dp_coeffs(DP, exp_eps, fixed_formula, D, n_models, w, eps_hat, deps_dDP, static_cast<cudaStream_t>(stream));
}

int pgf_perturb_gate_fwd_ex(const float* x0, int d0, long long ld0, const float* x1, int d1, long long ld1, const float* x2,
                         int d2, long long ld2, const float* w, const float* eps_hat, int B, int noise_mode,
                         const float* lap, const float* gum, unsigned long long seed, unsigned int offset,
                         unsigned long long row0, float tau, int hard, int want_gate, void* out, int out_dtype,
                         long long ld_out, unsigned char* gate_idx, float* row_min, float* row_max, int n_models,
                         long long sx0, long long sx1, long long sx2, long long s_coef, long long s_out,
                         unsigned long long seed_step, const unsigned long long* model_seeds, const void* step_state,
                         const long long* src_rows, int gather, int n_rep, void* stream) {
  if (B == 0 || n_models == 0) return PGF_OK;  // empty batch: nothing to do
  PGF_CHECK_ARG(n_models > 0 && (sx0 % 4) == 0 && (sx1 % 4) == 0 && (sx2 % 4) == 0 && (s_coef % 4) == 0 && (s_out % 4) == 0,
                "pgf_perturb_gate_fwd: n_models < 0 or model strides not multiples of 4");
  PGF_CHECK_ARG(B > 0, "pgf_perturb_gate_fwd: B < 0");
  PGF_CHECK_ARG(x0 && d0 > 0, "pgf_perturb_gate_fwd: first feature block is required");
  PGF_CHECK_ARG(d1 >= 0 && d2 >= 0 && (d1 == 0 || x1) && (d2 == 0 || x2), "pgf_perturb_gate_fwd: block pointer/size mismatch");
  PGF_CHECK_ARG(d1 > 0 || d2 == 0, "pgf_perturb_gate_fwd: block 2 given without block 1");
  PGF_CHECK_ARG((d0 % 4) == 0 && (d1 % 4) == 0 && (d2 % 4) == 0, "pgf_perturb_gate_fwd: block widths must be multiples of 4");
  PGF_CHECK_ARG((ld0 % 4) == 0 && (ld1 % 4) == 0 && (ld2 % 4) == 0 && aligned16(x0) && aligned16(x1) && aligned16(x2),
                "pgf_perturb_gate_fwd: feature blocks must be 16-byte aligned with row strides multiple of 4");
  PGF_CHECK_ARG(out && aligned16(out) && (ld_out % 4) == 0, "pgf_perturb_gate_fwd: out must be 16-byte aligned");
  PGF_CHECK_ARG(out_dtype == PGF_DT_F32 || out_dtype == PGF_DT_BF16, "pgf_perturb_gate_fwd: bad out_dtype");
  PGF_CHECK_ARG(noise_mode >= 0 && noise_mode <= 2, "pgf_perturb_gate_fwd: bad noise_mode %d", noise_mode);
  if (noise_mode != PGF_NOISE_NONE) PGF_CHECK_ARG(w && eps_hat, "pgf_perturb_gate_fwd: w / eps_hat required");
  if (noise_mode == PGF_NOISE_INJECTED) PGF_CHECK_ARG(lap && aligned16(lap), "pgf_perturb_gate_fwd: injected mode needs lap");
  bool gate = want_gate != 0 && noise_mode != PGF_NOISE_NONE;
  if (noise_mode == PGF_NOISE_INJECTED && !gum) gate = false;
  PGF_CHECK_ARG(tau > 0.f, "pgf_perturb_gate_fwd: tau must be > 0");
  PGF_CHECK_ARG(n_rep >= 1 && n_rep <= 4096, "pgf_perturb_gate_fwd_ex: n_rep must be in 1..4096");
  PerturbFwdArgs a = {};
  a.n_rep = 1;
  a.x[0] = x0; a.x[1] = x1; a.x[2] = x2;
  a.ld[0] = ld0; a.ld[1] = ld1; a.ld[2] = ld2;
  a.sx[0] = sx0; a.sx[1] = sx1; a.sx[2] = sx2;
  a.d[0] = d0; a.d[1] = d1; a.d[2] = d2;
  a.D = d0 + d1 + d2;
  a.B = B;
  a.n_models = n_models; a.s_coef = s_coef; a.s_out = s_out; a.seed_step = seed_step; a.model_seeds = model_seeds;
  memset(&a.rk, 0, sizeof(a.rk));
  a.w = w; a.eps_hat = eps_hat; a.lap = lap; a.gum = gum;
  a.seed = seed; a.offset = offset; a.row0 = row0;
  a.tau = tau; a.inv_tau = 1.0f / tau; a.hard = hard;
  a.out = out; a.ld_out = ld_out; a.gate_idx = gate_idx; a.row_min = row_min; a.row_max = row_max;
  a.st = static_cast<const StepState*>(step_state); a.src_rows = src_rows; a.gather = gather != 0; a.n_rep = n_rep;
  return perturb_gate_fwd(a, noise_mode, out_dtype, gate, static_cast<cudaStream_t>(stream));
}

int pgf_perturb_gate_fwd(const float* x0, int d0, long long ld0, const float* x1, int d1, long long ld1, const float* x2,
                         int d2, long long ld2, const float* w, const float* eps_hat, int B, int noise_mode,
                         const float* lap, const float* gum, unsigned long long seed, unsigned int offset,
                         unsigned long long row0, float tau, int hard, int want_gate, void* out, int out_dtype,
                         long long ld_out, unsigned char* gate_idx, float* row_min, float* row_max, int n_models,
                         long long sx0, long long sx1, long long sx2, long long s_coef, long long s_out,
                         unsigned long long seed_step, const unsigned long long* model_seeds, void* stream) {
  return pgf_perturb_gate_fwd_ex(x0, d0, ld0, x1, d1, ld1, x2, d2, ld2, w, eps_hat, B, noise_mode, lap, gum, seed, offset, row0, tau,
                                 hard, want_gate, out, out_dtype, ld_out, gate_idx, row_min, row_max, n_models, sx0, sx1, sx2, s_coef,
                                 s_out, seed_step, model_seeds, nullptr, nullptr, 0, 1, stream);
}

size_t pgf_perturb_gate_bwd_dp_workspace(int B, int D, int n_models) {
  if (B <= 0 || D <= 0 || n_models <= 0) return 0;
  return static_cast<size_t>(n_models) * perturb_bwd_slabs(B, D, n_models) * D * sizeof(float);
}

int pgf_perturb_gate_bwd_dp(const void* dF, int dF_dtype, long long ld, long long s_dF, int B, int D, int n_models,
                            int noise_mode, const float* lap, unsigned long long seed, unsigned long long seed_step,
                            const unsigned long long* model_seeds, unsigned int offset, unsigned long long row0,
                            const float* deps_dDP, long long s_coef,
                            float* workspace, size_t workspace_bytes, float* dDP, long long s_dDP, int accumulate,
                            void* stream) {
  PGF_CHECK_ARG(D > 0 && (D % 4) == 0 && dDP && deps_dDP && n_models >= 0, "pgf_perturb_gate_bwd_dp: bad D / NULL outputs");
  if (n_models == 0) return PGF_OK;
  if (B == 0) {
    if (!accumulate)
      for (int m = 0; m < n_models; ++m) cudaMemsetAsync(dDP + m * s_dDP, 0, sizeof(float) * D, static_cast<cudaStream_t>(stream));
    return PGF_OK;
  }
  PGF_CHECK_ARG(B > 0 && dF && aligned16(dF) && (ld % 4) == 0 && (s_dF % 4) == 0,
                "pgf_perturb_gate_bwd_dp: dF must be 16-byte aligned");
  PGF_CHECK_ARG(noise_mode == PGF_NOISE_INJECTED || noise_mode == PGF_NOISE_PHILOX, "pgf_perturb_gate_bwd_dp: bad noise_mode");
  if (noise_mode == PGF_NOISE_INJECTED) PGF_CHECK_ARG(lap, "pgf_perturb_gate_bwd_dp: injected mode needs lap");
  PGF_CHECK_ARG(workspace, "pgf_perturb_gate_bwd_dp: workspace is NULL");
  return perturb_gate_bwd_dp(dF, dF_dtype, ld, s_dF, B, D, n_models, noise_mode, lap, seed, seed_step, model_seeds, offset, row0,
                             deps_dDP, s_coef, workspace, workspace_bytes, dDP, s_dDP, accumulate, static_cast<cudaStream_t>(stream));
}

int pgf_minmax_norm_bwd(const float* x0, int d0, long long ld0, const float* x1, int d1, long long ld1, const float* x2,
                        int d2, long long ld2, const void* dn, int dn_dtype, long long ld_dn, int B, float* dx0,
                        long long ldd0, float* dx1, long long ldd1, float* dx2, long long ldd2, void* stream) {
  if (B == 0) return PGF_OK;
  PGF_CHECK_ARG(B > 0 && x0 && dx0 && dn && d0 > 0, "pgf_minmax_norm_bwd: NULL argument");
  PGF_CHECK_ARG((d1 == 0 || (x1 && dx1)) && (d2 == 0 || (x2 && dx2)), "pgf_minmax_norm_bwd: block pointer/size mismatch");
  PGF_CHECK_ARG((d0 % 4) == 0 && (d1 % 4) == 0 && (d2 % 4) == 0, "pgf_minmax_norm_bwd: block widths must be multiples of 4");
  NormBwdArgs a;
  a.x[0] = x0; a.x[1] = x1; a.x[2] = x2;
  a.ld[0] = ld0; a.ld[1] = ld1; a.ld[2] = ld2;
  a.dx[0] = dx0; a.dx[1] = dx1; a.dx[2] = dx2;
  a.ld_dx[0] = ldd0; a.ld_dx[1] = ldd1; a.ld_dx[2] = ldd2;
  a.d[0] = d0; a.d[1] = d1; a.d[2] = d2;
  a.D = d0 + d1 + d2;
  a.B = B;
  a.dn = dn;
  a.ld_dn = ld_dn;
  return minmax_norm_bwd(a, dn_dtype, static_cast<cudaStream_t>(stream));
}

int pgf_linear_fwd(const float* X, long long ldx, long long sX, const float* W, long long sW, const float* bias,
                   long long sb, float* Y, long long ldy, long long sY, int B, int N, int K, int act, int n_models,
                   void* stream) {
  if (B == 0 || n_models == 0) return PGF_OK;
  PGF_CHECK_ARG(X && W && Y && B > 0 && N > 0 && K > 0 && n_models > 0, "pgf_linear_fwd: NULL or non-positive argument");
  PGF_CHECK_ARG((K % 4) == 0 && (ldx % 4) == 0 && aligned16(X) && aligned16(W) && (sX % 4) == 0 && (sW % 4) == 0,
                "pgf_linear_fwd: K, ldx, strides must be multiples of 4 and X, W 16-byte aligned");
  LinFwdArgs a;
  a.X = X; a.ldx = ldx; a.sX = sX; a.W = W; a.sW = sW; a.bias = bias; a.sb = sb; a.Y = Y; a.ldy = ldy; a.sY = sY;
  a.B = B; a.N = N; a.K = K; a.act = act;
  return linear_fwd(a, n_models, static_cast<cudaStream_t>(stream));
}

// Few models per GPU at the reference batch: the TMA-ring kernels of linear_stream.cu (short launches, ramp-up and tail
// matter).  Many models per launch: the slab kernels of linear_wide.cu (millisecond launches, page locality matters).
// PGF_LINEAR_WIDE=0/1 forces one or the other.
static bool wide_regime(int B, int n_models) {
  static const char* env = getenv("PGF_LINEAR_WIDE");
  if (env) return env[0] == '1';
  return B <= 8 && n_models >= PGF_WIDE_MODELS;
}

size_t pgf_linear_bwd_dx_workspace(int B, int N, int K, int n_models) {
  if (B <= 0 || n_models <= 0) return 0;
  const size_t ring = linear_dx_workspace(B, N, K, n_models);
  if (!wide_regime(B, n_models)) return ring;
  const size_t wide = linear_dx_workspace_wide(B, N, K, n_models);
  return wide > ring ? wide : ring;
}

int pgf_linear_bwd_dx(const float* dY, long long ldy, long long sdY, const float* W, long long sW, const float* mask_src,
                      int mask_mode, long long ld_mask, long long s_mask, float* dX, long long ldx, long long sdX, int B, int N, int K,
                      int n_models, float* workspace, size_t workspace_bytes, void* stream) {
  if (B == 0 || n_models == 0) return PGF_OK;
  PGF_CHECK_ARG(dY && W && dX && workspace && B > 0 && N > 0 && K > 0, "pgf_linear_bwd_dx: NULL or non-positive argument");
  PGF_CHECK_ARG((K % 4) == 0 && (ldx % 4) == 0 && aligned16(W) && aligned16(dX) && (sW % 4) == 0 && (sdX % 4) == 0 &&
                    (!mask_src || ((ld_mask % 4) == 0 && aligned16(mask_src) && (s_mask % 4) == 0)),
                "pgf_linear_bwd_dx: K, ldx, strides must be multiples of 4 and W, dX, mask 16-byte aligned");
  if (linear_dx_narrow_applies(B, N))   // a narrow layer at a large batch (the classifier behind the module API): elementwise
    return linear_bwd_dx_narrow(dY, ldy, sdY, W, sW, mask_src, mask_mode, ld_mask, s_mask, dX, ldx, sdX, B, N, K, n_models,
                                static_cast<cudaStream_t>(stream));
  if (wide_regime(B, n_models))
    return linear_bwd_dx_wide(dY, ldy, sdY, W, sW, mask_src, mask_mode, ld_mask, s_mask, dX, ldx, sdX, B, N, K, n_models, workspace,
                              workspace_bytes, static_cast<cudaStream_t>(stream));
  return linear_bwd_dx(dY, ldy, sdY, W, sW, mask_src, mask_mode, ld_mask, s_mask, dX, ldx, sdX, B, N, K, n_models, workspace,
                       workspace_bytes, static_cast<cudaStream_t>(stream));
}

size_t pgf_linear_bwd_dw_workspace(int B, int N, int K, int n_models) {
  if (B <= 0 || N <= 0 || K <= 0 || n_models <= 0) return 0;
  return linear_dw_batch_workspace(B, N, K, n_models);
}

int pgf_linear_bwd_dw_ex(const float* dY, long long ldy, long long sdY, const float* X, long long ldx, long long sX, float* dW,
                         long long sdW, float* db, long long sdb, int B, int N, int K, int accumulate, int n_models,
                         float* workspace, size_t workspace_bytes, void* stream) {
  if (n_models > 0 && B > 0 && workspace && linear_dw_batch_applies(B, N) && workspace_bytes >= linear_dw_batch_workspace(B, N, K, n_models)) {
    PGF_CHECK_ARG(dY && X && dW && N > 0 && K > 0, "pgf_linear_bwd_dw_ex: NULL or non-positive argument");
    PGF_CHECK_ARG((K % 4) == 0 && (ldx % 4) == 0 && aligned16(X) && aligned16(dW) && (sX % 4) == 0 && (sdW % 4) == 0 && aligned16(workspace),
                  "pgf_linear_bwd_dw_ex: K, ldx, strides must be multiples of 4 and X, dW, workspace 16-byte aligned");
    LinDwArgs a;
    a.dY = dY; a.ldy = ldy; a.sdY = sdY; a.X = X; a.ldx = ldx; a.sX = sX; a.dW = dW; a.sdW = sdW; a.db = db; a.sdb = sdb;
    a.B = B; a.N = N; a.K = K; a.rows_per_cta = 0; a.accumulate = accumulate;
    return linear_bwd_dw_batch(a, n_models, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
  }
  return pgf_linear_bwd_dw(dY, ldy, sdY, X, ldx, sX, dW, sdW, db, sdb, B, N, K, accumulate, n_models, stream);
}

int pgf_linear_bwd_dw(const float* dY, long long ldy, long long sdY, const float* X, long long ldx, long long sX, float* dW,
                      long long sdW, float* db, long long sdb, int B, int N, int K, int accumulate, int n_models,
                      void* stream) {
  if (n_models == 0) return PGF_OK;
  PGF_CHECK_ARG(dW && N > 0 && K > 0 && B >= 0, "pgf_linear_bwd_dw: NULL or non-positive argument");
  if (B == 0) {
    if (!accumulate) {
      for (int m = 0; m < n_models; ++m) {
        cudaMemsetAsync(dW + m * sdW, 0, sizeof(float) * N * K, static_cast<cudaStream_t>(stream));
        if (db) cudaMemsetAsync(db + m * sdb, 0, sizeof(float) * N, static_cast<cudaStream_t>(stream));
      }
    }
    return PGF_OK;
  }
  PGF_CHECK_ARG(dY && X, "pgf_linear_bwd_dw: NULL argument");
  PGF_CHECK_ARG((K % 4) == 0 && (ldx % 4) == 0 && aligned16(X) && aligned16(dW) && (sX % 4) == 0 && (sdW % 4) == 0,
                "pgf_linear_bwd_dw: K, ldx, strides must be multiples of 4 and X, dW 16-byte aligned");
  LinDwArgs a;
  a.dY = dY; a.ldy = ldy; a.sdY = sdY; a.X = X; a.ldx = ldx; a.sX = sX; a.dW = dW; a.sdW = sdW; a.db = db; a.sdb = sdb;
  a.B = B; a.N = N; a.K = K; a.rows_per_cta = 0; a.accumulate = accumulate;
  return linear_bwd_dw(a, n_models, static_cast<cudaStream_t>(stream));
}

int pgf_gemm_bf16(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, void* C, long long ldc,
                  int M, int N, int K, int epi, const float* bias, void* aux, long long ld_aux, int stream_k,
                  float* col_partial, void* stream) {
  PGF_CHECK_ARG(A && B && C, "pgf_gemm_bf16: NULL operand");
  PGF_CHECK_ARG((epi >= 0 && epi <= 7 && epi != 3) || epi == 9, "pgf_gemm_bf16: bad epilogue %d", epi);
  if (epi == PGF_EPI_BIAS_RELU_BF16 || epi == PGF_EPI_BIAS_TANH_BF16 || epi == PGF_EPI_BIAS_F32 || epi == PGF_EPI_BIAS_TANH_F32)
    PGF_CHECK_ARG(bias && aligned16(bias), "pgf_gemm_bf16: epilogue needs a 16-byte aligned bias");
  GemmArgs g = {};
  g.M = M; g.N = N; g.K = K; g.C = C; g.ldc = ldc; g.bias = bias; g.aux = aux; g.ld_aux = ld_aux; g.epi = epi;
  g.stream_k = stream_k; g.col_partial = col_partial;
  return gemm_bf16(A, lda, a_mn, B, ldb, b_mn, g, static_cast<cudaStream_t>(stream));
}

int pgf_gemm_bf16x3(const void* A3, long long lda, long long a_plane, int a_mn, const void* B3, long long ldb, long long b_plane,
                    int b_mn, float* C, long long ldc, int M, int N, int K, int epi, const float* bias, int k_slabs,
                    void* stream) {
  PGF_CHECK_ARG(A3 && B3 && C, "pgf_gemm_bf16x3: NULL operand");
  PGF_CHECK_ARG(epi == PGF_EPI_ATOMIC_F32 || epi == PGF_EPI_STORE_F32 || epi == PGF_EPI_BIAS_F32,
                "pgf_gemm_bf16x3: epilogue %d is not an exact fp32 one (4, 5 or 6)", epi);
  PGF_CHECK_ARG(epi != PGF_EPI_BIAS_F32 || (bias && aligned16(bias)), "pgf_gemm_bf16x3: epilogue 6 needs a 16-byte aligned bias");
  PGF_CHECK_ARG((k_slabs > 0) == (epi == PGF_EPI_ATOMIC_F32), "pgf_gemm_bf16x3: k_slabs goes with the accumulating epilogue (4) only");
  GemmArgs g = {};
  g.M = M; g.N = N; g.K = K; g.C = C; g.ldc = ldc; g.bias = bias; g.epi = epi; g.stream_k = k_slabs;
  // plane pairs, smallest products first (0 = hi, 1 = mid, 2 = lo): (lo,hi) (hi,lo) (mid,mid) (mid,hi) (hi,mid) (hi,hi)
  g.nseg = 6;
  g.seg_a = 2u | (0u << 2) | (1u << 4) | (1u << 6) | (0u << 8) | (0u << 10);
  g.seg_b = 0u | (2u << 2) | (1u << 4) | (0u << 6) | (1u << 8) | (0u << 10);
  g.a_plane = a_plane; g.b_plane = b_plane;
  return gemm_bf16(A3, lda, a_mn, B3, ldb, b_mn, g, static_cast<cudaStream_t>(stream));
}

int pgf_split3(const float* src, long long ld, int R, int C, const float* bias, int act, const void* mask_plane,
               long long ld_mask, float* out_f32, long long ld_out, void* planes, long long ldp, long long plane_stride,
               void* stream) {
  if (R <= 0 || C <= 0) return PGF_OK;
  PGF_CHECK_ARG(src && (out_f32 || planes), "pgf_split3: NULL argument");
  PGF_CHECK_ARG(act >= 0 && act <= 2, "pgf_split3: bad activation %d", act);
  PGF_CHECK_ARG((C % 8) == 0 && (ld % 4) == 0 && aligned16(src) && (!bias || aligned16(bias)) &&
                    (!out_f32 || (aligned16(out_f32) && (ld_out % 4) == 0)) &&
                    (!planes || (aligned16(planes) && (ldp % 8) == 0 && (plane_stride % 8) == 0)) &&
                    (!mask_plane || (aligned16(mask_plane) && (ld_mask % 8) == 0)),
                "pgf_split3: C must be a multiple of 8, rows 16-byte aligned");
  return split3(src, ld, R, C, bias, act, mask_plane, ld_mask, out_f32, ld_out, planes, ldp, plane_stride,
                static_cast<cudaStream_t>(stream));
}

int pgf_gemm_partial_rows(int M) { return gemm_partial_rows(M); }

int pgf_set_sm_reserve(int n_sms) {
  set_sm_reserve(n_sms);
  return PGF_OK;
}

int pgf_reduce_partials(const float* partial, int rows, int N, const float* coef, float* out, int accumulate, void* stream) {
  PGF_CHECK_ARG(partial && out && rows > 0 && N > 0, "pgf_reduce_partials: bad argument");
  return reduce_partials(partial, rows, N, coef, out, accumulate, static_cast<cudaStream_t>(stream));
}

int pgf_gemm_bf16_ddp(const void* A, long long lda, const void* B, long long ldb, int b_mn, int M, int N, int K,
                      unsigned long long seed, unsigned int offset, unsigned long long row0, const float* deps_dDP,
                      float* workspace, size_t workspace_bytes, float* dDP, int accumulate, void* stream) {
  PGF_CHECK_ARG(A && B && deps_dDP && dDP && workspace, "pgf_gemm_bf16_ddp: NULL argument");
  PGF_CHECK_ARG(M > 0 && N > 0 && K > 0, "pgf_gemm_bf16_ddp: empty problem");
  const int rows = gemm_partial_rows(M);
  PGF_CHECK_ARG(workspace_bytes >= static_cast<size_t>(rows) * N * sizeof(float), "pgf_gemm_bf16_ddp: workspace too small");
  GemmArgs g = {};
  g.M = M; g.N = N; g.K = K; g.C = workspace; g.ldc = N; g.epi = 8 /* PGF_EPI_DDP_PARTIAL */;
  g.col_partial = workspace; g.seed = seed; g.offset = offset; g.row0 = row0;
  const int rc = gemm_bf16(A, lda, 0, B, ldb, b_mn, g, static_cast<cudaStream_t>(stream));
  if (rc != PGF_OK) return rc;
  return reduce_partials(workspace, rows, N, deps_dDP, dDP, accumulate, static_cast<cudaStream_t>(stream));
}

size_t pgf_cls_ce_workspace(int B, int H, int n_models) {
  if (B <= 0 || n_models <= 0) return 0;
  return cls_ce_workspace(B, H, n_models);
}

int pgf_cls_ce(const void* h, int h_dtype, long long ldh, long long sh, const float* Wc, long long sWc, const float* bc,
               long long sbc, const long long* labels, long long slabels, int B, int H, int n_models, float loss_scale,
               float grad_scale, int backward, int through_tanh, float* logits, long long slogits, long long* pred,
               long long spred, float* stats, void* dz, int dz_dtype, long long lddz, long long sdz, float* dWc,
               long long sdWc, float* dbc, long long sdbc, float* dz_colsum, long long sdz_colsum, float* workspace,
               size_t workspace_bytes, void* stream) {
  if (n_models == 0) return PGF_OK;
  PGF_CHECK_ARG(B > 0, "pgf_cls_ce: empty batch (the reference divides by B: cross_entropy mean over 0 rows is NaN)");
  PGF_CHECK_ARG(h && Wc && bc && workspace, "pgf_cls_ce: NULL argument");
  PGF_CHECK_ARG(labels || !backward, "pgf_cls_ce: backward needs labels");
  PGF_CHECK_ARG(aligned16(h) && aligned16(Wc) && (ldh % 4) == 0 && (sWc % 4) == 0 && (sh % 4) == 0,
                "pgf_cls_ce: h, Wc must be 16-byte aligned");
  if (backward && dz) PGF_CHECK_ARG(aligned16(dz) && (lddz % 4) == 0 && (sdz % 4) == 0, "pgf_cls_ce: dz must be 16-byte aligned");
  CeArgs a = {};
  a.h = h; a.ldh = ldh; a.sh = sh; a.Wc = Wc; a.sWc = sWc; a.bc = bc; a.sbc = sbc; a.labels = labels; a.slab = slabels;
  a.logits = logits; a.slogits = slogits; a.pred = pred; a.spred = spred; a.dz = dz; a.lddz = lddz; a.sdz = sdz;
  a.partial = workspace; a.B = B; a.H = H; a.grad_scale = grad_scale; a.through_tanh = through_tanh;
  PGF_CHECK_ARG(!dz_colsum || (backward && dz), "pgf_cls_ce: dz_colsum needs backward and dz");
  return cls_ce(a, h_dtype, dz_dtype, backward, n_models, loss_scale, stats, dWc, sdWc, dbc, sdbc, dz_colsum, sdz_colsum,
                workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int pgf_adam_step(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n, int step, float lr,
                  float beta1, float beta2, float eps, float grad_scale, void* stream) {
  PGF_CHECK_ARG(n >= 0 && step >= 1, "pgf_adam_step: n < 0 or step < 1");
  if (n == 0) return PGF_OK;
  PGF_CHECK_ARG(p && g && m && v, "pgf_adam_step: NULL argument");
  return adam_step(p, g, m, v, bf16_shadow, n, step, lr, beta1, beta2, eps, grad_scale, static_cast<cudaStream_t>(stream));
}

int pgf_adam_step_strided(float* p, const float* g, float* m, float* v, void* bf16_shadow, long long n, long long model_stride,
                          int n_models, int step, float lr, float beta1, float beta2, float eps, float grad_scale,
                          void* stream) {
  PGF_CHECK_ARG(n >= 0 && step >= 1 && n_models >= 0, "pgf_adam_step_strided: bad n / step / n_models");
  if (n == 0 || n_models == 0) return PGF_OK;
  PGF_CHECK_ARG(p && g && m && v && (model_stride % 4) == 0 && ((n % 4) == 0 || n_models == 1),
                "pgf_adam_step_strided: NULL argument, or segment length / stride not a multiple of 4");
  return adam_step(p, g, m, v, bf16_shadow, n, step, lr, beta1, beta2, eps, grad_scale, static_cast<cudaStream_t>(stream),
                   model_stride, n_models);
}

int pgf_linear_adam_step(const float* dY, long long ldy, long long sdY, const float* X, long long ldx, long long sX, int B,
                         int N, int K, float* W, float* mW, float* vW, float* bias, float* mb, float* vb, long long sP,
                         int step, float lr, float beta1, float beta2, float eps, float grad_scale, int n_models,
                         void* stream) {
  if (n_models == 0 || N == 0 || K == 0) return PGF_OK;
  PGF_CHECK_ARG(dY && X && W && mW && vW && B > 0 && N > 0 && K > 0 && step >= 1, "pgf_linear_adam_step: bad argument");
  PGF_CHECK_ARG((K % 4) == 0 && (ldx % 4) == 0 && (sX % 4) == 0 && (sP % 4) == 0 && aligned16(X) && aligned16(W) &&
                    aligned16(mW) && aligned16(vW),
                "pgf_linear_adam_step: K, ldx, strides must be multiples of 4 and X, W, mW, vW 16-byte aligned");
  PGF_CHECK_ARG(!bias || (mb && vb), "pgf_linear_adam_step: bias needs its moment buffers");
  LinAdamArgs a = {};
  LinAdamLayer& l = a.l[0];
  l.dY = dY; l.ldy = ldy; l.sdY = sdY; l.X = X; l.ldx = ldx; l.sX = sX; l.W = W; l.mW = mW; l.vW = vW;
  l.bias = bias; l.mb = mb; l.vb = vb; l.N = N; l.K = K;
  a.n_layers = 1; a.sP = sP; a.B = B; a.st = nullptr; a.adv = StepAdvance{};
  a.c = make_adam_coef(step, lr, beta1, beta2, eps, grad_scale);
  if (wide_regime(B, n_models)) return linear_adam_step_wide(a, n_models, static_cast<cudaStream_t>(stream));
  return linear_adam_step(a, n_models, static_cast<cudaStream_t>(stream));
}

int pgf_memcpy_peer_async(void* dst, int dst_device, const void* src, int src_device, size_t nbytes, void* stream) {
  if (nbytes == 0) return PGF_OK;
  PGF_CHECK_ARG(dst && src, "pgf_memcpy_peer_async: NULL pointer");
  if (dst_device != src_device) {
    int cur = -1, can = 0;
    PGF_CUDA_CALL(cudaGetDevice(&cur));
    const int other = cur == src_device ? dst_device : src_device;
    PGF_CUDA_CALL(cudaDeviceCanAccessPeer(&can, cur, other));
    PGF_CHECK_ARG(can, "pgf_memcpy_peer_async: device %d cannot access device %d directly", cur, other);
    const cudaError_t e = cudaDeviceEnablePeerAccess(other, 0);
    if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) {
      set_error("pgf_memcpy_peer_async: cudaDeviceEnablePeerAccess(%d) failed: %s", other, cudaGetErrorString(e));
      return PGF_ERR_CUDA;
    }
    (void)cudaGetLastError();   // clear the sticky 'already enabled'
  }
  PGF_CUDA_CALL(cudaMemcpyPeerAsync(dst, dst_device, src, src_device, nbytes, static_cast<cudaStream_t>(stream)));
  return PGF_OK;
}

int pgf_fill_zero(void* p, size_t nbytes, void* stream) {
  if (nbytes == 0) return PGF_OK;
  PGF_CHECK_ARG(p, "pgf_fill_zero: NULL pointer");
  PGF_CUDA_CALL(cudaMemsetAsync(p, 0, nbytes, static_cast<cudaStream_t>(stream)));
  return PGF_OK;
}

int pgf_cast_f32_to_bf16(const float* src, void* dst, long long n, void* stream) {
  if (n == 0) return PGF_OK;
  PGF_CHECK_ARG(src && dst && n > 0 && aligned16(src) && (reinterpret_cast<uintptr_t>(dst) & 7) == 0, "pgf_cast_f32_to_bf16: bad argument");
  return cast_f32_to_bf16(src, dst, n, static_cast<cudaStream_t>(stream));
}

size_t pgf_colsum_workspace(int B, int N) {
  if (B <= 0 || N <= 0) return 0;
  return static_cast<size_t>(colsum_slabs(B, N)) * N * sizeof(float);
}

int pgf_colsum(const void* x, int dtype, long long ld, int B, int N, float* out, float* workspace, size_t workspace_bytes,
               void* stream) {
  PGF_CHECK_ARG(x && out && workspace && B > 0 && N > 0 && (N % 4) == 0, "pgf_colsum: bad argument");
  return colsum(x, dtype, ld, B, N, out, workspace, workspace_bytes, static_cast<cudaStream_t>(stream));
}

int pgf_prigumbel_coef(const float* w, const float* gumbel, int H, float exp_eps, float tau, int hard,
                       unsigned long long seed, unsigned int offset, float* coef, float* wloss, void* stream) {
  PGF_CHECK_ARG(w && coef && wloss && H > 0 && (H % 4) == 0 && aligned16(coef), "pgf_prigumbel_coef: NULL argument or H not a multiple of 4");
  PGF_CHECK_ARG(tau > 0.f, "pgf_prigumbel_coef: tau must be positive");
  return prigumbel_coef(w, gumbel, H, exp_eps, tau, hard, seed, offset, coef, wloss, static_cast<cudaStream_t>(stream));
}

int pgf_prigumbel_fwd(const float* z, long long ldz, const float* coef, const float* lap, float eps,
                      unsigned long long seed, unsigned int offset, unsigned long long row0, float* out, long long ld_out,
                      float* row_min, float* row_max, int B, int H, void* stream) {
  if (B == 0) return PGF_OK;
  PGF_CHECK_ARG(z && coef && out && B > 0 && H > 0, "pgf_prigumbel_fwd: NULL or non-positive argument");
  PGF_CHECK_ARG((H % 4) == 0 && (ldz % 4) == 0 && (ld_out % 4) == 0 && aligned16(z) && aligned16(coef) && aligned16(out),
                "pgf_prigumbel_fwd: H and leading dimensions must be multiples of 4, pointers 16-byte aligned");
  PGF_CHECK_ARG(lap != nullptr || eps > 0.f, "pgf_prigumbel_fwd: eps must be positive in Philox mode");
  PriGumbelArgs a;
  a.z = z; a.ldz = ldz; a.coef = coef; a.lap = lap; a.inv_eps = lap ? 0.f : 1.0f / eps;
  a.k0 = static_cast<unsigned int>(seed); a.k1 = static_cast<unsigned int>(seed >> 32); a.offset = offset; a.row0 = row0;
  a.out = out; a.ld_out = ld_out; a.row_min = row_min; a.row_max = row_max; a.B = B; a.H = H;
  return prigumbel_fwd(a, static_cast<cudaStream_t>(stream));
}

size_t pgf_prigumbel_bwd_workspace(int B, int H) {
  if (B <= 0 || H <= 0) return 0;
  return static_cast<size_t>(prigumbel_bwd_slabs(B)) * H * sizeof(float);
}

int pgf_prigumbel_bwd(const float* z, long long ldz, const float* coef, const float* dout, long long ld_dout,
                      const float* wloss, float exp_eps, float wloss_scale, float* dz, long long ld_dz, float* dw,
                      int accumulate, int B, int H, float* workspace, size_t workspace_bytes, void* stream) {
  if (B == 0) return PGF_OK;
  PGF_CHECK_ARG(z && coef && dout && dz && workspace && B > 0 && H > 0, "pgf_prigumbel_bwd: NULL or non-positive argument");
  PGF_CHECK_ARG(dw == nullptr || wloss != nullptr, "pgf_prigumbel_bwd: dw needs wloss from pgf_prigumbel_coef");
  PGF_CHECK_ARG((H % 4) == 0 && (ldz % 4) == 0 && (ld_dout % 4) == 0 && (ld_dz % 4) == 0 && aligned16(z) && aligned16(coef) &&
                    aligned16(dout) && aligned16(dz),
                "pgf_prigumbel_bwd: H and leading dimensions must be multiples of 4, pointers 16-byte aligned");
  if (workspace_bytes < pgf_prigumbel_bwd_workspace(B, H)) {
    set_error("pgf_prigumbel_bwd: workspace of %zu bytes, need %zu", workspace_bytes, pgf_prigumbel_bwd_workspace(B, H));
    return PGF_ERR_WORKSPACE;
  }
  PriGumbelBwdArgs a;
  a.z = z; a.ldz = ldz; a.coef = coef; a.dout = dout; a.ld_dout = ld_dout; a.dz = dz; a.ld_dz = ld_dz; a.partial = workspace;
  a.B = B; a.H = H;
  return prigumbel_bwd(a, exp_eps, wloss_scale, wloss, dw, accumulate, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
