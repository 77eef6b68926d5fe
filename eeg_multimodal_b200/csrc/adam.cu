// Fused Adam for the two parameter groups of the reference's step (past_acc.py:155-160:
// `Adam(DP_params, lr)` and `Adam(model_params, lr)`, torch defaults betas=(.9,.999), eps=1e-8,
// no weight decay, no amsgrad), over a flat fp32 parameter buffer, grouped over models.
// Mirrors torch.optim.Adam's single-tensor op order:
//   m.lerp_(g, 1-b1); v.mul_(b2).addcmul_(g, g, 1-b2);
//   denom = sqrt(v)/sqrt(1-b2^t) + eps;  p.addcdiv_(m, denom, -lr/(1-b1^t))
// Optionally refreshes a bf16 shadow copy of the parameters (tensor-core GEMM operands).
// HBM traffic: read 16 B/param (p,g,m,v), write 12 B/param (+2 with the bf16 shadow).
#include "pgf_kernels.cuh"

namespace pgf {

__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                   float* __restrict__ m, float* __restrict__ v,
                                                   __nv_bfloat16* __restrict__ shadow, long long n, const AdamCoef c,
                                                   long long model_stride) {
  // blockIdx.y walks the models of a strided group (segment [0,n) of every model's flat buffer)
  const long long base = static_cast<long long>(blockIdx.y) * model_stride;
  p += base; g += base; m += base; v += base;
  if (shadow) shadow += base;
  const long long n4 = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 P = reinterpret_cast<float4*>(p)[i];
    const float4 G = reinterpret_cast<const float4*>(g)[i];
    float4 M = reinterpret_cast<float4*>(m)[i];
    float4 V = reinterpret_cast<float4*>(v)[i];
    float pe[4] = {P.x, P.y, P.z, P.w}, ge[4] = {G.x, G.y, G.z, G.w}, me[4] = {M.x, M.y, M.z, M.w},
          ve[4] = {V.x, V.y, V.z, V.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) adam_update(pe[e], me[e], ve[e], ge[e], c);
    reinterpret_cast<float4*>(p)[i] = make_float4(pe[0], pe[1], pe[2], pe[3]);
    reinterpret_cast<float4*>(m)[i] = make_float4(me[0], me[1], me[2], me[3]);
    reinterpret_cast<float4*>(v)[i] = make_float4(ve[0], ve[1], ve[2], ve[3]);
    if (shadow) {
      uint2 u;
      u.x = pack_bf16x2(pe[0], pe[1]);
      u.y = pack_bf16x2(pe[2], pe[3]);
      reinterpret_cast<uint2*>(shadow)[i] = u;
    }
  }
  // tail (n % 4 elements), handled by the first threads of block 0
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) {
    const long long i = (n4 << 2) + threadIdx.x;
    float pp = p[i], mm = m[i], vv = v[i];
    adam_update(pp, mm, vv, g[i], c);
    p[i] = pp; m[i] = mm; v[i] = vv;
    if (shadow) shadow[i] = __float2bfloat16_rn(pp);
  }
}

int adam_step(float* p, const float* g, float* m, float* v, void* shadow, long long n, int step, float lr, float b1,
              float b2, float eps, float grad_scale, cudaStream_t s, long long model_stride, int n_models) {
  if (n <= 0 || n_models <= 0) return PGF_OK;
  if ((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
       reinterpret_cast<uintptr_t>(v)) & 15) {
    set_error("pgf_adam_step: buffers must be 16-byte aligned");
    return PGF_ERR_ARG;
  }
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = (8LL * num_sms() + n_models - 1) / n_models;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<dim3(static_cast<unsigned>(blocks), n_models), 256, 0, s>>>(p, g, m, v, static_cast<__nv_bfloat16*>(shadow), n,
                                                                            make_adam_coef(step, lr, b1, b2, eps, grad_scale),
                                                                            model_stride);
  PGF_CUDA_LAUNCH_CHECK("pgf_adam_step");
  return PGF_OK;
}

// ---------------------------------------------------------------------------------------------
// small layout helpers for the tensor-core path
// ---------------------------------------------------------------------------------------------
__global__ void cast_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, long long n) {
  const long long n4 = n >> 2;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(src)[i];
    uint2 u;
    u.x = pack_bf16x2(x.x, x.y);
    u.y = pack_bf16x2(x.z, x.w);
    reinterpret_cast<uint2*>(dst)[i] = u;
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 3)) dst[(n4 << 2) + threadIdx.x] = __float2bfloat16_rn(src[(n4 << 2) + threadIdx.x]);
}

int cast_f32_to_bf16(const float* src, void* dst, long long n, cudaStream_t s) {
  long long blocks = (n / 4 + 255) / 256;
  const long long cap = 8LL * num_sms();
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  cast_bf16_kernel<<<static_cast<unsigned>(blocks), 256, 0, s>>>(src, static_cast<__nv_bfloat16*>(dst), n);
  PGF_CUDA_LAUNCH_CHECK("pgf_cast_f32_to_bf16");
  return PGF_OK;
}

// column sums of a [B,N] bf16/fp32 matrix (bias gradients of the tensor-core path):
// stage 1 partial[slab][N], stage 2 deterministic sum.
template <typename T>
__global__ void __launch_bounds__(128) colsum_kernel(const T* __restrict__ x, long long ld, int B, int N, int rows_per_slab,
                                                     float* __restrict__ partial) {
  const int c = (blockIdx.x * blockDim.x + threadIdx.x) * 4;
  if (c >= N) return;
  const int r0 = blockIdx.y * rows_per_slab, r1 = min(B, r0 + rows_per_slab);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int r = r0; r < r1; ++r) {
    float4 v;
    if (sizeof(T) == 4) {
      v = ldg_stream(reinterpret_cast<const float4*>(reinterpret_cast<const float*>(x) + static_cast<long long>(r) * ld + c));
    } else {
      const uint2 u = ldg_stream_u2(reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(x) + static_cast<long long>(r) * ld + c));
      const float2 a = unpack_bf16x2(u.x), b = unpack_bf16x2(u.y);
      v = make_float4(a.x, a.y, b.x, b.y);
    }
    acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
  }
  *reinterpret_cast<float4*>(partial + static_cast<long long>(blockIdx.y) * N + c) = acc;
}

__global__ void colsum_finalize_kernel(const float* __restrict__ partial, int nslab, int N, float* __restrict__ out) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= N) return;
  float s = 0.f;
  for (int i = 0; i < nslab; ++i) s += partial[static_cast<long long>(i) * N + c];
  out[c] = s;
}

int colsum_slabs(int B, int N) {
  const int cctas = (N / 4 + 127) / 128;
  int slabs = (4 * num_sms() + cctas - 1) / cctas;
  const int max_slabs = (B + 63) / 64;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  return slabs;
}

int colsum(const void* x, int dtype, long long ld, int B, int N, float* out, float* workspace, size_t workspace_bytes,
           cudaStream_t s) {
  const int slabs = colsum_slabs(B, N);
  if (workspace_bytes < static_cast<size_t>(slabs) * N * sizeof(float)) {
    set_error("pgf_colsum: workspace too small");
    return PGF_ERR_WORKSPACE;
  }
  const int rows = (B + slabs - 1) / slabs;
  const dim3 grid((N / 4 + 127) / 128, slabs);
  if (dtype == PGF_DT_F32)
    colsum_kernel<float><<<grid, 128, 0, s>>>(static_cast<const float*>(x), ld, B, N, rows, workspace);
  else
    colsum_kernel<__nv_bfloat16><<<grid, 128, 0, s>>>(static_cast<const __nv_bfloat16*>(x), ld, B, N, rows, workspace);
  PGF_CUDA_LAUNCH_CHECK("pgf_colsum");
  colsum_finalize_kernel<<<(N + 255) / 256, 256, 0, s>>>(workspace, slabs, N, out);
  PGF_CUDA_LAUNCH_CHECK("pgf_colsum(finalize)");
  return PGF_OK;
}

}  // namespace pgf
