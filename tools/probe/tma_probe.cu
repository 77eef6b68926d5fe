// Pure streaming-delivery probe for the TMA-fed ring of csrc/linear_stream.cu: how fast can 148 persistent CTAs pull an
// fp32 [rows][K] matrix (>> L2) through an 8-stage x 16 KB shared-memory ring, depending on how a stage is requested?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/probe/tma_probe tools/probe/tma_probe.cu -lcuda
//   ./tools/probe/tma_probe            (prints GB/s per variant)
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void bar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(u32(b)), "r"(c)); }
__device__ __forceinline__ void bar_wait(uint64_t* b, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(u32(b)), "r"(parity) : "memory");
  } while (!ok);
}
__device__ __forceinline__ void bar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(u32(b)) : "memory"); }
__device__ __forceinline__ void bar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(u32(b)), "r"(bytes) : "memory"); }

constexpr int STAGE_BYTES = 16384;
constexpr int WARPS = 8;

// mode 0/1: 2-D tensor map, box {bw floats, bh rows} (bw*bh*4 = 16 KB); mode 2: 1-D bulk copies of `piece` bytes, 16 KB / piece per stage
__global__ void __launch_bounds__(288, 1) probe(const __grid_constant__ CUtensorMap tm, const float* W, long long rows, int K, int mode,
                                                int bw, int bh, int piece, int stages, float* sink) {
  extern __shared__ __align__(1024) unsigned char smem[];
  float* ring = (float*)smem;
  uint64_t* full = (uint64_t*)(smem + (size_t)stages * STAGE_BYTES);
  uint64_t* empty = full + stages;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) { bar_init(full + s, 1); bar_init(empty + s, WARPS); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  __syncthreads();
  // work: the matrix as a sequence of stage-sized units; contiguous unit range per CTA
  const int kpieces = mode == 2 ? (K * 4 + piece - 1) / piece : K / bw;     // pieces along a row
  const int rows_per_stage = mode == 2 ? STAGE_BYTES / piece : bh;
  const long long row_groups = rows / rows_per_stage;
  const long long units = row_groups * kpieces;
  const long long lo = units * blockIdx.x / gridDim.x, hi = units * (blockIdx.x + 1) / gridDim.x;
  if (warp == WARPS) {
    int stage = 0; uint32_t phase = 0;
    for (long long u = lo; u < hi; ++u) {
      const long long rg = u / kpieces; const int kp = (int)(u % kpieces);
      if (lane == 0) { bar_wait(empty + stage, phase ^ 1u); bar_expect(full + stage, mode == 2 ? (uint32_t)(rows_per_stage * piece) : (uint32_t)STAGE_BYTES); }
      __syncwarp();
      float* dst = ring + (size_t)stage * (STAGE_BYTES / 4);
      if (mode != 2) {
        if (lane == 0)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(u32(dst)),
                       "l"(&tm), "r"(kp * bw), "r"((int)(rg * bh)), "r"(u32(full + stage)) : "memory");
      } else {
        for (int r = lane; r < rows_per_stage; r += 32)
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(u32(dst + (size_t)r * (piece / 4))),
                       "l"(W + (rg * rows_per_stage + r) * K + (long long)kp * (piece / 4)), "r"(piece), "r"(u32(full + stage)) : "memory");
      }
      if (++stage == stages) { stage = 0; phase ^= 1u; }
    }
    return;
  }
  int stage = 0; uint32_t phase = 0;
  float acc = 0.f;
  for (long long u = lo; u < hi; ++u) {
    bar_wait(full + stage, phase);
    acc += ring[(size_t)stage * (STAGE_BYTES / 4) + threadIdx.x];     // touch the stage
    __syncwarp();
    if (lane == 0) bar_arrive(empty + stage);
    if (++stage == stages) { stage = 0; phase ^= 1u; }
  }
  if (acc == 123.456f) sink[0] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const int K = 2304;
  const long long rows = 6LL * 2304 * 8;        // 8 x the six-model W1 set = 1.02 GB
  float* W; float* sink;
  CK(cudaMalloc(&W, rows * K * 4)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(W, 0, rows * K * 4));
  void* fp = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fp, cudaEnableDefault, &q));
  EncodeFn enc = (EncodeFn)fp;
  struct V { const char* name; int mode, bw, bh, piece, stages; CUtensorMapL2promotion l2; } vs[] = {
      {"tensor map box 128 x 32 rows, L2 promo 256B", 0, 128, 32, 0, 8, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"tensor map box 128 x 32 rows, L2 promo 128B", 0, 128, 32, 0, 8, CU_TENSOR_MAP_L2_PROMOTION_L2_128B},
      {"tensor map box 128 x 32 rows, no promo", 0, 128, 32, 0, 8, CU_TENSOR_MAP_L2_PROMOTION_NONE},
      {"tensor map box 256 x 16 rows", 1, 256, 16, 0, 8, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"tensor map box 128 x 32 rows, 12 stages", 0, 128, 32, 0, 12, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"tensor map box 128 x 32 rows, 4 stages", 0, 128, 32, 0, 4, CU_TENSOR_MAP_L2_PROMOTION_L2_256B},
      {"1-D bulk 512 B x 32 rows", 2, 0, 0, 512, 8, CU_TENSOR_MAP_L2_PROMOTION_NONE},
      {"1-D bulk 2304 B x ~7 rows (quarter rows)", 2, 0, 0, 2304, 8, CU_TENSOR_MAP_L2_PROMOTION_NONE},
      {"1-D bulk 1024 B x 16 rows", 2, 0, 0, 1024, 8, CU_TENSOR_MAP_L2_PROMOTION_NONE},
      {"1-D bulk 512 B x 32 rows, 12 stages", 2, 0, 0, 512, 12, CU_TENSOR_MAP_L2_PROMOTION_NONE},
  };
  for (auto& v : vs) {
    CUtensorMap tm;
    if (v.mode != 2) {
      cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows}; cuuint64_t strides[1] = {(cuuint64_t)K * 4};
      cuuint32_t box[2] = {(cuuint32_t)v.bw, (cuuint32_t)v.bh}; cuuint32_t es[2] = {1, 1};
      CUresult r = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, W, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       v.l2, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) { printf("%s: encode failed %d\n", v.name, (int)r); continue; }
    }
    int piece = v.piece;
    if (v.mode == 2 && STAGE_BYTES % piece) piece = v.piece;   // 2304 B pieces: 7 rows x 2304 = 16128 B (stage not full); account below
    const size_t smem = (size_t)v.stages * STAGE_BYTES + 2 * v.stages * 8;
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    // stage accounting for non-dividing pieces
    float best = 1e9f;
    for (int it = 0; it < 4; ++it) {
      CK(cudaEventRecord(e0));
      probe<<<148, 288, smem>>>(tm, W, rows, K, v.mode, v.bw, v.bh, piece, v.stages, sink);
      CK(cudaEventRecord(e1));
      CK(cudaEventSynchronize(e1));
      float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
      if (it > 0 && ms < best) best = ms;
    }
    CK(cudaGetLastError());
    double bytes = (double)rows * K * 4;
    if (v.mode == 2) {   // bytes actually requested
      const int rps = STAGE_BYTES / piece; const int kp = (K * 4 + piece - 1) / piece;
      bytes = (double)(rows / rps) * kp * rps * piece;
    }
    printf("%-52s %7.3f ms  %7.0f GB/s\n", v.name, best, bytes / best / 1e6);
  }
  return 0;
}
