// Kernel (b), large-batch regime: bf16 x bf16 -> fp32 GEMM on the 5th-gen tensor cores
// (tcgen05.mma, accumulators in TMEM, operands staged by TMA into 128B-swizzled shared memory),
// with the fusion-MLP epilogues fused in.  Hand-written PTX, no CUTLASS.
//
// Reference ops replaced: the cuBLAS SGEMMs behind nn.Linear of `fc_layers` and their autograd
// (models.py:46-51,80; past_acc.py:87-92,137) when the batch is large enough to be a real dense
// contraction (SURVEY.md section 8d "Roofline for (b) - tensor pipe").
//
//   C[M,N] = A[M,K] . B[N,K]^T        A, B bf16; fp32 accumulate
// Either operand may be "MN-major" (stored [K,M] / [K,N], i.e. the transposed matrix row-major),
// which is what the weight-gradient GEMMs need (dW = dZ^T . X contracts over the batch) -- the
// UMMA descriptors read those layouts directly, no transposed copies are ever materialised.
//
// Structure (persistent; one CTA per SM, CTA pairs = clusters of 2 whenever M > 128; 320 threads per CTA):
//   warp 0     TMA producer: 5-stage ring of {A 128x64, B 128x64 (its half of the pair's 256 columns)} bf16 tiles
//   warp 1     TMEM allocator + single-thread tcgen05.mma issuer (cta_group::2: 256x256x16 per instruction, leader CTA)
//   warps 2-9  epilogue, two groups of four warps on alternate column groups: tcgen05.ld -> bias / activation / mask
//              -> 128B-swizzled shared staging -> TMA store (or TMA fp32 reduce-add)
// TMEM holds two 128x256 fp32 accumulators per CTA (all 512 columns) so the epilogue of tile i overlaps
// the MMAs of tile i+1.  Work is either whole tiles round-robin, or -- for the K=batch
// weight-gradient GEMMs whose tile count does not fill 148 SMs -- a split over K slabs (slab-major units),
// partial tiles combined by TMA fp32 reduce-adds into a zero-initialised C.
//
// fp32-parity mode (nseg = 6, pgf_gemm_bf16x3): each fp32 operand is given as three bf16 planes hi/mid/lo with
// hi+mid+lo == x exactly (pgf_split3); the contraction runs over six plane pairs
//   (lo,hi) (hi,lo) (mid,mid) (mid,hi) (hi,mid) (hi,hi)      -- smallest terms first
// as six K segments of ONE accumulation, i.e. every product the fp32 result needs down to 2^-24 relative.  The
// producer warp walks the segments by moving the plane coordinate of 3-D tensor maps; MMA issuer and epilogue
// only see six times as many k-blocks.
#include <cuda.h>
#include <stdlib.h>

#include "pgf_kernels.cuh"

namespace pgf {

#define PGF_EPI_STORE_BF16 0        // C = acc
#define PGF_EPI_BIAS_RELU_BF16 1    // C = relu(acc + bias[n])
#define PGF_EPI_BIAS_TANH_BF16 2    // C = tanh(acc + bias[n])
#define PGF_EPI_ATOMIC_F32 4        // C(fp32) += acc                    (stream-K partials)
#define PGF_EPI_STORE_F32 5         // C(fp32) = acc
#define PGF_EPI_BIAS_F32 6          // C(fp32) = acc + bias[n]
#define PGF_EPI_BIAS_TANH_F32 7     // C(fp32) = tanh(acc + bias[n])
#define PGF_EPI_DDP_PARTIAL 8       // no C: col_partial[m/128][n] = sum over 128 rows of acc * Laplace(row, n)
#define PGF_EPI_BITMASK_BF16 9      // C = acc * bit(aux, m, n)          (aux uint32 [M, N/32]: ReLU sign bits written by epi 1)

constexpr int BM = 128, BN = 256, BK = 64, UMMA_K = 16;  // BM = accumulator rows per CTA (TMEM lanes)
constexpr int GEMM_THREADS = 320;  // TMA warp, MMA warp, 2 x 4 epilogue warps
// CG = CTAs per MMA (tcgen05 cta_group).  CG=2: a CTA pair (cluster of 2, one TPC) owns a 256x256
// tile; each CTA stages its own 128 A rows and HALF of the B tile, the MMA reads both halves, so
// shared-memory fill traffic per flop drops by a third and the ring gets 6 stages instead of 4.
template <int CG> struct Cfg {
  static constexpr int A_BYTES = BM * BK * 2;            // 16 KB
  static constexpr int B_BYTES = (BN / CG) * BK * 2;     // 32 KB (CG=1) / 16 KB (CG=2)
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = CG == 1 ? 3 : 5;
  static constexpr int EPI_BYTES = 4 * 16384;  // 2 epilogue groups x 2 output staging buffers, [128 rows][128 B] each
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + EPI_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
// Arrive on a barrier living in another CTA of the cluster (address from mapa).
// The accumulator hand-back (epilogue -> MMA warp of the leader CTA) orders TENSOR-memory accesses, which the tcgen05
// fences on both sides do; it publishes no global or shared data.  `release.cluster` lowers to MEMBAR.ALL.GPU + ERRBAR +
// CGAERRBAR, i.e. the arriving lane waits for every global store it has in flight (column partials, sign bits) once per
// tile (11 % of the K=768 kernel's stall samples sat on that ERRBAR); the default form is a bare SYNCS.ARRIVE.
__device__ __forceinline__ void mbar_arrive_cluster_nofence(uint32_t cluster_bar) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_bar) : "memory");
}
__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// TMA tile load; completes `bytes` on mbarrier `bar`.  CG=2: `bar` is the LEADER CTA's barrier
// (shared::cluster address), the data lands in the issuing CTA's own shared memory.
template <int CG>
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  if (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  if (CG == 1) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  } else {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(dst),
        "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
  }
}
template <int CG>
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accum) {
  if (CG == 1) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  } else {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accum)
        : "memory");
  }
}
// make the mbarrier (same offset in every CTA of the MMA group) track completion of all MMAs issued so far
template <int CG>
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  if (CG == 1) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
  } else {
    const uint16_t mask = 3;
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"(mask)
                 : "memory");
  }
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, "
      "%32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, "
      "%48, %49, %50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]),
        "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]), "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]),
        "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]), "=r"(v[46]), "=r"(v[47]), "=r"(v[48]),
        "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]), "=r"(v[55]), "=r"(v[56]),
        "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// UMMA shared-memory descriptor, SWIZZLE_128B, version 1 (sm_100).  Field layout as in the PTX ISA
// "tcgen05 shared memory descriptor": start>>4 [0,14), LBO>>4 [16,30), SBO>>4 [32,46),
// version [46,48), layout_type [61,64) (2 = 128B swizzle).
//   K-major tile  [rows][64 bf16 = 128 B]: 8-row groups 1024 B apart (SBO); LBO unused (1).
//   MN-major tile [64 k][64 mn = 128 B] boxes: 8-k groups 1024 B apart (SBO), next 64-wide
//   MN box 8192 B further (LBO).
template <bool MN_MAJOR>
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr) {
  const uint64_t lbo = MN_MAJOR ? (8192u >> 4) : 1u;
  const uint64_t sbo = 1024u >> 4;
  return static_cast<uint64_t>((saddr & 0x3FFFFu) >> 4) | (lbo << 16) | (sbo << 32) | (1ull << 46) | (2ull << 61);
}

// Column sums across the 32 lanes of a warp for 32 per-lane values (lane = accumulator row, f[i] = column i):
// a 5-step butterfly that halves the number of live values per step (16+8+4+2+1 = 31 shuffles).  Returns, on
// lane l, the sum over the warp's 32 rows of column l.  f is destroyed.
__device__ __forceinline__ float warp_colsum32(float (&f)[32], int lane) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? f[i] : f[i + o];
      const float keep = up ? f[i + o] : f[i];
      f[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return f[0];
}

// tanh for the bf16-input GEMM epilogues: 1 - 2/(2^(2x*log2 e) + 1) with MUFU.EX2 + MUFU.RCP (6 instructions, absolute
// error <= 3e-7, exact saturation at +-1) instead of tanhf's ~25: the epilogue's arithmetic is energy the power-capped
// tensor pipe can use, and its inputs already carry bf16 rounding (2^-9).  The fp32 parity path keeps tanhf.
__device__ __forceinline__ float tanh_fast(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * 2.8853900817779268f));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(e + 1.0f));
  return fmaf(-2.0f, r, 1.0f);
}

struct WorkUnit {
  int tile, kb0, kb1;
};

// Work distribution over `nworkers` MMA groups (CTAs or CTA pairs), round-robin over units.
//   stream_k == 0: a unit is a whole output tile.
//   stream_k == S: K is cut into S slabs and a unit is (slab, tile), slab-major: the groups that run
//   concurrently work on the SAME K slab of both operands, so each slab is read from HBM once and
//   served to the other tiles from L2 (the weight-gradient GEMMs contract over the batch: K = 65,536,
//   operands 335 MB each, far larger than L2).  Partial tiles are combined by fp32 reductions.
struct Scheduler {
  int num_tiles, kb_k, nseg, nslab, tile_major, worker, nworkers, it;
  __device__ Scheduler(const GemmArgs& g, int tile_m, int worker_, int nworkers_) {
    const int mb = (g.M + tile_m - 1) / tile_m, nb = (g.N + BN - 1) / BN;
    num_tiles = mb * nb;
    kb_k = (g.K + BK - 1) / BK;
    nseg = g.nseg > 1 ? g.nseg : 1;
    nslab = g.stream_k > 0 ? g.stream_k : 1;
    tile_major = g.tile_major;
    worker = worker_;
    nworkers = nworkers_;
    it = 0;
  }
  // fp32-parity mode (nseg plane pairs): a slab is a range [a, b) of the ORIGINAL k-blocks and the unit runs all nseg
  // segments over it, kb0 = a * nseg .. kb1 = b * nseg; k-block kb of the unit is plane pair (kb - kb0) / (b - a) at
  // original k-block a + (kb - kb0) % (b - a).  The tensor core truncates when it accumulates, an error that grows with
  // the number of MMAs issued while the accumulator is at full magnitude, i.e. with b - a (the last segment, hi x hi):
  // slabs bound that chain, and their partial tiles are combined by round-to-nearest fp32 adds.
  __device__ bool next(WorkUnit& u) {
    const long long total = static_cast<long long>(num_tiles) * nslab;
    while (true) {
      const long long unit = worker + static_cast<long long>(it) * nworkers;
      ++it;
      if (unit >= total) return false;
      int slab;
      if (tile_major) {  // the slabs of one tile run back to back: its fp32 partial sums meet in L2 (large C)
        u.tile = static_cast<int>(unit / nslab);
        slab = static_cast<int>(unit - static_cast<long long>(u.tile) * nslab);
      } else {           // slab-major: concurrent units share the K slab of both operands (large operands, small C)
        slab = static_cast<int>(unit / num_tiles);
        u.tile = static_cast<int>(unit - static_cast<long long>(slab) * num_tiles);
      }
      u.kb0 = static_cast<int>(static_cast<long long>(kb_k) * slab / nslab) * nseg;
      u.kb1 = static_cast<int>(static_cast<long long>(kb_k) * (slab + 1) / nslab) * nseg;
      if (u.kb1 > u.kb0) return true;  // more slabs than k-blocks: skip the empty ones (all roles agree)
    }
  }
};

// epilogue output: shared -> global through TMA (plain store, or fp32 add-reduction for split-K partials)
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src), "r"(c0),
               "r"(c1)
               : "memory");
}
// the same store with an L2 eviction-priority hint: an output far larger than L2 that the NEXT kernel streams (H1, dZ1:
// 335 MB) only displaces the operand panels this kernel keeps re-reading if it is written with the normal priority
__device__ __forceinline__ void tma_store_2d_hint(const CUtensorMap* map, uint32_t src, int c0, int c1, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3}], [%1], %4;" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1), "l"(policy)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map), "r"(src),
               "r"(c0), "r"(c1)
               : "memory");
}

template <bool A_MN, bool B_MN, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                    const __grid_constant__ CUtensorMap tmC, const GemmArgs g) {
  using C = Cfg<CG>;
  constexpr int STAGES = C::STAGES, STAGE_BYTES = C::STAGE_BYTES, A_STAGE_BYTES = C::A_BYTES;
  constexpr int TILE_M = BM * CG, BN_CTA = BN / CG;
  extern __shared__ uint8_t smem_raw[];
  // 128B-swizzle atoms must be 1024-byte aligned
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_al = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t epi_base = smem_base + STAGES * STAGE_BYTES;  // out staging [2][16 KB], mask staging [2][16 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + STAGES * STAGE_BYTES + C::EPI_BYTES);
  // bars[0..S) full, [S..2S) empty, [2S..2S+2) tmem_full, [2S+2..2S+4) tmem_empty, tmem slot
  const uint32_t full_bar = smem_u32(bars), empty_bar = smem_u32(bars + STAGES);
  const uint32_t tfull_bar = smem_u32(bars + 2 * STAGES), tempty_bar = smem_u32(bars + 2 * STAGES + 2);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nb_n = (g.N + BN - 1) / BN;
  const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;  // rank 0 = leader: issues the MMAs, owns full/tmem_empty
  const int worker = blockIdx.x / CG, nworkers = gridDim.x / CG;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmC) : "memory");
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar + 8 * s, 1);    // the leader's arrive.expect_tx (+ the TMA bytes of every CTA of the group)
      mbar_init(empty_bar + 8 * s, 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(tfull_bar + 8 * s, 1);
      mbar_init(tempty_bar + 8 * s, 8 * CG);  // one arrival per epilogue warp of the MMA group
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 1) {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // =============================== TMA producer (every CTA) ===============================
    if (lane == 0) {
      Scheduler sched(g, TILE_M, worker, nworkers);
      WorkUnit u;
      uint32_t stage = 0, phase = 0;
      while (sched.next(u)) {
        const int m0 = (u.tile / nb_n) * TILE_M + static_cast<int>(rank) * BM;
        const int n0 = (u.tile % nb_n) * BN + static_cast<int>(rank) * BN_CTA;
        for (int kb = u.kb0; kb < u.kb1; ++kb) {
          mbar_wait(empty_bar + 8 * stage, phase ^ 1);
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
          const uint32_t fb_local = full_bar + 8 * stage;
          const uint32_t fb = CG == 2 ? mapa_u32(fb_local, 0) : fb_local;  // the leader's barrier
          // Only the leader arrives (expecting the bytes of the whole group).  The peer needs no arrival of
          // its own: it refills slot s only after ITS empty[s] fired, i.e. after the MMAs that consumed the
          // previous contents retired, which is after the leader's previous full[s] phase completed; so its
          // complete_tx always lands in the phase it belongs to (and a remote arrive.release.cluster per
          // k-block would serialise the peer's producer on a cluster-scope fence).
          if (rank == 0) mbar_expect_tx(fb_local, STAGE_BYTES * CG);
          if (g.nseg > 1) {
            // fp32-parity mode: k-block kb of the unit = original k-block a + j % len of plane pair j / len
            const int len = (u.kb1 - u.kb0) / g.nseg, j = kb - u.kb0;
            const int seg = j / len, k0 = (u.kb0 / g.nseg + (j - seg * len)) * BK;
            const int pa = static_cast<int>((g.seg_a >> (2 * seg)) & 3u), pb = static_cast<int>((g.seg_b >> (2 * seg)) & 3u);
            if (A_MN) {
#pragma unroll
              for (int j = 0; j < BM / 64; ++j) tma_load_3d<CG>(sa + j * 8192, &tmA, fb, m0 + 64 * j, k0, pa);
            } else {
              tma_load_3d<CG>(sa, &tmA, fb, k0, m0, pa);
            }
            if (B_MN) {
#pragma unroll
              for (int j = 0; j < BN_CTA / 64; ++j) tma_load_3d<CG>(sb + j * 8192, &tmB, fb, n0 + 64 * j, k0, pb);
            } else {
              tma_load_3d<CG>(sb, &tmB, fb, k0, n0, pb);
            }
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
            continue;
          }
          const int k0 = kb * BK;
          if (A_MN) {
#pragma unroll
            for (int j = 0; j < BM / 64; ++j) tma_load_2d<CG>(sa + j * 8192, &tmA, fb, m0 + 64 * j, k0);
          } else {
            tma_load_2d<CG>(sa, &tmA, fb, k0, m0);
          }
          if (B_MN) {
#pragma unroll
            for (int j = 0; j < BN_CTA / 64; ++j) tma_load_2d<CG>(sb + j * 8192, &tmB, fb, n0 + 64 * j, k0);
          } else {
            tma_load_2d<CG>(sb, &tmB, fb, k0, n0);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // =============================== MMA issuer (leader CTA, one thread) ===============================
    if (lane == 0 && rank == 0) {
      // instruction descriptor: fp32 accum, bf16 x bf16, majors, N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(A_MN) << 15) |
                             (static_cast<uint32_t>(B_MN) << 16) | (static_cast<uint32_t>(BN >> 3) << 17) |
                             (static_cast<uint32_t>(TILE_M >> 4) << 24);
      Scheduler sched(g, TILE_M, worker, nworkers);
      WorkUnit u;
      uint32_t stage = 0, phase = 0, unit = 0;
      while (sched.next(u)) {
        const uint32_t acc = unit & 1, acc_phase = (unit >> 1) & 1;
        mbar_wait(tempty_bar + 8 * acc, acc_phase ^ 1);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (int kb = u.kb0; kb < u.kb1; ++kb) {
          mbar_wait(full_bar + 8 * stage, phase);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          const uint32_t sa = smem_base + stage * STAGE_BYTES, sb = sa + A_STAGE_BYTES;
          const uint64_t adesc = make_smem_desc<A_MN>(sa), bdesc = make_smem_desc<B_MN>(sb);
#pragma unroll
          for (int k = 0; k < BK / UMMA_K; ++k) {
            // advance along K: K-major +32 B inside the swizzle atom; MN-major +2 k-groups (2 KB)
            const uint64_t ao = static_cast<uint64_t>((A_MN ? 2048 * k : 32 * k) >> 4);
            const uint64_t bo = static_cast<uint64_t>((B_MN ? 2048 * k : 32 * k) >> 4);
            umma_bf16<CG>(tmem_d, adesc + ao, bdesc + bo, idesc, (kb > u.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit<CG>(empty_bar + 8 * stage);  // frees the smem slot (in every CTA of the group) when these MMAs retire
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        umma_commit<CG>(tfull_bar + 8 * acc);  // accumulator ready for the epilogue warps of the group
        ++unit;
      }
    }
  } else {
    // =============================== epilogue (warps 2..9, every CTA) ===============================
    // TMEM -> registers -> (bias / activation / ReLU sign bits) -> 128B-swizzled shared staging -> TMA store
    // (or TMA fp32 add-reduction for split-K partials).  The thread <-> accumulator-row mapping of
    // tcgen05.ld would make direct global stores touch 32 different lines per instruction; staging
    // through shared memory hands the global traffic to the TMA engine in full 128-byte rows.
    // Two epilogue groups (EG) of four warps: a warp may only read the TMEM lane quarter (warp % 4), so each
    // group covers all 128 accumulator rows, and the groups take alternate column groups of the tile --
    // two warps per SM sub-partition keep the short-K GEMMs (whose epilogue outweighs their 12 k-blocks of
    // MMAs) and the Philox-regenerating dDP epilogue off the critical path.  Each group owns two staging
    // buffers, one named barrier and one TMA-store leader; one barrier per column group (the leader waits
    // for the previous store's shared-memory reads before arriving, which frees the other buffer).
    const int eg = (warp - 2) >> 2;
    const int quarter = warp & 3;  // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const bool leader_thread = ((warp - 2) & 3) == 0 && lane == 0;
    const uint32_t bar_id = 1u + static_cast<uint32_t>(eg);
    const bool out_f32 = g.epi == PGF_EPI_ATOMIC_F32 || g.epi == PGF_EPI_STORE_F32 || g.epi == PGF_EPI_BIAS_F32 ||
                         g.epi == PGF_EPI_BIAS_TANH_F32;
    const bool has_bias = g.epi == PGF_EPI_BIAS_RELU_BF16 || g.epi == PGF_EPI_BIAS_TANH_BF16 || g.epi == PGF_EPI_BIAS_F32 ||
                          g.epi == PGF_EPI_BIAS_TANH_F32;
    // ReLU sign bits, one uint32 per (row, 32 columns): written by the forward epilogue, read back by the
    // backward one -- 1/16 of the bytes of the bf16 activation tile the mask would otherwise be derived from
    const bool mask_out = g.epi == PGF_EPI_BIAS_RELU_BF16 && g.aux != nullptr;
    const bool mask_in = g.epi == PGF_EPI_BITMASK_BF16;
    uint32_t* mask_words = static_cast<uint32_t*>(const_cast<void*>(g.aux));
    const bool want_cs = g.col_partial != nullptr && !out_f32;
    const int GW = out_f32 ? 32 : 64;  // columns per 128-byte staging row
    const uint32_t stage_base = epi_base + static_cast<uint32_t>(eg) * 32768u;
    const int neg = g.epi_groups;  // 2, or 1 = only the first group works (tuning knob; idle warps just release TMEM)
    const uint32_t swz = static_cast<uint32_t>(row & 7);
    const uint32_t row_off = static_cast<uint32_t>(row) * 128u;
    Scheduler sched(g, TILE_M, worker, nworkers);
    WorkUnit u;
    uint32_t unit = 0, gcount = 0;
    const uint32_t tempty_leader = CG == 2 ? mapa_u32(tempty_bar, 0) : tempty_bar;
    uint64_t store_policy = 0;
    if (g.store_hint) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(store_policy));
    while (sched.next(u)) {
      const uint32_t acc = unit & 1, acc_phase = (unit >> 1) & 1;
      const int m0 = (u.tile / nb_n) * TILE_M + static_cast<int>(rank) * BM, n0 = (u.tile % nb_n) * BN;
      uint4 mwa = make_uint4(0u, 0u, 0u, 0u), mwb = make_uint4(0u, 0u, 0u, 0u);
      if (mask_in && m0 + row < g.M) {  // this row's 256 sign bits of the tile: issued before the accumulator wait
        const uint4* mp = reinterpret_cast<const uint4*>(mask_words + static_cast<long long>(m0 + row) * g.ld_aux + (n0 >> 5));
        mwa = __ldg(mp);
        if (n0 + 128 < g.N) mwb = __ldg(mp + 1);
      }
      mbar_wait(tfull_bar + 8 * acc, acc_phase);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(quarter * 32) << 16);
      // column partials (bias gradient / dDP): one row per 32-row accumulator slab = per warp, no cross-warp step
      float* cp_row = g.col_partial ? g.col_partial + (static_cast<long long>(m0 >> 5) + quarter) * g.N : nullptr;
      if (g.epi == PGF_EPI_DDP_PARTIAL) {
        // dX tile never leaves the SM: multiply by the regenerated Laplace noise of (global row, column) and
        // reduce over the warp's 32 accumulator rows.
        const uint32_t grow = static_cast<uint32_t>(g.row0 + static_cast<unsigned long long>(m0 + row));
#pragma unroll 1
        for (int n = n0 + 32 * eg; eg < neg && n < n0 + BN && n < g.N; n += 32 * neg) {
          uint32_t v[32];
          tmem_ld32(taddr + (n - n0), v);
          float f[32];
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint4 r = philox4x32_10_rk(static_cast<uint32_t>((n >> 2) + q), grow, PGF_STREAM_LAPLACE, g.offset, g.rk);
            f[4 * q + 0] = __uint_as_float(v[4 * q + 0]) * laplace_from_bits(r.x);
            f[4 * q + 1] = __uint_as_float(v[4 * q + 1]) * laplace_from_bits(r.y);
            f[4 * q + 2] = __uint_as_float(v[4 * q + 2]) * laplace_from_bits(r.z);
            f[4 * q + 3] = __uint_as_float(v[4 * q + 3]) * laplace_from_bits(r.w);
          }
          const float cs = warp_colsum32(f, lane);
          if (n + lane < g.N) cp_row[n + lane] = cs;
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
          if (CG == 2) mbar_arrive_cluster_nofence(tempty_leader + 8 * acc);
          else mbar_arrive(tempty_bar + 8 * acc);
        }
        ++unit;
        continue;
      }
#pragma unroll 1
      for (int n = n0 + GW * eg; eg < neg && n < n0 + BN && n < g.N; n += neg * GW) {  // warp-uniform
        const uint32_t sbuf = stage_base + (gcount & 1) * 16384;
        uint32_t mw[2] = {0u, 0u};
        if (out_f32) {
          uint32_t v[32];
          tmem_ld32(taddr + (n - n0), v);
          float f[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[i]);
          if (has_bias) {
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
              if (n + i < g.N) {
                const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + n + i));
                f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
              }
            }
            if (g.epi == PGF_EPI_BIAS_TANH_F32) {
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = tanh_fast(f[i]);
            }
          }
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t dst = sbuf + row_off + ((static_cast<uint32_t>(c) ^ swz) << 4);
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "f"(f[4 * c]), "f"(f[4 * c + 1]),
                         "f"(f[4 * c + 2]), "f"(f[4 * c + 3])
                         : "memory");
          }
        } else {
          if (mask_in) {
            const int gi = (n - n0) >> 6;  // 64-column group inside the tile
            mw[0] = gi == 0 ? mwa.x : (gi == 1 ? mwa.z : (gi == 2 ? mwb.x : mwb.z));
            mw[1] = gi == 0 ? mwa.y : (gi == 1 ? mwa.w : (gi == 2 ? mwb.y : mwb.w));
          }
          uint32_t v[64];
          tmem_ld64(taddr + (n - n0), v);  // the whole 64-column group in one TMEM load
#pragma unroll
          for (int half = 0; half < 2; ++half) {
            float f[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) f[i] = __uint_as_float(v[32 * half + i]);
            if (has_bias) {
#pragma unroll
              for (int i = 0; i < 32; i += 4) {
                if (n + 32 * half + i < g.N) {
                  const float4 b = __ldg(reinterpret_cast<const float4*>(g.bias + n + 32 * half + i));
                  f[i] += b.x; f[i + 1] += b.y; f[i + 2] += b.z; f[i + 3] += b.w;
                }
              }
              if (g.epi == PGF_EPI_BIAS_RELU_BF16) {
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = fmaxf(f[i], 0.f);
                if (mask_out) {
                  uint32_t m = 0u;
#pragma unroll
                  for (int i = 0; i < 32; ++i) m |= (f[i] > 0.f ? 1u : 0u) << i;
                  mw[half] = m;
                }
              } else {
#pragma unroll
                for (int i = 0; i < 32; ++i) f[i] = tanh_fast(f[i]);
              }
            } else if (mask_in) {
              const uint32_t m = mw[half];
#pragma unroll
              for (int i = 0; i < 32; ++i) f[i] = (m & (1u << i)) ? f[i] : 0.f;
            }
#pragma unroll
            for (int c = 0; c < 4; ++c) {
              const uint32_t dst = sbuf + row_off + ((static_cast<uint32_t>(4 * half + c) ^ swz) << 4);
              asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(dst), "r"(pack_bf16x2(f[8 * c], f[8 * c + 1])),
                           "r"(pack_bf16x2(f[8 * c + 2], f[8 * c + 3])), "r"(pack_bf16x2(f[8 * c + 4], f[8 * c + 5])),
                           "r"(pack_bf16x2(f[8 * c + 6], f[8 * c + 7]))
                           : "memory");
            }
            // fused bias gradient: column sums of the (fp32) epilogue values over this warp's 32 rows
            if (want_cs) {
              const float cs = warp_colsum32(f, lane);
              if (n + 32 * half + lane < g.N) cp_row[n + 32 * half + lane] = cs;
            }
          }
        }
        if (mask_out && m0 + row < g.M)
          *reinterpret_cast<uint2*>(mask_words + static_cast<long long>(m0 + row) * g.ld_aux + (n >> 5)) = make_uint2(mw[0], mw[1]);
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy smem writes -> visible to TMA
        // the previous store of this group has finished reading its staging buffer (= the one the next column group
        // will overwrite); arriving at the barrier below publishes that to the other three warps
        if (leader_thread) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        asm volatile("bar.sync %0, 128;" ::"r"(bar_id) : "memory");
        if (leader_thread) {
          if (g.epi == PGF_EPI_ATOMIC_F32) tma_reduce_add_2d(&tmC, sbuf, n, m0);
          else if (g.store_hint) tma_store_2d_hint(&tmC, sbuf, n, m0, store_policy);
          else tma_store_2d(&tmC, sbuf, n, m0);  // rows >= M / columns >= N are clipped by the tensor map
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
        ++gcount;
      }
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster_nofence(tempty_leader + 8 * acc);
        else mbar_arrive(tempty_bar + 8 * acc);
      }
      ++unit;
    }
    if (leader_thread) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // all outputs landed before exit
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  if (CG == 2) cluster_sync_all(); else __syncthreads();  // nobody frees TMEM / exits while the peer still needs it
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    if (CG == 1)
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// row-major tensor [planes][rows, cols] (cols contiguous, row stride ld elements, `plane` elements between planes),
// box {box_cols, box_rows}, 128B swizzle.  planes == 0: plain 2-D map.
static int make_tmap(CUtensorMap* map, const void* ptr, long long rows, long long cols, long long ld, int box_cols,
                     int box_rows, bool f32 = false, int planes = 0, long long plane = 0) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) {
    set_error("pgf_gemm_bf16: cuTensorMapEncodeTiled not available from the driver");
    return PGF_ERR_CUDA;
  }
  const cuuint64_t esz = f32 ? 4 : 2;
  const cuuint64_t dims[3] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows), static_cast<cuuint64_t>(planes)};
  const cuuint64_t strides[2] = {static_cast<cuuint64_t>(ld) * esz, static_cast<cuuint64_t>(plane) * esz};
  const cuuint32_t box[3] = {static_cast<cuuint32_t>(box_cols), static_cast<cuuint32_t>(box_rows), 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  const CUresult r = fn(map, f32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, planes > 0 ? 3 : 2,
                        const_cast<void*>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("pgf_gemm_bf16: cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld ld=%lld planes=%d", static_cast<int>(r), rows,
              cols, ld, planes);
    return PGF_ERR_CUDA;
  }
  return PGF_OK;
}

// SMs the persistent GEMM grids leave free (pgf_set_sm_reserve): while a collective is in flight on another stream its
// kernels need SMs of their own -- a one-CTA-per-SM grid that does not fit next to them has its last clusters wait for
// the first ones to finish, which costs far more than the SMs given up.
static thread_local int g_sm_reserve = 0;
void set_sm_reserve(int n) { g_sm_reserve = n < 0 ? 0 : n; }

int gemm_bf16(const void* A, long long lda, int a_mn, const void* B, long long ldb, int b_mn, const GemmArgs& g_in,
              cudaStream_t s) {
  GemmArgs g = g_in;
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return PGF_OK;
  const bool out_f32 = g.epi == PGF_EPI_ATOMIC_F32 || g.epi == PGF_EPI_STORE_F32 || g.epi == PGF_EPI_BIAS_F32 ||
                       g.epi == PGF_EPI_BIAS_TANH_F32;
  if ((g.N % 8) || (lda % 8) || (ldb % 8) || (g.epi != PGF_EPI_DDP_PARTIAL && (g.ldc % (out_f32 ? 4 : 8))) ||
      ((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B) | reinterpret_cast<uintptr_t>(g.C)) & 15)) {
    set_error("pgf_gemm_bf16: N, lda, ldb must be multiples of 8, ldc of 8 (bf16 out) / 4 (fp32 out), pointers 16-byte aligned");
    return PGF_ERR_ARG;
  }
  CUtensorMap tmA, tmB, tmC;
  int rc;
  const int planes = g.nseg > 1 ? 3 : 0;
  if (planes) {
    if (!out_f32 || g.col_partial || g.nseg > 8 || (g.a_plane % 8) || (g.b_plane % 8)) {
      set_error("pgf_gemm_bf16x3: fp32-output epilogues only (4, 5, 6), no column partials, plane strides multiples of 8");
      return PGF_ERR_ARG;
    }
  } else {
    g.nseg = 1;
  }
  // output staging rows are 128 bytes: 64 bf16 or 32 fp32 columns x 128 accumulator rows per TMA store
  const bool no_c = g.epi == PGF_EPI_DDP_PARTIAL;
  if (no_c) {
    if (!g.col_partial || g.stream_k) {
      set_error("pgf_gemm_bf16: the dDP epilogue needs the column-partial workspace and no split-K");
      return PGF_ERR_ARG;
    }
    g.rk = philox_make_keys(g.seed);
  } else if (g.col_partial && out_f32) {
    set_error("pgf_gemm_bf16: column partials are fused into the bf16-output epilogues only");
    return PGF_ERR_ARG;
  }
  if (g.epi == PGF_EPI_BITMASK_BF16 || (g.epi == PGF_EPI_BIAS_RELU_BF16 && g.aux)) {
    if (!g.aux || (g.N % 128) || (g.ld_aux % 4) || (reinterpret_cast<uintptr_t>(g.aux) & 15) || g.ld_aux * 32 < g.N) {
      set_error("pgf_gemm_bf16: the ReLU bitmask needs N %% 128 == 0, a 16-byte aligned uint32 [M, ld_aux >= N/32] buffer, ld_aux %% 4 == 0");
      return PGF_ERR_ARG;
    }
  }
  // K-major operand [R,K]: box {64 k, BM|BN rows}.  MN-major operand stored [K,R]: box {64 r, 64 k}.
  rc = a_mn ? make_tmap(&tmA, A, g.K, g.M, lda, 64, 64, false, planes, g.a_plane)
            : make_tmap(&tmA, A, g.M, g.K, lda, BK, BM, false, planes, g.a_plane);
  if (rc != PGF_OK) return rc;
  if (no_c) {
    tmC = tmA;  // never dereferenced
  } else {
    rc = make_tmap(&tmC, g.C, g.M, g.N, g.ldc, out_f32 ? 32 : 64, BM, out_f32);
    if (rc != PGF_OK) return rc;
  }
  static const bool force_1cta_b = getenv("PGF_GEMM_1CTA") != nullptr;
  const int cg_b = (!force_1cta_b && g.M > BM) ? 2 : 1;
  rc = b_mn ? make_tmap(&tmB, B, g.K, g.N, ldb, 64, 64, false, planes, g.b_plane)
            : make_tmap(&tmB, B, g.N, g.K, ldb, BK, BN / cg_b, false, planes, g.b_plane);
  if (rc != PGF_OK) return rc;

  // CTA pairs (cta_group::2) whenever there is at least one full 256-row tile of work per pair
  static const bool force_1cta = getenv("PGF_GEMM_1CTA") != nullptr;
  const int cg = (!force_1cta && g.M > BM) ? 2 : 1;
  const int tile_m = BM * cg;
  const int tiles = ((g.M + tile_m - 1) / tile_m) * ((g.N + BN - 1) / BN);
  const int kb_k = (g.K + BK - 1) / BK, kb_total = kb_k * g.nseg;
  const int sms_usable = num_sms() - g_sm_reserve > 2 * cg ? num_sms() - g_sm_reserve : 2 * cg;
  const int workers_max = sms_usable / cg;
  int workers = tiles < workers_max ? tiles : workers_max;
  if (g.stream_k) {
    if (g.epi != PGF_EPI_ATOMIC_F32) {
      set_error("pgf_gemm_bf16: split-K needs the fp32 reduction epilogue");
      return PGF_ERR_ARG;
    }
    // number of K slabs: small enough that one slab of both operands stays L2-resident (<= 64 MB),
    // then whatever minimises waves x (slab length + epilogue) over the candidates.  stream_k > 1: the caller's
    // count (the fp32-parity mode bounds the length of one TMEM accumulation chain this way).
    const double slab_bytes_per_kb = 2.0 * BK * (static_cast<double>(g.M) + g.N);
    int s_min = static_cast<int>(slab_bytes_per_kb * kb_total / (64.0 * 1024 * 1024)) + 1;
    if (s_min > kb_k) s_min = kb_k;
    int best = s_min;
    double best_cost = 1e300;
    for (int S = s_min; S <= s_min + 24 && S <= kb_k; ++S) {
      const long long units = static_cast<long long>(tiles) * S;
      const long long waves = (units + workers_max - 1) / workers_max;
      const double cost = static_cast<double>(waves) * (static_cast<double>(kb_total) / S + 4.0);
      if (cost < best_cost) { best_cost = cost; best = S; }
    }
    if (g.stream_k > 1) best = g.stream_k < kb_k ? g.stream_k : kb_k;
    // a C far larger than L2 (forward-type GEMMs cut into slabs for accuracy): the slabs of a tile run back to back
    g.tile_major = static_cast<double>(g.M) * g.N * 4.0 > 48.0 * 1024 * 1024 ? 1 : 0;
    g.stream_k = best;
    const long long units = static_cast<long long>(tiles) * best;
    workers = static_cast<int>(units < workers_max ? units : workers_max);
  }
  // evict_first on plain output stores of tensors far larger than L2 (PGF_GEMM_STORE_HINT=0/1 overrides)
  static const int hint_env = getenv("PGF_GEMM_STORE_HINT") ? atoi(getenv("PGF_GEMM_STORE_HINT")) : -1;
  g.store_hint = hint_env >= 0 ? hint_env : 0;
  static const bool one_epi_group = getenv("PGF_GEMM_EPI_GROUPS") != nullptr && getenv("PGF_GEMM_EPI_GROUPS")[0] == '1';
  g.epi_groups = one_epi_group ? 1 : 2;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(workers * cg);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cg;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t lerr = cudaSuccess;
#define PGF_GEMM_LAUNCH(AM, BMN, CGV)                                                                                     \
  do {                                                                                                                    \
    cfg.dynamicSmemBytes = Cfg<CGV>::SMEM_BYTES;                                                                          \
    ensure_dynamic_smem(reinterpret_cast<const void*>(gemm_bf16_tc_kernel<AM, BMN, CGV>), Cfg<CGV>::SMEM_BYTES);          \
    lerr = cudaLaunchKernelEx(&cfg, gemm_bf16_tc_kernel<AM, BMN, CGV>, tmA, tmB, tmC, g);                                      \
  } while (0)
#define PGF_GEMM_DISPATCH(CGV)                               \
  do {                                                       \
    if (a_mn && b_mn) PGF_GEMM_LAUNCH(true, true, CGV);      \
    else if (a_mn) PGF_GEMM_LAUNCH(true, false, CGV);        \
    else if (b_mn) PGF_GEMM_LAUNCH(false, true, CGV);        \
    else PGF_GEMM_LAUNCH(false, false, CGV);                 \
  } while (0)
  if (cg == 2) PGF_GEMM_DISPATCH(2);
  else PGF_GEMM_DISPATCH(1);
#undef PGF_GEMM_DISPATCH
#undef PGF_GEMM_LAUNCH
  if (lerr != cudaSuccess) {
    set_error("pgf_gemm_bf16: launch failed: %s", cudaGetErrorString(lerr));
    return PGF_ERR_CUDA;
  }
  PGF_CUDA_LAUNCH_CHECK("pgf_gemm_bf16");
  return PGF_OK;
}

// rows of the column-partial workspace for an M-row output: one per 32-row accumulator slab (= per epilogue warp)
int gemm_partial_rows(int M) {
  if (M <= 0) return 0;
  static const bool force_1cta = getenv("PGF_GEMM_1CTA") != nullptr;
  const int cg = (!force_1cta && M > BM) ? 2 : 1;
  const int tile_m = BM * cg;
  return ((M + tile_m - 1) / tile_m) * (tile_m / 32);
}

}  // namespace pgf
