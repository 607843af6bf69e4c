// prefill_tc.cuh — prefill (many positions at once) on the 5th-generation tensor cores.
//
// The decode kernel (megakernel.cuh) is an M=1 streaming GEMV and deliberately stays off the tensor cores.  A prompt of
// S positions is GEMM-shaped: per layer  [S,D]x[D,6144], [S,4096]x[4096,D], [S,D]x[D,2F], [S,F]x[F,D]  (+ lm_head).
// These run here as a hand-written sm_100a GEMM:  D[128 tokens, NT features] (fp32, TMEM) += X tile x W tile^T
//   * operands: TMA tensor-map loads (cp.async.bulk.tensor, SWIZZLE_128B, K-major) into a shared-memory ring;
//     the weight map is a 3-D view {TK, N, K/TK} of the K-panel layout the decode kernel streams (one layout, two kernels)
//   * math: tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> fp32), M=128, N=NT, K=16, issued by ONE thread,
//     accumulator in tensor memory, completion via tcgen05.commit -> mbarrier
//   * epilogue: 4 warps tcgen05.ld their 32 TMEM lanes (= 32 tokens), apply the fused epilogue (bf16 store / residual
//     add / last-row logits) with the same rounding points as the decode kernel
// Replaces the reference's prefill pass through torchtune modules (llm_utils.py:416-489 with S = prompt length;
// sharded_inference_engine.py:108-193,230-370).  Around the GEMMs: RMSNorm, RoPE + KV append, causal attention over the
// prompt, SwiGLU, arg-max — small row-parallel kernels.
#pragma once
#include <cuda.h>   // CUtensorMap + enums only; cuTensorMapEncodeTiled is fetched with cudaGetDriverEntryPoint (no -lcuda)

#include "kernels.cuh"

namespace xb {

constexpr int TC_M = 128;        // tokens per GEMM tile = UMMA M
constexpr int TC_BK = 64;        // K elements per ring stage = one 128-byte swizzle row
constexpr int TC_THREADS = 256;  // warp 0 TMA producer, warp 1 MMA issuer, warp 2 TMEM allocator, warps 4..7 epilogue

enum { TCE_STORE = 0, TCE_RESID = 1, TCE_HEAD_LAST = 2, TCE_QKV = 3, TCE_SWIGLU = 4 };

constexpr int TC_CHUNK = 1024;   // tokens per prefill pass: every weight tile is fetched from HBM once per chunk and reused
                                 // by the chunk's 128-token tiles out of L2 (grid x = token tile, so they run together)

struct TcGemmArgs {
  int S;            // valid token rows of the chunk (grid.x = ceil(S/128) token tiles; HEAD_LAST: only the last tile)
  int N, K, tk;     // output features, reduction length, K-panel width of the weight layout
  bf16* out;        // STORE: [S, ld] bf16;  RESID: residual stream x [S, ld], updated in place
  float* logits;    // HEAD_LAST: [N] fp32, row S-1 only (bf16-valued)
  int ld;           // row stride of `out` in elements
  int mt0;          // first token tile of this launch (0, or the last tile for HEAD_LAST)
  const int* tile_info;   // sparse MoE (NULL otherwise): {first row of this expert's block, its 128-row tiles}; CTAs beyond exit
  int grid_m;             // sparse MoE: token tiles to launch (the most one expert can get); 0 = ceil(S / 128) - mt0
  // TCE_QKV: RoPE + KV append fused into the QKV GEMM (out = rotated q [S, q_rows]); TCE_SWIGLU: out = act [S, N/2]
  bf16 *kc, *vc;          // K / V cache of (request, layer): [Hkv][L][hd]
  const float2* rope;     // [L][hd/2] (cos, sin)
  int q_rows, kv_rows, hd, L, pos0;
  // split-K (one 128-token tile: the GEMM is weight-streaming bound and N / NT tiles do not fill the GPU): grid.z = ksplit CTAs
  // share an output tile, each sweeps K / ksplit; fp32 partial tiles go through `ws`, the CTA drawing the last ticket of the tile
  // adds them in split order (deterministic) and runs the epilogue
  int ksplit;             // 1 = off
  float* ws;              // [tile][ksplit][128][NT] fp32
  unsigned int* ws_ticket;   // [tile], zero between launches (the last CTA resets its entry)
  int* err;               // sticky device error word (StepCtl::err): 3 = a bounded wait of the persistent GEMM expired
};

__host__ __device__ constexpr int tc_stage_bytes(int nt) { return TC_M * TC_BK * 2 + nt * TC_BK * 2; }
__host__ __device__ constexpr int tc_smem_bytes(int nt, int ns) { return ns * tc_stage_bytes(nt) + 1024 + 256; }
// CTAs per SM the launch bounds ask for: two when two rings fit, so one CTA's TMEM->global epilogue overlaps the other's MMAs
__host__ __device__ constexpr int tc_occupancy(int nt, int ns) { return 2 * tc_smem_bytes(nt, ns) <= 226 * 1024 && 2 * nt <= 512 ? 2 : 1; }

// ---- PTX wrappers (forms as in CUTLASS cute/arch/{copy_sm90_tma,copy_sm100,mma_sm100_umma,tmem_allocator_sm100}.hpp) ----
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint64_t* bar) {
  asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t* slot_smem, int ncols) {   // whole warp
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, int ncols) {      // whole warp
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
// shared-memory matrix descriptor, K-major, SWIZZLE_128B (cute/arch/mma_sm100_desc.hpp SmemDescriptor):
// start>>4 [0,14) | LBO>>4 [16,30) (=1, unused for swizzled K-major) | SBO>>4 [32,46) (8 rows x 128 B = 1024) | version=1 [46,48) | layout=2 [61,64)
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)(1024 >> 4) << 32) | ((uint64_t)1 << 46) |
         ((uint64_t)2 << 61);
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8}, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {   // arrives on `bar` when all MMAs issued so far have completed
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
                 "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
               : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- CTA-pair (cta_group::2) forms: one 256 x NT tile per cluster of two CTAs.  Each CTA stages its own 128 rows of A and HALF
//      of the B tile, the leader's single thread issues tcgen05.mma.cta_group::2 (reads both CTAs' shared memory, writes both
//      CTAs' tensor memory), commits are multicast to the same mbarrier in both CTAs (forms as in CUTLASS
//      cute/arch/{copy_sm100_tma,mma_sm100_umma,tmem_allocator_sm100}.hpp and cutlass/arch/barrier.h) ----
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {   // same offset in CTA `rank` of the cluster
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA loads whose completion is counted on an mbarrier that may live in the PEER CTA (cluster address `bar_cluster`)
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, int c0, int c1, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, int c0, int c1, int c2, uint32_t bar_cluster) {
  asm volatile("cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
                   smem_u32(dst)),
               "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void tmem_alloc_2sm(uint32_t* slot_smem, int ncols) {   // one warp in EACH CTA of the pair, same slot offset
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot_smem)), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_2sm(uint32_t taddr, int ncols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %6, %7, %8, %9, %10, %11, %12}, p;\n\t}"
      ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u), "r"(0u)
      : "memory");
}
// arrives on `bar` (same offset) in every CTA of `mask` once all MMAs issued so far by this thread have completed
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar, uint16_t mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)), "h"(mask)
               : "memory");
}
__host__ __device__ constexpr int tc2_stage_bytes(int nt) { return TC_M * TC_BK * 2 + (nt / 2) * TC_BK * 2; }   // per CTA: own A rows + half of B
__host__ __device__ constexpr int tc2_smem_bytes(int nt, int ns) { return ns * tc2_stage_bytes(nt) + 1024 + 256; }

// ------------------------------------------------------------------------------------------------------------
// row-parallel helpers around the GEMMs
// ------------------------------------------------------------------------------------------------------------
// x[s, :] = embed[tokens[s], :]   (first shard)
__global__ void tc_gather_kernel(const int* tokens, const bf16* embed, bf16* x, int D) {
  const uint4* s4 = reinterpret_cast<const uint4*>(embed + (size_t)tokens[blockIdx.x] * D);
  uint4* d4 = reinterpret_cast<uint4*>(x + (size_t)blockIdx.x * D);
  for (int c = threadIdx.x; c < D / 8; c += blockDim.x) d4[c] = s4[c];
}

// lowest-index arg-max of logits[0..V) -> ctl->token   (single block of 1024 threads)
__global__ void __launch_bounds__(1024) tc_argmax_kernel(const float* logits, int V, StepCtl* ctl) {
  __shared__ float sv[32];
  __shared__ int si[32];
  float bv = -INFINITY;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < V; i += 1024) {
    const float v = logits[i];
    if (v > bv) { bv = v; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, bv, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
  }
  if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 32; ++w)
      if (sv[w] > bv || (sv[w] == bv && si[w] < bi)) { bv = sv[w]; bi = si[w]; }
    ctl->token = bi;
  }
}

// sparse MoE prefill: rows of the gathered buffers per expert.  One block: count -> 128-row-aligned offsets -> assign rows.
// exp_info[2e] = first row of expert e, [2e+1] = its 128-row tiles, [2E] = total rows; row_of[2s+k] = row of token s's k-th
// expert (ascending index); row_token[row] = token or -1 (padding).
__global__ void __launch_bounds__(1024) moe_plan_kernel(const int* route_e, int S, int E, int rows_max, int* row_of, int* row_token,
                                                        int* exp_info) {
  __shared__ int cnt[8], off[9];
  if (threadIdx.x < 8) cnt[threadIdx.x] = 0;
  __syncthreads();
  for (int i = threadIdx.x; i < 2 * S; i += 1024) atomicAdd(&cnt[route_e[i]], 1);
  __syncthreads();
  if (threadIdx.x == 0) {
    off[0] = 0;
    for (int x = 0; x < E; ++x) {
      const int tiles = (cnt[x] + TC_M - 1) / TC_M;
      off[x + 1] = off[x] + tiles * TC_M;
      exp_info[2 * x] = off[x];
      exp_info[2 * x + 1] = tiles;
    }
    exp_info[2 * E] = off[E];
  }
  for (int r = threadIdx.x; r < rows_max; r += 1024) row_token[r] = -1;
  __syncthreads();
  // rows inside an expert block in token order (deterministic): one warp-free serial pass per expert would be slow, so each
  // thread ranks its own pair by counting earlier pairs of the same expert
  for (int i = threadIdx.x; i < 2 * S; i += 1024) {
    const int x = route_e[i];
    int rank = 0;
    for (int j = 0; j < i; ++j) rank += (route_e[j] == x) ? 1 : 0;
    const int r = off[x] + rank;
    row_of[i] = r;
    row_token[r] = i >> 1;
  }
}

}  // namespace xb
