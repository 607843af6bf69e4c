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

enum { TCE_STORE = 0, TCE_RESID = 1, TCE_HEAD_LAST = 2 };

struct TcGemmArgs {
  int S;            // valid token rows of this tile (<= 128)
  int N, K, tk;     // output features, reduction length, K-panel width of the weight layout
  bf16* out;        // STORE: [S, ld] bf16;  RESID: residual stream x [S, ld], updated in place
  float* logits;    // HEAD_LAST: [N] fp32, row S-1 only (bf16-valued)
  int ld;           // row stride of `out` in elements
};

__host__ __device__ constexpr int tc_stage_bytes(int nt) { return TC_M * TC_BK * 2 + nt * TC_BK * 2; }
__host__ __device__ constexpr int tc_nstage(int nt) { return nt <= 64 ? 8 : 6; }
__host__ __device__ constexpr int tc_smem_bytes(int nt) { return tc_nstage(nt) * tc_stage_bytes(nt) + 1024 + 256; }

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
// instruction descriptor (InstrDescriptor): D=f32 [4,6)=1, A=bf16 [7,10)=1, B=bf16 [10,13)=1, A/B K-major, N>>3 [17,23), M>>4 [24,29)
__host__ __device__ constexpr uint32_t umma_idesc_bf16(int m, int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
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

// ------------------------------------------------------------------------------------------------------------
// D[128, NT] = X[128, K] * W[n0:n0+NT, K]^T   — one output tile per CTA, K swept through a TMA/mbarrier ring
// ------------------------------------------------------------------------------------------------------------
template <int NT, int EPI>
__global__ void __launch_bounds__(TC_THREADS, 1)
tc_gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcGemmArgs a) {
  extern __shared__ __align__(1024) uint8_t tc_smem[];
  constexpr int A_BYTES = TC_M * TC_BK * 2;
  constexpr int STAGE = tc_stage_bytes(NT);
  constexpr int NS = tc_nstage(NT);
  constexpr int TMEM_COLS = NT < 32 ? 32 : NT;      // power of two >= 32
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(tc_smem) + 1023) & ~(uintptr_t)1023);
  uint64_t* full = reinterpret_cast<uint64_t*>(base + NS * STAGE);
  uint64_t* empty = full + NS;
  uint64_t* tmem_full = empty + NS;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_full + 1);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * NT;
  const int nkb = a.K / TC_BK;

  if (threadIdx.x == 0) {
    for (int i = 0; i < NS; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], 1); }
    mbar_init(tmem_full, 1);
    fence_mbar_init();
  }
  if (warp == 2) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0 && lane == 0) {
    // ===== TMA producer =====
    for (int kb = 0; kb < nkb; ++kb) {
      const int st = kb % NS;
      mbar_wait(&empty[st], (((unsigned)kb / NS) & 1u) ^ 1u);
      uint8_t* sa = base + (size_t)st * STAGE;
      mbar_expect_tx(&full[st], STAGE);
      tma_load_2d(sa, &tmA, kb * TC_BK, 0, &full[st]);
      const int kcol = kb * TC_BK;
      tma_load_3d(sa + A_BYTES, &tmB, kcol % a.tk, n0, kcol / a.tk, &full[st]);
    }
  } else if (warp == 1 && lane == 0) {
    // ===== MMA issuer (single thread) =====
    constexpr uint32_t idesc = umma_idesc_bf16(TC_M, NT);
    for (int kb = 0; kb < nkb; ++kb) {
      const int st = kb % NS;
      mbar_wait(&full[st], ((unsigned)kb / NS) & 1u);
      tc_fence_after();
      const uint32_t sa = smem_u32(base + (size_t)st * STAGE);
      const uint32_t sb = sa + A_BYTES;
#pragma unroll
      for (int k = 0; k < TC_BK / 16; ++k)      // 4 x (K=16) per 128-byte swizzle row: advance the start address by 32 B
        umma_bf16(tmem_base, umma_desc_sw128(sa + k * 32), umma_desc_sw128(sb + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
      umma_commit(&empty[st]);                  // frees the smem stage once these MMAs have read it
    }
    umma_commit(tmem_full);                     // accumulator complete
  } else if (warp >= 4) {
    // ===== epilogue: warp q owns TMEM lanes 32q..32q+31 = token rows =====
    const int q = warp - 4;
    const int row = q * 32 + lane;
    mbar_wait(tmem_full, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < NT; c0 += 16) {
      uint32_t v[16];
      tmem_ld16(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      const int n = n0 + c0;
      if (EPI == TCE_STORE) {
        if (row < a.S) {
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            __nv_bfloat162 p = __floats2bfloat162_rn(__uint_as_float(v[2 * j]), __uint_as_float(v[2 * j + 1]));
            o[j] = *reinterpret_cast<uint32_t*>(&p);
          }
          uint4* dst = reinterpret_cast<uint4*>(a.out + (size_t)row * a.ld + n);
          dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
          dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      } else if (EPI == TCE_RESID) {
        // h = x + Linear(..): Linear out -> bf16, add in fp32 -> bf16 (App. A.5, A.7); in place on the residual stream
        if (row < a.S) {
          uint4* px = reinterpret_cast<uint4*>(a.out + (size_t)row * a.ld + n);
          uint4 r0 = px[0], r1 = px[1];
          const uint32_t rr[8] = {r0.x, r0.y, r0.z, r0.w, r1.x, r1.y, r1.z, r1.w};
          uint32_t o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float y0 = bf_lo(rr[j]) + round_bf(__uint_as_float(v[2 * j]));
            const float y1 = bf_hi(rr[j]) + round_bf(__uint_as_float(v[2 * j + 1]));
            __nv_bfloat162 p = __floats2bfloat162_rn(y0, y1);
            o[j] = *reinterpret_cast<uint32_t*>(&p);
          }
          px[0] = make_uint4(o[0], o[1], o[2], o[3]);
          px[1] = make_uint4(o[4], o[5], o[6], o[7]);
        }
      } else {
        if (row == a.S - 1) {
#pragma unroll
          for (int j = 0; j < 16; ++j) a.logits[n + j] = round_bf(__uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------------------
// row-parallel helpers around the GEMMs
// ------------------------------------------------------------------------------------------------------------
// x[s, :] = embed[tokens[s], :]   (first shard)
__global__ void tc_gather_kernel(const int* tokens, const bf16* embed, bf16* x, int D) {
  const uint4* s4 = reinterpret_cast<const uint4*>(embed + (size_t)tokens[blockIdx.x] * D);
  uint4* d4 = reinterpret_cast<uint4*>(x + (size_t)blockIdx.x * D);
  for (int c = threadIdx.x; c < D / 8; c += blockDim.x) d4[c] = s4[c];
}

// RMSNorm of row s (torchtune<=0.6 / HF order, App. A.1): fp32 normalise -> bf16 -> * scale (bf16) -> bf16.  block = 256
__global__ void __launch_bounds__(256) tc_rmsnorm_kernel(const bf16* x, const bf16* w, bf16* out, int D, float eps) {
  __shared__ float red[8];
  const int s = blockIdx.x;
  const uint4* x4 = reinterpret_cast<const uint4*>(x + (size_t)s * D);
  float ss = 0.f;
  for (int c = threadIdx.x; c < D / 8; c += 256) {
    float f[8];
    unpack8(x4[c], f);
#pragma unroll
    for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
  }
  ss = warp_sum(ss);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ss;
  __syncthreads();
  float tot = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) tot += red[i];
  const float rstd = 1.0f / sqrtf(tot / (float)D + eps);
  uint4* o4 = reinterpret_cast<uint4*>(out + (size_t)s * D);
  for (int c = threadIdx.x; c < D / 8; c += 256) {
    float f[8], sc[8];
    unpack8(x4[c], f);
    unpack8(reinterpret_cast<const uint4*>(w)[c], sc);
    uint32_t ov[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      __nv_bfloat162 p = __floats2bfloat162_rn(round_bf(f[2 * j] * rstd) * sc[2 * j], round_bf(f[2 * j + 1] * rstd) * sc[2 * j + 1]);
      ov[j] = *reinterpret_cast<uint32_t*>(&p);
    }
    o4[c] = make_uint4(ov[0], ov[1], ov[2], ov[3]);
  }
}

// qkv [S, (Hq+2Hkv)*hd] (bf16, interleaved-pair rows) -> q_all [S, Hq*hd] rotated, K/V cache rows pos0+s appended.
__global__ void tc_rope_kv_kernel(const bf16* qkv, bf16* q_all, bf16* k_cache, bf16* v_cache, const float2* rope, int Hq, int Hkv,
                                  int hd, int L, int pos0) {
  const int s = blockIdx.x, pos = pos0 + s;
  const int q_rows = Hq * hd, kv_rows = Hkv * hd, nq = q_rows + 2 * kv_rows;
  const __nv_bfloat162* in = reinterpret_cast<const __nv_bfloat162*>(qkv + (size_t)s * nq);
  for (int p = threadIdx.x; p < nq / 2; p += blockDim.x) {
    const int n = 2 * p;
    const __nv_bfloat162 v = in[p];
    const float x0 = __low2float(v), x1 = __high2float(v);
    if (n < q_rows + kv_rows) {
      const int ih = n % hd;
      const float2 cs = rope[(size_t)pos * (hd / 2) + (ih >> 1)];
      const __nv_bfloat162 o = __floats2bfloat162_rn(__fsub_rn(__fmul_rn(x0, cs.x), __fmul_rn(x1, cs.y)),
                                                      __fadd_rn(__fmul_rn(x1, cs.x), __fmul_rn(x0, cs.y)));
      if (n < q_rows) {
        *reinterpret_cast<__nv_bfloat162*>(q_all + (size_t)s * q_rows + n) = o;
      } else {
        const int nk = n - q_rows;
        *reinterpret_cast<__nv_bfloat162*>(k_cache + ((size_t)(nk / hd) * L + pos) * hd + ih) = o;
      }
    } else {
      const int nv = n - q_rows - kv_rows;
      *reinterpret_cast<__nv_bfloat162*>(v_cache + ((size_t)(nv / hd) * L + pos) * hd + (nv % hd)) = v;
    }
  }
}

// causal attention of query token s (position pos0+s) for kv head g over cache positions [0, pos0+s]; grid (S, Hkv), block 128
template <int HD, int R>
__global__ void __launch_bounds__(128) tc_attn_prefill_kernel(const bf16* q_all, const bf16* k_cache, const bf16* v_cache, bf16* out,
                                                              int Hq, int L, int pos0, float scale) {
  constexpr int LPP = HD / 8, PPW = 32 / LPP, AW = 4;
  const int s = blockIdx.x, g = blockIdx.y;
  const int ctx = pos0 + s + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPP, li = lane % LPP;
  float q[R][8], acc[R][8], m[R], l[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    unpack8(*reinterpret_cast<const uint4*>(q_all + ((size_t)s * Hq + g * R + r) * HD + li * 8), q[r]);
    m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[r][e] = 0.f;
  }
  const bf16* kb = k_cache + (size_t)g * L * HD + li * 8;
  const bf16* vb = v_cache + (size_t)g * L * HD + li * 8;
  for (int tb = warp * PPW; tb < ctx; tb += AW * PPW) {
    const int t = tb + sub;
    const bool valid = t < ctx;
    uint4 kc = make_uint4(0, 0, 0, 0), vc = kc;
    if (valid) {
      kc = *reinterpret_cast<const uint4*>(kb + (size_t)t * HD);
      vc = *reinterpret_cast<const uint4*>(vb + (size_t)t * HD);
    }
    float vf[8];
    unpack8(vc, vf);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float sc = dot8(kc, q[r], 0.f);
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) sc += __shfl_xor_sync(0xffffffffu, sc, o);
      if (valid) {
        sc *= scale;
        const float mn = fmaxf(m[r], sc);
        const float corr = expf(m[r] - mn);
        const float p = expf(sc - mn);
        l[r] = l[r] * corr + p;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[r][e] = acc[r][e] * corr + p * vf[e];
        m[r] = mn;
      }
    }
  }
#pragma unroll
  for (int o = 16; o >= LPP; o >>= 1) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float mo = __shfl_xor_sync(0xffffffffu, m[r], o);
      const float lo = __shfl_xor_sync(0xffffffffu, l[r], o);
      const float M = fmaxf(m[r], mo);
      const float fa = (m[r] == -INFINITY) ? 0.f : expf(m[r] - M);
      const float fb = (mo == -INFINITY) ? 0.f : expf(mo - M);
      const bool lower = (lane & o) == 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float ao = __shfl_xor_sync(0xffffffffu, acc[r][e], o);
        acc[r][e] = lower ? (acc[r][e] * fa + ao * fb) : (ao * fb + acc[r][e] * fa);
      }
      l[r] = lower ? (l[r] * fa + lo * fb) : (lo * fb + l[r] * fa);
      m[r] = M;
    }
  }
  __shared__ float s_ml[AW][R][2];
  __shared__ float s_acc[AW][R][HD];
  if (sub == 0) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (li == 0) { s_ml[warp][r][0] = m[r]; s_ml[warp][r][1] = l[r]; }
#pragma unroll
      for (int e = 0; e < 8; ++e) s_acc[warp][r][li * 8 + e] = acc[r][e];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < R * HD; idx += 128) {
    const int r = idx / HD, d = idx - r * HD;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < AW; ++w) M = fmaxf(M, s_ml[w][r][0]);
    float Ls = 0.f, A = 0.f;
#pragma unroll
    for (int w = 0; w < AW; ++w) {
      const float mw = s_ml[w][r][0];
      const float f = (mw == -INFINITY) ? 0.f : expf(mw - M);
      Ls += s_ml[w][r][1] * f;
      A += s_acc[w][r][d] * f;
    }
    out[((size_t)s * Hq + g * R + r) * HD + d] = f2bf(A / Ls);
  }
}

// gu [S, 2F] (gate_m, up_m interleaved) -> act [S, F] = bf16( bf16(silu(g)) * u )   (llm_utils.py:491-500)
// grid (S, ceil(F/2048)), block 256: each thread turns two 16-byte loads (8 pairs) into one 16-byte store
__global__ void __launch_bounds__(256) tc_swiglu_kernel(const bf16* gu, bf16* act, int F) {
  const int s = blockIdx.x;
  const int m0 = (blockIdx.y * 256 + threadIdx.x) * 8;
  if (m0 >= F) return;
  const uint4* in = reinterpret_cast<const uint4*>(gu + (size_t)s * 2 * F + 2 * (size_t)m0);
  const uint4 a0 = in[0], a1 = in[1];
  const uint32_t w[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
  uint32_t o[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const float g0 = bf_lo(w[2 * j]), u0 = bf_hi(w[2 * j]);
    const float g1 = bf_lo(w[2 * j + 1]), u1 = bf_hi(w[2 * j + 1]);
    __nv_bfloat162 p = __floats2bfloat162_rn(round_bf(g0 / (1.0f + expf(-g0))) * u0, round_bf(g1 / (1.0f + expf(-g1))) * u1);
    o[j] = *reinterpret_cast<uint32_t*>(&p);
  }
  *reinterpret_cast<uint4*>(act + (size_t)s * F + m0) = make_uint4(o[0], o[1], o[2], o[3]);
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

}  // namespace xb
