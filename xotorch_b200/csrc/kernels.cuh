// kernels.cuh — hand-written sm_100a kernels of the shard-decode step.
//
// One decode step of a shard (layers [a..b]) is the launch sequence
//     ingest -> { qkv -> attn -> oproj -> gateup -> down } x layers [-> head]
// replacing the ~50 eager launches per layer of the reference path (SURVEY.md §2.2 K1-K12;
// /root/reference/xotorch/inference/llm_utils.py:416-489 driving torchtune modules).
//
// All five matrix kernels are the same HBM-bound streaming GEMV (M=1):  y[n] = sum_k W[n,k] x[k], W row-major
// [N,K] bf16 streamed exactly once with TMA bulk copies (cp.async.bulk -> SASS UBLKCP) into per-warp
// shared-memory rings guarded by mbarriers; x (<= 56 KiB) lives in shared memory as bf16; fp32 accumulation,
// warp-shuffle reduction, one bf16 rounding per output.  They differ in prologue
// (RMSNorm of the residual stream vs plain copy) and epilogue (RoPE + KV append, residual add, SwiGLU, logits +
// argmax).  Tensor cores are deliberately not used at M=1 (≈2 FLOP/B; SURVEY.md §7 kernel notes).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace xb {

typedef __nv_bfloat16 bf16;

constexpr int NW = 8;                 // consumer(=producer) warps per GEMV CTA
constexpr int GEMV_THREADS = NW * 32;
constexpr int MAX_RING_STEPS = 4096;  // timestamp / token log capacity of one ring run

// Per-engine control block in device memory.
struct StepCtl {
  int pos;         // position (0-based) of the token processed by the current step
  int step;        // global ring step number of the current step
  int run_step0;   // ring step number at which the current ring run started (indexes the logs)
  int err;         // sticky: 1 = flag wait timed out
  int token;       // last-shard argmax of the most recent head kernel
  int pad[3];
};

// Inbox: lives in the RECEIVER's HBM; written by the previous shard over NVLink (or locally for a 1-shard ring).
struct InboxHdr {
  unsigned int flag_hidden[2];   // == step+1 when hidden[slot] holds the activation of ring step `step`
  unsigned int flag_token[2];    // == step+1 when token[slot] holds the token id feeding ring step `step`
  int token[2];
  int pad[26];                   // header is 128 bytes; bf16 hidden[2][D] follows
};
static_assert(sizeof(InboxHdr) == 128, "InboxHdr must be 128 bytes");

// ------------------------------------------------------------------------------------------------------------
// PTX helpers
// ------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// TMA 1-D bulk copy global -> shared, completion counted on an mbarrier (SASS: UBLKCP).
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar,
                                         uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
__device__ __forceinline__ uint64_t globaltimer_ns() {
  uint64_t t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }
__device__ __forceinline__ float bf2f(bf16 v) { return __bfloat162float(v); }
__device__ __forceinline__ bf16 f2bf(float v) { return __float2bfloat16_rn(v); }
__device__ __forceinline__ float round_bf(float v) { return __bfloat162float(__float2bfloat16_rn(v)); }

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  f[0] = bf_lo(v.x); f[1] = bf_hi(v.x); f[2] = bf_lo(v.y); f[3] = bf_hi(v.y);
  f[4] = bf_lo(v.z); f[5] = bf_hi(v.z); f[6] = bf_lo(v.w); f[7] = bf_hi(v.w);
}
__device__ __forceinline__ float dot8(const uint4& w, const float (&x)[8], float acc) {
  acc = fmaf(bf_lo(w.x), x[0], acc); acc = fmaf(bf_hi(w.x), x[1], acc);
  acc = fmaf(bf_lo(w.y), x[2], acc); acc = fmaf(bf_hi(w.y), x[3], acc);
  acc = fmaf(bf_lo(w.z), x[4], acc); acc = fmaf(bf_hi(w.z), x[5], acc);
  acc = fmaf(bf_lo(w.w), x[6], acc); acc = fmaf(bf_hi(w.w), x[7], acc);
  return acc;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Spin until *flag == want (acquire, system scope: the writer may be a peer GPU).  Bounded by a wall-clock
// timeout so a broken ring can never hang the GPU (sets ctl->err instead).
__device__ __forceinline__ bool wait_flag(const unsigned int* flag, unsigned int want, StepCtl* ctl) {
  if (ld_acquire_sys(flag) == want) return true;
  const uint64_t t0 = globaltimer_ns();
  while (ld_acquire_sys(flag) != want) {
    if (*reinterpret_cast<volatile int*>(&ctl->err)) return false;
    if (globaltimer_ns() - t0 > 4000000000ull) {  // 4 s
      ctl->err = 1;
      return false;
    }
    __nanosleep(40);
  }
  return true;
}

// ------------------------------------------------------------------------------------------------------------
// Completion of the LAST kernel of a step: the last CTA to finish publishes the step's result.
// ------------------------------------------------------------------------------------------------------------
struct StepEnd {
  unsigned int* ticket;        // device counter (library-owned), reset by the last CTA
  StepCtl* ctl;
  unsigned int* post_flag;     // base of the 2-entry flag array in the NEXT inbox (NULL: nothing to post)
  int post_step_offset;        // 0: result feeds the SAME ring step of the next shard (hidden); 1: feeds step+1 (token)
  int advance;                 // 1: ring mode -> ctl->pos++, ctl->step++ and log timestamps
  unsigned long long* t_out;   // [MAX_RING_STEPS] log (ring mode)
};

// Called by every CTA after its stores are done (all threads).  `peer` => stores went to another GPU.
__device__ __forceinline__ bool step_end_arrive(const StepEnd& se) {
  __shared__ int s_last;
  if (se.post_flag) __threadfence_system();   // every writer orders its own (possibly peer) stores first
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    unsigned int t = atomicAdd(se.ticket, 1u);
    s_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  return s_last != 0;
}
__device__ __forceinline__ void step_end_publish(const StepEnd& se) {  // thread 0 of the last CTA
  *se.ticket = 0;
  const int step = se.ctl->step;
  if (se.advance) {
    int li = step - se.ctl->run_step0;
    if (li >= 0 && li < MAX_RING_STEPS) se.t_out[li] = globaltimer_ns();
  }
  if (se.post_flag) {
    __threadfence_system();
    const int ps = step + se.post_step_offset;
    st_release_sys(se.post_flag + (ps & 1), (unsigned int)(ps + 1));
  }
  if (se.advance) {
    se.ctl->pos = se.ctl->pos + 1;
    se.ctl->step = step + 1;
  }
}

// ------------------------------------------------------------------------------------------------------------
// ingest: first kernel of a step.  Brings the step's input into the local residual stream x[D].
//   ring mode  (src_* NULL): wait for the inbox flag of ctl->step, read token/hidden from the inbox slot;
//   direct mode            : host gave pos; read token/hidden from a local staging buffer.
// Replaces tok_embeddings lookup (llm_utils.py:448-449) and the H2D of the received hidden state
// (sharded_inference_engine.py:254-258).
// ------------------------------------------------------------------------------------------------------------
struct IngestArgs {
  StepCtl* ctl;
  InboxHdr* inbox;                 // own inbox
  const bf16* embed;               // [V, D] (first shard) or NULL
  bf16* x;                         // residual stream [D]
  int D;
  int direct;                      // 1: direct mode
  int pos;                         // direct mode: position to set
  const int* src_token;            // direct mode, first shard
  const bf16* src_hidden;          // direct mode, other shards
  unsigned long long* t_in;        // ring-mode log
};

__global__ void __launch_bounds__(256) ingest_kernel(IngestArgs a) {
  __shared__ int s_tok;
  __shared__ int s_ok;
  const bf16* src = nullptr;
  if (a.direct) {
    if (threadIdx.x == 0) {
      a.ctl->pos = a.pos;
      s_tok = a.embed ? *a.src_token : 0;
      s_ok = 1;
    }
    __syncthreads();
    src = a.embed ? a.embed + (size_t)s_tok * a.D : a.src_hidden;
  } else {
    const int step = a.ctl->step;
    const int slot = step & 1;
    if (threadIdx.x == 0) {
      bool ok;
      if (a.embed) {
        ok = wait_flag(&a.inbox->flag_token[slot], (unsigned)(step + 1), a.ctl);
        s_tok = ok ? a.inbox->token[slot] : 0;
      } else {
        ok = wait_flag(&a.inbox->flag_hidden[slot], (unsigned)(step + 1), a.ctl);
        s_tok = 0;
      }
      s_ok = ok;
      int li = step - a.ctl->run_step0;
      if (li >= 0 && li < MAX_RING_STEPS) a.t_in[li] = globaltimer_ns();
    }
    __syncthreads();
    src = a.embed ? a.embed + (size_t)s_tok * a.D
                  : reinterpret_cast<const bf16*>(reinterpret_cast<const char*>(a.inbox) + sizeof(InboxHdr)) +
                        (size_t)slot * a.D;
  }
  const uint4* s4 = reinterpret_cast<const uint4*>(src);
  uint4* d4 = reinterpret_cast<uint4*>(a.x);
  for (int i = threadIdx.x; i < a.D / 8; i += blockDim.x) {
    uint4 v;
    // volatile-ish load: inbox data was written by a peer; the acquire above orders it, bypass L1 anyway
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(s4 + i));
    d4[i] = v;
  }
}

// ------------------------------------------------------------------------------------------------------------
// Streaming GEMV
// ------------------------------------------------------------------------------------------------------------
enum { PRO_NORM = 0, PRO_PLAIN = 1 };
enum { EPI_QKV = 0, EPI_RESID = 1, EPI_SWIGLU = 2, EPI_HEAD = 3 };

struct GemvArgs {
  const bf16* W;        // [N, K] row-major bf16 (device layout, see engine.cu)
  int N, K;
  int stages;           // ring depth per warp
  const bf16* x_in;     // PRO_NORM: residual stream [K]; PRO_PLAIN: activation [K]
  const bf16* norm_w;   // PRO_NORM: RMSNorm scale [K]
  float eps;
  StepCtl* ctl;
  // EPI_QKV
  bf16* q_out;          // [Hq*hd] (interleaved-pair order inside each head)
  bf16* k_cache;        // [Hkv, L, hd] of this layer
  bf16* v_cache;
  const float2* rope;   // [max_pos, hd/2] (cos, sin)
  int q_rows, kv_rows, hd, L;
  // EPI_RESID
  const bf16* res;      // residual input [N]
  bf16* out;            // output [N] (may alias res) — EPI_SWIGLU: act[N/2]
  bf16* out_peer;       // optional second destination: next inbox hidden base (2 slots of N) or NULL
  // EPI_HEAD
  float* logits;        // [N] fp32 (bf16 values)
  float* cta_best_val;  // [gridDim.x]
  int* cta_best_idx;    // [gridDim.x]
  int* tokens_out;      // [MAX_RING_STEPS] token log (ring mode)
  int* next_token;      // next (= first shard's) inbox token[2] base or NULL
  // completion (only used when `final_kernel`)
  int final_kernel;
  StepEnd end;
};

template <int TR, int TK, int PRO, int EPI>
__global__ void __launch_bounds__(GEMV_THREADS, 1) gemv_kernel(const GemvArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int STAGE_BYTES = TR * TK * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = a.K, S = a.stages;

  // smem carve-up: [ring NW*S*STAGE_BYTES][x_s K bf16][bars NW*S u64][red 32 f32]
  uint8_t* ring_base = smem;
  bf16* x_s = reinterpret_cast<bf16*>(smem + (size_t)NW * S * STAGE_BYTES);
  uint64_t* bar_base = reinterpret_cast<uint64_t*>(reinterpret_cast<uint8_t*>(x_s) + (size_t)K * 2);
  float* red = reinterpret_cast<float*>(bar_base + NW * S);

  // rows -> CTAs (contiguous groups of TR rows), groups -> warps (round robin), tiles = group x K-chunk
  const int ngroups = a.N / TR;
  const int gpc = ngroups / gridDim.x, grem = ngroups % gridDim.x;
  const int g0 = blockIdx.x * gpc + min((int)blockIdx.x, grem);
  const int gn = gpc + ((int)blockIdx.x < grem ? 1 : 0);
  const int KC = K / TK;
  const int my_groups = gn > warp ? (gn - warp + NW - 1) / NW : 0;
  const int ntiles = my_groups * KC;

  uint8_t* ring = ring_base + (size_t)warp * S * STAGE_BYTES;
  uint64_t* bars = bar_base + warp * S;
  const uint64_t policy = policy_evict_first();

  auto issue = [&](int i) {  // lane 0 only
    const int gl = warp + (i / KC) * NW;
    const int kc = i - (i / KC) * KC;
    const size_t row0 = (size_t)(g0 + gl) * TR;
    const int st = i % S;
    uint8_t* dst = ring + (size_t)st * STAGE_BYTES;
    mbar_expect_tx(&bars[st], STAGE_BYTES);
    if (KC == 1) {
      bulk_g2s(dst, a.W + row0 * K, STAGE_BYTES, &bars[st], policy);
    } else {
#pragma unroll
      for (int r = 0; r < TR; ++r)
        bulk_g2s(dst + r * TK * 2, a.W + (row0 + r) * K + (size_t)kc * TK, TK * 2, &bars[st], policy);
    }
  };

  if (lane == 0) {
    for (int s = 0; s < S; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
    const int n0 = ntiles < S ? ntiles : S;
    for (int i = 0; i < n0; ++i) issue(i);   // weights do not depend on the previous kernel: prefetch first
  }
  __syncwarp();

  // ---- prologue: build x_s (bf16) ----
  if (PRO == PRO_NORM) {
    // RMSNorm, torchtune<=0.6 / HF order: fp32 normalise -> bf16 -> * scale (bf16) -> bf16
    float ss = 0.f;
    for (int c = threadIdx.x; c < K / 8; c += GEMV_THREADS) {
      uint4 v = reinterpret_cast<const uint4*>(a.x_in)[c];
      reinterpret_cast<uint4*>(x_s)[c] = v;
      float f[8];
      unpack8(v, f);
#pragma unroll
      for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
    }
    ss = warp_sum(ss);
    if (lane == 0) red[warp] = ss;
    __syncthreads();
    float tot = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) tot += red[w];
    const float rs = 1.0f / sqrtf(tot / (float)K + a.eps);
    for (int c = threadIdx.x; c < K / 8; c += GEMV_THREADS) {
      uint4 v = reinterpret_cast<uint4*>(x_s)[c];
      uint4 g = reinterpret_cast<const uint4*>(a.norm_w)[c];
      float f[8], s[8];
      unpack8(v, f);
      unpack8(g, s);
      uint32_t o[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float y0 = round_bf(f[2 * j] * rs) * s[2 * j];
        float y1 = round_bf(f[2 * j + 1] * rs) * s[2 * j + 1];
        __nv_bfloat162 p = __floats2bfloat162_rn(y0, y1);
        o[j] = *reinterpret_cast<uint32_t*>(&p);
      }
      reinterpret_cast<uint4*>(x_s)[c] = make_uint4(o[0], o[1], o[2], o[3]);
    }
  } else {
    for (int c = threadIdx.x; c < K / 8; c += GEMV_THREADS)
      reinterpret_cast<uint4*>(x_s)[c] = reinterpret_cast<const uint4*>(a.x_in)[c];
  }
  __syncthreads();

  const int pos = a.ctl->pos;
  float best_v = -INFINITY;
  int best_i = 0x7fffffff;

  // ---- main loop: this warp's private pipeline ----
  float acc[TR];
#pragma unroll
  for (int r = 0; r < TR; ++r) acc[r] = 0.f;
  for (int i = 0; i < ntiles; ++i) {
    const int st = i % S;
    const uint32_t phase = (uint32_t)(i / S) & 1u;
    const int kc = i - (i / KC) * KC;
    mbar_wait(&bars[st], phase);
    const uint4* t4 = reinterpret_cast<const uint4*>(ring + (size_t)st * STAGE_BYTES);
    const uint4* x4 = reinterpret_cast<const uint4*>(x_s + (size_t)kc * TK);
#pragma unroll
    for (int c = 0; c < TK / 256; ++c) {
      float xf[8];
      unpack8(x4[c * 32 + lane], xf);
#pragma unroll
      for (int r = 0; r < TR; ++r) acc[r] = dot8(t4[r * (TK / 8) + c * 32 + lane], xf, acc[r]);
    }
    __syncwarp();
    if (lane == 0 && i + S < ntiles) issue(i + S);
    if (kc == KC - 1) {
      float v[TR];
#pragma unroll
      for (int r = 0; r < TR; ++r) {
        v[r] = warp_sum(acc[r]);
        acc[r] = 0.f;
      }
      const int row0 = (g0 + warp + (i / KC) * NW) * TR;
      // ---- epilogues (lane p handles pair p / row p) ----
      if (EPI == EPI_QKV) {
        // rounding points: Linear out -> bf16; RoPE in fp32 with separate mul/mul/sub (eager torch) -> bf16 (App. A.2, A.7)
        if (lane < TR / 2) {
          float v0 = lane == 0 ? v[0] : v[TR > 2 ? 2 : 0];
          float v1 = lane == 0 ? v[1] : v[TR > 2 ? 3 : 1];
          const int n = row0 + 2 * lane;
          const float x0 = round_bf(v0), x1 = round_bf(v1);
          if (n < a.q_rows + a.kv_rows) {
            const int i_in_head = n % a.hd;
            const float2 cs = a.rope[(size_t)pos * (a.hd / 2) + (i_in_head >> 1)];
            const float o0 = __fsub_rn(__fmul_rn(x0, cs.x), __fmul_rn(x1, cs.y));
            const float o1 = __fadd_rn(__fmul_rn(x1, cs.x), __fmul_rn(x0, cs.y));
            __nv_bfloat162 p = __floats2bfloat162_rn(o0, o1);
            if (n < a.q_rows) {
              *reinterpret_cast<__nv_bfloat162*>(a.q_out + n) = p;
            } else {
              const int nk = n - a.q_rows;
              const int h = nk / a.hd;
              *reinterpret_cast<__nv_bfloat162*>(a.k_cache + ((size_t)h * a.L + pos) * a.hd + i_in_head) = p;
            }
          } else {
            const int nv = n - a.q_rows - a.kv_rows;
            const int h = nv / a.hd, ih = nv % a.hd;
            __nv_bfloat162 p = __floats2bfloat162_rn(x0, x1);
            *reinterpret_cast<__nv_bfloat162*>(a.v_cache + ((size_t)h * a.L + pos) * a.hd + ih) = p;
          }
        }
      } else if (EPI == EPI_RESID) {
        // h = x + Linear(..): Linear out -> bf16, add in fp32 -> bf16 (App. A.5, A.7)
        if (lane < TR) {
          float vv = v[0];
#pragma unroll
          for (int r = 1; r < TR; ++r) vv = lane == r ? v[r] : vv;
          const int n = row0 + lane;
          const bf16 o = f2bf(bf2f(a.res[n]) + round_bf(vv));
          a.out[n] = o;
          if (a.out_peer) a.out_peer[(size_t)(a.ctl->step & 1) * a.N + n] = o;
        }
      } else if (EPI == EPI_SWIGLU) {
        // rows interleaved (2m: gate_m, 2m+1: up_m): act[m] = bf16( bf16(silu(bf16 g)) * bf16 u )  (llm_utils.py:491-500)
        if (lane < TR / 2) {
          float g = round_bf(lane == 0 ? v[0] : v[TR > 2 ? 2 : 0]);
          float u = round_bf(lane == 0 ? v[1] : v[TR > 2 ? 3 : 1]);
          const float sg = round_bf(g / (1.0f + expf(-g)));
          a.out[(row0 >> 1) + lane] = f2bf(sg * u);
        }
      } else {  // EPI_HEAD: logits (bf16-valued fp32) + running argmax (lowest index wins ties)
#pragma unroll
        for (int r = 0; r < TR; ++r) {
          const float l = round_bf(v[r]);
          if (lane == 0) a.logits[row0 + r] = l;
          if (l > best_v) { best_v = l; best_i = row0 + r; }   // rows ascend within a warp: first max kept
        }
      }
    }
  }

  if (EPI == EPI_HEAD) {
    // CTA argmax -> per-CTA slot; the last CTA reduces all slots (order independent: max with lowest index)
    __shared__ float s_bv[NW];
    __shared__ int s_bi[NW];
    if (lane == 0) { s_bv[warp] = best_v; s_bi[warp] = best_i; }
    __syncthreads();
    if (threadIdx.x == 0) {
      float bv = s_bv[0]; int bi = s_bi[0];
      for (int w = 1; w < NW; ++w)
        if (s_bv[w] > bv || (s_bv[w] == bv && s_bi[w] < bi)) { bv = s_bv[w]; bi = s_bi[w]; }
      a.cta_best_val[blockIdx.x] = bv;
      a.cta_best_idx[blockIdx.x] = bi;
    }
  }

  if (a.final_kernel) {
    const bool last = step_end_arrive(a.end);
    if (last && threadIdx.x == 0) {
      if (EPI == EPI_HEAD) {
        __threadfence();
        float bv = -INFINITY; int bi = 0x7fffffff;
        for (int c = 0; c < (int)gridDim.x; ++c) {
          const float cv = *reinterpret_cast<volatile float*>(a.cta_best_val + c);
          const int ci = *reinterpret_cast<volatile int*>(a.cta_best_idx + c);
          if (cv > bv || (cv == bv && ci < bi)) { bv = cv; bi = ci; }
        }
        a.ctl->token = bi;
        const int step = a.ctl->step;
        if (a.end.advance) {
          const int li = step - a.ctl->run_step0;
          if (li >= 0 && li < MAX_RING_STEPS) a.tokens_out[li] = bi;
        }
        if (a.next_token) a.next_token[(step + 1) & 1] = bi;   // feeds ring step+1 of the first shard (peer store)
      }
      step_end_publish(a.end);
    }
  }
}

// ------------------------------------------------------------------------------------------------------------
// GQA decode attention over the VALID prefix [0, pos] of a contiguous KV cache (no K/V head materialisation,
// no work on the unwritten tail: SURVEY.md §2.2 K5/K6, App. B7).  Split-K over the context with an online
// softmax per warp; the last CTA of a KV head (ticket) merges the splits in fixed order and writes bf16.
// Replaces KVCache.update + expand + F.scaled_dot_product_attention (App. A.3-A.4).
// grid = (nsplit, Hkv), block = 128.
// ------------------------------------------------------------------------------------------------------------
struct AttnArgs {
  const bf16* q;        // [Hq*hd]
  const bf16* k_cache;  // [Hkv, L, hd]
  const bf16* v_cache;
  bf16* out;            // [Hq*hd]
  float* part_ml;       // [Hkv*R][nsplit][2]
  float* part_acc;      // [Hkv*R][nsplit][hd]
  unsigned int* tickets;  // [Hkv]
  StepCtl* ctl;
  int L;
  float scale;
};

template <int HD, int R>
__global__ void __launch_bounds__(128) attn_decode_kernel(const AttnArgs a) {
  constexpr int E = HD / 32;
  constexpr int AW = 4;
  const int split = blockIdx.x, nsplit = gridDim.x, g = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int ctx = a.ctl->pos + 1;
  const int chunk = (ctx + nsplit - 1) / nsplit;
  const int t0 = split * chunk;
  const int t1 = min(ctx, t0 + chunk);

  float q[R][E], acc[R][E], m[R], l[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
    for (int e = 0; e < E; ++e) {
      q[r][e] = bf2f(a.q[(size_t)(g * R + r) * HD + lane * E + e]);
      acc[r][e] = 0.f;
    }
  }
  const bf16* kb = a.k_cache + (size_t)g * a.L * HD;
  const bf16* vb = a.v_cache + (size_t)g * a.L * HD;
  for (int t = t0 + warp; t < t1; t += AW) {
    float kf[E], vf[E];
#pragma unroll
    for (int e = 0; e < E; e += 2) {
      __nv_bfloat162 kk = *reinterpret_cast<const __nv_bfloat162*>(kb + (size_t)t * HD + lane * E + e);
      __nv_bfloat162 vv = *reinterpret_cast<const __nv_bfloat162*>(vb + (size_t)t * HD + lane * E + e);
      kf[e] = __low2float(kk); kf[e + 1] = __high2float(kk);
      vf[e] = __low2float(vv); vf[e + 1] = __high2float(vv);
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float s = 0.f;
#pragma unroll
      for (int e = 0; e < E; ++e) s = fmaf(q[r][e], kf[e], s);
      s = warp_sum(s) * a.scale;
      const float mn = fmaxf(m[r], s);
      const float corr = expf(m[r] - mn);
      const float p = expf(s - mn);
      l[r] = l[r] * corr + p;
#pragma unroll
      for (int e = 0; e < E; ++e) acc[r][e] = acc[r][e] * corr + p * vf[e];
      m[r] = mn;
    }
  }
  // merge the 4 warps in shared memory (fixed order)
  __shared__ float s_m[AW][R], s_l[AW][R];
  __shared__ float s_acc[AW][R][HD];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    if (lane == 0) { s_m[warp][r] = m[r]; s_l[warp][r] = l[r]; }
#pragma unroll
    for (int e = 0; e < E; ++e) s_acc[warp][r][lane * E + e] = acc[r][e];
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < R * HD; idx += 128) {
    const int r = idx / HD, d = idx % HD;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < AW; ++w) M = fmaxf(M, s_m[w][r]);
    float Lsum = 0.f, A = 0.f;
    if (M > -INFINITY) {
#pragma unroll
      for (int w = 0; w < AW; ++w) {
        const float f = expf(s_m[w][r] - M);   // exp(-inf) = 0 for idle warps
        Lsum += s_l[w][r] * f;
        A += s_acc[w][r][d] * f;
      }
    }
    const size_t hr = (size_t)(g * R + r);
    a.part_acc[(hr * nsplit + split) * HD + d] = A;
    if (d == 0) {
      a.part_ml[(hr * nsplit + split) * 2 + 0] = M;
      a.part_ml[(hr * nsplit + split) * 2 + 1] = Lsum;
    }
  }
  // ticket: last split CTA of this kv head merges
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    unsigned int t = atomicAdd(&a.tickets[g], 1u);
    s_last = (t == (unsigned)nsplit - 1);
    if (s_last) a.tickets[g] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int idx = threadIdx.x; idx < R * HD; idx += 128) {
    const int r = idx / HD, d = idx % HD;
    const size_t hr = (size_t)(g * R + r);
    const volatile float* ml = a.part_ml + hr * nsplit * 2;
    const volatile float* pa = a.part_acc + hr * nsplit * HD;
    float M = -INFINITY;
    for (int j = 0; j < nsplit; ++j) M = fmaxf(M, ml[j * 2]);
    float Lsum = 0.f, A = 0.f;
    for (int j = 0; j < nsplit; ++j) {
      const float mj = ml[j * 2];
      if (mj > -INFINITY) {
        const float f = expf(mj - M);
        Lsum += ml[j * 2 + 1] * f;
        A += pa[j * HD + d] * f;
      }
    }
    a.out[hr * HD + d] = f2bf(A / Lsum);
  }
}

// ------------------------------------------------------------------------------------------------------------
// small utility kernels
// ------------------------------------------------------------------------------------------------------------
__global__ void set_ctl_kernel(StepCtl* ctl, int pos, int step, int run_step0) {
  ctl->pos = pos; ctl->step = step; ctl->run_step0 = run_step0;
}
__global__ void f32_to_bf16_kernel(const float* in, bf16* out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = f2bf(in[i]);
}
__global__ void bf16_to_f32_kernel(const bf16* in, float* out, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    out[i] = bf2f(in[i]);
}
// post a hidden vector / token into the next inbox and release the flag (used once, after prefill)
__global__ void post_kernel(InboxHdr* next, const bf16* hidden, int D, const int* token, int step) {
  const int slot = step & 1;
  if (hidden) {
    bf16* dst = reinterpret_cast<bf16*>(reinterpret_cast<char*>(next) + sizeof(InboxHdr)) + (size_t)slot * D;
    for (int i = threadIdx.x; i < D; i += blockDim.x) dst[i] = hidden[i];
  } else if (threadIdx.x == 0) {
    next->token[slot] = *token;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    st_release_sys(hidden ? &next->flag_hidden[slot] : &next->flag_token[slot], (unsigned)(step + 1));
  }
}

// Row re-layout while loading: dst row r <- src row map(r)
//   mode 0: identity; mode 1: q/k interleave (_permute, llm_utils.py:175-183) with `hd`;
__global__ void relayout_rows_kernel(const bf16* src, bf16* dst, int rows, int K, int mode, int hd, int dst_stride_rows,
                                     int dst_row_off) {
  // dst row index = dst_row_off + r * dst_stride_rows  (stride 2 interleaves gate/up)
  const int r = blockIdx.x;
  if (r >= rows) return;
  int sr = r;
  if (mode == 1) {
    const int h = r / hd, i = r % hd;
    sr = h * hd + (i & 1) * (hd / 2) + (i >> 1);
  }
  const uint4* s4 = reinterpret_cast<const uint4*>(src + (size_t)sr * K);
  uint4* d4 = reinterpret_cast<uint4*>(dst + ((size_t)dst_row_off + (size_t)r * dst_stride_rows) * K);
  for (int c = threadIdx.x; c < K / 8; c += blockDim.x) d4[c] = s4[c];
}

// Counter-based synthetic weights, bit-identical to xotorch_b200/synth.py (splitmix64 -> Irwin-Hall(4)).
__device__ __forceinline__ float synth_value(unsigned long long base, unsigned long long idx, float scale, float mean) {
  unsigned long long z = base + idx;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  const int s = (int)(z & 0xffff) + (int)((z >> 16) & 0xffff) + (int)((z >> 32) & 0xffff) + (int)(z >> 48);
  return __fadd_rn(__fmul_rn((float)(s - 131070), scale), mean);
}
// dst row r (of `rows`) takes HF row map(r) of the tensor; same mapping as relayout_rows_kernel.
__global__ void synth_rows_kernel(bf16* dst, int rows, int K, int mode, int hd, int dst_stride_rows, int dst_row_off,
                                  unsigned long long base, float scale, float mean) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  int sr = r;
  if (mode == 1) {
    const int h = r / hd, i = r % hd;
    sr = h * hd + (i & 1) * (hd / 2) + (i >> 1);
  }
  bf16* d = dst + ((size_t)dst_row_off + (size_t)r * dst_stride_rows) * K;
  for (int c = threadIdx.x; c < K; c += blockDim.x)
    d[c] = f2bf(synth_value(base, (unsigned long long)sr * (unsigned long long)K + c, scale, mean));
}

}  // namespace xb
