// megakernel.cuh — the persistent shard-step kernel (sm_100a).
//
// ONE launch runs `n_steps` consecutive steps (tokens) of a shard: for every step
//     ingest -> { qkv -> attention -> o-proj -> gate/up -> down } x layers [-> lm_head + argmax] -> publish
// replacing the reference's per-token Python/eager path (SURVEY.md §3.3;
// /root/reference/xotorch/inference/llm_utils.py:416-489, general_mha.py:178-254, node.py:109-147,424-443).
//
// Why one kernel: at batch 1 every matrix is read exactly once per token, so the step is HBM-bound, but one
// matrix is only 5-35 us of HBM time and a kernel boundary costs several us of drained memory pipe.  Here the
// weight stream never stops:
//   * grid = one CTA per SM, 8 warps; every warp owns a private ring of S shared-memory stages fed by TMA bulk
//     copies (cp.async.bulk, mbarrier complete_tx) and walks ONE tile list that spans all ops, layers and steps.
//     A stage is refilled as soon as it is consumed, so while a CTA waits at a grid barrier or does attention its
//     next weights are already in flight (weights never depend on activations).
//   * ops are separated by grid barriers (release/acquire on a monotonic counter); activations (<= 56 KiB) travel
//     through L2.
//   * the inter-shard hop is fused: the down-proj epilogue of a shard's last layer stores the bf16 hidden state
//     straight into the NEXT shard's inbox over NVLink (peer stores) and CTA 0 releases a flag; the next shard's
//     CTAs acquire that flag before their first RMSNorm.  The last shard's lm_head epilogue produces the argmax
//     token and stores it into shard 0's inbox the same way (ring closure), so no host code runs per token.
#pragma once
#include "kernels.cuh"

namespace xb {

constexpr int NCW = 8;                    // consumer warps (also run prologues, attention, grid barriers)
constexpr int W_PROD = 8;                 // TMA producer warp
constexpr int W_EPI = 9;                  // epilogue warp (cross-warp reduction + fused epilogues)
constexpr int MK_THREADS = 320;
constexpr int CONS_THREADS = NCW * 32;    // 256
constexpr int MK_NSLOT = 4;               // partial-sum hand-off slots consumer warps -> epilogue warp
// ring stage = mk_rows(tk) weight rows (whole RoPE / gate-up pairs) x tk columns: 16 KiB x 12 stages (Llama sizes, measured
// 1.4 % faster than 32 KiB x 6) or 8 KiB x 12 (tiny parity shapes)
__host__ __device__ constexpr int mk_rows(int tk) { return tk >= 2048 ? 4 : 8; }
__host__ __device__ constexpr int mk_stage_bytes(int tk) { return mk_rows(tk) * tk * 2; }
__host__ __device__ constexpr int mk_nstage(int tk) { return 12; }   // default ring depth (8 for very wide models, see engine.cu)


__device__ __forceinline__ void bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

struct LayerDev {
  const bf16 *attn_norm, *wqkv, *wo, *mlp_norm, *wgu, *wdown;
  bf16 *k_cache, *v_cache;   // [Hkv, L, hd] each
};

struct MegaArgs {
  const LayerDev* layers;
  int nl, D, F, Hq, Hkv, hd, V, L;
  float eps, attn_scale;
  const float2* rope;
  const bf16 *embed, *final_norm, *lm_head;
  int first, last;
  bf16 *x, *q, *attn_out, *act;
  float* logits;
  float *part_ml, *part_acc;
  unsigned int* tickets;     // [Hkv]
  float* cta_best_val;
  int* cta_best_idx;
  int nsplit;
  StepCtl* ctl;
  InboxHdr *inbox, *next_inbox;
  int* tokens_log;
  unsigned long long *t_in, *t_out;
  unsigned int* bar_counter; // grid barrier counter (host zeroes it before every launch)
  int mode;                  // 0 ring, 1 direct
  int n_steps, step0, pos0;
  const int* src_tokens;     // direct, first shard: [n_steps]
  const bf16* src_hidden;    // direct, other shards: [n_steps, D]
  bf16* out_hidden;          // direct, non-last shard: [n_steps, D]
  int xbuf_bytes;            // shared-memory activation buffer size
  unsigned long long* dbg;   // optional timeline [n_ops+1][4] of CTA dbg_cta for step dbg_step (NULL = off)
  int dbg_step, dbg_cta;
};

enum { OP_QKV = 0, OP_O = 1, OP_GU = 2, OP_DOWN = 3, OP_HEAD = 4 };

struct OpInfo {
  const bf16* W;
  int N, K, kind, li;
};

__device__ __forceinline__ OpInfo op_info(const MegaArgs& a, const LayerDev* layers, int op) {
  OpInfo o;
  if (op == 4 * a.nl) {
    o.W = a.lm_head; o.N = a.V; o.K = a.D; o.kind = OP_HEAD; o.li = a.nl;
    return o;
  }
  o.li = op >> 2;
  o.kind = op & 3;
  const LayerDev& l = layers[o.li];
  switch (o.kind) {
    case OP_QKV: o.W = l.wqkv; o.N = (a.Hq + 2 * a.Hkv) * a.hd; o.K = a.D; break;
    case OP_O: o.W = l.wo; o.N = a.D; o.K = a.Hq * a.hd; break;
    case OP_GU: o.W = l.wgu; o.N = 2 * a.F; o.K = a.D; break;
    default: o.W = l.wdown; o.N = a.D; o.K = a.F; break;
  }
  return o;
}

// acc (packed fp32x2) += 8 bf16 weights (one 16-byte shared-memory load) * 8 x values (4 packed fp32x2): 8 unpack ops
// + 4 FFMA2 (Blackwell packed fp32 FMA).  Lane-local; the two halves of acc are added at the end of the row group.
__device__ __forceinline__ float2 dot8_f2(const uint4& w, const float2 (&x)[4], float2 acc) {
  acc = __ffma2_rn(make_float2(bf_lo(w.x), bf_hi(w.x)), x[0], acc);
  acc = __ffma2_rn(make_float2(bf_lo(w.y), bf_hi(w.y)), x[1], acc);
  acc = __ffma2_rn(make_float2(bf_lo(w.z), bf_hi(w.z)), x[2], acc);
  acc = __ffma2_rn(make_float2(bf_lo(w.w), bf_hi(w.w)), x[3], acc);
  return acc;
}

// Transpose-reduce NV per-lane values over the 32 lanes of a warp: NV-1 exchange shuffles + log2(32/NV) butterfly
// steps (9 shuffles for NV=8 instead of 40).  Afterwards lane L holds the warp total of value index L / (32/NV).
template <int NV>
__device__ __forceinline__ float warp_reduce_rows(float (&v)[NV], int lane) {
  int off = 16;
#pragma unroll
  for (int n = NV; n > 1; n >>= 1) {
    const int half = n >> 1;
    const bool upper = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float keep = upper ? v[i + half] : v[i];
      const float send = upper ? v[i] : v[i + half];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
#pragma unroll
  for (; off > 0; off >>= 1) v[0] += __shfl_xor_sync(0xffffffffu, v[0], off);
  return v[0];
}

// rows -> CTAs: contiguous ranges of row PAIRS (RoPE / gate-up pairs stay together), balanced to +-1 pair
__device__ __forceinline__ void cta_pairs(int N, int& p0, int& pn) {
  const int npairs = N / 2;
  const int ppc = npairs / gridDim.x, prem = npairs % gridDim.x;
  p0 = blockIdx.x * ppc + min((int)blockIdx.x, prem);
  pn = ppc + ((int)blockIdx.x < prem ? 1 : 0);
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

// Grid-wide barrier on a monotonic counter.  Returns false on abort (peer timeout / error flag).
// Participants: the 8 consumer warps + the epilogue warp (named barrier 2); the producer warp never stops streaming.
__device__ __forceinline__ bool grid_barrier(const MegaArgs& a, unsigned int& target, volatile int* s_abort) {
  bar_sync(2, CONS_THREADS + 32);
  target += gridDim.x;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(a.bar_counter, 1u);
    if (ld_acquire_gpu(a.bar_counter) < target) {
      const uint64_t t0 = globaltimer_ns();
      while (ld_acquire_gpu(a.bar_counter) < target) {
        if (*reinterpret_cast<volatile int*>(&a.ctl->err) || globaltimer_ns() - t0 > 6000000000ull) {
          a.ctl->err = 2;
          *s_abort = 1;
          break;
        }
      }
    }
  }
  bar_sync(2, CONS_THREADS + 32);
  return *s_abort == 0;
}

// ------------------------------------------------------------------------------------------------------------
// attention phase (consumer warps only): one (kv head, context split) item per CTA, 8 warps, 16-byte K/V loads (HD/8 lanes per position)
// ------------------------------------------------------------------------------------------------------------
// K/V registers of the first position each lane group handles; loaded BEFORE the grid barrier that closes the QKV op
// when that position was written by an earlier step (only position ctx-1 is new), so the L2 latency overlaps the barrier.
struct AttnPre {
  uint4 k, v;
  bool have;
};

template <int HD>
__device__ __forceinline__ AttnPre attn_prefetch(const MegaArgs& a, const LayerDev& lw, int ctx) {
  constexpr int LPP = HD / 8;
  constexpr int PPW = 32 / LPP;
  AttnPre pre;
  pre.k = make_uint4(0, 0, 0, 0); pre.v = pre.k; pre.have = false;
  const int nsplit = a.nsplit;
  const int item = blockIdx.x;
  if (item >= a.Hkv * nsplit) return pre;
  const int g = item / nsplit, split = item - g * nsplit;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPP, li = lane % LPP;
  const int chunk = (ctx + nsplit - 1) / nsplit;
  const int t0 = split * chunk;
  const int t1 = min(ctx, t0 + chunk);
  const int t = t0 + warp * PPW + sub;
  if (t < t1 && t < ctx - 1) {
    pre.k = __ldcg(reinterpret_cast<const uint4*>(lw.k_cache + ((size_t)g * a.L + t) * HD + li * 8));
    pre.v = __ldcg(reinterpret_cast<const uint4*>(lw.v_cache + ((size_t)g * a.L + t) * HD + li * 8));
    pre.have = true;
  }
  return pre;
}

template <int HD, int R>
__device__ __forceinline__ void attn_phase(const MegaArgs& a, const LayerDev& lw, int ctx, float* scratch, const AttnPre& pre,
                                           unsigned long long* stamp) {
  constexpr int LPP = HD / 8;        // lanes per position
  constexpr int PPW = 32 / LPP;      // positions per warp load
  const int nsplit = a.nsplit;
  const int item = blockIdx.x;
  if (item >= a.Hkv * nsplit) return;
  const int g = item / nsplit, split = item - g * nsplit;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int sub = lane / LPP, li = lane % LPP;
  const int chunk = (ctx + nsplit - 1) / nsplit;
  const int t0 = split * chunk;
  const int t1 = min(ctx, t0 + chunk);

  float q[R][8], acc[R][8], m[R], l[R];
#pragma unroll
  for (int r = 0; r < R; ++r) {
    const uint4 qv = __ldcg(reinterpret_cast<const uint4*>(a.q + (size_t)(g * R + r) * HD + li * 8));
    unpack8(qv, q[r]);
    m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[r][e] = 0.f;
  }
  const bf16* kb = lw.k_cache + (size_t)g * a.L * HD + li * 8;
  const bf16* vb = lw.v_cache + (size_t)g * a.L * HD + li * 8;
  const int stride = NCW * PPW;
  int t = t0 + warp * PPW + sub;
  uint4 kn = pre.k, vn = pre.v;
  if (t < t1 && !pre.have) {
    kn = __ldcg(reinterpret_cast<const uint4*>(kb + (size_t)t * HD));
    vn = __ldcg(reinterpret_cast<const uint4*>(vb + (size_t)t * HD));
  }
  // all lanes of a warp run the same number of iterations (shuffles inside): iterate on the warp's base position
  for (int tb = t0 + warp * PPW; tb < t1; tb += stride, t += stride) {
    const uint4 kc = kn, vc = vn;
    const bool valid = t < t1;
    const int tn = t + stride;
    if (tn < t1) {
      kn = __ldcg(reinterpret_cast<const uint4*>(kb + (size_t)tn * HD));
      vn = __ldcg(reinterpret_cast<const uint4*>(vb + (size_t)tn * HD));
    }
    float vf[8];
    unpack8(vc, vf);
#pragma unroll
    for (int r = 0; r < R; ++r) {
      float s = dot8(kc, q[r], 0.f);
#pragma unroll
      for (int o = LPP / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (valid) {
        s *= a.attn_scale;
        const float mn = fmaxf(m[r], s);
        const float corr = expf(m[r] - mn);
        const float p = expf(s - mn);
        l[r] = l[r] * corr + p;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[r][e] = acc[r][e] * corr + p * vf[e];
        m[r] = mn;
      }
    }
  }
  if (stamp && threadIdx.x == 0) stamp[1] = globaltimer_ns();
  // merge the PPW lane groups of the warp (fixed order), then the 8 warps through shared memory
#pragma unroll
  for (int o = 16; o >= LPP; o >>= 1) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float mo = __shfl_xor_sync(0xffffffffu, m[r], o);
      const float lo = __shfl_xor_sync(0xffffffffu, l[r], o);
      const float M = fmaxf(m[r], mo);
      const float fa = (m[r] == -INFINITY) ? 0.f : expf(m[r] - M);
      const float fb = (mo == -INFINITY) ? 0.f : expf(mo - M);
      // keep the combination order identical on both partners: lower lane group first
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
  float* s_ml = scratch;                         // [NCW][R][2]
  float* s_acc = scratch + NCW * R * 2;           // [NCW][R][HD]
  if (sub == 0) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (li == 0) { s_ml[(warp * R + r) * 2] = m[r]; s_ml[(warp * R + r) * 2 + 1] = l[r]; }
#pragma unroll
      for (int e = 0; e < 8; ++e) s_acc[(warp * R + r) * HD + li * 8 + e] = acc[r][e];
    }
  }
  bar_sync(1, CONS_THREADS);
  for (int idx = threadIdx.x; idx < R * HD; idx += CONS_THREADS) {
    const int r = idx / HD, d = idx - r * HD;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < NCW; ++w) M = fmaxf(M, s_ml[(w * R + r) * 2]);
    float Ls = 0.f, A = 0.f;
#pragma unroll
    for (int w = 0; w < NCW; ++w) {
      const float mw = s_ml[(w * R + r) * 2];
      const float f = (mw == -INFINITY) ? 0.f : expf(mw - M);
      Ls += s_ml[(w * R + r) * 2 + 1] * f;
      A += s_acc[(w * R + r) * HD + d] * f;
    }
    const size_t hr = (size_t)(g * R + r);
    a.part_acc[(hr * nsplit + split) * HD + d] = A;
    if (d == 0) {
      a.part_ml[(hr * nsplit + split) * 2] = M;
      a.part_ml[(hr * nsplit + split) * 2 + 1] = Ls;
    }
  }
  if (stamp && threadIdx.x == 0) stamp[2] = globaltimer_ns();
  // ticket: the last split CTA of this kv head merges all splits (fixed order) and writes bf16
  __shared__ int s_last;
  bar_sync(1, CONS_THREADS);
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int tk = atomicAdd(&a.tickets[g], 1u);
    s_last = (tk == (unsigned)nsplit - 1);
    if (s_last) a.tickets[g] = 0;
  }
  bar_sync(1, CONS_THREADS);
  if (stamp && threadIdx.x == 0) stamp[3] = globaltimer_ns();
  if (!s_last) return;
  __threadfence();
  for (int idx = threadIdx.x; idx < R * HD; idx += CONS_THREADS) {
    const int r = idx / HD, d = idx - r * HD;
    const size_t hr = (size_t)(g * R + r);
    const float* ml = a.part_ml + hr * nsplit * 2;
    const float* pa = a.part_acc + hr * nsplit * HD + d;
    float M = -INFINITY;
    for (int j = 0; j < nsplit; ++j) M = fmaxf(M, __ldcg(ml + j * 2));
    float Ls = 0.f, A = 0.f;
#pragma unroll 6
    for (int j = 0; j < nsplit; ++j) {
      const float mj = __ldcg(ml + j * 2);
      const float f = (mj == -INFINITY) ? 0.f : expf(mj - M);
      Ls += __ldcg(ml + j * 2 + 1) * f;
      A += __ldcg(pa + (size_t)j * HD) * f;
    }
    a.attn_out[hr * HD + d] = f2bf(A / Ls);
  }
}

// ------------------------------------------------------------------------------------------------------------
// the kernel: 10 warps = 8 consumers + 1 TMA producer + 1 epilogue
//
// Ring stage = 8 weight rows x TK columns.  A CTA owns a contiguous range of row pairs and walks it in groups of 8
// rows (the last group of an op may be shorter: only its real rows are loaded).  Every consumer warp works on EVERY
// stage (K-split): column block cb (256 columns) of row subset rs belongs to warp rs*CB+cb — perfect balance inside
// the CTA and one in-order weight stream per SM.  TK=2048: 8 warps x (8 rows x 256 cols); TK=512 (tiny parity
// shapes): 4 row subsets x 2 column blocks, 2 rows per warp.  The inner loop is deliberately lean (the consumers must
// drain the ring several times faster than HBM fills it, or stalls at barriers are never recovered): one 16-byte
// shared load + 8 unpack + 4 packed FFMA2 per 8 weights, the x chunk converted once per stage, a 9-shuffle
// transpose-reduce per row group.  Partials go to the epilogue warp through a 4-slot mbarrier mailbox; it adds them
// in fixed order (deterministic) and applies the fused epilogue.
// ------------------------------------------------------------------------------------------------------------
template <int TK, int HD, int R, int NSTAGE>
__global__ void __launch_bounds__(MK_THREADS, 1) shard_step_kernel(const MegaArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int MK_ROWS = mk_rows(TK);            // weight rows per ring stage
  constexpr int ROWB = TK * 2;                    // bytes per stage row
  constexpr int STAGE_BYTES = mk_stage_bytes(TK);
  constexpr int CB = TK / 256;                    // column blocks per stage (a lane covers 8 columns, a warp 256)
  constexpr int RS = NCW / CB;                    // row subsets per stage
  constexpr int RPW = MK_ROWS / RS;               // rows per consumer warp per stage
  static_assert(CB * RS == NCW && RPW * RS == MK_ROWS, "TK must be 256, 512, 1024 or 2048");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // smem: [ring][xbuf][full NSTAGE][empty NSTAGE][done NSLOT][free NSLOT][part NSLOT*NCW*8 f32][red 32 f32]
  uint8_t* ring = smem;
  uint8_t* xbuf = smem + (size_t)NSTAGE * STAGE_BYTES;
  bf16* x_s = reinterpret_cast<bf16*>(xbuf);
  uint64_t* full = reinterpret_cast<uint64_t*>(xbuf + a.xbuf_bytes);
  uint64_t* empty = full + NSTAGE;
  uint64_t* done = empty + NSTAGE;
  uint64_t* freeb = done + MK_NSLOT;
  float* part = reinterpret_cast<float*>(freeb + MK_NSLOT);   // [NSLOT][NCW][8]
  float* red = part + MK_NSLOT * NCW * 8;
  LayerDev* s_layers = reinterpret_cast<LayerDev*>(red + 32);   // layer table copy: no L2 round trip per op
  __shared__ volatile int s_abort;
  __shared__ int s_tok;

  const int nops_all = 4 * a.nl + (a.last ? 1 : 0);
  if (threadIdx.x == 0) {
    s_abort = 0;
    for (int i = 0; i < NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NCW); }
    for (int i = 0; i < MK_NSLOT; ++i) { mbar_init(&done[i], NCW); mbar_init(&freeb[i], 1); }
    fence_mbar_init();
  }
  for (int i = threadIdx.x; i < a.nl * (int)(sizeof(LayerDev) / 8); i += MK_THREADS)
    reinterpret_cast<unsigned long long*>(s_layers)[i] = reinterpret_cast<const unsigned long long*>(a.layers)[i];
  __syncthreads();   // the only CTA-wide barrier: roles diverge below and never meet on barrier 0 again

  auto op_enabled = [&](int op, int step) -> bool {
    if (op < 4 * a.nl) return true;
    return a.last && (a.mode == 0 || step == a.n_steps - 1);
  };

  // =========================================== producer warp ===========================================
  // One cursor walks the tile list (step, op, row group, K chunk) of this CTA across ops and steps and feeds the
  // shared-memory ring with ONE TMA bulk copy per stage (K-panel weight layout); it only ever blocks on a full ring,
  // never on a grid barrier.  (HBM->L2 prefetch ahead of the ring — always, or only while the ring is full — was tried
  // and measured SLOWER on B200: every byte then crosses the L2 twice and L2-slice throughput becomes the cap.)
  if (warp == W_PROD) {
    if (lane != 0) return;
    const uint64_t policy = policy_evict_first();
    // tile cursor: (step, op, group, K chunk) of this CTA, flattened index `idx`
    struct Cur { int s, op, g, kc, p0, pn, ng, KC, K, N; const bf16* W; unsigned int idx; bool end; };
    auto cur_seek = [&](Cur& c) {        // position on the first tile at or after (c.s, c.op)
      c.g = 0; c.kc = 0;
      while (true) {
        if (c.s >= a.n_steps) { c.end = true; return; }
        if (c.op >= nops_all) { c.op = 0; ++c.s; continue; }
        if (!op_enabled(c.op, c.s)) { ++c.op; continue; }
        const OpInfo o = op_info(a, s_layers, c.op);
        cta_pairs(o.N, c.p0, c.pn);
        c.ng = (c.pn + MK_ROWS / 2 - 1) / (MK_ROWS / 2);
        if (c.ng == 0) { ++c.op; continue; }
        c.W = o.W; c.K = o.K; c.N = o.N; c.KC = o.K / TK;
        return;
      }
    };
    auto cur_next = [&](Cur& c) {
      ++c.idx;
      if (++c.kc == c.KC) { c.kc = 0; if (++c.g == c.ng) { ++c.op; cur_seek(c); } }
    };
    Cur ld{0, 0, 0, 0, 0, 0, 0, 1, 0, 0, nullptr, 0u, false};
    cur_seek(ld);
    bool aborted = false;
    while (!ld.end) {
      const unsigned int p_cnt = ld.idx;
      const int st = p_cnt % NSTAGE;
      const uint32_t par = ((p_cnt / NSTAGE) & 1u) ^ 1u;
      while (!mbar_try_wait(&empty[st], par)) {
        if (s_abort) { aborted = true; break; }
      }
      if (aborted) break;
      uint8_t* dst = ring + (size_t)st * STAGE_BYTES;
      const size_t row0 = (size_t)(ld.p0 + (MK_ROWS / 2) * ld.g) * 2;
      const int nrows = min(MK_ROWS, 2 * (ld.pn - (MK_ROWS / 2) * ld.g));
      // K-panel layout [K/TK][N][TK]: the nrows x TK stage is contiguous in HBM -> ONE bulk copy of up to 32 KiB
      mbar_expect_tx(&full[st], (uint32_t)nrows * TK * 2);
      bulk_g2s(dst, ld.W + ((size_t)ld.kc * ld.N + row0) * TK, (uint32_t)nrows * TK * 2, &full[st], policy);
      cur_next(ld);
    }
    // never exit with bulk copies in flight: wait for the completion of everything issued and not yet consumed
    const unsigned int p_end = ld.idx;
    const unsigned int lo = p_end > (unsigned)NSTAGE ? p_end - NSTAGE : 0;
    for (unsigned int i = lo; i < p_end; ++i) mbar_wait(&full[i % NSTAGE], (i / NSTAGE) & 1u);
    return;
  }

  // ================================== consumer warps 0..7 and epilogue warp ==================================
  const bool is_epi = (warp == W_EPI);
  const int cb = warp % CB, rs = warp / CB;       // consumer role: column block, row subset
  int c_st = 0;                                   // ring position of the next stage to consume (consumers)
  uint32_t c_ph = 0;
  unsigned int g_cnt = 0;                         // row groups handed over (consumers) / finished (epilogue)
  unsigned int bar_target = 0;
  bool ok = true;

  for (int s = 0; s < a.n_steps && ok; ++s) {
    const int pos = a.pos0 + s;
    const int step = a.step0 + s;
    const int slot = step & 1;
    // ---------------- ingest: where does this step's input come from? ----------------
    if (threadIdx.x == 0) {
      if (a.mode == 1) {
        s_tok = a.first ? a.src_tokens[s] : 0;
      } else {
        bool got;
        if (a.first) {
          got = wait_flag(&a.inbox->flag_token[slot], (unsigned)(step + 1), a.ctl);
          s_tok = got ? *reinterpret_cast<volatile int*>(&a.inbox->token[slot]) : 0;
        } else {
          got = wait_flag(&a.inbox->flag_hidden[slot], (unsigned)(step + 1), a.ctl);
          s_tok = 0;
        }
        if (!got) s_abort = 1;
        if (blockIdx.x == 0 && s < MAX_RING_STEPS) a.t_in[s] = globaltimer_ns();
      }
    }
    bar_sync(2, CONS_THREADS + 32);
    if (s_abort) { ok = false; break; }
    const bf16* x_src;
    if (a.first) x_src = a.embed + (size_t)s_tok * a.D;
    else if (a.mode == 1) x_src = a.src_hidden + (size_t)s * a.D;
    else x_src = reinterpret_cast<const bf16*>(reinterpret_cast<const char*>(a.inbox) + sizeof(InboxHdr)) + (size_t)slot * a.D;

    const bool head_on = op_enabled(4 * a.nl, s);
    const int nops = 4 * a.nl + (head_on ? 1 : 0);
    float best_v = -INFINITY;   // epilogue warp: per-lane running argmax of the head op
    int best_i = 0x7fffffff;
    const bool dbg_on = a.dbg && s == a.dbg_step && (int)blockIdx.x == a.dbg_cta && threadIdx.x == 0;

    for (int op = 0; op < nops && ok; ++op) {
      const OpInfo o = op_info(a, s_layers, op);
      const LayerDev* lw = o.kind == OP_HEAD ? nullptr : &s_layers[o.li];
      const int K = o.K, KC = K / TK;
      int p0, pn;
      cta_pairs(o.N, p0, pn);
      const int ngroups = (pn + MK_ROWS / 2 - 1) / (MK_ROWS / 2);
      if (dbg_on) a.dbg[op * 4 + 0] = globaltimer_ns();

      if (!is_epi) {
        // ---------------- prologue (consumers): activation vector -> shared memory (bf16) ----------------
        const bf16* xin;
        const bf16* nw = nullptr;
        if (o.kind == OP_QKV) { xin = (o.li == 0) ? x_src : a.x; nw = lw->attn_norm; }
        else if (o.kind == OP_O) xin = a.attn_out;
        else if (o.kind == OP_GU) { xin = a.x; nw = lw->mlp_norm; }
        else if (o.kind == OP_DOWN) xin = a.act;
        else { xin = a.x; nw = a.final_norm; }
        const uint4* xin4 = reinterpret_cast<const uint4*>(xin);
        if (nw) {
          // RMSNorm, torchtune<=0.6 / HF order (App. A.1): fp32 normalise -> bf16 -> * scale (bf16) -> bf16.
          // x and the norm weights are fetched together into registers (ONE L2 round trip on the critical path: the ring
          // is full and HBM idles while the consumers sit in this prologue), normalised from registers, stored once.
          constexpr int PC = 4;                                   // chunks of 8 per thread held in registers: K <= 8192
          const bool seed_x = (o.kind == OP_QKV && o.li == 0 && blockIdx.x == 0);   // CTA 0 materialises the residual stream
          const int nch = K / 8;
          uint4 xv[PC], gv[PC];
          float ss = 0.f;
#pragma unroll
          for (int j = 0; j < PC; ++j) {
            const int c = threadIdx.x + j * CONS_THREADS;
            if (c < nch) { xv[j] = __ldcg(xin4 + c); gv[j] = __ldg(reinterpret_cast<const uint4*>(nw) + c); }
          }
#pragma unroll
          for (int j = 0; j < PC; ++j) {
            const int c = threadIdx.x + j * CONS_THREADS;
            if (c < nch) {
              if (seed_x) reinterpret_cast<uint4*>(a.x)[c] = xv[j];
              float f[8];
              unpack8(xv[j], f);
#pragma unroll
              for (int i = 0; i < 8; ++i) ss = fmaf(f[i], f[i], ss);
            }
          }
          for (int c = threadIdx.x + PC * CONS_THREADS; c < nch; c += CONS_THREADS) {   // K > 8192: spill through x_s
            const uint4 v = __ldcg(xin4 + c);
            reinterpret_cast<uint4*>(x_s)[c] = v;
            if (seed_x) reinterpret_cast<uint4*>(a.x)[c] = v;
            float f[8];
            unpack8(v, f);
#pragma unroll
            for (int i = 0; i < 8; ++i) ss = fmaf(f[i], f[i], ss);
          }
          ss = warp_sum(ss);
          if (lane == 0) red[warp] = ss;
          bar_sync(1, CONS_THREADS);
          float tot = 0.f;
#pragma unroll
          for (int w = 0; w < NCW; ++w) tot += red[w];
          const float rstd = 1.0f / sqrtf(tot / (float)K + a.eps);
          auto norm_store = [&](int c, const uint4& v, const uint4& g) {
            float f[8], sc[8];
            unpack8(v, f);
            unpack8(g, sc);
            uint32_t ov[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float y0 = round_bf(f[2 * i] * rstd) * sc[2 * i];
              const float y1 = round_bf(f[2 * i + 1] * rstd) * sc[2 * i + 1];
              __nv_bfloat162 p2 = __floats2bfloat162_rn(y0, y1);
              ov[i] = *reinterpret_cast<uint32_t*>(&p2);
            }
            reinterpret_cast<uint4*>(x_s)[c] = make_uint4(ov[0], ov[1], ov[2], ov[3]);
          };
#pragma unroll
          for (int j = 0; j < PC; ++j) {
            const int c = threadIdx.x + j * CONS_THREADS;
            if (c < nch) norm_store(c, xv[j], gv[j]);
          }
          for (int c = threadIdx.x + PC * CONS_THREADS; c < nch; c += CONS_THREADS)
            norm_store(c, reinterpret_cast<uint4*>(x_s)[c], __ldg(reinterpret_cast<const uint4*>(nw) + c));
        } else {
          for (int c = threadIdx.x; c < K / 8; c += CONS_THREADS) reinterpret_cast<uint4*>(x_s)[c] = __ldcg(xin4 + c);
        }
        bar_sync(1, CONS_THREADS);
        if (dbg_on) {
          a.dbg[op * 4 + 1] = globaltimer_ns();
          int landed = 0;    // ring stages already filled when consumption of this op starts
          for (int i = 0; i < NSTAGE; ++i) {
            const int si = (c_st + i) % NSTAGE;
            landed += mbar_try_wait(&full[si], si >= c_st ? c_ph : (c_ph ^ 1u)) ? 1 : 0;
          }
          a.dbg[(nops_all + 1 + op) * 4] = (unsigned long long)landed;
        }
        // ---------------- consume: every warp takes its RPW rows x 256 columns of every stage ----------------
        // Software pipeline over the CTA's tile list of this op: the registers of tile t+1 are loaded before the math of
        // tile t is issued and its stage is released after that math (the loads have long completed, so the release
        // never stalls); the row-group reduction + mailbox hand-off of group g overlaps the loads of group g+1.
        const uint8_t* wlane = ring + (size_t)(rs * RPW) * ROWB + (size_t)(cb * 256 + lane * 8) * 2;
        const bf16* xlane = x_s + cb * 256 + lane * 8;
        const int T = ngroups * KC;
        float2 acc[RPW];
#pragma unroll
        for (int r = 0; r < RPW; ++r) acc[r] = make_float2(0.f, 0.f);
        int ld_st = c_st;                  // ring slot of the tile whose registers were loaded last
        auto load_tile = [&](int kc, uint4 (&w)[RPW], uint4& xv) {
          mbar_wait(&full[c_st], c_ph);
          const uint8_t* wp = wlane + (size_t)c_st * STAGE_BYTES;
#pragma unroll
          for (int r = 0; r < RPW; ++r) w[r] = *reinterpret_cast<const uint4*>(wp + r * ROWB);
          xv = *reinterpret_cast<const uint4*>(xlane + (size_t)kc * TK);
          ld_st = c_st;
          if (++c_st == NSTAGE) { c_st = 0; c_ph ^= 1u; }
        };
        auto release = [&](int st) {       // all lanes' loads of that stage are complete by now
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[st]);
        };
        auto compute = [&](const uint4 (&w)[RPW], const uint4& xv) {
          float2 xf[4];
          xf[0] = make_float2(bf_lo(xv.x), bf_hi(xv.x)); xf[1] = make_float2(bf_lo(xv.y), bf_hi(xv.y));
          xf[2] = make_float2(bf_lo(xv.z), bf_hi(xv.z)); xf[3] = make_float2(bf_lo(xv.w), bf_hi(xv.w));
#pragma unroll
          for (int r = 0; r < RPW; ++r) acc[r] = dot8_f2(w[r], xf, acc[r]);
        };
        auto finish_group = [&]() {
          float v[RPW];
#pragma unroll
          for (int r = 0; r < RPW; ++r) { v[r] = acc[r].x + acc[r].y; acc[r] = make_float2(0.f, 0.f); }
          const int ps = g_cnt % MK_NSLOT;
          const uint32_t fpar = ((g_cnt / MK_NSLOT) & 1u) ^ 1u;
          const float tot = warp_reduce_rows<RPW>(v, lane);       // lane L: total of row L / (32/RPW)
          mbar_wait(&freeb[ps], fpar);                            // all lanes poll (uniform): no divergent wait
          if ((lane & (32 / RPW - 1)) == 0) part[(ps * NCW + warp) * 8 + lane / (32 / RPW)] = tot;
          __syncwarp();
          if (lane == 0) mbar_arrive(&done[ps]);
          ++g_cnt;
        };
        uint4 wA[RPW], wB[RPW], xA, xB;
        int kcA = 0, stA = 0, stB = 0;
        if (T > 0) { load_tile(0, wA, xA); stA = ld_st; }
        for (int t = 0; t < T; t += 2) {
          const int kcB = (kcA + 1 == KC) ? 0 : kcA + 1;
          if (t + 1 < T) { load_tile(kcB, wB, xB); stB = ld_st; }
          compute(wA, xA);
          release(stA);
          if (dbg_on && op >= 20 && op < 24 && t < 60) a.dbg[(2 * nops_all + 2) * 4 + (op - 20) * 64 + t] = globaltimer_ns();
          if (kcA == KC - 1) finish_group();
          if (t + 1 >= T) break;
          kcA = (kcB + 1 == KC) ? 0 : kcB + 1;
          if (t + 2 < T) { load_tile(kcA, wA, xA); stA = ld_st; }
          compute(wB, xB);
          release(stB);
          if (dbg_on && op >= 20 && op < 24 && t + 1 < 60) a.dbg[(2 * nops_all + 2) * 4 + (op - 20) * 64 + t + 1] = globaltimer_ns();
          if (kcB == KC - 1) finish_group();
        }
      } else {
        // ---------------- epilogue warp: reduce the 8 partials per row in fixed order, fused epilogues ----------------
        // Operands the epilogue needs from global memory (residual pair for O/DOWN, RoPE cos/sin for QKV) are
        // preloaded for the row pairs of this CTA at op start — lane i holds pairs i and i+32 — while the consumers are
        // still in their prologue; per group they are redistributed with a shuffle.  (A dependent L2 load per group
        // would throttle the whole CTA to one group per ~0.7 us through the mailbox back-pressure.)
        const bool pre_ok = pn <= 64;
        const int q_rows = a.Hq * a.hd, kv_rows = a.Hkv * a.hd;
        float2 pre[2] = {make_float2(0.f, 0.f), make_float2(0.f, 0.f)};
        if (pre_ok && (o.kind == OP_O || o.kind == OP_DOWN || o.kind == OP_QKV)) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int pi = lane + 32 * h;
            if (pi < pn) {
              const int np = (p0 + pi) * 2;
              if (o.kind != OP_QKV) {
                const __nv_bfloat162 rv = *reinterpret_cast<const __nv_bfloat162*>(a.x + np);
                pre[h] = make_float2(__low2float(rv), __high2float(rv));
              } else if (np < q_rows + kv_rows) {
                pre[h] = a.rope[(size_t)pos * (a.hd / 2) + ((np % a.hd) >> 1)];
              }
            }
          }
        }
        for (int g = 0; g < ngroups; ++g) {
          const int pi = (MK_ROWS / 2) * g + (lane & 3);        // pair index inside the CTA handled by lanes 0..MK_ROWS/2-1
          const float2 s0 = make_float2(__shfl_sync(0xffffffffu, pre[0].x, pi & 31), __shfl_sync(0xffffffffu, pre[0].y, pi & 31));
          const float2 s1 = make_float2(__shfl_sync(0xffffffffu, pre[1].x, pi & 31), __shfl_sync(0xffffffffu, pre[1].y, pi & 31));
          const float2 pg = (pi & 32) ? s1 : s0;
          const int ps = g_cnt % MK_NSLOT;
          mbar_wait(&done[ps], (g_cnt / MK_NSLOT) & 1u);
          float v0 = 0.f, v1 = 0.f;
          if (lane < MK_ROWS / 2) {
            const int prs = (2 * lane) / RPW, pr0 = (2 * lane) % RPW;     // row subset / row inside the subset of pair `lane`
#pragma unroll
            for (int c = 0; c < CB; ++c) {
              const float2 pv = *reinterpret_cast<const float2*>(part + (ps * NCW + prs * CB + c) * 8 + pr0);
              v0 += pv.x; v1 += pv.y;
            }
          }
          ++g_cnt;
          __syncwarp();
          if (lane == 0) mbar_arrive(&freeb[ps]);
          if (lane < MK_ROWS / 2 && pi < pn) {
            const int n = (p0 + pi) * 2;
            if (o.kind == OP_QKV) {
              // Linear out -> bf16; RoPE in fp32 with separate mul/mul/sub like eager torch -> bf16 (App. A.2, A.7)
              const float x0 = round_bf(v0), x1 = round_bf(v1);
              if (n < q_rows + kv_rows) {
                const int ih = n % a.hd;
                const float2 cs = pre_ok ? pg : a.rope[(size_t)pos * (a.hd / 2) + (ih >> 1)];
                const float o0 = __fsub_rn(__fmul_rn(x0, cs.x), __fmul_rn(x1, cs.y));
                const float o1 = __fadd_rn(__fmul_rn(x1, cs.x), __fmul_rn(x0, cs.y));
                const __nv_bfloat162 p2 = __floats2bfloat162_rn(o0, o1);
                if (n < q_rows) {
                  *reinterpret_cast<__nv_bfloat162*>(a.q + n) = p2;
                } else {
                  const int nk = n - q_rows;
                  *reinterpret_cast<__nv_bfloat162*>(lw->k_cache + ((size_t)(nk / a.hd) * a.L + pos) * a.hd + ih) = p2;
                }
              } else {
                const int nv = n - q_rows - kv_rows;
                *reinterpret_cast<__nv_bfloat162*>(lw->v_cache + ((size_t)(nv / a.hd) * a.L + pos) * a.hd + (nv % a.hd)) =
                    __floats2bfloat162_rn(x0, x1);
              }
            } else if (o.kind == OP_O || o.kind == OP_DOWN) {
              // h = x + Linear(..): Linear out -> bf16, add in fp32 -> bf16 (App. A.5, A.7)
              float r0, r1;
              if (pre_ok) { r0 = pg.x; r1 = pg.y; }
              else {
                const __nv_bfloat162 rv = *reinterpret_cast<const __nv_bfloat162*>(a.x + n);
                r0 = __low2float(rv); r1 = __high2float(rv);
              }
              const __nv_bfloat162 p2 = __floats2bfloat162_rn(r0 + round_bf(v0), r1 + round_bf(v1));
              *reinterpret_cast<__nv_bfloat162*>(a.x + n) = p2;
              if (o.kind == OP_DOWN && o.li == a.nl - 1 && !a.last) {
                if (a.mode == 0) {   // fused hop: peer store into the next shard's inbox slot over NVLink
                  bf16* dst = reinterpret_cast<bf16*>(reinterpret_cast<char*>(a.next_inbox) + sizeof(InboxHdr)) + (size_t)slot * a.D;
                  *reinterpret_cast<__nv_bfloat162*>(dst + n) = p2;
                } else {
                  *reinterpret_cast<__nv_bfloat162*>(a.out_hidden + (size_t)s * a.D + n) = p2;
                }
              }
            } else if (o.kind == OP_GU) {
              // rows interleaved (gate_m, up_m): act[m] = bf16( bf16(silu(bf16 g)) * bf16 u )  (llm_utils.py:491-500)
              const float gg = round_bf(v0), uu = round_bf(v1);
              const float sg = round_bf(gg / (1.0f + expf(-gg)));
              a.act[n >> 1] = f2bf(sg * uu);
            } else {
              // logits (bf16-valued fp32) + running argmax; groups ascend so the first maximum is the lowest index
              const float l0 = round_bf(v0), l1 = round_bf(v1);
              *reinterpret_cast<float2*>(a.logits + n) = make_float2(l0, l1);
              if (l0 > best_v) { best_v = l0; best_i = n; }
              if (l1 > best_v) { best_v = l1; best_i = n + 1; }
            }
          }
        }
        if (o.kind == OP_HEAD) {   // CTA argmax: lowest index among equal maxima
#pragma unroll
          for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(0xffffffffu, best_v, off);
            const int oi = __shfl_xor_sync(0xffffffffu, best_i, off);
            if (ov > best_v || (ov == best_v && oi < best_i)) { best_v = ov; best_i = oi; }
          }
          if (lane == 0) { a.cta_best_val[blockIdx.x] = best_v; a.cta_best_idx[blockIdx.x] = best_i; }
        }
        if (o.kind == OP_DOWN && o.li == a.nl - 1 && !a.last && a.mode == 0) __threadfence_system();  // order the peer stores
      }
      if (dbg_on) a.dbg[op * 4 + 2] = globaltimer_ns();
      // ---------------- op boundary ----------------
      const bool hop_out = (o.kind == OP_DOWN && o.li == a.nl - 1 && !a.last && a.mode == 0);
      AttnPre apre;
      apre.have = false;
      if (o.kind == OP_QKV && !is_epi) apre = attn_prefetch<HD>(a, *lw, pos + 1);   // old K/V: overlap the barrier
      ok = grid_barrier(a, bar_target, &s_abort);
      if (!ok) break;
      if (o.kind == OP_QKV) {
        unsigned long long* stamp = (dbg_on && op == 20) ? a.dbg + (2 * nops_all + 2) * 4 + 48 : nullptr;
        if (stamp) stamp[0] = globaltimer_ns();
        if (!is_epi) attn_phase<HD, R>(a, *lw, pos + 1, reinterpret_cast<float*>(xbuf), apre, stamp);
        if (stamp) stamp[4] = globaltimer_ns();
        ok = grid_barrier(a, bar_target, &s_abort);
        if (!ok) break;
        if (stamp) stamp[5] = globaltimer_ns();
      }
      if (dbg_on) a.dbg[op * 4 + 3] = globaltimer_ns();
      // ---------------- publish (CTA 0, after the barrier that closes the producing op) ----------------
      if (blockIdx.x == 0 && threadIdx.x == 0) {
        if (hop_out) {
          if (s < MAX_RING_STEPS) a.t_out[s] = globaltimer_ns();
          __threadfence_system();
          st_release_sys(&a.next_inbox->flag_hidden[slot], (unsigned)(step + 1));
        } else if (o.kind == OP_HEAD) {
          float bv = -INFINITY; int bi = 0x7fffffff;
          for (int c = 0; c < (int)gridDim.x; ++c) {
            const float cv = __ldcg(a.cta_best_val + c);
            const int ci = __ldcg(a.cta_best_idx + c);
            if (cv > bv || (cv == bv && ci < bi)) { bv = cv; bi = ci; }
          }
          a.ctl->token = bi;
          if (a.mode == 0) {
            if (s < MAX_RING_STEPS) { a.tokens_log[s] = bi; a.t_out[s] = globaltimer_ns(); }
            const int ns = step + 1;               // the token feeds ring step+1 of the first shard
            *reinterpret_cast<volatile int*>(&a.next_inbox->token[ns & 1]) = bi;
            __threadfence_system();
            st_release_sys(&a.next_inbox->flag_token[ns & 1], (unsigned)(ns + 1));
          }
        }
      }
    }
  }
  if (!ok && threadIdx.x == 0) s_abort = 1;   // let the producer drain and leave
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.ctl->pos = a.pos0 + a.n_steps;
    a.ctl->step = a.step0 + a.n_steps;
  }
}

}  // namespace xb
