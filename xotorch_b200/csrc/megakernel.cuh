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
constexpr int MK_STAGE_BYTES = 16384;     // one ring stage: TR rows x TK columns of bf16
constexpr int MK_NSTAGE = 12;             // CTA-level ring depth (192 KiB in flight per SM)
constexpr int MK_NSLOT = 4;               // partial-sum hand-off slots consumer warps -> epilogue warp

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

__device__ __forceinline__ OpInfo op_info(const MegaArgs& a, int op) {
  OpInfo o;
  if (op == 4 * a.nl) {
    o.W = a.lm_head; o.N = a.V; o.K = a.D; o.kind = OP_HEAD; o.li = a.nl;
    return o;
  }
  o.li = op >> 2;
  o.kind = op & 3;
  const LayerDev& l = a.layers[o.li];
  switch (o.kind) {
    case OP_QKV: o.W = l.wqkv; o.N = (a.Hq + 2 * a.Hkv) * a.hd; o.K = a.D; break;
    case OP_O: o.W = l.wo; o.N = a.D; o.K = a.Hq * a.hd; break;
    case OP_GU: o.W = l.wgu; o.N = 2 * a.F; o.K = a.D; break;
    default: o.W = l.wdown; o.N = a.D; o.K = a.F; break;
  }
  return o;
}

// rows -> CTAs: contiguous groups of TR rows, balanced to +-1 group
__device__ __forceinline__ void cta_groups(int N, int TR, int& g0, int& gn) {
  const int ngroups = N / TR;
  const int gpc = ngroups / gridDim.x, grem = ngroups % gridDim.x;
  g0 = blockIdx.x * gpc + min((int)blockIdx.x, grem);
  gn = gpc + ((int)blockIdx.x < grem ? 1 : 0);
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
template <int HD, int R>
__device__ __forceinline__ void attn_phase(const MegaArgs& a, const LayerDev& lw, int ctx, float* scratch) {
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
    const uint4 qv = *reinterpret_cast<const uint4*>(a.q + (size_t)(g * R + r) * HD + li * 8);
    unpack8(qv, q[r]);
    m[r] = -INFINITY; l[r] = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[r][e] = 0.f;
  }
  const bf16* kb = lw.k_cache + (size_t)g * a.L * HD + li * 8;
  const bf16* vb = lw.v_cache + (size_t)g * a.L * HD + li * 8;
  const int stride = NCW * PPW;
  int t = t0 + warp * PPW + sub;
  uint4 kn = make_uint4(0, 0, 0, 0), vn = kn;
  if (t < t1) {
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
// Ring stage = TR rows x TK columns (16 KiB).  Column block cb (256 columns) of row subset rs (4 rows) belongs to
// consumer warp rs*CB+cb, so EVERY consumer warp works on EVERY stage (K-split): perfect balance inside the CTA and
// one in-order weight stream per SM.  Per row group the 8 warps hand their partial sums to the epilogue warp
// through a 4-slot mbarrier mailbox; it adds them in fixed order (deterministic) and applies the fused epilogue.
//   TK=2048: CB=8, RS=1, TR=4   (Llama-3 sizes)        TK=512: CB=2, RS=4, TR=16   (tiny parity shapes)
// ------------------------------------------------------------------------------------------------------------
template <int TK, int HD, int R>
__global__ void __launch_bounds__(MK_THREADS, 1) shard_step_kernel(const MegaArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int CB = TK / 256;          // column blocks per stage
  constexpr int RS = NCW / CB;          // row subsets per stage
  constexpr int TR = 4 * RS;            // rows per stage
  static_assert(TR * TK * 2 == MK_STAGE_BYTES, "stage must be 16 KiB");
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // smem: [ring][xbuf][full NSTAGE][empty NSTAGE][done NSLOT][free NSLOT][part NSLOT*NCW*4 f32][red 32 f32]
  uint8_t* ring = smem;
  uint8_t* xbuf = smem + (size_t)MK_NSTAGE * MK_STAGE_BYTES;
  bf16* x_s = reinterpret_cast<bf16*>(xbuf);
  uint64_t* full = reinterpret_cast<uint64_t*>(xbuf + a.xbuf_bytes);
  uint64_t* empty = full + MK_NSTAGE;
  uint64_t* done = empty + MK_NSTAGE;
  uint64_t* freeb = done + MK_NSLOT;
  float* part = reinterpret_cast<float*>(freeb + MK_NSLOT);   // [NSLOT][NCW][4]
  float* red = part + MK_NSLOT * NCW * 4;
  __shared__ volatile int s_abort;
  __shared__ int s_tok;

  const int nops_all = 4 * a.nl + (a.last ? 1 : 0);
  if (threadIdx.x == 0) {
    s_abort = 0;
    for (int i = 0; i < MK_NSTAGE; ++i) { mbar_init(&full[i], 1); mbar_init(&empty[i], NCW); }
    for (int i = 0; i < MK_NSLOT; ++i) { mbar_init(&done[i], NCW); mbar_init(&freeb[i], 1); }
    fence_mbar_init();
  }
  __syncthreads();   // the only CTA-wide barrier: roles diverge below and never meet on barrier 0 again

  auto op_enabled = [&](int op, int step) -> bool {
    if (op < 4 * a.nl) return true;
    return a.last && (a.mode == 0 || step == a.n_steps - 1);
  };

  // =========================================== producer warp ===========================================
  // One cursor walks the tile list (step, op, row group, K chunk) of this CTA across ops and steps and feeds the
  // shared-memory ring with TMA bulk copies; it only ever blocks on a full ring, never on a grid barrier.
  // (An HBM->L2 prefetch cursor running ahead of the ring was tried and measured SLOWER on B200: every byte then
  //  crosses the L2 twice and the L2-slice throughput, not HBM, becomes the cap.)
  if (warp == W_PROD) {
    if (lane != 0) return;
    const uint64_t policy = policy_evict_first();
    struct Cur { int s, op, g, kc, g0, gn, KC, K; const bf16* W; bool end; };
    auto cur_seek = [&](Cur& c) {        // position on the first tile at or after (c.s, c.op)
      c.g = 0; c.kc = 0;
      while (true) {
        if (c.s >= a.n_steps) { c.end = true; return; }
        if (c.op >= nops_all) { c.op = 0; ++c.s; continue; }
        if (!op_enabled(c.op, c.s)) { ++c.op; continue; }
        const OpInfo o = op_info(a, c.op);
        cta_groups(o.N, TR, c.g0, c.gn);
        if (c.gn == 0) { ++c.op; continue; }
        c.W = o.W; c.K = o.K; c.KC = o.K / TK;
        return;
      }
    };
    auto cur_next = [&](Cur& c) {
      if (++c.kc == c.KC) { c.kc = 0; if (++c.g == c.gn) { ++c.op; cur_seek(c); } }
    };
    Cur ld{0, 0, 0, 0, 0, 0, 1, 0, nullptr, false};
    cur_seek(ld);
    unsigned int p_cnt = 0;
    bool aborted = false;
    while (!ld.end) {
      const int st = p_cnt % MK_NSTAGE;
      const uint32_t par = ((p_cnt / MK_NSTAGE) & 1u) ^ 1u;
      while (!mbar_try_wait(&empty[st], par)) {
        if (s_abort) { aborted = true; break; }
      }
      if (aborted) break;
      uint8_t* dst = ring + (size_t)st * MK_STAGE_BYTES;
      const size_t row0 = (size_t)(ld.g0 + ld.g) * TR;
      mbar_expect_tx(&full[st], MK_STAGE_BYTES);
      if (ld.KC == 1) {
        bulk_g2s(dst, ld.W + row0 * ld.K, MK_STAGE_BYTES, &full[st], policy);
      } else {
#pragma unroll 4
        for (int r = 0; r < TR; ++r)
          bulk_g2s(dst + r * TK * 2, ld.W + (row0 + r) * ld.K + (size_t)ld.kc * TK, TK * 2, &full[st], policy);
      }
      ++p_cnt;
      cur_next(ld);
    }
    // never exit with bulk copies in flight: wait for the completion of everything issued and not yet consumed
    const unsigned int lo = p_cnt > MK_NSTAGE ? p_cnt - MK_NSTAGE : 0;
    for (unsigned int i = lo; i < p_cnt; ++i) mbar_wait(&full[i % MK_NSTAGE], (i / MK_NSTAGE) & 1u);
    return;
  }

  // ================================== consumer warps 0..7 and epilogue warp ==================================
  const bool is_epi = (warp == W_EPI);
  const int cb = warp % CB, rs = warp / CB;      // consumer role
  unsigned int c_cnt = 0;                         // stages consumed (consumers)
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
      const OpInfo o = op_info(a, op);
      const LayerDev* lw = o.kind == OP_HEAD ? nullptr : &a.layers[o.li];
      const int K = o.K, KC = K / TK;
      int g0, gn;
      cta_groups(o.N, TR, g0, gn);
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
          // RMSNorm, torchtune<=0.6 / HF order (App. A.1): fp32 normalise -> bf16 -> * scale (bf16) -> bf16
          float ss = 0.f;
          const bool seed_x = (o.kind == OP_QKV && o.li == 0 && blockIdx.x == 0);   // CTA 0 materialises the residual stream
          for (int c = threadIdx.x; c < K / 8; c += CONS_THREADS) {
            const uint4 v = __ldcg(xin4 + c);
            reinterpret_cast<uint4*>(x_s)[c] = v;
            if (seed_x) reinterpret_cast<uint4*>(a.x)[c] = v;
            float f[8];
            unpack8(v, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) ss = fmaf(f[j], f[j], ss);
          }
          ss = warp_sum(ss);
          if (lane == 0) red[warp] = ss;
          bar_sync(1, CONS_THREADS);
          float tot = 0.f;
#pragma unroll
          for (int w = 0; w < NCW; ++w) tot += red[w];
          const float rstd = 1.0f / sqrtf(tot / (float)K + a.eps);
          for (int c = threadIdx.x; c < K / 8; c += CONS_THREADS) {
            const uint4 v = reinterpret_cast<uint4*>(x_s)[c];
            const uint4 g = __ldg(reinterpret_cast<const uint4*>(nw) + c);
            float f[8], sc[8];
            unpack8(v, f);
            unpack8(g, sc);
            uint32_t ov[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float y0 = round_bf(f[2 * j] * rstd) * sc[2 * j];
              const float y1 = round_bf(f[2 * j + 1] * rstd) * sc[2 * j + 1];
              __nv_bfloat162 p2 = __floats2bfloat162_rn(y0, y1);
              ov[j] = *reinterpret_cast<uint32_t*>(&p2);
            }
            reinterpret_cast<uint4*>(x_s)[c] = make_uint4(ov[0], ov[1], ov[2], ov[3]);
          }
        } else {
          for (int c = threadIdx.x; c < K / 8; c += CONS_THREADS) reinterpret_cast<uint4*>(x_s)[c] = __ldcg(xin4 + c);
        }
        bar_sync(1, CONS_THREADS);
        if (dbg_on) {
          a.dbg[op * 4 + 1] = globaltimer_ns();
          int landed = 0;    // ring stages already filled when consumption of this op starts
          for (int i = 0; i < MK_NSTAGE; ++i)
            landed += mbar_try_wait(&full[(c_cnt + i) % MK_NSTAGE], ((c_cnt + i) / MK_NSTAGE) & 1u) ? 1 : 0;
          a.dbg[(nops_all + 1 + op) * 4] = (unsigned long long)landed;
        }
        // ---------------- consume: every warp takes its 4 rows x 256 columns of every stage ----------------
        // Software pipelined: the registers of tile t+1 are loaded (and its stage released) before the FMAs of
        // tile t, so mbarrier-wait + LDS latency of the next stage hides behind the math of the current one.
        const int T = gn * KC;
        const uint4* xs4 = reinterpret_cast<const uint4*>(x_s) + cb * 32 + lane;
        const uint4* ring4 = reinterpret_cast<const uint4*>(ring) + (size_t)(rs * 4) * (TK / 8) + cb * 32 + lane;
        auto load_tile = [&](unsigned int cnt, int kc, uint4 (&w)[4], uint4& xv) {
          const int st = cnt % MK_NSTAGE;
          mbar_wait(&full[st], (cnt / MK_NSTAGE) & 1u);
          const uint4* t4 = ring4 + (size_t)st * (MK_STAGE_BYTES / 16);
#pragma unroll
          for (int r = 0; r < 4; ++r) w[r] = t4[r * (TK / 8)];
          xv = xs4[kc * (TK / 8)];
          __syncwarp();
          if (lane == 0) mbar_arrive(&empty[st]);   // release: the tile now lives in registers
        };
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        auto finish_group = [&]() {
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[r] = warp_sum(acc[r]);
          if (lane == 0) {
            const int ps = g_cnt % MK_NSLOT;
            mbar_wait(&freeb[ps], ((g_cnt / MK_NSLOT) & 1u) ^ 1u);
            *reinterpret_cast<float4*>(part + (ps * NCW + warp) * 4) = make_float4(acc[0], acc[1], acc[2], acc[3]);
            mbar_arrive(&done[ps]);
          }
          ++g_cnt;
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[r] = 0.f;
        };
        auto compute = [&](const uint4 (&w)[4], const uint4& xv) {
          float xf[8];
          unpack8(xv, xf);
#pragma unroll
          for (int r = 0; r < 4; ++r) acc[r] = dot8(w[r], xf, acc[r]);
        };
        uint4 wA[4], wB[4], xA, xB;
        int kcA = 0;                       // K-chunk index of the tile held in A
        if (T > 0) load_tile(c_cnt, 0, wA, xA);
        for (int t = 0; t < T; t += 2) {
          // ---- tile t in A; prefetch t+1 into B ----
          const int kcB = (kcA + 1 == KC) ? 0 : kcA + 1;
          if (t + 1 < T) load_tile(c_cnt + 1, kcB, wB, xB);
          compute(wA, xA);
          ++c_cnt;
          if (kcA == KC - 1) finish_group();
          if (t + 1 >= T) break;
          // ---- tile t+1 in B; prefetch t+2 into A ----
          kcA = (kcB + 1 == KC) ? 0 : kcB + 1;
          if (t + 2 < T) load_tile(c_cnt + 1, kcA, wA, xA);
          compute(wB, xB);
          ++c_cnt;
          if (kcB == KC - 1) finish_group();
        }
      } else {
        // ---------------- epilogue warp: reduce the 8 partials per row in fixed order, fused epilogues ----------------
        // Operands the epilogue needs from global memory (residual pair for O/DOWN, RoPE cos/sin for QKV) are
        // preloaded for ALL row pairs of this CTA at op start — lane i holds pair i — while the consumers are still in
        // their prologue; per group they are redistributed with a shuffle.  (A dependent L2 load per group would
        // throttle the whole CTA to one group per ~0.7 us through the mailbox back-pressure.)
        const int npairs = gn * (TR / 2);
        const bool pre_ok = npairs <= 32;
        const int q_rows = a.Hq * a.hd, kv_rows = a.Hkv * a.hd;
        float2 pre = make_float2(0.f, 0.f);
        if (pre_ok && lane < npairs) {
          const int np = g0 * TR + 2 * lane;
          if (o.kind == OP_O || o.kind == OP_DOWN) {
            const __nv_bfloat162 rv = *reinterpret_cast<const __nv_bfloat162*>(a.x + np);
            pre = make_float2(__low2float(rv), __high2float(rv));
          } else if (o.kind == OP_QKV && np < q_rows + kv_rows) {
            pre = a.rope[(size_t)pos * (a.hd / 2) + ((np % a.hd) >> 1)];
          }
        }
        for (int g = 0; g < gn; ++g) {
          const float2 pg = make_float2(__shfl_sync(0xffffffffu, pre.x, (g * (TR / 2) + lane) & 31),
                                        __shfl_sync(0xffffffffu, pre.y, (g * (TR / 2) + lane) & 31));
          const int ps = g_cnt % MK_NSLOT;
          mbar_wait(&done[ps], (g_cnt / MK_NSLOT) & 1u);
          float v0 = 0.f, v1 = 0.f;
          const int prs = (2 * lane) / 4, pr0 = (2 * lane) % 4;     // lane p owns the row pair (2p, 2p+1) of the group
          if (lane < TR / 2) {
#pragma unroll
            for (int c = 0; c < CB; ++c) {
              const float2 pv = *reinterpret_cast<const float2*>(part + (ps * NCW + prs * CB + c) * 4 + pr0);
              v0 += pv.x; v1 += pv.y;
            }
          }
          ++g_cnt;
          __syncwarp();
          if (lane == 0) mbar_arrive(&freeb[ps]);
          if (lane < TR / 2) {
            const int n = (g0 + g) * TR + 2 * lane;
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
      ok = grid_barrier(a, bar_target, &s_abort);
      if (!ok) break;
      if (o.kind == OP_QKV) {
        if (!is_epi) attn_phase<HD, R>(a, *lw, pos + 1, reinterpret_cast<float*>(xbuf));
        ok = grid_barrier(a, bar_target, &s_abort);
        if (!ok) break;
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
