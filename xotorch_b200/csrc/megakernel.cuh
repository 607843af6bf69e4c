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

constexpr int MK_TR = 2;     // weight rows per tile (RoPE pair / gate-up pair)
constexpr int MK_S = 3;      // ring stages per warp

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

// rows -> CTAs: contiguous groups of MK_TR rows, balanced to +-1 group
__device__ __forceinline__ void cta_groups(int N, int& g0, int& gn) {
  const int ngroups = N / MK_TR;
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
__device__ __forceinline__ bool grid_barrier(const MegaArgs& a, unsigned int& target, int* s_abort) {
  __syncthreads();
  if (threadIdx.x == 0) {
    target += gridDim.x;
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
  __syncthreads();
  return *reinterpret_cast<volatile int*>(s_abort) == 0;
}

// ------------------------------------------------------------------------------------------------------------
// attention phase: one (kv head, context split) item per CTA, 8 warps, 16-byte K/V loads (HD/8 lanes per position)
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
  const int stride = NW * PPW;
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
  float* s_ml = scratch;                         // [NW][R][2]
  float* s_acc = scratch + NW * R * 2;           // [NW][R][HD]
  if (sub == 0) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      if (li == 0) { s_ml[(warp * R + r) * 2] = m[r]; s_ml[(warp * R + r) * 2 + 1] = l[r]; }
#pragma unroll
      for (int e = 0; e < 8; ++e) s_acc[(warp * R + r) * HD + li * 8 + e] = acc[r][e];
    }
  }
  __syncthreads();
  for (int idx = threadIdx.x; idx < R * HD; idx += GEMV_THREADS) {
    const int r = idx / HD, d = idx - r * HD;
    float M = -INFINITY;
#pragma unroll
    for (int w = 0; w < NW; ++w) M = fmaxf(M, s_ml[(w * R + r) * 2]);
    float Ls = 0.f, A = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) {
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
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int tk = atomicAdd(&a.tickets[g], 1u);
    s_last = (tk == (unsigned)nsplit - 1);
    if (s_last) a.tickets[g] = 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int idx = threadIdx.x; idx < R * HD; idx += GEMV_THREADS) {
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
// the kernel
// ------------------------------------------------------------------------------------------------------------
template <int TK, int HD, int R>
__global__ void __launch_bounds__(GEMV_THREADS, 1) shard_step_kernel(const MegaArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  constexpr int STAGE_BYTES = MK_TR * TK * 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // smem: [ring NW*S*STAGE][xbuf xbuf_bytes][bars NW*S u64][red 32 f32][misc]
  uint8_t* ring = smem + (size_t)warp * MK_S * STAGE_BYTES;
  uint8_t* xbuf = smem + (size_t)NW * MK_S * STAGE_BYTES;
  bf16* x_s = reinterpret_cast<bf16*>(xbuf);
  uint64_t* bars = reinterpret_cast<uint64_t*>(xbuf + a.xbuf_bytes) + warp * MK_S;
  float* red = reinterpret_cast<float*>(xbuf + a.xbuf_bytes + NW * MK_S * 8);
  __shared__ int s_abort;
  __shared__ int s_tok;
  __shared__ float s_bv[NW];
  __shared__ int s_bi[NW];

  const int nops_all = 4 * a.nl + (a.last ? 1 : 0);
  const uint64_t policy = policy_evict_first();
  if (threadIdx.x == 0) s_abort = 0;
  if (lane == 0) {
    for (int s = 0; s < MK_S; ++s) mbar_init(&bars[s], 1);
    fence_mbar_init();
  }
  __syncthreads();

  // ---- prefetch cursor (uniform per warp): walks (step, op, group, kchunk) over the whole launch ----
  int p_step = 0, p_op = 0, p_grp = warp, p_kc = 0, p_g0 = 0, p_gn = 0, p_KC = 1;
  const bf16* p_W = nullptr;
  int p_K = 0;
  bool p_done = false;
  unsigned int p_cnt = 0, c_cnt = 0;   // tiles issued / consumed by this warp
  auto op_enabled = [&](int op, int step) -> bool {
    if (op < 4 * a.nl) return true;
    return a.last && (a.mode == 0 || step == a.n_steps - 1);
  };
  auto p_load_op = [&]() {   // position the cursor on the first tile of (p_step, p_op) or later
    while (true) {
      if (p_step >= a.n_steps) { p_done = true; return; }
      if (p_op >= nops_all) { p_op = 0; ++p_step; continue; }
      if (!op_enabled(p_op, p_step)) { ++p_op; continue; }
      const OpInfo o = op_info(a, p_op);
      cta_groups(o.N, p_g0, p_gn);
      if (warp >= p_gn) { ++p_op; continue; }     // this warp has no group in this op
      p_W = o.W; p_K = o.K; p_KC = o.K / TK; p_grp = warp; p_kc = 0;
      return;
    }
  };
  auto p_issue = [&]() {     // issue the tile under the cursor, advance
    if (lane == 0) {
      const int st = p_cnt % MK_S;
      uint8_t* dst = ring + (size_t)st * STAGE_BYTES;
      const size_t row0 = (size_t)(p_g0 + p_grp) * MK_TR;
      mbar_expect_tx(&bars[st], STAGE_BYTES);
      if (p_KC == 1) {
        bulk_g2s(dst, p_W + row0 * p_K, STAGE_BYTES, &bars[st], policy);
      } else {
#pragma unroll
        for (int r = 0; r < MK_TR; ++r)
          bulk_g2s(dst + r * TK * 2, p_W + (row0 + r) * p_K + (size_t)p_kc * TK, TK * 2, &bars[st], policy);
      }
    }
    ++p_cnt;
    if (++p_kc == p_KC) {
      p_kc = 0;
      p_grp += NW;
      if (p_grp >= p_gn) { ++p_op; p_load_op(); }
    }
  };
  p_load_op();
  for (int i = 0; i < MK_S && !p_done; ++i) p_issue();   // weights do not depend on activations: start streaming now

  unsigned int bar_target = 0;
  bool ok = true;

  for (int s = 0; s < a.n_steps && ok; ++s) {
    const int pos = a.pos0 + s;
    const int step = a.step0 + s;
    const int slot = step & 1;
    // ---------------- ingest: where does this step's input come from? ----------------
    const bf16* x_src;
    if (a.mode == 1) {
      if (a.first) {
        if (threadIdx.x == 0) s_tok = a.src_tokens[s];
        __syncthreads();
        x_src = a.embed + (size_t)s_tok * a.D;
      } else {
        x_src = a.src_hidden + (size_t)s * a.D;
      }
    } else {
      if (threadIdx.x == 0) {
        bool got;
        if (a.first) {
          got = wait_flag(&a.inbox->flag_token[slot], (unsigned)(step + 1), a.ctl);
          s_tok = got ? *reinterpret_cast<volatile int*>(&a.inbox->token[slot]) : 0;
        } else {
          got = wait_flag(&a.inbox->flag_hidden[slot], (unsigned)(step + 1), a.ctl);
        }
        if (!got) s_abort = 1;
        if (blockIdx.x == 0 && s < MAX_RING_STEPS) a.t_in[s] = globaltimer_ns();
      }
      __syncthreads();
      if (s_abort) { ok = false; break; }
      x_src = a.first ? a.embed + (size_t)s_tok * a.D
                      : reinterpret_cast<const bf16*>(reinterpret_cast<const char*>(a.inbox) + sizeof(InboxHdr)) + (size_t)slot * a.D;
    }
    const bool head_on = op_enabled(4 * a.nl, s);
    const int nops = 4 * a.nl + (head_on ? 1 : 0);
    float best_v = -INFINITY;
    int best_i = 0x7fffffff;

    for (int op = 0; op < nops && ok; ++op) {
      const OpInfo o = op_info(a, op);
      const LayerDev* lw = o.kind == OP_HEAD ? nullptr : &a.layers[o.li];
      const int K = o.K, KC = K / TK;
      // ---------------- prologue: activation vector -> shared memory (bf16) ----------------
      {
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
          for (int c = threadIdx.x; c < K / 8; c += GEMV_THREADS) {
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
          __syncthreads();
          float tot = 0.f;
#pragma unroll
          for (int w = 0; w < NW; ++w) tot += red[w];
          const float rs = 1.0f / sqrtf(tot / (float)K + a.eps);
          for (int c = threadIdx.x; c < K / 8; c += GEMV_THREADS) {
            const uint4 v = reinterpret_cast<uint4*>(x_s)[c];
            const uint4 g = __ldg(reinterpret_cast<const uint4*>(nw) + c);
            float f[8], sc[8];
            unpack8(v, f);
            unpack8(g, sc);
            uint32_t ov[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float y0 = round_bf(f[2 * j] * rs) * sc[2 * j];
              const float y1 = round_bf(f[2 * j + 1] * rs) * sc[2 * j + 1];
              __nv_bfloat162 p = __floats2bfloat162_rn(y0, y1);
              ov[j] = *reinterpret_cast<uint32_t*>(&p);
            }
            reinterpret_cast<uint4*>(x_s)[c] = make_uint4(ov[0], ov[1], ov[2], ov[3]);
          }
        } else {
          for (int c = threadIdx.x; c < K / 8; c += GEMV_THREADS) reinterpret_cast<uint4*>(x_s)[c] = __ldcg(xin4 + c);
        }
        __syncthreads();
      }
      // ---------------- stream this warp's tiles of the op ----------------
      int g0, gn;
      cta_groups(o.N, g0, gn);
      for (int grp = warp; grp < gn; grp += NW) {
        float acc[MK_TR];
#pragma unroll
        for (int r = 0; r < MK_TR; ++r) acc[r] = 0.f;
        for (int kc = 0; kc < KC; ++kc) {
          const int st = c_cnt % MK_S;
          mbar_wait(&bars[st], (c_cnt / MK_S) & 1u);
          const uint4* t4 = reinterpret_cast<const uint4*>(ring + (size_t)st * STAGE_BYTES);
          const uint4* x4 = reinterpret_cast<const uint4*>(x_s + (size_t)kc * TK);
#pragma unroll
          for (int c = 0; c < TK / 256; ++c) {
            float xf[8];
            unpack8(x4[c * 32 + lane], xf);
#pragma unroll
            for (int r = 0; r < MK_TR; ++r) acc[r] = dot8(t4[r * (TK / 8) + c * 32 + lane], xf, acc[r]);
          }
          ++c_cnt;
          __syncwarp();
          if (!p_done) p_issue();
        }
        const float v0 = warp_sum(acc[0]), v1 = warp_sum(acc[1]);
        const int n = (g0 + grp) * MK_TR;
        // ---------------- epilogues ----------------
        if (o.kind == OP_QKV) {
          // Linear out -> bf16; RoPE in fp32 with separate mul/mul/sub like eager torch -> bf16 (App. A.2, A.7)
          if (lane == 0) {
            const float x0 = round_bf(v0), x1 = round_bf(v1);
            const int q_rows = a.Hq * a.hd, kv_rows = a.Hkv * a.hd;
            if (n < q_rows + kv_rows) {
              const int ih = n % a.hd;
              const float2 cs = a.rope[(size_t)pos * (a.hd / 2) + (ih >> 1)];
              const float o0 = __fsub_rn(__fmul_rn(x0, cs.x), __fmul_rn(x1, cs.y));
              const float o1 = __fadd_rn(__fmul_rn(x1, cs.x), __fmul_rn(x0, cs.y));
              const __nv_bfloat162 p = __floats2bfloat162_rn(o0, o1);
              if (n < q_rows) {
                *reinterpret_cast<__nv_bfloat162*>(a.q + n) = p;
              } else {
                const int nk = n - q_rows;
                *reinterpret_cast<__nv_bfloat162*>(lw->k_cache + ((size_t)(nk / a.hd) * a.L + pos) * a.hd + ih) = p;
              }
            } else {
              const int nv = n - q_rows - kv_rows;
              *reinterpret_cast<__nv_bfloat162*>(lw->v_cache + ((size_t)(nv / a.hd) * a.L + pos) * a.hd + (nv % a.hd)) =
                  __floats2bfloat162_rn(x0, x1);
            }
          }
        } else if (o.kind == OP_O || o.kind == OP_DOWN) {
          // h = x + Linear(..): Linear out -> bf16, add in fp32 -> bf16 (App. A.5, A.7)
          if (lane == 0) {
            const __nv_bfloat162 rv = *reinterpret_cast<const __nv_bfloat162*>(a.x + n);
            const __nv_bfloat162 p =
                __floats2bfloat162_rn(__low2float(rv) + round_bf(v0), __high2float(rv) + round_bf(v1));
            *reinterpret_cast<__nv_bfloat162*>(a.x + n) = p;
            if (o.kind == OP_DOWN && o.li == a.nl - 1 && !a.last) {
              if (a.mode == 0) {   // fused hop: peer store into the next shard's inbox slot
                bf16* dst = reinterpret_cast<bf16*>(reinterpret_cast<char*>(a.next_inbox) + sizeof(InboxHdr)) + (size_t)slot * a.D;
                *reinterpret_cast<__nv_bfloat162*>(dst + n) = p;
              } else {
                *reinterpret_cast<__nv_bfloat162*>(a.out_hidden + (size_t)s * a.D + n) = p;
              }
            }
          }
        } else if (o.kind == OP_GU) {
          // rows interleaved (gate_m, up_m): act[m] = bf16( bf16(silu(bf16 g)) * bf16 u )  (llm_utils.py:491-500)
          if (lane == 0) {
            const float g = round_bf(v0), u = round_bf(v1);
            const float sg = round_bf(g / (1.0f + expf(-g)));
            a.act[n >> 1] = f2bf(sg * u);
          }
        } else {
          // logits (bf16-valued fp32) + running argmax, lowest index wins ties (rows ascend per warp)
          const float l0 = round_bf(v0), l1 = round_bf(v1);
          if (lane == 0) *reinterpret_cast<float2*>(a.logits + n) = make_float2(l0, l1);
          if (l0 > best_v) { best_v = l0; best_i = n; }
          if (l1 > best_v) { best_v = l1; best_i = n + 1; }
        }
      }
      // ---------------- op boundary ----------------
      if (o.kind == OP_HEAD) {
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
      const bool hop_out = (o.kind == OP_DOWN && o.li == a.nl - 1 && !a.last && a.mode == 0);
      if (hop_out) __threadfence_system();   // every writer orders its own peer stores before the barrier
      ok = grid_barrier(a, bar_target, &s_abort);
      if (!ok) break;
      if (o.kind == OP_QKV) {
        attn_phase<HD, R>(a, *lw, pos + 1, reinterpret_cast<float*>(xbuf));
        ok = grid_barrier(a, bar_target, &s_abort);
        if (!ok) break;
      }
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
  // ---- drain: never exit with bulk copies in flight ----
  while (c_cnt < p_cnt) {
    mbar_wait(&bars[c_cnt % MK_S], (c_cnt / MK_S) & 1u);
    ++c_cnt;
  }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    a.ctl->pos = a.pos0 + a.n_steps;
    a.ctl->step = a.step0 + a.n_steps;
  }
}

}  // namespace xb
