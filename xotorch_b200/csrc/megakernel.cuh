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
//   * grid = one CTA per SM, 10 warps: a TMA producer warp, 8 K-split consumer warps, an epilogue warp.  The producer
//     walks ONE tile list that spans all ops, layers and steps and feeds a CTA-level ring of 12 x 16 KiB shared-memory
//     stages with one bulk copy per stage (cp.async.bulk + mbarrier complete_tx; K-panel weight layout).  A stage is
//     refilled as soon as it is consumed, so while a CTA waits at a grid barrier or does attention its next weights
//     are already in flight (dense weights never depend on activations; sparse-MoE expert weights wait for the router).
//   * every consumer warp takes its 256 columns of every stage (packed FFMA2), partial sums go through a 4-slot
//     mbarrier mailbox to the epilogue warp: fixed-order reduction + fused RoPE / KV append / residual / SwiGLU / arg-max.
//   * ops are separated by grid barriers (release/acquire on a monotonic counter); activations (<= 56 KiB) travel
//     through L2; attention runs between the QKV and O ops as (kv head x context split) items with a ticket merge.
//   * the inter-shard hop is fused: the down-proj epilogue of a shard's last layer stores the 16-bit hidden state
//     straight into the NEXT shard's inbox over NVLink (peer stores) and CTA 0 releases a flag; the next shard's
//     CTAs acquire that flag before their first RMSNorm.  The last shard's lm_head epilogue produces the argmax
//     token and stores it into shard 0's inbox the same way (ring closure), so no host code runs per token.
//   * several requests can be in flight: ring step s serves request s % nreq (own KV cache, inbox slot s % (nreq+1)).
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
  const bf16 *attn_norm, *wqkv, *wo, *mlp_norm, *wgu, *wdown;   // MoE: wgu / wdown = expert 0, expert e at + e * stride
  const bf16* router;        // MoE: [E, D] row-major (NULL for a dense MLP)
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
  int nsplit, nsplit_min;   // context splits per kv head: upper bound (CTAs / kv heads) and the count used for short contexts
  StepCtl* ctl;
  InboxHdr *inbox, *next_inbox;
  int* tokens_log;
  unsigned long long *t_in, *t_out;
  unsigned int* bar_counter; // grid barrier counter (host zeroes it before every launch)
  int mode;                  // 0 ring, 1 direct
  // direct mode, one step (the API hop, xotb200_forward_hop): the hidden state arrives in / leaves through the inbox's API slots
  // — the same fused peer store + release/acquire flag as ring mode, but one engine call per shard per token
  int hop_in, hop_out;       // 1: input from this engine's inbox API slot / output into the next engine's
  unsigned int hop_seq_in, hop_seq_out;   // hop sequence numbers (slot = seq % API_SLOTS, flag value = seq + 1)
  int n_steps, step0, pos0;
  int nreq;                  // requests in flight (ring mode): ring step s serves request s % nreq at position pos0 + s / nreq
  int nslots;                // inbox slots in use = nreq + 1 (a shard is never more than nreq-1 steps ahead of the next one)
  int req0;                  // direct mode: the request whose KV cache this call works on
  const int* req_map;        // ring mode, optional [nreq]: KV cache of the i-th active request (NULL: i)
  const int* pos_tab;        // ring mode, optional [nreq]: position of the i-th active request's first step of this launch (NULL: pos0)
  int* tok_host;             // last shard, optional: host-mapped token log [MAX_RING_STEPS] + counter, readable while the kernel runs
  unsigned int* tok_count_host;
  size_t kv_req_stride;      // elements between the KV caches of consecutive requests
  const int* src_tokens;     // direct, first shard: [n_steps]
  const bf16* src_hidden;    // direct, other shards: [n_steps, D]
  bf16* out_hidden;          // direct, non-last shard: [n_steps, D]
  // Mixtral-style sparse MoE MLP (E == 0: dense).  Every CTA evaluates the router redundantly from the normalised x it
  // already holds in shared memory (E rows of D: no extra grid barrier), the gate/up and down ops then run as TWO
  // segments (the selected experts in ascending index) over the same row partition.
  int E, topk;
  size_t gu_stride, dn_stride;   // elements between consecutive experts' gate/up and down matrices
  int layer_tab_bytes;           // size of the shared-memory layer table (the MoE stash follows it)
  int* route_log;                // [L][nl] experts chosen for request 0 at (position, local layer): lo | hi << 8 (parity tests)
  int xbuf_bytes;            // shared-memory activation buffer size
  int l2_ahead;              // producer look-ahead into L2 while the ring is full (stages, 0 = off), see the producer
  int max_inflight;          // producer window: bulk copies issued and not yet landed (0 = the whole ring), see the producer
  unsigned int pace_q;       // producer pacing: 2^-16 ns per byte of this CTA's weight stream (0 = unpaced), see the producer
  unsigned long long* dbg;   // optional timeline [n_ops+1][4] of CTA dbg_cta for step dbg_step (NULL = off)
  int dbg_step, dbg_cta;
  unsigned long long* dbg_all;   // optional [n_ops][256][3]: (consume start, tiles done, landed stages) of EVERY CTA for step dbg_step
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

}  // namespace xb
