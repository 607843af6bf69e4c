// engine.cu — host side of libxotb200.so: engine object, weight layout + strict loader, request state (one KV cache per
// request in flight), the launchers (persistent decode kernel in direct and ring mode, tensor-core prefill incl. the
// sparse-MoE plan) and the C ABI declared in include/xotb200.h.
//
// HBM layout of a shard (16-bit weights: bf16 or IEEE half, 256-byte aligned, execution order so one decode step walks HBM linearly).
// Every streamed matrix [N, K] is stored as K-panels [K/TK][N][TK]: any 8 consecutive rows of one K chunk — one ring
// stage of the persistent kernel — are contiguous, so a stage is ONE TMA bulk copy of up to 32 KiB (the single producer
// thread can issue only ~14 M bulk copies/s; with 4 KiB row-wise copies it could not outrun HBM).
//   per layer: attn_norm[D] | Wqkv[(Hq+2Hkv)*hd, D] | Wo[D, Hq*hd] | mlp_norm[D] | Wgu[2F, D] | Wdown[D, F]
//     Wqkv rows: q heads then k heads in interleaved-pair order (the reference's `_permute`,
//       /root/reference/xotorch/inference/llm_utils.py:175-183) so a RoPE pair is two ADJACENT rows and the rotation
//       fuses into the GEMV epilogue; v rows unchanged.
//     Wgu rows: gate_m, up_m interleaved so SwiGLU fuses into the epilogue.
//     sparse MoE (Mixtral): Wgu / Wdown hold the E experts back to back (each in the same K-panel layout) + router[E, D].
//   first shard: embed[V, D];  last shard: final_norm[D] | lm_head[V, D] (a second, K-panelled copy of embed when tied).
//   KV cache: [request][local_layer][k|v][Hkv][L][hd] (same [Hkv, L, hd] order as torchtune KVCache).
#include <cuda_runtime.h>
#include <dirent.h>
#include <fcntl.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <thread>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/xotb200.h"
#include "kernels.cuh"
#include "megakernel.cuh"
#include "prefill_tc.cuh"

// Decode GEMV inner loop: 1 = mma.sync tensor-core path (default), 0 = the round-1 packed-FFMA2 loop (kept for A/B builds:
// make EXTRA=-DXB_GEMV_MMA=0).  XB_TILE_STAMPS=1 adds per-tile timestamps to the debug timeline (tools/quick_bench.py).
#ifndef XB_GEMV_MMA
#define XB_GEMV_MMA 1
#endif
#ifndef XOTB200_DEFAULT_PACE_GBS
#define XOTB200_DEFAULT_PACE_GBS 7300.0
#endif
#ifndef XB_GEMV_PAIRS
#define XB_GEMV_PAIRS 0
#endif
#ifndef XB_TILE_STAMPS
#define XB_TILE_STAMPS 0
#endif

// The dtype-dependent helpers and kernels are compiled twice: xb::bf (bfloat16 weights) and xb::fp (IEEE half weights).
namespace xb {
namespace bf {
#define XB_DT 0
#include "dtype_impl.inc"
#include "megakernel_impl.inc"
#include "prefill_tc_impl.inc"
#undef XB_DT
}  // namespace bf
namespace fp {
#define XB_DT 1
#include "dtype_impl.inc"
#include "megakernel_impl.inc"
#include "prefill_tc_impl.inc"
#undef XB_DT
}  // namespace fp
}  // namespace xb
// launch the bf16 or the fp16 instantiation of a kernel expression
#define DT_LAUNCH(e, call)                 \
  do {                                     \
    if ((e)->cfg.dtype == XOTB200_DTYPE_F16) { using namespace xb::fp; call; } \
    else { using namespace xb::bf; call; } \
  } while (0)

using namespace xb;

// --------------------------------------------------------------------------------------------------------------
static thread_local char g_err[1024] = "";
static void set_err(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
#define CK(call)                                                                                            \
  do {                                                                                                      \
    cudaError_t _e = (call);                                                                                \
    if (_e != cudaSuccess) {                                                                                \
      set_err("%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__);                  \
      return XOTB200_ECUDA;                                                                                 \
    }                                                                                                       \
  } while (0)
#define REQUIRE(cond, ...)                \
  do {                                    \
    if (!(cond)) {                        \
      set_err(__VA_ARGS__);               \
      return XOTB200_EINVAL;              \
    }                                     \
  } while (0)

struct TensorSlot {
  bf16* dst;        // base of the destination matrix
  int rows, K;      // HF shape [rows, K] (1-D tensors: rows=1)
  int mode;         // 0 identity, 1 q/k interleave
  int stride, off;  // destination row = off + r*stride
  int tensor_id;    // synth.py tensor id
  float std, mean;
  bool loaded;
  int n_total;      // rows of the destination matrix
  int panel_tk;     // 0: plain row-major; else K-panel layout [K/tk][n_total][tk] (streamed GEMV weights)
};

struct LayerW {
  bf16 *attn_norm, *wqkv, *wo, *mlp_norm, *wgu, *wdown;   // MoE: wgu / wdown hold the E experts back to back
  bf16* router;                                           // MoE: [E, D] row-major
};

struct xotb200_engine {
  xotb200_model_cfg cfg;
  int device = 0, sm_count = 148;
  int start = -1, end = -1;
  bool first = false, last = false;
  bf16* blob = nullptr;
  size_t blob_bytes = 0;
  std::vector<LayerW> layers;
  bf16 *embed = nullptr, *final_norm = nullptr, *lm_head = nullptr;
  std::multimap<std::string, TensorSlot> slots;   // HF name -> destination(s) (a tied embedding feeds two)
  int tk = 0;                                     // columns per ring stage, fixed per engine
  // request state
  bf16* kv = nullptr;              // [nreq][local layer][k|v][Hkv][L][hd]
  int L = 0;
  int nreq = 1, cur_req = 0;       // requests in flight through the ring / request selected for xotb200_forward
  size_t kv_req_stride = 0;
  float2* rope = nullptr;
  int rope_len = 0;
  // activations / scratch
  bf16 *x = nullptr, *q = nullptr, *attn_out = nullptr, *act = nullptr;
  float* logits = nullptr;
  float *part_ml = nullptr, *part_acc = nullptr;
  unsigned int* tickets = nullptr;  // [Hkv] attention + [1] step end
  float* cta_best_val = nullptr;
  int* cta_best_idx = nullptr;
  int nsplit = 1, nsplit_min = 1;
  StepCtl* ctl = nullptr;
  InboxHdr* inbox = nullptr;
  InboxHdr* next_inbox = nullptr;
  bool next_ipc = false;
  // API hop (xotb200_forward_hop): one engine call per shard per token with the hidden state moved by the producing kernel
  xotb200_engine* next_engine = nullptr;   // same-process link: needed for multi-token (prompt) hops, which are peer copies
  unsigned int hop_seq_next = 0;           // sender side: next hop number of this link (reset when the link is made)
  std::mutex hop_mu;                       // receiver side: guards the growth of hop_buf
  bf16* hop_buf[API_SLOTS] = {};           // receiver side: [n_tokens, D] prompt hidden states pushed by the previous engine
  int hop_cap[API_SLOTS] = {};
  int hop_in = 0, hop_out = 0;             // per-launch arguments of the API hop (see MegaArgs)
  // continuous batching: the active requests of the next ring launch (KV slot + position each); NULL/0 = all nreq from pos0
  int *d_req_map = nullptr, *d_pos_tab = nullptr;
  int n_active = 0;
  // generated tokens streamed to the host while the persistent kernel runs (pinned, mapped)
  int* tok_host = nullptr;
  int* tok_host_dev = nullptr;
  volatile unsigned int* tok_count_host = nullptr;
  unsigned int* tok_count_dev = nullptr;
  unsigned int hop_seq_in = 0, hop_seq_out = 0;
  int* tokens_log = nullptr;
  int* route_log = nullptr;         // sparse MoE: experts chosen per (position, local layer) of request 0
  unsigned long long *t_in = nullptr, *t_out = nullptr;
  // staging for the host-buffer API
  int stage_cap = 0;
  int* stage_tokens = nullptr;
  float* stage_f32 = nullptr;
  bf16 *stage_in = nullptr, *stage_out = nullptr;
  bool have_result = false;
  cudaStream_t stream = nullptr;
  LayerDev* d_layers = nullptr;
  unsigned int* bar_counter = nullptr;
  void* mega = nullptr;
  size_t mega_smem = 0;
  int mega_xbuf = 0, mega_grid = 0, mega_nstage = 0;
  int cta_limit = 0;
  // tensor-core prefill (prefill_tc.cuh)
  int l2_ahead = 0;                // decode producer look-ahead into L2 (stages per CTA)
  int max_inflight = 0;            // decode producer window (bulk copies in flight per CTA, 0 = ring depth)
  unsigned int pace_q = 0;         // decode producer pacing, 2^-16 ns per byte per CTA (0 = off)
  float* tc_ws = nullptr;          // split-K partial tiles of the one-tile prefill GEMMs, and their tickets
  size_t tc_ws_bytes = 0;
  unsigned int* tc_ws_ticket = nullptr;
  bool tc_fuse = true;             // RoPE/KV-append and SwiGLU fused into GEMM epilogues (XOTB200_TC_FUSE=0: separate kernels, A/B)
  int prefill_mode = 0;            // 0 auto (TC when n_tokens >= 8), 1 always the decode kernel, 2 always TC
  bf16 *tc_x = nullptr, *tc_xn = nullptr, *tc_qkv = nullptr, *tc_q = nullptr, *tc_ao = nullptr, *tc_gu = nullptr, *tc_act = nullptr;
  CUtensorMap tm_xn, tm_ao, tm_act;
  std::map<int, std::vector<CUtensorMap>> tm_w;   // per N-tile width: 4 per local layer (qkv, o, gu, down) + head
  int tc_cap = 0;                  // token rows the tensor-core scratch buffers currently hold
  // sparse MoE prefill: gathered per-expert row blocks (rows_max = 2 * tc_cap + E * 128) and the device-side plan
  bf16 *moe_xg = nullptr, *moe_gu = nullptr, *moe_act = nullptr, *moe_y = nullptr;
  int *moe_route_e = nullptr, *moe_row_of = nullptr, *moe_row_token = nullptr, *moe_exp_info = nullptr;
  float* moe_route_w = nullptr;
  int moe_rows = 0;
  CUtensorMap tm_moe_xg, tm_moe_act;
  unsigned long long* dbg = nullptr;
  int dbg_step = -1, dbg_cta = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int64_t launches = 0;
  int64_t load_bytes = 0;          // xotb200_load_shard: bytes moved host -> HBM, wall seconds, seconds spent copying out of the mapping
  double load_seconds = 0.0, load_memcpy_seconds = 0.0;
  int last_ring_steps = 0;
  int ring_next_step = 0;          // first ring step number no run on this engine has used (inbox flags are never reused)
};

static size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Every copy / memset is ordered on the ENGINE stream (created cudaStreamNonBlocking: the legacy default stream does
// not synchronise with it, so plain cudaMemcpy/cudaMemset would race with the engine's kernels).
#define MEMSET_SYNC(e, ptr, val, bytes)                         \
  do {                                                          \
    CK(cudaMemsetAsync((ptr), (val), (bytes), (e)->stream));    \
    CK(cudaStreamSynchronize((e)->stream));                     \
  } while (0)
#define MEMCPY_SYNC(e, dst, src, bytes, kind)                        \
  do {                                                               \
    CK(cudaMemcpyAsync((dst), (src), (bytes), (kind), (e)->stream)); \
    CK(cudaStreamSynchronize((e)->stream));                          \
  } while (0)

// --------------------------------------------------------------------------------------------------------------
// megakernel dispatch
// --------------------------------------------------------------------------------------------------------------
typedef void (*mega_fn)(const MegaArgs);

#define PICK_MEGA_NS(NSP)                                                                                   \
  template <int TK, int HD, int NS, bool MOE>                                                                \
  static mega_fn pick_mega_r_##NSP(int R) {                                                                  \
    switch (R) {                                                                                             \
      case 2: return xb::NSP::shard_step_kernel<TK, HD, 2, NS, MOE>;                                         \
      case 4: return xb::NSP::shard_step_kernel<TK, HD, 4, NS, MOE>;                                         \
      case 8: return xb::NSP::shard_step_kernel<TK, HD, 8, NS, MOE>;                                         \
      default: return nullptr;                                                                               \
    }                                                                                                        \
  }                                                                                                          \
  static mega_fn pick_mega_##NSP(int tk, int hd, int R, int ns, bool moe) {                                  \
    if (moe) { /* sparse MoE: 2F activations -> 8-stage ring at Mixtral sizes, 12 at the tiny parity shapes */ \
      if (tk == 2048 && hd == 128 && ns == 8) return pick_mega_r_##NSP<2048, 128, 8, true>(R);              \
      if (tk == 2048 && hd == 128 && ns == 12) return pick_mega_r_##NSP<2048, 128, 12, true>(R);            \
      if (tk == 512 && hd == 128 && ns == 12) return pick_mega_r_##NSP<512, 128, 12, true>(R);              \
      if (tk == 512 && hd == 64 && ns == 12) return pick_mega_r_##NSP<512, 64, 12, true>(R);                \
      return nullptr;                                                                                        \
    }                                                                                                        \
    if (tk == 2048 && hd == 128 && ns == 12) return pick_mega_r_##NSP<2048, 128, 12, false>(R);             \
    if (tk == 2048 && hd == 64 && ns == 12) return pick_mega_r_##NSP<2048, 64, 12, false>(R);               \
    if (tk == 2048 && hd == 128 && ns == 8) return pick_mega_r_##NSP<2048, 128, 8, false>(R); /* very wide (70B) */ \
    if (tk == 512 && hd == 128 && ns == 12) return pick_mega_r_##NSP<512, 128, 12, false>(R);               \
    if (tk == 512 && hd == 64 && ns == 12) return pick_mega_r_##NSP<512, 64, 12, false>(R);                 \
    return nullptr;                                                                                          \
  }
PICK_MEGA_NS(bf)
PICK_MEGA_NS(fp)
static mega_fn pick_mega(int tk, int hd, int R, int ns, int dtype, bool moe) {
  return dtype == XOTB200_DTYPE_F16 ? pick_mega_fp(tk, hd, R, ns, moe) : pick_mega_bf(tk, hd, R, ns, moe);
}

static const size_t SMEM_BUDGET = 232448;  // 227 KiB opt-in maximum per CTA on sm_100

// Resolve the kernel instantiation, shared-memory size and grid for this engine (once per engine).
static int prepare_mega(xotb200_engine* e) {
  if (e->mega) return 0;
  const xotb200_model_cfg& c = e->cfg;
  const int qdim = c.n_heads * c.head_dim;
  const int tk = e->tk;
  REQUIRE(tk == 2048 || tk == 512, "no shard set");
  const int R = c.n_heads / c.n_kv_heads;
  const int grid0 = e->cta_limit > 0 ? std::min(e->sm_count, e->cta_limit) : e->sm_count;
  const size_t moe_stash = c.n_experts > 0 ? align_up(((size_t)c.dim / 2 + grid0 - 1) / grid0 * 8, 16) : 0;   // float2 per row pair of a CTA
  const size_t kmax = (size_t)std::max(std::max(c.dim, c.ffn * (c.n_experts > 0 ? 2 : 1)), qdim);
  const size_t attn_scratch = ((size_t)NCW * R * 2 + (size_t)NCW * R * c.head_dim) * 4;
  const size_t xbuf = align_up(std::max(kmax * 2, attn_scratch), 128);
  const size_t fixed = xbuf + (size_t)(2 * mk_nstage(tk) + 2 * MK_NSLOT) * 8 + (size_t)MK_NSLOT * NCW * 8 * 4 + 128 + 128 +
                       (size_t)c.n_layers * sizeof(LayerDev) + moe_stash + 2048 /* static __shared__ */;
  REQUIRE(fixed + 4 * (size_t)mk_stage_bytes(tk) <= SMEM_BUDGET, "activation width %zu too large for a 4-stage ring", kmax);
  const int nstage = (SMEM_BUDGET - fixed) / mk_stage_bytes(tk) >= 12 ? 12 : 8;
  REQUIRE(fixed + (size_t)nstage * mk_stage_bytes(tk) <= SMEM_BUDGET, "activation width %zu too large for an 8-stage ring", kmax);
  mega_fn fn = pick_mega(tk, c.head_dim, R, nstage, c.dtype, c.n_experts > 0);
  REQUIRE(fn != nullptr, "no megakernel instantiation for tk=%d head_dim=%d gqa=%d stages=%d", tk, c.head_dim, R, nstage);
  const size_t smem = (size_t)nstage * mk_stage_bytes(tk) + fixed - 2048;
  e->mega_nstage = nstage;
  CK(cudaFuncSetAttribute((const void*)fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, (const void*)fn, MK_THREADS, smem));
  REQUIRE(per_sm >= 1, "megakernel does not fit on an SM");
  e->mega = (void*)fn;
  e->mega_smem = smem;
  e->mega_xbuf = (int)xbuf;
  e->mega_grid = e->sm_count;   // one CTA per SM, all co-resident (cooperative launch checks it)
  if (e->cta_limit > 0) e->mega_grid = std::min(e->mega_grid, e->cta_limit);
  e->nsplit = std::max(1, std::min(32, e->mega_grid / c.n_kv_heads));
  e->nsplit_min = std::min(e->nsplit, 12);
  if (const char* v = getenv("XOTB200_NSPLIT")) e->nsplit = std::max(1, std::min(e->nsplit, atoi(v)));   // tuning hooks
  if (const char* v = getenv("XOTB200_NSPLIT_MIN")) e->nsplit_min = std::max(1, atoi(v));
  e->nsplit_min = std::min(e->nsplit_min, e->nsplit);
  // producer pacing (see shard_step_kernel): target AGGREGATE weight-stream rate of the grid in GB/s, shared evenly by the CTAs
  double pace_gbs = XOTB200_DEFAULT_PACE_GBS;
  if (const char* v = getenv("XOTB200_PACE_GBS")) pace_gbs = atof(v);
  if (const char* v = getenv("XOTB200_L2_AHEAD")) e->l2_ahead = std::max(0, atoi(v));
  if (const char* v = getenv("XOTB200_INFLIGHT")) e->max_inflight = std::max(0, atoi(v));
  e->pace_q = pace_gbs > 0 ? (unsigned int)(65536.0 * e->mega_grid / pace_gbs + 0.5) : 0u;
  return 0;
}

static size_t dbg_cap(const xotb200_engine* e) { return (size_t)(4 * e->cfg.n_layers + 2) * 4 * 2 + 4 * 64 + 8; }

// Launch the persistent shard-step kernel for n_steps steps.  mode 0 = ring (inbox in, next inbox out),
// mode 1 = direct (staging buffers).
static int launch_mega(xotb200_engine* e, int mode, int n_steps, int step0, int pos0, cudaStream_t st) {
  if (int rc = prepare_mega(e)) return rc;
  const xotb200_model_cfg& c = e->cfg;
  MegaArgs a{};
  a.layers = e->d_layers; a.nl = e->end - e->start + 1;
  a.D = c.dim; a.F = c.ffn; a.Hq = c.n_heads; a.Hkv = c.n_kv_heads; a.hd = c.head_dim; a.V = c.vocab; a.L = e->L;
  a.eps = c.norm_eps; a.attn_scale = 1.0f / sqrtf((float)c.head_dim); a.rope = e->rope;
  a.embed = e->embed; a.final_norm = e->final_norm; a.lm_head = e->lm_head;
  a.first = e->first ? 1 : 0; a.last = e->last ? 1 : 0;
  a.x = e->x; a.q = e->q; a.attn_out = e->attn_out; a.act = e->act; a.logits = e->logits;
  a.part_ml = e->part_ml; a.part_acc = e->part_acc; a.tickets = e->tickets;
  a.cta_best_val = e->cta_best_val; a.cta_best_idx = e->cta_best_idx; a.nsplit = e->nsplit; a.nsplit_min = e->nsplit_min;
  a.ctl = e->ctl; a.inbox = e->inbox; a.next_inbox = e->next_inbox; a.tokens_log = e->tokens_log;
  a.t_in = e->t_in; a.t_out = e->t_out; a.bar_counter = e->bar_counter;
  a.mode = mode; a.n_steps = n_steps; a.step0 = step0; a.pos0 = pos0;
  const int nact = (mode == 0 && e->n_active > 0) ? e->n_active : e->nreq;
  a.nreq = nact; a.nslots = nact + 1; a.req0 = e->cur_req; a.kv_req_stride = e->kv_req_stride;
  a.req_map = (mode == 0 && e->n_active > 0) ? e->d_req_map : nullptr;
  a.pos_tab = (mode == 0 && e->n_active > 0) ? e->d_pos_tab : nullptr;
  a.tok_host = (mode == 0 && e->last) ? e->tok_host_dev : nullptr;
  a.tok_count_host = e->tok_count_dev;
  a.src_tokens = e->stage_tokens; a.src_hidden = e->stage_in; a.out_hidden = e->stage_out;
  a.hop_in = mode == 1 ? e->hop_in : 0; a.hop_out = mode == 1 ? e->hop_out : 0;
  a.hop_seq_in = e->hop_seq_in; a.hop_seq_out = e->hop_seq_out;
  a.xbuf_bytes = e->mega_xbuf;
  a.pace_q = e->pace_q;
  a.max_inflight = e->max_inflight;
  a.l2_ahead = e->l2_ahead;
  a.E = c.n_experts; a.topk = c.experts_per_tok;
  a.gu_stride = (size_t)2 * c.ffn * c.dim; a.dn_stride = (size_t)c.dim * c.ffn;
  a.layer_tab_bytes = c.n_layers * (int)sizeof(LayerDev);
  a.route_log = e->route_log;
  a.dbg = (mode == 0 && e->dbg_step >= 0) ? e->dbg : nullptr; a.dbg_step = e->dbg_step; a.dbg_cta = e->dbg_cta;
  a.dbg_all = a.dbg ? e->dbg + dbg_cap(e) : nullptr;
  CK(cudaMemsetAsync(e->bar_counter, 0, 4, st));
  void* params[] = {&a};
  CK(cudaLaunchCooperativeKernel((const void*)e->mega, dim3(e->mega_grid), dim3(MK_THREADS), params, e->mega_smem, st));
  e->launches++;
  return 0;
}

// --------------------------------------------------------------------------------------------------------------
// tensor-core prefill: tensor maps, scratch, driver
// --------------------------------------------------------------------------------------------------------------
typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static encode_tiled_fn get_encode_tiled() {
  static encode_tiled_fn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  }
  return fn;
}
// activations [rows, K] row-major bf16 (rows a multiple of 128) -> box {64 k, 128 rows}, SWIZZLE_128B
static int make_tmap_act(CUtensorMap* m, bf16* ptr, int K, int rows = TC_M) {
  encode_tiled_fn enc = get_encode_tiled();
  REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)K * 2};
  const cuuint32_t box[2] = {(cuuint32_t)TC_BK, (cuuint32_t)TC_M};
  const cuuint32_t es[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, ptr, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(activation K=%d) failed: %d", K, (int)r);
  return 0;
}
// weights in K-panel layout [K/tk][N][tk] -> 3-D map {tk, N, K/tk}, box {64 k, nt rows, 1}, SWIZZLE_128B
static int make_tmap_w(CUtensorMap* m, const bf16* ptr, int N, int K, int tk, int nt) {
  encode_tiled_fn enc = get_encode_tiled();
  REQUIRE(enc != nullptr, "cuTensorMapEncodeTiled not available from the driver");
  const cuuint64_t dims[3] = {(cuuint64_t)tk, (cuuint64_t)N, (cuuint64_t)(K / tk)};
  const cuuint64_t strides[2] = {(cuuint64_t)tk * 2, (cuuint64_t)N * tk * 2};
  const cuuint32_t box[3] = {(cuuint32_t)TC_BK, (cuuint32_t)nt, 1};
  const cuuint32_t es[3] = {1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<bf16*>(ptr), dims, strides, box, es,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  REQUIRE(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled(weights N=%d K=%d tk=%d) failed: %d", N, K, tk, (int)r);
  return 0;
}
// GEMM tile configuration {N tile, ring stages}.  One 128-token tile (S <= 128) is HBM-bound: many CTAs, deep rings.
// Several token tiles are tensor-bound: measured on B200 (8B shape, 1024 / 4096 tokens, whole prefill): 256x4 16.2 / 80 ms,
// 128x3 (two CTAs per SM, epilogue of one under the MMAs of the other) 17.4 / 88, 128x6 21.1 / 99, 64x8 34.7 / 153.  XOTB200_TC_SMALL / XOTB200_TC_BIG = "nt,ns" override (tuning hook).
struct TcCfg { int nt, ns; };
static TcCfg tc_env(const char* name, TcCfg dflt) {
  const char* v = getenv(name);
  TcCfg c = dflt;
  if (v && sscanf(v, "%d,%d", &c.nt, &c.ns) != 2) c = dflt;
  return c;
}
static TcCfg tc_cfg(int N, int S) {
  static const TcCfg small = tc_env("XOTB200_TC_SMALL", {64, 8}), big = tc_env("XOTB200_TC_BIG", {256, 4});
  TcCfg c = S > TC_M ? big : (N <= 8192 ? small : TcCfg{128, 6});
  if (N % c.nt != 0 && S > TC_M && N % 128 == 0) c = {128, 3};
  if (N % c.nt != 0) c = {64, 8};
  return c;
}
static int tc_nt(int N) { return tc_cfg(N, 1).nt; }

// cudaFuncSetAttribute(max dynamic smem) once per (device, kernel).  The cache is process-wide and engines on different GPUs
// run on different host threads (ctypes drops the GIL), so it is guarded by a mutex.
static int ensure_smem_attr(int device, const void* fn, int bytes) {
  static std::mutex mu;
  static std::map<std::pair<int, const void*>, bool> done;
  std::lock_guard<std::mutex> lock(mu);
  auto key = std::make_pair(device, fn);
  if (!done.count(key)) {
    CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    done[key] = true;
  }
  return 0;
}

template <int NT, int NS>
static int launch_tc_gemm_nt(xotb200_engine* e, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcGemmArgs& a, int epi,
                             cudaStream_t st) {
  void (*fn)(CUtensorMap, CUtensorMap, TcGemmArgs) = nullptr;
  const bool f16 = e->cfg.dtype == XOTB200_DTYPE_F16;
  switch (epi) {
    case TCE_STORE: fn = f16 ? xb::fp::tc_gemm_kernel<NT, TCE_STORE, NS> : xb::bf::tc_gemm_kernel<NT, TCE_STORE, NS>; break;
    case TCE_RESID: fn = f16 ? xb::fp::tc_gemm_kernel<NT, TCE_RESID, NS> : xb::bf::tc_gemm_kernel<NT, TCE_RESID, NS>; break;
    case TCE_QKV: fn = f16 ? xb::fp::tc_gemm_kernel<NT, TCE_QKV, NS> : xb::bf::tc_gemm_kernel<NT, TCE_QKV, NS>; break;
    case TCE_SWIGLU: fn = f16 ? xb::fp::tc_gemm_kernel<NT, TCE_SWIGLU, NS> : xb::bf::tc_gemm_kernel<NT, TCE_SWIGLU, NS>; break;
    default: fn = f16 ? xb::fp::tc_gemm_kernel<NT, TCE_HEAD_LAST, NS> : xb::bf::tc_gemm_kernel<NT, TCE_HEAD_LAST, NS>; break;
  }
  if (int rc = ensure_smem_attr(e->device, (const void*)fn, tc_smem_bytes(NT, NS))) return rc;
  const int mtiles = a.grid_m > 0 ? a.grid_m : (epi == TCE_HEAD_LAST ? 1 : (a.S + TC_M - 1) / TC_M - a.mt0);
  fn<<<dim3(mtiles, a.N / NT, a.ksplit > 1 ? a.ksplit : 1), TC_THREADS, tc_smem_bytes(NT, NS), st>>>(tmA, tmB, a);
  CK(cudaGetLastError());
  e->launches++;
  return 0;
}
// persistent variant (one CTA per SM, two TMEM accumulators): used when a launch has more output tiles than SMs
template <int NT, int NS>
static int launch_tc_gemm_persistent(xotb200_engine* e, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcGemmArgs& a, int epi,
                                     int tiles, cudaStream_t st) {
  void (*fn)(CUtensorMap, CUtensorMap, TcGemmArgs) = nullptr;
  const bool f16 = e->cfg.dtype == XOTB200_DTYPE_F16;
  switch (epi) {
    case TCE_STORE: fn = f16 ? xb::fp::tc_gemm_persistent_kernel<NT, TCE_STORE, NS> : xb::bf::tc_gemm_persistent_kernel<NT, TCE_STORE, NS>; break;
    case TCE_RESID: fn = f16 ? xb::fp::tc_gemm_persistent_kernel<NT, TCE_RESID, NS> : xb::bf::tc_gemm_persistent_kernel<NT, TCE_RESID, NS>; break;
    case TCE_QKV: fn = f16 ? xb::fp::tc_gemm_persistent_kernel<NT, TCE_QKV, NS> : xb::bf::tc_gemm_persistent_kernel<NT, TCE_QKV, NS>; break;
    case TCE_SWIGLU: fn = f16 ? xb::fp::tc_gemm_persistent_kernel<NT, TCE_SWIGLU, NS> : xb::bf::tc_gemm_persistent_kernel<NT, TCE_SWIGLU, NS>; break;
    default: fn = f16 ? xb::fp::tc_gemm_persistent_kernel<NT, TCE_HEAD_LAST, NS> : xb::bf::tc_gemm_persistent_kernel<NT, TCE_HEAD_LAST, NS>; break;
  }
  if (int rc = ensure_smem_attr(e->device, (const void*)fn, tc_smem_bytes(NT, NS))) return rc;
  fn<<<std::min(tiles, e->sm_count), TC_THREADS, tc_smem_bytes(NT, NS), st>>>(tmA, tmB, a);
  CK(cudaGetLastError());
  e->launches++;
  return 0;
}

// CTA-pair variant (cta_group::2): clusters of two CTAs, one 256-token x NT tile per pair; tmB is the weight map with a box of NT/2 rows
template <int NT, int NS>
static int launch_tc_gemm_2sm(xotb200_engine* e, const CUtensorMap& tmA, const CUtensorMap& tmB_half, const TcGemmArgs& a_in, int epi, cudaStream_t st) {
  TcGemmArgs a = a_in;
  a.err = &e->ctl->err;
  void (*fn)(CUtensorMap, CUtensorMap, TcGemmArgs) = nullptr;
  const bool f16 = e->cfg.dtype == XOTB200_DTYPE_F16;
  switch (epi) {
    case TCE_STORE: fn = f16 ? xb::fp::tc_gemm_2sm_kernel<NT, TCE_STORE, NS> : xb::bf::tc_gemm_2sm_kernel<NT, TCE_STORE, NS>; break;
    case TCE_RESID: fn = f16 ? xb::fp::tc_gemm_2sm_kernel<NT, TCE_RESID, NS> : xb::bf::tc_gemm_2sm_kernel<NT, TCE_RESID, NS>; break;
    case TCE_QKV: fn = f16 ? xb::fp::tc_gemm_2sm_kernel<NT, TCE_QKV, NS> : xb::bf::tc_gemm_2sm_kernel<NT, TCE_QKV, NS>; break;
    case TCE_SWIGLU: fn = f16 ? xb::fp::tc_gemm_2sm_kernel<NT, TCE_SWIGLU, NS> : xb::bf::tc_gemm_2sm_kernel<NT, TCE_SWIGLU, NS>; break;
    default: set_err("no CTA-pair GEMM for epilogue %d", epi); return XOTB200_EINVAL;
  }
  if (int rc = ensure_smem_attr(e->device, (const void*)fn, tc2_smem_bytes(NT, NS))) return rc;
  const int total = ((a.S + 2 * TC_M - 1) / (2 * TC_M)) * (a.N / NT);
  const int pairs = std::max(1, std::min(total, e->sm_count / 2));
  fn<<<2 * pairs, TC_THREADS, tc2_smem_bytes(NT, NS), st>>>(tmA, tmB_half, a);
  CK(cudaGetLastError());
  e->launches++;
  return 0;
}

static int launch_tc_gemm(xotb200_engine* e, const CUtensorMap& tmA, const CUtensorMap& tmB, const TcGemmArgs& a_in, int epi, TcCfg c,
                          cudaStream_t st) {
  TcGemmArgs a = a_in;
  a.err = &e->ctl->err;
  REQUIRE(a.N % c.nt == 0 && a.K % TC_BK == 0, "tensor-core GEMM needs N %% %d == 0 and K %% 64 == 0 (N=%d K=%d)", c.nt, a.N, a.K);
  {
    static const bool persist = !(getenv("XOTB200_TC_PERSIST") && atoi(getenv("XOTB200_TC_PERSIST")) == 0);
    const int mt = a.grid_m > 0 ? a.grid_m : (epi == TCE_HEAD_LAST ? 1 : (a.S + TC_M - 1) / TC_M - a.mt0);
    const int tiles = mt * (a.N / c.nt);
    if (persist && tiles > e->sm_count) {
      if (c.nt == 256 && c.ns == 4) return launch_tc_gemm_persistent<256, 4>(e, tmA, tmB, a, epi, tiles, st);
      if (c.nt == 128 && (c.ns == 6 || c.ns == 3)) return launch_tc_gemm_persistent<128, 6>(e, tmA, tmB, a, epi, tiles, st);
    }
  }
#define TC_CASE(NT_, NS_) if (c.nt == NT_ && c.ns == NS_) return launch_tc_gemm_nt<NT_, NS_>(e, tmA, tmB, a, epi, st)
  TC_CASE(32, 8); TC_CASE(64, 8); TC_CASE(64, 4); TC_CASE(128, 6); TC_CASE(128, 3); TC_CASE(256, 4); TC_CASE(256, 2);
#undef TC_CASE
  set_err("no tensor-core GEMM instantiation for tile %d x %d stages", c.nt, c.ns);
  return XOTB200_EINVAL;
}

// weight tensor maps for N-tile width nt: 4 per local layer (qkv, o, gu, down) + head; built on first use of that width
static int ensure_tm_w(xotb200_engine* e, int nt) {
  if (e->tm_w.count(nt)) return 0;
  const xotb200_model_cfg& c = e->cfg;
  const int D = c.dim, F = c.ffn, qd = c.n_heads * c.head_dim, nqkv = (c.n_heads + 2 * c.n_kv_heads) * c.head_dim;
  const int nl = e->end - e->start + 1;
  std::vector<CUtensorMap> v((size_t)4 * nl + 1, CUtensorMap{});
  auto mk = [&](CUtensorMap* m, const bf16* w, int N, int K) { return N % nt ? 0 : make_tmap_w(m, w, N, K, e->tk, nt); };
  for (int li = 0; li < nl; ++li) {
    const LayerW& w = e->layers[li];
    if (int rc = mk(&v[4 * li + 0], w.wqkv, nqkv, D)) return rc;
    if (int rc = mk(&v[4 * li + 1], w.wo, D, qd)) return rc;
    if (int rc = mk(&v[4 * li + 2], w.wgu, 2 * F, D)) return rc;
    if (int rc = mk(&v[4 * li + 3], w.wdown, D, F)) return rc;
  }
  if (e->last)
    if (int rc = mk(&v[4 * nl], e->lm_head, c.vocab, D)) return rc;
  if (c.n_experts > 0) {   // sparse MoE: per-expert gate/up and down maps follow at 4*nl + 1 + (li*E + x)*2 + {0, 1}
    const int E = c.n_experts;
    v.resize((size_t)4 * nl + 1 + (size_t)2 * nl * E, CUtensorMap{});
    for (int li = 0; li < nl; ++li)
      for (int x = 0; x < E; ++x) {
        const LayerW& w = e->layers[li];
        if (int rc = mk(&v[4 * nl + 1 + (li * E + x) * 2 + 0], w.wgu + (size_t)x * 2 * F * D, 2 * F, D)) return rc;
        if (int rc = mk(&v[4 * nl + 1 + (li * E + x) * 2 + 1], w.wdown + (size_t)x * D * F, D, F)) return rc;
      }
  }
  e->tm_w[nt] = std::move(v);
  return 0;
}

static int prepare_tc(xotb200_engine* e, int n_tokens) {
  const xotb200_model_cfg& c = e->cfg;
  const size_t D = c.dim, F = c.ffn, qd = (size_t)c.n_heads * c.head_dim, nqkv = (size_t)(c.n_heads + 2 * c.n_kv_heads) * c.head_dim;
  const int need = (std::min(n_tokens, TC_CHUNK) + TC_M - 1) / TC_M * TC_M;
  if (need > e->tc_cap) {   // scratch for one chunk of `need` token rows (grown on demand, kept across requests)
    CK(cudaStreamSynchronize(e->stream));
    bf16** bufs[] = {&e->tc_x, &e->tc_xn, &e->tc_qkv, &e->tc_q, &e->tc_ao, &e->tc_gu, &e->tc_act};
    const size_t fdense = c.n_experts > 0 ? 8 : F;   // sparse MoE shards use the gathered buffers below instead of tc_gu / tc_act
    const size_t cols[] = {D, D, nqkv, qd, qd, 2 * fdense, fdense};
    for (int i = 0; i < 7; ++i) {
      if (*bufs[i]) CK(cudaFree(*bufs[i]));
      *bufs[i] = nullptr;
      CK(cudaMalloc(bufs[i], (size_t)need * cols[i] * 2));
      MEMSET_SYNC(e, *bufs[i], 0, (size_t)need * cols[i] * 2);
    }
    e->tc_cap = need;
    if (c.n_experts > 0) {
      const size_t rows = (size_t)2 * need + (size_t)c.n_experts * TC_M;
      void** mb[] = {(void**)&e->moe_xg, (void**)&e->moe_gu, (void**)&e->moe_act, (void**)&e->moe_y, (void**)&e->moe_route_e,
                     (void**)&e->moe_route_w, (void**)&e->moe_row_of, (void**)&e->moe_row_token, (void**)&e->moe_exp_info};
      const size_t mbytes[] = {rows * D * 2, rows * 2 * F * 2, rows * F * 2, rows * D * 2, (size_t)need * 2 * 4, (size_t)need * 2 * 4,
                               (size_t)need * 2 * 4, rows * 4, (size_t)(2 * c.n_experts + 1) * 4};
      for (int i = 0; i < 9; ++i) {
        if (*mb[i]) CK(cudaFree(*mb[i]));
        *mb[i] = nullptr;
        CK(cudaMalloc(mb[i], mbytes[i]));
        MEMSET_SYNC(e, *mb[i], 0, mbytes[i]);
      }
      e->moe_rows = (int)rows;
      if (int rc = make_tmap_act(&e->tm_moe_xg, e->moe_xg, (int)D, (int)rows)) return rc;
      if (int rc = make_tmap_act(&e->tm_moe_act, e->moe_act, (int)F, (int)rows)) return rc;
    }
    if (int rc = make_tmap_act(&e->tm_xn, e->tc_xn, (int)D, need)) return rc;
    if (int rc = make_tmap_act(&e->tm_ao, e->tc_ao, (int)qd, need)) return rc;
    if (c.n_experts == 0)
      if (int rc = make_tmap_act(&e->tm_act, e->tc_act, (int)F, need)) return rc;
  }
  return 0;
}

// one GEMM of the prefill: weight `widx` of the shard (see ensure_tm_w), tile configuration from (N, tokens in flight)
static int tc_gemm(xotb200_engine* e, const CUtensorMap& tmA, int widx, const TcGemmArgs& a, int epi, cudaStream_t st) {
  const TcCfg c = tc_cfg(a.N, epi == TCE_HEAD_LAST ? 1 : a.S);
  // ONE token tile (S <= 128): the GEMM streams the weight once and is bound by how many SMs pull on it.  N / NT tiles
  // of the narrow GEMMs (O, down: 32 tiles of 128; QKV: 48) leave most SMs idle and every CTA re-reads the whole [128, K]
  // activation tile; split-K puts ksplit CTAs on each tile, each reading only its K slice of both operands.
  static const bool use_splitk = !(getenv("XOTB200_TC_SPLITK") && atoi(getenv("XOTB200_TC_SPLITK")) == 0);
  if (use_splitk && a.S <= TC_M && a.mt0 == 0 && a.tile_info == nullptr && a.N % 128 == 0 &&
      (epi == TCE_STORE || epi == TCE_RESID || epi == TCE_QKV)) {
    const int tiles = a.N / 128, nkb = a.K / TC_BK;
    int ks = std::min(4, e->sm_count / std::max(1, tiles));
    while (ks > 1 && nkb % ks != 0) --ks;
    if (ks >= 2 && nkb / ks >= 4) {
      const size_t need = (size_t)tiles * ks * TC_M * 128 * 4;
      if (need > e->tc_ws_bytes) {
        CK(cudaStreamSynchronize(st));
        if (e->tc_ws) CK(cudaFree(e->tc_ws));
        e->tc_ws = nullptr;
        CK(cudaMalloc(&e->tc_ws, need));
        e->tc_ws_bytes = need;
      }
      if (!e->tc_ws_ticket) {
        CK(cudaMalloc(&e->tc_ws_ticket, 4096 * 4));
        MEMSET_SYNC(e, e->tc_ws_ticket, 0, 4096 * 4);
      }
      REQUIRE(tiles <= 4096, "split-K ticket table too small for N=%d", a.N);
      if (int rc = ensure_tm_w(e, 128)) return rc;
      TcGemmArgs b = a;
      b.ksplit = ks; b.ws = e->tc_ws; b.ws_ticket = e->tc_ws_ticket; b.err = &e->ctl->err;
      return launch_tc_gemm_nt<128, 6>(e, tmA, e->tm_w[128][widx], b, epi, st);
    }
  }
  // several 256-token tiles in flight and a 256-wide N tile: CTA pairs (each CTA stages half of the weight tile)
  static const bool use_2sm = !(getenv("XOTB200_TC_2SM") && atoi(getenv("XOTB200_TC_2SM")) == 0);
  if (use_2sm && epi != TCE_HEAD_LAST && a.tile_info == nullptr && a.mt0 == 0 && c.nt == 256 && a.S > 2 * TC_M && a.N % 256 == 0 &&
      ((a.S + 2 * TC_M - 1) / (2 * TC_M)) * (a.N / 256) >= e->sm_count / 4) {
    // (128-wide pair tiles for the narrow GEMMs — 2.6 waves instead of 1.3 for QKV — were measured SLOWER: 76 vs 62 us; the
    //  256-wide tile's operand economy outweighs the partial wave)
    if (int rc = ensure_tm_w(e, 128)) return rc;
    return launch_tc_gemm_2sm<256, 6>(e, tmA, e->tm_w[128][widx], a, epi, st);
  }
  if (int rc = ensure_tm_w(e, c.nt)) return rc;
  return launch_tc_gemm(e, tmA, e->tm_w[c.nt][widx], a, epi, c, st);
}

// one expert GEMM of a sparse-MoE prefill: rows [exp_info[2x], +128 * exp_info[2x+1]) of the gathered buffer, weight map `widx`;
// a.S = rows of the gathered buffer (all of them valid to write), grid.x = the most tiles one expert can get
static int tc_gemm_expert(xotb200_engine* e, const CUtensorMap& tmA, int widx, TcGemmArgs a, int max_tiles, int avg_rows, cudaStream_t st) {
  const TcCfg c = tc_cfg(a.N, avg_rows);
  if (int rc = ensure_tm_w(e, c.nt)) return rc;
  a.grid_m = max_tiles;
  return launch_tc_gemm(e, tmA, e->tm_w[c.nt][widx], a, TCE_STORE, c, st);
}

template <int HD>
static int launch_tc_attn(xotb200_engine* e, int S, int R, bf16* kc, bf16* vc, int pos0, cudaStream_t st) {
  constexpr int SMEM = (64 + 4 * 64) * HD * 2;          // Q tile + 2 x (K, V) tiles
  const bool f16 = e->cfg.dtype == XOTB200_DTYPE_F16;
  auto fn = f16 ? xb::fp::tc_flash_attn_kernel<HD> : xb::bf::tc_flash_attn_kernel<HD>;
  if (int rc = ensure_smem_attr(e->device, (const void*)fn, SMEM)) return rc;
  const float scale_log2 = 1.4426950408889634f / sqrtf((float)HD);
  fn<<<dim3((S + 63) / 64, e->cfg.n_heads), 128, SMEM, st>>>(e->tc_q, kc, vc, e->tc_ao, S, e->cfg.n_heads, R, e->L, pos0, scale_log2);
  CK(cudaGetLastError());
  e->launches++;
  return 0;
}

// Prefill n_tokens positions starting at pos0 on the tensor cores, in chunks of up to TC_CHUNK tokens (layer-major
// inside a chunk, so a chunk streams the shard's weights once).  Inputs: stage_tokens /
// stage_in; outputs: stage_out (non-last shard) or logits + ctl->token of the last position (last shard).
static int prefill_tc(xotb200_engine* e, int n_tokens, int pos0, cudaStream_t st) {
  if (int rc = prepare_tc(e, n_tokens)) return rc;
  const xotb200_model_cfg& c = e->cfg;
  const int D = c.dim, F = c.ffn, Hq = c.n_heads, Hkv = c.n_kv_heads, hd = c.head_dim;
  const int qd = Hq * hd, nqkv = (Hq + 2 * Hkv) * hd, nl = e->end - e->start + 1;
  for (int m0 = 0; m0 < n_tokens; m0 += TC_CHUNK) {
    const int S = std::min(TC_CHUNK, n_tokens - m0);
    if (e->first) {
      tc_gather_kernel<<<S, 128, 0, st>>>(e->stage_tokens + m0, e->embed, e->tc_x, D);
      CK(cudaGetLastError());
      e->launches++;
    } else {
      CK(cudaMemcpyAsync(e->tc_x, e->stage_in + (size_t)m0 * D, (size_t)S * D * 2, cudaMemcpyDeviceToDevice, st));
    }
    for (int li = 0; li < nl; ++li) {
      const LayerW& w = e->layers[li];
      bf16* kc = e->kv + (size_t)e->cur_req * e->kv_req_stride + ((size_t)li * 2 + 0) * Hkv * e->L * hd;
      bf16* vc = e->kv + (size_t)e->cur_req * e->kv_req_stride + ((size_t)li * 2 + 1) * Hkv * e->L * hd;
      DT_LAUNCH(e, (tc_rmsnorm_kernel<<<S, 256, 0, st>>>(e->tc_x, w.attn_norm, e->tc_xn, D, c.norm_eps)));
      CK(cudaGetLastError());
      if (e->tc_fuse) {     // RoPE + KV append in the QKV GEMM's epilogue: rotated q -> tc_q, k / v -> the cache rows of this chunk
        TcGemmArgs g{S, nqkv, D, e->tk, e->tc_q, nullptr, qd, 0};
        g.kc = kc; g.vc = vc; g.rope = e->rope; g.q_rows = qd; g.kv_rows = Hkv * hd; g.hd = hd; g.L = e->L; g.pos0 = pos0 + m0;
        if (int rc = tc_gemm(e, e->tm_xn, 4 * li + 0, g, TCE_QKV, st)) return rc;
      } else {
        TcGemmArgs g{S, nqkv, D, e->tk, e->tc_qkv, nullptr, nqkv, 0};
        if (int rc = tc_gemm(e, e->tm_xn, 4 * li + 0, g, TCE_STORE, st)) return rc;
        DT_LAUNCH(e, (tc_rope_kv_kernel<<<S, 256, 0, st>>>(e->tc_qkv, e->tc_q, kc, vc, e->rope, Hq, Hkv, hd, e->L, pos0 + m0)));
        CK(cudaGetLastError());
        e->launches++;
      }
      if (int rc = (hd == 128 ? launch_tc_attn<128>(e, S, Hq / Hkv, kc, vc, pos0 + m0, st)
                              : launch_tc_attn<64>(e, S, Hq / Hkv, kc, vc, pos0 + m0, st)))
        return rc;
      TcGemmArgs o{S, D, qd, e->tk, e->tc_x, nullptr, D, 0};
      if (int rc = tc_gemm(e, e->tm_ao, 4 * li + 1, o, TCE_RESID, st)) return rc;
      DT_LAUNCH(e, (tc_rmsnorm_kernel<<<S, 256, 0, st>>>(e->tc_x, w.mlp_norm, e->tc_xn, D, c.norm_eps)));
      CK(cudaGetLastError());
      if (c.n_experts > 0) {
        // ---- sparse MoE MLP: router -> plan -> gather -> per-expert gate/up -> SwiGLU -> per-expert down -> combine ----
        const int E = c.n_experts, rows = e->moe_rows, max_tiles = (S + TC_M - 1) / TC_M, avg_rows = std::max(1, 2 * S / E);
        int* rlog = (e->route_log && e->cur_req == 0) ? e->route_log + (size_t)(pos0 + m0) * nl : nullptr;
        DT_LAUNCH(e, (moe_router_kernel<<<S, 256, 0, st>>>(e->tc_xn, w.router, D, E, e->moe_route_e, e->moe_route_w, rlog, nl, li)));
        CK(cudaGetLastError());
        moe_plan_kernel<<<1, 1024, 0, st>>>(e->moe_route_e, S, E, rows, e->moe_row_of, e->moe_row_token, e->moe_exp_info);
        CK(cudaGetLastError());
        DT_LAUNCH(e, (moe_gather_kernel<<<rows, 128, 0, st>>>(e->tc_xn, e->moe_row_token, e->moe_exp_info, E, e->moe_xg, D)));
        CK(cudaGetLastError());
        for (int x = 0; x < E; ++x) {
          TcGemmArgs u{rows, 2 * F, D, e->tk, e->moe_gu, nullptr, 2 * F, 0, e->moe_exp_info + 2 * x};
          if (int rc = tc_gemm_expert(e, e->tm_moe_xg, 4 * nl + 1 + (li * E + x) * 2 + 0, u, max_tiles, avg_rows, st)) return rc;
        }
        DT_LAUNCH(e, (tc_swiglu_kernel<<<dim3(rows, (F + 2047) / 2048), 256, 0, st>>>(e->moe_gu, e->moe_act, F, e->moe_exp_info + 2 * E)));
        CK(cudaGetLastError());
        for (int x = 0; x < E; ++x) {
          TcGemmArgs d{rows, D, F, e->tk, e->moe_y, nullptr, D, 0, e->moe_exp_info + 2 * x};
          if (int rc = tc_gemm_expert(e, e->tm_moe_act, 4 * nl + 1 + (li * E + x) * 2 + 1, d, max_tiles, avg_rows, st)) return rc;
        }
        DT_LAUNCH(e, (moe_combine_kernel<<<S, 256, 0, st>>>(e->tc_x, e->moe_y, e->moe_row_of, e->moe_route_w, D)));
        CK(cudaGetLastError());
        e->launches += 7;   // 2 RMSNorm + router, plan, gather, SwiGLU, combine (RoPE/KV counted above when it is a kernel)
        continue;
      }
      if (e->tc_fuse) {     // SwiGLU in the gate/up GEMM's epilogue: [S, F] written instead of [S, 2F] + a kernel
        TcGemmArgs u{S, 2 * F, D, e->tk, e->tc_act, nullptr, F, 0};
        if (int rc = tc_gemm(e, e->tm_xn, 4 * li + 2, u, TCE_SWIGLU, st)) return rc;
      } else {
        TcGemmArgs u{S, 2 * F, D, e->tk, e->tc_gu, nullptr, 2 * F, 0};
        if (int rc = tc_gemm(e, e->tm_xn, 4 * li + 2, u, TCE_STORE, st)) return rc;
        DT_LAUNCH(e, (tc_swiglu_kernel<<<dim3(S, (F + 2047) / 2048), 256, 0, st>>>(e->tc_gu, e->tc_act, F, nullptr)));
        CK(cudaGetLastError());
        e->launches++;
      }
      TcGemmArgs d{S, D, F, e->tk, e->tc_x, nullptr, D, 0};
      if (int rc = tc_gemm(e, e->tm_act, 4 * li + 3, d, TCE_RESID, st)) return rc;
      e->launches += 2;   // the two RMSNorm kernels (the GEMM / attention launchers count themselves)
    }
    if (e->last) {
      if (m0 + S == n_tokens) {   // logits of the last position only (sharded_inference_engine.py:365-368)
        const int mt_last = (S - 1) / TC_M, r0 = mt_last * TC_M;   // only the tile holding the last position
        DT_LAUNCH(e, (tc_rmsnorm_kernel<<<S - r0, 256, 0, st>>>(e->tc_x + (size_t)r0 * D, e->final_norm, e->tc_xn + (size_t)r0 * D, D, c.norm_eps)));
        CK(cudaGetLastError());
        TcGemmArgs h{S, c.vocab, D, e->tk, nullptr, e->logits, 0, mt_last};
        if (int rc = tc_gemm(e, e->tm_xn, 4 * nl, h, TCE_HEAD_LAST, st)) return rc;
        tc_argmax_kernel<<<1, 1024, 0, st>>>(e->logits, c.vocab, e->ctl);
        CK(cudaGetLastError());
        e->launches += 2;
      }
    } else {
      CK(cudaMemcpyAsync(e->stage_out + (size_t)m0 * D, e->tc_x, (size_t)S * D * 2, cudaMemcpyDeviceToDevice, st));
    }
  }
  // the residual stream of the last position is what xotb200_post_to_next hands to the next shard
  CK(cudaMemcpyAsync(e->x, e->tc_x + (size_t)((n_tokens - 1) % TC_CHUNK) * D, (size_t)D * 2, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// --------------------------------------------------------------------------------------------------------------
// minimal JSON (safetensors header) parser
// --------------------------------------------------------------------------------------------------------------
struct StEntry {
  std::string dtype;
  std::vector<int64_t> shape;
  int64_t begin = 0, end = 0;
};
struct JsonP {
  const char* p;
  const char* e;
  bool ok = true;
  void ws() { while (p < e && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
  bool eat(char c) { ws(); if (p < e && *p == c) { ++p; return true; } return false; }
  std::string str() {
    ws();
    std::string s;
    if (p >= e || *p != '"') { ok = false; return s; }
    ++p;
    while (p < e && *p != '"') {
      if (*p == '\\' && p + 1 < e) { ++p; }
      s.push_back(*p++);
    }
    if (p < e) ++p; else ok = false;
    return s;
  }
  int64_t num() {
    ws();
    char* q = nullptr;
    long long v = strtoll(p, &q, 10);
    if (q == p) ok = false;
    p = q;
    return v;
  }
  void skip() {  // skip any value
    ws();
    if (p >= e) { ok = false; return; }
    if (*p == '"') { str(); return; }
    if (*p == '{') { ++p; if (eat('}')) return; do { str(); if (!eat(':')) { ok = false; return; } skip(); } while (ok && eat(',')); if (!eat('}')) ok = false; return; }
    if (*p == '[') { ++p; if (eat(']')) return; do { skip(); } while (ok && eat(',')); if (!eat(']')) ok = false; return; }
    while (p < e && *p != ',' && *p != '}' && *p != ']') ++p;
  }
};

static bool parse_st_header(const char* s, size_t n, std::map<std::string, StEntry>& out) {
  JsonP j{s, s + n};
  if (!j.eat('{')) return false;
  if (j.eat('}')) return true;
  do {
    std::string name = j.str();
    if (!j.ok || !j.eat(':')) return false;
    if (name == "__metadata__") { j.skip(); continue; }
    if (!j.eat('{')) return false;
    StEntry en;
    do {
      std::string k = j.str();
      if (!j.eat(':')) return false;
      if (k == "dtype") en.dtype = j.str();
      else if (k == "shape") {
        if (!j.eat('[')) return false;
        if (!j.eat(']')) { do { en.shape.push_back(j.num()); } while (j.eat(',')); if (!j.eat(']')) return false; }
      } else if (k == "data_offsets") {
        if (!j.eat('[')) return false;
        en.begin = j.num(); if (!j.eat(',')) return false; en.end = j.num(); if (!j.eat(']')) return false;
      } else j.skip();
    } while (j.ok && j.eat(','));
    if (!j.eat('}')) return false;
    out[name] = en;
  } while (j.ok && j.eat(','));
  return j.ok && j.eat('}');
}

// --------------------------------------------------------------------------------------------------------------
// helpers
// --------------------------------------------------------------------------------------------------------------
static void free_shard(xotb200_engine* e) {
  if (e->d_layers) { cudaFree(e->d_layers); e->d_layers = nullptr; }
  if (e->blob) { cudaFree(e->blob); e->blob = nullptr; }
  if (e->kv) { cudaFree(e->kv); e->kv = nullptr; }
  if (e->route_log) { cudaFree(e->route_log); e->route_log = nullptr; }
  e->layers.clear();
  e->slots.clear();
  e->mega = nullptr;
  e->tm_w.clear();
  e->embed = e->final_norm = e->lm_head = nullptr;
  e->start = e->end = -1;
  e->L = 0;
}

static int upload_slot(xotb200_engine* e, TensorSlot& s, const void* host) {
  const size_t bytes = (size_t)s.rows * s.K * 2;
  if (s.mode == 0 && s.stride == 1 && s.panel_tk == 0) {
    CK(cudaMemcpyAsync(s.dst + (size_t)s.off * s.K, host, bytes, cudaMemcpyHostToDevice, e->stream));
    CK(cudaStreamSynchronize(e->stream));
  } else {
    bf16* tmp = nullptr;
    CK(cudaMalloc(&tmp, bytes));
    cudaError_t er = cudaMemcpyAsync(tmp, host, bytes, cudaMemcpyHostToDevice, e->stream);
    if (er == cudaSuccess) {
      relayout_rows_kernel<<<s.rows, 128, 0, e->stream>>>(tmp, s.dst, s.rows, s.K, s.mode, e->cfg.head_dim, s.stride, s.off, s.n_total, s.panel_tk);
      er = cudaGetLastError();
      e->launches++;
    }
    cudaError_t er2 = cudaStreamSynchronize(e->stream);
    cudaFree(tmp);
    CK(er);
    CK(er2);
  }
  s.loaded = true;
  return 0;
}

// --------------------------------------------------------------------------------------------------------------
// Streaming shard loader: safetensors file (mmap, page cache) -> pinned double buffer -> HBM.
// The reference reads every *.safetensors of the directory fully into host tensors and copies module by module
// (llm_utils.py:205-218,230-318).  Here a tensor is cut into chunks of whole rows (<= STAGE_BYTES); worker threads copy chunk
// i+1 from the mapping into pinned buffer B while chunk i goes pinned buffer A -> device (DMA at PCIe rate, no pageable bounce
// through the driver's staging buffer) -> row re-layout kernel -> its place in the weight blob.  No per-tensor cudaMalloc /
// cudaFree / stream synchronise; one synchronise per file.
// --------------------------------------------------------------------------------------------------------------
struct StreamLoader {
  static constexpr size_t STAGE_BYTES = (size_t)64 << 20;
  xotb200_engine* e = nullptr;
  void* pinned[2] = {nullptr, nullptr};
  bf16* dtemp[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};      // buffer i may be overwritten once done[i] has completed
  int cur = 0, nthreads = 4;
  double t_memcpy = 0.0;
  int64_t bytes = 0;

  int init(xotb200_engine* eng) {
    e = eng;
    for (int i = 0; i < 2; ++i) {
      CK(cudaHostAlloc(&pinned[i], STAGE_BYTES, cudaHostAllocDefault));
      CK(cudaMalloc(&dtemp[i], STAGE_BYTES));
      CK(cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming));
    }
    const unsigned hc = std::thread::hardware_concurrency();
    nthreads = (int)std::max(1u, std::min(8u, hc ? hc / 2 : 4u));
    if (const char* v = getenv("XOTB200_LOAD_THREADS")) nthreads = std::max(1, std::min(32, atoi(v)));
    return 0;
  }
  ~StreamLoader() {
    for (int i = 0; i < 2; ++i) {
      if (done[i]) cudaEventDestroy(done[i]);
      if (dtemp[i]) cudaFree(dtemp[i]);
      if (pinned[i]) cudaFreeHost(pinned[i]);
    }
  }
  // parallel copy out of the file mapping (page-cache reads scale with threads; one thread tops out at a few GB/s)
  void copy_in(void* dst, const void* src, size_t n) {
    const auto t0 = std::chrono::steady_clock::now();
    const int T = (int)std::min<size_t>((size_t)nthreads, std::max<size_t>(1, n >> 20));
    if (T <= 1) memcpy(dst, src, n);
    else {
      std::vector<std::thread> th;
      const size_t per = (n / T + 4095) & ~(size_t)4095;
      for (int t = 0; t < T; ++t) {
        const size_t a = (size_t)t * per, b = std::min(n, a + per);
        if (a < b) th.emplace_back([=] { memcpy((char*)dst + a, (const char*)src + a, b - a); });
      }
      for (auto& x : th) x.join();
    }
    t_memcpy += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  }
  // one tensor of the file into one destination slot
  int upload(TensorSlot& s, const char* host) {
    const size_t row_bytes = (size_t)s.K * 2;
    const bool plain = s.mode == 0 && s.stride == 1 && s.panel_tk == 0;
    int unit = s.mode == 1 ? e->cfg.head_dim : 1;                                   // q/k interleave permutes rows inside a head
    int rows_per = (int)std::max<size_t>(1, STAGE_BYTES / row_bytes) / unit * unit;
    REQUIRE(rows_per >= unit, "a block of %d rows of %d columns does not fit the %zu-byte staging buffer", unit, s.K, STAGE_BYTES);
    for (int r0 = 0; r0 < s.rows; r0 += rows_per) {
      const int rows = std::min(rows_per, s.rows - r0);
      const size_t n = (size_t)rows * row_bytes;
      CK(cudaEventSynchronize(done[cur]));                                           // previous use of this buffer pair finished
      copy_in(pinned[cur], host + (size_t)r0 * row_bytes, n);
      if (plain) {
        CK(cudaMemcpyAsync(s.dst + (size_t)(s.off + r0) * s.K, pinned[cur], n, cudaMemcpyHostToDevice, e->stream));
      } else {
        CK(cudaMemcpyAsync(dtemp[cur], pinned[cur], n, cudaMemcpyHostToDevice, e->stream));
        relayout_rows_kernel<<<rows, 128, 0, e->stream>>>(dtemp[cur], s.dst, rows, s.K, s.mode, e->cfg.head_dim, s.stride,
                                                          s.off + r0 * s.stride, s.n_total, s.panel_tk);
        CK(cudaGetLastError());
        e->launches++;
      }
      CK(cudaEventRecord(done[cur], e->stream));
      cur ^= 1;
      bytes += (int64_t)n;
    }
    s.loaded = true;
    return 0;
  }
};

// --------------------------------------------------------------------------------------------------------------
// C ABI
// --------------------------------------------------------------------------------------------------------------
extern "C" {

const char* xotb200_last_error(void) { return g_err; }
const char* xotb200_version(void) { return "xotb200 0.1.0 sm_100a"; }

int xotb200_create(const xotb200_model_cfg* cfg, int device, xotb200_engine** out) {
  REQUIRE(cfg && out, "null argument");
  REQUIRE(cfg->dtype == XOTB200_DTYPE_BF16 || cfg->dtype == XOTB200_DTYPE_F16, "dtype must be bf16 (0) or f16 (1), got %d", cfg->dtype);
  REQUIRE(cfg->head_dim == 64 || cfg->head_dim == 128, "head_dim must be 64 or 128 (got %d)", cfg->head_dim);
  REQUIRE(cfg->n_heads % cfg->n_kv_heads == 0, "n_heads %% n_kv_heads != 0");
  const int R = cfg->n_heads / cfg->n_kv_heads;
  REQUIRE(R == 1 || R == 2 || R == 4 || R == 8, "GQA ratio %d unsupported", R);
  REQUIRE(cfg->dim % 512 == 0 && cfg->ffn % 512 == 0 && (cfg->n_heads * cfg->head_dim) % 512 == 0,
          "dim, ffn and n_heads*head_dim must be multiples of 512");
  REQUIRE(cfg->vocab % 2 == 0, "vocab must be even (row pairs)");
  REQUIRE(cfg->n_experts == 0 || (cfg->n_experts >= 2 && cfg->n_experts <= NCW && cfg->experts_per_tok == 2),
          "sparse MoE: 2..%d experts with 2 experts per token are implemented (got %d / %d)", NCW, cfg->n_experts, cfg->experts_per_tok);
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  REQUIRE(device >= 0 && device < ndev, "device %d out of range (%d visible)", device, ndev);
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  REQUIRE(prop.major == 10, "libxotb200 is sm_100a-only; device %d is sm_%d%d", device, prop.major, prop.minor);
  xotb200_engine* e = new xotb200_engine();
  e->cfg = *cfg;
  e->device = device;
  e->sm_count = prop.multiProcessorCount;
  const int D = cfg->dim, Hq = cfg->n_heads, Hkv = cfg->n_kv_heads, hd = cfg->head_dim;
  e->nsplit = std::max(1, std::min(32, e->sm_count / Hkv));
  if (const char* v = getenv("XOTB200_TC_FUSE")) e->tc_fuse = atoi(v) != 0;
  CK(cudaStreamCreateWithFlags(&e->stream, cudaStreamNonBlocking));
  CK(cudaEventCreate(&e->ev0));
  CK(cudaEventCreate(&e->ev1));
  CK(cudaMalloc(&e->x, (size_t)D * 2));
  CK(cudaMalloc(&e->q, (size_t)Hq * hd * 2));
  CK(cudaMalloc(&e->attn_out, (size_t)Hq * hd * 2));
  CK(cudaMalloc(&e->act, (size_t)cfg->ffn * 2 * (cfg->n_experts > 0 ? 2 : 1)));   // MoE: both selected experts' activations
  CK(cudaMalloc(&e->logits, (size_t)cfg->vocab * 4));
  CK(cudaMalloc(&e->part_ml, (size_t)Hq * e->nsplit * 2 * 4));
  CK(cudaMalloc(&e->part_acc, (size_t)Hq * e->nsplit * hd * 4));
  CK(cudaMalloc(&e->tickets, (Hkv + 1) * 4));
  MEMSET_SYNC(e, e->tickets, 0, (Hkv + 1) * 4);
  CK(cudaMalloc(&e->cta_best_val, 1024 * 4));
  CK(cudaMalloc(&e->cta_best_idx, 1024 * 4));
  CK(cudaMalloc(&e->bar_counter, 4));
  CK(cudaMalloc(&e->ctl, sizeof(StepCtl)));
  MEMSET_SYNC(e, e->ctl, 0, sizeof(StepCtl));
  const size_t inbox_bytes = sizeof(InboxHdr) + (size_t)(INBOX_SLOTS + API_SLOTS) * D * 2;
  CK(cudaMalloc(&e->inbox, inbox_bytes));
  MEMSET_SYNC(e, e->inbox, 0, inbox_bytes);
  CK(cudaMalloc(&e->tokens_log, MAX_RING_STEPS * 4));
  CK(cudaMalloc(&e->d_req_map, INBOX_SLOTS * 4));
  CK(cudaMalloc(&e->d_pos_tab, INBOX_SLOTS * 4));
  {
    void* hp = nullptr;
    CK(cudaHostAlloc(&hp, (MAX_RING_STEPS + 16) * 4, cudaHostAllocMapped | cudaHostAllocPortable));
    memset(hp, 0, (MAX_RING_STEPS + 16) * 4);
    e->tok_host = static_cast<int*>(hp) + 16;                       // [0]: counter (own cache line), [16..]: token ids
    e->tok_count_host = static_cast<volatile unsigned int*>(hp);
    void* dp = nullptr;
    CK(cudaHostGetDevicePointer(&dp, hp, 0));
    e->tok_count_dev = static_cast<unsigned int*>(dp);
    e->tok_host_dev = static_cast<int*>(dp) + 16;
  }
  CK(cudaMalloc(&e->t_in, MAX_RING_STEPS * 8));
  CK(cudaMalloc(&e->t_out, MAX_RING_STEPS * 8));
  MEMSET_SYNC(e, e->t_in, 0, MAX_RING_STEPS * 8);
  MEMSET_SYNC(e, e->t_out, 0, MAX_RING_STEPS * 8);
  *out = e;
  return 0;
}

int xotb200_destroy(xotb200_engine* e) {
  if (!e) return 0;
  cudaSetDevice(e->device);
  cudaDeviceSynchronize();
  free_shard(e);
  if (e->next_ipc && e->next_inbox) cudaIpcCloseMemHandle(e->next_inbox);
  void* tcp[] = {e->tc_ws, e->tc_ws_ticket, e->tc_x, e->tc_xn, e->tc_qkv, e->tc_q, e->tc_ao, e->tc_gu, e->tc_act, e->moe_xg, e->moe_gu, e->moe_act, e->moe_y,
                 e->moe_route_e, e->moe_route_w, e->moe_row_of, e->moe_row_token, e->moe_exp_info};
  for (void* p : tcp)
    if (p) cudaFree(p);
  for (bf16* p : e->hop_buf)
    if (p) cudaFree(p);
  if (e->tok_count_host) cudaFreeHost(const_cast<unsigned int*>(e->tok_count_host));
  void* ptrs[] = {e->d_req_map, e->d_pos_tab, e->x, e->q, e->attn_out, e->act, e->logits, e->part_ml, e->part_acc, e->tickets, e->cta_best_val,
                  e->cta_best_idx, e->ctl, e->bar_counter, e->dbg, e->inbox, e->tokens_log, e->t_in, e->t_out, e->rope, e->stage_tokens,
                  e->stage_f32, e->stage_in, e->stage_out};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (e->ev0) cudaEventDestroy(e->ev0);
  if (e->ev1) cudaEventDestroy(e->ev1);
  if (e->stream) cudaStreamDestroy(e->stream);
  delete e;
  return 0;
}

int xotb200_set_shard(xotb200_engine* e, int start_layer, int end_layer) {
  REQUIRE(e, "null engine");
  const xotb200_model_cfg& c = e->cfg;
  REQUIRE(0 <= start_layer && start_layer <= end_layer && end_layer < c.n_layers, "bad layer range [%d, %d] of %d",
          start_layer, end_layer, c.n_layers);
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  free_shard(e);
  e->start = start_layer;
  e->end = end_layer;
  e->first = start_layer == 0;
  e->last = end_layer == c.n_layers - 1;
  const size_t D = c.dim, F = c.ffn, Hq = c.n_heads, Hkv = c.n_kv_heads, hd = c.head_dim, V = c.vocab;
  const size_t nl = end_layer - start_layer + 1;
  const size_t A = 256;
  const size_t sz_norm = align_up(D * 2, A), sz_qkv = align_up((Hq + 2 * Hkv) * hd * D * 2, A),
               sz_o = align_up(D * Hq * hd * 2, A), sz_gu = align_up(2 * F * D * 2, A), sz_dn = align_up(D * F * 2, A);
  const size_t E = c.n_experts, nexp = E ? E : 1;
  const size_t sz_router = E ? align_up(E * D * 2, A) : 0;
  const size_t per_layer = 2 * sz_norm + sz_qkv + sz_o + nexp * (sz_gu + sz_dn) + sz_router;
  const size_t sz_emb = align_up(V * D * 2, A);
  const bool own_head = e->last;   // the head is streamed in K-panel layout, the embedding is gathered row-wise: a tied
                                   // model keeps both copies when one shard holds both
  const int qdim = c.n_heads * c.head_dim;
  e->tk = (c.dim % 2048 == 0 && c.ffn % 2048 == 0 && qdim % 2048 == 0) ? 2048 : 512;
  const int tk = e->tk;
  size_t total = nl * per_layer + (e->first ? sz_emb : 0) + (e->last ? sz_norm : 0) + (own_head ? sz_emb : 0);
  CK(cudaMalloc(&e->blob, total));
  e->blob_bytes = total;
  char* p = reinterpret_cast<char*>(e->blob);
  auto take = [&](size_t n) { char* r = p; p += n; return reinterpret_cast<bf16*>(r); };
  if (e->first) e->embed = take(sz_emb);
  e->layers.resize(nl);
  const int nqkv = (int)((Hq + 2 * Hkv) * hd);
  for (size_t li = 0; li < nl; ++li) {
    LayerW& w = e->layers[li];
    w.attn_norm = take(sz_norm); w.wqkv = take(sz_qkv); w.wo = take(sz_o);
    w.mlp_norm = take(sz_norm); w.wgu = take(nexp * sz_gu); w.wdown = take(nexp * sz_dn);
    w.router = E ? take(sz_router) : nullptr;
    const int l = start_layer + (int)li;
    const std::string pre = "model.layers." + std::to_string(l) + ".";
    const int t = l * 16;
    auto add = [&](const char* nm, bf16* dst, int rows, int K, int mode, int stride, int off, int kind, float mean,
                   int n_total, int ptk) {
      e->slots.insert({pre + nm, TensorSlot{dst, rows, K, mode, stride, off, t + kind, 0.02f, mean, false, n_total, ptk}});
    };
    add("input_layernorm.weight", w.attn_norm, 1, (int)D, 0, 1, 0, 0, 1.0f, 1, 0);
    add("self_attn.q_proj.weight", w.wqkv, (int)(Hq * hd), (int)D, 1, 1, 0, 1, 0.f, nqkv, tk);
    add("self_attn.k_proj.weight", w.wqkv, (int)(Hkv * hd), (int)D, 1, 1, (int)(Hq * hd), 2, 0.f, nqkv, tk);
    add("self_attn.v_proj.weight", w.wqkv, (int)(Hkv * hd), (int)D, 0, 1, (int)((Hq + Hkv) * hd), 3, 0.f, nqkv, tk);
    add("self_attn.o_proj.weight", w.wo, (int)D, (int)(Hq * hd), 0, 1, 0, 4, 0.f, (int)D, tk);
    add("post_attention_layernorm.weight", w.mlp_norm, 1, (int)D, 0, 1, 0, 5, 1.0f, 1, 0);
    if (E) {
      // on-disk Mixtral names (w1 = gate, w3 = up, w2 = down); tensor ids as in synth.py: G_MOE + layer*64 + ...
      const int tm = (1 << 21) + l * 64 - t;
      add("block_sparse_moe.gate.weight", w.router, (int)E, (int)D, 0, 1, 0, tm, 0.f, (int)E, 0);
      for (int x = 0; x < (int)E; ++x) {
        const std::string ex = "block_sparse_moe.experts." + std::to_string(x) + ".";
        bf16* gu = w.wgu + (size_t)x * (sz_gu / 2);
        bf16* dn = w.wdown + (size_t)x * (sz_dn / 2);
        add((ex + "w1.weight").c_str(), gu, (int)F, (int)D, 0, 2, 0, tm + 1 + 3 * x, 0.f, (int)(2 * F), tk);
        add((ex + "w3.weight").c_str(), gu, (int)F, (int)D, 0, 2, 1, tm + 2 + 3 * x, 0.f, (int)(2 * F), tk);
        add((ex + "w2.weight").c_str(), dn, (int)D, (int)F, 0, 1, 0, tm + 3 + 3 * x, 0.f, (int)D, tk);
      }
    } else {
      add("mlp.gate_proj.weight", w.wgu, (int)F, (int)D, 0, 2, 0, 6, 0.f, (int)(2 * F), tk);
      add("mlp.up_proj.weight", w.wgu, (int)F, (int)D, 0, 2, 1, 7, 0.f, (int)(2 * F), tk);
      add("mlp.down_proj.weight", w.wdown, (int)D, (int)F, 0, 1, 0, 8, 0.f, (int)D, tk);
    }
  }
  if (e->first)
    e->slots.insert({"model.embed_tokens.weight", TensorSlot{e->embed, (int)V, (int)D, 0, 1, 0, (1 << 20), 0.02f, 0.f, false, (int)V, 0}});
  if (e->last) {
    e->final_norm = take(sz_norm);
    e->slots.insert({"model.norm.weight", TensorSlot{e->final_norm, 1, (int)D, 0, 1, 0, (1 << 20) + 1, 0.02f, 1.0f, false, 1, 0}});
    e->lm_head = take(sz_emb);
    if (c.tie_embeddings)   // tied: the head IS the embedding table (general_mha.py:42-44,125-126), stored again as K-panels
      e->slots.insert({"model.embed_tokens.weight", TensorSlot{e->lm_head, (int)V, (int)D, 0, 1, 0, (1 << 20), 0.02f, 0.f, false, (int)V, tk}});
    else
      e->slots.insert({"lm_head.weight", TensorSlot{e->lm_head, (int)V, (int)D, 0, 1, 0, (1 << 20) + 2, 0.02f, 0.f, false, (int)V, tk}});
  }
  return 0;
}

int xotb200_load_tensor(xotb200_engine* e, const char* hf_name, const void* host_data, const int64_t* shape, int ndim) {
  REQUIRE(e && hf_name && host_data && shape, "null argument");
  REQUIRE(e->start >= 0, "no shard set");
  CK(cudaSetDevice(e->device));
  auto range = e->slots.equal_range(hf_name);
  if (range.first == range.second) return 1;  // not part of this shard
  const int64_t rows = ndim == 2 ? shape[0] : 1, K = ndim == 2 ? shape[1] : shape[0];
  for (auto it = range.first; it != range.second; ++it) {
    TensorSlot& s = it->second;
    REQUIRE((ndim == 1 || ndim == 2) && rows == s.rows && K == s.K, "tensor %s: shape mismatch (got %lld x %lld, want %d x %d)",
            hf_name, (long long)rows, (long long)K, s.rows, s.K);
    if (int rc = upload_slot(e, s, host_data)) return rc;
  }
  return 0;
}

int xotb200_check_loaded(xotb200_engine* e) {
  REQUIRE(e && e->start >= 0, "no shard set");
  for (auto& kv : e->slots)
    if (!kv.second.loaded) {
      set_err("strict loader: tensor '%s' of shard [%d, %d] was never loaded", kv.first.c_str(), e->start, e->end);
      return XOTB200_ESTRICT;
    }
  return 0;
}

int xotb200_load_shard(xotb200_engine* e, const char* dir, int start_layer, int end_layer) {
  REQUIRE(e && dir, "null argument");
  if (int rc = xotb200_set_shard(e, start_layer, end_layer)) return rc;
  DIR* d = opendir(dir);
  if (!d) { set_err("cannot open directory %s", dir); return XOTB200_EIO; }
  std::vector<std::string> files;
  while (struct dirent* de = readdir(d)) {
    std::string n = de->d_name;
    if (n.size() > 12 && n.substr(n.size() - 12) == ".safetensors") files.push_back(std::string(dir) + "/" + n);
  }
  closedir(d);
  std::sort(files.begin(), files.end());
  if (files.empty()) { set_err("no *.safetensors in %s", dir); return XOTB200_EIO; }
  CK(cudaSetDevice(e->device));
  const bool streaming = !(getenv("XOTB200_LOAD_STREAM") && atoi(getenv("XOTB200_LOAD_STREAM")) == 0);   // 0: the round-1 per-tensor path (A/B)
  StreamLoader ld;
  if (streaming)
    if (int rc = ld.init(e)) return rc;
  const auto t_load0 = std::chrono::steady_clock::now();
  e->load_bytes = 0;
  for (auto& f : files) {
    int fd = open(f.c_str(), O_RDONLY);
    if (fd < 0) { set_err("cannot open %s", f.c_str()); return XOTB200_EIO; }
    struct stat sb;
    if (fstat(fd, &sb) != 0 || sb.st_size < 8) { close(fd); set_err("cannot stat %s (or shorter than a safetensors header)", f.c_str()); return XOTB200_EIO; }
    void* m = mmap(nullptr, sb.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (m == MAP_FAILED) { set_err("mmap failed for %s", f.c_str()); return XOTB200_EIO; }
    posix_madvise(m, sb.st_size, POSIX_MADV_SEQUENTIAL);
    const char* base = static_cast<const char*>(m);
    uint64_t hlen = 0;
    memcpy(&hlen, base, 8);
    std::map<std::string, StEntry> hdr;
    int rc = 0;
    if (hlen > (uint64_t)sb.st_size - 8 || !parse_st_header(base + 8, hlen, hdr)) {
      set_err("bad safetensors header in %s", f.c_str());
      rc = XOTB200_EIO;
    }
    const uint64_t data_bytes = (uint64_t)sb.st_size - 8 - hlen;
    for (auto it = hdr.begin(); rc == 0 && it != hdr.end(); ++it) {
      auto sl = e->slots.find(it->first);
      if (sl == e->slots.end()) {
        // strict: a tensor of a layer this shard OWNS that has no slot (e.g. q_proj.bias of a biased-attention
        // checkpoint) would be silently ignored and give wrong logits -> refuse the checkpoint
        int l = -1;
        if (sscanf(it->first.c_str(), "model.layers.%d.", &l) == 1 && l >= e->start && l <= e->end &&
            it->first.find("rotary_emb.inv_freq") == std::string::npos) {
          set_err("strict loader: tensor '%s' belongs to layer %d of this shard but the engine has no use for it (unsupported architecture variant)",
                  it->first.c_str(), l);
          rc = XOTB200_ESTRICT;
        }
        continue;
      }
      const StEntry& en = it->second;
      const char* want = e->cfg.dtype == XOTB200_DTYPE_F16 ? "F16" : "BF16";
      if (en.dtype != want) { set_err("tensor %s: dtype %s but the engine was created for %s", it->first.c_str(), en.dtype.c_str(), want); rc = XOTB200_EIO; break; }
      uint64_t elems = 1;
      bool shape_ok = !en.shape.empty() && en.shape.size() <= 2;
      for (int64_t d : en.shape) { if (d <= 0 || d > (int64_t)1 << 40) shape_ok = false; else elems *= (uint64_t)d; }
      if (!shape_ok || en.begin < 0 || en.end < en.begin || (uint64_t)en.end > data_bytes || (uint64_t)(en.end - en.begin) != elems * 2) {
        set_err("tensor %s: data_offsets [%lld, %lld) do not match its shape / the file size", it->first.c_str(), (long long)en.begin, (long long)en.end);
        rc = XOTB200_EIO;
        break;
      }
      const char* data = base + 8 + hlen + en.begin;
      if (!streaming) {
        rc = xotb200_load_tensor(e, it->first.c_str(), data, en.shape.data(), (int)en.shape.size());
        if (rc == 1) rc = 0;
        e->load_bytes += en.end - en.begin;
        continue;
      }
      const int64_t rows = en.shape.size() == 2 ? en.shape[0] : 1, K = en.shape.size() == 2 ? en.shape[1] : en.shape[0];
      auto range = e->slots.equal_range(it->first);
      for (auto sl2 = range.first; rc == 0 && sl2 != range.second; ++sl2) {       // a tied embedding feeds two destinations
        TensorSlot& ts = sl2->second;
        if (rows != ts.rows || K != ts.K) {
          set_err("tensor %s: shape mismatch (got %lld x %lld, want %d x %d)", it->first.c_str(), (long long)rows, (long long)K, ts.rows, ts.K);
          rc = XOTB200_EINVAL;
          break;
        }
        rc = ld.upload(ts, data);
      }
    }
    if (cudaStreamSynchronize(e->stream) != cudaSuccess && rc == 0) { set_err("stream error while loading %s", f.c_str()); rc = XOTB200_ECUDA; }
    munmap(m, sb.st_size);                        // only after every copy out of the mapping has completed
    if (rc) return rc;
  }
  if (streaming) e->load_bytes = ld.bytes;
  e->load_seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_load0).count();
  e->load_memcpy_seconds = ld.t_memcpy;
  return xotb200_check_loaded(e);
}

int xotb200_init_synthetic(xotb200_engine* e, uint64_t seed) {
  REQUIRE(e && e->start >= 0, "no shard set");
  CK(cudaSetDevice(e->device));
  const double ih4_std = sqrt(4.0 * (65536.0 * 65536.0 - 1.0) / 12.0);
  for (auto& kv : e->slots) {
    TensorSlot& s = kv.second;
    const unsigned long long base = seed * 0x9E3779B97F4A7C15ull + (unsigned long long)s.tensor_id * 0xD1B54A32D192ED03ull;
    const float scale = (float)((double)s.std / ih4_std);
    DT_LAUNCH(e, (synth_rows_kernel<<<s.rows, 256, 0, e->stream>>>(s.dst, s.rows, s.K, s.mode, e->cfg.head_dim, s.stride, s.off, s.n_total, s.panel_tk, base, scale, s.mean)));
    CK(cudaGetLastError());
    e->launches++;
    s.loaded = true;
  }
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

int xotb200_begin_batch(xotb200_engine* e, int max_total_len, int n_requests) {
  REQUIRE(e && e->start >= 0, "no shard set");
  REQUIRE(n_requests >= 1 && n_requests <= INBOX_SLOTS - 1, "n_requests must be 1..%d", INBOX_SLOTS - 1);
  REQUIRE(max_total_len > 0 && max_total_len <= e->cfg.max_pos, "max_total_len %d out of range (max_position_embeddings %d)",
          max_total_len, e->cfg.max_pos);
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  const xotb200_model_cfg& c = e->cfg;
  const size_t nl = e->end - e->start + 1;
  if (max_total_len > e->L || n_requests != e->nreq) {
    if (e->kv) { CK(cudaFree(e->kv)); e->kv = nullptr; }
    e->L = std::max(e->L, max_total_len);
    e->nreq = n_requests;
    e->kv_req_stride = nl * 2 * (size_t)c.n_kv_heads * e->L * c.head_dim;
    CK(cudaMalloc(&e->kv, (size_t)e->nreq * e->kv_req_stride * 2));
    if (c.n_experts > 0) {
      if (e->route_log) { CK(cudaFree(e->route_log)); e->route_log = nullptr; }
      CK(cudaMalloc(&e->route_log, (size_t)e->L * nl * 4));
    }
  }
  if (e->route_log) CK(cudaMemsetAsync(e->route_log, 0xff, (size_t)e->L * nl * 4, e->stream));
  if (max_total_len > e->rope_len) {
    // RoPE table: theta_j in fp32, angle = fp32(pos * theta_j), cos/sin correctly rounded
    const int half = c.head_dim / 2;
    std::vector<float> theta(half);
    for (int j = 0; j < half; ++j) {
      float ex = (float)(2 * j) / (float)c.head_dim;
      float f = 1.0f / powf(c.rope_theta, ex);
      if (c.rope_kind == XOTB200_ROPE_LLAMA3) {
        const double low_wl = (double)c.rope_old_context_len / c.rope_low_freq_factor;
        const double high_wl = (double)c.rope_old_context_len / c.rope_high_freq_factor;
        const double wavelen = 2.0 * M_PI / (double)f;
        if (wavelen < high_wl) {
        } else if (wavelen > low_wl) {
          f = f / c.rope_factor;
        } else {
          const double smooth = ((double)c.rope_old_context_len / wavelen - c.rope_low_freq_factor) /
                                (c.rope_high_freq_factor - c.rope_low_freq_factor);
          f = (float)((1.0 - smooth) * (double)f / c.rope_factor + smooth * (double)f);
        }
      }
      theta[j] = f;
    }
    std::vector<float2> tab((size_t)max_total_len * half);
    for (int p = 0; p < max_total_len; ++p)
      for (int j = 0; j < half; ++j) {
        const float ang = (float)p * theta[j];
        tab[(size_t)p * half + j] = make_float2((float)cos((double)ang), (float)sin((double)ang));
      }
    if (e->rope) { CK(cudaFree(e->rope)); e->rope = nullptr; }
    CK(cudaMalloc(&e->rope, tab.size() * sizeof(float2)));
    MEMCPY_SYNC(e, e->rope, tab.data(), tab.size() * sizeof(float2), cudaMemcpyHostToDevice);
    e->rope_len = max_total_len;
  }
  {  // device-side layer table (weights + this request's KV pointers)
    std::vector<LayerDev> h(nl);
    for (size_t li = 0; li < nl; ++li) {
      const LayerW& w = e->layers[li];
      h[li] = LayerDev{w.attn_norm, w.wqkv, w.wo, w.mlp_norm, w.wgu, w.wdown, w.router,
                       e->kv + (li * 2 + 0) * (size_t)c.n_kv_heads * e->L * c.head_dim,
                       e->kv + (li * 2 + 1) * (size_t)c.n_kv_heads * e->L * c.head_dim};
    }
    if (!e->d_layers) CK(cudaMalloc(&e->d_layers, nl * sizeof(LayerDev)));
    MEMCPY_SYNC(e, e->d_layers, h.data(), nl * sizeof(LayerDev), cudaMemcpyHostToDevice);
  }
  // reset_caches (llm_utils.py:409-414): zero like torchtune's KVCache.reset
  CK(cudaMemsetAsync(e->kv, 0, (size_t)e->nreq * e->kv_req_stride * 2, e->stream));
  e->cur_req = 0;
  CK(cudaMemsetAsync(e->ctl, 0, sizeof(StepCtl), e->stream));
  CK(cudaStreamSynchronize(e->stream));
  e->have_result = false;
  return 0;
}

int xotb200_begin_request(xotb200_engine* e, int max_total_len) { return xotb200_begin_batch(e, max_total_len, 1); }

int xotb200_select_request(xotb200_engine* e, int request) {
  REQUIRE(e && request >= 0 && request < e->nreq, "request index out of range (batch of %d)", e ? e->nreq : 0);
  e->cur_req = request;
  return 0;
}

// sticky device-side error word -> status code (3: a bounded wait inside the persistent tensor-core GEMM expired)
static int check_device_error(xotb200_engine* e) {
  StepCtl h;
  CK(cudaMemcpyAsync(&h, e->ctl, sizeof(h), cudaMemcpyDeviceToHost, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  if (h.err) {
    MEMSET_SYNC(e, &e->ctl->err, 0, 4);
    set_err(h.err == 3 ? "tensor-core GEMM: a pipeline wait expired (device error %d)" : "device-side wait timed out (device error %d)", h.err);
    return XOTB200_ETIMEOUT;
  }
  return 0;
}

static int ensure_staging(xotb200_engine* e, int n) {
  if (n <= e->stage_cap) return 0;
  const size_t D = e->cfg.dim;
  for (void* p : {(void*)e->stage_tokens, (void*)e->stage_f32, (void*)e->stage_in, (void*)e->stage_out})
    if (p) cudaFree(p);
  e->stage_tokens = nullptr; e->stage_f32 = nullptr; e->stage_in = e->stage_out = nullptr;
  e->stage_cap = 0;
  CK(cudaMalloc(&e->stage_tokens, (size_t)n * 4));
  CK(cudaMalloc(&e->stage_f32, (size_t)n * D * 4));
  CK(cudaMalloc(&e->stage_in, (size_t)n * D * 2));
  CK(cudaMalloc(&e->stage_out, (size_t)n * D * 2));
  e->stage_cap = n;
  return 0;
}

int xotb200_forward(xotb200_engine* e, const int32_t* tokens, const float* hidden_f32, int n_tokens, int pos0,
                    float* out_f32) {
  REQUIRE(e && e->start >= 0, "no shard set");
  REQUIRE((tokens != nullptr) != (hidden_f32 != nullptr), "exactly one of tokens / hidden must be given");
  REQUIRE((tokens != nullptr) == e->first, "%s shard expects %s", e->first ? "first" : "non-first",
          e->first ? "token ids" : "a hidden state");
  REQUIRE(n_tokens >= 1 && pos0 >= 0, "bad n_tokens/pos0");
  REQUIRE(e->L > 0, "begin_request was not called");
  REQUIRE(pos0 + n_tokens <= e->L, "KV cache overflow: pos0 %d + n %d > capacity %d", pos0, n_tokens, e->L);
  REQUIRE(out_f32, "null output");
  if (int rc = xotb200_check_loaded(e)) return rc;
  CK(cudaSetDevice(e->device));
  if (int rc = ensure_staging(e, n_tokens)) return rc;
  const size_t D = e->cfg.dim;
  cudaStream_t st = e->stream;
  if (tokens) {
    for (int i = 0; i < n_tokens; ++i)
      REQUIRE(tokens[i] >= 0 && tokens[i] < e->cfg.vocab, "token id %d out of range", tokens[i]);
    CK(cudaMemcpyAsync(e->stage_tokens, tokens, (size_t)n_tokens * 4, cudaMemcpyHostToDevice, st));
  } else {
    CK(cudaMemcpyAsync(e->stage_f32, hidden_f32, (size_t)n_tokens * D * 4, cudaMemcpyHostToDevice, st));
    DT_LAUNCH(e, (f32_to_bf16_kernel<<<256, 256, 0, st>>>(e->stage_f32, e->stage_in, (size_t)n_tokens * D)));
    CK(cudaGetLastError());
    e->launches++;
  }
  const bool use_tc = e->prefill_mode == 2 || (e->prefill_mode == 0 && n_tokens >= 8);
  if (use_tc) {
    if (int rc = prefill_tc(e, n_tokens, pos0, st)) return rc;     // GEMM-shaped: 5th-gen tensor cores
  } else {
    if (int rc = launch_mega(e, 1, n_tokens, 0, pos0, st)) return rc;   // one launch walks all n_tokens positions
  }
  if (e->last) {
    CK(cudaMemcpyAsync(out_f32, e->logits, (size_t)e->cfg.vocab * 4, cudaMemcpyDeviceToHost, st));
  } else {
    DT_LAUNCH(e, (bf16_to_f32_kernel<<<256, 256, 0, st>>>(e->stage_out, e->stage_f32, (size_t)n_tokens * D)));
    CK(cudaGetLastError());
    e->launches++;
    CK(cudaMemcpyAsync(out_f32, e->stage_f32, (size_t)n_tokens * D * 4, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  if (use_tc)
    if (int rc = check_device_error(e)) return rc;
  e->have_result = true;
  return 0;
}

// Same step with the hidden state staying in device memory on both sides (16-bit, the engine's dtype): the hop between
// shards is then 2*S*D bytes moved GPU to GPU by the caller (NCCL send/recv, cudaMemcpyPeer) instead of a host round trip.
int xotb200_forward_device(xotb200_engine* e, const int32_t* tokens, const void* hidden_dev, int n_tokens, int pos0,
                           void* out_hidden_dev, float* logits_f32) {
  REQUIRE(e && e->start >= 0, "no shard set");
  REQUIRE((tokens != nullptr) != (hidden_dev != nullptr), "exactly one of tokens / hidden must be given");
  REQUIRE((tokens != nullptr) == e->first, "%s shard expects %s", e->first ? "first" : "non-first",
          e->first ? "token ids" : "a hidden state");
  REQUIRE(n_tokens >= 1 && pos0 >= 0, "bad n_tokens/pos0");
  REQUIRE(e->L > 0, "begin_request was not called");
  REQUIRE(pos0 + n_tokens <= e->L, "KV cache overflow: pos0 %d + n %d > capacity %d", pos0, n_tokens, e->L);
  REQUIRE(e->last ? (logits_f32 != nullptr) : (out_hidden_dev != nullptr), "null output");
  if (int rc = xotb200_check_loaded(e)) return rc;
  CK(cudaSetDevice(e->device));
  if (int rc = ensure_staging(e, n_tokens)) return rc;
  const size_t D = e->cfg.dim;
  cudaStream_t st = e->stream;
  if (tokens) {
    for (int i = 0; i < n_tokens; ++i)
      REQUIRE(tokens[i] >= 0 && tokens[i] < e->cfg.vocab, "token id %d out of range", tokens[i]);
    CK(cudaMemcpyAsync(e->stage_tokens, tokens, (size_t)n_tokens * 4, cudaMemcpyHostToDevice, st));
  } else {
    CK(cudaMemcpyAsync(e->stage_in, hidden_dev, (size_t)n_tokens * D * 2, cudaMemcpyDeviceToDevice, st));
  }
  const bool use_tc = e->prefill_mode == 2 || (e->prefill_mode == 0 && n_tokens >= 8);
  if (use_tc) {
    if (int rc = prefill_tc(e, n_tokens, pos0, st)) return rc;
  } else {
    if (int rc = launch_mega(e, 1, n_tokens, 0, pos0, st)) return rc;
  }
  if (e->last) CK(cudaMemcpyAsync(logits_f32, e->logits, (size_t)e->cfg.vocab * 4, cudaMemcpyDeviceToHost, st));
  else CK(cudaMemcpyAsync(out_hidden_dev, e->stage_out, (size_t)n_tokens * D * 2, cudaMemcpyDeviceToDevice, st));
  CK(cudaStreamSynchronize(st));
  e->have_result = true;
  return 0;
}

// ---- the API hop: one engine call per shard per token like xotb200_forward, but the hidden state never visits the host.
//      decode (n_tokens == 1): the producing kernel's last epilogue stores it into the NEXT engine's inbox (peer store over
//      NVLink + release flag, the ring-mode hop); the call returns WITHOUT synchronising — the next engine's kernel acquires the
//      flag on the device, so it may be launched before this one has finished.
//      prompts (n_tokens > 1): one peer copy of the 16-bit [n_tokens, D] block into the next engine's receive buffer
//      (same-process links only; across processes use xotb200_forward_device + the caller's transport).
int xotb200_forward_hop(xotb200_engine* e, const int32_t* tokens, int64_t seq_in, int n_tokens, int pos0, int64_t* seq_out,
                        float* logits_f32) {
  REQUIRE(e && e->start >= 0, "no shard set");
  REQUIRE((tokens != nullptr) == e->first, "%s shard expects %s", e->first ? "first" : "non-first", e->first ? "token ids" : "a hop number");
  REQUIRE(e->first || seq_in >= 0, "seq_in must name the hop to consume");
  REQUIRE(n_tokens >= 1 && pos0 >= 0, "bad n_tokens/pos0");
  REQUIRE(e->L > 0, "begin_request was not called");
  REQUIRE(pos0 + n_tokens <= e->L, "KV cache overflow: pos0 %d + n %d > capacity %d", pos0, n_tokens, e->L);
  if (e->last) REQUIRE(logits_f32 != nullptr, "null output");
  else REQUIRE(seq_out != nullptr && e->next_inbox != nullptr, "next engine not linked");
  if (int rc = xotb200_check_loaded(e)) return rc;
  CK(cudaSetDevice(e->device));
  if (int rc = ensure_staging(e, n_tokens)) return rc;
  const size_t D = e->cfg.dim;
  cudaStream_t st = e->stream;
  {   // an error left behind by an earlier asynchronous hop (flag wait timed out) surfaces here
    StepCtl h;
    CK(cudaMemcpyAsync(&h, e->ctl, sizeof(h), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (h.err) {
      MEMSET_SYNC(e, &e->ctl->err, 0, 4);
      set_err("an earlier hop on this engine timed out waiting for the previous shard's flag");
      return XOTB200_ETIMEOUT;
    }
  }
  const bool decode = n_tokens == 1;
  if (tokens) {
    for (int i = 0; i < n_tokens; ++i)
      REQUIRE(tokens[i] >= 0 && tokens[i] < e->cfg.vocab, "token id %d out of range", tokens[i]);
    CK(cudaMemcpyAsync(e->stage_tokens, tokens, (size_t)n_tokens * 4, cudaMemcpyHostToDevice, st));
  } else if (!decode) {
    const int slot = (int)(seq_in % API_SLOTS);
    REQUIRE(e->hop_buf[slot] && e->hop_cap[slot] >= n_tokens, "hop %lld: no prompt block of %d tokens was pushed into this engine", (long long)seq_in, n_tokens);
    CK(cudaMemcpyAsync(e->stage_in, e->hop_buf[slot], (size_t)n_tokens * D * 2, cudaMemcpyDeviceToDevice, st));
  }
  unsigned int seq = 0;
  if (!e->last) seq = e->hop_seq_next++;
  e->hop_in = (!e->first && decode) ? 1 : 0;
  e->hop_out = (!e->last && decode) ? 1 : 0;
  e->hop_seq_in = (unsigned int)seq_in;
  e->hop_seq_out = seq;
  const bool use_tc = e->prefill_mode == 2 || (e->prefill_mode == 0 && n_tokens >= 8);
  int rc = use_tc ? prefill_tc(e, n_tokens, pos0, st) : launch_mega(e, 1, n_tokens, 0, pos0, st);
  e->hop_in = e->hop_out = 0;
  if (rc) return rc;
  if (e->last) {
    CK(cudaMemcpyAsync(logits_f32, e->logits, (size_t)e->cfg.vocab * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
  } else {
    if (!decode) {
      xotb200_engine* nx = e->next_engine;
      REQUIRE(nx != nullptr, "multi-token hops need a same-process link (xotb200_link_next_local); across processes use xotb200_forward_device");
      const int slot = (int)(seq % API_SLOTS);
      {
        std::lock_guard<std::mutex> lock(nx->hop_mu);
        if (nx->hop_cap[slot] < n_tokens) {
          CK(cudaSetDevice(nx->device));
          if (nx->hop_buf[slot]) { CK(cudaFree(nx->hop_buf[slot])); nx->hop_buf[slot] = nullptr; nx->hop_cap[slot] = 0; }
          CK(cudaMalloc(&nx->hop_buf[slot], (size_t)n_tokens * D * 2));
          nx->hop_cap[slot] = n_tokens;
          CK(cudaSetDevice(e->device));
        }
      }
      CK(cudaMemcpyPeerAsync(nx->hop_buf[slot], nx->device, e->stage_out, e->device, (size_t)n_tokens * D * 2, st));
      CK(cudaStreamSynchronize(st));
    }
    *seq_out = (int64_t)seq;
  }
  e->have_result = true;
  return 0;
}

int xotb200_last_token(xotb200_engine* e, int32_t* token) {
  REQUIRE(e && token, "null argument");
  REQUIRE(e->last && e->have_result, "no last-shard result available");
  CK(cudaSetDevice(e->device));
  StepCtl h;
  MEMCPY_SYNC(e, &h, e->ctl, sizeof(h), cudaMemcpyDeviceToHost);
  *token = h.token;
  return 0;
}

int xotb200_inbox_export(xotb200_engine* e, void* out) {
  REQUIRE(e && out, "null argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == XOTB200_IPC_HANDLE_BYTES, "ipc handle size");
  CK(cudaSetDevice(e->device));
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, e->inbox));
  memcpy(out, &h, sizeof(h));
  return 0;
}

static void drop_next(xotb200_engine* e) {
  if (e->next_ipc && e->next_inbox) cudaIpcCloseMemHandle(e->next_inbox);
  e->next_inbox = nullptr;
  e->next_ipc = false;
  e->next_engine = nullptr;
}
// A new link starts its API-hop numbering at 0 and clears the receiver's API flags (they hold seq + 1 of the previous link).
static int reset_hop_link(xotb200_engine* e) {
  e->hop_seq_next = 0;
  CK(cudaMemsetAsync(e->next_inbox->flag_api, 0, sizeof(e->next_inbox->flag_api), e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

int xotb200_link_next_ipc(xotb200_engine* e, const void* handle) {
  REQUIRE(e && handle, "null argument");
  CK(cudaSetDevice(e->device));
  drop_next(e);
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  e->next_inbox = static_cast<InboxHdr*>(p);
  e->next_ipc = true;
  return reset_hop_link(e);
}

int xotb200_link_next_local(xotb200_engine* e, xotb200_engine* next) {
  REQUIRE(e && next, "null argument");
  CK(cudaSetDevice(e->device));
  drop_next(e);
  if (next->device != e->device) {
    int can = 0;
    CK(cudaDeviceCanAccessPeer(&can, e->device, next->device));
    REQUIRE(can, "device %d cannot access peer %d", e->device, next->device);
    cudaError_t er = cudaDeviceEnablePeerAccess(next->device, 0);
    if (er != cudaSuccess && er != cudaErrorPeerAccessAlreadyEnabled) CK(er);
    cudaGetLastError();
  }
  e->next_inbox = next->inbox;
  e->next_engine = next;
  return reset_hop_link(e);
}

int xotb200_post_to_next(xotb200_engine* e, int step) {
  REQUIRE(e && e->next_inbox, "next inbox not linked");
  REQUIRE(e->have_result, "no forward result to post");
  CK(cudaSetDevice(e->device));
  if (e->last)
    post_kernel<<<1, 256, 0, e->stream>>>(e->next_inbox, nullptr, e->cfg.dim, &e->ctl->token, step, e->nreq + 1);
  else
    post_kernel<<<1, 256, 0, e->stream>>>(e->next_inbox, e->x, e->cfg.dim, nullptr, step, e->nreq + 1);
  CK(cudaGetLastError());
  e->launches++;
  e->ring_next_step = std::max(e->ring_next_step, step + 1);
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

int xotb200_ring_launch(xotb200_engine* e, int n_steps, int step0, int pos0) {
  REQUIRE(e && e->start >= 0 && e->next_inbox, "ring not linked");
  REQUIRE(n_steps >= 1 && n_steps <= MAX_RING_STEPS, "n_steps out of range (1..%d)", MAX_RING_STEPS);
  REQUIRE(step0 >= 0 && pos0 >= 0, "negative step0 / pos0");
  REQUIRE(n_steps % e->nreq == 0, "n_steps %d must be a multiple of the %d requests in flight", n_steps, e->nreq);
  REQUIRE(e->L > 0 && pos0 + n_steps / e->nreq <= e->L, "KV cache overflow: pos0 %d + %d steps per request > capacity %d", pos0,
          n_steps / e->nreq, e->L);
  if (int rc = xotb200_check_loaded(e)) return rc;
  CK(cudaSetDevice(e->device));
  cudaStream_t st = e->stream;
  CK(cudaStreamSynchronize(st));                 // the previous run's kernel must not still be writing the host token log
  *e->tok_count_host = 0;
  CK(cudaEventRecord(e->ev0, st));
  if (int rc = launch_mega(e, 0, n_steps, step0, pos0, st)) return rc;
  CK(cudaEventRecord(e->ev1, st));
  e->last_ring_steps = n_steps;
  e->ring_next_step = std::max(e->ring_next_step, step0 + n_steps);
  return 0;
}

// Continuous batching: a ring launch over an explicit set of ACTIVE requests, each with its own KV slot and its
// own position — requests join (after a prefill into a free slot) and leave between launches, the others keep their caches.
int xotb200_ring_launch_slots(xotb200_engine* e, int n_steps, int step0, const int32_t* kv_slots, const int32_t* positions, int n_active) {
  REQUIRE(e && kv_slots && positions, "null argument");
  REQUIRE(n_active >= 1 && n_active <= e->nreq, "n_active %d out of range (batch of %d KV caches)", n_active, e ? e->nreq : 0);
  REQUIRE(n_steps >= 1 && n_steps % n_active == 0, "n_steps %d must be a positive multiple of the %d active requests", n_steps, n_active);
  for (int i = 0; i < n_active; ++i) {
    REQUIRE(kv_slots[i] >= 0 && kv_slots[i] < e->nreq, "kv slot %d out of range", kv_slots[i]);
    REQUIRE(positions[i] >= 0 && positions[i] + n_steps / n_active <= e->L, "request in slot %d: position %d + %d tokens exceeds the KV capacity %d",
            kv_slots[i], positions[i], n_steps / n_active, e->L);
    for (int j = 0; j < i; ++j) REQUIRE(kv_slots[j] != kv_slots[i], "kv slot %d named twice", kv_slots[i]);
  }
  CK(cudaSetDevice(e->device));
  CK(cudaMemcpyAsync(e->d_req_map, kv_slots, (size_t)n_active * 4, cudaMemcpyHostToDevice, e->stream));
  CK(cudaMemcpyAsync(e->d_pos_tab, positions, (size_t)n_active * 4, cudaMemcpyHostToDevice, e->stream));
  e->n_active = n_active;
  const int saved_nreq = e->nreq;
  e->nreq = n_active;                       // the divisibility / capacity checks of xotb200_ring_launch apply to the active set
  const int rc = xotb200_ring_launch(e, n_steps, step0, 0);
  e->nreq = saved_nreq;
  e->n_active = 0;
  return rc;
}

// Tokens the last shard has produced so far in the CURRENT ring launch, readable while the persistent kernel is still running
// (the kernel writes them to pinned, mapped host memory): what a Node.on_token bridge streams from (node.py:119-147).
int xotb200_ring_poll(xotb200_engine* e, int32_t* tokens_out, int from, int max_n, int* n_done) {
  REQUIRE(e && n_done && from >= 0 && max_n >= 0, "bad argument");
  REQUIRE(e->last, "tokens are produced on the last shard");
  const unsigned int done = *e->tok_count_host;
  std::atomic_thread_fence(std::memory_order_acquire);
  const int avail = std::max(0, std::min((int)done - from, max_n));
  if (tokens_out)
    for (int i = 0; i < avail; ++i) tokens_out[i] = e->tok_host[from + i];
  *n_done = (int)done;
  return 0;
}

// Seed ring step `step` of the FIRST shard with an explicit token id (this engine is the last shard; `n_active` requests will be
// in flight in the run that consumes it).  Used when the set of active requests changes between launches.
int xotb200_post_token(xotb200_engine* e, int step, int32_t token, int n_active) {
  REQUIRE(e && e->next_inbox, "next inbox not linked");
  REQUIRE(e->last, "tokens are posted by the last shard");
  REQUIRE(step >= 0 && n_active >= 1 && n_active <= INBOX_SLOTS - 1 && token >= 0 && token < e->cfg.vocab, "bad argument");
  CK(cudaSetDevice(e->device));
  CK(cudaMemcpyAsync(e->d_pos_tab + INBOX_SLOTS - 1, &token, 4, cudaMemcpyHostToDevice, e->stream));   // scratch word
  post_kernel<<<1, 32, 0, e->stream>>>(e->next_inbox, nullptr, e->cfg.dim, e->d_pos_tab + INBOX_SLOTS - 1, step, n_active + 1);
  CK(cudaGetLastError());
  e->launches++;
  e->ring_next_step = std::max(e->ring_next_step, step + 1);
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

// Zero the KV cache of one request slot (what reset_caches does for the reference's single request, llm_utils.py:409-414).
int xotb200_clear_request(xotb200_engine* e, int request) {
  REQUIRE(e && e->kv && request >= 0 && request < e->nreq, "request index out of range");
  CK(cudaSetDevice(e->device));
  CK(cudaMemsetAsync(e->kv + (size_t)request * e->kv_req_stride, 0, e->kv_req_stride * 2, e->stream));
  CK(cudaStreamSynchronize(e->stream));
  return 0;
}

int xotb200_ring_wait(xotb200_engine* e, int32_t* tokens_out, int n_steps, float* ms_out) {
  REQUIRE(e, "null engine");
  CK(cudaSetDevice(e->device));
  CK(cudaStreamSynchronize(e->stream));
  StepCtl h;
  MEMCPY_SYNC(e, &h, e->ctl, sizeof(h), cudaMemcpyDeviceToHost);
  if (h.err) {
    MEMSET_SYNC(e, &e->ctl->err, 0, 4);
    set_err("ring step %d: timed out waiting for the previous shard's flag", h.step);
    return XOTB200_ETIMEOUT;
  }
  if (tokens_out) {
    REQUIRE(e->last, "tokens are produced on the last shard");
    REQUIRE(n_steps <= e->last_ring_steps, "n_steps exceeds the last ring run");
    MEMCPY_SYNC(e, tokens_out, e->tokens_log, (size_t)n_steps * 4, cudaMemcpyDeviceToHost);
  }
  if (ms_out) CK(cudaEventElapsedTime(ms_out, e->ev0, e->ev1));
  e->have_result = true;
  return 0;
}

int xotb200_ring_next_step(xotb200_engine* e, int* step) {
  REQUIRE(e && step, "null argument");
  *step = e->ring_next_step;
  return 0;
}

int xotb200_ring_timestamps(xotb200_engine* e, uint64_t* t_in, uint64_t* t_out, int n_steps) {
  REQUIRE(e && t_in && t_out && n_steps <= MAX_RING_STEPS, "bad argument");
  CK(cudaSetDevice(e->device));
  MEMCPY_SYNC(e, t_in, e->t_in, (size_t)n_steps * 8, cudaMemcpyDeviceToHost);
  MEMCPY_SYNC(e, t_out, e->t_out, (size_t)n_steps * 8, cudaMemcpyDeviceToHost);
  return 0;
}

int xotb200_set_cta_limit(xotb200_engine* e, int max_ctas) {
  REQUIRE(e && max_ctas >= 0, "bad argument");
  REQUIRE(max_ctas == 0 || max_ctas >= e->cfg.n_kv_heads, "cta limit must be 0 (off) or >= n_kv_heads");
  e->cta_limit = max_ctas;
  e->mega = nullptr;   // re-resolve grid / splits on the next launch
  return 0;
}

int xotb200_set_prefill_mode(xotb200_engine* e, int mode) {
  REQUIRE(e && mode >= 0 && mode <= 2, "prefill mode must be 0 (auto), 1 (decode kernel) or 2 (tensor cores)");
  e->prefill_mode = mode;
  return 0;
}

int xotb200_debug_gemm(xotb200_engine* e, const void* a_bf16, const void* w_bf16, int S, int N, int K, int tk, void* out_bf16) {
  REQUIRE(e && a_bf16 && w_bf16 && out_bf16, "null argument");
  REQUIRE(S >= 1 && S <= TC_M && K % tk == 0 && tk % TC_BK == 0 && N % tc_nt(N) == 0, "bad GEMM shape S=%d N=%d K=%d tk=%d", S, N, K, tk);
  CK(cudaSetDevice(e->device));
  bf16 *dA = nullptr, *dW0 = nullptr, *dW = nullptr, *dO = nullptr;
  struct Free4 {   // every exit path (CK returns early) releases the four scratch allocations
    bf16 **a, **b, **c, **d;
    ~Free4() { for (bf16** p : {a, b, c, d}) if (*p) cudaFree(*p); }
  } guard{&dA, &dW0, &dW, &dO};
  CK(cudaMalloc(&dA, (size_t)TC_M * K * 2));
  CK(cudaMalloc(&dW0, (size_t)N * K * 2));
  CK(cudaMalloc(&dW, (size_t)N * K * 2));
  CK(cudaMalloc(&dO, (size_t)TC_M * N * 2));
  MEMSET_SYNC(e, dA, 0, (size_t)TC_M * K * 2);
  MEMCPY_SYNC(e, dA, a_bf16, (size_t)S * K * 2, cudaMemcpyHostToDevice);
  MEMCPY_SYNC(e, dW0, w_bf16, (size_t)N * K * 2, cudaMemcpyHostToDevice);
  relayout_rows_kernel<<<N, 128, 0, e->stream>>>(dW0, dW, N, K, 0, e->cfg.head_dim, 1, 0, N, tk);
  CK(cudaGetLastError());
  CUtensorMap tA, tB;
  int rc = make_tmap_act(&tA, dA, K);
  if (!rc) rc = make_tmap_w(&tB, dW, N, K, tk, tc_nt(N));
  if (!rc) {
    TcGemmArgs g{S, N, K, tk, dO, nullptr, N, 0};
    rc = launch_tc_gemm(e, tA, tB, g, TCE_STORE, tc_cfg(N, S), e->stream);
  }
  cudaError_t er = cudaStreamSynchronize(e->stream);
  if (!rc && er == cudaSuccess) {
    er = cudaMemcpyAsync(out_bf16, dO, (size_t)S * N * 2, cudaMemcpyDeviceToHost, e->stream);
    if (er == cudaSuccess) er = cudaStreamSynchronize(e->stream);
  }
  if (rc) return rc;
  CK(er);
  return 0;
}

int xotb200_debug_routing(xotb200_engine* e, int32_t* out, int n_positions) {
  REQUIRE(e && out && e->start >= 0, "bad argument");
  REQUIRE(e->cfg.n_experts > 0 && e->route_log, "not a sparse-MoE shard with an open request");
  REQUIRE(n_positions >= 1 && n_positions <= e->L, "n_positions %d out of range (KV capacity %d)", n_positions, e->L);
  CK(cudaSetDevice(e->device));
  const size_t nl = e->end - e->start + 1;
  MEMCPY_SYNC(e, out, e->route_log, (size_t)n_positions * nl * 4, cudaMemcpyDeviceToHost);
  return 0;
}

int xotb200_debug_timeline(xotb200_engine* e, int step, int cta, uint64_t* out, int n_u64) {
  REQUIRE(e, "null engine");
  CK(cudaSetDevice(e->device));
  // [one CTA's timeline: dbg_cap words][every CTA: (4 * n_layers + 1) ops x 256 CTAs x 3 words]
  const size_t cap = dbg_cap(e) + (size_t)(4 * e->cfg.n_layers + 1) * 256 * 3;
  if (!e->dbg) { CK(cudaMalloc(&e->dbg, cap * 8)); MEMSET_SYNC(e, e->dbg, 0, cap * 8); }
  if (out) {
    REQUIRE(n_u64 >= 0 && (size_t)n_u64 <= cap, "n_u64 too large (max %zu)", cap);
    CK(cudaStreamSynchronize(e->stream));
    MEMCPY_SYNC(e, out, e->dbg, (size_t)n_u64 * 8, cudaMemcpyDeviceToHost);
  }
  e->dbg_step = step;
  e->dbg_cta = cta;
  return 0;
}

int xotb200_kernel_launches(xotb200_engine* e, int64_t* count) {
  REQUIRE(e && count, "null argument");
  *count = e->launches;
  return 0;
}

int xotb200_load_stats(xotb200_engine* e, int64_t* bytes, double* seconds, double* memcpy_seconds) {
  REQUIRE(e && bytes && seconds && memcpy_seconds, "null argument");
  *bytes = e->load_bytes; *seconds = e->load_seconds; *memcpy_seconds = e->load_memcpy_seconds;
  return 0;
}

int xotb200_weight_bytes(xotb200_engine* e, int64_t* bytes) {
  REQUIRE(e && bytes, "null argument");
  *bytes = (int64_t)e->blob_bytes;
  return 0;
}

int xotb200_step_bytes(xotb200_engine* e, int ctx, int64_t* bytes) {
  REQUIRE(e && bytes && e->start >= 0, "bad argument");
  const xotb200_model_cfg& c = e->cfg;
  const int64_t D = c.dim, F = c.ffn, Hq = c.n_heads, Hkv = c.n_kv_heads, hd = c.head_dim, V = c.vocab;
  const int64_t nl = e->end - e->start + 1;
  const int64_t mlp = c.n_experts > 0 ? (int64_t)c.n_experts * D + (int64_t)c.experts_per_tok * 3 * D * F : 3 * D * F;   // active experts only
  int64_t b = nl * 2 * (D * Hq * hd + 2 * D * Hkv * hd + Hq * hd * D + mlp + 2 * D);
  if (e->last) b += 2 * (V * D + D);
  if (e->first) b += 2 * D;
  b += 2 * 2 * Hkv * hd * nl * (int64_t)ctx + 2 * 2 * Hkv * hd * nl;
  *bytes = b;
  return 0;
}

}  // extern "C"
