/*
 * xotb200.h — C ABI of libxotb200.so: the B200-native (sm_100a) shard-decode engine.
 *
 * This is the drop-in boundary for xotorch's ONE hot path: the per-shard Llama-style decode/prefill step behind
 * `InferenceEngine.infer_tensor` and the inter-shard hidden-state hop behind `PeerHandle.send_tensor`.
 * The reference is pure Python and has no FFI for this path; the closest precedent is the external "Cheetah"
 * C++ service it talks to over a Unix socket (/root/reference/xotorch/inference/cheetah/sharded_inference_engine.py:331-457).
 * Each entry point below cites the reference interface it replaces (paths relative to /root/reference).
 * The ctypes binding a maintainer would add is xotorch_b200/_lib.py (shown in INTEGRATION.md).
 *
 * Conventions: every function returns 0 on success or a negative XOTB200_E* code; the message is available
 * from xotb200_last_error() (thread-local).  All pointers are plain host pointers unless the name says `dev`.
 * Host buffers are caller-owned; device buffers are library-owned.  No torch types cross this boundary.
 * Calls on one engine must be serialised by the caller (the reference funnels all engine work through a
 * 1-thread executor: xotorch/inference/torch/sharded_inference_engine.py:46).
 */
#ifndef XOTB200_H
#define XOTB200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XOTB200_OK 0
#define XOTB200_EINVAL (-1)   /* bad argument / wrong state               */
#define XOTB200_ECUDA (-2)    /* a CUDA runtime call failed               */
#define XOTB200_EIO (-3)      /* safetensors / file problem               */
#define XOTB200_ESTRICT (-4)  /* strict loader: tensor missing or unknown */
#define XOTB200_ETIMEOUT (-5) /* device-side wait on a peer flag expired  */

#define XOTB200_DTYPE_BF16 0
#define XOTB200_DTYPE_F16 1 /* IEEE half weights / activations / KV (the 70B card upstream is fp16); fp32 accumulation as for bf16 */

#define XOTB200_ROPE_PLAIN 0  /* RotaryPositionalEmbeddings (general_mha.py:57-63, and llama ids w/o rope_scaling: App. B3) */
#define XOTB200_ROPE_LLAMA3 1 /* Llama3ScaledRoPE (general_mha.py:33-40)                                              */

/* Mirror of the dict built by load_model_config (xotorch/inference/llm_utils.py:79-126). */
typedef struct xotb200_model_cfg {
  int32_t n_layers;       /* num_layers       <- num_hidden_layers        */
  int32_t dim;            /* embed_dim        <- hidden_size              */
  int32_t ffn;            /* intermediate_dim <- intermediate_size        */
  int32_t n_heads;        /* num_heads                                     */
  int32_t n_kv_heads;     /* num_kv_heads                                  */
  int32_t head_dim;       /* head_dim (64 or 128)                          */
  int32_t vocab;          /* vocab_size                                    */
  int32_t max_pos;        /* max_seq_len      <- max_position_embeddings   */
  float rope_theta;       /* rope_base                                     */
  float norm_eps;         /* norm_eps                                      */
  int32_t tie_embeddings; /* tied lm_head (general_mha.py:42-44,125-126)  */
  int32_t dtype;          /* XOTB200_DTYPE_*                               */
  int32_t rope_kind;      /* XOTB200_ROPE_*                                */
  float rope_factor;      /* HF rope_scaling.factor (App. B2)              */
  float rope_low_freq_factor;
  float rope_high_freq_factor;
  int32_t rope_old_context_len;
  int32_t n_experts;        /* num_local_experts: Mixtral-style sparse MoE MLP, 2..8 (0 = dense SwiGLU MLP)   */
  int32_t experts_per_tok;  /* num_experts_per_tok: must be 2 when n_experts > 0                                */
} xotb200_model_cfg;

typedef struct xotb200_engine xotb200_engine;

/* Thread-local message of the last failing call on this thread. */
const char* xotb200_last_error(void);
/* Library / build identification ("xotb200 <version> sm_100a"). */
const char* xotb200_version(void);

/* ---- lifetime: replaces ShardedGeneralModel.__init__ / clear_model
 *      (xotorch/inference/torch/models/general_mha.py:144-176; sharded_inference_engine.py:85-106) ---- */
int xotb200_create(const xotb200_model_cfg* cfg, int device, xotb200_engine** out);
int xotb200_destroy(xotb200_engine* e);

/* ---- shard selection + weights: replaces TorchDynamicShardInferenceEngine.ensure_shard -> start_model ->
 *      load_model_weights_torchtune (sharded_inference_engine.py:372-421; llm_utils.py:185-333).
 *      Inclusive layer range as in Shard (xotorch/inference/shard.py:4-39).  Embedding is held only by the first
 *      shard, final norm + lm_head only by the last (App. B6).  The loader is STRICT (App. B6). ---- */
int xotb200_set_shard(xotb200_engine* e, int start_layer, int end_layer);
/* set_shard + read every *.safetensors in `dir` (HF names, HF layout; q/k re-laid out like _permute, llm_utils.py:175-183).
 * Streams: file mapping -> pinned double buffer (parallel copy) -> DMA -> re-layout kernel, chunk i+1 overlapping chunk i; strict:
 * missing tensors, unknown tensors of an owned layer, wrong dtype/shape and inconsistent data_offsets are errors. */
int xotb200_load_shard(xotb200_engine* e, const char* safetensors_dir, int start_layer, int end_layer);
/* one tensor by HF name from a host buffer (16-bit patterns in the engine's dtype); tensors outside the shard are ignored (returns 1) */
int xotb200_load_tensor(xotb200_engine* e, const char* hf_name, const void* host_data, const int64_t* shape, int ndim);
/* fill the current shard with the counter-based synthetic weights of xotorch_b200/synth.py, generated on device */
int xotb200_init_synthetic(xotb200_engine* e, uint64_t seed);
/* 0 if every tensor of the current shard has been loaded, else XOTB200_ESTRICT (message lists the first missing) */
int xotb200_check_loaded(xotb200_engine* e);

/* ---- request state: replaces encode()'s state/KV reset + setup_cache
 *      (sharded_inference_engine.py:71-83,133-147) and ShardTransformerDecoder.setup_caches/reset_caches
 *      (llm_utils.py:369-414).  KV capacity comes from the request (App. B7). ---- */
int xotb200_begin_request(xotb200_engine* e, int max_total_len);
/* Several independent requests in flight through the ring (each keeps its own KV cache; each is still a batch-1 stream):
 * ring step s serves request s % n_requests at position pos0 + s / n_requests, so with n_requests >= number of shards
 * every GPU of the ring is busy all the time (the reference keeps per-request state in Node but its engine
 * is single-request, sharded_inference_engine.py:250).  1 <= n_requests <= 8.  xotb200_forward / xotb200_post_to_next act on
 * the request chosen with xotb200_select_request (0 after begin_batch). */
int xotb200_begin_batch(xotb200_engine* e, int max_total_len, int n_requests);
int xotb200_select_request(xotb200_engine* e, int request);

/* ---- the step: replaces infer_tensor -> ShardedGeneralModel.generate -> ShardTransformerDecoder.forward
 *      (sharded_inference_engine.py:230-370; general_mha.py:178-254; llm_utils.py:416-489).
 *      Exactly one of tokens / hidden_f32 is non-NULL (ndim==2 ids vs ndim==3 hidden, sharded_inference_engine.py:254-263).
 *      n_tokens >= 1 positions starting at pos0 (pos0==0 => prefill).  Output, to caller-owned host memory:
 *        non-last shard: out_f32[n_tokens*dim]  = hidden state after the shard's last layer (bf16 values widened);
 *        last shard    : out_f32[vocab]         = logits of the LAST position (bf16 values widened), and the
 *                        lowest-index argmax is stored for xotb200_last_token(). ---- */
int xotb200_forward(xotb200_engine* e, const int32_t* tokens, const float* hidden_f32, int n_tokens, int pos0,
                    float* out_f32);
/* The same step with the hidden state left in DEVICE memory on both sides, 16-bit in the engine's dtype ([n_tokens, dim]):
 * the payload of PeerHandle.send_tensor (grpc_peer_handle.py:111-130) becomes 2*S*D bytes that the caller moves GPU to GPU
 * (NCCL send/recv, cudaMemcpyPeerAsync) — bit-identical to the fp32 wire because 16-bit -> f32 -> 16-bit is exact.
 * hidden_dev / out_hidden_dev are device pointers valid on the engine's device (the caller has made the input visible:
 * stream-synchronised); out_hidden_dev is written when the call returns.  Last shard: logits_f32 (host) as in xotb200_forward. */
int xotb200_forward_device(xotb200_engine* e, const int32_t* tokens, const void* hidden_dev, int n_tokens, int pos0,
                           void* out_hidden_dev, float* logits_f32);
/* The API hop — what replaces PeerHandle.send_tensor's payload (grpc_peer_handle.py:111-130; receiver grpc_server.py:73-88) when
 * both engines sit on one NVSwitch box: the same step as xotb200_forward, one call per shard per token, but the hidden state goes
 * from the producing kernel straight into the NEXT engine's HBM and only a hop NUMBER crosses the API (the descriptor the peer
 * handle carries).  First shard: `tokens` (host) != NULL; other shards: seq_in = the number the previous engine's call returned.
 * Non-last shards return the number of the hop they produced in *seq_out; the last shard writes the last position's logits to
 * logits_f32 (host) like xotb200_forward.
 *   n_tokens == 1 (decode): the last epilogue of the persistent kernel stores the 16-bit hidden state into the next engine's inbox
 *     (peer store over NVLink) and releases a flag; the consuming kernel acquires it on the device.  A non-last shard's call
 *     returns WITHOUT synchronising its stream, so the next engine's call may be issued at once: no NCCL, no host sync per hop.
 *   n_tokens  > 1 (prompt): one peer copy of the [n_tokens, dim] block into the next engine's receive buffer; needs a same-process
 *     link (xotb200_link_next_local).  Across processes move prompts with xotb200_forward_device + the caller's transport.
 * At most 8 hops may be in flight on one link (the reference keeps one per request). */
int xotb200_forward_hop(xotb200_engine* e, const int32_t* tokens, int64_t seq_in, int n_tokens, int pos0, int64_t* seq_out,
                        float* logits_f32);
/* lowest-index argmax of the most recent last-shard logits (on-device "greedy"; sharded_inference_engine.py:208-228 at temp 0) */
int xotb200_last_token(xotb200_engine* e, int32_t* token);

/* ---- device-resident ring (replaces Node.forward_tensor -> PeerHandle.send_tensor per token:
 *      xotorch/orchestration/node.py:424-443; xotorch/networking/grpc/grpc_peer_handle.py:111-130) ----
 * Every engine owns an inbox in its own HBM.  A shard's last kernel stores its bf16 hidden state (or, on the last
 * shard, the argmax token id) straight into the NEXT engine's inbox over NVLink and releases a flag; the next
 * shard's first kernel acquires the flag.  No host involvement per token. */
#define XOTB200_IPC_HANDLE_BYTES 64
int xotb200_inbox_export(xotb200_engine* e, void* ipc_handle_out /* XOTB200_IPC_HANDLE_BYTES */);
int xotb200_link_next_ipc(xotb200_engine* e, const void* ipc_handle /* of the next engine's inbox */);
int xotb200_link_next_local(xotb200_engine* e, xotb200_engine* next); /* same process (may be e itself) */
/* Hand the result of the most recent xotb200_forward (hidden state of its last position, or last token) to the
 * next engine's inbox as step `step` — used once after prefill to seed the ring. */
int xotb200_post_to_next(xotb200_engine* e, int step);
/* Launch the persistent kernel for n_steps ring steps starting at ring step `step0`, position `pos0`; asynchronous.
 * With a batch of B requests n_steps must be a multiple of B (n_steps / B tokens per request) and step0 a multiple of B. */
int xotb200_ring_launch(xotb200_engine* e, int n_steps, int step0, int pos0);
/* Continuous batching: the same launch over an explicit set of n_active requests of the batch (xotb200_begin_batch), request i
 * living in KV cache kv_slots[i] and standing at position positions[i]; ring step s serves active request s % n_active.  Requests
 * join (prefill into a free slot with xotb200_select_request + xotb200_forward, then xotb200_post_token) and leave between
 * launches — the per-slot position table the reference's Node keeps per request (node.py:119-147) moved next to the kernel. */
int xotb200_ring_launch_slots(xotb200_engine* e, int n_steps, int step0, const int32_t* kv_slots, const int32_t* positions, int n_active);
/* Last shard: seed ring step `step` of the first shard with an explicit token id (n_active = requests in flight in that run). */
int xotb200_post_token(xotb200_engine* e, int step, int32_t token, int n_active);
/* Last shard: tokens produced so far by the CURRENT ring launch, readable while its persistent kernel is still running (pinned
 * mapped host memory written by the kernel).  Copies up to max_n ids starting at index `from`; *n_done = tokens produced so far.
 * This is what a Node.on_token bridge streams from (the reference fires on_token once per token, node.py:127-137). */
int xotb200_ring_poll(xotb200_engine* e, int32_t* tokens_out, int from, int max_n, int* n_done);
/* Zero the KV cache of one request slot of the batch (reset_caches for ONE request, llm_utils.py:409-414). */
int xotb200_clear_request(xotb200_engine* e, int request);
/* Wait for the engine's stream; on the last shard copies the n_steps generated token ids (int32) to tokens_out
 * (may be NULL).  ms_out (may be NULL) = device time of the launched steps by CUDA events. */
int xotb200_ring_wait(xotb200_engine* e, int32_t* tokens_out, int n_steps, float* ms_out);
/* First ring step number that no xotb200_post_to_next / xotb200_ring_launch on this engine has used yet.  Inbox flags hold
 * `step + 1` and are never cleared, so step numbers must never be reused on an engine: a driver that is (re)created over
 * existing engines, or that recovers from XOTB200_ETIMEOUT, continues from the maximum of this value over its engines. */
int xotb200_ring_next_step(xotb200_engine* e, int* step);
/* Per-step device timestamps (globaltimer ns) of the most recent ring run: t_in = input flag observed,
 * t_out = result posted.  Arrays of n_steps uint64 each (hop latency accounting). */
int xotb200_ring_timestamps(xotb200_engine* e, uint64_t* t_in, uint64_t* t_out, int n_steps);

/* ---- tuning / debugging ----
 * Limit the persistent kernel to max_ctas CTAs (0 = one per SM).  Needed when several engines share ONE GPU and run
 * ring mode concurrently: their persistent kernels must be co-resident (sum of limits <= SM count). */
int xotb200_set_cta_limit(xotb200_engine* e, int max_ctas);
/* Prefill path: 0 = automatic (>= 8 positions per call run as tensor-core GEMMs: tcgen05 + TMEM + TMA tensor maps,
 * prefill_tc.cuh; fewer run through the decode kernel), 1 = always the decode kernel, 2 = always tensor cores. */
int xotb200_set_prefill_mode(xotb200_engine* e, int mode);
/* Unit-test hook for the tensor-core GEMM: out[S,N] (bf16 bits) = A[S,K] * W[N,K]^T with fp32 accumulation, W re-laid out
 * as K-panels of width tk exactly like a model weight.  S <= 128. */
int xotb200_debug_gemm(xotb200_engine* e, const void* a_bf16, const void* w_bf16, int S, int N, int K, int tk, void* out_bf16);
/* Sparse-MoE shards: the experts the router chose for request 0 at positions [0, n_positions) of the open request,
 * out[pos * n_local_layers + local_layer] = lower expert | higher expert << 8 (-1 = position not run).  Parity tests use it
 * to tell a routing near-tie (a discrete, legitimately implementation-dependent choice) from an arithmetic error. */
int xotb200_debug_routing(xotb200_engine* e, int32_t* out, int n_positions);
/* Per-op device timeline of CTA `cta` at ring step `step` (-1 disables): for op i the four globaltimer stamps
 * [prologue start, tiles start, tiles done, barrier exit] at out[4*i..4*i+3].  Call once to arm (out may be NULL),
 * run the ring, call again to read. */
int xotb200_debug_timeline(xotb200_engine* e, int step, int cta, uint64_t* out, int n_u64);

/* ---- introspection ---- */
int xotb200_kernel_launches(xotb200_engine* e, int64_t* count); /* kernels launched by this engine so far */
/* the most recent xotb200_load_shard: bytes moved into HBM, wall seconds of the whole load, seconds of it spent copying out of
 * the file mapping into the pinned staging buffers (the rest overlaps: DMA + re-layout kernels) */
int xotb200_load_stats(xotb200_engine* e, int64_t* bytes, double* seconds, double* memcpy_seconds);
int xotb200_weight_bytes(xotb200_engine* e, int64_t* bytes);    /* resident weight bytes of the current shard */
int xotb200_step_bytes(xotb200_engine* e, int ctx, int64_t* bytes); /* algorithmic HBM bytes of one decode step at context ctx */

#ifdef __cplusplus
}
#endif
#endif /* XOTB200_H */
