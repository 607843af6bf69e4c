// kernels.cuh — shared device helpers (PTX wrappers for mbarrier / TMA bulk copies / system-scope flags, bf16
// unpack + dot products) and the small utility kernels (weight re-layout, synthetic weights, dtype conversion,
// inbox post).  The hot path itself is the persistent kernel in megakernel.cuh.
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
constexpr int INBOX_SLOTS = 9;   // ring mode: up to 8 requests in flight through a ring + 1
constexpr int API_SLOTS = 8;     // API hop (xotb200_forward_hop, one engine call per shard per token): hops in flight per link
struct InboxHdr {
  unsigned int flag_hidden[INBOX_SLOTS];   // == step+1 when hidden[slot] holds the activation of ring step `step`
  unsigned int flag_token[INBOX_SLOTS];    // == step+1 when token[slot] holds the token id feeding ring step `step`
  int token[INBOX_SLOTS];
  unsigned int flag_api[API_SLOTS];        // == seq+1 when hidden[INBOX_SLOTS + seq % API_SLOTS] holds API hop `seq` of this link
  int pad[64 - 3 * INBOX_SLOTS - API_SLOTS];   // header is 256 bytes; bf16 hidden[INBOX_SLOTS + API_SLOTS][D] follows
};
static_assert(sizeof(InboxHdr) == 256, "InboxHdr must be 256 bytes");

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
// small utility kernels
// ------------------------------------------------------------------------------------------------------------
// post a hidden vector / token into the next inbox and release the flag (used once, after prefill)
__global__ void post_kernel(InboxHdr* next, const bf16* hidden, int D, const int* token, int step, int nslots) {
  const int slot = step % nslots;
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

// Row re-layout while loading: destination row rd = dst_row_off + r*dst_stride_rows takes source row map(r)
//   mode 0: identity; mode 1: q/k interleave (_permute, llm_utils.py:175-183) with `hd`.
// Destination layout: panel_tk == 0 -> plain row-major [n_total, K]; else K-panels [K/panel_tk][n_total][panel_tk]
// (element (rd, c) at ((c / tk) * n_total + rd) * tk + c % tk) so that any run of consecutive rows of one K chunk is
// contiguous: a whole ring stage of the persistent kernel is ONE TMA bulk copy.
__device__ __forceinline__ size_t panel_index(int rd, int c, int K, int n_total, int tk) {
  if (tk == 0) return (size_t)rd * K + c;
  return ((size_t)(c / tk) * n_total + rd) * tk + (c % tk);
}
__device__ __forceinline__ int src_row_of(int r, int mode, int hd) {
  if (mode != 1) return r;
  const int h = r / hd, i = r % hd;
  return h * hd + (i & 1) * (hd / 2) + (i >> 1);
}
__global__ void relayout_rows_kernel(const bf16* src, bf16* dst, int rows, int K, int mode, int hd, int dst_stride_rows,
                                     int dst_row_off, int n_total, int panel_tk) {
  const int r = blockIdx.x;
  if (r >= rows) return;
  const int sr = src_row_of(r, mode, hd);
  const int rd = dst_row_off + r * dst_stride_rows;
  const uint4* s4 = reinterpret_cast<const uint4*>(src + (size_t)sr * K);
  for (int c8 = threadIdx.x; c8 < K / 8; c8 += blockDim.x)
    *reinterpret_cast<uint4*>(dst + panel_index(rd, c8 * 8, K, n_total, panel_tk)) = s4[c8];
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
}  // namespace xb
