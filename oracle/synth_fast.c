/* TEST INFRASTRUCTURE (oracle/): multi-threaded C twin of xotorch_b200/synth.py::synth_values_f32 so the CPU
 * baseline can materialise Llama-3-8B-shaped bf16 weights in seconds.  Bit-identical to the torch generator and to
 * the device generator (csrc/kernels.cuh::synth_value); checked in tests/test_oracle.py.
 * Build: gcc -O3 -fopenmp -shared -fPIC -o oracle/_build/libsynthfast.so oracle/synth_fast.c */
#include <omp.h>
#include <stdint.h>
#include <string.h>

/* The GPU box exposes every host core but runs the container under a CPU quota: size the team from the caller. */
void synth_set_threads(int n) { if (n > 0) omp_set_num_threads(n); }

static inline uint16_t f32_to_bf16_rne(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40); /* NaN */
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

/* out[i] = bf16( f32(s_i - 131070) * scale + mean ),  i in [start, start+count) of tensor `base` */
void synth_fill_bf16(uint16_t* out, uint64_t base, uint64_t start, uint64_t count, float scale, float mean) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)count; ++i) {
    uint64_t z = base + start + (uint64_t)i;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    int32_t s = (int32_t)(z & 0xffff) + (int32_t)((z >> 16) & 0xffff) + (int32_t)((z >> 32) & 0xffff) + (int32_t)(z >> 48);
    volatile float prod = (float)(s - 131070) * scale; /* no FMA contraction: mul then add, like torch */
    out[i] = f32_to_bf16_rne(prod + mean);
  }
}

/* IEEE half variant (the 70B card upstream is fp16): same counter stream, fp32 value rounded to nearest-even half.
 * gcc >= 12 on x86-64 converts float -> _Float16 with round-to-nearest-even (checked against torch in tests/test_oracle.py). */
void synth_fill_f16(uint16_t* out, uint64_t base, uint64_t start, uint64_t count, float scale, float mean) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < (int64_t)count; ++i) {
    uint64_t z = base + start + (uint64_t)i;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z = z ^ (z >> 31);
    int32_t s = (int32_t)(z & 0xffff) + (int32_t)((z >> 16) & 0xffff) + (int32_t)((z >> 32) & 0xffff) + (int32_t)(z >> 48);
    volatile float prod = (float)(s - 131070) * scale;
    const _Float16 h = (_Float16)(prod + mean);
    memcpy(&out[i], &h, 2);
  }
}
