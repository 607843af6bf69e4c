"""The lane / fragment mapping of the decode kernel's tensor-core GEMV (xotorch_b200/csrc/megakernel_impl.inc, `mma_gemv`),
emulated in numpy from the PTX fragment layout of mma.sync.m16n8k16 — the check that was run before the kernel first went to
the GPU, kept as a test so the comment in the source stays true.  No GPU, no library: pure index arithmetic.

  A (row-major, 16x16): a0: (g, 2t..2t+1)  a1: (g+8, 2t..)  a2: (g, 2t+8..)  a3: (g+8, 2t+8..)      g = lane >> 2, t = lane & 3
  B (col-major, 16x8) : b0: (k = 2t..2t+1, n = g)  b1: (k = 2t+8.., n = g)
  D (16x8)            : c0,c1: (g, 2t..2t+1)  c2,c3: (g+8, 2t..2t+1)
"""
import numpy as np
import pytest


def _emulate(RPW: int, W: np.ndarray, x: np.ndarray) -> np.ndarray:
  """one warp, one ring stage slice: W [RPW rows, 256 columns], x [256]; returns the RPW row totals as the kernel forms them"""
  LPR = 32 // RPW            # lanes per weight row inside one MMA block
  NWIN = LPR // 4            # 32-column windows per MMA block
  MBLK = 256 // RPW          # columns per MMA block
  # x in shared memory: the two middle 32-bit words of every 16-byte granule swapped (xs_perm)
  xs = x.reshape(-1, 4, 2)[:, [0, 2, 1, 3], :].reshape(-1)
  acc = [np.zeros((32, 4)), np.zeros((32, 4))]       # accE / accO per lane
  for m in range(RPW):
    A = np.zeros((16, 16))
    B = np.zeros((16, 8))
    for lane in range(32):
      g, t = lane >> 2, lane & 3
      # one 16-byte load: 8 weights of row lane // LPR at column 8 * (lane % LPR) of block m; register order = fragment order
      w = W[lane // LPR, m * MBLK + (lane % LPR) * 8: m * MBLK + (lane % LPR) * 8 + 8].reshape(4, 2)
      A[g, 2 * t:2 * t + 2], A[g + 8, 2 * t:2 * t + 2] = w[0], w[1]
      A[g, 2 * t + 8:2 * t + 10], A[g + 8, 2 * t + 8:2 * t + 10] = w[2], w[3]
      # one 8-byte load: window bn >> 1, half bn & 1
      bn = (lane >> 2) & (2 * NWIN - 1)
      xo = m * MBLK + (bn >> 1) * 32 + t * 8 + (bn & 1) * 4
      b = xs[xo:xo + 4].reshape(2, 2)
      B[2 * t:2 * t + 2, g], B[2 * t + 8:2 * t + 10, g] = b[0], b[1]
    D = A @ B
    for lane in range(32):
      g, t = lane >> 2, lane & 3
      acc[m & 1][lane] += [D[g, 2 * t], D[g, 2 * t + 1], D[g + 8, 2 * t], D[g + 8, 2 * t + 1]]
  v = (acc[0][:, 0] + acc[1][:, 0]) + (acc[0][:, 3] + acc[1][:, 3])       # c0 + c3 of every lane
  tot = np.array([sum(v[(lane & ~(LPR - 1)) + 5 * h] for h in range(NWIN)) for lane in range(32)])
  return tot[::LPR]           # lanes 0, LPR, 2*LPR, ... hold rows 0, 1, 2, ...


@pytest.mark.parametrize("RPW", [4, 2, 8])      # TK=2048 (4 rows per stage), tiny parity shapes (2), and an 8-row stage
def test_diagonal_of_the_mma_is_the_gemv(RPW):
  rng = np.random.default_rng(RPW)
  W = rng.standard_normal((RPW, 256))
  x = rng.standard_normal(256)
  got = _emulate(RPW, W, x)
  assert np.allclose(got, W @ x, rtol=0, atol=1e-12)


def test_rows_do_not_leak_into_each_other():
  """a stale (garbage) row of a short last stage must not touch the other rows: D[i][*] depends on A row i only"""
  rng = np.random.default_rng(0)
  W = rng.standard_normal((4, 256))
  x = rng.standard_normal(256)
  ref = _emulate(4, W, x)
  W2 = W.copy()
  W2[3] = 1e30                                      # garbage in the last row
  got = _emulate(4, W2, x)
  assert np.array_equal(got[:3], ref[:3])
