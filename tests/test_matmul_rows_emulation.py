"""CPU emulation of the thread / tile index arithmetic of the experimental skinny matmul (tensorfrost_b200/csrc/matmul_rows.cu).

The kernel was written after the round's GPU budget ended, so its index math is checked here by replaying it in numpy, one python
"thread" at a time with the same formulas (B staging with zero padding, A chunks of 32 k through a 36-float pitch, TM x 4*CH register
tiles, column chunks c*(BN/CH) + tx*4, guarded stores): every output element must be written exactly once with the right dot product.
This pins the design, not the compiled code; the GPU test (tests/test_library_gpu.py::test_matmul_rows) is enabled with TFCUDA_EXPERIMENTAL=1."""
import numpy as np
import pytest

THREADS, KC, APITCH = 256, 32, 36


def emulate(a, b, txn, ch, tm, grid=3):
    r, k = a.shape
    n = b.shape[1]
    bn, ty_count = txn * 4 * ch, THREADS // txn
    br = ty_count * tm
    kpad = (k + 3) & ~3
    bs = np.zeros((kpad, bn), np.float32)
    for e in range(kpad * bn):  # the staging loop (all threads together)
        kk, nn = divmod(e, bn)
        bs[kk, nn] = b[kk, nn] if (kk < k and nn < n) else 0.0
    c = np.full((r, n), np.nan, np.float32)
    writes = np.zeros((r, n), np.int32)
    tiles = (r + br - 1) // br
    vecs = br * (KC // 4)
    chunks = (k + KC - 1) // KC
    for block in range(grid):
        for tile in range(block, tiles, grid):
            r0 = tile * br
            acc = np.zeros((THREADS, tm, 4 * ch), np.float32)
            for chunk in range(chunks):
                k0 = chunk * KC
                a_s = np.zeros((br, APITCH), np.float32)
                for v in range(vecs):  # fetch + stash of every thread / slot
                    rr, kq = v // (KC // 4), (v % (KC // 4)) * 4
                    gr, gk = r0 + rr, k0 + kq
                    val = np.zeros(4, np.float32)
                    if gr < r and gk < k:
                        for j in range(4):
                            if gk + j < k:
                                val[j] = a[gr, gk + j]
                    a_s[rr, kq:kq + 4] = val
                klen = min(KC, kpad - k0)
                for tid in range(THREADS):
                    tx, ty = tid % txn, tid // txn
                    for kq in range(0, klen, 4):
                        for i in range(tm):
                            a4 = a_s[ty * tm + i, kq:kq + 4]
                            for kk in range(4):
                                for cc in range(ch):
                                    col = cc * (bn // ch) + tx * 4
                                    acc[tid, i, cc * 4:cc * 4 + 4] += a4[kk] * bs[k0 + kq + kk, col:col + 4]
            for tid in range(THREADS):
                tx, ty = tid % txn, tid // txn
                for i in range(tm):
                    gr = r0 + ty * tm + i
                    if gr >= r:
                        continue
                    for cc in range(ch):
                        col = cc * (bn // ch) + tx * 4
                        for j in range(4):
                            if col + j < n:
                                c[gr, col + j] = acc[tid, i, cc * 4 + j]
                                writes[gr, col + j] += 1
    return c, writes


@pytest.mark.parametrize("r,k,n,cfg", [(70, 48, 128, (16, 2, 4)), (130, 128, 12, (4, 1, 4)), (300, 12, 128, (16, 2, 4)), (65, 128, 48, (16, 1, 4)),
                                       (257, 7, 5, (4, 1, 4)), (129, 36, 30, (8, 1, 4)), (3, 4, 128, (16, 2, 4))])
def test_index_arithmetic_covers_every_output_once(r, k, n, cfg):
    rng = np.random.default_rng(r + k + n)
    a, b = rng.standard_normal((r, k)).astype(np.float32), rng.standard_normal((k, n)).astype(np.float32)
    c, writes = emulate(a, b, *cfg)
    assert (writes == 1).all()
    want = a.astype(np.float64) @ b.astype(np.float64)
    assert np.abs(c - want).max() <= 1e-4 * np.abs(want).max()
