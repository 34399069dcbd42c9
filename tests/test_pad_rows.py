"""The GPU path never computes the zero-scoring pad rows that the SSE2 layout appends to the read (ssw.c:108, 346); it rebuilds
their effect on the per-column maximum from the bottom-row H values (sw_finish.cuh).  This test pins that formula against the
scalar oracle, which does compute the pad rows cell by cell."""
import ctypes as ct

import numpy as np

from oracle import oracle


class Ends(ct.Structure):
    _fields_ = [(n, ct.c_int32) for n in ("score", "ref", "read", "score2", "ref2")]


def oracle_colmax(read, ref, mat, gapO, gapE, byte_mode, bias):
    P = oracle.port_lib()
    cm = np.zeros(len(ref), dtype=np.int32)
    e = Ends()
    P.oracle_score_pass(read.ctypes.data_as(ct.c_void_p), len(read), ref.ctypes.data_as(ct.c_void_p), len(ref), mat.ctypes.data_as(ct.c_void_p), 5,
                        gapO, gapE, byte_mode, bias, 0, -1, 15, ct.byref(e), cm.ctypes.data_as(ct.c_void_p))
    return cm, e


def real_rows_dp(read, ref, mat, gapO, gapE):
    R = len(read)
    H = np.zeros(R + 1, dtype=np.int64); E = np.zeros(R + 1, dtype=np.int64)
    cm = np.zeros(len(ref), dtype=np.int64); B = np.zeros(len(ref), dtype=np.int64)
    for i, t in enumerate(ref):
        Hn = np.zeros(R + 1, dtype=np.int64); F = 0
        for r in range(R):
            h = (H[r - 1] if r > 0 else 0) + mat[t * 5 + read[r]]
            v = max(0, h, E[r], F); Hn[r] = v
            E[r] = max(0, E[r] - gapE, v - gapO); F = max(0, F - gapE, v - gapO)
        H = Hn; cm[i] = Hn[:R].max(); B[i] = Hn[R - 1]
    return cm, B


def padded(cm, B, P, gapO, gapE):
    out = cm.copy()
    for c in range(len(cm)):
        best = out[c]
        for d in range(1, min(P, c) + 1):
            best = max(best, B[c - d])
        for d in range(P + 1, c + 1):
            best = max(best, B[c - d] - gapO - (d - P - 1) * gapE)
        out[c] = best
    return out


def test_pad_row_formula_matches_oracle():
    rng = np.random.default_rng(5)
    checked = 0
    for _ in range(150):
        alpha = int(rng.integers(2, 5)); R = int(rng.integers(1, 50)); L = int(rng.integers(1, 70))
        match = int(rng.integers(1, 4)); mism = int(rng.integers(1, 5)); gapE = int(rng.integers(1, 3)); gapO = gapE + int(rng.integers(1, 6))
        mat = np.full((5, 5), -mism, dtype=np.int8)
        for i in range(4):
            mat[i, i] = match
        mat = mat.reshape(-1)
        ref = rng.integers(0, alpha, size=L).astype(np.int8)
        st = int(rng.integers(0, L))
        read = np.resize(ref[st:], R).copy().astype(np.int8)
        mut = rng.random(R) < 0.15
        read[mut] = rng.integers(0, alpha, size=int(mut.sum()))
        cm, B = real_rows_dp(read, ref, mat, gapO, gapE)
        for byte_mode, W in ((1, 16), (0, 8)):
            ocm, e = oracle_colmax(read, ref, mat, gapO, gapE, byte_mode, mism)
            if byte_mode and e.score == 255:
                continue
            P = (W - R % W) % W
            assert np.array_equal(padded(cm, B, P, gapO, gapE), ocm.astype(np.int64)), (R, L, byte_mode)
            checked += 1
    assert checked > 200
