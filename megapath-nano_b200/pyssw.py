"""Host-side mirror of MegaPath-Nano's bin/realignment/pyssw.py on top of the B200 engine.

Same public surface as the reference binding (pyssw.py:51-147): `SSW(match, mismatch, gap_open, gap_extend, lib_path)`,
`set_reference_sequence(ref)`, `align(query) -> (score, cigar_string, ref_begin)`, `get_cigar(...)`, `to_int(seq)`, so
FastPassAligner (fast_align_reads2ref.py:19-49) and local_realignment.py:321-327 can use it unchanged.  Differences:
  * `align()` goes through the very same C symbols (ssw_init / ssw_align / align_destroy) but they execute on the GPU;
  * `align_batch(queries)` is new: all queries against the current reference in ONE submit through the batched C ABI
    (include/mpn_ssw_batch.h) -- the call the ONT realignment loop should make instead of one ctypes round trip per read;
  * the profile is released with init_destroy (the reference leaks it, pyssw.py:141-147).
There is no CPU path: a missing library or GPU raises."""
import ctypes as ct
import os
import numpy as np

from . import batch as _batch

_HERE = os.path.dirname(os.path.abspath(__file__))
lib_path = os.path.join(_HERE, "realign", "libssw.so")

_CODES = np.full(256, 4, dtype=np.int8)            # A C G T -> 0..3, anything else -> N (4), case-insensitive (pyssw.py:61-98)
for _i, _c in enumerate("ACGT"):
    _CODES[ord(_c)] = _i
    _CODES[ord(_c.lower())] = _i
_CODES[ord("N")] = 4
_CODES[ord("n")] = 4


class CAlignRes(ct.Structure):                       # s_align, include/ssw.h
    _fields_ = [("nScore", ct.c_uint16), ("nScore2", ct.c_uint16), ("nRefBeg", ct.c_int32), ("nRefEnd", ct.c_int32),
                ("nQryBeg", ct.c_int32), ("nQryEnd", ct.c_int32), ("nRefEnd2", ct.c_int32),
                ("sCigar", ct.POINTER(ct.c_uint32)), ("nCigarLen", ct.c_int32)]


class SSW(object):
    _OPS = "MIDNSHP=X"

    def __init__(self, match=4, mismatch=6, gap_open=8, gap_extend=2, lib_path=lib_path):
        if not os.path.exists(lib_path):
            raise RuntimeError(f"{lib_path} not built; this aligner has no CPU fallback")
        self.match, self.mismatch, self.gap_open, self.gap_extend = match, mismatch, gap_open, gap_extend
        self.lib_path = lib_path
        self.lEle = ["A", "C", "G", "T", "N"]
        self.mat_np = self._matrix()
        self.mat = (ct.c_int8 * 25)(*[int(v) for v in self.mat_np])
        L = ct.cdll.LoadLibrary(lib_path)
        L.ssw_init.argtypes = [ct.POINTER(ct.c_int8), ct.c_int32, ct.POINTER(ct.c_int8), ct.c_int32, ct.c_int8]
        L.ssw_init.restype = ct.c_void_p
        L.init_destroy.argtypes = [ct.c_void_p]
        L.ssw_align.argtypes = [ct.c_void_p, ct.POINTER(ct.c_int8), ct.c_int32, ct.c_uint8, ct.c_uint8, ct.c_uint8, ct.c_uint16, ct.c_int32, ct.c_int32]
        L.ssw_align.restype = ct.POINTER(CAlignRes)
        L.align_destroy.argtypes = [ct.POINTER(CAlignRes)]
        self._lib = L
        self._engine = None
        self.reference = None

    def _matrix(self):
        """+match on the ACGT diagonal, -mismatch elsewhere among ACGT, 0 against N (pyssw.py:61-79)."""
        m = np.zeros((5, 5), dtype=np.int8)
        m[:4, :4] = -self.mismatch
        for i in range(4):
            m[i, i] = self.match
        return m.reshape(-1)

    def to_int(self, seq):
        if isinstance(seq, str):
            seq = seq.encode("latin-1", "replace")
        return _CODES[np.frombuffer(seq, dtype=np.uint8)].copy()

    def set_reference_sequence(self, reference):
        self.reference = reference
        self.rNum = self.to_int(reference)
        self.reference_len = len(reference)

    def get_cigar(self, cigar, ref_position_start, ref_position_end, query_position_start, query_position_end, query):
        out = []
        if query_position_start > 0:
            out.append(f"{query_position_start}S")
        for x in cigar:
            op = int(x) & 15
            c = "M" if op > 8 else self._OPS[op]
            out.append(f"{int(x) >> 4}{'=' if c == 'M' else c}")
        aligned = query_position_end - query_position_start + 1
        if aligned < len(query):
            out.append(f"{len(query) - aligned}S")
        return "".join(out)

    @staticmethod
    def _mask_len(qlen):
        return 15 if qlen <= 30 else qlen          # pyssw.py:142

    def align(self, query):
        """One query against the reference set by set_reference_sequence: (score, cigar string, ref_begin)."""
        q = self.to_int(query)
        qp = q.ctypes.data_as(ct.POINTER(ct.c_int8))
        prof = self._lib.ssw_init(qp, len(q), self.mat, 5, 2)
        try:
            res = self._lib.ssw_align(prof, self.rNum.ctypes.data_as(ct.POINTER(ct.c_int8)), self.reference_len, self.gap_open, self.gap_extend,
                                      2, 0, 0, self._mask_len(len(q)))
            r = res.contents
            cig = [r.sCigar[k] for k in range(r.nCigarLen)]
            out = (r.nScore, self.get_cigar(cig, r.nRefBeg, r.nRefEnd, r.nQryBeg, r.nQryEnd, query), r.nRefBeg)
            self._lib.align_destroy(res)
        finally:
            self._lib.init_destroy(prof)
        return out

    def align_batch(self, queries):
        """All queries against the current reference in one GPU submit; returns a list of (score, cigar string, ref_begin)."""
        if self._engine is None:
            self._engine = _batch.Engine()
        qs = [self.to_int(q) for q in queries]
        n = len(qs)
        if n == 0:
            return []
        rl = np.array([len(q) for q in qs], dtype=np.int64)
        # one arena: the reference once, then the queries; every pair points at the same reference span
        seq = np.concatenate([self.rNum] + qs)
        rd_start = self.reference_len + np.concatenate([[0], np.cumsum(rl)[:-1]]).astype(np.int64)
        b = type("B", (), {})()
        b.mat, b.n, b.gapO, b.gapE, b.score_size, b.flag, b.filters, b.filterd = self.mat_np, 5, self.gap_open, self.gap_extend, 2, 2, 0, 0
        masklen = np.array([self._mask_len(int(l)) for l in rl], dtype=np.int32)
        rec, cig = self._engine.align_spans(b, seq, rd_start, rl.astype(np.int32), np.zeros(n, np.int64), np.full(n, self.reference_len, np.int32), masklen)
        out = []
        for i in range(n):
            o, l = int(rec["cigar_off"][i]), int(rec["cigar_len"][i])
            out.append((int(rec["score1"][i]),
                        self.get_cigar(cig[o:o + l], int(rec["ref_begin1"][i]), int(rec["ref_end1"][i]), int(rec["read_begin1"][i]), int(rec["read_end1"][i]), queries[i]),
                        int(rec["ref_begin1"][i])))
        return out
