"""ctypes binding of the batched C ABI (include/mpn_ssw_batch.h).  Host-side plumbing only: every DP cell is computed by
the CUDA kernels in csrc/; if libmpn_ssw.so is missing or no GPU is present this module raises -- there is no CPU path."""
import ctypes as ct
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPN_SSW_LIB") or os.path.join(HERE, "libmpn_ssw.so")   # override: A/B runs of kernel variants

FIELDS = ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigarLen")


class MpnParams(ct.Structure):
    _fields_ = [("mat", ct.POINTER(ct.c_int8)), ("n", ct.c_int32), ("gapO", ct.c_int32), ("gapE", ct.c_int32), ("score_size", ct.c_int32),
                ("flag", ct.c_int32), ("filters", ct.c_int32), ("filterd", ct.c_int32)]


RESULT_DTYPE = np.dtype([("score1", np.uint16), ("score2", np.uint16), ("ref_begin1", np.int32), ("ref_end1", np.int32),
                         ("read_begin1", np.int32), ("read_end1", np.int32), ("ref_end2", np.int32), ("cigar_len", np.int32),
                         ("status", np.int32), ("cigar_off", np.int64)], align=True)

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(f"{LIB_PATH} is not built (run `python -c 'import __graft_entry__ as g; g.build()'`); there is no CPU fallback")
        L = ct.CDLL(LIB_PATH)
        L.mpn_engine_create.restype = ct.c_void_p
        L.mpn_engine_create.argtypes = [ct.c_int]
        L.mpn_engine_destroy.argtypes = [ct.c_void_p]
        L.mpn_engine_set_stream.argtypes = [ct.c_void_p, ct.c_void_p]
        L.mpn_engine_stats.argtypes = [ct.c_void_p] + [ct.POINTER(ct.c_int64)] * 4
        L.mpn_engine_set_profile.argtypes = [ct.c_void_p, ct.c_int]
        L.mpn_engine_phase_ms.argtypes = [ct.c_void_p, ct.POINTER(ct.c_float)]
        L.mpn_engine_phase_ms_mean.argtypes = [ct.c_void_p, ct.POINTER(ct.c_float), ct.POINTER(ct.c_int)]
        L.mpn_batch_upload.restype = ct.c_void_p
        L.mpn_batch_upload.argtypes = [ct.c_void_p, ct.POINTER(MpnParams), ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64]
        L.mpn_batch_run.argtypes = [ct.c_void_p]
        L.mpn_batch_fetch.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64]
        L.mpn_batch_free.argtypes = [ct.c_void_p]
        L.mpn_align_batch.argtypes = [ct.c_void_p, ct.POINTER(MpnParams), ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64,
                                      ct.c_void_p, ct.c_void_p, ct.c_int64]
        L.mpn_align_batch_spans.argtypes = [ct.c_void_p, ct.POINTER(MpnParams), ct.c_void_p, ct.c_int64, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p,
                                            ct.c_void_p, ct.c_int64, ct.c_void_p, ct.c_void_p, ct.c_int64]
        L.mpn_align_batch_packed4.argtypes = [ct.c_void_p, ct.POINTER(MpnParams), ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64,
                                              ct.c_void_p, ct.c_void_p, ct.c_int64]
        L.mpn_pack4.argtypes = [ct.c_void_p, ct.c_int64, ct.c_void_p]
        L.mpn_pack4.restype = None
        L.mpn_pool_create.restype = ct.c_void_p
        L.mpn_pool_create.argtypes = [ct.c_void_p, ct.c_int]
        L.mpn_pool_destroy.argtypes = [ct.c_void_p]
        L.mpn_pool_ndev.argtypes = [ct.c_void_p]
        L.mpn_pool_engine.restype = ct.c_void_p
        L.mpn_pool_engine.argtypes = [ct.c_void_p, ct.c_int]
        L.mpn_pool_align_batch.argtypes = [ct.c_void_p, ct.POINTER(MpnParams), ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64,
                                           ct.c_void_p, ct.c_void_p, ct.c_int64]
        L.mpn_pool_align_batch_spans.argtypes = [ct.c_void_p, ct.POINTER(MpnParams), ct.c_void_p, ct.c_int64, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p,
                                                 ct.c_void_p, ct.c_int64, ct.c_void_p, ct.c_void_p, ct.c_int64]
        L.mpn_pool_last_shares.argtypes = [ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_void_p]
        assert ct.sizeof(MpnParams) == 40 and RESULT_DTYPE.itemsize == 40, (ct.sizeof(MpnParams), RESULT_DTYPE.itemsize)
        _lib = L
    return _lib


def _ptr(a):
    """address of a numpy array or a torch tensor (pinned host memory)"""
    if hasattr(a, "data_ptr"):
        return ct.c_void_p(a.data_ptr())
    return ct.c_void_p(a.ctypes.data)


class Engine:
    """One GPU, one stream.  `align(batch)` is the call a user makes: host buffers in, host records out."""

    def __init__(self, device=-1):
        self.L = lib()
        self.h = self.L.mpn_engine_create(device)
        if not self.h:
            raise RuntimeError("mpn_engine_create failed: no CUDA device (this engine has no CPU fallback)")

    def close(self):
        if self.h:
            self.L.mpn_engine_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_stream(self, cuda_stream_ptr):
        self.L.mpn_engine_set_stream(self.h, ct.c_void_p(cuda_stream_ptr))

    def set_profile(self, on=True):
        self.L.mpn_engine_set_profile(self.h, 1 if on else 0)

    def phase_ms(self):
        """ms of {forward, finish, reverse, trace} of the last run (needs set_profile(True))"""
        v = (ct.c_float * 4)()
        if self.L.mpn_engine_phase_ms(self.h, v):
            return None
        return dict(forward=v[0], finish=v[1], reverse=v[2], trace=v[3])

    def phase_ms_mean(self):
        """the same averaged over the runs since set_profile(True) (at most the last 16): (dict, number of runs)"""
        v = (ct.c_float * 4)()
        n = ct.c_int(0)
        if self.L.mpn_engine_phase_ms_mean(self.h, v, ct.byref(n)):
            return None, 0
        return dict(forward=v[0], finish=v[1], reverse=v[2], trace=v[3]), n.value

    def stats(self):
        v = [ct.c_int64(0) for _ in range(4)]
        self.L.mpn_engine_stats(self.h, *[ct.byref(x) for x in v])
        return dict(launches=v[0].value, pairs=v[1].value, cells=v[2].value, wide_pairs=v[3].value)

    def _params(self, b, keep):
        mat = np.ascontiguousarray(b.mat, dtype=np.int8)
        keep.append(mat)
        return MpnParams(mat.ctypes.data_as(ct.POINTER(ct.c_int8)), int(b.n), int(b.gapO), int(b.gapE), int(b.score_size), int(b.flag),
                         int(b.filters), int(b.filterd))

    def upload(self, b):
        keep = []
        p = self._params(b, keep)
        h = self.L.mpn_batch_upload(self.h, ct.byref(p), _ptr(b.reads), _ptr(b.read_off), _ptr(b.refs), _ptr(b.ref_off), _ptr(b.masklen), int(b.npairs))
        if not h:
            raise RuntimeError("mpn_batch_upload failed (bad arguments)")
        return h

    def run(self, h):
        rc = self.L.mpn_batch_run(h)
        if rc:
            raise RuntimeError(f"mpn_batch_run -> {rc}")

    def fetch(self, h, npairs, cigar_cap, out=None, cig=None):
        out = np.zeros(npairs, dtype=RESULT_DTYPE) if out is None else out
        cig = np.zeros(max(cigar_cap, 1), dtype=np.uint32) if cig is None else cig
        rc = self.L.mpn_batch_fetch(h, _ptr(out), _ptr(cig), int(cigar_cap))
        if rc:
            raise RuntimeError(f"mpn_batch_fetch -> {rc}")
        return out, cig

    def free(self, h):
        self.L.mpn_batch_free(h)

    def align(self, b, cigar_cap=None, out=None, cig=None):
        """The public call: host buffers in, host records out (mpn_align_batch; large batches are pipelined in chunks inside).
        b: anything with the PairBatch fields (workloads.PairBatch).  Returns (records[npairs] of RESULT_DTYPE, cigar arena)."""
        n = int(b.npairs)
        if cigar_cap is None:
            cigar_cap = n * 24 + int(len(b.reads)) // 4 + 4096
        out = np.zeros(n, dtype=RESULT_DTYPE) if out is None else out
        cig = np.zeros(max(cigar_cap, 1), dtype=np.uint32) if cig is None else cig
        keep = []
        p = self._params(b, keep)
        rc = self.L.mpn_align_batch(self.h, ct.byref(p), _ptr(b.reads), _ptr(b.read_off), _ptr(b.refs), _ptr(b.ref_off), _ptr(b.masklen), n,
                                    _ptr(out), _ptr(cig), int(cigar_cap))
        if rc:
            raise RuntimeError(f"mpn_align_batch -> {rc}")
        return out, cig


def pack4(codes):
    """int8 codes -> nibble-packed uint8 array for Engine.align_packed4 (mpn_pack4: base i in the low / high nibble of byte i // 2)"""
    codes = np.ascontiguousarray(codes, dtype=np.int8)
    out = np.zeros((len(codes) + 1) // 2, dtype=np.uint8)
    lib().mpn_pack4(_ptr(codes), len(codes), _ptr(out))
    return out


def pack2(codes):
    """int8 codes -> (four-per-byte uint8 array, exceptions) for Engine.align_packed2 (mpn_pack2: base i in bits 2 (i & 3) of byte i // 4; codes
    above 3 are stored as 0 and listed as position << 4 | code, sorted)"""
    codes = np.ascontiguousarray(codes, dtype=np.int8)
    out = np.zeros((len(codes) + 3) // 4, dtype=np.uint8)
    L = lib()
    L.mpn_pack2.restype = ct.c_int64
    L.mpn_pack2.argtypes = [ct.c_void_p, ct.c_int64, ct.c_void_p, ct.c_void_p, ct.c_int64]
    exc = np.zeros(1024, dtype=np.int64)
    n = L.mpn_pack2(_ptr(codes), len(codes), _ptr(out), _ptr(exc), len(exc))
    if n > len(exc):
        exc = np.zeros(n, dtype=np.int64)
        n = L.mpn_pack2(_ptr(codes), len(codes), _ptr(out), _ptr(exc), len(exc))
    return out, exc[:n].copy()


def _align_packed2(self, b, reads2, read_exc, refs2, ref_exc, cigar_cap=None, out=None, cig=None):
    """mpn_align_batch_packed2: b carries offsets (in bases), maskLen and scoring as usual; reads2 / refs2 are the four-per-byte streams,
    read_exc / ref_exc their exception lists (int64 arrays or pinned tensors)."""
    n = int(b.npairs)
    if cigar_cap is None:
        cigar_cap = n * 24 + int(b.read_off[-1]) // 4 + 4096
    out = np.zeros(n, dtype=RESULT_DTYPE) if out is None else out
    cig = np.zeros(max(cigar_cap, 1), dtype=np.uint32) if cig is None else cig
    keep = []
    p = self._params(b, keep)
    self.L.mpn_align_batch_packed2.argtypes = [ct.c_void_p, ct.POINTER(MpnParams), ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64, ct.c_void_p, ct.c_void_p, ct.c_void_p, ct.c_int64,
                                               ct.c_void_p, ct.c_int64, ct.c_void_p, ct.c_void_p, ct.c_int64]
    rc = self.L.mpn_align_batch_packed2(self.h, ct.byref(p), _ptr(reads2), _ptr(b.read_off), _ptr(read_exc) if len(read_exc) else None, len(read_exc),
                                        _ptr(refs2), _ptr(b.ref_off), _ptr(ref_exc) if len(ref_exc) else None, len(ref_exc), _ptr(b.masklen), n, _ptr(out), _ptr(cig), int(cigar_cap))
    if rc:
        raise RuntimeError(f"mpn_align_batch_packed2 -> {rc}")
    return out, cig


def _align_packed4(self, b, reads4, refs4, cigar_cap=None, out=None, cig=None):
    """mpn_align_batch_packed4: b carries offsets (in bases), maskLen and scoring as usual; reads4 / refs4 are the nibble-packed streams."""
    n = int(b.npairs)
    if cigar_cap is None:
        cigar_cap = n * 24 + int(b.read_off[-1]) // 4 + 4096
    out = np.zeros(n, dtype=RESULT_DTYPE) if out is None else out
    cig = np.zeros(max(cigar_cap, 1), dtype=np.uint32) if cig is None else cig
    keep = []
    p = self._params(b, keep)
    rc = self.L.mpn_align_batch_packed4(self.h, ct.byref(p), _ptr(reads4), _ptr(b.read_off), _ptr(refs4), _ptr(b.ref_off), _ptr(b.masklen), n, _ptr(out), _ptr(cig), int(cigar_cap))
    if rc:
        raise RuntimeError(f"mpn_align_batch_packed4 -> {rc}")
    return out, cig


Engine.align_packed4 = _align_packed4
Engine.align_packed2 = _align_packed2


class Pool:
    """One batch over several GPUs of the box (mpn_pool_*): an engine + a host thread per device, ranges of about equal cost pulled from
    one queue, records written straight into the caller's arrays -- no collective.  devices: list of ordinals,
    an int (the first n devices) or None (every visible device)."""

    def __init__(self, devices=None):
        self.L = lib()
        if devices is None:
            self.h = self.L.mpn_pool_create(None, 0)
        elif isinstance(devices, int):
            self.h = self.L.mpn_pool_create(None, int(devices))
        else:
            arr = (ct.c_int * len(devices))(*[int(d) for d in devices])
            self.h = self.L.mpn_pool_create(arr, len(devices))
        if not self.h:
            raise RuntimeError("mpn_pool_create failed: no CUDA device / bad device list (there is no CPU fallback)")
        self.ndev = self.L.mpn_pool_ndev(self.h)

    def close(self):
        if self.h:
            self.L.mpn_pool_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def align(self, b, cigar_cap=None, out=None, cig=None):
        n = int(b.npairs)
        if cigar_cap is None:
            cigar_cap = n * 24 + int(len(b.reads)) // 4 + 4096 + 2048 * max(self.ndev, 1) * 16
        out = np.zeros(n, dtype=RESULT_DTYPE) if out is None else out
        cig = np.zeros(max(cigar_cap, 1), dtype=np.uint32) if cig is None else cig
        keep = []
        p = Engine._params(self, b, keep)
        rc = self.L.mpn_pool_align_batch(self.h, ct.byref(p), _ptr(b.reads), _ptr(b.read_off), _ptr(b.refs), _ptr(b.ref_off), _ptr(b.masklen), n,
                                         _ptr(out), _ptr(cig), int(cigar_cap))
        if rc:
            raise RuntimeError(f"mpn_pool_align_batch -> {rc}")
        return out, cig

    def align_spans(self, b, seq, rd_start, rd_len, rf_start, rf_len, masklen, cigar_cap=None):
        seq = np.ascontiguousarray(seq, dtype=np.int8)
        rd_start = np.ascontiguousarray(rd_start, dtype=np.int64); rf_start = np.ascontiguousarray(rf_start, dtype=np.int64)
        rd_len = np.ascontiguousarray(rd_len, dtype=np.int32); rf_len = np.ascontiguousarray(rf_len, dtype=np.int32)
        masklen = np.ascontiguousarray(masklen, dtype=np.int32)
        n = len(rd_start)
        if cigar_cap is None:
            cigar_cap = int(rd_len.sum() + rf_len.sum()) + 16 * n + 4096
        out = np.zeros(n, dtype=RESULT_DTYPE)
        cig = np.zeros(max(cigar_cap, 1), dtype=np.uint32)
        keep = []
        p = Engine._params(self, b, keep)
        rc = self.L.mpn_pool_align_batch_spans(self.h, ct.byref(p), _ptr(seq), len(seq), _ptr(rd_start), _ptr(rd_len), _ptr(rf_start), _ptr(rf_len), _ptr(masklen), n,
                                               _ptr(out), _ptr(cig), int(cigar_cap))
        if rc:
            raise RuntimeError(f"mpn_pool_align_batch_spans -> {rc}")
        return out, cig

    def last_shares(self):
        """per device: (host wall ms, pairs, forward cells) of its share of the last batch"""
        ms = np.zeros(self.ndev, dtype=np.float64); pr = np.zeros(self.ndev, dtype=np.int64); ce = np.zeros(self.ndev, dtype=np.int64)
        self.L.mpn_pool_last_shares(self.h, _ptr(ms), _ptr(pr), _ptr(ce))
        return [dict(ms=float(ms[k]), pairs=int(pr[k]), cells=int(ce[k])) for k in range(self.ndev)]


def _align_spans(self, b, seq, rd_start, rd_len, rf_start, rf_len, masklen, cigar_cap=None):
    """Pairs that share sequences (mpn_align_batch_spans): `seq` is one int8 arena, pair i aligns seq[rd_start[i] : +rd_len[i]] to
    seq[rf_start[i] : +rf_len[i]].  `b` carries the scoring / flag fields of a PairBatch.  Returns (records, cigar arena)."""
    seq = np.ascontiguousarray(seq, dtype=np.int8)
    rd_start = np.ascontiguousarray(rd_start, dtype=np.int64); rf_start = np.ascontiguousarray(rf_start, dtype=np.int64)
    rd_len = np.ascontiguousarray(rd_len, dtype=np.int32); rf_len = np.ascontiguousarray(rf_len, dtype=np.int32)
    masklen = np.ascontiguousarray(masklen, dtype=np.int32)
    n = len(rd_start)
    if cigar_cap is None:
        cigar_cap = n * 24 + int(rd_len.sum()) // 4 + 4096
    out = np.zeros(n, dtype=RESULT_DTYPE)
    cig = np.zeros(max(cigar_cap, 1), dtype=np.uint32)
    keep = []
    p = self._params(b, keep)
    rc = self.L.mpn_align_batch_spans(self.h, ct.byref(p), _ptr(seq), len(seq), _ptr(rd_start), _ptr(rd_len), _ptr(rf_start), _ptr(rf_len), _ptr(masklen), n,
                                      _ptr(out), _ptr(cig), int(cigar_cap))
    if rc == -3:                                       # MPN_E_CIGAR_SPACE: size for the worst case once
        return _align_spans(self, b, seq, rd_start, rd_len, rf_start, rf_len, masklen, cigar_cap=int(rd_len.sum() + rf_len.sum()) + 16 * n)
    if rc:
        raise RuntimeError(f"mpn_align_batch_spans -> {rc}")
    return out, cig


Engine.align_spans = _align_spans


def as_table(rec, cig, cigar_cap=64):
    """records -> (int32 [n, 8] in FIELDS order with cigarLen = -1 for NULL results, uint32 [n, cigar_cap] cigar words)"""
    n = len(rec)
    t = np.zeros((n, 8), dtype=np.int32)
    for k, f in enumerate(("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigar_len")):
        t[:, k] = rec[f]
    null = rec["status"] != 0
    t[null] = 0
    t[null, 7] = -1
    c = np.zeros((n, cigar_cap), dtype=np.uint32)
    lens = np.minimum(np.where(null, 0, rec["cigar_len"]), cigar_cap)
    for i in np.nonzero(lens > 0)[0]:
        o = int(rec["cigar_off"][i])
        c[i, :lens[i]] = cig[o:o + lens[i]]
    return t, c
