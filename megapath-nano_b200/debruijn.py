"""Host-side mirror of the reference's ctypes binding of the `debruijn_graph` shared object
(realign_illumina_reads.py:30,33 loader, :46-48 `DBGPointer`, :541-562 the get_consensus call), on the Boost-free assembler
megapath-nano_b200/csrc/debruijn_assemble.cpp (include/debruijn_graph.h).

  get_consensus(ref, reads, base_qualities)   one window, exactly the reference's call -> sorted candidate haplotypes
  consensus_windows(windows)                  NEW: many windows in one call, spread over the host threads

Host code: this is the step that produces the haplotypes the realigner (realigner.py) aligns on the GPU."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPN_DBG_LIB") or os.path.join(HERE, "realign", "debruijn_graph")
min_low_base_quality = 15            # realign_illumina_reads.py:549


class DBGPointer(ctypes.Structure):                         # realign_illumina_reads.py:46-48 (the C struct has 500 slots, debruijn_graph.h:40-44)
    _fields_ = [("consensus_size", ctypes.c_int),
                ("consensus", ctypes.c_char_p * 500)]


_libs = {}


def load(path=None):
    path = path or LIB_PATH
    if path not in _libs:
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is not built (run __graft_entry__.build())")
        L = ctypes.cdll.LoadLibrary(path)
        L.get_consensus.restype = ctypes.POINTER(DBGPointer)
        L.get_consensus.argtypes = [ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_int]
        L.free_memory.restype = None
        L.free_memory.argtypes = [ctypes.POINTER(DBGPointer), ctypes.c_int]
        if hasattr(L, "mpn_dbg_consensus_packed"):          # absent in the reference's own object (the binding drives both, as the parity tests do)
            L.mpn_dbg_consensus_packed.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p,
                                                   ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_longlong)]
            L.mpn_dbg_free.argtypes = [ctypes.c_void_p]
        _libs[path] = L
    return _libs[path]


def byte(x):
    return x if isinstance(x, bytes) else x.encode()


def low_quality_field(base_quality):
    """realign_illumina_reads.py:548-549: positions of bases below quality 15, joined by ' '"""
    return " ".join(str(i) for i, q in enumerate(base_quality) if int(q) < min_low_base_quality)


def get_consensus(ref, reads, low_quality_fields, lib_path=None):
    """reads: list of str; low_quality_fields: list of str as built by low_quality_field (one per read).
    Same call sequence as realign_illumina_reads.py:551-562."""
    L = load(lib_path)
    p = L.get_consensus(byte(ref), byte(",".join(reads)), byte(",".join(low_quality_fields)), len(reads))
    n = p.contents.consensus_size
    consensus = [item.decode() for item in p.contents.consensus[:n]]
    L.free_memory(p, n)
    return consensus


def consensus_windows(windows, lib_path=None, with_k=False):
    """windows: iterable of (ref, reads, low_quality_fields).  One call of mpn_dbg_consensus_packed -> list of haplotype lists."""
    import array
    L = load(lib_path)
    parts = []
    for ref, reads, lowq in windows:
        parts.extend((ref, ",".join(reads), ",".join(lowq)))
    nw = len(parts) // 3
    if parts:
        parts.append("")
    text = "\0".join(parts).encode() if parts else b""
    counts, ks = array.array("i", [0] * max(nw, 1)), array.array("i", [0] * max(nw, 1))
    out, out_bytes = ctypes.c_void_p(), ctypes.c_longlong(0)
    rc = L.mpn_dbg_consensus_packed(text, len(text), nw, ctypes.c_void_p(counts.buffer_info()[0]), ctypes.c_void_p(ks.buffer_info()[0]),
                                    ctypes.byref(out), ctypes.byref(out_bytes))
    if rc:
        raise RuntimeError(f"mpn_dbg_consensus_packed -> {rc}")
    flat = ctypes.string_at(out.value, out_bytes.value).decode().split("\0")[:-1] if out_bytes.value else []
    L.mpn_dbg_free(out)
    res, at = [], 0
    for w in range(nw):
        res.append(flat[at:at + counts[w]]); at += counts[w]
    return (res, list(ks[:nw])) if with_k else res


expand_align_ref_region = 20         # realign_illumina_reads.py:18


def regions_from_windows(windows, lib_path=None):
    """The chain of realign_illumina_reads.py:551-593 for many windows (anything with chrom, chrom_start, win_start, win_end, reads, positions, cigars, low_quality): assemble the candidate haplotypes
    of every window (one mpn_dbg_consensus_packed call), drop windows without an alternative haplotype (:564-566), widen the reference
    to the span of the reads + 20 bp (:567-577) and build the realigner's inputs (:593).  Returns (regions, kept window indices);
    the regions go to realigner.realign_regions."""
    import importlib
    Region = importlib.import_module(__package__ + ".realigner").Region
    cons = consensus_windows([(w.chrom[w.win_start:w.win_end], w.reads, w.low_quality) for w in windows], lib_path)
    regions, kept = [], []
    for k, (w, consensus) in enumerate(zip(windows, cons)):
        ref = w.chrom[w.win_start:w.win_end]
        if len(consensus) == 0 or (len(consensus) == 1 and consensus[0] == ref) or len(w.reads) == 0:
            continue
        start_pos, end_pos = w.chrom_start + w.win_start, w.chrom_start + w.win_end
        min_read_start = min(w.positions)
        max_read_end = max(p + len(r) for p, r in zip(w.positions, w.reads))
        tmp_ref_start = max(w.chrom_start, min(min_read_start, start_pos) - expand_align_ref_region)
        tmp_ref_end = min(w.chrom_start + len(w.chrom), max(max_read_end, end_pos) + expand_align_ref_region)
        ref_prefix = w.chrom[tmp_ref_start - w.chrom_start:w.win_start]
        ref_suffix = w.chrom[w.win_end:tmp_ref_end - w.chrom_start]
        n = min(1000, len(w.reads))
        regions.append(Region(ref_prefix + ref + ref_suffix, [ref_prefix + c + ref_suffix for c in consensus],
                                        [r.upper() for r in w.reads[:n]], list(w.positions[:n]), list(w.cigars[:n]), tmp_ref_start, len(ref_prefix), len(ref_suffix)))
        kept.append(k)
    return regions, kept
