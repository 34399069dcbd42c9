"""Host-side mirror of the reference's ctypes binding of the `realigner` shared object
(realign_illumina_reads.py:29-43 `StructPointer`, :586-612 the realign_reads / free_memory calls), on the B200 engine.

  realign_reads(region)            one region, exactly the reference's call -> (positions, cigar strings)
  realign_regions(list of regions) NEW: many regions, every Smith-Waterman pair of all of them in ONE GPU batch

There is no CPU path: the shared object aborts without a CUDA device."""
import ctypes
import dataclasses
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("MPN_REALIGNER_LIB") or os.path.join(HERE, "realign", "realigner")
max_region_reads_num = 1000          # realign_illumina_reads.py:17


class StructPointer(ctypes.Structure):                      # realign_illumina_reads.py:40-43 <-> struct_str_arr, realigner.h:42-46
    _fields_ = [("position", ctypes.c_int * max_region_reads_num),
                ("cigar_string", ctypes.c_char_p * max_region_reads_num)]


class MpnRegion(ctypes.Structure):                          # include/realigner.h mpn_region
    _fields_ = [("seqs", ctypes.POINTER(ctypes.c_char_p)), ("positions", ctypes.POINTER(ctypes.c_int)), ("cigars", ctypes.POINTER(ctypes.c_char_p)),
                ("read_size", ctypes.c_int), ("reference", ctypes.c_char_p), ("haplotypes", ctypes.c_char_p),
                ("ref_start", ctypes.c_int), ("ref_prefix", ctypes.c_int), ("ref_suffix", ctypes.c_int)]


_libs = {}


def load(path=None):
    path = path or LIB_PATH
    if path not in _libs:
        if not os.path.exists(path):
            raise RuntimeError(f"{path} is not built (run __graft_entry__.build()); there is no CPU fallback")
        L = ctypes.cdll.LoadLibrary(path)
        L.realign_reads.restype = ctypes.POINTER(StructPointer)
        L.free_memory.restype = None
        L.free_memory.argtypes = [ctypes.POINTER(StructPointer), ctypes.c_int]
        _libs[path] = L
    return _libs[path]



@dataclasses.dataclass
class Region:
    """one realigner call: the arguments of realign_reads (realigner.cpp:856, realign_illumina_reads.py:586-605) as Python values"""
    reference: str
    haplotypes: list
    reads: list
    positions: list
    cigars: list
    ref_start: int
    ref_prefix: int
    ref_suffix: int

def byte(x):
    return x if isinstance(x, bytes) else x.encode()


def _marshal(region):
    n = min(max_region_reads_num, len(region.reads))
    seq_list = (ctypes.c_char_p * n)(*[byte(s) for s in region.reads[:n]])
    position_list = (ctypes.c_int * n)(*[int(p) for p in region.positions[:n]])
    cigars_list = (ctypes.c_char_p * n)(*[byte(c) for c in region.cigars[:n]])
    return n, seq_list, position_list, cigars_list


def realign_reads(region, lib_path=None):
    """region: anything with reference, haplotypes (list), reads, positions, cigars, ref_start, ref_prefix, ref_suffix
    (realigner.Region).  Same call sequence as realign_illumina_reads.py:586-612; works against the reference's own
    `realigner` as well (lib_path) -- that is how the parity tests drive both."""
    L = load(lib_path)
    n, seq_list, position_list, cigars_list = _marshal(region)
    L.realign_reads.argtypes = [ctypes.c_char_p * n, ctypes.c_int * n, ctypes.c_char_p * n, ctypes.c_char_p, ctypes.c_char_p,
                                ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
    p = L.realign_reads(seq_list, position_list, cigars_list, ctypes.c_char_p(byte(region.reference)), ctypes.c_char_p(byte(" ".join(region.haplotypes))),
                        int(region.ref_start), int(region.ref_prefix), int(region.ref_suffix), n)
    positions = list(p.contents.position[:n])
    cigars = [c.decode() for c in p.contents.cigar_string[:n]]
    L.free_memory(p, n)
    return positions, cigars


def realign_regions(regions, lib_path=None):
    """NEW batched entry point (mpn_realign_regions): returns [(positions, cigars)] per region."""
    L = load(lib_path)
    nr = len(regions)
    arr = (MpnRegion * max(nr, 1))()
    keep = []
    for k, region in enumerate(regions):
        n, seq_list, position_list, cigars_list = _marshal(region)
        ref_b, hap_b = byte(region.reference), byte(" ".join(region.haplotypes))
        keep.append((seq_list, position_list, cigars_list, ref_b, hap_b))
        arr[k] = MpnRegion(ctypes.cast(seq_list, ctypes.POINTER(ctypes.c_char_p)), ctypes.cast(position_list, ctypes.POINTER(ctypes.c_int)),
                           ctypes.cast(cigars_list, ctypes.POINTER(ctypes.c_char_p)), n, ref_b, hap_b,
                           int(region.ref_start), int(region.ref_prefix), int(region.ref_suffix))
    out = (ctypes.POINTER(StructPointer) * max(nr, 1))()
    L.mpn_realign_regions.argtypes = [ctypes.POINTER(MpnRegion), ctypes.c_int, ctypes.POINTER(ctypes.POINTER(StructPointer))]
    rc = L.mpn_realign_regions(arr, nr, out)
    if rc:
        raise RuntimeError(f"mpn_realign_regions -> {rc}")
    res = []
    for k in range(nr):
        n = arr[k].read_size
        res.append((list(out[k].contents.position[:n]), [c.decode() for c in out[k].contents.cigar_string[:n]]))
        L.free_memory(out[k], n)
    return res


def realign_regions_packed(regions, lib_path=None):
    """Same result as realign_regions through the flat-buffer entry point (mpn_realign_regions_packed): one bytes join and three int
    arrays instead of an array of C strings per region -- the marshalling cost of 10^5 reads drops from ~0.15 s to ~10 ms."""
    import array
    L = load(lib_path)
    nr = len(regions)
    parts, nreads, geom, positions = [], array.array("i"), array.array("i"), array.array("i")
    for rg in regions:
        n = min(max_region_reads_num, len(rg.reads))
        parts.append(rg.reference); parts.append(" ".join(rg.haplotypes))
        parts.extend(rg.reads[:n]); parts.extend(rg.cigars[:n])
        nreads.append(n); geom.extend((int(rg.ref_start), int(rg.ref_prefix), int(rg.ref_suffix)))
        pos = rg.positions[:n]
        positions.extend(pos if all(type(p) is int for p in pos) else [int(p) for p in pos])     # plain ints go in without per-item conversion
    if parts:
        parts.append("")                                                # the join then ends with the last terminator: no second 20 MB copy
    text = "\0".join(parts).encode() if parts else b""
    total = len(positions)
    out_pos = (ctypes.c_int * max(total, 1))()
    out_cig = ctypes.c_void_p(); out_bytes = ctypes.c_longlong(0)
    L.mpn_realign_regions_packed.argtypes = [ctypes.c_char_p, ctypes.c_longlong, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                             ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_longlong)]
    L.mpn_realign_free.argtypes = [ctypes.c_void_p]
    addr = lambda a: ctypes.c_void_p(a.buffer_info()[0]) if len(a) else ctypes.c_void_p(0)
    dummy = array.array("i", [0])
    rc = L.mpn_realign_regions_packed(text, len(text), nr, addr(nreads) if nr else addr(dummy), addr(geom) if nr else addr(dummy), addr(positions) if total else addr(dummy),
                                      out_pos, ctypes.byref(out_cig), ctypes.byref(out_bytes))
    if rc:
        raise RuntimeError(f"mpn_realign_regions_packed -> {rc}")
    cigs = ctypes.string_at(out_cig.value, out_bytes.value).decode().split("\0")[:-1] if out_bytes.value else []
    L.mpn_realign_free(out_cig)
    res, at = [], 0
    for n in nreads:
        res.append((list(out_pos[at:at + n]), cigs[at:at + n])); at += n
    return res


def last_stats(lib_path=None):
    L = load(lib_path)
    pairs, cells = ctypes.c_longlong(0), ctypes.c_longlong(0)
    secs = (ctypes.c_double * 3)()
    L.mpn_realign_last_stats(ctypes.byref(pairs), ctypes.byref(cells), secs)
    return dict(pairs=pairs.value, cells=cells.value, fast_pass_s=secs[0], gpu_s=secs[1], compose_s=secs[2])


def fastpass_only(regions, which=0, lib_path=None):
    """Test / A-B hook (mpn_realign_fastpass_only): the k-mer fast pass alone.  which = 0: GPU kernel, 1: host k-mer index.
    Returns (hap_scores, places[(score, pos)], kernel_ms), both flat in region / haplotype / read order."""
    import array
    L = load(lib_path)
    nr = len(regions)
    arr = (MpnRegion * max(nr, 1))()
    keep = []
    nh = npl = 0
    for k, region in enumerate(regions):
        n, seq_list, position_list, cigars_list = _marshal(region)
        ref_b, hap_b = byte(region.reference), byte(" ".join(region.haplotypes))
        keep.append((seq_list, position_list, cigars_list, ref_b, hap_b))
        arr[k] = MpnRegion(ctypes.cast(seq_list, ctypes.POINTER(ctypes.c_char_p)), ctypes.cast(position_list, ctypes.POINTER(ctypes.c_int)),
                           ctypes.cast(cigars_list, ctypes.POINTER(ctypes.c_char_p)), n, ref_b, hap_b,
                           int(region.ref_start), int(region.ref_prefix), int(region.ref_suffix))
        nhap = len(" ".join(region.haplotypes).split())
        nh += nhap; npl += nhap * n
    scores = (ctypes.c_int * max(nh, 1))()
    places = (ctypes.c_int * max(2 * npl, 1))()
    ms = ctypes.c_double(0)
    L.mpn_realign_fastpass_only.argtypes = [ctypes.POINTER(MpnRegion), ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.POINTER(ctypes.c_double)]
    rc = L.mpn_realign_fastpass_only(arr, nr, int(which), scores, places, ctypes.byref(ms))
    if rc:
        raise RuntimeError(f"mpn_realign_fastpass_only -> {rc}")
    return list(scores[:nh]), list(places[:2 * npl]), ms.value
