"""The C-ABI boundary without a GPU: the libraries load, export every function the headers in include/ declare, and the
struct layouts that ctypes callers mirror (pyssw.py:7-16) are what the reference's are."""
import ctypes as ct
import importlib
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "megapath-nano_b200")
LIBS = [os.path.join(PKG, "libmpn_ssw.so"), os.path.join(PKG, "realign", "libssw.so"), os.path.join(PKG, "realign", "realigner")]


def declared_functions(header):
    src = open(os.path.join(ROOT, "include", header)).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"^\s*(?:[A-Za-z_][\w\s\*]*?)\b(\w+)\s*\([^;{]*\)\s*;", src, flags=re.M)
    return [n for n in names if n not in ("defined",)]


@pytest.fixture(scope="module")
def built():
    if not all(os.path.exists(p) for p in LIBS):
        subprocess.run(["bash", os.path.join(ROOT, "build.sh")], check=True)
    return LIBS


def test_headers_declare_the_reference_entry_points():
    assert set(declared_functions("ssw.h")) >= {"ssw_init", "init_destroy", "ssw_align", "align_destroy"}
    assert set(declared_functions("mpn_ssw_batch.h")) >= {"mpn_engine_create", "mpn_align_batch", "mpn_align_batch_spans", "mpn_batch_upload", "mpn_batch_upload_spans",
                                                          "mpn_batch_run", "mpn_batch_fetch", "mpn_batch_free"}
    assert set(declared_functions("realigner.h")) >= {"realign_reads", "free_memory", "mpn_realign_regions"}


def test_libraries_export_every_declared_symbol(built):
    want = declared_functions("ssw.h") + declared_functions("mpn_ssw_batch.h") + declared_functions("realigner.h")
    for path in built:
        lib = ct.CDLL(path)
        for name in want:
            assert hasattr(lib, name), (path, name)
    # the assembler is its own shared object (its free_memory has a different argument type than the realigner's)
    dbg = os.path.join(PKG, "realign", "debruijn_graph")
    if not os.path.exists(dbg):
        subprocess.run(["bash", os.path.join(ROOT, "build.sh"), "dbg"], check=True)
    lib = ct.CDLL(dbg)
    assert set(declared_functions("debruijn_graph.h")) >= {"get_consensus", "free_memory", "mpn_dbg_consensus_packed"}
    for name in declared_functions("debruijn_graph.h"):
        assert hasattr(lib, name), (dbg, name)


def test_s_align_layout_is_the_reference_layout(tmp_path):
    # compile a 10-line C program against include/ssw.h and compare offsets with the ctypes mirror the reference's callers use
    src = tmp_path / "lay.c"
    src.write_text('#include <stddef.h>\n#include "ssw.h"\nint main(void){printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(s_align), offsetof(s_align,score1),'
                   'offsetof(s_align,score2), offsetof(s_align,ref_begin1), offsetof(s_align,ref_end1), offsetof(s_align,read_begin1), offsetof(s_align,read_end1),'
                   'offsetof(s_align,ref_end2), offsetof(s_align,cigar), offsetof(s_align,cigarLen)); return 0;}\n')
    exe = tmp_path / "lay"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [40, 0, 2, 4, 8, 12, 16, 20, 24, 32]
    pyssw = importlib.import_module("megapath-nano_b200.pyssw")
    assert ct.sizeof(pyssw.CAlignRes) == 40 and pyssw.CAlignRes.sCigar.offset == 24 and pyssw.CAlignRes.nCigarLen.offset == 32


def test_struct_str_arr_layout_and_cpp_front_end_symbols(built, tmp_path):
    """struct_str_arr is what realign_illumina_reads.py:40-43 mirrors (1000 ints, 1000 pointers), and the C++ front end of
    include/ssw_cpp.h links: a translation unit that uses the class compiles and resolves against libmpn_ssw.so"""
    R = importlib.import_module("megapath-nano_b200.realigner")
    assert ct.sizeof(R.StructPointer) == 1000 * 4 + 1000 * 8 and R.StructPointer.cigar_string.offset == 4000
    src = tmp_path / "lay2.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "realigner.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(struct_str_arr), offsetof(struct_str_arr, cigar_string), sizeof(mpn_region)); return 0;}\n')
    exe = tmp_path / "lay2"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [12000, 4000, ct.sizeof(R.MpnRegion)]
    cpp = tmp_path / "use.cpp"
    cpp.write_text('#include "ssw_cpp.h"\nint probe(){ StripedSmithWaterman::Aligner a(4, 6, 8, 2); StripedSmithWaterman::Filter f; StripedSmithWaterman::Alignment al;'
                   ' std::vector<StripedSmithWaterman::PairView> pv; std::vector<StripedSmithWaterman::Alignment> out; a.SetReferenceSequence("ACGT", 4);'
                   ' std::vector<StripedSmithWaterman::SeqView> pool; std::vector<StripedSmithWaterman::PairIndex> pi;'
                   ' return (int)a.AlignPairs(pv, f, &out) + (int)a.AlignIndexed(pool, pi, f, &out) + (int)sizeof(al); }\n')
    so = tmp_path / "use.so"
    subprocess.run(["g++", "-std=c++17", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"), "-o", str(so), str(cpp), built[0], "-Wl,-z,defs",
                    "-Wl,-rpath," + PKG], check=True)


def test_cigar_helpers_match_reference_encoding(tmp_path):
    src = tmp_path / "cig.c"
    src.write_text('#include "ssw.h"\nint main(void){const char* ops="MIDNSHP=X"; for(int i=0;i<9;i++){unsigned v=to_cigar_int(7,ops[i]); if((v&15)!=(unsigned)i||cigar_int_to_len(v)!=7||cigar_int_to_op(v)!=ops[i]) return 1;}'
                   ' if(cigar_int_to_op(0x7f)!=\'M\') return 2; if(to_cigar_int(3,\'?\')!=(3u<<4)) return 3; return 0;}\n')
    exe = tmp_path / "cig"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    assert subprocess.run([str(exe)]).returncode == 0


def test_no_gpu_means_loud_failure_not_fallback(built):
    """without a CUDA device engine creation must fail (NULL), never silently compute on the CPU"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    B = importlib.import_module("megapath-nano_b200.batch")
    with pytest.raises(RuntimeError):
        B.Engine()


def test_product_never_imports_the_oracle():
    for dirpath, _, files in os.walk(PKG):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt, os.path.join(dirpath, f)


def test_fastpass_and_assembler_struct_layouts(tmp_path):
    """the structs bindings have to mirror: mpn_fp_region / mpn_placement (include/mpn_ssw_batch.h) and dbg_str_arr (include/debruijn_graph.h)"""
    src = tmp_path / "lay2.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include <stdint.h>\n#include "mpn_ssw_batch.h"\n#include "debruijn_graph.h"\n'
                   'int main(void){printf("%zu %zu %zu %zu %zu %zu %zu\\n", sizeof(mpn_fp_region), offsetof(mpn_fp_region, hap_first), offsetof(mpn_fp_region, suffix),'
                   'sizeof(mpn_placement), sizeof(dbg_str_arr), offsetof(dbg_str_arr, consensus), sizeof(mpn_result)); return 0;}\n')
    exe = tmp_path / "lay2"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()]
    assert got == [32, 8, 28, 8, 8 + 500 * 8, 8, 40]
    D = importlib.import_module("megapath-nano_b200.debruijn")
    assert ct.sizeof(D.DBGPointer) == got[4] and D.DBGPointer.consensus.offset == got[5]


def test_pack4_is_host_code_and_matches_numpy(built):
    """mpn_pack4 (the host helper of the nibble-packed entry point) needs no GPU: base i -> low / high nibble of byte i // 2"""
    import numpy as np
    B = importlib.import_module("megapath-nano_b200.batch")
    rng = np.random.default_rng(3)
    for n in (0, 1, 2, 7, 1000, 12345):
        codes = rng.integers(0, 5, size=n, dtype=np.int8)
        got = B.pack4(codes)
        pad = np.concatenate([codes, np.zeros(n % 2, np.int8)]).astype(np.uint8)
        want = (pad[0::2] | (pad[1::2] << 4)).astype(np.uint8)
        assert got.dtype == np.uint8 and (got == want).all()


def test_pack2_is_host_code_and_matches_numpy(built):
    """mpn_pack2 (the host helper of the 2-bit entry point) needs no GPU: base i -> bits 2 (i & 3) of byte i // 4, codes above 3 as exceptions"""
    import numpy as np
    B = importlib.import_module("megapath-nano_b200.batch")
    rng = np.random.default_rng(4)
    for n in (0, 1, 3, 4, 9, 1000, 12345):
        codes = rng.integers(0, 5, size=n, dtype=np.int8)
        got, exc = B.pack2(codes)
        pad = np.concatenate([codes, np.zeros((-n) % 4, np.int8)]).astype(np.uint8)
        pad = np.where(pad > 3, 0, pad).astype(np.uint8)
        want = (pad[0::4] | (pad[1::4] << 2) | (pad[2::4] << 4) | (pad[3::4] << 6)).astype(np.uint8)
        assert got.dtype == np.uint8 and (got == want).all()
        pos = np.nonzero(codes > 3)[0]
        assert (exc == ((pos.astype(np.int64) << 4) | codes[pos].astype(np.int64))).all()
