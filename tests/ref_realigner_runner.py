"""Runs the compiled REFERENCE `realigner` (oracle/_ref/realigner_ref) over regions given as JSON on stdin and prints the
results as JSON.  Must stay free of numpy / torch imports: the reference object carries a statically linked libstdc++ whose
GNU-unique symbols clash with the system libstdc++ those packages load (std::regex then crashes), so the parity tests
run this file in a clean subprocess.  TEST INFRASTRUCTURE ONLY.

  python tests/ref_realigner_runner.py <path to realigner shared object>  < regions.json  > results.json"""
import ctypes
import json
import sys


class StructPointer(ctypes.Structure):
    _fields_ = [("position", ctypes.c_int * 1000), ("cigar_string", ctypes.c_char_p * 1000)]


def main():
    L = ctypes.cdll.LoadLibrary(sys.argv[1])
    L.realign_reads.restype = ctypes.POINTER(StructPointer)
    L.free_memory.restype = None
    L.free_memory.argtypes = [ctypes.POINTER(StructPointer), ctypes.c_int]
    out = []
    for rg in json.load(sys.stdin):
        n = min(1000, len(rg["reads"]))
        seqs = (ctypes.c_char_p * n)(*[s.encode() for s in rg["reads"][:n]])
        pos = (ctypes.c_int * n)(*rg["positions"][:n])
        cig = (ctypes.c_char_p * n)(*[c.encode() for c in rg["cigars"][:n]])
        L.realign_reads.argtypes = [ctypes.c_char_p * n, ctypes.c_int * n, ctypes.c_char_p * n, ctypes.c_char_p, ctypes.c_char_p,
                                    ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int]
        p = L.realign_reads(seqs, pos, cig, ctypes.c_char_p(rg["reference"].encode()), ctypes.c_char_p(" ".join(rg["haplotypes"]).encode()),
                            rg["ref_start"], rg["ref_prefix"], rg["ref_suffix"], n)
        out.append([list(p.contents.position[:n]), [c.decode() for c in p.contents.cigar_string[:n]]])
        L.free_memory(p, n)
    json.dump(out, sys.stdout)


if __name__ == "__main__":
    main()
