"""ctypes access to the parity checkers.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
It never touches the GPU and nothing in megapath-nano_b200/ imports it.

  ref_lib()      the UNMODIFIED reference ssw.c compiled into oracle/_ref/libssw_ref.so (oracle/Makefile)
  port           the scalar restatement oracle/ssw_oracle.c (liboracle.so)
  run_batch(...) one pair per call, `threads` POSIX threads (oracle/batch_driver.c), returns (results[n,8], cigars, seconds)
"""
import ctypes as ct
import os
import subprocess
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_ORACLE_SO = os.path.join(HERE, "liboracle.so")
_REF_SO = os.path.join(HERE, "_ref", "libssw_ref.so")
_REALIGNER_REF = os.path.join(HERE, "_ref", "realigner_ref")
_DBG_REF = os.path.join(HERE, "_ref", "debruijn_graph_ref")

FIELDS = ("score1", "score2", "ref_begin1", "ref_end1", "read_begin1", "read_end1", "ref_end2", "cigarLen")


def build(force=False):
    """Compile liboracle.so and, if /root/reference is present, oracle/_ref (outputs stay under oracle/)."""
    need = force or not os.path.exists(_ORACLE_SO) or (
        os.path.getmtime(_ORACLE_SO) < max(os.path.getmtime(os.path.join(HERE, f)) for f in ("ssw_oracle.c", "batch_driver.c")))
    have_ref_src = os.path.exists("/root/reference/bin/realignment/realign/ssw.c")
    if need:
        subprocess.run(["make", "-C", HERE, os.path.join(HERE, "liboracle.so")], check=True, capture_output=True)
    if have_ref_src and (force or not os.path.exists(_REF_SO) or not os.path.exists(_REALIGNER_REF) or not os.path.exists(_DBG_REF)):
        subprocess.run(["make", "-C", HERE, "ref"], check=True, capture_output=True)


_HOSTTEST_SO = os.path.join(HERE, "_hosttest", "realigner_hosttest.so")


def build_hosttest(force=False):
    """the product's host-side realigner sources linked against the CPU checkers (oracle/host_shim.cpp); returns the path"""
    build()
    srcs = [os.path.join(HERE, "host_shim.cpp"), os.path.join(HERE, "ssw_oracle.c"), os.path.join(HERE, "batch_driver.c")]
    pkg = os.path.join(HERE, "..", "megapath-nano_b200", "csrc")
    srcs += [os.path.join(pkg, f) for f in ("realign_region.cpp", "ssw_cpp_layer.cpp", "host_shared.h")]
    if force or not os.path.exists(_HOSTTEST_SO) or os.path.getmtime(_HOSTTEST_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.run(["make", "-C", HERE, "hosttest"], check=True, capture_output=True)
    return _HOSTTEST_SO


_REFSRC = "/root/reference/bin/realignment"
_CALLERS = os.path.join(HERE, "..", "baseline", "_ref")


def callers_dir():
    return os.path.abspath(_CALLERS)


def build_ref_callers(force=False):
    """The reference's OWN callers of the hot path, unmodified, for the boundary tests (tests/test_gpu_reference_callers.py):
      baseline/_ref/{pyssw.py, fast_align_reads2ref.py}          copied as they are (they load realign/libssw.so next to them)
      baseline/_ref/{ssw_cpp.cpp, ssw_cpp.h, realigner.cpp, realigner.h, ssw.c, ssw.h(as ssw_ref.h)}    copied as they are
      baseline/_ref/realigner_refsrc_on_product    reference ssw_cpp.cpp + realigner.cpp compiled against include/ssw.h and LINKED TO THE PRODUCT
                                                   libssw.so -- INTEGRATION.md section 1's recipe (reference README.md:45 with ssw.c replaced)
      baseline/_ref/ssw_cpp_probe_ref              tests/cpp/ssw_cpp_probe_common.cpp on the reference's ssw_cpp.cpp + ssw.c (all CPU)
    baseline/_ref is git-ignored and travels to the GPU box; nothing here is product code.  Returns the directory, or None when neither
    the reference sources nor a previous build are present."""
    import shutil
    d = callers_dir()
    root = os.path.abspath(os.path.join(HERE, ".."))
    obj = os.path.join(d, "realigner_refsrc_on_product")
    probe = os.path.join(d, "ssw_cpp_probe_ref")
    have_src = os.path.exists(os.path.join(_REFSRC, "pyssw.py"))
    if not have_src:
        return d if os.path.exists(obj) and os.path.exists(os.path.join(d, "pyssw.py")) else None
    os.makedirs(d, exist_ok=True)
    for f in ("pyssw.py", "fast_align_reads2ref.py"):
        shutil.copyfile(os.path.join(_REFSRC, f), os.path.join(d, f))
    for f in ("ssw_cpp.cpp", "ssw_cpp.h", "realigner.cpp", "realigner.h", "ssw.c"):
        shutil.copyfile(os.path.join(_REFSRC, "realign", f), os.path.join(d, f))
    prod = os.path.join(root, "megapath-nano_b200", "realign")
    if os.path.exists(os.path.join(prod, "libssw.so")) and (force or not os.path.exists(obj) or os.path.getmtime(obj) < os.path.getmtime(os.path.join(prod, "libssw.so"))):
        # `#include "ssw.h"` in the copied sources resolves to include/ssw.h (no ssw.h is copied next to them)
        subprocess.run(["g++", "-std=c++14", "-O1", "-shared", "-fPIC", "-w", "-I", os.path.join(root, "include"), "-o", obj,
                        os.path.join(d, "ssw_cpp.cpp"), os.path.join(d, "realigner.cpp"), os.path.join(prod, "libssw.so"), "-Wl,-rpath," + prod], check=True)
    src = os.path.join(root, "tests", "cpp", "ssw_cpp_probe_common.cpp")
    if force or not os.path.exists(probe) or os.path.getmtime(probe) < os.path.getmtime(src):
        refh = os.path.join(d, "_refhdr")
        os.makedirs(refh, exist_ok=True)
        shutil.copyfile(os.path.join(_REFSRC, "realign", "ssw.h"), os.path.join(refh, "ssw.h"))          # the all-CPU probe uses the reference's own header
        subprocess.run(["gcc", "-O2", "-w", "-I", refh, "-c", "-o", os.path.join(d, "ssw_ref.o"), os.path.join(d, "ssw.c")], check=True)
        subprocess.run(["g++", "-std=c++14", "-O1", "-w", "-I", refh, "-I", d, "-o", probe, src, os.path.join(d, "ssw_cpp.cpp"), os.path.join(d, "ssw_ref.o")], check=True)
    return d


_port = None
_ref = None


def port_lib():
    global _port
    if _port is None:
        build()
        _port = ct.CDLL(_ORACLE_SO)
        _port.oracle_run_batch.restype = ct.c_double
        _port.oracle_run_batch_port.restype = ct.c_double
    return _port


def have_ref():
    return os.path.exists(_REF_SO)


def require_ref():
    """GPU-tier parity is only ever claimed against the compiled UNMODIFIED reference: build it when the sources are present, otherwise
    raise (tests call pytest.fail on this; nothing degrades to the restatement)."""
    if not (os.path.exists(_REF_SO) and os.path.exists(_REALIGNER_REF)):
        try:
            build()
        except Exception:
            pass
    if not (os.path.exists(_REF_SO) and os.path.exists(_REALIGNER_REF)):
        raise RuntimeError("oracle/_ref/{libssw_ref.so,realigner_ref} missing: run `make -C oracle ref` where /root/reference exists "
                           "(the prebuilt files travel to the GPU box); the GPU parity tier never falls back to the restatement")
    return _REF_SO


def ref_lib():
    global _ref
    if _ref is None:
        build()
        _ref = ct.CDLL(_REF_SO)
    return _ref


def ref_path():
    return _REF_SO


def realigner_ref_path():
    return _REALIGNER_REF


def dbg_ref_path():
    """the reference's debruijn_graph.cpp, unmodified, compiled over oracle/boost_shim (None when not built)"""
    if not os.path.exists(_DBG_REF):
        try:
            build()
        except Exception:
            pass
    return _DBG_REF if os.path.exists(_DBG_REF) else None


def _p(a, t):
    return a.ctypes.data_as(ct.POINTER(t))


def run_batch(reads, read_off, refs, ref_off, masklen, mat, n, gapO=8, gapE=2, flag=1, filters=0, filterd=32767, score_size=2,
              threads=1, impl="ref", cigar_cap=64, lib=None):
    """impl: 'ref' (compiled reference ssw.c), 'port' (scalar restatement), or 'lib' with lib=<CDLL exporting the ssw.h ABI>."""
    reads = np.ascontiguousarray(reads, dtype=np.int8)
    refs = np.ascontiguousarray(refs, dtype=np.int8)
    read_off = np.ascontiguousarray(read_off, dtype=np.int64)
    ref_off = np.ascontiguousarray(ref_off, dtype=np.int64)
    masklen = np.ascontiguousarray(masklen, dtype=np.int32)
    mat = np.ascontiguousarray(mat, dtype=np.int8)
    npairs = len(read_off) - 1
    out = np.zeros((npairs, 8), dtype=np.int32)
    cig = np.zeros((npairs, cigar_cap), dtype=np.uint32)
    P = port_lib()
    common = (_p(reads, ct.c_int8), _p(read_off, ct.c_int64), _p(refs, ct.c_int8), _p(ref_off, ct.c_int64), _p(masklen, ct.c_int32),
              _p(mat, ct.c_int8), ct.c_int32(n), ct.c_int32(gapO), ct.c_int32(gapE), ct.c_int32(flag), ct.c_int32(filters), ct.c_int32(filterd),
              ct.c_int32(score_size), ct.c_int64(npairs), ct.c_int32(threads), _p(out, ct.c_int32), _p(cig, ct.c_uint32), ct.c_int32(cigar_cap))
    if impl == "port":
        secs = P.oracle_run_batch_port(*common)
    else:
        L = ref_lib() if impl == "ref" else lib
        fns = [ct.cast(getattr(L, nm), ct.c_void_p) for nm in ("ssw_init", "ssw_align", "init_destroy", "align_destroy")]
        secs = P.oracle_run_batch(*fns, *common)
    return out, cig, secs
