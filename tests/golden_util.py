"""Loads tests/golden/ssw_golden.npz (made by tests/golden/make_golden.py from the compiled reference ssw.c)."""
import importlib
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
w = importlib.import_module("workloads")
CAP = 192


def golden_cases():
    z = np.load(os.path.join(HERE, "golden", "ssw_golden.npz"))
    for k in range(int(z["ncases"])):
        pre = f"c{k}_"
        n, gapO, gapE, flag, filters, filterd, score_size = [int(v) for v in z[pre + "params"]]
        b = w.PairBatch(z[pre + "reads"], z[pre + "read_off"], z[pre + "refs"], z[pre + "ref_off"], z[pre + "masklen"], mat=z[pre + "mat"], n=n,
                        gapO=gapO, gapE=gapE, flag=flag, filters=filters, filterd=filterd, score_size=score_size, name=str(z[pre + "name"]))
        yield k, b, z[pre + "res"], z[pre + "cigar"]


def diff(res_a, cig_a, res_b, cig_b):
    """indices of pairs whose 8 fields or CIGAR words (first CAP) differ"""
    cap = min(cig_a.shape[1], cig_b.shape[1])
    bad = (res_a != res_b).any(axis=1) | (cig_a[:, :cap] != cig_b[:, :cap]).any(axis=1)
    return np.nonzero(bad)[0]
