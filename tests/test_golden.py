"""Golden vectors: outputs of the REAL reference (oracle/_ref/libsbref.so built from
/root/reference) on small seeded inputs, committed under tests/golden/ by tests/golden/generate.py.
The oracle must reproduce them (copies bit for bit)."""
import glob
import os

import numpy as np

from tests import cases as C

HERE = os.path.dirname(os.path.abspath(__file__))


def load_golden():
    out = []
    for f in sorted(glob.glob(os.path.join(HERE, "golden", "*.npz"))):
        z = np.load(f, allow_pickle=True)
        case = z["case"].item()
        g = dict(kind=str(z["kind"]), case=case)
        if g["kind"] in ("copy", "masked_copy"):
            g["v0"], g["v1"] = C.make_copy_data(case, int(z["seed"]), consistent=True)
            g["want"] = [z["want_%d" % j] for j in range(len(g["v1"]))]
            g["m0"], g["m1"] = C.make_masks(case, int(z["seed"])) if g["kind"] == "masked_copy" \
                else (None, None)
        else:
            g["v0"], g["v1"], g["vr"] = C.make_contraction_data(case, int(z["seed"]))
            g["want"] = [z["want_%d" % j] for j in range(len(g["vr"]))]
        out.append((os.path.basename(f), g))
    return out


def test_golden_files_exist():
    assert len(load_golden()) >= 40


def test_oracle_reproduces_golden_vectors():
    for name, g in load_golden():
        case = g["case"]
        if g["kind"] in ("copy", "masked_copy"):
            got = C.oracle_copy(case, g["v0"], g["v1"], g["m0"], g["m1"])
            for j, (x, w) in enumerate(zip(got, g["want"])):
                assert C.bits_equal(x, w), (name, j)
        else:
            got = C.oracle_contraction(case, g["v0"], g["v1"], g["vr"])
            tol = 1e-12 if case["T"] in (np.dtype(np.float64), np.dtype(np.complex128)) else 1e-5
            for x, w in zip(got, g["want"]):
                if w.size:
                    assert np.linalg.norm(x - w) <= tol * max(np.linalg.norm(w), np.sqrt(w.size)), name
