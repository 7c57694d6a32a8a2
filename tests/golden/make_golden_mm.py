#!/usr/bin/env python
"""Regenerates tests/golden/mm/*.mtx and tests/golden/mm_expected.npz.

Run in the authoring container (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden_mm.py
The .mtx files are small hand-shaped MatrixMarket inputs (comments, blank lines, upper-case
banner tokens, exponents, a symmetric file with explicit zeros, integer and pattern storage).
The expected arrays are what the REFERENCE's own reader (src/utils/mmread.cpp + the mmio.c it
vendors, behind oracle/mm_ref_shim.cpp) returns for each file, and what its mmutils.hpp
unfolding makes of the symmetric one."""
import ctypes
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

FILES = {
    "general_real.mtx": """%%MatrixMarket matrix coordinate real general
% a comment
%another

5 4 7
1 1 1.5
2 3 -2.25e-3
5 4 1e300
3 1 0.1
3 3 3
4 2 -7.0E+2
1 4 .5
""",
    "symmetric_real.mtx": """%%MatrixMarket MATRIX Coordinate REAL Symmetric
4 4 6
1 1 2.0
2 1 -1.0
3 2 0.0
3 3 2.0
4 1 0.333333333333333314829616256247
4 4 0
""",
    "integer_general.mtx": """%%MatrixMarket matrix coordinate integer general
3 3 4
1 2 7
2 2 -3
3 1 0
3 3 12
""",
    "pattern_general.mtx": """%%MatrixMarket matrix coordinate pattern general
3 5 4
1 5
2 2
3 1
3 4
""",
}


def main():
    from tests import util
    R = util.mm_ref_lib()
    assert R is not None, "build oracle/_ref/libmm_ref.so first (make -C oracle ref)"
    out = {}
    for name, text in FILES.items():
        path = os.path.join(HERE, "mm", name)
        with open(path, "w") as f:
            f.write(text)
        props = (ctypes.c_int * 6)()
        assert R.ref_mm_properties(path.encode(), props) == 1
        key = name[:-4]
        out[key + "/props"] = np.array(list(props), dtype=np.int32)
        n = props[2]
        rows, cols = np.zeros(n, np.int32), np.zeros(n, np.int32)
        for suffix, dt in (("float", np.float32), ("double", np.float64), ("int", np.int32)):
            vals = np.zeros(n, dt)
            rc = getattr(R, f"ref_mm_load_{suffix}")(path.encode(), vals.ctypes.data, rows.ctypes.data, cols.ctypes.data)
            out[f"{key}/{suffix}/rc"] = np.array([rc], np.int32)
            if rc == 0:
                out[f"{key}/{suffix}/vals"], out[f"{key}/rows"], out[f"{key}/cols"] = vals.copy(), rows.copy(), cols.copy()
        rc = R.ref_mm_load_pattern(path.encode(), rows.ctypes.data, cols.ctypes.data)
        out[f"{key}/pattern/rc"] = np.array([rc], np.int32)
        if rc == 0:
            out[f"{key}/rows"], out[f"{key}/cols"] = rows.copy(), cols.copy()
        if props[5] == 1:
            for suffix, dt in (("float", np.float32), ("double", np.float64)):
                vals = out[f"{key}/{suffix}/vals"]
                r, c = out[f"{key}/rows"].copy(), out[f"{key}/cols"].copy()
                m = getattr(R, f"ref_mm_unfolded_size_{suffix}")(vals.ctypes.data, r.ctypes.data, c.ctypes.data, n)
                ur, uc, uv = np.zeros(m, np.int32), np.zeros(m, np.int32), np.zeros(m, dt)
                getattr(R, f"ref_mm_unfold_{suffix}")(ur.ctypes.data, uc.ctypes.data, uv.ctypes.data, r.ctypes.data,
                                                      c.ctypes.data, vals.ctypes.data, n)
                out[f"{key}/{suffix}/unfolded_rows"], out[f"{key}/{suffix}/unfolded_cols"] = ur, uc
                out[f"{key}/{suffix}/unfolded_vals"] = uv
    np.savez_compressed(os.path.join(HERE, "mm_expected.npz"), **out)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
