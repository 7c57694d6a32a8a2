#!/usr/bin/env python
"""Regenerates tests/golden/conversions.npz from the REFERENCE library.

Run in the authoring container (needs /root/reference and `make -C oracle ref`):
    python tests/golden/make_golden.py
For each seeded input matrix the reference's own host conversion code
(src/core/ell.c, hell.c, dia.c, hdia.cpp, built unmodified into
oracle/_ref/libspgpu_ref.so) produces every metadata array; they are stored so
the CPU test-suite can check our C port bit-exactly without the reference."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from spgpu_b200 import capi, formats as F, generators as G  # noqa: E402


def cases():
    """name -> Coo.  Kept small: the fixture is committed."""
    out = {}
    # 70-row 1-D 3-point Laplacian, row 40 gets six extra entries (40, 3k)=0.5
    rows, cols, vals = [], [], []
    for i in range(70):
        for j, v in ((i - 1, -1.0), (i, 2.0), (i + 1, -1.0)):
            if 0 <= j < 70:
                rows.append(i); cols.append(j); vals.append(v)
        if i == 40:
            for k in range(6):
                rows.append(40); cols.append(3 * k); vals.append(0.5)
    out["survey_b1"] = F.Coo(np.array(rows, np.int32), np.array(cols, np.int32), np.array(vals, np.float64), 70, 70, 0)
    out["lap2d_13x9"] = G.laplace2d_5pt(13, 9)
    out["st27_6"] = G.stencil3d_27pt(6)
    out["lap7_7"] = G.laplace3d_7pt(7)
    out["powerlaw_700"] = G.powerlaw(700, 6, 200, 128, 3, np.float32)
    out["banded_z_500_b1"] = G.banded_complex(500, 9, 40, 11, np.complex128, base=1)
    out["ragged_c_97x83"] = G.random_coo(97, 83, (0, 11), 5, np.complex64, 0)
    out["ragged_d_33x65_b1"] = G.random_coo(33, 65, (0, 7), 6, np.float64, 1)
    out["unsorted_s_64x64"] = G.random_coo(64, 64, (0, 9), 7, np.float32, 0, sort_cols=False)
    out["ctest_100"] = F.Coo((np.arange(200) % 100).astype(np.int32), (np.arange(200) % 100).astype(np.int32),
                             np.ones(200, np.float32), 100, 100, 0)
    return out


def convert_all(coo, L, hack):
    """every conversion output, as a flat dict of arrays (values as raw bytes)"""
    d = {}
    for ell_base in (0, 1):
        ell = F.coo_to_ell(coo, ell_base, L)
        p = f"ell{ell_base}_"
        d[p + "rs"], d[p + "indices"], d[p + "values"] = ell.rs, ell.indices, ell.values.view(np.uint8)
        d[p + "meta"] = np.array([ell.pitch, ell.maxnnz], np.int64)
        if ell_base == 0:
            oell = F.ell_to_oell(ell, L)
            d["oell_ridx"], d["oell_rs"], d["oell_indices"] = oell.ridx, oell.rs, oell.indices
            d["oell_values"] = oell.values.view(np.uint8)
            hell = F.ell_to_hell(ell, hack, L)
            d["hell_hack_offsets"] = hell.hack_offsets
            d["hell_height"] = np.array([hell.height], np.int64)
            live = live_mask(hell)
            d["hell_indices_live"] = hell.indices[live]
            d["hell_values_live"] = hell.values[live].view(np.uint8)
    dia = F.coo_to_dia(coo, L)
    d["dia_offsets"], d["dia_values"] = dia.offsets, dia.values.view(np.uint8)
    d["dia_meta"] = np.array([dia.pitch, dia.diags], np.int64)
    hd = F.coo_to_hdia(coo, hack, L)
    d["hdia_hack_offsets"], d["hdia_offsets"], d["hdia_values"] = hd.hack_offsets, hd.offsets, hd.values.view(np.uint8)
    hd2 = F.dia_to_hdia(dia, hack, L)
    d["dhdia_hack_offsets"], d["dhdia_offsets"], d["dhdia_values"] = hd2.hack_offsets, hd2.offsets, hd2.values.view(np.uint8)
    return d


def live_mask(hell):
    """positions of the slots ellToHell writes (padding content is undefined)"""
    m = np.zeros(hell.values.shape[0], dtype=bool)
    hs = hell.hack_size
    for h in range(hell.hack_offsets.shape[0]):
        rows = hell.rs[h * hs:(h + 1) * hs]
        for k in range(int(rows.max()) if rows.size else 0):
            m[int(hell.hack_offsets[h]) + k * hs + np.nonzero(rows > k)[0]] = True
    return m


def main():
    ref = capi.SpgpuLib(os.path.join(ROOT, "oracle", "_ref", "libspgpu_ref.so"), ext=False)
    blob = {}
    for name, coo in cases().items():
        for hack in (32, 64):
            for k, v in convert_all(coo, ref, hack).items():
                blob[f"{name}/h{hack}/{k}"] = v
    path = os.path.join(ROOT, "tests", "golden", "conversions.npz")
    np.savez_compressed(path, **blob)
    print(f"wrote {path}: {len(blob)} arrays, {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
