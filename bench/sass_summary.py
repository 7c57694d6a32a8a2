#!/usr/bin/env python
"""Opcode histograms (memory, math, synchronisation instructions) of the hot kernels in spgpu_b200/lib/libspgpu.so from
`cuobjdump -sass` -- the proof that the build carries sm_100a code with the instructions DESIGN.md names (LDG.E.EF.LTC64B,
UBLKCP, LDG.E.STRONG.SYS, ACQBULK / PREEXIT of programmatic dependent launch ...).  No GPU needed.
    python bench/sass_summary.py > profiles/sass_r2_hot_kernels.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "spgpu_b200", "lib", "libspgpu.so")
KEEP = re.compile(r"^(LDG|STG|LDS|STS|LDL|STL|ATOM|RED|DFMA|DADD|DMUL|FFMA|SHFL|UBLKCP|SYNCS|BAR|MEMBAR|CCTL|NANOSLEEP|ACQBULK|PREEXIT|CALL|UTMA|FENCE|ERRBAR)")
WANT = [
    ("default cfg5 kernel", "hell_spmv_kernel<double, 8, 32, 10>"),
    ("cfg3 kernel", "hell_spmv_kernel<float, 8, 32, 12>"),
    ("cfg4 kernel", "hell_spmv_kernel<double2, 4, 32, 8>"),
    ("cfg2 kernel", "hdia_spmv_kernel<double, 9, 32, 8, false, 128>"),
    ("cfg1 kernel", "ell_spmv_short_kernel<double, 5, 1, 16, 128>"),
    ("DIA", "dia_spmv_kernel<double, "),
    ("fused SpMV + halo (multi-GPU)", "spmv_halo_kernel<double, HellRowBody<double, 8, 32>, 10, false, true>"),
    ("fused SpMV + halo + dot", "spmv_halo_kernel<double, HellRowBody<double, 8, 32>, 10, true, true>"),
    ("fused SpMV + dot (one GPU)", "spmv_halo_kernel<double, HellRowBody<double, 8, 32>, 12, true, false>"),
    ("fused HDIA SpMV + halo", "spmv_halo_kernel<double, HdiaRowBody<double, 9, 32>, 8, false, true>"),
    ("bulk-async (TMA) HELL pipeline, hellVariant=3", "hell_spmv_bulk_kernel<double, 32, 8>"),
    ("bulk-async HDIA pipeline, hdiaVariant=4", "hdia_spmv_bulk_kernel<double, 32, "),
    ("dot", "reduce_kernel<double, OpDot<double>, 4>"),
    ("axpby", "ew_kernel<double, 2, "),
    ("fused CG update (+ all-reduce in the last CTA)", "cg_update_kernel<double>"),
    ("fold (+ all-reduce)", "fold_partials_kernel<double>"),
    ("separate halo exchange", "halo_exchange_kernel("),
    ("all-reduce kernel", "allreduce_sum_kernel<double>"),
]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    archs = collections.Counter(re.findall(r"arch = (sm_\w+)", sass))
    funcs = {}
    name = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = m.group(1)
            funcs[name] = []
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and name:
            funcs[name].append(m.group(1))
    demangled = dict(zip(funcs, subprocess.run(["c++filt"], input="\n".join(funcs), capture_output=True, text=True).stdout.splitlines()))
    print("SASS opcode histograms of the hot kernels in spgpu_b200/lib/libspgpu.so (cuobjdump -sass; bench/sass_summary.py)")
    print(f"cubins: {dict(archs)}\n")
    for title, key in WANT:
        hit = [f for f, d in demangled.items() if key in d]
        if not hit:
            print(f"## {title}: {key}: NOT FOUND")
            continue
        f = hit[0]
        ops = funcs[f]
        hist = collections.Counter(o for o in ops if KEEP.match(o))
        print(f"## {title}\n   {demangled[f][:150]}\n   instructions: {len(ops)}")
        print("   " + ", ".join(f"{o} x{c}" for o, c in sorted(hist.items(), key=lambda kv: (-kv[1], kv[0]))) + "\n")


if __name__ == "__main__":
    sys.exit(main())
