#!/usr/bin/env python
"""NTT sweep only (bench.py's ntt leg) as a compact table; FHE_B200_LIB selects an A/B build of the library.
usage: python tools/ntt_bench.py [--reps R] [--min LOGN] [--max LOGN]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import _pkg  # noqa: E402

pkg = _pkg.load_package()
import bench  # noqa: E402

a = sys.argv[1:]
opt = lambda k, d: int(a[a.index(k) + 1]) if k in a else d
ctx = pkg.Context(0)
ctx.use_torch_stream()
hbm_peak = bench.peaks()[0]
rows = bench.ntt_sweep(pkg, ctx, torch, hbm_peak, opt("--reps", 20), list(range(opt("--min", 10), opt("--max", 16) + 1)), 4096)
print("lib:", pkg.LIB_PATH)
for r in rows:
    print("N=2^%-2d u%d  fwd %7.1f GB/s (%.3f ms)  inv %7.1f GB/s (%.3f ms)" % (r["log_n"], r["word_bits"], r["fwd_gbs"], r["fwd_ms"], r["inv_gbs"], r["inv_ms"]))
