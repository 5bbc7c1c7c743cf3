#!/usr/bin/env python
"""NTT throughput at N = 2^10..2^12 for 4096 / 16384 / 65536 polynomials (how much of the configured batch is launch, ramp and tail)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, _pkg
pkg = _pkg.load_package()
import bench
ctx = pkg.Context(0); ctx.use_torch_stream()
hbm = bench.peaks()[0]
for batch in (4096, 16384, 65536):
    rows = bench.ntt_sweep(pkg, ctx, torch, hbm, 20, [10, 11, 12], batch)
    for r in rows:
        print("batch %6d N=2^%-2d u%d  fwd %7.1f GB/s (%.4f ms)  inv %7.1f GB/s" % (batch, r["log_n"], r["word_bits"], r["fwd_gbs"], r["fwd_ms"], r["inv_gbs"]))
