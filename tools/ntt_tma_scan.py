#!/usr/bin/env python
"""NTT sweep at N = 2^10..2^12 for one setting of the TMA knobs (FHE_B200_NTT_TMA, FHE_B200_NTT_TMA_DEPTH read from the environment)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, _pkg
pkg = _pkg.load_package()
import bench
ctx = pkg.Context(0); ctx.use_torch_stream()
hbm = bench.peaks()[0]
rows = bench.ntt_sweep(pkg, ctx, torch, hbm, 20, [10, 11, 12], int(sys.argv[1]) if len(sys.argv) > 1 else 4096)
print("tma=%s depth=%s: " % (os.environ.get("FHE_B200_NTT_TMA"), os.environ.get("FHE_B200_NTT_TMA_DEPTH")) +
      "  ".join("2^%d/u%d %4.0f|%4.0f" % (r["log_n"], r["word_bits"], r["fwd_gbs"], r["inv_gbs"]) for r in rows))
