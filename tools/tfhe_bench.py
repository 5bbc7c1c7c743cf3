#!/usr/bin/env python
"""TFHE PBS / CKKS hom-mult timing helper (single GPU): python tools/tfhe_bench.py [tfhe] [ckks] [--batch B]"""
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import _pkg  # noqa: E402

pkg = _pkg.load_package()
from learn_fhe_b200 import ckks, tfhe  # noqa: E402

args = sys.argv[1:]
ctx = pkg.Context(0)
ctx.use_torch_stream()


def timed(fn, reps):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / reps


if "bf" in args:
    print("u32 butterfly rate (T bf/s):", ctx.butterfly_rate())
if "peaks" in args:
    print("fp64 peaks (T instr/s):", ctx.fp64_peak(), "int32:", ctx.int32_peak())
if "tfhe" in args or not args:
    batch = int(args[args.index("--batch") + 1]) if "--batch" in args else 2048
    P = tfhe.bootstrapping_testing_param()
    rng = np.random.default_rng(1)
    n, N, k = P.n, P.big_n, P.k
    brk = rng.integers(0, 1 << 63, size=(n, (k + 1) * P.bs_d, k + 1, N), dtype=np.uint64)
    ksk_a = rng.integers(0, 1 << 63, size=(k * N * P.ks_d, n), dtype=np.uint64)
    ksk_b = rng.integers(0, 1 << 63, size=(k * N * P.ks_d,), dtype=np.uint64)
    bk = tfhe.BootstrappingKey(ctx, P, brk, ksk_a, ksk_b)
    lut = pkg.to_dev(rng.integers(0, 1 << 63, size=N, dtype=np.uint64))
    cts = pkg.to_dev(rng.integers(0, 1 << 63, size=(batch, n + 1), dtype=np.uint64))
    out = torch.empty_like(cts)
    modes = [int(m) for m in args[args.index("--modes") + 1].split(",")] if "--modes" in args else [0, 1, 2]
    for mode in modes:
        bk.set_mode(mode)
        ctx.prof_begin()
        ms = timed(lambda: tfhe.Bootstrapping.bootstrap_dev(bk, lut, cts, out), 2)
        print("tfhe pbs mode %d batch %d: %.2f ms -> %.0f PBS/s" % (mode, batch, ms, batch / ms * 1e3), ctx.prof_end())
if "ckks" in args:
    log_n, L = 16, 8
    count = int(args[args.index("--count") + 1]) if "--count" in args else 16
    P = ckks.CkksParam.new(ctx, log_n, 55, L)
    rng = np.random.default_rng(2)
    ksk = np.stack([np.stack([rng.integers(0, q, size=P.n, dtype=np.uint64) for q in P.qs + P.ps]) for _ in range(2)])
    rlk = ckks.CkksKeySwitchingKey(P, ksk)
    ct0 = torch.stack([torch.stack([torch.stack([torch.randint(0, q, (P.n,), dtype=torch.int64, device="cuda") for q in P.qs]) for _ in range(2)])
                       for _ in range(count)])
    ct1 = ct0.clone()
    out = torch.empty((count, 2, L - 1, P.n), dtype=torch.int64, device="cuda")
    ctx.prof_begin()
    ms = timed(lambda: ckks.Ckks.mul_dev(P, rlk, L, ct0, ct1, out), 2)
    print("ckks mul N=2^16 l=8 count %d: %.2f ms -> %.1f mult/s" % (count, ms, count / ms * 1e3), ctx.prof_end())
