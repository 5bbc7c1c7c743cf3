#!/usr/bin/env python
"""Short single-GPU command for ncu captures: one FHEW bootstrap batch and/or a few batched NTTs.
usage: python tools/prof_cmd.py [fhew] [ntt] [--batch B]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import _pkg  # noqa: E402

pkg = _pkg.load_package()
from learn_fhe_b200 import fhew  # noqa: E402
import bench  # noqa: E402

args = sys.argv[1:]
batch = int(args[args.index("--batch") + 1]) if "--batch" in args else 148 * 8
ctx = pkg.Context(0)
if "fhew" in args or not args:
    param = fhew.single_key_testing_param(bench.FHEW_T_Q)
    bk = fhew.BootstrappingKey(ctx, param, *bench.synth_fhew_key(param, 1))
    f = pkg.to_dev(fhew.gate_poly(param, [1, 1, 1, 0]))
    cin = pkg.to_dev(bench.synth_cts(param, batch, 2))
    cout = torch.empty_like(cin)
    for _ in range(2):
        fhew.Bootstrapping.bootstrap_dev(bk, f, cin, cout, post_add=fhew.big_q_by_8(param))
    ctx.sync()
if "ntt" in args:
    for log_n, b in ((12, 4096), (16, 512)):
        for bits in (64, 32):
            q = pkg.first_two_adic_prime(55 if bits == 64 else 28, log_n + 1)
            t = torch.zeros((b << log_n) // (1 if bits == 64 else 2), dtype=torch.int64, device="cuda")
            for _ in range(2):
                ctx.call("fhe_ntt_fwd_u%d" % bits, q, log_n, b, pkg.dptr(t))
                ctx.call("fhe_ntt_inv_u%d" % bits, q, log_n, b, pkg.dptr(t))
            ctx.sync()
if "ntt16" in args:  # BASELINE configs[1]'s largest case, for the DRAM-traffic sum: 3 forward transforms of 4096 x 2^16 u64
    q = pkg.first_two_adic_prime(55, 17)
    t = torch.zeros(4096 << 16, dtype=torch.int64, device="cuda")
    for _ in range(3):
        ctx.call("fhe_ntt_fwd_u64", q, 16, 4096, pkg.dptr(t))
    ctx.sync()
print("prof_cmd done, launches:", ctx.launches)
