#!/usr/bin/env python
"""Throughput of the 64-bit-modulus FHEW path at the parameter size of examples/multi_key_uint8.rs:15-29 (N = 2048, 55-bit Q,
d = 5, LWE n = 600) on synthetic key material: python tools/fhew_wide_bench.py [--batch B]"""
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

a = sys.argv[1:]
batch = int(a[a.index("--batch") + 1]) if "--batch" in a else 592
ctx = pkg.Context(0)
ctx.use_torch_stream()
log_n = 11
q = pkg.first_two_adic_prime(55, log_n + 1)
param = pkg.FhewParam(log_n=log_n, big_q=q, p=4, rlwe_log_b=11, rlwe_d=5, rgsw_log_b=11, rgsw_d=5, n_s=600, q_ks=1 << 20, ks_log_b=4,
                      ks_d=5, w=10)
bk = fhew.BootstrappingKey(ctx, param, *bench.synth_fhew_key(param, 5))
f = pkg.to_dev(fhew.gate_poly(param, [1, 1, 1, 0]))
cin = pkg.to_dev(bench.synth_cts(param, batch, 6))
cout = torch.empty_like(cin)
post = fhew.big_q_by_8(param)
fhew.Bootstrapping.bootstrap_dev(bk, f, cin, cout, post_add=post)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2):
    fhew.Bootstrapping.bootstrap_dev(bk, f, cin, cout, post_add=post)
e1.record()
e1.synchronize()
ms = e0.elapsed_time(e1) / 2
print("FHEW 64-bit path, N=2048 Q=2^55 d=5 n_s=600: batch %d in %.1f ms -> %.0f gates/s (key %.0f MB)" % (batch, ms, batch / ms * 1e3, bk.nbytes / 1e6))
