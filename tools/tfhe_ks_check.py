#!/usr/bin/env python
"""Checksum of a 2000-ciphertext TFHE mode-3 batch on synthetic keys: compares kernel variants (FHE_B200_TFHE_KEY_SMEM=0/1) bit for bit."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, _pkg
pkg = _pkg.load_package()
from learn_fhe_b200 import tfhe
ctx = pkg.Context(0); ctx.use_torch_stream()
P = tfhe.bootstrapping_testing_param()
rng = np.random.default_rng(1)
n, N, k = P.n, P.big_n, P.k
brk = rng.integers(0, 1 << 63, size=(n, (k + 1) * P.bs_d, k + 1, N), dtype=np.uint64)
ksk_a = rng.integers(0, 1 << 63, size=(k * N * P.ks_d, n), dtype=np.uint64)
ksk_b = rng.integers(0, 1 << 63, size=(k * N * P.ks_d,), dtype=np.uint64)
bk = tfhe.BootstrappingKey(ctx, P, brk, ksk_a, ksk_b)
lut = pkg.to_dev(rng.integers(0, 1 << 63, size=N, dtype=np.uint64))
cts = pkg.to_dev(rng.integers(0, 1 << 63, size=(2000, n + 1), dtype=np.uint64))
cts[3, 5] = 0
out = torch.empty_like(cts)
bk.set_mode(3)
tfhe.Bootstrapping.bootstrap_dev(bk, lut, cts, out); ctx.sync()
print("checksum", int(out.sum().item()) & 0xFFFFFFFFFFFF, int(out[3].sum().item()) & 0xFFFFFFFF)
