"""The CUDA library against the committed fixtures of tests/golden/tfhe_ckks.json, which were computed by the pure-Python
restatement of the reference alone (tests/pyref.py; generator tests/golden/make_golden.py) - no oracle in the loop."""
import json
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tfhe_ckks.json")))
A = lambda v: np.ascontiguousarray(v, dtype=np.uint64)


def test_fft64_product_fixture(pkg, ctx):
    from learn_fhe_b200 import tfhe
    for c in GOLD["fft64_mul"]:
        got = tfhe.nega_cyclic_fft64_mul_assign_rt(ctx, A(c["a"]).copy(), A(c["b"]))
        assert [int(x) for x in got] == c["out"]


def test_tggsw_and_tlwe_fixture(pkg, ctx):
    from learn_fhe_b200 import tfhe
    g, ks = GOLD["tggsw"], GOLD["tlwe_key_switch"]
    kn = g["k"] * g["n"]
    assert len(ks["a"]) == kn
    n_lwe = len(ks["ksk_a"][0])
    param = pkg.TfheParam(log_p=4, padding=1, n=n_lwe, ks_log_b=ks["log_b"], ks_d=ks["d"], log_big_n=g["n"].bit_length() - 1, k=g["k"],
                          bs_log_b=g["log_b"], bs_d=g["d"])
    brk = np.tile(A(g["rows"])[None], (n_lwe, 1, 1, 1))  # the same TGGSW ciphertext at every LWE index
    bk = tfhe.BootstrappingKey(ctx, param, brk, A(ks["ksk_a"]), A(ks["ksk_b"]))
    idx = np.array([0, n_lwe - 1], dtype=np.uint32)
    ct0, ct1 = np.stack([A(g["ct0"])] * 2), np.stack([A(g["ct1"])] * 2)
    got = tfhe.Tggsw.external_product(bk, idx, ct0)
    assert got[0].tolist() == g["external_product"] and got[1].tolist() == g["external_product"]
    got = tfhe.Tggsw.cmux(bk, idx, ct0, ct1)
    assert got[0].tolist() == g["cmux"] and got[1].tolist() == g["cmux"]
    ct = A([ks["a"] + [ks["b"]]])
    assert tfhe.Tlwe.key_switch(bk, ct)[0].tolist() == ks["out"]
    bk.free()


def test_programmable_bootstrap_fixture(pkg, ctx):
    from learn_fhe_b200 import tfhe
    g = GOLD["tfhe_pbs"]
    param = pkg.TfheParam(log_p=g["log_p"], padding=g["padding"], n=g["n"], ks_log_b=g["ks_log_b"], ks_d=g["ks_d"],
                          log_big_n=g["big_n"].bit_length() - 1, k=g["k"], bs_log_b=g["bs_log_b"], bs_d=g["bs_d"])
    bk = tfhe.BootstrappingKey(ctx, param, A(g["brk"]), A(g["ksk_a"]), A(g["ksk_b"]))
    got = tfhe.Bootstrapping.bootstrap(bk, tfhe.encode_lut(bk.param, A(g["v"])), A(g["cts"]))
    assert got.tolist() == g["out"]
    bk.free()


def test_rescale_and_ckks_fixture(pkg, ctx):
    from learn_fhe_b200 import ckks
    for c in GOLD["rns_rescale_k"]:
        x = A(c["x"]).T.copy()[None]  # [batch 1][limb][coefficient]; ring degree = number of coefficients (power of two: 8)
        got = ckks.rescale_k(ctx, c["qs"], c["k"], x)
        assert [[int(v) for v in got[0][:, i]] for i in range(x.shape[2])] == c["out"]
    g = GOLD["ckks"]
    P = ckks.CkksParam(ctx, g["log_n"], g["qs"], g["ps"])
    rlk = ckks.CkksKeySwitchingKey(P, A(g["ksk"]))
    ct0, ct1 = A(g["ct0"])[None], A(g["ct1"])[None]
    assert ckks.Ckks.mul(P, rlk, ct0, ct1)[0].tolist() == g["mul"]
    assert ckks.Ckks.key_switch(P, rlk, ct0, 0)[0].tolist() == g["key_switch_ct0"]
    keys = [ckks.CkksKeySwitchingKey(P, A(k)) for k in g["rot_keys"]]
    assert ckks.Ckks.key_switch(P, keys[0], ct0, 5)[0].tolist() == g["rotate5_ct0"]
    assert ckks.Ckks.mul_constant(P, A(g["pts"][0]), ct0)[0].tolist() == g["mul_constant_pt0_ct0"]
    mm = g["mul_mat"]
    baby = [(t, keys[0] if t else None) for t in mm["baby_t"]]
    giant = [(t, keys[1] if t else None) for t in mm["giant_t"]]
    assert ckks.Ckks.mul_mat(P, baby, giant, np.array(mm["present"], dtype=np.uint8), A(g["pts"]), ct0)[0].tolist() == mm["out"]
    for k in keys:
        k.free()
    rlk.free()
    P.free()
