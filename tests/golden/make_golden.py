#!/usr/bin/env python
"""Regenerates tests/golden/*.json from tests/pyref.py (the pure-Python restatement of the reference written from the
Rust sources).  The reference itself holds no golden vectors (every test draws from an entropy-seeded RNG, SURVEY.md
§4) and cannot be run here (no Rust toolchain), so these fixtures pin the *restated* semantics: any later change to
the oracle, the CUDA kernels or pyref that alters a bit shows up as a diff against the committed files.

usage: python tests/golden/make_golden.py"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import pyref  # noqa: E402


def rnd(seed, n, q):
    return [int(x) for x in np.random.default_rng(seed).integers(0, q, size=n, dtype=np.uint64)]


def main():
    g = {}
    # --- NTT (fft.rs:40-77, fft/zq.rs:58-67) ---
    ntt = []
    for bits, log_n in ((28, 3), (28, 6), (45, 5), (55, 7), (61, 4)):
        q = pyref.two_adic_primes(bits, log_n + 1, 1)[0]
        a = rnd(100 + log_n, 1 << log_n, q)
        a[0], a[1] = 0, q - 1
        ntt.append({"q": q, "a": a, "fwd": pyref.ntt_fwd(q, a), "generator": pyref.zq_generator(q)})
    g["ntt"] = ntt
    # FHEW-T modulus: root and first twiddles
    q = pyref.two_adic_primes(28, 10, 1)[0]
    tw, twi = pyref.compute_twiddle(q)
    g["fhew_t_modulus"] = {"q": q, "generator": pyref.zq_generator(q), "omega": pyref.zq_two_adic_generator(q, 10),
                           "tw_first16": tw[:16], "tw_inv_first16": twi[:16], "tw_len": len(tw)}
    # --- negacyclic product vs schoolbook ---
    q = pyref.two_adic_primes(45, 5, 1)[0]
    a, b = rnd(1, 16, q), rnd(2, 16, q)
    g["negacyclic_mul"] = {"q": q, "a": a, "b": b, "out": pyref.schoolbook_negacyclic(a, b, q)}
    # --- decomposers (decompose.rs) ---
    dz = []
    for q, log_b, d in ((268409857, 7, 4), (1 << 16, 4, 4), (pyref.two_adic_primes(55, 12, 1)[0], 11, 5), (268409857, 5, 4), (97, 2, 3)):
        v = rnd(7, 24, q) + [0, 1, q - 1, q // 2, q // 2 + 1, q // 2 - 1]
        dz.append({"q": q, "log_b": log_b, "d": d, "v": v, "digits": [pyref.decompose_zq(q, log_b, d, x) for x in v]})
    g["decompose_zq"] = dz
    dt = []
    for log_b, d in ((23, 1), (4, 5), (8, 8), (2, 8), (1, 3)):
        v = rnd(8, 24, 1 << 64) + [0, 1, (1 << 64) - 1, 1 << 63, (1 << 63) - 1]
        dt.append({"log_b": log_b, "d": d, "v": v, "digits": [pyref.decompose_t64(log_b, d, x) for x in v]})
    g["decompose_t64"] = dt
    # --- mod switches (zq.rs:128-140) ---
    ms = []
    for q, qp in ((268409857, 1 << 16), (1 << 16, 1024), (1024, 268409857)):
        v = rnd(9, 40, q) + [0, 1, q - 1, q // 2]
        ms.append({"q": q, "qp": qp, "v": v, "mod_switch": [pyref.zq_mod_switch(q, x, qp) for x in v],
                   "mod_switch_odd": [pyref.zq_mod_switch_odd(q, x, qp) for x in v]})
    g["mod_switch"] = ms
    # --- automorphism / monomial (avec.rs:34-50, ring.rs:299-313) ---
    q = 268409857
    a = rnd(10, 16, q)
    g["automorphism"] = [{"q": q, "a": a, "t": t, "out": pyref.automorphism(a, t, q)} for t in (5, -5, 31, 3, 33)]
    g["monomial_mul"] = [{"q": q, "a": a, "k": k, "out": pyref.monomial_mul(a, k, q)} for k in (0, 1, 15, 16, 17, 31, -1, -17, 35)]
    # --- RNS fast base conversion (rns.rs:331-345) ---
    qs = pyref.two_adic_primes(55, 5, 6)
    rns = []
    for i in range(8):
        x = [rnd(20 + i, 1, qi)[0] for qi in qs[:3]]
        rns.append({"x": x, "out": pyref.rns_extend_bases(qs[:3], qs[3:], x)})
    g["rns_extend_bases"] = {"qs": qs[:3], "ps": qs[3:], "cases": rns}
    # --- LMKCDEY schedule + tiny FHEW bootstrap in the REFERENCE dataflow (schoolbook products) ---
    P = {"n": 16, "log_n": 4, "big_q": pyref.two_adic_primes(20, 5, 1)[0], "q_ks": 1 << 10, "ks_log_b": 2, "ks_d": 5,
         "rgsw_log_b": 5, "rgsw_d": 4, "rlwe_log_b": 4, "rlwe_d": 5, "n_s": 6, "w": 3, "p": 4}
    n, q = P["n"], P["big_q"]
    rng = np.random.default_rng(4242)
    I = lambda hi, shape: rng.integers(0, hi, size=shape, dtype=np.uint64)  # noqa: E731
    keys = {"ksk_a": I(P["q_ks"], (n * P["ks_d"], P["n_s"])).tolist(), "ksk_b": I(P["q_ks"], (n * P["ks_d"],)).tolist(),
            "brk": I(q, (P["n_s"], 2 * P["rgsw_d"], 2, n)).tolist(), "ak": I(q, (P["w"] + 1, P["rlwe_d"], 2, n)).tolist(),
            "ak_t": [2 * n - 5] + [pow(5, v, 2 * n) for v in range(1, P["w"] + 1)]}
    f, q8 = pyref.fhew_gate_poly(P, [1, 1, 1, 0])
    cases = []
    for i in range(3):
        ct = (I(q, (n,)).tolist(), int(I(q, (1,))[0]))
        a2, b2 = ct
        ksa = [pyref.zq_mod_switch(q, v, P["q_ks"]) for v in a2]
        ksb = pyref.zq_mod_switch(q, b2, P["q_ks"])
        ka, kb = pyref.lwe_key_switch(P["q_ks"], P["ks_log_b"], P["ks_d"], keys["ksk_a"], keys["ksk_b"], ksa, ksb)
        pro = [pyref.zq_mod_switch_odd(P["q_ks"], v, 2 * n) for v in ka] + [pyref.zq_mod_switch_odd(P["q_ks"], kb, 2 * n)]
        sched = pyref.blind_rotate_schedule(n, P["w"], pro[:-1])
        oa, ob = pyref.fhew_bootstrap(P, keys, f, ct)
        cases.append({"ct": ct[0] + [ct[1]], "prologue": pro, "schedule": [[0 if k == "ext" else 1, j] for k, j in sched],
                      "out": oa + [(ob + q8) % q]})
    g["fhew_tiny"] = {"param": P, "keys": keys, "table": [1, 1, 1, 0], "f": f, "post_add": q8, "cases": cases}
    with open(os.path.join(HERE, "util_fhew.json"), "w") as fh:
        json.dump(g, fh, separators=(",", ":"))
    print("wrote", os.path.join(HERE, "util_fhew.json"), os.path.getsize(os.path.join(HERE, "util_fhew.json")), "bytes")


def main_tfhe_ckks():
    """tests/golden/tfhe_ckks.json: the f64 FFT torus product, a TGGSW external product / CMUX, a TLWE key switch, rescale_k and a
    whole Ckks::mul, all computed by pyref alone (no oracle, no CUDA) on seeded inputs."""
    g = {}
    M = 1 << 64
    ff = []
    for n, log_b in ((2, 8), (8, 8), (16, 23), (64, 12)):
        a = rnd(300 + n, n, M)
        b = [(x - (1 << (log_b - 1))) % M for x in rnd(400 + n, n, 1 << log_b)]
        ff.append({"a": a, "b": b, "out": pyref.fft64_negacyclic_mul(a, b)})
    g["fft64_mul"] = ff
    k, d, log_b, n = 1, 2, 8, 16
    rows = [[rnd(500 + 10 * r + c, n, M) for c in range(k + 1)] for r in range((k + 1) * d)]
    ct0 = [rnd(600 + c, n, M) for c in range(k + 1)]
    ct1 = [rnd(610 + c, n, M) for c in range(k + 1)]
    g["tggsw"] = {"k": k, "d": d, "log_b": log_b, "n": n, "rows": rows, "ct0": ct0, "ct1": ct1,
                  "external_product": pyref.tggsw_external_product(log_b, d, rows, ct0), "cmux": pyref.tggsw_cmux(log_b, d, rows, ct0, ct1)}
    kn, n_out, ks_log_b, ks_d = 16, 5, 4, 5
    ksk_a = [rnd(700 + i, n_out, M) for i in range(kn * ks_d)]
    ksk_b = rnd(799, kn * ks_d, M)
    a, b = rnd(800, kn, M), rnd(801, 1, M)[0]
    oa, ob = pyref.tlwe_key_switch(ks_log_b, ks_d, ksk_a, ksk_b, a, b)
    g["tlwe_key_switch"] = {"log_b": ks_log_b, "d": ks_d, "ksk_a": ksk_a, "ksk_b": ksk_b, "a": a, "b": b, "out": oa + [ob]}
    # a whole programmable bootstrap at tiny parameters: n = 4, N = 16, k = 1, TGGSW (8, 2), key switch (4, 5), p = 2^4 + padding
    n_lwe, nb, kk, bl, bd = 4, 16, 1, 8, 2
    brk = [[[rnd(1100 + 100 * i + 10 * r + c, nb, M) for c in range(kk + 1)] for r in range((kk + 1) * bd)] for i in range(n_lwe)]
    pk_a = [rnd(1500 + i, n_lwe, M) for i in range(kk * nb * 5)]
    pk_b = rnd(1599, kk * nb * 5, M)
    v = [i % 16 for i in range(nb)]
    cts = [rnd(1600 + i, n_lwe + 1, M) for i in range(3)]
    cts[1][0] = 0
    outs = []
    for ct in cts:
        oa, ob = pyref.tfhe_bootstrap(4, 1, kk, bl, bd, 4, 5, brk, pk_a, pk_b, v, ct)
        outs.append(oa + [ob])
    g["tfhe_pbs"] = {"log_p": 4, "padding": 1, "n": n_lwe, "big_n": nb, "k": kk, "bs_log_b": bl, "bs_d": bd, "ks_log_b": 4, "ks_d": 5,
                     "brk": brk, "ksk_a": pk_a, "ksk_b": pk_b, "v": v, "cts": cts, "out": outs}
    primes = pyref.two_adic_primes(55, 4, 6)
    rs = []
    for nq, kk in ((4, 1), (6, 3), (5, 2)):
        qs = primes[:nq]
        x = [[rnd(900 + 7 * nq + i, 1, q)[0] for i, q in enumerate(qs)] for _ in range(6)] + [[q - 1 for q in qs], [0] * nq]
        rs.append({"qs": qs, "k": kk, "x": x, "out": [pyref.rns_rescale_k(qs, kk, c) for c in x]})
    g["rns_rescale_k"] = rs
    n, big_l = 8, 3
    qs, ps = primes[:big_l], primes[big_l:2 * big_l]
    mods = qs + ps
    rlk_b = [rnd(1000 + i, n, q) for i, q in enumerate(mods)]
    rlk_a = [rnd(1010 + i, n, q) for i, q in enumerate(mods)]
    ct0 = ([rnd(1020 + i, n, q) for i, q in enumerate(qs)], [rnd(1030 + i, n, q) for i, q in enumerate(qs)])
    ct1 = ([rnd(1040 + i, n, q) for i, q in enumerate(qs)], [rnd(1050 + i, n, q) for i, q in enumerate(qs)])
    pb, pa = pyref.ckks_mul(qs, ps, rlk_b, rlk_a, ct0, ct1)
    sb, sa = pyref.ckks_key_switch(qs, ps, rlk_b, rlk_a, ct0[0], ct0[1])
    # rotation by t = 5, plaintext product, and a 2 x 2 BSGS matrix product (one absent diagonal) with two more keys
    keys = [([rnd(1200 + 20 * w + i, n, q) for i, q in enumerate(mods)], [rnd(1210 + 20 * w + i, n, q) for i, q in enumerate(mods)]) for w in range(2)]
    rot5 = pyref.ckks_rotate(qs, ps, 5, keys[0][0], keys[0][1], ct0[0], ct0[1])
    pts = [[rnd(1300 + 10 * j + i, n, q) for i, q in enumerate(qs)] for j in range(3)]
    mc = pyref.ckks_mul_constant(qs, pts[0], ct0[0], ct0[1])
    baby, giant, present = [(0, None), (5, keys[0])], [(0, None), (2 * n - 1, keys[1])], [[1, 1], [0, 1]]
    mm = pyref.ckks_mul_mat(qs, ps, baby, giant, present, pts, ct0[0], ct0[1])
    g["ckks"] = {"log_n": 3, "qs": qs, "ps": ps, "ksk": [rlk_b, rlk_a], "ct0": [ct0[0], ct0[1]], "ct1": [ct1[0], ct1[1]],
                 "mul": [pb, pa], "key_switch_ct0": [sb, sa], "rot_keys": [[k[0], k[1]] for k in keys], "rotate5_ct0": [rot5[0], rot5[1]],
                 "pts": pts, "mul_constant_pt0_ct0": [mc[0], mc[1]], "mul_mat": {"baby_t": [0, 5], "giant_t": [0, 2 * n - 1], "present": present,
                                                                              "out": [mm[0], mm[1]]}}
    path = os.path.join(HERE, "tfhe_ckks.json")
    with open(path, "w") as fh:
        json.dump(g, fh, separators=(",", ":"))
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
    main_tfhe_ckks()
