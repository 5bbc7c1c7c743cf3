#!/usr/bin/env python
"""TEST INFRASTRUCTURE.  Writes files in the schema of the Rust dumpers (oracle/pin/*.rs) from the C++ ORACLE, so that the
replay code of tests/refpin.py is exercised in every CPU test run.  These files pin NOTHING (oracle against oracle); the
pin is tests/golden/ref/, which only the reference's own binaries can produce (oracle/pin/apply.sh).
usage: python oracle/pin/selfcheck.py OUTDIR"""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import orc  # noqa: E402

L = lambda a: np.asarray(a).tolist()


def util_section():
    g = {"ntt": [], "decompose_zq": [], "decompose_t64": [], "mod_switch": [], "automorphism": [], "monomial_mul": [], "fft64_mul": [],
         "rns_rescale_k": []}
    for bits, log_n in ((28, 3), (28, 9), (45, 5), (55, 7)):
        q = orc.two_adic_primes(bits, log_n + 1, 1)[0]
        a = orc.residues(100 + log_n, 1 << log_n, q)
        g["ntt"].append({"q": q, "a": L(a), "fwd": L(orc.ntt_fwd(q, a)), "generator": 0})
    q = orc.two_adic_primes(45, 5, 1)[0]
    a, b = orc.residues(1, 16, q), orc.residues(2, 16, q)
    g["negacyclic_mul"] = {"q": q, "a": L(a), "b": L(b), "out": L(orc.ntt_mul(q, a, b))}
    for q, log_b, d in ((268409857, 7, 4), (1 << 16, 4, 4), (97, 2, 3)):
        v = orc.residues(7, 24, q)
        g["decompose_zq"].append({"q": q, "log_b": log_b, "d": d, "v": L(v), "digits": L(orc.decompose_zq(q, log_b, d, v).T)})
    for log_b, d in ((23, 1), (4, 5), (7, 3)):
        v = orc.splitmix64(8, 24)
        g["decompose_t64"].append({"log_b": log_b, "d": d, "v": L(v), "digits": L(orc.decompose_t64(log_b, d, v).T)})
    for q, qp in ((268409857, 1 << 16), (1 << 16, 1024)):
        v = orc.residues(9, 40, q)
        g["mod_switch"].append({"q": q, "qp": qp, "v": L(v), "mod_switch": L(orc.mod_switch(q, qp, v)),
                                "mod_switch_odd": L(orc.mod_switch(q, qp, v, odd=True))})
    q = 268409857
    for t in (5, -5, 25):
        a = orc.residues(20 + t, 16, q)
        g["automorphism"].append({"q": q, "a": L(a), "t": t, "out": L(orc.automorphism_zq(q, a, t))})
    for k in (0, 1, 17, -1):
        a = orc.residues(40 + k, 16, q)
        g["monomial_mul"].append({"q": q, "a": L(a), "k": k, "out": L(orc.monomial_mul_zq(q, a, k))})
    for log_n, log_b in ((1, 8), (6, 17), (11, 23)):
        n = 1 << log_n
        a = orc.splitmix64(300 + log_n, n)
        b = ((orc.splitmix64(400 + log_n, n) % np.uint64(1 << log_b)).astype(np.int64) - (1 << (log_b - 1))).astype(np.uint64)
        g["fft64_mul"].append({"a": L(a), "b": L(b), "out": L(orc.fft64_mul(a, b))})
    pr = orc.two_adic_primes(55, 8, 6)
    qs, ps = pr[:3], pr[3:]
    cases = []
    for i in range(6):
        x = np.array([orc.residues(500 + i, 1, m)[0] for m in qs], dtype=np.uint64)
        cases.append({"x": L(x), "out": L(orc.rns_extend_bases(qs, ps, x.reshape(-1, 1))[3:, 0])})
    g["rns_extend_bases"] = {"qs": qs, "ps": ps, "cases": cases}
    for nq, k in ((3, 1), (6, 3)):
        m = pr[:nq]
        x = np.stack([orc.residues(600 + i, 8, q) for i, q in enumerate(m)])
        g["rns_rescale_k"].append({"qs": m, "k": k, "x": L(x.T), "out": L(orc.rns_rescale_k(m, k, x).T)})
    return g


def fhew_section(log_q, log_n, log_b, d, n_s, log_q_ks, ks, w, seed):
    P = orc.fhew_testing_param()
    P.log_n, P.big_q, P.p = log_n, orc.two_adic_primes(log_q, log_n + 1, 1)[0], 4
    P.rlwe_log_b = P.rgsw_log_b = log_b
    P.rlwe_d = P.rgsw_d = d
    P.n_s, P.q_ks, P.ks_log_b, P.ks_d, P.w = n_s, 1 << log_q_ks, ks[0], ks[1], w
    K = orc.FhewKey(P, seed)
    ex = K.export()
    f = orc.fhew_gate_poly(P, [1, 1, 1, 0])
    q8 = int(round(P.big_q / 8.0))
    bits = np.array([0, 0, 1, 1, 0, 1, 0, 1], dtype=np.int32)
    cts = K.encrypt(bits, 7)
    lin = (cts[:4] + cts[4:]) % np.uint64(P.big_q)
    cases = []
    for i in range(4):
        out = K.op([1, 1, 1, 0], lin[i:i + 1])[0]
        cases.append({"ct": L(lin[i]), "prologue": L(K.prologue(lin[i:i + 1])[0]), "out": L(out), "bit": int(K.decrypt(out[None])[0])})
    steps = []
    n = 1 << log_n
    for idx in (0, n_s - 1):
        acc = orc.residues(900 + idx, 2 * n, P.big_q).reshape(2, n)
        steps.append({"kind": "external_product", "idx": idx, "acc": L(acc), "out": L(K.external_product(idx, acc))})
    for idx in (0, w):
        acc = orc.residues(950 + idx, 2 * n, P.big_q).reshape(2, n)
        steps.append({"kind": "automorphism", "idx": idx, "acc": L(acc), "out": L(K.automorphism(idx, acc))})
    param = {k: int(getattr(P, k)) for k in ("log_n", "big_q", "p", "rlwe_log_b", "rlwe_d", "rgsw_log_b", "rgsw_d", "n_s", "q_ks", "ks_log_b", "ks_d", "w")}
    param["n"] = n
    return {"param": param, "keys": {k: L(ex[k]) for k in ("ksk_a", "ksk_b", "brk", "ak", "ak_t")}, "table": [1, 1, 1, 0], "f": L(f), "post_add": q8,
            "cases": cases, "steps": steps}


def tfhe_section():
    out = {"tfhe_pbs": [], "tggsw": [], "tlwe_key_switch": []}
    for n, big_n, k, bs, ks, seed in ((4, 16, 1, (8, 2), (4, 5), 31), (3, 64, 2, (8, 3), (4, 5), 32)):
        P = orc.tfhe_testing_param()
        P.n, P.big_n, P.k, P.bs_log_b, P.bs_d, P.ks_log_b, P.ks_d = n, big_n, k, bs[0], bs[1], ks[0], ks[1]
        K = orc.TfheKey(P, seed)
        ex = K.export()
        p = 1 << P.log_p
        table = ((3 * np.arange(p) + 1) % p).astype(np.uint64)
        v = K.lut_poly(table)
        cts = K.encrypt(np.array([0, 1, 7, 15], dtype=np.uint64), 5)
        out["tfhe_pbs"].append({"log_p": P.log_p, "padding": P.padding, "n": n, "big_n": big_n, "k": k, "bs_log_b": bs[0], "bs_d": bs[1],
                                "ks_log_b": ks[0], "ks_d": ks[1], "brk": L(ex["brk"]), "ksk_a": L(ex["ksk_a"]), "ksk_b": L(ex["ksk_b"]),
                                "v": L(v), "cts": L(cts), "out": L(K.bootstrap(v, cts))})
        ct0 = orc.splitmix64(60 + seed, (k + 1) * big_n).reshape(k + 1, big_n)
        ct1 = orc.splitmix64(70 + seed, (k + 1) * big_n).reshape(k + 1, big_n)
        out["tggsw"].append({"k": k, "d": bs[1], "log_b": bs[0], "n": big_n, "rows": L(ex["brk"][0]), "ct0": L(ct0), "ct1": L(ct1),
                             "external_product": L(K.external_product(0, ct0)), "cmux": L(ct0 + K.external_product(0, ct1 - ct0))})
        big = orc.splitmix64(80 + seed, k * big_n + 1)
        out["tlwe_key_switch"].append({"log_b": ks[0], "d": ks[1], "ksk_a": L(ex["ksk_a"]), "ksk_b": L(ex["ksk_b"]), "a": L(big[:-1]), "b": int(big[-1]),
                                       "out": L(K.key_switch(big))})
    return out


def ckks_section():
    cases = []
    for log_n, log_qi, big_l, seed in ((3, 55, 3, 41), (6, 55, 4, 42)):
        n = 1 << log_n
        t = pow(5, 1, 2 * n)
        K = orc.CkksKey(log_n, log_qi, big_l, seed, auto_ts=(t,))
        ct0 = K.encrypt((np.arange(n, dtype=np.int64) * 3) % 17 - 8, big_l, 7)
        ct1 = K.encrypt((np.arange(n, dtype=np.int64) * 5) % 5 - 2, big_l, 8)
        mul = K.mul(ct0, ct1)
        cases.append({"log_n": log_n, "qs": K.qs, "ps": K.ps, "sk": L(K.sk()), "ksk": L(K.ksk(-1)), "ct0": L(ct0), "ct1": L(ct1), "mul": L(mul),
                      "mul_again": L(K.mul(mul, mul)), "key_switch_ct0": L(K.key_switch(-1, ct0)),
                      "rot_keys": [{"j": 1, "t": t, "ksk": L(K.ksk(0))}], "rotate1_ct0": L(K.key_switch(0, ct0, apply_auto=True))})
    return {"ckks": cases}


def main(out):
    orc.build()
    orc.lib()
    os.makedirs(out, exist_ok=True)
    dump = lambda name, obj: json.dump(obj, open(os.path.join(out, name), "w"), separators=(",", ":"))
    dump("ref_util.json", util_section())
    dump("ref_fhew.json", {"fhew_tiny": fhew_section(20, 4, 5, 4, 6, 10, (2, 5), 3, 0x201), "fhew_n64": fhew_section(28, 6, 7, 4, 12, 16, (4, 4), 10, 0x202)})
    dump("ref_tfhe.json", tfhe_section())
    dump("ref_ckks.json", ckks_section())


if __name__ == "__main__":
    main(sys.argv[1])
