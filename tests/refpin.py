"""Replays known-answer files written by the REFERENCE itself (oracle/pin/*.rs run with cargo in a checkout of
han0110/learn-fhe, see oracle/pin/README.md) against the C++ oracle and, under -m gpu, against the CUDA library.
One set of check functions serves both the real files (tests/golden/ref/ref_*.json, present only after somebody with a
Rust toolchain has run oracle/pin/apply.sh) and a self-generated set in the same schema (oracle/pin/selfcheck.py, which
keeps this code exercised in every CPU run but pins nothing)."""
import json
import os

import numpy as np

A = lambda v: np.ascontiguousarray(v, dtype=np.uint64)
HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "golden", "ref")
FILES = ("ref_util.json", "ref_fhew.json", "ref_tfhe.json", "ref_ckks.json")


def load(dirname, name):
    p = os.path.join(dirname, name)
    return json.load(open(p)) if os.path.exists(p) else None


# ---- oracle side --------------------------------------------------------------------------------------------------------------
def check_util(orc, G):
    n = 0
    for c in G.get("ntt", []):
        assert (orc.ntt_fwd(c["q"], A(c["a"])) == A(c["fwd"])).all(), ("ntt", c["q"], len(c["a"]))
        assert (orc.ntt_inv(c["q"], A(c["fwd"])) == A(c["a"])).all()
        n += 1
    if "negacyclic_mul" in G:
        c = G["negacyclic_mul"]
        assert (orc.ntt_mul(c["q"], A(c["a"]), A(c["b"])) == A(c["out"])).all()
        n += 1
    for c in G.get("decompose_zq", []):
        assert (orc.decompose_zq(c["q"], c["log_b"], c["d"], A(c["v"])) == A(c["digits"]).T).all(), ("decompose_zq", c["q"], c["log_b"], c["d"])
        n += 1
    for c in G.get("decompose_t64", []):
        assert (orc.decompose_t64(c["log_b"], c["d"], A(c["v"])) == A(c["digits"]).T).all(), ("decompose_t64", c["log_b"], c["d"])
        n += 1
    for c in G.get("mod_switch", []):
        assert (orc.mod_switch(c["q"], c["qp"], A(c["v"])) == A(c["mod_switch"])).all()
        assert (orc.mod_switch(c["q"], c["qp"], A(c["v"]), odd=True) == A(c["mod_switch_odd"])).all()
        n += 1
    for c in G.get("automorphism", []):
        assert (orc.automorphism_zq(c["q"], A(c["a"]), c["t"]) == A(c["out"])).all()
        n += 1
    for c in G.get("monomial_mul", []):
        assert (orc.monomial_mul_zq(c["q"], A(c["a"]), c["k"]) == A(c["out"])).all()
        n += 1
    for c in G.get("fft64_mul", []):
        assert [int(x) for x in orc.fft64_mul(A(c["a"]), A(c["b"]))] == c["out"], ("fft64_mul", len(c["a"]))
        n += 1
    r = G.get("rns_extend_bases")
    if r:
        for c in r["cases"]:
            out = orc.rns_extend_bases(r["qs"], r["ps"], A(c["x"]).reshape(-1, 1))
            assert [int(x) for x in out[len(r["qs"]):, 0]] == c["out"] or [int(x) for x in out[:, 0]] == c["out"]
            n += 1
    for c in G.get("rns_rescale_k", []):
        x = A(c["x"]).T  # [limb][coefficient]
        got = orc.rns_rescale_k(c["qs"], c["k"], x)
        assert [[int(v) for v in got[:, i]] for i in range(x.shape[1])] == c["out"], ("rescale_k", c["k"])
        n += 1
    return n


def fhew_key(orc, g):
    P = orc.FhewParamC(**{k: g["param"][k] for k in ("log_n", "big_q", "p", "rlwe_log_b", "rlwe_d", "rgsw_log_b", "rgsw_d", "n_s", "q_ks",
                                                     "ks_log_b", "ks_d", "w")})
    k = g["keys"]
    return P, (A(k["ksk_a"]), A(k["ksk_b"]), A(k["brk"]), A(k["ak"]), np.array(k["ak_t"], dtype=np.int64))


def check_fhew(orc, G):
    n = 0
    for name, g in G.items():
        P, keys = fhew_key(orc, g)
        K = orc.FhewKey.from_arrays(P, *keys)
        f = A(g["f"])
        assert (orc.fhew_gate_poly(P, g["table"]) == f).all(), name
        for c in g["cases"]:
            ct = A(c["ct"]).reshape(1, -1)
            assert [int(x) for x in K.prologue(ct)[0]] == c["prologue"], (name, "prologue")
            out = K.bootstrap(f, ct)
            out[0, -1] = (int(out[0, -1]) + g["post_add"]) % P.big_q
            assert [int(x) for x in out[0]] == c["out"], (name, "bootstrap")
            assert [int(x) for x in K.op(g["table"], ct)[0]] == c["out"]
            n += 1
        for s in g.get("steps", []):
            fn = K.external_product if s["kind"] == "external_product" else K.automorphism
            assert fn(s["idx"], A(s["acc"])).tolist() == s["out"], (name, s["kind"], s["idx"])
            n += 1
    return n


def tfhe_param(orc, g, n=None):
    P = orc.tfhe_testing_param()
    P.log_p, P.padding = g.get("log_p", 4), g.get("padding", 1)
    P.n, P.big_n, P.k = n if n is not None else g["n"], g["big_n"], g["k"]
    P.bs_log_b, P.bs_d, P.ks_log_b, P.ks_d = g["bs_log_b"], g["bs_d"], g["ks_log_b"], g["ks_d"]
    return P


def check_tfhe(orc, G):
    n = 0
    for g in G.get("tfhe_pbs", []):
        P = tfhe_param(orc, g)
        K = orc.TfheKey.from_arrays(P, A(g["brk"]), A(g["ksk_a"]), A(g["ksk_b"]))
        got = K.bootstrap(A(g["v"]), A(g["cts"]))
        assert got.tolist() == g["out"], ("tfhe_pbs", g["big_n"], g["bs_log_b"], g["bs_d"])
        n += 1
    for g in G.get("tggsw", []):
        rows = A(g["rows"])
        P = tfhe_param(orc, dict(big_n=g["n"], k=g["k"], bs_log_b=g["log_b"], bs_d=g["d"], ks_log_b=4, ks_d=1), n=1)
        kn = g["k"] * g["n"]
        K = orc.TfheKey.from_arrays(P, rows[None], np.zeros((kn, 1), dtype=np.uint64), np.zeros(kn, dtype=np.uint64))
        ct0, ct1 = A(g["ct0"]), A(g["ct1"])
        assert K.external_product(0, ct0).tolist() == g["external_product"], ("tggsw", g["n"], g["k"], g["d"])
        assert (ct0 + K.external_product(0, ct1 - ct0)).tolist() == g["cmux"]
        n += 1
    for g in G.get("tlwe_key_switch", []):
        ksk_a = A(g["ksk_a"])
        rows, n_out = ksk_a.shape
        kn = rows // g["d"]
        P = tfhe_param(orc, dict(big_n=kn, k=1, bs_log_b=8, bs_d=1, ks_log_b=g["log_b"], ks_d=g["d"]), n=n_out)
        K = orc.TfheKey.from_arrays(P, np.zeros((n_out, 2, 2, kn), dtype=np.uint64), ksk_a, A(g["ksk_b"]))
        assert [int(x) for x in K.key_switch(A(g["a"] + [g["b"]]))] == g["out"], ("tlwe_key_switch", kn)
        n += 1
    return n


def check_ckks(orc, G):
    n = 0
    for g in G.get("ckks", []):
        auto = [(k["t"], A(k["ksk"])) for k in g.get("rot_keys", [])]
        K = orc.CkksKey.from_arrays(g["log_n"], g["qs"], g["ps"], A(g["ksk"]), auto)
        ct0, ct1 = A(g["ct0"]), A(g["ct1"])
        mul = K.mul(ct0, ct1)
        assert mul.tolist() == g["mul"], ("ckks mul", g["log_n"])
        if "mul_again" in g:
            assert K.mul(mul, mul).tolist() == g["mul_again"]
        assert K.key_switch(-1, ct0).tolist() == g["key_switch_ct0"]
        if auto:
            assert K.key_switch(0, ct0, apply_auto=True).tolist() == g["rotate1_ct0"]
        n += 1
    return n


# ---- CUDA side (through the C ABI) --------------------------------------------------------------------------------------------------
def check_gpu(pkg, ctx, orc, dirname):
    from learn_fhe_b200 import ckks, fhew, tfhe, util
    n = 0
    G = load(dirname, "ref_util.json") or {}
    for c in G.get("ntt", []):
        x = A(c["a"]).copy()[None]
        util.nega_cyclic_ntt_in_place(ctx, c["q"], x)
        assert (x[0] == A(c["fwd"])).all()
        n += 1
    for c in G.get("fft64_mul", []):
        assert [int(v) for v in tfhe.nega_cyclic_fft64_mul_assign_rt(ctx, A(c["a"]).copy(), A(c["b"]))] == c["out"]
        n += 1
    for c in G.get("rns_rescale_k", []):
        x = A(c["x"]).T.copy()[None]
        got = ckks.rescale_k(ctx, c["qs"], c["k"], x)
        assert [[int(v) for v in got[0][:, i]] for i in range(x.shape[2])] == c["out"]
        n += 1
    for name, g in (load(dirname, "ref_fhew.json") or {}).items():
        P, keys = fhew_key(orc, g)
        param = pkg.FhewParam(**{k: g["param"][k] for k in ("log_n", "big_q", "p", "rlwe_log_b", "rlwe_d", "rgsw_log_b", "rgsw_d", "n_s", "q_ks",
                                                            "ks_log_b", "ks_d", "w")})
        bk = fhew.BootstrappingKey(ctx, param, *keys)
        cts = A([c["ct"] for c in g["cases"]])
        got = fhew.Bootstrapping.bootstrap(bk, A(g["f"]), cts, post_add=g["post_add"])
        assert got.tolist() == [c["out"] for c in g["cases"]], name
        bk.free()
        n += 1
    for g in (load(dirname, "ref_tfhe.json") or {}).get("tfhe_pbs", []):
        param = pkg.TfheParam(log_p=g["log_p"], padding=g["padding"], n=g["n"], ks_log_b=g["ks_log_b"], ks_d=g["ks_d"],
                              log_big_n=g["big_n"].bit_length() - 1, k=g["k"], bs_log_b=g["bs_log_b"], bs_d=g["bs_d"])
        bk = tfhe.BootstrappingKey(ctx, param, A(g["brk"]), A(g["ksk_a"]), A(g["ksk_b"]))
        got = tfhe.Bootstrapping.bootstrap(bk, tfhe.encode_lut(bk.param, A(g["v"])), A(g["cts"]))
        assert got.tolist() == g["out"], ("tfhe_pbs", g["big_n"])
        bk.free()
        n += 1
    for g in (load(dirname, "ref_ckks.json") or {}).get("ckks", []):
        P = ckks.CkksParam(ctx, g["log_n"], g["qs"], g["ps"])
        rlk = ckks.CkksKeySwitchingKey(P, A(g["ksk"]))
        ct0, ct1 = A(g["ct0"])[None], A(g["ct1"])[None]
        assert ckks.Ckks.mul(P, rlk, ct0, ct1)[0].tolist() == g["mul"], ("ckks mul", g["log_n"])
        assert ckks.Ckks.key_switch(P, rlk, ct0, 0)[0].tolist() == g["key_switch_ct0"]
        for k in g.get("rot_keys", []):
            rk = ckks.CkksKeySwitchingKey(P, A(k["ksk"]))
            assert ckks.Ckks.key_switch(P, rk, ct0, k["t"])[0].tolist() == g["rotate1_ct0"]
            rk.free()
        rlk.free()
        P.free()
        n += 1
    return n
