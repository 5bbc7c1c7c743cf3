"""CPU tier: the C++ oracle (oracle/liborc.so) against (a) the reference's own property tests restated with the same
sweeps, (b) the independent pure-Python restatement tests/pyref.py, (c) the committed golden fixtures."""
import json
import os

import numpy as np
import pytest

import pyref

GOLD = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "util_fhew.json")))


def A(x):
    return np.array(x, dtype=np.uint64)


# ---- (a) reference property tests ------------------------------------------------------------------------------------
def test_ntt_round_trip_and_schoolbook(orc):
    """util/src/ring/fft/zq.rs:94-116, ring.rs:442-452: log_n 0..9, ten 45-bit primes each."""
    for log_n in range(0, 10):
        n = 1 << log_n
        for k, q in enumerate(orc.two_adic_primes(45, log_n + 1, 10)):
            a, b = orc.residues(3 * log_n + k, n, q), orc.residues(5 * log_n + k + 1, n, q)
            assert (orc.ntt_inv(q, orc.ntt_fwd(q, a)) == a).all()
            if log_n <= 7:
                assert (orc.ntt_mul(q, a, b) == orc.schoolbook_zq(q, a, b)).all()


def test_ntt_evaluation_order(orc):
    """out[i] = a(psi^(2*brev(i)+1)), psi = omega^(2^s / 2n) — checked by direct evaluation."""
    log_n, n = 4, 16
    q = orc.two_adic_primes(28, 10, 1)[0]
    a = [int(x) for x in orc.residues(1, n, q)]
    s = ((q - 1) & -(q - 1)).bit_length() - 1
    psi = pow(pyref.zq_two_adic_generator(q, s), (1 << s) // (2 * n), q)
    out = orc.ntt_fwd(q, A(a))
    for i in range(n):
        br = int(format(i, "04b")[::-1], 2)
        x = pow(psi, 2 * br + 1, q)
        assert int(out[i]) == sum(c * pow(x, e, q) for e, c in enumerate(a)) % q


def test_decomposer_recomposition(orc):
    """Digits recompose to the rounded value up to the documented wrap (Zq: error <= 2^log_q - q)."""
    for q, log_b, d in ((268409857, 7, 4), (1 << 16, 4, 4), (orc.two_adic_primes(55, 12, 1)[0], 11, 5)):
        v = orc.residues(5, 2000, q)
        digs = orc.decompose_zq(q, log_b, d, v)
        lq, rb, bases = orc.decomposor_zq_info(q, log_b, d)
        centred = np.where(digs < (q >> 1), digs.astype(np.int64), digs.astype(np.int64) - np.int64(q))
        assert centred.max() <= (1 << (log_b - 1)) and centred.min() >= -(1 << (log_b - 1)) + 1
        rec = sum(int(bases[k]) * centred[k].astype(object) for k in range(d))
        err = np.array([min((int(r) - int(x)) % q, (int(x) - int(r)) % q) for r, x in zip(rec, v)], dtype=object)
        assert max(err) <= (1 << max(rb, 1)) + ((1 << lq) - q)
    for log_b, d in ((23, 1), (4, 5), (8, 8)):
        v = orc.splitmix64(6, 2000)
        digs = orc.decompose_t64(log_b, d, v).astype(np.int64)
        assert digs.max() <= (1 << (log_b - 1)) and digs.min() >= -(1 << (log_b - 1))
        rb = 64 - log_b * d
        rec = sum((digs[k].astype(object) << (rb + k * log_b)) for k in range(d))
        err = [min((int(r) - int(x)) % (1 << 64), (int(x) - int(r)) % (1 << 64)) for r, x in zip(rec, v)]
        assert max(err) <= (1 << rb) >> 1


def test_rns_extend_preserves_value(orc):
    """util/src/ring/rns.rs:373-386: 8 -> 16 primes of 55 bits, centred value preserved."""
    import math
    primes = orc.two_adic_primes(55, 5, 16)
    qs, ps = primes[:8], primes[8:]
    big_q = math.prod(qs)
    rng = np.random.default_rng(1)
    n = 16
    vals = [int.from_bytes(rng.bytes(55), "little") % big_q - big_q // 2 for _ in range(n)]
    x = A([[v % qi for v in vals] for qi in qs])
    out = orc.rns_extend_bases(qs, ps, x)
    for j, p in enumerate(ps):
        assert [int(t) for t in out[8 + j]] == [v % p for v in vals]
    assert (out[:8] == x).all()


def test_fhew_gates_decrypt(orc, fhew_setup):
    """fhew/boolean.rs:257-286 at single_key_testing_param: NAND / AND / XOR truth tables through the oracle."""
    P, K, _ = fhew_setup
    a = np.array([0, 0, 1, 1], dtype=np.int32)
    b = np.array([0, 1, 0, 1], dtype=np.int32)
    ca, cb = K.encrypt(a, 100), K.encrypt(b, 200)
    q = np.uint64(P.big_q)
    lin = (ca + cb) % q
    assert (K.decrypt(K.op([1, 1, 1, 0], lin, threads=4)) == 1 - (a & b)).all()
    assert (K.decrypt(K.op([0, 0, 0, 1], lin, threads=4)) == (a & b)).all()
    d = (ca + (q - cb)) % q
    assert (K.decrypt(K.op([0, 1, 1, 1], (d + d) % q, threads=4)) == (a ^ b)).all()


def test_fhew_step_census(orc, fhew_setup):
    """~100 external products and 101..117 automorphisms per bootstrap at FHEW-T."""
    P, K, _ = fhew_setup
    cts = K.encrypt(np.array([0, 1, 1, 0, 1, 0, 0, 1], dtype=np.int32), 9)
    for row in K.prologue(cts):
        st = orc.fhew_schedule(P, row[:P.n_s])
        n_ext, n_auto = int((st[:, 0] == 0).sum()), int((st[:, 0] == 1).sum())
        assert 95 <= n_ext <= 100 and 98 <= n_auto <= 120
        assert (row[:P.n_s] % 2 == 1).sum() + (row[:P.n_s] == 0).sum() == P.n_s  # odd or zero (zq.rs:132-140)


# ---- (b) oracle vs the independent Python restatement ------------------------------------------------------------------
def test_oracle_matches_pyref_util(orc):
    for bits, log_n in ((28, 5), (45, 6), (55, 4), (61, 3)):
        q = orc.two_adic_primes(bits, log_n + 1, 1)[0]
        assert q == pyref.two_adic_primes(bits, log_n + 1, 1)[0]
        a = orc.residues(log_n, 1 << log_n, q)
        fwd = pyref.ntt_fwd(q, [int(x) for x in a])
        assert [int(x) for x in orc.ntt_fwd(q, a)] == fwd
        assert [int(x) for x in orc.ntt_inv(q, A(fwd))] == [int(x) for x in a]
        f, i = orc.twiddles(q)
        pf, pi = pyref.compute_twiddle(q)
        assert [int(x) for x in f[:64]] == pf[:64] and [int(x) for x in i[:64]] == pi[:64] and len(f) == len(pf)
    q = 268409857
    v = orc.residues(2, 300, q)
    for log_b, d in ((7, 4), (5, 4), (9, 3), (14, 2)):
        got = orc.decompose_zq(q, log_b, d, v)
        exp = np.array([pyref.decompose_zq(q, log_b, d, int(x)) for x in v], dtype=np.uint64).T
        assert (got == exp).all(), (log_b, d)
    w = orc.splitmix64(3, 300)
    for log_b, d in ((23, 1), (4, 5), (7, 3), (1, 3)):
        got = orc.decompose_t64(log_b, d, w)
        exp = np.array([pyref.decompose_t64(log_b, d, int(x)) for x in w], dtype=np.uint64).T
        assert (got == exp).all(), (log_b, d)
    for qq, qp in ((268409857, 1 << 16), (1 << 16, 1024)):
        v = orc.residues(4, 500, qq)
        assert [int(x) for x in orc.mod_switch(qq, qp, v)] == [pyref.zq_mod_switch(qq, int(x), qp) for x in v]
        assert [int(x) for x in orc.mod_switch(qq, qp, v, odd=True)] == [pyref.zq_mod_switch_odd(qq, int(x), qp) for x in v]
    a = orc.residues(5, 32, q)
    for t in (5, -5, 63, 7):
        assert [int(x) for x in orc.automorphism_zq(q, a, t)] == pyref.automorphism([int(x) for x in a], t, q)
    for k in (0, 1, 31, 32, 33, 63, -1, -40, 100):
        assert [int(x) for x in orc.monomial_mul_zq(q, a, k)] == pyref.monomial_mul([int(x) for x in a], k, q)


# ---- (c) golden fixtures -----------------------------------------------------------------------------------------------
def test_golden_util(orc):
    for c in GOLD["ntt"]:
        assert (orc.ntt_fwd(c["q"], A(c["a"])) == A(c["fwd"])).all()
        assert (orc.ntt_inv(c["q"], A(c["fwd"])) == A(c["a"])).all()
    m = GOLD["fhew_t_modulus"]
    f, i = orc.twiddles(m["q"])
    assert len(f) == m["tw_len"] and [int(x) for x in f[:16]] == m["tw_first16"] and [int(x) for x in i[:16]] == m["tw_inv_first16"]
    assert m["q"] == 268409857 and m["generator"] == 5 and m["omega"] == 28892341
    c = GOLD["negacyclic_mul"]
    assert (orc.ntt_mul(c["q"], A(c["a"]), A(c["b"])) == A(c["out"])).all()
    for c in GOLD["decompose_zq"]:
        assert (orc.decompose_zq(c["q"], c["log_b"], c["d"], A(c["v"])) == A(c["digits"]).T).all()
    for c in GOLD["decompose_t64"]:
        assert (orc.decompose_t64(c["log_b"], c["d"], A(c["v"])) == A(c["digits"]).T).all()
    for c in GOLD["mod_switch"]:
        assert (orc.mod_switch(c["q"], c["qp"], A(c["v"])) == A(c["mod_switch"])).all()
        assert (orc.mod_switch(c["q"], c["qp"], A(c["v"]), odd=True) == A(c["mod_switch_odd"])).all()
    for c in GOLD["automorphism"]:
        assert (orc.automorphism_zq(c["q"], A(c["a"]), c["t"]) == A(c["out"])).all()
    for c in GOLD["monomial_mul"]:
        assert (orc.monomial_mul_zq(c["q"], A(c["a"]), c["k"]) == A(c["out"])).all()
    r = GOLD["rns_extend_bases"]
    for c in r["cases"]:
        out = orc.rns_extend_bases(r["qs"], r["ps"], A(c["x"]).reshape(-1, 1))
        assert [int(x) for x in out[:, 0]] == c["out"]


def golden_fhew_tiny(orc):
    g = GOLD["fhew_tiny"]
    P = orc.FhewParamC(**{k: g["param"][k] for k in ("log_n", "big_q", "p", "rlwe_log_b", "rlwe_d", "rgsw_log_b", "rgsw_d", "n_s",
                                                     "q_ks", "ks_log_b", "ks_d", "w")})
    k = g["keys"]
    return g, P, (A(k["ksk_a"]), A(k["ksk_b"]), A(k["brk"]), A(k["ak"]), np.array(k["ak_t"], dtype=np.int64))


def test_golden_fhew_tiny(orc):
    """Full bootstrap in the REFERENCE dataflow (pyref, schoolbook products) == the oracle, stage by stage."""
    g, P, keys = golden_fhew_tiny(orc)
    K = orc.FhewKey.from_arrays(P, *keys)
    f = A(g["f"])
    for c in g["cases"]:
        ct = A(c["ct"]).reshape(1, -1)
        pro = K.prologue(ct)
        assert [int(x) for x in pro[0]] == c["prologue"]
        st = orc.fhew_schedule(P, pro[0][:P.n_s])
        assert [[int(a), int(b)] for a, b in st] == c["schedule"]
        out = K.bootstrap(f, ct)
        out[0, -1] = (int(out[0, -1]) + g["post_add"]) % P.big_q
        assert [int(x) for x in out[0]] == c["out"]
        assert [int(x) for x in K.op(g["table"], ct)[0]] == c["out"]


def test_oracle_matches_pyref_f64_fft_and_tfhe(orc):
    """A11/A12: the C++ oracle against the independent pure-Python restatement of c64.rs / fft.rs / tggsw.rs / tlwe.rs:
    every torus word identical (the f64 operation order is part of the contract)."""
    for log_n in range(0, 7):
        n = 1 << log_n
        a = orc.splitmix64(0x600 + log_n, n)
        for log_b in (8, 23):
            dig = (orc.splitmix64(0x700 + log_n + log_b, n) % np.uint64(1 << log_b)).astype(np.int64) - (1 << (log_b - 1))
            b = dig.astype(np.uint64)
            assert [int(x) for x in orc.fft64_mul(a, b)] == pyref.fft64_negacyclic_mul([int(x) for x in a], [int(x) for x in b]), (log_n, log_b)
    for v in (0.0, 0.5, -0.5, 1.5, 2.5, -2.5, 2.0 ** 52 + 1, -(2.0 ** 63), 2.0 ** 63, 2.0 ** 64, 2.0 ** 64 + 4096, 3.0 * 2.0 ** 80, -1.0e30, 1e-30):
        exact = int(pyref.rust_round(v)) if abs(v) < 2.0 ** 53 else int(v)
        assert pyref.f64_mod_u64(v) == exact % (1 << 64), v
    for k, d, log_b in ((1, 2, 8), (2, 3, 6), (1, 1, 23)):
        P = orc.tfhe_testing_param()
        P.n, P.big_n, P.k, P.bs_log_b, P.bs_d, P.ks_log_b, P.ks_d = 3, 16, k, log_b, d, 4, 5
        K = orc.TfheKey(P, 0x5EED0021 + k)
        ex = K.export()
        glwe = orc.splitmix64(5 + k, (k + 1) * P.big_n).reshape(k + 1, P.big_n)
        other = orc.splitmix64(9 + k, (k + 1) * P.big_n).reshape(k + 1, P.big_n)
        L = lambda m: [[int(x) for x in r] for r in m]
        for i in range(P.n):
            rows = [L(r) for r in ex["brk"][i]]
            assert L(K.external_product(i, glwe)) == pyref.tggsw_external_product(log_b, d, rows, L(glwe)), (k, i)
            cm = pyref.tggsw_cmux(log_b, d, rows, L(glwe), L(other))
            ref = glwe + K.external_product(i, other - glwe)
            assert L(ref) == cm
        ct = orc.splitmix64(77 + k, k * P.big_n + 1)
        a_out, b_out = pyref.tlwe_key_switch(P.ks_log_b, P.ks_d, L(ex["ksk_a"]), [int(x) for x in ex["ksk_b"]], [int(x) for x in ct[:-1]], int(ct[-1]))
        assert [int(x) for x in K.key_switch(ct)] == a_out + [b_out]


def test_oracle_matches_pyref_tfhe_bootstrap(orc):
    """A11: whole programmable bootstraps (mod switch, n exact-FFT CMUX steps, sample extract, key switch) at tiny parameters."""
    L = lambda m: [[int(x) for x in r] for r in m]
    for k, d, log_b in ((1, 2, 8), (2, 1, 12)):
        P = orc.tfhe_testing_param()
        P.n, P.big_n, P.k, P.bs_log_b, P.bs_d, P.ks_log_b, P.ks_d = 4, 16, k, log_b, d, 4, 5
        K = orc.TfheKey(P, 0x5EED0041 + k)
        ex = K.export()
        cts = K.encrypt(np.arange(4, dtype=np.uint64), 9)
        cts[1, 0] = 0  # a zero mask word (the CUDA kernel skips that CMUX; the reference computes acc + ext(0))
        v = K.lut_poly(np.arange(1 << P.log_p, dtype=np.uint64))
        ref = K.bootstrap(v, cts)
        brk = [[L(r) for r in ex["brk"][i]] for i in range(P.n)]
        for c in range(4):
            a, b = pyref.tfhe_bootstrap(P.log_p, P.padding, k, log_b, d, P.ks_log_b, P.ks_d, brk, L(ex["ksk_a"]), [int(x) for x in ex["ksk_b"]],
                                        [int(x) for x in v], [int(x) for x in cts[c]])
            assert [int(x) for x in ref[c]] == a + [b], (k, c)


def test_oracle_matches_pyref_rns_and_ckks(orc):
    """A13/A14: rescale_k (both branches), key switch and Ckks::mul against the pure-Python big-integer restatement of
    rns.rs:99-132, 331-345 and ckks.rs:255-293 (polynomial products by schoolbook)."""
    primes = orc.two_adic_primes(55, 5, 8)
    n = 8
    for nq, k in ((4, 1), (6, 3), (8, 4), (2, 1)):
        qs = primes[:nq]
        x = np.stack([orc.residues(17 * nq + i, n, q) for i, q in enumerate(qs)])
        x[:, 0] = [q - 1 for q in qs]
        got = orc.rns_rescale_k(qs, k, x)
        for c in range(n):
            assert [int(v) for v in got[:, c]] == pyref.rns_rescale_k(qs, k, [int(v) for v in x[:, c]]), (nq, k, c)
    log_n, big_l = 3, 3
    K = orc.CkksKey(log_n, 55, big_l, 0x5EED0031, auto_ts=(5,))
    L = lambda m: [[int(v) for v in r] for r in m]
    rng = np.random.default_rng(5)
    ksk = K.ksk(-1)  # [2 (b, a)][2L (qs then ps)][N]
    for level in (3, 2):
        ct0 = K.encrypt(rng.integers(-1000, 1000, size=K.n, dtype=np.int64), level, 1)
        ct1 = K.encrypt(rng.integers(-1000, 1000, size=K.n, dtype=np.int64), level, 2)
        idx = list(range(level)) + list(range(big_l, 2 * big_l))
        kb, ka = L(ksk[0][idx]), L(ksk[1][idx])
        qs_l = K.qs[:level]
        ref = K.key_switch(-1, ct0, apply_auto=False)
        b, a = pyref.ckks_key_switch(qs_l, K.ps, kb, ka, L(ct0[0]), L(ct0[1]))
        assert L(ref[0]) == b and L(ref[1]) == a, level
        prod = K.mul(ct0, ct1)
        pb, pa = pyref.ckks_mul(qs_l, K.ps, kb, ka, (L(ct0[0]), L(ct0[1])), (L(ct1[0]), L(ct1[1])))
        assert L(prod[0]) == pb and L(prod[1]) == pa, level
        rk = K.ksk(0)  # rotation key for t = 5
        rot = K.key_switch(0, ct0, apply_auto=True)
        rb, ra = pyref.ckks_rotate(qs_l, K.ps, 5, L(rk[0][idx]), L(rk[1][idx]), L(ct0[0]), L(ct0[1]))
        assert L(rot[0]) == rb and L(rot[1]) == ra, level
        if level >= 2:
            pt = [[int(v) for v in rng.integers(0, q, size=K.n, dtype=np.uint64)] for q in qs_l]
            limb = np.stack([np.stack([orc.ntt_mul(q, ct0[h, t], A(pt[t])) for t, q in enumerate(qs_l)]) for h in range(2)])
            mc = K.rescale(limb)
            mb, ma = pyref.ckks_mul_constant(qs_l, pt, L(ct0[0]), L(ct0[1]))
            assert L(mc[0]) == mb and L(mc[1]) == ma, level


GOLD2 = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "tfhe_ckks.json")))


def test_golden_tfhe_ckks_primitives(orc):
    """tests/golden/tfhe_ckks.json (written by pyref alone): the oracle reproduces the f64 FFT products and rescale_k words."""
    for c in GOLD2["fft64_mul"]:
        assert [int(x) for x in orc.fft64_mul(A(c["a"]), A(c["b"]))] == c["out"]
    for c in GOLD2["rns_rescale_k"]:
        x = A(c["x"]).T  # [limb][coefficient]
        got = orc.rns_rescale_k(c["qs"], c["k"], x)
        assert [[int(v) for v in got[:, i]] for i in range(x.shape[1])] == c["out"]
