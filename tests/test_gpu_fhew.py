"""FHEW / LMKCDEY parity: every stage of Bootstrapping::bootstrap against the oracle, bit for bit (FHEW-T parameters,
scheme/fhew/src/fhew/boolean.rs:225-239), plus the gate truth tables of boolean.rs:241-318."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def bk(pkg, ctx, fhew_setup):
    from learn_fhe_b200 import fhew
    P, K, ex = fhew_setup
    param = fhew.single_key_testing_param(P.big_q)
    key = fhew.BootstrappingKey(ctx, param, ex["ksk_a"], ex["ksk_b"], ex["brk"], ex["ak"], ex["ak_t"])
    yield key
    key.free()


def _inputs(K, P, count, seed):
    rng = np.random.default_rng(seed)
    bits = rng.integers(0, 2, size=2 * count).astype(np.int32)
    cts = K.encrypt(bits, seed)
    lin = (cts[:count] + cts[count:]) % np.uint64(P.big_q)
    return bits[:count], bits[count:], lin


def test_prologue_and_key_switch(pkg, ctx, orc, fhew_setup, bk):
    P, K, _ = fhew_setup
    _, _, lin = _inputs(K, P, 13, 5)
    d_in = pkg.to_dev(lin)
    out = torch.empty((13, P.n_s + 1), dtype=torch.int64, device="cuda")
    ctx.call("fhe_fhew_prologue_batch", bk.h, 13, pkg.dptr(d_in), pkg.dptr(out))
    ctx.sync()
    assert (pkg.to_host(out) == K.prologue(lin)).all()
    ks_in = orc.mod_switch(P.big_q, P.q_ks, lin)
    d_ks = pkg.to_dev(ks_in)  # keep alive: a temporary would be recycled by torch's allocator before the kernel runs
    ctx.call("fhe_lwe_key_switch_batch", bk.h, 13, pkg.dptr(d_ks), pkg.dptr(out))
    ctx.sync()
    assert (pkg.to_host(out) == K.key_switch(ks_in)).all()


def test_external_product_and_automorphism(pkg, ctx, orc, fhew_setup, bk):
    P, K, _ = fhew_setup
    count = 12
    acc = orc.residues(9, count * 2 * P.n, P.big_q).reshape(count, 2, P.n)
    acc[0, 0, :3] = [0, P.big_q - 1, P.big_q // 2]
    d_acc = pkg.to_dev(acc)
    out = torch.empty_like(d_acc)
    idx = np.array([0, 1, 7, 99, 50, 3, 42, 98, 11, 12, 13, 14], dtype=np.uint32)
    d_idx = pkg.to_dev(idx)
    ctx.call("fhe_fhew_external_product", bk.h, count, pkg.dptr(d_idx), pkg.dptr(d_acc), pkg.dptr(out))
    ctx.sync()
    got = pkg.to_host(out)
    for i in range(count):
        assert (got[i] == K.external_product(int(idx[i]), acc[i])).all(), i
    vidx = np.array([0, 1, 2, 10, 5, 9, 3, 4, 6, 7, 8, 10], dtype=np.uint32)
    d_vidx = pkg.to_dev(vidx)
    ctx.call("fhe_fhew_automorphism", bk.h, count, pkg.dptr(d_vidx), pkg.dptr(d_acc), pkg.dptr(out))
    ctx.sync()
    got = pkg.to_host(out)
    for i in range(count):
        assert (got[i] == K.automorphism(int(vidx[i]), acc[i])).all(), i


def test_blind_rotate_accumulator(pkg, ctx, orc, fhew_setup, bk):
    from learn_fhe_b200 import fhew
    P, K, _ = fhew_setup
    _, _, lin = _inputs(K, P, 3, 6)
    ct2n = K.prologue(lin)
    f = fhew.gate_poly(bk.param, [1, 1, 1, 0])
    out = torch.empty((3, 2, P.n), dtype=torch.int64, device="cuda")
    d_f, d_ct2n = pkg.to_dev(f), pkg.to_dev(ct2n)
    ctx.call("fhe_fhew_blind_rotate_batch", bk.h, pkg.dptr(d_f), 3, pkg.dptr(d_ct2n), pkg.dptr(out))
    ctx.sync()
    got = pkg.to_host(out)
    for i in range(3):
        assert (got[i] == K.blind_rotate(f, ct2n[i])).all(), i


def test_nand_bit_exact_and_decrypts(pkg, ctx, orc, fhew_setup, bk):
    """BASELINE config 1: FHEW NAND at the repo's test parameter set; LWE outputs bit-identical to the oracle."""
    from learn_fhe_b200 import fhew
    P, K, _ = fhew_setup
    m0, m1, lin = _inputs(K, P, 16, 7)
    got = fhew.Fhew.op(bk, [1, 1, 1, 0], lin)
    ref = K.op([1, 1, 1, 0], lin, threads=8)
    assert (got == ref).all()
    assert (K.decrypt(got) == 1 - (m0 & m1)).all()


def test_gate_truth_tables(pkg, ctx, orc, fhew_setup, bk):
    """boolean.rs:257-286: not/and/nand/or/nor/xor/xnor/majority over all inputs (decrypt == plaintext op)."""
    from learn_fhe_b200 import fhew
    P, K, _ = fhew_setup
    ops = {"and": lambda a, b: a & b, "nand": lambda a, b: 1 - (a & b), "or": lambda a, b: a | b, "nor": lambda a, b: 1 - (a | b),
           "xor": lambda a, b: a ^ b, "xnor": lambda a, b: 1 - (a ^ b)}
    a = np.array([0, 0, 1, 1] * 4, dtype=np.int32)
    b = np.array([0, 1, 0, 1] * 4, dtype=np.int32)
    ca, cb = K.encrypt(a, 100), K.encrypt(b, 200)
    for name, fn in ops.items():
        out = fhew.Fhew.gate(bk, name, ca, cb)
        assert (K.decrypt(out) == fn(a, b)).all(), name
    a3 = np.array([(i >> 2) & 1 for i in range(8)], dtype=np.int32)
    b3 = np.array([(i >> 1) & 1 for i in range(8)], dtype=np.int32)
    c3 = np.array([i & 1 for i in range(8)], dtype=np.int32)
    out = fhew.Fhew.gate(bk, "majority", K.encrypt(a3, 1), K.encrypt(b3, 2), K.encrypt(c3, 3))
    assert (K.decrypt(out) == ((a3 + b3 + c3) >= 2).astype(np.int32)).all()
    assert (K.decrypt(fhew.Fhew.not_(bk.param, ca)) == 1 - a).all()


def test_large_batch_consistency(pkg, ctx, orc, fhew_setup, bk):
    """Size-independent property at a bench-like batch: identical ciphertexts give identical outputs, all decrypt right."""
    from learn_fhe_b200 import fhew
    P, K, _ = fhew_setup
    m0, m1, lin = _inputs(K, P, 32, 8)
    big = np.tile(lin, (16, 1))  # 512 gates
    got = fhew.Fhew.op(bk, [1, 1, 1, 0], big)
    assert (got.reshape(16, 32, -1) == got[:32][None]).all()
    assert (K.decrypt(got[:32]) == 1 - (m0 & m1)).all()
    ref = K.op([1, 1, 1, 0], lin[:4], threads=4)
    assert (got[:4] == ref).all()


def test_golden_tiny_parameter_set(pkg, ctx, orc):
    """Committed fixture (tests/golden/util_fhew.json, produced by pyref in the reference dataflow): N=16, 20-bit Q."""
    from learn_fhe_b200 import fhew
    from test_cpu_oracle import golden_fhew_tiny
    g, P, keys = golden_fhew_tiny(orc)
    param = pkg.FhewParam(**{k: g["param"][k] for k in ("log_n", "big_q", "p", "rlwe_log_b", "rlwe_d", "rgsw_log_b", "rgsw_d", "n_s",
                                                       "q_ks", "ks_log_b", "ks_d", "w")})
    key = fhew.BootstrappingKey(ctx, param, *keys)
    cts = np.array([c["ct"] for c in g["cases"]], dtype=np.uint64)
    got = fhew.Bootstrapping.bootstrap(key, np.array(g["f"], dtype=np.uint64), cts, post_add=g["post_add"])
    assert (got == np.array([c["out"] for c in g["cases"]], dtype=np.uint64)).all()
    key.free()


def test_empty_batch_and_bad_parameters(pkg, ctx, orc, fhew_setup, bk):
    from learn_fhe_b200 import fhew
    P, K, ex = fhew_setup
    out = fhew.Fhew.op(bk, [1, 1, 1, 0], np.zeros((0, P.n + 1), dtype=np.uint64))
    assert out.shape == (0, P.n + 1)
    bad = fhew.single_key_testing_param(P.big_q)
    bad.big_q = 268409859  # not prime: the reference panics (ring.rs:258), here FHE_EINVAL
    with pytest.raises(pkg.FheError):
        fhew.BootstrappingKey(ctx, bad, ex["ksk_a"], ex["ksk_b"], ex["brk"], ex["ak"], ex["ak_t"])


def test_out_of_range_indices_and_exponents_are_rejected(pkg, ctx, orc, fhew_setup, bk):
    """The reference panics on a key index out of bounds and on an even non-zero blind-rotation exponent
    (bootstrapping.rs:217-222); the util-level device entry points return FHE_EINVAL before any table is indexed."""
    from learn_fhe_b200 import fhew
    P, K, ex = fhew_setup
    acc = pkg.to_dev(orc.residues(9, 2 * 2 * P.n, P.big_q).reshape(2, 2, P.n))
    out = torch.empty_like(acc)
    for name, bad in (("fhe_fhew_external_product", P.n_s), ("fhe_fhew_external_product", 0x8001), ("fhe_fhew_automorphism", 11)):
        d_idx = pkg.to_dev(np.array([0, bad], dtype=np.uint32))
        with pytest.raises(pkg.FheError):
            ctx.call(name, bk.h, 2, pkg.dptr(d_idx), pkg.dptr(acc), pkg.dptr(out))
    _, _, lin = _inputs(K, P, 2, 6)
    ct2n = K.prologue(lin)
    f = pkg.to_dev(fhew.gate_poly(bk.param, [1, 1, 1, 0]))
    o = torch.empty((2, 2, P.n), dtype=torch.int64, device="cuda")
    for pos, val in ((3, 2 * P.n), (5, 6), (P.n_s, 2 * P.n + 1)):  # too large / even non-zero mask word / body too large
        bad = ct2n.copy()
        bad[1, pos] = val
        d = pkg.to_dev(bad)
        with pytest.raises(pkg.FheError):
            ctx.call("fhe_fhew_blind_rotate_batch", bk.h, pkg.dptr(f), 2, pkg.dptr(d), pkg.dptr(o))
    d = pkg.to_dev(ct2n)
    ctx.call("fhe_fhew_blind_rotate_batch", bk.h, pkg.dptr(f), 2, pkg.dptr(d), pkg.dptr(o))  # the context stays usable
    ctx.call("fhe_fhew_key_check_error", bk.h)
    assert (pkg.to_host(o)[0] == K.blind_rotate(fhew.gate_poly(bk.param, [1, 1, 1, 0]), ct2n[0])).all()
    # upload-time checks: key-switching key words must be < q_ks; n_s must leave room for the schedule scratch
    ksk_bad = ex["ksk_a"].copy()
    ksk_bad[0, 0] = P.q_ks
    with pytest.raises(pkg.FheError):
        fhew.BootstrappingKey(ctx, bk.param, ksk_bad, ex["ksk_b"], ex["brk"], ex["ak"], ex["ak_t"])


def test_host_path_pipelining_matches_device_path(pkg, ctx, orc, fhew_setup):
    """fhe_fhew_bootstrap_batch_host splits large batches into chunks (copy streams overlapped with the kernels): every
    ciphertext must equal the device-resident path, chunk boundaries included (9001 is not a multiple of the chunk size)."""
    import torch
    from learn_fhe_b200 import fhew
    P, K, ex = fhew_setup
    param = fhew.single_key_testing_param(P.big_q)
    bk = fhew.BootstrappingKey(ctx, param, ex["ksk_a"], ex["ksk_b"], ex["brk"], ex["ak"], ex["ak_t"])
    count = 9001
    bits = np.random.default_rng(8).integers(0, 2, size=64).astype(np.int32)
    base = K.encrypt(bits, 77)
    cts = np.ascontiguousarray(np.tile(base, (count // 64 + 1, 1))[:count])
    cts[:, :-1] = (cts[:, :-1] + np.arange(count, dtype=np.uint64)[:, None]) % np.uint64(P.big_q)  # distinct masks, any phase
    f = fhew.gate_poly(param, [1, 1, 1, 0])
    post = fhew.big_q_by_8(param)
    host = fhew.Bootstrapping.bootstrap(bk, f, cts, post_add=post)
    d_in, d_f = pkg.to_dev(cts), pkg.to_dev(f)
    d_out = torch.empty_like(d_in)
    fhew.Bootstrapping.bootstrap_dev(bk, d_f, d_in, d_out, post_add=post)
    ctx.sync()
    assert (pkg.to_host(d_out) == host).all()
    ref = K.op([1, 1, 1, 0], cts[[0, 2250, 2251, 4500, 9000]], threads=5)
    assert (host[[0, 2250, 2251, 4500, 9000]] == ref).all()
    bk.free()


# ---- 64-bit modulus path (the parameter shape of examples/multi_key_uint8.rs:15-29) -------------------
def _wide_setup(pkg, ctx, orc, log_n, n_s, seed, bits=55):
    """55-bit Q = two_adic_primes(55, log_n + 1).next(), RLWE / RGSW decomposor (11, 5), LWE q = 2^20 with (4, 5), w = 10."""
    from learn_fhe_b200 import fhew
    P = orc.fhew_testing_param()
    P.log_n, P.big_q = log_n, orc.two_adic_primes(bits, log_n + 1, 1)[0]
    P.rlwe_log_b = P.rgsw_log_b = 11
    P.rlwe_d = P.rgsw_d = 5
    P.n_s, P.q_ks, P.ks_log_b, P.ks_d, P.w = n_s, 1 << 20, 4, 5, 10
    K = orc.FhewKey(P, seed)
    ex = K.export()
    param = pkg.FhewParam(log_n=log_n, big_q=P.big_q, p=4, rlwe_log_b=11, rlwe_d=5, rgsw_log_b=11, rgsw_d=5, n_s=n_s, q_ks=1 << 20,
                          ks_log_b=4, ks_d=5, w=10)
    key = fhew.BootstrappingKey(ctx, param, ex["ksk_a"], ex["ksk_b"], ex["brk"], ex["ak"], ex["ak_t"])
    return P, K, param, key


@pytest.mark.parametrize("log_n,bits", [(5, 55), (8, 55), (7, 60)])
def test_wide_modulus_steps_and_gates_reduced(pkg, ctx, orc, log_n, bits):
    """55-bit Q runs the lazy 64-bit butterflies (Q < 2^56), 60-bit Q the generic ones."""
    from learn_fhe_b200 import fhew
    P, K, param, key = _wide_setup(pkg, ctx, orc, log_n, 24, 0x5EED0009, bits)
    count = 6
    acc = orc.residues(19, count * 2 * P.n, P.big_q).reshape(count, 2, P.n)
    acc[0, 0, :3] = [0, P.big_q - 1, P.big_q // 2]
    d_acc = pkg.to_dev(acc)
    out = torch.empty_like(d_acc)
    idx = np.array([0, 1, 7, 23, 12, 3], dtype=np.uint32)
    d_idx = pkg.to_dev(idx)
    ctx.call("fhe_fhew_external_product", key.h, count, pkg.dptr(d_idx), pkg.dptr(d_acc), pkg.dptr(out))
    ctx.sync()
    got = pkg.to_host(out)
    for i in range(count):
        assert (got[i] == K.external_product(int(idx[i]), acc[i])).all(), i
    vidx = np.array([0, 1, 2, 10, 5, 9], dtype=np.uint32)
    d_vidx = pkg.to_dev(vidx)
    ctx.call("fhe_fhew_automorphism", key.h, count, pkg.dptr(d_vidx), pkg.dptr(d_acc), pkg.dptr(out))
    ctx.sync()
    got = pkg.to_host(out)
    for i in range(count):
        assert (got[i] == K.automorphism(int(vidx[i]), acc[i])).all(), i
    bits = np.array([0, 0, 1, 1, 0, 1, 0, 1] * 3, dtype=np.int32)
    cts = K.encrypt(bits, 17)
    half = len(bits) // 2
    lin = (cts[:half] + cts[half:]) % np.uint64(P.big_q)
    for name in ("nand", "xor"):
        table, pre = fhew.GATES[name]
        x = fhew.Fhew.linear(param, pre, (cts[:half], cts[half:]))
        got = fhew.Fhew.op(key, table, x)
        assert (got == K.op(table, x, threads=4)).all(), name
    nand = fhew.Fhew.op(key, [1, 1, 1, 0], lin)
    assert (K.decrypt(nand) == 1 - (bits[:half] & bits[half:])).all()
    key.free()


def test_wide_modulus_at_the_multi_key_parameter_size(pkg, ctx, orc):
    """N = 2048, Q 55 bits, d = 5, LWE n = 600 (examples/multi_key_uint8.rs:15-29): decryptions of a batch of gates and one
    ciphertext compared word for word with the oracle."""
    from learn_fhe_b200 import fhew
    P, K, param, key = _wide_setup(pkg, ctx, orc, 11, 600, 0x5EED000A)
    bits = np.array([0, 0, 1, 1, 0, 1, 0, 1] * 2, dtype=np.int32)
    cts = K.encrypt(bits, 23)
    half = len(bits) // 2
    lin = (cts[:half] + cts[half:]) % np.uint64(P.big_q)
    got = fhew.Fhew.op(key, [1, 1, 1, 0], lin)
    assert (K.decrypt(got) == 1 - (bits[:half] & bits[half:])).all()
    assert (got[:1] == K.op([1, 1, 1, 0], lin[:1], threads=1)).all()
    key.free()


@pytest.mark.parametrize("log_n", list(range(2, 10)))
def test_reference_ring_sweep_45bit_primes_nine_digits(pkg, ctx, orc, log_n):
    """The shapes of the reference's own RLWE / RGSW tests (testing_n_q, rlwe.rs:337-342: log_n x 45-bit primes of
    two_adic_primes(45, log_n + 1); decomposor log_b = 5, d = 9, rgsw.rs:163-227, rlwe.rs:344-459): external product,
    automorphism + key switch and gate bootstraps through the 64-bit kernels, every word equal to the oracle's."""
    from learn_fhe_b200 import fhew
    for pi, q in enumerate(orc.two_adic_primes(45, log_n + 1, 3)):
        P = orc.fhew_testing_param()
        P.log_n, P.big_q = log_n, q
        P.rlwe_log_b = P.rgsw_log_b = 5
        P.rlwe_d = P.rgsw_d = 9
        P.n_s, P.q_ks, P.ks_log_b, P.ks_d, P.w = 6, 1 << 16, 4, 4, 3
        K = orc.FhewKey(P, 0x5EED0100 + 16 * log_n + pi)
        ex = K.export()
        param = pkg.FhewParam(log_n=log_n, big_q=q, p=4, rlwe_log_b=5, rlwe_d=9, rgsw_log_b=5, rgsw_d=9, n_s=P.n_s, q_ks=P.q_ks,
                              ks_log_b=4, ks_d=4, w=P.w)
        key = fhew.BootstrappingKey(ctx, param, ex["ksk_a"], ex["ksk_b"], ex["brk"], ex["ak"], ex["ak_t"])
        count = P.n_s
        acc = orc.residues(31 + pi, count * 2 * P.n, q).reshape(count, 2, P.n)
        acc[0, 0, :2] = [0, q - 1]
        d_acc = pkg.to_dev(acc)
        out = torch.empty_like(d_acc)
        d_idx = pkg.to_dev(np.arange(count, dtype=np.uint32))
        ctx.call("fhe_fhew_external_product", key.h, count, pkg.dptr(d_idx), pkg.dptr(d_acc), pkg.dptr(out))
        ctx.sync()
        got = pkg.to_host(out)
        for i in range(count):
            assert (got[i] == K.external_product(i, acc[i])).all(), (log_n, q, i)
        nv = P.w + 1
        d_vidx = pkg.to_dev(np.arange(nv, dtype=np.uint32))
        ctx.call("fhe_fhew_automorphism", key.h, nv, pkg.dptr(d_vidx), pkg.dptr(d_acc), pkg.dptr(out))
        ctx.sync()
        got = pkg.to_host(out)
        for v in range(nv):
            assert (got[v] == K.automorphism(v, acc[v])).all(), (log_n, q, v)
        if log_n >= 4:  # a gate needs N >= 16 for the four plateaus of the test polynomial to be distinguishable
            bits = np.array([0, 0, 1, 1, 0, 1, 0, 1], dtype=np.int32)
            cts = K.encrypt(bits, 5)
            lin = (cts[:4] + cts[4:]) % np.uint64(q)
            assert (fhew.Fhew.op(key, [1, 1, 1, 0], lin) == K.op([1, 1, 1, 0], lin, threads=2)).all(), (log_n, q)
        key.free()


def test_multi_key_internal_product_and_key_share_merge(pkg, ctx, orc):
    """Rgsw::internal_product (rgsw.rs:130-150) bit-exact against the oracle - reduced ring and the
    multi-key example's own size (N = 2048, 55-bit Q, decomposor (11, 5), examples/multi_key_uint8.rs:15-29) - and
    Bootstrapping::key_share_merge (bootstrapping.rs:295-320) on the device: merged key == host merge, and the bootstrap under
    it decrypts with the SUM of the parties' secrets."""
    import multikey
    from test_cpu_multikey import small_param
    from learn_fhe_b200 import fhew
    for log_n, bits, log_b, d in ((6, 45, 9, 5), (11, 55, 11, 5), (9, 28, 7, 4)):
        q = orc.two_adic_primes(bits, log_n + 1, 1)[0]
        n = 1 << log_n
        param = pkg.FhewParam(log_n=log_n, big_q=q, p=4, rlwe_log_b=log_b, rlwe_d=d, rgsw_log_b=log_b, rgsw_d=d, n_s=4, q_ks=1 << 16, ks_log_b=4,
                              ks_d=4, w=3)
        ct0 = orc.residues(11 + log_n, 3 * 2 * d * 2 * n, q).reshape(3, 2 * d, 2, n)
        ct1 = orc.residues(12 + log_n, 3 * 2 * d * 2 * n, q).reshape(3, 2 * d, 2, n)
        got = fhew.Rgsw.internal_product(ctx, param, ct0, ct1)
        for c in range(3 if log_n < 11 else 1):
            assert (got[c] == orc.rgsw_internal_product(q, log_n, log_b, d, ct0[c], ct1[c])).all(), (log_n, c)
    for parties in (2, 3):
        P = small_param(orc)
        M = multikey.MultiKey(orc, P, parties, 17 + parties)
        param = pkg.FhewParam(log_n=P.log_n, big_q=P.big_q, p=P.p, rlwe_log_b=P.rlwe_log_b, rlwe_d=P.rlwe_d, rgsw_log_b=P.rgsw_log_b,
                              rgsw_d=P.rgsw_d, n_s=P.n_s, q_ks=P.q_ks, ks_log_b=P.ks_log_b, ks_d=P.ks_d, w=P.w)
        got = fhew.key_share_merge(ctx, param, M.crs, M.shares)
        ref = M.merge_reference()
        for g, r in zip(got, ref):
            assert g.shape == r.shape and (g == r).all()
        if parties > 2:
            continue  # merge parity only: a third party's noise does not fit this reduced modulus (tests/test_cpu_multikey.py)
        bk = fhew.BootstrappingKey(ctx, param, got[0], got[1], got[2], got[3], M.ak_t)
        bits = np.array([0, 0, 1, 1, 0, 1, 0, 1])
        cts = M.encrypt(bits)
        lin = (cts[:4] + cts[4:]) % np.uint64(P.big_q)
        out = fhew.Fhew.op(bk, [1, 1, 1, 0], lin)
        assert (M.decrypt(out) == 1 - (bits[:4] & bits[4:])).all()
        K = orc.FhewKey.from_arrays(P, *ref, M.ak_t)
        assert (out == K.op([1, 1, 1, 0], lin)).all()
        bk.free()
