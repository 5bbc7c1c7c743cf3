"""Parity of the CUDA NTT kernels with the oracle (restating util/src/ring/fft/zq.rs:94-116 property sweeps)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(pkg, ctx, q, a, fwd, bits):
    from learn_fhe_b200 import util
    log_n = a.shape[-1].bit_length() - 1
    t = pkg.to_dev(a.astype(np.uint32) if bits == 32 else a)
    (util.ntt_fwd_dev if fwd else util.ntt_inv_dev)(ctx, q, t, log_n, bits)
    ctx.sync()
    return pkg.to_host(t).astype(np.uint64)


@pytest.mark.parametrize("log_n", list(range(0, 17)))
def test_ntt_u64_matches_oracle(pkg, ctx, orc, log_n):
    n = 1 << log_n
    batch = 5 if log_n < 14 else 3
    for q in orc.two_adic_primes(55, log_n + 1, 2) + orc.two_adic_primes(61, log_n + 1, 1):
        a = orc.residues(0x5EED0000 + log_n, n * batch, q).reshape(batch, n)
        ref = orc.ntt_fwd(q, a, threads=4)
        got = _run(pkg, ctx, q, a, True, 64)
        assert (got == ref).all(), (log_n, q)
        back = _run(pkg, ctx, q, ref, False, 64)
        assert (back == a).all(), (log_n, q)


@pytest.mark.parametrize("log_n", list(range(0, 17)))
def test_ntt_u32_matches_oracle(pkg, ctx, orc, log_n):
    n = 1 << log_n
    batch = 5 if log_n < 14 else 3
    for bits in (28, 30):
        q = orc.two_adic_primes(bits, log_n + 1, 1)[0]
        a = orc.residues(0x5EED1000 + log_n, n * batch, q).reshape(batch, n)
        ref = orc.ntt_fwd(q, a, threads=4)
        got = _run(pkg, ctx, q, a, True, 32)
        assert (got == ref).all(), (log_n, q)
        back = _run(pkg, ctx, q, ref, False, 32)
        assert (back == a).all(), (log_n, q)


def test_reference_round_trip_and_schoolbook_sweep(pkg, ctx, orc):
    """fft/zq.rs:94-116 + ring.rs:442-452: log_n 0..9, ten 45-bit primes each, through the host-slice C ABI."""
    from learn_fhe_b200 import util
    for log_n in range(0, 10):
        n = 1 << log_n
        for k, q in enumerate(orc.two_adic_primes(45, log_n + 1, 10)):
            a = orc.residues(11 * log_n + k, n, q)
            b = orc.residues(97 * log_n + k, n, q)
            x = a.copy()
            util.nega_cyclic_ntt_in_place(ctx, q, x)
            assert (x == orc.ntt_fwd(q, a)).all()
            util.nega_cyclic_intt_in_place(ctx, q, x)
            assert (x == a).all()
            y = a.copy()
            util.nega_cyclic_ntt_mul_assign(ctx, q, y, b)
            assert (y == orc.schoolbook_zq(q, a, b)).all()


def test_ntt_large_batch_linearity(pkg, ctx, orc):
    """Full-size property check (BASELINE config 2 shape): NTT(a+b) == NTT(a)+NTT(b), iNTT(NTT(a)) == a at N=2^16."""
    import torch
    from learn_fhe_b200 import util
    log_n, batch = 16, 64
    q = orc.two_adic_primes(55, log_n + 1, 1)[0]
    n = 1 << log_n
    a = orc.residues(1, n * batch, q)
    b = orc.residues(2, n * batch, q)
    s = (a + b) % np.uint64(q)
    ta, tb, ts = pkg.to_dev(a), pkg.to_dev(b), pkg.to_dev(s)
    for t in (ta, tb, ts):
        util.ntt_fwd_dev(ctx, q, t, log_n)
    tsum = torch.empty_like(ta)
    util.vec_add_dev(ctx, q, ta, tb, tsum)
    ctx.sync()
    assert torch.equal(tsum, ts)
    util.ntt_inv_dev(ctx, q, ta, log_n)
    ctx.sync()
    assert (pkg.to_host(ta) == a).all()
    # spot-check three polynomials against the oracle
    for i in (0, 17, 63):
        assert (pkg.to_host(tb).reshape(batch, n)[i] == orc.ntt_fwd(q, b.reshape(batch, n)[i])).all()


def test_ntt_rejects_bad_modulus(pkg, ctx):
    import torch
    from learn_fhe_b200 import util
    t = torch.zeros(16, dtype=torch.int64, device="cuda")
    with pytest.raises(pkg.FheError):
        util.ntt_fwd_dev(ctx, 15, t, 4)  # not prime (reference panics)
    with pytest.raises(pkg.FheError):
        util.ntt_fwd_dev(ctx, 13, t, 4)  # 2-adicity too small for n = 16


def test_twiddle_table_matches_reference_rule(pkg, ctx, orc):
    from learn_fhe_b200 import util
    for q in orc.two_adic_primes(28, 10, 2) + orc.two_adic_primes(55, 12, 1):
        f_ref, i_ref = orc.twiddles(q)
        ln = min(len(f_ref), 1 << 11)
        f, i = util.twiddles(ctx, q, ln)
        assert (f == f_ref[:ln]).all() and (i == i_ref[:ln]).all()


def test_ntt_degree_2_17_and_rns_batches(pkg, ctx, orc):
    """Largest supported degree (column S=4 + 2^13 tiles) and the multi-modulus entry point with ragged polynomial counts."""
    log_n = 17
    n = 1 << log_n
    for bits, word in ((55, 64), (28, 32)):
        q = orc.two_adic_primes(bits, log_n + 1, 1)[0]
        a = orc.residues(0x5EED2000 + bits, n * 2, q).reshape(2, n)
        ref = orc.ntt_fwd(q, a, threads=2)
        got = _run(pkg, ctx, q, a, True, word)
        assert (got == ref).all(), (log_n, word)
        assert (_run(pkg, ctx, q, ref, False, word) == a).all(), (log_n, word)
    # fhe_ntt_fwd_rns / inv_rns: [batch][limbs][n], limb = polynomial index % limbs; 3 x 5 = 15 polynomials (not a multiple of 4)
    import torch
    for log_n in (9, 12, 14):
        n = 1 << log_n
        qs = orc.two_adic_primes(55, log_n + 1, 5)
        x = np.stack([np.stack([orc.residues(7 * b + i, n, q) for i, q in enumerate(qs)]) for b in range(3)])
        t = pkg.to_dev(x)
        qa = np.ascontiguousarray(qs, dtype=np.uint64)
        ctx.call("fhe_ntt_fwd_rns", pkg.hptr(qa), len(qs), log_n, 3, pkg.dptr(t))
        ctx.sync()
        got = pkg.to_host(t)
        for b in range(3):
            for i, q in enumerate(qs):
                assert (got[b, i] == orc.ntt_fwd(q, x[b, i])).all(), (log_n, b, i)
        ctx.call("fhe_ntt_inv_rns", pkg.hptr(qa), len(qs), log_n, 3, pkg.dptr(t))
        ctx.sync()
        assert (pkg.to_host(t) == x).all()
        ctx.call("fhe_ntt_fwd_rns", pkg.hptr(qa), len(qs), log_n, 0, pkg.dptr(t))  # empty batch is a no-op
