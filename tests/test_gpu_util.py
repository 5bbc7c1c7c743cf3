"""Parity of the util-level CUDA kernels (element-wise, automorphism, monomial, mod-switch, decomposition) with the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_elementwise_ops(pkg, ctx, orc):
    from learn_fhe_b200 import util
    for q in (orc.two_adic_primes(55, 12, 1)[0], orc.two_adic_primes(61, 5, 1)[0], 268409857, 12289, 7):
        n = 5000
        a = orc.residues(3, n, q)
        b = orc.residues(4, n, q)
        a[:4] = [0, q - 1, 0, q - 1]
        b[:4] = [0, q - 1, q - 1, 0]
        ta, tb = pkg.to_dev(a), pkg.to_dev(b)
        out = torch.empty_like(ta)
        for name, fn in (("mul", util.pointwise_mul_dev), ("add", util.vec_add_dev), ("sub", util.vec_sub_dev)):
            fn(ctx, q, ta, tb, out)
            ctx.sync()
            assert (pkg.to_host(out) == orc.vec_op(name, q, a, b)).all(), (name, q)
        util.vec_neg_dev(ctx, q, ta, out)
        ctx.sync()
        assert (pkg.to_host(out) == orc.vec_op("neg", q, a)).all()
        acc = pkg.to_dev(b.copy())
        util.pointwise_mac_dev(ctx, q, ta, tb, acc)
        ctx.sync()
        exp = orc.vec_op("add", q, b, orc.vec_op("mul", q, a, b))
        assert (pkg.to_host(acc) == exp).all()
        sc = int(orc.residues(9, 1, q)[0])
        util.vec_scalar_mul_dev(ctx, q, ta, sc, out)
        ctx.sync()
        assert (pkg.to_host(out) == orc.vec_op("mul", q, a, np.full(n, sc, dtype=np.uint64))).all()


@pytest.mark.parametrize("log_n", [0, 1, 4, 9, 11])
def test_automorphism_and_monomial(pkg, ctx, orc, log_n):
    from learn_fhe_b200 import util
    n = 1 << log_n
    q = 268409857
    a = orc.residues(5 + log_n, n, q)
    a[0] = 0
    w = orc.splitmix64(6 + log_n, n)
    ta, tw = pkg.to_dev(a), pkg.to_dev(w)
    out = torch.empty_like(ta)
    for t in (1, 5, -5, 25, 2 * n - 1, -1, 3, 2 * n + 5, -(2 * n) - 3):
        if n == 1 and t % 2 == 0:
            continue
        util.automorphism_dev(ctx, q, ta, log_n, t, out)
        ctx.sync()
        assert (pkg.to_host(out) == orc.automorphism_zq(q, a, t)).all(), t
        util.automorphism_dev(ctx, 0, tw, log_n, t, out)
        ctx.sync()
        assert (pkg.to_host(out) == orc.automorphism_t64(w, t)).all(), t
    for k in (0, 1, n - 1, n, n + 1, 2 * n - 1, -1, -n, 5 * n + 3, -7 * n - 2):
        util.monomial_mul_dev(ctx, q, ta, log_n, k, out)
        ctx.sync()
        assert (pkg.to_host(out) == orc.monomial_mul_zq(q, a, k)).all(), k
        util.monomial_mul_dev(ctx, 0, tw, log_n, k, out)
        ctx.sync()
        assert (pkg.to_host(out) == orc.monomial_mul_t64(w, k)).all(), k


def test_mod_switch(pkg, ctx, orc):
    from learn_fhe_b200 import util
    cases = [(268409857, 1 << 16), (1 << 16, 1024), (1 << 20, 4096), (orc.two_adic_primes(55, 12, 1)[0], 1 << 20), (1024, 268409857)]
    for q, qp in cases:
        n = 20000
        a = orc.residues(77, n, q)
        a[:3] = [0, q - 1, q // 2]
        # values around rounding ties of v*qp/q
        ta = pkg.to_dev(a)
        out = torch.empty_like(ta)
        util.mod_switch_dev(ctx, q, qp, ta, out)
        ctx.sync()
        assert (pkg.to_host(out) == orc.mod_switch(q, qp, a)).all(), (q, qp)
        util.mod_switch_dev(ctx, q, qp, ta, out, odd=True)
        ctx.sync()
        assert (pkg.to_host(out) == orc.mod_switch(q, qp, a, odd=True)).all(), (q, qp)
    # exhaustive over Z_{2^16} -> Z_1024 (the FHEW-T mod_switch_odd domain)
    a = np.arange(1 << 16, dtype=np.uint64)
    ta = pkg.to_dev(a)
    out = torch.empty_like(ta)
    util.mod_switch_dev(ctx, 1 << 16, 1024, ta, out, odd=True)
    ctx.sync()
    assert (pkg.to_host(out) == orc.mod_switch(1 << 16, 1024, a, odd=True)).all()


def test_decompose_zq(pkg, ctx, orc):
    from learn_fhe_b200 import util
    cases = [(268409857, 7, 4), (1 << 16, 4, 4), (268409857, 5, 4), (orc.two_adic_primes(55, 12, 1)[0], 11, 5),
             (orc.two_adic_primes(54, 10, 1)[0], 6, 9), (1 << 20, 4, 5), (orc.two_adic_primes(45, 10, 1)[0], 5, 9), (268409857, 14, 2)]
    for q, log_b, d in cases:
        n = 8192
        a = orc.residues(31, n, q)
        a[:6] = [0, 1, q - 1, q // 2, q // 2 + 1, (q // 2) - 1]
        ta = pkg.to_dev(a)
        out = torch.empty((d, n), dtype=torch.int64, device="cuda")
        util.decompose_zq_dev(ctx, q, log_b, d, ta, out)
        ctx.sync()
        assert (pkg.to_host(out) == orc.decompose_zq(q, log_b, d, a)).all(), (q, log_b, d)


def test_decompose_t64(pkg, ctx, orc):
    from learn_fhe_b200 import util
    for log_b, d in [(23, 1), (4, 5), (8, 8), (7, 3), (2, 8), (16, 4), (1, 3)]:
        n = 8192
        a = orc.splitmix64(41, n)
        a[:4] = [0, 1, (1 << 64) - 1, 1 << 63]
        ta = pkg.to_dev(a)
        out = torch.empty((d, n), dtype=torch.int64, device="cuda")
        util.decompose_t64_dev(ctx, log_b, d, ta, out)
        ctx.sync()
        assert (pkg.to_host(out) == orc.decompose_t64(log_b, d, a)).all(), (log_b, d)
        o2 = torch.empty_like(ta)
        for bits in (0, 1, 41, 44, 52, 63):
            util.rounding_shr_t64_dev(ctx, bits, ta, o2)
            ctx.sync()
            assert (pkg.to_host(o2) == orc.rounding_shr_t64(a, bits)).all(), bits
