"""Bootstrapping::slot_to_coeff / coeff_to_slot (scheme/ckks/src/bootstrapping.rs:73-108) as chained
fhe_ckks_mul_mat calls over the reference's r = 3 factor matrices (sfft.rs:75-104), diagonals encoded once on the host, the
rotation-key set derived from the BSGS plans (bootstrapping.rs:56-71).  Every intermediate ciphertext equals, limb for limb,
the composition of the oracle's rotate / mul_constant / add in the reference's order; decryptions decode to the special FFT
of the input slots (the reference's own test, bootstrapping.rs:121-143)."""
import numpy as np
import pytest

import ckks_boot_ref as ref

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_n,backend", [(2, "mp"), (4, "mp"), (6, "mp"), (9, "f64")])
def test_slot_to_coeff_and_back_match_the_oracle_chain(pkg, ctx, orc, log_n, backend):
    from learn_fhe_b200 import ckks, ckks_bootstrapping as cb
    big_l = 8
    K0 = orc.CkksKey(log_n, 55, big_l, 1)
    P = ckks.CkksParam(ctx, log_n, K0.qs, K0.ps)
    bp = cb.BootstrappingParam(P, 3, backend)
    js = bp.rotation_indices()
    K = orc.CkksKey(log_n, 55, big_l, 0x5EED0005, auto_ts=tuple(bp.rotation_exponent(j) for j in js))
    key_index = {j: i for i, j in enumerate(js)}
    bk = cb.BootstrappingKey(bp, lambda j: K.ksk(key_index[j]))
    rng = np.random.default_rng(log_n)
    count = 2
    m0 = rng.uniform(-1, 1, (count, bp.l)) + 1j * rng.uniform(-1, 1, (count, bp.l))
    cts = []
    for c in range(count):
        z = cb.sifft(bp.B, [bp.B.mp.mpc(complex(x)) for x in m0[c]] if backend == "mp" else m0[c])
        re, im = bp.B.re_im(z)
        ints = bp.B.trunc_scaled(np.concatenate([re, im]), P.qs[big_l - 1])
        cts.append(K.encrypt(np.array(ints, dtype=np.int64), big_l, 30 + c))
    ct0 = np.stack(cts)
    ct1 = cb.Bootstrapping.slot_to_coeff(bk, ct0)
    n_mats = len(bp.sfft_fmats)
    assert ct1.shape == (count, 2, big_l - n_mats, P.n)
    for c in range(count):
        want = ref.chain(orc, K, key_index, bp, "sfft", ct0[c])
        assert (ct1[c] == want).all(), ("slot_to_coeff", log_n, c)
        got = ref.decode(P, 1, ref.crt_centered(K.qs[:ct1.shape[2]], K.decrypt(ct1[c])))
        m1 = ref.sfft(ref.bit_reverse(m0[c]))
        assert np.abs(got - m1).max() < 1e-4 * max(1.0, np.abs(m1).max())
    if big_l - 2 * n_mats >= 1:
        ct2 = cb.Bootstrapping.coeff_to_slot(bk, ct1)
        for c in range(count):
            assert (ct2[c] == ref.chain(orc, K, key_index, bp, "sifft", ct1[c])).all(), ("coeff_to_slot", log_n, c)
            got = ref.decode(P, 1, ref.crt_centered(K.qs[:ct2.shape[2]], K.decrypt(ct2[c])))
            assert np.abs(got - m0[c]).max() < 1e-4
    bk.free()
    P.free()


def test_bootstrapping_key_generated_on_the_device(pkg, ctx, orc):
    """BootstrappingKey.key_gen: the rotation keys of the BSGS plans come from fhe_ckks_keygen (counter stream) and equal the
    checker's keys for the same seed and exponents, so SlotToCoeff under them is the oracle chain, and it decrypts under the
    returned secret."""
    from learn_fhe_b200 import ckks, ckks_bootstrapping as cb
    log_n, big_l, seed = 6, 8, 0x5EED0B00
    K0 = orc.CkksKey(log_n, 55, big_l, 1)
    P = ckks.CkksParam(ctx, log_n, K0.qs, K0.ps)
    bp = cb.BootstrappingParam(P, 3, "mp")
    js = bp.rotation_indices()
    bk, sk = cb.BootstrappingKey.key_gen(bp, seed)
    K = orc.CkksKey(log_n, 55, big_l, seed, auto_ts=tuple(bp.rotation_exponent(j) for j in js), ctr=True)
    assert (sk == K.sk()).all() and sorted(bk.rtk) == sorted(js)
    key_index = {j: i for i, j in enumerate(js)}
    rng = np.random.default_rng(3)
    m0 = rng.uniform(-1, 1, bp.l) + 1j * rng.uniform(-1, 1, bp.l)
    z = cb.sifft(bp.B, [bp.B.mp.mpc(complex(x)) for x in m0])
    re, im = bp.B.re_im(z)
    ints = bp.B.trunc_scaled(np.concatenate([re, im]), P.qs[big_l - 1])
    ct0 = K.encrypt(np.array(ints, dtype=np.int64), big_l, 77)[None]
    ct1 = cb.Bootstrapping.slot_to_coeff(bk, ct0)
    assert (ct1[0] == ref.chain(orc, K, key_index, bp, "sfft", ct0[0])).all()
    got = ref.decode(P, 1, ref.crt_centered(K.qs[:ct1.shape[2]], K.decrypt(ct1[0])))
    m1 = ref.sfft(ref.bit_reverse(m0))
    assert np.abs(got - m1).max() < 1e-4 * max(1.0, np.abs(m1).max())
    # the relinearisation key generated alongside
    a = K.encrypt(rng.integers(-99, 99, size=P.n, dtype=np.int64), big_l, 1)[None]
    b = K.encrypt(rng.integers(-99, 99, size=P.n, dtype=np.int64), big_l, 2)[None]
    assert (ckks.Ckks.mul(P, bk.rlk, a, b) == K.mul(a, b)).all()
    bk.free()
    P.free()
