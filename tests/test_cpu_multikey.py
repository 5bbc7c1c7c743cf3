"""Multi-key FHEW on the oracle: Rgsw::internal_product (rgsw.rs:130-150) and the key-share merge of
bootstrapping.rs:295-320, checked through the properties the reference's example relies on (examples/multi_key_uint8.rs)."""
import numpy as np
import pytest

import multikey


def small_param(orc, log_n=8, bits=45, log_b=9, d=5, n_s=6, w=3):
    """Reduced multi-key set.  N must stay large against n_s: the odd mod switch to 2N leaves a rounding error of about
    0.6 sqrt(n_s) |s| on the blind-rotation exponent, and the gate tolerates 2N / 8 (here 64 against ~7 for three parties)."""
    P = orc.fhew_testing_param()
    P.log_n, P.big_q, P.p = log_n, orc.two_adic_primes(bits, log_n + 1, 1)[0], 4
    P.rlwe_log_b = P.rgsw_log_b = log_b
    P.rlwe_d = P.rgsw_d = d
    P.n_s, P.q_ks, P.ks_log_b, P.ks_d, P.w = n_s, 1 << 20, 4, 5, w
    return P


def test_internal_product_rows_are_external_products(orc):
    """internal_product(ct0, ct1)[r] == external_product(ct0, ct1[r]): the evaluation-domain dot of rgsw.rs:136-147 is exact, so it
    equals the coefficient-form dot of rgsw.rs:116-128 bit for bit."""
    P = small_param(orc)
    K = orc.FhewKey(P, 5)
    brk = K.export()["brk"]
    ct1 = orc.residues(3, 2 * P.rgsw_d * 2 * P.n, P.big_q).reshape(2 * P.rgsw_d, 2, P.n)
    got = orc.rgsw_internal_product(P.big_q, P.log_n, P.rgsw_log_b, P.rgsw_d, brk[2], ct1)
    for r in range(2 * P.rgsw_d):
        assert (got[r] == K.external_product(2, ct1[r])).all(), r


def test_merged_key_bootstraps_under_the_sum_of_secrets(orc):
    """Two parties, as in examples/multi_key_uint8.rs (const N: usize = 2): the left fold of internal products multiplies the
    brk noise by about sqrt(2 d N) B / 2 per extra party, so a third party does not fit a 45-bit modulus."""
    parties = 2
    P = small_param(orc)
    M = multikey.MultiKey(orc, P, parties, 7 + parties)
    ksk_a, ksk_b, brk, ak = M.merge_reference()
    K = orc.FhewKey.from_arrays(P, ksk_a, ksk_b, brk, ak, M.ak_t)
    bits = np.array([0, 0, 1, 1, 0, 1, 0, 1])
    cts = M.encrypt(bits)
    assert (M.decrypt(cts) == bits).all()
    lin = (cts[:4] + cts[4:]) % np.uint64(P.big_q)
    assert (M.decrypt(K.op([1, 1, 1, 0], lin)) == 1 - (bits[:4] & bits[4:])).all()
