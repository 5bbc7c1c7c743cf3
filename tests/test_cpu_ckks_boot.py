"""Host half of SURVEY.md 8f rank 2 (CoeffToSlot / SlotToCoeff): the factor matrices, BSGS plans and diagonal encodings of
learn-fhe_b200/ckks_bootstrapping.py against the reference's own properties (scheme/ckks/src/sfft.rs:128-139 factorisation test,
scheme/ckks/src/bootstrapping.rs:121-143 round trip), with the oracle standing in for the device."""
import importlib.util
import os
import sys

import numpy as np
import pytest

import ckks_boot_ref as ref

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _mod():
    """ckks_bootstrapping.py without importing the package __init__ (which needs the CUDA library): host-only code."""
    import _pkg
    pkg = _pkg.load_package()
    from learn_fhe_b200 import ckks_bootstrapping
    return ckks_bootstrapping


def _dense(mat):
    n = mat.n
    d = np.zeros((n, n), dtype=np.complex128)
    for j, v in mat.diags.items():
        for i in range(n):
            d[i, (j + i) % n] = complex(v[i])
    return d


@pytest.mark.parametrize("log_l", [1, 2, 3, 5])
def test_factor_matrices_multiply_to_the_special_fft(log_l):
    cb = _mod()
    B = cb._backend("f64")
    n = 1 << log_l
    full = np.eye(n, dtype=np.complex128)
    for m in cb.sfft_fmats(B, n):
        full = full @ _dense(m)
    inv = np.eye(n, dtype=np.complex128)
    for m in cb.sifft_fmats(B, n):
        inv = inv @ _dense(m)
    assert np.allclose(full @ inv, np.eye(n), atol=1e-9)
    # the matrix applied to a slot vector is sfft of its bit reversal (bootstrapping.rs:133: m1 = sfft(bit_reverse(m0)))
    m0 = np.random.default_rng(log_l).standard_normal(n) + 1j * np.random.default_rng(log_l + 9).standard_normal(n)
    assert np.allclose(full @ m0, ref.sfft(ref.bit_reverse(m0)), atol=1e-9)
    # grouping by r and the BSGS plan cover every diagonal exactly once
    for r in (1, 2, 3):
        for mat in cb._group(cb.sfft_fmats(B, n), r):
            bs = mat.bsgs()
            assert sorted(i + j for i, js in bs.items() for j in js) == sorted(mat.diags)


@pytest.mark.parametrize("log_n,backend", [(3, "mp"), (5, "mp"), (6, "f64")])
def test_slot_to_coeff_to_slot_on_the_oracle(orc, log_n, backend):
    """The reference's round-trip test (bootstrapping.rs:121-143) with the oracle evaluating every homomorphic operation:
    decrypt(slot_to_coeff(enc(m0))) decodes to sfft(bit_reverse(m0)), and coeff_to_slot brings it back."""
    cb = _mod()

    class P:  # the fields of CkksParam the host code reads
        pass
    big_l = 8
    K0 = orc.CkksKey(log_n, 55, big_l, 1)
    P.log_n, P.n, P.big_l, P.qs, P.ps = log_n, 1 << log_n, big_l, K0.qs, K0.ps
    bp = cb.BootstrappingParam(P, 3, backend)
    js = bp.rotation_indices()
    K = orc.CkksKey(log_n, 55, big_l, 0x5EED0005, auto_ts=tuple(bp.rotation_exponent(j) for j in js))
    key_index = {j: i for i, j in enumerate(js)}
    rng = np.random.default_rng(log_n)
    m0 = (rng.uniform(-1, 1, bp.l) + 1j * rng.uniform(-1, 1, bp.l))
    z = cb.sifft(bp.B, [bp.B.mp.mpc(complex(x)) for x in m0] if backend == "mp" else m0)
    re, im = bp.B.re_im(z)
    ints = bp.B.trunc_scaled(np.concatenate([re, im]), P.qs[big_l - 1])
    assert np.allclose(ref.decode(P, 1, ints), m0, atol=1e-9)  # encode / decode agree
    ct0 = K.encrypt(np.array(ints, dtype=np.int64), big_l, 3)
    ct1 = ref.chain(orc, K, key_index, bp, "sfft", ct0)
    n_mats = len(bp.sfft_fmats)
    assert ct1.shape[1] == big_l - n_mats
    m1 = ref.sfft(ref.bit_reverse(m0))
    # the scale stays q_last: every mul_constant multiplies by scale and rescales by the dropped prime (both ~2^55)
    got1 = ref.decode(P, 1, ref.crt_centered(K.qs[:ct1.shape[1]], K.decrypt(ct1)))
    assert np.abs(got1 - m1).max() < 1e-4 * max(1.0, np.abs(m1).max()), np.abs(got1 - m1).max()
    if big_l - 2 * n_mats >= 1:
        ct2 = ref.chain(orc, K, key_index, bp, "sifft", ct1)
        got2 = ref.decode(P, 1, ref.crt_centered(K.qs[:ct2.shape[1]], K.decrypt(ct2)))
        assert np.abs(got2 - m0).max() < 1e-4, np.abs(got2 - m0).max()
