"""key generation on the device and the serialised key format.
fhe_fhew_keygen evaluates Bootstrapping::key_gen (scheme/fhew/src/bootstrapping.rs:122-146) on the GPU from the counter-based
stream of csrc/keygen_stream.cuh; oracle/orc_keygen.hpp evaluates the same formulas on the same stream on the CPU.  The two keys
must be equal word for word, and gates evaluated under the device key must decrypt correctly with the returned secrets."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _param(pkg, orc, log_n, bits, log_b, d, n_s, q_ks_bits, ks, w):
    q = orc.two_adic_primes(bits, log_n + 1, 1)[0]
    kw = dict(log_n=log_n, big_q=q, p=4, rlwe_log_b=log_b, rlwe_d=d, rgsw_log_b=log_b, rgsw_d=d, n_s=n_s, q_ks=1 << q_ks_bits, ks_log_b=ks[0],
              ks_d=ks[1], w=w)
    P = orc.fhew_testing_param()
    for k, v in kw.items():
        setattr(P, k, v)
    return pkg.FhewParam(**kw), P


@pytest.mark.parametrize("shape", [(9, 28, 7, 4, 100, 16, (4, 4), 10),   # FHEW-T (fhew/boolean.rs:225-239): fast 32-bit kernels
                                   (6, 28, 7, 4, 12, 16, (4, 4), 3),     # reduced ring: generic 32-bit kernels
                                   (7, 45, 9, 5, 10, 20, (4, 5), 4)])    # 64-bit residues (the multi-key example's word size)
def test_device_keygen_equals_host_keygen_on_the_same_stream(pkg, ctx, orc, shape):
    from learn_fhe_b200 import fhew
    param, P = _param(pkg, orc, *shape)
    seed = 0x5EED0700 + shape[0]
    bk, z, s, ex = fhew.BootstrappingKey.key_gen(ctx, param, seed, export=True)
    K = orc.FhewKey.ctr(P, seed)
    ref = K.export()
    assert (z == ref["z"]).all() and (s == ref["s"]).all()
    for k in ("ksk_a", "ksk_b", "brk", "ak"):
        assert (ex[k] == ref[k]).all(), k
    # the key object built on the device evaluates gates exactly like the oracle under the same key, and they decrypt
    bits = np.random.default_rng(shape[0]).integers(0, 2, size=32).astype(np.int32)
    cts = K.encrypt(bits, 9)
    lin = (cts[:16] + cts[16:]) % np.uint64(P.big_q)
    got = fhew.Fhew.op(bk, [1, 1, 1, 0], lin)
    assert (got == K.op([1, 1, 1, 0], lin, threads=4)).all()
    assert (K.decrypt(got) == 1 - (bits[:16] & bits[16:])).all()
    # without the export nothing but the two secrets leaves the device; same key again
    bk2, z2, s2 = fhew.BootstrappingKey.key_gen(ctx, param, seed)
    assert (z2 == z).all() and (fhew.Fhew.op(bk2, [1, 1, 1, 0], lin) == got).all()
    bk2.free()
    # serialised format: round trip reproduces the key bit for bit (same outputs), damaged blobs are rejected
    blob = bk.serialize()
    bk3 = fhew.BootstrappingKey.deserialize(ctx, blob)
    assert bk3.param.n_s == param.n_s and bk3.param.big_q == param.big_q
    assert (fhew.Fhew.op(bk3, [1, 1, 1, 0], lin) == got).all()
    assert (bk3.serialize() == blob).all()
    bk3.free()
    for damage in ("magic", "version", "truncate", "section"):
        bad = blob.copy()
        if damage == "magic":
            bad[0] ^= 0xFF
        elif damage == "version":
            bad[8] = 9
        elif damage == "truncate":
            bad = bad[:-8]
        else:
            bad[32] ^= 1  # brk_bytes
        with pytest.raises(pkg.FheError):
            fhew.BootstrappingKey.deserialize(ctx, bad)
    bk.free()


@pytest.mark.parametrize("log_n,big_l", [(4, 3), (10, 4), (13, 8)])
def test_ckks_device_keygen_equals_host_keygen_on_the_same_stream(pkg, ctx, orc, log_n, big_l):
    """fhe_ckks_keygen (ckks.rs:139-184 on the GPU) == oracle/orc_keygen.hpp on the same counter stream: secret, relinearisation key
    and two automorphism keys word for word; products / rotations under the device keys equal the oracle's; serialise round trip."""
    from learn_fhe_b200 import ckks
    n = 1 << log_n
    ts = (5, 2 * n - 1)
    seed = 0x5EED0800 + log_n
    K = orc.CkksKey(log_n, 55, big_l, seed, auto_ts=ts, ctr=True)
    P = ckks.CkksParam(ctx, log_n, K.qs, K.ps)
    sk, rlk, autk, ex = ckks.key_gen(P, seed, ts, export=True)
    assert (sk == K.sk()).all()
    assert (ex[0] == K.ksk(-1)).all()
    for i in range(len(ts)):
        assert (ex[1 + i] == K.ksk(i)).all(), i
    rng = np.random.default_rng(log_n)
    ct0 = K.encrypt(rng.integers(-99, 99, size=n, dtype=np.int64), big_l, 1)[None]
    ct1 = K.encrypt(rng.integers(-99, 99, size=n, dtype=np.int64), big_l, 2)[None]
    assert (ckks.Ckks.mul(P, rlk, ct0, ct1) == K.mul(ct0, ct1)).all()
    assert (ckks.Ckks.key_switch(P, autk[0], ct0, ts[0])[0] == K.key_switch(0, ct0[0], apply_auto=True)).all()
    blob = rlk.serialize()
    again = ckks.CkksKeySwitchingKey.deserialize(P, blob)
    assert (ckks.Ckks.mul(P, again, ct0, ct1) == K.mul(ct0, ct1)).all() and (again.serialize() == blob).all()
    bad = blob.copy()
    bad[16] ^= 1  # log_n of the header
    with pytest.raises(pkg.FheError):
        ckks.CkksKeySwitchingKey.deserialize(P, bad)
    for k in [rlk, again] + autk:
        k.free()
    P.free()


@pytest.mark.parametrize("n,big_n,k,bs,ks", [(6, 64, 2, (8, 3), (4, 5)), (12, 512, 1, (10, 2), (4, 5)), (16, 2048, 1, (23, 1), (4, 5))])
def test_tfhe_device_keygen_equals_host_keygen_on_the_same_stream(pkg, ctx, orc, n, big_n, k, bs, ks):
    """fhe_tfhe_keygen (tfhe/bootstrapping.rs:59-76 on the GPU) == oracle/orc_keygen.hpp on the same counter stream: secrets, every
    TGGSW row (the mask-secret products are the reference's f64 FFT products, bit-identical) and the TLWE key-switching key
    word for word; programmable bootstraps under the device key equal the oracle's under the same key and decrypt."""
    from learn_fhe_b200 import tfhe
    P = orc.tfhe_testing_param()
    P.n, P.big_n, P.k, P.bs_log_b, P.bs_d, P.ks_log_b, P.ks_d = n, big_n, k, bs[0], bs[1], ks[0], ks[1]
    param = pkg.TfheParam(log_p=P.log_p, padding=P.padding, n=n, ks_log_b=ks[0], ks_d=ks[1], log_big_n=big_n.bit_length() - 1, k=k, bs_log_b=bs[0],
                          bs_d=bs[1])
    seed = 0x5EED0900 + n
    bk, z, s, ex = tfhe.BootstrappingKey.key_gen(ctx, param, P.tlwe_std, P.tglwe_std, seed, export=True)
    K = orc.TfheKey.ctr(P, seed)
    ref = K.export()
    assert (z == ref["z"]).all() and (s == ref["s"]).all()
    for key in ("brk", "ksk_a", "ksk_b"):
        assert (ex[key] == ref[key]).all(), key
    msgs = np.arange(8, dtype=np.uint64) % np.uint64(1 << P.log_p)
    cts = K.encrypt(msgs, 4)
    v = K.lut_poly(np.arange(1 << P.log_p, dtype=np.uint64))
    got = tfhe.Bootstrapping.bootstrap(bk, tfhe.encode_lut(param, v), cts)
    assert (got == K.bootstrap(v, cts, threads=4)).all()
    if big_n >= 512:  # the tiny ring has no noise margin for a look-up
        assert (K.decrypt(got)[0] == msgs).all()
    # serialised key: round trip, same outputs in every mode the parameters support, corrupted blobs rejected
    blob = bk.serialize()
    again = tfhe.BootstrappingKey.deserialize(ctx, blob)
    assert (again.param.n, again.param.log_big_n, again.param.bs_d) == (param.n, param.log_big_n, param.bs_d)
    assert (tfhe.Bootstrapping.bootstrap(again, tfhe.encode_lut(param, v), cts) == got).all()
    assert (again.serialize() == blob).all()
    if k == 1 and big_n >= 512 and bs[0] * bs[1] <= 31:
        for mode in (2, 3):
            bk.set_mode(mode)
            again.set_mode(mode)
            assert (tfhe.Bootstrapping.bootstrap(again, tfhe.encode_lut(param, v), cts) == tfhe.Bootstrapping.bootstrap(bk, tfhe.encode_lut(param, v), cts)).all()
    for bad in (blob[:-8], np.concatenate([blob, np.zeros(8, dtype=np.uint8)]), np.concatenate([np.zeros(8, dtype=np.uint8), blob[8:]])):
        with pytest.raises(pkg.FheError):
            tfhe.BootstrappingKey.deserialize(ctx, bad)
    wrong = blob.copy()
    wrong[8] = 9  # version
    with pytest.raises(pkg.FheError):
        tfhe.BootstrappingKey.deserialize(ctx, wrong)
    again.free()
    bk.free()
