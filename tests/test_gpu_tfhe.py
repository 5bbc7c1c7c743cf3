"""Parity of the TFHE CUDA path with the oracle: f64 FFT torus product (util/src/ring/fft/c64.rs:150-208), TGGSW external
product, TLWE key switch and programmable bootstrapping (scheme/tfhe/src/bootstrapping.rs:138-165).  The kernels evaluate
the reference's floating-point algorithm in the reference's operation order, so every comparison is BIT-EXACT on the raw
torus words (stronger than the reference's own bound log2(err) <= 64 + log_b + log_n - 53); decryptions are checked too."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("log_n", list(range(0, 13)))
def test_fft64_mul_matches_oracle(pkg, ctx, orc, log_n):
    from learn_fhe_b200 import tfhe
    n = 1 << log_n
    batch = 5
    a = orc.splitmix64(0x5EED2000 + log_n, n * batch).reshape(batch, n)
    for log_b in (12, 17, 23):
        dig = (orc.splitmix64(log_b * 100 + log_n, n * batch) % np.uint64(1 << log_b)).astype(np.int64) - (1 << (log_b - 1))
        b = dig.astype(np.uint64).reshape(batch, n)
        ref = orc.fft64_mul(a, b, threads=4)
        got = tfhe.nega_cyclic_fft64_mul_assign_rt(ctx, a.copy(), b)
        assert (got == ref).all(), (log_n, log_b)
    # exact for small operands (c64.rs:169-184)
    sa = (orc.splitmix64(1 + log_n, n) % np.uint64(32)).astype(np.uint64)
    sb = (orc.splitmix64(2 + log_n, n) % np.uint64(32)).astype(np.uint64)
    assert (tfhe.nega_cyclic_fft64_mul_assign_rt(ctx, sa.copy(), sb) == orc.schoolbook_t64(sa, sb)).all()


def test_fft64_error_bound_of_reference(pkg, ctx, orc):
    """c64.rs:186-208: full-range torus x b-bit digits, log2(max err vs exact product) <= 64 + log_b + log_n - 53."""
    from learn_fhe_b200 import tfhe
    for log_n in (8, 9):
        n = 1 << log_n
        for log_b in (12, 17):
            a = orc.splitmix64(log_n * 7 + log_b, n)
            dig = (orc.splitmix64(log_n * 11 + log_b, n) % np.uint64(1 << log_b)).astype(np.int64) - (1 << (log_b - 1))
            b = dig.astype(np.uint64)
            got = tfhe.nega_cyclic_fft64_mul_assign_rt(ctx, a.copy(), b)
            exact = orc.schoolbook_t64(a, b)
            err = (got - exact).astype(np.int64)
            assert np.abs(err).max() < 2.0 ** (64 + log_b + log_n - 53)


def _small_param(orc, n=40, big_n=256, k=2, bs_log_b=8, bs_d=4, ks_log_b=4, ks_d=5):
    P = orc.tfhe_testing_param()
    P.n, P.big_n, P.k, P.bs_log_b, P.bs_d, P.ks_log_b, P.ks_d = n, big_n, k, bs_log_b, bs_d, ks_log_b, ks_d
    return P


def _upload(pkg, ctx, P, ex):
    from learn_fhe_b200 import tfhe
    param = pkg.TfheParam(log_p=P.log_p, padding=P.padding, n=P.n, ks_log_b=P.ks_log_b, ks_d=P.ks_d,
                          log_big_n=P.big_n.bit_length() - 1, k=P.k, bs_log_b=P.bs_log_b, bs_d=P.bs_d)
    return tfhe.BootstrappingKey(ctx, param, ex["brk"], ex["ksk_a"], ex["ksk_b"])


@pytest.mark.parametrize("k,bs_d,big_n", [(2, 4, 256), (2, 8, 256), (1, 1, 512), (1, 3, 1024)])
def test_external_product_keyswitch_blind_rotate_reduced(pkg, ctx, orc, k, bs_d, big_n):
    """tggsw.rs:134-181 / tlwe.rs:162-192 shapes (N=256, k=2, d=8-style) at reduced n."""
    from learn_fhe_b200 import tfhe
    P = _small_param(orc, k=k, bs_d=bs_d, big_n=big_n, bs_log_b=23 if bs_d == 1 else 8)
    K = orc.TfheKey(P, 0x5EED0003)
    bk = _upload(pkg, ctx, P, K.export())
    count = 5
    glwe = orc.splitmix64(3, count * (P.k + 1) * P.big_n).reshape(count, P.k + 1, P.big_n)
    idx = np.array([0, 1, P.n - 1, 7, 7], dtype=np.uint32)
    got = tfhe.Tggsw.external_product(bk, idx, glwe)
    for c in range(count):
        assert (got[c] == K.external_product(int(idx[c]), glwe[c])).all(), c
    # cmux = ct0 + external_product(b, ct1 - ct0) (tggsw.rs:114-121), wrapping torus arithmetic
    other = orc.splitmix64(33, glwe.size).reshape(glwe.shape)
    got = tfhe.Tggsw.cmux(bk, idx, glwe, other)
    for c in range(count):
        assert (got[c] == glwe[c] + K.external_product(int(idx[c]), other[c] - glwe[c])).all(), c
    ext = orc.splitmix64(4, count * (P.k * P.big_n + 1)).reshape(count, -1)
    got = tfhe.Tlwe.key_switch(bk, ext)
    for c in range(count):
        assert (got[c] == K.key_switch(ext[c])).all()
    msgs = np.arange(count, dtype=np.uint64) % np.uint64(1 << P.log_p)
    cts = K.encrypt(msgs, 9)
    cts[1, 2] = 0
    v = K.lut_poly(np.arange(1 << P.log_p, dtype=np.uint64))
    got = tfhe.Bootstrapping.blind_rotate_extract(bk, tfhe.encode_lut(bk.param, v), cts)
    for c in range(count):
        assert (got[c] == K.blind_rotate_extract(v, cts[c])).all(), c
    full = tfhe.Bootstrapping.bootstrap(bk, tfhe.encode_lut(bk.param, v), cts)
    assert (full == K.bootstrap(v, cts, threads=4)).all()
    assert tfhe.Bootstrapping.bootstrap(bk, tfhe.encode_lut(bk.param, v), cts[:0]).shape == (0, P.n + 1)
    bk.free()


def test_pbs_reference_parameters(pkg, ctx, orc):
    """tfhe/bootstrapping.rs:138-165: LUTs identity / double / parity over all 16 messages at n=1024, N=2048, k=1,
    log_b=23, d=1 — raw TLWE outputs bit-identical to the oracle and decryptions equal to the table."""
    from learn_fhe_b200 import tfhe
    P = orc.tfhe_testing_param()
    K = orc.TfheKey(P, 0x5EED0003)
    bk = _upload(pkg, ctx, P, K.export())
    p = 1 << P.log_p
    msgs = np.arange(p, dtype=np.uint64)
    cts = K.encrypt(msgs, 21)
    for name, table in (("identity", msgs), ("double", (2 * msgs) % p), ("parity", msgs % 2)):
        v = K.lut_poly(table.astype(np.uint64))
        got = tfhe.Bootstrapping.bootstrap(bk, tfhe.encode_lut(bk.param, v), cts)
        assert (K.decrypt(got)[0] == table).all(), name
        ref = K.bootstrap(v, cts[:6], threads=6)
        assert (got[:6] == ref).all(), name
    bk.free()


def test_pbs_fourier_accumulation_mode(pkg, ctx, orc):
    """Optional mode 1 (products summed in the Fourier domain): not the reference's rounding order, so the contract is the
    north star's floating-point one - raw outputs within a stated bound of the reference's, decrypted results identical.
    Bound: per CMUX and coefficient at most (k+1)d roundings of <= 2^(64 + log_b + log_n - 53) (c64.rs:186-208) are
    replaced by one; over n CMUX steps and the key switch the phase moves by < 2^52 here (plaintext scale 2^59)."""
    from learn_fhe_b200 import tfhe
    P = orc.tfhe_testing_param()
    K = orc.TfheKey(P, 0x5EED0003)
    bk = _upload(pkg, ctx, P, K.export())
    p = 1 << P.log_p
    msgs = np.arange(p, dtype=np.uint64)
    cts = K.encrypt(msgs, 21)
    v = K.lut_poly(((3 * msgs + 1) % p).astype(np.uint64))
    lut = tfhe.encode_lut(bk.param, v)
    exact = tfhe.Bootstrapping.bootstrap(bk, lut, cts)
    bk.set_mode(True)
    fast = tfhe.Bootstrapping.bootstrap(bk, lut, cts)
    bk.set_mode(False)
    assert (tfhe.Bootstrapping.bootstrap(bk, lut, cts) == exact).all()  # the default mode is restored exactly
    assert (K.decrypt(fast)[0] == (3 * msgs + 1) % p).all() and (K.decrypt(exact)[0] == (3 * msgs + 1) % p).all()
    ph_fast, ph_exact = K.decrypt(fast)[1], K.decrypt(exact)[1]
    diff = (ph_fast - ph_exact).astype(np.int64)
    assert np.abs(diff).max() < 2 ** 52, np.abs(diff).max()
    assert not (fast == exact).all()  # it really is a different rounding order
    bk.free()


@pytest.mark.parametrize("bs_d,big_n,n,bs_log_b", [(1, 2048, 12, 23), (3, 1024, 9, 7), (2, 512, 8, 10), (2, 2048, 5, 12), (1, 512, 20, 23)])
def test_fused_mode_equals_cpu_replay(pkg, ctx, orc, hostsim, bs_d, big_n, n, bs_log_b):
    """Modes 2 and 3 (tfhe_fast.cuh): the CUDA kernel evaluates exactly the per-thread logic tests/hostsim replays on the CPU (every
    operation is an explicit IEEE multiply / add / fma), so blind-rotation outputs are bit-identical to the replay, whose
    distance from the reference dataflow is bounded in tests/test_cpu_hostsim.py::test_kernel_logic_tfhe_fast_mode."""
    import ctypes as C
    from learn_fhe_b200 import tfhe
    u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
    H = hostsim
    H.sim_tfhe_fast_key.restype = C.c_void_p
    H.sim_tfhe_fast_key.argtypes = [C.c_uint] * 4 + [u64p]
    H.sim_tfhe_fast_key_free.argtypes = [C.c_void_p]
    H.sim_tfhe_fast_blind_rotate_extract.argtypes = [C.c_void_p, u64p, u64p, u64p]
    H.sim_tfhe_fast_blind_rotate_extract32.argtypes = [C.c_void_p, u64p, u64p, u64p]
    P = _small_param(orc, n=n, big_n=big_n, k=1, bs_d=bs_d, bs_log_b=bs_log_b)
    K = orc.TfheKey(P, 0x5EED0003)
    ex = K.export()
    bk = _upload(pkg, ctx, P, ex)
    h = H.sim_tfhe_fast_key(big_n.bit_length() - 1, P.n, P.bs_log_b, P.bs_d, ex["brk"].reshape(-1))
    count = 300  # more ciphertexts than resident CTAs would need at full size; exercises the grid-stride loop at small ones
    msgs = np.arange(count, dtype=np.uint64) % np.uint64(1 << P.log_p)
    cts = K.encrypt(msgs, 9)
    cts[1, 2] = 0
    v = K.lut_poly(np.arange(1 << P.log_p, dtype=np.uint64))
    lut = tfhe.encode_lut(bk.param, v)
    bk.set_mode(0)
    exact = tfhe.Bootstrapping.bootstrap(bk, lut, cts[:32])
    for mode, replay in ((2, H.sim_tfhe_fast_blind_rotate_extract), (3, H.sim_tfhe_fast_blind_rotate_extract32)):  # 64- / 32-bit accumulator words
        bk.set_mode(mode)
        got = tfhe.Bootstrapping.blind_rotate_extract(bk, lut, cts)
        for c in (0, 1, 2, 150, count - 1):
            out = np.zeros(P.big_n + 1, dtype=np.uint64)
            replay(h, lut, cts[c], out)
            assert (got[c] == out).all(), (mode, c)
        # end to end: decrypts like the reference dataflow
        full = tfhe.Bootstrapping.bootstrap(bk, lut, cts[:32])
        assert (K.decrypt(full)[0] == K.decrypt(exact)[0]).all(), mode
    H.sim_tfhe_fast_key_free(h)
    bk.free()


def test_pbs_fused_mode_reference_parameters(pkg, ctx, orc):
    """Modes 2 and 3 at TFHE-T (tfhe/bootstrapping.rs:141-152), all 16 messages x 3 LUTs x 8 fresh encryptions: decryptions identical
    to the table (and to the bit-identical mode).  Once one digit of one CMUX rounds the other way the two modes hold
    different, equally valid encryptions, so their phases differ by the output noise itself (measured <= 2^52.4 here); both
    stay within 2^55 of the encoded message, 1/8 of the decoding margin 2^58 (plaintext scale 2^59)."""
    from learn_fhe_b200 import tfhe
    P = orc.tfhe_testing_param()
    K = orc.TfheKey(P, 0x5EED0003)
    bk = _upload(pkg, ctx, P, K.export())
    p = 1 << P.log_p
    msgs = np.tile(np.arange(p, dtype=np.uint64), 8)
    cts = K.encrypt(msgs, 21)
    for name, table in (("identity", np.arange(p)), ("double", (2 * np.arange(p)) % p), ("parity", np.arange(p) % 2)):
        v = K.lut_poly(table.astype(np.uint64))
        lut = tfhe.encode_lut(bk.param, v)
        bk.set_mode(0)
        exact = tfhe.Bootstrapping.bootstrap(bk, lut, cts)
        want = table[msgs.astype(np.int64)]
        m_exact, ph_exact = K.decrypt(exact)
        enc = (want.astype(np.uint64) << np.uint64(64 - (P.log_p + P.padding))).astype(np.uint64)
        assert (m_exact == want).all(), name
        assert np.abs((ph_exact - enc).astype(np.int64)).max() < 2 ** 55, name
        worst = {}
        for mode in (2, 3):  # 3: the accumulator keeps the top 32 bits of every torus word - the same bounds hold
            bk.set_mode(mode)
            fast = tfhe.Bootstrapping.bootstrap(bk, lut, cts)
            m_fast, ph_fast = K.decrypt(fast)
            assert (m_fast == want).all(), (name, mode)
            worst[mode] = np.abs((ph_fast - enc).astype(np.int64)).max()
            assert worst[mode] < 2 ** 55, (name, mode)
            assert np.abs((ph_fast - ph_exact).astype(np.int64)).max() < 2 ** 54, (name, mode)
        print("tfhe fused noise", name, {m: float(np.log2(float(w))) for m, w in worst.items()})
    bk.free()


def test_tfhe_parameter_errors(pkg, ctx, orc):
    from learn_fhe_b200 import tfhe
    P = _small_param(orc, k=1, bs_d=1, big_n=64, bs_log_b=23)
    K = orc.TfheKey(P, 1)
    ex = K.export()
    bad = pkg.TfheParam(log_p=4, padding=1, n=P.n, ks_log_b=4, ks_d=5, log_big_n=6, k=1, bs_log_b=23, bs_d=3)  # 69 bits
    with pytest.raises(pkg.FheError):
        tfhe.BootstrappingKey(ctx, bad, ex["brk"], ex["ksk_a"], ex["ksk_b"])
    # a bootstrapping-key index >= n is a slice index out of bounds in the reference: FHE_EINVAL, not a wild device read
    good = pkg.TfheParam(log_p=P.log_p, padding=P.padding, n=P.n, ks_log_b=P.ks_log_b, ks_d=P.ks_d, log_big_n=6, k=1, bs_log_b=23, bs_d=1)
    bk = tfhe.BootstrappingKey(ctx, good, ex["brk"], ex["ksk_a"], ex["ksk_b"])
    glwe = orc.splitmix64(3, 2 * 2 * 64).reshape(2, 2, 64)
    with pytest.raises(pkg.FheError):
        tfhe.Tggsw.external_product(bk, np.array([0, P.n], dtype=np.uint32), glwe)
    with pytest.raises(pkg.FheError):
        tfhe.Tggsw.cmux(bk, np.array([P.n + 5, 0], dtype=np.uint32), glwe, glwe)
    with pytest.raises(pkg.FheError):
        bk.set_mode(2)  # no fused specialisation for N = 64
    assert tfhe.Tggsw.external_product(bk, np.array([0, P.n - 1], dtype=np.uint32), glwe).shape == glwe.shape
    bk.free()
