"""Full-size runs of BASELINE.json's configurations through the C ABI.  The oracle cannot evaluate
these batches in seconds, so every element of the batch is checked through size-independent properties of the domain
(round trip, linearity, known transforms, decrypt-equals-plaintext, commutativity, batch-position independence) and a few
elements of each batch are additionally compared bit for bit with the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------------------------- C2: NTT sweep
@pytest.mark.parametrize("bits", [64, 32])
@pytest.mark.parametrize("log_n", list(range(10, 17)))
def test_ntt_sweep_4096_polynomials(pkg, ctx, orc, log_n, bits):
    """configs[1]: N = 2^10..2^16, 4096 polynomials, q = first prime of two_adic_primes(55 | 28, log_n + 1)."""
    from learn_fhe_b200 import util
    n, batch = 1 << log_n, 4096
    q = orc.two_adic_primes(55 if bits == 64 else 28, log_n + 1, 1)[0]
    dt = torch.int64 if bits == 64 else torch.int32
    g = torch.Generator(device="cuda")
    g.manual_seed(0x5EED0000 + 2 * log_n + (bits == 32))
    a = torch.randint(0, q, (batch, n), device="cuda", generator=g)
    b = torch.randint(0, q, (batch, n), device="cuda", generator=g)
    a[0].zero_()            # NTT(0) = 0
    a[1].fill_(q - 1)       # largest residue everywhere
    a[2].zero_()
    a[2, 0] = 1             # NTT(1) = (1, ..., 1)
    s = (a + b) % q
    spot = [1, 3, 1234, batch - 1]
    host_a = a[spot].cpu().numpy().astype(np.uint64)
    a, b, s = a.to(dt), b.to(dt), s.to(dt)
    fa, fb, fs = a.clone(), b.clone(), s.clone()
    torch.cuda.synchronize()
    for t in (fa, fb, fs):
        util.ntt_fwd_dev(ctx, q, t, log_n, bits)
    ctx.sync()
    assert int(fa.min()) >= 0 and int(fa.max()) < q  # canonical residues
    assert torch.equal((fa.long() + fb.long()) % q, fs.long()), "forward transform is not additive on some polynomial"
    assert not fa[0].any() and bool((fa[2] == 1).all())
    ref = orc.ntt_fwd(q, host_a, threads=4)
    assert (fa[spot].cpu().numpy().astype(np.uint64) == ref).all()
    torch.cuda.synchronize()
    for t, src in ((fa, a), (fs, s)):
        util.ntt_inv_dev(ctx, q, t, log_n, bits)
        ctx.sync()
        assert torch.equal(t, src), "inverse(forward(x)) != x on some polynomial"


# ---------------------------------------------------------------------------------------------------------------- C1: FHEW gates
def test_fhew_16384_nand_gates(pkg, ctx, orc, fhew_setup):
    """configs[0] at the headline batch: 16384 NAND gates on real encryptions under the seeded FHEW-T key."""
    from learn_fhe_b200 import fhew
    P, K, ex = fhew_setup
    param = fhew.single_key_testing_param(P.big_q)
    bk = fhew.BootstrappingKey(ctx, param, ex["ksk_a"], ex["ksk_b"], ex["brk"], ex["ak"], ex["ak_t"])
    count = 16384
    rng = np.random.default_rng(77)
    bits = rng.integers(0, 2, size=2 * count).astype(np.int32)
    cts = K.encrypt(bits, 77)
    lin = (cts[:count] + cts[count:]) % np.uint64(P.big_q)
    table = [1, 1, 1, 0]
    out = fhew.Fhew.op(bk, table, lin)  # host buffers (the e2e path of bench.py)
    assert (K.decrypt(out) == 1 - (bits[:count] & bits[count:])).all()
    # device-resident entry point: same words
    d_in, d_f = pkg.to_dev(lin), pkg.to_dev(fhew.gate_poly(param, table))
    d_out = torch.empty_like(d_in)
    fhew.Bootstrapping.bootstrap_dev(bk, d_f, d_in, d_out, post_add=fhew.big_q_by_8(param))
    ctx.sync()
    assert (pkg.to_host(d_out) == out).all()
    # a ciphertext's result does not depend on its position in the batch or on its neighbours
    rev = fhew.Fhew.op(bk, table, np.ascontiguousarray(lin[::-1]))
    assert (rev[::-1] == out).all()
    spot = [0, 1, 4097, count - 1]
    assert (out[spot] == K.op(table, lin[spot], threads=4)).all()
    bk.free()


# ---------------------------------------------------------------------------------------------------------------- C3: TFHE PBS
def test_tfhe_16384_programmable_bootstraps(pkg, ctx, orc):
    """configs[2] at N = 2048 (tfhe/bootstrapping.rs:141-152 parameters): 16384 PBS with an arbitrary 16-entry table."""
    from learn_fhe_b200 import tfhe
    P = orc.tfhe_testing_param()
    K = orc.TfheKey(P, 0x5EED0003)
    ex = K.export()
    param = pkg.TfheParam(log_p=P.log_p, padding=P.padding, n=P.n, ks_log_b=P.ks_log_b, ks_d=P.ks_d,
                          log_big_n=P.big_n.bit_length() - 1, k=P.k, bs_log_b=P.bs_log_b, bs_d=P.bs_d)
    bk = tfhe.BootstrappingKey(ctx, param, ex["brk"], ex["ksk_a"], ex["ksk_b"])
    count, p = 16384, 1 << P.log_p
    rng = np.random.default_rng(78)
    msgs = rng.integers(0, p, size=count).astype(np.uint64)
    table = ((5 * np.arange(p) + 3) % p).astype(np.uint64)
    cts = K.encrypt(msgs, 78)
    v = K.lut_poly(table)
    lut = tfhe.encode_lut(bk.param, v)
    got = tfhe.Bootstrapping.bootstrap(bk, lut, cts)
    assert (K.decrypt(got)[0] == table[msgs.astype(np.int64)]).all()
    rev = tfhe.Bootstrapping.bootstrap(bk, lut, np.ascontiguousarray(cts[::-1]))
    assert (rev[::-1] == got).all()
    spot = [0, 8191, count - 1]
    assert (got[spot] == K.bootstrap(v, cts[spot], threads=3)).all()
    # the fused bounded-error modes (3 is the bench headline): every one of the 16384 outputs decrypts to the table, its phase stays
    # within 2^54 of the bit-identical mode's (decoding margin 2^58), and an output does not depend on its position in the batch
    ph_exact = K.decrypt(got)[1]
    for mode in (2, 3):
        bk.set_mode(mode)
        fused = tfhe.Bootstrapping.bootstrap(bk, lut, cts)
        m, ph = K.decrypt(fused)
        assert (m == table[msgs.astype(np.int64)]).all(), mode
        assert np.abs((ph - ph_exact).astype(np.int64)).max() < 2 ** 54, mode
        rev = tfhe.Bootstrapping.bootstrap(bk, lut, np.ascontiguousarray(cts[::-1]))
        assert (rev[::-1] == fused).all(), mode
    bk.free()


# ---------------------------------------------------------------------------------------------------------------- C4: CKKS mul
def test_ckks_512_pairs_n65536_full_chain(pkg, ctx, orc):
    """configs[3]: Ckks::mul at N = 2^16, log_qi = 55, L = 8 (+ 8 special primes), level 8 -> 7, 512 ciphertext pairs."""
    from learn_fhe_b200 import ckks
    log_n, big_l, count = 16, 8, 512
    K = orc.CkksKey(log_n, 55, big_l, 0x5EED0004)
    P = ckks.CkksParam(ctx, log_n, K.qs, K.ps)
    rlk = ckks.CkksKeySwitchingKey(P, K.ksk(-1))
    n = 1 << log_n
    g = torch.Generator(device="cuda")
    g.manual_seed(0x5EED0004)

    def rand_ct():
        t = torch.empty((count, 2, big_l, n), dtype=torch.int64, device="cuda")
        for i, q in enumerate(K.qs):
            t[:, :, i, :] = torch.randint(0, q, (count, 2, n), device="cuda", generator=g)
        return t

    ct0, ct1 = rand_ct(), rand_ct()
    rng = np.random.default_rng(79)
    real = {}
    for slot in (0, count - 1):  # real encryptions of small plaintexts in two slots
        e0 = K.encrypt(rng.integers(-(1 << 20), 1 << 20, size=n, dtype=np.int64), big_l, 900 + slot)
        e1 = K.encrypt(rng.integers(-(1 << 20), 1 << 20, size=n, dtype=np.int64), big_l, 950 + slot)
        real[slot] = (e0, e1)
        ct0[slot] = pkg.to_dev(e0)
        ct1[slot] = pkg.to_dev(e1)
    ct0[7], ct1[7] = ct0[0], ct1[0]  # the same pair at another batch position
    out = torch.empty((count, 2, big_l - 1, n), dtype=torch.int64, device="cuda")
    swapped = torch.empty_like(out)
    torch.cuda.synchronize()
    ckks.Ckks.mul_dev(P, rlk, big_l, ct0, ct1, out)
    ckks.Ckks.mul_dev(P, rlk, big_l, ct1, ct0, swapped)
    ctx.sync()
    for i, q in enumerate(K.qs[:big_l - 1]):
        assert int(out[:, :, i, :].min()) >= 0 and int(out[:, :, i, :].max()) < q
    assert torch.equal(out, swapped), "Ckks::mul is not commutative on some pair"
    assert torch.equal(out[7], out[0])
    for slot, (e0, e1) in real.items():
        ref = K.mul(e0, e1)
        got = pkg.to_host(out[slot].contiguous())
        assert (got == ref).all(), slot
        assert (K.decrypt(got) == K.decrypt(ref)).all()
    rlk.free()
    P.free()
