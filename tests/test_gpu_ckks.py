"""Parity of the RNS / CKKS CUDA path with the oracle (restating util/src/ring/rns.rs:373-386 and the integer part of
scheme/ckks/src/ckks.rs:303-415: every limb of every output bit-identical; decryptions identical)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _rns_random(orc, seed, moduli, batch, n):
    x = np.zeros((batch, len(moduli), n), dtype=np.uint64)
    for b in range(batch):
        for i, q in enumerate(moduli):
            x[b, i] = orc.residues(seed + 131 * b + i, n, q)
    return x


@pytest.mark.parametrize("log_n", [0, 3, 9, 12])
def test_extend_bases_matches_oracle(pkg, ctx, orc, log_n):
    """rns.rs:373-386 shape: 8 -> 16 primes of 55 bits; plus ragged base sizes."""
    from learn_fhe_b200 import ckks
    n = 1 << log_n
    primes = orc.two_adic_primes(55, 13, 16)
    for nq, np_ in ((8, 8), (1, 3), (3, 1), (16, 2)):
        qs = primes[:nq]
        ps = primes[nq:nq + np_] if nq < 16 else orc.two_adic_primes(54, 13, np_)
        x = _rns_random(orc, 7 + log_n, qs, 3, n)
        x[0, :, 0] = 0
        x[0, :, n - 1] = [q - 1 for q in qs]
        got = ckks.extend_bases(ctx, qs, ps, x)
        for b in range(3):
            assert (got[b] == orc.rns_extend_bases(qs, ps, x[b])).all(), (log_n, nq, np_, b)


@pytest.mark.parametrize("log_n", [0, 4, 10])
def test_rescale_k_matches_oracle(pkg, ctx, orc, log_n):
    from learn_fhe_b200 import ckks
    n = 1 << log_n
    primes = orc.two_adic_primes(55, 11, 12)
    for nq, k in ((8, 1), (2, 1), (12, 4), (9, 8), (3, 2)):
        qs = primes[:nq]
        x = _rns_random(orc, 1000 + log_n, qs, 2, n)
        got = ckks.rescale_k(ctx, qs, k, x)
        for b in range(2):
            assert (got[b] == orc.rns_rescale_k(qs, k, x[b])).all(), (log_n, nq, k, b)


def test_rns_argument_errors(pkg, ctx, orc):
    """duplicate moduli panic in the reference (rns.rs:84) -> FHE_EINVAL; k out of range (rns.rs:104)."""
    from learn_fhe_b200 import ckks
    p = orc.two_adic_primes(55, 5, 3)
    x = np.zeros((1, 2, 16), dtype=np.uint64)
    with pytest.raises(pkg.FheError):
        ckks.extend_bases(ctx, p[:2], [p[0]], x)
    with pytest.raises(pkg.FheError):
        ckks.rescale_k(ctx, p[:2], 2, x)


@pytest.fixture(scope="module", params=[(1, 8), (4, 3), (6, 8), (9, 4), (12, 3)])
def ckks_setup(request, pkg, ctx, orc):
    from learn_fhe_b200 import ckks
    log_n, big_l = request.param
    K = orc.CkksKey(log_n, 55, big_l, 0x5EED0004, auto_ts=(5, -1))
    P = ckks.CkksParam(ctx, log_n, K.qs, K.ps)
    rlk = ckks.CkksKeySwitchingKey(P, K.ksk(-1))
    yield K, P, rlk
    rlk.free()
    P.free()


def _small_pt(seed, n, bound=1 << 20):
    rng = np.random.default_rng(seed)
    return rng.integers(-bound, bound, size=n, dtype=np.int64)


def test_mul_relin_rescale_matches_oracle(pkg, ctx, orc, ckks_setup):
    from learn_fhe_b200 import ckks
    K, P, rlk = ckks_setup
    for level in range(P.big_l, 1, -1):
        count = 3
        ct0 = np.stack([K.encrypt(_small_pt(10 * level + i, K.n), level, 100 + i) for i in range(count)])
        ct1 = np.stack([K.encrypt(_small_pt(20 * level + i, K.n), level, 200 + i) for i in range(count)])
        ref = K.mul(ct0, ct1, threads=3)
        got = ckks.Ckks.mul(P, rlk, ct0, ct1)
        assert got.shape == ref.shape
        assert (got == ref).all(), ("mul", P.log_n, level)
        assert (K.decrypt(got[0]) == K.decrypt(ref[0])).all()


def test_host_paths_pipelined_over_ragged_chunks(pkg, ctx, orc, ckks_setup, monkeypatch):
    """The host-slice entry points overlap H2D / kernels / D2H over chunks of the batch (two copy streams + events): any chunk
    count, including ones that do not divide the batch, gives the words of the one-chunk path (Ckks::mul and the NTT)."""
    from learn_fhe_b200 import ckks, util
    K, P, rlk = ckks_setup
    level, count = P.big_l, 7
    ct0 = np.stack([K.encrypt(_small_pt(300 + i, K.n), level, 400 + i) for i in range(count)])
    ct1 = np.stack([K.encrypt(_small_pt(500 + i, K.n), level, 600 + i) for i in range(count)])
    ref = K.mul(ct0, ct1, threads=4)
    q = K.qs[0]
    a = np.stack([orc.residues(900 + i, K.n, q) for i in range(11)])
    want = a.copy()
    monkeypatch.setenv("FHE_B200_HOST_CHUNKS", "1")
    util.nega_cyclic_ntt_in_place(ctx, q, want)
    for chunks in ("1", "2", "3", "7", "64"):
        monkeypatch.setenv("FHE_B200_HOST_CHUNKS", chunks)
        assert (ckks.Ckks.mul(P, rlk, ct0, ct1) == ref).all(), chunks
        x = a.copy()
        util.nega_cyclic_ntt_in_place(ctx, q, x)
        assert (x == want).all(), chunks
        util.nega_cyclic_intt_in_place(ctx, q, x)
        assert (x == a).all(), chunks


def test_mul_chain_all_levels(pkg, ctx, orc, ckks_setup):
    """ckks.rs:378-398: L-1 chained multiplications, every intermediate ciphertext bit-identical."""
    from learn_fhe_b200 import ckks
    K, P, rlk = ckks_setup
    level = P.big_l
    g = K.encrypt(_small_pt(1, K.n, 4), level, 1)[None]
    r = g.copy()
    other = K.encrypt(_small_pt(2, K.n, 4), level, 2)[None]
    while level > 1:
        g = ckks.Ckks.mul(P, rlk, g, other[:, :, :level])
        r = K.mul(r, other[:, :, :level])
        assert (g == r).all(), level
        level -= 1


def test_key_switch_rotate_conjugate_rescale(pkg, ctx, orc, ckks_setup):
    from learn_fhe_b200 import ckks
    K, P, rlk = ckks_setup
    for level in (P.big_l, 1):
        cts = np.stack([K.encrypt(_small_pt(40 + i, K.n), level, 300 + i) for i in range(2)])
        # plain key switch with the relinearisation key (no automorphism)
        got = ckks.Ckks.key_switch(P, rlk, cts, 0)
        for i in range(2):
            assert (got[i] == K.key_switch(-1, cts[i], apply_auto=False)).all()
        for which, t in enumerate(K.auto_ts):
            ksk = ckks.CkksKeySwitchingKey(P, K.ksk(which))
            got = ckks.Ckks.key_switch(P, ksk, cts, t)
            for i in range(2):
                assert (got[i] == K.key_switch(which, cts[i], apply_auto=True)).all(), (level, t)
            ksk.free()
    cts = np.stack([K.encrypt(_small_pt(50 + i, K.n), P.big_l, 400 + i) for i in range(2)])
    got = ckks.Ckks.rescale(P, cts)
    for i in range(2):
        assert (got[i] == K.rescale(cts[i])).all()


def test_mul_constant_matches_oracle(pkg, ctx, orc, ckks_setup):
    """ckks.rs:250-253 on an encoded plaintext: limb-wise negacyclic products with the ciphertext halves, then rescale."""
    from learn_fhe_b200 import ckks
    K, P, rlk = ckks_setup
    level = P.big_l
    cts = np.stack([K.encrypt(_small_pt(60 + i, K.n), level, 500 + i) for i in range(3)])
    for per_ct in (False, True):
        npt = 3 if per_ct else 1
        pts = np.stack([np.stack([np.mod(_small_pt(70 + j, K.n, 1 << 30), q).astype(np.uint64) for q in P.qs[:level]]) for j in range(npt)])
        got = ckks.Ckks.mul_constant(P, pts if per_ct else pts[0], cts)
        for i in range(3):
            pt = pts[i if per_ct else 0]
            prod = np.stack([np.stack([orc.ntt_mul(P.qs[t], cts[i, h, t], pt[t]) for t in range(level)]) for h in range(2)])
            assert (got[i] == K.rescale(prod)).all(), (per_ct, i)


def test_mul_mat_bsgs_matches_oracle(pkg, ctx, orc, ckks_setup):
    """Bootstrapping::mul_mat (scheme/ckks/src/bootstrapping.rs:92-108) with a 2 x 2 BSGS plan (baby rotations {0, 5}, giant
    rotations {0, -1 (conjugation key stands in for a second rotation key)}, one absent diagonal), random encoded diagonals:
    equal, limb for limb, to the composition of the oracle's rotate / mul_constant / add in the reference's order."""
    from learn_fhe_b200 import ckks
    K, P, rlk = ckks_setup
    level = P.big_l
    q = lambda lv: [np.uint64(x) for x in P.qs[:lv]]
    cts = np.stack([K.encrypt(_small_pt(80 + i, K.n), level, 600 + i) for i in range(2)])
    keys = [ckks.CkksKeySwitchingKey(P, K.ksk(w)) for w in range(len(K.auto_ts))]
    baby = [(0, None), (K.auto_ts[0], keys[0])]
    giant = [(0, None), (K.auto_ts[1], keys[1])]
    present = np.array([[1, 1], [0, 1]], dtype=np.uint8)
    pts = np.stack([np.stack([np.mod(_small_pt(90 + j, K.n, 1 << 30), int(m)).astype(np.uint64) for m in P.qs[:level]]) for j in range(3)])
    got = ckks.Ckks.mul_mat(P, baby, giant, present, pts, cts)

    def rns_add(a, b, lv):
        return np.stack([np.stack([(a[h, t] + b[h, t]) % q(lv)[t] for t in range(lv)]) for h in range(2)])

    def mul_constant(pt, ct):
        prod = np.stack([np.stack([orc.ntt_mul(P.qs[t], ct[h, t], pt[t]) for t in range(level)]) for h in range(2)])
        return K.rescale(prod)

    for c in range(2):
        rot = [cts[c], K.key_switch(0, cts[c], apply_auto=True)]
        inner0 = rns_add(mul_constant(pts[0], rot[0]), mul_constant(pts[1], rot[1]), level - 1)
        inner1 = mul_constant(pts[2], rot[1])
        ref = rns_add(inner0, K.key_switch(1, inner1, apply_auto=True), level - 1)
        assert (got[c] == ref).all(), c
    for k in keys:
        k.free()


def test_ckks_level_errors(pkg, ctx, orc, ckks_setup):
    from learn_fhe_b200 import ckks
    K, P, rlk = ckks_setup
    ct = np.zeros((1, 2, 1, K.n), dtype=np.uint64)
    with pytest.raises(pkg.FheError):
        ckks.Ckks.mul(P, rlk, ct, ct)  # level 1 has no limb to drop
