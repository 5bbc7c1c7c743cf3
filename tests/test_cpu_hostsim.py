"""CPU tier: the CUDA kernels' per-thread logic (the __host__ __device__ headers of learn-fhe_b200/csrc compiled with
g++ and replayed sequentially by tests/hostsim) against the oracle — the closest a GPU-less box gets to the kernels."""
import ctypes as C
import json
import os

import numpy as np
import pytest

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(dtype=np.uint32, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


class SimParam(C.Structure):
    _fields_ = [("log_n", C.c_uint), ("big_q", C.c_uint64), ("p", C.c_uint64), ("rlwe_log_b", C.c_uint), ("rlwe_d", C.c_uint),
                ("rgsw_log_b", C.c_uint), ("rgsw_d", C.c_uint), ("n_s", C.c_uint), ("q_ks", C.c_uint64), ("ks_log_b", C.c_uint),
                ("ks_d", C.c_uint), ("w", C.c_uint)]


@pytest.fixture(scope="module")
def H(hostsim):
    hostsim.sim_ntt_u64.argtypes = [C.c_uint64, C.c_uint, C.c_int, u64p, C.c_int, C.c_uint]
    hostsim.sim_ntt_u32.argtypes = [C.c_uint64, C.c_uint, C.c_int, u32p, C.c_int, C.c_uint]
    hostsim.sim_fhew_key_upload.restype = C.c_void_p
    hostsim.sim_fhew_key_upload.argtypes = [C.POINTER(SimParam), u64p, u64p, u64p, u64p, i64p]
    hostsim.sim_fhew_key_free.argtypes = [C.c_void_p]
    hostsim.sim_fhew_prologue.argtypes = [C.c_void_p, u64p, u64p, C.c_int, C.c_int, C.c_uint]
    hostsim.sim_fhew_step.argtypes = [C.c_void_p, C.c_uint, u64p, u64p, C.c_uint]
    hostsim.sim_fhew_blind_rotate_extract.argtypes = [C.c_void_p, u64p, u64p, C.c_uint64, u64p, u64p, C.c_uint]
    hostsim.sim_mul_u64.restype = C.c_uint64
    hostsim.sim_mul_u64.argtypes = [C.c_uint64] * 3
    hostsim.sim_mul_u32.restype = C.c_uint32
    hostsim.sim_mul_u32.argtypes = [C.c_uint32] * 3
    hostsim.sim_swz32.restype = C.c_uint
    hostsim.sim_swz64.restype = C.c_uint
    return hostsim


def sim_key(H, P, ex):
    sp = SimParam(*[getattr(P, f[0]) for f in SimParam._fields_])
    h = H.sim_fhew_key_upload(C.byref(sp), ex["ksk_a"].reshape(-1), ex["ksk_b"].reshape(-1), ex["brk"].reshape(-1),
                              ex["ak"].reshape(-1), np.ascontiguousarray(ex["ak_t"], dtype=np.int64))
    assert h
    return h


def test_modmul_primitives(H, orc):
    rng = np.random.default_rng(0)
    for q in (orc.two_adic_primes(61, 5, 1)[0], orc.two_adic_primes(55, 12, 1)[0], 268409857, (1 << 62) - 57, 3):
        for _ in range(300):
            a, b = int(rng.integers(0, q)), int(rng.integers(0, q))
            assert H.sim_mul_u64(q, a, b) == a * b % q
        assert H.sim_mul_u64(q, q - 1, q - 1) == (q - 1) * (q - 1) % q
    for q in (268409857, (1 << 30) - 35, 12289):
        for _ in range(300):
            a, b = int(rng.integers(0, q)), int(rng.integers(0, q))
            assert H.sim_mul_u32(q, a, b) == a * b % q


def test_swizzle_is_a_permutation(H):
    for fn, nmax in ((H.sim_swz32, 1 << 13), (H.sim_swz64, 1 << 13)):
        assert sorted(fn(p) for p in range(nmax)) == list(range(nmax))


@pytest.mark.parametrize("log_n", list(range(0, 17)))
def test_kernel_logic_ntt(H, orc, log_n):
    n = 1 << log_n
    for bits, fn, dt in ((55, H.sim_ntt_u64, np.uint64), (28, H.sim_ntt_u32, np.uint32), (61, H.sim_ntt_u64, np.uint64),
                         (30, H.sim_ntt_u32, np.uint32)):
        if log_n == 0 and bits == 61:
            continue
        q = orc.two_adic_primes(bits, log_n + 1, 1)[0]
        a = orc.residues(log_n, n, q)
        ref = orc.ntt_fwd(q, a)
        for c in ([log_n] if log_n <= 13 else [12, 13]):
            x = a.astype(dt).copy()
            assert fn(q, log_n, c, x, 1, 37) == 0 and (x.astype(np.uint64) == ref).all(), (log_n, bits, c, "fwd")
            assert fn(q, log_n, c, x, 0, 64) == 0 and (x.astype(np.uint64) == a).all(), (log_n, bits, c, "inv")


def test_kernel_logic_fhew(H, orc, fhew_setup):
    P, K, ex = fhew_setup
    h = sim_key(H, P, ex)
    bits = np.array([0, 0, 1, 1, 0, 1, 0, 1], dtype=np.int32)
    cts = K.encrypt(bits, 7)
    lin = (cts[:4] + cts[4:]) % np.uint64(P.big_q)
    pro = K.prologue(lin)
    for i in range(4):
        out = np.zeros(P.n_s + 1, dtype=np.uint64)
        H.sim_fhew_prologue(h, lin[i], out, 1, 1, 128)
        assert (out == pro[i]).all()
    acc = orc.residues(9, 2 * P.n, P.big_q).reshape(2, P.n)
    out = np.zeros_like(acc)
    for j in (0, 5, 99):
        H.sim_fhew_step(h, j, acc.reshape(-1), out.reshape(-1), 128)
        assert (out == K.external_product(j, acc)).all()
    for v in (0, 1, 10):
        H.sim_fhew_step(h, 0x8000 | v, acc.reshape(-1), out.reshape(-1), 96)
        assert (out == K.automorphism(v, acc)).all()
    f = orc.fhew_gate_poly(P, [1, 1, 1, 0])
    q8 = int(round(P.big_q / 8.0))
    ref = K.op([1, 1, 1, 0], lin[:2], threads=2)
    for i in range(2):
        o = np.zeros(P.n + 1, dtype=np.uint64)
        a = np.zeros(2 * P.n, dtype=np.uint64)
        assert H.sim_fhew_blind_rotate_extract(h, f, pro[i], q8, o, a, 128) == 0
        assert (o == ref[i]).all()
        assert (a.reshape(2, -1) == K.blind_rotate(f, pro[i])).all()
    H.sim_fhew_key_free(h)


@pytest.mark.parametrize("bits", [55, 60])
def test_kernel_logic_fhew_64bit_modulus(H, orc, bits):
    """The generic FHEW kernels instantiated for a 64-bit modulus (fhew.cu `wide`: the parameter shape of
    examples/multi_key_uint8.rs:15-29 - 55-bit Q, decomposors (11, 5), LWE q = 2^20 - at N = 64): single steps, the LWE
    prologue and whole gate bootstraps against the oracle, bit-exact.  55 bits: lazy butterflies of the fast NTT path
    (Q < 2^56); 60 bits: the generic per-butterfly reductions."""
    P = orc.fhew_testing_param()
    P.log_n, P.big_q = 6, orc.two_adic_primes(bits, 7, 1)[0]
    P.rlwe_log_b = P.rgsw_log_b = 11
    P.rlwe_d = P.rgsw_d = 5
    P.n_s, P.q_ks, P.ks_log_b, P.ks_d, P.w = 12, 1 << 20, 4, 5, 10
    K = orc.FhewKey(P, 0x5EED000B)
    ex = K.export()
    sp = SimParam(*[getattr(P, f[0]) for f in SimParam._fields_])
    H.sim_fhew64_key_upload.restype = C.c_void_p
    H.sim_fhew64_key_upload.argtypes = [C.POINTER(SimParam), u64p, u64p, u64p, u64p, i64p]
    H.sim_fhew64_key_free.argtypes = [C.c_void_p]
    H.sim_fhew64_step.argtypes = [C.c_void_p, C.c_uint, u64p, u64p, C.c_uint]
    H.sim_fhew64_prologue.argtypes = [C.c_void_p, u64p, u64p, C.c_int, C.c_int, C.c_uint]
    H.sim_fhew64_blind_rotate_extract.argtypes = [C.c_void_p, u64p, u64p, C.c_uint64, u64p, u64p, C.c_uint]
    U = lambda v: np.ascontiguousarray(v, dtype=np.uint64).reshape(-1)
    ak_t = np.ascontiguousarray(ex["ak_t"], dtype=np.int64)
    h = H.sim_fhew64_key_upload(C.byref(sp), U(ex["ksk_a"]), U(ex["ksk_b"]), U(ex["brk"]), U(ex["ak"]), ak_t)
    assert h
    acc = orc.residues(29, 2 * P.n, P.big_q).reshape(2, P.n)
    acc[0, :3] = [0, P.big_q - 1, P.big_q // 2]
    out = np.zeros_like(acc)
    for j in (0, 5, 11):
        assert H.sim_fhew64_step(h, j, acc.reshape(-1), out.reshape(-1), 48) == 0
        assert (out == K.external_product(j, acc)).all(), j
    for v in (0, 1, 10):
        assert H.sim_fhew64_step(h, 0x8000 | v, acc.reshape(-1), out.reshape(-1), 48) == 0
        assert (out == K.automorphism(v, acc)).all(), v
    bits = np.array([0, 0, 1, 1, 0, 1, 0, 1], dtype=np.int32)
    cts = K.encrypt(bits, 7)
    lin = (cts[:4] + cts[4:]) % np.uint64(P.big_q)
    pro = K.prologue(lin)
    f = orc.fhew_gate_poly(P, [1, 1, 1, 0])
    q8 = int(round(P.big_q / 8.0))
    ref = K.op([1, 1, 1, 0], lin, threads=2)
    for i in range(4):
        o = np.zeros(P.n_s + 1, dtype=np.uint64)
        assert H.sim_fhew64_prologue(h, lin[i], o, 1, 1, 32) == 0
        assert (o == pro[i]).all()
        o = np.zeros(P.n + 1, dtype=np.uint64)
        a = np.zeros(2 * P.n, dtype=np.uint64)
        assert H.sim_fhew64_blind_rotate_extract(h, f, pro[i], q8, o, a, 40) == 0
        assert (o == ref[i]).all()
    assert (K.decrypt(ref) == 1 - (bits[:4] & bits[4:])).all()
    H.sim_fhew64_key_free(h)


def test_kernel_logic_fhew_golden_tiny(H, orc):
    """The kernel logic on the committed tiny-parameter fixture (produced by pyref in the reference dataflow)."""
    from test_cpu_oracle import golden_fhew_tiny
    g, P, keys = golden_fhew_tiny(orc)
    ex = dict(ksk_a=keys[0], ksk_b=keys[1], brk=keys[2], ak=keys[3], ak_t=keys[4])
    h = sim_key(H, P, ex)
    f = np.array(g["f"], dtype=np.uint64)
    for c in g["cases"]:
        pro = np.zeros(P.n_s + 1, dtype=np.uint64)
        H.sim_fhew_prologue(h, np.array(c["ct"], dtype=np.uint64), pro, 1, 1, 32)
        assert [int(x) for x in pro] == c["prologue"]
        o = np.zeros(P.n + 1, dtype=np.uint64)
        a = np.zeros(2 * P.n, dtype=np.uint64)
        assert H.sim_fhew_blind_rotate_extract(h, f, pro, g["post_add"], o, a, 32) == 0
        assert [int(x) for x in o] == c["out"]
    H.sim_fhew_key_free(h)


@pytest.mark.parametrize("log_n", list(range(9, 18)))
def test_kernel_logic_ntt_fast(H, orc, log_n):
    """Second-generation (lazy-reduction) NTT passes of ntt_fast.cuh, replayed with the launcher's geometry."""
    H.sim_ntt_fast_u64.argtypes = [C.c_uint64, C.c_uint, C.c_int, u64p, C.c_int]
    H.sim_ntt_fast_u32.argtypes = [C.c_uint64, C.c_uint, C.c_int, u32p, C.c_int]
    n = 1 << log_n
    cases = [(55, H.sim_ntt_fast_u64, np.uint64), (28, H.sim_ntt_fast_u32, np.uint32)]
    if log_n <= 13:
        cases += [(40, H.sim_ntt_fast_u64, np.uint64), (20, H.sim_ntt_fast_u32, np.uint32)]
    for bits, fn, dt in cases:
        q = orc.two_adic_primes(bits, log_n + 1, 1)[0]
        a = orc.residues(100 + log_n, n, q)
        a[:4] = [0, q - 1, 1, q - 1]  # extremes
        ref = orc.ntt_fwd(q, a)
        logts = [-1] if log_n <= 13 or log_n == 17 else [-1, 13]
        for logt in logts:
            x = a.astype(dt).copy()
            assert fn(q, log_n, logt, x, 1) == 0 and (x.astype(np.uint64) == ref).all(), (log_n, bits, logt, "fwd")
            assert fn(q, log_n, logt, x, 0) == 0 and (x.astype(np.uint64) == a).all(), (log_n, bits, logt, "inv")


def test_lz64_one_multiply_barrett(H):
    """Lz64::barrett2 (one 32-bit high multiply): for moduli of every bit length 3..56 and lazy values up to the
    path's largest bound 256 q (edges included) the result is congruent and below 2q, and canon() is exact."""
    rng = np.random.default_rng(11)
    H.sim_lz64_barrett2.argtypes = [C.c_uint64, C.c_void_p, C.c_size_t]
    for k in range(3, 57):
        for q in {(1 << k) - 1, (1 << (k - 1)) + 1, int(rng.integers((1 << (k - 1)) + 1, 1 << k)) | 1}:
            edge = [0, 1, q - 1, q, q + 1, 2 * q - 1, 2 * q, 255 * q, 256 * q - 1, 128 * q, 128 * q - 1]
            mult = rng.integers(0, 256, size=2000, dtype=np.uint64) * np.uint64(q)
            xs = np.concatenate([np.array(edge, dtype=np.uint64), mult + rng.integers(0, q, size=2000, dtype=np.uint64),
                                 mult[:50], mult[:50] + np.uint64(q - 1)])
            assert int(xs.max()) < 256 * q
            assert H.sim_lz64_barrett2(q, xs.ctypes.data, xs.size) == 0, (k, q)


def test_fast_swizzle(H):
    H.sim_swz2_32.restype = C.c_uint
    H.sim_swz2_64.restype = C.c_uint
    for fn, lanes_per_phase, words_per_bank in ((H.sim_swz2_32, 32, 1), (H.sim_swz2_64, 16, 1)):
        assert sorted(fn(p) for p in range(1 << 13)) == list(range(1 << 13))
        # linear over XOR
        for a, b in ((5, 64), (0x1234, 0x0F00), (777, 8)):
            assert fn(a ^ b) == fn(a) ^ fn(b)
        # radix-8 pass at stride 8 (LL = 3): lanes = consecutive groups; every j must be bank-conflict-free per phase
        nb = lanes_per_phase
        for j in range(8):
            for first in (0, lanes_per_phase, 5 * lanes_per_phase):
                banks = set()
                for lane in range(lanes_per_phase):
                    g = first + lane
                    lo, hi = g & 7, g >> 3
                    banks.add(fn((hi << 6) | (j << 3) | lo) % nb)
                assert len(banks) == lanes_per_phase, (fn, j, first)


def test_kernel_logic_rns(H, orc):
    """rns_core.cuh (extend_bases / rescale_k per-coefficient device logic) against the oracle (rns.rs:83-132, 331-345)."""
    H.sim_rns_extend.argtypes = [u64p, C.c_size_t, u64p, C.c_size_t, u64p, u64p, C.c_size_t]
    H.sim_rns_rescale.argtypes = [u64p, C.c_size_t, C.c_size_t, u64p, u64p, C.c_size_t]
    primes = orc.two_adic_primes(55, 10, 16)
    n = 64
    U = lambda v: np.ascontiguousarray(v, dtype=np.uint64)
    for nq, np_ in ((8, 8), (1, 2), (5, 1), (12, 4)):
        qs, ps = primes[:nq], primes[nq:nq + np_]
        x = np.stack([orc.residues(3 * nq + i, n, q) for i, q in enumerate(qs)])
        x[:, 0] = 0
        x[:, 1] = [q - 1 for q in qs]
        out = np.zeros((np_, n), dtype=np.uint64)
        assert H.sim_rns_extend(U(qs), nq, U(ps), np_, x.reshape(-1), out.reshape(-1), n) == 0
        assert (out == orc.rns_extend_bases(qs, ps, x)[nq:]).all(), (nq, np_)
    for nq, k in ((8, 1), (2, 1), (16, 8), (9, 8), (3, 2)):
        qs = primes[:nq]
        x = np.stack([orc.residues(7 * nq + i, n, q) for i, q in enumerate(qs)])
        out = np.zeros((nq - k, n), dtype=np.uint64)
        assert H.sim_rns_rescale(U(qs), nq, k, x.reshape(-1), out.reshape(-1), n) == 0
        assert (out == orc.rns_rescale_k(qs, k, x)).all(), (nq, k)
    # Ckks::mul's two consecutive rescales fused into one pass (rns_rescale2_coeff)
    H.sim_rns_rescale2.argtypes = [u64p, C.c_size_t, C.c_size_t, u64p, u64p, C.c_size_t]
    for nq, k in ((16, 8), (6, 4), (4, 2), (11, 3)):
        qs = primes[:nq]
        x = np.stack([orc.residues(11 * nq + i, n, q) for i, q in enumerate(qs)])
        x[:, 0] = 0
        x[:, 1] = [q - 1 for q in qs]
        out = np.zeros((nq - k - 1, n), dtype=np.uint64)
        assert H.sim_rns_rescale2(U(qs), nq, k, x.reshape(-1), out.reshape(-1), n) == 0
        assert (out == orc.rns_rescale_k(qs[:nq - k], 1, orc.rns_rescale_k(qs, k, x))).all(), (nq, k)
    # a wider last kept limb: the canonical (non-lazy) finishing path
    qs = primes[:3] + orc.two_adic_primes(58, 10, 1) + primes[3:6]
    x = np.stack([orc.residues(900 + i, n, q) for i, q in enumerate(qs)])
    x[:, 1] = [q - 1 for q in qs]
    out = np.zeros((3, n), dtype=np.uint64)
    assert H.sim_rns_rescale2(U(qs), 7, 3, x.reshape(-1), out.reshape(-1), n) == 0
    assert (out == orc.rns_rescale_k(qs[:4], 1, orc.rns_rescale_k(qs, 3, x))).all()
    # the three accumulation modes of rns_extend_coeff at the largest fan-in (16 source limbs, all residues q_i - 1):
    # exact 128-bit sum (targets in [2^33, 2^59), incl. primes just below 2^59), lazy Shoup sum (forced), canonical (61-bit)
    big = orc.two_adic_primes(59, 8, 16)
    for targets, env in ((orc.two_adic_primes(59, 8, 18)[16:], None), (primes[:2], "1"), (orc.two_adic_primes(61, 8, 2), None)):
        x = np.stack([orc.residues(500 + i, n, q) for i, q in enumerate(big)])
        x[:, :8] = np.array([[q - 1] * 8 for q in big], dtype=np.uint64)
        out = np.zeros((2, n), dtype=np.uint64)
        if env:
            os.environ["FHE_B200_RNS_NO_WIDE"] = env
        try:
            assert H.sim_rns_extend(U(big), 16, U(targets), 2, x.reshape(-1), out.reshape(-1), n) == 0
        finally:
            os.environ.pop("FHE_B200_RNS_NO_WIDE", None)
        assert (out == orc.rns_extend_bases(big, targets, x)[16:]).all(), targets
    # mixed-width moduli (source residues larger than the target modulus)
    qs, ps = orc.two_adic_primes(55, 8, 3), orc.two_adic_primes(30, 8, 2) + orc.two_adic_primes(40, 8, 1)
    x = np.stack([orc.residues(99 + i, n, q) for i, q in enumerate(qs)])
    out = np.zeros((3, n), dtype=np.uint64)
    assert H.sim_rns_extend(U(qs), 3, U(ps), 3, x.reshape(-1), out.reshape(-1), n) == 0
    assert (out == orc.rns_extend_bases(qs, ps, x)[3:]).all()


def _tfhe_small_param(orc, n=24, big_n=64, k=1, bs_log_b=8, bs_d=3, ks_log_b=4, ks_d=5):
    P = orc.tfhe_testing_param()
    P.n, P.big_n, P.k, P.bs_log_b, P.bs_d, P.ks_log_b, P.ks_d = n, big_n, k, bs_log_b, bs_d, ks_log_b, ks_d
    return P


def test_kernel_logic_fft64_mul(H, orc):
    """tfhe_core.cuh f64 FFT product (twist, radix passes, untwist, f64_mod_u64) is bit-identical to the oracle's restatement
    of util/src/ring/fft/c64.rs:11-108 — full-range torus words times signed digits, the case of c64.rs:186-208."""
    H.sim_fft64_mul.argtypes = [u64p, u64p, C.c_uint, C.c_uint]
    for log_n in range(1, 12):
        n = 1 << log_n
        a = orc.splitmix64(31 + log_n, n)
        digits = (orc.splitmix64(77 + log_n, n) % np.uint64(1 << 17)).astype(np.int64) - (1 << 16)
        b = digits.astype(np.uint64)
        ref = orc.fft64_mul(a, b)
        x = a.copy()
        assert H.sim_fft64_mul(x, b, log_n, 37) == 0
        assert (x == ref).all(), log_n
        # small operands: exact == schoolbook (c64.rs:169-184)
        sa = (orc.splitmix64(5 + log_n, n) % np.uint64(64)).astype(np.uint64)
        sb = (orc.splitmix64(6 + log_n, n) % np.uint64(64)).astype(np.uint64)
        y = sa.copy()
        H.sim_fft64_mul(y, sb, log_n, 8)
        assert (y == orc.schoolbook_t64(sa, sb)).all()


@pytest.mark.parametrize("k,bs_d,big_n,n", [(1, 1, 64, 24), (1, 3, 64, 24), (2, 2, 64, 24), (1, 1, 2048, 5), (1, 2, 512, 6), (1, 1, 16, 6)])
def test_kernel_logic_tfhe(H, orc, k, bs_d, big_n, n):
    """TGGSW external product and the CMUX blind rotation (+ sample extract) of tfhe_core.cuh against the oracle
    (tggsw.rs:100-121, tfhe/bootstrapping.rs:84-104), reduced parameters, every torus word bit-identical."""
    H.sim_tfhe_key.restype = C.c_void_p
    H.sim_tfhe_key.argtypes = [C.c_uint] * 5 + [u64p]
    H.sim_tfhe_key_free.argtypes = [C.c_void_p]
    H.sim_tfhe_external_product.argtypes = [C.c_void_p, C.c_uint, u64p, u64p, C.c_uint]
    H.sim_tfhe_blind_rotate_extract.argtypes = [C.c_void_p, u64p, u64p, u64p, C.c_uint]
    P = _tfhe_small_param(orc, n=n, big_n=big_n, k=k, bs_d=bs_d, bs_log_b=23 if bs_d == 1 else 8)
    K = orc.TfheKey(P, 0x5EED0003)
    ex = K.export()
    log_n = P.big_n.bit_length() - 1
    h = H.sim_tfhe_key(log_n, P.k, P.n, P.bs_log_b, P.bs_d, ex["brk"].reshape(-1))
    glwe = orc.splitmix64(9, (P.k + 1) * P.big_n).reshape(P.k + 1, P.big_n)
    for i in (0, P.n // 2, P.n - 1):
        out = np.zeros_like(glwe)
        H.sim_tfhe_external_product(h, i, glwe.reshape(-1), out.reshape(-1), 48)
        assert (out == K.external_product(i, glwe)).all(), i
    cts = K.encrypt(np.arange(3, dtype=np.uint64), 5)
    cts[0, 3] = 0  # a zero mask word: the kernel skips that CMUX
    v = K.lut_poly(np.arange(1 << P.log_p, dtype=np.uint64))
    lut = (v << np.uint64(64 - (P.log_p + P.padding))).astype(np.uint64)
    for ct in cts:
        out = np.zeros(P.k * P.big_n + 1, dtype=np.uint64)
        H.sim_tfhe_blind_rotate_extract(h, lut, ct, out, 40)
        assert (out == K.blind_rotate_extract(v, ct)).all()
    H.sim_tfhe_key_free(h)


def test_kernel_logic_fhew_fast(H, orc, fhew_setup):
    """Fast blind-rotation path (fhew_fast.cuh: 3+4+2 passes, MAC fused with the first inverse pass) against the oracle:
    single steps and whole gate bootstraps, bit-exact."""
    P, K, ex = fhew_setup
    H.sim_fhew_fast_step.argtypes = [C.c_void_p, C.c_uint, u64p, u64p]
    H.sim_fhew_fast_blind_rotate_extract.argtypes = [C.c_void_p, u64p, u64p, C.c_uint64, u64p, u64p]
    h = sim_key(H, P, ex)
    acc = orc.residues(9, 2 * P.n, P.big_q).reshape(2, P.n)
    acc[0, :3] = [0, P.big_q - 1, (P.big_q - 1) // 2]
    acc[1, :3] = [P.big_q - 1, 0, (P.big_q + 1) // 2]
    out = np.zeros_like(acc)
    for j in (0, 5, 99):
        assert H.sim_fhew_fast_step(h, j, acc.reshape(-1), out.reshape(-1)) == 0
        assert (out == K.external_product(j, acc)).all(), j
    for v in (0, 1, 10):
        assert H.sim_fhew_fast_step(h, 0x8000 | v, acc.reshape(-1), out.reshape(-1)) == 0
        assert (out == K.automorphism(v, acc)).all(), v
    bits = np.array([0, 0, 1, 1, 0, 1, 0, 1], dtype=np.int32)
    cts = K.encrypt(bits, 7)
    lin = (cts[:4] + cts[4:]) % np.uint64(P.big_q)
    pro = K.prologue(lin)
    f = orc.fhew_gate_poly(P, [1, 1, 1, 0])
    q8 = int(round(P.big_q / 8.0))
    ref = K.op([1, 1, 1, 0], lin[:2], threads=2)
    for i in range(2):
        o = np.zeros(P.n + 1, dtype=np.uint64)
        a = np.zeros(2 * P.n, dtype=np.uint64)
        assert H.sim_fhew_fast_blind_rotate_extract(h, f, pro[i], q8, o, a) == 0
        assert (a.reshape(2, -1) == K.blind_rotate(f, pro[i])).all()
        assert (o == ref[i]).all()
        # the kernel's 64-thread barriers: either half may run ahead of the other between two full barriers
        H.sim_fhew_fast_blind_rotate_extract_drift.argtypes = [C.c_void_p, u64p, u64p, C.c_uint64, u64p, C.c_int]
        for drift in (1, 2):
            o2 = np.zeros(P.n + 1, dtype=np.uint64)
            assert H.sim_fhew_fast_blind_rotate_extract_drift(h, f, pro[i], q8, o2, drift) == 0
            assert (o2 == ref[i]).all(), drift
    H.sim_fhew_key_free(h)


def _rot(poly, e):
    """poly * X^e over T64[X]/(X^N + 1), e in [0, 2N) (ring.rs:299-313)."""
    n = len(poly)
    out = np.empty_like(poly)
    for c in range(n):
        src = (c - e) % (2 * n)
        out[c] = poly[src % n] if src < n else np.uint64((1 << 64) - int(poly[src % n])) if poly[src % n] else np.uint64(0)
    return out


@pytest.mark.parametrize("bs_d,big_n,n,bs_log_b", [(1, 2048, 6, 23), (1, 1024, 6, 23), (3, 1024, 5, 7), (2, 512, 6, 10), (2, 2048, 3, 12), (3, 512, 4, 8)])
def test_kernel_logic_tfhe_fast_mode(H, orc, bs_d, big_n, n, bs_log_b):
    """Bounded-error TFHE path (tfhe_fast.cuh: 5 fused passes, Fourier-domain accumulation, FMA butterflies) against the oracle's
    restatement of the reference dataflow: every CMUX output within (k+1)d * 2^(64 + log_b + log_n - 53) torus units of the
    reference's (c64.rs:186-208 is the reference's own per-product bound), blind rotation decrypts to the same message."""
    u64 = np.uint64
    H.sim_tfhe_fast_key.restype = C.c_void_p
    H.sim_tfhe_fast_key.argtypes = [C.c_uint] * 4 + [u64p]
    H.sim_tfhe_fast_key_free.argtypes = [C.c_void_p]
    H.sim_tfhe_fast_cmux.argtypes = [C.c_void_p, C.c_uint, C.c_uint, u64p]
    H.sim_tfhe_fast_blind_rotate_extract.argtypes = [C.c_void_p, u64p, u64p, u64p]
    P = _tfhe_small_param(orc, n=n, big_n=big_n, k=1, bs_d=bs_d, bs_log_b=bs_log_b)
    K = orc.TfheKey(P, 0x5EED0003)
    ex = K.export()
    log_n = P.big_n.bit_length() - 1
    h = H.sim_tfhe_fast_key(log_n, P.n, P.bs_log_b, P.bs_d, ex["brk"].reshape(-1))
    assert h
    bound = 2 * bs_d * 2.0 ** (64 + bs_log_b + log_n - 53)
    acc = orc.splitmix64(19, 2 * P.big_n).reshape(2, P.big_n)
    worst = 0
    for step, e in ((0, 1), (P.n - 1, P.big_n), (1, 2 * P.big_n - 1), (2, 777 % (2 * P.big_n))):
        rot = np.stack([_rot(acc[0], e), _rot(acc[1], e)])
        ref = acc + K.external_product(step, rot - acc)
        got = acc.copy()
        assert H.sim_tfhe_fast_cmux(h, step, e, got.reshape(-1)) == 0
        err = np.abs((got - ref).astype(np.int64)).max()
        worst = max(worst, int(err))
        assert err < bound, (step, e, err, bound)
    assert worst > 0  # a different rounding order, not the bit-identical path
    msgs = np.arange(3, dtype=np.uint64)
    cts = K.encrypt(msgs, 5)
    cts[0, 1] = 0  # a zero mask word: that CMUX is skipped
    v = K.lut_poly(np.arange(1 << P.log_p, dtype=np.uint64))
    lut = (v << u64(64 - (P.log_p + P.padding))).astype(np.uint64)
    # after a digit flips the raw words are a different (equally valid) encryption: compare the phases b - <a, s> instead
    sk = ex["s"].astype(np.int64).astype(np.uint64)
    phase = lambda t: (int(t[-1]) - int((t[:-1] * sk).sum(dtype=np.uint64))) % (1 << 64)
    dec = lambda ph: ((ph + (1 << (log_delta - 1))) >> log_delta) % (1 << (P.log_p + P.padding))
    log_delta = 64 - (P.log_p + P.padding)
    H.sim_tfhe_fast_blind_rotate_extract32.argtypes = [C.c_void_p, u64p, u64p, u64p]
    for ct in cts:
        ref = K.blind_rotate_extract(v, ct)
        for fn in (H.sim_tfhe_fast_blind_rotate_extract, H.sim_tfhe_fast_blind_rotate_extract32):  # 64- and 32-bit accumulator words
            out = np.zeros(P.big_n + 1, dtype=np.uint64)
            assert fn(h, lut, ct, out) == 0
            d = (phase(out) - phase(ref) + (1 << 63)) % (1 << 64) - (1 << 63)
            assert abs(d) < 2 ** (log_delta - 4), (d, log_delta)
            assert dec(phase(out)) == dec(phase(ref))
        assert (out & u64(0xFFFFFFFF) == 0).all()
    H.sim_tfhe_fast_key_free(h)
