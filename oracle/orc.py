"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liborc.so (CPU restatement of the reference's polynomial-ring path).
Importable only from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs; the product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE, "liborc.so"])


class FhewParamC(C.Structure):
    _fields_ = [("log_n", C.c_uint), ("big_q", C.c_uint64), ("p", C.c_uint64), ("rlwe_log_b", C.c_uint),
                ("rlwe_d", C.c_uint), ("rgsw_log_b", C.c_uint), ("rgsw_d", C.c_uint), ("n_s", C.c_uint),
                ("q_ks", C.c_uint64), ("ks_log_b", C.c_uint), ("ks_d", C.c_uint), ("w", C.c_uint)]

    @property
    def n(self):
        return 1 << self.log_n


class TfheParamC(C.Structure):
    _fields_ = [("log_p", C.c_uint), ("padding", C.c_uint), ("n", C.c_uint), ("tlwe_std", C.c_double),
                ("ks_log_b", C.c_uint), ("ks_d", C.c_uint), ("big_n", C.c_uint), ("k", C.c_uint),
                ("tglwe_std", C.c_double), ("bs_log_b", C.c_uint), ("bs_d", C.c_uint)]


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liborc.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_last_error.restype = C.c_char_p
        for name in ("orc_zq_generator", "orc_zq_two_adic_generator", "orc_zq_pow", "orc_zq_inv", "orc_zq_mul",
                     "orc_zq_add", "orc_zq_sub", "orc_zq_neg", "orc_f64_mod_u64"):
            getattr(L, name).restype = C.c_uint64
        L.orc_zq_generator.argtypes = [C.c_uint64]
        L.orc_zq_two_adic_generator.argtypes = [C.c_uint64, C.c_uint]
        L.orc_zq_pow.argtypes = [C.c_uint64] * 3
        L.orc_zq_inv.argtypes = [C.c_uint64] * 2
        for nm in ("orc_zq_mul", "orc_zq_add", "orc_zq_sub"):
            getattr(L, nm).argtypes = [C.c_uint64] * 3
        L.orc_zq_neg.argtypes = [C.c_uint64] * 2
        L.orc_zq_to_i64.restype = C.c_int64
        L.orc_zq_to_i64.argtypes = [C.c_uint64] * 2
        L.orc_f64_mod_u64.argtypes = [C.c_double]
        L.orc_is_prime.argtypes = [C.c_uint64]
        L.orc_two_adic_primes.argtypes = [C.c_uint, C.c_uint, C.c_size_t, u64p]
        L.orc_twiddles.restype = C.c_long
        L.orc_twiddles.argtypes = [C.c_uint64, C.c_void_p, C.c_void_p, C.c_size_t]
        for nm in ("orc_vec_add", "orc_vec_sub", "orc_vec_mul"):
            getattr(L, nm).argtypes = [C.c_uint64, u64p, u64p, u64p, C.c_size_t]
            getattr(L, nm).restype = None
        L.orc_vec_neg.argtypes = [C.c_uint64, u64p, u64p, C.c_size_t]
        L.orc_vec_neg.restype = None
        for nm in ("orc_mod_switch", "orc_mod_switch_odd"):
            getattr(L, nm).argtypes = [C.c_uint64, C.c_uint64, u64p, u64p, C.c_size_t]
            getattr(L, nm).restype = None
        for nm in ("orc_ntt_fwd", "orc_ntt_inv"):
            getattr(L, nm).argtypes = [C.c_uint64, u64p, C.c_size_t, C.c_size_t, C.c_int]
        L.orc_ntt_mul.argtypes = [C.c_uint64, u64p, u64p, C.c_size_t, C.c_size_t, C.c_int]
        L.orc_schoolbook_zq.argtypes = [C.c_uint64, u64p, u64p, u64p, C.c_size_t]
        L.orc_schoolbook_t64.argtypes = [u64p, u64p, u64p, C.c_size_t]
        L.orc_fft64_mul.argtypes = [u64p, u64p, C.c_size_t, C.c_size_t, C.c_int]
        L.orc_automorphism_zq.argtypes = [C.c_uint64, u64p, u64p, C.c_size_t, C.c_int64]
        L.orc_automorphism_t64.argtypes = [u64p, u64p, C.c_size_t, C.c_int64]
        L.orc_monomial_mul_zq.argtypes = [C.c_uint64, u64p, C.c_size_t, C.c_int64]
        L.orc_monomial_mul_t64.argtypes = [u64p, C.c_size_t, C.c_int64]
        L.orc_decompose_zq.argtypes = [C.c_uint64, C.c_uint, C.c_uint, u64p, C.c_size_t, u64p]
        L.orc_decompose_t64.argtypes = [C.c_uint, C.c_uint, u64p, C.c_size_t, u64p]
        L.orc_decomposor_zq_info.argtypes = [C.c_uint64, C.c_uint, C.c_uint, C.POINTER(C.c_uint), C.POINTER(C.c_uint), u64p]
        L.orc_rounding_shr_t64.argtypes = [u64p, u64p, C.c_size_t, C.c_uint]
        L.orc_rns_extend_bases.argtypes = [u64p, C.c_size_t, u64p, C.c_size_t, u64p, u64p, C.c_size_t]
        L.orc_rns_switch_bases.argtypes = [u64p, C.c_size_t, u64p, C.c_size_t, u64p, u64p, C.c_size_t]
        L.orc_rns_rescale_k.argtypes = [u64p, C.c_size_t, C.c_size_t, u64p, u64p, C.c_size_t]
        # FHEW
        L.orc_fhew_testing_param.argtypes = [C.POINTER(FhewParamC)]
        L.orc_fhew_testing_param.restype = None
        L.orc_fhew_keygen.argtypes = [C.POINTER(FhewParamC), C.c_uint64]
        L.orc_fhew_keygen.restype = C.c_void_p
        L.orc_fhew_key_free.argtypes = [C.c_void_p]
        L.orc_fhew_key_free.restype = None
        L.orc_fhew_key_import.argtypes = [C.POINTER(FhewParamC), u64p, u64p, u64p, u64p, i64p]
        L.orc_fhew_key_import.restype = C.c_void_p
        L.orc_fhew_keygen_ctr.restype = C.c_void_p
        L.orc_fhew_key_export.argtypes = [C.c_void_p] + [C.c_void_p] * 7
        L.orc_fhew_encrypt.argtypes = [C.c_void_p, i32p, C.c_size_t, C.c_uint64, u64p]
        L.orc_fhew_decrypt.argtypes = [C.c_void_p, u64p, C.c_size_t, i32p]
        L.orc_fhew_phase.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p]
        L.orc_fhew_gate_poly.argtypes = [C.POINTER(FhewParamC), i32p, u64p]
        L.orc_fhew_op.argtypes = [C.c_void_p, i32p, u64p, C.c_size_t, u64p, C.c_int]
        L.orc_fhew_bootstrap.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t, u64p, C.c_int]
        L.orc_fhew_prologue.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p]
        L.orc_lwe_key_switch.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p]
        L.orc_fhew_schedule.argtypes = [C.POINTER(FhewParamC), u64p, i32p, C.c_size_t]
        L.orc_fhew_schedule.restype = C.c_long
        L.orc_fhew_external_product.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p]
        L.orc_fhew_automorphism.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p]
        L.orc_fhew_blind_rotate.argtypes = [C.c_void_p, u64p, u64p, u64p]
        L.orc_rlwe_decrypt.argtypes = [C.c_void_p, u64p, u64p]
        # TFHE
        L.orc_tfhe_testing_param.argtypes = [C.POINTER(TfheParamC)]
        L.orc_tfhe_testing_param.restype = None
        L.orc_tfhe_keygen.argtypes = [C.POINTER(TfheParamC), C.c_uint64]
        L.orc_tfhe_keygen.restype = C.c_void_p
        L.orc_rgsw_internal_product.argtypes = [C.c_uint64, C.c_uint, C.c_uint, C.c_uint, u64p, u64p, u64p]
        L.orc_tfhe_key_import.restype = C.c_void_p
        L.orc_tfhe_keygen_ctr.restype = C.c_void_p
        L.orc_tfhe_keygen_ctr.argtypes = [C.c_void_p, C.c_uint64]
        L.orc_tfhe_key_import.argtypes = [C.c_void_p, u64p, u64p, u64p]
        L.orc_ckks_key_import.restype = C.c_void_p
        L.orc_ckks_keygen_ctr.restype = C.c_void_p
        L.orc_ckks_keygen_ctr.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_uint64, i64p, C.c_size_t]
        L.orc_ckks_key_import.argtypes = [C.c_uint, u64p, u64p, C.c_size_t, u64p, i64p, u64p, C.c_size_t]
        L.orc_tfhe_key_free.argtypes = [C.c_void_p]
        L.orc_tfhe_key_free.restype = None
        L.orc_tfhe_key_export.argtypes = [C.c_void_p] + [C.c_void_p] * 5
        L.orc_tfhe_encrypt.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_uint64, u64p]
        L.orc_tfhe_decrypt.argtypes = [C.c_void_p, u64p, C.c_size_t, C.c_void_p, C.c_void_p]
        L.orc_tfhe_lut_poly.argtypes = [C.POINTER(TfheParamC), u64p, u64p]
        L.orc_tfhe_bootstrap.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t, u64p, C.c_int]
        L.orc_tfhe_blind_rotate_extract.argtypes = [C.c_void_p, u64p, u64p, u64p]
        L.orc_tfhe_external_product.argtypes = [C.c_void_p, C.c_size_t, u64p, u64p]
        L.orc_tfhe_key_switch.argtypes = [C.c_void_p, u64p, u64p]
        # CKKS
        L.orc_ckks_keygen.argtypes = [C.c_uint, C.c_uint, C.c_uint, C.c_uint64, i64p, C.c_size_t]
        L.orc_ckks_keygen.restype = C.c_void_p
        L.orc_ckks_key_free.argtypes = [C.c_void_p]
        L.orc_ckks_key_free.restype = None
        L.orc_ckks_moduli.argtypes = [C.c_void_p, u64p, u64p]
        L.orc_ckks_ksk_export.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.orc_ckks_encrypt.argtypes = [C.c_void_p, i64p, C.c_size_t, C.c_uint64, u64p]
        L.orc_ckks_decrypt.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p]
        L.orc_ckks_mul.argtypes = [C.c_void_p, u64p, u64p, C.c_size_t, C.c_size_t, u64p, C.c_int]
        L.orc_ckks_key_switch.argtypes = [C.c_void_p, C.c_int, C.c_int, u64p, C.c_size_t, u64p]
        L.orc_ckks_rescale.argtypes = [C.c_void_p, u64p, C.c_size_t, u64p]
        _LIB = L
    return _LIB


def _ck(rc):
    if rc != 0:
        raise RuntimeError("oracle: " + lib().orc_last_error().decode())


def U(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


def splitmix64(seed, n):
    """Counter-based generator shared by oracle-side tests, bench and the CUDA tests."""
    idx = np.arange(1, n + 1, dtype=np.uint64)
    with np.errstate(over="ignore"):
        z = np.uint64(seed) + idx * np.uint64(0x9E3779B97F4A7C15)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def residues(seed, n, q):
    return splitmix64(seed, n) % np.uint64(q)


# ------------------------------------------------------------------ util level
def two_adic_primes(bits, log_n, count):
    out = np.zeros(count, dtype=np.uint64)
    _ck(lib().orc_two_adic_primes(bits, log_n, count, out))
    return [int(x) for x in out]


def twiddles(q):
    n = lib().orc_twiddles(q, None, None, 0)
    if n < 0:
        _ck(-1)
    f = np.zeros(n, dtype=np.uint64)
    i = np.zeros(n, dtype=np.uint64)
    lib().orc_twiddles(q, f.ctypes.data, i.ctypes.data, n)
    return f, i


def ntt_fwd(q, a, threads=1):
    a = U(a).copy()
    n = a.shape[-1]
    _ck(lib().orc_ntt_fwd(q, a.reshape(-1), n, a.size // n, threads))
    return a


def ntt_inv(q, a, threads=1):
    a = U(a).copy()
    n = a.shape[-1]
    _ck(lib().orc_ntt_inv(q, a.reshape(-1), n, a.size // n, threads))
    return a


def ntt_mul(q, a, b, threads=1):
    a = U(a).copy()
    b = U(b)
    n = a.shape[-1]
    _ck(lib().orc_ntt_mul(q, a.reshape(-1), b.reshape(-1), n, a.size // n, threads))
    return a


def schoolbook_zq(q, a, b):
    out = np.zeros_like(U(a))
    _ck(lib().orc_schoolbook_zq(q, U(a), U(b), out, len(out)))
    return out


def schoolbook_t64(a, b):
    out = np.zeros_like(U(a))
    _ck(lib().orc_schoolbook_t64(U(a), U(b), out, len(out)))
    return out


def fft64_mul(a, b, threads=1):
    a = U(a).copy()
    b = U(b)
    n = a.shape[-1]
    _ck(lib().orc_fft64_mul(a.reshape(-1), b.reshape(-1), n, a.size // n, threads))
    return a


def vec_op(name, q, a, b=None):
    a = U(a)
    out = np.zeros_like(a)
    if b is None:
        getattr(lib(), "orc_vec_" + name)(q, a.reshape(-1), out.reshape(-1), a.size)
    else:
        getattr(lib(), "orc_vec_" + name)(q, a.reshape(-1), U(b).reshape(-1), out.reshape(-1), a.size)
    return out


def mod_switch(q, qp, a, odd=False):
    a = U(a)
    out = np.zeros_like(a)
    (lib().orc_mod_switch_odd if odd else lib().orc_mod_switch)(q, qp, a.reshape(-1), out.reshape(-1), a.size)
    return out


def automorphism_zq(q, a, t):
    out = np.zeros_like(U(a))
    _ck(lib().orc_automorphism_zq(q, U(a), out, len(out), t))
    return out


def automorphism_t64(a, t):
    out = np.zeros_like(U(a))
    _ck(lib().orc_automorphism_t64(U(a), out, len(out), t))
    return out


def monomial_mul_zq(q, a, k):
    a = U(a).copy()
    _ck(lib().orc_monomial_mul_zq(q, a, len(a), k))
    return a


def monomial_mul_t64(a, k):
    a = U(a).copy()
    _ck(lib().orc_monomial_mul_t64(a, len(a), k))
    return a


def decompose_zq(q, log_b, d, a):
    a = U(a)
    out = np.zeros((d, a.size), dtype=np.uint64)
    _ck(lib().orc_decompose_zq(q, log_b, d, a.reshape(-1), a.size, out.reshape(-1)))
    return out


def decompose_t64(log_b, d, a):
    a = U(a)
    out = np.zeros((d, a.size), dtype=np.uint64)
    _ck(lib().orc_decompose_t64(log_b, d, a.reshape(-1), a.size, out.reshape(-1)))
    return out


def decomposor_zq_info(q, log_b, d):
    lq, rb = C.c_uint(), C.c_uint()
    bases = np.zeros(d, dtype=np.uint64)
    _ck(lib().orc_decomposor_zq_info(q, log_b, d, C.byref(lq), C.byref(rb), bases))
    return lq.value, rb.value, bases


def rounding_shr_t64(a, bits):
    a = U(a)
    out = np.zeros_like(a)
    _ck(lib().orc_rounding_shr_t64(a.reshape(-1), out.reshape(-1), a.size, bits))
    return out


def rns_extend_bases(qs, ps, x):
    x = U(x)
    n = x.shape[-1]
    out = np.zeros((len(qs) + len(ps), n), dtype=np.uint64)
    _ck(lib().orc_rns_extend_bases(U(qs), len(qs), U(ps), len(ps), x.reshape(-1), out.reshape(-1), n))
    return out


def rns_switch_bases(qs, ps, x):
    x = U(x)
    n = x.shape[-1]
    out = np.zeros((len(ps), n), dtype=np.uint64)
    _ck(lib().orc_rns_switch_bases(U(qs), len(qs), U(ps), len(ps), x.reshape(-1), out.reshape(-1), n))
    return out


def rns_rescale_k(qs, k, x):
    x = U(x)
    n = x.shape[-1]
    out = np.zeros((len(qs) - k, n), dtype=np.uint64)
    _ck(lib().orc_rns_rescale_k(U(qs), len(qs), k, x.reshape(-1), out.reshape(-1), n))
    return out


# ------------------------------------------------------------------ FHEW
def fhew_testing_param():
    p = FhewParamC()
    lib().orc_fhew_testing_param(C.byref(p))
    return p


class FhewKey:
    def __init__(self, param, seed, _handle=None):
        self.param = param
        self.h = _handle if _handle is not None else lib().orc_fhew_keygen(C.byref(param), seed)
        if not self.h:
            _ck(-1)

    @classmethod
    def ctr(cls, param, seed):
        """Key generation fed by the counter-based stream of the device keygen (oracle/orc_keygen.hpp): the checker of
        fhe_fhew_keygen."""
        return cls(param, seed, _handle=lib().orc_fhew_keygen_ctr(C.byref(param), seed))

    @classmethod
    def from_arrays(cls, param, ksk_a, ksk_b, brk, ak, ak_t):
        """Public evaluation keys only (no secrets: encrypt/decrypt are unavailable); reference layout as in export()."""
        h = lib().orc_fhew_key_import(C.byref(param), U(ksk_a).reshape(-1), U(ksk_b).reshape(-1), U(brk).reshape(-1),
                                      U(ak).reshape(-1), np.ascontiguousarray(ak_t, dtype=np.int64))
        return cls(param, 0, _handle=h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_fhew_key_free(self.h)
            self.h = None

    def export(self):
        P = self.param
        n = P.n
        out = dict(
            ksk_a=np.zeros((n * P.ks_d, P.n_s), dtype=np.uint64), ksk_b=np.zeros(n * P.ks_d, dtype=np.uint64),
            brk=np.zeros((P.n_s, 2 * P.rgsw_d, 2, n), dtype=np.uint64),
            ak=np.zeros((P.w + 1, P.rlwe_d, 2, n), dtype=np.uint64), ak_t=np.zeros(P.w + 1, dtype=np.int64),
            z=np.zeros(n, dtype=np.int64), s=np.zeros(P.n_s, dtype=np.int64))
        _ck(lib().orc_fhew_key_export(self.h, *[out[k].ctypes.data for k in ("ksk_a", "ksk_b", "brk", "ak", "ak_t", "z", "s")]))
        return out

    def encrypt(self, bits, seed):
        bits = np.ascontiguousarray(bits, dtype=np.int32)
        cts = np.zeros((len(bits), self.param.n + 1), dtype=np.uint64)
        _ck(lib().orc_fhew_encrypt(self.h, bits, len(bits), seed, cts.reshape(-1)))
        return cts

    def decrypt(self, cts):
        cts = U(cts)
        out = np.zeros(len(cts), dtype=np.int32)
        _ck(lib().orc_fhew_decrypt(self.h, cts.reshape(-1), len(cts), out))
        return out

    def phase(self, cts):
        cts = U(cts)
        out = np.zeros(len(cts), dtype=np.uint64)
        _ck(lib().orc_fhew_phase(self.h, cts.reshape(-1), len(cts), out))
        return out

    def op(self, table, cts, threads=1):
        cts = U(cts)
        out = np.zeros_like(cts)
        _ck(lib().orc_fhew_op(self.h, np.ascontiguousarray(table, dtype=np.int32), cts.reshape(-1), len(cts), out.reshape(-1), threads))
        return out

    def bootstrap(self, f, cts, threads=1):
        cts = U(cts)
        out = np.zeros_like(cts)
        _ck(lib().orc_fhew_bootstrap(self.h, U(f), cts.reshape(-1), len(cts), out.reshape(-1), threads))
        return out

    def prologue(self, cts):
        cts = U(cts)
        out = np.zeros((len(cts), self.param.n_s + 1), dtype=np.uint64)
        _ck(lib().orc_fhew_prologue(self.h, cts.reshape(-1), len(cts), out.reshape(-1)))
        return out

    def key_switch(self, cts):
        cts = U(cts)
        out = np.zeros((len(cts), self.param.n_s + 1), dtype=np.uint64)
        _ck(lib().orc_lwe_key_switch(self.h, cts.reshape(-1), len(cts), out.reshape(-1)))
        return out

    def external_product(self, j, acc):
        acc = U(acc)
        out = np.zeros_like(acc)
        _ck(lib().orc_fhew_external_product(self.h, j, acc.reshape(-1), out.reshape(-1)))
        return out

    def automorphism(self, v, acc):
        acc = U(acc)
        out = np.zeros_like(acc)
        _ck(lib().orc_fhew_automorphism(self.h, v, acc.reshape(-1), out.reshape(-1)))
        return out

    def blind_rotate(self, f, ct2n):
        out = np.zeros((2, self.param.n), dtype=np.uint64)
        _ck(lib().orc_fhew_blind_rotate(self.h, U(f), U(ct2n), out.reshape(-1)))
        return out

    def rlwe_decrypt(self, acc):
        out = np.zeros(self.param.n, dtype=np.uint64)
        _ck(lib().orc_rlwe_decrypt(self.h, U(acc).reshape(-1), out))
        return out


def rgsw_internal_product(q, log_n, log_b, d, ct0, ct1):
    """Rgsw::internal_product (rgsw.rs:130-150): RGSW ciphertexts [2d][2][n] over Z_q."""
    ct0, ct1 = U(ct0), U(ct1)
    out = np.zeros_like(ct0)
    _ck(lib().orc_rgsw_internal_product(q, log_n, log_b, d, ct0.reshape(-1), ct1.reshape(-1), out.reshape(-1)))
    return out


def fhew_gate_poly(param, table):
    f = np.zeros(param.n, dtype=np.uint64)
    _ck(lib().orc_fhew_gate_poly(C.byref(param), np.ascontiguousarray(table, dtype=np.int32), f))
    return f


def fhew_schedule(param, a):
    cap = 4 * (param.n_s + param.n)
    steps = np.zeros(2 * cap, dtype=np.int32)
    cnt = lib().orc_fhew_schedule(C.byref(param), U(a), steps, cap)
    if cnt < 0:
        _ck(-1)
    return steps[: 2 * cnt].reshape(-1, 2)


# ------------------------------------------------------------------ TFHE
def tfhe_testing_param():
    p = TfheParamC()
    lib().orc_tfhe_testing_param(C.byref(p))
    return p


class TfheKey:
    def __init__(self, param, seed, _handle=None):
        self.param = param
        self.h = _handle if _handle is not None else lib().orc_tfhe_keygen(C.byref(param), seed)
        if not self.h:
            _ck(-1)

    @classmethod
    def ctr(cls, param, seed):
        """Key generation fed by the counter stream of the device keygen (oracle/orc_keygen.hpp): the checker of fhe_tfhe_keygen."""
        return cls(param, seed, _handle=lib().orc_tfhe_keygen_ctr(C.byref(param), seed))

    @classmethod
    def from_arrays(cls, param, brk, ksk_a, ksk_b):
        """Public evaluation keys only (no secrets: encrypt / decrypt are unavailable); layout as in export()."""
        h = lib().orc_tfhe_key_import(C.byref(param), U(brk).reshape(-1), U(ksk_a).reshape(-1), U(ksk_b).reshape(-1))
        return cls(param, 0, _handle=h)

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_tfhe_key_free(self.h)
            self.h = None

    def export(self):
        P = self.param
        kn = P.k * P.big_n
        out = dict(brk=np.zeros((P.n, (P.k + 1) * P.bs_d, P.k + 1, P.big_n), dtype=np.uint64),
                   ksk_a=np.zeros((kn * P.ks_d, P.n), dtype=np.uint64), ksk_b=np.zeros(kn * P.ks_d, dtype=np.uint64),
                   z=np.zeros(P.n, dtype=np.int64), s=np.zeros(kn, dtype=np.int64))
        _ck(lib().orc_tfhe_key_export(self.h, *[out[k].ctypes.data for k in ("brk", "ksk_a", "ksk_b", "z", "s")]))
        return out

    def encrypt(self, msgs, seed):
        msgs = U(msgs)
        cts = np.zeros((len(msgs), self.param.n + 1), dtype=np.uint64)
        _ck(lib().orc_tfhe_encrypt(self.h, msgs, len(msgs), seed, cts.reshape(-1)))
        return cts

    def decrypt(self, cts):
        cts = U(cts)
        m = np.zeros(len(cts), dtype=np.uint64)
        ph = np.zeros(len(cts), dtype=np.uint64)
        _ck(lib().orc_tfhe_decrypt(self.h, cts.reshape(-1), len(cts), m.ctypes.data, ph.ctypes.data))
        return m, ph

    def lut_poly(self, table):
        v = np.zeros(self.param.big_n, dtype=np.uint64)
        _ck(lib().orc_tfhe_lut_poly(C.byref(self.param), U(table), v))
        return v

    def bootstrap(self, v, cts, threads=1):
        cts = U(cts)
        out = np.zeros_like(cts)
        _ck(lib().orc_tfhe_bootstrap(self.h, U(v), cts.reshape(-1), len(cts), out.reshape(-1), threads))
        return out

    def blind_rotate_extract(self, v, ct):
        out = np.zeros(self.param.k * self.param.big_n + 1, dtype=np.uint64)
        _ck(lib().orc_tfhe_blind_rotate_extract(self.h, U(v), U(ct), out))
        return out

    def external_product(self, i, glwe):
        glwe = U(glwe)
        out = np.zeros_like(glwe)
        _ck(lib().orc_tfhe_external_product(self.h, i, glwe.reshape(-1), out.reshape(-1)))
        return out

    def key_switch(self, ct):
        out = np.zeros(self.param.n + 1, dtype=np.uint64)
        _ck(lib().orc_tfhe_key_switch(self.h, U(ct), out))
        return out


# ------------------------------------------------------------------ CKKS
class CkksKey:
    @classmethod
    def from_arrays(cls, log_n, qs, ps, rlk, auto=()):
        """Public evaluation keys only: explicit moduli, rlk [2][2L][N] and automorphism keys [(t, ksk [2][2L][N]), ...]."""
        self = cls.__new__(cls)
        self.log_n, self.big_l, self.n = log_n, len(qs), 1 << log_n
        self.qs, self.ps = [int(q) for q in qs], [int(p) for p in ps]
        self.auto_ts = [int(t) for t, _ in auto]
        ts = np.ascontiguousarray(self.auto_ts, dtype=np.int64)
        ak = U(np.stack([U(k) for _, k in auto])).reshape(-1) if auto else np.zeros(1, dtype=np.uint64)
        self.h = lib().orc_ckks_key_import(log_n, U(self.qs), U(self.ps), self.big_l, U(rlk).reshape(-1), ts if len(ts) else np.zeros(1, dtype=np.int64),
                                           ak, len(auto))
        if not self.h:
            _ck(-1)
        return self

    def __init__(self, log_n, log_qi, big_l, seed, auto_ts=(), ctr=False):
        """ctr=True: the key generation fed by the counter stream of the device keygen (oracle/orc_keygen.hpp)."""
        self.log_n, self.big_l = log_n, big_l
        self.n = 1 << log_n
        ts = np.ascontiguousarray(list(auto_ts), dtype=np.int64)
        self.auto_ts = [int(t) for t in ts]
        gen = lib().orc_ckks_keygen_ctr if ctr else lib().orc_ckks_keygen
        self.h = gen(log_n, log_qi, big_l, seed, ts if len(ts) else np.zeros(1, dtype=np.int64), len(ts))
        if not self.h:
            _ck(-1)
        qs = np.zeros(big_l, dtype=np.uint64)
        ps = np.zeros(big_l, dtype=np.uint64)
        _ck(lib().orc_ckks_moduli(self.h, qs, ps))
        self.qs, self.ps = [int(x) for x in qs], [int(x) for x in ps]

    def __del__(self):
        if getattr(self, "h", None):
            lib().orc_ckks_key_free(self.h)
            self.h = None

    def ksk(self, which=-1):
        out = np.zeros((2, 2 * self.big_l, self.n), dtype=np.uint64)
        _ck(lib().orc_ckks_ksk_export(self.h, which, out.ctypes.data, None))
        return out

    def sk(self):
        sk = np.zeros(self.n, dtype=np.int64)
        _ck(lib().orc_ckks_ksk_export(self.h, -1, None, sk.ctypes.data))
        return sk

    def encrypt(self, pt, level, seed):
        ct = np.zeros((2, level, self.n), dtype=np.uint64)
        _ck(lib().orc_ckks_encrypt(self.h, np.ascontiguousarray(pt, dtype=np.int64), level, seed, ct.reshape(-1)))
        return ct

    def decrypt(self, ct):
        ct = U(ct)
        level = ct.shape[1]
        pt = np.zeros((level, self.n), dtype=np.uint64)
        _ck(lib().orc_ckks_decrypt(self.h, ct.reshape(-1), level, pt.reshape(-1)))
        return pt

    def mul(self, ct0, ct1, threads=1):
        ct0, ct1 = U(ct0), U(ct1)
        batched = ct0.ndim == 4
        c0 = ct0 if batched else ct0[None]
        c1 = ct1 if batched else ct1[None]
        level = c0.shape[2]
        out = np.zeros((c0.shape[0], 2, level - 1, self.n), dtype=np.uint64)
        _ck(lib().orc_ckks_mul(self.h, c0.reshape(-1), c1.reshape(-1), level, c0.shape[0], out.reshape(-1), threads))
        return out if batched else out[0]

    def key_switch(self, which, ct, apply_auto=True):
        ct = U(ct)
        out = np.zeros_like(ct)
        _ck(lib().orc_ckks_key_switch(self.h, which, 1 if apply_auto else 0, ct.reshape(-1), ct.shape[1], out.reshape(-1)))
        return out

    def rescale(self, ct):
        ct = U(ct)
        out = np.zeros((2, ct.shape[1] - 1, self.n), dtype=np.uint64)
        _ck(lib().orc_ckks_rescale(self.h, ct.reshape(-1), ct.shape[1], out.reshape(-1)))
        return out
