"""Host-side mirror of the reference `util` crate entry points on the hot path (util/src/lib.rs:8-21), forwarding to
the C ABI.  Names follow the reference; arrays are numpy uint64 (host forms) or torch CUDA int64 tensors (device
forms, suffix `_dev`).  Nothing here computes on the CPU."""
import ctypes as C

import numpy as np

from . import dptr, hptr


def _log2(n):
    assert n > 0 and n & (n - 1) == 0, "length must be a power of two"
    return n.bit_length() - 1


# --- util/src/ring/fft/zq.rs:27-36 -------------------------------------------------------------------------------
def nega_cyclic_ntt_in_place(ctx, q, a):
    """a: numpy uint64 [..., n] host array, transformed in place (forward: natural -> bit-reversed)."""
    n = a.shape[-1]
    ctx.call("fhe_ntt_fwd_host", q, hptr(a), n, a.size // n)
    return a


def nega_cyclic_intt_in_place(ctx, q, a):
    n = a.shape[-1]
    ctx.call("fhe_ntt_inv_host", q, hptr(a), n, a.size // n)
    return a


def nega_cyclic_ntt_mul_assign(ctx, q, a, b):
    """fft/zq.rs:14-25: a <- a * b (coefficient form)."""
    n = a.shape[-1]
    ctx.call("fhe_negacyclic_mul_host", q, hptr(a), hptr(b), n, a.size // n)
    return a


def ntt_fwd_dev(ctx, q, t, log_n, bits=64):
    batch = t.numel() >> log_n
    ctx.call("fhe_ntt_fwd_u64" if bits == 64 else "fhe_ntt_fwd_u32", q, log_n, batch, dptr(t))
    return t


def ntt_inv_dev(ctx, q, t, log_n, bits=64):
    batch = t.numel() >> log_n
    ctx.call("fhe_ntt_inv_u64" if bits == 64 else "fhe_ntt_inv_u32", q, log_n, batch, dptr(t))
    return t


def twiddles(ctx, q, length):
    f = np.zeros(length, dtype=np.uint64)
    i = np.zeros(length, dtype=np.uint64)
    ctx.call("fhe_twiddles_host", q, length, hptr(f), hptr(i))
    return f, i


# --- element-wise (zq.rs:156-196, avec.rs:166-291, ring.rs:266-270) --------------------------------------------------
def _ew(ctx, name, q, a, b, out):
    ctx.call(name, q, a.numel(), dptr(a), dptr(b), dptr(out))
    return out


def pointwise_mul_dev(ctx, q, a, b, out):
    return _ew(ctx, "fhe_pointwise_mul_u64", q, a, b, out)


def pointwise_mac_dev(ctx, q, a, b, acc):
    return _ew(ctx, "fhe_pointwise_mac_u64", q, a, b, acc)


def vec_add_dev(ctx, q, a, b, out):
    return _ew(ctx, "fhe_vec_add_u64", q, a, b, out)


def vec_sub_dev(ctx, q, a, b, out):
    return _ew(ctx, "fhe_vec_sub_u64", q, a, b, out)


def vec_neg_dev(ctx, q, a, out):
    ctx.call("fhe_vec_neg_u64", q, a.numel(), dptr(a), dptr(out))
    return out


def vec_scalar_mul_dev(ctx, q, a, scalar, out):
    ctx.call("fhe_vec_scalar_mul_u64", q, a.numel(), dptr(a), scalar, dptr(out))
    return out


# --- avec.rs:34-50, ring.rs:299-313 ---------------------------------------------------------------------------------
def automorphism_dev(ctx, q, a, log_n, t, out):
    """q == 0 selects T64 (wrapping negation)."""
    ctx.call("fhe_automorphism_u64", q, log_n, a.numel() >> log_n, t, dptr(a), dptr(out))
    return out


def monomial_mul_dev(ctx, q, a, log_n, k, out):
    ctx.call("fhe_monomial_mul_u64", q, log_n, a.numel() >> log_n, k, dptr(a), dptr(out))
    return out


# --- zq.rs:128-140 ---------------------------------------------------------------------------------------------------
def mod_switch_dev(ctx, q, q_prime, a, out, odd=False):
    ctx.call("fhe_mod_switch_odd_u64" if odd else "fhe_mod_switch_u64", q, q_prime, a.numel(), dptr(a), dptr(out))
    return out


# --- misc/decompose.rs -------------------------------------------------------------------------------------------------
def decompose_zq_dev(ctx, q, log_b, d, a, out):
    """out: [d, a.numel()] limb-major."""
    ctx.call("fhe_decompose_zq", q, log_b, d, a.numel(), dptr(a), dptr(out))
    return out


def decompose_t64_dev(ctx, log_b, d, a, out):
    ctx.call("fhe_decompose_t64", log_b, d, a.numel(), dptr(a), dptr(out))
    return out


def rounding_shr_t64_dev(ctx, bits, a, out):
    ctx.call("fhe_rounding_shr_t64", bits, a.numel(), dptr(a), dptr(out))
    return out


# --- ring/fft/c64.rs:11-56 ----------------------------------------------------------------------------------------------
def nega_cyclic_fft64_mul_assign_rt(ctx, a, b):
    n = a.shape[-1]
    ctx.call("fhe_fft64_negacyclic_mul_host", hptr(a), hptr(b), n, a.size // n)
    return a


def fft64_mul_dev(ctx, a, b, log_n):
    ctx.call("fhe_fft64_negacyclic_mul", log_n, a.numel() >> log_n, dptr(a), dptr(b))
    return a


# --- ring/rns.rs:83-132 ---------------------------------------------------------------------------------------------------
def _u64arr(xs):
    return np.ascontiguousarray(xs, dtype=np.uint64)


def rns_extend_bases_dev(ctx, qs, ps, log_n, x, out):
    qs, ps = _u64arr(qs), _u64arr(ps)
    batch = x.numel() // (len(qs) << log_n)
    ctx.call("fhe_rns_extend_bases", hptr(qs), len(qs), hptr(ps), len(ps), log_n, batch, dptr(x), dptr(out))
    return out


def rns_rescale_k_dev(ctx, qs, k, log_n, x, out):
    qs = _u64arr(qs)
    batch = x.numel() // (len(qs) << log_n)
    ctx.call("fhe_rns_rescale_k", hptr(qs), len(qs), k, log_n, batch, dptr(x), dptr(out))
    return out
