"""Host-side mirror of the CKKS / RNS call sites (scheme/ckks/src/ckks.rs:123-129,250-293; util/src/ring/rns.rs:83-132)
over the C ABI.  Ciphertexts are [count][2 (b, a)][level][N] uint64 arrays (tuple order of ckks.rs:112-121), coefficient
form, limb-major (rns.rs:21).  Nothing here computes on the CPU."""
import ctypes as C

import numpy as np

from . import dptr, hptr, to_dev, to_host


def _u64(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


# --- util/src/ring/rns.rs -----------------------------------------------------------------------------------------------
def extend_bases(ctx, qs, ps, x):
    """RnsRq::extend_bases (rns.rs:83-91): x [batch][len(qs)][n] -> [batch][len(qs)+len(ps)][n]."""
    import torch
    x = _u64(x)
    batch, nq, n = x.shape
    d_in = to_dev(x, ctx.device)
    d_out = torch.empty((batch, nq + len(ps), n), dtype=torch.int64, device=d_in.device)
    qs_a, ps_a = _u64(qs), _u64(ps)  # keep the arrays alive across the call
    ctx.call("fhe_rns_extend_bases", hptr(qs_a), nq, hptr(ps_a), len(ps), n.bit_length() - 1, batch, dptr(d_in), dptr(d_out))
    ctx.sync()
    return to_host(d_out)


def rescale_k(ctx, qs, k, x):
    """RnsRq::rescale_k (rns.rs:103-118): x [batch][len(qs)][n] -> [batch][len(qs)-k][n]."""
    import torch
    x = _u64(x)
    batch, nq, n = x.shape
    d_in = to_dev(x, ctx.device)
    d_out = torch.empty((batch, nq - k, n), dtype=torch.int64, device=d_in.device)
    qs_a = _u64(qs)
    ctx.call("fhe_rns_rescale_k", hptr(qs_a), nq, k, n.bit_length() - 1, batch, dptr(d_in), dptr(d_out))
    ctx.sync()
    return to_host(d_out)


# --- scheme/ckks/src/ckks.rs ----------------------------------------------------------------------------------------------
class CkksParam:
    """CkksParam::new (ckks.rs:19-35): qs = first L, ps = next L of two_adic_primes(log_qi, log_n + 1)."""

    def __init__(self, ctx, log_n, qs, ps):
        assert len(qs) == len(ps)
        self.ctx, self.log_n, self.n, self.big_l = ctx, log_n, 1 << log_n, len(qs)
        self.qs, self.ps = [int(q) for q in qs], [int(p) for p in ps]
        h = C.c_void_p()
        qs_a, ps_a = _u64(self.qs), _u64(self.ps)
        ctx.call("fhe_ckks_create", log_n, hptr(qs_a), hptr(ps_a), self.big_l, C.byref(h))
        self.h = h

    @classmethod
    def new(cls, ctx, log_n, log_qi, big_l):
        from . import two_adic_primes
        primes = two_adic_primes(log_qi, log_n + 1, 2 * big_l)
        return cls(ctx, log_n, primes[:big_l], primes[big_l:])

    def free(self):
        if getattr(self, "h", None):
            self.ctx.L.fhe_ckks_destroy(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class CkksKeySwitchingKey:
    """CkksKeySwitchingKey (ckks.rs:154-162), uploaded from the reference layout [2 (b, a)][2L][N] (coefficient form)."""

    def __init__(self, param, ksk):
        self.param = param
        ksk = _u64(ksk)
        assert ksk.shape == (2, 2 * param.big_l, param.n)
        h = C.c_void_p()
        param.ctx.call("fhe_ckks_ksk_upload", param.h, hptr(ksk), C.byref(h))
        self.h = h

    @classmethod
    def _adopt(cls, param, h):
        self = cls.__new__(cls)
        self.param, self.h = param, C.c_void_p(h) if not isinstance(h, C.c_void_p) else h
        return self

    def serialize(self):
        size = int(self.param.ctx.L.fhe_ckks_ksk_serialized_size(self.h))
        buf = np.zeros(size, dtype=np.uint8)
        self.param.ctx.call("fhe_ckks_ksk_serialize", self.param.h, self.h, hptr(buf), size)
        return buf

    @classmethod
    def deserialize(cls, param, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        h = C.c_void_p()
        param.ctx.call("fhe_ckks_ksk_deserialize", param.h, hptr(blob), blob.size, C.byref(h))
        return cls._adopt(param, h)

    def free(self):
        if getattr(self, "h", None):
            self.param.ctx.L.fhe_ckks_ksk_free(self.param.ctx.h, self.h)
            self.h = None

    @property
    def nbytes(self):
        return int(self.param.ctx.L.fhe_ckks_ksk_bytes(self.h))

    def broadcast(self, dist, root=0):
        """One-time NCCL broadcast of the evaluation-form key from `root` (every rank holds a key object of the same shape)."""
        from . import nccl_comm_ptr
        comm = nccl_comm_ptr(dist, "cuda:%d" % self.param.ctx.device)
        self.param.ctx.call("fhe_ckks_ksk_broadcast", self.h, comm, root)

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


def key_gen(param, seed, auto_ts=(), export=False):
    """Ckks::sk_gen + rlk_gen + one automorphism key per exponent (ckks.rs:139-184) on the device from the counter stream of `seed`:
    returns (sk [N] int64, rlk, [automorphism keys]) and, with export=True, the coefficient-form keys [1 + n][2][2L][N] too."""
    ts = np.ascontiguousarray(list(auto_ts), dtype=np.int64)
    sk = np.zeros(param.n, dtype=np.int64)
    handles = (C.c_void_p * (1 + len(ts)))()
    ex = np.zeros((1 + len(ts), 2, 2 * param.big_l, param.n), dtype=np.uint64) if export else None
    param.ctx.call("fhe_ckks_keygen", param.h, seed, len(ts), hptr(ts) if len(ts) else None, hptr(sk), C.cast(handles, C.c_void_p),
                   hptr(ex) if export else None)
    keys = [CkksKeySwitchingKey._adopt(param, handles[i]) for i in range(1 + len(ts))]
    return (sk, keys[0], keys[1:], ex) if export else (sk, keys[0], keys[1:])


class Ckks:
    @staticmethod
    def mul(param, rlk, ct0, ct1):
        """Ckks::mul (ckks.rs:255-267), host arrays [count][2][l][N] -> [count][2][l-1][N]."""
        ct0, ct1 = _u64(ct0), _u64(ct1)
        count, _, level, n = ct0.shape
        out = np.zeros((count, 2, level - 1, n), dtype=np.uint64)
        param.ctx.call("fhe_ckks_mul_relin_rescale_batch_host", param.h, rlk.h, level, count, hptr(ct0), hptr(ct1), hptr(out))
        return out

    @staticmethod
    def mul_dev(param, rlk, level, ct0, ct1, out):
        """Device-resident form (torch int64 tensors)."""
        param.ctx.call("fhe_ckks_mul_relin_rescale_batch", param.h, rlk.h, level, ct0.shape[0], dptr(ct0), dptr(ct1), dptr(out))
        return out

    @staticmethod
    def key_switch(param, ksk, ct, t=0):
        """Ckks::key_switch (ckks.rs:284-293), preceded by X -> X^t when t != 0 (rotate / conjugate, ckks.rs:274-282)."""
        import torch
        ct = _u64(ct)
        count, _, level, n = ct.shape
        d_in = to_dev(ct, param.ctx.device)
        d_out = torch.empty_like(d_in)
        param.ctx.call("fhe_ckks_key_switch", param.h, ksk.h, t, level, count, dptr(d_in), dptr(d_out))
        param.ctx.sync()
        return to_host(d_out)

    @staticmethod
    def mul_constant(param, pt, ct):
        """Ckks::mul_constant on an encoded plaintext (ckks.rs:250-253): pt [level][N] (shared) or [count][level][N]."""
        import torch
        ct, pt = _u64(ct), _u64(pt)
        count, _, level, n = ct.shape
        pt_count = 1 if pt.ndim == 2 else pt.shape[0]
        d_ct, d_pt = to_dev(ct, param.ctx.device), to_dev(pt, param.ctx.device)
        d_out = torch.empty((count, 2, level - 1, n), dtype=torch.int64, device=d_ct.device)
        param.ctx.call("fhe_ckks_mul_plain_rescale_batch", param.h, level, count, pt_count, dptr(d_pt), dptr(d_ct), dptr(d_out))
        param.ctx.sync()
        return to_host(d_out)

    @staticmethod
    def mul_mat(param, baby, giant, present, pts, ct):
        """Bootstrapping::mul_mat (scheme/ckks/src/bootstrapping.rs:92-108).  baby / giant: lists of (t, CkksKeySwitchingKey or
        None); present [n_giant][n_baby] 0/1; pts [number present][level][N] encoded diagonals in row-major (i, j) order."""
        import torch
        from . import CkksRot
        ct, pts = _u64(ct), _u64(pts)
        count, _, level, n = ct.shape
        mk = lambda lst: (CkksRot * len(lst))(*[CkksRot(t, k.h if k is not None else None) for t, k in lst])
        b_arr, g_arr = mk(baby), mk(giant)
        pres = np.ascontiguousarray(present, dtype=np.uint8)
        d_ct, d_pts = to_dev(ct, param.ctx.device), to_dev(pts, param.ctx.device)
        d_out = torch.empty((count, 2, level - 1, n), dtype=torch.int64, device=d_ct.device)
        param.ctx.call("fhe_ckks_mul_mat", param.h, level, count, len(baby), C.cast(b_arr, C.c_void_p), len(giant), C.cast(g_arr, C.c_void_p),
                       hptr(pres), dptr(d_pts), dptr(d_ct), dptr(d_out))
        param.ctx.sync()
        return to_host(d_out)

    @staticmethod
    def rescale(param, ct):
        """CkksCiphertext::rescale (ckks.rs:123-125)."""
        import torch
        ct = _u64(ct)
        count, _, level, n = ct.shape
        d_in = to_dev(ct, param.ctx.device)
        d_out = torch.empty((count, 2, level - 1, n), dtype=torch.int64, device=d_in.device)
        param.ctx.call("fhe_ckks_rescale", param.h, level, count, dptr(d_in), dptr(d_out))
        param.ctx.sync()
        return to_host(d_out)
