"""Host-side mirror of the TFHE call sites (scheme/tfhe/src/{tlwe,tglwe,tggsw,bootstrapping}.rs; util/src/ring.rs:315-320)
over the C ABI.  Torus words are numpy uint64 (host forms) or torch CUDA int64 tensors (`_dev`).  Nothing here computes
on the CPU."""
import ctypes as C

import numpy as np

from . import TfheParam, dptr, hptr, to_dev, to_host


def _u64(x):
    return np.ascontiguousarray(x, dtype=np.uint64)


def bootstrapping_testing_param():
    """tfhe/bootstrapping.rs:141-152: log_p 4, padding 1, n 1024, ks (4, 5), N 2048, k 1, TGGSW (23, 1)."""
    return TfheParam(log_p=4, padding=1, n=1024, ks_log_b=4, ks_d=5, log_big_n=11, k=1, bs_log_b=23, bs_d=1)


def nega_cyclic_fft64_mul_assign_rt(ctx, a, b):
    """util/src/ring/fft/c64.rs:11-17 (`Rt *= &Rt`): a <- a * b over T64[X]/(X^n+1); a, b host arrays [..., n]."""
    n = a.shape[-1]
    b = _u64(b)
    ctx.call("fhe_fft64_negacyclic_mul_host", hptr(a), hptr(b), n, a.size // n)
    return a


def encode_lut(param, v):
    """Tglwe::encode (tglwe.rs:80-84 -> tlwe.rs:113-116): m << log_delta."""
    log_delta = 64 - (param.log_p + param.padding)
    return (_u64(v) << np.uint64(log_delta)).astype(np.uint64)


class BootstrappingKey:
    """Device-resident BootstrappingKey (tfhe/bootstrapping.rs:40-46): bsk polynomials in the twisted Fourier domain."""

    def __init__(self, ctx, param, brk, ksk_a, ksk_b):
        self.ctx, self.param = ctx, param
        brk, ksk_a, ksk_b = _u64(brk), _u64(ksk_a), _u64(ksk_b)
        h = C.c_void_p()
        ctx.call("fhe_tfhe_key_upload", C.byref(param), hptr(brk), hptr(ksk_a), hptr(ksk_b), C.byref(h))
        self.h = h

    @classmethod
    def key_gen(cls, ctx, param, tlwe_std, tglwe_std, seed, export=False):
        """Bootstrapping::key_gen (tfhe/bootstrapping.rs:59-76) on the device from the counter stream of `seed`: returns
        (key, z [n], s [kN]) and, with export=True, the key in the upload layout (brk, ksk_a, ksk_b) too."""
        kn = param.k * param.big_n
        z, s = np.zeros(param.n, dtype=np.int64), np.zeros(kn, dtype=np.int64)
        ex = None
        if export:
            ex = dict(brk=np.zeros((param.n, (param.k + 1) * param.bs_d, param.k + 1, param.big_n), dtype=np.uint64),
                      ksk_a=np.zeros((kn * param.ks_d, param.n), dtype=np.uint64), ksk_b=np.zeros(kn * param.ks_d, dtype=np.uint64))
        P = lambda k: hptr(ex[k]) if ex else None
        h = C.c_void_p()
        ctx.call("fhe_tfhe_keygen", C.byref(param), tlwe_std, tglwe_std, seed, hptr(z), hptr(s), P("brk"), P("ksk_a"), P("ksk_b"), C.byref(h))
        self = cls.__new__(cls)
        self.ctx, self.param, self.h = ctx, param, h
        return (self, z, s, ex) if export else (self, z, s)

    def serialize(self):
        """The key as bytes: header, parameters and the device images (fhe_tfhe_key_serialize)."""
        size = int(self.ctx.L.fhe_tfhe_key_serialized_size(self.h))
        buf = np.zeros(size, dtype=np.uint8)
        self.ctx.call("fhe_tfhe_key_serialize", self.h, hptr(buf), size)
        return buf

    @classmethod
    def deserialize(cls, ctx, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        h = C.c_void_p()
        ctx.call("fhe_tfhe_key_deserialize", hptr(blob), blob.size, C.byref(h))
        self = cls.__new__(cls)
        # the parameters travel in the blob: fhe_tfhe_param follows the 56-byte header
        self.ctx, self.h = ctx, h
        self.param = TfheParam.from_buffer_copy(blob[56:56 + C.sizeof(TfheParam)].tobytes())
        return self

    def free(self):
        if getattr(self, "h", None):
            self.ctx.L.fhe_tfhe_key_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def set_mode(self, mode):
        """0 / False (default): the reference's dataflow, bit-identical torus words; 1 / True: Fourier-domain accumulation (one
        rounding per output coefficient; within the reference's error bound, decryptions identical); 2: the fused
        bounded-error blind rotation (same contract as 1; k = 1, N in {512, 1024, 2048}); 3: mode 2 with the accumulator kept as
        the top 32 bits of every torus word (each increment is a sum of f64 products of magnitude ~2^90 whose bits below 2^35 are
        rounding noise in every mode, the reference's included; output words have zero low halves)."""
        self.ctx.call("fhe_tfhe_key_set_mode", self.h, int(mode))

    @property
    def nbytes(self):
        return int(self.ctx.L.fhe_tfhe_key_bytes(self.h))

    def broadcast(self, dist, root=0):
        from . import nccl_comm_ptr
        comm = nccl_comm_ptr(dist, "cuda:%d" % self.ctx.device)
        self.ctx.call("fhe_tfhe_key_broadcast", self.h, comm, root)


class Bootstrapping:
    @staticmethod
    def bootstrap(bk, lut_encoded, ct):
        """Bootstrapping::bootstrap (tfhe/bootstrapping.rs:78-82) on a host batch ct [count, n+1]; lut_encoded [N]."""
        ct, lut = _u64(ct), _u64(lut_encoded)
        out = np.empty_like(ct)
        bk.ctx.call("fhe_tfhe_pbs_batch_host", bk.h, hptr(lut), ct.shape[0], hptr(ct), hptr(out))
        return out

    @staticmethod
    def bootstrap_dev(bk, lut_dev, ct_dev, out_dev):
        bk.ctx.call("fhe_tfhe_pbs_batch", bk.h, dptr(lut_dev), ct_dev.shape[0], dptr(ct_dev), dptr(out_dev))
        return out_dev

    @staticmethod
    def blind_rotate_extract(bk, lut_encoded, ct):
        """blind_rotate + sample_extract(0) (tfhe/bootstrapping.rs:84-96, tglwe.rs:115-127): -> [count, kN+1]."""
        import torch
        ct = _u64(ct)
        d_in, d_lut = to_dev(ct, bk.ctx.device), to_dev(_u64(lut_encoded), bk.ctx.device)
        d_out = torch.empty((ct.shape[0], bk.param.k * bk.param.big_n + 1), dtype=torch.int64, device=d_in.device)
        bk.ctx.call("fhe_tfhe_blind_rotate_extract_batch", bk.h, dptr(d_lut), ct.shape[0], dptr(d_in), dptr(d_out))
        bk.ctx.sync()
        return to_host(d_out)


class Tggsw:
    @staticmethod
    def external_product(bk, idx, glwe):
        """Tggsw::external_product(brk[idx[i]], glwe_i) (tggsw.rs:100-112): glwe [count, k+1, N]."""
        import torch
        glwe = _u64(glwe)
        d_in = to_dev(glwe, bk.ctx.device)
        d_idx = to_dev(np.ascontiguousarray(idx, dtype=np.uint32), bk.ctx.device)
        d_out = torch.empty_like(d_in)
        bk.ctx.call("fhe_tfhe_external_product", bk.h, glwe.shape[0], dptr(d_idx), dptr(d_in), dptr(d_out))
        bk.ctx.sync()
        return to_host(d_out)


    @staticmethod
    def cmux(bk, idx, ct0, ct1):
        """Tggsw::cmux(brk[idx[i]], ct0_i, ct1_i) (tggsw.rs:114-121): glwe [count, k+1, N]."""
        import torch
        ct0, ct1 = _u64(ct0), _u64(ct1)
        d0, d1 = to_dev(ct0, bk.ctx.device), to_dev(ct1, bk.ctx.device)
        d_idx = to_dev(np.ascontiguousarray(idx, dtype=np.uint32), bk.ctx.device)
        d_out = torch.empty_like(d0)
        bk.ctx.call("fhe_tfhe_cmux", bk.h, ct0.shape[0], dptr(d_idx), dptr(d0), dptr(d1), dptr(d_out))
        bk.ctx.sync()
        return to_host(d_out)


class Tlwe:
    @staticmethod
    def key_switch(bk, ct):
        """Tlwe::key_switch (tlwe.rs:144-153): [count, kN+1] -> [count, n+1]."""
        import torch
        ct = _u64(ct)
        d_in = to_dev(ct, bk.ctx.device)
        d_out = torch.empty((ct.shape[0], bk.param.n + 1), dtype=torch.int64, device=d_in.device)
        bk.ctx.call("fhe_tlwe_key_switch_batch", bk.h, ct.shape[0], dptr(d_in), dptr(d_out))
        bk.ctx.sync()
        return to_host(d_out)
