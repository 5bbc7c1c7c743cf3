"""Host-side mirror of the FHEW call sites (scheme/fhew/src/{bootstrapping,fhew,lwe,rgsw,rlwe}.rs) over the C ABI."""
import ctypes as C

import numpy as np

from . import FhewParam, dptr, hptr

# Table 1 of ePrint 2020/086 as used by fhew.rs:58-67: (table, linear pre-op name)
GATES = {
    "and": ([0, 0, 0, 1], "add"), "nand": ([1, 1, 1, 0], "add"), "or": ([0, 1, 1, 1], "add"), "nor": ([1, 0, 0, 0], "add"),
    "xor": ([0, 1, 1, 1], "sub2"), "xnor": ([1, 0, 0, 0], "sub2"), "majority": ([0, 0, 0, 1], "add3"),
}


def single_key_testing_param(big_q):
    """fhew/boolean.rs:225-239 (big_q = two_adic_primes(28, 10).next())."""
    return FhewParam(log_n=9, big_q=big_q, p=4, rlwe_log_b=7, rlwe_d=4, rgsw_log_b=7, rgsw_d=4, n_s=100, q_ks=1 << 16,
                     ks_log_b=4, ks_d=4, w=10)


def big_q_by_8(param):
    """bootstrapping.rs:62-64: Zq::from_f64(Q, Q as f64 / 8.0)"""
    return int(round(param.big_q / 8.0)) % param.big_q


def gate_poly(param, table):
    """fhew.rs:31-36: each table entry repeated q/8 = N/4 times, value -Q/8 or +Q/8."""
    q8 = big_q_by_8(param)
    vals = [(param.big_q - q8) % param.big_q, q8]
    return np.repeat(np.array([vals[t] for t in table], dtype=np.uint64), param.n // 4)


class BootstrappingKey:
    """Device-resident BootstrappingKey (bootstrapping.rs:92-99): brk / ak rows pre-transformed to evaluation form."""

    def __init__(self, ctx, param, ksk_a, ksk_b, brk, ak, ak_t):
        self.ctx, self.param = ctx, param
        u = lambda x: np.ascontiguousarray(x, dtype=np.uint64)
        ksk_a, ksk_b, brk, ak = u(ksk_a), u(ksk_b), u(brk), u(ak)
        ak_t = np.ascontiguousarray(ak_t, dtype=np.int64)
        h = C.c_void_p()
        ctx.call("fhe_fhew_key_upload", C.byref(param), hptr(ksk_a), hptr(ksk_b), hptr(brk), hptr(ak), hptr(ak_t), C.byref(h))
        self.h = h

    @classmethod
    def _adopt(cls, ctx, param, h):
        self = cls.__new__(cls)
        self.ctx, self.param, self.h = ctx, param, h
        return self

    @classmethod
    def key_gen(cls, ctx, param, seed, export=False):
        """Bootstrapping::key_gen (bootstrapping.rs:122-146) on the device from the counter-based stream of `seed`.  Returns
        (key, z, s) - the RLWE and LWE secrets as int64 arrays - and, with export=True, also the coefficient-form key in the
        reference layout (ksk_a, ksk_b, brk, ak) for parity checks."""
        z, s = np.zeros(param.n, dtype=np.int64), np.zeros(param.n_s, dtype=np.int64)
        h = C.c_void_p()
        ex = None
        if export:
            ex = dict(ksk_a=np.zeros((param.n * param.ks_d, param.n_s), dtype=np.uint64), ksk_b=np.zeros(param.n * param.ks_d, dtype=np.uint64),
                      brk=np.zeros((param.n_s, 2 * param.rgsw_d, 2, param.n), dtype=np.uint64),
                      ak=np.zeros((param.w + 1, param.rlwe_d, 2, param.n), dtype=np.uint64))
        P = lambda k: hptr(ex[k]) if ex else None
        ctx.call("fhe_fhew_keygen", C.byref(param), seed, hptr(z), hptr(s), P("ksk_a"), P("ksk_b"), P("brk"), P("ak"), C.byref(h))
        key = cls._adopt(ctx, param, h)
        return (key, z, s, ex) if export else (key, z, s)

    def serialize(self):
        """The key as bytes: header, parameters and the device images (fhe_fhew_key_serialize)."""
        size = int(self.ctx.L.fhe_fhew_key_serialized_size(self.h))
        buf = np.zeros(size, dtype=np.uint8)
        self.ctx.call("fhe_fhew_key_serialize", self.h, hptr(buf), size)
        return buf

    @classmethod
    def deserialize(cls, ctx, blob):
        blob = np.ascontiguousarray(blob, dtype=np.uint8)
        h = C.c_void_p()
        ctx.call("fhe_fhew_key_deserialize", hptr(blob), blob.size, C.byref(h))
        # the parameters travel in the blob: fhe_fhew_param follows the 56-byte header
        param = FhewParam.from_buffer_copy(blob[56:56 + C.sizeof(FhewParam)].tobytes())
        return cls._adopt(ctx, param, h)

    def free(self):
        if getattr(self, "h", None):
            self.ctx.L.fhe_fhew_key_free(self.ctx.h, self.h)
            self.h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    @property
    def nbytes(self):
        return int(self.ctx.L.fhe_fhew_key_bytes(self.h))

    def broadcast(self, dist, root=0):
        """One-time NCCL broadcast of the transformed key buffers from `root`."""
        from . import nccl_comm_ptr
        comm = nccl_comm_ptr(dist, "cuda:%d" % self.ctx.device)
        self.ctx.call("fhe_fhew_key_broadcast", self.h, comm, root)

    def time_kernels(self, f_dev, ct_dev, out_dev, post_add, reps=3):
        """Per-kernel device time of bootstrap_dev (CUDA events on the launching stream, fhe_prof_*) and the roofline
        record of the dominant kernel (blind rotation)."""
        ctx, P = self.ctx, self.param
        count = ct_dev.shape[0]
        Bootstrapping.bootstrap_dev(self, f_dev, ct_dev, out_dev, post_add)
        ctx.prof_begin()
        for _ in range(reps):
            Bootstrapping.bootstrap_dev(self, f_dev, ct_dev, out_dev, post_add)
        prof = ctx.prof_end()
        total = sum(v["ms"] for v in prof.values())
        kernels = {k: {"ms_per_launch": v["ms"] / v["launches"], "launches_per_step": v["launches"] // reps,
                       "share": v["ms"] / total} for k, v in prof.items()}
        peak = ctx.int32_peak()
        br_ms = prof["fhew_blind_rotate_kernel"]["ms"] / prof["fhew_blind_rotate_kernel"]["launches"]
        # algorithmic 32x32 multiplies per bootstrap in the fused dataflow: Shoup butterfly = 3, MAC = 1;
        # step counts are the exact census of this batch (schedule_counts)
        n, lg = P.n, P.log_n
        bf = (n // 2) * lg
        import torch
        ct2n = torch.empty((count, P.n_s + 1), dtype=torch.int64, device=ct_dev.device)
        ctx.call("fhe_fhew_prologue_batch", self.h, count, dptr(ct_dev), dptr(ct2n))
        ctx.sync()
        n_ext, n_auto = schedule_counts(P, ct2n.cpu().numpy())
        n_bf = float(n_ext.sum()) * (2 * P.rgsw_d + 2) * bf + float(n_auto.sum()) * (P.rlwe_d + 2) * bf
        n_mac = float(n_ext.sum()) * 2 * P.rgsw_d * 2 * n + float(n_auto.sum()) * P.rlwe_d * 2 * n
        mults = 3 * n_bf + n_mac
        achieved = mults / (br_ms * 1e-3) / 1e12
        # The bound is the INT32 multiply pipe.  A Shoup butterfly needs two low products (IMAD) and one high product
        # (IMAD.HI, which occupies the pipe for two issue slots); a MAC is one 32x32->64 IMAD.WIDE.  The peak is the rate the
        # pipe sustains for exactly this mix, from the three rates measured on this device (fhe_diag_int32_peak).
        t_min = (2 * n_bf / peak["imad"] + n_bf / peak["imad_hi"] + n_mac / peak["imad_wide"]) / 1e12 if peak["imad"] else None
        peak_mix = mults / t_min / 1e12 if t_min else None
        key_bytes = self.nbytes
        io_bytes = count * ((P.n_s + 1) * 4 + (n + 1) * 8)
        roofline = {"bound": "int32", "kernel": "fhew_blind_rotate_kernel", "achieved": achieved, "peak": peak_mix,
                    "unit": "Tmul/s (algorithmic 32-bit multiplies; peak = measured rate of the INT32 multiply pipe for the "
                            "algorithm's own mix: 2 IMAD + 1 IMAD.HI per butterfly, 1 IMAD.WIDE per MAC)",
                    "frac": achieved / peak_mix if peak_mix else None, "traffic": None,
                    "frac_vs_plain_imad_rate": achieved / peak["imad"] if peak["imad"] else None,
                    "hbm_algorithmic_bytes_per_launch": int(io_bytes + key_bytes),
                    "hbm_gbs": (io_bytes + key_bytes) / (br_ms * 1e-3) / 1e9,
                    "int32_peaks_tops": peak, "ms_per_launch": br_ms, "bootstraps_per_launch": count,
                    "ext_products_per_bootstrap": float(n_ext.mean()), "automorphisms_per_bootstrap": float(n_auto.mean()),
                    "algorithmic_butterflies_per_bootstrap": n_bf / count, "algorithmic_macs_per_bootstrap": n_mac / count,
                    "algorithmic_mults_per_bootstrap": mults / count}
        return {"kernels": kernels, "roofline": roofline}


def schedule_counts(P, ct2n):
    """Number of (external products, automorphisms) blind_rotate_core (bootstrapping.rs:172-209) performs for each LWE
    ciphertext mod 2N in ct2n [count, n_s+1] — host-side census used by the roofline accounting only."""
    n, m, half = P.n, 2 * P.n, P.n // 2
    a = np.asarray(ct2n, dtype=np.int64)[:, :P.n_s]
    dlog = np.full(m, -1, dtype=np.int64)
    sign = np.zeros(m, dtype=np.int64)
    pw = 1
    for l in range(half):
        dlog[pw], sign[pw] = l, 1
        dlog[(m - pw) % m], sign[(m - pw) % m] = l, 0
        pw = pw * 5 % m
    count = a.shape[0]
    present = np.zeros((2, count, half), dtype=bool)  # [side (0 = minus, 1 = plus)][ct][l]
    nz = a != 0
    rows = np.nonzero(nz)[0]
    vals = a[nz]
    present[sign[vals], rows, dlog[vals]] = True
    autos = np.ones(count, dtype=np.int64)  # the t = -g automorphism between the two sweeps
    for side in range(2):
        v = np.zeros(count, dtype=np.int64)
        for l in range(half - 1, 0, -1):
            v += 1
            trig = present[side][:, l - 1] | (v == P.w) | (l == 1)
            autos += trig
            v[trig] = 0
    return nz.sum(axis=1), autos


class Bootstrapping:
    @staticmethod
    def bootstrap(bk, f, ct, post_add=0):
        """bootstrapping.rs:149-155 on a host batch ct [count, N+1] (numpy uint64); f [N]."""
        ct = np.ascontiguousarray(ct, dtype=np.uint64)
        f = np.ascontiguousarray(f, dtype=np.uint64)
        out = np.empty_like(ct)
        bk.ctx.call("fhe_fhew_bootstrap_batch_host", bk.h, hptr(f), post_add, ct.shape[0], hptr(ct), hptr(out))
        return out

    @staticmethod
    def bootstrap_dev(bk, f_dev, ct_dev, out_dev, post_add=0):
        count = ct_dev.shape[0]
        bk.ctx.call("fhe_fhew_bootstrap_batch", bk.h, dptr(f_dev), post_add, count, dptr(ct_dev), dptr(out_dev))
        return out_dev


class Rgsw:
    @staticmethod
    def internal_product(ctx, param, ct0, ct1):
        """Rgsw::internal_product (rgsw.rs:130-150) on host batches of RGSW ciphertexts [count, 2d, 2 (a, b), N] (or one
        ciphertext [2d, 2, N]) over Z_Q with the RGSW decomposor of `param`."""
        import torch
        from . import to_dev, to_host
        ct0, ct1 = np.ascontiguousarray(ct0, dtype=np.uint64), np.ascontiguousarray(ct1, dtype=np.uint64)
        single = ct0.ndim == 3
        if single:
            ct0, ct1 = ct0[None], ct1[None]
        assert ct0.shape == ct1.shape and ct0.shape[1:] == (2 * param.rgsw_d, 2, param.n)
        d0, d1 = to_dev(ct0, ctx.device), to_dev(ct1, ctx.device)
        out = torch.empty_like(d0)
        ctx.call("fhe_rgsw_internal_product", param.big_q, param.log_n, param.rgsw_log_b, param.rgsw_d, ct0.shape[0], dptr(d0), dptr(d1), dptr(out))
        ctx.sync()
        r = to_host(out)
        return r[0] if single else r


def key_share_merge(ctx, param, crs, shares):
    """Bootstrapping::key_share_merge (bootstrapping.rs:295-320) on the device.  crs: dict(ksk [N ks_d, n_s] mod q_ks,
    ak [w+1, rlwe_d, N] mod Q); every share: dict(ksk [N ks_d] mod q_ks, brk [n_s, 2 rgsw_d, 2, N], ak [w+1, rlwe_d, N]).
    ksk / ak shares are summed (lwe.rs:228-238, rlwe.rs:294-326), the brk shares of each LWE index are folded with
    Rgsw::internal_product, all n_s indices in one batched call per party.  Returns (ksk_a, ksk_b, brk, ak) in the layout of
    BootstrappingKey / fhe_fhew_key_upload."""
    import torch
    from . import to_dev, to_host
    from . import util
    dev = lambda a: to_dev(np.ascontiguousarray(a, dtype=np.uint64), ctx.device)
    ksk_b = dev(shares[0]["ksk"])
    ak_b = dev(shares[0]["ak"])
    brk = dev(shares[0]["brk"])
    for sh in shares[1:]:
        nxt = dev(sh["ksk"])
        util.vec_add_dev(ctx, param.q_ks, ksk_b, nxt, ksk_b)
        nxt_ak = dev(sh["ak"])
        util.vec_add_dev(ctx, param.big_q, ak_b, nxt_ak, ak_b)
        nxt_brk = dev(sh["brk"])
        out = torch.empty_like(brk)
        ctx.call("fhe_rgsw_internal_product", param.big_q, param.log_n, param.rgsw_log_b, param.rgsw_d, brk.shape[0], dptr(brk), dptr(nxt_brk), dptr(out))
        ctx.sync()
        brk = out
    ctx.sync()
    ak = np.stack([np.ascontiguousarray(crs["ak"], dtype=np.uint64), to_host(ak_b)], axis=2)  # [w+1][d][2 (a, b)][N]
    return np.ascontiguousarray(crs["ksk"], dtype=np.uint64), to_host(ksk_b), to_host(brk), ak


class Fhew:
    @staticmethod
    def linear(param, name, cts):
        q = np.uint64(param.big_q)
        if name == "add":
            return (cts[0] + cts[1]) % q
        if name == "add3":
            return (cts[0] + cts[1] + cts[2]) % q
        if name == "sub2":  # (ct0 - ct1).double()
            d = (cts[0] + (q - cts[1])) % q
            return (d + d) % q
        raise ValueError(name)

    @staticmethod
    def op(bk, table, ct):
        """fhew.rs:31-39: bootstrap with the gate polynomial, then b += Q/8."""
        return Bootstrapping.bootstrap(bk, gate_poly(bk.param, table), ct, post_add=big_q_by_8(bk.param))

    @staticmethod
    def gate(bk, name, *cts):
        table, lin = GATES[name]
        return Fhew.op(bk, table, Fhew.linear(bk.param, lin, cts))

    @staticmethod
    def not_(param, ct):
        """fhew.rs:27-29: (-a, -b + Q/4) — linear, no bootstrap."""
        q = np.uint64(param.big_q)
        out = (q - ct) % q
        q4 = np.uint64(int(round(param.big_q / 4.0)) % param.big_q)
        out[..., -1] = (out[..., -1] + q4) % q
        return out
