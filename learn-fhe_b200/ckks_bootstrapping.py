"""Host-side mirror of scheme/ckks/src/bootstrapping.rs (CoeffToSlot / SlotToCoeff, the part of CKKS bootstrapping the
reference implements) over the C ABI.

What runs where:
  * ONE-TIME, on the host: the special-FFT factor matrices (scheme/ckks/src/sfft.rs:75-104), their grouping by r
    (bootstrapping.rs:22-32), the baby-step giant-step plan of every grouped matrix (util/src/misc/matrix.rs:45-52, 125-158),
    the rotated diagonals of the plan encoded as RNS plaintext polynomials (Ckks::encode, ckks.rs:186-199: special inverse FFT,
    scale by q_last, truncate to integers) and the set of rotation exponents whose keys are needed (bootstrapping.rs:56-71).
    The reference recomputes plan and encodings inside every mul_mat call with 256-bit floats; here they are computed once
    (mpmath, `backend="mp"`, 320-bit) or, for timing runs at sizes where that takes minutes, in complex128 (`backend="f64"`).
  * PER CIPHERTEXT BATCH, on the device: the chain of `fhe_ckks_mul_mat` calls (rotations = automorphism + key switch, plaintext
    products + rescale, limb-wise sums), one per grouped matrix, in the reference's order (mats.iter().rev().fold).
Rotation keys are CkksKeySwitchingKey objects: uploaded on one rank and sent to the others with `broadcast_keys`.
Hoisting (sharing extend_bases across the rotations of one ciphertext) is NOT done: extend_bases does not commute bit for
bit with the sign flips of an automorphism (its f64 quotient estimate, rns.rs:331-345, rounds differently on -x), and the
contract of this path is limb-for-limb equality with the reference's order of operations."""
import numpy as np

from . import ckks as _ckks


# ---- numbers: complex128 or mpmath complex in numpy object arrays -------------------------------------------------------------------
class _F64:
    name = "f64"

    @staticmethod
    def cis(k, n4):  # e^(2 pi i k / n4)
        return np.exp(2j * np.pi * (np.asarray(k, dtype=np.float64) / n4))

    @staticmethod
    def const(v, n):
        return np.full(n, v, dtype=np.complex128)

    @staticmethod
    def conj(a):
        return np.conj(a)

    @staticmethod
    def trunc_scaled(x, scale):  # BigInt::from(F256) truncates toward zero (f256.rs:213-239)
        return [int(v) for v in np.trunc(x * float(scale))]

    @staticmethod
    def re_im(a):
        return a.real, a.imag


class _MP:
    name = "mp"

    def __init__(self, prec=320):
        import mpmath
        self.mp = mpmath
        self.prec = prec

    def cis(self, k, n4):
        mp = self.mp
        with mp.workprec(self.prec):
            return np.array([mp.expjpi(mp.mpf(2 * int(x)) / n4) for x in np.atleast_1d(k)], dtype=object)

    def const(self, v, n):
        return np.array([self.mp.mpc(v)] * n, dtype=object)

    def conj(self, a):
        return np.array([self.mp.conj(x) for x in a], dtype=object)

    def trunc_scaled(self, x, scale):
        mp = self.mp
        with mp.workprec(self.prec):
            out = []
            for v in x:
                y = v * scale
                out.append(int(mp.floor(y)) if y >= 0 else -int(mp.floor(-y)))
            return out

    def re_im(self, a):
        return np.array([x.real for x in a], dtype=object), np.array([x.imag for x in a], dtype=object)


def _backend(name):
    return _MP() if name == "mp" else _F64()


def _w(B, n):
    """sfft.rs:40-73: twiddles in powers-of-5-mod-4n order, n/2 of them: e^(2 pi i (5^k mod 4n) / 4n)."""
    pw, out = 1, []
    for _ in range(n // 2):
        out.append(pw)
        pw = pw * 5 % (4 * n)
    return B.cis(np.array(out, dtype=np.int64), 4 * n)


def _rot(a, j):  # AVec::rot_iter(j) (avec.rs:28-31): rotate left by j mod len
    return np.roll(a, -(j % len(a)))


class BabyStepGiantStep:
    """matrix.rs:125-158: idx -> (i = (idx / k) k, j = idx % k)."""

    def __init__(self, indices, k):
        self.k, self.map = k, {}
        for idx in indices:
            self.map.setdefault((idx // k) * k, set()).add(idx % k)

    def is_(self):
        return sorted(self.map)

    def js(self):
        return sorted(set(j for v in self.map.values() for j in v))

    def ijs(self):
        return sorted(set(self.is_()) | set(self.js()))

    def items(self):
        return [(i, sorted(self.map[i])) for i in sorted(self.map)]


class DiagSparseMatrix:
    """util/src/misc/matrix.rs:19-123: {diagonal index j: vector of n entries}, entry (i, (i + j) mod n) = diag_j[i]."""

    def __init__(self, B, n, diags):
        self.B, self.n, self.diags = B, n, dict(sorted(diags.items()))

    def mul(self, rhs):  # matrix.rs:90-105
        out = {}
        for i, a in self.diags.items():
            for j, b in rhs.diags.items():
                k = (i + j) % self.n
                v = a * _rot(b, i)
                out[k] = out[k] + v if k in out else v
        return DiagSparseMatrix(self.B, self.n, out)

    def inv(self):  # matrix.rs:66-80 (key n - j is NOT reduced: j = 0 becomes n, as in the reference)
        return DiagSparseMatrix(self.B, self.n, {self.n - j: self.B.conj(_rot(d, self.n - j)) / 2 for j, d in self.diags.items()})

    def bsgs(self):  # matrix.rs:45-52: the k in 1..=max_j with the fewest non-zero rotation indices (first minimum)
        js = list(self.diags)
        best = None
        for k in range(1, max(js) + 1):
            cand = BabyStepGiantStep(js, k)
            cost = sum(1 for j in cand.ijs() if j != 0)
            if best is None or cost < best[0]:
                best = (cost, cand)
        return best[1]


def sfft_fmats(B, n):
    """sfft.rs:75-99: the log2(n) butterfly factors of the special FFT over n slots."""
    log_n = n.bit_length() - 1
    mats = []
    for log_k in range(log_n):
        m = 1 << (log_n - 1 - log_k)
        w = _w(B, 2 * m)
        one, zero = B.const(1, m), B.const(0, m)
        tile = lambda pat: np.concatenate([pat] * (n // len(pat)))
        diag_zero = tile(np.concatenate([one, -w]))
        if log_k == 0:
            mats.append(DiagSparseMatrix(B, n, {0: diag_zero, n - m: tile(np.concatenate([w, one]))}))
        else:
            mats.append(DiagSparseMatrix(B, n, {0: diag_zero, n - m: tile(np.concatenate([zero, one])), m: tile(np.concatenate([w, zero]))}))
    return mats


def sifft_fmats(B, n):  # sfft.rs:102-104
    return [m.inv() for m in reversed(sfft_fmats(B, n))]


def _group(mats, r):  # bootstrapping.rs:24-25: product of every chunk of r factors
    out = []
    for c in range(0, len(mats), r):
        acc = mats[c]
        for m in mats[c + 1:c + r]:
            acc = acc.mul(m)
        out.append(acc)
    return out


def sifft(B, z):
    """sfft.rs:21-36 (the inverse special FFT Ckks::encode applies to the slot vector)."""
    z = np.array(z, dtype=object if B.name == "mp" else np.complex128)
    n = len(z)
    log_n = n.bit_length() - 1
    for log_m in range(log_n - 1, -1, -1):
        m = 1 << log_m
        t = B.conj(_w(B, 2 * m))
        z = z.reshape(-1, 2 * m)
        a, b = z[:, :m].copy(), z[:, m:].copy()
        z = np.concatenate([a + b, (a - b) * t[None, :]], axis=1).reshape(-1)  # Butterfly::dif (fft.rs:100-106)
    idx = [int(format(i, "0%db" % log_n)[::-1], 2) if log_n else 0 for i in range(n)]
    if n > 2:
        z = z[idx]  # misc.rs:29-42 bit_reverse (identity for n <= 2)
    return z / n


def pow5(log_n, j):
    """CkksParam::pow5 (ckks.rs:47-49): 5^j mod 2N."""
    return pow(5, j, 2 << log_n)


class BootstrappingParam:
    """bootstrapping.rs:14-41 + everything that can be precomputed about the two matrix chains."""

    def __init__(self, param, r, backend="mp"):
        self.param, self.r = param, r
        self.B = B = _backend(backend)
        self.l = param.n // 2
        self.sfft_fmats = _group(sfft_fmats(B, self.l), r)
        self.sifft_fmats = _group(sifft_fmats(B, self.l), r)

    def rotation_indices(self, chains=("sfft", "sifft")):
        """bootstrapping.rs:61-65: every non-zero baby / giant index of every grouped matrix (of the chains asked for)."""
        js = []
        for mat in (self.sfft_fmats if "sfft" in chains else []) + (self.sifft_fmats if "sifft" in chains else []):
            for j in mat.bsgs().ijs():
                if j != 0 and j not in js:
                    js.append(j)
        return js

    def rotation_exponent(self, j):
        """Ckks::rtk_gen / Ckks::rotate (ckks.rs:174-184, 279-282): X -> X^(5^(j mod l) mod 2N)."""
        return pow5(self.param.log_n, j % self.l)

    def encode(self, m, level):
        """Ckks::encode (ckks.rs:186-199) at `level`: slots [l] -> RNS plaintext [level][N] (uint64)."""
        z = sifft(self.B, m)
        re, im = self.B.re_im(z)
        scale = self.param.qs[self.param.big_l - 1]
        if self.B.name == "f64":  # timing path: the scaled values fit an int64 (|z| <= 1, scale ~ 2^55)
            ints = np.trunc(np.concatenate([re, im]) * float(scale)).astype(np.int64)
            return np.stack([np.mod(ints, np.int64(q)).astype(np.uint64) for q in self.param.qs[:level]])
        ints = self.B.trunc_scaled(np.concatenate([re, im]), scale)
        return np.array([[v % q for v in ints] for q in self.param.qs[:level]], dtype=np.uint64)

    def plan(self, mat, level):
        """Everything fhe_ckks_mul_mat needs for one grouped matrix at `level`: baby / giant rotation indices, the present
        mask and the encoded diagonals mat.diag(i + j).rot_iter(-i) (bootstrapping.rs:100-102) in (i, j) row-major order."""
        bs = mat.bsgs()
        baby, giant = bs.js(), bs.is_()
        present = np.zeros((len(giant), len(baby)), dtype=np.uint8)
        pts = []
        for gi, (i, js) in enumerate(bs.items()):
            for j in js:
                present[gi, baby.index(j)] = 1
                pts.append(self.encode(_rot(mat.diags[i + j], -i), level))
        return dict(baby=baby, giant=giant, present=present, pts=np.stack(pts))


class BootstrappingKey:
    """bootstrapping.rs:43-71: the rotation keys of every index the two chains use, as device-resident key-switching keys.
    `ksk_for(j)` returns the reference-layout key [2][2L][N] for rotation index j (rank 0; other ranks pass zeros of that shape
    and receive the key through broadcast_keys)."""

    def __init__(self, bparam, ksk_for, chains=("sfft", "sifft")):
        self.bparam = bparam
        self.rtk = {j: _ckks.CkksKeySwitchingKey(bparam.param, ksk_for(j)) for j in bparam.rotation_indices(chains)}
        self._plans, self._dev_plans = {}, {}

    @classmethod
    def key_gen(cls, bparam, seed, chains=("sfft", "sifft")):
        """Bootstrapping::key_gen (bootstrapping.rs:56-71) on the device: the secret, the relinearisation key and one rotation key
        per index of the chains' BSGS plans from the counter stream of `seed` (fhe_ckks_keygen); nothing but the secret leaves the
        device.  Returns (key, sk [N] int64); the relinearisation key is `key.rlk`."""
        js = bparam.rotation_indices(chains)
        sk, rlk, autk = _ckks.key_gen(bparam.param, seed, [bparam.rotation_exponent(j) for j in js])
        self = cls.__new__(cls)
        self.bparam, self.rlk = bparam, rlk
        self.rtk = dict(zip(js, autk))
        self._plans, self._dev_plans = {}, {}
        return self, sk

    def broadcast_keys(self, dist, root=0):
        for j in sorted(self.rtk):
            self.rtk[j].broadcast(dist, root)

    @property
    def nbytes(self):
        return sum(k.nbytes for k in self.rtk.values())

    def free(self):
        for k in self.rtk.values():
            k.free()
        self.rtk = {}
        if getattr(self, "rlk", None) is not None:
            self.rlk.free()
            self.rlk = None

    def plan(self, which, idx, level):
        key = (which, idx, level)
        if key not in self._plans:
            mats = self.bparam.sfft_fmats if which == "sfft" else self.bparam.sifft_fmats
            self._plans[key] = self.bparam.plan(mats[idx], level)
        return self._plans[key]


class Bootstrapping:
    @staticmethod
    def _rots(bk, idxs):
        return [(0, None) if j == 0 else (bk.bparam.rotation_exponent(j), bk.rtk[j]) for j in idxs]

    @staticmethod
    def mul_mat_dev(bk, which, idx, ct_dev, out_dev=None):
        """Device-resident form of mul_mat: ct_dev [count][2][level][N] (torch int64 CUDA) -> [count][2][level-1][N]; the
        encoded diagonals of the plan are uploaded once and stay on the device."""
        import ctypes as C
        import torch
        from . import CkksRot, dptr, hptr, to_dev
        P = bk.bparam.param
        level, count = ct_dev.shape[2], ct_dev.shape[0]
        key = (which, idx, level)
        if key not in bk._dev_plans:
            p = bk.plan(which, idx, level)
            mk = lambda lst: (CkksRot * len(lst))(*[CkksRot(t, k.h if k is not None else None) for t, k in lst])
            bk._dev_plans[key] = dict(nb=len(p["baby"]), ng=len(p["giant"]), baby=mk(Bootstrapping._rots(bk, p["baby"])),
                                      giant=mk(Bootstrapping._rots(bk, p["giant"])), present=np.ascontiguousarray(p["present"], dtype=np.uint8),
                                      pts=to_dev(p["pts"], P.ctx.device))
        d = bk._dev_plans[key]
        if out_dev is None:
            out_dev = torch.empty((count, 2, level - 1, P.n), dtype=torch.int64, device=ct_dev.device)
        P.ctx.call("fhe_ckks_mul_mat", P.h, level, count, d["nb"], C.cast(d["baby"], C.c_void_p), d["ng"], C.cast(d["giant"], C.c_void_p),
                   hptr(d["present"]), dptr(d["pts"]), dptr(ct_dev), dptr(out_dev))
        return out_dev

    @staticmethod
    def chain_dev(bk, which, ct_dev):
        """slot_to_coeff (which = "sfft") / coeff_to_slot ("sifft") on a device-resident batch."""
        mats = bk.bparam.sfft_fmats if which == "sfft" else bk.bparam.sifft_fmats
        for idx in range(len(mats) - 1, -1, -1):
            ct_dev = Bootstrapping.mul_mat_dev(bk, which, idx, ct_dev)
        return ct_dev

    @staticmethod
    def mul_mat(bk, which, idx, ct):
        """Bootstrapping::mul_mat (bootstrapping.rs:92-108) of grouped matrix `idx` of chain `which` on a host batch
        ct [count][2][level][N] -> [count][2][level-1][N]."""
        level = ct.shape[2]
        p = bk.plan(which, idx, level)
        return _ckks.Ckks.mul_mat(bk.bparam.param, Bootstrapping._rots(bk, p["baby"]), Bootstrapping._rots(bk, p["giant"]), p["present"], p["pts"], ct)

    @staticmethod
    def _mul_mats(bk, which, ct):  # bootstrapping.rs:82-90: mats.iter().rev().fold(ct, mul_mat)
        mats = bk.bparam.sfft_fmats if which == "sfft" else bk.bparam.sifft_fmats
        for idx in range(len(mats) - 1, -1, -1):
            ct = Bootstrapping.mul_mat(bk, which, idx, ct)
        return ct

    @staticmethod
    def slot_to_coeff(bk, ct):
        """bootstrapping.rs:73-75"""
        return Bootstrapping._mul_mats(bk, "sfft", ct)

    @staticmethod
    def coeff_to_slot(bk, ct):
        """bootstrapping.rs:77-79"""
        return Bootstrapping._mul_mats(bk, "sifft", ct)
