"""TEST INFRASTRUCTURE: multi-party key-share generation of scheme/fhew (crs_gen, pk_share_gen / merge, key_share_gen:
bootstrapping.rs:232-294, lwe.rs:173-226, rlwe.rs:220-314, rgsw.rs:75-105) restated with numpy + the oracle's ring product,
so that Bootstrapping::key_share_merge (the part the library implements on the GPU) can be tested end to end: the merged key
must bootstrap ciphertexts that decrypt correctly under the SUM of the parties' secrets."""
import numpy as np


def _dg(rng, size):  # distribution.rs:23-47 dg(3.2, 6): discrete Gaussian, support [-6 sigma, 6 sigma]; any small noise works here
    return np.clip(np.rint(rng.normal(0.0, 3.2, size)), -19, 19).astype(np.int64)


def _zq(q, v):
    return (np.asarray(v, dtype=np.int64) % np.int64(q)).astype(np.uint64) if q < (1 << 62) else None


def _auto_i64(v, t):  # AVec<i64>::automorphism (avec.rs:34-50)
    n = len(v)
    out = np.zeros_like(v)
    for i in range(n):
        it = (i * (t % (2 * n))) % (2 * n)
        if it < n:
            out[it] = v[i]
        else:
            out[it - n] = -v[i]
    return out


def _add(q, *xs):
    acc = 0
    for x in xs:
        acc = (acc + x.astype(object)) % q
    return np.array(acc, dtype=np.uint64)


def _bases(q, log_b, d):  # Base2Decomposor<Zq>::new (decompose.rs:49-64)
    log_q = (q - 1).bit_length()
    rb = max(0, log_q - log_b * d)
    return [(1 << (rb + log_b * k)) % q for k in range(d)]


class MultiKey:
    def __init__(self, orc, P, parties, seed):
        self.P, self.orc = P, orc
        rng = np.random.default_rng(seed)
        q, n, qks, n_s, w = P.big_q, P.n, P.q_ks, P.n_s, P.w
        mul = lambda a, s: orc.ntt_mul(q, a, _zq(q, s))
        uni = lambda mod, size: rng.integers(0, mod, size=size, dtype=np.uint64)
        ak_t = [(2 * n - 5) % (2 * n)] + [pow(5, v, 2 * n) for v in range(1, w + 1)]  # bootstrapping.rs:86-89
        self.ak_t = np.array([t if t < n else t - 2 * n for t in ak_t], dtype=np.int64)
        self.crs = dict(pk=uni(q, n), ksk=uni(qks, (n * P.ks_d, n_s)), ak=uni(q, (w + 1, P.rlwe_d, n)))
        self.z = [_dg(rng, n) for _ in range(parties)]
        self.s = [_dg(rng, n_s) for _ in range(parties)]
        # collective public key (rlwe.rs:220-237): b = a * sum z_i + sum e_i
        pk_b = _add(q, *[_add(q, mul(self.crs["pk"], z), _zq(q, _dg(rng, n))) for z in self.z])
        self.pk = (self.crs["pk"], pk_b)
        gb, rbases, kb = _bases(q, P.rgsw_log_b, P.rgsw_d), _bases(q, P.rlwe_log_b, P.rlwe_d), _bases(qks, P.ks_log_b, P.ks_d)
        self.shares = []
        for z, s in zip(self.z, self.s):
            # LWE key-switching key share (lwe.rs:214-226): pt = power_up(-z) limb-major, b = <a, s> + pt + e  (mod q_ks)
            ksk = np.zeros(n * P.ks_d, dtype=np.uint64)
            for k in range(P.ks_d):
                for c in range(n):
                    idx = k * n + c
                    dot = int((self.crs["ksk"][idx].astype(object) * s.astype(object)).sum())
                    ksk[idx] = (dot + (-int(z[c])) * kb[k] + int(_dg(rng, 1)[0])) % qks
            # brk share (bootstrapping.rs:277-284): Rgsw::pk_encrypt(pk, X^{s_j})  (rgsw.rs:75-105, rlwe.rs:158-170)
            brk = np.zeros((n_s, 2 * P.rgsw_d, 2, n), dtype=np.uint64)
            for j in range(n_s):
                m = np.zeros(n, dtype=np.int64)
                e = int(s[j]) % (2 * n)
                m[e % n] = 1 if e < n else -1
                for r in range(2 * P.rgsw_d):
                    u = rng.choice(np.array([-1, 0, 1], dtype=np.int64), size=n, p=[0.25, 0.5, 0.25])  # zo(0.5)
                    a = _add(q, mul(self.pk[0], u), _zq(q, _dg(rng, n)))
                    b = _add(q, mul(self.pk[1], u), _zq(q, _dg(rng, n)))
                    pt = _zq(q, (m.astype(object) * gb[r % P.rgsw_d]) % q)
                    if r < P.rgsw_d:
                        a = _add(q, a, pt)
                    else:
                        b = _add(q, b, pt)
                    brk[j, r, 0], brk[j, r, 1] = a, b
            # automorphism key shares (rlwe.rs:280-314): b_k = a_k * z + e + (-z(X^t)) B^k
            ak = np.zeros((w + 1, P.rlwe_d, n), dtype=np.uint64)
            for v in range(w + 1):
                za = _auto_i64(z, int(self.ak_t[v]))
                for k in range(P.rlwe_d):
                    pt = _zq(q, ((-za).astype(object) * rbases[k]) % q)
                    ak[v, k] = _add(q, mul(self.crs["ak"][v, k], z), _zq(q, _dg(rng, n)), pt)
            self.shares.append(dict(ksk=ksk, brk=brk, ak=ak))
        self.z_sum = np.sum(self.z, axis=0)
        self.rng = rng

    def merge_reference(self):
        """key_share_merge with host arithmetic and the oracle's Rgsw::internal_product: (ksk_a, ksk_b, brk, ak)."""
        P, orc = self.P, self.orc
        ksk_b = _add(P.q_ks, *[sh["ksk"] for sh in self.shares])
        ak_b = _add(P.big_q, *[sh["ak"] for sh in self.shares])
        brk = self.shares[0]["brk"]
        for sh in self.shares[1:]:
            brk = np.stack([orc.rgsw_internal_product(P.big_q, P.log_n, P.rgsw_log_b, P.rgsw_d, brk[j], sh["brk"][j]) for j in range(P.n_s)])
        return self.crs["ksk"], ksk_b, brk, np.stack([self.crs["ak"], ak_b], axis=2)

    def encrypt(self, bits):
        """LWE encryptions of bits under the collective secret (lwe.rs:130-140: b = <a, z> + pt + e, pt = m * round(Q / p))."""
        P = self.P
        n, q = P.n, P.big_q
        delta = int(round(q / P.p))
        cts = np.zeros((len(bits), n + 1), dtype=np.uint64)
        for i, m in enumerate(bits):
            a = self.rng.integers(0, q, size=n, dtype=np.uint64)
            dot = int((a.astype(object) * self.z_sum.astype(object)).sum())
            cts[i, :n] = a
            cts[i, n] = (dot + int(m) * delta + int(_dg(self.rng, 1)[0])) % q
        return cts

    def decrypt(self, cts):
        P = self.P
        n, q = P.n, P.big_q
        out = []
        for ct in cts:
            ph = (int(ct[n]) - int((ct[:n].astype(object) * self.z_sum.astype(object)).sum())) % q
            out.append(int(round(ph * P.p / q)) % P.p)
        return np.array(out)
