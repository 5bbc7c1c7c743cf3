"""TEST INFRASTRUCTURE ONLY — second, independent restatement of the reference's algorithms in pure Python ints.

Written directly from the Rust sources (file:line cited per function) without looking at oracle/*.hpp, so that the
C++ oracle (oracle/liborc.so) can be cross-checked against an independent reading of the reference
(tests/test_cpu_oracle_vs_pyref.py) and so that tests/golden/*.json can be regenerated
(tests/golden/make_golden.py).  Python loops only: small sizes.

Products of polynomials use the schoolbook negacyclic rule (util/src/ring.rs:421-440) — independent of any NTT.
"""
import math


# ---- util/src/zq.rs -------------------------------------------------------------------------------------------------
def zq_to_i64(q, v):  # zq.rs:71-77
    return v if v < (q >> 1) else v - q


def zq_center_u64(q, v):  # zq.rs:83-89 (two's complement in a u64)
    return v if v < (q >> 1) else ((~(q - v)) + 1) & 0xFFFFFFFFFFFFFFFF


def rust_round(x):  # f64::round: half away from zero
    if abs(x) >= 2.0 ** 52:  # already an integer (x + 0.5 would round to even there)
        return x
    f = math.floor(abs(x))
    r = f + 1 if abs(x) - f >= 0.5 else f  # abs(x) - f is exact
    return r if x >= 0 else -r


def zq_from_f64(q, x):  # zq.rs:59-61
    return int(rust_round(x)) % q


def zq_generator(q):  # zq.rs:99-105
    order = q - 1
    for g in range(1, order):
        if pow(g, order >> 1, q) == order:
            return g
    raise ValueError("no generator")


def zq_two_adic_generator(q, log_n):  # zq.rs:107-109
    return pow(zq_generator(q), (q - 1) >> log_n, q)


def zq_mod_switch(q, v, qp):  # zq.rs:128-130   (v as f64 * q' as f64) / q as f64
    return zq_from_f64(qp, (float(v) * float(qp)) / float(q))


def zq_mod_switch_odd(q, v, qp):  # zq.rs:132-140
    x = (float(v) * float(qp)) / float(q)
    u = math.floor(x)
    if u == 0:
        return int(rust_round(x)) % qp
    return (int(u) | 1) % qp


def is_prime(n):
    if n < 2:
        return False
    for p in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        if n % p == 0:
            return n == p
    d, r = n - 1, 0
    while d % 2 == 0:
        d //= 2
        r += 1
    for a in (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37):
        x = pow(a, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(r - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def two_adic_primes(bits, log_n, count):  # zq.rs:325-329
    lo, hi = 1 << (bits - log_n - 1), 1 << (bits - log_n)
    out = []
    k = hi - 1
    while k >= lo and len(out) < count:
        c = (k << log_n) + 1
        if is_prime(c):
            out.append(c)
        k -= 1
    return out


# ---- util/src/misc.rs:29-42, util/src/ring/fft/zq.rs:58-67 -----------------------------------------------------------
def bit_reverse(vals):
    n = len(vals)
    vals = list(vals)
    if n > 2:
        lg = n.bit_length() - 1
        for i in range(n):
            j = int(format(i, "0%db" % lg)[::-1], 2)
            if i < j:
                vals[i], vals[j] = vals[j], vals[i]
    return vals


def compute_twiddle(q):
    s = ((q - 1) & -(q - 1)).bit_length() - 1  # trailing zeros of q-1
    w = zq_two_adic_generator(q, s)
    tw = [pow(w, i, q) for i in range(1 << (s - 1))]
    tw_inv = [pow(v, q - 2, q) for v in tw]
    return bit_reverse(tw), bit_reverse(tw_inv)


# ---- util/src/ring/fft.rs:40-77 ----------------------------------------------------------------------------------------
def ntt_fwd(q, a, tw=None):
    a = list(a)
    tw = tw or compute_twiddle(q)[0]
    log_n = len(a).bit_length() - 1
    for layer in range(log_n):
        m, size = 1 << layer, 1 << (log_n - layer - 1)
        for i in range(m):
            t = tw[m + i]
            base = i * 2 * size
            for k in range(size):
                u, v = a[base + k], a[base + size + k]
                tb = t * v % q
                a[base + k], a[base + size + k] = (u + tb) % q, (u - tb) % q
    return a


def ntt_inv(q, a, tw_inv=None):
    a = list(a)
    tw_inv = tw_inv or compute_twiddle(q)[1]
    n = len(a)
    log_n = n.bit_length() - 1
    for layer in reversed(range(log_n)):
        m, size = 1 << layer, 1 << (log_n - layer - 1)
        for i in range(m):
            t = tw_inv[m + i]
            base = i * 2 * size
            for k in range(size):
                u, v = a[base + k], a[base + size + k]
                a[base + k], a[base + size + k] = (u + v) % q, (u - v) * t % q
    n_inv = pow(n % q, q - 2, q)
    return [x * n_inv % q for x in a]


def schoolbook_negacyclic(a, b, mod):  # ring.rs:421-440; mod = q, or 2^64 for the torus
    n = len(a)
    out = [0] * n
    for i in range(n):
        for j in range(n):
            k = i + j
            if k < n:
                out[k] = (out[k] + a[i] * b[j]) % mod
            else:
                out[k - n] = (out[k - n] - a[i] * b[j]) % mod
    return out


# ---- util/src/avec.rs:34-50, util/src/ring.rs:299-313 ------------------------------------------------------------------
def automorphism(a, t, mod):
    n = len(a)
    t %= 2 * n
    out = list(a)
    for i in range(n):
        it = (i * t) % (2 * n)
        if it < n:
            out[it] = a[i]
        else:
            out[it - n] = (-a[i]) % mod
    return out


def monomial_mul(a, k, mod):
    n = len(a)
    i = k % (2 * n)
    r = i % n
    out = a[n - r:] + a[:n - r] if r else list(a)  # rotate_right(i % n)
    if i < n:
        return [(-v) % mod for v in out[:i]] + out[i:]
    return out[:i - n] + [(-v) % mod for v in out[i - n:]]


# ---- util/src/misc/decompose.rs ----------------------------------------------------------------------------------------
def decomposor_zq(q, log_b, d):  # decompose.rs:49-64
    log_q = (q - 1).bit_length() if q > 1 else 0  # next_power_of_two().ilog2()
    return max(0, log_q - log_b * d)  # rounding_bits


def decompose_zq(q, log_b, d, v):  # decompose.rs:42-46, 91-112
    rb = decomposor_zq(q, log_b, d)
    rounded = (v + ((1 << rb) >> 1) % q) % q
    x = zq_center_u64(q, (rounded >> rb) % q)
    b_by_2, mask, neg_b = 1 << (log_b - 1), (1 << log_b) - 1, q - (1 << log_b)
    out = []
    for _ in range(d):
        limb = x & mask
        carry = 1 if limb + (x & 1) > b_by_2 else 0
        x = ((x >> log_b) + carry) & 0xFFFFFFFFFFFFFFFF
        out.append((limb + carry * neg_b) % q)
    return out


def decompose_t64(log_b, d, v):  # decompose.rs:66-81, 114-135
    rb = max(0, 64 - log_b * d)
    M = 0xFFFFFFFFFFFFFFFF
    v = ((v + ((1 << rb) >> 1)) & M) >> rb
    mask = (1 << log_b) - 1
    out = []
    for _ in range(d):
        limb = v & mask
        v >>= log_b
        carry = ((((limb - 1) & M) | v) & limb) >> (log_b - 1)
        v = (v + carry) & M
        out.append((limb - (carry << log_b)) & M)
    return out


def decompose_poly_zq(q, log_b, d, poly):  # limb-major collection impl, decompose.rs:137-155
    per = [decompose_zq(q, log_b, d, v) for v in poly]
    return [[per[i][k] for i in range(len(poly))] for k in range(d)]


# ---- scheme/fhew/src ------------------------------------------------------------------------------------------------------
def lwe_key_switch(q_ks, log_b, d, ksk_a, ksk_b, a, b):  # lwe.rs:151-160 (digits flattened limb-major)
    digs = decompose_poly_zq(q_ks, log_b, d, a)
    flat = [x for limb in digs for x in limb]
    n_s = len(ksk_a[0])
    out_a = [sum(ksk_a[k][j] * flat[k] for k in range(len(flat))) % q_ks for j in range(n_s)]
    out_b = (sum(ksk_b[k] * flat[k] for k in range(len(flat))) + b) % q_ks
    return out_a, out_b


def rgsw_external_product(q, log_b, d, rows, acc_a, acc_b):  # rgsw.rs:116-128; rows = [(a_poly, b_poly)] * 2d
    limbs = decompose_poly_zq(q, log_b, d, acc_a) + decompose_poly_zq(q, log_b, d, acc_b)
    n = len(acc_a)
    out_a, out_b = [0] * n, [0] * n
    for (ra, rb), limb in zip(rows, limbs):
        pa, pb = schoolbook_negacyclic(ra, limb, q), schoolbook_negacyclic(rb, limb, q)
        out_a = [(x + y) % q for x, y in zip(out_a, pa)]
        out_b = [(x + y) % q for x, y in zip(out_b, pb)]
    return out_a, out_b


def rlwe_automorphism(q, log_b, d, rows, t, acc_a, acc_b):  # rlwe.rs:177-191
    a, b = automorphism(acc_a, t, q), automorphism(acc_b, t, q)
    limbs = decompose_poly_zq(q, log_b, d, a)
    n = len(a)
    out_a, out_b = [0] * n, list(b)
    for (ra, rb), limb in zip(rows, limbs):
        pa, pb = schoolbook_negacyclic(ra, limb, q), schoolbook_negacyclic(rb, limb, q)
        out_a = [(x + y) % q for x, y in zip(out_a, pa)]
        out_b = [(x + y) % q for x, y in zip(out_b, pb)]
    return out_a, out_b


def i_minus_i_plus(n, a):  # bootstrapping.rs:212-231
    m = 2 * n
    minus, plus = {}, {}
    g = 1
    for l in range(n // 2):
        minus[(-g) % m] = l
        plus[g] = l
        g = g * 5 % m
    i_minus, i_plus = [[] for _ in range(n // 2)], [[] for _ in range(n // 2)]
    for i, ai in enumerate(a):
        if ai in minus and ai not in plus:
            i_minus[minus[ai]].append(i)
        elif ai in plus and ai not in minus:
            i_plus[plus[ai]].append(i)
        elif ai == 0:
            pass
        else:
            raise ValueError("unreachable!() in the reference")
    return i_minus, i_plus


def blind_rotate_schedule(n, w, a):
    """Sequence of ('ext', j) / ('auto', v) steps of blind_rotate_core (bootstrapping.rs:172-209)."""
    i_minus, i_plus = i_minus_i_plus(n, a)
    steps, v = [], 0
    for side, sets in ((0, i_minus), (1, i_plus)):
        for l in range(len(sets) - 1, 0, -1):
            steps += [("ext", j) for j in sets[l]]
            v += 1
            if sets[l - 1] or v == w or l == 1:
                steps.append(("auto", v))
                v = 0
        steps += [("ext", j) for j in sets[0]]
        if side == 0:
            steps.append(("auto", 0))
    return steps


def sample_extract0(q, acc_a, acc_b):  # rlwe.rs:193-202 with i = 0
    a = [acc_a[0]] + [(-v) % q for v in reversed(acc_a[1:])]
    return a, acc_b[0]


def fhew_bootstrap(P, keys, f, ct):
    """Bootstrapping::bootstrap (bootstrapping.rs:149-155).  P: dict of parameters; keys: dict with ksk_a [N*d][n_s],
    ksk_b, brk [n_s][2d][2][N], ak [w+1][d][2][N], ak_t; ct = (a list, b) mod Q.  Returns (a list, b) mod Q."""
    n, q = P["n"], P["big_q"]
    a, b = ct
    a = [zq_mod_switch(q, v, P["q_ks"]) for v in a]
    b = zq_mod_switch(q, b, P["q_ks"])
    a, b = lwe_key_switch(P["q_ks"], P["ks_log_b"], P["ks_d"], keys["ksk_a"], keys["ksk_b"], a, b)
    a = [zq_mod_switch_odd(P["q_ks"], v, 2 * n) for v in a]
    b = zq_mod_switch_odd(P["q_ks"], b, 2 * n)
    g = 5
    fp = monomial_mul(automorphism(list(f), -g, q), zq_to_i64(2 * n, b * g % (2 * n)), q)
    acc_a, acc_b = [0] * n, fp
    for kind, idx in blind_rotate_schedule(n, P["w"], a):
        if kind == "ext":
            rows = [(list(r[0]), list(r[1])) for r in keys["brk"][idx]]
            acc_a, acc_b = rgsw_external_product(q, P["rgsw_log_b"], P["rgsw_d"], rows, acc_a, acc_b)
        else:
            rows = [(list(r[0]), list(r[1])) for r in keys["ak"][idx]]
            acc_a, acc_b = rlwe_automorphism(q, P["rlwe_log_b"], P["rlwe_d"], rows, int(keys["ak_t"][idx]), acc_a, acc_b)
    return sample_extract0(q, acc_a, acc_b)


def fhew_gate_poly(P, table):  # fhew.rs:31-36
    q8 = zq_from_f64(P["big_q"], P["big_q"] / 8.0)
    vals = [(-q8) % P["big_q"], q8]
    return [vals[t] for t in table for _ in range(P["n"] // 4)], q8


# ---- util/src/ring/rns.rs ----------------------------------------------------------------------------------------------------
def rns_extend_bases(qs, ps, x):
    """Rns::extend_bases (rns.rs:331-345) on one coefficient: x = residues mod qs -> residues mod qs + ps."""
    big_q = math.prod(qs)
    q_hats = [big_q // qi for qi in qs]
    q_hat_invs = [pow(qh % qi, qi - 2, qi) for qh, qi in zip(q_hats, qs)]
    vs = [xi * inv % qi for xi, inv, qi in zip(x, q_hat_invs, qs)]
    est = 0.0
    for vi, qi in zip(vs, qs):  # sequential f64 sum, ascending i, no FMA
        est = est + (1.0 / float(qi)) * float(vi)
    u = int(rust_round(est))
    out = list(x)
    for p in ps:
        acc = sum((qh % p) * vi for qh, vi in zip(q_hats, vs)) % p
        out.append((acc - u * (big_q % p)) % p)
    return out


def rns_rescale_k(qs, k, x):
    """RnsRq::rescale_k (rns.rs:99-132) on one coefficient: x = residues mod qs (kept ++ dropped); returns residues mod kept.
    round(): every limb += (P >> 1) mod q_i; k == 1: subtract the dropped limb's non-centred value, else subtract the
    base extension (switch_bases) of the dropped limbs; div(): multiply by P^-1 mod q_i."""
    kept, dropped = qs[:len(qs) - k], qs[len(qs) - k:]
    p = math.prod(dropped)
    r = [(xi + (p >> 1) % qi) % qi for xi, qi in zip(x, qs)]
    rk, rd = r[:len(kept)], r[len(kept):]
    if k == 1:
        sub = [rd[0] % qi for qi in kept]  # Zq -= u64: the value is reduced mod q_i
    else:
        sub = rns_extend_bases(dropped, kept, rd)[k:]
    return [((a - b) % qi) * pow(p % qi, qi - 2, qi) % qi for a, b, qi in zip(rk, sub, kept)]


# ---- util/src/ring/fft/c64.rs + ring/fft.rs:7-35 (f64 negacyclic product of torus polynomials) ----------------------------------
# Every f64 operation below is a separate Python float operation (IEEE double, round to nearest, never fused), in the order of
# num_complex 0.4.6 `Mul` (re = a.re*b.re - a.im*b.im, im = a.re*b.im + a.im*b.re) and of the reference's loops.
def f64_mod_u64(v):  # c64.rs:69-85
    import struct
    bits = struct.unpack("<Q", struct.pack("<d", v))[0]
    sign, exponent = bits >> 63, (bits >> 52) & 0x7FF
    mantissa = ((bits << 11) | 0x8000000000000000) & 0xFFFFFFFFFFFFFFFF
    shift = 1086 - exponent
    if -63 <= shift <= 0:
        value = (mantissa << -shift) & 0xFFFFFFFFFFFFFFFF
    elif 1 <= shift <= 64:
        value = (((mantissa >> (shift - 1)) + 1) & 0xFFFFFFFFFFFFFFFF) >> 1
    else:
        value = 0
    return value if sign == 0 else (-value) & 0xFFFFFFFFFFFFFFFF


def _cmul(a, b):
    return (a[0] * b[0] - a[1] * b[1], a[0] * b[1] + a[1] * b[0])


def c64_twiddle(n):  # c64.rs:98-108: cis((i as f64 * PI) / n as f64), i < n
    return [(math.cos((float(i) * math.pi) / float(n)), math.sin((float(i) * math.pi) / float(n))) for i in range(n)]


def fft_in_place(a, tw_bo):  # fft.rs:7-18 (Butterfly::dit: tb = t * b; a + tb, a - tb)
    n = len(a)
    for layer in reversed(range(n.bit_length() - 1)):
        size = 1 << layer
        for c in range(n // (2 * size)):
            t = tw_bo[c]
            for j in range(size):
                i0, i1 = c * 2 * size + j, c * 2 * size + size + j
                tb = _cmul(t, a[i1])
                a[i0], a[i1] = (a[i0][0] + tb[0], a[i0][1] + tb[1]), (a[i0][0] - tb[0], a[i0][1] - tb[1])


def ifft_in_place(a, tw_inv_bo, n_inv):  # fft.rs:22-35 (Butterfly::dif: a + b, (a - b) * t; then *= n_inv)
    n = len(a)
    for layer in range(n.bit_length() - 1):
        size = 1 << layer
        for c in range(n // (2 * size)):
            t = tw_inv_bo[c]
            for j in range(size):
                i0, i1 = c * 2 * size + j, c * 2 * size + size + j
                s = (a[i0][0] + a[i1][0], a[i0][1] + a[i1][1])
                d = _cmul((a[i0][0] - a[i1][0], a[i0][1] - a[i1][1]), t)
                a[i0], a[i1] = s, d
    for i in range(n):
        a[i] = (a[i][0] * n_inv, a[i][1] * n_inv)


def t64_to_i64(v):  # torus.rs:20-25
    return v - (1 << 64) if v >> 63 else v


def fft64_negacyclic_mul(a, b):
    """nega_cyclic_fft64_mul_assign_rt (c64.rs:11-56) on torus words (u64): returns a * b over T64[X]/(X^n + 1)."""
    n = len(a)
    if n == 1:
        return [(a[0] * b[0]) & 0xFFFFFFFFFFFFFFFF]
    m = n // 2
    tw_n = c64_twiddle(n)   # twist: entries i < n/2 of the table for n
    tw_m = c64_twiddle(m)   # transform of the n/2 complex points: table for n/2, bit-reversed
    tw_bo = bit_reverse(tw_m)
    tw_inv_bo = bit_reverse([(c, -s) for c, s in tw_m])

    def twisted(x):  # formula 8 of ePrint 2021/480
        return [_cmul((float(t64_to_i64(x[i])), float(t64_to_i64(x[i + m]))), tw_n[i]) for i in range(m)]

    ca, cb = twisted(a), twisted(b)
    fft_in_place(ca, tw_bo)
    fft_in_place(cb, tw_bo)
    ca = [_cmul(x, y) for x, y in zip(ca, cb)]
    ifft_in_place(ca, tw_inv_bo, 1.0 / float(m))
    lo, hi = [], []
    for i in range(m):  # formula 10
        c = _cmul(ca[i], (tw_n[i][0], -tw_n[i][1]))
        lo.append(f64_mod_u64(c[0]))
        hi.append(f64_mod_u64(c[1]))
    return lo + hi


# ---- scheme/tfhe/src ------------------------------------------------------------------------------------------------------------
def tggsw_external_product(log_b, d, rows, glwe):
    """Tggsw::external_product (tggsw.rs:100-112): rows[(k+1) d][k+1][N] torus words (the TGLWE rows of one TGGSW ciphertext:
    d rows per mask polynomial a_j, then d for the body), glwe[k+1][N] = (a_0 .. a_{k-1}, b).  Every row x limb product is
    an f64 FFT product rounded to a torus word; the sum is wrapping."""
    n = len(glwe[0])
    limbs = []
    for v in glwe:  # chain![a.., b].flat_map(decompose): limb-major per polynomial (decompose.rs:137-155)
        per = [decompose_t64(log_b, d, x) for x in v]
        limbs += [[per[i][j] for i in range(n)] for j in range(d)]
    out = []
    for c in range(len(glwe)):
        acc = [0] * n
        for row, limb in zip(rows, limbs):
            prod = fft64_negacyclic_mul(list(row[c]), limb)
            acc = [(x + y) & 0xFFFFFFFFFFFFFFFF for x, y in zip(acc, prod)]
        out.append(acc)
    return out


def tggsw_cmux(log_b, d, rows, ct0, ct1):  # tggsw.rs:114-121: ct0 + external_product(b, ct1 - ct0)
    M = 0xFFFFFFFFFFFFFFFF
    diff = [[(y - x) & M for x, y in zip(p0, p1)] for p0, p1 in zip(ct0, ct1)]
    ext = tggsw_external_product(log_b, d, rows, diff)
    return [[(x + y) & M for x, y in zip(p0, e)] for p0, e in zip(ct0, ext)]


def tlwe_key_switch(log_b, d, ksk_a, ksk_b, a, b):
    """Tlwe::key_switch (tlwe.rs:144-153): a' = ksk.a . limbs, b' = ksk.b . limbs + b over the limb-major (flattened) digits of
    a; the key encrypts the powered-up -sk1 (tlwe.rs:100-111), so the sums are added.  Wrapping u64 arithmetic."""
    M = 0xFFFFFFFFFFFFFFFF
    per = [decompose_t64(log_b, d, x) for x in a]
    digs = [per[i][j] for j in range(d) for i in range(len(a))]
    n = len(ksk_a[0])
    out_a = [0] * n
    out_b = b
    for dg, ra, rb in zip(digs, ksk_a, ksk_b):
        out_a = [(x + dg * y) & M for x, y in zip(out_a, ra)]
        out_b = (out_b + dg * rb) & M
    return out_a, out_b


# ---- scheme/ckks/src/ckks.rs -------------------------------------------------------------------------------------------------------
def _rns_poly(f, moduli, polys):  # apply a per-coefficient RNS map to limb-major polynomials
    n = len(polys[0])
    cols = [f([p[i] for p in polys]) for i in range(n)]
    return [[c[j] for c in cols] for j in range(len(cols[0]))]


def ckks_key_switch(qs_l, ps, ksk_b, ksk_a, ct_b, ct_a):
    """Ckks::key_switch (ckks.rs:284-293): ct_a extended to qs_l ++ ps, multiplied limb-wise with the key (given on the same
    limbs, coefficient form), rescale_k by the special primes; b += ct_b."""
    moduli = list(qs_l) + list(ps)
    ext = _rns_poly(lambda x: rns_extend_bases(list(qs_l), list(ps), x), moduli, ct_a)
    prod_b = [schoolbook_negacyclic(k, e, m) for k, e, m in zip(ksk_b, ext, moduli)]
    prod_a = [schoolbook_negacyclic(k, e, m) for k, e, m in zip(ksk_a, ext, moduli)]
    rb = _rns_poly(lambda x: rns_rescale_k(moduli, len(ps), x), moduli, prod_b)
    ra = _rns_poly(lambda x: rns_rescale_k(moduli, len(ps), x), moduli, prod_a)
    return [[(x + y) % q for x, y in zip(p, c)] for p, c, q in zip(rb, ct_b, qs_l)], ra


def ckks_mul(qs_l, ps, rlk_b, rlk_a, ct0, ct1):
    """Ckks::mul (ckks.rs:255-272): tensor, relinearise d2 with ct_b = 0, add, rescale by the last prime."""
    (b0, a0), (b1, a1) = ct0, ct1
    mul = lambda x, y: [schoolbook_negacyclic(p, r, q) for p, r, q in zip(x, y, qs_l)]
    add = lambda x, y: [[(u + v) % q for u, v in zip(p, r)] for p, r, q in zip(x, y, qs_l)]
    d0, d1, d2 = mul(b0, b1), add(mul(b0, a1), mul(a0, b1)), mul(a0, a1)
    zero = [[0] * len(d2[0]) for _ in qs_l]
    rb, ra = ckks_key_switch(qs_l, ps, rlk_b, rlk_a, zero, d2)
    sb, sa = add(d0, rb), add(d1, ra)
    return (_rns_poly(lambda x: rns_rescale_k(list(qs_l), 1, x), qs_l, sb), _rns_poly(lambda x: rns_rescale_k(list(qs_l), 1, x), qs_l, sa))


def t64_rounding_shr(v, bits):  # decompose.rs:115-118 on a torus word
    return ((v + ((1 << bits) >> 1)) & 0xFFFFFFFFFFFFFFFF) >> bits


def tfhe_bootstrap(log_p, padding, k, bs_log_b, bs_d, ks_log_b, ks_d, brk, ksk_a, ksk_b, v, ct):
    """Bootstrapping::bootstrap (tfhe/bootstrapping.rs:78-110): v = test polynomial over Z_p (N entries), ct = TLWE [n + 1];
    brk[n][(k+1) d][k+1][N].  blind_rotate: acc = (0, encode(v)).rotate(-b~), then acc = cmux(brk_i, acc, acc.rotate(a~_i));
    sample_extract(0) (tglwe.rs:115-127); Tlwe::key_switch."""
    M = 1 << 64
    n_big = len(v)
    log_delta = 64 - (log_p + padding)
    pt = [(int(m) << log_delta) % M for m in v]  # Tlwe::encode (tlwe.rs:113-116)
    rb = 64 - (2 * n_big).bit_length() + 1
    a = [t64_rounding_shr(x, rb) for x in ct[:-1]]
    b = t64_rounding_shr(ct[-1], rb)
    acc = [[0] * n_big for _ in range(k)] + [pt]
    acc = [monomial_mul(p, -b, M) for p in acc]
    for rows, ai in zip(brk, a):
        acc = tggsw_cmux(bs_log_b, bs_d, rows, acc, [monomial_mul(p, ai, M) for p in acc])
    ext_a = []
    for p in acc[:-1]:
        ext_a += [p[0]] + [(-x) % M for x in reversed(p[1:])]
    return tlwe_key_switch(ks_log_b, ks_d, ksk_a, ksk_b, ext_a, acc[-1][0])


def ckks_rotate(qs_l, ps, t, ksk_b, ksk_a, ct_b, ct_a):
    """Ckks::rotate / conjugate (ckks.rs:274-282): X -> X^t on both polynomials (limb-wise, avec.rs:34-50), then key_switch."""
    rb = [automorphism(p, t, q) for p, q in zip(ct_b, qs_l)]
    ra = [automorphism(p, t, q) for p, q in zip(ct_a, qs_l)]
    return ckks_key_switch(qs_l, ps, ksk_b, ksk_a, rb, ra)


def ckks_mul_constant(qs_l, pt, ct_b, ct_a):
    """Ckks::mul_constant (ckks.rs:250-253) on an already encoded plaintext pt [l][N]: (pt * b, pt * a).rescale()."""
    mul = lambda x: [schoolbook_negacyclic(p, r, q) for p, r, q in zip(pt, x, qs_l)]
    res = lambda x: _rns_poly(lambda c: rns_rescale_k(list(qs_l), 1, c), qs_l, x)
    return res(mul(ct_b)), res(mul(ct_a))


def ckks_mul_mat(qs_l, ps, baby, giant, present, pts, ct_b, ct_a):
    """Bootstrapping::mul_mat (ckks/src/bootstrapping.rs:92-108): baby / giant = [(t, (ksk_b, ksk_a) or None)], pts = encoded
    diagonals in row-major (i, j) order of the present pairs.  out = sum_i rot_{g_i}(sum_j mul_constant(pt_ij, rot_{b_j}(ct)))."""
    rot = lambda t, key, b, a: (b, a) if t == 0 else ckks_rotate(qs_l if len(b) == len(qs_l) else qs_l[:len(b)], ps, t, key[0], key[1], b, a)
    add = lambda x, y, mods: [[(u + v) % q for u, v in zip(p, r)] for p, r, q in zip(x, y, mods)]
    rots = [rot(t, key, ct_b, ct_a) for t, key in baby]
    lower = qs_l[:-1]
    out, it = None, iter(pts)
    for i, (t, key) in enumerate(giant):
        inner = None
        for j in range(len(baby)):
            if not present[i][j]:
                continue
            term = ckks_mul_constant(qs_l, next(it), rots[j][0], rots[j][1])
            inner = term if inner is None else (add(inner[0], term[0], lower), add(inner[1], term[1], lower))
        if t != 0:  # keys restricted to the limbs that remain after the rescale
            kb = [key[0][x] for x in list(range(len(lower))) + list(range(len(qs_l), len(qs_l) + len(ps)))]
            ka = [key[1][x] for x in list(range(len(lower))) + list(range(len(qs_l), len(qs_l) + len(ps)))]
            inner = ckks_rotate(lower, ps, t, kb, ka, inner[0], inner[1])
        out = inner if out is None else (add(out[0], inner[0], lower), add(out[1], inner[1], lower))
    return out
