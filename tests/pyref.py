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
    return math.floor(x + 0.5) if x >= 0 else -math.floor(-x + 0.5)


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
