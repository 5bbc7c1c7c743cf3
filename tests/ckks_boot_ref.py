"""TEST INFRASTRUCTURE: Bootstrapping::mul_mat / slot_to_coeff / coeff_to_slot (scheme/ckks/src/bootstrapping.rs:73-108) composed
from the ORACLE's Ckks::rotate / mul_constant / add in the reference's order, on the plans and encoded diagonals of
learn-fhe_b200/ckks_bootstrapping.py - the checker of the device chain - plus Ckks::decode for the functional check."""
import numpy as np


def rns_add(qs, a, b):
    lv = a.shape[1]
    return np.stack([np.stack([(a[h, t] + b[h, t]) % np.uint64(qs[t]) for t in range(lv)]) for h in range(2)])


def mul_constant(orc, K, pt, ct):
    """Ckks::mul_constant (ckks.rs:250-253) on an already encoded plaintext: (pt * b, pt * a).rescale()."""
    level = ct.shape[1]
    prod = np.stack([np.stack([orc.ntt_mul(K.qs[t], ct[h, t], pt[t]) for t in range(level)]) for h in range(2)])
    return K.rescale(prod)


def mul_mat(orc, K, key_index, plan, ct):
    """bootstrapping.rs:92-108: ct_rot[j] = rotate(j, ct); inner_i = sum_j mul_constant(diag_rot(i, j), ct_rot[j]);
    out = sum_i rotate(i, inner_i).  key_index[j] = index of the oracle's automorphism key for rotation index j."""
    rot = lambda j, c: c if j == 0 else K.key_switch(key_index[j], c, apply_auto=True)
    ct_rot = {j: rot(j, ct) for j in plan["baby"]}
    out, p = None, 0
    for gi, i in enumerate(plan["giant"]):
        inner = None
        for bj, j in enumerate(plan["baby"]):
            if not plan["present"][gi, bj]:
                continue
            term = mul_constant(orc, K, plan["pts"][p], ct_rot[j])
            p += 1
            inner = term if inner is None else rns_add(K.qs, inner, term)
        g = rot(i, inner)
        out = g if out is None else rns_add(K.qs, out, g)
    return out


def chain(orc, K, key_index, bparam, which, ct):
    mats = bparam.sfft_fmats if which == "sfft" else bparam.sifft_fmats
    for idx in range(len(mats) - 1, -1, -1):
        ct = mul_mat(orc, K, key_index, bparam.plan(mats[idx], ct.shape[1]), ct)
    return ct


def crt_centered(qs, limbs):
    """[level][n] residues -> centred integers (python ints)."""
    big_q = 1
    for q in qs:
        big_q *= q
    out = []
    for c in range(limbs.shape[1]):
        x = 0
        for q, r in zip(qs, limbs[:, c]):
            m = big_q // q
            x = (x + int(r) * m * pow(m, -1, q)) % big_q
        out.append(x - big_q if x > big_q // 2 else x)
    return out


def sfft(z):
    """sfft.rs:7-19 in complex128 (decode side of the functional check)."""
    z = np.array(z, dtype=np.complex128)
    n = len(z)
    log_n = n.bit_length() - 1
    if n > 2:
        z = z[[int(format(i, "0%db" % log_n)[::-1], 2) for i in range(n)]]
    for log_m in range(log_n):
        m = 1 << log_m
        pw, w = 1, []
        for _ in range(m):
            w.append(pw)
            pw = pw * 5 % (8 * m)
        t = np.exp(2j * np.pi * np.array(w) / (8 * m))
        z = z.reshape(-1, 2 * m)
        a, tb = z[:, :m], z[:, m:] * t[None, :]
        z = np.concatenate([a + tb, a - tb], axis=1).reshape(-1)
    return z


def decode(param, scale_pow, coeffs):
    """Ckks::decode (ckks.rs:201-213): slots = sfft((re + i im) / scale^scale_pow)."""
    l = len(coeffs) // 2
    s = float(param.qs[param.big_l - 1]) ** scale_pow
    return sfft(np.array([complex(coeffs[i] / s, coeffs[l + i] / s) for i in range(l)]))


def bit_reverse(v):
    n = len(v)
    if n <= 2:
        return np.array(v)
    log_n = n.bit_length() - 1
    return np.array(v)[[int(format(i, "0%db" % log_n)[::-1], 2) for i in range(n)]]
