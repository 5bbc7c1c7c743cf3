// ORACLE — TEST INFRASTRUCTURE ONLY (see orc_util.hpp header).
// Restatement of util/src/ring/rns.rs (RnsRq: extend_bases / switch_bases / rescale_k / mul).
// BigUint constants of the reference (rns.rs:287-322) are mathematically determined residues; they
// are computed here with modular arithmetic only (no big integers needed).
#pragma once
#include "orc_util.hpp"

namespace orc {

// A limb-major RNS polynomial: limbs[i] is the polynomial mod qs[i]  (rns.rs:21)
struct RnsPoly {
    Vec qs;
    std::vector<Vec> limbs;
    size_t n() const { return limbs.empty() ? 0 : limbs[0].size(); }
};

static inline u64 prod_mod(const Vec& xs, u64 p, size_t skip = (size_t)-1) {
    u64 r = 1 % p;
    for (size_t j = 0; j < xs.size(); ++j)
        if (j != skip) r = zq_mul(p, r, xs[j] % p);
    return r;
}

// rns.rs:278-322 struct Rns / new / with_ps
struct Rns {
    Vec qs, ps;
    Vec q_hats_inv_qs;               // (Q/q_i)^-1 mod q_i
    std::vector<double> q_fracs;     // 1.0 / q_i
    std::vector<Vec> q_hats_ps;      // [p][i] = (Q/q_i) mod p
    std::vector<Vec> uq_ps;          // [p][u] = (u*Q) mod p, u = 0..=len(qs)
    explicit Rns(const Vec& qs_) : qs(qs_) {
        for (size_t i = 0; i < qs.size(); ++i) {
            q_hats_inv_qs.push_back(zq_inv(qs[i], prod_mod(qs, qs[i], i)));
            q_fracs.push_back(1.0 / (double)qs[i]);
        }
    }
    Rns& with_ps(const Vec& ps_) {
        ps = ps_;
        for (u64 p : ps) {
            Vec row;
            for (size_t i = 0; i < qs.size(); ++i) row.push_back(prod_mod(qs, p, i));
            q_hats_ps.push_back(row);
            u64 qmod = prod_mod(qs, p);
            Vec uq;
            for (size_t u = 0; u <= qs.size(); ++u) uq.push_back(zq_mul(p, (u64)u % p, qmod));
            uq_ps.push_back(uq);
        }
        return *this;
    }
    // rns.rs:331-345 extend_bases for one coefficient.  f64 sum is sequential, i ascending, no FMA.
    void extend_bases(const u64* vqs, u64* vps) const {
        size_t l = qs.size();
        Vec vs(l);
        for (size_t i = 0; i < l; ++i) vs[i] = zq_mul(qs[i], vqs[i], q_hats_inv_qs[i]);
        // Iterator::sum::<f64>() starts from 0.0 (std impl: fold(0.0, |a, b| a + b)); -0.0 vs 0.0 irrelevant
        volatile double acc = 0.0;
        for (size_t i = 0; i < l; ++i) {
            volatile double term = q_fracs[i] * (double)vs[i];
            acc = acc + term;
        }
        size_t u = (size_t)std::round(acc);
        for (size_t k = 0; k < ps.size(); ++k) {
            u64 p = ps[k];
            // Dot (misc.rs:50-62): Sum starts from the first product, then adds the rest
            u64 s = zq_mul(p, q_hats_ps[k][0], zq_from_u64(p, vs[0]));
            for (size_t i = 1; i < l; ++i) s = zq_add(p, s, zq_mul(p, q_hats_ps[k][i], zq_from_u64(p, vs[i])));
            vps[k] = zq_sub(p, s, uq_ps[k][u]);
        }
    }
};
// NOTE on `q_hats_pi.dot(&vs)` (rns.rs:344): lhs items are &Zq (mod p), rhs items are &u64, so each
// product is Zq * u64 = Zq * Zq::from_u64(p, v) (zq.rs impl_op_with_primitive) — as restated above.

// rns.rs:83-91 extend_bases: appends limbs for ps
static inline RnsPoly rns_extend_bases(const RnsPoly& x, const Vec& ps) {
    Rns rns(x.qs);
    rns.with_ps(ps);
    size_t n = x.n(), l = x.qs.size();
    RnsPoly out = x;
    for (u64 p : ps) {
        out.qs.push_back(p);
        out.limbs.emplace_back(n, 0);
    }
    Vec vq(l), vp(ps.size());
    for (size_t c = 0; c < n; ++c) {
        for (size_t i = 0; i < l; ++i) vq[i] = x.limbs[i][c];
        rns.extend_bases(vq.data(), vp.data());
        for (size_t k = 0; k < ps.size(); ++k) out.limbs[l + k][c] = vp[k];
    }
    return out;
}
// rns.rs:93-97 switch_bases
static inline RnsPoly rns_switch_bases(const RnsPoly& x, const Vec& ps) {
    RnsPoly e = rns_extend_bases(x, ps);
    RnsPoly out;
    size_t l = x.qs.size();
    out.qs.assign(e.qs.begin() + l, e.qs.end());
    out.limbs.assign(e.limbs.begin() + l, e.limbs.end());
    return out;
}
// floor(P/2) mod q for P = prod(ps) (P odd): (P-1)/2 mod q = (P mod q - 1) * 2^-1 mod q   (rns.rs:120-125 `p >> 1`)
static inline u64 p_half_mod(const Vec& ps, u64 q) {
    bool odd = true;
    for (u64 p : ps) odd = odd && (p & 1);
    if (!odd) throw std::runtime_error("p_half_mod expects odd moduli");
    u64 pm = prod_mod(ps, q);
    return zq_mul(q, zq_sub(q, pm, 1 % q), zq_inv(q, 2 % q));
}
// rns.rs:99-132 rescale_k (round, subtract, div)
static inline RnsPoly rns_rescale_k(const RnsPoly& x, size_t k) {
    assert(k > 0 && k < x.qs.size());
    size_t l = x.qs.size() - k, n = x.n();
    Vec qs(x.qs.begin(), x.qs.begin() + l), ps(x.qs.begin() + l, x.qs.end());
    RnsPoly s = x;
    // round(): add (P>>1) mod q_i to EVERY limb (including the ones to be dropped)
    for (size_t i = 0; i < s.qs.size(); ++i) {
        u64 ph = p_half_mod(ps, s.qs[i]);
        // for the dropped limbs P mod p_j == 0 for j in ps... (P>>1) mod p_j is still well defined:
        // p_half_mod handles it since prod_mod gives 0 -> (0-1)/2 mod p_j = (p_j-1)/2
        for (size_t c = 0; c < n; ++c) s.limbs[i][c] = zq_add(s.qs[i], s.limbs[i][c], ph);
    }
    RnsPoly out;
    out.qs = qs;
    out.limbs.assign(s.limbs.begin(), s.limbs.begin() + l);
    if (k == 1) {
        // rns.rs:109-111: *vq -= vp.to_u64()  (Zq -= u64: from_u64 reduces the NON-centred value)
        const Vec& rp = s.limbs[l];
        for (size_t i = 0; i < l; ++i)
            for (size_t c = 0; c < n; ++c) out.limbs[i][c] = zq_sub(qs[i], out.limbs[i][c], zq_from_u64(qs[i], rp[c]));
    } else {
        RnsPoly rps;
        rps.qs = ps;
        rps.limbs.assign(s.limbs.begin() + l, s.limbs.end());
        RnsPoly sw = rns_switch_bases(rps, qs);
        for (size_t i = 0; i < l; ++i)
            for (size_t c = 0; c < n; ++c) out.limbs[i][c] = zq_sub(qs[i], out.limbs[i][c], sw.limbs[i][c]);
    }
    // div(): multiply by P^-1 mod q_i
    for (size_t i = 0; i < l; ++i) {
        u64 pinv = zq_inv(qs[i], prod_mod(ps, qs[i]));
        for (size_t c = 0; c < n; ++c) out.limbs[i][c] = zq_mul(qs[i], out.limbs[i][c], pinv);
    }
    return out;
}
// rns.rs:143-158 MulAssign: keep lhs limbs whose modulus is also in rhs (lhs order), multiply limb-wise
// (coefficient form => ring.rs:256-264 => nega_cyclic_ntt_mul_assign, 3 transforms per limb)
static inline RnsPoly rns_mul(const RnsPoly& a, const RnsPoly& b) {
    RnsPoly out;
    for (size_t i = 0; i < a.qs.size(); ++i) {
        for (size_t j = 0; j < b.qs.size(); ++j) {
            if (a.qs[i] == b.qs[j]) {
                Vec limb = a.limbs[i];
                nega_cyclic_ntt_mul_assign(a.qs[i], limb.data(), b.limbs[j].data(), limb.size());
                out.qs.push_back(a.qs[i]);
                out.limbs.push_back(std::move(limb));
            }
        }
    }
    return out;
}
static inline RnsPoly rns_add(const RnsPoly& a, const RnsPoly& b) {
    assert(a.qs == b.qs);
    RnsPoly out = a;
    for (size_t i = 0; i < a.qs.size(); ++i)
        for (size_t c = 0; c < a.n(); ++c) out.limbs[i][c] = zq_add(a.qs[i], a.limbs[i][c], b.limbs[i][c]);
    return out;
}
static inline RnsPoly rns_sub(const RnsPoly& a, const RnsPoly& b) {
    assert(a.qs == b.qs);
    RnsPoly out = a;
    for (size_t i = 0; i < a.qs.size(); ++i)
        for (size_t c = 0; c < a.n(); ++c) out.limbs[i][c] = zq_sub(a.qs[i], a.limbs[i][c], b.limbs[i][c]);
    return out;
}
static inline RnsPoly rns_neg(const RnsPoly& a) {
    RnsPoly out = a;
    for (size_t i = 0; i < a.qs.size(); ++i)
        for (size_t c = 0; c < a.n(); ++c) out.limbs[i][c] = zq_neg(a.qs[i], a.limbs[i][c]);
    return out;
}
static inline RnsPoly rns_zero(const Vec& qs, size_t n) {
    RnsPoly out;
    out.qs = qs;
    out.limbs.assign(qs.size(), Vec(n, 0));
    return out;
}
// rns.rs:62-64 from_i64
static inline RnsPoly rns_from_i64(const Vec& qs, const std::vector<i64>& v) {
    RnsPoly out = rns_zero(qs, v.size());
    for (size_t i = 0; i < qs.size(); ++i)
        for (size_t c = 0; c < v.size(); ++c) out.limbs[i][c] = zq_from_i64(qs[i], v[c]);
    return out;
}
// rns.rs:74-77 automorphism
static inline RnsPoly rns_automorphism(const RnsPoly& a, i64 t) {
    RnsPoly out = a;
    for (size_t i = 0; i < a.qs.size(); ++i) out.limbs[i] = automorphism_zq(a.qs[i], a.limbs[i].data(), a.n(), t);
    return out;
}

}  // namespace orc
