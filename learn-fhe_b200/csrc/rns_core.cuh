// RNS base conversion / rescale building blocks (K14/K15 of SURVEY.md §2), __host__ __device__.
// Reference: util/src/ring/rns.rs — Rns::new/with_ps (287-322), Rns::extend_bases (331-345), RnsRq::rescale_k / round / div
// (99-132).  All integer results are canonical residues, so any exact evaluation order is bit-identical; the only
// order-sensitive part is the f64 overflow estimate u = round(sum_i (1/q_i) * v_i), which is evaluated exactly like the
// reference: sequential sum from 0.0, i ascending, separate multiply and add (no FMA).
#pragma once
#include "fhew_core.cuh"  // f64 helpers
#include "modarith.cuh"

namespace fhe {

static constexpr int RNS_MAXL = 16;  // max source limbs of one base conversion (kept in registers)

// Conversion table of one (source base qs -> target base ps) pair.  Two layouts with identical member names so the
// per-coefficient code is shared: RnsExtTab (pointers; tests/hostsim) and RnsExtTabV (fixed-size arrays, passed BY VALUE as a
// kernel parameter so every constant is a uniform constant-bank operand instead of a global load).  2-D tables use the fixed
// strides RNS_MAXL / RNS_MAXL + 1.
struct RnsExtTab {
    int nq, np;
    int lazy;                  // 2: every p_k in [2^33, 2^59): exact 128-bit sum of the nq <= 16 products, reduced once (c64, m32)
                               // 1: every p_k < 2^58, so nq <= 16 lazy Shoup products ([0, 4 p_k) each) sum below 2^64
    const uint64_t* c64;       // [np] 2^64 mod p_k, Shoup companion in c64_sh; m32 = floor(2^64 / p_k) (< 2^31)
    const uint64_t* c64_sh;
    const uint64_t* m32;
    const Mod64* mq;           // [nq]
    const uint64_t* qhat_inv;  // [nq]  (Q/q_i)^-1 mod q_i, with Shoup companion in qhat_inv_sh
    const uint64_t* qhat_inv_sh;
    const double* frac;        // [nq]  1.0 / q_i
    const Mod64* mp;           // [np]
    const uint64_t* qhat_ps;   // [np][RNS_MAXL]     (Q/q_i) mod p_k
    const uint64_t* qhat_ps_sh;  // [np][RNS_MAXL]   Shoup companions floor(qhat_ps * 2^64 / p_k)
    const uint64_t* uq_ps;     // [np][RNS_MAXL + 1] (u*Q) mod p_k
};
struct RnsExtTabV {
    int nq, np;
    int lazy, pad_;
    Mod64 mq[RNS_MAXL];
    uint64_t qhat_inv[RNS_MAXL], qhat_inv_sh[RNS_MAXL];
    double frac[RNS_MAXL];
    Mod64 mp[RNS_MAXL];
    uint64_t qhat_ps[RNS_MAXL * RNS_MAXL], qhat_ps_sh[RNS_MAXL * RNS_MAXL];
    uint64_t uq_ps[RNS_MAXL * (RNS_MAXL + 1)];
    uint64_t c64[RNS_MAXL], c64_sh[RNS_MAXL], m32[RNS_MAXL];
};

// rescale_k (rns.rs:99-132) over moduli kept (l) ++ dropped (k)
struct RescaleTab {
    RnsExtTab ext;  // dropped -> kept (unused when k == 1)
    int l, k;
    const Mod64* m_all;     // [l + k]
    const uint64_t* ph;     // [l + k]  (P >> 1) mod q_i, P = prod(dropped)
    const uint64_t* pinv;   // [l]      P^-1 mod q_i (+ Shoup companion)
    const uint64_t* pinv_sh;
};
struct RescaleTabV {
    RnsExtTabV ext;
    int l, k;
    Mod64 m_all[2 * RNS_MAXL];
    uint64_t ph[2 * RNS_MAXL];
    uint64_t pinv[RNS_MAXL], pinv_sh[RNS_MAXL];
};

// v mod m.q for an arbitrary 64-bit v
HD uint64_t rns_reduce_u64(const Mod64& m, uint64_t v) {
    if (v < m.q) return v;
    if (m.s >= 32) {
        U128 x;
        x.lo = v;
        x.hi = 0;
        return m.reduce128(x);
    }
    return v % m.q;
}
HD double f64_add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    volatile double r = a + b;
    return r;
#endif
}
HD double u64_to_f64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __ull2double_rn(v);
#else
    return (double)v;
#endif
}

// Rns::extend_bases for one coefficient (rns.rs:331-345): x[i] = residue mod q_i (i < nq) -> y[k] = residue mod p_k.
// `emit(k, y)` receives the outputs; `begin(k)` runs before the products of target k (rescale_k requests the kept limb there, so its
// global-memory latency hides behind them).
// Targets are visited in the order first, 0, 1, ... (skipping first).
template <typename Tab, typename Emit, typename Begin>
HD void rns_extend_coeff(const Tab& T, const uint64_t* x /* [RNS_MAXL], first nq valid */, Emit emit, Begin begin, int first = 0) {
    uint64_t v[RNS_MAXL];
    double acc = 0.0;
#pragma unroll
    for (int i = 0; i < RNS_MAXL; ++i) {
        if (i < T.nq) {
            const Mod64 m = T.mq[i];
            v[i] = m.redq(m.shoup_lazy(x[i], T.qhat_inv[i], T.qhat_inv_sh[i]));
            acc = f64_add_rn(acc, f64_mul_rn(T.frac[i], u64_to_f64(v[i])));
        } else {
            v[i] = 0;
        }
    }
    const uint32_t u = (uint32_t)f64_round_half_away(acc);
    for (int kk = 0; kk < T.np; ++kk) {
        const int k = kk == 0 ? first : (kk <= first ? kk - 1 : kk);
        begin(k);
        const Mod64 m = T.mp[k];
        const uint64_t* qh = T.qhat_ps + (size_t)k * RNS_MAXL;
        uint64_t s = 0;
        if (T.lazy == 2) {
            // sum_i qhat_i * v_i < 16 * 2^59 * 2^64 < 2^128 accumulated exactly (4 wide products per term instead of the 9
            // of a Shoup product), then hi * 2^64 + lo reduced once: hi by a lazy Shoup product with c64 = 2^64 mod p
            // ([0, 4p)), lo by a one-multiply Barrett on its top 32 bits ([0, 3p) for p >= 2^33: the quotient estimate
            // floor(floor(lo / 2^32) * m32 / 2^32) is floor(lo / p) - {0, 1, 2}).  Same canonical value as the reference's
            // per-term canonical Zq arithmetic.
            unsigned __int128 acc = 0;
#pragma unroll
            for (int i = 0; i < RNS_MAXL; ++i) {
                if (i >= T.nq) break;
                acc += (unsigned __int128)qh[i] * v[i];
            }
            const uint64_t hi = (uint64_t)(acc >> 64), lo = (uint64_t)acc;
            const uint64_t r1 = T.c64[k] * hi - mulhi_u64_approx(T.c64_sh[k], hi) * m.q;
            const uint64_t r0 = lo - (uint64_t)mulhi_u32((uint32_t)(lo >> 32), (uint32_t)T.m32[k]) * m.q;
            s = r1 + r0;  // < 7p < 2^62
            s = umin_(s, s - 4 * m.q);
            s = umin_(s, s - m.q2);
            s = umin_(s, s - m.q);
        } else if (T.lazy) {
            // constant * variable products by Shoup (valid for ANY 64-bit v, so v_i needs no reduction mod p_k first); the
            // canonical value of the sum is what the reference's per-term canonical Zq arithmetic yields
            const uint64_t* qs = T.qhat_ps_sh + (size_t)k * RNS_MAXL;
#pragma unroll
            for (int i = 0; i < RNS_MAXL; ++i) {
                if (i >= T.nq) break;  // uniform early exit: predicated-off iterations would still cost issue slots
                s += qh[i] * v[i] - mulhi_u64_approx(qs[i], v[i]) * m.q;  // [0, 4 p_k): quotient off by <= 2
            }
            s = rns_reduce_u64(m, s);
        } else {
#pragma unroll
            for (int i = 0; i < RNS_MAXL; ++i) {
                if (i >= T.nq) break;
                s = m.add(s, m.mul(qh[i], rns_reduce_u64(m, v[i])));
            }
        }
        emit(k, m.sub(s, T.uq_ps[(size_t)k * (RNS_MAXL + 1) + u]));
    }
}

template <typename Tab, typename Emit>
HD void rns_extend_coeff(const Tab& T, const uint64_t* x, Emit emit) {
    rns_extend_coeff(T, x, emit, [](int) {});
}

// rescale_k for one coefficient: x(i) reads limb i of the input (already including any pre-addend, canonical);
// emit(i, y) receives the kept limbs before any post-addend.
template <typename Tab, typename Load, typename Emit>
HD void rns_rescale_coeff(const Tab& R, Load x, Emit emit) {
    const int l = R.l, k = R.k;
    // round(): s_i = x_i + (P >> 1) mod q_i, for every limb (rns.rs:120-125)
    auto rounded = [&](int i) { return R.m_all[i].add(x(i), R.ph[i]); };
    if (k == 1) {
        const uint64_t vp = rounded(l);  // non-centred value of the dropped limb (rns.rs:109-111)
        // this path is bound by global-memory latency: issue every load before the first use (static register indices)
        uint64_t xk[RNS_MAXL];
#pragma unroll
        for (int i = 0; i < RNS_MAXL; ++i) xk[i] = i < l ? rounded(i) : 0;
#pragma unroll
        for (int i = 0; i < RNS_MAXL; ++i) {
            if (i >= l) break;
            const Mod64 m = R.m_all[i];
            const uint64_t v = m.sub(xk[i], rns_reduce_u64(m, vp));
            emit(i, m.redq(m.shoup_lazy(v, R.pinv[i], R.pinv_sh[i])));
        }
    } else {
        uint64_t xs[RNS_MAXL];
#pragma unroll
        for (int j = 0; j < RNS_MAXL; ++j) xs[j] = j < k ? rounded(l + j) : 0;
        // (loading the kept limbs up front and unrolling the target loop was measured: more registers, lower occupancy, slower;
        // one limb ahead of its use costs two registers)
        uint64_t kept = 0;
        rns_extend_coeff(
            R.ext, xs,
            [&](int i, uint64_t y) {
                const Mod64 m = R.m_all[i];
                const uint64_t v = m.sub(m.add(kept, R.ph[i]), y);
                emit(i, m.redq(m.shoup_lazy(v, R.pinv[i], R.pinv_sh[i])));  // div(): * P^-1 (rns.rs:127-132)
            },
            [&](int i) { kept = x(i); });
    }
}

// rescale_k by the k dropped limbs of R followed by rescale_k(1) of the result's last limb (tables of the second step in R1:
// R1.l = R.l - 1, R1.k = 1) for one coefficient, without materialising the intermediate: exactly rns_rescale_coeff(R1, .)
// applied to the limbs rns_rescale_coeff(R, .) emits.  The last kept limb is converted first because every other output needs it.
struct Rescale1TabV {
    int l;
    int fast;  // 1: every modulus < 2^61 and the dropped limb of the second step < 2 q_i for every kept q_i (lazy finishing, see below)
    Mod64 m_all[RNS_MAXL + 1];
    uint64_t ph[RNS_MAXL + 1];
    uint64_t pinv[RNS_MAXL], pinv_sh[RNS_MAXL];
};
template <typename Tab, typename Tab1, typename Load, typename Emit>
HD void rns_rescale2_coeff(const Tab& R, const Tab1& R1, Load x, Emit emit) {
    const int l = R.l, k = R.k;
    uint64_t xs[RNS_MAXL];
#pragma unroll
    for (int j = 0; j < RNS_MAXL; ++j) xs[j] = j < k ? R.m_all[l + j].add(x(l + j), R.ph[l + j]) : 0;
    uint64_t kept = 0, vp = 0;
    rns_extend_coeff(
        R.ext, xs,
        [&](int i, uint64_t y) {
            const Mod64 m = R.m_all[i];
            const uint64_t v = m.sub(m.add(kept, R.ph[i]), y);
            const uint64_t r = m.redq(m.shoup_lazy(v, R.pinv[i], R.pinv_sh[i]));  // limb i of rescale_k(x, k)
            if (i == l - 1) {
                vp = m.add(r, R1.ph[l - 1]);  // rounded dropped limb of the second step (rns.rs:109-111)
            } else if (R1.fast) {
                // same canonical result with lazy intermediates: u = kept + P/2 + q - y < 3q, t = u P^-1 in [0, 4q) (3-multiply high
                // word), w = t + q_last/2 + q - (vp mod q) < 6q < 2^64, and one canonical reduction at the very end
                const uint64_t u = kept + R.ph[i] + m.q - y;
                const uint64_t t = R.pinv[i] * u - mulhi_u64_approx(R.pinv_sh[i], u) * m.q;
                const uint64_t w = t + R1.ph[i] + m.q - umin_(vp, vp - m.q);
                emit(i, m.canon4(R1.pinv[i] * w - mulhi_u64_approx(R1.pinv_sh[i], w) * m.q));
            } else {
                const uint64_t w = m.sub(m.add(r, R1.ph[i]), rns_reduce_u64(m, vp));
                emit(i, m.redq(m.shoup_lazy(w, R1.pinv[i], R1.pinv_sh[i])));
            }
        },
        [&](int i) { kept = x(i); }, l - 1);
}

}  // namespace fhe
