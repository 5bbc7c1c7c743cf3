// Host-only construction of the RNS constant tables (Rns::new / with_ps, util/src/ring/rns.rs:287-322; round / div
// constants of rescale_k, rns.rs:120-132).  The reference computes them with BigUint; they are mathematically determined
// residues, computed here with 128-bit modular arithmetic only.  Shared by ckks.cu and tests/hostsim.
#pragma once
#include <cstdlib>
#include <cstdint>
#include <vector>

#include "modarith.cuh"
#include "rns_core.cuh"

namespace fhe {

inline Mod64 host_make_mod64(uint64_t q) {
    Mod64 m;
    m.q = q;
    m.q2 = 2 * q;
    unsigned s = 0;
    while (s < 64 && (q >> s)) ++s;
    if (s < 2) s = 2;
    m.s = s;
    m.mu = (uint64_t)((((u128_t)1) << (2 * s)) / q);
    return m;
}
inline uint64_t host_prod_mod(const std::vector<uint64_t>& xs, uint64_t p, size_t skip = (size_t)-1) {
    uint64_t r = 1 % p;
    for (size_t j = 0; j < xs.size(); ++j)
        if (j != skip) r = host_mulmod(r, xs[j] % p, p);
    return r;
}
inline uint64_t host_inv_any(uint64_t a, uint64_t q) {  // extended Euclid: q need not be prime; 0 if not invertible
    __int128 t = 0, nt = 1, r = q, nr = a % q;
    while (nr != 0) {
        __int128 qu = r / nr;
        __int128 tmp = t - qu * nt;
        t = nt;
        nt = tmp;
        tmp = r - qu * nr;
        r = nr;
        nr = tmp;
    }
    if (r != 1) return 0;
    if (t < 0) t += q;
    return (uint64_t)t;
}

struct RnsExtHost {
    std::vector<Mod64> mq, mp;
    std::vector<uint64_t> qhat_inv, qhat_inv_sh, qhat_ps, qhat_ps_sh, uq_ps, c64, c64_sh, m32;
    int lazy = 1;
    std::vector<double> frac;
    void build(const std::vector<uint64_t>& qs, const std::vector<uint64_t>& ps) {
        const size_t nq = qs.size(), np = ps.size();
        mq.resize(nq);
        mp.resize(np);
        qhat_inv.resize(nq);
        qhat_inv_sh.resize(nq);
        frac.resize(nq);
        qhat_ps.assign(np * RNS_MAXL, 0);
        qhat_ps_sh.assign(np * RNS_MAXL, 0);
        lazy = nq <= 16 ? 1 : 0;
        bool wide = nq <= 16;
        c64.assign(np, 0);
        c64_sh.assign(np, 0);
        m32.assign(np, 0);
        uq_ps.assign(np * (RNS_MAXL + 1), 0);
        for (size_t i = 0; i < nq; ++i) {
            mq[i] = host_make_mod64(qs[i]);
            qhat_inv[i] = host_inv_any(host_prod_mod(qs, qs[i], i), qs[i]);
            qhat_inv_sh[i] = host_shoup64(qhat_inv[i], qs[i]);
            frac[i] = 1.0 / (double)qs[i];
        }
        for (size_t k = 0; k < np; ++k) {
            mp[k] = host_make_mod64(ps[k]);
            if (ps[k] >= (1ull << 58)) lazy = 0;
            if (ps[k] < (1ull << 33) || ps[k] >= (1ull << 59)) wide = false;
            c64[k] = (uint64_t)((((u128_t)1) << 64) % ps[k]);
            c64_sh[k] = host_shoup64(c64[k], ps[k]);
            m32[k] = (uint64_t)((((u128_t)1) << 64) / ps[k]);
            for (size_t i = 0; i < nq; ++i) {
                qhat_ps[k * RNS_MAXL + i] = host_prod_mod(qs, ps[k], i);
                qhat_ps_sh[k * RNS_MAXL + i] = host_shoup64(qhat_ps[k * RNS_MAXL + i], ps[k]);
            }
            const uint64_t qmod = host_prod_mod(qs, ps[k]);
            for (size_t u = 0; u <= nq; ++u) uq_ps[k * (RNS_MAXL + 1) + u] = host_mulmod(u % ps[k], qmod, ps[k]);
        }
        if (wide && getenv("FHE_B200_RNS_NO_WIDE") == nullptr) lazy = 2;
    }
    // by-value form (kernel parameter); requires nq, np <= RNS_MAXL
    void fill(RnsExtTabV& t) const {
        t.nq = (int)mq.size();
        t.np = (int)mp.size();
        t.lazy = lazy;
        t.pad_ = 0;
        for (size_t i = 0; i < mq.size(); ++i) {
            t.mq[i] = mq[i];
            t.qhat_inv[i] = qhat_inv[i];
            t.qhat_inv_sh[i] = qhat_inv_sh[i];
            t.frac[i] = frac[i];
        }
        for (size_t k = 0; k < mp.size(); ++k) {
            t.mp[k] = mp[k];
            t.c64[k] = c64[k];
            t.c64_sh[k] = c64_sh[k];
            t.m32[k] = m32[k];
        }
        for (size_t i = 0; i < qhat_ps.size(); ++i) {
            t.qhat_ps[i] = qhat_ps[i];
            t.qhat_ps_sh[i] = qhat_ps_sh[i];
        }
        for (size_t i = 0; i < uq_ps.size(); ++i) t.uq_ps[i] = uq_ps[i];
    }
    RnsExtTab view() const {
        RnsExtTab t;
        t.nq = (int)mq.size();
        t.np = (int)mp.size();
        t.mq = mq.data();
        t.qhat_inv = qhat_inv.data();
        t.qhat_inv_sh = qhat_inv_sh.data();
        t.frac = frac.data();
        t.mp = mp.data();
        t.qhat_ps = qhat_ps.data();
        t.qhat_ps_sh = qhat_ps_sh.data();
        t.lazy = lazy;
        t.c64 = c64.data();
        t.c64_sh = c64_sh.data();
        t.m32 = m32.data();
        t.uq_ps = uq_ps.data();
        return t;
    }
};

// rescale_k over moduli `all` = kept (l) ++ dropped (k)
struct RescaleHost {
    std::vector<uint64_t> kept, dropped, ph, pinv, pinv_sh;
    std::vector<Mod64> m_all;
    void build(const std::vector<uint64_t>& all, size_t k) {
        const size_t l = all.size() - k;
        kept.assign(all.begin(), all.begin() + l);
        dropped.assign(all.begin() + l, all.end());
        m_all.resize(all.size());
        ph.resize(all.size());
        pinv.resize(l);
        pinv_sh.resize(l);
        for (size_t i = 0; i < all.size(); ++i) {
            const uint64_t q = all[i], pm = host_prod_mod(dropped, q);
            m_all[i] = host_make_mod64(q);
            // (P >> 1) mod q = (P - 1) / 2 mod q for odd P (if q divides P this is (q - 1) / 2)
            ph[i] = host_mulmod((pm + q - 1) % q, host_inv_any(2 % q, q), q);
        }
        for (size_t i = 0; i < l; ++i) {
            pinv[i] = host_inv_any(host_prod_mod(dropped, kept[i]), kept[i]);
            pinv_sh[i] = host_shoup64(pinv[i], kept[i]);
        }
    }
};

}  // namespace fhe
