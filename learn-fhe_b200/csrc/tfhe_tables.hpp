// Host-only construction of the f64 twiddle tables of the torus FFT path, exactly like compute_twiddle
// (util/src/ring/fft/c64.rs:98-108): cis((i as f64 * PI) / n as f64), conjugates, and the bit-reversed chunk twiddles
// fft_in_place indexes (ring/fft.rs:9-35).  Shared by tfhe.cu and tests/hostsim.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "tfhe_core.cuh"
#include "tfhe_fast.cuh"

namespace fhe {

// layout: [tw (m) | tw_inv (m) | tw_bo (mb) | tw_inv_bo (mb)], m = n/2, mb = max(m/2, 1)
struct FftTabHost {
    std::vector<Cx> data;
    size_t m = 0, mb = 0;
    unsigned log_n = 0;
    void build(unsigned log_n_) {
        log_n = log_n_;
        const size_t n = (size_t)1 << log_n;
        m = n / 2;
        mb = m / 2 > 1 ? m / 2 : 1;
        data.assign(2 * m + 2 * mb, Cx{0.0, 0.0});
        for (size_t j = 0; j < m; ++j) {
            volatile double num = (double)j * M_PI;
            const double ang = num / (double)n;
            data[j] = Cx{std::cos(ang), std::sin(ang)};
            data[m + j] = Cx{data[j].re, -data[j].im};
        }
        unsigned lgm = 0;
        while (((size_t)1 << lgm) < m) ++lgm;
        for (size_t c = 0; c < mb; ++c) {  // bit_reverse(twiddle(m))[c] = cis(brev(c) * pi / m)
            size_t i = 0;
            for (unsigned b = 0; b < lgm; ++b)
                if (c & ((size_t)1 << b)) i |= (size_t)1 << (lgm - 1 - b);
            volatile double num = (double)i * M_PI;
            const double ang = num / (double)m;
            const Cx t{std::cos(ang), std::sin(ang)};
            data[2 * m + c] = t;
            data[2 * m + mb + c] = Cx{t.re, -t.im};
        }
    }
    FftTab view(const Cx* base) const {
        FftTab T;
        T.lg = (int)log_n - 1;
        T.tw = base;
        T.tw_inv = base + m;
        T.tw_bo = base + 2 * m;
        T.tw_inv_bo = base + 2 * m + mb;
        T.m_inv = 1.0 / (double)m;
        return T;
    }
};

// Tables of the bounded-error fast path (tfhe_fast.cuh) for ring degree n = 2^log_n, m = n/2 = 2^lg, zeta = e^(i pi / n):
// layout [W (m) | V (m/2, >= 1) | U (m) | W2 | W3 | V4 | V5] (the last four are the per-pass copies for the split r1 + r2 + r3).
// Angles are evaluated in long double and rounded once.
struct FastFftTabHost {
    std::vector<Cx> data;
    size_t m = 0, mv = 0, o_w2 = 0, o_w3 = 0, o_v4 = 0, o_v5 = 0;
    unsigned log_n = 0;
    Cx u0[16];
    static Cx cis_pi(long double num, long double den) {
        const long double ang = 3.14159265358979323846264338327950288L * num / den;
        return Cx{(double)cosl(ang), (double)sinl(ang)};
    }
    void build(unsigned log_n_, int r1, int r2, int r3) {
        log_n = log_n_;
        const size_t n = (size_t)1 << log_n;
        m = n / 2;
        mv = m / 2 > 1 ? m / 2 : 1;
        data.assign(2 * m + mv, Cx{1.0, 0.0});
        unsigned lg = 0;
        while (((size_t)1 << lg) < m) ++lg;
        // exponent of the root of chunk c at level l: e(0, 0) = m; e(l+1, 2c) = e(l, c) / 2; e(l+1, 2c+1) = e(l, c) / 2 + n
        std::vector<uint64_t> e{(uint64_t)m}, nx;
        for (unsigned l = 0; l < lg; ++l) {
            nx.assign(e.size() * 2, 0);
            for (size_t c = 0; c < e.size(); ++c) {
                nx[2 * c] = (e[c] / 2) % (2 * n);
                nx[2 * c + 1] = (e[c] / 2 + n) % (2 * n);
                data[((size_t)1 << l) + c] = cis_pi((long double)nx[2 * c], (long double)n);
            }
            e.swap(nx);
        }
        for (size_t i = 0; i < mv; ++i) data[m + i] = cis_pi(-2.0L * (long double)i, (long double)m);
        for (size_t p = 0; p < m; ++p) {
            const Cx u = cis_pi(-(long double)p, (long double)n);
            data[m + mv + p] = Cx{u.re / (double)m, u.im / (double)m};
        }
        for (size_t i = 0; i < 16; ++i) u0[i] = (i << (lg - r1)) < m ? cis_pi(-(long double)(i << (lg - r1)), (long double)n) : Cx{1.0, 0.0};
        const Cx *W = data.data(), *V = data.data() + m;
        std::vector<Cx> ext;
        o_w2 = data.size();
        for (int u = 0; u < r2; ++u)
            for (int top = 0; top < (1 << u); ++top)
                for (size_t hi = 0; hi < ((size_t)1 << r1); ++hi) ext.push_back(W[((size_t)1 << (r1 + u)) + (hi << u) + top]);
        o_w3 = data.size() + ext.size();
        for (int u = 0; u < r3; ++u)
            for (int top = 0; top < (1 << u); ++top)
                for (size_t g = 0; g < (m >> r3); ++g) ext.push_back(W[((size_t)1 << (r1 + r2 + u)) + (g << u) + top]);
        o_v4 = data.size() + ext.size();
        for (int uu = 0; uu < r2; ++uu)
            for (int low = 0; low < (1 << uu); ++low)
                for (size_t lo = 0; lo < ((size_t)1 << r3); ++lo) ext.push_back(V[((((size_t)low << r3) | lo) << (r1 + r2 - 1 - uu))]);
        o_v5 = data.size() + ext.size();
        for (int uu = 0; uu < r1; ++uu)
            for (size_t lo = 0; lo < (m >> r1); ++lo) ext.push_back(V[lo << (r1 - 1 - uu)]);
        data.insert(data.end(), ext.begin(), ext.end());
    }
    FastFftTab view(const Cx* base) const {
        FastFftTab T;
        T.W = base;
        T.V = base + m;
        T.U = base + m + mv;
        T.W2 = base + o_w2;
        T.W3 = base + o_w3;
        T.V4 = base + o_v4;
        T.V5 = base + o_v5;
        for (size_t i = 0; i < 16; ++i) T.w0[i] = i < m ? data[i] : Cx{1.0, 0.0};
        for (size_t i = 0; i < 16; ++i) T.u0[i] = u0[i];
        return T;
    }
};

}  // namespace fhe
