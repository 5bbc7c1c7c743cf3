// Host-only construction of the f64 twiddle tables of the torus FFT path, exactly like compute_twiddle
// (util/src/ring/fft/c64.rs:98-108): cis((i as f64 * PI) / n as f64), conjugates, and the bit-reversed chunk twiddles
// fft_in_place indexes (ring/fft.rs:9-35).  Shared by tfhe.cu and tests/hostsim.
#pragma once
#include <cmath>
#include <cstdint>
#include <vector>

#include "tfhe_core.cuh"

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

}  // namespace fhe
