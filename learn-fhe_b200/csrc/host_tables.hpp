// Host-only table construction shared by the context (ctx.cu) and the CPU simulation used in tests
// (tests/hostsim).  Root choice and ordering follow util/src/zq.rs:99-109 and
// util/src/ring/fft/zq.rs:58-67: tw[j] = omega^(brev_{s-1}(j)), omega = g0^((q-1) >> s).
#pragma once
#include <cstdint>
#include <vector>

#include "modarith.cuh"

namespace fhe {

inline bool host_is_prime_u64(uint64_t n) {
    if (n < 2) return false;
    static const uint64_t bases[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (uint64_t p : bases)
        if (n % p == 0) return n == p;
    uint64_t d = n - 1;
    int r = 0;
    while (!(d & 1)) {
        d >>= 1;
        ++r;
    }
    for (uint64_t a : bases) {
        uint64_t x = host_powmod(a, d, n);
        if (x == 1 || x == n - 1) continue;
        bool composite = true;
        for (int i = 1; i < r; ++i) {
            x = host_mulmod(x, x, n);
            if (x == n - 1) {
                composite = false;
                break;
            }
        }
        if (composite) return false;
    }
    return true;
}
inline size_t host_brev(size_t x, unsigned bits) {
    size_t r = 0;
    for (unsigned i = 0; i < bits; ++i)
        if (x >> i & 1) r |= (size_t)1 << (bits - 1 - i);
    return r;
}
// returns false if q is not prime or lacks the 2-adicity for a table of `len` entries (len = max ring degree)
inline bool host_build_twiddles(uint64_t q, size_t len, std::vector<uint64_t>& fwd, std::vector<uint64_t>& inv) {
    if (q <= 2 || !host_is_prime_u64(q)) return false;
    uint64_t order = q - 1;
    unsigned s = (unsigned)__builtin_ctzll(order);
    unsigned lg = 0;
    while (((size_t)1 << lg) < len) ++lg;
    if (lg + 1 > s) return false;
    uint64_t g = 0;
    for (uint64_t c = 1; c < order; ++c)
        if (host_powmod(c, order >> 1, q) == order) {
            g = c;
            break;
        }
    if (!g) return false;
    uint64_t psi = host_powmod(g, order >> s, q);
    for (unsigned i = 0; i < s - 1 - lg; ++i) psi = host_mulmod(psi, psi, q);
    uint64_t psi_inv = host_powmod(psi, q - 2, q);
    std::vector<uint64_t> pw(len), pwi(len);
    pw[0] = pwi[0] = 1;
    for (size_t i = 1; i < len; ++i) {
        pw[i] = host_mulmod(pw[i - 1], psi, q);
        pwi[i] = host_mulmod(pwi[i - 1], psi_inv, q);
    }
    fwd.resize(len);
    inv.resize(len);
    for (size_t j = 0; j < len; ++j) {
        size_t r = host_brev(j, lg);
        fwd[j] = pw[r];
        inv[j] = pwi[r];
    }
    return true;
}

}  // namespace fhe
