// Deterministic key-generation streams (key generation on the device).
// The reference draws key material from an `RngCore` passed in by the caller (scheme/fhew/src/bootstrapping.rs:122-146,
// scheme/tfhe/src/bootstrapping.rs:59-76, scheme/ckks/src/ckks.rs:154-184); which generator is the caller's business.  A GPU
// cannot replay a sequential generator, so the device key generation defines its own COUNTER-BASED stream: value = f(seed,
// domain, index), every word independent of the order in which threads produce it.  The same integer-only functions run on the
// host (oracle/orc_keygen.hpp, the checker) and on the device (csrc/keygen.cu), so a device-generated key can be compared
// word for word with a host-generated one.
//   ks_u64      splitmix64 finaliser of (seed, domain, index)
//   ks_uniform  floor(u64 * q / 2^64): uniform over [0, q) up to a bias below q / 2^64
//   ks_gauss    discrete Gaussian, sigma = 3.2, support [-19, 19] (distribution.rs:23-47 `dg(3.2, 6)`: 6 sigma tail cut) by
//               inversion on a 38-entry cumulative table with 64-bit resolution (no floating point: identical everywhere)
//   ks_ternary  zo(0.5) (distribution.rs:10-21): -1, +1 with probability 1/4 each, else 0
#pragma once
#include <cstdint>

#include "modarith.cuh"

namespace fhe {

HD uint64_t ks_u64(uint64_t seed, uint32_t domain, uint64_t index) {
    uint64_t z = seed + 0x9E3779B97F4A7C15ull * (index + 1) + ((uint64_t)domain << 56) * 0xD6E8FEB86659FD93ull;
    z ^= (uint64_t)domain * 0xA0761D6478BD642Full;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
HD uint64_t ks_uniform(uint64_t seed, uint32_t domain, uint64_t index, uint64_t q) { return mulhi_u64(ks_u64(seed, domain, index), q); }
HD int64_t ks_gauss(uint64_t seed, uint32_t domain, uint64_t index) {
    // cumulative distribution of exp(-x^2 / (2 * 3.2^2)), x = -19 .. 18, scaled to 2^64 (tools: mpmath, 200 bits)
    const uint64_t cdt[38] = {
    0x0000000bd798c7acull, 0x00000053f5e234f7ull, 0x000001e24b89625aull, 0x000009adbc7900e1ull,
    0x00002d17ca1e3a82ull, 0x0000bf05ff772ee2ull, 0x0002e06843bbf95dull, 0x000a19062dceb033ull,
    0x00204c0ded173b17ull, 0x005e3163ffc656a8ull, 0x00fab6a7e614dfaeull, 0x0261b16e99bc14f8ull,
    0x054c69348bbc9867ull, 0x0acd276539458231ull, 0x143796b3c337f00bull, 0x22d43df09de40e7aull,
    0x37651d96aa4fb429ull, 0x51a5da2b40f1bff1ull, 0x700ad4bf5a6e80e2ull, 0x8ff52b40a5917f1dull,
    0xae5a25d4bf0e400eull, 0xc89ae26955b04bd6ull, 0xdd2bc20f621bf185ull, 0xebc8694c3cc80ff4ull,
    0xf532d89ac6ba7dceull, 0xfab396cb74436798ull, 0xfd9e4e916643eb07ull, 0xff05495819eb2051ull,
    0xffa1ce9c0039a957ull, 0xffdfb3f212e8c4e8ull, 0xfff5e6f9d2314fccull, 0xfffd1f97bc4406a2ull,
    0xffff40fa0088d11dull, 0xffffd2e835e1c57dull, 0xfffff6524386ff1eull, 0xfffffe1db4769da5ull,
    0xffffffac0a1dcb08ull, 0xfffffff428673853ull,
    };
    const uint64_t u = ks_u64(seed, domain, index);
    int lo = 0, hi = 38;  // number of thresholds <= u
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdt[mid] <= u)
            lo = mid + 1;
        else
            hi = mid;
    }
    return (int64_t)lo - 19;
}
HD int64_t ks_ternary(uint64_t seed, uint32_t domain, uint64_t index) {
    const uint64_t u = ks_u64(seed, domain, index) >> 62;
    return u == 0 ? -1 : (u == 1 ? 1 : 0);
}
// Torus noise (distribution.rs:49-54 `tdg(sigma)`: the fractional part of a normal variate scaled by 2^64) without floating point:
// an Irwin-Hall variate - the sum of twelve 32-bit uniforms, mean 6 * 2^32, standard deviation exactly 2^32 - scaled by
// sigma_q = round(sigma * 2^64) / 2^32.  Approximately Gaussian (support +-6 sigma, like the cut of `dg`), integer-only and
// therefore identical on host and device; needs sigma < 2^-8 so that the 128-bit product fits the shifts below.
HD uint64_t ks_tgauss(uint64_t seed, uint32_t domain, uint64_t index, uint64_t sigma_q) {
    uint64_t sum = 0;
    for (int j = 0; j < 6; ++j) {
        const uint64_t w = ks_u64(seed, domain, index * 8 + j);
        sum += (w >> 32) + (w & 0xFFFFFFFFull);
    }
    const int64_t sc = (int64_t)sum - (6ll << 32);
    const uint64_t mag = (uint64_t)(sc < 0 ? -sc : sc);
    const uint64_t hi = mulhi_u64(mag, sigma_q), lo = mag * sigma_q;
    const uint64_t e = (hi << 32) | (lo >> 32);
    return sc < 0 ? (uint64_t)(0 - e) : e;
}
// binary secret (distribution.rs:6-8): 0 / 1 with probability 1/2
HD int64_t ks_binary(uint64_t seed, uint32_t domain, uint64_t index) { return (int64_t)(ks_u64(seed, domain, index) >> 63); }

// stream domains of the CKKS key generation: secret, then (mask, error) per key: key 0 = relinearisation key, 1.. = automorphism keys
enum : uint32_t { KS_CKKS_SK = 16, KS_CKKS_KEY0 = 17 };
HD uint32_t ks_ckks_a(uint32_t key) { return KS_CKKS_KEY0 + 2 * key; }
HD uint32_t ks_ckks_e(uint32_t key) { return KS_CKKS_KEY0 + 2 * key + 1; }
// stream domains of the TFHE key generation
enum : uint32_t { KS_TFHE_Z = 40, KS_TFHE_S = 41, KS_TFHE_BRK_A = 42, KS_TFHE_BRK_E = 43, KS_TFHE_KSK_A = 44, KS_TFHE_KSK_E = 45 };
// stream domains of the FHEW key generation
enum : uint32_t { KS_FHEW_Z = 1, KS_FHEW_S = 2, KS_FHEW_KSK_A = 3, KS_FHEW_KSK_E = 4, KS_FHEW_BRK_A = 5, KS_FHEW_BRK_E = 6, KS_FHEW_AK_A = 7, KS_FHEW_AK_E = 8 };

}  // namespace fhe
