// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
//
// CPU restatement (C++17, no dependencies) of the `util` crate of han0110/learn-fhe, following the
// reference's *structure*: every modular add/mul is an `unsigned __int128 %`, transforms are radix-2
// in place, a coefficient-form product is three transforms.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may load this library.
//
// Parity status: the reference (Rust) cannot be built in this image and holds no golden vectors /
// known-answer tests (all its tests use entropy-seeded RNGs).  *Functional* parity is pinned by
// re-stating the reference's own property tests against this oracle (tests/test_oracle_*.py);
// *bitwise* parity is pinned only via this restatement plus an independent pure-Python restatement
// (tests/golden/gen_golden.py) — i.e. "bitwise parity unpinned by the reference itself".
//
// Every function cites the reference file:line (relative to /root/reference) it follows.
#pragma once
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <mutex>
#include <stdexcept>
#include <vector>

namespace orc {

typedef unsigned __int128 u128;
typedef uint64_t u64;
typedef int64_t i64;
typedef std::vector<u64> Vec;

// ------------------------------------------------------------------------------------------
// Zq  (util/src/zq.rs)
// ------------------------------------------------------------------------------------------

// zq.rs:44-47  from_u128
static inline u64 zq_from_u128(u64 q, u128 v) { return (u64)(v % (u128)q); }
// zq.rs:49-52  from_u64
static inline u64 zq_from_u64(u64 q, u64 v) { return v % q; }
// zq.rs:54-57  from_i64 (rem_euclid)
static inline u64 zq_from_i64(u64 q, i64 v) {
    i64 r = v % (i64)q;
    if (r < 0) r += (i64)q;
    return (u64)r;
}
// zq.rs:59-61  from_f64: f64::round is half-away-from-zero == std::round
static inline u64 zq_from_f64(u64 q, double v) { return zq_from_i64(q, (i64)std::round(v)); }
// zq.rs:71-77  to_i64 (centred; note strict `<` against q>>1)
static inline i64 zq_to_i64(u64 q, u64 v) { return v < (q >> 1) ? (i64)v : (i64)v - (i64)q; }
// zq.rs:83-89  to_center_u64
static inline u64 zq_to_center_u64(u64 q, u64 v) { return v < (q >> 1) ? v : (~(q - v)) + 1; }
// zq.rs:156-162 neg
static inline u64 zq_neg(u64 q, u64 v) { return zq_from_u64(q, q - v); }
// zq.rs:174-180 add
static inline u64 zq_add(u64 q, u64 a, u64 b) { return zq_from_u128(q, (u128)a + (u128)b); }
// zq.rs:182-188 sub = add(neg)
static inline u64 zq_sub(u64 q, u64 a, u64 b) { return zq_add(q, a, zq_neg(q, b)); }
// zq.rs:190-196 mul
static inline u64 zq_mul(u64 q, u64 a, u64 b) { return zq_from_u128(q, (u128)a * (u128)b); }
// zq.rs:111-117 pow (BigUint::modpow in the reference; any square-and-multiply gives the same value)
static inline u64 zq_pow(u64 q, u64 v, u64 e) {
    u64 r = 1 % q, b = v % q;
    while (e) {
        if (e & 1) r = zq_mul(q, r, b);
        b = zq_mul(q, b, b);
        e >>= 1;
    }
    return r;
}
// zq.rs:123-126 inv (extended gcd, canonical representative)
static inline u64 zq_inv(u64 q, u64 v) {
    if (v == 0) throw std::runtime_error("zq_inv(0)");
    i64 a = (i64)v, b = (i64)q, x0 = 1, x1 = 0;
    while (b != 0) {
        i64 t = a / b;
        i64 r = a - t * b;
        a = b;
        b = r;
        i64 x = x0 - t * x1;
        x0 = x1;
        x1 = x;
    }
    return zq_from_i64(q, x0);
}
// zq.rs:99-105 generator: smallest g with g^((q-1)/2) == q-1
static inline u64 zq_generator(u64 q) {
    u64 order = q - 1;
    for (u64 g = 1; g < order; ++g)
        if (zq_pow(q, g, order >> 1) == order) return g;
    throw std::runtime_error("no generator");
}
// zq.rs:107-109 two_adic_generator
static inline u64 zq_two_adic_generator(u64 q, unsigned log_n) { return zq_pow(q, zq_generator(q), (q - 1) >> log_n); }
// zq.rs:128-130 mod_switch — (v as f64 * q' as f64) / q as f64, then round
static inline u64 zq_mod_switch(u64 q, u64 v, u64 qp) {
    volatile double num = (double)v * (double)qp;  // volatile: forbid contraction / reassociation
    return zq_from_f64(qp, num / (double)q);
}
// zq.rs:132-140 mod_switch_odd
static inline u64 zq_mod_switch_odd(u64 q, u64 v, u64 qp) {
    volatile double num = (double)v * (double)qp;
    double x = num / (double)q;
    double u = std::floor(x);
    if (u == 0.0) return zq_from_u64(qp, (u64)std::round(x));
    return zq_from_u64(qp, ((u64)u) | 1);
}

// zq.rs:337-342 is_prime — the reference uses probably_prime(.,20); deterministic Miller-Rabin
// over the first 12 prime bases decides primality exactly for all u64, so both agree on primes.
static inline bool is_prime(u64 n) {
    if (n < 2) return false;
    static const u64 small[] = {2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37};
    for (u64 p : small) {
        if (n % p == 0) return n == p;
    }
    u64 d = n - 1;
    unsigned r = 0;
    while ((d & 1) == 0) {
        d >>= 1;
        ++r;
    }
    for (u64 a : small) {
        u64 x = zq_pow(n, a, d);
        if (x == 1 || x == n - 1) continue;
        bool comp = true;
        for (unsigned i = 1; i < r; ++i) {
            x = zq_mul(n, x, x);
            if (x == n - 1) {
                comp = false;
                break;
            }
        }
        if (comp) return false;
    }
    return true;
}
// zq.rs:325-329 two_adic_primes(bits, log_n): descending k in [2^(bits-log_n-1), 2^(bits-log_n)), q = k*2^log_n + 1
static inline Vec two_adic_primes(unsigned bits, unsigned log_n, size_t count) {
    assert(bits > log_n);
    u64 mn = 1ull << (bits - log_n - 1), mx = 1ull << (bits - log_n);
    Vec out;
    for (u64 k = mx; k-- > mn && out.size() < count;) {
        u64 q = (k << log_n) + 1;
        if (is_prime(q)) out.push_back(q);
    }
    return out;
}

// ------------------------------------------------------------------------------------------
// T64  (util/src/torus.rs) — wrapping u64; torus.rs:49-86
// ------------------------------------------------------------------------------------------
static inline u64 t64_from_f64(double v) { return (u64)(i64)std::round(v); }  // torus.rs:28-33
static inline double t64_to_f64(u64 v) { return (double)(i64)v; }             // torus.rs:24-26

// ------------------------------------------------------------------------------------------
// misc.rs:29-42 bit_reverse (identity for n <= 2)
// ------------------------------------------------------------------------------------------
template <typename T>
static inline void bit_reverse(std::vector<T>& v) {
    size_t n = v.size();
    if (n > 2) {
        assert((n & (n - 1)) == 0);
        unsigned lg = 0;
        while ((1ull << lg) < n) ++lg;
        for (size_t i = 0; i < n; ++i) {
            size_t j = 0;
            for (unsigned b = 0; b < lg; ++b)
                if (i >> b & 1) j |= (size_t)1 << (lg - 1 - b);
            if (i < j) std::swap(v[i], v[j]);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Negacyclic NTT  (util/src/ring/fft.rs:37-115, util/src/ring/fft/zq.rs:14-67)
// ------------------------------------------------------------------------------------------
struct Twiddle {
    Vec fwd, inv;  // bit-reversed order
};
// fft/zq.rs:58-67 compute_twiddle
static inline Twiddle compute_twiddle(u64 q) {
    u64 order = q - 1;
    unsigned s = __builtin_ctzll(order);
    u64 root = zq_two_adic_generator(q, s);
    size_t len = (size_t)1 << (s - 1);
    Twiddle t;
    t.fwd.resize(len);
    t.inv.resize(len);
    u64 p = 1 % q;
    for (size_t i = 0; i < len; ++i) {  // zq.rs:119-121 powers()
        t.fwd[i] = p;
        p = zq_mul(q, p, root);
    }
    for (size_t i = 0; i < len; ++i) t.inv[i] = zq_inv(q, t.fwd[i]);
    bit_reverse(t.fwd);
    bit_reverse(t.inv);
    return t;
}
// fft/zq.rs:38-56 twiddle(): global cache behind a mutex
static inline const Twiddle& twiddle(u64 q) {
    static std::mutex mu;
    static std::map<u64, Twiddle>* cache = new std::map<u64, Twiddle>();
    std::lock_guard<std::mutex> g(mu);
    auto it = cache->find(q);
    if (it == cache->end()) {
        if (!is_prime(q)) throw std::runtime_error("twiddle: q not prime");
        it = cache->emplace(q, compute_twiddle(q)).first;
    }
    return it->second;
}

// fft.rs:94-101 Butterfly::dit
static inline void bf_dit(u64 q, u64& a, u64& b, u64 t) {
    u64 tb = zq_mul(q, t, b);
    u64 c = zq_add(q, a, tb);
    u64 d = zq_sub(q, a, tb);
    a = c;
    b = d;
}
// fft.rs:103-109 Butterfly::dif
static inline void bf_dif(u64 q, u64& a, u64& b, u64 t) {
    u64 c = zq_add(q, a, b);
    u64 d = zq_mul(q, zq_sub(q, a, b), t);
    a = c;
    b = d;
}
// fft.rs:40-54 nega_cyclic_fft_in_place (Alg. 1 of ePrint 2016/504); the `m == 0` branch is dead.
static inline void nega_cyclic_fft_in_place(u64 q, u64* a, size_t n, const Vec& tw) {
    assert((n & (n - 1)) == 0);
    unsigned log_n = 0;
    while (((size_t)1 << log_n) < n) ++log_n;
    for (unsigned layer = 0; layer < log_n; ++layer) {
        size_t m = (size_t)1 << layer, size = (size_t)1 << (log_n - layer - 1);
        for (size_t i = 0; i < m; ++i) {
            u64 t = tw[m + i];
            u64* u = a + 2 * size * i;
            u64* v = u + size;
            for (size_t j = 0; j < size; ++j) bf_dit(q, u[j], v[j], t);
        }
    }
}
// fft.rs:59-77 nega_cyclic_ifft_in_place (Alg. 2), final multiply by n^-1
static inline void nega_cyclic_ifft_in_place(u64 q, u64* a, size_t n, const Vec& tw_inv, u64 n_inv) {
    unsigned log_n = 0;
    while (((size_t)1 << log_n) < n) ++log_n;
    for (unsigned layer = log_n; layer-- > 0;) {
        size_t m = (size_t)1 << layer, size = (size_t)1 << (log_n - layer - 1);
        for (size_t i = 0; i < m; ++i) {
            u64 t = tw_inv[m + i];
            u64* u = a + 2 * size * i;
            u64* v = u + size;
            for (size_t j = 0; j < size; ++j) bf_dif(q, u[j], v[j], t);
        }
    }
    for (size_t i = 0; i < n; ++i) a[i] = zq_mul(q, a[i], n_inv);
}
// fft/zq.rs:27-30
static inline void nega_cyclic_ntt_in_place(u64 q, u64* a, size_t n) {
    const Twiddle& t = twiddle(q);
    if (n > 1 && t.fwd.size() < n) throw std::runtime_error("ntt: 2-adicity too small for n");
    nega_cyclic_fft_in_place(q, a, n, t.fwd);
}
// fft/zq.rs:32-36
static inline void nega_cyclic_intt_in_place(u64 q, u64* a, size_t n) {
    const Twiddle& t = twiddle(q);
    if (n > 1 && t.inv.size() < n) throw std::runtime_error("intt: 2-adicity too small for n");
    u64 n_inv = zq_inv(q, zq_from_u64(q, (u64)n));
    nega_cyclic_ifft_in_place(q, a, n, t.inv, n_inv);
}
// fft/zq.rs:14-25 nega_cyclic_ntt_mul_assign: NTT(a), NTT(clone b), pointwise, iNTT
static inline void nega_cyclic_ntt_mul_assign(u64 q, u64* a, const u64* b, size_t n) {
    nega_cyclic_ntt_in_place(q, a, n);
    Vec bb(b, b + n);
    nega_cyclic_ntt_in_place(q, bb.data(), n);
    for (size_t i = 0; i < n; ++i) a[i] = zq_mul(q, a[i], bb[i]);
    nega_cyclic_intt_in_place(q, a, n);
}
// ring.rs:421-440 nega_cyclic_schoolbook_mul (test helper in the reference; independent check here)
static inline Vec schoolbook_zq(u64 q, const u64* a, const u64* b, size_t n) {
    Vec c(n);
    for (size_t i = 0; i < n; ++i) c[i] = zq_mul(q, a[i], b[0]);
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 1; j < n; ++j) {
            u64 p = zq_mul(q, a[i], b[j]);
            if (i + j < n)
                c[i + j] = zq_add(q, c[i + j], p);
            else
                c[i + j - n] = zq_sub(q, c[i + j - n], p);
        }
    return c;
}
static inline Vec schoolbook_t64(const u64* a, const u64* b, size_t n) {
    Vec c(n);
    for (size_t i = 0; i < n; ++i) c[i] = a[i] * b[0];
    for (size_t i = 0; i < n; ++i)
        for (size_t j = 1; j < n; ++j) {
            u64 p = a[i] * b[j];
            if (i + j < n)
                c[i + j] += p;
            else
                c[i + j - n] -= p;
        }
    return c;
}

// ------------------------------------------------------------------------------------------
// avec.rs:34-50 automorphism; ring.rs:299-313 monomial multiply
// ------------------------------------------------------------------------------------------
template <typename Neg>
static inline Vec automorphism_generic(const u64* in, size_t n, i64 t, Neg neg) {
    Vec v(in, in + n);
    i64 m = 2 * (i64)n;
    size_t tt = (size_t)(((t % m) + m) % m);
    for (size_t i = 0; i < n; ++i) {
        size_t it = (i * tt) % (2 * n);
        if (it < n)
            v[it] = in[i];
        else
            v[it - n] = neg(in[i]);
    }
    return v;
}
static inline Vec automorphism_zq(u64 q, const u64* in, size_t n, i64 t) {
    return automorphism_generic(in, n, t, [q](u64 x) { return zq_neg(q, x); });
}
static inline Vec automorphism_t64(const u64* in, size_t n, i64 t) {
    return automorphism_generic(in, n, t, [](u64 x) { return (u64)(0 - x); });
}
template <typename Neg>
static inline void monomial_mul_generic(u64* a, size_t n, i64 k, Neg neg) {
    i64 m = 2 * (i64)n;
    size_t i = (size_t)(((k % m) + m) % m);
    size_t r = i % n;
    // slice::rotate_right(r)
    Vec tmp(a, a + n);
    for (size_t j = 0; j < n; ++j) a[(j + r) % n] = tmp[j];
    if (i < n) {
        for (size_t j = 0; j < i; ++j) a[j] = neg(a[j]);
    } else {
        for (size_t j = i - n; j < n; ++j) a[j] = neg(a[j]);
    }
}
static inline void monomial_mul_zq(u64 q, u64* a, size_t n, i64 k) {
    monomial_mul_generic(a, n, k, [q](u64 x) { return zq_neg(q, x); });
}
static inline void monomial_mul_t64(u64* a, size_t n, i64 k) {
    monomial_mul_generic(a, n, k, [](u64 x) { return (u64)(0 - x); });
}

// ------------------------------------------------------------------------------------------
// misc/decompose.rs
// ------------------------------------------------------------------------------------------
struct DecomposorZq {  // decompose.rs:49-64
    u64 q;
    unsigned log_q, log_b, d, rounding_bits;
    DecomposorZq() : q(0), log_q(0), log_b(0), d(0), rounding_bits(0) {}
    DecomposorZq(u64 q_, unsigned log_b_, unsigned d_) : q(q_), log_b(log_b_), d(d_) {
        // q.next_power_of_two().ilog2()
        u64 p = 1;
        log_q = 0;
        while (p < q) {
            p <<= 1;
            ++log_q;
        }
        rounding_bits = log_q > log_b * d ? log_q - log_b * d : 0;
    }
    // decompose.rs:25-27 log_bases; :54-55 bases[i] = 2^(rounding_bits + i*log_b) mod q
    u64 base(unsigned i) const { return zq_from_u64(q, 1ull << (rounding_bits + i * log_b)); }
    // decompose.rs:92-95 rounding_shr
    u64 rounding_shr(u64 v) const {
        u64 rounded = zq_add(q, v, zq_from_u64(q, (1ull << rounding_bits) >> 1));
        return zq_from_u64(q, rounded >> rounding_bits);
    }
    // decompose.rs:42-46 + 101-111: digits of one element, least-significant first
    void decompose(u64 v, u64* out) const {
        u64 b_by_2 = 1ull << (log_b - 1), mask = (1ull << log_b) - 1, neg_b = q - (1ull << log_b);
        u64 x = zq_to_center_u64(q, rounding_shr(v));
        for (unsigned k = 0; k < d; ++k) {
            u64 limb = x & mask;
            u64 carry = (limb + (x & 1) > b_by_2) ? 1 : 0;
            x >>= log_b;
            x += carry;
            out[k] = zq_from_u64(q, limb + carry * neg_b);
        }
    }
    // decompose.rs:137-155 collection impl: limb-major (digit k of every element, then digit k+1 ...)
    void decompose_vec(const u64* in, size_t n, u64* out /* d*n */) const {
        Vec tmp(d);
        for (size_t i = 0; i < n; ++i) {
            decompose(in[i], tmp.data());
            for (unsigned k = 0; k < d; ++k) out[(size_t)k * n + i] = tmp[k];
        }
    }
};
struct DecomposorT64 {  // decompose.rs:66-81
    unsigned log_b, d, rounding_bits;
    DecomposorT64() : log_b(0), d(0), rounding_bits(0) {}
    DecomposorT64(unsigned log_b_, unsigned d_) : log_b(log_b_), d(d_) {
        rounding_bits = 64 > log_b * d ? 64 - log_b * d : 0;
    }
    u64 base(unsigned i) const { return 1ull << (rounding_bits + i * log_b); }
    // decompose.rs:115-118
    static u64 rounding_shr_bits(u64 v, unsigned bits) {
        u64 rounded = v + ((bits >= 64 ? 0 : (1ull << bits)) >> 1);
        return bits >= 64 ? 0 : rounded >> bits;
    }
    u64 rounding_shr(u64 v) const { return rounding_shr_bits(v, rounding_bits); }
    // decompose.rs:124-134
    void decompose(u64 v0, u64* out) const {
        u64 mask = (1ull << log_b) - 1;
        u64 v = rounding_shr(v0);
        for (unsigned k = 0; k < d; ++k) {
            u64 limb = v & mask;
            v >>= log_b;
            u64 carry = ((limb - 1) | v) & limb;
            carry >>= (log_b - 1);
            v += carry;
            out[k] = limb - (carry << log_b);
        }
    }
    void decompose_vec(const u64* in, size_t n, u64* out) const {
        Vec tmp(d);
        for (size_t i = 0; i < n; ++i) {
            decompose(in[i], tmp.data());
            for (unsigned k = 0; k < d; ++k) out[(size_t)k * n + i] = tmp[k];
        }
    }
};

// ------------------------------------------------------------------------------------------
// f64 FFT negacyclic product over T64  (util/src/ring/fft/c64.rs, util/src/ring/fft.rs:7-35)
// ------------------------------------------------------------------------------------------
struct C64 {
    double re, im;
};
// num_complex Mul: (a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re); no FMA contraction (Rust never contracts)
static inline C64 c_mul(C64 a, C64 b) {
    volatile double p0 = a.re * b.re, p1 = a.im * b.im, p2 = a.re * b.im, p3 = a.im * b.re;
    return C64{p0 - p1, p2 + p3};
}
static inline C64 c_add(C64 a, C64 b) { return C64{a.re + b.re, a.im + b.im}; }
static inline C64 c_sub(C64 a, C64 b) { return C64{a.re - b.re, a.im - b.im}; }
struct Twiddle64 {
    std::vector<C64> tw, tw_inv, tw_bo, tw_inv_bo;  // c64.rs:98-108
};
// c64.rs:98-108 compute_twiddle(n): cis(i*pi/n)
static inline Twiddle64 compute_twiddle64(size_t n) {
    Twiddle64 t;
    t.tw.resize(n);
    t.tw_inv.resize(n);
    for (size_t i = 0; i < n; ++i) {
        volatile double num = (double)i * M_PI;
        double ang = num / (double)n;
        t.tw[i] = C64{std::cos(ang), std::sin(ang)};
        t.tw_inv[i] = C64{t.tw[i].re, -t.tw[i].im};
    }
    t.tw_bo = t.tw;
    t.tw_inv_bo = t.tw_inv;
    bit_reverse(t.tw_bo);
    bit_reverse(t.tw_inv_bo);
    return t;
}
// c64.rs:88-95: grow-only global table.  NOTE (reference behaviour): the table is recomputed only when
// a larger n is requested, and smaller n index it with a stride (c64.rs:24,36) or use a prefix of the
// bit-reversed table (c64.rs:59,64), so the values used for a given n depend on the largest n seen so far
// only through libm rounding of i*pi/n' with n' >= n.  We key the table by the exact n requested so the
// oracle is history-independent; for power-of-two n'/n the angles i*pi/n are bit-identical either way
// because (i*k)*pi/(n*k) and i*pi/n round identically when k is a power of two.
static inline const Twiddle64& twiddle64(size_t n) {
    static std::mutex mu;
    static std::map<size_t, Twiddle64>* cache = new std::map<size_t, Twiddle64>();
    std::lock_guard<std::mutex> g(mu);
    auto it = cache->find(n);
    if (it == cache->end()) it = cache->emplace(n, compute_twiddle64(n)).first;
    return it->second;
}
// fft.rs:9-19 fft_in_place (cyclic, DIF-shaped loop order with Butterfly::dit)
static inline void fft_in_place_c64(C64* a, size_t n, const std::vector<C64>& tw_bo) {
    unsigned lg = 0;
    while (((size_t)1 << lg) < n) ++lg;
    for (unsigned layer = lg; layer-- > 0;) {
        size_t size = (size_t)1 << layer;
        size_t chunks = n / (2 * size);
        for (size_t c = 0; c < chunks; ++c) {
            C64 t = tw_bo[c];
            C64* u = a + 2 * size * c;
            C64* v = u + size;
            for (size_t j = 0; j < size; ++j) {
                C64 tb = c_mul(t, v[j]);  // dit: tb = t*b; a+tb; a-tb
                C64 x = c_add(u[j], tb), y = c_sub(u[j], tb);
                u[j] = x;
                v[j] = y;
            }
        }
    }
}
// fft.rs:23-35 ifft_in_place
static inline void ifft_in_place_c64(C64* a, size_t n, const std::vector<C64>& tw_inv_bo, double n_inv) {
    unsigned lg = 0;
    while (((size_t)1 << lg) < n) ++lg;
    for (unsigned layer = 0; layer < lg; ++layer) {
        size_t size = (size_t)1 << layer;
        size_t chunks = n / (2 * size);
        for (size_t c = 0; c < chunks; ++c) {
            C64 t = tw_inv_bo[c];
            C64* u = a + 2 * size * c;
            C64* v = u + size;
            for (size_t j = 0; j < size; ++j) {
                C64 x = c_add(u[j], v[j]);           // dif: c = a+b; d = (a-b)*t
                C64 y = c_mul(c_sub(u[j], v[j]), t);
                u[j] = x;
                v[j] = y;
            }
        }
    }
    for (size_t i = 0; i < n; ++i) {  // C64 *= &f64  (num_complex: re*=s, im*=s)
        a[i].re *= n_inv;
        a[i].im *= n_inv;
    }
}
// c64.rs:69-85 f64_mod_u64
static inline u64 f64_mod_u64(double v) {
    u64 bits;
    std::memcpy(&bits, &v, 8);
    u64 sign = bits >> 63;
    u64 exponent = (bits >> 52) & 0x7ff;
    u64 mantissa = (bits << 11) | 0x8000000000000000ull;
    i64 shift = 1086 - (i64)exponent;
    u64 value;
    if (shift >= -63 && shift <= 0)
        value = mantissa << (-shift);
    else if (shift >= 1 && shift <= 64)
        value = (shift - 1 >= 64 ? 0 : ((mantissa >> (shift - 1)) + 1)) >> 1;
    else
        value = 0;
    return sign == 0 ? value : (u64)(0 - value);
}
// c64.rs:20-28 to_c64_twisted — twiddle(n)[0] strided by len/n; we request the table for exactly n.
static inline std::vector<C64> to_c64_twisted(const u64* a, size_t n) {
    // the reference indexes twiddle(a.len()) — table built for n (entries cis(i*pi/n)) — with step len/n = 1
    const Twiddle64& t = twiddle64(n);
    std::vector<C64> c(n / 2);
    for (size_t j = 0; j < n / 2; ++j) c[j] = c_mul(C64{t64_to_f64(a[j]), t64_to_f64(a[j + n / 2])}, t.tw[j]);
    return c;
}
// c64.rs:31-41 assign_from_c64_twisted
static inline void assign_from_c64_twisted(u64* a, size_t n, const std::vector<C64>& c) {
    const Twiddle64& t = twiddle64(n);
    for (size_t j = 0; j < n / 2; ++j) {
        C64 x = c_mul(c[j], t.tw_inv[j]);
        a[j] = f64_mod_u64(x.re);
        a[j + n / 2] = f64_mod_u64(x.im);
    }
}
// c64.rs:58-67: nega_cyclic_fft64_in_place(a) looks up twiddle(a.len()) where a.len() = n/2 complex
// points and uses the bit-reversed table [2]; a prefix of the bit-reversed table of a larger table
// equals the bit-reversed table cis(brev(k)*pi/len) restricted to the first len/2 entries.
static inline void nega_cyclic_fft64_in_place(C64* a, size_t len) { fft_in_place_c64(a, len, twiddle64(len).tw_bo); }
static inline void nega_cyclic_ifft64_in_place(C64* a, size_t len) {
    ifft_in_place_c64(a, len, twiddle64(len).tw_inv_bo, 1.0 / (double)len);
}
// c64.rs:11-17, 43-56
static inline void nega_cyclic_fft64_mul_assign_rt(u64* a, const u64* b, size_t n) {
    if (n == 1) {
        a[0] *= b[0];
        return;
    }
    std::vector<C64> ca = to_c64_twisted(a, n), cb = to_c64_twisted(b, n);
    nega_cyclic_fft64_in_place(ca.data(), n / 2);
    nega_cyclic_fft64_in_place(cb.data(), n / 2);
    for (size_t i = 0; i < n / 2; ++i) ca[i] = c_mul(ca[i], cb[i]);
    nega_cyclic_ifft64_in_place(ca.data(), n / 2);
    assign_from_c64_twisted(a, n, ca);
}

}  // namespace orc
