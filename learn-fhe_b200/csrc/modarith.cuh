// Modular arithmetic primitives for the sm_100a kernels (u32 path: q <= 2^30, u64 path: q <= 2^62).
// Replaces the reference's `u128 %` Zq arithmetic (util/src/zq.rs:156-196) with Shoup / Barrett forms
// whose *canonical* results are identical.  Everything is __host__ __device__ so that the index and
// arithmetic logic can be exercised on the CPU by tests/hostsim (test infrastructure; the product
// path only ever runs the device compilation).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define HD __host__ __device__ __forceinline__
#define DEV __device__ __forceinline__
#else
#define HD inline
#define DEV inline
#endif

#if !defined(__CUDACC__)
// host-only compilation (tests/hostsim): minimal stand-ins for the CUDA vector types used in shared headers
struct uint2 {
    unsigned int x, y;
};
struct double2 {
    double x, y;
};
struct uint4 {
    unsigned int x, y, z, w;
};
#endif

namespace fhe {

typedef unsigned __int128 u128_t;

HD uint32_t mulhi_u32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
HD uint64_t mulhi_u64(uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
    return __umul64hi(a, b);
#else
    return (uint64_t)(((u128_t)a * b) >> 64);
#endif
}
// hi64(a*b) from three 32x32 partial products; result in {Q-2, Q-1, Q} for Q = floor(a*b / 2^64)
HD uint64_t mulhi_u64_approx(uint64_t a, uint64_t b) {
    uint32_t al = (uint32_t)a, ah = (uint32_t)(a >> 32), bl = (uint32_t)b, bh = (uint32_t)(b >> 32);
    // only the high words of the cross terms are needed: IMAD.HI (2 issue slots) instead of IMAD.WIDE (2.4)
    return (uint64_t)ah * bh + (uint64_t)mulhi_u32(al, bh) + (uint64_t)mulhi_u32(ah, bl);
}
// floor(wp * y / 2^32) on the FP64 pipe.  c = wp 2^-32 and K = 2^52 - wp 2^20 are exact doubles; with Y = 2^52 + y (y placed in the
// low word of the double 2^52) the product-sum Y c + K equals y wp 2^-32 + 2^52 exactly, and one DFMA rounded towards minus
// infinity lands on 2^52 + floor(y wp / 2^32): the low word of the result IS mulhi_u32(wp, y), for every y and wp.
struct F64Quot {
    double c, k;
};
HD F64Quot make_f64_quot(uint32_t wp) {
    F64Quot f;
#if defined(__CUDA_ARCH__)
    f.c = __dadd_rn(__hiloint2double(0x41300000, (int)wp), -0x1p20);  // (2^20 + wp 2^-32) - 2^20
    f.k = __fma_rn(f.c, -0x1p52, 0x1p52);
#else
    f.c = (double)wp * 0x1p-32;
    f.k = 0x1p52 - (double)wp * 0x1p20;
#endif
    return f;
}
HD uint32_t mulhi_u32_f64(const F64Quot& f, uint32_t y) {
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(__fma_rd(__hiloint2double(0x43300000, (int)y), f.c, f.k));
#else
    return (uint32_t)(((uint64_t)(uint32_t)(f.c * 0x1p32) * y) >> 32);
#endif
}
HD uint32_t umin_(uint32_t a, uint32_t b) { return a < b ? a : b; }
// a + b for sums that cannot wrap (a + b < 2^32), forced onto the ALU pipe (VIADDMNMX).  ptxas otherwise places many
// two-operand integer adds on the IMAD pipe as IMAD.IADD, and that pipe is the measured bottleneck of every modular kernel
// here (64 lanes/clk/SM on B200; ncu: sm__pipe_fmaheavy_cycles_active 58-66 %).
HD uint32_t alu_add(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    return __viaddmin_u32(a, b, 0xFFFFFFFFu);
#else
    return a + b;
#endif
}
HD uint64_t umin_(uint64_t a, uint64_t b) { return a < b ? a : b; }

// ------------------------------------------------------------------------------------------------
// 32-bit modulus.  Values live in a u32; lazy ranges [0,2q) / [0,4q) need q <= 2^30.
// ------------------------------------------------------------------------------------------------
struct Mod32 {
    typedef uint32_t W;
    typedef uint64_t W2;  // MAC accumulator
    uint32_t q, q2;
    uint64_t mu;  // floor(2^64 / q): Barrett constant for reducing 64-bit accumulators
    static constexpr int BITS = 32;
    HD static uint32_t mulhi(uint32_t a, uint32_t b) { return mulhi_u32(a, b); }
    // x in [0,4q) -> [0,2q)
    HD uint32_t red2q(uint32_t x) const { return umin_(x, x - q2); }
    // x in [0,2q) -> [0,q)
    HD uint32_t redq(uint32_t x) const { return umin_(x, x - q); }
    HD uint32_t canon4(uint32_t x) const { return redq(red2q(x)); }
    // w * y mod q in [0,2q) for ANY y < 2^32, given wp = floor(w * 2^32 / q)
    HD uint32_t shoup_lazy(uint32_t y, uint32_t w, uint32_t wp) const { return w * y - mulhi_u32(wp, y) * q; }
    HD uint32_t add(uint32_t a, uint32_t b) const { return redq(a + b); }          // canonical in, canonical out
    HD uint32_t sub(uint32_t a, uint32_t b) const { return redq(a + q - b); }
    HD uint32_t neg(uint32_t a) const { return a == 0 ? 0 : q - a; }
    // reduce a 64-bit value (any) to canonical
    HD uint32_t reduce64(uint64_t x) const {
        uint64_t qh = mulhi_u64(x, mu);          // floor(x/q) - {0,1}
        uint64_t r = x - qh * (uint64_t)q;       // [0, 2q)
        return redq((uint32_t)r);
    }
    HD uint32_t mul(uint32_t a, uint32_t b) const { return reduce64((uint64_t)a * b); }
    // MAC accumulator: acc += a*b with a,b < 2^32; caller bounds the number of terms so acc < 2^64
    HD static void mac(uint64_t& acc, uint32_t a, uint32_t b) { acc += (uint64_t)a * b; }
    HD uint32_t reduce_acc(uint64_t acc) const { return reduce64(acc); }
    HD static uint32_t shoup_companion_host(uint32_t w, uint32_t q_) { return (uint32_t)(((uint64_t)w << 32) / q_); }
};

// ------------------------------------------------------------------------------------------------
// 64-bit modulus, q <= 2^62 ("safe": exact mulhi, Harvey ranges [0,2q)/[0,4q)).
// ------------------------------------------------------------------------------------------------
struct U128 {
    uint64_t lo, hi;
};
HD U128 mul_wide_u64(uint64_t a, uint64_t b) {
    U128 r;
    r.lo = a * b;
    r.hi = mulhi_u64(a, b);
    return r;
}
struct Mod64 {
    typedef uint64_t W;
    typedef U128 W2;
    uint64_t q, q2;
    uint64_t mu;      // floor(2^(2s) / q), s = bit length used by the 128->64 Barrett
    uint32_t s;       // q < 2^s, s <= 62
    static constexpr int BITS = 64;
    HD uint64_t red2q(uint64_t x) const { return umin_(x, x - q2); }
    HD uint64_t redq(uint64_t x) const { return umin_(x, x - q); }
    HD uint64_t canon4(uint64_t x) const { return redq(red2q(x)); }
    HD uint64_t shoup_lazy(uint64_t y, uint64_t w, uint64_t wp) const { return w * y - mulhi_u64(wp, y) * q; }
    HD uint64_t add(uint64_t a, uint64_t b) const { return redq(a + b); }
    HD uint64_t sub(uint64_t a, uint64_t b) const { return redq(a + q - b); }
    HD uint64_t neg(uint64_t a) const { return a == 0 ? 0 : q - a; }
    // 128-bit value x < 2^(2s) -> canonical.  Barrett: x1 = x >> (s-1) (< 2^(s+1)); qh = (x1*mu) >> (s+1)
    HD uint64_t reduce128(U128 x) const {
        uint64_t x1 = (x.lo >> (s - 1)) | (x.hi << (65 - s));  // s >= 2
        U128 p = mul_wide_u64(x1, mu);
        uint64_t qh = (p.lo >> (s + 1)) | (p.hi << (63 - s));
        uint64_t r = x.lo - qh * q;  // [0, 3q)
        r = umin_(r, r - q2);
        return redq(r);
    }
    HD uint64_t mul(uint64_t a, uint64_t b) const { return reduce128(mul_wide_u64(a, b)); }
    // 128-bit MAC accumulator; every term a*b < q^2 <= 2^(2s) and terms are reduced on overflow risk by the caller
    HD static void mac(U128& acc, uint64_t a, uint64_t b) {
        U128 p = mul_wide_u64(a, b);
        uint64_t lo = acc.lo + p.lo;
        acc.hi += p.hi + (lo < acc.lo ? 1 : 0);
        acc.lo = lo;
    }
};

// host-side helpers used by the context when it builds tables (setup only)
inline uint64_t host_mulmod(uint64_t a, uint64_t b, uint64_t q) { return (uint64_t)(((u128_t)a * b) % q); }
inline uint64_t host_powmod(uint64_t b, uint64_t e, uint64_t q) {
    uint64_t r = 1 % q;
    b %= q;
    while (e) {
        if (e & 1) r = host_mulmod(r, b, q);
        b = host_mulmod(b, b, q);
        e >>= 1;
    }
    return r;
}
inline uint64_t host_shoup64(uint64_t w, uint64_t q) { return (uint64_t)((((u128_t)w) << 64) / q); }
inline uint32_t host_shoup32(uint32_t w, uint32_t q) { return (uint32_t)((((uint64_t)w) << 32) / q); }

}  // namespace fhe
