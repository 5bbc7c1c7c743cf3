// TFHE blind rotation, bounded-error ("fast") mode: one CMUX = 5 fused register passes around 4 shared-memory exchanges.
// __host__ __device__ so tests/hostsim can replay the kernel logic on the CPU.
//
// Contract (BASELINE.json north_star: "where the reference uses a floating-point FFT, coefficient error stays within a stated
// bound and decrypted results stay bit-identical"): this mode computes the same TGGSW external product as
// scheme/tfhe/src/tggsw.rs:100-112 with the same signed digits (util/src/misc/decompose.rs:114-135) and the same f64 ring
// product idea as util/src/ring/fft/c64.rs:11-56 (fold n reals into n/2 complex, twist, FFT, pointwise, inverse), but
//   * the (k+1)d products of one output are summed in the Fourier domain ((k+1) inverse transforms per CMUX, not (k+1)^2 d),
//   * every butterfly is the 6-FMA form (a + w b, 2a - (a + w b)), the twist is merged into the forward twiddles
//     (evaluation at the roots of X^(n/2) = i, like a negacyclic NTT), the inverse is a decimation-in-time transform with
//     position twiddles followed by the untwist, and
//   * f64 -> torus rounding is x - 2^64 rint(x 2^-64) (one rounding to nearest; ties to even instead of away from zero).
// Every output coefficient therefore carries ONE rounding of a sum whose exact value equals the reference's exact sum; the
// error against the exact negacyclic product obeys the reference's own per-product bound 2^(64 + log_b + log_n - 53)
// (c64.rs:186-208) times (k+1)d terms.  The bit-identical mode (tfhe_core.cuh) stays the parity mode.
//
// Pass structure for N/2 = 2^LG complex points, LG = R1 + R2 + R3 (forward levels 0..LG-1, level l pairs distance 2^(LG-1-l)):
//   P1  diff = rot(acc, e) - acc, signed digits, -> f64, forward levels [0, R1) with compile-time twiddles   (acc -> X)
//   P2  forward levels [R1, R1+R2)                                                                          (X -> X)
//   P3  forward levels [R1+R2, LG) of every limb, multiply-accumulate against the key rows (streamed from L2, coalesced
//       layout), inverse levels [R1+R2, LG) with constant twiddles of both outputs                          (X -> X)
//   P4  inverse levels [R1, R1+R2)                                                                          (X -> X)
//   P5  inverse levels [0, R1), untwist * 1/m, round to torus, acc += result                                (X -> acc)
// 5 barriers per CMUX; 64 KiB of shared memory per ciphertext at TFHE-T (acc 32 KiB + X 32 KiB).
#pragma once
#include "tfhe_core.cuh"

namespace fhe {

HD double f64_fma_rn(double a, double b, double c) {
#if defined(__CUDA_ARCH__)
    return __fma_rn(a, b, c);
#else
    return __builtin_fma(a, b, c);
#endif
}
// (a, b) <- (a + w b, a - w b), 6 fused multiply-adds
HD void bf6(Cx& a, Cx& b, const Cx w) {
    const double tr = f64_fma_rn(-w.im, b.im, f64_fma_rn(w.re, b.re, a.re));
    const double ti = f64_fma_rn(w.im, b.re, f64_fma_rn(w.re, b.im, a.im));
    b.re = f64_fma_rn(2.0, a.re, -tr);
    b.im = f64_fma_rn(2.0, a.im, -ti);
    a.re = tr;
    a.im = ti;
}
// w = 1 and w = -i
HD void bf_one(Cx& a, Cx& b) {
    const Cx s{f64_add_rn(a.re, b.re), f64_add_rn(a.im, b.im)};
    b = Cx{f64_sub_rn(a.re, b.re), f64_sub_rn(a.im, b.im)};
    a = s;
}
HD void bf_minus_i(Cx& a, Cx& b) {  // w b = (b.im, -b.re)
    const Cx s{f64_add_rn(a.re, b.im), f64_sub_rn(a.im, b.re)};
    b = Cx{f64_sub_rn(a.re, b.im), f64_add_rn(a.im, b.re)};
    a = s;
}
// x mod 2^64 rounded to nearest (|x| < 2^115): t = rint(x 2^-64) by the 1.5 * 2^52 trick, y = x - 2^64 t exactly
HD uint64_t f64_to_torus(double x) {
    const double t = f64_mul_rn(x, 0x1p-64);
    const double r = f64_sub_rn(f64_add_rn(t, 0x1.8p52), 0x1.8p52);
    const double y = f64_fma_rn(-r, 0x1p64, x);
#if defined(__CUDA_ARCH__)
    return (uint64_t)__double2ll_rn(y);  // |y| <= 2^63; +2^63 saturates to 2^63 - 1 (one torus ulp, inside the bound)
#else
    if (y >= 0x1p63) return 0x7fffffffffffffffull;
    if (y <= -0x1p63) return 0x8000000000000000ull;
    return (uint64_t)(long long)__builtin_rint(y);
#endif
}
// the top 32 bits of x mod 2^64, rounded to nearest: y = x - 2^64 rint(x 2^-64) as above (|y| <= 2^63), then rint(y 2^-32) sits
// in the low word of y 2^-32 + 1.5 * 2^52
HD uint32_t f64_to_torus32(double x) {
    const double t = f64_mul_rn(x, 0x1p-64);
    const double r = f64_sub_rn(f64_add_rn(t, 0x1.8p52), 0x1.8p52);
    const double y = f64_fma_rn(-r, 0x1p64, x);
#if defined(__CUDA_ARCH__)
    return (uint32_t)__double2loint(__fma_rn(y, 0x1p-32, 0x1.8p52));
#else
    return (uint32_t)(int64_t)__builtin_rint(y * 0x1p-32);
#endif
}
// Accumulator word of the fused path: uint64_t (the full torus word) or uint32_t (its top half; every increment is a sum of f64
// products of magnitude ~2^90 and carries no information below bit 35, see DESIGN.md "TFHE fused kernel")
template <typename A>
HD A f64_to_acc(double x);
template <>
HD uint64_t f64_to_acc<uint64_t>(double x) {
    return f64_to_torus(x);
}
template <>
HD uint32_t f64_to_acc<uint32_t>(double x) {
    return f64_to_torus32(x);
}
HD uint32_t acc_hi_word(uint64_t v) { return (uint32_t)(v >> 32); }
HD uint32_t acc_hi_word(uint32_t v) { return v; }
HD uint64_t acc_to_t64(uint64_t v) { return v; }
HD uint64_t acc_to_t64(uint32_t v) { return (uint64_t)v << 32; }
template <typename A>
HD A t64_to_acc(uint64_t v);
template <>
HD uint64_t t64_to_acc<uint64_t>(uint64_t v) {
    return v;
}
template <>
HD uint32_t t64_to_acc<uint32_t>(uint64_t v) {
    return (uint32_t)((v + 0x80000000ull) >> 32);
}
HD double i32_to_f64(int32_t v) {
#if defined(__CUDA_ARCH__)
    return __int2double_rn(v);
#else
    return (double)v;
#endif
}

// all D signed digits of v (decompose.rs:114-135), least significant first.  The fast path requires log_b d <= 31, so the
// rounding shift (rounding_bits >= 33) only involves the high word of v: (v + 2^(rb-1)) >> rb = (hi(v) + 2^(rb-33)) >> (rb-32)
// (the wrapping add of the reference wraps the high word alike), and the limb / carry chain runs on 32-bit values.
struct FastDigits {
    uint32_t log_b, sh, rnd, mask;  // sh = rounding_bits - 32, rnd = 2^(rounding_bits - 33), mask = 2^log_b - 1
};
inline FastDigits make_fast_digits(const DecompT64& dp) {
    FastDigits f;
    f.log_b = dp.log_b;
    f.sh = dp.rounding_bits - 32;
    f.rnd = 1u << (dp.rounding_bits - 33);
    f.mask = (1u << dp.log_b) - 1u;
    return f;
}
// tables of the fast path for ring degree n = 2m
struct FastFftTab {
    const Cx* W;  // [m]   forward chunk twiddles: level l, chunk c uses W[2^l + c] = zeta^e(l+1, 2c), zeta = e^(i pi / n)
    const Cx* V;  // [m/2] inverse position twiddles V[i] = e^(-2 pi i / m * i)
    const Cx* U;  // [m]   untwist and scale: zeta^(-p) / m
    // per-pass copies laid out so that the threads of a warp read consecutive 16-byte words (one twiddle index t per row):
    const Cx* W2;  // [2^R2 - 1][2^R1]      P2: row (2^u - 1 + top), column hi   = W[2^(R1+u) + (hi << u) + top]
    const Cx* W3;  // [2^R3 - 1][m >> R3]   P3: row (2^u - 1 + top), column g    = W[2^(R1+R2+u) + (g << u) + top]
    const Cx* V4;  // [2^R2 - 1][2^R3]      P4: row (h - 1 + low), column lo     = V[((low << R3) | lo) << (R1+u)], h = 2^(R2-1-u)
    const Cx* V5;  // [R1][m >> R1]         P5: row uu (h = 2^uu), column lo     = V[lo << (R1-1-uu)]; the twiddle of position
                   //                           low > 0 is that times the constant e^(-i pi low / h)
    Cx w0[16];     // W[0..15]: twiddles of the first forward pass (levels 0..3), kept in the kernel's constant bank
    Cx u0[16];     // zeta^(-(i << (LG-R1))): U[lo + (i << L)] = U[lo] * u0[i] (P5 loads one untwist factor per thread)
};
struct TfheFastDev {
    int log_n;
    uint32_t n_lwe;
    FastDigits dig;
    FastFftTab fft;
    const Cx* key;  // [n_lwe][2^R3][2d][2][m >> R3]: key row r, output o at spectral position (g << R3) | i stored at
                    // (((step 2^R3 + i) 2d + r) 2 + o) (m >> R3) + g  (coalesced over the P3 unit index g)
};

template <int D>
HD void t64_digits(const FastDigits& fd, uint32_t hi /* high word of the torus value */, int32_t* dig) {
    uint32_t v = (hi + fd.rnd) >> fd.sh;
#pragma unroll
    for (int k = 0; k < D; ++k) {
        const uint32_t limb = v & fd.mask;
        v = k + 1 < D ? v >> fd.log_b : 0u;  // after the last digit nothing is left (log_b d bits in total)
        const uint32_t carry = (((limb - 1u) | v) & limb) >> (fd.log_b - 1u);
        v += carry;
        dig[k] = (int32_t)(limb - (carry << fd.log_b));
    }
}

// forward register pass over levels l0 .. l0+R-1 of the group with chunk index `hi` at level l0; element j of the group
// sits at distance j << L.  tw(u, top) returns W[2^(l0+u) + (hi << u) + top].
template <int R, typename Tw>
HD void fast_fwd_regs(Cx* x, Tw tw) {
#pragma unroll
    for (int u = 0; u < R; ++u) {
        const int h = 1 << (R - 1 - u);
#pragma unroll
        for (int top = 0; top < (1 << u); ++top) {
            const Cx w = tw(u, top);
#pragma unroll
            for (int low = 0; low < h; ++low) {
                const int j = (top << (R - u)) | low;
                bf6(x[j], x[j + h], w);
            }
        }
    }
}
// inverse register pass over the same levels (processed from l0+R-1 down to l0): decimation in time with position
// twiddles; the element pair at level l0+u has half-size 2^(L+R-1-u) and position ((low << L) | lo) inside its half block
// tw(uu, low) returns V[((low << L) | lo) << (l0+u)], u = R-1-uu (stage with 2^uu positions `low`)
template <int R, typename Tw>
HD void fast_inv_regs(Cx* x, Tw tw) {
#pragma unroll
    for (int uu = 0; uu < R; ++uu) {
        const int u = R - 1 - uu, h = 1 << uu;
#pragma unroll
        for (int low = 0; low < h; ++low) {
            const Cx w = tw(uu, low);
#pragma unroll
            for (int top = 0; top < (1 << u); ++top) {
                const int j = (top << (R - u)) | low;
                bf6(x[j], x[j + h], w);
            }
        }
    }
}
// b * e^(-i pi low / h) for h in {1, 2, 4, 8} (compile-time low, h after unrolling)
HD Cx cx_rot_const(const Cx b, int low, int h) {
    constexpr double S = 0.70710678118654752440, C8 = 0.92387953251128675613, S8 = 0.38268343236508977173;
    if (low == 0) return b;
    if (2 * low == h) return Cx{b.im, -b.re};
    if (4 * low == h) return Cx{f64_mul_rn(f64_add_rn(b.re, b.im), S), f64_mul_rn(f64_sub_rn(b.im, b.re), S)};
    if (4 * low == 3 * h) return Cx{f64_mul_rn(f64_sub_rn(b.im, b.re), S), f64_mul_rn(f64_add_rn(b.re, b.im), -S)};
    // h = 8, odd low: (c - i s), c = cos(pi low / 8), s = sin(pi low / 8)
    const double c = low == 1 ? C8 : low == 3 ? S8 : low == 5 ? -S8 : -C8, sn = (low == 1 || low == 7) ? S8 : C8;
    return Cx{f64_fma_rn(b.im, sn, f64_mul_rn(b.re, c)), f64_fma_rn(-b.re, sn, f64_mul_rn(b.im, c))};
}
// the inverse pass adjacent to the pointwise product (L = 0, lo = 0): twiddles e^(-i pi low / h) are constants
template <int R>
HD void fast_inv_regs_const(Cx* x) {
    static_assert(R >= 1 && R <= 3, "constant inverse pass: radix 2, 4 or 8");
    constexpr double S = 0.70710678118654752440;
#pragma unroll
    for (int uu = 0; uu < R; ++uu) {
        const int u = R - 1 - uu, h = 1 << uu;
#pragma unroll
        for (int low = 0; low < h; ++low) {
#pragma unroll
            for (int top = 0; top < (1 << u); ++top) {
                const int j = (top << (R - u)) | low;
                // w = e^(-i pi low / h): h = 1: 1;  h = 2: 1, -i;  h = 4: 1, (1-i)/sqrt2, -i, (-1-i)/sqrt2
                if (low == 0)
                    bf_one(x[j], x[j + h]);
                else if (2 * low == h)
                    bf_minus_i(x[j], x[j + h]);
                else if (4 * low == h)
                    bf6(x[j], x[j + h], Cx{S, -S});
                else
                    bf6(x[j], x[j + h], Cx{-S, -S});
            }
        }
    }
}

template <int LG_, int R1_, int R2_, int R3_, int D_>
struct TfheFastCfg {
    static constexpr int LG = LG_, R1 = R1_, R2 = R2_, R3 = R3_, D = D_;
    static_assert(R1 + R2 + R3 == LG, "pass split must cover every level");
    static_assert(R1 <= 4 && R3 <= 3, "first pass uses the 16 constant twiddles, last pass the constant inverse");
    static constexpr uint32_t M = 1u << LG, N = 2u << LG, NL = 2 * D;
    static constexpr uint32_t U1 = 2u << (LG - R1);       // P1 / P5 units: (component, group)
    static constexpr uint32_t U2 = NL << (LG - R2);       // P2 units: (limb, group)
    static constexpr uint32_t U3 = 1u << (LG - R3);       // P3 units: group (all limbs, both outputs)
    static constexpr uint32_t U4 = 2u << (LG - R2);       // P4 units: (output, group)
    static constexpr size_t KEY_STRIDE = (size_t)2 * NL * M;  // complex words per CMUX step
};
template <typename C, typename A = uint64_t>
HD size_t tfhe_fast_smem_bytes(uint32_t n_lwe) {
    return (size_t)2 * C::N * sizeof(A) + (size_t)C::NL * C::M * sizeof(Cx) + (((size_t)n_lwe * 2 + 15) & ~(size_t)15);
}

// ---- P1: rotate-subtract, decompose, first forward pass -----------------------------------------------------------------------
template <typename C, typename A>
HD void tfhe_fast_p1(const TfheFastDev& P, const A* __restrict__ acc, Cx* __restrict__ X, uint32_t unit, uint32_t e) {
    constexpr int L = C::LG - C::R1, NE = 1 << C::R1;
    constexpr uint32_t N = C::N, M = C::M;
    const uint32_t j = unit >> L, lo = unit & ((1u << L) - 1u);
    const A* a = acc + (size_t)j * N;
    int32_t dig[C::D][NE][2];
    const uint32_t from0 = (lo + 2 * N - e) & (2 * N - 1);
#pragma unroll
    for (int i = 0; i < NE; ++i) {
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint32_t c = lo + ((uint32_t)i << L) + (uint32_t)half * M;
            const uint32_t from = (from0 + ((uint32_t)i << L) + (uint32_t)half * M) & (2 * N - 1);
            const A r = a[from & (N - 1)];
            const A diff = (A)(((from & N) ? (A)(0 - r) : r) - a[c]);
            int32_t dd[C::D];
            t64_digits<C::D>(P.dig, acc_hi_word(diff), dd);
#pragma unroll
            for (int k = 0; k < C::D; ++k) dig[k][i][half] = dd[k];
        }
    }
#pragma unroll
    for (int k = 0; k < C::D; ++k) {
        Cx x[NE];
#pragma unroll
        for (int i = 0; i < NE; ++i) x[i] = Cx{i32_to_f64(dig[k][i][0]), i32_to_f64(dig[k][i][1])};
        fast_fwd_regs<C::R1>(x, [&](int u, int top) { return P.fft.w0[(1 << u) + top]; });
        Cx* f = X + ((size_t)(j * C::D + k) << C::LG);
        const uint32_t p0 = swz_cx(lo);
#pragma unroll
        for (int i = 0; i < NE; ++i) f[p0 ^ swz_cx((uint32_t)i << L)] = x[i];
    }
}
// ---- P2 / P4: middle passes ---------------------------------------------------------------------------------------------------------
template <typename C, bool FWD>
HD void tfhe_fast_mid(const TfheFastDev& P, Cx* __restrict__ X, uint32_t unit) {
    constexpr int L = C::R3, NE = 1 << C::R2, LGR = C::LG - C::R2;
    const uint32_t g = unit & ((1u << LGR) - 1u);
    Cx* f = X + ((size_t)(unit >> LGR) << C::LG);
    const uint32_t lo = g & ((1u << L) - 1u), hi = g >> L;
    const uint32_t p0 = swz_cx((hi << (L + C::R2)) | lo);
    Cx x[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) x[i] = f[p0 ^ swz_cx((uint32_t)i << L)];
    if (FWD)
        fast_fwd_regs<C::R2>(x, [&](int u, int top) { return ld_cx(P.fft.W2 + (((1 << u) - 1 + top) << C::R1) + hi); });
    else
        fast_inv_regs<C::R2>(x, [&](int uu, int low) { return ld_cx(P.fft.V4 + (((1 << uu) - 1 + low) << C::R3) + lo); });
#pragma unroll
    for (int i = 0; i < NE; ++i) f[p0 ^ swz_cx((uint32_t)i << L)] = x[i];
}
// key rows: every ciphertext reads the same rows at step i, and the CTAs of one SM walk the steps almost in lockstep, so with
// TFHE_KEY_L1 = 1 the rows go through L1 (one CTA's miss could be the others' hit); 0: L2 only (ld.global.cg).  Measured: 326.0 vs
// 327.3 ms per 16 384 PBS - 64 KiB of rows per step do not survive in the ~70 KiB of L1 left beside the twiddle tables.
#ifndef TFHE_KEY_L1
#define TFHE_KEY_L1 0
#endif
HD Cx ld_key(const Cx* p) { return TFHE_KEY_L1 ? ld_cx(p) : ld_cx_stream(p); }
// ---- P3: last forward pass of every limb, multiply-accumulate with the key, first inverse pass ----------------------------------------
// key rows of limb r at the unit's spectral positions I0 .. I1-1: k[2 i + o]
template <typename C, int I0 = 0, int I1 = (1 << C::R3)>
HD void tfhe_fast_p3_keys(const Cx* __restrict__ key, uint32_t g, uint32_t r, Cx* k) {
#pragma unroll
    for (int i = I0; i < I1; ++i) {
        k[2 * i] = ld_key(key + ((size_t)((i * C::NL + r) * 2 + 0) * C::U3 + g));
        k[2 * i + 1] = ld_key(key + ((size_t)((i * C::NL + r) * 2 + 1) * C::U3 + g));
    }
}
// NPRE > 0: kpre holds the rows of limb 0 at the unit's first NPRE spectral positions, requested by the kernel before P2 so that
// their L2 latency hides behind that pass (r02 ncu: half of P3's stall samples are long_scoreboard on these loads)
// KS: `key` points at this step's rows already staged in shared memory (same [position][row][output][unit] order): plain loads
template <typename C, int NPRE = 0, bool KS = false>
HD void tfhe_fast_p3(const TfheFastDev& P, Cx* __restrict__ X, const Cx* __restrict__ key, uint32_t g, const Cx* kpre) {
    constexpr int NE = 1 << C::R3;
    constexpr uint32_t G = C::U3;
    Cx o0[NE], o1[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) o0[i] = o1[i] = Cx{0.0, 0.0};
    const uint32_t p0 = swz_cx(g << C::R3);
#ifndef TFHE_P3K
#define TFHE_P3K 4  // key rows of this many spectral positions are requested before the forward pass of each limb (mode 3, 16 384 PBS: 0 / 2 / 4 / 8 -> 341 / 339 / 326 / 341 ms)
#endif
    constexpr int NPK = KS ? 0 : (TFHE_P3K < NE ? TFHE_P3K : NE);
#pragma unroll 1
    for (uint32_t r = 0; r < C::NL; ++r) {
        const Cx* f = X + ((size_t)r << C::LG);
        Cx kq[NPK > 0 ? 2 * NPK : 1];
        if (NPK > 0) tfhe_fast_p3_keys<C, 0, NPK>(key, g, r, kq);
        Cx x[NE];
#pragma unroll
        for (int i = 0; i < NE; ++i) x[i] = f[p0 ^ (uint32_t)i];  // swz_cx(i) = i for i < 8
        fast_fwd_regs<C::R3>(x, [&](int u, int top) { return ld_cx(P.fft.W3 + (size_t)((1 << u) - 1 + top) * G + g); });
#pragma unroll
        for (int i = 0; i < NE; ++i) {
            Cx k0, k1;
            if (NPRE > 0 && i < NPRE && r == 0) {
                k0 = kpre[2 * i];
                k1 = kpre[2 * i + 1];
            } else if (NPK > 0 && i < NPK) {
                k0 = kq[2 * i];
                k1 = kq[2 * i + 1];
            } else if (KS) {
                k0 = key[(size_t)((i * C::NL + r) * 2 + 0) * G + g];
                k1 = key[(size_t)((i * C::NL + r) * 2 + 1) * G + g];
            } else {
                k0 = ld_key(key + ((size_t)((i * C::NL + r) * 2 + 0) * G + g));
                k1 = ld_key(key + ((size_t)((i * C::NL + r) * 2 + 1) * G + g));
            }
            o0[i] = Cx{f64_fma_rn(-x[i].im, k0.im, f64_fma_rn(x[i].re, k0.re, o0[i].re)), f64_fma_rn(x[i].im, k0.re, f64_fma_rn(x[i].re, k0.im, o0[i].im))};
            o1[i] = Cx{f64_fma_rn(-x[i].im, k1.im, f64_fma_rn(x[i].re, k1.re, o1[i].re)), f64_fma_rn(x[i].im, k1.re, f64_fma_rn(x[i].re, k1.im, o1[i].im))};
        }
    }
    fast_inv_regs_const<C::R3>(o0);
    fast_inv_regs_const<C::R3>(o1);
    Cx* f1 = X + C::M;
#pragma unroll
    for (int i = 0; i < NE; ++i) {
        X[p0 ^ (uint32_t)i] = o0[i];
        f1[p0 ^ (uint32_t)i] = o1[i];
    }
}
// ---- P5: last inverse pass, untwist, round, accumulate ---------------------------------------------------------------------------------
template <typename C, typename A>
HD void tfhe_fast_p5(const TfheFastDev& P, A* __restrict__ acc, const Cx* __restrict__ X, uint32_t unit) {
    constexpr int L = C::LG - C::R1, NE = 1 << C::R1;
    const uint32_t o = unit >> L, lo = unit & ((1u << L) - 1u);
    const Cx* f = X + ((size_t)o << C::LG);
    A* a = acc + (size_t)o * C::N;
    const uint32_t p0 = swz_cx(lo);
    Cx x[NE];
#pragma unroll
    for (int i = 0; i < NE; ++i) x[i] = f[p0 ^ swz_cx((uint32_t)i << L)];
    Cx base = Cx{1.0, 0.0};
    fast_inv_regs<C::R1>(x, [&](int uu, int low) {
        if (low == 0) base = ld_cx(P.fft.V5 + ((size_t)uu << L) + lo);
        return cx_rot_const(base, low, 1 << uu);
    });
#pragma unroll
    const Cx ul = ld_cx(P.fft.U + lo);
#pragma unroll
    for (int i = 0; i < NE; ++i) {
        const uint32_t p = lo + ((uint32_t)i << L);
        const Cx u = i == 0 ? ul : Cx{f64_fma_rn(-ul.im, P.fft.u0[i].im, f64_mul_rn(ul.re, P.fft.u0[i].re)), f64_fma_rn(ul.im, P.fft.u0[i].re, f64_mul_rn(ul.re, P.fft.u0[i].im))};
        const double re = f64_fma_rn(-x[i].im, u.im, f64_mul_rn(x[i].re, u.re));
        const double im = f64_fma_rn(x[i].im, u.re, f64_mul_rn(x[i].re, u.im));
        a[p] += f64_to_acc<A>(re);
        a[p + C::M] += f64_to_acc<A>(im);
    }
}

// One CMUX step acc <- acc + external_product(brk_step, acc.rotate(e) - acc) (tggsw.rs:100-121) on the accumulator in `acc`
// ([2][N] torus words).  run(units, f) calls f(unit) for every unit in [0, units) - spread over the CTA's threads on the
// device, a plain loop in tests/hostsim - followed by a barrier.
template <typename C, typename A, typename Run>
HD void tfhe_fast_cmux(const TfheFastDev& P, A* acc, Cx* X, uint32_t step, uint32_t e, Run run) {
    const Cx* key = P.key + (size_t)step * C::KEY_STRIDE;
    run(C::U1, [&](uint32_t u) { tfhe_fast_p1<C>(P, acc, X, u, e); });
    run(C::U2, [&](uint32_t u) { tfhe_fast_mid<C, true>(P, X, u); });
    run(C::U3, [&](uint32_t u) { tfhe_fast_p3<C>(P, X, key, u, nullptr); });
    run(C::U4, [&](uint32_t u) { tfhe_fast_mid<C, false>(P, X, u); });
    run(C::U1, [&](uint32_t u) { tfhe_fast_p5<C>(P, acc, X, u); });
}

// ---- key transform: one torus polynomial [N] -> its forward spectrum in storage order (position p of the bit-reversed
// output), plain radix-2 levels (setup path; the same twiddle table as the passes above) ----------------------------------------
// level l of the forward transform on s[0..m): every thread handles butterflies b = tid, tid + nthr, ...
HD void tfhe_fast_key_level(Cx* s, int lg, int l, const Cx* __restrict__ W, uint32_t b) {
    const uint32_t h = 1u << (lg - 1 - l), c = b >> (lg - 1 - l), jj = b & (h - 1);
    const uint32_t i0 = (c << (lg - l)) | jj;
    const Cx w = W[(1u << l) + c];
    const Cx t = Cx{f64_fma_rn(-w.im, s[i0 + h].im, f64_mul_rn(w.re, s[i0 + h].re)), f64_fma_rn(w.im, s[i0 + h].re, f64_mul_rn(w.re, s[i0 + h].im))};
    const Cx a = s[i0];
    s[i0] = cx_add(a, t);
    s[i0 + h] = cx_sub(a, t);
}
// destination of spectral position p of key polynomial (step, r, o) in the coalesced P3 layout
template <typename C>
HD size_t tfhe_fast_key_index(uint32_t step, uint32_t r, uint32_t o, uint32_t p) {
    const uint32_t g = p >> C::R3, i = p & ((1u << C::R3) - 1u);
    return (size_t)step * C::KEY_STRIDE + ((size_t)((i * C::NL + r) * 2 + o)) * C::U3 + g;
}

// (log2(N/2), d) -> pass split; f is called with a value of the matching TfheFastCfg type.  Returns false when the fast path
// has no specialisation (the caller then stays on the generic kernels of tfhe_core.cuh).
template <typename F>
inline bool tfhe_fast_dispatch(int lg, unsigned d, F f) {
#define FHE_TFHE_FAST_CASE(LGv, R1v, R2v, R3v, Dv) \
    if (lg == LGv && d == Dv) {                    \
        f(TfheFastCfg<LGv, R1v, R2v, R3v, Dv>{});  \
        return true;                               \
    }
    FHE_TFHE_FAST_CASE(10, 4, 3, 3, 1)
    FHE_TFHE_FAST_CASE(10, 4, 3, 3, 2)
    FHE_TFHE_FAST_CASE(9, 3, 3, 3, 1)
    FHE_TFHE_FAST_CASE(9, 3, 3, 3, 2)
    FHE_TFHE_FAST_CASE(9, 3, 3, 3, 3)
    FHE_TFHE_FAST_CASE(8, 3, 3, 2, 1)
    FHE_TFHE_FAST_CASE(8, 3, 3, 2, 2)
    FHE_TFHE_FAST_CASE(8, 3, 3, 2, 3)
#undef FHE_TFHE_FAST_CASE
    return false;
}

}  // namespace fhe
