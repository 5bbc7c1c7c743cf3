// TFHE building blocks: T64 gadget decomposition, the f64 complex FFT negacyclic product of the reference, CMUX steps.
// __host__ __device__ so tests/hostsim can replay the kernel logic on the CPU.
//
// Reference call sites replaced (util/src unless noted):
//   Base2Decomposor<T64>::decompose / T64::rounding_shr        misc/decompose.rs:66-81, 114-135
//   nega_cyclic_fft64_mul_assign_rt and helpers                 ring/fft/c64.rs:11-108, ring/fft.rs:7-35
//   Tggsw::external_product / cmux, Tglwe rotate/sample_extract scheme/tfhe/src/tggsw.rs:100-121, tglwe.rs:61-66,115-127
//   Bootstrapping::blind_rotate / mod_switch                    scheme/tfhe/src/bootstrapping.rs:84-104
// Floating point: every f64 operation is a single IEEE round-to-nearest multiply, add or subtract in the order the
// reference performs it (num_complex Mul: re = a.re*b.re - a.im*b.im, im = a.re*b.im + a.im*b.re; dit/dif butterflies of
// ring/fft.rs:94-109; final `*= 1/len`), never contracted into FMAs, so every torus word equals the reference's.
#pragma once
#include "fhew_core.cuh"  // f64_mul_rn
#include "modarith.cuh"
#include "rns_core.cuh"   // f64_add_rn

namespace fhe {

HD double f64_sub_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    volatile double r = a - b;
    return r;
#endif
}
struct alignas(16) Cx {  // 16-byte aligned: one LDS.128 / LDG.128 per complex point
    double re, im;
};
HD Cx cx_mul(Cx a, Cx b) {
    const double p0 = f64_mul_rn(a.re, b.re), p1 = f64_mul_rn(a.im, b.im), p2 = f64_mul_rn(a.re, b.im), p3 = f64_mul_rn(a.im, b.re);
    return Cx{f64_sub_rn(p0, p1), f64_add_rn(p2, p3)};
}
HD Cx cx_add(Cx a, Cx b) { return Cx{f64_add_rn(a.re, b.re), f64_add_rn(a.im, b.im)}; }
HD Cx cx_sub(Cx a, Cx b) { return Cx{f64_sub_rn(a.re, b.re), f64_sub_rn(a.im, b.im)}; }

// c64.rs:69-85 f64_mod_u64: round-to-nearest (ties away in magnitude) of v modulo 2^64
HD uint64_t f64_mod_u64_dev(double v) {
#if defined(__CUDA_ARCH__)
    const uint64_t bits = (uint64_t)__double_as_longlong(v);
#else
    uint64_t bits;
    __builtin_memcpy(&bits, &v, 8);
#endif
    const uint64_t sign = bits >> 63;
    const int64_t exponent = (int64_t)((bits >> 52) & 0x7ff);
    const uint64_t mantissa = (bits << 11) | 0x8000000000000000ull;
    const int64_t shift = 1086 - exponent;
    uint64_t value;
    if (shift >= -63 && shift <= 0)
        value = mantissa << (-shift);
    else if (shift >= 1 && shift <= 64)
        value = ((mantissa >> (shift - 1)) + 1) >> 1;
    else
        value = 0;
    return sign == 0 ? value : (uint64_t)(0 - value);
}
// T64::to_i64() as f64 (torus.rs:20-25)
HD double t64_to_f64(uint64_t v) {
#if defined(__CUDA_ARCH__)
    return __ll2double_rn((long long)v);
#else
    return (double)(int64_t)v;
#endif
}

// ---- T64 gadget decomposition (decompose.rs:66-81, 114-135) -------------------------------------------------------------
struct DecompT64 {
    uint32_t log_b, d, rounding_bits;
};
inline DecompT64 make_decomp_t64(uint32_t log_b, uint32_t d) {
    DecompT64 p;
    p.log_b = log_b;
    p.d = d;
    p.rounding_bits = 64 > log_b * d ? 64 - log_b * d : 0;
    return p;
}
HD uint64_t t64_rounding_shr_dev(uint64_t v, uint32_t bits) {
    if (bits == 0) return v;
    if (bits >= 64) return 0;
    return (v + ((1ull << bits) >> 1)) >> bits;
}
// digit `which` (0 = least significant) of v
HD uint64_t t64_digit(const DecompT64& dp, uint64_t v, uint32_t which) {
    v = t64_rounding_shr_dev(v, dp.rounding_bits);
    const uint64_t mask = (1ull << dp.log_b) - 1;
    uint64_t dig = 0;
    for (uint32_t k = 0; k <= which; ++k) {
        const uint64_t limb = v & mask;
        v >>= dp.log_b;
        const uint64_t carry = (((limb - 1) | v) & limb) >> (dp.log_b - 1);
        v += carry;
        dig = limb - (carry << dp.log_b);
    }
    return dig;
}

// global loads of complex values: `ld_cx` for tables that should stay in L1 (read-only path), `ld_cx_stream` for the key
// spectra, which are read once per CMUX and must not evict the twiddle tables from L1 (L2 only: ld.global.cg)
HD Cx ld_cx(const Cx* p) {
#if defined(__CUDA_ARCH__)
    const double2 v = __ldg(reinterpret_cast<const double2*>(p));
    return Cx{v.x, v.y};
#else
    return *p;
#endif
}
HD Cx ld_cx_stream(const Cx* p) {
#if defined(__CUDA_ARCH__)
    const double2 v = __ldcg(reinterpret_cast<const double2*>(p));
    return Cx{v.x, v.y};
#else
    return *p;
#endif
}

// ---- shared-memory swizzle for 16-byte complex elements -------------------------------------------------------------------
HD uint32_t swz_cx(uint32_t p) { return p ^ ((p >> 3) & 7u); }

// first-pass radix of an FFT of 2^lg complex points (then radix-8 passes); lg <= 4 is a single pass
HD constexpr int fft_r1(int lg) { return lg <= 4 ? lg : (lg % 3 == 0 ? 3 : (lg % 3 == 1 ? 4 : 2)); }

// ---- register passes (same index scheme as the NTT: element j of a group at (hi << (L+R)) | (j << L) | lo) ----------------
// forward (fft.rs:9-19): stages l0 .. l0+R-1, chunk twiddle tw_bo[(hi << u) + top], dit butterfly
template <int R>
HD void fft_fwd_regs(Cx* x, const Cx* __restrict__ tw_bo, uint32_t hi) {
#pragma unroll
    for (int u = 0; u < R; ++u) {
        const int h = 1 << (R - 1 - u);
#pragma unroll
        for (int top = 0; top < (1 << u); ++top) {
            const Cx t = ld_cx(tw_bo + ((hi << u) + top));
#pragma unroll
            for (int low = 0; low < h; ++low) {
                const int j = (top << (R - u)) | low;
                const Cx tb = cx_mul(t, x[j + h]);
                const Cx a = x[j];
                x[j] = cx_add(a, tb);
                x[j + h] = cx_sub(a, tb);
            }
        }
    }
}
// inverse (fft.rs:23-35 without the final scaling): dif butterfly, stages l0+R-1 down to l0
template <int R>
HD void fft_inv_regs(Cx* x, const Cx* __restrict__ tw_inv_bo, uint32_t hi) {
#pragma unroll
    for (int u = R - 1; u >= 0; --u) {
        const int h = 1 << (R - 1 - u);
#pragma unroll
        for (int top = 0; top < (1 << u); ++top) {
            const Cx t = ld_cx(tw_inv_bo + ((hi << u) + top));
#pragma unroll
            for (int low = 0; low < h; ++low) {
                const int j = (top << (R - u)) | low;
                const Cx a = x[j], b = x[j + h];
                x[j] = cx_add(a, b);
                x[j + h] = cx_mul(cx_sub(a, b), t);
            }
        }
    }
}

// A pass over one FFT of 2^lg points held in (swizzled) shared memory `s`; `load(p)` / `store(p, v)` may be replaced by
// the caller to fuse the producer / consumer of the first / last pass.  unit = group index in [0, 2^(lg-R)).
template <int R, bool FWD, typename Load, typename Store>
HD void fft_pass_unit(int lg, int l0, uint32_t grp, const Cx* __restrict__ tw, Load load, Store store) {
    const int L = lg - l0 - R;
    const uint32_t lo = grp & ((1u << L) - 1u), hi = grp >> L;
    const uint32_t base = (hi << (L + R)) | lo;
    Cx x[1 << R];
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) x[j] = load(base | ((uint32_t)j << L));
    if (FWD)
        fft_fwd_regs<R>(x, tw, hi);
    else
        fft_inv_regs<R>(x, tw, hi);
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) store(base | ((uint32_t)j << L), x[j]);
}
template <bool FWD, typename Load, typename Store>
HD void fft_pass_unit_dyn(int r, int lg, int l0, uint32_t grp, const Cx* __restrict__ tw, Load load, Store store) {
    switch (r) {
        case 1: fft_pass_unit<1, FWD>(lg, l0, grp, tw, load, store); break;
        case 2: fft_pass_unit<2, FWD>(lg, l0, grp, tw, load, store); break;
        case 3: fft_pass_unit<3, FWD>(lg, l0, grp, tw, load, store); break;
        default: fft_pass_unit<4, FWD>(lg, l0, grp, tw, load, store); break;
    }
}

// device-side view of the f64 tables for ring degree n (m = n/2 complex points, lg = log2 m)
struct FftTab {
    int lg;                // log2(n/2)
    const Cx* tw;          // [n/2]  cis(j*pi/n)            (c64.rs:20-28 twist)
    const Cx* tw_inv;      // [n/2]  conj
    const Cx* tw_bo;       // [m/2 (>=1)] bit-reversed cis(i*pi/m) prefix used by fft_in_place (chunk twiddles)
    const Cx* tw_inv_bo;   // [m/2 (>=1)]
    double m_inv;          // 1.0 / m
};

// pass plan of an FFT of 2^lg points: pass i covers stages l0[i] .. l0[i]+r[i]-1
struct FftPlan {
    int n;
    int l0[6], r[6];
};
HD FftPlan make_fft_plan(int lg) {
    FftPlan p;
    p.n = 0;
    int t = 0;
    if (lg == 0) return p;
    const int r1 = fft_r1(lg);
    p.l0[p.n] = 0;
    p.r[p.n] = r1;
    ++p.n;
    t = r1;
    while (t < lg) {
        p.l0[p.n] = t;
        p.r[p.n] = 3;
        ++p.n;
        t += 3;
    }
    return p;
}

}  // namespace fhe

namespace fhe {

// ---- CTA-level FFT over `nf` transforms stored back to back in shared memory ------------------------------------------------
// `run(phase)` executes phase(tid, nthr) for every thread of the CTA followed by a barrier (device: __syncthreads();
// tests/hostsim: a sequential loop over tid).
template <bool FWD, typename Run>
HD void fft_run(Cx* s, uint32_t nf, const FftTab& T, Run run) {
    const int lg = T.lg;
    // pass plan in closed form (no arrays: they would live in local memory): pass 0 has fft_r1(lg) stages, the rest 3
    const int r1 = fft_r1(lg), npass = lg == 0 ? 0 : 1 + (lg - r1) / 3;
    const Cx* tw = FWD ? T.tw_bo : T.tw_inv_bo;
    for (int pp = 0; pp < npass; ++pp) {
        const int pi = FWD ? pp : npass - 1 - pp;
        const int r = pi == 0 ? r1 : 3, l0 = pi == 0 ? 0 : r1 + 3 * (pi - 1);
        const uint32_t lgroups = (uint32_t)(lg - r);
        run([&](uint32_t tid, uint32_t nthr) {
            const uint32_t total = nf << lgroups;
            for (uint32_t u = tid; u < total; u += nthr) {
                Cx* f = s + ((size_t)(u >> lgroups) << lg);
                fft_pass_unit_dyn<FWD>(
                    r, lg, l0, u & ((1u << lgroups) - 1u), tw, [&](uint32_t p) { return f[swz_cx(p)]; },
                    [&](uint32_t p, Cx v) { f[swz_cx(p)] = v; });
            }
        });
    }
}

// to_c64_twisted (c64.rs:20-28) of one polynomial given as a coefficient functor: F[swz(p)] = (c(p), c(p + m)) * tw[p]
template <typename Coef>
HD void fft_twist_in(Cx* f, const FftTab& T, Coef coef, uint32_t tid, uint32_t nthr) {
    const uint32_t m = 1u << T.lg;
    for (uint32_t p = tid; p < m; p += nthr) f[swz_cx(p)] = cx_mul(Cx{t64_to_f64(coef(p)), t64_to_f64(coef(p + m))}, ld_cx(T.tw + p));
}
// tail of ifft_in_place (`*= 1/len`, fft.rs:32-34) + assign_from_c64_twisted (c64.rs:31-41) for element p: (lo, hi) words
HD void fft_untwist_out(const Cx* f, const FftTab& T, uint32_t p, uint64_t& lo, uint64_t& hi) {
    Cx c = f[swz_cx(p)];
    c.re = f64_mul_rn(c.re, T.m_inv);
    c.im = f64_mul_rn(c.im, T.m_inv);
    const Cx x = cx_mul(c, ld_cx(T.tw_inv + p));
    lo = f64_mod_u64_dev(x.re);
    hi = f64_mod_u64_dev(x.im);
}

// ---- TGGSW external product / CMUX on one TGLWE accumulator held in shared memory ----------------------------------------------
struct TfheDev {
    int log_n;            // ring degree N = 2^log_n
    uint32_t k;           // GLWE dimension
    uint32_t n_lwe;       // TLWE dimension n
    DecompT64 bs_dec;     // TGGSW decomposor
    FftTab fft;
    const Cx* brk;        // [n_lwe][(k+1)*d rows][(k+1) outputs][N/2] Fourier-domain key polynomials
    uint32_t fourier_acc; // 0: reference dataflow (every row*limb product rounded on its own: bit-identical torus words);
                          // 1: sum the products in the Fourier domain, one inverse transform per output (error below the
                          //    reference's own bound, decryptions identical; (k+1) instead of (k+1)^2 d inverse FFTs per CMUX)
};
// shared memory: acc[(k+1)][N] u64 | F[(k+1)*d][N/2] Cx | P[(k+1)*d][N/2] Cx
HD size_t tfhe_smem_bytes(uint32_t k, uint32_t d, int log_n) {
    const size_t n = (size_t)1 << log_n;
    return (size_t)(k + 1) * n * 8 + 2 * (size_t)(k + 1) * d * (n / 2) * sizeof(Cx);
}
// rot(src, e)[c] for e in [0, 2N): coefficient c of src * X^e (ring.rs:299-313; tglwe.rs:61-66)
HD uint64_t t64_rot_coef(const uint64_t* src, uint32_t n, uint32_t e, uint32_t c) {
    const uint32_t from = (c + 2 * n - e) & (2 * n - 1);
    const uint64_t v = src[from & (n - 1)];
    return from >= n ? (uint64_t)(0 - v) : v;
}
// One external product with TGGSW `key` ([(k+1)d][(k+1)][N/2] Cx) applied to the polynomials given by `src(j, c)`
// (j = component 0..k, c = coefficient); the result component o is passed to `sink(o, c_lo, lo, c_hi, hi)` as two
// coefficients at a time; `sink` is called by exactly one thread per (o, coefficient) after all reads of `src`.
template <typename Src, typename Sink, typename Run>
HD void tfhe_external_product(const TfheDev& P, Cx* F, Cx* Pb, const Cx* __restrict__ key, Src src, Sink sink, Run run) {
    const uint32_t n = 1u << P.log_n, m = n >> 1, d = P.bs_dec.d, nl = (P.k + 1) * d;
    const int lg = P.fft.lg;
    // limbs in the order [a_0 digits.., a_{k-1} digits.., b digits] (tggsw.rs:106-108): limb r = j*d + i
    run([&](uint32_t tid, uint32_t nthr) {
        for (uint32_t r = 0; r < nl; ++r) {
            const uint32_t j = r / d, i = r % d;
            fft_twist_in(F + ((size_t)r << lg), P.fft, [&](uint32_t c) { return t64_digit(P.bs_dec, src(j, c), i); }, tid, nthr);
        }
    });
    fft_run<true>(F, nl, P.fft, run);
    for (uint32_t o = 0; o <= P.k; ++o) {
        // each product row_r.{a_o | b} * limb_r is inverse-transformed and rounded on its own (Dot sums T64 products,
        // misc.rs:59-61), exactly like the reference
        run([&](uint32_t tid, uint32_t nthr) {
#pragma unroll 4
            for (uint32_t u = tid; u < nl * m; u += nthr) {
                const uint32_t r = u >> lg, p = u & (m - 1);
                Pb[((size_t)r << lg) + swz_cx(p)] = cx_mul(F[((size_t)r << lg) + swz_cx(p)], ld_cx_stream(key + (((size_t)(r * (P.k + 1) + o) << lg) + p)));
            }
        });
        fft_run<false>(Pb, nl, P.fft, run);
        run([&](uint32_t tid, uint32_t nthr) {
            for (uint32_t p = tid; p < m; p += nthr) {
                uint64_t slo = 0, shi = 0;
                for (uint32_t r = 0; r < nl; ++r) {
                    uint64_t lo, hi;
                    fft_untwist_out(Pb + ((size_t)r << lg), P.fft, p, lo, hi);
                    slo += lo;
                    shi += hi;
                }
                sink(o, p, slo, p + m, shi, true);
            }
        });
    }
}
// ---- compile-time specialisation for N/2 = 2^LG points (LG = 8, 9, 10) -------------------------------------------------------------
// Same schedule, arithmetic and operation order as tfhe_external_product; every pass has its radix, stride and swizzle
// masks as compile-time constants (swz_cx is linear over XOR: swz(base | j << L) = swz(base) ^ const_j), 32-bit indices
// and constant trip counts.  The generic version spends ~3/4 of its issue slots on index arithmetic; this one was written
// after the ncu source view showed the kernel at 48 % issue utilisation with only 27 % of instructions on the FP64 pipe.
template <int R, bool FWD, int LG, int L0>
HD void fft_pass_unit_c(uint32_t grp, const Cx* __restrict__ tw, Cx* f) {
    constexpr int L = LG - L0 - R;
    const uint32_t lo = grp & ((1u << L) - 1u), hi = grp >> L;
    const uint32_t P0 = swz_cx((hi << (L + R)) | lo);
    Cx x[1 << R];
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) x[j] = f[P0 ^ swz_cx((uint32_t)j << L)];
    if (FWD)
        fft_fwd_regs<R>(x, tw, hi);
    else
        fft_inv_regs<R>(x, tw, hi);
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) f[P0 ^ swz_cx((uint32_t)j << L)] = x[j];
}
template <int R, bool FWD, int LG, int L0>
HD void fft_pass_c(Cx* s, uint32_t nf, const Cx* __restrict__ tw, uint32_t tid, uint32_t nthr) {
    constexpr uint32_t LGR = LG - R;
    for (uint32_t u = tid; u < (nf << LGR); u += nthr) fft_pass_unit_c<R, FWD, LG, L0>(u & ((1u << LGR) - 1u), tw, s + ((u >> LGR) << LG));
}
template <int LG, bool FWD, typename Run>
HD void fft_run_c(Cx* s, uint32_t nf, const FftTab& T, Run run) {
    constexpr int R1 = fft_r1(LG), NPASS = 1 + (LG - R1) / 3;
    static_assert(NPASS >= 2 && NPASS <= 4, "specialised for 5 <= LG <= 13");
    const Cx* tw = FWD ? T.tw_bo : T.tw_inv_bo;
    if (FWD) {
        run([&](uint32_t tid, uint32_t nthr) { fft_pass_c<R1, true, LG, 0>(s, nf, tw, tid, nthr); });
        run([&](uint32_t tid, uint32_t nthr) { fft_pass_c<3, true, LG, R1>(s, nf, tw, tid, nthr); });
        if (NPASS > 2) run([&](uint32_t tid, uint32_t nthr) { fft_pass_c<3, true, LG, (NPASS > 2 ? R1 + 3 : R1)>(s, nf, tw, tid, nthr); });
        if (NPASS > 3) run([&](uint32_t tid, uint32_t nthr) { fft_pass_c<3, true, LG, (NPASS > 3 ? R1 + 6 : R1)>(s, nf, tw, tid, nthr); });
    } else {
        if (NPASS > 3) run([&](uint32_t tid, uint32_t nthr) { fft_pass_c<3, false, LG, (NPASS > 3 ? R1 + 6 : R1)>(s, nf, tw, tid, nthr); });
        if (NPASS > 2) run([&](uint32_t tid, uint32_t nthr) { fft_pass_c<3, false, LG, (NPASS > 2 ? R1 + 3 : R1)>(s, nf, tw, tid, nthr); });
        run([&](uint32_t tid, uint32_t nthr) { fft_pass_c<3, false, LG, R1>(s, nf, tw, tid, nthr); });
        run([&](uint32_t tid, uint32_t nthr) { fft_pass_c<R1, false, LG, 0>(s, nf, tw, tid, nthr); });
    }
}
template <int LG, typename Src, typename Sink, typename Run>
HD void tfhe_external_product_c(const TfheDev& P, Cx* F, Cx* Pb, const Cx* __restrict__ key, Src src, Sink sink, Run run) {
    constexpr uint32_t m = 1u << LG;
    const uint32_t d = P.bs_dec.d, nl = (P.k + 1) * d, ko = P.k + 1;
    const FftTab& T = P.fft;
    run([&](uint32_t tid, uint32_t nthr) {
        for (uint32_t r = 0; r < nl; ++r) {
            const uint32_t j = r / d, i = r % d;
            Cx* f = F + (r << LG);
            for (uint32_t p = tid; p < m; p += nthr)
                f[swz_cx(p)] = cx_mul(Cx{t64_to_f64(t64_digit(P.bs_dec, src(j, p), i)), t64_to_f64(t64_digit(P.bs_dec, src(j, p + m), i))}, ld_cx(T.tw + p));
        }
    });
    fft_run_c<LG, true>(F, nl, T, run);
    for (uint32_t o = 0; o <= P.k; ++o) {
        run([&](uint32_t tid, uint32_t nthr) {
            for (uint32_t r = 0; r < nl; ++r) {
                const Cx* f = F + (r << LG);
                Cx* pb = Pb + (r << LG);
                const Cx* kk = key + ((size_t)(r * ko + o) << LG);
#pragma unroll 4
                for (uint32_t p = tid; p < m; p += nthr) pb[swz_cx(p)] = cx_mul(f[swz_cx(p)], ld_cx_stream(kk + p));
            }
        });
        fft_run_c<LG, false>(Pb, nl, T, run);
        run([&](uint32_t tid, uint32_t nthr) {
            for (uint32_t p = tid; p < m; p += nthr) {
                uint64_t slo = 0, shi = 0;
                const Cx ti = ld_cx(T.tw_inv + p);
                for (uint32_t r = 0; r < nl; ++r) {
                    Cx c = Pb[(r << LG) + swz_cx(p)];
                    c.re = f64_mul_rn(c.re, T.m_inv);
                    c.im = f64_mul_rn(c.im, T.m_inv);
                    const Cx x = cx_mul(c, ti);
                    slo += f64_mod_u64_dev(x.re);
                    shi += f64_mod_u64_dev(x.im);
                }
                sink(o, p, slo, p + m, shi, true);
            }
        });
    }
}
// Fourier-domain accumulation variant (TfheDev::fourier_acc): out_o = IFFT(sum_r FFT(limb_r) o key[r][o]).  NOT the
// reference's rounding order: each output coefficient carries one rounding instead of (k+1)d, so it differs from the
// reference's torus words by at most the reference's own per-product error bound 2^(64 + log_b + log_n - 53) (c64.rs:186-208)
// times (k+1)d, far below the plaintext scale; decrypted results are identical (tests/test_gpu_tfhe.py).
template <int LG, typename Src, typename Sink, typename Run>
HD void tfhe_external_product_acc_c(const TfheDev& P, Cx* F, Cx* Pb, const Cx* __restrict__ key, Src src, Sink sink, Run run) {
    constexpr uint32_t m = 1u << LG;
    const uint32_t d = P.bs_dec.d, nl = (P.k + 1) * d, ko = P.k + 1;
    const FftTab& T = P.fft;
    run([&](uint32_t tid, uint32_t nthr) {
        for (uint32_t r = 0; r < nl; ++r) {
            const uint32_t j = r / d, i = r % d;
            Cx* f = F + (r << LG);
            for (uint32_t p = tid; p < m; p += nthr)
                f[swz_cx(p)] = cx_mul(Cx{t64_to_f64(t64_digit(P.bs_dec, src(j, p), i)), t64_to_f64(t64_digit(P.bs_dec, src(j, p + m), i))}, ld_cx(T.tw + p));
        }
    });
    fft_run_c<LG, true>(F, nl, T, run);
    run([&](uint32_t tid, uint32_t nthr) {
        for (uint32_t u = tid; u < ko * m; u += nthr) {
            const uint32_t o = u >> LG, p = u & (m - 1);
            Cx s = cx_mul(F[swz_cx(p)], ld_cx_stream(key + ((size_t)o << LG) + p));
            for (uint32_t r = 1; r < nl; ++r) s = cx_add(s, cx_mul(F[(r << LG) + swz_cx(p)], ld_cx_stream(key + ((size_t)(r * ko + o) << LG) + p)));
            Pb[(o << LG) + swz_cx(p)] = s;
        }
    });
    fft_run_c<LG, false>(Pb, ko, T, run);
    run([&](uint32_t tid, uint32_t nthr) {
        for (uint32_t u = tid; u < ko * m; u += nthr) {
            const uint32_t o = u >> LG, p = u & (m - 1);
            Cx c = Pb[(o << LG) + swz_cx(p)];
            c.re = f64_mul_rn(c.re, T.m_inv);
            c.im = f64_mul_rn(c.im, T.m_inv);
            const Cx x = cx_mul(c, ld_cx(T.tw_inv + p));
            sink(o, p, f64_mod_u64_dev(x.re), p + m, f64_mod_u64_dev(x.im), true);
        }
    });
}
template <typename Src, typename Sink, typename Run>
HD void tfhe_external_product_any(const TfheDev& P, Cx* F, Cx* Pb, const Cx* __restrict__ key, Src src, Sink sink, Run run) {
    if (P.fourier_acc) {
        switch (P.fft.lg) {
            case 10: tfhe_external_product_acc_c<10>(P, F, Pb, key, src, sink, run); return;
            case 9: tfhe_external_product_acc_c<9>(P, F, Pb, key, src, sink, run); return;
            case 8: tfhe_external_product_acc_c<8>(P, F, Pb, key, src, sink, run); return;
            default: break;  // other sizes: reference dataflow
        }
    }
    switch (P.fft.lg) {
        case 10: tfhe_external_product_c<10>(P, F, Pb, key, src, sink, run); break;
        case 9: tfhe_external_product_c<9>(P, F, Pb, key, src, sink, run); break;
        case 8: tfhe_external_product_c<8>(P, F, Pb, key, src, sink, run); break;
        default: tfhe_external_product(P, F, Pb, key, src, sink, run); break;
    }
}
// Measured alternatives (B200, TFHE-T, batch 2048): folding the twist into the first forward pass, the pointwise
// product into the first inverse pass and the untwist into the last inverse pass removes up to 4 shared-memory round trips
// and 5 barriers per CMUX but leaves passes with half of the threads idle; it ran at 10.5k (fully fused) and 11.9k (radix-8
// fused passes only) PBS/s against 12.3k for this plain schedule; repeated on top of the compile-time specialisation below
// (3+4+3 passes with twist and pointwise product fused into the radix-8 end passes): 12.7k against 14.2k.  Longer
// per-thread dependency chains cost more than the saved shared-memory round trips at 16 warps per SM, so the plain
// schedule stays.
// acc <- cmux(brk_i, acc, acc.rotate(e)) = acc + external_product(brk_i, acc.rotate(e) - acc)   (tggsw.rs:114-121)
template <typename Run>
HD void tfhe_cmux_step(const TfheDev& P, uint64_t* acc, Cx* F, Cx* Pb, uint32_t i, uint32_t e, Run run) {
    const uint32_t n = 1u << P.log_n, d = P.bs_dec.d;
    const Cx* key = P.brk + (((size_t)i * (P.k + 1) * d * (P.k + 1)) << P.fft.lg);
    tfhe_external_product_any(
        P, F, Pb, key, [&](uint32_t j, uint32_t c) { return t64_rot_coef(acc + (size_t)j * n, n, e, c) - acc[(size_t)j * n + c]; },
        [&](uint32_t o, uint32_t c0, uint64_t v0, uint32_t c1, uint64_t v1, bool) {
            acc[(size_t)o * n + c0] += v0;
            acc[(size_t)o * n + c1] += v1;
        },
        run);
}

}  // namespace fhe
