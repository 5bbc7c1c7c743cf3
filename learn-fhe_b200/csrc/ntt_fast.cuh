// Fast-path negacyclic NTT building blocks (second generation), sm_100a.
//
// Same transform as ntt_core.cuh (reference: util/src/ring/fft.rs:40-77 + util/src/ring/fft/zq.rs:58-67; natural ->
// bit-reversed forward, bit-reversed -> natural inverse incl. n^-1) but organised around the measured bottleneck of
// the first generation (instruction issue, not HBM or the IMAD pipe: ~48 SASS instructions per u64 butterfly):
//   * long lazy-reduction chains: forward butterflies never reduce x (values grow by 2q (u32) / 4q (u64) per stage
//     inside the word's headroom; one min() per element where the bound would overflow, one Barrett at the end);
//     u64 products use a 3-multiply approximate high word (result in [0,4q));
//   * the first pass of a tile reads global memory straight into registers and the last pass writes registers
//     straight to global memory (vectorised 16-byte accesses), so a polynomial crosses shared memory once per
//     intermediate pass only;
//   * an XOR swizzle that is GF(2)-linear, so the swizzled address of element j of a radix group is
//     swz(base) ^ const_j: one LOP per element instead of a swizzle evaluation;
//   * per-limb modulus descriptors, so one launch transforms a whole RNS batch [polys][N] with limb = poly % limbs.
// Preconditions of the fast path (checked by the launcher, which otherwise uses the generic kernels of
// ntt_kernels.cuh): u32: q < 2^28; u64: q < 2^56; 9 <= log_n <= 17.
// Everything is __host__ __device__ so tests/hostsim can replay the passes thread by thread on the CPU.
#pragma once
#ifndef FAST_PAIR32
#define FAST_PAIR32 1
#endif
#ifndef FAST_HALF64
#define FAST_HALF64 1
#endif
#ifndef FAST_HALF32
#define FAST_HALF32 0
#endif
#ifndef FAST_MIN_NTHR
#define FAST_MIN_NTHR 32  // smallest CTA: tiles covered by fewer threads are batched PB per CTA.  128 measured 5-11 % faster than 256 at N = 2^10, 2^11; 32 (one tile per CTA for the 32-thread u32 2^10 tile) another 6-13 % at 4096 x 2^10 u32 (2.01 -> 2.27 TB/s) and +8 % at 65 536 polynomials, neutral elsewhere
#endif
#include "modarith.cuh"
#include "ntt_core.cuh"

namespace fhe {

// ---- swizzle (word index within one polynomial tile) ---------------------------------------------------------------
// u32: bits 5,6,7 -> bits 2,3,4; u64: bits 4,5,6 -> bits 1,2,3.  Conflict-free for (a) 32 consecutive words,
// (b) radix-8 groups at stride 8 (lanes over 3 low bits + group bits at 6..), (c) 16-byte vector accesses of 8
// contiguous words per lane.  Linear over XOR: swz(a ^ b) = swz(a) ^ swz(b).
template <typename W>
HD uint32_t swz2(uint32_t p);
template <>
HD uint32_t swz2<uint32_t>(uint32_t p) {
    return p ^ ((p >> 3) & 0x1Cu);
}
template <>
HD uint32_t swz2<uint64_t>(uint32_t p) {
    return p ^ ((p >> 3) & 0xEu);
}

// ---- lazy modular arithmetic ------------------------------------------------------------------------------------------
// u32, q < 2^28: 16q <= 2^32.  Forward values may grow to 16q; inverse values stay in [0,2q).
struct Lz32 {
    typedef uint32_t W;
    static constexpr int BITS = 32;
    static constexpr int FWD_GROW = 2;    // bound growth per forward stage, in units of q
    static constexpr int FWD_LIMIT = 16;  // values must stay < FWD_LIMIT * q
    uint32_t q, q2, q8, mu;               // mu = floor(2^32 / q)
    HD uint32_t mul_lazy(uint32_t y, uint32_t w, uint32_t wp) const { return w * y - mulhi_u32(wp, y) * q; }  // [0,2q), any y
    // x' = x + t*y, y' = x + 2q - t*y; the two-operand add is pinned to the ALU pipe (alu_add) so that the butterfly costs
    // exactly its 4 algorithmic slots on the IMAD pipe, the measured bottleneck.
    HD void bf_fwd(uint32_t& x, uint32_t& y, TwPair<uint32_t> t) const {
        const uint32_t x0 = x;
        const uint32_t ty = t.w * y - mulhi_u32(t.wp, y) * q;  // IMAD, IMAD.HI, IMAD: 4 slots on the IMAD pipe
        x = alu_add(x0, ty);                                   // VIADDMNMX (ALU)
        y = x0 + q2 - ty;                                      // IADD3 (ALU)
    }
    HD uint32_t pre_red(uint32_t x) const { return umin_(x, x - q8); }  // [0,16q) -> [0,8q)
    HD uint32_t canon(uint32_t x) const {                               // any x -> [0,q)
        const uint32_t r = x - mulhi_u32(x, mu) * q;                    // [0,2q)
        return umin_(r, r - q);
    }
    // inverse: x,y in [0,2q) -> [0,2q)
    HD void bf_inv(uint32_t& x, uint32_t& y, TwPair<uint32_t> t, int /*stage*/) const {
        const uint32_t s = alu_add(x, y), d = x + q2 - y;
        x = umin_(s, s - q2);
        y = mul_lazy(d, t.w, t.wp);
    }
    HD void bf_inv_last(uint32_t& x, uint32_t& y, TwPair<uint32_t> ninv, TwPair<uint32_t> wninv, int /*stage*/) const {
        const uint32_t s = alu_add(x, y), d = x + q2 - y;
        x = mul_lazy(s, ninv.w, ninv.wp);
        y = mul_lazy(d, wninv.w, wninv.wp);
    }
    HD uint32_t inv_pass_fix(uint32_t x) const { return x; }        // nothing to do: the invariant is [0,2q)
    HD uint32_t inv_canon(uint32_t x) const { return umin_(x, x - q); }  // [0,2q) -> [0,q)
};

// u64, q < 2^56.  Products via a 3-multiply approximate high word: result in [0,4q).
// Forward: growth 4q per stage, never reduced (17 stages: 69q < 2^64).  Inverse: pass-level invariant "every value
// < 16q"; inside a pass of R <= 4 stages the un-multiplied sum chain reaches 16q * 2^R <= 256q and is Barrett-reduced
// at the end of the pass (elements 0 and, for R = 4, 1 of each register group).
struct Lz64 {
    typedef uint64_t W;
    static constexpr int BITS = 64;
    static constexpr int FWD_GROW = 4;
    static constexpr int FWD_LIMIT = 128;  // < 2^64 / q for q < 2^56 (only checked by the planner)
    uint64_t q, nq, q2, q4, q16;           // nq = 2^64 - q
    uint32_t mu32, sh_e, sh_f, pad_;       // one-multiply Barrett: see barrett2
    HD uint64_t mul_lazy(uint64_t y, uint64_t w, uint64_t wp) const { return w * y + mulhi_u64_approx(wp, y) * nq; }  // [0,4q)
    HD void bf_fwd(uint64_t& x, uint64_t& y, TwPair<uint64_t> t) const {
        const uint64_t ty = mul_lazy(y, t.w, t.wp);
        const uint64_t x0 = x;
        x = x0 + ty;
        y = x0 + q4 - ty;
    }
    HD uint64_t pre_red(uint64_t x) const { return x; }
    // x < 2^8 * 2^k (k = bit length of q; every lazy bound of this path is <= 256 q) -> [0, 2q) with ONE 32-bit high
    // multiply: xs = x >> e < 2^32 (e = max(k - 24, 0)), mu32 = floor(2^(32+f+e) / q) < 2^32 (f = k - 1 - e), and
    // Qh = (xs * mu32) >> (32 + f) is floor(x / q) or one less (error terms x / 2^(k+31) + 2^e / q < 1).
    HD uint64_t barrett2(uint64_t x) const {
        const uint32_t qh = mulhi_u32((uint32_t)(x >> sh_e), mu32) >> sh_f;
        return x + (uint64_t)qh * nq;
    }
    HD uint64_t canon(uint64_t x) const {
        const uint64_t r = barrett2(x);
        return umin_(r, r - q);
    }
    // stage = 0 for the first stage executed inside the pass (operands < 16q), 1 for the next (sum chain < 32q), ...
    HD void bf_inv(uint64_t& x, uint64_t& y, TwPair<uint64_t> t, int stage) const {
        const uint64_t s = x + y, d = x + (q16 << stage) - y;
        x = s;
        y = mul_lazy(d, t.w, t.wp);
    }
    HD void bf_inv_last(uint64_t& x, uint64_t& y, TwPair<uint64_t> ninv, TwPair<uint64_t> wninv, int stage) const {
        const uint64_t s = x + y, d = x + (q16 << stage) - y;
        x = mul_lazy(s, ninv.w, ninv.wp);
        y = mul_lazy(d, wninv.w, wninv.wp);
    }
    HD uint64_t inv_pass_fix(uint64_t x) const { return barrett2(x); }
    HD uint64_t inv_canon(uint64_t x) const {  // [0,4q) -> [0,q)
        uint64_t r = umin_(x, x - q2);
        return umin_(r, r - q);
    }
};

// host-side constructors (setup only; shared by the launcher and tests/hostsim)
inline Lz32 make_lz32(uint64_t q) {
    Lz32 m;
    m.q = (uint32_t)q;
    m.q2 = (uint32_t)(2 * q);
    m.q8 = (uint32_t)(8 * q);
    m.mu = (uint32_t)((1ull << 32) / q);
    return m;
}
inline Lz64 make_lz64(uint64_t q) {
    Lz64 m;
    m.q = q;
    m.nq = 0 - q;
    m.q2 = 2 * q;
    m.q4 = 4 * q;
    m.q16 = 16 * q;
    int k = 0;
    while ((q >> k) != 0) ++k;  // q < 2^k
    m.sh_e = (uint32_t)(k > 24 ? k - 24 : 0);
    m.sh_f = (uint32_t)(k - 1) - m.sh_e;
    m.mu32 = (uint32_t)((((unsigned __int128)1) << (32 + m.sh_f + m.sh_e)) / q);
    m.pad_ = 0;
    return m;
}

// per-limb descriptor (device array; limb = polynomial index % limbs)
template <typename L>
struct FastLimb {
    L m;
    const TwPair<typename L::W>* tw;   // forward table, bit-reversed order
    const TwPair<typename L::W>* itw;  // inverse table
    TwPair<typename L::W> ninv, wninv; // n^-1 and itw[1] * n^-1 for the ring degree of this launch
};

// read-only (ld.global.nc) twiddle pair fetch
HD TwPair<uint32_t> ld_tw(const TwPair<uint32_t>* p) {
#if defined(__CUDA_ARCH__)
    const uint2 v = __ldg(reinterpret_cast<const uint2*>(p));
    return TwPair<uint32_t>{v.x, v.y};
#else
    return *p;
#endif
}
HD TwPair<uint64_t> ld_tw(const TwPair<uint64_t>* p) {
#if defined(__CUDA_ARCH__)
    const ulonglong2 v = __ldg(reinterpret_cast<const ulonglong2*>(p));
    return TwPair<uint64_t>{v.x, v.y};
#else
    return *p;
#endif
}

// ---- register passes ---------------------------------------------------------------------------------------------------
// element j of a group sits at tile position (hi << (L+R)) | (j << L) | lo; the pass covers global stages
// l0 .. l0+R-1; tb = 2^l0 + (position >> (logN - l0)); twiddle of sub-stage u, pair prefix `top`: tw[(tb << u) + top].
// NC: twiddles are in global memory and fetched through the read-only path; false: plain loads (shared-memory tables)
template <typename L, int R, bool NC = true>
HD void fast_fwd_regs(const L& m, typename L::W* x, const TwPair<typename L::W>* __restrict__ tw, uint32_t tb) {
#pragma unroll
    for (int u = 0; u < R; ++u) {
        const int h = 1 << (R - 1 - u);
#pragma unroll
        for (int top = 0; top < (1 << u); ++top) {
            const TwPair<typename L::W> t = NC ? ld_tw(tw + ((tb << u) + top)) : tw[(tb << u) + top];
#pragma unroll
            for (int low = 0; low < h; ++low) {
                const int j = (top << (R - u)) | low;
                m.bf_fwd(x[j], x[j + h], t);
            }
        }
    }
}
// LAST: global stage 0 is part of this pass and is the final stage of the transform: fold n^-1.
template <typename L, int R, bool LAST, bool NC = true>
HD void fast_inv_regs(const L& m, typename L::W* x, const TwPair<typename L::W>* __restrict__ itw, uint32_t tb,
                      TwPair<typename L::W> ninv, TwPair<typename L::W> wninv) {
#pragma unroll
    for (int u = R - 1; u >= 0; --u) {
        const int h = 1 << (R - 1 - u);
        const int stage = R - 1 - u;
#pragma unroll
        for (int top = 0; top < (1 << u); ++top) {
            if (LAST && u == 0) {
#pragma unroll
                for (int low = 0; low < h; ++low) m.bf_inv_last(x[low], x[low + h], ninv, wninv, stage);
            } else {
                const TwPair<typename L::W> t = NC ? ld_tw(itw + ((tb << u) + top)) : itw[(tb << u) + top];
#pragma unroll
                for (int low = 0; low < h; ++low) {
                    const int j = (top << (R - u)) | low;
                    m.bf_inv(x[j], x[j + h], t, stage);
                }
            }
        }
    }
}

template <typename L, int R, bool LAST, bool NC = true>
HD void fast_inv_regs_flat(const L& m, typename L::W* x, const TwPair<typename L::W>* __restrict__ itw, uint32_t tb,
                      TwPair<typename L::W> ninv, TwPair<typename L::W> wninv) {
    // fast_inv_regs with every stage as one flat loop over its 2^(R-1) butterflies (constant trip count: nvcc left the nested `low < h`
    // form partly rolled for R = 4 inside the tile kernel, which put x[] in local memory); butterfly b of stage `stage`: top = b >> stage.
    // Only the radix-16 tile plan (FAST_R16_64) uses it: the FHEW kernel measured 1 % slower with this form.
#pragma unroll
    for (int stage = 0; stage < R; ++stage) {
        const int u = R - 1 - stage;
        const int h = 1 << stage;
        TwPair<typename L::W> t[1 << (R - 1)];
        if (!(LAST && u == 0)) {
#pragma unroll
            for (int top = 0; top < (1 << (R - 1)); ++top)
                if (top < (1 << u)) t[top] = NC ? ld_tw(itw + ((tb << u) + top)) : itw[(tb << u) + top];
        }
#pragma unroll
        for (int b = 0; b < (1 << (R - 1)); ++b) {
            const int top = b >> stage, low = b & (h - 1);
            const int j = (top << (R - u)) | low;
            if (LAST && u == 0)
                m.bf_inv_last(x[j], x[j + h], ninv, wninv, stage);
            else
                m.bf_inv(x[j], x[j + h], t[top], stage);
        }
    }
}

// Two groups that share every twiddle (adjacent positions lo, lo + 1 of the same pass): each twiddle is fetched once.
template <typename L, int R, bool NC = true>
HD void fast_fwd_regs2(const L& m, typename L::W* x, typename L::W* y, const TwPair<typename L::W>* __restrict__ tw, uint32_t tb) {
#pragma unroll
    for (int u = 0; u < R; ++u) {
        const int h = 1 << (R - 1 - u);
#pragma unroll
        for (int top = 0; top < (1 << u); ++top) {
            const TwPair<typename L::W> t = NC ? ld_tw(tw + ((tb << u) + top)) : tw[(tb << u) + top];
#pragma unroll
            for (int low = 0; low < h; ++low) {
                const int j = (top << (R - u)) | low;
                m.bf_fwd(x[j], x[j + h], t);
                m.bf_fwd(y[j], y[j + h], t);
            }
        }
    }
}
template <typename L, int R, bool LAST, bool NC = true>
HD void fast_inv_regs2(const L& m, typename L::W* x, typename L::W* y, const TwPair<typename L::W>* __restrict__ itw, uint32_t tb,
                       TwPair<typename L::W> ninv, TwPair<typename L::W> wninv) {
#pragma unroll
    for (int u = R - 1; u >= 0; --u) {
        const int h = 1 << (R - 1 - u);
        const int stage = R - 1 - u;
#pragma unroll
        for (int top = 0; top < (1 << u); ++top) {
            if (LAST && u == 0) {
#pragma unroll
                for (int low = 0; low < h; ++low) {
                    m.bf_inv_last(x[low], x[low + h], ninv, wninv, stage);
                    m.bf_inv_last(y[low], y[low + h], ninv, wninv, stage);
                }
            } else {
                const TwPair<typename L::W> t = NC ? ld_tw(itw + ((tb << u) + top)) : itw[(tb << u) + top];
#pragma unroll
                for (int low = 0; low < h; ++low) {
                    const int j = (top << (R - u)) | low;
                    m.bf_inv(x[j], x[j + h], t, stage);
                    m.bf_inv(y[j], y[j + h], t, stage);
                }
            }
        }
    }
}
// 8-byte access to two adjacent 32-bit words (even word address)
HD void ld_pair(const uint32_t* p, uint32_t& a, uint32_t& b) {
#if defined(__CUDA_ARCH__)
    const uint2 v = *reinterpret_cast<const uint2*>(p);
    a = v.x;
    b = v.y;
#else
    a = p[0];
    b = p[1];
#endif
}
HD void st_pair(uint32_t* p, uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint2*>(p) = make_uint2(a, b);
#else
    p[0] = a;
    p[1] = b;
#endif
}

// ---- 16-byte vector helpers (8 logically contiguous words at swizzled word address P0, P0 = swz(8g)) ----------------
#if defined(__CUDA_ARCH__)
DEV void ld_vec8(const uint32_t* s, uint32_t P0, uint32_t* x) {
    const uint4 a = *reinterpret_cast<const uint4*>(s + P0);
    const uint4 b = *reinterpret_cast<const uint4*>(s + (P0 ^ 4u));
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
DEV void st_vec8(uint32_t* s, uint32_t P0, const uint32_t* x) {
    *reinterpret_cast<uint4*>(s + P0) = make_uint4(x[0], x[1], x[2], x[3]);
    *reinterpret_cast<uint4*>(s + (P0 ^ 4u)) = make_uint4(x[4], x[5], x[6], x[7]);
}
DEV void ld_vec8(const uint64_t* s, uint32_t P0, uint64_t* x) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(s + (P0 ^ (2u * m)));
        x[2 * m] = a.x;
        x[2 * m + 1] = a.y;
    }
}
DEV void st_vec8(uint64_t* s, uint32_t P0, const uint64_t* x) {
#pragma unroll
    for (int m = 0; m < 4; ++m) *reinterpret_cast<ulonglong2*>(s + (P0 ^ (2u * m))) = make_ulonglong2(x[2 * m], x[2 * m + 1]);
}
// 16 consecutive 64-bit words of a tile (the swizzle of a 16-aligned position only touches address bits 1-3)
DEV void ld_vec16(const uint64_t* s, uint32_t P0, uint64_t* x) {
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const ulonglong2 a = *reinterpret_cast<const ulonglong2*>(s + (P0 ^ (2u * m)));
        x[2 * m] = a.x;
        x[2 * m + 1] = a.y;
    }
}
DEV void st_vec16(uint64_t* s, uint32_t P0, const uint64_t* x) {
#pragma unroll
    for (int m = 0; m < 8; ++m) *reinterpret_cast<ulonglong2*>(s + (P0 ^ (2u * m))) = make_ulonglong2(x[2 * m], x[2 * m + 1]);
}
DEV void ldg_vec16(const uint64_t* g, uint64_t* x) {
#pragma unroll
    for (int m = 0; m < 8; ++m) {
        const ulonglong2 a = reinterpret_cast<const ulonglong2*>(g)[m];
        x[2 * m] = a.x;
        x[2 * m + 1] = a.y;
    }
}
DEV void stg_vec16(uint64_t* g, const uint64_t* x) {
#pragma unroll
    for (int m = 0; m < 8; ++m) reinterpret_cast<ulonglong2*>(g)[m] = make_ulonglong2(x[2 * m], x[2 * m + 1]);
}
DEV void ld_vec16(const uint32_t*, uint32_t, uint32_t*) {}  // radix-16 tile passes exist for 64-bit words only
DEV void st_vec16(uint32_t*, uint32_t, const uint32_t*) {}
DEV void ldg_vec16(const uint32_t*, uint32_t*) {}
DEV void stg_vec16(uint32_t*, const uint32_t*) {}
// 8 contiguous words in global memory (16-byte aligned)
DEV void ldg_vec8(const uint32_t* g, uint32_t* x) {
    const uint4 a = reinterpret_cast<const uint4*>(g)[0], b = reinterpret_cast<const uint4*>(g)[1];
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
DEV void stg_vec8(uint32_t* g, const uint32_t* x) {
    reinterpret_cast<uint4*>(g)[0] = make_uint4(x[0], x[1], x[2], x[3]);
    reinterpret_cast<uint4*>(g)[1] = make_uint4(x[4], x[5], x[6], x[7]);
}
DEV void ldg_vec8(const uint64_t* g, uint64_t* x) {
#pragma unroll
    for (int m = 0; m < 4; ++m) {
        const ulonglong2 a = reinterpret_cast<const ulonglong2*>(g)[m];
        x[2 * m] = a.x;
        x[2 * m + 1] = a.y;
    }
}
DEV void stg_vec8(uint64_t* g, const uint64_t* x) {
#pragma unroll
    for (int m = 0; m < 4; ++m) reinterpret_cast<ulonglong2*>(g)[m] = make_ulonglong2(x[2 * m], x[2 * m + 1]);
}
#else
template <typename W>
inline void ld_vec8(const W* s, uint32_t P0, W* x) {
    for (uint32_t j = 0; j < 8; ++j) x[j] = s[P0 ^ j];
}
template <typename W>
inline void st_vec8(W* s, uint32_t P0, const W* x) {
    for (uint32_t j = 0; j < 8; ++j) s[P0 ^ j] = x[j];
}
template <typename W>
inline void ldg_vec8(const W* g, W* x) {
    for (int j = 0; j < 8; ++j) x[j] = g[j];
}
template <typename W>
inline void stg_vec8(W* g, const W* x) {
    for (int j = 0; j < 8; ++j) g[j] = x[j];
}
template <typename W>
inline void ld_vec16(const W* s, uint32_t P0, W* x) {
    for (uint32_t j = 0; j < 16; ++j) x[j] = s[P0 ^ j];
}
template <typename W>
inline void st_vec16(W* s, uint32_t P0, const W* x) {
    for (uint32_t j = 0; j < 16; ++j) s[P0 ^ j] = x[j];
}
template <typename W>
inline void ldg_vec16(const W* g, W* x) {
    for (int j = 0; j < 16; ++j) x[j] = g[j];
}
template <typename W>
inline void stg_vec16(W* g, const W* x) {
    for (int j = 0; j < 16; ++j) g[j] = x[j];
}
#endif

// ---- tile passes ---------------------------------------------------------------------------------------------------------
// A tile is 2^LOGT contiguous coefficients = chunk k (of 2^s0) of a polynomial of degree 2^(s0+LOGT).  The plan is a
// first pass of R1 stages (2 <= R1 <= 4) followed by (LOGT-R1)/3 radix-8 passes; the inverse runs it backwards.
// `tid` enumerates the TPP threads that co-operate on one tile.

// forward first pass: global -> registers -> shared
template <typename L, int LOGT, int R1, int TPP>
HD void fast_fwd_first(const FastLimb<L>& d, const typename L::W* g, typename L::W* s, int s0, uint32_t k, bool pre_red,
                       uint32_t tid) {
    typedef typename L::W W;
    constexpr int LL = LOGT - R1;
    const uint32_t tb = (1u << s0) + k;
#pragma unroll 1
    for (uint32_t grp = tid; grp < (1u << LL); grp += TPP) {
        W x[1 << R1];
#pragma unroll
        for (int j = 0; j < (1 << R1); ++j) x[j] = g[grp + ((uint32_t)j << LL)];
        if (pre_red) {
#pragma unroll
            for (int j = 0; j < (1 << R1); ++j) x[j] = d.m.pre_red(x[j]);
        }
        fast_fwd_regs<L, R1>(d.m, x, d.tw, tb);
        const uint32_t P0 = swz2<W>(grp);
#pragma unroll
        for (int j = 0; j < (1 << R1); ++j) s[P0 ^ swz2<W>((uint32_t)j << LL)] = x[j];
    }
}
// paired first pass (32-bit words, R1 <= 3): groups grp, grp + 1 -> 8-byte global loads and shared stores, shared twiddles
template <typename L, int LOGT, int R1, int TPP>
HD void fast_fwd_first2(const FastLimb<L>& d, const uint32_t* g, uint32_t* s, int s0, uint32_t k, bool pre_red, uint32_t tid) {
    typedef uint32_t W;
    constexpr int LL = LOGT - R1;
    const uint32_t tb = (1u << s0) + k;
#pragma unroll 1
    for (uint32_t grp = tid << 1; grp < (1u << LL); grp += 2 * TPP) {
        W x[1 << R1], y[1 << R1];
#pragma unroll
        for (int j = 0; j < (1 << R1); ++j) ld_pair(g + grp + ((uint32_t)j << LL), x[j], y[j]);
        if (pre_red) {
#pragma unroll
            for (int j = 0; j < (1 << R1); ++j) {
                x[j] = d.m.pre_red(x[j]);
                y[j] = d.m.pre_red(y[j]);
            }
        }
        fast_fwd_regs2<L, R1>(d.m, x, y, d.tw, tb);
        const uint32_t P0 = swz2<W>(grp);
#pragma unroll
        for (int j = 0; j < (1 << R1); ++j) st_pair(s + (P0 ^ swz2<W>((uint32_t)j << LL)), x[j], y[j]);
    }
}
// forward middle pass (radix 2^RM, shared -> shared); local stages t0 .. t0+RM-1, LL = LOGT - t0 - RM >= 3
template <typename L, int LOGT, int TPP, int t0, int RM = 3>
HD void fast_fwd_mid(const FastLimb<L>& d, typename L::W* s, int s0, uint32_t k, bool pre_red, uint32_t tid) {
    typedef typename L::W W;
    constexpr int LL = LOGT - t0 - RM, NE = 1 << RM;
#pragma unroll 1
    for (uint32_t grp = tid; grp < (1u << (LOGT - RM)); grp += TPP) {
        const uint32_t lo = grp & ((1u << LL) - 1u), hi = grp >> LL;
        const uint32_t P0 = swz2<W>((hi << (LL + RM)) | lo);
        W x[NE];
#pragma unroll
        for (int j = 0; j < NE; ++j) x[j] = s[P0 ^ swz2<W>((uint32_t)j << LL)];
        if (pre_red) {
#pragma unroll
            for (int j = 0; j < NE; ++j) x[j] = d.m.pre_red(x[j]);
        }
        const uint32_t tb = (1u << (s0 + t0)) + (k << t0) + hi;
        fast_fwd_regs<L, RM>(d.m, x, d.tw, tb);
#pragma unroll
        for (int j = 0; j < NE; ++j) s[P0 ^ swz2<W>((uint32_t)j << LL)] = x[j];
    }
}
// forward middle pass for 32-bit words, two adjacent groups (lo even, lo + 1) per iteration: the XOR swizzle leaves word-address
// bits 0..1 alone, so the pair is one aligned 8-byte shared access, and both groups use the same twiddles.  Half as many
// shared-memory wavefronts, twiddle fetches and address computations per butterfly as fast_fwd_mid.
template <typename L, int LOGT, int TPP, int t0>
HD void fast_fwd_mid2(const FastLimb<L>& d, uint32_t* s, int s0, uint32_t k, bool pre_red, uint32_t tid) {
    typedef uint32_t W;
    constexpr int LL = LOGT - t0 - 3;
    static_assert(LL >= 1, "paired groups need a stride of at least two words");
#pragma unroll 1
    for (uint32_t gp = tid; gp < (1u << (LOGT - 4)); gp += TPP) {
        const uint32_t lo = (gp & ((1u << (LL - 1)) - 1u)) << 1, hi = gp >> (LL - 1);
        const uint32_t P0 = swz2<W>((hi << (LL + 3)) | lo);
        W x[8], y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ld_pair(s + (P0 ^ swz2<W>((uint32_t)j << LL)), x[j], y[j]);
        if (pre_red) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                x[j] = d.m.pre_red(x[j]);
                y[j] = d.m.pre_red(y[j]);
            }
        }
        const uint32_t tb = (1u << (s0 + t0)) + (k << t0) + hi;
        fast_fwd_regs2<L, 3>(d.m, x, y, d.tw, tb);
#pragma unroll
        for (int j = 0; j < 8; ++j) st_pair(s + (P0 ^ swz2<W>((uint32_t)j << LL)), x[j], y[j]);
    }
}
// forward last pass (radix 2^RM, LL = 0): shared -> registers -> global, canonical output
template <typename L, int LOGT, int TPP, int RM = 3>
HD void fast_fwd_last(const FastLimb<L>& d, const typename L::W* s, typename L::W* __restrict__ g, int s0, uint32_t k, bool pre_red,
                      uint32_t tid) {
    typedef typename L::W W;
    constexpr int t0 = LOGT - RM, NE = 1 << RM;
#pragma unroll 1
    for (uint32_t grp = tid; grp < (1u << (LOGT - RM)); grp += TPP) {
        W x[NE];
        if (RM == 4)
            ld_vec16(s, swz2<W>(grp << RM), x);
        else
            ld_vec8(s, swz2<W>(grp << RM), x);
        if (pre_red) {
#pragma unroll
            for (int j = 0; j < NE; ++j) x[j] = d.m.pre_red(x[j]);
        }
        const uint32_t tb = (1u << (s0 + t0)) + (k << t0) + grp;
        fast_fwd_regs<L, RM>(d.m, x, d.tw, tb);
#pragma unroll
        for (int j = 0; j < NE; ++j) x[j] = d.m.canon(x[j]);
        if (RM == 4)
            stg_vec16(g + (grp << RM), x);
        else
            stg_vec8(g + (grp << RM), x);
    }
}
// inverse first pass (radix 2^RM, LL = 0): global -> registers -> shared
template <typename L, int LOGT, int TPP, int RM = 3>
HD void fast_inv_first(const FastLimb<L>& d, const typename L::W* g, typename L::W* s, int s0, uint32_t k, uint32_t tid) {
    typedef typename L::W W;
    constexpr int t0 = LOGT - RM, NE = 1 << RM;
#pragma unroll 1
    for (uint32_t grp = tid; grp < (1u << (LOGT - RM)); grp += TPP) {
        W x[NE];
        if (RM == 4)
            ldg_vec16(g + (grp << RM), x);
        else
            ldg_vec8(g + (grp << RM), x);
        const uint32_t tb = (1u << (s0 + t0)) + (k << t0) + grp;
        if (RM == 4)
            fast_inv_regs_flat<L, RM, false>(d.m, x, d.itw, tb, d.ninv, d.wninv);
        else
            fast_inv_regs<L, RM, false>(d.m, x, d.itw, tb, d.ninv, d.wninv);
        x[0] = d.m.inv_pass_fix(x[0]);
        if (RM == 4) x[1] = d.m.inv_pass_fix(x[1]);  // the sum chain of a 4-stage pass reaches 32q on element 1 (see Lz64)
        if (RM == 4)
            st_vec16(s, swz2<W>(grp << RM), x);
        else
            st_vec8(s, swz2<W>(grp << RM), x);
    }
}
template <typename L, int LOGT, int TPP, int t0, int RM = 3>
HD void fast_inv_mid(const FastLimb<L>& d, typename L::W* s, int s0, uint32_t k, uint32_t tid) {
    typedef typename L::W W;
    constexpr int LL = LOGT - t0 - RM, NE = 1 << RM;
#pragma unroll 1
    for (uint32_t grp = tid; grp < (1u << (LOGT - RM)); grp += TPP) {
        const uint32_t lo = grp & ((1u << LL) - 1u), hi = grp >> LL;
        const uint32_t P0 = swz2<W>((hi << (LL + RM)) | lo);
        W x[NE];
#pragma unroll
        for (int j = 0; j < NE; ++j) x[j] = s[P0 ^ swz2<W>((uint32_t)j << LL)];
        const uint32_t tb = (1u << (s0 + t0)) + (k << t0) + hi;
        if (RM == 4)
            fast_inv_regs_flat<L, RM, false>(d.m, x, d.itw, tb, d.ninv, d.wninv);
        else
            fast_inv_regs<L, RM, false>(d.m, x, d.itw, tb, d.ninv, d.wninv);
        x[0] = d.m.inv_pass_fix(x[0]);
        if (RM == 4) x[1] = d.m.inv_pass_fix(x[1]);
#pragma unroll
        for (int j = 0; j < NE; ++j) s[P0 ^ swz2<W>((uint32_t)j << LL)] = x[j];
    }
}
template <typename L, int LOGT, int TPP, int t0>
HD void fast_inv_mid2(const FastLimb<L>& d, uint32_t* s, int s0, uint32_t k, uint32_t tid) {  // see fast_fwd_mid2
    typedef uint32_t W;
    constexpr int LL = LOGT - t0 - 3;
    static_assert(LL >= 1, "paired groups need a stride of at least two words");
#pragma unroll 1
    for (uint32_t gp = tid; gp < (1u << (LOGT - 4)); gp += TPP) {
        const uint32_t lo = (gp & ((1u << (LL - 1)) - 1u)) << 1, hi = gp >> (LL - 1);
        const uint32_t P0 = swz2<W>((hi << (LL + 3)) | lo);
        W x[8], y[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) ld_pair(s + (P0 ^ swz2<W>((uint32_t)j << LL)), x[j], y[j]);
        const uint32_t tb = (1u << (s0 + t0)) + (k << t0) + hi;
        fast_inv_regs2<L, 3, false>(d.m, x, y, d.itw, tb, d.ninv, d.wninv);
        x[0] = d.m.inv_pass_fix(x[0]);
        y[0] = d.m.inv_pass_fix(y[0]);
#pragma unroll
        for (int j = 0; j < 8; ++j) st_pair(s + (P0 ^ swz2<W>((uint32_t)j << LL)), x[j], y[j]);
    }
}
// inverse last pass (R1 stages at local stages 0..R1-1): shared -> registers -> global.
// FINAL: s0 == 0, this is the end of the transform: fold n^-1 and write canonical residues; otherwise the column
// kernel follows and the pass invariant (u32 [0,2q), u64 < 16q) is restored instead.
template <typename L, int LOGT, int R1, int TPP, bool FINAL>
HD void fast_inv_last(const FastLimb<L>& d, const typename L::W* s, typename L::W* __restrict__ g, int s0, uint32_t k, uint32_t tid) {
    typedef typename L::W W;
    constexpr int LL = LOGT - R1;
    const uint32_t tb = (1u << s0) + k;
#pragma unroll 1
    for (uint32_t grp = tid; grp < (1u << LL); grp += TPP) {
        const uint32_t P0 = swz2<W>(grp);
        W x[1 << R1];
#pragma unroll
        for (int j = 0; j < (1 << R1); ++j) x[j] = s[P0 ^ swz2<W>((uint32_t)j << LL)];
        fast_inv_regs<L, R1, FINAL>(d.m, x, d.itw, tb, d.ninv, d.wninv);
        if (FINAL) {
#pragma unroll
            for (int j = 0; j < (1 << R1); ++j) x[j] = d.m.inv_canon(x[j]);
        } else {
            x[0] = d.m.inv_pass_fix(x[0]);
            if (R1 == 4) x[1] = d.m.inv_pass_fix(x[1]);
        }
#pragma unroll
        for (int j = 0; j < (1 << R1); ++j) g[grp + ((uint32_t)j << LL)] = x[j];
    }
}

template <typename L, int LOGT, int R1, int TPP, bool FINAL>
HD void fast_inv_last2(const FastLimb<L>& d, const uint32_t* s, uint32_t* __restrict__ g, int s0, uint32_t k, uint32_t tid) {  // see fast_fwd_first2
    typedef uint32_t W;
    constexpr int LL = LOGT - R1;
    const uint32_t tb = (1u << s0) + k;
#pragma unroll 1
    for (uint32_t grp = tid << 1; grp < (1u << LL); grp += 2 * TPP) {
        const uint32_t P0 = swz2<W>(grp);
        W x[1 << R1], y[1 << R1];
#pragma unroll
        for (int j = 0; j < (1 << R1); ++j) ld_pair(s + (P0 ^ swz2<W>((uint32_t)j << LL)), x[j], y[j]);
        fast_inv_regs2<L, R1, FINAL>(d.m, x, y, d.itw, tb, d.ninv, d.wninv);
        if (FINAL) {
#pragma unroll
            for (int j = 0; j < (1 << R1); ++j) {
                x[j] = d.m.inv_canon(x[j]);
                y[j] = d.m.inv_canon(y[j]);
            }
        } else {
            x[0] = d.m.inv_pass_fix(x[0]);
            y[0] = d.m.inv_pass_fix(y[0]);
        }
#pragma unroll
        for (int j = 0; j < (1 << R1); ++j) st_pair(g + grp + ((uint32_t)j << LL), x[j], y[j]);
    }
}

// first-pass radix of a 2^logt tile (the remaining stages are radix-8 passes)
HD constexpr int fast_r1(int logt) { return logt % 3 == 0 ? 3 : (logt % 3 == 1 ? 4 : 2); }
// ---- tile geometry and pass sequence (shared by the kernel and tests/hostsim) -----------------------------------------------
// PAIR (32-bit words only): middle passes process two adjacent groups per thread, so half as many threads cover a tile.
// FAST_R16_64 = 1: the 2^12 tile of 64-bit words runs three radix-16 passes (4 + 4 + 4) instead of four radix-8 ones: two instead
// of three round trips through shared memory, one group per thread and pass.  Measured (4096 x 2^12 / 2^16, B200): forward 1.82-1.84
// vs 1.84 TB/s / 1.48-1.50 vs 1.50, inverse 1.59-1.63 vs 1.84 / 1.31-1.32 vs 1.49 at 80 or 64 registers: the tile is bound by the
// multiply pipe, not by the shared-memory round trips.  Off.
#ifndef FAST_R16_64
#define FAST_R16_64 0
#endif
template <typename L, int LOGT>
struct FastGeom {
    static constexpr int RM = (L::BITS == 64 && FAST_R16_64 && LOGT == 12) ? 4 : 3;  // radix of the passes after the first
    static constexpr int R1 = RM == 4 ? 4 : fast_r1(LOGT);
    static constexpr int RMAX = R1 > RM ? R1 : RM;
    static constexpr int PAIR = (L::BITS == 32 && FAST_PAIR32) ? 1 : 0;
    // 64-bit words: no pairing (twice the registers), but for the big tiles half the threads per tile, each walking two groups
    // per pass: 256-thread CTAs (4 per SM) interleave their load / compute / store phases better than 512-thread ones
    // (measured +6-13 % at 2^12, the tile of N = 2^12 and 2^16; smaller tiles already run 256-thread CTAs)
    static constexpr int HALF = (RM == 3 && ((L::BITS == 64 && FAST_HALF64 && LOGT >= 12) || (L::BITS == 32 && FAST_HALF32 && LOGT >= 12))) ? 1 : 0;
    static constexpr int TPP = 1 << (LOGT - RMAX - PAIR - HALF);
    static constexpr int PB = TPP >= FAST_MIN_NTHR ? 1 : FAST_MIN_NTHR / TPP;
    static constexpr int NTHR = TPP * PB;
    static constexpr int NP3 = (LOGT - R1) / RM;  // passes after the first (middle passes + the last one)
};
template <typename L, int LOGT, int t0>
HD void fast_fwd_mid_any(const FastLimb<L>& d, typename L::W* s, int s0, uint32_t k, bool pre_red, uint32_t tid) {
    typedef FastGeom<L, LOGT> G;
    if constexpr (G::PAIR != 0)
        fast_fwd_mid2<L, LOGT, G::TPP, t0>(d, s, s0, k, pre_red, tid);
    else
        fast_fwd_mid<L, LOGT, G::TPP, t0, G::RM>(d, s, s0, k, pre_red, tid);
}
template <typename L, int LOGT, int t0>
HD void fast_inv_mid_any(const FastLimb<L>& d, typename L::W* s, int s0, uint32_t k, uint32_t tid) {
    typedef FastGeom<L, LOGT> G;
    if constexpr (G::PAIR != 0)
        fast_inv_mid2<L, LOGT, G::TPP, t0>(d, s, s0, k, tid);
    else
        fast_inv_mid<L, LOGT, G::TPP, t0, G::RM>(d, s, s0, k, tid);
}
template <typename L, int LOGT>
HD void fast_fwd_first_any(const FastLimb<L>& d, const typename L::W* g, typename L::W* s, int s0, uint32_t k, bool pre_red, uint32_t tid) {
    typedef FastGeom<L, LOGT> G;
    if constexpr (G::PAIR != 0 && G::R1 <= 3)
        fast_fwd_first2<L, LOGT, G::R1, G::TPP>(d, g, s, s0, k, pre_red, tid);
    else
        fast_fwd_first<L, LOGT, G::R1, G::TPP>(d, g, s, s0, k, pre_red, tid);
}
template <typename L, int LOGT, bool FINAL>
HD void fast_inv_last_any(const FastLimb<L>& d, const typename L::W* s, typename L::W* __restrict__ g, int s0, uint32_t k, uint32_t tid) {
    typedef FastGeom<L, LOGT> G;
    if constexpr (G::PAIR != 0 && G::R1 <= 3)
        fast_inv_last2<L, LOGT, G::R1, G::TPP, FINAL>(d, s, g, s0, k, tid);
    else
        fast_inv_last<L, LOGT, G::R1, G::TPP, FINAL>(d, s, g, s0, k, tid);
}
// tile size used for a ring of degree 2^log_n: rings larger than one tile take a column kernel (S <= 4 stages, HBM-bound)
// first; measured best tile 2^11 for N = 2^14, 2^15
HD constexpr int fast_tile_logt(int log_n) { return log_n <= 13 ? log_n : (log_n <= 15 ? 11 : log_n - 4); }

// ---- column passes (first S forward stages / last S inverse stages of a polynomial too large for one tile) ------------
// the polynomial is viewed as a [2^S][2^lc] row-major matrix; one thread transforms one column in registers.
template <typename L, int S>
HD void fast_fwd_column(const FastLimb<L>& d, const typename L::W* gin, typename L::W* g, int lc, uint32_t col) {
    typedef typename L::W W;
    W x[1 << S];
#pragma unroll
    for (int j = 0; j < (1 << S); ++j) x[j] = gin[col + ((size_t)j << lc)];
    fast_fwd_regs<L, S>(d.m, x, d.tw, 1u);
#pragma unroll
    for (int j = 0; j < (1 << S); ++j) g[col + ((size_t)j << lc)] = x[j];
}
template <typename L, int S>
HD void fast_inv_column(const FastLimb<L>& d, const typename L::W* gin, typename L::W* g, int lc, uint32_t col) {
    typedef typename L::W W;
    W x[1 << S];
#pragma unroll
    for (int j = 0; j < (1 << S); ++j) x[j] = gin[col + ((size_t)j << lc)];
    fast_inv_regs<L, S, true>(d.m, x, d.itw, 1u, d.ninv, d.wninv);
#pragma unroll
    for (int j = 0; j < (1 << S); ++j) g[col + ((size_t)j << lc)] = d.m.inv_canon(x[j]);
}

// tile plan: first-pass radix for a tile of LOGT stages

// Bounds bookkeeping for the u32 forward path (host side, planner): returns the pre_red bit mask for the passes of the
// tile (bit 0 = first pass, bit i = i-th radix-8 pass after it) given the bound (in units of q) of the values entering
// the tile, or -1 if impossible.
inline int fast_plan_prered32(int logt, int bound_in) {
    const int r1 = fast_r1(logt);
    if (bound_in > 16) return -1;
    int mask = 0, b = bound_in;
    if (b + 2 * r1 > 16) {
        mask |= 1;
        b = 8;
    }
    b += 2 * r1;
    const int np = (logt - r1) / 3;
    for (int i = 0; i < np; ++i) {
        if (b + 6 > 16) {
            mask |= 2 << i;
            b = 8;
        }
        b += 6;
    }
    return mask;
}

}  // namespace fhe
