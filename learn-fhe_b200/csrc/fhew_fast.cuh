// FHEW / LMKCDEY blind rotation, fast path for N = 512, Q < 2^28 (the reference's single-key parameter set,
// scheme/fhew/src/fhew/boolean.rs:225-239).  Same reference call sites and the same exact arithmetic as fhew_core.cuh
// (rgsw.rs:116-128, rlwe.rs:177-202, bootstrapping.rs:158-231); what changes is the schedule inside one step:
//
//   128 threads per ciphertext.  Forward transforms are split 3 + 4 + 2 stages:
//     P1  thread (g, h): reads acc_h[g + 64 j] (or the permuted a(X^t)), decomposes in registers and runs the first
//         radix-8 pass of its digit polynomials before anything is written: digits never exist untransformed in memory;
//     P2  radix-16 pass in shared memory;
//     P3  thread t owns evaluation points 4t..4t+3 of EVERY digit polynomial: last radix-4 pass, multiply-accumulate against
//         the key rows (two coalesced 16-byte loads per row), Barrett, and the first radix-4 pass of BOTH inverse transforms,
//         all in registers: no exchange of partial sums, no separate MAC pass;
//     P4  inverse radix-16 pass (2 polynomials); P5 inverse radix-8 pass with n^-1 folded in, (+ b(X^t)), canonical result
//         written to the other accumulator buffer (ping-pong, so permuted reads of the old accumulator stay valid).
//   5 barriers per step (the first generation needs 9) and about a third of its instructions: lazy forward butterflies
//   (digits enter as digit + Q < 2 Q, values stay < 16 Q < 2^32, one min() before P3), twiddles in shared memory, swizzled addresses formed as swz(base) ^ const_j.
// __host__ __device__ so tests/hostsim can replay it.
#pragma once
#include "fhew_core.cuh"
#include "ntt_fast.cuh"

#ifndef FF_P3_UNROLL
#define FF_P3_UNROLL 4  // rows of the P3 MAC loop unrolled together
#endif
#define FF_PRAGMA_(x) _Pragma(#x)
#define FF_UNROLL(n) FF_PRAGMA_(unroll n)

namespace fhe {

static constexpr int FF_LOGN = 9;
static constexpr int FF_N = 1 << FF_LOGN;
static constexpr int FF_THREADS = 128;

struct FhewFastDev {
    Lz32 m;
    uint64_t mu64;  // floor(2^64 / Q): Barrett constant for the 64-bit MAC accumulators
    uint32_t n_s, w;
    DecompParam g_dec;  // RGSW decomposor (log_b * d <= 32, d <= 4)
    DecompParam r_dec;  // RLWE key-switch decomposor
    const uint4* brk4;  // [n_s][2 d_g rows][128][2]: {a(4t..4t+3)}, {b(4t..4t+3)} evaluation form
    const uint4* ak4;   // [w+1][d_r rows][128][2]
    const uint16_t* dlog;
    const TwPair<uint32_t>* tw;   // global tables (copied to shared memory when the kernel starts)
    const TwPair<uint32_t>* itw;
    TwPair<uint32_t> ninv, wninv;
    uint32_t ak_t[40];     // automorphism exponents t mod 2N for ak[0..w]
    uint32_t ak_tinv[40];  // t^-1 mod 2N
};

// shared memory (32-bit words): acc[2 (a,b)][N] | dig[8][N] | tw[N pairs] | itw[N pairs] | steps | a2n
// (during an automorphism step dig[7] parks b(X^t): its digit polynomials occupy dig[0..d_r), d_r <= 4)
struct FhewFastSmem {
    uint32_t* acc;
    uint32_t* dig;
    const TwPair<uint32_t>* tw;
    const TwPair<uint32_t>* itw;
};
// FF_TW_NC = 1: twiddle tables stay in global memory and are read through the read-only L1 path (8 KB less shared memory
// per CTA); 0: copied into shared memory by every CTA
#ifndef FF_TW_NC
#define FF_TW_NC 0
#endif
// digit slot that receives the b half of a step's result (the a half goes to slot 0); see ff_step
static constexpr uint32_t FF_RB = 4;
HD constexpr size_t ff_fixed_words() { return (size_t)2 * FF_N + 8 * FF_N + (FF_TW_NC ? 0 : 4 * FF_N); }
HD TwPair<uint32_t> ff_tw(const TwPair<uint32_t>* p) { return FF_TW_NC ? ld_tw(p) : *p; }

// swizzle of the digit polynomials: bits 5,6,7 -> 2,3,4 and bit 8 -> 2 (the radix-16 pass at stride 4 needs bit 8)
HD uint32_t swzf(uint32_t p) { return p ^ ((p >> 3) & 0x1Cu) ^ ((p >> 6) & 4u); }

HD uint32_t ff_reduce64(const FhewFastDev& P, uint64_t x) {  // any x -> [0, Q)
    const uint64_t qh = mulhi_u64(x, P.mu64);
    const uint32_t r = (uint32_t)(x - qh * (uint64_t)P.m.q);  // [0, 2Q)
    return umin_(r, r - P.m.q);
}

// value of poly(X^t) at coefficient c, read from the canonical polynomial `src`: +-src[c * tinv mod 2N]  (avec.rs:34-50)
HD uint32_t ff_perm_coef(const FhewFastDev& P, const uint32_t* src, uint32_t tinv, uint32_t c) {
    const uint32_t i = (c * tinv) & (2 * FF_N - 1);
    const uint32_t v = src[i & (FF_N - 1)];
    return i >= (uint32_t)FF_N ? P.m.q - v - (v == 0 ? P.m.q : 0) : v;  // neg(0) = 0
}

// ---- P1: decompose + first forward radix-8 pass (stages 0..2, stride 64) --------------------------------------------------------
// Signed-digit state of one coefficient (decompose.rs:92-111 with zq.rs:83-89), 32-bit working word (log_b * d <= 32):
// start: x = centred(rounding_shr(v)); step: emits the next digit as a residue mod Q and advances x.
// The centred value is offset by K = B^d (log_b * d <= 30): the d digits only depend on x mod B^d, and with the offset every
// intermediate is a small non-negative number, so the two-operand adds cannot wrap and can be pinned to the ALU pipe.
HD uint32_t ff_dec_start(const FhewFastDev& P, const DecompParam& dp, uint32_t v) {
    uint32_t r = alu_add(v, (uint32_t)dp.half);
    r = umin_(r, r - P.m.q);
    const uint32_t sh = r >> dp.rounding_bits;
    const uint32_t k = 1u << (dp.log_b * dp.d);
    return sh < (P.m.q >> 1) ? sh + k : sh + (k - P.m.q);
}
// For log_b >= 2 the reference's carry rule `limb + (x & 1) > B/2` is `limb > B/2` (B/2 is even, so the tie limb = B/2 has
// an even limb and never carries), hence with h = B/2 - 1: signed digit = ((x + h) & (B-1)) - h, next x = (x + h) >> log_b
// (bits above the log_b * d consumed ones are irrelevant).  The digit is returned as the representative digit + Q in
// (0, 2Q): the forward butterflies are lazy, so no canonical form is needed.
HD uint32_t ff_dec_step(const FhewFastDev& P, const DecompParam& dp, uint32_t& x) {
    const uint32_t mask = (1u << dp.log_b) - 1u, hb = (1u << (dp.log_b - 1)) - 1u;
    const uint32_t t = alu_add(x, hb);
    x = t >> dp.log_b;
    return alu_add(t & mask, P.m.q - hb);
}
// Thread (g, h) holds the 8 coefficients at positions g + 64 j (values from acc_in[h] for an external product, from the
// permuted a(X^t) for an automorphism) and walks the digits; digit k becomes polynomial `pbase + k` if lo <= k < hi.
// where thread g parks coefficient g + 64 j of b(X^t) during an automorphism step: slot 7 at the SAME swizzled address the
// thread itself uses for its digit stores, so that a later P1 writing slot 7 (fused with this step's P5) only ever overwrites
// words its own thread has already consumed
HD uint32_t ff_park(uint32_t g, int j) { return 7u * FF_N + (swzf(g) ^ swzf((uint32_t)j << 6)); }
// digit walk of P1 on decomposition states st[8] (positions g + 64 j): digit k becomes polynomial `pbase + k` if lo <= k < hi
HD void ff_p1_digits(const FhewFastDev& P, const FhewFastSmem& S, const DecompParam& dp, uint32_t* st, uint32_t g, uint32_t lo, uint32_t hi,
                     uint32_t pbase) {
    const uint32_t P0 = swzf(g);
#pragma unroll 1
    for (uint32_t k = 0; k < dp.d; ++k) {
        uint32_t x[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) x[j] = ff_dec_step(P, dp, st[j]);
        if (k >= lo && k < hi) {
            fast_fwd_regs<Lz32, 3, FF_TW_NC != 0>(P.m, x, S.tw, 1u);
            uint32_t* d = S.dig + ((pbase + k) << FF_LOGN);
#pragma unroll
            for (int j = 0; j < 8; ++j) d[P0 ^ swzf((uint32_t)j << 6)] = x[j];
        }
    }
}
HD void ff_p1(const FhewFastDev& P, const FhewFastSmem& S, const DecompParam& dp, const uint32_t* acc_in, bool is_auto, uint32_t tinv,
              uint32_t g, uint32_t h, uint32_t lo, uint32_t hi, uint32_t pbase) {
    uint32_t st[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const uint32_t c = g + 64u * j;
        uint32_t v;
        if (is_auto) {
            const uint32_t i = (c * tinv) & (2 * FF_N - 1);  // a(X^t)[c] = +-a[c * t^-1 mod 2N]   (avec.rs:34-50)
            v = acc_in[i & (FF_N - 1)];
            if (i >= (uint32_t)FF_N) v = v == 0 ? 0 : P.m.q - v;
        } else {
            v = acc_in[(h << FF_LOGN) + c];
        }
        st[j] = ff_dec_start(P, dp, v);
    }
    if (is_auto && h == 1) {  // park b(X^t) for P5: the accumulator is updated in place
#pragma unroll
        for (int j = 0; j < 8; ++j) S.dig[ff_park(g, j)] = ff_perm_coef(P, acc_in + FF_N, tinv, g + 64u * j);
    }
    ff_p1_digits(P, S, dp, st, g, lo, hi, pbase);
}
// ---- P2: forward radix-16 pass (stages 3..6, stride 4) on polynomial `poly`, group `grp` in [0, 32) ------------------------------
HD void ff_p2(const FhewFastDev& P, const FhewFastSmem& S, uint32_t poly, uint32_t grp) {
    const uint32_t lo = grp & 3u, hi = grp >> 2;
    const uint32_t P0 = swzf((hi << 6) | lo);
    uint32_t* d = S.dig + (poly << FF_LOGN);
    uint32_t x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = d[P0 ^ swzf((uint32_t)j << 2)];
    fast_fwd_regs<Lz32, 4, FF_TW_NC != 0>(P.m, x, S.tw, 8u + hi);
#pragma unroll
    for (int j = 0; j < 16; ++j) d[P0 ^ swzf((uint32_t)j << 2)] = x[j];
}
// ---- P3: last forward radix-4 pass (stages 7, 8) of `rows` digit polynomials + MAC + first inverse radix-4 pass -------------------
HD void ff_ld4(const uint32_t* d, uint32_t P0, uint32_t* x) {
#if defined(__CUDA_ARCH__)
    const uint4 v = *reinterpret_cast<const uint4*>(d + P0);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
#else
    for (int i = 0; i < 4; ++i) x[i] = d[P0 + i];
#endif
}
HD void ff_st4(uint32_t* d, uint32_t P0, const uint32_t* x) {
#if defined(__CUDA_ARCH__)
    *reinterpret_cast<uint4*>(d + P0) = make_uint4(x[0], x[1], x[2], x[3]);
#else
    for (int i = 0; i < 4; ++i) d[P0 + i] = x[i];
#endif
}
HD uint4 ff_ldg4(const uint4* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
HD void ff_p3(const FhewFastDev& P, const FhewFastSmem& S, const uint4* __restrict__ key /* [rows][128][2] */, uint32_t rows, uint32_t t) {
    const uint32_t P0 = swzf(t << 2);
    uint64_t sa[4] = {0, 0, 0, 0}, sb[4] = {0, 0, 0, 0};
    const TwPair<uint32_t> t0 = ff_tw(S.tw + 128u + t), t1 = ff_tw(S.tw + 256u + 2u * t), t2 = ff_tw(S.tw + 257u + 2u * t);
    FF_UNROLL(FF_P3_UNROLL)
    for (uint32_t k = 0; k < rows; ++k) {
        uint32_t x[4];
        ff_ld4(S.dig + (k << FF_LOGN), P0, x);
        const uint4 ka = ff_ldg4(key + ((size_t)k * FF_THREADS + t) * 2), kb = ff_ldg4(key + ((size_t)k * FF_THREADS + t) * 2 + 1);
#pragma unroll
        for (int i = 0; i < 4; ++i) x[i] = P.m.pre_red(x[i]);  // < 15 Q after 7 stages -> < 8 Q
        P.m.bf_fwd(x[0], x[2], t0);                              // stages 7, 8: < 12 Q < 2^32
        P.m.bf_fwd(x[1], x[3], t0);
        P.m.bf_fwd(x[0], x[1], t1);
        P.m.bf_fwd(x[2], x[3], t2);
        sa[0] += (uint64_t)ka.x * x[0]; sa[1] += (uint64_t)ka.y * x[1]; sa[2] += (uint64_t)ka.z * x[2]; sa[3] += (uint64_t)ka.w * x[3];
        sb[0] += (uint64_t)kb.x * x[0]; sb[1] += (uint64_t)kb.y * x[1]; sb[2] += (uint64_t)kb.z * x[2]; sb[3] += (uint64_t)kb.w * x[3];
    }
    uint32_t y[8];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        y[i] = ff_reduce64(P, sa[i]);
        y[4 + i] = ff_reduce64(P, sb[i]);
    }
    const TwPair<uint32_t> i0 = ff_tw(S.itw + 128u + t), i1 = ff_tw(S.itw + 256u + 2u * t), i2 = ff_tw(S.itw + 257u + 2u * t);
#pragma unroll
    for (int o = 0; o < 8; o += 4) {  // first inverse radix-4 pass (stages 8, 7) of a then b
        P.m.bf_inv(y[o + 0], y[o + 1], i1, 0);
        P.m.bf_inv(y[o + 2], y[o + 3], i2, 0);
        P.m.bf_inv(y[o + 0], y[o + 2], i0, 1);
        P.m.bf_inv(y[o + 1], y[o + 3], i0, 1);
    }
    ff_st4(S.dig, P0, y);
    ff_st4(S.dig + FF_RB * FF_N, P0, y + 4);
}
// ---- P4: inverse radix-16 pass (stages 6..3) on the result polynomial in digit slot `poly` (0: a, FF_RB: b) -----------------------
HD void ff_p4(const FhewFastDev& P, const FhewFastSmem& S, uint32_t poly, uint32_t grp) {
    const uint32_t lo = grp & 3u, hi = grp >> 2;
    const uint32_t P0 = swzf((hi << 6) | lo);
    uint32_t* d = S.dig + (poly << FF_LOGN);
    uint32_t x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = d[P0 ^ swzf((uint32_t)j << 2)];
    fast_inv_regs<Lz32, 4, false, FF_TW_NC != 0>(P.m, x, S.itw, 8u + hi, P.ninv, P.wninv);
#pragma unroll
    for (int j = 0; j < 16; ++j) d[P0 ^ swzf((uint32_t)j << 2)] = x[j];
}
// The same pass with the 16 elements of a group split over a LANE PAIR (lane, lane ^ 16): thread `part` runs the radix-8 half
// (stages 6..4) on elements 8 part .. 8 part + 7 - exactly fast_inv_regs<3> with chunk index 2 (8 + hi) + part - and the two
// threads then share the 8 butterflies of stage 3 (pairs (low, low + 8), one twiddle): part 0 takes low = 0..3, part 1
// low = 4..7, each receiving the four partner values it needs by one __shfl_xor each.  All four warps of the CTA work on P4
// (with ff_p4 only the first warp of each half did: r01 ncu, barrier stalls 22 %).  Same butterflies on the same values, so the
// result is bit-identical to ff_p4, which tests/hostsim keeps replaying.
#if defined(__CUDA_ARCH__)
DEV void ff_p4_split(const FhewFastDev& P, const FhewFastSmem& S, uint32_t poly, uint32_t grp, uint32_t part) {
    const uint32_t lo = grp & 3u, hi = grp >> 2;
    const uint32_t P0 = swzf((hi << 6) | lo);
    uint32_t* d = S.dig + (poly << FF_LOGN);
    uint32_t x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = d[P0 ^ swzf((uint32_t)(part * 8 + j) << 2)];
    fast_inv_regs<Lz32, 3, false, FF_TW_NC != 0>(P.m, x, S.itw, ((8u + hi) << 1) + part, P.ninv, P.wninv);
    const TwPair<uint32_t> t = ff_tw(S.itw + 8u + hi);
    // part 0 sends x[4..7] and keeps x[0..3]; part 1 sends x[0..3] (elements 8..11) and keeps x[4..7] (elements 12..15)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t send = part ? x[i] : x[4 + i];
        const uint32_t recv = __shfl_xor_sync(0xFFFFFFFFu, send, 16);
        uint32_t a = part ? recv : x[i], b = part ? x[4 + i] : recv;  // elements (4 part + i, 4 part + i + 8)
        P.m.bf_inv(a, b, t, 3);
        d[P0 ^ swzf((uint32_t)(4 * part + i) << 2)] = a;
        d[P0 ^ swzf((uint32_t)(4 * part + i + 8) << 2)] = b;
    }
}
#endif
// ---- P5: inverse radix-8 pass (stages 2..0, n^-1 folded) of polynomial h, canonical result (+ add(j)) -> acc_out[h][g + 64 j] -------
// v[j] = canonical coefficient g + 64 j of result polynomial h
template <typename Add>
HD void ff_p5_regs(const FhewFastDev& P, const FhewFastSmem& S, uint32_t g, uint32_t h, bool do_add, Add add, uint32_t* v) {
    const uint32_t P0 = swzf(g);
    const uint32_t* d = S.dig + ((h * FF_RB) << FF_LOGN);
    uint32_t x[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] = d[P0 ^ swzf((uint32_t)j << 6)];
    fast_inv_regs<Lz32, 3, true, FF_TW_NC != 0>(P.m, x, S.itw, 1u, P.ninv, P.wninv);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        v[j] = P.m.inv_canon(x[j]);
        if (do_add) {
            v[j] = alu_add(v[j], add(j));
            v[j] = umin_(v[j], v[j] - P.m.q);
        }
    }
}
template <typename Add>
HD void ff_p5(const FhewFastDev& P, const FhewFastSmem& S, uint32_t* acc_out, uint32_t g, uint32_t h, bool do_add, Add add) {
    uint32_t v[8];
    ff_p5_regs(P, S, g, h, do_add, add, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) acc_out[(h << FF_LOGN) + g + 64u * j] = v[j];
}

// One step on the accumulator S.acc (updated in place) as five phases.  Threads 0..63 form half 0, threads 64..127 half 1.
// Data flow between the halves (what lets most barriers be 64-thread barriers, so the halves can drift apart):
//   P1 (ext)  half h decomposes acc[h] into digit slots [h d, (h+1) d)            reads acc[h]            (own half)
//   P1 (auto) half h transforms its share [lo_h, hi_h) of the digits of a(X^t)    reads acc[0], acc[1]    (BOTH halves: full barrier before)
//   P2        half h runs the radix-16 pass on the slots it wrote in P1                                   (own half)
//   P3        every thread needs every slot at its four evaluation points; results a -> slot 0, b -> slot FF_RB   (full barrier before and after)
//   P4, P5    half h finishes result slot h FF_RB and writes acc[h]                                        (own half)
// Slot 0 lies in half 0's P1/P2 range and slot FF_RB = 4 in half 1's for an external product with d = 4; for the narrower
// cases (d < 4, automorphisms) half 1's slots start below 4 and slot 4 is free, so a half never touches a slot the other
// half may still be reading.  dig[7] (the parked b(X^t)) is written and read by half 1 only.
enum : int { FF_SYNC_HALF = 0, FF_SYNC_FULL = 1 };
struct FfStep {
    bool is_auto;
    uint32_t d, rows, tinv;
    const uint4* key;
};
HD FfStep ff_decode(const FhewFastDev& P, uint32_t step) {
    FfStep s;
    s.is_auto = (step & FHEW_STEP_AUTO) != 0;
    const uint32_t idx = step & 0x7FFFu;
    s.d = s.is_auto ? P.r_dec.d : P.g_dec.d;
    s.rows = s.is_auto ? s.d : 2 * s.d;
    s.key = s.is_auto ? P.ak4 + (size_t)idx * s.d * FF_THREADS * 2 : P.brk4 + (size_t)idx * (2 * s.d) * FF_THREADS * 2;
    s.tinv = s.is_auto ? P.ak_tinv[idx] : 0;
    return s;
}
// digit slots half h owns in P1 / P2
HD void ff_half_slots(const FfStep& s, uint32_t h, uint32_t& lo, uint32_t& hi) {
    if (s.is_auto) {
        const uint32_t half = (s.d + 1) / 2;
        lo = h * half;
        hi = h == 0 ? half : s.d;
    } else {
        lo = h * s.d;
        hi = lo + s.d;
    }
}
// external product: digits of acc.a -> slots [0, d), digits of acc.b -> [d, 2d)   (rgsw.rs:122-124)
// automorphism:     digits of a(X^t) -> slots [0, d), split between the halves    (rlwe.rs:182)
HD void ff_phase1(const FhewFastDev& P, const FhewFastSmem& S, const FfStep& s, uint32_t tid) {
    const uint32_t g = tid & 63u, h = tid >> 6;
    uint32_t lo, hi;
    ff_half_slots(s, h, lo, hi);
    const DecompParam& dp = s.is_auto ? P.r_dec : P.g_dec;
    if (s.is_auto)
        ff_p1(P, S, dp, S.acc, true, s.tinv, g, h, lo, hi, 0u);
    else
        ff_p1(P, S, dp, S.acc, false, 0u, g, h, 0u, s.d, lo);
}
HD void ff_phase2(const FhewFastDev& P, const FhewFastSmem& S, const FfStep& s, uint32_t tid) {
    const uint32_t l = tid & 63u, h = tid >> 6;
    uint32_t lo, hi;
    ff_half_slots(s, h, lo, hi);
#pragma unroll 1
    for (uint32_t u = l; u < (hi - lo) * 32u; u += 64u) ff_p2(P, S, lo + (u >> 5), u & 31u);
}
HD void ff_phase3(const FhewFastDev& P, const FhewFastSmem& S, const FfStep& s, uint32_t tid) { ff_p3(P, S, s.key, s.rows, tid); }
// FF_P4_SPLIT = 1: P4 groups split over lane pairs with __shfl_xor, all four warps busy (ff_p4_split); 0: one warp per half.
// Measured on B200 (r02, FHEW-T, batch 16 384, two runs each): split 489.6 k gates/s (32.60 ms per launch), one warp per half
// 500.8 k (31.85 ms).  With 7 CTAs resident per SM the idle warps of one CTA are covered by the other CTAs, and the split adds
// 4 shuffles + 8 selects per thread and doubles the threads that take part in the phase's barrier: it LOSES 2.2 %, so it is off.
#ifndef FF_P4_SPLIT
#define FF_P4_SPLIT 0
#endif
HD void ff_phase4(const FhewFastDev& P, const FhewFastSmem& S, const FfStep& s, uint32_t tid) {
#if defined(__CUDA_ARCH__) && FF_P4_SPLIT
    // thread (half h, warp w of the half, lane): group 16 w + (lane & 15), part lane >> 4
    ff_p4_split(P, S, (tid >> 6) * FF_RB, ((tid >> 5) & 1u) * 16u + (tid & 15u), (tid >> 4) & 1u);
#else
    if ((tid & 32u) == 0) ff_p4(P, S, (tid >> 6) * FF_RB, tid & 31u);  // first warp of each half
#endif
}
HD void ff_phase5(const FhewFastDev& P, const FhewFastSmem& S, const FfStep& s, uint32_t tid) {
    const uint32_t g = tid & 63u, h = tid >> 6;
    // key switch adds the (permuted) body: b' = sum ksk.b_k * limb_k + b(X^t)   (rlwe.rs:184)
    ff_p5(P, S, S.acc, g, h, s.is_auto && h == 1, [&](int j) { return S.dig[ff_park(g, j)]; });
}
// P5 of step `s` fused with P1 of the NEXT step `nx` when that one is an external product: thread (g, h) decomposes exactly the
// coefficients g + 64 j of acc[h] that it has just produced, so they stay in registers - no barrier, no shared-memory round
// trip, and acc itself is not written (nobody reads it before the next P5 rewrites it).  P5 reads result slot h FF_RB and P1
// writes slots [h d, (h + 1) d) at the thread's own 8 positions only, all loads of the slot come first.
HD void ff_phase51(const FhewFastDev& P, const FhewFastSmem& S, const FfStep& s, const FfStep& nx, uint32_t tid) {
    const uint32_t g = tid & 63u, h = tid >> 6;
    uint32_t v[8];
    ff_p5_regs(P, S, g, h, s.is_auto && h == 1, [&](int j) { return S.dig[ff_park(g, j)]; }, v);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = ff_dec_start(P, P.g_dec, v[j]);
    ff_p1_digits(P, S, P.g_dec, v, g, 0u, nx.d, h * nx.d);
}
// run(phase, scope): phase(tid) for every thread of the CTA, then a barrier of the given scope.  The caller has executed a
// full barrier after initialising S.acc; a full barrier follows the last phase.
template <typename Run>
HD void ff_run_steps(const FhewFastDev& P, const FhewFastSmem& S, const uint16_t* steps, uint32_t ns, Run run) {
    if (ns == 0) return;
    FfStep s = ff_decode(P, steps[0]);
    run([&](uint32_t tid) { ff_phase1(P, S, s, tid); }, FF_SYNC_HALF);
    for (uint32_t i = 0; i < ns; ++i) {
        run([&](uint32_t tid) { ff_phase2(P, S, s, tid); }, FF_SYNC_FULL);
        run([&](uint32_t tid) { ff_phase3(P, S, s, tid); }, FF_SYNC_FULL);
        run([&](uint32_t tid) { ff_phase4(P, S, s, tid); }, FF_SYNC_HALF);
        if (i + 1 == ns) {
            run([&](uint32_t tid) { ff_phase5(P, S, s, tid); }, FF_SYNC_FULL);
            break;
        }
        const FfStep nx = ff_decode(P, steps[i + 1]);
        if (nx.is_auto) {  // the automorphism reads both halves of acc
            run([&](uint32_t tid) { ff_phase5(P, S, s, tid); }, FF_SYNC_FULL);
            run([&](uint32_t tid) { ff_phase1(P, S, nx, tid); }, FF_SYNC_HALF);
        } else {
            run([&](uint32_t tid) { ff_phase51(P, S, s, nx, tid); }, FF_SYNC_HALF);
        }
        s = nx;
    }
}
// one isolated step (tests, util-level callers)
template <typename Run>
HD void ff_step(const FhewFastDev& P, const FhewFastSmem& S, uint32_t step, Run run) {
    const uint16_t one = (uint16_t)step;
    ff_run_steps(P, S, &one, 1u, run);
}
HD bool ff_step_is_auto(uint32_t step) { return (step & FHEW_STEP_AUTO) != 0; }

// acc init (bootstrapping.rs:158-169): acc = (0, f(X^-g) * X^(b*g))
template <typename FT>
HD void ff_init(const FhewFastDev& P, uint32_t* acc, const FT* __restrict__ f, uint32_t b2n, uint32_t tid) {
    const uint32_t m2 = 2 * FF_N - 1;
    const uint32_t t = (2 * FF_N - 5) & m2, e = (b2n * 5) & m2;
    for (uint32_t i = tid; i < (uint32_t)FF_N; i += FF_THREADS) {
        const uint32_t pos = (i * t + e) & m2;
        uint32_t v = (uint32_t)f[i];
        if (pos >= (uint32_t)FF_N) v = v == 0 ? 0 : P.m.q - v;
        acc[i] = 0;
        acc[FF_N + (pos & (FF_N - 1))] = v;
    }
}
// Rlwe::sample_extract(ct, 0) (rlwe.rs:193-202) + post_add on the body
template <typename OT>
HD void ff_extract(const FhewFastDev& P, const uint32_t* acc, uint32_t post_add, OT* out, uint32_t tid) {
    for (uint32_t k = tid; k < (uint32_t)FF_N; k += FF_THREADS) {
        const uint32_t v = k == 0 ? acc[0] : acc[FF_N - k];
        out[k] = (OT)(k == 0 || v == 0 ? v : P.m.q - v);
    }
    if (tid == 0) {
        uint32_t b = acc[FF_N] + post_add;
        out[FF_N] = (OT)umin_(b, b - P.m.q);
    }
}

}  // namespace fhe
