// FHEW / LMKCDEY blind-rotation building blocks (generic path: Mod32 for Q < 2^30, Mod64 above), __host__ __device__ so that
// tests/hostsim can replay the kernel logic on the CPU.
//
// Reference call sites replaced (all under scheme/fhew/src unless noted):
//   Base2Decomposor<Zq>::decompose        util/src/misc/decompose.rs:42-46, 91-112
//   Rgsw::external_product                rgsw.rs:116-128
//   Rlwe::automorphism / key_switch       rlwe.rs:177-191  (+ util/src/avec.rs:34-50)
//   Bootstrapping::blind_rotate(_core)    bootstrapping.rs:158-231
//   Rlwe::sample_extract                  rlwe.rs:193-202
// Dataflow differs from the reference (keys are pre-transformed once; products are accumulated in the
// evaluation domain: 2d forward + 2 inverse transforms per external product instead of 48), but all
// arithmetic is exact mod Q, so every accumulator value is bit-identical.
#pragma once
#include <cmath>

#include "modarith.cuh"
#include "ntt_core.cuh"
#include "ntt_fast.cuh"

namespace fhe {

struct DecompParam {
    uint32_t log_b, d, rounding_bits;
    uint64_t half;   // ((1 << rounding_bits) >> 1) mod q
    uint64_t neg_b;  // q - 2^log_b
};
// decompose.rs:49-64 (log_q = next_power_of_two(q).ilog2())
inline DecompParam make_decomp(uint64_t q, uint32_t log_b, uint32_t d) {
    DecompParam p;
    p.log_b = log_b;
    p.d = d;
    uint32_t log_q = 0;
    while (log_q < 64 && ((uint64_t)1 << log_q) < q) ++log_q;
    p.rounding_bits = log_q > log_b * d ? log_q - log_b * d : 0;
    p.half = (((uint64_t)1 << p.rounding_bits) >> 1) % q;
    p.neg_b = q - ((uint64_t)1 << log_b);
    return p;
}

// Signed base-2^log_b digits of one residue (decompose.rs:92-95 rounding_shr, zq.rs:83-89 to_center_u64,
// decompose.rs:101-111).  X is the working word: uint64_t mirrors the reference literally; uint32_t is
// equivalent whenever log_b * d <= 32 (only the low log_b*d bits of the centred value are ever consumed).
// Digits are produced least-significant first as residues mod q; `emit(k, digit)` receives them.
template <typename X, typename Emit>
HD void decompose_zq(uint64_t q, const DecompParam& dp, uint64_t v, Emit emit) {
    uint64_t rounded = v + dp.half;
    if (rounded >= q) rounded -= q;
    uint64_t sh = rounded >> dp.rounding_bits;
    X x = (sh < (q >> 1)) ? (X)sh : (X)(sh - q);  // two's complement centred value
    const X mask = ((X)1 << dp.log_b) - 1;
    const X b_by_2 = (X)1 << (dp.log_b - 1);
    for (uint32_t k = 0; k < dp.d; ++k) {
        X limb = x & mask;
        X carry = (limb + (x & 1) > b_by_2) ? 1 : 0;
        x >>= dp.log_b;
        x += carry;
        uint64_t dig = (uint64_t)limb + (carry ? dp.neg_b : 0);
        if (dig >= q) dig %= q;  // only reachable when 2^log_b > q
        emit(k, dig);
    }
}

// ---- LMKCDEY schedule --------------------------------------------------------------------------------
// dlog table (host-built, 2N entries of u16): for odd a in [0, 2N): l | 0x8000 if a = -g^l, l if a = +g^l (g = 5,
// l < N/2); 0xFFFF for even a (only a == 0 is legal: it is skipped, bootstrapping.rs:220).
inline void build_dlog_table(uint32_t n, uint16_t* tab /* 2n */) {
    uint32_t m = 2 * n;
    for (uint32_t i = 0; i < m; ++i) tab[i] = 0xFFFF;
    uint64_t pw = 1 % m;
    for (uint32_t l = 0; l < n / 2; ++l) {
        tab[pw] = (uint16_t)l;
        tab[(m - pw) % m] = (uint16_t)(l | 0x8000);
        pw = (pw * 5) % m;
    }
    if (n == 1) tab[1] = 0;  // degenerate ring, never used for bootstrapping
}
// Step encoding: bit 15 = 1 -> automorphism with ak[idx], 0 -> external product with brk[idx].
#define FHEW_STEP_AUTO 0x8000u
// Sequential construction of blind_rotate_core's control flow (bootstrapping.rs:172-209) for one LWE mask
// `a` (n_s entries mod 2N).  `cnt`/`pos` are scratch arrays of N entries (u16), `sorted` of n_s entries.
// Returns the number of steps written, or 0xFFFFFFFF if an even non-zero exponent is met (`unreachable!`).
template <typename AT>
HD uint32_t build_schedule(uint32_t n, uint32_t n_s, uint32_t w, const AT* a, const uint16_t* dlog, uint16_t* cnt /* n */,
                           uint16_t* sorted /* n_s */, uint16_t* steps) {
    const uint32_t half = n / 2;
    // bucket index: minus side [0, half), plus side [half, n)
    for (uint32_t i = 0; i < n; ++i) cnt[i] = 0;
    bool bad = false;
    for (uint32_t j = 0; j < n_s; ++j) {
        uint32_t aj = (uint32_t)a[j];
        if (aj == 0) continue;
        uint16_t e = dlog[aj];
        if (e == 0xFFFF) {
            bad = true;
            continue;
        }
        uint32_t bucket = (e & 0x8000) ? (e & 0x7FFF) : half + e;
        cnt[bucket]++;
    }
    if (bad) return 0xFFFFFFFFu;
    // exclusive prefix -> start offsets (stored back into cnt as running insert positions)
    uint32_t run = 0;
    for (uint32_t i = 0; i < n; ++i) {
        uint32_t c = cnt[i];
        cnt[i] = (uint16_t)run;
        run += c;
    }
    for (uint32_t j = 0; j < n_s; ++j) {  // stable: j ascending inside each bucket (Vec::push order)
        uint32_t aj = (uint32_t)a[j];
        if (aj == 0) continue;
        uint16_t e = dlog[aj];
        uint32_t bucket = (e & 0x8000) ? (e & 0x7FFF) : half + e;
        sorted[cnt[bucket]++] = (uint16_t)j;
    }
    // now cnt[bucket] = end offset of the bucket; start = end of previous bucket (0 for bucket 0)
    uint32_t ns = 0, v = 0;
    for (uint32_t side = 0; side < 2; ++side) {
        const uint32_t off = side * half;
        for (uint32_t l = half - 1; l >= 1; --l) {
            uint32_t b = off + l;
            uint32_t st = b == 0 ? 0 : cnt[b - 1], en = cnt[b];
            for (uint32_t x = st; x < en; ++x) steps[ns++] = sorted[x];
            v += 1;
            uint32_t bprev = off + l - 1;
            uint32_t pst = bprev == 0 ? 0 : cnt[bprev - 1], pen = cnt[bprev];
            if (pen > pst || v == w || l == 1) {
                steps[ns++] = (uint16_t)(FHEW_STEP_AUTO | v);
                v = 0;
            }
        }
        {
            uint32_t b = off;
            uint32_t st = b == 0 ? 0 : cnt[b - 1], en = cnt[b];
            for (uint32_t x = st; x < en; ++x) steps[ns++] = sorted[x];
        }
        if (side == 0) steps[ns++] = (uint16_t)(FHEW_STEP_AUTO | 0);
    }
    return ns;
}
inline uint32_t max_schedule_steps(uint32_t n, uint32_t n_s) { return n_s + n + 2; }

// ---- per-CTA accumulator steps ---------------------------------------------------------------------------
// M = Mod32 (Q < 2^30, the single-key parameter sets) or Mod64 (Q < 2^62: the 54/55-bit multi-key set of
// examples/multi_key_uint8.rs:15-29); W = M::W is the word every residue, twiddle and key entry is stored in.
template <typename W>
struct alignas(2 * sizeof(W)) KeyPair {
    W x, y;  // {a, b} of one key row at one evaluation point
};
template <typename M>
struct FhewDevT {
    typedef typename M::W W;
    M m;
    int log_n;
    uint32_t n_s, w;
    DecompParam g_dec;  // RGSW decomposor
    DecompParam r_dec;  // RLWE key-switch decomposor
    uint32_t small_digits;  // 1 if both decomposors satisfy log_b * d <= 32 (u32 digit extraction)
    const TwPair<W>* tw;
    const TwPair<W>* itw;
    TwPair<W> ninv, wninv;
    const KeyPair<W>* brk;  // [n_s][2*g_d][N] of {a, b} in evaluation form
    const KeyPair<W>* ak;   // [w+1][r_d][N] of {a, b} in evaluation form
    const uint16_t* dlog;
    uint32_t ak_t[40];  // automorphism exponents t mod 2N for ak[0..w]
    // 64-bit moduli below 2^56: the transforms use the lazy butterflies of the fast NTT path (ntt_fast.cuh: no reduction in
    // the forward direction - at most (1 + 4 log N) q < 128 q -, one Barrett per pass on the sum chain in the inverse)
    uint32_t lazy;
    Lz64 lz;
    W c64;  // 2^64 mod q (64-bit moduli: folds the high word of the exact 128-bit MAC sums)
};
typedef FhewDevT<Mod32> FhewDev;
template <int R>
HD void fhew_fwd_group_lz(const Lz64& m, uint64_t* s, int c, int t0, uint32_t g, const TwPair<uint64_t>* __restrict__ tw) {
    const int L = c - t0 - R;
    const uint32_t lo = g & ((1u << L) - 1u), hi = g >> L;
    const uint32_t base = (hi << (L + R)) | lo;
    uint64_t x[1 << R];
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) x[j] = s[swz<uint64_t>(base | ((uint32_t)j << L))];
    fast_fwd_regs<Lz64, R>(m, x, tw, (1u << t0) + hi);
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) s[swz<uint64_t>(base | ((uint32_t)j << L))] = x[j];
}
// pass invariant: every value < 16 q on entry and on exit; LAST (t0 == 0): n^-1 folded, canonical results
template <int R, bool LAST>
HD void fhew_inv_group_lz(const Lz64& m, uint64_t* s, int c, int t0, uint32_t g, const TwPair<uint64_t>* __restrict__ itw, TwPair<uint64_t> ninv,
                          TwPair<uint64_t> wninv) {
    const int L = c - t0 - R;
    const uint32_t lo = g & ((1u << L) - 1u), hi = g >> L;
    const uint32_t base = (hi << (L + R)) | lo;
    uint64_t x[1 << R];
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) x[j] = s[swz<uint64_t>(base | ((uint32_t)j << L))];
    fast_inv_regs<Lz64, R, LAST>(m, x, itw, (1u << t0) + hi, ninv, wninv);
    if (LAST) {
#pragma unroll
        for (int j = 0; j < (1 << R); ++j) x[j] = m.inv_canon(x[j]);
    } else {
        x[0] = m.inv_pass_fix(x[0]);
    }
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) s[swz<uint64_t>(base | ((uint32_t)j << L))] = x[j];
}

// Shared-memory working set of one accumulator (all polynomials swizzled with swz<W>):
//   acc_a[N], acc_b[N]   coefficient form, canonical
//   dig[kmax][N]         digit polynomials / evaluation-domain products
template <typename W>
HD W* fhew_dig(W* smem, uint32_t n, uint32_t k) { return smem + (size_t)(2 + k) * n; }

// Phase D (external product): digits of acc.a -> dig[0..d), digits of acc.b -> dig[d..2d)   (rgsw.rs:122-124)
template <typename M>
HD void fhew_phase_decomp_ext(const FhewDevT<M>& P, typename M::W* smem, uint32_t tid, uint32_t nthr) {
    typedef typename M::W W;
    const uint32_t n = 1u << P.log_n, d = P.g_dec.d;
    for (uint32_t i = tid; i < n; i += nthr) {
        const uint32_t si = swz<W>(i);
        for (uint32_t h = 0; h < 2; ++h) {
            const W v = smem[h * n + si];
            W* base = fhew_dig(smem, n, h * d) + si;
            if (P.small_digits)
                decompose_zq<uint32_t>(P.m.q, P.g_dec, v, [&](uint32_t k, uint64_t dg) { base[(size_t)k * n] = (W)dg; });
            else
                decompose_zq<uint64_t>(P.m.q, P.g_dec, v, [&](uint32_t k, uint64_t dg) { base[(size_t)k * n] = (W)dg; });
        }
    }
}
// Phase D (automorphism + key switch), step 1: digits of a(X^t) -> dig[0..d)   (rlwe.rs:80-82,182; avec.rs:34-50)
template <typename M>
HD void fhew_phase_decomp_auto_a(const FhewDevT<M>& P, typename M::W* smem, uint32_t t, uint32_t tid, uint32_t nthr) {
    typedef typename M::W W;
    const uint32_t n = 1u << P.log_n;
    for (uint32_t i = tid; i < n; i += nthr) {
        const uint32_t it = (i * t) & (2 * n - 1);
        W v = smem[swz<W>(i)];
        if (it >= n) v = P.m.neg(v);
        W* base = fhew_dig(smem, n, 0) + swz<W>(it & (n - 1));
        if (P.small_digits)
            decompose_zq<uint32_t>(P.m.q, P.r_dec, v, [&](uint32_t k, uint64_t dg) { base[(size_t)k * n] = (W)dg; });
        else
            decompose_zq<uint64_t>(P.m.q, P.r_dec, v, [&](uint32_t k, uint64_t dg) { base[(size_t)k * n] = (W)dg; });
    }
}
// step 2 (after a barrier; acc_a is dead): b(X^t) -> acc_a region
template <typename M>
HD void fhew_phase_auto_b(const FhewDevT<M>& P, typename M::W* smem, uint32_t t, uint32_t tid, uint32_t nthr) {
    typedef typename M::W W;
    const uint32_t n = 1u << P.log_n;
    for (uint32_t i = tid; i < n; i += nthr) {
        const uint32_t it = (i * t) & (2 * n - 1);
        W v = smem[n + swz<W>(i)];
        if (it >= n) v = P.m.neg(v);
        smem[swz<W>(it & (n - 1))] = v;
    }
}
// forward / inverse NTT passes over `npoly` consecutive digit polynomials
template <int R, typename M>
HD void fhew_fwd_pass(const FhewDevT<M>& P, typename M::W* polys, uint32_t npoly, int t0, uint32_t tid, uint32_t nthr) {
    const int c = P.log_n;
    const uint32_t lg_groups = (uint32_t)(c - R);
    const uint32_t total = npoly << lg_groups;
    for (uint32_t u = tid; u < total; u += nthr) {
        const uint32_t poly = u >> lg_groups, g = u & ((1u << lg_groups) - 1u);
        if constexpr (sizeof(typename M::W) == 8) {
            if (P.lazy) {
                fhew_fwd_group_lz<R>(P.lz, polys + ((size_t)poly << c), c, t0, g, P.tw);
                continue;
            }
        }
        fwd_tile_group<M, R>(P.m, polys + ((size_t)poly << c), c, t0, 0, 0, g, P.tw);
    }
}
template <int R, bool LAST, typename M>
HD void fhew_inv_pass(const FhewDevT<M>& P, typename M::W* polys, uint32_t npoly, int t0, uint32_t tid, uint32_t nthr) {
    const int c = P.log_n;
    const uint32_t lg_groups = (uint32_t)(c - R);
    const uint32_t total = npoly << lg_groups;
    for (uint32_t u = tid; u < total; u += nthr) {
        const uint32_t poly = u >> lg_groups, g = u & ((1u << lg_groups) - 1u);
        if constexpr (sizeof(typename M::W) == 8) {
            if (P.lazy) {
                fhew_inv_group_lz<R, LAST>(P.lz, polys + ((size_t)poly << c), c, t0, g, P.itw, P.ninv, P.wninv);
                continue;
            }
        }
        inv_tile_group<M, R, LAST>(P.m, polys + ((size_t)poly << c), c, t0, 0, 0, g, P.itw, P.ninv, P.wninv);
    }
}
template <typename M>
HD void fhew_fwd_pass_dyn(const FhewDevT<M>& P, typename M::W* polys, uint32_t npoly, int t0, int r, uint32_t tid, uint32_t nthr) {
    if (r == 3)
        fhew_fwd_pass<3>(P, polys, npoly, t0, tid, nthr);
    else if (r == 2)
        fhew_fwd_pass<2>(P, polys, npoly, t0, tid, nthr);
    else
        fhew_fwd_pass<1>(P, polys, npoly, t0, tid, nthr);
}
template <typename M>
HD void fhew_inv_pass_dyn(const FhewDevT<M>& P, typename M::W* polys, uint32_t npoly, int t0, int r, uint32_t tid, uint32_t nthr) {
    if (t0 == 0) {
        if (r == 3)
            fhew_inv_pass<3, true>(P, polys, npoly, t0, tid, nthr);
        else if (r == 2)
            fhew_inv_pass<2, true>(P, polys, npoly, t0, tid, nthr);
        else
            fhew_inv_pass<1, true>(P, polys, npoly, t0, tid, nthr);
    } else {
        if (r == 3)
            fhew_inv_pass<3, false>(P, polys, npoly, t0, tid, nthr);
        else if (r == 2)
            fhew_inv_pass<2, false>(P, polys, npoly, t0, tid, nthr);
        else
            fhew_inv_pass<1, false>(P, polys, npoly, t0, tid, nthr);
    }
}
// Phase M: evaluation-domain multiply-accumulate against `rows` pre-transformed key rows {a,b}[rows][N];
// results overwrite dig[0] (a) and dig[1] (b) at the same index (each index is owned by one thread).
// Digits come out of the forward transform in [0,4q) and are canonicalised here so that rows * q^2 < 2^64.
template <typename M>
HD void fhew_phase_mac(const FhewDevT<M>& P, typename M::W* smem, const KeyPair<typename M::W>* __restrict__ key, uint32_t rows, uint32_t tid,
                       uint32_t nthr) {
    typedef typename M::W W;
    const uint32_t n = 1u << P.log_n;
    W* dig = fhew_dig(smem, n, 0);
    for (uint32_t i = tid; i < n; i += nthr) {
        const uint32_t si = swz<W>(i);
        if constexpr (sizeof(W) == 4) {
            uint64_t sa = 0, sb = 0;
            for (uint32_t k = 0; k < rows; ++k) {
                const KeyPair<W> kv = key[(size_t)k * n + i];
                const uint32_t dg = P.m.canon4(dig[(size_t)k * n + si]);
                sa += (uint64_t)kv.x * dg;
                sb += (uint64_t)kv.y * dg;
            }
            dig[si] = P.m.reduce64(sa);
            dig[n + si] = P.m.reduce64(sb);
        } else if (P.lazy && P.m.s >= 33) {
            // 2^32 <= q < 2^56 and rows <= 32: the sum of rows products is < 2^117, accumulated exactly in 128 bits (four wide products per
            // term instead of a full Barrett product) and reduced once: hi 2^64 + lo = hi (2^64 mod q) + lo (mod q).  The canonical
            // result equals the per-term canonical Zq arithmetic of the reference.  Key rows are fetched two ahead of their use.
            unsigned __int128 sa = 0, sb = 0;
            KeyPair<W> k0 = key[i], k1 = rows > 1 ? key[(size_t)n + i] : k0;
            for (uint32_t k = 0; k < rows; ++k) {
                const KeyPair<W> kv = k0;
                k0 = k1;
                if (k + 2 < rows) k1 = key[(size_t)(k + 2) * n + i];
                const W dg = P.lz.canon(dig[(size_t)k * n + si]);
                sa += (unsigned __int128)kv.x * dg;
                sb += (unsigned __int128)kv.y * dg;
            }
            auto fold = [&](unsigned __int128 v) {
                const W hi = (W)(v >> 64), lo = (W)v;  // hi < 2^53, lo < 2^64: both inside the Barrett input range 2^(2s), s >= 33
                U128 h, l;
                h.lo = hi;
                h.hi = 0;
                l.lo = lo;
                l.hi = 0;
                return P.m.add(P.m.mul(P.m.reduce128(h), P.c64), P.m.reduce128(l));
            };
            dig[si] = fold(sa);
            dig[n + si] = fold(sb);
        } else {  // 64-bit modulus: every product is reduced (rows * q^2 does not fit the 128-bit Barrett's input range)
            W sa = 0, sb = 0;
            for (uint32_t k = 0; k < rows; ++k) {
                const KeyPair<W> kv = key[(size_t)k * n + i];
                const W dg = P.m.canon4(dig[(size_t)k * n + si]);
                sa = P.m.add(sa, P.m.mul(kv.x, dg));
                sb = P.m.add(sb, P.m.mul(kv.y, dg));
            }
            dig[si] = sa;
            dig[n + si] = sb;
        }
    }
}
// Phase F: acc <- (dig[0], dig[1] (+ b(X^t) parked in the acc_a region when add_b))
template <typename M>
HD void fhew_phase_finish(const FhewDevT<M>& P, typename M::W* smem, bool add_b, uint32_t tid, uint32_t nthr) {
    typedef typename M::W W;
    const uint32_t n = 1u << P.log_n;
    const W* dig = fhew_dig(smem, n, 0);
    for (uint32_t i = tid; i < n; i += nthr) {
        const uint32_t si = swz<W>(i);
        W a = P.m.redq(dig[si]);
        W b = P.m.redq(dig[n + si]);
        if (add_b) b = P.m.add(b, smem[si]);
        smem[si] = a;
        smem[n + si] = b;
    }
}
// acc init (bootstrapping.rs:158-169): acc = (0, f(X^-g) * X^(b*g)); both maps are signed permutations, composed here.
template <typename M, typename FT>
HD void fhew_phase_init(const FhewDevT<M>& P, typename M::W* smem, const FT* __restrict__ f, uint32_t b2n, uint32_t tid, uint32_t nthr) {
    typedef typename M::W W;
    const uint32_t n = 1u << P.log_n, m2 = 2 * n - 1;
    const uint32_t t = (2 * n - 5) & m2;   // -g mod 2N
    const uint32_t e = (b2n * 5) & m2;     // b*g mod 2N (the centred value and its residue give the same monomial)
    for (uint32_t i = tid; i < n; i += nthr) {
        const uint32_t pos = (i * t + e) & m2;
        W v = (W)f[i];
        if (pos >= n) v = P.m.neg(v);
        smem[swz<W>(i)] = 0;
        smem[n + swz<W>(pos & (n - 1))] = v;
    }
}

// One full step (external product or automorphism).  `run(phase)` executes phase(tid, nthr) for every thread of
// the CTA followed by a barrier: on the device run = { phase(threadIdx.x, blockDim.x); __syncthreads(); },
// in tests/hostsim it loops tid sequentially.
template <typename M, typename Run>
HD void fhew_step(const FhewDevT<M>& P, typename M::W* smem, uint32_t step, Run run) {
    typedef typename M::W W;
    const uint32_t n = 1u << P.log_n;
    const PassPlan plan = make_plan(P.log_n);
    const bool is_auto = (step & FHEW_STEP_AUTO) != 0;
    const uint32_t idx = step & 0x7FFFu;
    const uint32_t rows = is_auto ? P.r_dec.d : 2 * P.g_dec.d;
    const KeyPair<W>* key = is_auto ? P.ak + (size_t)idx * rows * n : P.brk + (size_t)idx * rows * n;
    const uint32_t t = is_auto ? P.ak_t[idx] : 0;
    W* dig = fhew_dig(smem, n, 0);
    if (!is_auto)
        run([&](uint32_t tid, uint32_t nthr) { fhew_phase_decomp_ext(P, smem, tid, nthr); });
    else
        run([&](uint32_t tid, uint32_t nthr) { fhew_phase_decomp_auto_a(P, smem, t, tid, nthr); });
    for (int pi = 0; pi < plan.n; ++pi) {
        run([&](uint32_t tid, uint32_t nthr) {
            // b(X^t) is parked in the (now dead) acc_a region; disjoint from dig[], so it shares the first pass's phase
            if (is_auto && pi == 0) fhew_phase_auto_b(P, smem, t, tid, nthr);
            fhew_fwd_pass_dyn(P, dig, rows, plan.t0[pi], plan.r[pi], tid, nthr);
        });
    }
    run([&](uint32_t tid, uint32_t nthr) { fhew_phase_mac(P, smem, key, rows, tid, nthr); });
    for (int pi = plan.n - 1; pi >= 0; --pi)
        run([&](uint32_t tid, uint32_t nthr) { fhew_inv_pass_dyn(P, dig, 2, plan.t0[pi], plan.r[pi], tid, nthr); });
    run([&](uint32_t tid, uint32_t nthr) { fhew_phase_finish(P, smem, is_auto, tid, nthr); });
}


// ---- LWE side: mod switches and key switch (lwe.rs:90-99, 151-160; zq.rs:128-140) ------------------------------
// IEEE double ops that must not be contracted or reassociated (the reference computes (v*q')/q in f64).
HD double f64_mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    volatile double r = a * b;
    return r;
#endif
}
HD double f64_div_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __ddiv_rn(a, b);
#else
    volatile double r = a / b;
    return r;
#endif
}
HD double f64_round_half_away(double x) {  // Rust f64::round
    double t = trunc(x);
    double frac = fabs(x - t);  // exact
    if (frac >= 0.5) t += (x < 0.0 ? -1.0 : 1.0);
    return t;
}
// Zq::mod_switch (zq.rs:128-130): from_f64(q', (v as f64 * q' as f64) / q as f64)
HD uint64_t zq_mod_switch_dev(uint64_t v, uint64_t q, uint64_t qp) {
    double x = f64_div_rn(f64_mul_rn((double)v, (double)qp), (double)q);
    long long r = (long long)f64_round_half_away(x);
    long long m = r % (long long)qp;
    if (m < 0) m += (long long)qp;
    return (uint64_t)m;
}
// Zq::mod_switch_odd (zq.rs:132-140)
HD uint64_t zq_mod_switch_odd_dev(uint64_t v, uint64_t q, uint64_t qp) {
    double x = f64_div_rn(f64_mul_rn((double)v, (double)qp), (double)q);
    double u = floor(x);
    if (u == 0.0) return ((uint64_t)f64_round_half_away(x)) % qp;
    return (((uint64_t)u) | 1ull) % qp;
}

struct LweKsDev {
    uint64_t big_q;      // modulus of the incoming ciphertext (Q); ignored when !switch_in
    uint32_t n;          // input dimension N
    uint32_t n_s;        // output dimension
    uint64_t q_ks;       // power of two <= 2^32
    uint64_t q_out;      // 2N when switch_out
    DecompParam ks_dec;
    const uint32_t* ksk;  // [N * d_ks][n_s + 1]: a_0..a_{n_s-1}, b ; index = digit*N + coefficient (lwe.rs:114,156)
    uint32_t switch_in, switch_out;
};
// Phase 1 for ciphertext slot g of the CTA: digits of mod_switch(a_i) -> digs[g][k*N + i]; returns b' via *b_out (thread 0)
template <typename CT>
HD void lwe_phase_digits(const LweKsDev& P, const CT* __restrict__ ct, uint32_t* digs, uint32_t* b_out, uint32_t tid, uint32_t nthr) {
    for (uint32_t i = tid; i < P.n; i += nthr) {
        uint64_t v = (uint64_t)ct[i];
        if (P.switch_in) v = zq_mod_switch_dev(v, P.big_q, P.q_ks);
        decompose_zq<uint32_t>(P.q_ks, P.ks_dec, v, [&](uint32_t k, uint64_t dg) { digs[(size_t)k * P.n + i] = (uint32_t)dg; });
    }
    if (tid == 0) {
        uint64_t b = (uint64_t)ct[P.n];
        if (P.switch_in) b = zq_mod_switch_dev(b, P.big_q, P.q_ks);
        *b_out = (uint32_t)b;
    }
}
// Phase 2: output column j (j == n_s is the body) for G ciphertext slots; arithmetic mod 2^32 is exact mod q_ks | 2^32
template <int G>
HD void lwe_phase_gemv(const LweKsDev& P, const uint32_t* digs /* [G][N*d] */, uint32_t j, uint32_t* acc /* [G] */) {
    const uint32_t len = P.n * P.ks_dec.d, ld = P.n_s + 1;
    for (int g = 0; g < G; ++g) acc[g] = 0;
    // (four key rows per iteration with 16-byte digit loads and 8 slots per CTA was measured: 5 % slower)
    for (uint32_t idx = 0; idx < len; ++idx) {
        const uint32_t kv = P.ksk[(size_t)idx * ld + j];
        for (int g = 0; g < G; ++g) acc[g] += kv * digs[(size_t)g * len + idx];
    }
}
HD uint64_t lwe_phase_out(const LweKsDev& P, uint32_t acc, uint32_t j, uint32_t b_in) {
    uint64_t v = (uint64_t)acc;
    if (j == P.n_s) v += b_in;
    v &= (P.q_ks - 1);
    if (P.switch_out) v = zq_mod_switch_odd_dev(v, P.q_ks, P.q_out);
    return v;
}

// Rlwe::sample_extract(ct, 0) (rlwe.rs:193-202) from the swizzled accumulator: out = [a_0, -a_{N-1}, .., -a_1, b_0 + post_add]
template <typename M, typename OT>
HD void fhew_phase_extract(const FhewDevT<M>& P, const typename M::W* smem, typename M::W post_add, OT* out, uint32_t tid, uint32_t nthr) {
    typedef typename M::W W;
    const uint32_t n = 1u << P.log_n;
    for (uint32_t k = tid; k < n; k += nthr) {
        W v = k == 0 ? smem[swz<W>(0)] : P.m.neg(smem[swz<W>(n - k)]);
        out[k] = (OT)v;
    }
    if (tid == 0) out[n] = (OT)P.m.add(smem[n + swz<W>(0)], post_add);
}

}  // namespace fhe
