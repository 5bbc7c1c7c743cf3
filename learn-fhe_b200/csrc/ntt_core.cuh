// Negacyclic NTT building blocks shared by every kernel that transforms polynomials.
//
// Semantics follow the reference exactly (util/src/ring/fft.rs:40-77 with the twiddle rule of
// util/src/ring/fft/zq.rs:58-67): forward = Cooley-Tukey, natural-order input -> bit-reversed output,
// stage l uses table entries [2^l, 2^(l+1)); inverse = Gentleman-Sande, bit-reversed -> natural, then
// * n^-1 (fused into the last stage here; modular arithmetic is exact so results are identical).
//
// Harvey lazy butterflies: forward values live in [0,4q), inverse values in [0,2q); canonicalised once
// at the end.  Twiddles are stored as interleaved (w, w') pairs, w' = floor(w * 2^BITS / q).
//
// A "pass" performs R consecutive stages on 2^R elements held in registers.  Passes read/write a
// shared-memory tile through a bank-conflict-free XOR swizzle.  All functions are __host__ __device__
// so tests/hostsim can run the exact index/arithmetic logic sequentially on the CPU.
#pragma once
#include "modarith.cuh"

namespace fhe {

template <typename W>
struct TwPair {
    W w, wp;
};

// ---- butterflies ---------------------------------------------------------------------------------
// fft.rs:94-101 (dit), lazy: x,y in [0,4q) -> [0,4q)
template <typename A>
HD void bf_fwd(const A& m, typename A::W& x, typename A::W& y, TwPair<typename A::W> t) {
    typename A::W xr = m.red2q(x);
    typename A::W ty = m.shoup_lazy(y, t.w, t.wp);
    x = xr + ty;
    y = xr + m.q2 - ty;
}
// fft.rs:103-109 (dif), lazy: x,y in [0,2q) -> [0,2q)
template <typename A>
HD void bf_inv(const A& m, typename A::W& x, typename A::W& y, TwPair<typename A::W> t) {
    typename A::W s = x + y;
    typename A::W d = x + m.q2 - y;
    x = m.red2q(s);
    y = m.shoup_lazy(d, t.w, t.wp);
}
// last inverse stage with n^-1 folded in: x' = (x+y)*ninv, y' = (x-y)*(w*ninv); outputs in [0,2q)
template <typename A>
HD void bf_inv_last(const A& m, typename A::W& x, typename A::W& y, TwPair<typename A::W> ninv, TwPair<typename A::W> wninv) {
    typename A::W s = x + y;
    typename A::W d = x + m.q2 - y;
    x = m.shoup_lazy(s, ninv.w, ninv.wp);
    y = m.shoup_lazy(d, wninv.w, wninv.wp);
}

// ---- register passes -----------------------------------------------------------------------------
// R forward stages on x[0 .. 2^R).  Element j of the group sits at polynomial position
//   (hi << (L+R)) | (j << L) | lo ; the pass covers global stages l0 .. l0+R-1 and `tb` = 2^l0 + hiIdx where
// hiIdx = position >> (logN - l0).  Twiddle index of sub-stage u for the pair whose upper-bit prefix is
// `top` (u bits) is (tb << u) + top  — i.e. 2^(l0+u) + (position >> (logN - l0 - u)).
template <typename A, int R>
HD void fwd_pass_regs(const A& m, typename A::W* x, const TwPair<typename A::W>* __restrict__ tw, uint32_t tb) {
#pragma unroll
    for (int u = 0; u < R; ++u) {
        const int h = 1 << (R - 1 - u);
#pragma unroll
        for (int pi = 0; pi < (1 << (R - 1)); ++pi) {
            const int top = pi >> (R - 1 - u);
            const int low = pi & (h - 1);
            const int j = (top << (R - u)) | low;
            TwPair<typename A::W> t = tw[(tb << u) + top];
            bf_fwd(m, x[j], x[j + h], t);
        }
    }
}
// R inverse stages (global stages l0+R-1 down to l0) on x[0 .. 2^R); `itw` is the inverse table.
// If LAST, global stage 0 is the final one of the transform (l0 must be 0) and n^-1 is folded in.
template <typename A, int R, bool LAST>
HD void inv_pass_regs(const A& m, typename A::W* x, const TwPair<typename A::W>* __restrict__ itw, uint32_t tb,
                      TwPair<typename A::W> ninv, TwPair<typename A::W> wninv) {
#pragma unroll
    for (int u = R - 1; u >= 0; --u) {
        const int h = 1 << (R - 1 - u);
#pragma unroll
        for (int pi = 0; pi < (1 << (R - 1)); ++pi) {
            const int top = pi >> (R - 1 - u);
            const int low = pi & (h - 1);
            const int j = (top << (R - u)) | low;
            if (LAST && u == 0) {
                bf_inv_last(m, x[j], x[j + h], ninv, wninv);
            } else {
                TwPair<typename A::W> t = itw[(tb << u) + top];
                bf_inv(m, x[j], x[j + h], t);
            }
        }
    }
}

// ---- shared-memory swizzle -------------------------------------------------------------------------
// Conflict-free for (a) 32 (16 for 8-byte words) consecutive elements, (b) the stride-8 accesses of the
// last radix-8 pass, (c) the stride-64 groups-of-8 accesses of the second-to-last radix-8 pass.
template <typename W>
HD uint32_t swz(uint32_t p);
template <>
HD uint32_t swz<uint32_t>(uint32_t p) {
    return p ^ ((p >> 5) & 7u) ^ (((p >> 6) & 3u) << 3);
}
template <>
HD uint32_t swz<uint64_t>(uint32_t p) {
    return p ^ ((p >> 4) & 7u) ^ (((p >> 6) & 1u) << 3);
}

// ---- tile passes -----------------------------------------------------------------------------------
// One thread-unit of a pass over a tile of 2^c elements held in (swizzled) shared memory `s`.
//   t0 = first local stage of the pass, R = stages in the pass, g = group id in [0, 2^(c-R)).
//   Global stage of local stage t is s0 + t; the tile is chunk `k` (of 2^s0 chunks) of the polynomial.
template <typename A, int R>
HD void fwd_tile_group(const A& m, typename A::W* s, int c, int t0, int s0, uint32_t k, uint32_t g,
                       const TwPair<typename A::W>* __restrict__ tw) {
    typedef typename A::W W;
    const int L = c - t0 - R;
    const uint32_t lo = g & ((1u << L) - 1u), hi = g >> L;
    const uint32_t base = (hi << (L + R)) | lo;
    W x[1 << R];
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) x[j] = s[swz<W>(base | ((uint32_t)j << L))];
    const uint32_t tb = (1u << (s0 + t0)) + (k << t0) + hi;
    fwd_pass_regs<A, R>(m, x, tw, tb);
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) s[swz<W>(base | ((uint32_t)j << L))] = x[j];
}
template <typename A, int R, bool LAST>
HD void inv_tile_group(const A& m, typename A::W* s, int c, int t0, int s0, uint32_t k, uint32_t g,
                       const TwPair<typename A::W>* __restrict__ itw, TwPair<typename A::W> ninv, TwPair<typename A::W> wninv) {
    typedef typename A::W W;
    const int L = c - t0 - R;
    const uint32_t lo = g & ((1u << L) - 1u), hi = g >> L;
    const uint32_t base = (hi << (L + R)) | lo;
    W x[1 << R];
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) x[j] = s[swz<W>(base | ((uint32_t)j << L))];
    const uint32_t tb = (1u << (s0 + t0)) + (k << t0) + hi;
    inv_pass_regs<A, R, LAST>(m, x, itw, tb, ninv, wninv);
#pragma unroll
    for (int j = 0; j < (1 << R); ++j) s[swz<W>(base | ((uint32_t)j << L))] = x[j];
}

// Pass plan for a tile of c stages: an optional remainder pass (c % 3 stages) first, then radix-8 passes.
// Forward runs the plan front to back; inverse back to front.
struct PassPlan {
    int n;       // number of passes
    int t0[8];   // first local stage of each pass
    int r[8];    // stages per pass
};
HD PassPlan make_plan(int c) {
    PassPlan p;
    p.n = 0;
    int t = 0;
    int rem = c % 3;
    if (c > 0 && rem != 0) {
        p.t0[p.n] = 0;
        p.r[p.n] = rem;
        ++p.n;
        t = rem;
    }
    while (t < c) {
        p.t0[p.n] = t;
        p.r[p.n] = 3;
        ++p.n;
        t += 3;
    }
    return p;
}

// All groups of one forward pass handled by "thread" tid of nthr (the CUDA kernels call this with
// threadIdx.x / blockDim.x and __syncthreads() between passes; hostsim loops tid sequentially).
template <typename A>
HD void fwd_tile_pass(const A& m, typename A::W* s, int c, int t0, int r, int s0, uint32_t k, uint32_t tid, uint32_t nthr,
                      const TwPair<typename A::W>* __restrict__ tw) {
    const uint32_t groups = 1u << (c - r);
    for (uint32_t g = tid; g < groups; g += nthr) {
        if (r == 3)
            fwd_tile_group<A, 3>(m, s, c, t0, s0, k, g, tw);
        else if (r == 2)
            fwd_tile_group<A, 2>(m, s, c, t0, s0, k, g, tw);
        else
            fwd_tile_group<A, 1>(m, s, c, t0, s0, k, g, tw);
    }
}
template <typename A>
HD void inv_tile_pass(const A& m, typename A::W* s, int c, int t0, int r, int s0, uint32_t k, uint32_t tid, uint32_t nthr,
                      const TwPair<typename A::W>* __restrict__ itw, bool last, TwPair<typename A::W> ninv,
                      TwPair<typename A::W> wninv) {
    const uint32_t groups = 1u << (c - r);
    for (uint32_t g = tid; g < groups; g += nthr) {
        if (last) {
            if (r == 3)
                inv_tile_group<A, 3, true>(m, s, c, t0, s0, k, g, itw, ninv, wninv);
            else if (r == 2)
                inv_tile_group<A, 2, true>(m, s, c, t0, s0, k, g, itw, ninv, wninv);
            else
                inv_tile_group<A, 1, true>(m, s, c, t0, s0, k, g, itw, ninv, wninv);
        } else {
            if (r == 3)
                inv_tile_group<A, 3, false>(m, s, c, t0, s0, k, g, itw, ninv, wninv);
            else if (r == 2)
                inv_tile_group<A, 2, false>(m, s, c, t0, s0, k, g, itw, ninv, wninv);
            else
                inv_tile_group<A, 1, false>(m, s, c, t0, s0, k, g, itw, ninv, wninv);
        }
    }
}

// Column ("strided") transform of R = 2^S rows held in registers: the first S forward stages (or the
// last S inverse stages) of a polynomial viewed as a [2^S][N/2^S] row-major matrix.  Every column uses
// the same twiddles: index (1 << u) + top.
template <typename A, int S>
HD void fwd_column_regs(const A& m, typename A::W* x, const TwPair<typename A::W>* __restrict__ tw) {
    fwd_pass_regs<A, S>(m, x, tw, 1u);
}
template <typename A, int S>
HD void inv_column_regs(const A& m, typename A::W* x, const TwPair<typename A::W>* __restrict__ itw, TwPair<typename A::W> ninv,
                        TwPair<typename A::W> wninv) {
    inv_pass_regs<A, S, true>(m, x, itw, 1u, ninv, wninv);
}

}  // namespace fhe
