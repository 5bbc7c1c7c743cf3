// Batched negacyclic NTT / iNTT kernels (K1/K2 of SURVEY.md §2), sm_100a.
// Replaces util/src/ring/fft/zq.rs:27-36 (+ ring/fft.rs:40-77).
//
//   ntt_tile_kernel   : one CTA transforms one tile (2^c contiguous coefficients, c <= 13) entirely in shared
//                       memory; persistent CTAs loop over (polynomial, tile) work items.  For N = 2^c the tile is
//                       the whole polynomial.  128-bit coalesced global loads/stores, swizzled smem tile,
//                       radix-8 register passes, Shoup twiddles (w, w') fetched as 8/16-byte pairs.
//   ntt_column_kernel : for N > 2^c, the 2^S-point column transforms (S = logN - c <= 4) that precede (forward)
//                       or follow (inverse) the tile kernel; pure register kernel, uniform twiddles.
#pragma once
#include <cuda_runtime.h>

#include "ntt_core.cuh"

namespace fhe {

template <typename W>
struct Vec16;
template <>
struct Vec16<uint32_t> {
    typedef uint4 T;
    static constexpr int N = 4;
};
template <>
struct Vec16<uint64_t> {
    typedef ulonglong2 T;
    static constexpr int N = 2;
};

template <typename A>
struct NttArgs {
    typename A::W* data;   // in place
    const TwPair<typename A::W>* tw;  // forward or inverse table (bit-reversed order, length >= N)
    A m;
    int log_n;
    int c;                 // tile = 2^c coefficients
    unsigned long long n_items;  // batch * 2^(log_n - c)
    TwPair<typename A::W> ninv, wninv;
};

template <typename A, bool FWD>
__global__ void __launch_bounds__(1024) ntt_tile_kernel(NttArgs<A> a) {
    typedef typename A::W W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W* s = reinterpret_cast<W*>(smem_raw);
    const int c = a.c, s0 = a.log_n - a.c;
    const uint32_t C = 1u << c;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const PassPlan plan = make_plan(c);
    constexpr int VN = Vec16<W>::N;
    typedef typename Vec16<W>::T V;
    const bool final_out = FWD || s0 == 0;
    for (unsigned long long item = blockIdx.x; item < a.n_items; item += gridDim.x) {
        const uint32_t k = (uint32_t)(item & ((1ull << s0) - 1ull));
        const unsigned long long b = item >> s0;
        W* g = a.data + (b << a.log_n) + ((unsigned long long)k << c);
        if (C >= (uint32_t)VN) {
            const V* gv = reinterpret_cast<const V*>(g);
            for (uint32_t i = tid; i < C / VN; i += nthr) {
                V v = gv[i];
                const W* e = reinterpret_cast<const W*>(&v);
#pragma unroll
                for (int j = 0; j < VN; ++j) s[swz<W>(i * VN + j)] = e[j];
            }
        } else {
            for (uint32_t i = tid; i < C; i += nthr) s[swz<W>(i)] = g[i];
        }
        __syncthreads();
        if (FWD) {
            for (int pi = 0; pi < plan.n; ++pi) {
                fwd_tile_pass<A>(a.m, s, c, plan.t0[pi], plan.r[pi], s0, k, tid, nthr, a.tw);
                __syncthreads();
            }
        } else {
            for (int pi = plan.n - 1; pi >= 0; --pi) {
                const bool last = (s0 == 0) && (plan.t0[pi] == 0);
                inv_tile_pass<A>(a.m, s, c, plan.t0[pi], plan.r[pi], s0, k, tid, nthr, a.tw, last, a.ninv, a.wninv);
                __syncthreads();
            }
        }
        if (C >= (uint32_t)VN) {
            V* gv = reinterpret_cast<V*>(g);
            for (uint32_t i = tid; i < C / VN; i += nthr) {
                V v;
                W* e = reinterpret_cast<W*>(&v);
#pragma unroll
                for (int j = 0; j < VN; ++j) {
                    W x = s[swz<W>(i * VN + j)];
                    if (final_out) x = FWD ? a.m.canon4(x) : a.m.redq(x);
                    e[j] = x;
                }
                gv[i] = v;
            }
        } else {
            for (uint32_t i = tid; i < C; i += nthr) {
                W x = s[swz<W>(i)];
                if (final_out) x = FWD ? a.m.canon4(x) : a.m.redq(x);
                g[i] = x;
            }
        }
        __syncthreads();
    }
}

template <typename A, int S, bool FWD>
__global__ void __launch_bounds__(256) ntt_column_kernel(NttArgs<A> a, unsigned long long batch) {
    typedef typename A::W W;
    const int lc = a.log_n - S;  // log2(columns)
    const unsigned long long total = batch << lc;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const unsigned long long b = idx >> lc;
        const uint32_t col = (uint32_t)(idx & ((1ull << lc) - 1ull));
        W* g = a.data + (b << a.log_n) + col;
        W x[1 << S];
#pragma unroll
        for (int j = 0; j < (1 << S); ++j) x[j] = g[(unsigned long long)j << lc];
        if (FWD) {
            fwd_column_regs<A, S>(a.m, x, a.tw);
        } else {
            inv_column_regs<A, S>(a.m, x, a.tw, a.ninv, a.wninv);
#pragma unroll
            for (int j = 0; j < (1 << S); ++j) x[j] = a.m.redq(x[j]);
        }
#pragma unroll
        for (int j = 0; j < (1 << S); ++j) g[(unsigned long long)j << lc] = x[j];
    }
}

}  // namespace fhe
