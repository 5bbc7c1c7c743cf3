// Fast-path batched negacyclic NTT kernels + launcher (see ntt_fast.cuh for the design).
//   ntt_fast_tile_kernel   : PB tiles per CTA (TPP threads each); first pass global->regs->smem, radix-8 passes in
//                            shared memory, last pass smem->regs->global.  One work item = (tile k, polynomial b),
//                            k-major so that CTAs running together share twiddle lines in L1/L2.
//   ntt_fast_column_kernel : register-only 2^S-point column transforms for N > 2^13 (first S forward stages / last S
//                            inverse stages), coalesced along the row direction.
// Replaces util/src/ring/fft/zq.rs:27-36 (+ ring/fft.rs:40-77) for every caller: fhe_ntt_*, CKKS limb batches, key
// upload.  Moduli outside the lazy-reduction preconditions fall back to the generic kernels (ntt_launch.cu).
#include <atomic>
#include <cstdlib>
#include <algorithm>

#include "ctx.cuh"
#include "ntt_fast.cuh"

namespace fhe {

template <typename L>
struct FastArgs {
    typename L::W* data;        // destination (and source unless `in` differs)
    const typename L::W* in;    // source of the kernel launched first; == data for in-place transforms
    const FastLimb<L>* limbs;
    uint32_t nl;
    uint32_t limb0;             // limb of polynomial 0 of this launch (chunked launches keep limb = (limb0 + b) % nl)
    uint32_t n_polys;
    int log_n;
    int s0;
    uint32_t pre_red;
    unsigned long long n_items;  // n_polys << s0
};


// resident threads per SM the tile kernels are compiled for (register cap = 65536 / this): u64 1024 (64 registers),
// u32 1280 (48 registers; measured best for the paired-group passes: 1024 and 1536 are 3-8 % slower on some sizes)
#ifndef FAST_TMA_DEFAULT
#define FAST_TMA_DEFAULT 0
#endif
#ifndef FAST_OCC32
#define FAST_OCC32 1280
#endif
#ifndef FAST_OCC64
#define FAST_OCC64 1024
#endif
// measured exceptions for 64-bit words: the forward 2^11 tile (also the tile of N = 2^14, 2^15) is 9-12 % faster at 51 registers,
// the 2^10 tile 3-5 % faster at 85
#ifndef FAST_OCC64_R16
#define FAST_OCC64_R16 768  // the radix-16 plan of the 2^12 tile (FAST_R16_64): 16 values + twiddles per thread
#endif
template <typename L, int LOGT, bool FWD>
struct FastOcc {
    static constexpr int value = L::BITS == 32 ? FAST_OCC32
                                               : (FastGeom<L, LOGT>::RM == 4 ? FAST_OCC64_R16 : (LOGT == 11 && FWD ? 1280 : (LOGT == 10 ? 768 : FAST_OCC64)));
};
template <typename L, int LOGT, bool FWD, bool FINAL>
__global__ void __launch_bounds__(FastGeom<L, LOGT>::NTHR, FastOcc<L, LOGT, FWD>::value / FastGeom<L, LOGT>::NTHR)
ntt_fast_tile_kernel(FastArgs<L> a) {
    typedef typename L::W W;
    typedef FastGeom<L, LOGT> G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t pslot = threadIdx.x / G::TPP, tid = threadIdx.x % G::TPP;
    W* s = reinterpret_cast<W*>(smem_raw) + ((size_t)pslot << LOGT);
    const unsigned long long item = (unsigned long long)blockIdx.x * G::PB + pslot;
    const bool active = item < a.n_items;
    uint32_t k = 0, b = 0;
    if (active) {
        k = (uint32_t)(item / a.n_polys);
        b = (uint32_t)(item - (unsigned long long)k * a.n_polys);
    }
    const FastLimb<L>& d = a.limbs[(a.limb0 + b) % a.nl];
    const size_t off = ((size_t)b << a.log_n) + ((size_t)k << LOGT);
    W* g = a.data + off;
    const W* gin = a.in + off;
    if (FWD) {
        if (active) fast_fwd_first_any<L, LOGT>(d, gin, s, a.s0, k, (a.pre_red & 1u) != 0, tid);
        __syncthreads();
        if (G::NP3 > 1) {
            if (active) fast_fwd_mid_any<L, LOGT, G::R1>(d, s, a.s0, k, ((a.pre_red >> 1) & 1u) != 0, tid);
            __syncthreads();
        }
        if (G::NP3 > 2) {
            if (active) fast_fwd_mid_any<L, LOGT, (G::NP3 > 2 ? G::R1 + G::RM : G::R1)>(d, s, a.s0, k, ((a.pre_red >> 2) & 1u) != 0, tid);
            __syncthreads();
        }
        static_assert(G::NP3 <= 3, "at most two middle passes");
        if (active) fast_fwd_last<L, LOGT, G::TPP, G::RM>(d, s, g, a.s0, k, ((a.pre_red >> G::NP3) & 1u) != 0, tid);
    } else {
        if (active) fast_inv_first<L, LOGT, G::TPP, G::RM>(d, gin, s, a.s0, k, tid);
        __syncthreads();
        if (G::NP3 > 2) {
            if (active) fast_inv_mid_any<L, LOGT, (G::NP3 > 2 ? G::R1 + G::RM : G::R1)>(d, s, a.s0, k, tid);
            __syncthreads();
        }
        if (G::NP3 > 1) {
            if (active) fast_inv_mid_any<L, LOGT, G::R1>(d, s, a.s0, k, tid);
            __syncthreads();
        }
        if (active) fast_inv_last_any<L, LOGT, FINAL>(d, s, g, a.s0, k, tid);
    }
}

// ---- TMA-staged persistent tile kernel ------------------------------------------------------------------------------------------------
// Same passes as ntt_fast_tile_kernel, but the CTA is persistent over blocks of PB tiles and the NEXT block's coefficients
// are brought from HBM into a staging buffer by the bulk-copy engine (cp.async.bulk global -> shared, completion counted on an
// mbarrier) while the current block is being transformed: the first pass reads the staged tile from shared memory instead
// of waiting on global loads, so HBM latency is hidden by a copy that needs no registers and no warps, not by occupancy.
// Shared memory: work[PB << LOGT] | stage[PB << LOGT] | mbarrier.
DEV uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
DEV void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
DEV void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
DEV void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
DEV void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "NTT_TMA_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra NTT_TMA_DONE;\n"
        "bra NTT_TMA_WAIT;\n"
        "NTT_TMA_DONE:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
template <typename L, int LOGT, bool FWD>
struct FastTmaOcc {  // resident threads per SM the TMA variant is compiled for (shared memory: 2 tiles per slot)
    static constexpr int value = L::BITS == 32 ? 1024 : (LOGT >= 12 ? 768 : 1024);
};
template <typename L, int LOGT, bool FWD, bool FINAL>
__global__ void __launch_bounds__(FastGeom<L, LOGT>::NTHR, FastTmaOcc<L, LOGT, FWD>::value / FastGeom<L, LOGT>::NTHR)
ntt_fast_tile_tma_kernel(FastArgs<L> a, unsigned long long n_blocks) {
    typedef typename L::W W;
    typedef FastGeom<L, LOGT> G;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr uint32_t TILE_BYTES = (uint32_t)sizeof(W) << LOGT;
    const uint32_t pslot = threadIdx.x / G::TPP, tid = threadIdx.x % G::TPP;
    W* work = reinterpret_cast<W*>(smem_raw) + ((size_t)pslot << LOGT);
    W* stage_all = reinterpret_cast<W*>(smem_raw) + ((size_t)G::PB << LOGT);
    const W* stage = stage_all + ((size_t)pslot << LOGT);
    uint64_t* bar = reinterpret_cast<uint64_t*>(stage_all + ((size_t)G::PB << LOGT));
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    auto issue = [&](unsigned long long blk) {  // one thread: request the PB tiles of block `blk`
        const unsigned long long first = blk * G::PB;
        const uint32_t cnt = (uint32_t)min((unsigned long long)G::PB, a.n_items - first);
        mbar_expect_tx(bar, cnt * TILE_BYTES);
        for (uint32_t p = 0; p < cnt; ++p) {
            const unsigned long long it = first + p;
            const uint32_t kk = (uint32_t)(it / a.n_polys), bb = (uint32_t)(it - (unsigned long long)kk * a.n_polys);
            bulk_g2s(stage_all + ((size_t)p << LOGT), a.in + (((size_t)bb << a.log_n) + ((size_t)kk << LOGT)), TILE_BYTES, bar);
        }
    };
    unsigned long long blk = blockIdx.x;
    if (blk < n_blocks && threadIdx.x == 0) issue(blk);
    uint32_t parity = 0;
    for (; blk < n_blocks; blk += gridDim.x) {
        const unsigned long long item = blk * G::PB + pslot;
        const bool active = item < a.n_items;
        uint32_t k = 0, b = 0;
        if (active) {
            k = (uint32_t)(item / a.n_polys);
            b = (uint32_t)(item - (unsigned long long)k * a.n_polys);
        }
        const FastLimb<L>& d = a.limbs[(a.limb0 + b) % a.nl];
        W* g = a.data + (((size_t)b << a.log_n) + ((size_t)k << LOGT));
        mbar_wait(bar, parity);
        parity ^= 1u;
        if (FWD) {
            if (active) fast_fwd_first_any<L, LOGT>(d, stage, work, a.s0, k, (a.pre_red & 1u) != 0, tid);
        } else {
            if (active) fast_inv_first<L, LOGT, G::TPP, G::RM>(d, stage, work, a.s0, k, tid);
        }
        __syncthreads();  // the staged block is consumed: the copy engine may refill the buffer
        if (blk + gridDim.x < n_blocks && threadIdx.x == 0) issue(blk + gridDim.x);
        if (FWD) {
            if (G::NP3 > 1) {
                if (active) fast_fwd_mid_any<L, LOGT, G::R1>(d, work, a.s0, k, ((a.pre_red >> 1) & 1u) != 0, tid);
                __syncthreads();
            }
            if (G::NP3 > 2) {
                if (active) fast_fwd_mid_any<L, LOGT, (G::NP3 > 2 ? G::R1 + G::RM : G::R1)>(d, work, a.s0, k, ((a.pre_red >> 2) & 1u) != 0, tid);
                __syncthreads();
            }
            if (active) fast_fwd_last<L, LOGT, G::TPP, G::RM>(d, work, g, a.s0, k, ((a.pre_red >> G::NP3) & 1u) != 0, tid);
        } else {
            if (G::NP3 > 2) {
                if (active) fast_inv_mid_any<L, LOGT, (G::NP3 > 2 ? G::R1 + G::RM : G::R1)>(d, work, a.s0, k, tid);
                __syncthreads();
            }
            if (G::NP3 > 1) {
                if (active) fast_inv_mid_any<L, LOGT, G::R1>(d, work, a.s0, k, tid);
                __syncthreads();
            }
            if (active) fast_inv_last_any<L, LOGT, FINAL>(d, work, g, a.s0, k, tid);
        }
        __syncthreads();  // work is rewritten by the next block's first pass
    }
}
#ifndef FAST_COL_MINB
#define FAST_COL_MINB 3  // 80 registers, 3 CTAs per SM: measured best for the HBM-bound column pass (4: spills; unbounded: 114-124 registers)
#endif
template <typename L, int S, bool FWD>
__global__ void __launch_bounds__(256, FAST_COL_MINB) ntt_fast_column_kernel(FastArgs<L> a) {
    const int lc = a.log_n - S;
    const unsigned long long total = (unsigned long long)a.n_polys << lc;
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const uint32_t b = (uint32_t)(idx >> lc);
        const uint32_t col = (uint32_t)(idx & ((1ull << lc) - 1ull));
        const FastLimb<L>& d = a.limbs[(a.limb0 + b) % a.nl];
        typename L::W* g = a.data + ((size_t)b << a.log_n);
        const typename L::W* gin = a.in + ((size_t)b << a.log_n);
        if (FWD)
            fast_fwd_column<L, S>(d, gin, g, lc, col);
        else
            fast_inv_column<L, S>(d, gin, g, lc, col);
    }
}

// ---- host side ---------------------------------------------------------------------------------------------------------
template <typename L>
L make_lazy(uint64_t q);
template <>
Lz32 make_lazy<Lz32>(uint64_t q) {
    return make_lz32(q);
}
template <>
Lz64 make_lazy<Lz64>(uint64_t q) {
    return make_lz64(q);
}

template <typename L>
static bool fast_modulus_ok(uint64_t q) {
    return L::BITS == 32 ? (q > 2 && q < (1ull << 28)) : (q > 2 && q < (1ull << 56));
}

// device table of limb descriptors for (qs, log_n), cached in the context
template <typename L>
static fhe_status get_fast_limbs(fhe_ctx* ctx, const uint64_t* qs, size_t nl, unsigned log_n, const FastLimb<L>** out) {
    typedef typename L::W W;
    std::vector<uint64_t> key(qs, qs + nl);
    key.push_back(((uint64_t)log_n << 8) | (uint64_t)L::BITS);
    std::lock_guard<std::mutex> lock(ctx->mu);
    auto it = ctx->fast_limbs.find(key);
    if (it != ctx->fast_limbs.end()) {
        *out = (const FastLimb<L>*)it->second;
        return FHE_OK;
    }
    std::vector<FastLimb<L>> h(nl);
    for (size_t i = 0; i < nl; ++i) {
        const NttTable* t;
        FHE_CHECK(get_ntt_table_locked(ctx, qs[i], L::BITS, (size_t)1 << log_n, &t));
        h[i].m = make_lazy<L>(qs[i]);
        h[i].tw = (const TwPair<W>*)t->d_fwd;
        h[i].itw = (const TwPair<W>*)t->d_inv;
        const uint64_t ninv = host_invmod(((uint64_t)1 << log_n) % qs[i], qs[i]);
        h[i].ninv = make_twpair<W>(ninv, qs[i]);
        h[i].wninv = make_twpair<W>(host_mulmod(t->h_inv[1], ninv, qs[i]), qs[i]);
    }
    void* d = nullptr;
    FHE_CUDA(ctx, cudaMalloc(&d, nl * sizeof(FastLimb<L>)));
    FHE_CUDA(ctx, cudaMemcpy(d, h.data(), nl * sizeof(FastLimb<L>), cudaMemcpyHostToDevice));
    ctx->fast_limbs[key] = d;
    *out = (const FastLimb<L>*)d;
    return FHE_OK;
}

// which tile kernel: FHE_B200_NTT_TMA = 0 (direct global loads, one tile block per CTA), 1 (TMA-staged persistent, both word
// sizes), 32 / 64 (TMA-staged for that word size only).  Measured on B200 (r02, 4096 polynomials, fwd GB/s, direct -> TMA):
// u32 2^10 2020 -> 1818, 2^11 2486 -> 2436, 2^12 2719 -> 2414, 2^13 2636 -> 2366; u64 2^10 1932 -> 1616, 2^12 1841 -> 1691,
// 2^13 1730 -> 1189: the staged variant LOSES 2-31 % everywhere.  The staging buffer doubles the shared memory per tile
// (fewer resident CTAs) and turns the first pass's global loads into shared-memory loads on the unit that is already the
// busiest (LSU wavefronts 67-73 % for u32), while the many small CTAs of the direct kernel already overlap their load,
// compute and store phases.  The direct kernel therefore stays the default; the staged one is kept selectable.
static bool fast_use_tma(int bits) {
    static const int mode = [] {
        const char* e = getenv("FHE_B200_NTT_TMA");
        return e ? atoi(e) : FAST_TMA_DEFAULT;
    }();
    return mode == 1 || mode == bits;
}
template <typename L, int LOGT, bool FWD, bool FINAL>
static fhe_status launch_fast_tile(fhe_ctx* ctx, FastArgs<L>& a) {
    typedef FastGeom<L, LOGT> G;
    auto kern = ntt_fast_tile_kernel<L, LOGT, FWD, FINAL>;
    const size_t smem = (size_t)G::PB * sizeof(typename L::W) << LOGT;
    // the attribute is per device: remember which devices of this process already have it (one bit per device ordinal)
    static std::atomic<uint64_t> attr_done{0};
    const uint64_t dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_done.load(std::memory_order_relaxed) & dev_bit)) {
        FHE_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        attr_done.fetch_or(dev_bit, std::memory_order_relaxed);
    }
    const unsigned long long grid = (a.n_items + G::PB - 1) / G::PB;
    if (grid == 0) return FHE_OK;
    if (grid > 0x7FFFFFFFull) return fail(ctx, FHE_EINVAL, "batch too large for one launch");
    if (fast_use_tma(L::BITS)) {
        auto tk = ntt_fast_tile_tma_kernel<L, LOGT, FWD, FINAL>;
        const size_t tsmem = 2 * smem + 16;
        static std::atomic<uint64_t> tattr_done{0};
        static int occ = 0;
        if (!(tattr_done.load(std::memory_order_relaxed) & dev_bit)) {
            FHE_CUDA(ctx, cudaFuncSetAttribute(tk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tsmem));
            FHE_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, tk, G::NTHR, tsmem));
            tattr_done.fetch_or(dev_bit, std::memory_order_relaxed);
        }
        if (occ >= 1) {
            // blocks per persistent CTA the grid is sized for: with one block each (a batch that fits one wave) nothing is pipelined
            static const int depth = [] {
                const char* e = getenv("FHE_B200_NTT_TMA_DEPTH");
                return e ? std::max(1, atoi(e)) : 1;
            }();
            const unsigned tgrid = (unsigned)std::min<unsigned long long>((grid + depth - 1) / depth, (unsigned long long)ctx->sm_count * occ);
            tk<<<tgrid, G::NTHR, tsmem, ctx->stream>>>(a, grid);
            return after_launch(ctx, FWD ? "ntt_fast_tile_tma_kernel<fwd>" : "ntt_fast_tile_tma_kernel<inv>");
        }
    }
    kern<<<(unsigned)grid, G::NTHR, smem, ctx->stream>>>(a);
    return after_launch(ctx, FWD ? "ntt_fast_tile_kernel<fwd>" : "ntt_fast_tile_kernel<inv>");
}
template <typename L, bool FWD, bool FINAL>
static fhe_status launch_fast_tile_logt(fhe_ctx* ctx, FastArgs<L>& a, int logt) {
    switch (logt) {
        case 9: return launch_fast_tile<L, 9, FWD, FINAL>(ctx, a);
        case 10: return launch_fast_tile<L, 10, FWD, FINAL>(ctx, a);
        case 11: return launch_fast_tile<L, 11, FWD, FINAL>(ctx, a);
        case 12: return launch_fast_tile<L, 12, FWD, FINAL>(ctx, a);
        case 13: return launch_fast_tile<L, 13, FWD, FINAL>(ctx, a);
        default: return fail(ctx, FHE_EUNSUPPORTED, "fast NTT: tile size 2^%d not instantiated", logt);
    }
}
template <typename L, int S, bool FWD>
static fhe_status launch_fast_column_s(fhe_ctx* ctx, FastArgs<L>& a) {
    const unsigned long long total = (unsigned long long)a.n_polys << (a.log_n - S);
    const unsigned long long grid = std::min<unsigned long long>((total + 255) / 256, (unsigned long long)ctx->sm_count * 32);
    if (grid == 0) return FHE_OK;
    ntt_fast_column_kernel<L, S, FWD><<<(unsigned)grid, 256, 0, ctx->stream>>>(a);
    return after_launch(ctx, FWD ? "ntt_fast_column_kernel<fwd>" : "ntt_fast_column_kernel<inv>");
}
template <typename L, bool FWD>
static fhe_status launch_fast_column(fhe_ctx* ctx, FastArgs<L>& a, int S) {
    switch (S) {
        case 1: return launch_fast_column_s<L, 1, FWD>(ctx, a);
        case 2: return launch_fast_column_s<L, 2, FWD>(ctx, a);
        case 3: return launch_fast_column_s<L, 3, FWD>(ctx, a);
        case 4: return launch_fast_column_s<L, 4, FWD>(ctx, a);
        default: return fail(ctx, FHE_EUNSUPPORTED, "fast NTT: column transform of 2^%d rows not instantiated", S);
    }
}

// returns FHE_EUNSUPPORTED (without touching ctx->err) when the fast path does not apply
template <typename L>
static fhe_status launch_ntt_fast(fhe_ctx* ctx, const uint64_t* qs, size_t nl, unsigned log_n, size_t n_polys,
                                  const typename L::W* d_src, typename L::W* d_a, bool fwd) {
    if (log_n < 9 || log_n > 17 || nl == 0) return FHE_EUNSUPPORTED;
    for (size_t i = 0; i < nl; ++i)
        if (!fast_modulus_ok<L>(qs[i])) return FHE_EUNSUPPORTED;
    if (n_polys == 0) return FHE_OK;
    if (n_polys > 0x7FFFFFFFull) return fail(ctx, FHE_EINVAL, "batch too large");
    int logt = fast_tile_logt((int)log_n);
    if (const char* e = getenv("FHE_B200_NTT_LOGT")) {  // tuning knob: tile size for rings larger than one tile
        const int v = atoi(e);
        if (log_n > 13 && v >= 9 && v <= 13 && (int)log_n - v <= 4) logt = v;
    }
    const int S = (int)log_n - logt;
    FastArgs<L> a;
    a.data = d_a;
    a.in = d_src ? d_src : d_a;
    FHE_CHECK(get_fast_limbs<L>(ctx, qs, nl, log_n, &a.limbs));
    a.nl = (uint32_t)nl;
    a.limb0 = 0;
    a.n_polys = (uint32_t)n_polys;
    a.log_n = (int)log_n;
    a.s0 = S;
    a.n_items = (unsigned long long)n_polys << S;
    a.pre_red = 0;
    if (L::BITS == 32 && fwd) {
        const int mask = fast_plan_prered32(logt, 1 + 2 * S);
        if (mask < 0) return FHE_EUNSUPPORTED;
        a.pre_red = (uint32_t)mask;
    }
    if (S == 0) return fwd ? launch_fast_tile_logt<L, true, true>(ctx, a, logt) : launch_fast_tile_logt<L, false, true>(ctx, a, logt);
    // Rings larger than one tile take two kernels (column pass + tile pass) with a round trip through memory between them
    // (r01 ncu: 972 MB of DRAM traffic for 537 MB algorithmic at 512 x 2^16 u64).  FHE_B200_NTT_L2_MB = m walks the batch in
    // chunks of m MiB so that the intermediate polynomials are still L2-resident when the second kernel reads them and are
    // overwritten there before eviction (DRAM then sees one read and one write per polynomial).  Measured on B200 (r02,
    // 4096 polynomials, 48 MiB chunks): SLOWER - 2^16 u64 1.20 instead of 1.50 TB/s, 2^14 u32 1.85 instead of 2.09 TB/s - the
    // transform is bound by the integer pipe, not by DRAM, and 128 short launches cost more in ramp-up / tail than the saved
    // traffic; so the default is one launch pair for the whole batch.
    static const size_t l2_budget = [] {
        if (const char* e = getenv("FHE_B200_NTT_L2_MB")) return (size_t)std::max(1, atoi(e)) << 20;  // tuning knob
        return ~(size_t)0 >> 1;
    }();
    const size_t poly_bytes = sizeof(typename L::W) << log_n;
    const size_t per_chunk = std::max<size_t>(1, l2_budget / (poly_bytes * (d_src && d_src != d_a ? 2 : 1)));
    for (size_t p0 = 0; p0 < n_polys; p0 += per_chunk) {
        const size_t cnt = std::min(per_chunk, n_polys - p0);
        FastArgs<L> c = a;
        c.data = d_a + (p0 << log_n);
        c.in = (d_src ? d_src : d_a) + (p0 << log_n);
        c.n_polys = (uint32_t)cnt;
        c.n_items = (unsigned long long)cnt << S;
        // limb of polynomial b is limbs[b % nl]: keep the phase when a chunk does not start at a multiple of nl
        c.limb0 = (uint32_t)(p0 % nl);
        if (fwd) {
            FHE_CHECK((launch_fast_column<L, true>(ctx, c, S)));
            c.in = c.data;
            FHE_CHECK((launch_fast_tile_logt<L, true, true>(ctx, c, logt)));
        } else {
            FHE_CHECK((launch_fast_tile_logt<L, false, false>(ctx, c, logt)));
            c.in = c.data;
            FHE_CHECK((launch_fast_column<L, false>(ctx, c, S)));
        }
    }
    return FHE_OK;
}

fhe_status launch_ntt_fast_u64(fhe_ctx* ctx, const uint64_t* qs, size_t nl, unsigned log_n, size_t n_polys, const uint64_t* d_src,
                               uint64_t* d_a, bool fwd) {
    return launch_ntt_fast<Lz64>(ctx, qs, nl, log_n, n_polys, d_src, d_a, fwd);
}
fhe_status launch_ntt_fast_u32(fhe_ctx* ctx, const uint64_t* qs, size_t nl, unsigned log_n, size_t n_polys, const uint32_t* d_src,
                               uint32_t* d_a, bool fwd) {
    return launch_ntt_fast<Lz32>(ctx, qs, nl, log_n, n_polys, d_src, d_a, fwd);
}

}  // namespace fhe
