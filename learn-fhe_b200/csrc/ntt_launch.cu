// Launch logic + C ABI for the batched negacyclic NTT (fhe_ntt_*), host-slice wrappers included.
#include <atomic>
#include <algorithm>
#include <cstdlib>

#include "ctx.cuh"
#include "ntt_kernels.cuh"

namespace fhe {

// tile size policy: whole polynomial per CTA up to 2^CMAX coefficients; above that 2^CTILE tiles + column kernel
template <typename W>
struct TilePolicy;
template <>
struct TilePolicy<uint64_t> {
    static constexpr int CMAX = 13;   // 64 KiB tile
    static constexpr int CTILE = 12;  // 32 KiB tiles when split
};
template <>
struct TilePolicy<uint32_t> {
    static constexpr int CMAX = 13;   // 32 KiB tile
    static constexpr int CTILE = 13;
};

template <typename A, bool FWD>
static fhe_status launch_tile(fhe_ctx* ctx, NttArgs<A>& a) {
    typedef typename A::W W;
    const size_t smem = sizeof(W) << a.c;
    int threads = std::max(32, std::min(1024, (1 << a.c) / 8));
    auto kern = ntt_tile_kernel<A, FWD>;
    // the attribute is per device: remember which devices of this process already have it (one bit per device ordinal)
    static std::atomic<uint64_t> attr_done{0};
    const uint64_t dev_bit = 1ull << (ctx->device & 63);
    if (!(attr_done.load(std::memory_order_relaxed) & dev_bit)) {
        FHE_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_done.fetch_or(dev_bit, std::memory_order_relaxed);
    }
    int occ = 1;
    FHE_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) occ = 1;
    unsigned long long grid = std::min<unsigned long long>(a.n_items, (unsigned long long)ctx->sm_count * occ);
    if (grid == 0) return FHE_OK;
    kern<<<(unsigned)grid, threads, smem, ctx->stream>>>(a);
    return after_launch(ctx, "ntt_tile_kernel");
}

template <typename A, int S, bool FWD>
static fhe_status launch_column_s(fhe_ctx* ctx, NttArgs<A>& a, size_t batch) {
    unsigned long long total = (unsigned long long)batch << (a.log_n - S);
    int threads = 256;
    unsigned long long grid = std::min<unsigned long long>((total + threads - 1) / threads, (unsigned long long)ctx->sm_count * 16);
    if (grid == 0) return FHE_OK;
    ntt_column_kernel<A, S, FWD><<<(unsigned)grid, threads, 0, ctx->stream>>>(a, batch);
    return after_launch(ctx, "ntt_column_kernel");
}
template <typename A, bool FWD>
static fhe_status launch_column(fhe_ctx* ctx, NttArgs<A>& a, size_t batch, int S) {
    switch (S) {
        case 1: return launch_column_s<A, 1, FWD>(ctx, a, batch);
        case 2: return launch_column_s<A, 2, FWD>(ctx, a, batch);
        case 3: return launch_column_s<A, 3, FWD>(ctx, a, batch);
        case 4: return launch_column_s<A, 4, FWD>(ctx, a, batch);
        default: return fail(ctx, FHE_EUNSUPPORTED, "column transform of 2^%d rows not supported", S);
    }
}

template <typename A>
static fhe_status launch_ntt(fhe_ctx* ctx, uint64_t q, unsigned log_n, size_t batch, typename A::W* d_a, bool fwd) {
    typedef typename A::W W;
    FHE_REQUIRE(ctx, log_n <= 17, "log_n %u too large (max 17)", log_n);
    if (batch == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_a != nullptr, "null data pointer");
    if (log_n == 0) {  // degree-1 ring: both transforms are the identity (n^-1 = 1); only the modulus is checked
        std::lock_guard<std::mutex> lock(ctx->mu);
        const ModInfo* mi;
        return get_mod_info(ctx, q, &mi);
    }
    const NttTable* t;
    FHE_CHECK(get_ntt_table(ctx, q, A::BITS, (size_t)1 << log_n, &t));
    NttArgs<A> a;
    a.data = d_a;
    a.m = make_mod<A>(q);
    a.log_n = (int)log_n;
    a.c = (int)log_n <= TilePolicy<W>::CMAX ? (int)log_n : TilePolicy<W>::CTILE;
    a.n_items = (unsigned long long)batch << (log_n - a.c);
    uint64_t ninv = host_invmod(((uint64_t)1 << log_n) % q, q);
    a.ninv = make_twpair<W>(ninv, q);
    a.wninv = make_twpair<W>(host_mulmod(t->h_inv[1], ninv, q), q);
    const int S = (int)log_n - a.c;
    if (fwd) {
        a.tw = (const TwPair<W>*)t->d_fwd;
        if (S > 0) FHE_CHECK((launch_column<A, true>(ctx, a, batch, S)));
        return launch_tile<A, true>(ctx, a);
    } else {
        a.tw = (const TwPair<W>*)t->d_inv;
        FHE_CHECK((launch_tile<A, false>(ctx, a)));
        if (S > 0) return launch_column<A, false>(ctx, a, batch, S);
        return FHE_OK;
    }
}

static bool g_force_generic = false;  // FHE_B200_NTT_GENERIC=1: always use the first-generation kernels (A/B tests)
static bool use_fast() {
    static bool init = false;
    if (!init) {
        const char* e = getenv("FHE_B200_NTT_GENERIC");
        g_force_generic = e && e[0] == '1';
        init = true;
    }
    return !g_force_generic;
}

fhe_status launch_ntt_u64(fhe_ctx* ctx, uint64_t q, unsigned log_n, size_t batch, uint64_t* d_a, bool fwd) {
    if (use_fast()) {
        fhe_status st = launch_ntt_fast_u64(ctx, &q, 1, log_n, batch, nullptr, d_a, fwd);
        if (st != FHE_EUNSUPPORTED) return st;
    }
    return launch_ntt<Mod64>(ctx, q, log_n, batch, d_a, fwd);
}
fhe_status launch_ntt_u32(fhe_ctx* ctx, uint32_t q, unsigned log_n, size_t batch, uint32_t* d_a, bool fwd) {
    if (use_fast()) {
        uint64_t q64 = q;
        fhe_status st = launch_ntt_fast_u32(ctx, &q64, 1, log_n, batch, nullptr, d_a, fwd);
        if (st != FHE_EUNSUPPORTED) return st;
    }
    return launch_ntt<Mod32>(ctx, q, log_n, batch, d_a, fwd);
}
fhe_status launch_ntt_rns_u64(fhe_ctx* ctx, const uint64_t* qs, size_t nl, unsigned log_n, size_t n_polys, uint64_t* d_a, bool fwd) {
    return launch_ntt_rns_u64_oop(ctx, qs, nl, log_n, n_polys, nullptr, d_a, fwd);
}
fhe_status launch_ntt_rns_u64_oop(fhe_ctx* ctx, const uint64_t* qs, size_t nl, unsigned log_n, size_t n_polys, const uint64_t* d_src,
                                  uint64_t* d_a, bool fwd) {
    if (n_polys == 0) return FHE_OK;
    FHE_REQUIRE(ctx, nl > 0 && n_polys % nl == 0, "polynomial count %zu is not a multiple of the limb count %zu", n_polys, nl);
    if (use_fast()) {
        fhe_status st = launch_ntt_fast_u64(ctx, qs, nl, log_n, n_polys, d_src, d_a, fwd);
        if (st != FHE_EUNSUPPORTED) return st;
    }
    if (d_src && d_src != d_a)
        FHE_CUDA(ctx, cudaMemcpyAsync(d_a, d_src, (n_polys << log_n) * sizeof(uint64_t), cudaMemcpyDeviceToDevice, ctx->stream));
    // generic kernels take one modulus per launch
    for (size_t b = 0; b < n_polys / nl; ++b)
        for (size_t i = 0; i < nl; ++i) FHE_CHECK(launch_ntt<Mod64>(ctx, qs[i], log_n, 1, d_a + ((b * nl + i) << log_n), fwd));
    return FHE_OK;
}

static bool pow2_log(size_t n, unsigned* lg) {
    if (n == 0 || (n & (n - 1))) return false;
    unsigned l = 0;
    while (((size_t)1 << l) < n) ++l;
    *lg = l;
    return true;
}

static fhe_status ntt_host(fhe_ctx* ctx, uint64_t q, uint64_t* a, size_t n, size_t batch, bool fwd) {
    unsigned lg;
    FHE_REQUIRE(ctx, pow2_log(n, &lg), "polynomial length %zu is not a power of two", n);
    if (batch == 0) return FHE_OK;
    size_t bytes = n * batch * sizeof(uint64_t);
    void* d;
    FHE_CHECK(ensure_scratch(ctx, bytes, &d));
    // Pipelined over chunks of polynomials (two copy streams + events): the H2D copy of chunk c+1 and the D2H copy of chunk c-1
    // overlap the transform of chunk c, so a large batch moves at the full-duplex PCIe rate instead of one direction at a time.
    size_t nchunk = (bytes >= ((size_t)32 << 20) && batch >= 8) ? 8 : 1;
    if (const char* e = getenv("FHE_B200_HOST_CHUNKS")) nchunk = std::max<size_t>(1, std::min<size_t>((size_t)atoi(e), batch));  // tuning knob
    if (nchunk == 1) {
        FHE_CUDA(ctx, cudaMemcpyAsync(d, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
        FHE_CHECK(launch_ntt_u64(ctx, q, lg, batch, (uint64_t*)d, fwd));
        FHE_CUDA(ctx, cudaMemcpyAsync(a, d, bytes, cudaMemcpyDeviceToHost, ctx->stream));
        FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return FHE_OK;
    }
    const size_t cs = (batch + nchunk - 1) / nchunk;
    if (!ctx->copy_in) FHE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    if (!ctx->copy_out) FHE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> ev(2 * nchunk + 1);
    const size_t fence = 2 * nchunk;
    for (auto& e : ev) FHE_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    fhe_status st = FHE_OK;
    auto cu = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && st == FHE_OK) st = fail(ctx, FHE_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    };
    cu(cudaEventRecord(ev[fence], ctx->stream), "event");  // the scratch buffer is free once earlier work on the stream is done
    cu(cudaStreamWaitEvent(ctx->copy_in, ev[fence], 0), "wait");
    for (size_t c = 0; c < nchunk && st == FHE_OK; ++c) {
        const size_t off = c * cs;
        if (off >= batch) break;
        const size_t cnt = std::min(cs, batch - off);
        uint64_t* dc = (uint64_t*)d + off * n;
        cu(cudaMemcpyAsync(dc, a + off * n, cnt * n * 8, cudaMemcpyHostToDevice, ctx->copy_in), "H2D");
        cu(cudaEventRecord(ev[2 * c], ctx->copy_in), "event");
        cu(cudaStreamWaitEvent(ctx->stream, ev[2 * c], 0), "wait");
        if (st == FHE_OK) st = launch_ntt_u64(ctx, q, lg, cnt, dc, fwd);
        cu(cudaEventRecord(ev[2 * c + 1], ctx->stream), "event");
        cu(cudaStreamWaitEvent(ctx->copy_out, ev[2 * c + 1], 0), "wait");
        cu(cudaMemcpyAsync(a + off * n, dc, cnt * n * 8, cudaMemcpyDeviceToHost, ctx->copy_out), "D2H");
    }
    cu(cudaEventRecord(ev[fence], ctx->copy_out), "event");
    cu(cudaStreamWaitEvent(ctx->stream, ev[fence], 0), "wait");
    if (st == FHE_OK)
        cu(cudaStreamSynchronize(ctx->stream), "sync");
    else
        cudaDeviceSynchronize();
    for (auto& e : ev) cudaEventDestroy(e);
    return st;
}

}  // namespace fhe

using namespace fhe;

extern "C" {

fhe_status fhe_ntt_fwd_u64(fhe_ctx* ctx, uint64_t q, unsigned log_n, size_t batch, uint64_t* d_a) {
    if (!ctx) return FHE_EINVAL;
    return launch_ntt_u64(ctx, q, log_n, batch, d_a, true);
}
fhe_status fhe_ntt_inv_u64(fhe_ctx* ctx, uint64_t q, unsigned log_n, size_t batch, uint64_t* d_a) {
    if (!ctx) return FHE_EINVAL;
    return launch_ntt_u64(ctx, q, log_n, batch, d_a, false);
}
fhe_status fhe_ntt_fwd_u32(fhe_ctx* ctx, uint32_t q, unsigned log_n, size_t batch, uint32_t* d_a) {
    if (!ctx) return FHE_EINVAL;
    return launch_ntt_u32(ctx, q, log_n, batch, d_a, true);
}
fhe_status fhe_ntt_inv_u32(fhe_ctx* ctx, uint32_t q, unsigned log_n, size_t batch, uint32_t* d_a) {
    if (!ctx) return FHE_EINVAL;
    return launch_ntt_u32(ctx, q, log_n, batch, d_a, false);
}
// [batch][limbs][n]: polynomial p = b*limbs + i uses modulus qs[p % limbs]; one launch for the whole batch
fhe_status fhe_ntt_fwd_rns(fhe_ctx* ctx, const uint64_t* qs, size_t limbs, unsigned log_n, size_t batch, uint64_t* d_a) {
    if (!ctx || !qs) return FHE_EINVAL;
    return launch_ntt_rns_u64(ctx, qs, limbs, log_n, batch * limbs, d_a, true);
}
fhe_status fhe_ntt_inv_rns(fhe_ctx* ctx, const uint64_t* qs, size_t limbs, unsigned log_n, size_t batch, uint64_t* d_a) {
    if (!ctx || !qs) return FHE_EINVAL;
    return launch_ntt_rns_u64(ctx, qs, limbs, log_n, batch * limbs, d_a, false);
}
fhe_status fhe_ntt_fwd_host(fhe_ctx* ctx, uint64_t q, uint64_t* a, size_t n, size_t batch) {
    if (!ctx || !a) return FHE_EINVAL;
    return ntt_host(ctx, q, a, n, batch, true);
}
fhe_status fhe_ntt_inv_host(fhe_ctx* ctx, uint64_t q, uint64_t* a, size_t n, size_t batch) {
    if (!ctx || !a) return FHE_EINVAL;
    return ntt_host(ctx, q, a, n, batch, false);
}

}  // extern "C"
