// Context implementation: device selection, streams, memory helpers and the twiddle-table cache.
// Table contents follow compute_twiddle (util/src/ring/fft/zq.rs:58-67) with the root choice of
// Zq::generator / two_adic_generator (util/src/zq.rs:99-109): tw[j] = omega^(brev_{s-1}(j)).
#include "ctx.cuh"

#include <dlfcn.h>

#include <algorithm>

#include <cstring>
#include "host_tables.hpp"

namespace fhe {

bool host_is_prime(uint64_t n) { return host_is_prime_u64(n); }

fhe_status get_mod_info(fhe_ctx* ctx, uint64_t q, const ModInfo** out) {
    auto it = ctx->mods.find(q);
    if (it == ctx->mods.end()) {
        FHE_REQUIRE(ctx, q > 2 && host_is_prime(q), "modulus %llu is not an odd prime (the reference panics here too)", (unsigned long long)q);
        ModInfo mi;
        mi.q = q;
        uint64_t order = q - 1;
        mi.two_adicity = (unsigned)__builtin_ctzll(order);
        uint64_t g = 0;
        for (uint64_t c = 1; c < order; ++c)
            if (host_powmod(c, order >> 1, q) == order) {
                g = c;
                break;
            }
        FHE_REQUIRE(ctx, g != 0, "no generator for %llu", (unsigned long long)q);
        mi.root_2s = host_powmod(g, order >> mi.two_adicity, q);
        it = ctx->mods.emplace(q, mi).first;
    }
    *out = &it->second;
    return FHE_OK;
}

fhe_status get_ntt_table(fhe_ctx* ctx, uint64_t q, int bits, size_t len, const NttTable** out) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    return get_ntt_table_locked(ctx, q, bits, len, out);
}

fhe_status get_ntt_table_locked(fhe_ctx* ctx, uint64_t q, int bits, size_t len, const NttTable** out) {
    const ModInfo* mi;
    FHE_CHECK(get_mod_info(ctx, q, &mi));
    if (len < 2) len = 2;
    FHE_REQUIRE(ctx, (len & (len - 1)) == 0, "table length must be a power of two");
    unsigned lg = 0;
    while (((size_t)1 << lg) < len) ++lg;
    FHE_REQUIRE(ctx, lg + 1 <= mi->two_adicity, "q = %llu has 2-adicity %u: negacyclic NTT of degree %zu needs q = 1 mod %zu",
                (unsigned long long)q, mi->two_adicity, len, 2 * len);
    if (bits == 32) FHE_REQUIRE(ctx, q < (1ull << 30), "u32 path needs q < 2^30");
    if (bits == 64) FHE_REQUIRE(ctx, q < (1ull << 62), "u64 path needs q < 2^62");
    NttTable& t = ctx->tables[std::make_pair(q, bits)];
    if (t.len < len) {
        // prefix property: entries [0, len) of the reference's 2^(s-1)-entry table are psi^(brev_lg(j)), psi = omega^(2^(s-1-lg))
        FHE_REQUIRE(ctx, host_build_twiddles(q, len, t.h_fwd, t.h_inv), "cannot build twiddles for q = %llu", (unsigned long long)q);
        // older (shorter) tables stay alive until the context dies: keys and limb descriptors may still point at them
        // (prefix property: they remain valid for the degrees they were built for)
        if (t.d_fwd) ctx->retired.push_back(t.d_fwd);
        if (t.d_inv) ctx->retired.push_back(t.d_inv);
        t.d_fwd = t.d_inv = nullptr;
        size_t bytes = len * (bits == 32 ? sizeof(TwPair<uint32_t>) : sizeof(TwPair<uint64_t>));
        FHE_CUDA(ctx, cudaMalloc(&t.d_fwd, bytes));
        FHE_CUDA(ctx, cudaMalloc(&t.d_inv, bytes));
        if (bits == 32) {
            std::vector<TwPair<uint32_t>> f(len), v(len);
            for (size_t j = 0; j < len; ++j) {
                f[j] = make_twpair<uint32_t>(t.h_fwd[j], q);
                v[j] = make_twpair<uint32_t>(t.h_inv[j], q);
            }
            FHE_CUDA(ctx, cudaMemcpy(t.d_fwd, f.data(), bytes, cudaMemcpyHostToDevice));
            FHE_CUDA(ctx, cudaMemcpy(t.d_inv, v.data(), bytes, cudaMemcpyHostToDevice));
        } else {
            std::vector<TwPair<uint64_t>> f(len), v(len);
            for (size_t j = 0; j < len; ++j) {
                f[j] = make_twpair<uint64_t>(t.h_fwd[j], q);
                v[j] = make_twpair<uint64_t>(t.h_inv[j], q);
            }
            FHE_CUDA(ctx, cudaMemcpy(t.d_fwd, f.data(), bytes, cudaMemcpyHostToDevice));
            FHE_CUDA(ctx, cudaMemcpy(t.d_inv, v.data(), bytes, cudaMemcpyHostToDevice));
        }
        t.len = len;
    }
    *out = &t;
    return FHE_OK;
}

fhe_status ensure_scratch(fhe_ctx* ctx, size_t bytes, void** out) {
    if (ctx->scratch_bytes < bytes) {
        // stream-ordered safety: earlier kernels may still be using the old scratch
        FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->scratch) cudaFree(ctx->scratch);
        ctx->scratch = nullptr;
        ctx->scratch_bytes = 0;
        size_t want = bytes + bytes / 4;
        FHE_CUDA(ctx, cudaMalloc(&ctx->scratch, want));
        ctx->scratch_bytes = want;
    }
    *out = ctx->scratch;
    return FHE_OK;
}

__global__ void validate_below_kernel(const uint32_t* __restrict__ idx, size_t count, uint32_t limit, int* __restrict__ flag) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += (size_t)gridDim.x * blockDim.x)
        if (idx[i] >= limit) atomicOr(flag, 1);
}
fhe_status validate_below(fhe_ctx* ctx, const uint32_t* d_idx, size_t count, uint32_t limit, const char* what) {
    if (!ctx->d_flag) FHE_CUDA(ctx, cudaMalloc((void**)&ctx->d_flag, sizeof(int)));
    FHE_CUDA(ctx, cudaMemsetAsync(ctx->d_flag, 0, sizeof(int), ctx->stream));
    const unsigned grid = (unsigned)std::min<size_t>((count + 255) / 256, 1024);
    validate_below_kernel<<<grid, 256, 0, ctx->stream>>>(d_idx, count, limit, ctx->d_flag);
    FHE_CHECK(after_launch(ctx, "validate_below_kernel"));
    int h = 0;
    FHE_CUDA(ctx, cudaMemcpyAsync(&h, ctx->d_flag, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h) return fail(ctx, FHE_EINVAL, "%s out of range (must be < %u)", what, limit);
    return FHE_OK;
}

fhe_status ensure_stage_d(fhe_ctx* ctx, int slot, size_t bytes, void** out) {
    if (ctx->stage_d_bytes[slot] < bytes) {
        FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ctx->stage_d[slot]) cudaFree(ctx->stage_d[slot]);
        ctx->stage_d[slot] = nullptr;
        ctx->stage_d_bytes[slot] = 0;
        FHE_CUDA(ctx, cudaMalloc(&ctx->stage_d[slot], bytes));
        ctx->stage_d_bytes[slot] = bytes;
    }
    *out = ctx->stage_d[slot];
    return FHE_OK;
}

}  // namespace fhe

using namespace fhe;

extern "C" {

const char* fhe_version(void) { return "fhe_b200 0.1 (sm_100a)"; }

fhe_status fhe_ctx_create(int device, fhe_ctx** out) {
    if (!out) return FHE_EINVAL;
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0 || device < 0 || device >= count) return FHE_ECUDA;  // no CPU fallback by design
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return FHE_ECUDA;
    if (prop.major < 10) return FHE_EUNSUPPORTED;  // kernels are built for sm_100a only
    if (cudaSetDevice(device) != cudaSuccess) return FHE_ECUDA;
    fhe_ctx* c = new fhe_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking) != cudaSuccess) {
        delete c;
        return FHE_ECUDA;
    }
    c->stream = c->own_stream;
    *out = c;
    return FHE_OK;
}

void fhe_ctx_destroy(fhe_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    for (auto& kv : ctx->tables) {
        if (kv.second.d_fwd) cudaFree(kv.second.d_fwd);
        if (kv.second.d_inv) cudaFree(kv.second.d_inv);
    }
    for (auto& kv : ctx->fft_tables)
        if (kv.second.first) cudaFree(kv.second.first);
    for (void* p : ctx->retired) cudaFree(p);
    for (auto& fn : ctx->cleanup) fn();
    for (auto& kv : ctx->fast_limbs) cudaFree(kv.second);
    if (ctx->scratch) cudaFree(ctx->scratch);
    if (ctx->d_flag) cudaFree(ctx->d_flag);
    for (int i = 0; i < 3; ++i)
        if (ctx->stage_d[i]) cudaFree(ctx->stage_d[i]);
    for (auto& pe : ctx->prof_events) cudaEventDestroy(pe.second);
    if (ctx->prof_start) cudaEventDestroy(ctx->prof_start);
    if (ctx->copy_in) cudaStreamDestroy(ctx->copy_in);
    if (ctx->copy_out) cudaStreamDestroy(ctx->copy_out);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

fhe_status fhe_ctx_set_stream(fhe_ctx* ctx, void* cuda_stream) {
    if (!ctx) return FHE_EINVAL;
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return FHE_OK;
}

fhe_status fhe_sync(fhe_ctx* ctx) {
    if (!ctx) return FHE_EINVAL;
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FHE_OK;
}

const char* fhe_last_error(const fhe_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
uint64_t fhe_launch_count(const fhe_ctx* ctx) { return ctx ? ctx->launches : 0; }
int fhe_sm_count(const fhe_ctx* ctx) { return ctx ? ctx->sm_count : 0; }

fhe_status fhe_malloc(fhe_ctx* ctx, size_t bytes, void** d_ptr) {
    if (!ctx || !d_ptr) return FHE_EINVAL;
    cudaError_t e = cudaMalloc(d_ptr, bytes ? bytes : 1);
    if (e != cudaSuccess) return fail(ctx, FHE_ENOMEM, "cudaMalloc(%zu): %s", bytes, cudaGetErrorString(e));
    return FHE_OK;
}
fhe_status fhe_free(fhe_ctx* ctx, void* d_ptr) {
    if (!ctx) return FHE_EINVAL;
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FHE_CUDA(ctx, cudaFree(d_ptr));
    return FHE_OK;
}
fhe_status fhe_memcpy_h2d(fhe_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
    if (!ctx) return FHE_EINVAL;
    FHE_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
    return FHE_OK;
}
fhe_status fhe_memcpy_d2h(fhe_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
    if (!ctx) return FHE_EINVAL;
    FHE_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FHE_OK;
}

// ---- per-launch timing log ----------------------------------------------------------------------------------------
fhe_status fhe_prof_begin(fhe_ctx* ctx) {
    if (!ctx) return FHE_EINVAL;
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    for (auto& pe : ctx->prof_events) cudaEventDestroy(pe.second);
    ctx->prof_events.clear();
    if (!ctx->prof_start) FHE_CUDA(ctx, cudaEventCreate(&ctx->prof_start));
    FHE_CUDA(ctx, cudaEventRecord(ctx->prof_start, ctx->stream));
    ctx->prof_on = true;
    return FHE_OK;
}
fhe_status fhe_prof_end(fhe_ctx* ctx, char* buf, size_t cap) {
    if (!ctx || !buf || cap < 64) return FHE_EINVAL;
    ctx->prof_on = false;
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    std::map<std::string, std::pair<double, unsigned long long>> agg;  // name -> (ms, launches)
    cudaEvent_t prev = ctx->prof_start;
    for (auto& pe : ctx->prof_events) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, prev, pe.second) == cudaSuccess) {
            auto& a = agg[pe.first];
            a.first += ms;
            a.second += 1;
        }
        prev = pe.second;
    }
    for (auto& pe : ctx->prof_events) cudaEventDestroy(pe.second);
    ctx->prof_events.clear();
    std::string js = "{";
    bool first = true;
    for (auto& kv : agg) {
        char tmp[256];
        snprintf(tmp, sizeof tmp, "%s\"%s\": {\"ms\": %.6f, \"launches\": %llu}", first ? "" : ", ", kv.first.c_str(), kv.second.first,
                 kv.second.second);
        js += tmp;
        first = false;
    }
    js += "}";
    if (js.size() + 1 > cap) return fail(ctx, FHE_EINVAL, "fhe_prof_end: buffer too small (%zu needed)", js.size() + 1);
    memcpy(buf, js.c_str(), js.size() + 1);
    return FHE_OK;
}

// Zq two_adic_primes (util/src/zq.rs:325-329): primes q = 1 mod 2^log_n below 2^bits, descending
fhe_status fhe_two_adic_primes(unsigned bits, unsigned log_n, size_t count, uint64_t* out) {
    if (!out || bits < 2 || bits > 62 || log_n >= bits) return FHE_EINVAL;
    const uint64_t step = 1ull << log_n;
    uint64_t c = (1ull << bits) - step + 1;  // largest value = 1 mod 2^log_n below 2^bits
    const uint64_t lo = 1ull << (bits - 1);
    size_t got = 0;
    while (got < count && c > lo) {
        if (host_is_prime(c)) out[got++] = c;
        c -= step;
    }
    return got == count ? FHE_OK : FHE_EINVAL;
}

// one-time key distribution: ncclBroadcast resolved at run time from the NCCL already loaded in the process
// (torch's bundled libnccl when called from Python, the system libnccl.so.2 otherwise) - no link-time dependency.
fhe_status fhe_keys_broadcast(fhe_ctx* ctx, void* nccl_comm, int root, void* d_buf, size_t bytes) {
    if (!ctx) return FHE_EINVAL;
    FHE_REQUIRE(ctx, nccl_comm && d_buf, "null communicator or buffer");
    typedef int (*bcast_fn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
    static bcast_fn fn = nullptr;
    if (!fn) {
        fn = (bcast_fn)dlsym(RTLD_DEFAULT, "ncclBroadcast");
        if (!fn) {
            void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (h) fn = (bcast_fn)dlsym(h, "ncclBroadcast");
        }
    }
    if (!fn) return fail(ctx, FHE_EUNSUPPORTED, "ncclBroadcast not found (no NCCL in the process and libnccl.so.2 not loadable)");
    int rc = fn(d_buf, d_buf, bytes, /*ncclChar*/ 0, root, nccl_comm, ctx->stream);
    if (rc != 0) return fail(ctx, FHE_ECUDA, "ncclBroadcast failed with %d", rc);
    return FHE_OK;
}

fhe_status fhe_twiddles_host(fhe_ctx* ctx, uint64_t q, size_t len, uint64_t* fwd, uint64_t* inv) {
    if (!ctx) return FHE_EINVAL;
    const NttTable* t;
    FHE_CHECK(get_ntt_table(ctx, q, 64, len, &t));
    for (size_t i = 0; i < len; ++i) {
        if (fwd) fwd[i] = t->h_fwd[i];
        if (inv) inv[i] = t->h_inv[i];
    }
    return FHE_OK;
}

}  // extern "C"
