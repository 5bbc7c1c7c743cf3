// Context: device, stream, twiddle-table cache (replaces the reference's global Mutex<HashMap<q, tables>>,
// util/src/ring/fft/zq.rs:38-56), error reporting.  Host-side C++ (setup path only).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/fhe_b200.h"
#include "modarith.cuh"
#include "ntt_core.cuh"

namespace fhe {

struct NttTable {
    void* d_fwd = nullptr;  // TwPair<W>[len], bit-reversed order
    void* d_inv = nullptr;
    size_t len = 0;
    std::vector<uint64_t> h_fwd, h_inv;  // plain values (host copy, for fhe_twiddles_host / key setup)
    uint64_t omega = 0;                   // primitive 2*len-th root used
};

struct ModInfo {
    uint64_t q = 0;
    unsigned two_adicity = 0;
    uint64_t root_2s = 0;  // omega = g0^((q-1) >> s): primitive 2^s-th root, g0 = smallest g with g^((q-1)/2) = q-1
};

}  // namespace fhe

struct fhe_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    // copy streams of the pipelined *_host entry points (created on first use)
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    int sm_count = 0;
    std::string err;
    uint64_t launches = 0;
    std::mutex mu;
    std::map<uint64_t, fhe::ModInfo> mods;
    std::map<std::pair<uint64_t, int>, fhe::NttTable> tables;  // (q, word bits)
    std::map<int, std::pair<void*, size_t>> fft_tables;        // log_len -> (device table, bytes)   (tfhe)
    std::map<std::vector<uint64_t>, void*> fast_limbs;         // (moduli..., log_n<<8 | bits) -> device FastLimb<L>[]
    std::vector<void*> retired;                                // superseded twiddle tables (freed with the context)
    std::map<std::vector<uint64_t>, void*> rns_tabs;           // RNS base-conversion / rescale tables (ckks.cu)
    std::vector<std::function<void()>> cleanup;                // run by fhe_ctx_destroy
    void* scratch = nullptr;
    size_t scratch_bytes = 0;
    // per-launch timing log (fhe_prof_begin / fhe_prof_end): one event after every launch on the context's stream
    bool prof_on = false;
    cudaEvent_t prof_start = nullptr;
    std::vector<std::pair<const char*, cudaEvent_t>> prof_events;
    // pinned + device staging for the *_host entry points (grow-only)
    void* stage_h[2] = {nullptr, nullptr};
    size_t stage_h_bytes[2] = {0, 0};
    void* stage_d[3] = {nullptr, nullptr, nullptr};
    size_t stage_d_bytes[3] = {0, 0, 0};
    int* d_flag = nullptr;  // device word for the argument checks of the util-level entry points (validate_below)
};

namespace fhe {

inline fhe_status fail(fhe_ctx* ctx, fhe_status st, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    if (ctx) ctx->err = buf;
    return st;
}

#define FHE_CUDA(ctx, call)                                                                             \
    do {                                                                                                \
        cudaError_t e__ = (call);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fhe::fail(ctx, FHE_ECUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)
#define FHE_CHECK(st)              \
    do {                           \
        fhe_status s__ = (st);     \
        if (s__ != FHE_OK) return s__; \
    } while (0)
#define FHE_REQUIRE(ctx, cond, ...) \
    do {                            \
        if (!(cond)) return fhe::fail(ctx, FHE_EINVAL, __VA_ARGS__); \
    } while (0)

// launch bookkeeping + error check after a kernel launch
inline fhe_status after_launch(fhe_ctx* ctx, const char* what) {
    ctx->launches++;
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(ctx, FHE_ECUDA, "launch %s: %s", what, cudaGetErrorString(e));
    if (ctx->prof_on) {
        cudaEvent_t ev;
        if (cudaEventCreate(&ev) == cudaSuccess) {
            cudaEventRecord(ev, ctx->stream);
            ctx->prof_events.emplace_back(what, ev);
        }
    }
    return FHE_OK;
}

bool host_is_prime(uint64_t n);
fhe_status get_mod_info(fhe_ctx* ctx, uint64_t q, const ModInfo** out);
// table with at least `len` entries for modulus q in the given word width (32 or 64)
fhe_status get_ntt_table(fhe_ctx* ctx, uint64_t q, int bits, size_t len, const NttTable** out);
fhe_status get_ntt_table_locked(fhe_ctx* ctx, uint64_t q, int bits, size_t len, const NttTable** out);  // ctx->mu held
fhe_status ensure_scratch(fhe_ctx* ctx, size_t bytes, void** out);
// argument check of a device index array: FHE_EINVAL (naming `what`) unless every d_idx[i] < limit.  The reference panics on
// such inputs (slice index out of bounds); here they must not become out-of-bounds device reads.  Synchronises the stream.
fhe_status validate_below(fhe_ctx* ctx, const uint32_t* d_idx, size_t count, uint32_t limit, const char* what);
// grow-only device staging buffer `slot` (0..2) for the *_host entry points
fhe_status ensure_stage_d(fhe_ctx* ctx, int slot, size_t bytes, void** out);

template <typename A>
A make_mod(uint64_t q);
template <>
inline Mod32 make_mod<Mod32>(uint64_t q) {
    Mod32 m;
    m.q = (uint32_t)q;
    m.q2 = (uint32_t)(2 * q);
    m.mu = q > 1 ? (uint64_t)((((u128_t)1) << 64) / q) : 0;
    return m;
}
template <>
inline Mod64 make_mod<Mod64>(uint64_t q) {
    Mod64 m;
    m.q = q;
    m.q2 = 2 * q;
    unsigned s = 0;
    while (s < 64 && (q >> s)) ++s;  // bit length
    if (s < 2) s = 2;
    m.s = s;
    m.mu = (uint64_t)((((u128_t)1) << (2 * s)) / q);
    return m;
}
template <typename W>
inline TwPair<W> make_twpair(uint64_t w, uint64_t q);
template <>
inline TwPair<uint32_t> make_twpair<uint32_t>(uint64_t w, uint64_t q) {
    return TwPair<uint32_t>{(uint32_t)w, host_shoup32((uint32_t)w, (uint32_t)q)};
}
template <>
inline TwPair<uint64_t> make_twpair<uint64_t>(uint64_t w, uint64_t q) {
    return TwPair<uint64_t>{w, host_shoup64(w, q)};
}
inline uint64_t host_invmod(uint64_t a, uint64_t q) { return host_powmod(a, q - 2, q); }  // q prime

// launchers (ntt_launch.cu)
fhe_status launch_ntt_u64(fhe_ctx* ctx, uint64_t q, unsigned log_n, size_t batch, uint64_t* d_a, bool fwd);
fhe_status launch_ntt_u32(fhe_ctx* ctx, uint32_t q, unsigned log_n, size_t batch, uint32_t* d_a, bool fwd);
// multi-modulus batches: polynomial p (of n_polys, contiguous) uses modulus qs[p % nl]
fhe_status launch_ntt_rns_u64(fhe_ctx* ctx, const uint64_t* qs, size_t nl, unsigned log_n, size_t n_polys, uint64_t* d_a, bool fwd);
// out-of-place form: d_src (may be null = in place) -> d_dst
fhe_status launch_ntt_rns_u64_oop(fhe_ctx* ctx, const uint64_t* qs, size_t nl, unsigned log_n, size_t n_polys, const uint64_t* d_src,
                                  uint64_t* d_dst, bool fwd);
// fast path (ntt_fast_launch.cu): FHE_EUNSUPPORTED (ctx->err untouched) when its preconditions do not hold
fhe_status launch_ntt_fast_u64(fhe_ctx* ctx, const uint64_t* qs, size_t nl, unsigned log_n, size_t n_polys, const uint64_t* d_src,
                               uint64_t* d_a, bool fwd);
fhe_status launch_ntt_fast_u32(fhe_ctx* ctx, const uint64_t* qs, size_t nl, unsigned log_n, size_t n_polys, const uint32_t* d_src,
                               uint32_t* d_a, bool fwd);

}  // namespace fhe
