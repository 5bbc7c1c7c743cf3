// FHEW / LMKCDEY gate bootstrapping on sm_100a (K8-K12 of SURVEY.md §2).
//
//   fhew_prologue_kernel      mod_switch -> Lwe::key_switch -> mod_switch_odd   (lwe.rs:90-99,151-160)
//   fhew_blind_rotate_kernel  persistent CTA per ciphertext: LMKCDEY schedule built in shared memory, accumulator
//                             (2N words) and digit polynomials resident in shared memory across all ~208 steps; per
//                             step: decompose -> 2d (or d) forward NTT -> MAC against pre-transformed key rows
//                             streamed from L2 -> 2 inverse NTT -> (+ b(X^t)); epilogue sample_extract (+ Q/8)
//                             (bootstrapping.rs:149-231, rgsw.rs:116-128, rlwe.rs:177-202, fhew.rs:31-39)
//   fhew_step_kernel          one external product / automorphism per accumulator (parity tests, util-level callers)
// All per-thread logic lives in fhew_core.cuh (shared with tests/hostsim).
#include <algorithm>
#include <cstdlib>
#include <vector>

#include <cstring>

#include "ctx.cuh"
#include "fhew_core.cuh"
#include "fhew_fast.cuh"

struct fhe_fhew_key {
    fhe_fhew_param param;
    fhe::FhewDev P;                      // Q < 2^30; its scalar fields (log_n, n_s, w, decomposors, ak_t) are always filled
    fhe::FhewDevT<fhe::Mod64> P64;       // wide: 2^30 <= Q < 2^62 (examples/multi_key_uint8.rs: 55-bit Q, N = 2048, d = 5)
    bool wide = false;
    fhe::LweKsDev K;
    uint32_t kmax = 0;
    void* d_brk = nullptr;
    void* d_ak = nullptr;
    void* d_ksk = nullptr;
    void* d_dlog = nullptr;
    int* d_err = nullptr;
    size_t brk_bytes = 0, ak_bytes = 0, ksk_bytes = 0;
    // fast path (fhew_fast.cuh): N = 512, Q < 2^28, small digits, instantiated (d_g, d_r)
    bool fast = false;
    fhe::FhewFastDev F;
    void* d_brk4 = nullptr;
    void* d_ak4 = nullptr;
};

namespace fhe {

static constexpr int BR_THREADS = 128;
static constexpr int BR_THREADS_WIDE = 512;
static constexpr int PRO_G = 4;       // ciphertext slots per prologue CTA (key rows are loaded once per CTA iteration) ...
static constexpr int PRO_G_BIG = 4;   // ... and when PRO_G slots of N * d digits would exceed the shared memory (N = 2048, d = 5: 4 x 40 KB)
static constexpr int PRO_THREADS = 128;

template <typename W>
__global__ void fhew_pack_rows_kernel(const W* __restrict__ ab /* [rows][2][N] eval form */, KeyPair<W>* __restrict__ out, uint32_t n,
                                      unsigned long long rows) {
    unsigned long long total = rows * n;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        unsigned long long r = i / n;
        uint32_t c = (uint32_t)(i - r * n);
        out[i] = KeyPair<W>{ab[(r * 2) * n + c], ab[(r * 2 + 1) * n + c]};
    }
}

template <typename CT, typename OT, int PRO_G>
__global__ void __launch_bounds__(PRO_THREADS) fhew_prologue_kernel(LweKsDev K, const CT* __restrict__ ct_in, OT* __restrict__ out,
                                                                      unsigned long long count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* digs = reinterpret_cast<uint32_t*>(smem_raw);
    const uint32_t len = K.n * K.ks_dec.d;
    uint32_t* b_in = digs + (size_t)PRO_G * len;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    const unsigned long long groups = (count + PRO_G - 1) / PRO_G;
    for (unsigned long long grp = blockIdx.x; grp < groups; grp += gridDim.x) {
        const unsigned long long base = grp * PRO_G;
        for (int g = 0; g < PRO_G; ++g) {
            unsigned long long c = base + g < count ? base + g : count - 1;  // clamp (duplicates are not stored)
            lwe_phase_digits(K, ct_in + c * (K.n + 1), digs + (size_t)g * len, b_in + g, tid, nthr);
        }
        __syncthreads();
        for (uint32_t j = tid; j <= K.n_s; j += nthr) {
            uint32_t acc[PRO_G];
            lwe_phase_gemv<PRO_G>(K, digs, j, acc);
            for (int g = 0; g < PRO_G; ++g)
                if (base + g < count) out[(base + g) * (K.n_s + 1) + j] = (OT)lwe_phase_out(K, acc[g], j, b_in[g]);
        }
        __syncthreads();
    }
}

// mode 0: out = LWE ciphertext [N+1] (sample_extract + post_add); mode 1: out = accumulator [2][N]
// NT = threads per CTA the kernel is compiled for: 128 for 32-bit moduli (several CTAs per SM), BR_THREADS_WIDE for 64-bit
// moduli at N = 2048, where the working set (196 KB) admits one CTA per SM and the CTA itself has to fill it
template <typename M, typename FT, typename OT, int NT>
__global__ void __launch_bounds__(NT) fhew_blind_rotate_kernel(FhewDevT<M> P, uint32_t kmax, const FT* __restrict__ f,
                                                                          const uint32_t* __restrict__ ct2n, typename M::W post_add,
                                                                          unsigned long long count, OT* __restrict__ out, int mode,
                                                                          int* __restrict__ err) {
    typedef typename M::W W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W* smem = reinterpret_cast<W*>(smem_raw);
    const uint32_t n = 1u << P.log_n;
    uint16_t* steps = reinterpret_cast<uint16_t*>(smem + (size_t)(2 + kmax) * n);
    const uint32_t max_steps = P.n_s + n + 2;
    uint32_t* a2n = reinterpret_cast<uint32_t*>(steps + ((max_steps + 1) & ~1u));
    __shared__ uint32_t ns_sh;
    // schedule scratch aliases the (not yet used) digit region
    uint16_t* cnt = reinterpret_cast<uint16_t*>(fhew_dig(smem, n, 0));
    uint16_t* sorted = cnt + n;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    auto run = [&](auto phase) {
        phase(tid, nthr);
        __syncthreads();
    };
    for (unsigned long long ct = blockIdx.x; ct < count; ct += gridDim.x) {
        const uint32_t* src = ct2n + ct * (P.n_s + 1);
        for (uint32_t j = tid; j <= P.n_s; j += nthr) a2n[j] = src[j];
        __syncthreads();
        if (tid == 0) ns_sh = build_schedule(n, P.n_s, P.w, a2n, P.dlog, cnt, sorted, steps);
        fhew_phase_init(P, smem, f, a2n[P.n_s], tid, nthr);
        __syncthreads();
        uint32_t ns = ns_sh;
        if (ns == 0xFFFFFFFFu) {  // reference: unreachable!() (bootstrapping.rs:221)
            if (tid == 0) atomicExch(err, 1);
            ns = 0;
        }
        for (uint32_t s = 0; s < ns; ++s) fhew_step(P, smem, steps[s], run);
        if (mode == 0) {
            fhew_phase_extract(P, smem, post_add, out + ct * (n + 1), tid, nthr);
        } else {
            OT* o = out + ct * 2ull * n;
            for (uint32_t i = tid; i < n; i += nthr) {
                o[i] = (OT)smem[swz<W>(i)];
                o[n + i] = (OT)smem[n + swz<W>(i)];
            }
        }
        __syncthreads();
    }
}

template <typename M, typename IT, typename OT>
__global__ void __launch_bounds__(BR_THREADS) fhew_step_kernel(FhewDevT<M> P, uint32_t kmax, uint32_t kind_flag, const uint32_t* __restrict__ idx,
                                                                 const IT* __restrict__ acc_in, OT* __restrict__ acc_out, unsigned long long count) {
    typedef typename M::W W;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    W* smem = reinterpret_cast<W*>(smem_raw);
    const uint32_t n = 1u << P.log_n;
    const uint32_t tid = threadIdx.x, nthr = blockDim.x;
    auto run = [&](auto phase) {
        phase(tid, nthr);
        __syncthreads();
    };
    for (unsigned long long c = blockIdx.x; c < count; c += gridDim.x) {
        const IT* in = acc_in + c * 2ull * n;
        for (uint32_t i = tid; i < n; i += nthr) {
            smem[swz<W>(i)] = (W)in[i];
            smem[n + swz<W>(i)] = (W)in[n + i];
        }
        __syncthreads();
        fhew_step(P, smem, kind_flag | idx[c], run);
        OT* o = acc_out + c * 2ull * n;
        for (uint32_t i = tid; i < n; i += nthr) {
            o[i] = (OT)smem[swz<W>(i)];
            o[n + i] = (OT)smem[n + swz<W>(i)];
        }
        __syncthreads();
    }
}

// {a,b}[rows][N] (uint2) -> [rows][128][2] uint4: {a(4t..4t+3)}, {b(4t..4t+3)}
__global__ void fhew_pack4_kernel(const uint2* __restrict__ ab, uint4* __restrict__ out, unsigned long long rows) {
    const unsigned long long total = rows * FF_THREADS;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const uint2* src = ab + i * 4;
        out[i * 2] = make_uint4(src[0].x, src[1].x, src[2].x, src[3].x);
        out[i * 2 + 1] = make_uint4(src[0].y, src[1].y, src[2].y, src[3].y);
    }
}

// Fast path (fhew_fast.cuh).  mode 0: out = LWE ciphertext [N+1] (sample_extract + post_add); mode 1: accumulator [2][N]
#ifndef FF_MINB
#define FF_MINB 8  // register target (64): shared memory still limits residency to 7 CTAs per SM, but this allocation measured 2 % faster than 72 or 80 registers (56: slower)
#endif
template <typename FT, typename OT>
__global__ void __launch_bounds__(FF_THREADS, FF_MINB) fhew_blind_rotate_fast_kernel(FhewFastDev P, const FT* __restrict__ f,
                                                                             const uint32_t* __restrict__ ct2n, uint32_t post_add,
                                                                             unsigned long long count, OT* __restrict__ out, int mode,
                                                                             int* __restrict__ err) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint32_t* words = reinterpret_cast<uint32_t*>(smem_raw);
    FhewFastSmem S;
    S.acc = words;
    S.dig = words + 2 * FF_N;
#if FF_TW_NC
    S.tw = P.tw;
    S.itw = P.itw;
#else
    TwPair<uint32_t>* stw = reinterpret_cast<TwPair<uint32_t>*>(words + 10 * FF_N);
    S.tw = stw;
    S.itw = stw + FF_N;
#endif
    uint16_t* steps = reinterpret_cast<uint16_t*>(words + ff_fixed_words());
    const uint32_t max_steps = P.n_s + FF_N + 2;
    uint32_t* a2n = reinterpret_cast<uint32_t*>(steps + ((max_steps + 1) & ~1u));
    __shared__ uint32_t ns_sh;
    uint16_t* cnt = reinterpret_cast<uint16_t*>(S.dig);  // schedule scratch aliases the digit region
    uint16_t* sorted = cnt + FF_N;
    const uint32_t tid = threadIdx.x;
#if !FF_TW_NC
    for (uint32_t i = tid; i < (uint32_t)FF_N; i += FF_THREADS) {
        stw[i] = P.tw[i];
        stw[FF_N + i] = P.itw[i];
    }
#endif
    // 64-thread barriers (ids 1, 2) where only the threads of one half exchange data (fhew_fast.cuh, ff_step)
    auto run = [&](auto phase, int scope) {
        phase(tid);
        if (scope == FF_SYNC_FULL)
            __syncthreads();
        else
            if (tid < 64u)
                asm volatile("bar.sync 1, 64;" ::: "memory");
            else
                asm volatile("bar.sync 2, 64;" ::: "memory");
    };
    for (unsigned long long ct = blockIdx.x; ct < count; ct += gridDim.x) {
        const uint32_t* src = ct2n + ct * (P.n_s + 1);
        for (uint32_t j = tid; j <= P.n_s; j += FF_THREADS) a2n[j] = src[j];
        __syncthreads();
        if (tid == 0) ns_sh = build_schedule((uint32_t)FF_N, P.n_s, P.w, a2n, P.dlog, cnt, sorted, steps);
        ff_init(P, S.acc, f, a2n[P.n_s], tid);
        __syncthreads();
        uint32_t ns = ns_sh;
        if (ns == 0xFFFFFFFFu) {  // reference: unreachable!() (bootstrapping.rs:221)
            if (tid == 0) atomicExch(err, 1);
            ns = 0;
        }
        ff_run_steps(P, S, steps, ns, run);
        const uint32_t* acc = S.acc;
        if (mode == 0) {
            ff_extract(P, acc, post_add, out + ct * (FF_N + 1), tid);
        } else {
            OT* o = out + ct * 2ull * FF_N;
            for (uint32_t i = tid; i < 2u * FF_N; i += FF_THREADS) o[i] = (OT)acc[i];
        }
        __syncthreads();
    }
}

static size_t br_fast_smem_bytes(const fhe_fhew_key* key) {
    const uint32_t max_steps = key->F.n_s + FF_N + 2;
    return ff_fixed_words() * 4 + (size_t)((max_steps + 1) & ~1u) * 2 + (size_t)(key->F.n_s + 1) * 4 + 16;
}

static size_t br_smem_bytes(const fhe_fhew_key* key) {
    const uint32_t n = 1u << key->P.log_n;
    const uint32_t max_steps = key->P.n_s + n + 2;
    return (size_t)(2 + key->kmax) * n * (key->wide ? 8 : 4) + (size_t)((max_steps + 1) & ~1u) * 2 + (size_t)(key->P.n_s + 1) * 4 + 16;
}

template <typename K>
static fhe_status persistent_grid(fhe_ctx* ctx, K kern, int threads, size_t smem, unsigned long long items, unsigned* grid) {
    FHE_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    int occ = 0;
    FHE_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) return fail(ctx, FHE_EUNSUPPORTED, "kernel does not fit on an SM (smem %zu bytes)", smem);
    *grid = (unsigned)std::min<unsigned long long>(items, (unsigned long long)ctx->sm_count * occ);
    return FHE_OK;
}

template <typename W>
static fhe_status rows_to_eval_t(fhe_ctx* ctx, const fhe_fhew_key* key, W* d_tmp, size_t rows, void** d_out, size_t* bytes);
template <typename W>
static fhe_status upload_rows_eval_t(fhe_ctx* ctx, const fhe_fhew_key* key, const uint64_t* rows_ab, size_t rows, void** d_out, size_t* bytes) {
    const uint32_t n = 1u << key->P.log_n;
    const size_t words = rows * 2 * n;
    std::vector<W> h(words);
    const uint64_t q = key->param.big_q;
    for (size_t i = 0; i < words; ++i) {
        FHE_REQUIRE(ctx, rows_ab[i] < q, "key coefficient out of range");
        h[i] = (W)rows_ab[i];
    }
    W* d_tmp = nullptr;
    FHE_CUDA(ctx, cudaMalloc(&d_tmp, words * sizeof(W)));
    cudaError_t e = cudaMemcpyAsync(d_tmp, h.data(), words * sizeof(W), cudaMemcpyHostToDevice, ctx->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
    fhe_status st = e == cudaSuccess ? FHE_OK : fail(ctx, FHE_ECUDA, "key upload: %s", cudaGetErrorString(e));
    if (st == FHE_OK) st = rows_to_eval_t<W>(ctx, key, d_tmp, rows, d_out, bytes);
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_tmp);
    return st;
}
// device-resident coefficient-form rows [rows][2 (a, b)][N] (overwritten) -> evaluation-form KeyPair rows in a new buffer
template <typename W>
static fhe_status rows_to_eval_t(fhe_ctx* ctx, const fhe_fhew_key* key, W* d_tmp, size_t rows, void** d_out, size_t* bytes) {
    const uint32_t n = 1u << key->P.log_n;
    const uint64_t q = key->param.big_q;
    fhe_status st = FHE_OK;
    {
        if (sizeof(W) == 4)
            st = launch_ntt_u32(ctx, (uint32_t)q, (unsigned)key->P.log_n, rows * 2, (uint32_t*)d_tmp, true);
        else
            st = launch_ntt_u64(ctx, q, (unsigned)key->P.log_n, rows * 2, (uint64_t*)d_tmp, true);
    }
    if (st == FHE_OK) {
        *bytes = rows * n * sizeof(KeyPair<W>);
        if (cudaMalloc(d_out, *bytes) != cudaSuccess) st = fail(ctx, FHE_ENOMEM, "key alloc");
    }
    if (st == FHE_OK) {
        unsigned long long total = (unsigned long long)rows * n;
        unsigned grid = (unsigned)std::min<unsigned long long>((total + 255) / 256, (unsigned long long)ctx->sm_count * 8);
        fhew_pack_rows_kernel<W><<<grid, 256, 0, ctx->stream>>>(d_tmp, (KeyPair<W>*)*d_out, n, rows);
        st = after_launch(ctx, "fhew_pack_rows_kernel");
    }
    return st;
}
static fhe_status upload_rows_eval(fhe_ctx* ctx, const fhe_fhew_key* key, const uint64_t* rows_ab, size_t rows, void** d_out, size_t* bytes) {
    return key->wide ? upload_rows_eval_t<uint64_t>(ctx, key, rows_ab, rows, d_out, bytes)
                     : upload_rows_eval_t<uint32_t>(ctx, key, rows_ab, rows, d_out, bytes);
}

static fhe_status run_prologue(fhe_ctx* ctx, const fhe_fhew_key* key, size_t count, const uint64_t* d_ct_in, bool sw_in, bool sw_out,
                               uint32_t* d_out32, uint64_t* d_out64) {
    LweKsDev K = key->K;
    K.switch_in = sw_in;
    K.switch_out = sw_out;
    const bool big = ((size_t)PRO_G * K.n * K.ks_dec.d + PRO_G) * 4 > 160 * 1024;
    const size_t g = big ? PRO_G_BIG : PRO_G;
    const size_t smem = (g * K.n * K.ks_dec.d + g) * 4;
    unsigned grid;
    unsigned long long groups = (count + g - 1) / g;
#define FHE_PROLOGUE_LAUNCH(OT_, G_, OUT_)                                                   \
    do {                                                                                     \
        auto kern = fhew_prologue_kernel<uint64_t, OT_, G_>;                                 \
        FHE_CHECK(persistent_grid(ctx, kern, PRO_THREADS, smem, groups, &grid));             \
        kern<<<grid, PRO_THREADS, smem, ctx->stream>>>(K, d_ct_in, OUT_, count);             \
    } while (0)
    if (d_out32) {
        if (big)
            FHE_PROLOGUE_LAUNCH(uint32_t, PRO_G_BIG, d_out32);
        else
            FHE_PROLOGUE_LAUNCH(uint32_t, PRO_G, d_out32);
    } else {
        if (big)
            FHE_PROLOGUE_LAUNCH(uint64_t, PRO_G_BIG, d_out64);
        else
            FHE_PROLOGUE_LAUNCH(uint64_t, PRO_G, d_out64);
    }
#undef FHE_PROLOGUE_LAUNCH
    return after_launch(ctx, "fhew_prologue_kernel");
}

static bool fhew_force_generic() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("FHE_B200_FHEW_GENERIC");
        v = (e && e[0] == '1') ? 1 : 0;
    }
    return v == 1;
}
template <typename OT>
static fhe_status run_blind_rotate_fast(fhe_ctx* ctx, const fhe_fhew_key* key, const uint64_t* d_f, const uint32_t* d_ct2n,
                                        uint64_t post_add, size_t count, OT* d_out, int mode, bool reset_err) {
    const size_t smem = br_fast_smem_bytes(key);
    auto kern = fhew_blind_rotate_fast_kernel<uint64_t, OT>;
    unsigned grid;
    FHE_CHECK(persistent_grid(ctx, kern, FF_THREADS, smem, count, &grid));
    if (reset_err) FHE_CUDA(ctx, cudaMemsetAsync(key->d_err, 0, sizeof(int), ctx->stream));
    kern<<<grid, FF_THREADS, smem, ctx->stream>>>(key->F, d_f, d_ct2n, (uint32_t)post_add, count, d_out, mode, key->d_err);
    return after_launch(ctx, "fhew_blind_rotate_kernel");
}
static bool fhew_fast_instantiated(unsigned dg, unsigned dr) { return dg >= 1 && dg <= 4 && dr >= 1 && dr <= 4; }

template <typename OT>
static fhe_status run_blind_rotate(fhe_ctx* ctx, const fhe_fhew_key* key, const uint64_t* d_f, const uint32_t* d_ct2n, uint64_t post_add,
                                   size_t count, OT* d_out, int mode, bool reset_err = true) {
    if (key->fast && !fhew_force_generic()) {
        return run_blind_rotate_fast<OT>(ctx, key, d_f, d_ct2n, post_add, count, d_out, mode, reset_err);
    }
    const size_t smem = br_smem_bytes(key);
    unsigned grid;
    if (key->wide) {
        auto kern = fhew_blind_rotate_kernel<Mod64, uint64_t, OT, BR_THREADS_WIDE>;
        // small rings leave room for several CTAs per SM: keep 128 threads there
        const int nt = smem > 64 * 1024 ? BR_THREADS_WIDE : BR_THREADS;
        FHE_CHECK(persistent_grid(ctx, kern, nt, smem, count, &grid));
        if (reset_err) FHE_CUDA(ctx, cudaMemsetAsync(key->d_err, 0, sizeof(int), ctx->stream));
        kern<<<grid, nt, smem, ctx->stream>>>(key->P64, key->kmax, d_f, d_ct2n, post_add, count, d_out, mode, key->d_err);
        return after_launch(ctx, "fhew_blind_rotate_kernel");
    }
    auto kern = fhew_blind_rotate_kernel<Mod32, uint64_t, OT, BR_THREADS>;
    FHE_CHECK(persistent_grid(ctx, kern, BR_THREADS, smem, count, &grid));
    if (reset_err) FHE_CUDA(ctx, cudaMemsetAsync(key->d_err, 0, sizeof(int), ctx->stream));
    kern<<<grid, BR_THREADS, smem, ctx->stream>>>(key->P, key->kmax, d_f, d_ct2n, (uint32_t)post_add, count, d_out, mode, key->d_err);
    return after_launch(ctx, "fhew_blind_rotate_kernel");
}

// [count][n_s+1] u64 words mod 2N -> u32, checking what Bootstrapping::blind_rotate assumes of its input (bootstrapping.rs:217-222):
// every word < 2N and every mask word odd or zero
__global__ void fhew_narrow_check_kernel(const uint64_t* __restrict__ in, uint32_t* __restrict__ out, size_t words, uint32_t row, uint32_t two_n,
                                         int* __restrict__ err) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < words; i += (size_t)gridDim.x * blockDim.x) {
        const uint64_t v = in[i];
        const bool body = (i % row) == row - 1;
        if (v >= two_n || (!body && v != 0 && (v & 1) == 0)) atomicOr(err, 1);
        out[i] = (uint32_t)v;
    }
}

static fhe_status check_err_flag(fhe_ctx* ctx, const fhe_fhew_key* key) {
    int h = 0;
    FHE_CUDA(ctx, cudaMemcpyAsync(&h, key->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    if (h) return fail(ctx, FHE_EINVAL, "blind rotation met an even non-zero exponent (reference: unreachable!, bootstrapping.rs:221)");
    return FHE_OK;
}

}  // namespace fhe

using namespace fhe;

extern "C" {

}  // extern "C"

#include "fhew_key.cuh"
fhe_status fhew_key_build(fhe_ctx* ctx, const fhe_fhew_param* pp, const int64_t* ak_t, const FhewKeySource& src, fhe_fhew_key** out) {
    const uint64_t *ksk_a = src.host_ksk_a, *ksk_b = src.host_ksk_b, *brk = src.host_brk, *ak = src.host_ak;
    if (!ctx || !pp || !out) return FHE_EINVAL;
    *out = nullptr;
    FHE_REQUIRE(ctx, ak_t && ((ksk_a && ksk_b && brk && ak) || (src.img_brk && src.img_ak && src.img_ksk) ||
                              (src.dev_brk_rows && src.dev_ak_rows && src.dev_ksk)),
                "null key pointer");
    FHE_REQUIRE(ctx, pp->log_n >= 2 && pp->log_n <= 11, "FHEW path supports 4 <= N <= 2048 (got log_n = %u)", pp->log_n);
    FHE_REQUIRE(ctx, pp->big_q < (1ull << 62), "FHEW path needs Q < 2^62");
    const bool wide = pp->big_q >= (1ull << 30);  // 64-bit residues (generic kernels only)
    FHE_REQUIRE(ctx, wide ? (2 * pp->rgsw_d <= 32 && pp->rlwe_d <= 32) : (2 * pp->rgsw_d <= 16 && pp->rlwe_d <= 16),
                "too many digits: 2 * rgsw_d <= 16 and rlwe_d <= 16 for Q < 2^30 (a u64 accumulates that many unreduced q^2-sized "
                "products), <= 32 above");
    FHE_REQUIRE(ctx, pp->q_ks >= 2 && pp->q_ks <= (1ull << 32) && (pp->q_ks & (pp->q_ks - 1)) == 0, "q_ks must be a power of two <= 2^32");
    FHE_REQUIRE(ctx, pp->w >= 1 && pp->w < 40, "window w must be in [1, 39]");
    FHE_REQUIRE(ctx, pp->rgsw_d >= 1 && pp->rlwe_d >= 1 && pp->ks_d >= 1 && pp->rgsw_log_b >= 1 && pp->rlwe_log_b >= 1 && pp->ks_log_b >= 1,
                "decomposor parameters must be positive");
    FHE_REQUIRE(ctx, pp->rgsw_log_b * pp->rgsw_d <= 64 && pp->rlwe_log_b * pp->rlwe_d <= 64 && pp->ks_log_b * pp->ks_d <= 32,
                "decomposor log_b * d too large");
    // (the check against `wide` follows below, once it is known)
    FHE_REQUIRE(ctx, pp->n_s >= 1 && pp->n_s < 32768, "n_s out of range");
    const uint32_t n = 1u << pp->log_n;
    // the LMKCDEY schedule scratch (u16 cnt[N] + sorted[n_s]) aliases the digit region of kmax * N residue words
    FHE_REQUIRE(ctx, 2 * ((size_t)n + pp->n_s) <= (size_t)std::max(2 * pp->rgsw_d, pp->rlwe_d) * n * (wide ? 8 : 4),
                "n_s = %u too large for N = %u with these decomposors (schedule scratch of 2 (N + n_s) bytes must fit the digit region)", pp->n_s, n);
    if (ksk_a) {
        for (size_t i = 0; i < (size_t)n * pp->ks_d * pp->n_s; ++i) FHE_REQUIRE(ctx, ksk_a[i] < pp->q_ks, "ksk_a word out of range (>= q_ks)");
        for (size_t i = 0; i < (size_t)n * pp->ks_d; ++i) FHE_REQUIRE(ctx, ksk_b[i] < pp->q_ks, "ksk_b word out of range (>= q_ks)");
    }
    const NttTable* t;
    FHE_CHECK(get_ntt_table(ctx, pp->big_q, wide ? 64 : 32, n, &t));
    fhe_fhew_key* key = new fhe_fhew_key();
    key->param = *pp;
    key->wide = wide;
    const uint64_t q = pp->big_q;
    FhewDev& P = key->P;
    if (!wide) P.m = make_mod<Mod32>(q);
    P.lazy = 0;
    P.log_n = (int)pp->log_n;
    P.n_s = pp->n_s;
    P.w = pp->w;
    P.g_dec = make_decomp(q, pp->rgsw_log_b, pp->rgsw_d);
    P.r_dec = make_decomp(q, pp->rlwe_log_b, pp->rlwe_d);
    P.small_digits = (pp->rgsw_log_b * pp->rgsw_d <= 32 && pp->rlwe_log_b * pp->rlwe_d <= 32) ? 1 : 0;
    const uint64_t ninv = host_invmod(n % q, q);
    if (!wide) {
        P.tw = (const TwPair<uint32_t>*)t->d_fwd;
        P.itw = (const TwPair<uint32_t>*)t->d_inv;
        P.ninv = make_twpair<uint32_t>(ninv, q);
        P.wninv = make_twpair<uint32_t>(host_mulmod(t->h_inv[1], ninv, q), q);
    }
    for (unsigned v = 0; v <= pp->w; ++v) P.ak_t[v] = (uint32_t)(((ak_t[v] % (int64_t)(2 * n)) + 2 * n) % (2 * n));
    if (wide) {
        FhewDevT<Mod64>& W = key->P64;
        W.m = make_mod<Mod64>(q);
        W.log_n = P.log_n;
        W.n_s = P.n_s;
        W.w = P.w;
        W.g_dec = P.g_dec;
        W.r_dec = P.r_dec;
        W.small_digits = P.small_digits;
        W.lazy = q < (1ull << 56) ? 1u : 0u;
        W.lz = make_lz64(q);
        W.c64 = (uint64_t)((((u128_t)1) << 64) % q);
        W.tw = (const TwPair<uint64_t>*)t->d_fwd;
        W.itw = (const TwPair<uint64_t>*)t->d_inv;
        W.ninv = make_twpair<uint64_t>(ninv, q);
        W.wninv = make_twpair<uint64_t>(host_mulmod(t->h_inv[1], ninv, q), q);
        for (unsigned v = 0; v <= pp->w; ++v) W.ak_t[v] = P.ak_t[v];
    }
    key->kmax = std::max(2 * pp->rgsw_d, pp->rlwe_d);
    const size_t n_brk_rows = (size_t)pp->n_s * 2 * pp->rgsw_d, n_ak_rows = (size_t)(pp->w + 1) * pp->rlwe_d;
    fhe_status st = FHE_OK;
    if (brk) {
        st = upload_rows_eval(ctx, key, brk, n_brk_rows, &key->d_brk, &key->brk_bytes);
        if (st == FHE_OK) st = upload_rows_eval(ctx, key, ak, n_ak_rows, &key->d_ak, &key->ak_bytes);
    } else if (src.img_brk) {
        const size_t pair = wide ? sizeof(KeyPair<uint64_t>) : sizeof(KeyPair<uint32_t>);
        key->brk_bytes = n_brk_rows * n * pair;
        key->ak_bytes = n_ak_rows * n * pair;
        if (cudaMalloc(&key->d_brk, key->brk_bytes) != cudaSuccess || cudaMalloc(&key->d_ak, key->ak_bytes) != cudaSuccess ||
            cudaMemcpy(key->d_brk, src.img_brk, key->brk_bytes, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(key->d_ak, src.img_ak, key->ak_bytes, cudaMemcpyHostToDevice) != cudaSuccess)
            st = fail(ctx, FHE_ECUDA, "key image upload failed");
    } else {
        st = wide ? rows_to_eval_t<uint64_t>(ctx, key, (uint64_t*)src.dev_brk_rows, n_brk_rows, &key->d_brk, &key->brk_bytes)
                  : rows_to_eval_t<uint32_t>(ctx, key, (uint32_t*)src.dev_brk_rows, n_brk_rows, &key->d_brk, &key->brk_bytes);
        if (st == FHE_OK)
            st = wide ? rows_to_eval_t<uint64_t>(ctx, key, (uint64_t*)src.dev_ak_rows, n_ak_rows, &key->d_ak, &key->ak_bytes)
                      : rows_to_eval_t<uint32_t>(ctx, key, (uint32_t*)src.dev_ak_rows, n_ak_rows, &key->d_ak, &key->ak_bytes);
        if (st == FHE_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, FHE_ECUDA, "key transform failed");
    }
    if (st == FHE_OK) {
        std::vector<uint16_t> dlog(2 * n);
        build_dlog_table(n, dlog.data());
        if (cudaMalloc(&key->d_dlog, dlog.size() * 2) != cudaSuccess || cudaMalloc((void**)&key->d_err, sizeof(int)) != cudaSuccess ||
            cudaMemcpy(key->d_dlog, dlog.data(), dlog.size() * 2, cudaMemcpyHostToDevice) != cudaSuccess)
            st = fail(ctx, FHE_ECUDA, "dlog upload failed");
    }
    if (st == FHE_OK) {
        LweKsDev& K = key->K;
        K.big_q = q;
        K.n = n;
        K.n_s = pp->n_s;
        K.q_ks = pp->q_ks;
        K.q_out = 2 * n;
        K.ks_dec = make_decomp(pp->q_ks, pp->ks_log_b, pp->ks_d);
        K.switch_in = K.switch_out = 1;
        const size_t len = (size_t)n * pp->ks_d, ld = pp->n_s + 1;
        key->ksk_bytes = len * ld * 4;
        if (src.dev_ksk) {
            key->d_ksk = src.dev_ksk;  // adopted
        } else {
            std::vector<uint32_t> h;
            if (ksk_a) {
                h.resize(len * ld);
                for (size_t idx = 0; idx < len; ++idx) {
                    for (size_t j = 0; j < pp->n_s; ++j) h[idx * ld + j] = (uint32_t)ksk_a[idx * pp->n_s + j];
                    h[idx * ld + pp->n_s] = (uint32_t)ksk_b[idx];
                }
            }
            if (cudaMalloc(&key->d_ksk, key->ksk_bytes) != cudaSuccess ||
                cudaMemcpy(key->d_ksk, ksk_a ? (const void*)h.data() : src.img_ksk, key->ksk_bytes, cudaMemcpyHostToDevice) != cudaSuccess)
                st = fail(ctx, FHE_ECUDA, "ksk upload failed");
        }
        K.ksk = (const uint32_t*)key->d_ksk;
    }
    if (st != FHE_OK) {
        fhe_fhew_key_free(ctx, key);
        return st;
    }
    P.brk = (const KeyPair<uint32_t>*)key->d_brk;
    P.ak = (const KeyPair<uint32_t>*)key->d_ak;
    P.dlog = (const uint16_t*)key->d_dlog;
    key->P64.brk = (const KeyPair<uint64_t>*)key->d_brk;
    key->P64.ak = (const KeyPair<uint64_t>*)key->d_ak;
    key->P64.dlog = P.dlog;
    // fast path structures
    if (!wide && pp->log_n == FF_LOGN && q < (1ull << 28) && P.small_digits && pp->rgsw_log_b >= 2 && pp->rlwe_log_b >= 2 && pp->rgsw_log_b * pp->rgsw_d <= 30 &&
        pp->rlwe_log_b * pp->rlwe_d <= 30 &&
        fhew_fast_instantiated(pp->rgsw_d, pp->rlwe_d)) {
        FhewFastDev& F = key->F;
        F.m.q = (uint32_t)q;
        F.m.q2 = (uint32_t)(2 * q);
        F.m.q8 = (uint32_t)(8 * q);
        F.m.mu = (uint32_t)((1ull << 32) / q);
        F.mu64 = (uint64_t)((((u128_t)1) << 64) / q);
        F.n_s = pp->n_s;
        F.w = pp->w;
        F.g_dec = P.g_dec;
        F.r_dec = P.r_dec;
        F.dlog = P.dlog;
        F.tw = P.tw;
        F.itw = P.itw;
        F.ninv = P.ninv;
        F.wninv = P.wninv;
        for (unsigned v = 0; v <= pp->w; ++v) {
            F.ak_t[v] = P.ak_t[v];
            uint32_t inv = 1;  // t^-1 mod 2N by Newton iteration (t odd)
            for (int it = 0; it < 5; ++it) inv = inv * (2u - P.ak_t[v] * inv);
            F.ak_tinv[v] = inv & (2 * n - 1);
        }
        const size_t brk_rows = (size_t)pp->n_s * 2 * pp->rgsw_d, ak_rows = (size_t)(pp->w + 1) * pp->rlwe_d;
        if (cudaMalloc(&key->d_brk4, brk_rows * n * 8) != cudaSuccess || cudaMalloc(&key->d_ak4, ak_rows * n * 8) != cudaSuccess) {
            fhe_fhew_key_free(ctx, key);
            return fail(ctx, FHE_ENOMEM, "fast key alloc");
        }
        fhew_pack4_kernel<<<(unsigned)std::min<size_t>((brk_rows * FF_THREADS + 255) / 256, 4096), 256, 0, ctx->stream>>>(
            (const uint2*)key->d_brk, (uint4*)key->d_brk4, brk_rows);
        fhe_status st2 = after_launch(ctx, "fhew_pack4_kernel");
        if (st2 == FHE_OK) {
            fhew_pack4_kernel<<<(unsigned)std::min<size_t>((ak_rows * FF_THREADS + 255) / 256, 4096), 256, 0, ctx->stream>>>(
                (const uint2*)key->d_ak, (uint4*)key->d_ak4, ak_rows);
            st2 = after_launch(ctx, "fhew_pack4_kernel");
        }
        if (st2 == FHE_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st2 = fail(ctx, FHE_ECUDA, "fast key packing failed");
        if (st2 != FHE_OK) {
            fhe_fhew_key_free(ctx, key);
            return st2;
        }
        F.brk4 = (const uint4*)key->d_brk4;
        F.ak4 = (const uint4*)key->d_ak4;
        key->fast = true;
    }
    *out = key;
    return FHE_OK;
}

extern "C" {

fhe_status fhe_fhew_key_upload(fhe_ctx* ctx, const fhe_fhew_param* pp, const uint64_t* ksk_a, const uint64_t* ksk_b, const uint64_t* brk,
                               const uint64_t* ak, const int64_t* ak_t, fhe_fhew_key** out) {
    if (!ctx || !pp || !out) return FHE_EINVAL;
    *out = nullptr;
    FHE_REQUIRE(ctx, ksk_a && ksk_b && brk && ak && ak_t, "null key pointer");
    FhewKeySource src;
    src.host_ksk_a = ksk_a;
    src.host_ksk_b = ksk_b;
    src.host_brk = brk;
    src.host_ak = ak;
    return fhew_key_build(ctx, pp, ak_t, src, out);
}

// ---- serialised key: header | fhe_fhew_param | ak_t[w+1] | brk image | ak image | ksk image -------------------
// The images are the device buffers themselves (evaluation-form rows, packed ksk), so loading a key is three host-to-device
// copies and no transform.  Versioned; a blob written for another parameter struct size or version is rejected.
struct FhewBlobHeader {
    char magic[8];
    uint32_t version, kind;
    uint64_t param_bytes, n_ak_t, brk_bytes, ak_bytes, ksk_bytes;
};
static const char FHEW_BLOB_MAGIC[8] = {'F', 'H', 'E', 'B', '2', '0', '0', 'K'};
size_t fhe_fhew_key_serialized_size(const fhe_fhew_key* key) {
    if (!key) return 0;
    return sizeof(FhewBlobHeader) + sizeof(fhe_fhew_param) + (key->param.w + 1) * sizeof(int64_t) + key->brk_bytes + key->ak_bytes + key->ksk_bytes;
}
fhe_status fhe_fhew_key_serialize(fhe_ctx* ctx, const fhe_fhew_key* key, void* buf, size_t cap) {
    if (!ctx || !key || !buf) return FHE_EINVAL;
    FHE_REQUIRE(ctx, cap >= fhe_fhew_key_serialized_size(key), "buffer too small for the serialised key");
    FhewBlobHeader h;
    memcpy(h.magic, FHEW_BLOB_MAGIC, 8);
    h.version = 1;
    h.kind = 1;
    h.param_bytes = sizeof(fhe_fhew_param);
    h.n_ak_t = key->param.w + 1;
    h.brk_bytes = key->brk_bytes;
    h.ak_bytes = key->ak_bytes;
    h.ksk_bytes = key->ksk_bytes;
    unsigned char* p = (unsigned char*)buf;
    memcpy(p, &h, sizeof h);
    p += sizeof h;
    memcpy(p, &key->param, sizeof(fhe_fhew_param));
    p += sizeof(fhe_fhew_param);
    const uint32_t two_n = 2u << key->P.log_n;
    for (unsigned v = 0; v <= key->param.w; ++v) {  // exponents are stored reduced mod 2N
        const int64_t t = (int64_t)key->P.ak_t[v];
        memcpy(p + v * sizeof(int64_t), &t, sizeof t);
        (void)two_n;
    }
    p += h.n_ak_t * sizeof(int64_t);
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FHE_CUDA(ctx, cudaMemcpy(p, key->d_brk, key->brk_bytes, cudaMemcpyDeviceToHost));
    p += key->brk_bytes;
    FHE_CUDA(ctx, cudaMemcpy(p, key->d_ak, key->ak_bytes, cudaMemcpyDeviceToHost));
    p += key->ak_bytes;
    FHE_CUDA(ctx, cudaMemcpy(p, key->d_ksk, key->ksk_bytes, cudaMemcpyDeviceToHost));
    return FHE_OK;
}
fhe_status fhe_fhew_key_deserialize(fhe_ctx* ctx, const void* buf, size_t len, fhe_fhew_key** out) {
    if (!ctx || !buf || !out) return FHE_EINVAL;
    *out = nullptr;
    FhewBlobHeader h;
    FHE_REQUIRE(ctx, len >= sizeof h, "serialised key truncated");
    memcpy(&h, buf, sizeof h);
    FHE_REQUIRE(ctx, memcmp(h.magic, FHEW_BLOB_MAGIC, 8) == 0 && h.kind == 1, "not a serialised FHEW key");
    FHE_REQUIRE(ctx, h.version == 1 && h.param_bytes == sizeof(fhe_fhew_param), "serialised key of another format version");
    const unsigned char* p = (const unsigned char*)buf + sizeof h;
    fhe_fhew_param pp;
    FHE_REQUIRE(ctx, len >= sizeof h + sizeof pp, "serialised key truncated");
    memcpy(&pp, p, sizeof pp);
    p += sizeof pp;
    FHE_REQUIRE(ctx, h.n_ak_t == (uint64_t)pp.w + 1 && pp.w < 40 && pp.log_n <= 11, "serialised key header inconsistent");
    const size_t n = (size_t)1 << pp.log_n, pair = pp.big_q >= (1ull << 30) ? 16 : 8;
    FHE_REQUIRE(ctx, h.brk_bytes == (size_t)pp.n_s * 2 * pp.rgsw_d * n * pair && h.ak_bytes == (size_t)(pp.w + 1) * pp.rlwe_d * n * pair &&
                         h.ksk_bytes == n * pp.ks_d * ((size_t)pp.n_s + 1) * 4,
                "serialised key sections do not match its parameters");
    FHE_REQUIRE(ctx, len == sizeof h + sizeof pp + h.n_ak_t * 8 + h.brk_bytes + h.ak_bytes + h.ksk_bytes, "serialised key length mismatch");
    std::vector<int64_t> ak_t(h.n_ak_t);
    memcpy(ak_t.data(), p, h.n_ak_t * 8);
    p += h.n_ak_t * 8;
    FhewKeySource src;
    src.img_brk = p;
    src.img_ak = p + h.brk_bytes;
    src.img_ksk = p + h.brk_bytes + h.ak_bytes;
    return fhew_key_build(ctx, &pp, ak_t.data(), src, out);
}

void fhe_fhew_key_free(fhe_ctx* ctx, fhe_fhew_key* key) {
    if (!key) return;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    if (key->d_brk) cudaFree(key->d_brk);
    if (key->d_ak) cudaFree(key->d_ak);
    if (key->d_ksk) cudaFree(key->d_ksk);
    if (key->d_brk4) cudaFree(key->d_brk4);
    if (key->d_ak4) cudaFree(key->d_ak4);
    if (key->d_dlog) cudaFree(key->d_dlog);
    if (key->d_err) cudaFree(key->d_err);
    delete key;
}

fhe_status fhe_fhew_prologue_batch(fhe_ctx* ctx, const fhe_fhew_key* key, size_t count, const uint64_t* d_ct_in, uint64_t* d_out) {
    if (!ctx || !key) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    return run_prologue(ctx, key, count, d_ct_in, true, true, nullptr, d_out);
}
fhe_status fhe_lwe_key_switch_batch(fhe_ctx* ctx, const fhe_fhew_key* key, size_t count, const uint64_t* d_ct_in, uint64_t* d_out) {
    if (!ctx || !key) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    return run_prologue(ctx, key, count, d_ct_in, false, false, nullptr, d_out);
}

static fhe_status fhew_step_api(fhe_ctx* ctx, const fhe_fhew_key* key, uint32_t flag, size_t count, const uint32_t* d_idx,
                                const uint64_t* d_acc_in, uint64_t* d_acc_out) {
    if (!ctx || !key) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_idx && d_acc_in && d_acc_out, "null pointer");
    // brk[idx] needs idx < n_s, ak[idx] needs idx <= w (the reference panics on a slice index out of bounds)
    FHE_CHECK(validate_below(ctx, d_idx, count, flag ? key->P.w + 1 : key->P.n_s, flag ? "automorphism key index" : "bootstrapping key index"));
    const uint32_t n = 1u << key->P.log_n;
    const size_t smem = (size_t)(2 + key->kmax) * n * (key->wide ? 8 : 4);
    unsigned grid;
    if (key->wide) {
        auto kern = fhew_step_kernel<Mod64, uint64_t, uint64_t>;
        FHE_CHECK(persistent_grid(ctx, kern, BR_THREADS, smem, count, &grid));
        kern<<<grid, BR_THREADS, smem, ctx->stream>>>(key->P64, key->kmax, flag, d_idx, d_acc_in, d_acc_out, count);
        return after_launch(ctx, "fhew_step_kernel");
    }
    auto kern = fhew_step_kernel<Mod32, uint64_t, uint64_t>;
    FHE_CHECK(persistent_grid(ctx, kern, BR_THREADS, smem, count, &grid));
    kern<<<grid, BR_THREADS, smem, ctx->stream>>>(key->P, key->kmax, flag, d_idx, d_acc_in, d_acc_out, count);
    return after_launch(ctx, "fhew_step_kernel");
}
fhe_status fhe_fhew_external_product(fhe_ctx* ctx, const fhe_fhew_key* key, size_t count, const uint32_t* d_idx, const uint64_t* d_acc_in,
                                     uint64_t* d_acc_out) {
    return fhew_step_api(ctx, key, 0u, count, d_idx, d_acc_in, d_acc_out);
}
fhe_status fhe_fhew_automorphism(fhe_ctx* ctx, const fhe_fhew_key* key, size_t count, const uint32_t* d_idx, const uint64_t* d_acc_in,
                                 uint64_t* d_acc_out) {
    return fhew_step_api(ctx, key, FHEW_STEP_AUTO, count, d_idx, d_acc_in, d_acc_out);
}

fhe_status fhe_fhew_blind_rotate_batch(fhe_ctx* ctx, const fhe_fhew_key* key, const uint64_t* d_f, size_t count, const uint64_t* d_ct2n,
                                       uint64_t* d_acc_out) {
    if (!ctx || !key) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_f && d_ct2n && d_acc_out, "null pointer");
    // narrow the [count][n_s+1] u64 input to u32 on the device and validate it before any table is indexed with it
    void* scratch;
    const size_t words = count * (key->P.n_s + 1);
    FHE_CHECK(ensure_scratch(ctx, words * 4, &scratch));
    FHE_CUDA(ctx, cudaMemsetAsync(key->d_err, 0, sizeof(int), ctx->stream));
    fhew_narrow_check_kernel<<<(unsigned)std::min<size_t>((words + 255) / 256, 2048), 256, 0, ctx->stream>>>(
        d_ct2n, (uint32_t*)scratch, words, key->P.n_s + 1, 2u << key->P.log_n, key->d_err);
    FHE_CHECK(after_launch(ctx, "fhew_narrow_check_kernel"));
    {
        int h = 0;
        FHE_CUDA(ctx, cudaMemcpyAsync(&h, key->d_err, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
        FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (h) return fail(ctx, FHE_EINVAL, "blind rotation input must hold words < 2N with odd or zero mask words (bootstrapping.rs:217-222)");
    }
    FHE_CHECK(run_blind_rotate<uint64_t>(ctx, key, d_f, (const uint32_t*)scratch, 0, count, d_acc_out, 1));
    return check_err_flag(ctx, key);
}

fhe_status fhe_fhew_bootstrap_batch(fhe_ctx* ctx, const fhe_fhew_key* key, const uint64_t* d_f, uint64_t post_add, size_t count,
                                    const uint64_t* d_ct_in, uint64_t* d_ct_out) {
    if (!ctx || !key) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_f && d_ct_in && d_ct_out, "null pointer");
    FHE_REQUIRE(ctx, post_add < key->param.big_q, "post_add out of range");
    void* scratch;
    FHE_CHECK(ensure_scratch(ctx, count * (key->P.n_s + 1) * 4, &scratch));
    FHE_CHECK(run_prologue(ctx, key, count, d_ct_in, true, true, (uint32_t*)scratch, nullptr));
    return run_blind_rotate<uint64_t>(ctx, key, d_f, (const uint32_t*)scratch, post_add, count, d_ct_out, 0);
}

fhe_status fhe_fhew_bootstrap_batch_host(fhe_ctx* ctx, const fhe_fhew_key* key, const uint64_t* f, uint64_t post_add, size_t count,
                                         const uint64_t* ct_in, uint64_t* ct_out) {
    if (!ctx || !key) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, f && ct_in && ct_out, "null pointer");
    const uint32_t n = 1u << key->P.log_n;
    const size_t ct_bytes = count * (n + 1) * 8, f_bytes = (size_t)n * 8;
    void *d_in, *d_out, *d_f;
    FHE_CHECK(ensure_stage_d(ctx, 0, ct_bytes, &d_in));
    FHE_CHECK(ensure_stage_d(ctx, 1, ct_bytes, &d_out));
    FHE_CHECK(ensure_stage_d(ctx, 2, f_bytes, &d_f));
    FHE_REQUIRE(ctx, post_add < key->param.big_q, "post_add out of range");
    // Pipelined over chunks: H2D of chunk c+1 and D2H of chunk c-1 overlap the kernels of chunk c (two copy streams + events)
    size_t nchunk = count >= 8192 ? 4 : (count >= 2048 ? 2 : 1);  // measured: 8 or 16 chunks lose more to partial waves than they hide
    if (const char* e = getenv("FHE_B200_HOST_CHUNKS")) nchunk = std::max<size_t>(1, std::min<size_t>((size_t)atoi(e), 64));  // tuning knob
    const size_t cs = (count + nchunk - 1) / nchunk, row = (size_t)n + 1;
    if (!ctx->copy_in) FHE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    if (!ctx->copy_out) FHE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    void* scratch;
    FHE_CHECK(ensure_scratch(ctx, count * (key->P.n_s + 1) * 4, &scratch));
    std::vector<cudaEvent_t> ev(2 * nchunk + 1);  // per chunk: input landed, kernels done; last: fence
    const size_t fence = 2 * nchunk;
    for (auto& e : ev) FHE_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    fhe_status st = FHE_OK;
    auto cu = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && st == FHE_OK) st = fail(ctx, FHE_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    };
    cu(cudaMemcpyAsync(d_f, f, f_bytes, cudaMemcpyHostToDevice, ctx->stream), "H2D f");
    cu(cudaMemsetAsync(key->d_err, 0, sizeof(int), ctx->stream), "memset");
    cu(cudaEventRecord(ev[fence], ctx->stream), "event");  // staging buffers are free once earlier work on the stream is done
    cu(cudaStreamWaitEvent(ctx->copy_in, ev[fence], 0), "wait");
    for (size_t c = 0; c < nchunk && st == FHE_OK; ++c) {
        const size_t off = c * cs, cnt = std::min(cs, count - off);
        if (off >= count) break;
        const uint64_t* h_in = ct_in + off * row;
        uint64_t* dd_in = (uint64_t*)d_in + off * row;
        uint64_t* dd_out = (uint64_t*)d_out + off * row;
        uint32_t* sc = (uint32_t*)scratch + off * (key->P.n_s + 1);
        cu(cudaMemcpyAsync(dd_in, h_in, cnt * row * 8, cudaMemcpyHostToDevice, ctx->copy_in), "H2D");
        cu(cudaEventRecord(ev[2 * c], ctx->copy_in), "event");
        cu(cudaStreamWaitEvent(ctx->stream, ev[2 * c], 0), "wait");
        if (st == FHE_OK) st = run_prologue(ctx, key, cnt, dd_in, true, true, sc, nullptr);
        if (st == FHE_OK) st = run_blind_rotate<uint64_t>(ctx, key, (const uint64_t*)d_f, sc, post_add, cnt, dd_out, 0, false);
        cu(cudaEventRecord(ev[2 * c + 1], ctx->stream), "event");
        cu(cudaStreamWaitEvent(ctx->copy_out, ev[2 * c + 1], 0), "wait");
        cu(cudaMemcpyAsync(ct_out + off * row, dd_out, cnt * row * 8, cudaMemcpyDeviceToHost, ctx->copy_out), "D2H");
    }
    cu(cudaEventRecord(ev[fence], ctx->copy_out), "event");
    cu(cudaStreamWaitEvent(ctx->stream, ev[fence], 0), "wait");
    if (st == FHE_OK) st = check_err_flag(ctx, key);  // synchronises the stream (which now waits for the last D2H)
    else cudaDeviceSynchronize();
    for (auto& e : ev) cudaEventDestroy(e);
    return st;
}

// The device-resident entry points (fhe_fhew_bootstrap_batch) are asynchronous and do not read the key's error flag; this
// call synchronises and reports it (FHE_EINVAL if a blind rotation since the last reset met an even non-zero exponent).
fhe_status fhe_fhew_key_check_error(fhe_ctx* ctx, const fhe_fhew_key* key) {
    if (!ctx || !key) return FHE_EINVAL;
    return check_err_flag(ctx, key);
}

// one-time distribution of the (already transformed) key buffers from `root` (no upstream analogue)
fhe_status fhe_fhew_key_broadcast(fhe_ctx* ctx, fhe_fhew_key* key, void* nccl_comm, int root) {
    if (!ctx || !key) return FHE_EINVAL;
    FHE_CHECK(fhe_keys_broadcast(ctx, nccl_comm, root, key->d_brk, key->brk_bytes));
    FHE_CHECK(fhe_keys_broadcast(ctx, nccl_comm, root, key->d_ak, key->ak_bytes));
    FHE_CHECK(fhe_keys_broadcast(ctx, nccl_comm, root, key->d_ksk, key->ksk_bytes));
    if (key->fast) {
        FHE_CHECK(fhe_keys_broadcast(ctx, nccl_comm, root, key->d_brk4, key->brk_bytes));
        FHE_CHECK(fhe_keys_broadcast(ctx, nccl_comm, root, key->d_ak4, key->ak_bytes));
    }
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FHE_OK;
}
size_t fhe_fhew_key_bytes(const fhe_fhew_key* key) { return key ? key->brk_bytes + key->ak_bytes + key->ksk_bytes : 0; }

}  // extern "C"
