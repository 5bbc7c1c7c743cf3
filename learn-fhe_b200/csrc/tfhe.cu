// TFHE programmable bootstrapping on sm_100a (K7, K11 (T64 form), K13 of SURVEY.md §2).
//
//   tfhe_fft_mul_kernel        `Rt *= &Rt` (util/src/ring.rs:315-320 -> ring/fft/c64.rs:11-56): one CTA per product
//   tfhe_key_fft_kernel        one-time conversion of TGGSW key polynomials to the twisted Fourier domain (c64.rs:20-28 + fft)
//   tfhe_blind_rotate_kernel   persistent CTA per ciphertext: mod_switch, acc = (0, lut).rotate(-b~), n CMUX steps with the
//                              accumulator, digit spectra and products resident in shared memory, bsk rows streamed from
//                              L2; epilogue sample_extract(0)          (tfhe/bootstrapping.rs:84-104, tggsw.rs:100-121,
//                                                                         tglwe.rs:61-66,115-127)
//   tfhe_ext_kernel            one Tggsw::external_product per TGLWE (parity tests / util-level callers)
//   tlwe_digits_kernel +
//   tlwe_key_switch_kernel     Tlwe::key_switch (tlwe.rs:144-153): signed digits (limb-major) x ksk, wrapping u64 GEMM
// All per-thread logic lives in tfhe_core.cuh (shared with tests/hostsim).  The product path is the reference's own
// floating-point algorithm evaluated in the same operation order: results are bit-identical, not merely within the
// error bound of c64.rs:186-208.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "ctx.cuh"
#include "keygen_stream.cuh"
#include "tfhe_core.cuh"
#include "tfhe_fast.cuh"
#include "tfhe_tables.hpp"

struct fhe_tfhe_key {
    fhe_tfhe_param param;
    fhe::TfheDev P;
    fhe::DecompT64 ks_dec;
    void* d_brk = nullptr;  // Cx [n][(k+1)d][(k+1)][N/2]
    void* d_ksk = nullptr;  // u64 [(kN) d_ks][n+1]
    void* d_ksk_colsum = nullptr;  // u64 [n+1]: 2^(ks_log_b - 1) * (sum over rows of ksk[row][j])  (digit-offset correction)
    size_t brk_bytes = 0, ksk_bytes = 0;
    // bounded-error fast path (tfhe_fast.cuh): second image of the bsk in that path's transform and layout, its tables
    fhe::TfheFastDev F;
    void* d_brk_fast = nullptr;
    void* d_fast_tab = nullptr;
    size_t brk_fast_bytes = 0;
    bool fast_ok = false;  // a specialisation exists for (N, d) and k = 1
    int mode = 0;          // 0 bit-identical reference dataflow, 1 Fourier-domain accumulation (generic kernels), 2 fast path
};

namespace fhe {

static constexpr int TFHE_THREADS = 256;
static constexpr int KS_G = 16;     // ciphertexts per key-switch CTA
static constexpr int KS_COLS = 128; // output columns per key-switch CTA
static constexpr int KS_CH = 256;   // digit rows staged in shared memory per iteration

// f64 tables of ring degree n = 2^log_n, computed on the host exactly like compute_twiddle (c64.rs:98-108):
// cis((i as f64 * PI) / n as f64)
fhe_status get_fft_tab(fhe_ctx* ctx, unsigned log_n, FftTab* out) {
    FHE_REQUIRE(ctx, log_n >= 1 && log_n <= 13, "f64 FFT path supports ring degrees 2..8192");
    std::lock_guard<std::mutex> lock(ctx->mu);
    FftTabHost h;
    auto it = ctx->fft_tables.find((int)log_n);
    if (it == ctx->fft_tables.end()) {
        h.build(log_n);
        void* d = nullptr;
        FHE_CUDA(ctx, cudaMalloc(&d, h.data.size() * sizeof(Cx)));
        FHE_CUDA(ctx, cudaMemcpy(d, h.data.data(), h.data.size() * sizeof(Cx), cudaMemcpyHostToDevice));
        it = ctx->fft_tables.emplace((int)log_n, std::make_pair(d, h.data.size() * sizeof(Cx))).first;
    } else {
        h.log_n = log_n;
        h.m = ((size_t)1 << log_n) / 2;
        h.mb = h.m / 2 > 1 ? h.m / 2 : 1;
    }
    *out = h.view((const Cx*)it->second.first);
    return FHE_OK;
}

#define TFHE_RUN                                        \
    auto run = [&](auto phase) {                        \
        phase((uint32_t)threadIdx.x, (uint32_t)blockDim.x); \
        __syncthreads();                                \
    }

// a <- a * b over T64[X]/(X^n + 1)   (c64.rs:43-56)
__global__ void __launch_bounds__(TFHE_THREADS) tfhe_fft_mul_kernel(FftTab T, unsigned long long batch, uint64_t* __restrict__ a,
                                                                     const uint64_t* __restrict__ b) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx* s = reinterpret_cast<Cx*>(smem_raw);
    const uint32_t m = 1u << T.lg;
    TFHE_RUN;
    for (unsigned long long item = blockIdx.x; item < batch; item += gridDim.x) {
        uint64_t* pa = a + item * 2ull * m;
        const uint64_t* pb = b + item * 2ull * m;
        run([&](uint32_t tid, uint32_t nthr) {
            fft_twist_in(s, T, [&](uint32_t c) { return pa[c]; }, tid, nthr);
            fft_twist_in(s + m, T, [&](uint32_t c) { return pb[c]; }, tid, nthr);
        });
        fft_run<true>(s, 2, T, run);
        run([&](uint32_t tid, uint32_t nthr) {
            for (uint32_t p = tid; p < m; p += nthr) s[swz_cx(p)] = cx_mul(s[swz_cx(p)], s[m + swz_cx(p)]);
        });
        fft_run<false>(s, 1, T, run);
        run([&](uint32_t tid, uint32_t nthr) {
            for (uint32_t p = tid; p < m; p += nthr) {
                uint64_t lo, hi;
                fft_untwist_out(s, T, p, lo, hi);
                pa[p] = lo;
                pa[p + m] = hi;
            }
        });
    }
}

// polys [count][n] u64 -> spectra [count][n/2] Cx (logical index order = fft_in_place output order)
__global__ void __launch_bounds__(TFHE_THREADS) tfhe_key_fft_kernel(FftTab T, unsigned long long count, const uint64_t* __restrict__ polys,
                                                                     Cx* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx* s = reinterpret_cast<Cx*>(smem_raw);
    const uint32_t m = 1u << T.lg;
    TFHE_RUN;
    for (unsigned long long item = blockIdx.x; item < count; item += gridDim.x) {
        const uint64_t* p = polys + item * 2ull * m;
        run([&](uint32_t tid, uint32_t nthr) { fft_twist_in(s, T, [&](uint32_t c) { return p[c]; }, tid, nthr); });
        fft_run<true>(s, 1, T, run);
        run([&](uint32_t tid, uint32_t nthr) {
            for (uint32_t i = tid; i < m; i += nthr) out[item * m + i] = s[swz_cx(i)];
        });
    }
}

// blind_rotate + sample_extract(0): ct_in [count][n_lwe+1] -> out [count][kN+1]
__global__ void __launch_bounds__(TFHE_THREADS, 2) tfhe_blind_rotate_kernel(TfheDev P, const uint64_t* __restrict__ lut,
                                                                          const uint64_t* __restrict__ ct_in, unsigned long long count,
                                                                          uint64_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t n = 1u << P.log_n, k = P.k, d = P.bs_dec.d;
    uint64_t* acc = reinterpret_cast<uint64_t*>(smem_raw);
    Cx* F = reinterpret_cast<Cx*>(acc + (size_t)(k + 1) * n);
    Cx* Pb = F + (size_t)(k + 1) * d * (n / 2);
    const uint32_t rb = 64 - (P.log_n + 1);  // tfhe/bootstrapping.rs:99-104
    TFHE_RUN;
    for (unsigned long long ct = blockIdx.x; ct < count; ct += gridDim.x) {
        const uint64_t* src = ct_in + ct * (P.n_lwe + 1);
        const uint32_t bt = (uint32_t)t64_rounding_shr_dev(src[P.n_lwe], rb) & (2 * n - 1);
        const uint32_t e0 = (2 * n - bt) & (2 * n - 1);  // rotate(-b~)
        run([&](uint32_t tid, uint32_t nthr) {
            for (uint32_t c = tid; c < n; c += nthr) {
                for (uint32_t j = 0; j < k; ++j) acc[(size_t)j * n + c] = 0;
                acc[(size_t)k * n + c] = t64_rot_coef(lut, n, e0, c);
            }
        });
        for (uint32_t i = 0; i < P.n_lwe; ++i) {
            const uint32_t e = (uint32_t)t64_rounding_shr_dev(src[i], rb) & (2 * n - 1);
            if (e == 0) continue;  // rotate(0) - acc = 0: the external product of zero is exactly zero
            tfhe_cmux_step(P, acc, F, Pb, i, e, run);
        }
        // Tglwe::sample_extract(ct, 0) (tglwe.rs:115-127)
        uint64_t* o = out + ct * ((unsigned long long)k * n + 1);
        for (uint32_t c = threadIdx.x; c < k * n; c += blockDim.x) {
            const uint32_t j = c >> P.log_n, x = c & (n - 1);
            o[c] = x == 0 ? acc[(size_t)j * n] : (uint64_t)(0 - acc[(size_t)j * n + (n - x)]);
        }
        if (threadIdx.x == 0) o[(size_t)k * n] = acc[(size_t)k * n];
        __syncthreads();
    }
}

// ---- bounded-error fast path (tfhe_fast.cuh) ------------------------------------------------------------------------------------------
static constexpr int TFHE_FAST_THREADS = 128;
// tuning knobs (A/B builds): CTAs per SM the fused kernel is compiled for (3: 168 registers; 2: 255) and the number of spectral
// positions whose limb-0 key rows are requested before P2 (0: none)
#ifndef TFHE_FAST_MINB
#define TFHE_FAST_MINB 3
#endif
#ifndef TFHE_PRE
#define TFHE_PRE 0
#endif
#ifndef TFHE_KEY_SMEM_DEFAULT
#define TFHE_KEY_SMEM_DEFAULT 0
#endif
// key polynomial (step, r, o) [N] torus words -> forward spectrum in the coalesced P3 layout; one CTA per polynomial
template <typename C>
__global__ void __launch_bounds__(256) tfhe_fast_key_kernel(TfheFastDev P, unsigned long long polys, const uint64_t* __restrict__ src,
                                                            Cx* __restrict__ dst) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    Cx* s = reinterpret_cast<Cx*>(smem_raw);
    for (unsigned long long item = blockIdx.x; item < polys; item += gridDim.x) {
        const uint64_t* p = src + item * C::N;
        for (uint32_t i = threadIdx.x; i < C::M; i += blockDim.x) s[i] = Cx{t64_to_f64(p[i]), t64_to_f64(p[i + C::M])};
        __syncthreads();
        for (int l = 0; l < C::LG; ++l) {
            for (uint32_t b = threadIdx.x; b < C::M / 2; b += blockDim.x) tfhe_fast_key_level(s, C::LG, l, P.fft.W, b);
            __syncthreads();
        }
        const uint32_t o = (uint32_t)(item % 2), r = (uint32_t)((item / 2) % C::NL), step = (uint32_t)(item / (2 * C::NL));
        for (uint32_t i = threadIdx.x; i < C::M; i += blockDim.x) dst[tfhe_fast_key_index<C>(step, r, o, i)] = s[i];
        __syncthreads();
    }
}
// blind_rotate + sample_extract(0) for k = 1: ct_in [count][n_lwe+1] -> out [count][N+1]; persistent CTA per ciphertext,
// accumulator (2 N torus words) and the exchange buffer (2 d N/2 complex) resident in shared memory across all n CMUX steps
// A = uint64_t: accumulator in full torus words (mode 2); uint32_t: its top half only (mode 3)
template <typename C, typename A>
__global__ void __launch_bounds__(TFHE_FAST_THREADS, TFHE_FAST_MINB) tfhe_blind_rotate_fast_kernel(TfheFastDev P, const uint64_t* __restrict__ lut,
                                                                                   const uint64_t* __restrict__ ct_in, unsigned long long count,
                                                                                   uint64_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t N = C::N;
    Cx* X = reinterpret_cast<Cx*>(smem_raw);  // 16-byte aligned entries first
    A* acc = reinterpret_cast<A*>(X + (size_t)C::NL * C::M);
    uint16_t* ex = reinterpret_cast<uint16_t*>(acc + 2 * N);
    const uint32_t rb = 64 - (C::LG + 2);  // tfhe/bootstrapping.rs:99-104: switch to Z_{2N}
    auto run = [&](uint32_t units, auto f) {
        for (uint32_t u = threadIdx.x; u < units; u += TFHE_FAST_THREADS) f(u);
        __syncthreads();
    };
    for (unsigned long long ct = blockIdx.x; ct < count; ct += gridDim.x) {
        const uint64_t* src = ct_in + ct * (P.n_lwe + 1);
        for (uint32_t i = threadIdx.x; i < P.n_lwe; i += TFHE_FAST_THREADS) ex[i] = (uint16_t)((uint32_t)t64_rounding_shr_dev(src[i], rb) & (2 * N - 1));
        const uint32_t bt = (uint32_t)t64_rounding_shr_dev(src[P.n_lwe], rb) & (2 * N - 1);
        const uint32_t e0 = (2 * N - bt) & (2 * N - 1);  // rotate(-b~)
        for (uint32_t c = threadIdx.x; c < N; c += TFHE_FAST_THREADS) {
            acc[c] = 0;
            acc[N + c] = t64_to_acc<A>(t64_rot_coef(lut, N, e0, c));
        }
        __syncthreads();
        for (uint32_t i = 0; i < P.n_lwe; ++i) {
            const uint32_t e = ex[i];
            if (e == 0) continue;  // rotate(0) - acc = 0: the external product of zero is exactly zero
            constexpr int NPRE = TFHE_PRE < (1 << C::R3) ? TFHE_PRE : (1 << C::R3);
            if constexpr (NPRE > 0 && C::U1 == TFHE_FAST_THREADS && C::U2 == 2 * TFHE_FAST_THREADS && C::U3 == TFHE_FAST_THREADS &&
                          C::U4 == 2 * TFHE_FAST_THREADS) {
                // the five phases of tfhe_fast_cmux with one P1 / P3 / P5 unit and two P2 / P4 units per thread; limb 0's key rows are
                // requested from L2 before P2
                const uint32_t t = threadIdx.x;
                const Cx* key = P.key + (size_t)i * C::KEY_STRIDE;
                tfhe_fast_p1<C>(P, acc, X, t, e);
                __syncthreads();
                Cx kq[2 * NPRE];
                tfhe_fast_p3_keys<C, 0, NPRE>(key, t, 0, kq);
                tfhe_fast_mid<C, true>(P, X, t);
                tfhe_fast_mid<C, true>(P, X, t + TFHE_FAST_THREADS);
                __syncthreads();
                tfhe_fast_p3<C, NPRE>(P, X, key, t, kq);
                __syncthreads();
                tfhe_fast_mid<C, false>(P, X, t);
                tfhe_fast_mid<C, false>(P, X, t + TFHE_FAST_THREADS);
                __syncthreads();
                tfhe_fast_p5<C>(P, acc, X, t);
                __syncthreads();
            } else {
                tfhe_fast_cmux<C>(P, acc, X, i, e, run);
            }
        }
        uint64_t* o = out + ct * ((unsigned long long)N + 1);
        for (uint32_t x = threadIdx.x; x < N; x += TFHE_FAST_THREADS) o[x] = acc_to_t64(x == 0 ? acc[0] : (A)(0 - acc[N - x]));
        if (threadIdx.x == 0) o[N] = acc_to_t64(acc[N]);
        __syncthreads();
    }
}

// Variant of the fused kernel (32-bit accumulator words) whose key rows travel global -> shared memory by cp.async one whole CMUX step
// ahead: every thread copies exactly the 16-byte words its own P3 unit will read (no cross-thread hand-over, so cp.async.wait_group is
// the only synchronisation), P3 then reads them with plain shared loads and never waits on L2.  Shared memory: X | acc | one step of
// key rows; the mod-switched exponents are recomputed from the ciphertext (one broadcast load per step, fetched a step ahead).
DEV void cp_async16(void* smem_dst, const void* gsrc) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
DEV void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
DEV void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }
template <typename C>
__global__ void __launch_bounds__(TFHE_FAST_THREADS, 2) tfhe_blind_rotate_fast_ks_kernel(TfheFastDev P, const uint64_t* __restrict__ lut,
                                                                                    const uint64_t* __restrict__ ct_in, unsigned long long count,
                                                                                    uint64_t* __restrict__ out) {
    typedef uint32_t A;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr uint32_t N = C::N;
    constexpr uint32_t ROWS = (uint32_t)(C::KEY_STRIDE / C::U3);  // rows of U3 words per step
    Cx* X = reinterpret_cast<Cx*>(smem_raw);
    Cx* kbuf = X + (size_t)C::NL * C::M;
    A* acc = reinterpret_cast<A*>(kbuf + C::KEY_STRIDE);
    const uint32_t rb = 64 - (C::LG + 2);
    const uint32_t t = threadIdx.x;
    auto run = [&](uint32_t units, auto f) {
        for (uint32_t u = t; u < units; u += TFHE_FAST_THREADS) f(u);
        __syncthreads();
    };
    auto stage = [&](uint32_t step) {  // this thread's words of every row of `step` (units t, t + THREADS, ... of P3)
        const Cx* src = P.key + (size_t)step * C::KEY_STRIDE;
        for (uint32_t g = t; g < C::U3; g += TFHE_FAST_THREADS)
#pragma unroll 8
            for (uint32_t row = 0; row < ROWS; ++row) cp_async16(kbuf + (size_t)row * C::U3 + g, src + (size_t)row * C::U3 + g);
        cp_async_commit();
    };
    for (unsigned long long ct = blockIdx.x; ct < count; ct += gridDim.x) {
        const uint64_t* src = ct_in + ct * (P.n_lwe + 1);
        stage(0);
        const uint32_t bt = (uint32_t)t64_rounding_shr_dev(src[P.n_lwe], rb) & (2 * N - 1);
        const uint32_t e0 = (2 * N - bt) & (2 * N - 1);  // rotate(-b~)
        for (uint32_t c = t; c < N; c += TFHE_FAST_THREADS) {
            acc[c] = 0;
            acc[N + c] = t64_to_acc<A>(t64_rot_coef(lut, N, e0, c));
        }
        uint64_t a_next = src[0];
        __syncthreads();
        for (uint32_t i = 0; i < P.n_lwe; ++i) {
            const uint32_t e = (uint32_t)t64_rounding_shr_dev(a_next, rb) & (2 * N - 1);
            if (i + 1 < P.n_lwe) a_next = src[i + 1];
            if (e != 0) {  // rotate(0) - acc = 0: the external product of zero is exactly zero
                run(C::U1, [&](uint32_t u) { tfhe_fast_p1<C>(P, acc, X, u, e); });
                run(C::U2, [&](uint32_t u) { tfhe_fast_mid<C, true>(P, X, u); });
                cp_async_wait_all();  // own words only: no barrier needed
                run(C::U3, [&](uint32_t u) { tfhe_fast_p3<C, 0, true>(P, X, kbuf, u, nullptr); });
                if (i + 1 < P.n_lwe) stage(i + 1);
                run(C::U4, [&](uint32_t u) { tfhe_fast_mid<C, false>(P, X, u); });
                run(C::U1, [&](uint32_t u) { tfhe_fast_p5<C>(P, acc, X, u); });
            } else {
                cp_async_wait_all();
                if (i + 1 < P.n_lwe) stage(i + 1);
            }
        }
        cp_async_wait_all();
        uint64_t* o = out + ct * ((unsigned long long)N + 1);
        for (uint32_t x = t; x < N; x += TFHE_FAST_THREADS) o[x] = acc_to_t64(x == 0 ? acc[0] : (A)(0 - acc[N - x]));
        if (t == 0) o[N] = acc_to_t64(acc[N]);
        __syncthreads();
    }
}
template <typename C>
static size_t tfhe_fast_ks_smem_bytes() {
    return ((size_t)C::NL * C::M + C::KEY_STRIDE) * sizeof(Cx) + (size_t)2 * C::N * sizeof(uint32_t);
}

// Tggsw::external_product(brk[idx[c]], glwe_c) (in1 == nullptr), or Tggsw::cmux(brk[idx[c]], ct0 = in, ct1 = in1) =
// ct0 + external_product(b, ct1 - ct0) (tggsw.rs:100-121): glwe [count][k+1][N]
__global__ void __launch_bounds__(TFHE_THREADS, 2) tfhe_ext_kernel(TfheDev P, const uint32_t* __restrict__ idx, const uint64_t* __restrict__ in,
                                                                 const uint64_t* __restrict__ in1, unsigned long long count,
                                                                 uint64_t* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t n = 1u << P.log_n, k = P.k, d = P.bs_dec.d;
    uint64_t* acc = reinterpret_cast<uint64_t*>(smem_raw);
    Cx* F = reinterpret_cast<Cx*>(acc + (size_t)(k + 1) * n);
    Cx* Pb = F + (size_t)(k + 1) * d * (n / 2);
    TFHE_RUN;
    for (unsigned long long c = blockIdx.x; c < count; c += gridDim.x) {
        const uint64_t* g = in + c * (unsigned long long)(k + 1) * n;
        const uint64_t* g1 = in1 ? in1 + c * (unsigned long long)(k + 1) * n : nullptr;
        uint64_t* o = out + c * (unsigned long long)(k + 1) * n;
        const Cx* key = P.brk + (((size_t)idx[c] * (k + 1) * d * (k + 1)) << P.fft.lg);
        tfhe_external_product_any(
            P, F, Pb, key, [&](uint32_t j, uint32_t x) { return g1 ? g1[(size_t)j * n + x] - g[(size_t)j * n + x] : g[(size_t)j * n + x]; },
            [&](uint32_t oo, uint32_t c0, uint64_t v0, uint32_t c1, uint64_t v1, bool first) {
                const uint64_t b0 = first ? (g1 ? g[(size_t)oo * n + c0] : 0) : o[(size_t)oo * n + c0];
                const uint64_t b1 = first ? (g1 ? g[(size_t)oo * n + c1] : 0) : o[(size_t)oo * n + c1];
                o[(size_t)oo * n + c0] = b0 + v0;
                o[(size_t)oo * n + c1] = b1 + v1;
            },
            run);
    }
}

// signed digits of the mask of `count` ciphertexts [count][len+1], limb-major (index = digit*len + coefficient,
// tlwe.rs:106,149), packed for the GEMM kernel as dig[group][index][KS_G] (missing ciphertexts of the last group = 0)
__global__ void __launch_bounds__(256) tlwe_digits_kernel(DecompT64 dp, uint32_t len, unsigned long long count, const uint64_t* __restrict__ ct,
                                                          int32_t* __restrict__ dig) {
    const unsigned long long groups = (count + KS_G - 1) / KS_G;
    const unsigned long long total = groups * len * KS_G, stride = (unsigned long long)gridDim.x * blockDim.x;
    const uint64_t mask = (1ull << dp.log_b) - 1;
    for (unsigned long long t = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
        const uint32_t g = (uint32_t)(t % KS_G);
        const unsigned long long r = t / KS_G;
        const uint32_t coef = (uint32_t)(r % len);
        const unsigned long long grp = r / len;
        const unsigned long long c = grp * KS_G + g;
        uint64_t v = c < count ? t64_rounding_shr_dev(ct[c * (len + 1) + coef], dp.rounding_bits) : 0;
        for (uint32_t k = 0; k < dp.d; ++k) {
            const uint64_t limb = v & mask;
            v >>= dp.log_b;
            const uint64_t carry = (((limb - 1) | v) & limb) >> (dp.log_b - 1);
            v += carry;
            // stored with the offset B/2 so that the GEMM multiplies by a small UNSIGNED factor (two 32-bit multiplies per term
            // instead of three for a sign-extended one); the epilogue takes (B/2) * column sum of the key off again
            dig[((grp * dp.d + k) * len + coef) * KS_G + g] = (int32_t)(int64_t)(limb - (carry << dp.log_b)) + (int32_t)(1u << (dp.log_b - 1));
        }
    }
}
// out[c][j] = sum_idx ksk[idx][j] * dig[c][idx] (+ b_in for the body column j == n_out), wrapping mod 2^64
__global__ void __launch_bounds__(KS_COLS) tlwe_key_switch_kernel(uint32_t rows /* len * d */, uint32_t n_out, uint32_t len,
                                                                   unsigned long long count, const int32_t* __restrict__ dig,
                                                                   const uint64_t* __restrict__ ksk, const uint64_t* __restrict__ colsum_half,
                                                                   const uint64_t* __restrict__ ct_in, uint64_t* __restrict__ out) {
    __shared__ __align__(16) int32_t sd[KS_CH * KS_G];
    const uint32_t j = blockIdx.x * KS_COLS + threadIdx.x, ld = n_out + 1;
    const unsigned long long grp = blockIdx.y;
    const int32_t* gd = dig + grp * (unsigned long long)rows * KS_G;
    uint64_t acc[KS_G];
#pragma unroll
    for (int g = 0; g < KS_G; ++g) acc[g] = 0;
    for (uint32_t base = 0; base < rows; base += KS_CH) {
        const uint32_t chunk = min((uint32_t)KS_CH, rows - base);
        for (uint32_t t = threadIdx.x; t < chunk * KS_G / 4; t += blockDim.x)
            reinterpret_cast<int4*>(sd)[t] = reinterpret_cast<const int4*>(gd + (size_t)base * KS_G)[t];
        __syncthreads();
        if (j < ld) {
#pragma unroll 4
            for (uint32_t r = 0; r < chunk; ++r) {
                const uint64_t kv = ksk[(size_t)(base + r) * ld + j];
                const int4* row = reinterpret_cast<const int4*>(sd + r * KS_G);
#pragma unroll
                for (int q4 = 0; q4 < KS_G / 4; ++q4) {
                    const int4 dv = row[q4];
                    acc[4 * q4 + 0] += kv * (uint64_t)(uint32_t)dv.x;
                    acc[4 * q4 + 1] += kv * (uint64_t)(uint32_t)dv.y;
                    acc[4 * q4 + 2] += kv * (uint64_t)(uint32_t)dv.z;
                    acc[4 * q4 + 3] += kv * (uint64_t)(uint32_t)dv.w;
                }
            }
        }
        __syncthreads();
    }
    if (j < ld) {
#pragma unroll
        for (int g = 0; g < KS_G; ++g) {
            const unsigned long long c = grp * KS_G + g;
            if (c < count) {
                uint64_t v = acc[g] - colsum_half[j];  // undo the digit offset: sum (d + B/2) k = sum d k + (B/2) sum k
                if (j == n_out) v += ct_in[c * (len + 1) + len];
                out[c * ld + j] = v;
            }
        }
    }
}

template <typename K>
static fhe_status tfhe_grid(fhe_ctx* ctx, K kern, size_t smem, unsigned long long items, unsigned* grid) {
    FHE_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)std::max<size_t>(smem, 1024)));
    int occ = 0;
    FHE_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TFHE_THREADS, smem));
    if (occ < 1) return fail(ctx, FHE_EUNSUPPORTED, "TFHE kernel does not fit on an SM (%zu bytes of shared memory)", smem);
    *grid = (unsigned)std::min<unsigned long long>(items, (unsigned long long)ctx->sm_count * occ);
    return FHE_OK;
}

static fhe_status run_key_switch(fhe_ctx* ctx, const fhe_tfhe_key* key, size_t count, const uint64_t* d_in, uint64_t* d_out) {
    const fhe_tfhe_param& pp = key->param;
    const uint32_t len = pp.k << pp.log_big_n, rows = len * pp.ks_d;
    const size_t groups = (count + KS_G - 1) / KS_G;
    void* scratch;
    FHE_CHECK(ensure_scratch(ctx, groups * rows * KS_G * sizeof(int32_t), &scratch));
    const unsigned long long total = (unsigned long long)groups * len * KS_G;
    const unsigned dgrid = (unsigned)std::min<unsigned long long>((total + 255) / 256, (unsigned long long)ctx->sm_count * 16);
    tlwe_digits_kernel<<<dgrid, 256, 0, ctx->stream>>>(key->ks_dec, len, count, d_in, (int32_t*)scratch);
    FHE_CHECK(after_launch(ctx, "tlwe_digits_kernel"));
    FHE_REQUIRE(ctx, groups <= 65535, "key-switch batch too large for one launch (max %d ciphertexts)", 65535 * KS_G);
    dim3 grid((pp.n + 1 + KS_COLS - 1) / KS_COLS, (unsigned)groups);
    tlwe_key_switch_kernel<<<grid, KS_COLS, 0, ctx->stream>>>(rows, pp.n, len, count, (const int32_t*)scratch, (const uint64_t*)key->d_ksk,
                                                              (const uint64_t*)key->d_ksk_colsum, d_in, d_out);
    return after_launch(ctx, "tlwe_key_switch_kernel");
}

static fhe_status run_blind_rotate(fhe_ctx* ctx, const fhe_tfhe_key* key, const uint64_t* d_lut, size_t count, const uint64_t* d_in,
                                   uint64_t* d_out) {
    if (key->mode >= 2) {
        fhe_status st = FHE_OK;
        tfhe_fast_dispatch(key->P.log_n - 1, key->P.bs_dec.d, [&](auto cfg) {
            typedef decltype(cfg) C;
            auto launch = [&](auto kern, size_t smem) {
                int occ = 0;
                if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess ||
                    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, TFHE_FAST_THREADS, smem) != cudaSuccess || occ < 1) {
                    st = fail(ctx, FHE_ECUDA, "tfhe_blind_rotate_fast_kernel does not fit (%zu bytes of shared memory)", smem);
                    return;
                }
                const unsigned grid = (unsigned)std::min<unsigned long long>(count, (unsigned long long)ctx->sm_count * occ);
                kern<<<grid, TFHE_FAST_THREADS, smem, ctx->stream>>>(key->F, d_lut, d_in, count, d_out);
                st = after_launch(ctx, "tfhe_blind_rotate_fast_kernel");
            };
            // mode 3 with the key rows staged through shared memory when two such CTAs fit on an SM (FHE_B200_TFHE_KEY_SMEM=0 disables)
            static const bool ks_on = [] {
                const char* e = getenv("FHE_B200_TFHE_KEY_SMEM");
                return e ? atoi(e) != 0 : TFHE_KEY_SMEM_DEFAULT != 0;
            }();
            if (key->mode == 3 && ks_on && tfhe_fast_ks_smem_bytes<C>() <= (size_t)112 * 1024)
                launch(tfhe_blind_rotate_fast_ks_kernel<C>, tfhe_fast_ks_smem_bytes<C>());
            else if (key->mode == 3)
                launch(tfhe_blind_rotate_fast_kernel<C, uint32_t>, tfhe_fast_smem_bytes<C, uint32_t>(key->F.n_lwe));
            else
                launch(tfhe_blind_rotate_fast_kernel<C, uint64_t>, tfhe_fast_smem_bytes<C, uint64_t>(key->F.n_lwe));
        });
        return st;
    }
    const size_t smem = tfhe_smem_bytes(key->P.k, key->P.bs_dec.d, key->P.log_n);
    unsigned grid;
    FHE_CHECK(tfhe_grid(ctx, tfhe_blind_rotate_kernel, smem, count, &grid));
    tfhe_blind_rotate_kernel<<<grid, TFHE_THREADS, smem, ctx->stream>>>(key->P, d_lut, d_in, count, d_out);
    return after_launch(ctx, "tfhe_blind_rotate_kernel");
}

// second image of the bsk for the fast path, from the raw torus polynomials still on the device in d_raw ([polys][N])
// (img != nullptr: the image itself, a host copy of d_brk_fast written by fhe_tfhe_key_serialize, is uploaded instead)
static fhe_status build_fast_key(fhe_ctx* ctx, fhe_tfhe_key* key, const uint64_t* d_raw, size_t polys, const void* img = nullptr, size_t img_bytes = 0) {
    const fhe_tfhe_param& pp = key->param;
    key->fast_ok = false;
    if (pp.k != 1 || pp.bs_log_b * pp.bs_d > 31 || pp.n > 65535) return FHE_OK;
    FastFftTabHost h;
    fhe_status st = FHE_OK;
    const bool have = tfhe_fast_dispatch((int)pp.log_big_n - 1, pp.bs_d, [&](auto cfg) {
        typedef decltype(cfg) C;
        h.build(pp.log_big_n, C::R1, C::R2, C::R3);
        const size_t tab_bytes = h.data.size() * sizeof(Cx);
        key->brk_fast_bytes = (size_t)pp.n * C::KEY_STRIDE * sizeof(Cx);
        if (cudaMalloc(&key->d_fast_tab, tab_bytes) != cudaSuccess || cudaMalloc(&key->d_brk_fast, key->brk_fast_bytes) != cudaSuccess ||
            cudaMemcpyAsync(key->d_fast_tab, h.data.data(), tab_bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess) {
            st = fail(ctx, FHE_ENOMEM, "fast-path bsk image alloc / upload failed");
            return;
        }
        key->F.log_n = (int)pp.log_big_n;
        key->F.n_lwe = pp.n;
        key->F.dig = make_fast_digits(key->P.bs_dec);
        key->F.fft = h.view((const Cx*)key->d_fast_tab);
        key->F.key = (const Cx*)key->d_brk_fast;
        if (img) {
            if (img_bytes != key->brk_fast_bytes || cudaMemcpyAsync(key->d_brk_fast, img, img_bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
                cudaStreamSynchronize(ctx->stream) != cudaSuccess)
                st = fail(ctx, FHE_EINVAL, "fast-path bsk image does not match the parameters");
            return;
        }
        const size_t smem = C::M * sizeof(Cx);
        auto kern = tfhe_fast_key_kernel<C>;
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
            st = fail(ctx, FHE_ECUDA, "tfhe_fast_key_kernel attribute");
            return;
        }
        const unsigned grid = (unsigned)std::min<size_t>(polys, (size_t)ctx->sm_count * 8);
        kern<<<grid, 256, smem, ctx->stream>>>(key->F, polys, d_raw, (Cx*)key->d_brk_fast);
        st = after_launch(ctx, "tfhe_fast_key_kernel");
        if (st == FHE_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, FHE_ECUDA, "fast-path bsk transform failed");
    });
    if (st == FHE_OK) key->fast_ok = have;
    return st;
}

}  // namespace fhe

using namespace fhe;

extern "C" {

fhe_status fhe_fft64_negacyclic_mul(fhe_ctx* ctx, unsigned log_n, size_t batch, uint64_t* d_a, const uint64_t* d_b) {
    if (!ctx) return FHE_EINVAL;
    if (batch == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_a && d_b, "null pointer");
    if (log_n == 0) {  // c64.rs:12-15: a[0] *= b[0]
        std::vector<uint64_t> ha(batch), hb(batch);
        FHE_CUDA(ctx, cudaMemcpyAsync(ha.data(), d_a, batch * 8, cudaMemcpyDeviceToHost, ctx->stream));
        FHE_CUDA(ctx, cudaMemcpyAsync(hb.data(), d_b, batch * 8, cudaMemcpyDeviceToHost, ctx->stream));
        FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        for (size_t i = 0; i < batch; ++i) ha[i] *= hb[i];
        FHE_CUDA(ctx, cudaMemcpyAsync(d_a, ha.data(), batch * 8, cudaMemcpyHostToDevice, ctx->stream));
        FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        return FHE_OK;
    }
    FftTab T;
    FHE_CHECK(get_fft_tab(ctx, log_n, &T));
    const size_t smem = ((size_t)2 << T.lg) * sizeof(Cx);
    unsigned grid;
    FHE_CHECK(tfhe_grid(ctx, tfhe_fft_mul_kernel, smem, batch, &grid));
    tfhe_fft_mul_kernel<<<grid, TFHE_THREADS, smem, ctx->stream>>>(T, batch, d_a, d_b);
    return after_launch(ctx, "tfhe_fft_mul_kernel");
}

fhe_status fhe_fft64_negacyclic_mul_host(fhe_ctx* ctx, uint64_t* a, const uint64_t* b, size_t n, size_t batch) {
    if (!ctx || !a || !b) return FHE_EINVAL;
    FHE_REQUIRE(ctx, n >= 1 && (n & (n - 1)) == 0, "polynomial length %zu is not a power of two", n);
    if (batch == 0) return FHE_OK;
    unsigned lg = 0;
    while (((size_t)1 << lg) < n) ++lg;
    const size_t bytes = n * batch * 8;
    void *da, *db;
    FHE_CHECK(ensure_stage_d(ctx, 0, bytes, &da));
    FHE_CHECK(ensure_stage_d(ctx, 1, bytes, &db));
    FHE_CUDA(ctx, cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, ctx->stream));
    FHE_CUDA(ctx, cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, ctx->stream));
    FHE_CHECK(fhe_fft64_negacyclic_mul(ctx, lg, batch, (uint64_t*)da, (const uint64_t*)db));
    FHE_CUDA(ctx, cudaMemcpyAsync(a, da, bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FHE_OK;
}

}  // extern "C"

// column sums of the merged ksk [rows][ld], scaled by B/2 (the digit-offset correction of tlwe_key_switch_kernel)
__global__ void tfhe_ksk_colsum_kernel(const uint64_t* __restrict__ ksk, size_t rows, uint32_t ld, uint32_t shift, uint64_t* __restrict__ out) {
    const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= ld) return;
    uint64_t s = 0;
    for (size_t r = 0; r < rows; ++r) s += ksk[r * ld + j];
    out[j] = s << shift;
}
// Key object from either the reference layout on the HOST (brk, ksk_a, ksk_b: fhe_tfhe_key_upload) or buffers already on the DEVICE
// (d_brk_raw [polys][N] torus words in the same order, d_ksk_merged [(kN) d_ks][n+1]: fhe_tfhe_keygen; d_ksk_merged is adopted).
// host copies of a key's device buffers (fhe_tfhe_key_serialize); fast == nullptr when the key had no fused-path image
struct TfheKeyImages {
    const void *brk = nullptr, *ksk = nullptr, *colsum = nullptr, *fast = nullptr;
    size_t brk_bytes = 0, ksk_bytes = 0, colsum_bytes = 0, fast_bytes = 0;
};
static fhe_status tfhe_key_build(fhe_ctx* ctx, const fhe_tfhe_param* pp, const uint64_t* brk, const uint64_t* ksk_a, const uint64_t* ksk_b,
                                 uint64_t* d_brk_raw, uint64_t* d_ksk_merged, fhe_tfhe_key** out, const TfheKeyImages* img = nullptr) {
    if (!ctx || !pp || !out) return FHE_EINVAL;
    *out = nullptr;
    FHE_REQUIRE(ctx, (brk && ksk_a && ksk_b) || (d_brk_raw && d_ksk_merged) || img, "null key pointer");
    FHE_REQUIRE(ctx, pp->log_big_n >= 1 && pp->log_big_n <= 12, "TFHE ring degree 2^%u out of range (2..4096)", pp->log_big_n);
    FHE_REQUIRE(ctx, pp->k >= 1 && pp->k <= 4, "GLWE dimension k must be in [1, 4]");
    FHE_REQUIRE(ctx, pp->n >= 1 && pp->n <= 65535, "TLWE dimension out of range");
    FHE_REQUIRE(ctx, pp->bs_log_b >= 1 && pp->bs_d >= 1 && pp->bs_log_b * pp->bs_d <= 64 && pp->bs_log_b <= 52,
                "TGGSW decomposor out of range (log_b * d <= 64)");
    FHE_REQUIRE(ctx, pp->ks_log_b >= 1 && pp->ks_d >= 1 && pp->ks_log_b * pp->ks_d <= 64 && pp->ks_log_b <= 31,
                "TLWE key-switch decomposor out of range (log_b <= 31, log_b * d <= 64)");
    FHE_REQUIRE(ctx, pp->log_p + pp->padding < 64, "log_p + padding must be < 64");
    const size_t n = (size_t)1 << pp->log_big_n, m = n / 2;
    const size_t smem = tfhe_smem_bytes(pp->k, pp->bs_d, (int)pp->log_big_n);
    FHE_REQUIRE(ctx, smem <= 220 * 1024, "TFHE parameters need %zu bytes of shared memory per ciphertext (max 220 KiB)", smem);
    fhe_tfhe_key* key = new fhe_tfhe_key();
    key->param = *pp;
    key->ks_dec = make_decomp_t64(pp->ks_log_b, pp->ks_d);
    TfheDev& P = key->P;
    P.log_n = (int)pp->log_big_n;
    P.k = pp->k;
    P.n_lwe = pp->n;
    P.bs_dec = make_decomp_t64(pp->bs_log_b, pp->bs_d);
    P.fourier_acc = 0;
    fhe_status st = get_fft_tab(ctx, pp->log_big_n, &P.fft);
    // brk: [n][(k+1)d][(k+1)][N] torus words -> Fourier domain
    const size_t polys = (size_t)pp->n * (pp->k + 1) * pp->bs_d * (pp->k + 1);
    uint64_t* d_tmp = d_brk_raw;
    key->brk_bytes = polys * m * sizeof(Cx);
    if (st == FHE_OK && img) {  // serialised key: the device images as they were, no transform
        const size_t rows = (size_t)pp->k * n * pp->ks_d, ld = pp->n + 1;
        key->ksk_bytes = rows * ld * 8;
        if (img->brk_bytes != key->brk_bytes || img->ksk_bytes != key->ksk_bytes || img->colsum_bytes != ld * 8)
            st = fail(ctx, FHE_EINVAL, "serialised key sections do not match its parameters");
        if (st == FHE_OK && (cudaMalloc(&key->d_brk, key->brk_bytes) != cudaSuccess || cudaMalloc(&key->d_ksk, key->ksk_bytes) != cudaSuccess ||
                             cudaMalloc(&key->d_ksk_colsum, ld * 8) != cudaSuccess))
            st = fail(ctx, FHE_ENOMEM, "key alloc");
        if (st == FHE_OK && (cudaMemcpy(key->d_brk, img->brk, key->brk_bytes, cudaMemcpyHostToDevice) != cudaSuccess ||
                             cudaMemcpy(key->d_ksk, img->ksk, key->ksk_bytes, cudaMemcpyHostToDevice) != cudaSuccess ||
                             cudaMemcpy(key->d_ksk_colsum, img->colsum, ld * 8, cudaMemcpyHostToDevice) != cudaSuccess))
            st = fail(ctx, FHE_ECUDA, "key image upload failed");
        if (st == FHE_OK && img->fast) st = build_fast_key(ctx, key, nullptr, polys, img->fast, img->fast_bytes);
        if (st == FHE_OK && img->fast && !key->fast_ok) st = fail(ctx, FHE_EINVAL, "serialised key carries a fused-path image its parameters do not support");
        if (st != FHE_OK) {
            fhe_tfhe_key_free(ctx, key);
            return st;
        }
        P.brk = (const Cx*)key->d_brk;
        *out = key;
        return FHE_OK;
    }
    if (st == FHE_OK && !d_tmp && cudaMalloc((void**)&d_tmp, polys * n * 8) != cudaSuccess) st = fail(ctx, FHE_ENOMEM, "bsk staging alloc");
    if (st == FHE_OK && cudaMalloc(&key->d_brk, key->brk_bytes) != cudaSuccess) st = fail(ctx, FHE_ENOMEM, "bsk alloc");
    if (st == FHE_OK && brk && cudaMemcpyAsync(d_tmp, brk, polys * n * 8, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess)
        st = fail(ctx, FHE_ECUDA, "bsk upload");
    if (st == FHE_OK) {
        unsigned grid;
        st = tfhe_grid(ctx, tfhe_key_fft_kernel, m * sizeof(Cx), polys, &grid);
        if (st == FHE_OK) {
            tfhe_key_fft_kernel<<<grid, TFHE_THREADS, m * sizeof(Cx), ctx->stream>>>(P.fft, polys, d_tmp, (Cx*)key->d_brk);
            st = after_launch(ctx, "tfhe_key_fft_kernel");
        }
    }
    if (st == FHE_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, FHE_ECUDA, "bsk transform failed");
    if (st == FHE_OK) st = build_fast_key(ctx, key, d_tmp, polys);
    if (d_tmp && !d_brk_raw) cudaFree(d_tmp);
    if (st == FHE_OK && d_ksk_merged) {  // already merged on the device: adopt, compute the correction sums there
        const size_t rows = (size_t)pp->k * n * pp->ks_d, ld = pp->n + 1;
        key->ksk_bytes = rows * ld * 8;
        key->d_ksk = d_ksk_merged;
        if (cudaMalloc(&key->d_ksk_colsum, ld * 8) != cudaSuccess) st = fail(ctx, FHE_ENOMEM, "ksk alloc");
        if (st == FHE_OK) {
            tfhe_ksk_colsum_kernel<<<(unsigned)((ld + 127) / 128), 128, 0, ctx->stream>>>((const uint64_t*)key->d_ksk, rows, (uint32_t)ld, pp->ks_log_b - 1,
                                                                                          (uint64_t*)key->d_ksk_colsum);
            st = after_launch(ctx, "tfhe_ksk_colsum_kernel");
            if (st == FHE_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, FHE_ECUDA, "ksk column sums failed");
        }
    }
    // ksk: [(kN) d_ks][n] + [(kN) d_ks] -> [(kN) d_ks][n+1]
    if (st == FHE_OK && !d_ksk_merged) {
        const size_t rows = (size_t)pp->k * n * pp->ks_d, ld = pp->n + 1;
        std::vector<uint64_t> h(rows * ld);
        for (size_t r = 0; r < rows; ++r) {
            for (size_t j = 0; j < pp->n; ++j) h[r * ld + j] = ksk_a[r * pp->n + j];
            h[r * ld + pp->n] = ksk_b[r];
        }
        key->ksk_bytes = h.size() * 8;
        std::vector<uint64_t> cs(ld, 0);
        for (size_t r = 0; r < rows; ++r)
            for (size_t j = 0; j < ld; ++j) cs[j] += h[r * ld + j];
        for (size_t j = 0; j < ld; ++j) cs[j] <<= (pp->ks_log_b - 1);
        if (cudaMalloc(&key->d_ksk, key->ksk_bytes) != cudaSuccess || cudaMalloc(&key->d_ksk_colsum, ld * 8) != cudaSuccess ||
            cudaMemcpy(key->d_ksk, h.data(), key->ksk_bytes, cudaMemcpyHostToDevice) != cudaSuccess ||
            cudaMemcpy(key->d_ksk_colsum, cs.data(), ld * 8, cudaMemcpyHostToDevice) != cudaSuccess)
            st = fail(ctx, FHE_ECUDA, "ksk upload failed");
    }
    if (st != FHE_OK) {
        if (d_ksk_merged) key->d_ksk = nullptr;  // stays with the caller on failure
        fhe_tfhe_key_free(ctx, key);
        return st;
    }
    P.brk = (const Cx*)key->d_brk;
    *out = key;
    return FHE_OK;
}

// ---- key generation on the device (tfhe/bootstrapping.rs:59-76) ------------------------------------------------
// masks of the TGGSW rows: A [rows][k][N] uniform torus words; Srep [rows][k][N] = the TGLWE secret polynomials repeated per row
__global__ void __launch_bounds__(256) tfhe_kg_masks_kernel(uint64_t seed, uint32_t log_n, uint32_t k, unsigned long long rows, const int64_t* __restrict__ s,
                                                            uint64_t* __restrict__ a, uint64_t* __restrict__ srep) {
    const unsigned long long total = (rows * k) << log_n, stride = (unsigned long long)gridDim.x * blockDim.x;
    const uint32_t n = 1u << log_n;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        a[i] = ks_u64(seed, KS_TFHE_BRK_A, i);
        srep[i] = (uint64_t)s[(((i >> log_n) % k) << log_n) + (i & (n - 1))];
    }
}
// out [rows][(k+1)][N] in the upload order (a_0 .. a_{k-1}, b): b = sum_j (a_j s_j) + e (+ the plaintext z_i B^d at coefficient 0 of the
// component the row's index selects, tggsw.rs:84-87)
__global__ void __launch_bounds__(256) tfhe_kg_rows_kernel(uint64_t seed, uint64_t sigma_q, uint32_t log_n, uint32_t k, uint32_t d, uint32_t log_b,
                                                           uint32_t rounding_bits, unsigned long long rows, const int64_t* __restrict__ z,
                                                           const uint64_t* __restrict__ a, const uint64_t* __restrict__ as, uint64_t* __restrict__ out) {
    const unsigned long long total = rows << log_n, stride = (unsigned long long)gridDim.x * blockDim.x;
    const uint32_t n = 1u << log_n, rows_per = (k + 1) * d;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const unsigned long long R = i >> log_n;
        const uint32_t c = (uint32_t)(i & (n - 1)), r = (uint32_t)(R % rows_per), comp = r / d, dig = r % d;
        const uint64_t pt = c == 0 ? (uint64_t)z[R / rows_per] << (rounding_bits + dig * log_b) : 0;
        uint64_t b = ks_tgauss(seed, KS_TFHE_BRK_E, i, sigma_q);
        for (uint32_t j = 0; j < k; ++j) {
            const unsigned long long src = ((R * k + j) << log_n) + c;
            b += as[src];
            out[((R * (k + 1) + j) << log_n) + c] = a[src] + (j == comp ? pt : 0);
        }
        out[((R * (k + 1) + k) << log_n) + c] = b + (comp == k ? pt : 0);
    }
}
// TLWE key-switching key (tlwe.rs:100-111, 122-132) in the merged layout [(kN) d_ks][n + 1]
__global__ void __launch_bounds__(128) tfhe_kg_ksk_kernel(uint64_t seed, uint64_t sigma_q, uint32_t n_lwe, uint32_t kn, uint32_t d_ks, uint32_t log_b,
                                                          uint32_t rounding_bits, const int64_t* __restrict__ z, const int64_t* __restrict__ s,
                                                          uint64_t* __restrict__ ksk) {
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < (unsigned long long)kn * d_ks;
         idx += (unsigned long long)gridDim.x * blockDim.x) {
        const uint32_t dig = (uint32_t)(idx / kn), i = (uint32_t)(idx % kn);
        uint64_t* row = ksk + idx * (n_lwe + 1);
        uint64_t dot = 0;
        for (uint32_t j = 0; j < n_lwe; ++j) {
            const uint64_t a = ks_u64(seed, KS_TFHE_KSK_A, idx * n_lwe + j);
            row[j] = a;
            dot += a * (uint64_t)z[j];
        }
        row[n_lwe] = dot + ks_tgauss(seed, KS_TFHE_KSK_E, idx, sigma_q) + ((uint64_t)(-s[i]) << (rounding_bits + dig * log_b));
    }
}

extern "C" {

fhe_status fhe_tfhe_key_upload(fhe_ctx* ctx, const fhe_tfhe_param* pp, const uint64_t* brk, const uint64_t* ksk_a, const uint64_t* ksk_b,
                               fhe_tfhe_key** out) {
    if (!ctx || !pp || !out) return FHE_EINVAL;
    *out = nullptr;
    FHE_REQUIRE(ctx, brk && ksk_a && ksk_b, "null key pointer");
    return tfhe_key_build(ctx, pp, brk, ksk_a, ksk_b, nullptr, nullptr, out);
}

fhe_status fhe_tfhe_keygen(fhe_ctx* ctx, const fhe_tfhe_param* pp, double tlwe_std, double tglwe_std, uint64_t seed, int64_t* z_out, int64_t* s_out,
                           uint64_t* brk_out, uint64_t* ksk_a_out, uint64_t* ksk_b_out, fhe_tfhe_key** out) {
    if (!ctx || !pp || !out) return FHE_EINVAL;
    *out = nullptr;
    FHE_REQUIRE(ctx, pp->log_big_n >= 1 && pp->log_big_n <= 12 && pp->k >= 1 && pp->k <= 4 && pp->n >= 1 && pp->n <= 65535, "TFHE parameters out of range");
    FHE_REQUIRE(ctx, tlwe_std > 0 && tlwe_std < 0.00390625 && tglwe_std > 0 && tglwe_std < 0.00390625, "noise standard deviations must be in (0, 2^-8)");
    FHE_REQUIRE(ctx, pp->bs_log_b * pp->bs_d <= 64 && pp->ks_log_b * pp->ks_d <= 64, "decomposor log_b * d <= 64");
    const uint32_t log_n = pp->log_big_n, N = 1u << log_n, k = pp->k, kn = k * N;
    const size_t rows = (size_t)pp->n * (k + 1) * pp->bs_d, ksk_rows = (size_t)kn * pp->ks_d, ld = pp->n + 1;
    const uint64_t sq_tlwe = (uint64_t)llround(tlwe_std * 18446744073709551616.0), sq_tglwe = (uint64_t)llround(tglwe_std * 18446744073709551616.0);
    const DecompT64 bsd = make_decomp_t64(pp->bs_log_b, pp->bs_d), ksd = make_decomp_t64(pp->ks_log_b, pp->ks_d);
    std::vector<int64_t> z(pp->n), s(kn);
    for (uint32_t i = 0; i < pp->n; ++i) z[i] = ks_binary(seed, KS_TFHE_Z, i);
    for (uint32_t i = 0; i < kn; ++i) s[i] = ks_binary(seed, KS_TFHE_S, i);
    int64_t *d_z = nullptr, *d_s = nullptr;
    uint64_t *d_a = nullptr, *d_as = nullptr, *d_srep = nullptr, *d_raw = nullptr, *d_ksk = nullptr;
    fhe_status st = FHE_OK;
    auto cu = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && st == FHE_OK) st = fail(ctx, FHE_ECUDA, "tfhe keygen %s: %s", what, cudaGetErrorString(e));
    };
    cu(cudaMalloc(&d_z, pp->n * 8), "alloc");
    cu(cudaMalloc(&d_s, (size_t)kn * 8), "alloc");
    cu(cudaMalloc(&d_a, rows * kn * 8), "alloc");
    cu(cudaMalloc(&d_as, rows * kn * 8), "alloc");
    cu(cudaMalloc(&d_srep, rows * kn * 8), "alloc");
    cu(cudaMalloc(&d_raw, rows * (k + 1) * N * 8), "alloc");
    cu(cudaMalloc(&d_ksk, ksk_rows * ld * 8), "alloc");
    if (st == FHE_OK) {
        cu(cudaMemcpyAsync(d_z, z.data(), pp->n * 8, cudaMemcpyHostToDevice, ctx->stream), "copy");
        cu(cudaMemcpyAsync(d_s, s.data(), (size_t)kn * 8, cudaMemcpyHostToDevice, ctx->stream), "copy");
    }
    auto grid = [&](unsigned long long total) { return (unsigned)std::min<unsigned long long>((total + 255) / 256, (unsigned long long)ctx->sm_count * 16); };
    if (st == FHE_OK) {
        tfhe_kg_masks_kernel<<<grid((unsigned long long)rows * kn), 256, 0, ctx->stream>>>(seed, log_n, k, rows, d_s, d_a, d_srep);
        st = after_launch(ctx, "tfhe_kg_masks_kernel");
    }
    cu(cudaMemcpyAsync(d_as, d_a, rows * kn * 8, cudaMemcpyDeviceToDevice, ctx->stream), "copy");
    // a_j * s_j: the reference's own f64 FFT product (`&a * sk`, tglwe.rs:97-100 -> ring/fft/c64.rs), bit-identical kernel
    if (st == FHE_OK) st = fhe_fft64_negacyclic_mul(ctx, log_n, rows * k, d_as, d_srep);
    if (st == FHE_OK) {
        tfhe_kg_rows_kernel<<<grid((unsigned long long)rows << log_n), 256, 0, ctx->stream>>>(seed, sq_tglwe, log_n, k, pp->bs_d, pp->bs_log_b, bsd.rounding_bits,
                                                                                              rows, d_z, d_a, d_as, d_raw);
        st = after_launch(ctx, "tfhe_kg_rows_kernel");
    }
    if (st == FHE_OK) {
        tfhe_kg_ksk_kernel<<<(unsigned)std::min<size_t>((ksk_rows + 127) / 128, (size_t)ctx->sm_count * 16), 128, 0, ctx->stream>>>(
            seed, sq_tlwe, pp->n, kn, pp->ks_d, pp->ks_log_b, ksd.rounding_bits, d_z, d_s, d_ksk);
        st = after_launch(ctx, "tfhe_kg_ksk_kernel");
    }
    if (st == FHE_OK) cu(cudaStreamSynchronize(ctx->stream), "sync");
    if (st == FHE_OK && brk_out) cu(cudaMemcpy(brk_out, d_raw, rows * (k + 1) * N * 8, cudaMemcpyDeviceToHost), "export");
    if (st == FHE_OK && (ksk_a_out || ksk_b_out)) {
        std::vector<uint64_t> h(ksk_rows * ld);
        cu(cudaMemcpy(h.data(), d_ksk, h.size() * 8, cudaMemcpyDeviceToHost), "export");
        for (size_t r = 0; r < ksk_rows; ++r) {
            if (ksk_a_out) memcpy(ksk_a_out + r * pp->n, h.data() + r * ld, (size_t)pp->n * 8);
            if (ksk_b_out) ksk_b_out[r] = h[r * ld + pp->n];
        }
    }
    if (st == FHE_OK) {
        st = tfhe_key_build(ctx, pp, nullptr, nullptr, nullptr, d_raw, d_ksk, out);
        if (st == FHE_OK) d_ksk = nullptr;  // adopted by the key
    }
    cudaStreamSynchronize(ctx->stream);
    for (void* p : {(void*)d_z, (void*)d_s, (void*)d_a, (void*)d_as, (void*)d_srep, (void*)d_raw, (void*)d_ksk})
        if (p) cudaFree(p);
    if (st == FHE_OK) {
        if (z_out) std::copy(z.begin(), z.end(), z_out);
        if (s_out) std::copy(s.begin(), s.end(), s_out);
    }
    return st;
}

// ---- serialised key: header | fhe_tfhe_param | bsk image (Fourier domain) | ksk image | ksk column sums |
// fused-path bsk image (when the parameters have one).  The images are the device buffers themselves: loading is four copies.
struct TfheBlobHeader {
    char magic[8];
    uint32_t version, kind;
    uint64_t param_bytes, brk_bytes, ksk_bytes, colsum_bytes, fast_bytes;
};
static const char TFHE_BLOB_MAGIC[8] = {'F', 'H', 'E', 'B', '2', '0', '0', 'K'};
size_t fhe_tfhe_key_serialized_size(const fhe_tfhe_key* key) {
    if (!key) return 0;
    return sizeof(TfheBlobHeader) + sizeof(fhe_tfhe_param) + key->brk_bytes + key->ksk_bytes + ((size_t)key->param.n + 1) * 8 +
           (key->fast_ok ? key->brk_fast_bytes : 0);
}
fhe_status fhe_tfhe_key_serialize(fhe_ctx* ctx, const fhe_tfhe_key* key, void* buf, size_t cap) {
    if (!ctx || !key || !buf) return FHE_EINVAL;
    FHE_REQUIRE(ctx, cap >= fhe_tfhe_key_serialized_size(key), "buffer too small for the serialised key");
    TfheBlobHeader h;
    memcpy(h.magic, TFHE_BLOB_MAGIC, 8);
    h.version = 1;
    h.kind = 2;
    h.param_bytes = sizeof(fhe_tfhe_param);
    h.brk_bytes = key->brk_bytes;
    h.ksk_bytes = key->ksk_bytes;
    h.colsum_bytes = ((size_t)key->param.n + 1) * 8;
    h.fast_bytes = key->fast_ok ? key->brk_fast_bytes : 0;
    unsigned char* p = (unsigned char*)buf;
    memcpy(p, &h, sizeof h);
    p += sizeof h;
    memcpy(p, &key->param, sizeof(fhe_tfhe_param));
    p += sizeof(fhe_tfhe_param);
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FHE_CUDA(ctx, cudaMemcpy(p, key->d_brk, h.brk_bytes, cudaMemcpyDeviceToHost));
    p += h.brk_bytes;
    FHE_CUDA(ctx, cudaMemcpy(p, key->d_ksk, h.ksk_bytes, cudaMemcpyDeviceToHost));
    p += h.ksk_bytes;
    FHE_CUDA(ctx, cudaMemcpy(p, key->d_ksk_colsum, h.colsum_bytes, cudaMemcpyDeviceToHost));
    p += h.colsum_bytes;
    if (h.fast_bytes) FHE_CUDA(ctx, cudaMemcpy(p, key->d_brk_fast, h.fast_bytes, cudaMemcpyDeviceToHost));
    return FHE_OK;
}
fhe_status fhe_tfhe_key_deserialize(fhe_ctx* ctx, const void* buf, size_t len, fhe_tfhe_key** out) {
    if (!ctx || !buf || !out) return FHE_EINVAL;
    *out = nullptr;
    TfheBlobHeader h;
    FHE_REQUIRE(ctx, len >= sizeof h, "serialised key truncated");
    memcpy(&h, buf, sizeof h);
    FHE_REQUIRE(ctx, memcmp(h.magic, TFHE_BLOB_MAGIC, 8) == 0 && h.kind == 2, "not a serialised TFHE key");
    FHE_REQUIRE(ctx, h.version == 1 && h.param_bytes == sizeof(fhe_tfhe_param), "serialised key of another format version");
    FHE_REQUIRE(ctx, len >= sizeof h + sizeof(fhe_tfhe_param), "serialised key truncated");
    const unsigned char* p = (const unsigned char*)buf + sizeof h;
    fhe_tfhe_param pp;
    memcpy(&pp, p, sizeof pp);
    p += sizeof pp;
    FHE_REQUIRE(ctx, h.brk_bytes <= len && h.ksk_bytes <= len && h.colsum_bytes <= len && h.fast_bytes <= len &&
                         len == sizeof h + sizeof pp + h.brk_bytes + h.ksk_bytes + h.colsum_bytes + h.fast_bytes,
                "serialised key length mismatch");
    TfheKeyImages img;
    img.brk = p;
    img.brk_bytes = h.brk_bytes;
    img.ksk = p + h.brk_bytes;
    img.ksk_bytes = h.ksk_bytes;
    img.colsum = p + h.brk_bytes + h.ksk_bytes;
    img.colsum_bytes = h.colsum_bytes;
    if (h.fast_bytes) {
        img.fast = p + h.brk_bytes + h.ksk_bytes + h.colsum_bytes;
        img.fast_bytes = h.fast_bytes;
    }
    return tfhe_key_build(ctx, &pp, nullptr, nullptr, nullptr, nullptr, nullptr, out, &img);
}

void fhe_tfhe_key_free(fhe_ctx* ctx, fhe_tfhe_key* key) {
    if (!key) return;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    if (key->d_brk) cudaFree(key->d_brk);
    if (key->d_ksk) cudaFree(key->d_ksk);
    if (key->d_ksk_colsum) cudaFree(key->d_ksk_colsum);
    if (key->d_brk_fast) cudaFree(key->d_brk_fast);
    if (key->d_fast_tab) cudaFree(key->d_fast_tab);
    delete key;
}
fhe_status fhe_tfhe_key_set_mode(fhe_ctx* ctx, fhe_tfhe_key* key, int mode) {
    if (!ctx || !key) return FHE_EINVAL;
    FHE_REQUIRE(ctx, mode >= 0 && mode <= 3,
                "mode must be 0 (reference dataflow, bit-identical), 1 (Fourier-domain accumulation), 2 (bounded-error fused path) or 3 "
                "(fused path with 32-bit accumulator words)");
    if (mode >= 2 && !key->fast_ok)
        return fail(ctx, FHE_EUNSUPPORTED, "the fused bounded-error path needs k = 1, log_b d <= 31, d <= 3 (2 at N = 2048) and N in {512, 1024, 2048}");
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    key->mode = mode;
    key->P.fourier_acc = mode != 0;  // the stand-alone external product / CMUX entry points use the generic kernels
    return FHE_OK;
}
size_t fhe_tfhe_key_bytes(const fhe_tfhe_key* key) { return key ? key->brk_bytes + key->ksk_bytes + key->brk_fast_bytes : 0; }
fhe_status fhe_tfhe_key_broadcast(fhe_ctx* ctx, fhe_tfhe_key* key, void* nccl_comm, int root) {
    if (!ctx || !key) return FHE_EINVAL;
    FHE_CHECK(fhe_keys_broadcast(ctx, nccl_comm, root, key->d_brk, key->brk_bytes));
    FHE_CHECK(fhe_keys_broadcast(ctx, nccl_comm, root, key->d_ksk, key->ksk_bytes));
    FHE_CHECK(fhe_keys_broadcast(ctx, nccl_comm, root, key->d_ksk_colsum, ((size_t)key->param.n + 1) * 8));
    if (key->fast_ok) FHE_CHECK(fhe_keys_broadcast(ctx, nccl_comm, root, key->d_brk_fast, key->brk_fast_bytes));
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FHE_OK;
}

fhe_status fhe_tfhe_blind_rotate_extract_batch(fhe_ctx* ctx, const fhe_tfhe_key* key, const uint64_t* d_lut, size_t count,
                                               const uint64_t* d_ct_in, uint64_t* d_out) {
    if (!ctx || !key) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_lut && d_ct_in && d_out, "null pointer");
    return run_blind_rotate(ctx, key, d_lut, count, d_ct_in, d_out);
}

fhe_status fhe_tlwe_key_switch_batch(fhe_ctx* ctx, const fhe_tfhe_key* key, size_t count, const uint64_t* d_ct_in, uint64_t* d_ct_out) {
    if (!ctx || !key) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_ct_in && d_ct_out, "null pointer");
    return run_key_switch(ctx, key, count, d_ct_in, d_ct_out);
}

fhe_status fhe_tfhe_pbs_batch(fhe_ctx* ctx, const fhe_tfhe_key* key, const uint64_t* d_lut, size_t count, const uint64_t* d_ct_in,
                              uint64_t* d_ct_out) {
    if (!ctx || !key) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_lut && d_ct_in && d_ct_out, "null pointer");
    const size_t ext_words = ((size_t)key->param.k << key->param.log_big_n) + 1;
    void* mid;
    FHE_CHECK(ensure_stage_d(ctx, 2, count * ext_words * 8, &mid));
    FHE_CHECK(run_blind_rotate(ctx, key, d_lut, count, d_ct_in, (uint64_t*)mid));
    return run_key_switch(ctx, key, count, (const uint64_t*)mid, d_ct_out);
}

fhe_status fhe_tfhe_pbs_batch_host(fhe_ctx* ctx, const fhe_tfhe_key* key, const uint64_t* lut, size_t count, const uint64_t* ct_in,
                                   uint64_t* ct_out) {
    if (!ctx || !key) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, lut && ct_in && ct_out, "null pointer");
    const size_t n = (size_t)1 << key->param.log_big_n, ct_bytes = count * (key->param.n + 1) * 8;
    void *d_in, *d_out;
    FHE_CHECK(ensure_stage_d(ctx, 0, ct_bytes + n * 8, &d_in));
    FHE_CHECK(ensure_stage_d(ctx, 1, ct_bytes, &d_out));
    uint64_t* d_lut = (uint64_t*)d_in + count * (key->param.n + 1);
    FHE_CUDA(ctx, cudaMemcpyAsync(d_in, ct_in, ct_bytes, cudaMemcpyHostToDevice, ctx->stream));
    FHE_CUDA(ctx, cudaMemcpyAsync(d_lut, lut, n * 8, cudaMemcpyHostToDevice, ctx->stream));
    FHE_CHECK(fhe_tfhe_pbs_batch(ctx, key, d_lut, count, (const uint64_t*)d_in, (uint64_t*)d_out));
    FHE_CUDA(ctx, cudaMemcpyAsync(ct_out, d_out, ct_bytes, cudaMemcpyDeviceToHost, ctx->stream));
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FHE_OK;
}

fhe_status fhe_tfhe_external_product(fhe_ctx* ctx, const fhe_tfhe_key* key, size_t count, const uint32_t* d_idx, const uint64_t* d_glwe_in,
                                     uint64_t* d_glwe_out) {
    if (!ctx || !key) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_idx && d_glwe_in && d_glwe_out && d_glwe_in != d_glwe_out, "null or aliased pointer");
    FHE_CHECK(validate_below(ctx, d_idx, count, key->P.n_lwe, "bootstrapping key index"));
    const size_t smem = tfhe_smem_bytes(key->P.k, key->P.bs_dec.d, key->P.log_n);
    unsigned grid;
    FHE_CHECK(tfhe_grid(ctx, tfhe_ext_kernel, smem, count, &grid));
    tfhe_ext_kernel<<<grid, TFHE_THREADS, smem, ctx->stream>>>(key->P, d_idx, d_glwe_in, nullptr, count, d_glwe_out);
    return after_launch(ctx, "tfhe_ext_kernel");
}

fhe_status fhe_tfhe_cmux(fhe_ctx* ctx, const fhe_tfhe_key* key, size_t count, const uint32_t* d_idx, const uint64_t* d_ct0, const uint64_t* d_ct1,
                         uint64_t* d_out) {
    if (!ctx || !key) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_idx && d_ct0 && d_ct1 && d_out && d_out != d_ct0 && d_out != d_ct1, "null or aliased pointer");
    FHE_CHECK(validate_below(ctx, d_idx, count, key->P.n_lwe, "bootstrapping key index"));
    const size_t smem = tfhe_smem_bytes(key->P.k, key->P.bs_dec.d, key->P.log_n);
    unsigned grid;
    FHE_CHECK(tfhe_grid(ctx, tfhe_ext_kernel, smem, count, &grid));
    tfhe_ext_kernel<<<grid, TFHE_THREADS, smem, ctx->stream>>>(key->P, d_idx, d_ct0, d_ct1, count, d_out);
    return after_launch(ctx, "tfhe_ext_kernel");
}

}  // extern "C"
