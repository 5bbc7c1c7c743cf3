// Key generation on the device.  fhe_fhew_keygen evaluates Bootstrapping::key_gen
// (scheme/fhew/src/bootstrapping.rs:122-146 with lwe.rs:108-119,130-140, rlwe.rs:109-156, rgsw.rs:84-105) for a whole key on
// the GPU: uniform masks and Gaussian errors come from the counter-based stream of keygen_stream.cuh (one independent word per
// (domain, index), so threads may produce them in any order), the products a * z run through the batched NTT, and the rows go
// straight into the evaluation-form key buffers without ever visiting the host.  oracle/orc_keygen.hpp evaluates the same
// formulas on the same stream on the CPU: the keys are equal word for word (tests/test_gpu_keygen.py).
#include <algorithm>
#include <vector>

#include "fhew_core.cuh"
#include "fhew_key.cuh"
#include "keygen_stream.cuh"

namespace fhe {

// rows [rows][2][n] of W words: a = uniform (domain da), b = e (domain de) as a residue; a also copied to a64 [rows][n] for the product
template <typename W>
__global__ void __launch_bounds__(256) kg_rows_kernel(uint64_t seed, uint32_t da, uint32_t de, uint64_t q, uint32_t log_n, unsigned long long rows,
                                                      W* __restrict__ out, uint64_t* __restrict__ a64) {
    const unsigned long long total = rows << log_n, stride = (unsigned long long)gridDim.x * blockDim.x;
    const uint32_t n = 1u << log_n;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        const unsigned long long r = i >> log_n;
        const uint32_t c = (uint32_t)(i & (n - 1));
        const uint64_t a = ks_uniform(seed, da, i, q);
        const int64_t e = ks_gauss(seed, de, i);
        out[((r * 2) << log_n) + c] = (W)a;
        out[((r * 2 + 1) << log_n) + c] = (W)(e < 0 ? q - (uint64_t)(-e) : (uint64_t)e);
        a64[i] = a;
    }
}
// a64[r][c] <- a64[r][c] * z_eval[c]  (evaluation domain)
__global__ void __launch_bounds__(256) kg_mul_bcast_kernel(Mod64 m, uint32_t log_n, unsigned long long total, uint64_t* __restrict__ a64,
                                                           const uint64_t* __restrict__ z_eval) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride)
        a64[i] = m.mul(a64[i], z_eval[i & ((1u << log_n) - 1)]);
}
// rows[r].b += a64[r]  (a * z, coefficient form)
template <typename W>
__global__ void __launch_bounds__(256) kg_add_b_kernel(uint64_t q, uint32_t log_n, unsigned long long total, W* __restrict__ out,
                                                       const uint64_t* __restrict__ az) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const uint32_t n = 1u << log_n;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
        W* b = out + (((i >> log_n) * 2 + 1) << log_n) + (i & (n - 1));
        const uint64_t v = (uint64_t)*b + az[i];
        *b = (W)(v >= q ? v - q : v);
    }
}
// RGSW plaintext X^{s_j} B^k: one coefficient per row, added to a (rows k < d) or b (rows d + k)   (rgsw.rs:101-103)
template <typename W>
__global__ void kg_brk_pt_kernel(uint64_t q, uint32_t log_n, uint32_t n_s, uint32_t d, const int64_t* __restrict__ s, const uint64_t* __restrict__ bases,
                                 W* __restrict__ rows) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n_s * d) return;
    const uint32_t j = t / d, k = t % d, n = 1u << log_n;
    const uint32_t e = (uint32_t)(((s[j] % (int64_t)(2 * n)) + 2 * n) % (2 * n));  // ring.rs:299-313: X^e, e mod 2N
    const uint32_t pos = e & (n - 1);
    const uint64_t v = e < n ? bases[k] : (bases[k] ? q - bases[k] : 0);
    W* ra = rows + ((((size_t)j * 2 * d + k) * 2) << log_n) + pos;
    W* rb = rows + ((((size_t)j * 2 * d + d + k) * 2 + 1) << log_n) + pos;
    uint64_t x = (uint64_t)*ra + v;
    *ra = (W)(x >= q ? x - q : x);
    x = (uint64_t)*rb + v;
    *rb = (W)(x >= q ? x - q : x);
}
// automorphism-key plaintext (-z(X^t)) B^k added to b of row (v, k)   (rlwe.rs:109-132; avec.rs:34-50 on the i64 secret)
template <typename W>
__global__ void __launch_bounds__(256) kg_ak_pt_kernel(Mod64 m, uint32_t log_n, uint32_t nv, uint32_t d, const int64_t* __restrict__ z,
                                                       const uint32_t* __restrict__ ak_t, const uint64_t* __restrict__ bases, W* __restrict__ rows) {
    const uint32_t n = 1u << log_n;
    const unsigned long long total = (unsigned long long)nv * d * n, stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const uint32_t i = (uint32_t)(idx & (n - 1));
        const uint32_t r = (uint32_t)(idx >> log_n), v = r / d, k = r % d;
        const uint32_t it = (uint32_t)(((unsigned long long)i * ak_t[v]) & (2ull * n - 1));
        const int64_t za = it < n ? z[i] : -z[i];                           // z(X^t) at coefficient it mod n
        const int64_t neg = -za;
        const uint64_t res = neg < 0 ? m.q - (uint64_t)(-neg) % m.q : (uint64_t)neg % m.q;
        const uint64_t pt = m.mul(bases[k], res == m.q ? 0 : res);
        W* b = rows + (((size_t)r * 2 + 1) << log_n) + (it & (n - 1));
        const uint64_t x = (uint64_t)*b + pt;
        *b = (W)(x >= m.q ? x - m.q : x);
    }
}
// LWE key-switching key (lwe.rs:108-119, 130-140), q_ks a power of two <= 2^32: row idx = digit * N + coefficient,
// packed [idx][n_s + 1] u32 = (a_0 .. a_{n_s-1}, b), b = <a, s> + base_k (-z_i) + e
__global__ void __launch_bounds__(128) kg_ksk_kernel(uint64_t seed, uint64_t q_ks, uint32_t log_n, uint32_t n_s, uint32_t ks_d, const int64_t* __restrict__ z,
                                                     const int64_t* __restrict__ s, const uint64_t* __restrict__ bases, uint32_t* __restrict__ ksk) {
    const uint32_t n = 1u << log_n;
    const uint64_t mask = q_ks - 1;
    for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < n * ks_d; idx += gridDim.x * blockDim.x) {
        const uint32_t k = idx >> log_n, i = idx & (n - 1);
        uint64_t dot = 0;
        uint32_t* row = ksk + (size_t)idx * (n_s + 1);
        for (uint32_t j = 0; j < n_s; ++j) {
            const uint64_t a = ks_uniform(seed, KS_FHEW_KSK_A, (uint64_t)idx * n_s + j, q_ks);
            row[j] = (uint32_t)a;
            dot += a * (uint64_t)s[j];  // two's complement wrap: exact modulo the power of two q_ks
        }
        const uint64_t pt = bases[k] * (uint64_t)(-z[i]);
        row[n_s] = (uint32_t)((dot + pt + (uint64_t)ks_gauss(seed, KS_FHEW_KSK_E, idx)) & mask);
    }
}

static std::vector<uint64_t> decomp_bases(uint64_t q, unsigned log_b, unsigned d) {  // Base2Decomposor<Zq>::new (decompose.rs:49-64)
    const DecompParam dp = make_decomp(q, log_b, d);
    std::vector<uint64_t> b(d);
    for (unsigned k = 0; k < d; ++k) b[k] = (uint64_t)((((u128_t)1) << (dp.rounding_bits + log_b * k)) % q);
    return b;
}

template <typename W>
static fhe_status fhew_keygen_t(fhe_ctx* ctx, const fhe_fhew_param* pp, uint64_t seed, int64_t* z_out, int64_t* s_out, uint64_t* ksk_a_out,
                                uint64_t* ksk_b_out, uint64_t* brk_out, uint64_t* ak_out, fhe_fhew_key** out) {
    const uint32_t log_n = pp->log_n, n = 1u << log_n, nv = pp->w + 1;
    const uint64_t q = pp->big_q;
    const size_t brk_rows = (size_t)pp->n_s * 2 * pp->rgsw_d, ak_rows = (size_t)nv * pp->rlwe_d, max_rows = std::max(brk_rows, ak_rows);
    // secrets: tiny, drawn on the host from the same stream and uploaded (they are returned to the caller anyway)
    std::vector<int64_t> z(n), s(pp->n_s);
    for (uint32_t i = 0; i < n; ++i) z[i] = ks_gauss(seed, KS_FHEW_Z, i);
    for (uint32_t j = 0; j < pp->n_s; ++j) s[j] = ks_gauss(seed, KS_FHEW_S, j);
    std::vector<uint32_t> ak_t(nv);
    std::vector<int64_t> ak_t_signed(nv);
    {
        const uint32_t m2 = 2 * n;
        uint64_t g = 5 % m2, pw = 1;
        ak_t[0] = m2 - (uint32_t)g;  // -g   (bootstrapping.rs:86-89)
        for (uint32_t v = 1; v < nv; ++v) {
            pw = pw * g % m2;
            ak_t[v] = (uint32_t)pw;
        }
        for (uint32_t v = 0; v < nv; ++v) ak_t_signed[v] = ak_t[v] < n ? (int64_t)ak_t[v] : (int64_t)ak_t[v] - (int64_t)m2;  // Zq::to_i64 (zq.rs:71-77)
    }
    const std::vector<uint64_t> gb = decomp_bases(q, pp->rgsw_log_b, pp->rgsw_d), rb = decomp_bases(q, pp->rlwe_log_b, pp->rlwe_d),
                                kb = decomp_bases(pp->q_ks, pp->ks_log_b, pp->ks_d);
    // device scratch: z, s, ak_t, bases, z_eval, a64
    int64_t *d_z = nullptr, *d_s = nullptr;
    uint32_t* d_akt = nullptr;
    uint64_t *d_bases = nullptr, *d_zeval = nullptr, *d_a64 = nullptr;
    W *d_brk = nullptr, *d_ak = nullptr;
    uint32_t* d_ksk = nullptr;
    fhe_status st = FHE_OK;
    auto cu = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && st == FHE_OK) st = fail(ctx, FHE_ECUDA, "keygen %s: %s", what, cudaGetErrorString(e));
    };
    const size_t nb = gb.size() + rb.size() + kb.size();
    std::vector<uint64_t> bases(gb);
    bases.insert(bases.end(), rb.begin(), rb.end());
    bases.insert(bases.end(), kb.begin(), kb.end());
    std::vector<uint64_t> zres(n);
    for (uint32_t i = 0; i < n; ++i) zres[i] = z[i] < 0 ? q - (uint64_t)(-z[i]) : (uint64_t)z[i];
    cu(cudaMalloc(&d_z, n * 8), "alloc");
    cu(cudaMalloc(&d_s, pp->n_s * 8), "alloc");
    cu(cudaMalloc(&d_akt, nv * 4), "alloc");
    cu(cudaMalloc(&d_bases, nb * 8), "alloc");
    cu(cudaMalloc(&d_zeval, n * 8), "alloc");
    cu(cudaMalloc(&d_a64, max_rows * n * 8), "alloc");
    cu(cudaMalloc(&d_brk, brk_rows * 2 * n * sizeof(W)), "alloc");
    cu(cudaMalloc(&d_ak, ak_rows * 2 * n * sizeof(W)), "alloc");
    const size_t ksk_words = (size_t)n * pp->ks_d * (pp->n_s + 1);
    cu(cudaMalloc(&d_ksk, ksk_words * 4), "alloc");
    if (st == FHE_OK) {
        cu(cudaMemcpyAsync(d_z, z.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream), "copy");
        cu(cudaMemcpyAsync(d_s, s.data(), pp->n_s * 8, cudaMemcpyHostToDevice, ctx->stream), "copy");
        cu(cudaMemcpyAsync(d_akt, ak_t.data(), nv * 4, cudaMemcpyHostToDevice, ctx->stream), "copy");
        cu(cudaMemcpyAsync(d_bases, bases.data(), nb * 8, cudaMemcpyHostToDevice, ctx->stream), "copy");
        cu(cudaMemcpyAsync(d_zeval, zres.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream), "copy");
    }
    if (st == FHE_OK) st = launch_ntt_u64(ctx, q, log_n, 1, d_zeval, true);
    const Mod64 m = make_mod<Mod64>(q);
    auto grid = [&](unsigned long long total) { return (unsigned)std::min<unsigned long long>((total + 255) / 256, (unsigned long long)ctx->sm_count * 16); };
    // rows: a uniform, b = a * z + e   (rlwe.rs:146-156)
    auto make_rows = [&](W* d_rows, size_t rows, uint32_t da, uint32_t de) {
        const unsigned long long total = (unsigned long long)rows << log_n;
        kg_rows_kernel<W><<<grid(total), 256, 0, ctx->stream>>>(seed, da, de, q, log_n, rows, d_rows, d_a64);
        if (st == FHE_OK) st = after_launch(ctx, "kg_rows_kernel");
        if (st == FHE_OK) st = launch_ntt_u64(ctx, q, log_n, rows, d_a64, true);
        kg_mul_bcast_kernel<<<grid(total), 256, 0, ctx->stream>>>(m, log_n, total, d_a64, d_zeval);
        if (st == FHE_OK) st = after_launch(ctx, "kg_mul_bcast_kernel");
        if (st == FHE_OK) st = launch_ntt_u64(ctx, q, log_n, rows, d_a64, false);
        kg_add_b_kernel<W><<<grid(total), 256, 0, ctx->stream>>>(q, log_n, total, d_rows, d_a64);
        if (st == FHE_OK) st = after_launch(ctx, "kg_add_b_kernel");
    };
    if (st == FHE_OK) {
        make_rows(d_brk, brk_rows, KS_FHEW_BRK_A, KS_FHEW_BRK_E);
        kg_brk_pt_kernel<W><<<(pp->n_s * pp->rgsw_d + 127) / 128, 128, 0, ctx->stream>>>(q, log_n, pp->n_s, pp->rgsw_d, d_s, d_bases, d_brk);
        if (st == FHE_OK) st = after_launch(ctx, "kg_brk_pt_kernel");
    }
    if (st == FHE_OK) {
        make_rows(d_ak, ak_rows, KS_FHEW_AK_A, KS_FHEW_AK_E);
        kg_ak_pt_kernel<W><<<grid((unsigned long long)ak_rows << log_n), 256, 0, ctx->stream>>>(m, log_n, nv, pp->rlwe_d, d_z, d_akt, d_bases + gb.size(), d_ak);
        if (st == FHE_OK) st = after_launch(ctx, "kg_ak_pt_kernel");
    }
    if (st == FHE_OK) {
        kg_ksk_kernel<<<(n * pp->ks_d + 127) / 128, 128, 0, ctx->stream>>>(seed, pp->q_ks, log_n, pp->n_s, pp->ks_d, d_z, d_s, d_bases + gb.size() + rb.size(), d_ksk);
        st = after_launch(ctx, "kg_ksk_kernel");
    }
    if (st == FHE_OK) cu(cudaStreamSynchronize(ctx->stream), "sync");
    // optional export of the coefficient-form key in the reference layout (parity tests; a production caller passes null)
    if (st == FHE_OK && (brk_out || ak_out)) {
        std::vector<W> h(std::max(brk_rows, ak_rows) * 2 * n);
        if (brk_out) {
            cu(cudaMemcpy(h.data(), d_brk, brk_rows * 2 * n * sizeof(W), cudaMemcpyDeviceToHost), "export");
            for (size_t i = 0; i < brk_rows * 2 * n; ++i) brk_out[i] = h[i];
        }
        if (ak_out) {
            cu(cudaMemcpy(h.data(), d_ak, ak_rows * 2 * n * sizeof(W), cudaMemcpyDeviceToHost), "export");
            for (size_t i = 0; i < ak_rows * 2 * n; ++i) ak_out[i] = h[i];
        }
    }
    if (st == FHE_OK && (ksk_a_out || ksk_b_out)) {
        std::vector<uint32_t> h(ksk_words);
        cu(cudaMemcpy(h.data(), d_ksk, ksk_words * 4, cudaMemcpyDeviceToHost), "export");
        const size_t ld = pp->n_s + 1;
        for (size_t idx = 0; idx < (size_t)n * pp->ks_d; ++idx) {
            if (ksk_a_out)
                for (size_t j = 0; j < pp->n_s; ++j) ksk_a_out[idx * pp->n_s + j] = h[idx * ld + j];
            if (ksk_b_out) ksk_b_out[idx] = h[idx * ld + pp->n_s];
        }
    }
    if (st == FHE_OK) {
        FhewKeySource src;
        src.dev_brk_rows = d_brk;
        src.dev_ak_rows = d_ak;
        src.dev_ksk = d_ksk;
        st = fhew_key_build(ctx, pp, ak_t_signed.data(), src, out);
        if (st == FHE_OK) d_ksk = nullptr;  // adopted by the key
    }
    cudaStreamSynchronize(ctx->stream);
    for (void* p : {(void*)d_z, (void*)d_s, (void*)d_akt, (void*)d_bases, (void*)d_zeval, (void*)d_a64, (void*)d_brk, (void*)d_ak, (void*)d_ksk})
        if (p) cudaFree(p);
    if (st == FHE_OK) {
        if (z_out) std::copy(z.begin(), z.end(), z_out);
        if (s_out) std::copy(s.begin(), s.end(), s_out);
    }
    return st;
}

}  // namespace fhe

using namespace fhe;

extern "C" fhe_status fhe_fhew_keygen(fhe_ctx* ctx, const fhe_fhew_param* pp, uint64_t seed, int64_t* z_out, int64_t* s_out, uint64_t* ksk_a_out,
                                      uint64_t* ksk_b_out, uint64_t* brk_out, uint64_t* ak_out, fhe_fhew_key** out) {
    if (!ctx || !pp || !out) return FHE_EINVAL;
    *out = nullptr;
    FHE_REQUIRE(ctx, pp->log_n >= 2 && pp->log_n <= 11 && pp->big_q >= 3 && pp->big_q < (1ull << 62) && host_is_prime(pp->big_q),
                "key generation needs 4 <= N <= 2048 and a prime Q < 2^62");
    FHE_REQUIRE(ctx, pp->q_ks >= 2 && pp->q_ks <= (1ull << 32) && (pp->q_ks & (pp->q_ks - 1)) == 0, "q_ks must be a power of two <= 2^32");
    FHE_REQUIRE(ctx, pp->w >= 1 && pp->w < 40 && pp->n_s >= 1 && pp->rgsw_d >= 1 && pp->rlwe_d >= 1 && pp->ks_d >= 1, "bad parameters");
    return pp->big_q >= (1ull << 30) ? fhew_keygen_t<uint64_t>(ctx, pp, seed, z_out, s_out, ksk_a_out, ksk_b_out, brk_out, ak_out, out)
                                     : fhew_keygen_t<uint32_t>(ctx, pp, seed, z_out, s_out, ksk_a_out, ksk_b_out, brk_out, ak_out, out);
}
