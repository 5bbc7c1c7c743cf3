// Element-wise / permutation / decomposition kernels of the util layer (K3-K7, K10, K12 of SURVEY.md §2) and the
// composed coefficient-form product.  HBM-bound streaming kernels: grid-stride, one 8-byte word per thread access
// (coalesced), grids sized to a multiple of the SM count.
#include <algorithm>

#include "ctx.cuh"
#include "fhew_core.cuh"

namespace fhe {

enum { OP_MUL = 0, OP_MAC = 1, OP_ADD = 2, OP_SUB = 3, OP_NEG = 4, OP_SCALAR = 5 };

template <int OP>
__global__ void __launch_bounds__(256) zq_elementwise_kernel(Mod64 m, unsigned long long count, const uint64_t* __restrict__ a,
                                                             const uint64_t* __restrict__ b, uint64_t scalar, uint64_t scalar_shoup,
                                                             uint64_t* __restrict__ out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        uint64_t x = a[i], r;
        if (OP == OP_MUL)
            r = m.mul(x, b[i]);
        else if (OP == OP_MAC)
            r = m.add(out[i], m.mul(x, b[i]));
        else if (OP == OP_ADD)
            r = m.add(x, b[i]);
        else if (OP == OP_SUB)
            r = m.sub(x, b[i]);
        else if (OP == OP_NEG)
            r = m.neg(x);
        else
            r = m.redq(m.shoup_lazy(x, scalar, scalar_shoup));
        out[i] = r;
    }
}

// avec.rs:34-50 (q != 0: Zq negation, q == 0: wrapping negation)
__global__ void __launch_bounds__(256) automorphism_kernel(uint64_t q, int log_n, unsigned long long batch, uint32_t t,
                                                           const uint64_t* __restrict__ in, uint64_t* __restrict__ out) {
    const unsigned long long total = batch << log_n, stride = (unsigned long long)gridDim.x * blockDim.x;
    const uint32_t n = 1u << log_n;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const unsigned long long b = idx >> log_n;
        const uint32_t i = (uint32_t)(idx & (n - 1));
        const uint32_t it = (uint32_t)(((unsigned long long)i * t) & (2ull * n - 1));
        uint64_t v = in[idx];
        if (it >= n) v = q ? (v == 0 ? 0 : q - v) : (uint64_t)(0 - v);
        out[(b << log_n) + (it & (n - 1))] = v;
    }
}
// ring.rs:299-313: out = in * X^k, k already reduced mod 2N
__global__ void __launch_bounds__(256) monomial_mul_kernel(uint64_t q, int log_n, unsigned long long batch, uint32_t k,
                                                           const uint64_t* __restrict__ in, uint64_t* __restrict__ out) {
    const unsigned long long total = batch << log_n, stride = (unsigned long long)gridDim.x * blockDim.x;
    const uint32_t n = 1u << log_n;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const unsigned long long b = idx >> log_n;
        const uint32_t j = (uint32_t)(idx & (n - 1));  // output index: gather form keeps the stores coalesced
        // out[j] = +-in[(j - k) mod 2N]
        const uint32_t src = (j + 2 * n - k) & (2 * n - 1);
        uint64_t v = in[(b << log_n) + (src & (n - 1))];
        if (src >= n) v = q ? (v == 0 ? 0 : q - v) : (uint64_t)(0 - v);
        out[idx] = v;
    }
}
__global__ void __launch_bounds__(256) mod_switch_kernel(uint64_t q, uint64_t qp, int odd, unsigned long long count,
                                                         const uint64_t* __restrict__ in, uint64_t* __restrict__ out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        out[i] = odd ? zq_mod_switch_odd_dev(in[i], q, qp) : zq_mod_switch_dev(in[i], q, qp);
}
__global__ void __launch_bounds__(256) decompose_zq_kernel(uint64_t q, DecompParam dp, unsigned long long count,
                                                           const uint64_t* __restrict__ in, uint64_t* __restrict__ out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        decompose_zq<uint64_t>(q, dp, in[i], [&](uint32_t k, uint64_t dg) { out[(unsigned long long)k * count + i] = dg; });
}
// decompose.rs:114-135
__device__ __forceinline__ uint64_t t64_rounding_shr(uint64_t v, unsigned bits) {
    if (bits >= 64) return 0;
    return (v + ((1ull << bits) >> 1)) >> bits;
}
__global__ void __launch_bounds__(256) decompose_t64_kernel(unsigned log_b, unsigned d, unsigned rounding_bits, unsigned long long count,
                                                            const uint64_t* __restrict__ in, uint64_t* __restrict__ out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    const uint64_t mask = (1ull << log_b) - 1;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride) {
        uint64_t v = t64_rounding_shr(in[i], rounding_bits);
        for (unsigned k = 0; k < d; ++k) {
            uint64_t limb = v & mask;
            v >>= log_b;
            uint64_t carry = (((limb - 1) | v) & limb) >> (log_b - 1);
            v += carry;
            out[(unsigned long long)k * count + i] = limb - (carry << log_b);
        }
    }
}
__global__ void __launch_bounds__(256) rounding_shr_t64_kernel(unsigned bits, unsigned long long count, const uint64_t* __restrict__ in,
                                                               uint64_t* __restrict__ out) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < count; i += stride)
        out[i] = t64_rounding_shr(in[i], bits);
}

static unsigned ew_grid(fhe_ctx* ctx, unsigned long long count) {
    unsigned long long blocks = (count + 255) / 256;
    unsigned long long cap = (unsigned long long)ctx->sm_count * 8;
    return (unsigned)std::max<unsigned long long>(1, std::min(blocks, cap));
}

template <int OP>
static fhe_status launch_ew(fhe_ctx* ctx, uint64_t q, size_t count, const uint64_t* a, const uint64_t* b, uint64_t scalar, uint64_t* out) {
    if (!ctx) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, q >= 2 && q < (1ull << 62), "modulus out of range for the u64 path");
    FHE_REQUIRE(ctx, a && out && (b || OP == OP_NEG || OP == OP_SCALAR), "null pointer");
    Mod64 m = make_mod<Mod64>(q);
    uint64_t sc = scalar % q;
    zq_elementwise_kernel<OP><<<ew_grid(ctx, count), 256, 0, ctx->stream>>>(m, count, a, b, sc, host_shoup64(sc, q), out);
    return after_launch(ctx, "zq_elementwise_kernel");
}

fhe_status pointwise_mul_dev(fhe_ctx* ctx, uint64_t q, size_t count, const uint64_t* a, const uint64_t* b, uint64_t* out) {
    return launch_ew<OP_MUL>(ctx, q, count, a, b, 0, out);
}

// Rgsw::internal_product (scheme/fhew/src/rgsw.rs:130-150), evaluation-domain dot products: output row (c, r), component h' is
//   sum over limbs (h, k) of e0[c][h d + k][h'] o dig[k][((c rows + r) 2 + h) n ..]      (all operands in evaluation form)
// e0: ct0 rows transformed [count][rows][2][n]; dig: fhe_decompose_zq of the whole ct1 array, limb-major [d][count rows 2 n].
__global__ void __launch_bounds__(256) rgsw_ip_mac_kernel(Mod64 m, uint32_t log_n, uint32_t d, unsigned long long count,
                                                          const uint64_t* __restrict__ e0, const uint64_t* __restrict__ dig,
                                                          uint64_t* __restrict__ out) {
    const uint32_t n = 1u << log_n, rows = 2 * d;
    const unsigned long long total = count * rows * n, stride = (unsigned long long)gridDim.x * blockDim.x;
    const unsigned long long plane = count * rows * 2ull * n;  // words per digit plane
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const uint32_t i = (uint32_t)(idx & (n - 1));
        const unsigned long long cr = idx >> log_n, c = cr / rows;
        uint64_t sa = 0, sb = 0;
        for (uint32_t h = 0; h < 2; ++h)
            for (uint32_t k = 0; k < d; ++k) {
                const uint64_t x = dig[k * plane + ((cr * 2 + h) << log_n) + i];
                const uint64_t* row = e0 + (((c * rows + h * d + k) * 2) << log_n) + i;
                sa = m.add(sa, m.mul(row[0], x));
                sb = m.add(sb, m.mul(row[n], x));
            }
        out[((cr * 2) << log_n) + i] = sa;
        out[((cr * 2 + 1) << log_n) + i] = sb;
    }
}

}  // namespace fhe

using namespace fhe;

extern "C" {

fhe_status fhe_pointwise_mul_u64(fhe_ctx* ctx, uint64_t q, size_t count, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out) {
    return launch_ew<OP_MUL>(ctx, q, count, d_a, d_b, 0, d_out);
}
fhe_status fhe_pointwise_mac_u64(fhe_ctx* ctx, uint64_t q, size_t count, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_acc) {
    return launch_ew<OP_MAC>(ctx, q, count, d_a, d_b, 0, d_acc);
}
fhe_status fhe_vec_add_u64(fhe_ctx* ctx, uint64_t q, size_t count, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out) {
    return launch_ew<OP_ADD>(ctx, q, count, d_a, d_b, 0, d_out);
}
fhe_status fhe_vec_sub_u64(fhe_ctx* ctx, uint64_t q, size_t count, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out) {
    return launch_ew<OP_SUB>(ctx, q, count, d_a, d_b, 0, d_out);
}
fhe_status fhe_vec_neg_u64(fhe_ctx* ctx, uint64_t q, size_t count, const uint64_t* d_a, uint64_t* d_out) {
    return launch_ew<OP_NEG>(ctx, q, count, d_a, nullptr, 0, d_out);
}
fhe_status fhe_vec_scalar_mul_u64(fhe_ctx* ctx, uint64_t q, size_t count, const uint64_t* d_a, uint64_t scalar, uint64_t* d_out) {
    return launch_ew<OP_SCALAR>(ctx, q, count, d_a, nullptr, scalar, d_out);
}

fhe_status fhe_automorphism_u64(fhe_ctx* ctx, uint64_t q, unsigned log_n, size_t batch, int64_t t, const uint64_t* d_in, uint64_t* d_out) {
    if (!ctx) return FHE_EINVAL;
    if (batch == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_in && d_out && d_in != d_out, "automorphism is out of place");
    FHE_REQUIRE(ctx, log_n <= 20, "log_n too large");
    const int64_t m = 2ll << log_n;
    const uint32_t tt = (uint32_t)(((t % m) + m) % m);
    automorphism_kernel<<<ew_grid(ctx, (unsigned long long)batch << log_n), 256, 0, ctx->stream>>>(q, (int)log_n, batch, tt, d_in, d_out);
    return after_launch(ctx, "automorphism_kernel");
}
fhe_status fhe_monomial_mul_u64(fhe_ctx* ctx, uint64_t q, unsigned log_n, size_t batch, int64_t k, const uint64_t* d_in, uint64_t* d_out) {
    if (!ctx) return FHE_EINVAL;
    if (batch == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_in && d_out && d_in != d_out, "monomial multiply is out of place");
    FHE_REQUIRE(ctx, log_n <= 20, "log_n too large");
    const int64_t m = 2ll << log_n;
    const uint32_t kk = (uint32_t)(((k % m) + m) % m);
    monomial_mul_kernel<<<ew_grid(ctx, (unsigned long long)batch << log_n), 256, 0, ctx->stream>>>(q, (int)log_n, batch, kk, d_in, d_out);
    return after_launch(ctx, "monomial_mul_kernel");
}
fhe_status fhe_mod_switch_u64(fhe_ctx* ctx, uint64_t q, uint64_t q_prime, size_t count, const uint64_t* d_in, uint64_t* d_out) {
    if (!ctx) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, q >= 1 && q_prime >= 1 && q_prime < (1ull << 62), "modulus out of range");
    mod_switch_kernel<<<ew_grid(ctx, count), 256, 0, ctx->stream>>>(q, q_prime, 0, count, d_in, d_out);
    return after_launch(ctx, "mod_switch_kernel");
}
fhe_status fhe_mod_switch_odd_u64(fhe_ctx* ctx, uint64_t q, uint64_t q_prime, size_t count, const uint64_t* d_in, uint64_t* d_out) {
    if (!ctx) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, q >= 1 && q_prime >= 1 && q_prime < (1ull << 62), "modulus out of range");
    mod_switch_kernel<<<ew_grid(ctx, count), 256, 0, ctx->stream>>>(q, q_prime, 1, count, d_in, d_out);
    return after_launch(ctx, "mod_switch_kernel");
}
fhe_status fhe_decompose_zq(fhe_ctx* ctx, uint64_t q, unsigned log_b, unsigned d, size_t count, const uint64_t* d_in, uint64_t* d_out) {
    if (!ctx) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, q >= 2 && log_b >= 1 && d >= 1 && log_b < 63 && (unsigned long long)log_b * d <= 64, "bad decomposor parameters");
    decompose_zq_kernel<<<ew_grid(ctx, count), 256, 0, ctx->stream>>>(q, make_decomp(q, log_b, d), count, d_in, d_out);
    return after_launch(ctx, "decompose_zq_kernel");
}
fhe_status fhe_decompose_t64(fhe_ctx* ctx, unsigned log_b, unsigned d, size_t count, const uint64_t* d_in, uint64_t* d_out) {
    if (!ctx) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, log_b >= 1 && d >= 1 && log_b < 64, "bad decomposor parameters");
    const unsigned rb = 64 > log_b * d ? 64 - log_b * d : 0;
    decompose_t64_kernel<<<ew_grid(ctx, count), 256, 0, ctx->stream>>>(log_b, d, rb, count, d_in, d_out);
    return after_launch(ctx, "decompose_t64_kernel");
}
fhe_status fhe_rounding_shr_t64(fhe_ctx* ctx, unsigned bits, size_t count, const uint64_t* d_in, uint64_t* d_out) {
    if (!ctx) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    rounding_shr_t64_kernel<<<ew_grid(ctx, count), 256, 0, ctx->stream>>>(bits, count, d_in, d_out);
    return after_launch(ctx, "rounding_shr_t64_kernel");
}

// nega_cyclic_ntt_mul_assign (fft/zq.rs:14-25): NTT(a), NTT(copy of b), pointwise, iNTT — b is left untouched
fhe_status fhe_negacyclic_mul_u64(fhe_ctx* ctx, uint64_t q, unsigned log_n, size_t batch, uint64_t* d_a, const uint64_t* d_b) {
    if (!ctx) return FHE_EINVAL;
    if (batch == 0) return FHE_OK;
    const size_t count = batch << log_n;
    void* tmp;
    FHE_CHECK(ensure_scratch(ctx, count * 8, &tmp));
    FHE_CUDA(ctx, cudaMemcpyAsync(tmp, d_b, count * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    FHE_CHECK(launch_ntt_u64(ctx, q, log_n, batch, d_a, true));
    FHE_CHECK(launch_ntt_u64(ctx, q, log_n, batch, (uint64_t*)tmp, true));
    FHE_CHECK(pointwise_mul_dev(ctx, q, count, d_a, (const uint64_t*)tmp, d_a));
    return launch_ntt_u64(ctx, q, log_n, batch, d_a, false);
}
fhe_status fhe_negacyclic_mul_host(fhe_ctx* ctx, uint64_t q, uint64_t* a, const uint64_t* b, size_t n, size_t batch) {
    if (!ctx || !a || !b) return FHE_EINVAL;
    FHE_REQUIRE(ctx, n && !(n & (n - 1)), "polynomial length %zu is not a power of two", n);
    if (batch == 0) return FHE_OK;
    unsigned lg = 0;
    while (((size_t)1 << lg) < n) ++lg;
    const size_t bytes = n * batch * 8;
    uint64_t *da = nullptr, *db = nullptr;
    fhe_status st = FHE_OK;
    if (cudaMalloc(&da, bytes) != cudaSuccess || cudaMalloc(&db, bytes) != cudaSuccess) st = fail(ctx, FHE_ENOMEM, "device allocation failed");
    if (st == FHE_OK && (cudaMemcpyAsync(da, a, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess ||
                         cudaMemcpyAsync(db, b, bytes, cudaMemcpyHostToDevice, ctx->stream) != cudaSuccess))
        st = fail(ctx, FHE_ECUDA, "h2d failed");
    if (st == FHE_OK) st = fhe_negacyclic_mul_u64(ctx, q, lg, batch, da, db);
    if (st == FHE_OK && cudaMemcpyAsync(a, da, bytes, cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) st = fail(ctx, FHE_ECUDA, "d2h failed");
    cudaStreamSynchronize(ctx->stream);
    cudaFree(da);
    cudaFree(db);
    return st;
}

// Rgsw::internal_product (rgsw.rs:130-150) on `count` pairs of RGSW ciphertexts [count][2d rows][2 (a, b)][n] over Z_q, coefficient
// form in and out: ct0's rows -> evaluation form, every row of ct1 decomposed (a digits then b digits, decompose.rs), limbs
// transformed, dotted with ct0's a / b columns, transformed back.  out must not alias the inputs.
fhe_status fhe_rgsw_internal_product(fhe_ctx* ctx, uint64_t q, unsigned log_n, unsigned log_b, unsigned d, size_t count, const uint64_t* d_ct0,
                                     const uint64_t* d_ct1, uint64_t* d_out) {
    if (!ctx) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_ct0 && d_ct1 && d_out && d_out != d_ct0 && d_out != d_ct1, "null or aliased pointer");
    FHE_REQUIRE(ctx, q >= 2 && q < (1ull << 62) && host_is_prime(q), "internal product needs an NTT-friendly prime modulus < 2^62");
    FHE_REQUIRE(ctx, log_n >= 1 && log_n <= 16 && log_b >= 1 && d >= 1 && log_b < 63 && (unsigned long long)log_b * d <= 64, "bad parameters");
    const size_t n = (size_t)1 << log_n, rows = 2 * (size_t)d, words = count * rows * 2 * n;
    void* ws;
    FHE_CHECK(ensure_scratch(ctx, (1 + (size_t)d) * words * 8, &ws));
    uint64_t* e0 = (uint64_t*)ws;
    uint64_t* dig = e0 + words;
    FHE_CUDA(ctx, cudaMemcpyAsync(e0, d_ct0, words * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    FHE_CHECK(launch_ntt_u64(ctx, q, log_n, count * rows * 2, e0, true));
    decompose_zq_kernel<<<ew_grid(ctx, words), 256, 0, ctx->stream>>>(q, make_decomp(q, log_b, d), words, d_ct1, dig);
    FHE_CHECK(after_launch(ctx, "decompose_zq_kernel"));
    FHE_CHECK(launch_ntt_u64(ctx, q, log_n, (size_t)d * count * rows * 2, dig, true));
    rgsw_ip_mac_kernel<<<ew_grid(ctx, count * rows * n), 256, 0, ctx->stream>>>(make_mod<Mod64>(q), log_n, d, count, e0, dig, d_out);
    FHE_CHECK(after_launch(ctx, "rgsw_ip_mac_kernel"));
    return launch_ntt_u64(ctx, q, log_n, count * rows * 2, d_out, false);
}

}  // extern "C"
