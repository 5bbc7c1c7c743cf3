// Pipe-peak microbenchmarks: measured INT32 multiply throughput of the device the context
// runs on.  The modular-NTT kernels are bound by the fma pipe's integer multiply rate (IMAD / IMAD.HI / IMAD.WIDE),
// for which MEASURED_PEAKS.json carries no figure; bench.py divides the blind-rotation kernel's algorithmic
// multiply count by these numbers.
#include "ctx.cuh"

namespace fhe {

// mode 0: 32-bit IMAD (x = x * a + b); mode 1: IMAD.HI (__umulhi); mode 2: IMAD.WIDE (u32 x u32 + u64)
template <int MODE>
__global__ void __launch_bounds__(256) imad_peak_kernel(uint32_t a, uint32_t b, int iters, uint32_t* __restrict__ sink) {
    constexpr int ILP = 8;
    uint32_t x[ILP];
    uint64_t w[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) {
        x[j] = threadIdx.x * 2654435761u + j * 40503u + blockIdx.x;
        w[j] = x[j];
    }
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                if (MODE == 0)
                    x[j] = x[j] * a + b;
                else if (MODE == 1)
                    x[j] = __umulhi(x[j], a) + b;
                else
                    w[j] = (uint64_t)(uint32_t)w[j] * a + w[j];
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc ^= x[j] ^ (uint32_t)w[j] ^ (uint32_t)(w[j] >> 32);
    if (acc == 0x12345u) sink[0] = acc;  // keeps the chains alive without a store in the common case
}

template <int MODE>
static fhe_status run_peak(fhe_ctx* ctx, double* tops) {
    const int iters = 4096;
    const unsigned grid = (unsigned)ctx->sm_count * 8;
    void* sink;
    FHE_CHECK(ensure_scratch(ctx, 256, &sink));
    cudaEvent_t e0, e1;
    FHE_CUDA(ctx, cudaEventCreate(&e0));
    FHE_CUDA(ctx, cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        FHE_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        imad_peak_kernel<MODE><<<grid, 256, 0, ctx->stream>>>(0x9E3779B1u, 12345u, iters, (uint32_t*)sink);
        FHE_CHECK(after_launch(ctx, "imad_peak_kernel"));
        FHE_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        FHE_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        FHE_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        double ops = (double)grid * 256 * (double)iters * 8 * 8;
        double t = ops / (ms * 1e-3) / 1e12;
        if (rep > 0 && t > best) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tops = best;
    return FHE_OK;
}

// FP64 pipe: mode 0 DADD, 1 DMUL, 2 DFMA (8 independent chains per thread; values stay finite: |x| oscillates around 1)
template <int MODE>
__global__ void __launch_bounds__(256) fp64_peak_kernel(double a, double b, int iters, double* __restrict__ sink) {
    constexpr int ILP = 8;
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = 1.0 + 1e-3 * (double)((threadIdx.x + j * 37 + blockIdx.x) & 255);
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int r = 0; r < 8; ++r) {
#pragma unroll
            for (int j = 0; j < ILP; ++j) {
                if (MODE == 0)
                    x[j] = __dadd_rn(x[j], (r & 1) ? a : -a);
                else if (MODE == 1)
                    x[j] = __dmul_rn(x[j], (r & 1) ? b : 1.0 / b);
                else
                    x[j] = __fma_rn(x[j], (r & 1) ? b : 1.0 / b, (r & 1) ? a : -a);
            }
        }
    }
    double acc = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) acc += x[j];
    if (acc == 0.12345) sink[0] = acc;
}
template <int MODE>
static fhe_status run_fp64_peak(fhe_ctx* ctx, double* tops) {
    const int iters = 2048;
    const unsigned grid = (unsigned)ctx->sm_count * 8;
    void* sink;
    FHE_CHECK(ensure_scratch(ctx, 256, &sink));
    cudaEvent_t e0, e1;
    FHE_CUDA(ctx, cudaEventCreate(&e0));
    FHE_CUDA(ctx, cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        FHE_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        fp64_peak_kernel<MODE><<<grid, 256, 0, ctx->stream>>>(0.5, 1.0000001, iters, (double*)sink);
        FHE_CHECK(after_launch(ctx, "fp64_peak_kernel"));
        FHE_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        FHE_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        FHE_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double ops = (double)grid * 256 * (double)iters * 8 * 8;
        const double t = ops / (ms * 1e-3) / 1e12;
        if (rep > 0 && t > best) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tops = best;
    return FHE_OK;
}


// Radix-16 register passes of lazy u32 butterflies (the FHEW / u32 NTT inner loop), two ways of forming the Shoup quotient
// floor(wp * y / 2^32): MODE 0 IMAD.HI (multiply pipe, half rate); MODE 1 one DFMA rounded down on the FP64 pipe
// (mulhi_u32_f64, modarith.cuh).  Twiddles come from shared memory like in the real passes.
template <int MODE>
__global__ void __launch_bounds__(128) bf_rate_kernel(uint32_t q, int iters, uint32_t* __restrict__ sink) {
    __shared__ TwPair<uint32_t> tw[256];
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
        const uint32_t w = (i * 2654435761u + 12345u) % q;
        tw[i] = TwPair<uint32_t>{w, (uint32_t)(((uint64_t)w << 32) / q)};
    }
    __syncthreads();
    const uint32_t q2 = 2 * q, q8 = 8 * q;
    uint32_t x[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) x[j] = (threadIdx.x * 977u + j * 131u + blockIdx.x) % q;
    for (int it = 0; it < iters; ++it) {
        const uint32_t tb = (uint32_t)(it * 5 + (threadIdx.x >> 5)) & 15u;  // warp-uniform: broadcast reads, like the real passes
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int h = 1 << (3 - u);
#pragma unroll
            for (int top = 0; top < (1 << u); ++top) {
                const TwPair<uint32_t> t = tw[((tb << u) + top) & 255u];
                F64Quot fq;
                if (MODE == 1) fq = make_f64_quot(t.wp);
#pragma unroll
                for (int low = 0; low < h; ++low) {
                    const int j = (top << (4 - u)) | low;
                    const uint32_t y = x[j + h], x0 = x[j];
                    const uint32_t hi = MODE == 1 ? mulhi_u32_f64(fq, y) : mulhi_u32(t.wp, y);
                    const uint32_t ty = t.w * y - hi * q;
                    x[j] = alu_add(x0, ty);
                    x[j + h] = x0 + q2 - ty;
                }
            }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) x[j] = umin_(x[j], x[j] - q8);
    }
    uint32_t acc = 0;
#pragma unroll
    for (int j = 0; j < 16; ++j) acc ^= x[j];
    sink[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE>
static fhe_status run_bf_rate(fhe_ctx* ctx, double* tbf, uint32_t* checksum) {
    const int iters = 512;
    const unsigned grid = (unsigned)ctx->sm_count * 8;
    void* sink;
    FHE_CHECK(ensure_scratch(ctx, (size_t)grid * 128 * 4, &sink));
    cudaEvent_t e0, e1;
    FHE_CUDA(ctx, cudaEventCreate(&e0));
    FHE_CUDA(ctx, cudaEventCreate(&e1));
    double best = 0;
    for (int rep = 0; rep < 4; ++rep) {
        FHE_CUDA(ctx, cudaEventRecord(e0, ctx->stream));
        bf_rate_kernel<MODE><<<grid, 128, 0, ctx->stream>>>(268369921u, iters, (uint32_t*)sink);
        FHE_CHECK(after_launch(ctx, "bf_rate_kernel"));
        FHE_CUDA(ctx, cudaEventRecord(e1, ctx->stream));
        FHE_CUDA(ctx, cudaEventSynchronize(e1));
        float ms = 0;
        FHE_CUDA(ctx, cudaEventElapsedTime(&ms, e0, e1));
        const double t = (double)grid * 128 * (double)iters * 32 / (ms * 1e-3) / 1e12;
        if (rep > 0 && t > best) best = t;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    std::vector<uint32_t> h((size_t)grid * 128);
    FHE_CUDA(ctx, cudaMemcpy(h.data(), sink, h.size() * 4, cudaMemcpyDeviceToHost));
    uint32_t c = 0;
    for (uint32_t v : h) c = c * 31u + v;
    *checksum = c;
    *tbf = best;
    return FHE_OK;
}

}  // namespace fhe

using namespace fhe;

extern "C" fhe_status fhe_diag_fp64_peak(fhe_ctx* ctx, double* dadd_tops, double* dmul_tops, double* dfma_tops) {
    if (!ctx) return FHE_EINVAL;
    double a = 0, b = 0, c = 0;
    FHE_CHECK(run_fp64_peak<0>(ctx, &a));
    FHE_CHECK(run_fp64_peak<1>(ctx, &b));
    FHE_CHECK(run_fp64_peak<2>(ctx, &c));
    if (dadd_tops) *dadd_tops = a;
    if (dmul_tops) *dmul_tops = b;
    if (dfma_tops) *dfma_tops = c;
    return FHE_OK;
}

extern "C" fhe_status fhe_diag_int32_peak(fhe_ctx* ctx, double* imad_tops, double* imad_hi_tops, double* imad_wide_tops) {
    if (!ctx) return FHE_EINVAL;
    double a = 0, b = 0, c = 0;
    FHE_CHECK(run_peak<0>(ctx, &a));
    FHE_CHECK(run_peak<1>(ctx, &b));
    FHE_CHECK(run_peak<2>(ctx, &c));
    if (imad_tops) *imad_tops = a;
    if (imad_hi_tops) *imad_hi_tops = b;
    if (imad_wide_tops) *imad_wide_tops = c;
    return FHE_OK;
}

// T butterflies/s of the u32 radix-16 register pass with the Shoup quotient on the multiply pipe (IMAD.HI) and on the FP64 pipe
// (DFMA rounded down); *same_out = 1 when both kernels produced identical words.
extern "C" fhe_status fhe_diag_butterfly_rate(fhe_ctx* ctx, double* imad_hi_tbf, double* dfma_tbf, int* same_out) {
    if (!ctx) return FHE_EINVAL;
    double a = 0, b = 0;
    uint32_t ca = 0, cb = 0;
    FHE_CHECK(run_bf_rate<0>(ctx, &a, &ca));
    FHE_CHECK(run_bf_rate<1>(ctx, &b, &cb));
    if (imad_hi_tbf) *imad_hi_tbf = a;
    if (dfma_tbf) *dfma_tbf = b;
    if (same_out) *same_out = ca == cb ? 1 : 0;
    return FHE_OK;
}
