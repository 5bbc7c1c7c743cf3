// RNS base conversion / rescale kernels and the CKKS ciphertext pipeline on sm_100a (K14-K16 of SURVEY.md §2).
//
// Reference call sites replaced:
//   RnsRq::extend_bases / switch_bases / rescale_k / round / div      util/src/ring/rns.rs:83-132, 287-345
//   RnsRq *= RnsRq (moduli intersection, per-limb negacyclic product)  util/src/ring/rns.rs:143-158
//   Ckks::mul / relinearize / key_switch / rotate / conjugate          scheme/ckks/src/ckks.rs:255-293
//   CkksCiphertext::rescale / automorphism                             scheme/ckks/src/ckks.rs:123-129
// Dataflow (differs from the reference, results identical because every step is exact modular arithmetic on canonical
// residues; the only floating-point step, the overflow estimate of extend_bases, sees the same integers in the same
// order): inputs are transformed once, the tensor product d0,d1,d2 and the key products are taken in the evaluation
// domain against a key-switching key that was transformed at upload, so one Ckks::mul costs 9l+3L transforms instead of
// the reference's 3*(4l + 2(l+L)).  Limb-batched transforms use the multi-modulus fast NTT (ntt_fast_launch.cu).
#include <algorithm>
#include <cstring>
#include <functional>
#include <vector>

#include "ctx.cuh"
#include "keygen_stream.cuh"
#include "rns_core.cuh"
#include "rns_tables.hpp"

namespace fhe {

// ---- host-built tables (values computed by rns_tables.hpp, shared with tests/hostsim) --------------------------------
// The tables travel BY VALUE as kernel parameters (RnsExtTabV / RescaleTabV, 7.5 / 9.3 KiB: CUDA 12 large kernel
// parameters), so they live in the constant bank; nothing is uploaded.
template <typename T>
static T* upload_vec(const std::vector<T>& h) {
    T* d = nullptr;
    if (h.empty()) return nullptr;
    if (cudaMalloc((void**)&d, h.size() * sizeof(T)) != cudaSuccess) return nullptr;
    if (cudaMemcpy(d, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice) != cudaSuccess) {
        cudaFree(d);
        return nullptr;
    }
    return d;
}
struct RnsExtOwned {
    RnsExtTabV tab;
    bool ok = false;
    void build(const std::vector<uint64_t>& qs, const std::vector<uint64_t>& ps) {
        ok = qs.size() <= (size_t)RNS_MAXL && ps.size() <= (size_t)RNS_MAXL;
        if (!ok) return;
        RnsExtHost h;
        h.build(qs, ps);
        memset(&tab, 0, sizeof tab);
        h.fill(tab);
    }
    void release() {}
};
struct RescaleOwned {
    RescaleTabV tab;
    bool ok = false;
    void build(const std::vector<uint64_t>& all, size_t k) {
        RescaleHost h;
        h.build(all, k);
        ok = h.kept.size() <= (size_t)RNS_MAXL && k <= (size_t)RNS_MAXL;
        if (!ok) return;
        memset(&tab, 0, sizeof tab);
        if (k > 1) {
            RnsExtHost e;
            e.build(h.dropped, h.kept);
            e.fill(tab.ext);
        }
        tab.l = (int)h.kept.size();
        tab.k = (int)k;
        for (size_t i = 0; i < all.size(); ++i) {
            tab.m_all[i] = h.m_all[i];
            tab.ph[i] = h.ph[i];
        }
        for (size_t i = 0; i < h.kept.size(); ++i) {
            tab.pinv[i] = h.pinv[i];
            tab.pinv_sh[i] = h.pinv_sh[i];
        }
    }
    void release() {}
};

// ---- kernels -----------------------------------------------------------------------------------------------------------
// in [B][in_limbs][n] (limbs in_off .. in_off+nq-1 are the source base) -> out [B][out_limbs][n] at limbs out_off..
__global__ void __launch_bounds__(256) rns_extend_kernel(const __grid_constant__ RnsExtTabV T, int log_n, unsigned long long batch, const uint64_t* __restrict__ in,
                                                         int in_limbs, int in_off, uint64_t* __restrict__ out, int out_limbs, int out_off) {
    const unsigned long long total = batch << log_n, stride = (unsigned long long)gridDim.x * blockDim.x;
    const size_t n = (size_t)1 << log_n;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const unsigned long long b = idx >> log_n;
        const size_t c = (size_t)(idx & (n - 1));
        const uint64_t* src = in + (b * in_limbs + in_off) * n + c;
        uint64_t x[RNS_MAXL];
#pragma unroll
        for (int i = 0; i < RNS_MAXL; ++i) x[i] = i < T.nq ? src[(size_t)i * n] : 0;
        uint64_t* dst = out + (b * out_limbs + out_off) * n + c;
        rns_extend_coeff(T, x, [&](int k, uint64_t y) { dst[(size_t)k * n] = y; });
    }
}
// copy limbs [0, nq) of every batch element into a wider layout (the "input limbs copied" half of extend_bases)
__global__ void __launch_bounds__(256) rns_copy_limbs_kernel(int log_n, unsigned long long batch, int nq, const uint64_t* __restrict__ in,
                                                             int in_limbs, uint64_t* __restrict__ out, int out_limbs) {
    const size_t n = (size_t)1 << log_n;
    const unsigned long long total = batch * nq * n, stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const unsigned long long b = idx / ((unsigned long long)nq * n);
        const unsigned long long r = idx - b * nq * n;
        out[b * out_limbs * n + r] = in[b * in_limbs * n + r];
    }
}
// rescale_k: in [B][l+k][n] -> out [B][l][n] (+ post [B][l][n]; if post_even_only only for even b)
__global__ void __launch_bounds__(256, 3) rns_rescale_kernel(const __grid_constant__ RescaleTabV R, int log_n, unsigned long long batch, const uint64_t* __restrict__ in,
                                                          const uint64_t* __restrict__ post, int post_even_only,
                                                          uint64_t* __restrict__ out) {
    const unsigned long long total = batch << log_n, stride = (unsigned long long)gridDim.x * blockDim.x;
    const size_t n = (size_t)1 << log_n;
    const int l = R.l, k = R.k;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const unsigned long long b = idx >> log_n;
        const size_t c = (size_t)(idx & (n - 1));
        const size_t ibase = b * (size_t)(l + k) * n + c, obase = b * (size_t)l * n + c;
        auto load = [&](int i) { return in[ibase + (size_t)i * n]; };
        rns_rescale_coeff(R, load, [&](int i, uint64_t v) {
            if (post && (!post_even_only || (b & 1ull) == 0)) v = R.m_all[i].add(v, post[obase + (size_t)i * n]);
            out[obase + (size_t)i * n] = v;
        });
    }
}
// rescale_k(k) then rescale_k(1) in one pass: in [B][l+k][n] -> out [B][l-1][n]   (Ckks::mul: key_switch's division by P followed by
// rescale, ckks.rs:266 + 123-125)
__global__ void __launch_bounds__(256, 3) rns_rescale2_kernel(const __grid_constant__ RescaleTabV R, const __grid_constant__ Rescale1TabV R1, int log_n,
                                                           unsigned long long batch, const uint64_t* __restrict__ in, uint64_t* __restrict__ out) {
    const unsigned long long total = batch << log_n, stride = (unsigned long long)gridDim.x * blockDim.x;
    const size_t n = (size_t)1 << log_n;
    const int l = R.l, k = R.k;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const unsigned long long b = idx >> log_n;
        const size_t c = (size_t)(idx & (n - 1));
        const size_t ibase = b * (size_t)(l + k) * n + c, obase = b * (size_t)(l - 1) * n + c;
        rns_rescale2_coeff(R, R1, [&](int i) { return in[ibase + (size_t)i * n]; }, [&](int i, uint64_t v) { out[obase + (size_t)i * n] = v; });
    }
}

// Row-wise elementwise kernels: blockIdx.y walks the rows (one limb of one polynomial: the modulus and every base offset are uniform,
// so no per-element division), blockIdx.x / threadIdx.x walk the row in coefficient pairs (16-byte accesses).
static constexpr int CKKS_ROW_THREADS = 256;
struct U64x2 {
    uint64_t x, y;
};
DEV U64x2 ld2(const uint64_t* p) {
    const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
    return U64x2{v.x, v.y};
}
DEV void st2(uint64_t* p, uint64_t x, uint64_t y) { *reinterpret_cast<ulonglong2*>(p) = make_ulonglong2(x, y); }
static dim3 row_grid(fhe_ctx* ctx, unsigned log_n, unsigned long long rows) {
    const unsigned long long pairs = ((unsigned long long)1 << log_n) / 2;
    const unsigned gx = (unsigned)std::max<unsigned long long>(1, pairs / (CKKS_ROW_THREADS * 4));  // ~4 pairs per thread and row
    const unsigned long long want = std::max<unsigned long long>(1, (unsigned long long)ctx->sm_count * 16 / gx);
    return dim3(gx, (unsigned)std::min<unsigned long long>(std::min<unsigned long long>(rows, want), 65535), 1);
}

// tensor product in the evaluation domain: e0, e1 [C][2 (b,a)][l][n] -> d01 [C][2 (d0,d1)][l][n], d2 [C][l][n]  (ckks.rs:262-266)
__global__ void __launch_bounds__(CKKS_ROW_THREADS) ckks_tensor_kernel(const Mod64* __restrict__ mods, int l, int log_n, unsigned long long count,
                                                                       const uint64_t* __restrict__ e0, const uint64_t* __restrict__ e1,
                                                                       uint64_t* __restrict__ d01, uint64_t* __restrict__ d2) {
    const size_t n = (size_t)1 << log_n, ln = (size_t)l * n;
    const uint32_t rows = (uint32_t)(count * l);
    for (uint32_t row = blockIdx.y; row < rows; row += gridDim.y) {
        const uint32_t c = row / (uint32_t)l, j = row - c * (uint32_t)l;
        const Mod64 m = mods[j];
        const size_t o = (size_t)c * 2 * ln + (size_t)j * n;
        const uint64_t *pb0 = e0 + o, *pa0 = pb0 + ln, *pb1 = e1 + o, *pa1 = pb1 + ln;
        uint64_t *q0 = d01 + o, *q1 = q0 + ln, *q2 = d2 + (size_t)c * ln + (size_t)j * n;
        for (size_t x = 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); x < n; x += 2 * (size_t)gridDim.x * blockDim.x) {
            const U64x2 b0 = ld2(pb0 + x), a0 = ld2(pa0 + x), b1 = ld2(pb1 + x), a1 = ld2(pa1 + x);
            st2(q0 + x, m.mul(b0.x, b1.x), m.mul(b0.y, b1.y));
            st2(q1 + x, m.add(m.mul(b0.x, a1.x), m.mul(a0.x, b1.x)), m.add(m.mul(b0.y, a1.y), m.mul(a0.y, b1.y)));
            st2(q2 + x, m.mul(a0.x, a1.x), m.mul(a0.y, a1.y));
        }
    }
}
// key products: x = [xq (l limbs, eval) ; xp (L limbs, eval)] against ksk [2 (b,a)][2L][n] eval -> out [C][2][l+L][n]  (ckks.rs:289-291)
__global__ void __launch_bounds__(CKKS_ROW_THREADS) ckks_keymul_kernel(const Mod64* __restrict__ mods /* [2L]: qs then ps */, int l, int big_l,
                                                                       int log_n, unsigned long long count, const uint64_t* __restrict__ xq,
                                                                       const uint64_t* __restrict__ xp, const uint64_t* __restrict__ ksk,
                                                                       const uint64_t* __restrict__ addend /* [C][2][l][n] or null */,
                                                                       const uint64_t* __restrict__ pmod /* [L]: P mod q_j */,
                                                                       uint64_t* __restrict__ out) {
    const size_t n = (size_t)1 << log_n;
    const uint32_t le = (uint32_t)(l + big_l), rows = (uint32_t)(count * le);
    for (uint32_t row = blockIdx.y; row < rows; row += gridDim.y) {
        const uint32_t c = row / le, j = row - c * le;
        const uint32_t kl = j < (uint32_t)l ? j : (uint32_t)big_l + (j - (uint32_t)l);
        const Mod64 m = mods[kl];
        const uint64_t* pv = j < (uint32_t)l ? xq + ((size_t)c * l + j) * n : xp + ((size_t)c * big_l + (j - l)) * n;
        const uint64_t *pkb = ksk + (size_t)kl * n, *pka = ksk + ((size_t)2 * big_l + kl) * n;
        uint64_t *ob = out + ((size_t)c * 2 * le + j) * n, *oa = ob + (size_t)le * n;
        // Ckks::mul adds (d0, d1) to the relinearised pair after the division by P (ckks.rs:266, rns.rs:127-132).  On the kept
        // limbs rescale_k is (x_i + P/2 - ext_i) * P^-1, so adding P * d_i to x_i here - in the evaluation domain, before the
        // inverse transform - yields exactly d_i + rescale_k(x)_i and saves the inverse transforms of d0 and d1.
        const bool add = addend && j < (uint32_t)l;
        const uint64_t* pd0 = add ? addend + ((size_t)c * 2 * l + j) * n : nullptr;
        const uint64_t* pd1 = add ? pd0 + (size_t)l * n : nullptr;
        const uint64_t pm = add ? pmod[j] : 0;
        for (size_t x = 2 * ((size_t)blockIdx.x * blockDim.x + threadIdx.x); x < n; x += 2 * (size_t)gridDim.x * blockDim.x) {
            const U64x2 v = ld2(pv + x), wb = ld2(pkb + x), wa = ld2(pka + x);
            uint64_t b0 = m.mul(wb.x, v.x), b1 = m.mul(wb.y, v.y), a0 = m.mul(wa.x, v.x), a1 = m.mul(wa.y, v.y);
            if (add) {
                const U64x2 d0 = ld2(pd0 + x), d1 = ld2(pd1 + x);
                b0 = m.add(b0, m.mul(pm, d0.x));
                b1 = m.add(b1, m.mul(pm, d0.y));
                a0 = m.add(a0, m.mul(pm, d1.x));
                a1 = m.add(a1, m.mul(pm, d1.y));
            }
            st2(ob + x, b0, b1);
            st2(oa + x, a0, a1);
        }
    }
}
// plaintext x ciphertext in the evaluation domain: e [C][2][l][n] *= pe [1 or C][l][n] (limb-wise)   (ckks.rs:250-253)
__global__ void __launch_bounds__(256) ckks_ptmul_kernel(const Mod64* __restrict__ mods, int l, int log_n, unsigned long long count, int pt_per_ct,
                                                         const uint64_t* __restrict__ pe, const uint64_t* src /* may be e */, uint64_t* e) {
    const size_t n = (size_t)1 << log_n, ln = (size_t)l * n;
    const unsigned long long total = count * 2 * ln, stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const unsigned long long c = idx / (2 * ln);
        const size_t r = (size_t)(idx % ln);
        const Mod64 m = mods[r >> log_n];
        e[idx] = m.mul(src[idx], pe[(pt_per_ct ? c * ln : 0) + r]);
    }
}
// limb-wise modular addition of RNS polynomials: out = a + b over [polys][n] with modulus mods[poly % l]   (rns.rs Add impls)
__global__ void __launch_bounds__(256) rns_add_kernel(const Mod64* __restrict__ mods, int l, int log_n, unsigned long long polys,
                                                      const uint64_t* __restrict__ a, const uint64_t* __restrict__ b, uint64_t* __restrict__ out) {
    const unsigned long long total = polys << log_n, stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride)
        out[idx] = mods[(idx >> log_n) % l].add(a[idx], b[idx]);
}
// RnsRq::automorphism (rns.rs:74-77 -> avec.rs:34-50) on [polys][n] with modulus mods[poly % l]
__global__ void __launch_bounds__(256) rns_automorphism_kernel(const Mod64* __restrict__ mods, int l, int log_n, unsigned long long polys, uint32_t t,
                                                               const uint64_t* __restrict__ in, uint64_t* __restrict__ out) {
    const uint32_t n = 1u << log_n;
    const unsigned long long total = polys << log_n, stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const unsigned long long p = idx >> log_n;
        const uint32_t i = (uint32_t)(idx & (n - 1));
        const uint32_t it = (uint32_t)(((unsigned long long)i * t) & (2ull * n - 1));
        uint64_t v = in[idx];
        if (it >= n) v = mods[p % l].neg(v);
        out[(p << log_n) + (it & (n - 1))] = v;
    }
}

// ---- key generation on the device (ckks.rs:154-162 ksk_gen, 215-225 sk_encrypt) --------------------------------
// key buffer [2 (b, a)][2L][n], coefficient form: a_i[c] uniform mod (q|p)_i from the counter stream; b_i[c] = e[c] + P sk'[c] mod
// (q|p)_i (P = prod ps; the same small e for every limb, RnsRq::sample_i64); the - a * sk term is subtracted afterwards
__global__ void __launch_bounds__(256) ckks_kg_rows_kernel(uint64_t seed, uint32_t da, uint32_t de, const Mod64* __restrict__ mods,
                                                           const uint64_t* __restrict__ pmod_all /* [2L]: P mod modulus */, int limbs, int log_n,
                                                           const int64_t* __restrict__ skp, uint64_t* __restrict__ key) {
    const size_t n = (size_t)1 << log_n;
    const unsigned long long total = (unsigned long long)limbs << log_n, stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const int i = (int)(idx >> log_n);
        const size_t c = (size_t)(idx & (n - 1));
        const Mod64 m = mods[i];
        auto res = [&](int64_t v) { return v < 0 ? m.q - (uint64_t)(-v) % m.q : (uint64_t)v % m.q; };
        uint64_t sp = res(skp[c]);
        if (sp == m.q) sp = 0;
        uint64_t e = res(ks_gauss(seed, de, c));
        if (e == m.q) e = 0;
        key[idx] = m.add(e, m.mul(pmod_all[i], sp));
        key[total + idx] = ks_uniform(seed, da, idx, m.q);
    }
}
// dst[i][c] = residue of the small integer v[c] modulo limb i
__global__ void __launch_bounds__(256) ckks_kg_residues_kernel(const Mod64* __restrict__ mods, int limbs, int log_n, const int64_t* __restrict__ v,
                                                               uint64_t* __restrict__ dst) {
    const size_t n = (size_t)1 << log_n;
    const unsigned long long total = (unsigned long long)limbs << log_n, stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
        const Mod64 m = mods[idx >> log_n];
        const int64_t x = v[idx & (n - 1)];
        const uint64_t r = x < 0 ? m.q - (uint64_t)(-x) % m.q : (uint64_t)x % m.q;
        dst[idx] = r == m.q ? 0 : r;
    }
}
// t <- t o s (evaluation domain, per-limb modulus); then, after the inverse transform, b <- b - t
__global__ void __launch_bounds__(256) ckks_kg_mul_kernel(const Mod64* __restrict__ mods, int log_n, unsigned long long total, uint64_t* __restrict__ t,
                                                          const uint64_t* __restrict__ s) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) t[idx] = mods[idx >> log_n].mul(t[idx], s[idx]);
}
__global__ void __launch_bounds__(256) ckks_kg_sub_kernel(const Mod64* __restrict__ mods, int log_n, unsigned long long total, uint64_t* __restrict__ b,
                                                          const uint64_t* __restrict__ t) {
    const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
    for (unsigned long long idx = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) b[idx] = mods[idx >> log_n].sub(b[idx], t[idx]);
}

static unsigned stream_grid(fhe_ctx* ctx, unsigned long long work) {
    unsigned long long blocks = (work + 255) / 256, cap = (unsigned long long)ctx->sm_count * 16;
    return (unsigned)std::max<unsigned long long>(1, std::min(blocks, cap));
}

static fhe_status check_moduli(fhe_ctx* ctx, const std::vector<uint64_t>& v) {
    for (size_t i = 0; i < v.size(); ++i) {
        FHE_REQUIRE(ctx, v[i] > 2 && (v[i] & 1) && v[i] < (1ull << 62), "RNS moduli must be odd and < 2^62");
        for (size_t j = 0; j < i; ++j) FHE_REQUIRE(ctx, v[i] != v[j], "RNS moduli must be pairwise distinct (rns.rs:84)");
    }
    return FHE_OK;
}

// cached table lookup: key = kind, qs..., 0, ps...
template <typename Owned, typename Build>
static fhe_status cached_table(fhe_ctx* ctx, std::vector<uint64_t> key, Build build, const Owned** out) {
    std::lock_guard<std::mutex> lock(ctx->mu);
    auto it = ctx->rns_tabs.find(key);
    if (it == ctx->rns_tabs.end()) {
        Owned* o = new Owned();
        build(*o);
        if (!o->ok) {
            o->release();
            delete o;
            return fail(ctx, FHE_EINVAL, "RNS table: at most %d source / target limbs per conversion", RNS_MAXL);
        }
        ctx->cleanup.push_back([o]() {
            o->release();
            delete o;
        });
        it = ctx->rns_tabs.emplace(key, (void*)o).first;
    }
    *out = (const Owned*)it->second;
    return FHE_OK;
}
static fhe_status get_ext(fhe_ctx* ctx, const std::vector<uint64_t>& qs, const std::vector<uint64_t>& ps, const RnsExtOwned** out) {
    std::vector<uint64_t> key{1};
    key.insert(key.end(), qs.begin(), qs.end());
    key.push_back(0);
    key.insert(key.end(), ps.begin(), ps.end());
    return cached_table<RnsExtOwned>(ctx, key, [&](RnsExtOwned& o) { o.build(qs, ps); }, out);
}
static fhe_status get_rescale(fhe_ctx* ctx, const std::vector<uint64_t>& all, size_t k, const RescaleOwned** out) {
    std::vector<uint64_t> key{2, (uint64_t)k};
    key.insert(key.end(), all.begin(), all.end());
    return cached_table<RescaleOwned>(ctx, key, [&](RescaleOwned& o) { o.build(all, k); }, out);
}

static fhe_status run_extend(fhe_ctx* ctx, const std::vector<uint64_t>& qs, const std::vector<uint64_t>& ps, unsigned log_n, size_t batch,
                             const uint64_t* d_in, int in_limbs, int in_off, uint64_t* d_out, int out_limbs, int out_off) {
    FHE_REQUIRE(ctx, qs.size() >= 1 && qs.size() <= (size_t)RNS_MAXL, "base conversion supports 1..%d source limbs", RNS_MAXL);
    const RnsExtOwned* t;
    FHE_CHECK(get_ext(ctx, qs, ps, &t));
    rns_extend_kernel<<<stream_grid(ctx, (unsigned long long)batch << log_n), 256, 0, ctx->stream>>>(t->tab, (int)log_n, batch, d_in, in_limbs,
                                                                                                    in_off, d_out, out_limbs, out_off);
    return after_launch(ctx, "rns_extend_kernel");
}
static fhe_status run_rescale(fhe_ctx* ctx, const std::vector<uint64_t>& all, size_t k, unsigned log_n, size_t batch, const uint64_t* d_in,
                              const uint64_t* d_post, bool post_even_only, uint64_t* d_out) {
    FHE_REQUIRE(ctx, k >= 1 && k < all.size(), "rescale_k needs 0 < k < number of limbs (rns.rs:104)");
    FHE_REQUIRE(ctx, k <= (size_t)RNS_MAXL, "rescale_k supports dropping at most %d limbs", RNS_MAXL);
    const RescaleOwned* t;
    FHE_CHECK(get_rescale(ctx, all, k, &t));
    rns_rescale_kernel<<<stream_grid(ctx, (unsigned long long)batch << log_n), 256, 0, ctx->stream>>>(t->tab, (int)log_n, batch, d_in, d_post,
                                                                                                     post_even_only ? 1 : 0, d_out);
    return after_launch(ctx, "rns_rescale_kernel");
}

// rescale_k(all, k) followed by rescale_k(kept, 1), fused (kept = all[0 .. all.size() - k), at least two limbs)
static fhe_status run_rescale2(fhe_ctx* ctx, const std::vector<uint64_t>& all, size_t k, unsigned log_n, size_t batch, const uint64_t* d_in,
                               uint64_t* d_out) {
    FHE_REQUIRE(ctx, k >= 2 && k + 2 <= all.size() && k <= (size_t)RNS_MAXL, "fused rescale needs k >= 2 dropped and >= 2 kept limbs");
    const std::vector<uint64_t> kept(all.begin(), all.end() - k);
    const RescaleOwned *t, *t1;
    FHE_CHECK(get_rescale(ctx, all, k, &t));
    FHE_CHECK(get_rescale(ctx, kept, 1, &t1));
    Rescale1TabV r1;
    memset(&r1, 0, sizeof r1);
    r1.l = t1->tab.l;
    r1.fast = 1;
    for (size_t i = 0; i < kept.size(); ++i)
        if (all[i] >= (1ull << 61) || kept.back() > 2 * kept[i]) r1.fast = 0;
    for (size_t i = 0; i < kept.size(); ++i) {
        r1.m_all[i] = t1->tab.m_all[i];
        r1.ph[i] = t1->tab.ph[i];
    }
    for (size_t i = 0; i + 1 < kept.size(); ++i) {
        r1.pinv[i] = t1->tab.pinv[i];
        r1.pinv_sh[i] = t1->tab.pinv_sh[i];
    }
    rns_rescale2_kernel<<<stream_grid(ctx, (unsigned long long)batch << log_n), 256, 0, ctx->stream>>>(t->tab, r1, (int)log_n, batch, d_in, d_out);
    return after_launch(ctx, "rns_rescale2_kernel");
}

}  // namespace fhe

using namespace fhe;

struct fhe_ckks_ctx {
    void* ws2 = nullptr;  // second grow-only workspace (fhe_ckks_mul_mat, which calls entry points that use `ws`)
    size_t ws2_bytes = 0;
    unsigned log_n = 0;
    size_t big_l = 0;
    std::vector<uint64_t> qs, ps;
    Mod64* d_mods = nullptr;  // [2L]: qs then ps
    uint64_t* d_pmod = nullptr;  // [L]: (product of the special primes) mod q_j
    // grow-only workspace
    void* ws = nullptr;
    size_t ws_bytes = 0;
};
struct fhe_ckks_ksk {
    uint64_t* d_eval = nullptr;  // [2 (b,a)][2L][n] evaluation form
    size_t bytes = 0;
};

namespace fhe {
static fhe_status ckks_ws(fhe_ctx* ctx, fhe_ckks_ctx* ck, size_t bytes, uint64_t** out) {
    if (ck->ws_bytes < bytes) {
        FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ck->ws) cudaFree(ck->ws);
        ck->ws = nullptr;
        ck->ws_bytes = 0;
        if (cudaMalloc(&ck->ws, bytes) != cudaSuccess) return fail(ctx, FHE_ENOMEM, "CKKS workspace of %zu bytes", bytes);
        ck->ws_bytes = bytes;
    }
    *out = (uint64_t*)ck->ws;
    return FHE_OK;
}
// bytes of workspace one chunk of a batched operation may use (FHE_B200_CKKS_WS_MB overrides; tuning knob)
static size_t ckks_ws_budget() {
    static const size_t v = [] {
        const char* e = getenv("FHE_B200_CKKS_WS_MB");
        const long mb = e ? atol(e) : 0;
        return mb > 0 ? (size_t)mb << 20 : (size_t)6 << 30;
    }();
    return v;
}
static std::vector<uint64_t> level_qs(const fhe_ckks_ctx* ck, size_t l) { return std::vector<uint64_t>(ck->qs.begin(), ck->qs.begin() + l); }
static std::vector<uint64_t> level_qps(const fhe_ckks_ctx* ck, size_t l) {
    std::vector<uint64_t> v = level_qs(ck, l);
    v.insert(v.end(), ck->ps.begin(), ck->ps.end());
    return v;
}

// Ckks::key_switch core (ckks.rs:284-293) on `count` polynomials a (coefficient form [count][l][n]) whose evaluation form
// a_eval [count][l][n] is already available: r [count][2][l][n] = rescale_k(ksk * extend(a), L) (+ post on the b half).
// scratch: xp [count][L][n], kk [count][2][l+L][n]
// d01_eval (optional, [count][2][l][n] evaluation form): added to the result (see ckks_keymul_kernel)
static fhe_status key_switch_core(fhe_ctx* ctx, fhe_ckks_ctx* ck, const fhe_ckks_ksk* ksk, size_t l, size_t count, const uint64_t* a_coeff,
                                  const uint64_t* a_eval, uint64_t* xp, uint64_t* kk, const uint64_t* post, uint64_t* r,
                                  const uint64_t* d01_eval = nullptr, uint64_t* rescaled_out = nullptr) {
    const unsigned log_n = ck->log_n;
    const size_t L = ck->big_l;
    const std::vector<uint64_t> qs = level_qs(ck, l), qps = level_qps(ck, l);
    FHE_CHECK(run_extend(ctx, qs, ck->ps, log_n, count, a_coeff, (int)l, 0, xp, (int)L, 0));
    FHE_CHECK(launch_ntt_rns_u64(ctx, ck->ps.data(), L, log_n, count * L, xp, true));
    ckks_keymul_kernel<<<row_grid(ctx, log_n, (unsigned long long)count * (l + L)), CKKS_ROW_THREADS, 0, ctx->stream>>>(
        ck->d_mods, (int)l, (int)L, (int)log_n, count, a_eval, xp, ksk->d_eval, d01_eval, ck->d_pmod, kk);
    FHE_CHECK(after_launch(ctx, "ckks_keymul_kernel"));
    FHE_CHECK(launch_ntt_rns_u64(ctx, qps.data(), l + L, log_n, count * 2 * (l + L), kk, false));
    // rescaled_out: the caller wants rescale(r) (one more limb dropped) and not r itself
    if (rescaled_out && !post && L >= 2 && l >= 2) return run_rescale2(ctx, qps, L, log_n, count * 2, kk, rescaled_out);
    FHE_CHECK(run_rescale(ctx, qps, L, log_n, count * 2, kk, post, true, r));
    if (rescaled_out) return run_rescale(ctx, qs, 1, log_n, count * 2, r, nullptr, false, rescaled_out);
    return FHE_OK;
}
}  // namespace fhe

extern "C" {

// ---- util-level RNS entry points ---------------------------------------------------------------------------------------
fhe_status fhe_rns_extend_bases(fhe_ctx* ctx, const uint64_t* qs, size_t nq, const uint64_t* ps, size_t np, unsigned log_n, size_t batch,
                                const uint64_t* d_in, uint64_t* d_out) {
    if (!ctx || !qs || !ps) return FHE_EINVAL;
    if (batch == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_in && d_out && log_n <= 20, "bad arguments");
    std::vector<uint64_t> vq(qs, qs + nq), vp(ps, ps + np), all(vq);
    all.insert(all.end(), vp.begin(), vp.end());
    FHE_CHECK(check_moduli(ctx, all));
    rns_copy_limbs_kernel<<<stream_grid(ctx, (unsigned long long)batch * nq << log_n), 256, 0, ctx->stream>>>((int)log_n, batch, (int)nq, d_in,
                                                                                                             (int)nq, d_out, (int)(nq + np));
    FHE_CHECK(after_launch(ctx, "rns_copy_limbs_kernel"));
    return run_extend(ctx, vq, vp, log_n, batch, d_in, (int)nq, 0, d_out, (int)(nq + np), (int)nq);
}
fhe_status fhe_rns_rescale_k(fhe_ctx* ctx, const uint64_t* qs, size_t nq, size_t k, unsigned log_n, size_t batch, const uint64_t* d_in,
                             uint64_t* d_out) {
    if (!ctx || !qs) return FHE_EINVAL;
    if (batch == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_in && d_out && log_n <= 20, "bad arguments");
    std::vector<uint64_t> all(qs, qs + nq);
    FHE_CHECK(check_moduli(ctx, all));
    return run_rescale(ctx, all, k, log_n, batch, d_in, nullptr, false, d_out);
}

// ---- CKKS ----------------------------------------------------------------------------------------------------------------
fhe_status fhe_ckks_create(fhe_ctx* ctx, unsigned log_n, const uint64_t* qs, const uint64_t* ps, size_t big_l, fhe_ckks_ctx** out) {
    if (!ctx || !qs || !ps || !out) return FHE_EINVAL;
    *out = nullptr;
    FHE_REQUIRE(ctx, log_n >= 1 && log_n <= 17, "CKKS ring degree 2^%u out of range", log_n);
    FHE_REQUIRE(ctx, big_l >= 1 && big_l <= (size_t)RNS_MAXL, "CKKS supports 1..%d ciphertext primes", RNS_MAXL);
    std::vector<uint64_t> all(qs, qs + big_l);
    all.insert(all.end(), ps, ps + big_l);
    FHE_CHECK(check_moduli(ctx, all));
    for (uint64_t q : all) {  // every modulus must support the degree-2^log_n negacyclic NTT (panics in the reference otherwise)
        const NttTable* t;
        FHE_CHECK(get_ntt_table(ctx, q, 64, (size_t)1 << log_n, &t));
    }
    fhe_ckks_ctx* ck = new fhe_ckks_ctx();
    ck->log_n = log_n;
    ck->big_l = big_l;
    ck->qs.assign(qs, qs + big_l);
    ck->ps.assign(ps, ps + big_l);
    std::vector<Mod64> m(2 * big_l);
    for (size_t i = 0; i < 2 * big_l; ++i) m[i] = make_mod<Mod64>(all[i]);
    ck->d_mods = upload_vec(m);
    std::vector<uint64_t> pmod(big_l);
    for (size_t j = 0; j < big_l; ++j) {
        uint64_t v = 1 % qs[j];
        for (size_t i = 0; i < big_l; ++i) v = host_mulmod(v, ps[i] % qs[j], qs[j]);
        pmod[j] = v;
    }
    ck->d_pmod = upload_vec(pmod);
    if (!ck->d_mods || !ck->d_pmod) {
        if (ck->d_mods) cudaFree(ck->d_mods);
        if (ck->d_pmod) cudaFree(ck->d_pmod);
        delete ck;
        return fail(ctx, FHE_ENOMEM, "CKKS modulus table upload failed");
    }
    *out = ck;
    return FHE_OK;
}
void fhe_ckks_destroy(fhe_ctx* ctx, fhe_ckks_ctx* ck) {
    if (!ck) return;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    if (ck->d_mods) cudaFree(ck->d_mods);
    if (ck->d_pmod) cudaFree(ck->d_pmod);
    if (ck->ws) cudaFree(ck->ws);
    if (ck->ws2) cudaFree(ck->ws2);
    delete ck;
}
fhe_status fhe_ckks_ksk_upload(fhe_ctx* ctx, fhe_ckks_ctx* ck, const uint64_t* ksk, fhe_ckks_ksk** out) {
    if (!ctx || !ck || !ksk || !out) return FHE_EINVAL;
    *out = nullptr;
    const size_t n = (size_t)1 << ck->log_n, L2 = 2 * ck->big_l, words = 2 * L2 * n;
    std::vector<uint64_t> qps = level_qps(ck, ck->big_l);
    for (size_t h = 0; h < 2; ++h)
        for (size_t i = 0; i < L2; ++i)
            for (size_t c = 0; c < n; ++c) FHE_REQUIRE(ctx, ksk[(h * L2 + i) * n + c] < qps[i], "key coefficient out of range");
    fhe_ckks_ksk* k = new fhe_ckks_ksk();
    if (cudaMalloc((void**)&k->d_eval, words * 8) != cudaSuccess) {
        delete k;
        return fail(ctx, FHE_ENOMEM, "key alloc");
    }
    cudaError_t e = cudaMemcpyAsync(k->d_eval, ksk, words * 8, cudaMemcpyHostToDevice, ctx->stream);
    fhe_status st = e == cudaSuccess ? FHE_OK : fail(ctx, FHE_ECUDA, "key upload: %s", cudaGetErrorString(e));
    if (st == FHE_OK) st = launch_ntt_rns_u64(ctx, qps.data(), L2, ck->log_n, 2 * L2, k->d_eval, true);
    if (st == FHE_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = fail(ctx, FHE_ECUDA, "key transform failed");
    if (st != FHE_OK) {
        cudaFree(k->d_eval);
        delete k;
        return st;
    }
    k->bytes = words * 8;
    *out = k;
    return FHE_OK;
}
size_t fhe_ckks_ksk_bytes(const fhe_ckks_ksk* ksk) { return ksk ? ksk->bytes : 0; }
fhe_status fhe_ckks_ksk_broadcast(fhe_ctx* ctx, fhe_ckks_ksk* ksk, void* nccl_comm, int root) {
    if (!ctx || !ksk) return FHE_EINVAL;
    FHE_CHECK(fhe_keys_broadcast(ctx, nccl_comm, root, ksk->d_eval, ksk->bytes));
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    return FHE_OK;
}
void fhe_ckks_ksk_free(fhe_ctx* ctx, fhe_ckks_ksk* ksk) {
    if (!ksk) return;
    if (ctx) cudaStreamSynchronize(ctx->stream);
    if (ksk->d_eval) cudaFree(ksk->d_eval);
    delete ksk;
}

// Key generation on the device: Ckks::sk_gen (ckks.rs:139-141, zo(0.5)), rlk_gen (164-167: ksk_gen(sk, sk^2)) and
// one automorphism key per exponent in auto_ts (cjk_gen / rtk_gen, 169-184: ksk_gen(sk, sk(X^t))), every key generated on the GPU
// from the counter-based stream of csrc/keygen_stream.cuh and left there in evaluation form.  keys_out receives 1 + n_auto
// handles (rlk first); sk_out [N] the secret; export_out (optional, HOST, [1 + n_auto][2 (b, a)][2L][N]) the coefficient-form keys
// in the layout of fhe_ckks_ksk_upload for parity checks.
fhe_status fhe_ckks_keygen(fhe_ctx* ctx, fhe_ckks_ctx* ck, uint64_t seed, size_t n_auto, const int64_t* auto_ts, int64_t* sk_out,
                           fhe_ckks_ksk** keys_out, uint64_t* export_out) {
    if (!ctx || !ck || !keys_out || (n_auto && !auto_ts)) return FHE_EINVAL;
    const unsigned log_n = ck->log_n;
    const size_t n = (size_t)1 << log_n, L2 = 2 * ck->big_l, plane = L2 * n;
    for (size_t i = 0; i <= n_auto; ++i) keys_out[i] = nullptr;
    std::vector<int64_t> sk(n);
    for (size_t c = 0; c < n; ++c) sk[c] = ks_ternary(seed, KS_CKKS_SK, c);
    const std::vector<uint64_t> qps = level_qps(ck, ck->big_l);
    std::vector<uint64_t> pmod_all(L2, 0);
    for (size_t i = 0; i < ck->big_l; ++i) {
        uint64_t v = 1 % qps[i];
        for (uint64_t p : ck->ps) v = host_mulmod(v, p % qps[i], qps[i]);
        pmod_all[i] = v;  // P mod p_j = 0 for the special primes themselves
    }
    int64_t *d_sk = nullptr, *d_skp = nullptr;
    uint64_t *d_sk_eval = nullptr, *d_tmp = nullptr, *d_pmod = nullptr;
    fhe_status st = FHE_OK;
    auto cu = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && st == FHE_OK) st = fail(ctx, FHE_ECUDA, "ckks keygen %s: %s", what, cudaGetErrorString(e));
    };
    cu(cudaMalloc(&d_sk, n * 8), "alloc");
    cu(cudaMalloc(&d_skp, n * 8), "alloc");
    cu(cudaMalloc(&d_sk_eval, plane * 8), "alloc");
    cu(cudaMalloc(&d_tmp, plane * 8), "alloc");
    cu(cudaMalloc(&d_pmod, L2 * 8), "alloc");
    if (st == FHE_OK) {
        cu(cudaMemcpyAsync(d_sk, sk.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream), "copy");
        cu(cudaMemcpyAsync(d_pmod, pmod_all.data(), L2 * 8, cudaMemcpyHostToDevice, ctx->stream), "copy");
    }
    const unsigned grid = stream_grid(ctx, plane);
    if (st == FHE_OK) {
        ckks_kg_residues_kernel<<<grid, 256, 0, ctx->stream>>>(ck->d_mods, (int)L2, (int)log_n, d_sk, d_sk_eval);
        st = after_launch(ctx, "ckks_kg_residues_kernel");
    }
    if (st == FHE_OK) st = launch_ntt_rns_u64(ctx, qps.data(), L2, log_n, L2, d_sk_eval, true);
    // sk^2 as small integers: exact through the first prime (|coefficient| <= N << q_0 / 2), the reference's i64 Karatsuba gives the same
    std::vector<int64_t> sk_sq(n);
    if (st == FHE_OK) {
        const uint64_t q0 = ck->qs[0];
        cu(cudaMemcpyAsync(d_tmp, d_sk_eval, n * 8, cudaMemcpyDeviceToDevice, ctx->stream), "copy");  // limb 0 of sk in evaluation form
        if (st == FHE_OK) st = fhe_pointwise_mul_u64(ctx, q0, n, d_tmp, d_tmp, d_tmp);
        if (st == FHE_OK) st = launch_ntt_u64(ctx, q0, log_n, 1, d_tmp, false);
        std::vector<uint64_t> h(n);
        cu(cudaMemcpyAsync(h.data(), d_tmp, n * 8, cudaMemcpyDeviceToHost, ctx->stream), "copy");
        cu(cudaStreamSynchronize(ctx->stream), "sync");
        for (size_t c = 0; c < n; ++c) sk_sq[c] = h[c] < (q0 >> 1) ? (int64_t)h[c] : (int64_t)h[c] - (int64_t)q0;
    }
    for (size_t key = 0; key <= n_auto && st == FHE_OK; ++key) {
        std::vector<int64_t> skp = sk_sq;
        if (key > 0) {  // sk(X^t) on the i64 secret (avec.rs:34-50)
            const int64_t m2 = 2 * (int64_t)n, t = ((auto_ts[key - 1] % m2) + m2) % m2;
            for (size_t i = 0; i < n; ++i) {
                const size_t it = (size_t)(((unsigned long long)i * (unsigned long long)t) % (unsigned long long)m2);
                if (it < n)
                    skp[it] = sk[i];
                else
                    skp[it - n] = -sk[i];
            }
        }
        fhe_ckks_ksk* k = new fhe_ckks_ksk();
        k->bytes = 2 * plane * 8;
        if (cudaMalloc((void**)&k->d_eval, k->bytes) != cudaSuccess) {
            delete k;
            st = fail(ctx, FHE_ENOMEM, "key alloc");
            break;
        }
        keys_out[key] = k;
        cu(cudaMemcpyAsync(d_skp, skp.data(), n * 8, cudaMemcpyHostToDevice, ctx->stream), "copy");
        ckks_kg_rows_kernel<<<grid, 256, 0, ctx->stream>>>(seed, ks_ckks_a((uint32_t)key), ks_ckks_e((uint32_t)key), ck->d_mods, d_pmod, (int)L2, (int)log_n,
                                                           d_skp, k->d_eval);
        if (st == FHE_OK) st = after_launch(ctx, "ckks_kg_rows_kernel");
        cu(cudaMemcpyAsync(d_tmp, k->d_eval + plane, plane * 8, cudaMemcpyDeviceToDevice, ctx->stream), "copy");
        if (st == FHE_OK) st = launch_ntt_rns_u64(ctx, qps.data(), L2, log_n, L2, d_tmp, true);
        ckks_kg_mul_kernel<<<grid, 256, 0, ctx->stream>>>(ck->d_mods, (int)log_n, plane, d_tmp, d_sk_eval);
        if (st == FHE_OK) st = after_launch(ctx, "ckks_kg_mul_kernel");
        if (st == FHE_OK) st = launch_ntt_rns_u64(ctx, qps.data(), L2, log_n, L2, d_tmp, false);
        ckks_kg_sub_kernel<<<grid, 256, 0, ctx->stream>>>(ck->d_mods, (int)log_n, plane, k->d_eval, d_tmp);  // b = -(a sk) + e + P sk'
        if (st == FHE_OK) st = after_launch(ctx, "ckks_kg_sub_kernel");
        if (export_out) {
            cu(cudaMemcpyAsync(export_out + key * 2 * plane, k->d_eval, 2 * plane * 8, cudaMemcpyDeviceToHost, ctx->stream), "export");
            cu(cudaStreamSynchronize(ctx->stream), "sync");
        }
        if (st == FHE_OK) st = launch_ntt_rns_u64(ctx, qps.data(), L2, log_n, 2 * L2, k->d_eval, true);
        cu(cudaStreamSynchronize(ctx->stream), "sync");  // skp (host vector) is re-used by the next key
    }
    cudaStreamSynchronize(ctx->stream);
    for (void* p : {(void*)d_sk, (void*)d_skp, (void*)d_sk_eval, (void*)d_tmp, (void*)d_pmod})
        if (p) cudaFree(p);
    if (st != FHE_OK) {
        for (size_t i = 0; i <= n_auto; ++i) {
            fhe_ckks_ksk_free(ctx, keys_out[i]);
            keys_out[i] = nullptr;
        }
        return st;
    }
    if (sk_out) std::copy(sk.begin(), sk.end(), sk_out);
    return FHE_OK;
}
// serialised key-switching key: header {magic "FHEB200K", version 1, kind 3, log_n, L, bytes} | the evaluation-form image
struct CkksBlobHeader {
    char magic[8];
    uint32_t version, kind;
    uint64_t log_n, big_l, bytes;
};
size_t fhe_ckks_ksk_serialized_size(const fhe_ckks_ksk* ksk) { return ksk ? sizeof(CkksBlobHeader) + ksk->bytes : 0; }
fhe_status fhe_ckks_ksk_serialize(fhe_ctx* ctx, const fhe_ckks_ctx* ck, const fhe_ckks_ksk* ksk, void* buf, size_t cap) {
    if (!ctx || !ck || !ksk || !buf) return FHE_EINVAL;
    FHE_REQUIRE(ctx, cap >= fhe_ckks_ksk_serialized_size(ksk), "buffer too small for the serialised key");
    CkksBlobHeader h;
    memcpy(h.magic, "FHEB200K", 8);
    h.version = 1;
    h.kind = 3;
    h.log_n = ck->log_n;
    h.big_l = ck->big_l;
    h.bytes = ksk->bytes;
    memcpy(buf, &h, sizeof h);
    FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
    FHE_CUDA(ctx, cudaMemcpy((unsigned char*)buf + sizeof h, ksk->d_eval, ksk->bytes, cudaMemcpyDeviceToHost));
    return FHE_OK;
}
fhe_status fhe_ckks_ksk_deserialize(fhe_ctx* ctx, const fhe_ckks_ctx* ck, const void* buf, size_t len, fhe_ckks_ksk** out) {
    if (!ctx || !ck || !buf || !out) return FHE_EINVAL;
    *out = nullptr;
    CkksBlobHeader h;
    FHE_REQUIRE(ctx, len >= sizeof h, "serialised key truncated");
    memcpy(&h, buf, sizeof h);
    FHE_REQUIRE(ctx, memcmp(h.magic, "FHEB200K", 8) == 0 && h.kind == 3, "not a serialised CKKS key-switching key");
    FHE_REQUIRE(ctx, h.version == 1, "serialised key of another format version");
    const size_t want = 2 * 2 * ck->big_l * (((size_t)1) << ck->log_n) * 8;
    FHE_REQUIRE(ctx, h.log_n == ck->log_n && h.big_l == ck->big_l && h.bytes == want && len == sizeof h + want,
                "serialised key does not match this CKKS context (ring degree / modulus chain / length)");
    fhe_ckks_ksk* k = new fhe_ckks_ksk();
    k->bytes = want;
    if (cudaMalloc((void**)&k->d_eval, want) != cudaSuccess ||
        cudaMemcpy(k->d_eval, (const unsigned char*)buf + sizeof h, want, cudaMemcpyHostToDevice) != cudaSuccess) {
        if (k->d_eval) cudaFree(k->d_eval);
        delete k;
        return fail(ctx, FHE_ECUDA, "key image upload failed");
    }
    *out = k;
    return FHE_OK;
}

fhe_status fhe_ckks_mul_relin_rescale_batch(fhe_ctx* ctx, fhe_ckks_ctx* ck, const fhe_ckks_ksk* rlk, size_t level, size_t count,
                                            const uint64_t* d_ct0, const uint64_t* d_ct1, uint64_t* d_out) {
    if (!ctx || !ck || !rlk) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_ct0 && d_ct1 && d_out, "null pointer");
    FHE_REQUIRE(ctx, level >= 2 && level <= ck->big_l, "level must be in [2, L] (rescale needs a limb to drop)");
    const unsigned log_n = ck->log_n;
    const size_t n = (size_t)1 << log_n, l = level, L = ck->big_l, le = l + L;
    const std::vector<uint64_t> qs = level_qs(ck, l);
    // per-pair workspace (words): e0, e1 (2l each) | d01 (2l) | d2 eval (l) | d2 coeff (l) | xp (L) | kk (2(l+L)) | r (2l)
    const size_t per = (2 * l + 2 * l + 2 * l + l + l + L + 2 * le + 2 * l) * n;
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(count, ckks_ws_budget() / (per * 8)));
    uint64_t* ws;
    FHE_CHECK(ckks_ws(ctx, ck, chunk * per * 8, &ws));
    uint64_t* e0 = ws;
    uint64_t* e1 = e0 + chunk * 2 * l * n;
    uint64_t* d01 = e1 + chunk * 2 * l * n;
    uint64_t* d2e = d01 + chunk * 2 * l * n;
    uint64_t* d2c = d2e + chunk * l * n;
    uint64_t* xp = d2c + chunk * l * n;
    uint64_t* kk = xp + chunk * L * n;
    uint64_t* r = kk + chunk * 2 * le * n;
    for (size_t base = 0; base < count; base += chunk) {
        const size_t c = std::min(chunk, count - base);
        const uint64_t* in0 = d_ct0 + base * 2 * l * n;
        const uint64_t* in1 = d_ct1 + base * 2 * l * n;
        FHE_CHECK(launch_ntt_rns_u64_oop(ctx, qs.data(), l, log_n, c * 2 * l, in0, e0, true));
        FHE_CHECK(launch_ntt_rns_u64_oop(ctx, qs.data(), l, log_n, c * 2 * l, in1, e1, true));
        ckks_tensor_kernel<<<row_grid(ctx, log_n, (unsigned long long)c * l), CKKS_ROW_THREADS, 0, ctx->stream>>>(ck->d_mods, (int)l, (int)log_n, c, e0, e1,
                                                                                                         d01, d2e);
        FHE_CHECK(after_launch(ctx, "ckks_tensor_kernel"));
        FHE_CHECK(launch_ntt_rns_u64_oop(ctx, qs.data(), l, log_n, c * l, d2e, d2c, false));
        // relinearize(d2) with ct_b = 0, plus (d0, d1) folded in before the inverse transforms; then rescale (ckks.rs:266, 123-125)
        FHE_CHECK(key_switch_core(ctx, ck, rlk, l, c, d2c, d2e, xp, kk, nullptr, r, d01, d_out + base * 2 * (l - 1) * n));
    }
    return FHE_OK;
}

fhe_status fhe_ckks_mul_relin_rescale_batch_host(fhe_ctx* ctx, fhe_ckks_ctx* ck, const fhe_ckks_ksk* rlk, size_t level, size_t count,
                                                 const uint64_t* ct0, const uint64_t* ct1, uint64_t* out) {
    if (!ctx || !ck || !rlk) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, ct0 && ct1 && out, "null pointer");
    FHE_REQUIRE(ctx, level >= 2 && level <= ck->big_l, "level must be in [2, L]");
    const size_t n = (size_t)1 << ck->log_n;
    const size_t in_bytes = count * 2 * level * n * 8, out_bytes = count * 2 * (level - 1) * n * 8;
    void *d0, *d1, *dout;
    FHE_CHECK(ensure_stage_d(ctx, 0, in_bytes, &d0));
    FHE_CHECK(ensure_stage_d(ctx, 1, in_bytes, &d1));
    FHE_CHECK(ensure_stage_d(ctx, 2, out_bytes, &dout));
    // Pipelined over chunks: the H2D copies of chunk c+1 and the D2H copy of chunk c-1 overlap the kernels of chunk c (two copy
    // streams + events; PCIe is full duplex).  The operation moves 4 level + 2 (level - 1) limbs per pair over the bus for ~0.1 ms
    // of kernels per pair, so the host path is bound by the H2D direction once the three stages overlap.
    size_t nchunk = count >= 32 ? 8 : (count >= 8 ? 4 : (count >= 2 ? 2 : 1));
    if (const char* e = getenv("FHE_B200_HOST_CHUNKS")) nchunk = std::max<size_t>(1, std::min<size_t>((size_t)atoi(e), count));  // tuning knob
    const size_t cs = (count + nchunk - 1) / nchunk, in_pair = 2 * level * n, out_pair = 2 * (level - 1) * n;
    if (!ctx->copy_in) FHE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_in, cudaStreamNonBlocking));
    if (!ctx->copy_out) FHE_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->copy_out, cudaStreamNonBlocking));
    std::vector<cudaEvent_t> ev(2 * nchunk + 1);  // per chunk: inputs landed, kernels done; last: fence
    const size_t fence = 2 * nchunk;
    for (auto& e : ev) FHE_CUDA(ctx, cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
    fhe_status st = FHE_OK;
    auto cu = [&](cudaError_t e, const char* what) {
        if (e != cudaSuccess && st == FHE_OK) st = fail(ctx, FHE_ECUDA, "%s: %s", what, cudaGetErrorString(e));
    };
    cu(cudaEventRecord(ev[fence], ctx->stream), "event");  // the staging buffers are free once earlier work on the stream is done
    cu(cudaStreamWaitEvent(ctx->copy_in, ev[fence], 0), "wait");
    for (size_t c = 0; c < nchunk && st == FHE_OK; ++c) {
        const size_t off = c * cs;
        if (off >= count) break;
        const size_t cnt = std::min(cs, count - off);
        uint64_t *c0 = (uint64_t*)d0 + off * in_pair, *c1 = (uint64_t*)d1 + off * in_pair, *co = (uint64_t*)dout + off * out_pair;
        cu(cudaMemcpyAsync(c0, ct0 + off * in_pair, cnt * in_pair * 8, cudaMemcpyHostToDevice, ctx->copy_in), "H2D");
        cu(cudaMemcpyAsync(c1, ct1 + off * in_pair, cnt * in_pair * 8, cudaMemcpyHostToDevice, ctx->copy_in), "H2D");
        cu(cudaEventRecord(ev[2 * c], ctx->copy_in), "event");
        cu(cudaStreamWaitEvent(ctx->stream, ev[2 * c], 0), "wait");
        if (st == FHE_OK) st = fhe_ckks_mul_relin_rescale_batch(ctx, ck, rlk, level, cnt, c0, c1, co);
        cu(cudaEventRecord(ev[2 * c + 1], ctx->stream), "event");
        cu(cudaStreamWaitEvent(ctx->copy_out, ev[2 * c + 1], 0), "wait");
        cu(cudaMemcpyAsync(out + off * out_pair, co, cnt * out_pair * 8, cudaMemcpyDeviceToHost, ctx->copy_out), "D2H");
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

fhe_status fhe_ckks_key_switch(fhe_ctx* ctx, fhe_ckks_ctx* ck, const fhe_ckks_ksk* ksk, int64_t t, size_t level, size_t count,
                               const uint64_t* d_ct, uint64_t* d_out) {
    if (!ctx || !ck || !ksk) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_ct && d_out, "null pointer");
    FHE_REQUIRE(ctx, level >= 1 && level <= ck->big_l, "level must be in [1, L]");
    const unsigned log_n = ck->log_n;
    const size_t n = (size_t)1 << log_n, l = level, L = ck->big_l, le = l + L;
    const std::vector<uint64_t> qs = level_qs(ck, l);
    // per-ciphertext workspace (words): ct' (2l) | a coeff (l) | a eval (l) | xp (L) | kk (2(l+L))
    const size_t per = (2 * l + l + l + L + 2 * le) * n;
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(count, ckks_ws_budget() / (per * 8)));
    uint64_t* ws;
    FHE_CHECK(ckks_ws(ctx, ck, chunk * per * 8, &ws));
    uint64_t* cta = ws;
    uint64_t* ac = cta + chunk * 2 * l * n;
    uint64_t* ae = ac + chunk * l * n;
    uint64_t* xp = ae + chunk * l * n;
    uint64_t* kk = xp + chunk * L * n;
    const int64_t m2 = 2 * (int64_t)n;
    const uint32_t tt = (uint32_t)(((t % m2) + m2) % m2);
    for (size_t base = 0; base < count; base += chunk) {
        const size_t c = std::min(chunk, count - base);
        const uint64_t* in = d_ct + base * 2 * l * n;
        const uint64_t* ct = in;
        if (t != 0) {  // CkksCiphertext::automorphism (ckks.rs:127-129)
            rns_automorphism_kernel<<<stream_grid(ctx, (unsigned long long)c * 2 * l << log_n), 256, 0, ctx->stream>>>(ck->d_mods, (int)l, (int)log_n,
                                                                                                                       c * 2 * l, tt, in, cta);
            FHE_CHECK(after_launch(ctx, "rns_automorphism_kernel"));
            ct = cta;
        }
        // gather the a halves ([c][1][l][n]) contiguously
        FHE_CUDA(ctx, cudaMemcpy2DAsync(ac, l * n * 8, ct + l * n, 2 * l * n * 8, l * n * 8, c, cudaMemcpyDeviceToDevice, ctx->stream));
        FHE_CHECK(launch_ntt_rns_u64_oop(ctx, qs.data(), l, log_n, c * l, ac, ae, true));
        FHE_CHECK(key_switch_core(ctx, ck, ksk, l, c, ac, ae, xp, kk, ct, d_out + base * 2 * l * n));
    }
    return FHE_OK;
}

// Ckks::mul_constant after encoding (ckks.rs:250-253): (pt * ct.b, pt * ct.a).rescale(); pt [pt_count][level][N] coefficient form,
// pt_count == 1 (shared by the batch) or == count
fhe_status fhe_ckks_mul_plain_rescale_batch(fhe_ctx* ctx, fhe_ckks_ctx* ck, size_t level, size_t count, size_t pt_count, const uint64_t* d_pt,
                                            const uint64_t* d_ct, uint64_t* d_out) {
    if (!ctx || !ck) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_pt && d_ct && d_out, "null pointer");
    FHE_REQUIRE(ctx, level >= 2 && level <= ck->big_l, "level must be in [2, L] (rescale needs a limb to drop)");
    FHE_REQUIRE(ctx, pt_count == 1 || pt_count == count, "pt_count must be 1 or count");
    const unsigned log_n = ck->log_n;
    const size_t n = (size_t)1 << log_n, l = level;
    const std::vector<uint64_t> qs = level_qs(ck, l);
    const size_t per = 2 * l * n;
    const size_t chunk = std::max<size_t>(1, std::min<size_t>(count, ((size_t)4 << 30) / (per * 8)));
    uint64_t* ws;
    FHE_CHECK(ckks_ws(ctx, ck, (pt_count * l * n + chunk * per) * 8, &ws));
    uint64_t* pe = ws;
    uint64_t* e = pe + pt_count * l * n;
    FHE_CHECK(launch_ntt_rns_u64_oop(ctx, qs.data(), l, log_n, pt_count * l, d_pt, pe, true));
    for (size_t base = 0; base < count; base += chunk) {
        const size_t c = std::min(chunk, count - base);
        FHE_CHECK(launch_ntt_rns_u64_oop(ctx, qs.data(), l, log_n, c * 2 * l, d_ct + base * per, e, true));
        ckks_ptmul_kernel<<<stream_grid(ctx, (unsigned long long)c * per), 256, 0, ctx->stream>>>(
            ck->d_mods, (int)l, (int)log_n, c, pt_count == 1 ? 0 : 1, pe + (pt_count == 1 ? 0 : base * l * n), e, e);
        FHE_CHECK(after_launch(ctx, "ckks_ptmul_kernel"));
        FHE_CHECK(launch_ntt_rns_u64(ctx, qs.data(), l, log_n, c * 2 * l, e, false));
        FHE_CHECK(run_rescale(ctx, qs, 1, log_n, c * 2, e, nullptr, false, d_out + base * 2 * (l - 1) * n));
    }
    return FHE_OK;
}

// Bootstrapping::mul_mat (scheme/ckks/src/bootstrapping.rs:92-108): baby-step / giant-step product of a diagonal-sparse matrix
// with `count` ciphertexts: out = sum_i rot_{g_i}( sum_j mul_constant(pt_ij, rot_{b_j}(ct)) ), every mul_constant rescaling
// before the sums exactly like the reference.  The BSGS plan and the encoded diagonals come from the host (misc/matrix.rs,
// sfft.rs: out of scope); this evaluates them with the rotation / plaintext-product kernels above.
fhe_status fhe_ckks_mul_mat(fhe_ctx* ctx, fhe_ckks_ctx* ck, size_t level, size_t count, size_t n_baby, const fhe_ckks_rot* baby, size_t n_giant,
                            const fhe_ckks_rot* giant, const uint8_t* present, const uint64_t* d_pts, const uint64_t* d_ct, uint64_t* d_out) {
    if (!ctx || !ck) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, baby && giant && present && d_pts && d_ct && d_out && n_baby >= 1 && n_giant >= 1, "null pointer or empty plan");
    FHE_REQUIRE(ctx, level >= 2 && level <= ck->big_l, "level must be in [2, L] (mul_constant rescales)");
    const unsigned log_n = ck->log_n;
    const size_t n = (size_t)1 << log_n, l = level, ct_in = count * 2 * l * n, ct_out = count * 2 * (l - 1) * n;
    for (size_t j = 0; j < n_baby; ++j) FHE_REQUIRE(ctx, baby[j].t == 0 || baby[j].key, "baby step %zu has no rotation key", j);
    for (size_t i = 0; i < n_giant; ++i) {
        FHE_REQUIRE(ctx, giant[i].t == 0 || giant[i].key, "giant step %zu has no rotation key", i);
        bool any = false;
        for (size_t j = 0; j < n_baby; ++j) any = any || present[i * n_baby + j];
        FHE_REQUIRE(ctx, any, "giant step %zu has no diagonal", i);
    }
    const size_t words = n_baby * ct_in + ct_in + l * n + 3 * ct_out;
    if (ck->ws2_bytes < words * 8) {
        FHE_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
        if (ck->ws2) cudaFree(ck->ws2);
        ck->ws2 = nullptr;
        ck->ws2_bytes = 0;
        if (cudaMalloc(&ck->ws2, words * 8) != cudaSuccess) return fail(ctx, FHE_ENOMEM, "mul_mat workspace of %zu bytes", words * 8);
        ck->ws2_bytes = words * 8;
    }
    uint64_t* rot = (uint64_t*)ck->ws2;             // [n_baby] rotated inputs, transformed ONCE to evaluation form
    uint64_t* prod = rot + n_baby * ct_in;          // one plaintext product (evaluation -> coefficient form)
    uint64_t* pe = prod + ct_in;                    // one diagonal in evaluation form
    uint64_t* inner = pe + l * n;                   // sum over baby steps of one giant step
    uint64_t* tmp = inner + ct_out;
    uint64_t* tmp2 = tmp + ct_out;
    auto add = [&](const uint64_t* a, const uint64_t* b, uint64_t* o) -> fhe_status {
        rns_add_kernel<<<stream_grid(ctx, (unsigned long long)ct_out), 256, 0, ctx->stream>>>(ck->d_mods, (int)(l - 1), (int)log_n,
                                                                                               count * 2 * (l - 1), a, b, o);
        return after_launch(ctx, "rns_add_kernel");
    };
    // every rotated input is used by up to n_giant diagonals: its forward transform is shared (the reference transforms
    // inside each polynomial product; the products are exact, so the results are the same words)
    const std::vector<uint64_t> qs = level_qs(ck, l);
    for (size_t j = 0; j < n_baby; ++j) {
        uint64_t* e = rot + j * ct_in;
        if (baby[j].t == 0) {
            FHE_CHECK(launch_ntt_rns_u64_oop(ctx, qs.data(), l, log_n, count * 2 * l, d_ct, e, true));
        } else {
            FHE_CHECK(fhe_ckks_key_switch(ctx, ck, baby[j].key, baby[j].t, l, count, d_ct, e));
            FHE_CHECK(launch_ntt_rns_u64(ctx, qs.data(), l, log_n, count * 2 * l, e, true));
        }
    }
    size_t pt_idx = 0;
    for (size_t i = 0; i < n_giant; ++i) {
        bool first = true;
        for (size_t j = 0; j < n_baby; ++j) {
            if (!present[i * n_baby + j]) continue;
            const uint64_t* pt = d_pts + pt_idx * l * n;
            ++pt_idx;
            // mul_constant (ckks.rs:250-253): limb-wise product with the encoded diagonal, then rescale
            FHE_CHECK(launch_ntt_rns_u64_oop(ctx, qs.data(), l, log_n, l, pt, pe, true));
            ckks_ptmul_kernel<<<stream_grid(ctx, (unsigned long long)ct_in), 256, 0, ctx->stream>>>(ck->d_mods, (int)l, (int)log_n, count, 0, pe,
                                                                                                    rot + j * ct_in, prod);
            FHE_CHECK(after_launch(ctx, "ckks_ptmul_kernel"));
            FHE_CHECK(launch_ntt_rns_u64(ctx, qs.data(), l, log_n, count * 2 * l, prod, false));
            FHE_CHECK(run_rescale(ctx, qs, 1, log_n, count * 2, prod, nullptr, false, first ? inner : tmp));
            if (!first) FHE_CHECK(add(inner, tmp, inner));
            first = false;
        }
        const uint64_t* term = inner;
        if (giant[i].t != 0) {
            FHE_CHECK(fhe_ckks_key_switch(ctx, ck, giant[i].key, giant[i].t, l - 1, count, inner, tmp2));
            term = tmp2;
        }
        if (i == 0)
            FHE_CUDA(ctx, cudaMemcpyAsync(d_out, term, ct_out * 8, cudaMemcpyDeviceToDevice, ctx->stream));
        else
            FHE_CHECK(add(d_out, term, d_out));
    }
    return FHE_OK;
}

fhe_status fhe_ckks_rescale(fhe_ctx* ctx, fhe_ckks_ctx* ck, size_t level, size_t count, const uint64_t* d_ct, uint64_t* d_out) {
    if (!ctx || !ck) return FHE_EINVAL;
    if (count == 0) return FHE_OK;
    FHE_REQUIRE(ctx, d_ct && d_out, "null pointer");
    FHE_REQUIRE(ctx, level >= 2 && level <= ck->big_l, "level must be in [2, L]");
    return run_rescale(ctx, level_qs(ck, level), 1, ck->log_n, count * 2, d_ct, nullptr, false, d_out);
}

}  // extern "C"
