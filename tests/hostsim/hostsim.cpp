// TEST INFRASTRUCTURE ONLY — CPU simulation of the CUDA kernels' per-thread logic.
//
// The kernels in learn-fhe_b200/csrc keep all index arithmetic and modular arithmetic in __host__ __device__
// functions.  This file compiles those same headers with g++ and replays each kernel's control flow
// sequentially (thread id loops where the kernel has threads, nothing where it has __syncthreads()), so that
// the logic can be checked against the oracle in the CPU-only test tier.  It is never loaded by the product
// path (libfhe_b200.so has no CPU fallback).
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <vector>

#include "../../learn-fhe_b200/csrc/fhew_core.cuh"
#include "../../learn-fhe_b200/csrc/host_tables.hpp"
#include "../../learn-fhe_b200/csrc/modarith.cuh"
#include "../../learn-fhe_b200/csrc/ntt_core.cuh"
#include "../../learn-fhe_b200/csrc/ntt_fast.cuh"

using namespace fhe;

template <typename A>
static A make_mod_sim(uint64_t q);
template <>
Mod32 make_mod_sim<Mod32>(uint64_t q) {
    Mod32 m;
    m.q = (uint32_t)q;
    m.q2 = (uint32_t)(2 * q);
    m.mu = (uint64_t)((((u128_t)1) << 64) / q);
    return m;
}
template <>
Mod64 make_mod_sim<Mod64>(uint64_t q) {
    Mod64 m;
    m.q = q;
    m.q2 = 2 * q;
    unsigned s = 0;
    while (s < 64 && (q >> s)) ++s;
    if (s < 2) s = 2;
    m.s = s;
    m.mu = (uint64_t)((((u128_t)1) << (2 * s)) / q);
    return m;
}
template <typename W>
static TwPair<W> twp(uint64_t w, uint64_t q);
template <>
TwPair<uint32_t> twp<uint32_t>(uint64_t w, uint64_t q) {
    return TwPair<uint32_t>{(uint32_t)w, host_shoup32((uint32_t)w, (uint32_t)q)};
}
template <>
TwPair<uint64_t> twp<uint64_t>(uint64_t w, uint64_t q) {
    return TwPair<uint64_t>{w, host_shoup64(w, q)};
}

template <typename A, int S>
static void sim_column(const A& m, typename A::W* data, int log_n, const TwPair<typename A::W>* tw, bool fwd,
                       TwPair<typename A::W> ninv, TwPair<typename A::W> wninv) {
    typedef typename A::W W;
    const int lc = log_n - S;
    for (uint32_t col = 0; col < (1u << lc); ++col) {
        W x[1 << S];
        for (int j = 0; j < (1 << S); ++j) x[j] = data[((size_t)j << lc) + col];
        if (fwd) {
            fwd_column_regs<A, S>(m, x, tw);
        } else {
            inv_column_regs<A, S>(m, x, tw, ninv, wninv);
            for (int j = 0; j < (1 << S); ++j) x[j] = m.redq(x[j]);
        }
        for (int j = 0; j < (1 << S); ++j) data[((size_t)j << lc) + col] = x[j];
    }
}

// mirrors ntt_tile_kernel for one polynomial
template <typename A>
static void sim_tiles(const A& m, typename A::W* data, int log_n, int c, const TwPair<typename A::W>* tw, bool fwd, uint32_t nthr,
                      TwPair<typename A::W> ninv, TwPair<typename A::W> wninv) {
    typedef typename A::W W;
    const int s0 = log_n - c;
    const uint32_t C = 1u << c;
    std::vector<W> s(C);
    const PassPlan plan = make_plan(c);
    const bool final_out = fwd || s0 == 0;
    for (uint32_t k = 0; k < (1u << s0); ++k) {
        W* g = data + ((size_t)k << c);
        for (uint32_t i = 0; i < C; ++i) s[swz<W>(i)] = g[i];
        if (fwd) {
            for (int pi = 0; pi < plan.n; ++pi)
                for (uint32_t tid = 0; tid < nthr; ++tid) fwd_tile_pass<A>(m, s.data(), c, plan.t0[pi], plan.r[pi], s0, k, tid, nthr, tw);
        } else {
            for (int pi = plan.n - 1; pi >= 0; --pi) {
                const bool last = (s0 == 0) && (plan.t0[pi] == 0);
                for (uint32_t tid = 0; tid < nthr; ++tid)
                    inv_tile_pass<A>(m, s.data(), c, plan.t0[pi], plan.r[pi], s0, k, tid, nthr, tw, last, ninv, wninv);
            }
        }
        for (uint32_t i = 0; i < C; ++i) {
            W x = s[swz<W>(i)];
            if (final_out) x = fwd ? m.canon4(x) : m.redq(x);
            g[i] = x;
        }
    }
}

template <typename A>
static int sim_ntt(uint64_t q, unsigned log_n, int c, typename A::W* a, int fwd, uint32_t nthr) {
    typedef typename A::W W;
    if (log_n == 0) return 0;
    std::vector<uint64_t> hf, hi;
    if (!host_build_twiddles(q, (size_t)1 << log_n, hf, hi)) return -1;
    size_t n = (size_t)1 << log_n;
    std::vector<TwPair<W>> tf(n), ti(n);
    for (size_t j = 0; j < n; ++j) {
        tf[j] = twp<W>(hf[j], q);
        ti[j] = twp<W>(hi[j], q);
    }
    A m = make_mod_sim<A>(q);
    uint64_t ninv = host_powmod(n % q, q - 2, q);
    TwPair<W> pn = twp<W>(ninv, q), pw = twp<W>(host_mulmod(hi[1], ninv, q), q);
    if (c > (int)log_n) c = (int)log_n;
    const int S = (int)log_n - c;
    if (S > 4) return -2;
    if (fwd) {
        if (S == 1) sim_column<A, 1>(m, a, log_n, tf.data(), true, pn, pw);
        if (S == 2) sim_column<A, 2>(m, a, log_n, tf.data(), true, pn, pw);
        if (S == 3) sim_column<A, 3>(m, a, log_n, tf.data(), true, pn, pw);
        if (S == 4) sim_column<A, 4>(m, a, log_n, tf.data(), true, pn, pw);
        sim_tiles<A>(m, a, log_n, c, tf.data(), true, nthr, pn, pw);
    } else {
        sim_tiles<A>(m, a, log_n, c, ti.data(), false, nthr, pn, pw);
        if (S == 1) sim_column<A, 1>(m, a, log_n, ti.data(), false, pn, pw);
        if (S == 2) sim_column<A, 2>(m, a, log_n, ti.data(), false, pn, pw);
        if (S == 3) sim_column<A, 3>(m, a, log_n, ti.data(), false, pn, pw);
        if (S == 4) sim_column<A, 4>(m, a, log_n, ti.data(), false, pn, pw);
    }
    return 0;
}

extern "C" {

int sim_ntt_u64(uint64_t q, unsigned log_n, int c, uint64_t* a, int fwd, unsigned nthr) { return sim_ntt<Mod64>(q, log_n, c, a, fwd, nthr); }
int sim_ntt_u32(uint64_t q, unsigned log_n, int c, uint32_t* a, int fwd, unsigned nthr) { return sim_ntt<Mod32>(q, log_n, c, a, fwd, nthr); }

// modular primitives (host compilation of the same inline functions the kernels use)
uint64_t sim_mul_u64(uint64_t q, uint64_t a, uint64_t b) { return make_mod_sim<Mod64>(q).mul(a, b); }
uint32_t sim_mul_u32(uint32_t q, uint32_t a, uint32_t b) { return make_mod_sim<Mod32>(q).mul(a, b); }
uint32_t sim_reduce64_u32(uint32_t q, uint64_t x) { return make_mod_sim<Mod32>(q).reduce64(x); }
uint64_t sim_mulhi_approx(uint64_t a, uint64_t b) { return mulhi_u64_approx(a, b); }
uint64_t sim_shoup_u64(uint64_t q, uint64_t y, uint64_t w) { return make_mod_sim<Mod64>(q).shoup_lazy(y, w, host_shoup64(w, q)); }
uint32_t sim_shoup_u32(uint32_t q, uint32_t y, uint32_t w) { return make_mod_sim<Mod32>(q).shoup_lazy(y, w, host_shoup32(w, q)); }
// swizzle bank-conflict analysis helper: returns max multiplicity of a bank among `lanes` consecutive lanes
unsigned sim_swz32(unsigned p) { return swz<uint32_t>(p); }
unsigned sim_swz64(unsigned p) { return swz<uint64_t>(p); }

}  // extern "C"

#include "hostsim_fast.inc"
#include "hostsim_fhew.inc"
