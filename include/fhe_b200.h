/*
 * fhe_b200.h — C ABI of libfhe_b200.so: the B200-native (sm_100a) polynomial-ring engine that stands in for the
 * hot path of han0110/learn-fhe (util crate ring/polynomial/decomposition API + the FHEW / TFHE / CKKS call sites).
 *
 * The reference has no FFI boundary: its de-facto boundary is the `util` crate's public Rust API
 * (util/src/lib.rs:8-21) and the scheme-level associated functions built on it.  Every entry point below cites the
 * reference interface (file:line relative to the reference root) that a Rust `-sys` shim would forward to it; see
 * INTEGRATION.md for the binding.
 *
 * Conventions
 *   - plain pointers and sizes only; every function returns an fhe_status (0 = ok) and never unwinds.  Conditions
 *     that `assert!`/`unwrap()`-panic in the reference (length mismatch, non power-of-two degree, modulus without
 *     enough 2-adicity, ...) are reported as FHE_EINVAL.
 *   - `Zq` values are canonical residues in [0, q) stored structure-of-arrays as uint64_t (the reference stores
 *     {q, v} per element, util/src/zq.rs:21-26); `T64` values are raw uint64_t torus words (util/src/torus.rs:12).
 *   - pointers named d_* are DEVICE pointers (cudaMalloc / torch tensors / fhe_malloc); functions with a `_host`
 *     suffix take HOST pointers and stage through the device inside the call (these are the drop-in forms of the
 *     reference's in-place slice functions and what bench.py's `e2e` times).
 *   - all device work is enqueued on the context's stream; `fhe_sync` waits for it.  `_host` functions return
 *     after the result is in the caller's buffer.
 *   - polynomials are contiguous coefficient arrays of length n = 2^log_n; batches are contiguous polynomials;
 *     RNS polynomials are limb-major (util/src/ring/rns.rs:21); decompositions are limb-major
 *     (util/src/misc/decompose.rs:137-155).
 *   - there is NO CPU fallback: every entry point fails with FHE_ECUDA if no sm_100-class device is usable.
 */
#ifndef FHE_B200_H
#define FHE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int fhe_status;
enum { FHE_OK = 0, FHE_EINVAL = 1, FHE_ECUDA = 2, FHE_ENOMEM = 3, FHE_EUNSUPPORTED = 4 };

typedef struct fhe_ctx fhe_ctx;

/* ---- context ------------------------------------------------------------------------------------------------ */
/* One context per device/process: owns the stream, the twiddle / Shoup tables keyed by modulus (the reference
 * keeps a global Mutex<HashMap<q, ..>>, util/src/ring/fft/zq.rs:38-56) and uploaded keys. */
fhe_status fhe_ctx_create(int device, fhe_ctx** out);
void fhe_ctx_destroy(fhe_ctx* ctx);
/* use an externally owned cudaStream_t (e.g. torch.cuda.current_stream().cuda_stream); NULL restores the own stream */
fhe_status fhe_ctx_set_stream(fhe_ctx* ctx, void* cuda_stream);
fhe_status fhe_sync(fhe_ctx* ctx);
const char* fhe_last_error(const fhe_ctx* ctx);
/* number of kernels this context has launched so far (bench.py's gpu_launches) */
uint64_t fhe_launch_count(const fhe_ctx* ctx);
int fhe_sm_count(const fhe_ctx* ctx);
const char* fhe_version(void);

/* per-launch timing log: after fhe_prof_begin every kernel launch of this context is followed by a CUDA event on the
 * context's stream; fhe_prof_end synchronises and writes a JSON object {"kernel name": {"ms": total, "launches": n}}
 * (time between consecutive events = device time of that launch when launches are back to back). bench.py's
 * roofline leg uses it; it is off by default and costs nothing then. */
fhe_status fhe_prof_begin(fhe_ctx* ctx);
fhe_status fhe_prof_end(fhe_ctx* ctx, char* json_buf, size_t cap);

/* measured integer-multiply peaks of this device in 10^12 thread-level operations per second: 32-bit IMAD, IMAD.HI
 * (__umulhi) and IMAD.WIDE (u32 x u32 + u64).  The modular kernels' "binding roofline" when they are not HBM-bound. */
fhe_status fhe_diag_int32_peak(fhe_ctx* ctx, double* imad_tops, double* imad_hi_tops, double* imad_wide_tops);
/* measured FP64 pipe rates (10^12 thread-instructions per second): DADD, DMUL, DFMA - the denominators of the TFHE roofline */
fhe_status fhe_diag_fp64_peak(fhe_ctx* ctx, double* dadd_tops, double* dmul_tops, double* dfma_tops);
/* measured rate (10^12 butterflies per second) of the u32 lazy radix-16 register pass with the Shoup quotient formed by IMAD.HI on
 * the multiply pipe and by one DFMA (rounded down) on the FP64 pipe; *same_out = 1 when both variants produced identical words */
fhe_status fhe_diag_butterfly_rate(fhe_ctx* ctx, double* imad_hi_tbf, double* dfma_tbf, int* same_out);

fhe_status fhe_malloc(fhe_ctx* ctx, size_t bytes, void** d_ptr);
fhe_status fhe_free(fhe_ctx* ctx, void* d_ptr);
fhe_status fhe_memcpy_h2d(fhe_ctx* ctx, void* d_dst, const void* h_src, size_t bytes);
fhe_status fhe_memcpy_d2h(fhe_ctx* ctx, void* h_dst, const void* d_src, size_t bytes);

/* two_adic_primes(bits, log_n) (util/src/zq.rs:325-329): the first `count` primes q < 2^bits with q = 1 mod 2^log_n,
 * descending.  Host-side setup helper (CkksParam::new, test parameter sets). */
fhe_status fhe_two_adic_primes(unsigned bits, unsigned log_n, size_t count, uint64_t* out);

/* ---- Zq negacyclic NTT (util/src/ring/fft/zq.rs:27-36, ring.rs:140-144,180-184) ------------------------------- */
/* In place, `batch` polynomials of degree 2^log_n.  Forward: natural order in, bit-reversed order out (exactly the
 * reference's evaluation order and root choice); inverse: the converse, including the n^-1 factor.
 * u64 path: prime q < 2^62.  u32 path: prime q < 2^30 (FHEW's 28-bit Q). */
fhe_status fhe_ntt_fwd_u64(fhe_ctx* ctx, uint64_t q, unsigned log_n, size_t batch, uint64_t* d_a);
fhe_status fhe_ntt_inv_u64(fhe_ctx* ctx, uint64_t q, unsigned log_n, size_t batch, uint64_t* d_a);
fhe_status fhe_ntt_fwd_u32(fhe_ctx* ctx, uint32_t q, unsigned log_n, size_t batch, uint32_t* d_a);
fhe_status fhe_ntt_inv_u32(fhe_ctx* ctx, uint32_t q, unsigned log_n, size_t batch, uint32_t* d_a);
/* RNS form: `limbs` moduli qs[i]; data is [batch][limbs][n] (limb-major per polynomial, rns.rs:21) */
fhe_status fhe_ntt_fwd_rns(fhe_ctx* ctx, const uint64_t* qs, size_t limbs, unsigned log_n, size_t batch, uint64_t* d_a);
fhe_status fhe_ntt_inv_rns(fhe_ctx* ctx, const uint64_t* qs, size_t limbs, unsigned log_n, size_t batch, uint64_t* d_a);
/* host-slice forms of nega_cyclic_ntt_in_place / nega_cyclic_intt_in_place (fft/zq.rs:27-36) */
fhe_status fhe_ntt_fwd_host(fhe_ctx* ctx, uint64_t q, uint64_t* a, size_t n, size_t batch);
fhe_status fhe_ntt_inv_host(fhe_ctx* ctx, uint64_t q, uint64_t* a, size_t n, size_t batch);
/* bit-reversed twiddle table the context uses for q (first `len` entries; compute_twiddle, fft/zq.rs:58-67) */
fhe_status fhe_twiddles_host(fhe_ctx* ctx, uint64_t q, size_t len, uint64_t* fwd, uint64_t* inv);

/* ---- Zq element-wise and ring ops ------------------------------------------------------------------------------ */
/* coefficient-form product a <- a * b in Z_q[X]/(X^n+1): nega_cyclic_ntt_mul_assign (fft/zq.rs:14-25),
 * `Rq *= &Rq` (ring.rs:256-264).  b is not modified. */
fhe_status fhe_negacyclic_mul_u64(fhe_ctx* ctx, uint64_t q, unsigned log_n, size_t batch, uint64_t* d_a, const uint64_t* d_b);
fhe_status fhe_negacyclic_mul_host(fhe_ctx* ctx, uint64_t q, uint64_t* a, const uint64_t* b, size_t n, size_t batch);
/* evaluation-form (or any element-wise) ops over `count` residues: ring.rs:266-270, avec.rs:166-291, zq.rs:156-196 */
fhe_status fhe_pointwise_mul_u64(fhe_ctx* ctx, uint64_t q, size_t count, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out);
fhe_status fhe_pointwise_mac_u64(fhe_ctx* ctx, uint64_t q, size_t count, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_acc);
fhe_status fhe_vec_add_u64(fhe_ctx* ctx, uint64_t q, size_t count, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out);
fhe_status fhe_vec_sub_u64(fhe_ctx* ctx, uint64_t q, size_t count, const uint64_t* d_a, const uint64_t* d_b, uint64_t* d_out);
fhe_status fhe_vec_neg_u64(fhe_ctx* ctx, uint64_t q, size_t count, const uint64_t* d_a, uint64_t* d_out);
fhe_status fhe_vec_scalar_mul_u64(fhe_ctx* ctx, uint64_t q, size_t count, const uint64_t* d_a, uint64_t scalar, uint64_t* d_out);
/* AVec::automorphism (avec.rs:34-50): out[(i*t) mod 2n (negated past n)] = in[i]; q == 0 selects T64 (wrapping) */
fhe_status fhe_automorphism_u64(fhe_ctx* ctx, uint64_t q, unsigned log_n, size_t batch, int64_t t, const uint64_t* d_in, uint64_t* d_out);
/* `poly * (X ^ k)` (ring.rs:299-313, 380-406); q == 0 selects T64 */
fhe_status fhe_monomial_mul_u64(fhe_ctx* ctx, uint64_t q, unsigned log_n, size_t batch, int64_t k, const uint64_t* d_in, uint64_t* d_out);
/* Zq::mod_switch / mod_switch_odd (zq.rs:128-140; avec.rs:61-67), f64 arithmetic bit-identical to the reference */
fhe_status fhe_mod_switch_u64(fhe_ctx* ctx, uint64_t q, uint64_t q_prime, size_t count, const uint64_t* d_in, uint64_t* d_out);
fhe_status fhe_mod_switch_odd_u64(fhe_ctx* ctx, uint64_t q, uint64_t q_prime, size_t count, const uint64_t* d_in, uint64_t* d_out);

/* ---- gadget decomposition (util/src/misc/decompose.rs) --------------------------------------------------------- */
/* Base2Decomposor<Zq>::decompose (decompose.rs:42-64,91-112): out is [d][count] limb-major, digits as residues mod q */
fhe_status fhe_decompose_zq(fhe_ctx* ctx, uint64_t q, unsigned log_b, unsigned d, size_t count, const uint64_t* d_in, uint64_t* d_out);
/* Base2Decomposor<T64>::decompose (decompose.rs:66-81,114-135): out is [d][count] limb-major, wrapping digits */
fhe_status fhe_decompose_t64(fhe_ctx* ctx, unsigned log_b, unsigned d, size_t count, const uint64_t* d_in, uint64_t* d_out);
/* T64::rounding_shr (decompose.rs:115-118) */
fhe_status fhe_rounding_shr_t64(fhe_ctx* ctx, unsigned bits, size_t count, const uint64_t* d_in, uint64_t* d_out);

/* ---- T64 negacyclic product by f64 FFT (util/src/ring/fft/c64.rs:11-108; `Rt *= &Rt`, ring.rs:315-320) ---------- */
/* a <- a * b over the torus, bit-identical to the reference's floating-point algorithm (same operation order,
 * no FMA contraction). */
fhe_status fhe_fft64_negacyclic_mul(fhe_ctx* ctx, unsigned log_n, size_t batch, uint64_t* d_a, const uint64_t* d_b);
fhe_status fhe_fft64_negacyclic_mul_host(fhe_ctx* ctx, uint64_t* a, const uint64_t* b, size_t n, size_t batch);

/* ---- RNS (util/src/ring/rns.rs) -------------------------------------------------------------------------------- */
/* RnsRq::extend_bases (rns.rs:83-91, 331-345): in [batch][nq][n] -> out [batch][nq+np][n] (input limbs copied) */
fhe_status fhe_rns_extend_bases(fhe_ctx* ctx, const uint64_t* qs, size_t nq, const uint64_t* ps, size_t np, unsigned log_n,
                                size_t batch, const uint64_t* d_in, uint64_t* d_out);
/* RnsRq::rescale_k (rns.rs:99-132): in [batch][nq][n] -> out [batch][nq-k][n] */
fhe_status fhe_rns_rescale_k(fhe_ctx* ctx, const uint64_t* qs, size_t nq, size_t k, unsigned log_n, size_t batch,
                             const uint64_t* d_in, uint64_t* d_out);

/* ---- FHEW / LMKCDEY (scheme/fhew/src/{lwe,rlwe,rgsw,bootstrapping,fhew}.rs) ------------------------------------- */
typedef struct fhe_fhew_param {
    unsigned log_n;      /* ring degree N = 2^log_n                                   (rlwe.rs:13-20)       */
    uint64_t big_q;      /* RLWE/RGSW modulus Q, prime, Q = 1 mod 2N: < 2^30 (boolean.rs:225-239; 32-bit kernels, fused fast path at
                          * N = 512) or < 2^62 (examples/multi_key_uint8.rs:15-29: 55 bits, N = 2048, d = 5; 64-bit generic kernels) */
    uint64_t p;          /* plaintext modulus                                                             */
    unsigned rlwe_log_b, rlwe_d;   /* RLWE key-switch decomposor (automorphism keys)  (rlwe.rs:17-19)       */
    unsigned rgsw_log_b, rgsw_d;   /* RGSW decomposor                                 (rgsw.rs:18-27)       */
    unsigned n_s;        /* LWE_s dimension                                           (lwe.rs:17-27)        */
    uint64_t q_ks;       /* LWE_s modulus, a power of two <= 2^32                                          */
    unsigned ks_log_b, ks_d;       /* LWE key-switch decomposor                       (lwe.rs:29-32)        */
    unsigned w;          /* LMKCDEY window                                            (bootstrapping.rs:21-32) */
} fhe_fhew_param;
typedef struct fhe_fhew_key fhe_fhew_key;
/* Upload a BootstrappingKey (bootstrapping.rs:92-99) given in the reference's coefficient-form layout (HOST pointers):
 *   ksk_a [N*ks_d][n_s], ksk_b [N*ks_d]       index = digit*N + coefficient        (lwe.rs:108-119)
 *   brk   [n_s][2*rgsw_d][2 (a,b)][N]          RGSW rows                            (rgsw.rs:84-105)
 *   ak    [w+1][rlwe_d][2 (a,b)][N], ak_t[w+1] auto keys for t = -g, g^1..g^w       (bootstrapping.rs:86-89,134-137)
 * brk/ak rows are transformed once to evaluation form (u32) on the device. */
fhe_status fhe_fhew_key_upload(fhe_ctx* ctx, const fhe_fhew_param* param, const uint64_t* ksk_a, const uint64_t* ksk_b,
                               const uint64_t* brk, const uint64_t* ak, const int64_t* ak_t, fhe_fhew_key** out);
/* Key generation ON THE DEVICE: Bootstrapping::key_gen (bootstrapping.rs:122-146; rlwe.rs:109-156,
 * rgsw.rs:84-105, lwe.rs:108-140) for the whole key, directly into the evaluation-form device buffers.  The reference takes
 * its randomness from the caller's RngCore; here it is the counter-based stream of csrc/keygen_stream.cuh (splitmix64 of
 * (seed, domain, index); uniform masks, discrete Gaussian sigma = 3.2 cut at 6 sigma).  z_out [N] / s_out [n_s] receive the
 * RLWE / LWE secrets (int64, host).  The four *_out HOST pointers (reference layout of fhe_fhew_key_upload) are optional: a
 * caller that only wants the key passes NULL and nothing but the two secrets ever leaves the device. */
fhe_status fhe_fhew_keygen(fhe_ctx* ctx, const fhe_fhew_param* param, uint64_t seed, int64_t* z_out, int64_t* s_out, uint64_t* ksk_a_out,
                           uint64_t* ksk_b_out, uint64_t* brk_out, uint64_t* ak_out, fhe_fhew_key** out);
/* Serialised key (the reference has no key format): header {magic "FHEB200K", version 1} | fhe_fhew_param | ak_t[w+1] | the
 * three device images (evaluation-form brk / ak rows, packed ksk).  Loading is three host-to-device copies, no transform.
 * A blob of another version, of inconsistent section sizes or of the wrong length is rejected with FHE_EINVAL. */
size_t fhe_fhew_key_serialized_size(const fhe_fhew_key* key);
fhe_status fhe_fhew_key_serialize(fhe_ctx* ctx, const fhe_fhew_key* key, void* buf, size_t cap);
fhe_status fhe_fhew_key_deserialize(fhe_ctx* ctx, const void* buf, size_t len, fhe_fhew_key** out);
void fhe_fhew_key_free(fhe_ctx* ctx, fhe_fhew_key* key);
/* device bytes held by the key (brk + ak + ksk) and its one-time NCCL broadcast from `root` (every rank must hold a
 * key object of the same parameters, e.g. uploaded from zeros; see fhe_keys_broadcast) */
size_t fhe_fhew_key_bytes(const fhe_fhew_key* key);
fhe_status fhe_fhew_key_broadcast(fhe_ctx* ctx, fhe_fhew_key* key, void* nccl_comm, int root);
/* Bootstrapping::bootstrap (bootstrapping.rs:149-155) on `count` LWE ciphertexts [a_0..a_{N-1}, b] mod Q with test
 * polynomial f (N residues, shared by the batch); out has the same layout.  `post_add` is added to every output
 * body (Fhew::op adds Q/8, fhew.rs:37-38; pass 0 for a bare bootstrap). */
fhe_status fhe_fhew_bootstrap_batch(fhe_ctx* ctx, const fhe_fhew_key* key, const uint64_t* d_f, uint64_t post_add, size_t count,
                                    const uint64_t* d_ct_in, uint64_t* d_ct_out);
fhe_status fhe_fhew_bootstrap_batch_host(fhe_ctx* ctx, const fhe_fhew_key* key, const uint64_t* f, uint64_t post_add, size_t count,
                                         const uint64_t* ct_in, uint64_t* ct_out);
/* The device-pointer entry points above are asynchronous and defer the reference's `unreachable!` check (an even non-zero
 * blind-rotation exponent, bootstrapping.rs:217-222; cannot happen after mod_switch_odd): this call synchronises the context's
 * stream and returns FHE_EINVAL if any bootstrap with this key met one since the flag was last reset (every launch resets it). */
fhe_status fhe_fhew_key_check_error(fhe_ctx* ctx, const fhe_fhew_key* key);
/* Rgsw::internal_product (scheme/fhew/src/rgsw.rs:130-150), the operation Bootstrapping::key_share_merge (bootstrapping.rs:295-320)
 * folds the parties' brk shares with: `count` pairs of RGSW ciphertexts [count][2d rows][2 (a, b)][n] over Z_q (prime, NTT
 * friendly, < 2^62), coefficient form in and out, decomposor (log_b, d); out[c] = internal_product(ct0[c], ct1[c]), every row
 * bit-identical to Rgsw::external_product(ct0[c], row of ct1[c]).  d_out must not alias the inputs. */
fhe_status fhe_rgsw_internal_product(fhe_ctx* ctx, uint64_t q, unsigned log_n, unsigned log_b, unsigned d, size_t count, const uint64_t* d_ct0,
                                     const uint64_t* d_ct1, uint64_t* d_out);
/* first three steps of bootstrap (mod_switch -> Lwe::key_switch -> mod_switch_odd; lwe.rs:90-99,151-160):
 * out [count][n_s+1] residues mod 2N */
fhe_status fhe_fhew_prologue_batch(fhe_ctx* ctx, const fhe_fhew_key* key, size_t count, const uint64_t* d_ct_in, uint64_t* d_out);
/* Lwe::key_switch alone (lwe.rs:151-160): in [count][N+1] mod q_ks -> out [count][n_s+1] mod q_ks */
fhe_status fhe_lwe_key_switch_batch(fhe_ctx* ctx, const fhe_fhew_key* key, size_t count, const uint64_t* d_ct_in, uint64_t* d_out);
/* Rgsw::external_product(brk[j], acc) (rgsw.rs:116-128) and Rlwe::automorphism(ak[v], acc) (rlwe.rs:188-191) on
 * `count` accumulators [a (N), b (N)] in coefficient form; idx[i] selects the key per accumulator. */
fhe_status fhe_fhew_external_product(fhe_ctx* ctx, const fhe_fhew_key* key, size_t count, const uint32_t* d_idx,
                                     const uint64_t* d_acc_in, uint64_t* d_acc_out);
fhe_status fhe_fhew_automorphism(fhe_ctx* ctx, const fhe_fhew_key* key, size_t count, const uint32_t* d_idx,
                                 const uint64_t* d_acc_in, uint64_t* d_acc_out);
/* blind rotation only (bootstrapping.rs:158-209): in [count][n_s+1] mod 2N -> acc [count][2][N] */
fhe_status fhe_fhew_blind_rotate_batch(fhe_ctx* ctx, const fhe_fhew_key* key, const uint64_t* d_f, size_t count,
                                       const uint64_t* d_ct2n, uint64_t* d_acc_out);

/* ---- TFHE programmable bootstrapping (scheme/tfhe/src/{tlwe,tglwe,tggsw,bootstrapping}.rs) ---------------------- */
typedef struct fhe_tfhe_param {
    unsigned log_p, padding;       /* tlwe.rs:19-49                                   */
    unsigned n;                    /* TLWE dimension                                  */
    unsigned ks_log_b, ks_d;       /* TLWE key-switch decomposor (tlwe.rs:29-32)      */
    unsigned log_big_n;            /* ring degree N = 2^log_big_n (tglwe.rs:20-26)    */
    unsigned k;                    /* GLWE dimension                                  */
    unsigned bs_log_b, bs_d;       /* TGGSW decomposor (tggsw.rs:19-32)               */
} fhe_tfhe_param;
typedef struct fhe_tfhe_key fhe_tfhe_key;
/* Upload a BootstrappingKey (tfhe/bootstrapping.rs:40-46), HOST pointers, reference layout:
 *   brk   [n][(k+1)*bs_d][(k+1) polys a_0..a_{k-1}, b][N]   (tggsw.rs:73-89)
 *   ksk_a [(k*N)*ks_d][n], ksk_b [(k*N)*ks_d]                index = digit*(kN) + coefficient (tlwe.rs:100-111)
 * brk polynomials are converted once to the twisted Fourier domain (c64.rs:20-28 + forward FFT). */
fhe_status fhe_tfhe_key_upload(fhe_ctx* ctx, const fhe_tfhe_param* param, const uint64_t* brk, const uint64_t* ksk_a,
                               const uint64_t* ksk_b, fhe_tfhe_key** out);
/* Key generation ON THE DEVICE: Bootstrapping::key_gen (scheme/tfhe/src/bootstrapping.rs:59-76; tggsw.rs:73-89,
 * tglwe.rs:92-103, tlwe.rs:96-111,122-132) for the whole key - n TGGSW encryptions of the TLWE secret bits and the TLWE
 * key-switching key - from the counter-based stream of csrc/keygen_stream.cuh (uniform torus masks; torus noise = an
 * integer-only Irwin-Hall variate of standard deviation std * 2^64; the mask-secret products go through the same f64 FFT
 * product as the reference's `&a * sk`).  z_out [n] / s_out [kN]: the binary secrets (int64, host).  brk_out / ksk_a_out /
 * ksk_b_out (optional, HOST) export the key in the layout of fhe_tfhe_key_upload. */
fhe_status fhe_tfhe_keygen(fhe_ctx* ctx, const fhe_tfhe_param* param, double tlwe_std, double tglwe_std, uint64_t seed, int64_t* z_out, int64_t* s_out,
                           uint64_t* brk_out, uint64_t* ksk_a_out, uint64_t* ksk_b_out, fhe_tfhe_key** out);
/* Serialised key: header {magic "FHEB200K", version 1, kind 2} | fhe_tfhe_param | the device images (bsk in the twisted Fourier
 * domain, merged ksk, its digit-offset column sums, and the fused path's bsk image when the parameters have one).  Loading is
 * four host-to-device copies, no transform; inconsistent or truncated blobs are rejected with FHE_EINVAL. */
size_t fhe_tfhe_key_serialized_size(const fhe_tfhe_key* key);
fhe_status fhe_tfhe_key_serialize(fhe_ctx* ctx, const fhe_tfhe_key* key, void* buf, size_t cap);
fhe_status fhe_tfhe_key_deserialize(fhe_ctx* ctx, const void* buf, size_t len, fhe_tfhe_key** out);
void fhe_tfhe_key_free(fhe_ctx* ctx, fhe_tfhe_key* key);
/* Evaluation mode of every product made with this key.  0 (default): the reference's dataflow - each row * limb product is
 * inverse-transformed and rounded on its own (misc.rs:59-61), raw torus words bit-identical to the reference.  1: products
 * are summed in the Fourier domain and rounded once per output ((k+1) instead of (k+1)^2 d inverse FFTs per CMUX); torus
 * words then differ from the reference's by less than its own error bound (c64.rs:186-208), decryptions are identical.
 * 2: the fused bounded-error blind rotation (k = 1; N = 512, 1024, 2048): same digits and the same exact sums as mode 1, FMA
 * butterflies, twist merged into the forward twiddles, one rounding per output coefficient; each CMUX output is within
 * (k+1) d 2^(64 + log_b + log_n - 53) of the reference's, decryptions are identical (FHE_EUNSUPPORTED for other shapes).
 * 3: mode 2 with the accumulator held as the top 32 bits of every torus word (each CMUX increment is a sum of f64 products of
 * magnitude ~2^90 with nothing but rounding noise below 2^35 - in the reference's own dataflow too - so the extra rounding of
 * at most 2^31 per step is invisible next to it; the low halves of the output words are zero).
 * The stand-alone external product / CMUX entry points evaluate modes 1, 2 and 3 alike. */
fhe_status fhe_tfhe_key_set_mode(fhe_ctx* ctx, fhe_tfhe_key* key, int mode);
/* device bytes held by the key (Fourier-domain bsk + ksk) and its one-time NCCL broadcast from `root` */
size_t fhe_tfhe_key_bytes(const fhe_tfhe_key* key);
fhe_status fhe_tfhe_key_broadcast(fhe_ctx* ctx, fhe_tfhe_key* key, void* nccl_comm, int root);
/* Bootstrapping::bootstrap (tfhe/bootstrapping.rs:78-82) on `count` TLWE ciphertexts [a (n), b]; `d_lut` is the
 * already-encoded test polynomial (N torus words: Tglwe::encode(v), tglwe.rs:80-84), shared by the batch. */
fhe_status fhe_tfhe_pbs_batch(fhe_ctx* ctx, const fhe_tfhe_key* key, const uint64_t* d_lut, size_t count, const uint64_t* d_ct_in,
                              uint64_t* d_ct_out);
fhe_status fhe_tfhe_pbs_batch_host(fhe_ctx* ctx, const fhe_tfhe_key* key, const uint64_t* lut, size_t count, const uint64_t* ct_in,
                                   uint64_t* ct_out);
/* Tggsw::external_product(brk[idx[i]], glwe_i) (tggsw.rs:100-112): glwe [count][k+1][N] */
fhe_status fhe_tfhe_external_product(fhe_ctx* ctx, const fhe_tfhe_key* key, size_t count, const uint32_t* d_idx,
                                     const uint64_t* d_glwe_in, uint64_t* d_glwe_out);
/* Tggsw::cmux(brk[idx[i]], ct0_i, ct1_i) = ct0 + external_product(b, ct1 - ct0) (tggsw.rs:114-121): [count][k+1][N] */
fhe_status fhe_tfhe_cmux(fhe_ctx* ctx, const fhe_tfhe_key* key, size_t count, const uint32_t* d_idx, const uint64_t* d_ct0,
                         const uint64_t* d_ct1, uint64_t* d_out);
/* blind_rotate + sample_extract(0) (tfhe/bootstrapping.rs:84-96, tglwe.rs:115-127): out [count][kN+1] */
fhe_status fhe_tfhe_blind_rotate_extract_batch(fhe_ctx* ctx, const fhe_tfhe_key* key, const uint64_t* d_lut, size_t count,
                                               const uint64_t* d_ct_in, uint64_t* d_out);
/* Tlwe::key_switch (tlwe.rs:144-153): in [count][kN+1] -> out [count][n+1] */
fhe_status fhe_tlwe_key_switch_batch(fhe_ctx* ctx, const fhe_tfhe_key* key, size_t count, const uint64_t* d_ct_in, uint64_t* d_ct_out);

/* ---- CKKS RNS ciphertext arithmetic (scheme/ckks/src/ckks.rs:123-129, 250-293) ---------------------------------- */
typedef struct fhe_ckks_ctx fhe_ckks_ctx;
/* qs / ps as chosen by CkksParam::new (ckks.rs:19-35): big_l ciphertext primes and big_l special primes */
fhe_status fhe_ckks_create(fhe_ctx* ctx, unsigned log_n, const uint64_t* qs, const uint64_t* ps, size_t big_l, fhe_ckks_ctx** out);
void fhe_ckks_destroy(fhe_ctx* ctx, fhe_ckks_ctx* ck);
typedef struct fhe_ckks_ksk fhe_ckks_ksk;
/* Upload a CkksKeySwitchingKey (ckks.rs:154-162) in reference layout (HOST): [2 (b, a)][2*big_l limbs: qs then ps][N],
 * coefficient form; transformed once to evaluation form. */
fhe_status fhe_ckks_ksk_upload(fhe_ctx* ctx, fhe_ckks_ctx* ck, const uint64_t* ksk, fhe_ckks_ksk** out);
void fhe_ckks_ksk_free(fhe_ctx* ctx, fhe_ckks_ksk* ksk);
/* device bytes of the evaluation-form key and its one-time NCCL broadcast from `root` (keys are generated / uploaded on one rank,
 * 16 MiB per key-switching key at N = 2^16, L = 8) */
size_t fhe_ckks_ksk_bytes(const fhe_ckks_ksk* ksk);
/* Key generation ON THE DEVICE: Ckks::sk_gen (ckks.rs:139-141), rlk_gen (164-167) and one automorphism key
 * per exponent of auto_ts (cjk_gen / rtk_gen, 169-184: t = -1 or 5^j mod 2N), each a ksk_gen (154-162) whose masks and errors come
 * from the counter-based stream of csrc/keygen_stream.cuh; the keys stay on the device in evaluation form.  keys_out receives
 * 1 + n_auto handles (relinearisation key first), sk_out [N] the ternary secret (int64, host); export_out is optional (HOST,
 * [1 + n_auto][2 (b, a)][2L][N], the coefficient-form keys in the layout of fhe_ckks_ksk_upload). */
fhe_status fhe_ckks_keygen(fhe_ctx* ctx, fhe_ckks_ctx* ck, uint64_t seed, size_t n_auto, const int64_t* auto_ts, int64_t* sk_out,
                           fhe_ckks_ksk** keys_out, uint64_t* export_out);
/* Serialised key-switching key: header {magic "FHEB200K", version 1, kind 3, log_n, L, bytes} | the evaluation-form device image;
 * rejected with FHE_EINVAL when it does not match the context it is loaded into. */
size_t fhe_ckks_ksk_serialized_size(const fhe_ckks_ksk* ksk);
fhe_status fhe_ckks_ksk_serialize(fhe_ctx* ctx, const fhe_ckks_ctx* ck, const fhe_ckks_ksk* ksk, void* buf, size_t cap);
fhe_status fhe_ckks_ksk_deserialize(fhe_ctx* ctx, const fhe_ckks_ctx* ck, const void* buf, size_t len, fhe_ckks_ksk** out);
fhe_status fhe_ckks_ksk_broadcast(fhe_ctx* ctx, fhe_ckks_ksk* ksk, void* nccl_comm, int root);
/* Ckks::mul = tensor product + relinearize + rescale (ckks.rs:255-272) on `count` pairs at level l (l limbs):
 * ct layout [count][2 (b, a)][l][N] coefficient form (ckks.rs:112-121); out [count][2][l-1][N] */
fhe_status fhe_ckks_mul_relin_rescale_batch(fhe_ctx* ctx, fhe_ckks_ctx* ck, const fhe_ckks_ksk* rlk, size_t level, size_t count,
                                            const uint64_t* d_ct0, const uint64_t* d_ct1, uint64_t* d_out);
fhe_status fhe_ckks_mul_relin_rescale_batch_host(fhe_ctx* ctx, fhe_ckks_ctx* ck, const fhe_ckks_ksk* rlk, size_t level, size_t count,
                                                 const uint64_t* ct0, const uint64_t* ct1, uint64_t* out);
/* Ckks::key_switch (ckks.rs:284-293), optionally preceded by the automorphism X -> X^t (rotate / conjugate,
 * ckks.rs:274-282; t == 0 means none): [count][2][l][N] -> same shape */
fhe_status fhe_ckks_key_switch(fhe_ctx* ctx, fhe_ckks_ctx* ck, const fhe_ckks_ksk* ksk, int64_t t, size_t level, size_t count,
                               const uint64_t* d_ct, uint64_t* d_out);
/* Ckks::mul_constant on an already-encoded plaintext (ckks.rs:250-253): (pt * b, pt * a).rescale().
 * d_pt [pt_count][level][N] coefficient form with pt_count == 1 (shared) or count; out [count][2][level-1][N] */
fhe_status fhe_ckks_mul_plain_rescale_batch(fhe_ctx* ctx, fhe_ckks_ctx* ck, size_t level, size_t count, size_t pt_count,
                                            const uint64_t* d_pt, const uint64_t* d_ct, uint64_t* d_out);
/* One rotation of a BSGS plan: t = 5^j mod 2N (ckks.rs:279-282; 0 = no rotation, key may then be NULL) */
typedef struct fhe_ckks_rot {
    int64_t t;
    const fhe_ckks_ksk* key;
} fhe_ckks_rot;
/* Bootstrapping::mul_mat (scheme/ckks/src/bootstrapping.rs:92-108; the building block of coeff_to_slot / slot_to_coeff):
 *   out = sum_i rot_{giant[i]}( sum_j mul_constant(pt_ij, rot_{baby[j]}(ct)) )
 * with every mul_constant rescaling before the sums, exactly as the reference.  `present` [n_giant][n_baby] marks the (i, j)
 * pairs that carry a diagonal; d_pts holds their encoded plaintexts [number present][level][N] (coefficient form) in
 * row-major (i, j) order.  ct [count][2][level][N] -> out [count][2][level-1][N].  The BSGS plan and the diagonal encoding
 * (misc/matrix.rs, sfft.rs, 256-bit floats) are host-side and not part of this library. */
fhe_status fhe_ckks_mul_mat(fhe_ctx* ctx, fhe_ckks_ctx* ck, size_t level, size_t count, size_t n_baby, const fhe_ckks_rot* baby,
                            size_t n_giant, const fhe_ckks_rot* giant, const uint8_t* present, const uint64_t* d_pts,
                            const uint64_t* d_ct, uint64_t* d_out);
/* CkksCiphertext::rescale (ckks.rs:123-125): [count][2][l][N] -> [count][2][l-1][N] */
fhe_status fhe_ckks_rescale(fhe_ctx* ctx, fhe_ckks_ctx* ck, size_t level, size_t count, const uint64_t* d_ct, uint64_t* d_out);

/* ---- multi-GPU: one-time key distribution (no upstream analogue) --------------------------------- */
/* Broadcast a device buffer from rank `root` to every rank of an already-initialised NCCL communicator
 * (ncclComm_t passed as void*), on the context's stream. */
fhe_status fhe_keys_broadcast(fhe_ctx* ctx, void* nccl_comm, int root, void* d_buf, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* FHE_B200_H */
