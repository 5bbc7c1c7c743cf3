// Entry points declared in include/fhe_b200.h whose kernels are not written yet: they fail loudly with
// FHE_EUNSUPPORTED (never a CPU fallback).  Each function leaves this file when its CUDA implementation lands.
#include "ctx.cuh"

using namespace fhe;
#define PENDING(ctx, what) return (ctx) ? fail((ctx), FHE_EUNSUPPORTED, what " is not implemented yet") : FHE_EINVAL

extern "C" {
fhe_status fhe_fft64_negacyclic_mul(fhe_ctx* ctx, unsigned, size_t, uint64_t*, const uint64_t*) { PENDING(ctx, "fhe_fft64_negacyclic_mul"); }
fhe_status fhe_fft64_negacyclic_mul_host(fhe_ctx* ctx, uint64_t*, const uint64_t*, size_t, size_t) { PENDING(ctx, "fhe_fft64_negacyclic_mul_host"); }
fhe_status fhe_rns_extend_bases(fhe_ctx* ctx, const uint64_t*, size_t, const uint64_t*, size_t, unsigned, size_t, const uint64_t*, uint64_t*) { PENDING(ctx, "fhe_rns_extend_bases"); }
fhe_status fhe_rns_rescale_k(fhe_ctx* ctx, const uint64_t*, size_t, size_t, unsigned, size_t, const uint64_t*, uint64_t*) { PENDING(ctx, "fhe_rns_rescale_k"); }
fhe_status fhe_tfhe_key_upload(fhe_ctx* ctx, const fhe_tfhe_param*, const uint64_t*, const uint64_t*, const uint64_t*, fhe_tfhe_key**) { PENDING(ctx, "fhe_tfhe_key_upload"); }
void fhe_tfhe_key_free(fhe_ctx*, fhe_tfhe_key*) {}
fhe_status fhe_tfhe_pbs_batch(fhe_ctx* ctx, const fhe_tfhe_key*, const uint64_t*, size_t, const uint64_t*, uint64_t*) { PENDING(ctx, "fhe_tfhe_pbs_batch"); }
fhe_status fhe_tfhe_pbs_batch_host(fhe_ctx* ctx, const fhe_tfhe_key*, const uint64_t*, size_t, const uint64_t*, uint64_t*) { PENDING(ctx, "fhe_tfhe_pbs_batch_host"); }
fhe_status fhe_tfhe_external_product(fhe_ctx* ctx, const fhe_tfhe_key*, size_t, const uint32_t*, const uint64_t*, uint64_t*) { PENDING(ctx, "fhe_tfhe_external_product"); }
fhe_status fhe_tfhe_blind_rotate_extract_batch(fhe_ctx* ctx, const fhe_tfhe_key*, const uint64_t*, size_t, const uint64_t*, uint64_t*) { PENDING(ctx, "fhe_tfhe_blind_rotate_extract_batch"); }
fhe_status fhe_tlwe_key_switch_batch(fhe_ctx* ctx, const fhe_tfhe_key*, size_t, const uint64_t*, uint64_t*) { PENDING(ctx, "fhe_tlwe_key_switch_batch"); }
fhe_status fhe_ckks_create(fhe_ctx* ctx, unsigned, const uint64_t*, const uint64_t*, size_t, fhe_ckks_ctx**) { PENDING(ctx, "fhe_ckks_create"); }
void fhe_ckks_destroy(fhe_ctx*, fhe_ckks_ctx*) {}
fhe_status fhe_ckks_ksk_upload(fhe_ctx* ctx, fhe_ckks_ctx*, const uint64_t*, fhe_ckks_ksk**) { PENDING(ctx, "fhe_ckks_ksk_upload"); }
void fhe_ckks_ksk_free(fhe_ctx*, fhe_ckks_ksk*) {}
fhe_status fhe_ckks_mul_relin_rescale_batch(fhe_ctx* ctx, fhe_ckks_ctx*, const fhe_ckks_ksk*, size_t, size_t, const uint64_t*, const uint64_t*, uint64_t*) { PENDING(ctx, "fhe_ckks_mul_relin_rescale_batch"); }
fhe_status fhe_ckks_mul_relin_rescale_batch_host(fhe_ctx* ctx, fhe_ckks_ctx*, const fhe_ckks_ksk*, size_t, size_t, const uint64_t*, const uint64_t*, uint64_t*) { PENDING(ctx, "fhe_ckks_mul_relin_rescale_batch_host"); }
fhe_status fhe_ckks_key_switch(fhe_ctx* ctx, fhe_ckks_ctx*, const fhe_ckks_ksk*, int64_t, size_t, size_t, const uint64_t*, uint64_t*) { PENDING(ctx, "fhe_ckks_key_switch"); }
fhe_status fhe_ckks_rescale(fhe_ctx* ctx, fhe_ckks_ctx*, size_t, size_t, const uint64_t*, uint64_t*) { PENDING(ctx, "fhe_ckks_rescale"); }
}
