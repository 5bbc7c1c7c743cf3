// Internal interface between fhew.cu (key object construction) and keygen.cu (device key generation).
#pragma once
#include "ctx.cuh"

// Where the three key buffers of a BootstrappingKey come from.  Exactly one group is set:
//   host_*: the reference's coefficient-form layout on the host (fhe_fhew_key_upload): range-checked, transformed on the device;
//   img_*:  the device images themselves on the host (fhe_fhew_key_deserialize): copied as they are;
//   dev_*:  coefficient-form rows [rows][2 (a, b)][N] (u32 words for Q < 2^30, u64 above; overwritten) and the packed ksk
//           [N ks_d][n_s + 1] u32 already on the device (fhe_fhew_keygen): rows transformed, ksk buffer adopted by the key.
struct FhewKeySource {
    const uint64_t *host_ksk_a = nullptr, *host_ksk_b = nullptr, *host_brk = nullptr, *host_ak = nullptr;
    const void *img_brk = nullptr, *img_ak = nullptr, *img_ksk = nullptr;
    void *dev_brk_rows = nullptr, *dev_ak_rows = nullptr, *dev_ksk = nullptr;
};
fhe_status fhew_key_build(fhe_ctx* ctx, const fhe_fhew_param* pp, const int64_t* ak_t, const FhewKeySource& src, fhe_fhew_key** out);
