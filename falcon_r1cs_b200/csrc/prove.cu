// TEMPORARY stubs; replaced as the subsystems land.
#include "ctx.hpp"
#define NOTIMPL(name) { frcs_set_error(name ": not implemented yet"); return FRCS_E_INVALID_ARG; }
extern "C" {
int32_t frcs_load_pk(frcs_ctx*, const frcs_pk_view*) NOTIMPL("frcs_load_pk")
int32_t frcs_prove_batch(frcs_ctx*, uint64_t, const uint16_t*, const uint16_t*, const uint16_t*, const uint64_t*, const uint64_t*, uint64_t*, int32_t*) NOTIMPL("frcs_prove_batch")
int32_t frcs_prove_from_z(frcs_ctx*, uint64_t, const uint64_t*, const uint64_t*, const uint64_t*, uint64_t*) NOTIMPL("frcs_prove_from_z")
int32_t frcs_prove_batch_dev(frcs_ctx*, uint64_t, const uint16_t*, const uint16_t*, const uint16_t*, const uint64_t*, const uint64_t*, uint64_t*, int32_t*, void*) NOTIMPL("frcs_prove_batch_dev")
int32_t frcs_proof_compress(const uint64_t*, uint8_t*) NOTIMPL("frcs_proof_compress")
int32_t frcs_imad_peak(frcs_ctx*, double*) NOTIMPL("frcs_imad_peak")
}
