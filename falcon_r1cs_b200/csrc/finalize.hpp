// Host tail of the prover (finalize.cpp).
#pragma once
#include <cstdint>
// msm: the five MSM results as XYZZ points (A | B1 | L | H in G1, 24 u64 each; B2 in G2, 48 u64);
// r, s: Montgomery Fr; proof: A (12) | B (24) | C (12) affine Montgomery.
void host_finalize_proof(const uint64_t* msm, const uint64_t* r_mont, const uint64_t* s_mont, uint64_t* proof);
void host_compress_proof(const uint64_t* proof_affine, uint8_t* out192);
// sums the MSM results of `n_shards` base-range shards (each PROOF_MSM_WORDS = 144 u64, same layout as `msm`
// above) into out144
void host_sum_partials(uint32_t n_shards, const uint64_t* const* partials, uint64_t* out144);
