// Batched independent molecules (BASELINE.json configs[0..1]) -- placeholder until the fused
// one-CTA-per-molecule kernel lands.
#include "common.cuh"

extern "C" int mchrg_b200_singlepoint_batch(const mchrg_b200_model* model, int32_t nmol, const int64_t* offsets,
                                            const int32_t* num, const double* xyz, const double* charge, double* qvec,
                                            double* energy, double* gradient, int32_t* status) {
   mchrg::set_last_error("singlepoint_batch is not available in this build");
   return MCHRG_B200_ERR_MODEL;
}
