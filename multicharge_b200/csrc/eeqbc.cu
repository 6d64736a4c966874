// EEQ-BC 2025 model (model/eeqbc.f90) -- placeholder until the device path lands.
#include "common.cuh"

extern "C" {
int mchrg_b200_new_eeqbc_model(mchrg_b200_context* ctx, int32_t nid, const double* chi, const double* rad,
                               const double* eta, const double* kcnchi, const double* kqchi, const double* kqeta,
                               double kcnrad, const double* cap, const double* avg_cn, const double* rvdw, double kbc,
                               double cutoff, double cn_exp, const double* rcov, const double* en, double norm_exp,
                               mchrg_b200_model** model) {
   mchrg::set_last_error("EEQ-BC is not available in this build");
   return MCHRG_B200_ERR_MODEL;
}
int mchrg_b200_new_eeqbc2025_model(mchrg_b200_context* ctx, int32_t nid, const int32_t* num, const double* rvdw,
                                   const double* en_pauling, mchrg_b200_model** model) {
   mchrg::set_last_error("EEQ-BC is not available in this build");
   return MCHRG_B200_ERR_MODEL;
}
}
