// Model object (class(mchrg_model_type)) and the per-call device view of a structure.
#pragma once
#include <vector>

#include "common.cuh"

namespace mchrg {

// largest molecule the one-CTA-per-molecule kernels of batch.cu hold (batch::NP_MAX)
constexpr int BATCH_MAX_ATOMS = 208;

// mctc_ncoord parameters (cn_count%erf / erf_en), SURVEY.md 3.4
struct NcoordPar {
   double kcn = 7.5;
   double norm_exp = 1.0;
   double cutoff = 25.0;
   double cut = -1.0;  // log-cut (cn_max), <= 0: none
};

// per-atom parameters gathered from the per-species model arrays (one small kernel per call)
struct AtomPar {
   const double* rad = nullptr;
   const double* eta = nullptr;
   const double* chi = nullptr;
   const double* kcnchi = nullptr;
   const double* rcov = nullptr;
   // EEQ-BC
   const double* kqchi = nullptr;
   const double* kqeta = nullptr;
   const double* cap = nullptr;
   const double* avg_cn = nullptr;
   const double* en = nullptr;
   const double* rcov_en = nullptr;
};

// device view of type(structure_type)
struct DevMol {
   int nat = 0;
   const int* id = nullptr;      // 1-based species
   const double* xyz = nullptr;  // [3*nat]
   double charge = 0.0;
   bool periodic = false;
   double lattice[9] = {0};
};

}  // namespace mchrg

struct mchrg_b200_model {
   mchrg_b200_context* ctx = nullptr;
   int kind = 1;  // 1 = EEQ, 2 = EEQ-BC
   int nid = 0;
   // host copies (model/type.F90:42-63) and device copies, per species
   std::vector<double> chi, rad, eta, kcnchi, rcov;
   std::vector<double> kqchi, kqeta, cap, avg_cn, rvdw, en, rcov_bc;
   double kcnrad = 0.0, kbc = 0.0, norm_exp = 1.0;
   mchrg::NcoordPar ncoord, ncoord_en;
   // device: [chi | rad | eta | kcnchi | rcov | kqchi | kqeta | cap | avg_cn | en] each nid, then rvdw nid*nid
   double* d_par = nullptr;
   const double* dev(int slot) const { return d_par + (size_t)slot * nid; }
   const double* dev_rvdw() const { return d_par + (size_t)10 * nid; }
};
