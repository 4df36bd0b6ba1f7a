// internal.h -- shared declarations between the kernel translation units and the C-ABI layer.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "materials.cuh"

namespace ad4sm {

// device view of one batch (SoA copy of a Vector{CEElem}/Vector{CPElem}, elements.jl:71-111)
struct BatchDev {
  int kind, D, P, N;
  int mf;  // MatFamily
  int pad;
  MatDev mat;
  long long ne;
  long long s0;          // first staging slot of this batch
  const int* nodes;      // [ne][N] 0-based
  const double* gradN;   // [ne][D][P][N]
  const double* wgt;     // [ne][P]
  const double* Nval;    // [ne][P][N] or nullptr
};

enum EvalMode { MODE_ELASTIC = 0, MODE_PF_U = 1, MODE_PF_D = 2, MODE_PHI_R = 3 };

// staging written by K1, consumed by K2
struct Staging {
  double* Ke;    // [slots][N][N][DB][DB]   full element tangent, block-row major
  double* re;    // [slots][N][DB]
  double* phie;  // [slots]
};

// assembly maps (built once per plan)
struct AsmMaps {
  long long n_nodes;
  int N;                       // nodes per element
  int max_nbr;                 // max neighbours of a node (incl. itself)
  const long long* adj_ptr;    // [n_nodes+1]
  const unsigned* adj;         // [n_adj]  slot*N + local node, in `elems` order per node
  const unsigned short* pos;   // [n_adj][N] position of each element node in the row node's neighbour list
  const long long* nbr_ptr;    // [n_nodes+1]
  const int* nbr;              // [n_nbr] sorted neighbour node ids (0-based)
};

// ---- launchers (kernels.cu) ----
cudaError_t launch_k1(const BatchDev& b, int mode, const double* u, int dpn, const double* dfield, double* hist,
                      const Staging& st, bool want_K, cudaStream_t s);
cudaError_t launch_k2(const AsmMaps& m, int DB, int dpn, const Staging& st, bool want_K, double* nzval, double* r,
                      unsigned long long* n_zero, cudaStream_t s);
cudaError_t launch_phi_reduce(const double* phie, long long n, double* partials, int n_partials, double* phi,
                              cudaStream_t s);
cudaError_t launch_pattern(const AsmMaps& m, int dpn, long long* colptr, long long* rowval, cudaStream_t s);
cudaError_t launch_dropzeros(long long n_dof, long long nnz, const long long* colptr, const long long* rowval,
                             const double* nzval, long long* colptr_out, long long* rowval_out, double* nzval_out,
                             long long* nnz_out_host, cudaStream_t s);
cudaError_t launch_fp64_peak(double* tflops, double* mhz);
int phi_reduce_partials();

}  // namespace ad4sm
