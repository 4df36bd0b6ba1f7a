// internal.h -- shared declarations between the kernel translation unit (kernels.cu) and the C-ABI layer (api.cu).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "materials.cuh"

namespace ad4sm {

enum EvalMode { MODE_ELASTIC = 0, MODE_PF_U = 1, MODE_PF_D = 2, MODE_PHI_R = 3 };

// assembly maps (built once per plan, api.cu) + outputs of one field
struct AsmArgs {
  long long n_nodes, n_blocks;  // node rows, node-pair blocks of the pattern
  int N, DB, dpn;               // nodes/element, block size computed by K1, dofs per node of the assembled field
  const unsigned* blk_row;      // [n_blocks] row node of each block
  const unsigned* nbr_ptr;      // [n_nodes+1] first block of each row node (blocks = sorted neighbour nodes)
  const unsigned* nbr;          // [n_blocks]  neighbour (column) node of each block
  const unsigned* blk_ptr;      // [n_blocks+1] source list of each block
  const unsigned* src;          // [ne*N*N] (slot*NPB + pair index)*2 + transposed, in `elems` order per block
  const unsigned* adj_ptr;      // [n_nodes+1]
  const unsigned* adj;          // [ne*N] slot*N + local node, in `elems` order per node
  const double* Ke;
  const double* re;
  double* nzval;                // [dpn*dpn*n_blocks]
  double* r;                    // [dpn*n_nodes]
  unsigned long long* n_zero;   // [1] stored exact zeros in nzval
  // destination-sorted staging (DESIGN.md "K2s"): contributions of one unordered node pair are contiguous
  long long n_pairs;            // unordered node pairs {i <= p} of the pattern (0: not built)
  const unsigned* cptr;         // [n_pairs+1] contribution range of each pair (in `elems` order inside a pair)
  const unsigned* out_up;       // [n_pairs] nzval offset of block (i, p)
  const unsigned* out_lo;       // [n_pairs] nzval offset of block (p, i)   (== out_up for i == p)
  const unsigned* nn_pack;      // [n_pairs] row lengths: nn(i) | nn(p) << 16
  int k2s_direct;               // K2s with per-thread loads instead of the TMA-staged reads (AD4SM_K2S_DIRECT=1)
  int k2s_exp;                  // TIMING EXPERIMENT ONLY (AD4SM_EXP_K2S): 1 K2s leaves the lower triangle unwritten, 2 writes nothing
  const double* C;              // [ne*NPB][CBLK] contribution blocks, canonical orientation (row node <= col node)
  // peer push (double-buffered HALO rows): adj entries >= halo_adj0 read their residual row halo_adj_shift entries later
  unsigned long long halo_adj0, halo_adj_shift;
  // split assembly of a partitioned run (ad4sm_eval_assemble_split_device): K3 covers the node rows [node0, node1), K2s the
  // pairs [pair0, pair1); keep_zero_count: this launch continues the exact-zero count of the previous one
  long long node0 = 0, node1 = 0, pair0 = 0, pair1 = 0;
  int keep_zero_count = 0;
};
constexpr int CBLK3 = 10;       // doubles per 3x3 contribution block: 9 + 1 pad = 80 B, 16-byte aligned

// One K1 launch = one batch (SoA copy of a Vector{CEElem}/Vector{CPElem}, elements.jl:71-111).
struct K1Args {
  MatDev mat;
  long long ne;         // elements in the batch
  long long slot0;      // first staging slot of the batch
  int D, N, P;          // element dimension, nodes, Gauss points
  int mf;               // MatFamily
  int dpn;              // rows of u (dofs per node of the displacement field)
  int want_K;           // 0: ϕ and r only (makeϕr)
  int cas;              // 1: axisymmetric CAS batch (D = 2; Nval = N / X0, the hoop row; 3-D material on a 3x3 F)
  const int* nodes;     // [ne][N] 0-based
  const double* gradN;  // [ne][D][P][N]
  const double* wgt;    // [ne][P]
  const double* Nval;   // [ne][P][N] or nullptr
  const double* u;      // [n_nodes][dpn]
  const double* dfield; // [n_nodes] or nullptr
  double* hist;         // [ne][P][2] (this batch) or nullptr
  // staging written by K1, consumed by K2/K3
  double* Ke;           // [slots][NPB][DB*DB]  cyclic symmetric block storage (element.cuh)
  double* re;           // [slots][N][DB]
  double* phie;         // [slots]
  // destination-sorted staging: cpos[slot*NPB + block] = position << 1 | transposed; nullptr = element-major staging
  const unsigned* cpos;
  double* C;
  int exp;              // TIMING EXPERIMENT ONLY (AD4SM_EXP_K1, wrong results): 1 = the sorted epilogue stores nothing
  int sorted_tma;       // sorted blocks leave through shared memory + one 80-byte bulk copy each (else plain STG.128)
  // peer staging: slots [first, first+count) are also written to the owner rank's staging over NVLink
  int n_xr;
  struct PeerRange {
    long long first, count;
    double* ke;  // peer Ke + remote_first_slot * width, i.e. indexed by (slot - first)
    double* re;
  } xr[4];
  // general export map (in-library partitioned evaluation, ad4sm_dist_attach): own slot -> row of the element-major export
  // buffers xke / xre, or -1; nullptr = nothing to export this way
  const int* xmap;
  double* xke;
  double* xre;
};

// launch shape of the persistent K1 for a batch
struct K1Config {
  int persistent;  // 0: thread-per-element kernel (one Gauss point)
  int EW;          // elements per warp iteration
  int warps;       // warps per CTA
  int grid;        // CTAs
  int staged;      // blocks/residual leave through the shared-memory staging + bulk stores
};
bool k1_config(int D, int N, int P, bool pf, long long ne, K1Config* c);

cudaError_t launch_k1(const K1Args& a, int mode, cudaStream_t s);
const char* k1_unsupported_reason(int D, int N, int P, int mf, int mode);
cudaError_t launch_k2(const AsmArgs& a, cudaStream_t s);  // Kt values
cudaError_t launch_k3(const AsmArgs& a, cudaStream_t s);  // residual
cudaError_t launch_k2s(const AsmArgs& a, cudaStream_t s); // Kt values from the destination-sorted staging
cudaError_t launch_halo_scatter(const double* Ke, const unsigned* cpos, double* C, long long slot0, long long n_slots, int NPB,
                                cudaStream_t s);
int phi_reduce_partials();
cudaError_t launch_phi_reduce(const double* phie, long long n, double* partials, double* phi, cudaStream_t s);
cudaError_t launch_pattern(const AsmArgs& a, long long* colptr, long long* rowval, cudaStream_t s);
// exact-zero compaction (dropzeros, elements.toolkit.jl:202); tmp must hold n_dof+1 long longs
// assembly-only entry: element duals [ne][stride] = (v | g[n] | h[n(n+1)/2] packed upper triangle, column-major) -> the
// element-major staging (Ke blocks, re, phie); `index`: record of each slot (nullptr = identity)
cudaError_t launch_duals_to_staging(const double* Phi, long long stride, const long long* index, long long n_slots, int N, int D,
                                    double* Ke, double* re, double* phie, cudaStream_t s);
cudaError_t launch_dropzeros(long long n_dof, const long long* colptr, const long long* rowval, const double* nzval,
                             long long* colptr_out, long long* rowval_out, double* nzval_out, void* cub_tmp,
                             size_t cub_tmp_bytes, cudaStream_t s);
size_t dropzeros_tmp_bytes(long long n_dof);
cudaError_t launch_fp64_peak(double* tflops, double* mhz);

// ---- batched D1 post-processing (post.cu; elements.toolkit.jl:8-132) ---------------------------------------------------
struct PostArgs {
  MatDev mat;
  long long ne, slot0;
  int D, N, P, mf, dpn, want_stress;
  const int* nodes;
  const double* gradN;
  const double* wgt;
  const double* u;
  const long long* index;  // position of each slot in `elems` (nullptr: slot0 + e)
  double *F, *detJ, *vol, *sigma, *Pk;  // outputs in `elems` order; any may be nullptr
};
cudaError_t launch_post(const PostArgs& a, cudaStream_t s);

// ---- batched element constructors (construct.cu; elasticelements.jl:35-197, phasefieldelements.jl:9-145) ------------------
enum ConstructShape { CS_TRIA3 = 1, CS_QUAD4 = 2, CS_TET4 = 3, CS_HEX8 = 4, CS_WDG6 = 5 };
struct ConstructArgs {
  int shape, P, ldc;
  int n_wfac;               // weight factors per Gauss point (tensor rules: D, wedge: 1)
  int n1d;                  // tensor rules: points per direction (fixes the order in which V is summed); 0 otherwise
  long long ne;
  const long long* nodes;   // device [ne][N], 1-based
  const double* coords;     // device [n_nodes][ldc]
  double pt[27][3];         // Gauss points, storage order (ξ index fastest)
  double wfac[27][3];       // wgt = detJ * wfac[0] * wfac[1] * ..., left to right
  double *gradN, *wgt, *Nval, *V;  // device outputs; Nval, V may be nullptr
};
cudaError_t launch_construct(const ConstructArgs& a, cudaStream_t s);

// ---- constraint-equation catalogue (consteq.cu; solvers.jl:17-22, 54-73) ----------------------------------------------------
struct ConstEqArgs {
  long long n_eqs;
  const long long* dof_ptr;  // [n_eqs+1]
  const long long* dofs;     // 1-based linear indices into u[:]
  const double* c;           // linear coefficients, aligned with dofs
  const double* c0;          // [n_eqs]
  const long long* q_ptr;    // [n_eqs+1] offsets into Q (an equation without Q has an empty range), or nullptr
  const double* Q;           // per equation n x n, column-major, symmetric
  const double* u;           // [n_dof]
  const double* lambda;      // [n_eqs] or nullptr
  double *veqs, *grad, *hess;
};
cudaError_t launch_consteq(const ConstEqArgs& a, cudaStream_t s);

// ---- NCCL plumbing of the partitioned evaluation (dist.cu) ---------------------------------------------------------
struct DistComm;
struct DistXfer {
  int send, peer;   // send != 0: ncclSend, else ncclRecv
  double* buf;
  size_t count;     // doubles
};
}  // namespace ad4sm
#include <string>
namespace ad4sm {
int dist_unique_id(void* out, std::string* err);
DistComm* dist_comm_init(int nranks, int rank, const void* id_bytes, int device, std::string* err);
void dist_comm_destroy(DistComm* c);
int dist_comm_size(const DistComm* c);
int dist_comm_rank(const DistComm* c);
int dist_exchange(DistComm* c, int n, const DistXfer* x, cudaStream_t s, std::string* err);
int dist_allgather(DistComm* c, const double* src, double* dst, size_t count, cudaStream_t s, std::string* err);
cudaError_t launch_gather_rows(const double* src, const long long* rows, long long n_rows, int width, double* dst, cudaStream_t s);
cudaError_t launch_phi_rank_sum(const double* allphi, int n, double* out, cudaStream_t s);

// ---- Newton-step glue (newton.cu; solvers.jl:183-184, 211-224, 257-291) ----------------------------------------------
struct NewtonArgs {
  long long n_nodes, nfree;
  int dpn;
  const unsigned* nbr_ptr;      // node-block pattern of the field (AsmArgs)
  const unsigned* nbr;
  const int* fidx;              // [n_dof] >= 0: rank among the free dofs; < 0: -1 - rank among the constrained dofs
  const int* free_list;         // [nfree] dof id of each free dof (ascending = findall(bfree[:]))
  const double* nzval;          // structural CSC values of the last evaluation
  const long long* colptr_ff;   // [nfree+1] 1-based column pointers of Kt[ifree,ifree] (nullptr in the counting pass)
  long long* cnt;               // counting pass: [nfree] free rows per column
  long long* rowval_ff;         // pattern pass
  double* nzval_ff;             // value pass
  const double* du;             // [ncnst] Δu of the constrained dofs, or nullptr
  double* fc;                   // [nfree] (Kt[ifree,icnst]*Δucnst)
};
int newton_partials();
cudaError_t launch_newton_columns(const NewtonArgs& a, cudaStream_t s);
size_t newton_scan_tmp_bytes(long long n);
cudaError_t launch_newton_colptr(long long* cnt, long long nfree, void* tmp, size_t tmp_bytes, cudaStream_t s);
cudaError_t launch_newton_res(long long nfree, const int* free_list, const double* r, const double* fe, const double* fc, int kind,
                              double* res, long long ncnst, const int* cnst_list, double* partials, double* norms, cudaStream_t s);

}  // namespace ad4sm
