// api.cu -- C ABI of libad4sm_b200.so (include/ad4sm_b200.h): plan construction (SoA upload, structural
// pattern, element -> block source maps) and the evaluation entry points that replace the reference's
//   makeϕrKt(elems, u)            /root/reference/src/elasticelements.jl:356-358  (-> _assemble_ϕrKt :208-221)
//   makeϕrKt(Φ, elems, u)         /root/reference/src/elements.toolkit.jl:169-203
//   makeϕr                        /root/reference/src/elements.toolkit.jl:204-215
//   makeϕrKt(elems, u, d)         /root/reference/src/phasefieldelements.jl:228-238
//   makeϕrKt_d(elems, u, d[,ϕmax])/root/reference/src/phasefieldelements.jl:239-262
// There is no CPU fallback: every compute entry point needs a CUDA device.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <string>
#include <vector>

#include "../../include/ad4sm_b200.h"
#include "element.cuh"
#include "internal.h"

using namespace ad4sm;

static thread_local std::string g_err;
static int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}
#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e_ = (call);                                                                         \
    if (e_ != cudaSuccess)                                                                           \
      return fail(e_ == cudaErrorMemoryAllocation ? AD4SM_ENOMEM : AD4SM_ENODEV,                     \
                  std::string(#call) + ": " + cudaGetErrorString(e_));                               \
  } while (0)

namespace {

struct BatchPlan {
  K1Args k;            // device pointers + shape (u/dfield/hist/want_K are filled per evaluation)
  long long hist_off;  // offset (in Gauss points) of the batch in the concatenated history array
  bool halo;           // AD4SM_BATCH_HALO: staging filled by the exchange step, not by K1
};

struct FieldOut {
  int dpn = 0;
  long long n_dof = 0, nnz = 0;
  double* nzval = nullptr;
  double* r = nullptr;
  long long* colptr = nullptr;  // structural pattern, built on first use
  long long* rowval = nullptr;
  bool evaluated = false, have_K = false;
  long long n_zero = 0;  // host copy (valid after a host-API evaluation / fetch)
  // dropzeros result of the last evaluation (built on demand)
  long long *dz_colptr = nullptr, *dz_rowval = nullptr;
  double* dz_nzval = nullptr;
  long long dz_cap = 0;
  bool dz_valid = false;
  // Newton-step glue (ad4sm_newton_*): free / constrained split of the field's dofs and the pattern of Kt[ifree,ifree]
  struct Newton {
    bool ready = false;
    long long nfree = 0, ncnst = 0, nnz_ff = 0;
    int *fidx = nullptr, *free_list = nullptr, *cnst_list = nullptr;
    long long *colptr_ff = nullptr, *rowval_ff = nullptr;
    double *nzval_ff = nullptr, *res = nullptr, *fc = nullptr, *fe = nullptr, *du = nullptr, *partials = nullptr, *norms = nullptr;
    bool reduced = false;
  } nw;
};

struct EvSet {
  cudaEvent_t e[5];  // [0] start, [1] K1 done, [4] assembly start (after any exchange), [2] K2/K3 done, [3] ϕ done
};

}  // namespace

struct ad4sm_plan {
  int device = 0;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaStream_t copy_stream = nullptr;          // host API: ϕ and r travel back while K2 still runs
  cudaEvent_t ev_r_ready = nullptr, ev_copied = nullptr;
  // host API, pipelined upload: u (and d) travel in node chunks on the copy stream and K1 starts on the element
  // slice whose nodes have arrived, instead of waiting for the whole array (0.45 ms of PCIe for 1 M Hex8)
  static constexpr int H2D_MAX = 4;
  int n_h2d = 0;                  // 0: one copy, one launch
  long long h2d_elem[H2D_MAX + 1] = {0, 0, 0, 0, 0};  // element slices of the (single) batch
  long long h2d_node[H2D_MAX + 1] = {0, 0, 0, 0, 0};  // slice s needs the nodes below h2d_node[s + 1]
  cudaEvent_t ev_h2d[H2D_MAX] = {nullptr, nullptr, nullptr, nullptr}, ev_start = nullptr;
  bool h2d_active = false;        // set by eval_host around the evaluation
  long long* key_dev = nullptr;   // [ne] position of each slot in `elems` (orig_index), nullptr = identity (assembly-only entry)
  // peer push (AD4SM_OPT_PEER_PUSH): export ranges leave by copy-engine P2P copies into double-buffered HALO rows
  bool push_opt = false;          // option set: the next ad4sm_peer_attach builds push links
  bool push_mode = false;         // links attached in push mode (a rank may have none of its own and still receive)
  int n_push = 0;
  struct PushLink {
    long long first, count;       // local slots
    long long remote[2];          // first row in the peer's staging, per receive buffer
    double *ke, *re;              // peer staging bases (IPC-mapped)
  } push[AD4SM_MAX_PEER_LINKS];
  long long push_end = 0;         // elements [0, push_end) cover every pushed range: K1 runs them first
  int parity = 0;                 // receive buffer of the current evaluation (same on every rank: they evaluate in lockstep)
  bool push_pending = false;      // the element stage of the current evaluation pushed: its assembly reads buffer `parity`
  long long n_halo = 0;           // HALO slots = rows of one receive buffer
  cudaEvent_t ev_exports = nullptr, ev_pushed = nullptr;
  int N = 0, D = 0, dpn = 0;
  long long n_nodes = 0, ne = 0, n_blocks = 0, n_gp = 0;
  bool has_cp = false;
  std::vector<BatchPlan> batches;
  std::vector<void*> allocs;  // every device allocation, for destroy
  unsigned *blk_row = nullptr, *nbr_ptr = nullptr, *nbr = nullptr, *blk_ptr = nullptr, *src = nullptr, *adj_ptr = nullptr,
           *adj = nullptr;
  double *Ke = nullptr, *re = nullptr, *phie = nullptr, *partials = nullptr, *phi_dev = nullptr;
  unsigned long long* n_zero_dev = nullptr;
  double *u_dev = nullptr, *d_dev = nullptr, *hist_dev = nullptr;
  FieldOut field[2];
  int last_field = -1;
  // profiling
  bool prof = false;
  std::vector<EvSet> ev_pool;
  size_t ev_used = 0;
  double prof_ms[AD4SM_PROF_NCLASS] = {0, 0, 0, 0};
  long long prof_launches[AD4SM_PROF_NCLASS] = {0, 0, 0, 0};
  EvSet* ev_cur = nullptr;  // events of the evaluation in flight between the elements and assemble stages
  int elements_mode = -1;
  // destination-sorted staging (3-D elements, dpn == 3, evaluated batches before the HALO batches)
  bool sorted_built = false, sorted_on = true, sorted_last = false;
  long long n_pairs = 0;
  unsigned *cpos = nullptr, *cptr = nullptr, *out_up = nullptr, *out_lo = nullptr, *nn_pack = nullptr;
  double* C = nullptr;
  long long n_own_sorted = 0;  // slots [n_own_sorted, ne) are HALO: scattered into C before K2s
  // split assembly (ad4sm_eval_assemble_split_device): node rows < split_node -- and the pairs numbered under them,
  // [0, split_pair) -- receive no HALO contribution and can be assembled before the interface exchange has completed
  long long split_node = 0, split_pair = 0;
  int split_done = 0;          // 1: the interior part of the pending evaluation has been assembled
  bool halo_prepared = false;  // the HALO scatter and the interface rows of r have been done on the exchange stream
  int n_own_batches = 0;       // evaluated (non-HALO) batches; the partitioned-run features (export ranges, peer links) know one
  bool dist_aware = false;     // the caller declared its export ranges / peer links (ad4sm_peer_attach / ad4sm_export_ranges)
  // peer staging / export ranges
  int n_xr = 0, n_xr_local = 0;
  K1Args::PeerRange xr[AD4SM_MAX_PEER_LINKS] = {};
  std::vector<void*> ipc_open;
  // in-library partitioned evaluation (ad4sm_dist_attach)
  struct DistPeer {
    int rank = 0;
    long long n_send = 0, recv_first = 0, n_recv = 0;
    long long *rows_x = nullptr, *rows_s = nullptr;  // device: rows in the export buffers / staging slots, in the peer's HALO order
    long long first_x = -1, first_s = -1;            // >= 0: the list is the contiguous range starting there (sent in place)
    double *ske = nullptr, *sre = nullptr;           // packed send buffers (non-contiguous lists)
  };
  ad4sm_dist* dcomm = nullptr;
  std::vector<DistPeer> dpeers;
  std::vector<cudaEvent_t> dprof;  // developer tool (AD4SM_DIST_PROF): 4 timing events per partitioned evaluation
  int* xmap = nullptr;
  double *xke = nullptr, *xre = nullptr, *allphi = nullptr, *phi_local = nullptr;
  long long n_export = 0;
};

struct ad4sm_dist {
  ad4sm::DistComm* c;
  int device;
};

template <class T> static int dev_alloc(ad4sm_plan* p, T** ptr, size_t count) {
  *ptr = nullptr;
  // +64 bytes: the TMA bulk copies of K1 round the tail group's wgt / node-id / N rows up to 16 bytes
  CK(cudaMalloc((void**)ptr, count * sizeof(T) + 64));
  p->allocs.push_back((void*)*ptr);
  return 0;
}
template <class T> static int dev_upload(ad4sm_plan* p, T** ptr, const T* host, size_t count) {
  int rc = dev_alloc(p, ptr, count);
  if (rc) return rc;
  if (count) CK(cudaMemcpy(*ptr, host, count * sizeof(T), cudaMemcpyHostToDevice));
  return 0;
}

static int mat_family(const ad4sm_batch& b, int* mf, std::string* why) {
  const int kind = b.mat_kind;
  if (b.elem_kind == AD4SM_ELEM_CAS) {  // 3x3 F (in-plane block + hoop stretch): the 3-D elastic materials
    switch (kind) {
      case AD4SM_MAT_HOOKE: *mf = MF_HOOKE; return 0;
      case AD4SM_MAT_NEOHOOKE: *mf = MF_NEO; return 0;
      case AD4SM_MAT_MOONEYRIVLIN: *mf = MF_MOONEY; return 0;
      case AD4SM_MAT_OGDEN: *mf = MF_OGDEN; return 0;
    }
    *why = "CAS (axisymmetric) elements evaluate a 3x3 F: the material must be Hooke, NeoHooke, MooneyRivlin or Ogden";
    return -1;
  }
  if (kind == AD4SM_MAT_PHASEFIELD) {
    const int base = b.mat_base;
    const bool dhn = b.mat_flags & AD4SM_F_DHN;
    if (base == AD4SM_MAT_HOOKE || base == AD4SM_MAT_HOOKE2D) {
      *mf = MF_HOOKE;
      return 0;
    }
    if (base == AD4SM_MAT_NEOHOOKE && !dhn) {
      if (b.dim == 2 && !(b.mat_flags & AD4SM_F_NH2D_EXT)) {
        *why = "PhaseField{NeoHooke} indexes F[5..9]: BoundsError for 2x2 F in the reference "
               "(phasefieldmaterials.jl:42-47); set AD4SM_F_NH2D_EXT for the plane-strain embedding";
        return -1;
      }
      *mf = MF_PFNEO;
      return 0;
    }
    *why = "PhaseField base material must be Hooke/Hooke2D (ATn, DHn) or NeoHooke (ATn): the reference's generic "
           "methods read mat.mat.E/ν (phasefieldmaterials.jl:20-36,56-72)";
    return -1;
  }
  switch (kind) {
    case AD4SM_MAT_HOOKE:
      if (b.dim != 3) {
        *why = "Hooke is 3-D only (indexes E[5], E[9], elasticmaterials.jl:214-216); use Hooke2D";
        return -1;
      }
      *mf = MF_HOOKE;
      return 0;
    case AD4SM_MAT_HOOKE2D:
      if (b.dim != 2) {
        *why = "Hooke2D needs a 2-D element";
        return -1;
      }
      *mf = MF_HOOKE;
      return 0;
    case AD4SM_MAT_NEOHOOKE: *mf = MF_NEO; return 0;
    case AD4SM_MAT_MOONEYRIVLIN: *mf = MF_MOONEY; return 0;
    case AD4SM_MAT_OGDEN: *mf = MF_OGDEN; return 0;
  }
  *why = "unknown material kind";
  return -1;
}


// ---- batched constructors: ad4sm_mesh -> kernel arguments (Gauss points in the reference's storage order) --------------
static int shape_dims(int shape, int n1d, int* D, int* N, int* P) {
  switch (shape) {
    case AD4SM_SHAPE_TRIA3: *D = 2, *N = 3, *P = 1; return 0;
    case AD4SM_SHAPE_TET4: *D = 3, *N = 4, *P = 1; return 0;
    case AD4SM_SHAPE_WDG6: *D = 3, *N = 6, *P = 6; return 0;
    case AD4SM_SHAPE_QUAD4:
      if (n1d < 1 || n1d > 3) return -1;
      *D = 2, *N = 4, *P = n1d * n1d;
      return 0;
    case AD4SM_SHAPE_HEX8:
      if (n1d < 1 || n1d > 3) return -1;
      *D = 3, *N = 8, *P = n1d * n1d * n1d;
      return 0;
  }
  return -1;
}
static int construct_args(const ad4sm_mesh* m, ConstructArgs* a) {
  if (!m) return fail(AD4SM_EINVAL, "mesh is NULL");
  int D, N, P;
  if (shape_dims(m->shape, m->n_gp_1d, &D, &N, &P))
    return fail(AD4SM_EINVAL, "mesh: unknown shape, or a tensor rule with other than 1..3 points per direction");
  if (m->coord_dim < D) return fail(AD4SM_EINVAL, "mesh: coord_dim < element dimension");
  if (!m->coords || m->n_nodes <= 0) return fail(AD4SM_EINVAL, "mesh: coords missing");
  memset(a, 0, sizeof(*a));
  a->shape = m->shape;
  a->P = P;
  a->ldc = m->coord_dim;
  if (m->shape == AD4SM_SHAPE_QUAD4 || m->shape == AD4SM_SHAPE_HEX8) {
    const int n = m->n_gp_1d;
    a->n1d = n;
    a->n_wfac = D;
    for (int g = 0; g < P; g++) {  // storage index g = ii + n (jj + n kk): Nx[ii,jj,kk] flattened column-major
      const int idx[3] = {g % n, (g / n) % n, g / (n * n)};
      for (int k = 0; k < D; k++) {
        a->pt[g][k] = m->gp_xi[idx[k]];
        a->wfac[g][k] = m->gp_w[idx[k]];
      }
    }
  } else if (m->shape == AD4SM_SHAPE_WDG6) {  // elasticelements.jl:169-171
    const double s3 = sqrt(3.0) / 3.0;
    const double xy[3][2] = {{2.0 / 3.0, 1.0 / 6.0}, {1.0 / 6.0, 2.0 / 3.0}, {1.0 / 6.0, 1.0 / 6.0}};
    a->n_wfac = 1;
    for (int g = 0; g < 6; g++) {
      a->pt[g][0] = xy[g / 2][0];
      a->pt[g][1] = xy[g / 2][1];
      a->pt[g][2] = (g % 2) ? -s3 : s3;
      a->wfac[g][0] = 1.0 / 3.0;
    }
  }
  return 0;
}
// nodes / coords: host; outputs: device.  Temporary device copies of the inputs live for the duration of the call.
static int construct_run(const ad4sm_mesh* m, long long ne, const int64_t* nodes, double* gradN, double* wgt, double* Nval, double* V,
                         cudaStream_t s) {
  ConstructArgs a;
  int rc = construct_args(m, &a);
  if (rc) return rc;
  int D, N, P;
  shape_dims(m->shape, m->n_gp_1d, &D, &N, &P);
  for (long long i = 0; i < ne * N; i++)
    if (nodes[i] < 1 || nodes[i] > m->n_nodes) return fail(AD4SM_EINVAL, "node id out of range 1..n_nodes");
  long long* nd = nullptr;
  double* xc = nullptr;
  cudaError_t e = cudaMalloc((void**)&nd, sizeof(long long) * (size_t)ne * N);
  if (e == cudaSuccess) e = cudaMalloc((void**)&xc, sizeof(double) * (size_t)m->n_nodes * m->coord_dim);
  if (e == cudaSuccess) e = cudaMemcpyAsync(nd, nodes, sizeof(long long) * (size_t)ne * N, cudaMemcpyHostToDevice, s);
  if (e == cudaSuccess) e = cudaMemcpyAsync(xc, m->coords, sizeof(double) * (size_t)m->n_nodes * m->coord_dim, cudaMemcpyHostToDevice, s);
  a.ne = ne;
  a.nodes = nd;
  a.coords = xc;
  a.gradN = gradN;
  a.wgt = wgt;
  a.Nval = Nval;
  a.V = V;
  if (e == cudaSuccess) e = launch_construct(a, s);
  if (e == cudaSuccess) e = cudaStreamSynchronize(s);
  cudaFree(nd);
  cudaFree(xc);
  if (e != cudaSuccess)
    return fail(e == cudaErrorMemoryAllocation ? AD4SM_ENOMEM : AD4SM_ENODEV, std::string("construct: ") + cudaGetErrorString(e));
  return 0;
}

extern "C" {

int ad4sm_abi_version(void) { return AD4SM_ABI_VERSION; }
const char* ad4sm_last_error(void) { return g_err.c_str(); }

int ad4sm_device_count(int* count) {
  if (!count) return fail(AD4SM_EINVAL, "count is NULL");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *count = 0;
    return fail(AD4SM_ENODEV, std::string("cudaGetDeviceCount: ") + cudaGetErrorString(e));
  }
  *count = n;
  return AD4SM_OK;
}

static int select_device(int32_t device) {
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(AD4SM_ENODEV, "no CUDA device: libad4sm_b200 has no CPU fallback");
  int dev = device;
  if (dev < 0) CK(cudaGetDevice(&dev));
  if (dev >= ndev) return fail(AD4SM_EINVAL, "device ordinal out of range");
  CK(cudaSetDevice(dev));
  return 0;
}

int ad4sm_construct_device(const ad4sm_mesh* mesh, int64_t n_elems, const int64_t* nodes, int32_t device, double* gradN_dev,
                           double* wgt_dev, double* Nval_dev, double* V_dev) {
  if (!mesh || n_elems < 0 || (n_elems > 0 && (!nodes || !gradN_dev || !wgt_dev))) return fail(AD4SM_EINVAL, "construct: arguments missing");
  int rc = select_device(device);
  if (rc) return rc;
  return construct_run(mesh, n_elems, nodes, gradN_dev, wgt_dev, Nval_dev, V_dev, (cudaStream_t)0);
}

int ad4sm_construct(const ad4sm_mesh* mesh, int64_t n_elems, const int64_t* nodes, int32_t device, double* gradN, double* wgt,
                    double* Nval, double* V) {
  if (!mesh || n_elems < 0 || (n_elems > 0 && (!nodes || !gradN || !wgt))) return fail(AD4SM_EINVAL, "construct: arguments missing");
  int D, N, P;
  if (shape_dims(mesh->shape, mesh->n_gp_1d, &D, &N, &P))
    return fail(AD4SM_EINVAL, "mesh: unknown shape, or a tensor rule with other than 1..3 points per direction");
  int rc = select_device(device);
  if (rc) return rc;
  const size_t ng = (size_t)n_elems * D * P * N, nw = (size_t)n_elems * P, nv = (size_t)n_elems * P * N;
  double* buf = nullptr;
  CK(cudaMalloc((void**)&buf, sizeof(double) * (ng + nw + nv + (size_t)n_elems) + 64));
  double *g = buf, *w = g + ng, *v = w + nw, *vol = v + nv;
  rc = construct_run(mesh, n_elems, nodes, g, w, Nval ? v : nullptr, V ? vol : nullptr, (cudaStream_t)0);
  cudaError_t e = cudaSuccess;
  if (!rc) e = cudaMemcpy(gradN, g, sizeof(double) * ng, cudaMemcpyDeviceToHost);
  if (!rc && e == cudaSuccess) e = cudaMemcpy(wgt, w, sizeof(double) * nw, cudaMemcpyDeviceToHost);
  if (!rc && e == cudaSuccess && Nval) e = cudaMemcpy(Nval, v, sizeof(double) * nv, cudaMemcpyDeviceToHost);
  if (!rc && e == cudaSuccess && V) e = cudaMemcpy(V, vol, sizeof(double) * (size_t)n_elems, cudaMemcpyDeviceToHost);
  cudaFree(buf);
  if (rc) return rc;
  if (e != cudaSuccess) return fail(AD4SM_ENODEV, std::string("construct: ") + cudaGetErrorString(e));
  return AD4SM_OK;
}

int ad4sm_consteq_eval(const ad4sm_consteq* q, int64_t n_dof, const double* u, const double* lambda, int32_t device, double* veqs,
                       double* grad, double* hess) {
  if (!q || !u || !veqs || !grad || n_dof <= 0) return fail(AD4SM_EINVAL, "consteq: arguments missing");
  if (q->n_eqs < 0 || (q->n_eqs > 0 && (!q->dof_ptr || !q->dofs || !q->c || !q->c0))) return fail(AD4SM_EINVAL, "consteq: arrays missing");
  if (q->n_eqs == 0) return AD4SM_OK;
  const long long nd = q->dof_ptr[q->n_eqs], nq = q->q_ptr ? q->q_ptr[q->n_eqs] : 0;
  for (long long ii = 0; ii < q->n_eqs; ii++) {
    const long long n = q->dof_ptr[ii + 1] - q->dof_ptr[ii];
    if (n < 0) return fail(AD4SM_EINVAL, "consteq: dof_ptr must be non-decreasing");
    if (q->q_ptr) {
      const long long m = q->q_ptr[ii + 1] - q->q_ptr[ii];
      if (m != 0 && m != n * n) return fail(AD4SM_EINVAL, "consteq: Q of an equation must be empty or n x n");
    }
  }
  if (nq > 0 && !q->Q) return fail(AD4SM_EINVAL, "consteq: Q missing");
  for (long long k = 0; k < nd; k++)
    if (q->dofs[k] < 1 || q->dofs[k] > n_dof) return fail(AD4SM_EINVAL, "consteq: dof index out of range 1..length(u) (BoundsError)");
  int rc = select_device(device);
  if (rc) return rc;
  // one device buffer: [dof_ptr | dofs | q_ptr] int64, then [c | c0 | Q | u | lambda | veqs | grad | hess] doubles
  const size_t ni = (size_t)(q->n_eqs + 1) * 2 + (size_t)nd;
  const size_t nf = (size_t)nd * 2 + (size_t)q->n_eqs * 3 + (size_t)nq * 2 + (size_t)n_dof;
  char* buf = nullptr;
  CK(cudaMalloc((void**)&buf, ni * 8 + nf * 8 + 64));
  long long* ip = (long long*)buf;
  double* fp = (double*)(buf + ni * 8);
  ConstEqArgs a;
  memset(&a, 0, sizeof(a));
  a.n_eqs = q->n_eqs;
  long long *d_ptr = ip, *d_dofs = d_ptr + q->n_eqs + 1, *d_qptr = d_dofs + nd;
  double *d_c = fp, *d_c0 = d_c + nd, *d_Q = d_c0 + q->n_eqs, *d_u = d_Q + nq, *d_lam = d_u + n_dof, *d_v = d_lam + q->n_eqs,
         *d_g = d_v + q->n_eqs, *d_h = d_g + nd;
  cudaError_t e = cudaMemcpy(d_ptr, q->dof_ptr, 8 * (size_t)(q->n_eqs + 1), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_dofs, q->dofs, 8 * (size_t)nd, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && q->q_ptr) e = cudaMemcpy(d_qptr, q->q_ptr, 8 * (size_t)(q->n_eqs + 1), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_c, q->c, 8 * (size_t)nd, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_c0, q->c0, 8 * (size_t)q->n_eqs, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && nq) e = cudaMemcpy(d_Q, q->Q, 8 * (size_t)nq, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(d_u, u, 8 * (size_t)n_dof, cudaMemcpyHostToDevice);
  if (e == cudaSuccess && lambda) e = cudaMemcpy(d_lam, lambda, 8 * (size_t)q->n_eqs, cudaMemcpyHostToDevice);
  a.dof_ptr = d_ptr, a.dofs = d_dofs, a.q_ptr = q->q_ptr ? d_qptr : nullptr;
  a.c = d_c, a.c0 = d_c0, a.Q = d_Q, a.u = d_u, a.lambda = lambda ? d_lam : nullptr;
  a.veqs = d_v, a.grad = d_g, a.hess = (hess && nq) ? d_h : nullptr;
  if (e == cudaSuccess) e = launch_consteq(a, (cudaStream_t)0);
  if (e == cudaSuccess) e = cudaMemcpy(veqs, d_v, 8 * (size_t)q->n_eqs, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess) e = cudaMemcpy(grad, d_g, 8 * (size_t)nd, cudaMemcpyDeviceToHost);
  if (e == cudaSuccess && a.hess) e = cudaMemcpy(hess, d_h, 8 * (size_t)nq, cudaMemcpyDeviceToHost);
  cudaFree(buf);
  if (e != cudaSuccess) return fail(AD4SM_ENODEV, std::string("consteq: ") + cudaGetErrorString(e));
  return AD4SM_OK;
}

void ad4sm_plan_destroy(ad4sm_plan* p) {
  if (!p) return;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  for (void* m : p->ipc_open) cudaIpcCloseMemHandle(m);
  for (void* a : p->allocs) cudaFree(a);
  for (int f = 0; f < 2; f++) {
    cudaFree(p->field[f].dz_colptr);
    cudaFree(p->field[f].dz_rowval);
    cudaFree(p->field[f].dz_nzval);
  }
  for (auto& s : p->ev_pool)
    for (int i = 0; i < 5; i++) cudaEventDestroy(s.e[i]);
  if (p->ev_r_ready) cudaEventDestroy(p->ev_r_ready);
  if (p->ev_copied) cudaEventDestroy(p->ev_copied);
  if (p->ev_start) cudaEventDestroy(p->ev_start);
  if (!p->dprof.empty()) {  // AD4SM_DIST_PROF: average intervals from the end of the element stage, last half of the evaluations
    cudaDeviceSynchronize();
    const size_t n = p->dprof.size() / 4, k0 = n / 2;
    double a[3] = {0, 0, 0};
    for (size_t k = k0; k < n; k++)
      for (int j = 0; j < 3; j++) {
        float ms = 0.f;
        cudaEventElapsedTime(&ms, p->dprof[4 * k], p->dprof[4 * k + 1 + j]);
        a[j] += ms;
      }
    fprintf(stderr, "DISTPROF(lib) %zu evaluations: element stage end -> exchange done %.4f ms, -> interior done %.4f ms, -> step end %.4f ms\n",
            n - k0, a[0] / (n - k0), a[1] / (n - k0), a[2] / (n - k0));
    for (cudaEvent_t e : p->dprof) cudaEventDestroy(e);
  }
  if (p->ev_exports) cudaEventDestroy(p->ev_exports);
  if (p->ev_pushed) cudaEventDestroy(p->ev_pushed);
  for (cudaEvent_t e : p->ev_h2d)
    if (e) cudaEventDestroy(e);
  if (p->copy_stream) cudaStreamDestroy(p->copy_stream);
  if (p->own_stream) cudaStreamDestroy(p->own_stream);
  delete p;
}

int ad4sm_plan_create(const ad4sm_batch* batches, int32_t n_batches, int64_t n_nodes, int32_t dofs_per_node,
                      int32_t device, ad4sm_plan** plan_out) {
  if (!plan_out) return fail(AD4SM_EINVAL, "plan_out is NULL");
  *plan_out = nullptr;
  if (!batches || n_batches <= 0) return fail(AD4SM_EINVAL, "makeϕrKt: `elems` is empty (elasticelements.jl:210)");
  if (n_nodes <= 0 || n_nodes >= (1ll << 31)) return fail(AD4SM_EINVAL, "n_nodes out of range");
  const int N = batches[0].n_elem_nodes, D = batches[0].dim;
  long long ne = 0, n_gp = 0;
  bool has_cp = false;
  std::vector<int> mfs(n_batches);
  for (int b = 0; b < n_batches; b++) {
    const ad4sm_batch& B = batches[b];
    if (B.n_elem_nodes != N)
      return fail(AD4SM_EINVAL, "all elements must have the same number of dofs (elasticelements.jl:212-214)");
    if (B.dim != D) return fail(AD4SM_EINVAL, "all batches must have the same dimension D");
    const bool halo = (B.batch_flags & AD4SM_BATCH_HALO) != 0;
    if (B.n_elems < 0 || (B.n_elems > 0 && (!B.nodes || (!halo && !B.mesh && (!B.gradN || !B.wgt)))))
      return fail(AD4SM_EINVAL, "batch arrays missing");
    if (B.elem_kind == AD4SM_ELEM_CP && B.n_elems > 0 && !halo && !B.mesh && !B.Nval)
      return fail(AD4SM_EINVAL, "CPElem batch needs Nval");
    if (B.elem_kind == AD4SM_ELEM_CAS) {
      if (B.dim != 2 || (B.n_elem_nodes != 3 && B.n_elem_nodes != 4)) return fail(AD4SM_EINVAL, "CAS batches are 2-D with 3 or 4 nodes (ASTria / ASQuad)");
      if (B.mesh) return fail(AD4SM_EINVAL, "CAS batches have no device constructor: pass gradN, wgt and the hoop row N/X0");
      if (B.n_elems > 0 && !halo && !B.Nval) return fail(AD4SM_EINVAL, "CAS batch needs Nval = N / X0 (the hoop row)");
    }
    if (B.mesh && !halo) {  // constructed on the device from the coordinates: the declared shape must be the mesh's
      int mD, mN, mP;
      if (shape_dims(B.mesh->shape, B.mesh->n_gp_1d, &mD, &mN, &mP) || mD != B.dim || mN != B.n_elem_nodes || mP != B.n_gauss)
        return fail(AD4SM_EINVAL, "batch.mesh: shape / rule do not match dim, n_elem_nodes, n_gauss of the batch");
      if (B.mesh->n_nodes != n_nodes) return fail(AD4SM_EINVAL, "batch.mesh: n_nodes differs from the plan's");
    }
    if (B.elem_kind != AD4SM_ELEM_CE && B.elem_kind != AD4SM_ELEM_CP && B.elem_kind != AD4SM_ELEM_CAS)
      return fail(AD4SM_EINVAL, "unknown elem_kind");
    std::string why;
    if (mat_family(B, &mfs[b], &why)) return fail(AD4SM_EINVAL, why);
    const char* r = k1_unsupported_reason(B.dim, B.n_elem_nodes, B.n_gauss, mfs[b], 0);
    if (r) return fail(AD4SM_EINVAL, r);
    has_cp |= (B.elem_kind == AD4SM_ELEM_CP);
    ne += B.n_elems;
    n_gp += B.n_elems * B.n_gauss;
  }
  if (ne == 0) return fail(AD4SM_EINVAL, "makeϕrKt: `elems` is empty (elasticelements.jl:210)");
  if (dofs_per_node < D) return fail(AD4SM_EINVAL, "size(u,1) must be >= element dimension (getF reads rows 1:D, toolkit:13)");
  const int NPB = n_pair_blocks(N);
  if ((double)ne * N * N >= 4.0e9 || (double)ne * NPB * 2 >= 4.0e9)
    return fail(AD4SM_EINVAL, "too many elements for 32-bit source maps on one device; partition the mesh");

  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(AD4SM_ENODEV, "no CUDA device: libad4sm_b200 has no CPU fallback");
  int dev = device;
  if (dev < 0) CK(cudaGetDevice(&dev));
  if (dev >= ndev) return fail(AD4SM_EINVAL, "device ordinal out of range");
  CK(cudaSetDevice(dev));

  // AD4SM_PLAN_TIMING=1: wall time of every stage of the plan construction to stderr (developer tool)
  static const bool plan_timing = getenv("AD4SM_PLAN_TIMING") != nullptr;
  auto t_last = std::chrono::steady_clock::now();
  auto stage = [&](const char* what) {
    if (!plan_timing) return;
    const auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "PLAN %-28s %.3f s\n", what, std::chrono::duration<double>(t - t_last).count());
    t_last = t;
  };
  // ---- host-side maps ---------------------------------------------------------------------------
  // slots: position in the concatenation of the batches; `key`: position in `elems` (summation order)
  std::vector<int> nodes0((size_t)ne * N);
  std::vector<long long> key(ne);
  bool keyed = false;
  {
    long long s = 0;
    for (int b = 0; b < n_batches; b++) {
      const ad4sm_batch& B = batches[b];
      for (long long e = 0; e < B.n_elems; e++, s++) {
        for (int a = 0; a < N; a++) {
          const long long nd = B.nodes[e * N + a];
          if (nd < 1 || nd > n_nodes) return fail(AD4SM_EINVAL, "node id out of range 1..n_nodes");
          nodes0[s * N + a] = (int)(nd - 1);
        }
        key[s] = B.orig_index ? B.orig_index[e] : s;
        keyed |= (B.orig_index != nullptr);
      }
    }
  }
  std::vector<unsigned> order(ne);  // slots in summation order
  for (long long s = 0; s < ne; s++) order[s] = (unsigned)s;
  if (keyed) std::stable_sort(order.begin(), order.end(), [&](unsigned x, unsigned y) { return key[x] < key[y]; });

  stage("nodes, keys, order");
  // node -> (slot, local node) adjacency, lists in summation order
  std::vector<unsigned> adj_ptr(n_nodes + 1, 0), adj((size_t)ne * N);
  for (size_t i = 0; i < nodes0.size(); i++) adj_ptr[nodes0[i] + 1]++;
  for (long long i = 0; i < n_nodes; i++) adj_ptr[i + 1] += adj_ptr[i];
  {
    std::vector<unsigned> cur(adj_ptr.begin(), adj_ptr.end() - 1);
    for (long long o = 0; o < ne; o++) {
      const unsigned s = order[o];
      for (int a = 0; a < N; a++) adj[cur[nodes0[(size_t)s * N + a]]++] = s * (unsigned)N + a;
    }
  }
  stage("adjacency");
  // neighbour lists (sorted, unique) = block columns of each row node
  std::vector<unsigned> nbr_ptr(n_nodes + 1, 0);
#pragma omp parallel
  {
    std::vector<int> tmp;
#pragma omp for schedule(static)
    for (long long i = 0; i < n_nodes; i++) {
      tmp.clear();
      for (unsigned s = adj_ptr[i]; s < adj_ptr[i + 1]; s++) {
        const unsigned slot = adj[s] / N;
        for (int b = 0; b < N; b++) tmp.push_back(nodes0[(size_t)slot * N + b]);
      }
      std::sort(tmp.begin(), tmp.end());
      nbr_ptr[i + 1] = (unsigned)(std::unique(tmp.begin(), tmp.end()) - tmp.begin());
    }
  }
  {
    unsigned long long tot = 0;
    for (long long i = 0; i < n_nodes; i++) {
      tot += nbr_ptr[i + 1];
      if (tot >= 0xffffffffull) return fail(AD4SM_EINVAL, "too many node-pair blocks for one device; partition the mesh");
      nbr_ptr[i + 1] = (unsigned)tot;
    }
  }
  stage("neighbour counts");
  const long long n_blocks = nbr_ptr[n_nodes];
  std::vector<unsigned> nbr(n_blocks), blk_row(n_blocks), blk_ptr(n_blocks + 1, 0);
  // pidx[s * N + b]: position of node b of adjacency entry s's element in the neighbour list of the entry's node -- found
  // once (binary search) and reused by the source map and by the destination-sorted tables
  bool small_rows = true;
  for (long long i = 0; i < n_nodes && small_rows; i++) small_rows = nbr_ptr[i + 1] - nbr_ptr[i] <= 0xffffu;
  std::vector<unsigned short> pidx(small_rows ? (size_t)ne * N * N : 0);
#pragma omp parallel
  {
    std::vector<int> tmp;
    std::vector<unsigned> cnt;
#pragma omp for schedule(static)
    for (long long i = 0; i < n_nodes; i++) {
      tmp.clear();
      for (unsigned s = adj_ptr[i]; s < adj_ptr[i + 1]; s++) {
        const unsigned slot = adj[s] / N;
        for (int b = 0; b < N; b++) tmp.push_back(nodes0[(size_t)slot * N + b]);
      }
      std::sort(tmp.begin(), tmp.end());
      tmp.erase(std::unique(tmp.begin(), tmp.end()), tmp.end());
      const unsigned b0 = nbr_ptr[i];
      for (size_t p = 0; p < tmp.size(); p++) {
        nbr[b0 + p] = (unsigned)tmp[p];
        blk_row[b0 + p] = (unsigned)i;
      }
      // source counts per block
      cnt.assign(tmp.size(), 0);
      for (unsigned s = adj_ptr[i]; s < adj_ptr[i + 1]; s++) {
        const unsigned slot = adj[s] / N;
        for (int b = 0; b < N; b++) {
          const int nb = nodes0[(size_t)slot * N + b];
          const size_t p = std::lower_bound(tmp.begin(), tmp.end(), nb) - tmp.begin();
          if (small_rows) pidx[(size_t)s * N + b] = (unsigned short)p;
          cnt[p]++;
        }
      }
      for (size_t p = 0; p < tmp.size(); p++) blk_ptr[b0 + p + 1] = cnt[p];
    }
  }
  auto pos_of = [&](long long i, unsigned s, int b) -> unsigned {  // position of that node in row i's neighbour list
    if (small_rows) return pidx[(size_t)s * N + b];
    const unsigned* lo = nbr.data() + nbr_ptr[i];
    const unsigned nb = (unsigned)nodes0[(size_t)(adj[s] / N) * N + b];
    return (unsigned)(std::lower_bound(lo, (const unsigned*)(nbr.data() + nbr_ptr[i + 1]), nb) - lo);
  };
  stage("neighbour lists, block counts");
  for (long long k = 0; k < n_blocks; k++) blk_ptr[k + 1] += blk_ptr[k];
  std::vector<unsigned> src((size_t)ne * N * N);
#pragma omp parallel
  {
    std::vector<unsigned> cur;
#pragma omp for schedule(static)
    for (long long i = 0; i < n_nodes; i++) {
      const unsigned b0 = nbr_ptr[i], nn = nbr_ptr[i + 1] - b0;
      cur.assign(nn, 0);
      for (unsigned s = adj_ptr[i]; s < adj_ptr[i + 1]; s++) {
        const unsigned slot = adj[s] / N, a = adj[s] % N;
        for (int b = 0; b < N; b++) {
          const unsigned p = pos_of(i, s, b);
          int idx, tr;
          pair_slot(N, (int)a, b, idx, tr);
          src[blk_ptr[b0 + p] + cur[p]++] = ((slot * (unsigned)NPB + (unsigned)idx) << 1) | (unsigned)tr;
        }
      }
    }
  }

  stage("source map");
  // ---- device side ------------------------------------------------------------------------------
  ad4sm_plan* p = new ad4sm_plan;
  p->device = dev;
  p->N = N;
  p->D = D;
  p->dpn = dofs_per_node;
  p->n_nodes = n_nodes;
  p->ne = ne;
  p->n_blocks = n_blocks;
  p->n_gp = n_gp;
  p->has_cp = has_cp;
  int rc = 0;
#define TRY(x)                   \
  do {                           \
    rc = (x);                    \
    if (rc) {                    \
      ad4sm_plan_destroy(p);     \
      return rc;                 \
    }                            \
  } while (0)
  {
    cudaError_t e = cudaStreamCreateWithFlags(&p->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) {
      delete p;
      return fail(AD4SM_ENODEV, std::string("cudaStreamCreate: ") + cudaGetErrorString(e));
    }
    p->stream = p->own_stream;
    // the copy stream also carries the interface exchange of ad4sm_dist_eval*, whose few CTAs must not queue behind the
    // interior assembly running on the main stream: highest priority
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if (const char* e = getenv("AD4SM_COPY_PRIO"))  // experiment knob: 0 = default priority
      if (atoi(e) == 0) prio_hi = 0;
    if (cudaStreamCreateWithPriority(&p->copy_stream, cudaStreamNonBlocking, prio_hi) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_r_ready, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&p->ev_copied, cudaEventDisableTiming) != cudaSuccess) {
      ad4sm_plan_destroy(p);
      return fail(AD4SM_ENODEV, "cudaStreamCreate/cudaEventCreate failed");
    }
  }
  // pipelined upload (host API): a single ordinary batch (HALO batches may follow it) whose node numbering follows the element order (structured /
  // bandwidth-reduced meshes).  Slices of 1/16, 1/4, 1/2 and all of the elements, on 32-element boundaries (keeps the
  // 16-byte alignment of every per-element array); each needs the nodes below its running maximum.
  {
    const char* env = getenv("AD4SM_H2D_SLICES");
    const bool want = !env || atoi(env) != 0;
    int n_ord = 0;
    for (int b = 0; b < n_batches; b++) n_ord += !(batches[b].batch_flags & AD4SM_BATCH_HALO);
    const long long ne_ord = (n_ord == 1 && !(batches[0].batch_flags & AD4SM_BATCH_HALO)) ? batches[0].n_elems : 0;
    if (want && ne_ord >= 65536) {
      const long long frac16[ad4sm_plan::H2D_MAX] = {1, 4, 8, 16};
      long long mx = 0, e = 0;
      bool ok = true;
      for (int c = 0; c < ad4sm_plan::H2D_MAX; c++) {
        const long long e1 = c == ad4sm_plan::H2D_MAX - 1 ? ne_ord : ((ne_ord * frac16[c] / 16) & ~31ll);
        for (; e < e1; e++)
          for (int a = 0; a < N; a++) mx = std::max<long long>(mx, nodes0[(size_t)e * N + a] + 1);
        p->h2d_elem[c + 1] = e1;
        p->h2d_node[c + 1] = c == ad4sm_plan::H2D_MAX - 1 ? n_nodes : mx;
        if (c == 0 && mx > n_nodes / 4) ok = false;  // the first slice already needs most of u: nothing to overlap
      }
      if (ok) {
        p->n_h2d = ad4sm_plan::H2D_MAX;
        bool evok = cudaEventCreateWithFlags(&p->ev_start, cudaEventDisableTiming) == cudaSuccess;
        for (int c = 0; c < ad4sm_plan::H2D_MAX && evok; c++)
          evok = cudaEventCreateWithFlags(&p->ev_h2d[c], cudaEventDisableTiming) == cudaSuccess;
        if (!evok) {
          ad4sm_plan_destroy(p);
          return fail(AD4SM_ENODEV, "cudaEventCreate failed");
        }
      }
    }
  }
  stage("streams, upload slices");
  if (keyed) TRY(dev_upload(p, &p->key_dev, key.data(), key.size()));
  TRY(dev_upload(p, &p->blk_row, blk_row.data(), blk_row.size()));
  TRY(dev_upload(p, &p->nbr_ptr, nbr_ptr.data(), nbr_ptr.size()));
  TRY(dev_upload(p, &p->nbr, nbr.data(), nbr.size()));
  TRY(dev_upload(p, &p->blk_ptr, blk_ptr.data(), blk_ptr.size()));
  TRY(dev_upload(p, &p->src, src.data(), src.size()));
  TRY(dev_upload(p, &p->adj_ptr, adj_ptr.data(), adj_ptr.size()));
  TRY(dev_upload(p, &p->adj, adj.data(), adj.size()));
  {
    long long s0 = 0, h0 = 0;
    for (int b = 0; b < n_batches; b++) {
      const ad4sm_batch& B = batches[b];
      BatchPlan bp;
      memset(&bp, 0, sizeof(bp));
      bp.k.mat.kind = B.mat_kind;
      bp.k.mat.base = B.mat_base;
      bp.k.mat.flags = B.mat_flags;
      for (int i = 0; i < 8; i++) bp.k.mat.p[i] = B.mat_params[i];
      bp.k.ne = B.n_elems;
      bp.k.slot0 = s0;
      bp.k.D = D;
      bp.k.N = N;
      bp.k.P = B.n_gauss;
      bp.k.mf = mfs[b];
      bp.k.dpn = dofs_per_node;
      bp.k.cas = B.elem_kind == AD4SM_ELEM_CAS;
      bp.hist_off = h0;
      bp.halo = (B.batch_flags & AD4SM_BATCH_HALO) != 0;
      if (!bp.halo) {
        int* nd = nullptr;
        TRY(dev_upload(p, &nd, nodes0.data() + (size_t)s0 * N, (size_t)B.n_elems * N));
        bp.k.nodes = nd;
        double* x = nullptr;
        if (B.mesh) {  // batched constructor on the device: no host ∇N, no upload
          double *g = nullptr, *w = nullptr, *v = nullptr;
          TRY(dev_alloc(p, &g, (size_t)B.n_elems * D * B.n_gauss * N));
          TRY(dev_alloc(p, &w, (size_t)B.n_elems * B.n_gauss));
          if (B.elem_kind == AD4SM_ELEM_CP) TRY(dev_alloc(p, &v, (size_t)B.n_elems * B.n_gauss * N));
          TRY(construct_run(B.mesh, B.n_elems, B.nodes, g, w, v, nullptr, p->stream));
          bp.k.gradN = g;
          bp.k.wgt = w;
          bp.k.Nval = v;
        } else {
          TRY(dev_upload(p, &x, B.gradN, (size_t)B.n_elems * D * B.n_gauss * N));
          bp.k.gradN = x;
          TRY(dev_upload(p, &x, B.wgt, (size_t)B.n_elems * B.n_gauss));
          bp.k.wgt = x;
          if (B.elem_kind == AD4SM_ELEM_CP || B.elem_kind == AD4SM_ELEM_CAS) {
            TRY(dev_upload(p, &x, B.Nval, (size_t)B.n_elems * B.n_gauss * N));
            bp.k.Nval = x;
          }
        }
      }
      p->batches.push_back(bp);
      s0 += B.n_elems;
      h0 += B.n_elems * B.n_gauss;
    }
  }
  stage("map + element uploads / constructors");
  {  // HALO slots get a second set of rows behind the staging (receive buffer B of the peer-push protocol)
    long long nh = 0;
    for (int b = 0; b < n_batches; b++)
      if (batches[b].batch_flags & AD4SM_BATCH_HALO) nh += batches[b].n_elems;
    p->n_halo = nh;
  }
  TRY(dev_alloc(p, &p->Ke, (size_t)(ne + p->n_halo) * NPB * D * D));
  TRY(dev_alloc(p, &p->re, (size_t)(ne + p->n_halo) * N * D));
  TRY(dev_alloc(p, &p->phie, (size_t)ne));
  CK(cudaMemset(p->phie, 0, sizeof(double) * (size_t)ne));  // HALO slots stay 0: they add nothing to ϕ on this rank
  TRY(dev_alloc(p, &p->partials, (size_t)phi_reduce_partials()));
  TRY(dev_alloc(p, &p->phi_dev, 1));
  TRY(dev_alloc(p, &p->n_zero_dev, 1));
  TRY(dev_alloc(p, &p->u_dev, (size_t)n_nodes * dofs_per_node));
  FieldOut& fu = p->field[AD4SM_FIELD_U];
  fu.dpn = dofs_per_node;
  fu.n_dof = n_nodes * dofs_per_node;
  fu.nnz = n_blocks * dofs_per_node * dofs_per_node;
  TRY(dev_alloc(p, &fu.nzval, (size_t)fu.nnz));
  TRY(dev_alloc(p, &fu.r, (size_t)fu.n_dof));
  if (has_cp) {
    FieldOut& fd = p->field[AD4SM_FIELD_D];
    fd.dpn = 1;
    fd.n_dof = n_nodes;
    fd.nnz = n_blocks;
    TRY(dev_alloc(p, &fd.nzval, (size_t)fd.nnz));
    TRY(dev_alloc(p, &fd.r, (size_t)fd.n_dof));
    TRY(dev_alloc(p, &p->d_dev, (size_t)n_nodes));
    TRY(dev_alloc(p, &p->hist_dev, (size_t)n_gp * 2));
  }
  stage("output allocations");
  // ---- destination-sorted staging: contributions grouped by unordered node pair, in `elems` order inside a pair ----
  {
    const char* env = getenv("AD4SM_SORTED");
    // every evaluated (non-HALO) batch must come before the HALO batches and have a K1 that writes sorted blocks
    bool shape_ok = !(batches[0].batch_flags & AD4SM_BATCH_HALO);
    bool seen_halo = false;
    long long n_own = 0;
    int n_own_batches = 0;
    for (int b = 0; b < n_batches && shape_ok; b++) {
      const ad4sm_batch& B = batches[b];
      if (B.batch_flags & AD4SM_BATCH_HALO) {
        seen_halo = true;
        continue;
      }
      K1Config c;
      shape_ok = !seen_halo && k1_config(D, N, B.n_gauss, B.elem_kind == AD4SM_ELEM_CP && B.mat_kind == AD4SM_MAT_PHASEFIELD, B.n_elems, &c);
      n_own += B.n_elems;
      n_own_batches++;
    }
    p->n_own_batches = n_own_batches;
    if (!(env && env[0] == '0') && shape_ok && D == 3 && dofs_per_node == 3 &&
        (double)n_blocks * 9.0 < 4.0e9 && (double)ne * NPB < 2.0e9 && n_blocks < (1ll << 31)) {
      // pairs {i <= p} are numbered row by row; in a (sorted) neighbour list the upper ones are the tail
      std::vector<unsigned> up0(n_nodes + 1, 0), first_up(n_nodes);
#pragma omp parallel for schedule(static)
      for (long long i = 0; i < n_nodes; i++) {
        const unsigned* lo = nbr.data() + nbr_ptr[i];
        const unsigned* hi = nbr.data() + nbr_ptr[i + 1];
        first_up[i] = (unsigned)(std::lower_bound(lo, hi, (unsigned)i) - lo);
        up0[i + 1] = (unsigned)(hi - lo) - first_up[i];
      }
      for (long long i = 0; i < n_nodes; i++) up0[i + 1] += up0[i];
      const long long U = up0[n_nodes];
      auto find_blk = [&](unsigned row, unsigned col) -> unsigned {
        const unsigned* lo = nbr.data() + nbr_ptr[row];
        const unsigned* hi = nbr.data() + nbr_ptr[row + 1];
        return (unsigned)(std::lower_bound(lo, hi, col) - nbr.data());
      };
      const size_t n_contrib = (size_t)ne * NPB;
      std::vector<unsigned> cptr(U + 1, 0), cpos(n_contrib);
      // Every contribution (slot, stored block idx) belongs to exactly one pair {i <= p}; seen from the smaller node i it
      // is found by scanning adj[i], which lists the elements in summation (`elems`) order.  Two passes over the row
      // nodes, both parallel and free of atomics: (1) contributions per pair, (2) positions.  Inside one element (only
      // an element that names a node twice can contribute to a pair more than once) the order is that of idx.
      struct Item {
        unsigned idx, p, flag;
      };
      auto scan_row = [&](long long i, std::vector<Item>& items, auto&& emit) {
        for (unsigned s = adj_ptr[i]; s < adj_ptr[i + 1];) {
          const unsigned slot = adj[s] / N;
          items.clear();
          unsigned s1 = s;
          for (; s1 < adj_ptr[i + 1] && adj[s1] / N == slot; s1++) {
            const int a = (int)(adj[s1] % N);
            for (int b = 0; b < N; b++) {
              const unsigned nb = (unsigned)nodes0[(size_t)slot * N + b];
              if (nb < (unsigned)i || (nb == (unsigned)i && b < a)) continue;  // the smaller node's row / once per local pair
              int idx, tr;
              pair_slot(N, a, b, idx, tr);
              // the stored block holds K(a', b'), (a', b') = (a, b) or, transposed, (b, a); flag = node(a') > node(b')
              items.push_back({(unsigned)idx, pos_of(i, s1, b) - first_up[i], (tr && nb > (unsigned)i) ? 1u : 0u});
            }
          }
          if (s1 - s > 1) std::sort(items.begin(), items.end(), [](const Item& x, const Item& y) { return x.idx < y.idx; });
          for (const Item& it : items) emit(slot, it);
          s = s1;
        }
      };
#pragma omp parallel
      {
        std::vector<Item> items;
#pragma omp for schedule(static)
        for (long long i = 0; i < n_nodes; i++)
          scan_row(i, items, [&](unsigned, const Item& it) { cptr[up0[i] + it.p + 1]++; });
      }
      for (long long u = 0; u < U; u++) cptr[u + 1] += cptr[u];
#pragma omp parallel
      {
        std::vector<Item> items;
        std::vector<unsigned> cur;
#pragma omp for schedule(static)
        for (long long i = 0; i < n_nodes; i++) {
          cur.assign(up0[i + 1] - up0[i], 0);
          scan_row(i, items, [&](unsigned slot, const Item& it) {
            cpos[(size_t)slot * NPB + it.idx] = ((cptr[up0[i] + it.p] + cur[it.p]++) << 1) | it.flag;
          });
        }
      }
      stage("sorted: dest, cptr, cpos");
      std::vector<unsigned> out_up(U), out_lo(U), nn_pack(U);
      const unsigned dd = (unsigned)(dofs_per_node * dofs_per_node);
      bool ok = true;
#pragma omp parallel for schedule(static)
      for (long long i = 0; i < n_nodes; i++)
        for (unsigned b = nbr_ptr[i]; b < nbr_ptr[i + 1]; b++) {
          const unsigned pn = nbr[b];
          if (pn < (unsigned)i) continue;
          const unsigned u = up0[i] + (b - nbr_ptr[i]) - first_up[i];
          const unsigned nni = nbr_ptr[i + 1] - nbr_ptr[i], nnp = nbr_ptr[pn + 1] - nbr_ptr[pn];
          if (nni > 0xffffu || nnp > 0xffffu) ok = false;
          out_up[u] = nbr_ptr[i] * dd + (b - nbr_ptr[i]) * (unsigned)dofs_per_node;
          const unsigned b2 = find_blk(pn, (unsigned)i);
          out_lo[u] = nbr_ptr[pn] * dd + (b2 - nbr_ptr[pn]) * (unsigned)dofs_per_node;
          nn_pack[u] = nni | (nnp << 16);
        }
      stage("sorted: pair tables");
      if (ok) {
        int rc2 = dev_upload(p, &p->cpos, cpos.data(), cpos.size());
        if (!rc2) rc2 = dev_upload(p, &p->cptr, cptr.data(), cptr.size());
        if (!rc2) rc2 = dev_upload(p, &p->out_up, out_up.data(), out_up.size());
        if (!rc2) rc2 = dev_upload(p, &p->out_lo, out_lo.data(), out_lo.size());
        if (!rc2) rc2 = dev_upload(p, &p->nn_pack, nn_pack.data(), nn_pack.size());
        if (!rc2) rc2 = dev_alloc(p, &p->C, n_contrib * (size_t)CBLK3);
        if (rc2) {
          ad4sm_plan_destroy(p);
          return rc2;
        }
        p->n_pairs = U;
        p->sorted_built = true;
        p->n_own_sorted = n_own;
        // first node row with a HALO slot in its adjacency: rows before it (and their pairs, which are numbered row by
        // row under the smaller node) are complete without the exchange.  Rounded down to whole K2s CTAs (128 pairs).
        long long sn = n_nodes;
        if (n_own < ne) {
#pragma omp parallel for schedule(static) reduction(min : sn)
          for (long long i = 0; i < n_nodes; i++) {
            if (i >= sn) continue;
            for (unsigned s2 = adj_ptr[i]; s2 < adj_ptr[i + 1]; s2++)
              if ((long long)(adj[s2] / N) >= n_own) {
                sn = i;
                break;
              }
          }
        }
        while (sn > 0 && (up0[sn] % 128u) != 0) sn--;
        p->split_node = sn;
        p->split_pair = up0[sn];
      }
    }
  }
#undef TRY
  stage("sorted: uploads");
  *plan_out = p;
  return AD4SM_OK;
}

static int check_field(ad4sm_plan* p, int32_t field) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  if (field != AD4SM_FIELD_U && field != AD4SM_FIELD_D) return fail(AD4SM_EINVAL, "unknown field");
  if (field == AD4SM_FIELD_D && !p->has_cp) return fail(AD4SM_EINVAL, "plan has no CPElem batch: no phase-field dofs");
  return 0;
}

static AsmArgs asm_args(ad4sm_plan* p, int field) {
  AsmArgs a;
  a.n_nodes = p->n_nodes;
  a.n_blocks = p->n_blocks;
  a.N = p->N;
  a.DB = field == AD4SM_FIELD_D ? 1 : p->D;
  a.dpn = p->field[field].dpn;
  a.blk_row = p->blk_row;
  a.nbr_ptr = p->nbr_ptr;
  a.nbr = p->nbr;
  a.blk_ptr = p->blk_ptr;
  a.src = p->src;
  a.adj_ptr = p->adj_ptr;
  a.adj = p->adj;
  a.Ke = p->Ke;
  a.re = p->re;
  a.nzval = p->field[field].nzval;
  a.r = p->field[field].r;
  a.n_zero = p->n_zero_dev;
  a.n_pairs = (p->sorted_built && field == AD4SM_FIELD_U) ? p->n_pairs : 0;
  a.cptr = p->cptr;
  a.out_up = p->out_up;
  a.out_lo = p->out_lo;
  a.nn_pack = p->nn_pack;
  a.C = p->C;
  static int k2s_exp = -1, k2s_direct = -1;
  if (k2s_exp < 0) k2s_exp = getenv("AD4SM_EXP_K2S") ? atoi(getenv("AD4SM_EXP_K2S")) : 0;
  if (k2s_direct < 0) k2s_direct = getenv("AD4SM_K2S_DIRECT") ? atoi(getenv("AD4SM_K2S_DIRECT")) : 0;
  a.k2s_exp = k2s_exp;
  a.k2s_direct = k2s_direct;
  a.halo_adj0 = (unsigned long long)p->n_own_sorted * p->N;
  a.halo_adj_shift = (p->push_pending && p->parity) ? (unsigned long long)p->n_halo * p->N : 0ull;
  a.node0 = 0;
  a.node1 = a.n_nodes;
  a.pair0 = 0;
  a.pair1 = a.n_pairs;
  return a;
}

static int ensure_pattern(ad4sm_plan* p, int field) {
  FieldOut& f = p->field[field];
  if (f.colptr) return 0;
  CK(cudaSetDevice(p->device));
  int rc = dev_alloc(p, &f.colptr, (size_t)f.n_dof + 1);
  if (rc) return rc;
  rc = dev_alloc(p, &f.rowval, (size_t)f.nnz);
  if (rc) return rc;
  CK(launch_pattern(asm_args(p, field), f.colptr, f.rowval, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return 0;
}

int ad4sm_plan_pattern_size(ad4sm_plan* p, int32_t field, int64_t* n_dof, int64_t* nnz) {
  int rc = check_field(p, field);
  if (rc) return rc;
  if (n_dof) *n_dof = p->field[field].n_dof;
  if (nnz) *nnz = p->field[field].nnz;
  return AD4SM_OK;
}

int ad4sm_plan_pattern(ad4sm_plan* p, int32_t field, int64_t* colptr, int64_t* rowval) {
  int rc = check_field(p, field);
  if (rc) return rc;
  rc = ensure_pattern(p, field);
  if (rc) return rc;
  FieldOut& f = p->field[field];
  if (colptr) CK(cudaMemcpy(colptr, f.colptr, sizeof(long long) * (f.n_dof + 1), cudaMemcpyDeviceToHost));
  if (rowval) CK(cudaMemcpy(rowval, f.rowval, sizeof(long long) * f.nnz, cudaMemcpyDeviceToHost));
  return AD4SM_OK;
}

int ad4sm_set_stream(ad4sm_plan* p, void* cuda_stream) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  CK(cudaSetDevice(p->device));
  CK(cudaStreamSynchronize(p->stream));
  p->stream = cuda_stream ? (cudaStream_t)cuda_stream : p->own_stream;
  return AD4SM_OK;
}

static EvSet* next_events(ad4sm_plan* p) {
  if (p->ev_used == p->ev_pool.size()) {
    EvSet s;
    for (int i = 0; i < 5; i++)
      if (cudaEventCreate(&s.e[i]) != cudaSuccess) return nullptr;
    p->ev_pool.push_back(s);
  }
  return &p->ev_pool[p->ev_used++];
}

static int drain_events(ad4sm_plan* p) {
  if (p->ev_used == 0) return 0;
  CK(cudaStreamSynchronize(p->stream));
  for (size_t i = 0; i < p->ev_used; i++) {
    EvSet& s = p->ev_pool[i];
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, s.e[0], s.e[1]));
    p->prof_ms[AD4SM_PROF_ELEMENT] += ms;
    CK(cudaEventElapsedTime(&ms, s.e[4], s.e[2]));
    p->prof_ms[AD4SM_PROF_ASSEMBLE] += ms;
    CK(cudaEventElapsedTime(&ms, s.e[1], s.e[4]));
    p->prof_ms[AD4SM_PROF_GATHER] += ms;  // interface exchange between the element and assembly stages
    CK(cudaEventElapsedTime(&ms, s.e[2], s.e[3]));
    p->prof_ms[AD4SM_PROF_REDUCE] += ms;
  }
  p->ev_used = 0;
  return 0;
}

static int check_eval_args(ad4sm_plan* p, int32_t mode, const double* u_dev, const double* d_dev) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  if (mode < 0 || mode > 3) return fail(AD4SM_EINVAL, "mode must be 0..3");
  if (!u_dev) return fail(AD4SM_EINVAL, "u is NULL");
  const bool needs_d = (mode == MODE_PF_U || mode == MODE_PF_D);
  if (needs_d && !d_dev) return fail(AD4SM_EINVAL, "d is NULL");
  if (needs_d && !p->has_cp) return fail(AD4SM_EINVAL, "phase-field evaluation needs CPElem batches");
  for (const BatchPlan& b : p->batches) {
    const bool is_pf = b.k.mat.kind == MAT_PHASEFIELD;
    if (needs_d && (!is_pf || (!b.k.Nval && !b.halo)))
      return fail(AD4SM_EINVAL, "makeϕrKt(elems,u,d) / makeϕrKt_d need CPElem with a PhaseField material");
    if (!needs_d && is_pf)
      return fail(AD4SM_EINVAL, "PhaseField material evaluated without d: the reference has no getϕ(F, ::PhaseField)");
  }
  return 0;
}

// `internal`: called from ad4sm_eval_device, which also runs the assembly and may therefore pick a staging layout of its
// own (destination-sorted).  The public two-stage API promises the element-major
// staging of ad4sm_staging_view_get between the stages (interface exchange of partitioned runs).
static int eval_elements_impl(ad4sm_plan* p, int32_t mode, const double* u_dev, const double* d_dev, double* hist_dev,
                              bool internal) {
  int rc = check_eval_args(p, mode, u_dev, d_dev);
  if (rc) return rc;
  const bool want_K = mode != MODE_PHI_R;
  const int kmode = mode == MODE_PHI_R ? MODE_ELASTIC : mode;
  CK(cudaSetDevice(p->device));
  p->ev_cur = p->prof ? next_events(p) : nullptr;
  if (p->ev_cur) CK(cudaEventRecord(p->ev_cur->e[0], p->stream));
  long long launches = 0;
  // the two-stage API promises element-major staging between the stages; a caller that declared which slot ranges
  // it reads (export ranges / peer links) gets the sorted path, with K1 materialising exactly those ranges
  const bool kmode_ok = want_K && (mode == MODE_ELASTIC || mode == MODE_PF_U);
  const bool sorted_ok = kmode_ok && p->sorted_built && p->sorted_on && (internal || (p->dist_aware && p->n_own_batches == 1));

  const bool sorted = sorted_ok;
  if (p->push_mode && !internal && !sorted)
    return fail(AD4SM_ESTATE, "peer push is set up for Kt evaluations on the destination-sorted path only (modes 0 and 1)");
  p->sorted_last = sorted;
  for (const BatchPlan& b : p->batches) {
    if (b.halo) continue;
    K1Args a = b.k;
    a.u = u_dev;
    a.dfield = d_dev;
    a.hist = (mode == MODE_PF_D && hist_dev) ? hist_dev + b.hist_off * 2 : nullptr;
    a.want_K = want_K;
    a.Ke = p->Ke;
    a.re = p->re;
    a.phie = p->phie;
    // two-stage API of a partitioned run: K1 also writes the interface rows for the owners (peer memory), or -- on the
    // sorted path -- materialises the declared local ranges element-major (the element-major path writes everything so)
    if (p->n_xr && !internal && (sorted || !p->n_xr_local)) {
      if (b.k.slot0 == 0) {
        a.n_xr = p->n_xr;
        for (int x = 0; x < p->n_xr; x++) a.xr[x] = p->xr[x];
      }
    }
    if (p->xmap && sorted && !internal && b.k.slot0 == 0 && !p->push_mode) {  // (push transport: the linked ranges above are what leaves)
      a.xmap = p->xmap;
      a.xke = p->xke;
      a.xre = p->xre;
    }
    if (sorted) {
      a.cpos = p->cpos;
      a.C = p->C;
      static int tma = -1;
      if (tma < 0) {
        const char* e = getenv("AD4SM_SORTED_TMA");
        tma = e ? atoi(e) : 0;  // measured: 36 x 80-byte bulk copies per element make K1 2.74 ms; plain STG.128 1.93 ms
      }
      a.sorted_tma = tma;
      static int k1exp = -1;
      if (k1exp < 0) k1exp = getenv("AD4SM_EXP_K1") ? atoi(getenv("AD4SM_EXP_K1")) : 0;
      a.exp = k1exp;
    }
    // Sub-launches over element slices of the batch: (i) host API -- slice c starts after chunk c of u has arrived;
    // (ii) peer push -- the slices covering the pushed ranges run first, then the copy engine ships them while the
    // rest of K1 runs.
    const bool push = p->push_mode && !internal && sorted && b.k.slot0 == 0;
    if (p->h2d_active || push) {
      long long cut[ad4sm_plan::H2D_MAX + 3];
      int nc = 0;
      cut[nc++] = 0;
      if (p->h2d_active)
        for (int c = 1; c < p->n_h2d; c++) cut[nc++] = p->h2d_elem[c];
      if (push && p->push_end > 0 && p->push_end < a.ne) cut[nc++] = p->push_end;
      cut[nc++] = a.ne;
      std::sort(cut, cut + nc);
      bool pushed = !push;
      for (int k = 0; k + 1 < nc; k++) {
        const long long e0 = cut[k], e1 = cut[k + 1];
        if (p->h2d_active)
          for (int c = 0; c < p->n_h2d; c++)
            if (p->h2d_elem[c] < e1) CK(cudaStreamWaitEvent(p->stream, p->ev_h2d[c], 0));  // every chunk the slice may touch
        if (e1 > e0) {
          K1Args as = a;
          as.ne = e1 - e0;
          as.slot0 = a.slot0 + e0;
          as.nodes = a.nodes + e0 * a.N;
          as.gradN = a.gradN + e0 * (long long)(a.D * a.P * a.N);
          as.wgt = a.wgt + e0 * a.P;
          if (a.Nval) as.Nval = a.Nval + e0 * (long long)(a.P * a.N);
          if (a.hist) as.hist = a.hist + e0 * (long long)(a.P * 2);
          cudaError_t e = launch_k1(as, kmode, p->stream);
          if (e != cudaSuccess) return fail(AD4SM_ENODEV, std::string("K1 launch: ") + cudaGetErrorString(e));
          launches++;
        }
        if (!pushed && e1 >= p->push_end) {  // the pushed ranges are staged: ship them on the copy stream
          pushed = true;
          const long long kw = (long long)n_pair_blocks(p->N) * p->D * p->D, rw = (long long)p->N * p->D;
          CK(cudaEventRecord(p->ev_exports, p->stream));
          CK(cudaStreamWaitEvent(p->copy_stream, p->ev_exports, 0));
          for (int x = 0; x < p->n_push; x++) {
            const ad4sm_plan::PushLink& L = p->push[x];
            const long long row = L.remote[p->parity];
            CK(cudaMemcpyAsync(L.ke + row * kw, p->Ke + L.first * kw, sizeof(double) * L.count * kw, cudaMemcpyDeviceToDevice,
                               p->copy_stream));
            CK(cudaMemcpyAsync(L.re + row * rw, p->re + L.first * rw, sizeof(double) * L.count * rw, cudaMemcpyDeviceToDevice,
                               p->copy_stream));
          }
          CK(cudaEventRecord(p->ev_pushed, p->copy_stream));
        }
      }
      if (push) {
        CK(cudaStreamWaitEvent(p->stream, p->ev_pushed, 0));  // what follows on the stream (the barrier) sees the pushes done
        p->push_pending = true;
      }
      continue;
    }
    if (p->h2d_active)
      for (int c = 0; c < p->n_h2d; c++) CK(cudaStreamWaitEvent(p->stream, p->ev_h2d[c], 0));
    cudaError_t e = launch_k1(a, kmode, p->stream);
    if (e != cudaSuccess) return fail(AD4SM_ENODEV, std::string("K1 launch: ") + cudaGetErrorString(e));
    launches++;
  }
  // ϕ = Σ ϕ_e right behind K1: it needs no interface exchange (HALO slots hold 0), so a partitioned run can ship its
  // partial while the assembly is still to come
  CK(launch_phi_reduce(p->phie, p->ne, p->partials, p->phi_dev, p->stream));
  if (p->ev_cur) CK(cudaEventRecord(p->ev_cur->e[1], p->stream));
  if (p->prof) {
    p->prof_launches[AD4SM_PROF_ELEMENT] += launches;
    p->prof_launches[AD4SM_PROF_REDUCE] += 2;
  }
  p->elements_mode = mode;
  p->split_done = 0;
  p->halo_prepared = false;
  return AD4SM_OK;
}

// part 0: everything (or, after a part-1 call, the remainder); part 1: only the interior (rows / pairs without HALO
// contributions), leaving the evaluation pending.
static int assemble_impl(ad4sm_plan* p, int32_t mode, int part) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  if (mode < 0 || mode > 3) return fail(AD4SM_EINVAL, "mode must be 0..3");
  if (p->elements_mode != mode) return fail(AD4SM_ESTATE, "ad4sm_eval_assemble_device: no element evaluation of this mode pending");
  const int field = mode == MODE_PF_D ? AD4SM_FIELD_D : AD4SM_FIELD_U;
  const bool want_K = mode != MODE_PHI_R;
  CK(cudaSetDevice(p->device));
  EvSet* ev = p->ev_cur;
  const bool splittable = want_K && p->sorted_last && p->n_own_sorted < p->ne && p->split_pair > 0;
  if (part == 1) {
    if (p->split_done) return fail(AD4SM_ESTATE, "ad4sm_eval_assemble_split_device: the interior part has already been assembled");
    if (ev) CK(cudaEventRecord(ev->e[4], p->stream));
    if (splittable) {
      AsmArgs a = asm_args(p, field);
      a.node1 = p->split_node;
      a.pair1 = p->split_pair;
      CK(launch_k3(a, p->stream));
      CK(launch_k2s(a, p->stream));
      if (p->prof) p->prof_launches[AD4SM_PROF_ASSEMBLE] += 2;
    }
    p->split_done = splittable ? 2 : 1;
    return AD4SM_OK;
  }
  const bool rest = p->split_done == 2;  // K3 / K2s of the interior already ran
  if (ev && !p->split_done) CK(cudaEventRecord(ev->e[4], p->stream));  // exchange time (if any) lies between e[1] and e[4]
  // r and ϕ first: the host API copies them back while the (much longer) Kt assembly is still running
  AsmArgs a = asm_args(p, field);
  if (rest) {
    a.node0 = p->split_node;
    a.pair0 = p->split_pair;
    a.keep_zero_count = 1;
  }
  const bool prepared = rest && p->halo_prepared;  // ad4sm_eval_halo_prepare_device ran behind the exchange
  if (!prepared) CK(launch_k3(a, p->stream));
  CK(cudaEventRecord(p->ev_r_ready, p->stream));
  if (want_K) {
    if (p->sorted_last) {
      if (p->n_own_sorted < p->ne && !prepared) {
        const long long kw = (long long)n_pair_blocks(p->N) * p->D * p->D;
        const double* rows = p->Ke + ((p->push_pending && p->parity) ? p->n_halo * kw : 0);  // receive buffer in use (peer push)
        CK(launch_halo_scatter(rows, p->cpos, p->C, p->n_own_sorted, p->ne - p->n_own_sorted, n_pair_blocks(p->N), p->stream));
      }
      CK(launch_k2s(a, p->stream));
    }
    else CK(launch_k2(a, p->stream));
  }
  if (ev) CK(cudaEventRecord(ev->e[2], p->stream));
  if (ev) CK(cudaEventRecord(ev->e[3], p->stream));
  if (p->prof) {
    p->prof_launches[AD4SM_PROF_ASSEMBLE] += want_K ? 2 : 1;
  }
  p->split_done = 0;
  p->halo_prepared = false;
  p->ev_cur = nullptr;
  p->elements_mode = -1;
  if (p->push_pending) {  // the next evaluation's pushes land in the other receive buffer
    p->parity ^= 1;
    p->push_pending = false;
  }
  FieldOut& f = p->field[field];
  f.evaluated = true;
  f.have_K = want_K;
  f.dz_valid = false;
  f.n_zero = -1;
  p->last_field = field;
  return AD4SM_OK;
}
// After ad4sm_eval_assemble_split_device and once the interface rows have arrived: the HALO scatter into the sorted staging
// and the interface rows of r, on `stream` (the stream the exchange ran on) -- off the main stream, which is still busy with
// the interior.  The caller orders `stream` behind the exchange and the main stream behind `stream`.
int ad4sm_eval_halo_prepare_device(ad4sm_plan* p, int32_t mode, void* cuda_stream) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  if (p->elements_mode != mode || mode < 0 || mode > 3) return fail(AD4SM_ESTATE, "ad4sm_eval_halo_prepare_device: no element evaluation of this mode pending");
  if (p->split_done != 2 || p->halo_prepared) return AD4SM_OK;  // nothing was split off (or already done): the assembly does it all
  CK(cudaSetDevice(p->device));
  cudaStream_t s = cuda_stream ? (cudaStream_t)cuda_stream : p->copy_stream;
  const int field = mode == MODE_PF_D ? AD4SM_FIELD_D : AD4SM_FIELD_U;
  AsmArgs a = asm_args(p, field);
  a.node0 = p->split_node;
  CK(launch_k3(a, s));
  if (p->sorted_last && p->n_own_sorted < p->ne) {
    const long long kw = (long long)n_pair_blocks(p->N) * p->D * p->D;
    const double* rows = p->Ke + ((p->push_pending && p->parity) ? p->n_halo * kw : 0);
    CK(launch_halo_scatter(rows, p->cpos, p->C, p->n_own_sorted, p->ne - p->n_own_sorted, n_pair_blocks(p->N), s));
  }
  p->halo_prepared = true;
  return AD4SM_OK;
}
int ad4sm_eval_assemble_device(ad4sm_plan* p, int32_t mode) { return assemble_impl(p, mode, 0); }
int ad4sm_eval_assemble_split_device(ad4sm_plan* p, int32_t mode) { return assemble_impl(p, mode, 1); }

int ad4sm_eval_elements_device(ad4sm_plan* p, int32_t mode, const double* u_dev, const double* d_dev, double* hist_dev) {
  return eval_elements_impl(p, mode, u_dev, d_dev, hist_dev, false);
}

int ad4sm_eval_device(ad4sm_plan* p, int32_t mode, const double* u_dev, const double* d_dev, double* hist_dev) {
  int rc = eval_elements_impl(p, mode, u_dev, d_dev, hist_dev, true);
  if (rc) return rc;
  return ad4sm_eval_assemble_device(p, mode);
}

int ad4sm_staging_view_get(ad4sm_plan* p, int32_t field, ad4sm_staging_view* v) {
  int rc = check_field(p, field);
  if (rc) return rc;
  if (!v) return fail(AD4SM_EINVAL, "view is NULL");
  const int DB = field == AD4SM_FIELD_D ? 1 : p->D;
  v->n_slots = p->ne;
  v->ke_width = (int64_t)n_pair_blocks(p->N) * DB * DB;
  v->re_width = (int64_t)p->N * DB;
  v->Ke_dev = p->Ke;
  v->re_dev = p->re;
  v->phie_dev = p->phie;
  return AD4SM_OK;
}

int ad4sm_set_option(ad4sm_plan* p, int32_t option, int64_t value) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  switch (option) {
    case AD4SM_OPT_SORTED: p->sorted_on = value != 0; return AD4SM_OK;
    case AD4SM_OPT_PEER_PUSH: p->push_opt = value != 0; return AD4SM_OK;
  }
  return fail(AD4SM_EINVAL, "unknown option");
}
int ad4sm_get_info(ad4sm_plan* p, int32_t what, int64_t* value) {
  if (!p || !value) return fail(AD4SM_EINVAL, "plan or value is NULL");
  switch (what) {
    case AD4SM_INFO_SORTED: *value = p->sorted_built ? (p->sorted_on ? 2 : 1) : 0; return AD4SM_OK;
    case AD4SM_INFO_DIST_TRANSPORT: *value = p->dcomm ? (p->push_mode ? 2 : 1) : 0; return AD4SM_OK;
    case AD4SM_INFO_SPLIT_PAIRS: *value = (p->sorted_built && p->n_own_sorted < p->ne) ? p->split_pair : 0; return AD4SM_OK;
  }
  return fail(AD4SM_EINVAL, "unknown info id");
}

int ad4sm_peer_export(ad4sm_plan* p, void* handles_out) {
  if (!p || !handles_out) return fail(AD4SM_EINVAL, "plan or handles_out is NULL");
  CK(cudaSetDevice(p->device));
  cudaIpcMemHandle_t h[2];
  CK(cudaIpcGetMemHandle(&h[0], p->Ke));
  CK(cudaIpcGetMemHandle(&h[1], p->re));
  static_assert(sizeof(h) == AD4SM_PEER_HANDLE_BYTES, "handle size");
  memcpy(handles_out, h, sizeof(h));
  return AD4SM_OK;
}

int ad4sm_peer_attach(ad4sm_plan* p, int32_t n_links, const ad4sm_peer_link* links) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  if (n_links < 0 || n_links > AD4SM_MAX_PEER_LINKS || (n_links && !links)) return fail(AD4SM_EINVAL, "bad peer links");
  CK(cudaSetDevice(p->device));
  const long long kw = (long long)n_pair_blocks(p->N) * p->D * p->D, rw = (long long)p->N * p->D;
  p->n_xr = 0;
  p->n_push = 0;
  p->push_mode = false;
  if (p->push_opt) {  // copy-engine pushes of locally materialised ranges (header: AD4SM_OPT_PEER_PUSH)
    if (!p->sorted_built || p->n_own_batches != 1)
      return fail(AD4SM_ESTATE, "AD4SM_OPT_PEER_PUSH needs a plan with destination-sorted staging and a single evaluated batch");
    if (!p->ev_exports) {
      CK(cudaEventCreateWithFlags(&p->ev_exports, cudaEventDisableTiming));
      CK(cudaEventCreateWithFlags(&p->ev_pushed, cudaEventDisableTiming));
    }
    long long end = 0;
    for (int i = 0; i < n_links; i++) {
      const ad4sm_peer_link& L = links[i];
      if (L.first_slot < 0 || L.count <= 0 || L.first_slot + L.count > p->n_own_sorted || !L.peer_handles || L.remote_first_slot_b <= L.remote_first_slot)
        return fail(AD4SM_EINVAL, "peer link range outside the plan's own elements");
      cudaIpcMemHandle_t h[2];
      memcpy(h, L.peer_handles, sizeof(h));
      void *ke = nullptr, *re = nullptr;
      CK(cudaIpcOpenMemHandle(&ke, h[0], cudaIpcMemLazyEnablePeerAccess));
      p->ipc_open.push_back(ke);
      CK(cudaIpcOpenMemHandle(&re, h[1], cudaIpcMemLazyEnablePeerAccess));
      p->ipc_open.push_back(re);
      p->push[i] = {L.first_slot, L.count, {L.remote_first_slot, L.remote_first_slot_b}, (double*)ke, (double*)re};
      // the range is materialised element-major in local memory by K1 (like ad4sm_export_ranges)
      p->xr[i].first = L.first_slot;
      p->xr[i].count = L.count;
      p->xr[i].ke = p->Ke + L.first_slot * kw;
      p->xr[i].re = p->re + L.first_slot * rw;
      end = std::max<long long>(end, L.first_slot + L.count);
    }
    p->push_mode = true;
    p->n_push = n_links;
    p->n_xr = n_links;
    p->n_xr_local = n_links;
    p->push_end = std::min<long long>(p->n_own_sorted, (end + 31) & ~31ll);
    {  // the first K1 slice ends on a whole wave of the persistent kernel (all warps of all SMs busy to its end)
      K1Config c;
      const BatchPlan& b0 = p->batches[0];
      if (k1_config(p->D, p->N, b0.k.P, false, b0.k.ne, &c) && c.persistent) {
        const long long wave = (long long)c.grid * c.warps * c.EW;
        if (wave > 0) p->push_end = std::min<long long>(p->n_own_sorted, (p->push_end + wave - 1) / wave * wave);
      }
    }
    p->parity = 0;
    p->push_pending = false;
    p->dist_aware = true;
    return AD4SM_OK;
  }
  for (int i = 0; i < n_links; i++) {
    const ad4sm_peer_link& L = links[i];
    if (L.first_slot < 0 || L.count <= 0 || L.first_slot + L.count > p->ne || !L.peer_handles)
      return fail(AD4SM_EINVAL, "peer link range outside the plan");
    cudaIpcMemHandle_t h[2];
    memcpy(h, L.peer_handles, sizeof(h));
    void *ke = nullptr, *re = nullptr;
    CK(cudaIpcOpenMemHandle(&ke, h[0], cudaIpcMemLazyEnablePeerAccess));
    p->ipc_open.push_back(ke);
    CK(cudaIpcOpenMemHandle(&re, h[1], cudaIpcMemLazyEnablePeerAccess));
    p->ipc_open.push_back(re);
    p->xr[i].first = L.first_slot;
    p->xr[i].count = L.count;
    p->xr[i].ke = (double*)ke + L.remote_first_slot * kw;
    p->xr[i].re = (double*)re + L.remote_first_slot * rw;
  }
  p->n_xr = n_links;
  p->n_xr_local = 0;
  p->dist_aware = true;
  return AD4SM_OK;
}

int ad4sm_export_ranges(ad4sm_plan* p, int32_t n_ranges, const int64_t* first_slot, const int64_t* count) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  if (n_ranges < 0 || n_ranges > AD4SM_MAX_PEER_LINKS || (n_ranges && (!first_slot || !count)))
    return fail(AD4SM_EINVAL, "bad export ranges");
  const long long kw = (long long)n_pair_blocks(p->N) * p->D * p->D, rw = (long long)p->N * p->D;
  for (int i = 0; i < n_ranges; i++) {
    if (first_slot[i] < 0 || count[i] <= 0 || first_slot[i] + count[i] > p->ne) return fail(AD4SM_EINVAL, "export range outside the plan");
    p->xr[i].first = first_slot[i];
    p->xr[i].count = count[i];
    p->xr[i].ke = p->Ke + first_slot[i] * kw;
    p->xr[i].re = p->re + first_slot[i] * rw;
  }
  p->n_xr = n_ranges;
  p->n_xr_local = n_ranges;
  p->dist_aware = true;
  return AD4SM_OK;
}

int ad4sm_sync(ad4sm_plan* p) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  CK(cudaSetDevice(p->device));
  CK(cudaStreamSynchronize(p->stream));
  if (p->copy_stream) CK(cudaStreamSynchronize(p->copy_stream));
  return AD4SM_OK;
}

int ad4sm_device_view_get(ad4sm_plan* p, int32_t field, ad4sm_device_view* v) {
  int rc = check_field(p, field);
  if (rc) return rc;
  if (!v) return fail(AD4SM_EINVAL, "view is NULL");
  rc = ensure_pattern(p, field);
  if (rc) return rc;
  FieldOut& f = p->field[field];
  v->n_dof = f.n_dof;
  v->nnz = f.nnz;
  v->phi_dev = p->phi_dev;
  v->r_dev = f.r;
  v->nzval_dev = f.nzval;
  v->colptr_dev = (const int64_t*)f.colptr;
  v->rowval_dev = (const int64_t*)f.rowval;
  v->n_zero_dev = (const int64_t*)p->n_zero_dev;
  return AD4SM_OK;
}

// host-pointer evaluation shared by the four reference-facing entry points
// Host inputs -> device, asynchronously.  With a pipelined plan (n_h2d > 0) u and d travel in node chunks on the copy
// stream, each followed by an event that the K1 launch of the matching element slice waits for (h2d_active).
static int upload_inputs(ad4sm_plan* p, int mode, const double* u, const double* d, double* const* hist) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  if (mode < 0 || mode > 3) return fail(AD4SM_EINVAL, "mode must be 0..3");
  if (!u) return fail(AD4SM_EINVAL, "u is NULL");
  CK(cudaSetDevice(p->device));
  cudaStream_t s = p->stream;
  const bool needs_d = (mode == MODE_PF_U || mode == MODE_PF_D);
  if (needs_d) {
    if (!d) return fail(AD4SM_EINVAL, "d is NULL");
    if (!p->has_cp) return fail(AD4SM_EINVAL, "phase-field evaluation needs CPElem batches");
  }
  if (hist)
    for (size_t b = 0; b < p->batches.size(); b++)
      if (!hist[b] && !p->batches[b].halo) return fail(AD4SM_EINVAL, "hist pointer of a batch is NULL");
  if (p->n_h2d > 0) {  // node chunks on the copy stream; K1 slice c waits for ev_h2d[c]
    CK(cudaEventRecord(p->ev_start, s));
    CK(cudaStreamWaitEvent(p->copy_stream, p->ev_start, 0));
    for (int c = 0; c < p->n_h2d; c++) {
      const long long n0 = p->h2d_node[c], n1 = p->h2d_node[c + 1];
      if (n1 > n0) {
        CK(cudaMemcpyAsync(p->u_dev + n0 * p->dpn, u + n0 * p->dpn, sizeof(double) * (n1 - n0) * p->dpn, cudaMemcpyHostToDevice,
                           p->copy_stream));
        if (needs_d) CK(cudaMemcpyAsync(p->d_dev + n0, d + n0, sizeof(double) * (n1 - n0), cudaMemcpyHostToDevice, p->copy_stream));
      }
      CK(cudaEventRecord(p->ev_h2d[c], p->copy_stream));
    }
  } else {
    CK(cudaMemcpyAsync(p->u_dev, u, sizeof(double) * p->n_nodes * p->dpn, cudaMemcpyHostToDevice, s));
    if (needs_d) CK(cudaMemcpyAsync(p->d_dev, d, sizeof(double) * p->n_nodes, cudaMemcpyHostToDevice, s));
  }
  if (hist)
    for (const BatchPlan& b : p->batches) {
      if (b.halo) continue;
      const double* h = hist[&b - p->batches.data()];
      CK(cudaMemcpyAsync(p->hist_dev + b.hist_off * 2, h, sizeof(double) * 2 * b.k.ne * b.k.P, cudaMemcpyHostToDevice, s));
    }
  return AD4SM_OK;
}
// ϕ and r of the last assembly to the host on the copy stream, as soon as K3 and the ϕ reduction are done (the Kt
// assembly keeps running on the main stream); the main stream then waits for the copy, so that a following evaluation
// cannot overwrite r before it has left.
static int download_phi_r(ad4sm_plan* p, int field, double* phi, double* r) {
  FieldOut& f = p->field[field];
  CK(cudaStreamWaitEvent(p->copy_stream, p->ev_r_ready, 0));
  if (phi) CK(cudaMemcpyAsync(phi, p->phi_dev, sizeof(double), cudaMemcpyDeviceToHost, p->copy_stream));
  if (r) CK(cudaMemcpyAsync(r, f.r, sizeof(double) * f.n_dof, cudaMemcpyDeviceToHost, p->copy_stream));
  CK(cudaEventRecord(p->ev_copied, p->copy_stream));
  CK(cudaStreamWaitEvent(p->stream, p->ev_copied, 0));
  return AD4SM_OK;
}

static int eval_host(ad4sm_plan* p, int mode, const double* u, const double* d, double* const* hist, double* phi,
                     double* r, int64_t* nnz_out) {
  int rc = upload_inputs(p, mode, u, d, hist);
  if (rc) return rc;
  cudaStream_t s = p->stream;
  const bool needs_d = (mode == MODE_PF_U || mode == MODE_PF_D);
  p->h2d_active = p->n_h2d > 0;
  rc = ad4sm_eval_device(p, mode, p->u_dev, needs_d ? p->d_dev : nullptr, hist ? p->hist_dev : nullptr);
  p->h2d_active = false;
  if (rc) {
    if (p->n_h2d > 0) cudaStreamSynchronize(p->copy_stream);  // the caller's buffers must not be in flight on return
    return rc;
  }
  const int field = mode == MODE_PF_D ? AD4SM_FIELD_D : AD4SM_FIELD_U;
  FieldOut& f = p->field[field];
  unsigned long long nz = 0;
  rc = download_phi_r(p, field, phi, r);
  if (rc) return rc;
  if (mode != MODE_PHI_R) CK(cudaMemcpyAsync(&nz, p->n_zero_dev, sizeof(nz), cudaMemcpyDeviceToHost, s));
  if (hist)
    for (const BatchPlan& b : p->batches) {
      if (b.halo) continue;
      double* h = hist[&b - p->batches.data()];
      CK(cudaMemcpyAsync(h, p->hist_dev + b.hist_off * 2, sizeof(double) * 2 * b.k.ne * b.k.P, cudaMemcpyDeviceToHost, s));
    }
  CK(cudaStreamSynchronize(s));
  if (mode != MODE_PHI_R) {
    f.n_zero = (long long)nz;
    if (nnz_out) *nnz_out = f.nnz - f.n_zero;
  }
  return AD4SM_OK;
}

// mesh-partitioned runs with host-resident inputs / outputs: the two ends of the two-stage API
int ad4sm_eval_elements_host(ad4sm_plan* p, int32_t mode, const double* u, const double* d, double* const* hist) {
  int rc = upload_inputs(p, mode, u, d, hist);
  if (rc) return rc;
  const bool needs_d = (mode == MODE_PF_U || mode == MODE_PF_D);
  p->h2d_active = p->n_h2d > 0;
  rc = eval_elements_impl(p, mode, p->u_dev, needs_d ? p->d_dev : nullptr, hist ? p->hist_dev : nullptr, false);
  p->h2d_active = false;
  if (rc && p->n_h2d > 0) cudaStreamSynchronize(p->copy_stream);
  return rc;
}
int ad4sm_fetch_phi_r_async(ad4sm_plan* p, int32_t field, double* phi, double* r) {
  int rc = check_field(p, field);
  if (rc) return rc;
  if (!p->field[field].evaluated) return fail(AD4SM_ESTATE, "ad4sm_fetch_phi_r_async before any evaluation of this field");
  CK(cudaSetDevice(p->device));
  return download_phi_r(p, field, phi, r);
}

int ad4sm_eval(ad4sm_plan* p, const double* u, double* phi, double* r, int64_t* nnz_out) {
  return eval_host(p, MODE_ELASTIC, u, nullptr, nullptr, phi, r, nnz_out);
}
int ad4sm_eval_phi_r(ad4sm_plan* p, const double* u, double* phi, double* r) {
  return eval_host(p, MODE_PHI_R, u, nullptr, nullptr, phi, r, nullptr);
}
int ad4sm_eval_pf_u(ad4sm_plan* p, const double* u, const double* d, double* phi, double* r, int64_t* nnz_out) {
  return eval_host(p, MODE_PF_U, u, d, nullptr, phi, r, nnz_out);
}
int ad4sm_eval_pf_d(ad4sm_plan* p, const double* u, const double* d, double* const* hist, double* phi, double* r,
                    int64_t* nnz_out) {
  return eval_host(p, MODE_PF_D, u, d, hist, phi, r, nnz_out);
}

// makeϕrKt(Φ, elems, u) (elements.toolkit.jl:169-203): assembly of caller-supplied element duals
static int assemble_duals_impl(ad4sm_plan* p, const void* Phi, int64_t stride, bool want_K, double* phi, double* r, int64_t* nnz_out) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  if (!Phi) return fail(AD4SM_EINVAL, "Phi is NULL");
  const long long n = (long long)p->N * p->D, m = n * (n + 1) / 2;
  if (p->dpn != p->D) return fail(AD4SM_EINVAL, "assembly of element duals needs size(u,1) == element dimension");
  if (stride < 1 + n + (want_K ? m : 0))
    return fail(AD4SM_EINVAL, want_K ? "stride smaller than 1 + n + n(n+1)/2 doubles (D2{n,m,Float64})"
                                     : "stride smaller than 1 + n doubles (D1{n,Float64})");
  for (const BatchPlan& b : p->batches)
    if (b.halo) return fail(AD4SM_EINVAL, "ad4sm_assemble_duals on a plan with HALO batches");
  CK(cudaSetDevice(p->device));
  cudaStream_t s = p->stream;
  double* dPhi = nullptr;
  const size_t bytes = sizeof(double) * (size_t)p->ne * (size_t)stride;
  cudaError_t e = cudaMalloc(&dPhi, bytes);
  if (e != cudaSuccess) return fail(AD4SM_ENOMEM, std::string("cudaMalloc of the element duals: ") + cudaGetErrorString(e));
  auto done = [&](int rc) {
    cudaStreamSynchronize(s);
    cudaFree(dPhi);
    return rc;
  };
  if (cudaMemcpyAsync(dPhi, Phi, bytes, cudaMemcpyHostToDevice, s) != cudaSuccess) return done(fail(AD4SM_ENODEV, "upload of the element duals failed"));
  if (launch_duals_to_staging(dPhi, stride, p->key_dev, p->ne, p->N, p->D, want_K ? p->Ke : nullptr, p->re, p->phie, s) != cudaSuccess)
    return done(fail(AD4SM_ENODEV, "duals_to_staging launch failed"));
  const bool sorted = want_K && p->sorted_built && p->sorted_on;
  if (sorted && launch_halo_scatter(p->Ke, p->cpos, p->C, 0, p->n_own_sorted, n_pair_blocks(p->N), s) != cudaSuccess)
    return done(fail(AD4SM_ENODEV, "staging scatter launch failed"));
  if (launch_phi_reduce(p->phie, p->ne, p->partials, p->phi_dev, s) != cudaSuccess)  // ϕ = Σ v (toolkit:186)
    return done(fail(AD4SM_ENODEV, "ϕ reduction launch failed"));
  p->ev_cur = nullptr;
  p->sorted_last = sorted;
  const int mode = want_K ? MODE_ELASTIC : MODE_PHI_R;
  p->elements_mode = mode;
  p->split_done = 0;
  int rc = ad4sm_eval_assemble_device(p, mode);
  if (rc) return done(rc);
  FieldOut& f = p->field[AD4SM_FIELD_U];
  unsigned long long nz = 0;
  rc = download_phi_r(p, AD4SM_FIELD_U, phi, r);
  if (rc) return done(rc);
  if (want_K && cudaMemcpyAsync(&nz, p->n_zero_dev, sizeof(nz), cudaMemcpyDeviceToHost, s) != cudaSuccess)
    return done(fail(AD4SM_ENODEV, "copy failed"));
  if (cudaStreamSynchronize(s) != cudaSuccess) return done(fail(AD4SM_ENODEV, "assembly failed"));
  cudaFree(dPhi);
  if (want_K) {
    f.n_zero = (long long)nz;
    if (nnz_out) *nnz_out = f.nnz - f.n_zero;
  }
  return AD4SM_OK;
}
int ad4sm_assemble_duals(ad4sm_plan* p, const void* Phi, int64_t stride, double* phi, double* r, int64_t* nnz_out) {
  return assemble_duals_impl(p, Phi, stride, true, phi, r, nnz_out);
}
// makeϕr(Φ, elems, u) (elements.toolkit.jl:204-215): records (v | g[n] | ...), D1 or D2
int ad4sm_assemble_duals_phi_r(ad4sm_plan* p, const void* Phi, int64_t stride, double* phi, double* r) {
  return assemble_duals_impl(p, Phi, stride, false, phi, r, nullptr);
}

int ad4sm_fetch_csc(ad4sm_plan* p, int64_t* colptr, int64_t* rowval, double* nzval) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  if (p->last_field < 0) return fail(AD4SM_ESTATE, "fetch before any evaluation");
  const int field = p->last_field;
  FieldOut& f = p->field[field];
  if (!f.have_K) return fail(AD4SM_ESTATE, "the last evaluation computed ϕ and r only");
  int rc = ensure_pattern(p, field);
  if (rc) return rc;
  CK(cudaSetDevice(p->device));
  if (f.n_zero < 0) {  // device-path evaluation: read the zero count now
    unsigned long long nz = 0;
    CK(cudaStreamSynchronize(p->stream));
    CK(cudaMemcpy(&nz, p->n_zero_dev, sizeof(nz), cudaMemcpyDeviceToHost));
    f.n_zero = (long long)nz;
  }
  if (f.n_zero == 0) {  // nothing to drop: the structural pattern is the result
    if (colptr) CK(cudaMemcpy(colptr, f.colptr, sizeof(long long) * (f.n_dof + 1), cudaMemcpyDeviceToHost));
    if (rowval) CK(cudaMemcpy(rowval, f.rowval, sizeof(long long) * f.nnz, cudaMemcpyDeviceToHost));
    if (nzval) CK(cudaMemcpy(nzval, f.nzval, sizeof(double) * f.nnz, cudaMemcpyDeviceToHost));
    return AD4SM_OK;
  }
  const long long nnz_out = f.nnz - f.n_zero;
  if (!f.dz_valid) {
    if (!f.dz_colptr) CK(cudaMalloc((void**)&f.dz_colptr, sizeof(long long) * (f.n_dof + 1)));
    if (f.dz_cap < nnz_out || !f.dz_rowval) {
      cudaFree(f.dz_rowval);
      cudaFree(f.dz_nzval);
      f.dz_rowval = nullptr;
      f.dz_nzval = nullptr;
      const long long cap = std::max<long long>(nnz_out, 1);
      CK(cudaMalloc((void**)&f.dz_rowval, sizeof(long long) * cap));
      CK(cudaMalloc((void**)&f.dz_nzval, sizeof(double) * cap));
      f.dz_cap = cap;
    }
    void* tmp = nullptr;
    const size_t tb = dropzeros_tmp_bytes(f.n_dof);
    CK(cudaMalloc(&tmp, tb ? tb : 1));
    cudaError_t e = launch_dropzeros(f.n_dof, f.colptr, f.rowval, f.nzval, f.dz_colptr, f.dz_rowval, f.dz_nzval, tmp, tb,
                                     p->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
    cudaFree(tmp);
    CK(e);
    f.dz_valid = true;
  }
  if (colptr) CK(cudaMemcpy(colptr, f.dz_colptr, sizeof(long long) * (f.n_dof + 1), cudaMemcpyDeviceToHost));
  if (rowval) CK(cudaMemcpy(rowval, f.dz_rowval, sizeof(long long) * nnz_out, cudaMemcpyDeviceToHost));
  if (nzval) CK(cudaMemcpy(nzval, f.dz_nzval, sizeof(double) * nnz_out, cudaMemcpyDeviceToHost));
  return AD4SM_OK;
}

// ---- Newton-step glue (solvers.jl:183-184, 211-224, 257-291; kernels in newton.cu) ------------------------------------
static NewtonArgs newton_args(ad4sm_plan* p, int field) {
  FieldOut& f = p->field[field];
  NewtonArgs a = {};
  a.n_nodes = p->n_nodes;
  a.nfree = f.nw.nfree;
  a.dpn = f.dpn;
  a.nbr_ptr = p->nbr_ptr;
  a.nbr = p->nbr;
  a.fidx = f.nw.fidx;
  a.free_list = f.nw.free_list;
  a.nzval = f.nzval;
  return a;
}

int ad4sm_newton_setup(ad4sm_plan* p, int32_t field, const uint8_t* bfree, int64_t* nfree_out, int64_t* ncnst_out, int64_t* nnz_ff_out) {
  int rc = check_field(p, field);
  if (rc) return rc;
  if (!bfree) return fail(AD4SM_EINVAL, "bfree is NULL");
  FieldOut& f = p->field[field];
  if (f.n_dof >= (1ll << 31)) return fail(AD4SM_EINVAL, "too many dofs for the 32-bit free/constrained index");
  if (f.dpn > 8) return fail(AD4SM_EINVAL, "Newton glue supports up to 8 dofs per node");
  CK(cudaSetDevice(p->device));
  FieldOut::Newton& w = f.nw;
  // ifreeu = findall(bfreeu[:]), icnstu = findall(.!bfreeu[:])   (solvers.jl:183-184)
  std::vector<int> fidx(f.n_dof), fl, cl;
  for (long long k = 0; k < f.n_dof; k++) {
    if (bfree[k]) {
      fidx[k] = (int)fl.size();
      fl.push_back((int)k);
    } else {
      fidx[k] = -1 - (int)cl.size();
      cl.push_back((int)k);
    }
  }
  for (void* q : {(void*)w.fidx, (void*)w.free_list, (void*)w.cnst_list, (void*)w.colptr_ff, (void*)w.rowval_ff, (void*)w.nzval_ff, (void*)w.res,
                  (void*)w.fc, (void*)w.fe, (void*)w.du, (void*)w.partials, (void*)w.norms})
    if (q) {  // a previous split: release it
      cudaFree(q);
      p->allocs.erase(std::remove(p->allocs.begin(), p->allocs.end(), q), p->allocs.end());
    }
  w = FieldOut::Newton();
  w.nfree = (long long)fl.size();
  w.ncnst = (long long)cl.size();
  rc = dev_upload(p, &w.fidx, fidx.data(), fidx.size());
  if (!rc) rc = dev_upload(p, &w.free_list, fl.data(), fl.size());
  if (!rc) rc = dev_upload(p, &w.cnst_list, cl.data(), cl.size());
  if (!rc) rc = dev_alloc(p, &w.colptr_ff, (size_t)w.nfree + 1);
  if (!rc) rc = dev_alloc(p, &w.res, (size_t)w.nfree);
  if (!rc) rc = dev_alloc(p, &w.fc, (size_t)w.nfree);
  if (!rc) rc = dev_alloc(p, &w.fe, (size_t)f.n_dof);
  if (!rc) rc = dev_alloc(p, &w.du, (size_t)w.ncnst);
  if (!rc) rc = dev_alloc(p, &w.partials, (size_t)2 * newton_partials());
  if (!rc) rc = dev_alloc(p, &w.norms, 2);
  if (rc) return rc;
  // column counts of Kt[ifree,ifree] -> 1-based column pointers
  CK(cudaMemsetAsync(w.colptr_ff, 0, sizeof(long long) * (w.nfree + 1), p->stream));
  NewtonArgs a = newton_args(p, field);
  a.cnt = w.colptr_ff;
  CK(launch_newton_columns(a, p->stream));
  void* tmp = nullptr;
  const size_t tb = newton_scan_tmp_bytes(w.nfree + 1);
  CK(cudaMalloc(&tmp, tb ? tb : 1));
  cudaError_t e = launch_newton_colptr(w.colptr_ff, w.nfree, tmp, tb, p->stream);
  long long last = 1;
  if (e == cudaSuccess) e = cudaMemcpyAsync(&last, w.colptr_ff + w.nfree, sizeof(long long), cudaMemcpyDeviceToHost, p->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(p->stream);
  cudaFree(tmp);
  CK(e);
  w.nnz_ff = last - 1;
  rc = dev_alloc(p, &w.nzval_ff, (size_t)std::max<long long>(w.nnz_ff, 1));
  if (rc) return rc;
  w.ready = true;
  if (nfree_out) *nfree_out = w.nfree;
  if (ncnst_out) *ncnst_out = w.ncnst;
  if (nnz_ff_out) *nnz_ff_out = w.nnz_ff;
  return AD4SM_OK;
}

static int newton_ensure_rowval(ad4sm_plan* p, int field) {
  FieldOut::Newton& w = p->field[field].nw;
  if (w.rowval_ff) return 0;
  int rc = dev_alloc(p, &w.rowval_ff, (size_t)std::max<long long>(w.nnz_ff, 1));
  if (rc) return rc;
  NewtonArgs a = newton_args(p, field);
  a.colptr_ff = w.colptr_ff;
  a.rowval_ff = w.rowval_ff;
  CK(launch_newton_columns(a, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return 0;
}

int ad4sm_newton_pattern(ad4sm_plan* p, int32_t field, int64_t* colptr_ff, int64_t* rowval_ff) {
  int rc = check_field(p, field);
  if (rc) return rc;
  FieldOut::Newton& w = p->field[field].nw;
  if (!w.ready) return fail(AD4SM_ESTATE, "ad4sm_newton_pattern before ad4sm_newton_setup");
  CK(cudaSetDevice(p->device));
  if (colptr_ff) CK(cudaMemcpy(colptr_ff, w.colptr_ff, sizeof(long long) * (w.nfree + 1), cudaMemcpyDeviceToHost));
  if (rowval_ff) {
    rc = newton_ensure_rowval(p, field);
    if (rc) return rc;
    CK(cudaMemcpy(rowval_ff, w.rowval_ff, sizeof(long long) * w.nnz_ff, cudaMemcpyDeviceToHost));
  }
  return AD4SM_OK;
}

int ad4sm_newton_reduce_device(ad4sm_plan* p, int32_t field, int32_t kind, const double* fe_dev, const double* ducnst_dev) {
  int rc = check_field(p, field);
  if (rc) return rc;
  FieldOut& f = p->field[field];
  FieldOut::Newton& w = f.nw;
  if (!w.ready) return fail(AD4SM_ESTATE, "ad4sm_newton_reduce before ad4sm_newton_setup");
  if (!f.evaluated || !f.have_K) return fail(AD4SM_ESTATE, "ad4sm_newton_reduce needs a Kt evaluation of this field first");
  if (kind != 0 && kind != 1) return fail(AD4SM_EINVAL, "kind must be 0 (predictor) or 1 (corrector)");
  CK(cudaSetDevice(p->device));
  NewtonArgs a = newton_args(p, field);
  a.colptr_ff = w.colptr_ff;
  a.nzval_ff = w.nzval_ff;
  const bool want_fc = kind == 0 && ducnst_dev && w.ncnst > 0;
  if (want_fc) {
    a.du = ducnst_dev;
    a.fc = w.fc;
  }
  CK(launch_newton_columns(a, p->stream));
  CK(launch_newton_res(w.nfree, w.free_list, f.r, fe_dev, want_fc ? w.fc : nullptr, kind, w.res, w.ncnst, w.cnst_list, w.partials,
                       w.norms, p->stream));
  w.reduced = true;
  return AD4SM_OK;
}

int ad4sm_newton_reduce(ad4sm_plan* p, int32_t field, int32_t kind, const double* fe, const double* ducnst, double* res,
                        double* nzval_ff, double* norms) {
  int rc = check_field(p, field);
  if (rc) return rc;
  FieldOut& f = p->field[field];
  FieldOut::Newton& w = f.nw;
  if (!w.ready) return fail(AD4SM_ESTATE, "ad4sm_newton_reduce before ad4sm_newton_setup");
  CK(cudaSetDevice(p->device));
  cudaStream_t s = p->stream;
  if (fe) CK(cudaMemcpyAsync(w.fe, fe, sizeof(double) * f.n_dof, cudaMemcpyHostToDevice, s));
  if (kind == 0 && ducnst && w.ncnst > 0) CK(cudaMemcpyAsync(w.du, ducnst, sizeof(double) * w.ncnst, cudaMemcpyHostToDevice, s));
  rc = ad4sm_newton_reduce_device(p, field, kind, fe ? w.fe : nullptr, (kind == 0 && ducnst) ? w.du : nullptr);
  if (rc) return rc;
  if (res) CK(cudaMemcpyAsync(res, w.res, sizeof(double) * w.nfree, cudaMemcpyDeviceToHost, s));
  if (norms) CK(cudaMemcpyAsync(norms, w.norms, sizeof(double) * 2, cudaMemcpyDeviceToHost, s));
  if (nzval_ff) CK(cudaMemcpyAsync(nzval_ff, w.nzval_ff, sizeof(double) * w.nnz_ff, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return AD4SM_OK;
}

int ad4sm_newton_view_get(ad4sm_plan* p, int32_t field, ad4sm_newton_view* v) {
  int rc = check_field(p, field);
  if (rc) return rc;
  if (!v) return fail(AD4SM_EINVAL, "view is NULL");
  FieldOut::Newton& w = p->field[field].nw;
  if (!w.ready) return fail(AD4SM_ESTATE, "ad4sm_newton_view_get before ad4sm_newton_setup");
  CK(cudaSetDevice(p->device));
  rc = newton_ensure_rowval(p, field);
  if (rc) return rc;
  v->nfree = w.nfree;
  v->ncnst = w.ncnst;
  v->nnz_ff = w.nnz_ff;
  v->res_dev = w.res;
  v->nzval_ff_dev = w.nzval_ff;
  v->norms_dev = w.norms;
  v->colptr_ff_dev = (const int64_t*)w.colptr_ff;
  v->rowval_ff_dev = (const int64_t*)w.rowval_ff;
  return AD4SM_OK;
}

// ---- the partitioned evaluation behind the C ABI (NCCL inside the library; plumbing in dist.cu) ------------------------
int ad4sm_dist_unique_id(void* id_out) {
  if (!id_out) return fail(AD4SM_EINVAL, "id_out is NULL");
  std::string err;
  if (dist_unique_id(id_out, &err)) return fail(AD4SM_ENCCL, err);
  return AD4SM_OK;
}

int ad4sm_dist_init(int32_t nranks, int32_t rank, const void* nccl_id, int32_t device, ad4sm_dist** comm_out) {
  if (!comm_out) return fail(AD4SM_EINVAL, "comm_out is NULL");
  *comm_out = nullptr;
  if (nranks < 1 || rank < 0 || rank >= nranks || !nccl_id) return fail(AD4SM_EINVAL, "bad rank / nranks / id");
  int dev = device;
  if (dev < 0) CK(cudaGetDevice(&dev));
  CK(cudaSetDevice(dev));
  std::string err;
  DistComm* c = dist_comm_init(nranks, rank, nccl_id, dev, &err);
  if (!c) return fail(AD4SM_ENCCL, err);
  *comm_out = new ad4sm_dist{c, dev};
  return AD4SM_OK;
}

void ad4sm_dist_destroy(ad4sm_dist* comm) {
  if (!comm) return;
  cudaSetDevice(comm->device);
  dist_comm_destroy(comm->c);
  delete comm;
}

// Copy-engine push transport for the in-library partitioned evaluation: when EVERY rank's send lists are contiguous slot
// ranges (z-slabs and other contiguous partitions), the plans are destination-sorted, and CUDA IPC works between the ranks,
// the interface rows travel like in dist.py's default transport -- K1 evaluates them first, the copy engine pushes them into
// the owner's double-buffered HALO rows over NVLink while the rest of K1 runs -- and ncclSend/ncclRecv (a few channels per
// peer: ~0.14 ms per 28 MB interface, measured) is not used at all; the ϕ all-gather is the barrier.  Collective: every
// rank all-gathers (IPC handles, n_own, n_slots, its receive ranges, "I can push") and all take the same decision.
// AD4SM_DIST_PUSH=0 keeps the send/recv exchange.
static int dist_try_push(ad4sm_plan* p, ad4sm_dist* comm, int n_peers, const ad4sm_dist_peer* peers) {
  const int nranks = dist_comm_size(comm->c), me = dist_comm_rank(comm->c);
  constexpr int REC = 32;  // 8-byte words per rank: [0,16) handles, 16 n_own, 17 n_slots, 18 can_push, 19 n_entries, 20.. (rank, first, count) x 4
  const char* env = getenv("AD4SM_DIST_PUSH");
  bool can = !(env && atoi(env) == 0) && p->sorted_built && p->n_own_batches == 1 && n_peers <= AD4SM_MAX_PEER_LINKS;
  for (const ad4sm_plan::DistPeer& d : p->dpeers)
    if (d.n_send && d.first_s < 0) can = false;
  std::vector<long long> mine(REC, 0), all((size_t)REC * nranks, 0);
  if (can) {
    cudaIpcMemHandle_t h[2];
    if (cudaIpcGetMemHandle(&h[0], p->Ke) != cudaSuccess || cudaIpcGetMemHandle(&h[1], p->re) != cudaSuccess) {
      cudaGetLastError();
      can = false;
    } else {
      memcpy(mine.data(), h, sizeof(h));
    }
  }
  mine[16] = p->batches[0].k.ne;
  mine[17] = p->ne;
  mine[18] = can ? 1 : 0;
  int ne_ = 0;
  for (int i = 0; i < n_peers && ne_ < 4; i++)
    if (peers[i].n_recv > 0) {
      mine[20 + 3 * ne_] = peers[i].rank;
      mine[21 + 3 * ne_] = peers[i].recv_first_slot;
      mine[22 + 3 * ne_] = peers[i].n_recv;
      ne_++;
    }
  mine[19] = ne_;
  double *d_mine = nullptr, *d_all = nullptr;
  std::string err;
  auto gather = [&]() -> int {  // all ranks' records (a tiny collective on the plan's stream)
    if (!d_mine) {
      CK(cudaMalloc(&d_mine, REC * 8));
      CK(cudaMalloc(&d_all, (size_t)REC * 8 * nranks));
    }
    CK(cudaMemcpyAsync(d_mine, mine.data(), REC * 8, cudaMemcpyHostToDevice, p->stream));
    if (dist_allgather(comm->c, d_mine, d_all, REC, p->stream, &err)) return fail(AD4SM_ENCCL, err);
    CK(cudaMemcpyAsync(all.data(), d_all, (size_t)REC * 8 * nranks, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
  };
  auto done = [&](int rc) {
    cudaFree(d_mine);
    cudaFree(d_all);
    return rc;
  };
  int rc = gather();
  if (rc) return done(rc);
  bool everyone = true;
  for (int g = 0; g < nranks; g++) everyone = everyone && all[(size_t)REC * g + 18] != 0;
  if (!everyone) return done(AD4SM_OK);  // send/recv exchange
  // links: my contiguous send ranges -> the receive ranges the owners announced for me
  ad4sm_peer_link links[AD4SM_MAX_PEER_LINKS];
  int nl = 0;
  bool ok = true;
  for (const ad4sm_plan::DistPeer& d : p->dpeers) {
    if (!d.n_send) continue;
    const long long* R = all.data() + (size_t)REC * d.rank;
    long long first = -1;
    for (int k = 0; k < (int)R[19]; k++)
      if (R[20 + 3 * k] == me && R[22 + 3 * k] == d.n_send) first = R[21 + 3 * k];
    if (first < 0) {
      ok = false;
      break;
    }
    links[nl].first_slot = d.first_s;
    links[nl].count = d.n_send;
    links[nl].remote_first_slot = first;
    links[nl].peer_handles = R;  // the first 128 bytes of the record
    links[nl].remote_first_slot_b = R[17] + (first - R[16]);
    nl++;
  }
  if (ok) {
    p->push_opt = true;
    ok = ad4sm_peer_attach(p, nl, links) == AD4SM_OK;
  }
  // second round: did every rank attach?  otherwise all fall back together
  mine[18] = ok ? 1 : 0;
  rc = gather();
  if (rc) return done(rc);
  everyone = true;
  for (int g = 0; g < nranks; g++) everyone = everyone && all[(size_t)REC * g + 18] != 0;
  if (!everyone) {
    p->push_opt = false;
    ad4sm_peer_attach(p, 0, nullptr);
    p->dist_aware = true;
    p->n_xr = 0;
    p->n_xr_local = 0;
    p->push_mode = false;
  }
  return done(AD4SM_OK);
}

int ad4sm_dist_attach(ad4sm_plan* p, ad4sm_dist* comm, int32_t n_peers, const ad4sm_dist_peer* peers) {
  if (!p || !comm) return fail(AD4SM_EINVAL, "plan or comm is NULL");
  if (n_peers < 0 || (n_peers && !peers)) return fail(AD4SM_EINVAL, "bad peer list");
  if (comm->device != p->device) return fail(AD4SM_EINVAL, "communicator and plan live on different devices");
  if (p->batches.empty() || p->batches[0].halo) return fail(AD4SM_EINVAL, "the first batch of a partitioned plan must be the rank's own elements");
  for (size_t b = 1; b < p->batches.size(); b++)
    if (!p->batches[b].halo) return fail(AD4SM_EINVAL, "ad4sm_dist_attach needs a plan of one own batch followed by HALO batches");
  CK(cudaSetDevice(p->device));
  const long long n_own = p->batches[0].k.ne;
  const int nranks = dist_comm_size(comm->c), me = dist_comm_rank(comm->c);
  // export rows: one per own element some neighbour needs, numbered in (peer, list) order so that a partition whose
  // interface elements each go to one neighbour (slabs) sends every list in place
  std::vector<int> xmap(n_own, -1);
  long long n_export = 0;
  std::vector<std::vector<long long>> rows_x(n_peers), rows_s(n_peers);
  for (int i = 0; i < n_peers; i++) {
    const ad4sm_dist_peer& P = peers[i];
    if (P.rank < 0 || P.rank >= nranks || P.rank == me) return fail(AD4SM_EINVAL, "peer rank out of range");
    if (P.n_send < 0 || P.n_recv < 0 || (P.n_send && !P.send_slots)) return fail(AD4SM_EINVAL, "bad peer send list");
    if (P.n_recv && (P.recv_first_slot < n_own || P.recv_first_slot + P.n_recv > p->ne))
      return fail(AD4SM_EINVAL, "peer receive range outside the plan's HALO slots");
    for (long long k = 0; k < P.n_send; k++) {
      const long long sl = P.send_slots[k];
      if (sl < 0 || sl >= n_own) return fail(AD4SM_EINVAL, "send slot outside the plan's own elements");
      if (xmap[sl] < 0) xmap[sl] = (int)n_export++;
      rows_x[i].push_back(xmap[sl]);
      rows_s[i].push_back(sl);
    }
  }
  p->dpeers.clear();
  p->dcomm = comm;
  p->n_export = n_export;
  const long long kw = (long long)n_pair_blocks(p->N) * p->D * p->D, rw = (long long)p->N * p->D;  // widest field (u)
  int rc = 0;
  if (!p->allphi) rc = dev_alloc(p, &p->allphi, (size_t)nranks);
  if (!rc && !p->phi_local) rc = dev_alloc(p, &p->phi_local, 1);
  if (!rc && p->sorted_built && p->n_own_batches == 1 && n_export > 0) {
    rc = dev_upload(p, &p->xmap, xmap.data(), xmap.size());
    if (!rc) rc = dev_alloc(p, &p->xke, (size_t)(n_export * kw));
    if (!rc) rc = dev_alloc(p, &p->xre, (size_t)(n_export * rw));
  }
  if (rc) return rc;
  auto contiguous = [](const std::vector<long long>& v) -> long long {
    for (size_t k = 1; k < v.size(); k++)
      if (v[k] != v[0] + (long long)k) return -1;
    return v.empty() ? 0 : v[0];
  };
  for (int i = 0; i < n_peers; i++) {
    ad4sm_plan::DistPeer d;
    d.rank = peers[i].rank;
    d.n_send = peers[i].n_send;
    d.recv_first = peers[i].recv_first_slot;
    d.n_recv = peers[i].n_recv;
    d.first_x = contiguous(rows_x[i]);
    d.first_s = contiguous(rows_s[i]);
    if (d.n_send && (d.first_x < 0 || d.first_s < 0)) {
      rc = dev_upload(p, &d.rows_x, rows_x[i].data(), rows_x[i].size());
      if (!rc) rc = dev_upload(p, &d.rows_s, rows_s[i].data(), rows_s[i].size());
      if (!rc) rc = dev_alloc(p, &d.ske, (size_t)(d.n_send * kw));
      if (!rc) rc = dev_alloc(p, &d.sre, (size_t)(d.n_send * rw));
      if (rc) return rc;
    }
    p->dpeers.push_back(d);
  }
  p->dist_aware = true;   // the element stage may use the destination-sorted staging: the exported rows are materialised apart
  p->n_xr = 0;
  p->n_xr_local = 0;
  p->push_mode = false;
  return dist_try_push(p, comm, n_peers, peers);
}

// element stage -> grouped neighbour exchange -> assembly -> ϕ all-gather + rank-ordered sum, all on the plan's stream
static int dist_eval_stages(ad4sm_plan* p, int32_t mode, const double* u_dev, const double* d_dev, double* hist_dev) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  if (!p->dcomm) return fail(AD4SM_ESTATE, "ad4sm_dist_eval before ad4sm_dist_attach");
  // the push transport serves the Kt evaluations on the sorted path (modes 0, 1); ϕ,r-only and d-phase evaluations of the
  // same plan take the send/recv exchange of the element-major rows
  const bool pushable = mode == MODE_ELASTIC || mode == MODE_PF_U;
  const bool push_saved = p->push_mode;
  if (!pushable) p->push_mode = false;
  int rc = eval_elements_impl(p, mode, u_dev, d_dev, hist_dev, false);
  const bool pushed = p->push_mode && p->push_pending;
  p->push_mode = push_saved;
  if (rc) return rc;
  const int DB = mode == MODE_PF_D ? 1 : p->D;
  const long long kw = (long long)n_pair_blocks(p->N) * DB * DB, rw = (long long)p->N * DB;
  const bool want_K = mode != MODE_PHI_R;
  const bool from_x = p->sorted_last && p->xmap;   // K1 wrote the exported rows into xke / xre
  cudaStream_t s = p->stream;
  std::vector<DistXfer> xf;
  for (ad4sm_plan::DistPeer& d : p->dpeers) {
    if (pushed) break;  // the copy engine has pushed the interface rows during K1 (dist_try_push)
    if (d.n_send) {
      const double* kbase = from_x ? p->xke : p->Ke;
      const double* rbase = from_x ? p->xre : p->re;
      const long long first = from_x ? d.first_x : d.first_s;
      double *ksend, *rsend;
      if (first >= 0) {
        ksend = const_cast<double*>(kbase) + first * kw;
        rsend = const_cast<double*>(rbase) + first * rw;
      } else {
        const long long* rows = from_x ? d.rows_x : d.rows_s;
        if (want_K) CK(launch_gather_rows(kbase, rows, d.n_send, (int)kw, d.ske, s));
        CK(launch_gather_rows(rbase, rows, d.n_send, (int)rw, d.sre, s));
        ksend = d.ske;
        rsend = d.sre;
      }
      if (want_K) xf.push_back({1, d.rank, ksend, (size_t)(d.n_send * kw)});
      xf.push_back({1, d.rank, rsend, (size_t)(d.n_send * rw)});
    }
    if (d.n_recv) {
      if (want_K) xf.push_back({0, d.rank, p->Ke + d.recv_first * kw, (size_t)(d.n_recv * kw)});
      xf.push_back({0, d.rank, p->re + d.recv_first * rw, (size_t)(d.n_recv * rw)});
    }
  }
  // The exchange and the ϕ all-gather run on the copy stream while the main stream assembles the interior (the rows and
  // pairs no HALO slot contributes to); only the interface part of the assembly waits for them.
  std::string err;
  const int nranks = dist_comm_size(p->dcomm->c);
  cudaStream_t xs = p->copy_stream;
  if (!p->ev_exports) {
    CK(cudaEventCreateWithFlags(&p->ev_exports, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&p->ev_pushed, cudaEventDisableTiming));
  }
  CK(cudaMemcpyAsync(p->phi_local, p->phi_dev, sizeof(double), cudaMemcpyDeviceToDevice, s));  // this rank's partial ϕ
  CK(cudaEventRecord(p->ev_exports, s));
  CK(cudaStreamWaitEvent(xs, p->ev_exports, 0));
  if (!xf.empty() && dist_exchange(p->dcomm->c, (int)xf.size(), xf.data(), xs, &err)) return fail(AD4SM_ENCCL, err);
  if (dist_allgather(p->dcomm->c, p->phi_local, p->allphi, 1, xs, &err)) return fail(AD4SM_ENCCL, err);
  static const bool dprof = getenv("AD4SM_DIST_PROF") != nullptr;
  cudaEvent_t pe[4] = {nullptr, nullptr, nullptr, nullptr};
  if (dprof && p->dprof.size() < 4 * 4096) {
    for (int k = 0; k < 4; k++) {
      CK(cudaEventCreate(&pe[k]));
      p->dprof.push_back(pe[k]);
    }
    CK(cudaEventRecord(pe[0], s));   // (after the element stage and the enqueue of the exchange)
    CK(cudaEventRecord(pe[1], xs));  // exchange + all-gather complete
  }
  rc = ad4sm_eval_assemble_split_device(p, mode);
  if (rc) return rc;
  if (pe[2]) CK(cudaEventRecord(pe[2], s));  // interior assembled
  // behind the exchange, still on its stream: HALO scatter + the interface rows of r (the main stream is busy with the interior)
  rc = ad4sm_eval_halo_prepare_device(p, mode, xs);
  if (rc) return rc;
  CK(cudaEventRecord(p->ev_pushed, xs));
  CK(cudaStreamWaitEvent(s, p->ev_pushed, 0));
  rc = ad4sm_eval_assemble_device(p, mode);
  if (rc) return rc;
  // ϕ: the partials of all ranks, summed in rank order, into phi_dev (the global ϕ)
  CK(launch_phi_rank_sum(p->allphi, nranks, p->phi_dev, s));
  if (pe[3]) CK(cudaEventRecord(pe[3], s));
  return AD4SM_OK;
}

int ad4sm_dist_eval_device(ad4sm_plan* p, int32_t mode, const double* u_dev, const double* d_dev, double* hist_dev) {
  return dist_eval_stages(p, mode, u_dev, d_dev, hist_dev);
}

int ad4sm_dist_eval(ad4sm_plan* p, int32_t mode, const double* u, const double* d, double* phi_global, double* r_local) {
  int rc = upload_inputs(p, mode, u, d, nullptr);
  if (rc) return rc;
  const bool needs_d = (mode == MODE_PF_U || mode == MODE_PF_D);
  p->h2d_active = p->n_h2d > 0;
  rc = dist_eval_stages(p, mode, p->u_dev, needs_d ? p->d_dev : nullptr, nullptr);
  p->h2d_active = false;
  if (rc) {
    if (p->n_h2d > 0) cudaStreamSynchronize(p->copy_stream);
    return rc;
  }
  const int field = mode == MODE_PF_D ? AD4SM_FIELD_D : AD4SM_FIELD_U;
  rc = download_phi_r(p, field, nullptr, r_local);   // r leaves on the copy stream while Kt is still being assembled
  if (rc) return rc;
  if (phi_global) CK(cudaMemcpyAsync(phi_global, p->phi_dev, sizeof(double), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return AD4SM_OK;
}

// ---- batched D1 post-processing (elements.toolkit.jl:8-132; kernel in post.cu) ----------------------------------------
int ad4sm_post_device(ad4sm_plan* p, const double* u_dev, double* F, double* detJ, double* vol, double* sigma, double* Pk) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  if (!u_dev) return fail(AD4SM_EINVAL, "u is NULL");
  for (const BatchPlan& b : p->batches) {
    if (b.halo) return fail(AD4SM_EINVAL, "post-processing on a plan with HALO batches: evaluate each rank's own elements with its own plan");
    if (b.k.cas) return fail(AD4SM_EINVAL, "post-processing of CAS batches is not defined (the reference has no getF for them)");
    if ((sigma || Pk) && b.k.mat.kind == MAT_PHASEFIELD)
      return fail(AD4SM_EINVAL, "getσ / getP need an elastic material: the reference has no getϕ(F, ::PhaseField) (phasefieldmaterials.jl:20)");
  }
  CK(cudaSetDevice(p->device));
  for (const BatchPlan& b : p->batches) {
    PostArgs a = {};
    a.mat = b.k.mat;
    a.ne = b.k.ne;
    a.slot0 = b.k.slot0;
    a.D = b.k.D;
    a.N = b.k.N;
    a.P = b.k.P;
    a.mf = b.k.mf;
    a.dpn = b.k.dpn;
    a.want_stress = (sigma || Pk) ? 1 : 0;
    a.nodes = b.k.nodes;
    a.gradN = b.k.gradN;
    a.wgt = b.k.wgt;
    a.u = u_dev;
    a.index = p->key_dev ? p->key_dev + b.k.slot0 : nullptr;
    a.F = F;
    a.detJ = detJ;
    a.vol = vol;
    a.sigma = sigma;
    a.Pk = Pk;
    cudaError_t e = launch_post(a, p->stream);
    if (e != cudaSuccess) return fail(AD4SM_ENODEV, std::string("post-processing launch: ") + cudaGetErrorString(e));
  }
  return AD4SM_OK;
}

int ad4sm_post(ad4sm_plan* p, const double* u, double* F, double* detJ, double* vol, double* sigma, double* Pk) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  if (!u) return fail(AD4SM_EINVAL, "u is NULL");
  CK(cudaSetDevice(p->device));
  cudaStream_t s = p->stream;
  const size_t ne = (size_t)p->ne, dd = (size_t)p->D * p->D;
  double* outs[5] = {F, detJ, vol, sigma, Pk};
  const size_t width[5] = {dd, 1, 1, dd, dd};
  double* dev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
  auto release = [&]() {
    for (double* d : dev) cudaFree(d);
  };
  for (int i = 0; i < 5; i++)
    if (outs[i] && cudaMalloc((void**)&dev[i], sizeof(double) * ne * width[i]) != cudaSuccess) {
      release();
      return fail(AD4SM_ENOMEM, "cudaMalloc of the post-processing outputs failed");
    }
  cudaError_t e = cudaMemcpyAsync(p->u_dev, u, sizeof(double) * p->n_nodes * p->dpn, cudaMemcpyHostToDevice, s);
  int rc = e == cudaSuccess ? ad4sm_post_device(p, p->u_dev, dev[0], dev[1], dev[2], dev[3], dev[4]) : fail(AD4SM_ENODEV, "upload of u failed");
  for (int i = 0; i < 5 && !rc; i++)
    if (outs[i] && cudaMemcpyAsync(outs[i], dev[i], sizeof(double) * ne * width[i], cudaMemcpyDeviceToHost, s) != cudaSuccess)
      rc = fail(AD4SM_ENODEV, "download of the post-processing outputs failed");
  if (cudaStreamSynchronize(s) != cudaSuccess && !rc) rc = fail(AD4SM_ENODEV, "post-processing failed");
  release();
  return rc;
}

int ad4sm_profile_enable(ad4sm_plan* p, int32_t on) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  CK(cudaSetDevice(p->device));
  int rc = drain_events(p);
  if (rc) return rc;
  p->prof = on != 0;
  return AD4SM_OK;
}

int ad4sm_profile_read(ad4sm_plan* p, double ms[AD4SM_PROF_NCLASS], int64_t launches[AD4SM_PROF_NCLASS], int32_t reset) {
  if (!p) return fail(AD4SM_EINVAL, "plan is NULL");
  CK(cudaSetDevice(p->device));
  int rc = drain_events(p);
  if (rc) return rc;
  for (int i = 0; i < AD4SM_PROF_NCLASS; i++) {
    if (ms) ms[i] = p->prof_ms[i];
    if (launches) launches[i] = p->prof_launches[i];
    if (reset) {
      p->prof_ms[i] = 0.0;
      p->prof_launches[i] = 0;
    }
  }
  return AD4SM_OK;
}

int ad4sm_measure_fp64_peak(int32_t device, double* tflops, double* sm_mhz_est) {
  if (!tflops) return fail(AD4SM_EINVAL, "tflops is NULL");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return fail(AD4SM_ENODEV, "no CUDA device");
  if (device >= 0) CK(cudaSetDevice(device));
  double tf = 0.0, mhz = 0.0;
  CK(launch_fp64_peak(&tf, &mhz));
  *tflops = tf;
  if (sm_mhz_est) *sm_mhz_est = mhz;
  return AD4SM_OK;
}

}  // extern "C"
