// kernels.cu -- CUDA kernels of the makeϕrKt hot path (sm_100a).
//
//   K1  element energy / gradient / Hessian          replaces the @threads loop of _assemble_ϕrKt
//                                                     (/root/reference/src/elasticelements.jl:216-218)
//   K2  deterministic assembly of Kt (segmented left fold over a precomputed source map, no atomics)
//   K3  residual assembly                             replaces makeϕrKt(Φ, elems, u)
//   ϕ   fixed-tree reduction                          (/root/reference/src/elements.toolkit.jl:169-203)
//   pattern / dropzeros                               sparse(...) structure and dropzeros (toolkit:202)
//
// K1 thread mapping (Hex8: LP = 8 lanes per element, 16 elements per 128-thread CTA):
//   phase 1  lane q evaluates Gauss point q: gathers u, builds F, evaluates the material tangent in
//            registers and leaves (w·A[45], w·P[9], w·ψ) in a shared-memory record;
//   phase 2  lane q owns column node q: streams the 8 Gauss-point records (shared-memory broadcast) and
//            accumulates its 4-5 node blocks K(a,q) = Σ_gp ∇N_aᵀ (A ∇N_q) in registers -- FP64-FMA bound,
//            no cross-lane floating-point reduction.
// Elements with one Gauss point and N <= 4 (Tet04, Tria) run one thread per element, all in registers.
#include <cub/cub.cuh>

#include "element.cuh"
#include "internal.h"

namespace ad4sm {

constexpr int K1_THREADS = 128;

__host__ __device__ inline int k1_rec_stride(int rec_size) { return (rec_size + 1) & ~1; }
// shared-memory stride (doubles) between the elements of a CTA: the 32/LP elements of a warp read the
// same record offset at the same time (broadcast within an element), so successive elements are
// skewed by 16/(32/LP) doubles to land in different banks.
__host__ __device__ inline int k1_elem_stride(int P, int rstride, int LP) {
  const int ew = 32 / LP;
  return ((P * rstride + 15) & ~15) + 16 / ew;
}

template <int D, int N> __device__ __forceinline__ void load_gp_inputs(const K1Args& a, long long e, int gp, double* G,
                                                                       double* ue, const int* nd) {
  const int P = a.P;
  const double* gN = a.gradN + e * (long long)(D * P * N);
#pragma unroll
  for (int k = 0; k < D; k++)
#pragma unroll
    for (int n = 0; n < N; n++) G[k * N + n] = __ldg(gN + (k * P + gp) * N + n);
#pragma unroll
  for (int n = 0; n < N; n++)
#pragma unroll
    for (int j = 0; j < D; j++) ue[n * D + j] = __ldg(a.u + (long long)nd[n] * a.dpn + j);
}

// ---- K1, u-dual, LP lanes per element -----------------------------------------------------------
template <int D, int N, int LP, int MF, bool PF>
__global__ void __launch_bounds__(K1_THREADS, 2) k1_u_kernel(const K1Args a, const int rstride, const int estride) {
  using R = GpRec<D>;
  constexpr int EPC = K1_THREADS / LP;
  constexpr int DD = D * D;
  extern __shared__ double sm[];
  const int el = threadIdx.x / LP, q = threadIdx.x % LP;
  const long long e = (long long)blockIdx.x * EPC + el;
  const int P = a.P;
  double* recs = sm + el * estride;
  const bool valid = e < a.ne;
  if (valid) {
    int nd[N];
#pragma unroll
    for (int n = 0; n < N; n++) nd[n] = __ldg(a.nodes + e * N + n);
    for (int gp = q; gp < P; gp += LP) {
      double G[D * N], ue[N * D], de[N], Nv[N];
      load_gp_inputs<D, N>(a, e, gp, G, ue, nd);
      if (PF) {
#pragma unroll
        for (int n = 0; n < N; n++) {
          de[n] = __ldg(a.dfield + nd[n]);
          Nv[n] = __ldg(a.Nval + (e * P + gp) * N + n);
        }
      }
      gauss_point_u<D, N, MF, PF>(a.mat, G, ue, de, Nv, __ldg(a.wgt + e * P + gp), recs + gp * rstride);
    }
  }
  __syncthreads();
  if (valid && q < N) {
    const long long slot = a.slot0 + e;
    const double* gN = a.gradN + e * (long long)(D * P * N);
    double* Ke = a.want_K ? a.Ke + slot * (long long)(n_pair_blocks(N) * DD) : nullptr;
    column_u<D, N>(q, P, recs, rstride, gN, Ke, a.re + slot * (long long)(N * D));
    if (q == 0) {  // ϕ_e = Σ_gp w ψ, in Gauss-point order (elasticelements.jl:199-206)
      double s = 0.0;
      for (int gp = 0; gp < P; gp++) s += recs[gp * rstride + R::OFF_PSI];
      a.phie[slot] = s;
    }
  }
}

// ---- K1, u-dual, one thread per element (P == 1) ------------------------------------------------
template <int D, int N, int MF, bool PF> __global__ void __launch_bounds__(K1_THREADS) k1_u_kernel_t1(const K1Args a) {
  using R = GpRec<D>;
  constexpr int DD = D * D;
  const long long e = (long long)blockIdx.x * K1_THREADS + threadIdx.x;
  if (e >= a.ne) return;
  int nd[N];
#pragma unroll
  for (int n = 0; n < N; n++) nd[n] = __ldg(a.nodes + e * N + n);
  double G[D * N], ue[N * D], de[N], Nv[N];
  load_gp_inputs<D, N>(a, e, 0, G, ue, nd);
  if (PF) {
#pragma unroll
    for (int n = 0; n < N; n++) {
      de[n] = __ldg(a.dfield + nd[n]);
      Nv[n] = __ldg(a.Nval + e * N + n);
    }
  }
  double rec[R::SIZE];
  gauss_point_u<D, N, MF, PF>(a.mat, G, ue, de, Nv, __ldg(a.wgt + e), rec);
  const long long slot = a.slot0 + e;
  double* Ke = a.want_K ? a.Ke + slot * (long long)(n_pair_blocks(N) * DD) : nullptr;
  double* re = a.re + slot * (long long)(N * D);
#pragma unroll
  for (int q = 0; q < N; q++) column_u<D, N, 1>(q, 1, rec, 0, G, Ke, re);  // G is [D][1][N]
  a.phie[slot] = rec[R::OFF_PSI];
}

// ---- K1, d-dual (makeϕrKt_d) -----------------------------------------------------------------------
template <int D, int N, int LP>
__global__ void __launch_bounds__(K1_THREADS, 2) k1_d_kernel(const K1Args a, const int rstride, const int estride) {
  constexpr int EPC = K1_THREADS / LP;
  extern __shared__ double sm[];
  const int el = threadIdx.x / LP, q = threadIdx.x % LP;
  const long long e = (long long)blockIdx.x * EPC + el;
  const int P = a.P;
  double* recs = sm + el * estride;
  const bool valid = e < a.ne;
  if (valid) {
    int nd[N];
#pragma unroll
    for (int n = 0; n < N; n++) nd[n] = __ldg(a.nodes + e * N + n);
    for (int gp = q; gp < P; gp += LP) {
      double G[D * N], ue[N * D], de[N], Nv[N];
      load_gp_inputs<D, N>(a, e, gp, G, ue, nd);
#pragma unroll
      for (int n = 0; n < N; n++) {
        de[n] = __ldg(a.dfield + nd[n]);
        Nv[n] = __ldg(a.Nval + (e * P + gp) * N + n);
      }
      double* h = a.hist ? a.hist + (e * P + gp) * 2 : nullptr;
      gauss_point_d<D, N>(a.mat, G, Nv, ue, de, __ldg(a.wgt + e * P + gp), h, recs + gp * rstride);
    }
  }
  __syncthreads();
  if (valid && q < N) {
    const long long slot = a.slot0 + e;
    const double* gN = a.gradN + e * (long long)(D * P * N);
    const double* Nv = a.Nval + e * (long long)(P * N);
    double* Ke = a.want_K ? a.Ke + slot * (long long)n_pair_blocks(N) : nullptr;
    column_d<D, N>(q, P, recs, rstride, gN, Nv, Ke, a.re + slot * (long long)N);
    if (q == 0) {
      double s = 0.0;
      for (int gp = 0; gp < P; gp++) s += recs[gp * rstride + 2];
      a.phie[slot] = s;
    }
  }
}

template <int D, int N> __global__ void __launch_bounds__(K1_THREADS) k1_d_kernel_t1(const K1Args a) {
  const long long e = (long long)blockIdx.x * K1_THREADS + threadIdx.x;
  if (e >= a.ne) return;
  int nd[N];
#pragma unroll
  for (int n = 0; n < N; n++) nd[n] = __ldg(a.nodes + e * N + n);
  double G[D * N], ue[N * D], de[N], Nv[N];
  load_gp_inputs<D, N>(a, e, 0, G, ue, nd);
#pragma unroll
  for (int n = 0; n < N; n++) {
    de[n] = __ldg(a.dfield + nd[n]);
    Nv[n] = __ldg(a.Nval + e * N + n);
  }
  double rec[GpRecD<D>::SIZE];
  gauss_point_d<D, N>(a.mat, G, Nv, ue, de, __ldg(a.wgt + e), a.hist ? a.hist + e * 2 : nullptr, rec);
  const long long slot = a.slot0 + e;
  double* Ke = a.want_K ? a.Ke + slot * (long long)n_pair_blocks(N) : nullptr;
  double* re = a.re + slot * (long long)N;
#pragma unroll
  for (int q = 0; q < N; q++) column_d<D, N, 1>(q, 1, rec, 0, G, Nv, Ke, re);
  a.phie[slot] = rec[2];
}

// ---- K1 dispatch ---------------------------------------------------------------------------------
static bool lanes_for(int D, int N, int P, int& LP) {
  if ((D == 2 && N == 3) || (D == 3 && N == 4)) {
    LP = (P == 1) ? 1 : 4;
    return true;
  }
  if (D == 2 && N == 4) {
    LP = 4;
    return true;
  }
  if (D == 3 && (N == 6 || N == 8)) {
    LP = 8;
    return true;
  }
  return false;
}

const char* k1_unsupported_reason(int D, int N, int P, int mf, int mode) {
  int LP;
  if (!lanes_for(D, N, P, LP)) return "unsupported element shape (D, N): supported are (2,3), (2,4), (3,4), (3,6), (3,8)";
  if (P < 1 || P > 64) return "number of Gauss points must be in 1..64";
  if (mf == MF_OGDEN) return "Ogden: the reference defines only the scalar energy (no derivative); not implemented";
  if (mf < 0 || mf > MF_OGDEN) return "unknown material family";
  (void)mode;
  return nullptr;
}

template <int D, int N, int LP, int MF, bool PF> static cudaError_t launch_u_t(const K1Args& a, cudaStream_t s) {
  if constexpr (LP == 1) {
    const unsigned grid = (unsigned)((a.ne + K1_THREADS - 1) / K1_THREADS);
    k1_u_kernel_t1<D, N, MF, PF><<<grid, K1_THREADS, 0, s>>>(a);
    return cudaGetLastError();
  } else {
    constexpr int EPC = K1_THREADS / LP;
    const int rstride = k1_rec_stride(GpRec<D>::SIZE);
    const int estride = k1_elem_stride(a.P, rstride, LP);
    const size_t smem = (size_t)EPC * estride * sizeof(double);
    auto kern = k1_u_kernel<D, N, LP, MF, PF>;
    if (smem > 48 * 1024) {
      cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err != cudaSuccess) return err;
    }
    const unsigned grid = (unsigned)((a.ne + EPC - 1) / EPC);
    kern<<<grid, K1_THREADS, smem, s>>>(a, rstride, estride);
    return cudaGetLastError();
  }
}

template <int D, int N, int LP> static cudaError_t launch_u_mf(const K1Args& a, bool pf, cudaStream_t s) {
  switch (a.mf) {
    case MF_HOOKE: return pf ? launch_u_t<D, N, LP, MF_HOOKE, true>(a, s) : launch_u_t<D, N, LP, MF_HOOKE, false>(a, s);
    case MF_NEO: return launch_u_t<D, N, LP, MF_NEO, false>(a, s);
    case MF_MOONEY: return launch_u_t<D, N, LP, MF_MOONEY, false>(a, s);
    case MF_PFNEO: return launch_u_t<D, N, LP, MF_PFNEO, true>(a, s);
    default: return cudaErrorInvalidValue;
  }
}

template <int D, int N, int LP> static cudaError_t launch_d_t(const K1Args& a, cudaStream_t s) {
  if constexpr (LP == 1) {
    const unsigned grid = (unsigned)((a.ne + K1_THREADS - 1) / K1_THREADS);
    k1_d_kernel_t1<D, N><<<grid, K1_THREADS, 0, s>>>(a);
    return cudaGetLastError();
  } else {
    constexpr int EPC = K1_THREADS / LP;
    const int rstride = k1_rec_stride(GpRecD<D>::SIZE);
    const int estride = k1_elem_stride(a.P, rstride, LP);
    const size_t smem = (size_t)EPC * estride * sizeof(double);
    auto kern = k1_d_kernel<D, N, LP>;
    if (smem > 48 * 1024) {
      cudaError_t err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (err != cudaSuccess) return err;
    }
    const unsigned grid = (unsigned)((a.ne + EPC - 1) / EPC);
    kern<<<grid, K1_THREADS, smem, s>>>(a, rstride, estride);
    return cudaGetLastError();
  }
}

// mode: MODE_ELASTIC / MODE_PF_U (u-dual) or MODE_PF_D (d-dual); a.want_K selects ϕ,r-only evaluation.
cudaError_t launch_k1(const K1Args& a, int mode, cudaStream_t s) {
  if (a.ne == 0) return cudaSuccess;
  int LP;
  if (!lanes_for(a.D, a.N, a.P, LP)) return cudaErrorInvalidValue;
  const bool is_pf = a.mat.kind == MAT_PHASEFIELD;
  if ((mode == MODE_ELASTIC) == is_pf) return cudaErrorInvalidValue;  // PhaseField needs d; elastic has none
  const bool pf = (mode == MODE_PF_U);
#define SHAPE(DD, NN, LL)                 \
  if (a.D == DD && a.N == NN && LP == LL) \
    return mode == MODE_PF_D ? launch_d_t<DD, NN, LL>(a, s) : launch_u_mf<DD, NN, LL>(a, pf, s);
  SHAPE(2, 3, 1) SHAPE(2, 3, 4) SHAPE(2, 4, 4) SHAPE(3, 4, 1) SHAPE(3, 4, 4) SHAPE(3, 6, 8) SHAPE(3, 8, 8)
#undef SHAPE
  return cudaErrorInvalidValue;
}

// ---- K2: Kt values.  One thread per (row node i, row component j, neighbour p): sums the DB x DB source
// blocks of node pair (i, nbr[p]) in `elems` order (left fold, like Julia's sparse(), toolkit:202) and writes
// the dpn consecutive values of CSC column / CSR row (i, j).  Consecutive threads write consecutive memory.
template <int DB> __global__ void __launch_bounds__(256) k2_kernel(const AsmArgs a) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int dpn = a.dpn;
  if (g >= a.n_blocks * dpn) return;
  const unsigned i = __ldg(a.blk_row + g / dpn);
  const long long nb0 = __ldg(a.nbr_ptr + i);
  const int nn = (int)(__ldg(a.nbr_ptr + i + 1) - nb0);
  const int gl = (int)(g - nb0 * dpn);
  const int j = gl / nn, p = gl - j * nn;
  const long long blk = nb0 + p;
  double acc[DB];
#pragma unroll
  for (int l = 0; l < DB; l++) acc[l] = 0.0;
  if (j < DB) {
    const unsigned s0 = __ldg(a.blk_ptr + blk), s1 = __ldg(a.blk_ptr + blk + 1);
    for (unsigned s = s0; s < s1; s++) {
      const unsigned code = __ldg(a.src + s);
      const double* B = a.Ke + (long long)(code >> 1) * (DB * DB);
      if (code & 1u) {
#pragma unroll
        for (int l = 0; l < DB; l++) acc[l] += B[l * DB + j];
      } else {
#pragma unroll
        for (int l = 0; l < DB; l++) acc[l] += B[j * DB + l];
      }
    }
  }
  double* out = a.nzval + nb0 * (dpn * dpn) + (long long)j * dpn * nn + (long long)p * dpn;
  int nz = 0;
#pragma unroll
  for (int l = 0; l < DB; l++) {
    out[l] = acc[l];
    nz += (acc[l] == 0.0);
  }
  for (int l = DB; l < dpn; l++) {
    out[l] = 0.0;
    nz++;
  }
  if (nz) atomicAdd(a.n_zero, (unsigned long long)nz);  // integer atomic: order-independent
}

// ---- K3: residual.  One thread per (node, component): left fold over the node's elements in `elems` order
// (r[idxii] += grad, toolkit:190).
__global__ void __launch_bounds__(256) k3_kernel(const AsmArgs a) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int dpn = a.dpn;
  if (g >= a.n_nodes * dpn) return;
  const long long i = g / dpn;
  const int j = (int)(g - i * dpn);
  double acc = 0.0;
  if (j < a.DB) {
    const unsigned s0 = __ldg(a.adj_ptr + i), s1 = __ldg(a.adj_ptr + i + 1);
    for (unsigned s = s0; s < s1; s++) acc += a.re[(long long)__ldg(a.adj + s) * a.DB + j];
  }
  a.r[g] = acc;
}

cudaError_t launch_k2(const AsmArgs& a, bool want_K, cudaStream_t s) {
  if (want_K) {
    cudaError_t err = cudaMemsetAsync(a.n_zero, 0, sizeof(unsigned long long), s);
    if (err != cudaSuccess) return err;
    const long long nt = a.n_blocks * a.dpn;
    if (nt > 0) {
      const unsigned grid = (unsigned)((nt + 255) / 256);
      switch (a.DB) {
        case 1: k2_kernel<1><<<grid, 256, 0, s>>>(a); break;
        case 2: k2_kernel<2><<<grid, 256, 0, s>>>(a); break;
        case 3: k2_kernel<3><<<grid, 256, 0, s>>>(a); break;
        default: return cudaErrorInvalidValue;
      }
      err = cudaGetLastError();
      if (err != cudaSuccess) return err;
    }
  }
  const long long nr = a.n_nodes * a.dpn;
  if (nr > 0) k3_kernel<<<(unsigned)((nr + 255) / 256), 256, 0, s>>>(a);
  return cudaGetLastError();
}

// ---- ϕ = Σ_e ϕ_e: fixed two-level tree (chunk per block, then one block), run-to-run reproducible --------
constexpr int PHI_PARTIALS = 1024;
int phi_reduce_partials() { return PHI_PARTIALS; }

__device__ __forceinline__ double block_sum_256(double v, double* sh) {
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  return sh[0];
}
__global__ void __launch_bounds__(256) phi_partial_kernel(const double* __restrict__ x, long long n, double* partials) {
  __shared__ double sh[256];
  const long long chunk = (n + PHI_PARTIALS - 1) / PHI_PARTIALS;
  const long long b0 = blockIdx.x * chunk;
  long long b1 = b0 + chunk;
  if (b1 > n) b1 = n;
  double v = 0.0;
  for (long long i = b0 + threadIdx.x; i < b1; i += 256) v += x[i];
  const double t = block_sum_256(v, sh);
  if (threadIdx.x == 0) partials[blockIdx.x] = t;
}
__global__ void __launch_bounds__(256) phi_final_kernel(const double* __restrict__ partials, double* phi) {
  __shared__ double sh[256];
  double v = 0.0;
  for (int i = threadIdx.x; i < PHI_PARTIALS; i += 256) v += partials[i];
  const double t = block_sum_256(v, sh);
  if (threadIdx.x == 0) *phi = t;
}
cudaError_t launch_phi_reduce(const double* phie, long long n, double* partials, double* phi, cudaStream_t s) {
  phi_partial_kernel<<<PHI_PARTIALS, 256, 0, s>>>(phie, n, partials);
  phi_final_kernel<<<1, 256, 0, s>>>(partials, phi);
  return cudaGetLastError();
}

// ---- structural CSC pattern (Int64, 1-based) of the assembled field ----------------------------------------
__global__ void __launch_bounds__(256) pattern_kernel(const AsmArgs a, long long* colptr, long long* rowval) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int dpn = a.dpn;
  if (g > a.n_nodes * dpn) return;
  if (g == a.n_nodes * dpn) {
    colptr[g] = a.n_blocks * dpn * dpn + 1;
    return;
  }
  const long long i = g / dpn;
  const int j = (int)(g - i * dpn);
  const long long nb0 = a.nbr_ptr[i];
  const int nn = (int)(a.nbr_ptr[i + 1] - nb0);
  const long long base = nb0 * (dpn * dpn) + (long long)j * dpn * nn;
  colptr[g] = base + 1;
  for (int p = 0; p < nn; p++) {
    const long long c = (long long)a.nbr[nb0 + p] * dpn;
    for (int l = 0; l < dpn; l++) rowval[base + p * dpn + l] = c + l + 1;
  }
}
cudaError_t launch_pattern(const AsmArgs& a, long long* colptr, long long* rowval, cudaStream_t s) {
  const long long n = a.n_nodes * a.dpn + 1;
  pattern_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(a, colptr, rowval);
  return cudaGetLastError();
}

// ---- dropzeros (toolkit:202): per-column non-zero counts -> exclusive scan -> compaction -----------------
__global__ void __launch_bounds__(256) dz_count_kernel(long long n_dof, const long long* colptr, const double* nzval,
                                                       long long* cnt) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_dof) return;
  long long k = 0;
  if (c < n_dof)
    for (long long s = colptr[c] - 1; s < colptr[c + 1] - 1; s++) k += (nzval[s] != 0.0);
  cnt[c] = k;
}
__global__ void __launch_bounds__(256) dz_fill_kernel(long long n_dof, const long long* colptr, const long long* rowval,
                                                      const double* nzval, long long* colptr_out, long long* rowval_out,
                                                      double* nzval_out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_dof) return;
  const long long o0 = colptr_out[c];  // exclusive scan of the counts (0-based)
  if (c < n_dof) {
    long long o = o0;
    for (long long s = colptr[c] - 1; s < colptr[c + 1] - 1; s++) {
      const double v = nzval[s];
      if (v != 0.0) {
        rowval_out[o] = rowval[s];
        nzval_out[o] = v;
        o++;
      }
    }
  }
}
__global__ void __launch_bounds__(256) dz_onebase_kernel(long long n, long long* colptr_out) {
  const long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < n) colptr_out[c] += 1;
}
size_t dropzeros_tmp_bytes(long long n_dof) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (long long*)nullptr, (long long*)nullptr, n_dof + 1);
  return bytes;
}
cudaError_t launch_dropzeros(long long n_dof, const long long* colptr, const long long* rowval, const double* nzval,
                             long long* colptr_out, long long* rowval_out, double* nzval_out, void* cub_tmp,
                             size_t cub_tmp_bytes, cudaStream_t s) {
  const unsigned grid = (unsigned)((n_dof + 1 + 255) / 256);
  // counts are staged in colptr_out's tail-compatible scratch: use rowval_out? no -- scan in place on colptr_out
  dz_count_kernel<<<grid, 256, 0, s>>>(n_dof, colptr, nzval, colptr_out);
  cudaError_t err = cub::DeviceScan::ExclusiveSum(cub_tmp, cub_tmp_bytes, colptr_out, colptr_out, n_dof + 1, s);
  if (err != cudaSuccess) return err;
  dz_fill_kernel<<<grid, 256, 0, s>>>(n_dof, colptr, rowval, nzval, colptr_out, rowval_out, nzval_out);
  dz_onebase_kernel<<<grid, 256, 0, s>>>(n_dof + 1, colptr_out);
  return cudaGetLastError();
}

// ---- FP64 FMA peak (roofline denominator for K1; MEASURED_PEAKS.json has no FP64 entry) --------------------
__global__ void __launch_bounds__(256) dfma_peak_kernel(double* out, int iters, double a, double b) {
  double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
  for (int i = 0; i < iters; i++) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      x0 = fma(x0, a, b);
      x1 = fma(x1, a, b);
      x2 = fma(x2, a, b);
      x3 = fma(x3, a, b);
      x4 = fma(x4, a, b);
      x5 = fma(x5, a, b);
      x6 = fma(x6, a, b);
      x7 = fma(x7, a, b);
    }
  }
  out[(long long)blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
cudaError_t launch_fp64_peak(double* tflops, double* mhz) {
  int dev = 0, sms = 0, khz = 0;
  cudaError_t err = cudaGetDevice(&dev);
  if (err != cudaSuccess) return err;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
  const int blocks = sms * 8, threads = 256, iters = 4096;
  double* out = nullptr;
  err = cudaMalloc(&out, sizeof(double) * blocks * threads);
  if (err != cudaSuccess) return err;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  double best = 0.0;
  for (int rep = 0; rep < 6; rep++) {  // first reps warm up clocks
    cudaEventRecord(e0);
    dfma_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
    cudaEventRecord(e1);
    err = cudaEventSynchronize(e1);
    if (err != cudaSuccess) break;
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    const double fl = 2.0 * 64.0 * iters * (double)blocks * threads;
    const double tf = fl / (ms * 1e-3) / 1e12;
    if (rep >= 2 && tf > best) best = tf;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  if (err != cudaSuccess) return err;
  *tflops = best;
  *mhz = khz / 1000.0;
  return cudaSuccess;
}

}  // namespace ad4sm
