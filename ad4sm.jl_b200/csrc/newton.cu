// newton.cu -- device side of the Newton-step glue around makeϕrKt.
//
// Replaces, for the nEqs == 0 branch of Solvers.solvestep!, what the reference does on the host after every
// makeϕrKt (/root/reference/src/solvers.jl):
//   ifreeu / icnstu = findall(bfreeu[:]) / findall(.!bfreeu[:])                          :183-184
//   predictor  res = fi[ifreeu] - fe[ifreeu];  res .-= Kt[ifreeu,icnstu]*Δucnst          :220-221
//   corrector  res = fe[ifreeu] - fi[ifreeu];  norm(fi[icnstu]), norm(res)               :262-265
//   the matrix handed to lu(): Kt[ifreeu,ifreeu]                                          :223, 285
// so that only res, two norms and the VALUES of Kt[ifree,ifree] (pattern built once) have to leave the device.
// Kt is symmetric and stored with every CSC column contiguous (= CSR row), node-block structured: column (i, jc) holds,
// for each neighbour node p of i in ascending order, dpn consecutive rows p*dpn + l.
#include <cub/cub.cuh>

#include "internal.h"

namespace ad4sm {

__device__ __forceinline__ int warp_excl_scan(int v, int lane, int& total) {
  int x = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int y = __shfl_up_sync(0xffffffffu, x, o);
    if (lane >= o) x += y;
  }
  total = __shfl_sync(0xffffffffu, x, 31);
  return x - v;
}

// One warp per free column f (dof c = free_list[f]); lane <-> neighbour block p of the column's node.
//   cnt != nullptr        : cnt[f] = number of free rows in the column                     (pattern, pass 1)
//   rowval_ff != nullptr  : 1-based free-numbered row indices at colptr_ff[f]-1 ...        (pattern, pass 2)
//   nzval_ff != nullptr   : the values, same positions                                      (every Newton iteration)
//   du != nullptr         : fc[f] = Σ_{j constrained} Kt[c, j] * du[rank(j)]  (= (Kt[ifree,icnst]*Δucnst)[f], Kt symmetric);
//                           products and sums separately rounded like SparseArrays' C[i] += nzv*x (no FMA contraction),
//                           partial sums of the (<= 32 x dpn) entries of a chunk combined by a fixed butterfly
__global__ void __launch_bounds__(256) newton_column_kernel(const NewtonArgs a) {
  const long long f = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (f >= a.nfree) return;
  const int dpn = a.dpn;
  const long long c = a.free_list[f];
  const long long i = c / dpn;
  const int jc = (int)(c - i * dpn);
  const long long nb0 = a.nbr_ptr[i];
  const int nn = (int)(a.nbr_ptr[i + 1] - nb0);
  const long long base = nb0 * (dpn * dpn) + (long long)jc * dpn * nn;
  const long long out0 = a.colptr_ff ? a.colptr_ff[f] - 1 : 0;
  int off = 0;
  double acc = 0.0;
  for (int p0 = 0; p0 < nn; p0 += 32) {
    const int p = p0 + lane;
    int ids[8], nfree = 0;
    if (p < nn) {
      const long long pn = a.nbr[nb0 + p];
      for (int l = 0; l < dpn; l++) {
        ids[l] = a.fidx[pn * dpn + l];
        nfree += ids[l] >= 0;
      }
    }
    int tot;
    int pos = warp_excl_scan(nfree, lane, tot);
    if (p < nn && (a.rowval_ff || a.nzval_ff || a.du)) {
      for (int l = 0; l < dpn; l++) {
        if (ids[l] >= 0) {
          const long long o = out0 + off + pos;
          if (a.rowval_ff) a.rowval_ff[o] = (long long)ids[l] + 1;
          if (a.nzval_ff) a.nzval_ff[o] = a.nzval[base + (long long)p * dpn + l];
          pos++;
        } else if (a.du) {
          acc = __dadd_rn(acc, __dmul_rn(a.nzval[base + (long long)p * dpn + l], a.du[-1 - ids[l]]));
        }
      }
    }
    off += tot;
  }
  if (a.cnt && lane == 0) a.cnt[f] = off;
  if (a.du) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc = __dadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, o));
    if (lane == 0) a.fc[f] = acc;
  }
}

__global__ void __launch_bounds__(256) newton_onebase_kernel(long long n, long long* x) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) x[k] += 1;
}

// predictor (kind 0): res = (fi - fe) - fc        solvers.jl:220-221
// corrector (kind 1): res = fe - fi                solvers.jl:262
__global__ void __launch_bounds__(256) newton_res_kernel(long long nfree, const int* free_list, const double* r, const double* fe,
                                                         const double* fc, int kind, double* res) {
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= nfree) return;
  const long long c = free_list[f];
  const double fi = r[c], fev = fe ? fe[c] : 0.0;
  res[f] = kind == 0 ? ((fi - fev) - (fc ? fc[f] : 0.0)) : (fev - fi);
}

// Σ x[idx[k]]² (idx == nullptr: x[k]²): fixed two-level tree, run-to-run reproducible
constexpr int NW_PARTIALS = 512;
__global__ void __launch_bounds__(256) newton_sumsq_partial_kernel(const double* x, const int* idx, long long n, double* partials) {
  __shared__ double sh[256];
  const long long chunk = (n + NW_PARTIALS - 1) / NW_PARTIALS;
  const long long b0 = blockIdx.x * chunk;
  long long b1 = b0 + chunk;
  if (b1 > n) b1 = n;
  double v = 0.0;
  for (long long k = b0 + threadIdx.x; k < b1; k += 256) {
    const double t = idx ? x[idx[k]] : x[k];
    v += t * t;
  }
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) partials[blockIdx.x] = sh[0];
}
__global__ void __launch_bounds__(256) newton_norm_final_kernel(const double* partials, double* out) {
  __shared__ double sh[256];
  double v = 0.0;
  for (int k = threadIdx.x; k < NW_PARTIALS; k += 256) v += partials[k];
  sh[threadIdx.x] = v;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if ((int)threadIdx.x < s) sh[threadIdx.x] += sh[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = sqrt(sh[0]);
}

int newton_partials() { return NW_PARTIALS; }

cudaError_t launch_newton_columns(const NewtonArgs& a, cudaStream_t s) {
  if (a.nfree > 0) {
    const long long threads = a.nfree * 32;
    newton_column_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, s>>>(a);
  }
  return cudaGetLastError();
}

size_t newton_scan_tmp_bytes(long long n) {
  size_t bytes = 0;
  cub::DeviceScan::ExclusiveSum(nullptr, bytes, (long long*)nullptr, (long long*)nullptr, n);
  return bytes;
}
// cnt[0..nfree) column counts (cnt[nfree] = 0) -> 1-based colptr in place
cudaError_t launch_newton_colptr(long long* cnt, long long nfree, void* tmp, size_t tmp_bytes, cudaStream_t s) {
  cudaError_t e = cub::DeviceScan::ExclusiveSum(tmp, tmp_bytes, cnt, cnt, nfree + 1, s);
  if (e != cudaSuccess) return e;
  newton_onebase_kernel<<<(unsigned)((nfree + 1 + 255) / 256), 256, 0, s>>>(nfree + 1, cnt);
  return cudaGetLastError();
}

cudaError_t launch_newton_res(long long nfree, const int* free_list, const double* r, const double* fe, const double* fc, int kind,
                              double* res, long long ncnst, const int* cnst_list, double* partials, double* norms, cudaStream_t s) {
  if (nfree > 0) newton_res_kernel<<<(unsigned)((nfree + 255) / 256), 256, 0, s>>>(nfree, free_list, r, fe, fc, kind, res);
  newton_sumsq_partial_kernel<<<NW_PARTIALS, 256, 0, s>>>(res, nullptr, nfree, partials);
  newton_norm_final_kernel<<<1, 256, 0, s>>>(partials, norms);                       // norms[0] = norm(res)
  newton_sumsq_partial_kernel<<<NW_PARTIALS, 256, 0, s>>>(r, cnst_list, ncnst, partials + NW_PARTIALS);
  newton_norm_final_kernel<<<1, 256, 0, s>>>(partials + NW_PARTIALS, norms + 1);     // norms[1] = norm(fi[icnst])
  return cudaGetLastError();
}

}  // namespace ad4sm
