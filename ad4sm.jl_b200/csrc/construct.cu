// construct.cu -- batched element constructors on the device: coordinates -> (∇N, wgt, N, V) for all elements of a
// homogeneous group at once.  Replaces, per element, the reference constructors
//   Tria / TriaP      /root/reference/src/elasticelements.jl:35-46,    phasefieldelements.jl:9-24
//   Quad / QuadP      /root/reference/src/elasticelements.jl:50-79,    phasefieldelements.jl:25-54   (QuadR/QuadPR: 1-point rule)
//   Tet04 / Tet04P    /root/reference/src/elasticelements.jl:81-96,    phasefieldelements.jl:129-145
//   Hex08 / Hex08P    /root/reference/src/elasticelements.jl:129-162,  phasefieldelements.jl:56-94   (Hex08R/Hex08PR: 1-point rule)
//   Wdg06 / Wdg06P    /root/reference/src/elasticelements.jl:163-197,  phasefieldelements.jl:95-128
// The reference differentiates the shape functions with D1 duals at each Gauss point, forms J[j][i] = ∂x_i/∂ξ_j and
// solves J \ ∂N/∂ξ; here the same quantities are written out (the dual products only ever add exact zeros) and the
// D x D system is solved by Gaussian elimination with partial pivoting, the algorithm behind `\`.
// Output layouts are the plan's: ∇N [ne][D][P][N], wgt [ne][P], N [ne][P][N]; Gauss points in the reference's storage
// order (ξ index fastest: `tuple(Nx...)` flattens Nx[ii,jj,kk] column-major).
#include "internal.h"

namespace ad4sm {

namespace {

template <int SHAPE> struct ShapeTraits;
template <> struct ShapeTraits<CS_QUAD4> { static constexpr int D = 2, N = 4; };
template <> struct ShapeTraits<CS_HEX8> { static constexpr int D = 3, N = 8; };
template <> struct ShapeTraits<CS_WDG6> { static constexpr int D = 3, N = 6; };

// shape-function values and parametric derivatives dN[j][a] = ∂N_a/∂ξ_j at the point x
template <int SHAPE> __device__ __forceinline__ void shape_eval(const double* x, double* Nv, double (*dN)[8]) {
  if constexpr (SHAPE == CS_QUAD4) {  // elasticelements.jl:57
    const double xm = 1.0 - x[0], xp = 1.0 + x[0], em = 1.0 - x[1], ep = 1.0 + x[1];
    Nv[0] = xm * em / 4, Nv[1] = xp * em / 4, Nv[2] = xp * ep / 4, Nv[3] = xm * ep / 4;
    dN[0][0] = -em / 4, dN[0][1] = em / 4, dN[0][2] = ep / 4, dN[0][3] = -ep / 4;
    dN[1][0] = -xm / 4, dN[1][1] = -xp / 4, dN[1][2] = xp / 4, dN[1][3] = xm / 4;
  } else if constexpr (SHAPE == CS_HEX8) {  // elasticelements.jl:134-137
#pragma unroll
    for (int a = 0; a < 8; a++) {
      const double sx = ((a & 1) ^ ((a >> 1) & 1)) ? 1.0 : -1.0;  // -,+,+,-,-,+,+,-
      const double se = (a & 2) ? 1.0 : -1.0;                      // -,-,+,+,-,-,+,+
      const double sz = (a & 4) ? 1.0 : -1.0;
      const double fx = 1.0 + sx * x[0], fe = 1.0 + se * x[1], fz = 1.0 + sz * x[2];
      Nv[a] = fx * fe * fz / 8;
      dN[0][a] = sx * fe * fz / 8;
      dN[1][a] = fx * se * fz / 8;
      dN[2][a] = fx * fe * sz / 8;
    }
  } else {  // wedge, elasticelements.jl:166-167
    const double t = 1.0 - x[0] - x[1], zm = 1.0 - x[2], zp = 1.0 + x[2];
    Nv[0] = zm * t / 2, Nv[1] = zm * x[0] / 2, Nv[2] = zm * x[1] / 2;
    Nv[3] = zp * t / 2, Nv[4] = zp * x[0] / 2, Nv[5] = zp * x[1] / 2;
    dN[0][0] = -zm / 2, dN[0][1] = zm / 2, dN[0][2] = 0.0, dN[0][3] = -zp / 2, dN[0][4] = zp / 2, dN[0][5] = 0.0;
    dN[1][0] = -zm / 2, dN[1][1] = 0.0, dN[1][2] = zm / 2, dN[1][3] = -zp / 2, dN[1][4] = 0.0, dN[1][5] = zp / 2;
    dN[2][0] = -t / 2, dN[2][1] = -x[0] / 2, dN[2][2] = -x[1] / 2, dN[2][3] = t / 2, dN[2][4] = x[0] / 2, dN[2][5] = x[1] / 2;
  }
}

// detJ as written in the reference (elements.toolkit.jl:33-36)
template <int D> __device__ __forceinline__ double det_small(const double (*J)[3]) {
  if constexpr (D == 2) return J[0][0] * J[1][1] - J[0][1] * J[1][0];
  return J[0][0] * (J[1][1] * J[2][2] - J[1][2] * J[2][1]) - J[0][1] * (J[1][0] * J[2][2] - J[1][2] * J[2][0]) +
         J[0][2] * (J[1][0] * J[2][1] - J[1][1] * J[2][0]);
}

// G = J \ B (B: D x N), Gaussian elimination with partial pivoting on a copy of J; everything stays in registers
template <int D, int N> __device__ __forceinline__ void solve_small(const double (*J)[3], double (*B)[8]) {
  double A[D][D];
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) A[i][j] = J[i][j];
#pragma unroll
  for (int c = 0; c < D - 1; c++) {
    // pivot: largest |A[r][c]|, r >= c (first one on ties, like idamax); rows are exchanged by conditional swaps so
    // that no array is indexed with a run-time value
#pragma unroll
    for (int r = c + 1; r < D; r++) {
      if (fabs(A[r][c]) > fabs(A[c][c])) {
#pragma unroll
        for (int j = 0; j < D; j++) {
          const double t = A[c][j];
          A[c][j] = A[r][j];
          A[r][j] = t;
        }
#pragma unroll
        for (int a = 0; a < N; a++) {
          const double t = B[c][a];
          B[c][a] = B[r][a];
          B[r][a] = t;
        }
      }
    }
#pragma unroll
    for (int r = c + 1; r < D; r++) {
      const double l = A[r][c] / A[c][c];
#pragma unroll
      for (int j = c + 1; j < D; j++) A[r][j] -= l * A[c][j];
#pragma unroll
      for (int a = 0; a < N; a++) B[r][a] -= l * B[c][a];
    }
  }
#pragma unroll
  for (int r = D - 1; r >= 0; r--) {
#pragma unroll
    for (int a = 0; a < N; a++) {
      double s = B[r][a];
#pragma unroll
      for (int j = r + 1; j < D; j++) s -= A[r][j] * B[j][a];
      B[r][a] = s / A[r][r];
    }
  }
}

// isoparametric shapes: one thread per (element, Gauss point)
template <int SHAPE> __global__ void __launch_bounds__(128) construct_iso_kernel(const ConstructArgs a) {
  constexpr int D = ShapeTraits<SHAPE>::D, N = ShapeTraits<SHAPE>::N;
  const int P = a.P;
  const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= a.ne * P) return;
  const long long e = t / P;
  const int gp = (int)(t - e * P);
  double x[3] = {a.pt[gp][0], a.pt[gp][1], a.pt[gp][2]};
  double Nv[8], dN[3][8];
  shape_eval<SHAPE>(x, Nv, dN);
  double J[3][3];
#pragma unroll
  for (int j = 0; j < D; j++)
#pragma unroll
    for (int i = 0; i < D; i++) J[j][i] = 0.0;
#pragma unroll
  for (int n = 0; n < N; n++) {  // p = Σ_a N_a p0_a, left to right (elasticelements.jl:68, 149)
    const long long nd = a.nodes[e * N + n] - 1;
    double X[D];
#pragma unroll
    for (int i = 0; i < D; i++) X[i] = __ldg(a.coords + nd * a.ldc + i);
#pragma unroll
    for (int j = 0; j < D; j++)
#pragma unroll
      for (int i = 0; i < D; i++) J[j][i] += dN[j][n] * X[i];
  }
  double w = det_small<D>(J);
#pragma unroll
  for (int k = 0; k < 3; k++)
    if (k < a.n_wfac) w *= a.wfac[gp][k];  // detJ(F)*wξ*wη*wζ, left to right (elasticelements.jl:74, 157)
  solve_small<D, N>(J, dN);
  a.wgt[e * P + gp] = w;
#pragma unroll
  for (int k = 0; k < D; k++)
#pragma unroll
    for (int n = 0; n < N; n++) a.gradN[((e * D + k) * P + gp) * N + n] = dN[k][n];
  if (a.Nval) {
#pragma unroll
    for (int n = 0; n < N; n++) a.Nval[(e * P + gp) * N + n] = Nv[n];
  }
}

// constant-strain shapes: one thread per element
__global__ void __launch_bounds__(128) construct_tria_kernel(const ConstructArgs a) {  // elasticelements.jl:38-46
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.ne) return;
  double x[3], y[3];
#pragma unroll
  for (int n = 0; n < 3; n++) {
    const long long nd = a.nodes[e * 3 + n] - 1;
    x[n] = __ldg(a.coords + nd * a.ldc);
    y[n] = __ldg(a.coords + nd * a.ldc + 1);
  }
  const double Delta = x[0] * y[1] - x[1] * y[0] - x[0] * y[2] + x[2] * y[0] + x[1] * y[2] - x[2] * y[1];
  double* g = a.gradN + e * 6;
  g[0] = (y[1] - y[2]) / Delta, g[1] = (y[2] - y[0]) / Delta, g[2] = (y[0] - y[1]) / Delta;
  g[3] = (x[2] - x[1]) / Delta, g[4] = (x[0] - x[2]) / Delta, g[5] = (x[1] - x[0]) / Delta;
  const double A = fabs(Delta) / 2;
  a.wgt[e] = A;
  if (a.V) a.V[e] = A;
  if (a.Nval) a.Nval[e * 3] = a.Nval[e * 3 + 1] = a.Nval[e * 3 + 2] = 1.0 / 3.0;
}
__global__ void __launch_bounds__(128) construct_tet_kernel(const ConstructArgs a) {  // elasticelements.jl:85-94
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.ne) return;
  double X[4][3];
#pragma unroll
  for (int n = 0; n < 4; n++) {
    const long long nd = a.nodes[e * 4 + n] - 1;
#pragma unroll
    for (int i = 0; i < 3; i++) X[n][i] = __ldg(a.coords + nd * a.ldc + i);
  }
  // A = [1 p]; C = inv(A): column a of C holds the coefficients of N_a = C[0][a] + C[1..3][a]·x, so rows 2..4 are ∇N.
  // With the edge matrix E[b][i] = p_{b+1,i} - p_{0,i}: det(A) = det(E), [∇N_1 ∇N_2 ∇N_3] = inv(E), ∇N_0 = -Σ.
  double E[3][3];
#pragma unroll
  for (int b = 0; b < 3; b++)
#pragma unroll
    for (int i = 0; i < 3; i++) E[b][i] = X[b + 1][i] - X[0][i];
  double G[3][8];  // G[i][b]: solve E · G = I
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int b = 0; b < 3; b++) G[i][b] = (i == b) ? 1.0 : 0.0;
  const double det = det_small<3>(E);
  solve_small<3, 3>(E, G);
  double* g = a.gradN + e * 12;
#pragma unroll
  for (int i = 0; i < 3; i++) {
    g[i * 4 + 0] = -(G[i][0] + G[i][1] + G[i][2]);
    g[i * 4 + 1] = G[i][0], g[i * 4 + 2] = G[i][1], g[i * 4 + 3] = G[i][2];
  }
  const double V = det / 6;  // signed, no abs (elasticelements.jl:92)
  a.wgt[e] = V;
  if (a.V) a.V[e] = V;
  if (a.Nval) a.Nval[e * 4] = a.Nval[e * 4 + 1] = a.Nval[e * 4 + 2] = a.Nval[e * 4 + 3] = 0.25;
}

// V += wgt in the reference's loop order (ii outermost, kk innermost; the storage index has ii fastest)
__global__ void __launch_bounds__(128) construct_volume_kernel(const ConstructArgs a) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.ne) return;
  const int n = a.n1d, D = a.n_wfac;
  double V = 0.0;
  if (n > 0) {
    const int nk = D == 3 ? n : 1;
    for (int ii = 0; ii < n; ii++)
      for (int jj = 0; jj < n; jj++)
        for (int kk = 0; kk < nk; kk++) V += a.wgt[e * a.P + ii + n * (jj + n * kk)];
  } else {
    for (int gp = 0; gp < a.P; gp++) V += a.wgt[e * a.P + gp];
  }
  a.V[e] = V;
}

}  // namespace

cudaError_t launch_construct(const ConstructArgs& a, cudaStream_t s) {
  if (a.ne == 0) return cudaSuccess;
  const unsigned T = 128;
  const unsigned g_el = (unsigned)((a.ne + T - 1) / T);
  const unsigned g_gp = (unsigned)((a.ne * a.P + T - 1) / T);
  switch (a.shape) {
    case CS_TRIA3: construct_tria_kernel<<<g_el, T, 0, s>>>(a); return cudaGetLastError();
    case CS_TET4: construct_tet_kernel<<<g_el, T, 0, s>>>(a); return cudaGetLastError();
    case CS_QUAD4: construct_iso_kernel<CS_QUAD4><<<g_gp, T, 0, s>>>(a); break;
    case CS_HEX8: construct_iso_kernel<CS_HEX8><<<g_gp, T, 0, s>>>(a); break;
    case CS_WDG6: construct_iso_kernel<CS_WDG6><<<g_gp, T, 0, s>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess || !a.V) return err;
  construct_volume_kernel<<<g_el, T, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace ad4sm
