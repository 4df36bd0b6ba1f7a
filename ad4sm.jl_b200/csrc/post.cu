// post.cu -- batched post-processing with first derivatives only: what the reference evaluates element
// by element on the host after every load step (/root/reference/src/elements.toolkit.jl):
//   getF(elem, u) / get∇u(elem, u)   :8-27    volume average  I + Σ_gp wgt ∇u_gp / V
//   detJ(elem, u)                     :33-47   Σ_gp wgt det F_gp / V
//   getV(elem, u)                     :52-58   Σ_gp wgt det F_gp                      (getV(elems,u) = their sum, :59)
//   getσ(elem, u)                     :68-87   Σ_gp wgt (1/J P Fᵀ)_gp / V,  P = ∂ϕ/∂F  (D1 duals over F in the reference)
//   getP(elem, u)                     :117-132 Σ_gp wgt Pᵀ_gp / V
// with V = elem.V = Σ_gp wgt for every working constructor (elasticelements.jl:92, 143-161).  One thread per element; the
// material's first derivative comes from the same invariant-level tangent code K1 uses (materials.cuh).
#include "element.cuh"
#include "internal.h"

namespace ad4sm {

template <int D, int N, int MF> __global__ void __launch_bounds__(128) post_kernel(const PostArgs a) {
  constexpr int DD = D * D;
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= a.ne) return;
  const int P = a.P;
  double ue[N * D];
#pragma unroll
  for (int n = 0; n < N; n++) {
    const long long nd = a.nodes[e * N + n];
#pragma unroll
    for (int j = 0; j < D; j++) ue[n * D + j] = a.u[nd * a.dpn + j];
  }
  double Fs[DD], Ss[DD], Ps[DD], Js = 0.0, V = 0.0;
#pragma unroll
  for (int i = 0; i < DD; i++) Fs[i] = Ss[i] = Ps[i] = 0.0;
  const double* gN = a.gradN + e * (long long)(D * P * N);
  for (int gp = 0; gp < P; gp++) {
    double G[D * N], F[DD];
#pragma unroll
    for (int k = 0; k < D; k++)
#pragma unroll
      for (int n = 0; n < N; n++) G[k * N + n] = gN[(k * P + gp) * N + n];
    deformation_gradient<D, N>(G, ue, F);  // F[k*D+j] = F_jk: the reference's column-major linear index
    const double w = a.wgt[e * P + gp];
    double J;
    if (D == 2)
      J = F[0] * F[3] - F[2] * F[1];                                   // F11 F22 - F12 F21      (toolkit:33)
    else
      J = F[0] * (F[4] * F[8] - F[7] * F[5]) - F[3] * (F[1] * F[8] - F[7] * F[2]) + F[6] * (F[1] * F[5] - F[4] * F[2]);  // :34-36
    V += w;
    Js += w * J;
#pragma unroll
    for (int i = 0; i < DD; i++) Fs[i] += w * (F[i] - ((i % (D + 1)) == 0 ? 1.0 : 0.0));  // ∇u = F - I   (:21-26)
    if (a.want_stress) {
      Tangent<D> t;
      material_tangent<D, MF>(a.mat, F, nullptr, t);
      const double iJ = 1.0 / J;
#pragma unroll
      for (int i = 0; i < D; i++)
#pragma unroll
        for (int j = 0; j < D; j++) {
          double s = 0.0;  // σ_ij = Σ_k (1/J P_ik) F_jk                                  (:86)
#pragma unroll
          for (int k = 0; k < D; k++) s += (iJ * t.P[k * D + i]) * F[k * D + j];
          Ss[j * D + i] += w * s;
          Ps[j * D + i] += w * t.P[i * D + j];  // getP returns P1K' (:131): entry (i,j) of the result is P_ji
        }
    }
  }
  const long long o = a.index ? a.index[e] : a.slot0 + e;  // position in `elems`
  if (a.F) {
#pragma unroll
    for (int i = 0; i < DD; i++) a.F[o * DD + i] = Fs[i] / V + ((i % (D + 1)) == 0 ? 1.0 : 0.0);
  }
  if (a.detJ) a.detJ[o] = Js / V;
  if (a.vol) a.vol[o] = Js;
  if (a.sigma) {
#pragma unroll
    for (int i = 0; i < DD; i++) a.sigma[o * DD + i] = Ss[i] / V;
  }
  if (a.Pk) {
#pragma unroll
    for (int i = 0; i < DD; i++) a.Pk[o * DD + i] = Ps[i] / V;
  }
}

template <int D, int N> static cudaError_t launch_post_mf(const PostArgs& a, cudaStream_t s) {
  const unsigned grid = (unsigned)((a.ne + 127) / 128);
  switch (a.mf) {
    case MF_HOOKE: post_kernel<D, N, MF_HOOKE><<<grid, 128, 0, s>>>(a); break;
    case MF_NEO: post_kernel<D, N, MF_NEO><<<grid, 128, 0, s>>>(a); break;
    case MF_MOONEY: post_kernel<D, N, MF_MOONEY><<<grid, 128, 0, s>>>(a); break;
    case MF_OGDEN: post_kernel<D, N, MF_OGDEN><<<grid, 128, 0, s>>>(a); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

cudaError_t launch_post(const PostArgs& a, cudaStream_t s) {
  if (a.ne == 0) return cudaSuccess;
#define SHAPE(DD_, NN_) \
  if (a.D == DD_ && a.N == NN_) return launch_post_mf<DD_, NN_>(a, s);
  SHAPE(2, 3) SHAPE(2, 4) SHAPE(3, 4) SHAPE(3, 6) SHAPE(3, 8)
#undef SHAPE
  return cudaErrorInvalidValue;
}

}  // namespace ad4sm
