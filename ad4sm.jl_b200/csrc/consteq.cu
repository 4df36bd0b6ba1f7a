// consteq.cu -- constraint-equation evaluation for a fixed catalogue of closed-form constraints.
//
// The reference evaluates, for every ConstEq (func, iDoFs) (/root/reference/src/solvers.jl:17-22, 54-73):
//     ϕ = eqn.func(adiff.D2(u[iDoFs]));   veqs[ii] = val(ϕ);   reqs[iDoFs, ii] = grad(ϕ);   Keqs[iDoFs, iDoFs] += λ[ii] hess(ϕ)
// `func` is an arbitrary Julia closure and cannot cross a C ABI; what can is the family the closures used for periodic
// boundary conditions, ties, rigid links and prescribed averages belong to -- polynomials of degree <= 2 in the dofs:
//     ϕ(v) = ½ vᵀ Q v + cᵀ v - c0,      v = u[iDoFs]        (Q symmetric, possibly absent: a LINEAR constraint)
// whose D2 evaluation is  val = ϕ(v),  grad = Q v + c,  hess = Q.  One thread per equation; the outputs keep the
// per-equation layout of the inputs (gradient aligned with iDoFs, λ-scaled Hessian aligned with Q), which the host side
// turns into the two sparse matrices exactly like the reference's indexed assignments do.
#include "internal.h"

namespace ad4sm {

__global__ void __launch_bounds__(128) consteq_kernel(const ConstEqArgs a) {
  const long long ii = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (ii >= a.n_eqs) return;
  const long long d0 = a.dof_ptr[ii], n = a.dof_ptr[ii + 1] - d0;
  const long long* dofs = a.dofs + d0;
  const double* c = a.c + d0;
  const double* Q = (a.q_ptr && a.q_ptr[ii + 1] > a.q_ptr[ii]) ? a.Q + a.q_ptr[ii] : nullptr;
  const double lam = a.lambda ? a.lambda[ii] : 0.0;
  double val = 0.0;
  for (long long k = 0; k < n; k++) {
    const double vk = a.u[dofs[k] - 1];
    double g = c[k];
    if (Q) {
      double qv = 0.0;
      for (long long l = 0; l < n; l++) qv += Q[k + l * n] * a.u[dofs[l] - 1];  // column-major, symmetric
      g += qv;
      val += 0.5 * qv * vk;
    }
    val += c[k] * vk;
    a.grad[d0 + k] = g;
  }
  a.veqs[ii] = val - a.c0[ii];
  if (Q && a.hess) {
    double* H = a.hess + a.q_ptr[ii];
    for (long long t = 0; t < n * n; t++) H[t] = lam * Q[t];
  }
}

cudaError_t launch_consteq(const ConstEqArgs& a, cudaStream_t s) {
  if (a.n_eqs == 0) return cudaSuccess;
  consteq_kernel<<<(unsigned)((a.n_eqs + 127) / 128), 128, 0, s>>>(a);
  return cudaGetLastError();
}

}  // namespace ad4sm
