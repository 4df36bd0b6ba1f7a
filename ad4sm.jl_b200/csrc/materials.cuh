// materials.cuh -- strain-energy densities of AD4SM.Materials with their first and second
// derivatives w.r.t. the deformation gradient, evaluated per Gauss point in registers.
//
// Reference formulas (energies only; derivatives come from adiff duals there):
//   Hooke / Hooke2D             /root/reference/src/elasticmaterials.jl:202-245
//   HyperEla dispatcher         /root/reference/src/elasticmaterials.jl:138-147
//   invariants                  /root/reference/src/materials.jl:98-114
//   NeoHooke / MooneyRivlin     /root/reference/src/elasticmaterials.jl:173-197
//   get1stinvariants/hyddev     /root/reference/src/materials.jl:117-155, elasticmaterials.jl:264-333
//   PhaseField ATn / DHn        /root/reference/src/phasefieldmaterials.jl:20-135
//
// Output convention: F is addressed like the reference's linear index, F[k*D+j] = F_jk.
//   P[I]      = dψ/dF_I                      (I = k*D+j)
//   A[pk(I,J)]= d²ψ/dF_I dF_J, I<=J, packed column-major upper triangle pk = J(J+1)/2+I
// which is exactly the (g, h) payload of the reference's `getϕ(adiff.D2(valF), mat)::D2{9,45}`
// (elasticelements.jl:346).
#pragma once
#include "dual2.cuh"

namespace ad4sm {

enum MatKind { MAT_HOOKE = 1, MAT_HOOKE2D = 2, MAT_NEOHOOKE = 3, MAT_MOONEY = 4, MAT_OGDEN = 5, MAT_PHASEFIELD = 6 };
enum MatFlag { F_SMALL = 1u, F_PLANE_STRESS = 2u, F_DHN = 4u, F_PF_GRAD_INTENDED = 8u, F_NH2D_EXT = 16u };

// material families the kernels are specialised on
enum MatFamily {
  MF_HOOKE = 0,  // Hooke, Hooke2D and PhaseField over them: quadratic in the strain
  MF_NEO = 1,    // NeoHooke (f(I1, J))
  MF_MOONEY = 2, // MooneyRivlin (f(I1, I2, J))
  MF_PFNEO = 3,  // PhaseField{NeoHooke,:ATn}
  MF_OGDEN = 4   // Ogden (principal stretches)
};

struct MatDev {
  int kind, base;
  unsigned flags;
  int pad;
  double p[8];
};

// phase-field state at a Gauss point (u-phase): value of d, |∇d|²
struct PfPoint {
  double d, gd2;
};

template <int D> struct Tangent {
  static constexpr int DD = D * D;
  static constexpr int NA = DD * (DD + 1) / 2;
  double psi;
  double P[DD];
  double A[NA];
};

AD4SM_HD constexpr int pk(int I, int J) { return I <= J ? J * (J + 1) / 2 + I : I * (I + 1) / 2 + J; }

AD4SM_DI void lame(double Es, double nu, double& lam, double& mu) {  // elasticmaterials.jl:211-212
  lam = Es * nu / (1 + nu) / (1 - 2 * nu);
  mu = Es / 2 / (1 + nu);
}

// ---------------------------------------------------------------------------------------------
// strain E (symmetric D x D, full storage E[k*D+j])   elasticmaterials.jl:204-208
// ---------------------------------------------------------------------------------------------
template <int D> AD4SM_DI void strain(const double* F, bool small, double* E) {
  if (small) {
#pragma unroll
    for (int j = 0; j < D; j++)
#pragma unroll
      for (int k = 0; k < D; k++) E[k * D + j] = (F[k * D + j] + F[j * D + k] - (j == k ? 2.0 : 0.0)) / 2.0;
  } else {
#pragma unroll
    for (int j = 0; j < D; j++)
#pragma unroll
      for (int k = j; k < D; k++) {
        double c = 0.0;
#pragma unroll
        for (int i = 0; i < D; i++) c += F[j * D + i] * F[k * D + i];  // C_jk = Σ_i F_ij F_ik
        const double e = (c - (j == k ? 1.0 : 0.0)) / 2.0;
        E[k * D + j] = e;
        E[j * D + k] = e;
      }
  }
}

// ψ = p ΣE_ii² + q Σ_{i<j} E_ii E_jj + s Σ_{i<j} E_ij²   and its F-derivatives.
// S = dψ/dE (symmetric), ℂ_iiii = 2p, ℂ_iijj = q, ℂ_ijij = ℂ_ijji = s/2.
template <int D>
AD4SM_DI void quad_strain_tangent(const double* F, const double* E, bool small, double p, double q, double s,
                                  Tangent<D>& t) {
  constexpr int DD = D * D;
  double S[DD];
  double tr = 0.0, psi = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) tr += E[i * D + i];
#pragma unroll
  for (int i = 0; i < D; i++) {
    S[i * D + i] = 2.0 * p * E[i * D + i] + q * (tr - E[i * D + i]);
    psi += p * E[i * D + i] * E[i * D + i];
#pragma unroll
    for (int j = i + 1; j < D; j++) {
      S[j * D + i] = S[i * D + j] = s * E[j * D + i];
      psi += q * E[i * D + i] * E[j * D + j] + s * E[j * D + i] * E[j * D + i];
    }
  }
  t.psi = psi;
  const double hs = 0.5 * s;
  if (small) {
#pragma unroll
    for (int I = 0; I < DD; I++) t.P[I] = S[I];
#pragma unroll
    for (int J = 0; J < DD; J++)
#pragma unroll
      for (int I = 0; I <= J; I++) {
        const int j = I % D, k = I / D, l = J % D, m = J / D;
        double a = 0.0;
        if (j == k && l == m) a = (j == l) ? 2.0 * p : q;
        if (j != k) a = hs * (((j == l && k == m) ? 1.0 : 0.0) + ((j == m && k == l) ? 1.0 : 0.0));
        t.A[pk(I, J)] = a;
      }
  } else {
    // P = F S
#pragma unroll
    for (int k = 0; k < D; k++)
#pragma unroll
      for (int j = 0; j < D; j++) {
        double a = 0.0;
#pragma unroll
        for (int pp = 0; pp < D; pp++) a += F[pp * D + j] * S[k * D + pp];
        t.P[k * D + j] = a;
      }
    // b = F F^T (for the [k==m] term we need Σ_{p≠k} F_jp F_lp = b_jl - F_jk F_lk)
    double b[DD];
#pragma unroll
    for (int j = 0; j < D; j++)
#pragma unroll
      for (int l = j; l < D; l++) {
        double a = 0.0;
#pragma unroll
        for (int pp = 0; pp < D; pp++) a += F[pp * D + j] * F[pp * D + l];
        b[l * D + j] = b[j * D + l] = a;
      }
#pragma unroll
    for (int J = 0; J < DD; J++)
#pragma unroll
      for (int I = 0; I <= J; I++) {
        const int j = I % D, k = I / D, l = J % D, m = J / D;
        double a = F[k * D + j] * F[m * D + l] * ((k == m) ? 2.0 * p : q);
        if (j == l) a += S[m * D + k];
        if (k == m)
          a += hs * (b[l * D + j] - F[k * D + j] * F[k * D + l]);
        else
          a += hs * F[m * D + j] * F[k * D + l];
        t.A[pk(I, J)] = a;
      }
  }
}

// coefficients (p, q, s) of the Hooke family incl. PhaseField degradation.
// c  : E33 = c (E11+E22) (plane-stress phase-field invariants, elasticmaterials.jl:327), else 0
// g  : (1-d)², 1 for the undamaged materials
template <int D>
AD4SM_DI void hooke_family_coeffs(const MatDev& m, const double* E, double g, double& p, double& q, double& s) {
  const int kind = (m.kind == MAT_PHASEFIELD) ? m.base : m.kind;
  const double Es = m.p[0], nu = m.p[1];
  if (m.kind != MAT_PHASEFIELD) {
    if (kind == MAT_HOOKE2D && (m.flags & F_PLANE_STRESS)) {  // elasticmaterials.jl:244
      p = Es / (1 - nu * nu);
      q = nu * p;
      s = Es / (1 + nu);
    } else {  // :214-216, :232
      double lam, mu;
      lame(Es, nu, lam, mu);
      p = mu + lam / 2;
      q = lam;
      s = 2 * mu;
    }
    return;
  }
  double lam, mu;
  lame(Es, nu, lam, mu);
  const double c = (kind == MAT_HOOKE2D && (m.flags & F_PLANE_STRESS)) ? nu / (nu - 1) : 0.0;
  double tr = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) tr += E[i * D + i];
  const bool pos = (1.0 + c) * tr >= 0.0;  // I1 >= 0
  const double amu = g;
  double kappa;
  if (!(m.flags & F_DHN)) {  // phasefieldmaterials.jl:29-33
    const double al = pos ? g : 1.0;
    kappa = al * (lam / 2) * (1 + c) * (1 + c) + amu * mu * c * c;
  } else {  // :65-69, K = λ + μ/3 as written
    const double K = lam + mu / 3;
    const double ak = pos ? g : 1.0;
    kappa = ak * (K / 2) * (1 + c) * (1 + c) + amu * mu * (c * c - (1 + c) * (1 + c) / 3.0);
  }
  p = kappa + amu * mu;
  q = 2 * kappa;
  s = 2 * amu * mu;
}

// ---------------------------------------------------------------------------------------------
// generic assembly of P and A from NV scalar "invariants" with known F-derivatives
// ---------------------------------------------------------------------------------------------
template <int DD, int NV>
AD4SM_DI void chain_invariants(const Dual2<NV>& f, const double (*dV)[DD], double* P, double* A) {
  double W[NV][DD];
#pragma unroll
  for (int v = 0; v < NV; v++)
#pragma unroll
    for (int I = 0; I < DD; I++) {
      double a = 0.0;
#pragma unroll
      for (int w = 0; w < NV; w++) a += f.h[pk(v, w)] * dV[w][I];
      W[v][I] = a;
    }
#pragma unroll
  for (int I = 0; I < DD; I++) {
    double a = 0.0;
#pragma unroll
    for (int v = 0; v < NV; v++) a += f.g[v] * dV[v][I];
    P[I] = a;
  }
#pragma unroll
  for (int J = 0; J < DD; J++)
#pragma unroll
    for (int I = 0; I <= J; I++) {
      double a = 0.0;
#pragma unroll
      for (int v = 0; v < NV; v++) a += dV[v][I] * W[v][J];
      A[pk(I, J)] = a;
    }
}

AD4SM_DI double det3(const double* F) {
  return F[0] * (F[4] * F[8] - F[7] * F[5]) - F[3] * (F[1] * F[8] - F[7] * F[2]) + F[6] * (F[1] * F[5] - F[4] * F[2]);
}
// cof_jk = dJ/dF_jk (cyclic formula)
AD4SM_DI void cof3(const double* F, double* c) {
#pragma unroll
  for (int k = 0; k < 3; k++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      const int j1 = (j + 1) % 3, j2 = (j + 2) % 3, k1 = (k + 1) % 3, k2 = (k + 2) % 3;
      c[k * 3 + j] = F[k1 * 3 + j1] * F[k2 * 3 + j2] - F[k2 * 3 + j1] * F[k1 * 3 + j2];
    }
}
// A += fJ * d²J/dF dF  (3-D): ε_jln ε_kmp F_np
AD4SM_DI void add_d2J3(const double* F, double fJ, double* A) {
#pragma unroll
  for (int J = 0; J < 9; J++)
#pragma unroll
    for (int I = 0; I < J; I++) {
      const int j = I % 3, k = I / 3, l = J % 3, m = J / 3;
      if (j != l && k != m) {
        const int n = 3 - j - l, pp = 3 - k - m;
        const double sg = ((((l - j + 3) % 3) == 1) ? 1.0 : -1.0) * ((((m - k + 3) % 3) == 1) ? 1.0 : -1.0);
        A[pk(I, J)] += fJ * sg * F[pp * 3 + n];
      }
    }
}

// NeoHooke / MooneyRivlin energy on duals of (I1, [I2,] J)   elasticmaterials.jl:173-197
template <int NV, bool USE_I2>
AD4SM_DI Dual2<NV> hyper_energy(const MatDev& m, const Dual2<NV>& I1, const Dual2<NV>& I2, const Dual2<NV>& J) {
  if (USE_I2) {
    const double C1 = m.p[0], C2 = m.p[1], K = m.p[2];
    if (K < 0) return C1 * (I1 - 3.0) + C2 * (I2 - 3.0);
    const double cb = cbrt(J.v);
    const double jm23 = 1.0 / (cb * cb);
    Dual2<NV> a = d2_pow_given(J, -2.0 / 3, jm23);
    Dual2<NV> b = d2_pow_given(J, -4.0 / 3, jm23 * jm23);
    return C1 * (I1 * a - 3.0) + C2 * (I2 * b - 3.0) + K * d2_sq(J - 1.0);
  } else {
    const double C1 = m.p[0], K = m.p[1];
    if (K < 0) return C1 * (I1 - 3.0);
    const double cb = cbrt(J.v);
    Dual2<NV> a = d2_pow_given(J, -2.0 / 3, 1.0 / (cb * cb));
    return C1 * (I1 * a - 3.0) + K * d2_sq(J - 1.0);
  }
}

// 3-D hyperelastic tangent.  J = sqrt(I3) = |det F| like the reference (elasticmaterials.jl:179,192).
template <bool USE_I2> AD4SM_DI void hyper_tangent3(const MatDev& m, const double* F, Tangent<3>& t) {
  constexpr int NV = USE_I2 ? 3 : 2;
  constexpr int VJ = NV - 1;
  double dV[NV][9];
  double I1 = 0.0;
#pragma unroll
  for (int I = 0; I < 9; I++) I1 += F[I] * F[I];
  const double dj = det3(F);
  const double sgn = dj < 0 ? -1.0 : 1.0;
  double cf[9];
  cof3(F, cf);
#pragma unroll
  for (int I = 0; I < 9; I++) {
    dV[0][I] = 2.0 * F[I];
    dV[VJ][I] = sgn * cf[I];
  }
  double C[9], b[9], I2 = 0.0;
  if (USE_I2) {
#pragma unroll
    for (int j = 0; j < 3; j++)
#pragma unroll
      for (int k = j; k < 3; k++) {
        double c = 0.0, bb = 0.0;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          c += F[j * 3 + i] * F[k * 3 + i];   // C_jk = Σ_i F_ij F_ik
          bb += F[i * 3 + j] * F[i * 3 + k];  // b_jk = Σ_i F_ji F_ki
        }
        C[k * 3 + j] = C[j * 3 + k] = c;
        b[k * 3 + j] = b[j * 3 + k] = bb;
      }
    // materials.jl:112
    I2 = C[0] * C[4] + C[4] * C[8] + C[0] * C[8] - C[3] * C[3] - C[6] * C[6] - C[7] * C[7];
    // dI2/dF = 2 (I1 F - F C)
#pragma unroll
    for (int k = 0; k < 3; k++)
#pragma unroll
      for (int j = 0; j < 3; j++) {
        double fc = 0.0;
#pragma unroll
        for (int pp = 0; pp < 3; pp++) fc += F[pp * 3 + j] * C[k * 3 + pp];
        dV[1][k * 3 + j] = 2.0 * (I1 * F[k * 3 + j] - fc);
      }
  }
  Dual2<NV> dI1 = d2_var<NV>(I1, 0);
  Dual2<NV> dI2 = USE_I2 ? d2_var<NV>(I2, 1) : d2_const<NV>(0.0);
  Dual2<NV> dJ = d2_var<NV>(sgn * dj, VJ);
  Dual2<NV> f = hyper_energy<NV, USE_I2>(m, dI1, dI2, dJ);
  t.psi = f.v;
  chain_invariants<9, NV>(f, dV, t.P, t.A);
  // second derivatives of the invariants
  const double f1 = f.g[0], fJ = f.g[VJ];
#pragma unroll
  for (int I = 0; I < 9; I++) t.A[pk(I, I)] += 2.0 * f1;
  add_d2J3(F, fJ * sgn, t.A);
  if (USE_I2) {
    const double f2 = 2.0 * f.g[1];
#pragma unroll
    for (int J = 0; J < 9; J++)
#pragma unroll
      for (int I = 0; I <= J; I++) {
        const int j = I % 3, k = I / 3, l = J % 3, mm = J / 3;
        double a = 2.0 * F[k * 3 + j] * F[mm * 3 + l] - F[mm * 3 + j] * F[k * 3 + l];
        if (j == l && k == mm) a += I1;
        if (j == l) a -= C[k * 3 + mm];
        if (k == mm) a -= b[l * 3 + j];
        t.A[pk(I, J)] += f2 * a;
      }
  }
}

// 2-D hyperelastic tangent on (t = ΣF², J2 = det F)  elasticmaterials.jl:138-147, materials.jl:98-104
template <bool USE_I2> AD4SM_DI void hyper_tangent2(const MatDev& m, const double* F, Tangent<2>& t) {
  const double tt = F[0] * F[0] + F[1] * F[1] + F[2] * F[2] + F[3] * F[3];
  const double j2 = F[0] * F[3] - F[1] * F[2];
  double dV[2][4];
#pragma unroll
  for (int I = 0; I < 4; I++) dV[0][I] = 2.0 * F[I];
  dV[1][0] = F[3];
  dV[1][3] = F[0];
  dV[1][1] = -F[2];
  dV[1][2] = -F[1];
  const double K = USE_I2 ? m.p[2] : m.p[1];
  Dual2<2> T = d2_var<2>(tt, 0), J2 = d2_var<2>(j2, 1);
  Dual2<2> J2sq = d2_sq(J2);
  Dual2<2> C33 = K < 0 ? d2_inv(J2sq) : d2_const<2>(1.0);
  Dual2<2> I1 = T + C33;
  Dual2<2> I2 = J2sq + T * C33;
  Dual2<2> I3 = J2sq * C33;
  Dual2<2> f;
  if (K < 0)
    f = hyper_energy<2, USE_I2>(m, I1, I2, d2_const<2>(1.0));
  else
    f = hyper_energy<2, USE_I2>(m, I1, I2, d2_sqrt(I3));
  t.psi = f.v;
  chain_invariants<4, 2>(f, dV, t.P, t.A);
#pragma unroll
  for (int I = 0; I < 4; I++) t.A[pk(I, I)] += 2.0 * f.g[0];
  t.A[pk(0, 3)] += f.g[1];  // d²J2/dF11 dF22 = 1
  t.A[pk(1, 2)] -= f.g[1];  // d²J2/dF21 dF12 = -1
}

// PhaseField{NeoHooke,:ATn}   phasefieldmaterials.jl:37-54  (+ 2-D plane-strain embedding)
template <int D> AD4SM_DI void pfneo_tangent(const MatDev& m, const double* F, double g, Tangent<D>& t) {
  constexpr int DD = D * D;
  const double C1 = m.p[0], K = m.p[1];
  double I1 = (D == 2) ? 1.0 : 0.0;
#pragma unroll
  for (int I = 0; I < DD; I++) I1 += F[I] * F[I];
  double dV[2][DD];
#pragma unroll
  for (int I = 0; I < DD; I++) dV[0][I] = 2.0 * F[I];
  double J;
  if (D == 3) {
    J = det3(F);
    cof3(F, dV[1]);
  } else {
    J = F[0] * F[3] - F[1] * F[2];
    dV[1][0] = F[3];
    dV[1][3] = F[0];
    dV[1][1] = -F[2];
    dV[1][2] = -F[1];
  }
  Dual2<2> dI1 = d2_var<2>(I1, 0), dJ = d2_var<2>(J, 1);
  Dual2<2> f = C1 * (dI1 - 3.0 - 2.0 * d2_log(dJ)) + K * d2_sq(dJ - 1.0);
  if (J >= 1.0) f = g * f;
  t.psi = f.v;
  chain_invariants<DD, 2>(f, dV, t.P, t.A);
#pragma unroll
  for (int I = 0; I < DD; I++) t.A[pk(I, I)] += 2.0 * f.g[0];
  if (D == 3) {
    add_d2J3(F, f.g[1], t.A);
  } else {
    t.A[pk(0, 3)] += f.g[1];
    t.A[pk(1, 2)] -= f.g[1];
  }
}

// ---------------------------------------------------------------------------------------------
// Ogden (principal stretches)   elasticmaterials.jl:148-172
//   λ_i = sqrt(σ_i(FᵀF));  K < 0: ψ = μ/α (Σ λ_i^α − 3);  else J = Π λ_i, ψ = μ/α (Σ λ_i^α J^(−α/3) − 3) + K (J−1)².
//   2-D: F is embedded with F33 = 1/det F (K < 0, plane stress) or 1 (:155-160), i.e. λ3 = 1/(λ1 λ2) or 1.
// The reference defines only this scalar energy: its `Ogden` cannot be constructed (struct/`new` arity mismatch,
// :114-120) and `svdvals` has no method for duals (:162), so r and Kt have no reference value.  The
// derivatives here are the exact ones of that energy, from the spectral representation F = Σ λ_i n_i ⊗ N_i:
//   P = Σ_i f_i n_i⊗N_i
//   A = Σ_ij f_ij (n_i⊗N_i)⊗(n_j⊗N_j) + Σ_{i≠j} a_ij (n_i⊗N_j)⊗(n_i⊗N_j) + Σ_{i≠j} b_ij (n_i⊗N_j)⊗(n_j⊗N_i)
//   a_ij = (λ_i f_i − λ_j f_j)/(λ_i² − λ_j²),  b_ij = (λ_j f_i − λ_i f_j)/(λ_i² − λ_j²)        (Ogden 1984, §6.1.4)
// Because λ_i f_i = μ s λ_i^α + c with s and c common to all i (s = J^(−α/3) or 1), the quotients are divided
// differences of x^(α/2) and x^((α−2)/2) in x = λ², evaluated without cancellation -- repeated stretches (u = 0:
// λ = (1,1,1)) are regular points of these formulas, and any orthonormal eigenbasis may be used there.
// ---------------------------------------------------------------------------------------------
// (x_i^p − x_j^p)/(x_i − x_j) for x > 0, given xip = x_i^p and xjp = x_j^p (no further pow calls)
AD4SM_DI double divdiff_pow(double xi, double xj, double xip, double xjp, double p) {
  const double t = (xi - xj) / xj;
  if (fabs(t) < 2e-3) {  // series of ((1+t)^p − 1)/t, truncation error t^5/720·|p…(p−5)| < 1e-16
    return (xjp / xj) * p *
           (1.0 + (p - 1.0) * t / 2.0 *
                      (1.0 + (p - 2.0) * t / 3.0 * (1.0 + (p - 3.0) * t / 4.0 * (1.0 + (p - 4.0) * t / 5.0))));
  }
  return (xip - xjp) / (xi - xj);  // cancellation error <= eps / 2e-3
}

// Eigen-decomposition of a symmetric DxD matrix S (full storage, destroyed): cyclic Jacobi, at most 8 sweeps, stopped
// when the off-diagonal mass is below 1e-30 of the diagonal (deterministic; robust for repeated eigenvalues).
// V columns = eigenvectors, S diagonal = eigenvalues.
template <int D> AD4SM_DI void jacobi_eig(double* S, double* V) {
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = 0; j < D; j++) V[i * D + j] = (i == j) ? 1.0 : 0.0;
  for (int sweep = 0; sweep < (D == 2 ? 1 : 8); sweep++) {
    if (D == 3) {
      const double off = S[1] * S[1] + S[2] * S[2] + S[5] * S[5];
      const double dg = S[0] * S[0] + S[4] * S[4] + S[8] * S[8];
      if (off <= 1e-30 * dg) break;
    }
#pragma unroll
    for (int p = 0; p < D - 1; p++)
#pragma unroll
      for (int q = p + 1; q < D; q++) {
        const double apq = S[p * D + q];
        if (apq != 0.0) {
          const double theta = (S[q * D + q] - S[p * D + p]) / (2.0 * apq);
          const double t = (theta >= 0.0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          const double c = 1.0 / sqrt(t * t + 1.0), sn = t * c;
#pragma unroll
          for (int k = 0; k < D; k++) {  // S <- S J
            const double skp = S[k * D + p], skq = S[k * D + q];
            S[k * D + p] = c * skp - sn * skq;
            S[k * D + q] = sn * skp + c * skq;
          }
#pragma unroll
          for (int k = 0; k < D; k++) {  // S <- Jᵀ S
            const double spk = S[p * D + k], sqk = S[q * D + k];
            S[p * D + k] = c * spk - sn * sqk;
            S[q * D + k] = sn * spk + c * sqk;
          }
#pragma unroll
          for (int k = 0; k < D; k++) {
            const double vkp = V[k * D + p], vkq = V[k * D + q];
            V[k * D + p] = c * vkp - sn * vkq;
            V[k * D + q] = sn * vkp + c * vkq;
          }
        }
      }
  }
}

// energy on duals of the D in-plane stretches (the third stretch of the 2-D embedding is a function of them).
// lam_a[i] = λ_i^α is supplied by the caller (one exp/log pair per stretch for everything in this material).
template <int D> AD4SM_DI Dual2<D> ogden_energy(const MatDev& m, const Dual2<D>* lam, const double* lam_a, double& s_out) {
  const double alpha = m.p[0], mu = m.p[1], K = m.p[2];
  Dual2<D> S = d2_pow_given(lam[0], alpha, lam_a[0]);
#pragma unroll
  for (int i = 1; i < D; i++) S = S + d2_pow_given(lam[i], alpha, lam_a[i]);
  Dual2<D> J = lam[0];
#pragma unroll
  for (int i = 1; i < D; i++) J = J * lam[i];
  if (D == 2) S = (K < 0) ? S + d2_pow(J, -alpha) : S + 1.0;  // λ3 = 1/(λ1 λ2) or 1
  if (K < 0) {
    s_out = 1.0;
    return (mu / alpha) * (S - 3.0);
  }
  const Dual2<D> sc = d2_pow(J, -alpha / 3.0);  // 3-D: J = λ1λ2λ3; 2-D (K >= 0): λ3 = 1, J = λ1λ2
  s_out = sc.v;
  return (mu / alpha) * (S * sc - 3.0) + K * d2_sq(J - 1.0);
}

// A += c · sym(x yᵀ) restricted to the packed upper triangle:  A[pk(I,J)] += x[I]·y[J]   (caller makes it symmetric)
template <int DD> AD4SM_DI void rank1_packed(double* A, const double* x, const double* y) {
#pragma unroll
  for (int J = 0; J < DD; J++)
#pragma unroll
    for (int I = 0; I <= J; I++) A[pk(I, J)] += x[I] * y[J];
}

template <int D> AD4SM_DI void ogden_tangent(const MatDev& m, const double* F, Tangent<D>& t) {
  constexpr int DD = D * D;
  const double alpha = m.p[0], mu = m.p[1];
  // C = FᵀF (F[k*D+j] = F_jk), eigen-decomposition C = V diag(x) Vᵀ
  double C[DD], V[DD];
#pragma unroll
  for (int a = 0; a < D; a++)
#pragma unroll
    for (int b = a; b < D; b++) {
      double c = 0.0;
#pragma unroll
      for (int i = 0; i < D; i++) c += F[a * D + i] * F[b * D + i];
      C[a * D + b] = C[b * D + a] = c;
    }
  jacobi_eig<D>(C, V);
  double lam[D], x[D], lam_a[D];
#pragma unroll
  for (int i = 0; i < D; i++) {
    x[i] = C[i * D + i];
    lam[i] = sqrt(x[i]);
    lam_a[i] = exp(0.5 * alpha * log(x[i]));  // λ^α = x^(α/2); x^((α−2)/2) = λ^α / x
  }
  // n_i = F N_i / λ_i
  double U[DD], ilam[D];
#pragma unroll
  for (int i = 0; i < D; i++) ilam[i] = 1.0 / lam[i];
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int r = 0; r < D; r++) {
      double a = 0.0;
#pragma unroll
      for (int k = 0; k < D; k++) a += F[k * D + r] * V[k * D + i];
      U[r * D + i] = a * ilam[i];
    }
  Dual2<D> dl[D];
#pragma unroll
  for (int i = 0; i < D; i++) dl[i] = d2_var<D>(lam[i], i);
  double s;
  const Dual2<D> f = ogden_energy<D>(m, dl, lam_a, s);
  t.psi = f.v;
  const double c0 = lam[0] * f.g[0] - mu * s * lam_a[0];  // λ_i f_i = μ s λ_i^α + c0 for every i
#pragma unroll
  for (int I = 0; I < DD; I++) t.P[I] = 0.0;
#pragma unroll
  for (int I = 0; I < Tangent<D>::NA; I++) t.A[I] = 0.0;
  // q_ij[I] = (n_i ⊗ N_j)[I] = n_i[j'] N_j[k] in the F[k*D+j'] addressing, I = k*D + j'
  auto qvec = [&](int i, int j, double* q) {
#pragma unroll
    for (int I = 0; I < DD; I++) q[I] = U[(I % D) * D + i] * V[(I / D) * D + j];
  };
  // diagonal part: A += Σ_i q_ii w_iᵀ with w_i = Σ_j f_ij q_jj  (f symmetric, so the sum is symmetric);  P = Σ f_i q_ii.
  // q_jj is recomputed from U, V inside the sums instead of being held (27 doubles fewer live next to A).
#pragma unroll
  for (int i = 0; i < D; i++) {
    double qi[DD], w[DD];
    qvec(i, i, qi);
#pragma unroll
    for (int I = 0; I < DD; I++) {
      double a = 0.0;
#pragma unroll
      for (int j = 0; j < D; j++) a += f.h[pk(i, j)] * (U[(I % D) * D + j] * V[(I / D) * D + j]);
      w[I] = a;
      t.P[I] += f.g[i] * qi[I];
    }
    rank1_packed<DD>(t.A, qi, w);
  }
  // pair part: A += q1 (a q1 + b q2)ᵀ + q2 (a q2 + b q1)ᵀ,  q1 = q_ij, q2 = q_ji
#pragma unroll
  for (int i = 0; i < D; i++)
#pragma unroll
    for (int j = i + 1; j < D; j++) {
      const double a = mu * s * divdiff_pow(x[i], x[j], lam_a[i], lam_a[j], 0.5 * alpha);
      const double b = mu * s * lam[i] * lam[j] * divdiff_pow(x[i], x[j], lam_a[i] / x[i], lam_a[j] / x[j], 0.5 * alpha - 1.0) -
                       c0 / (lam[i] * lam[j]);
      double q1[DD], q2[DD], y[DD];
      qvec(i, j, q1);
      qvec(j, i, q2);
#pragma unroll
      for (int I = 0; I < DD; I++) y[I] = a * q1[I] + b * q2[I];
      rank1_packed<DD>(t.A, q1, y);
#pragma unroll
      for (int I = 0; I < DD; I++) y[I] = a * q2[I] + b * q1[I];
      rank1_packed<DD>(t.A, q2, y);
    }
}

// ---------------------------------------------------------------------------------------------
// dispatcher: ψ, P, A for one Gauss point.  pf = nullptr for plain elastic materials.
// For PhaseField the d-dependent but F-independent term Gc/(2 l0) (dⁿ + l0²|∇d|²) is added to ψ.
// ---------------------------------------------------------------------------------------------
template <int D, int MF> AD4SM_DI void material_tangent(const MatDev& m, const double* F, const PfPoint* pf, Tangent<D>& t) {
  double g = 1.0;
  if (m.kind == MAT_PHASEFIELD && pf) g = (1.0 - pf->d) * (1.0 - pf->d);
  if (MF == MF_HOOKE) {
    double E[D * D], p, q, s;
    const bool small = m.flags & F_SMALL;
    strain<D>(F, small, E);
    hooke_family_coeffs<D>(m, E, g, p, q, s);
    quad_strain_tangent<D>(F, E, small, p, q, s, t);
  } else if (MF == MF_NEO) {
    if constexpr (D == 3) hyper_tangent3<false>(m, F, t);
    else hyper_tangent2<false>(m, F, t);
  } else if (MF == MF_MOONEY) {
    if constexpr (D == 3) hyper_tangent3<true>(m, F, t);
    else hyper_tangent2<true>(m, F, t);
  } else if (MF == MF_PFNEO) {
    pfneo_tangent<D>(m, F, g, t);
  } else if (MF == MF_OGDEN) {
    ogden_tangent<D>(m, F, t);
  }
  if (m.kind == MAT_PHASEFIELD && pf) {
    const double l0 = m.p[4], Gc = m.p[5], n = m.p[6];
    const double dn = (n == 2.0) ? pf->d * pf->d : ((n == 1.0) ? pf->d : pow(pf->d, n));
    t.psi += Gc / (2 * l0) * (dn + l0 * l0 * pf->gd2);
  }
}

// ---------------------------------------------------------------------------------------------
// d-phase: value-only split of the elastic energy  ψ = (1-d)² ψ⁺ + ψ⁻   (phasefieldmaterials.jl:20-135)
// with_hist selects the `λ/2` (instead of K/2) compressive term of the DHn history variant (:131).
// ---------------------------------------------------------------------------------------------
template <int D> AD4SM_DI void pf_energy_split(const MatDev& m, const double* F, bool with_hist, double& plus, double& minus,
                                              bool& tensile) {
  if (m.base == MAT_NEOHOOKE && !(m.flags & F_DHN)) {
    const double C1 = m.p[0], K = m.p[1];
    double I1 = (D == 2) ? 1.0 : 0.0;
#pragma unroll
    for (int I = 0; I < D * D; I++) I1 += F[I] * F[I];
    const double J = (D == 3) ? det3(F) : (F[0] * F[3] - F[1] * F[2]);
    const double e = C1 * (I1 - 3 - 2 * log(J)) + K * (J - 1) * (J - 1);
    tensile = J >= 1.0;
    plus = tensile ? e : 0.0;
    minus = tensile ? 0.0 : e;
    return;
  }
  double E[D * D];
  strain<D>(F, m.flags & F_SMALL, E);
  double lam, mu;
  lame(m.p[0], m.p[1], lam, mu);
  const double nu = m.p[1];
  const bool ps = (m.base == MAT_HOOKE2D) && (m.flags & F_PLANE_STRESS);
  double tr = 0.0, sq = 0.0;
#pragma unroll
  for (int i = 0; i < D; i++) {
    tr += E[i * D + i];
    sq += E[i * D + i] * E[i * D + i];
#pragma unroll
    for (int j = i + 1; j < D; j++) sq += 2.0 * E[j * D + i] * E[j * D + i];
  }
  const double E33 = ps ? nu / (nu - 1) * tr : 0.0;
  const double I1 = tr + E33;
  const double I1sq = sq + E33 * E33;
  tensile = I1 >= 0.0;
  if (!(m.flags & F_DHN)) {
    const double vol = lam / 2 * I1 * I1, dev = mu * I1sq;
    plus = tensile ? vol + dev : dev;
    minus = tensile ? 0.0 : vol;
  } else {
    const double K = lam + mu / 3;
    const double ed = I1sq - I1 * I1 / 3.0;
    const double dev = mu * ed;
    plus = tensile ? K / 2 * I1 * I1 + dev : dev;
    minus = tensile ? 0.0 : ((with_hist ? lam : K) / 2 * I1 * I1);
  }
}

}  // namespace ad4sm
