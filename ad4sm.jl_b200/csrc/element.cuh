// element.cuh -- per-element building blocks of K1 (element energy / gradient / Hessian).
//
// Replaces, for one element, the reference's Gauss loops
//   getϕ(elem::C3DE, u::Array{<:D2})      /root/reference/src/elasticelements.jl:339-350
//   _integrate_ϕ                           /root/reference/src/elasticelements.jl:199-206
//   getϕ(elem::CPElem, u, d[, ϕmax])       /root/reference/src/phasefieldelements.jl:159-223
//   getF / get_d_and_∇d / chain operator × /root/reference/src/elements.toolkit.jl:8-20,159-167,231-245
// The reference differentiates with D2{3N,·} duals; here each Gauss point yields the material
// tangent (ψ, P = ∂ψ/∂F, A = ∂²ψ/∂F∂F, materials.cuh) and the chain to the element dofs is the
// exact two-stage contraction  K_(a,j),(b,l) = Σ_gp w Σ_k,m ∇N_a,k A_(kj)(ml) ∇N_b,m  (F is linear
// in u, so this is what `×` computes, toolkit:231-245).
//
// Work split (kernels.cu): a Gauss point is evaluated by one thread ("phase 1", gauss_point*),
// which leaves w·A, w·P, w·ψ in a GP record; then thread q of the element owns column node b = q
// and accumulates over the Gauss points the blocks K(a,b), a = (q+o) mod N, o = 0..N/2
// ("phase 2", column_*).  No cross-thread floating-point reduction is needed.
//
// Everything here is __host__ __device__ so tests/cxx can run the same code on the CPU.
#pragma once
#include "materials.cuh"

namespace ad4sm {

// ---- element block storage ("cyclic" symmetric layout) ------------------------------------------
// An element tangent has N x N node blocks of DB x DB; only NPB = N(N+1)/2 are stored.  Block
// (q, o) holds  M[j][l] = K_((q+o)%N, j),(q, l)  at index o*N + q;  o <= N/2, and for even N the
// o = N/2 blocks only for q < N/2.
AD4SM_HD constexpr int n_pair_blocks(int N) { return N * (N + 1) / 2; }
AD4SM_HD constexpr int n_offsets(int N, int q) {  // number of o's owned by column q
  return (N % 2) ? (N + 1) / 2 : (q < N / 2 ? N / 2 + 1 : N / 2);
}
// where K(a,b) (row node a, col node b) lives: index and whether to read it transposed
AD4SM_HD inline void pair_slot(int N, int a, int b, int& idx, int& tr) {
  const int o1 = ((a - b) % N + N) % N;
  if (2 * o1 < N || (2 * o1 == N && b < N / 2)) {
    idx = o1 * N + b;
    tr = 0;
  } else {
    idx = (((b - a) % N + N) % N) * N + a;
    tr = 1;
  }
}

// ---- GP record -------------------------------------------------------------------------------------
template <int D> struct GpRec {
  static constexpr int DD = D * D;
  static constexpr int NA = DD * (DD + 1) / 2;
  static constexpr int OFF_A = 0, OFF_P = NA, OFF_PSI = NA + DD;
  static constexpr int SIZE = NA + DD + 1;
};
// d-phase record: [0] w ψ'', [1] w ψ', [2] w ψ, [3] w c (c = Gc l0, "intended" gradient term, else 0), [4..4+D) ∇d
template <int D> struct GpRecD {
  static constexpr int SIZE = 4 + D;
};

// 16-byte shared-memory access helpers.  A GP record / gradient tile that lives in shared memory is read
// with LDS.128 (a quarter-warp = the 8 lanes of one element read the same address: one broadcast
// wavefront); a record held in registers (thread-per-element path) or on the host is read as scalars.
// compiler-only fence: bounds how far shared-memory loads are hoisted (register pressure of phase 2)
#ifdef __CUDA_ARCH__
#define AD4SM_SCHED_FENCE() asm volatile("" ::: "memory")
#else
#define AD4SM_SCHED_FENCE() ((void)0)
#endif
#ifdef __CUDACC__
struct VecMem {
  const double* p;  // 16-byte aligned
  AD4SM_DI double operator()(int i) const {
    const double2 v = reinterpret_cast<const double2*>(p)[i >> 1];
    return (i & 1) ? v.y : v.x;
  }
};
AD4SM_DI void store2(double* p, int i2, double a, double b) { reinterpret_cast<double2*>(p)[i2] = make_double2(a, b); }
#else
AD4SM_DI void store2(double* p, int i2, double a, double b) {
  p[2 * i2] = a;
  p[2 * i2 + 1] = b;
}
#endif
struct ScalarMem {
  const double* p;
  AD4SM_DI double operator()(int i) const { return p[i]; }
};

// F[k*D+j] = δ_jk + Σ_a ∇N_a,k u_a,j   (toolkit:8-20); G[k*N+a], ue[a*D+j]
template <int D, int N> AD4SM_DI void deformation_gradient(const double* G, const double* ue, double* F) {
#pragma unroll
  for (int k = 0; k < D; k++)
#pragma unroll
    for (int j = 0; j < D; j++) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < N; a++) s += G[k * N + a] * ue[a * D + j];
      F[k * D + j] = (j == k) ? s + 1.0 : s;
    }
}

// phase 1, u-dual paths (elastic and phase-field u-phase): writes the GP record.
//   G [D][N] physical shape-function gradients at this Gauss point, ue [N][D] nodal displacements,
//   de [N] nodal phase field and Nv [N] shape-function values (PF only), w = elem.wgt[gp].
//   psi_out: nullptr -> w·ψ is stored in the record (slot OFF_PSI); else it is written there and the record holds
//            only w·A and w·P (NA + DD doubles, an even count for D = 2, 3).
template <int D, int N, int MF, bool PF>
AD4SM_DI void gauss_point_u_eval(const MatDev& m, const double* G, const double* ue, const double* de, const double* Nv,
                                 Tangent<D>& t) {
  double F[D * D];
  deformation_gradient<D, N>(G, ue, F);
  PfPoint pf;
  if (PF) {  // get_d_and_∇d, toolkit:159-167
    double dv = 0.0;
#pragma unroll
    for (int a = 0; a < N; a++) dv += Nv[a] * de[a];
    double g2 = 0.0;
#pragma unroll
    for (int k = 0; k < D; k++) {
      double s = 0.0;
#pragma unroll
      for (int a = 0; a < N; a++) s += G[k * N + a] * de[a];
      g2 += s * s;
    }
    pf.d = dv;
    pf.gd2 = g2;
  }
  material_tangent<D, MF>(m, F, PF ? &pf : nullptr, t);
}

// record = [w·A (NA) | w·P (DD) | w·ψ], written in 16-byte pairs (rec is 16-byte aligned, capacity SIZE+1)
template <int D> AD4SM_DI void gp_record_store(const Tangent<D>& t, double w, double* rec, double* psi_out) {
  using R = GpRec<D>;
  auto val = [&](int i) -> double {
    return i < R::NA ? w * t.A[i] : (i < R::NA + R::DD ? w * t.P[i - R::NA] : (i == R::OFF_PSI ? w * t.psi : 0.0));
  };
  if (psi_out) {
    static_assert((R::NA + R::DD) % 2 == 0, "record body must be a whole number of 16-byte pairs");
    *psi_out = w * t.psi;
#pragma unroll
    for (int i2 = 0; i2 < (R::NA + R::DD) / 2; i2++) store2(rec, i2, val(2 * i2), val(2 * i2 + 1));
  } else {
#pragma unroll
    for (int i2 = 0; i2 < (R::SIZE + 1) / 2; i2++) store2(rec, i2, val(2 * i2), val(2 * i2 + 1));
  }
}

template <int D, int N, int MF, bool PF>
AD4SM_DI void gauss_point_u(const MatDev& m, const double* G, const double* ue, const double* de, const double* Nv,
                            double w, double* rec, double* psi_out = nullptr) {
  Tangent<D> t;
  gauss_point_u_eval<D, N, MF, PF>(m, G, ue, de, Nv, t);
  gp_record_store<D>(t, w, rec, psi_out);
}

// phase 1, d-dual path (makeϕrKt_d, phasefieldelements.jl:239-262; energies phasefieldmaterials.jl:20-135).
// hist: nullptr or this Gauss point's (ϕmax⁺, ϕmax⁻) pair, updated in place (phasefieldelements.jl:180).
template <int D, int N>
AD4SM_DI void gauss_point_d(const MatDev& m, const double* G, const double* Nv, const double* ue, const double* de,
                            double w, double* hist, double* rec) {
  double F[D * D];
  deformation_gradient<D, N>(G, ue, F);
  double plus, minus;
  bool tensile;
  pf_energy_split<D>(m, F, hist != nullptr, plus, minus, tensile);
  if (hist) {
    const bool neo = (m.base == MAT_NEOHOOKE) && !(m.flags & F_DHN);
    if (tensile) {
      const double h = plus > hist[0] ? plus : hist[0];
      hist[0] = h;
      plus = h;
    } else if (!neo) {
      const double h = plus > hist[1] ? plus : hist[1];
      hist[1] = h;
      plus = h;
    }
  }
  double dv = 0.0;
#pragma unroll
  for (int a = 0; a < N; a++) dv += Nv[a] * de[a];
  double gd[D], g2 = 0.0;
#pragma unroll
  for (int k = 0; k < D; k++) {
    double s = 0.0;
#pragma unroll
    for (int a = 0; a < N; a++) s += G[k * N + a] * de[a];
    gd[k] = s;
    g2 += s * s;
  }
  const double l0 = m.p[4], Gc = m.p[5], n = m.p[6];
  const double c0 = Gc / (2 * l0);
  // dⁿ and its derivatives with the reference's `^` rule (adiff.jl:138-139)
  double pn, pn1, pn2;
  if (n == 0.0) {
    pn = 1.0, pn1 = 0.0, pn2 = 0.0;
  } else if (n == 1.0) {
    pn = dv, pn1 = 1.0, pn2 = 0.0;
  } else if (n == 2.0) {
    pn = dv * dv, pn1 = 2.0 * dv, pn2 = 2.0;
  } else {
    pn = pow(dv, n), pn1 = n * pow(dv, n - 1.0), pn2 = n * (n - 1.0) * pow(dv, n - 2.0);
  }
  const double omd = 1.0 - dv;
  const double psi = omd * omd * plus + minus + c0 * (pn + l0 * l0 * g2);
  const double s1 = -2.0 * omd * plus + c0 * pn1;
  const double s2 = 2.0 * plus + c0 * pn2;
  rec[0] = w * s2;
  rec[1] = w * s1;
  rec[2] = w * psi;
  rec[3] = (m.flags & F_PF_GRAD_INTENDED) ? w * (2.0 * c0 * l0 * l0) : 0.0;
#pragma unroll
  for (int k = 0; k < D; k++) rec[4 + k] = gd[k];
}

// ---- CAS: axisymmetric continuum element (EXTENSION: the reference has the constructors ASTria / ASQuad,
// elasticelements.jl:235-284, but neither the struct nor its getϕ -- SURVEY §9-7) ---------------------------------------
// Energy  Φ_e = Σ_gp wgt ϕ(F),  F = [1+u_r,r  u_r,z  0;  u_z,r  1+u_z,z  0;  0  0  1 + u_r/r]  with the 3-D material ϕ;
// dofs (u_r, u_z) per node; wgt carries 2π r (elasticelements.jl:249, 279).  F is linear in u, so -- as for the other
// elements -- gradient and Hessian are the exact chain of the material tangent (P, A) through  ∂F/∂u:
//   ∂F_(k,j)/∂u_(a,j) = ∇N_a,k (k, j in {r, z}),   ∂F_33/∂u_(a,r) = N_a / r =: H_a   (the "hoop row", K1Args.Nval).
// One thread per element; Kf = full (2N)x(2N) tangent [row dof][col dof], dof = 2*node + comp.
template <int N, int MF>
AD4SM_DI void cas_element(const MatDev& m, int P, const double* gN, const double* wg, const double* Hr, const double* ue,
                          bool want_K, double* Kf, double* r, double& phi) {
  constexpr int ND = 2 * N;
  constexpr int I5[5] = {0, 1, 3, 4, 8};  // the F components that depend on u, as linear indices k*3 + j of the 3x3 F
  phi = 0.0;
#pragma unroll
  for (int i = 0; i < ND; i++) r[i] = 0.0;
  if (want_K) {
#pragma unroll
    for (int i = 0; i < ND * ND; i++) Kf[i] = 0.0;
  }
  for (int gp = 0; gp < P; gp++) {
    double G[2 * N], H[N];
#pragma unroll
    for (int k = 0; k < 2; k++)
#pragma unroll
      for (int a = 0; a < N; a++) G[k * N + a] = gN[(k * P + gp) * N + a];
#pragma unroll
    for (int a = 0; a < N; a++) H[a] = Hr[gp * N + a];
    double F2[4];
    deformation_gradient<2, N>(G, ue, F2);
    double hoop = 0.0;
#pragma unroll
    for (int a = 0; a < N; a++) hoop += H[a] * ue[a * 2];
    double F[9] = {F2[0], F2[1], 0.0, F2[2], F2[3], 0.0, 0.0, 0.0, hoop + 1.0};
    Tangent<3> t;
    material_tangent<3, MF>(m, F, nullptr, t);
    const double w = wg[gp];
    phi += w * t.psi;
    // B_(a,j)[I]: the derivative of F component I5[I] with respect to dof (a, j)
    auto Bv = [&](int a, int j, int I) -> double {
      return I < 4 ? ((I & 1) == j ? G[(I >> 1) * N + a] : 0.0) : (j == 0 ? H[a] : 0.0);
    };
#pragma unroll
    for (int a = 0; a < N; a++)
#pragma unroll
      for (int j = 0; j < 2; j++) {
        double s = 0.0;
#pragma unroll
        for (int I = 0; I < 5; I++) s += Bv(a, j, I) * t.P[I5[I]];
        r[a * 2 + j] += w * s;
      }
    if (want_K) {
#pragma unroll
      for (int b = 0; b < N; b++)
#pragma unroll
        for (int l = 0; l < 2; l++) {
          double T[5];  // T[I] = Σ_J A_(I,J) B_(b,l)[J]
#pragma unroll
          for (int I = 0; I < 5; I++) {
            double s = 0.0;
#pragma unroll
            for (int J = 0; J < 5; J++) s += t.A[pk(I5[I], I5[J])] * Bv(b, l, J);
            T[I] = w * s;
          }
#pragma unroll
          for (int a = 0; a < N; a++)
#pragma unroll
            for (int j = 0; j < 2; j++) {
              double s = Kf[(a * 2 + j) * ND + b * 2 + l];
#pragma unroll
              for (int I = 0; I < 5; I++) s += Bv(a, j, I) * T[I];
              Kf[(a * 2 + j) * ND + b * 2 + l] = s;
            }
        }
    }
  }
}
// the full element tangent into the cyclic block storage Ke [NPB][2*2]; diagonal node blocks mirrored like column_u_acc
template <int N> AD4SM_DI void cas_store(const double* Kf, double* Ke) {
  constexpr int ND = 2 * N;
#pragma unroll
  for (int q = 0; q < N; q++)
#pragma unroll
    for (int o = 0; o <= N / 2; o++)
      if (o < n_offsets(N, q)) {
        const int a = (q + o) % N;
#pragma unroll
        for (int j = 0; j < 2; j++)
#pragma unroll
          for (int l = 0; l < 2; l++) {
            const int jj = (o == 0 && j > l) ? l : j, ll = (o == 0 && j > l) ? j : l;  // lower entry := upper entry
            Ke[(o * N + q) * 4 + j * 2 + l] = Kf[(a * 2 + jj) * ND + q * 2 + ll];
          }
      }
}

// T[(k,j)][l] = Σ_m A_(kj)(ml) ∇N_b,m  from the packed symmetric A (81 FMAs for D = 3)
template <int D, class Mem> AD4SM_DI void tangent_times_gradient(const Mem& A, const double* gb, double (*T)[D]) {
  constexpr int DD = D * D;
#pragma unroll
  for (int I = 0; I < DD; I++)
#pragma unroll
    for (int l = 0; l < D; l++) T[I][l] = 0.0;
#pragma unroll
  for (int J = 0; J < DD; J++)
#pragma unroll
    for (int I = 0; I <= J; I++) {
      const double a = A(J * (J + 1) / 2 + I);
      T[I][J % D] += a * gb[J / D];
      if (I != J) T[J][I % D] += a * gb[I / D];
    }
}

// blk[j][l] += Σ_k ∇N_a,k T[(k,j)][l]
template <int D> AD4SM_DI void add_block(const double* ga, const double (*T)[D], double* blk) {
#pragma unroll
  for (int j = 0; j < D; j++)
#pragma unroll
    for (int l = 0; l < D; l++) {
      double s = blk[j * D + l];
#pragma unroll
      for (int k = 0; k < D; k++) s += ga[k] * T[k * D + j][l];
      blk[j * D + l] = s;
    }
}

// phase 2 (u-dual): column node q of one element.
//   recs: P GP records, `rstride` doubles apart, read through Mem (VecMem: shared memory, ScalarMem: registers/host)
//   gS:   the element's ∇N: ∇N_a,k at Gauss point gp is gS[gp*gstride + k*kstride + a]
//   Ke:   the element's block storage [NPB][D*D] (may be nullptr when only ϕ, r are wanted)
//   re:   the element's residual [N][D].
// PT: compile-time number of Gauss points (0 = use the run-time P), so that a register-resident
// record (P = 1, thread-per-element path) is indexed statically.
template <int D, int N, int PT, class Mem>
AD4SM_DI void column_u_acc(int q, int Prt, const double* recs, int rstride, const double* gS, int gstride, int kstride,
                           bool want_K, double (*acc)[D * D], double* rq) {
  using R = GpRec<D>;
  const int P = PT ? PT : Prt;
  constexpr int DD = D * D;
  constexpr int NO = N / 2 + 1;  // max offsets per column
#pragma unroll
  for (int o = 0; o < NO; o++)
#pragma unroll
    for (int i = 0; i < DD; i++) acc[o][i] = 0.0;
#pragma unroll
  for (int j = 0; j < D; j++) rq[j] = 0.0;
  // row-node pointers (a = (q + o) mod N), hoisted: inside the loop only gp * gstride + k * kstride is added
  const double* ga[NO];
#pragma unroll
  for (int o = 0; o < NO; o++) {
    int a = q + o;
    if (a >= N) a -= N;
    ga[o] = gS + a;
  }
#if defined(AD4SM_VAR_UNROLL2)
#pragma unroll 2
#endif
  for (int gp = 0; gp < P; gp++) {
    const Mem rec{recs + gp * rstride};
    const int goff = gp * gstride;
    double gb[D];
#pragma unroll
    for (int k = 0; k < D; k++) gb[k] = ga[0][goff + k * kstride];
    // r_(q,j) += Σ_k wP_(kj) ∇N_q,k
#pragma unroll
    for (int j = 0; j < D; j++)
#pragma unroll
      for (int k = 0; k < D; k++) rq[j] += rec(R::OFF_P + k * D + j) * gb[k];
    if (want_K) {
      // one block column l at a time, so that only D*D intermediate values are live next to the accumulators:
      //   T_l[(k,j)] = Σ_m A_(kj)(ml) ∇N_q,m ;   K(a,q)[j][l] += Σ_k ∇N_a,k T_l[(k,j)]
#pragma unroll
      for (int l = 0; l < D; l++) {
#if !defined(AD4SM_VAR_NOFENCE)
        AD4SM_SCHED_FENCE();  // keep the record loads of one column from being hoisted over the previous one
#endif
        double Tl[DD];
#pragma unroll
        for (int I = 0; I < DD; I++) {
          double t = 0.0;
#pragma unroll
          for (int m = 0; m < D; m++) t += rec(pk(I, m * D + l)) * gb[m];
          Tl[I] = t;
        }
        // every lane accumulates all NO offsets: for even N the o = N/2 block of the columns q >= N/2 is a
        // duplicate that is never stored, but skipping it would only add a divergent branch (the warp issues the
        // instructions for the other lanes anyway)
#pragma unroll
        for (int o = 0; o < NO; o++) {
#pragma unroll
          for (int k = 0; k < D; k++) {
            const double gak = ga[o][goff + k * kstride];
#pragma unroll
            for (int j = 0; j < D; j++) acc[o][j * D + l] += gak * Tl[k * D + j];
          }
        }
      }
    }
  }
  // the diagonal node block is symmetric in exact arithmetic; mirror its upper triangle so that Kt is
  // bitwise symmetric like the reference's (both triangles read one packed slot, adiff.jl:160)
#pragma unroll
  for (int j = 0; j < D; j++)
#pragma unroll
    for (int l = j + 1; l < D; l++) acc[0][l * D + j] = acc[0][j * D + l];
}

// write column q's blocks into the element's block storage Ke [NPB][D*D] (global or shared memory)
template <int D, int N> AD4SM_DI void column_u_store(int q, const double (*acc)[D * D], double* Ke) {
  constexpr int DD = D * D;
  constexpr int NO = N / 2 + 1;
  const int no = n_offsets(N, q);
#pragma unroll
  for (int o = 0; o < NO; o++)
    if (o < no) {
#pragma unroll
      for (int i = 0; i < DD; i++) Ke[(o * N + q) * DD + i] = acc[o][i];
    }
}

template <int D, int N, int PT, class Mem>
AD4SM_DI void column_u(int q, int Prt, const double* recs, int rstride, const double* gS, int gstride, int kstride,
                       double* Ke, double* re) {
  double acc[N / 2 + 1][D * D], rq[D];
  column_u_acc<D, N, PT, Mem>(q, Prt, recs, rstride, gS, gstride, kstride, Ke != nullptr, acc, rq);
#pragma unroll
  for (int j = 0; j < D; j++) re[q * D + j] = rq[j];
  if (Ke) column_u_store<D, N>(q, acc, Ke);
}

// phase 2 (d-dual): column node q; blocks are scalars.  gS as in column_u; Nv: [gp][a] shape-function values,
// `nstride` doubles between Gauss points.
template <int D, int N, int PT = 0>
AD4SM_DI void column_d(int q, int Prt, const double* recs, int rstride, const double* gS, int gstride, const double* Nv,
                       int nstride, double* Ke, double* re) {
  constexpr int NO = N / 2 + 1;
  const int P = PT ? PT : Prt;
  const int no = n_offsets(N, q);
  double acc[NO];
#pragma unroll
  for (int o = 0; o < NO; o++) acc[o] = 0.0;
  double rq = 0.0;
  for (int gp = 0; gp < P; gp++) {
    const double* rec = recs + gp * rstride;
    const double* g = gS + gp * gstride;
    const double* nv = Nv + gp * nstride;
    const double ws2 = rec[0], ws1 = rec[1], wc = rec[3];
    const double nb = nv[q];
    double gb[D], gdgb = 0.0;
#pragma unroll
    for (int k = 0; k < D; k++) {
      gb[k] = g[k * N + q];
      gdgb += rec[4 + k] * gb[k];
    }
    rq += ws1 * nb + wc * gdgb;
    if (Ke) {
#pragma unroll
      for (int o = 0; o < NO; o++)
        if (o < no) {
          int a = q + o;
          if (a >= N) a -= N;
          double gg = 0.0;
#pragma unroll
          for (int k = 0; k < D; k++) gg += g[k * N + a] * gb[k];
          acc[o] += ws2 * (nv[a] * nb) + wc * gg;
        }
    }
  }
  re[q] = rq;
  if (Ke) {
#pragma unroll
    for (int o = 0; o < NO; o++)
      if (o < no) Ke[o * N + q] = acc[o];
  }
}

}  // namespace ad4sm
