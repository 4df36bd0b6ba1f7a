// TEST INFRASTRUCTURE: compiles the product's device headers (csrc/*.cuh) for the host with g++ so the
// per-Gauss-point material tangents and the per-element column contraction can be checked against the
// oracle without a GPU.  It mirrors the sequence kernels.cu runs (phase 1 per Gauss point, phase 2 per
// column node); it is never used by the product.
#include <cstring>
#include <vector>

#include "element.cuh"

using namespace ad4sm;

static MatDev make_mat(const int* desc, const double* params) {
  MatDev m;
  m.kind = desc[0];
  m.base = desc[1];
  m.flags = (unsigned)desc[2];
  m.pad = 0;
  for (int i = 0; i < 8; i++) m.p[i] = params[i];
  return m;
}

template <int D> static int tangent_d(int mf, const MatDev& m, const double* F, const PfPoint* pf, double* psi, double* P,
                                      double* A) {
  Tangent<D> t;
  switch (mf) {
    case MF_HOOKE: material_tangent<D, MF_HOOKE>(m, F, pf, t); break;
    case MF_NEO: material_tangent<D, MF_NEO>(m, F, pf, t); break;
    case MF_MOONEY: material_tangent<D, MF_MOONEY>(m, F, pf, t); break;
    case MF_PFNEO: material_tangent<D, MF_PFNEO>(m, F, pf, t); break;
    case MF_OGDEN: material_tangent<D, MF_OGDEN>(m, F, pf, t); break;
    default: return -1;
  }
  *psi = t.psi;
  for (int i = 0; i < D * D; i++) P[i] = t.P[i];
  for (int i = 0; i < Tangent<D>::NA; i++) A[i] = t.A[i];
  return 0;
}

template <int D, int N, int MF, bool PF>
static void element_u_t(const MatDev& m, int P, const double* gN, const double* wgt, const double* Nv, const double* ue,
                        const double* de, double* phi, double* re, double* Ke) {
  using R = GpRec<D>;
  constexpr int RS = R::SIZE + 1;  // record capacity (written in pairs)
  std::vector<double> recs((size_t)P * RS), GS((size_t)P * D * N);
  for (int gp = 0; gp < P; gp++) {
    double* G = GS.data() + (size_t)gp * D * N;  // [gp][k][a], the layout the kernels keep in shared memory
    for (int k = 0; k < D; k++)
      for (int a = 0; a < N; a++) G[k * N + a] = gN[(k * P + gp) * N + a];
    gauss_point_u<D, N, MF, PF>(m, G, ue, de, PF ? Nv + gp * N : nullptr, wgt[gp], recs.data() + (size_t)gp * RS);
  }
  double s = 0.0;
  for (int gp = 0; gp < P; gp++) s += recs[(size_t)gp * RS + R::OFF_PSI];
  *phi = s;
  for (int q = 0; q < N; q++) column_u<D, N, 0, ScalarMem>(q, P, recs.data(), RS, gN, N, P * N, Ke, re);  // raw [k][gp][a] layout
}

template <int D, int N>
static int element_u_mf(int mf, bool pf, const MatDev& m, int P, const double* gN, const double* wgt, const double* Nv,
                        const double* ue, const double* de, double* phi, double* re, double* Ke) {
#define CASE(MFV)                                                                                   \
  case MFV:                                                                                         \
    if (pf) element_u_t<D, N, MFV, true>(m, P, gN, wgt, Nv, ue, de, phi, re, Ke);                   \
    else element_u_t<D, N, MFV, false>(m, P, gN, wgt, Nv, ue, de, phi, re, Ke);                     \
    return 0;
  switch (mf) {
    CASE(MF_HOOKE) CASE(MF_NEO) CASE(MF_MOONEY) CASE(MF_PFNEO) CASE(MF_OGDEN)
  }
#undef CASE
  return -1;
}

template <int D, int N>
static void element_d_t(const MatDev& m, int P, const double* gN, const double* wgt, const double* Nv, const double* ue,
                        const double* de, double* hist, double* phi, double* re, double* Ke) {
  constexpr int RS = GpRecD<D>::SIZE;
  std::vector<double> recs((size_t)P * RS), GS((size_t)P * D * N);
  for (int gp = 0; gp < P; gp++) {
    double* G = GS.data() + (size_t)gp * D * N;
    for (int k = 0; k < D; k++)
      for (int a = 0; a < N; a++) G[k * N + a] = gN[(k * P + gp) * N + a];
    gauss_point_d<D, N>(m, G, Nv + gp * N, ue, de, wgt[gp], hist ? hist + 2 * gp : nullptr, recs.data() + (size_t)gp * RS);
  }
  double s = 0.0;
  for (int gp = 0; gp < P; gp++) s += recs[(size_t)gp * RS + 2];
  *phi = s;
  for (int q = 0; q < N; q++) column_d<D, N>(q, P, recs.data(), RS, GS.data(), D * N, Nv, N, Ke, re);
}

extern "C" {

int h_tangent(int D, int mf, const int* desc, const double* params, const double* F, int has_pf, double d, double gd2,
              double* psi, double* P, double* A) {
  MatDev m = make_mat(desc, params);
  PfPoint pf{d, gd2};
  if (D == 2) return tangent_d<2>(mf, m, F, has_pf ? &pf : nullptr, psi, P, A);
  if (D == 3) return tangent_d<3>(mf, m, F, has_pf ? &pf : nullptr, psi, P, A);
  return -1;
}

void h_pair_slot(int N, int a, int b, int* idx, int* tr) { pair_slot(N, a, b, *idx, *tr); }

// one element, u-dual: Ke is the block storage [NPB][D*D], re [N][D]
int h_element_u(int D, int N, int mf, int pf, const int* desc, const double* params, int P, const double* gN,
                const double* wgt, const double* Nv, const double* ue, const double* de, double* phi, double* re,
                double* Ke) {
  MatDev m = make_mat(desc, params);
#define RUN(DD, NN) \
  if (D == DD && N == NN) return element_u_mf<DD, NN>(mf, pf != 0, m, P, gN, wgt, Nv, ue, de, phi, re, Ke);
  RUN(2, 3) RUN(2, 4) RUN(3, 4) RUN(3, 6) RUN(3, 8)
#undef RUN
  return -2;
}

int h_element_d(int D, int N, const int* desc, const double* params, int P, const double* gN, const double* wgt,
                const double* Nv, const double* ue, const double* de, double* hist, double* phi, double* re, double* Ke) {
  MatDev m = make_mat(desc, params);
#define RUN(DD, NN)                                                       \
  if (D == DD && N == NN) {                                               \
    element_d_t<DD, NN>(m, P, gN, wgt, Nv, ue, de, hist, phi, re, Ke);    \
    return 0;                                                             \
  }
  RUN(2, 3) RUN(2, 4) RUN(3, 4) RUN(3, 6) RUN(3, 8)
#undef RUN
  return -2;
}
}
