// ORACLE -- test infrastructure only; never linked into or called by the product path.
//
// C++ restatement of the *reference algorithm* of AD4SM.jl v0.0.7 `makeϕrKt`:
//   - forward second-order duals                      /root/reference/src/adiff.jl:29-42, 88-94, 131-150
//   - kinematics getF / get_d_and_∇d                   /root/reference/src/elements.toolkit.jl:8-20, 159-167
//   - chain operator ×                                 /root/reference/src/elements.toolkit.jl:231-245
//   - element energies (generic + 3-D chain path)      /root/reference/src/elasticelements.jl:199-206, 339-350
//                                                      /root/reference/src/phasefieldelements.jl:159-223
//   - material energies                                /root/reference/src/elasticmaterials.jl:138-245
//                                                      /root/reference/src/materials.jl:98-155
//                                                      /root/reference/src/phasefieldmaterials.jl:20-135
//   - threaded element loop + SERIAL COO->CSC assembly /root/reference/src/elasticelements.jl:208-221
//                                                      /root/reference/src/elements.toolkit.jl:169-203
// It evaluates the same operations in the same association order as the reference (built with
// -ffp-contract=off), including the promotion of reals to full duals (adiff.jl:54,78-80), so it
// doubles as the timed "reference-algorithm restatement (C++)" CPU baseline.
//
// PARITY UNPINNED BY REFERENCE TESTS: the reference ships no tests/golden vectors and Julia is not
// installed here; this file is pinned by the Python restatement (oracle/*.py), analytic
// known-answer tests, sympy Hessians and mpmath finite differences (tests/test_oracle_*.py).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

typedef int64_t i64;

// ------------------------------------------------------------------------------------------------
// duals
// ------------------------------------------------------------------------------------------------
template <int N> struct D1 {
  double v;
  double g[N];
};
template <int N> struct D2 {
  static constexpr int n = N;
  static constexpr int m = N * (N + 1) / 2;
  double v;
  double g[N];
  double h[N * (N + 1) / 2];
};

template <int N> static inline D2<N> lift(double x) {  // adiff.jl:54
  D2<N> r;
  r.v = x;
  for (int i = 0; i < N; i++) r.g[i] = 0.0;
  for (int i = 0; i < D2<N>::m; i++) r.h[i] = 0.0;
  return r;
}
template <int N> static inline D1<N> lift1(double x) {  // adiff.jl:55
  D1<N> r;
  r.v = x;
  for (int i = 0; i < N; i++) r.g[i] = 0.0;
  return r;
}

// D1 ops (adiff.jl:111-114)
template <int N> static inline D1<N> operator+(const D1<N>& x, const D1<N>& y) {
  D1<N> r;
  r.v = x.v + y.v;
  for (int i = 0; i < N; i++) r.g[i] = x.g[i] + y.g[i];
  return r;
}
template <int N> static inline D1<N> operator*(const D1<N>& x, const D1<N>& y) {
  D1<N> r;
  r.v = x.v * y.v;
  for (int i = 0; i < N; i++) r.g[i] = x.v * y.g[i] + y.v * x.g[i];
  return r;
}
template <int N> static inline D1<N> operator*(double a, const D1<N>& y) { return lift1<N>(a) * y; }
template <int N> static inline D1<N> operator+(const D1<N>& x, double b) { return x + lift1<N>(b); }

// D2 ops (adiff.jl:132-140)
template <int N> static inline D2<N> operator+(const D2<N>& x, const D2<N>& y) {
  D2<N> r;
  r.v = x.v + y.v;
  for (int i = 0; i < N; i++) r.g[i] = x.g[i] + y.g[i];
  for (int i = 0; i < D2<N>::m; i++) r.h[i] = x.h[i] + y.h[i];
  return r;
}
template <int N> static inline D2<N> operator-(const D2<N>& x, const D2<N>& y) {
  D2<N> r;
  r.v = x.v - y.v;
  for (int i = 0; i < N; i++) r.g[i] = x.g[i] - y.g[i];
  for (int i = 0; i < D2<N>::m; i++) r.h[i] = x.h[i] - y.h[i];
  return r;
}
template <int N> static inline D2<N> operator-(const D2<N>& x) {
  D2<N> r;
  r.v = -x.v;
  for (int i = 0; i < N; i++) r.g[i] = -x.g[i];
  for (int i = 0; i < D2<N>::m; i++) r.h[i] = -x.h[i];
  return r;
}
template <int N> static inline D2<N> operator*(const D2<N>& x, const D2<N>& y) {  // adiff.jl:135
  D2<N> r;
  r.v = x.v * y.v;
  for (int i = 0; i < N; i++) r.g[i] = x.v * y.g[i] + y.v * x.g[i];
  int k = 0;
  for (int j = 0; j < N; j++)
    for (int i = 0; i <= j; i++, k++)
      r.h[k] = ((x.v * y.h[k] + y.v * x.h[k]) + x.g[i] * y.g[j]) + y.g[i] * x.g[j];
  return r;
}
template <int N> static inline D2<N> inv(const D2<N>& x) {  // adiff.jl:136
  D2<N> r;
  r.v = 1 / x.v;
  const double a = -1 / (x.v * x.v);
  const double b = 2 / (x.v * x.v * x.v);
  const double c = 1 / (x.v * x.v);
  for (int i = 0; i < N; i++) r.g[i] = a * x.g[i];
  int k = 0;
  for (int j = 0; j < N; j++)
    for (int i = 0; i <= j; i++, k++) r.h[k] = b * (x.g[i] * x.g[j]) - c * x.h[k];
  return r;
}
template <int N> static inline D2<N> operator/(const D2<N>& x, const D2<N>& y) { return x * inv(y); }  // :137
static inline double powd(double v, double n) { return std::pow(v, n); }
template <int N> static inline D2<N> dpow(const D2<N>& x, double n) {  // adiff.jl:138-139
  if (n == 0) {
    D2<N> one = lift<N>(1.0);
    return one;
  }
  if (n == 1) return x;
  D2<N> r;
  r.v = powd(x.v, n);
  const double a = n * powd(x.v, n - 1);
  const double b = n * (n - 1) * powd(x.v, n - 2);
  for (int i = 0; i < N; i++) r.g[i] = a * x.g[i];
  int k = 0;
  for (int j = 0; j < N; j++)
    for (int i = 0; i <= j; i++, k++) r.h[k] = b * (x.g[i] * x.g[j]) + a * x.h[k];
  return r;
}
template <int N> static inline D2<N> dlog(const D2<N>& x) {  // adiff.jl:140
  D2<N> r;
  r.v = std::log(x.v);
  const double v2 = x.v * x.v;
  for (int i = 0; i < N; i++) r.g[i] = x.g[i] / x.v;
  int k = 0;
  for (int j = 0; j < N; j++)
    for (int i = 0; i <= j; i++, k++) r.h[k] = -(x.g[i] * x.g[j]) / v2 + x.h[k] / x.v;
  return r;
}
template <int N> static inline D2<N> dsqrt(const D2<N>& x) { return dpow(x, 0.5); }  // :147
// mixed real/dual: promotion to a full dual, then the dual op (adiff.jl:78-80)
template <int N> static inline D2<N> operator*(double a, const D2<N>& y) { return lift<N>(a) * y; }
template <int N> static inline D2<N> operator*(const D2<N>& x, double b) { return x * lift<N>(b); }
template <int N> static inline D2<N> operator+(const D2<N>& x, double b) { return x + lift<N>(b); }
template <int N> static inline D2<N> operator+(double a, const D2<N>& y) { return lift<N>(a) + y; }
template <int N> static inline D2<N> operator-(const D2<N>& x, double b) { return x - lift<N>(b); }
template <int N> static inline D2<N> operator-(double a, const D2<N>& y) { return lift<N>(a) - y; }
template <int N> static inline D2<N> operator/(const D2<N>& x, double b) { return x * inv(lift<N>(b)); }
template <int N> static inline D2<N> operator/(double a, const D2<N>& y) { return lift<N>(a) * inv(y); }

// plain-double versions so the material templates also run on Float64
static inline double dpow(double x, double n) { return n == 0 ? 1.0 : (n == 1 ? x : std::pow(x, n)); }
static inline double dlog(double x) { return std::log(x); }
static inline double dsqrt(double x) { return std::sqrt(x); }
static inline double val(double x) { return x; }
template <int N> static inline double val(const D2<N>& x) { return x.v; }
template <int N> static inline double val(const D1<N>& x) { return x.v; }
template <class T> struct Zero {
  static T get() { return T(0); }
};
template <int N> struct Zero<D2<N>> {
  static D2<N> get() { return lift<N>(0.0); }
};
template <int N> struct Zero<D1<N>> {
  static D1<N> get() { return lift1<N>(0.0); }
};

// conversion Real -> dual with zero derivatives (adiff.jl:54), identity otherwise
template <class R> struct Conv {
  static R from(const R& x) { return x; }
  static R from(double x) { return Zero<R>::get() + x; }
};
template <> struct Conv<double> {
  static double from(double x) { return x; }
};

// ------------------------------------------------------------------------------------------------
// materials
// ------------------------------------------------------------------------------------------------
enum MatKind { HOOKE = 1, HOOKE2D = 2, NEOHOOKE = 3, MOONEY = 4, OGDEN = 5, PHASEFIELD = 6 };
enum MatFlags { F_SMALL = 1, F_PLANE_STRESS = 2, F_DHN = 4, F_PF_GRAD_INTENDED = 8, F_NH2D_EXT = 16 };
struct Mat {
  int kind;    // MatKind; for PHASEFIELD the base kind is in `base`
  int base;    // base MatKind of a PhaseField
  int flags;
  double p[8];  // Hooke/Hooke2D: E, nu | NeoHooke: C1, K | MooneyRivlin: C1, C2, K | Ogden: alpha, mu, K
                // PhaseField: p[0..3] base params, p[4]=l0, p[5]=Gc, p[6]=n
};

// F is addressed with the reference's linear index: F(k), k = 1..D*D, column-major.
template <class T, int D> struct Mx {
  T a[D * D];
  T& operator()(int k) { return a[k - 1]; }
  const T& operator()(int k) const { return a[k - 1]; }
  T& at(int i, int j) { return a[j * D + i]; }  // 0-based (row i, col j)
  const T& at(int i, int j) const { return a[j * D + i]; }
};

template <class T, int D> static inline Mx<T, D> FtF(const Mx<T, D>& F) {  // transpose(F)F, k left to right
  Mx<T, D> C;
  for (int i = 0; i < D; i++)
    for (int j = 0; j < D; j++) {
      T s = F.at(0, i) * F.at(0, j);
      for (int k = 1; k < D; k++) s = s + F.at(k, i) * F.at(k, j);
      C.at(i, j) = s;
    }
  return C;
}
template <class T, int D> static inline Mx<T, D> strain(const Mx<T, D>& F, bool small) {
  Mx<T, D> E;
  if (small) {  // (F+transpose(F)-2I)/2
    for (int i = 0; i < D; i++)
      for (int j = 0; j < D; j++) E.at(i, j) = (F.at(i, j) + F.at(j, i) - (i == j ? 2.0 : 0.0)) / 2.0;
  } else {  // (transpose(F)F-I)/2
    Mx<T, D> C = FtF<T, D>(F);
    for (int i = 0; i < D; i++)
      for (int j = 0; j < D; j++) E.at(i, j) = (C.at(i, j) - (i == j ? 1.0 : 0.0)) / 2.0;
  }
  return E;
}
static inline void lame(double Es, double nu, double& lam, double& mu) {  // elasticmaterials.jl:211-212
  lam = Es * nu / (1 + nu) / (1 - 2 * nu);
  mu = Es / 2 / (1 + nu);
}
template <class T> static inline T sq(const T& x) { return dpow(x, 2.0); }

template <class T> static T phi_hooke(const Mx<T, 3>& F, const double* p, int flags) {  // :202-219
  Mx<T, 3> E = strain<T, 3>(F, flags & F_SMALL);
  double lam, mu;
  lame(p[0], p[1], lam, mu);
  T phi = (mu + lam / 2) * (sq(E(1)) + sq(E(5)) + sq(E(9)));
  phi = phi + lam * (E(1) * E(5) + E(5) * E(9) + E(9) * E(1));
  phi = phi + (2 * mu) * (sq(E(2)) + sq(E(3)) + sq(E(6)));
  return phi;
}
template <class T> static T phi_hooke2d(const Mx<T, 2>& F, const double* p, int flags) {  // :220-245
  Mx<T, 2> E = strain<T, 2>(F, flags & F_SMALL);
  const double Es = p[0], nu = p[1];
  if (!(flags & F_PLANE_STRESS)) {
    double lam, mu;
    lame(Es, nu, lam, mu);
    return (mu + lam / 2) * (sq(E(1)) + sq(E(4))) + lam * (E(1) * E(4)) + (2 * mu) * sq(E(2));
  }
  return (Es / (1 - nu * nu)) * (sq(E(1)) + sq(E(4)) + nu * (E(1) * E(4))) + (Es / (1 + nu)) * sq(E(2));
}
template <class T> static void invariants3(const Mx<T, 3>& C, T& I1, T& I2, T& I3) {  // materials.jl:111-114
  I1 = C(1) + C(5) + C(9);
  I2 = C(1) * C(5) + C(5) * C(9) + C(1) * C(9) - sq(C(2)) - sq(C(3)) - sq(C(6));
  I3 = C(1) * C(5) * C(9) + 2.0 * C(2) * C(3) * C(6) - C(1) * sq(C(6)) - C(5) * sq(C(3)) - C(9) * sq(C(2));
}
template <class T> static void invariants2(const Mx<T, 2>& C, const T& C33, T& I1, T& I2, T& I3) {  // :98-104
  I1 = C(1) + C(4) + C33;
  I2 = C(1) * C(4) + C(4) * C33 + C(1) * C33 - sq(C(2));
  I3 = C(1) * C(4) * C33 - C33 * sq(C(2));
}
template <class T> static T phi_I(T I1, T I2, T I3, int kind, const double* p) {  // :173-197
  if (kind == MOONEY) {
    const double C1 = p[0], C2 = p[1], K = p[2];
    if (K < 0) return C1 * (I1 - 3.0) + C2 * (I2 - 3.0);
    T J = dsqrt(I3);
    I1 = I1 * dpow(J, -2.0 / 3);
    I2 = I2 * dpow(J, -4.0 / 3);
    return C1 * (I1 - 3.0) + C2 * (I2 - 3.0) + K * sq(J - 1.0);
  }
  const double C1 = p[0], K = p[1];
  if (K < 0) return C1 * (I1 - 3.0);
  T J = dsqrt(I3);
  I1 = I1 * dpow(J, -2.0 / 3);
  return C1 * (I1 - 3.0) + K * sq(J - 1.0);
}
// Ogden (elasticmaterials.jl:148-172).  The reference defines only the scalar energy: λ_i = sqrt(svdvals(FᵀF)_i);
//   K < 0: μ/α (Σ λ_i^α − 3);   else J = Π λ_i,  μ/α (Σ λ_i^α / J^(α/3) − 3) + K (J − 1)²
// and cannot differentiate it (svdvals on duals).  The oracle therefore differentiates the SAME energy
// through a smooth matrix function instead of an eigen-decomposition: Σ λ_i^α = tr(U^α), U = (FᵀF)^(1/2) by the Newton
// iteration U <- (U + C U⁻¹)/2 from U = I (quadratic; ‖C − I‖ < 1/2 here), in the element's dual arithmetic -- which has
// no singular point at repeated stretches and shares nothing with the product's spectral divided-difference tangent
// (materials.cuh).  Integer α only (tr(U^α) by repeated multiplication); J = sqrt(det C).
template <class T> static Mx<T, 3> inv3(const Mx<T, 3>& A) {
  Mx<T, 3> B;
  const T c00 = A.at(1, 1) * A.at(2, 2) - A.at(1, 2) * A.at(2, 1);
  const T c01 = A.at(1, 2) * A.at(2, 0) - A.at(1, 0) * A.at(2, 2);
  const T c02 = A.at(1, 0) * A.at(2, 1) - A.at(1, 1) * A.at(2, 0);
  const T det = A.at(0, 0) * c00 + A.at(0, 1) * c01 + A.at(0, 2) * c02;
  const T id = 1.0 / det;
  B.at(0, 0) = c00 * id;
  B.at(1, 0) = c01 * id;
  B.at(2, 0) = c02 * id;
  B.at(0, 1) = (A.at(0, 2) * A.at(2, 1) - A.at(0, 1) * A.at(2, 2)) * id;
  B.at(1, 1) = (A.at(0, 0) * A.at(2, 2) - A.at(0, 2) * A.at(2, 0)) * id;
  B.at(2, 1) = (A.at(0, 1) * A.at(2, 0) - A.at(0, 0) * A.at(2, 1)) * id;
  B.at(0, 2) = (A.at(0, 1) * A.at(1, 2) - A.at(0, 2) * A.at(1, 1)) * id;
  B.at(1, 2) = (A.at(0, 2) * A.at(1, 0) - A.at(0, 0) * A.at(1, 2)) * id;
  B.at(2, 2) = (A.at(0, 0) * A.at(1, 1) - A.at(0, 1) * A.at(1, 0)) * id;
  return B;
}
template <class T> static Mx<T, 3> mul3(const Mx<T, 3>& A, const Mx<T, 3>& B) {
  Mx<T, 3> R;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) R.at(i, j) = A.at(i, 0) * B.at(0, j) + A.at(i, 1) * B.at(1, j) + A.at(i, 2) * B.at(2, j);
  return R;
}
template <class T> static T phi_ogden3(const Mx<T, 3>& F, const double* p) {
  const double alpha = p[0], mu = p[1], K = p[2];
  const int ia = (int)alpha;
  if ((double)ia != alpha || ia < 1 || ia > 8) throw std::runtime_error("oracle: Ogden needs an integer alpha in 1..8");
  const Mx<T, 3> C = FtF<T, 3>(F);
  Mx<T, 3> U;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) U.at(i, j) = Zero<T>::get() + (i == j ? 1.0 : 0.0);
  for (int it = 0; it < 9; it++) {
    const Mx<T, 3> W = mul3<T>(C, inv3<T>(U));
    for (int k = 0; k < 9; k++) U.a[k] = 0.5 * (U.a[k] + W.a[k]);
    for (int i = 0; i < 3; i++)  // keep the iterate symmetric (C and U commute in exact arithmetic)
      for (int j = i + 1; j < 3; j++) {
        const T s = 0.5 * (U.at(i, j) + U.at(j, i));
        U.at(i, j) = s;
        U.at(j, i) = s;
      }
  }
  Mx<T, 3> Pw = U;
  for (int k = 1; k < ia; k++) Pw = mul3<T>(Pw, U);
  const T S = Pw.at(0, 0) + Pw.at(1, 1) + Pw.at(2, 2);  // Σ λ_i^α
  if (K < 0) return (mu / alpha) * (S - 3.0);
  T I1, I2, I3;
  invariants3(C, I1, I2, I3);
  const T J = dsqrt(I3);
  return (mu / alpha) * (S * dpow(J, -alpha / 3) - 3.0) + K * sq(J - 1.0);
}
template <class T> static T phi_hyper(const Mx<T, 3>& F, int kind, const double* p) {  // :138-147
  if (kind == OGDEN) return phi_ogden3<T>(F, p);
  Mx<T, 3> C = FtF<T, 3>(F);
  T I1, I2, I3;
  invariants3(C, I1, I2, I3);
  return phi_I(I1, I2, I3, kind, p);
}
template <class T> static T phi_hyper(const Mx<T, 2>& F, int kind, const double* p) {
  if (kind == OGDEN) throw std::runtime_error("oracle: 2-D Ogden is not restated in C++ (use the Python oracle)");
  Mx<T, 2> C = FtF<T, 2>(F);
  const double K = kind == MOONEY ? p[2] : p[1];
  T L3 = K < 0 ? dpow(F(1) * F(4) - F(2) * F(3), -2.0) : (Zero<T>::get() + 1.0);
  T I1, I2, I3;
  invariants2(C, L3, I1, I2, I3);
  return phi_I(I1, I2, I3, kind, p);
}
template <class T> static T getphi(const Mx<T, 3>& F, const Mat& m) {
  if (m.kind == HOOKE) return phi_hooke<T>(F, m.p, m.flags);
  return phi_hyper<T>(F, m.kind, m.p);
}
template <class T> static T getphi(const Mx<T, 2>& F, const Mat& m) {
  if (m.kind == HOOKE2D) return phi_hooke2d<T>(F, m.p, m.flags);
  return phi_hyper<T>(F, m.kind, m.p);
}

// get1stinvariants / gethyddevdecomp  (materials.jl:117-155, elasticmaterials.jl:264-333)
template <class T> static void first_inv(const Mx<T, 3>& F, const Mat& m, T& I1, T& I1sq) {
  Mx<T, 3> E = strain<T, 3>(F, m.flags & F_SMALL);
  I1 = E(1) + E(5) + E(9);
  I1sq = sq(E(1)) + sq(E(5)) + sq(E(9)) + 2.0 * (sq(E(2)) + sq(E(3)) + sq(E(6)));
}
template <class T> static void first_inv(const Mx<T, 2>& F, const Mat& m, T& I1, T& I1sq) {
  Mx<T, 2> E = strain<T, 2>(F, m.flags & F_SMALL);
  if (!(m.flags & F_PLANE_STRESS)) {
    I1 = E(1) + E(4);
    I1sq = sq(E(1)) + sq(E(4)) + 2.0 * sq(E(3));
  } else {
    const double nu = m.p[1];
    T E33 = (nu / (nu - 1)) * (E(1) + E(4));
    I1 = E(1) + E(4) + E33;
    I1sq = sq(E(1)) + sq(E(4)) + sq(E33) + 2.0 * sq(E(3));
  }
}
template <class T> static void hyddev(const Mx<T, 3>& F, const Mat& m, T& I1, T& ed) {
  Mx<T, 3> E = strain<T, 3>(F, m.flags & F_SMALL);
  I1 = E(1) + E(5) + E(9);
  Mx<T, 3> e;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) e.at(i, j) = i == j ? E.at(i, j) - I1 / 3.0 : E.at(i, j);
  T s = e(1) * e(1);
  for (int k = 2; k <= 9; k++) s = s + e(k) * e(k);
  ed = s;
}
template <class T> static void hyddev(const Mx<T, 2>& F, const Mat& m, T& I1, T& ed) {
  Mx<T, 2> E = strain<T, 2>(F, m.flags & F_SMALL);
  T E33 = Zero<T>::get();
  const bool ps = m.flags & F_PLANE_STRESS;
  if (ps) {
    const double nu = m.p[1];
    E33 = (nu / (nu - 1)) * (E(1) + E(4));
    I1 = E(1) + E(4) + E33;
  } else {
    I1 = E(1) + E(4);
  }
  Mx<T, 2> e;
  for (int i = 0; i < 2; i++)
    for (int j = 0; j < 2; j++) e.at(i, j) = i == j ? E.at(i, j) - I1 / 3.0 : E.at(i, j);
  T s = e(1) * e(1);
  for (int k = 2; k <= 4; k++) s = s + e(k) * e(k);
  ed = ps ? s + sq(E33 - I1 / 3.0) : s + sq(I1 / 3.0);
}

template <class T> static inline T detF(const Mx<T, 3>& F) {  // phasefieldmaterials.jl:42-44
  return F(1) * F(5) * F(9) - F(1) * F(6) * F(8) - F(2) * F(4) * F(9) + F(2) * F(6) * F(7) + F(3) * F(4) * F(8) -
         F(3) * F(5) * F(7);
}
template <class T> static inline T detF(const Mx<T, 2>& F) { return F(1) * F(4) - F(2) * F(3); }
template <class T> static inline T sumF2(const Mx<T, 3>& F) {  // :45-47
  return sq(F(1)) + sq(F(2)) + sq(F(3)) + sq(F(4)) + sq(F(5)) + sq(F(6)) + sq(F(7)) + sq(F(8)) + sq(F(9));
}
template <class T> static inline T sumF2(const Mx<T, 2>& F) {  // plane-strain embedding
  return sq(F(1)) + sq(F(2)) + sq(F(3)) + sq(F(4)) + 1.0;
}

// PhaseField energies.  TF: type of F entries, TD: type of d.  Result type R is whichever is dual.
// hist: nullptr (no history) or pointer to 2 doubles, updated in place (phasefieldmaterials.jl:75-135).
template <class R, class TF, class TD, int D>
static R getphi_pf(const Mx<TF, D>& F, const TD& d, const TD* gd, const Mat& m, double* hist) {
  const double l0 = m.p[4], Gc = m.p[5], n = m.p[6];
  TD gg = gd[0] * gd[0];
  for (int k = 1; k < D; k++) gg = gg + gd[k] * gd[k];
  TD gamma = dpow(d, n) + (l0 * l0) * gg;
  TD omd2 = sq(1.0 - d);
  R psi;
  if (m.base == NEOHOOKE && !(m.flags & F_DHN)) {
    const double C1 = m.p[0], K = m.p[1];
    TF J = detF<TF>(F);
    TF I1 = sumF2<TF>(F);
    if (!hist) {
      TF phiel = C1 * (I1 - 3.0 - 2.0 * dlog(J)) + K * sq(J - 1.0);
      psi = val(J) >= 1 ? Conv<R>::from(omd2 * phiel) : Conv<R>::from(phiel);
    } else {
      if (val(J) >= 1) {
        const double e = val(C1 * (I1 - 3.0 - 2.0 * dlog(J)) + K * sq(J - 1.0));
        hist[0] = e > hist[0] ? e : hist[0];  // max(ϕmax, e)
        psi = Conv<R>::from(omd2 * hist[0]);
      } else {
        psi = Conv<R>::from(C1 * (I1 - 3.0 - 2.0 * dlog(J)) + K * sq(J - 1.0));
      }
    }
  } else {
    double lam, mu;
    lame(m.p[0], m.p[1], lam, mu);
    if (!(m.flags & F_DHN)) {
      TF I1, I1sq;
      first_inv<TF>(F, m, I1, I1sq);
      if (!hist) {
        if (val(I1) >= 0)
          psi = Conv<R>::from(omd2 * ((lam / 2) * sq(I1) + mu * I1sq));
        else
          psi = Conv<R>::from((lam / 2) * sq(I1) + omd2 * mu * I1sq);
      } else {
        if (val(I1) >= 0) {
          const double e = val((lam / 2) * sq(I1) + mu * I1sq);
          hist[0] = e > hist[0] ? e : hist[0];
          psi = Conv<R>::from(omd2 * hist[0]);
        } else {
          const double e = val(mu * I1sq);
          hist[1] = e > hist[1] ? e : hist[1];
          psi = Conv<R>::from((lam / 2) * sq(I1) + omd2 * hist[1]);
        }
      }
    } else {
      TF I1, ed;
      hyddev<TF>(F, m, I1, ed);
      const double K = lam + mu / 3;
      if (!hist) {
        if (val(I1) >= 0)
          psi = Conv<R>::from(omd2 * ((K / 2) * sq(I1) + mu * ed));
        else
          psi = Conv<R>::from((K / 2) * sq(I1) + omd2 * mu * ed);
      } else {
        if (val(I1) >= 0) {
          const double e = val((K / 2) * sq(I1) + mu * ed);
          hist[0] = e > hist[0] ? e : hist[0];
          psi = Conv<R>::from(omd2 * hist[0]);
        } else {
          const double e = val(mu * ed);
          hist[1] = e > hist[1] ? e : hist[1];
          psi = Conv<R>::from((lam / 2) * sq(I1) + omd2 * hist[1]);  // λ/2 as written (:131)
        }
      }
    }
  }
  return psi + (Gc / (2 * l0)) * gamma;
}

// ------------------------------------------------------------------------------------------------
// element layer
// ------------------------------------------------------------------------------------------------
struct Batch {
  int kind;  // 0 = CEElem, 1 = CPElem, 2 = CAS (axisymmetric extension; Nval = N / X0)
  int D, P, N;
  Mat mat;
  i64 ne;
  const i64* nodes;     // [ne][N], 1-based
  const double* gradN;  // [ne][D][P][N]
  const double* wgt;    // [ne][P]
  const double* Nval;   // [ne][P][N]  (CPElem)
};

// Elements.dot: s = zero(T); s += a[i]*b[i]   (elements.jl:22-28)
template <class T, int N> static inline T edot(const double* a, const T* b, int stride) {
  T s = Zero<T>::get();
  for (int i = 0; i < N; i++) s = s + a[i] * b[i * stride];
  return s;
}

// getF: F[j,k] = ∇N[k][gp]·u[j,:] (+I)   (toolkit:8-20);  ue is D x N column-major: ue[a*D + j]
template <class T, int D, int N> static inline Mx<T, D> getF(const double* gN, int P, int gp, const T* ue) {
  Mx<T, D> F;
  for (int j = 0; j < D; j++)
    for (int k = 0; k < D; k++) F.at(j, k) = edot<T, N>(gN + (k * P + gp) * N, ue + j, D);
  for (int j = 0; j < D; j++) F.at(j, j) = F.at(j, j) + 1.0;
  return F;
}

// chain operator ×  (toolkit:231-245): ϕ::D2{DD}, F linear-indexed D1{n}
template <int DD, int n> static void chain_add(const D2<DD>& phi, const D1<n>* F, D2<n>& acc) {
  D2<n> r;
  r.v = phi.v;
  for (int p = 0; p < n; p++) r.g[p] = 0.0;
  for (int p = 0; p < D2<n>::m; p++) r.h[p] = 0.0;
  for (int ii = 0; ii < DD; ii++)
    for (int p = 0; p < n; p++) r.g[p] = r.g[p] + phi.g[ii] * F[ii].g[p];
  for (int ii = 1; ii < DD; ii++)
    for (int jj = 0; jj < ii; jj++) {
      const double hij = phi.h[ii * (ii + 1) / 2 + jj];
      int k = 0;
      for (int q = 0; q < n; q++)
        for (int p = 0; p <= q; p++, k++)
          r.h[k] = r.h[k] + hij * (F[ii].g[p] * F[jj].g[q] + F[jj].g[p] * F[ii].g[q]);
    }
  for (int ii = 0; ii < DD; ii++) {
    const double hii = phi.h[ii * (ii + 1) / 2 + ii];
    int k = 0;
    for (int q = 0; q < n; q++)
      for (int p = 0; p <= q; p++, k++) r.h[k] = r.h[k] + (hii * F[ii].g[p]) * F[ii].g[q];
  }
  acc = acc + r;
}

template <int n> static void store(const D2<n>& phi, double* val_, double* g, double* H) {
  *val_ = phi.v;
  for (int i = 0; i < n; i++) g[i] = phi.g[i];
  int k = 0;
  for (int j = 0; j < n; j++)
    for (int i = 0; i <= j; i++, k++) {
      H[j * n + i] = phi.h[k];
      H[i * n + j] = phi.h[k];
    }
}

enum Mode { M_ELASTIC = 0, M_PF_U = 1, M_PF_D = 2 };

// u-dual element energy (elastic or phase-field u-phase)
template <int D, int N>
static void elem_u(const Batch& b, i64 e, int mode, const double* u, int dpn, const double* dfield, double* val_,
                   double* g, double* H) {
  constexpr int n = D * N;
  const int P = b.P;
  const i64* nd = b.nodes + e * N;
  const double* gN = b.gradN + e * (i64)D * P * N;
  const double* w = b.wgt + e * P;
  const double* Nv = b.Nval ? b.Nval + e * (i64)P * N : nullptr;
  double de[N];
  if (mode == M_PF_U)
    for (int a = 0; a < N; a++) de[a] = dfield[nd[a] - 1];
  D2<n>* phi = new D2<n>(lift<n>(0.0));
  if (D == 3 && (b.kind == 0 || mode == M_PF_U)) {
    // chain path (elasticelements.jl:339-350, phasefieldelements.jl:194-207)
    D1<n>* ue = new D1<n>[n];
    for (int a = 0; a < N; a++)
      for (int j = 0; j < D; j++) {
        D1<n>& x = ue[a * D + j];
        x.v = u[(nd[a] - 1) * (i64)dpn + j];
        for (int q = 0; q < n; q++) x.g[q] = (q == a * D + j) ? 1.0 : 0.0;
      }
    for (int gp = 0; gp < P; gp++) {
      Mx<D1<n>, D> F = getF<D1<n>, D, N>(gN, P, gp, ue);
      Mx<D2<D * D>, D> Fd;
      for (int k = 0; k < D * D; k++) {
        Fd.a[k] = lift<D * D>(F.a[k].v);
        Fd.a[k].g[k] = 1.0;
      }
      D2<D * D> dphi;
      if (mode == M_PF_U) {
        double dv = edot<double, N>(Nv + gp * N, de, 1);
        double gd[D];
        for (int j = 0; j < D; j++) gd[j] = edot<double, N>(gN + (j * P + gp) * N, de, 1);
        dphi = getphi_pf<D2<D * D>, D2<D * D>, double, D>(Fd, dv, gd, b.mat, nullptr);
      } else {
        dphi = getphi<D2<D * D>>(Fd, b.mat);
      }
      chain_add<D * D, n>(w[gp] * dphi, F.a, *phi);
    }
    delete[] ue;
  } else {
    // generic path: all arithmetic in D2{n}  (elasticelements.jl:199-206, phasefieldelements.jl:159-169)
    D2<n>* ue = new D2<n>[n];
    for (int a = 0; a < N; a++)
      for (int j = 0; j < D; j++) {
        ue[a * D + j] = lift<n>(u[(nd[a] - 1) * (i64)dpn + j]);
        ue[a * D + j].g[a * D + j] = 1.0;
      }
    for (int gp = 0; gp < P; gp++) {
      if (b.kind == 2) {
        // CAS (axisymmetric; EXTENSION, no reference getϕ -- SURVEY §9-7, constructors elasticelements.jl:235-284):
        // 3x3 F = in-plane block I + ∇u (rows r, z) and the hoop stretch 1 + (N·u_r)/X0; Nval holds N/X0 here
        if constexpr (D == 2) {
          Mx<D2<n>, 2> F2 = getF<D2<n>, 2, N>(gN, P, gp, ue);
          Mx<D2<n>, 3> F3;
          for (int k = 0; k < 9; k++) F3.a[k] = lift<n>(0.0);
          for (int j = 0; j < 2; j++)
            for (int k = 0; k < 2; k++) F3.at(j, k) = F2.at(j, k);
          F3.at(2, 2) = edot<D2<n>, N>(Nv + gp * N, ue, D) + 1.0;
          *phi = *phi + w[gp] * getphi<D2<n>>(F3, b.mat);
        }
        continue;
      }
      Mx<D2<n>, D> F = getF<D2<n>, D, N>(gN, P, gp, ue);
      if (mode == M_PF_U) {
        double dv = edot<double, N>(Nv + gp * N, de, 1);
        double gd[D];
        for (int j = 0; j < D; j++) gd[j] = edot<double, N>(gN + (j * P + gp) * N, de, 1);
        *phi = *phi + w[gp] * getphi_pf<D2<n>, D2<n>, double, D>(F, dv, gd, b.mat, nullptr);
      } else {
        *phi = *phi + w[gp] * getphi<D2<n>>(F, b.mat);
      }
    }
    delete[] ue;
  }
  store<n>(*phi, val_, g, H);
  delete phi;
}

// d-dual element energy (phase-field d-phase), optional history [P][2]
template <int D, int N>
static void elem_d(const Batch& b, i64 e, const double* u, int dpn, const double* dfield, double* hist, double* val_,
                   double* g, double* H) {
  constexpr int n = N;
  const int P = b.P;
  const i64* nd = b.nodes + e * N;
  const double* gN = b.gradN + e * (i64)D * P * N;
  const double* w = b.wgt + e * P;
  const double* Nv = b.Nval + e * (i64)P * N;
  double ue[D * N];
  for (int a = 0; a < N; a++)
    for (int j = 0; j < D; j++) ue[a * D + j] = u[(nd[a] - 1) * (i64)dpn + j];
  D2<n> de[N];
  for (int a = 0; a < N; a++) {
    de[a] = lift<n>(dfield[nd[a] - 1]);
    de[a].g[a] = 1.0;
  }
  const bool intended = b.mat.flags & F_PF_GRAD_INTENDED;
  D2<n> phi = lift<n>(0.0);
  for (int gp = 0; gp < P; gp++) {
    Mx<double, D> F = getF<double, D, N>(gN, P, gp, ue);
    D2<n> dv = edot<D2<n>, N>(Nv + gp * N, de, 1);
    D2<n> gd[D];
    for (int j = 0; j < D; j++) {
      gd[j] = edot<D2<n>, N>(gN + (j * P + gp) * N, de, 1);
      if (!intended) gd[j] = lift<n>(gd[j].v);  // Float64 truncation, toolkit:162-165 + adiff.jl:74
    }
    phi = phi + w[gp] * getphi_pf<D2<n>, double, D2<n>, D>(F, dv, gd, b.mat, hist ? hist + (e * P + gp) * 2 : nullptr);
  }
  store<n>(phi, val_, g, H);
}

template <int D, int N>
static void run_batch(const Batch& b, int mode, const double* u, int dpn, const double* d, double* hist, double* val_,
                      double* g, double* H, int nthreads) {
  const int n = mode == M_PF_D ? N : D * N;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel for schedule(static)
  for (i64 e = 0; e < b.ne; e++) {
    if (mode == M_PF_D)
      elem_d<D, N>(b, e, u, dpn, d, hist, val_ + e, g + e * n, H + e * (i64)n * n);
    else
      elem_u<D, N>(b, e, mode, u, dpn, d, val_ + e, g + e * n, H + e * (i64)n * n);
  }
}

extern "C" {

int oracle_max_threads() {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// Per-element duals Φ[e] = (val, grad[n], hess[n x n] full symmetric, column-major).
// mat_desc = {kind, base, flags}; returns 0 on success, <0 for an unsupported (D, N).
int oracle_elem_duals(int ekind, int D, int P, int N, const int* mat_desc, const double* mat_params, i64 ne,
                      const i64* nodes, const double* gradN, const double* wgt, const double* Nval, int mode,
                      const double* u, int dpn, const double* d, double* hist, int nthreads, double* val_, double* g,
                      double* H) {
  Batch b;
  b.kind = ekind;
  b.D = D;
  b.P = P;
  b.N = N;
  b.mat.kind = mat_desc[0];
  b.mat.base = mat_desc[1];
  b.mat.flags = mat_desc[2];
  for (int i = 0; i < 8; i++) b.mat.p[i] = mat_params[i];
  b.ne = ne;
  b.nodes = nodes;
  b.gradN = gradN;
  b.wgt = wgt;
  b.Nval = Nval;
#define RUN(DD, NN)                                                          \
  if (D == DD && N == NN) {                                                  \
    run_batch<DD, NN>(b, mode, u, dpn, d, hist, val_, g, H, nthreads);       \
    return 0;                                                                \
  }
  RUN(2, 3) RUN(2, 4) RUN(3, 4) RUN(3, 6) RUN(3, 8)
#undef RUN
  return -1;
}

// Serial assembly  (toolkit:169-203): COO triplets in element order -> CSC with duplicates
// summed as a left fold in input order (Julia `sparse`), then `dropzeros`.
struct AsmResult {
  std::vector<i64> colptr, rowval;
  std::vector<double> nzval;
};

void* oracle_assemble(i64 ne, int N, int dpn, int n, const i64* nodes, const double* val_, const double* g,
                      const double* H, i64 ndof, double* phi_out, double* r) {
  // n = dofs per element = dpn*N  (toolkit:187: idx = LinearIndices(u)[:, nodes][:])
  const i64 Nt = ne * (i64)n * n;
  std::vector<i64> II(Nt), JJ(Nt);
  std::vector<double> KK(Nt);
  for (i64 i = 0; i < ndof; i++) r[i] = 0.0;
  double phi = 0;
  std::vector<i64> idx(n);
  i64 N1 = 0;
  for (i64 e = 0; e < ne; e++) {
    for (int a = 0; a < N; a++)
      for (int j = 0; j < dpn; j++) idx[a * dpn + j] = (nodes[e * N + a] - 1) * dpn + j + 1;
    phi += val_[e];
    for (int p = 0; p < n; p++) r[idx[p] - 1] += g[e * n + p];
    for (int q = 0; q < n; q++)
      for (int p = 0; p < n; p++) {
        II[N1] = idx[p];
        JJ[N1] = idx[q];
        KK[N1] = H[e * (i64)n * n + q * n + p];
        N1++;
      }
  }
  *phi_out = phi;
  // --- Julia sparse(): counting sort by row (stable), fold duplicates per row in input order,
  //     then transpose to CSC (stable => rows ascending in each column).
  std::vector<i64> rp(ndof + 2, 0);
  for (i64 k = 0; k < Nt; k++) rp[II[k] + 1]++;
  for (i64 i = 1; i <= ndof + 1; i++) rp[i] += rp[i - 1];
  std::vector<i64> cj(Nt);
  std::vector<double> cv(Nt);
  {
    std::vector<i64> pos(rp.begin(), rp.end() - 1);
    for (i64 k = 0; k < Nt; k++) {
      i64 p = pos[II[k]]++;
      cj[p] = JJ[k];
      cv[p] = KK[k];
    }
  }
  std::vector<i64>().swap(II);
  std::vector<i64>().swap(JJ);
  std::vector<double>().swap(KK);
  std::vector<i64> last(ndof + 1, -1);
  i64 w = 0;
  std::vector<i64> rp2(ndof + 2, 0);
  for (i64 i = 1; i <= ndof; i++) {
    const i64 rs = w;
    for (i64 k = rp[i]; k < rp[i + 1]; k++) {
      const i64 j = cj[k];
      if (last[j] >= rs) {
        cv[last[j]] += cv[k];
      } else {
        last[j] = w;
        cj[w] = j;
        cv[w] = cv[k];
        w++;
      }
    }
    rp2[i + 1] = w;
  }
  // dropzeros + transpose to CSC
  AsmResult* R = new AsmResult;
  R->colptr.assign(ndof + 1, 1);
  {
    std::vector<i64> cnt(ndof + 1, 0);
    for (i64 i = 1; i <= ndof; i++)
      for (i64 k = rp2[i]; k < rp2[i + 1]; k++)
        if (cv[k] != 0.0) cnt[cj[k]]++;
    R->colptr[0] = 1;
    for (i64 j = 1; j <= ndof; j++) R->colptr[j] = R->colptr[j - 1] + cnt[j];
  }
  const i64 nnz = R->colptr[ndof] - 1;
  R->rowval.resize(nnz);
  R->nzval.resize(nnz);
  {
    std::vector<i64> pos(R->colptr.begin(), R->colptr.end() - 1);
    for (i64 i = 1; i <= ndof; i++)
      for (i64 k = rp2[i]; k < rp2[i + 1]; k++)
        if (cv[k] != 0.0) {
          i64 p = pos[cj[k] - 1]++ - 1;
          R->rowval[p] = i;
          R->nzval[p] = cv[k];
        }
  }
  return R;
}
i64 oracle_result_nnz(void* h) { return (i64)((AsmResult*)h)->nzval.size(); }
void oracle_result_copy(void* h, i64* colptr, i64* rowval, double* nzval) {
  AsmResult* R = (AsmResult*)h;
  std::memcpy(colptr, R->colptr.data(), R->colptr.size() * sizeof(i64));
  std::memcpy(rowval, R->rowval.data(), R->rowval.size() * sizeof(i64));
  std::memcpy(nzval, R->nzval.data(), R->nzval.size() * sizeof(double));
}
void oracle_result_free(void* h) { delete (AsmResult*)h; }

}  // extern "C"
