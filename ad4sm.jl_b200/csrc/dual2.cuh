// dual2.cuh -- second-order forward duals that live in registers.
//
// Same algebra as the reference's `adiff.D2{N,M,T}` (/root/reference/src/adiff.jl:38-42,131-150):
// value, N-gradient and packed upper-triangular Hessian, (i<=j) -> j(j+1)/2+i.  In the kernels the
// duals run over the 2-3 scalar invariants a material depends on (N <= 3, so a dual is <= 10
// doubles); the chain to the D*D entries of F and on to the element dofs is done analytically
// (materials.cuh, element kernels), which is the "AD over F followed by the B^T.A.B contraction"
// form north_star allows, at a fraction of the flops of D2{9,45}.
#pragma once
#include <cmath>
#ifdef __CUDACC__
#include <cuda_runtime.h>
#define AD4SM_DI __host__ __device__ __forceinline__
#define AD4SM_HD __host__ __device__
#else
// host-only compilation (tests/cxx builds these headers with g++ to check the math without a GPU)
#define AD4SM_DI inline
#define AD4SM_HD
#endif

namespace ad4sm {

template <int N> struct Dual2 {
  static constexpr int M = N * (N + 1) / 2;
  double v;
  double g[N];
  double h[M];
};


template <int N> AD4SM_DI Dual2<N> d2_const(double x) {
  Dual2<N> r;
  r.v = x;
#pragma unroll
  for (int i = 0; i < N; i++) r.g[i] = 0.0;
#pragma unroll
  for (int i = 0; i < Dual2<N>::M; i++) r.h[i] = 0.0;
  return r;
}
// seed variable k (adiff.jl:58-62)
template <int N> AD4SM_DI Dual2<N> d2_var(double x, int k) {
  Dual2<N> r = d2_const<N>(x);
#pragma unroll
  for (int i = 0; i < N; i++)
    if (i == k) r.g[i] = 1.0;
  return r;
}
template <int N> AD4SM_DI Dual2<N> operator+(const Dual2<N>& x, const Dual2<N>& y) {  // adiff.jl:132
  Dual2<N> r;
  r.v = x.v + y.v;
#pragma unroll
  for (int i = 0; i < N; i++) r.g[i] = x.g[i] + y.g[i];
#pragma unroll
  for (int i = 0; i < Dual2<N>::M; i++) r.h[i] = x.h[i] + y.h[i];
  return r;
}
template <int N> AD4SM_DI Dual2<N> operator-(const Dual2<N>& x, const Dual2<N>& y) {  // adiff.jl:133
  Dual2<N> r;
  r.v = x.v - y.v;
#pragma unroll
  for (int i = 0; i < N; i++) r.g[i] = x.g[i] - y.g[i];
#pragma unroll
  for (int i = 0; i < Dual2<N>::M; i++) r.h[i] = x.h[i] - y.h[i];
  return r;
}
template <int N> AD4SM_DI Dual2<N> operator*(const Dual2<N>& x, const Dual2<N>& y) {  // adiff.jl:135
  Dual2<N> r;
  r.v = x.v * y.v;
#pragma unroll
  for (int i = 0; i < N; i++) r.g[i] = x.v * y.g[i] + y.v * x.g[i];
  int k = 0;
#pragma unroll
  for (int j = 0; j < N; j++)
#pragma unroll
    for (int i = 0; i <= j; i++, k++) r.h[k] = x.v * y.h[k] + y.v * x.h[k] + x.g[i] * y.g[j] + y.g[i] * x.g[j];
  return r;
}
// real (.) dual: the reference promotes the real to a dual with zero derivatives (adiff.jl:78-80);
// the values are identical to scaling / shifting.
template <int N> AD4SM_DI Dual2<N> operator*(double a, const Dual2<N>& y) {
  Dual2<N> r;
  r.v = a * y.v;
#pragma unroll
  for (int i = 0; i < N; i++) r.g[i] = a * y.g[i];
#pragma unroll
  for (int i = 0; i < Dual2<N>::M; i++) r.h[i] = a * y.h[i];
  return r;
}
template <int N> AD4SM_DI Dual2<N> operator*(const Dual2<N>& y, double a) { return a * y; }
template <int N> AD4SM_DI Dual2<N> operator+(const Dual2<N>& x, double b) {
  Dual2<N> r = x;
  r.v = x.v + b;
  return r;
}
template <int N> AD4SM_DI Dual2<N> operator+(double b, const Dual2<N>& x) { return x + b; }
template <int N> AD4SM_DI Dual2<N> operator-(const Dual2<N>& x, double b) {
  Dual2<N> r = x;
  r.v = x.v - b;
  return r;
}
template <int N> AD4SM_DI Dual2<N> operator-(double b, const Dual2<N>& x) {
  Dual2<N> r;
  r.v = b - x.v;
#pragma unroll
  for (int i = 0; i < N; i++) r.g[i] = -x.g[i];
#pragma unroll
  for (int i = 0; i < Dual2<N>::M; i++) r.h[i] = -x.h[i];
  return r;
}
// generic scalar function application: f(x) given f, f', f'' at x.v
// (the pattern of adiff.jl:136-147: (f, f' g, f'' g⊗g + f' h))
template <int N> AD4SM_DI Dual2<N> d2_apply(const Dual2<N>& x, double f0, double f1, double f2) {
  Dual2<N> r;
  r.v = f0;
#pragma unroll
  for (int i = 0; i < N; i++) r.g[i] = f1 * x.g[i];
  int k = 0;
#pragma unroll
  for (int j = 0; j < N; j++)
#pragma unroll
    for (int i = 0; i <= j; i++, k++) r.h[k] = f2 * (x.g[i] * x.g[j]) + f1 * x.h[k];
  return r;
}
template <int N> AD4SM_DI Dual2<N> d2_inv(const Dual2<N>& x) {  // adiff.jl:136
  const double i1 = 1.0 / x.v;
  return d2_apply(x, i1, -i1 * i1, 2.0 * i1 * i1 * i1);
}
template <int N> AD4SM_DI Dual2<N> d2_sq(const Dual2<N>& x) {  // x^2, adiff.jl:139
  return d2_apply(x, x.v * x.v, 2.0 * x.v, 2.0);
}
template <int N> AD4SM_DI Dual2<N> d2_log(const Dual2<N>& x) {  // adiff.jl:140
  const double i1 = 1.0 / x.v;
  return d2_apply(x, log(x.v), i1, -i1 * i1);
}
template <int N> AD4SM_DI Dual2<N> d2_sqrt(const Dual2<N>& x) {  // adiff.jl:147 (x^0.5)
  const double s = sqrt(x.v);
  return d2_apply(x, s, 0.5 / s, -0.25 / (s * x.v));
}
// x^p for real p (adiff.jl:138); vp = x.v^p supplied by the caller so cbrt-based forms can be used
template <int N> AD4SM_DI Dual2<N> d2_pow_given(const Dual2<N>& x, double p, double vp) {
  const double i1 = 1.0 / x.v;
  return d2_apply(x, vp, p * vp * i1, p * (p - 1.0) * vp * i1 * i1);
}
template <int N> AD4SM_DI Dual2<N> d2_pow(const Dual2<N>& x, double p) { return d2_pow_given(x, p, pow(x.v, p)); }

}  // namespace ad4sm
