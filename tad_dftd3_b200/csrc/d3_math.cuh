// Scalar layer of the D3 kernels: one set of device functions, instantiated on
//   float, double                     (energies and analytic first derivatives)
//   Dual<float>, Dual<double>         (value + one tangent: the second-order /
//                                      Hessian-vector path is the directional
//                                      derivative of the analytic gradient)
//
// Everything the pair kernels need is expressed through `+ - *`, `rcp`, `rsq`,
// `sqr`, `ex` and `val`, so the kernels are written once.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace d3 {

template <typename T>
struct Dual {
  T v, d;
  __host__ __device__ Dual() {}
  __host__ __device__ Dual(T v_) : v(v_), d(T(0)) {}
  __host__ __device__ Dual(T v_, T d_) : v(v_), d(d_) {}
};

#define D3_HD __host__ __device__ __forceinline__

template <typename T> D3_HD Dual<T> operator+(Dual<T> a, Dual<T> b) { return Dual<T>(a.v + b.v, a.d + b.d); }
template <typename T> D3_HD Dual<T> operator-(Dual<T> a, Dual<T> b) { return Dual<T>(a.v - b.v, a.d - b.d); }
template <typename T> D3_HD Dual<T> operator-(Dual<T> a) { return Dual<T>(-a.v, -a.d); }
template <typename T> D3_HD Dual<T> operator*(Dual<T> a, Dual<T> b) { return Dual<T>(a.v * b.v, a.v * b.d + a.d * b.v); }
template <typename T> D3_HD Dual<T> operator+(Dual<T> a, T b) { return Dual<T>(a.v + b, a.d); }
template <typename T> D3_HD Dual<T> operator+(T a, Dual<T> b) { return Dual<T>(a + b.v, b.d); }
template <typename T> D3_HD Dual<T> operator-(Dual<T> a, T b) { return Dual<T>(a.v - b, a.d); }
template <typename T> D3_HD Dual<T> operator-(T a, Dual<T> b) { return Dual<T>(a - b.v, -b.d); }
template <typename T> D3_HD Dual<T> operator*(Dual<T> a, T b) { return Dual<T>(a.v * b, a.d * b); }
template <typename T> D3_HD Dual<T> operator*(T a, Dual<T> b) { return Dual<T>(a * b.v, a * b.d); }
template <typename T> D3_HD Dual<T>& operator+=(Dual<T>& a, Dual<T> b) { a.v += b.v; a.d += b.d; return a; }
template <typename T> D3_HD Dual<T>& operator-=(Dual<T>& a, Dual<T> b) { a.v -= b.v; a.d -= b.d; return a; }

// ---- traits ---------------------------------------------------------------
template <typename S> struct Traits;
template <> struct Traits<float> {
  using base = float; using wide = double; static constexpr bool dual = false;
};
template <> struct Traits<double> {
  using base = double; using wide = double; static constexpr bool dual = false;
};
template <> struct Traits<Dual<float>> {
  using base = float; using wide = Dual<double>; static constexpr bool dual = true;
};
template <> struct Traits<Dual<double>> {
  using base = double; using wide = Dual<double>; static constexpr bool dual = true;
};

// machine epsilon of the base type (the reference clamps squared distances at
// finfo(dtype).eps, tad-mctc storch.cdist)
template <typename T> D3_HD T eps_of();
template <> D3_HD float eps_of<float>() { return 1.1920928955078125e-07f; }
template <> D3_HD double eps_of<double>() { return 2.220446049250313e-16; }

// ---- value / tangent access -------------------------------------------------
D3_HD float val(float a) { return a; }
D3_HD double val(double a) { return a; }
template <typename T> D3_HD T val(Dual<T> a) { return a.v; }

D3_HD float tan_of(float) { return 0.f; }
D3_HD double tan_of(double) { return 0.0; }
template <typename T> D3_HD T tan_of(Dual<T> a) { return a.d; }

template <typename S, typename T> D3_HD S make(T v, T d);
template <> D3_HD float make<float, float>(float v, float) { return v; }
template <> D3_HD double make<double, double>(double v, double) { return v; }
template <> D3_HD Dual<float> make<Dual<float>, float>(float v, float d) { return Dual<float>(v, d); }
template <> D3_HD Dual<double> make<Dual<double>, double>(double v, double d) { return Dual<double>(v, d); }

// ---- elementary functions ---------------------------------------------------
D3_HD float rcp(float a) { return 1.0f / a; }
D3_HD double rcp(double a) { return 1.0 / a; }
template <typename T> D3_HD Dual<T> rcp(Dual<T> a) {
  T r = rcp(a.v);
  return Dual<T>(r, -a.d * r * r);
}

// IEEE division (the weight normalisation divides denormals by denormals)
D3_HD float dv(float a, float b) { return a / b; }
D3_HD double dv(double a, double b) { return a / b; }
template <typename T> D3_HD Dual<T> dv(Dual<T> a, Dual<T> b) {
  T q = a.v / b.v;
  return Dual<T>(q, (a.d - q * b.d) / b.v);
}

D3_HD float rsq(float a) { return rsqrtf(a); }
D3_HD double rsq(double a) { return rsqrt(a); }
template <typename T> D3_HD Dual<T> rsq(Dual<T> a) {
  T r = rsq(a.v);
  return Dual<T>(r, T(-0.5) * a.d * r * r * r);
}

D3_HD float sqr(float a) { return sqrtf(a); }
D3_HD double sqr(double a) { return sqrt(a); }
template <typename T> D3_HD Dual<T> sqr(Dual<T> a) {
  T s = sqr(a.v);
  // (sqrt at exactly 0: the reference clamps |C6 C6 C6| at eps before its sqrt, whose gradient is 0 there)
  return Dual<T>(s, s > T(0) ? T(0.5) * a.d / s : T(0));
}

D3_HD float ex(float a) { return expf(a); }
D3_HD double ex(double a) { return exp(a); }
template <typename T> D3_HD Dual<T> ex(Dual<T> a) {
  T e = ex(a.v);
  return Dual<T>(e, e * a.d);
}

D3_HD float er(float a) { return erff(a); }
D3_HD double er(double a) { return erf(a); }
template <typename T> D3_HD Dual<T> er(Dual<T> a) {
  return Dual<T>(er(a.v), T(1.1283791670955126) * ex(-(a.v * a.v)) * a.d);   // 2 / sqrt(pi)
}

D3_HD float lg(float a) { return logf(a); }
D3_HD double lg(double a) { return log(a); }
template <typename T> D3_HD Dual<T> lg(Dual<T> a) { return Dual<T>(lg(a.v), a.d / a.v); }

D3_HD float ab(float a) { return fabsf(a); }
D3_HD double ab(double a) { return fabs(a); }
template <typename T> D3_HD Dual<T> ab(Dual<T> a) { return a.v < T(0) ? Dual<T>(-a.v, -a.d) : a; }

// x^p for x > 0 and a plain (tangent-free) exponent
template <typename S, typename B> D3_HD S pw(S x, B p) { return ex(p * lg(x)); }

// ---- fast, branch-free variants for the pair loops --------------------------
// Arguments in the pair loops are positive, finite and far from the
// denormal range, so the special-case handling of the library routines (which
// costs about half of the issue slots, see profiles/r1a_summary.md) is dropped:
//   frsq : MUFU.RSQ64H seed (2^-22) + one cubic step      -> 5 FP64 instructions
//   frcp : MUFU.RCP64H seed (2^-22) + one cubic step      -> 3 FP64 instructions
//   fex  : 2^n * P11(r), |r| <= ln2/2, Taylor degree 11   -> 16 FP64 instructions
// each good to a few ulp (error of the cubic step ~ seed^3 = 2^-66).
D3_HD float frsq(float a) { return rsqrtf(a); }
D3_HD float frcp(float a) { return 1.0f / a; }
D3_HD float fex(float a) { return expf(a); }

D3_HD double frsq(double a) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double t = a * y;
  double e = fma(-t, y, 1.0);
  double p = fma(0.375, e, 0.5);
  double q = e * y;
  return fma(q, p, y);
#else
  return 1.0 / sqrt(a);
#endif
}

D3_HD double frcp(double a) {
#ifdef __CUDA_ARCH__
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));
  double e = fma(-a, y, 1.0);
  double f = fma(e, e, e);
  return fma(y, f, y);
#else
  return 1.0 / a;
#endif
}

// exp(x) for -700 <= x <= 700 (callers clamp).  The Taylor coefficients live in
// constant memory so that each DFMA takes its coefficient as a c[][] operand
// (as literals they cost two UMOV each, 4 % of all issue slots in profile r1c).
static __device__ __constant__ double kExpC[12] = {
    2.505210838544172e-08,   // 1/11!
    2.755731922398589e-07,   // 1/10!
    2.7557319223985893e-06,  // 1/9!
    2.48015873015873e-05,    // 1/8!
    1.984126984126984e-04,   // 1/7!
    1.388888888888889e-03,   // 1/6!
    8.333333333333333e-03,   // 1/5!
    4.1666666666666664e-02,  // 1/4!
    1.6666666666666666e-01,  // 1/3!
    0.5, 1.0, 1.0};
static __device__ __constant__ double kExpR[4] = {
    1.4426950408889634,            // log2(e)
    6755399441055744.0,            // 1.5 * 2^52
    -6.93147180369123816490e-01,   // -ln2 (high part)
    -1.90821492927058770002e-10};  // -ln2 (low part)

D3_HD double fex(double x) {
#ifdef __CUDA_ARCH__
  double t = fma(x, kExpR[0], kExpR[1]);
  int n = __double2loint(t);
  double nf = t - kExpR[1];
  double r = fma(nf, kExpR[2], x);
  r = fma(nf, kExpR[3], r);
  double p = kExpC[0];
#pragma unroll
  for (int k = 1; k < 12; ++k) p = fma(p, r, kExpC[k]);
  return __hiloint2double(__double2hiint(p) + (n << 20), __double2loint(p));
#else
  return exp(x);
#endif
}

template <typename T> D3_HD Dual<T> frcp(Dual<T> a) {
  T r = frcp(a.v);
  return Dual<T>(r, -a.d * r * r);
}
template <typename T> D3_HD Dual<T> frsq(Dual<T> a) {
  T r = frsq(a.v);
  return Dual<T>(r, T(-0.5) * a.d * r * r * r);
}
template <typename T> D3_HD Dual<T> fex(Dual<T> a) {
  T e = fex(a.v);
  return Dual<T>(e, e * a.d);
}

// 1/c for the dC6/C6 factors of the three-body backward: a pair whose C6 is exactly 0 (zero block of a
// custom table, underflow in float) has no C6 derivative in the reference (eps clamp before the sqrt).
template <typename S> D3_HD S rcp0(S c) { return val(c) != val(S(0)) ? frcp(c) : S(0); }

// max(a, lo) on the value (tangent follows the selected branch)
D3_HD float fmx(float a, float lo) { return a < lo ? lo : a; }
D3_HD double fmx(double a, double lo) { return a < lo ? lo : a; }

// smallest addend that keeps 1/sqrt(r^2) finite for coincident atoms
template <typename T> D3_HD T tiny_of();
template <> D3_HD float tiny_of<float>() { return 1e-30f; }
template <> D3_HD double tiny_of<double>() { return 1e-300; }
template <typename T> D3_HD Dual<T> fmx(Dual<T> a, T lo) { return a.v < lo ? Dual<T>(lo) : a; }

// conversions between working precision and the (float64) weight stage
D3_HD double widen(float a) { return (double)a; }
D3_HD double widen(double a) { return a; }
D3_HD Dual<double> widen(Dual<float> a) { return Dual<double>((double)a.v, (double)a.d); }
D3_HD Dual<double> widen(Dual<double> a) { return a; }

template <typename S> D3_HD S narrow(typename Traits<S>::wide a);
template <> D3_HD float narrow<float>(double a) { return (float)a; }
template <> D3_HD double narrow<double>(double a) { return a; }
template <> D3_HD Dual<float> narrow<Dual<float>>(Dual<double> a) { return Dual<float>((float)a.v, (float)a.d); }
template <> D3_HD Dual<double> narrow<Dual<double>>(Dual<double> a) { return a; }

// integer power by repeated squaring is written out at the call sites.

}  // namespace d3
