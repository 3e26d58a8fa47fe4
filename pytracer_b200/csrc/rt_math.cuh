// rt_math.cuh — scalar/vector helpers shared by the fp32 and fp64 instantiations.
//
// Every expression is written in the operation order of the reference's Python
// (geometry.py / transformations.py), so that the fp64 instantiation — compiled with
// --fmad=false — rounds exactly like CPython does, and the fp32 instantiation lets the
// compiler fuse multiply-adds.
#pragma once
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#define RT_DEV __device__ __forceinline__

template <typename T> struct Num;

// single-instruction MUFU forms (approx, flush-to-zero): 1-2 ulp, no slow-path call, no range fix-up
RT_DEV float fast_rcp(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
RT_DEV float fast_sqrt(float x) {
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}

template <> struct Num<float> {
  static constexpr bool is_f64 = false;
  static RT_DEV float sqrt(float x) { return fast_sqrt(x); }
  static RT_DEV float rsqrt(float x) { return rsqrtf(x); }
  static RT_DEV float div(float a, float b) { return a * fast_rcp(b); }
  static RT_DEV float floor(float x) { return floorf(x); }
  static RT_DEV float abs(float x) { return fabsf(x); }
  static RT_DEV float max(float a, float b) { return fmaxf(a, b); }
  static RT_DEV float min(float a, float b) { return fminf(a, b); }
  static RT_DEV float acos(float x) { return acosf(x); }
  static RT_DEV float atan2(float y, float x) { return atan2f(y, x); }
  // cos/sin of phi in [0, 2pi]: shift into [-pi, pi] where the MUFU approximations are at
  // their best (abs. error 2^-21), cos(phi) = -cos(phi - pi), sin(phi) = -sin(phi - pi).
  static RT_DEV void sincos(float phi, float* s, float* c) {
    float x = phi - 3.14159265358979323846f;
    *s = -__sinf(x);
    *c = -__cosf(x);
  }
  static RT_DEV float inf() { return __int_as_float(0x7f800000); }
  static RT_DEV long long floor_ll(float x) { return (long long)__float2int_rd(x); }  // parity only
};

template <> struct Num<double> {
  static constexpr bool is_f64 = true;
  static RT_DEV double sqrt(double x) { return ::sqrt(x); }
  static RT_DEV double rsqrt(double x) { return 1.0 / ::sqrt(x); }
  static RT_DEV double div(double a, double b) { return a / b; }
  static RT_DEV double floor(double x) { return ::floor(x); }
  static RT_DEV double abs(double x) { return ::fabs(x); }
  static RT_DEV double max(double a, double b) { return ::fmax(a, b); }
  static RT_DEV double min(double a, double b) { return ::fmin(a, b); }
  static RT_DEV double acos(double x) { return ::acos(x); }
  static RT_DEV double atan2(double y, double x) { return ::atan2(y, x); }
  static RT_DEV void sincos(double phi, double* s, double* c) { ::sincos(phi, s, c); }
  static RT_DEV double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
  static RT_DEV long long floor_ll(double x) { return (long long)::floor(x); }
};

template <typename T> struct V3 {
  T x, y, z;
};

template <typename T> RT_DEV V3<T> mk3(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> RT_DEV V3<T> operator+(V3<T> a, V3<T> b) { return mk3<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> RT_DEV V3<T> operator-(V3<T> a, V3<T> b) { return mk3<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> RT_DEV V3<T> operator-(V3<T> a) { return mk3<T>(-a.x, -a.y, -a.z); }
template <typename T> RT_DEV V3<T> operator*(T s, V3<T> a) { return mk3<T>(s * a.x, s * a.y, s * a.z); }
template <typename T> RT_DEV V3<T> mul3(V3<T> a, V3<T> b) { return mk3<T>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <typename T> RT_DEV T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
template <typename T> RT_DEV T max3(V3<T> a) { return Num<T>::max(Num<T>::max(a.x, a.y), a.z); }
template <typename T> RT_DEV V3<T> normalize(V3<T> a) {  // Vec.normalize, geometry.py:129-135
  if (Num<T>::is_f64) {
    T n = Num<T>::sqrt(dot(a, a));
    return mk3<T>(a.x / n, a.y / n, a.z / n);
  }
  T r = Num<T>::rsqrt(dot(a, a));
  return r * a;
}
template <typename T, typename U> RT_DEV V3<T> cast3(V3<U> a) { return mk3<T>((T)a.x, (T)a.y, (T)a.z); }

// Transformation * Point / Vec / Normal on a row-major 3x4 block (transformations.py:58-86)
template <typename T> RT_DEV V3<T> xf_point(const T* m, V3<T> p) {
  return mk3<T>(p.x * m[0] + p.y * m[1] + p.z * m[2] + m[3],
                p.x * m[4] + p.y * m[5] + p.z * m[6] + m[7],
                p.x * m[8] + p.y * m[9] + p.z * m[10] + m[11]);
}
template <typename T> RT_DEV V3<T> xf_vec(const T* m, V3<T> v) {
  return mk3<T>(v.x * m[0] + v.y * m[1] + v.z * m[2],
                v.x * m[4] + v.y * m[5] + v.z * m[6],
                v.x * m[8] + v.y * m[9] + v.z * m[10]);
}
template <typename T> RT_DEV V3<T> xf_normal(const T* invm, V3<T> n) {  // transpose of the inverse
  return mk3<T>(n.x * invm[0] + n.y * invm[4] + n.z * invm[8],
                n.x * invm[1] + n.y * invm[5] + n.z * invm[9],
                n.x * invm[2] + n.y * invm[6] + n.z * invm[10]);
}

template <typename T> struct Ray {
  V3<T> o, d;
  T tmin, tmax;
};
