// rt_device.cuh — device restatement of the per-ray functions of the reference, templated on the
// arithmetic type T (float: production path-tracing arithmetic; double: bit-faithful path of the
// deterministic renderers, compiled without multiply-add fusion).
//
//   closest hit  World.ray_intersection   world.py:51-69     -> scan_*/finish_hit
//   any hit      World.is_point_visible   world.py:71-80     -> any_hit_*
//   shapes       Sphere / Plane           shapes.py:97-198
//   pigments     Pigment.get_color        materials.py:58,70-82,96-100
//   BRDFs        eval / scatter_ray       materials.py:129-196, ONB geometry.py:247-262
//   cameras      fire_ray                 camera.py:59-78,103-124; imagetracer.py:48-58
#pragma once
#include "rt_math.cuh"
#include "rt_pcg.cuh"
#include "../../include/rt_api.h"

// ---------------------------------------------------------------- device-side scene tables
struct DevPigment {
  int32_t kind, steps, tex_w, tex_h;
  cudaTextureObject_t tex;  // float4 texels, point filter (fp32 path)
  const double* texels64;   // same texels in fp64 (fp64 path)
  float c1[3], c2[3];
  double c1d[3], c2d[3];
};
struct DevMaterial {
  int32_t brdf_kind, brdf_pigment, emitted_pigment;
  int32_t uses_uv;  // 0 when both pigments are uniform: the hit record can skip atan2 / acos
  double threshold;
};
struct DevLight {
  double pos[3], color[3], radius;
};

// Shapes are stored sorted by kind: spheres [0, n_spheres) then planes [n_spheres, n_shapes), each
// group in World.shapes order; `orig` maps back to the index in World.shapes.  Within a group a
// strict '<' scan keeps the first shape on ties like world.py:62; across groups ties are resolved
// on `orig`, so the winner is the one the reference's single loop would pick.
template <typename T> struct SceneView {
  const T* invm;  // [n][12] inverse transforms, the only per-shape data the scan loops read
  const T* m;     // [n][12]
  const int32_t* orig;
  const int32_t* material;  // by sorted index
  const DevMaterial* materials;
  const DevPigment* pigments;
  const DevLight* lights;
  int32_t n_shapes, n_spheres, n_lights;
};

template <typename T> struct Hit {
  int32_t idx;   // sorted index, -1 = miss
  T t;
  V3<T> point, normal;  // world space, normal normalised
  T u, v;
};

// ---------------------------------------------------------------- shape tests (t only)
// Sphere.ray_intersection shapes.py:97-121 up to the choice of the root; returns t, +inf = miss
// (a valid t is always < tmax <= inf, so the closest-hit scan needs a single compare).
template <typename T> RT_DEV T sphere_t(const T* im, const Ray<T>& r) {
  V3<T> o = xf_point(im, r.o);
  V3<T> d = xf_vec(im, r.d);
  T a = dot(d, d);
  T b = (T)2 * dot(o, d);
  T c = dot(o, o) - (T)1;
  T delta = b * b - (T)4 * a * c;
  if (delta <= (T)0) return Num<T>::inf();
  T sd = Num<T>::sqrt(delta);
  T den = (T)2 * a;
  T t1 = Num<T>::div(-b - sd, den);
  T t2 = Num<T>::div(-b + sd, den);
  if (t1 > r.tmin && t1 < r.tmax) return t1;
  if (t2 > r.tmin && t2 < r.tmax) return t2;
  return Num<T>::inf();
}

// Plane.ray_intersection shapes.py:163-174; only row 2 of the inverse transform is needed for t.
template <typename T> RT_DEV T plane_t(const T* im, const Ray<T>& r) {
  T oz = r.o.x * im[8] + r.o.y * im[9] + r.o.z * im[10] + im[11];
  T dz = r.d.x * im[8] + r.d.y * im[9] + r.d.z * im[10];
  if (Num<T>::abs(dz) < (T)1e-5) return Num<T>::inf();
  T t = Num<T>::div(-oz, dz);
  if (t <= r.tmin || t >= r.tmax) return Num<T>::inf();
  return t;
}

// world.py:62 keeps the FIRST shape of World.shapes on equal t; spheres are scanned before planes
// here, so a plane that ties with the current best sphere wins iff it came first in World.shapes.
// Out of line: ties are rare and the two global loads must not be hoisted into the scan loop.
static __device__ __noinline__ bool plane_wins_tie(const int32_t* __restrict__ orig, int plane, int best) {
  return orig[plane] < orig[best];
}

template <typename T>
RT_DEV void scan_planes(const T* __restrict__ xf, int begin, int end, int n_spheres,
                        const int32_t* __restrict__ orig, const Ray<T>& r, T& best_t, int& best) {
  for (int i = max(begin, n_spheres); i < end; ++i) {
    T t = plane_t(xf + 12 * (i - begin), r);
    if (t < best_t) { best_t = t; best = i; }
    else if (t == best_t && best >= 0 && best < n_spheres) {
      if (plane_wins_tie(orig, i, best)) { best_t = t; best = i; }
    }
  }
}

// Closest-hit scan over sorted shapes [begin, end) held at `xf` (xf[0] is shape `begin`); the data
// may live in shared or global memory.  All lanes walk the same shapes: loads are broadcasts.
// Start with best_t = +inf, best = -1.
template <typename T>
RT_DEV void scan_closest(const T* __restrict__ xf, int begin, int end, int n_spheres,
                         const int32_t* __restrict__ orig, const Ray<T>& r, T& best_t, int& best) {
  int s_end = min(end, n_spheres);
#pragma unroll 2
  for (int i = begin; i < s_end; ++i) {
    T t = sphere_t(xf + 12 * (i - begin), r);
    if (t < best_t) { best_t = t; best = i; }
  }
  scan_planes<T>(xf, begin, end, n_spheres, orig, r, best_t, best);
}

// shapes.py:133-151 / :191-198 — true as soon as one shape blocks the segment
template <typename T>
RT_DEV bool scan_any(const T* __restrict__ xf, int begin, int end, int n_spheres, const Ray<T>& r) {
  int s_end = min(end, n_spheres);
  for (int i = begin; i < s_end; ++i) {
    const T* im = xf + 12 * (i - begin);
    V3<T> o = xf_point(im, r.o);
    V3<T> d = xf_vec(im, r.d);
    T a = dot(d, d);
    T b = (T)2 * dot(o, d);
    T c = dot(o, o) - (T)1;
    T delta = b * b - (T)4 * a * c;
    if (delta <= (T)0) continue;
    T sd = Num<T>::sqrt(delta);
    T den = (T)2 * a;
    T t1 = Num<T>::div(-b - sd, den);
    T t2 = Num<T>::div(-b + sd, den);
    if ((r.tmin < t1 && t1 < r.tmax) || (r.tmin < t2 && t2 < r.tmax)) return true;
  }
  for (int i = max(begin, n_spheres); i < end; ++i) {
    const T* im = xf + 12 * (i - begin);
    T oz = r.o.x * im[8] + r.o.y * im[9] + r.o.z * im[10] + im[11];
    T dz = r.d.x * im[8] + r.d.y * im[9] + r.d.z * im[10];
    if (Num<T>::abs(dz) < (T)1e-5) continue;
    T t = Num<T>::div(-oz, dz);
    if (r.tmin < t && t < r.tmax) return true;
  }
  return false;
}

// ---------------------------------------------------------------- fp32 production scans
// Same mathematics as sphere_t<float> (shapes.py:97-121), arranged for the FMA pipe: the ray is
// taken to the sphere's frame with 18 FFMA, a, b/2 and c cost 9 more, delta/4 = (b/2)^2 - a c two:
// 29 FMA-pipe instructions and three broadcast LDS.128 per sphere, nothing else in the common case.
// Spheres are tested four at a time; the few whose line is crossed (delta > 0) go to a per-lane
// candidate list and only those get the sqrt / reciprocal / range tests after the sweep — so the
// sweep itself has one (rarely taken) branch per four spheres.
RT_DEV float sphere_qdelta(const float* __restrict__ im, const Ray<float>& r, float& a, float& hb) {
  const float4 r0 = reinterpret_cast<const float4*>(im)[0];
  const float4 r1 = reinterpret_cast<const float4*>(im)[1];
  const float4 r2 = reinterpret_cast<const float4*>(im)[2];
  const float px = fmaf(r0.x, r.o.x, fmaf(r0.y, r.o.y, fmaf(r0.z, r.o.z, r0.w)));
  const float py = fmaf(r1.x, r.o.x, fmaf(r1.y, r.o.y, fmaf(r1.z, r.o.z, r1.w)));
  const float pz = fmaf(r2.x, r.o.x, fmaf(r2.y, r.o.y, fmaf(r2.z, r.o.z, r2.w)));
  const float dx = fmaf(r0.x, r.d.x, fmaf(r0.y, r.d.y, r0.z * r.d.z));
  const float dy = fmaf(r1.x, r.d.x, fmaf(r1.y, r.d.y, r1.z * r.d.z));
  const float dz = fmaf(r2.x, r.d.x, fmaf(r2.y, r.d.y, r2.z * r.d.z));
  a = fmaf(dx, dx, fmaf(dy, dy, dz * dz));
  hb = fmaf(px, dx, fmaf(py, dy, pz * dz));
  const float c = fmaf(px, px, fmaf(py, py, fmaf(pz, pz, -1.0f)));
  return fmaf(hb, hb, -a * c);
}

// roots of the crossed sphere: t = (-b/2 -+ sqrt(delta/4)) / a, first one inside (tmin, tmax)
RT_DEV float sphere_root(float a, float hb, float qd, float tmin, float tmax) {
  const float sd = fast_sqrt(qd), inv = fast_rcp(a);
  const float t1 = (-hb - sd) * inv, t2 = (-hb + sd) * inv;
  if (t1 > tmin && t1 < tmax) return t1;
  if (t2 > tmin && t2 < tmax) return t2;
  return Num<float>::inf();
}

#define RT_CAND_CAP 16

// Sweeps spheres [i0, i1) (data at xf, xf[0] = sphere `base`); returns the number of crossed spheres,
// the first RT_CAND_CAP of them in cand[] in ascending order.
RT_DEV int sweep_spheres(const float* __restrict__ xf, int base, int i0, int i1, const Ray<float>& r, int* cand) {
  int nc = 0;
  int i = i0;
  for (; i + 4 <= i1; i += 4) {
    float a, hb;
    const float* q = xf + 12 * (i - base);
    const float d0 = sphere_qdelta(q, r, a, hb);
    const float d1 = sphere_qdelta(q + 12, r, a, hb);
    const float d2 = sphere_qdelta(q + 24, r, a, hb);
    const float d3 = sphere_qdelta(q + 36, r, a, hb);
    if (fmaxf(fmaxf(d0, d1), fmaxf(d2, d3)) > 0.0f) {
      if (d0 > 0.0f) { cand[nc & (RT_CAND_CAP - 1)] = i; ++nc; }
      if (d1 > 0.0f) { cand[nc & (RT_CAND_CAP - 1)] = i + 1; ++nc; }
      if (d2 > 0.0f) { cand[nc & (RT_CAND_CAP - 1)] = i + 2; ++nc; }
      if (d3 > 0.0f) { cand[nc & (RT_CAND_CAP - 1)] = i + 3; ++nc; }
    }
  }
  for (; i < i1; ++i) {
    float a, hb;
    if (sphere_qdelta(xf + 12 * (i - base), r, a, hb) > 0.0f) { cand[nc & (RT_CAND_CAP - 1)] = i; ++nc; }
  }
  return nc;
}

template <>
RT_DEV void scan_closest<float>(const float* __restrict__ xf, int begin, int end, int n_spheres,
                                const int32_t* __restrict__ orig, const Ray<float>& r, float& best_t, int& best) {
  const int s_end = min(end, n_spheres);
  if (begin < s_end) {
    int cand[RT_CAND_CAP];
    const int nc = sweep_spheres(xf, begin, begin, s_end, r, cand);
    if (nc <= RT_CAND_CAP) {
      for (int j = 0; j < nc; ++j) {  // ascending index: strict '<' keeps the first shape on ties
        const int i = cand[j];
        float a, hb;
        const float qd = sphere_qdelta(xf + 12 * (i - begin), r, a, hb);
        const float t = sphere_root(a, hb, qd, r.tmin, r.tmax);
        if (t < best_t) { best_t = t; best = i; }
      }
    } else {  // a line through more than RT_CAND_CAP spheres: plain pass
      for (int i = begin; i < s_end; ++i) {
        float a, hb;
        const float qd = sphere_qdelta(xf + 12 * (i - begin), r, a, hb);
        if (qd > 0.0f) {
          const float t = sphere_root(a, hb, qd, r.tmin, r.tmax);
          if (t < best_t) { best_t = t; best = i; }
        }
      }
    }
  }
  scan_planes<float>(xf, begin, end, n_spheres, orig, r, best_t, best);
}

template <>
RT_DEV bool scan_any<float>(const float* __restrict__ xf, int begin, int end, int n_spheres, const Ray<float>& r) {
  const int s_end = min(end, n_spheres);
  // planes first: a handful of shapes that often decide the query (ground, sky)
  for (int i = max(begin, n_spheres); i < end; ++i)
    if (plane_t<float>(xf + 12 * (i - begin), r) < Num<float>::inf()) return true;
  if (begin < s_end) {
    int cand[RT_CAND_CAP];
    const int nc = sweep_spheres(xf, begin, begin, s_end, r, cand);
    if (nc <= RT_CAND_CAP) {
      for (int j = 0; j < nc; ++j) {
        float a, hb;
        const float qd = sphere_qdelta(xf + 12 * (cand[j] - begin), r, a, hb);
        const float sd = fast_sqrt(qd), inv = fast_rcp(a);
        const float t1 = (-hb - sd) * inv, t2 = (-hb + sd) * inv;
        if ((r.tmin < t1 && t1 < r.tmax) || (r.tmin < t2 && t2 < r.tmax)) return true;
      }
    } else {
      for (int i = begin; i < s_end; ++i) {
        float a, hb;
        const float qd = sphere_qdelta(xf + 12 * (i - begin), r, a, hb);
        if (qd > 0.0f) {
          const float sd = fast_sqrt(qd), inv = fast_rcp(a);
          const float t1 = (-hb - sd) * inv, t2 = (-hb + sd) * inv;
          if ((r.tmin < t1 && t1 < r.tmax) || (r.tmin < t2 && t2 < r.tmax)) return true;
        }
      }
    }
  }
  return false;
}

// Hit record of the winning shape (shapes.py:123-131 / :176-189) + world.py:66-67
template <typename T>
RT_DEV void finish_hit(const SceneView<T>& sc, const Ray<T>& r, T t, int idx, Hit<T>& h, bool normalise = true,
                       bool force_uv = false) {
  const T* im = sc.invm + 12 * (size_t)idx;
  const T* mm = sc.m + 12 * (size_t)idx;
  V3<T> o = xf_point(im, r.o);
  V3<T> d = xf_vec(im, r.d);
  V3<T> hp = o + t * d;
  h.idx = idx;
  h.t = t;
  h.point = xf_point(mm, hp);
  V3<T> n;
  h.u = h.v = (T)0;
  if (idx < sc.n_spheres) {
    n = (dot(hp, d) < (T)0) ? hp : -hp;  // shapes.py:45-54
    if (force_uv || sc.materials[sc.material[idx]].uses_uv) {
      T u = Num<T>::atan2(hp.y, hp.x) / (T)(2.0 * 3.14159265358979323846);  // shapes.py:36-42
      h.u = (u >= (T)0) ? u : u + (T)1;
      T z = Num<T>::min((T)1, Num<T>::max((T)-1, hp.z));  // the reference would raise outside [-1,1]
      h.v = Num<T>::acos(z) / (T)3.14159265358979323846;
    }
  } else {
    n = mk3<T>((T)0, (T)0, (d.z < (T)0) ? (T)1 : (T)-1);
    h.u = hp.x - Num<T>::floor(hp.x);
    h.v = hp.y - Num<T>::floor(hp.y);
  }
  n = xf_normal(im, n);
  if (normalise) {
    if (Num<T>::is_f64) {  // Normal.normalize geometry.py:220-226
      T nn = Num<T>::sqrt(n.x * n.x + n.y * n.y + n.z * n.z);
      n = mk3<T>(n.x / nn, n.y / nn, n.z / nn);
    } else {
      n = normalize(n);
    }
  }
  h.normal = n;
}

// ---------------------------------------------------------------- pigments
template <typename T> RT_DEV V3<T> pig_c1(const DevPigment& p);
template <> RT_DEV V3<float> pig_c1<float>(const DevPigment& p) { return mk3<float>(p.c1[0], p.c1[1], p.c1[2]); }
template <> RT_DEV V3<double> pig_c1<double>(const DevPigment& p) { return mk3<double>(p.c1d[0], p.c1d[1], p.c1d[2]); }
template <typename T> RT_DEV V3<T> pig_c2(const DevPigment& p);
template <> RT_DEV V3<float> pig_c2<float>(const DevPigment& p) { return mk3<float>(p.c2[0], p.c2[1], p.c2[2]); }
template <> RT_DEV V3<double> pig_c2<double>(const DevPigment& p) { return mk3<double>(p.c2d[0], p.c2d[1], p.c2d[2]); }

template <typename T> RT_DEV V3<T> pig_texel(const DevPigment& p, int col, int row);
// out of line: image pigments are the rare case and the fetch must not be if-converted into the
// uniform / checkered paths
static __device__ __noinline__ float4 fetch_texel(cudaTextureObject_t tex, float x, float y) {
  return tex2D<float4>(tex, x, y);  // texture unit, nearest texel
}
template <> RT_DEV V3<float> pig_texel<float>(const DevPigment& p, int col, int row) {
  float4 t = fetch_texel(p.tex, col + 0.5f, row + 0.5f);
  return mk3<float>(t.x, t.y, t.z);
}
template <> RT_DEV V3<double> pig_texel<double>(const DevPigment& p, int col, int row) {
  const double* t = p.texels64 + 3 * ((size_t)row * p.tex_w + col);
  return mk3<double>(t[0], t[1], t[2]);
}

template <typename T> RT_DEV V3<T> pigment_color(const DevPigment* pigs, int idx, T u, T v) {
  const DevPigment& p = pigs[idx];
  if (p.kind == RT_PIGMENT_UNIFORM) return pig_c1<T>(p);  // materials.py:58
  if (p.kind == RT_PIGMENT_CHECKERED) {                   // materials.py:96-100
    long long iu = Num<T>::floor_ll(u * (T)p.steps);
    long long iv = Num<T>::floor_ll(v * (T)p.steps);
    return ((iu & 1) == (iv & 1)) ? pig_c1<T>(p) : pig_c2<T>(p);
  }
  int col = (int)(u * (T)p.tex_w);                        // materials.py:70-82
  int row = (int)(v * (T)p.tex_h);
  col = min(max(col, 0), p.tex_w - 1);
  row = min(max(row, 0), p.tex_h - 1);
  return pig_texel<T>(p, col, row);
}

// ---------------------------------------------------------------- BRDFs
template <typename T> RT_DEV T normalized_dot(V3<T> a, V3<T> b) {  // geometry.py:265-276
  return dot(normalize(a), normalize(b));
}
template <typename T> RT_DEV T clamped_acos(T x) {
  return Num<T>::acos(Num<T>::min((T)1, Num<T>::max((T)-1, x)));
}

// BRDF.eval: DiffuseBRDF materials.py:129-130, SpecularBRDF materials.py:164-173
template <typename T>
RT_DEV V3<T> brdf_eval(const SceneView<T>& sc, const DevMaterial& mat, V3<T> normal, V3<T> in_dir,
                       V3<T> out_dir, T u, T v) {
  V3<T> c = pigment_color<T>(sc.pigments, mat.brdf_pigment, u, v);
  if (mat.brdf_kind == RT_BRDF_DIFFUSE) return (T)(1.0 / 3.14159265358979323846) * c;
  T theta_in = clamped_acos(normalized_dot(normal, in_dir));
  T theta_out = clamped_acos(normalized_dot(normal, out_dir));
  if (Num<T>::abs(theta_in - theta_out) < (T)mat.threshold) return c;
  return mk3<T>((T)0, (T)0, (T)0);
}

// create_onb_from_z geometry.py:247-262 (Duff et al.)
template <typename T> RT_DEV void onb_from_z(V3<T> n, V3<T>& e1, V3<T>& e2) {
  T sign = (n.z > (T)0) ? (T)1 : (T)-1;
  T a = Num<T>::div((T)-1, sign + n.z);
  T b = n.x * n.y * a;
  e1 = mk3<T>((T)1 + sign * n.x * n.x * a, sign * b, -sign * n.x);
  e2 = mk3<T>(b, sign + n.y * n.y * a, -n.y);
}

// DiffuseBRDF.scatter_ray materials.py:132-152: cosine-weighted direction around `normal`
template <typename T> RT_DEV V3<T> diffuse_dir(V3<T> normal, T u1, T u2) {
  V3<T> e1, e2;
  onb_from_z(normal, e1, e2);
  T cos_theta = Num<T>::sqrt(u1), sin_theta = Num<T>::sqrt((T)1 - u1);
  T phi = (T)(2.0 * 3.14159265358979323846) * u2;
  T s, c;
  Num<T>::sincos(phi, &s, &c);
  return (cos_theta * (c * e1) + cos_theta * (s * e2)) + sin_theta * normal;
}

// SpecularBRDF.scatter_ray materials.py:175-196: mirror direction
template <typename T> RT_DEV V3<T> specular_dir(V3<T> incoming, V3<T> normal) {
  V3<T> rd = normalize(incoming);
  V3<T> n = normalize(normal);
  T dp = dot(n, rd);
  return rd - dp * ((T)2 * n);
}

// ---------------------------------------------------------------- camera + image tracer
// Always evaluated in fp64 and then rounded to T: the primary rays are the reference's own rays
// (camera.py:59-78,103-124 after imagetracer.py:56-58), whatever T the tracing uses.
struct DevCamera {
  int32_t kind, _pad;
  double dist, aspect;
  double m[12];
};

// Camera.fire_ray(u, v), camera.py:59-78 / :103-124
RT_DEV void camera_fire_f64(const DevCamera& cam, double u, double v, V3<double>& o, V3<double>& d) {
  double sy = __dmul_rn(__dsub_rn(1.0, __dmul_rn(2.0, u)), cam.aspect);
  double sz = __dsub_rn(__dmul_rn(2.0, v), 1.0);
  V3<double> lo, ld;
  if (cam.kind == RT_CAMERA_PERSPECTIVE) {
    lo = mk3<double>(-cam.dist, 0.0, 0.0);
    ld = mk3<double>(cam.dist, sy, sz);
  } else {
    lo = mk3<double>(-1.0, sy, sz);
    ld = mk3<double>(1.0, 0.0, 0.0);
  }
  const double* m = cam.m;
  // explicit round-to-nearest ops: no fused multiply-add even in the fp32 translation unit
#define RT_ROW3(v, r) __dadd_rn(__dadd_rn(__dmul_rn(v.x, m[r]), __dmul_rn(v.y, m[r + 1])), __dmul_rn(v.z, m[r + 2]))
  o = mk3<double>(__dadd_rn(RT_ROW3(lo, 0), m[3]), __dadd_rn(RT_ROW3(lo, 4), m[7]), __dadd_rn(RT_ROW3(lo, 8), m[11]));
  d = mk3<double>(RT_ROW3(ld, 0), RT_ROW3(ld, 4), RT_ROW3(ld, 8));
#undef RT_ROW3
}

// ImageTracer.fire_ray, imagetracer.py:48-58
RT_DEV void camera_ray_f64(const DevCamera& cam, int width, int height, int col, int row,
                           double u_pixel, double v_pixel, V3<double>& o, V3<double>& d) {
  double u = __ddiv_rn(__dadd_rn((double)col, u_pixel), (double)width);
  double v = __dsub_rn(1.0, __ddiv_rn(__dadd_rn((double)row, v_pixel), (double)height));
  camera_fire_f64(cam, u, v, o, d);
}

// jitter of imagetracer.py:88-93 from the two draws of this sample
RT_DEV void jitter_f64(Pcg& aa, int ic, int ir, int S, double& u_pixel, double& v_pixel) {
  double r1 = pcg_random_float<double>(aa);
  double r2 = pcg_random_float<double>(aa);
  u_pixel = __ddiv_rn(__dadd_rn((double)ic, r1), (double)S);
  v_pixel = __ddiv_rn(__dadd_rn((double)ir, r2), (double)S);
}
