// rt_bvh.cuh — bounding-volume hierarchy over the spheres of a scene: an explicit,
// separately reported mode (rt_render_params.accel = RT_ACCEL_BVH).  The default path evaluates
// World.ray_intersection's loop over ALL shapes like the reference (world.py:55-64) and sits on the FP32
// roof; this one skips the spheres whose (padded) box the ray misses, which changes the roofline class —
// memory latency and divergence instead of FMA throughput — and the cost per ray from O(N) to O(log N).
//
// Results are those of the linear scan — bit for bit in fp64, and in fp32 wherever the fp32 quadratic
// is meaningful (see the note at the end of this comment): a sphere that survives the culling is tested by
// the very same function on the very same table (sphere_t_at / sphere_t<double> / sphere_blocks on
// invm), ties on t are resolved towards the lower index like the scan's strict '<' in ascending order,
// and the culling is conservative — boxes are computed in fp64 and padded (host, rt_bvh.h) far beyond
// the rounding of the fp32 slab test, so no sphere the scan would hit is ever skipped.  Planes are
// unbounded: they stay outside the tree and are scanned like before.
// fp32 note: for a ray that starts thousands of units from a sphere (ground points near the horizon
// shooting shadow or scatter rays back at the scene) |o'|^2 ~ 1e8 swamps the radius 1 and the sign of the
// fp32 discriminant is rounding noise: the linear scan then reports phantom hits on spheres the ray only
// passes at a distance.  The tree does not test spheres whose box the ray misses, so it is free of those —
// the two fp32 images differ on ~1e-4 of the pixels of such frames, and the tree is the faithful one.
//
// Node = 64 bytes = four 16-byte loads: the boxes of BOTH children and where they lead, so one fetch
// decides two subtrees (the usual layout for GPU traversal):
//   q0 = lo0.xyz, hi0.x   q1 = hi0.yz, lo1.xy   q2 = lo1.z, hi1.xyz   q3 = {ref0, ref1, -, -} as ints
// (child references: see RT_BVH_DONE below; a leaf's spheres are prims[first .. first + count)).
#pragma once
#include <climits>

#include "rt_device.cuh"

#define RT_BVH_STACK 48  // rt_api.cu refuses trees deeper than this minus 2 (rt_bvh.h keeps them far shallower)

struct BvhRay {
  float ox, oy, oz, ix, iy, iz;  // origin, 1 / direction (zeros replaced by a tiny value of the same sign)
};
RT_DEV float bvh_safe_inv(float d) {
  const float a = fabsf(d) < 1e-30f ? copysignf(1e-30f, d) : d;
  return 1.0f / a;
}
template <typename T> RT_DEV BvhRay bvh_ray(const Ray<T>& r) {
  BvhRay b;
  b.ox = (float)r.o.x; b.oy = (float)r.o.y; b.oz = (float)r.o.z;
  b.ix = bvh_safe_inv((float)r.d.x); b.iy = bvh_safe_inv((float)r.d.y); b.iz = bvh_safe_inv((float)r.d.z);
  return b;
}
// entry / exit parameter of the ray through a box (slab test); hit iff near <= far
RT_DEV void bvh_slab(const BvhRay& r, float lx, float ly, float lz, float hx, float hy, float hz, float& tn, float& tf) {
  const float ax = (lx - r.ox) * r.ix, bx = (hx - r.ox) * r.ix;
  const float ay = (ly - r.oy) * r.iy, by = (hy - r.oy) * r.iy;
  const float az = (lz - r.oz) * r.iz, bz = (hz - r.oz) * r.iz;
  tn = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fminf(az, bz));
  tf = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
}

// conservative comparisons: lower bounds are pushed down, upper bounds up, by a relative 1e-6
RT_DEV float bvh_dn(float x) { return x - fabsf(x) * 1e-6f; }
RT_DEV float bvh_up(float x) { return x + fabsf(x) * 1e-6f; }
RT_DEV bool bvh_box_hit(float tn, float tf, float tmin, float limit) {
  return bvh_dn(tn) <= bvh_up(tf) && bvh_up(tf) >= tmin && bvh_dn(tn) <= limit;
}

template <typename T> RT_DEV T bvh_sphere_t(const SceneView<T>& sc, int i, const Ray<T>& r, int origin);
// fp32: the function the linear path resolves its candidates with (sphere_t_at, rt_device.cuh)
template <> RT_DEV float bvh_sphere_t<float>(const SceneView<float>& sc, int i, const Ray<float>& r, int origin) {
  return sphere_t_at(sc.invm + 12 * (size_t)i, r, i == origin);
}
template <> RT_DEV double bvh_sphere_t<double>(const SceneView<double>& sc, int i, const Ray<double>& r, int) {
  return sphere_t<double>(sc.invm + 12 * (size_t)i, r);
}
template <typename T> RT_DEV bool bvh_sphere_blocks(const SceneView<T>& sc, int i, const Ray<T>& r);
template <> RT_DEV bool bvh_sphere_blocks<float>(const SceneView<float>& sc, int i, const Ray<float>& r) {
  return sphere_blocks(sc.invm + 12 * (size_t)i, r);
}
template <> RT_DEV bool bvh_sphere_blocks<double>(const SceneView<double>& sc, int i, const Ray<double>& r) {
  return scan_any<double>(sc.invm + 12 * (size_t)i, i, i + 1, sc.n_spheres, r);  // the scan's own test on one sphere
}

// A child reference is one int: >= 0 an inner node, < 0 a leaf holding (first, count) as
// -(1 + first * 64 + (count - 1)) with count <= 64, RT_BVH_DONE when a walk is over.
#define RT_BVH_DONE INT_MIN
RT_DEV int bvh_leaf_first(int ref) { return (-ref - 1) >> 6; }
RT_DEV int bvh_leaf_count(int ref) { return ((-ref - 1) & 63) + 1; }

// Closest sphere, as a resumable walk in two kinds of steps, so that a warp can run the lanes that are
// testing boxes together and the lanes that are testing spheres together (the two do not share a
// single instruction; mixing them in one loop body leaves 5 of 32 lanes active on incoherent rays):
//   bvh_inner_step   w.ref >= 0: fetch the node, test both child boxes, descend into the nearer hit
//                    child (the other one waits on the stack with its entry distance) or pop
//   bvh_leaf_step    w.ref < 0: ONE sphere of the leaf, by the scan's own test; after the last one, pop
// best_t / best may carry the result of another scan in; the margins on the box tests (bvh_dn / bvh_up)
// absorb the rounding of the slab arithmetic itself.
// Where a walk keeps the subtrees it still has to visit: a per-thread array (local memory) for the
// one-thread-per-pixel kernels, or a lane-interleaved slice of shared memory for the path tracer's
// hierarchy kernel (rt_warp_bvh.cuh), whose 512 resident walks per SM would otherwise thrash the L1.
struct BvhLocalStack {
  int ref[RT_BVH_STACK];
  float t[RT_BVH_STACK];
  RT_DEV void put(int i, int r, float d) { ref[i] = r; t[i] = d; }
  RT_DEV void get(int i, int& r, float& d) const { r = ref[i]; d = t[i]; }
};
struct BvhSharedStack {
  float2* base;  // entry i of this lane at base[32 * i]
  RT_DEV void put(int i, int r, float d) { base[32 * i] = make_float2(__int_as_float(r), d); }
  RT_DEV void get(int i, int& r, float& d) const { const float2 v = base[32 * i]; r = __float_as_int(v.x); d = v.y; }
};
struct BvhSrc {  // the tree: global memory, or a copy in shared memory when it is small
  const float4* nodes;
  const int32_t* prims;
};
template <typename T> RT_DEV BvhSrc bvh_global_src(const SceneView<T>& sc) {
  BvhSrc b;
  b.nodes = sc.bvh_nodes;
  b.prims = sc.bvh_prims;
  return b;
}

template <typename Stack> struct BvhWalk {
  BvhRay br;
  float tmin, tmax, limit;
  int ref, leaf_k, sp;
  Stack stack;
};
template <typename Stack> RT_DEV void bvh_pop(BvhWalk<Stack>& w) {  // skipping subtrees that start beyond the best hit found since they were pushed
  while (w.sp > 0) {
    --w.sp;
    int r; float d;
    w.stack.get(w.sp, r, d);
    if (bvh_dn(d) <= w.limit) { w.ref = r; w.leaf_k = 0; return; }
  }
  w.ref = RT_BVH_DONE;
}
template <typename T, typename Stack> RT_DEV void bvh_begin(const SceneView<T>& sc, const Ray<T>& r, T best_t, BvhWalk<Stack>& w) {
  w.br = bvh_ray(r);
  w.tmin = (float)r.tmin * 0.999999f;
  w.tmax = (float)r.tmax;
  w.limit = fminf(w.tmax, (float)best_t) * 1.000001f;
  w.ref = sc.n_spheres > 0 ? 0 : RT_BVH_DONE;
  w.leaf_k = 0;
  w.sp = 0;
}
template <typename Stack> RT_DEV void bvh_inner_step(const BvhSrc& src, BvhWalk<Stack>& w) {
  const float4* q = src.nodes + 4 * (size_t)w.ref;
  const float4 q0 = q[0], q1 = q[1], q2 = q[2];
  const int4 q3 = reinterpret_cast<const int4*>(q)[3];
  float n0, f0, n1, f1;
  bvh_slab(w.br, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, n0, f0);
  bvh_slab(w.br, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, n1, f1);
  const bool h0 = bvh_box_hit(n0, f0, w.tmin, w.limit);
  const bool h1 = bvh_box_hit(n1, f1, w.tmin, w.limit);
  if (h0 && h1) {
    const bool first0 = n0 <= n1;
    w.stack.put(w.sp, first0 ? q3.y : q3.x, first0 ? n1 : n0);
    ++w.sp;
    w.ref = first0 ? q3.x : q3.y;
    w.leaf_k = 0;
  } else if (h0 || h1) {
    w.ref = h0 ? q3.x : q3.y;
    w.leaf_k = 0;
  } else {
    bvh_pop(w);
  }
}
template <typename T, typename Stack>
RT_DEV void bvh_leaf_step(const SceneView<T>& sc, const BvhSrc& src, const Ray<T>& r, T& best_t, int& best, int origin, BvhWalk<Stack>& w) {
  const int i = src.prims[bvh_leaf_first(w.ref) + w.leaf_k];
  const T t = bvh_sphere_t<T>(sc, i, r, origin);
  if (t < best_t || (t == best_t && best >= 0 && i < best)) { best_t = t; best = i; }
  if (++w.leaf_k == bvh_leaf_count(w.ref)) {
    w.limit = fminf(w.tmax, (float)best_t) * 1.000001f;
    bvh_pop(w);
  }
}
template <typename T>
RT_DEV void bvh_closest_spheres(const SceneView<T>& sc, const Ray<T>& r, T& best_t, int& best, int origin) {
  BvhWalk<BvhLocalStack> w;
  const BvhSrc src = bvh_global_src(sc);
  bvh_begin<T>(sc, r, best_t, w);
  while (w.ref != RT_BVH_DONE) {
    if (w.ref >= 0) bvh_inner_step(src, w);
    else bvh_leaf_step<T>(sc, src, r, best_t, best, origin, w);
  }
}

// Does any sphere block the segment (World.is_point_visible, world.py:76-78)?  No ordering, early exit.
template <typename T> RT_DEV bool bvh_any_sphere(const SceneView<T>& sc, const Ray<T>& r) {
  if (sc.n_spheres == 0) return false;
  const BvhRay br = bvh_ray(r);
  const float tmin = (float)r.tmin * 0.999999f, limit = (float)r.tmax * 1.000001f;
  int stack[RT_BVH_STACK];
  int sp = 0, ref = 0;
  while (true) {
    if (ref >= 0) {
      const float4* q = sc.bvh_nodes + 4 * (size_t)ref;
      const float4 q0 = __ldg(q), q1 = __ldg(q + 1), q2 = __ldg(q + 2);
      const int4 q3 = __ldg(reinterpret_cast<const int4*>(q + 3));
      float n0, f0, n1, f1;
      bvh_slab(br, q0.x, q0.y, q0.z, q0.w, q1.x, q1.y, n0, f0);
      bvh_slab(br, q1.z, q1.w, q2.x, q2.y, q2.z, q2.w, n1, f1);
      const bool h0 = bvh_box_hit(n0, f0, tmin, limit);
      const bool h1 = bvh_box_hit(n1, f1, tmin, limit);
      if (h0 && h1) { stack[sp++] = q3.y; ref = q3.x; continue; }
      if (h0 || h1) { ref = h0 ? q3.x : q3.y; continue; }
    } else {
      const int first = bvh_leaf_first(ref), count = bvh_leaf_count(ref);
      for (int k = 0; k < count; ++k)
        if (bvh_sphere_blocks<T>(sc, __ldg(sc.bvh_prims + first + k), r)) return true;
    }
    if (sp == 0) return false;
    ref = stack[--sp];
  }
}

// World.ray_intersection with the tree: spheres through the hierarchy, planes scanned (they are few and
// unbounded); same winner as closest_all<T>.
template <typename T>
RT_DEV void closest_bvh(const SceneView<T>& sc, const Ray<T>& r, T& best_t, int& best, int origin = -1) {
  bvh_closest_spheres<T>(sc, r, best_t, best, origin);
  if constexpr (sizeof(T) == 4) {
    scan_plane_block(sc.invm + 12 * (size_t)sc.n_spheres, sc.n_spheres, sc.n_shapes - sc.n_spheres, sc.orig, r, best_t, best, origin);
  } else {
    scan_planes<T>(sc.invm, 0, sc.n_shapes, sc.n_spheres, sc.orig, r, best_t, best);
  }
}
template <typename T> RT_DEV bool any_bvh(const SceneView<T>& sc, const Ray<T>& r) {
  if constexpr (sizeof(T) == 4) {
    if (any_plane_blocks(sc.invm + 12 * (size_t)sc.n_spheres, sc.n_shapes - sc.n_spheres, r)) return true;
    return bvh_any_sphere<T>(sc, r);
  } else {
    // the fp64 scan tests spheres before planes; the answer (any blocker) does not depend on the order
    if (bvh_any_sphere<T>(sc, r)) return true;
    return scan_any<T>(sc.invm + 12 * (size_t)sc.n_spheres, sc.n_spheres, sc.n_shapes, sc.n_spheres, r);
  }
}
