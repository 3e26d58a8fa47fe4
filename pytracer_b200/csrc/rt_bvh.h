// rt_bvh.h — host-side builder of the sphere hierarchy traversed by rt_bvh.cuh.
// Top-down, binned surface-area heuristic (16 bins per axis), leaves of at most RT_BVH_LEAF spheres.
// Boxes are those of the transformed unit spheres (an affine image of the unit sphere spans
// centre_k +- |row k of the 3x3 block| along axis k), computed in fp64 and PADDED: by 2e-4 of the
// box's own extent plus 4e-6 of the extent of the whole set.  The padding is what makes the culling
// conservative against the fp32 arithmetic of both the slab test and the sphere test (a grazing hit
// the fp32 quadratic still reports lies within ~1e-5 of the sphere's extent from the true surface).
// Why that suffices for the slab test too (it runs in fp32 even under the fp64 kernels): rounding the ray
// origin to fp32 moves it by at most 6e-8 |o|.  For an origin inside ~66 scene extents that is below the
// absolute pad (4e-6 of the scene extent); for one further out, every box is about |o| away, so the error
// of the slab parameters is ~1e-7 RELATIVE, which the traversal's own margins (bvh_dn / bvh_up: 1e-6
// relative on every entry / exit parameter) absorb.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <vector>

#ifndef RT_BVH_LEAF
#define RT_BVH_LEAF 1  // measured on B200 (profiles/r2_bvh_sweep.log): 1 | 2 | 4 spheres per leaf = 4.60 | 4.68 | 4.19 Grays/s on config 4, 7.36 | 6.98 | 6.70 on config 5 (fp64)
#endif
#define RT_BVH_SAH_DEPTH 20  // below this depth ranges are halved (balanced), so that the depth stays under
                             // RT_BVH_SAH_DEPTH + log2(n) < RT_BVH_STACK of rt_bvh.cuh for any scene that fits a GPU
#define RT_BVH_BINS 16

struct BvhHostNode {  // 64 bytes, the device layout of rt_bvh.cuh
  float lo0[3], hi0[3], lo1[3], hi1[3];
  int32_t ref0, ref1, pad0, pad1;  // >= 0: inner node; < 0: leaf, -(1 + first * 64 + (count - 1))
};
inline int32_t bvh_leaf_ref(int first, int count) { return -(1 + first * 64 + (count - 1)); }

struct BvhBuild {
  std::vector<BvhHostNode> nodes;
  std::vector<int32_t> prims;  // sphere (sorted) indices, leaf ranges point in here
  int max_depth = 0;
};

namespace bvh_detail {
struct Box {
  double lo[3], hi[3];
  Box() { for (int k = 0; k < 3; ++k) { lo[k] = DBL_MAX; hi[k] = -DBL_MAX; } }
  void grow(const Box& b) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], b.lo[k]); hi[k] = std::max(hi[k], b.hi[k]); } }
  void grow(const double* p) { for (int k = 0; k < 3; ++k) { lo[k] = std::min(lo[k], p[k]); hi[k] = std::max(hi[k], p[k]); } }
  double area() const {
    const double dx = hi[0] - lo[0], dy = hi[1] - lo[1], dz = hi[2] - lo[2];
    return (dx < 0 || dy < 0 || dz < 0) ? 0.0 : 2.0 * (dx * dy + dy * dz + dz * dx);
  }
};
struct Prim {
  Box box;
  double c[3];
  int32_t idx;
};
struct Child {
  Box box;
  int32_t c, n;  // n == 0: inner node c; n > 0: leaf prims[c .. c + n)
};

inline Child build_range(std::vector<Prim>& p, int b, int e, int depth, BvhBuild& out) {
  Child me;
  Box cb;
  for (int i = b; i < e; ++i) { me.box.grow(p[i].box); cb.grow(p[i].c); }
  out.max_depth = std::max(out.max_depth, depth);
  const int count = e - b;
  auto make_leaf = [&]() {
    me.c = (int32_t)out.prims.size();
    me.n = count;
    // ascending sphere index inside a leaf (not required for correctness: ties are resolved on the index)
    std::sort(p.begin() + b, p.begin() + e, [](const Prim& x, const Prim& y) { return x.idx < y.idx; });
    for (int i = b; i < e; ++i) out.prims.push_back(p[i].idx);
    return me;
  };
  if (count <= RT_BVH_LEAF) return make_leaf();
  // binned SAH over the three axes
  int best_axis = -1, best_split = -1;
  double best_cost = DBL_MAX;
  for (int axis = 0; axis < 3 && depth < RT_BVH_SAH_DEPTH; ++axis) {
    const double lo = cb.lo[axis], ext = cb.hi[axis] - cb.lo[axis];
    if (!(ext > 0)) continue;
    Box bins[RT_BVH_BINS];
    int cnt[RT_BVH_BINS] = {0};
    for (int i = b; i < e; ++i) {
      int k = (int)((p[i].c[axis] - lo) / ext * RT_BVH_BINS);
      k = std::min(std::max(k, 0), RT_BVH_BINS - 1);
      bins[k].grow(p[i].box);
      cnt[k]++;
    }
    double right_area[RT_BVH_BINS];
    int right_cnt[RT_BVH_BINS];
    Box acc;
    int c = 0;
    for (int k = RT_BVH_BINS - 1; k > 0; --k) { acc.grow(bins[k]); c += cnt[k]; right_area[k] = acc.area(); right_cnt[k] = c; }
    Box left;
    int lc = 0;
    for (int k = 0; k < RT_BVH_BINS - 1; ++k) {
      left.grow(bins[k]);
      lc += cnt[k];
      if (lc == 0 || right_cnt[k + 1] == 0) continue;
      const double cost = left.area() * lc + right_area[k + 1] * right_cnt[k + 1];
      if (cost < best_cost) { best_cost = cost; best_axis = axis; best_split = k; }
    }
  }
  int mid;
  if (best_axis >= 0) {
    const double lo = cb.lo[best_axis], ext = cb.hi[best_axis] - cb.lo[best_axis];
    auto it = std::partition(p.begin() + b, p.begin() + e, [&](const Prim& q) {
      int k = (int)((q.c[best_axis] - lo) / ext * RT_BVH_BINS);
      k = std::min(std::max(k, 0), RT_BVH_BINS - 1);
      return k <= best_split;
    });
    mid = (int)(it - p.begin());
  } else {  // deep in the tree, or all centroids coincide: halve the range along its widest axis
    int axis = 0;
    for (int k = 1; k < 3; ++k)
      if (cb.hi[k] - cb.lo[k] > cb.hi[axis] - cb.lo[axis]) axis = k;
    mid = b + count / 2;
    std::nth_element(p.begin() + b, p.begin() + mid, p.begin() + e, [axis](const Prim& x, const Prim& y) { return x.c[axis] < y.c[axis]; });
  }
  if (mid <= b || mid >= e) mid = b + count / 2;
  const int32_t my_index = (int32_t)out.nodes.size();
  out.nodes.push_back(BvhHostNode());
  const Child l = build_range(p, b, mid, depth + 1, out);
  const Child r = build_range(p, mid, e, depth + 1, out);
  BvhHostNode& n = out.nodes[my_index];
  for (int k = 0; k < 3; ++k) {
    // round outwards when narrowing to fp32
    n.lo0[k] = std::nextafterf((float)l.box.lo[k], -INFINITY); n.hi0[k] = std::nextafterf((float)l.box.hi[k], INFINITY);
    n.lo1[k] = std::nextafterf((float)r.box.lo[k], -INFINITY); n.hi1[k] = std::nextafterf((float)r.box.hi[k], INFINITY);
  }
  n.ref0 = l.n > 0 ? bvh_leaf_ref(l.c, l.n) : l.c;
  n.ref1 = r.n > 0 ? bvh_leaf_ref(r.c, r.n) : r.c;
  n.pad0 = n.pad1 = 0;
  me.c = my_index;
  me.n = 0;
  return me;
}
}  // namespace bvh_detail

// m64: [n_spheres][12] row-major 3x4 sphere transformations (sorted order).  The root is always node 0
// (a scene of one leaf gets a root whose second child is an empty, never-hit box).
inline BvhBuild bvh_build(const double* m64, int n_spheres) {
  using namespace bvh_detail;
  BvhBuild out;
  if (n_spheres <= 0) return out;
  std::vector<Prim> p(n_spheres);
  Box all;
  for (int i = 0; i < n_spheres; ++i) {
    const double* m = m64 + 12 * (size_t)i;
    for (int k = 0; k < 3; ++k) {
      const double h = std::sqrt(m[4 * k] * m[4 * k] + m[4 * k + 1] * m[4 * k + 1] + m[4 * k + 2] * m[4 * k + 2]);
      p[i].c[k] = m[4 * k + 3];
      p[i].box.lo[k] = m[4 * k + 3] - h;
      p[i].box.hi[k] = m[4 * k + 3] + h;
    }
    p[i].idx = i;
    all.grow(p[i].box);
  }
  double scene_ext = 0.0;
  for (int k = 0; k < 3; ++k) scene_ext = std::max(scene_ext, std::max(std::fabs(all.lo[k]), std::fabs(all.hi[k])));
  for (int k = 0; k < 3; ++k) scene_ext = std::max(scene_ext, all.hi[k] - all.lo[k]);
  for (auto& q : p)
    for (int k = 0; k < 3; ++k) {
      const double pad = 2e-4 * (q.box.hi[k] - q.box.lo[k]) + 4e-6 * scene_ext;
      q.box.lo[k] -= pad;
      q.box.hi[k] += pad;
    }
  out.nodes.reserve(n_spheres);
  const Child root = build_range(p, 0, n_spheres, 0, out);
  if (root.n > 0) {  // everything in one leaf: wrap it
    BvhHostNode n;
    for (int k = 0; k < 3; ++k) {
      n.lo0[k] = std::nextafterf((float)root.box.lo[k], -INFINITY); n.hi0[k] = std::nextafterf((float)root.box.hi[k], INFINITY);
      n.lo1[k] = n.hi1[k] = 3.0e38f;
    }
    // child 1: a point at the far corner of fp32 space; if a ray ever "hits" it, it re-tests prims[0],
    // which cannot change the result
    n.ref0 = bvh_leaf_ref(root.c, root.n); n.ref1 = bvh_leaf_ref(root.c, 1);
    n.pad0 = n.pad1 = 0;
    out.nodes.insert(out.nodes.begin(), n);
  }
  return out;
}
