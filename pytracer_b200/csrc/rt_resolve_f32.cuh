// rt_resolve_f32.cuh — OnOff / Flat / PointLight in fp32 (render.py:52,65,157), the throughput path of
// the deterministic renderers (BASELINE config 5).  Same structure as k_resolve<T> — one thread per
// pixel, strata in the reference's order, shapes staged through shared memory by the block — but the
// strata of a pixel are traced TWO AT A TIME against each staged sphere pair: six broadcast LDS.128 feed
// 60 packed FMAs (FFMA2) instead of 30, so the shared-memory pipe stops throttling the FMA pipe.
// Scenes larger than 96 KB of pairs are swept in 48 KB chunks through a DOUBLE BUFFER filled by the TMA
// engine (cp.async.bulk + mbarrier, rt_tma.cuh): one thread posts the copy of chunk c+1, everybody sweeps
// chunk c; a block barrier per chunk only guards the buffer about to be overwritten.
#pragma once
#include "rt_kernels.cuh"
#include "rt_tma.cuh"

struct ResolveSample {
  Ray<float> ray;
  int cand[RT_CAND_CAP];
  int nc;
  float best_t;
  int best;
  bool mine;
  Hit<float> h;
  V3<float> color;
};

RT_DEV void resolve_shade(const SceneView<float>& sc, const RenderArgs& a, ResolveSample& q, const V3<float>& background) {
  q.color = background;
  q.h.idx = -1;
  if (q.mine && q.best >= 0) {
    finish_hit<float>(sc, q.ray, q.best_t, q.best, q.h);
    if (a.algorithm == RT_ALGO_ONOFF) q.color = load3<float>(a.onoff);
    else if (a.algorithm == RT_ALGO_FLAT) q.color = flat_color<float>(sc, q.h);
    else {
      const DevMaterial& mat = sc.materials[sc.material[q.h.idx]];
      q.color = load3<float>(a.ambient) + pigment_color<float>(sc.pigments, mat.emitted_pigment, q.h.u, q.h.v);
    }
  }
}

// BVH (sphere hierarchy instead of the staged sweep) is a template parameter: the walk needs none of the
// sweep's registers, so its instantiation is held to 80 and three blocks share an SM.
// Out of line on purpose: the two instantiations below must round the light term identically (inlined,
// the compiler fuses its products into the caller's additions differently in each, and the linear and the
// hierarchy image of the same frame would differ in the last bit of every lit pixel).
static __device__ __noinline__ V3<float> light_term_f32(const SceneView<float>& sc, const Hit<float>& h, V3<float> ray_dir, int l) {
  return light_term<float>(sc, h, ray_dir, l);
}

template <bool BVH>
__global__ void __launch_bounds__(RT_RESOLVE_THREADS, BVH ? 3 : 2)
k_resolve_f32(const __grid_constant__ SceneView<float> sc, const __grid_constant__ RenderArgs a, const int chunk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int n_planes = sc.n_shapes - sc.n_spheres;
  constexpr bool bvh = BVH;  // sphere hierarchy (rt_bvh.cuh): no pair staging, every thread walks the tree
  const bool single = bvh || sc.n_pairs <= chunk;
  const bool planes_smem = n_planes <= RT_PLANES_SMEM_MAX;
  const float* planes_g = sc.packed + 24 * (size_t)sc.n_pairs;
  float* sh_planes = reinterpret_cast<float*>(smem_raw);
  float4* sh_pairs = reinterpret_cast<float4*>(smem_raw + (planes_smem ? (size_t)n_planes * 48 : 0));
  const float* planes = planes_smem ? sh_planes : planes_g;

  const PixelMap pm = make_pixel_map(a);
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = p < pm.n_pixels;
  int col = 0, row = 0;
  if (active) pm.locate(p, col, row);
  const long long pix = (long long)row * a.width + col;
  const int S2 = a.S > 0 ? a.S * a.S : 1;

  // double buffer + two mbarriers behind it (chunked scenes); one buffer when everything fits
  float4* buf[2] = {sh_pairs, sh_pairs + (single ? 0 : 6 * (size_t)chunk)};
  uint64_t* bars = reinterpret_cast<uint64_t*>(sh_pairs + 6 * (size_t)chunk * (single ? 1 : 2));
  uint32_t parity[2] = {0u, 0u};
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  if (planes_smem) stage_bytes(sh_planes, planes_g, (size_t)n_planes * 48);
  __syncthreads();
  if (single && sc.n_pairs > 0 && !bvh) {  // the whole pair list, once, by the copy engine
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(&bars[0], (uint32_t)sc.n_pairs * 96u);
      tma_load_1d(buf[0], sc.packed, (uint32_t)sc.n_pairs * 96u, &bars[0]);
    }
    mbar_wait(&bars[0], 0);
    parity[0] = 1;
  }

  Pcg aa;
  aa.inc = a.aa_inc;
  aa.state = (a.S > 0 && active) ? pcg_jump(a.aa_state, 2ull * (unsigned long long)pix * S2, a.jump) : 0;
  V3<float> cum = mk3<float>(0.f, 0.f, 0.f);
  int last_hit = -1;
  unsigned int n_closest = 0, n_shadow = 0, n_samples = 0;
  const V3<float> background = load3<float>(a.background);
  const int n_chunks = single ? 1 : (sc.n_pairs + chunk - 1) / chunk;

  auto post_chunk = [&](int c) {  // thread 0: start the bulk copy of chunk c into buffer c & 1
    const int b0 = c * chunk, b1 = min(b0 + chunk, sc.n_pairs);
    const uint32_t bytes = (uint32_t)(b1 - b0) * 96u;
    mbar_arrive_expect_tx(&bars[c & 1], bytes);
    tma_load_1d(buf[c & 1], sc.packed + 24 * (size_t)b0, bytes, &bars[c & 1]);
  };

  // one sweep of all sphere pairs for one or two rays (block-uniform control flow)
  auto sweep = [&](bool two, bool on0, bool on1, const Ray<float>& r0, const Ray<float>& r1, int* c0, int& n0, int* c1, int& n1) {
    if (single) {
      if (two) { if (on0 || on1) sweep_pairs2(buf[0], 0, 0, sc.n_pairs, r0, r1, sc.gate_a, sc.gate_t, c0, n0, c1, n1); }
      else if (on0) sweep_pairs<true>(buf[0], 0, 0, sc.n_pairs, pack_ray(r0, sc.gate_a, sc.gate_t), c0, n0);
      return;
    }
    __syncthreads();  // the previous sweep is done with both buffers
    if (threadIdx.x == 0) post_chunk(0);
    for (int c = 0; c < n_chunks; ++c) {
      const int b0 = c * chunk, b1 = min(b0 + chunk, sc.n_pairs);
      if (c + 1 < n_chunks) {
        if (c >= 1) __syncthreads();  // chunk c - 1, which lived in the buffer about to be refilled, is consumed
        if (threadIdx.x == 0) post_chunk(c + 1);
      }
      mbar_wait(&bars[c & 1], parity[c & 1]);
      parity[c & 1] ^= 1u;
      if (two) { if (on0 || on1) sweep_pairs2(buf[c & 1], b0, b0, b1, r0, r1, sc.gate_a, sc.gate_t, c0, n0, c1, n1); }
      else if (on0) sweep_pairs<true>(buf[c & 1], b0, b0, b1, pack_ray(r0, sc.gate_a, sc.gate_t), c0, n0);
    }
  };

  for (int s = 0; s < S2; s += 2) {  // block-uniform trip count; strata s and s + 1 travel together
    const bool two = s + 1 < S2;
    ResolveSample q0, q1;
    q0.mine = active && stratum_is_mine(a, s);
    q1.mine = two && active && stratum_is_mine(a, s + 1);
    if (active) {  // jitter draws are consumed for every stratum, in order
      q0.ray = primary_ray<float>(a, col, row, s, aa);
      q1.ray = two ? primary_ray<float>(a, col, row, s + 1, aa) : q0.ray;
    }
    q0.nc = q1.nc = 0;
    q0.best = q1.best = -1;
    q0.best_t = q1.best_t = Num<float>::inf();
    if (!bvh) sweep(two, q0.mine, q1.mine, q0.ray, q1.ray, q0.cand, q0.nc, q1.cand, q1.nc);
    if (q0.mine) {
      if (bvh) bvh_closest_spheres<float>(sc, q0.ray, q0.best_t, q0.best, -1);
      else resolve_candidates(sc.invm, sc.n_spheres, q0.cand, q0.nc, q0.ray, q0.best_t, q0.best);
      scan_plane_block(planes, sc.n_spheres, n_planes, sc.orig, q0.ray, q0.best_t, q0.best);
      ++n_closest; ++n_samples;
    }
    if (q1.mine) {
      if (bvh) bvh_closest_spheres<float>(sc, q1.ray, q1.best_t, q1.best, -1);
      else resolve_candidates(sc.invm, sc.n_spheres, q1.cand, q1.nc, q1.ray, q1.best_t, q1.best);
      scan_plane_block(planes, sc.n_spheres, n_planes, sc.orig, q1.ray, q1.best_t, q1.best);
      ++n_closest; ++n_samples;
    }
    resolve_shade(sc, a, q0, background);
    resolve_shade(sc, a, q1, background);
    if (q0.mine) last_hit = q0.best >= 0 ? sc.orig[q0.best] : -1;
    if (q1.mine) last_hit = q1.best >= 0 ? sc.orig[q1.best] : -1;
    if (a.algorithm == RT_ALGO_POINTLIGHT) {
      for (int l = 0; l < sc.n_lights; ++l) {  // render.py:168-191
        bool need0 = q0.mine && q0.best >= 0, need1 = q1.mine && q1.best >= 0;
        Ray<float> s0, s1;
        const V3<float> lp = load3<float>(sc.lights[l].pos);
        bool blocked0 = false, blocked1 = false;
        if (need0) { s0 = shadow_ray<float>(lp, q0.h.point); ++n_shadow; blocked0 = any_plane_blocks(planes, n_planes, s0); }
        if (need1) { s1 = shadow_ray<float>(lp, q1.h.point); ++n_shadow; blocked1 = any_plane_blocks(planes, n_planes, s1); }
        if (!need0) s0 = need1 ? s1 : q0.ray;
        if (!need1) s1 = s0;
        const bool go0 = need0 && !blocked0, go1 = need1 && !blocked1;
        int c0[RT_CAND_CAP], c1[RT_CAND_CAP];
        int m0 = 0, m1 = 0;
        if (bvh) {
          if (go0) blocked0 = bvh_any_sphere<float>(sc, s0);
          if (go1) blocked1 = bvh_any_sphere<float>(sc, s1);
        } else {
          if (single || __syncthreads_or(go0 || go1)) sweep(two, go0, go1, s0, s1, c0, m0, c1, m1);
          if (go0) blocked0 = any_candidate_blocks(sc.invm, sc.n_spheres, c0, m0, s0);
          if (go1) blocked1 = any_candidate_blocks(sc.invm, sc.n_spheres, c1, m1, s1);
        }
        if (need0 && !blocked0) q0.color = q0.color + light_term_f32(sc, q0.h, q0.ray.d, l);
        if (need1 && !blocked1) q1.color = q1.color + light_term_f32(sc, q1.h, q1.ray.d, l);
      }
    }
    if (q0.mine) cum = (a.S > 0) ? cum + q0.color : q0.color;
    if (q1.mine) cum = cum + q1.color;
  }
  if (active) {
    if (a.S > 0) cum = (1.0f / (float)S2) * cum;  // imagetracer.py:99-101
    store_pixel<float>(a, pm.at(p, col, row), cum);
    if (a.out_hit) a.out_hit[pm.at(p, col, row)] = a.hit_mode == RT_HIT_RAY_COUNT ? (int)(n_closest + n_shadow) : last_hit;
  }
  block_count_add(a.counters + CNT_CLOSEST, n_closest);
  block_count_add(a.counters + CNT_SHADOW, n_shadow);
  block_count_add(a.counters + CNT_SAMPLES, n_samples);
}

inline cudaError_t launch_resolve_f32(const SceneView<float>& sc, const RenderArgs& a, cudaStream_t st, LaunchInfo* info) {
  PixelMap pm = make_pixel_map(a);
  if (pm.n_pixels == 0) return cudaSuccess;
  void (*kern)(const SceneView<float>, const RenderArgs, const int) = sc.accel ? k_resolve_f32<true> : k_resolve_f32<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * RT_SMEM_SHAPE_BYTES + 17 * 1024);
  if (e != cudaSuccess) return e;
  const int n_planes = sc.n_shapes - sc.n_spheres;
  const size_t planes_bytes = n_planes <= RT_PLANES_SMEM_MAX ? (size_t)n_planes * 48 : 0;
  // everything in one chunk while it fits 96 KB of shared memory, else 48 KB chunks (512 pairs)
  int chunk = (size_t)sc.n_pairs * 96 <= 2 * RT_SMEM_SHAPE_BYTES ? (sc.n_pairs > 0 ? sc.n_pairs : 1) : RT_SMEM_SHAPE_BYTES / 96;
  if (sc.accel) chunk = 1;  // hierarchy traversal: only the plane block and the (unused) barriers
  const bool single = sc.n_pairs <= chunk || sc.accel;
  size_t smem = planes_bytes + (size_t)chunk * 96 * (single ? 1 : 2) + 16;  // + two mbarriers
  long long blocks = (pm.n_pixels + RT_RESOLVE_THREADS - 1) / RT_RESOLVE_THREADS;
  kern<<<(unsigned)blocks, RT_RESOLVE_THREADS, smem, st>>>(sc, a, chunk);
  if (info) { info->n_launches += 1; info->variant = 0; }
  return cudaGetLastError();
}
