// rt_resolve_hybrid.cuh — OnOff / Flat / PointLight (render.py:52,65,157) with the reference's own fp64
// decisions at the speed of an fp32 sweep: what RT_PRECISION_AUTO resolves to for these renderers.
//
// Why a third path.  The bit-faithful fp64 kernel (k_resolve<double>) walks World.ray_intersection's loop
// (world.py:55-64) over every shape in fp64 without multiply-add fusion: exact, but 12x slower than the
// fp32 sweep on 4 096 spheres.  The fp32 kernel is fast but its hit decisions are not the reference's.
// Here the O(N) part is an fp32 *gate* that is provably conservative, and everything that decides a
// pixel — roots, closest-hit order, ties, shading — is the unchanged fp64 code on the few spheres that
// pass the gate.  The image is the fp64 kernel's image bit for bit (tests: full-size config 2 / config 5).
//
// The gate.  Sphere.ray_intersection (shapes.py:97-121) misses iff delta/4 = (o'.d')^2 - |d'|^2 (|o'|^2-1)
// <= 0, with o' = M o + t, d' = M d the ray in the sphere's frame.  All rays of one sweep share their
// origin: primary rays start at the perspective camera's position, and a shadow ray P -> L
// (world.py:71-80) lies on the same line as the ray L -> P, whose origin is the light.  So per (origin,
// sphere) the host-side part is precomputed in fp64 by k_co_prep — p = M O + t rounded to fp32, and
// c' = (|p|^2 - 1) - eta (|p|^2 + 1) rounded DOWN (stored negated) — and the sweep only needs d' = M d (9 FMA),
// |d'|^2 (3), p.d' (3) and (p.d')^2 + |d'|^2 (-c') (2): 17 packed FMAs per sphere pair and ray instead of 30, with no
// cancellation in p (the fp32 form M o + t loses |t| / |p| digits when the scene sits far from the
// world's origin).  eta bounds every rounding of the fp32 evaluation (derivation at k_co_prep), so
//     reference delta > 0   ==>   gate > 0:
// no sphere the reference hits is ever dropped; spheres passing the gate without being hit cost one fp64
// test each and change nothing.  Orthogonal cameras (no common origin) keep the plain fp64 kernel.
//
// Schedule: one thread per pixel, the strata of a pixel FOUR at a time against each staged pair record
// (seven broadcast LDS.128 feed 68 FFMA2: the shared-memory pipe runs at 41 % of the FMA pipe's pace);
// tables larger than 96 KB stream through a double buffer filled by the TMA engine (cp.async.bulk +
// mbarrier) like k_resolve_f32.
#pragma once
#include "rt_kernels.cuh"
#include "rt_tma.cuh"

#define RT_CO_REC 28          // floats per sphere PAIR: 9 matrix entries, p (3), -c', pad — element-interleaved
#define RT_CO_REC_BYTES 112
#define RT_CO_RESIDENT_BYTES (96 * 1024)
#define RT_CO_CHUNK_BYTES (48 * 1024)

// Per (origin k, sphere i) record for the gate.  Error budget, u = 2^-24, all norms Euclidean:
//   M^ = fl(M), d^ = fl(d): entries within u;  D^ = fl(M^ d^) by a 3-term FMA chain:
//     |D^ - D|_i <= 5u (|M||d|)_i  ==>  |D^ - D| <= eD |D|,  eD = 5u |M|_F |M^-1|_F   (|D| >= |d| / |M^-1|)
//   p^ = fl(P): |p^ - P| <= u |P|;   a^ = fl(D^.D^): |a^ - a| <= (2 eD + 4u) a
//   hb^ = fl(p^.D^): |hb^ - hb| <= (eD + 4u) |P| sqrt(a)   ==>  |hb^^2 - hb^2| <= (2 eD + 9u) a S,  S = |P|^2
//   fl(hb^^2 + a^ (-c')), one product and one fused multiply-add: + u (hb^2 + a |c'|) + |a^ - a| |c'|
//   total <= a ((4 eD + 14u) S + (2 eD + 5u)) <= a (4 eD + 14u)(S + 1).
// eta = (22 kappa + 16) u covers it with a tenth to spare (4 eD + 14u = (20 kappa + 14) u), kappa = |M|_F |M^-1|_F >= 3;
// the reference's own fp64 roundings of delta (1e-16 relative) disappear in that margin.
__global__ void __launch_bounds__(256)
k_co_prep(const __grid_constant__ SceneView<double> sc, const double ox, const double oy, const double oz,
          float* __restrict__ out, const int n_pairs, const int n_origins) {
  const int per = 2 * n_pairs;
  const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long long)per * n_origins) return;
  const int k = (int)(idx / per), i = (int)(idx - (long long)k * per);
  float* rec = out + ((size_t)k * n_pairs + (size_t)(i >> 1)) * RT_CO_REC + (i & 1);
  if (i >= sc.n_spheres) {  // padding of an odd sphere count: zero matrix => a = 0, gate = 0, never a candidate
#pragma unroll
    for (int e = 0; e < 14; ++e) rec[2 * e] = 0.0f;
    return;
  }
  const double* im = sc.invm + 12 * (size_t)i;
  const double* mm = sc.m + 12 * (size_t)i;
  V3<double> O = mk3<double>(ox, oy, oz);
  if (k > 0) O = mk3<double>(sc.lights[k - 1].pos[0], sc.lights[k - 1].pos[1], sc.lights[k - 1].pos[2]);
  const V3<double> P = xf_point<double>(im, O);
  const double S = dot(P, P);
  double fi = 0.0, fm = 0.0;
#pragma unroll
  for (int r = 0; r < 3; ++r)
#pragma unroll
    for (int c = 0; c < 3; ++c) { fi += im[4 * r + c] * im[4 * r + c]; fm += mm[4 * r + c] * mm[4 * r + c]; }
  const double kappa = ::sqrt(fi) * ::sqrt(fm);
  const double u = 5.9604644775390625e-08;  // 2^-24
  const double eta = (22.0 * kappa + 16.0) * u;
  const double cp = (S - 1.0) - eta * (S + 1.0);
  float px = (float)P.x, py = (float)P.y, pz = (float)P.z;
  float fc = __double2float_ru(-cp);  // -c', rounded up: c' itself rounds down
  const bool finite = isfinite(px) && isfinite(py) && isfinite(pz) && isfinite(fc) && isfinite((float)kappa);
  if (!finite) { px = py = pz = 0.0f; fc = 3.0e38f; }  // degenerate transformation: always a candidate
  const int at[9] = {0, 1, 2, 4, 5, 6, 8, 9, 10};
#pragma unroll
  for (int e = 0; e < 9; ++e) rec[2 * e] = (float)im[at[e]];
  rec[2 * 9] = px; rec[2 * 10] = py; rec[2 * 11] = pz; rec[2 * 12] = fc; rec[2 * 13] = 0.0f;
}

template <int R> struct CoDirs {
  float x[R], y[R], z[R];
};

// The gate over sphere pairs [p0, p1) whose records start at shared-memory address `recs` (pair `base`), for R rays
// that share the records' origin.  Spheres that pass go to cand[r][] (ring of RT_CAND_CAP, nc[r] counts
// all of them) in ascending sphere index.  Out of line so that the loop gets its own register allocation,
// whatever the caller keeps alive around it.  A ray with a zero direction passes nowhere (a = 0).
template <int OFF> RT_DEV float4 lds128(uint32_t addr) {  // the records are in shared memory: LDS with an immediate offset
  float4 v;
  asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr), "n"(OFF));
  return v;
}
template <int R>
static __device__ __noinline__ void co_sweep(const uint32_t recs, int base, int p0, int p1, const CoDirs<R> d,
                                             int* __restrict__ cand, int* __restrict__ nc) {
  f32x2 bx[R], by[R], bz[R];
  int n[R];
#pragma unroll
  for (int r = 0; r < R; ++r) { bx[r] = pk2(d.x[r], d.x[r]); by[r] = pk2(d.y[r], d.y[r]); bz[r] = pk2(d.z[r], d.z[r]); n[r] = nc[r]; }
#pragma unroll 1
  for (int p = p0; p < p1; ++p) {
    const uint32_t q = recs + (uint32_t)(p - base) * RT_CO_REC_BYTES;
    const float4 v0 = lds128<0>(q), v1 = lds128<16>(q), v2 = lds128<32>(q), v3 = lds128<48>(q), v4 = lds128<64>(q),
                 v5 = lds128<80>(q), v6 = lds128<96>(q);
    const f32x2 m00 = pk2(v0.x, v0.y), m01 = pk2(v0.z, v0.w), m02 = pk2(v1.x, v1.y);
    const f32x2 m10 = pk2(v1.z, v1.w), m11 = pk2(v2.x, v2.y), m12 = pk2(v2.z, v2.w);
    const f32x2 m20 = pk2(v3.x, v3.y), m21 = pk2(v3.z, v3.w), m22 = pk2(v4.x, v4.y);
    const f32x2 px = pk2(v4.z, v4.w), py = pk2(v5.x, v5.y), pz = pk2(v5.z, v5.w), ncp = pk2(v6.x, v6.y);
    float lo[R], hi[R];
    float top = 0.0f;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const f32x2 dx = fma2(m00, bx[r], fma2(m01, by[r], mul2(m02, bz[r])));
      const f32x2 dy = fma2(m10, bx[r], fma2(m11, by[r], mul2(m12, bz[r])));
      const f32x2 dz = fma2(m20, bx[r], fma2(m21, by[r], mul2(m22, bz[r])));
      const f32x2 a = fma2(dx, dx, fma2(dy, dy, mul2(dz, dz)));
      const f32x2 hb = fma2(px, dx, fma2(py, dy, mul2(pz, dz)));
      upk2(fma2(a, ncp, mul2(hb, hb)), lo[r], hi[r]);
      top = fmaxf(top, fmaxf(lo[r], hi[r]));
    }
    if (top > 0.0f) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        if (lo[r] > 0.0f) { cand[r * RT_CAND_CAP + (n[r] & (RT_CAND_CAP - 1))] = 2 * p; ++n[r]; }
        if (hi[r] > 0.0f) { cand[r * RT_CAND_CAP + (n[r] & (RT_CAND_CAP - 1))] = 2 * p + 1; ++n[r]; }
      }
    }
  }
#pragma unroll
  for (int r = 0; r < R; ++r) nc[r] = n[r];
}

// World.ray_intersection (world.py:51-69) on the spheres that passed the gate + every plane, in the
// reference's arithmetic and order: ascending index, strict '<' (the first shape wins ties).
RT_DEV void hyb_closest(const SceneView<double>& sc, const Ray<double>& r, const int* cand, int nc, double& best_t, int& best) {
  if (nc <= RT_CAND_CAP) {
    for (int j = 0; j < nc; ++j) {
      const int i = cand[j];
      const double t = sphere_t<double>(sc.invm + 12 * (size_t)i, r);
      if (t < best_t) { best_t = t; best = i; }
    }
  } else {  // the line crosses more spheres than the list holds: the reference's plain loop
    for (int i = 0; i < sc.n_spheres; ++i) {
      const double t = sphere_t<double>(sc.invm + 12 * (size_t)i, r);
      if (t < best_t) { best_t = t; best = i; }
    }
  }
  scan_planes<double>(sc.invm, 0, sc.n_shapes, sc.n_spheres, sc.orig, r, best_t, best);
}

// World.is_point_visible's loop (world.py:76-78) on the gated spheres + every plane
RT_DEV bool hyb_blocked(const SceneView<double>& sc, const Ray<double>& sr, const int* cand, int nc) {
  if (nc <= RT_CAND_CAP) {
    for (int j = 0; j < nc; ++j) {
      const int i = cand[j];
      if (scan_any<double>(sc.invm + 12 * (size_t)i, i, i + 1, sc.n_spheres, sr)) return true;
    }
  } else if (scan_any<double>(sc.invm, 0, sc.n_spheres, sc.n_spheres, sr)) {
    return true;
  }
  return scan_any<double>(sc.invm + 12 * (size_t)sc.n_spheres, sc.n_spheres, sc.n_shapes, sc.n_spheres, sr);
}

template <int R>
__global__ void __launch_bounds__(RT_RESOLVE_THREADS, 2)
k_resolve_hyb(const __grid_constant__ SceneView<double> sc, const __grid_constant__ RenderArgs a,
              const float* __restrict__ co, const int n_pairs, const int n_origins, const int chunk) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // resident: the tables of all origins fit and are loaded once; otherwise every sweep streams the table
  // of its origin through two chunk buffers
  const bool resident = (size_t)n_origins * n_pairs * RT_CO_REC_BYTES <= RT_CO_RESIDENT_BYTES;
  float4* const sh = reinterpret_cast<float4*>(smem_raw);
  float4* const buf[2] = {sh, sh + 7 * (size_t)chunk};
  uint64_t* const bars = reinterpret_cast<uint64_t*>(sh + 7 * (size_t)(resident ? n_origins * n_pairs : 2 * chunk));
  uint32_t parity[2] = {0u, 0u};
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (resident && n_pairs > 0) {
    const uint32_t bytes = (uint32_t)(n_origins * n_pairs) * RT_CO_REC_BYTES;
    if (threadIdx.x == 0) {
      mbar_arrive_expect_tx(&bars[0], bytes);
      tma_load_1d(sh, co, bytes, &bars[0]);
    }
    mbar_wait(&bars[0], 0);
    parity[0] = 1;
  }
  const int n_chunks = resident ? 1 : (n_pairs + chunk - 1) / chunk;

  const PixelMap pm = make_pixel_map(a);
  const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const bool active = p < pm.n_pixels;
  int col = 0, row = 0;
  if (active) pm.locate(p, col, row);
  const long long pix = (long long)row * a.width + col;
  const int S2 = a.S > 0 ? a.S * a.S : 1;

  // one gate sweep of every sphere pair for this thread's R rays from origin k (block-uniform control flow)
  auto sweep = [&](int k, bool on, const CoDirs<R>& d, int* cand, int* nc) {
    if (n_pairs == 0) return;
    if (resident) {
      if (on) co_sweep<R>(smem_u32(sh + 7 * (size_t)k * n_pairs), 0, 0, n_pairs, d, cand, nc);
      return;
    }
    const float* table = co + (size_t)k * n_pairs * RT_CO_REC;
    auto post = [&](int c) {  // thread 0: bulk copy of chunk c into buffer c & 1
      const int b0 = c * chunk, b1 = min(b0 + chunk, n_pairs);
      const uint32_t bytes = (uint32_t)(b1 - b0) * RT_CO_REC_BYTES;
      mbar_arrive_expect_tx(&bars[c & 1], bytes);
      tma_load_1d(buf[c & 1], table + (size_t)b0 * RT_CO_REC, bytes, &bars[c & 1]);
    };
    __syncthreads();  // the previous sweep is done with both buffers
    if (threadIdx.x == 0) post(0);
    for (int c = 0; c < n_chunks; ++c) {
      const int b0 = c * chunk, b1 = min(b0 + chunk, n_pairs);
      if (c + 1 < n_chunks) {
        if (c >= 1) __syncthreads();  // chunk c - 1, in the buffer about to be refilled, is consumed
        if (threadIdx.x == 0) post(c + 1);
      }
      mbar_wait(&bars[c & 1], parity[c & 1]);
      parity[c & 1] ^= 1u;
      if (on) co_sweep<R>(smem_u32(buf[c & 1]), b0, b0, b1, d, cand, nc);
    }
  };

  Pcg aa;
  aa.inc = a.aa_inc;
  aa.state = (a.S > 0 && active) ? pcg_jump(a.aa_state, 2ull * (unsigned long long)pix * S2, a.jump) : 0;
  V3<double> cum = mk3<double>(0.0, 0.0, 0.0);
  int last_hit = -1;
  unsigned int n_closest = 0, n_shadow = 0, n_samples = 0;
  const V3<double> background = load3<double>(a.background);
  const bool point_light = a.algorithm == RT_ALGO_POINTLIGHT;

  for (int s0 = 0; s0 < S2; s0 += R) {  // block-uniform trip count; R strata travel together
    Ray<double> ray[R];
    bool mine[R];
    bool any_mine = false;
    CoDirs<R> gd;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const int s = s0 + r;
      mine[r] = active && s < S2 && stratum_is_mine(a, s);
      gd.x[r] = gd.y[r] = gd.z[r] = 0.0f;
      if (active && s < S2) ray[r] = primary_ray<double>(a, col, row, s, aa);  // jitter draws are consumed for every stratum, in order
      if (mine[r]) { gd.x[r] = (float)ray[r].d.x; gd.y[r] = (float)ray[r].d.y; gd.z[r] = (float)ray[r].d.z; any_mine = true; }
    }
    int cand[R * RT_CAND_CAP];
    int nc[R];
#pragma unroll
    for (int r = 0; r < R; ++r) nc[r] = 0;
    sweep(0, any_mine, gd, cand, nc);

    Hit<double> h[R];
    V3<double> color[R];
    bool lit[R];
    bool any_lit = false;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      color[r] = background;
      h[r].idx = -1;
      lit[r] = false;
      if (!mine[r]) continue;
      double best_t = Num<double>::inf();
      int best = -1;
      hyb_closest(sc, ray[r], cand + r * RT_CAND_CAP, nc[r], best_t, best);
      ++n_closest; ++n_samples;
      last_hit = best >= 0 ? sc.orig[best] : -1;
      if (best < 0) continue;
      finish_hit<double>(sc, ray[r], best_t, best, h[r]);
      if (a.algorithm == RT_ALGO_ONOFF) color[r] = load3<double>(a.onoff);
      else if (a.algorithm == RT_ALGO_FLAT) color[r] = flat_color<double>(sc, h[r]);
      else {  // render.py:163-167
        const DevMaterial& mat = sc.materials[sc.material[best]];
        color[r] = load3<double>(a.ambient) + pigment_color<double>(sc.pigments, mat.emitted_pigment, h[r].u, h[r].v);
        lit[r] = true;
        any_lit = true;
      }
    }
    if (point_light) {
      for (int l = 0; l < sc.n_lights; ++l) {  // render.py:168-191
        if (!resident && !__syncthreads_or(any_lit)) continue;  // nobody in the block needs this sweep
        const V3<double> lp = load3<double>(sc.lights[l].pos);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          gd.x[r] = gd.y[r] = gd.z[r] = 0.0f;
          nc[r] = 0;
          if (lit[r]) {  // the line of the shadow ray, taken from the light's side
            gd.x[r] = (float)(lp.x - h[r].point.x); gd.y[r] = (float)(lp.y - h[r].point.y); gd.z[r] = (float)(lp.z - h[r].point.z);
          }
        }
        sweep(1 + l, any_lit, gd, cand, nc);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          if (!lit[r]) continue;
          const Ray<double> sr = shadow_ray<double>(lp, h[r].point);
          ++n_shadow;
          if (!hyb_blocked(sc, sr, cand + r * RT_CAND_CAP, nc[r])) color[r] = color[r] + light_term<double>(sc, h[r], ray[r].d, l);
        }
      }
    }
#pragma unroll
    for (int r = 0; r < R; ++r)
      if (mine[r]) cum = (a.S > 0) ? cum + color[r] : color[r];
  }
  if (active) {
    if (a.S > 0) cum = (1.0 / (double)S2) * cum;  // imagetracer.py:99-101
    store_pixel<double>(a, pm.at(p, col, row), cum);
    if (a.out_hit) a.out_hit[pm.at(p, col, row)] = a.hit_mode == RT_HIT_RAY_COUNT ? (int)(n_closest + n_shadow) : last_hit;
  }
  block_count_add(a.counters + CNT_CLOSEST, n_closest);
  block_count_add(a.counters + CNT_SHADOW, n_shadow);
  block_count_add(a.counters + CNT_SAMPLES, n_samples);
}

// `co` must hold (1 + n_lights) * ceil(n_spheres / 2) records (resolve_hybrid_table_bytes)

inline cudaError_t launch_resolve_hybrid_impl(const SceneView<double>& sc, const RenderArgs& a, float* co, cudaStream_t st, LaunchInfo* info) {
  PixelMap pm = make_pixel_map(a);
  if (pm.n_pixels == 0) return cudaSuccess;
  const int n_pairs = (sc.n_spheres + 1) / 2;
  const int n_origins = 1 + (a.algorithm == RT_ALGO_POINTLIGHT ? sc.n_lights : 0);
  // the camera's position: Camera.fire_ray's origin (camera.py:103-124), the same for every pixel
  // (camera_fire_f64 evaluates it as ((lx m0 + 0 m1) + 0 m2) + m3; the gate's margin is 1e9 times any difference)
  const double* cm = a.cam.m;
  const double ox = -a.cam.dist * cm[0] + cm[3], oy = -a.cam.dist * cm[4] + cm[7], oz = -a.cam.dist * cm[8] + cm[11];
  if (n_pairs > 0) {
    const long long items = 2ll * n_pairs * n_origins;
    k_co_prep<<<(unsigned)((items + 255) / 256), 256, 0, st>>>(sc, ox, oy, oz, co, n_pairs, n_origins);
    if (info) info->n_launches += 1;
  }
  const size_t table_bytes = (size_t)n_origins * n_pairs * RT_CO_REC_BYTES;
  const bool resident = table_bytes <= RT_CO_RESIDENT_BYTES;
  const int chunk = resident ? (n_pairs > 0 ? n_pairs : 1) : RT_CO_CHUNK_BYTES / RT_CO_REC_BYTES;
  const size_t smem = (resident ? (table_bytes ? table_bytes : 16) : 2 * (size_t)chunk * RT_CO_REC_BYTES) + 16;
  const int S2 = a.S > 0 ? a.S * a.S : 1;
  void (*kern)(const SceneView<double>, const RenderArgs, const float*, const int, const int, const int) =
      S2 >= 4 ? k_resolve_hyb<4> : k_resolve_hyb<1>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * RT_CO_CHUNK_BYTES + 1024);
  if (e != cudaSuccess) return e;
  const long long blocks = (pm.n_pixels + RT_RESOLVE_THREADS - 1) / RT_RESOLVE_THREADS;
  kern<<<(unsigned)blocks, RT_RESOLVE_THREADS, smem, st>>>(sc, a, co, n_pairs, n_origins, chunk);
  if (info) { info->n_launches += 1; info->variant = 0; }
  return cudaGetLastError();
}
