// rt_warp.cuh — PathTracer (render.py:99-139) as a warp-cooperative wavefront (fp32).
//
// Why: on demo.txt 48 % of the samples trace one ray while 6.6 % trace up to 1 111; a thread-per-
// sample walk leaves most lanes idle.  Here a warp owns a group of pixels and keeps, in shared
// memory, a LIFO stack of *scatter records* — one per surface interaction that still owes
// num_of_rays scattered rays {point, normal | mirror direction, child throughput, depth, rng}.
// Every iteration the 32 lanes take the next 32 owed rays off the top of the stack (a record is
// shared by up to N consecutive lanes: broadcast reads), each lane scatters, traces and shades its
// ray, and the lanes whose ray must branch again push a new record; positions come from one ballot
// (warp-level compaction).  Lanes stay full whatever single samples do, and the closest-hit loop
// over the shapes stays perfectly convergent: all lanes walk the same shape list.
//
// The estimator is the reference's (N children per interaction, Russian roulette from rr_limit,
// children cut beyond max_depth); it is linear, so each traced ray adds
// throughput * (emitted | background) to its pixel.  Random numbers: a record carries a 64-bit
// state; child i owns z = mix64(state + (i+1)*phi64) (rt_pcg.cuh: child_stream) — its two scatter
// uniforms are the halves of z, a roulette draw is a PCG32 step from z — so the image depends only
// on (sample index, position in the tree) and not on lane assignment, GPU count or partition.
#pragma once
#include "rt_kernels.cuh"

#define RT_WARP_MAX_THREADS 256
#ifndef RT_WARP_SPLIT
#define RT_WARP_SPLIT 1
#endif


struct __align__(16) ScatterRec {
  float4 a;  // P.xyz, meta (slot | kind << 5 | child depth << 6 (10 bits) | origin shape << 16 (0xFFFF: none))
  float4 b;  // nd.xyz (normal, or mirror direction for a specular surface), w.r
  float4 c;  // w.g, w.b, rng state lo, hi
};

struct WarpCfg {
  int cap;            // records per warp stack
  int group;          // pixels per task (G)
  int per_pixel;      // strata of a pixel traced by this rank (L)
  int rounds;         // ceil(G * L / 32)
  int per_warp_bytes;
  int shape_bytes;    // shared memory reserved for the shape block (0 = read from global)
  long long n_tasks;
  // values every iteration needs, precomputed so that they are constant-bank operands instead of
  // registers: background colour in fp32, 1/N, 1/spp, the multiply-shift magic for x / N
  float bg[3], inv_n, inv_spp;
  unsigned n_magic;
  double inv_w, inv_h, inv_s;  // 1 / width, 1 / height, 1 / samples_per_side (fp64, host-rounded)
  unsigned long long n_samples;  // samples this launch traces (all tasks), added to the counter once
};

// SMALL instantiation: every table of the scene lives in shared memory at FIXED offsets (capacity for
// the largest small scene, 3.3 KB), re-packed by the staging loop into what the per-ray code reads:
//   SM_INVM / SM_M   [16][12] fp32 transforms (three LDS.128 per shape)
//   SM_ORIG          [16] index in World.shapes
//   SM_SHAPE_MAT     [16] the shape's material as one int4 {brdf_kind, brdf_pigment, emitted_pigment, flags}
//                    (no material-index indirection)
//   SM_PIG           [32] 48-byte pigment records {kind, steps, tex_w, tex_h | c1.xyz, c2.x | c2.yz, texture handle}
#define RT_SMALL_MAX_SPHERES 8
#define RT_SMALL_MAX_SHAPES 16
#define RT_SMALL_MAX_MATERIALS 16
#define RT_SMALL_MAX_PIGMENTS 32
enum {
  SM_INVM = 0,
  SM_M = SM_INVM + RT_SMALL_MAX_SHAPES * 48,
  SM_ORIG = SM_M + RT_SMALL_MAX_SHAPES * 48,
  SM_SHAPE_MAT = SM_ORIG + RT_SMALL_MAX_SHAPES * 4,
  SM_PIG = SM_SHAPE_MAT + RT_SMALL_MAX_SHAPES * 16,
  SM_PIG_STRIDE = 48,
  SM_BYTES = SM_PIG + RT_SMALL_MAX_PIGMENTS * SM_PIG_STRIDE
};
#ifndef RT_SMALL_THREADS
#define RT_SMALL_THREADS 256
#endif
#ifndef RT_SMALL_MINB
#define RT_SMALL_MINB 3
#endif

RT_DEV void stage_words(void* sh, const void* g, int bytes) {  // 4-byte granularity
  const uint32_t* src = reinterpret_cast<const uint32_t*>(g);
  uint32_t* dst = reinterpret_cast<uint32_t*>(sh);
  for (int i = threadIdx.x; i < bytes / 4; i += blockDim.x) dst[i] = __ldg(src + i);
}

// ---- shared memory by 32-bit window address ------------------------------------------------------
// On sm_100 the address an LDS takes is (cluster CTA id << 24) + offset, and the compiler re-derives
// that base (S2R SR_CgaCtaId + LEA, ~25 cycles of latency each time) at nearly every access when
// registers are short: ncu counted 11 such sequences per iteration of the drain loop on demo.txt, 6 %
// of the instruction stream.  The wavefront kernel therefore keeps ONE base address per warp that has
// been passed through a shuffle (not re-derivable) and addresses its work stack and the staged scene
// tables relative to it with explicit ld.shared / st.shared.
RT_DEV float4 lds4(uint32_t a) {  // read/write data (work stack): ordered against st.shared and __syncwarp
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a) : "memory");
  return v;
}
RT_DEV void sts4(uint32_t a, float4 v) {
  asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" :: "r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
// tables that are constant once staged: no "memory" clobber, so loads may move across ordinary memory
// operations; volatile keeps them behind the __syncthreads() that ends the staging
RT_DEV float4 lds4c(uint32_t a) {
  float4 v;
  asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
  return v;
}
RT_DEV int4 lds4ic(uint32_t a) {
  int4 v;
  asm volatile("ld.shared.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
  return v;
}
RT_DEV int ldsic(uint32_t a) {
  int v;
  asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(a));
  return v;
}
RT_DEV Rows3 lds_rows(uint32_t a) {
  Rows3 m;
  m.r0 = lds4c(a); m.r1 = lds4c(a + 16); m.r2 = lds4c(a + 32);
  return m;
}

// Closest hit over a handful of shapes staged at `sb` (demo.txt: one sphere, two planes): World.ray_intersection
// (world.py:51-69) without a sweep / candidate list, every shape tested directly and without branches —
// with 32 different rays per warp some lane crosses every shape anyway.  Rays of the path tracer are
// unbounded (tmax = inf, ray.py:31), so only tmin is tested.  Spheres and planes keep separate minima
// (strict '<': the first of each kind wins a tie like world.py:62) and the rare sphere-vs-plane tie is
// resolved once at the end on the index in World.shapes.
RT_DEV void closest_small(uint32_t sb, int n_spheres, int n_shapes, const Ray<float>& r, int origin, float& best_t, int& best) {
  const float inf = Num<float>::inf();
  float ts = inf, tp = inf;
  int is = -1, ip = -1;
  uint32_t a = sb + SM_INVM;
#pragma unroll 1
  for (int i = 0; i < n_spheres; ++i, a += 48) {
    float tm, half;
    bool ok = sphere_cross_rows(lds_rows(a), r, tm, half);
    const float t1 = tm - half;
    float t = (t1 > r.tmin) ? t1 : tm + half;   // the first root beyond tmin, shapes.py:112-119
    if (i == origin) { t = 2.0f * tm; ok = true; }  // see sphere_t_at
    if (ok && t > r.tmin && t < ts) { ts = t; is = i; }
  }
#pragma unroll 1
  for (int i = n_spheres; i < n_shapes; ++i, a += 48) {
    const float4 row = lds4c(a + 32);
    const float oz = fmaf(r.o.x, row.x, fmaf(r.o.y, row.y, fmaf(r.o.z, row.z, row.w)));
    const float dz = fmaf(r.d.x, row.x, fmaf(r.d.y, row.y, r.d.z * row.z));
    const float t = -oz * fast_rcp(dz);
    if (!(fabsf(dz) < 1e-5f) && t > r.tmin && t < tp && i != origin) { tp = t; ip = i; }
  }
  bool plane = tp < ts;
  if (tp == ts && ip >= 0 && is >= 0) plane = ldsic(sb + SM_ORIG + 4 * ip) < ldsic(sb + SM_ORIG + 4 * is);
  best_t = plane ? tp : ts;
  best = plane ? ip : is;
}

// Pigment.get_color (materials.py:58, :70-82, :96-100) from a staged 48-byte record
RT_DEV V3<float> pigment_color_small(uint32_t pa, float u, float v) {
  const int4 h = lds4ic(pa);
  const float4 q = lds4c(pa + 16);
  if (h.x == RT_PIGMENT_UNIFORM) return mk3<float>(q.x, q.y, q.z);
  const float4 w = lds4c(pa + 32);
  if (h.x == RT_PIGMENT_CHECKERED) {
    const int iu = __float2int_rd(u * (float)h.y), iv = __float2int_rd(v * (float)h.y);
    return (((iu ^ iv) & 1) == 0) ? mk3<float>(q.x, q.y, q.z) : mk3<float>(q.w, w.x, w.y);
  }
  int col = (int)(u * (float)h.z), row = (int)(v * (float)h.w);
  col = min(max(col, 0), h.z - 1);
  row = min(max(row, 0), h.w - 1);
  const cudaTextureObject_t tex = ((unsigned long long)__float_as_uint(w.w) << 32) | (unsigned long long)__float_as_uint(w.z);
  const float4 t = fetch_texel(tex, col + 0.5f, row + 0.5f);
  return mk3<float>(t.x, t.y, t.z);
}

// Primary ray of the path tracer's warp kernels.  Same formulas as primary_ray<T> (imagetracer.py:48-58,
// :88-93, camera.py) in fp64, with the five divisions by width / height / S / 0xFFFFFFFF replaced by
// multiplications with the host-rounded reciprocals: a correctly rounded fp64 division costs ~30
// instructions and a sample needs six of them.  The fp64 values differ from the reference's by at most an
// ulp (1e-16), so the fp32 ray they round to is the reference's ray except when a coordinate sits within
// 1e-16 of an fp32 rounding boundary (~1e-8 of the rays) — immaterial for this renderer, whose parity is
// statistical; the deterministic renderers, the replay mode and the ray probes keep the divisions.
RT_DEV Ray<float> primary_ray_mul(const RenderArgs& a, double inv_w, double inv_h, double inv_s, int col, int row, int s, Pcg& aa) {
  double up = 0.5, vp = 0.5;
  if (a.S > 0) {
    const int ir = a.S <= 1024 ? (int)(((float)s + 0.5f) * (float)inv_s) : s / a.S;  // s / S, exact (s < S^2 <= 2^20)
    const int ic = s - ir * a.S;
    const double r1 = (double)pcg_random(aa) * (1.0 / 4294967295.0);
    const double r2 = (double)pcg_random(aa) * (1.0 / 4294967295.0);
    up = ((double)ic + r1) * inv_s;
    vp = ((double)ir + r2) * inv_s;
  }
  const double u = ((double)col + up) * inv_w;
  const double v = 1.0 - ((double)row + vp) * inv_h;
  V3<double> o, d;
  camera_fire_f64(a.cam, u, v, o, d);
  Ray<float> r;
  r.o = cast3<float>(o);
  r.d = cast3<float>(d);
  r.tmin = 1.0e-5f;
  r.tmax = Num<float>::inf();
  return r;
}

RT_DEV int own_stratum(const RenderArgs& a, int ls) {
  return (a.part_mode == RT_PART_SPP && a.part_count > 1) ? a.part_rank + ls * a.part_count : ls;
}

// SMALL: scenes of at most 8 spheres (demo.txt): the sweep is not unrolled and the kernel is held to
// 80 registers so that three CTAs (24 warps) fit an SM; large scenes are FMA-bound in the unrolled
// sweep and keep the 128-register budget.
// ACC: how per-pixel sums are kept while a task is in flight.
//   ACC_REG   one pixel per task: lane-private registers, one warp reduction at the end
//   ACC_LANES 2..8 pixels per task: acc[slot][channel][lane] in shared memory — every lane adds into
//             its own column (no conflicts, no atomics), reduced over lanes at the end
//   ACC_SEG   more pixels per task: segmented warp scan over the lanes of a record, then one
//             shared-memory atomic per record
enum { ACC_REG = 0, ACC_LANES = 1, ACC_SEG = 2 };
#define RT_ACC_LANES_MAX_GROUP 8

template <bool SHAPES_SMEM, int ACC, bool SMALL>
__global__ void __launch_bounds__(SMALL ? RT_SMALL_THREADS : RT_WARP_MAX_THREADS, SMALL ? RT_SMALL_MINB : 2)
k_pt_warp(const __grid_constant__ SceneView<float> sc, const __grid_constant__ RenderArgs a,
          const __grid_constant__ WarpCfg cfg) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const unsigned FULL = 0xffffffffu;
  const int warp = threadIdx.x >> 5;
  ScanSrc<float> src = global_src<float>(sc);
  if (SMALL) {
    stage_words(smem_raw + SM_INVM, sc.invm, sc.n_shapes * 48);
    stage_words(smem_raw + SM_M, sc.m, sc.n_shapes * 48);
    stage_words(smem_raw + SM_ORIG, sc.orig, sc.n_shapes * 4);
    for (int i = threadIdx.x; i < sc.n_shapes; i += blockDim.x) {
      const DevMaterial& mt = sc.materials[sc.material[i]];
      reinterpret_cast<int4*>(smem_raw + SM_SHAPE_MAT)[i] = make_int4(mt.brdf_kind, mt.brdf_pigment, mt.emitted_pigment, mt.flags);
    }
    for (int i = threadIdx.x; i < sc.n_pigments; i += blockDim.x) {
      const DevPigment& pg = sc.pigments[i];
      float4* o = reinterpret_cast<float4*>(smem_raw + SM_PIG + i * SM_PIG_STRIDE);
      o[0] = make_float4(__int_as_float(pg.kind), __int_as_float(pg.steps), __int_as_float(pg.tex_w), __int_as_float(pg.tex_h));
      o[1] = make_float4(pg.c1[0], pg.c1[1], pg.c1[2], pg.c2[0]);
      o[2] = make_float4(pg.c2[1], pg.c2[2], __uint_as_float((uint32_t)pg.tex), __uint_as_float((uint32_t)((unsigned long long)pg.tex >> 32)));
    }
    __syncthreads();
  } else if (SHAPES_SMEM) {
    stage_bytes(smem_raw, sc.packed, (size_t)sc.n_pairs * 96 + (size_t)(sc.n_shapes - sc.n_spheres) * 48);
    __syncthreads();
    src.pairs = reinterpret_cast<const float4*>(smem_raw);
    src.planes = reinterpret_cast<const float*>(smem_raw) + 24 * (size_t)sc.n_pairs;
  }
  // lane index and the two shared-window addresses the loop lives on, each passed through a shuffle so
  // that they are values held (or spilled) instead of special-register sequences re-derived at every use
  const int lane0 = threadIdx.x & 31;
  const int lane = __shfl_sync(FULL, lane0, lane0);
  const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(smem_raw);
  // (lane 0's value of a lane-dependent expression: a shuffle of a warp-uniform value would be folded away)
  const uint32_t sb = __shfl_sync(FULL, smem0 + (uint32_t)lane0, 0);  // staged scene tables (SMALL)
  const int wofs = (SMALL ? (int)SM_BYTES : cfg.shape_bytes) + warp * cfg.per_warp_bytes;
  const uint32_t wb = __shfl_sync(FULL, smem0 + (uint32_t)(wofs + lane0), 0);  // this warp's record stack
  const uint32_t curb = wb + (uint32_t)cfg.cap * (uint32_t)sizeof(ScatterRec);  // the partly consumed record
  constexpr bool MULTI_SLOT = ACC != ACC_REG;
  int* slot_rays = reinterpret_cast<int*>(smem_raw + wofs + (size_t)(cfg.cap + 1) * sizeof(ScatterRec));  // [32], multi-pixel tasks + RT_HIT_RAY_COUNT
  float* acc = reinterpret_cast<float*>(slot_rays + 32);   // ACC_SEG: [32][3]; ACC_LANES: [G][3][32]
  const bool count_rays = a.out_hit != nullptr && a.hit_mode == RT_HIT_RAY_COUNT;

  const PixelMap pm = make_pixel_map(a);
  const int N = a.num_of_rays;
  const int S2 = a.S > 0 ? a.S * a.S : 1;
  const int G = cfg.group, L = cfg.per_pixel;
  // x / N for 0 <= x <= 32 as a multiply-shift (exact for N <= 1024; beyond that x / N is 0 or 1)
  auto div_n = [&](int x) -> int { return cfg.n_magic ? (int)(((unsigned)x * cfg.n_magic) >> 16) : (x >= N ? 1 : 0); };
  const float inv_n = cfg.inv_n, inv_spp = cfg.inv_spp;
  // rays are counted per task from warp-uniform quantities (every lane holds the same sum)
  unsigned int n_rays = 0;
  bool overflow = false;

  while (true) {
    long long task = 0;
    if (lane == 0) task = (long long)atomicAdd(a.counters + CNT_TASK, 1ull);
    task = __shfl_sync(FULL, task, 0);
    if (task >= cfg.n_tasks) break;
    const long long p0 = task * G;
    // one pixel per task: the jitter stream is taken to the pixel's first draw once, lanes jump on by 2 s
    uint64_t aa_task = a.aa_state;
    if (ACC == ACC_REG && a.S > 0 && p0 < pm.n_pixels) {
      int col, row;
      pm.locate(p0, col, row);
      aa_task = pcg_jump(a.aa_state, 2ull * (unsigned long long)((long long)row * a.width + col) * S2, a.jump);
    }
    float sr = 0.f, sg = 0.f, sb_ = 0.f;  // lane-private sums (single-slot tasks)
    unsigned int task_rays = 0;           // warp-uniform
    const int task_prims = (int)min((long long)G, pm.n_pixels - p0) * L;  // primaries of this task
    if (MULTI_SLOT) {
      if (ACC == ACC_SEG) { for (int i = lane; i < 96; i += 32) acc[i] = 0.f; }
      else { for (int i = 0; i < 3 * cfg.group; ++i) acc[i * 32 + lane] = 0.f; }
      slot_rays[lane] = 0;
      __syncwarp();
    }

    // All samples of the task start first (cfg.rounds warp-wide primary iterations pushing onto one
    // stack), then the stack drains once: the partially filled iterations at the end of a drain are
    // paid once per task instead of once per 32 samples.
    {
      int top = 0, cur_rem = 0, cur_done = 0;
      int round = 0;
      bool primary_phase = true;
      while (true) {
        // ---------------- pick this lane's ray
        bool active;
        int slot = 0, depth = 0, seg_start = lane, origin = -1;
        V3<float> thr = mk3<float>(1.f, 1.f, 1.f);
        Ray<float> ray;
        Pcg rng;
        long long pix = 0;
        bool last_of_pixel = false;
        if (primary_phase) {
          const int j = round * 32 + lane;
          slot = j / L;
          const int ls = j - slot * L;
          const long long p = p0 + slot;
          active = (j < G * L) && (p < pm.n_pixels);
          task_rays += (unsigned)min(max(task_prims - round * 32, 0), 32);
          if (active) {
            int col, row;
            pm.locate(p, col, row);
            pix = (long long)row * a.width + col;
            const int s = own_stratum(a, ls);
            const unsigned long long k = (unsigned long long)pix * S2 + s;
            Pcg aa;
            aa.inc = a.aa_inc;
            if (ACC == ACC_REG) aa.state = (a.S > 0) ? pcg_jump(aa_task, 2ull * (unsigned)s, a.jump) : 0;
            else aa.state = (a.S > 0) ? pcg_jump(a.aa_state, 2ull * k, a.jump) : 0;
            ray = primary_ray_mul(a, cfg.inv_w, cfg.inv_h, cfg.inv_s, col, row, s, aa);
            rng = pcg_seed(a.pt_state, (a.pt_inc >> 1) + k);
            last_of_pixel = (ls == L - 1);
          }
          if (ACC == ACC_SEG) seg_start = max(0, slot * L - round * 32);
        } else {
          const long long avail = (long long)cur_rem + (long long)top * N;
          if (avail == 0) break;
          const int take = (int)min(avail, 32ll);
          active = lane < take;
          task_rays += (unsigned)take;
          // the first `from_cur` lanes finish the partly consumed record, the others take whole records
          // off the top, N lanes each: one address per lane, one set of three broadcast LDS.128
          const int from_cur = min(cur_rem, take);
          const int jj = max(lane - from_cur, 0);
          const int r = div_n(jj);
          const bool on_cur = lane < from_cur;
          const int child = on_cur ? cur_done + lane : jj - r * N;
          if (ACC == ACC_SEG) seg_start = on_cur ? 0 : from_cur + r * N;
          const uint32_t ra = (on_cur || !active) ? curb : wb + (uint32_t)(top - 1 - r) * (uint32_t)sizeof(ScatterRec);
          ScatterRec rec;
          rec.a = lds4(ra); rec.b = lds4(ra + 16); rec.c = lds4(ra + 32);
          // bookkeeping, identical in all lanes
          const int rest = take - from_cur;
          const int full = div_n(rest), part = rest - full * N;
          cur_rem -= from_cur;
          cur_done += from_cur;
          __syncwarp();  // every lane has read its record
          int new_top = top - full;
          if (part > 0) {  // the next record is only partly consumed: it becomes `cur`
            new_top -= 1;
            if (lane < 3) sts4(curb + 16 * lane, lds4(wb + (uint32_t)new_top * (uint32_t)sizeof(ScatterRec) + 16 * lane));
            cur_rem = N - part;
            cur_done = part;
          }
          __syncwarp();
          top = new_top;
          if (active) {
            const int meta = __float_as_int(rec.a.w);
            slot = meta & 31;
            depth = (meta >> 6) & 1023;
            origin = (int)((unsigned)meta >> 16);
            if (origin == 0xFFFF) origin = -1;
            thr = mk3<float>(rec.b.w, rec.c.x, rec.c.y);
            const uint64_t base = ((uint64_t)__float_as_uint(rec.c.w) << 32) | (uint64_t)__float_as_uint(rec.c.z);
            rng.state = child_stream(base, child);
            rng.inc = a.pt_inc;
            ray.o = mk3<float>(rec.a.x, rec.a.y, rec.a.z);
            ray.tmax = Num<float>::inf();
            const V3<float> nd = mk3<float>(rec.b.x, rec.b.y, rec.b.z);
            if (((meta >> 5) & 1) == RT_BRDF_DIFFUSE) {
              const float u1 = unit_from_u32((uint32_t)(rng.state >> 32));  // the two halves of the child's stream value
              const float u2 = unit_from_u32((uint32_t)rng.state);
              ray.d = diffuse_dir<float>(nd, u1, u2);
              ray.tmin = 1.0e-3f;
            } else {
              ray.d = nd;
              ray.tmin = 1e-5f;
            }
          }
        }

        // ---------------- trace + shade
        V3<float> contrib = mk3<float>(0.f, 0.f, 0.f);
        bool push = false;
        ScatterRec out;
        float best_t = Num<float>::inf();
        int best = -1;
        if (!SMALL && RT_WARP_SPLIT) {  // large scenes: the whole warp sweeps together, half-warps on different pairs
          if (!active) { ray.o = mk3<float>(0.f, 0.f, 0.f); ray.d = mk3<float>(0.f, 0.f, 0.f); ray.tmin = 0.f; ray.tmax = 0.f; }
          closest_all_warp(sc, src, ray, active, best_t, best, origin);
        }
        if (active) {
          if (SMALL) closest_small(sb, sc.n_spheres, sc.n_shapes, ray, origin, best_t, best);
          else if (!RT_WARP_SPLIT) closest_all_f32<true>(sc, src, ray, best_t, best, origin);
          const bool found = best >= 0;
          if (count_rays) {
            if (MULTI_SLOT) atomicAdd(&slot_rays[slot], 1);
          } else if (primary_phase) {  // (warp-uniform)
            if (last_of_pixel && a.out_hit)
              a.out_hit[pm.compact ? p0 + slot : pix] = !found ? -1 : (SMALL ? ldsic(sb + SM_ORIG + 4 * best) : sc.orig[best]);
          }
          if (!found) {
            contrib = mk3<float>(thr.x * cfg.bg[0], thr.y * cfg.bg[1], thr.z * cfg.bg[2]);
          } else {
            // The hit record is built lazily (rt_device.cuh local_hit / local_uv / world_frame): most rays
            // of a tree are its leaves, whose children render.py:100-101 cuts — for those only the emitted
            // colour matters, which for a uniform pigment needs no record at all.
            int brdf_kind, brdf_pig, emit_pig, mf;
            Rows3 im;
            const bool sphere = best < sc.n_spheres;
            if (SMALL) {
              const int4 mt = lds4ic(sb + SM_SHAPE_MAT + 16 * best);
              brdf_kind = mt.x; brdf_pig = mt.y; emit_pig = mt.z; mf = mt.w;
            } else {
              const DevMaterial& mat = sc.materials[sc.material[best]];
              brdf_kind = mat.brdf_kind; brdf_pig = mat.brdf_pigment; emit_pig = mat.emitted_pigment; mf = mat.flags;
            }
            auto rows_of = [&](bool inverse) -> Rows3 {
              if (SMALL) return lds_rows(sb + (inverse ? SM_INVM : SM_M) + 48 * best);
              return rows_at((inverse ? sc.invm : sc.m) + 12 * (size_t)best);
            };
            auto pig_uniform = [&](int idx) -> V3<float> {
              if (SMALL) { const float4 q = lds4c(sb + SM_PIG + idx * SM_PIG_STRIDE + 16); return mk3<float>(q.x, q.y, q.z); }
              return pig_c1<float>(sc.pigments[idx]);
            };
            auto pig_at = [&](int idx, float u, float v) -> V3<float> {
              if (SMALL) return pigment_color_small(sb + SM_PIG + idx * SM_PIG_STRIDE, u, v);
              return pigment_color<float>(sc.pigments, idx, u, v);
            };
            float u = 0.f, v = 0.f;
            if (depth >= a.max_depth) {
              // every child would come back BLACK, so neither the roulette draw (render.py:116-123) nor
              // the BRDF colour can change the result: emitted radiance only
              if (!(mf & MAT_EMIT_BLACK)) {
                if (mf & MAT_UV_EMIT) local_uv<float>(local_hit_rows(rows_of(true), ray, best_t), sphere, u, v);
                contrib = mul3(thr, (mf & MAT_UV_EMIT) ? pig_at(emit_pig, u, v) : pig_uniform(emit_pig));
              }
            } else {
              const bool scatters = !(mf & MAT_NO_SCATTER);
              LocalHit<float> lh;
              if (scatters || (mf & MAT_USES_UV)) { im = rows_of(true); lh = local_hit_rows(im, ray, best_t); }
              if (mf & MAT_USES_UV) local_uv<float>(lh, sphere, u, v);
              if (!(mf & MAT_EMIT_BLACK))
                contrib = mul3(thr, (mf & MAT_UV_EMIT) ? pig_at(emit_pig, u, v) : pig_uniform(emit_pig));
              if (scatters) {
                V3<float> hit_color = (mf & MAT_UV_BRDF) ? pig_at(brdf_pig, u, v) : pig_uniform(brdf_pig);
                const float lum = max3(hit_color);
                bool go_on = true;
                if (depth >= a.rr_limit) {  // render.py:116-123
                  const float q = fmaxf(0.05f, 1.f - lum);
                  if (pcg_random_float<float>(rng) > q) hit_color = fast_rcp(1.f - q) * hit_color;
                  else go_on = false;
                }
                if (go_on && lum > 0.f) {
                  push = true;
                  V3<float> point, normal;
                  world_frame_rows(im, rows_of(false), lh, sphere, point, normal);
                  const V3<float> nd = (brdf_kind == RT_BRDF_DIFFUSE) ? normal : specular_dir<float>(ray.d, normal);
                  const V3<float> w = inv_n * mul3(thr, hit_color);
                  out.a = make_float4(point.x, point.y, point.z,
                                      __int_as_float(slot | (brdf_kind << 5) | ((depth + 1) << 6) |
                                                     ((best < 0xFFFF ? best : 0xFFFF) << 16)));
                  out.b = make_float4(nd.x, nd.y, nd.z, w.x);
                  out.c = make_float4(w.y, w.z, __uint_as_float((uint32_t)rng.state), __uint_as_float((uint32_t)(rng.state >> 32)));
                }
              }
            }
          }
        }

        // ---------------- accumulate
        if (ACC == ACC_SEG) {
          // lanes of one record (or one pixel, for primaries) are contiguous: segmented scan
#pragma unroll
          for (int d = 1; d < 32; d <<= 1) {
            const float r = __shfl_up_sync(FULL, contrib.x, d);
            const float g = __shfl_up_sync(FULL, contrib.y, d);
            const float b = __shfl_up_sync(FULL, contrib.z, d);
            if (lane - d >= seg_start) { contrib.x += r; contrib.y += g; contrib.z += b; }
          }
          const int next_start = __shfl_down_sync(FULL, seg_start, 1);
          const bool tail = (lane == 31) || (next_start != seg_start);
          if (active && tail && (contrib.x != 0.f || contrib.y != 0.f || contrib.z != 0.f)) {
            atomicAdd(&acc[3 * slot + 0], contrib.x);
            atomicAdd(&acc[3 * slot + 1], contrib.y);
            atomicAdd(&acc[3 * slot + 2], contrib.z);
          }
        } else if (ACC == ACC_LANES) {
          if (active) {  // this lane's own column of the pixel's accumulator: no other lane touches it
            float* col = acc + (3 * slot) * 32 + lane;
            col[0] += contrib.x; col[32] += contrib.y; col[64] += contrib.z;
          }
        } else {
          sr += contrib.x; sg += contrib.y; sb_ += contrib.z;
        }

        // ---------------- push the new records: one ballot gives every lane its slot
        const unsigned pmask = __ballot_sync(FULL, push);
        const int npush = __popc(pmask);
        if (top + npush > cfg.cap) {
          overflow = true;
        } else {
          if (push) {
            const uint32_t pa = wb + (uint32_t)(top + __popc(pmask & ((1u << lane) - 1u))) * (uint32_t)sizeof(ScatterRec);
            sts4(pa, out.a); sts4(pa + 16, out.b); sts4(pa + 32, out.c);
          }
          top += npush;
        }
        __syncwarp();
        if (primary_phase && ++round >= cfg.rounds) primary_phase = false;
      }
    }
    n_rays += task_rays;

    // ---------------- write the pixels of this task
    if (ACC == ACC_SEG) {
      __syncwarp();
      const long long p = p0 + lane;
      if (lane < G && p < pm.n_pixels) {
        int col, row;
        pm.locate(p, col, row);
        store_pixel<float>(a, pm.at(p, col, row),
                           mk3<float>(acc[3 * lane] * inv_spp, acc[3 * lane + 1] * inv_spp, acc[3 * lane + 2] * inv_spp));
        if (count_rays) a.out_hit[pm.at(p, col, row)] = slot_rays[lane];
      }
      __syncwarp();
    } else if (ACC == ACC_LANES) {
      __syncwarp();
      for (int g = 0; g < G; ++g) {
        float r = acc[(3 * g) * 32 + lane], gr = acc[(3 * g + 1) * 32 + lane], b = acc[(3 * g + 2) * 32 + lane];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
          r += __shfl_xor_sync(FULL, r, d);
          gr += __shfl_xor_sync(FULL, gr, d);
          b += __shfl_xor_sync(FULL, b, d);
        }
        const long long p = p0 + g;
        if (lane == 0 && p < pm.n_pixels) {
          int col, row;
          pm.locate(p, col, row);
          store_pixel<float>(a, pm.at(p, col, row), mk3<float>(r * inv_spp, gr * inv_spp, b * inv_spp));
          if (count_rays) a.out_hit[pm.at(p, col, row)] = slot_rays[g];
        }
      }
      __syncwarp();
    } else {
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        sr += __shfl_xor_sync(FULL, sr, d);
        sg += __shfl_xor_sync(FULL, sg, d);
        sb_ += __shfl_xor_sync(FULL, sb_, d);
      }
      if (lane == 0 && p0 < pm.n_pixels) {
        int col, row;
        pm.locate(p0, col, row);
        store_pixel<float>(a, pm.at(p0, col, row), mk3<float>(sr * inv_spp, sg * inv_spp, sb_ * inv_spp));
        if (count_rays) a.out_hit[pm.at(p0, col, row)] = (int)task_rays;
      }
    }
  }
  if (overflow) atomicExch(a.counters + CNT_OVERFLOW, 1ull);
  if (lane == 0 && n_rays) atomicAdd(a.counters + CNT_CLOSEST, (unsigned long long)n_rays);
  if (blockIdx.x == 0 && threadIdx.x == 0) atomicAdd(a.counters + CNT_SAMPLES, cfg.n_samples);
}

inline cudaError_t launch_pt_warp_impl(const SceneView<float>& sc, const RenderArgs& a, cudaStream_t st,
                                       int sm_count, LaunchInfo* info, const char** why_not) {
  PixelMap pm = make_pixel_map(a);
  const int S2 = a.S > 0 ? a.S * a.S : 1;
  int L = S2;
  if (a.part_mode == RT_PART_SPP && a.part_count > 1)
    L = a.part_rank < S2 ? (S2 - a.part_rank + a.part_count - 1) / a.part_count : 0;
  if (pm.n_pixels == 0 || L == 0) return cudaSuccess;
  if (a.num_of_rays < 1) { *why_not = "num_of_rays must be >= 1"; return cudaErrorInvalidValue; }
  if (a.max_depth >= 1023) { *why_not = "max_depth >= 1023 is not supported by the warp variant (use mega)"; return cudaErrorInvalidValue; }
  WarpCfg cfg;
  cfg.per_pixel = L;
  // Samples per task: all strata of one pixel when there are at least 32 of them (demo.txt at 64 spp:
  // 64-sample tasks, lane-private sums), otherwise as many pixels as make one warp-full of samples.
  size_t shape_bytes = (size_t)sc.n_pairs * 96 + (size_t)(sc.n_shapes - sc.n_spheres) * 48;
  const bool shapes_smem = shape_bytes > 0 && shape_bytes <= 64 * 1024;
  const bool small = sc.n_shapes > 0 && sc.n_spheres <= RT_SMALL_MAX_SPHERES && sc.n_shapes <= RT_SMALL_MAX_SHAPES &&
                     sc.n_materials <= RT_SMALL_MAX_MATERIALS && sc.n_pigments <= RT_SMALL_MAX_PIGMENTS;
  // (multi-pixel tasks, L < 32: 32 samples per task keep the accumulator columns small enough for three
  // resident blocks per SM — measured on demo.txt split over 8 GPUs: 50.9 vs 41.4 Grays/s per GPU)
  int want = 32;
#ifdef RT_TUNING  // tuning builds only (-DRT_TUNING): samples per multi-pixel task, clamped to what the kernel supports
  if (const char* env = getenv("RT_WARP_WANT")) want = std::min(std::max(atoi(env), 1), 256);
#endif
  if (L >= 32) cfg.group = 1;
  else if (L >= 4) cfg.group = (want + L - 1) / L < RT_ACC_LANES_MAX_GROUP ? (want + L - 1) / L : RT_ACC_LANES_MAX_GROUP;
  else cfg.group = 32 / L;
  if (cfg.group * L > 32 && !small) cfg.group = 32 / L > 0 ? 32 / L : 1;
  cfg.rounds = (cfg.group * L + 31) / 32;
  cfg.n_tasks = (pm.n_pixels + cfg.group - 1) / cfg.group;
  // stack capacity: every primary of the task may leave one level-1 record, and at most ~32 records
  // per deeper tree level are alive at any time; with N == 1 a record is replaced
  // by at most one record
  const long long prims = (long long)cfg.rounds * 32;
  long long cap = a.num_of_rays == 1 ? prims + 32 : prims + 32ll * (long long)a.max_depth;
  if (cap < 64) cap = 64;
  int acc_mode = cfg.group == 1 ? ACC_REG : (cfg.group <= RT_ACC_LANES_MAX_GROUP ? ACC_LANES : ACC_SEG);
  cfg.shape_bytes = small ? (int)SM_BYTES : (shapes_smem ? (int)((shape_bytes + 15) / 16 * 16) : 0);
  const size_t limit = 200 * 1024;
  int warps = (small ? RT_SMALL_THREADS : RT_WARP_MAX_THREADS) / 32;
  size_t per_warp = 0, smem = 0;
  auto footprint = [&](int mode, int w, size_t& pw) {
    pw = (size_t)(cap + 1) * sizeof(ScatterRec);
    if (mode == ACC_SEG) pw += 32 * sizeof(int) + 96 * sizeof(float);
    if (mode == ACC_LANES) pw += 32 * sizeof(int) + (size_t)cfg.group * 96 * sizeof(float);
    return cfg.shape_bytes + pw * w;
  };
  // The accumulator columns of ACC_LANES cost 384 B per pixel and warp; where that takes a resident
  // block away from the SM (demo.txt split over 4 or 8 GPUs: 2 blocks instead of 3, -25 % throughput)
  // the segmented-scan accumulators (ACC_SEG, 512 B per warp in all) are the better trade.
  if (acc_mode == ACC_LANES) {
    size_t pw_l, pw_s;
    const size_t sm_bytes = 227 * 1024;
    const size_t blocks_lanes = sm_bytes / (footprint(ACC_LANES, warps, pw_l) + 1024);
    const size_t blocks_seg = sm_bytes / (footprint(ACC_SEG, warps, pw_s) + 1024);
    const size_t by_threads = 2048 / (warps * 32), by_regs = small ? RT_SMALL_MINB : 2;
    if (std::min(std::min(blocks_seg, by_threads), by_regs) > std::min(std::min(blocks_lanes, by_threads), by_regs)) acc_mode = ACC_SEG;
#ifdef RT_TUNING
    if (const char* env = getenv("RT_WARP_ACC")) { const int m = atoi(env); if (m == ACC_LANES || m == ACC_SEG) acc_mode = m; }
#endif
  }
  for (;; warps >>= 1) {
    smem = footprint(acc_mode, warps, per_warp);
    if (smem <= limit || warps == 1) break;
  }
  if (smem > limit) { *why_not = "max_depth needs a deeper work stack than shared memory holds"; return cudaErrorInvalidValue; }
  cfg.cap = (int)cap;
  cfg.per_warp_bytes = (int)per_warp;
  for (int k = 0; k < 3; ++k) cfg.bg[k] = (float)a.background[k];
  cfg.inv_n = 1.0f / (float)a.num_of_rays;
  cfg.inv_spp = 1.0f / (float)S2;
  cfg.n_magic = a.num_of_rays <= 1024 ? (65536u + (unsigned)a.num_of_rays - 1u) / (unsigned)a.num_of_rays : 0u;
  cfg.n_samples = (unsigned long long)pm.n_pixels * (unsigned long long)L;
  cfg.inv_w = 1.0 / (double)a.width;
  cfg.inv_h = 1.0 / (double)a.height;
  cfg.inv_s = a.S > 0 ? 1.0 / (double)a.S : 1.0;

  void (*kern)(const SceneView<float>, const RenderArgs, const WarpCfg);
#define RT_PICK(ACCM)                                                                   \
  (small ? k_pt_warp<true, ACCM, true> : (shapes_smem ? k_pt_warp<true, ACCM, false> : k_pt_warp<false, ACCM, false>))
  if (acc_mode == ACC_REG) kern = RT_PICK(ACC_REG);
  else if (acc_mode == ACC_LANES) kern = RT_PICK(ACC_LANES);
  else kern = RT_PICK(ACC_SEG);
#undef RT_PICK
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)limit);
  if (e != cudaSuccess) return e;
  int per_sm = 0;
  e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, warps * 32, smem);
  if (e != cudaSuccess) return e;
  if (per_sm < 1) per_sm = 1;
  long long blocks = (long long)per_sm * sm_count;  // persistent: every block stays resident
  long long needed = (cfg.n_tasks + warps - 1) / warps;
  if (blocks > needed) blocks = needed;
  kern<<<(unsigned)blocks, warps * 32, smem, st>>>(sc, a, cfg);
  if (info) { info->n_launches += 1; info->variant = RT_VARIANT_WARP; }
  return cudaGetLastError();
}
