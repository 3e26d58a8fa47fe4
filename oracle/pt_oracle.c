/*
 * pt_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A sequential fp64 CPU restatement of the rendering hot path of ziotom78/pytracer
 * (ImageTracer.fire_all_rays driving OnOff/Flat/PathTracer/PointLight renderers), written
 * from the reference's formulas so that every floating-point operation happens in the same
 * order as in the Python code (Python floats are C doubles, `math.*` is libm, `x**2` is
 * pow(x, 2.0)). Compiled with -ffp-contract=off so that no multiply-add is fused.
 * Parity is PINNED: tests/golden/make_golden.py runs the unmodified reference in the build
 * container and tests/test_oracle_golden.py checks this file against those outputs
 * bit for bit (images, hit indices, ray counts, PCG states).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may load this library. The product (pytracer_b200/) never does.
 *
 * Every function cites the reference file:line it restates (paths relative to
 * /root/reference/src/pytracer/).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "../include/rt_api.h"

typedef struct { double x, y, z; } v3;
typedef struct { double r, g, b; } col3;
typedef struct { v3 o, d; double tmin, tmax; int depth; } ray_t;
typedef struct { uint64_t state, inc; } pcg_t;

/* ------------------------------------------------------------------ pcg.py:22-62 */
static uint32_t pcg_random(pcg_t* p) {                       /* pcg.py:43-58 */
  uint64_t old = p->state;
  p->state = old * 6364136223846793005ULL + p->inc;
  uint32_t xorshifted = (uint32_t)(((old >> 18) ^ old) >> 27);
  uint32_t rot = (uint32_t)(old >> 59);
  return (xorshifted >> rot) | (xorshifted << ((-rot) & 31));
}
static double pcg_random_float(pcg_t* p) {                   /* pcg.py:60-62 */
  return (double)pcg_random(p) / 4294967295.0;
}
static void pcg_seed(pcg_t* p, uint64_t init_state, uint64_t init_seq) { /* pcg.py:29-41 */
  p->state = 0;
  p->inc = (init_seq << 1) | 1;
  pcg_random(p);
  p->state += init_state;
  pcg_random(p);
}

/* ------------------------------------------------------------------ geometry.py */
static double sq(double x) { return pow(x, 2.0); }           /* `x**2` in Vec.squared_norm */
static double vec_squared_norm(v3 a) {                       /* geometry.py:111-115 */
  return sq(a.x) + sq(a.y) + sq(a.z);
}
static double vec_norm(v3 a) { return sqrt(vec_squared_norm(a)); } /* geometry.py:117-119 */
static v3 vec_normalize(v3 a) {                              /* geometry.py:129-135 */
  double n = vec_norm(a);
  v3 r = { a.x / n, a.y / n, a.z / n };
  return r;
}
static double normal_norm(v3 a) {                            /* geometry.py:214-218, x*x form */
  return sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
}
static double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; } /* :107-109 */
static v3 neg(v3 a) { v3 r = { -a.x, -a.y, -a.z }; return r; }
static v3 scale(double s, v3 a) { v3 r = { s * a.x, s * a.y, s * a.z }; return r; } /* :33-34 */
static v3 add(v3 a, v3 b) { v3 r = { a.x + b.x, a.y + b.y, a.z + b.z }; return r; }
static v3 sub(v3 a, v3 b) { v3 r = { a.x - b.x, a.y - b.y, a.z - b.z }; return r; }

static double normalized_dot(v3 a, v3 b) {                   /* geometry.py:265-276 */
  return dot(vec_normalize(a), vec_normalize(b));
}

static void create_onb_from_z(v3 n, v3* e1, v3* e2, v3* e3) { /* geometry.py:247-262 */
  double sign = (n.z > 0.0) ? 1.0 : -1.0;
  double a = -1.0 / (sign + n.z);
  double b = n.x * n.y * a;
  e1->x = 1.0 + sign * n.x * n.x * a; e1->y = sign * b; e1->z = -sign * n.x;
  e2->x = b; e2->y = sign + n.y * n.y * a; e2->z = -n.y;
  *e3 = n;
}

/* ------------------------------------------------------------------ transformations.py:58-86 */
static v3 xf_point(const double* m, v3 p) {                  /* :66-77 (w == 1 for affine) */
  v3 r = { p.x * m[0] + p.y * m[1] + p.z * m[2] + m[3],
           p.x * m[4] + p.y * m[5] + p.z * m[6] + m[7],
           p.x * m[8] + p.y * m[9] + p.z * m[10] + m[11] };
  return r;
}
static v3 xf_vec(const double* m, v3 v) {                    /* :59-65 */
  v3 r = { v.x * m[0] + v.y * m[1] + v.z * m[2],
           v.x * m[4] + v.y * m[5] + v.z * m[6],
           v.x * m[8] + v.y * m[9] + v.z * m[10] };
  return r;
}
static v3 xf_normal(const double* invm, v3 n) {              /* :78-86, transpose of the inverse */
  v3 r = { n.x * invm[0] + n.y * invm[4] + n.z * invm[8],
           n.x * invm[1] + n.y * invm[5] + n.z * invm[9],
           n.x * invm[2] + n.y * invm[6] + n.z * invm[10] };
  return r;
}

/* ------------------------------------------------------------------ ray.py:52-69 */
static v3 ray_at(const ray_t* r, double t) { return add(r->o, scale(t, r->d)); }
static ray_t ray_transform(const ray_t* r, const double* m) {
  ray_t q = *r;
  q.o = xf_point(m, r->o);
  q.d = xf_vec(m, r->d);
  return q;
}

/* ------------------------------------------------------------------ shapes.py */
typedef struct {
  int shape;            /* index in World.shapes */
  double t;
  v3 world_point, normal;
  double u, v;
} hit_t;

static int sphere_intersection(const rt_scene_desc* s, int i, const ray_t* ray, hit_t* h) {
  /* shapes.py:97-131 */
  const double* m = s->shape_m + 12 * (size_t)i;
  const double* invm = s->shape_invm + 12 * (size_t)i;
  ray_t inv = ray_transform(ray, invm);
  v3 ov = inv.o;
  double a = vec_squared_norm(inv.d);
  double b = 2.0 * dot(ov, inv.d);
  double c = vec_squared_norm(ov) - 1.0;
  double delta = b * b - 4.0 * a * c;
  if (delta <= 0.0) return 0;
  double sd = sqrt(delta);
  double t1 = (-b - sd) / (2.0 * a);
  double t2 = (-b + sd) / (2.0 * a);
  double t;
  if (t1 > inv.tmin && t1 < inv.tmax) t = t1;
  else if (t2 > inv.tmin && t2 < inv.tmax) t = t2;
  else return 0;
  v3 hp = ray_at(&inv, t);
  h->shape = i;
  h->t = t;
  h->world_point = xf_point(m, hp);
  v3 n = (dot(hp, inv.d) < 0.0) ? hp : neg(hp);             /* shapes.py:45-54 */
  h->normal = xf_normal(invm, n);
  double u = atan2(hp.y, hp.x) / (2.0 * M_PI);               /* shapes.py:36-42 */
  h->u = (u >= 0.0) ? u : u + 1.0;
  h->v = acos(hp.z) / M_PI;
  return 1;
}

static int sphere_quick(const rt_scene_desc* s, int i, const ray_t* ray) { /* shapes.py:133-151 */
  const double* invm = s->shape_invm + 12 * (size_t)i;
  ray_t inv = ray_transform(ray, invm);
  v3 ov = inv.o;
  double a = vec_squared_norm(inv.d);
  double b = 2.0 * dot(ov, inv.d);
  double c = vec_squared_norm(ov) - 1.0;
  double delta = b * b - 4.0 * a * c;
  if (delta <= 0.0) return 0;
  double sd = sqrt(delta);
  double t1 = (-b - sd) / (2.0 * a);
  double t2 = (-b + sd) / (2.0 * a);
  return (inv.tmin < t1 && t1 < inv.tmax) || (inv.tmin < t2 && t2 < inv.tmax);
}

static int plane_intersection(const rt_scene_desc* s, int i, const ray_t* ray, hit_t* h) {
  /* shapes.py:163-189 */
  const double* m = s->shape_m + 12 * (size_t)i;
  const double* invm = s->shape_invm + 12 * (size_t)i;
  ray_t inv = ray_transform(ray, invm);
  if (fabs(inv.d.z) < 1e-5) return 0;
  double t = -inv.o.z / inv.d.z;
  if (t <= inv.tmin || t >= inv.tmax) return 0;
  v3 hp = ray_at(&inv, t);
  h->shape = i;
  h->t = t;
  h->world_point = xf_point(m, hp);
  v3 n = { 0.0, 0.0, (inv.d.z < 0.0) ? 1.0 : -1.0 };
  h->normal = xf_normal(invm, n);
  h->u = hp.x - floor(hp.x);
  h->v = hp.y - floor(hp.y);
  return 1;
}

static int plane_quick(const rt_scene_desc* s, int i, const ray_t* ray) { /* shapes.py:191-198 */
  const double* invm = s->shape_invm + 12 * (size_t)i;
  ray_t inv = ray_transform(ray, invm);
  if (fabs(inv.d.z) < 1e-5) return 0;
  double t = -inv.o.z / inv.d.z;
  return inv.tmin < t && t < inv.tmax;
}

/* ------------------------------------------------------------------ world.py */
typedef struct { uint64_t closest, shadow, samples; } counters_t;

static int world_ray_intersection(const rt_scene_desc* s, const ray_t* ray, hit_t* out,
                                  counters_t* cnt) {         /* world.py:51-69 */
  int found = 0;
  hit_t h;
  if (cnt) cnt->closest++;
  for (int i = 0; i < s->n_shapes; ++i) {
    int ok = (s->shape_kind[i] == RT_SHAPE_SPHERE) ? sphere_intersection(s, i, ray, &h)
                                                   : plane_intersection(s, i, ray, &h);
    if (!ok) continue;
    if (!found || h.t < out->t) { *out = h; found = 1; }
  }
  if (found) {
    double n = normal_norm(out->normal);                     /* Normal.normalize geometry.py:220-226 */
    out->normal.x /= n; out->normal.y /= n; out->normal.z /= n;
  }
  return found;
}

static int world_is_point_visible(const rt_scene_desc* s, v3 point, v3 observer,
                                  counters_t* cnt) {         /* world.py:71-80 */
  v3 direction = sub(point, observer);
  double dir_norm = vec_norm(direction);
  ray_t ray = { observer, direction, 1e-2 / dir_norm, 1.0, 0 };
  if (cnt) cnt->shadow++;
  for (int i = 0; i < s->n_shapes; ++i) {
    int blocked = (s->shape_kind[i] == RT_SHAPE_SPHERE) ? sphere_quick(s, i, &ray)
                                                        : plane_quick(s, i, &ray);
    if (blocked) return 0;
  }
  return 1;
}

/* ------------------------------------------------------------------ materials.py */
static col3 pigment_get_color(const rt_scene_desc* s, int pig, double u, double v) {
  const rt_pigment* p = &s->pigments[pig];
  col3 c;
  if (p->kind == RT_PIGMENT_UNIFORM) {                       /* materials.py:58-59 */
    c.r = p->color1[0]; c.g = p->color1[1]; c.b = p->color1[2];
  } else if (p->kind == RT_PIGMENT_CHECKERED) {              /* materials.py:96-100 */
    long long iu = (long long)floor(u * p->num_of_steps);
    long long iv = (long long)floor(v * p->num_of_steps);
    const double* k = ((iu & 1) == (iv & 1)) ? p->color1 : p->color2;
    c.r = k[0]; c.g = k[1]; c.b = k[2];
  } else {                                                   /* materials.py:70-82 */
    long long col = (long long)(u * p->tex_width);
    long long row = (long long)(v * p->tex_height);
    if (col >= p->tex_width) col = p->tex_width - 1;
    if (row >= p->tex_height) row = p->tex_height - 1;
    const double* t = s->texels + 3 * ((size_t)p->tex_offset + (size_t)row * p->tex_width + (size_t)col);
    c.r = t[0]; c.g = t[1]; c.b = t[2];
  }
  return c;
}

static col3 brdf_eval(const rt_scene_desc* s, const rt_material* mat, v3 normal, v3 in_dir,
                      v3 out_dir, double u, double v) {
  col3 c = pigment_get_color(s, mat->brdf_pigment, u, v);
  if (mat->brdf_kind == RT_BRDF_DIFFUSE) {                   /* materials.py:129-130 */
    double k = 1.0 / M_PI;
    c.r = c.r * k; c.g = c.g * k; c.b = c.b * k;
    return c;
  }
  /* SpecularBRDF.eval materials.py:164-173 */
  double theta_in = acos(normalized_dot(normal, in_dir));
  double theta_out = acos(normalized_dot(normal, out_dir));
  if (fabs(theta_in - theta_out) < mat->threshold_angle_rad) return c;
  c.r = c.g = c.b = 0.0;
  return c;
}

static ray_t brdf_scatter_ray(const rt_material* mat, pcg_t* pcg, v3 incoming_dir, v3 point,
                              v3 normal, int depth) {
  ray_t r;
  r.o = point; r.tmax = INFINITY; r.depth = depth;
  if (mat->brdf_kind == RT_BRDF_DIFFUSE) {                   /* materials.py:132-152 */
    v3 e1, e2, e3;
    create_onb_from_z(normal, &e1, &e2, &e3);
    double cos_theta_sq = pcg_random_float(pcg);
    double cos_theta = sqrt(cos_theta_sq), sin_theta = sqrt(1.0 - cos_theta_sq);
    double phi = 2.0 * M_PI * pcg_random_float(pcg);
    v3 a = scale(cos_theta, scale(cos(phi), e1));
    v3 b = scale(cos_theta, scale(sin(phi), e2));
    v3 c = scale(sin_theta, e3);
    r.d = add(add(a, b), c);
    r.tmin = 1.0e-3;
  } else {                                                   /* materials.py:175-196 */
    v3 rd = vec_normalize(incoming_dir);
    v3 n = vec_normalize(normal);
    double dp = dot(n, rd);
    r.d = sub(rd, scale(dp, scale(2.0, n)));
    r.tmin = 1e-5;
  }
  return r;
}

/* ------------------------------------------------------------------ render.py */
typedef struct {
  const rt_scene_desc* s;
  const rt_render_params* p;
  pcg_t* pt_pcg;
  counters_t* cnt;
  int last_shape;   /* shape hit by the last depth-0 ray (for the hit-index image) */
} ctx_t;

static col3 onoff_call(ctx_t* c, const ray_t* ray) {         /* render.py:52-53 */
  hit_t h;
  const double* k = c->p->background;
  c->last_shape = -1;
  if (world_ray_intersection(c->s, ray, &h, c->cnt)) { k = c->p->onoff_color; c->last_shape = h.shape; }
  col3 r = { k[0], k[1], k[2] };
  return r;
}

static col3 flat_call(ctx_t* c, const ray_t* ray) {          /* render.py:65-74 */
  hit_t h;
  c->last_shape = -1;
  if (!world_ray_intersection(c->s, ray, &h, c->cnt)) {
    col3 r = { c->p->background[0], c->p->background[1], c->p->background[2] };
    return r;
  }
  c->last_shape = h.shape;
  const rt_material* mat = &c->s->materials[c->s->shape_material[h.shape]];
  col3 a = pigment_get_color(c->s, mat->brdf_pigment, h.u, h.v);
  col3 e = pigment_get_color(c->s, mat->emitted_pigment, h.u, h.v);
  col3 r = { a.r + e.r, a.g + e.g, a.b + e.b };
  return r;
}

static col3 pathtracer_call(ctx_t* c, const ray_t* ray) {    /* render.py:99-139 */
  col3 black = { 0.0, 0.0, 0.0 };
  if (ray->depth > c->p->max_depth) return black;
  hit_t h;
  int found = world_ray_intersection(c->s, ray, &h, c->cnt);
  if (ray->depth == 0) c->last_shape = found ? h.shape : -1;
  if (!found) {
    col3 r = { c->p->background[0], c->p->background[1], c->p->background[2] };
    return r;
  }
  const rt_material* mat = &c->s->materials[c->s->shape_material[h.shape]];
  col3 hit_color = pigment_get_color(c->s, mat->brdf_pigment, h.u, h.v);
  col3 emitted = pigment_get_color(c->s, mat->emitted_pigment, h.u, h.v);
  double lum = fmax(fmax(hit_color.r, hit_color.g), hit_color.b);
  if (ray->depth >= c->p->rr_limit) {                        /* render.py:116-123 */
    double q = fmax(0.05, 1 - lum);
    if (pcg_random_float(c->pt_pcg) > q) {
      double k = 1.0 / (1.0 - q);
      hit_color.r = hit_color.r * k; hit_color.g = hit_color.g * k; hit_color.b = hit_color.b * k;
    } else {
      return emitted;
    }
  }
  col3 cum = black;
  if (lum > 0.0) {                                           /* render.py:126-137 */
    for (int i = 0; i < c->p->num_of_rays; ++i) {
      ray_t nr = brdf_scatter_ray(mat, c->pt_pcg, ray->d, h.world_point, h.normal, ray->depth + 1);
      col3 rad = pathtracer_call(c, &nr);
      cum.r = cum.r + hit_color.r * rad.r;
      cum.g = cum.g + hit_color.g * rad.g;
      cum.b = cum.b + hit_color.b * rad.b;
    }
  }
  double k = 1.0 / c->p->num_of_rays;                        /* render.py:139 */
  col3 r = { emitted.r + cum.r * k, emitted.g + cum.g * k, emitted.b + cum.b * k };
  return r;
}

static col3 pointlight_call(ctx_t* c, const ray_t* ray) {    /* render.py:157-193 */
  hit_t h;
  c->last_shape = -1;
  if (!world_ray_intersection(c->s, ray, &h, c->cnt)) {
    col3 r = { c->p->background[0], c->p->background[1], c->p->background[2] };
    return r;
  }
  c->last_shape = h.shape;
  const rt_scene_desc* s = c->s;
  const rt_material* mat = &s->materials[s->shape_material[h.shape]];
  col3 e = pigment_get_color(s, mat->emitted_pigment, h.u, h.v);
  col3 res = { c->p->ambient[0] + e.r, c->p->ambient[1] + e.g, c->p->ambient[2] + e.b };
  for (int l = 0; l < s->n_lights; ++l) {
    const rt_light* L = &s->lights[l];
    v3 lp = { L->position[0], L->position[1], L->position[2] };
    if (!world_is_point_visible(s, lp, h.world_point, c->cnt)) continue;
    v3 distance_vec = sub(h.world_point, lp);
    double distance = vec_norm(distance_vec);
    v3 in_dir = scale(1.0 / distance, distance_vec);
    double cos_theta = fmax(0.0, normalized_dot(neg(in_dir), h.normal));
    double distance_factor = (L->linear_radius > 0) ? sq(L->linear_radius / distance) : 1.0;
    col3 b = brdf_eval(s, mat, h.normal, in_dir, neg(ray->d), h.u, h.v);
    res.r = res.r + b.r * L->color[0] * cos_theta * distance_factor;
    res.g = res.g + b.g * L->color[1] * cos_theta * distance_factor;
    res.b = res.b + b.b * L->color[2] * cos_theta * distance_factor;
  }
  return res;
}

static col3 renderer_call(ctx_t* c, const ray_t* ray) {
  switch (c->p->algorithm) {
    case RT_ALGO_ONOFF: return onoff_call(c, ray);
    case RT_ALGO_FLAT: return flat_call(c, ray);
    case RT_ALGO_PATHTRACING: return pathtracer_call(c, ray);
    default: return pointlight_call(c, ray);
  }
}

/* ------------------------------------------------------------------ camera.py, imagetracer.py */
static ray_t camera_fire_ray(const rt_camera* cam, double u, double v) {
  ray_t r;
  r.tmin = 1.0e-5; r.tmax = INFINITY; r.depth = 0;
  if (cam->kind == RT_CAMERA_PERSPECTIVE) {                  /* camera.py:103-124 */
    r.o.x = -cam->screen_distance; r.o.y = 0.0; r.o.z = 0.0;
    r.d.x = cam->screen_distance; r.d.y = (1.0 - 2 * u) * cam->aspect_ratio; r.d.z = 2 * v - 1;
  } else {                                                   /* camera.py:59-78 */
    r.o.x = -1.0; r.o.y = (1.0 - 2 * u) * cam->aspect_ratio; r.o.z = 2 * v - 1;
    r.d.x = 1.0; r.d.y = 0.0; r.d.z = 0.0;
  }
  return ray_transform(&r, cam->m);
}

static ray_t tracer_fire_ray(const rt_render_params* p, int col, int row, double u_pixel,
                             double v_pixel) {               /* imagetracer.py:48-58 */
  double u = (col + u_pixel) / p->width;
  double v = 1.0 - (row + v_pixel) / p->height;
  return camera_fire_ray(&p->camera, u, v);
}

/*
 * ImageTracer.fire_all_rays (imagetracer.py:60-110) restricted to rows row_begin, row_begin +
 * row_step, ... < row_end (row_step = 1: the reference's loop; > 1: interleaved rows for the
 * multi-threaded CPU baseline, whose threads must be load-balanced).
 *   rgb          double[H][W][3], only the traced rows are written
 *   hit_index    optional int32[H][W]
 *   counters     uint64[3] = closest-hit queries, shadow queries, samples (accumulated)
 *   aa_io, pt_io {state, inc} of ImageTracer.pcg / PathTracer.pcg, advanced in place
 *   sample_states optional uint64[(row_end-row_begin)*W*max(1,S)^2]: PathTracer.pcg.state at
 *                the start of every sample (what RT_RNG_REPLAY consumes)
 */
int orc_render(const rt_scene_desc* s, const rt_render_params* p, int row_begin, int row_end,
               int row_step, double* rgb, int32_t* hit_index, uint64_t* counters, uint64_t* aa_io,
               uint64_t* pt_io, uint64_t* sample_states) {
  pcg_t aa = { aa_io[0], aa_io[1] };
  pcg_t pt = { pt_io[0], pt_io[1] };
  counters_t cnt = { 0, 0, 0 };
  ctx_t c = { s, p, &pt, &cnt, -1 };
  int S = p->samples_per_side;
  size_t k = 0;
  for (int row = row_begin; row < row_end; row += (row_step > 0 ? row_step : 1)) {
    for (int col = 0; col < p->width; ++col) {
      col3 px;
      if (S > 0) {
        col3 cum = { 0.0, 0.0, 0.0 };
        for (int ir = 0; ir < S; ++ir) {
          for (int ic = 0; ic < S; ++ic) {
            double u_pixel = (ic + pcg_random_float(&aa)) / S;
            double v_pixel = (ir + pcg_random_float(&aa)) / S;
            ray_t ray = tracer_fire_ray(p, col, row, u_pixel, v_pixel);
            if (sample_states) sample_states[k] = pt.state;
            ++k;
            cnt.samples++;
            col3 v = renderer_call(&c, &ray);
            cum.r = cum.r + v.r; cum.g = cum.g + v.g; cum.b = cum.b + v.b;
          }
        }
        double w = 1.0 / (double)((long long)S * S);
        px.r = cum.r * w; px.g = cum.g * w; px.b = cum.b * w;
      } else {
        ray_t ray = tracer_fire_ray(p, col, row, 0.5, 0.5);
        if (sample_states) sample_states[k] = pt.state;
        ++k;
        cnt.samples++;
        px = renderer_call(&c, &ray);
      }
      size_t off = (size_t)row * p->width + col;
      rgb[3 * off + 0] = px.r; rgb[3 * off + 1] = px.g; rgb[3 * off + 2] = px.b;
      if (hit_index) hit_index[off] = c.last_shape;
    }
  }
  aa_io[0] = aa.state; pt_io[0] = pt.state;
  if (counters) { counters[0] += cnt.closest; counters[1] += cnt.shadow; counters[2] += cnt.samples; }
  return 0;
}

/* Renderer.__call__(ray) for explicit rays, one after the other on one PCG stream */
int orc_trace_rays(const rt_scene_desc* s, const rt_render_params* p, const double* rays,
                   const int32_t* depth, int n, uint64_t* pt_io, double* out_rgb,
                   uint64_t* counters) {
  pcg_t pt = { pt_io[0], pt_io[1] };
  counters_t cnt = { 0, 0, 0 };
  ctx_t c = { s, p, &pt, &cnt, -1 };
  for (int i = 0; i < n; ++i) {
    const double* q = rays + 8 * (size_t)i;
    ray_t r = { { q[0], q[1], q[2] }, { q[3], q[4], q[5] }, q[6], q[7], depth ? depth[i] : 0 };
    col3 v = renderer_call(&c, &r);
    out_rgb[3 * i + 0] = v.r; out_rgb[3 * i + 1] = v.g; out_rgb[3 * i + 2] = v.b;
  }
  pt_io[0] = pt.state;
  if (counters) { counters[0] += cnt.closest; counters[1] += cnt.shadow; counters[2] += (uint64_t)n; }
  return 0;
}

int orc_intersect(const rt_scene_desc* s, const double* rays, int n, rt_hit* out) {
  for (int i = 0; i < n; ++i) {
    const double* q = rays + 8 * (size_t)i;
    ray_t r = { { q[0], q[1], q[2] }, { q[3], q[4], q[5] }, q[6], q[7], 0 };
    hit_t h;
    memset(&out[i], 0, sizeof(rt_hit));
    if (!world_ray_intersection(s, &r, &h, NULL)) { out[i].shape = -1; out[i].material = -1; continue; }
    out[i].shape = h.shape;
    out[i].material = s->shape_material[h.shape];
    out[i].t = h.t;
    out[i].world_point[0] = h.world_point.x; out[i].world_point[1] = h.world_point.y; out[i].world_point[2] = h.world_point.z;
    out[i].normal[0] = h.normal.x; out[i].normal[1] = h.normal.y; out[i].normal[2] = h.normal.z;
    out[i].uv[0] = h.u; out[i].uv[1] = h.v;
  }
  return 0;
}

int orc_is_point_visible(const rt_scene_desc* s, const double* pairs, int n, uint8_t* out) {
  for (int i = 0; i < n; ++i) {
    const double* q = pairs + 6 * (size_t)i;
    v3 point = { q[0], q[1], q[2] }, obs = { q[3], q[4], q[5] };
    out[i] = (uint8_t)world_is_point_visible(s, point, obs, NULL);
  }
  return 0;
}

int orc_camera_rays(const rt_render_params* p, uint64_t* aa_io, double* out) {
  pcg_t aa = { aa_io[0], aa_io[1] };
  int S = p->samples_per_side;
  size_t k = 0;
  for (int row = 0; row < p->height; ++row)
    for (int col = 0; col < p->width; ++col)
      for (int ir = 0; ir < (S > 0 ? S : 1); ++ir)
        for (int ic = 0; ic < (S > 0 ? S : 1); ++ic) {
          double up = 0.5, vp = 0.5;
          if (S > 0) {
            up = (ic + pcg_random_float(&aa)) / S;
            vp = (ir + pcg_random_float(&aa)) / S;
          }
          ray_t r = tracer_fire_ray(p, col, row, up, vp);
          double* q = out + 8 * k++;
          q[0] = r.o.x; q[1] = r.o.y; q[2] = r.o.z; q[3] = r.d.x; q[4] = r.d.y; q[5] = r.d.z;
          q[6] = r.tmin; q[7] = r.tmax;
        }
  aa_io[0] = aa.state;
  return 0;
}

int orc_pcg_seed(uint64_t init_state, uint64_t init_seq, uint64_t* state_inc) {
  pcg_t p;
  pcg_seed(&p, init_state, init_seq);
  state_inc[0] = p.state; state_inc[1] = p.inc;
  return 0;
}

int orc_pcg_draw(uint64_t* state_inc, int n, uint32_t* out) {
  pcg_t p = { state_inc[0], state_inc[1] };
  for (int i = 0; i < n; ++i) out[i] = pcg_random(&p);
  state_inc[0] = p.state;
  return 0;
}

int orc_pigment_color(const rt_scene_desc* s, int pigment, const double* uv, int n, double* out) {
  for (int i = 0; i < n; ++i) {
    col3 c = pigment_get_color(s, pigment, uv[2 * i], uv[2 * i + 1]);
    out[3 * i] = c.r; out[3 * i + 1] = c.g; out[3 * i + 2] = c.b;
  }
  return 0;
}

int orc_scatter(const rt_scene_desc* s, int material, const double* in, int n,
                uint64_t* state_inc, double* out_rays) {
  pcg_t p = { state_inc[0], state_inc[1] };
  for (int i = 0; i < n; ++i) {
    const double* q = in + 9 * (size_t)i;
    v3 d = { q[0], q[1], q[2] }, pt = { q[3], q[4], q[5] }, nn = { q[6], q[7], q[8] };
    ray_t r = brdf_scatter_ray(&s->materials[material], &p, d, pt, nn, 1);
    double* o = out_rays + 8 * (size_t)i;
    o[0] = r.o.x; o[1] = r.o.y; o[2] = r.o.z; o[3] = r.d.x; o[4] = r.d.y; o[5] = r.d.z;
    o[6] = r.tmin; o[7] = r.tmax;
  }
  state_inc[0] = p.state;
  return 0;
}

int orc_onb(const double* normals, int n, double* out) {
  for (int i = 0; i < n; ++i) {
    v3 nn = { normals[3 * i], normals[3 * i + 1], normals[3 * i + 2] }, e1, e2, e3;
    create_onb_from_z(nn, &e1, &e2, &e3);
    double* o = out + 9 * (size_t)i;
    o[0] = e1.x; o[1] = e1.y; o[2] = e1.z; o[3] = e2.x; o[4] = e2.y; o[5] = e2.z;
    o[6] = e3.x; o[7] = e3.y; o[8] = e3.z;
  }
  return 0;
}
