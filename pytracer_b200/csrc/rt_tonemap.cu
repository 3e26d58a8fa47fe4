// rt_tonemap.cu — the step right after the render: HdrImage.average_luminosity,
// normalize_image, clamp_image and the per-pixel map of write_ldr_image on the device
// (reference: hdrimages.py:120-171; called by main.py:209-215 `demo` / `render` and :226-230 `pfm2png`).
//
// Both kernels stream the fp32 framebuffer [n_pixels][3] once and are HBM-bound:
//   k_lum_sum   12 B read per pixel;  sum of log10(delta + luminosity) in fp64, deterministic
//               (fixed grid, per-block partials summed in a fixed order by the last block to finish)
//   k_tone_map  12 B read + 3 B (LDR) [+ 12 B (normalised HDR)] written per pixel
// Arithmetic is fp64 in the reference's operation order (this file is compiled with --fmad=false): a
// pixel read from a PFM file is an fp32 value the reference promotes to a Python float, exactly what
// (double)rgb[i] is here.
#include "rt_launch.h"

#define RT_TM_THREADS 256
#define RT_TM_MAX_BLOCKS 2048

static __device__ __forceinline__ float fast_rcp_tm(float x) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
  return r;
}
static __device__ __forceinline__ float4 ld_stream(const float4* p) { return __ldcs(p); }

// sum_i log10(v_i) = log10(prod_i m_i) + log10(2) * sum_i e_i with v_i = m_i 2^e_i, m_i in [1, 2): one
// integer add and one fp64 multiply per pixel instead of one fp64 log10 (≈ 60 fp64 instructions, which
// would make this kernel FP64-bound at a third of the HBM rate).  The mantissa product is folded back
// into (mantissa, exponent) every RT_TM_FOLD pixels (it stays below 2^RT_TM_FOLD), and one log10 per
// thread closes the sum.  Mathematically the same quantity; rounding differs from the reference's
// running sum by ~1e-16 per factor, i.e. below the reference's own accumulation error.
#define RT_TM_FOLD 256
struct LogAcc {
  double mant;      // product of mantissas since the last fold, in [1, 2^RT_TM_FOLD)
  long long expo;   // sum of binary exponents
  double logs;      // log10 of everything folded out so far that is not representable as above (zeros, NaNs)
  int pending;
};
static __device__ __forceinline__ void split(double v, double& m, int& e) {
  const long long b = __double_as_longlong(v);
  e = (int)((b >> 52) & 0x7ff) - 1023;
  m = __longlong_as_double((b & 0x800fffffffffffffLL) | 0x3ff0000000000000LL);
}
static __device__ __forceinline__ void fold(LogAcc& a) {
  double m; int e;
  split(a.mant, m, e);
  a.mant = m; a.expo += e; a.pending = 0;
}
static __device__ __forceinline__ void add_pixel(LogAcc& a, double delta, float r, float g, float b) {
  // Color.luminosity, colors.py:59-61 — max and min are exact in fp32, the sum and the halving in fp64
  const double v = delta + ((double)fmaxf(fmaxf(r, g), b) + (double)fminf(fminf(r, g), b)) / 2.0;
  const long long bits = __double_as_longlong(v);
  const int field = (int)((bits >> 52) & 0x7ff);
  if (bits > 0 && field != 0 && field != 0x7ff) {  // positive, normal, finite: every pixel of a real image
    double m; int e;
    split(v, m, e);
    a.mant *= m; a.expo += e;
    if (++a.pending == RT_TM_FOLD) fold(a);
  } else {
    a.logs += log10(v);  // 0 -> -inf, negative -> NaN, subnormal / inf: the libm value (the reference raises on <= 0)
  }
}
static __device__ __forceinline__ double close_acc(LogAcc& a) {
  fold(a);
  return a.logs + log10(a.mant) + (double)a.expo * 0.30102999566398119521;  // log10(2)
}

// Block sum -> partials[block]; the last block to finish adds the partials.  Every step has a fixed
// order (shuffle tree, then warps 0..7, then blocks strided by thread and the same tree again), so the
// result depends on the grid size only, not on scheduling.
static __device__ __forceinline__ double block_sum(double acc, double* red) {
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  double s = 0.0;
  for (int w = 0; w < RT_TM_THREADS / 32; ++w) s += red[w];
  return s;  // same value in every thread
}
static __device__ __forceinline__ void finish_sum(double acc, double* __restrict__ partials, unsigned int* __restrict__ done,
                                                  double* __restrict__ out) {
  __shared__ double red[RT_TM_THREADS / 32];
  __shared__ bool last;
  const double s = block_sum(acc, red);
  if (threadIdx.x == 0) {
    partials[blockIdx.x] = s;
    __threadfence();
    last = atomicAdd(done, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  double t = 0.0;
  for (unsigned b = threadIdx.x; b < gridDim.x; b += RT_TM_THREADS) t += reinterpret_cast<volatile double*>(partials)[b];
  t = block_sum(t, red);
  if (threadIdx.x == 0) { out[0] = t; *done = 0; }
}

// hdrimages.py:120-128: cumsum += log10(delta + pix.luminosity())
__global__ void __launch_bounds__(RT_TM_THREADS)
k_lum_sum(const float* __restrict__ rgb, long long n_pixels, double delta, double* __restrict__ partials,
          unsigned int* __restrict__ done, double* __restrict__ out) {
  LogAcc acc = {1.0, 0, 0.0, 0};
  const long long n_groups = n_pixels / 4;  // 4 pixels = 12 floats = three 16-byte loads
  const float4* v = reinterpret_cast<const float4*>(rgb);
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += (long long)gridDim.x * blockDim.x) {
    // default caching, not evict-first: a frame that fits the 126 MB L2 (up to 3840x2160) is still there
    // when the map kernel reads it next
    const float4 a = __ldg(v + 3 * g), b = __ldg(v + 3 * g + 1), c = __ldg(v + 3 * g + 2);
    add_pixel(acc, delta, a.x, a.y, a.z);
    add_pixel(acc, delta, a.w, b.x, b.y);
    add_pixel(acc, delta, b.z, b.w, c.x);
    add_pixel(acc, delta, c.y, c.z, c.w);
  }
  if (blockIdx.x == 0 && threadIdx.x == 0)  // the 0..3 pixels after the last full group
    for (long long p = n_groups * 4; p < n_pixels; ++p) add_pixel(acc, delta, rgb[3 * p], rgb[3 * p + 1], rgb[3 * p + 2]);
  finish_sum(close_acc(acc), partials, done, out);
}

// unaligned buffers: same sum, scalar loads
__global__ void __launch_bounds__(RT_TM_THREADS)
k_lum_sum_scalar(const float* __restrict__ rgb, long long n_pixels, double delta, double* __restrict__ partials,
                 unsigned int* __restrict__ done, double* __restrict__ out) {
  LogAcc acc = {1.0, 0, 0.0, 0};
  for (long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pixels; p += (long long)gridDim.x * blockDim.x)
    add_pixel(acc, delta, rgb[3 * p], rgb[3 * p + 1], rgb[3 * p + 2]);
  finish_sum(close_acc(acc), partials, done, out);
}

// normalize_image (pixel * (factor / luminosity)), clamp_image (x / (1 + x)) and write_ldr_image's
// int(255 * pow(c, 1 / gamma)) for one channel value (hdrimages.py:130-165)
struct ToneArgs {
  double scale;      // factor / luminosity, formed on the host exactly like hdrimages.py:139
  double inv_gamma;  // 1 / gamma
  int gamma_is_one;  // pow(x, 1.0) == x: skip the call
  int normalize, clamp;  // which of the two HdrImage steps run (RT_TONE_*)
};
static __device__ __forceinline__ double tone_channel(float c, const ToneArgs& t) {
  double x = (double)c;
  if (t.normalize) x = x * t.scale;
  if (t.clamp) x = x / (1.0 + x);
  return x;
}
static __device__ __forceinline__ unsigned int ldr_byte(double y, const ToneArgs& t) {
  const double g = t.gamma_is_one ? y : pow(y, t.inv_gamma);
  const double q = 255.0 * g;
  // int() truncates toward zero; PIL's putpixel then wraps out-of-range ints into a byte, which only
  // happens for negative or non-finite colours — those are clamped here
  if (!(q > 0.0)) return 0u;
  return q >= 255.0 ? 255u : (unsigned int)q;
}

// LDR bytes, LDR-only calls with gamma == 1 (what `render` and pfm2png do by default): the fp64
// division would make the kernel FP64-bound, so a byte is first computed in fp32.  Error of the fp32
// 255 * y against the fp64 value: scale and product roundings 2 x 2^-24, 1 + x 2^-24, MUFU.RCP one ulp
// (2^-23) and one fused multiply-add: below 5.4e-7 relative, i.e. 1.4e-4 absolute at most.  So whenever
// the fp32 value is further than RT_TM_GUARD = 5e-4 from an integer its floor IS the fp64 floor; the
// ~0.1 % of values inside the guard band are flagged and recomputed in fp64 afterwards (out of line,
// rarely entered) — bytes stay bit-exact.
//
// floor without the conversion unit (FRND / F2I run at a quarter of the FADD rate): h = q - 1/2,
// tmp = h + 1.5 * 2^23 rounds h to the nearest integer — floor(q) unless q is within 2^-16 of an integer,
// inside the guard band anyway — and leaves it in the low mantissa bits; d = h - (tmp - 1.5 * 2^23) is
// frac(q) - 1/2, exactly.
#define RT_TM_GUARD 5.0e-4f
#define RT_TM_MAGIC 12582912.0f
static __device__ __noinline__ unsigned int ldr_byte_exact(float c, const ToneArgs t) {
  return ldr_byte(tone_channel(c, t), t);
}
// inputs must be >= 0 (checked per quad by the caller); NaN fails every comparison and is flagged
template <bool NORM, bool CLAMP>
static __device__ __forceinline__ unsigned int ldr_byte_fast(float c, float scale32, float scale255, bool& unsafe) {
  const float x = NORM ? c * scale32 : c;
  const float a = NORM ? c * scale255 : c * 255.0f;
  const float h = CLAMP ? __fmaf_rn(a, fast_rcp_tm(1.0f + x), -0.5f) : a - 0.5f;
  const float tmp = h + RT_TM_MAGIC;
  const float d = h - (tmp - RT_TM_MAGIC);
  // safe: frac in (guard, 1 - guard), or the byte is 0 and frac < 1 - guard (q >= 0: it cannot go below 0)
  unsafe = !(d < 0.5f - RT_TM_GUARD && (d > RT_TM_GUARD - 0.5f || tmp == RT_TM_MAGIC));
  // without clamp_image a channel can exceed 1: the low mantissa byte would wrap modulo 256 where the
  // reference saturates (PIL clips), so anything from 255.5 up goes to the exact path (which returns 255)
  if (!CLAMP) unsafe = unsafe || !(h < 255.0f);
  return (unsigned int)__float_as_int(tmp) & 0xffu;
}

// The map is elementwise over the flat array of 3 n floats, so pixels need not stay together: lane l
// loads the l-th 16-byte quad of a 512-byte run (fully coalesced LDG.128) and stores its four bytes as
// one word (128 contiguous bytes per warp store); RT_TM_UNROLL independent quads per thread keep
// enough loads in flight.
#define RT_TM_UNROLL 4
// Two channel values per instruction: the arithmetic of ldr_byte_fast on Blackwell's packed fp32 pipe
// (mul / add / fma .rn.f32x2 — the same IEEE roundings as the scalar form, half the issue slots; at 14
// scalar instructions per value the kernel kept the issue port 67 % busy and lost to it a fifth of the
// copy bandwidth).  MUFU.RCP and the three comparisons per value stay scalar.
typedef unsigned long long tm_f32x2;
static __device__ __forceinline__ tm_f32x2 tm_pk(float lo, float hi) { tm_f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
static __device__ __forceinline__ void tm_upk(tm_f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
static __device__ __forceinline__ tm_f32x2 tm_mul2(tm_f32x2 a, tm_f32x2 b) { tm_f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
static __device__ __forceinline__ tm_f32x2 tm_add2(tm_f32x2 a, tm_f32x2 b) { tm_f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
static __device__ __forceinline__ tm_f32x2 tm_sub2(tm_f32x2 a, tm_f32x2 b) { tm_f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
static __device__ __forceinline__ tm_f32x2 tm_fma2(tm_f32x2 a, tm_f32x2 b, tm_f32x2 c) { tm_f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <bool NORM, bool CLAMP>
static __device__ __forceinline__ void ldr_pair_fast(float c0, float c1, float scale32, float scale255, unsigned int& b0,
                                                     unsigned int& b1, bool& unsafe0, bool& unsafe1) {
  const tm_f32x2 c = tm_pk(c0, c1);
  const tm_f32x2 x = NORM ? tm_mul2(c, tm_pk(scale32, scale32)) : c;
  const tm_f32x2 a = tm_mul2(c, NORM ? tm_pk(scale255, scale255) : tm_pk(255.0f, 255.0f));
  tm_f32x2 h;
  if (CLAMP) {
    float d0, d1;
    tm_upk(tm_add2(x, tm_pk(1.0f, 1.0f)), d0, d1);
    h = tm_fma2(a, tm_pk(fast_rcp_tm(d0), fast_rcp_tm(d1)), tm_pk(-0.5f, -0.5f));
  } else {
    h = tm_add2(a, tm_pk(-0.5f, -0.5f));
  }
  const tm_f32x2 magic = tm_pk(RT_TM_MAGIC, RT_TM_MAGIC);
  const tm_f32x2 tmp = tm_add2(h, magic);
  const tm_f32x2 d = tm_sub2(h, tm_sub2(tmp, magic));
  float t0, t1, e0, e1;
  tm_upk(tmp, t0, t1);
  tm_upk(d, e0, e1);
  // safe: frac in (guard, 1 - guard), or the byte is 0 and frac < 1 - guard (q >= 0: it cannot go below 0)
  unsafe0 = !(e0 < 0.5f - RT_TM_GUARD && (e0 > RT_TM_GUARD - 0.5f || t0 == RT_TM_MAGIC));
  unsafe1 = !(e1 < 0.5f - RT_TM_GUARD && (e1 > RT_TM_GUARD - 0.5f || t1 == RT_TM_MAGIC));
  if (!CLAMP) {  // values past 255.5 would wrap modulo 256 in the mantissa trick: exact path (saturates like PIL)
    float h0, h1;
    tm_upk(h, h0, h1);
    unsafe0 = unsafe0 || !(h0 < 255.0f);
    unsafe1 = unsafe1 || !(h1 < 255.0f);
  }
  b0 = (unsigned int)__float_as_int(t0) & 0xffu;
  b1 = (unsigned int)__float_as_int(t1) & 0xffu;
}

template <bool NORM, bool CLAMP>
static __device__ __forceinline__ unsigned int ldr_quad(float4 q, const ToneArgs& t, float scale32, float scale255) {
  const float in[4] = {q.x, q.y, q.z, q.w};
  unsigned int b[4];
  bool u[4];
  ldr_pair_fast<NORM, CLAMP>(q.x, q.y, scale32, scale255, b[0], b[1], u[0], u[1]);
  ldr_pair_fast<NORM, CLAMP>(q.z, q.w, scale32, scale255, b[2], b[3], u[2], u[3]);
  unsigned int w = b[0] | (b[1] << 8) | (b[2] << 16) | (b[3] << 24);
  unsigned int redo = (u[0] ? 1u : 0u) | (u[1] ? 2u : 0u) | (u[2] ? 4u : 0u) | (u[3] ? 8u : 0u);
  if (!(fminf(fminf(q.x, q.y), fminf(q.z, q.w)) >= 0.0f)) redo = 0xfu;  // a negative channel: fp64 for the quad
  while (redo) {  // fp64, the reference's own arithmetic, for the flagged channels
    const int k = __ffs(redo) - 1;
    redo &= redo - 1;
    const float cv = k == 0 ? in[0] : (k == 1 ? in[1] : (k == 2 ? in[2] : in[3]));
    w = (w & ~(0xffu << (8 * k))) | (ldr_byte_exact(cv, t) << (8 * k));
  }
  return w;
}
template <bool NORM, bool CLAMP>
__global__ void __launch_bounds__(RT_TM_THREADS)
k_tone_map_ldr(const float* __restrict__ rgb, long long n_pixels, ToneArgs t, unsigned char* __restrict__ out_ldr) {
  const long long n_floats = 3 * n_pixels, n_quads = n_floats / 4;
  const float4* v = reinterpret_cast<const float4*>(rgb);
  unsigned int* o = reinterpret_cast<unsigned int*>(out_ldr);
  const float scale32 = (float)t.scale, scale255 = 255.0f * scale32;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + (RT_TM_UNROLL - 1) * stride < n_quads; i += RT_TM_UNROLL * stride) {
    float4 q[RT_TM_UNROLL];
#pragma unroll
    for (int u = 0; u < RT_TM_UNROLL; ++u) q[u] = ld_stream(v + i + u * stride);
#pragma unroll
    for (int u = 0; u < RT_TM_UNROLL; ++u) __stcs(o + i + u * stride, ldr_quad<NORM, CLAMP>(q[u], t, scale32, scale255));
  }
  for (; i < n_quads; i += stride) __stcs(o + i, ldr_quad<NORM, CLAMP>(ld_stream(v + i), t, scale32, scale255));
  if (blockIdx.x == 0 && threadIdx.x < (int)(n_floats - 4 * n_quads)) {  // 0..3 floats after the last quad
    const long long j = 4 * n_quads + threadIdx.x;
    out_ldr[j] = (unsigned char)ldr_byte_exact(rgb[j], t);
  }
}

__global__ void __launch_bounds__(RT_TM_THREADS)
k_tone_map(const float* __restrict__ rgb, long long n_pixels, ToneArgs t, float* __restrict__ out_hdr,
           unsigned char* __restrict__ out_ldr, int vec_ok) {
  const long long n_groups = vec_ok ? n_pixels / 4 : 0;
  const float4* v = reinterpret_cast<const float4*>(rgb);
  for (long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x; g < n_groups; g += (long long)gridDim.x * blockDim.x) {
    const float4 a = ld_stream(v + 3 * g), b = ld_stream(v + 3 * g + 1), c = ld_stream(v + 3 * g + 2);
    const float in[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};
    double y[12];
#pragma unroll
    for (int k = 0; k < 12; ++k) y[k] = tone_channel(in[k], t);
    if (out_hdr) {
      float4* o = reinterpret_cast<float4*>(out_hdr) + 3 * g;
      __stcs(o, make_float4((float)y[0], (float)y[1], (float)y[2], (float)y[3]));
      __stcs(o + 1, make_float4((float)y[4], (float)y[5], (float)y[6], (float)y[7]));
      __stcs(o + 2, make_float4((float)y[8], (float)y[9], (float)y[10], (float)y[11]));
    }
    if (out_ldr) {
      unsigned int w[3];
#pragma unroll
      for (int k = 0; k < 3; ++k)
        w[k] = ldr_byte(y[4 * k], t) | (ldr_byte(y[4 * k + 1], t) << 8) | (ldr_byte(y[4 * k + 2], t) << 16) |
               (ldr_byte(y[4 * k + 3], t) << 24);
      unsigned int* o = reinterpret_cast<unsigned int*>(out_ldr) + 3 * g;
      __stcs(o, w[0]); __stcs(o + 1, w[1]); __stcs(o + 2, w[2]);
    }
  }
  // pixels not covered by full aligned groups
  const long long first = n_groups * 4;
  for (long long p = first + (long long)blockIdx.x * blockDim.x + threadIdx.x; p < n_pixels; p += (long long)gridDim.x * blockDim.x) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      const double y = tone_channel(rgb[3 * p + k], t);
      if (out_hdr) out_hdr[3 * p + k] = (float)y;
      if (out_ldr) out_ldr[3 * p + k] = (unsigned char)ldr_byte(y, t);
    }
  }
}

// persistent grid: every resident CTA slot of the device, grid-stride beyond
template <typename K> static int tm_blocks(K kernel, long long n_items, int sm_count) {
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, RT_TM_THREADS, 0) != cudaSuccess || per_sm < 1) per_sm = 1;
  long long want = (n_items + RT_TM_THREADS - 1) / RT_TM_THREADS;
  long long cap = (long long)sm_count * per_sm;
  if (cap > RT_TM_MAX_BLOCKS) cap = RT_TM_MAX_BLOCKS;
  if (want > cap) want = cap;
  return (int)(want < 1 ? 1 : want);
}

int tonemap_max_blocks() { return RT_TM_MAX_BLOCKS; }

cudaError_t launch_lum_sum(const float* d_rgb, long long n_pixels, double delta, double* d_partials,
                           unsigned int* d_done, double* d_out, int sm_count, cudaStream_t st) {
  const bool vec_ok = (reinterpret_cast<uintptr_t>(d_rgb) & 15) == 0;
  if (vec_ok) {
    const int blocks = tm_blocks(k_lum_sum, n_pixels / 4 + 1, sm_count);
    k_lum_sum<<<blocks, RT_TM_THREADS, 0, st>>>(d_rgb, n_pixels, delta, d_partials, d_done, d_out);
  } else {
    const int blocks = tm_blocks(k_lum_sum_scalar, n_pixels, sm_count);
    k_lum_sum_scalar<<<blocks, RT_TM_THREADS, 0, st>>>(d_rgb, n_pixels, delta, d_partials, d_done, d_out);
  }
  return cudaGetLastError();
}

cudaError_t launch_tone_map(const float* d_rgb, long long n_pixels, int flags, double scale, double gamma, float* d_out_hdr,
                            unsigned char* d_out_ldr, int sm_count, cudaStream_t st) {
  ToneArgs t;
  t.normalize = (flags & RT_TONE_NORMALIZE) != 0;
  t.clamp = (flags & RT_TONE_CLAMP) != 0;
  t.scale = scale;
  t.inv_gamma = 1.0 / gamma;
  t.gamma_is_one = gamma == 1.0;
  const int vec_ok = (reinterpret_cast<uintptr_t>(d_rgb) & 15) == 0 && (!d_out_hdr || (reinterpret_cast<uintptr_t>(d_out_hdr) & 15) == 0) &&
                     (!d_out_ldr || (reinterpret_cast<uintptr_t>(d_out_ldr) & 3) == 0);
  // fp32 fast path: LDR only, gamma 1, a positive finite scale whose fp32 products cannot overflow
  const bool scale_ok = !t.normalize || (scale > 1e-30 && scale < 1e30);
  if (vec_ok && !d_out_hdr && d_out_ldr && t.gamma_is_one && scale_ok) {
    void (*kern)(const float*, long long, ToneArgs, unsigned char*) =
        t.normalize ? (t.clamp ? k_tone_map_ldr<true, true> : k_tone_map_ldr<true, false>)
                    : (t.clamp ? k_tone_map_ldr<false, true> : k_tone_map_ldr<false, false>);
    const int blocks = tm_blocks(kern, (3 * n_pixels / 4 + RT_TM_UNROLL - 1) / RT_TM_UNROLL + 1, sm_count);
    kern<<<blocks, RT_TM_THREADS, 0, st>>>(d_rgb, n_pixels, t, d_out_ldr);
    return cudaGetLastError();
  }
  const int blocks = tm_blocks(k_tone_map, vec_ok ? n_pixels / 4 + 1 : n_pixels, sm_count);
  k_tone_map<<<blocks, RT_TM_THREADS, 0, st>>>(d_rgb, n_pixels, t, d_out_hdr, d_out_ldr, vec_ok);
  return cudaGetLastError();
}
