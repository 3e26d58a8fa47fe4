// rt_pcg.cuh — PCG32 XSH-RR on the device (reference: pcg.py:22-62), plus the two things a
// parallel renderer needs that the sequential reference does not: O(popcount) jump-ahead on a
// stream (LCG composition table built on the host) and hashed sub-streams for scatter records.
#pragma once
#include "rt_math.cuh"

#define RT_PCG_MULT 6364136223846793005ULL
#define RT_JUMP_BITS 44

struct Pcg {
  uint64_t state, inc;
};

RT_DEV uint32_t pcg_random(Pcg& p) {  // pcg.py:43-58
  uint64_t old = p.state;
  p.state = old * RT_PCG_MULT + p.inc;
  uint32_t xorshifted = (uint32_t)(((old >> 18) ^ old) >> 27);
  uint32_t rot = (uint32_t)(old >> 59);
  return __funnelshift_r(xorshifted, xorshifted, rot);  // 32-bit rotate right
}

// pcg.py:60-62: random() / 0xFFFFFFFF, in [0, 1] inclusive.  The fp64 form is the reference's
// correctly rounded quotient; the fp32 form rounds that quotient to float.
template <typename T> RT_DEV T pcg_random_float(Pcg& p);
template <> RT_DEV double pcg_random_float<double>(Pcg& p) { return (double)pcg_random(p) / 4294967295.0; }
template <> RT_DEV float pcg_random_float<float>(Pcg& p) {
  return __uint2float_rn(pcg_random(p)) * 2.3283064370807974e-10f;  // 1/0xFFFFFFFF
}

RT_DEV Pcg pcg_seed(uint64_t init_state, uint64_t init_seq) {  // pcg.py:29-41
  Pcg p;
  p.state = 0;
  p.inc = (init_seq << 1) | 1;
  pcg_random(p);
  p.state += init_state;
  pcg_random(p);
  return p;
}

// state -> state after `delta` draws, given the table {A_b, C_b} of x -> A_b x + C_b = LCG^(2^b)
// for this stream's increment (host: rt_api.cu build_jump_table).
struct JumpTable {
  uint64_t mult[RT_JUMP_BITS];
  uint64_t plus[RT_JUMP_BITS];
};

RT_DEV uint64_t pcg_jump(uint64_t state, uint64_t delta, const JumpTable& t) {
  delta &= (1ull << RT_JUMP_BITS) - 1ull;
#pragma unroll 1
  while (delta != 0) {  // one multiply-add per SET bit
    const int b = __ffsll((long long)delta) - 1;
    state = state * t.mult[b] + t.plus[b];
    delta &= delta - 1;
  }
  return state;
}

// Advance by a small number of draws without a table (used to skip the draws the reference
// spends on scatter rays it then cuts at depth > max_depth).
RT_DEV void pcg_skip(Pcg& p, uint32_t n) {
  uint64_t cur_mult = RT_PCG_MULT, cur_plus = p.inc, acc_mult = 1, acc_plus = 0;
  while (n) {
    if (n & 1) { acc_mult *= cur_mult; acc_plus = acc_plus * cur_mult + cur_plus; }
    cur_plus = (cur_mult + 1) * cur_plus;
    cur_mult *= cur_mult;
    n >>= 1;
  }
  p.state = acc_mult * p.state + acc_plus;
}

// splitmix64 finaliser: decorrelates the sub-stream of child `i` of a scatter record
RT_DEV uint64_t mix64(uint64_t z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

// Sub-stream of child `i` of a scatter record (the wavefront kernels): z = mix64(record state + (i+1) phi64).
// The two halves of z are the child's two scatter uniforms (materials.py:136-137 draws two numbers per
// scattered ray) and z itself is the state the child's own record hands on (a roulette draw, the only other
// number a ray may need, is a PCG step from z).  splitmix64's finaliser is a bijection with full avalanche, so
// the 2 x 32 bits are as good as two generator outputs — and cost no LCG steps (two 64-bit multiply-adds and
// two output permutations per ray, 5 % of the instruction stream of demo.txt).
RT_DEV uint64_t child_stream(uint64_t record_state, int child) {
  return mix64(record_state + (uint64_t)(child + 1) * 0x9E3779B97F4A7C15ULL);
}
RT_DEV float unit_from_u32(uint32_t x) { return __uint2float_rn(x) * 2.3283064370807974e-10f; }  // pcg.py:60-62: x / 0xFFFFFFFF in [0, 1]
