// TEST INFRASTRUCTURE — CPU restatement of the *project's own* random-stream specification.
//
// The reference draws its randomness from third-party crates that are absent from /root/reference
// (rand 0.10 ChaCha8Rng + rand_distr 0.6 ziggurat StandardNormal; no Cargo.lock, so not even pinned) and
// no reference test pins a value produced by them, so bit-parity with the Rust streams is "parity unpinned".
// What IS kept from the reference is the per-chain stream separation and the CONSUMPTION ORDER
// (SURVEY.md §8 a20): d normals at trajectory start (src/dynamics/transformed_hamiltonian.rs:698), one bool per
// doubling (src/nuts.rs:334), one random_bool per merge when other.log_size < self_log_size (src/nuts.rs:199-203),
// one Uniform(1-j,1+j) per step-size update (src/stepsize/adapt.rs:259-263).
//
// Stream definition (shared by oracle and GPU, stated in DESIGN.md §RNG):
//   Philox4x32-10; key = (seed lo32, seed hi32); counter = (c lo32, c hi32, stream lo32, stream hi32);
//   stream = global chain id + 1 (reference src/sampler.rs:1105-1106: set_stream(chain_id + 1)); c = per-chain event counter.
//   block(c) -> r0..r3;  a = r0 | r1<<32;  b = r2 | r3<<32
//   next_bool     : block(c++), r0 & 1
//   next_f64      : block(c++), (a>>11) * 2^-53                       in [0,1)
//   random_bool(p): next_f64() < p
//   uniform(lo,hi): fma(hi-lo, next_f64(), lo)
//   fill_normal(d): for pair p < ceil(d/2): block(c+p): u1 = ((a>>11)+1)*2^-53 in (0,1], u2 = (b>>11)*2^-53 in [0,1)
//                   r = sqrt(-2*det_log(u1)); (s,co) = det_sincos2pi(u2); out[2p] = r*co; out[2p+1] = r*s;  c += ceil(d/2)
//   det_log / det_sincos2pi use only IEEE +,-,*,/,fma,sqrt in a fixed order => bit-identical on x86 and sm_100.
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may use anything under oracle/.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>

namespace oracle {

struct PhiloxBlock {
  uint32_t r[4];
};

inline PhiloxBlock philox4x32_10(uint64_t seed, uint64_t stream, uint64_t counter) {
  uint32_t c0 = (uint32_t)counter, c1 = (uint32_t)(counter >> 32);
  uint32_t c2 = (uint32_t)stream, c3 = (uint32_t)(stream >> 32);
  uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
  for (int round = 0; round < 10; ++round) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c0;
    uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
    uint32_t n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
    uint32_t n3 = (uint32_t)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  return PhiloxBlock{{c0, c1, c2, c3}};
}

// log(u) for u in (0,1]: u = m*2^e, m in [sqrt(1/2), sqrt(2)); s=(m-1)/(m+1); log m = 2s*(1+z/3+z^2/5+...), z=s^2.
inline double det_log(double u) {
  uint64_t bits;
  std::memcpy(&bits, &u, 8);
  int e = (int)((bits >> 52) & 0x7ff) - 1023;
  bits = (bits & 0x000fffffffffffffull) | 0x3ff0000000000000ull;
  double m;
  std::memcpy(&m, &bits, 8);  // [1,2)
  if (m > 1.4142135623730951) {
    m = m * 0.5;  // exact
    e += 1;
  }
  double s = (m - 1.0) / (m + 1.0);
  double z = s * s;
  double p = 1.0 / 23.0;
  p = std::fma(p, z, 1.0 / 21.0);
  p = std::fma(p, z, 1.0 / 19.0);
  p = std::fma(p, z, 1.0 / 17.0);
  p = std::fma(p, z, 1.0 / 15.0);
  p = std::fma(p, z, 1.0 / 13.0);
  p = std::fma(p, z, 1.0 / 11.0);
  p = std::fma(p, z, 1.0 / 9.0);
  p = std::fma(p, z, 1.0 / 7.0);
  p = std::fma(p, z, 1.0 / 5.0);
  p = std::fma(p, z, 1.0 / 3.0);
  p = std::fma(p, z, 1.0);
  double logm = (2.0 * s) * p;
  return std::fma((double)e, 0.6931471805599453, logm);
}

// (sin, cos)(2*pi*u) for u in [0,1): quadrant q = floor(4u+0.5), r = u - q/4 in [-1/8,1/8], theta = 2*pi*r.
inline void det_sincos2pi(double u, double* s_out, double* c_out) {
  int q = (int)std::floor(u * 4.0 + 0.5);
  double r = u - 0.25 * (double)q;
  double t = r * 6.283185307179586;
  double t2 = t * t;
  // sin t = t * (1 - t2/3! + t2^2/5! - ... ) up to t^17
  double ps = 1.0 / 355687428096000.0;       // 1/17!
  ps = std::fma(ps, t2, -1.0 / 1307674368000.0);  // -1/15!
  ps = std::fma(ps, t2, 1.0 / 6227020800.0);      // 1/13!
  ps = std::fma(ps, t2, -1.0 / 39916800.0);       // -1/11!
  ps = std::fma(ps, t2, 1.0 / 362880.0);          // 1/9!
  ps = std::fma(ps, t2, -1.0 / 5040.0);           // -1/7!
  ps = std::fma(ps, t2, 1.0 / 120.0);             // 1/5!
  ps = std::fma(ps, t2, -1.0 / 6.0);              // -1/3!
  ps = std::fma(ps, t2, 1.0);
  double sn = t * ps;
  // cos t = 1 - t2/2! + ... up to t^18
  double pc = -1.0 / 6402373705728000.0;          // -1/18!
  pc = std::fma(pc, t2, 1.0 / 20922789888000.0);  // 1/16!
  pc = std::fma(pc, t2, -1.0 / 87178291200.0);    // -1/14!
  pc = std::fma(pc, t2, 1.0 / 479001600.0);       // 1/12!
  pc = std::fma(pc, t2, -1.0 / 3628800.0);        // -1/10!
  pc = std::fma(pc, t2, 1.0 / 40320.0);           // 1/8!
  pc = std::fma(pc, t2, -1.0 / 720.0);            // -1/6!
  pc = std::fma(pc, t2, 1.0 / 24.0);              // 1/4!
  pc = std::fma(pc, t2, -0.5);                    // -1/2!
  pc = std::fma(pc, t2, 1.0);
  double cs = pc;
  switch (q & 3) {
    case 0: *s_out = sn; *c_out = cs; break;
    case 1: *s_out = cs; *c_out = -sn; break;
    case 2: *s_out = -sn; *c_out = -cs; break;
    default: *s_out = -cs; *c_out = sn; break;
  }
}

struct Rng {
  uint64_t seed = 0, stream = 1, counter = 0;
  Rng() {}
  Rng(uint64_t seed_, uint64_t stream_, uint64_t counter_ = 0) : seed(seed_), stream(stream_), counter(counter_) {}

  static double u53(uint64_t a) { return (double)(a >> 11) * 0x1.0p-53; }
  bool next_bool() {
    PhiloxBlock b = philox4x32_10(seed, stream, counter++);
    return (b.r[0] & 1u) != 0;
  }
  double next_f64() {
    PhiloxBlock b = philox4x32_10(seed, stream, counter++);
    uint64_t a = (uint64_t)b.r[0] | ((uint64_t)b.r[1] << 32);
    return u53(a);
  }
  bool random_bool(double p) { return next_f64() < p; }
  double uniform(double lo, double hi) { return std::fma(hi - lo, next_f64(), lo); }
  static void normal_pair(uint64_t seed, uint64_t stream, uint64_t ctr, double* n0, double* n1) {
    PhiloxBlock b = philox4x32_10(seed, stream, ctr);
    uint64_t a = (uint64_t)b.r[0] | ((uint64_t)b.r[1] << 32);
    uint64_t bb = (uint64_t)b.r[2] | ((uint64_t)b.r[3] << 32);
    double u1 = (double)((a >> 11) + 1) * 0x1.0p-53;
    double u2 = (double)(bb >> 11) * 0x1.0p-53;
    double r = std::sqrt(-2.0 * det_log(u1));
    double s, c;
    det_sincos2pi(u2, &s, &c);
    *n0 = r * c;
    *n1 = r * s;
  }
  void fill_normal(double* out, size_t d) {
    size_t pairs = (d + 1) / 2;
    for (size_t p = 0; p < pairs; ++p) {
      double n0, n1;
      normal_pair(seed, stream, counter + p, &n0, &n1);
      out[2 * p] = n0;
      if (2 * p + 1 < d) out[2 * p + 1] = n1;
    }
    counter += pairs;
  }
};

}  // namespace oracle
