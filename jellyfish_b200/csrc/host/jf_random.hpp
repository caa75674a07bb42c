// jf_random.hpp -- the exact bit stream the reference hashes are seeded from.
//
// The reference draws its hash matrices from *unseeded* glibc random()
// (reference lib/misc.cc:66-72 `random_bits`, never preceded by srandom() in lib/ or
// sub_commands/).  To be independent of the host libc we restate glibc's default
// generator here: TYPE_3 additive feedback, degree 31, separation 3, seeded with 1
// through the 16807 Lehmer LCG, first 310 outputs discarded, each output r >> 1.
#ifndef JFB_RANDOM_HPP
#define JFB_RANDOM_HPP
#include <stdint.h>

namespace jfb {

class glibc_random {
  int32_t r_[34];
  int     f_, b_;   // front / rear indices into r_[0..30]
  uint32_t step() {
    uint32_t v = (uint32_t)r_[f_] + (uint32_t)r_[b_];
    r_[f_] = (int32_t)v;
    if(++f_ >= 31) f_ = 0;
    if(++b_ >= 31) b_ = 0;
    return v >> 1;
  }
public:
  explicit glibc_random(uint32_t seed = 1) { reseed(seed); }
  void reseed(uint32_t seed) {
    if(seed == 0) seed = 1;
    r_[0] = (int32_t)seed;
    for(int i = 1; i < 31; ++i) {
      // 16807 * r[i-1] % 2147483647 without overflow (Schrage)
      int64_t hi = r_[i - 1] / 127773, lo = r_[i - 1] % 127773;
      int64_t w  = 16807 * lo - 2836 * hi;
      if(w < 0) w += 2147483647;
      r_[i] = (int32_t)w;
    }
    f_ = 3; b_ = 0;
    for(int i = 0; i < 310; ++i) step();
  }
  // == random()
  long next() { return (long)step(); }
  // == random_bits(length), reference lib/misc.cc:66-72 (step ConstFloorLog2<RAND_MAX> = 30 bits: the 31-bit draws overlap by one bit)
  uint64_t bits(int length) {
    uint64_t res = 0;
    for(int i = 0; i < length; i += 30)
      res ^= (uint64_t)next() << i;
    return res & ((uint64_t)-1 >> (64 - length));
  }
};

} // namespace jfb
#endif
