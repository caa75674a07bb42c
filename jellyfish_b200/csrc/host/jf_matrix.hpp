// jf_matrix.hpp -- GF(2) rectangular matrix used as the k-mer hash.
//
// Behavioural contract (reference include/jellyfish/rectangular_binary_matrix.hpp):
//  * r x c bits, r <= 64, stored one uint64 per column                      (:30-38)
//  * times(v) = XOR of columns[c-1-i] over the set bits i of the 2k-bit key (:223-261):
//    the column index is REVERSED with respect to the bit index.
//  * identity (no columns) when the table is as large as the key space: v & mask (:225)
//  * the hash matrix of a table is the pseudo-inverse of a freshly drawn random
//    matrix (lib/rectangular_binary_matrix.cc:160-210,240-247; large_hash_array.hpp:992-1002)
// This is an independent implementation (row-reduction written against that contract).
#ifndef JFB_MATRIX_HPP
#define JFB_MATRIX_HPP
#include <stdint.h>
#include <vector>
#include <stdexcept>
#include <algorithm>
#include "jf_random.hpp"

namespace jfb {

class gf2_matrix {
  unsigned r_, c_;
  bool     identity_;            // "no columns" identity of the reference (r == c)
  std::vector<uint64_t> col_;    // c_ columns, bit j of col_[i] = row j (row 0 = LSB of the result)

  uint64_t cmask() const { return r_ >= 64 ? ~(uint64_t)0 : (((uint64_t)1 << r_) - 1); }
public:
  gf2_matrix() : r_(0), c_(0), identity_(true) { }
  gf2_matrix(unsigned r, unsigned c) : r_(r), c_(c), identity_(false), col_(c, 0) {
    if(r == 0 || r > 64 || c == 0) throw std::out_of_range("Invalid matrix size");
  }
  template<typename It>
  gf2_matrix(unsigned r, unsigned c, It raw) : r_(r), c_(c), identity_(false), col_(c, 0) {
    for(unsigned i = 0; i < c; ++i, ++raw) col_[i] = (uint64_t)*raw & cmask();
  }
  static gf2_matrix identity(unsigned c) { gf2_matrix m; m.r_ = m.c_ = c; m.identity_ = true; return m; }
  // the "low identity" r x c matrix: picks the low r bits of the vector
  static gf2_matrix low_identity(unsigned r, unsigned c) {
    gf2_matrix m(r, c);
    unsigned row = std::min(r, c), col = c - row;
    for(unsigned i = col; i < c; ++i) m.col_[i] = (uint64_t)1 << (row - 1 - (i - col));
    return m;
  }

  unsigned r() const { return r_; }
  unsigned c() const { return c_; }
  bool is_identity() const { return identity_; }
  uint64_t operator[](unsigned i) const { return identity_ ? ((uint64_t)1 << i) : col_[i]; }
  const std::vector<uint64_t>& columns() const { return col_; }
  unsigned nb_words() const { return (c_ + 63) / 64; }

  bool is_low_identity() const {
    if(identity_) return true;
    unsigned row = std::min(r_, c_), col = c_ - row;
    for(unsigned i = 0; i < col; ++i) if(col_[i]) return false;
    for(unsigned i = col; i < c_; ++i) if(col_[i] != ((uint64_t)1 << (row - 1 - (i - col)))) return false;
    return true;
  }

  bool operator==(const gf2_matrix& o) const {
    if(r_ != o.r_ || c_ != o.c_ || identity_ != o.identity_) return false;
    return identity_ || col_ == o.col_;
  }
  bool operator!=(const gf2_matrix& o) const { return !(*this == o); }

  // v: little-endian array of nb_words() 64-bit words holding the c-bit vector.
  uint64_t times(const uint64_t* v) const {
    if(identity_) return v[0] & cmask();
    uint64_t res = 0;
    for(unsigned i = 0; i < c_; ++i)
      if((v[i >> 6] >> (i & 63)) & 1) res ^= col_[c_ - 1 - i];
    return res;
  }

  void randomize(glibc_random& rng) {
    for(unsigned i = 0; i < c_; ++i) col_[i] = rng.bits(64) & cmask();
  }

  // The matrix is viewed as square (c x c) by stacking [I 0] on top of it: the
  // top c-r rows copy the high c-r bits of the vector, the bottom r rows are this
  // matrix.  Returns the bottom r rows of the inverse of that square matrix, i.e.
  // N such that N * [high bits of v : this*v] = low r bits of v.
  // Throws std::domain_error when singular.
  gf2_matrix pseudo_inverse() const {
    if(identity_) return *this;
    std::vector<uint64_t> piv(col_);
    gf2_matrix res = low_identity(r_, c_);
    const unsigned srow = std::min(r_, c_), scol = c_ - srow;
    // forward elimination on columns scol..c-1, pivot rows from the top (bit srow-1) down
    uint64_t mask = (uint64_t)1 << (srow - 1);
    for(unsigned i = scol; i < c_; ++i, mask >>= 1) {
      if(!(piv[i] & mask)) {
        unsigned j = i + 1;
        while(j < c_ && !(piv[j] & mask)) ++j;
        if(j == c_) throw std::domain_error("Matrix is singular");
        piv[i] ^= piv[j]; res.col_[i] ^= res.col_[j];
      }
      for(unsigned j = i + 1; j < c_; ++j)
        if(piv[j] & mask) { piv[j] ^= piv[i]; res.col_[j] ^= res.col_[i]; }
    }
    // back substitution: clear the pivot rows in every column to the left
    mask = (uint64_t)1 << (srow - 1);
    for(unsigned i = scol; i < c_; ++i, mask >>= 1)
      for(unsigned j = 0; j < i; ++j)
        if(piv[j] & mask) { piv[j] ^= piv[i]; res.col_[j] ^= res.col_[i]; }
    return res;
  }

  // Redraw until pseudo-invertible, return the pseudo-inverse (the table's hash matrix).
  gf2_matrix randomize_pseudo_inverse(glibc_random& rng) {
    for(;;) {
      randomize(rng);
      try { return pseudo_inverse(); } catch(std::domain_error&) { }
    }
  }
};

} // namespace jfb
#endif
