// ref_matrix -- prints hash matrices drawn by the UNMODIFIED reference library.  TEST INFRASTRUCTURE
// ONLY (built into oracle/_ref/, linked against the reference's own objects; used by
// scripts/make_matrix_golden.py to write tests/golden/matrix_golden.json and by tests/test_oracle.py).
//   ref_matrix R C [SKIP]   -> columns of the (SKIP+1)-th matrix that
//   RectangularBinaryMatrix(R, C).randomize_pseudo_inverse() returns (what large_hash::array keeps as
//   its hash matrix and file_header records: large_hash_array.hpp:992-1002), one per line.
#include <jellyfish/rectangular_binary_matrix.hpp>
#include <jellyfish/misc.hpp>
#include <cstdio>
#include <cstdlib>
int main(int argc, char** argv) {
  if(argc < 3) { fprintf(stderr, "usage: ref_matrix R C [SKIP]\n"); return 1; }
  unsigned r = atoi(argv[1]), c = atoi(argv[2]), skip = argc > 3 ? atoi(argv[3]) : 0;
  for(unsigned s = 0; ; ++s) {
    jellyfish::RectangularBinaryMatrix m(r, c);
    jellyfish::RectangularBinaryMatrix inv = m.randomize_pseudo_inverse();
    if(s < skip) continue;
    for(unsigned i = 0; i < c; ++i) printf("%llu\n", (unsigned long long)inv[i]);
    return 0;
  }
}
