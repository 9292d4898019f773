// Host build of the scalar building blocks in ns_math.cuh, for CPU unit tests
// (tests/test_native_math.py).  Not part of the product library.
#include "ns_math.cuh"

static const double kTab[NS_EXP_N] = {NS_EXP_TAB_VALUES};

extern "C" {
void nsh_exp64(const double* a, double* out, int64_t n) {
  for (int64_t i = 0; i < n; ++i) out[i] = ns_exp64_neg(a[i], kTab);
}
int nsh_interval_update(uint64_t nb, uint64_t nt, int precision, uint64_t* lo, uint64_t* hi) {
  return ns_interval_update(nb, nt, precision, lo, hi);
}
uint64_t nsh_read_bits(const uint32_t* words, int32_t pos, int32_t len, int count) {
  return ns_read_bits(words, pos, len, count);
}
void nsh_write_bits(uint32_t* words, int32_t pos, uint64_t value, int count) {
  ns_write_bits(words, pos, value, count);
}
uint32_t nsh_orderable(float f) { return ns_f32_orderable(f); }
}
