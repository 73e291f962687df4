// Host build of the sequential bin-stage logic (kalloc.h) so tests can exercise the exact code the
// CUDA sampler runs on a box without a GPU.  Test infrastructure; not linked into libsamble_b200.so.
#include "kalloc.h"

extern "C" void samble_host_num_points_to_choose(const float* bin_prob, const long long* cnt, int B, int nb,
                                                 int total, int* k_out) {
  for (int b = 0; b < B; ++b) samble::num_points_to_choose(bin_prob + b * nb, cnt + b * nb, nb, total, k_out + b * nb);
}

extern "C" float samble_host_aten_row_sum(const float* x, int n) { return samble::aten_row_sum(x, n); }
