/* Plain-C client of the C ABI (include/mapanything_b200.h): compiled with gcc by tests/test_cabi_cpu.py and linked against
 * libmapanything_b200.so.  Only HOST entry points are exercised (no GPU): version, error reporting, and the resampling
 * tables, whose window weights must sum to exactly 2^22 within rounding. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "mapanything_b200.h"

int main(void) {
  int ksize = 0, i, t;
  const int in_size = 1920, out_size = 522;
  int32_t *bounds, *coeffs;
  uint32_t* packed;
  if (ma_abi_version() != MA_ABI_VERSION) return 1;
  if (ma_resample_coeffs(0, 10, MA_FILTER_LANCZOS, &ksize, NULL, NULL) == MA_OK) return 2;
  if (strstr(ma_last_error(), "bad sizes") == NULL) return 3;
  if (ma_resample_coeffs(in_size, out_size, MA_FILTER_LANCZOS, &ksize, NULL, NULL) != MA_OK || ksize != 25) return 4;
  bounds = (int32_t*)malloc(sizeof(int32_t) * 2 * out_size);
  coeffs = (int32_t*)malloc(sizeof(int32_t) * ksize * out_size);
  packed = (uint32_t*)malloc(sizeof(uint32_t) * 3 * ((ksize + 3) / 4) * out_size);
  if (ma_resample_coeffs(in_size, out_size, MA_FILTER_LANCZOS, &ksize, bounds, coeffs) != MA_OK) return 5;
  for (i = 0; i < out_size; ++i) {
    long sum = 0;
    if (bounds[2 * i] < 0 || bounds[2 * i + 1] < 1 || bounds[2 * i] + bounds[2 * i + 1] > in_size) return 6;
    for (t = 0; t < ksize; ++t) sum += coeffs[t * out_size + i];
    if (labs(sum - (1L << 22)) > ksize) return 7; /* each tap rounds by at most 1/2 */
  }
  if (ma_resample_pack_coeffs(coeffs, ksize, out_size, packed) != MA_OK) return 8;
  printf("cabi host check ok: ksize %d, first window [%d, %d)\n", ksize, bounds[0], bounds[0] + bounds[1]);
  free(bounds);
  free(coeffs);
  free(packed);
  return 0;
}
