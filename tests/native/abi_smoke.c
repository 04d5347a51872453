/* Plain-C consumer of include/plonky_b200.h (compiled with gcc -std=c11 by tests/test_c_abi.py): proves the header is
 * valid C for an FFI binding and that a non-Python harness gets the reference's answers through the ABI.
 *   case 1  test_msm       (src/curve/curve_msm.rs:218-241): generators G, 2G, 3G of BLS12-377, w = 5
 *   case 2  fft_and_ifft   (src/fft.rs:164-185): degree 200, c_i = i * 1337 % 100 over Bls12377Scalar
 * Inputs and expected outputs come from abi_golden.h, generated from the big-integer oracle by the test. */
#include <stdio.h>
#include <string.h>
#include "plonky_b200.h"
#include "abi_golden.h"

static int check(int status, const char* what) {
  if (status != PLK_OK) {
    fprintf(stderr, "%s failed: %s (%s)\n", what, plk_status_string(status), plk_last_error_message());
    return 1;
  }
  return 0;
}

int main(void) {
  if (plk_abi_version() != PLK_ABI_VERSION) { fprintf(stderr, "ABI version mismatch\n"); return 2; }
  int devices = 0;
  if (check(plk_device_count(&devices), "plk_device_count") || devices < 1) { fprintf(stderr, "no CUDA device\n"); return 3; }
  if (check(plk_set_device(0), "plk_set_device")) return 3;

  /* ---- case 1: msm_precompute + msm_execute, and msm_parallel, against the oracle's point ---- */
  plk_msm_table* table = NULL;
  if (check(plk_msm_precompute(PLK_CURVE_BLS12_377, kMsmGeneratorsXyz, NULL, 3, 5, &table), "plk_msm_precompute")) return 4;
  if (plk_msm_table_len(table) != 3 || plk_msm_table_window(table) != 5) { fprintf(stderr, "table metadata\n"); return 4; }
  uint64_t out[18];
  uint8_t zero = 9;
  if (check(plk_msm_execute(table, kMsmScalars, 3, out, &zero), "plk_msm_execute")) return 4;
  if (zero != 0 || memcmp(out, kMsmExpectedXy, sizeof(kMsmExpectedXy)) != 0) { fprintf(stderr, "test_msm: wrong point\n"); return 5; }
  uint64_t out2[18];
  if (check(plk_msm_parallel(PLK_CURVE_BLS12_377, kMsmScalars, kMsmGeneratorsXyz, NULL, 3, 5, out2, &zero), "plk_msm_parallel")) return 4;
  if (zero != 0 || memcmp(out2, out, sizeof(out)) != 0) { fprintf(stderr, "msm_parallel != msm_execute\n"); return 5; }
  /* length mismatch is the reference's assert_eq! (curve_msm.rs:67): a status, never an abort */
  if (plk_msm_execute(table, kMsmScalars, 2, out2, &zero) != PLK_ELENGTH) { fprintf(stderr, "expected PLK_ELENGTH\n"); return 5; }
  plk_msm_free(table);

  /* ---- case 2: fft_precompute(200) -> size 256; fft_with_precomputation; ifft ---- */
  plk_fft_plan* plan = NULL;
  if (check(plk_fft_precompute(PLK_FIELD_BLS12_377_SCALAR, 200, &plan), "plk_fft_precompute")) return 6;
  if (plk_fft_size(plan) != 256) { fprintf(stderr, "plan size\n"); return 6; }
  static uint64_t points[256 * 4], back[256 * 4];
  if (check(plk_fft(plan, kFftCoefficients, 200, points), "plk_fft")) return 6;
  if (memcmp(points, kFftExpected, sizeof(kFftExpected)) != 0) { fprintf(stderr, "fft_and_ifft: wrong evaluations\n"); return 7; }
  if (check(plk_ifft_pow2(plan, points, back, 256), "plk_ifft_pow2")) return 6;
  if (memcmp(back, kFftCoefficients, 200 * 4 * 8) != 0) { fprintf(stderr, "fft_and_ifft: wrong interpolation\n"); return 7; }
  for (int i = 200 * 4; i < 256 * 4; ++i) if (back[i] != 0) { fprintf(stderr, "fft_and_ifft: padding not zero\n"); return 7; }
  if (plk_ifft_pow2(plan, points, back, 200) != PLK_ENOTPOW2) { fprintf(stderr, "expected PLK_ENOTPOW2\n"); return 7; }
  plk_fft_free(plan);
  printf("abi_smoke ok: test_msm and fft_and_ifft through the C ABI, %llu kernel launches\n", (unsigned long long)plk_kernel_launch_count());
  return 0;
}
