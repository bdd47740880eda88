/* Plain-C host program driving libfs2b200.so through include/fs2_b200.h only (no Python, no torch):
 * LengthRegulator of a tiny batch, checked against a scalar CPU loop that follows
 * core/duration_modeling/length_regulator.py:38-95.  Built and run by tests/test_c_abi.py on the GPU box:
 *   gcc -I include -I /usr/local/cuda/include tests/c_abi/length_regulator_host.c -o <exe> \
 *       -L fastspeech2_b200 -lfs2b200 -L /usr/local/cuda/lib64 -lcudart -Wl,-rpath,<repo>/fastspeech2_b200
 */
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "fs2_b200.h"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e_), __LINE__); return 2; } } while (0)
#define FS(x) do { int r_ = (x); if (r_ != FS2_OK) { printf("fs2 error %d: %s (line %d)\n", r_, fs2_last_error(), __LINE__); return 3; } } while (0)

int main(void) {
  enum { B = 3, T = 5, C = 8 };
  const int64_t ilens[B] = {5, 3, 4};
  int64_t ds[B][T] = {{2, 0, 3, 1, 1}, {1, 4, 2, 9, 9}, {0, 0, 0, 0, 7}}; /* row 2: all-zero slice -> ones, in place */
  float hs[B][T][C];
  for (int b = 0; b < B; ++b) for (int t = 0; t < T; ++t) for (int c = 0; c < C; ++c) hs[b][t][c] = (float)(100 * b + 10 * t + c) + 0.25f;
  printf("%s\n", fs2_version());

  float *d_hs, *d_out; int64_t *d_ds, *d_il, *d_ol, *d_stats; int32_t* d_cum;
  CK(cudaMalloc((void**)&d_hs, sizeof hs)); CK(cudaMalloc((void**)&d_ds, sizeof ds)); CK(cudaMalloc((void**)&d_il, sizeof ilens));
  CK(cudaMalloc((void**)&d_ol, B * 8)); CK(cudaMalloc((void**)&d_stats, 16)); CK(cudaMalloc((void**)&d_cum, B * T * 4));
  CK(cudaMemcpy(d_hs, hs, sizeof hs, cudaMemcpyHostToDevice)); CK(cudaMemcpy(d_ds, ds, sizeof ds, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(d_il, ilens, sizeof ilens, cudaMemcpyHostToDevice));

  FS(fs2_length_plan(d_ds, FS2_DUR_I64, d_il, 1.0f, B, T, /*mutate_ds=*/1, d_cum, d_ol, d_stats, NULL));
  int64_t stats[2], olens[B];
  CK(cudaMemcpy(stats, d_stats, 16, cudaMemcpyDeviceToHost));          /* the path's single host sync */
  CK(cudaMemcpy(olens, d_ol, sizeof olens, cudaMemcpyDeviceToHost));
  const int L = (int)stats[0];
  CK(cudaMalloc((void**)&d_out, (size_t)B * L * C * 4));
  FS(fs2_length_gather(d_hs, d_cum, d_il, B, T, C, d_out, L, NULL));
  float* out = (float*)malloc((size_t)B * L * C * 4);
  CK(cudaMemcpy(out, d_out, (size_t)B * L * C * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(ds, d_ds, sizeof ds, cudaMemcpyDeviceToHost));

  /* scalar restatement of the reference */
  int bad = 0, lmax = 0;
  for (int b = 0; b < B; ++b) {
    int64_t ref_d[T]; int64_t sum = 0;
    const int64_t orig[B][T] = {{2, 0, 3, 1, 1}, {1, 4, 2, 9, 9}, {0, 0, 0, 0, 7}};
    for (int t = 0; t < ilens[b]; ++t) { ref_d[t] = orig[b][t]; sum += ref_d[t]; }
    if (sum == 0) for (int t = 0; t < ilens[b]; ++t) ref_d[t] = 1;
    int j = 0;
    for (int t = 0; t < ilens[b]; ++t) {
      if (ds[b][t] != ref_d[t]) { printf("ds[%d][%d] = %lld, expected %lld\n", b, t, (long long)ds[b][t], (long long)ref_d[t]); ++bad; }
      for (int r = 0; r < ref_d[t]; ++r, ++j)
        for (int c = 0; c < C; ++c)
          if (j < L && memcmp(&out[((size_t)b * L + j) * C + c], &hs[b][t][c], 4)) ++bad;
    }
    if (olens[b] != j) { printf("olens[%d] = %lld, expected %d\n", b, (long long)olens[b], j); ++bad; }
    for (; j < L; ++j) for (int c = 0; c < C; ++c) if (out[((size_t)b * L + j) * C + c] != 0.0f) ++bad;
    if (olens[b] > lmax) lmax = (int)olens[b];
  }
  if (lmax != L || stats[1] != 0) ++bad;
  /* error path: NULL argument must be rejected with a message, not crash */
  if (fs2_length_gather(NULL, d_cum, d_il, B, T, C, d_out, L, NULL) != FS2_ERR_INVALID || !strstr(fs2_last_error(), "null")) ++bad;
  printf("Lmax=%d olens=%lld,%lld,%lld mismatches=%d\n", L, (long long)olens[0], (long long)olens[1], (long long)olens[2], bad);
  printf(bad ? "C_ABI_FAIL\n" : "C_ABI_OK\n");
  return bad ? 1 : 0;
}
