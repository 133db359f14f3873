// GPU self-test for the tcgen05 GEMM core: compares tic_gemm_bf16 against the SIMT reference for one
// (a_mn_major, b_mn_major, M, N, K) combination per process, so a trapped kernel cannot poison later cases.
//   usage: selftest a_mn b_mn M N K [d_bf16] [iters]
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "../../include/tic_b200.h"

#define CK(x)                                                                      \
  do {                                                                             \
    cudaError_t e = (x);                                                           \
    if (e != cudaSuccess) {                                                        \
      printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
      return 2;                                                                    \
    }                                                                              \
  } while (0)

static uint32_t rng_state = 12345u;
static float frand() {
  rng_state = rng_state * 1664525u + 1013904223u;
  return ((rng_state >> 8) & 0xFFFF) / 65536.0f - 0.5f;
}

int main(int argc, char** argv) {
  if (argc < 6) {
    printf("usage: selftest a_mn b_mn M N K [d_bf16] [iters]\n");
    return 1;
  }
  const int a_mn = atoi(argv[1]), b_mn = atoi(argv[2]), M = atoi(argv[3]), N = atoi(argv[4]), K = atoi(argv[5]);
  const int d_bf16 = argc > 6 ? atoi(argv[6]) : 0;
  const int iters = argc > 7 ? atoi(argv[7]) : 0;
  auto pad8 = [](int x) { return (x + 7) / 8 * 8; };
  const int64_t lda = a_mn ? pad8(M) : pad8(K), ldb = b_mn ? pad8(N) : pad8(K), ldd = pad8(N);
  const int64_t a_rows = a_mn ? K : M, b_rows = b_mn ? K : N;
  std::vector<__nv_bfloat16> hA(a_rows * lda), hB(b_rows * ldb);
  for (auto& x : hA) x = __float2bfloat16(frand());
  for (auto& x : hB) x = __float2bfloat16(frand());
  std::vector<float> hbias(N);
  for (auto& x : hbias) x = frand();
  __nv_bfloat16 *dA, *dB;
  float* dbias;
  void *dD, *dR;
  const size_t dbytes = (size_t)M * ldd * (d_bf16 ? 2 : 4);
  CK(cudaMalloc(&dA, hA.size() * 2));
  CK(cudaMalloc(&dB, hB.size() * 2));
  CK(cudaMalloc(&dbias, N * 4));
  CK(cudaMalloc(&dD, dbytes));
  CK(cudaMalloc(&dR, dbytes));
  CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dbias, hbias.data(), N * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dD, 0xFF, dbytes));
  CK(cudaMemset(dR, 0xFF, dbytes));
  int rc = tic_gemm_bf16(dA, nullptr, lda, a_mn, dB, nullptr, ldb, b_mn, dD, nullptr, ldd, d_bf16, M, N, K, 0.5f, dbias, 0, 0, nullptr);
  if (rc) {
    printf("tic_gemm_bf16 rc=%d: %s\n", rc, tic_last_error_string());
    return 3;
  }
  CK(cudaDeviceSynchronize());
  rc = tic_gemm_bf16_simt(dA, lda, a_mn, dB, ldb, b_mn, dR, ldd, d_bf16, M, N, K, 0.5f, dbias, 0, nullptr);
  if (rc) {
    printf("simt rc=%d: %s\n", rc, tic_last_error_string());
    return 3;
  }
  CK(cudaDeviceSynchronize());
  std::vector<uint8_t> hD(dbytes), hR(dbytes);
  CK(cudaMemcpy(hD.data(), dD, dbytes, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(hR.data(), dR, dbytes, cudaMemcpyDeviceToHost));
  double max_err = 0, max_ref = 0;
  int bad = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float x, r;
      if (d_bf16) {
        x = __bfloat162float(reinterpret_cast<__nv_bfloat16*>(hD.data())[m * ldd + n]);
        r = __bfloat162float(reinterpret_cast<__nv_bfloat16*>(hR.data())[m * ldd + n]);
      } else {
        x = reinterpret_cast<float*>(hD.data())[m * ldd + n];
        r = reinterpret_cast<float*>(hR.data())[m * ldd + n];
      }
      double e = fabs((double)x - r);
      if (!(e <= 1e30)) e = 1e30;
      if (e > max_err) max_err = e;
      if (fabs(r) > max_ref) max_ref = fabs(r);
      if (e > (d_bf16 ? 2e-2 : 2e-3) * (1.0 + fabs(r)) && bad < 5) {
        printf("  mismatch m=%d n=%d got %g ref %g\n", m, n, x, r);
        ++bad;
      }
    }
  printf("a_mn=%d b_mn=%d M=%d N=%d K=%d bf16out=%d  max_err=%.3e max_ref=%.3e  %s\n", a_mn, b_mn, M, N, K, d_bf16,
         max_err, max_ref, bad ? "FAIL" : "PASS");
  if (iters > 0 && !bad) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) tic_gemm_bf16(dA, nullptr, lda, a_mn, dB, nullptr, ldb, b_mn, dD, nullptr, ldd, d_bf16, M, N, K, 0.5f, dbias, 0, 0, nullptr);
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) tic_gemm_bf16(dA, nullptr, lda, a_mn, dB, nullptr, ldb, b_mn, dD, nullptr, ldd, d_bf16, M, N, K, 0.5f, dbias, 0, 0, nullptr);
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    ms /= iters;
    printf("  time %.3f ms  %.1f TFLOP/s\n", ms, 2.0 * M * N * K / ms * 1e-9);
    // the same launches as ONE CUDA graph (serialised kernel nodes): device-side time per kernel, no host submission cost
    cudaStream_t st;
    cudaStreamCreate(&st);
    cudaGraph_t g;
    cudaGraphExec_t ge;
    CK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
    for (int i = 0; i < iters; ++i) tic_gemm_bf16(dA, nullptr, lda, a_mn, dB, nullptr, ldb, b_mn, dD, nullptr, ldd, d_bf16, M, N, K, 0.5f, dbias, 0, 0, st);
    CK(cudaStreamEndCapture(st, &g));
    CK(cudaGraphInstantiate(&ge, g, 0));
    CK(cudaGraphLaunch(ge, st));
    CK(cudaStreamSynchronize(st));
    cudaEventRecord(e0, st);
    CK(cudaGraphLaunch(ge, st));
    cudaEventRecord(e1, st);
    CK(cudaEventSynchronize(e1));
    cudaEventElapsedTime(&ms, e0, e1);
    printf("  as a graph of %d serialised nodes: %.2f us per kernel\n", iters, 1e3 * ms / iters);
  }
  return bad ? 4 : 0;
}
