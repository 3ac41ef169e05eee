// Stand-alone check + timing of vpb_dwconv7x7_nhwc through the C ABI (no Python, starts in < 1 s):
//   parity against oracle/c/dwconv_ref.c on ragged small shapes (bit-level: both round fp32 sums to bf16,
//   so the tolerance is 1 bf16 ulp of the result), then CUDA-event timing at the ConvNeXt-XXL stage shapes
//   (768 px, B = 8).  Built by tools/build_dwconv_check.sh into tools/_bin/; prints JSON lines.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "visper_b200.h"

extern "C" void oracle_dwconv7x7_nhwc(const float*, const float*, const float*, float*, int, int, int, int);

static uint32_t rng = 12345u;
static float frand() {
  rng = rng * 1664525u + 1013904223u;
  return ((rng >> 8) & 0xffff) / 32768.f - 1.f;
}
static float bf(float x) { return __bfloat162float(__float2bfloat16(x)); }

static const char* g_variant = "scalar";

static int parity(int B, int H, int W, int C) {
  const size_t n = (size_t)B * H * W * C;
  std::vector<float> in(n), w(C * 49), bias(C), ref(n);
  for (auto& v : in) v = bf(frand());
  for (auto& v : w) v = bf(frand() * 0.2f);
  for (auto& v : bias) v = bf(frand());
  oracle_dwconv7x7_nhwc(in.data(), w.data(), bias.data(), ref.data(), B, H, W, C);
  std::vector<__nv_bfloat16> hin(n), hw(49 * C), hb(C), hout(n);
  for (size_t i = 0; i < n; ++i) hin[i] = __float2bfloat16(in[i]);
  for (int c = 0; c < C; ++c) {
    hb[c] = __float2bfloat16(bias[c]);
    for (int t = 0; t < 49; ++t) hw[(size_t)t * C + c] = __float2bfloat16(w[c * 49 + t]);
  }
  __nv_bfloat16 *din, *dw, *db, *dout;
  cudaMalloc(&din, n * 2); cudaMalloc(&dw, 49 * C * 2); cudaMalloc(&db, C * 2); cudaMalloc(&dout, n * 2);
  cudaMemcpy(din, hin.data(), n * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), 49 * C * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), C * 2, cudaMemcpyHostToDevice);
  cudaMemset(dout, 0xff, n * 2);
  int rc = vpb_dwconv7x7_nhwc(din, dw, db, dout, B, H, W, C, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(hout.data(), dout, n * 2, cudaMemcpyDeviceToHost);
  double maxerr = 0, maxref = 0;
  size_t bad = 0;
  for (size_t i = 0; i < n; ++i) {
    const float g = __bfloat162float(hout[i]);
    const double d = fabs((double)g - ref[i]);
    const double tol = fabs(ref[i]) * 0.0079 + 1e-3;  // 1 bf16 ulp (2^-7) + fp32-vs-double slack
    if (!(d <= tol)) ++bad;
    if (d > maxerr) maxerr = d;
    if (fabs(ref[i]) > maxref) maxref = fabs(ref[i]);
  }
  printf("{\"check\": \"dwconv7x7 parity\", \"variant\": \"%s\", \"B\": %d, \"H\": %d, \"W\": %d, \"C\": %d, \"rc\": %d, \"cuda\": \"%s\", "
         "\"max_abs_err\": %.5g, \"max_abs_ref\": %.5g, \"bad\": %zu, \"ok\": %s}\n",
         g_variant, B, H, W, C, rc, cudaGetErrorString(e), maxerr, maxref, bad, (rc == 0 && e == cudaSuccess && bad == 0) ? "true" : "false");
  cudaFree(din); cudaFree(dw); cudaFree(db); cudaFree(dout);
  return (rc == 0 && e == cudaSuccess && bad == 0) ? 0 : 1;
}

static void timing(int B, int H, int W, int C, int iters) {
  const size_t n = (size_t)B * H * W * C;
  __nv_bfloat16 *din, *dw, *db, *dout, *flush;
  const size_t fl = 256u << 20;
  cudaMalloc(&din, n * 2); cudaMalloc(&dw, 49 * C * 2); cudaMalloc(&db, C * 2); cudaMalloc(&dout, n * 2);
  cudaMalloc(&flush, fl);
  cudaMemset(din, 0, n * 2); cudaMemset(dw, 0, 49 * C * 2); cudaMemset(db, 0, C * 2);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int i = 0; i < 2; ++i) vpb_dwconv7x7_nhwc(din, dw, db, dout, B, H, W, C, nullptr);
  float total = 0;
  for (int i = 0; i < iters; ++i) {
    cudaMemsetAsync(flush, i, fl, nullptr);  // L2 flush between timed launches
    cudaEventRecord(e0, nullptr);
    vpb_dwconv7x7_nhwc(din, dw, db, dout, B, H, W, C, nullptr);
    cudaEventRecord(e1, nullptr);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); total += ms;
  }
  const double ms = total / iters, bytes = 2.0 * n * 2, flop = 98.0 * n;
  printf("{\"bench\": \"dwconv7x7\", \"variant\": \"%s\", \"B\": %d, \"H\": %d, \"W\": %d, \"C\": %d, \"ms\": %.4f, \"GBps\": %.1f, "
         "\"TFLOPs_fp32\": %.2f, \"l2_flush\": true}\n", g_variant, B, H, W, C, ms, bytes / ms * 1e-6, flop / ms * 1e-9);
  cudaFree(din); cudaFree(dw); cudaFree(db); cudaFree(dout); cudaFree(flush);
}

// scalar and packed (VPB_OPT_DWCONV_FFMA2) variants must agree bit for bit
static int identical(int B, int H, int W, int C) {
  const size_t n = (size_t)B * H * W * C;
  std::vector<__nv_bfloat16> hin(n), hw(49 * C), hb(C), o0(n), o1(n);
  for (auto& v : hin) v = __float2bfloat16(frand());
  for (auto& v : hw) v = __float2bfloat16(frand() * 0.2f);
  for (auto& v : hb) v = __float2bfloat16(frand());
  __nv_bfloat16 *din, *dw, *db, *dout;
  cudaMalloc(&din, n * 2); cudaMalloc(&dw, 49 * C * 2); cudaMalloc(&db, C * 2); cudaMalloc(&dout, n * 2);
  cudaMemcpy(din, hin.data(), n * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(dw, hw.data(), 49 * C * 2, cudaMemcpyHostToDevice);
  cudaMemcpy(db, hb.data(), C * 2, cudaMemcpyHostToDevice);
  vpb_set_option(VPB_OPT_DWCONV_FFMA2, 0);
  vpb_dwconv7x7_nhwc(din, dw, db, dout, B, H, W, C, nullptr);
  cudaMemcpy(o0.data(), dout, n * 2, cudaMemcpyDeviceToHost);
  vpb_set_option(VPB_OPT_DWCONV_FFMA2, 1);
  vpb_dwconv7x7_nhwc(din, dw, db, dout, B, H, W, C, nullptr);
  cudaError_t e = cudaDeviceSynchronize();
  cudaMemcpy(o1.data(), dout, n * 2, cudaMemcpyDeviceToHost);
  vpb_set_option(VPB_OPT_DWCONV_FFMA2, 0);
  const bool same = memcmp(o0.data(), o1.data(), n * 2) == 0;
  printf("{\"check\": \"dwconv7x7 packed == scalar\", \"B\": %d, \"H\": %d, \"W\": %d, \"C\": %d, \"cuda\": \"%s\", \"bit_identical\": %s}\n",
         B, H, W, C, cudaGetErrorString(e), same ? "true" : "false");
  cudaFree(din); cudaFree(dw); cudaFree(db); cudaFree(dout);
  return same && e == cudaSuccess ? 0 : 1;
}

int main(int argc, char** argv) {
  int fails = 0;
  for (int packed = 0; packed < 2; ++packed) {
    vpb_set_option(VPB_OPT_DWCONV_FFMA2, packed);
    g_variant = packed ? "ffma2" : "scalar";
    fails += parity(1, 5, 9, 64);     // smaller than the halo, single tile
    fails += parity(2, 24, 24, 128);  // W = 16 + 8 ragged, H = 3 tiles
    fails += parity(1, 17, 35, 192);  // ragged both ways
  }
  fails += identical(2, 48, 48, 256);
  printf("{\"dwconv_parity_failures\": %d}\n", fails);
  if (argc > 1 && atoi(argv[1]) == 0) return fails;
  const int B = 8;
  for (int rep = 0; rep < 2; ++rep)
    for (int packed = 0; packed < 2; ++packed) {  // alternating A/B on one box
      vpb_set_option(VPB_OPT_DWCONV_FFMA2, packed);
      g_variant = packed ? "ffma2" : "scalar";
      timing(B, 192, 192, 384, 5);
      timing(B, 96, 96, 768, 5);
      timing(B, 48, 48, 1536, 5);
      timing(B, 24, 24, 3072, 5);
    }
  return fails;
}
