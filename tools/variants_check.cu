// Stand-alone A/B of the kernel variants that are OFF by default (written after round 1's GPU budget was spent),
// through the C ABI only — no Python, so a run costs seconds of GPU time instead of a torch import:
//   gemm_epi8   VPB_OPT_GEMM_EPI8      CTA-pair GEMM, eight epilogue warps (K <= 1024): bit identity + times
//   gather_flat VPB_OPT_GATHER_FLAT    grid-stride row gather: bit identity + times
//   attn_tc64   VPB_OPT_ATTN_FWD_TC64  ViT attention (head_dim 64, non-causal) on the tcgen05 forward: max diff + times
// Every timing alternates off / on inside one process (same box, same clocks) with an L2 flush before each launch.
// Built by tools/build_dwconv_check.sh into tools/_bin/variants_check; prints JSON lines; exit code = failures.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "visper_b200.h"

typedef __nv_bfloat16 bf16;
static uint32_t rng = 2024u;
static float frand() {
  rng = rng * 1664525u + 1013904223u;
  return ((rng >> 8) & 0xffff) / 32768.f - 1.f;
}
static bf16* dev_rand(size_t n, float scale) {
  std::vector<bf16> h(n);
  for (auto& v : h) v = __float2bfloat16(frand() * scale);
  bf16* d;
  cudaMalloc(&d, n * 2);
  cudaMemcpy(d, h.data(), n * 2, cudaMemcpyHostToDevice);
  return d;
}
static void* g_flush = nullptr;
static void flush_l2(int tag) {
  if (!g_flush) cudaMalloc(&g_flush, 256u << 20);
  cudaMemsetAsync(g_flush, tag & 0xff, 256u << 20, nullptr);
}
template <class F>
static float time_ms(F&& launch, int iters) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  launch();
  float total = 0;
  for (int i = 0; i < iters; ++i) {
    flush_l2(i);
    cudaEventRecord(e0, nullptr);
    launch();
    cudaEventRecord(e1, nullptr);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    total += ms;
  }
  return total / iters;
}
static double max_abs_diff(const bf16* a, const bf16* b, size_t n, double* maxref) {
  std::vector<bf16> ha(n), hb(n);
  cudaMemcpy(ha.data(), a, n * 2, cudaMemcpyDeviceToHost);
  cudaMemcpy(hb.data(), b, n * 2, cudaMemcpyDeviceToHost);
  double d = 0, r = 0;
  for (size_t i = 0; i < n; ++i) {
    const double x = __bfloat162float(ha[i]), y = __bfloat162float(hb[i]);
    if (!(fabs(x - y) <= d)) d = fabs(x - y);  // NaN propagates into d
    if (fabs(y) > r) r = fabs(y);
  }
  if (maxref) *maxref = r;
  return d;
}
static bool same_bits(const void* a, const void* b, size_t bytes) {
  std::vector<uint8_t> ha(bytes), hb(bytes);
  cudaMemcpy(ha.data(), a, bytes, cudaMemcpyDeviceToHost);
  cudaMemcpy(hb.data(), b, bytes, cudaMemcpyDeviceToHost);
  return memcmp(ha.data(), hb.data(), bytes) == 0;
}

static int gemm_epi8(int M, int N, int K, int act, bool res) {
  bf16 *A = dev_rand((size_t)M * K, 1.f), *W = dev_rand((size_t)N * K, 1.f / sqrtf((float)K)), *b = dev_rand(N, 1.f);
  bf16* R = res ? dev_rand((size_t)M * N, 1.f) : nullptr;
  bf16 *C0, *C1;
  cudaMalloc(&C0, (size_t)M * N * 2);
  cudaMalloc(&C1, (size_t)M * N * 2);
  int rc = 0;
  auto run = [&](bf16* C) { rc |= vpb_gemm_bf16(A, K, 0, W, K, 0, C, N, M, N, K, act, b, R, N, nullptr, 0, nullptr); };
  vpb_set_option(VPB_OPT_GEMM_EPI8, 0);
  run(C0);
  vpb_set_option(VPB_OPT_GEMM_EPI8, 1);
  run(C1);
  cudaError_t e = cudaDeviceSynchronize();
  const bool same = same_bits(C0, C1, (size_t)M * N * 2);
  float ms[2][2];
  for (int rep = 0; rep < 2; ++rep)
    for (int v = 0; v < 2; ++v) {
      vpb_set_option(VPB_OPT_GEMM_EPI8, v);
      ms[rep][v] = time_ms([&] { run(C1); }, 5);
    }
  vpb_set_option(VPB_OPT_GEMM_EPI8, 0);
  const double fl = 2.0 * M * N * K;
  printf("{\"variant\": \"gemm_epi8\", \"M\": %d, \"N\": %d, \"K\": %d, \"act\": %d, \"residual\": %s, \"rc\": %d, \"cuda\": \"%s\", "
         "\"bit_identical\": %s, \"ms_off\": [%.4f, %.4f], \"ms_on\": [%.4f, %.4f], \"tflops_off\": %.0f, \"tflops_on\": %.0f}\n",
         M, N, K, act, res ? "true" : "false", rc, cudaGetErrorString(e), same ? "true" : "false", ms[0][0], ms[1][0], ms[0][1],
         ms[1][1], fl / ms[1][0] * 1e-9, fl / ms[1][1] * 1e-9);
  cudaFree(A); cudaFree(W); cudaFree(b); cudaFree(R); cudaFree(C0); cudaFree(C1);
  return (rc == 0 && e == cudaSuccess && same) ? 0 : 1;
}

static int gather_flat(int rows_src, int n, int D) {
  bf16* src = dev_rand((size_t)rows_src * D, 1.f);
  std::vector<int> hidx(n);
  for (int i = 0; i < n; ++i) {
    rng = rng * 1664525u + 1013904223u;
    hidx[i] = (rng >> 4) % 17 == 0 ? -1 : (int)((rng >> 8) % (uint32_t)rows_src);
  }
  int* idx;
  cudaMalloc(&idx, n * 4);
  cudaMemcpy(idx, hidx.data(), n * 4, cudaMemcpyHostToDevice);
  bf16 *o0, *o1;
  cudaMalloc(&o0, (size_t)n * D * 2);
  cudaMalloc(&o1, (size_t)n * D * 2);
  int rc = 0;
  auto run = [&](bf16* o) { rc |= vpb_gather_rows(o, D, n, D, nullptr, idx, src, D, nullptr, 0, nullptr, 0, nullptr, 0, nullptr); };
  vpb_set_option(VPB_OPT_GATHER_FLAT, 0);
  run(o0);
  vpb_set_option(VPB_OPT_GATHER_FLAT, 1);
  run(o1);
  cudaError_t e = cudaDeviceSynchronize();
  const bool same = same_bits(o0, o1, (size_t)n * D * 2);
  float ms[2];
  for (int v = 0; v < 2; ++v) {
    vpb_set_option(VPB_OPT_GATHER_FLAT, v);
    ms[v] = time_ms([&] { run(o1); }, 5);
  }
  vpb_set_option(VPB_OPT_GATHER_FLAT, 0);
  const double bytes = 2.0 * n * D * 2;
  printf("{\"variant\": \"gather_flat\", \"rows\": %d, \"D\": %d, \"rc\": %d, \"cuda\": \"%s\", \"bit_identical\": %s, "
         "\"ms_off\": %.4f, \"ms_on\": %.4f, \"GBps_off\": %.0f, \"GBps_on\": %.0f}\n",
         n, D, rc, cudaGetErrorString(e), same ? "true" : "false", ms[0], ms[1], bytes / ms[0] * 1e-6, bytes / ms[1] * 1e-6);
  cudaFree(src); cudaFree(idx); cudaFree(o0); cudaFree(o1);
  return (rc == 0 && e == cudaSuccess && same) ? 0 : 1;
}

// forward attention on the packed [B*S, 3*H*hd] projection: option `opt` off vs on
static int attn_ab(const char* name, int opt, int B, int H, int KVH, int S, int hd, int causal, double tol, int opt2 = -1) {
  auto set = [&](int v) {
    vpb_set_option(opt, v);
    if (opt2 >= 0) vpb_set_option(opt2, v);
  };
  const int ld = (H + 2 * KVH) * hd;
  bf16* qkv = dev_rand((size_t)B * S * ld, 1.f);
  bf16 *o0, *o1;
  float *l0, *l1;
  cudaMalloc(&o0, (size_t)B * S * H * hd * 2);
  cudaMalloc(&o1, (size_t)B * S * H * hd * 2);
  cudaMalloc(&l0, (size_t)B * H * S * 4);
  cudaMalloc(&l1, (size_t)B * H * S * 4);
  int rc = 0;
  auto run = [&](bf16* o, float* l) {
    rc |= vpb_attn_fwd(qkv, ld, qkv + H * hd, ld, qkv + (H + KVH) * hd, ld, nullptr, 0, nullptr, 0, o, H * hd, l, B, H, KVH, S, S,
                       0, hd, 1.f / sqrtf((float)hd), causal, 0, nullptr);
  };
  set(0);
  run(o0, l0);
  set(1);
  run(o1, l1);
  cudaError_t e = cudaDeviceSynchronize();
  double mref = 0;
  const double d = max_abs_diff(o1, o0, (size_t)B * S * H * hd, &mref);
  float ms[2][2];
  for (int rep = 0; rep < 2; ++rep)
    for (int v = 0; v < 2; ++v) {
      set(v);
      ms[rep][v] = time_ms([&] { run(o1, l1); }, 5);
    }
  set(0);
  const double fl = 4.0 * B * H * (double)S * S * hd * (causal ? 0.5 : 1.0);
  const bool ok = rc == 0 && e == cudaSuccess && d <= tol * mref;
  printf("{\"variant\": \"%s\", \"B\": %d, \"H\": %d, \"KVH\": %d, \"S\": %d, \"hd\": %d, \"causal\": %d, \"rc\": %d, \"cuda\": \"%s\", "
         "\"max_abs_diff\": %.4g, \"max_abs_ref\": %.4g, \"ok\": %s, \"ms_off\": [%.4f, %.4f], \"ms_on\": [%.4f, %.4f], "
         "\"tflops_off\": %.0f, \"tflops_on\": %.0f}\n",
         name, B, H, KVH, S, hd, causal, rc, cudaGetErrorString(e), d, mref, ok ? "true" : "false", ms[0][0], ms[1][0], ms[0][1],
         ms[1][1], fl / ms[1][0] * 1e-9, fl / ms[1][1] * 1e-9);
  cudaFree(qkv); cudaFree(o0); cudaFree(o1); cudaFree(l0); cudaFree(l1);
  return ok ? 0 : 1;
}

int main(int argc, char** argv) {
  const char* only = argc > 1 ? argv[1] : "";
  auto want = [&](const char* n) { return !only[0] || strcmp(only, n) == 0; };
  int fails = 0;
  if (want("gemm_epi8")) {
    fails += gemm_epi8(320000, 768, 192, VPB_ACT_GELU, false);    // Swin-L stage 1 fc1
    fails += gemm_epi8(320000, 192, 768, VPB_ACT_NONE, true);     // Swin-L stage 1 fc2 (+residual): N < 256 → 1-CTA kernel, unchanged
    fails += gemm_epi8(294912, 1536, 384, VPB_ACT_GELU, false);   // ConvNeXt-XXL stage 0 fc1
    fails += gemm_epi8(73728, 3072, 768, VPB_ACT_GELU, false);    // ConvNeXt-XXL stage 1 fc1
    fails += gemm_epi8(73728, 768, 3072, VPB_ACT_NONE, true);     // K > 1024: must take the default kernel
    fails += gemm_epi8(20000, 3072, 768, VPB_ACT_GELU, false);    // Swin-L stage 3 fc1
    fails += gemm_epi8(8192, 520, 200, VPB_ACT_GELU, true);       // ragged N and K
  }
  if (want("gather_flat")) {
    fails += gather_flat(320000, 332928, 192);                    // Swin-L stage 1 window partition
    fails += gather_flat(20000, 28800, 768);                      // stage 3
    fails += gather_flat(73728, 18432, 768);                      // ConvNeXt 2x2 merge block
    fails += gather_flat(4096, 16384, 4096);                      // splice-sized rows: D > 2048 keeps the default kernel
  }
  if (want("attn_tc64")) {
    fails += attn_ab("attn_tc64", VPB_OPT_ATTN_FWD_TC64, 8, 16, 16, 577, 64, 0, 2e-2);   // CLIP ViT-L/14-336
    fails += attn_ab("attn_tc64", VPB_OPT_ATTN_FWD_TC64, 8, 16, 16, 1370, 64, 0, 2e-2);  // DINOv2-L @518
    fails += attn_ab("attn_tc64", VPB_OPT_ATTN_FWD_TC64, 2, 4, 4, 70, 64, 0, 2e-2);      // shorter than one tile
  }
  printf("{\"variants_check_failures\": %d}\n", fails);
  return fails;
}
