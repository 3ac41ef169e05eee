// Depthwise 7x7 convolution (stride 1, zero pad 3) of the ConvNeXt block, NHWC bf16:
// /root/reference/ola_vlm/model/multimodal_encoder/clip_convnext_encoder.py:159-162 runs
// `stage(x)` of timm's ConvNeXt (un-vendored; timm==1.0.8, setup.py:20), whose block starts with
// `conv_dw = Conv2d(C, C, 7, padding=3, groups=C)`.
//
// Not a GEMM: 49 MACs per output element, no reuse across channels.  One CTA owns a
// 64-channel slice (one 128-byte line per pixel) of a TH x TW output tile: the (TH+6) x (TW+6)
// input halo tile is staged once in shared memory with zero-filling cp.async, the 49 x 64 filter
// taps sit beside it, and every thread keeps one channel PAIR of one output row — TW x 2 fp32
// accumulators — in registers, so an input value read from shared memory feeds 7 taps x 2
// channels.  A warp is 32 channel pairs of one row: every shared-memory access is one conflict-free
// 128-byte row and every global store is a full line.
// Algorithmic bytes: 2 * B*H*W*C * 2 (read + write; halos are L2 hits); 98 FLOP per element, so at
// ~12 FLOP/B the kernel sits on the fp32 FMA pipe, not on HBM.
#include "common.cuh"
#include "visper_b200.h"

namespace vpb {

#define ST(s) reinterpret_cast<cudaStream_t>(s)

constexpr int DW_K = 7, DW_PAD = 3, DW_CH = 64;

// acc += a * b on a channel pair as ONE packed instruction (IEEE fp32 FMA per half, so bit-identical to two fmaf)
__device__ __forceinline__ void fma_pair(float2& acc, const float2& a, const float2& b) {
  uint64_t ra = *reinterpret_cast<const uint64_t*>(&a), rb = *reinterpret_cast<const uint64_t*>(&b);
  uint64_t rc = *reinterpret_cast<uint64_t*>(&acc);
  asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(rc) : "l"(ra), "l"(rb));
  acc = *reinterpret_cast<float2*>(&rc);
}

template <int TH, int TW, bool PACKED>
__global__ void __launch_bounds__(TH * 32)
dwconv7x7_kernel(const bf16* __restrict__ in, const bf16* __restrict__ w49, const bf16* __restrict__ bias,
                 bf16* __restrict__ out, int H, int W, int C, int tiles_x) {
  constexpr int IH = TH + DW_K - 1, IW = TW + DW_K - 1;
  __shared__ __align__(16) bf16 s_in[IH * IW * DW_CH];
  __shared__ __align__(16) bf16 s_w[DW_K * DW_K * DW_CH];

  const int tile = blockIdx.x;
  const int ty0 = (tile / tiles_x) * TH, tx0 = (tile % tiles_x) * TW;
  const int c0 = blockIdx.y * DW_CH;
  const int b = blockIdx.z;
  const bf16* img = in + (int64_t)b * H * W * C + c0;

  // stage the halo tile (zero outside the image) and the filter slice
  for (int i = threadIdx.x; i < IH * IW * (DW_CH / 8); i += TH * 32) {
    const int ch8 = i % (DW_CH / 8);
    const int p = i / (DW_CH / 8);
    const int iy = ty0 - DW_PAD + p / IW, ix = tx0 - DW_PAD + p % IW;
    const bool ok = iy >= 0 && iy < H && ix >= 0 && ix < W;
    const bf16* g = ok ? img + ((int64_t)iy * W + ix) * C + ch8 * 8 : img;
    cp_async16(smem_u32(s_in + p * DW_CH + ch8 * 8), g, ok);
  }
  for (int i = threadIdx.x; i < DW_K * DW_K * (DW_CH / 8); i += TH * 32) {
    const int ch8 = i % (DW_CH / 8), tap = i / (DW_CH / 8);
    cp_async16(smem_u32(s_w + tap * DW_CH + ch8 * 8), w49 + (int64_t)tap * C + c0 + ch8 * 8, true);
  }
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  const int cp = threadIdx.x & 31;  // channel pair within the slice
  const int oy = threadIdx.x >> 5;  // output row within the tile
  const bf162* sin2 = reinterpret_cast<const bf162*>(s_in);
  const bf162* sw2 = reinterpret_cast<const bf162*>(s_w);

  float2 acc[TW];
  {
    const float2 bv = bias ? __bfloat1622float2(*reinterpret_cast<const bf162*>(bias + c0 + 2 * cp))
                           : make_float2(0.f, 0.f);
#pragma unroll
    for (int x = 0; x < TW; ++x) acc[x] = bv;
  }
#pragma unroll
  for (int ky = 0; ky < DW_K; ++ky) {
    float2 row[IW];
#pragma unroll
    for (int x = 0; x < IW; ++x) row[x] = __bfloat1622float2(sin2[((oy + ky) * IW + x) * (DW_CH / 2) + cp]);
#pragma unroll
    for (int kx = 0; kx < DW_K; ++kx) {
      const float2 wv = __bfloat1622float2(sw2[(ky * DW_K + kx) * (DW_CH / 2) + cp]);
#pragma unroll
      for (int x = 0; x < TW; ++x) {
        if constexpr (PACKED) {
          fma_pair(acc[x], row[x + kx], wv);
        } else {
          acc[x].x = fmaf(row[x + kx].x, wv.x, acc[x].x);
          acc[x].y = fmaf(row[x + kx].y, wv.y, acc[x].y);
        }
      }
    }
  }

  const int y = ty0 + oy;
  if (y < H) {
    bf16* orow = out + (((int64_t)b * H + y) * W) * C + c0 + 2 * cp;
#pragma unroll
    for (int x = 0; x < TW; ++x)
      if (tx0 + x < W) *reinterpret_cast<uint32_t*>(orow + (int64_t)(tx0 + x) * C) = pack2(acc[x].x, acc[x].y);
  }
}

}  // namespace vpb

using namespace vpb;

extern "C" int vpb_dwconv7x7_nhwc(const void* in, const void* w49, const void* bias, void* out, int B,
                                  int H, int W, int C, void* stream) {
  VPB_CHECK(B > 0 && H > 0 && W > 0 && C > 0 && C % DW_CH == 0 && B <= 65535,
            "dwconv7x7: bad shape B=%d H=%d W=%d C=%d (C must be a multiple of 64)", B, H, W, C);
  VPB_CHECK(in != out, "dwconv7x7: in-place is not supported (halo reads)");
  constexpr int TH = 8, TW = 16;
  const int tiles_x = (W + TW - 1) / TW, tiles_y = (H + TH - 1) / TH;
  dim3 grid(tiles_x * tiles_y, C / DW_CH, B);
  if (get_option(VPB_OPT_DWCONV_FFMA2))
    dwconv7x7_kernel<TH, TW, true><<<grid, TH * 32, 0, ST(stream)>>>((const bf16*)in, (const bf16*)w49,
                                                                    (const bf16*)bias, (bf16*)out, H, W, C, tiles_x);
  else
    dwconv7x7_kernel<TH, TW, false><<<grid, TH * 32, 0, ST(stream)>>>((const bf16*)in, (const bf16*)w49,
                                                                     (const bf16*)bias, (bf16*)out, H, W, C, tiles_x);
  VPB_LAUNCH_OK();
  return 0;
}
