// Layout / resampling kernels of the frozen DPT depth decoder (DAv2_Head) behind `depth_preds`:
// /root/reference/ola_vlm/model/aux_heads/da_v2_head.py:181-321, called under no_grad from
// base_ola_vlm.py:462-470.  Activations are NHWC bf16 ([B*H*W, C] rows), every convolution is an
// im2col (this file) + the tcgen05 GEMM with the bias / ReLU / skip-add fused in its epilogue,
// ConvTranspose2d(k = stride) is a GEMM + pixel shuffle.  All HBM-bound, 128-bit accesses.
#include "common.cuh"
#include "visper_b200.h"

namespace vpb {

#define ST(s) reinterpret_cast<cudaStream_t>(s)

static inline int grid_for(int64_t n, int threads) {
  int64_t g = (n + threads - 1) / threads;
  const int64_t cap = (int64_t)num_sms() * 32;
  return (int)(g < cap ? (g > 0 ? g : 1) : cap);
}

// out[(b,oy,ox), (ky,kx,c)] = relu?(in[b, oy*s-1+ky, ox*s-1+kx, c]) (zero outside), 3x3, pad 1.
// One CTA per output image row (b, oy); its threads walk the row's Wo*9*C/8 16-byte elements with 32-bit index
// arithmetic (round 1 ran one flat grid-stride loop with four 64-bit div/mod per element and reached 1.7 TB/s of
// stores: ALU-bound).  Consecutive threads write consecutive 16-byte elements of the [Wo, 9*C] output row block and
// read runs of C contiguous input channels.
__global__ void __launch_bounds__(256)
im2col3x3_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int B, int H, int W, int C,
                 int Ho, int Wo, int stride, int relu_in) {
  const int cv = C >> 3;
  const int per_px = 9 * cv;
  const int per_row = Wo * per_px;
  const int b = blockIdx.x / Ho, oy = blockIdx.x - b * Ho;
  const bf16* img = in + (int64_t)b * H * W * C;
  bf16* orow = out + (int64_t)blockIdx.x * per_row * 8;
  const int iy0 = oy * stride - 1;
  for (int j = threadIdx.x; j < per_row; j += 256) {
    const int ox = j / per_px;
    const int r = j - ox * per_px;
    const int tap = r / cv;
    const int c8 = r - tap * cv;
    const int ky = tap / 3;
    const int iy = iy0 + ky, ix = ox * stride - 1 + (tap - ky * 3);
    uint4 v = make_uint4(0, 0, 0, 0);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
      v = ldg16(img + ((int64_t)iy * W + ix) * C + c8 * 8);
      if (relu_in) {
        float f[8];
        unpack8(v, f);
#pragma unroll
        for (int q = 0; q < 8; ++q) f[q] = fmaxf(f[q], 0.f);
        v = pack8(f);
      }
    }
    stg16(orow + (int64_t)j * 8, v);
  }
}

// F.interpolate(mode="bilinear") on NHWC; AC = align_corners.  align_corners=False uses torch's
// half-pixel rule: src = max((dst + 0.5) * in/out - 0.5, 0).
template <bool AC>
__global__ void __launch_bounds__(256)
bilinear_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, int B, int Hi, int Wi, int Ho,
                int Wo, int C) {
  const int cv = C >> 3;
  const int64_t total = (int64_t)B * Ho * Wo * cv;
  const float sy = AC ? (Ho > 1 ? (float)(Hi - 1) / (float)(Ho - 1) : 0.f) : (float)Hi / (float)Ho;
  const float sx = AC ? (Wo > 1 ? (float)(Wi - 1) / (float)(Wo - 1) : 0.f) : (float)Wi / (float)Wo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    int64_t r = i / cv;
    const int ox = (int)(r % Wo);
    r /= Wo;
    const int oy = (int)(r % Ho);
    const int b = (int)(r / Ho);
    const float fy = AC ? sy * oy : fmaxf(sy * (oy + 0.5f) - 0.5f, 0.f);
    const float fx = AC ? sx * ox : fmaxf(sx * (ox + 0.5f) - 0.5f, 0.f);
    int y0 = (int)fy, x0 = (int)fx;
    if (y0 > Hi - 1) y0 = Hi - 1;
    if (x0 > Wi - 1) x0 = Wi - 1;
    const int y1 = y0 + (y0 < Hi - 1 ? 1 : 0), x1 = x0 + (x0 < Wi - 1 ? 1 : 0);
    const float ly = fy - y0, lx = fx - x0;
    const bf16* base = in + (int64_t)b * Hi * Wi * C + c8 * 8;
    float a[8], bb[8], c[8], d[8], o[8];
    unpack8(ldg16(base + ((int64_t)y0 * Wi + x0) * C), a);
    unpack8(ldg16(base + ((int64_t)y0 * Wi + x1) * C), bb);
    unpack8(ldg16(base + ((int64_t)y1 * Wi + x0) * C), c);
    unpack8(ldg16(base + ((int64_t)y1 * Wi + x1) * C), d);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      o[j] = (1.f - ly) * ((1.f - lx) * a[j] + lx * bb[j]) + ly * ((1.f - lx) * c[j] + lx * d[j]);
    stg16(out + i * 8, pack8(o));
  }
}

// ConvTranspose2d with kernel == stride == k after its GEMM:
// in [B*H*W, k*k*C] with column (ky*k+kx)*C + c  →  out[b, y*k+ky, x*k+kx, c] + bias[c]
__global__ void __launch_bounds__(256)
pixel_shuffle_kernel(const bf16* __restrict__ in, const bf16* __restrict__ bias, bf16* __restrict__ out,
                     int B, int H, int W, int C, int k) {
  const int cv = C >> 3;
  const int64_t total = (int64_t)B * H * W * k * k * cv;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % cv);
    int64_t r = i / cv;
    const int tap = (int)(r % (k * k));
    r /= (k * k);
    const int x = (int)(r % W);
    r /= W;
    const int y = (int)(r % H);
    const int b = (int)(r / H);
    float f[8], bv[8];
    unpack8(ldg16_stream(in + i * 8), f);
    if (bias) {
      unpack8(ldg16(bias + c8 * 8), bv);
#pragma unroll
      for (int j = 0; j < 8; ++j) f[j] += bv[j];
    }
    const int oy = y * k + tap / k, ox = x * k + tap % k;
    stg16(out + ((((int64_t)b * H * k + oy) * W * k + ox) * C + c8 * 8), pack8(f));
  }
}

// 1x1 convolution to ONE channel (+ReLU): out[p] = act(sum_c in[p,c] * w[c] + bias), fp32 out
__global__ void __launch_bounds__(256)
conv1x1_to1_kernel(const bf16* __restrict__ in, const bf16* __restrict__ w, const bf16* __restrict__ bias,
                   float* __restrict__ out, int64_t P, int C, int relu) {
  const int cv = C >> 3;
  const float b0 = bias ? __bfloat162float(bias[0]) : 0.f;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < P;
       p += (int64_t)gridDim.x * blockDim.x) {
    float acc = b0;
    for (int c8 = 0; c8 < cv; ++c8) {
      float a[8], ww[8];
      unpack8(ldg16_stream(in + p * C + c8 * 8), a);
      unpack8(ldg16(w + c8 * 8), ww);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(a[j], ww[j], acc);
    }
    out[p] = relu ? fmaxf(acc, 0.f) : acc;
  }
}

// per-image min-max normalisation (base_ola_vlm.py:466-469): one CTA per image
__global__ void __launch_bounds__(1024)
minmax_normalize_kernel(const float* __restrict__ in, float* __restrict__ out, int64_t n) {
  __shared__ float red[33];
  const float* x = in + (int64_t)blockIdx.x * n;
  float* y = out + (int64_t)blockIdx.x * n;
  float mn = INFINITY, mx = -INFINITY;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    const float v = x[i];
    mn = fminf(mn, v);
    mx = fmaxf(mx, v);
  }
  mx = block_max(mx, red);
  mn = -block_max(-mn, red);
  const float inv = 1.f / (mx - mn);  // constant image → inf/nan, as in the reference
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) y[i] = (x[i] - mn) * inv;
}

}  // namespace vpb

using namespace vpb;

extern "C" int vpb_im2col3x3_nhwc(const void* in, void* out, int B, int H, int W, int C, int stride,
                                  int relu_in, void* stream) {
  VPB_CHECK(B > 0 && H > 0 && W > 0 && C > 0 && C % 8 == 0 && (stride == 1 || stride == 2),
            "im2col3x3: bad shape B=%d H=%d W=%d C=%d stride=%d", B, H, W, C, stride);
  const int Ho = (H - 1) / stride + 1, Wo = (W - 1) / stride + 1;
  im2col3x3_kernel<<<B * Ho, 256, 0, ST(stream)>>>((const bf16*)in, (bf16*)out, B, H, W, C, Ho, Wo, stride, relu_in);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_bilinear_nhwc(const void* in, void* out, int B, int Hi, int Wi, int Ho, int Wo, int C,
                                 void* stream) {
  VPB_CHECK(B > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C % 8 == 0, "bilinear: bad shape");
  const int64_t total = (int64_t)B * Ho * Wo * (C / 8);
  bilinear_kernel<true><<<grid_for(total, 256), 256, 0, ST(stream)>>>((const bf16*)in, (bf16*)out, B, Hi, Wi,
                                                                      Ho, Wo, C);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_bilinear_nhwc_half_pixel(const void* in, void* out, int B, int Hi, int Wi, int Ho, int Wo,
                                            int C, void* stream) {
  VPB_CHECK(B > 0 && Hi > 0 && Wi > 0 && Ho > 0 && Wo > 0 && C % 8 == 0, "bilinear: bad shape");
  const int64_t total = (int64_t)B * Ho * Wo * (C / 8);
  bilinear_kernel<false><<<grid_for(total, 256), 256, 0, ST(stream)>>>((const bf16*)in, (bf16*)out, B, Hi, Wi,
                                                                       Ho, Wo, C);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_pixel_shuffle_nhwc(const void* in, const void* bias, void* out, int B, int H, int W,
                                      int C, int k, void* stream) {
  VPB_CHECK(B > 0 && H > 0 && W > 0 && C % 8 == 0 && k > 0, "pixel_shuffle: bad shape");
  const int64_t total = (int64_t)B * H * W * k * k * (C / 8);
  pixel_shuffle_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>((const bf16*)in, (const bf16*)bias,
                                                                     (bf16*)out, B, H, W, C, k);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_conv1x1_to1(const void* in, const void* w, const void* bias, float* out, int64_t P,
                               int C, int relu, void* stream) {
  VPB_CHECK(P > 0 && C % 8 == 0, "conv1x1_to1: bad shape");
  conv1x1_to1_kernel<<<grid_for(P, 256), 256, 0, ST(stream)>>>((const bf16*)in, (const bf16*)w,
                                                              (const bf16*)bias, out, P, C, relu);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_minmax_normalize(const float* in, float* out, int B, int64_t n, void* stream) {
  VPB_CHECK(B > 0 && n > 0, "minmax_normalize: bad shape");
  minmax_normalize_kernel<<<B, 1024, 0, ST(stream)>>>(in, out, n);
  VPB_LAUNCH_OK();
  return 0;
}
