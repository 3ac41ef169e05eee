/* TEST INFRASTRUCTURE — CPU oracle, never linked into the product.
 * Plain-C restatement of the depthwise 7x7 convolution that opens every ConvNeXt block of the
 * reference's CLIP-ConvNeXt tower: /root/reference/ola_vlm/model/multimodal_encoder/
 * clip_convnext_encoder.py:159-162 runs `stage(x)` of timm's ConvNeXt (timm==1.0.8, setup.py:20, not
 * vendored): `conv_dw = nn.Conv2d(C, C, kernel_size=7, padding=3, groups=C)`, i.e. torch conv2d
 * cross-correlation, zero padding.  NHWC float in/out; the caller rounds to bf16 where it wants to
 * model storage.  Pinned against torch.nn.functional.conv2d in tests/test_convnext_cpu.py. */
#include <stdint.h>

void oracle_dwconv7x7_nhwc(const float* in, const float* w /* [C][7][7] torch layout */,
                           const float* bias, float* out, int B, int H, int W, int C) {
  for (int b = 0; b < B; ++b)
    for (int y = 0; y < H; ++y)
      for (int x = 0; x < W; ++x)
        for (int c = 0; c < C; ++c) {
          double acc = bias ? bias[c] : 0.0;
          for (int ky = 0; ky < 7; ++ky) {
            const int iy = y + ky - 3;
            if (iy < 0 || iy >= H) continue;
            for (int kx = 0; kx < 7; ++kx) {
              const int ix = x + kx - 3;
              if (ix < 0 || ix >= W) continue;
              acc += (double)in[(((int64_t)b * H + iy) * W + ix) * C + c] * (double)w[(c * 7 + ky) * 7 + kx];
            }
          }
          out[(((int64_t)b * H + y) * W + x) * C + c] = (float)acc;
        }
}
