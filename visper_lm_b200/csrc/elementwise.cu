// HBM-bound elementwise / gather kernels: RoPE (in place on the packed QKV buffer), SwiGLU,
// activation backward, im2col for the CLIP patch embedding, CLIP embedding assembly, row
// gather / gather-sum / scatter-add for the multimodal splice, group mean for task tokens,
// transpose and small utilities.  All use 128-bit coalesced accesses.
//
// Reference call sites: HF apply_rotary_pos_emb / LlamaMLP (silu(gate)*up) / CLIPVisionEmbeddings
// behind /root/reference/ola_vlm/model/language_model/ola_llama.py:105 and
// multimodal_encoder/clip_encoder.py:56; the splice of
// /root/reference/ola_vlm/model/ola_arch.py:345-444 (+ append_special_tokens :224-254).
#include "common.cuh"
#include "visper_b200.h"

namespace vpb {

// ---------------------------------------------------------------------------------------------
// RoPE
// ---------------------------------------------------------------------------------------------
__global__ void rope_table_kernel(float* __restrict__ cs, float* __restrict__ sn, int max_pos,
                                  int half, float theta) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= max_pos * half) return;
  const int pos = idx / half, i = idx - pos * half;
  // HF: inv_freq = 1 / base ** (arange(0, dim, 2) / dim) in fp32; freqs = pos * inv_freq
  const float inv_freq = 1.0f / powf(theta, (float)(2 * i) / (float)(2 * half));
  const float ang = (float)pos * inv_freq;
  cs[idx] = cosf(ang);
  sn[idx] = sinf(ang);
}

// x: [M, ld] packed heads; rotates `nheads` heads of width hd starting at column col0.
// out[i] = x[i]*c - x[i+h]*s ; out[i+h] = x[i+h]*c + x[i]*s   (sign = -1 → inverse rotation)
__global__ void __launch_bounds__(128)
rope_kernel(bf16* __restrict__ x, int64_t ld, const int* __restrict__ pos_ids, int seq_len,
            const float* __restrict__ cs, const float* __restrict__ sn, int nheads, int hd,
            float sign) {
  const int m = blockIdx.x;
  const int half = hd >> 1;
  const int vph = half >> 3;  // 8-wide vectors per half head
  const int pos = pos_ids ? pos_ids[m] : (m % seq_len);
  const float* c = cs + (int64_t)pos * half;
  const float* s = sn + (int64_t)pos * half;
  bf16* row = x + (int64_t)m * ld;
  for (int t = threadIdx.x; t < nheads * vph; t += blockDim.x) {
    const int h = t / vph, v = t - h * vph;
    bf16* p1 = row + h * hd + v * 8;
    bf16* p2 = p1 + half;
    float a[8], b[8], o1[8], o2[8];
    unpack8(*reinterpret_cast<const uint4*>(p1), a);
    unpack8(*reinterpret_cast<const uint4*>(p2), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      // explicit rounding points: the fused QKV-GEMM epilogue (gemm_tcgen05.cu, EPI_ROPE) uses the
      // same expressions and must give the same bits
      const float cc = c[v * 8 + j], ss = sign * s[v * 8 + j];
      o1[j] = __fmaf_rn(a[j], cc, -__fmul_rn(b[j], ss));
      o2[j] = __fmaf_rn(b[j], cc, __fmul_rn(a[j], ss));
    }
    stg16(p1, pack8(o1));
    stg16(p2, pack8(o2));
  }
}

// ---------------------------------------------------------------------------------------------
// SwiGLU and activation backward
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
swiglu_fwd_kernel(const bf16* __restrict__ gu, int64_t ldgu, bf16* __restrict__ h, int64_t ldh,
                  int M, int F) {
  const int vpr = F >> 3;
  const int64_t total = (int64_t)M * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / vpr), v = (int)(i - (int64_t)m * vpr);
    float g[8], u[8], o[8];
    unpack8(ldg16_stream(gu + (int64_t)m * ldgu + v * 8), g);
    unpack8(ldg16_stream(gu + (int64_t)m * ldgu + F + v * 8), u);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = g[j] / (1.f + __expf(-g[j])) * u[j];
    stg16(h + (int64_t)m * ldh + v * 8, pack8(o));
  }
}

__global__ void __launch_bounds__(256)
swiglu_bwd_kernel(const bf16* __restrict__ gu, int64_t ldgu, const bf16* __restrict__ dh,
                  int64_t lddh, bf16* __restrict__ dgu, int64_t lddgu, int M, int F) {
  const int vpr = F >> 3;
  const int64_t total = (int64_t)M * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / vpr), v = (int)(i - (int64_t)m * vpr);
    float g[8], u[8], d[8], dg[8], du[8];
    unpack8(ldg16_stream(gu + (int64_t)m * ldgu + v * 8), g);
    unpack8(ldg16_stream(gu + (int64_t)m * ldgu + F + v * 8), u);
    unpack8(ldg16_stream(dh + (int64_t)m * lddh + v * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float sg = 1.f / (1.f + __expf(-g[j]));
      const float silu = g[j] * sg;
      du[j] = d[j] * silu;
      dg[j] = d[j] * u[j] * (sg + silu * (1.f - sg));
    }
    stg16(dgu + (int64_t)m * lddgu + v * 8, pack8(dg));
    stg16(dgu + (int64_t)m * lddgu + F + v * 8, pack8(du));
  }
}

__device__ __forceinline__ float act_grad(float x, int act) {
  switch (act) {
    case VPB_ACT_GELU: {
      const float cdf = 0.5f * (1.f + erff(x * 0.70710678118654752f));
      const float pdf = 0.3989422804014327f * __expf(-0.5f * x * x);
      return cdf + x * pdf;
    }
    case VPB_ACT_QUICK_GELU: {
      const float s = 1.f / (1.f + __expf(-1.702f * x));
      return s + 1.702f * x * s * (1.f - s);
    }
    case VPB_ACT_RELU:
      return x > 0.f ? 1.f : 0.f;
    default:
      return 1.f;
  }
}

// dx = dy * act'(pre)
__global__ void __launch_bounds__(256)
act_bwd_kernel(const bf16* __restrict__ pre, int64_t ldp, const bf16* __restrict__ dy, int64_t lddy,
               bf16* __restrict__ dx, int64_t lddx, int M, int N, int act) {
  const int vpr = N >> 3;
  const int64_t total = (int64_t)M * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int m = (int)(i / vpr), v = (int)(i - (int64_t)m * vpr);
    float p[8], d[8], o[8];
    unpack8(ldg16_stream(pre + (int64_t)m * ldp + v * 8), p);
    unpack8(ldg16_stream(dy + (int64_t)m * lddy + v * 8), d);
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = d[j] * act_grad(p[j], act);
    stg16(dx + (int64_t)m * lddx + v * 8, pack8(o));
  }
}

// out = alpha*a (+ beta*b), bf16, flat 8-wide
__global__ void __launch_bounds__(256)
axpby_kernel(const bf16* __restrict__ a, const bf16* __restrict__ b, bf16* __restrict__ out,
             float alpha, float beta, int64_t nvec) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    float x[8], o[8];
    unpack8(ldg16_stream(a + i * 8), x);
    if (b) {
      float y[8];
      unpack8(ldg16_stream(b + i * 8), y);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = alpha * x[j] + beta * y[j];
    } else {
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = alpha * x[j];
    }
    stg16(out + i * 8, pack8(o));
  }
}

// out = in * (*scale) with the scalar read from device memory (autograd grad_output, no host sync)
__global__ void __launch_bounds__(256)
scale_dev_kernel(const bf16* __restrict__ in, bf16* __restrict__ out, const float* __restrict__ sc,
                 int64_t nvec) {
  const float s = *sc;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec;
       i += (int64_t)gridDim.x * blockDim.x) {
    float x[8];
    unpack8(ldg16_stream(in + i * 8), x);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] *= s;
    stg16(out + i * 8, pack8(x));
  }
}

// ---------------------------------------------------------------------------------------------
// CLIP patch embedding helpers
// ---------------------------------------------------------------------------------------------
// images [B,3,H,W] bf16 → patches [B*gh*gw, Kpad]; column = c*P*P + ky*P + kx (conv weight order)
__global__ void __launch_bounds__(256)
im2col_kernel(const bf16* __restrict__ img, bf16* __restrict__ out, int B, int H, int W, int P,
              int Kpad) {
  const int gh = H / P, gw = W / P;
  const int K = 3 * P * P;
  const int vpr = Kpad >> 3;
  const int64_t total = (int64_t)B * gh * gw * vpr;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i % vpr);
    const int64_t r = i / vpr;
    const int px = (int)(r % gw);
    const int py = (int)((r / gw) % gh);
    const int b = (int)(r / ((int64_t)gw * gh));
    float o[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int col = v * 8 + j;
      float val = 0.f;
      if (col < K) {
        const int c = col / (P * P);
        const int rem = col - c * P * P;
        const int ky = rem / P, kx = rem - ky * P;
        val = __bfloat162float(
            img[(((int64_t)b * 3 + c) * H + (py * P + ky)) * W + (px * P + kx)]);
      }
      o[j] = val;
    }
    stg16(out + r * Kpad + v * 8, pack8(o));
  }
}

// hidden[b, 0] = cls + pos[0]; hidden[b, 1+p] = patch[b, p] + pos[1+p]
__global__ void __launch_bounds__(128)
clip_embed_kernel(const bf16* __restrict__ patch, const bf16* __restrict__ cls,
                  const bf16* __restrict__ pos, bf16* __restrict__ out, int npatch, int D) {
  const int r = blockIdx.x;  // over B*(npatch+1)
  const int S = npatch + 1;
  const int b = r / S, t = r - b * S;
  const bf16* src = t == 0 ? cls : patch + ((int64_t)b * npatch + (t - 1)) * D;
  for (int v = threadIdx.x; v < (D >> 3); v += blockDim.x) {
    float a[8], p[8];
    unpack8(ldg16(src + v * 8), a);
    unpack8(ldg16(pos + (int64_t)t * D + v * 8), p);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += p[j];
    stg16(out + (int64_t)r * D + v * 8, pack8(a));
  }
}

// ---------------------------------------------------------------------------------------------
// row gather / gather-sum / scatter-add (the multimodal splice and its backward)
// ---------------------------------------------------------------------------------------------
struct GatherSrc {
  const bf16* p[4];
  int64_t ld[4];
};
// out[r] = src[kind[r]][index[r]]  (kind < 0 or index < 0 → zeros)
__global__ void __launch_bounds__(128)
gather_rows_kernel(bf16* __restrict__ out, int64_t ldo, const int* __restrict__ kind,
                   const int* __restrict__ index, GatherSrc s, int D) {
  const int r = blockIdx.x;
  const int k = kind ? kind[r] : 0;
  const int idx = index[r];
  const bool valid = k >= 0 && idx >= 0;
  const bf16* src = valid ? s.p[k] + (int64_t)idx * s.ld[k] : nullptr;
  for (int v = threadIdx.x; v < (D >> 3); v += blockDim.x) {
    uint4 val = make_uint4(0, 0, 0, 0);
    if (valid) val = ldg16_stream(src + v * 8);
    stg16(out + (int64_t)r * ldo + v * 8, val);
  }
}

// EXPERIMENTAL (VPB_OPT_GATHER_FLAT, off until measured): the same gather as a flat grid-stride loop over
// (row, 16-byte vector) pairs.  One CTA per row leaves 104 of 128 threads idle on a 192-wide Swin row and
// launches 333 k CTAs per call (seg teacher gathers: 3.4 ms per batch against ~0.8 ms of HBM time).
__global__ void __launch_bounds__(256)
gather_rows_flat_kernel(bf16* __restrict__ out, int64_t ldo, const int* __restrict__ kind,
                        const int* __restrict__ index, GatherSrc s, int D, int nrows) {
  const int vec = D >> 3;
  const int64_t total = (int64_t)nrows * vec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / vec), v = (int)(i - (int64_t)r * vec);
    const int k = kind ? kind[r] : 0;
    const int idx = index[r];
    uint4 val = make_uint4(0, 0, 0, 0);
    if (k >= 0 && idx >= 0) val = ldg16_stream(s.p[k] + (int64_t)idx * s.ld[k] + v * 8);
    stg16(out + (int64_t)r * ldo + v * 8, val);
  }
}

// out[s] = scale * sum_{c<cnt} src[index[s*cnt+c]]   (index < 0 skipped), fp32 accumulate
__global__ void __launch_bounds__(128)
gather_sum_rows_kernel(bf16* __restrict__ out, int64_t ldo, const int* __restrict__ index, int cnt,
                       const bf16* __restrict__ src, int64_t lds, int D, float scale) {
  const int s = blockIdx.x;
  for (int v = threadIdx.x; v < (D >> 3); v += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int c = 0; c < cnt; ++c) {
      const int idx = index[s * cnt + c];
      if (idx >= 0) {
        float x[8];
        unpack8(ldg16_stream(src + (int64_t)idx * lds + v * 8), x);
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += x[j];
      }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= scale;
    stg16(out + (int64_t)s * ldo + v * 8, pack8(acc));
  }
}

// dst_f32[index[r]] += src[r]   (embedding-table gradient; index < 0 skipped)
__global__ void __launch_bounds__(128)
scatter_add_rows_kernel(float* __restrict__ dst, int64_t ldd, const int* __restrict__ index,
                        const bf16* __restrict__ src, int64_t lds, int D) {
  const int r = blockIdx.x;
  const int idx = index[r];
  if (idx < 0) return;
  for (int v = threadIdx.x; v < (D >> 3); v += blockDim.x) {
    float x[8];
    unpack8(ldg16_stream(src + (int64_t)r * lds + v * 8), x);
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(dst + (int64_t)idx * ldd + v * 8 + j, x[j]);
  }
}

// dst[index[r]] += src[r]  (bf16, indices unique within one call; index < 0 skipped)
__global__ void __launch_bounds__(128)
add_rows_kernel(bf16* __restrict__ dst, int64_t ldd, const int* __restrict__ index,
                const bf16* __restrict__ src, int64_t lds, int D) {
  const int r = blockIdx.x;
  const int idx = index[r];
  if (idx < 0) return;
  for (int v = threadIdx.x; v < (D >> 3); v += blockDim.x) {
    float a[8], b[8];
    unpack8(*reinterpret_cast<const uint4*>(dst + (int64_t)idx * ldd + v * 8), a);
    unpack8(ldg16_stream(src + (int64_t)r * lds + v * 8), b);
#pragma unroll
    for (int j = 0; j < 8; ++j) a[j] += b[j];
    stg16(dst + (int64_t)idx * ldd + v * 8, pack8(a));
  }
}

// in [groups*gsize, D] → out [groups, D] mean over consecutive gsize rows (task-token pooling,
// ola_arch.py:225-228) ; backward broadcasts dout/gsize.
__global__ void __launch_bounds__(128)
group_mean_kernel(const bf16* __restrict__ in, int64_t ldi, bf16* __restrict__ out, int64_t ldo,
                  int gsize, int D) {
  const int g = blockIdx.x;
  for (int v = threadIdx.x; v < (D >> 3); v += blockDim.x) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int r = 0; r < gsize; ++r) {
      float x[8];
      unpack8(ldg16(in + ((int64_t)g * gsize + r) * ldi + v * 8), x);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += x[j];
    }
    const float inv = 1.f / gsize;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] *= inv;
    stg16(out + (int64_t)g * ldo + v * 8, pack8(acc));
  }
}
__global__ void __launch_bounds__(128)
group_mean_bwd_kernel(const bf16* __restrict__ dout, int64_t ldo, bf16* __restrict__ din,
                      int64_t ldi, int gsize, int D) {
  const int r = blockIdx.x;
  const int g = r / gsize;
  const float inv = 1.f / gsize;
  for (int v = threadIdx.x; v < (D >> 3); v += blockDim.x) {
    float x[8];
    unpack8(ldg16(dout + (int64_t)g * ldo + v * 8), x);
#pragma unroll
    for (int j = 0; j < 8; ++j) x[j] *= inv;
    stg16(din + (int64_t)r * ldi + v * 8, pack8(x));
  }
}

// out[c, r] = in[r, c]  (32x32 tiles through shared memory)
__global__ void __launch_bounds__(256)
transpose_kernel(const bf16* __restrict__ in, int64_t ldi, bf16* __restrict__ out, int64_t ldo,
                 int R, int C) {
  __shared__ bf16 tile[32][34];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  for (int i = ty; i < 32; i += 8) {
    const int r = r0 + i, c = c0 + tx;
    tile[i][tx] = (r < R && c < C) ? in[(int64_t)r * ldi + c] : __float2bfloat16(0.f);
  }
  __syncthreads();
  for (int i = ty; i < 32; i += 8) {
    const int c = c0 + i, r = r0 + tx;
    if (r < R && c < C) out[(int64_t)c * ldo + r] = tile[tx][i];
  }
}

__global__ void __launch_bounds__(256)
cast_f32_bf16_kernel(const float* __restrict__ in, bf16* __restrict__ out, int64_t n, float scale) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float2bfloat16(in[i] * scale);
}

static inline int grid_for(int64_t work, int block) {
  int64_t g = (work + block - 1) / block;
  const int64_t cap = (int64_t)148 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace vpb

using namespace vpb;
#define ST(s) ((cudaStream_t)(s))

extern "C" int vpb_rope_table(float* cos_t, float* sin_t, int max_pos, int head_dim, float theta,
                              void* stream) {
  VPB_CHECK(head_dim % 16 == 0 && max_pos > 0, "rope_table: head_dim=%d max_pos=%d", head_dim, max_pos);
  const int half = head_dim / 2;
  const int n = max_pos * half;
  rope_table_kernel<<<(n + 255) / 256, 256, 0, ST(stream)>>>(cos_t, sin_t, max_pos, half, theta);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_rope_inplace(void* x, int64_t ld, int M, int seq_len, const int* pos_ids,
                                const float* cos_t, const float* sin_t, int nheads, int head_dim,
                                int inverse, void* stream) {
  VPB_CHECK(head_dim % 16 == 0 && ld % 8 == 0 && M > 0 && seq_len > 0, "rope: bad shape");
  rope_kernel<<<M, 128, 0, ST(stream)>>>((bf16*)x, ld, pos_ids, seq_len, cos_t, sin_t, nheads,
                                         head_dim, inverse ? -1.f : 1.f);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_swiglu_fwd(const void* gu, int64_t ldgu, void* h, int64_t ldh, int M, int F,
                              void* stream) {
  VPB_CHECK(F % 8 == 0 && ldgu % 8 == 0 && ldh % 8 == 0 && M > 0, "swiglu: bad shape");
  swiglu_fwd_kernel<<<grid_for((int64_t)M * (F / 8), 256), 256, 0, ST(stream)>>>(
      (const bf16*)gu, ldgu, (bf16*)h, ldh, M, F);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_swiglu_bwd(const void* gu, int64_t ldgu, const void* dh, int64_t lddh, void* dgu,
                              int64_t lddgu, int M, int F, void* stream) {
  VPB_CHECK(F % 8 == 0 && ldgu % 8 == 0 && lddh % 8 == 0 && lddgu % 8 == 0 && M > 0,
            "swiglu_bwd: bad shape");
  swiglu_bwd_kernel<<<grid_for((int64_t)M * (F / 8), 256), 256, 0, ST(stream)>>>(
      (const bf16*)gu, ldgu, (const bf16*)dh, lddh, (bf16*)dgu, lddgu, M, F);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_act_bwd(const void* pre, int64_t ldp, const void* dy, int64_t lddy, void* dx,
                           int64_t lddx, int M, int N, int act, void* stream) {
  VPB_CHECK(N % 8 == 0 && ldp % 8 == 0 && lddy % 8 == 0 && lddx % 8 == 0 && M > 0, "act_bwd: bad shape");
  act_bwd_kernel<<<grid_for((int64_t)M * (N / 8), 256), 256, 0, ST(stream)>>>(
      (const bf16*)pre, ldp, (const bf16*)dy, lddy, (bf16*)dx, lddx, M, N, act);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_axpby(const void* a, const void* b, void* out, float alpha, float beta,
                         int64_t n, void* stream) {
  VPB_CHECK(n % 8 == 0 && n > 0, "axpby: n=%lld must be a positive multiple of 8", (long long)n);
  axpby_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>((const bf16*)a, (const bf16*)b,
                                                             (bf16*)out, alpha, beta, n / 8);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_scale_dev(const void* in, void* out, const float* scale, int64_t n, void* stream) {
  VPB_CHECK(n % 8 == 0 && n > 0, "scale_dev: n=%lld must be a positive multiple of 8", (long long)n);
  scale_dev_kernel<<<grid_for(n / 8, 256), 256, 0, ST(stream)>>>((const bf16*)in, (bf16*)out, scale, n / 8);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_im2col_patches(const void* images, void* out, int B, int H, int W, int patch,
                                  int Kpad, void* stream) {
  VPB_CHECK(H % patch == 0 && W % patch == 0 && Kpad % 8 == 0 && Kpad >= 3 * patch * patch,
            "im2col: bad shape");
  const int64_t total = (int64_t)B * (H / patch) * (W / patch) * (Kpad / 8);
  im2col_kernel<<<grid_for(total, 256), 256, 0, ST(stream)>>>((const bf16*)images, (bf16*)out, B, H,
                                                              W, patch, Kpad);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_clip_embed(const void* patch, const void* cls, const void* pos, void* out, int B,
                              int npatch, int D, void* stream) {
  VPB_CHECK(D % 8 == 0 && B > 0, "clip_embed: bad shape");
  clip_embed_kernel<<<B * (npatch + 1), 128, 0, ST(stream)>>>((const bf16*)patch, (const bf16*)cls,
                                                              (const bf16*)pos, (bf16*)out, npatch, D);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_gather_rows(void* out, int64_t ldo, int nrows, int D, const int* kind,
                               const int* index, const void* src0, int64_t ld0, const void* src1,
                               int64_t ld1, const void* src2, int64_t ld2, const void* src3,
                               int64_t ld3, void* stream) {
  VPB_CHECK(D % 8 == 0 && nrows > 0 && ldo % 8 == 0, "gather_rows: bad shape");
  GatherSrc s;
  s.p[0] = (const bf16*)src0; s.ld[0] = ld0;
  s.p[1] = (const bf16*)src1; s.ld[1] = ld1;
  s.p[2] = (const bf16*)src2; s.ld[2] = ld2;
  s.p[3] = (const bf16*)src3; s.ld[3] = ld3;
  if (get_option(VPB_OPT_GATHER_FLAT) && D <= 2048) {
    const int64_t total = (int64_t)nrows * (D >> 3);
    const int64_t want = (total + 255) / 256, cap = (int64_t)num_sms() * 16;
    gather_rows_flat_kernel<<<(int)(want < cap ? want : cap), 256, 0, ST(stream)>>>((bf16*)out, ldo, kind, index, s,
                                                                                   D, nrows);
    VPB_LAUNCH_OK();
    return 0;
  }
  gather_rows_kernel<<<nrows, 128, 0, ST(stream)>>>((bf16*)out, ldo, kind, index, s, D);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_gather_sum_rows(void* out, int64_t ldo, int nslots, int cnt, const int* index,
                                   const void* src, int64_t lds, int D, float scale, void* stream) {
  VPB_CHECK(D % 8 == 0 && nslots > 0 && cnt > 0, "gather_sum_rows: bad shape");
  gather_sum_rows_kernel<<<nslots, 128, 0, ST(stream)>>>((bf16*)out, ldo, index, cnt,
                                                         (const bf16*)src, lds, D, scale);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_scatter_add_rows(float* dst, int64_t ldd, int nrows, const int* index,
                                    const void* src, int64_t lds, int D, void* stream) {
  VPB_CHECK(D % 8 == 0 && nrows > 0, "scatter_add_rows: bad shape");
  scatter_add_rows_kernel<<<nrows, 128, 0, ST(stream)>>>(dst, ldd, index, (const bf16*)src, lds, D);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_add_rows(void* dst, int64_t ldd, int nrows, const int* index, const void* src,
                            int64_t lds, int D, void* stream) {
  VPB_CHECK(D % 8 == 0 && nrows > 0, "add_rows: bad shape");
  add_rows_kernel<<<nrows, 128, 0, ST(stream)>>>((bf16*)dst, ldd, index, (const bf16*)src, lds, D);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_group_mean(const void* in, int64_t ldi, void* out, int64_t ldo, int groups,
                              int gsize, int D, void* stream) {
  VPB_CHECK(D % 8 == 0 && groups > 0 && gsize > 0, "group_mean: bad shape");
  group_mean_kernel<<<groups, 128, 0, ST(stream)>>>((const bf16*)in, ldi, (bf16*)out, ldo, gsize, D);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_group_mean_bwd(const void* dout, int64_t ldo, void* din, int64_t ldi, int groups,
                                  int gsize, int D, void* stream) {
  VPB_CHECK(D % 8 == 0 && groups > 0 && gsize > 0, "group_mean_bwd: bad shape");
  group_mean_bwd_kernel<<<groups * gsize, 128, 0, ST(stream)>>>((const bf16*)dout, ldo, (bf16*)din,
                                                                ldi, gsize, D);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_transpose(const void* in, int64_t ldi, void* out, int64_t ldo, int R, int C,
                             void* stream) {
  VPB_CHECK(R > 0 && C > 0, "transpose: bad shape");
  dim3 grid((C + 31) / 32, (R + 31) / 32);
  transpose_kernel<<<grid, 256, 0, ST(stream)>>>((const bf16*)in, ldi, (bf16*)out, ldo, R, C);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_cast_f32_bf16(const float* in, void* out, int64_t n, float scale, void* stream) {
  VPB_CHECK(n > 0, "cast: n=%lld", (long long)n);
  cast_f32_bf16_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>(in, (bf16*)out, n, scale);
  VPB_LAUNCH_OK();
  return 0;
}
