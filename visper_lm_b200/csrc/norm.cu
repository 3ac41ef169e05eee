// Row normalisation kernels (HBM-bound): RMSNorm / LayerNorm forward + backward and the column
// reductions that produce their weight / bias gradients.  One CTA per row, 128-bit coalesced
// loads, row cached in registers, fp32 statistics with warp-shuffle reductions.
//
// Reference semantics: HF LlamaRMSNorm / Phi3RMSNorm (transformers modeling_llama.py: x.float(),
// x * rsqrt(mean(x²)+eps), cast to input dtype, then * weight) reached from
// /root/reference/ola_vlm/model/language_model/ola_llama.py:105; torch.nn.LayerNorm in
// CLIPEncoderLayer (clip_encoder.py:56) and PerceiverAttention/FeedForward/norm_out
// (ola_vlm/model/multimodal_projector/resampler.py:9-16,39-40,186).
#include "common.cuh"
#include "visper_b200.h"

namespace vpb {

constexpr int NORM_THREADS = 256;
constexpr int NORM_MAXV = 4;  // vectors of 8 per thread → D <= 8192

template <bool RMS>
__global__ void __launch_bounds__(NORM_THREADS)
norm_fwd_kernel(const bf16* __restrict__ x, int64_t ldx, const bf16* __restrict__ w,
                const bf16* __restrict__ b, bf16* __restrict__ y, int64_t ldy,
                float* __restrict__ mean_out, float* __restrict__ rstd_out, int D, float eps) {
  __shared__ float red[33];
  const int row = blockIdx.x;
  const int nvec = D >> 3;
  const bf16* xr = x + (int64_t)row * ldx;
  float v[NORM_MAXV][8];
  float s = 0.f, ss = 0.f;
#pragma unroll
  for (int k = 0; k < NORM_MAXV; ++k) {
    const int i = threadIdx.x + k * NORM_THREADS;
    if (i < nvec) {
      unpack8(ldg16_stream(xr + i * 8), v[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        s += v[k][j];
        ss += v[k][j] * v[k][j];
      }
    }
  }
  float mean = 0.f, rstd;
  if (RMS) {
    ss = block_sum(ss, red);
    rstd = rsqrtf(ss / D + eps);
  } else {
    s = block_sum(s, red);
    mean = s / D;
    // two-pass variance from the register cache (matches torch's numerics better than E[x²]-m²)
    float vs = 0.f;
#pragma unroll
    for (int k = 0; k < NORM_MAXV; ++k) {
      const int i = threadIdx.x + k * NORM_THREADS;
      if (i < nvec) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float d = v[k][j] - mean;
          vs += d * d;
        }
      }
    }
    vs = block_sum(vs, red);
    rstd = rsqrtf(vs / D + eps);
  }
  if (threadIdx.x == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  bf16* yr = y + (int64_t)row * ldy;
#pragma unroll
  for (int k = 0; k < NORM_MAXV; ++k) {
    const int i = threadIdx.x + k * NORM_THREADS;
    if (i < nvec) {
      float wv[8], o[8];
      unpack8(ldg16(w + i * 8), wv);
      if (RMS) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          o[j] = wv[j] * __bfloat162float(__float2bfloat16(v[k][j] * rstd));  // HF rounding order
      } else {
        float bv[8];
        unpack8(ldg16(b + i * 8), bv);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = (v[k][j] - mean) * rstd * wv[j] + bv[j];
      }
      stg16(yr + i * 8, pack8(o));
    }
  }
}

// LayerNorm forward for MANY SHORT rows (the Swin teacher: up to 330k rows of 192 / 384 / 768): one warp
// per row, eight rows per CTA, shuffle-only reductions — the CTA-per-row kernel above leaves 7/8 of its
// threads idle and pays three block reductions per 384-byte row there.  Same two-pass arithmetic.
constexpr int WNORM_MAXV = 3;      // 32 lanes x 3 vectors of 8 → D <= 768
constexpr int WNORM_MIN_ROWS = 16384;
__global__ void __launch_bounds__(256)
layernorm_fwd_warp_kernel(const bf16* __restrict__ x, int64_t ldx, const bf16* __restrict__ w,
                          const bf16* __restrict__ b, bf16* __restrict__ y, int64_t ldy,
                          float* __restrict__ mean_out, float* __restrict__ rstd_out, int M, int D,
                          float eps) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= M) return;
  const int nvec = D >> 3;
  const bf16* xr = x + (int64_t)row * ldx;
  float v[WNORM_MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < WNORM_MAXV; ++k) {
    const int i = lane + k * 32;
    if (i < nvec) {
      unpack8(ldg16_stream(xr + i * 8), v[k]);
#pragma unroll
      for (int j = 0; j < 8; ++j) s += v[k][j];
    }
  }
  const float mean = warp_sum(s) / D;
  float vs = 0.f;
#pragma unroll
  for (int k = 0; k < WNORM_MAXV; ++k) {
    const int i = lane + k * 32;
    if (i < nvec) {
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        const float d = v[k][j] - mean;
        vs += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(vs) / D + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  bf16* yr = y + (int64_t)row * ldy;
#pragma unroll
  for (int k = 0; k < WNORM_MAXV; ++k) {
    const int i = lane + k * 32;
    if (i < nvec) {
      float wv[8], bv[8], o[8];
      unpack8(ldg16(w + i * 8), wv);
      unpack8(ldg16(b + i * 8), bv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = (v[k][j] - mean) * rstd * wv[j] + bv[j];
      stg16(yr + i * 8, pack8(o));
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// RMSNorm of the decoder rows, high-occupancy variants of the CTA-per-row kernels.
// ncu on norm_fwd_kernel<true> at M = 16384, D = 4096 (profiles/r02_ncu_hbm_kernels.csv): 58 registers → 4 CTAs =
// 4 rows = 32 KB of loads in flight per SM, DRAM 44 % busy, 0.63 of the HBM copy rate where a plain copy of the
// same 268 MB reaches 0.83: the kernel is bound by the memory latency it can cover, not by bandwidth.  These
// variants keep the row PACKED (bf16, 16 B per vector) in registers across the block reduction and unpack it twice
// instead, which fits 32 (forward) / 40 (backward) registers: 8 / 6 resident CTAs per SM, twice the bytes in flight.
// (Two other designs were measured and dropped — one warp per row with the row in registers: 243 registers, and a
// persistent cp.async ring of rows in shared memory: no faster, ALU-latency-bound with 8 warps per SM —
// profiles/r02_norm_experiments.txt.)
// ---------------------------------------------------------------------------------------------------
template <int MAXV>
__global__ void __launch_bounds__(NORM_THREADS, 8)
rmsnorm_fwd_packed_kernel(const bf16* __restrict__ x, int64_t ldx, const bf16* __restrict__ w,
                          bf16* __restrict__ y, int64_t ldy, float* __restrict__ rstd_out, int D, float eps) {
  __shared__ float red[33];
  const int row = blockIdx.x;
  const int nvec = D >> 3;
  const bf16* xr = x + (int64_t)row * ldx;
  uint4 xv[MAXV];
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int i = threadIdx.x + k * NORM_THREADS;
    xv[k] = i < nvec ? ldg16_stream(xr + i * 8) : make_uint4(0, 0, 0, 0);
  }
  float ss = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    float v[8];
    unpack8(xv[k], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) ss = fmaf(v[j], v[j], ss);
  }
  ss = block_sum(ss, red);
  const float rstd = rsqrtf(ss / D + eps);
  if (threadIdx.x == 0 && rstd_out) rstd_out[row] = rstd;
  bf16* yr = y + (int64_t)row * ldy;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int i = threadIdx.x + k * NORM_THREADS;
    if (i < nvec) {
      asm volatile("" : "+r"(xv[k].x), "+r"(xv[k].y), "+r"(xv[k].z), "+r"(xv[k].w));  // re-unpack, do not keep floats
      float v[8], wv[8], o[8];
      unpack8(xv[k], v);
      unpack8(ldg16(w + i * 8), wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = wv[j] * __bfloat162float(__float2bfloat16(v[j] * rstd));  // HF rounding order
      stg16(yr + i * 8, pack8(o));
    }
  }
}

template <int MAXV>
__global__ void __launch_bounds__(NORM_THREADS, 6)
rmsnorm_bwd_packed_kernel(const bf16* __restrict__ dy, int64_t lddy, const bf16* __restrict__ x, int64_t ldx,
                          const bf16* __restrict__ w, const float* __restrict__ rstd_in,
                          const bf16* __restrict__ dres, int64_t lddres, bf16* __restrict__ dx, int64_t lddx, int D) {
  __shared__ float red[33];
  const int row = blockIdx.x;
  const int nvec = D >> 3;
  const bf16* xr = x + (int64_t)row * ldx;
  const bf16* dyr = dy + (int64_t)row * lddy;
  uint4 xv[MAXV], gv[MAXV];
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int i = threadIdx.x + k * NORM_THREADS;
    xv[k] = i < nvec ? ldg16_stream(xr + i * 8) : make_uint4(0, 0, 0, 0);
    gv[k] = i < nvec ? ldg16_stream(dyr + i * 8) : make_uint4(0, 0, 0, 0);
  }
  const float rstd = rstd_in[row];
  float sgx = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int i = threadIdx.x + k * NORM_THREADS;
    if (i < nvec) {
      float xh[8], g[8], wv[8];
      unpack8(xv[k], xh);
      unpack8(gv[k], g);
      unpack8(ldg16(w + i * 8), wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) sgx = fmaf(g[j] * wv[j], xh[j] * rstd, sgx);
    }
  }
  sgx = block_sum(sgx, red) / D;
  bf16* dxr = dx + (int64_t)row * lddx;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int i = threadIdx.x + k * NORM_THREADS;
    if (i < nvec) {
      asm volatile("" : "+r"(xv[k].x), "+r"(xv[k].y), "+r"(xv[k].z), "+r"(xv[k].w));
      asm volatile("" : "+r"(gv[k].x), "+r"(gv[k].y), "+r"(gv[k].z), "+r"(gv[k].w));
      float xh[8], g[8], wv[8], o[8];
      unpack8(xv[k], xh);
      unpack8(gv[k], g);
      unpack8(ldg16(w + i * 8), wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (g[j] * wv[j] - (xh[j] * rstd) * sgx);
      if (dres) {
        float r[8];
        unpack8(ldg16_stream(dres + (int64_t)row * lddres + i * 8), r);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += r[j];
      }
      stg16(dxr + i * 8, pack8(o));
    }
  }
}

// dx = rstd * (g - [mean(g)] - xhat * mean(g*xhat)),  g = dy * w ;  dx (+)= dres if given
template <bool RMS>
__global__ void __launch_bounds__(NORM_THREADS)
norm_bwd_kernel(const bf16* __restrict__ dy, int64_t lddy, const bf16* __restrict__ x, int64_t ldx,
                const bf16* __restrict__ w, const float* __restrict__ mean_in,
                const float* __restrict__ rstd_in, const bf16* __restrict__ dres, int64_t lddres,
                bf16* __restrict__ dx, int64_t lddx, int D) {
  __shared__ float red[33];
  const int row = blockIdx.x;
  const int nvec = D >> 3;
  const float mean = RMS ? 0.f : mean_in[row];
  const float rstd = rstd_in[row];
  const bf16* xr = x + (int64_t)row * ldx;
  const bf16* dyr = dy + (int64_t)row * lddy;
  float xh[NORM_MAXV][8], g[NORM_MAXV][8];
  float sg = 0.f, sgx = 0.f;
#pragma unroll
  for (int k = 0; k < NORM_MAXV; ++k) {
    const int i = threadIdx.x + k * NORM_THREADS;
    if (i < nvec) {
      float wv[8];
      unpack8(ldg16_stream(xr + i * 8), xh[k]);
      unpack8(ldg16_stream(dyr + i * 8), g[k]);
      unpack8(ldg16(w + i * 8), wv);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        xh[k][j] = (xh[k][j] - mean) * rstd;
        g[k][j] *= wv[j];
        sg += g[k][j];
        sgx += g[k][j] * xh[k][j];
      }
    }
  }
  sgx = block_sum(sgx, red) / D;
  if (!RMS) sg = block_sum(sg, red) / D; else sg = 0.f;
  bf16* dxr = dx + (int64_t)row * lddx;
#pragma unroll
  for (int k = 0; k < NORM_MAXV; ++k) {
    const int i = threadIdx.x + k * NORM_THREADS;
    if (i < nvec) {
      float o[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = rstd * (g[k][j] - sg - xh[k][j] * sgx);
      if (dres) {
        float r[8];
        unpack8(ldg16_stream(dres + (int64_t)row * lddres + i * 8), r);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] += r[j];
      }
      stg16(dxr + i * 8, pack8(o));
    }
  }
}

// out[n] += sum_m a[m,n] * f(m,n),  f = 1 | (b[m,n]-mean[m])*rstd[m]   (fp32 atomics, out zeroed)
// grid (ceil(N/256), row_splits), block 256 (8 warps × 32 lanes, 8 columns per lane).
__global__ void __launch_bounds__(256)
colsum_kernel(const bf16* __restrict__ a, int64_t lda, const bf16* __restrict__ b, int64_t ldb,
              const float* __restrict__ mean, const float* __restrict__ rstd,
              float* __restrict__ out, int M, int N) {
  __shared__ float sm[8][256];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int n0 = blockIdx.x * 256 + lane * 8;
  const int rows_per = (M + gridDim.y - 1) / gridDim.y;
  const int m_begin = blockIdx.y * rows_per;
  const int m_end = min(M, m_begin + rows_per);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  if (n0 + 8 <= N) {
    // two rows per warp in flight (four 16-byte loads): one row at a time reached 0.57 of the HBM rate
    auto row = [&](int m, const uint4& ua, const uint4& ub) {
      float av[8];
      unpack8(ua, av);
      if (b) {
        float bv[8];
        unpack8(ub, bv);
        const float mu = mean ? mean[m] : 0.f;
        const float rs = rstd ? rstd[m] : 1.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += av[j] * (bv[j] - mu) * rs;
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[j] += av[j];
      }
    };
    int m = m_begin + wid;
    for (; m + 8 < m_end; m += 16) {
      const uint4 a0 = ldg16_stream(a + (int64_t)m * lda + n0);
      const uint4 a1 = ldg16_stream(a + (int64_t)(m + 8) * lda + n0);
      uint4 b0 = a0, b1 = a1;
      if (b) {
        b0 = ldg16_stream(b + (int64_t)m * ldb + n0);
        b1 = ldg16_stream(b + (int64_t)(m + 8) * ldb + n0);
      }
      row(m, a0, b0);
      row(m + 8, a1, b1);
    }
    if (m < m_end) {
      const uint4 a0 = ldg16_stream(a + (int64_t)m * lda + n0);
      const uint4 b0 = b ? ldg16_stream(b + (int64_t)m * ldb + n0) : a0;
      row(m, a0, b0);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) sm[wid][lane * 8 + j] = acc[j];
  __syncthreads();
  const int c = threadIdx.x;
  float t = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) t += sm[k][c];
  const int n = blockIdx.x * 256 + c;
  if (n < N) atomicAdd(out + n, t);
}

}  // namespace vpb

using namespace vpb;

static int check_norm(int M, int D) {
  VPB_CHECK(M > 0, "norm: M=%d", M);
  VPB_CHECK(D % 8 == 0 && D >= 8 && D <= 8 * NORM_THREADS * NORM_MAXV, "norm: unsupported D=%d", D);
  return 0;
}

extern "C" int vpb_rmsnorm_fwd(const void* x, int64_t ldx, const void* w, void* y, int64_t ldy,
                               float* rstd, int M, int D, float eps, void* stream) {
  if (check_norm(M, D)) return -1;
  if (!get_option(VPB_OPT_NORM_R1)) {
    if (D <= 8 * NORM_THREADS * 2)
      rmsnorm_fwd_packed_kernel<2><<<M, NORM_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (const bf16*)w,
                                                                                (bf16*)y, ldy, rstd, D, eps);
    else
      rmsnorm_fwd_packed_kernel<4><<<M, NORM_THREADS, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (const bf16*)w,
                                                                                (bf16*)y, ldy, rstd, D, eps);
    VPB_LAUNCH_OK();
    return 0;
  }
  norm_fwd_kernel<true><<<M, NORM_THREADS, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, ldx, (const bf16*)w, nullptr, (bf16*)y, ldy, nullptr, rstd, D, eps);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_rmsnorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx,
                               const void* w, const float* rstd, const void* dres, int64_t lddres,
                               void* dx, int64_t lddx, int M, int D, void* stream) {
  if (check_norm(M, D)) return -1;
  if (!get_option(VPB_OPT_NORM_R1)) {
    if (D <= 8 * NORM_THREADS * 2)
      rmsnorm_bwd_packed_kernel<2><<<M, NORM_THREADS, 0, (cudaStream_t)stream>>>(
          (const bf16*)dy, lddy, (const bf16*)x, ldx, (const bf16*)w, rstd, (const bf16*)dres, lddres, (bf16*)dx, lddx, D);
    else
      rmsnorm_bwd_packed_kernel<4><<<M, NORM_THREADS, 0, (cudaStream_t)stream>>>(
          (const bf16*)dy, lddy, (const bf16*)x, ldx, (const bf16*)w, rstd, (const bf16*)dres, lddres, (bf16*)dx, lddx, D);
    VPB_LAUNCH_OK();
    return 0;
  }
  norm_bwd_kernel<true><<<M, NORM_THREADS, 0, (cudaStream_t)stream>>>(
      (const bf16*)dy, lddy, (const bf16*)x, ldx, (const bf16*)w, nullptr, rstd, (const bf16*)dres,
      lddres, (bf16*)dx, lddx, D);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_layernorm_fwd(const void* x, int64_t ldx, const void* w, const void* b, void* y,
                                 int64_t ldy, float* mean, float* rstd, int M, int D, float eps,
                                 void* stream) {
  if (check_norm(M, D)) return -1;
  if (D <= 8 * 32 * WNORM_MAXV && M >= WNORM_MIN_ROWS) {
    layernorm_fwd_warp_kernel<<<(M + 7) / 8, 256, 0, (cudaStream_t)stream>>>(
        (const bf16*)x, ldx, (const bf16*)w, (const bf16*)b, (bf16*)y, ldy, mean, rstd, M, D, eps);
    VPB_LAUNCH_OK();
    return 0;
  }
  norm_fwd_kernel<false><<<M, NORM_THREADS, 0, (cudaStream_t)stream>>>(
      (const bf16*)x, ldx, (const bf16*)w, (const bf16*)b, (bf16*)y, ldy, mean, rstd, D, eps);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_layernorm_bwd(const void* dy, int64_t lddy, const void* x, int64_t ldx,
                                 const void* w, const float* mean, const float* rstd,
                                 const void* dres, int64_t lddres, void* dx, int64_t lddx, int M,
                                 int D, void* stream) {
  if (check_norm(M, D)) return -1;
  norm_bwd_kernel<false><<<M, NORM_THREADS, 0, (cudaStream_t)stream>>>(
      (const bf16*)dy, lddy, (const bf16*)x, ldx, (const bf16*)w, mean, rstd, (const bf16*)dres,
      lddres, (bf16*)dx, lddx, D);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_colsum(const void* a, int64_t lda, const void* b, int64_t ldb, const float* mean,
                          const float* rstd, float* out, int M, int N, void* stream) {
  VPB_CHECK(M > 0 && N > 0 && N % 8 == 0, "colsum: M=%d N=%d", M, N);
  VPB_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * N, (cudaStream_t)stream));
  int splits = (M + 63) / 64;
  if (splits > 128) splits = 128;
  dim3 grid((N + 255) / 256, splits);
  colsum_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)a, lda, (const bf16*)b, ldb,
                                                        mean, rstd, out, M, N);
  VPB_LAUNCH_OK();
  return 0;
}
