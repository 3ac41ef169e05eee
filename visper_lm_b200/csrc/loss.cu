// Loss kernels (HBM-bound, fp32 math on bf16 inputs).
//
//  * next-token cross-entropy: one CTA per logits row, online-softmax (running max / sum) in a
//    single read, then the gradient (softmax - onehot)/count written in place over the logits.
//    Reference: /root/reference/ola_vlm/model/language_model/ola_llama.py:121-136
//    (lm_head → .float() → shift → CrossEntropyLoss(mean over labels != -100)).
//  * embedding-distillation loss: smooth-L1 + InfoNCE over flattened, L2-normalised embeddings
//    with (all-gathered) targets.  One launch reads every pred/target element once per tile and
//    produces all B×B' dot products, the squared norms and the smooth-L1 sums; a 1-CTA finalize
//    does the scaled cross-entropy.  Reference: base_ola_vlm.py:289-320 (_emb_loss) and
//    ola_utils.py:108-125 (calculate_contrastive_loss).
#include "common.cuh"
#include "visper_b200.h"

namespace vpb {

// ---------------------------------------------------------------------------------------------
// NTP cross-entropy
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ long long shifted_label(const long long* labels, int64_t grow, int T,
                                                   int shift) {
  if (!shift) return labels[grow];
  const int t = (int)(grow % T);
  return (t + 1 < T) ? labels[grow + 1] : -100;
}

__global__ void __launch_bounds__(256)
ce_count_kernel(const long long* __restrict__ labels, int64_t R, int T, int shift,
                float* __restrict__ count_out) {
  __shared__ float red[33];
  float c = 0.f;
  for (int64_t r = threadIdx.x; r < R; r += blockDim.x)
    c += (shifted_label(labels, r, T, shift) != -100) ? 1.f : 0.f;
  c = block_sum(c, red);
  if (threadIdx.x == 0) *count_out = c;
}

// logits: [R, V] bf16 (row stride ld) for global rows row0..row0+R-1. Writes row_loss[row0+r]
// (0 for ignored rows) and, if write_grad, overwrites logits with gscale*(softmax-onehot)/count.
__global__ void __launch_bounds__(256)
ce_fwd_bwd_kernel(bf16* __restrict__ logits, int64_t ld, const long long* __restrict__ labels,
                  int64_t row0, int V, int T, int shift, float* __restrict__ row_loss,
                  const float* __restrict__ count, float gscale, int write_grad) {
  __shared__ float red[33];
  const int64_t grow = row0 + blockIdx.x;
  bf16* lr = logits + (int64_t)blockIdx.x * ld;
  const long long label = shifted_label(labels, grow, T, shift);
  const int nvec = V >> 3;
  if (label == -100) {
    if (threadIdx.x == 0) row_loss[grow] = 0.f;
    if (write_grad)
      for (int i = threadIdx.x; i < nvec; i += blockDim.x) stg16(lr + i * 8, make_uint4(0, 0, 0, 0));
    return;
  }
  float mx = -INFINITY, sm = 0.f;
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
    float v[8];
    unpack8(*reinterpret_cast<const uint4*>(lr + i * 8), v);
    float lm = v[0];
#pragma unroll
    for (int j = 1; j < 8; ++j) lm = fmaxf(lm, v[j]);
    if (lm > mx) {
      sm *= __expf(mx - lm);
      mx = lm;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) sm += __expf(v[j] - mx);
  }
  const float gmx = block_max(mx, red);
  sm *= __expf(mx - gmx);  // mx = -inf (thread saw nothing) → 0
  const float gsum = block_sum(sm, red);
  const float lse = gmx + logf(gsum);
  if (threadIdx.x == 0) row_loss[grow] = lse - __bfloat162float(lr[label]);
  if (write_grad) {
    const float sc = gscale / *count;
    __syncthreads();  // the label logit has been read
    for (int i = threadIdx.x; i < nvec; i += blockDim.x) {
      float v[8];
      unpack8(*reinterpret_cast<const uint4*>(lr + i * 8), v);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float p = __expf(v[j] - lse);
        if ((long long)(i * 8 + j) == label) p -= 1.f;
        v[j] = p * sc;
      }
      stg16(lr + i * 8, pack8(v));
    }
  }
}

// loss = sum(row_loss) / count   (deterministic single-CTA reduction)
__global__ void __launch_bounds__(1024)
ce_finalize_kernel(const float* __restrict__ row_loss, int64_t R, const float* __restrict__ count,
                   float* __restrict__ loss_out) {
  __shared__ float red[33];
  float s = 0.f;
  for (int64_t r = threadIdx.x; r < R; r += blockDim.x) s += row_loss[r];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *loss_out = s / *count;
}

// ---------------------------------------------------------------------------------------------
// distillation loss
// ---------------------------------------------------------------------------------------------
constexpr int DL_THREADS = 256;
constexpr int DL_CHUNK = 8192;  // elements of the flattened embedding per CTA
constexpr int DL_TI = 8, DL_TJ = 8;

__device__ __forceinline__ float smooth_l1(float d) {
  const float a = fabsf(d);
  return a < 1.f ? 0.5f * d * d : a - 0.5f;
}

// partials layout per chunk c: [B*Bt dots | B pn2 | Bt tn2 | B sl1]
__global__ void __launch_bounds__(DL_THREADS)
distill_partials_kernel(const bf16* __restrict__ pred, int64_t ldp, const bf16* __restrict__ tgt,
                        int64_t ldt, int64_t n, int B, int Bt, int off,
                        float* __restrict__ partials) {
  __shared__ float red[DL_TI * DL_TJ][DL_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int64_t e0 = (int64_t)blockIdx.x * DL_CHUNK;
  const int stride = B * Bt + B + Bt + B;
  float* out = partials + (int64_t)blockIdx.x * stride;
  const int iters = DL_CHUNK / (DL_THREADS * 8);

  auto block_reduce_store = [&](float* acc, int cnt, float* dst, auto&& dst_index) {
    // acc[0..cnt) per thread → block sums; dst_index(k) gives the output slot or -1
    for (int k = 0; k < cnt; ++k) {
      const float w = warp_sum(acc[k]);
      if (lane == 0) red[k][wid] = w;
    }
    __syncthreads();
    if (threadIdx.x < cnt) {
      float t = 0.f;
#pragma unroll
      for (int w = 0; w < DL_THREADS / 32; ++w) t += red[threadIdx.x][w];
      const int slot = dst_index(threadIdx.x);
      if (slot >= 0) dst[slot] = t;
    }
    __syncthreads();
  };

  for (int it = 0; it < B; it += DL_TI) {
    for (int jt = 0; jt < Bt; jt += DL_TJ) {
      float acc[DL_TI * DL_TJ];
#pragma unroll
      for (int k = 0; k < DL_TI * DL_TJ; ++k) acc[k] = 0.f;
      float pn[DL_TI], tn[DL_TJ];
#pragma unroll
      for (int k = 0; k < DL_TI; ++k) pn[k] = 0.f;
#pragma unroll
      for (int k = 0; k < DL_TJ; ++k) tn[k] = 0.f;
      for (int s = 0; s < iters; ++s) {
        const int64_t e = e0 + ((int64_t)s * DL_THREADS + threadIdx.x) * 8;
        if (e >= n) continue;  // n % 8 == 0
        float p[DL_TI][8];
#pragma unroll
        for (int ii = 0; ii < DL_TI; ++ii) {
          if (it + ii < B) unpack8(ldg16(pred + (int64_t)(it + ii) * ldp + e), p[ii]);
          else {
#pragma unroll
            for (int x = 0; x < 8; ++x) p[ii][x] = 0.f;
          }
        }
#pragma unroll
        for (int jj = 0; jj < DL_TJ; ++jj) {
          float t[8];
          if (jt + jj < Bt) unpack8(ldg16(tgt + (int64_t)(jt + jj) * ldt + e), t);
          else {
#pragma unroll
            for (int x = 0; x < 8; ++x) t[x] = 0.f;
          }
#pragma unroll
          for (int ii = 0; ii < DL_TI; ++ii) {
            float a = 0.f;
#pragma unroll
            for (int x = 0; x < 8; ++x) a += p[ii][x] * t[x];
            acc[ii * DL_TJ + jj] += a;
          }
          if (it == 0) {
#pragma unroll
            for (int x = 0; x < 8; ++x) tn[jj] += t[x] * t[x];
          }
        }
        if (jt == 0) {
#pragma unroll
          for (int ii = 0; ii < DL_TI; ++ii)
#pragma unroll
            for (int x = 0; x < 8; ++x) pn[ii] += p[ii][x] * p[ii][x];
        }
      }
      block_reduce_store(acc, DL_TI * DL_TJ, out, [&](int k) {
        const int i = it + k / DL_TJ, j = jt + k % DL_TJ;
        return (i < B && j < Bt) ? i * Bt + j : -1;
      });
      if (jt == 0)
        block_reduce_store(pn, DL_TI, out + B * Bt, [&](int k) { return it + k < B ? it + k : -1; });
      if (it == 0)
        block_reduce_store(tn, DL_TJ, out + B * Bt + B,
                           [&](int k) { return jt + k < Bt ? jt + k : -1; });
    }
  }
  // smooth-L1 of pred_i against its own target (row i + off)
  for (int i = 0; i < B; ++i) {
    float s1 = 0.f;
    const int j = i + off;
    for (int s = 0; s < iters; ++s) {
      const int64_t e = e0 + ((int64_t)s * DL_THREADS + threadIdx.x) * 8;
      if (e >= n) continue;
      float p[8], t[8];
      unpack8(ldg16(pred + (int64_t)i * ldp + e), p);
      unpack8(ldg16(tgt + (int64_t)j * ldt + e), t);
#pragma unroll
      for (int x = 0; x < 8; ++x) s1 += smooth_l1(p[x] - t[x]);
    }
    block_reduce_store(&s1, 1, out + B * Bt + B + Bt, [&](int) { return i; });
  }
}

// Single CTA. Reduces the chunk partials in a fixed order, then:
//   cos_ij = p_i·t_j / (max(|p_i|,1e-12) max(|t_j|,1e-12)) ; z_ij = cos_ij * min(exp(tau),100)
//   ce_i = logsumexp_j z_ij - z_{i,i+off}
//   sl1 = sum_i m_i S_i / (B n) ; con = cw * mean_i(ce_i) * mean_i(m_i) ; loss = sl1 + con
// out[0..2] = loss, sl1, con ; out[3] = d loss / d tau.
// coef layout (for the backward kernel): [B a_i | B d_i | B*Bt c_ij]
__global__ void __launch_bounds__(256)
distill_finalize_kernel(const float* __restrict__ partials, int nchunks, int B, int Bt, int off,
                        int64_t n, const float* __restrict__ tau, const float* __restrict__ mask,
                        float cw, float* __restrict__ out, float* __restrict__ coef,
                        float* __restrict__ stats) {
  extern __shared__ float sh[];
  const int stride = B * Bt + B + Bt + B;
  float* tot = sh;           // [stride]
  float* ce = sh + stride;   // [B]
  float* aux = ce + B;       // [B] d tau contributions
  for (int k = threadIdx.x; k < stride; k += blockDim.x) {
    float t = 0.f;
    for (int c = 0; c < nchunks; ++c) t += partials[(int64_t)c * stride + k];
    tot[k] = t;
    if (stats) stats[k] = t;
  }
  __syncthreads();
  const float* dots = tot;
  const float* pn2 = tot + B * Bt;
  const float* tn2 = pn2 + B;
  const float* sl1s = tn2 + Bt;
  const float et = __expf(*tau);
  const float scale = fminf(et, 100.f);
  const float dscale = et < 100.f ? et : 0.f;
  float mmean = 0.f;
  for (int i = 0; i < B; ++i) mmean += mask ? mask[i] : 1.f;
  mmean /= B;
  const float gce = cw * mmean / B;  // d loss / d ce_i
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float pni = fmaxf(sqrtf(pn2[i]), 1e-12f);
    float mx = -INFINITY;
    for (int j = 0; j < Bt; ++j) {
      const float z = dots[i * Bt + j] / (pni * fmaxf(sqrtf(tn2[j]), 1e-12f)) * scale;
      mx = fmaxf(mx, z);
    }
    float sm = 0.f;
    for (int j = 0; j < Bt; ++j) {
      const float z = dots[i * Bt + j] / (pni * fmaxf(sqrtf(tn2[j]), 1e-12f)) * scale;
      sm += __expf(z - mx);
    }
    const float lse = mx + logf(sm);
    const int lab = i + off;
    const float zl = dots[i * Bt + lab] / (pni * fmaxf(sqrtf(tn2[lab]), 1e-12f)) * scale;
    ce[i] = lse - zl;
    // backward coefficients
    float di = 0.f, dt = 0.f;
    for (int j = 0; j < Bt; ++j) {
      const float tnj = fmaxf(sqrtf(tn2[j]), 1e-12f);
      const float cosv = dots[i * Bt + j] / (pni * tnj);
      const float z = cosv * scale;
      const float G = gce * (__expf(z - lse) - (j == lab ? 1.f : 0.f));  // d loss / d z_ij
      coef[2 * B + i * Bt + j] = G * scale / (pni * tnj);                // * t_j
      di -= G * scale * cosv / (pni * pni);                              // * p_i
      dt += G * cosv * dscale;
    }
    coef[i] = (mask ? mask[i] : 1.f) / ((float)B * (float)n);  // * smooth_l1'(p - t)
    coef[B + i] = di;
    aux[i] = dt;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float s1 = 0.f, cem = 0.f, dt = 0.f;
    for (int i = 0; i < B; ++i) {
      s1 += (mask ? mask[i] : 1.f) * sl1s[i];
      cem += ce[i];
      dt += aux[i];
    }
    s1 /= ((float)B * (float)n);
    const float con = cw * (cem / B) * mmean;
    out[0] = s1 + con;
    out[1] = s1;
    out[2] = con;
    out[3] = dt;
  }
}

// dpred[i,e] = g * ( a_i * sl1'(p_ie - t_{i+off,e}) + d_i p_ie + sum_j c_ij t_je )
__global__ void __launch_bounds__(DL_THREADS)
distill_bwd_kernel(const bf16* __restrict__ pred, int64_t ldp, const bf16* __restrict__ tgt,
                   int64_t ldt, int64_t n, int B, int Bt, int off, const float* __restrict__ coef,
                   const float* __restrict__ gout, bf16* __restrict__ dpred, int64_t lddp) {
  const float g = gout ? *gout : 1.f;
  const int64_t nvec = n >> 3;
  for (int64_t vi = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; vi < nvec;
       vi += (int64_t)gridDim.x * blockDim.x) {
    const int64_t e = vi * 8;
    for (int it = 0; it < B; it += DL_TI) {
      float acc[DL_TI][8];
#pragma unroll
      for (int ii = 0; ii < DL_TI; ++ii) {
        const int i = it + ii;
        if (i < B) {
          float p[8], t[8];
          unpack8(ldg16(pred + (int64_t)i * ldp + e), p);
          unpack8(ldg16(tgt + (int64_t)(i + off) * ldt + e), t);
          const float a = coef[i], d = coef[B + i];
#pragma unroll
          for (int x = 0; x < 8; ++x) {
            const float df = p[x] - t[x];
            const float s1g = fabsf(df) < 1.f ? df : (df > 0.f ? 1.f : -1.f);
            acc[ii][x] = a * s1g + d * p[x];
          }
        } else {
#pragma unroll
          for (int x = 0; x < 8; ++x) acc[ii][x] = 0.f;
        }
      }
      for (int j = 0; j < Bt; ++j) {
        float t[8];
        unpack8(ldg16(tgt + (int64_t)j * ldt + e), t);
#pragma unroll
        for (int ii = 0; ii < DL_TI; ++ii) {
          const int i = it + ii;
          const float c = i < B ? coef[2 * B + i * Bt + j] : 0.f;
#pragma unroll
          for (int x = 0; x < 8; ++x) acc[ii][x] += c * t[x];
        }
      }
#pragma unroll
      for (int ii = 0; ii < DL_TI; ++ii) {
        const int i = it + ii;
        if (i < B) {
#pragma unroll
          for (int x = 0; x < 8; ++x) acc[ii][x] *= g;
          stg16(dpred + (int64_t)i * lddp + e, pack8(acc[ii]));
        }
      }
    }
  }
}

}  // namespace vpb

using namespace vpb;
#define ST(s) ((cudaStream_t)(s))

extern "C" int vpb_ce_count(const int64_t* labels, int64_t R, int T, int shift, float* count_out,
                            void* stream) {
  VPB_CHECK(R > 0 && T > 0, "ce_count: bad shape");
  ce_count_kernel<<<1, 256, 0, ST(stream)>>>((const long long*)labels, R, T, shift, count_out);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_ce_fwd_bwd(void* logits, int64_t ld, const int64_t* labels, int64_t row0, int R,
                              int V, int T, int shift, float* row_loss, const float* count,
                              float gscale, int write_grad, void* stream) {
  VPB_CHECK(R > 0 && V % 8 == 0 && ld % 8 == 0, "ce: bad shape R=%d V=%d", R, V);
  ce_fwd_bwd_kernel<<<R, 256, 0, ST(stream)>>>((bf16*)logits, ld, (const long long*)labels, row0, V,
                                               T, shift, row_loss, count, gscale, write_grad);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_ce_finalize(const float* row_loss, int64_t R, const float* count,
                               float* loss_out, void* stream) {
  ce_finalize_kernel<<<1, 1024, 0, ST(stream)>>>(row_loss, R, count, loss_out);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int64_t vpb_distill_workspace_floats(int B, int Bt, int64_t n) {
  const int64_t nchunks = (n + DL_CHUNK - 1) / DL_CHUNK;
  const int64_t stride = (int64_t)B * Bt + B + Bt + B;
  return nchunks * stride;
}

extern "C" int vpb_distill_loss_fwd(const void* pred, int64_t ldp, const void* tgt, int64_t ldt,
                                    int64_t n, int B, int Bt, int off, const float* tau,
                                    const float* mask, float contrastive_weight, float* workspace,
                                    float* out4, float* coef, float* stats, void* stream) {
  VPB_CHECK(n % 8 == 0 && n > 0 && B > 0 && Bt >= B + off && off >= 0, "distill: bad shape");
  VPB_CHECK(ldp % 8 == 0 && ldt % 8 == 0, "distill: row strides must be multiples of 8");
  const int nchunks = (int)((n + DL_CHUNK - 1) / DL_CHUNK);
  const int stride = B * Bt + B + Bt + B;
  VPB_CHECK((stride + 2 * B) * sizeof(float) <= 200 * 1024, "distill: batch too large");
  distill_partials_kernel<<<nchunks, DL_THREADS, 0, ST(stream)>>>(
      (const bf16*)pred, ldp, (const bf16*)tgt, ldt, n, B, Bt, off, workspace);
  VPB_LAUNCH_OK();
  const size_t sh = (stride + 2 * B) * sizeof(float);
  if (sh > 48 * 1024) {
    VPB_CUDA(cudaFuncSetAttribute(distill_finalize_kernel,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
  }
  distill_finalize_kernel<<<1, 256, sh, ST(stream)>>>(workspace, nchunks, B, Bt, off, n, tau, mask,
                                                      contrastive_weight, out4, coef, stats);
  VPB_LAUNCH_OK();
  return 0;
}

extern "C" int vpb_distill_loss_bwd(const void* pred, int64_t ldp, const void* tgt, int64_t ldt,
                                    int64_t n, int B, int Bt, int off, const float* coef,
                                    const float* gout, void* dpred, int64_t lddp, void* stream) {
  VPB_CHECK(n % 8 == 0 && n > 0 && B > 0 && lddp % 8 == 0, "distill_bwd: bad shape");
  int64_t blocks = (n / 8 + DL_THREADS - 1) / DL_THREADS;
  if (blocks > 148 * 8) blocks = 148 * 8;
  distill_bwd_kernel<<<(int)blocks, DL_THREADS, 0, ST(stream)>>>(
      (const bf16*)pred, ldp, (const bf16*)tgt, ldt, n, B, Bt, off, coef, gout, (bf16*)dpred, lddp);
  VPB_LAUNCH_OK();
  return 0;
}
