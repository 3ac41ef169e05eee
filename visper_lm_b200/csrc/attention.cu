// Flash-attention forward + backward (online softmax, never materialises the S×S matrix).
// Round-1 implementation on the warp-level tensor path (mma.sync m16n8k16 bf16, ldmatrix,
// cp.async double buffering, XOR-swizzled shared memory); the tcgen05/TMEM version replaces it
// in a later round.  Covers every attention on the hot path with one template:
//   * decoder causal GQA, head_dim 128 (Llama-3) / 96 (Phi-3) — replaces flash_attn_func selected
//     by attn_implementation="flash_attention_2" (/root/reference/ola_vlm/train/ola_vlm_train_mem.py:5)
//   * CLIP ViT-L patch attention, 16 heads × 64, S = 577, non-causal (clip_encoder.py:56)
//   * PerceiverAttention of the embedding-predictor heads, 4 heads × 32, queries = latents,
//     keys = cat(context, latents) given as TWO key/value segments so the concat is never built
//     (/root/reference/ola_vlm/model/multimodal_projector/resampler.py:46-75)
// Q/K/V/O are addressed in place inside packed row-major projections: row = b*seq + i,
// head h at columns [h*HD, (h+1)*HD).
#include "common.cuh"
#include "visper_b200.h"

namespace vpb {

struct AttnParams {
  const bf16 *q, *k, *v, *k2, *v2;
  bf16* o;
  float* lse;  // [B,H,sq] natural-log LSE of the scaled scores
  int64_t ldq, ldk, ldv, ldk2, ldv2, ldo;
  int B, H, KVH, sq, sk, sk2;
  float scale;
  int window;  // causal sliding window: key j visible to query i iff 0 <= i + off - j <= window (0: off)
  // backward only
  const bf16* dO;
  int64_t lddo;
  const float* delta;  // [B,H,sq]
  bf16 *dq, *dk, *dv, *dk2, *dv2;
  int64_t lddq, lddk, lddv, lddk2, lddv2;
  // additive score bias (forward only, vpb_attn_fwd_bias): scores = scale*q.k + bias[h] + bias_mask[b % mask_mod]
  const float* bias;       // [H, sq, sk] fp32
  const float* bias_mask;  // [mask_mod, sq, sk] fp32 or null
  int mask_mod;
};

template <int HD>
struct Cfg {
  static constexpr int HDP = (HD == 96) ? 128 : HD;  // padded smem row (elements)
  static constexpr int CPR = HDP / 8;                 // 16-byte chunks per smem row
  static constexpr int KS = HD / 16;                  // k-steps over head_dim
  static constexpr int ND = HD / 8;                   // 8-wide n-blocks over head_dim
};

template <int CPR>
__device__ __forceinline__ int swz(int row, int chunk) {
  if (CPR >= 8) return chunk ^ (row & 7);
  return chunk ^ ((row >> 1) & 3);
}
template <int CPR>
__device__ __forceinline__ uint32_t saddr(uint32_t base, int row, int chunk) {
  return base + (uint32_t)((row * CPR + swz<CPR>(row, chunk)) * 16);
}

// cp.async a [ROWS x HD] tile (rows >= rows_valid zero-filled) into swizzled shared memory
template <int ROWS, int HD, int NT>
__device__ __forceinline__ void load_tile(uint32_t sbase, const bf16* g, int64_t ld, int rows_valid) {
  constexpr int CPR = Cfg<HD>::CPR;
  constexpr int CH = HD / 8;
  for (int idx = threadIdx.x; idx < ROWS * CH; idx += NT) {
    const int r = idx / CH, c = idx - r * CH;
    const bool ok = r < rows_valid;
    cp_async16(saddr<CPR>(sbase, r, c), ok ? (const void*)(g + (int64_t)r * ld + c * 8) : (const void*)g, ok);
  }
}

// A fragment (16 rows starting at r0, k-step ks) from a row-major [rows][HD] tile
template <int CPR>
__device__ __forceinline__ void ld_a(uint32_t* f, uint32_t sbase, int r0, int ks, int lane) {
  ldsm_x4(f, saddr<CPR>(sbase, r0 + (lane & 15), 2 * ks + (lane >> 4)));
}
// B fragments for n-blocks nb, nb+1 (tile rows = n), k-step ks: f = {b0(nb), b1(nb), b0(nb+1), b1(nb+1)}
template <int CPR>
__device__ __forceinline__ void ld_b(uint32_t* f, uint32_t sbase, int nb, int ks, int lane) {
  const int id = lane >> 3;
  ldsm_x4(f, saddr<CPR>(sbase, (nb + (id >> 1)) * 8 + (lane & 7), 2 * ks + (id & 1)));
}
// B fragments from a tile whose rows are the contraction index: rows kk*16..+16, column chunks nd, nd+1
template <int CPR>
__device__ __forceinline__ void ld_bt(uint32_t* f, uint32_t sbase, int kk, int nd, int lane) {
  const int id = lane >> 3;
  ldsm_x4_t(f, saddr<CPR>(sbase, kk * 16 + (id & 1) * 8 + (lane & 7), nd + (id >> 1)));
}

constexpr float LOG2E = 1.4426950408889634f;

// =============================================================================================
// forward
// =============================================================================================
template <int HD, bool CAUSAL, bool BIAS = false>
__global__ void __launch_bounds__(256)
attn_fwd_kernel(const AttnParams p) {
  using C = Cfg<HD>;
  constexpr int CPR = C::CPR, KS = C::KS, ND = C::ND, HDP = C::HDP;
  constexpr int BMQ = 128, BN = 64, NT = 256;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK0 = sQ + BMQ * HDP * 2;
  const uint32_t sV0 = sK0 + 2 * BN * HDP * 2;

  const int qt = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.H / p.KVH);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int q0 = qt * BMQ;
  const int off = p.sk - p.sq;  // causal diagonal offset

  int n1 = (p.sk + BN - 1) / BN;
  int n2 = (p.sk2 + BN - 1) / BN;
  int jt0 = 0;  // first key tile any row of this query tile can see (sliding window)
  if (CAUSAL) {
    int kv_end = q0 + BMQ + off;
    if (kv_end > p.sk) kv_end = p.sk;
    if (kv_end < 0) kv_end = 0;
    n1 = (kv_end + BN - 1) / BN;
    n2 = 0;
    if (p.window > 0) {
      const int lo = q0 + off - p.window;
      jt0 = lo > 0 ? lo / BN : 0;
      if (jt0 > n1) jt0 = n1;
    }
  }
  const int ntiles = n1 + n2;

  const bf16* qg = p.q + ((int64_t)b * p.sq + q0) * p.ldq + h * HD;
  {
    int rv = p.sq - q0;
    load_tile<BMQ, HD, NT>(sQ, qg, p.ldq, rv < BMQ ? rv : BMQ);
  }
  auto issue_kv = [&](int jt, int buf) {
    const bool seg2 = jt >= n1;
    const int j0 = (seg2 ? jt - n1 : jt) * BN;
    const int len = seg2 ? p.sk2 : p.sk;
    const int64_t ldk = seg2 ? p.ldk2 : p.ldk, ldv = seg2 ? p.ldv2 : p.ldv;
    const bf16* kg = (seg2 ? p.k2 : p.k) + ((int64_t)b * len + j0) * ldk + kvh * HD;
    const bf16* vg = (seg2 ? p.v2 : p.v) + ((int64_t)b * len + j0) * ldv + kvh * HD;
    int rv = len - j0;
    if (rv > BN) rv = BN;
    load_tile<BN, HD, NT>(sK0 + buf * BN * HDP * 2, kg, ldk, rv);
    load_tile<BN, HD, NT>(sV0 + buf * BN * HDP * 2, vg, ldv, rv);
  };
  if (ntiles > jt0) issue_kv(jt0, jt0 & 1);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  uint32_t qf[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) ld_a<CPR>(qf[ks], sQ, warp * 16, ks, lane);

  float o[ND][4];
#pragma unroll
  for (int i = 0; i < ND; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const float sl2 = p.scale * LOG2E;

  for (int jt = jt0; jt < ntiles; ++jt) {
    const int buf = jt & 1;
    if (jt + 1 < ntiles) {
      issue_kv(jt + 1, buf ^ 1);
      cp_async_commit();
      cp_async_wait<1>();
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t sK = sK0 + buf * BN * HDP * 2, sV = sV0 + buf * BN * HDP * 2;

    float s[BN / 8][4];
#pragma unroll
    for (int i = 0; i < BN / 8; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
      for (int nb = 0; nb < BN / 8; nb += 2) {
        uint32_t bf[4];
        ld_b<CPR>(bf, sK, nb, ks, lane);
        mma16816(s[nb], qf[ks], bf);
        mma16816(s[nb + 1], qf[ks], bf + 2);
      }
    }
    const bool seg2 = jt >= n1;
    const int j0 = (seg2 ? jt - n1 : jt) * BN;
    const int len = seg2 ? p.sk2 : p.sk;
    if constexpr (BIAS) {
      // s holds the unscaled q.k; the softmax below scales by p.scale, so the additive terms go in
      // divided by it.  Fragment element e of n-block nb: row g + 8*(e>>1), column 2t + (e&1).
      const float inv = 1.f / p.scale;
      const float* bh = p.bias + (int64_t)h * p.sq * p.sk;
      const float* mw = p.bias_mask ? p.bias_mask + (int64_t)(b % p.mask_mod) * p.sq * p.sk : nullptr;
#pragma unroll
      for (int r = 0; r < 2; ++r) {
        const int i = q0 + warp * 16 + g + r * 8;
        if (i < p.sq) {
#pragma unroll
          for (int nb = 0; nb < BN / 8; ++nb) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
              const int j = j0 + nb * 8 + 2 * t + c;
              if (j < p.sk) {
                float a = __ldg(bh + (int64_t)i * p.sk + j);
                if (mw) a += __ldg(mw + (int64_t)i * p.sk + j);
                s[nb][2 * r + c] = fmaf(a, inv, s[nb][2 * r + c]);
              }
            }
          }
        }
      }
    }
    // masking (boundary tiles only)
    const bool win = CAUSAL && p.window > 0;
    const bool need_mask = (j0 + BN > len) || (CAUSAL && (j0 + BN - 1 > q0 + off)) ||
                           (win && j0 < q0 + BMQ - 1 + off - p.window);
    if (need_mask) {
#pragma unroll
      for (int nb = 0; nb < BN / 8; ++nb) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int j = j0 + nb * 8 + 2 * t + (e & 1);
          const int i = q0 + warp * 16 + g + (e >> 1) * 8;
          const bool ok = (j < len) && (!CAUSAL || j <= i + off) && (!win || j >= i + off - p.window);
          if (!ok) s[nb][e] = -INFINITY;
        }
      }
    }
    // online softmax
    float mnew[2], alpha[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float mx = mrow[r];
#pragma unroll
      for (int nb = 0; nb < BN / 8; ++nb) mx = fmaxf(mx, fmaxf(s[nb][2 * r], s[nb][2 * r + 1]));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 1));
      mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, 2));
      mnew[r] = mx;
      const float base = (mx == -INFINITY) ? 0.f : mx;
      alpha[r] = exp2f((mrow[r] - base) * sl2);  // mrow = -inf → 0
      mrow[r] = mx;
      lrow[r] *= alpha[r];
      const float mb = base * sl2;
#pragma unroll
      for (int nb = 0; nb < BN / 8; ++nb) {
        const float p0 = exp2f(s[nb][2 * r] * sl2 - mb);
        const float p1 = exp2f(s[nb][2 * r + 1] * sl2 - mb);
        s[nb][2 * r] = p0;
        s[nb][2 * r + 1] = p1;
        lrow[r] += p0 + p1;
      }
    }
#pragma unroll
    for (int nd = 0; nd < ND; ++nd) {
      o[nd][0] *= alpha[0];
      o[nd][1] *= alpha[0];
      o[nd][2] *= alpha[1];
      o[nd][3] *= alpha[1];
    }
    // O += P V
#pragma unroll
    for (int kk = 0; kk < BN / 16; ++kk) {
      uint32_t pf[4];
      pf[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
      pf[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
      pf[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      pf[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int nd = 0; nd < ND; nd += 2) {
        uint32_t vf[4];
        ld_bt<CPR>(vf, sV, kk, nd, lane);
        mma16816(o[nd], pf, vf);
        mma16816(o[nd + 1], pf, vf + 2);
      }
    }
    __syncthreads();
  }

  // finalize: normalise, stage through sQ (each warp owns its 16 rows), coalesced 16-byte stores
  float inv[2], lse[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    float l = lrow[r];
    l += __shfl_xor_sync(0xffffffffu, l, 1);
    l += __shfl_xor_sync(0xffffffffu, l, 2);
    inv[r] = l > 0.f ? 1.f / l : 0.f;
    lse[r] = (l > 0.f) ? mrow[r] * p.scale + logf(l) : -INFINITY;
  }
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) {
    const int r0 = warp * 16 + g;
    uint32_t a0 = saddr<CPR>(sQ, r0, nd) + t * 4;
    uint32_t a1 = saddr<CPR>(sQ, r0 + 8, nd) + t * 4;
    uint32_t v0 = pack2(o[nd][0] * inv[0], o[nd][1] * inv[0]);
    uint32_t v1 = pack2(o[nd][2] * inv[1], o[nd][3] * inv[1]);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a0), "r"(v0));
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a1), "r"(v1));
  }
  __syncwarp();
  bf16* og = p.o + ((int64_t)b * p.sq + q0) * p.ldo + h * HD;
  for (int idx = lane; idx < 16 * ND; idx += 32) {
    const int r = warp * 16 + idx / ND, c = idx % ND;
    if (q0 + r < p.sq) {
      uint4 val;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                   : "r"(saddr<CPR>(sQ, r, c)));
      stg16(og + (int64_t)r * p.ldo + c * 8, val);
    }
  }
  if (p.lse && t == 0) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = q0 + warp * 16 + g + r * 8;
      if (i < p.sq) p.lse[((int64_t)b * p.H + h) * p.sq + i] = lse[r];
    }
  }
}

// =============================================================================================
// backward
// =============================================================================================
// delta[b,h,i] = sum_d dO[b,i,h,d] * O[b,i,h,d]
// One warp per (b,i) row, HD/8 lanes per head (128-bit coalesced loads of both tensors), all index
// arithmetic in 32 bits and hoisted out of the head loop (the first version spent its time in
// 64-bit div/mod: ALU-bound at 81 % issue-slot use for an HBM-bound reduction).
template <int LPH>  // lanes per head = HD / 8 (4, 8, 12→16 padded, 16)
__global__ void __launch_bounds__(256)
attn_delta_kernel(const bf16* __restrict__ o, int64_t ldo, const bf16* __restrict__ dO, int64_t lddo,
                  float* __restrict__ delta, int B, int H, int sq, int HD) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + warp;  // b*sq + i
  if (row >= B * sq) return;
  const int b = row / sq, i = row - b * sq;
  constexpr int HPW = 32 / LPH;  // heads per warp pass
  const int sub = lane / LPH, l = lane % LPH;
  const bool lane_ok = l * 8 < HD;
  const bf16* orow = o + (int64_t)row * ldo + l * 8;
  const bf16* drow = dO + (int64_t)row * lddo + l * 8;
  float* dout = delta + (int64_t)b * H * sq + i;
#pragma unroll 4
  for (int h0 = 0; h0 < H; h0 += HPW) {
    const int h = h0 + sub;
    float acc = 0.f;
    if (h < H && lane_ok) {
      float a[8], d[8];
      unpack8(ldg16_stream(orow + h * HD), a);
      unpack8(ldg16_stream(drow + h * HD), d);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(a[j], d[j], acc);
    }
#pragma unroll
    for (int s = LPH / 2; s > 0; s >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, s);
    if (l == 0 && h < H) dout[(int64_t)h * sq] = acc;
  }
}

static void launch_delta(const bf16* o, int64_t ldo, const bf16* dO, int64_t lddo, float* delta, int B,
                         int H, int sq, int HD, cudaStream_t st) {
  const int grid = (B * sq + 7) / 8;
  if (HD <= 32) attn_delta_kernel<4><<<grid, 256, 0, st>>>(o, ldo, dO, lddo, delta, B, H, sq, HD);
  else if (HD <= 64) attn_delta_kernel<8><<<grid, 256, 0, st>>>(o, ldo, dO, lddo, delta, B, H, sq, HD);
  else attn_delta_kernel<16><<<grid, 256, 0, st>>>(o, ldo, dO, lddo, delta, B, H, sq, HD);
}

// dK, dV for one 64-row key tile of one kv head (all query heads of its GQA group, all query tiles)
template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(128)
attn_bwd_dkdv_kernel(const AttnParams p) {
  using C = Cfg<HD>;
  constexpr int CPR = C::CPR, KS = C::KS, ND = C::ND, HDP = C::HDP;
  constexpr int BKV = 64, BQ = (HD > 64) ? 32 : 64, NT = 128;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sK = smem_u32(smem);
  const uint32_t sV = sK + BKV * HDP * 2;
  const uint32_t sQ = sV + BKV * HDP * 2;
  const uint32_t sdO = sQ + BQ * HDP * 2;
  float* sLse = reinterpret_cast<float*>(smem + (2 * BKV + 2 * BQ) * HDP * 2);
  float* sDelta = sLse + BQ;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int n1 = (p.sk + BKV - 1) / BKV;
  const bool seg2 = (int)blockIdx.x >= n1;
  const int kv0 = (seg2 ? blockIdx.x - n1 : blockIdx.x) * BKV;
  const int len = seg2 ? p.sk2 : p.sk;
  const int kvh = blockIdx.y, b = blockIdx.z;
  const int G = p.H / p.KVH;
  const int off = p.sk - p.sq;
  const int64_t ldk = seg2 ? p.ldk2 : p.ldk, ldv = seg2 ? p.ldv2 : p.ldv;
  const bf16* kg = (seg2 ? p.k2 : p.k) + ((int64_t)b * len + kv0) * ldk + kvh * HD;
  const bf16* vg = (seg2 ? p.v2 : p.v) + ((int64_t)b * len + kv0) * ldv + kvh * HD;
  int kv_valid = len - kv0;
  if (kv_valid > BKV) kv_valid = BKV;
  load_tile<BKV, HD, NT>(sK, kg, ldk, kv_valid);
  load_tile<BKV, HD, NT>(sV, vg, ldv, kv_valid);
  cp_async_commit();

  float dk[ND][4], dv[ND][4];
#pragma unroll
  for (int i = 0; i < ND; ++i) {
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
  }
  const float sl2 = p.scale * LOG2E;
  const int nq_tiles = (p.sq + BQ - 1) / BQ;
  int qt_begin = 0, qt_end = nq_tiles;
  const bool win = CAUSAL && p.window > 0;
  if (CAUSAL) {
    int first = kv0 - off;  // first query row that can see key kv0
    if (first < 0) first = 0;
    qt_begin = first / BQ;
    if (win) {  // last query row that can see the tile's last key
      const int last = kv0 + BKV - 1 - off + p.window;
      if (last < 0) qt_end = 0;
      else if (last / BQ + 1 < qt_end) qt_end = last / BQ + 1;
    }
  }

  for (int hq = kvh * G; hq < (kvh + 1) * G; ++hq) {
    for (int qt = qt_begin; qt < qt_end; ++qt) {
      const int q0 = qt * BQ;
      __syncthreads();  // previous tile fully consumed
      int q_valid = p.sq - q0;
      if (q_valid > BQ) q_valid = BQ;
      load_tile<BQ, HD, NT>(sQ, p.q + ((int64_t)b * p.sq + q0) * p.ldq + hq * HD, p.ldq, q_valid);
      load_tile<BQ, HD, NT>(sdO, p.dO + ((int64_t)b * p.sq + q0) * p.lddo + hq * HD, p.lddo, q_valid);
      cp_async_commit();
      for (int i = threadIdx.x; i < BQ; i += NT) {
        const bool ok = i < q_valid;
        const int64_t li = ((int64_t)b * p.H + hq) * p.sq + q0 + i;
        sLse[i] = ok ? p.lse[li] * LOG2E : 0.f;
        sDelta[i] = ok ? p.delta[li] : 0.f;
      }
      cp_async_wait<0>();
      __syncthreads();

      // S^T = K_w Q^T and dP^T = V_w dO^T : [16 kv rows x BQ]
      float st[BQ / 8][4], dpt[BQ / 8][4];
#pragma unroll
      for (int i = 0; i < BQ / 8; ++i) {
        st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
        dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
      }
#pragma unroll
      for (int ks = 0; ks < KS; ++ks) {
        uint32_t kf[4], vf[4];
        ld_a<CPR>(kf, sK, warp * 16, ks, lane);
        ld_a<CPR>(vf, sV, warp * 16, ks, lane);
#pragma unroll
        for (int nb = 0; nb < BQ / 8; nb += 2) {
          uint32_t bq[4], bo[4];
          ld_b<CPR>(bq, sQ, nb, ks, lane);
          ld_b<CPR>(bo, sdO, nb, ks, lane);
          mma16816(st[nb], kf, bq);
          mma16816(st[nb + 1], kf, bq + 2);
          mma16816(dpt[nb], vf, bo);
          mma16816(dpt[nb + 1], vf, bo + 2);
        }
      }
      // P^T and dS^T (in place: st ← P^T, dpt ← dS^T)
#pragma unroll
      for (int nb = 0; nb < BQ / 8; ++nb) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int qc = nb * 8 + 2 * t + (e & 1);
          const int kr = warp * 16 + g + (e >> 1) * 8;
          const bool ok = (qc < q_valid) && (kr < kv_valid) && (!CAUSAL || kv0 + kr <= q0 + qc + off) &&
                          (!win || kv0 + kr >= q0 + qc + off - p.window);
          const float pv = ok ? exp2f(st[nb][e] * sl2 - sLse[qc]) : 0.f;
          st[nb][e] = pv;
          dpt[nb][e] = pv * (dpt[nb][e] - sDelta[qc]);
        }
      }
      // dV += P^T dO ; dK += dS^T Q   (contraction over the BQ query rows)
#pragma unroll
      for (int kk = 0; kk < BQ / 16; ++kk) {
        uint32_t pf[4], sf[4];
        pf[0] = pack2(st[2 * kk][0], st[2 * kk][1]);
        pf[1] = pack2(st[2 * kk][2], st[2 * kk][3]);
        pf[2] = pack2(st[2 * kk + 1][0], st[2 * kk + 1][1]);
        pf[3] = pack2(st[2 * kk + 1][2], st[2 * kk + 1][3]);
        sf[0] = pack2(dpt[2 * kk][0], dpt[2 * kk][1]);
        sf[1] = pack2(dpt[2 * kk][2], dpt[2 * kk][3]);
        sf[2] = pack2(dpt[2 * kk + 1][0], dpt[2 * kk + 1][1]);
        sf[3] = pack2(dpt[2 * kk + 1][2], dpt[2 * kk + 1][3]);
#pragma unroll
        for (int nd = 0; nd < ND; nd += 2) {
          uint32_t f1[4], f2[4];
          ld_bt<CPR>(f1, sdO, kk, nd, lane);
          ld_bt<CPR>(f2, sQ, kk, nd, lane);
          mma16816(dv[nd], pf, f1);
          mma16816(dv[nd + 1], pf, f1 + 2);
          mma16816(dk[nd], sf, f2);
          mma16816(dk[nd + 1], sf, f2 + 2);
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  // stage dK through sQ/sdO? BQ may be 32 rows only — use sK (dK) and sV (dV): all reads are done.
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) {
    const int r0 = warp * 16 + g;
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr<CPR>(sK, r0, nd) + t * 4),
                 "r"(pack2(dk[nd][0] * p.scale, dk[nd][1] * p.scale)));
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr<CPR>(sK, r0 + 8, nd) + t * 4),
                 "r"(pack2(dk[nd][2] * p.scale, dk[nd][3] * p.scale)));
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr<CPR>(sV, r0, nd) + t * 4),
                 "r"(pack2(dv[nd][0], dv[nd][1])));
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr<CPR>(sV, r0 + 8, nd) + t * 4),
                 "r"(pack2(dv[nd][2], dv[nd][3])));
  }
  __syncwarp();
  const int64_t lddk = seg2 ? p.lddk2 : p.lddk, lddv = seg2 ? p.lddv2 : p.lddv;
  bf16* dkg = (seg2 ? p.dk2 : p.dk) + ((int64_t)b * len + kv0) * lddk + kvh * HD;
  bf16* dvg = (seg2 ? p.dv2 : p.dv) + ((int64_t)b * len + kv0) * lddv + kvh * HD;
  for (int idx = lane; idx < 16 * ND; idx += 32) {
    const int r = warp * 16 + idx / ND, c = idx % ND;
    if (r < kv_valid) {
      uint4 a, d;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w)
                   : "r"(saddr<CPR>(sK, r, c)));
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(d.x), "=r"(d.y), "=r"(d.z), "=r"(d.w)
                   : "r"(saddr<CPR>(sV, r, c)));
      stg16(dkg + (int64_t)r * lddk + c * 8, a);
      stg16(dvg + (int64_t)r * lddv + c * 8, d);
    }
  }
}

// dQ for one 64-row query tile of one head
template <int HD, bool CAUSAL>
__global__ void __launch_bounds__(128)
attn_bwd_dq_kernel(const AttnParams p) {
  using C = Cfg<HD>;
  constexpr int CPR = C::CPR, KS = C::KS, ND = C::ND, HDP = C::HDP;
  constexpr int BQ = 64, BKV = (HD > 64) ? 32 : 64, NT = 128;
  extern __shared__ __align__(128) uint8_t smem[];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sdO = sQ + BQ * HDP * 2;
  const uint32_t sK = sdO + BQ * HDP * 2;
  const uint32_t sV = sK + BKV * HDP * 2;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int qt = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.H / p.KVH);
  const int q0 = qt * BQ;
  const int off = p.sk - p.sq;
  int q_valid = p.sq - q0;
  if (q_valid > BQ) q_valid = BQ;

  load_tile<BQ, HD, NT>(sQ, p.q + ((int64_t)b * p.sq + q0) * p.ldq + h * HD, p.ldq, q_valid);
  load_tile<BQ, HD, NT>(sdO, p.dO + ((int64_t)b * p.sq + q0) * p.lddo + h * HD, p.lddo, q_valid);
  cp_async_commit();

  float lse2[2], dl[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = warp * 16 + g + r * 8;
    const bool ok = i < q_valid;
    const int64_t li = ((int64_t)b * p.H + h) * p.sq + q0 + i;
    lse2[r] = ok ? p.lse[li] * LOG2E : 0.f;
    dl[r] = ok ? p.delta[li] : 0.f;
  }

  int n1 = (p.sk + BKV - 1) / BKV;
  int n2 = (p.sk2 + BKV - 1) / BKV;
  int jt0 = 0;
  const bool win = CAUSAL && p.window > 0;
  if (CAUSAL) {
    int kv_end = q0 + BQ + off;
    if (kv_end > p.sk) kv_end = p.sk;
    if (kv_end < 0) kv_end = 0;
    n1 = (kv_end + BKV - 1) / BKV;
    n2 = 0;
    if (win) {
      const int lo = q0 + off - p.window;
      jt0 = lo > 0 ? lo / BKV : 0;
      if (jt0 > n1) jt0 = n1;
    }
  }
  float dq[ND][4];
#pragma unroll
  for (int i = 0; i < ND; ++i) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  const float sl2 = p.scale * LOG2E;

  for (int jt = jt0; jt < n1 + n2; ++jt) {
    const bool seg2 = jt >= n1;
    const int j0 = (seg2 ? jt - n1 : jt) * BKV;
    const int len = seg2 ? p.sk2 : p.sk;
    const int64_t ldk = seg2 ? p.ldk2 : p.ldk, ldv = seg2 ? p.ldv2 : p.ldv;
    int kv_valid = len - j0;
    if (kv_valid > BKV) kv_valid = BKV;
    __syncthreads();
    load_tile<BKV, HD, NT>(sK, (seg2 ? p.k2 : p.k) + ((int64_t)b * len + j0) * ldk + kvh * HD, ldk, kv_valid);
    load_tile<BKV, HD, NT>(sV, (seg2 ? p.v2 : p.v) + ((int64_t)b * len + j0) * ldv + kvh * HD, ldv, kv_valid);
    cp_async_commit();
    cp_async_wait<0>();
    __syncthreads();

    float s[BKV / 8][4], dp[BKV / 8][4];
#pragma unroll
    for (int i = 0; i < BKV / 8; ++i) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t qf[4], of[4];
      ld_a<CPR>(qf, sQ, warp * 16, ks, lane);
      ld_a<CPR>(of, sdO, warp * 16, ks, lane);
#pragma unroll
      for (int nb = 0; nb < BKV / 8; nb += 2) {
        uint32_t bk[4], bv[4];
        ld_b<CPR>(bk, sK, nb, ks, lane);
        ld_b<CPR>(bv, sV, nb, ks, lane);
        mma16816(s[nb], qf, bk);
        mma16816(s[nb + 1], qf, bk + 2);
        mma16816(dp[nb], of, bv);
        mma16816(dp[nb + 1], of, bv + 2);
      }
    }
#pragma unroll
    for (int nb = 0; nb < BKV / 8; ++nb) {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int jc = nb * 8 + 2 * t + (e & 1);
        const int r = e >> 1;
        const int i = warp * 16 + g + r * 8;
        const bool ok = (i < q_valid) && (jc < kv_valid) && (!CAUSAL || j0 + jc <= q0 + i + off) &&
                        (!win || j0 + jc >= q0 + i + off - p.window);
        const float pv = ok ? exp2f(s[nb][e] * sl2 - lse2[r]) : 0.f;
        s[nb][e] = pv * (dp[nb][e] - dl[r]);  // dS
      }
    }
#pragma unroll
    for (int kk = 0; kk < BKV / 16; ++kk) {
      uint32_t sf[4];
      sf[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
      sf[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
      sf[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
      sf[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
      for (int nd = 0; nd < ND; nd += 2) {
        uint32_t kf[4];
        ld_bt<CPR>(kf, sK, kk, nd, lane);
        mma16816(dq[nd], sf, kf);
        mma16816(dq[nd + 1], sf, kf + 2);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) {
    const int r0 = warp * 16 + g;
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr<CPR>(sQ, r0, nd) + t * 4),
                 "r"(pack2(dq[nd][0] * p.scale, dq[nd][1] * p.scale)));
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(saddr<CPR>(sQ, r0 + 8, nd) + t * 4),
                 "r"(pack2(dq[nd][2] * p.scale, dq[nd][3] * p.scale)));
  }
  __syncwarp();
  bf16* dqg = p.dq + ((int64_t)b * p.sq + q0) * p.lddq + h * HD;
  for (int idx = lane; idx < 16 * ND; idx += 32) {
    const int r = warp * 16 + idx / ND, c = idx % ND;
    if (r < q_valid) {
      uint4 a;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w)
                   : "r"(saddr<CPR>(sQ, r, c)));
      stg16(dqg + (int64_t)r * p.lddq + c * 8, a);
    }
  }
}

// =============================================================================================
// EXPERIMENTAL (VPB_OPT_WIN_ATTN_V2, off by default, not yet run on hardware): window attention with score
// bias for windows of up to 144 tokens, head_dim 32.  One CTA per (window, head); nine warps own 16 query
// rows each and see ALL keys in one pass (18 n-blocks of scores in registers), so there is no online
// softmax, no second 128-row query tile with 16 live rows, and the bias comes in as float2 loads.
// =============================================================================================
constexpr int WIN_MAX = 144;
// MINB = resident CTAs per SM the register budget is cut for: 2 → 96 registers and ~260 B of spills,
// 1 → no spills but nine warps per SM; which one wins has to be measured.
template <int MINB>
__global__ void __launch_bounds__(288, MINB)
win_attn_fwd_kernel(const AttnParams p) {
  constexpr int HD = 32, CPR = Cfg<HD>::CPR, KS = Cfg<HD>::KS, ND = Cfg<HD>::ND, NT = 288, NB = WIN_MAX / 8;
  __shared__ __align__(128) uint8_t smem[3 * WIN_MAX * HD * 2];
  const uint32_t sQ = smem_u32(smem);
  const uint32_t sK = sQ + WIN_MAX * HD * 2;
  const uint32_t sV = sK + WIN_MAX * HD * 2;
  const int b = blockIdx.x, h = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int S = p.sq;  // == p.sk <= WIN_MAX

  load_tile<WIN_MAX, HD, NT>(sQ, p.q + (int64_t)b * S * p.ldq + h * HD, p.ldq, S);
  load_tile<WIN_MAX, HD, NT>(sK, p.k + (int64_t)b * S * p.ldk + h * HD, p.ldk, S);
  load_tile<WIN_MAX, HD, NT>(sV, p.v + (int64_t)b * S * p.ldv + h * HD, p.ldv, S);
  cp_async_commit();
  cp_async_wait<0>();
  __syncthreads();

  uint32_t qf[KS][4];
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) ld_a<CPR>(qf[ks], sQ, warp * 16, ks, lane);
  float s[NB][4];
#pragma unroll
  for (int i = 0; i < NB; ++i) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
  for (int ks = 0; ks < KS; ++ks) {
#pragma unroll
    for (int nb = 0; nb < NB; nb += 2) {
      uint32_t bf[4];
      ld_b<CPR>(bf, sK, nb, ks, lane);
      mma16816(s[nb], qf[ks], bf);
      mma16816(s[nb + 1], qf[ks], bf + 2);
    }
  }
  // scores = scale * q.k + bias[h] (+ mask[window]); -inf outside the window.  Element (2r+c) of n-block nb
  // is row g + 8r, column 8 nb + 2t + c.
  const float* bh = p.bias + (int64_t)h * S * S;
  const float* mw = p.bias_mask ? p.bias_mask + (int64_t)(b % p.mask_mod) * S * S : nullptr;
  const bool even = (S & 1) == 0;
  float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int i = warp * 16 + g + r * 8;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const int j = nb * 8 + 2 * t;
      float a0 = -INFINITY, a1 = -INFINITY;
      if (i < S && j < S) {
        const int64_t off = (int64_t)i * S + j;
        float b0, b1 = 0.f;
        if (even) {  // j even, S even → 8-byte aligned pair, both columns inside the window
          const float2 bb = __ldg(reinterpret_cast<const float2*>(bh + off));
          b0 = bb.x;
          b1 = bb.y;
          if (mw) {
            const float2 mm = __ldg(reinterpret_cast<const float2*>(mw + off));
            b0 += mm.x;
            b1 += mm.y;
          }
        } else {
          b0 = __ldg(bh + off) + (mw ? __ldg(mw + off) : 0.f);
          if (j + 1 < S) b1 = __ldg(bh + off + 1) + (mw ? __ldg(mw + off + 1) : 0.f);
        }
        a0 = fmaf(s[nb][2 * r], p.scale, b0);
        if (j + 1 < S) a1 = fmaf(s[nb][2 * r + 1], p.scale, b1);
      }
      s[nb][2 * r] = a0;
      s[nb][2 * r + 1] = a1;
      mx[r] = fmaxf(mx[r], fmaxf(a0, a1));
    }
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
    mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
  }
  float inv[2], lsum[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const float base = (mx[r] == -INFINITY) ? 0.f : mx[r] * LOG2E;
    float sum = 0.f;
#pragma unroll
    for (int nb = 0; nb < NB; ++nb) {
      const float p0 = exp2f(fmaf(s[nb][2 * r], LOG2E, -base));
      const float p1 = exp2f(fmaf(s[nb][2 * r + 1], LOG2E, -base));
      s[nb][2 * r] = p0;
      s[nb][2 * r + 1] = p1;
      sum += p0 + p1;
    }
    sum += __shfl_xor_sync(0xffffffffu, sum, 1);
    sum += __shfl_xor_sync(0xffffffffu, sum, 2);
    lsum[r] = sum;
    inv[r] = sum > 0.f ? 1.f / sum : 0.f;
  }
  float o[ND][4];
#pragma unroll
  for (int i = 0; i < ND; ++i) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
#pragma unroll
  for (int kk = 0; kk < WIN_MAX / 16; ++kk) {
    uint32_t pf[4];
    pf[0] = pack2(s[2 * kk][0], s[2 * kk][1]);
    pf[1] = pack2(s[2 * kk][2], s[2 * kk][3]);
    pf[2] = pack2(s[2 * kk + 1][0], s[2 * kk + 1][1]);
    pf[3] = pack2(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
    for (int nd = 0; nd < ND; nd += 2) {
      uint32_t vf[4];
      ld_bt<CPR>(vf, sV, kk, nd, lane);
      mma16816(o[nd], pf, vf);
      mma16816(o[nd + 1], pf, vf + 2);
    }
  }
  // each warp stages its own 16 output rows over its own (already consumed) Q rows, then 16-byte stores
  __syncwarp();
#pragma unroll
  for (int nd = 0; nd < ND; ++nd) {
    const int r0 = warp * 16 + g;
    const uint32_t a0 = saddr<CPR>(sQ, r0, nd) + t * 4;
    const uint32_t a1 = saddr<CPR>(sQ, r0 + 8, nd) + t * 4;
    const uint32_t v0 = pack2(o[nd][0] * inv[0], o[nd][1] * inv[0]);
    const uint32_t v1 = pack2(o[nd][2] * inv[1], o[nd][3] * inv[1]);
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a0), "r"(v0));
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(a1), "r"(v1));
  }
  __syncwarp();
  bf16* og = p.o + (int64_t)b * S * p.ldo + h * HD;
  for (int idx = lane; idx < 16 * ND; idx += 32) {
    const int r = warp * 16 + idx / ND, c = idx % ND;
    if (r < S) {
      uint4 val;
      asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                   : "=r"(val.x), "=r"(val.y), "=r"(val.z), "=r"(val.w)
                   : "r"(saddr<CPR>(sQ, r, c)));
      stg16(og + (int64_t)r * p.ldo + c * 8, val);
    }
  }
  if (p.lse && t == 0) {
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int i = warp * 16 + g + r * 8;
      if (i < S) p.lse[((int64_t)b * p.H + h) * S + i] = lsum[r] > 0.f ? mx[r] + logf(lsum[r]) : -INFINITY;
    }
  }
}

// ---------------------------------------------------------------------------------------------
template <int HD, bool CAUSAL, bool BIAS = false>
static int launch_fwd(const AttnParams& p, cudaStream_t st) {
  constexpr int HDP = Cfg<HD>::HDP;
  constexpr int SMEM = (128 + 4 * 64) * HDP * 2;
  auto kern = attn_fwd_kernel<HD, CAUSAL, BIAS>;
  static bool cfg = false;
  if (!cfg) {
    VPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    cfg = true;
  }
  dim3 grid((p.sq + 127) / 128, p.H, p.B);
  kern<<<grid, 256, SMEM, st>>>(p);
  VPB_LAUNCH_OK();
  return 0;
}
template <int HD, bool CAUSAL>
static int launch_bwd(const AttnParams& p, cudaStream_t st) {
  constexpr int HDP = Cfg<HD>::HDP;
  constexpr int BSMALL = (HD > 64) ? 32 : 64;
  {
    constexpr int SMEM = (2 * 64 + 2 * BSMALL) * HDP * 2 + 2 * BSMALL * 4;
    auto kern = attn_bwd_dkdv_kernel<HD, CAUSAL>;
    static bool cfg = false;
    if (!cfg) {
      VPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      cfg = true;
    }
    dim3 grid((p.sk + 63) / 64 + (p.sk2 + 63) / 64, p.KVH, p.B);
    kern<<<grid, 128, SMEM, st>>>(p);
    VPB_LAUNCH_OK();
  }
  {
    constexpr int SMEM = (2 * 64 + 2 * BSMALL) * HDP * 2;
    auto kern = attn_bwd_dq_kernel<HD, CAUSAL>;
    static bool cfg = false;
    if (!cfg) {
      VPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
      cfg = true;
    }
    dim3 grid((p.sq + 63) / 64, p.H, p.B);
    kern<<<grid, 128, SMEM, st>>>(p);
    VPB_LAUNCH_OK();
  }
  return 0;
}

static int check_attn(const AttnParams& p, int HD) {
  VPB_CHECK(HD == 32 || HD == 64 || HD == 96 || HD == 128, "attention: unsupported head_dim %d", HD);
  VPB_CHECK(p.B > 0 && p.H > 0 && p.KVH > 0 && p.H % p.KVH == 0 && p.sq > 0 && p.sk > 0 && p.sk2 >= 0,
            "attention: bad shape B=%d H=%d KVH=%d sq=%d sk=%d sk2=%d", p.B, p.H, p.KVH, p.sq, p.sk, p.sk2);
  VPB_CHECK(p.ldq % 8 == 0 && p.ldk % 8 == 0 && p.ldv % 8 == 0 && p.ldo % 8 == 0,
            "attention: row strides must be multiples of 8 elements");
  return 0;
}

}  // namespace vpb

namespace vpb {
int attn_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                void* o, int64_t ldo, float* lse, int B, int H, int KVH, int sq, int sk, int head_dim,
                float scale, int causal, int window, cudaStream_t st);
int attn_bwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                const void* dO, int64_t lddo, const float* lse, const float* delta, void* dq,
                int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int KVH,
                int sq, int sk, int head_dim, float scale, int causal, int window,
                const float* rope_cos, const float* rope_sin, const int* rope_pos, int* rope_fused,
                cudaStream_t st);
}
using namespace vpb;

#define DISPATCH_HD(HDv, CAUSALv, FN, ...)                                  \
  do {                                                                      \
    if (CAUSALv) {                                                          \
      switch (HDv) {                                                        \
        case 32: return FN<32, true>(__VA_ARGS__);                          \
        case 64: return FN<64, true>(__VA_ARGS__);                          \
        case 96: return FN<96, true>(__VA_ARGS__);                          \
        default: return FN<128, true>(__VA_ARGS__);                         \
      }                                                                     \
    } else {                                                                \
      switch (HDv) {                                                        \
        case 32: return FN<32, false>(__VA_ARGS__);                         \
        case 64: return FN<64, false>(__VA_ARGS__);                         \
        case 96: return FN<96, false>(__VA_ARGS__);                         \
        default: return FN<128, false>(__VA_ARGS__);                        \
      }                                                                     \
    }                                                                       \
  } while (0)

extern "C" int vpb_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                            int64_t ldv, const void* k2, int64_t ldk2, const void* v2, int64_t ldv2,
                            void* o, int64_t ldo, float* lse, int B, int H, int KVH, int sq, int sk,
                            int sk2, int head_dim, float scale, int causal, int window,
                            void* stream) {
  AttnParams p = {};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v;
  p.k2 = (const bf16*)k2; p.v2 = (const bf16*)v2;
  p.o = (bf16*)o; p.lse = lse;
  p.window = (causal && window > 0 && window < sk) ? window : 0;
  VPB_CHECK(window >= 0 && (window == 0 || causal), "attention: a sliding window needs causal=1");
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldk2 = ldk2; p.ldv2 = ldv2; p.ldo = ldo;
  p.B = B; p.H = H; p.KVH = KVH; p.sq = sq; p.sk = sk; p.sk2 = k2 ? sk2 : 0;
  p.scale = scale;
  if (check_attn(p, head_dim)) return -1;
  VPB_CHECK(!(causal && p.sk2 > 0), "attention: causal with a second key segment is not supported");
  // tcgen05 path: head_dim 128 (Llama-3) and 96 (Phi-3); the 96 kernels need full 128-row tiles to exist.
  // head_dim 64, non-causal (the CLIP / DINOv2 ViT towers) rides the same forward kernel behind
  // VPB_OPT_ATTN_FWD_TC64: written after the round's GPU budget was spent, off until validated on hardware.
  const bool tc_hd = ((head_dim == 128 || head_dim == 96) && !(causal && sk < sq && (head_dim == 96 || p.window > 0))) ||
                     (head_dim == 64 && !causal && get_option(VPB_OPT_ATTN_FWD_TC64));
  if (tc_hd && p.sk2 == 0 && !get_option(VPB_OPT_ATTN_LEGACY_FWD) &&
      (reinterpret_cast<uintptr_t>(q) & 15) == 0 && (reinterpret_cast<uintptr_t>(k) & 15) == 0 &&
      (reinterpret_cast<uintptr_t>(v) & 15) == 0)
    return attn_fwd_tc(q, ldq, k, ldk, v, ldv, o, ldo, lse, B, H, KVH, sq, sk, head_dim, scale, causal,
                       p.window, (cudaStream_t)stream);
  DISPATCH_HD(head_dim, causal, launch_fwd, p, (cudaStream_t)stream);
}

// Window attention with an additive score bias (Swin: relative-position bias per head + the
// shifted-window mask per window), forward only, head_dim 32, non-causal, one K/V segment.
extern "C" int vpb_attn_fwd_bias(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                 int64_t ldv, void* o, int64_t ldo, float* lse, int B, int H, int sq,
                                 int sk, int head_dim, float scale, const float* bias,
                                 const float* bias_mask, int mask_mod, void* stream) {
  AttnParams p = {};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v;
  p.o = (bf16*)o; p.lse = lse;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
  p.B = B; p.H = H; p.KVH = H; p.sq = sq; p.sk = sk; p.sk2 = 0;
  p.scale = scale;
  p.bias = bias; p.bias_mask = bias_mask; p.mask_mod = bias_mask ? mask_mod : 1;
  if (check_attn(p, head_dim)) return -1;
  VPB_CHECK(head_dim == 32, "attn_fwd_bias: head_dim %d (only 32 is built)", head_dim);
  VPB_CHECK(bias != nullptr && scale > 0.f, "attn_fwd_bias: bias table missing or scale <= 0");
  VPB_CHECK(!bias_mask || (mask_mod > 0 && B % mask_mod == 0), "attn_fwd_bias: B=%d is not a multiple of mask_mod=%d", B, mask_mod);
  if (get_option(VPB_OPT_WIN_ATTN_V2) && sq == sk && sq <= WIN_MAX) {  // experimental one-pass window kernel
    if (get_option(VPB_OPT_WIN_ATTN_V2) == 2)
      win_attn_fwd_kernel<1><<<dim3(B, H), 288, 0, (cudaStream_t)stream>>>(p);
    else
      win_attn_fwd_kernel<2><<<dim3(B, H), 288, 0, (cudaStream_t)stream>>>(p);
    VPB_LAUNCH_OK();
    return 0;
  }
  VPB_CHECK(B <= 65535, "attn_fwd_bias: B=%d windows exceed gridDim.z; split the batch", B);
  return launch_fwd<32, false, true>(p, (cudaStream_t)stream);
}

extern "C" int vpb_rope_inplace(void* x, int64_t ld, int M, int seq_len, const int* pos_ids,
                                const float* cos_t, const float* sin_t, int nheads, int head_dim,
                                int inverse, void* stream);

// rope_cos != null: dQ and dK come back rotated by the inverse RoPE (the gradient w.r.t. the
// pre-rotation projections) — in the tcgen05 v2 epilogues when they run, by the rope kernel otherwise
static int attn_bwd_impl(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                         int64_t ldv, const void* k2, int64_t ldk2, const void* v2, int64_t ldv2,
                         const void* o, int64_t ldo, const void* dO, int64_t lddo,
                         const float* lse, float* delta, void* dq, int64_t lddq, void* dk,
                         int64_t lddk, void* dv, int64_t lddv, void* dk2, int64_t lddk2,
                         void* dv2, int64_t lddv2, int B, int H, int KVH, int sq, int sk, int sk2,
                         int head_dim, float scale, int causal, int window, const float* rope_cos,
                         const float* rope_sin, const int* rope_pos, int* rope_fused, void* stream) {
  if (rope_fused) *rope_fused = 0;
  AttnParams p = {};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v;
  p.k2 = (const bf16*)k2; p.v2 = (const bf16*)v2;
  p.lse = const_cast<float*>(lse);
  p.window = (causal && window > 0 && window < sk) ? window : 0;
  VPB_CHECK(window >= 0 && (window == 0 || causal), "attention: a sliding window needs causal=1");
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldk2 = ldk2; p.ldv2 = ldv2; p.ldo = ldo;
  p.B = B; p.H = H; p.KVH = KVH; p.sq = sq; p.sk = sk; p.sk2 = k2 ? sk2 : 0;
  p.scale = scale;
  p.dO = (const bf16*)dO; p.lddo = lddo; p.delta = delta;
  p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv; p.dk2 = (bf16*)dk2; p.dv2 = (bf16*)dv2;
  p.lddq = lddq; p.lddk = lddk; p.lddv = lddv; p.lddk2 = lddk2; p.lddv2 = lddv2;
  if (check_attn(p, head_dim)) return -1;
  VPB_CHECK(lddo % 8 == 0 && lddq % 8 == 0 && lddk % 8 == 0 && lddv % 8 == 0, "attention bwd: strides");
  VPB_CHECK(!(causal && p.sk2 > 0), "attention: causal with a second key segment is not supported");
  VPB_CHECK((int64_t)B * sq < (1ll << 31), "attention bwd: B*sq too large");
  launch_delta((const bf16*)o, ldo, (const bf16*)dO, lddo, delta, B, H, sq, head_dim, (cudaStream_t)stream);
  VPB_LAUNCH_OK();
  auto al16 = [](const void* x) { return (reinterpret_cast<uintptr_t>(x) & 15) == 0; };
  const bool tc_hd = (head_dim == 128 || head_dim == 96) && !(causal && sk < sq && (head_dim == 96 || p.window > 0));
  if (tc_hd && p.sk2 == 0 && !get_option(VPB_OPT_ATTN_LEGACY_BWD) && al16(q) &&
      al16(k) && al16(v) && al16(dO))
    return attn_bwd_tc(q, ldq, k, ldk, v, ldv, dO, lddo, lse, delta, dq, lddq, dk, lddk, dv, lddv, B,
                       H, KVH, sq, sk, head_dim, scale, causal, p.window, rope_cos, rope_sin, rope_pos,
                       rope_fused, (cudaStream_t)stream);
  DISPATCH_HD(head_dim, causal, launch_bwd, p, (cudaStream_t)stream);
}

extern "C" int vpb_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                            int64_t ldv, const void* k2, int64_t ldk2, const void* v2, int64_t ldv2,
                            const void* o, int64_t ldo, const void* dO, int64_t lddo,
                            const float* lse, float* delta, void* dq, int64_t lddq, void* dk,
                            int64_t lddk, void* dv, int64_t lddv, void* dk2, int64_t lddk2,
                            void* dv2, int64_t lddv2, int B, int H, int KVH, int sq, int sk, int sk2,
                            int head_dim, float scale, int causal, int window, void* stream) {
  return attn_bwd_impl(q, ldq, k, ldk, v, ldv, k2, ldk2, v2, ldv2, o, ldo, dO, lddo, lse, delta, dq,
                       lddq, dk, lddk, dv, lddv, dk2, lddk2, dv2, lddv2, B, H, KVH, sq, sk, sk2,
                       head_dim, scale, causal, window, nullptr, nullptr, nullptr, nullptr, stream);
}

extern "C" int vpb_attn_bwd_rope(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                 int64_t ldv, const void* o, int64_t ldo, const void* dO,
                                 int64_t lddo, const float* lse, float* delta, void* dq, int64_t lddq,
                                 void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int KVH,
                                 int seq_len, int head_dim, float scale, int causal, int window,
                                 const float* cos_t, const float* sin_t, const int* pos_ids,
                                 void* stream) {
  VPB_CHECK(cos_t && sin_t, "attn_bwd_rope: rotary tables missing");
  int fused = 0;
  if (attn_bwd_impl(q, ldq, k, ldk, v, ldv, nullptr, 0, nullptr, 0, o, ldo, dO, lddo, lse, delta, dq,
                    lddq, dk, lddk, dv, lddv, nullptr, 0, nullptr, 0, B, H, KVH, seq_len, seq_len, 0,
                    head_dim, scale, causal, window, cos_t, sin_t, pos_ids, &fused, stream))
    return -1;
  if (fused) return 0;
  // kernels without the fused epilogue (legacy / v1 / head_dim != 128): rotate dQ and dK afterwards
  if (vpb_rope_inplace(dq, lddq, B * seq_len, seq_len, pos_ids, cos_t, sin_t, H, head_dim, 1, stream))
    return -1;
  return vpb_rope_inplace(dk, lddk, B * seq_len, seq_len, pos_ids, cos_t, sin_t, KVH, head_dim, 1, stream);
}
