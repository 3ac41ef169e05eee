// tcgen05 / TMEM flash attention for head_dim 128 (the Llama-3 decoder: causal GQA, 32 q-heads /
// 8 kv-heads).  Replaces flash_attn's FA2 (mma.sync) kernels selected by
// attn_implementation="flash_attention_2" (/root/reference/ola_vlm/train/ola_vlm_train_mem.py:5).
//
// Forward, one CTA per (128-query tile, head, batch), 192 threads:
//   warp 0    TMA producer: Q once, then a 2-stage ring of K/V tiles (128 keys x 128 dims, two
//             128-byte-swizzled 64-column boxes each) read in place from the packed QKV projection
//   warp 1    TMEM allocator + single-thread tcgen05.mma issuer:
//               S_j  = Q · K_jᵀ        (SS, K-major A and B)      → TMEM S[j&1]   (128 fp32 columns)
//               O   += P_j · V_j       (SS, V as MN-major B)      → TMEM O        (128 fp32 columns)
//             S_{j+1} is issued before P_j is awaited, so QKᵀ of the next tile overlaps the softmax
//   warps 2-5 softmax: one query row per thread (TMEM lane = row): tcgen05.ld the 128 scores, online
//             max / sum in the log2 domain with LAZY rescaling (O in TMEM is only multiplied when the
//             running max moved by more than 2^8), P_j → bf16 → swizzled shared memory for the PV MMA.
// Saves LSE [B,H,sq] in natural-log units, same contract as the mma.sync kernels in attention.cu.
#include "common.cuh"
#include "visper_b200.h"

namespace vpb {

struct AttnTcParams {
  bf16* o;
  float* lse;
  int64_t ldo;
  int B, H, KVH, sq, sk;
  float scale;
};

namespace tc {
constexpr int BM = 128, BN = 128, HD = 128;
constexpr int TILE_BYTES = 128 * 128 * 2;  // 32 KB: two 64-column chunks of [128 rows x 128 B]
constexpr int CHUNK_BYTES = 16384;
constexpr int OFF_Q = 0;
constexpr int OFF_KV = TILE_BYTES;                   // stage s: K at OFF_KV + s*2*TILE, V right after
constexpr int OFF_P = OFF_KV + 4 * TILE_BYTES;
constexpr int OFF_BAR = OFF_P + TILE_BYTES;
constexpr int SMEM_BYTES = OFF_BAR + 256 + 1024;
constexpr float LOG2E = 1.4426950408889634f;
}  // namespace tc

template <bool CAUSAL>
__global__ void __launch_bounds__(192, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [2]
  uint64_t* kv_empty = bars + 3;  // [2]
  uint64_t* s_full = bars + 5;    // [2]
  uint64_t* p_full = bars + 7;
  uint64_t* pv_done = bars + 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;  // heavy tiles first
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.H / p.KVH);
  const int q0 = qt * BM;
  const int off = p.sk - p.sq;
  int kv_end = p.sk;
  if (CAUSAL) {
    kv_end = q0 + BM + off;
    if (kv_end > p.sk) kv_end = p.sk;
    if (kv_end < 0) kv_end = 0;
  }
  const int ntiles = (kv_end + BN - 1) / BN;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(p_full, 4);
    mbar_init(pv_done, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t TM_S = tmem_base;        // S[0] at +0, S[1] at +128
  const uint32_t TM_O = tmem_base + 256;  // O accumulator

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, TILE_BYTES);
      tma_load_2d(smem + OFF_Q, &tmQ, q_full, h * HD, b * p.sq + q0);
      tma_load_2d(smem + OFF_Q + CHUNK_BYTES, &tmQ, q_full, h * HD + 64, b * p.sq + q0);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j & 1;
        mbar_wait(&kv_empty[s], ((j >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * TILE_BYTES);
        uint8_t* sk = smem + OFF_KV + s * 2 * TILE_BYTES;
        uint8_t* sv = sk + TILE_BYTES;
        const int row = b * p.sk + j * BN;
        tma_load_2d(sk, &tmK, &kv_full[s], kvh * HD, row);
        tma_load_2d(sk + CHUNK_BYTES, &tmK, &kv_full[s], kvh * HD + 64, row);
        tma_load_2d(sv, &tmV, &kv_full[s], kvh * HD, row);
        tma_load_2d(sv + CHUNK_BYTES, &tmV, &kv_full[s], kvh * HD + 64, row);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && ntiles > 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(BM, BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(BM, HD, 0, 1);
      const uint32_t sq_addr = smem_u32(smem + OFF_Q);
      const uint32_t sp_addr = smem_u32(smem + OFF_P);
      auto issue_s = [&](int j) {
        const int s = j & 1;
        mbar_wait(&kv_full[s], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t sk_addr = smem_u32(smem + OFF_KV + s * 2 * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t o = (k >> 2) * CHUNK_BYTES + (k & 3) * 32;
          umma_bf16(TM_S + s * BN, make_smem_desc(sq_addr + o, 16, 1024),
                    make_smem_desc(sk_addr + o, 16, 1024), idesc_s, k != 0);
        }
        umma_commit(&s_full[s]);
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_s(j + 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        const uint32_t sv_addr = smem_u32(smem + OFF_KV + (j & 1) * 2 * TILE_BYTES + TILE_BYTES);
#pragma unroll
        for (int k = 0; k < BN / 16; ++k) {
          const uint32_t oa = (k >> 2) * CHUNK_BYTES + (k & 3) * 32;  // P: K-major over keys
          umma_bf16(TM_O, make_smem_desc(sp_addr + oa, 16, 1024),
                    make_smem_desc(sv_addr + k * 2048, CHUNK_BYTES, 1024), idesc_pv,
                    (j | k) != 0);
        }
        umma_commit(pv_done);
        umma_commit(&kv_empty[j & 1]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    float m_used = -INFINITY, l = 0.f;
    uint8_t* prow = smem + OFF_P + row * 128;
    for (int j = 0; j < ntiles; ++j) {
      const int sb = j & 1;
      mbar_wait(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      uint32_t r[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(TM_S + lane_addr + sb * BN + c * 32, r + c * 32);
      tmem_ld_wait();
      const int j0 = j * BN;
      const bool need_mask = (j0 + BN > p.sk) || (CAUSAL && (j0 + BN - 1 > q0 + off));
      float mx = -INFINITY;
      if (need_mask) {
        const int lim = CAUSAL ? min(p.sk - 1, q0 + row + off) : p.sk - 1;  // last visible key
#pragma unroll
        for (int c = 0; c < 128; ++c) {
          float v = __uint_as_float(r[c]) * sl2;
          if (j0 + c > lim) v = -INFINITY;
          r[c] = __float_as_uint(v);
          mx = fmaxf(mx, v);
        }
      } else {
#pragma unroll
        for (int c = 0; c < 128; ++c) {
          const float v = __uint_as_float(r[c]) * sl2;
          r[c] = __float_as_uint(v);
          mx = fmaxf(mx, v);
        }
      }
      // lazy rescale: keep the old reference max unless the new one is > 2^8 larger
      const bool upd = mx > m_used + 8.f;
      float alpha = 1.f;
      if (upd) {
        alpha = exp2f(m_used - mx);  // m_used = -inf → 0
        m_used = mx;
        l *= alpha;
      }
      if (j > 0) {
        mbar_wait(pv_done, (j - 1) & 1);  // PV_{j-1} finished: O and the P buffer are ours
        tc_fence_after();
        if (__any_sync(0xffffffffu, upd)) {
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld32(TM_O + lane_addr + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(TM_O + lane_addr + c * 32, o);
          }
          tmem_st_wait();
        }
      }
      const float mb = (m_used == -INFINITY) ? 0.f : m_used;
      float sum = 0.f;
#pragma unroll
      for (int u = 0; u < 16; ++u) {  // 16-byte units of 8 keys
        float pv[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          pv[i] = exp2f(__uint_as_float(r[u * 8 + i]) - mb);
          sum += pv[i];
        }
        const uint4 packed = pack8(pv);
        const int chunk = u >> 3, uu = u & 7;
        *reinterpret_cast<uint4*>(prow + chunk * CHUNK_BYTES + ((uu ^ (row & 7)) << 4)) = packed;
      }
      l += sum;
      fence_proxy_async();  // generic-proxy smem writes → visible to the tensor-core (async) proxy
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: O / l → bf16 → global, LSE
    const bool row_ok = q0 + row < p.sq;
    if (ntiles > 0) {
      mbar_wait(pv_done, (ntiles - 1) & 1);
      tc_fence_after();
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    bf16* orow = p.o + ((int64_t)b * p.sq + q0 + row) * p.ldo + h * HD;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      if (ntiles > 0) {
        tmem_ld32(TM_O + lane_addr + c * 32, o);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) o[i] = 0;
      }
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(o[g * 8 + i]) * inv;
          stg16(orow + c * 32 + g * 8, pack8(v));
        }
      }
    }
    if (p.lse && row_ok)
      p.lse[((int64_t)b * p.H + h) * p.sq + q0 + row] =
          l > 0.f ? m_used * 0.6931471805599453f + logf(l) : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <bool CAUSAL>
static int launch_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                         int64_t ldv, const AttnTcParams& p, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV;
  if (make_tmap_2d(&tmQ, q, (uint64_t)p.H * tc::HD, (uint64_t)p.B * p.sq, (uint64_t)ldq, 64, 128)) return -1;
  if (make_tmap_2d(&tmK, k, (uint64_t)p.KVH * tc::HD, (uint64_t)p.B * p.sk, (uint64_t)ldk, 64, 128)) return -1;
  if (make_tmap_2d(&tmV, v, (uint64_t)p.KVH * tc::HD, (uint64_t)p.B * p.sk, (uint64_t)ldv, 64, 128)) return -1;
  auto kern = attn_fwd_tc_kernel<CAUSAL>;
  static bool cfg = false;
  if (!cfg) {
    VPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    cfg = true;
  }
  dim3 grid((p.sq + tc::BM - 1) / tc::BM, p.H, p.B);
  kern<<<grid, 192, tc::SMEM_BYTES, st>>>(tmQ, tmK, tmV, p);
  VPB_LAUNCH_OK();
  return 0;
}

// entry used by vpb_attn_fwd (attention.cu) for head_dim 128, single K/V segment
int attn_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                void* o, int64_t ldo, float* lse, int B, int H, int KVH, int sq, int sk, float scale,
                int causal, cudaStream_t st) {
  AttnTcParams p;
  p.o = (bf16*)o;
  p.lse = lse;
  p.ldo = ldo;
  p.B = B; p.H = H; p.KVH = KVH; p.sq = sq; p.sk = sk;
  p.scale = scale;
  return causal ? launch_fwd_tc<true>(q, ldq, k, ldk, v, ldv, p, st)
                : launch_fwd_tc<false>(q, ldq, k, ldk, v, ldv, p, st);
}

}  // namespace vpb
