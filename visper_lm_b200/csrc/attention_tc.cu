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
#include <type_traits>
#include "common.cuh"
#include "visper_b200.h"

namespace vpb {

struct AttnTcParams {
  bf16* o;
  float* lse;
  int64_t ldo;
  int B, H, KVH, sq, sk;
  float scale;
  int window;  // causal sliding window: key j visible to query i iff 0 <= i + off - j <= window (0: off)
};

namespace tc {
constexpr int BM = 128, BN = 128;  // head_dim (128 or 96) is a template parameter of the kernel
constexpr int TILE_BYTES = 128 * 128 * 2;  // 32 KB: two 64-column chunks of [128 rows x 128 B]
constexpr int CHUNK_BYTES = 16384;
constexpr int OFF_Q = 0;
constexpr int KST = 3, VST = 2;                      // K ring runs ahead of the V ring
constexpr int OFF_K = TILE_BYTES;                    // K stage s at OFF_K + s*TILE
constexpr int OFF_V = OFF_K + KST * TILE_BYTES;      // V stage s at OFF_V + s*TILE
constexpr int OFF_BAR = OFF_V + VST * TILE_BYTES;
constexpr int OFF_X = OFF_BAR + 256;  // softmax exchange: [2][NS][128] maxima + [NS][128] sums, NS <= 4: 1536 floats
constexpr int SMEM_BYTES = OFF_X + 6144 + 1024;
constexpr int THREADS = 320;  // TMA warp + MMA warp + 8 softmax warps
constexpr float LOG2E = 1.4426950408889634f;
}  // namespace tc

__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// Two variants of this kernel were measured on B200 in round 2 and dropped (profiles/r02_variants_ab.txt): every
// fourth exponential as a degree-3 polynomial on the FMA pipe (0.440 vs 0.427 ms — the softmax warps are bound by
// instruction issue and latency, not by the MUFU rate) and Q resident in TMEM as the A operand of QK^T (0.450 ms).
// NS = softmax warps per TMEM lane quarter (they split the 128 key columns): 2 → 8 softmax warps.  NS = 4 (16 warps,
// 32 scores per thread) is supported by the code below and was measured no faster (0.447 vs 0.431 ms,
// profiles/r02_attn_fwd_experiments.txt), like two other restructurings of this one-work-item-per-CTA kernel (three
// score buffers with software-pipelined TMEM reads: 0.497 ms; two co-resident CTAs per SM: 0.453 ms).  What they all
// share is the per-CTA fixed cost — see attn_fwd_tc_persist_kernel below, which is the default.
template <bool CAUSAL, int HD, int NS>
__global__ void __launch_bounds__(64 + 128 * NS, 1)
attn_fwd_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                   const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* k_full = bars + 1;    // [3]
  uint64_t* k_empty = bars + 4;   // [3]
  uint64_t* v_full = bars + 7;    // [2]
  uint64_t* v_empty = bars + 9;   // [2]
  uint64_t* s_full = bars + 11;   // [2]
  uint64_t* p_full = bars + 13;
  uint64_t* pv_done = bars + 14;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 16);

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
  // sliding window: the first key tile any row of this query tile can see; the loops below run
  // over ntiles tiles starting there (ring indices stay 0-based)
  int jb = 0;
  if (CAUSAL && p.window > 0) {
    const int lo = q0 + off - p.window;
    jb = lo > 0 ? lo / BN : 0;
    if (jb * BN > kv_end) jb = kv_end / BN;
  }
  const int ntiles = (kv_end + BN - 1) / BN - jb;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < KST; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
    }
    for (int i = 0; i < VST; ++i) {
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
    }
    mbar_init(p_full, 4 * NS);
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
      // head_dim 64 (ViT towers) is ONE 64-column chunk per tile; 96 / 128 add a second TMA box
      constexpr uint32_t TX_BYTES = HD > 64 ? TILE_BYTES : CHUNK_BYTES;
      mbar_arrive_expect_tx(q_full, TX_BYTES);
      tma_load_2d(smem + OFF_Q, &tmQ, q_full, h * HD, b * p.sq + q0);
      if constexpr (HD > 64) tma_load_2d(smem + OFF_Q + CHUNK_BYTES, &tmQ, q_full, h * HD + 64, b * p.sq + q0);
      auto load_k = [&](int j) {
        const int s = j % KST;
        mbar_wait(&k_empty[s], ((j / KST) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[s], TX_BYTES);
        uint8_t* sk = smem + OFF_K + s * TILE_BYTES;
        const int row = b * p.sk + (j + jb) * BN;
        tma_load_2d(sk, &tmK, &k_full[s], kvh * HD, row);
        if constexpr (HD > 64) tma_load_2d(sk + CHUNK_BYTES, &tmK, &k_full[s], kvh * HD + 64, row);
      };
      if (ntiles > 0) load_k(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) load_k(j + 1);  // K runs one tile ahead of V
        const int s = j % VST;
        mbar_wait(&v_empty[s], ((j / VST) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[s], TX_BYTES);
        uint8_t* sv = smem + OFF_V + s * TILE_BYTES;
        const int row = b * p.sk + (j + jb) * BN;
        tma_load_2d(sv, &tmV, &v_full[s], kvh * HD, row);
        if constexpr (HD > 64) tma_load_2d(sv + CHUNK_BYTES, &tmV, &v_full[s], kvh * HD + 64, row);
      }
    }
  } else if (warp == 1) {
    if (ntiles > 0) {  // warp-uniform loop, one elected lane issues (descriptors stay uniform)
      const bool leader = elect_one();
      constexpr uint32_t idesc_s = make_idesc_bf16(BM, BN, 0, 0);
      constexpr uint32_t idesc_pv = make_idesc_bf16(BM, HD, 0, 1);
      const uint64_t q_desc = make_smem_desc(smem_u32(smem + OFF_Q), 16, 1024);
      const uint64_t k_desc0 = make_smem_desc(smem_u32(smem + OFF_K), 16, 1024);
      const uint64_t v_desc0 = make_smem_desc(smem_u32(smem + OFF_V), CHUNK_BYTES, 1024);  // MN-major
      auto issue_s = [&](int j) {
        const int s = j % KST, sb = j & 1;
        mbar_wait_spin(&k_full[s], (j / KST) & 1);
        tc_fence_after();
        if (leader) {
          const uint64_t k_desc = desc_adv(k_desc0, s * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < HD / 16; ++k) {
            const uint32_t o = (k >> 2) * CHUNK_BYTES + (k & 3) * 32;
            umma_bf16(TM_S + sb * BN, desc_adv(q_desc, o), desc_adv(k_desc, o), idesc_s, k != 0);
          }
          umma_commit(&s_full[sb]);
          umma_commit(&k_empty[s]);
        }
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_s(j + 1);
        mbar_wait_spin(&v_full[j % VST], (j / VST) & 1);
        mbar_wait_spin(p_full, j & 1);
        tc_fence_after();
        if (leader) {
          const uint64_t v_desc = desc_adv(v_desc0, (j % VST) * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BN / 16; ++k) {
            // A = P_j in TMEM, written by the softmax warps over the first 64 columns of S[j&1]
            umma_bf16_ts(TM_O, TM_S + (j & 1) * BN + k * 8, desc_adv(v_desc, k * 2048), idesc_pv,
                         (j | k) != 0);
          }
          umma_commit(pv_done);
          umma_commit(&v_empty[j % VST]);
        }
      }
    }
  } else {
    // 4*NS softmax warps: warps w, w+4, ... share the 32 TMEM lanes of quarter w&3 and split the 128 key
    // columns into NS parts of CW, so every SM sub-partition runs NS softmax warps.
    constexpr int CW = 128 / NS;          // score columns per thread
    constexpr int OC = 4 / NS;            // 32-column chunks of O per thread (rescale / epilogue)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;     // part index 0..NS-1
    const int row = quarter * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    float m_used = -INFINITY, l = 0.f;
    float* xchg = reinterpret_cast<float*>(smem + OFF_X);  // [2 parity][NS parts][128 rows] + [NS][128]
    for (int j = 0; j < ntiles; ++j) {
      const int sb = j & 1;
      mbar_wait_spin(&s_full[sb], (j >> 1) & 1);
      tc_fence_after();
      uint32_t r[CW];
      tmem_ld32(TM_S + lane_addr + sb * BN + half * CW, r);
      if constexpr (CW == 64) tmem_ld32(TM_S + lane_addr + sb * BN + half * CW + 32, r + 32);
      tmem_ld_wait();
      const int jt0 = (j + jb) * BN;  // first key of this tile
      const int j0 = jt0 + half * CW;
      const bool win = CAUSAL && p.window > 0;
      const bool need_mask = (jt0 + BN > p.sk) || (CAUSAL && (jt0 + BN - 1 > q0 + off)) ||
                             (win && jt0 < q0 + BM - 1 + off - p.window);
      if (need_mask) {
        const int lim = CAUSAL ? min(p.sk - 1, q0 + row + off) : p.sk - 1;  // last visible key
        const int lo = win ? q0 + row + off - p.window : 0;                 // first visible key
#pragma unroll
        for (int c = 0; c < CW; ++c)
          if (j0 + c > lim || j0 + c < lo) r[c] = 0xff800000u;  // -inf
      }
      // max of the RAW scores (scaled once, scale > 0); four independent chains for ILP
      float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
      for (int c = 0; c < CW; c += 4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) mx4[e] = fmaxf(mx4[e], __uint_as_float(r[c + e]));
      }
      float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
      xchg[(sb * NS + half) * 128 + row] = mx;
      named_bar_sync(1, 128 * NS);
#pragma unroll
      for (int o = 1; o < NS; ++o) mx = fmaxf(mx, xchg[(sb * NS + ((half + o) % NS)) * 128 + row]);
      mx *= sl2;
      // lazy rescale: keep the old reference max unless the new one is > 2^8 larger
      const bool upd = mx > m_used + 8.f;
      float alpha = 1.f;
      if (upd) {
        alpha = exp2f(m_used - mx);  // m_used = -inf → 0
        m_used = mx;
        l *= alpha;
      }
      // exponentials first (registers only) so they overlap the PV MMA of the previous tile
      const float mb = (m_used == -INFINITY) ? 0.f : m_used;
      float sum4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
      for (int c = 0; c < CW; c += 4) {
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const float pv = ex2_approx(fmaf(__uint_as_float(r[c + e]), sl2, -mb));
          sum4[e] += pv;
          r[c + e] = __float_as_uint(pv);
        }
      }
      const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
      if (j > 0 && __any_sync(0xffffffffu, upd)) {
        mbar_wait(pv_done, (j - 1) & 1);  // PV_{j-1} finished: O may be rescaled
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < OC; ++c) {
          const int oc = (half * OC + c) * 32;  // this thread's 32-column chunk of O
          if (oc >= HD) break;                  // head_dim 96 / 64: the last parts own fewer chunks
          uint32_t o[32];
          tmem_ld32(TM_O + lane_addr + oc, o);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
          tmem_st32(TM_O + lane_addr + oc, o);
        }
      }
      // P_j (bf16 pairs) overwrites the row's own scores in TMEM: packed columns [0,64) of S[sb]
      {
        uint32_t w[CW / 2];
#pragma unroll
        for (int i = 0; i < CW / 2; ++i) w[i] = pack2(__uint_as_float(r[2 * i]), __uint_as_float(r[2 * i + 1]));
        if constexpr (CW == 64) tmem_st32(TM_S + lane_addr + sb * BN + half * 32, w);
        else tmem_st16(TM_S + lane_addr + sb * BN + half * 16, w);
        tmem_st_wait();
      }
      l += sum;
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    // epilogue: combine the NS partial row sums, O / l → bf16 → global, LSE
    float* lx = xchg + 2 * NS * 128;
    lx[half * 128 + row] = l;
    named_bar_sync(1, 128 * NS);
#pragma unroll
    for (int o = 1; o < NS; ++o) l += lx[((half + o) % NS) * 128 + row];
    const bool row_ok = q0 + row < p.sq;
    if (ntiles > 0) {
      mbar_wait(pv_done, (ntiles - 1) & 1);
      tc_fence_after();
    }
    const float inv = l > 0.f ? 1.f / l : 0.f;
    bf16* orow = p.o + ((int64_t)b * p.sq + q0 + row) * p.ldo + h * HD + half * OC * 32;
#pragma unroll 1
    for (int c = 0; c < OC; ++c) {
      if ((half * OC + c) * 32 >= HD) break;
      uint32_t o[32];
      if (ntiles > 0) {
        tmem_ld32(TM_O + lane_addr + (half * OC + c) * 32, o);
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
    if (p.lse && row_ok && half == 0)
      p.lse[((int64_t)b * p.H + h) * p.sq + q0 + row] =
          l > 0.f ? m_used * 0.6931471805599453f + logf(l) : -INFINITY;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// =============================================================================================
// forward, persistent (default)
// =============================================================================================
// tools/attn_overhead_probe.py (profiles/r02_attn_overhead_probe.jsonl) fits the kernel above as
//     time per CTA = 6.07 us + 1.10 us x key tiles
// on B200: a 128-query work item pays ~11 500 cycles that are not tile work — CTA launch behind the previous CTA's
// exit (192 KB of shared memory, 512 TMEM columns: nothing co-resides), barrier init + TMEM allocation, tensor-map
// fetch and the Q / K_0 round trip, the first Q·Kᵀ, the pipeline ramp, the O epilogue, deallocation.  A causal
// S = 2048 item has 8.5 key tiles on average, so 39 % of the 0.427 ms was that fixed cost (27.7 items per SM x 6 us).
// Here ONE CTA per SM stays resident and walks a heavy-first list of work items.  Its TMA warp runs ahead across item
// boundaries (Q double-buffered, the K / V rings never drain), its MMA warp issues the first Q·Kᵀ of item n+1 before
// it waits for the last P of item n, O is double-buffered in TMEM and a separate EPILOGUE warpgroup writes item n out
// while the softmax warps are already on item n+1, and barriers / TMEM are set up once.  Same arithmetic in the same order as the kernel above:
// bit-identical output and LSE.
namespace tcp {
constexpr int KST = 2, VST = 2;
constexpr int OFF_Q = 0;                                   // Q[2]
constexpr int OFF_K = 2 * tc::TILE_BYTES;                  // K[2]
constexpr int OFF_V = OFF_K + KST * tc::TILE_BYTES;        // V[2]
constexpr int OFF_BAR = OFF_V + VST * tc::TILE_BYTES;      // 6 x 32 KB of tiles
constexpr int OFF_X = OFF_BAR + 256;                       // max exchange [2][2][128] + per-item [2][l0|l1|m][128]: 1280 floats
constexpr int SMEM_BYTES = OFF_X + 5120 + 1024;
constexpr int THREADS = 512;
}  // namespace tcp

// n-th work item of this CTA in a persistent kernel.  The work lists are sorted heavy-first and dealt round by round; dealing
// every second round in reverse CTA order ("snake") evens out who gets the heavier side of the rounds in which the item size
// steps down: the busiest CTA is 0.6 % above the mean instead of 2.6 % (forward, B=8, H=32, S=2048), 1.2 % instead of 4.7 % at
// Phi-3's B=4.  Neighbouring CTAs still hold neighbouring items (the heads of a GQA group share K/V in L2).
__device__ __forceinline__ int persist_item(int n) {
  return n * (int)gridDim.x + ((n & 1) ? (int)gridDim.x - 1 - (int)blockIdx.x : (int)blockIdx.x);
}

struct FwdItem {
  int h, b, kvh, q0, jb, ntiles;
};

// work item w of the heavy-first list: query tiles in descending order (causal: most key tiles first), the heads of
// one batch entry adjacent (a GQA group's CTAs read the same K / V tiles while they are hot in L2)
template <bool CAUSAL>
__device__ __forceinline__ bool fwd_item(const AttnTcParams& p, int w, int nqt, FwdItem& it) {
  const int per = p.H * p.B;
  if (w >= nqt * per) return false;
  const int qi = w / per, r = w - qi * per;
  const int qt = CAUSAL ? nqt - 1 - qi : qi;
  it.h = r % p.H;
  it.b = r / p.H;
  it.kvh = it.h / (p.H / p.KVH);
  it.q0 = qt * tc::BM;
  const int off = p.sk - p.sq;
  int kv_end = p.sk;
  if (CAUSAL) {
    kv_end = it.q0 + tc::BM + off;
    if (kv_end > p.sk) kv_end = p.sk;
  }
  it.jb = 0;
  if (CAUSAL && p.window > 0) {  // sliding window: first key tile any row of this query tile can see
    const int lo = it.q0 + off - p.window;
    it.jb = lo > 0 ? lo / tc::BN : 0;
    if (it.jb * tc::BN > kv_end) it.jb = kv_end / tc::BN;
  }
  it.ntiles = (kv_end + tc::BN - 1) / tc::BN - it.jb;  // >= 1: the launcher only takes shapes with sk >= sq
  return true;
}

template <bool CAUSAL, int HD>
__global__ void __launch_bounds__(512, 1)
attn_fwd_tc_persist_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                           const __grid_constant__ CUtensorMap tmV, const AttnTcParams p) {
  using namespace tc;
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment by OFFSET on the __shared__ array (not through an integer cast): the compiler keeps the
  // shared address space, so the softmax exchange slots are LDS / STS instead of generic loads through L1TEX
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + tcp::OFF_BAR);
  uint64_t* q_full = bars + 0;    // [2]
  uint64_t* q_empty = bars + 2;   // [2] every Q·Kᵀ of the item that used the buffer has completed
  uint64_t* k_full = bars + 4;    // [2]
  uint64_t* k_empty = bars + 6;   // [2]
  uint64_t* v_full = bars + 8;    // [2]
  uint64_t* v_empty = bars + 10;  // [2]
  uint64_t* s_full = bars + 12;   // [2]
  // p_full / pv_done are indexed by tile parity as well: a waiter is never more than two tiles behind the arriving
  // side (the scores of tile g+2 are only issued after P·V of tile g), so with two barriers a phase can not be
  // observed one wrap late whatever the TMA latency does to the relative timing
  uint64_t* p_full = bars + 14;   // [2]
  uint64_t* pv_done = bars + 16;  // [2]
  uint64_t* o_free = bars + 18;   // [2] the epilogue warps have read O[n&1] and the row sums of item n
  uint64_t* o_ready = bars + 20;  // [2] every P·V of item n has completed (committed after its last tile)
  uint64_t* l_ready = bars + 22;  // [2] the softmax warps have published the row sums / maxima of item n
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 24);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nqt = (p.sq + BM - 1) / BM;
  const int off = p.sk - p.sq;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&o_free[i], 4);
      mbar_init(&o_ready[i], 1);
      mbar_init(&l_ready[i], 8);
      mbar_init(&p_full[i], 8);
      mbar_init(&pv_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t TM_S = tmem_base;        // S[0] at +0, S[1] at +128
  const uint32_t TM_O = tmem_base + 256;  // O[0] at +256, O[1] at +384
  constexpr uint32_t TX_BYTES = HD > 64 ? TILE_BYTES : CHUNK_BYTES;

  // 512 threads = four warpgroups: {TMA warp, MMA warp, two idle warps}, two softmax warpgroups, one epilogue
  // warpgroup.  128 registers per thread at launch; the data-movement and epilogue groups hand registers to the
  // softmax groups (setmaxnreg INSIDE each role's branch, so that ptxas budgets each region separately).
  if (warp < 4) {
  setmaxnreg_dec<96>();
  if (warp == 0) {
    if (lane == 0) {
      // global tile sequence (item 0 tile 0, item 0 tile 1, ..., item 1 tile 0, ...); g counts tiles, K is requested
      // one tile ahead of V, and the Q of an item right before its first K
      auto load_q = [&](int n, const FwdItem& it) {
        const int s = n & 1;
        mbar_wait(&q_empty[s], ((n >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[s], TX_BYTES);
        uint8_t* sq_ = smem + tcp::OFF_Q + s * TILE_BYTES;
        tma_load_2d(sq_, &tmQ, &q_full[s], it.h * HD, it.b * p.sq + it.q0);
        if constexpr (HD > 64) tma_load_2d(sq_ + CHUNK_BYTES, &tmQ, &q_full[s], it.h * HD + 64, it.b * p.sq + it.q0);
      };
      auto load_k = [&](int g, const FwdItem& it, int j) {
        const int s = g % tcp::KST;
        mbar_wait(&k_empty[s], ((g / tcp::KST) & 1) ^ 1);
        mbar_arrive_expect_tx(&k_full[s], TX_BYTES);
        uint8_t* sk_ = smem + tcp::OFF_K + s * TILE_BYTES;
        const int row = it.b * p.sk + (j + it.jb) * BN;
        tma_load_2d(sk_, &tmK, &k_full[s], it.kvh * HD, row);
        if constexpr (HD > 64) tma_load_2d(sk_ + CHUNK_BYTES, &tmK, &k_full[s], it.kvh * HD + 64, row);
      };
      auto load_v = [&](int g, const FwdItem& it, int j) {
        const int s = g % tcp::VST;
        mbar_wait(&v_empty[s], ((g / tcp::VST) & 1) ^ 1);
        mbar_arrive_expect_tx(&v_full[s], TX_BYTES);
        uint8_t* sv_ = smem + tcp::OFF_V + s * TILE_BYTES;
        const int row = it.b * p.sk + (j + it.jb) * BN;
        tma_load_2d(sv_, &tmV, &v_full[s], it.kvh * HD, row);
        if constexpr (HD > 64) tma_load_2d(sv_ + CHUNK_BYTES, &tmV, &v_full[s], it.kvh * HD + 64, row);
      };
      FwdItem cur, nxt;
      int n = 0, g = 0;
      bool have = fwd_item<CAUSAL>(p, persist_item(0), nqt, cur);
      if (have) {
        load_q(0, cur);
        load_k(0, cur, 0);
      }
      while (have) {
        const bool have_next = fwd_item<CAUSAL>(p, persist_item(n + 1), nqt, nxt);
        // Q of the NEXT item is requested as soon as its buffer can be free (the previous item's last Q·Kᵀ was issued
        // two producer tiles ago): Q rows are read once, so this is a DRAM round trip that must not sit in front of
        // the next item's first MMA
        const int jq = cur.ntiles > 1 ? 1 : 0;
        for (int j = 0; j < cur.ntiles; ++j, ++g) {
          if (j + 1 < cur.ntiles) {
            load_k(g + 1, cur, j + 1);
          } else if (have_next) {  // the tile after this item's last one is the next item's first
            load_k(g + 1, nxt, 0);
          }
          load_v(g, cur, j);
          if (j == jq && have_next) load_q(n + 1, nxt);
        }
        cur = nxt;
        have = have_next;
        ++n;
      }
    }
  } else if (warp == 1) {
    const bool leader = elect_one();  // warp-uniform loops, one elected lane issues (descriptors stay uniform)
    constexpr uint32_t idesc_s = make_idesc_bf16(BM, BN, 0, 0);
    constexpr uint32_t idesc_pv = make_idesc_bf16(BM, HD, 0, 1);
    const uint64_t q_desc0 = make_smem_desc(smem_u32(smem + tcp::OFF_Q), 16, 1024);
    const uint64_t k_desc0 = make_smem_desc(smem_u32(smem + tcp::OFF_K), 16, 1024);
    const uint64_t v_desc0 = make_smem_desc(smem_u32(smem + tcp::OFF_V), CHUNK_BYTES, 1024);  // MN-major
    // S of global tile g (item n): Q[n&1] · K[g%2]ᵀ → S[g&1]; `last` = the item's last tile (its Q buffer is then free)
    auto issue_s = [&](int g, int n, bool last) {
      const int s = g % tcp::KST, sb = g & 1;
      mbar_wait_spin(&k_full[s], (g / tcp::KST) & 1);
      tc_fence_after();
      if (leader) {
        const uint64_t q_desc = desc_adv(q_desc0, (n & 1) * TILE_BYTES);
        const uint64_t k_desc = desc_adv(k_desc0, s * TILE_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t o = (k >> 2) * CHUNK_BYTES + (k & 3) * 32;
          umma_bf16(TM_S + sb * BN, desc_adv(q_desc, o), desc_adv(k_desc, o), idesc_s, k != 0);
        }
        umma_commit(&s_full[sb]);
        umma_commit(&k_empty[s]);
        if (last) umma_commit(&q_empty[n & 1]);
      }
    };
    FwdItem cur, nxt;
    int n = 0, g = 0;
    bool have = fwd_item<CAUSAL>(p, persist_item(0), nqt, cur);
    if (have) {
      mbar_wait(&q_full[0], 0);
      issue_s(0, 0, cur.ntiles == 1);
    }
    while (have) {
      bool have_next = false;
      for (int j = 0; j < cur.ntiles; ++j, ++g) {
        // the next tile's scores are issued before this tile's P is awaited: Q·Kᵀ overlaps the softmax, also across
        // the boundary between two work items
        if (j + 1 < cur.ntiles) {
          issue_s(g + 1, n, j + 2 == cur.ntiles);
        } else {
          have_next = fwd_item<CAUSAL>(p, persist_item(n + 1), nqt, nxt);
          if (have_next) {
            mbar_wait_spin(&q_full[(n + 1) & 1], ((n + 1) >> 1) & 1);
            issue_s(g + 1, n + 1, nxt.ntiles == 1);
          }
        }
        mbar_wait_spin(&v_full[g % tcp::VST], (g / tcp::VST) & 1);
        mbar_wait_spin(&p_full[g & 1], (g >> 1) & 1);
        if (j == 0) mbar_wait_spin(&o_free[n & 1], ((n >> 1) & 1) ^ 1);  // the epilogue warps are done with O[n&1] (item n-2)
        tc_fence_after();
        if (leader) {
          const uint64_t v_desc = desc_adv(v_desc0, (g % tcp::VST) * TILE_BYTES);
#pragma unroll
          for (int k = 0; k < BN / 16; ++k) {
            // A = P in TMEM, written by the softmax warps over the first 64 columns of S[g&1]
            umma_bf16_ts(TM_O + (n & 1) * 128, TM_S + (g & 1) * BN + k * 8, desc_adv(v_desc, k * 2048), idesc_pv,
                         (j | k) != 0);
          }
          umma_commit(&pv_done[g & 1]);
          umma_commit(&v_empty[g % tcp::VST]);
          if (j + 1 == cur.ntiles) umma_commit(&o_ready[n & 1]);  // O of this work item is complete
        }
      }
      cur = nxt;
      have = have_next;
      ++n;
    }
  }
  } else if (warp < 12) {
    setmaxnreg_inc<168>();
    // 8 softmax warps: warp pair (w, w+4) shares the 32 TMEM lanes of quarter w&3 and splits the 128 key columns
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = quarter * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    float* xchg = reinterpret_cast<float*>(smem + tcp::OFF_X);  // [2 parity][2 half][128 rows] + [2][128]
    const bool win = CAUSAL && p.window > 0;
    FwdItem it;
    int g = 0;
    for (int n = 0; fwd_item<CAUSAL>(p, persist_item(n), nqt, it); ++n) {
      const int q0 = it.q0;
      float m_used = -INFINITY, l = 0.f;
      const uint32_t tm_o = TM_O + (n & 1) * 128;
      for (int j = 0; j < it.ntiles; ++j, ++g) {
        const int sb = g & 1;
        mbar_wait_spin(&s_full[sb], (g >> 1) & 1);
        tc_fence_after();
        uint32_t r[64];
        tmem_ld32(TM_S + lane_addr + sb * BN + half * 64, r);
        tmem_ld32(TM_S + lane_addr + sb * BN + half * 64 + 32, r + 32);
        tmem_ld_wait();
        const int jt0 = (j + it.jb) * BN;  // first key of this tile
        const int j0 = jt0 + half * 64;
        const bool need_mask = (jt0 + BN > p.sk) || (CAUSAL && (jt0 + BN - 1 > q0 + off)) ||
                               (win && jt0 < q0 + BM - 1 + off - p.window);
        if (need_mask) {
          const int lim = CAUSAL ? min(p.sk - 1, q0 + row + off) : p.sk - 1;  // last visible key
          const int vis = lim - j0;                                            // last visible column of this half
          if (!win) {  // one compare per score (every causal item has a diagonal tile: 12 % of the tiles at S = 2048)
#pragma unroll
            for (int c = 0; c < 64; ++c)
              if (c > vis) r[c] = 0xff800000u;  // -inf
          } else {
            const int lov = q0 + row + off - p.window - j0;  // first visible column
#pragma unroll
            for (int c = 0; c < 64; ++c)
              if (c > vis || c < lov) r[c] = 0xff800000u;
          }
        }
        float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
        for (int c = 0; c < 64; c += 4) {
#pragma unroll
          for (int e = 0; e < 4; ++e) mx4[e] = fmaxf(mx4[e], __uint_as_float(r[c + e]));
        }
        float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3]));
        xchg[(sb * 2 + half) * 128 + row] = mx;
        named_bar_sync(1 + quarter, 64);  // only the two warps of this lane quarter exchange (and overwrite each other's S columns with P)
        mx = fmaxf(mx, xchg[(sb * 2 + (half ^ 1)) * 128 + row]) * sl2;
        const bool upd = mx > m_used + 8.f;  // lazy rescale: keep the reference max unless it moved by > 2^8
        float alpha = 1.f;
        if (upd) {
          alpha = exp2f(m_used - mx);  // m_used = -inf → 0
          m_used = mx;
          l *= alpha;
        }
        const float mb = (m_used == -INFINITY) ? 0.f : m_used;
        float sum4[4] = {0.f, 0.f, 0.f, 0.f};
        // (Measured after this change and dropped, profiles/r02_attn_fwd_experiments.txt: the next tile's scores read
        // from TMEM and reduced to their maximum under this tile's exponentials — 0.42 ms, the wait for S of tile g+1
        // serialises the softmax behind P·V(g-1) + Q·Kᵀ(g+1) on the in-order tensor pipe; the two warps of a lane quarter
        // on ALTERNATE tiles with all 128 columns per thread — 0.342 ms, the per-tile chain S -> softmax -> P·V -> S(g+2)
        // gets longer with twice the columns per thread; every fourth exponential as a polynomial — 0.301 ms.)
        // sixteen columns at a time, each chunk packed and stored to TMEM before the next one's exponentials: with the
        // stores as ordering points ptxas mixes the FFMA / FADD / pack work into the MUFU stream.  (Written as one loop
        // of 64 exponentials followed by the packing, the SASS was 64 back-to-back MUFU.EX2 — 8 issue cycles each, the
        // FMA pipe idle — and both softmax warps of a scheduler run that phase at the same time: 0.330 -> 0.295 ms.)
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          uint32_t w[8];
#pragma unroll
          for (int e = 0; e < 16; e += 2) {
            const float p0 = ex2_approx(fmaf(__uint_as_float(r[ch * 16 + e]), sl2, -mb));
            const float p1 = ex2_approx(fmaf(__uint_as_float(r[ch * 16 + e + 1]), sl2, -mb));
            sum4[e & 2] += p0;
            sum4[(e & 2) + 1] += p1;
            w[e >> 1] = pack2(p0, p1);
          }
          tmem_st8(TM_S + lane_addr + sb * BN + half * 32 + ch * 8, w);
        }
        const float sum = (sum4[0] + sum4[1]) + (sum4[2] + sum4[3]);
        if (j > 0 && __any_sync(0xffffffffu, upd)) {
          mbar_wait_spin(&pv_done[(g - 1) & 1], ((g - 1) >> 1) & 1);  // P·V of the previous tile finished: O may be rescaled
          tc_fence_after();
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            if (half * 64 + c * 32 >= HD) break;  // head_dim 96 / 64: the second half owns fewer chunks
            uint32_t o[32];
            tmem_ld32(tm_o + lane_addr + half * 64 + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(tm_o + lane_addr + half * 64 + c * 32, o);
          }
        }
        tmem_st_wait();
        l += sum;
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[sb]);
      }
      // end of this work item: publish the half-row sums and the reference maximum for the epilogue warps and go on
      // with the next item.  The slot of item n-2 must have been read (o_free is arrived after that read).
      mbar_wait(&o_free[n & 1], ((n >> 1) & 1) ^ 1);
      float* lm = xchg + 512 + (n & 1) * 384;  // [l half 0 | l half 1 | m] x 128 rows
      lm[half * 128 + row] = l;
      if (half == 0) lm[256 + row] = m_used;
      __syncwarp();
      if (lane == 0) mbar_arrive(&l_ready[n & 1]);
    }
  } else {
    // epilogue warpgroup: one thread per query row; O / l → bf16 → global and the LSE of work item n while the
    // softmax warps are already on item n+1 (the per-item code was 23 % of the softmax warps' time — TMEM read of O,
    // 8 KB of global stores per warp, logf — in the ncu source view of the kernel that did it in line)
    setmaxnreg_dec<72>();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float* xchg = reinterpret_cast<const float*>(smem + tcp::OFF_X);
    FwdItem it;
    for (int n = 0; fwd_item<CAUSAL>(p, persist_item(n), nqt, it); ++n) {
      mbar_wait(&l_ready[n & 1], (n >> 1) & 1);
      const float* lm = xchg + 512 + (n & 1) * 384;
      const float l = lm[row] + lm[128 + row];
      const float m = lm[256 + row];
      mbar_wait(&o_ready[n & 1], (n >> 1) & 1);
      tc_fence_after();
      const uint32_t tm_o = TM_O + (n & 1) * 128 + lane_addr;
      const bool row_ok = it.q0 + row < p.sq;
      const float inv = l > 0.f ? 1.f / l : 0.f;
      bf16* orow = p.o + ((int64_t)it.b * p.sq + it.q0 + row) * p.ldo + it.h * HD;
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t o[32];
        tmem_ld32(tm_o + c * 32, o);
        tmem_ld_wait();
        if (c == HD / 32 - 1) {  // the accumulator (and the row sums) have been read: hand the buffers back
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_free[n & 1]);
        }
        if (row_ok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(o[q * 8 + i]) * inv;
            stg16(orow + c * 32 + q * 8, pack8(v));
          }
        }
      }
      if (p.lse && row_ok)
        p.lse[((int64_t)it.b * p.H + it.h) * p.sq + it.q0 + row] = l > 0.f ? m * 0.6931471805599453f + logf(l) : -INFINITY;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// bounded polling for the MMA issuers that watch several barriers (a broken invariant traps)
__device__ __forceinline__ void poll_guard(uint32_t& spins, bool did) {
  if (did) {
    spins = 0;
  } else if (++spins > (1u << 28)) {
    printf("[vpb] attention MMA issuer starved (block %d,%d,%d)\n", blockIdx.x, blockIdx.y, blockIdx.z);
    __trap();
  }
}

template <bool CAUSAL, int HD>
static int launch_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                         int64_t ldv, const AttnTcParams& p, cudaStream_t st) {
  CUtensorMap tmQ, tmK, tmV;
  if (make_tmap_2d(&tmQ, q, (uint64_t)p.H * HD, (uint64_t)p.B * p.sq, (uint64_t)ldq, 64, 128)) return -1;
  if (make_tmap_2d(&tmK, k, (uint64_t)p.KVH * HD, (uint64_t)p.B * p.sk, (uint64_t)ldk, 64, 128)) return -1;
  if (make_tmap_2d(&tmV, v, (uint64_t)p.KVH * HD, (uint64_t)p.B * p.sk, (uint64_t)ldv, 64, 128)) return -1;
  dim3 grid((p.sq + tc::BM - 1) / tc::BM, p.H, p.B);
  // default: persistent kernel (every work item needs at least one key tile: sk >= sq, the self-attention shapes);
  // VPB_OPT_ATTN_FWD_NS2 = 1 selects the one-work-item-per-CTA kernel
  static int n_sm = 0;
  if (n_sm == 0) {
    int dev = 0;
    VPB_CUDA(cudaGetDevice(&dev));
    VPB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  const int items = (int)grid.x * p.H * p.B;
  // with only a few work items per SM (the ViT towers: 640 items of 5 key tiles) the static round-robin of the
  // persistent kernel loses more to its ragged last round than it saves: 0.061 vs 0.053 ms at B=8, S=577
  const int fwd_mode = get_option(VPB_OPT_ATTN_FWD_NS2);  // 1: never persistent, 2: persistent whatever the item count (tests)
  if (fwd_mode != 1 && p.sk >= p.sq && p.sk > 0 && (fwd_mode == 2 || items >= 8 * n_sm)) {
    auto kernp = attn_fwd_tc_persist_kernel<CAUSAL, HD>;
    static bool cfgp = false;
    if (!cfgp) {
      VPB_CUDA(cudaFuncSetAttribute(kernp, cudaFuncAttributeMaxDynamicSharedMemorySize, tcp::SMEM_BYTES));
      cfgp = true;
    }
    kernp<<<n_sm, tcp::THREADS, tcp::SMEM_BYTES, st>>>(tmQ, tmK, tmV, p);
    VPB_LAUNCH_OK();
    return 0;
  }
  auto kern = attn_fwd_tc_kernel<CAUSAL, HD, 2>;
  static bool cfg = false;
  if (!cfg) {
    VPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, tc::SMEM_BYTES));
    cfg = true;
  }
  kern<<<grid, tc::THREADS, tc::SMEM_BYTES, st>>>(tmQ, tmK, tmV, p);
  VPB_LAUNCH_OK();
  return 0;
}

// =============================================================================================
// backward (head_dim 128)
// =============================================================================================
static long long* g_trace_buffer = nullptr;  // device buffer of >= 16*512 int64, or null
void set_trace_buffer(void* p) { g_trace_buffer = static_cast<long long*>(p); }

struct AttnTcBwdParams {
  const float* lse;    // [B,H,sq] natural log
  const float* delta;  // [B,H,sq] rowsum(dO*O)
  bf16 *dq, *dk, *dv;
  int64_t lddq, lddk, lddv;
  int B, H, KVH, sq, sk;
  float scale;
  int window;  // see AttnTcParams
  long long* trace;  // optional timeline buffer (vpb_set_trace_buffer): clock64 stamps of CTA 0
  // optional fused INVERSE rotary embedding of dQ / dK in the epilogues (head_dim 128, sq == sk):
  // cos/sin tables [max_pos, 64], positions per row or null (position = row index in the sequence)
  const float* rope_cos;
  const float* rope_sin;
  const int* rope_pos;
};

// cos / sin of one row's 32 rotary angles (fp32, as vpb_rope_table wrote them).  Loaded BEFORE the
// wait on the last MMA so the L2 latency of the table rows (one row per lane: nothing coalesces)
// hides behind the tail of the tensor work.
struct RopeRow {
  float4 c[8], s[8];
};
__device__ __forceinline__ void load_rope_row(RopeRow& r, const float* cs, const float* sn) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    r.c[i] = __ldg(reinterpret_cast<const float4*>(cs) + i);
    r.s[i] = __ldg(reinterpret_cast<const float4*>(sn) + i);
  }
}

// One 32-column chunk pair (x = columns [32g, 32g+32), y = the same + 64) of a gradient row:
// scale, bf16-round like the unfused store, rotate by the INVERSE angle (the backward of HF
// apply_rotary_pos_emb), store.  Same expressions as rope_kernel (elementwise.cu) → same bits.
template <bool ROPE>
__device__ __forceinline__ void store_pair_inverse_rope(bf16* o1p, bf16* o2p, const uint32_t* xa,
                                                        const uint32_t* xb, float scale,
                                                        const RopeRow& r) {
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float a[8], b[8], o1[8], o2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      a[i] = __uint_as_float(xa[q * 8 + i]) * scale;
      b[i] = __uint_as_float(xb[q * 8 + i]) * scale;
    }
    const uint4 pa = pack8(a), pb = pack8(b);
    if constexpr (ROPE) {
      unpack8(pa, a);
      unpack8(pb, b);
      const float4 c0 = r.c[2 * q], c1 = r.c[2 * q + 1], s0 = r.s[2 * q], s1 = r.s[2 * q + 1];
      const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
      const float sv[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float ss = -1.f * sv[i];
        o1[i] = __fmaf_rn(a[i], cc[i], -__fmul_rn(b[i], ss));
        o2[i] = __fmaf_rn(b[i], cc[i], __fmul_rn(a[i], ss));
      }
      stg16(o1p + q * 8, pack8(o1));
      stg16(o2p + q * 8, pack8(o2));
    } else {
      stg16(o1p + q * 8, pa);
      stg16(o2p + q * 8, pb);
    }
  }
}

// timeline instrumentation of the v2 backward kernels: row `slot`, column `it` (512 columns per row)
#define VPB_TRACE(slot, it)                                                        \
  do {                                                                             \
    if (p.trace && lin_cta == 0 && (it) < 512) p.trace[(slot) * 512 + (it)] = clock64(); \
  } while (0)

namespace tcb {
constexpr int HD = 128;
constexpr float LOG2E = 1.4426950408889634f;
// ---- dK/dV kernel: 128-key tile per CTA, 64-query tiles streamed ----
constexpr int A_BKV = 128, A_BQ = 64, A_ST = 3;
constexpr int A_OFF_K = 0, A_OFF_V = 32768;
constexpr int A_OFF_QD = 65536;              // stage s: Q (16 KB) then dO (16 KB)
constexpr int A_OFF_P = A_OFF_QD + A_ST * 32768;   // P^T  [128 keys x 64 queries] 16 KB
constexpr int A_OFF_DS = A_OFF_P + 16384;          // dS^T 16 KB
constexpr int A_OFF_LD = A_OFF_DS + 16384;         // [2][2][64] floats: lse*log2e | delta
constexpr int A_OFF_BAR = A_OFF_LD + 1024;
constexpr int A_SMEM = A_OFF_BAR + 256 + 1024;
// ---- dQ kernel: 128-query tile per CTA, 64-key tiles streamed ----
constexpr int B_BQ = 128, B_BKV = 64, B_ST = 3;
constexpr int B_OFF_Q = 0, B_OFF_DO = 32768;
constexpr int B_OFF_KV = 65536;              // stage s: K (16 KB) then V (16 KB)
constexpr int B_OFF_DS = B_OFF_KV + B_ST * 32768;  // dS [128 queries x 64 keys] 16 KB
constexpr int B_OFF_BAR = B_OFF_DS + 16384;
constexpr int B_SMEM = B_OFF_BAR + 256 + 1024;
}  // namespace tcb

// dK_j = scale * sum_i dS_ij^T Q_i ,  dV_j = sum_i P_ij^T dO_i   (sum over the GQA group's heads too)
template <bool CAUSAL>
__global__ void __launch_bounds__(320, 1)
attn_bwd_dkdv_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                        const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                        const AttnTcBwdParams p) {
  using namespace tcb;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_OFF_BAR);
  uint64_t* kv_full = bars + 0;
  uint64_t* qd_full = bars + 1;    // [3]
  uint64_t* qd_empty = bars + 4;   // [3]
  uint64_t* sd_full = bars + 7;    // [2]
  uint64_t* pds_full = bars + 9;
  uint64_t* pds_empty = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);
  float* ld_buf = reinterpret_cast<float*>(smem + A_OFF_LD);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int kv0 = blockIdx.x * A_BKV;
  const int kvh = blockIdx.y, b = blockIdx.z;
  const int G = p.H / p.KVH;
  const int off = p.sk - p.sq;
  const int nq_tiles = (p.sq + A_BQ - 1) / A_BQ;
  int qt_begin = 0;
  if (CAUSAL) {
    int first = kv0 - off;
    if (first < 0) first = 0;
    qt_begin = first / A_BQ;
    if (qt_begin > nq_tiles) qt_begin = nq_tiles;
  }
  const int nper = nq_tiles - qt_begin;
  const int nit = G * nper;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(kv_full, 1);
    for (int i = 0; i < A_ST; ++i) {
      mbar_init(&qd_full[i], 1);
      mbar_init(&qd_empty[i], 1);
    }
    mbar_init(&sd_full[0], 1);
    mbar_init(&sd_full[1], 1);
    mbar_init(pds_full, 8);
    mbar_init(pds_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t TM_S = tmem_base;         // S^T[2]  : 2 x 64 columns
  const uint32_t TM_DP = tmem_base + 128;  // dP^T[2] : 2 x 64 columns
  const uint32_t TM_DV = tmem_base + 256;  // 128 columns
  const uint32_t TM_DK = tmem_base + 384;  // 128 columns

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 65536);
      const int krow = b * p.sk + kv0;
      tma_load_2d(smem + A_OFF_K, &tmK, kv_full, kvh * HD, krow);
      tma_load_2d(smem + A_OFF_K + 16384, &tmK, kv_full, kvh * HD + 64, krow);
      tma_load_2d(smem + A_OFF_V, &tmV, kv_full, kvh * HD, krow);
      tma_load_2d(smem + A_OFF_V + 16384, &tmV, kv_full, kvh * HD + 64, krow);
      for (int it = 0; it < nit; ++it) {
        const int st = it % A_ST;
        const int hq = kvh * G + it / nper;
        const int qrow = b * p.sq + (qt_begin + it % nper) * A_BQ;
        mbar_wait(&qd_empty[st], ((it / A_ST) & 1) ^ 1);
        mbar_arrive_expect_tx(&qd_full[st], 32768);
        uint8_t* sq_ = smem + A_OFF_QD + st * 32768;
        tma_load_2d(sq_, &tmQ, &qd_full[st], hq * HD, qrow);
        tma_load_2d(sq_ + 8192, &tmQ, &qd_full[st], hq * HD + 64, qrow);
        tma_load_2d(sq_ + 16384, &tmDO, &qd_full[st], hq * HD, qrow);
        tma_load_2d(sq_ + 24576, &tmDO, &qd_full[st], hq * HD + 64, qrow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nit > 0) {
      constexpr uint32_t idesc_sd = make_idesc_bf16(128, A_BQ, 0, 0);   // [128 keys x 64 queries]
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, HD, 0, 1);    // A K-major, B MN-major
      const uint32_t k_addr = smem_u32(smem + A_OFF_K), v_addr = smem_u32(smem + A_OFF_V);
      const uint32_t p_addr = smem_u32(smem + A_OFF_P), ds_addr = smem_u32(smem + A_OFF_DS);
      auto issue_sd = [&](int it) {
        const int st = it % A_ST, sb = it & 1;
        mbar_wait(&qd_full[st], (it / A_ST) & 1);
        tc_fence_after();
        const uint32_t q_addr = smem_u32(smem + A_OFF_QD + st * 32768);
        const uint32_t do_addr = q_addr + 16384;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t oa = (k >> 2) * 16384 + (k & 3) * 32;  // K / V tiles: 128-row chunks
          const uint32_t ob = (k >> 2) * 8192 + (k & 3) * 32;   // Q / dO tiles: 64-row chunks
          umma_bf16(TM_S + sb * A_BQ, make_smem_desc(k_addr + oa, 16, 1024),
                    make_smem_desc(q_addr + ob, 16, 1024), idesc_sd, k != 0);
        }
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t oa = (k >> 2) * 16384 + (k & 3) * 32;
          const uint32_t ob = (k >> 2) * 8192 + (k & 3) * 32;
          umma_bf16(TM_DP + sb * A_BQ, make_smem_desc(v_addr + oa, 16, 1024),
                    make_smem_desc(do_addr + ob, 16, 1024), idesc_sd, k != 0);
        }
        umma_commit(&sd_full[sb]);
      };
      mbar_wait(kv_full, 0);
      issue_sd(0);
      for (int it = 0; it < nit; ++it) {
        if (it + 1 < nit) issue_sd(it + 1);
        mbar_wait(pds_full, it & 1);
        tc_fence_after();
        const int st = it % A_ST;
        const uint32_t q_addr = smem_u32(smem + A_OFF_QD + st * 32768);
        const uint32_t do_addr = q_addr + 16384;
#pragma unroll
        for (int k = 0; k < A_BQ / 16; ++k) {  // contraction over the 64 queries
          umma_bf16(TM_DV, make_smem_desc(p_addr + k * 32, 16, 1024),
                    make_smem_desc(do_addr + k * 2048, 8192, 1024), idesc_acc, (it | k) != 0);
        }
#pragma unroll
        for (int k = 0; k < A_BQ / 16; ++k) {
          umma_bf16(TM_DK, make_smem_desc(ds_addr + k * 32, 16, 1024),
                    make_smem_desc(q_addr + k * 2048, 8192, 1024), idesc_acc, (it | k) != 0);
        }
        umma_commit(pds_empty);
        umma_commit(&qd_empty[st]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;     // which 32 of the 64 query columns this warp handles
    const int row = quarter * 32 + lane;  // key row within the tile
    const int tid = threadIdx.x - 64;     // 0..255 among the softmax warps
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    const bool key_ok = kv0 + row < p.sk;
    uint8_t* prow = smem + A_OFF_P + row * 128;
    uint8_t* dsrow = smem + A_OFF_DS + row * 128;
    for (int it = 0; it < nit; ++it) {
      const int sb = it & 1;
      const int hq = kvh * G + it / nper;
      const int q0 = (qt_begin + it % nper) * A_BQ;
      // stage lse*log2e / delta of the 64 queries (double buffered by iteration parity)
      if (tid < 128) {
        float* buf = ld_buf + sb * 128;
        const int c = tid & 63;
        const bool ok = q0 + c < p.sq;
        const int64_t li = ((int64_t)b * p.H + hq) * p.sq + q0 + c;
        if (tid < 64) buf[c] = ok ? p.lse[li] * LOG2E : 0.f;
        else buf[64 + c] = ok ? p.delta[li] : 0.f;
      }
      mbar_wait(&sd_full[sb], (it >> 1) & 1);
      tc_fence_after();
      uint32_t s[32], d[32];
      tmem_ld32(TM_S + lane_addr + sb * A_BQ + half * 32, s);
      tmem_ld32(TM_DP + lane_addr + sb * A_BQ + half * 32, d);
      named_bar_sync(1, 256);  // lse/delta staged
      tmem_ld_wait();
      const float* lbuf = ld_buf + sb * 128;
      const bool need_mask = (q0 + A_BQ > p.sq) || !key_ok ||
                             (CAUSAL && (kv0 + A_BKV - 1 > q0 + off));
      const float4* l4 = reinterpret_cast<const float4*>(lbuf) + half * 8;
#pragma unroll
      for (int c4 = 0; c4 < 8; ++c4) {
        const float4 ls = l4[c4], dl4 = l4[16 + c4];
        const float lsv[4] = {ls.x, ls.y, ls.z, ls.w}, dlv[4] = {dl4.x, dl4.y, dl4.z, dl4.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          const int c = c4 * 4 + e;
          s[c] = __float_as_uint(ex2_approx(fmaf(__uint_as_float(s[c]), sl2, -lsv[e])));
          d[c] = __float_as_uint(__uint_as_float(d[c]) - dlv[e]);
        }
      }
      if (need_mask) {  // ONE warp-uniform branch per tile (a per-element branch costs 3 control
                        // instructions + a branch-resolve stall per score: r01 ncu source view)
        // visible queries are qc >= first_q; everything is masked for an out-of-range key row
        const int first_q = key_ok ? (CAUSAL ? kv0 + row - off : 0) : 0x7fffffff;
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          const int qc = q0 + half * 32 + c;
          if (qc < first_q || qc >= p.sq) s[c] = 0u;
        }
      }
#pragma unroll
      for (int c = 0; c < 32; ++c) d[c] = __float_as_uint(__uint_as_float(s[c]) * __uint_as_float(d[c]));
      if (it > 0) {
        mbar_wait(pds_empty, (it - 1) & 1);  // previous dV/dK MMAs finished reading P^T / dS^T
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float a[8], g[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          a[i] = __uint_as_float(s[u * 8 + i]);
          g[i] = __uint_as_float(d[u * 8 + i]);
        }
        const int so = ((half * 4 + u) ^ (row & 7)) << 4;
        *reinterpret_cast<uint4*>(prow + so) = pack8(a);
        *reinterpret_cast<uint4*>(dsrow + so) = pack8(g);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(pds_full);
    }
    // epilogue: each of the four warps sharing a lane quarter writes 32 of the 128 head-dim columns
    if (nit > 0) {
      mbar_wait(pds_empty, (nit - 1) & 1);
      tc_fence_after();
    }
    bf16* dkrow = p.dk + ((int64_t)b * p.sk + kv0 + row) * p.lddk + kvh * HD;
    bf16* dvrow = p.dv + ((int64_t)b * p.sk + kv0 + row) * p.lddv + kvh * HD;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c = half * 2 + cc;
      uint32_t a[32], g[32];
      if (nit > 0) {
        tmem_ld32(TM_DK + lane_addr + c * 32, a);
        tmem_ld32(TM_DV + lane_addr + c * 32, g);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = g[i] = 0;
      }
      if (key_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float x[8], y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            x[i] = __uint_as_float(a[q * 8 + i]) * p.scale;
            y[i] = __uint_as_float(g[q * 8 + i]);
          }
          stg16(dkrow + c * 32 + q * 8, pack8(x));
          stg16(dvrow + c * 32 + q * 8, pack8(y));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// dQ_i = scale * sum_j dS_ij K_j
template <bool CAUSAL>
__global__ void __launch_bounds__(320, 1)
attn_bwd_dq_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                      const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                      const AttnTcBwdParams p) {
  using namespace tcb;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF_BAR);
  uint64_t* q_full = bars + 0;
  uint64_t* kv_full = bars + 1;   // [3]
  uint64_t* kv_empty = bars + 4;  // [3]
  uint64_t* sd_full = bars + 7;   // [2]
  uint64_t* ds_full = bars + 9;
  uint64_t* ds_empty = bars + 10;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = CAUSAL ? (gridDim.x - 1 - blockIdx.x) : blockIdx.x;
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.H / p.KVH);
  const int q0 = qt * B_BQ;
  const int off = p.sk - p.sq;
  int kv_end = p.sk;
  if (CAUSAL) {
    kv_end = q0 + B_BQ + off;
    if (kv_end > p.sk) kv_end = p.sk;
    if (kv_end < 0) kv_end = 0;
  }
  const int nit = (kv_end + B_BKV - 1) / B_BKV;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    for (int i = 0; i < B_ST; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    mbar_init(&sd_full[0], 1);
    mbar_init(&sd_full[1], 1);
    mbar_init(ds_full, 8);
    mbar_init(ds_empty, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t TM_S = tmem_base;         // S[2]  : 2 x 64
  const uint32_t TM_DP = tmem_base + 128;  // dP[2] : 2 x 64
  const uint32_t TM_DQ = tmem_base + 256;  // 128

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, 65536);
      const int qrow = b * p.sq + q0;
      tma_load_2d(smem + B_OFF_Q, &tmQ, q_full, h * HD, qrow);
      tma_load_2d(smem + B_OFF_Q + 16384, &tmQ, q_full, h * HD + 64, qrow);
      tma_load_2d(smem + B_OFF_DO, &tmDO, q_full, h * HD, qrow);
      tma_load_2d(smem + B_OFF_DO + 16384, &tmDO, q_full, h * HD + 64, qrow);
      for (int it = 0; it < nit; ++it) {
        const int st = it % B_ST;
        const int krow = b * p.sk + it * B_BKV;
        mbar_wait(&kv_empty[st], ((it / B_ST) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[st], 32768);
        uint8_t* sk_ = smem + B_OFF_KV + st * 32768;
        tma_load_2d(sk_, &tmK, &kv_full[st], kvh * HD, krow);
        tma_load_2d(sk_ + 8192, &tmK, &kv_full[st], kvh * HD + 64, krow);
        tma_load_2d(sk_ + 16384, &tmV, &kv_full[st], kvh * HD, krow);
        tma_load_2d(sk_ + 24576, &tmV, &kv_full[st], kvh * HD + 64, krow);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && nit > 0) {
      constexpr uint32_t idesc_sd = make_idesc_bf16(128, B_BKV, 0, 0);  // [128 queries x 64 keys]
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, HD, 0, 1);
      const uint32_t q_addr = smem_u32(smem + B_OFF_Q), do_addr = smem_u32(smem + B_OFF_DO);
      const uint32_t ds_addr = smem_u32(smem + B_OFF_DS);
      auto issue_sd = [&](int it) {
        const int st = it % B_ST, sb = it & 1;
        mbar_wait(&kv_full[st], (it / B_ST) & 1);
        tc_fence_after();
        const uint32_t k_addr = smem_u32(smem + B_OFF_KV + st * 32768);
        const uint32_t v_addr = k_addr + 16384;
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t oa = (k >> 2) * 16384 + (k & 3) * 32;
          const uint32_t ob = (k >> 2) * 8192 + (k & 3) * 32;
          umma_bf16(TM_S + sb * B_BKV, make_smem_desc(q_addr + oa, 16, 1024),
                    make_smem_desc(k_addr + ob, 16, 1024), idesc_sd, k != 0);
        }
#pragma unroll
        for (int k = 0; k < HD / 16; ++k) {
          const uint32_t oa = (k >> 2) * 16384 + (k & 3) * 32;
          const uint32_t ob = (k >> 2) * 8192 + (k & 3) * 32;
          umma_bf16(TM_DP + sb * B_BKV, make_smem_desc(do_addr + oa, 16, 1024),
                    make_smem_desc(v_addr + ob, 16, 1024), idesc_sd, k != 0);
        }
        umma_commit(&sd_full[sb]);
      };
      mbar_wait(q_full, 0);
      issue_sd(0);
      for (int it = 0; it < nit; ++it) {
        if (it + 1 < nit) issue_sd(it + 1);
        mbar_wait(ds_full, it & 1);
        tc_fence_after();
        const int st = it % B_ST;
        const uint32_t k_addr = smem_u32(smem + B_OFF_KV + st * 32768);
#pragma unroll
        for (int k = 0; k < B_BKV / 16; ++k) {  // contraction over the 64 keys
          umma_bf16(TM_DQ, make_smem_desc(ds_addr + k * 32, 16, 1024),
                    make_smem_desc(k_addr + k * 2048, 8192, 1024), idesc_dq, (it | k) != 0);
        }
        umma_commit(ds_empty);
        umma_commit(&kv_empty[st]);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;     // which 32 of the 64 key columns this warp handles
    const int row = quarter * 32 + lane;  // query row within the tile
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    const bool row_ok = q0 + row < p.sq;
    const int64_t li = ((int64_t)b * p.H + h) * p.sq + q0 + row;
    const float lse2 = row_ok ? p.lse[li] * LOG2E : 0.f;
    const float dl = row_ok ? p.delta[li] : 0.f;
    uint8_t* dsrow = smem + B_OFF_DS + row * 128;
    for (int it = 0; it < nit; ++it) {
      const int sb = it & 1;
      const int j0 = it * B_BKV;
      mbar_wait(&sd_full[sb], (it >> 1) & 1);
      tc_fence_after();
      uint32_t s[32], d[32];
      tmem_ld32(TM_S + lane_addr + sb * B_BKV + half * 32, s);
      tmem_ld32(TM_DP + lane_addr + sb * B_BKV + half * 32, d);
      tmem_ld_wait();
      const bool need_mask = !row_ok || (j0 + B_BKV > p.sk) || (CAUSAL && (j0 + B_BKV - 1 > q0 + off));
      const int lim = CAUSAL ? min(p.sk - 1, q0 + row + off) : p.sk - 1;
#pragma unroll
      for (int c = 0; c < 32; ++c) s[c] = __float_as_uint(ex2_approx(fmaf(__uint_as_float(s[c]), sl2, -lse2)));
      if (need_mask) {  // one warp-uniform branch, not one per score
        const int vis = row_ok ? lim - (j0 + half * 32) : -1;  // last visible column of this chunk
#pragma unroll
        for (int c = 0; c < 32; ++c)
          if (c > vis) s[c] = 0u;
      }
#pragma unroll
      for (int c = 0; c < 32; ++c)
        d[c] = __float_as_uint(__uint_as_float(s[c]) * (__uint_as_float(d[c]) - dl));
      if (it > 0) mbar_wait(ds_empty, (it - 1) & 1);
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float g[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] = __uint_as_float(d[u * 8 + i]);
        *reinterpret_cast<uint4*>(dsrow + (((half * 4 + u) ^ (row & 7)) << 4)) = pack8(g);
      }
      fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(ds_full);
    }
    if (nit > 0) {
      mbar_wait(ds_empty, (nit - 1) & 1);
      tc_fence_after();
    }
    bf16* dqrow = p.dq + ((int64_t)b * p.sq + q0 + row) * p.lddq + h * HD;
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c = half * 2 + cc;
      uint32_t a[32];
      if (nit > 0) {
        tmem_ld32(TM_DQ + lane_addr + c * 32, a);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = 0;
      }
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float x[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(a[q * 8 + i]) * p.scale;
          stg16(dqrow + c * 32 + q * 8, pack8(x));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// =============================================================================================
// backward v2: ping-pong softmax groups
// =============================================================================================
// The v1 kernels above run ONE softmax group per CTA, so every iteration is a serial chain
// (S/dP MMA → TMEM load → exp / dS → smem → dV/dK MMA) and the tensor pipe idles ~65 % of the time
// (profiles/r01_ncu_attn_tc_8warp.csv).  v2 splits the 8 softmax warps into two groups of 4 (one
// warp per TMEM lane quarter) that work on ALTERNATE iterations with their own S/dP TMEM buffers
// and their own P/dS shared-memory buffers, releases the S/dP buffers as soon as they are in
// registers (s_free), and lets the single MMA thread POLL (mbarrier.test_wait) for whichever of
// "next S/dP" / "next accumulate" is ready instead of issuing in a fixed order.  lse/delta of a
// query tile are prefetched one iteration ahead.  The dQ kernel additionally pairs a heavy and a
// light causal query tile in one CTA (balanced work, half the prologues) with dQ double-buffered
// in TMEM.
namespace tcb2 {  // head_dim (128 or 96) is a template parameter of the kernels
constexpr float LOG2E = 1.4426950408889634f;
// TS mode (default): P^T / dS^T are written back over their own scores in TMEM and feed the
// accumulate MMAs as TMEM A operands (tcgen05.mma with [tmem] A) — no P/dS shared-memory tiles, no
// generic→async proxy fence, and the 64 KB they occupied become two more TMA stages (5 instead of 3).
// SS mode (VPB_OPT_ATTN_BWD_SS): P^T / dS^T staged through shared memory as K-major A tiles.
constexpr int A_BKV = 128, A_BQ = 64, A_ST_SS = 3, A_ST_TS = 5;
constexpr int A_OFF_K = 0, A_OFF_V = 32768;
constexpr int A_OFF_QD = 65536;                      // stage s: Q (16 KB) then dO (16 KB)
constexpr int A_OFF_PDS = A_OFF_QD + A_ST_SS * 32768;  // SS mode, group g: P^T (16 KB) then dS^T (16 KB)
constexpr int A_OFF_LD = A_OFF_PDS + 2 * 32768;      // [2 groups][2 parity][64 lse | 64 delta] floats
constexpr int A_OFF_BAR = A_OFF_LD + 2048;
constexpr int A_SMEM = A_OFF_BAR + 256;              // 231 680 B: needs the 1024-aligned base
constexpr int B_BQ = 128, B_BKV = 64, B_ST_SS = 4, B_ST_TS = 5;
constexpr int B_OFF_Q = 0, B_OFF_DO = 32768;
constexpr int B_OFF_KV = 65536;                      // stage s: K (16 KB) then V (16 KB)
constexpr int B_OFF_DS = B_OFF_KV + B_ST_SS * 32768; // SS mode, group g: dS [128 queries x 64 keys] 16 KB
constexpr int B_OFF_BAR = B_OFF_DS + 2 * 16384;
constexpr int B_SMEM = B_OFF_BAR + 256;
}  // namespace tcb2

// SPLIT (TS only): instead of two groups on alternate iterations, all 8 softmax warps work on EVERY
// iteration, two per TMEM lane quarter splitting the 64 query columns — the exp / dS phase of one
// group turned out to be the serial bottleneck (one warp per sub-partition issues ~0.26 IPC; the
// timeline in profiles/r01_attn_dkdv_timeline_cta0.txt shows the two groups' phases do not overlap).
template <bool CAUSAL, int HD, bool TS, bool SPLIT>
__global__ void __launch_bounds__(320, 1)
attn_bwd_dkdv_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                         const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                         const AttnTcBwdParams p) {
  using namespace tcb2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + A_OFF_BAR);
  constexpr int A_ST = TS ? A_ST_TS : A_ST_SS;
  uint64_t* kv_full = bars + 0;
  uint64_t* qd_full = bars + 1;     // [A_ST <= 5]
  uint64_t* qd_empty = bars + 6;    // [A_ST <= 5]
  uint64_t* sd_full = bars + 11;    // [2] S^T/dP^T of buffer b complete
  uint64_t* s_free = bars + 13;     // [2] SS mode: group b has its S^T/dP^T in registers
  uint64_t* pds_full = bars + 15;   // [2] group b wrote P^T/dS^T
  uint64_t* pds_empty = bars + 17;  // [2] SS mode: dV/dK MMAs reading smem buffer b retired
  uint64_t* all_done = bars + 19;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 20);
  uint64_t* ld_full = bars + 22;    // [3] SPLIT mode: lse / delta of an iteration staged in ld_buf[it % 3]
  float* ld_buf = reinterpret_cast<float*>(smem + A_OFF_LD);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // CTAs are dispatched in linear blockIdx order; with causal masking the early key tiles carry the
  // most work, so the linear id is re-read as (key tile major, (kv head, batch) minor): every SM
  // starts on heavy tiles and the tail of the grid is made of the lightest ones.
  const int lin = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
  const int lin_cta = lin;
  const int nhb = gridDim.y * gridDim.z;
  const int kv0 = (lin / nhb) * A_BKV;
  const int kvh = (lin % nhb) % gridDim.y, b = (lin % nhb) / gridDim.y;
  const int G = p.H / p.KVH;
  const int off = p.sk - p.sq;
  const int nq_tiles = (p.sq + A_BQ - 1) / A_BQ;
  int qt_begin = 0;
  if (CAUSAL) {
    int first = kv0 - off;
    if (first < 0) first = 0;
    qt_begin = first / A_BQ;
    if (qt_begin > nq_tiles) qt_begin = nq_tiles;
  }
  int qt_end = nq_tiles;
  const bool win = CAUSAL && p.window > 0;
  if (win) {  // last query row that still sees this tile's last key
    const int last = kv0 + A_BKV - 1 - off + p.window;
    if (last / A_BQ + 1 < qt_end) qt_end = last / A_BQ + 1;
    if (qt_end < qt_begin) qt_end = qt_begin;
  }
  const int nper = qt_end - qt_begin;
  const int nit = G * nper;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023) {
      printf("[vpb] dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(kv_full, 1);
    for (int i = 0; i < A_ST; ++i) {
      mbar_init(&qd_full[i], 1);
      mbar_init(&qd_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sd_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&pds_full[i], SPLIT ? 8 : 4);
      mbar_init(&pds_empty[i], 1);
    }
    mbar_init(all_done, 1);
    for (int i = 0; i < 3; ++i) mbar_init(&ld_full[i], 4);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t TM_S = tmem_base;         // S^T[2]  : 2 x 64 columns
  const uint32_t TM_DP = tmem_base + 128;  // dP^T[2] : 2 x 64 columns
  const uint32_t TM_DV = tmem_base + 256;  // 128 columns
  const uint32_t TM_DK = tmem_base + 384;  // 128 columns

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 65536);
      const int krow = b * p.sk + kv0;
      tma_load_2d(smem + A_OFF_K, &tmK, kv_full, kvh * HD, krow);
      tma_load_2d(smem + A_OFF_K + 16384, &tmK, kv_full, kvh * HD + 64, krow);
      tma_load_2d(smem + A_OFF_V, &tmV, kv_full, kvh * HD, krow);
      tma_load_2d(smem + A_OFF_V + 16384, &tmV, kv_full, kvh * HD + 64, krow);
      int hq = kvh * G, qi = 0, st = 0, ph = 1;
      for (int it = 0; it < nit; ++it) {
        const int qrow = b * p.sq + (qt_begin + qi) * A_BQ;
        mbar_wait(&qd_empty[st], ph);
        VPB_TRACE(0, it);  // producer: stage free, loads issued
        mbar_arrive_expect_tx(&qd_full[st], 32768);
        uint8_t* sq_ = smem + A_OFF_QD + st * 32768;
        tma_load_2d(sq_, &tmQ, &qd_full[st], hq * HD, qrow);
        tma_load_2d(sq_ + 8192, &tmQ, &qd_full[st], hq * HD + 64, qrow);
        tma_load_2d(sq_ + 16384, &tmDO, &qd_full[st], hq * HD, qrow);
        tma_load_2d(sq_ + 24576, &tmDO, &qd_full[st], hq * HD + 64, qrow);
        if (++qi == nper) {
          qi = 0;
          ++hq;
        }
        if (++st == A_ST) {
          st = 0;
          ph ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // The whole warp runs this loop (warp-uniform control flow keeps the descriptor arithmetic on
    // the uniform datapath — the issuing thread's own instruction stream was the limiter when the
    // descriptors were rebuilt per MMA inside an `if (lane == 0)` region); one elected lane issues.
    if (nit > 0) {
      const bool leader = elect_one();
      constexpr uint32_t idesc_sd = make_idesc_bf16(128, A_BQ, 0, 0);   // [128 keys x 64 queries]
      constexpr uint32_t idesc_acc = make_idesc_bf16(128, HD, 0, 1);    // A K-major, B MN-major
      const uint64_t k_desc = make_smem_desc(smem_u32(smem + A_OFF_K), 16, 1024);
      const uint64_t v_desc = make_smem_desc(smem_u32(smem + A_OFF_V), 16, 1024);
      const uint64_t q_desc0 = make_smem_desc(smem_u32(smem + A_OFF_QD), 16, 1024);       // K-major view
      const uint64_t q_mn0 = make_smem_desc(smem_u32(smem + A_OFF_QD), 8192, 1024);      // MN-major view
      const uint64_t p_desc0 = make_smem_desc(smem_u32(smem + A_OFF_PDS), 16, 1024);
      mbar_wait(kv_full, 0);
      int n_sd = 0, n_acc = 0;
      uint32_t spins = 0;
      while (n_acc < nit) {
        bool did = false;
        if (n_sd < nit) {
          const int sb = n_sd & 1, st = n_sd % A_ST;
          // buffer sb is free once iteration n_sd-2 is done with it: SS — its scores are in the
          // group's registers (s_free); TS — its P^T/dS^T were consumed, i.e. the dV/dK MMAs of
          // n_sd-2 were ISSUED (tcgen05.mma of one thread execute in order)
          bool ok = (n_sd < 2) || (TS ? (n_acc >= n_sd - 1) : mbar_test(&s_free[sb], ((n_sd >> 1) - 1) & 1));
          ok = ok && mbar_test(&qd_full[st], (n_sd / A_ST) & 1);
          if (ok) {
            tc_fence_after();
            if (leader) VPB_TRACE(1, n_sd);  // MMA: S/dP issue
            if (leader) {
              const uint64_t q_desc = desc_adv(q_desc0, st * 32768);
              const uint64_t do_desc = desc_adv(q_desc, 16384);
#pragma unroll
              for (int k = 0; k < HD / 16; ++k) {
                const uint32_t oa = (k >> 2) * 16384 + (k & 3) * 32;  // K / V tiles: 128-row chunks
                const uint32_t ob = (k >> 2) * 8192 + (k & 3) * 32;   // Q / dO tiles: 64-row chunks
                umma_bf16(TM_S + sb * A_BQ, desc_adv(k_desc, oa), desc_adv(q_desc, ob), idesc_sd, k != 0);
              }
#pragma unroll
              for (int k = 0; k < HD / 16; ++k) {
                const uint32_t oa = (k >> 2) * 16384 + (k & 3) * 32;
                const uint32_t ob = (k >> 2) * 8192 + (k & 3) * 32;
                umma_bf16(TM_DP + sb * A_BQ, desc_adv(v_desc, oa), desc_adv(do_desc, ob), idesc_sd, k != 0);
              }
              umma_commit(&sd_full[sb]);
            }
            ++n_sd;
            did = true;
          }
        }
        if (n_acc < n_sd && mbar_test(&pds_full[n_acc & 1], (n_acc >> 1) & 1)) {
          tc_fence_after();
          const int st = n_acc % A_ST;
          if (leader) VPB_TRACE(2, n_acc);  // MMA: dV/dK issue
          if (leader) {
            const uint64_t q_mn = desc_adv(q_mn0, st * 32768);
            const uint64_t do_mn = desc_adv(q_mn, 16384);
            const uint64_t p_desc = desc_adv(p_desc0, (n_acc & 1) * 32768);
            const uint64_t ds_desc = desc_adv(p_desc, 16384);
            if constexpr (TS) {
              const uint32_t gcol = (n_acc & 1) * A_BQ;  // P^T over S^T[g], dS^T over dP^T[g]: 32 packed columns
#pragma unroll
              for (int k = 0; k < A_BQ / 16; ++k) {  // contraction over the 64 queries
                // packed columns of k-step k: contiguous (k*8), or per column-half in SPLIT mode
                const uint32_t ac = SPLIT ? (k >> 1) * 32 + (k & 1) * 8 : k * 8;
                umma_bf16_ts(TM_DV, TM_S + gcol + ac, desc_adv(do_mn, k * 2048), idesc_acc, (n_acc | k) != 0);
              }
#pragma unroll
              for (int k = 0; k < A_BQ / 16; ++k) {
                const uint32_t ac = SPLIT ? (k >> 1) * 32 + (k & 1) * 8 : k * 8;
                umma_bf16_ts(TM_DK, TM_DP + gcol + ac, desc_adv(q_mn, k * 2048), idesc_acc, (n_acc | k) != 0);
              }
            } else {
#pragma unroll
              for (int k = 0; k < A_BQ / 16; ++k) {  // contraction over the 64 queries
                umma_bf16(TM_DV, desc_adv(p_desc, k * 32), desc_adv(do_mn, k * 2048), idesc_acc,
                          (n_acc | k) != 0);
              }
#pragma unroll
              for (int k = 0; k < A_BQ / 16; ++k) {
                umma_bf16(TM_DK, desc_adv(ds_desc, k * 32), desc_adv(q_mn, k * 2048), idesc_acc,
                          (n_acc | k) != 0);
              }
            }
            umma_commit(&pds_empty[n_acc & 1]);
            umma_commit(&qd_empty[st]);
          }
          ++n_acc;
          did = true;
        }
        poll_guard(spins, did);
      }
      if (leader) umma_commit(all_done);
    }
  } else {
    const int quarter = warp & 3;
    const int g = (warp - 2) >> 2;        // softmax group: iterations it ≡ g (mod 2)
    const int row = quarter * 32 + lane;  // key row within the tile == TMEM lane
    const int gtid = threadIdx.x - 64 - g * 128;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    const bool key_ok = kv0 + row < p.sk;
    uint8_t* prow = smem + A_OFF_PDS + g * 32768 + row * 128;
    uint8_t* dsrow = prow + 16384;
    float* lbase = ld_buf + g * 256;
    auto fetch = [&](int hq, int qi) -> float {  // this thread's share of the tile's lse*log2e | delta
      const int c = gtid & 63;
      const int q = (qt_begin + qi) * A_BQ + c;
      if (q >= p.sq) return 0.f;
      const int64_t li = ((int64_t)b * p.H + hq) * p.sq + q;
      return gtid < 64 ? p.lse[li] * LOG2E : p.delta[li];
    };
    auto advance = [&](int& hq, int& qi, int steps) {  // (head, query tile) `steps` iterations later
      for (int s_ = 0; s_ < steps; ++s_)
        if (++qi == nper) {
          qi = 0;
          ++hq;
        }
    };
    if constexpr (SPLIT) {
      static_assert(!SPLIT || TS, "SPLIT needs P/dS in TMEM");
      const int half = g;                   // which 32 of the 64 query columns this warp handles
      const int tid = threadIdx.x - 64;     // 0..255 among the softmax warps
      // (head, query tile) of an iteration are tracked incrementally: the div/mod by the runtime
      // tile count cost ~600 dependent cycles at the top of every iteration (timeline trace)
      auto fetch8 = [&](int hq, int qi) -> float {  // threads 0..127 stage lse*log2e | delta of the tile
        const int q = (qt_begin + qi) * A_BQ + (tid & 63);
        if (tid >= 128 || q >= p.sq) return 0.f;
        const int64_t li = ((int64_t)b * p.H + hq) * p.sq + q;
        return tid < 64 ? p.lse[li] : p.delta[li];  // raw: a multiply here would make the warp wait for the load at once
      };
      // lse / delta staging without a block barrier.  Round 1 staged iteration it+1 at the end of iteration it behind
      // a 256-thread bar.sync and fetched it+2 into a register meanwhile; the ncu source view
      // (profiles/r02_ncu_attn_baseline.csv) put 19 % of the softmax warps' samples on that barrier and another 8 %
      // on the global load arriving late.  Now: three staging buffers, the four writer warps arrive on ld_full[it % 3]
      // after their stores (readers wait on it — written an iteration earlier, so the wait is free), and two fetches
      // are in flight (iterations it+2 and it+3).  Buffer (it+1) % 3 is rewritten at the end of iteration it: its last
      // readers were in iteration it-2, and S/dP of iteration it — which this warp has seen complete — is only issued
      // after every warp handed over P/dS of it-2.
      int hq_n = kvh * G, qi_n = 0;  // (head, query tile) of the iteration whose values are fetched next
      auto step = [&]() {
        if (++qi_n == nper) {
          qi_n = 0;
          ++hq_n;
        }
      };
      auto publish = [&](int it_, float val) {  // writer warps (tid < 128 = the first four softmax warps)
        if (tid < 128) {
          ld_buf[(it_ % 3) * 128 + tid] = tid < 64 ? val * LOG2E : val;  // lse in log2 units | delta
          __syncwarp();
          if (lane == 0) mbar_arrive(&ld_full[it_ % 3]);
        }
      };
      int qi_c = 0;  // query tile of the current iteration
      float pre_a = 0.f, pre_b = 0.f;
      if (nit > 0) {
        publish(0, fetch8(hq_n, qi_n));
        step();
      }
      if (nit > 1) pre_a = fetch8(hq_n, qi_n);  // iteration 1
      step();
      if (nit > 2) pre_b = fetch8(hq_n, qi_n);  // iteration 2
      step();
      for (int it = 0; it < nit; ++it) {
        const int sb = it & 1;
        float* lbuf = ld_buf + (it % 3) * 128;
        const int q0 = (qt_begin + qi_c) * A_BQ;
        if (++qi_c == nper) qi_c = 0;
        // tile-level, hence warp-uniform: the two bodies below hold .sync.aligned TMEM stores (a per-row test such as
        // !key_ok would split the warp between them)
        const bool need_mask = (q0 + A_BQ > p.sq) || (kv0 + A_BKV > p.sk) ||
                               (CAUSAL && (kv0 + A_BKV - 1 > q0 + off)) ||
                               (win && (q0 + A_BQ - 1 + off - p.window > kv0));
        if (quarter == 0 && lane == 0 && half == 0) VPB_TRACE(3, it);
        mbar_wait_spin(&sd_full[sb], (it >> 1) & 1);
        mbar_wait_spin(&ld_full[it % 3], (it / 3) & 1);
        if (quarter == 0 && lane == 0 && half == 0) VPB_TRACE(4, it);
        tc_fence_after();
        uint32_t s[32], d[32];
        tmem_ld32(TM_S + lane_addr + sb * A_BQ + half * 32, s);
        tmem_ld32(TM_DP + lane_addr + sb * A_BQ + half * 32, d);
        tmem_ld_wait();
        if (quarter == 0 && lane == 0 && half == 0) VPB_TRACE(5, it);
        const float4* l4 = reinterpret_cast<const float4*>(lbuf) + half * 8;
        // Two straight-line chunks of sixteen query columns, each stored to TMEM as soon as it is packed: ptxas then
        // mixes the FFMA / FADD / FMUL / pack / LDS work into the MUFU stream (one 32-long MUFU burst per warp, with both
        // warps of a scheduler in that phase at the same time, left the other pipes idle).  MASK is a compile-time
        // copy of the body: a branch inside it would fence the two kinds of work into separate blocks again.
        auto body = [&](auto mask_c) {
          constexpr bool MASK = decltype(mask_c)::value;
          const int first_q = key_ok ? (CAUSAL ? kv0 + row - off : 0) : 0x7fffffff;
          const int last_q = win ? min(p.sq - 1, kv0 + row - off + p.window) : p.sq - 1;
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t wp[8], wd[8];
#pragma unroll
            for (int c4 = 0; c4 < 4; ++c4) {
              const float4 ls = l4[ch * 4 + c4], dl4 = l4[16 + ch * 4 + c4];
              const float lsv[4] = {ls.x, ls.y, ls.z, ls.w}, dlv[4] = {dl4.x, dl4.y, dl4.z, dl4.w};
              float pv[4], dv[4];
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int c = ch * 16 + c4 * 4 + e;
                pv[e] = ex2_approx(fmaf(__uint_as_float(s[c]), sl2, -lsv[e]));
                if (MASK) {
                  const int qc = q0 + half * 32 + c;
                  if (qc < first_q || qc > last_q) pv[e] = 0.f;
                }
                dv[e] = pv[e] * (__uint_as_float(d[c]) - dlv[e]);
              }
              wp[c4 * 2] = pack2(pv[0], pv[1]);
              wp[c4 * 2 + 1] = pack2(pv[2], pv[3]);
              wd[c4 * 2] = pack2(dv[0], dv[1]);
              wd[c4 * 2 + 1] = pack2(dv[2], dv[3]);
            }
            // bf16 pairs back over this warp's OWN score columns: packed columns [32*half, 32*half+16)
            tmem_st8(TM_S + lane_addr + sb * A_BQ + half * 32 + ch * 8, wp);
            tmem_st8(TM_DP + lane_addr + sb * A_BQ + half * 32 + ch * 8, wd);
          }
        };
        if (need_mask) body(std::true_type{});  // one warp-uniform branch per tile, never one per score
        else body(std::false_type{});
        if (quarter == 0 && lane == 0 && half == 0) VPB_TRACE(6, it);
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pds_full[sb]);
        if (quarter == 0 && lane == 0 && half == 0) VPB_TRACE(7, it);
        // stage iteration it+1, keep it+2 in a register, fetch it+3
        if (it + 1 < nit) publish(it + 1, pre_a);
        pre_a = pre_b;
        if (it + 3 < nit) pre_b = fetch8(hq_n, qi_n);
        step();
      }
    } else {
    int hq_c = kvh * G, qi_c = 0;
    if (nper > 0) advance(hq_c, qi_c, g);
    float pre = (g < nit) ? fetch(hq_c, qi_c) : 0.f;
    int k = 0;
    for (int it = g; it < nit; it += 2, ++k) {
      float* lbuf = lbase + (k & 1) * 128;
      lbuf[gtid] = pre;
      const int q0 = (qt_begin + qi_c) * A_BQ;
      advance(hq_c, qi_c, 2);
      if (it + 2 < nit) pre = fetch(hq_c, qi_c);  // in flight during this iteration
      const bool need_mask = (q0 + A_BQ > p.sq) || !key_ok ||
                             (CAUSAL && (kv0 + A_BKV - 1 > q0 + off)) ||
                             (win && (q0 + A_BQ - 1 + off - p.window > kv0));
      if (quarter == 0 && lane == 0) VPB_TRACE(3, it);  // softmax: ready to wait for S/dP
      mbar_wait_spin(&sd_full[g], k & 1);
      if (quarter == 0 && lane == 0) VPB_TRACE(4, it);  // softmax: S/dP complete seen
      tc_fence_after();
      // all 64 S^T / dP^T columns go to registers first, so the buffer is handed back to the MMA
      // issuer (S/dP of iteration it+2) before any of the exp / dS work starts
      uint32_t sall[64], dall[64];
      tmem_ld32(TM_S + lane_addr + g * A_BQ, sall);
      tmem_ld32(TM_S + lane_addr + g * A_BQ + 32, sall + 32);
      tmem_ld32(TM_DP + lane_addr + g * A_BQ, dall);
      tmem_ld32(TM_DP + lane_addr + g * A_BQ + 32, dall + 32);
      named_bar_sync(1 + g, 128);  // lse/delta staged — under the TMEM load latency
      tmem_ld_wait();
      if (quarter == 0 && lane == 0) VPB_TRACE(5, it);  // softmax: scores in registers
      if constexpr (!TS) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[g]);
      }
#pragma unroll
      for (int hc = 0; hc < 2; ++hc) {  // two chunks of 32 query columns
        uint32_t* s = sall + hc * 32;
        uint32_t* d = dall + hc * 32;
        const float4* l4 = reinterpret_cast<const float4*>(lbuf) + hc * 8;
#pragma unroll
        for (int c4 = 0; c4 < 8; ++c4) {
          const float4 ls = l4[c4], dl4 = l4[16 + c4];
          const float lsv[4] = {ls.x, ls.y, ls.z, ls.w}, dlv[4] = {dl4.x, dl4.y, dl4.z, dl4.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c4 * 4 + e;
            s[c] = __float_as_uint(ex2_approx(fmaf(__uint_as_float(s[c]), sl2, -lsv[e])));
            d[c] = __float_as_uint(__uint_as_float(d[c]) - dlv[e]);
          }
        }
        if (need_mask) {  // one warp-uniform branch per chunk, never one per score
          const int first_q = key_ok ? (CAUSAL ? kv0 + row - off : 0) : 0x7fffffff;
          const int last_q = win ? min(p.sq - 1, kv0 + row - off + p.window) : p.sq - 1;
#pragma unroll
          for (int c = 0; c < 32; ++c) {
            const int qc = q0 + hc * 32 + c;
            if (qc < first_q || qc > last_q) s[c] = 0u;
          }
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) d[c] = __float_as_uint(__uint_as_float(s[c]) * __uint_as_float(d[c]));
        if constexpr (TS) {
          // bf16 pairs back into TMEM over this thread's own scores: columns [16*hc, 16*hc+16)
          uint32_t wp[16], wd[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            wp[i] = pack2(__uint_as_float(s[2 * i]), __uint_as_float(s[2 * i + 1]));
            wd[i] = pack2(__uint_as_float(d[2 * i]), __uint_as_float(d[2 * i + 1]));
          }
          tmem_st16(TM_S + lane_addr + g * A_BQ + hc * 16, wp);
          tmem_st16(TM_DP + lane_addr + g * A_BQ + hc * 16, wd);
        } else {
          if (hc == 0 && k > 0) mbar_wait(&pds_empty[g], (k - 1) & 1);  // dV/dK of it-2 read this buffer
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float a[8], gg[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              a[i] = __uint_as_float(s[u * 8 + i]);
              gg[i] = __uint_as_float(d[u * 8 + i]);
            }
            const int so = ((hc * 4 + u) ^ (row & 7)) << 4;
            *reinterpret_cast<uint4*>(prow + so) = pack8(a);
            *reinterpret_cast<uint4*>(dsrow + so) = pack8(gg);
          }
        }
      }
      if (quarter == 0 && lane == 0) VPB_TRACE(6, it);  // softmax: P/dS computed and stored (issued)
      if constexpr (TS) tmem_st_wait();
      else fence_proxy_async();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&pds_full[g]);
      if (quarter == 0 && lane == 0) VPB_TRACE(7, it);  // softmax: arrived
    }
    }
    [[maybe_unused]] RopeRow rr;
    if constexpr (HD == 128) {
      if (p.rope_cos && key_ok) {
        const int m = b * p.sk + kv0 + row;
        const int pos = p.rope_pos ? p.rope_pos[m] : kv0 + row;
        load_rope_row(rr, p.rope_cos + (int64_t)pos * 64 + g * 32, p.rope_sin + (int64_t)pos * 64 + g * 32);
      }
    }
    if (nit > 0) {
      mbar_wait(all_done, 0);
      tc_fence_after();
    }
    bf16* dkrow = p.dk + ((int64_t)b * p.sk + kv0 + row) * p.lddk + kvh * HD;
    bf16* dvrow = p.dv + ((int64_t)b * p.sk + kv0 + row) * p.lddv + kvh * HD;
    if constexpr (HD == 128) {
      // warp half g owns the column pair [32g, 32g+32) | [64+32g, ...): inverse RoPE of dK in place
#pragma unroll 1
      for (int which = 0; which < 2; ++which) {  // 0: dK (scaled, rotated back), 1: dV
        uint32_t a[32], a2[32];
        if (nit > 0) {
          const uint32_t tm = (which ? TM_DV : TM_DK) + lane_addr + g * 32;
          tmem_ld32(tm, a);
          tmem_ld32(tm + 64, a2);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int i = 0; i < 32; ++i) a[i] = a2[i] = 0;
        }
        if (key_ok) {
          bf16* orow = which ? dvrow : dkrow;
          if (which == 0 && p.rope_cos)
            store_pair_inverse_rope<true>(orow + g * 32, orow + 64 + g * 32, a, a2, p.scale, rr);
          else
            store_pair_inverse_rope<false>(orow + g * 32, orow + 64 + g * 32, a, a2, which ? 1.f : p.scale, rr);
        }
      }
    } else {
#pragma unroll 1
    for (int cc = 0; cc < 2; ++cc) {
      const int c = g * 2 + cc;
      if (c * 32 >= HD) break;  // head_dim 96: three 32-column chunks
      uint32_t a[32], gg[32];
      if (nit > 0) {
        tmem_ld32(TM_DK + lane_addr + c * 32, a);
        tmem_ld32(TM_DV + lane_addr + c * 32, gg);
        tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) a[i] = gg[i] = 0;
      }
      if (key_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float x[8], y[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            x[i] = __uint_as_float(a[q * 8 + i]) * p.scale;
            y[i] = __uint_as_float(gg[q * 8 + i]);
          }
          stg16(dkrow + c * 32 + q * 8, pack8(x));
          stg16(dvrow + c * 32 + q * 8, pack8(y));
        }
      }
    }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// dQ_i = scale * sum_j dS_ij K_j — two query tiles (heavy + light) per CTA, ping-pong groups
template <bool CAUSAL, int HD, bool TS, bool SPLIT>
__global__ void __launch_bounds__(320, 1)
attn_bwd_dq_tc2_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                       const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                       const AttnTcBwdParams p) {
  using namespace tcb2;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + B_OFF_BAR);
  constexpr int B_ST = TS ? B_ST_TS : B_ST_SS;
  uint64_t* q_full = bars + 0;
  uint64_t* q_empty = bars + 1;
  uint64_t* kv_full = bars + 2;    // [B_ST <= 5]
  uint64_t* kv_empty = bars + 7;   // [B_ST <= 5]
  uint64_t* sd_full = bars + 12;   // [2]
  uint64_t* s_free = bars + 14;    // [2] SS mode
  uint64_t* ds_full = bars + 16;   // [2]
  uint64_t* ds_empty = bars + 18;  // [2] SS mode
  uint64_t* tile_done = bars + 20; // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int h = blockIdx.y, b = blockIdx.z;
  const int kvh = h / (p.H / p.KVH);
  const int off = p.sk - p.sq;
  const int nqt = (p.sq + B_BQ - 1) / B_BQ;
  // tile 0 = the heavier (later) query tile, tile 1 = its mirror image
  int qt[2] = {nqt - 1 - (int)blockIdx.x, (int)blockIdx.x};
  const int ntl = (qt[0] != qt[1]) ? 2 : 1;
  int nit[2] = {0, 0};
  int jb[2] = {0, 0};  // first key tile a query tile can see (sliding window)
  const bool win = CAUSAL && p.window > 0;
  for (int t = 0; t < ntl; ++t) {
    int kv_end = p.sk;
    if (CAUSAL) {
      kv_end = qt[t] * B_BQ + B_BQ + off;
      if (kv_end > p.sk) kv_end = p.sk;
    }
    if (win) {
      const int lo = qt[t] * B_BQ + off - p.window;
      jb[t] = lo > 0 ? lo / B_BKV : 0;
    }
    nit[t] = (kv_end + B_BKV - 1) / B_BKV - jb[t];  // >= 1: the launcher guarantees sk >= sq for causal
  }
  const int total = nit[0] + nit[1];

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023) {
      printf("[vpb] dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int i = 0; i < B_ST; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&sd_full[i], 1);
      mbar_init(&s_free[i], 4);
      mbar_init(&ds_full[i], SPLIT ? 8 : 4);
      mbar_init(&ds_empty[i], 1);
      mbar_init(&tile_done[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t TM_S = tmem_base;         // S[2]  : 2 x 64
  const uint32_t TM_DP = tmem_base + 128;  // dP[2] : 2 x 64
  const uint32_t TM_DQ = tmem_base + 256;  // dQ[2] : 2 x 128 (one per query tile)

  if (warp == 0) {
    if (lane == 0) {
      int n = 0;
      for (int t = 0; t < ntl; ++t) {
        if (t > 0) mbar_wait(q_empty, 0);  // every S/dP MMA of tile 0 retired: Q/dO smem is free
        mbar_arrive_expect_tx(q_full, 65536);
        const int qrow = b * p.sq + qt[t] * B_BQ;
        tma_load_2d(smem + B_OFF_Q, &tmQ, q_full, h * HD, qrow);
        tma_load_2d(smem + B_OFF_Q + 16384, &tmQ, q_full, h * HD + 64, qrow);
        tma_load_2d(smem + B_OFF_DO, &tmDO, q_full, h * HD, qrow);
        tma_load_2d(smem + B_OFF_DO + 16384, &tmDO, q_full, h * HD + 64, qrow);
        for (int it = 0; it < nit[t]; ++it, ++n) {
          const int st = n % B_ST;
          const int krow = b * p.sk + (it + jb[t]) * B_BKV;
          mbar_wait(&kv_empty[st], ((n / B_ST) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[st], 32768);
          uint8_t* sk_ = smem + B_OFF_KV + st * 32768;
          tma_load_2d(sk_, &tmK, &kv_full[st], kvh * HD, krow);
          tma_load_2d(sk_ + 8192, &tmK, &kv_full[st], kvh * HD + 64, krow);
          tma_load_2d(sk_ + 16384, &tmV, &kv_full[st], kvh * HD, krow);
          tma_load_2d(sk_ + 24576, &tmV, &kv_full[st], kvh * HD + 64, krow);
        }
      }
    }
  } else if (warp == 1) {
    {  // warp-uniform issue loop, one elected lane issues (see the dK/dV kernel)
      const bool leader = elect_one();
      constexpr uint32_t idesc_sd = make_idesc_bf16(128, B_BKV, 0, 0);  // [128 queries x 64 keys]
      constexpr uint32_t idesc_dq = make_idesc_bf16(128, HD, 0, 1);
      const uint64_t q_desc = make_smem_desc(smem_u32(smem + B_OFF_Q), 16, 1024);
      const uint64_t do_desc = make_smem_desc(smem_u32(smem + B_OFF_DO), 16, 1024);
      const uint64_t kv_desc0 = make_smem_desc(smem_u32(smem + B_OFF_KV), 16, 1024);    // K-major view
      const uint64_t kv_mn0 = make_smem_desc(smem_u32(smem + B_OFF_KV), 8192, 1024);   // MN-major view
      const uint64_t ds_desc0 = make_smem_desc(smem_u32(smem + B_OFF_DS), 16, 1024);
      int n_sd = 0, n_dq = 0;
      uint32_t spins = 0;
      while (n_dq < total) {
        bool did = false;
        if (n_sd < total) {
          const int t = n_sd < nit[0] ? 0 : 1;
          const int i = n_sd - (t ? nit[0] : 0);
          const int sb = n_sd & 1, st = n_sd % B_ST;
          // see the dK/dV kernel: TS frees a score buffer when the dQ MMAs of n_sd-2 are issued
          bool ok = (n_sd < 2) || (TS ? (n_dq >= n_sd - 1) : mbar_test(&s_free[sb], ((n_sd >> 1) - 1) & 1));
          ok = ok && mbar_test(&kv_full[st], (n_sd / B_ST) & 1);
          if (ok && i == 0) ok = mbar_test(q_full, t);
          if (ok) {
            tc_fence_after();
            if (leader) {
              const uint64_t k_desc = desc_adv(kv_desc0, st * 32768);
              const uint64_t v_desc = desc_adv(k_desc, 16384);
#pragma unroll
              for (int k = 0; k < HD / 16; ++k) {
                const uint32_t oa = (k >> 2) * 16384 + (k & 3) * 32;
                const uint32_t ob = (k >> 2) * 8192 + (k & 3) * 32;
                umma_bf16(TM_S + sb * B_BKV, desc_adv(q_desc, oa), desc_adv(k_desc, ob), idesc_sd, k != 0);
              }
#pragma unroll
              for (int k = 0; k < HD / 16; ++k) {
                const uint32_t oa = (k >> 2) * 16384 + (k & 3) * 32;
                const uint32_t ob = (k >> 2) * 8192 + (k & 3) * 32;
                umma_bf16(TM_DP + sb * B_BKV, desc_adv(do_desc, oa), desc_adv(v_desc, ob), idesc_sd, k != 0);
              }
              umma_commit(&sd_full[sb]);
              if (i == nit[t] - 1) umma_commit(q_empty);  // last reader of this tile's Q / dO
            }
            ++n_sd;
            did = true;
          }
        }
        if (n_dq < n_sd && mbar_test(&ds_full[n_dq & 1], (n_dq >> 1) & 1)) {
          tc_fence_after();
          const int t = n_dq < nit[0] ? 0 : 1;
          const int i = n_dq - (t ? nit[0] : 0);
          const int st = n_dq % B_ST;
          if (leader) {
            const uint64_t k_mn = desc_adv(kv_mn0, st * 32768);
            const uint64_t ds_desc = desc_adv(ds_desc0, (n_dq & 1) * 16384);
#pragma unroll
            for (int k = 0; k < B_BKV / 16; ++k) {  // contraction over the 64 keys
              if constexpr (TS)
                umma_bf16_ts(TM_DQ + t * 128,
                             TM_DP + (n_dq & 1) * B_BKV + (SPLIT ? (k >> 1) * 32 + (k & 1) * 8 : k * 8),
                             desc_adv(k_mn, k * 2048), idesc_dq, (i | k) != 0);
              else
                umma_bf16(TM_DQ + t * 128, desc_adv(ds_desc, k * 32), desc_adv(k_mn, k * 2048), idesc_dq,
                          (i | k) != 0);
            }
            umma_commit(&ds_empty[n_dq & 1]);
            umma_commit(&kv_empty[st]);
            if (i == nit[t] - 1) umma_commit(&tile_done[t]);
          }
          ++n_dq;
          did = true;
        }
        poll_guard(spins, did);
      }
    }
  } else {
    const int quarter = warp & 3;
    const int g = (warp - 2) >> 2;        // softmax group: running iterations n ≡ g (mod 2)
    const int row = quarter * 32 + lane;  // query row within a tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    float lse2[2] = {0.f, 0.f}, dl[2] = {0.f, 0.f};
    bool row_ok[2] = {false, false};
    for (int t = 0; t < ntl; ++t) {
      row_ok[t] = qt[t] * B_BQ + row < p.sq;
      if (row_ok[t]) {
        const int64_t li = ((int64_t)b * p.H + h) * p.sq + qt[t] * B_BQ + row;
        lse2[t] = p.lse[li] * LOG2E;
        dl[t] = p.delta[li];
      }
    }
    uint8_t* dsrow = smem + B_OFF_DS + g * 16384 + row * 128;
    if constexpr (SPLIT) {
      static_assert(!SPLIT || TS, "SPLIT needs dS in TMEM");
      const int half = g;  // which 32 of the 64 key columns this warp handles
      for (int n = 0; n < total; ++n) {
        const int sb = n & 1;
        const int t = n < nit[0] ? 0 : 1;
        const int it = n - (t ? nit[0] : 0);
        const int q0 = qt[t] * B_BQ;
        const int j0 = (it + (t ? jb[1] : jb[0])) * B_BKV;
        const bool rok = t ? row_ok[1] : row_ok[0];
        const float l2 = t ? lse2[1] : lse2[0];
        const float dlt = t ? dl[1] : dl[0];
        const bool need_mask = !rok || (j0 + B_BKV > p.sk) || (CAUSAL && (j0 + B_BKV - 1 > q0 + off)) ||
                               (win && j0 < q0 + B_BQ - 1 + off - p.window);
        mbar_wait_spin(&sd_full[sb], (n >> 1) & 1);
        tc_fence_after();
        uint32_t s[32], d[32];
        tmem_ld32(TM_S + lane_addr + sb * B_BKV + half * 32, s);
        tmem_ld32(TM_DP + lane_addr + sb * B_BKV + half * 32, d);
        tmem_ld_wait();
#pragma unroll
        for (int c = 0; c < 32; ++c) {
          s[c] = __float_as_uint(ex2_approx(fmaf(__uint_as_float(s[c]), sl2, -l2)));
        }
        if (need_mask) {  // one warp-uniform branch per tile, never one per score
          const int lim = CAUSAL ? min(p.sk - 1, q0 + row + off) : p.sk - 1;
          const int vis = rok ? lim - (j0 + half * 32) : -1;                      // last visible column
          const int lov = win ? q0 + row + off - p.window - (j0 + half * 32) : 0;  // first visible column
#pragma unroll
          for (int c = 0; c < 32; ++c)
            if (c > vis || c < lov) s[c] = 0u;
        }
        uint32_t wd[16];
#pragma unroll
        for (int i = 0; i < 16; ++i)
          wd[i] = pack2(__uint_as_float(s[2 * i]) * (__uint_as_float(d[2 * i]) - dlt),
                        __uint_as_float(s[2 * i + 1]) * (__uint_as_float(d[2 * i + 1]) - dlt));
        tmem_st16(TM_DP + lane_addr + sb * B_BKV + half * 32, wd);  // over this warp's own dP columns
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ds_full[sb]);
      }
    } else {
    int k = 0;
    for (int n = g; n < total; n += 2, ++k) {
      const int t = n < nit[0] ? 0 : 1;
      const int it = n - (t ? nit[0] : 0);
      const int q0 = qt[t] * B_BQ;
      const int j0 = (it + (t ? jb[1] : jb[0])) * B_BKV;
      const bool rok = t ? row_ok[1] : row_ok[0];
      const float l2 = t ? lse2[1] : lse2[0];
      const float dlt = t ? dl[1] : dl[0];
      const bool need_mask = !rok || (j0 + B_BKV > p.sk) || (CAUSAL && (j0 + B_BKV - 1 > q0 + off)) ||
                             (win && j0 < q0 + B_BQ - 1 + off - p.window);
      const int lim = CAUSAL ? min(p.sk - 1, q0 + row + off) : p.sk - 1;
      const int lov = win ? q0 + row + off - p.window - j0 : 0;  // first visible column of this tile
      mbar_wait_spin(&sd_full[g], k & 1);
      tc_fence_after();
      // all 64 S / dP columns go to registers first, so the buffer is handed back to the MMA
      // issuer (S/dP of iteration n+2) before any of the exp / dS work starts
      uint32_t s[64], d[64];
      tmem_ld32(TM_S + lane_addr + g * B_BKV, s);
      tmem_ld32(TM_S + lane_addr + g * B_BKV + 32, s + 32);
      tmem_ld32(TM_DP + lane_addr + g * B_BKV, d);
      tmem_ld32(TM_DP + lane_addr + g * B_BKV + 32, d + 32);
      tmem_ld_wait();
      if constexpr (!TS) {
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[g]);
      }
#pragma unroll
      for (int c = 0; c < 64; ++c) s[c] = __float_as_uint(ex2_approx(fmaf(__uint_as_float(s[c]), sl2, -l2)));
      if (need_mask) {  // one warp-uniform branch per tile, never one per score
        const int vis = rok ? lim - j0 : -1;  // last visible column of this tile
#pragma unroll
        for (int c = 0; c < 64; ++c)
          if (c > vis || c < lov) s[c] = 0u;
      }
#pragma unroll
      for (int c = 0; c < 64; ++c)
        d[c] = __float_as_uint(__uint_as_float(s[c]) * (__uint_as_float(d[c]) - dlt));
      if constexpr (TS) {
        uint32_t wd[32];  // dS as bf16 pairs over this thread's own dP scores in TMEM
#pragma unroll
        for (int i = 0; i < 32; ++i) wd[i] = pack2(__uint_as_float(d[2 * i]), __uint_as_float(d[2 * i + 1]));
        tmem_st32(TM_DP + lane_addr + g * B_BKV, wd);
        tmem_st_wait();
      } else {
        if (k > 0) mbar_wait(&ds_empty[g], (k - 1) & 1);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          float gg[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) gg[i] = __uint_as_float(d[u * 8 + i]);
          *reinterpret_cast<uint4*>(dsrow + ((u ^ (row & 7)) << 4)) = pack8(gg);
        }
        fence_proxy_async();
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&ds_full[g]);
    }
    }
    // epilogue: both query tiles' dQ sit in TMEM; the 8 warps split the 128 head-dim columns
    for (int t = 0; t < ntl; ++t) {
      const bool rok = t ? row_ok[1] : row_ok[0];
      [[maybe_unused]] RopeRow rr;
      if constexpr (HD == 128) {
        if (p.rope_cos && rok) {
          const int m = b * p.sq + qt[t] * B_BQ + row;
          const int pos = p.rope_pos ? p.rope_pos[m] : qt[t] * B_BQ + row;
          load_rope_row(rr, p.rope_cos + (int64_t)pos * 64 + g * 32, p.rope_sin + (int64_t)pos * 64 + g * 32);
        }
      }
      mbar_wait(&tile_done[t], 0);
      tc_fence_after();
      bf16* dqrow = p.dq + ((int64_t)b * p.sq + qt[t] * B_BQ + row) * p.lddq + h * HD;
      if constexpr (HD == 128) {
        // warp half g owns the column pair [32g, 32g+32) | [64+32g, 64+32g+32): the two halves of a
        // rotary pair sit in one thread, so the inverse RoPE of dQ can happen here
        uint32_t a[32], bb[32];
        tmem_ld32(TM_DQ + t * 128 + lane_addr + g * 32, a);
        tmem_ld32(TM_DQ + t * 128 + lane_addr + 64 + g * 32, bb);
        tmem_ld_wait();
        if (rok) {
          if (p.rope_cos) store_pair_inverse_rope<true>(dqrow + g * 32, dqrow + 64 + g * 32, a, bb, p.scale, rr);
          else store_pair_inverse_rope<false>(dqrow + g * 32, dqrow + 64 + g * 32, a, bb, p.scale, rr);
        }
      } else {
#pragma unroll 1
        for (int cc = 0; cc < 2; ++cc) {
          const int c = g * 2 + cc;
          if (c * 32 >= HD) break;
          uint32_t a[32];
          tmem_ld32(TM_DQ + t * 128 + lane_addr + c * 32, a);
          tmem_ld_wait();
          if (rok) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float x[8];
#pragma unroll
              for (int i = 0; i < 8; ++i) x[i] = __uint_as_float(a[q * 8 + i]) * p.scale;
              stg16(dqrow + c * 32 + q * 8, pack8(x));
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

// =============================================================================================
// dQ, persistent (default for self-attention shapes without the fused inverse RoPE)
// =============================================================================================
// The same two measures that took the forward from 0.430 to 0.330 ms (profiles/r02_attn_fwd_experiments.txt), applied
// to the dQ kernel above, whose 2048 CTAs at B=8,H=32,S=2048 (13.8 per SM) each pay the per-CTA fixed cost, reload
// Q / dO between their two query tiles with the tensor pipe drained, and write 64 KB of dQ from the softmax warps:
// one CTA per SM walks a heavy-first list of 128-query work items; Q / dO are double-buffered in shared memory and
// requested one item ahead, the K/V ring never drains, lse / delta of the next item are prefetched into registers,
// dQ is double-buffered in TMEM and written out by a separate EPILOGUE warpgroup while the softmax warps and the
// tensor pipe are already on the next item.  Same MMAs, same order: bit-identical to the kernel above.
namespace tcq {
constexpr int BQ = 128, BKV = 64, ST = 3;
constexpr int OFF_QD = 0;                       // [2] x (Q 32 KB | dO 32 KB)
constexpr int OFF_KV = 2 * 65536;               // stage s: K (16 KB) then V (16 KB)
constexpr int OFF_BAR = OFF_KV + ST * 32768;    // 224 KB of tiles
constexpr int SMEM = OFF_BAR + 512;
constexpr int THREADS = 512;
constexpr float LOG2E = 1.4426950408889634f;
}  // namespace tcq

struct DqItem {
  int h, b, kvh, q0, jb, nit;
};

template <bool CAUSAL>
__device__ __forceinline__ bool dq_item(const AttnTcBwdParams& p, int w, int nqt, DqItem& it) {
  const int per = p.H * p.B;
  if (w >= nqt * per) return false;
  const int qi = w / per, r = w - qi * per;
  const int qt = CAUSAL ? nqt - 1 - qi : qi;  // heavy query tiles first
  it.h = r % p.H;
  it.b = r / p.H;
  it.kvh = it.h / (p.H / p.KVH);
  it.q0 = qt * tcq::BQ;
  const int off = p.sk - p.sq;
  int kv_end = p.sk;
  if (CAUSAL) {
    kv_end = it.q0 + tcq::BQ + off;
    if (kv_end > p.sk) kv_end = p.sk;
  }
  it.jb = 0;
  if (CAUSAL && p.window > 0) {
    const int lo = it.q0 + off - p.window;
    it.jb = lo > 0 ? lo / tcq::BKV : 0;
  }
  it.nit = (kv_end + tcq::BKV - 1) / tcq::BKV - it.jb;  // >= 1: the launcher guarantees sk >= sq
  return true;
}

template <bool CAUSAL, int HD>
__global__ void __launch_bounds__(512, 1)
attn_bwd_dq_persist_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmDO,
                           const __grid_constant__ CUtensorMap tmK, const __grid_constant__ CUtensorMap tmV,
                           const AttnTcBwdParams p) {
  using namespace tcq;
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* q_full = bars + 0;     // [2] Q / dO of item n landed in buffer n&1
  uint64_t* q_empty = bars + 2;    // [2] every S / dP MMA of the item in that buffer has completed
  uint64_t* kv_full = bars + 4;    // [ST]
  uint64_t* kv_empty = bars + 7;   // [ST]
  uint64_t* sd_full = bars + 10;   // [2] S / dP of iteration x (buffer x&1) complete
  uint64_t* ds_full = bars + 12;   // [2] the softmax warps wrote dS of iteration x
  uint64_t* dq_ready = bars + 14;  // [2] every dQ MMA of item n has completed
  uint64_t* dq_free = bars + 16;   // [2] the epilogue warps have read dQ[n&1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int off = p.sk - p.sq;
  const int nqt = (p.sq + BQ - 1) / BQ;
  const bool win = CAUSAL && p.window > 0;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023) {
      printf("[vpb] dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmQ);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmK);
    tma_prefetch_desc(&tmV);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
      mbar_init(&sd_full[i], 1);
      mbar_init(&ds_full[i], 8);
      mbar_init(&dq_ready[i], 1);
      mbar_init(&dq_free[i], 4);
    }
    for (int i = 0; i < ST; ++i) {
      mbar_init(&kv_full[i], 1);
      mbar_init(&kv_empty[i], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t TM_S = tmem_base;         // S[2]  : 2 x 64
  const uint32_t TM_DP = tmem_base + 128;  // dP[2] : 2 x 64
  const uint32_t TM_DQ = tmem_base + 256;  // dQ[2] : 2 x 128 (item n in half n&1)

  if (warp < 4) {
  setmaxnreg_dec<96>();
  if (warp == 0) {
    if (lane == 0) {
      auto load_q = [&](int n, const DqItem& it) {
        const int s_ = n & 1;
        mbar_wait(&q_empty[s_], ((n >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[s_], 65536);
        uint8_t* sq_ = smem + OFF_QD + s_ * 65536;
        const int qrow = it.b * p.sq + it.q0;
        tma_load_2d(sq_, &tmQ, &q_full[s_], it.h * HD, qrow);
        tma_load_2d(sq_ + 16384, &tmQ, &q_full[s_], it.h * HD + 64, qrow);
        tma_load_2d(sq_ + 32768, &tmDO, &q_full[s_], it.h * HD, qrow);
        tma_load_2d(sq_ + 49152, &tmDO, &q_full[s_], it.h * HD + 64, qrow);
      };
      DqItem cur, nxt;
      int n = 0, x = 0;
      bool have = dq_item<CAUSAL>(p, persist_item(0), nqt, cur);
      if (have) load_q(0, cur);
      while (have) {
        const bool have_next = dq_item<CAUSAL>(p, persist_item(n + 1), nqt, nxt);
        const int iq = cur.nit > 1 ? 1 : 0;  // the next item's Q / dO are requested after this item's first K/V tiles
        for (int i = 0; i < cur.nit; ++i, ++x) {
          const int st = x % ST;
          const int krow = cur.b * p.sk + (i + cur.jb) * BKV;
          mbar_wait(&kv_empty[st], ((x / ST) & 1) ^ 1);
          mbar_arrive_expect_tx(&kv_full[st], 32768);
          uint8_t* sk_ = smem + OFF_KV + st * 32768;
          tma_load_2d(sk_, &tmK, &kv_full[st], cur.kvh * HD, krow);
          tma_load_2d(sk_ + 8192, &tmK, &kv_full[st], cur.kvh * HD + 64, krow);
          tma_load_2d(sk_ + 16384, &tmV, &kv_full[st], cur.kvh * HD, krow);
          tma_load_2d(sk_ + 24576, &tmV, &kv_full[st], cur.kvh * HD + 64, krow);
          if (i == iq && have_next) load_q(n + 1, nxt);
        }
        cur = nxt;
        have = have_next;
        ++n;
      }
    }
  } else if (warp == 1) {
    // warp-uniform issue loop, one elected lane issues; S / dP of iteration x+1 and dQ of iteration x are issued in
    // whichever order their inputs become ready (mbarrier.test_wait polling), also across work-item boundaries
    const bool leader = elect_one();
    constexpr uint32_t idesc_sd = make_idesc_bf16(128, BKV, 0, 0);  // [128 queries x 64 keys]
    constexpr uint32_t idesc_dq = make_idesc_bf16(128, HD, 0, 1);
    const uint64_t q_desc0 = make_smem_desc(smem_u32(smem + OFF_QD), 16, 1024);
    const uint64_t kv_desc0 = make_smem_desc(smem_u32(smem + OFF_KV), 16, 1024);   // K-major view
    const uint64_t kv_mn0 = make_smem_desc(smem_u32(smem + OFF_KV), 8192, 1024);  // MN-major view
    DqItem si, di;                 // work items of the next S/dP issue and of the next dQ issue
    int sn = 0, dn = 0;            // their indices in this CTA's list
    int s_i = 0, d_i = 0;          // iteration inside the item
    bool s_have = dq_item<CAUSAL>(p, persist_item(0), nqt, si);
    bool d_have = s_have;
    di = si;
    int n_sd = 0, n_dq = 0;
    uint32_t spins = 0;
    while (d_have) {
      bool did = false;
      if (s_have) {
        const int sb = n_sd & 1, st = n_sd % ST;
        // a score buffer is free when the dQ MMAs of iteration n_sd-2 have been ISSUED (they read dS in place and
        // the tensor pipe executes one thread's MMAs in order)
        bool ok = (n_sd < 2) || (n_dq >= n_sd - 1);
        ok = ok && mbar_test(&kv_full[st], (n_sd / ST) & 1);
        if (ok && s_i == 0) ok = mbar_test(&q_full[sn & 1], (sn >> 1) & 1);
        if (ok) {
          tc_fence_after();
          if (leader) {
            const uint64_t q_desc = desc_adv(q_desc0, (sn & 1) * 65536);
            const uint64_t do_desc = desc_adv(q_desc, 32768);
            const uint64_t k_desc = desc_adv(kv_desc0, st * 32768);
            const uint64_t v_desc = desc_adv(k_desc, 16384);
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) {
              const uint32_t oa = (k >> 2) * 16384 + (k & 3) * 32;
              const uint32_t ob = (k >> 2) * 8192 + (k & 3) * 32;
              umma_bf16(TM_S + sb * BKV, desc_adv(q_desc, oa), desc_adv(k_desc, ob), idesc_sd, k != 0);
            }
#pragma unroll
            for (int k = 0; k < HD / 16; ++k) {
              const uint32_t oa = (k >> 2) * 16384 + (k & 3) * 32;
              const uint32_t ob = (k >> 2) * 8192 + (k & 3) * 32;
              umma_bf16(TM_DP + sb * BKV, desc_adv(do_desc, oa), desc_adv(v_desc, ob), idesc_sd, k != 0);
            }
            umma_commit(&sd_full[sb]);
            if (s_i == si.nit - 1) umma_commit(&q_empty[sn & 1]);  // last reader of this item's Q / dO
          }
          ++n_sd;
          if (++s_i == si.nit) {
            s_i = 0;
            ++sn;
            s_have = dq_item<CAUSAL>(p, persist_item(sn), nqt, si);
          }
          did = true;
        }
      }
      if (n_dq < n_sd && mbar_test(&ds_full[n_dq & 1], (n_dq >> 1) & 1) &&
          (d_i != 0 || mbar_test(&dq_free[dn & 1], ((dn >> 1) & 1) ^ 1))) {
        tc_fence_after();
        const int st = n_dq % ST;
        if (leader) {
          const uint64_t k_mn = desc_adv(kv_mn0, st * 32768);
#pragma unroll
          for (int k = 0; k < BKV / 16; ++k) {  // contraction over the 64 keys; A = dS in TMEM over dP (column-split halves)
            umma_bf16_ts(TM_DQ + (dn & 1) * 128, TM_DP + (n_dq & 1) * BKV + (k >> 1) * 32 + (k & 1) * 8,
                         desc_adv(k_mn, k * 2048), idesc_dq, (d_i | k) != 0);
          }
          umma_commit(&kv_empty[st]);
          if (d_i == di.nit - 1) umma_commit(&dq_ready[dn & 1]);
        }
        ++n_dq;
        if (++d_i == di.nit) {
          d_i = 0;
          ++dn;
          d_have = dq_item<CAUSAL>(p, persist_item(dn), nqt, di);
        }
        did = true;
      }
      poll_guard(spins, did);
    }
  }
  } else if (warp < 12) {
    setmaxnreg_inc<152>();
    // 8 softmax warps, two per TMEM lane quarter splitting the 64 key columns of a tile
    const int quarter = warp & 3;
    const int half = (warp - 4) >> 2;
    const int row = quarter * 32 + lane;  // query row within the tile == TMEM lane
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    const float sl2 = p.scale * LOG2E;
    DqItem it, nx;
    bool have = dq_item<CAUSAL>(p, persist_item(0), nqt, it);
    float l2 = 0.f, dlt = 0.f;
    bool rok = false;
    if (have) {
      rok = it.q0 + row < p.sq;
      if (rok) {
        const int64_t li = ((int64_t)it.b * p.H + it.h) * p.sq + it.q0 + row;
        l2 = p.lse[li] * LOG2E;
        dlt = p.delta[li];
      }
    }
    int x = 0;
    for (int n = 0; have; ++n) {
      // lse / delta of the NEXT item: in flight during this item's iterations
      const bool have_next = dq_item<CAUSAL>(p, persist_item(n + 1), nqt, nx);
      float l2n = 0.f, dln = 0.f;
      bool rokn = false;
      if (have_next) {
        rokn = nx.q0 + row < p.sq;
        if (rokn) {
          const int64_t li = ((int64_t)nx.b * p.H + nx.h) * p.sq + nx.q0 + row;
          l2n = p.lse[li];  // raw, scaled when the item becomes current: a multiply here would wait for the load
          dln = p.delta[li];
        }
      }
      const int q0 = it.q0;
      for (int i = 0; i < it.nit; ++i, ++x) {
        const int sb = x & 1;
        const int j0 = (i + it.jb) * BKV;
        // tile-level, hence warp-uniform (the bodies below hold .sync.aligned TMEM stores)
        const bool need_mask = (q0 + BQ > p.sq) || (j0 + BKV > p.sk) || (CAUSAL && (j0 + BKV - 1 > q0 + off)) ||
                               (win && j0 < q0 + BQ - 1 + off - p.window);
        mbar_wait_spin(&sd_full[sb], (x >> 1) & 1);
        tc_fence_after();
        uint32_t s_[32], d[32];
        tmem_ld32(TM_S + lane_addr + sb * BKV + half * 32, s_);
        tmem_ld32(TM_DP + lane_addr + sb * BKV + half * 32, d);
        tmem_ld_wait();
        // two straight-line chunks of sixteen key columns (see attn_bwd_dkdv_tc2_kernel): MUFU mixed with the FMA-pipe work
        auto body = [&](auto mask_c) {
          constexpr bool MASK = decltype(mask_c)::value;
          const int lim = CAUSAL ? min(p.sk - 1, q0 + row + off) : p.sk - 1;
          const int vis = rok ? lim - (j0 + half * 32) : -1;                      // last visible column
          const int lov = win ? q0 + row + off - p.window - (j0 + half * 32) : 0;  // first visible column
#pragma unroll
          for (int ch = 0; ch < 2; ++ch) {
            uint32_t wd[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const int c = ch * 16 + 2 * i;
              float p0 = ex2_approx(fmaf(__uint_as_float(s_[c]), sl2, -l2));
              float p1 = ex2_approx(fmaf(__uint_as_float(s_[c + 1]), sl2, -l2));
              if (MASK) {
                if (c > vis || c < lov) p0 = 0.f;
                if (c + 1 > vis || c + 1 < lov) p1 = 0.f;
              }
              wd[i] = pack2(p0 * (__uint_as_float(d[c]) - dlt), p1 * (__uint_as_float(d[c + 1]) - dlt));
            }
            tmem_st8(TM_DP + lane_addr + sb * BKV + half * 32 + ch * 8, wd);  // over this warp's own dP columns
          }
        };
        if (need_mask) body(std::true_type{});  // one warp-uniform branch per tile, never one per score
        else body(std::false_type{});
        tmem_st_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&ds_full[sb]);
      }
      it = nx;
      have = have_next;
      l2 = l2n * LOG2E;
      dlt = dln;
      rok = rokn;
    }
  } else {
    // epilogue warpgroup: one thread per query row writes dQ of item n (scaled, bf16) while the others are on n+1
    setmaxnreg_dec<112>();
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = (uint32_t)(quarter * 32) << 16;
    DqItem it;
    for (int n = 0; dq_item<CAUSAL>(p, persist_item(n), nqt, it); ++n) {
      mbar_wait(&dq_ready[n & 1], (n >> 1) & 1);
      tc_fence_after();
      const bool rok = it.q0 + row < p.sq;
      bf16* dqrow = p.dq + ((int64_t)it.b * p.sq + it.q0 + row) * p.lddq + it.h * HD;
      const uint32_t tm = TM_DQ + (n & 1) * 128 + lane_addr;
#pragma unroll
      for (int c = 0; c < HD / 32; ++c) {
        uint32_t a[32];
        tmem_ld32(tm + c * 32, a);
        tmem_ld_wait();
        if (c == HD / 32 - 1) {  // the accumulator has been read: the MMA warp may start item n+2 in it
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&dq_free[n & 1]);
        }
        if (rok) {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            float v[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(a[q * 8 + i]) * p.scale;
            stg16(dqrow + c * 32 + q * 8, pack8(v));
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 512);
}

template <bool CAUSAL>
static int launch_bwd_tc_v1(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                         int64_t ldv, const void* dO, int64_t lddo, const AttnTcBwdParams& p,
                         cudaStream_t st) {
  using namespace tcb;
  const uint64_t qcols = (uint64_t)p.H * HD, kcols = (uint64_t)p.KVH * HD;
  const uint64_t qrows = (uint64_t)p.B * p.sq, krows = (uint64_t)p.B * p.sk;
  {
    CUtensorMap tmQ, tmDO, tmK, tmV;
    if (make_tmap_2d(&tmQ, q, qcols, qrows, (uint64_t)ldq, 64, A_BQ)) return -1;
    if (make_tmap_2d(&tmDO, dO, qcols, qrows, (uint64_t)lddo, 64, A_BQ)) return -1;
    if (make_tmap_2d(&tmK, k, kcols, krows, (uint64_t)ldk, 64, A_BKV)) return -1;
    if (make_tmap_2d(&tmV, v, kcols, krows, (uint64_t)ldv, 64, A_BKV)) return -1;
    auto kern = attn_bwd_dkdv_tc_kernel<CAUSAL>;
    static bool cfg = false;
    if (!cfg) {
      VPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM));
      cfg = true;
    }
    dim3 grid((p.sk + A_BKV - 1) / A_BKV, p.KVH, p.B);
    kern<<<grid, 320, A_SMEM, st>>>(tmQ, tmDO, tmK, tmV, p);
    VPB_LAUNCH_OK();
  }
  {
    CUtensorMap tmQ, tmDO, tmK, tmV;
    if (make_tmap_2d(&tmQ, q, qcols, qrows, (uint64_t)ldq, 64, B_BQ)) return -1;
    if (make_tmap_2d(&tmDO, dO, qcols, qrows, (uint64_t)lddo, 64, B_BQ)) return -1;
    if (make_tmap_2d(&tmK, k, kcols, krows, (uint64_t)ldk, 64, B_BKV)) return -1;
    if (make_tmap_2d(&tmV, v, kcols, krows, (uint64_t)ldv, 64, B_BKV)) return -1;
    auto kern = attn_bwd_dq_tc_kernel<CAUSAL>;
    static bool cfg = false;
    if (!cfg) {
      VPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
      cfg = true;
    }
    dim3 grid((p.sq + B_BQ - 1) / B_BQ, p.H, p.B);
    kern<<<grid, 320, B_SMEM, st>>>(tmQ, tmDO, tmK, tmV, p);
    VPB_LAUNCH_OK();
  }
  return 0;
}

template <bool CAUSAL, int HD, bool TS, bool SPLIT>
static int launch_bwd_tc_v2(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                            int64_t ldv, const void* dO, int64_t lddo, const AttnTcBwdParams& p,
                            cudaStream_t st) {
  using namespace tcb2;
  const uint64_t qcols = (uint64_t)p.H * HD, kcols = (uint64_t)p.KVH * HD;
  const uint64_t qrows = (uint64_t)p.B * p.sq, krows = (uint64_t)p.B * p.sk;
  {
    CUtensorMap tmQ, tmDO, tmK, tmV;
    if (make_tmap_2d(&tmQ, q, qcols, qrows, (uint64_t)ldq, 64, A_BQ)) return -1;
    if (make_tmap_2d(&tmDO, dO, qcols, qrows, (uint64_t)lddo, 64, A_BQ)) return -1;
    if (make_tmap_2d(&tmK, k, kcols, krows, (uint64_t)ldk, 64, A_BKV)) return -1;
    if (make_tmap_2d(&tmV, v, kcols, krows, (uint64_t)ldv, 64, A_BKV)) return -1;
    auto kern = attn_bwd_dkdv_tc2_kernel<CAUSAL, HD, TS, SPLIT>;
    static bool cfg = false;
    if (!cfg) {
      VPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, A_SMEM));
      cfg = true;
    }
    dim3 grid((p.sk + A_BKV - 1) / A_BKV, p.KVH, p.B);
    kern<<<grid, 320, A_SMEM, st>>>(tmQ, tmDO, tmK, tmV, p);
    VPB_LAUNCH_OK();
  }
  {
    CUtensorMap tmQ, tmDO, tmK, tmV;
    if (make_tmap_2d(&tmQ, q, qcols, qrows, (uint64_t)ldq, 64, B_BQ)) return -1;
    if (make_tmap_2d(&tmDO, dO, qcols, qrows, (uint64_t)lddo, 64, B_BQ)) return -1;
    if (make_tmap_2d(&tmK, k, kcols, krows, (uint64_t)ldk, 64, B_BKV)) return -1;
    if (make_tmap_2d(&tmV, v, kcols, krows, (uint64_t)ldv, 64, B_BKV)) return -1;
    const int nqt = (p.sq + B_BQ - 1) / B_BQ;
    static int n_sm = 0;
    if (n_sm == 0) {
      int dev = 0;
      VPB_CUDA(cudaGetDevice(&dev));
      VPB_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    }
    bool persistent = false;
    if constexpr (TS && SPLIT) {
      // persistent dQ kernel: self-attention shapes (every work item sees at least one key tile), no fused inverse
      // RoPE, enough work items per SM for the static round-robin to balance
      persistent = !get_option(VPB_OPT_ATTN_BWD_DQ_R1) && p.rope_cos == nullptr && p.sk >= p.sq && p.sk > 0 &&
                   nqt * p.H * p.B >= 8 * n_sm;
      if (persistent) {
        auto kernp = attn_bwd_dq_persist_kernel<CAUSAL, HD>;
        static bool cfgp = false;
        if (!cfgp) {
          VPB_CUDA(cudaFuncSetAttribute(kernp, cudaFuncAttributeMaxDynamicSharedMemorySize, tcq::SMEM));
          cfgp = true;
        }
        kernp<<<n_sm, tcq::THREADS, tcq::SMEM, st>>>(tmQ, tmDO, tmK, tmV, p);
        VPB_LAUNCH_OK();
      }
    }
    if (!persistent) {
      auto kern = attn_bwd_dq_tc2_kernel<CAUSAL, HD, TS, SPLIT>;
      static bool cfg = false;
      if (!cfg) {
        VPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, B_SMEM));
        cfg = true;
      }
      dim3 grid((nqt + 1) / 2, p.H, p.B);
      kern<<<grid, 320, B_SMEM, st>>>(tmQ, tmDO, tmK, tmV, p);
      VPB_LAUNCH_OK();
    }
  }
  return 0;
}

static bool bwd_tc_takes_v1(int window, bool causal, int sq, int sk) {
  return window == 0 && (get_option(VPB_OPT_ATTN_TC_BWD_V1) || (causal && sk < sq));
}

template <bool CAUSAL>
static int launch_bwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                         int64_t ldv, const void* dO, int64_t lddo, const AttnTcBwdParams& p,
                         int head_dim, cudaStream_t st) {
  // v2 assumes every query tile sees at least one key tile (attention.cu only routes sk >= sq here
  // for head_dim 96; head_dim 128 can still fall back to the v1 kernels)
  const bool ss = get_option(VPB_OPT_ATTN_BWD_SS) != 0;
  const bool pingpong = get_option(VPB_OPT_ATTN_BWD_PINGPONG) != 0;
#define VPB_BWD_V2(HDv)                                                                              \
  (ss ? launch_bwd_tc_v2<CAUSAL, HDv, false, false>(q, ldq, k, ldk, v, ldv, dO, lddo, p, st)          \
      : pingpong ? launch_bwd_tc_v2<CAUSAL, HDv, true, false>(q, ldq, k, ldk, v, ldv, dO, lddo, p, st) \
             : launch_bwd_tc_v2<CAUSAL, HDv, true, true>(q, ldq, k, ldk, v, ldv, dO, lddo, p, st))
  // (a caller asking for the fused inverse RoPE only gets here with head_dim 128 on the v2 kernels:
  // attn_bwd_tc clears the request otherwise)
  if (head_dim == 96) return VPB_BWD_V2(96);
  if (bwd_tc_takes_v1(p.window, CAUSAL, p.sq, p.sk))
    return launch_bwd_tc_v1<CAUSAL>(q, ldq, k, ldk, v, ldv, dO, lddo, p, st);
  return VPB_BWD_V2(128);
#undef VPB_BWD_V2
}

// entry used by vpb_attn_bwd (attention.cu) after the delta kernel: head_dim 128 / 96, one K/V segment
int attn_bwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                const void* dO, int64_t lddo, const float* lse, const float* delta, void* dq,
                int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int H, int KVH,
                int sq, int sk, int head_dim, float scale, int causal, int window,
                const float* rope_cos, const float* rope_sin, const int* rope_pos, int* rope_fused,
                cudaStream_t st) {
  AttnTcBwdParams p;
  p.window = causal ? window : 0;
  // inverse RoPE of dQ / dK in the epilogues: v2 kernels, head_dim 128, self-attention shapes
  const bool fuse = rope_cos && head_dim == 128 && sq == sk &&
                    !bwd_tc_takes_v1(p.window, causal != 0, sq, sk);
  p.rope_cos = fuse ? rope_cos : nullptr;
  p.rope_sin = fuse ? rope_sin : nullptr;
  p.rope_pos = fuse ? rope_pos : nullptr;
  if (rope_fused) *rope_fused = fuse ? 1 : 0;
  p.trace = g_trace_buffer;
  p.lse = lse; p.delta = delta;
  p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv;
  p.lddq = lddq; p.lddk = lddk; p.lddv = lddv;
  p.B = B; p.H = H; p.KVH = KVH; p.sq = sq; p.sk = sk;
  p.scale = scale;
  return causal ? launch_bwd_tc<true>(q, ldq, k, ldk, v, ldv, dO, lddo, p, head_dim, st)
                : launch_bwd_tc<false>(q, ldq, k, ldk, v, ldv, dO, lddo, p, head_dim, st);
}

// entry used by vpb_attn_fwd (attention.cu) for head_dim 128 / 96, single K/V segment
int attn_fwd_tc(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                void* o, int64_t ldo, float* lse, int B, int H, int KVH, int sq, int sk, int head_dim,
                float scale, int causal, int window, cudaStream_t st) {
  AttnTcParams p;
  p.window = causal ? window : 0;
  p.o = (bf16*)o;
  p.lse = lse;
  p.ldo = ldo;
  p.B = B; p.H = H; p.KVH = KVH; p.sq = sq; p.sk = sk;
  p.scale = scale;
  if (head_dim == 64)  // the ViT towers' non-causal attention (VPB_OPT_ATTN_FWD_TC64, default on)
    return launch_fwd_tc<false, 64>(q, ldq, k, ldk, v, ldv, p, st);
  if (head_dim == 96)
    return causal ? launch_fwd_tc<true, 96>(q, ldq, k, ldk, v, ldv, p, st)
                  : launch_fwd_tc<false, 96>(q, ldq, k, ldk, v, ldv, p, st);
  return causal ? launch_fwd_tc<true, 128>(q, ldq, k, ldk, v, ldv, p, st)
                : launch_fwd_tc<false, 128>(q, ldq, k, ldk, v, ldv, p, st);
}

}  // namespace vpb
