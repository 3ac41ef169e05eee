// Persistent warp-specialised bf16 GEMM for sm_100a: TMA (128B swizzle) → shared memory ring →
// tcgen05.mma (accumulators in TMEM, double buffered) → tcgen05.ld epilogue with fused
// bias / activation / residual.  One CTA per SM, 192 threads:
//   warp 0   : TMA producer (one elected lane)
//   warp 1   : TMEM allocator + tcgen05.mma issuer (one elected lane)
//   warps 2-5: epilogue (TMEM lane quarter = warp_idx % 4)
//
// C[M,N] = epilogue( A · Bᵀ ) with fp32 accumulation, bf16 in / bf16 out.
//   a_layout 0: A stored [M,K], K contiguous ("K-major")      1: A stored [K,M], M contiguous
//   b_layout 0: B stored [N,K], K contiguous (nn.Linear weight) 1: B stored [K,N], N contiguous
// so that forward (x·Wᵀ: 0/0), dgrad (dY·W: 0/1) and wgrad (dYᵀ·X: 1/1) all run without
// materialising a transpose — the MN-major cases are expressed in the UMMA descriptors.
//
// Replaces the cuBLAS calls behind every F.linear on the reference hot path
// (HF modeling_llama/phi3/clip → torch.nn.functional.linear).
#include "common.cuh"
#include "visper_b200.h"

namespace vpb {

struct GemmArgs {
  bf16* C;
  int64_t ldc;
  const bf16* bias;  // [N] or null
  const bf16* res;   // [M,N] residual added after the activation, or null
  int64_t ldr;
  bf16* aux;  // optional copy of the pre-activation (bias added), for activation backward
  int64_t ldaux;
  int M, N, K;
  int act;  // VPB_ACT_*
  int F;    // SwiGLU epilogues: width of one half of the packed gate|up buffer
  int aux_tiled;  // SwiGLU epilogues: g|u saved in the tile-major layout (see gu_tiled_ptr)
  int group_m;    // row blocks per L2 panel of the tile order (host-chosen from K)
  int l2_hints;   // CTA-pair kernel: evict_last for the A panel, evict_first for B tiles
  // EPI_ROPE: rotary embedding of the first rope_heads 128-wide heads of the output row
  const float* rope_cos;  // [max_pos, 64]
  const float* rope_sin;
  const int* rope_pos;    // [M] or null (position = row % rope_T)
  int rope_T, rope_heads;
};

// Tile-major layout of the saved gate|up activations: [M/128 row blocks][F/32 chunks][128 rows]
// [32 gate | 32 up] bf16.  The epilogue thread that owns a row touches one contiguous 128-byte
// line per chunk and a warp touches 4 KB contiguous, in the forward store and in the backward load
// alike (the row-major layout makes both 64-byte pieces at a 2F-element stride).
__device__ __forceinline__ bf16* gu_tiled_ptr(bf16* base, int F, int row, int chunk) {
  return base + ((int64_t)((row >> 7) * (F >> 5) + chunk) * 128 + (row & 127)) * 64;
}

// Epilogue variants.  EPI_SWIGLU_FWD: the B tile is 128 gate rows + 128 up rows of the fused
// gate|up weight, so accumulator columns [0,128) / [128,256) hold gate / up of the SAME 128 output
// columns; the epilogue writes h = silu(g)·u to C and (optionally) the bf16 g|u pair to aux.
// EPI_SWIGLU_BWD: the accumulator is dh = dY·W_down; the epilogue reads the saved g|u from aux and
// writes d_gate | d_up to C.  Both replace a full extra pass over [M, 2F] in HBM.
constexpr int EPI_STD = 0;
constexpr int EPI_SWIGLU_FWD = 1;
constexpr int EPI_SWIGLU_BWD = 2;
// EPI_ROPE: the fused QKV projection; a 256-column tile is two 128-wide heads, the thread that owns a
// row rotates (x[j], x[j+64]) by its position's angle before the store (HF apply_rotary_pos_emb,
// rotate_half convention) — V heads pass through.
constexpr int EPI_ROPE = 3;

constexpr int BM = 128;
constexpr int BK = 64;

__device__ __forceinline__ float act_apply(float x, int act) {
  switch (act) {
    case VPB_ACT_GELU:
      return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
    case VPB_ACT_QUICK_GELU:
      return x / (1.f + __expf(-1.702f * x));
    case VPB_ACT_RELU:
      return fmaxf(x, 0.f);
    default:
      return x;
  }
}

// Epilogue of one accumulator tile for one warp: thread `lane` owns output row `row` and walks
// the BN fp32 accumulator columns at TMEM address `taddr` (lane quarter already applied).
// [c_begin, c_end): the 32-column chunks of the tile this warp drains (EPI_STD only; the fused epilogues
// always take the whole tile) — the 8-epilogue-warp variant gives each warp of a lane quarter one half.
template <int BN, int EPI>
__device__ __forceinline__ void gemm_epilogue_tile(const GemmArgs& args, uint32_t taddr, int row, int nb,
                                                   int c_begin = 0, int c_end = BN / 32) {
  constexpr int BNO = (EPI == EPI_SWIGLU_FWD) ? BN / 2 : BN;
  const bool row_ok = row < args.M;
  bf16* crow = args.C + (int64_t)row * args.ldc;
  const bf16* rrow = args.res ? args.res + (int64_t)row * args.ldr : nullptr;
  bf16* xrow = args.aux ? args.aux + (int64_t)row * args.ldaux : nullptr;
  if constexpr (EPI == EPI_SWIGLU_FWD) {
#pragma unroll 1
    for (int c = 0; c < BNO / 32; ++c) {
      const int n_base = nb * BNO + c * 32;
      uint32_t rg[32], ru[32];
      tmem_ld32(taddr + c * 32, rg);
      tmem_ld32(taddr + BNO + c * 32, ru);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int n = n_base + g * 8;
          float gv[8], uv[8], o[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            gv[j] = __uint_as_float(rg[g * 8 + j]);
            uv[j] = __uint_as_float(ru[g * 8 + j]);
          }
          // round to bf16 first: h is then bit-identical to swiglu_fwd_kernel on the stored g|u
          const uint4 pg = pack8(gv), pu = pack8(uv);
          if (args.aux) {
            if (args.aux_tiled) {
              bf16* t = gu_tiled_ptr(args.aux, args.F, row, n_base >> 5);
              stg16(t + g * 8, pg);
              stg16(t + 32 + g * 8, pu);
            } else {
              stg16(xrow + n, pg);
              stg16(xrow + args.F + n, pu);
            }
          }
          unpack8(pg, gv);
          unpack8(pu, uv);
#pragma unroll
          for (int j = 0; j < 8; ++j) o[j] = gv[j] / (1.f + __expf(-gv[j])) * uv[j];
          stg16(crow + n, pack8(o));
        }
      }
    }
  } else if constexpr (EPI == EPI_ROPE) {
    static_assert(EPI != EPI_ROPE || BN == 256, "rope epilogue: two 128-wide heads per tile");
    const int pos = args.rope_pos ? (row_ok ? args.rope_pos[row] : 0) : (row % args.rope_T);
    const float* cs = args.rope_cos + (int64_t)pos * 64;
    const float* sn = args.rope_sin + (int64_t)pos * 64;
#pragma unroll 1
    for (int hh = 0; hh < 2; ++hh) {
      const int head = nb * 2 + hh;
      if (head * 128 >= args.N) break;  // warp-uniform
      const bool rot = head < args.rope_heads;
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {  // columns [32c, 32c+32) pair with [64+32c, 64+32c+32)
        uint32_t ra[32], rb[32];
        tmem_ld32(taddr + hh * 128 + c * 32, ra);
        tmem_ld32(taddr + hh * 128 + 64 + c * 32, rb);
        tmem_ld_wait();
        if (row_ok) {
          bf16* o1p = crow + head * 128 + c * 32;
          bf16* o2p = o1p + 64;
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            float a[8], b[8], o1[8], o2[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              a[j] = __uint_as_float(ra[g * 8 + j]);
              b[j] = __uint_as_float(rb[g * 8 + j]);
            }
            // the unfused path stores the projection in bf16 and rotates that: same rounding here
            const uint4 pa = pack8(a), pb = pack8(b);
            if (rot) {
              unpack8(pa, a);
              unpack8(pb, b);
              const float4 c0 = __ldg(reinterpret_cast<const float4*>(cs + c * 32 + g * 8));
              const float4 c1 = __ldg(reinterpret_cast<const float4*>(cs + c * 32 + g * 8 + 4));
              const float4 s0 = __ldg(reinterpret_cast<const float4*>(sn + c * 32 + g * 8));
              const float4 s1 = __ldg(reinterpret_cast<const float4*>(sn + c * 32 + g * 8 + 4));
              const float cc[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
              const float ss[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
#pragma unroll
              for (int j = 0; j < 8; ++j) {
                o1[j] = __fmaf_rn(a[j], cc[j], -__fmul_rn(b[j], ss[j]));
                o2[j] = __fmaf_rn(b[j], cc[j], __fmul_rn(a[j], ss[j]));
              }
              stg16(o1p + g * 8, pack8(o1));
              stg16(o2p + g * 8, pack8(o2));
            } else {
              stg16(o1p + g * 8, pa);
              stg16(o2p + g * 8, pb);
            }
          }
        }
      }
    }
  } else if constexpr (EPI == EPI_SWIGLU_BWD) {
#pragma unroll 1
    for (int c = 0; c < BN / 32; ++c) {
      const int n_base = nb * BN + c * 32;
      if (n_base >= args.N) break;  // warp-uniform
      uint4 lg[4], lu[4];
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int n = n_base + g * 8;
        if (row_ok && n + 8 <= args.N) {
          if (args.aux_tiled) {
            const bf16* t = gu_tiled_ptr(args.aux, args.F, row, n_base >> 5);
            lg[g] = ldg16_stream(t + g * 8);
            lu[g] = ldg16_stream(t + 32 + g * 8);
          } else {
            lg[g] = ldg16_stream(xrow + n);
            lu[g] = ldg16_stream(xrow + args.F + n);
          }
        }
      }
      uint32_t r[32];
      tmem_ld32(taddr + c * 32, r);
      tmem_ld_wait();
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const int n = n_base + g * 8;
        if (row_ok && n + 8 <= args.N) {
          float d[8], gv[8], uv[8], dg[8], du[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) d[j] = __uint_as_float(r[g * 8 + j]);
          unpack8(pack8(d), d);  // dh rounded to bf16 as the unfused dgrad GEMM would store it
          unpack8(lg[g], gv);
          unpack8(lu[g], uv);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float sg = 1.f / (1.f + __expf(-gv[j]));
            const float silu = gv[j] * sg;
            du[j] = d[j] * silu;
            dg[j] = d[j] * uv[j] * (sg + silu * (1.f - sg));
          }
          stg16(crow + n, pack8(dg));
          stg16(crow + args.F + n, pack8(du));
        }
      }
    }
  } else {
#pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    const int n_base = nb * BN + c * 32;
    if (n_base >= args.N) break;  // warp-uniform
    uint32_t r[32];
    tmem_ld32(taddr + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      const int n = n_base + g * 8;
      if (n + 8 <= args.N) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __uint_as_float(r[g * 8 + j]);
        if (args.bias) {
          float b[8];
          unpack8(ldg16(args.bias + n), b);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += b[j];
        }
        if (row_ok) {
          if (xrow) stg16(xrow + n, pack8(v));
          if (args.act != VPB_ACT_NONE) {
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] = act_apply(v[j], args.act);
          }
          if (rrow) {
            float q[8];
            unpack8(*reinterpret_cast<const uint4*>(rrow + n), q);
#pragma unroll
            for (int j = 0; j < 8; ++j) v[j] += q[j];
          }
          stg16(crow + n, pack8(v));
        }
      }
    }
  }
  }
}

template <int BN, bool A_MN, bool B_MN, int EPI = EPI_STD>
__global__ void __launch_bounds__(192, 1)
gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                    const __grid_constant__ CUtensorMap tmB, const GemmArgs args) {
  constexpr int A_BYTES = BM * BK * 2;
  constexpr int B_BYTES = BN * BK * 2;
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int STAGES = (BN == 256) ? 4 : 6;
  constexpr uint32_t TMEM_COLS = 2 * BN;  // two accumulator stages (512 or 256 columns)

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) &
                                             ~static_cast<uintptr_t>(1023));
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty = full + STAGES;
  uint64_t* tfull = empty + STAGES;
  uint64_t* tempty = tfull + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  const int num_m = (args.M + BM - 1) / BM;
  constexpr int BNO = (EPI == EPI_SWIGLU_FWD) ? BN / 2 : BN;  // output columns per tile
  const int num_n = (args.N + BNO - 1) / BNO;
  const int num_tiles = num_m * num_n;
  const int num_kb = (args.K + BK - 1) / BK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // tile order: groups of GROUP_M row-blocks swept over all column-blocks, so one wave of
  // 148 CTAs shares a ~2048-row A panel and a ~2300-row B panel in L2.
  auto decode_tile = [&](int t, int& mb, int& nb) {
    const int per_group = args.group_m * num_n;
    const int g = t / per_group;
    const int first_m = g * args.group_m;
    const int gsize = min(args.group_m, num_m - first_m);
    const int r = t - g * per_group;
    mb = first_m + r % gsize;
    nb = r / gsize;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        int mb, nb;
        decode_tile(t, mb, nb);
        const int m0 = mb * BM, n0 = nb * BNO;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&full[s], STAGE_BYTES);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d(sa, &tmA, &full[s], kb * BK, m0);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)
              tma_load_2d(sa + c * 8192, &tmA, &full[s], m0 + c * 64, kb * BK);
          }
          if constexpr (EPI == EPI_SWIGLU_FWD) {
            static_assert(EPI != EPI_SWIGLU_FWD || (!B_MN && BN == 256), "swiglu fwd: K-major B, BN 256");
            tma_load_2d(sb, &tmB, &full[s], kb * BK, n0);                       // gate rows
            tma_load_2d(sb + BNO * BK * 2, &tmB, &full[s], kb * BK, args.F + n0);  // up rows
          } else if constexpr (!B_MN) {
            tma_load_2d(sb, &tmB, &full[s], kb * BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(sb + c * 8192, &tmB, &full[s], n0 + c * 64, kb * BK);
          }
        }
      }
    }
  } else if (warp == 1) {
    {  // warp-uniform loop (descriptor arithmetic on the uniform datapath), one elected lane issues
      const bool leader = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      // descriptor start-address increment (bytes) per UMMA_K = 16 slice
      constexpr uint32_t A_KADV = A_MN ? 16 * 128 : 16 * 2;
      constexpr uint32_t B_KADV = B_MN ? 16 * 128 : 16 * 2;
      const uint32_t s0 = smem_u32(smem);
      const uint64_t adesc0 = A_MN ? make_smem_desc(s0, 8192, 1024) : make_smem_desc(s0, 16, 1024);
      const uint64_t bdesc0 = B_MN ? make_smem_desc(s0 + A_BYTES, 8192, 1024) : make_smem_desc(s0 + A_BYTES, 16, 1024);
      uint32_t it = 0, tile_iter = 0;
      for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tile_iter) {
        const uint32_t acc = tile_iter & 1;
        const uint32_t acc_ph = (tile_iter >> 1) & 1;
        mbar_wait(&tempty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (leader) {
            const uint64_t adesc = desc_adv(adesc0, s * STAGE_BYTES);
            const uint64_t bdesc = desc_adv(bdesc0, s * STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              umma_bf16(d_tmem, desc_adv(adesc, k * A_KADV), desc_adv(bdesc, k * B_KADV), idesc,
                        (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit(&empty[s]);  // frees the smem slot when these MMAs retire
          }
        }
        if (leader) umma_commit(&tfull[acc]);  // accumulator complete → epilogue
      }
    }
  } else {
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32)
    uint32_t tile_iter = 0;
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++tile_iter) {
      int mb, nb;
      decode_tile(t, mb, nb);
      const uint32_t acc = tile_iter & 1;
      const uint32_t acc_ph = (tile_iter >> 1) & 1;
      mbar_wait(&tfull[acc], acc_ph);
      tc_fence_after();
      gemm_epilogue_tile<BN, EPI>(args, tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN,
                                  mb * BM + quarter * 32 + lane, nb);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ----------------------------------------------------------------------------------------------
// CTA-pair variant (cta_group::2): two CTAs of a cluster compute one 256 x 256 tile.
//   * each CTA stages its own 128 rows of A and HALF of the B tile (128 of the 256 N rows), so a
//     stage is 32 KB (6-deep ring) instead of 48 KB and every B byte is fetched and held once per
//     pair — a third less TMA / shared-memory traffic per FLOP than the 1-CTA kernel;
//   * both producers' TMA loads complete on the LEADER's "full" barrier; the leader's single MMA
//     thread issues tcgen05.mma.cta_group::2 (M = 256: 128 TMEM lanes in each CTA) and its commits
//     multicast to the "empty" / "accumulator full" barriers of both CTAs;
//   * each CTA's four epilogue warps drain their own TMEM half; the accumulator-free barrier lives
//     in the leader and counts the epilogue warps of both CTAs (remote mbarrier.arrive).
// ----------------------------------------------------------------------------------------------
constexpr int PAIR_BN = 256;
constexpr int PAIR_STAGES = 6;

// EW = epilogue warps per CTA: 4 (default), or 8 (EXPERIMENTAL, VPB_OPT_GEMM_EPI8, EPI_STD only, not yet
// measured on hardware): two warps per TMEM lane quarter, each draining half of the 256 columns.  For K <= 1024 a
// tile is <= 64 tcgen05.mma (~4.2 k tensor cycles at M = 256) while four warps need >= 6 k issue cycles for 256
// columns of bias + exact-erf GELU + pack + store per thread, so the epilogue, not the tensor pipe, paces the kernel.
template <bool A_MN, bool B_MN, int EPI = EPI_STD, int EW = 4>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(64 + 32 * EW, 1)
gemm_tcgen05_pair_kernel(const __grid_constant__ CUtensorMap tmA,
                         const __grid_constant__ CUtensorMap tmB, const GemmArgs args) {
  constexpr int BN = PAIR_BN, STAGES = PAIR_STAGES;
  constexpr int A_BYTES = BM * BK * 2;        // this CTA's 128 rows of A
  constexpr int B_BYTES = (BN / 2) * BK * 2;  // this CTA's half of the B tile
  constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  constexpr int BNO = (EPI == EPI_SWIGLU_FWD) ? BN / 2 : BN;  // output columns per tile

  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);  // used in the leader
  uint64_t* empty = full + STAGES;   // per CTA: the pair's MMAs have consumed this CTA's stage
  uint64_t* tfull = empty + STAGES;  // per CTA: accumulator stage complete
  uint64_t* tempty = tfull + 2;      // leader: both CTAs' epilogues drained the stage
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool is_leader_cta = rank == 0;

  const int num_m = (args.M + 2 * BM - 1) / (2 * BM);
  const int num_n = (args.N + BNO - 1) / BNO;
  const int num_tiles = num_m * num_n;
  const int num_kb = (args.K + BK - 1) / BK;
  const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    if (smem_u32(smem) & 1023) {
      printf("[vpb] dynamic shared memory base is not 1024-byte aligned\n");
      __trap();
    }
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    for (int i = 0; i < STAGES; ++i) {
      mbar_init(&full[i], 1);
      mbar_init(&empty[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull[i], 1);
      mbar_init(&tempty[i], 2 * EW);  // the epilogue warps of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // barriers of both CTAs initialised before any remote arrive / TMA
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  auto decode_tile = [&](int t, int& mb, int& nb) {
    const int per_group = args.group_m * num_n;
    const int g = t / per_group;
    const int first_m = g * args.group_m;
    const int gsize = min(args.group_m, num_m - first_m);
    const int r = t - g * per_group;
    mb = first_m + r % gsize;
    nb = r / gsize;
  };

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters) {
        int mb, nb;
        decode_tile(t, mb, nb);
        const int m0 = mb * 2 * BM + (int)rank * BM;
        int nrow;  // first of this CTA's 128 B rows
        if constexpr (EPI == EPI_SWIGLU_FWD) nrow = (rank == 0 ? 0 : args.F) + nb * BNO;
        else nrow = nb * BN + (int)rank * (BN / 2);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&empty[s], ph ^ 1);
          if (is_leader_cta) mbar_arrive_expect_tx(&full[s], 2 * STAGE_BYTES);  // both CTAs' bytes
          const uint32_t fbar = mapa_u32(&full[s], 0);
          uint8_t* sa = smem + s * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          // L2 priorities follow the tile order: the A panel of a group is re-read for every column
          // block of the sweep (keep it), a B tile is shared by the row blocks of ONE wave (let it go)
          // (measured in-step, same box: 1 = A last + B first loses 2.8 %; default 0 = no hints)
          const uint64_t pa = (args.l2_hints == 1 || args.l2_hints == 2) ? L2_EVICT_LAST : L2_EVICT_NORMAL;
          const uint64_t pb = args.l2_hints == 1 ? L2_EVICT_FIRST : (args.l2_hints == 3 ? L2_EVICT_LAST : L2_EVICT_NORMAL);
          if constexpr (!A_MN) {
            tma_load_2d_pair_hint(sa, &tmA, fbar, kb * BK, m0, pa);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) tma_load_2d_pair_hint(sa + c * 8192, &tmA, fbar, m0 + c * 64, kb * BK, pa);
          }
          if constexpr (!B_MN) {
            tma_load_2d_pair_hint(sb, &tmB, fbar, kb * BK, nrow, pb);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 128; ++c) tma_load_2d_pair_hint(sb + c * 8192, &tmB, fbar, nrow + c * 64, kb * BK, pb);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (is_leader_cta) {  // warp-uniform loop, one elected lane issues for the pair
      const bool leader = elect_one();
      constexpr uint32_t idesc = make_idesc_bf16(2 * BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
      constexpr uint32_t A_KADV = A_MN ? 16 * 128 : 16 * 2;
      constexpr uint32_t B_KADV = B_MN ? 16 * 128 : 16 * 2;
      const uint32_t s0 = smem_u32(smem);
      const uint64_t adesc0 = A_MN ? make_smem_desc(s0, 8192, 1024) : make_smem_desc(s0, 16, 1024);
      const uint64_t bdesc0 = B_MN ? make_smem_desc(s0 + A_BYTES, 8192, 1024) : make_smem_desc(s0 + A_BYTES, 16, 1024);
      uint32_t it = 0, tile_iter = 0;
      for (int t = cluster_id; t < num_tiles; t += num_clusters, ++tile_iter) {
        const uint32_t acc = tile_iter & 1;
        const uint32_t acc_ph = (tile_iter >> 1) & 1;
        mbar_wait(&tempty[acc], acc_ph ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BN;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES;
          const uint32_t ph = (it / STAGES) & 1;
          mbar_wait(&full[s], ph);
          tc_fence_after();
          if (leader) {
            const uint64_t adesc = desc_adv(adesc0, s * STAGE_BYTES);
            const uint64_t bdesc = desc_adv(bdesc0, s * STAGE_BYTES);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              umma_bf16_pair(d_tmem, desc_adv(adesc, k * A_KADV), desc_adv(bdesc, k * B_KADV), idesc,
                             (kb | k) != 0 ? 1u : 0u);
            }
            umma_commit_pair(&empty[s], 3);  // frees the stage in both CTAs
          }
        }
        if (leader) umma_commit_pair(&tfull[acc], 3);  // accumulator complete in both CTAs
      }
    }
  } else {
    const int quarter = warp & 3;  // TMEM lanes [32*quarter, 32*quarter+32)
    uint32_t tile_iter = 0;
    for (int t = cluster_id; t < num_tiles; t += num_clusters, ++tile_iter) {
      int mb, nb;
      decode_tile(t, mb, nb);
      const uint32_t acc = tile_iter & 1;
      const uint32_t acc_ph = (tile_iter >> 1) & 1;
      mbar_wait(&tfull[acc], acc_ph);
      tc_fence_after();
      if constexpr (EW == 8) {
        static_assert(EW == 4 || EPI == EPI_STD, "8 epilogue warps: standard epilogue only");
        const int part = (warp - 2) >> 2;  // warps 2-5 take columns [0,128), warps 6-9 [128,256)
        gemm_epilogue_tile<BN, EPI>(args, tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN,
                                    mb * 2 * BM + (int)rank * BM + quarter * 32 + lane, nb, part * (BN / 64),
                                    (part + 1) * (BN / 64));
      } else {
        gemm_epilogue_tile<BN, EPI>(args, tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN,
                                    mb * 2 * BM + (int)rank * BM + quarter * 32 + lane, nb);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(mapa_u32(&tempty[acc], 0));
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // the peer may still be reading this CTA's smem / signalling its barriers
  if (warp == 1) tmem_dealloc_pair(tmem_base, 512);
}

// ----------------------------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// 2-D bf16 tensor map: `inner` contiguous elements, `outer` rows `ld` elements apart.
int make_tmap_2d(CUtensorMap* m, const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld,
                 uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn fn = get_encode_fn();
  VPB_CHECK(fn != nullptr, "cuTensorMapEncodeTiled entry point unavailable (no CUDA driver?)");
  VPB_CHECK((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "TMA operand must be 16-byte aligned");
  VPB_CHECK((ld * 2) % 16 == 0, "TMA row stride must be a multiple of 16 bytes (ld=%llu)",
            (unsigned long long)ld);
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides,
                  box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  VPB_CHECK(r == CUDA_SUCCESS, "cuTensorMapEncodeTiled failed with %d", (int)r);
  return 0;
}

static int g_num_sms = 0;
int num_sms() {
  if (!g_num_sms) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  return g_num_sms;
}

// Tile order: groups of `panel rows` of A are swept over all column blocks, so the A panel stays
// in L2 while B streams past it once per group.  A bigger panel means fewer passes over B (the
// weights); it must still fit next to the streaming operand in the 126 MB L2: ~32 MB of A.
static int g_panel_mb = 32;
static int panel_rows_for(int K) {
  int64_t rows = ((int64_t)g_panel_mb << 20) / ((int64_t)K * 2);
  rows = (rows / 256) * 256;
  if (rows < 512) rows = 512;
  if (rows > 8192) rows = 8192;
  return (int)rows;
}

template <int BN, bool A_MN, bool B_MN, int EPI = EPI_STD>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& args,
                       cudaStream_t stream) {
  constexpr int STAGES = (BN == 256) ? 4 : 6;
  constexpr int SMEM = STAGES * (BM * BK * 2 + BN * BK * 2) + 1024 + 256;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, EPI>;
  static bool configured = false;
  if (!configured) {
    VPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  constexpr int BNO = (EPI == EPI_SWIGLU_FWD) ? BN / 2 : BN;
  const int tiles = ((args.M + BM - 1) / BM) * ((args.N + BNO - 1) / BNO);
  const int grid = tiles < num_sms() ? tiles : num_sms();
  GemmArgs a2 = args;
  a2.group_m = panel_rows_for(args.K) / BM;
  a2.l2_hints = 0;
  kern<<<grid, 192, SMEM, stream>>>(tmA, tmB, a2);
  VPB_LAUNCH_OK();
  return 0;
}


template <bool A_MN, bool B_MN, int EPI = EPI_STD, int EW = 4>
static int launch_gemm_pair(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmArgs& args,
                            cudaStream_t stream) {
  if constexpr (EPI == EPI_STD && EW == 4) {
    if (args.K <= 1024 && get_option(VPB_OPT_GEMM_EPI8))
      return launch_gemm_pair<A_MN, B_MN, EPI, 8>(tmA, tmB, args, stream);
  }
  constexpr int SMEM = PAIR_STAGES * (BM * BK * 2 + (PAIR_BN / 2) * BK * 2) + 256;
  auto kern = gemm_tcgen05_pair_kernel<A_MN, B_MN, EPI, EW>;
  static bool configured = false;
  if (!configured) {
    VPB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  constexpr int BNO = (EPI == EPI_SWIGLU_FWD) ? PAIR_BN / 2 : PAIR_BN;
  const int tiles = ((args.M + 2 * BM - 1) / (2 * BM)) * ((args.N + BNO - 1) / BNO);
  const int pairs = num_sms() / 2;
  const int grid = 2 * (tiles < pairs ? tiles : pairs);
  GemmArgs a2 = args;
  a2.group_m = panel_rows_for(args.K) / (2 * BM);
  a2.l2_hints = get_option(VPB_OPT_GEMM_L2_HINTS);
  kern<<<grid, 64 + 32 * EW, SMEM, stream>>>(tmA, tmB, a2);
  VPB_LAUNCH_OK();
  return 0;
}

// CTA pairs win on large problems (less TMA / shared-memory traffic per FLOP); on small ones the
// coarser 256 x 256 tiles can quantise into more waves than the 1-CTA kernel, so compare the two
// wave efficiencies (measured: CLIP QKV M=4616 N=3072 is 3.08 pair waves vs exactly 3 single waves).
static bool use_pair(int M, int N) {
  if (get_option(VPB_OPT_GEMM_PANEL_MB) > 0) g_panel_mb = get_option(VPB_OPT_GEMM_PANEL_MB);
  if (get_option(VPB_OPT_GEMM_1CTA)) return false;
  if (N < 256 || M < 256) return false;
  const int pairs = num_sms() / 2, sms = num_sms();
  const int64_t tp = (int64_t)((M + 255) / 256) * ((N + 255) / 256);
  if (tp < pairs) return false;
  const int64_t t1 = (int64_t)((M + 127) / 128) * ((N + 255) / 256);
  const double eff_p = (double)tp / (double)(((tp + pairs - 1) / pairs) * pairs);
  const double eff_1 = (double)t1 / (double)(((t1 + sms - 1) / sms) * sms);
  return eff_p + 0.05 >= eff_1;
}

}  // namespace vpb

using namespace vpb;

extern "C" int vpb_gemm_bf16(const void* A, int64_t lda, int a_layout, const void* B, int64_t ldb,
                             int b_layout, void* C, int64_t ldc, int M, int N, int K, int act,
                             const void* bias, const void* residual, int64_t ldr, void* aux,
                             int64_t ldaux, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VPB_CHECK(M > 0 && N > 0 && K > 0, "gemm: empty problem M=%d N=%d K=%d", M, N, K);
  VPB_CHECK(N % 8 == 0, "gemm: N=%d must be a multiple of 8", N);
  VPB_CHECK(ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0,
            "gemm: C must be 16-byte aligned with ldc %% 8 == 0");
  VPB_CHECK(!residual || (ldr % 8 == 0 && (reinterpret_cast<uintptr_t>(residual) & 15) == 0),
            "gemm: residual alignment");
  VPB_CHECK(!aux || (ldaux % 8 == 0 && (reinterpret_cast<uintptr_t>(aux) & 15) == 0),
            "gemm: aux alignment");
  VPB_CHECK(!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0, "gemm: bias alignment");
  VPB_CHECK(act >= VPB_ACT_NONE && act <= VPB_ACT_RELU, "gemm: unknown activation %d", act);

  const int tiles256 = ((M + BM - 1) / BM) * ((N + 255) / 256);
  const bool bn256 = tiles256 >= num_sms() && N >= 256;
  const bool pair = use_pair(M, N);
  const int BN = pair ? 128 : (bn256 ? 256 : 128);  // B rows per TMA box (a pair CTA loads half a tile)

  CUtensorMap tmA, tmB;
  if (a_layout == 0) {
    if (make_tmap_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM)) return -1;
  } else {
    if (make_tmap_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, 64, BK)) return -1;
  }
  if (b_layout == 0) {
    if (make_tmap_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, BN)) return -1;
  } else {
    if (make_tmap_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, 64, BK)) return -1;
  }
  GemmArgs args;
  args.C = static_cast<bf16*>(C);
  args.ldc = ldc;
  args.bias = static_cast<const bf16*>(bias);
  args.res = static_cast<const bf16*>(residual);
  args.ldr = ldr;
  args.aux = static_cast<bf16*>(aux);
  args.ldaux = ldaux;
  args.M = M;
  args.N = N;
  args.K = K;
  args.act = act;
  args.F = 0;
  args.aux_tiled = 0;
  args.rope_cos = args.rope_sin = nullptr;
  args.rope_pos = nullptr;
  args.rope_T = 1;
  args.rope_heads = 0;

  if (pair) {
    switch ((a_layout ? 2 : 0) | (b_layout ? 1 : 0)) {
      case 0: return launch_gemm_pair<false, false>(tmA, tmB, args, stream);
      case 1: return launch_gemm_pair<false, true>(tmA, tmB, args, stream);
      case 2: return launch_gemm_pair<true, false>(tmA, tmB, args, stream);
      default: return launch_gemm_pair<true, true>(tmA, tmB, args, stream);
    }
  }
  const int key = (bn256 ? 4 : 0) | (a_layout ? 2 : 0) | (b_layout ? 1 : 0);
  switch (key) {
    case 0: return launch_gemm<128, false, false>(tmA, tmB, args, stream);
    case 1: return launch_gemm<128, false, true>(tmA, tmB, args, stream);
    case 2: return launch_gemm<128, true, false>(tmA, tmB, args, stream);
    case 3: return launch_gemm<128, true, true>(tmA, tmB, args, stream);
    case 4: return launch_gemm<256, false, false>(tmA, tmB, args, stream);
    case 5: return launch_gemm<256, false, true>(tmA, tmB, args, stream);
    case 6: return launch_gemm<256, true, false>(tmA, tmB, args, stream);
    default: return launch_gemm<256, true, true>(tmA, tmB, args, stream);
  }
}

// gu[M,2F] = A·Wguᵀ (optional store) and h[M,F] = silu(gate)·up in ONE launch.
extern "C" int vpb_gemm_swiglu_fwd(const void* A, int64_t lda, const void* Wgu, int64_t ldw,
                                   void* gu, int64_t ldgu, int gu_tiled, void* h, int64_t ldh,
                                   int M, int F, int K, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VPB_CHECK(M > 0 && F > 0 && K > 0, "gemm_swiglu_fwd: empty problem M=%d F=%d K=%d", M, F, K);
  VPB_CHECK(F % 128 == 0, "gemm_swiglu_fwd: F=%d must be a multiple of 128", F);
  VPB_CHECK(ldh % 8 == 0 && (reinterpret_cast<uintptr_t>(h) & 15) == 0, "gemm_swiglu_fwd: h alignment");
  VPB_CHECK(!gu || (ldgu % 8 == 0 && (reinterpret_cast<uintptr_t>(gu) & 15) == 0),
            "gemm_swiglu_fwd: gu alignment");
  CUtensorMap tmA, tmB;
  if (make_tmap_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM)) return -1;
  if (make_tmap_2d(&tmB, Wgu, (uint64_t)K, (uint64_t)(2 * (int64_t)F), (uint64_t)ldw, BK, 128)) return -1;
  GemmArgs args;
  args.C = static_cast<bf16*>(h);
  args.ldc = ldh;
  args.bias = nullptr;
  args.res = nullptr;
  args.ldr = 0;
  args.aux = static_cast<bf16*>(gu);
  args.ldaux = ldgu;
  args.M = M;
  args.N = F;
  args.K = K;
  args.act = VPB_ACT_NONE;
  args.F = F;
  args.aux_tiled = gu_tiled;
  args.rope_cos = args.rope_sin = nullptr;
  args.rope_pos = nullptr;
  args.rope_T = 1;
  args.rope_heads = 0;
  if (use_pair(M, 2 * F)) return launch_gemm_pair<false, false, EPI_SWIGLU_FWD>(tmA, tmB, args, stream);
  return launch_gemm<256, false, false, EPI_SWIGLU_FWD>(tmA, tmB, args, stream);
}

// dgu[M,2F] = swiglu'(gu) ∘ (dY·W) where W is the down projection: b_layout 1 → W is [K=D, F]
// (the nn.Linear weight read MN-major), b_layout 0 → W is a K-major transposed copy [F, D].
extern "C" int vpb_gemm_swiglu_bwd(const void* dY, int64_t lddy, const void* W, int64_t ldw,
                                   int b_layout, const void* gu, int64_t ldgu, int gu_tiled,
                                   void* dgu, int64_t lddgu, int M, int F, int K, void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VPB_CHECK(M > 0 && F > 0 && K > 0, "gemm_swiglu_bwd: empty problem M=%d F=%d K=%d", M, F, K);
  VPB_CHECK(F % 8 == 0, "gemm_swiglu_bwd: F=%d must be a multiple of 8", F);
  VPB_CHECK(!gu_tiled || F % 32 == 0, "gemm_swiglu_bwd: tiled g|u needs F %% 32 == 0 (F=%d)", F);
  VPB_CHECK(gu && dgu && ldgu % 8 == 0 && lddgu % 8 == 0 &&
                (reinterpret_cast<uintptr_t>(gu) & 15) == 0 && (reinterpret_cast<uintptr_t>(dgu) & 15) == 0,
            "gemm_swiglu_bwd: gu/dgu alignment");
  CUtensorMap tmA, tmB;
  if (make_tmap_2d(&tmA, dY, (uint64_t)K, (uint64_t)M, (uint64_t)lddy, BK, BM)) return -1;
  const bool pair = use_pair(M, F);
  if (b_layout == 0) {
    if (make_tmap_2d(&tmB, W, (uint64_t)K, (uint64_t)F, (uint64_t)ldw, BK, pair ? 128 : 256)) return -1;
  } else {
    if (make_tmap_2d(&tmB, W, (uint64_t)F, (uint64_t)K, (uint64_t)ldw, 64, BK)) return -1;
  }
  GemmArgs args;
  args.C = static_cast<bf16*>(dgu);
  args.ldc = lddgu;
  args.bias = nullptr;
  args.res = nullptr;
  args.ldr = 0;
  args.aux = const_cast<bf16*>(static_cast<const bf16*>(gu));
  args.ldaux = ldgu;
  args.M = M;
  args.N = F;
  args.K = K;
  args.act = VPB_ACT_NONE;
  args.F = F;
  args.aux_tiled = gu_tiled;
  args.rope_cos = args.rope_sin = nullptr;
  args.rope_pos = nullptr;
  args.rope_T = 1;
  args.rope_heads = 0;
  if (pair) {
    if (b_layout == 0) return launch_gemm_pair<false, false, EPI_SWIGLU_BWD>(tmA, tmB, args, stream);
    return launch_gemm_pair<false, true, EPI_SWIGLU_BWD>(tmA, tmB, args, stream);
  }
  if (b_layout == 0) return launch_gemm<256, false, false, EPI_SWIGLU_BWD>(tmA, tmB, args, stream);
  return launch_gemm<256, false, true, EPI_SWIGLU_BWD>(tmA, tmB, args, stream);
}

// C[M,N] = rope(A·Bᵀ): the fused QKV projection of a decoder layer with rotary embedding applied to
// the first `rope_heads` 128-wide heads in the epilogue (head_dim 128, N % 256 == 0).
extern "C" int vpb_gemm_rope_bf16(const void* A, int64_t lda, const void* B, int64_t ldb, void* C,
                                  int64_t ldc, int M, int N, int K, const float* cos_t,
                                  const float* sin_t, int seq_len, const int* pos_ids, int rope_heads,
                                  void* stream_) {
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  VPB_CHECK(M > 0 && N > 0 && K > 0 && N % 256 == 0, "gemm_rope: N=%d must be a positive multiple of 256", N);
  VPB_CHECK(ldc % 8 == 0 && (reinterpret_cast<uintptr_t>(C) & 15) == 0, "gemm_rope: C alignment");
  VPB_CHECK(cos_t && sin_t && seq_len > 0 && rope_heads >= 0 && rope_heads * 128 <= N, "gemm_rope: bad rope arguments");
  const bool pair = use_pair(M, N);
  CUtensorMap tmA, tmB;
  if (make_tmap_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, BK, BM)) return -1;
  if (make_tmap_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, BK, pair ? 128 : 256)) return -1;
  GemmArgs args;
  args.C = static_cast<bf16*>(C);
  args.ldc = ldc;
  args.bias = nullptr;
  args.res = nullptr;
  args.ldr = 0;
  args.aux = nullptr;
  args.ldaux = 0;
  args.M = M;
  args.N = N;
  args.K = K;
  args.act = VPB_ACT_NONE;
  args.F = 0;
  args.aux_tiled = 0;
  args.rope_cos = cos_t;
  args.rope_sin = sin_t;
  args.rope_pos = pos_ids;
  args.rope_T = seq_len;
  args.rope_heads = rope_heads;
  if (pair) return launch_gemm_pair<false, false, EPI_ROPE>(tmA, tmB, args, stream);
  return launch_gemm<256, false, false, EPI_ROPE>(tmA, tmB, args, stream);
}
