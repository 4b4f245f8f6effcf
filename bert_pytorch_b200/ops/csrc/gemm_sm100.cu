// Persistent, warp-specialised tcgen05 GEMM for sm_100a with fused epilogues.
//
//   D[M,N] = epilogue( A[M,K] * B[N,K]^T )        bf16 operands, fp32 accumulation in TMEM
//
// One CTA per SM loops over 128 x BLOCK_N output tiles.  Warp 0 (one lane) is the TMA producer,
// warp 1 (one lane) issues tcgen05.mma, warps 2-5 drain the accumulator (tcgen05.ld) and run the
// epilogue.  Operand tiles travel global -> shared memory with cp.async.bulk.tensor (128B swizzle)
// through a STAGES-deep mbarrier ring; the accumulator is double buffered in TMEM so the epilogue of
// tile i overlaps the main loop of tile i+1.
//
// Both operands can be K-major (reduction dim contiguous) or MN-major, selected per instantiation, so
// the same kernel serves the three GEMMs of a linear layer without any transpose copy
// (SURVEY.md K5/K13/K16/K17/K21/K23 and their backward call sites):
//     forward  Y  = X  W^T        A K-major (X[M,K]),     B K-major  (W[N,K])
//     dgrad    dX = dY W          A K-major (dY[M,N']),   B MN-major (W[N',K'] read as [Kred][N])
//     wgrad    dW = dY^T X        A MN-major (dY[M,N']),  B MN-major (X[M,K'])   (+ split-K, fp32 atomics)
//
// Epilogues (runtime switch, warp uniform): bias, bias+GELU(erf) with the pre-activation saved for
// backward, bias+dropout+residual, residual add, multiply by GELU'(aux), fp32 accumulate-into (wgrad into
// the gradient arena), bias+tanh, plain fp32.
#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <unordered_map>

#include "common.cuh"
#include "gemm_sm100.h"

namespace b200 {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;          // 64 bf16 = 128 B = one swizzle atom row
constexpr int UMMA_K = 16;
constexpr int NUM_THREADS = 320;     // single-CTA kernel: warp0 TMA, warp1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int PAIR_THREADS = 352;    // pair kernel: warps 0..7 epilogue, 8 and 10 TMA producers (even / odd ring slots), 9 MMA
constexpr int EPI_THREADS = 256;
constexpr int SMEM_BUDGET = 196608;  // bytes for the operand ring (192 KB)

template <int BLOCK_N>
struct Cfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = SMEM_BUDGET / STAGE_BYTES;
  static constexpr int TMEM_COLS = 2 * BLOCK_N;  // double buffered accumulator (256 or 512 columns)
  static constexpr int SMEM_TOTAL = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

constexpr int PAIR_M = 256, PAIR_N = 256, PAIR_STAGE = 32768;   // CTA-pair kernel (below)

struct GemmArgs {
  int M, N, K;
  int m_blocks, n_blocks, k_blocks, k_splits, k_per_split;
  int epi;
  void* out;
  int ldo;
  __nv_bfloat16* aux_out;
  const __nv_bfloat16* bias;
  const __nv_bfloat16* res;
  int ldr;
  float* colsum;             // optional fp32 column sums of the bf16 output (bias gradients)
  unsigned char* mask_out;   // optional dropout keep bits, [M][N / 8] bytes (GemmCall::mask_out)
  const unsigned char* mask_in;   // the same bits as input: applied instead of generated (GemmCall::mask_in)
  int warp_epi;              // pair kernel: warp-local staged epilogue (default) instead of the CTA-wide staging tile
  Seed seed;
  unsigned int stream;
  unsigned int drop_thresh16;
  float drop_scale;
  float alpha;
  const float* scale_a;      // fp8: device dequantisation factors (nullptr for bf16 operands)
  const float* scale_b;
  unsigned int fp8_fmt;      // bit0: A is e5m2 (else e4m3), bit1: B is e5m2
  int stream_k;              // pair kernel: 1 = contiguous (tile, k-block) ranges per cluster; 2 = tail split (below)
  int k_main;                // tail split: k-blocks [0, k_main) of tile t go to cluster t, the rest to the helper clusters
  // GEMM -> reduce-scatter fusion (EPI_ACCUM_F32 into a gradient arena that lives in NVLink symmetric memory):
  // peer_world > 1 makes every accumulation atomic (peers add into this arena concurrently); peer_push adds the
  // tile -- plus, once per tile, the locally accumulated value -- straight into the arena of the rank that owns
  // that shard, so the gradient is reduced tile by tile while the backward pass is still running.
  int peer_world, peer_rank, peer_push;
  long long peer_off;        // element offset of `out` inside the arena
  long long peer_per;        // shard size in elements: owner(e) = min(e / peer_per, world - 1)
  float* peer_base[16];      // gradient arena of every rank (symmetric-memory mapping)
  // measurement knobs (tools/gemm_lab.py; all zero in production): bit0 = no operand refill after the ring's first
  // fill (tensor-pipe ceiling of the issue loop), bit1 = every load fetches tile (0, 0) (L2-hit-only supply),
  // bits 8..15 = L2 prefetch distance in k-blocks; lab_stats (optional): per cluster {total, wait_full, wait_tmem,
  // k-blocks} clock cycles of the MMA issuer
  unsigned int lab;
  unsigned long long* lab_stats;
  int mn3d;                  // MN-major bf16 operands arrive as ONE 3-D box [2][64 k][64 mn] per CTA and stage (make_tmap_mn3d)
};

// Drain one 128 x BLOCK_N fp32 accumulator tile (this warp's 32 TMEM lanes) through the selected epilogue.
// `taddr` already carries the lane quarter; `row` is this thread's global output row.
template <int BLOCK_N>
__device__ __forceinline__ void epilogue_tile(const GemmArgs& p, uint32_t taddr, int row, bool row_ok, int n_base,
                                              int c_begin, int c_end, float alpha, unsigned long long seed, bool add_local = true) {
  #pragma unroll 1
  for (int c = c_begin; c < c_end; ++c) {
    const int n0 = n_base + c * 32;
    if (n0 >= p.N) break;  // warp uniform
    uint32_t v[32];
    tmem_ld_32x32(taddr + c * 32, v);
    tmem_ld_wait();
    float f[32];
    if (alpha != 1.f) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * alpha;
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
    }
    const int ncols = min(32, p.N - n0);  // multiple of 8
    const size_t roff = (size_t)row * p.ldr + n0;
    const size_t ooff = (size_t)row * p.ldo + n0;

    if (p.epi == EPI_ACCUM_F32) {
      if (row_ok) {
        float* o = reinterpret_cast<float*>(p.out) + ooff;
        if (p.peer_world > 1) {
          // a 32-column chunk crosses at most one shard boundary (shards are multiples of 2048 elements)
          const long long e0 = p.peer_off + (long long)ooff;
          const int own0 = (int)min((long long)(p.peer_world - 1), e0 / p.peer_per);
          const long long bound = own0 + 1 < p.peer_world ? (long long)(own0 + 1) * p.peer_per : (1ll << 62);
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            if (j < ncols) {
              const int owner = (e0 + j >= bound) ? own0 + 1 : own0;
              if (!p.peer_push || owner == p.peer_rank) {
                asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(f[j]),
                             "f"(f[j + 1]), "f"(f[j + 2]), "f"(f[j + 3]) : "memory");
              } else {
                float4 v = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                if (add_local) {                       // what the earlier micro-steps accumulated here
                  const float4 cur = *reinterpret_cast<const float4*>(o + j);
                  v.x += cur.x; v.y += cur.y; v.z += cur.z; v.w += cur.w;
                }
                float* dst = p.peer_base[owner] + (e0 + j);
                asm volatile("red.relaxed.sys.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(dst), "f"(v.x), "f"(v.y),
                             "f"(v.z), "f"(v.w) : "memory");
              }
            }
          }
          continue;
        }
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          if (j < ncols) {
            // fire-and-forget reduction at the L2 (one contributor per element and launch unless split-K: the sum is
            // deterministic either way).  The load-add-store chain it replaces was DRAM-latency bound: 25 % of the
            // weight-gradient kernel's time (gemm_lab: 0.125 ms with, 0.099 ms without the epilogue)
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o + j), "f"(f[j]),
                         "f"(f[j + 1]), "f"(f[j + 2]), "f"(f[j + 3])
                         : "memory");
          }
        }
      }
      continue;
    }
    if (p.epi == EPI_F32) {
      if (row_ok) {
        float* o = reinterpret_cast<float*>(p.out) + ooff;
#pragma unroll
        for (int j = 0; j < 32; j += 4)
          if (j < ncols) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
      }
      continue;
    }
    // ---- bias (same 32 values for every lane: broadcast loads)
    if (p.epi == EPI_BIAS || p.epi == EPI_BIAS_GELU || p.epi == EPI_BIAS_DROP_RES || p.epi == EPI_BIAS_TANH ||
        p.epi == EPI_BIAS_GELU_DG) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        if (g * 8 < ncols) {
          const uint4 b = __ldg(reinterpret_cast<const uint4*>(p.bias + n0 + g * 8));
          const uint32_t w[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 bb = unpack_bf16(w[t]);
            f[g * 8 + 2 * t] += bb.x;
            f[g * 8 + 2 * t + 1] += bb.y;
          }
        }
      }
    }
    if (p.epi == EPI_BIAS_GELU) {
      if (row_ok) {  // save the pre-activation, then activate
        __nv_bfloat16* a = p.aux_out + ooff;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (g * 8 < ncols)
            *reinterpret_cast<uint4*>(a + g * 8) =
                make_uint4(pack_bf16(f[g * 8], f[g * 8 + 1]), pack_bf16(f[g * 8 + 2], f[g * 8 + 3]),
                           pack_bf16(f[g * 8 + 4], f[g * 8 + 5]), pack_bf16(f[g * 8 + 6], f[g * 8 + 7]));
      }
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = gelu_erf(f[j]);
    } else if (p.epi == EPI_BIAS_GELU_DG) {
      float dg[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        const GeluParts gp = gelu_parts(f[j]);
        dg[j] = fmaf(f[j] * 0.3989422804014327f, gp.e, gp.Phi);
        f[j] *= gp.Phi;
      }
      if (row_ok) {
        __nv_bfloat16* a = p.aux_out + ooff;
#pragma unroll
        for (int g = 0; g < 4; ++g)
          if (g * 8 < ncols)
            *reinterpret_cast<uint4*>(a + g * 8) =
                make_uint4(pack_bf16(dg[g * 8], dg[g * 8 + 1]), pack_bf16(dg[g * 8 + 2], dg[g * 8 + 3]),
                           pack_bf16(dg[g * 8 + 4], dg[g * 8 + 5]), pack_bf16(dg[g * 8 + 6], dg[g * 8 + 7]));
      }
    } else if (p.epi == EPI_BIAS_TANH) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = tanhf(f[j]);
    } else if (p.epi == EPI_BIAS_DROP_RES || p.epi == EPI_ADD || p.epi == EPI_DGELU || p.epi == EPI_MUL) {
      if (p.epi == EPI_BIAS_DROP_RES && p.drop_thresh16 != 0 && row_ok) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g * 8 < ncols) {
            const uint64_t e8 = ((uint64_t)row * (uint64_t)p.N + (uint64_t)(n0 + g * 8)) >> 3;
            const Keep8 keep = dropout_keep8(seed, p.stream, e8, p.drop_thresh16);
#pragma unroll
            for (int t = 0; t < 8; ++t) f[g * 8 + t] = keep[t] ? f[g * 8 + t] * p.drop_scale : 0.f;
          }
        }
      }
      if (row_ok && p.res != nullptr) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          if (g * 8 < ncols) {
            const uint4 r = *reinterpret_cast<const uint4*>(p.res + roff + g * 8);
            const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 rr = unpack_bf16(w[t]);
              if (p.epi == EPI_DGELU) {
                f[g * 8 + 2 * t] *= dgelu_erf(rr.x);
                f[g * 8 + 2 * t + 1] *= dgelu_erf(rr.y);
              } else if (p.epi == EPI_MUL) {
                f[g * 8 + 2 * t] *= rr.x;
                f[g * 8 + 2 * t + 1] *= rr.y;
              } else {
                f[g * 8 + 2 * t] += rr.x;
                f[g * 8 + 2 * t + 1] += rr.y;
              }
            }
          }
        }
      }
    }
    if (p.colsum != nullptr && row_ok) {   // small-problem path: one atomic per element (the pair kernel reduces in smem)
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (j < ncols) atomicAdd(p.colsum + n0 + j, __bfloat162float(__float2bfloat16(f[j])));
    }
    if (row_ok) {
      __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + ooff;
#pragma unroll
      for (int g = 0; g < 4; ++g)
        if (g * 8 < ncols)
          *reinterpret_cast<uint4*>(o + g * 8) =
              make_uint4(pack_bf16(f[g * 8], f[g * 8 + 1]), pack_bf16(f[g * 8 + 2], f[g * 8 + 3]),
                         pack_bf16(f[g * 8 + 4], f[g * 8 + 5]), pack_bf16(f[g * 8 + 6], f[g * 8 + 7]));
    }
  }
}

template <bool A_MN, bool B_MN, int BLOCK_N>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
            const GemmArgs p) {
  using C = Cfg<BLOCK_N>;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + C::STAGES * C::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + C::STAGES;
  uint64_t* tmem_full = empty_bar + C::STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < C::STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], EPI_THREADS);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_base_slot, C::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  const int total_tiles = p.m_blocks * p.n_blocks * p.k_splits;

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int nb = tile % p.n_blocks;
        const int mb = (tile / p.n_blocks) % p.m_blocks;
        const int ks = tile / (p.n_blocks * p.m_blocks);
        const int kb0 = ks * p.k_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.k_per_split);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % C::STAGES;
          const uint32_t ph = (it / C::STAGES) & 1;
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * C::STAGE_BYTES;
          uint8_t* sb = sa + C::A_BYTES;
          mbar_arrive_expect_tx(&full_bar[s], C::STAGE_BYTES);
          if constexpr (!A_MN) {
            tma_load_2d(sa, &tmap_a, &full_bar[s], kb * BLOCK_K, mb * BLOCK_M);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_M / 64; ++j)
              tma_load_2d(sa + j * 8192, &tmap_a, &full_bar[s], mb * BLOCK_M + j * 64, kb * BLOCK_K);
          }
          if constexpr (!B_MN) {
            tma_load_2d(sb, &tmap_b, &full_bar[s], kb * BLOCK_K, nb * BLOCK_N);
          } else {
#pragma unroll
            for (int j = 0; j < BLOCK_N / 64; ++j)
              tma_load_2d(sb + j * 8192, &tmap_b, &full_bar[s], nb * BLOCK_N + j * 64, kb * BLOCK_K);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N, A_MN, B_MN);
      uint32_t it = 0, tile_it = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_it) {
        const int ks = tile / (p.n_blocks * p.m_blocks);
        const int kb0 = ks * p.k_per_split;
        const int kb1 = min(p.k_blocks, kb0 + p.k_per_split);
        const uint32_t as = tile_it & 1, aph = (tile_it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const int s = it % C::STAGES;
          const uint32_t ph = (it / C::STAGES) & 1;
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * C::STAGE_BYTES);
          const uint32_t sb = sa + C::A_BYTES;
          const uint64_t da0 = A_MN ? umma_smem_desc_sw128(sa, 8192, 1024) : umma_smem_desc_sw128(sa, 16, 1024);
          const uint64_t db0 = B_MN ? umma_smem_desc_sw128(sb, 8192, 1024) : umma_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
            // advance inside the tile: K-major +32 B per UMMA_K, MN-major +16 rows * 128 B
            const uint64_t da = da0 + (uint64_t)(A_MN ? (kk * 2048) >> 4 : (kk * 32) >> 4);
            const uint64_t db = db0 + (uint64_t)(B_MN ? (kk * 2048) >> 4 : (kk * 32) >> 4);
            umma_bf16_ss(tmem_d, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);  // smem slot reusable once these MMAs retire
        }
        umma_commit(&tmem_full[as]);   // accumulator complete -> epilogue
      }
    }
  } else {
    // ------------------------------------------------------------------ epilogue (8 warps: 4 lane quarters x 2 column halves)
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const unsigned long long seed = p.drop_thresh16 != 0 ? p.seed.value() : 0ull;
    const int half = (warp - 2) >> 2;
    constexpr int CH = BLOCK_N / 64;   // 32-column chunks per warp
    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++tile_it) {
      const int nb = tile % p.n_blocks;
      const int mb = (tile / p.n_blocks) % p.m_blocks;
      const int ks = tile / (p.n_blocks * p.m_blocks);
      const int kb0 = ks * p.k_per_split;
      const int kb1 = min(p.k_blocks, kb0 + p.k_per_split);
      const uint32_t as = tile_it & 1, aph = (tile_it >> 1) & 1;
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      const int row = mb * BLOCK_M + q * 32 + lane;
      const bool row_ok = row < p.M && kb1 > kb0;
      const uint32_t taddr = tmem_base + as * BLOCK_N + (uint32_t(q * 32) << 16);
      epilogue_tile<BLOCK_N>(p, taddr, row, row_ok, nb * BLOCK_N, half * CH, half * CH + CH, p.alpha, seed, ks == 0);
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem_base, C::TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------------------
// TMA-staged epilogue of the CTA-pair kernel (bf16 outputs).  The residual / pre-activation tile arrives in
// shared memory by TMA while the main loop of the tile is still running, every thread combines its
// accumulator row segment with it in place, and the finished [128 x 256] tile leaves through TMA stores:
// all global traffic of the epilogue is full-line and asynchronous (the register path issues 16-byte
// accesses to 32 different rows per instruction and stalls on each load: measured 2-2.5x the time of the
// plain GEMM for the GELU / dGELU variants).
//   chunk layout in the staging tile: sub-tile j = columns [64j, 64j+64), rows of 128 B, 128B swizzle.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t cstage_offset(int r, int col) {   // col multiple of 8
  const int sub = col >> 6, ck = (col & 63) >> 3;
  return (uint32_t)(sub * 16384 + r * 128 + ((ck ^ (r & 7)) << 4));
}
__device__ __forceinline__ void half_bar_sync(int half) {    // the four epilogue warps of one column half (barriers 2, 3)
  asm volatile("bar.sync %0, 128;" ::"r"(half + 2) : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(EPI_THREADS) : "memory"); }

// pass over this warp's 4 chunks; MODE 0: acc (+bias) (+dropout) (+res | *dgelu(res) | tanh) -> staging
//                                 MODE 1: GELU first pass: pre-activation -> staging
//                                 MODE 2: GELU second pass: staging -> gelu(staging)
template <int MODE>
__device__ __forceinline__ void staged_pass(const GemmArgs& p, uint8_t* sC, uint32_t taddr, int r, int row, int n_base,
                                            int c_begin, bool use_res, float alpha, unsigned long long seed) {
#pragma unroll 1
  for (int c = c_begin; c < c_begin + 4; ++c) {
    const int col0 = c * 32;
    const int n0 = n_base + col0;
    if (n0 >= p.N) break;
    float f[32];
    if (MODE == 2) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {
        const uint4 u = *reinterpret_cast<const uint4*>(sC + cstage_offset(r, col0 + g * 8));
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 x = unpack_bf16(w[t]);
          f[g * 8 + 2 * t] = gelu_erf(x.x);
          f[g * 8 + 2 * t + 1] = gelu_erf(x.y);
        }
      }
    } else {
      // the bias loads are issued before the TMEM read so both latencies overlap; acc * alpha + bias is one FFMA
      // per element, and nothing at all when there is no bias and alpha == 1 (the epilogue is instruction bound)
      const int ncols = min(32, p.N - n0);
      uint4 bq[4];
      if (p.bias != nullptr) {
#pragma unroll
        for (int g = 0; g < 4; ++g)
          bq[g] = g * 8 < ncols ? __ldg(reinterpret_cast<const uint4*>(p.bias + n0 + g * 8)) : make_uint4(0, 0, 0, 0);
      }
      uint32_t v[32];
      tmem_ld_32x32(taddr + col0, v);
      tmem_ld_wait();
      if (p.bias != nullptr) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const uint32_t w[4] = {bq[g].x, bq[g].y, bq[g].z, bq[g].w};
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 bb = unpack_bf16(w[t]);
            f[g * 8 + 2 * t] = fmaf(__uint_as_float(v[g * 8 + 2 * t]), alpha, bb.x);
            f[g * 8 + 2 * t + 1] = fmaf(__uint_as_float(v[g * 8 + 2 * t + 1]), alpha, bb.y);
          }
        }
      } else if (alpha != 1.f) {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]) * alpha;
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      }
      if (MODE == 0) {
        if (p.epi == EPI_BIAS_TANH) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = tanhf(f[j]);
        }
        if (p.epi == EPI_BIAS_DROP_RES && p.drop_thresh16 != 0) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint64_t e8 = ((uint64_t)row * (uint64_t)p.N + (uint64_t)(n0 + g * 8)) >> 3;
            const Keep8 keep = dropout_keep8(seed, p.stream, e8, p.drop_thresh16);
#pragma unroll
            for (int t = 0; t < 8; ++t) f[g * 8 + t] = keep[t] ? f[g * 8 + t] * p.drop_scale : 0.f;
          }
        }
        if (use_res) {
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const uint4 u = *reinterpret_cast<const uint4*>(sC + cstage_offset(r, col0 + g * 8));
            const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 rr = unpack_bf16(w[t]);
              if (p.epi == EPI_DGELU) {
                f[g * 8 + 2 * t] *= dgelu_erf(rr.x);
                f[g * 8 + 2 * t + 1] *= dgelu_erf(rr.y);
              } else if (p.epi == EPI_MUL) {
                f[g * 8 + 2 * t] *= rr.x;
                f[g * 8 + 2 * t + 1] *= rr.y;
              } else {
                f[g * 8 + 2 * t] += rr.x;
                f[g * 8 + 2 * t + 1] += rr.y;
              }
            }
          }
        }
      }
    }
#pragma unroll
    for (int g = 0; g < 4; ++g)
      *reinterpret_cast<uint4*>(sC + cstage_offset(r, col0 + g * 8)) =
          make_uint4(pack_bf16(f[g * 8], f[g * 8 + 1]), pack_bf16(f[g * 8 + 2], f[g * 8 + 3]),
                     pack_bf16(f[g * 8 + 4], f[g * 8 + 5]), pack_bf16(f[g * 8 + 6], f[g * 8 + 7]));
  }
}

// Round 2b replacement of staged_pass<0>: the same CTA-wide staging tile and 32-column chunks, with (1) the epilogue
// operation a template parameter (a runtime switch inside the chunk loop cost hundreds of SASS instructions per chunk),
// (2) FMA-class math on packed fp32 pairs, one FFMA2 per pair for alpha / bias whatever the epilogue (absent bias = 0),
// (3) hoisted Philox keys and optional keep-bit output (mask_out), (4) the accumulator released right after the last
// TMEM read.  (A variant with 16-column sub-chunks and a load in flight ahead measured SLOWER here -- 1019 vs 709
// cycles per k-block on K = 1024 tiles: twice the tcgen05.wait::ld round trips, and the extra live registers pushed
// the bias vector into local memory.)
template <int EPI, bool HAS_BIAS>
__device__ __forceinline__ void staged_tile(const GemmArgs& p, uint8_t* sC, uint32_t taddr, int r, int row, int n_base,
                                            int col_begin, bool use_res, float alpha, const PhiloxKeys& keys, uint32_t dropT,
                                            uint32_t tmem_empty_addr, int c_lo = 0, int c_hi = 4, uint32_t* mbits_io = nullptr) {
  // chunks [c_lo, c_hi) of this warp's four 32-column chunks (the box-pipelined epilogue calls it once per 64-column
  // box); the accumulator is released after the read of chunk 3 (or of the last live chunk)
  const int n_w0 = n_base + col_begin;
  const int nch = n_w0 >= p.N ? 0 : min(4, (p.N - n_w0 + 31) >> 5);   // live 32-column chunks (warp uniform)
  const f32x2 alpha2 = f2_splat(alpha);
  const bool drop = EPI == EPI_BIAS_DROP_RES && dropT != 0;
  uint32_t mlocal[4] = {0u, 0u, 0u, 0u};
  uint32_t* mbits = mbits_io != nullptr ? mbits_io : mlocal;   // keep bits of this thread's 128 columns (mask_out / mask_in)
  const bool bits_in = EPI == EPI_BIAS_DROP_RES && drop && p.mask_in != nullptr;
  if (bits_in && c_lo == 0 && row < p.M && n_w0 + 128 <= p.N) {       // 16 bytes = this thread's 128 columns
    const uint4 mw = __ldg(reinterpret_cast<const uint4*>(p.mask_in + (size_t)row * (size_t)(p.N >> 3) + (size_t)(n_w0 >> 3)));
    mbits[0] = mw.x; mbits[1] = mw.y; mbits[2] = mw.z; mbits[3] = mw.w;
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    if (c < c_lo || c >= c_hi) continue;
    if (c >= nch) break;
    const int col0 = col_begin + c * 32, nn = n_base + col0;
    uint4 bq[4];
    if constexpr (HAS_BIAS) {
#pragma unroll
      for (int g = 0; g < 4; ++g)
        bq[g] = nn + g * 8 < p.N ? __ldg(reinterpret_cast<const uint4*>(p.bias + nn + g * 8)) : make_uint4(0, 0, 0, 0);
    }
    uint32_t v[32];
    tmem_ld_32x32(taddr + col0, v);
    tmem_ld_wait_dep(v);
    if (c + 1 >= nch) {                       // last read of the accumulator: the MMA warp may have it back
      tc_fence_before();
      mbar_arrive_cluster(tmem_empty_addr);
    }
#pragma unroll
    for (int g = 0; g < 4; ++g) {
      f32x2 x[4];
      if constexpr (HAS_BIAS) {               // one FFMA2 per pair for alpha and bias
        const uint32_t bw[4] = {bq[g].x, bq[g].y, bq[g].z, bq[g].w};
#pragma unroll
        for (int t = 0; t < 4; ++t)
          x[t] = f2_fma(f2_pack_u(v[g * 8 + 2 * t], v[g * 8 + 2 * t + 1]), alpha2, f2_from_bf16x2(bw[t]));
      } else {                                // the dgrad epilogues: the accumulator as it is (3 of 8 instructions per pair less)
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          x[t] = f2_pack_u(v[g * 8 + 2 * t], v[g * 8 + 2 * t + 1]);
          if (alpha != 1.f) x[t] = f2_mul(x[t], alpha2);
        }
      }
      if constexpr (EPI == EPI_BIAS_TANH) {
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const float2 f = f2_unpack(x[t]);
          x[t] = f2_pack(tanhf(f.x), tanhf(f.y));
        }
      }
      if constexpr (EPI == EPI_BIAS_DROP_RES) {
        if (bits_in) {
          const uint32_t kb = (mbits[c] >> (g * 8)) & 0xffu;
#pragma unroll
          for (int t = 0; t < 4; ++t)
            x[t] = f2_mul(x[t], f2_pack((kb >> (2 * t)) & 1u ? p.drop_scale : 0.f, (kb >> (2 * t + 1)) & 1u ? p.drop_scale : 0.f));
        } else if (drop) {
          const uint64_t e8 = ((uint64_t)row * (uint64_t)p.N + (uint64_t)(nn + g * 8)) >> 3;
          const uint4 rnd = philox7(keys, e8, p.stream);
          uint32_t kb = 0;
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const bool k0 = keep_bit(rnd, 2 * t, dropT), k1 = keep_bit(rnd, 2 * t + 1, dropT);
            x[t] = f2_mul(x[t], f2_pack(k0 ? p.drop_scale : 0.f, k1 ? p.drop_scale : 0.f));
            kb |= (k0 ? 1u : 0u) << (2 * t) | (k1 ? 1u : 0u) << (2 * t + 1);
          }
          mbits[c] |= kb << (g * 8);
        }
      }
      const uint32_t cell = smem_u32(sC) + cstage_offset(r, col0 + g * 8);
      if (use_res) {                          // own row of the residual / multiplier tile, replaced in place
        uint32_t rw[4];
        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(rw[0]), "=r"(rw[1]), "=r"(rw[2]), "=r"(rw[3]) : "r"(cell));
#pragma unroll
        for (int t = 0; t < 4; ++t) {
          const f32x2 rr = f2_from_bf16x2(rw[t]);
          if constexpr (EPI == EPI_MUL) x[t] = f2_mul(x[t], rr);
          else if constexpr (EPI == EPI_DGELU) {
            const float2 f = f2_unpack(x[t]), q = f2_unpack(rr);
            x[t] = f2_pack(f.x * dgelu_erf(q.x), f.y * dgelu_erf(q.y));
          } else x[t] = f2_add(x[t], rr);
        }
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(cell), "r"(f2_to_bf16x2(x[0])), "r"(f2_to_bf16x2(x[1])),
                   "r"(f2_to_bf16x2(x[2])), "r"(f2_to_bf16x2(x[3])) : "memory");
    }
  }
  if (nch == 0 && c_lo == 0) { tc_fence_before(); mbar_arrive_cluster(tmem_empty_addr); }
  if (c_hi == 4 && drop && !bits_in && p.mask_out != nullptr && row < p.M && n_w0 < p.N) {      // 16 bytes = this thread's 128 columns
    unsigned char* dst = p.mask_out + (size_t)row * (size_t)(p.N >> 3) + (size_t)(n_w0 >> 3);
    if (n_w0 + 128 <= p.N) *reinterpret_cast<uint4*>(dst) = make_uint4(mbits[0], mbits[1], mbits[2], mbits[3]);
    else {
      for (int j = 0; j < (p.N - n_w0) >> 3; ++j) dst[j] = (unsigned char)(mbits[j >> 2] >> ((j & 3) * 8));
    }
  }
}

// Column sums of one finished [128 rows x 64 columns] box of the staging tile by the 128 threads of its half: thread t
// owns the column pair (t & 31) over the row quarter (t >> 5).
__device__ __forceinline__ void staged_colsum_box(const GemmArgs& p, const uint8_t* box, int n_box, int t) {
  const int col = (t & 31) * 2, rq = t >> 5;
  const int n = n_box + col;
  if (n >= p.N) return;
  const uint32_t off = (uint32_t)((col & 7) * 2);
  const int ck = col >> 3;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
  for (int r = rq * 32; r < rq * 32 + 32; ++r) {
    const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(box + r * 128 + ((ck ^ (r & 7)) << 4) + off));
    s0 += f.x;
    s1 += f.y;
  }
  atomicAdd(p.colsum + n, s0);
  atomicAdd(p.colsum + n + 1, s1);
}

// Column sums of the finished bf16 tile in the staging buffer (bias gradients fused into a dgrad GEMM): thread t
// owns the column pair (t & 127) over the row half (t >> 7); a warp reads 128 contiguous (swizzle-permuted) bytes of
// one row per step -> conflict free.  Rows beyond M hold zeros (TMA zero-fills the operands), so no row guard.
__device__ __forceinline__ void staged_colsum(const GemmArgs& p, const uint8_t* sC, int n_base, int t) {
  const int col = (t & 127) * 2, rh = t >> 7;
  const int n = n_base + col;
  if (n >= p.N) return;
  const uint32_t base = (uint32_t)((col >> 6) * 16384) + (uint32_t)((col & 7) * 2);
  const int ck = (col & 63) >> 3;
  float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
  for (int r = rh * 64; r < rh * 64 + 64; ++r) {
    const float2 f = unpack_bf16(*reinterpret_cast<const uint32_t*>(sC + base + r * 128 + ((ck ^ (r & 7)) << 4)));
    s0 += f.x;
    s1 += f.y;
  }
  atomicAdd(p.colsum + n, s0);
  atomicAdd(p.colsum + n + 1, s1);
}

// ------------------------------------------------------------------------------------------------
// Warp-local staged epilogue of the pair kernel (round 2; every bf16 epilogue except the legacy two-pass EPI_BIAS_GELU).
// gemm_bench round 2: the CTA-wide staged epilogue cost the residual / dropout variants 11-32 % of the plain GEMM
// (attn-out 515 vs 762 TFLOP/s): one staging tile per CTA means the residual of tile i+1 can only be requested after
// tile i's stores have drained, and three CTA barriers per tile line the eight epilogue warps up behind the slowest.
// Here every epilogue warp owns two [32 rows x 64 columns] boxes (128B swizzle, 8 KB of the staging buffer):
//   * round k (64 columns) of a tile works in box k: the residual box arrives by TMA (per-warp mbarrier), each thread
//     combines its accumulator row segment with its own row of the box IN PLACE, and the elected lane sends the box
//     off with one TMA store;
//   * both residual boxes of the NEXT tile are requested as soon as this tile's stores have been issued, i.e. a whole
//     main loop ahead of their use;
//   * optional column sums (bias gradients) are taken from the finished box by the warp itself and leave as one
//     red.global.add.v2.f32 per lane and round.
// No CTA-wide barrier anywhere.
// ------------------------------------------------------------------------------------------------
struct WarpEpi {
  uint8_t* box;        // this warp's two boxes: box + k * 4096
  uint64_t* rbar;      // this warp's two residual barriers
  int q, half, lane;
};

__device__ __forceinline__ void warp_epi_request_res(const GemmArgs& p, const CUtensorMap* tmap_res, const WarpEpi& w,
                                                     int mn, int rank, int k) {
  const int nb = mn % p.n_blocks, mb = mn / p.n_blocks;
  const int n0 = nb * PAIR_N + w.half * 128 + k * 64;
  if (n0 >= p.N) return;
  mbar_arrive_expect_tx(&w.rbar[k], 4096);
  tma_load_2d(w.box + k * 4096, tmap_res, &w.rbar[k], n0, mb * PAIR_M + rank * 128 + w.q * 32);
}

__device__ __forceinline__ void warp_epi_tile(const GemmArgs& p, const CUtensorMap* tmap_out, const CUtensorMap* tmap_res,
                                              const WarpEpi& w, uint32_t taddr, int mn, int next_mn, int rank, bool use_res,
                                              uint32_t (&res_cnt)[2], float alpha, const PhiloxKeys& keys, uint32_t dropT,
                                              uint32_t tmem_empty_addr) {
  const int nb = mn % p.n_blocks, mb = mn / p.n_blocks;
  const int row_box0 = mb * PAIR_M + rank * 128 + w.q * 32;
  const int row = row_box0 + w.lane;
  const int lane = w.lane;
#pragma unroll 1
  for (int k = 0; k < 2; ++k) {
    const int col0 = w.half * 128 + k * 64;
    const int n0 = nb * PAIR_N + col0;
    const bool live = n0 < p.N;              // warp uniform
    uint8_t* box = w.box + k * 4096;
    if (live) {
      if (use_res) {                          // a box barrier completes once per tile in which the box is live
        mbar_wait(&w.rbar[k], res_cnt[k] & 1);
        ++res_cnt[k];
      } else {                                // the box still belongs to the previous tile's store of round k
        if (lane == 0) tma_store_wait_read<1>();
        __syncwarp();
      }
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int nn = n0 + c * 32;
        uint4 bq[4];
        if (p.bias != nullptr) {
#pragma unroll
          for (int g = 0; g < 4; ++g)
            bq[g] = nn + g * 8 < p.N ? __ldg(reinterpret_cast<const uint4*>(p.bias + nn + g * 8)) : make_uint4(0, 0, 0, 0);
        }
        uint32_t v[32];
        tmem_ld_32x32(taddr + col0 + c * 32, v);
        tmem_ld_wait();
        if (k == 1 && c == 1) {              // this thread is done with the accumulator
          tc_fence_before();
          mbar_arrive_cluster(tmem_empty_addr);
        }
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float f[8];
          if (p.bias != nullptr) {
            const uint32_t bw[4] = {bq[g].x, bq[g].y, bq[g].z, bq[g].w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 bb = unpack_bf16(bw[t]);
              f[2 * t] = fmaf(__uint_as_float(v[g * 8 + 2 * t]), alpha, bb.x);
              f[2 * t + 1] = fmaf(__uint_as_float(v[g * 8 + 2 * t + 1]), alpha, bb.y);
            }
          } else {
#pragma unroll
            for (int t = 0; t < 8; ++t) f[t] = __uint_as_float(v[g * 8 + t]) * alpha;
          }
          if (p.epi == EPI_BIAS_TANH) {
#pragma unroll
            for (int t = 0; t < 8; ++t) f[t] = tanhf(f[t]);
          }
          if (p.epi == EPI_BIAS_DROP_RES && dropT != 0) {
            const uint64_t e8 = ((uint64_t)row * (uint64_t)p.N + (uint64_t)(nn + g * 8)) >> 3;
            const uint4 r = philox7(keys, e8, p.stream);
#pragma unroll
            for (int t = 0; t < 8; ++t) f[t] = keep_bit(r, t, dropT) ? f[t] * p.drop_scale : 0.f;
          }
          uint4* cell = reinterpret_cast<uint4*>(box + lane * 128 + (((c * 4 + g) ^ (lane & 7)) << 4));
          if (use_res) {                      // own row of the residual box, replaced in place by the result
            const uint4 u = *cell;
            const uint32_t rw[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
            for (int t = 0; t < 4; ++t) {
              const float2 rr = unpack_bf16(rw[t]);
              if (p.epi == EPI_MUL) {
                f[2 * t] *= rr.x;
                f[2 * t + 1] *= rr.y;
              } else if (p.epi == EPI_DGELU) {
                f[2 * t] *= dgelu_erf(rr.x);
                f[2 * t + 1] *= dgelu_erf(rr.y);
              } else {
                f[2 * t] += rr.x;
                f[2 * t + 1] += rr.y;
              }
            }
          }
          *cell = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
        }
      }
    } else if (k == 1) {                     // nothing to drain in this round, but the accumulator must be released
      tc_fence_before();
      mbar_arrive_cluster(tmem_empty_addr);
    }
    if (!live) continue;
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(tmap_out, box, n0, row_box0);
      tma_store_commit();
    }
    if (p.colsum != nullptr) {               // lane l owns columns 2l, 2l+1 of the box: 32 rows, conflict free
      const int col = 2 * lane;
      const uint32_t off = (uint32_t)((col & 7) * 2);
      const int ck = col >> 3;
      float s0 = 0.f, s1 = 0.f;
#pragma unroll 8
      for (int r = 0; r < 32; ++r) {
        const float2 x = unpack_bf16(*reinterpret_cast<const uint32_t*>(box + r * 128 + ((ck ^ (r & 7)) << 4) + off));
        s0 += x.x;
        s1 += x.y;
      }
      if (n0 + col < p.N)
        asm volatile("red.global.add.v2.f32 [%0], {%1, %2};" ::"l"(p.colsum + n0 + col), "f"(s0), "f"(s1) : "memory");
    }
  }
  if (use_res && next_mn >= 0 && lane == 0) {   // both residual boxes of the next tile, a whole main loop ahead
    tma_store_wait_read<1>();                   // round 0's store has left box 0
    warp_epi_request_res(p, tmap_res, w, next_mn, rank, 0);
    tma_store_wait_read<0>();
    warp_epi_request_res(p, tmap_res, w, next_mn, rank, 1);
  }
}

// FFN-1 epilogue (EPI_BIAS_GELU_DG): x = acc + bias, out = gelu(x), aux = gelu'(x) -- the activation AND its derivative
// leave the GEMM, so neither a GELU pass nor a GELU' pass over HBM exists any more (the backward epilogue multiplies).
// Entirely warp local: every epilogue warp owns 8 KB of the staging buffer (two [32 rows x 64 columns] 128B-swizzled
// boxes), fills them from its TMEM lane quarter, and its elected lane sends them off with two TMA stores; the only
// wait is for the warp's own previous stores, one round (64 columns of math) earlier.  No CTA-wide barrier.
__device__ __forceinline__ void gelu_dg_warp(const GemmArgs& p, const CUtensorMap* tmap_out, const CUtensorMap* tmap_aux,
                                             uint8_t* sW, uint32_t taddr, int lane, int half, int n_tile0, int row_box0,
                                             float alpha, uint32_t tmem_empty_addr) {
  // Round 2b: (1) the accumulator chunks are software pipelined -- chunk c + 1 is on its way out of TMEM while chunk c is
  // being worked on (two warps per SMSP cannot hide a tcgen05.ld -> wait round trip per chunk otherwise), (2) the math
  // runs on packed fp32 pairs (gelu_dg2: 13 instead of 28 instructions per element; the K = 1024 tile leaves 32).
  const int col_base = half * 128;
  const int n_warp0 = n_tile0 + col_base;
  const f32x2 alpha2 = f2_splat(alpha);
  // 16-column sub-chunks, one in flight ahead of the one being worked on (2 x 16 accumulator registers live)
  const int nsub = n_warp0 >= p.N ? 0 : min(8, (p.N - n_warp0 + 15) >> 4);   // live sub-chunks (warp uniform)
  uint32_t va[16], vb[16];
  if (nsub > 0) tmem_ld_32x16(taddr + col_base, va);
#pragma unroll
  for (int sc = 0; sc < 8; ++sc) {
    if (sc >= nsub) break;
    uint32_t (&v)[16] = (sc & 1) ? vb : va;
    uint32_t (&vn)[16] = (sc & 1) ? va : vb;
    const int nn = n_warp0 + sc * 16;
    uint4 bq[2];
#pragma unroll
    for (int g = 0; g < 2; ++g)
      bq[g] = nn + g * 8 < p.N ? __ldg(reinterpret_cast<const uint4*>(p.bias + nn + g * 8)) : make_uint4(0, 0, 0, 0);
    tmem_ld_wait_dep16(v);
    if (sc + 1 < nsub) tmem_ld_32x16(taddr + col_base + (sc + 1) * 16, vn);
    else {                                     // this thread is done with the accumulator
      tc_fence_before();
      mbar_arrive_cluster(tmem_empty_addr);
    }
    if ((sc & 3) == 0) {                       // first write into this round's boxes: the previous round's stores have read them
      if (lane == 0) tma_store_wait_read<0>();
      __syncwarp();
    }
#pragma unroll
    for (int g = 0; g < 2; ++g) {
      const uint32_t w[4] = {bq[g].x, bq[g].y, bq[g].z, bq[g].w};
      uint32_t gq[4], dq[4];
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const f32x2 x = f2_fma(f2_pack_u(v[g * 8 + 2 * t], v[g * 8 + 2 * t + 1]), alpha2, f2_from_bf16x2(w[t]));
        const GeluDg2 r = gelu_dg2(x);
        gq[t] = f2_to_bf16x2(r.g);
        dq[t] = f2_to_bf16x2(r.dg);
      }
      const uint32_t off = (uint32_t)(lane * 128 + ((((sc & 3) * 2 + g) ^ (lane & 7)) << 4));
      *reinterpret_cast<uint4*>(sW + off) = make_uint4(gq[0], gq[1], gq[2], gq[3]);
      *reinterpret_cast<uint4*>(sW + 4096 + off) = make_uint4(dq[0], dq[1], dq[2], dq[3]);
    }
    if ((sc & 3) == 3 || sc == nsub - 1) {     // box complete: off it goes
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) {
        const int n0 = n_warp0 + (sc >> 2) * 64;
        tma_store_2d(tmap_out, sW, n0, row_box0);
        tma_store_2d(tmap_aux, sW + 4096, n0, row_box0);
        tma_store_commit();
      }
    }
  }
  if (nsub == 0) { tc_fence_before(); mbar_arrive_cluster(tmem_empty_addr); }
}

// ------------------------------------------------------------------------------------------------
// CTA-pair variant: a 2-CTA cluster (the two SMs of a TPC) computes one 256 x 256 tile with
// tcgen05.mma.cta_group::2.  Each CTA stages its own 128 rows of A and its own 128 columns of B per
// k-block (32 KB -> a 6-deep ring) and the tensor cores of both SMs read both halves, so the L2 -> SM
// operand traffic per FLOP is half that of the single-CTA 128 x 256 tile -- which is what limits that
// kernel (ncu: lts throughput ~42%, nothing else saturated).  Only the leader CTA (cluster rank 0)
// issues MMAs; completion is multicast to both CTAs' barriers; both CTAs run the epilogue on their own
// 128 accumulator rows (TMEM lanes).
// ------------------------------------------------------------------------------------------------
constexpr int PAIR_CSTAGE = 128 * PAIR_N * 2;   // bf16 output/residual staging tile of one CTA: 4 x [128 x 64] swizzled boxes
// operand ring: 5 stages next to the 64 KB staging tile of the bf16 epilogues, 7 stages for the fp32 (weight-gradient)
// epilogues that write straight from registers -- same 230 912 bytes either way
constexpr int PAIR_SMEM = 5 * PAIR_STAGE + PAIR_CSTAGE + 1024 + 512;
static_assert(7 * PAIR_STAGE + 1024 + 512 <= PAIR_SMEM, "7-stage ring must fit");

// Work decomposition of the pair kernel.  Classic: cluster c owns the (split, tile) units c, c + C, c + 2C, ...
// Stream-K (accumulating fp32 epilogue only): the m_blocks * n_blocks * k_blocks iteration space is cut into C
// equal contiguous ranges, so every cluster does the same number of MMAs regardless of how the tile count
// divides the machine (64 weight-gradient tiles on 74 CTA pairs would otherwise idle 14 % of the tensor cores);
// a range spans at most a few tiles and each piece is added with red.global.add.
// Tail split (stream_k == 2, fewer tiles T than clusters C, accumulating epilogue): cluster t < T runs k-blocks
// [0, k_main) of tile t -- all of them in lock step from k = 0, so clusters that share an operand panel still hit the
// same L2 lines at the same time (what plain stream-K loses) -- and the C - T otherwise idle clusters split the
// remaining k-blocks [k_main, K) of all tiles evenly among themselves.  64 weight-gradient tiles on 74 CTA pairs:
// 86 % -> ~99 % of the machine busy.
struct WorkSeg { int mn, kb0, kb1; };
__device__ __forceinline__ int seg_count(const GemmArgs& p, int c, int C) {
  if (p.stream_k == 2) {
    const int T = p.m_blocks * p.n_blocks, rt = p.k_blocks - p.k_main;
    if (c < T) return 1;
    const long long R = (long long)T * rt, H = C - T, h = c - T;
    const long long r0 = R * h / H, r1 = R * (h + 1) / H;
    return r1 > r0 ? (int)((r1 - 1) / rt - r0 / rt + 1) : 0;
  }
  if (!p.stream_k) {
    const int total = p.m_blocks * p.n_blocks * p.k_splits;
    return c < total ? (total - c + C - 1) / C : 0;
  }
  const long long total = (long long)p.m_blocks * p.n_blocks * p.k_blocks;
  const long long it0 = total * c / C, it1 = total * (c + 1) / C;
  return it1 > it0 ? (int)((it1 - 1) / p.k_blocks - it0 / p.k_blocks + 1) : 0;
}
__device__ __forceinline__ WorkSeg seg_get(const GemmArgs& p, int c, int C, int idx) {
  WorkSeg w;
  if (p.stream_k == 2) {
    const int T = p.m_blocks * p.n_blocks, rt = p.k_blocks - p.k_main;
    if (c < T) { w.mn = c; w.kb0 = 0; w.kb1 = p.k_main; return w; }
    const long long R = (long long)T * rt, H = C - T, h = c - T;
    const long long r0 = R * h / H, r1 = R * (h + 1) / H;
    const long long start = idx == 0 ? r0 : (r0 / rt + idx) * rt;
    const int inside = (int)(start % rt);
    w.mn = (int)(start / rt);
    w.kb0 = p.k_main + inside;
    w.kb1 = p.k_main + (int)min((long long)rt, inside + (r1 - start));
    return w;
  }
  if (!p.stream_k) {
    const int tile = c + idx * C;
    const int mnt = p.m_blocks * p.n_blocks;
    const int ks = tile / mnt;
    w.mn = tile % mnt;
    w.kb0 = ks * p.k_per_split;
    w.kb1 = min(p.k_blocks, w.kb0 + p.k_per_split);
    return w;
  }
  const long long total = (long long)p.m_blocks * p.n_blocks * p.k_blocks;
  const long long it0 = total * c / C, it1 = total * (c + 1) / C;
  const long long start = idx == 0 ? it0 : (it0 / p.k_blocks + idx) * p.k_blocks;
  w.mn = (int)(start / p.k_blocks);
  w.kb0 = (int)(start % p.k_blocks);
  w.kb1 = (int)min((long long)p.k_blocks, w.kb0 + (it1 - start));
  return w;
}

// The measurement knobs (GemmArgs::lab, tools/gemm_lab.py) exist only in builds with -DB200_GEMM_LAB
// (B200_NVCC_EXTRA=-DB200_GEMM_LAB python -m bert_pytorch_b200.ops.build): the handful of predicated instructions they
// add to the single-thread TMA-producer and MMA-issuer loops cost the weight-gradient GEMM 19 % in situ (62.9 -> 74.8 us)
// -- those two loops are the critical path of the kernel, every cycle of their period is a cycle the tensor cores wait.
#ifdef B200_GEMM_LAB
#define LABV(p) ((p).lab)
#define LABSTATS(p) ((p).lab_stats)
// phase clock of the CTA-wide staged epilogue (first epilogue thread of the leader CTA): lab_t[i] accumulates the cycles
// between mark i - 1 and mark i; written behind the MMA issuer's records (74 x 4) as 74 x 8 words
#define LAB_T(i) do { if (lab_clk) { const long long now_ = clock64(); lab_t[i] += now_ - lab_last; lab_last = now_; } } while (0)
#else
#define LABV(p) 0u
#define LABSTATS(p) ((unsigned long long*)nullptr)
#define LAB_T(i) do { } while (0)
#endif
// EC = epilogue class: every class is its own kernel, so the register allocation (168 per thread is the ceiling) and
// the instruction footprint of one epilogue do not pay for the others (the single runtime-switched kernel spilled).
enum { EC_F32 = 0, EC_GELU_DG = 1, EC_STAGED = 2 };
// WIDE (fp32 epilogues only): a 256 x 512 tile per cluster -- two N = 256 MMAs per k-step share the A operand, the two
// accumulators fill all 512 TMEM columns (no double buffering: a weight-gradient cluster owns one or two tiles with
// K = 12288, there is nothing to overlap the epilogue with).  Operand bytes per FLOP drop by a quarter (48 KB instead of
// 2 x 32 KB per 2 x 512 MMA cycles and CTA) -- which is what the TN kernel is short of: its MN-major TMA boxes arrive at
// 46 B/clk per SM (gemm_lab 'nomma': 710 cycles per 32 KB stage, the same on 18 or 74 clusters), the 256 x 256 tile
// needs 64.  Ring: 4 stages of 48 KB.
template <bool A_MN, bool B_MN, bool FP8, int EC, bool WIDE = false>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(PAIR_THREADS, 1)   // 11 warps: 3 on three of the SMSPs -> 168 registers
gemm_pair_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const __grid_constant__ CUtensorMap tmap_out, const __grid_constant__ CUtensorMap tmap_aux,
                 const __grid_constant__ CUtensorMap tmap_res, const GemmArgs p) {
  static_assert(!WIDE || (EC == EC_F32 && !FP8), "the 256 x 512 tile serves the fp32 (weight-gradient) epilogues of the bf16 path");
  constexpr int PAIR_STAGES = EC == EC_F32 ? (WIDE ? 4 : 7) : 5;
  constexpr int PAIR_STAGE = WIDE ? 49152 : b200::PAIR_STAGE;        // bytes per CTA and stage: A 16 KB + B 16 / 32 KB
  constexpr int TILE_N = WIDE ? 512 : PAIR_N;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sC = smem + PAIR_STAGES * PAIR_STAGE;          // staging tile (1024-aligned; absent in the 7-stage variant)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(sC + (PAIR_STAGES == 5 ? PAIR_CSTAGE : 0));
  uint64_t* empty_bar = full_bar + PAIR_STAGES;
  uint64_t* tmem_full = empty_bar + PAIR_STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint64_t* c_full = tmem_empty + 2;                      // residual tile landed in the staging buffer
  uint64_t* res_bar = c_full + 1;                         // [8 epilogue warps][2 boxes]: warp-local residual boxes landed
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(res_bar + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // Warp roles.  The SMSP arbiter prefers the HIGHEST warp id among its ready warps (B300_MICROARCH.md), so the two
  // single-thread roles that feed the tensor cores -- the TMA producer and the MMA issuer -- sit in the last two warps
  // (8, 9) and the epilogue in warps 0..7: with the roles the other way round (round 1) every burst of epilogue math
  // delayed the issue loop (gemm_lab: K = 1024 tiles ran at 709 cycles per k-block with, 573 without the epilogue,
  // although the MMA issuer never waited for a free accumulator).  lab bit8 restores the old order for A/B runs.
  // TWO producer warps (8: even iterations of the ring, 10: odd ones).  One producer's period is wait(empty) -> expect_tx
  // -> 2..4 TMA instructions, ~130 + 175..290 cycles per instruction (gemm_lab 'nomma': 530 / 650 / 710 / 820 cycles per
  // stage for 2 K-major / 1 + 1 / 2 three-dimensional / 4 two-dimensional MN-major boxes, the same on 18 and on 74
  // clusters): above the 512 cycles the MMAs of a stage take for every layout but K-major.  Two threads overlap their
  // issue latencies.
  constexpr int w_tma = 8, w_mma = 9, w_tma2 = 10;
  const int ew = warp;                                   // epilogue warp index 0..7 (meaningless for the feeders)
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;
  // elements per k-block: one 128-byte swizzle row of the operand type.  The fp8 variant moves the same bytes
  // per stage as bf16 (16 KB of A + 16 KB of B per CTA) and the same 32 bytes of K per MMA, at twice the FLOPs.
  constexpr int BK = FP8 ? 128 : BLOCK_K;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < PAIR_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 2 * EPI_THREADS);   // the epilogue threads of both CTAs of the pair
    }
    mbar_init(c_full, 1);
    for (int i = 0; i < 16; ++i) mbar_init(&res_bar[i], 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc_2sm(tmem_base_slot, 512);
  tc_fence_before();
  cluster_sync_all();                    // peer barriers are initialised before anything remote touches them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;
  if (!(LABV(p) & 512u)) {
    pdl_trigger();                       // the next kernel's CTAs may take this SM the moment this CTA leaves
    pdl_wait();                          // PDL: everything above overlapped the previous kernel's tail
  }

  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const int nseg = seg_count(p, cluster_id, nclusters);

  if (warp == w_tma || warp == w_tma2) {
    if (lane == 0) {                     // ---------------- TMA producers (both CTAs)
      const uint32_t my_parity = warp == w_tma ? 0u : 1u;
      uint32_t it = 0;
      int s = 0;                           // ring slot and phase of iteration `it`, kept incrementally: every instruction
      uint32_t ph = 0;                     // of this loop is on the critical path of the kernel
      for (int si = 0; si < nseg; ++si) {
        const WorkSeg w = seg_get(p, cluster_id, nclusters, si);
        const int nb = w.mn % p.n_blocks, mb = w.mn / p.n_blocks;
        const int kb0 = w.kb0, kb1 = w.kb1;
        const int m0 = mb * PAIR_M + (int)rank * 128, n0 = nb * TILE_N + (int)rank * 128;
        const int pf = (int)((LABV(p) >> 8) & 0xffu);
        for (int kb = kb0; kb < kb1; ++kb, ++it, ph ^= (++s == PAIR_STAGES) ? 1u : 0u, s = (s == PAIR_STAGES) ? 0 : s) {
          if ((it & 1u) != my_parity) continue;               // the other producer's slot
          mbar_wait(&empty_bar[s], ph ^ 1);
          uint8_t* sa = smem + s * PAIR_STAGE;
          uint8_t* sb = sa + 16384;
          if ((LABV(p) & 1u) && it >= (uint32_t)PAIR_STAGES) {          // lab: no refill, the MMAs re-read stale operands
            if (leader) mbar_arrive(&full_bar[s]);
            continue;
          }
          if (leader) mbar_arrive_expect_tx(&full_bar[s], 2 * PAIR_STAGE);
          const uint32_t fb = mapa_shared(smem_u32(&full_bar[s]), 0);
          int kbe = kb;
          if (LABV(p) & 16u) {                                           // lab: every cluster walks K from its own offset
            const int span = kb1 - kb0, off = (cluster_id * (int)((LABV(p) >> 16) & 0xffu)) % span;
            kbe = kb0 + (kb - kb0 + off) % span;
          }
          const int kc = (LABV(p) & 2u) ? (kb & 1) * BK : kbe * BK;
          const int ma = (LABV(p) & 2u) ? (int)rank * 128 : m0, na = (LABV(p) & 2u) ? (int)rank * 128 : n0;
          if (pf > 0 && kb + pf < kb1 && !p.mn3d) {                               // pull the operands of k-block kb + pf into L2
            const int kp = (kb + pf) * BK;
            if constexpr (!A_MN) tma_prefetch_2d(&tmap_a, kp, ma);
            else { tma_prefetch_2d(&tmap_a, ma, kp); if constexpr (!FP8) tma_prefetch_2d(&tmap_a, ma + 64, kp); }
            if constexpr (!B_MN) tma_prefetch_2d(&tmap_b, kp, na);
            else { tma_prefetch_2d(&tmap_b, na, kp); if constexpr (!FP8) tma_prefetch_2d(&tmap_b, na + 64, kp); }
          }
          if constexpr (!A_MN) {
            tma_load_2d_2sm(sa, &tmap_a, fb, kc, ma);
          } else if constexpr (FP8) {
            tma_load_2d_2sm(sa, &tmap_a, fb, ma, kc);          // 128 MN bytes x 128 K rows: one box
          } else if (p.mn3d) {                                // one 3-D box [2][64 k][64 mn] instead of two 2-D boxes: half the TMA
                                                              // operations per stage (gemm_lab: 816 -> 702 cycles per k-block, TN)
            tma_load_3d_2sm(sa, &tmap_a, fb, 0, kc, ma >> 6);
          } else {
            tma_load_2d_2sm(sa, &tmap_a, fb, ma, kc);
            tma_load_2d_2sm(sa + 8192, &tmap_a, fb, ma + 64, kc);
          }
#pragma unroll
          for (int j = 0; j < (WIDE ? 2 : 1); ++j) {          // WIDE: this CTA's 128 columns of both N = 256 halves
            uint8_t* sbj = sb + j * 16384;
            const int naj = na + j * 256;
            if constexpr (!B_MN) {
              tma_load_2d_2sm(sbj, &tmap_b, fb, kc, naj);
            } else if constexpr (FP8) {
              tma_load_2d_2sm(sbj, &tmap_b, fb, naj, kc);
            } else if (p.mn3d) {
              tma_load_3d_2sm(sbj, &tmap_b, fb, 0, kc, naj >> 6);
            } else {
              tma_load_2d_2sm(sbj, &tmap_b, fb, naj, kc);
              tma_load_2d_2sm(sbj + 8192, &tmap_b, fb, naj + 64, kc);
            }
          }
        }
      }
    }
  } else if (warp == w_mma) {
    if (lane == 0 && leader) {           // ---------------- MMA issuer (leader CTA only)
      const uint32_t idesc = FP8 ? umma_idesc_fp8(PAIR_M, PAIR_N, A_MN, B_MN, p.fp8_fmt & 1u, (p.fp8_fmt >> 1) & 1u)
                                 : umma_idesc_bf16(PAIR_M, PAIR_N, A_MN, B_MN);
      // bytes between consecutive MMAs along K inside one stage: K-major 32 B (16 bf16 / 32 fp8); MN-major one
      // MMA-K worth of rows (16 x 128 B for bf16, 32 x 128 B for fp8)
      constexpr uint32_t A_KSTEP = A_MN ? (FP8 ? 4096u : 2048u) : 32u;
      constexpr uint32_t B_KSTEP = B_MN ? (FP8 ? 4096u : 2048u) : 32u;
      uint32_t it = 0, tile_it = 0, ph = 0;
      int s = 0;
      const bool stats = LABSTATS(p) != nullptr;
      long long t_begin = 0, w_full = 0, w_tmem = 0;
      if (stats) t_begin = clock64();
      for (int si = 0; si < nseg; ++si, ++tile_it) {
        const WorkSeg w = seg_get(p, cluster_id, nclusters, si);
        const int kb0 = w.kb0, kb1 = w.kb1;
        const uint32_t as = WIDE ? 0u : (tile_it & 1), aph = WIDE ? (tile_it & 1) : ((tile_it >> 1) & 1);
        if (stats) { const long long t0 = clock64(); mbar_wait(&tmem_empty[as], aph ^ 1); w_tmem += clock64() - t0; }
        else mbar_wait(&tmem_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + as * PAIR_N;
        for (int kb = kb0; kb < kb1; ++kb, ++it, ph ^= (++s == PAIR_STAGES) ? 1u : 0u, s = (s == PAIR_STAGES) ? 0 : s) {
          if (stats) { const long long t0 = clock64(); mbar_wait(&full_bar[s], ph); w_full += clock64() - t0; }
          else mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * PAIR_STAGE);
          const uint32_t sb = sa + 16384;
          const uint64_t da0 = A_MN ? umma_smem_desc_sw128(sa, 8192, 1024) : umma_smem_desc_sw128(sa, 16, 1024);
          const uint64_t db0 = B_MN ? umma_smem_desc_sw128(sb, 8192, 1024) : umma_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < BLOCK_K / UMMA_K; ++kk) {
            if (LABV(p) & 64u) break;                 // lab: no MMAs at all -> the loop runs at the operand supply rate
            const uint64_t da = da0 + (uint64_t)((kk * A_KSTEP) >> 4);
            const uint64_t db = db0 + (uint64_t)((kk * B_KSTEP) >> 4);
            if constexpr (FP8) umma_fp8_ss_2sm(tmem_d, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
            else umma_bf16_ss_2sm(tmem_d, da, db, idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
            if constexpr (WIDE)                        // second N = 256 half: same A, the B tile 16 KB further on
              umma_bf16_ss_2sm(tmem_d + 256, da, db + (uint64_t)(16384 >> 4), idesc, (kb > kb0 || kk > 0) ? 1u : 0u);
          }
          umma_commit_2sm(&empty_bar[s], 3);     // both CTAs may refill this slot
        }
        umma_commit_2sm(&tmem_full[as], 3);      // both CTAs' epilogues may drain
      }
      if (stats) {
        unsigned long long* o = LABSTATS(p) + 4 * cluster_id;
        o[0] = (unsigned long long)(clock64() - t_begin); o[1] = (unsigned long long)w_full;
        o[2] = (unsigned long long)w_tmem; o[3] = it;
      }
    }
  } else {                               // ---------------- epilogue (both CTAs, own 128 rows)
    const int q = warp & 3;                  // TMEM lane quarter of this warp (hardware: warp id % 4)
    const int half = ew >> 2;
    const bool staged = EC != EC_F32 && !(LABV(p) & 128u);   // lab bit7: register path
    const bool use_res = EC == EC_STAGED && staged && p.res != nullptr &&
                         (p.epi == EPI_BIAS_DROP_RES || p.epi == EPI_ADD || p.epi == EPI_DGELU || p.epi == EPI_MUL);
    const bool issuer = ew == 0 && lane == 0;   // first epilogue thread drives the staging tile's TMA traffic
    const int r = q * 32 + lane;             // row inside the CTA tile == TMEM lane
    const unsigned long long seed = p.drop_thresh16 != 0 ? p.seed.value() : 0ull;
    float alpha = p.alpha;
    if constexpr (FP8) alpha *= __ldg(p.scale_a) * __ldg(p.scale_b);   // per-tensor dequantisation
    auto load_res = [&](int tile) {
      const int nb = tile % p.n_blocks;
      const int mb = (tile / p.n_blocks) % p.m_blocks;
      mbar_arrive_expect_tx(c_full, PAIR_CSTAGE);
#pragma unroll
      for (int j = 0; j < 4; ++j)
        tma_load_2d(sC + j * 16384, &tmap_res, c_full, nb * PAIR_N + j * 64, mb * PAIR_M + (int)rank * 128);
    };
    // warp-local staged epilogue (opt-in: B200_GEMM_WARP_EPI=1)
    const bool warp_local = EC == EC_STAGED && staged && p.warp_epi != 0 && p.epi != EPI_BIAS_GELU;
    WarpEpi we;
    we.box = sC + ew * 8192; we.rbar = res_bar + ew * 2; we.q = q; we.half = half; we.lane = lane;
    const PhiloxKeys keys = philox_keys(seed);
    const uint32_t dropT = p.drop_thresh16 << 16;
    uint32_t res_cnt[2] = {0u, 0u};
    if constexpr (EC == EC_STAGED) {
      if (warp_local) {
        if (use_res && nseg > 0 && lane == 0) {
          const int mn0 = seg_get(p, cluster_id, nclusters, 0).mn;
          warp_epi_request_res(p, &tmap_res, we, mn0, (int)rank, 0);
          warp_epi_request_res(p, &tmap_res, we, mn0, (int)rank, 1);
        }
      } else if (use_res && nseg > 0 && lane == 0 && (ew & 3) == 0) {   // the elected thread of each half: its two boxes
        const int mn0 = seg_get(p, cluster_id, nclusters, 0).mn;
        const int nb0 = mn0 % p.n_blocks, row00 = (mn0 / p.n_blocks) * PAIR_M + (int)rank * 128;
#pragma unroll
        for (int b = 0; b < 2; ++b) {
          const int box = half * 2 + b, n_box = nb0 * PAIR_N + box * 64;
          if (n_box < p.N) {
            mbar_arrive_expect_tx(&res_bar[box], 16384);
            tma_load_2d(sC + box * 16384, &tmap_res, &res_bar[box], n_box, row00);
          }
        }
      }
    }
    uint32_t tile_it = 0, res_phase = 0;
#ifdef B200_GEMM_LAB
    const bool lab_clk = LABSTATS(p) != nullptr && issuer && leader;
    long long lab_t[8] = {0, 0, 0, 0, 0, 0, 0, 0}, lab_last = lab_clk ? clock64() : 0;
#endif
    for (int si = 0; si < nseg; ++si, ++tile_it) {
      const WorkSeg wseg = seg_get(p, cluster_id, nclusters, si);
      const int mn = wseg.mn;
      const int nb = mn % p.n_blocks, mb = mn / p.n_blocks;
      const uint32_t as = WIDE ? 0u : (tile_it & 1), aph = WIDE ? (tile_it & 1) : ((tile_it >> 1) & 1);
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      const int row0 = mb * PAIR_M + (int)rank * 128;
      const int row = row0 + r;
      const uint32_t taddr = tmem_base + as * PAIR_N + (uint32_t(q * 32) << 16);
      if (LABV(p) & 12u) {                       // lab: bit2 = no epilogue at all, bit3 = drain the accumulator, store nothing
        if (LABV(p) & 8u) {
          uint32_t acc = 0;
#pragma unroll 1
          for (int c = half * 4; c < half * 4 + 4; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(taddr + c * 32, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc ^= v[j];
          }
          if (acc == 0x12345678u && LABSTATS(p) != nullptr) LABSTATS(p)[0] = acc;
        }
        tc_fence_before();
        mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[as]), 0));
        continue;
      }
      if (!staged) {
        if constexpr (EC != EC_GELU_DG) {
#pragma unroll 1
          for (int j = 0; j < (WIDE ? 2 : 1); ++j)
            epilogue_tile<PAIR_N>(p, taddr + j * 256, row, row < p.M, nb * TILE_N + j * 256, half * 4, half * 4 + 4, alpha, seed,
                                  wseg.kb0 == 0);
        }
        tc_fence_before();
        mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[as]), 0));
        continue;
      }
      if constexpr (EC == EC_GELU_DG) {      // warp-local staging + stores, no CTA-wide barrier (see gelu_dg_warp)
        gelu_dg_warp(p, &tmap_out, &tmap_aux, sC + ew * 8192, taddr, lane, half, nb * PAIR_N, row0 + q * 32, alpha,
                     mapa_shared(smem_u32(&tmem_empty[as]), 0));
      }
      if constexpr (EC == EC_STAGED) {
        if (warp_local) {
          const int next_mn = si + 1 < nseg ? seg_get(p, cluster_id, nclusters, si + 1).mn : -1;
          warp_epi_tile(p, &tmap_out, &tmap_res, we, taddr, mn, next_mn, (int)rank, use_res, res_cnt, alpha, keys, dropT,
                        mapa_shared(smem_u32(&tmem_empty[as]), 0));
          continue;
        }
        if (p.epi != EPI_BIAS_GELU) {
          // ---- box-pipelined epilogue (default).  The staging tile is four [128 x 64] boxes; the four warps of a column
          // half own two of them and work through them one after the other: residual box landed -> combine in place ->
          // half barrier -> TMA store.  The store of box 0 drains, and the residual of the NEXT tile's box 0 streams in,
          // while box 1 is being worked on -- the CTA-wide version (round 2a) waited 2.0 k cycles per tile for the store
          // to read the tile and only then asked for the next residual (gemm_lab phase clock): K = 1024 GEMMs with
          // three tiles per cluster ran at first main loop + 3 x epilogue.
          const uint32_t te = mapa_shared(smem_u32(&tmem_empty[as]), 0);
          const int nbase = nb * PAIR_N, cb = half * 128;
          const bool elected = (ew & 3) == 0 && lane == 0;             // first thread of this half
          const int tq = (ew & 3) * 32 + lane;                         // thread index inside the half
          uint32_t mbits[4] = {0u, 0u, 0u, 0u};
#pragma unroll 1
          for (int b = 0; b < 2; ++b) {
            const int box = half * 2 + b;
            const int n_box = nbase + box * 64;
            const bool live = n_box < p.N;                             // uniform over the half
            LAB_T(0);
            if (use_res) {
              if (live) mbar_wait(&res_bar[box], res_phase);
            } else {
              half_bar_sync(half);                                     // the elected thread has seen this box's previous store read it
            }
            LAB_T(1);
#define B200_STAGED(EPI_, BIAS_) staged_tile<EPI_, BIAS_>(p, sC, taddr, r, row, nbase, cb, use_res, alpha, keys, dropT, te, 2 * b, 2 * b + 2, mbits)
            const bool hb = p.bias != nullptr;
            switch (p.epi) {
              case EPI_BIAS_DROP_RES: B200_STAGED(EPI_BIAS_DROP_RES, true); break;
              case EPI_BIAS_TANH: B200_STAGED(EPI_BIAS_TANH, true); break;
              case EPI_MUL: if (hb) B200_STAGED(EPI_MUL, true); else B200_STAGED(EPI_MUL, false); break;
              case EPI_DGELU: if (hb) B200_STAGED(EPI_DGELU, true); else B200_STAGED(EPI_DGELU, false); break;
              default: if (hb) B200_STAGED(EPI_ADD, true); else B200_STAGED(EPI_ADD, false); break;   // none / bias / add
            }
#undef B200_STAGED
            LAB_T(2);
            fence_proxy_async();
            half_bar_sync(half);                                       // the box is complete
            LAB_T(3);
            if (elected && live) {
              tma_store_2d(&tmap_out, sC + box * 16384, n_box, row0);
              tma_store_commit();
            }
            if (p.colsum != nullptr && live) staged_colsum_box(p, sC + box * 16384, n_box, tq);   // while the store drains
          }
          LAB_T(4);
          res_phase ^= 1;
          // the column sums READ the boxes after their stores went out: nobody may ask for the next residual (a TMA write
          // into the same box) before every thread of the half is done with them (found by running the tests under
          // compute-sanitizer's timing: profiles/compute_sanitizer_racecheck_r2b.log)
          if (p.colsum != nullptr && use_res) half_bar_sync(half);
          if (elected) {                                               // boxes free again -> next tile's residual boxes
            const bool more = use_res && si + 1 < nseg;
            int nnb = 0, nrow0 = 0;
            if (more) {
              const int nmn = seg_get(p, cluster_id, nclusters, si + 1).mn;
              nnb = nmn % p.n_blocks;
              nrow0 = (nmn / p.n_blocks) * PAIR_M + (int)rank * 128;
            }
#pragma unroll
            for (int b = 0; b < 2; ++b) {
              // box 0's store is the older of two bulk groups only if box 1 was stored too
              if (b == 0 && nbase + (half * 2 + 1) * 64 < p.N) tma_store_wait_read<1>(); else tma_store_wait_read<0>();
              const int box = half * 2 + b, n_box = nnb * PAIR_N + box * 64;
              if (more && n_box < p.N) {
                mbar_arrive_expect_tx(&res_bar[box], 16384);
                tma_load_2d(sC + box * 16384, &tmap_res, &res_bar[box], n_box, nrow0);
              }
            }
          }
          LAB_T(5);
          continue;
        }
        // ---- legacy two-pass EPI_BIAS_GELU (B200_FUSED_GELU=0): CTA-wide staging
        staged_pass<1>(p, sC, taddr, r, row, nb * PAIR_N, half * 4, false, alpha, seed);
        tc_fence_before();
        mbar_arrive_cluster(mapa_shared(smem_u32(&tmem_empty[as]), 0));   // accumulator drained: MMA may reuse it
        fence_proxy_async();
        epi_bar_sync();
        if (issuer) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (nb * PAIR_N + j * 64 < p.N) tma_store_2d(&tmap_aux, sC + j * 16384, nb * PAIR_N + j * 64, row0);
          tma_store_commit();
          tma_store_wait_read<0>();
        }
        epi_bar_sync();
        staged_pass<2>(p, sC, taddr, r, row, nb * PAIR_N, half * 4, false, alpha, seed);
        fence_proxy_async();
        epi_bar_sync();
        if (issuer) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (nb * PAIR_N + j * 64 < p.N) tma_store_2d(&tmap_out, sC + j * 16384, nb * PAIR_N + j * 64, row0);
          tma_store_commit();
        }
        if (p.colsum != nullptr) staged_colsum(p, sC, nb * PAIR_N, ew * 32 + lane);   // while the stores drain
        if (issuer) tma_store_wait_read<0>();                 // staging tile free again
        epi_bar_sync();                                       // nobody touches the staging tile before that
      }
    }
    // shared memory must outlive the TMA stores' READS of it; their global writes are complete when the grid is (the
    // .read form is what CUTLASS epilogues wait for too) -- the full wait_group cost every bf16 GEMM its last ~1 us
    if (EC == EC_GELU_DG || warp_local) {
      if (lane == 0) tma_store_wait_read<0>();
    } else if (staged && lane == 0 && (ew & 3) == 0) {       // the elected thread of each half (ew == 0: also the legacy issuer)
      tma_store_wait_read<0>();
    }
#ifdef B200_GEMM_LAB
    if (lab_clk) {
      LAB_T(7);
      unsigned long long* o = LABSTATS(p) + 74 * 4 + 8 * cluster_id;
      for (int i = 0; i < 8; ++i) o[i] = (unsigned long long)lab_t[i];
    }
#endif
  }

  tc_fence_before();
  cluster_sync_all();                    // nobody exits while the peer can still signal or read its smem
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc_2sm(tmem_base, 512);
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || f == nullptr) {
      fprintf(stderr, "[b200] cuTensorMapEncodeTiled is unavailable (%s)\n", cudaGetErrorString(e));
      abort();
    }
    fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// 2D bf16 tensor map: `inner` contiguous elements per row, `outer` rows, row pitch `ld` elements,
// box = [box_outer][box_inner], 128B swizzle (box_inner must be 64 bf16).
static CUtensorMap make_tmap_2d(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                                uint32_t box_outer, uint32_t esize);

CUtensorMap make_tmap_2d_bf16(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                              uint32_t box_outer) {
  return make_tmap_2d(ptr, inner, outer, ld, box_inner, box_outer, 2);
}

// `esize` 2: bf16 (box_inner 64), 1: fp8 bytes (box_inner 128)
static CUtensorMap make_tmap_2d(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                                uint32_t box_outer, uint32_t esize) {
  struct Key {
    const void* p; uint64_t i, o, l; uint32_t bi, bo, es;
    bool operator==(const Key& k) const {
      return p == k.p && i == k.i && o == k.o && l == k.l && bi == k.bi && bo == k.bo && es == k.es;
    }
  };
  struct Hash {
    size_t operator()(const Key& k) const {
      size_t h = reinterpret_cast<size_t>(k.p);
      h = h * 1000003u ^ k.i; h = h * 1000003u ^ k.o; h = h * 1000003u ^ k.l; h = h * 1000003u ^ k.bi;
      h = h * 1000003u ^ k.es;
      return h * 1000003u ^ k.bo;
    }
  };
  static std::unordered_map<Key, CUtensorMap, Hash> cache;
  static std::mutex mu;
  Key key{ptr, inner, outer, ld, box_inner, box_outer, esize};
  std::lock_guard<std::mutex> lock(mu);
  auto itf = cache.find(key);
  if (itf != cache.end()) return itf->second;
  CUtensorMap m;
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * esize};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = get_encode_fn()(&m, esize == 2 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box,
                               estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    fprintf(stderr, "[b200] cuTensorMapEncodeTiled failed (%d): ptr=%p inner=%llu outer=%llu ld=%llu box=%ux%u\n",
            (int)r, ptr, (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld, box_inner,
            box_outer);
    abort();
  }
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  return m;
}

// MN-major bf16 operand as ONE box per CTA and stage: dims {64 mn (contiguous), K rows, MN / 64 groups}, box {64, BK, 2}
// -> shared memory [2][BK][64], the layout the two 2-D boxes produce.  MN must be a multiple of 64.
static CUtensorMapL2promotion mn_promotion() {     // measurement knob: B200_TMAP_PROMO=0..3 (none / 64 / 128 / 256 B)
  static const int v = []() { const char* e = getenv("B200_TMAP_PROMO"); return e ? atoi(e) : 3; }();
  return v == 0 ? CU_TENSOR_MAP_L2_PROMOTION_NONE : v == 1 ? CU_TENSOR_MAP_L2_PROMOTION_L2_64B
       : v == 2 ? CU_TENSOR_MAP_L2_PROMOTION_L2_128B : CU_TENSOR_MAP_L2_PROMOTION_L2_256B;
}
static CUtensorMap make_tmap_mn3d(const void* ptr, uint64_t mn, uint64_t k, uint64_t ld, uint32_t bk) {
  struct Key {
    const void* p; uint64_t mn, k, l; uint32_t bk;
    bool operator==(const Key& o) const { return p == o.p && mn == o.mn && k == o.k && l == o.l && bk == o.bk; }
  };
  struct Hash {
    size_t operator()(const Key& q) const {
      size_t h = reinterpret_cast<size_t>(q.p);
      h = h * 1000003u ^ q.mn; h = h * 1000003u ^ q.k; h = h * 1000003u ^ q.l;
      return h * 1000003u ^ q.bk;
    }
  };
  static std::unordered_map<Key, CUtensorMap, Hash> cache;
  static std::mutex mu;
  const Key key{ptr, mn, k, ld, bk};
  std::lock_guard<std::mutex> lock(mu);
  auto itf = cache.find(key);
  if (itf != cache.end()) return itf->second;
  CUtensorMap m;
  cuuint64_t dims[3] = {64, k, mn / 64};
  cuuint64_t strides[2] = {ld * 2, 128};
  cuuint32_t box[3] = {64, bk, 2};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = get_encode_fn()(&m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), dims, strides, box, estr,
                               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, mn_promotion(),
                               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { fprintf(stderr, "[b200] 3-D tensor map failed (%d)\n", (int)r); abort(); }
  if (cache.size() > 4096) cache.clear();
  cache.emplace(key, m);
  return m;
}

CUtensorMap make_tmap_2d_u8(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner,
                            uint32_t box_outer) {
  return make_tmap_2d(ptr, inner, outer, ld, box_inner, box_outer, 1);
}

static int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev;
    B200_CUDA_CHECK(cudaGetDevice(&dev));
    B200_CUDA_CHECK(cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev));
  }
  return n;
}

static unsigned int g_lab = []() { const char* e = getenv("B200_GEMM_LAB"); return e ? (unsigned int)strtoul(e, nullptr, 0) : 0u; }();
static unsigned long long* g_lab_stats = nullptr;
void gemm_lab(unsigned int flags, unsigned long long* stats) { g_lab = flags; g_lab_stats = stats; }

static void fill_peer(GemmArgs& p, const GemmCall& c) {
  p.lab = 0; p.lab_stats = nullptr;
  p.peer_world = c.peer_world; p.peer_rank = c.peer_rank; p.peer_push = c.peer_push;
  p.peer_off = c.peer_off; p.peer_per = c.peer_per;
  for (int i = 0; i < 16; ++i) p.peer_base[i] = i < c.peer_world ? c.peer_base[i] : nullptr;
}

template <bool A_MN, bool B_MN, int BLOCK_N>
static void launch(const GemmCall& c, cudaStream_t st) {
  using C = Cfg<BLOCK_N>;
  GemmArgs p;
  p.M = c.M; p.N = c.N; p.K = c.K;
  p.m_blocks = (c.M + BLOCK_M - 1) / BLOCK_M;
  p.n_blocks = (c.N + BLOCK_N - 1) / BLOCK_N;
  p.k_blocks = (c.K + BLOCK_K - 1) / BLOCK_K;
  p.k_splits = c.k_splits < 1 ? 1 : c.k_splits;
  if (p.k_splits > p.k_blocks) p.k_splits = p.k_blocks;
  p.k_per_split = (p.k_blocks + p.k_splits - 1) / p.k_splits;
  p.k_splits = (p.k_blocks + p.k_per_split - 1) / p.k_per_split;   // no empty splits
  p.epi = c.epi;
  p.out = c.out; p.ldo = c.ldo; p.aux_out = reinterpret_cast<__nv_bfloat16*>(c.aux_out);
  p.bias = reinterpret_cast<const __nv_bfloat16*>(c.bias);
  p.res = reinterpret_cast<const __nv_bfloat16*>(c.res); p.ldr = c.ldr;
  p.colsum = c.colsum; p.mask_out = nullptr; p.mask_in = nullptr;
  p.seed = Seed{c.seed, c.seed_step}; p.stream = c.stream;
  float pd = c.p_drop;
  p.drop_thresh16 = pd > 0.f ? (unsigned)(pd * 65536.f + 0.5f) : 0u;
  p.drop_scale = pd > 0.f ? 65536.f / (65536.f - (float)p.drop_thresh16) : 1.f;
  p.alpha = c.alpha;
  fill_peer(p, c);
  p.warp_epi = 0; p.mn3d = 0;
  p.scale_a = nullptr; p.scale_b = nullptr; p.fp8_fmt = 0; p.stream_k = 0; p.k_main = 0;
  // operand maps
  CUtensorMap ta = A_MN ? make_tmap_2d_bf16(c.A, c.M, c.K, c.lda, 64, BLOCK_K)
                        : make_tmap_2d_bf16(c.A, c.K, c.M, c.lda, BLOCK_K, BLOCK_M);
  CUtensorMap tb = B_MN ? make_tmap_2d_bf16(c.B, c.N, c.K, c.ldb, 64, BLOCK_K)
                        : make_tmap_2d_bf16(c.B, c.K, c.N, c.ldb, BLOCK_K, BLOCK_N);
  auto kern = gemm_kernel<A_MN, B_MN, BLOCK_N>;
  static bool configured = false;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM_TOTAL));
    configured = true;
  }
  const int tiles = p.m_blocks * p.n_blocks * p.k_splits;
  const int grid = tiles < num_sms() ? tiles : num_sms();
  if (grid <= 0) return;
  kern<<<grid, NUM_THREADS, C::SMEM_TOTAL, st>>>(ta, tb, p);
}

template <bool A_MN, bool B_MN, bool FP8>
static void launch_pair(const GemmCall& c, cudaStream_t st) {
  constexpr int BK = FP8 ? 128 : BLOCK_K;
  const bool wide = !FP8 && c.block_n == 1024 && (c.epi == EPI_ACCUM_F32 || c.epi == EPI_F32) && c.N % 512 == 0 &&
                    c.M % 64 == 0;                  // 256 x 512 tiles (aligned problems only; everything else: 256 x 256)
  GemmArgs p;
  p.M = c.M; p.N = c.N; p.K = c.K;
  p.m_blocks = (c.M + PAIR_M - 1) / PAIR_M;
  p.n_blocks = wide ? (c.N + 511) / 512 : (c.N + PAIR_N - 1) / PAIR_N;
  p.k_blocks = (c.K + BK - 1) / BK;
  p.k_splits = c.k_splits < 1 ? 1 : c.k_splits;
  if (p.k_splits > p.k_blocks) p.k_splits = p.k_blocks;
  p.k_per_split = (p.k_blocks + p.k_splits - 1) / p.k_splits;
  p.k_splits = (p.k_blocks + p.k_per_split - 1) / p.k_per_split;
  p.epi = c.epi;
  p.out = c.out; p.ldo = c.ldo; p.aux_out = reinterpret_cast<__nv_bfloat16*>(c.aux_out);
  p.bias = reinterpret_cast<const __nv_bfloat16*>(c.bias);
  p.res = reinterpret_cast<const __nv_bfloat16*>(c.res); p.ldr = c.ldr;
  p.colsum = c.colsum; p.mask_out = c.mask_out; p.mask_in = c.mask_in;
  p.seed = Seed{c.seed, c.seed_step}; p.stream = c.stream;
  float pd = c.p_drop;
  p.drop_thresh16 = pd > 0.f ? (unsigned)(pd * 65536.f + 0.5f) : 0u;
  p.drop_scale = pd > 0.f ? 65536.f / (65536.f - (float)p.drop_thresh16) : 1.f;
  p.alpha = c.alpha;
  fill_peer(p, c);
  // k_splits < 0 asks for stream-K (fp32 accumulate epilogue only): equal MMA work per cluster, pieces merged by
  // red.global.add.  Needs enough k-blocks per cluster to amortise the extra partial tiles.
  p.stream_k = 0; p.k_main = 0;
  if (c.k_splits == -2 && c.epi == EPI_ACCUM_F32) {       // tail split: see seg_count
    const int T = p.m_blocks * p.n_blocks, pairs_ = num_sms() / 2;
    const int kmain = (int)((long long)p.k_blocks * T / pairs_);
    if (T < pairs_ && kmain >= 4 && p.k_blocks - kmain >= 1) {
      p.stream_k = 2; p.k_main = kmain; p.k_splits = 2;     // k_splits > 1: the epilogue adds with red.global.add
    } else {
      p.k_splits = 1; p.k_per_split = p.k_blocks;
    }
  } else if (c.k_splits < 0 && c.epi == EPI_ACCUM_F32) {
    const long long iters = (long long)p.m_blocks * p.n_blocks * p.k_blocks;
    const int pairs_ = num_sms() / 2;
    if (iters >= 8ll * pairs_ && (p.m_blocks * p.n_blocks) % pairs_ != 0) {
      p.stream_k = 1;
      p.k_splits = 2;               // "more than one contributor per tile": the epilogue uses red.global.add
    }
  }
  p.scale_a = c.scale_a; p.scale_b = c.scale_b;
  p.lab = g_lab; p.lab_stats = g_lab_stats;
  p.fp8_fmt = (c.a_e5m2 ? 1u : 0u) | (c.b_e5m2 ? 2u : 0u);
  if (FP8 && (c.scale_a == nullptr || c.scale_b == nullptr)) {
    fprintf(stderr, "[b200] fp8 gemm needs the device dequantisation factors of both operands\n");
    abort();
  }
  // each CTA loads 128-row boxes of A and 128-column boxes of B (one 128-byte swizzle row wide)
  constexpr uint32_t ES = FP8 ? 1 : 2, BI = FP8 ? 128 : 64;
  CUtensorMap ta = A_MN ? make_tmap_2d(c.A, c.M, c.K, c.lda, BI, BK, ES)
                        : make_tmap_2d(c.A, c.K, c.M, c.lda, BK, 128, ES);
  CUtensorMap tb = B_MN ? make_tmap_2d(c.B, c.N, c.K, c.ldb, BI, BK, ES)
                        : make_tmap_2d(c.B, c.K, c.N, c.ldb, BK, 128, ES);
  p.mn3d = 0;
  if (!FP8 && (A_MN || B_MN) && !(p.lab & 32u) && !((A_MN && c.M % 64) || (B_MN && c.N % 64))) {   // lab bit5: 2-D boxes
    p.mn3d = 1;
    if (A_MN) ta = make_tmap_mn3d(c.A, c.M, c.K, c.lda, BK);
    if (B_MN) tb = make_tmap_mn3d(c.B, c.N, c.K, c.ldb, BK);
  }
  const bool f32_out = c.epi == EPI_ACCUM_F32 || c.epi == EPI_F32;
  // output / pre-activation / residual tiles of one CTA travel as 4 boxes of [128 rows x 64 columns]
  // (EPI_BIAS_GELU_DG: every epilogue warp stores its own [32 rows x 64 columns] boxes)
  // opt-in warp-local staged epilogue: B200_GEMM_WARP_EPI=1 (every layout) / =nt (forward GEMMs only)
  static const int warp_epi_on = []() { const char* e = getenv("B200_GEMM_WARP_EPI"); return !e ? 0 : (e[0] == '1' ? 1 : (e[0] == 'n' ? 2 : 0)); }();
  p.warp_epi = ((warp_epi_on == 1 || (warp_epi_on == 2 && !A_MN && !B_MN)) && c.epi != EPI_BIAS_GELU) ? 1 : 0;
  const uint32_t obox = (c.epi == EPI_BIAS_GELU_DG || p.warp_epi) ? 32 : 128;
  CUtensorMap to = f32_out ? ta : make_tmap_2d_bf16(c.out, c.N, c.M, c.ldo, 64, obox);
  CUtensorMap tx = (c.aux_out != nullptr && !f32_out) ? make_tmap_2d_bf16(c.aux_out, c.N, c.M, c.ldo, 64, obox) : to;
  CUtensorMap tr = (c.res != nullptr && !f32_out) ? make_tmap_2d_bf16(c.res, c.N, c.M, c.ldr, 64, obox) : to;
  auto kern = f32_out ? gemm_pair_kernel<A_MN, B_MN, FP8, EC_F32>
              : c.epi == EPI_BIAS_GELU_DG ? gemm_pair_kernel<A_MN, B_MN, FP8, EC_GELU_DG>
                                          : gemm_pair_kernel<A_MN, B_MN, FP8, EC_STAGED>;
  if constexpr (!FP8) {
    if (wide) kern = gemm_pair_kernel<A_MN, B_MN, false, EC_F32, true>;
  }
  static bool configured = false;
  if (!configured) {
    if constexpr (!FP8)
      B200_CUDA_CHECK(cudaFuncSetAttribute(gemm_pair_kernel<A_MN, B_MN, false, EC_F32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
    B200_CUDA_CHECK(cudaFuncSetAttribute(gemm_pair_kernel<A_MN, B_MN, FP8, EC_F32>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
    B200_CUDA_CHECK(cudaFuncSetAttribute(gemm_pair_kernel<A_MN, B_MN, FP8, EC_GELU_DG>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
    B200_CUDA_CHECK(cudaFuncSetAttribute(gemm_pair_kernel<A_MN, B_MN, FP8, EC_STAGED>, cudaFuncAttributeMaxDynamicSharedMemorySize, PAIR_SMEM));
    configured = true;
  }
  const int tiles = p.m_blocks * p.n_blocks * p.k_splits;
  int pairs = num_sms() / 2;
  if ((p.lab >> 24) != 0 && (int)(p.lab >> 24) < pairs) pairs = (int)(p.lab >> 24);   // lab: fewer clusters (per-SM vs chip-wide supply)
  const int grid = p.stream_k ? 2 * pairs : 2 * (tiles < pairs ? tiles : pairs);
  if (grid <= 0) return;
  launch_pdl(kern, dim3(grid), dim3(PAIR_THREADS), PAIR_SMEM, st, ta, tb, to, tx, tr, p);
}

void gemm_bf16(const GemmCall& c, cudaStream_t st) {
  if (c.fp8 && c.block_n != 512 && c.block_n != 1024) {
    fprintf(stderr, "[b200] fp8 operands are implemented on the CTA-pair kernel only (block_n=512)\n");
    abort();
  }
  if (c.block_n == 512 || c.block_n == 1024) {   // CTA-pair 256 x 256 tiles (1024: 256 x 512 for the fp32 epilogues)
    switch (c.layout) {
      case GEMM_NT: c.fp8 ? launch_pair<false, false, true>(c, st) : launch_pair<false, false, false>(c, st); return;
      case GEMM_NN: c.fp8 ? launch_pair<false, true, true>(c, st) : launch_pair<false, true, false>(c, st); return;
      case GEMM_TN: c.fp8 ? launch_pair<true, true, true>(c, st) : launch_pair<true, true, false>(c, st); return;
      default: fprintf(stderr, "[b200] bad gemm layout %d\n", c.layout); abort();
    }
  }
  const bool wide = c.block_n == 256;
  switch (c.layout) {
    case GEMM_NT: wide ? launch<false, false, 256>(c, st) : launch<false, false, 128>(c, st); break;
    case GEMM_NN: wide ? launch<false, true, 256>(c, st) : launch<false, true, 128>(c, st); break;
    case GEMM_TN: wide ? launch<true, true, 256>(c, st) : launch<true, true, 128>(c, st); break;
    default: fprintf(stderr, "[b200] bad gemm layout %d\n", c.layout); abort();
  }
}

}  // namespace b200
