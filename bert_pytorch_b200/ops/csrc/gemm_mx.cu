// Block-scaled fp8 GEMM (OCP MXFP8: e4m3 elements, one ue8m0 scale per 32 elements along K) on the 5th-generation
// tensor cores:  D[M,N] = (A * 2^sfa)[M,K] . (B * 2^sfb)[N,K]^T   with  tcgen05.mma.kind::mxf8f6f4.block_scale.
//
// The scale factors never touch the CUDA cores: they travel global -> shared memory (cp.async.bulk, 512 B per
// 128-row x 128-K block, already in the layout the tensor core wants) -> TMEM (tcgen05.cp, SASS UTCCP) and the
// MMA applies them per 32-element K group (the `sf_id` fields of the instruction descriptor select the byte).
// One CTA per 128 x 128 tile, 128 fp8 = one 128-byte swizzle row per k-block, double-buffered accumulator
// (2 x 128 TMEM columns) + 8 scale-factor columns, warp 0 = TMA / bulk-copy producer, warp 1 = UTCCP + MMA issuer,
// warps 2..9 = epilogue (bias / fp32 / bf16 output through the shared epilogue of gemm_sm100.cu's 1-CTA kernel is
// not reused here to keep this path self-contained: bf16 or fp32 output, optional bias).
//
// Scale-factor storage ("atom" layout of cutlass::detail::Sm1xxBlockScaledBasicChunk, K-major):
//   sf[(mb * KB + kb) * 512 + (r % 32) * 16 + (r / 32) * 4 + j]   for row r of 128-row block mb, K group j of k-block kb
// mx_quantize_kernel below writes exactly that.
#include "common.cuh"
#include "gemm_sm100.h"
#include "kernels.h"

namespace b200 {

namespace mx {
constexpr int BM = 128, BN = 128, BK = 128;          // BK in fp8 elements = bytes
constexpr int A_BYTES = BM * BK, B_BYTES = BN * BK;  // 16 KB each
constexpr int SF_BYTES = 512;                         // per 128 rows x 4 K groups
constexpr int STAGE = A_BYTES + B_BYTES + 2 * 1024;   // SFA / SFB slots padded to 1 KB
constexpr int STAGES = 5;
constexpr int THREADS = 320;
constexpr int SMEM = STAGES * STAGE + 1024 + 256;
}  // namespace mx

struct MxArgs {
  int M, N, K, m_blocks, n_blocks, k_blocks;
  void* out; int ldo; int out_f32;
  const __nv_bfloat16* bias;
  const uint8_t* sfa; const uint8_t* sfb;
};

// no-swizzle K-major descriptor of a 32 x 16 B scale-factor block: 8-row groups 128 B apart
__device__ __forceinline__ uint64_t sf_smem_desc(uint32_t smem_addr) {
  return (uint64_t)((smem_addr & 0x3FFFF) >> 4) | ((uint64_t)(128 >> 4) << 32) | ((uint64_t)1 << 46);
}
// instruction descriptor of kind::mxf8f6f4.block_scale (cute::UMMA::InstrDescriptorBlockScaled):
//   [4,6) B sf id  [7,10) A fmt (0 = e4m3)  [10,13) B fmt  [15]/[16] majors (0 = K)  [17,23) N >> 3  [23] scale fmt (1 = ue8m0)
//   [24,29) M >> 4  [29,31) A sf id
__device__ __forceinline__ uint32_t mx_idesc(uint32_t M, uint32_t N, uint32_t sf_id) {
  return (sf_id << 4) | ((N >> 3) << 17) | (1u << 23) | ((M >> 4) << 24) | (sf_id << 29);
}
__device__ __forceinline__ void tmem_cp_sf(uint32_t taddr, uint64_t sdesc) {
  asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(taddr), "l"(sdesc) : "memory");
}
__device__ __forceinline__ void umma_mx_ss(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate,
                                           uint32_t tsfa, uint32_t tsfb) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}"
      ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate), "r"(tsfa), "r"(tsfb)
      : "memory");
}
__device__ __forceinline__ void bulk_load(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(smem_dst)), "l"(gsrc), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(mx::THREADS, 1)
gemm_mx_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const MxArgs p) {
  using namespace mx;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full[s], 1); mbar_init(&tmem_empty[s], 256); }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *tmem_slot;
  const uint32_t t_sfa = tmem + 256, t_sfb = tmem + 260;      // 4 columns each (one per 32-row group)
  const int total = p.m_blocks * p.n_blocks;

  if (warp == 0) {
    if (lane == 0) {
      uint32_t it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int nb = tile % p.n_blocks, mb = tile / p.n_blocks;
        for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&empty_bar[s], ((it / STAGES) & 1) ^ 1);
          uint8_t* sa = smem + s * STAGE;
          uint8_t* sb = sa + A_BYTES;
          uint8_t* ssfa = sb + B_BYTES;
          uint8_t* ssfb = ssfa + 1024;
          mbar_arrive_expect_tx(&full_bar[s], A_BYTES + B_BYTES + 2 * SF_BYTES);
          tma_load_2d(sa, &tmap_a, &full_bar[s], kb * BK, mb * BM);
          tma_load_2d(sb, &tmap_b, &full_bar[s], kb * BK, nb * BN);
          bulk_load(ssfa, p.sfa + ((size_t)mb * p.k_blocks + kb) * SF_BYTES, SF_BYTES, &full_bar[s]);
          bulk_load(ssfb, p.sfb + ((size_t)nb * p.k_blocks + kb) * SF_BYTES, SF_BYTES, &full_bar[s]);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, tile_it = 0;
      for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tile_it) {
        const uint32_t as = tile_it & 1, aph = (tile_it >> 1) & 1;
        mbar_wait(&tmem_empty[as], aph ^ 1);
        tc_fence_after();
        const uint32_t tmem_d = tmem + as * BN;
        for (int kb = 0; kb < p.k_blocks; ++kb, ++it) {
          const int s = it % STAGES;
          mbar_wait(&full_bar[s], (it / STAGES) & 1);
          tc_fence_after();
          const uint32_t sa = smem_u32(smem + s * STAGE), sb = sa + A_BYTES, ssfa = sb + B_BYTES, ssfb = ssfa + 1024;
          // scale factors: shared memory -> TMEM; tcgen05.cp and tcgen05.mma of one thread execute in issue order
          tmem_cp_sf(t_sfa, sf_smem_desc(ssfa));
          tmem_cp_sf(t_sfb, sf_smem_desc(ssfb));
          const uint64_t da0 = umma_smem_desc_sw128(sa, 16, 1024), db0 = umma_smem_desc_sw128(sb, 16, 1024);
#pragma unroll
          for (int kk = 0; kk < BK / 32; ++kk)           // one MMA per 32-element scale group
            umma_mx_ss(tmem_d, da0 + (uint64_t)((kk * 32) >> 4), db0 + (uint64_t)((kk * 32) >> 4), mx_idesc(BM, BN, kk),
                       (kb > 0 || kk > 0) ? 1u : 0u, t_sfa, t_sfb);
          umma_commit(&empty_bar[s]);
        }
        umma_commit(&tmem_full[as]);
      }
    }
  } else {
    const int q = warp & 3, half = (warp - 2) >> 2;
    uint32_t tile_it = 0;
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x, ++tile_it) {
      const int nb = tile % p.n_blocks, mb = tile / p.n_blocks;
      const uint32_t as = tile_it & 1, aph = (tile_it >> 1) & 1;
      mbar_wait(&tmem_full[as], aph);
      tc_fence_after();
      const int row = mb * BM + q * 32 + lane;
      const uint32_t taddr = tmem + as * BN + (uint32_t(q * 32) << 16);
#pragma unroll 1
      for (int c = half * 2; c < half * 2 + 2; ++c) {
        const int n0 = nb * BN + c * 32;
        if (n0 >= p.N) break;
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        const int ncols = min(32, p.N - n0);
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.bias != nullptr) {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (j < ncols) f[j] += __bfloat162float(p.bias[n0 + j]);
        }
        if (row < p.M) {
          if (p.out_f32) {
            float* o = reinterpret_cast<float*>(p.out) + (size_t)row * p.ldo + n0;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (j < ncols) *reinterpret_cast<float4*>(o + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
          } else {
            __nv_bfloat16* o = reinterpret_cast<__nv_bfloat16*>(p.out) + (size_t)row * p.ldo + n0;
#pragma unroll
            for (int g = 0; g < 4; ++g)
              if (g * 8 < ncols)
                *reinterpret_cast<uint4*>(o + g * 8) =
                    make_uint4(pack_bf16(f[g * 8], f[g * 8 + 1]), pack_bf16(f[g * 8 + 2], f[g * 8 + 3]),
                               pack_bf16(f[g * 8 + 4], f[g * 8 + 5]), pack_bf16(f[g * 8 + 6], f[g * 8 + 7]));
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tmem_empty[as]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    tmem_dealloc(tmem, 512);
  }
}

// x: bf16 [R, K] row-major (K % 128 == 0) -> q: e4m3 bytes [R, K], sf: ue8m0 in the atom layout above
// ([ceil(R/128)][K/128][512]).  One thread per (row, 32-element group): 64 B in, 32 B + 1 B out.
__global__ void __launch_bounds__(256) mx_quantize_kernel(const __nv_bfloat16* __restrict__ x, uint8_t* __restrict__ q,
                                                          uint8_t* __restrict__ sf, int R, int K) {
  const int groups = K / 32;
  const long long gid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const int Rp = (R + 127) / 128 * 128;
  if (gid >= (long long)Rp * groups) return;
  const int row = (int)(gid / groups), g = (int)(gid % groups);
  const int kb = g >> 2, j = g & 3;
  const size_t sf_idx = ((size_t)(row >> 7) * (K / 128) + kb) * 512 + (row & 31) * 16 + ((row >> 5) & 3) * 4 + j;
  if (row >= R) { sf[sf_idx] = 127; return; }            // padding rows: scale 1, data comes in as TMA zero fill
  const uint4* src = reinterpret_cast<const uint4*>(x + (size_t)row * K + g * 32);
  float v[32];
  float amax = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint4 u = __ldg(src + i);
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 f = unpack_bf16(w[t]);
      v[i * 8 + 2 * t] = f.x; v[i * 8 + 2 * t + 1] = f.y;
      amax = fmaxf(amax, fmaxf(fabsf(f.x), fabsf(f.y)));
    }
  }
  // smallest power of two s with amax / s <= 448 (e4m3 max); ue8m0 stores log2(s) + 127
  int e = 127;
  if (amax > 0.f && isfinite(amax)) {
    e = (int)ceilf(log2f(amax * (1.0f / 448.0f))) + 127;
    e = max(1, min(254, e));
  }
  const float inv = exp2f((float)(127 - e));
  sf[sf_idx] = (uint8_t)e;
  uint32_t o[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) o[i] = fp8_cvt4(v[4 * i] * inv, v[4 * i + 1] * inv, v[4 * i + 2] * inv, v[4 * i + 3] * inv, false);
  uint4* dst = reinterpret_cast<uint4*>(q + (size_t)row * K + g * 32);
  dst[0] = make_uint4(o[0], o[1], o[2], o[3]);
  dst[1] = make_uint4(o[4], o[5], o[6], o[7]);
}

void mx_quantize(const void* x, void* q, void* sf, int R, int K, cudaStream_t st) {
  const long long work = (long long)((R + 127) / 128 * 128) * (K / 32);
  mx_quantize_kernel<<<(unsigned)((work + 255) / 256), 256, 0, st>>>((const __nv_bfloat16*)x, (uint8_t*)q, (uint8_t*)sf, R, K);
}

CUtensorMap make_tmap_2d_u8(const void* ptr, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_inner, uint32_t box_outer);

void gemm_mxfp8(const void* a, const void* sfa, const void* b, const void* sfb, void* out, int ldo, bool out_f32,
                const void* bias, int M, int N, int K, cudaStream_t st) {
  using namespace mx;
  MxArgs p;
  p.M = M; p.N = N; p.K = K;
  p.m_blocks = (M + BM - 1) / BM; p.n_blocks = (N + BN - 1) / BN; p.k_blocks = K / BK;
  p.out = out; p.ldo = ldo; p.out_f32 = out_f32 ? 1 : 0;
  p.bias = (const __nv_bfloat16*)bias;
  p.sfa = (const uint8_t*)sfa; p.sfb = (const uint8_t*)sfb;
  CUtensorMap ta = make_tmap_2d_u8(a, K, M, K, BK, BM);
  CUtensorMap tb = make_tmap_2d_u8(b, K, N, K, BK, BN);
  static bool configured = false;
  if (!configured) {
    B200_CUDA_CHECK(cudaFuncSetAttribute(gemm_mx_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM));
    configured = true;
  }
  int dev, sms;
  B200_CUDA_CHECK(cudaGetDevice(&dev));
  B200_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int tiles = p.m_blocks * p.n_blocks;
  const int grid = tiles < sms ? tiles : sms;
  if (grid > 0) gemm_mx_kernel<<<grid, THREADS, SMEM, st>>>(ta, tb, p);
}

}  // namespace b200
