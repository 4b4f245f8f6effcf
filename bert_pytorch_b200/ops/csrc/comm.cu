// Fused gradient all-reduce + unscale + partitioned LAMB + parameter all-gather over NVLink peer memory.
// ONE persistent cooperative kernel per optimizer step and rank; no NCCL call on this path
// (algorithm spec: parallel/sharded_lamb.py).
//
// Every rank maps every peer's gradient arena, fp32 parameter arena, bf16 shadow arena and a small
// scratch/flag pad (CUDA VMM symmetric memory; the handles are exchanged by the host once).  Rank r owns the
// contiguous shard [lo_r, hi_r) of the flat arena.  Phases (grid barriers inside the kernel, cross-GPU
// barriers through release/acquire flags in peer memory):
//   0  cross-GPU barrier: every rank's backward has finished writing its gradient arena
//   A  reduce-scatter: g[i] = (sum_p grad_p[i]) / (world * loss_scale) for i in the own shard, read from the
//      peers with plain P2P loads (or one multimem.ld_reduce through the NVLS multicast mapping);
//      inf/nan flag + partial sum of squares
//   1  all ranks exchange {sum g^2, found_inf} (P2P stores into the peers' pads) -> global grad norm; a set
//      found_inf makes every rank skip the update identically (GradScaler semantics, decided on device)
//   B  LAMB moments + update direction u on the shard, partial per-tensor ||p||^2, ||u||^2 (tensors may
//      straddle shard boundaries -> per-tensor partials)
//   2  exchange of the per-tensor partial norms
//   C  p -= lr * trust_ratio * u on the shard; new fp32 values and bf16 shadow are stored straight into every
//      peer's arenas (all-gather by P2P / multimem stores); the whole local gradient arena is zeroed
//   3  cross-GPU barrier: all pushes have landed before anyone's next forward reads the weights
#include "common.cuh"
#include "kernels.h"

namespace b200 {

constexpr int FUSED_THREADS = 512;
constexpr int MAX_WORLD = 16;

struct FusedLambArgs {
  int rank, world, use_multicast;
  float* grad[MAX_WORLD];                 // peers' gradient arenas (grad[rank] is local)
  float* param[MAX_WORLD];                // peers' fp32 parameter arenas
  __nv_bfloat16* shadow[MAX_WORLD];       // peers' bf16 shadow arenas
  float* pad[MAX_WORLD];                  // peers' scratch pads: [world][2 + 2*T] floats of exchange space
  unsigned int* flags[MAX_WORLD];         // peers' flag arrays: [4][world] epoch counters
  float* grad_mc;                         // multicast mappings (nullptr when unavailable)
  float* param_mc;
  __nv_bfloat16* shadow_mc;
  float* m;                               // local moments (arena sized; only the shard is touched)
  float* v;
  long long numel, lo, hi;
  const int* chunk_tensor;                // chunk table of the arena restricted to [lo, hi)
  const long long* chunk_start;
  const int* chunk_len;
  int nchunks, ntensors;
  const int* decay_flag;
  const int* prereduced;                  // per tensor: gradient already summed into the owner's arena (or null)
  float* stats;                           // local [4]: sumsq, found_inf, global sumsq, global inf
  float* norms;                           // local [2T] partial norms
  unsigned int* grid_bar;                 // local grid barrier counter (zeroed before launch)
  unsigned int epoch;
  float grad_mul;                         // 1 / (world * loss_scale)
  float lr, beta1, beta2, beta3, eps, weight_decay, bc1, bc2, max_grad_norm;
  int adam_w_mode, use_nvlamb;
  int push_master;                        // 0: fp32 master stays shard-local (peers get only the bf16 shadow)
};

__device__ __forceinline__ void grid_sync(unsigned int* bar, unsigned int& gen) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1u);
    const unsigned int target = (gen + 1) * gridDim.x;
    const uint64_t t0 = globaltimer_ns();
    unsigned int cur;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(bar) : "memory");
      if (globaltimer_ns() - t0 > B200_WATCHDOG_NS) { printf("[b200] grid barrier watchdog (block %d)\n", blockIdx.x); __trap(); }
    } while (cur < target);
  }
  __syncthreads();
  ++gen;
}

// full barrier across the ranks, executed by block 0 (callers follow it with a grid_sync)
__device__ __forceinline__ void peer_barrier(const FusedLambArgs& a, int slot) {
  if (blockIdx.x == 0) {
    __syncthreads();
    if (threadIdx.x < a.world) {
      const int peer = threadIdx.x;
      __threadfence_system();
      unsigned int* dst = a.flags[peer] + slot * a.world + a.rank;
      asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(dst), "r"(a.epoch) : "memory");
      const unsigned int* src = a.flags[a.rank] + slot * a.world + peer;
      const uint64_t t0 = globaltimer_ns();
      unsigned int cur;
      do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(cur) : "l"(src) : "memory");
        if (globaltimer_ns() - t0 > 4 * B200_WATCHDOG_NS) {
          printf("[b200] peer barrier watchdog: rank %d waits for rank %d slot %d epoch %u (saw %u)\n", a.rank, peer,
                 slot, a.epoch, cur);
          __trap();
        }
      } while ((int)(cur - a.epoch) < 0);
    }
    __syncthreads();
  }
}

__device__ __forceinline__ float block_sum_f(float v, float* sm) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sm[warp] = v;
  __syncthreads();
  float r = 0.f;
  if (warp == 0) {
    r = lane < (blockDim.x >> 5) ? sm[lane] : 0.f;
    r = warp_sum(r);
  }
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(FUSED_THREADS, 1) fused_allreduce_lamb_kernel(const FusedLambArgs a) {
  __shared__ float sm[32];
  __shared__ float s_bcast[4];
  unsigned int gen = 0;
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  float* lgrad = a.grad[a.rank];
  float* lparam = a.param[a.rank];

  // in-kernel timeline: block 0 / thread 0 stamps %globaltimer after every phase into stats[8..15]
  // (milliseconds since kernel entry): [8] wait for the slowest rank (barrier 0), [9] phase A (reduce-scatter),
  // [10] sync 1, [11] phase B, [12] sync 2, [13] phase C (apply + all-gather), [14] zero + final barrier, [15] total
  const unsigned long long t_enter = globaltimer_ns();
  unsigned long long t_prev = t_enter;
  auto stamp = [&](int slot) {
    if (gtid == 0) {
      const unsigned long long now = globaltimer_ns();
      a.stats[slot] = (float)(now - t_prev) * 1e-6f;
      t_prev = now;
    }
  };

  // ---- phase 0: everyone's gradients are complete
  if (gtid < 2 * a.ntensors) a.norms[gtid] = 0.f;
  if (gtid < 2) a.stats[gtid] = 0.f;
  peer_barrier(a, 0);
  grid_sync(a.grid_bar, gen);
  stamp(8);

  // ---- phase A: reduce-scatter the own shard
  // Tensors flagged `prereduced` already hold the sum over the ranks in the owner's arena: their weight-gradient
  // GEMMs of the last micro-step added every tile (plus the locally accumulated value) straight into the owner's
  // arena over NVLink (gemm_sm100.cu, peer_push) -- the reduce-scatter happened inside the backward pass.
  {
    float sq = 0.f;
    bool bad = false;
    auto reduce4 = [&](long long off, bool local_only) -> float4 {
      float4 acc;
      if (local_only) {
        acc = *reinterpret_cast<const float4*>(lgrad + off);
      } else if (a.use_multicast) {
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(acc.x), "=f"(acc.y), "=f"(acc.z), "=f"(acc.w) : "l"(a.grad_mc + off) : "memory");
      } else {
        // four peers' loads in flight per thread before the first add (NVLink latency is ~2 us: a dependent
        // load-add chain leaves the links idle); fixed rank order keeps the sum bitwise identical on every rank
        acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
        for (int p0 = 0; p0 < a.world; p0 += 4) {
          float4 g[4];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            g[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p0 + j < a.world)
              asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                           : "=f"(g[j].x), "=f"(g[j].y), "=f"(g[j].z), "=f"(g[j].w) : "l"(a.grad[p0 + j] + off) : "memory");
          }
#pragma unroll
          for (int j = 0; j < 4; ++j) { acc.x += g[j].x; acc.y += g[j].y; acc.z += g[j].z; acc.w += g[j].w; }
        }
      }
      acc.x *= a.grad_mul; acc.y *= a.grad_mul; acc.z *= a.grad_mul; acc.w *= a.grad_mul;
      const float q = acc.x * acc.x + acc.y * acc.y + acc.z * acc.z + acc.w * acc.w;
      bad |= !isfinite(q);
      sq += q;
      *reinterpret_cast<float4*>(lgrad + off) = acc;   // own shard: no peer reads this region
      return acc;
    };
    if (a.prereduced == nullptr) {
      const long long n4 = (a.hi - a.lo) >> 2;
      for (long long i = gtid; i < n4; i += gstride) reduce4(a.lo + (i << 2), false);
    } else {
      for (int c = blockIdx.x; c < a.nchunks; c += gridDim.x) {
        const long long off = a.chunk_start[c];
        const int n = a.chunk_len[c];
        const bool local_only = a.prereduced[a.chunk_tensor[c]] != 0;
        const int nv = (off & 3) ? 0 : (n >> 2);
        for (int i = threadIdx.x; i < nv; i += blockDim.x) reduce4(off + 4 * i, local_only);
        for (int i = 4 * nv + threadIdx.x; i < n; i += blockDim.x) {      // odd tails (e.g. the 2-element NSP bias)
          float acc = 0.f;
          if (local_only) acc = lgrad[off + i];
          else
            for (int p = 0; p < a.world; ++p) {
              float g;
              asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(g) : "l"(a.grad[p] + off + i) : "memory");
              acc += g;
            }
          acc *= a.grad_mul;
          bad |= !isfinite(acc);
          sq += acc * acc;
          lgrad[off + i] = acc;
        }
      }
    }
    sq = block_sum_f(sq, sm);
    if (threadIdx.x == 0) atomicAdd(a.stats, sq);
    if (bad) a.stats[1] = 1.f;
  }
  grid_sync(a.grid_bar, gen);
  stamp(9);

  // ---- sync 1: exchange {sumsq, inf}
  if (blockIdx.x == 0 && threadIdx.x < a.world) {
    float* dst = a.pad[threadIdx.x] + a.rank * 2;
    const float s0 = a.stats[0], s1 = a.stats[1];
    asm volatile("st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(s0), "f"(s1) : "memory");
  }
  peer_barrier(a, 1);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    float sq = 0.f, inf = 0.f;
    const float* mine = a.pad[a.rank];
    for (int p = 0; p < a.world; ++p) {
      float x, y;
      asm volatile("ld.relaxed.sys.global.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y) : "l"(mine + p * 2) : "memory");
      sq += x; inf += y;
    }
    a.stats[2] = sq;
    a.stats[3] = inf;
  }
  grid_sync(a.grid_bar, gen);
  stamp(10);
  const float gsumsq = a.stats[2];
  const bool skip = a.stats[3] != 0.f || !isfinite(gsumsq);

  // ---- phase B: moments + update direction on the shard, per-tensor partial norms
  if (!skip) {
    const float gnorm = sqrtf(gsumsq);
    const float clip = (a.max_grad_norm > 0.f && gnorm > a.max_grad_norm) ? gnorm / a.max_grad_norm : 1.f;
    const float inv_clip = 1.f / clip;
    const float rbc1 = 1.f / a.bc1, rbc2 = 1.f / a.bc2;
    for (int c = blockIdx.x; c < a.nchunks; c += gridDim.x) {
      const int t = a.chunk_tensor[c];
      const long long off = a.chunk_start[c];
      const int n = a.chunk_len[c];
      const float wd = a.decay_flag[t] ? a.weight_decay : 0.f;
      float sp = 0.f, su = 0.f;
      auto one = [&](float g, float pp, float& mm, float& vv) -> float {
        g *= inv_clip;
        if (!a.adam_w_mode) g += wd * pp;
        mm = a.beta1 * mm + a.beta3 * g;
        vv = a.beta2 * vv + (1.f - a.beta2) * g * g;
        float u = (mm * rbc1) / (sqrtf(vv * rbc2) + a.eps);
        if (a.adam_w_mode) u += wd * pp;
        sp += pp * pp; su += u * u;
        return u;
      };
      const int nv = (off & 3) ? 0 : (n >> 2);           // 16 B path (chunks start 4-aligned except odd tails)
      for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        const long long o = off + 4 * i;
        float4 g = *reinterpret_cast<const float4*>(lgrad + o);
        const float4 pp = *reinterpret_cast<const float4*>(lparam + o);
        float4 mm = *reinterpret_cast<const float4*>(a.m + o);
        float4 vv = *reinterpret_cast<const float4*>(a.v + o);
        g.x = one(g.x, pp.x, mm.x, vv.x); g.y = one(g.y, pp.y, mm.y, vv.y);
        g.z = one(g.z, pp.z, mm.z, vv.z); g.w = one(g.w, pp.w, mm.w, vv.w);
        *reinterpret_cast<float4*>(a.m + o) = mm;
        *reinterpret_cast<float4*>(a.v + o) = vv;
        *reinterpret_cast<float4*>(lgrad + o) = g;
      }
      for (int i = 4 * nv + threadIdx.x; i < n; i += blockDim.x) {
        float mm = a.m[off + i], vv = a.v[off + i];
        const float u = one(lgrad[off + i], lparam[off + i], mm, vv);
        a.m[off + i] = mm; a.v[off + i] = vv;
        lgrad[off + i] = u;
      }
      sp = block_sum_f(sp, sm);
      su = block_sum_f(su, sm);
      if (threadIdx.x == 0) {
        atomicAdd(a.norms + 2 * t, sp);
        atomicAdd(a.norms + 2 * t + 1, su);
      }
    }
  }
  grid_sync(a.grid_bar, gen);
  stamp(11);

  // ---- sync 2: exchange the per-tensor partial norms
  if (!skip && blockIdx.x == 0) {
    const int per = 2 * a.ntensors;
    for (int idx = threadIdx.x; idx < per * a.world; idx += blockDim.x) {
      const int p = idx / per, j = idx % per;
      float* dst = a.pad[p] + 2 * a.world + a.rank * per + j;
      asm volatile("st.relaxed.sys.global.f32 [%0], %1;" ::"l"(dst), "f"(a.norms[j]) : "memory");
    }
  }
  peer_barrier(a, 2);
  grid_sync(a.grid_bar, gen);
  stamp(12);

  // ---- phase C: apply on the shard, push fp32 + bf16 to every rank, zero the local gradients
  if (!skip) {
    const int per = 2 * a.ntensors;
    const float* mine = a.pad[a.rank] + 2 * a.world;
    for (int c = blockIdx.x; c < a.nchunks; c += gridDim.x) {
      const int t = a.chunk_tensor[c];
      const long long off = a.chunk_start[c];
      const int n = a.chunk_len[c];
      if (threadIdx.x == 0) {
        float ratio = a.lr;
        if (a.use_nvlamb || (a.decay_flag[t] && a.weight_decay != 0.f)) {
          float pn = 0.f, un = 0.f;
          for (int p = 0; p < a.world; ++p) {
            float x, y;
            asm volatile("ld.relaxed.sys.global.v2.f32 {%0, %1}, [%2];" : "=f"(x), "=f"(y)
                         : "l"(mine + p * per + 2 * t) : "memory");
            pn += x; un += y;
          }
          pn = sqrtf(pn); un = sqrtf(un);
          if (pn > 0.f && un > 0.f) ratio = a.lr * pn / un;
        }
        s_bcast[0] = ratio;
      }
      __syncthreads();
      const float ratio = s_bcast[0];
      // 1-D tensors (biases, LayerNorm) always travel in fp32: the engine reads LayerNorm parameters from the master copy
      const bool push_f32 = a.push_master || !a.decay_flag[t];
      const int nv = (off & 3) ? 0 : (n >> 2);
      for (int i = threadIdx.x; i < nv; i += blockDim.x) {
        const long long o = off + 4 * i;
        const float4 pp = *reinterpret_cast<const float4*>(lparam + o);
        const float4 u = *reinterpret_cast<const float4*>(lgrad + o);
        float4 np;
        np.x = pp.x - ratio * u.x; np.y = pp.y - ratio * u.y; np.z = pp.z - ratio * u.z; np.w = pp.w - ratio * u.w;
        uint2 nb;
        nb.x = pack_bf16(np.x, np.y); nb.y = pack_bf16(np.z, np.w);
        if (push_f32 && a.use_multicast) {
          asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.param_mc + o), "f"(np.x),
                       "f"(np.y), "f"(np.z), "f"(np.w) : "memory");
        } else if (push_f32) {
#pragma unroll 1
          for (int p = 0; p < a.world; ++p) *reinterpret_cast<float4*>(a.param[p] + o) = np;
        } else {
          *reinterpret_cast<float4*>(lparam + o) = np;
        }
        if (a.shadow_mc) {
          asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(a.shadow_mc + o),
                       "f"(__uint_as_float(nb.x)), "f"(__uint_as_float(nb.y)) : "memory");
        } else {
#pragma unroll 1
          for (int p = 0; p < a.world; ++p) *reinterpret_cast<uint2*>(a.shadow[p] + o) = nb;
        }
      }
      for (int i = 4 * nv + threadIdx.x; i < n; i += blockDim.x) {
        const float np = lparam[off + i] - ratio * lgrad[off + i];
        const __nv_bfloat16 nb = __float2bfloat16(np);
#pragma unroll 1
        for (int p = 0; p < a.world; ++p) {
          if (push_f32 || p == a.rank) a.param[p][off + i] = np;
          a.shadow[p][off + i] = nb;
        }
      }
      __syncthreads();
    }
  }
  // The update directions u live in the gradient arena: every block must be done with phase C before anyone
  // clears it (a missing barrier here silently dropped the update of late chunks).
  grid_sync(a.grid_bar, gen);
  stamp(13);
  // zero the whole local gradient arena (all peers finished reading it: they passed barrier 1)
  {
    float4* g4 = reinterpret_cast<float4*>(lgrad);
    const long long n4 = a.numel >> 2;
    for (long long i = gtid; i < n4; i += gstride) g4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  grid_sync(a.grid_bar, gen);
  peer_barrier(a, 3);
  stamp(14);
  if (gtid == 0) a.stats[15] = (float)(globaltimer_ns() - t_enter) * 1e-6f;
}

void fused_allreduce_lamb(const FusedLambLaunch& L, cudaStream_t st) {
  FusedLambArgs a;
  a.rank = L.rank; a.world = L.world; a.use_multicast = L.use_multicast;
  if (L.world > MAX_WORLD) { fprintf(stderr, "[b200] fused all-reduce supports <= %d ranks\n", MAX_WORLD); abort(); }
  for (int p = 0; p < L.world; ++p) {
    a.grad[p] = (float*)L.grad_ptrs[p]; a.param[p] = (float*)L.param_ptrs[p];
    a.shadow[p] = (__nv_bfloat16*)L.shadow_ptrs[p]; a.pad[p] = (float*)L.pad_ptrs[p];
    a.flags[p] = (unsigned int*)L.flag_ptrs[p];
  }
  a.grad_mc = (float*)L.grad_mc; a.param_mc = (float*)L.param_mc; a.shadow_mc = (__nv_bfloat16*)L.shadow_mc;
  a.m = L.m; a.v = L.v; a.numel = L.numel; a.lo = L.lo; a.hi = L.hi;
  a.chunk_tensor = L.chunk_tensor; a.chunk_start = L.chunk_start; a.chunk_len = L.chunk_len;
  a.nchunks = L.nchunks; a.ntensors = L.ntensors; a.decay_flag = L.decay_flag; a.prereduced = L.prereduced;
  a.stats = L.stats; a.norms = L.norms; a.grid_bar = L.grid_bar; a.epoch = L.epoch;
  a.grad_mul = L.grad_mul;
  a.lr = L.lr; a.beta1 = L.beta1; a.beta2 = L.beta2; a.beta3 = L.grad_averaging ? 1.f - L.beta1 : 1.f;
  a.eps = L.eps; a.weight_decay = L.weight_decay;
  a.bc1 = L.bias_correction ? 1.f - powf(L.beta1, (float)L.step) : 1.f;
  a.bc2 = L.bias_correction ? 1.f - powf(L.beta2, (float)L.step) : 1.f;
  a.max_grad_norm = L.max_grad_norm; a.adam_w_mode = L.adam_w_mode; a.use_nvlamb = L.use_nvlamb; a.push_master = L.push_master;
  B200_CUDA_CHECK(cudaMemsetAsync(L.grid_bar, 0, sizeof(unsigned int), st));
  int dev, sms;
  B200_CUDA_CHECK(cudaGetDevice(&dev));
  B200_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  void* args[] = {(void*)&a};
  // cooperative launch: all CTAs co-resident, required by the in-kernel grid barrier
  B200_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)fused_allreduce_lamb_kernel, dim3(sms), dim3(FUSED_THREADS), args, 0, st));
}

// ------------------------------------------------------------------------------------------------
// General fp32 all-reduce over a symmetric staging buffer (K-FAC factor statistics; any other
// small reduction that would otherwise be an ncclAllReduce).  Two-shot inside ONE kernel:
//   barrier -> rank r reduces slice r of every peer's buffer (P2P loads or one multimem.ld_reduce), scales,
//   and writes the result into every peer's buffer (P2P stores or multimem.st) -> barrier.
// In place: slice r is read only by rank r, and rank r overwrites it after reading.
// ------------------------------------------------------------------------------------------------
struct PeerAllreduceArgs {
  int rank, world, use_multicast;
  float* buf[MAX_WORLD];
  float* buf_mc;
  unsigned int* flags[MAX_WORLD];
  unsigned int* grid_bar;
  unsigned int epoch;
  long long n;          // multiple of 4
  float scale;
};

__global__ void __launch_bounds__(FUSED_THREADS, 1) peer_allreduce_kernel(const PeerAllreduceArgs a) {
  unsigned int gen = 0;
  FusedLambArgs b;                       // the barrier helpers only read rank / world / flags / epoch
  b.rank = a.rank; b.world = a.world; b.epoch = a.epoch;
  for (int p = 0; p < a.world; ++p) b.flags[p] = a.flags[p];
  peer_barrier(b, 0);
  grid_sync(a.grid_bar, gen);
  const long long n4 = a.n >> 2;
  const long long per = (n4 + a.world - 1) / a.world;
  const long long lo = per * a.rank, hi = min(n4, lo + per);
  const long long gtid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long gstride = (long long)gridDim.x * blockDim.x;
  for (long long i = lo + gtid; i < hi; i += gstride) {
    float4 acc;
    if (a.use_multicast) {
      asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(acc.x), "=f"(acc.y), "=f"(acc.z), "=f"(acc.w) : "l"(a.buf_mc + 4 * i) : "memory");
    } else {
      acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll 1
      for (int p = 0; p < a.world; ++p) {
        float4 g;
        asm volatile("ld.relaxed.sys.global.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(g.x), "=f"(g.y), "=f"(g.z), "=f"(g.w) : "l"(a.buf[p] + 4 * i) : "memory");
        acc.x += g.x; acc.y += g.y; acc.z += g.z; acc.w += g.w;
      }
    }
    acc.x *= a.scale; acc.y *= a.scale; acc.z *= a.scale; acc.w *= a.scale;
    if (a.use_multicast) {
      asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(a.buf_mc + 4 * i), "f"(acc.x),
                   "f"(acc.y), "f"(acc.z), "f"(acc.w) : "memory");
    } else {
#pragma unroll 1
      for (int p = 0; p < a.world; ++p) *reinterpret_cast<float4*>(a.buf[p] + 4 * i) = acc;
    }
  }
  grid_sync(a.grid_bar, gen);
  peer_barrier(b, 1);
}

void peer_allreduce(const PeerAllreduceLaunch& L, cudaStream_t st) {
  PeerAllreduceArgs a;
  a.rank = L.rank; a.world = L.world; a.use_multicast = L.use_multicast;
  if (L.world > MAX_WORLD) { fprintf(stderr, "[b200] peer all-reduce supports <= %d ranks\n", MAX_WORLD); abort(); }
  for (int p = 0; p < L.world; ++p) { a.buf[p] = (float*)L.buf_ptrs[p]; a.flags[p] = (unsigned int*)L.flag_ptrs[p]; }
  a.buf_mc = (float*)L.buf_mc; a.grid_bar = L.grid_bar; a.epoch = L.epoch; a.n = L.n; a.scale = L.scale;
  B200_CUDA_CHECK(cudaMemsetAsync(L.grid_bar, 0, sizeof(unsigned int), st));
  int dev, sms;
  B200_CUDA_CHECK(cudaGetDevice(&dev));
  B200_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const long long work = (L.n / 4 + L.world - 1) / L.world;
  int grid = (int)std::min<long long>(sms, std::max<long long>(1, (work + FUSED_THREADS - 1) / FUSED_THREADS));
  void* args[] = {(void*)&a};
  B200_CUDA_CHECK(cudaLaunchCooperativeKernel((void*)peer_allreduce_kernel, dim3(grid), dim3(FUSED_THREADS), args, 0, st));
}

}  // namespace b200
