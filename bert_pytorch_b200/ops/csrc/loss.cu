// Softmax cross-entropy over the (masked-position) MLM logits, forward + backward in one kernel
// (SURVEY.md K23/K24).  logits are bf16 [R, V]; rows whose target is < 0 are ignored (gradient zero).
// The mean is taken over the device-side count of valid targets, so no host sync is needed to know how
// many positions were masked.  The gradient overwrites the logits in place:
//     dlogits = (softmax - onehot) * grad_scale / count          loss += sum_r (lse_r - logit_r[t_r]) / count
#include "common.cuh"
#include "kernels.h"

namespace b200 {

constexpr int CE_THREADS = 256;

__global__ void __launch_bounds__(CE_THREADS)
softmax_ce_kernel(__nv_bfloat16* __restrict__ logits, int ld, const int* __restrict__ targets,
                  const int* __restrict__ count, float grad_scale, float* __restrict__ loss_out, int V) {
  __shared__ float sm[32];
  __shared__ float s_max, s_sum;
  const int r = blockIdx.x;
  __nv_bfloat16* row = logits + (size_t)r * ld;
  const int tgt = targets[r];
  const int nvec = V >> 3;  // V % 8 == 0
  uint4* row4 = reinterpret_cast<uint4*>(row);
  if (tgt < 0) {  // ignored position: zero gradient row
    for (int i = threadIdx.x; i < nvec; i += CE_THREADS) row4[i] = make_uint4(0, 0, 0, 0);
    return;
  }
  // pass 1: online max / sum of exp
  float m = -INFINITY, s = 0.f;
  for (int i = threadIdx.x; i < nvec; i += CE_THREADS) {
    const uint4 u = row4[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float f[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 p = unpack_bf16(w[t]);
      f[2 * t] = p.x; f[2 * t + 1] = p.y;
    }
    float lm = f[0];
#pragma unroll
    for (int t = 1; t < 8; ++t) lm = fmaxf(lm, f[t]);
    const float nm = fmaxf(m, lm);
    float acc = 0.f;
#pragma unroll
    for (int t = 0; t < 8; ++t) acc += __expf(f[t] - nm);
    s = s * __expf(m - nm) + acc;
    m = nm;
  }
  // block combine (max first, then rescaled sums)
  float wm = warp_max(m);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) sm[warp] = wm;
  __syncthreads();
  if (warp == 0) {
    float v = lane < CE_THREADS / 32 ? sm[lane] : -INFINITY;
    v = warp_max(v);
    if (lane == 0) s_max = v;
  }
  __syncthreads();
  const float gmax = s_max;
  float part = (m == -INFINITY) ? 0.f : s * __expf(m - gmax);
  part = warp_sum(part);
  if (lane == 0) sm[warp] = part;
  __syncthreads();
  if (warp == 0) {
    float v = lane < CE_THREADS / 32 ? sm[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) s_sum = v;
  }
  __syncthreads();
  const float inv_sum = 1.f / s_sum;
  const float cnt = (float)max(*count, 1);
  const float gs = grad_scale / cnt;
  if (threadIdx.x == 0) {
    const float lt = __bfloat162float(row[tgt]);
    atomicAdd(loss_out, (logf(s_sum) + gmax - lt) / cnt);
  }
  __syncthreads();  // the target logit was read before anyone overwrites it
  // pass 2: gradient in place
  for (int i = threadIdx.x; i < nvec; i += CE_THREADS) {
    const uint4 u = row4[i];
    const uint32_t w[4] = {u.x, u.y, u.z, u.w};
    float f[8];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
      const float2 p = unpack_bf16(w[t]);
      f[2 * t] = p.x; f[2 * t + 1] = p.y;
    }
#pragma unroll
    for (int t = 0; t < 8; ++t) {
      float pr = __expf(f[t] - gmax) * inv_sum;
      if (i * 8 + t == tgt) pr -= 1.f;
      f[t] = pr * gs;
    }
    row4[i] = make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
  }
}

void softmax_ce(void* logits, int ld, const int* targets, const int* count, float grad_scale, float* loss_out, int R,
                int V, cudaStream_t st) {
  if (R > 0) softmax_ce_kernel<<<R, CE_THREADS, 0, st>>>((__nv_bfloat16*)logits, ld, targets, count, grad_scale, loss_out, V);
}

}  // namespace b200
