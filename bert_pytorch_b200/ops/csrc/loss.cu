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

// ------------------------------------------------------------------------------------------------
// NSP head, everything after the pooler GEMM in ONE launch: the [B,H] x [H,2] classifier,
// its cross-entropy (ignore_index -1, mean over the valid rows), and the backward: dW[2,H], db[2] accumulated with
// atomics and the gradient of the pooler pre-activation  dz = (dlogits . W) * (1 - pooled^2)  (tanh' folded in).
// One block per sequence; pooled is the bf16 output of the pooler GEMM's bias+tanh epilogue.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
nsp_head_kernel(const __nv_bfloat16* __restrict__ pooled, const __nv_bfloat16* __restrict__ w, const float* __restrict__ bias,
                const long long* __restrict__ labels, int B, int H, float grad_scale, float* __restrict__ loss_out,
                __nv_bfloat16* __restrict__ dz, float* __restrict__ dw, float* __restrict__ db) {
  __shared__ float sm[32];
  __shared__ float s_bc[3];
  const int b = blockIdx.x;
  int n_valid = 0;
  for (int i = 0; i < B; ++i) n_valid += labels[i] >= 0 ? 1 : 0;      // B is ~100: cheaper than a second launch
  const long long y = labels[b];
  const __nv_bfloat16* x = pooled + (size_t)b * H;
  float a0 = 0.f, a1 = 0.f;
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    const float xv = __bfloat162float(x[i]);
    a0 += xv * __bfloat162float(w[i]);
    a1 += xv * __bfloat162float(w[H + i]);
  }
  auto block_sum = [&](float v) -> float {
    v = warp_sum(v);
    if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
    __syncthreads();
    float r = 0.f;
    for (int k = 0; k < (int)(blockDim.x >> 5); ++k) r += sm[k];
    __syncthreads();
    return r;
  };
  const float l0 = block_sum(a0) + bias[0], l1 = block_sum(a1) + bias[1];
  float d0 = 0.f, d1 = 0.f;
  if (y >= 0 && n_valid > 0) {
    const float m = fmaxf(l0, l1);
    const float lse = m + logf(expf(l0 - m) + expf(l1 - m));
    const float p0 = expf(l0 - lse), p1 = expf(l1 - lse);
    const float inv = 1.f / (float)n_valid;
    if (threadIdx.x == 0) atomicAdd(loss_out, (lse - (y == 0 ? l0 : l1)) * inv);
    d0 = (p0 - (y == 0 ? 1.f : 0.f)) * inv * grad_scale;
    d1 = (p1 - (y == 1 ? 1.f : 0.f)) * inv * grad_scale;
  }
  if (threadIdx.x == 0 && (d0 != 0.f || d1 != 0.f)) {
    atomicAdd(db, d0);
    atomicAdd(db + 1, d1);
  }
  for (int i = threadIdx.x; i < H; i += blockDim.x) {
    const float xv = __bfloat162float(x[i]);
    const float g = d0 * __bfloat162float(w[i]) + d1 * __bfloat162float(w[H + i]);
    dz[(size_t)b * H + i] = __float2bfloat16(g * (1.f - xv * xv));
    if (d0 != 0.f || d1 != 0.f) {
      atomicAdd(dw + i, d0 * xv);
      atomicAdd(dw + H + i, d1 * xv);
    }
  }
}

void nsp_head(const void* pooled, const void* w, const float* bias, const long long* labels, int B, int H, float grad_scale,
              float* loss_out, void* dz, float* dw, float* db, cudaStream_t st) {
  if (B > 0)
    nsp_head_kernel<<<B, 128, 0, st>>>((const __nv_bfloat16*)pooled, (const __nv_bfloat16*)w, bias, labels, B, H, grad_scale,
                                       loss_out, (__nv_bfloat16*)dz, dw, db);
}

}  // namespace b200
