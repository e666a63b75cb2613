// Shared device helpers for the dquartic B200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define DQ_API extern "C" __attribute__((visibility("default")))

#define DQ_LAUNCH_CHECK()                                \
  do {                                                   \
    cudaError_t e__ = cudaGetLastError();                \
    if (e__ != cudaSuccess) return (int)e__;             \
  } while (0)

namespace dq {

constexpr int kHeads = 4;
constexpr int kDimHead = 32;
constexpr int kHD = kHeads * kDimHead;  // 128

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// activations: 0 none, 1 SiLU, 2 GELU(erf), 3 Softplus (beta 1, threshold 20: torch.nn.Softplus defaults)
__device__ __forceinline__ float act_fwd(float z, int act) {
  if (act == 1) return z / (1.f + __expf(-z));
  if (act == 2) return 0.5f * z * (1.f + erff(z * 0.70710678118654752440f));
  if (act == 3) return z > 20.f ? z : log1pf(expf(z));
  return z;
}
__device__ __forceinline__ float act_bwd(float z, int act) {  // d act / d z
  if (act == 1) {
    float s = 1.f / (1.f + __expf(-z));
    return s * (1.f + z * (1.f - s));
  }
  if (act == 2) {
    float cdf = 0.5f * (1.f + erff(z * 0.70710678118654752440f));
    float pdf = 0.39894228040143267794f * __expf(-0.5f * z * z);
    return cdf + z * pdf;
  }
  if (act == 3) return z > 20.f ? 1.f : 1.f / (1.f + expf(-z));
  return 1.f;
}

// Block-wide sum of `nv` per-thread values (nv <= NV), result valid in threads [0, nv) of the block:
// thread i returns the total of value i.  `red` must hold (blockDim.x/32) * NV floats.
template <int NV>
__device__ __forceinline__ float block_reduce_vec(float (&v)[NV], float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    float s = warp_sum(v[i]);
    if (lane == 0) red[warp * NV + i] = s;
  }
  __syncthreads();
  float out = 0.f;
  if ((int)threadIdx.x < NV) {
    for (int w = 0; w < nwarp; ++w) out += red[w * NV + threadIdx.x];
  }
  __syncthreads();
  return out;
}

}  // namespace dq
