// GPU-side synthetic multiplexing from an HBM-resident slice pool: joint min/max of the two MS2 maps of each
// drawn pair (MS1 min/max from the first sample only), min-max normalisation with numpy's arithmetic
// (int32 pool: int32 subtraction, float64 true-divide, cast to float32; float32 pool: float32 throughout),
// and the 0.5/0.5 mix.  The pair indices come from the host (python `random`, bit-exact by construction).
// Replaces (reference /root/reference/dquartic): utils/data_loader.py:70-88 and model/model_interface.py:1071-1075.
#include <limits.h>
#include "common.cuh"

namespace dq {

// stats[item] = {ms2_min, ms2_max, ms1_min, ms1_max} as int32 (int pool) -- must be pre-initialised by stats_init
__global__ void stats_init_kernel(int* stats, int items, int is_float) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= items) return;
  if (is_float) {
    float* f = reinterpret_cast<float*>(stats);
    f[i * 4 + 0] = INFINITY; f[i * 4 + 1] = -INFINITY; f[i * 4 + 2] = INFINITY; f[i * 4 + 3] = -INFINITY;
  } else {
    stats[i * 4 + 0] = INT_MAX; stats[i * 4 + 1] = INT_MIN; stats[i * 4 + 2] = INT_MAX; stats[i * 4 + 3] = INT_MIN;
  }
}

__device__ __forceinline__ void atomicMinF(float* a, float v) {  // valid for any finite/inf floats
  if (v >= 0.f) atomicMin(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMax(reinterpret_cast<unsigned int*>(a), __float_as_uint(v));
}
__device__ __forceinline__ void atomicMaxF(float* a, float v) {
  if (v >= 0.f) atomicMax(reinterpret_cast<int*>(a), __float_as_int(v));
  else atomicMin(reinterpret_cast<unsigned int*>(a), __float_as_uint(v));
}

template <typename T>
__global__ void __launch_bounds__(256) minmax_kernel(const T* __restrict__ ms2, const T* __restrict__ ms1,
                                                     const long long* __restrict__ pairs, void* stats, long n2, int n1) {
  __shared__ T red[2 * 8];
  const int item = blockIdx.y;
  const long long i1 = pairs[item * 2], i2 = pairs[item * 2 + 1];
  const T* a = ms2 + (size_t)i1 * n2;
  const T* b = ms2 + (size_t)i2 * n2;
  T mn = a[0], mx = a[0];
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    T va = a[i], vb = b[i];
    mn = min(mn, min(va, vb));
    mx = max(mx, max(va, vb));
  }
  if (blockIdx.x == 0) {  // MS1 stats from sample 1 only (data_loader.py:72-73)
    const T* c = ms1 + (size_t)i1 * n1;
    T m1n = c[0], m1x = c[0];
    for (int i = threadIdx.x; i < n1; i += blockDim.x) { m1n = min(m1n, c[i]); m1x = max(m1x, c[i]); }
    if constexpr (sizeof(T) == 4 && !__is_same(T, float)) {
      atomicMin(reinterpret_cast<int*>(stats) + item * 4 + 2, (int)m1n);
      atomicMax(reinterpret_cast<int*>(stats) + item * 4 + 3, (int)m1x);
    } else {
      atomicMinF(reinterpret_cast<float*>(stats) + item * 4 + 2, (float)m1n);
      atomicMaxF(reinterpret_cast<float*>(stats) + item * 4 + 3, (float)m1x);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    T on = __shfl_xor_sync(0xffffffffu, mn, o), ox = __shfl_xor_sync(0xffffffffu, mx, o);
    mn = min(mn, on); mx = max(mx, ox);
  }
  if ((threadIdx.x & 31) == 0) { red[(threadIdx.x >> 5) * 2] = mn; red[(threadIdx.x >> 5) * 2 + 1] = mx; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w) { mn = min(mn, red[w * 2]); mx = max(mx, red[w * 2 + 1]); }
    if constexpr (sizeof(T) == 4 && !__is_same(T, float)) {
      atomicMin(reinterpret_cast<int*>(stats) + item * 4 + 0, (int)mn);
      atomicMax(reinterpret_cast<int*>(stats) + item * 4 + 1, (int)mx);
    } else {
      atomicMinF(reinterpret_cast<float*>(stats) + item * 4 + 0, (float)mn);
      atomicMaxF(reinterpret_cast<float*>(stats) + item * 4 + 1, (float)mx);
    }
  }
}

__device__ __forceinline__ float norm_i32(int v, int mn, int mx) { return (float)((double)(v - mn) / (double)(mx - mn)); }
__device__ __forceinline__ float norm_f32(float v, float mn, float mx) { return __fdiv_rn(__fsub_rn(v, mn), __fsub_rn(mx, mn)); }

// x0 = norm(ms2[i1]); other = norm(ms2[i2]); cond = w0*x0 + w1*other; ms1 outputs for both samples
template <typename T>
__global__ void __launch_bounds__(256) gather_norm_mix_kernel(const T* __restrict__ ms2, const T* __restrict__ ms1,
                                                              const long long* __restrict__ pairs, const void* stats,
                                                              float w0, float w1, float* __restrict__ x0,
                                                              float* __restrict__ other, float* __restrict__ cond,
                                                              float* __restrict__ ms1_1, float* __restrict__ ms1_2,
                                                              long n2, int n1) {
  const int item = blockIdx.y;
  const long long i1 = pairs[item * 2], i2 = pairs[item * 2 + 1];
  const T* a = ms2 + (size_t)i1 * n2;
  const T* b = ms2 + (size_t)i2 * n2;
  const T* st = reinterpret_cast<const T*>(stats) + item * 4;
  const T mn = st[0], mx = st[1], m1n = st[2], m1x = st[3];
  const long stride = (long)gridDim.x * blockDim.x;
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n2; i += stride) {
    float va, vb;
    if constexpr (__is_same(T, float)) { va = norm_f32(a[i], mn, mx); vb = norm_f32(b[i], mn, mx); }
    else { va = norm_i32(a[i], mn, mx); vb = norm_i32(b[i], mn, mx); }
    size_t o = (size_t)item * n2 + i;
    if (x0) x0[o] = va;
    if (other) other[o] = vb;
    if (cond) cond[o] = __fadd_rn(__fmul_rn(va, w0), __fmul_rn(vb, w1));
  }
  if (blockIdx.x == 0) {
    for (int i = threadIdx.x; i < n1; i += blockDim.x) {
      float v1, v2;
      if constexpr (__is_same(T, float)) {
        v1 = norm_f32(ms1[(size_t)i1 * n1 + i], m1n, m1x); v2 = norm_f32(ms1[(size_t)i2 * n1 + i], m1n, m1x);
      } else {
        v1 = norm_i32(ms1[(size_t)i1 * n1 + i], m1n, m1x); v2 = norm_i32(ms1[(size_t)i2 * n1 + i], m1n, m1x);
      }
      if (ms1_1) ms1_1[(size_t)item * n1 + i] = v1;
      if (ms1_2) ms1_2[(size_t)item * n1 + i] = v2;
    }
  }
}

}  // namespace dq
using namespace dq;

// pool dtype: 0 = int32, 1 = float32.  pairs: int64 (items, 2) on the device.  stats: 4 x 4 bytes per item.
DQ_API int dq_multiplex(const void* ms2, const void* ms1, int dtype, const long long* pairs, void* stats, float w0,
                        float w1, float* x0, float* other, float* cond, float* ms1_1, float* ms1_2, int items,
                        long n2, int n1, void* stream) {
  if (items <= 0) return 0;
  cudaStream_t st = (cudaStream_t)stream;
  stats_init_kernel<<<(items + 127) / 128, 128, 0, st>>>((int*)stats, items, dtype);
  DQ_LAUNCH_CHECK();
  long bx = (n2 + 256L * 16 - 1) / (256L * 16);
  if (bx > 64) bx = 64;
  if (bx < 1) bx = 1;
  dim3 grid((unsigned)bx, (unsigned)items);
  if (dtype == 0) {
    minmax_kernel<int><<<grid, 256, 0, st>>>((const int*)ms2, (const int*)ms1, pairs, stats, n2, n1);
    DQ_LAUNCH_CHECK();
    gather_norm_mix_kernel<int><<<grid, 256, 0, st>>>((const int*)ms2, (const int*)ms1, pairs, stats, w0, w1, x0, other, cond, ms1_1, ms1_2, n2, n1);
  } else if (dtype == 1) {
    minmax_kernel<float><<<grid, 256, 0, st>>>((const float*)ms2, (const float*)ms1, pairs, stats, n2, n1);
    DQ_LAUNCH_CHECK();
    gather_norm_mix_kernel<float><<<grid, 256, 0, st>>>((const float*)ms2, (const float*)ms1, pairs, stats, w0, w1, x0, other, cond, ms1_1, ms1_2, n2, n1);
  } else return -3;
  DQ_LAUNCH_CHECK();
  return 0;
}
