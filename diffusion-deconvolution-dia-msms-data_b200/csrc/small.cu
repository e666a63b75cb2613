// Tiny per-sample layers of the denoiser: sinusoidal time embedding, the time MLP, and ONE batched Linear that
// produces every ResnetBlock / ConditionalScaleShift (scale, shift) vector of the network in a single launch
// (the per-block `mlp.1` weights are laid out contiguously in the flat parameter buffer).
// Replaces (reference /root/reference/dquartic/model/unet1d.py): SinusoidalPosEmb 211-218, time_mlp 958-960,
// ResnetBlock.mlp 292-296/316, ConditionalScaleShift.to_scale_shift 664/677, Attention.to_k 535/555.
#include "common.cuh"

namespace dq {

__global__ void time_embed_kernel(const long long* __restrict__ t, float* __restrict__ out, int b, int dim, float neg_e) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  int half = dim / 2;
  if (i >= b * half) return;
  int s = i / half, k = i % half;
  float f = expf((float)k * neg_e);  // neg_e = -ln(theta)/(half-1), rounded to fp32 on the host like torch does
  float a = (float)t[s] * f;
  out[(size_t)s * dim + k] = sinf(a);
  out[(size_t)s * dim + half + k] = cosf(a);
}

// y[r][o] = bias[o] + sum_i x[r][i] * W[o][i]
__global__ void __launch_bounds__(256) linear_fwd_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                         const float* __restrict__ bias, float* __restrict__ y,
                                                         int rows, int in, int out) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)rows * out) return;
  int r = (int)(idx / out), o = (int)(idx % out);
  const float* xr = x + (size_t)r * in;
  const float* wr = W + (size_t)o * in;
  float acc = bias ? bias[o] : 0.f;
  for (int i = 0; i < in; ++i) acc = fmaf(xr[i], wr[i], acc);
  y[idx] = acc;
}

// dW[o][i] += sum_r dy[r][o] x[r][i] ; db[o] += sum_r dy[r][o]     (one thread per (o, i); i == in -> bias)
__global__ void __launch_bounds__(256) linear_bwd_w_kernel(const float* __restrict__ x, const float* __restrict__ dy,
                                                           float* __restrict__ dW, float* __restrict__ db, int rows,
                                                           int in, int out) {
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)out * (in + 1)) return;
  int o = (int)(idx / (in + 1)), i = (int)(idx % (in + 1));
  float acc = 0.f;
  if (i < in) {
    for (int r = 0; r < rows; ++r) acc = fmaf(dy[(size_t)r * out + o], x[(size_t)r * in + i], acc);
    dW[(size_t)o * in + i] += acc;
  } else if (db) {
    for (int r = 0; r < rows; ++r) acc += dy[(size_t)r * out + o];
    db[o] += acc;
  }
}

// dx[r][i] = sum_o W[o][i] dy[r][o]   (block per row, in <= 32)
__global__ void __launch_bounds__(256) linear_bwd_x_kernel(const float* __restrict__ W, const float* __restrict__ dy,
                                                           float* __restrict__ dx, int in, int out) {
  __shared__ float red[8 * 32];
  const int r = blockIdx.x;
  float acc[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) acc[i] = 0.f;
  for (int o = threadIdx.x; o < out; o += blockDim.x) {
    float d = dy[(size_t)r * out + o];
    const float* wr = W + (size_t)o * in;
#pragma unroll
    for (int i = 0; i < 32; ++i)
      if (i < in) acc[i] = fmaf(d, wr[i], acc[i]);
  }
  float tot = block_reduce_vec<32>(acc, red);
  if ((int)threadIdx.x < in) dx[(size_t)r * in + threadIdx.x] = tot;
}

__global__ void __launch_bounds__(256) act_fwd_kernel(const float* __restrict__ x, float* __restrict__ y, int act, long n) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = act_fwd(x[i], act);
}
__global__ void __launch_bounds__(256) act_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ xpre,
                                                      float* __restrict__ dx, int act, long n) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dx[i] = dy[i] * act_bwd(xpre[i], act);
}

// (B, C, L) <-> (B, L, C)
__global__ void __launch_bounds__(256) ncl_to_nlc_kernel(const float* __restrict__ in, float* __restrict__ out, int B,
                                                         int C, int L, int reverse) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long)B * C * L) return;
  int l = (int)(i % L), c = (int)((i / L) % C), b = (int)(i / ((long)L * C));
  size_t ncl = ((size_t)b * C + c) * L + l, nlc = ((size_t)b * L + l) * C + c;
  if (reverse) out[ncl] = in[nlc];
  else out[nlc] = in[ncl];
}

}  // namespace dq
using namespace dq;

DQ_API int dq_time_embed(const long long* t, float* out, int b, int dim, float neg_e, void* stream) {
  int n = b * (dim / 2);
  if (n <= 0) return 0;
  time_embed_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(t, out, b, dim, neg_e);
  DQ_LAUNCH_CHECK();
  return 0;
}
DQ_API int dq_linear_fwd(const float* x, const float* W, const float* bias, float* y, int rows, int in, int out,
                         void* stream) {
  long n = (long)rows * out;
  if (n <= 0) return 0;
  linear_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, W, bias, y, rows, in, out);
  DQ_LAUNCH_CHECK();
  return 0;
}
DQ_API int dq_linear_bwd(const float* x, const float* W, const float* dy, float* dx, float* dW, float* db, int rows,
                         int in, int out, void* stream) {
  if (rows <= 0 || in <= 0 || out <= 0) return 0;
  if (in > 32) return -3;
  cudaStream_t st = (cudaStream_t)stream;
  if (dW) {
    long n = (long)out * (in + 1);
    linear_bwd_w_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(x, dy, dW, db, rows, in, out);
    DQ_LAUNCH_CHECK();
  }
  if (dx) {
    linear_bwd_x_kernel<<<(unsigned)rows, 256, 0, st>>>(W, dy, dx, in, out);
    DQ_LAUNCH_CHECK();
  }
  return 0;
}
DQ_API int dq_act_fwd(const float* x, float* y, int act, long n, void* stream) {
  if (n <= 0) return 0;
  act_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(x, y, act, n);
  DQ_LAUNCH_CHECK();
  return 0;
}
DQ_API int dq_act_bwd(const float* dy, const float* xpre, float* dx, int act, long n, void* stream) {
  if (n <= 0) return 0;
  act_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(dy, xpre, dx, act, n);
  DQ_LAUNCH_CHECK();
  return 0;
}
DQ_API int dq_ncl_nlc(const float* in, float* out, int B, int C, int L, int reverse, void* stream) {
  long n = (long)B * C * L;
  if (n <= 0) return 0;
  ncl_to_nlc_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(in, out, B, C, L, reverse);
  DQ_LAUNCH_CHECK();
  return 0;
}
