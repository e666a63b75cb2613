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

// ---------------------------------------------------------------------------------------------- resampling glue
// The backward passes of Upsample (nearest x2 + Conv1d k3, unet1d.py:93-96) and Downsample (Conv1d k4 s2 p1, 110) run
// through the fused stride-1 backward kernel (conv_fused.cu) on re-indexed tensors:
//   up  : x_up[2j] = x_up[2j+1] = x[j];  dx[j] = dx_up[2j] + dx_up[2j+1]
//   down: the k4/s2 conv is a k3/s1 conv over the 2C-channel half-rate tensor [x_even; x_odd] with weights
//         even channel c: taps (0, w1, w3), odd channel c: taps (w0, w2, 0)
namespace dq {
__global__ void __launch_bounds__(256) upsample2x_kernel(const float* __restrict__ x, float* __restrict__ y, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float v = x[i];
    reinterpret_cast<float2*>(y)[i] = make_float2(v, v);
  }
}
__global__ void __launch_bounds__(256) fold2x_kernel(const float* __restrict__ d, float* __restrict__ dx, long n, int acc) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const float2 v = reinterpret_cast<const float2*>(d)[i];
    const float s = v.x + v.y;
    dx[i] = acc ? dx[i] + s : s;
  }
}
// x (rows, L) -> y viewed as (R, 2C, L/2): row = r*C + c goes to rows r*2C + c (even samples) and r*2C + C + c (odd)
__global__ void __launch_bounds__(256) s2d_kernel(const float* __restrict__ x, float* __restrict__ y, int C, int Lh, long n) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long row = i / Lh;
    const int j = (int)(i - row * Lh);
    const long r = row / C;
    const int c = (int)(row - r * C);
    const float2 v = reinterpret_cast<const float2*>(x)[i];
    y[((r * 2 * C + c)) * Lh + j] = v.x;
    y[((r * 2 * C + C + c)) * Lh + j] = v.y;
  }
}
__global__ void __launch_bounds__(256) d2s_kernel(const float* __restrict__ d, float* __restrict__ dx, int C, int Lh, long n, int acc) {
  for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
    const long row = i / Lh;
    const int j = (int)(i - row * Lh);
    const long r = row / C;
    const int c = (int)(row - r * C);
    float2 v = make_float2(d[((r * 2 * C + c)) * Lh + j], d[((r * 2 * C + C + c)) * Lh + j]);
    float2* dst = reinterpret_cast<float2*>(dx) + i;
    if (acc) { const float2 o = *dst; v.x += o.x; v.y += o.y; }
    *dst = v;
  }
}
// w4 (co, ci, 4) <-> w3 (co, 2ci, 3).  dir = 0: w3 = pack(w4);  dir = 1: w4 += unpack(w3)  (gradient)
__global__ void down_w_kernel(float* __restrict__ w4, float* __restrict__ w3, int co, int ci, int dir) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= co * ci) return;
  const int o = i / ci, c = i - o * ci;
  float* a = w4 + (size_t)i * 4;
  float* e = w3 + ((size_t)o * 2 * ci + c) * 3;        // even-sample channel c
  float* d = w3 + ((size_t)o * 2 * ci + ci + c) * 3;   // odd-sample channel c
  if (dir == 0) {
    e[0] = 0.f; e[1] = a[1]; e[2] = a[3];
    d[0] = a[0]; d[1] = a[2]; d[2] = 0.f;
  } else {
    a[0] += d[0]; a[1] += e[1]; a[2] += d[1]; a[3] += e[2];
  }
}
static inline unsigned ew_grid(long n) { long g = (n + 255) / 256; return (unsigned)(g > 148L * 16 ? 148L * 16 : g); }
}  // namespace dq

DQ_API int dq_upsample2x(const float* x, float* y, long n, void* stream) {
  if (n <= 0) return 0;
  dq::upsample2x_kernel<<<dq::ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, y, n);
  DQ_LAUNCH_CHECK();
  return 0;
}
DQ_API int dq_fold2x(const float* d, float* dx, long n, int acc, void* stream) {
  if (n <= 0) return 0;
  dq::fold2x_kernel<<<dq::ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d, dx, n, acc);
  DQ_LAUNCH_CHECK();
  return 0;
}
DQ_API int dq_s2d(const float* x, float* y, int R, int C, int L, void* stream) {
  if (L & 1) return -2;
  const long n = (long)R * C * (L / 2);
  if (n <= 0) return 0;
  dq::s2d_kernel<<<dq::ew_grid(n), 256, 0, (cudaStream_t)stream>>>(x, y, C, L / 2, n);
  DQ_LAUNCH_CHECK();
  return 0;
}
DQ_API int dq_d2s(const float* d, float* dx, int R, int C, int L, int acc, void* stream) {
  if (L & 1) return -2;
  const long n = (long)R * C * (L / 2);
  if (n <= 0) return 0;
  dq::d2s_kernel<<<dq::ew_grid(n), 256, 0, (cudaStream_t)stream>>>(d, dx, C, L / 2, n, acc);
  DQ_LAUNCH_CHECK();
  return 0;
}
DQ_API int dq_down_w(float* w4, float* w3, int co, int ci, int dir, void* stream) {
  if (co * ci <= 0) return 0;
  dq::down_w_kernel<<<(unsigned)((co * ci + 127) / 128), 128, 0, (cudaStream_t)stream>>>(w4, w3, co, ci, dir);
  DQ_LAUNCH_CHECK();
  return 0;
}
