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
  // blockIdx.y splits the rows (few outputs x many rows - to_k of the mid attention: 1152 threads walking 2176 rows
  // serially took 470 us); with more than one split the partial sums are added atomically
  long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= (long)out * (in + 1)) return;
  int o = (int)(idx / (in + 1)), i = (int)(idx % (in + 1));
  const int per = (rows + gridDim.y - 1) / gridDim.y;
  const int r_lo = blockIdx.y * per, r_hi = min(rows, r_lo + per);
  float acc = 0.f;
  if (i < in) {
    for (int r = r_lo; r < r_hi; ++r) acc = fmaf(dy[(size_t)r * out + o], x[(size_t)r * in + i], acc);
    if (gridDim.y == 1) dW[(size_t)o * in + i] += acc; else atomicAdd(dW + (size_t)o * in + i, acc);
  } else if (db) {
    for (int r = r_lo; r < r_hi; ++r) acc += dy[(size_t)r * out + o];
    if (gridDim.y == 1) db[o] += acc; else atomicAdd(db + o, acc);
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
    const unsigned gx = (unsigned)((n + 255) / 256);
    unsigned gy = 1;   // fill the GPU when there are few (output, input) pairs and many rows
    if (gx < 148 && rows >= 256) { gy = (296 + gx - 1) / gx; if ((int)gy > rows / 32) gy = (unsigned)(rows / 32); if (gy < 1) gy = 1; }
    linear_bwd_w_kernel<<<dim3(gx, gy), 256, 0, st>>>(x, dy, dW, db, rows, in, out);
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

// ------------------------------------------------------------------------------------------------------------------
// Backward of init_conv = Conv1d(2 -> COUT, k7, pad 3) over cat(ConditionalScaleShift(cond), x) (unet1d.py:1107-1117,
// 677-678) in ONE pass over (d, cond, x), without the data gradient ever existing.  Per sample s the pass produces the
// RAW correlations  G_s[co][ci][k] = sum_{rows of s, p} d[co][p] in_ci[p + k - 3]  (cond NOT scaled / shifted),
// D_s[co] = sum d[co][p] and the edge sums E_s[co][k] = sum of d[co][p] over the positions whose tap k falls outside the
// row.  Everything else is algebra on those 18 COUT numbers (initconv_bwd_finalize_kernel):
//   dW[co][0][k] += (1 + scale_s) G_s[co][0][k] + shift_s (D_s[co] - E_s[co][k]);   dW[co][1][k] += G_s[co][1][k]
//   db[co] += D_s[co];   d scale_s = sum_{co,k} W[co][0][k] G_s[co][0][k];   d shift_s = sum W[co][0][k] (D_s - E_s)
// Scratch record per sample and group of 4 output channels: G (4*2*7) | D (4) | E (4*7) = 88 floats.
constexpr int IC_REC = 88;
__global__ void __launch_bounds__(128) initconv_bwd_kernel(const float* __restrict__ d, const float* __restrict__ cond,
                                                           const float* __restrict__ x, float* __restrict__ scratch,
                                                           int cout, int L, int chunk, int rows_per_sample) {
  __shared__ float red[4 * 60];
  const int r = blockIdx.y, cg = blockIdx.z;
  const int n_begin = blockIdx.x * chunk, n_end = min(L, n_begin + chunk);
  const float* dr = d + ((size_t)r * cout + 4 * cg) * L;
  const float* in[2] = {cond + (size_t)r * L, x + (size_t)r * L};
  float acc[60];
#pragma unroll
  for (int i = 0; i < 60; ++i) acc[i] = 0.f;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int p = n_begin + 4 * threadIdx.x; p < n_end; p += 4 * 128) {
    float dv[4][4];
#pragma unroll
    for (int co = 0; co < 4; ++co) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(dr + (size_t)co * L + p));
      dv[co][0] = v.x; dv[co][1] = v.y; dv[co][2] = v.z; dv[co][3] = v.w;
      acc[56 + co] += (v.x + v.y) + (v.z + v.w);
    }
#pragma unroll
    for (int ci = 0; ci < 2; ++ci) {
      const float4 a = p >= 4 ? __ldg(reinterpret_cast<const float4*>(in[ci] + p - 4)) : z4;
      const float4 b = __ldg(reinterpret_cast<const float4*>(in[ci] + p));
      const float4 c = p + 4 < L ? __ldg(reinterpret_cast<const float4*>(in[ci] + p + 4)) : z4;
      const float w[12] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w, c.x, c.y, c.z, c.w};   // positions p - 4 .. p + 7
#pragma unroll
      for (int co = 0; co < 4; ++co)
#pragma unroll
        for (int k = 0; k < 7; ++k)
#pragma unroll
          for (int i = 0; i < 4; ++i) acc[(co * 2 + ci) * 7 + k] = fmaf(dv[co][i], w[i + k + 1], acc[(co * 2 + ci) * 7 + k]);
    }
    // taps that leave the row: k < 3 at positions p' < 3 - k, k > 3 at positions p' > L - 1 - (k - 3)
    if (p == 0 || p + 4 >= L) {
      float* e = scratch + ((size_t)(r / rows_per_sample) * gridDim.z + cg) * IC_REC + 60;
#pragma unroll
      for (int co = 0; co < 4; ++co) {
        if (p == 0) {
          atomicAdd(e + co * 7 + 0, dv[co][0] + dv[co][1] + dv[co][2]);
          atomicAdd(e + co * 7 + 1, dv[co][0] + dv[co][1]);
          atomicAdd(e + co * 7 + 2, dv[co][0]);
        }
        if (p + 4 >= L) {
          atomicAdd(e + co * 7 + 4, dv[co][3]);
          atomicAdd(e + co * 7 + 5, dv[co][3] + dv[co][2]);
          atomicAdd(e + co * 7 + 6, dv[co][3] + dv[co][2] + dv[co][1]);
        }
      }
    }
  }
  const float tot = block_reduce_vec<60>(acc, red);
  if (threadIdx.x < 60) atomicAdd(scratch + ((size_t)(r / rows_per_sample) * gridDim.z + cg) * IC_REC + threadIdx.x, tot);
}

__global__ void __launch_bounds__(64) initconv_bwd_finalize_kernel(const float* __restrict__ scratch, const float* __restrict__ w,
                                                                   const float* __restrict__ ss, int ss_stride, float* dw,
                                                                   float* db, float* dss, int cout) {
  __shared__ float red[2 * 2];
  const int s = blockIdx.x, cg = blockIdx.y, i = threadIdx.x;
  const float* rec = scratch + ((size_t)s * gridDim.y + cg) * IC_REC;
  const float scale = ss ? ss[(size_t)s * ss_stride] + 1.f : 1.f, shift = ss ? ss[(size_t)s * ss_stride + 1] : 0.f;
  float v[2] = {0.f, 0.f};
  if (i < 56) {
    const int k = i % 7, ci = (i / 7) % 2, co = 4 * cg + i / 14;
    const float g = rec[i];
    if (ci == 0) {
      const float sv = rec[56 + i / 14] - rec[60 + (i / 14) * 7 + k];   // D - E
      const float wv = w[((size_t)co * 2 + 0) * 7 + k];
      atomicAdd(dw + ((size_t)co * 2 + 0) * 7 + k, scale * g + shift * sv);
      v[0] = wv * g;
      v[1] = wv * sv;
    } else {
      atomicAdd(dw + ((size_t)co * 2 + 1) * 7 + k, g);
    }
  } else if (i < 60) {
    if (db) atomicAdd(db + 4 * cg + (i - 56), rec[i]);
  }
  const float tot = block_reduce_vec<2>(v, red);
  if (ss && dss) {
    if (i == 0) atomicAdd(dss + (size_t)s * ss_stride, tot);
    if (i == 1) atomicAdd(dss + (size_t)s * ss_stride + 1, tot);
  }
}

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

// Backward of init_conv (Conv1d(2 -> cout, k7, pad 3) over cat(cond * (scale + 1) + shift, x)): dW, db and the
// per-sample d scale / d shift of the ConditionalScaleShift in one pass, no data gradient tensor.  scratch: (b, cout / 4,
// 88) floats, ZEROED by the caller.  ss / dss: scale at [s * ss_stride], shift at [s * ss_stride + 1].  Returns 1 (nothing
// launched) unless cout % 4 == 0, L % 4 == 0 and the rows are 16-byte aligned.
DQ_API int dq_initconv_bwd(const float* d, const float* cond, const float* x, const float* ss, int ss_stride, const float* w,
                           float* dw, float* db, float* dss, float* scratch, int cout, int R, int L, int rows_per_sample,
                           void* stream) {
  if (R <= 0 || L <= 0) return 0;
  if ((cout & 3) || (L & 3) || ((((size_t)d | (size_t)cond | (size_t)x) & 15) != 0) || R % rows_per_sample) return 1;
  cudaStream_t st = (cudaStream_t)stream;
  const int nchunk = (L + 8191) / 8192;
  const int chunk = ((L + nchunk - 1) / nchunk + 511) / 512 * 512;
  dim3 grid((unsigned)((L + chunk - 1) / chunk), (unsigned)R, (unsigned)(cout / 4));
  dq::initconv_bwd_kernel<<<grid, 128, 0, st>>>(d, cond, x, scratch, cout, L, chunk, rows_per_sample);
  DQ_LAUNCH_CHECK();
  dim3 g2((unsigned)(R / rows_per_sample), (unsigned)(cout / 4));
  dq::initconv_bwd_finalize_kernel<<<g2, 64, 0, st>>>(scratch, w, ss, ss_stride, dw, db, dss, cout);
  DQ_LAUNCH_CHECK();
  return 0;
}
