// Tensor-core (mma.sync m16n8k8 TF32, fp32 accumulate) kernels of the fused LinearAttention
// (reference /root/reference/dquartic/model/unet1d.py:473-496 inside Residual(PreNorm(.)), 1017/1068).
//
// One warp = one head.  All per-position intermediates (q/k/v projections, softmax numerators, the 32x32 per-head
// contractions, the output projection) live in MMA register fragments, FlashAttention-style: the accumulator
// fragment of one MMA is re-used directly as the A (or B) operand of the next, so nothing but the normalised
// C-channel input tile is staged in shared memory and nothing of size 128 x L ever exists.
//
// Fragment conventions (PTX m16n8k8 .tf32, g = lane/4, t = lane%4):
//   C/D: c0 (g, 2t) c1 (g, 2t+1) c2 (g+8, 2t) c3 (g+8, 2t+1)
//   A  : a0 (g, k=t) a1 (g+8, t) a2 (g, t+4) a3 (g+8, t+4)        B: b0 (k=t, n=g) b1 (k=t+4, n=g)
// We use a fixed permutation of the k index inside every k8 block: slot t <-> actual k = 2t, slot t+4 <-> 2t+1.
// Then (c0, c2, c1, c3) of an accumulator tile IS an A fragment, (c0, c1) / (c2, c3) ARE B fragments of the
// transposed matrix, and operands that come from memory are simply loaded with the same permutation (one 64-bit load).
#include "linattn.cuh"

namespace dq {

template <int C>
struct TC {
  static constexpr int KC = (C + 7) / 8;                              // k8 steps over the C input channels
  static constexpr int CP = KC * 8;                                   // padded channel count
  static constexpr int XS = (CP == 8) ? 8 : (CP <= 24 ? 24 : 40);     // smem row stride: = 8 or 24 (mod 32) -> conflict-free
  static constexpr int CT = KC;                                       // n8 tiles over C output channels
};
constexpr int SP = 128;  // positions per staged sub-tile (= threads per CTA)

__device__ __forceinline__ uint32_t f2tf(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma8(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                     uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ float quad_max(float v) {
  v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
  return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
  v += __shfl_xor_sync(0xffffffffu, v, 1);
  return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

// Thread j normalises position n0 + j (RMSNorm over C with gain g) and writes the TF32-rounded row xn_s[j][0..CP).
template <int C>
__device__ __forceinline__ void stage_xn(const float* __restrict__ x, const float* __restrict__ g, int r, int L, int n0,
                                         int n_end, float* xn_s, float* inv_s) {
  using T = TC<C>;
  const int j = threadIdx.x, n = n0 + j;
  const bool ok = n < n_end;
  float v[C];
  float s2 = 0.f;
#pragma unroll
  for (int c = 0; c < C; ++c) {
    v[c] = ok ? __ldg(x + ((size_t)r * C + c) * L + n) : 0.f;
    s2 = fmaf(v[c], v[c], s2);
  }
  const float inv = 1.f / fmaxf(sqrtf(s2), 1e-12f);
  const float sc = inv * sqrtf((float)C);
  float* row = xn_s + j * T::XS;
#pragma unroll
  for (int c = 0; c < T::CP; ++c) row[c] = (c < C) ? __uint_as_float(f2tf(v[c < C ? c : 0] * sc * g[c < C ? c : 0])) : 0.f;
  if (inv_s) inv_s[j] = inv;
}

// B fragments of X^T (k = channel, n = position) for the two n8 tiles of slab s:  bx[j][ks] = {xn[pos][8ks+2t], [..+1]}
template <int C>
__device__ __forceinline__ void load_bx(const float* xn_s, int s, int g, int t, uint32_t (&bx)[2][TC<C>::KC][2]) {
  using T = TC<C>;
#pragma unroll
  for (int j = 0; j < 2; ++j)
#pragma unroll
    for (int ks = 0; ks < T::KC; ++ks) {
      float2 v = *reinterpret_cast<const float2*>(xn_s + (16 * s + 8 * j + g) * T::XS + 8 * ks + 2 * t);
      bx[j][ks][0] = __float_as_uint(v.x);
      bx[j][ks][1] = __float_as_uint(v.y);
    }
}
// A fragments of X (rows = positions g, g+8 of slab s; k = channel)
template <int C>
__device__ __forceinline__ void load_ax(const float* xn_s, int s, int g, int t, uint32_t (&ax)[TC<C>::KC][4]) {
  using T = TC<C>;
#pragma unroll
  for (int ks = 0; ks < T::KC; ++ks) {
    float2 lo = *reinterpret_cast<const float2*>(xn_s + (16 * s + g) * T::XS + 8 * ks + 2 * t);
    float2 hi = *reinterpret_cast<const float2*>(xn_s + (16 * s + 8 + g) * T::XS + 8 * ks + 2 * t);
    ax[ks][0] = __float_as_uint(lo.x); ax[ks][2] = __float_as_uint(lo.y);
    ax[ks][1] = __float_as_uint(hi.x); ax[ks][3] = __float_as_uint(hi.y);
  }
}

// ------------------------------------------------------------------------------------------- forward: stats
// K^T = Wk_h X^T and V^T = Wv_h X^T as (channels x positions) tiles; softmax over positions is then a row-wise
// online softmax; ctx[d][e] += P[d][n] V[e][n] re-uses the K^T accumulators as A and the V^T accumulators as B.
template <int C>
__global__ void __launch_bounds__(128) la_stats_tc_kernel(LAArgs a) {
  using T = TC<C>;
  __shared__ __align__(16) float xn_s[SP * T::XS];
  const int lane = threadIdx.x & 31, h = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int r = blockIdx.y, ch = blockIdx.x;
  const int n_begin = ch * a.chunk, n_end = min(a.L, n_begin + a.chunk);

  uint32_t wk[2][T::KC][4], wv[2][T::KC][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ks = 0; ks < T::KC; ++ks)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = h * 32 + 16 * mt + g + 8 * (i & 1), c = 8 * ks + 2 * t + (i >> 1);
        wk[mt][ks][i] = f2tf(c < C ? a.wqkv[(size_t)(kHD + d) * C + c] : 0.f);
        wv[mt][ks][i] = f2tf(c < C ? a.wqkv[(size_t)(2 * kHD + d) * C + c] : 0.f);
      }
  float m_run[2][2], s_run[2][2], ctx[2][4][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
    m_run[mt][0] = m_run[mt][1] = -INFINITY;
    s_run[mt][0] = s_run[mt][1] = 0.f;
#pragma unroll
    for (int ev = 0; ev < 4; ++ev)
#pragma unroll
      for (int i = 0; i < 4; ++i) ctx[mt][ev][i] = 0.f;
  }

  for (int n0 = n_begin; n0 < n_end; n0 += SP) {
    stage_xn<C>(a.x, a.g_pre, r, a.L, n0, n_end, xn_s, nullptr);
    __syncthreads();
    const int nslab = min(SP / 16, (n_end - n0 + 15) / 16);
    // pass A: row maxima of K^T over this sub-tile
    float mloc[2][2] = {{-INFINITY, -INFINITY}, {-INFINITY, -INFINITY}};
    for (int s = 0; s < nslab; ++s) {
      uint32_t bx[2][T::KC][2];
      load_bx<C>(xn_s, s, g, t, bx);
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < T::KC; ++ks) mma8(acc, wk[mt][ks][0], wk[mt][ks][1], wk[mt][ks][2], wk[mt][ks][3], bx[j][ks][0], bx[j][ks][1]);
          const int nb = n0 + 16 * s + 8 * j + 2 * t;
          const bool v0 = nb < n_end, v1 = nb + 1 < n_end;
          mloc[mt][0] = fmaxf(mloc[mt][0], fmaxf(v0 ? acc[0] : -INFINITY, v1 ? acc[1] : -INFINITY));
          mloc[mt][1] = fmaxf(mloc[mt][1], fmaxf(v0 ? acc[2] : -INFINITY, v1 ? acc[3] : -INFINITY));
        }
    }
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const float m_new = fmaxf(m_run[mt][hf], quad_max(mloc[mt][hf]));
        const float f = __expf(m_run[mt][hf] - m_new);
        s_run[mt][hf] *= f;
#pragma unroll
        for (int ev = 0; ev < 4; ++ev) { ctx[mt][ev][2 * hf] *= f; ctx[mt][ev][2 * hf + 1] *= f; }
        m_run[mt][hf] = m_new;
      }
    // pass B: P = exp(K^T - m), row sums, ctx += P V^T
    for (int s = 0; s < nslab; ++s) {
      uint32_t bx[2][T::KC][2];
      load_bx<C>(xn_s, s, g, t, bx);
      float kacc[2][2][4], vacc[2][2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
#pragma unroll
          for (int i = 0; i < 4; ++i) { kacc[mt][j][i] = 0.f; vacc[mt][j][i] = 0.f; }
#pragma unroll
          for (int ks = 0; ks < T::KC; ++ks) {
            mma8(kacc[mt][j], wk[mt][ks][0], wk[mt][ks][1], wk[mt][ks][2], wk[mt][ks][3], bx[j][ks][0], bx[j][ks][1]);
            mma8(vacc[mt][j], wv[mt][ks][0], wv[mt][ks][1], wv[mt][ks][2], wv[mt][ks][3], bx[j][ks][0], bx[j][ks][1]);
          }
          const int nb = n0 + 16 * s + 8 * j + 2 * t;
          const bool v0 = nb < n_end, v1 = nb + 1 < n_end;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool ok = (i & 1) ? v1 : v0;
            const float p = ok ? __expf(kacc[mt][j][i] - m_run[mt][i >> 1]) : 0.f;
            s_run[mt][i >> 1] += p;
            kacc[mt][j][i] = p;
          }
        }
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const uint32_t a0 = f2tf(kacc[mt][j][0]), a1 = f2tf(kacc[mt][j][2]), a2 = f2tf(kacc[mt][j][1]), a3 = f2tf(kacc[mt][j][3]);
#pragma unroll
          for (int ev = 0; ev < 4; ++ev) {
            const float* vv = vacc[ev >> 1][j];
            const uint32_t b0 = f2tf((ev & 1) ? vv[2] : vv[0]), b1 = f2tf((ev & 1) ? vv[3] : vv[1]);
            mma8(ctx[mt][ev], a0, a1, a2, a3, b0, b1);
          }
        }
    }
    __syncthreads();
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const float s = quad_sum(s_run[mt][hf]);
      const int d = h * 32 + 16 * mt + 8 * hf + g;
      float* po = a.part + (((size_t)r * a.nchunk + ch) * kHD + d) * 34;
      if (t == 0) { po[0] = m_run[mt][hf]; po[1] = s; }
#pragma unroll
      for (int ev = 0; ev < 4; ++ev)
        *reinterpret_cast<float2*>(po + 2 + 8 * ev + 2 * t) = make_float2(ctx[mt][ev][2 * hf], ctx[mt][ev][2 * hf + 1]);
    }
}

// ------------------------------------------------------------------------------------------- forward: output
// Per 16-position slab and head: Q = X Wq^T (positions x d) -> softmax over d (quad shuffles) -> O = Qs Ctx
// -> Y_h = O Wout_h^T, all in fragments; the four heads' Y_h meet in shared memory for bias + RMSNorm + residual.
template <int C>
__global__ void __launch_bounds__(128) la_out_tc_kernel(LAArgs a) {
  using T = TC<C>;
  constexpr int YS = T::CP + 4;  // row stride of the per-head partial-Y tile
  extern __shared__ float4 dyn_smem4[];
  float* xn_s = reinterpret_cast<float*>(dyn_smem4);  // SP * XS
  float* yp_s = xn_s + SP * T::XS;                    // 4 * SP * YS
  const int lane = threadIdx.x & 31, h = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int r = blockIdx.y;
  const int n_begin = blockIdx.x * a.chunk, n_end = min(a.L, n_begin + a.chunk);
  const float scale = rsqrtf((float)kDimHead);

  uint32_t bq[4][T::KC][2], bc[4][4][2], bo[4][T::CT][2];
#pragma unroll
  for (int dt = 0; dt < 4; ++dt)
#pragma unroll
    for (int ks = 0; ks < T::KC; ++ks)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int d = h * 32 + 8 * dt + g, c = 8 * ks + 2 * t + i;
        bq[dt][ks][i] = f2tf(c < C ? a.wqkv[(size_t)d * C + c] : 0.f);
      }
#pragma unroll
  for (int kd = 0; kd < 4; ++kd)
#pragma unroll
    for (int ev = 0; ev < 4; ++ev)
#pragma unroll
      for (int i = 0; i < 2; ++i)
        bc[kd][ev][i] = f2tf(a.ctx[((size_t)r * kHD + h * 32 + 8 * kd + 2 * t + i) * 32 + 8 * ev + g]);
#pragma unroll
  for (int ke = 0; ke < 4; ++ke)
#pragma unroll
    for (int ct = 0; ct < T::CT; ++ct)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = 8 * ct + g, e = h * 32 + 8 * ke + 2 * t + i;
        bo[ke][ct][i] = f2tf(c < C ? a.wout[(size_t)c * kHD + e] : 0.f);
      }

  for (int n0 = n_begin; n0 < n_end; n0 += SP) {
    stage_xn<C>(a.x, a.g_pre, r, a.L, n0, n_end, xn_s, nullptr);
    __syncthreads();
    const int nslab = min(SP / 16, (n_end - n0 + 15) / 16);
    for (int s = 0; s < nslab; ++s) {
      uint32_t ax[T::KC][4];
      load_ax<C>(xn_s, s, g, t, ax);
      float q[4][4];
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        q[dt][0] = q[dt][1] = q[dt][2] = q[dt][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < T::KC; ++ks) mma8(q[dt], ax[ks][0], ax[ks][1], ax[ks][2], ax[ks][3], bq[dt][ks][0], bq[dt][ks][1]);
      }
      // softmax over d for rows g (elements 0,1) and g+8 (elements 2,3)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float mx = -INFINITY;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) mx = fmaxf(mx, fmaxf(q[dt][2 * hf], q[dt][2 * hf + 1]));
        mx = quad_max(mx);
        float sm = 0.f;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) {
          q[dt][2 * hf] = __expf(q[dt][2 * hf] - mx);
          q[dt][2 * hf + 1] = __expf(q[dt][2 * hf + 1] - mx);
          sm += q[dt][2 * hf] + q[dt][2 * hf + 1];
        }
        const float f = scale / quad_sum(sm);
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) { q[dt][2 * hf] *= f; q[dt][2 * hf + 1] *= f; }
      }
      float o[4][4];
#pragma unroll
      for (int ev = 0; ev < 4; ++ev) o[ev][0] = o[ev][1] = o[ev][2] = o[ev][3] = 0.f;
#pragma unroll
      for (int kd = 0; kd < 4; ++kd) {
        const uint32_t a0 = f2tf(q[kd][0]), a1 = f2tf(q[kd][2]), a2 = f2tf(q[kd][1]), a3 = f2tf(q[kd][3]);
#pragma unroll
        for (int ev = 0; ev < 4; ++ev) mma8(o[ev], a0, a1, a2, a3, bc[kd][ev][0], bc[kd][ev][1]);
      }
      float y[T::CT][4];
#pragma unroll
      for (int ct = 0; ct < T::CT; ++ct) y[ct][0] = y[ct][1] = y[ct][2] = y[ct][3] = 0.f;
#pragma unroll
      for (int ke = 0; ke < 4; ++ke) {
        const uint32_t a0 = f2tf(o[ke][0]), a1 = f2tf(o[ke][2]), a2 = f2tf(o[ke][1]), a3 = f2tf(o[ke][3]);
#pragma unroll
        for (int ct = 0; ct < T::CT; ++ct) mma8(y[ct], a0, a1, a2, a3, bo[ke][ct][0], bo[ke][ct][1]);
      }
#pragma unroll
      for (int ct = 0; ct < T::CT; ++ct) {
        float* p0 = yp_s + ((size_t)h * SP + 16 * s + g) * YS + 8 * ct + 2 * t;
        *reinterpret_cast<float2*>(p0) = make_float2(y[ct][0], y[ct][1]);
        *reinterpret_cast<float2*>(p0 + 8 * YS) = make_float2(y[ct][2], y[ct][3]);
      }
    }
    __syncthreads();
    {
      const int j = threadIdx.x, n = n0 + j;
      if (n < n_end) {
        float y[C];
        float s2 = 0.f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          y[c] = a.bout[c] + yp_s[(0 * SP + j) * YS + c] + yp_s[(1 * SP + j) * YS + c] + yp_s[(2 * SP + j) * YS + c] +
                 yp_s[(3 * SP + j) * YS + c];
          s2 = fmaf(y[c], y[c], s2);
        }
        const float sc = sqrtf((float)C) / fmaxf(sqrtf(s2), 1e-12f);
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const size_t idx = ((size_t)r * C + c) * a.L + n;
          if (a.ypre) a.ypre[idx] = y[c];
          a.out[idx] = fmaf(y[c] * sc, a.g_out[c], __ldg(a.x + idx));
        }
      }
    }
    __syncthreads();
  }
}

template <int C>
static int fwd_tc(const LAArgs& a, cudaStream_t st) {
  using T = TC<C>;
  dim3 grid((unsigned)a.nchunk, (unsigned)a.R);
  la_stats_tc_kernel<C><<<grid, 128, 0, st>>>(a);
  DQ_LAUNCH_CHECK();
  la_combine_launch(a, st);
  DQ_LAUNCH_CHECK();
  size_t smem = sizeof(float) * (SP * T::XS + 4 * SP * (T::CP + 4));
  cudaFuncSetAttribute(la_out_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  la_out_tc_kernel<C><<<grid, 128, smem, st>>>(a);
  DQ_LAUNCH_CHECK();
  return 0;
}

int la_fwd_tc_dispatch(const LAArgs& a, int C, cudaStream_t st) {
  switch (C) {
    case 4: return fwd_tc<4>(a, st);
    case 8: return fwd_tc<8>(a, st);
    case 12: return fwd_tc<12>(a, st);
    case 16: return fwd_tc<16>(a, st);
    case 24: return fwd_tc<24>(a, st);
    case 32: return fwd_tc<32>(a, st);
    default: return -3;
  }
}

}  // namespace dq

namespace dq {

constexpr int RS = 36;  // row stride of the warp-private 16 x 32 transpose tiles (2*RS = 8 mod 32: conflict-free reads)

// store an accumulator tile set v[4][4] (rows = positions g / g+8 of the slab, cols = channel 8*tile + 2t, +1)
__device__ __forceinline__ void store_tile16x32(float* scr, const float (&v)[4][4], int g, int t) {
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
    *reinterpret_cast<float2*>(scr + g * RS + 8 * dt + 2 * t) = make_float2(v[dt][0], v[dt][1]);
    *reinterpret_cast<float2*>(scr + (g + 8) * RS + 8 * dt + 2 * t) = make_float2(v[dt][2], v[dt][3]);
  }
}
// A fragment of the TRANSPOSED tile: rows = channels 16*mt + g (+8), k = positions 8*j + 2t (+1)
__device__ __forceinline__ void load_At(const float* scr, int mt, int j, int g, int t, uint32_t (&A)[4]) {
  const float* p0 = scr + (8 * j + 2 * t) * RS + 16 * mt + g;
  A[0] = f2tf(p0[0]); A[1] = f2tf(p0[8]); A[2] = f2tf(p0[RS]); A[3] = f2tf(p0[RS + 8]);
}

// ------------------------------------------------------------------------------------------- backward: q path
// Per slab and head (fragments): Q -> Qs, Do = dY Wout_h, dQs = Do Ctx^T, dQr (softmax backward), dXn += dQr Wq.
// Reductions over positions (dCtx = Qs^T Do, G^T = Qs^T dY, dWq = dQr^T Xn) read Qs / Do / dQr back from
// warp-private shared tiles in the transposed role.  dWout = (Ctx^T G^T)^T is formed once per CTA at the end.
template <int C>
__global__ void __launch_bounds__(128) la_bwd_q_tc_kernel(LAArgs a) {
  using T = TC<C>;
  constexpr int YS = T::CP + 4;
  extern __shared__ float4 dyn_smem4[];
  float* xn_s = reinterpret_cast<float*>(dyn_smem4);   // SP * XS
  float* dy_s = xn_s + SP * T::XS;                     // SP * XS
  float* yp_s = dy_s + SP * T::XS;                     // SP * YS   (d xn_q, summed over heads with shared atomics)
  float* scr = yp_s + SP * YS;                         // 4 warps * 3 tiles * 16 * RS
  float* g_s = scr + 4 * 3 * 16 * RS;                  // 4 * 32 * CP
  float* acc_s = g_s + 4 * 32 * T::CP;                 // 2 * C  (d g_out, d b_out)
  const int lane = threadIdx.x & 31, h = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int r = blockIdx.y, ch = blockIdx.x;
  const int n_begin = ch * a.chunk, n_end = min(a.L, n_begin + a.chunk);
  const float scale = rsqrtf((float)kDimHead), inv_scale = sqrtf((float)kDimHead);
  const float sqrtC = sqrtf((float)C);
  float* scrQ = scr + (h * 3 + 0) * 16 * RS;
  float* scrD = scr + (h * 3 + 1) * 16 * RS;
  float* scrR = scr + (h * 3 + 2) * 16 * RS;

  uint32_t bq[4][T::KC][2], bwo[4][T::KC][2], bcA[4][4][2], bqT[4][T::CT][2];
#pragma unroll
  for (int dt = 0; dt < 4; ++dt)
#pragma unroll
    for (int ks = 0; ks < T::KC; ++ks)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = 8 * ks + 2 * t + i;
        bq[dt][ks][i] = f2tf(c < C ? a.wqkv[(size_t)(h * 32 + 8 * dt + g) * C + c] : 0.f);
        bwo[dt][ks][i] = f2tf(c < C ? a.wout[(size_t)c * kHD + h * 32 + 8 * dt + g] : 0.f);
      }
#pragma unroll
  for (int dt = 0; dt < 4; ++dt)
#pragma unroll
    for (int ke = 0; ke < 4; ++ke)
#pragma unroll
      for (int i = 0; i < 2; ++i)
        bcA[dt][ke][i] = f2tf(a.ctx[((size_t)r * kHD + h * 32 + 8 * dt + g) * 32 + 8 * ke + 2 * t + i]);
#pragma unroll
  for (int kd = 0; kd < 4; ++kd)
#pragma unroll
    for (int ct = 0; ct < T::CT; ++ct)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = 8 * ct + g;
        bqT[kd][ct][i] = f2tf(c < C ? a.wqkv[(size_t)(h * 32 + 8 * kd + 2 * t + i) * C + c] : 0.f);
      }
  float dctx[2][4][4], gT[2][T::CT][4], dwq[2][T::CT][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt) {
#pragma unroll
    for (int ev = 0; ev < 4; ++ev)
#pragma unroll
      for (int i = 0; i < 4; ++i) dctx[mt][ev][i] = 0.f;
#pragma unroll
    for (int ct = 0; ct < T::CT; ++ct)
#pragma unroll
      for (int i = 0; i < 4; ++i) { gT[mt][ct][i] = 0.f; dwq[mt][ct][i] = 0.f; }
  }
  if (threadIdx.x < 2 * C) acc_s[threadIdx.x] = 0.f;
  __syncthreads();

  for (int n0 = n_begin; n0 < n_end; n0 += SP) {
    stage_xn<C>(a.x, a.g_pre, r, a.L, n0, n_end, xn_s, nullptr);
    {  // d y = RMSNorm_out backward of d res, thread j = position
      const int j = threadIdx.x, n = n0 + j;
      const bool ok = n < n_end;
      float y[C], dr[C];
      float s2 = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const size_t idx = ((size_t)r * C + c) * a.L + n;
        y[c] = ok ? __ldg(a.ypre + idx) : 0.f;
        dr[c] = ok ? __ldg(a.dres + idx) : 0.f;
        s2 = fmaf(y[c], y[c], s2);
      }
      const float nrm = sqrtf(s2), inv = 1.f / fmaxf(nrm, 1e-12f);
      float dot = 0.f, dgl[C];
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const float uh = y[c] * inv;
        dgl[c] = dr[c] * uh * sqrtC;
        const float duh = dr[c] * a.g_out[c] * sqrtC;
        dot = fmaf(duh, uh, dot);
        y[c] = uh; dr[c] = duh;
      }
      float* row = dy_s + j * T::XS;
#pragma unroll
      for (int c = 0; c < T::CP; ++c) {
        float d = 0.f;
        if (c < C) {
          d = (nrm > 1e-12f) ? (dr[c < C ? c : 0] - y[c < C ? c : 0] * dot) * inv : dr[c < C ? c : 0] * inv;
          d = ok ? d : 0.f;
        }
        row[c] = __uint_as_float(f2tf(d));
        if (c < C) {
          const float s1 = warp_sum(dgl[c < C ? c : 0]), s2b = warp_sum(d);
          if (lane == 0) { atomicAdd(acc_s + c, s1); atomicAdd(acc_s + C + c, s2b); }
        }
      }
      float* yr = yp_s + j * YS;
#pragma unroll
      for (int c = 0; c < YS; ++c) yr[c] = 0.f;
    }
    __syncthreads();
    const int nslab = min(SP / 16, (n_end - n0 + 15) / 16);
    for (int s = 0; s < nslab; ++s) {
      float qs[4][4], dd[4][4], dqs[4][4];
      {
        uint32_t ax[T::KC][4];
        load_ax<C>(xn_s, s, g, t, ax);
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) {
          qs[dt][0] = qs[dt][1] = qs[dt][2] = qs[dt][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < T::KC; ++ks) mma8(qs[dt], ax[ks][0], ax[ks][1], ax[ks][2], ax[ks][3], bq[dt][ks][0], bq[dt][ks][1]);
        }
        load_ax<C>(dy_s, s, g, t, ax);
#pragma unroll
        for (int et = 0; et < 4; ++et) {
          dd[et][0] = dd[et][1] = dd[et][2] = dd[et][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < T::KC; ++ks) mma8(dd[et], ax[ks][0], ax[ks][1], ax[ks][2], ax[ks][3], bwo[et][ks][0], bwo[et][ks][1]);
        }
      }
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {  // softmax over d, times scale
        float mx = -INFINITY;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) mx = fmaxf(mx, fmaxf(qs[dt][2 * hf], qs[dt][2 * hf + 1]));
        mx = quad_max(mx);
        float sm = 0.f;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) {
          qs[dt][2 * hf] = __expf(qs[dt][2 * hf] - mx);
          qs[dt][2 * hf + 1] = __expf(qs[dt][2 * hf + 1] - mx);
          sm += qs[dt][2 * hf] + qs[dt][2 * hf + 1];
        }
        const float f = scale / quad_sum(sm);
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) { qs[dt][2 * hf] *= f; qs[dt][2 * hf + 1] *= f; }
      }
      // dQs = Do Ctx^T
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) dqs[dt][0] = dqs[dt][1] = dqs[dt][2] = dqs[dt][3] = 0.f;
#pragma unroll
      for (int ke = 0; ke < 4; ++ke) {
        const uint32_t a0 = f2tf(dd[ke][0]), a1 = f2tf(dd[ke][2]), a2 = f2tf(dd[ke][1]), a3 = f2tf(dd[ke][3]);
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) mma8(dqs[dt], a0, a1, a2, a3, bcA[dt][ke][0], bcA[dt][ke][1]);
      }
      store_tile16x32(scrQ, qs, g, t);
      store_tile16x32(scrD, dd, g, t);
      // softmax backward: dQr = Qs * (dQs - sum_d Qs dQs / scale)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        float ts = 0.f;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) ts += qs[dt][2 * hf] * dqs[dt][2 * hf] + qs[dt][2 * hf + 1] * dqs[dt][2 * hf + 1];
        ts = quad_sum(ts) * inv_scale;
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) {
          dqs[dt][2 * hf] = qs[dt][2 * hf] * (dqs[dt][2 * hf] - ts);
          dqs[dt][2 * hf + 1] = qs[dt][2 * hf + 1] * (dqs[dt][2 * hf + 1] - ts);
        }
      }
      store_tile16x32(scrR, dqs, g, t);
      {  // d xn_q += dQr Wq_h  -> shared accumulation over heads
        float dxn[T::CT][4];
#pragma unroll
        for (int ct = 0; ct < T::CT; ++ct) dxn[ct][0] = dxn[ct][1] = dxn[ct][2] = dxn[ct][3] = 0.f;
#pragma unroll
        for (int kd = 0; kd < 4; ++kd) {
          const uint32_t a0 = f2tf(dqs[kd][0]), a1 = f2tf(dqs[kd][2]), a2 = f2tf(dqs[kd][1]), a3 = f2tf(dqs[kd][3]);
#pragma unroll
          for (int ct = 0; ct < T::CT; ++ct) mma8(dxn[ct], a0, a1, a2, a3, bqT[kd][ct][0], bqT[kd][ct][1]);
        }
#pragma unroll
        for (int ct = 0; ct < T::CT; ++ct) {
          float* p0 = yp_s + (16 * s + g) * YS + 8 * ct + 2 * t;
          atomicAdd(p0, dxn[ct][0]); atomicAdd(p0 + 1, dxn[ct][1]);
          atomicAdd(p0 + 8 * YS, dxn[ct][2]); atomicAdd(p0 + 8 * YS + 1, dxn[ct][3]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t Bd[4][2], Bdy[T::CT][2], Bxn[T::CT][2];
#pragma unroll
        for (int ev = 0; ev < 4; ++ev) {
          Bd[ev][0] = f2tf(scrD[(8 * j + 2 * t) * RS + 8 * ev + g]);
          Bd[ev][1] = f2tf(scrD[(8 * j + 2 * t + 1) * RS + 8 * ev + g]);
        }
#pragma unroll
        for (int ct = 0; ct < T::CT; ++ct) {
          const int p = 16 * s + 8 * j + 2 * t;
          Bdy[ct][0] = __float_as_uint(dy_s[p * T::XS + 8 * ct + g]);
          Bdy[ct][1] = __float_as_uint(dy_s[(p + 1) * T::XS + 8 * ct + g]);
          Bxn[ct][0] = __float_as_uint(xn_s[p * T::XS + 8 * ct + g]);
          Bxn[ct][1] = __float_as_uint(xn_s[(p + 1) * T::XS + 8 * ct + g]);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          uint32_t Aq[4], Ar[4];
          load_At(scrQ, mt, j, g, t, Aq);
          load_At(scrR, mt, j, g, t, Ar);
#pragma unroll
          for (int ev = 0; ev < 4; ++ev) mma8(dctx[mt][ev], Aq[0], Aq[1], Aq[2], Aq[3], Bd[ev][0], Bd[ev][1]);
#pragma unroll
          for (int ct = 0; ct < T::CT; ++ct) {
            mma8(gT[mt][ct], Aq[0], Aq[1], Aq[2], Aq[3], Bdy[ct][0], Bdy[ct][1]);
            mma8(dwq[mt][ct], Ar[0], Ar[1], Ar[2], Ar[3], Bxn[ct][0], Bxn[ct][1]);
          }
        }
      }
      __syncwarp();
    }
    __syncthreads();
    {
      const int j = threadIdx.x, n = n0 + j;
      if (n < n_end) {
#pragma unroll
        for (int c = 0; c < C; ++c) a.dxnq[((size_t)r * C + c) * a.L + n] = yp_s[j * YS + c];
      }
    }
    __syncthreads();
  }
  // d ctx partial of this chunk
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      float* dp = a.dpart + (((size_t)r * a.nchunk + ch) * kHD + h * 32 + 16 * mt + 8 * hf + g) * 32;
#pragma unroll
      for (int ev = 0; ev < 4; ++ev)
        *reinterpret_cast<float2*>(dp + 8 * ev + 2 * t) = make_float2(dctx[mt][ev][2 * hf], dctx[mt][ev][2 * hf + 1]);
    }
  // d Wq (rows d, cols c)
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ct = 0; ct < T::CT; ++ct)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = h * 32 + 16 * mt + g + 8 * (i >> 1), c = 8 * ct + 2 * t + (i & 1);
        if (c < C) atomicAdd(a.dwqkv + (size_t)d * C + c, dwq[mt][ct][i]);
      }
  // d Wout[c][h*32+e] = sum_d ctx[d][e] G^T[d][c]
  float* gh = g_s + h * 32 * T::CP;
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ct = 0; ct < T::CT; ++ct)
#pragma unroll
      for (int i = 0; i < 4; ++i) gh[(16 * mt + g + 8 * (i >> 1)) * T::CP + 8 * ct + 2 * t + (i & 1)] = gT[mt][ct][i];
  __syncwarp();
  {
    const int e = lane;
    float accw[C];
#pragma unroll
    for (int c = 0; c < C; ++c) accw[c] = 0.f;
    for (int d = 0; d < 32; ++d) {
      const float cv = a.ctx[((size_t)r * kHD + h * 32 + d) * 32 + e];
#pragma unroll
      for (int c = 0; c < C; ++c) accw[c] = fmaf(cv, gh[d * T::CP + c], accw[c]);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) atomicAdd(a.dwout + (size_t)c * kHD + h * 32 + e, accw[c]);
  }
  __syncthreads();
  if (threadIdx.x < C) atomicAdd(a.dg_out + threadIdx.x, acc_s[threadIdx.x]);
  else if (threadIdx.x < 2 * C) atomicAdd(a.dbout + threadIdx.x - C, acc_s[threadIdx.x]);
}

// ------------------------------------------------------------------------------------------- backward: k/v path
template <int C>
__global__ void __launch_bounds__(128) la_bwd_kv_tc_kernel(LAArgs a) {
  using T = TC<C>;
  constexpr int YS = T::CP + 4;
  extern __shared__ float4 dyn_smem4[];
  float* xn_s = reinterpret_cast<float*>(dyn_smem4);   // SP * XS
  float* yp_s = xn_s + SP * T::XS;                     // SP * YS
  float* scr = yp_s + SP * YS;                         // 4 warps * 2 tiles * 16 * RS
  float* inv_s = scr + 4 * 2 * 16 * RS;                // SP
  float* acc_s = inv_s + SP;                           // C (d g_pre)
  const int lane = threadIdx.x & 31, h = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int r = blockIdx.y;
  const int n_begin = blockIdx.x * a.chunk, n_end = min(a.L, n_begin + a.chunk);
  const float sqrtC = sqrtf((float)C);
  float* scrK = scr + (h * 2 + 0) * 16 * RS;
  float* scrV = scr + (h * 2 + 1) * 16 * RS;

  uint32_t bwk[4][T::KC][2], bwv[4][T::KC][2], bdA[4][4][2], bdT[4][4][2], bkT[4][T::CT][2], bvT[4][T::CT][2];
  float cm[4][2], cs[4][2], cd[4][2];
#pragma unroll
  for (int dt = 0; dt < 4; ++dt) {
#pragma unroll
    for (int ks = 0; ks < T::KC; ++ks)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = 8 * ks + 2 * t + i;
        bwk[dt][ks][i] = f2tf(c < C ? a.wqkv[(size_t)(kHD + h * 32 + 8 * dt + g) * C + c] : 0.f);
        bwv[dt][ks][i] = f2tf(c < C ? a.wqkv[(size_t)(2 * kHD + h * 32 + 8 * dt + g) * C + c] : 0.f);
      }
#pragma unroll
    for (int k4 = 0; k4 < 4; ++k4)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        bdA[dt][k4][i] = f2tf(a.dctx[((size_t)r * kHD + h * 32 + 8 * dt + g) * 32 + 8 * k4 + 2 * t + i]);   // [d-tile][k=e]
        bdT[dt][k4][i] = f2tf(a.dctx[((size_t)r * kHD + h * 32 + 8 * dt + 2 * t + i) * 32 + 8 * k4 + g]);   // [k=d][e-tile]
      }
#pragma unroll
    for (int ct = 0; ct < T::CT; ++ct)
#pragma unroll
      for (int i = 0; i < 2; ++i) {
        const int c = 8 * ct + g;
        bkT[dt][ct][i] = f2tf(c < C ? a.wqkv[(size_t)(kHD + h * 32 + 8 * dt + 2 * t + i) * C + c] : 0.f);
        bvT[dt][ct][i] = f2tf(c < C ? a.wqkv[(size_t)(2 * kHD + h * 32 + 8 * dt + 2 * t + i) * C + c] : 0.f);
      }
#pragma unroll
    for (int i = 0; i < 2; ++i) {
      const size_t jd = (size_t)r * kHD + h * 32 + 8 * dt + 2 * t + i;
      cm[dt][i] = a.ms[jd * 2];
      cs[dt][i] = 1.f / a.ms[jd * 2 + 1];
      cd[dt][i] = a.sd[jd];
    }
  }
  float dwk[2][T::CT][4], dwv[2][T::CT][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ct = 0; ct < T::CT; ++ct)
#pragma unroll
      for (int i = 0; i < 4; ++i) { dwk[mt][ct][i] = 0.f; dwv[mt][ct][i] = 0.f; }
  if (threadIdx.x < C) acc_s[threadIdx.x] = 0.f;

  for (int n0 = n_begin; n0 < n_end; n0 += SP) {
    stage_xn<C>(a.x, a.g_pre, r, a.L, n0, n_end, xn_s, inv_s);
    {
      float* yr = yp_s + threadIdx.x * YS;
#pragma unroll
      for (int c = 0; c < YS; ++c) yr[c] = 0.f;
    }
    __syncthreads();
    const int nslab = min(SP / 16, (n_end - n0 + 15) / 16);
    for (int s = 0; s < nslab; ++s) {
      float kk[4][4], vv[4][4], dks[4][4], dv[4][4];
      {
        uint32_t ax[T::KC][4];
        load_ax<C>(xn_s, s, g, t, ax);
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) {
          kk[dt][0] = kk[dt][1] = kk[dt][2] = kk[dt][3] = 0.f;
          vv[dt][0] = vv[dt][1] = vv[dt][2] = vv[dt][3] = 0.f;
#pragma unroll
          for (int ks = 0; ks < T::KC; ++ks) {
            mma8(kk[dt], ax[ks][0], ax[ks][1], ax[ks][2], ax[ks][3], bwk[dt][ks][0], bwk[dt][ks][1]);
            mma8(vv[dt], ax[ks][0], ax[ks][1], ax[ks][2], ax[ks][3], bwv[dt][ks][0], bwv[dt][ks][1]);
          }
#pragma unroll
          for (int i = 0; i < 4; ++i) kk[dt][i] = __expf(kk[dt][i] - cm[dt][i & 1]) * cs[dt][i & 1];  // softmax_L(k)
        }
      }
#pragma unroll
      for (int dt = 0; dt < 4; ++dt) {
        dks[dt][0] = dks[dt][1] = dks[dt][2] = dks[dt][3] = 0.f;
        dv[dt][0] = dv[dt][1] = dv[dt][2] = dv[dt][3] = 0.f;
      }
#pragma unroll
      for (int k4 = 0; k4 < 4; ++k4) {
        const uint32_t v0 = f2tf(vv[k4][0]), v1 = f2tf(vv[k4][2]), v2 = f2tf(vv[k4][1]), v3 = f2tf(vv[k4][3]);
        const uint32_t k0 = f2tf(kk[k4][0]), k1 = f2tf(kk[k4][2]), k2 = f2tf(kk[k4][1]), k3 = f2tf(kk[k4][3]);
#pragma unroll
        for (int dt = 0; dt < 4; ++dt) {
          mma8(dks[dt], v0, v1, v2, v3, bdA[dt][k4][0], bdA[dt][k4][1]);  // dKs[n][d] = sum_e V[n][e] dctx[d][e]
          mma8(dv[dt], k0, k1, k2, k3, bdT[k4][dt][0], bdT[k4][dt][1]);   // dV[n][e]  = sum_d Ks[n][d] dctx[d][e]
        }
      }
#pragma unroll
      for (int dt = 0; dt < 4; ++dt)
#pragma unroll
        for (int i = 0; i < 4; ++i) dks[dt][i] = kk[dt][i] * (dks[dt][i] - cd[dt][i & 1]);  // d k_raw
      store_tile16x32(scrK, dks, g, t);
      store_tile16x32(scrV, dv, g, t);
      {
        float dxn[T::CT][4];
#pragma unroll
        for (int ct = 0; ct < T::CT; ++ct) dxn[ct][0] = dxn[ct][1] = dxn[ct][2] = dxn[ct][3] = 0.f;
#pragma unroll
        for (int k4 = 0; k4 < 4; ++k4) {
          const uint32_t a0 = f2tf(dks[k4][0]), a1 = f2tf(dks[k4][2]), a2 = f2tf(dks[k4][1]), a3 = f2tf(dks[k4][3]);
          const uint32_t e0 = f2tf(dv[k4][0]), e1 = f2tf(dv[k4][2]), e2 = f2tf(dv[k4][1]), e3 = f2tf(dv[k4][3]);
#pragma unroll
          for (int ct = 0; ct < T::CT; ++ct) {
            mma8(dxn[ct], a0, a1, a2, a3, bkT[k4][ct][0], bkT[k4][ct][1]);
            mma8(dxn[ct], e0, e1, e2, e3, bvT[k4][ct][0], bvT[k4][ct][1]);
          }
        }
#pragma unroll
        for (int ct = 0; ct < T::CT; ++ct) {
          float* p0 = yp_s + (16 * s + g) * YS + 8 * ct + 2 * t;
          atomicAdd(p0, dxn[ct][0]); atomicAdd(p0 + 1, dxn[ct][1]);
          atomicAdd(p0 + 8 * YS, dxn[ct][2]); atomicAdd(p0 + 8 * YS + 1, dxn[ct][3]);
        }
      }
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        uint32_t Bxn[T::CT][2];
#pragma unroll
        for (int ct = 0; ct < T::CT; ++ct) {
          const int p = 16 * s + 8 * j + 2 * t;
          Bxn[ct][0] = __float_as_uint(xn_s[p * T::XS + 8 * ct + g]);
          Bxn[ct][1] = __float_as_uint(xn_s[(p + 1) * T::XS + 8 * ct + g]);
        }
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          uint32_t Ak[4], Av[4];
          load_At(scrK, mt, j, g, t, Ak);
          load_At(scrV, mt, j, g, t, Av);
#pragma unroll
          for (int ct = 0; ct < T::CT; ++ct) {
            mma8(dwk[mt][ct], Ak[0], Ak[1], Ak[2], Ak[3], Bxn[ct][0], Bxn[ct][1]);
            mma8(dwv[mt][ct], Av[0], Av[1], Av[2], Av[3], Bxn[ct][0], Bxn[ct][1]);
          }
        }
      }
      __syncwarp();
    }
    __syncthreads();
    {  // thread j = position: RMSNorm_pre backward + residual gradient
      const int j = threadIdx.x, n = n0 + j;
      const bool ok = n < n_end;
      const float inv = inv_s[j];
      float uh[C], duh[C];
      float dot = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const size_t idx = ((size_t)r * C + c) * a.L + n;
        const float dxn = ok ? yp_s[j * YS + c] + __ldg(a.dxnq + idx) : 0.f;
        const float xv = ok ? __ldg(a.x + idx) : 0.f;
        uh[c] = xv * inv;
        const float dgc = warp_sum(dxn * uh[c] * sqrtC);
        if (lane == 0) atomicAdd(acc_s + c, dgc);
        duh[c] = dxn * a.g_pre[c] * sqrtC;
        dot = fmaf(duh[c], uh[c], dot);
      }
      if (ok) {
        const bool big = inv < 1e12f;
#pragma unroll
        for (int c = 0; c < C; ++c) {
          const size_t idx = ((size_t)r * C + c) * a.L + n;
          const float d = big ? (duh[c] - uh[c] * dot) * inv : duh[c] * inv;
          a.dx[idx] = __ldg(a.dres + idx) + d;
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int ct = 0; ct < T::CT; ++ct)
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int d = h * 32 + 16 * mt + g + 8 * (i >> 1), c = 8 * ct + 2 * t + (i & 1);
        if (c < C) {
          atomicAdd(a.dwqkv + (size_t)(kHD + d) * C + c, dwk[mt][ct][i]);
          atomicAdd(a.dwqkv + (size_t)(2 * kHD + d) * C + c, dwv[mt][ct][i]);
        }
      }
  __syncthreads();
  if (threadIdx.x < C) atomicAdd(a.dg_pre + threadIdx.x, acc_s[threadIdx.x]);
}

template <int C>
static int bwd_tc(const LAArgs& a, cudaStream_t st) {
  using T = TC<C>;
  constexpr int YS = T::CP + 4;
  dim3 grid((unsigned)a.nchunk, (unsigned)a.R);
  {
    size_t smem = sizeof(float) * (2 * SP * T::XS + SP * YS + 4 * 3 * 16 * RS + 4 * 32 * T::CP + 2 * C);
    cudaFuncSetAttribute(la_bwd_q_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    la_bwd_q_tc_kernel<C><<<grid, 128, smem, st>>>(a);
    DQ_LAUNCH_CHECK();
  }
  la_bwd_combine_launch(a, st);
  DQ_LAUNCH_CHECK();
  {
    size_t smem = sizeof(float) * (SP * T::XS + SP * YS + 4 * 2 * 16 * RS + SP + C);
    cudaFuncSetAttribute(la_bwd_kv_tc_kernel<C>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    la_bwd_kv_tc_kernel<C><<<grid, 128, smem, st>>>(a);
    DQ_LAUNCH_CHECK();
  }
  return 0;
}

int la_bwd_tc_dispatch(const LAArgs& a, int C, cudaStream_t st) {
  switch (C) {
    case 4: return bwd_tc<4>(a, st);
    case 8: return bwd_tc<8>(a, st);
    case 12: return bwd_tc<12>(a, st);
    case 16: return bwd_tc<16>(a, st);
    case 24: return bwd_tc<24>(a, st);
    case 32: return bwd_tc<32>(a, st);
    default: return -3;
  }
}

}  // namespace dq
