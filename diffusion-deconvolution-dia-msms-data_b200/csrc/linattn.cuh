// Shared declarations of the fused LinearAttention kernels (linattn.cu: fp32 CUDA-core reference kernels and the
// combine kernels; linattn_tc.cu: TF32 tensor-core kernels).
#pragma once
#include "common.cuh"

namespace dq {

constexpr int TP = 32;     // positions per tile
constexpr int LDS_ = 132;  // padded row stride of the [n][128] staging tiles (floats)

struct LAArgs {
  const float* x;      // (R, C, L) block input
  const float* g_pre;  // (C) PreNorm gain
  const float* wqkv;   // (384, C)
  const float* wout;   // (C, 128)
  const float* bout;   // (C)
  const float* g_out;  // (C)
  float* part;         // (R, nchunk, 128, 34) forward partials [m, s, ctx[32]]
  float* ctx;          // (R, 128, 32)  ctx[h*32+d][e]
  float* ms;           // (R, 128, 2)   max and sum of exp of k over L
  float* ypre;         // (R, C, L) to_out output before RMSNorm (saved for backward; may be null)
  float* out;          // (R, C, L)
  // backward
  const float* dres;   // (R, C, L) gradient of the block output
  float* dxnq;         // (R, C, L) scratch: q-path gradient w.r.t. the pre-normed input
  float* dpart;        // (R, nchunk, 128, 32) partial d ctx
  float* dctx;         // (R, 128, 32)
  float* sd;           // (R, 128)   sum_e dctx*ctx
  float* dx;           // (R, C, L)
  float* dwqkv;        // (384, C) accumulated
  float* dwout;        // (C, 128) accumulated
  float* dbout;        // (C) accumulated
  float* dg_out;       // (C) accumulated
  float* dg_pre;       // (C) accumulated
  int R, L, chunk, nchunk;
};


// defined in linattn.cu
void la_combine_launch(const LAArgs& a, cudaStream_t st);
void la_bwd_combine_launch(const LAArgs& a, cudaStream_t st);
// defined in linattn_tc.cu (tensor-core kernels)
int la_fwd_tc_dispatch(const LAArgs& a, int C, cudaStream_t st);
int la_bwd_tc_dispatch(const LAArgs& a, int C, cudaStream_t st);

}  // namespace dq
