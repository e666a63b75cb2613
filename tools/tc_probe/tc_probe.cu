// Bring-up probe for the tcgen05 LinearAttention kernels (sm_100a).  Standalone binary, run on the GPU box:
//   1. correctness of the exact tcgen05.mma operand configurations linattn_tc.cu relies on, against a CPU product:
//        T1  SS kind::tf32, A K-major / B K-major, no swizzle (core matrices 8 rows x 16 B)
//        T2  TS kind::tf32, A in TMEM (lane = row, column = k), B K-major no swizzle
//        T3  SS kind::f16 (bf16), A MN-major no swizzle from a [m/8][k][8] tile, B K-major no swizzle
//        T4  write-after-read through TMEM: an MMA that reads A from TMEM columns followed, without a wait, by an MMA
//            that writes its accumulator over the same columns
//   2. throughput of the pipes that bound those kernels: MUFU.EX2, a degree-3 polynomial exp2 on the FMA pipe, a mix
//      of the two, tcgen05.ld / tcgen05.st, FFMA vs FFMA2.
// Output: one line per test; exit code 0 iff every correctness test passed.
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include <cmath>

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(2); } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(done) : "r"(bar), "r"(parity) : "memory");
  return done != 0;
}
__device__ __forceinline__ bool mbar_wait(uint32_t bar, uint32_t parity) {   // bounded: false on timeout
  long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 2000000000LL) return false;
  }
  return true;
}
// no-swizzle shared-memory matrix descriptor: LBO = byte distance between core matrices adjacent in K,
// SBO = byte distance between core matrices adjacent in M/N (both majors), version 1 (sm_100)
__device__ __forceinline__ uint64_t umma_desc_ns(uint32_t saddr, uint32_t lbo, uint32_t sbo) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  return d;
}
constexpr uint32_t kTF32 = 2, kBF16 = 1;
__host__ __device__ constexpr uint32_t make_idesc(uint32_t fmt, int M, int N, int a_mn, int b_mn) {
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)a_mn << 15) | ((uint32_t)b_mn << 16) | ((uint32_t)(N >> 3) << 17) |
         ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void mma_ss_tf32(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ts_tf32(uint32_t d, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_ss_f16(uint32_t d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t addr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void proxy_fence() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

#define LD16_REGS(r) "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), \
                     "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
               : LD16_REGS(r) : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
               ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
                 "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]) : "memory");
}

// ------------------------------------------------------------------------------------------------ correctness
// One CTA of 128 threads.  mode selects the operand configuration.  A: (128, K) row-major fp32, B: (N, K) row-major fp32,
// D: (128, N) fp32.  K is a multiple of the instruction K (8 for tf32, 16 for bf16).
struct TestArgs {
  const float* A; const float* B; float* D; int N, K, mode; int* status;
};
constexpr int TEST_SMEM = 96 * 1024;

__global__ void __launch_bounds__(128) mma_test_kernel(TestArgs a) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  const int N = a.N, K = a.K;
  uint8_t* sA = smem;                 // up to 48 KB
  uint8_t* sB = smem + 48 * 1024;     // up to 48 KB
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(smem_u32(&tslot), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
  const uint32_t D_COL = 0, A_COL = 256;

  if (a.mode == 1 || a.mode == 2) {
    // K-major tf32 tiles: element (r, k): (r % 8) * 16 + (r / 8) * SBO + (k / 4) * LBO + (k % 4) * 4, LBO = 128, SBO = 128 * (K / 4)
    const uint32_t lbo = 128, sbo = 128 * (K / 4);
    if (a.mode == 1)
      for (int i = tid; i < 128 * K; i += 128) {
        const int r = i / K, k = i % K;
        *reinterpret_cast<float*>(sA + (r % 8) * 16 + (r / 8) * sbo + (k / 4) * lbo + (k % 4) * 4) = a.A[i];
      }
    for (int i = tid; i < N * K; i += 128) {
      const int r = i / K, k = i % K;
      *reinterpret_cast<float*>(sB + (r % 8) * 16 + (r / 8) * sbo + (k / 4) * lbo + (k % 4) * 4) = a.B[i];
    }
    if (a.mode == 2) {  // A into TMEM: thread = row, column = k
      for (int k0 = 0; k0 < K; k0 += 16) {
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(a.A[tid * K + k0 + j]);
        tmem_st16(lane_base + A_COL + k0, v);
      }
      tmem_st_wait();
    }
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = make_idesc(kTF32, 128, N, 0, 0);
      for (int k0 = 0; k0 < K; k0 += 8) {
        const uint64_t db = umma_desc_ns(smem_u32(sB) + (k0 / 4) * lbo, lbo, sbo);
        if (a.mode == 1) mma_ss_tf32(tm + D_COL, umma_desc_ns(smem_u32(sA) + (k0 / 4) * lbo, lbo, sbo), db, idesc, k0 > 0);
        else mma_ts_tf32(tm + D_COL, tm + A_COL + k0, db, idesc, k0 > 0);
      }
      umma_commit(smem_u32(&bar));
    }
  } else if (a.mode == 3) {
    // A (M = 128, K) MN-major bf16 from a [m/8][k][8] tile: element (m, k): (m % 8) * 2 + k * 16 + (m / 8) * (K * 16)
    //   -> core matrix = 8 k-rows x 16 B; K-direction stride (LBO) = 128, M-direction stride (SBO) = K * 16
    // B (N, K) K-major bf16: element (n, k): (n % 8) * 16 + (n / 8) * SBO + (k / 8) * LBO + (k % 8) * 2, LBO = 128, SBO = 128 * (K / 8)
    for (int i = tid; i < 128 * K; i += 128) {
      const int m = i / K, k = i % K;
      *reinterpret_cast<__nv_bfloat16*>(sA + (m % 8) * 2 + k * 16 + (m / 8) * (K * 16)) = __float2bfloat16(a.A[i]);
    }
    const uint32_t lbo = 128, sbo = 128 * (K / 8);
    for (int i = tid; i < N * K; i += 128) {
      const int r = i / K, k = i % K;
      *reinterpret_cast<__nv_bfloat16*>(sB + (r % 8) * 16 + (r / 8) * sbo + (k / 8) * lbo + (k % 8) * 2) = __float2bfloat16(a.B[i]);
    }
    proxy_fence();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t idesc = make_idesc(kBF16, 128, N, 1, 0);
      for (int k0 = 0; k0 < K; k0 += 16) {
        const uint64_t da = umma_desc_ns(smem_u32(sA) + k0 * 16, 128, K * 16);
        const uint64_t db = umma_desc_ns(smem_u32(sB) + (k0 / 8) * lbo, lbo, sbo);
        mma_ss_f16(tm + D_COL, da, db, idesc, k0 > 0);
      }
      umma_commit(smem_u32(&bar));
    }
  }
  const bool ok = mbar_wait(smem_u32(&bar), 0);
  tc_fence_after();
  if (!ok && tid == 0) *a.status = 1;
  if (ok) {
    for (int n0 = 0; n0 < N; n0 += 16) {
      uint32_t v[16];
      tmem_ld16_nowait(lane_base + D_COL + n0, v);
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 16; ++j) a.D[tid * N + n0 + j] = __uint_as_float(v[j]);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// T4: P (128 x 64 fp32 in TMEM columns [0, 64)); D2 (cols [256, 272)) = P . B2^T (TS, K = 64, N = 16) issued, then IMMEDIATELY
// S (cols [0, 64)) = A1 . B1^T (SS tf32, K = 8, N = 64) over the columns MMA2 reads.  Repeated `iters` times per CTA with
// different P; D2 is checked on the device against the fp32 product of the tf32-exact inputs.
__global__ void __launch_bounds__(128) war_test_kernel(int iters, int* mismatches, int* status) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  uint8_t* sA1 = smem;            // 128 x 8 tf32 K-major: 4 KB
  uint8_t* sB1 = smem + 4096;     // 64 x 8: 2 KB
  uint8_t* sB2 = smem + 8192;     // 16 x 64 tf32 K-major: LBO 128, SBO 128 * 16 = 2048: 4 KB
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(smem_u32(&tslot), 512);
  for (int i = tid; i < 128 * 8; i += 128) {
    const int r = i / 8, k = i % 8;
    *reinterpret_cast<float*>(sA1 + (r % 8) * 16 + (r / 8) * 256 + (k / 4) * 128 + (k % 4) * 4) = (float)((r * 7 + k * 3) % 11 - 5);
  }
  for (int i = tid; i < 64 * 8; i += 128) {
    const int r = i / 8, k = i % 8;
    *reinterpret_cast<float*>(sB1 + (r % 8) * 16 + (r / 8) * 256 + (k / 4) * 128 + (k % 4) * 4) = (float)((r * 5 + k) % 7 - 3);
  }
  for (int i = tid; i < 16 * 64; i += 128) {
    const int r = i / 64, k = i % 64;
    *reinterpret_cast<float*>(sB2 + (r % 8) * 16 + (r / 8) * 2048 + (k / 4) * 128 + (k % 4) * 4) = (float)((r + k * 3) % 5 - 2);
  }
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
  int bad = 0;
  uint32_t ph = 0;
  for (int it = 0; it < iters; ++it) {
    // P[row][k] = small integers depending on (row, k, it, block)
    float pv[64];
#pragma unroll
    for (int k0 = 0; k0 < 64; k0 += 16) {
      uint32_t v[16];
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        pv[k0 + j] = (float)((tid * 3 + (k0 + j) * 5 + it * 7 + blockIdx.x) % 13 - 6);
        v[j] = __float_as_uint(pv[k0 + j]);
      }
      tmem_st16(lane_base + k0, v);
    }
    tmem_st_wait();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      const uint32_t id2 = make_idesc(kTF32, 128, 16, 0, 0), id1 = make_idesc(kTF32, 128, 64, 0, 0);
      for (int k0 = 0; k0 < 64; k0 += 8)
        mma_ts_tf32(tm + 256, tm + k0, umma_desc_ns(smem_u32(sB2) + (k0 / 4) * 128, 128, 2048), id2, k0 > 0);
      mma_ss_tf32(tm + 0, umma_desc_ns(smem_u32(sA1), 128, 256), umma_desc_ns(smem_u32(sB1), 128, 256), id1, 0);
      umma_commit(smem_u32(&bar));
    }
    const bool ok = mbar_wait(smem_u32(&bar), ph);
    ph ^= 1;
    tc_fence_after();
    if (!ok) { if (tid == 0) *status = 1; break; }
    uint32_t d2[16];
    tmem_ld16_nowait(lane_base + 256, d2);
    tmem_ld_wait();
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      float ref = 0.f;
      for (int k = 0; k < 64; ++k) ref += pv[k] * (float)((n + k * 3) % 5 - 2);
      if (ref != __uint_as_float(d2[n])) ++bad;
    }
    // also check S (first 16 columns)
    uint32_t s[16];
    tmem_ld16_nowait(lane_base + 0, s);
    tmem_ld_wait();
#pragma unroll
    for (int n = 0; n < 16; ++n) {
      float ref = 0.f;
      for (int k = 0; k < 8; ++k) ref += (float)((tid * 7 + k * 3) % 11 - 5) * (float)((n * 5 + k) % 7 - 3);
      if (ref != __uint_as_float(s[n])) ++bad;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
  }
  if (bad) atomicAdd(mismatches, bad);
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ------------------------------------------------------------------------------------------------ throughput
__device__ __forceinline__ float fexp2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ unsigned long long pk2(float a, float b) { unsigned long long r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk2(unsigned long long v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ unsigned long long fma2(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ unsigned long long add2(unsigned long long a, unsigned long long b) {
  unsigned long long r; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
// 2^x for two values, x <= 0 assumed >= -126: round-to-nearest split, degree-3 polynomial (max rel 7.5e-5), exponent by IMAD
__device__ __forceinline__ void exp2_poly2(float x0, float x1, float& y0, float& y1) {
  const float kMagic = 12582912.f;
  const unsigned long long xm = pk2(kMagic, kMagic), nm = pk2(-kMagic, -kMagic);
  const unsigned long long x = pk2(x0, x1);
  const unsigned long long t = add2(x, xm);
  const unsigned long long n = add2(t, nm);
  float n0, n1; upk2(n, n0, n1);
  const unsigned long long f = add2(x, pk2(-n0, -n1));
  unsigned long long p = fma2(f, pk2(0.05517167f, 0.05517167f), pk2(0.24261113f, 0.24261113f));
  p = fma2(p, f, pk2(0.69326097f, 0.69326097f));
  p = fma2(p, f, pk2(0.99992806f, 0.99992806f));
  float p0, p1, t0, t1; upk2(p, p0, p1); upk2(t, t0, t1);
  y0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  y1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

// mode 0: all MUFU; 1: all polynomial; 2: 5 of 8 MUFU + 3 of 8 poly (pairs: 2 poly pairs... see below); 3: 4/8 each
__global__ void __launch_bounds__(128) exp_bench_kernel(float* out, int iters, int mode, float seed) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = -seed * (float)(i + 1) - 0.001f * threadIdx.x;
  float acc = 0.f;
  for (int it = 0; it < iters; ++it) {
    float e[32];
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) e[i] = fexp2(v[i]);
    } else if (mode == 1) {
#pragma unroll
      for (int i = 0; i < 32; i += 2) exp2_poly2(v[i], v[i + 1], e[i], e[i + 1]);
    } else if (mode == 2) {   // per 8: 6 MUFU, 2 poly
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
#pragma unroll
        for (int j = 0; j < 6; ++j) e[i + j] = fexp2(v[i + j]);
        exp2_poly2(v[i + 6], v[i + 7], e[i + 6], e[i + 7]);
      }
    } else {                  // per 8: 4 MUFU, 4 poly
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
#pragma unroll
        for (int j = 0; j < 4; ++j) e[i + j] = fexp2(v[i + j]);
        exp2_poly2(v[i + 4], v[i + 5], e[i + 4], e[i + 5]);
        exp2_poly2(v[i + 6], v[i + 7], e[i + 6], e[i + 7]);
      }
    }
    // cheap dependent use (one packed add per pair) so nothing is dead-code eliminated, and feed back into the inputs
    unsigned long long s = pk2(0.f, 0.f);
#pragma unroll
    for (int i = 0; i < 32; i += 2) s = add2(s, pk2(e[i], e[i + 1]));
    float s0, s1; upk2(s, s0, s1);
    acc += s0 + s1;
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] -= 1e-6f;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(128) fma_bench_kernel(float* out, int iters, int mode) {
  float v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = 0.001f * (i + threadIdx.x);
  const float a = 1.0001f, b = 1e-7f;
  const unsigned long long a2 = pk2(a, a), b2 = pk2(b, b);
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = fmaf(v[i], a, b);
    } else {
#pragma unroll
      for (int i = 0; i < 32; i += 2) {
        unsigned long long r = fma2(pk2(v[i], v[i + 1]), a2, b2);
        upk2(r, v[i], v[i + 1]);
      }
    }
  }
  float acc = 0.f;
#pragma unroll
  for (int i = 0; i < 32; ++i) acc += v[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

// tcgen05.ld / st throughput: each CTA (128 threads) allocates 128 columns; loop: ld x16 over 64 columns (4 loads), one wait
__global__ void __launch_bounds__(128) tmem_bench_kernel(float* out, int iters, int mode) {
  __shared__ uint32_t tslot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(smem_u32(&tslot), 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  const uint32_t lane_base = tm + ((uint32_t)(warp * 32) << 16);
  uint32_t acc = 0;
  uint32_t v[4][16];
#pragma unroll
  for (int c = 0; c < 4; ++c)
#pragma unroll
    for (int j = 0; j < 16; ++j) v[c][j] = threadIdx.x + j + c;
#pragma unroll
  for (int c = 0; c < 4; ++c) tmem_st16(lane_base + 16 * c, v[c]);
  tmem_st_wait();
  for (int it = 0; it < iters; ++it) {
    if (mode == 0) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld16_nowait(lane_base + 16 * c, v[c]);
      tmem_ld_wait();
#pragma unroll
      for (int c = 0; c < 4; ++c) acc ^= v[c][0] ^ v[c][15];
    } else {
#pragma unroll
      for (int c = 0; c < 4; ++c) { v[c][0] += it; tmem_st16(lane_base + 16 * c, v[c]); }
      tmem_st_wait();
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = (float)acc;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 128);
}


// ------------------------------------------------------------------------------------------------ MMA issue rate
// One CTA per SM; thread 0 issues `n` identical MMAs back to back, commits, waits.  mode 0: SS tf32 K-major N = nn, K = 8;
// 1: TS tf32 (A in TMEM) N = nn, K = 8;  2: SS bf16 A MN-major N = nn, K = 16.  Reports cycles per MMA (clock64 on the SM).
__global__ void __launch_bounds__(128) mma_rate_kernel(int n, int mode, int nn, int nacc, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tslot;
  const int tid = threadIdx.x, warp = tid >> 5;
  for (int i = tid; i < 48 * 1024 / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;   // finite in every format
  if (tid == 0) { mbar_init(smem_u32(&bar), 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  if (warp == 0) tmem_alloc(smem_u32(&tslot), 512);
  proxy_fence();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = tslot;
  if (tid == 0) {
    const uint32_t sa = smem_u32(smem), sb = sa + 32 * 1024;
    const uint64_t da = umma_desc_ns(sa, 128, mode == 2 ? 2048 : 256), db = umma_desc_ns(sb, 128, mode == 2 ? 2048 : 256);
    const uint32_t idesc = mode == 2 ? make_idesc(kBF16, 128, nn, 1, 0) : make_idesc(kTF32, 128, nn, 0, 0);
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
      const uint32_t d = tm + (uint32_t)((i % nacc) * 32);   // nacc independent accumulators (N <= 32 when nacc > 1)
      if (mode == 0) mma_ss_tf32(d, da, db, idesc, 1);
      else if (mode == 1) mma_ts_tf32(d, tm + 256, db, idesc, 1);
      else mma_ss_f16(d, da, db, idesc, 1);
    }
    umma_commit(smem_u32(&bar));
    mbar_wait(smem_u32(&bar), 0);
    const long long t1 = clock64();
    if (blockIdx.x == 0) *cycles = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tm, 512);
}

// ------------------------------------------------------------------------------------------------ host
static float rnd_small(uint32_t& s, int range) { s = s * 1664525u + 1013904223u; return (float)((int)((s >> 16) % (2 * range + 1)) - range); }

static bool run_mma_test(const char* name, int mode, int N, int K) {
  std::vector<float> A(128 * K), B(N * K), D(128 * N, -12345.f), R(128 * N, 0.f);
  uint32_t s = 1234u + mode * 77 + N + K;
  for (auto& v : A) v = rnd_small(s, 4);
  for (auto& v : B) v = rnd_small(s, 3);
  for (int m = 0; m < 128; ++m)
    for (int n = 0; n < N; ++n) {
      float acc = 0.f;
      for (int k = 0; k < K; ++k) acc += A[m * K + k] * B[n * K + k];
      R[m * N + n] = acc;
    }
  float *dA, *dB, *dD; int* dS;
  CK(cudaMalloc(&dA, A.size() * 4)); CK(cudaMalloc(&dB, B.size() * 4)); CK(cudaMalloc(&dD, D.size() * 4)); CK(cudaMalloc(&dS, 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dD, D.data(), D.size() * 4, cudaMemcpyHostToDevice));
  CK(cudaMemset(dS, 0, 4));
  CK(cudaFuncSetAttribute(mma_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TEST_SMEM));
  TestArgs a{dA, dB, dD, N, K, mode, dS};
  mma_test_kernel<<<1, 128, TEST_SMEM>>>(a);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: CUDA error %s\n", name, cudaGetErrorString(e)); exit(3); }
  int st = 0;
  CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost));
  int bad = 0, first = -1;
  for (size_t i = 0; i < D.size(); ++i) if (D[i] != R[i]) { if (first < 0) first = (int)i; ++bad; }
  printf("%s (N=%d K=%d): %s  mismatches %d / %zu  timeout %d", name, N, K, (bad == 0 && st == 0) ? "PASS" : "FAIL", bad, D.size(), st);
  if (first >= 0) printf("  first at (m=%d, n=%d): got %g want %g", first / N, first % N, D[first], R[first]);
  printf("\n");
  cudaFree(dA); cudaFree(dB); cudaFree(dD); cudaFree(dS);
  return bad == 0 && st == 0;
}

template <class F>
static float time_ms(F f, int reps = 5) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f();
  CK(cudaDeviceSynchronize());
  float best = 1e30f;
  for (int i = 0; i < reps; ++i) {
    cudaEventRecord(e0);
    f();
    cudaEventRecord(e1);
    CK(cudaEventSynchronize(e1));
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    if (ms < best) best = ms;
  }
  return best;
}

int main(int argc, char** argv) {
  int dev = 0; CK(cudaSetDevice(dev));
  cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, dev));
  int sms = prop.multiProcessorCount;
  int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, dev);
  printf("device %s, %d SMs, clock %d MHz\n", prop.name, sms, clk_khz / 1000);
  bool ok = true;
  ok &= run_mma_test("T1 SS tf32 K/K noswz", 1, 64, 8);
  ok &= run_mma_test("T1 SS tf32 K/K noswz", 1, 128, 16);
  ok &= run_mma_test("T1 SS tf32 K/K noswz", 1, 16, 24);
  ok &= run_mma_test("T2 TS tf32 A=TMEM", 2, 16, 64);
  ok &= run_mma_test("T2 TS tf32 A=TMEM", 2, 64, 32);
  ok &= run_mma_test("T3 SS bf16 A MN-major", 3, 16, 128);
  ok &= run_mma_test("T3 SS bf16 A MN-major", 3, 32, 64);
  {
    int *dM, *dS; CK(cudaMalloc(&dM, 4)); CK(cudaMalloc(&dS, 4)); CK(cudaMemset(dM, 0, 4)); CK(cudaMemset(dS, 0, 4));
    CK(cudaFuncSetAttribute(war_test_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * 1024));
    war_test_kernel<<<sms, 128, 16 * 1024>>>(2000, dM, dS);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("T4: CUDA error %s\n", cudaGetErrorString(e)); exit(3); }
    int m = 0, st = 0; CK(cudaMemcpy(&m, dM, 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(&st, dS, 4, cudaMemcpyDeviceToHost));
    printf("T4 TMEM write-after-read without wait: %s  mismatches %d timeout %d (informational)\n", (m == 0 && st == 0) ? "SAFE" : "UNSAFE", m, st);
  }
  float* dout; CK(cudaMalloc(&dout, (size_t)sms * 16 * 128 * 4));
  const double ghz = clk_khz * 1e-6;
  const char* exp_names[4] = {"MUFU.EX2 only", "poly only", "6 MUFU + 2 poly", "4 MUFU + 4 poly"};
  for (int occ : {4, 8}) {
    for (int mode = 0; mode < 4; ++mode) {
      const int iters = 4000;
      float ms = time_ms([&] { exp_bench_kernel<<<sms * occ, 128>>>(dout, iters, mode, 0.37f); });
      double elems = (double)sms * occ * 128 * 32 * iters;
      printf("exp  %-16s %2d warps/SM: %.3f ms  %.1f elem/clk/SM (at %d MHz)\n", exp_names[mode], occ * 4, ms, elems / (ms * 1e-3) / (ghz * 1e9) / sms, clk_khz / 1000);
    }
  }
  for (int mode = 0; mode < 2; ++mode) {
    const int iters = 20000, occ = 8;
    float ms = time_ms([&] { fma_bench_kernel<<<sms * occ, 128>>>(dout, iters, mode); });
    double fmas = (double)sms * occ * 128 * 32 * iters;
    printf("fma  %-6s %2d warps/SM: %.3f ms  %.1f FMA/clk/SM\n", mode ? "FFMA2" : "FFMA", occ * 4, ms, fmas / (ms * 1e-3) / (ghz * 1e9) / sms);
  }
  for (int occ : {1, 2, 4}) {
    for (int mode = 0; mode < 2; ++mode) {
      const int iters = 20000;
      float ms = time_ms([&] { tmem_bench_kernel<<<sms * occ, 128>>>(dout, iters, mode); });
      double bytes = (double)sms * occ * 128 * 64 * 4 * iters;
      printf("tmem %s %2d warps/SM: %.3f ms  %.1f B/clk/SM\n", mode ? "st" : "ld", occ * 4, ms, bytes / (ms * 1e-3) / (ghz * 1e9) / sms);
    }
  }
  {
    long long* dc; CK(cudaMalloc(&dc, 8));
    CK(cudaFuncSetAttribute(mma_rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
    const char* nm[3] = {"SS tf32 K=8", "TS tf32 K=8", "SS bf16 MN-major K=16"};
    for (int mode = 0; mode < 3; ++mode)
      for (int nacc : {1, 2, 4, 8})
        for (int nn : {16, 32, 128}) {
          if (nacc > 1 && nn > 32) continue;
          const int n = 2000;
          mma_rate_kernel<<<sms, 128, 64 * 1024>>>(n, mode, nn, nacc, dc);
          CK(cudaDeviceSynchronize());
          long long c = 0; CK(cudaMemcpy(&c, dc, 8, cudaMemcpyDeviceToHost));
          printf("mma rate %-22s N=%3d, %d accumulators: %.1f cycles / MMA\n", nm[mode], nn, nacc, (double)c / n);
        }
  }
  printf(ok ? "ALL CORRECTNESS TESTS PASSED\n" : "SOME CORRECTNESS TESTS FAILED\n");
  return ok ? 0 : 1;
}
