// eventgrad_b200 -- device helpers shared by the fused BatchNorm kernels (csrc/bn_act.cu, csrc/bn_act_v2.cu).
// Geometry: NHWC == row-major [M, C]; a CTA owns one 64-channel slice (blockIdx.x) and one row split
// (blockIdx.y); 8 threads x 8 channels (one 128-byte line) per row, 32 rows per pass.
#pragma once
#include "api.h"
#include "common.cuh"

namespace egb {

#define BN_THREADS 256
#define BN_RPP 32            // rows per pass (256 threads / 8 threads per row)
#define BN_SLICE 64          // channels per slice
#define BN_FWD_PASSES 8      // slab depth held in registers by the fused forward
#define BN_BWD_PASSES 4      // ... by the fused backward (3 tensors)

struct V8 {
  float v[8];
};
__device__ __forceinline__ V8 unpack_bf16x8(const uint4& u) {
  V8 r;
  const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float2 f = __bfloat1622float2(h[i]);
    r.v[2 * i] = f.x;
    r.v[2 * i + 1] = f.y;
  }
  return r;
}
__device__ __forceinline__ uint4 pack_bf16x8(const V8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  return u;
}
__device__ __forceinline__ uint4 ldg16(const __nv_bfloat16* p) { return *reinterpret_cast<const uint4*>(p); }
__device__ __forceinline__ void stg16(__nv_bfloat16* p, const uint4& u) { *reinterpret_cast<uint4*>(p) = u; }

// -------------------------------------------------------------------------------------------
// Reduce the per-thread 2x8 accumulators over the 32 row lanes, write this CTA's partial row
// [128] = {sum a[64] | sum b[64]} for its slice.
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_partials(const float (&a)[8], const float (&b)[8], float* smem /*[32][128]*/,
                                               int tx, int ty, float* partial_row) {
  float* row = smem + ty * 128;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    row[tx * 8 + e] = a[e];
    row[64 + tx * 8 + e] = b[e];
  }
  __syncthreads();
  if (threadIdx.x < 128) {
    float s = 0.f;
#pragma unroll 8
    for (int r = 0; r < BN_RPP; ++r) s += smem[r * 128 + threadIdx.x];
    partial_row[threadIdx.x] = s;
  }
}

// Last CTA of a slice: tot[c] = sum_r partial[r][c], fixed order, double.  256 threads = 32 float4
// columns x 8 row lanes, 8 loads in flight each; lanes folded through smem (8 x 128 doubles).
__device__ __forceinline__ void slice_combine(const float* partial, int nrows, double* sm /*[8][128]*/) {
  const int col = threadIdx.x & 31, bl = threadIdx.x >> 5;
  const float4* p4 = reinterpret_cast<const float4*>(partial) + col;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  int r = bl;
  for (; r + 56 < nrows; r += 64) {
    float4 x[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) x[u] = __ldcg(p4 + (size_t)(r + 8 * u) * 32);
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      a0 += (double)x[u].x; a1 += (double)x[u].y; a2 += (double)x[u].z; a3 += (double)x[u].w;
    }
  }
  for (; r < nrows; r += 8) {
    const float4 x0 = __ldcg(p4 + (size_t)r * 32);
    a0 += (double)x0.x; a1 += (double)x0.y; a2 += (double)x0.z; a3 += (double)x0.w;
  }
  __syncthreads();                                        // smem is being re-purposed
  double* dst = sm + bl * 128 + col * 4;
  dst[0] = a0; dst[1] = a1; dst[2] = a2; dst[3] = a3;
  __syncthreads();
  if (threadIdx.x < 128) {
    double t = sm[threadIdx.x];
#pragma unroll
    for (int l = 1; l < 8; ++l) t += sm[l * 128 + threadIdx.x];
    sm[threadIdx.x] = t;                                  // row 0 = totals
  }
  __syncthreads();
}

__device__ __forceinline__ bool elect_last_of_slice(unsigned int* ticket, unsigned int n) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(ticket, 1u);
    s_last = (prev == n - 1) ? 1 : 0;
    if (s_last) *ticket = 0u;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// mean / invstd / running stats for one slice from its totals (tot[0..63] = sum x, [64..127] = sum x^2)
__device__ __forceinline__ void finalize_stats(const BnParams& p, int slice, const double* tot) {
  if (threadIdx.x < BN_SLICE) {
    const int c = slice * BN_SLICE + threadIdx.x;
    const double invM = 1.0 / (double)p.M;
    const double mean = tot[threadIdx.x] * invM;
    double var = tot[64 + threadIdx.x] * invM - mean * mean;   // biased
    if (var < 0.0) var = 0.0;
    p.mean[c] = (float)mean;
    p.invstd[c] = rsqrtf((float)var + p.eps);
    if (p.run_mean != nullptr) {
      const double unb = p.M > 1 ? var * (double)p.M / (double)(p.M - 1) : var;
      p.run_mean[c] = (float)((1.0 - p.momentum) * (double)p.run_mean[c] + p.momentum * mean);
      p.run_var[c] = (float)((1.0 - p.momentum) * (double)p.run_var[c] + p.momentum * unb);
    }
  }
  if (threadIdx.x == 0 && slice == 0 && p.nbt != nullptr) *p.nbt += 1;
}

// epoch flag: slice-last CTA publishes, everyone else of the slice spins (bounded)
__device__ __forceinline__ void publish_flag(unsigned int* flag, unsigned int value) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(value) : "memory");
  }
}
__device__ __forceinline__ void wait_flag(const unsigned int* flag, unsigned int value, int* status) {
  if (threadIdx.x == 0) {
    const uint64_t t0 = globaltimer_ns();
    unsigned v;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
      if ((int)(v - value) >= 0) break;
      if (globaltimer_ns() - t0 > 2000000000ull) {       // 2 s: co-residency assumption violated
        if (status != nullptr) atomicExch(status, 2);
        break;
      }
    } while (true);
  }
  __syncthreads();
}

}  // namespace egb
