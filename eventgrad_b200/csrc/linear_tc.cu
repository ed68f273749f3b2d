// eventgrad_b200 -- fused Linear + bias (+ ReLU) on the 5th-generation tensor cores (sm_100a).
//
//   Y[M,N] = act( X[M,K] * W[N,K]^T + b[N] )      X, W bf16 (K-major), fp32 accumulate in TMEM
//
// This is the GEMM-shaped hot op of the reference's MNIST programs: the MLP of dmnist/cent and
// dmnist/decent runs Linear(784,128)+ReLU on the WHOLE per-rank shard every step
// (/root/reference/dmnist/cent/cent.cpp:16-35, :62-65: 60000/R x 784 x 128).
//
// Structure (one CTA = one 128-row tile of X, 128 threads):
//   * operands are staged in shared memory in the canonical K-major no-swizzle UMMA layout
//     [K/8 chunk][row][8 elements]: a core matrix = 8 rows x 16 bytes, contiguous (128 B);
//     SBO (next 8-row group) = 128 B, LBO (next 8-element K chunk) = rows*16 B;
//   * one elected thread issues `tcgen05.mma.cta_group::1.kind::f16` (M=128, N<=256, K=16 per
//     instruction) with 64-bit shared-memory descriptors and a 32-bit instruction descriptor;
//     completion is tracked with `tcgen05.commit` -> mbarrier;
//   * the fp32 accumulator lives in TMEM (128 lanes x N columns, allocated with tcgen05.alloc);
//     the epilogue reads it back with `tcgen05.ld.32x32b.x32`, adds the bias, applies ReLU and
//     stores bf16 or fp32.
// SASS evidence: UTCHMMA (tcgen05.mma), LDTM (tcgen05.ld), UTCBAR (commit) -- profiles/sass_evidence.md.
#include "api.h"
#include "common.cuh"

namespace egb {

#define LT_TM 128          // rows per CTA == UMMA M
#define LT_BK 64           // K elements staged per round (4 UMMA k-steps of 16)
#define LT_THREADS 128

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);               // start address, 16-byte units   [0,14)
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;      // leading-dimension byte offset   [16,30)
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;      // stride-dimension byte offset    [32,46)
  d |= 1ull << 46;                                        // descriptor version (Blackwell)  [46,48)
  // base_offset = 0, lbo_mode = 0, layout_type = SWIZZLE_NONE (0) in [61,64)
  return d;
}

// kind::f16 instruction descriptor: D=F32, A=B=BF16, both K-major, M=128, N=n
__device__ __forceinline__ uint32_t make_instr_desc(int n) {
  return (1u << 4)                 // c_format  F32
       | (1u << 7)                 // a_format  BF16
       | (1u << 10)                // b_format  BF16
       | ((uint32_t)(n >> 3) << 17)   // n_dim
       | ((uint32_t)(LT_TM >> 4) << 24);  // m_dim
}

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}

template <bool kOutBf16>
__global__ void __launch_bounds__(LT_THREADS) linear_tc_kernel(const LinearParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, K = p.K;
  const int row0 = blockIdx.x * LT_TM;
  // smem carve-up: A tile [8 chunks][128 rows][16 B] = 16 KB | B tile [8][N][16 B] | mbarrier | tmem ptr
  unsigned char* sA = smem_raw;
  unsigned char* sB = smem_raw + (LT_BK / 8) * LT_TM * 16;
  uint64_t* bar = reinterpret_cast<uint64_t*>(sB + (LT_BK / 8) * N * 16);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 1);
  const uint32_t bar_a = smem_u32(bar);
  const uint32_t ncols = (N <= 32) ? 32u : (N <= 64) ? 64u : (N <= 128) ? 128u : 256u;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const uint32_t idesc = make_instr_desc(N);

  const int num_kb = (K + LT_BK - 1) / LT_BK;
  uint32_t parity = 0;
  for (int kb = 0; kb < num_kb; ++kb) {
    const int k0 = kb * LT_BK;
    // ---- stage A (128 x 64) and B (N x 64): 8 consecutive threads fetch one 128-byte row segment ----
    for (int idx = tid; idx < LT_TM * (LT_BK / 8); idx += LT_THREADS) {
      const int r = idx >> 3, c = idx & 7;
      const int gr = row0 + r, gk = k0 + c * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (gr < p.M && gk < K) v = *reinterpret_cast<const uint4*>(p.x + (size_t)gr * K + gk);
      *reinterpret_cast<uint4*>(sA + ((size_t)c * LT_TM + r) * 16) = v;
    }
    for (int idx = tid; idx < N * (LT_BK / 8); idx += LT_THREADS) {
      const int r = idx >> 3, c = idx & 7;
      const int gk = k0 + c * 8;
      uint4 v = make_uint4(0, 0, 0, 0);
      if (gk < K) v = *reinterpret_cast<const uint4*>(p.w + (size_t)r * K + gk);
      *reinterpret_cast<uint4*>(sB + ((size_t)c * N + r) * 16) = v;
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> async proxy (UMMA)
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int rem = K - k0;
      const int ksteps = (rem >= LT_BK) ? (LT_BK / 16) : ((rem + 15) / 16);
      for (int s = 0; s < ksteps; ++s) {
        const uint64_t da = make_smem_desc(smem_u32(sA) + (uint32_t)s * 2u * LT_TM * 16u, LT_TM * 16u, 128u);
        const uint64_t db = make_smem_desc(smem_u32(sB) + (uint32_t)s * 2u * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
        const uint32_t accum = (kb > 0 || s > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
            : "memory");
      }
      // arrives on the mbarrier once every MMA issued so far has finished reading smem / writing TMEM
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar_a)
                   : "memory");
    }
    mbar_wait(bar_a, parity);
    parity ^= 1u;
  }
  // ---- epilogue: TMEM -> registers -> bias + ReLU -> global -------------------------------------
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const int row = row0 + warp * 32 + lane;                 // TMEM lane == tile row (M = 128)
  for (int cb = 0; cb < N; cb += 32) {
    uint32_t r[32];
    const uint32_t taddr = tmem_d + ((uint32_t)(warp * 32) << 16) + (uint32_t)cb;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    if (row < p.M) {
      const int nvalid = min(32, N - cb);
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (j < nvalid) {
          float v = __uint_as_float(r[j]) + (p.bias ? p.bias[cb + j] : 0.f);
          if (p.relu) v = fmaxf(v, 0.f);
          if (kOutBf16)
            reinterpret_cast<__nv_bfloat16*>(p.y)[(size_t)row * N + cb + j] = __float2bfloat16_rn(v);
          else
            reinterpret_cast<float*>(p.y)[(size_t)row * N + cb + j] = v;
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(ncols));
  }
}

cudaError_t launch_linear_tc(const LinearParams& p, cudaStream_t s) {
  if (p.N % 16 != 0 || p.N < 16 || p.N > 256 || p.K % 8 != 0 || p.M < 1) return cudaErrorInvalidValue;
  const size_t smem = (size_t)(LT_BK / 8) * LT_TM * 16 + (size_t)(LT_BK / 8) * p.N * 16 + 64;
  const int grid = (p.M + LT_TM - 1) / LT_TM;
  cudaError_t e;
  if (p.out_bf16) {
    e = cudaFuncSetAttribute(linear_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    linear_tc_kernel<true><<<grid, LT_THREADS, smem, s>>>(p);
  } else {
    e = cudaFuncSetAttribute(linear_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    linear_tc_kernel<false><<<grid, LT_THREADS, smem, s>>>(p);
  }
  return cudaGetLastError();
}

}  // namespace egb
