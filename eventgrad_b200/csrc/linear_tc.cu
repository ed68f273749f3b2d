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
//   * warp-specialised: warps 1-3 are producers that keep a 3-stage cp.async (LDGSTS) ring full, each
//     thread arriving on the stage's FULL mbarrier via cp.async.mbarrier.arrive; warp 0 / lane 0 is the
//     MMA issuer and returns slots through EMPTY mbarriers with tcgen05.commit -- no __syncthreads in
//     the main loop, MMAs issue back to back as stages fill;
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
#define LT_BK 32           // K elements per pipeline stage (2 UMMA k-steps of 16)
#define LT_STAGES 3        // cp.async ring depth: 3 x 16 KB (N=128) = 48 KB -> 4 CTAs/SM, one wave for M=60000
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

// 16-byte async copy global -> shared (LDGSTS); src_bytes = 0 zero-fills (rows >= M, K tail)
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}

template <bool kOutBf16>
__global__ void __launch_bounds__(LT_THREADS) linear_tc_kernel(const LinearParams p) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int N = p.N, K = p.K;
  const int row0 = blockIdx.x * LT_TM;
  // smem: LT_STAGES x { A tile [BK/8 chunks][128 rows][16 B] | B tile [BK/8][N][16 B] } | mbarriers | tmem ptr
  constexpr int CH = LT_BK / 8;
  constexpr int NPROD = LT_THREADS - 32;                       // warps 1..3 produce, warp 0 issues MMAs
  const uint32_t a_bytes = CH * LT_TM * 16, b_bytes = (uint32_t)CH * N * 16;
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + LT_STAGES * stage_bytes);
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * LT_STAGES + 1);
  const uint32_t smem0 = smem_u32(smem_raw);
  const uint32_t full0 = smem_u32(bars), empty0 = full0 + 8u * LT_STAGES, done_bar = empty0 + 8u * LT_STAGES;
  const uint32_t ncols = (N <= 32) ? 32u : (N <= 64) ? 64u : (N <= 128) ? 128u : 256u;

  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)),
                 "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  if (tid == 0) {
    for (int i = 0; i < LT_STAGES; ++i) {
      mbar_init(full0 + 8u * i, NPROD);       // one cp.async-completion arrival per producer thread
      mbar_init(empty0 + 8u * i, 1);          // one tcgen05.commit arrival
    }
    mbar_init(done_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_d = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  const int num_kb = (K + LT_BK - 1) / LT_BK;

  if (warp >= 1) {
    // ===== PRODUCERS: keep LT_STAGES k-blocks of LDGSTS copies in flight; never touch registers ====
    const int ptid = tid - 32;
    for (int kb = 0; kb < num_kb; ++kb) {
      const int st = kb % LT_STAGES;
      if (kb >= LT_STAGES) mbar_wait(empty0 + 8u * st, (uint32_t)((kb / LT_STAGES - 1) & 1));   // MMAs freed the slot
      const int k0 = kb * LT_BK;
      const uint32_t sa = smem0 + (uint32_t)st * stage_bytes, sb = sa + a_bytes;
      for (int idx = ptid; idx < LT_TM * CH; idx += NPROD) {
        const int r = idx / CH, c = idx % CH;
        const int gr = row0 + r, gk = k0 + c * 8;
        const bool ok = (gr < p.M) && (gk < K);
        cp_async16(sa + (uint32_t)(c * LT_TM + r) * 16u, p.x + (size_t)(ok ? gr : 0) * K + (ok ? gk : 0), ok ? 16 : 0);
      }
      for (int idx = ptid; idx < N * CH; idx += NPROD) {
        const int r = idx / CH, c = idx % CH;
        const int gk = k0 + c * 8;
        const bool ok = gk < K;
        cp_async16(sb + (uint32_t)(c * N + r) * 16u, p.w + (size_t)r * K + (ok ? gk : 0), ok ? 16 : 0);
      }
      // arrive on full[st] when all of THIS thread's copies above have landed
      asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(full0 + 8u * st) : "memory");
    }
  } else if (tid == 0) {
    // ===== MMA ISSUER: one thread, tcgen05.mma back to back as stages fill =========================
    const uint32_t idesc = make_instr_desc(N);
    for (int kb = 0; kb < num_kb; ++kb) {
      const int st = kb % LT_STAGES;
      mbar_wait(full0 + 8u * st, (uint32_t)((kb / LT_STAGES) & 1));
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // LDGSTS (generic proxy) -> UMMA (async proxy)
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const int rem = K - kb * LT_BK;
      const int ksteps = (rem >= LT_BK) ? (LT_BK / 16) : ((rem + 15) / 16);
      const uint32_t sa = smem0 + (uint32_t)st * stage_bytes, sb = sa + a_bytes;
      for (int s2 = 0; s2 < ksteps; ++s2) {
        const uint64_t da = make_smem_desc(sa + (uint32_t)s2 * 2u * LT_TM * 16u, LT_TM * 16u, 128u);
        const uint64_t db = make_smem_desc(sb + (uint32_t)s2 * 2u * (uint32_t)N * 16u, (uint32_t)N * 16u, 128u);
        const uint32_t accum = (kb > 0 || s2 > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accum)
            : "memory");
      }
      // free the slot for the producers once these MMAs have finished reading it
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(empty0 + 8u * st)
                   : "memory");
    }
    // accumulator complete (commits retire in issue order)
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(done_bar)
                 : "memory");
  }
  mbar_wait(done_bar, 0u);
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
      // N % 16 == 0, so every group of 8 columns is entirely inside or outside the matrix:
      // 16-byte (bf16) / 2 x 16-byte (fp32) vector stores, bias + ReLU fused here
#pragma unroll
      for (int g8 = 0; g8 < 4; ++g8) {
        const int c0 = cb + g8 * 8;
        if (c0 >= N) break;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          v[j] = __uint_as_float(r[g8 * 8 + j]) + (p.bias ? __ldg(p.bias + c0 + j) : 0.f);
          if (p.relu) v[j] = fmaxf(v[j], 0.f);
        }
        if (kOutBf16) {
          uint4 u;
          __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
          for (int j = 0; j < 4; ++j) h[j] = __floats2bfloat162_rn(v[2 * j], v[2 * j + 1]);
          *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(p.y) + (size_t)row * N + c0) = u;
        } else {
          float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(p.y) + (size_t)row * N + c0);
          dst[0] = make_float4(v[0], v[1], v[2], v[3]);
          dst[1] = make_float4(v[4], v[5], v[6], v[7]);
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
  const size_t smem = LT_STAGES * ((size_t)(LT_BK / 8) * LT_TM * 16 + (size_t)(LT_BK / 8) * p.N * 16) + 128;
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
