// eventgrad_b200 -- the convolutions of the ResNet family (3x3 stride 1/2, 1x1 stride 1/2, 3-channel stem: forward, data
// gradient, weight gradient) on the 5th-generation tensor cores at fp32 accuracy (sm_100a).
//
// The reference trains in fp32 (/root/reference/dcifar10/event/event.cpp:259-276 feeds fp32 tensors to
// torch::nn::Conv2d, resnet.hpp:3-9), and fp32 has no tensor-core path: cuDNN's fp32 kernels are SIMT (74 TFLOP/s
// peak on a B200) and take ~22 of the 24 ms of the headline step.  These kernels keep fp32 STORAGE and fp32
// ACCURACY but do the multiply-accumulate work with tcgen05.mma:
//
//   every fp32 operand is split once into three bf16 planes  x = x0 + x1 + x2  (8+8+8 mantissa bits: exact up to
//   2^-24 |x|), and the product is accumulated in fp32 in TMEM from the six significant cross terms
//       x0*w0 + x0*w1 + x1*w0 + x1*w1 + x0*w2 + x2*w0        (dropped: x1*w2, x2*w1, x2*w2 <= 2^-25 |x*w|)
//   so one "fp32 MMA" costs six bf16 MMAs: 2.25 PFLOP/s / 6 = 375 TFLOP/s of fp32-equivalent peak, 5x the SIMT
//   peak.  tests/test_gpu_conv_tc.py measures the error against an fp64 convolution next to cuDNN's fp32 error.
//
// Implicit GEMM, no im2col buffer: activations are NHWC planes [3][N][H][W][C]; the A tile of filter tap (r,s) is a
// SHIFTED window of the input, which a tiled 5-D TMA load fetches directly -- coordinates (c0, s-1, h0+r-1, n0, plane)
// with the hardware zero-filling everything outside the image (that IS the padding).
//
//   conv3x3_fprop_kernel   Y[pix, co] = sum_{tap,ci} X[pix+tap, ci] * Wt[co, tap, ci]        (forward; also the data
//                          gradient: X := dY planes, Wt := flipped/transposed weights [ci][8-tap][co])
//       persistent CTAs, tile 128 pixels x NT channels, K loop = 9 taps x Cin/64; warp 0 TMA producer, warp 1 MMA
//       issuer (24 tcgen05.mma per stage into a FRESH TMEM accumulator), warps 2-5 drain every k-block's accumulator
//       into fp32 registers (round-to-nearest adds: the tensor core itself truncates) and store the tile at the end.
//   conv3x3_wgrad_kernel   dW[tap, ci, co] = sum_pix X[pix+tap, ci] * dY[pix, co]            (weight gradient)
//       both operands are "MN-major" (the contiguous dimension is the channel, K = pixels), expressed with MN-major
//       SWIZZLE_128B descriptors over the very same TMA boxes; split over pixel ranges, partial sums in a workspace
//       (one split: dW is written directly).
//   conv_wgrad_reduce_kernel   fixed-tree (bitwise reproducible) sum of the partials + transpose to the OHWI layout.
//   conv_fprop_reduce_kernel   sum of the K-split partial tiles of the forward kernel on small grids (per-GPU batch 32).
//   Measured (B200, batch 256, 19.3 GFLOP per launch): forward 93-137 us, wgrad 84-163 us = 140-230 TFLOP/s
//   fp32-equivalent, 1e-7 rms against fp64 (cuDNN fp32: 2-4e-7); A/B history and ncu in profiles/README.md section 0a.
//   split3_kernel          fp32 -> three bf16 planes (flat; "parity" variant: the four (h%2, w%2) sub-images of the
//                          input of a STRIDE-2 conv, so that its taps become unit-stride boxes; "stem" variant: the 27
//                          (tap, rgb) values of the 3-channel stem gathered into one 64-wide K block = a 1x1 conv)
//   conv_wprep_kernel      weights fp32 [Co][T][Ci] -> planes [3][Co][T*Ci] (forward) and [3][Ci][T*Co] (data gradient).
//
// One kernel pair serves every convolution of the network through a TAP TABLE (ConvTcParams.dh/dw/src/wk): tap t
// reads source sub-image src[t] shifted by (dh[t], dw[t]) and the weight slice wk[t]:
//   3x3 stride 1      9 taps, shift (r-1, s-1)                      dgrad: same taps, weight slice 8-t of W^T
//   3x3 stride 2      9 taps over the parity images, shift in {-1,0}   dgrad: 4 launches, one per INPUT parity class
//                                                                   (1/2/2/4 taps over dY, scattered with stride 2)
//   1x1 stride 2      1 tap over parity image (0,0)                 dgrad: class (0,0) only, the rest is zero
//   stem (3 -> 64)    1 tap over the gathered 64-wide patches       (no data gradient: the input is the image)
#include "api.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace egb {
using namespace tc;

#define CV_THREADS 192
// TMEM accumulator ring: all 512 columns (8 accumulators at N tile 64, 4 at 128).  With two, the MMA of k-block i+2
// has to wait for the drain of k-block i (barrier wake-up + tcgen05.ld + 64-128 adds + arrive); a deeper ring takes that
// latency off the MMA's critical path.  (Measured neutral at batch 256 -- see the A/B list in profiles/README.md 0a.)
#define CV_HV 1   // independent TMEM accumulators per k-block that the K steps alternate between; 2 measured neutral (A/B list)
#define CV_NACC(NT) (512 / (CV_HV * (NT)))
#define CV_APLANE 16384u   // one bf16 plane of an operand tile with 128 rows: 128 x 128 B

// the six cross terms (plane of A, plane of B), largest last so that the small corrections are summed first
#define CV_TERM_A(t) ((t) == 0 ? 2 : ((t) == 2 || (t) == 3) ? 1 : 0)
#define CV_TERM_B(t) ((t) == 1 ? 2 : ((t) == 2 || (t) == 4) ? 1 : 0)

// ------------------------------------------------------------------------------------------------ forward / dgrad
template <int NT, int STAGES>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv3x3_fprop_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                     const ConvTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr uint32_t A_BYTES = 3u * CV_APLANE, B_PLANE = (uint32_t)NT * 128u, STAGE = A_BYTES + 3u * B_PLANE;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_unaligned = s_u32(smem_raw);
  const uint32_t smem0 = (smem_unaligned + 1023u) & ~1023u;
  unsigned char* smem_al = smem_raw + (smem0 - smem_unaligned);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + STAGES * STAGE);
  const uint32_t full0 = s_u32(bars), empty0 = full0 + 8u * STAGES;
  constexpr uint32_t NACC = CV_NACC(NT);
  const uint32_t tfull0 = empty0 + 8u * STAGES, tempty0 = tfull0 + 8u * NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NACC);
  constexpr uint32_t NCOLS = NACC * CV_HV * NT;            // the whole TMEM: 512 columns
  const int cpb = p.Ca / 64, num_kb_all = p.ntaps * cpb;
  const int n_tiles = p.Cb / NT;
  const int ksplits = p.ksplits;                           // > 1: few tiles (small batch): split the K loop over CTAs,
  const int kper = (num_kb_all + ksplits - 1) / ksplits;   //      partial tiles go to a workspace (launcher reduces)
  const int total = p.m_tiles * n_tiles * ksplits;
  const int tpi = p.bn == 1 ? p.H / p.bh : 1;              // tiles per image
#define CV_DECODE_WORK                                                         \
  const int ks = work % ksplits, tile_ = work / ksplits;                       \
  const int mt = tile_ / n_tiles, nt = tile_ - mt * n_tiles;                   \
  const int kb_lo = ks * kper, kb_hi = min(num_kb_all, kb_lo + kper);

  if (warp == 2) tmem_alloc(s_u32(tmem_slot), NCOLS);
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      bar_init(full0 + 8u * i, 1);
      bar_init(empty0 + 8u * i, 1);
    }
    for (int i = 0; i < (int)NACC; ++i) {
      bar_init(tfull0 + 8u * i, 1);
      bar_init(tempty0 + 8u * i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ===================== TMA PRODUCER =====================
    if (lane == 0) {
      uint32_t it = 0;
      for (int work = blockIdx.x; work < total; work += gridDim.x) {
        CV_DECODE_WORK
        const int n0 = p.bn == 1 ? mt / tpi : mt * p.bn;
        const int h0 = p.bn == 1 ? (mt - n0 * tpi) * p.bh : 0;
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          const int tap = kb / cpb, cc = kb - tap * cpb;
          const int dh = p.dh[tap], dw = p.dw[tap], src = p.src[tap], wk = p.wk[tap];
          bar_wait(empty0 + 8u * s, ph ^ 1u);
          bar_expect_tx(full0 + 8u * s, STAGE);
          const uint32_t sa = smem0 + s * STAGE;
#pragma unroll
          for (int pl = 0; pl < 3; ++pl)      // shifted window; rows outside the image arrive as zeros (= padding)
            tma_load_5d(sa + pl * CV_APLANE, &tmA, full0 + 8u * s, cc * 64, dw, h0 + dh, n0, pl * p.nsrc + src);
#pragma unroll
          for (int pl = 0; pl < 3; ++pl)
            tma_load_2d(sa + A_BYTES + pl * B_PLANE, &tmB, full0 + 8u * s, wk * p.Ca + cc * 64, pl * p.Cb + nt * NT);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA ISSUER (whole warp walks the loop, one elected lane issues) =====================
    {
      const uint32_t idesc = idesc_bf16_f32(128, NT, 0, 0);
      uint32_t it = 0;
      for (int work = blockIdx.x; work < total; work += gridDim.x) {
        CV_DECODE_WORK
        (void)mt; (void)nt;
        for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          const uint32_t as = it % NACC, aph = (it / NACC) & 1u;
          bar_wait(full0 + 8u * s, ph);                          // operands have landed
          bar_wait(tempty0 + 8u * as, aph ^ 1u);                 // the accumulator of k-block it-NACC has been drained
          tmem_fence_after();
          const uint32_t tmem_d = tmem_base + as * (CV_HV * NT);
          const uint32_t sa = smem0 + s * STAGE;
          // The tensor core TRUNCATES every accumulation to fp32 (measured: a bias of ~0.5 ulp per MMA towards zero,
          // growing linearly with K).  So each k-block starts a fresh accumulator that the epilogue warps add into
          // fp32 registers with round-to-nearest, and inside the k-block the five small correction terms come first:
          // only the last four MMAs (x0*w0) accumulate onto a full-magnitude value.
          if (elect_one_sync()) {
#pragma unroll
            for (int t = 0; t < 6; ++t) {
#pragma unroll
              for (int k4 = 0; k4 < 4; ++k4) {
                const uint64_t da = desc_k_sw128(sa + CV_TERM_A(t) * CV_APLANE) + (uint64_t)(k4 * 2);
                const uint64_t db = desc_k_sw128(sa + A_BYTES + CV_TERM_B(t) * B_PLANE) + (uint64_t)(k4 * 2);
                // consecutive MMAs alternate between CV_HV accumulators: no MMA waits for the accumulate of its predecessor
                umma_bf16(tmem_d + (uint32_t)(k4 % CV_HV) * NT, da, db, idesc, (k4 >= CV_HV || t > 0) ? 1u : 0u);
              }
            }
            umma_commit(empty0 + 8u * s);
            umma_commit(tfull0 + 8u * as);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ===================== EPILOGUE (4 warps, one TMEM lane quadrant each) =====================
    const int q = warp & 3;
    uint32_t it = 0;
    const long long M = (long long)p.N * p.H * p.W;
    for (int work = blockIdx.x; work < total; work += gridDim.x) {
      CV_DECODE_WORK
      float acc[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[j] = 0.f;
      for (int kb = kb_lo; kb < kb_hi; ++kb, ++it) {
        const uint32_t as = it % NACC, aph = (it / NACC) & 1u;
        bar_wait(tfull0 + 8u * as, aph);
        tmem_fence_after();
#pragma unroll
        for (int cb = 0; cb < NT; cb += 32) {
#pragma unroll
          for (int hv = 0; hv < CV_HV; ++hv) {
            uint32_t v[32];
            tmem_ld32(tmem_base + as * (CV_HV * NT) + (uint32_t)(hv * NT + cb) + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[cb + j] += __uint_as_float(v[j]);    // fp32 round-to-nearest
          }
        }
        tmem_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(tempty0 + 8u * as);
      }
      const long long row = (long long)mt * 128 + q * 32 + lane;
      if (ksplits > 1) {                                       // partial tile, dense rows: ws[ks][m_tiles*128][Cb]
        float4* dst = reinterpret_cast<float4*>(p.ws + (((size_t)ks * p.m_tiles * 128) + (size_t)row) * p.Cb + nt * NT);
#pragma unroll
        for (int j = 0; j < NT / 4; ++j) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
      } else if (row < M) {
        long long opix = row;
        if (p.os != 1 || p.OH != p.H || p.OW != p.W) {          // strided scatter (data gradient of a stride-2 conv)
          const int hw = p.H * p.W;
          const int n = (int)(row / hw), rem = (int)(row - (long long)n * hw);
          const int i = rem / p.W, j = rem - i * p.W;
          opix = ((long long)n * p.OH + i * p.os + p.op) * p.OW + j * p.os + p.oq;
        }
        float4* dst = reinterpret_cast<float4*>(p.out + (size_t)opix * p.Cb + nt * NT);
#pragma unroll
        for (int j = 0; j < NT / 4; ++j) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
      }
    }
  }
  tmem_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, NCOLS);
}

// out[pixel(row)][c] = sum_ks ws[ks][row][c] (fixed order), with the same output mapping as the un-split epilogue
__global__ void __launch_bounds__(256) conv_fprop_reduce_kernel(const ConvTcParams p) {
  const long long M = (long long)p.N * p.H * p.W;
  const int c4n = p.Cb / 4;
  const size_t zs = (size_t)p.m_tiles * 128 * p.Cb;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < M * c4n; i += (long long)gridDim.x * blockDim.x) {
    const long long row = i / c4n;
    const int c4 = (int)(i - row * c4n);
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int z = 0; z < p.ksplits; ++z) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(p.ws + z * zs + (size_t)row * p.Cb) + c4);
      a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w;
    }
    long long opix = row;
    if (p.os != 1 || p.OH != p.H || p.OW != p.W) {
      const int hw = p.H * p.W;
      const int n = (int)(row / hw), rem = (int)(row - (long long)n * hw);
      const int ii = rem / p.W, j = rem - ii * p.W;
      opix = ((long long)n * p.OH + ii * p.os + p.op) * p.OW + j * p.os + p.oq;
    }
    reinterpret_cast<float4*>(p.out + (size_t)opix * p.Cb)[c4] = a;
  }
}

// ------------------------------------------------------------------------------------------------ weight gradient
// grid (ceil(U/2), Cb/NT, splits); U = ntaps * Ca/64 "units" (tap, 64-channel block of the input); one CTA accumulates
// D[128 = two units][NT] over its range of 64-pixel K blocks and writes the partial to ws[split][unit*64 + ci][co].
template <int NT, int STAGES>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv3x3_wgrad_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmG,
                     const ConvTcParams p) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  constexpr uint32_t A_BYTES = 3u * CV_APLANE, B_PLANE = (uint32_t)NT * 128u, STAGE = A_BYTES + 3u * B_PLANE;
  constexpr uint32_t BLK = 8192u;                           // one [64 pixels][64 channels] block
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t smem_unaligned = s_u32(smem_raw);
  const uint32_t smem0 = (smem_unaligned + 1023u) & ~1023u;
  unsigned char* smem_al = smem_raw + (smem0 - smem_unaligned);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_al + STAGES * STAGE);
  const uint32_t full0 = s_u32(bars), empty0 = full0 + 8u * STAGES;
  constexpr uint32_t NACC = CV_NACC(NT);
  const uint32_t tfull0 = empty0 + 8u * STAGES, tempty0 = tfull0 + 8u * NACC;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 2 * NACC);
  constexpr uint32_t NCOLS = NACC * CV_HV * NT;
  const int cpb = p.Ca / 64, U = p.ntaps * cpb;
  const int u0 = 2 * blockIdx.x, u1 = (u0 + 1 < U) ? u0 + 1 : u0;
  const int per = (p.k_blocks + (int)gridDim.z - 1) / (int)gridDim.z;
  const int kb0 = blockIdx.z * per, kb1 = min(p.k_blocks, kb0 + per);
  const int nkb = kb1 > kb0 ? kb1 - kb0 : 0;
  const int bpi = p.bn == 1 ? p.H / p.bh : 1;              // 64-pixel blocks per image

  if (warp == 2) tmem_alloc(s_u32(tmem_slot), NCOLS);
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      bar_init(full0 + 8u * i, 1);
      bar_init(empty0 + 8u * i, 1);
    }
    for (int i = 0; i < (int)NACC; ++i) {
      bar_init(tfull0 + 8u * i, 1);
      bar_init(tempty0 + 8u * i, 4);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  tmem_fence_before();
  __syncthreads();
  tmem_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    if (lane == 0) {
      int tap[2], cb64[2];
      tap[0] = u0 / cpb; cb64[0] = u0 - tap[0] * cpb;
      tap[1] = u1 / cpb; cb64[1] = u1 - tap[1] * cpb;
      for (int i = 0; i < nkb; ++i) {
        const int kb = kb0 + i;
        const uint32_t s = (uint32_t)i % STAGES, ph = ((uint32_t)i / STAGES) & 1u;
        const int n0 = p.bn == 1 ? kb / bpi : kb * p.bn;
        const int h0 = p.bn == 1 ? (kb - n0 * bpi) * p.bh : 0;
        bar_wait(empty0 + 8u * s, ph ^ 1u);
        bar_expect_tx(full0 + 8u * s, STAGE);
        const uint32_t sa = smem0 + s * STAGE;
#pragma unroll
        for (int pl = 0; pl < 3; ++pl)
#pragma unroll
          for (int blk = 0; blk < 2; ++blk) {
            const int t = tap[blk];
            tma_load_5d(sa + pl * CV_APLANE + blk * BLK, &tmX, full0 + 8u * s, cb64[blk] * 64, p.dw[t], h0 + p.dh[t], n0,
                        pl * p.nsrc + p.src[t]);
          }
#pragma unroll
        for (int pl = 0; pl < 3; ++pl)
#pragma unroll
          for (int blk = 0; blk < NT / 64; ++blk)
            tma_load_5d(sa + A_BYTES + pl * B_PLANE + blk * BLK, &tmG, full0 + 8u * s, blockIdx.y * NT + blk * 64, 0, h0,
                        n0, pl);
      }
    }
  } else if (warp == 1) {
    {                                                       // whole warp walks the loop, one elected lane issues
      const uint32_t idesc = idesc_bf16_f32(128, NT, 1, 1);
      for (int i = 0; i < nkb; ++i) {
        const uint32_t s = (uint32_t)i % STAGES, ph = ((uint32_t)i / STAGES) & 1u;
        const uint32_t as = (uint32_t)i % NACC, aph = ((uint32_t)i / NACC) & 1u;
        bar_wait(full0 + 8u * s, ph);
        bar_wait(tempty0 + 8u * as, aph ^ 1u);
        tmem_fence_after();
        const uint32_t sa = smem0 + s * STAGE;
        // fresh accumulator per k-block, correction terms first (see conv3x3_fprop_kernel)
        if (elect_one_sync()) {
#pragma unroll
          for (int t = 0; t < 6; ++t) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {                 // 16 pixels (two 8-row atoms = 2048 B) per MMA
              const uint64_t da = desc_mn_sw128(sa + CV_TERM_A(t) * CV_APLANE + k4 * 2048u, BLK);
              const uint64_t db = desc_mn_sw128(sa + A_BYTES + CV_TERM_B(t) * B_PLANE + k4 * 2048u, BLK);
              umma_bf16(tmem_base + as * (CV_HV * NT) + (uint32_t)(k4 % CV_HV) * NT, da, db, idesc,
                        (k4 >= CV_HV || t > 0) ? 1u : 0u);
            }
          }
          umma_commit(empty0 + 8u * s);
          umma_commit(tfull0 + 8u * as);
        }
        __syncwarp();
      }
    }
  } else {
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int unit = m < 64 ? u0 : u1;
    const bool store = (m < 64) || (u1 != u0);
    float acc[NT];
#pragma unroll
    for (int j = 0; j < NT; ++j) acc[j] = 0.f;
    for (int i = 0; i < nkb; ++i) {
      const uint32_t as = (uint32_t)i % NACC, aph = ((uint32_t)i / NACC) & 1u;
      bar_wait(tfull0 + 8u * as, aph);
      tmem_fence_after();
#pragma unroll
      for (int cb = 0; cb < NT; cb += 32) {
#pragma unroll
        for (int hv = 0; hv < CV_HV; ++hv) {
          uint32_t v[32];
          tmem_ld32(tmem_base + as * (CV_HV * NT) + (uint32_t)(hv * NT + cb) + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[cb + j] += __uint_as_float(v[j]);
        }
      }
      tmem_fence_before();
      __syncwarp();
      if (lane == 0) bar_arrive(tempty0 + 8u * as);
    }
    if (store) {
      const size_t rows_all = (size_t)p.ntaps * p.Ca, my_row = (size_t)unit * 64 + (m & 63);
      if (gridDim.z == 1) {
        // one pixel split: this tile IS the result -- write dW[co][row] (OHWI) directly; lanes = consecutive rows
        float* dw = p.dwout + (size_t)(blockIdx.y * NT) * rows_all + my_row;
#pragma unroll
        for (int j = 0; j < NT; ++j) dw[(size_t)j * rows_all] = acc[j];
      } else {
        float4* dst = reinterpret_cast<float4*>(p.out + ((size_t)blockIdx.z * rows_all + my_row) * p.Cb + blockIdx.y * NT);
#pragma unroll
        for (int j = 0; j < NT / 4; ++j) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
      }
    }
  }
  tmem_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, NCOLS);
}

// dw[co][row] = sum_z ws[z][row][co]   (row = tap*Ca + ci; fixed summation order => bitwise reproducible)
__global__ void __launch_bounds__(256) conv_wgrad_reduce_kernel(const float* __restrict__ ws, float* __restrict__ dw,
                                                                int rows, int Cb, int splits) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  const int row0 = blockIdx.x * 32, co0 = blockIdx.y * 32;
  // sum over the splits as four interleaved chains (z mod 4) combined by a fixed tree: 16 loads in flight per thread
  // instead of 4 (the kernel is latency bound), and still bitwise reproducible
  float a4[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int c = 0; c < 4; ++c) a4[i][c] = 0.f;
  const size_t zs = (size_t)rows * Cb;
  for (int z0 = 0; z0 < splits; z0 += 4) {
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      if (z0 + c < splits) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const int r = row0 + ty + 8 * i;
          if (r < rows) a4[i][c] += __ldg(ws + (size_t)(z0 + c) * zs + (size_t)r * Cb + co0 + tx);
        }
      }
    }
  }
  float acc[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) acc[i] = (a4[i][0] + a4[i][1]) + (a4[i][2] + a4[i][3]);
#pragma unroll
  for (int i = 0; i < 4; ++i) tile[ty + 8 * i][tx] = acc[i];
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty + 8 * i, r = row0 + tx;
    if (r < rows) dw[(size_t)co * rows + r] = tile[tx][ty + 8 * i];
  }
}

// ------------------------------------------------------------------------------------------------ fp32 -> 3 x bf16
__device__ __forceinline__ void split3(float x, __nv_bfloat16& a, __nv_bfloat16& b, __nv_bfloat16& c) {
  a = __float2bfloat16_rn(x);
  const float r1 = __fsub_rn(x, __bfloat162float(a));        // exact
  b = __float2bfloat16_rn(r1);
  const float r2 = __fsub_rn(r1, __bfloat162float(b));       // exact
  c = __float2bfloat16_rn(r2);
}

__global__ void __launch_bounds__(256) split3_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                     size_t n8, size_t plane) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const F8 x = ld_f8(src + i * 8);
    __nv_bfloat16 a[8], b[8], c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split3(x.v[j], a[j], b[j], c[j]);
    *reinterpret_cast<uint4*>(dst + i * 8) = *reinterpret_cast<const uint4*>(a);
    *reinterpret_cast<uint4*>(dst + plane + i * 8) = *reinterpret_cast<const uint4*>(b);
    *reinterpret_cast<uint4*>(dst + 2 * plane + i * 8) = *reinterpret_cast<const uint4*>(c);
  }
}

// Input of a stride-2 conv: x [N][H][W][C] fp32 -> planes [3][4][N][H/2][W/2][C]; sub-image (h%2)*2 + (w%2) holds the
// pixels of that parity, so tap (r,s) of the strided conv is a UNIT-stride box of one sub-image shifted by -1 or 0.
__global__ void __launch_bounds__(256) split3_parity_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                            int N, int H, int W, int C) {
  const int c8 = C / 8;
  const size_t n8 = (size_t)N * H * W * c8, plane = (size_t)N * H * W * C;
  const int H2 = H / 2, W2 = W / 2;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i / c8;
    const int cg = (int)(i - pix * c8);
    const int w = (int)(pix % W);
    const size_t t = pix / W;
    const int h = (int)(t % H), n = (int)(t / H);
    const int sub = (h & 1) * 2 + (w & 1);
    const size_t o = ((((size_t)sub * N + n) * H2 + (h >> 1)) * W2 + (w >> 1)) * C + (size_t)cg * 8;
    const F8 x = ld_f8(src + i * 8);
    __nv_bfloat16 a[8], b[8], c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) split3(x.v[j], a[j], b[j], c[j]);
    *reinterpret_cast<uint4*>(dst + o) = *reinterpret_cast<const uint4*>(a);
    *reinterpret_cast<uint4*>(dst + plane + o) = *reinterpret_cast<const uint4*>(b);
    *reinterpret_cast<uint4*>(dst + 2 * plane + o) = *reinterpret_cast<const uint4*>(c);
  }
}

// Stem (Cin = 3): gather the 27 values (tap, rgb) of every output pixel's 3x3 window into ONE 64-wide K block
// (k = (r*3+s)*3 + c, zero beyond 27 and outside the image): the stem becomes a 1x1 convolution with K = 64.
// x [N][H][W][3] fp32 -> planes [3][N][H][W][64].
__global__ void __launch_bounds__(256) split3_stem_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst,
                                                          int N, int H, int W) {
  const size_t n8 = (size_t)N * H * W * 8, plane = (size_t)N * H * W * 64;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n8; i += (size_t)gridDim.x * blockDim.x) {
    const size_t pix = i >> 3;
    const int g = (int)(i & 7);
    const int w = (int)(pix % W);
    const size_t t = pix / W;
    const int h = (int)(t % H), n = (int)(t / H);
    __nv_bfloat16 a[8], b[8], c[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int k = g * 8 + j;
      float v = 0.f;
      if (k < 27) {
        const int tap = k / 3, ch = k - tap * 3;
        const int r = tap / 3, sx = tap - r * 3;
        const int hh = h + r - 1, ww = w + sx - 1;
        if (hh >= 0 && hh < H && ww >= 0 && ww < W) v = __ldg(src + (((size_t)n * H + hh) * W + ww) * 3 + ch);
      }
      split3(v, a[j], b[j], c[j]);
    }
    *reinterpret_cast<uint4*>(dst + i * 8) = *reinterpret_cast<const uint4*>(a);
    *reinterpret_cast<uint4*>(dst + plane + i * 8) = *reinterpret_cast<const uint4*>(b);
    *reinterpret_cast<uint4*>(dst + 2 * plane + i * 8) = *reinterpret_cast<const uint4*>(c);
  }
}

// Weights, once per step: w fp32 [Co][T][Ci] (OHWI) -> wp planes [3][Co][T*Ci] (same order: the forward's B operand)
// and, when wtp != null, wtp planes [3][Ci][T*Co] (wt[ci][t][co] = w[co][t][ci]: the data gradient's B operand).
// grid (Ci/32, Co/32, T), 256 threads (32 x 8).
__global__ void __launch_bounds__(256) conv_wprep_kernel(const float* __restrict__ w, __nv_bfloat16* __restrict__ wp,
                                                         __nv_bfloat16* __restrict__ wtp, int Co, int T, int Ci) {
  __shared__ float tile[32][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int ci0 = blockIdx.x * 32, co0 = blockIdx.y * 32, t = blockIdx.z;
  const size_t plane = (size_t)Co * T * Ci;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int co = co0 + ty + 8 * i;
    const size_t idx = ((size_t)co * T + t) * Ci + ci0 + tx;
    const float v = __ldg(w + idx);
    tile[ty + 8 * i][tx] = v;
    __nv_bfloat16 a, b, c;
    split3(v, a, b, c);
    wp[idx] = a;
    wp[plane + idx] = b;
    wp[2 * plane + idx] = c;
  }
  if (wtp == nullptr) return;
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int ci = ci0 + ty + 8 * i;
    const size_t idx = ((size_t)ci * T + t) * Co + co0 + tx;
    __nv_bfloat16 a, b, c;
    split3(tile[tx][ty + 8 * i], a, b, c);
    wtp[idx] = a;
    wtp[plane + idx] = b;
    wtp[2 * plane + idx] = c;
  }
}

// ------------------------------------------------------------------------------------------------ host side
static bool make_map_act(CUtensorMap* m, const void* base, int C, int W, int H, int N, int planes, int box_w, int box_h,
                         int box_n) {
  EgEncodeTiledFn enc = eg_get_encode_tiled();
  if (enc == nullptr) return false;
  const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, (cuuint64_t)planes};
  const cuuint64_t strides[4] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2,
                                 (cuuint64_t)N * H * W * C * 2};
  const cuuint32_t box[5] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, (cuuint32_t)box_n, 1};
  const cuuint32_t estr[5] = {1, 1, 1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}
static bool make_map_w(CUtensorMap* m, const void* base, uint64_t K, uint64_t rows, uint32_t box_rows) {
  EgEncodeTiledFn enc = eg_get_encode_tiled();
  if (enc == nullptr) return false;
  const cuuint64_t dims[2] = {K, rows};
  const cuuint64_t strides[1] = {K * 2};
  const cuuint32_t box[2] = {64, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// pixel tile of `rows` (128 or 64) rows: full image rows, bh of them, or bn whole images when an image is smaller
static bool tile_geometry(int N, int H, int W, int rows, int* bh, int* bn, int* tiles) {
  if (W < 1 || W > rows || rows % W) return false;
  int h = rows / W;
  if (h > H) h = H;
  if (H % h) return false;
  const int n = rows / (W * h);
  if (W * h * n != rows) return false;
  if (n > 1 && h != H) return false;
  *bh = h;
  *bn = n;
  *tiles = n == 1 ? N * (H / h) : (N + n - 1) / n;
  return true;
}

// (N, H, W) = the pixel grid of the GEMM rows (the conv's OUTPUT grid; dY's grid for a stride-2 data gradient)
bool conv_tc_supported(int N, int H, int W, int Ca, int Cb) {
  int bh, bn, t;
  return N >= 1 && Ca % 64 == 0 && Cb % 64 == 0 && Ca >= 64 && Cb >= 64 && tile_geometry(N, H, W, 128, &bh, &bn, &t) &&
         tile_geometry(N, H, W, 64, &bh, &bn, &t);
}

int conv_wgrad_splits(int N, int H, int W, int Ca, int Cb, int ntaps, int sm_count) {
  int bh, bn, kblocks;
  if (!tile_geometry(N, H, W, 64, &bh, &bn, &kblocks)) return 0;
  const int NT = (Cb % 128 == 0) ? 128 : 64;
  const int U = ntaps * (Ca / 64);
  const int work = ((U + 1) / 2) * (Cb / NT);
  int s = sm_count / work;
  if (s < 1) s = 1;
  if (s > kblocks) s = kblocks;
  return s;
}

static bool taps_ok(const ConvTcParams& p) {
  if (p.ntaps < 1 || p.ntaps > 9 || p.nsrc < 1 || p.nsrc > 4 || p.wtaps < 1 || p.wtaps > 9) return false;
  for (int t = 0; t < p.ntaps; ++t)
    if (p.src[t] < 0 || p.src[t] >= p.nsrc || p.wk[t] < 0 || p.wk[t] >= p.wtaps) return false;
  return true;
}

// K-loop splits of the forward kernel: only when the tiles alone leave most SMs idle (small per-GPU batch)
int conv_fprop_ksplits(int N, int H, int W, int Ca, int Cb, int ntaps, int sm_count) {
  int bh, bn, m_tiles;
  if (!tile_geometry(N, H, W, 128, &bh, &bn, &m_tiles)) return 0;
  const int NT = (Cb % 128 == 0) ? 128 : 64;
  const int tiles = m_tiles * (Cb / NT), num_kb = ntaps * (Ca / 64);
  if (tiles * 2 > sm_count || num_kb < 4) return 1;
  int ks = sm_count / tiles;
  if (ks > num_kb / 2) ks = num_kb / 2;                     // at least two k-blocks per split
  if (ks < 2) return 1;
  const int per = (num_kb + ks - 1) / ks;
  return (num_kb + per - 1) / per;                          // no empty split
}

int conv_fprop_mtiles(int N, int H, int W) {
  int bh, bn, m_tiles;
  return tile_geometry(N, H, W, 128, &bh, &bn, &m_tiles) ? m_tiles : 0;
}

template <int NT, int STAGES>
static cudaError_t fprop_launch(const CUtensorMap& tmA, const CUtensorMap& tmB, const ConvTcParams& p, int grid,
                                cudaStream_t s) {
  const size_t smem = (size_t)STAGES * (3 * CV_APLANE + 3 * NT * 128) + 256 + 1024;
  cudaError_t e = cudaFuncSetAttribute(conv3x3_fprop_kernel<NT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  conv3x3_fprop_kernel<NT, STAGES><<<grid, CV_THREADS, smem, s>>>(tmA, tmB, p);
  return cudaGetLastError();
}

// p.a = activation planes [3][nsrc][N][H][W][Ca], p.b = weight planes [3][Cb][wtaps*Ca], p.out fp32 [N][OH][OW][Cb]
cudaError_t launch_conv_fprop(const ConvTcParams& p0, int sm_count, cudaStream_t s) {
  ConvTcParams p = p0;
  if (!conv_tc_supported(p.N, p.H, p.W, p.Ca, p.Cb) || !taps_ok(p)) return cudaErrorInvalidValue;
  if (p.os < 1 || p.op < 0 || p.oq < 0 || p.op >= p.os || p.oq >= p.os || p.OH < p.H * p.os || p.OW < p.W * p.os)
    return cudaErrorInvalidValue;
  tile_geometry(p.N, p.H, p.W, 128, &p.bh, &p.bn, &p.m_tiles);
  if (p.ksplits < 1 || p.ws == nullptr) p.ksplits = 1;
  if (p.ksplits > 1) {                                      // the caller sized ws with conv_fprop_ksplits()
    const int want = conv_fprop_ksplits(p.N, p.H, p.W, p.Ca, p.Cb, p.ntaps, sm_count);
    if (p.ksplits != want) return cudaErrorInvalidValue;
  }
  CUtensorMap tmA, tmB;
  const int NT = (p.Cb % 128 == 0) ? 128 : 64;
  if (!make_map_w(&tmB, p.b, (uint64_t)p.wtaps * p.Ca, (uint64_t)3 * p.Cb, (uint32_t)NT)) return cudaErrorNotSupported;
  const int total = p.m_tiles * (p.Cb / NT) * p.ksplits;
  const int grid = total < sm_count ? total : sm_count;
  if (!make_map_act(&tmA, p.a, p.Ca, p.W, p.H, p.N, 3 * p.nsrc, p.W, p.bh, p.bn)) return cudaErrorNotSupported;
  eg_count_launch(EG_FAM_CONV, p.ksplits > 1 ? 2 : 1);
  cudaError_t e = NT == 128 ? fprop_launch<128, 2>(tmA, tmB, p, grid, s) : fprop_launch<64, 3>(tmA, tmB, p, grid, s);
  if (e != cudaSuccess || p.ksplits == 1) return e;
  const long long n4 = (long long)p.N * p.H * p.W * (p.Cb / 4);
  long long blocks = (n4 + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  conv_fprop_reduce_kernel<<<(unsigned)blocks, 256, 0, s>>>(p);
  return cudaGetLastError();
}

template <int NT, int STAGES>
static cudaError_t wgrad_launch(const CUtensorMap& tmX, const CUtensorMap& tmG, const ConvTcParams& p, dim3 grid,
                                cudaStream_t s) {
  const size_t smem = (size_t)STAGES * (3 * CV_APLANE + 3 * NT * 128) + 256 + 1024;
  cudaError_t e = cudaFuncSetAttribute(conv3x3_wgrad_kernel<NT, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  conv3x3_wgrad_kernel<NT, STAGES><<<grid, CV_THREADS, smem, s>>>(tmX, tmG, p);
  return cudaGetLastError();
}

// p.a = input planes [3][nsrc][N][H][W][Ca], p.b = dY planes [3][N][H][W][Cb], p.out = workspace
// [splits][ntaps*Ca][Cb], dw = [Cb][ntaps][Ca] (OHWI)
cudaError_t launch_conv_wgrad(const ConvTcParams& p0, float* dw, int splits, cudaStream_t s) {
  ConvTcParams p = p0;
  if (!conv_tc_supported(p.N, p.H, p.W, p.Ca, p.Cb) || !taps_ok(p) || splits < 1) return cudaErrorInvalidValue;
  tile_geometry(p.N, p.H, p.W, 64, &p.bh, &p.bn, &p.k_blocks);
  if (splits > p.k_blocks) splits = p.k_blocks;
  CUtensorMap tmX, tmG;
  const int NT = (p.Cb % 128 == 0) ? 128 : 64;
  if (!make_map_act(&tmX, p.a, p.Ca, p.W, p.H, p.N, 3 * p.nsrc, p.W, p.bh, p.bn)) return cudaErrorNotSupported;
  if (!make_map_act(&tmG, p.b, p.Cb, p.W, p.H, p.N, 3, p.W, p.bh, p.bn)) return cudaErrorNotSupported;
  const int U = p.ntaps * (p.Ca / 64);
  const dim3 grid((U + 1) / 2, p.Cb / NT, splits);
  p.dwout = dw;
  eg_count_launch(EG_FAM_CONV, splits > 1 ? 2 : 1);
  cudaError_t e = NT == 128 ? wgrad_launch<128, 2>(tmX, tmG, p, grid, s) : wgrad_launch<64, 3>(tmX, tmG, p, grid, s);
  if (e != cudaSuccess || splits == 1) return e;
  const int rows = p.ntaps * p.Ca;
  conv_wgrad_reduce_kernel<<<dim3((rows + 31) / 32, p.Cb / 32), 256, 0, s>>>(p.out, dw, rows, p.Cb, splits);
  return cudaGetLastError();
}

static unsigned split_blocks(size_t n8) {
  size_t blocks = (n8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  return (unsigned)blocks;
}

cudaError_t launch_split3(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t s) {
  if (n % 8) return cudaErrorInvalidValue;
  eg_count_launch(EG_FAM_CONV, 1);
  split3_kernel<<<split_blocks(n / 8), 256, 0, s>>>(src, dst, n / 8, n);
  return cudaGetLastError();
}

cudaError_t launch_split3_parity(const float* src, __nv_bfloat16* dst, int N, int H, int W, int C, cudaStream_t s) {
  if (C % 8 || H % 2 || W % 2) return cudaErrorInvalidValue;
  eg_count_launch(EG_FAM_CONV, 1);
  split3_parity_kernel<<<split_blocks((size_t)N * H * W * C / 8), 256, 0, s>>>(src, dst, N, H, W, C);
  return cudaGetLastError();
}

cudaError_t launch_split3_stem(const float* src, __nv_bfloat16* dst, int N, int H, int W, cudaStream_t s) {
  eg_count_launch(EG_FAM_CONV, 1);
  split3_stem_kernel<<<split_blocks((size_t)N * H * W * 8), 256, 0, s>>>(src, dst, N, H, W);
  return cudaGetLastError();
}

cudaError_t launch_conv_wprep(const float* w, __nv_bfloat16* wp, __nv_bfloat16* wtp, int Co, int T, int Ci, cudaStream_t s) {
  if (Co % 32 || Ci % 32 || T < 1) return cudaErrorInvalidValue;
  eg_count_launch(EG_FAM_CONV, 1);
  conv_wprep_kernel<<<dim3(Ci / 32, Co / 32, T), 256, 0, s>>>(w, wp, wtp, Co, T, Ci);
  return cudaGetLastError();
}

}  // namespace egb
