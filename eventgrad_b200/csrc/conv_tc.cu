// eventgrad_b200 -- 3x3 / stride-1 / pad-1 convolution on the 5th-generation tensor cores at fp32 accuracy (sm_100a).
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
//       SWIZZLE_128B descriptors over the very same TMA boxes; split over pixel ranges, partial sums in a workspace,
//   conv_wgrad_reduce_kernel   fixed-order (deterministic) sum of the partials + transpose to the OHWI weight layout.
//   split3_kernel          fp32 -> three bf16 planes.
#include "api.h"
#include "common.cuh"
#include "tc_common.cuh"

namespace egb {
using namespace tc;

#define CV_THREADS 192
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
  const uint32_t tfull0 = empty0 + 8u * STAGES, tempty0 = tfull0 + 16u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  constexpr uint32_t NCOLS = 2u * NT;                      // two accumulators (128 or 256 columns: powers of two)
  const int cpb = p.Ca / 64, num_kb = 9 * cpb;
  const int n_tiles = p.Cb / NT;
  const int total = p.m_tiles * n_tiles;
  const int tpi = p.bn == 1 ? p.H / p.bh : 1;              // tiles per image

  if (warp == 2) tmem_alloc(s_u32(tmem_slot), NCOLS);
  if (tid == 0) {
    for (int i = 0; i < STAGES; ++i) {
      bar_init(full0 + 8u * i, 1);
      bar_init(empty0 + 8u * i, 1);
    }
    for (int i = 0; i < 2; ++i) {
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
        const int mt = work / n_tiles, nt = work - mt * n_tiles;
        const int n0 = p.bn == 1 ? mt / tpi : mt * p.bn;
        const int h0 = p.bn == 1 ? (mt - n0 * tpi) * p.bh : 0;
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          const int tap = kb / cpb, cc = kb - tap * cpb;
          const int r = tap / 3, sx = tap - r * 3;
          bar_wait(empty0 + 8u * s, ph ^ 1u);
          bar_expect_tx(full0 + 8u * s, STAGE);
          const uint32_t sa = smem0 + s * STAGE;
#pragma unroll
          for (int pl = 0; pl < 3; ++pl)      // shifted window; rows outside the image arrive as zeros (= padding)
            tma_load_5d(sa + pl * CV_APLANE, &tmA, full0 + 8u * s, cc * 64, sx - 1, h0 + r - 1, n0, pl);
#pragma unroll
          for (int pl = 0; pl < 3; ++pl)
            tma_load_2d(sa + A_BYTES + pl * B_PLANE, &tmB, full0 + 8u * s, tap * p.Ca + cc * 64, pl * p.Cb + nt * NT);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA ISSUER =====================
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_f32(128, NT, 0, 0);
      uint32_t it = 0;
      for (int work = blockIdx.x; work < total; work += gridDim.x) {
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const uint32_t s = it % STAGES, ph = (it / STAGES) & 1u;
          const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
          bar_wait(full0 + 8u * s, ph);                          // operands have landed
          bar_wait(tempty0 + 8u * as, aph ^ 1u);                 // the accumulator of k-block it-2 has been drained
          tmem_fence_after();
          const uint32_t tmem_d = tmem_base + as * NT;
          const uint32_t sa = smem0 + s * STAGE;
          // The tensor core TRUNCATES every accumulation to fp32 (measured: a bias of ~0.5 ulp per MMA towards zero,
          // growing linearly with K).  So each k-block starts a fresh accumulator that the epilogue warps add into
          // fp32 registers with round-to-nearest, and inside the k-block the five small correction terms come first:
          // only the last four MMAs (x0*w0) accumulate onto a full-magnitude value.
#pragma unroll
          for (int t = 0; t < 6; ++t) {
#pragma unroll
            for (int k4 = 0; k4 < 4; ++k4) {
              const uint64_t da = desc_k_sw128(sa + CV_TERM_A(t) * CV_APLANE) + (uint64_t)(k4 * 2);
              const uint64_t db = desc_k_sw128(sa + A_BYTES + CV_TERM_B(t) * B_PLANE) + (uint64_t)(k4 * 2);
              umma_bf16(tmem_d, da, db, idesc, (k4 > 0 || t > 0) ? 1u : 0u);
            }
          }
          umma_commit(empty0 + 8u * s);
          umma_commit(tfull0 + 8u * as);
        }
      }
    }
  } else {
    // ===================== EPILOGUE (4 warps, one TMEM lane quadrant each) =====================
    const int q = warp & 3;
    uint32_t it = 0;
    const long long M = (long long)p.N * p.H * p.W;
    for (int work = blockIdx.x; work < total; work += gridDim.x) {
      const int mt = work / n_tiles, nt = work - mt * n_tiles;
      float acc[NT];
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[j] = 0.f;
      for (int kb = 0; kb < num_kb; ++kb, ++it) {
        const uint32_t as = it & 1u, aph = (it >> 1) & 1u;
        bar_wait(tfull0 + 8u * as, aph);
        tmem_fence_after();
#pragma unroll
        for (int cb = 0; cb < NT; cb += 32) {
          uint32_t v[32];
          tmem_ld32(tmem_base + as * NT + (uint32_t)cb + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
          for (int j = 0; j < 32; ++j) acc[cb + j] += __uint_as_float(v[j]);      // fp32 round-to-nearest
        }
        tmem_fence_before();
        __syncwarp();
        if (lane == 0) bar_arrive(tempty0 + 8u * as);
      }
      const long long row = (long long)mt * 128 + q * 32 + lane;
      if (row < M) {
        float4* dst = reinterpret_cast<float4*>(p.out + (size_t)row * p.Cb + nt * NT);
#pragma unroll
        for (int j = 0; j < NT / 4; ++j) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
      }
    }
  }
  tmem_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, NCOLS);
}

// ------------------------------------------------------------------------------------------------ weight gradient
// grid (ceil(U/2), Cb/NT, splits); U = 9 * Ca/64 "units" (tap, 64-channel block of the input); one CTA accumulates
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
  const uint32_t tfull0 = empty0 + 8u * STAGES, tempty0 = tfull0 + 16u;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  constexpr uint32_t NCOLS = 2u * NT;
  const int cpb = p.Ca / 64, U = 9 * cpb;
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
    for (int i = 0; i < 2; ++i) {
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
            const int r = tap[blk] / 3, sx = tap[blk] - r * 3;
            tma_load_5d(sa + pl * CV_APLANE + blk * BLK, &tmX, full0 + 8u * s, cb64[blk] * 64, sx - 1, h0 + r - 1, n0, pl);
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
    if (lane == 0) {
      const uint32_t idesc = idesc_bf16_f32(128, NT, 1, 1);
      for (int i = 0; i < nkb; ++i) {
        const uint32_t s = (uint32_t)i % STAGES, ph = ((uint32_t)i / STAGES) & 1u;
        const uint32_t as = (uint32_t)i & 1u, aph = ((uint32_t)i >> 1) & 1u;
        bar_wait(full0 + 8u * s, ph);
        bar_wait(tempty0 + 8u * as, aph ^ 1u);
        tmem_fence_after();
        const uint32_t sa = smem0 + s * STAGE;
        // fresh accumulator per k-block, correction terms first (see conv3x3_fprop_kernel)
#pragma unroll
        for (int t = 0; t < 6; ++t) {
#pragma unroll
          for (int k4 = 0; k4 < 4; ++k4) {                   // 16 pixels (two 8-row atoms = 2048 B) per MMA
            const uint64_t da = desc_mn_sw128(sa + CV_TERM_A(t) * CV_APLANE + k4 * 2048u, BLK);
            const uint64_t db = desc_mn_sw128(sa + A_BYTES + CV_TERM_B(t) * B_PLANE + k4 * 2048u, BLK);
            umma_bf16(tmem_base + as * NT, da, db, idesc, (k4 > 0 || t > 0) ? 1u : 0u);
          }
        }
        umma_commit(empty0 + 8u * s);
        umma_commit(tfull0 + 8u * as);
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
      const uint32_t as = (uint32_t)i & 1u, aph = ((uint32_t)i >> 1) & 1u;
      bar_wait(tfull0 + 8u * as, aph);
      tmem_fence_after();
#pragma unroll
      for (int cb = 0; cb < NT; cb += 32) {
        uint32_t v[32];
        tmem_ld32(tmem_base + as * NT + (uint32_t)cb + ((uint32_t)(q * 32) << 16), v);
#pragma unroll
        for (int j = 0; j < 32; ++j) acc[cb + j] += __uint_as_float(v[j]);
      }
      tmem_fence_before();
      __syncwarp();
      if (lane == 0) bar_arrive(tempty0 + 8u * as);
    }
    if (store) {
      float4* dst = reinterpret_cast<float4*>(
          p.out + ((size_t)blockIdx.z * (9 * p.Ca) + (size_t)unit * 64 + (m & 63)) * p.Cb + blockIdx.y * NT);
#pragma unroll
      for (int j = 0; j < NT / 4; ++j) dst[j] = make_float4(acc[4 * j], acc[4 * j + 1], acc[4 * j + 2], acc[4 * j + 3]);
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
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const size_t zs = (size_t)rows * Cb;
  for (int z = 0; z < splits; ++z) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int r = row0 + ty + 8 * i;
      if (r < rows) acc[i] += __ldg(ws + z * zs + (size_t)r * Cb + co0 + tx);
    }
  }
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

// ------------------------------------------------------------------------------------------------ host side
static bool make_map_act(CUtensorMap* m, const void* base, int C, int W, int H, int N, int box_w, int box_h, int box_n) {
  EgEncodeTiledFn enc = eg_get_encode_tiled();
  if (enc == nullptr) return false;
  const cuuint64_t dims[5] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N, 3};
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

bool conv_tc_supported(int N, int H, int W, int Ca, int Cb) {
  int bh, bn, t;
  return N >= 1 && Ca % 64 == 0 && Cb % 64 == 0 && Ca >= 64 && Cb >= 64 && tile_geometry(N, H, W, 128, &bh, &bn, &t) &&
         tile_geometry(N, H, W, 64, &bh, &bn, &t);
}

int conv_wgrad_splits(int N, int H, int W, int Ca, int Cb, int sm_count) {
  int bh, bn, kblocks;
  if (!tile_geometry(N, H, W, 64, &bh, &bn, &kblocks)) return 0;
  const int NT = (Cb % 128 == 0) ? 128 : 64;
  const int U = 9 * (Ca / 64);
  const int work = ((U + 1) / 2) * (Cb / NT);
  int s = sm_count / work;
  if (s < 1) s = 1;
  if (s > kblocks) s = kblocks;
  return s;
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

cudaError_t launch_conv3x3_fprop(const ConvTcParams& p0, int sm_count, cudaStream_t s) {
  ConvTcParams p = p0;
  if (!conv_tc_supported(p.N, p.H, p.W, p.Ca, p.Cb)) return cudaErrorInvalidValue;
  tile_geometry(p.N, p.H, p.W, 128, &p.bh, &p.bn, &p.m_tiles);
  CUtensorMap tmA, tmB;
  const int NT = (p.Cb % 128 == 0) ? 128 : 64;
  if (!make_map_act(&tmA, p.a, p.Ca, p.W, p.H, p.N, p.W, p.bh, p.bn)) return cudaErrorNotSupported;
  if (!make_map_w(&tmB, p.b, (uint64_t)9 * p.Ca, (uint64_t)3 * p.Cb, (uint32_t)NT)) return cudaErrorNotSupported;
  const int total = p.m_tiles * (p.Cb / NT);
  const int grid = total < sm_count ? total : sm_count;
  eg_count_launch(EG_FAM_CONV, 1);
  return NT == 128 ? fprop_launch<128, 2>(tmA, tmB, p, grid, s) : fprop_launch<64, 3>(tmA, tmB, p, grid, s);
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

// p.a = input planes [3][N][H][W][Ca], p.b = dY planes [3][N][H][W][Cb], p.out = workspace [splits][9*Ca][Cb],
// dw = [Cb][9][Ca] (OHWI)
cudaError_t launch_conv3x3_wgrad(const ConvTcParams& p0, float* dw, int splits, cudaStream_t s) {
  ConvTcParams p = p0;
  if (!conv_tc_supported(p.N, p.H, p.W, p.Ca, p.Cb) || splits < 1) return cudaErrorInvalidValue;
  tile_geometry(p.N, p.H, p.W, 64, &p.bh, &p.bn, &p.k_blocks);
  if (splits > p.k_blocks) splits = p.k_blocks;
  CUtensorMap tmX, tmG;
  const int NT = (p.Cb % 128 == 0) ? 128 : 64;
  if (!make_map_act(&tmX, p.a, p.Ca, p.W, p.H, p.N, p.W, p.bh, p.bn)) return cudaErrorNotSupported;
  if (!make_map_act(&tmG, p.b, p.Cb, p.W, p.H, p.N, p.W, p.bh, p.bn)) return cudaErrorNotSupported;
  const int U = 9 * (p.Ca / 64);
  const dim3 grid((U + 1) / 2, p.Cb / NT, splits);
  eg_count_launch(EG_FAM_CONV, 2);
  cudaError_t e = NT == 128 ? wgrad_launch<128, 2>(tmX, tmG, p, grid, s) : wgrad_launch<64, 3>(tmX, tmG, p, grid, s);
  if (e != cudaSuccess) return e;
  const int rows = 9 * p.Ca;
  conv_wgrad_reduce_kernel<<<dim3((rows + 31) / 32, p.Cb / 32), 256, 0, s>>>(p.out, dw, rows, p.Cb, splits);
  return cudaGetLastError();
}

cudaError_t launch_split3(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t s) {
  if (n % 8) return cudaErrorInvalidValue;
  const size_t n8 = n / 8;
  size_t blocks = (n8 + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  eg_count_launch(EG_FAM_CONV, 1);
  split3_kernel<<<(unsigned)blocks, 256, 0, s>>>(src, dst, n8, n);
  return cudaGetLastError();
}

}  // namespace egb
