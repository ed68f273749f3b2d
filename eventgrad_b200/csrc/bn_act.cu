// eventgrad_b200 -- fused BatchNorm2d (+ residual add) (+ ReLU), training forward + backward,
// channels-last bf16 activations, fp32 statistics/parameters.  sm_100a.
//
// Why: a launch-level profile of the flagship step (CIFAR ResNet, batch 256, bf16 autocast,
// profiles/launches_bench1_*.md) shows ATen's batch-norm + elementwise kernels taking ~2/3 of the
// GPU time while the tcgen05 convolutions take under 1/3.  The block pattern of the reference's
// ResNet (/root/reference/dcifar10/common/resnet.hpp:39-52: bn -> relu, bn -> += residual -> relu)
// is memory-bound glue; here it is 2 streaming kernels forward and 2 backward:
//
//   fwd  stats : per-channel sum / sum-of-squares partials over row slabs; the last CTA combines
//                them in fixed order (double), writes mean / invstd, updates running stats
//        apply : y = relu?( (x-mean)*invstd*gamma + beta (+ residual) )            -> bf16
//   bwd  reduce: dz = relu? dy*(y>0) : dy ;  sum dz, sum dz*xhat  -> dbeta, dgamma
//        dx    : dx = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)) ; dres = dz -> bf16
//
// Layout: NHWC == row-major [M = N*H*W, C].  A thread owns 8 consecutive channels (one 16-byte
// vector per row); TPR = C/8 threads span a row and 256/TPR rows are processed per CTA pass.
#include "api.h"
#include "common.cuh"

namespace egb {

#define BN_THREADS 256
#define BN_UNROLL 8

struct V8 {
  float v[8];
};

__device__ __forceinline__ V8 load_bf16x8(const __nv_bfloat16* p) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
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
__device__ __forceinline__ void store_bf16x8(__nv_bfloat16* p, const V8& r) {
  uint4 u;
  __nv_bfloat162* h = reinterpret_cast<__nv_bfloat162*>(&u);
#pragma unroll
  for (int i = 0; i < 4; ++i) h[i] = __floats2bfloat162_rn(r.v[2 * i], r.v[2 * i + 1]);
  *reinterpret_cast<uint4*>(p) = u;
}

// ---------------------------------------------------------------------------------------------
// Block-level: reduce per-thread 2x8 accumulators over the row lanes (ty) and write this CTA's
// partial [2][C] row; then the last CTA combines all partial rows in fixed order.
// smem: [rows_per_pass][2*C] floats = 256/TPR * 16*TPR * 4 B = 16 KB for every C.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_partials(const float (&a)[8], const float (&b)[8], float* smem, int C,
                                               int tx, int ty, int rpp, float* partial_row) {
  float* row = smem + (size_t)ty * 2 * C;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    row[tx * 8 + e] = a[e];
    row[C + tx * 8 + e] = b[e];
  }
  __syncthreads();
  for (int c = threadIdx.x; c < 2 * C; c += BN_THREADS) {
    float s = 0.f;
    for (int r = 0; r < rpp; ++r) s += smem[(size_t)r * 2 * C + c];
    partial_row[c] = s;
  }
}

// Last CTA: out[c] = sum_b partial[b][c] in double, fixed order.  All 256 threads take part:
// thread = (float4 column, row lane); each walks its rows with 4 loads in flight, then the row
// lanes are folded through shared memory in lane order.  smem needs lanes*W doubles.
__device__ __forceinline__ void final_combine(const float* partial, int nb, int W, double* sm) {
  const int W4 = W / 4;                                   // W = 2C is a multiple of 16
  const int lanes = (BN_THREADS >= W4) ? BN_THREADS / W4 : 1;
  for (int col0 = 0; col0 < W4; col0 += BN_THREADS) {     // only loops when W4 > 256 (C > 512)
    const int col = col0 + (int)(threadIdx.x % (lanes > 1 ? W4 : BN_THREADS));
    const int bl = (lanes > 1) ? (int)(threadIdx.x / W4) : 0;
    double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
    if (col < W4) {
      const float4* p4 = reinterpret_cast<const float4*>(partial) + col;
      int b = bl;
      for (; b + 7 * lanes < nb; b += 8 * lanes) {
        float4 x[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) x[u] = __ldcg(p4 + (size_t)(b + u * lanes) * W4);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          a0 += (double)x[u].x; a1 += (double)x[u].y; a2 += (double)x[u].z; a3 += (double)x[u].w;
        }
      }
      for (; b < nb; b += lanes) {
        const float4 x0 = __ldcg(p4 + (size_t)b * W4);
        a0 += (double)x0.x; a1 += (double)x0.y; a2 += (double)x0.z; a3 += (double)x0.w;
      }
      double* dst = sm + (size_t)bl * W + (size_t)col * 4;
      dst[0] = a0; dst[1] = a1; dst[2] = a2; dst[3] = a3;
    }
  }
  __syncthreads();
  if (lanes > 1) {
    for (int c = threadIdx.x; c < W; c += BN_THREADS) {
      double t = sm[c];
      for (int l = 1; l < lanes; ++l) t += sm[(size_t)l * W + c];
      sm[c] = t;                                           // row 0 holds the totals
    }
    __syncthreads();
  }
}

__device__ __forceinline__ bool elect_last_block(unsigned int* ticket) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(ticket, 1u);
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
    if (s_last) *ticket = 0u;
  }
  __syncthreads();
  if (s_last) __threadfence();
  return s_last != 0;
}

// ---------------------------------------------------------------------------------------------
// forward statistics
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BN_THREADS) bn_fwd_stats_kernel(const BnParams p) {
  extern __shared__ float smem[];
  const int C = p.C, TPR = C / 8, rpp = BN_THREADS / TPR;
  const int tx = threadIdx.x % TPR, ty = threadIdx.x / TPR;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  const long long stride = (long long)gridDim.x * rpp;
  long long row = (long long)blockIdx.x * rpp + ty;
  for (; row + (BN_UNROLL - 1) * stride < p.M; row += BN_UNROLL * stride) {
    V8 x[BN_UNROLL];
#pragma unroll
    for (int u = 0; u < BN_UNROLL; ++u) x[u] = load_bf16x8(p.x + (row + u * stride) * C + tx * 8);
#pragma unroll
    for (int u = 0; u < BN_UNROLL; ++u)
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s[e] += x[u].v[e];
        q[e] = fmaf(x[u].v[e], x[u].v[e], q[e]);
      }
  }
  for (; row < p.M; row += stride) {
    const V8 x = load_bf16x8(p.x + row * C + tx * 8);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s[e] += x.v[e];
      q[e] = fmaf(x.v[e], x.v[e], q[e]);
    }
  }
  block_partials(s, q, smem, C, tx, ty, rpp, p.partial + (size_t)blockIdx.x * 2 * C);
  if (!elect_last_block(p.ticket)) return;
  double* tot = reinterpret_cast<double*>(smem);          // 2C doubles (launch_bn sizes smem for it)
  final_combine(p.partial, gridDim.x, 2 * C, tot);
  const double invM = 1.0 / (double)p.M;
  for (int c = threadIdx.x; c < C; c += BN_THREADS) {
    const double mean = tot[c] * invM;
    double var = tot[C + c] * invM - mean * mean;            // biased variance
    if (var < 0.0) var = 0.0;
    p.mean[c] = (float)mean;
    p.invstd[c] = (float)(1.0 / sqrt(var + (double)p.eps));
    if (p.run_mean != nullptr) {                             // running stats (momentum, unbiased var)
      const double unb = p.M > 1 ? var * (double)p.M / (double)(p.M - 1) : var;
      p.run_mean[c] = (float)((1.0 - p.momentum) * (double)p.run_mean[c] + p.momentum * mean);
      p.run_var[c] = (float)((1.0 - p.momentum) * (double)p.run_var[c] + p.momentum * unb);
    }
  }
  if (threadIdx.x == 0 && p.nbt != nullptr) *p.nbt += 1;
}

// ---------------------------------------------------------------------------------------------
// forward apply: y = act(x*scale + shift (+res))
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BN_THREADS) bn_fwd_apply_kernel(const BnParams p) {
  const int C = p.C, TPR = C / 8, rpp = BN_THREADS / TPR;
  const int tx = threadIdx.x % TPR, ty = threadIdx.x / TPR;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = tx * 8 + e;
    sc[e] = p.gamma[c] * p.invstd[c];
    sh[e] = p.beta[c] - p.mean[c] * sc[e];
  }
  const long long stride = (long long)gridDim.x * rpp;
  const bool has_res = p.res != nullptr;
  for (long long row = (long long)blockIdx.x * rpp + ty; row < p.M; row += 2 * stride) {
    const long long row2 = row + stride;
    const bool two = row2 < p.M;
    const size_t o1 = (size_t)row * C + tx * 8, o2 = (size_t)row2 * C + tx * 8;
    V8 x1 = load_bf16x8(p.x + o1), x2, r1, r2;
    if (two) x2 = load_bf16x8(p.x + o2);
    if (has_res) {
      r1 = load_bf16x8(p.res + o1);
      if (two) r2 = load_bf16x8(p.res + o2);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = fmaf(x1.v[e], sc[e], sh[e]);
      if (has_res) v += r1.v[e];
      x1.v[e] = p.relu ? fmaxf(v, 0.f) : v;
    }
    store_bf16x8(p.y + o1, x1);
    if (two) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = fmaf(x2.v[e], sc[e], sh[e]);
        if (has_res) v += r2.v[e];
        x2.v[e] = p.relu ? fmaxf(v, 0.f) : v;
      }
      store_bf16x8(p.y + o2, x2);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward reduce: sum dz, sum dz*xhat   (dz = relu-masked dy)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_reduce_kernel(const BnParams p) {
  extern __shared__ float smem[];
  const int C = p.C, TPR = C / 8, rpp = BN_THREADS / TPR;
  const int tx = threadIdx.x % TPR, ty = threadIdx.x / TPR;
  float mu[8], is[8], s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mu[e] = p.mean[tx * 8 + e];
    is[e] = p.invstd[tx * 8 + e];
    s1[e] = s2[e] = 0.f;
  }
  const long long stride = (long long)gridDim.x * rpp;
  for (long long row = (long long)blockIdx.x * rpp + ty; row < p.M; row += 2 * stride) {
    const long long row2 = row + stride;
    const bool two = row2 < p.M;
    const size_t o1 = (size_t)row * C + tx * 8, o2 = (size_t)row2 * C + tx * 8;
    V8 d1 = load_bf16x8(p.dy + o1), x1 = load_bf16x8(p.x + o1), y1, d2, x2, y2;
    if (p.relu) y1 = load_bf16x8(p.y + o1);
    if (two) {
      d2 = load_bf16x8(p.dy + o2);
      x2 = load_bf16x8(p.x + o2);
      if (p.relu) y2 = load_bf16x8(p.y + o2);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float dz = (p.relu && !(y1.v[e] > 0.f)) ? 0.f : d1.v[e];
      s1[e] += dz;
      s2[e] = fmaf(dz, (x1.v[e] - mu[e]) * is[e], s2[e]);
    }
    if (two) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dz = (p.relu && !(y2.v[e] > 0.f)) ? 0.f : d2.v[e];
        s1[e] += dz;
        s2[e] = fmaf(dz, (x2.v[e] - mu[e]) * is[e], s2[e]);
      }
    }
  }
  block_partials(s1, s2, smem, C, tx, ty, rpp, p.partial + (size_t)blockIdx.x * 2 * C);
  if (!elect_last_block(p.ticket)) return;
  double* tot = reinterpret_cast<double*>(smem);
  final_combine(p.partial, gridDim.x, 2 * C, tot);
  for (int c = threadIdx.x; c < C; c += BN_THREADS) {
    p.dbeta[c] = (float)tot[c];
    p.dgamma[c] = (float)tot[C + c];
  }
}

// ---------------------------------------------------------------------------------------------
// backward dx (+ dres)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BN_THREADS) bn_bwd_dx_kernel(const BnParams p) {
  const int C = p.C, TPR = C / 8, rpp = BN_THREADS / TPR;
  const int tx = threadIdx.x % TPR, ty = threadIdx.x / TPR;
  float mu[8], is[8], k0[8], k1[8], k2[8];
  const float invM = 1.f / (float)p.M;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = tx * 8 + e;
    mu[e] = p.mean[c];
    is[e] = p.invstd[c];
    k0[e] = p.gamma[c] * is[e];              // dx = k0*(dz - k1 - xhat*k2)
    k1[e] = p.dbeta[c] * invM;
    k2[e] = p.dgamma[c] * invM;
  }
  const long long stride = (long long)gridDim.x * rpp;
  const bool want_dres = p.dres != nullptr;
  for (long long row = (long long)blockIdx.x * rpp + ty; row < p.M; row += 2 * stride) {
    const long long row2 = row + stride;
    const bool two = row2 < p.M;
    const size_t o1 = (size_t)row * C + tx * 8, o2 = (size_t)row2 * C + tx * 8;
    V8 d1 = load_bf16x8(p.dy + o1), x1 = load_bf16x8(p.x + o1), y1, d2, x2, y2;
    if (p.relu) y1 = load_bf16x8(p.y + o1);
    if (two) {
      d2 = load_bf16x8(p.dy + o2);
      x2 = load_bf16x8(p.x + o2);
      if (p.relu) y2 = load_bf16x8(p.y + o2);
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float dz = (p.relu && !(y1.v[e] > 0.f)) ? 0.f : d1.v[e];
      d1.v[e] = dz;
      x1.v[e] = k0[e] * (dz - k1[e] - (x1.v[e] - mu[e]) * is[e] * k2[e]);
    }
    store_bf16x8(p.dx + o1, x1);
    if (want_dres) store_bf16x8(p.dres + o1, d1);
    if (two) {
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        const float dz = (p.relu && !(y2.v[e] > 0.f)) ? 0.f : d2.v[e];
        d2.v[e] = dz;
        x2.v[e] = k0[e] * (dz - k1[e] - (x2.v[e] - mu[e]) * is[e] * k2[e]);
      }
      store_bf16x8(p.dx + o2, x2);
      if (want_dres) store_bf16x8(p.dres + o2, d2);
    }
  }
}

// ---------------------------------------------------------------------------------------------
static int bn_grid(const BnParams& p, int max_ctas, int rows_per_cta_min) {
  const int TPR = p.C / 8, rpp = BN_THREADS / TPR;
  long long passes = (p.M + rpp - 1) / rpp;
  long long g = (passes + rows_per_cta_min - 1) / rows_per_cta_min;
  if (g < 1) g = 1;
  if (g > max_ctas) g = max_ctas;
  return (int)g;
}

int bn_partial_rows(int sm_count) { return sm_count * 4; }

cudaError_t launch_bn(const BnParams& p, int which, int sm_count, cudaStream_t s) {
  const int TPR = p.C / 8;
  if (p.C % 8 != 0 || TPR < 1 || TPR > BN_THREADS || (BN_THREADS % TPR) != 0) return cudaErrorInvalidValue;
  size_t smem = (size_t)(BN_THREADS / TPR) * 2 * p.C * sizeof(float);   // 16 KB staging
  {   // final combine: lanes * 2C doubles
    const int W4 = 2 * p.C / 4;
    const int lanes = (BN_THREADS >= W4) ? BN_THREADS / W4 : 1;
    const size_t need = (size_t)lanes * 2 * p.C * sizeof(double);
    if (smem < need) smem = need;
  }
  // reductions: few, fat CTAs (8 x 16 B loads in flight per thread) keep the fixed-order combine of
  // the partial rows short: rows * 2C is bounded by ~32K floats
  int red_cap = 32768 / (2 * p.C);
  if (red_cap > bn_partial_rows(sm_count)) red_cap = bn_partial_rows(sm_count);
  if (red_cap < 16) red_cap = 16;
  const int red_grid = bn_grid(p, red_cap, 8);
  const int map_grid = bn_grid(p, sm_count * 8, 4);
  switch (which) {
    case 0: bn_fwd_stats_kernel<<<red_grid, BN_THREADS, smem, s>>>(p); break;
    case 1: bn_fwd_apply_kernel<<<map_grid, BN_THREADS, 0, s>>>(p); break;
    case 2: bn_bwd_reduce_kernel<<<red_grid, BN_THREADS, smem, s>>>(p); break;
    case 3: bn_bwd_dx_kernel<<<map_grid, BN_THREADS, 0, s>>>(p); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace egb
