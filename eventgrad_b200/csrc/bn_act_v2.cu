// eventgrad_b200 -- fused BatchNorm2d (+res)(+ReLU) v2: ReLU BIT MASK instead of re-reading y.  sm_100a.
//
// EXPERIMENTAL (opt-in with EGB_BN_V2=1, see ops/bn_act.py; csrc/bn_act.cu stays the default until
// this file has been run on hardware).
//
// Why: ncu on the split backward kernels of csrc/bn_act.cu (profiles/bn_slice_ncu_raw.md) shows both
// of them DRAM/latency bound at 2.4-2.8 TB/s with 105-107 registers (2 CTAs/SM); each reads THREE
// bf16 tensors (dy, x, y) where y is only needed for its sign.  Here the training forward emits one
// bit per element (y > 0, after the residual add) next to y, and the backward kernels read that
// mask: 1/16 of a tensor pass instead of a full one in both backward kernels -- 6R+2W -> 4.1R+2W tensor
// passes per BN+res+ReLU layer -- and 12 fewer live registers per row in flight (3 CTAs/SM).
//
// Mask layout (private to these kernels): [C/64 slices][M rows][8 bytes]; byte (slice,row,tx) holds the
// 8 channels slice*64 + tx*8 .. +7 of that row, bit e = channel tx*8+e.  A warp (4 rows x 8 threads)
// reads/writes 32 contiguous bytes = one sector.
//
// Block pattern served: /root/reference/dcifar10/common/resnet.hpp:39-52 (bn -> relu, bn -> += residual -> relu).
#include "bn_common.cuh"

namespace egb {

__device__ __forceinline__ void accumulate_stats(const uint4& u, float (&s)[8], float (&q)[8]) {
  const V8 x = unpack_bf16x8(u);
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    s[e] += x.v[e];
    q[e] = fmaf(x.v[e], x.v[e], q[e]);
  }
}

// ---------------------------------------------------------------------------------------------
// forward 1/2: per-slice sum / sum of squares (same scheme as bn_fwd_stats_kernel)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BN_THREADS, 3) bn2_fwd_stats_kernel(const BnParamsV2 pp) {
  const BnParams& p = pp.b;
  __shared__ __align__(16) float smem[BN_RPP * 128];
  const int slice = blockIdx.x, rs = blockIdx.y, RS = gridDim.y;
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const size_t coff = (size_t)slice * BN_SLICE + tx * 8;
  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  const long long stride = (long long)RS * BN_RPP;
  long long row = (long long)rs * BN_RPP + ty;
  for (; row + 7 * stride < p.M; row += 8 * stride) {
    uint4 xr[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) xr[u] = ldg16(p.x + (row + u * stride) * p.C + coff);
#pragma unroll
    for (int u = 0; u < 8; ++u) accumulate_stats(xr[u], s, q);
  }
  for (; row < p.M; row += stride) accumulate_stats(ldg16(p.x + row * p.C + coff), s, q);
  block_partials(s, q, smem, tx, ty, p.partial + ((size_t)slice * RS + rs) * 128);
  if (!elect_last_of_slice(p.ticket + slice, RS)) return;
  double* tot = reinterpret_cast<double*>(smem);
  slice_combine(p.partial + (size_t)slice * RS * 128, RS, tot);
  finalize_stats(p, slice, tot);
}

// ---------------------------------------------------------------------------------------------
// forward 2/2: y = act(x*sc + sh (+res)); mask bit = (y > 0)
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(BN_THREADS, 3) bn2_fwd_apply_kernel(const BnParamsV2 pp) {
  const BnParams& p = pp.b;
  const int slice = blockIdx.x, rs = blockIdx.y, RS = gridDim.y;
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const size_t coff = (size_t)slice * BN_SLICE + tx * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int c = slice * BN_SLICE + tx * 8 + e;
    sc[e] = p.gamma[c] * p.invstd[c];
    sh[e] = p.beta[c] - p.mean[c] * sc[e];
  }
  const long long stride = (long long)RS * BN_RPP;
  const bool has_res = p.res != nullptr;
  const bool relu = p.relu != 0;
  unsigned char* mrow = pp.mask + (size_t)slice * (size_t)p.M * 8 + tx;
  auto emit = [&](const uint4& xu, const uint4& ru, long long r2) {
    V8 x = unpack_bf16x8(xu);
    V8 r;
    if (has_res) r = unpack_bf16x8(ru);
    unsigned m = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      float v = fmaf(x.v[e], sc[e], sh[e]);
      if (has_res) v += r.v[e];
      if (relu) {
        // decide on the ROUNDED value so that the bit agrees with the stored bf16 y (a positive v can
        // only round to +0 if it is an fp32 denormal; keep the test exact anyway)
        const float vr = __bfloat162float(__float2bfloat16_rn(v));
        m |= (vr > 0.f) ? (1u << e) : 0u;
        v = fmaxf(v, 0.f);
      }
      x.v[e] = v;
    }
    stg16(p.y + r2 * p.C + coff, pack_bf16x8(x));
    if (relu) mrow[r2 * 8] = (unsigned char)m;
  };
  long long row = (long long)rs * BN_RPP + ty;
  for (; row + 2 * stride < p.M; row += 3 * stride) {      // 3 rows (x, res) in flight, no predicates
    uint4 xr[3], rr[3];
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      xr[u] = ldg16(p.x + (row + u * stride) * p.C + coff);
      if (has_res) rr[u] = ldg16(p.res + (row + u * stride) * p.C + coff);
    }
#pragma unroll
    for (int u = 0; u < 3; ++u) emit(xr[u], rr[u], row + u * stride);
  }
  for (; row < p.M; row += stride) {
    const uint4 xu = ldg16(p.x + row * p.C + coff);
    uint4 ru = make_uint4(0, 0, 0, 0);
    if (has_res) ru = ldg16(p.res + row * p.C + coff);
    emit(xu, ru, row);
  }
}

// dz = dy where the mask bit is set, else 0 -- done on the packed bf16 pairs (4 AND instead of 8 selects)
__device__ __forceinline__ uint4 apply_mask(const uint4& d, unsigned m) {
  uint4 o;
  const unsigned lo = 0x0000ffffu, hi = 0xffff0000u;
  o.x = d.x & (((m & 1u) ? lo : 0u) | ((m & 2u) ? hi : 0u));
  o.y = d.y & (((m & 4u) ? lo : 0u) | ((m & 8u) ? hi : 0u));
  o.z = d.z & (((m & 16u) ? lo : 0u) | ((m & 32u) ? hi : 0u));
  o.w = d.w & (((m & 64u) ? lo : 0u) | ((m & 128u) ? hi : 0u));
  return o;
}

// ---------------------------------------------------------------------------------------------
// backward 1/2: dbeta = sum dz, dgamma = sum dz * xhat
// ---------------------------------------------------------------------------------------------
#define BN2_RED_ROWS 3
__global__ void __launch_bounds__(BN_THREADS, 3) bn2_bwd_reduce_kernel(const BnParamsV2 pp) {
  const BnParams& p = pp.b;
  __shared__ __align__(16) float smem[BN_RPP * 128];
  const int slice = blockIdx.x, rs = blockIdx.y, RS = gridDim.y;
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const size_t coff = (size_t)slice * BN_SLICE + tx * 8;
  float mu[8], s1[8], s2[8];                              // s2 = sum dz*(x - mean); invstd is applied once at the end
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    mu[e] = p.mean[slice * BN_SLICE + tx * 8 + e];
    s1[e] = s2[e] = 0.f;
  }
  const bool relu = p.relu != 0;
  const unsigned char* mrow = pp.mask + (size_t)slice * (size_t)p.M * 8 + tx;
  const long long stride = (long long)RS * BN_RPP;
  for (long long row = (long long)rs * BN_RPP + ty; row < p.M; row += BN2_RED_ROWS * stride) {
    uint4 dr[BN2_RED_ROWS], xr[BN2_RED_ROWS];
    unsigned mk[BN2_RED_ROWS];
#pragma unroll
    for (int u = 0; u < BN2_RED_ROWS; ++u) {
      const long long r2 = row + u * stride;
      if (r2 < p.M) {
        dr[u] = ldg16(p.dy + r2 * p.C + coff);
        xr[u] = ldg16(p.x + r2 * p.C + coff);
        mk[u] = relu ? (unsigned)mrow[r2 * 8] : 0xffu;
      }
    }
#pragma unroll
    for (int u = 0; u < BN2_RED_ROWS; ++u) {
      if (row + u * stride >= p.M) continue;
      const V8 d = unpack_bf16x8(apply_mask(dr[u], mk[u])), x = unpack_bf16x8(xr[u]);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        s1[e] += d.v[e];
        s2[e] = fmaf(d.v[e], x.v[e] - mu[e], s2[e]);
      }
    }
  }
  block_partials(s1, s2, smem, tx, ty, p.partial + ((size_t)slice * RS + rs) * 128);
  if (!elect_last_of_slice(p.ticket + slice, RS)) return;
  double* tot = reinterpret_cast<double*>(smem);
  slice_combine(p.partial + (size_t)slice * RS * 128, RS, tot);
  if (threadIdx.x < BN_SLICE) {
    const int c = slice * BN_SLICE + threadIdx.x;
    p.dbeta[c] = (float)tot[threadIdx.x];
    p.dgamma[c] = (float)(tot[64 + threadIdx.x] * (double)p.invstd[c]);
  }
}

// ---------------------------------------------------------------------------------------------
// backward 2/2: dx = gamma*invstd*(dz - mean(dz) - xhat*mean(dz*xhat)); dres = dz
// ---------------------------------------------------------------------------------------------
#define BN2_DX_ROWS 3
__global__ void __launch_bounds__(BN_THREADS, 3) bn2_bwd_dx_kernel(const BnParamsV2 pp) {
  const BnParams& p = pp.b;
  const int slice = blockIdx.x, rs = blockIdx.y, RS = gridDim.y;
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const size_t coff = (size_t)slice * BN_SLICE + tx * 8;
  // dx = a*dz + b*x + c with a = gamma*is, b = -a*is*k2, c = -a*(k1 - mu*is*k2): 2 FMA per element
  float a[8], b[8], c[8];
  const float invM = 1.f / (float)p.M;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int ch = slice * BN_SLICE + tx * 8 + e;
    const float mu = p.mean[ch], is = p.invstd[ch];
    const float k1 = p.dbeta[ch] * invM, k2 = p.dgamma[ch] * invM;
    a[e] = p.gamma[ch] * is;
    b[e] = -a[e] * is * k2;
    c[e] = -a[e] * (k1 - mu * is * k2);
  }
  const bool relu = p.relu != 0;
  const unsigned char* mrow = pp.mask + (size_t)slice * (size_t)p.M * 8 + tx;
  const long long stride = (long long)RS * BN_RPP;
  const bool want_dres = p.dres != nullptr;
  auto emit = [&](const uint4& du, const uint4& xu, unsigned m, long long r2) {
    const uint4 dzp = apply_mask(du, m);
    const V8 d = unpack_bf16x8(dzp);
    V8 x = unpack_bf16x8(xu);
#pragma unroll
    for (int e = 0; e < 8; ++e) x.v[e] = fmaf(a[e], d.v[e], fmaf(b[e], x.v[e], c[e]));
    stg16(p.dx + r2 * p.C + coff, pack_bf16x8(x));
    if (want_dres) stg16(p.dres + r2 * p.C + coff, dzp);
  };
  long long row = (long long)rs * BN_RPP + ty;
  for (; row + (BN2_DX_ROWS - 1) * stride < p.M; row += BN2_DX_ROWS * stride) {
    uint4 dr[BN2_DX_ROWS], xr[BN2_DX_ROWS];
    unsigned mk[BN2_DX_ROWS];
#pragma unroll
    for (int u = 0; u < BN2_DX_ROWS; ++u) {
      const long long r2 = row + u * stride;
      dr[u] = ldg16(p.dy + r2 * p.C + coff);
      xr[u] = ldg16(p.x + r2 * p.C + coff);
      mk[u] = relu ? (unsigned)mrow[r2 * 8] : 0xffu;
    }
#pragma unroll
    for (int u = 0; u < BN2_DX_ROWS; ++u) emit(dr[u], xr[u], mk[u], row + u * stride);
  }
  for (; row < p.M; row += stride)
    emit(ldg16(p.dy + row * p.C + coff), ldg16(p.x + row * p.C + coff), relu ? (unsigned)mrow[row * 8] : 0xffu, row);
}

// which: 0 training forward (stats + apply + mask), 2 backward.  Eval forward stays on launch_bn.
cudaError_t launch_bn_v2(const BnParamsV2& pp, int which, int sm_count, cudaStream_t s) {
  const BnParams& p = pp.b;
  if (p.C % BN_SLICE != 0 || p.C / BN_SLICE > 32 || p.M < 1) return cudaErrorInvalidValue;
  if (p.relu && pp.mask == nullptr) return cudaErrorInvalidValue;
  const int slices = p.C / BN_SLICE;
  const long long passes = (p.M + BN_RPP - 1) / BN_RPP;
  auto grid = [&](long long passes_per_cta, long long cta_cap) {
    long long rsn = (passes + passes_per_cta - 1) / passes_per_cta;
    const long long cap = cta_cap / slices;
    if (rsn > cap) rsn = cap;
    if (rsn < 1) rsn = 1;
    return dim3((unsigned)slices, (unsigned)rsn);
  };
  const long long red_cap = 3LL * sm_count;              // reductions: one wave, 3 resident CTAs per SM
  const long long map_cap = 12LL * sm_count;             // maps: 4 waves
  if (which == 0) {
    bn2_fwd_stats_kernel<<<grid(8, red_cap), BN_THREADS, 0, s>>>(pp);
    bn2_fwd_apply_kernel<<<grid(3, map_cap), BN_THREADS, 0, s>>>(pp);
  } else if (which == 2) {
    bn2_bwd_reduce_kernel<<<grid(BN2_RED_ROWS * 2, red_cap), BN_THREADS, 0, s>>>(pp);
    bn2_bwd_dx_kernel<<<grid(BN2_DX_ROWS, map_cap), BN_THREADS, 0, s>>>(pp);
  } else {
    return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

}  // namespace egb
