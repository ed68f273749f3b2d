// eventgrad_b200 -- fused BatchNorm2d (+res)(+ReLU), ONE launch per direction with thread-block CLUSTERS
// and distributed shared memory.  sm_100a.  EXPERIMENTAL (EGB_BN_CLUSTER=1 on top of EGB_BN_V2=1).
//
// Why: at the per-GPU batch of the 8-GPU configuration (32 images) the step is kernel-count bound: every BN
// layer costs two launches per direction (stats -> apply, reduce -> dx) of ~4-5 us each, 112 launches per step,
// about a third of the 1.42 ms.  The earlier single-launch attempt (bn_*_fused_kernel in bn_act.cu) synchronised
// its CTAs through global-memory partials + an atomic ticket + a spin flag and was SLOWER (11-12 us per launch).
// Here a cluster of CS <= 16 CTAs owns one 64-channel slice:
//
//   1. every CTA streams its rows of the slice into shared memory with cp.async (x is read from HBM/L2 ONCE),
//   2. reduces them to a [128] partial in its own shared memory,
//   3. barrier.cluster (hardware, ~1 us) ... every CTA sums the CS partials straight out of its peers' shared
//      memory (DSMEM, ld.shared::cluster) in rank order -> identical statistics in every CTA,
//   4. normalises its slab from shared memory and stores y (+ the ReLU bit mask of bn_act_v2.cu).
//
// The backward is the same shape with two slabs (dy, x) and the bit mask.  Tensors whose slice does not fit
// (rows per CTA x 128 B x slabs > ~190 KB at the largest schedulable cluster) return "not taken" and the caller
// uses the two-launch path of bn_act_v2.cu.
//
// Block pattern served: /root/reference/dcifar10/common/resnet.hpp:39-52.
#include <cooperative_groups.h>

#include "bn_common.cuh"

namespace cg = cooperative_groups;

namespace egb {

#define BNC_FWD_ROWS 1536      // rows of one slab per CTA (192 KB)
#define BNC_BWD_ROWS 768       // two slabs

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Cluster-wide totals of the per-CTA partial rows (128 floats each), fixed rank order, double accumulate.
// On return tot[0..127] (shared, this CTA) holds the totals; includes the barrier that makes peers' partials
// visible and the one that keeps every CTA alive until all peers have read its partial.
__device__ __forceinline__ void cluster_totals(cg::cluster_group& cluster, float* my_part, double* tot) {
  cluster.sync();                                           // every CTA's partial is written (release/acquire)
  const unsigned cs = cluster.num_blocks();
  if (threadIdx.x < 128) {
    double t = 0.0;
    for (unsigned k = 0; k < cs; ++k) {
      const float* peer = cluster.map_shared_rank(my_part, k);
      t += (double)peer[threadIdx.x];
    }
    tot[threadIdx.x] = t;
  }
  cluster.sync();                                           // nobody exits / reuses `my_part` while a peer still reads it
}

// ===========================================================================================
// forward: stats + normalise (+res)(+relu)(+mask), one launch
// ===========================================================================================
__global__ void __launch_bounds__(BN_THREADS, 1) bn_cluster_fwd_kernel(const BnParamsV2 pp) {
  const BnParams& p = pp.b;
  extern __shared__ __align__(16) unsigned char slab_raw[];            // [rows_per_cta][128 B]
  __shared__ __align__(16) float red[BN_RPP * 128];
  __shared__ __align__(16) float part[128];
  __shared__ double tot[128];
  __shared__ float s_sc[BN_SLICE], s_sh[BN_SLICE];
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
  const int slice = blockIdx.x / cs;
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const size_t coff = (size_t)slice * BN_SLICE + tx * 8;
  const long long rows_per = ((p.M + cs - 1) / cs + BN_RPP - 1) / BN_RPP * BN_RPP;
  const long long row0 = (long long)crank * rows_per;
  long long nrows = p.M - row0;
  if (nrows > rows_per) nrows = rows_per;
  if (nrows < 0) nrows = 0;
  uint4* slab = reinterpret_cast<uint4*>(slab_raw);

  for (long long r = ty; r < nrows; r += BN_RPP) cp_async16(slab + r * 8 + tx, p.x + (row0 + r) * p.C + coff);
  cp_async_wait_all();
  __syncthreads();

  float s[8], q[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s[e] = q[e] = 0.f;
  for (long long r = ty; r < nrows; r += BN_RPP) {
    const V8 x = unpack_bf16x8(slab[r * 8 + tx]);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s[e] += x.v[e];
      q[e] = fmaf(x.v[e], x.v[e], q[e]);
    }
  }
  block_partials(s, q, red, tx, ty, part);
  __syncthreads();
  cluster_totals(cluster, part, tot);
  __syncthreads();
  if (threadIdx.x < BN_SLICE) {
    const int c = slice * BN_SLICE + threadIdx.x;
    const double invM = 1.0 / (double)p.M;
    const double mean = tot[threadIdx.x] * invM;
    double var = tot[64 + threadIdx.x] * invM - mean * mean;             // biased
    if (var < 0.0) var = 0.0;
    const float is = rsqrtf((float)var + p.eps);
    const float sc = p.gamma[c] * is;
    s_sc[threadIdx.x] = sc;
    s_sh[threadIdx.x] = p.beta[c] - (float)mean * sc;
    if (crank == 0) {                                                    // one writer per slice
      p.mean[c] = (float)mean;
      p.invstd[c] = is;
      if (p.run_mean != nullptr) {
        const double unb = p.M > 1 ? var * (double)p.M / (double)(p.M - 1) : var;
        p.run_mean[c] = (float)((1.0 - p.momentum) * (double)p.run_mean[c] + p.momentum * mean);
        p.run_var[c] = (float)((1.0 - p.momentum) * (double)p.run_var[c] + p.momentum * unb);
      }
    }
  }
  if (threadIdx.x == 0 && blockIdx.x == 0 && p.nbt != nullptr) *p.nbt += 1;
  __syncthreads();

  float sc[8], sh[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    sc[e] = s_sc[tx * 8 + e];
    sh[e] = s_sh[tx * 8 + e];
  }
  const bool has_res = p.res != nullptr, relu = p.relu != 0;
  unsigned char* mrow = pp.mask + (size_t)slice * (size_t)p.M * 8 + tx;
  for (long long r = ty; r < nrows; r += 2 * BN_RPP) {
    uint4 rr[2];
    if (has_res) {
#pragma unroll
      for (int u = 0; u < 2; ++u)
        if (r + u * BN_RPP < nrows) rr[u] = ldg16(p.res + (row0 + r + u * BN_RPP) * p.C + coff);
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const long long rl = r + u * BN_RPP;
      if (rl >= nrows) continue;
      V8 x = unpack_bf16x8(slab[rl * 8 + tx]);
      V8 rs;
      if (has_res) rs = unpack_bf16x8(rr[u]);
      unsigned m = 0;
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        float v = fmaf(x.v[e], sc[e], sh[e]);
        if (has_res) v += rs.v[e];
        if (relu) {
          const float vr = __bfloat162float(__float2bfloat16_rn(v));     // bit agrees with the stored bf16 y
          m |= (vr > 0.f) ? (1u << e) : 0u;
          v = fmaxf(v, 0.f);
        }
        x.v[e] = v;
      }
      stg16(p.y + (row0 + rl) * p.C + coff, pack_bf16x8(x));
      if (relu) mrow[(row0 + rl) * 8] = (unsigned char)m;
    }
  }
}

__device__ __forceinline__ uint4 mask_bf16x8(const uint4& d, unsigned m) {
  uint4 o;
  const unsigned lo = 0x0000ffffu, hi = 0xffff0000u;
  o.x = d.x & (((m & 1u) ? lo : 0u) | ((m & 2u) ? hi : 0u));
  o.y = d.y & (((m & 4u) ? lo : 0u) | ((m & 8u) ? hi : 0u));
  o.z = d.z & (((m & 16u) ? lo : 0u) | ((m & 32u) ? hi : 0u));
  o.w = d.w & (((m & 64u) ? lo : 0u) | ((m & 128u) ? hi : 0u));
  return o;
}

// ===========================================================================================
// backward: reduce + dx (+dres), one launch.  The dy slab is overwritten in place with dz = dy & mask.
// ===========================================================================================
__global__ void __launch_bounds__(BN_THREADS, 1) bn_cluster_bwd_kernel(const BnParamsV2 pp) {
  const BnParams& p = pp.b;
  extern __shared__ __align__(16) unsigned char slab_raw[];            // [rows][128 B] dz | [rows][128 B] x
  __shared__ __align__(16) float red[BN_RPP * 128];
  __shared__ __align__(16) float part[128];
  __shared__ double tot[128];
  __shared__ float s_a[BN_SLICE], s_b[BN_SLICE], s_c[BN_SLICE];
  cg::cluster_group cluster = cg::this_cluster();
  const int cs = (int)cluster.num_blocks(), crank = (int)cluster.block_rank();
  const int slice = blockIdx.x / cs;
  const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
  const size_t coff = (size_t)slice * BN_SLICE + tx * 8;
  const long long rows_per = ((p.M + cs - 1) / cs + BN_RPP - 1) / BN_RPP * BN_RPP;
  const long long row0 = (long long)crank * rows_per;
  long long nrows = p.M - row0;
  if (nrows > rows_per) nrows = rows_per;
  if (nrows < 0) nrows = 0;
  uint4* sdz = reinterpret_cast<uint4*>(slab_raw);
  uint4* sx = sdz + (size_t)rows_per * 8;
  const bool relu = p.relu != 0;
  const unsigned char* mrow = pp.mask + (size_t)slice * (size_t)p.M * 8 + tx;

  for (long long r = ty; r < nrows; r += BN_RPP) {
    cp_async16(sdz + r * 8 + tx, p.dy + (row0 + r) * p.C + coff);
    cp_async16(sx + r * 8 + tx, p.x + (row0 + r) * p.C + coff);
  }
  float mu[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) mu[e] = p.mean[slice * BN_SLICE + tx * 8 + e];
  cp_async_wait_all();
  __syncthreads();

  float s1[8], s2[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) s1[e] = s2[e] = 0.f;
  for (long long r = ty; r < nrows; r += BN_RPP) {
    uint4 dzp = sdz[r * 8 + tx];
    if (relu) {
      dzp = mask_bf16x8(dzp, (unsigned)mrow[(row0 + r) * 8]);
      sdz[r * 8 + tx] = dzp;                                             // each element is owned by one thread
    }
    const V8 d = unpack_bf16x8(dzp), x = unpack_bf16x8(sx[r * 8 + tx]);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      s1[e] += d.v[e];
      s2[e] = fmaf(d.v[e], x.v[e] - mu[e], s2[e]);
    }
  }
  block_partials(s1, s2, red, tx, ty, part);
  __syncthreads();
  cluster_totals(cluster, part, tot);
  __syncthreads();
  if (threadIdx.x < BN_SLICE) {
    const int c = slice * BN_SLICE + threadIdx.x;
    const float is = p.invstd[c], mean = p.mean[c];
    const float dbeta = (float)tot[threadIdx.x];
    const float dgamma = (float)(tot[64 + threadIdx.x] * (double)is);
    if (crank == 0) {
      p.dbeta[c] = dbeta;
      p.dgamma[c] = dgamma;
    }
    const float invM = 1.f / (float)p.M;
    const float k1 = dbeta * invM, k2 = dgamma * invM;
    const float a = p.gamma[c] * is;
    s_a[threadIdx.x] = a;                                                // dx = a*dz + b*x + c
    s_b[threadIdx.x] = -a * is * k2;
    s_c[threadIdx.x] = -a * (k1 - mean * is * k2);
  }
  __syncthreads();
  float a[8], b[8], c[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    a[e] = s_a[tx * 8 + e];
    b[e] = s_b[tx * 8 + e];
    c[e] = s_c[tx * 8 + e];
  }
  const bool want_dres = p.dres != nullptr;
  for (long long r = ty; r < nrows; r += BN_RPP) {
    const uint4 dzp = sdz[r * 8 + tx];
    const V8 d = unpack_bf16x8(dzp);
    V8 x = unpack_bf16x8(sx[r * 8 + tx]);
#pragma unroll
    for (int e = 0; e < 8; ++e) x.v[e] = fmaf(a[e], d.v[e], fmaf(b[e], x.v[e], c[e]));
    stg16(p.dx + (row0 + r) * p.C + coff, pack_bf16x8(x));
    if (want_dres) stg16(p.dres + (row0 + r) * p.C + coff, dzp);
  }
}

// Host-side plan shared by the launcher and the CPU test: cluster size and rows per CTA for M rows, or cs = 0 when the
// slice does not fit.  The kernels recompute rows_per with the same formula from cluster.num_blocks().
void bn_cluster_plan(long long M, int which, int* cs_out, long long* rows_per_out, size_t* smem_out) {
  const int slabs = (which == 0) ? 1 : 2;
  const long long cap = (which == 0) ? BNC_FWD_ROWS : BNC_BWD_ROWS;
  auto rows_per = [&](int cs) { return ((M + cs - 1) / cs + BN_RPP - 1) / BN_RPP * BN_RPP; };
  int cs = 1;
  while (cs < 8 && (M + cs - 1) / cs > 256) cs *= 2;    // ~256 rows per CTA: spread a slice over up to 8 SMs
  while (cs < 16 && rows_per(cs) > cap) cs *= 2;
  if (rows_per(cs) > cap) cs = 0;
  *cs_out = cs;
  *rows_per_out = cs ? rows_per(cs) : 0;
  *smem_out = cs ? (size_t)rows_per(cs) * 128 * slabs : 0;
}

// -------------------------------------------------------------------------------------------
// which: 0 training forward, 2 backward.  *taken = 1 if the single-launch cluster kernel was launched,
// 0 if the shape does not fit (caller falls back to launch_bn_v2).
cudaError_t launch_bn_cluster(const BnParamsV2& pp, int which, cudaStream_t s, int* taken) {
  const BnParams& p = pp.b;
  *taken = 0;
  if (p.C % BN_SLICE != 0 || p.C / BN_SLICE > 32 || p.M < 1) return cudaErrorInvalidValue;
  if (which != 0 && which != 2) return cudaErrorInvalidValue;
  if (p.relu && pp.mask == nullptr) return cudaErrorInvalidValue;
  static bool attr_done = false;
  static int max16_fwd = -1, max16_bwd = -1;            // is a 16-CTA cluster with full shared memory schedulable?
  int cs = 0;
  long long rows_per_cta = 0;
  size_t smem_bytes = 0;
  bn_cluster_plan(p.M, which, &cs, &rows_per_cta, &smem_bytes);
  if (cs == 0) return cudaSuccess;                      // does not fit: not taken
  if (!attr_done) {
    const int maxdyn = BNC_FWD_ROWS * 128;              // == BNC_BWD_ROWS * 256
    cudaError_t e = cudaFuncSetAttribute(bn_cluster_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxdyn);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(bn_cluster_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, maxdyn);
    if (e != cudaSuccess) return e;
    cudaFuncSetAttribute(bn_cluster_fwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    cudaFuncSetAttribute(bn_cluster_bwd_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    attr_done = true;
  }
  const int slices = p.C / BN_SLICE;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(slices * cs));
  cfg.blockDim = dim3(BN_THREADS);
  cfg.dynamicSmemBytes = smem_bytes;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeClusterDimension;
  at[0].val.clusterDim.x = (unsigned)cs;
  at[0].val.clusterDim.y = 1;
  at[0].val.clusterDim.z = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  if (cs == 16) {                                        // non-portable size: ask once whether it can be scheduled
    int& cached = (which == 0) ? max16_fwd : max16_bwd;
    if (cached < 0) {
      int n = 0;
      cudaLaunchConfig_t q = cfg;
      q.dynamicSmemBytes = (size_t)BNC_FWD_ROWS * 128;
      const cudaError_t e = (which == 0) ? cudaOccupancyMaxActiveClusters(&n, bn_cluster_fwd_kernel, &q)
                                         : cudaOccupancyMaxActiveClusters(&n, bn_cluster_bwd_kernel, &q);
      cached = (e == cudaSuccess && n > 0) ? 1 : 0;
      (void)cudaGetLastError();
    }
    if (!cached) return cudaSuccess;                     // not taken
  }
  const cudaError_t e = (which == 0) ? cudaLaunchKernelEx(&cfg, bn_cluster_fwd_kernel, pp)
                                     : cudaLaunchKernelEx(&cfg, bn_cluster_bwd_kernel, pp);
  if (e != cudaSuccess) return e;
  *taken = 1;
  return cudaGetLastError();
}

}  // namespace egb
