// eventgrad_b200 -- sparse (top-k) event exchange (K3) for sm_100a.
//
// spevent semantics (/root/reference/dcifar10/spevent/spevent.cpp:342-448): when tensor i
// fires, the k_i elements with the largest |theta - prev| travel as (value, index) pairs; the
// sender then sets prev[idx] = value; the receiver scatters the record into a persistent
// replica of that neighbour and mixes with the FULL replicas.
//
// Device design (all fired tensors of the model in one batched, segmented pass -- no per-tensor
// launches, no host round trip, nothing sorted):
//   1. exact k-th largest key per tensor by a 3-digit (11/11/10 bit) MSD radix select on the
//      monotone uint32 image of |diff|; histograms are segmented by tensor (a tile belongs to
//      exactly one tensor) and accumulated in shared memory;
//   2. per-tile (> tau, == tau) counts + a per-tensor exclusive scan give every selected
//      element a deterministic slot (ties at tau resolved towards the lowest index);
//   3. compaction writes (value, int32 index) straight into BOTH neighbours' inbox records
//      over NVLink and updates prev in the same pass;
//   4. sparse_apply scatters the freshly arrived records into the replicas (skipping records
//      that did not change), after which the dense mix+SGD kernel (gossip.cu) runs on
//      (theta, rep_l, rep_r).
#include "api.h"
#include "common.cuh"

namespace egb {

#define SP_BINS 2048
#define SP_MAX_TENSORS 4096

__device__ __forceinline__ uint32_t diff_key(float a, float b) {
  return __float_as_uint(fabsf(__fsub_rn(a, b)));   // non-negative floats order like uints
}

// number of valid (non-padding) elements of tile t
__device__ __forceinline__ int tile_valid(const TableDev& tab, int t, int i) {
  const int first = (t - tab.t_tile_start[i]) * EG_TILE;
  const int rem = tab.t_numel[i] - first;
  return rem < EG_TILE ? rem : EG_TILE;
}

// ---------------------------------------------------------------- 1. radix-select histograms
// pass 0: digit = key[31:21]; pass 1: key[20:10] among keys matching prefix; pass 2: key[9:0].
template <int PASS>
__global__ void __launch_bounds__(EG_THREADS, 4) sparse_hist_kernel(const SparseParams p) {
  __shared__ unsigned int sh[SP_BINS];
  const int tid = threadIdx.x;
  const int G = gridDim.x;
  const int per = (p.tab.n_tiles + G - 1) / G;          // blocked tile ranges: few tensor switches
  const int t0 = blockIdx.x * per, t1 = min(p.tab.n_tiles, t0 + per);
  int cur = -1;
  auto flush = [&](int tensor) {
    __syncthreads();
    if (tensor >= 0) {
      for (int bkt = tid; bkt < SP_BINS; bkt += EG_THREADS) {
        const unsigned c = sh[bkt];
        if (c) atomicAdd(p.hist + (size_t)tensor * SP_BINS + bkt, c);
      }
    }
    for (int bkt = tid; bkt < SP_BINS; bkt += EG_THREADS) sh[bkt] = 0u;
    __syncthreads();
  };
  flush(-1);
  for (int t = t0; t < t1; ++t) {
    const int i = p.tab.tile_tensor[t];
    if (!p.fire[i]) continue;
    if (i != cur) {
      flush(cur);
      cur = i;
    }
    const int valid = tile_valid(p.tab, t, i);
    const size_t base = (size_t)t * EG_TILE + (size_t)tid * EG_VEC;
    const F8 a = ld_f8(p.theta + base), b = ld_f8(p.prev + base);
    uint32_t prefix = 0;
    if (PASS > 0) prefix = p.sel_prefix[i];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (tid * EG_VEC + e >= valid) continue;
      const uint32_t key = diff_key(a.v[e], b.v[e]);
      if (PASS == 0) {
        atomicAdd(&sh[key >> 21], 1u);
      } else if (PASS == 1) {
        if ((key >> 21) == prefix) atomicAdd(&sh[(key >> 10) & 0x7FFu], 1u);
      } else {
        if ((key >> 10) == prefix) atomicAdd(&sh[key & 0x3FFu], 1u);
      }
    }
  }
  flush(cur);
}

// One CTA per tensor: walk the histogram from the top bucket down to the one holding the
// element of rank `remain` (1-based, counted from the largest); extend the prefix.
template <int PASS>
__global__ void __launch_bounds__(EG_THREADS) sparse_scan_kernel(const SparseParams p) {
  const int i = blockIdx.x;
  if (!p.fire[i]) return;
  __shared__ unsigned int wsum[EG_WARPS];
  __shared__ unsigned int s_digit, s_above;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned int* h = p.hist + (size_t)i * SP_BINS;
  const int nb = (PASS == 2) ? 1024 : SP_BINS;
  // descending order: position q <-> bucket nb-1-q ; thread owns 8 consecutive positions
  unsigned int loc[8], tot = 0;
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int q = tid * 8 + e;
    const unsigned c = (q < nb) ? h[nb - 1 - q] : 0u;
    loc[e] = c;
    tot += c;
    if (q < nb) h[nb - 1 - q] = 0u;            // ready for the next pass / next step
  }
  // block exclusive scan of tot
  unsigned int inc = tot;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const unsigned v = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += v;
  }
  if (lane == 31) wsum[warp] = inc;
  __syncthreads();
  unsigned int woff = 0;
  for (int w = 0; w < warp; ++w) woff += wsum[w];
  unsigned int excl = woff + inc - tot;
  const unsigned int remain = (PASS == 0) ? (unsigned)p.t_k[i] : p.sel_remain[i];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    if (remain > excl && remain <= excl + loc[e]) {       // exactly one (thread, e) satisfies this
      s_digit = (unsigned)(nb - 1 - (tid * 8 + e));
      s_above = excl;
    }
    excl += loc[e];
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t prev = (PASS == 0) ? 0u : p.sel_prefix[i];
    const int bits = (PASS == 2) ? 10 : 11;
    p.sel_prefix[i] = (prev << bits) | s_digit;
    p.sel_remain[i] = remain - s_above;                   // rank inside the chosen bucket
  }
}

// ---------------------------------------------------------------- 2. per-tile counts + scan
__global__ void __launch_bounds__(EG_THREADS, 4) sparse_count_kernel(const SparseParams p) {
  __shared__ unsigned int wg[EG_WARPS], we[EG_WARPS];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int t = blockIdx.x; t < p.tab.n_tiles; t += gridDim.x) {
    const int i = p.tab.tile_tensor[t];
    if (!p.fire[i]) continue;
    const uint32_t tau = p.sel_prefix[i];
    const int valid = tile_valid(p.tab, t, i);
    const size_t base = (size_t)t * EG_TILE + (size_t)tid * EG_VEC;
    const F8 a = ld_f8(p.theta + base), b = ld_f8(p.prev + base);
    unsigned g = 0, q = 0;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (tid * EG_VEC + e >= valid) continue;
      const uint32_t key = diff_key(a.v[e], b.v[e]);
      g += key > tau;
      q += key == tau;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      g += __shfl_xor_sync(0xffffffffu, g, o);
      q += __shfl_xor_sync(0xffffffffu, q, o);
    }
    if (lane == 0) {
      wg[warp] = g;
      we[warp] = q;
    }
    __syncthreads();
    if (tid == 0) {
      unsigned G2 = 0, Q2 = 0;
      for (int w = 0; w < EG_WARPS; ++w) {
        G2 += wg[w];
        Q2 += we[w];
      }
      p.tile_gt[t] = G2;
      p.tile_eq[t] = Q2;
    }
    __syncthreads();
  }
}

// one warp per tensor: exclusive prefix of (gt, eq) over the tensor's tiles, in place
__global__ void __launch_bounds__(EG_THREADS) sparse_tilescan_kernel(const SparseParams p) {
  const int lane = threadIdx.x & 31;
  const int i = blockIdx.x * EG_WARPS + (threadIdx.x >> 5);
  if (i >= p.tab.n_tensors || !p.fire[i]) return;
  const int ts = p.tab.t_tile_start[i], tc = p.tab.t_tile_count[i];
  unsigned cg = 0, ce = 0;
  for (int s = 0; s < tc; s += 32) {
    const int t = ts + s + lane;
    unsigned g = (s + lane < tc) ? p.tile_gt[t] : 0u, q = (s + lane < tc) ? p.tile_eq[t] : 0u;
    unsigned gi = g, qi = q;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned vg = __shfl_up_sync(0xffffffffu, gi, o), vq = __shfl_up_sync(0xffffffffu, qi, o);
      if (lane >= o) {
        gi += vg;
        qi += vq;
      }
    }
    if (s + lane < tc) {
      p.tile_gt[t] = cg + gi - g;
      p.tile_eq[t] = ce + qi - q;
    }
    cg += __shfl_sync(0xffffffffu, gi, 31);
    ce += __shfl_sync(0xffffffffu, qi, 31);
  }
  if (lane == 0) p.t_gt_total[i] = cg;
}

// ---------------------------------------------------------------- 3. compaction -> peers
__global__ void __launch_bounds__(EG_THREADS, 4) sparse_compact_kernel(const SparseParams p) {
  __shared__ unsigned int wg[EG_WARPS], we[EG_WARPS];
  __shared__ int s_last;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int step = *p.pass_num + 1;
  if (p.sync) {
    if (tid == 0) {   // WAR guard on the neighbours' record inboxes
      wait_ge(p.ack_from_l, (uint32_t)(step - 1), p.status, p.timeout_ns);
      wait_ge(p.ack_from_r, (uint32_t)(step - 1), p.status, p.timeout_ns);
    }
    __syncthreads();
  }
  for (int t = blockIdx.x; t < p.tab.n_tiles; t += gridDim.x) {
    const int i = p.tab.tile_tensor[t];
    if (!p.fire[i]) continue;
    const uint32_t tau = p.sel_prefix[i];
    const unsigned need_eq = p.sel_remain[i];
    const unsigned gt_total = p.t_gt_total[i];
    const int k = p.t_k[i];
    const int valid = tile_valid(p.tab, t, i);
    const int first = (t - p.tab.t_tile_start[i]) * EG_TILE;
    const size_t base = (size_t)t * EG_TILE + (size_t)tid * EG_VEC;
    const F8 a = ld_f8(p.theta + base);
    F8 b = ld_f8(p.prev + base);
    unsigned fg = 0, fe = 0, g = 0, q = 0;   // bit masks + counts
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      if (tid * EG_VEC + e >= valid) continue;
      const uint32_t key = diff_key(a.v[e], b.v[e]);
      if (key > tau) {
        fg |= 1u << e;
        ++g;
      } else if (key == tau) {
        fe |= 1u << e;
        ++q;
      }
    }
    unsigned gi = g, qi = q;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const unsigned vg = __shfl_up_sync(0xffffffffu, gi, o), vq = __shfl_up_sync(0xffffffffu, qi, o);
      if (lane >= o) {
        gi += vg;
        qi += vq;
      }
    }
    if (lane == 31) {
      wg[warp] = gi;
      we[warp] = qi;
    }
    __syncthreads();
    unsigned og = p.tile_gt[t], oe = p.tile_eq[t];
    for (int w = 0; w < warp; ++w) {
      og += wg[w];
      oe += we[w];
    }
    og += gi - g;
    oe += qi - q;
    const size_t ro = (size_t)p.t_rec_off[i];
    bool touched = false;
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      int pos = -1;
      if (fg & (1u << e)) {
        pos = (int)og++;
      } else if (fe & (1u << e)) {
        if (oe < need_eq) pos = (int)(gt_total + oe);
        ++oe;
      }
      if (pos >= 0 && pos < k) {
        const float val = a.v[e];
        const float idxw = __int_as_float(first + tid * EG_VEC + e);
        p.rec_to_l[ro + pos] = val;
        p.rec_to_l[ro + k + pos] = idxw;
        p.rec_to_r[ro + pos] = val;
        p.rec_to_r[ro + k + pos] = idxw;
        b.v[e] = val;                                   // prev[idx] <- value sent (spevent.cpp:407-413)
        touched = true;
      }
    }
    if (touched) st_f8(p.prev + base, b);
    __syncthreads();
  }
  // ---- publish: per-tensor sequence numbers (+ step-done flag in iter-sync mode) -------------
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    const unsigned prev = atomicAdd(p.ticket, 1u);
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence_system();
  for (int i = tid; i < p.tab.n_tensors; i += EG_THREADS) {
    if (p.fire[i]) {
      st_release_sys(p.seq_to_l + i, (uint32_t)step);
      st_release_sys(p.seq_to_r + i, (uint32_t)step);
    }
  }
  __syncthreads();
  if (tid == 0) {
    *p.ticket = 0u;
    fence_sys();
    st_release_sys(p.done_to_l, (uint32_t)step);
    st_release_sys(p.done_to_r, (uint32_t)step);
  }
}

// ---------------------------------------------------------------- 4. receive: scatter records
__global__ void __launch_bounds__(EG_THREADS, 4) sparse_apply_kernel(const SparseParams p) {
  __shared__ int s_last;
  const int tid = threadIdx.x;
  const int step = *p.pass_num + 1;
  if (p.sync) {
    if (tid == 0) {
      wait_ge(p.done_from_l, (uint32_t)step, p.status, p.timeout_ns);
      wait_ge(p.done_from_r, (uint32_t)step, p.status, p.timeout_ns);
    }
    __syncthreads();
  }
  // which records need applying?  iter-sync: only those rewritten since the last apply (values are
  // stable after the done-flag wait).  async: every record that has ever been written -- exactly
  // the reference, which re-scatters whatever the window holds on every step (idempotent).
  const int sz = p.tab.n_tensors;
  __shared__ unsigned char s_new[SP_MAX_TENSORS];
  for (int i = tid; i < sz; i += EG_THREADS) {
    const uint32_t sl = ld_acquire_sys(p.seq_from_l + i), sr = ld_acquire_sys(p.seq_from_r + i);
    const bool nl = p.sync ? (sl > p.applied_l[i]) : (sl > 0u);
    const bool nr = p.sync ? (sr > p.applied_r[i]) : (sr > 0u);
    s_new[i] = (unsigned char)((nl ? 1 : 0) | (nr ? 2 : 0));
  }
  __syncthreads();
  for (int i = 0; i < sz; ++i) {
    const bool newl = s_new[i] & 1, newr = s_new[i] & 2;
    if (!newl && !newr) continue;
    const int k = p.t_k[i];
    const int numel = p.tab.t_numel[i];
    const size_t ro = (size_t)p.t_rec_off[i];
    const size_t toff = (size_t)p.tab.t_tile_start[i] * EG_TILE;
    for (int c = blockIdx.x * EG_THREADS + tid; c < k; c += gridDim.x * EG_THREADS) {
      if (newl) {
        const float v = __ldcg(p.rec_from_l + ro + c);
        const int idx = __float_as_int(__ldcg(p.rec_from_l + ro + k + c));
        if (idx >= 0 && idx < numel) p.rep_l[toff + idx] = v;      // spevent.cpp:438-448
      }
      if (newr) {
        const float v = __ldcg(p.rec_from_r + ro + c);
        const int idx = __float_as_int(__ldcg(p.rec_from_r + ro + k + c));
        if (idx >= 0 && idx < numel) p.rep_r[toff + idx] = v;      // spevent.cpp:492-502
      }
    }
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(p.ticket, 1u);
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  if (p.sync) {
    for (int i = tid; i < sz; i += EG_THREADS) {
      p.applied_l[i] = ld_acquire_sys(p.seq_from_l + i);
      p.applied_r[i] = ld_acquire_sys(p.seq_from_r + i);
    }
  }
  __syncthreads();
  if (tid == 0) {
    *p.ticket = 0u;
    if (p.sync) {
      fence_sys();
      st_release_sys(p.ack_to_l, (uint32_t)step);   // records of `step` consumed
      st_release_sys(p.ack_to_r, (uint32_t)step);
    }
  }
}

// ------------------------------------------------------------------------------------------
cudaError_t launch_sparse_select_push(const SparseParams& p, int grid, cudaStream_t s) {
  const int sz = p.tab.n_tensors;
  eg_count_launch(EG_FAM_SPARSE, 9);
  sparse_hist_kernel<0><<<grid, EG_THREADS, 0, s>>>(p);
  sparse_scan_kernel<0><<<sz, EG_THREADS, 0, s>>>(p);
  sparse_hist_kernel<1><<<grid, EG_THREADS, 0, s>>>(p);
  sparse_scan_kernel<1><<<sz, EG_THREADS, 0, s>>>(p);
  sparse_hist_kernel<2><<<grid, EG_THREADS, 0, s>>>(p);
  sparse_scan_kernel<2><<<sz, EG_THREADS, 0, s>>>(p);
  sparse_count_kernel<<<grid, EG_THREADS, 0, s>>>(p);
  sparse_tilescan_kernel<<<(sz + EG_WARPS - 1) / EG_WARPS, EG_THREADS, 0, s>>>(p);
  sparse_compact_kernel<<<grid, EG_THREADS, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_sparse_apply(const SparseParams& p, int grid, cudaStream_t s) {
  eg_count_launch(EG_FAM_SPARSE, 1);
  sparse_apply_kernel<<<grid, EG_THREADS, 0, s>>>(p);
  return cudaGetLastError();
}

}  // namespace egb
