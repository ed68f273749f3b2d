// eventgrad_b200 -- receive side of the sparse (top-k) exchange, shared by csrc/sparse.cu (stand-alone launch) and
// csrc/gossip.cu (prologue of the mix+SGD kernel, so that a spevent step needs no separate apply launch).
#pragma once
#include "api.h"
#include "common.cuh"

namespace egb {

#define SP_BINS 2048
#define SP_MAX_TENSORS 1024     // == EG_MAX_OWN (asserted host-side)

// Grid-wide barrier of a persistent, co-resident grid (sense reversal on two words: arrivals, generation).
// Returns true on the LAST arriving CTA (before anybody is released) so it can do per-step bookkeeping.
__device__ __forceinline__ bool grid_barrier_arrive(unsigned int* bar, unsigned int* gen_out) {
  __shared__ int s_lastb;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned gen = *reinterpret_cast<volatile unsigned int*>(bar + 1);   // read BEFORE arriving
    *gen_out = gen;
    s_lastb = (atomicAdd(bar, 1u) == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  return s_lastb != 0;
}
__device__ __forceinline__ void grid_barrier_release(unsigned int* bar) {   // last arriver only, one thread
  bar[0] = 0u;
  __threadfence();
  atomicAdd(bar + 1, 1u);
}
__device__ __forceinline__ void grid_barrier_wait(unsigned int* bar, unsigned gen, int* status, uint64_t timeout_ns) {
  if (threadIdx.x == 0) {
    const uint64_t t0 = globaltimer_ns();
    while (*reinterpret_cast<volatile unsigned int*>(bar + 1) == gen) {
      __nanosleep(64);
      if (globaltimer_ns() - t0 > timeout_ns) {       // co-residency violated: never hang the GPU
        atomicExch(status, EG_ERR_TIMEOUT);
        break;
      }
    }
    __threadfence();
  }
  __syncthreads();
}

// Scatter the records that arrived since the last apply into the neighbour replicas (spevent.cpp:438-448, :492-502),
// then meet at a grid barrier (the mix that follows reads the replicas at arbitrary tiles); the last CTA to arrive
// records what was applied and acknowledges the records to the senders.
__device__ __forceinline__ void sparse_apply_prologue(const SparseParams& p, unsigned int* bar) {
  __shared__ unsigned char s_new[SP_MAX_TENSORS];
  __shared__ unsigned int s_gen;
  const int tid = threadIdx.x;
  const int step = *p.pass_num + 1;
  if (p.sync) {
    if (tid == 0) {
      wait_ge(p.done_from_l, (uint32_t)step, p.status, p.timeout_ns);
      wait_ge(p.done_from_r, (uint32_t)step, p.status, p.timeout_ns);
    }
    __syncthreads();
  }
  // which records need applying?  iter-sync: only those rewritten since the last apply (values are stable after the
  // done-flag wait).  async: every record that has ever been written -- exactly the reference, which re-scatters
  // whatever the window holds on every step (idempotent).
  const int sz = p.tab.n_tensors;
  for (int i = tid; i < sz; i += EG_THREADS) {
    const uint32_t sl = ld_acquire_sys(p.seq_from_l + i), sr = ld_acquire_sys(p.seq_from_r + i);
    const bool nl = p.sync ? (sl > p.applied_l[i]) : (sl > 0u);
    const bool nr = p.sync ? (sr > p.applied_r[i]) : (sr > 0u);
    s_new[i] = (unsigned char)((nl ? 1 : 0) | (nr ? 2 : 0));
  }
  // record start (in records, not words) of every tensor -> shared memory, so a thread can map a flat record number
  // to its tensor with a 10-step binary search; ALL records of ALL tensors are then spread evenly over the grid
  // (a per-tensor loop left most CTAs idle and serialised ~86 dependent global round trips on CTA 0)
  __shared__ int s_roff[SP_MAX_TENSORS + 1];
  for (int i = tid; i < sz; i += EG_THREADS) s_roff[i] = p.t_rec_off[i] >> 1;
  if (tid == 0) s_roff[sz] = (p.t_rec_off[sz - 1] >> 1) + p.t_k[sz - 1];
  __syncthreads();
  const int K = s_roff[sz];
  for (int r = blockIdx.x * EG_THREADS + tid; r < K; r += gridDim.x * EG_THREADS) {
    int lo = 0, hi = sz - 1;                             // last tensor with s_roff[i] <= r
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (s_roff[mid] <= r) lo = mid; else hi = mid - 1;
    }
    const int i = lo;
    const unsigned char nw = s_new[i];
    if (!nw) continue;
    const int c = r - s_roff[i];
    const int k = p.t_k[i];
    const int numel = p.tab.t_numel[i];
    const size_t ro = (size_t)p.t_rec_off[i];
    const size_t toff = (size_t)p.tab.t_tile_start[i] * EG_TILE;
    if (nw & 1) {
      const float v = __ldcg(p.rec_from_l + ro + c);
      const int idx = __float_as_int(__ldcg(p.rec_from_l + ro + k + c));
      if (idx >= 0 && idx < numel) p.rep_l[toff + idx] = v;      // spevent.cpp:438-448
    }
    if (nw & 2) {
      const float v = __ldcg(p.rec_from_r + ro + c);
      const int idx = __float_as_int(__ldcg(p.rec_from_r + ro + k + c));
      if (idx >= 0 && idx < numel) p.rep_r[toff + idx] = v;      // spevent.cpp:492-502
    }
  }
  const bool last = grid_barrier_arrive(bar, &s_gen);
  if (last) {                                            // every CTA has finished scattering
    if (p.sync) {
      for (int i = tid; i < sz; i += EG_THREADS) {
        p.applied_l[i] = ld_acquire_sys(p.seq_from_l + i);
        p.applied_r[i] = ld_acquire_sys(p.seq_from_r + i);
      }
    }
    __syncthreads();
    if (tid == 0) {
      if (p.sync) {
        fence_sys();
        st_release_sys(p.ack_to_l, (uint32_t)step);     // records of `step` consumed
        st_release_sys(p.ack_to_r, (uint32_t)step);
      }
      grid_barrier_release(bar);
    }
    __syncthreads();
  } else {
    grid_barrier_wait(bar, s_gen, p.status, p.timeout_ns);
  }
}

}  // namespace egb
