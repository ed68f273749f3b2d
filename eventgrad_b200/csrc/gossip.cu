// eventgrad_b200 -- fused event-triggered gossip step (K1/K2) for sm_100a.
//
// What one launch does on every rank (persistent grid, one tile of EG_TILE floats per CTA
// iteration):
//   1. push   : tiles of tensors whose trigger fired are stored straight into the LEFT
//               neighbour's "from-right" inbox and the RIGHT neighbour's "from-left" inbox
//               (peer-mapped pointers -> NVLink 5 / NVSwitch).  Below threshold nothing is
//               stored, so link bytes scale with events (MPI_Put semantics,
//               /root/reference/dcifar10/event/event.cpp:317-332).
//   2. sync   : iter mode publishes one release-flag per (tile, warp) and waits for the two
//               neighbours' matching flags, so exchange and math overlap slice by slice;
//               async mode reads whatever the inboxes hold (reference: unsynchronised window
//               reads, event.cpp:372-374).
//   3. mix+opt: theta <- ((theta+L)+R)/3 ; m <- mu*m+g ; theta <- theta - lr*m   (event.cpp:
//               459-461, :479), rounding identical to add_/add_/div_/SGD.
//   4. norm-on-write: per-warp sum theta_new^2 partials; the last CTA reduces them per tensor
//               in fixed order and runs the trigger FSM for the NEXT step (event.cpp:300-355),
//               so the step needs no extra pass over theta, no grid sync and no host sync.
//
// This file is compiled twice.  As gossip.cu it is the default kernel described above.  csrc/gossip_dbuf.cu
// defines EG_DBUF and includes it again to build the EXPERIMENTAL double-buffered dense variant
// (`gossip_step_kernel_dbuf`, decent only): the inbox has two slots, slot = step & 1, and the WAR ack
// (PROTOCOL.md 1.3) disappears -- a sender can be at most one step ahead of a receiver's reads because
// finishing step k+1 needs the receiver's flags of step k+1 (model-checked: tests/test_protocol_model.py).
#include "api.h"
#include "common.cuh"
#include "sparse_apply.cuh"

#ifdef EG_DBUF
#define EG_SYM(name) name##_dbuf
#define EG_SLOT_OFF(p, step) ((size_t)((step) & 1) * ((size_t)(p).tab.n_tiles * EG_TILE))
#else
#define EG_SYM(name) name
#define EG_SLOT_OFF(p, step) ((size_t)0)
#endif

namespace egb {

// Per-CTA tile metadata, resolved ONCE at kernel start (thread j resolves iteration j, so the
// dependent table look-ups tile->tensor->{fire, grad pointer, numel} run in parallel instead of
// sitting on every iteration's critical path in front of the streaming loads).
struct TileInfo {
  const void* gptr;   // table mode: address of this tile's first gradient element (null: flat arena)
  int tensor;
  int valid;          // elements of the tensor left from this tile's start (>= EG_TILE: full tile)
  int flags;          // bit0 fire, bit1 gradient is bf16
};
#define EG_TI_CACHE 64

// -------------------------------------------------------------------------------------------
// Trigger FSM update of ONE tensor for step `next_step` (scalar code, one lane).
// Also accounts the messages of the step that just ran (fire[i] still holds that decision).
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void fsm_update_tensor(const FsmDev& f, const TableDev& tab, int i, float norm,
                                                  float lnorm, float rnorm, bool have_recv, int next_step,
                                                  bool count) {
  const int sz = tab.n_tensors;
  float* row_next = nullptr;   // log row of the decision (step next_step)
  float* row_cur = nullptr;    // log row of the step that just ran (receive-side norms)
  if (f.log_ring != nullptr && f.log_cap > 0) {
    row_next = f.log_ring + (size_t)((next_step - 1) % f.log_cap) * sz * 5;
    if (next_step >= 2) row_cur = f.log_ring + (size_t)((next_step - 2) % f.log_cap) * sz * 5;
  }
  if (row_cur != nullptr && have_recv) {
    row_cur[i * 5 + 3] = lnorm;
    row_cur[i * 5 + 4] = rnorm;
  }
  f.cur_norm[i] = norm;
  if (!f.enabled) return;
  if (count && f.fire[i]) {   // +2 events per fired tensor, one per ring neighbour (event.cpp:319)
    atomicAdd(f.counters + 0, 2ull);
    atomicAdd(f.counters + 1, 2ull * (unsigned long long)tab.t_msg_bytes[i]);
    atomicAdd(f.counters + 2, 1ull);
  }
  const float value_diff = fabsf(__fsub_rn(norm, f.last_norm[i]));
  const float iter_diff = __fsub_rn((float)next_step, f.last_iter[i]);
  float th = (f.thres_type == 1) ? __fmul_rn(f.thres[i], f.horizon) : f.constant;
  const bool fire = (value_diff >= th) || (next_step < f.initial_comm_passes);
  if (row_next != nullptr) {
    row_next[i * 5 + 0] = norm;
    row_next[i * 5 + 1] = th;
    row_next[i * 5 + 2] = fire ? 1.f : 0.f;
  }
  if (fire) {
    const int H = f.history;
    float* sl = f.slopes + (size_t)i * H;
    double avg = 0.0;
    for (int j = 0; j < H - 1; ++j) {
      sl[j] = sl[j + 1];
      avg += (double)sl[j];
    }
    sl[H - 1] = __fdiv_rn(value_diff, iter_diff);
    avg += (double)sl[H - 1];
    avg /= (double)H;
    if (f.thres_type == 1) th = (float)avg;
    f.last_norm[i] = norm;
    f.last_iter[i] = (float)next_step;
  }
  f.thres[i] = th;
  f.fire[i] = fire ? 1 : 0;
}

// Fixed-order reduction of the per-warp partials of one tensor by ONE warp: every lane owns a
// strided subsequence (float4 loads, 4 independent accumulators), then a shuffle tree.
__device__ __forceinline__ double reduce_partials(const float* part, int n4, int lane) {
  const float4* p4 = reinterpret_cast<const float4*>(part);
  double a0 = 0.0, a1 = 0.0, a2 = 0.0, a3 = 0.0;
  int j = lane;
  for (; j + 96 < n4; j += 128) {
    const float4 x0 = __ldcg(p4 + j), x1 = __ldcg(p4 + j + 32), x2 = __ldcg(p4 + j + 64), x3 = __ldcg(p4 + j + 96);
    a0 += ((double)x0.x + (double)x0.y) + ((double)x0.z + (double)x0.w);
    a1 += ((double)x1.x + (double)x1.y) + ((double)x1.z + (double)x1.w);
    a2 += ((double)x2.x + (double)x2.y) + ((double)x2.z + (double)x2.w);
    a3 += ((double)x3.x + (double)x3.y) + ((double)x3.z + (double)x3.w);
  }
  for (; j < n4; j += 32) {
    const float4 x0 = __ldcg(p4 + j);
    a0 += ((double)x0.x + (double)x0.y) + ((double)x0.z + (double)x0.w);
  }
  return warp_sum_d((a0 + a1) + (a2 + a3));
}

// Norm reduction + trigger, spread over the grid with no serial tail and no per-tile fences:
// every CTA, once it has finished ALL its tiles, adds the number of tiles it contributed to each
// tensor's completion counter (one fence per CTA, <= tiles-per-CTA atomics).  The CTA whose add
// completes a tensor reduces that tensor's per-warp partials in fixed order (one warp per tensor)
// and runs the trigger FSM for it.  Tiles of a CTA ascend, so equal tensors are consecutive.
#define EG_MAX_OWN 1024   // >= number of parameter tensors (asserted host-side)
__device__ __forceinline__ void cta_finish_tensors(const GossipParams& p, const TileInfo* s_ti, int next_step,
                                                   bool count, bool recv_ok) {
  __shared__ int s_own[EG_MAX_OWN];
  __shared__ int s_nown;
  const int b = blockIdx.x, G = gridDim.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  __syncthreads();                       // every warp of this CTA has stored its partials
  if (tid == 0) {
    __threadfence();
    int nown = 0, cur = -1, cnt = 0;
    int j = 0;
    for (int t = b; ; t += G, ++j) {
      const int i = (t < p.tab.n_tiles) ? ((s_ti != nullptr && j < EG_TI_CACHE) ? s_ti[j].tensor : p.tab.tile_tensor[t]) : -2;
      if (i != cur) {
        if (cur >= 0) {
          const unsigned prev = atomicAdd(p.tensor_done + cur, (unsigned)cnt);
          if (prev + (unsigned)cnt == (unsigned)p.tab.t_tile_count[cur] && nown < EG_MAX_OWN) s_own[nown++] = cur;
        }
        cur = i;
        cnt = 0;
      }
      if (i == -2) break;
      ++cnt;
    }
    s_nown = nown;
  }
  __syncthreads();
  const int nown = s_nown;
  if (nown == 0) return;
  __threadfence();
  const bool recv = recv_ok && (p.tile_ss_l != nullptr);
  for (int k = warp; k < nown; k += EG_WARPS) {
    const int i = s_own[k];
    const size_t off = (size_t)p.tab.t_tile_start[i] * EG_WARPS;
    const int n4 = p.tab.t_tile_count[i] * EG_WARPS / 4;
    const double ss = reduce_partials(p.tile_ss + off, n4, lane);
    double sl = 0.0, sr = 0.0;
    if (recv) {
      sl = reduce_partials(p.tile_ss_l + off, n4, lane);
      sr = reduce_partials(p.tile_ss_r + off, n4, lane);
    }
    if (lane == 0) {
      p.tensor_done[i] = 0u;
      fsm_update_tensor(p.fsm, p.tab, i, (float)sqrt(ss), (float)sqrt(sl), (float)sqrt(sr), recv, next_step, count);
    }
  }
}

// -------------------------------------------------------------------------------------------
__device__ __forceinline__ void push_tile(const GossipParams& p, size_t base, const F8& th) {
  // push_r == nullptr: 2-rank ring, the one neighbour is both left and right and reads the single copy twice
  if (p.vec256_push) {
    st_f8(p.push_l + base, th);
    if (p.push_r != nullptr) st_f8(p.push_r + base, th);
  } else {
    st_f8_v4(p.push_l + base, th);
    if (p.push_r != nullptr) st_f8_v4(p.push_r + base, th);
  }
}
#ifdef EG_DBUF
__device__ __forceinline__ void push_tile_slot(const GossipParams& p, size_t base, const F8& th, int step) {
  push_tile(p, EG_SLOT_OFF(p, step) + base, th);
}
#endif

__device__ __forceinline__ TileInfo resolve_tile(const GossipParams& p, int t) {
  TileInfo ti;
  const int i = p.tab.tile_tensor[t];
  const int first = (t - p.tab.t_tile_start[i]) * EG_TILE;
  ti.tensor = i;
  ti.valid = p.tab.t_numel[i] - first;
  ti.flags = p.fsm.fire[i] ? 1 : 0;
  ti.gptr = nullptr;
  if (p.t_grad_ptr != nullptr) {
    const bool bf = p.t_grad_bf16[i] != 0;
    ti.flags |= bf ? 2 : 0;
    ti.gptr = reinterpret_cast<const char*>(p.t_grad_ptr[i]) + (size_t)first * (bf ? 2 : 4);
  }
  return ti;
}

__device__ __forceinline__ void fill_tile_cache(const GossipParams& p, TileInfo* s_ti) {
  const int j = threadIdx.x;
  if (j < EG_TI_CACHE) {
    const int t = blockIdx.x + j * gridDim.x;
    if (t < p.tab.n_tiles) s_ti[j] = resolve_tile(p, t);
  }
  __syncthreads();
}
__device__ __forceinline__ TileInfo tile_info(const GossipParams& p, const TileInfo* s_ti, int j, int t) {
  return (j < EG_TI_CACHE) ? s_ti[j] : resolve_tile(p, t);
}

// Gradient of this thread's 8 elements.  Flat mode: the fp32 grad arena.  Table mode: autograd's
// own gradient tensors are read in place (bf16 or fp32, exactly numel elements each) -- no
// AccumulateGrad kernels, no grad arena traffic, no zeroing.
__device__ __forceinline__ F8 load_grad(const GossipParams& p, const TileInfo& ti, size_t base, int tid) {
  if (ti.gptr == nullptr) return ld_f8(p.grad + base);
  const int off = tid * EG_VEC;
  F8 g;
  if (ti.flags & 2) {
    const __nv_bfloat16* gp = reinterpret_cast<const __nv_bfloat16*>(ti.gptr);
    if (off + EG_VEC <= ti.valid) {
      const uint4 u = *reinterpret_cast<const uint4*>(gp + off);
      const __nv_bfloat162* h = reinterpret_cast<const __nv_bfloat162*>(&u);
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float2 f = __bfloat1622float2(h[e]);
        g.v[2 * e] = f.x;
        g.v[2 * e + 1] = f.y;
      }
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) g.v[e] = (off + e < ti.valid) ? __bfloat162float(gp[off + e]) : 0.f;
    }
  } else {
    const float* gp = reinterpret_cast<const float*>(ti.gptr);
    if (off + EG_VEC <= ti.valid) {
      const float4 a = *reinterpret_cast<const float4*>(gp + off), b = *reinterpret_cast<const float4*>(gp + off + 4);
      g.v[0] = a.x; g.v[1] = a.y; g.v[2] = a.z; g.v[3] = a.w;
      g.v[4] = b.x; g.v[5] = b.y; g.v[6] = b.z; g.v[7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) g.v[e] = (off + e < ti.valid) ? gp[off + e] : 0.f;
    }
  }
  return g;
}

// mix + SGD + norm-on-write for one tile. `th` already holds theta_k for this thread's 8 floats.
template <bool kMom>
__device__ __forceinline__ void mix_tile(const GossipParams& p, int t, const TileInfo& ti, size_t base, F8 th,
                                         int lane, int warp, int step) {
  const bool logrecv = (p.tile_ss_l != nullptr);
  float ssl = 0.f, ssr = 0.f;
  // issue every load of the tile before the first dependent FP op (5 x 32 B in flight per lane)
  F8 L, R, m;
  if (p.do_mix) {
#ifdef EG_DBUF
    L = ld_f8_cg(p.inbox_l + EG_SLOT_OFF(p, step) + base);
    R = ld_f8_cg(p.inbox_r + EG_SLOT_OFF(p, step) + base);
#else
    L = ld_f8_cg(p.inbox_l + base);
    R = ld_f8_cg(p.inbox_r + base);
#endif
  }
  const F8 g = load_grad(p, ti, base, threadIdx.x);
  if (kMom) m = ld_f8(p.mom + base);
  if (p.do_mix) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      th.v[e] = __fdiv_rn(__fadd_rn(__fadd_rn(th.v[e], L.v[e]), R.v[e]), 3.0f);
      if (logrecv) {
        ssl = __fmaf_rn(L.v[e], L.v[e], ssl);
        ssr = __fmaf_rn(R.v[e], R.v[e], ssr);
      }
    }
  }
  float ss = 0.f;
  if (kMom) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      m.v[e] = __fadd_rn(__fmul_rn(m.v[e], p.mu), g.v[e]);
      th.v[e] = __fmaf_rn(m.v[e], -p.lr, th.v[e]);
      ss = __fmaf_rn(th.v[e], th.v[e], ss);
    }
    st_f8(p.mom + base, m);
  } else {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      th.v[e] = __fmaf_rn(g.v[e], -p.lr, th.v[e]);
      ss = __fmaf_rn(th.v[e], th.v[e], ss);
    }
  }
  st_f8(p.theta + base, th);
  if (p.shadow != nullptr) st_bf16x8(p.shadow + base, th);
  if (p.zero_grad && ti.gptr == nullptr) {
    F8 z;
#pragma unroll
    for (int e = 0; e < 8; ++e) z.v[e] = 0.f;
    st_f8(p.grad + base, z);
  }
  ss = warp_sum(ss);
  if (logrecv) {
    ssl = warp_sum(ssl);
    ssr = warp_sum(ssr);
  }
  if (!p.need_norm) return;
  (void)step;
  if (lane == 0) {
    p.tile_ss[(size_t)t * EG_WARPS + warp] = ss;
    if (logrecv) {
      p.tile_ss_l[(size_t)t * EG_WARPS + warp] = ssl;
      p.tile_ss_r[(size_t)t * EG_WARPS + warp] = ssr;
    }
  }
}

// Elect the last CTA of the grid; it acks and bumps the step counter (the trigger FSM itself is
// run per tensor by whichever warp completes that tensor, see cta_finish_tensors).
__device__ __forceinline__ void grid_tail(const GossipParams& p, int step) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(p.ticket, 1u);
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (!s_last) return;
  if (threadIdx.x == 0) {
    *p.ticket = 0u;
    *p.fsm.pass_num = step;
#ifndef EG_DBUF
    if (p.sync && p.send_ack) {
      // every CTA of this rank has finished reading its inboxes for `step`
      fence_sys();
      st_release_sys(p.ack_to_l, (uint32_t)step);
      st_release_sys(p.ack_to_r, (uint32_t)step);
    }
#endif
  }
}

// phase 1 of the split step: push only.  Depends on theta_k and the trigger decisions alone, so
// it is launched on a side stream at the START of step k and its NVLink traffic hides behind the
// forward/backward pass.  Publishes ONE "all pushed" flag per neighbour when every CTA is done.
__device__ __forceinline__ void push_phase(const GossipParams& p, int step) {
  __shared__ int s_last;
  const int b = blockIdx.x, G = gridDim.x, tid = threadIdx.x;
  if (p.sync) {
    __shared__ int s_ok;
    if (tid == 0) {   // WAR guard: neighbours have consumed what I pushed at step-1
      bool ok = wait_ge(p.ack_from_l, (uint32_t)(step - 1), p.status, p.timeout_ns);
      ok = wait_ge(p.ack_from_r, (uint32_t)(step - 1), p.status, p.timeout_ns) && ok;
      s_ok = ok ? 1 : 0;
    }
    __syncthreads();
    if (!s_ok) return;   // a wedged / dead peer (sticky status set): never overwrite memory it may still be reading
  }
  __shared__ TileInfo s_ti[EG_TI_CACHE];
  fill_tile_cache(p, s_ti);
  int j = 0;
  for (int t = b; t < p.tab.n_tiles; t += G, ++j) {
    if (!(tile_info(p, s_ti, j, t).flags & 1)) continue;
    const size_t base = (size_t)t * EG_TILE + (size_t)tid * EG_VEC;
    push_tile(p, base, ld_f8(p.theta + base));
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence_system();
    const unsigned prev = atomicAdd(p.ticket + 1, 1u);
    s_last = (prev == gridDim.x - 1) ? 1 : 0;
  }
  __syncthreads();
  if (s_last && tid == 0) {
    p.ticket[1] = 0u;
    __threadfence_system();
    st_release_sys(p.pushed_to_l, (uint32_t)step);
    st_release_sys(p.pushed_to_r, (uint32_t)step);
  }
}

template <bool kMom>
__global__ void __launch_bounds__(EG_THREADS, 4) EG_SYM(gossip_step_kernel)(const GossipParams p) {
  const int b = blockIdx.x, G = gridDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int step = *p.fsm.pass_num + 1;
  const int n_tiles = p.tab.n_tiles;
  const bool push = p.do_push != 0 && p.phase != 2;
  // Sticky failure: once any peer wait has timed out (this or an earlier launch) the step is a no-op on every
  // CTA -- nothing is stored into a peer that may be wedged or dead, theta stops changing, and the host raises
  // when it reads the status word.  (Uniform per CTA; the run is over, so a partial step does not matter.)
  if (p.do_push && *reinterpret_cast<volatile int*>(p.status) != EG_OK) return;

#ifdef EG_DBUF
  if (p.phase != 0 || !(p.sync && push)) return;   // the launcher rejects these; never reached
#else
  if (p.phase == 1) {
    push_phase(p, step);
    return;
  }
#endif
  __shared__ TileInfo s_ti[EG_TI_CACHE];
  fill_tile_cache(p, s_ti);
#ifndef EG_DBUF
  // spevent: receive side first -- scatter the records that arrived for this step into the replicas
  // (inbox_l / inbox_r point at them), grid barrier, then the dense mix below reads them tile by tile
  if (p.sparse != nullptr) sparse_apply_prologue(*p.sparse, p.sparse->bar);
#endif
  if (p.phase == 2 && p.sync && p.do_push) {
    // split step, second half: the neighbours' pushes of this step were issued during my backward
    __shared__ int s_pushed;
    if (tid == 0) {
      bool ok = wait_ge(p.pushed_from_l, (uint32_t)step, p.status, p.timeout_ns);
      ok = wait_ge(p.pushed_from_r, (uint32_t)step, p.status, p.timeout_ns) && ok;
      s_pushed = ok ? 1 : 0;
    }
    __syncthreads();
    if (!s_pushed) return;      // neighbour never pushed this step: do not average a stale inbox into theta
  }

  if (!(p.sync && push)) {
    // ---------------- async / split / no exchange: single pass, theta read once ---------------
    int j = 0;
    for (int t = b; t < n_tiles; t += G, ++j) {
      const size_t base = (size_t)t * EG_TILE + (size_t)tid * EG_VEC;
      const TileInfo ti = tile_info(p, s_ti, j, t);
      const F8 th = ld_f8(p.theta + base);
      if (push && (ti.flags & 1)) push_tile(p, base, th);
      mix_tile<kMom>(p, t, ti, base, th, lane, warp, step);
    }
  } else {
    // ---------------- iter-sync: per-WARP software pipeline, no block barriers ------------------
    // Every warp owns 256 floats of each of its CTA's tiles.  It pushes its slice of tile j and
    // publishes a release-flag (tile, warp) to both neighbours, then consumes tile j-D: waits for the
    // neighbours' flags of that slice (normally long set -- the ring is symmetric) and mixes it.
    // Warps never wait for each other, so flag latency and the sys-scope fence of one warp hide
    // behind the streaming of the other 31 warps on the SM.
#ifndef EG_DBUF
    __shared__ int s_ok;
    if (tid == 0) {
      // WAR guard: neighbours must have consumed what I pushed at step-1 before I overwrite it
      bool ok = wait_ge(p.ack_from_l, (uint32_t)(step - 1), p.status, p.timeout_ns);
      ok = wait_ge(p.ack_from_r, (uint32_t)(step - 1), p.status, p.timeout_ns) && ok;
      s_ok = ok;
    }
    __syncthreads();
    // ack timed out (or an earlier kernel already set the sticky status): the peer may still be reading what was
    // pushed last step -- store nothing, mix nothing; the host raises on the status word
    if (!s_ok) return;
#endif
    const int D = p.group_iters;                   // pipeline depth in tiles
    const int iters = (n_tiles + G - 1) / G;
    for (int j = 0; j < iters + D; ++j) {
      const int t = b + j * G;
      if (j < iters && t < n_tiles) {
        const bool fired = (tile_info(p, s_ti, j, t).flags & 1) != 0;
        if (fired) {
          const size_t base = (size_t)t * EG_TILE + (size_t)tid * EG_VEC;
#ifdef EG_DBUF
          push_tile_slot(p, base, ld_f8(p.theta + base), step);
#else
          push_tile(p, base, ld_f8(p.theta + base));
#endif
        }
        __syncwarp();
        if (lane == 0) {
          if (fired) fence_sys();                  // my warp's stores are performed before the flag
          st_release_sys(p.flag_to_l + (size_t)t * EG_WARPS + warp, (uint32_t)step);
          st_release_sys(p.flag_to_r + (size_t)t * EG_WARPS + warp, (uint32_t)step);
        }
      }
      const int t2 = b + (j - D) * G;
      if (j >= D && t2 < n_tiles) {
        if (lane == 0) {
          wait_ge(p.flag_from_l + (size_t)t2 * EG_WARPS + warp, (uint32_t)step, p.status, p.timeout_ns);
          wait_ge(p.flag_from_r + (size_t)t2 * EG_WARPS + warp, (uint32_t)step, p.status, p.timeout_ns);
        }
        __syncwarp();
        const size_t base = (size_t)t2 * EG_TILE + (size_t)tid * EG_VEC;
        mix_tile<kMom>(p, t2, tile_info(p, s_ti, j - D, t2), base, ld_f8(p.theta + base), lane, warp, step);
      }
    }
  }
  if (p.need_norm) cta_finish_tensors(p, s_ti, step + 1, p.do_mix != 0, true);
  grid_tail(p, step);
}

#ifndef EG_DBUF
// (Re)compute tile partials (+ bf16 shadow) from theta; optionally evaluate the trigger for the
// first step.  Used once at start-up and after theta is modified outside the step kernel
// (checkpoint restore).
__global__ void __launch_bounds__(EG_THREADS, 4) gossip_init_kernel(const GossipParams p, int run_fsm) {
  const int b = blockIdx.x, G = gridDim.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int next_step = *p.fsm.pass_num + 1;
  for (int t = b; t < p.tab.n_tiles; t += G) {
    const size_t base = (size_t)t * EG_TILE + (size_t)tid * EG_VEC;
    const F8 th = ld_f8(p.theta + base);
    float ss = 0.f;
#pragma unroll
    for (int e = 0; e < 8; ++e) ss = __fmaf_rn(th.v[e], th.v[e], ss);
    if (p.shadow != nullptr) st_bf16x8(p.shadow + base, th);
    ss = warp_sum(ss);
    if (lane == 0) {
      p.tile_ss[(size_t)t * EG_WARPS + warp] = ss;
      if (p.tile_ss_l != nullptr) {
        p.tile_ss_l[(size_t)t * EG_WARPS + warp] = 0.f;
        p.tile_ss_r[(size_t)t * EG_WARPS + warp] = 0.f;
      }
    }
  }
  if (run_fsm) cta_finish_tensors(p, nullptr, next_step, false, /*recv_ok=*/false);
}

// Trigger FSM alone, norms supplied by the caller (unit test against parallel/trigger.py).
__global__ void __launch_bounds__(EG_THREADS) fsm_decide_kernel(const FsmDev f, const TableDev t,
                                                               const float* ext_norm) {
  const int next_step = *f.pass_num + 1;
  for (int i = threadIdx.x; i < t.n_tensors; i += EG_THREADS)
    fsm_update_tensor(f, t, i, ext_norm[i], 0.f, 0.f, false, next_step, false);
  __syncthreads();
  if (threadIdx.x == 0) *f.pass_num += 1;
}

// -------------------------------------------------------------------------------------------
int gossip_max_grid(int device) {
  int sms = 0, per_sm = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  int per_sm2 = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, gossip_step_kernel<true>, EG_THREADS, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm2, gossip_step_kernel<false>, EG_THREADS, 0);
  if (per_sm2 < per_sm) per_sm = per_sm2;
  if (per_sm < 1) per_sm = 1;
  return sms * per_sm;
}

cudaError_t launch_gossip_step(const GossipParams& p, int grid, cudaStream_t s) {
  eg_count_launch(EG_FAM_GOSSIP, 1);
  if (p.mu != 0.f && p.mom != nullptr)
    gossip_step_kernel<true><<<grid, EG_THREADS, 0, s>>>(p);
  else
    gossip_step_kernel<false><<<grid, EG_THREADS, 0, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_gossip_init(const GossipParams& p, int grid, int run_fsm, cudaStream_t s) {
  eg_count_launch(EG_FAM_GOSSIP, 1);
  gossip_init_kernel<<<grid, EG_THREADS, 0, s>>>(p, run_fsm);
  return cudaGetLastError();
}

cudaError_t launch_fsm_decide(const FsmDev& f, const TableDev& t, const float* ext_norm, cudaStream_t s) {
  eg_count_launch(EG_FAM_GOSSIP, 1);
  fsm_decide_kernel<<<1, EG_THREADS, 0, s>>>(f, t, ext_norm);
  return cudaGetLastError();
}

#else   // EG_DBUF: only the step kernel + its launcher exist in this translation unit
int gossip_dbuf_max_grid(int device) {
  int sms = 0, a = 0, b = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, gossip_step_kernel_dbuf<true>, EG_THREADS, 0);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, gossip_step_kernel_dbuf<false>, EG_THREADS, 0);
  if (b < a) a = b;
  if (a < 1) a = 1;
  return sms * a;
}

// Dense iter-sync exchange only (decent): every tensor fires every step, inboxes hold 2 slots.
cudaError_t launch_gossip_step_dbuf(const GossipParams& p, int grid, cudaStream_t s) {
  if (p.phase != 0 || !p.sync || !p.do_push || !p.do_mix || p.fsm.enabled) return cudaErrorInvalidValue;
  eg_count_launch(EG_FAM_GOSSIP, 1);
  if (p.mu != 0.f && p.mom != nullptr)
    gossip_step_kernel_dbuf<true><<<grid, EG_THREADS, 0, s>>>(p);
  else
    gossip_step_kernel_dbuf<false><<<grid, EG_THREADS, 0, s>>>(p);
  return cudaGetLastError();
}
#endif

}  // namespace egb
