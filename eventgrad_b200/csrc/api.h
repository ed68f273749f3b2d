// eventgrad_b200 -- host-visible launch API of the sm_100a kernels.
// Pure CUDA (no torch headers) so the .cu files compile in seconds; bindings.cpp adapts tensors.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define EG_TILE 2048     // fp32 elements per tile == parallel/arena.py:TILE
#define EG_THREADS 256

namespace egb {

// ------------------------------------------------------------------ launch accounting (host)
// Every launcher of this extension counts the kernels it enqueues (per family), so "how many of OUR kernels ran
// in this region" is counted, not derived from a formula (bench.py `gpu_launches`; during CUDA-graph capture the
// Trainer records the per-step delta and multiplies by the replays).
enum { EG_FAM_GOSSIP = 0, EG_FAM_ALLREDUCE, EG_FAM_SPARSE, EG_FAM_BN, EG_FAM_LINEAR, EG_FAM_DATA, EG_FAM_CONV, EG_FAM_N };
void eg_count_launch(int family, int n);

// ------------------------------------------------------------------ tensor table (device)
struct TableDev {
  const int* tile_tensor;   // [n_tiles]  tile -> parameter tensor
  const int* t_tile_start;  // [n_tensors]
  const int* t_tile_count;  // [n_tensors]
  const int* t_numel;       // [n_tensors]
  const int* t_msg_bytes;   // [n_tensors] bytes one neighbour receives when tensor i fires
  int n_tiles;
  int n_tensors;
};

// ------------------------------------------------------------------ event-trigger FSM (device)
// Mirrors /root/reference/dcifar10/event/event.cpp:300-365 -- see parallel/trigger.py (oracle).
struct FsmDev {
  float* thres;        // [sz]
  float* last_norm;    // [sz]
  float* last_iter;    // [sz]
  float* slopes;       // [sz * history]
  int* fire;           // [sz] decision for the step that is about to run
  float* cur_norm;     // [sz] sender norm used for that decision
  unsigned long long* counters;  // [0] events (+2 per fire)  [1] payload bytes pushed  [2] fired tensors
  int* pass_num;       // completed steps (device-resident so CUDA graphs can replay the step)
  float* log_ring;     // optional [log_cap][sz][5] = norm, thres, fired, left_norm, right_norm
  int log_cap;
  float horizon;
  float constant;
  int thres_type;      // 1 adaptive, 0 constant
  int history;         // sent_history (<= 8)
  int initial_comm_passes;
  int enabled;         // 0: decent / cent (no trigger: fire[] stays all-ones)
};

struct SparseParams;
// ------------------------------------------------------------------ fused gossip step
// One launch = push theta_k to both ring neighbours (only tensors whose trigger fired) ->
// [iter-sync: flag handshake] -> (theta+L+R)/3 -> SGD(+momentum) -> write theta_{k+1} (+bf16
// shadow, zero grad) accumulating per-warp sum-of-squares -> last CTA evaluates the trigger FSM
// for step k+1 from those partials.  Replaces, per step, 86x{norm.item(), 2 MPI_Put, 2 memcpy,
// add_, add_, div_} + optimizer.step() of the reference.
struct GossipParams {
  float* theta;
  float* grad;
  float* mom;               // may be null when mu == 0
  const float* inbox_l;     // local: what my LEFT neighbour last pushed (or its sparse replica)
  const float* inbox_r;
  float* push_l;            // peer-mapped: LEFT neighbour's inbox_r
  float* push_r;            // peer-mapped: RIGHT neighbour's inbox_l
  __nv_bfloat16* shadow;    // optional bf16 copy of theta_{k+1} for the next forward
  const void* const* t_grad_ptr;  // optional [sz]: read each tensor's gradient in place (table mode)
  const int* t_grad_bf16;         // [sz] 1: that gradient is bf16, 0: fp32
  float* tile_ss;           // [n_tiles * 8] per-warp sum theta_{k+1}^2
  float* tile_ss_l;         // optional (logging): sum inbox_l^2
  float* tile_ss_r;
  // iter-sync handshake (all monotonic step counters)
  uint32_t* flag_from_l;    // local, written by left neighbour   [n_tiles * 8] (tile, warp)
  uint32_t* flag_from_r;
  uint32_t* flag_to_l;      // peer-mapped: left neighbour's flag_from_r
  uint32_t* flag_to_r;      // peer-mapped: right neighbour's flag_from_l
  uint32_t* ack_from_l;     // local: left neighbour finished reading what I pushed at step k
  uint32_t* ack_from_r;
  uint32_t* ack_to_l;       // peer-mapped
  uint32_t* ack_to_r;
  uint32_t* pushed_from_l;  // split step: left neighbour has finished ALL its pushes of step k
  uint32_t* pushed_from_r;
  uint32_t* pushed_to_l;    // peer-mapped
  uint32_t* pushed_to_r;
  unsigned int* ticket;     // [0] last-CTA election (step) [1] push phase
  unsigned int* tensor_done; // [sz] per-tensor count of per-warp partials written this launch
  int* status;              // sticky error word
  unsigned long long timeout_ns;
  TableDev tab;
  FsmDev fsm;
  float lr;
  float mu;
  int do_mix;               // 0: plain SGD (serial run)
  int do_push;              // 0: sparse mode (records are pushed by the top-k kernels) / serial
  int sync;                 // 1: iter-sync handshake, 0: async (reference RMA semantics)
  int send_ack;             // sync: write acks in the tail (0 when another kernel acks)
  int zero_grad;
  int group_iters;          // iter-sync software-pipeline depth in tiles (push j, mix j-D)
  int vec256_push;          // 1: 256-bit peer stores, 0: 2x128-bit
  const struct SparseParams* sparse;   // spevent: device copy of the sparse parameters -> the kernel's prologue scatters
                            // the arrived (value, index) records into inbox_l / inbox_r (the replicas) first
  int need_norm;            // 0: skip norm-on-write + trigger entirely (decent/cent without logs)
  int phase;                // 0 fused step | 1 push only (side stream, overlaps backward) | 2 wait+mix+SGD
};

int gossip_max_grid(int device);  // co-resident CTAs (persistent grid upper bound)
cudaError_t launch_gossip_step(const GossipParams& p, int grid, cudaStream_t s);
// experimental double-buffered dense variant (csrc/gossip_dbuf.cu): inbox_l/inbox_r/push_l/push_r address
// TWO consecutive slots of n_tiles*EG_TILE floats; requires phase 0, sync, do_push, do_mix, fsm disabled
// experimental copy-engine push of the split step (csrc/ce_push.cu): ack wait -> 2 x cudaMemcpyAsync -> flags
cudaError_t launch_ce_push(const GossipParams& p, cudaStream_t s);
int gossip_dbuf_max_grid(int device);
cudaError_t launch_gossip_step_dbuf(const GossipParams& p, int grid, cudaStream_t s);
// (re)compute tile_ss (+ shadow) from theta and evaluate the trigger for step pass_num+1.
cudaError_t launch_gossip_init(const GossipParams& p, int grid, int run_fsm, cudaStream_t s);
// Trigger FSM alone with externally supplied norms (unit tests against the oracle).
cudaError_t launch_fsm_decide(const FsmDev& f, const TableDev& t, const float* ext_norm, cudaStream_t s);

// ------------------------------------------------------------------ one-/two-shot all-reduce
// cent: grad <- sum_r grad_r / R fused with the SGD step; also final parameter averaging.
struct AllReduceParams {
  float* const* peer_bufs;  // device array [world] of peer-mapped pointers to the SAME buffer on each rank
  float* local;             // == peer_bufs[rank]
  float* theta;             // SGD mode: parameters (may alias nothing in `local`)
  float* mom;
  __nv_bfloat16* shadow;
  uint32_t* const* peer_flags;  // device array [world]: each rank's flag block (peer-mapped)
  uint32_t* flags;              // local flag block: [3][grid][world]
  unsigned int* ticket;
  int* status;
  int* step_ctr;            // device-resident launch counter for the flag protocol
  unsigned long long timeout_ns;
  int n_tiles;
  int rank;
  int world;
  float lr;
  float mu;
  int mode;                 // 0: average in place   1: average + SGD (+ zero grad)
  int two_shot;             // 0: every rank reads all peers   1: reduce-scatter + broadcast
  int zero_after;
  float* mc_local;          // NVLS variant only: multicast (NVSwitch) mapping of `local`; null otherwise
};
cudaError_t launch_allreduce(const AllReduceParams& p, int grid, cudaStream_t s);
// experimental NVLink-SHARP variant (csrc/allreduce_nvls.cu): multimem.ld_reduce + multimem.st, two-shot only
cudaError_t launch_allreduce_nvls(const AllReduceParams& p, int grid, cudaStream_t s);

// ------------------------------------------------------------------ sparse (top-k) exchange
struct SparseParams {
  const float* theta;
  float* prev;              // what this rank last sent (element-wise)
  float* rep_l;             // persistent replicas of the neighbours
  float* rep_r;
  // records: per tensor [vals(k_i) | idx(k_i)] at rec_off[i] (4-byte words)
  const float* rec_from_l;  // local inbox records written by neighbours
  const float* rec_from_r;
  float* rec_to_l;          // peer-mapped: left neighbour's rec_from_r
  float* rec_to_r;          // peer-mapped: right neighbour's rec_from_l
  uint32_t* seq_from_l;     // [sz] local: step at which the record of tensor i was last rewritten
  uint32_t* seq_from_r;
  uint32_t* seq_to_l;       // peer-mapped
  uint32_t* seq_to_r;
  uint32_t* applied_l;      // [sz] local bookkeeping: last seq applied to the replica
  uint32_t* applied_r;
  uint32_t* done_from_l;    // sync: neighbour finished writing all records of step k
  uint32_t* done_from_r;
  uint32_t* done_to_l;
  uint32_t* done_to_r;
  uint32_t* ack_from_l;
  uint32_t* ack_from_r;
  uint32_t* ack_to_l;
  uint32_t* ack_to_r;
  const int* t_k;           // [sz] k_i
  const int* t_rec_off;     // [sz] record offset (words)
  // scratch
  unsigned int* hist;       // [sz][2048] (left all-zero by every launch)
  unsigned int* sel_prefix; // [sz] pass 1: top 11 bits of tau; after pass 2: tau = the k-th largest |diff| key
  unsigned int* sel_remain; // [sz] pass 1: rank inside the bucket; after pass 2: number of keys == tau to select
  uint32_t* cand;           // [n_tiles * EG_TILE] candidate keys (low 21 bits), tensor i at t_tile_start[i]*EG_TILE
  unsigned int* cand_cnt;   // [sz]
  unsigned int* done1;      // [sz] per-tensor tile-completion counters of pass 1 / pass 2
  unsigned int* done2;
  unsigned long long* desc; // [n_tiles] look-back descriptors of pass 3
  unsigned int* bar;        // [2] grid barrier of the receive prologue (arrivals, generation)
  const int* fire;
  const int* pass_num;
  unsigned int* ticket;
  int* status;
  unsigned long long timeout_ns;
  TableDev tab;
  int sync;
};
// 3 launches: histogram+pick, candidates+exact threshold, ordered compaction straight into the neighbours' inboxes.
// `grid` must be a co-resident persistent grid (pass 3 uses a decoupled look-back between CTAs).
cudaError_t launch_sparse_select_push(const SparseParams& p, int grid, cudaStream_t s);
cudaError_t launch_sparse_apply(const SparseParams& p, int grid, cudaStream_t s);

// ------------------------------------------------------------------ data path
// uint8 NCHW batch -> normalised fp32/bf16 (NCHW or NHWC) with pad+flip+crop folded in.
cudaError_t launch_decode_augment(const uint8_t* in, void* out, const int* oy, const int* ox,
                                  const int* flip, int B, int C, int H, int W, int pad, float scale,
                                  float mean, float inv_std, int out_bf16, int nhwc, cudaStream_t s);

// ------------------------------------------------------------------ fused BatchNorm(+add)(+ReLU)
// NHWC activations [M = N*H*W, C] in fp32 (reference precision) or bf16, fp32 statistics / affine parameters
// (csrc/bn_act.cu).
struct BnParams {
  const void* x;              // BN input (conv output)
  const void* res;            // optional residual added before the activation
  void* y;                    // output (forward) / saved output for the ReLU mask (backward)
  const void* dy;             // backward: grad wrt y
  void* dx;                   // backward: grad wrt x
  void* dres;                 // backward: grad wrt residual (optional)
  int fp32;                   // 1: activations are float, 0: __nv_bfloat16
  const float* gamma;
  const float* beta;
  float* mean;                // saved batch mean   [C]
  float* invstd;              // saved 1/sqrt(var+eps)
  float* run_mean;            // optional running stats (updated in the stats kernel)
  float* run_var;
  long long* nbt;             // optional num_batches_tracked
  float* dgamma;              // backward outputs (fp32) -- also inputs of the dx kernel
  float* dbeta;
  float* partial;             // workspace [bn_partial_rows][128]
  unsigned int* ticket;       // [64]: per-slice tickets [0..31], global epoch ticket [63]
  unsigned int* flag;         // [32] per-slice epoch flags (fused kernels)
  unsigned int* epoch;        // [1]  launch epoch (device resident: graph-replay safe)
  int* status;                // sticky error word (2 = fused-kernel wait timed out)
  long long M;
  int C;
  float eps;
  float momentum;
  int relu;
  int fused_ok;               // allow the single-launch fused path when the grid is co-resident
  // fp32 only, optional: ALSO emit the result as three bf16 planes [3][M*C] (x = p0+p1+p2) -- the operand format of
  // the tensor-core convolutions (csrc/conv_tc.cu), so the consumer conv needs no separate split pass
  __nv_bfloat16* y_planes;    // forward: planes of y
  __nv_bfloat16* dx_planes;   // backward: planes of dx (= dY of the convolution that produced x)
};
int bn_partial_rows(int sm_count);
// fp32 NCHW variant (csrc/bn_nchw.cu): `hw` = H*W (multiple of 4), p.M = N*H*W.  Workspace: p.partial >= C*64*2
// floats, p.ticket >= C words (p.flag / p.epoch unused).
cudaError_t launch_bn_nchw(const BnParams& p, int hw, int which, int sm_count, cudaStream_t s);
// which: 0 training forward, 1 apply only (eval), 2 backward
cudaError_t launch_bn(const BnParams& p, int which, int sm_count, cudaStream_t s);

// ------------------------------------------------------------------ tcgen05 fused Linear(+bias)(+ReLU)
// Y[M,N] = act(X[M,K] * W[N,K]^T + b): bf16 operands, fp32 accumulation in TMEM (csrc/linear_tc_tma.cu).
struct LinearParams {
  const __nv_bfloat16* x;   // [M, K] row-major
  const __nv_bfloat16* w;   // [N, K] row-major (nn.Linear weight layout)
  const float* bias;        // [N] or null
  void* y;                  // [M, N] bf16 or fp32
  int M, N, K;              // N % 16 == 0, 16 <= N <= 256, K % 8 == 0
  int relu;
  int out_bf16;
};
// TMA + SWIZZLE_128B + persistent CTAs (one per SM) + double-buffered TMEM accumulator
cudaError_t launch_linear_tc_tma(const LinearParams& p, int sm_count, cudaStream_t s);

// ------------------------------------------------------------------ tcgen05 convolutions at fp32 accuracy
// (csrc/conv_tc.cu): fp32 tensors split into three bf16 planes, six bf16 MMAs per fp32 product, fp32 accumulation.
// One kernel pair + a tap table covers 3x3 stride 1/2, 1x1 stride 2 and the 3-channel stem (see conv_tc.cu).
struct ConvTcParams {
  const __nv_bfloat16* a;   // activation planes [3][nsrc][N][H][W][Ca] (forward: x; dgrad: dY; wgrad: x)
  const __nv_bfloat16* b;   // fprop/dgrad: weight planes [3][Cb][wtaps*Ca]; wgrad: dY planes [3][N][H][W][Cb]
  float* out;               // fprop/dgrad: [N][OH][OW][Cb]; wgrad: workspace [splits][ntaps*Ca][Cb]
  int N, H, W;              // pixel grid of the GEMM rows
  int Ca, Cb;
  int ntaps, nsrc, wtaps;   // taps of this launch, source sub-images per plane, taps in the weight matrix
  signed char dh[9], dw[9], src[9], wk[9];
  int OH, OW, os, op, oq;   // fprop: pixel (n,i,j) -> out[n][i*os+op][j*os+oq]
  float* dwout;             // wgrad: [Cb][ntaps*Ca] (OHWI); written directly by the kernel when there is one split
  float* ws;                // fprop: workspace [ksplits][m_tiles*128][Cb] when ksplits > 1
  int ksplits;              // fprop: K-loop splits (conv_fprop_ksplits(); 1 = none)
  int bh, bn, m_tiles, k_blocks;   // filled in by the launchers
};
bool conv_tc_supported(int N, int H, int W, int Ca, int Cb);
int conv_wgrad_splits(int N, int H, int W, int Ca, int Cb, int ntaps, int sm_count);
int conv_fprop_ksplits(int N, int H, int W, int Ca, int Cb, int ntaps, int sm_count);
int conv_fprop_mtiles(int N, int H, int W);   // 128-pixel tiles of the grid (workspace rows = 128 * this)
cudaError_t launch_conv_fprop(const ConvTcParams& p, int sm_count, cudaStream_t s);
cudaError_t launch_conv_wgrad(const ConvTcParams& p, float* dw, int splits, cudaStream_t s);
cudaError_t launch_split3(const float* src, __nv_bfloat16* dst, size_t n, cudaStream_t s);   // n % 8 == 0
cudaError_t launch_split3_parity(const float* src, __nv_bfloat16* dst, int N, int H, int W, int C, cudaStream_t s);
cudaError_t launch_split3_stem(const float* src, __nv_bfloat16* dst, int N, int H, int W, cudaStream_t s);
cudaError_t launch_conv_wprep(const float* w, __nv_bfloat16* wp, __nv_bfloat16* wtp, int Co, int T, int Ci, cudaStream_t s);

// ------------------------------------------------------------------ IPC window runtime
// The RMA-window replacement (MPI_Alloc_mem + MPI_Win_create, event.cpp:138-147).
struct IpcHandle {
  unsigned char bytes[64];
};
cudaError_t ipc_alloc(size_t nbytes, void** ptr, IpcHandle* h);
cudaError_t ipc_open(const IpcHandle& h, void** ptr);
cudaError_t ipc_close(void* ptr);
cudaError_t ipc_free(void* ptr);

}  // namespace egb
