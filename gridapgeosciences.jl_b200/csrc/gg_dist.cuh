// Device-side primitives of the multi-GPU path: halo exchange and deterministic all-reduce through peer memory.
//
// Replaces PartitionedArrays' `consistent!` (owner -> ghost update; reference call sites src/ODEs/StageOperators.jl:109,134,
// src/ODEs/DAE.jl:291,314) and the MPI_Allreduce of the Krylov dot products on the mass-solve path.  One process per GPU;
// every rank exports one cudaMalloc'ed "window" through CUDA IPC and maps the windows of all other ranks, so a kernel
// stores halo values and reduction operands straight into the peer's HBM over NVLink / NVSwitch and signals with a flag --
// no host round trip, no NCCL launch inside a solve (the messages are a few KB: latency, not bandwidth, is what counts).
//
// Protocol.  All ranks execute the same sequence of rounds; a round is numbered by a monotonically increasing epoch.
//   writer:  data stores to peer windows (any thread) ; grid barrier (device-scope release/acquire) ; ONE thread issues
//            __threadfence_system() -- fences are cumulative, so it also orders the other threads' stores -- and then
//            stores flag[me] = epoch into every peer window.  (A system-scope fence in every writer thread is not needed
//            and is expensive.)
//   reader:  one thread spins (volatile loads of its OWN window) until flag[r] >= epoch for all r ; grid barrier ; read data
// Reduction operands AND the exchange slot are double-buffered by the parity of the round that completes them (a rank can be
// at most one round ahead of a peer, because it needs that peer's flag to finish a round: data stored for round e is read
// after round e and before the reader posts e+1; the next store into the same half belongs to round e+2, whose writer has
// finished e+1, i.e. has seen the reader's flag e+1).  Sums run over ranks in rank order on every rank => bitwise identical totals
// everywhere, so all ranks take the same convergence decisions.  A wait that exceeds ~10 s sets an error word and every
// later wait returns at once: a lost peer becomes an error code, not a hung GPU.
#pragma once
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace gg {

constexpr int GG_MAX_RANKS = 16;
constexpr int GG_RED_WIDTH = 4;

struct WinHeader {
  unsigned long long flag[GG_MAX_RANKS];              // flag[src] = last epoch posted by rank src
  double red[2][GG_MAX_RANKS][GG_RED_WIDTH];          // reduction operands by epoch parity and source rank
  unsigned long long epoch;                           // this rank's round counter (persists across launches)
  unsigned long long error;                           // != 0 after a timed-out wait
  unsigned long long wait_ns, post_ns, rounds;        // measurement: time spent spinning / posting, rounds so far
  unsigned long long pad[3];
};

struct DistDev {
  int rank = 0, world = 1, force_multi = 0, timing = 0;
  WinHeader* const* win = nullptr;   // [world] window bases (peers' are IPC mappings)
  double* const* slot = nullptr;     // [world] exchange slots (doubles) behind the headers: two halves of slot_stride doubles
  int64_t slot_stride = 0;           // half used by an exchange = parity of the round (epoch) that completes it
  const uint8_t* owned = nullptr;    // [n local DOFs] 0 ghost, 1 owned, 2 owned and a ghost of some neighbour
  // DOFs this rank owns and a neighbour holds as ghosts, sorted by local id: local id, the receiver's local id and rank
  const int32_t *send_idx = nullptr, *send_remote = nullptr, *send_peer = nullptr;
  const int32_t* push_first = nullptr;  // [n local DOFs] first send entry of the row (rows with owned == 2)
  int64_t n_send = 0;
  // ghost DOFs of this rank
  const int32_t* recv_idx = nullptr;
  int64_t n_recv = 0;
};

__device__ __forceinline__ unsigned long long dist_now() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// every thread: store the owned values the neighbours need into THEIR exchange slots, at their local ids
// `epoch` is the caller's CURRENT round counter: the exchange is completed by the next round, epoch + 1
__device__ __forceinline__ void dist_push(const DistDev& d, unsigned long long epoch, const double* __restrict__ vec, int64_t tid,
                                          int64_t nth, int64_t limit) {
  const int64_t half = (int64_t)((epoch + 1ull) & 1ull) * d.slot_stride;
  for (int64_t k = tid; k < d.n_send; k += nth) {
    const int32_t i = d.send_idx[k];
    if (i < limit) d.slot[d.send_peer[k]][half + d.send_remote[k]] = vec[i];
  }
}

// push one freshly computed value of row i (owned[i] == 2) to every rank that holds the row as a ghost
__device__ __forceinline__ void dist_push_row(const DistDev& d, unsigned long long epoch, int64_t i, double v) {
  const int64_t half = (int64_t)((epoch + 1ull) & 1ull) * d.slot_stride;
  for (int64_t k = d.push_first[i]; k < d.n_send && d.send_idx[k] == i; ++k) d.slot[d.send_peer[k]][half + d.send_remote[k]] = v;
}

// every thread: copy the received ghost values from this rank's slot into vec; `epoch` is the round that completed the exchange
// (the counter as dist_allreduce left it)
__device__ __forceinline__ void dist_unpack(const DistDev& d, unsigned long long epoch, double* __restrict__ vec, int64_t tid,
                                            int64_t nth, int64_t limit) {
  const double* s = d.slot[d.rank] + (int64_t)(epoch & 1ull) * d.slot_stride;
  for (int64_t k = tid; k < d.n_recv; k += nth) {
    const int32_t i = d.recv_idx[k];
    if (i < limit) vec[i] = __ldcg(s + i);
  }
}

// One round: all-reduce of NV values (val holds the rank-local totals in every thread on entry, the global totals on exit)
// and, implicitly, completion of every dist_push issued before the grid barrier that precedes this call.
// Must be called by all threads of a cooperative grid; contains one grid barrier.
template <int NV>
__device__ __forceinline__ void dist_allreduce(cooperative_groups::grid_group& grid, const DistDev& d, unsigned long long& epoch,
                                               double (&val)[NV]) {
  ++epoch;
  const int par = (int)(epoch & 1ull);
  WinHeader* mine = d.win[d.rank];
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned long long tp = dist_now();
    for (int r = 0; r < d.world; ++r) {
      volatile double* dst = d.win[r]->red[par][d.rank];
#pragma unroll
      for (int v = 0; v < NV; ++v) dst[v] = val[v];
    }
    __threadfence_system();
    for (int r = 0; r < d.world; ++r)
      if (r != d.rank) *(volatile unsigned long long*)&d.win[r]->flag[d.rank] = epoch;
    const unsigned long long t0 = dist_now();
    for (int r = 0; r < d.world; ++r) {
      if (r == d.rank) continue;
      volatile unsigned long long* f = &mine->flag[r];
      while (*f < epoch) {
        if (*(volatile unsigned long long*)&mine->error) break;
        if (dist_now() - t0 > 10000000000ull) { *(volatile unsigned long long*)&mine->error = epoch; break; }
      }
    }
    __threadfence_system();
    const unsigned long long t1 = dist_now();
    if (d.timing) { mine->wait_ns += t1 - t0; mine->post_ns += t0 - tp; mine->rounds += 1; }
  }
  grid.sync();
  // one thread per CTA reads the operands (a volatile read by every thread of the grid serialises 2400 warps on one L2
  // line: measured 35 us on B200) and broadcasts the sums through shared memory
  __shared__ double bc[GG_RED_WIDTH];
  if (threadIdx.x == 0) {
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      double s = 0.0;
      for (int r = 0; r < d.world; ++r) s += *(volatile double*)&mine->red[par][r][v];
      bc[v] = s;
    }
  }
  __syncthreads();
#pragma unroll
  for (int v = 0; v < NV; ++v) val[v] = bc[v];
  __syncthreads();
}

}  // namespace gg
