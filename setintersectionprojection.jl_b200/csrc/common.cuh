// common.cuh — shared device utilities for the B200 (sm_100a) PARSDMM kernels.
//
// Everything on this path is HBM-bandwidth bound (no tensor-core work): the helpers here are
//   * 16-byte vector access types,
//   * a deterministic two-level reduction (warp shuffle -> shared -> per-block partial -> the last
//     block to finish folds the partials in a fixed order), accumulated in double even for float
//     data so that branch decisions (stop rules, rho/gamma adaptation, CG tolerance) do not depend
//     on the launch geometry,
//   * error handling macros.
// Compiled with -fmad=false: the reference (Julia, no fast-math) never contracts a*b+c.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <type_traits>

namespace sipb {

typedef long long i64;

constexpr int kThreads = 256;          // threads per CTA for every streaming kernel
constexpr int kMaxBlocks = 148 * 16;   // upper bound on the grid of any reducing kernel
constexpr int kMaxRed = 10;            // max number of simultaneous reductions per kernel

// ---------------------------------------------------------------------------------------------
// vector types: 16-byte accesses (float4 / double2)
// ---------------------------------------------------------------------------------------------
template <typename T> struct Vec;
template <> struct Vec<float> {
  typedef float4 type;
  static constexpr int W = 4;
};
template <> struct Vec<double> {
  typedef double2 type;
  static constexpr int W = 2;
};

template <typename T> __device__ __forceinline__ void vload(const T* p, T (&v)[Vec<T>::W]);
template <> __device__ __forceinline__ void vload<float>(const float* p, float (&v)[4]) {
  float4 t = *reinterpret_cast<const float4*>(p);
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void vload<double>(const double* p, double (&v)[2]) {
  double2 t = *reinterpret_cast<const double2*>(p);
  v[0] = t.x; v[1] = t.y;
}
// streaming (read-once) load: bypass L1 allocation
template <typename T> __device__ __forceinline__ void vload_stream(const T* p, T (&v)[Vec<T>::W]);
template <> __device__ __forceinline__ void vload_stream<float>(const float* p, float (&v)[4]) {
  float4 t = __ldcs(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void vload_stream<double>(const double* p, double (&v)[2]) {
  double2 t = __ldcs(reinterpret_cast<const double2*>(p));
  v[0] = t.x; v[1] = t.y;
}
template <typename T> __device__ __forceinline__ void vstore(T* p, const T (&v)[Vec<T>::W]);
template <> __device__ __forceinline__ void vstore<float>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void vstore<double>(double* p, const double (&v)[2]) {
  *reinterpret_cast<double2*>(p) = make_double2(v[0], v[1]);
}

// ---------------------------------------------------------------------------------------------
// deterministic reductions
// ---------------------------------------------------------------------------------------------
struct RedScratch {
  double* partials;        // [kMaxRed][kMaxBlocks]
  unsigned int* counter;   // ticket for "last block" detection (reset by the last block)
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// Sum NV per-thread doubles over the block; result valid in thread 0.
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double (*sm)[32]) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < NV; ++i) v[i] = warp_sum(v[i]);
  __syncthreads();   // protect sm reuse across calls
  if (lane == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) sm[i][w] = v[i];
  }
  __syncthreads();
  if (w == 0) {
    const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double t = lane < nw ? sm[i][lane] : 0.0;
      v[i] = warp_sum(t);
    }
  }
}

// Returns true (in every thread) for the last block of the grid to arrive.  `sys`: the block's global
// writes must become visible to PEER GPUs before the last block signals them (system-scope fence).
__device__ __forceinline__ bool last_block_ticket(unsigned int* counter, bool sys = false) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    if (sys) __threadfence_system();
    else __threadfence();
    unsigned int t = atomicAdd(counter, 1u);
    s_last = (t == gridDim.x - 1);
    if (s_last) *counter = 0u;   // all blocks have arrived; safe to reset for the next kernel
  }
  __syncthreads();
  return s_last != 0;
}

// Block-reduce NV values, publish the per-block partials, and let the last block fold them in a
// fixed order.  On return `v[0..NV)` holds the grid totals in thread 0 of the LAST block only;
// the function returns true in every thread of that block.
template <int NV>
__device__ __forceinline__ bool grid_sum(double (&v)[NV], const RedScratch& rs, bool sys = false) {
  __shared__ double sm[NV][32];
  block_sum<NV>(v, sm);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) rs.partials[i * kMaxBlocks + blockIdx.x] = v[i];
  }
  const bool last = last_block_ticket(rs.counter, sys);
  if (last) {
    __threadfence();
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      double t = 0.0;
      for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x)
        t += __ldcg(&rs.partials[i * kMaxBlocks + b]);
      v[i] = t;
    }
    block_sum<NV>(v, sm);
  }
  return last;
}

// ---------------------------------------------------------------------------------------------
// peer-memory collectives for slabs (one process per GPU, buffers mapped with CUDA IPC over NVLink)
//
// The CG has three tiny reductions and one halo per iteration; with NCCL each costs a kernel launch and
// ~10-15 us of latency on the critical path.  Here the reducing kernel's last block stores its partial sums
// straight into every peer's mailbox (NVLink peer stores + system fence + flag) and the consuming kernel
// sums the W partials in rank order (bit-identical on all ranks); the SpMV reads the neighbours' boundary
// plane of p through mapped peer pointers after waiting on a version flag the neighbour's p-update kernel
// publishes.  Two mailbox buffers suffice: a rank can only be one reduction ahead of its slowest peer.
// Every spin is bounded (kSpinLimit cycles); on time-out an error flag is raised and the kernel finishes.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxRanks = 16;
constexpr int kMailK = 4;
constexpr long long kSpinLimit = 6000000000ll;   // ~3 s at 1.9 GHz

struct PeerMail {
  double val[2][kMaxRanks][kMailK];
  unsigned long long flag[2][kMaxRanks];
  unsigned long long pver[2];        // [0]: version of the lower neighbour's p, [1]: of the upper neighbour's
};

struct CommDev {
  int on;                            // peer path active
  int rank, world;
  int has_lo, has_hi;
  int* err;                          // device flag: a bounded spin timed out
  unsigned long long* seq;           // device counter: mailbox reductions produced by this rank
  unsigned long long* pv;            // device counter: version of this rank's p vector
  unsigned long long* bseq;          // device-local flag: last reduction whose global value is in place
  PeerMail* mail[kMaxRanks];         // mail[q]: rank q's mailbox (peer mapped; mail[rank] is local)
};

__device__ __forceinline__ unsigned long long ld_vol(const unsigned long long* p) {
  return *reinterpret_cast<const volatile unsigned long long*>(p);
}
__device__ __forceinline__ double ld_vol(const double* p) { return *reinterpret_cast<const volatile double*>(p); }

// Called by every thread of ONE block (the last block of the reducing kernel); v valid in thread 0.
template <int K>
__device__ __forceinline__ void mail_publish(const CommDev& cd, const double (&v)[K]) {
  static_assert(K <= kMailK, "mailbox slot too small");
  __shared__ double sh[K];
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < K; ++i) sh[i] = v[i];
    s_seq = *cd.seq + 1ull;
  }
  __syncthreads();
  const unsigned long long seq = s_seq;
  const int buf = (int)(seq & 1ull);
  if ((int)threadIdx.x < cd.world) {
    PeerMail* pm = cd.mail[threadIdx.x];
#pragma unroll
    for (int i = 0; i < K; ++i) *reinterpret_cast<volatile double*>(&pm->val[buf][cd.rank][i]) = sh[i];
    __threadfence_system();
    *reinterpret_cast<volatile unsigned long long*>(&pm->flag[buf][cd.rank]) = seq;
  }
  __syncthreads();
  if (threadIdx.x == 0) *cd.seq = seq;
}

// Executed by ONE thread (a one-thread kernel between producer and consumer): sums the W partials in rank
// order.  A single poller per GPU: an earlier version let thread 0 of every block of the consumer poll the
// mailbox line, and the ~1200 pollers delayed the peers' NVLink writes to that very line by tens of us.
template <int K>
__device__ __forceinline__ void mail_collect(const CommDev& cd, double* out) {
  const unsigned long long seq = *cd.seq;
  const int buf = (int)(seq & 1ull);
  const PeerMail* me = cd.mail[cd.rank];
  double acc[K];
#pragma unroll
  for (int i = 0; i < K; ++i) acc[i] = 0.0;
  const long long t0 = clock64();
  for (int q = 0; q < cd.world; ++q) {
    while (ld_vol(&me->flag[buf][q]) != seq) {
      __nanosleep(64);
      if (clock64() - t0 > kSpinLimit) { *cd.err = 1; break; }
    }
    __threadfence_system();
#pragma unroll
    for (int i = 0; i < K; ++i) acc[i] += ld_vol(&me->val[buf][q][i]);
  }
#pragma unroll
  for (int i = 0; i < K; ++i) out[i] = acc[i];
}

// Collect inside the CONSUMER kernel (every thread of every block calls this before reading out[]): block 0 is
// the single poller of the peer-written mailbox line; it stores the global value and releases a device-LOCAL
// sequence flag on which thread 0 of the other blocks waits.  Saves the one-thread collector kernel (two kernel
// boundaries, ~5 us) per reduction.  Block 0 is dispatched first and waits for nothing on this GPU, so the
// other blocks cannot starve it.  Read out[] with ld_vol afterwards (L1 may hold the line from an earlier read).
template <int K>
__device__ __forceinline__ void mail_collect_all(const CommDev& cd, double* out) {
  if (threadIdx.x == 0) {
    const unsigned long long seq = *cd.seq;
    if (blockIdx.x == 0) {
      mail_collect<K>(cd, out);
      __threadfence();
      *reinterpret_cast<volatile unsigned long long*>(cd.bseq) = seq;
    } else {
      const long long t0 = clock64();
      while (ld_vol(cd.bseq) < seq) {
        __nanosleep(32);
        if (clock64() - t0 > kSpinLimit) { *cd.err = 1; break; }
      }
      __threadfence();
    }
  }
  __syncthreads();
}

// All-reduce inside the PRODUCER kernel: its last block publishes the rank's partial sums and then waits for the peers'
// partials itself (one poller per GPU, while nothing else runs), so the consumer kernel simply reads `out`.  Measured on
// 8 GPUs the consumer-side collect (block 0 polls the mailbox, ~700 other blocks poll a local release flag) cost ~30 us
// per reduction, far above the NVLink round trip.  Called by every thread of the last block; v valid in thread 0.
template <int K>
__device__ __forceinline__ void mail_allreduce(const CommDev& cd, const double (&v)[K], double* out) {
  mail_publish<K>(cd, v);
  if (threadIdx.x == 0) {
    mail_collect<K>(cd, out);
    __threadfence();
  }
}

// thread 0 of the LAST block of a kernel that rewrote p (every block fenced at system scope before its ticket)
__device__ __forceinline__ void p_publish(const CommDev& cd) {
  const unsigned long long v = *cd.pv + 1ull;
  *cd.pv = v;
  __threadfence_system();
  if (cd.has_lo) *reinterpret_cast<volatile unsigned long long*>(&cd.mail[cd.rank - 1]->pver[1]) = v;
  if (cd.has_hi) *reinterpret_cast<volatile unsigned long long*>(&cd.mail[cd.rank + 1]->pver[0]) = v;
}
// every thread of a block of the SpMV that touches a boundary plane: wait until that neighbour has
// published the current version of p (only the few blocks that own boundary rows poll)
__device__ __forceinline__ void p_wait(const CommDev& cd, bool need_lo, bool need_hi) {
  if (threadIdx.x == 0) {
    const unsigned long long v = *cd.pv;
    const PeerMail* me = cd.mail[cd.rank];
    const long long t0 = clock64();
    if (need_lo && cd.has_lo)
      while (ld_vol(&me->pver[0]) < v) { __nanosleep(64); if (clock64() - t0 > kSpinLimit) { *cd.err = 1; break; } }
    if (need_hi && cd.has_hi)
      while (ld_vol(&me->pver[1]) < v) { __nanosleep(64); if (clock64() - t0 > kSpinLimit) { *cd.err = 1; break; } }
    __threadfence_system();
  }
  __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Small all-reduces (<= kBigN 64-bit words) over peer memory: the once-per-iteration collectives of a slab solve — the
// batched log / adaptation sums, projector statistics, 256-bin radix histograms, tie counts — cost a kernel of one
// block and an NVLink round trip (~5 us) instead of an NCCL launch (~25-40 us each, ~10 per PARSDMM iteration).
// Every rank stores its words into every peer's slot, fences, raises a sequence flag; then it waits for all flags
// and reduces the slots in rank order (bit-identical on all ranks).  Two buffers: a rank can be at most one
// collective ahead of its slowest peer.
// ---------------------------------------------------------------------------------------------
constexpr int kBigN = 512;
struct PeerBig {
  unsigned long long data[2][kMaxRanks][kBigN];
  unsigned long long flag[2][kMaxRanks];
};
struct BigDev {
  int rank, world;
  int* err;
  unsigned long long* seq;           // device counter of collectives done by this rank
  PeerBig* box[kMaxRanks];           // box[q]: rank q's buffer (peer mapped; box[rank] is local)
};
// OP 0: sum of doubles, 1: sum of uint64, 2: min of uint64.  One block of kBigN threads, every rank launches it.
template <int OP>
__global__ void __launch_bounds__(kBigN) k_peer_allreduce(const __grid_constant__ BigDev bd, unsigned long long* buf, int count) {
  __shared__ unsigned long long s_seq;
  if (threadIdx.x == 0) s_seq = *bd.seq + 1ull;
  __syncthreads();
  const unsigned long long seq = s_seq;
  const int b = (int)(seq & 1ull);
  const int t = threadIdx.x;
  if (t < count) {
    const unsigned long long v = buf[t];
    for (int q = 0; q < bd.world; ++q)
      *reinterpret_cast<volatile unsigned long long*>(&bd.box[q]->data[b][bd.rank][t]) = v;
  }
  __threadfence_system();
  __syncthreads();
  if (t < bd.world) {
    *reinterpret_cast<volatile unsigned long long*>(&bd.box[t]->flag[b][bd.rank]) = seq;      // one flag store per peer
    const PeerBig* me = bd.box[bd.rank];
    const long long t0 = clock64();
    while (ld_vol(&me->flag[b][t]) != seq) {
      __nanosleep(64);
      if (clock64() - t0 > kSpinLimit) { *bd.err = 1; break; }
    }
    __threadfence_system();
  }
  __syncthreads();
  if (t < count) {
    const PeerBig* me = bd.box[bd.rank];
    if (OP == 0) {
      double acc = 0.0;
      for (int q = 0; q < bd.world; ++q) acc += __longlong_as_double((long long)ld_vol(&me->data[b][q][t]));
      buf[t] = (unsigned long long)__double_as_longlong(acc);
    } else if (OP == 1) {
      unsigned long long acc = 0ull;
      for (int q = 0; q < bd.world; ++q) acc += ld_vol(&me->data[b][q][t]);
      buf[t] = acc;
    } else {
      unsigned long long acc = ~0ull;
      for (int q = 0; q < bd.world; ++q) { const unsigned long long v = ld_vol(&me->data[b][q][t]); acc = v < acc ? v : acc; }
      buf[t] = acc;
    }
  }
  if (t == 0) *bd.seq = seq;
}

// ---------------------------------------------------------------------------------------------
// device-side loops: the CG iteration and the l1 threshold search run as the body of a CUDA-graph WHILE node; the
// kernel that decides convergence sets the loop condition (no host poll, no speculative launches)
// ---------------------------------------------------------------------------------------------
struct LoopCond {
  cudaGraphConditionalHandle h;
  int on;                 // 0: the kernel runs outside a graph (host-driven loop), the handle is unused
};
__device__ __forceinline__ void loop_set(const LoopCond& lc, bool keep_going) {
  if (lc.on) cudaGraphSetConditional(lc.h, keep_going ? 1u : 0u);
}

// ---------------------------------------------------------------------------------------------
// host-side error handling
// ---------------------------------------------------------------------------------------------
void set_error(const std::string& msg);

#define SIPB_CUDA_CHECK(expr)                                                              \
  do {                                                                                     \
    cudaError_t _e = (expr);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ::sipb::set_error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" + __FILE__ + \
                        ":" + std::to_string(__LINE__) + ")");                            \
      return SIPB_E_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

template <typename T> struct Eps;
template <> struct Eps<float> { static constexpr double v = 1.1920928955078125e-07; };
template <> struct Eps<double> { static constexpr double v = 2.220446049250313e-16; };

}  // namespace sipb
