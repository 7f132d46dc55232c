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

// Returns true (in every thread) for the last block of the grid to arrive.
__device__ __forceinline__ bool last_block_ticket(unsigned int* counter) {
  __shared__ int s_last;
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
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
__device__ __forceinline__ bool grid_sum(double (&v)[NV], const RedScratch& rs) {
  __shared__ double sm[NV][32];
  block_sum<NV>(v, sm);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) rs.partials[i * kMaxBlocks + blockIdx.x] = v[i];
  }
  const bool last = last_block_ticket(rs.counter);
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
