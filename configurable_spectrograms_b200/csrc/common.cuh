// Shared device helpers for libcsgpu (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <math_constants.h>
#include <stdint.h>

#include "csgpu.h"

struct csg_ctx {
  int device;
  cudaStream_t stream;
  bool own_stream;
  int sm_count;
  cudaEvent_t ev_start[32];
  cudaEvent_t ev_stop[32];
  cudaEvent_t ev_user[32];
  cudaStream_t side;      // copy-out stream: result read-backs that overlap the next step's uploads
  cudaEvent_t ev_fork;    // ctx stream -> side stream
  cudaEvent_t ev_side;    // last copy enqueued on the side stream
  int64_t launches;
  void* scratch;  // device scratch owned by the context (grow-only)
  size_t scratch_bytes;
  int raster_blocks_per_sm;  // csg_rasterise_blocks_per_sm(): cap on K3's persistent blocks per SM (0 = as many as fit)
  int stats_force_exact;  // csg_region_stats_force_exact(): K2a sends every percentile region to the exact select
  char err[512];
};

extern char g_csg_err[512];
int csg_fail(csg_ctx* ctx, int status, const char* fmt, ...);
int csg_scratch(csg_ctx* ctx, size_t bytes, void** out);  // stats.cu
// Fill on the ctx stream with a KERNEL (ctx.cu): unlike cudaMemsetAsync it never rides a copy-engine
// queue, where a fill waiting behind a spinning peer-wait kernel would hold up other streams' copies.
int csg_fill(csg_ctx* ctx, void* d_dst, int byte_value, size_t bytes);

#define CSG_CUDA(ctx, call)                                                                  \
  do {                                                                                       \
    cudaError_t _e = (call);                                                                 \
    if (_e != cudaSuccess)                                                                   \
      return csg_fail((ctx), CSG_ERR_CUDA, "%s failed: %s (%s:%d)", #call,                   \
                      cudaGetErrorString(_e), __FILE__, __LINE__);                           \
  } while (0)

#define CSG_LAUNCH_CHECK(ctx, name)                                                          \
  do {                                                                                       \
    cudaError_t _e = cudaGetLastError();                                                     \
    if (_e != cudaSuccess)                                                                   \
      return csg_fail((ctx), CSG_ERR_CUDA, "launch of %s failed: %s", name,                  \
                      cudaGetErrorString(_e));                                               \
    (ctx)->launches++;                                                                       \
  } while (0)

// ------------------------------------------------------------------ arithmetic in D
// numpy / matplotlib round after every operation; never let the compiler contract a*b+c.
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_rn(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float div_rn(float a, float b) { return __fdiv_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_rn(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double div_rn(double a, double b) { return __ddiv_rn(a, b); }

template <typename T>
__device__ __forceinline__ bool is_nan(T v) {
  return v != v;
}
__device__ __forceinline__ bool is_finite(float v) { return fabsf(v) <= 3.402823466e+38f; }
__device__ __forceinline__ bool is_finite(double v) { return fabs(v) <= 1.7976931348623157e+308; }

// ------------------------------------------------------- order-preserving float keys
template <typename T>
struct Key;
template <>
struct Key<float> {
  typedef uint32_t U;
  static constexpr int BITS = 32;
  static constexpr int POS_BITS = 31;  // finite positive floats: the raw bit pattern is monotone
  __device__ static __forceinline__ U key(float v) {
    U b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
  }
  __device__ static __forceinline__ float val(U k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
  }
  __device__ static __forceinline__ U bits(float v) { return __float_as_uint(v); }
};
template <>
struct Key<double> {
  typedef uint64_t U;
  static constexpr int BITS = 64;
  static constexpr int POS_BITS = 63;
  __device__ static __forceinline__ U key(double v) {
    U b = (U)__double_as_longlong(v);
    return (b & 0x8000000000000000ull) ? ~b : (b | 0x8000000000000000ull);
  }
  __device__ static __forceinline__ double val(U k) {
    return __longlong_as_double(
        (long long)((k & 0x8000000000000000ull) ? (k & 0x7fffffffffffffffull) : ~k));
  }
  __device__ static __forceinline__ U bits(double v) { return (U)__double_as_longlong(v); }
};

// ------------------------------------------------------------------- streaming loads
__device__ __forceinline__ float4 ld_stream4(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p));
  return r;
}
__device__ __forceinline__ double2 ld_stream2(const double* p) {
  double2 r;
  asm volatile("ld.global.nc.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(r.x), "=d"(r.y) : "l"(p));
  return r;
}
__device__ __forceinline__ float ld_stream1(const float* p) {
  float r;
  asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ double ld_stream1(const double* p) {
  double r;
  asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}

// ------------------------------------------------------------------ block reductions
template <typename V, typename Op>
__device__ __forceinline__ V warp_reduce(V v, Op op) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = op(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// all threads get the result; scratch needs 32 entries
template <typename V, typename Op>
__device__ __forceinline__ V block_reduce(V v, Op op, V identity, V* scratch) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_reduce(v, op);
  __syncthreads();
  if (lane == 0) scratch[warp] = v;
  __syncthreads();
  V r = (lane < nw) ? scratch[lane] : identity;
  r = warp_reduce(r, op);
  return r;
}

// ------------------------------------------------- numpy percentile arithmetic in D
// numpy _get_indexes/_get_gamma for one percentile over n valid samples, arithmetic in T
template <typename T>
__device__ void percentile_ranks(long long n, double p, long long& lo, long long& hi, T& gamma) {
  const T q = div_rn((T)p, (T)100);
  const T nm1 = (T)(n - 1);
  const T v = mul_rn(nm1, q);
  if (v >= nm1) {  // above bounds: both neighbours are the last element
    lo = hi = n - 1;
    gamma = T(0);
  } else if (v < T(0)) {
    lo = hi = 0;
    gamma = T(0);
  } else if (is_nan(v)) {
    lo = hi = n - 1;
    gamma = T(0);
  } else {
    const T fl = floor(v);
    lo = (long long)fl;
    if (lo > n - 1) lo = n - 1;
    hi = lo + 1;
    if (hi > n - 1) hi = n - 1;
    gamma = sub_rn(v, fl);
  }
}

template <typename T>
__device__ T numpy_lerp(T a, T b, T g) {
  const T d = sub_rn(b, a);
  T r = add_rn(a, mul_rn(d, g));
  if (g >= T(0.5)) r = sub_rn(b, mul_rn(d, sub_rn(T(1), g)));
  return r;
}
