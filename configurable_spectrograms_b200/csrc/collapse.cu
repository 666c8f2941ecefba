// K1 -- masked pitch-angle-range segmented reduction.
//
// Replaces np.nansum(cube, axis=1) (CS/constants.py:12; call sites CS/plotting.py:188,
// CS/fast/plotting.py:128,278, CS/fast/extrema.py:259) plus the pitch-angle gather
// (CS/fast/plotting.py:121-127) and the zoom "any non-NaN" probe (CS/plotting.py:597-603):
// one pass over the cube emits the unmasked sum, every pitch-angle group sum and the
// per-time-row non-NaN flags.  HBM-bound: T*P*E*s bytes in, (G+1)*T*E*s out.
//
// Summation order is numpy's, bit for bit:
//   layout TPE : ascending-p chain seeded with +0.0, NaN -> +0.0
//   layout TEP : (total) 8-accumulator pairwise over the contiguous pitch axis, then "+0.0";
//                (groups) the gather copies to C order, so the ascending-p chain again.
#include <string.h>

#include "common.cuh"

namespace {

constexpr int kBlock = 256;

template <typename T>
struct VecOf;
template <>
struct VecOf<float> {
  static constexpr int N = 4;
};
template <>
struct VecOf<double> {
  static constexpr int N = 2;
};

template <typename T, int VEC>
struct Chunk {
  T v[VEC];
};

__device__ __forceinline__ void load_chunk(Chunk<float, 4>& c, const float* p, bool aligned, int valid) {
  if (aligned) {
    float4 r = ld_stream4(p);
    c.v[0] = r.x, c.v[1] = r.y, c.v[2] = r.z, c.v[3] = r.w;
  } else {
#pragma unroll
    for (int i = 0; i < 4; ++i) c.v[i] = i < valid ? ld_stream1(p + i) : CUDART_NAN_F;  // padding lanes: ignored like NaN
  }
}
__device__ __forceinline__ void load_chunk(Chunk<double, 2>& c, const double* p, bool aligned, int valid) {
  if (aligned) {
    double2 r = ld_stream2(p);
    c.v[0] = r.x, c.v[1] = r.y;
  } else {
    c.v[0] = ld_stream1(p);
    c.v[1] = valid > 1 ? ld_stream1(p + 1) : CUDART_NAN;
  }
}

__device__ __forceinline__ void store_chunk(float* p, const float (&a)[4], bool aligned, int valid) {
  if (aligned)
    *reinterpret_cast<float4*>(p) = make_float4(a[0], a[1], a[2], a[3]);
  else
    for (int i = 0; i < valid; ++i) p[i] = a[i];
}
__device__ __forceinline__ void store_chunk(double* p, const double (&a)[2], bool aligned, int valid) {
  if (aligned)
    *reinterpret_cast<double2*>(p) = make_double2(a[0], a[1]);
  else
    for (int i = 0; i < valid; ++i) p[i] = a[i];
}

__device__ __forceinline__ int find_file(const csg_file_desc* files, int n_files, int block) {
  int lo = 0, hi = n_files - 1;
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (__ldg(&files[mid].first_block) <= block)
      lo = mid;
    else
      hi = mid - 1;
  }
  return lo;
}

// ---------------------------------------------------------------------------------
// layout TPE: thread = (time row, VEC consecutive energies); loop over pitch bins.
// Consecutive threads read consecutive 16-byte pieces of one (t,p) energy row, so every
// warp request covers whole 128-byte lines; P independent 128-bit loads per thread.
// ---------------------------------------------------------------------------------
template <typename T, int NG>
__global__ void __launch_bounds__(kBlock)
    collapse_tpe_kernel(const csg_file_desc* __restrict__ files, int n_files,
                        const uint8_t* __restrict__ pa_bits, int n_groups, T* __restrict__ sums,
                        uint8_t* __restrict__ row_flags) {
  extern __shared__ unsigned char smem_raw[];
  unsigned* s_flags = reinterpret_cast<unsigned*>(smem_raw);  // [kBlock + 1]
  uint8_t* s_bits = smem_raw + (kBlock + 1) * sizeof(unsigned);  // [P]

  const int fi = find_file(files, n_files, blockIdx.x);
  const csg_file_desc f = files[fi];
  constexpr int VEC = VecOf<T>::N;
  const int P = f.P, E = f.E;
  const int EV = (E + VEC - 1) / VEC;  // energy chunks per row (the last one may be partial)
  const long long n_items = (long long)f.T * EV;
  const long long item0 = (long long)(blockIdx.x - f.first_block) * kBlock;
  const long long item = item0 + threadIdx.x;
  const int t_first = (int)(item0 / EV);

  for (int p = threadIdx.x; p < P; p += kBlock) s_bits[p] = NG ? pa_bits[f.bits_off + p] : 0;
  for (int i = threadIdx.x; i <= kBlock; i += kBlock) s_flags[i] = 0;
  __syncthreads();

  if (item < n_items) {
    const int t = (int)(item / EV);
    const int c = (int)(item - (long long)t * EV);
    const T* cube = static_cast<const T*>(f.d_cube);
    const T* ptr = cube + ((long long)t * P) * E + (long long)c * VEC;
    const int valid = (E - c * VEC) < VEC ? (E - c * VEC) : VEC;
    const bool aligned = (E % VEC == 0) && ((reinterpret_cast<uintptr_t>(cube) & 15) == 0);

    T acc[NG + 1][VEC];
#pragma unroll
    for (int g = 0; g <= NG; ++g)
#pragma unroll
      for (int v = 0; v < VEC; ++v) acc[g][v] = T(0);
    unsigned flag = 0;

    constexpr int U = 8;
    int p = 0;
    for (; p + U <= P; p += U) {
      Chunk<T, VEC> x[U];
#pragma unroll
      for (int u = 0; u < U; ++u) load_chunk(x[u], ptr + (long long)(p + u) * E, aligned, valid);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const unsigned bits = s_bits[p + u];
        const unsigned member = (bits << 1) | 1u;
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
          const T xv = x[u].v[v];
          const bool ok = !is_nan(xv);
          const T z = ok ? xv : T(0);
          flag |= ok ? member : 0u;
          acc[0][v] = add_rn(acc[0][v], z);
#pragma unroll
          for (int g = 0; g < NG; ++g)
            if ((bits >> g) & 1u) acc[g + 1][v] = add_rn(acc[g + 1][v], z);
        }
      }
    }
    for (; p < P; ++p) {
      Chunk<T, VEC> x;
      load_chunk(x, ptr + (long long)p * E, aligned, valid);
      const unsigned bits = s_bits[p];
      const unsigned member = (bits << 1) | 1u;
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const T xv = x.v[v];
        const bool ok = !is_nan(xv);
        const T z = ok ? xv : T(0);
        flag |= ok ? member : 0u;
        acc[0][v] = add_rn(acc[0][v], z);
#pragma unroll
        for (int g = 0; g < NG; ++g)
          if ((bits >> g) & 1u) acc[g + 1][v] = add_rn(acc[g + 1][v], z);
      }
    }

    T* out = sums + f.sums_off + (long long)t * E + (long long)c * VEC;
    const long long plane = (long long)f.T * E;
    const bool out_aligned = (E % VEC == 0) && ((reinterpret_cast<uintptr_t>(sums + f.sums_off) & 15) == 0) &&
                             (((plane * sizeof(T)) & 15) == 0);
#pragma unroll
    for (int g = 0; g <= NG; ++g)
      if (g <= n_groups) store_chunk(out + g * plane, acc[g], out_aligned, valid);
    if (flag) atomicOr(&s_flags[t - t_first], flag);
  }
  __syncthreads();
  if (row_flags != nullptr) {
    // rows touched by this block: t_first .. t_last (a row can straddle two blocks)
    const long long last_item = (item0 + kBlock < n_items ? item0 + kBlock : n_items) - 1;
    const int n_rows = (int)(last_item / EV) - t_first + 1;
    for (int r = threadIdx.x; r < n_rows; r += kBlock) {
      const unsigned fl = s_flags[r];
      if (fl) {
        // byte-wide OR through the containing aligned word
        uint8_t* dst = row_flags + f.flags_off + t_first + r;
        const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
        unsigned* word = reinterpret_cast<unsigned*>(a & ~uintptr_t(3));
        atomicOr(word, (fl & 0xffu) << (8 * (a & 3)));
      }
    }
  }
}

// numpy pairwise_sum over n contiguous values (NaN already treated as 0), n >= 0
template <typename T, typename Load>
__device__ T pairwise_sum(Load load, int first, int n) {
  if (n < 8) {
    T r = T(-0.0);
    for (int i = 0; i < n; ++i) r = add_rn(r, load(first + i));
    return r;
  }
  if (n <= 128) {
    T r[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) r[k] = load(first + k);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
      for (int k = 0; k < 8; ++k) r[k] = add_rn(r[k], load(first + i + k));
    }
    T res = add_rn(add_rn(add_rn(r[0], r[1]), add_rn(r[2], r[3])),
                   add_rn(add_rn(r[4], r[5]), add_rn(r[6], r[7])));
    for (; i < n; ++i) res = add_rn(res, load(first + i));
    return res;
  }
  int n2 = n / 2;
  n2 -= n2 % 8;
  return add_rn(pairwise_sum<T>(load, first, n2), pairwise_sum<T>(load, first + n2, n - n2));
}

// ---------------------------------------------------------------------------------
// layout TEP: rows of P contiguous pitch samples.  A block stages ROWS rows in shared
// memory with coalesced 128-bit loads (row pitch P+1 words -> conflict-free row walks),
// then one thread per (t,e) row runs numpy's pairwise order for the total and the
// ascending-p chain for each pitch-angle group.
// ---------------------------------------------------------------------------------
template <typename T, int NG>
__global__ void collapse_tep_kernel(const csg_file_desc* __restrict__ files, int n_files,
                                    const uint8_t* __restrict__ pa_bits, int n_groups,
                                    T* __restrict__ sums, uint8_t* __restrict__ row_flags, int rows_per_block,
                                    int max_P) {
  extern __shared__ unsigned char smem_raw[];
  T* s_rows = reinterpret_cast<T*>(smem_raw);  // [rows_per_block][max_P + 1]
  uint8_t* s_bits = smem_raw + (size_t)rows_per_block * (max_P + 1) * sizeof(T);

  const int fi = find_file(files, n_files, blockIdx.x);
  const csg_file_desc f = files[fi];
  const int P = f.P, E = f.E;
  const long long n_rows = (long long)f.T * E;
  const long long row0 = (long long)(blockIdx.x - f.first_block) * rows_per_block;
  const int rows_here = (int)(n_rows - row0 < rows_per_block ? n_rows - row0 : rows_per_block);
  const T* cube = static_cast<const T*>(f.d_cube) + row0 * P;
  const int pitch = P + 1;

  for (int p = threadIdx.x; p < P; p += blockDim.x) s_bits[p] = NG ? pa_bits[f.bits_off + p] : 0;
  const long long n_elem = (long long)rows_here * P;
  constexpr int VEC = VecOf<T>::N;
  const bool aligned = ((reinterpret_cast<uintptr_t>(cube) & 15) == 0);
  if (aligned) {
    const long long n_vec = n_elem / VEC;
    for (long long i = threadIdx.x; i < n_vec; i += blockDim.x) {
      Chunk<T, VEC> x;
      load_chunk(x, cube + i * VEC, true, VEC);
#pragma unroll
      for (int v = 0; v < VEC; ++v) {
        const long long j = i * VEC + v;
        const int r = (int)(j / P), p = (int)(j - (long long)r * P);
        s_rows[r * pitch + p] = x.v[v];
      }
    }
    for (long long j = n_vec * VEC + threadIdx.x; j < n_elem; j += blockDim.x) {
      const int r = (int)(j / P), p = (int)(j - (long long)r * P);
      s_rows[r * pitch + p] = ld_stream1(cube + j);
    }
  } else {
    for (long long j = threadIdx.x; j < n_elem; j += blockDim.x) {
      const int r = (int)(j / P), p = (int)(j - (long long)r * P);
      s_rows[r * pitch + p] = ld_stream1(cube + j);
    }
  }
  __syncthreads();

  const int r = threadIdx.x;
  if (r < rows_here) {
    const T* row = s_rows + r * pitch;
    unsigned flag = 0;
    T acc[NG + 1];
#pragma unroll
    for (int g = 0; g <= NG; ++g) acc[g] = T(0);
    for (int p = 0; p < P; ++p) {
      const T xv = row[p];
      const bool ok = !is_nan(xv);
      const unsigned bits = s_bits[p];
      flag |= ok ? ((bits << 1) | 1u) : 0u;
      const T z = ok ? xv : T(0);
#pragma unroll
      for (int g = 0; g < NG; ++g)
        if ((bits >> g) & 1u) acc[g + 1] = add_rn(acc[g + 1], z);
    }
    auto load = [row](int i) {
      const T xv = row[i];
      return is_nan(xv) ? T(0) : xv;
    };
    acc[0] = add_rn(T(0), pairwise_sum<T>(load, 0, P));  // reduction seeded with +0.0

    const long long grow = row0 + r;  // = t*E + e
    T* out = sums + f.sums_off + grow;
#pragma unroll
    for (int g = 0; g <= NG; ++g)
      if (g <= n_groups) out[g * n_rows] = acc[g];
    if (row_flags != nullptr && flag) {
      uint8_t* dst = row_flags + f.flags_off + (grow / E);
      const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
      atomicOr(reinterpret_cast<unsigned*>(a & ~uintptr_t(3)), (flag & 0xffu) << (8 * (a & 3)));
    }
  }
}

// warp = one zoom window: any row with the group's bit set?
__global__ void window_any_kernel(const uint8_t* __restrict__ row_flags, const csg_flag_window* __restrict__ windows,
                                  int n_windows, const int32_t* __restrict__ pool, uint8_t* __restrict__ out) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= n_windows) return;
  const csg_flag_window win = windows[w];
  const uint8_t* fl = row_flags + win.flags_off;
  bool any = false;
  for (int i = lane; i < win.nt; i += 32) {
    const int row = win.rows_off < 0 ? win.t0 + i : __ldg(pool + win.rows_off + i);
    any |= ((fl[row] >> win.bit) & 1u) != 0;
  }
  any = __any_sync(0xffffffffu, any);
  if (lane == 0) out[w] = any ? 1 : 0;
}

inline int tep_rows_per_block(int P, int dtype) {
  const size_t es = dtype == CSG_F64 ? 8 : 4;
  int rows = 256;
  while (rows > 32 && (size_t)rows * (P + 1) * es > 96 * 1024) rows >>= 1;
  return rows;
}

template <typename T>
int launch_tpe(csg_ctx* ctx, const csg_file_desc* d_files, int n_files, int total_blocks,
               const uint8_t* d_pa_bits, int n_groups, int max_P, T* d_sums, uint8_t* d_row_flags) {
  const size_t smem = (kBlock + 1) * sizeof(unsigned) + (size_t)max_P;
  auto go = [&](auto kern) {
    kern<<<total_blocks, kBlock, smem, ctx->stream>>>(d_files, n_files, d_pa_bits, n_groups, d_sums, d_row_flags);
  };
  if (n_groups == 0)
    go(collapse_tpe_kernel<T, 0>);
  else if (n_groups <= 4)
    go(collapse_tpe_kernel<T, 4>);
  else
    go(collapse_tpe_kernel<T, CSG_MAX_GROUPS>);
  CSG_LAUNCH_CHECK(ctx, "collapse_tpe_kernel");
  return CSG_OK;
}

template <typename T>
int launch_tep(csg_ctx* ctx, const csg_file_desc* d_files, int n_files, int total_blocks,
               const uint8_t* d_pa_bits, int n_groups, int max_P, int dtype, T* d_sums, uint8_t* d_row_flags) {
  const int rows = tep_rows_per_block(max_P, dtype);
  const size_t smem = (size_t)rows * (max_P + 1) * sizeof(T) + (size_t)max_P;
  if (smem > 200 * 1024)
    return csg_fail(ctx, CSG_ERR_ARG, "layout TEP supports at most %d pitch bins per row for this dtype (got %d)",
                    (int)(200 * 1024 / (32 * sizeof(T))) - 2, max_P);
  auto go = [&](auto kern) {
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    kern<<<total_blocks, rows, smem, ctx->stream>>>(d_files, n_files, d_pa_bits, n_groups, d_sums, d_row_flags, rows, max_P);
  };
  if (n_groups == 0)
    go(collapse_tep_kernel<T, 0>);
  else if (n_groups <= 4)
    go(collapse_tep_kernel<T, 4>);
  else
    go(collapse_tep_kernel<T, CSG_MAX_GROUPS>);
  CSG_LAUNCH_CHECK(ctx, "collapse_tep_kernel");
  return CSG_OK;
}

}  // namespace

extern "C" {

int32_t csg_collapse_blocks(int32_t T, int32_t P, int32_t E, int dtype, int layout) {
  if (T <= 0 || E <= 0) return 0;
  if (layout == CSG_LAYOUT_TEP) {
    const int rows = tep_rows_per_block(P, dtype);
    return (int32_t)(((long long)T * E + rows - 1) / rows);
  }
  const int vec = dtype == CSG_F64 ? 2 : 4;
  const long long items = (long long)T * ((E + vec - 1) / vec);
  return (int32_t)((items + kBlock - 1) / kBlock);
}

int csg_collapse(csg_ctx* ctx, const csg_file_desc* d_files, int n_files, int total_blocks,
                 const uint8_t* d_pa_bits, int n_groups, int max_P, int dtype, int layout, void* d_sums,
                 uint8_t* d_row_flags) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_groups < 0 || n_groups > CSG_MAX_GROUPS) return csg_fail(ctx, CSG_ERR_ARG, "n_groups %d out of range", n_groups);
  if (n_groups > 0 && !d_pa_bits) return csg_fail(ctx, CSG_ERR_ARG, "d_pa_bits is NULL with n_groups > 0");
  if (dtype != CSG_F32 && dtype != CSG_F64) return csg_fail(ctx, CSG_ERR_ARG, "bad dtype %d", dtype);
  if (layout != CSG_LAYOUT_TPE && layout != CSG_LAYOUT_TEP) return csg_fail(ctx, CSG_ERR_ARG, "bad layout %d", layout);
  if (n_files <= 0 || total_blocks <= 0) return CSG_OK;
  if (!d_files || !d_sums) return csg_fail(ctx, CSG_ERR_ARG, "NULL table or output");
  if (max_P <= 0 || max_P > 32768) return csg_fail(ctx, CSG_ERR_ARG, "max_P %d out of range (1..32768)", max_P);
  if (layout == CSG_LAYOUT_TPE) {
    if (dtype == CSG_F32)
      return launch_tpe<float>(ctx, d_files, n_files, total_blocks, d_pa_bits, n_groups, max_P, (float*)d_sums, d_row_flags);
    return launch_tpe<double>(ctx, d_files, n_files, total_blocks, d_pa_bits, n_groups, max_P, (double*)d_sums, d_row_flags);
  }
  if (dtype == CSG_F32)
    return launch_tep<float>(ctx, d_files, n_files, total_blocks, d_pa_bits, n_groups, max_P, dtype, (float*)d_sums, d_row_flags);
  return launch_tep<double>(ctx, d_files, n_files, total_blocks, d_pa_bits, n_groups, max_P, dtype, (double*)d_sums, d_row_flags);
}

int csg_window_any(csg_ctx* ctx, const uint8_t* d_row_flags, const csg_flag_window* d_windows, int n_windows,
                   const int32_t* d_index_pool, uint8_t* d_out) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_windows <= 0) return CSG_OK;
  if (!d_row_flags || !d_windows || !d_out) return csg_fail(ctx, CSG_ERR_ARG, "NULL argument");
  const int blocks = (n_windows * 32 + 255) / 256;
  window_any_kernel<<<blocks, 256, 0, ctx->stream>>>(d_row_flags, d_windows, n_windows, d_index_pool, d_out);
  CSG_LAUNCH_CHECK(ctx, "window_any_kernel");
  return CSG_OK;
}

int csg_collapse_host(csg_ctx* ctx, const void* h_cube, int32_t T, int32_t P, int32_t E, int dtype, int layout,
                      const uint8_t* h_pa_bits, int n_groups, void* h_sums, uint8_t* h_row_flags) {
  if (!ctx) return CSG_ERR_ARG;
  if (T < 0 || P <= 0 || E <= 0) return csg_fail(ctx, CSG_ERR_ARG, "bad cube shape (%d,%d,%d)", T, P, E);
  if (dtype != CSG_F32 && dtype != CSG_F64) return csg_fail(ctx, CSG_ERR_ARG, "bad dtype %d", dtype);
  if (T == 0) return CSG_OK;
  const size_t es = dtype == CSG_F64 ? 8 : 4;
  const size_t cube_bytes = (size_t)T * P * E * es;
  const size_t sums_bytes = (size_t)(n_groups + 1) * T * E * es;
  const size_t flag_bytes = ((size_t)T + 3) & ~size_t(3);
  void *d_cube = nullptr, *d_sums = nullptr, *d_flags = nullptr, *d_bits = nullptr, *d_desc = nullptr;
  int st = CSG_OK;
  auto cleanup = [&]() {
    cudaStreamSynchronize(ctx->stream);
    cudaFree(d_cube), cudaFree(d_sums), cudaFree(d_flags), cudaFree(d_bits), cudaFree(d_desc);
  };
#define CSG_TRY(x)            \
  if ((st = (x)) != CSG_OK) { \
    cleanup();                \
    return st;                \
  }
  CSG_TRY(csg_dev_alloc(ctx, cube_bytes, &d_cube));
  CSG_TRY(csg_dev_alloc(ctx, sums_bytes, &d_sums));
  CSG_TRY(csg_dev_alloc(ctx, flag_bytes, &d_flags));
  CSG_TRY(csg_dev_alloc(ctx, (size_t)P, &d_bits));
  CSG_TRY(csg_dev_alloc(ctx, sizeof(csg_file_desc), &d_desc));
  csg_file_desc desc;
  memset(&desc, 0, sizeof(desc));
  desc.d_cube = d_cube;
  desc.T = T, desc.P = P, desc.E = E;
  const int blocks = csg_collapse_blocks(T, P, E, dtype, layout);
  CSG_TRY(csg_h2d(ctx, d_cube, h_cube, cube_bytes));
  CSG_TRY(csg_h2d(ctx, d_desc, &desc, sizeof(desc)));
  if (n_groups > 0) CSG_TRY(csg_h2d(ctx, d_bits, h_pa_bits, (size_t)P));
  CSG_TRY(csg_memset(ctx, d_flags, 0, flag_bytes));
  CSG_TRY(csg_collapse(ctx, (const csg_file_desc*)d_desc, 1, blocks, (const uint8_t*)d_bits, n_groups, P, dtype, layout,
                       d_sums, (uint8_t*)d_flags));
  CSG_TRY(csg_d2h(ctx, h_sums, d_sums, sums_bytes));
  if (h_row_flags) CSG_TRY(csg_d2h(ctx, h_row_flags, d_flags, (size_t)T));
  CSG_TRY(csg_sync(ctx));
#undef CSG_TRY
  cleanup();
  return CSG_OK;
}

}  // extern "C"
