// K1 -- masked pitch-angle-range segmented reduction.
//
// Replaces np.nansum(cube, axis=1) (CS/constants.py:12; call sites CS/plotting.py:188,
// CS/fast/plotting.py:128,278, CS/fast/extrema.py:259) plus the pitch-angle gather
// (CS/fast/plotting.py:121-127) and the zoom "any non-NaN" probe (CS/plotting.py:597-603):
// one pass over the cube emits the unmasked sum, every pitch-angle group sum and the
// per-time-row non-NaN flags.  HBM-bound: T*P*E*s bytes in, (G+1)*T*E*s out.
//
// Output layout is ENERGY-MAJOR: sums[file][g][e][t] with row pitch Tp = T rounded up to a
// multiple of 4 -- the orientation of matrix_plot = collapsed.T (CS/plotting.py:236) that
// imshow draws, so K2a / K2b / K3 stream contiguous time runs and the rasteriser needs no
// transpose.
//
// Summation order is numpy's, bit for bit:
//   layout TPE : ascending-p chain seeded with +0.0, NaN -> +0.0
//   layout TEP : (total) 8-accumulator pairwise over the contiguous pitch axis, then "+0.0";
//                (groups) the gather copies to C order, so the ascending-p chain again.
//
// Kernels
//   collapse_slab_kernel   (TPE, the FAST shapes) one time step of a (T,P,E) cube is one contiguous
//       P*E slab: every warp streams its time rows through a private ring of shared-memory
//       stages filled by cp.async.bulk (TMA, mbarrier complete_tx) -- bytes in flight do not
//       cost registers -- and walks the pitch axis in runs of constant group membership with a
//       loop body specialised per membership mask; a CTA covers 16 time rows so the transposed
//       result leaves as 64-byte row segments.
//   collapse_tpe_kernel    (TPE, any shape/alignment) register-staged generic path.
//   collapse_tep_kernel    (stored (T,E,P) view).
#include <stdlib.h>
#include <string.h>

#include <vector>

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

__device__ __forceinline__ int pitch_of(int T) { return (T + 3) & ~3; }

__device__ __forceinline__ void or_flag_byte(uint8_t* dst, unsigned fl) {
  // byte-wide OR through the containing aligned word
  const uintptr_t a = reinterpret_cast<uintptr_t>(dst);
  atomicOr(reinterpret_cast<unsigned*>(a & ~uintptr_t(3)), (fl & 0xffu) << (8 * (a & 3)));
}

// ---------------------------------------------------------------------------------
// TMA / mbarrier primitives (sm_90+ PTX; SASS: UBLKCP, SYNCS)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// global -> shared bulk copy, completion counted in bytes on the mbarrier; the cube is read
// exactly once, so it is marked evict-first in L2 (the sums written next should stay there)
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, uint64_t* bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
          smem_u32(dst)),
      "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}

// ---------------------------------------------------------------------------------
// slab kernel: persistent CTAs, one per SM, each owning a contiguous range of 16-row tiles.
// Every warp streams ITS time rows (row w, w+W, ... of each tile) through a private ring of
// shared-memory stages; lane 0 keeps the ring full with bulk copies that run ahead across
// tile boundaries, all lanes consume.  A tile's transposed result is staged in shared
// memory and leaves as 64-byte row segments.
// ---------------------------------------------------------------------------------
constexpr int kTileRows = 16;
constexpr int kOutPitch = kTileRows + 1;
constexpr int kSlabMaxWarps = 8;

template <typename T, int V>
struct LdsVec;
template <>
struct LdsVec<float, 4> {
  static __device__ __forceinline__ void load(const float* p, float (&x)[4]) {
    const float4 r = *reinterpret_cast<const float4*>(p);
    x[0] = r.x, x[1] = r.y, x[2] = r.z, x[3] = r.w;
  }
};
template <>
struct LdsVec<double, 2> {
  static __device__ __forceinline__ void load(const double* p, double (&x)[2]) {
    const double2 r = *reinterpret_cast<const double2*>(p);
    x[0] = r.x, x[1] = r.y;
  }
};

// one run of pitch bins [p0, p1) of a staged chunk (row pitch E), constant membership MASK
// (MASK < 0: membership read per bin from s_bits -- the generic body for more than 4 groups)
template <typename T, int NG, int NJ, int MASK>
__device__ __forceinline__ void slab_run(const T* __restrict__ stage, int p0, int p1, int E, int EV, int lane,
                                         const uint8_t* __restrict__ s_bits, int p_base, unsigned keep,
                                         T (&acc)[NJ][NG + 1][VecOf<T>::N], unsigned& flag) {
  constexpr int V = VecOf<T>::N;
  bool any = false;
  unsigned any_bits = 0;
#pragma unroll 4
  for (int p = p0; p < p1; ++p) {
    const unsigned bits = MASK < 0 ? (s_bits[p_base + p] & keep) : (unsigned)MASK;
    bool row_ok = false;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      // lanes past the last energy chunk redo the last one (results discarded): no divergence
      const int c = min(lane + 32 * j, EV - 1);
      T x[V];
      LdsVec<T, V>::load(stage + (size_t)p * E + c * V, x);
#pragma unroll
      for (int v = 0; v < V; ++v) {
        const bool ok = !is_nan(x[v]);
        const T z = ok ? x[v] : T(0);
        row_ok |= ok;
        acc[j][0][v] = add_rn(acc[j][0][v], z);
#pragma unroll
        for (int g = 0; g < NG; ++g)
          if ((bits >> g) & 1u) acc[j][g + 1][v] = add_rn(acc[j][g + 1][v], z);
      }
    }
    any |= row_ok;
    if (MASK < 0) any_bits |= row_ok ? ((bits << 1) | 1u) : 0u;
  }
  if (MASK < 0)
    flag |= any_bits;
  else
    flag |= any ? (((unsigned)MASK << 1) | 1u) : 0u;
}

template <typename T, int NG, int NJ>
__global__ void __launch_bounds__(kSlabMaxWarps * 32, 1)
    collapse_slab_kernel(const csg_file_desc* __restrict__ files, int n_files, int total_tiles,
                         const uint8_t* __restrict__ pa_bits, int n_groups, T* __restrict__ sums,
                         uint8_t* __restrict__ row_flags, int pc, int ns, int stage_elems, int max_P, int max_E) {
  constexpr int V = VecOf<T>::N;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  const int W = blockDim.x >> 5;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // carve shared memory: stages | output tile | barriers | runs | membership bytes
  T* s_stage = reinterpret_cast<T*>(smem_raw);
  size_t off = (size_t)W * ns * stage_elems * sizeof(T);
  T* s_out = reinterpret_cast<T*>(smem_raw + off);
  off += (size_t)(NG + 1) * max_E * kOutPitch * sizeof(T);
  off = (off + 7) & ~size_t(7);
  uint64_t* s_bar = reinterpret_cast<uint64_t*>(smem_raw + off);
  off += (size_t)W * ns * sizeof(uint64_t);
  int* s_runs = reinterpret_cast<int*>(smem_raw + off);  // {p0, p1, mask} triples; [3*max_P] = count, [+1] = alias
  off += (size_t)(3 * max_P + 2) * sizeof(int);
  uint8_t* s_bits = smem_raw + off;

  if (threadIdx.x < W * ns) mbar_init(&s_bar[threadIdx.x], 1);
  if (threadIdx.x == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();

  T* my_stage = s_stage + (size_t)warp * ns * stage_elems;
  uint64_t* my_bar = s_bar + warp * ns;
  uint64_t policy = 0;
  if (lane == 0) policy = policy_evict_first();
  unsigned ring = 0;  // stages this warp has consumed since the kernel started (slot / parity of the ring)

  // this CTA's contiguous range of tiles
  const int per_cta = (total_tiles + gridDim.x - 1) / gridDim.x;
  int tile = blockIdx.x * per_cta;
  const int tile_end = min(total_tiles, tile + per_cta);

  while (tile < tile_end) {
    const int fi = find_file(files, n_files, tile);
    const csg_file_desc f = files[fi];
    const int P = f.P, E = f.E, EV = E / V;
    const int file_tile0 = tile - f.first_block;
    const int file_tiles = (f.T + kTileRows - 1) / kTileRows;
    const int n_seg = min(tile_end - tile, file_tiles - file_tile0);  // tiles of this file in this CTA

    // ---- per-file tables: membership bytes, runs of constant membership, groups that hold every bin
    __syncthreads();
    for (int p = threadIdx.x; p < P; p += blockDim.x) s_bits[p] = NG ? pa_bits[f.bits_off + p] : 0;
    __syncthreads();
    if (threadIdx.x == 0) {
      unsigned all = 0xffu;
      for (int p = 0; p < P; ++p) all &= s_bits[p];
      int n = 0, start = 0;
      for (int p = 1; p <= P; ++p)
        if (p == P || s_bits[p] != s_bits[start]) {
          s_runs[3 * n] = start, s_runs[3 * n + 1] = p, s_runs[3 * n + 2] = s_bits[start] & ~all;
          ++n, start = p;
        }
      s_runs[3 * max_P] = n;
      s_runs[3 * max_P + 1] = (int)all;  // these groups equal the unmasked total: copied, not summed
    }
    __syncthreads();
    const int n_runs = s_runs[3 * max_P];
    const unsigned alias = (unsigned)s_runs[3 * max_P + 1];
    const unsigned keep = ~alias;

    const int n_ch = (P + pc - 1) / pc;  // stages per time row
    const int nfull = (kTileRows - warp + W - 1) / W;  // this warp's rows in a full tile
    auto rows_in = [&](int kk) {  // this warp's rows in tile kk of the segment (only a file's last tile is short)
      const int here = min(kTileRows, f.T - (file_tile0 + kk) * kTileRows);
      return warp < here ? (here - warp + W - 1) / W : 0;
    };
    const int total = (nfull * (n_seg - 1) + rows_in(n_seg - 1)) * n_ch;
    const T* cube = static_cast<const T*>(f.d_cube);

    auto issue = [&](int q) {  // lane 0: arm the stage's barrier and start its bulk copy
      const int ri = q / n_ch, ch = q - ri * n_ch;
      const int kk = ri / nfull, i = ri - kk * nfull;
      const int t = (file_tile0 + kk) * kTileRows + warp + W * i;
      const int pcn = min(pc, P - ch * pc);
      const unsigned bytes = (unsigned)((size_t)pcn * E * sizeof(T));
      const unsigned slot = (ring + (unsigned)q) % (unsigned)ns;
      mbar_expect_tx(&my_bar[slot], bytes);
      bulk_g2s(my_stage + (size_t)slot * stage_elems, cube + ((size_t)t * P + (size_t)ch * pc) * E, bytes, &my_bar[slot],
               policy);
    };
    if (lane == 0)
      for (int q = 0; q < ns - 1 && q < total; ++q) issue(q);

    int s = 0;  // stages consumed in this segment
    for (int kk = 0; kk < n_seg; ++kk) {
      const int t_base = (file_tile0 + kk) * kTileRows;
      const int rows_here = min(kTileRows, f.T - t_base);
      const int n_i = rows_in(kk);
      for (int i = 0; i < n_i; ++i) {
        T acc[NJ][NG + 1][V];
#pragma unroll
        for (int j = 0; j < NJ; ++j)
#pragma unroll
          for (int g = 0; g <= NG; ++g)
#pragma unroll
            for (int v = 0; v < V; ++v) acc[j][g][v] = T(0);
        unsigned flag = 0;
        for (int ch = 0; ch < n_ch; ++ch, ++s) {
          if (lane == 0 && s + ns - 1 < total) issue(s + ns - 1);  // its slot was drained in the previous iteration
          const unsigned g = ring + (unsigned)s;
          const unsigned slot = g % (unsigned)ns;
          mbar_wait(&my_bar[slot], (g / (unsigned)ns) & 1u);
          const T* stage = my_stage + (size_t)slot * stage_elems;
          const int p_lo = ch * pc, p_hi = min(P, p_lo + pc);
          for (int r = 0; r < n_runs; ++r) {
            const int a = max(s_runs[3 * r], p_lo), b = min(s_runs[3 * r + 1], p_hi);
            if (a >= b) continue;
            const int mask = s_runs[3 * r + 2];
            const int a0 = a - p_lo, b0 = b - p_lo;
            if (NG == 4) {
              switch (mask & 15) {
#define CSG_RUN(M)                                                                      \
  case M:                                                                               \
    slab_run<T, NG, NJ, M>(stage, a0, b0, E, EV, lane, s_bits, p_lo, keep, acc, flag);  \
    break;
                CSG_RUN(0)
                CSG_RUN(1)
                CSG_RUN(2)
                CSG_RUN(3)
                CSG_RUN(4)
                CSG_RUN(5)
                CSG_RUN(6)
                CSG_RUN(7)
                CSG_RUN(8)
                CSG_RUN(9)
                CSG_RUN(10)
                CSG_RUN(11)
                CSG_RUN(12)
                CSG_RUN(13)
                CSG_RUN(14)
                CSG_RUN(15)
#undef CSG_RUN
              }
            } else if (NG == 0) {
              slab_run<T, NG, NJ, 0>(stage, a0, b0, E, EV, lane, s_bits, p_lo, keep, acc, flag);
            } else {
              slab_run<T, NG, NJ, -1>(stage, a0, b0, E, EV, lane, s_bits, p_lo, keep, acc, flag);
            }
          }
          __syncwarp();  // every lane is done with this stage before lane 0 re-arms it
        }
        // ---- the row is complete: groups that hold every bin copy the total
        const int row = warp + W * i;
        if (flag & 1u) flag |= alias << 1;
#pragma unroll
        for (int j = 0; j < NJ; ++j) {
          const int c = lane + 32 * j;
          if (c < EV) {
#pragma unroll
            for (int g = 0; g <= NG; ++g) {
              const bool copy = g > 0 && ((alias >> (g - 1)) & 1u);
#pragma unroll
              for (int v = 0; v < V; ++v)
                s_out[((size_t)g * E + c * V + v) * kOutPitch + row] = copy ? acc[j][0][v] : acc[j][g][v];
            }
          }
        }
        const unsigned fl = __reduce_or_sync(0xffffffffu, flag);
        if (lane == 0 && row_flags != nullptr) row_flags[f.flags_off + t_base + row] = (uint8_t)fl;  // sole owner
      }
      __syncthreads();
      // ---- transposed tile -> sums[g][e][t_base .. t_base+rows_here): 64-byte segments
      {
        const int Tp = pitch_of(f.T);
        const long long plane = (long long)E * Tp;
        T* out = sums + f.sums_off + t_base;
        const int r = threadIdx.x & (kTileRows - 1);
        const int step = blockDim.x / kTileRows;
        const int n_ge = (n_groups + 1) * E;
        int ge = threadIdx.x / kTileRows;
        int g = ge / E, e = ge - g * E;
        for (; ge < n_ge; ge += step) {
          if (r < rows_here) out[g * plane + (long long)e * Tp + r] = s_out[(size_t)ge * kOutPitch + r];
          e += step;
          while (e >= E) e -= E, ++g;
        }
      }
      __syncthreads();
    }
    ring += (unsigned)total;
    tile += n_seg;
  }
}

// ---------------------------------------------------------------------------------
// layout TPE, generic: thread = (time row, VEC consecutive energies); loop over pitch bins.
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

    // energy-major output: element (e, t) at e*Tp + t (uncoalesced here; the slab kernel is the fast path)
    const int Tp = pitch_of(f.T);
    const long long plane = (long long)E * Tp;
    T* out = sums + f.sums_off + t;
#pragma unroll
    for (int g = 0; g <= NG; ++g)
      if (g <= n_groups)
#pragma unroll
        for (int v = 0; v < VEC; ++v)
          if (v < valid) out[g * plane + (long long)(c * VEC + v) * Tp] = acc[g][v];
    if (flag) atomicOr(&s_flags[t - t_first], flag);
  }
  __syncthreads();
  if (row_flags != nullptr) {
    // rows touched by this block: t_first .. t_last (a row can straddle two blocks)
    const long long last_item = (item0 + kBlock < n_items ? item0 + kBlock : n_items) - 1;
    const int n_rows = (int)(last_item / EV) - t_first + 1;
    for (int r = threadIdx.x; r < n_rows; r += kBlock) {
      const unsigned fl = s_flags[r];
      if (fl) or_flag_byte(row_flags + f.flags_off + t_first + r, fl);
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
    const int t = (int)(grow / E), e = (int)(grow - (long long)t * E);
    const int Tp = pitch_of(f.T);
    const long long plane = (long long)E * Tp;
    T* out = sums + f.sums_off + (long long)e * Tp + t;
#pragma unroll
    for (int g = 0; g <= NG; ++g)
      if (g <= n_groups) out[g * plane] = acc[g];
    if (row_flags != nullptr && flag) or_flag_byte(row_flags + f.flags_off + t, flag);
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

// ---- slab kernel configuration: warps per CTA, pitch bins per stage, stages per warp, shared memory
struct SlabCfg {
  int warps, pc, ns, stage_elems, nj;
  size_t smem;
  bool ok;
};
inline SlabCfg slab_config(int max_P, int max_E, int n_groups, int dtype) {
  SlabCfg c{};
  const size_t es = dtype == CSG_F64 ? 8 : 4;
  const int V = (int)(16 / es);
  c.ok = false;
  if (max_E % V != 0 || max_P <= 0 || max_E <= 0) return c;
  c.nj = (max_E / V + 31) / 32;
  if (c.nj > 3) return c;
  const int NG = n_groups == 0 ? 0 : (n_groups <= 4 ? 4 : CSG_MAX_GROUPS);
  const size_t fixed = (size_t)(NG + 1) * max_E * kOutPitch * es + 16 + (size_t)(3 * max_P + 2) * 4 + max_P + 64;
  const size_t budget = 224 * 1024;  // one persistent CTA per SM
  const size_t row = (size_t)max_E * es;
  int warps = 8, ns = 3;
  int pc = (int)(6144 / row);  // about 6 KB per stage
  if (pc < 1) pc = 1;
  if (pc > max_P) pc = max_P;
  // tuning overrides (scripts/k1_sweep.py)
  if (const char* e = getenv("CSG_SLAB_W")) warps = atoi(e) >= 1 && atoi(e) <= kSlabMaxWarps ? atoi(e) : warps;
  if (const char* e = getenv("CSG_SLAB_PC")) pc = atoi(e) > 0 ? (atoi(e) < max_P ? atoi(e) : max_P) : pc;
  if (const char* e = getenv("CSG_SLAB_NS")) ns = atoi(e) >= 2 ? atoi(e) : ns;
  auto total = [&](int pc_, int ns_) { return fixed + (size_t)warps * ns_ * (pc_ * row + 8); };
  while (total(pc, ns) > budget && pc > 1) pc = (pc + 1) / 2;
  if (total(pc, ns) > budget) ns = 2;
  if (total(pc, ns) > budget) return c;
  c.warps = warps, c.pc = pc, c.ns = ns, c.stage_elems = (int)(pc * (size_t)max_E), c.smem = total(pc, ns), c.ok = true;
  return c;
}

inline bool slab_file_ok(int32_t T, int32_t P, int32_t E, int dtype, const void* d_cube) {
  const size_t es = dtype == CSG_F64 ? 8 : 4;
  const int V = (int)(16 / es);
  if (T <= 0 || P <= 0 || E % V != 0) return false;
  if ((reinterpret_cast<uintptr_t>(d_cube) & 15) != 0) return false;
  if ((E / V + 31) / 32 > 3) return false;
  return true;
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

template <typename T, int NG>
int launch_slab_nj(csg_ctx* ctx, const SlabCfg& c, const csg_file_desc* d_files, int n_files, int total_blocks,
                   const uint8_t* d_pa_bits, int n_groups, int max_P, int max_E, T* d_sums, uint8_t* d_row_flags) {
  auto go = [&](auto kern) -> int {
    CSG_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c.smem));
    const int grid = total_blocks < ctx->sm_count ? total_blocks : ctx->sm_count;  // persistent: one CTA per SM
    kern<<<grid, c.warps * 32, c.smem, ctx->stream>>>(d_files, n_files, total_blocks, d_pa_bits, n_groups, d_sums,
                                                       d_row_flags, c.pc, c.ns, c.stage_elems, max_P, max_E);
    return CSG_OK;
  };
  int st;
  if (c.nj == 1)
    st = go(collapse_slab_kernel<T, NG, 1>);
  else if (c.nj == 2)
    st = go(collapse_slab_kernel<T, NG, 2>);
  else
    st = go(collapse_slab_kernel<T, NG, 3>);
  if (st != CSG_OK) return st;
  CSG_LAUNCH_CHECK(ctx, "collapse_slab_kernel");
  return CSG_OK;
}

template <typename T>
int launch_slab(csg_ctx* ctx, const SlabCfg& c, const csg_file_desc* d_files, int n_files, int total_blocks,
                const uint8_t* d_pa_bits, int n_groups, int max_P, int max_E, T* d_sums, uint8_t* d_row_flags) {
  if (n_groups == 0)
    return launch_slab_nj<T, 0>(ctx, c, d_files, n_files, total_blocks, d_pa_bits, n_groups, max_P, max_E, d_sums,
                                d_row_flags);
  if (n_groups <= 4)
    return launch_slab_nj<T, 4>(ctx, c, d_files, n_files, total_blocks, d_pa_bits, n_groups, max_P, max_E, d_sums,
                                d_row_flags);
  return launch_slab_nj<T, CSG_MAX_GROUPS>(ctx, c, d_files, n_files, total_blocks, d_pa_bits, n_groups, max_P, max_E, d_sums,
                                           d_row_flags);
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

int csg_collapse_kernel(int32_t T, int32_t P, int32_t E, int dtype, int layout, const void* d_cube) {
  if (layout == CSG_LAYOUT_TEP) return CSG_K1_GENERIC;
  return slab_file_ok(T, P, E, dtype, d_cube) ? CSG_K1_SLAB : CSG_K1_GENERIC;
}

int csg_slab_supported(int max_P, int max_E, int n_groups, int dtype) {
  return slab_config(max_P, max_E, n_groups, dtype).ok ? 1 : 0;
}

int32_t csg_collapse_blocks(int32_t T, int32_t P, int32_t E, int dtype, int layout, int kernel) {
  if (T <= 0 || E <= 0) return 0;
  if (layout == CSG_LAYOUT_TEP) {
    const int rows = tep_rows_per_block(P, dtype);
    return (int32_t)(((long long)T * E + rows - 1) / rows);
  }
  if (kernel == CSG_K1_SLAB) return (T + kTileRows - 1) / kTileRows;
  const int vec = dtype == CSG_F64 ? 2 : 4;
  const long long items = (long long)T * ((E + vec - 1) / vec);
  return (int32_t)((items + kBlock - 1) / kBlock);
}

int64_t csg_sums_elems(int32_t T, int32_t E, int n_groups) {
  if (T <= 0 || E <= 0) return 0;
  return (int64_t)(n_groups + 1) * E * ((T + 3) & ~3);
}

int csg_collapse(csg_ctx* ctx, const csg_file_desc* d_files, int n_files, int total_blocks,
                 const uint8_t* d_pa_bits, int n_groups, int max_P, int max_E, int dtype, int layout, int kernel,
                 void* d_sums, uint8_t* d_row_flags) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_groups < 0 || n_groups > CSG_MAX_GROUPS) return csg_fail(ctx, CSG_ERR_ARG, "n_groups %d out of range", n_groups);
  if (n_groups > 0 && !d_pa_bits) return csg_fail(ctx, CSG_ERR_ARG, "d_pa_bits is NULL with n_groups > 0");
  if (dtype != CSG_F32 && dtype != CSG_F64) return csg_fail(ctx, CSG_ERR_ARG, "bad dtype %d", dtype);
  if (layout != CSG_LAYOUT_TPE && layout != CSG_LAYOUT_TEP) return csg_fail(ctx, CSG_ERR_ARG, "bad layout %d", layout);
  if (n_files <= 0 || total_blocks <= 0) return CSG_OK;
  if (!d_files || !d_sums) return csg_fail(ctx, CSG_ERR_ARG, "NULL table or output");
  if (max_P <= 0 || max_P > 32768) return csg_fail(ctx, CSG_ERR_ARG, "max_P %d out of range (1..32768)", max_P);
  if (layout == CSG_LAYOUT_TPE) {
    if (kernel == CSG_K1_SLAB) {
      const SlabCfg c = slab_config(max_P, max_E, n_groups, dtype);
      if (!c.ok)
        return csg_fail(ctx, CSG_ERR_ARG, "slab kernel cannot stage (P=%d, E=%d): use CSG_K1_GENERIC for this table", max_P,
                        max_E);
      if (dtype == CSG_F32)
        return launch_slab<float>(ctx, c, d_files, n_files, total_blocks, d_pa_bits, n_groups, max_P, max_E, (float*)d_sums,
                                  d_row_flags);
      return launch_slab<double>(ctx, c, d_files, n_files, total_blocks, d_pa_bits, n_groups, max_P, max_E, (double*)d_sums,
                                 d_row_flags);
    }
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
  const int Tp = (T + 3) & ~3;
  const size_t sums_bytes = (size_t)csg_sums_elems(T, E, n_groups) * es;
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
  int kernel = csg_collapse_kernel(T, P, E, dtype, layout, d_cube);
  if (kernel == CSG_K1_SLAB && !csg_slab_supported(P, E, n_groups, dtype)) kernel = CSG_K1_GENERIC;
  const int blocks = csg_collapse_blocks(T, P, E, dtype, layout, kernel);
  CSG_TRY(csg_h2d(ctx, d_cube, h_cube, cube_bytes));
  CSG_TRY(csg_h2d(ctx, d_desc, &desc, sizeof(desc)));
  if (n_groups > 0) CSG_TRY(csg_h2d(ctx, d_bits, h_pa_bits, (size_t)P));
  CSG_TRY(csg_memset(ctx, d_flags, 0, flag_bytes));
  CSG_TRY(csg_collapse(ctx, (const csg_file_desc*)d_desc, 1, blocks, (const uint8_t*)d_bits, n_groups, P, E, dtype, layout,
                       kernel, d_sums, (uint8_t*)d_flags));
  // the device keeps sums energy-major [g][e][Tp]; the host entry hands back numpy's (T, E)
  std::vector<unsigned char> tmp(sums_bytes);
  CSG_TRY(csg_d2h(ctx, tmp.data(), d_sums, sums_bytes));
  if (h_row_flags) CSG_TRY(csg_d2h(ctx, h_row_flags, d_flags, (size_t)T));
  CSG_TRY(csg_sync(ctx));
#undef CSG_TRY
  for (int g = 0; g <= n_groups; ++g)
    for (int e = 0; e < E; ++e) {
      const unsigned char* src = tmp.data() + ((size_t)g * E + e) * Tp * es;
      unsigned char* dst = (unsigned char*)h_sums + ((size_t)g * T * E + e) * es;
      for (int t = 0; t < T; ++t) memcpy(dst + (size_t)t * E * es, src + (size_t)t * es, es);
    }
  cleanup();
  return CSG_OK;
}

}  // extern "C"
