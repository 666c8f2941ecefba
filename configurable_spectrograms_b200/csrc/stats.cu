// K2a -- per-region exact percentiles + min/max reductions.
//
// Replaces compute_percentile_bounds -> np.nanpercentile(matrix, p) (CS/percentile_utils.py:87-88;
// called at CS/fast/plotting.py:134,286 and CS/plotting.py:259) and the reductions
// safe_vmin = nanmin(matrix[isfinite & > 0]) (CS/plotting.py:261-262), nanmin/nanmax (:314-315).
//
// A region is a set of energy rows x a time range of one energy-major collapsed matrix, so its
// cells are contiguous runs.  One thread block per region, ONE pass over the cells:
//
//   sample   2048 evenly spread cells are sorted in shared memory; for each wanted percentile two
//            pivots bracket its quantile with a 3.5-sigma margin (pivots are sample VALUES, so
//            heavy ties -- spectrogram counts repeat massively -- collapse a bracket onto the
//            tied value instead of widening it)
//   pass     every cell is classified (NaN / inf / positive, min / max reductions) and compared
//            with the pivots: counts below / equal to each pivot; cells strictly inside a bracket
//            are appended to a small shared-memory candidate list (warp-aggregated append)
//   resolve  n is known now: numpy's rank arithmetic in D gives the two neighbour ranks of each
//            percentile; a rank is answered by a pivot (tie counts) or by the sorted candidates.
//
// No per-cell atomics, no histogram.  The rare region whose rank escapes its bracket (or whose
// candidates overflow) is flagged and redone by the multi-pass MSD radix select below, which is
// exact for any input.  Interpolation is numpy's lerp rounded after every operation
// (SURVEY.md Appendix B).
#include <stdlib.h>

#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kSample = 4096;   // storage shared by the sample and the candidate lists
constexpr int kDraw = 2048;     // sample size: two-sided brackets at the 1st / 99th percentile need m*(1-q) > margin
constexpr int kCand = 2048;     // candidate capacity per bracket (the lists reuse the sample's storage)
static_assert(2 * kCand == kSample, "candidate lists alias the sample buffer");
constexpr int kMaxCols = 1024;  // energy-row lists up to this length are staged in shared memory
constexpr int kDigitBits = 11;
constexpr int kBins = 1 << kDigitBits;
constexpr int kTargets = 4;  // (lo, hi) neighbours of two percentiles

// Bitonic sort of n_pow2 keys in shared memory (any power of two; used for short lists).
template <typename U>
__device__ void bitonic_sort_small(U* a, int n_pow2) {
  for (int k = 2; k <= n_pow2; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      __syncthreads();
      for (int i = threadIdx.x; i < n_pow2; i += kThreads) {
        const int l = i ^ j;
        if (l > i) {
          const U x = a[i], y = a[l];
          const bool up = (i & k) == 0;
          if ((x > y) == up) a[i] = y, a[l] = x;
        }
      }
    }
  __syncthreads();
}

// Bitonic sort of kThreads * C keys: every thread keeps C consecutive keys in registers.  Compare-
// exchange distances below C stay inside the thread, distances below 32 * C are warp shuffles, and
// only the few largest distances go through shared memory (two barriers each).
template <typename U, int C>
__device__ void bitonic_sort_regs(U* a) {
  constexpr int N = kThreads * C;
  const int t = threadIdx.x, base = t * C;
  U r[C];
  __syncthreads();  // the keys were just written by other threads
#pragma unroll
  for (int e = 0; e < C; ++e) r[e] = a[base + e];
  auto keep = [](U mine, U other, bool want_min) { return want_min == (mine < other) ? mine : other; };
#pragma unroll
  for (int k = 2; k <= N; k <<= 1) {
#pragma unroll
    for (int j = k >> 1; j > 0; j >>= 1) {
      if (j >= 32 * C) {  // partner in another warp: through shared memory
        __syncthreads();
#pragma unroll
        for (int e = 0; e < C; ++e) a[base + e] = r[e];
        __syncthreads();
        const bool up = (base & k) == 0, low = (base & j) == 0;
#pragma unroll
        for (int e = 0; e < C; ++e) r[e] = keep(r[e], a[(base + e) ^ j], low == up);
      } else if (j >= C) {  // partner in another lane, same element slot
        const bool up = (base & k) == 0, low = (base & j) == 0;
#pragma unroll
        for (int e = 0; e < C; ++e) {
          const U other = __shfl_xor_sync(0xffffffffu, r[e], j / C);
          r[e] = keep(r[e], other, low == up);
        }
      } else {  // both keys in this thread
#pragma unroll
        for (int e = 0; e < C; ++e) {
          const int l = e ^ j;
          if (l > e) {
            const bool up = ((base + e) & k) == 0;
            const U x = r[e], y = r[l];
            const bool swap = (x > y) == up;
            r[e] = swap ? y : x;
            r[l] = swap ? x : y;
          }
        }
      }
    }
  }
  __syncthreads();
#pragma unroll
  for (int e = 0; e < C; ++e) a[base + e] = r[e];
  __syncthreads();
}

template <typename U>
__device__ void block_sort(U* a, int n_pow2) {
  switch (n_pow2) {
    case kThreads * 2: bitonic_sort_regs<U, 2>(a); break;
    case kThreads * 4: bitonic_sort_regs<U, 4>(a); break;
    case kThreads * 8: bitonic_sort_regs<U, 8>(a); break;
    default: bitonic_sort_small(a, n_pow2);
  }
}

// Walk every cell of a region with all lanes busy: the (energy row, 16-byte vector) items of a
// contiguous-time region are flattened over the block (two vector loads in flight per thread);
// the few unaligned head / tail cells of every row and row-list regions take a scalar loop.
// fn(v, in) is called with block-uniform control flow (`in` = this lane holds a real cell), so
// it may use full-mask warp primitives.
template <bool B>
struct Flag {
  static constexpr bool value = B;
};

// c += (v OP r): one compare and one predicated add.  (Written as `c += cond ? 1 : 0` the compiler emits
// add-to-temporary / predicated move / move back -- three instructions per counter per cell.)
#define CSG_COUNT_IF(NAME, OP)                                                                              \
  __device__ __forceinline__ void NAME(unsigned& c, float v, float r) {                                     \
    asm("{ .reg .pred q; setp." OP ".f32 q, %1, %2; @q add.u32 %0, %0, 1; }" : "+r"(c) : "f"(v), "f"(r));    \
  }                                                                                                         \
  __device__ __forceinline__ void NAME(unsigned& c, double v, double r) {                                   \
    asm("{ .reg .pred q; setp." OP ".f64 q, %1, %2; @q add.u32 %0, %0, 1; }" : "+r"(c) : "d"(v), "d"(r));    \
  }
__device__ __forceinline__ float min_of(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ float max_of(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double min_of(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ double max_of(double a, double b) { return fmax(a, b); }
CSG_COUNT_IF(count_if_gt, "gt")
CSG_COUNT_IF(count_if_lt, "lt")
CSG_COUNT_IF(count_if_eq, "eq")
#undef CSG_COUNT_IF

template <typename T, typename Fn>
__device__ __forceinline__ void for_each_cell(const T* __restrict__ mats, const csg_region& rg,
                                              const int32_t* __restrict__ pool, const int* s_cols, bool cols_in_smem,
                                              Fn&& fn) {
  constexpr int V = 16 / sizeof(T);
  const int tid = threadIdx.x;
  const T* base = mats + rg.mat_off;
  const int ne = rg.ne, nt = rg.nt;
  auto col_of = [&](int j) { return cols_in_smem ? s_cols[j] : __ldg(pool + rg.cols_off + j); };
  const bool vec_ok = rg.rows_off < 0 && (rg.ld % V) == 0 && ((reinterpret_cast<uintptr_t>(mats) & 15) == 0);
  int head = 0, nvec = 0;
  if (vec_ok) {
    const int mis = (int)((rg.mat_off + rg.t0) % V);
    head = (V - mis) % V;
    if (head > nt) head = nt;
    nvec = (nt - head) / V;
  }
  if (nvec > 0) {
    const long long total = (long long)ne * nvec;
    // item -> (row, vector) advanced incrementally: no division in the loop
    const int stride = 2 * kThreads;
    const int dq = stride / nvec, dr = stride - dq * nvec;
    int ra = tid / nvec, va = tid - ra * nvec;
    int rb = (tid + kThreads) / nvec, vb = (tid + kThreads) - rb * nvec;
    for (long long it0 = 0; it0 < total; it0 += stride) {
      const bool ina = it0 + tid < total, inb = it0 + tid + kThreads < total;
      T a[V], b[V];
      if constexpr (sizeof(T) == 4) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 qa = ina ? __ldg(reinterpret_cast<const float4*>(base + (long long)col_of(ra) * rg.ld + rg.t0 + head) + va) : z;
        const float4 qb = inb ? __ldg(reinterpret_cast<const float4*>(base + (long long)col_of(rb) * rg.ld + rg.t0 + head) + vb) : z;
        a[0] = (T)qa.x, a[1] = (T)qa.y, a[V - 2] = (T)qa.z, a[V - 1] = (T)qa.w;
        b[0] = (T)qb.x, b[1] = (T)qb.y, b[V - 2] = (T)qb.z, b[V - 1] = (T)qb.w;
      } else {
        const double2 z = make_double2(0.0, 0.0);
        const double2 qa = ina ? __ldg(reinterpret_cast<const double2*>(base + (long long)col_of(ra) * rg.ld + rg.t0 + head) + va) : z;
        const double2 qb = inb ? __ldg(reinterpret_cast<const double2*>(base + (long long)col_of(rb) * rg.ld + rg.t0 + head) + vb) : z;
        a[0] = (T)qa.x, a[V - 1] = (T)qa.y;
        b[0] = (T)qb.x, b[V - 1] = (T)qb.y;
      }
#pragma unroll
      for (int v = 0; v < V; ++v) fn(a[v], ina);
      if (it0 + kThreads < total) {
#pragma unroll
        for (int v = 0; v < V; ++v) fn(b[v], inb);
      }
      ra += dq, va += dr;
      if (va >= nvec) va -= nvec, ++ra;
      rb += dq, vb += dr;
      if (vb >= nvec) vb -= nvec, ++rb;
    }
  }
  // scalar cells: the head and tail of every row (vector path), or every cell otherwise
  const int done = head + nvec * V;          // cells [head, done) of a row went through the vector loop
  const int per_row = nvec > 0 ? nt - (done - head) : nt;
  if (per_row > 0) {
    const long long total = (long long)ne * per_row;
    for (long long it0 = 0; it0 < total; it0 += kThreads) {
      const long long it = it0 + tid;
      const bool in = it < total;
      T v = T(0);
      if (in) {
        const int j = (int)(it / per_row);
        int k = (int)(it - (long long)j * per_row);
        if (nvec > 0 && k >= head) k += done - head;  // skip the vectorised middle
        const int t = rg.rows_off < 0 ? rg.t0 + k : __ldg(pool + rg.rows_off + k);
        v = __ldg(base + (long long)col_of(j) * rg.ld + t);
      }
      fn(v, in);
    }
  }
}

// The same walk for per-thread work: fn(v) is called for real cells only (no lane masks, no warp
// primitives inside fn); a thread takes whole 16-byte vectors, two in flight.
template <typename T, typename Fn>
__device__ __forceinline__ void for_each_real_cell(const T* __restrict__ mats, const csg_region& rg,
                                                   const int32_t* __restrict__ pool, const int* s_cols, bool cols_in_smem,
                                                   Fn&& fn) {
  constexpr int V = 16 / sizeof(T);
  const int tid = threadIdx.x;
  const T* base = mats + rg.mat_off;
  const int ne = rg.ne, nt = rg.nt;
  auto col_of = [&](int j) { return cols_in_smem ? s_cols[j] : __ldg(pool + rg.cols_off + j); };
  const bool vec_ok = rg.rows_off < 0 && (rg.ld % V) == 0 && ((reinterpret_cast<uintptr_t>(mats) & 15) == 0);
  int head = 0, nvec = 0;
  if (vec_ok) {
    const int mis = (int)((rg.mat_off + rg.t0) % V);
    head = (V - mis) % V;
    if (head > nt) head = nt;
    nvec = (nt - head) / V;
  }
  if (nvec > 0) {
    const int total = ne * nvec;  // vectors of the region (a region is far below 2^31 cells)
    const int stride = 2 * kThreads;
    const int dq = stride / nvec, dr = stride - dq * nvec;
    int ra = tid / nvec, va = tid - ra * nvec;
    int rb = (tid + kThreads) / nvec, vb = (tid + kThreads) - rb * nvec;
    const T* row0 = base + rg.t0 + head;
    for (int it = tid; it < total; it += stride) {
      const bool inb = it + kThreads < total;
      if constexpr (sizeof(T) == 4) {
        const float4 qa = __ldg(reinterpret_cast<const float4*>(row0 + (long long)col_of(ra) * rg.ld) + va);
        float4 qb = make_float4(0.f, 0.f, 0.f, 0.f);
        if (inb) qb = __ldg(reinterpret_cast<const float4*>(row0 + (long long)col_of(rb) * rg.ld) + vb);
        fn((T)qa.x), fn((T)qa.y), fn((T)qa.z), fn((T)qa.w);
        if (inb) fn((T)qb.x), fn((T)qb.y), fn((T)qb.z), fn((T)qb.w);
      } else {
        const double2 qa = __ldg(reinterpret_cast<const double2*>(row0 + (long long)col_of(ra) * rg.ld) + va);
        double2 qb = make_double2(0.0, 0.0);
        if (inb) qb = __ldg(reinterpret_cast<const double2*>(row0 + (long long)col_of(rb) * rg.ld) + vb);
        fn((T)qa.x), fn((T)qa.y);
        if (inb) fn((T)qb.x), fn((T)qb.y);
      }
      ra += dq, va += dr;
      if (va >= nvec) va -= nvec, ++ra;
      rb += dq, vb += dr;
      if (vb >= nvec) vb -= nvec, ++rb;
    }
  }
  // scalar cells: the head and tail of every row (vector path), or every cell otherwise
  const int done = head + nvec * V;  // cells [head, done) of a row went through the vector loop
  const int per_row = nvec > 0 ? nt - (done - head) : nt;
  if (per_row > 0) {
    const long long total = (long long)ne * per_row;
    for (long long it = tid; it < total; it += kThreads) {
      const int j = (int)(it / per_row);
      int k = (int)(it - (long long)j * per_row);
      if (nvec > 0 && k >= head) k += done - head;  // skip the vectorised middle
      const int t = rg.rows_off < 0 ? rg.t0 + k : __ldg(pool + rg.rows_off + k);
      fn(__ldg(base + (long long)col_of(j) * rg.ld + t));
    }
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads, 6)
    region_stats_kernel(const T* __restrict__ mats, const csg_region* __restrict__ regions,
                        const int32_t* __restrict__ pool, csg_region_stats* __restrict__ out,
                        uint8_t* __restrict__ todo, int force_exact) {
  typedef typename Key<T>::U U;
  constexpr U KEY_MIN = 0;       // below the key of every non-NaN value (-inf maps above 0)
  constexpr U KEY_MAX = ~(U)0;   // above the key of +inf
  __shared__ U s_buf[kSample];  // the sorted sample, then (pivots taken) the two candidate lists
  U* s_keys = s_buf;
  U(*s_cand)[kCand] = reinterpret_cast<U(*)[kCand]>(s_buf);
  __shared__ int s_ncand[2];
  __shared__ unsigned s_rare[3];  // NaN, +inf, -inf cells
  __shared__ int s_cols[kMaxCols];
  __shared__ unsigned s_u[32];
  __shared__ T s_t[32];
  __shared__ U s_pivot[2][2];
  __shared__ unsigned s_hist8[kThreads];  // 256 bins of the candidate-list radix select
  __shared__ unsigned long long s_sel_scratch[32];
  __shared__ int s_sel_digit;
  __shared__ long long s_sel_rank;
  __shared__ unsigned s_sel_eq;

  const csg_region rg = regions[blockIdx.x];
  if (threadIdx.x == 0) todo[blockIdx.x] = 0;
  if (rg.want_pct == 2) return;  // geometry only
  const bool pct = rg.want_pct == 1;
  const int tid = threadIdx.x;
  const bool cols_in_smem = rg.ne <= kMaxCols;
  const long long n_cells = (long long)rg.ne * rg.nt;

  if (cols_in_smem)
    for (int i = tid; i < rg.ne; i += kThreads) s_cols[i] = __ldg(pool + rg.cols_off + i);
  if (tid < 2) s_ncand[tid] = 0;
  if (tid < 3) s_rare[tid] = 0;
  __syncthreads();

  // ---- brackets
  const bool small = n_cells <= kCand;  // everything fits the candidate list: no sampling
  if (pct && !small) {
    // evenly spread sample positions (n_cells > kSample / 2, so neighbours may repeat a cell for
    // regions below kSample cells: harmless, the sample only places the pivots)
    for (int s = tid; s < kDraw; s += kThreads) {
      const long long cell = ((2 * (long long)s + 1) * n_cells) / (2 * kDraw);
      const int j = (int)(cell / rg.nt), k = (int)(cell - (long long)j * rg.nt);
      const int col = cols_in_smem ? s_cols[j] : __ldg(pool + rg.cols_off + j);
      const int t = rg.rows_off < 0 ? rg.t0 + k : __ldg(pool + rg.rows_off + k);
      const T v = __ldg(mats + rg.mat_off + (long long)col * rg.ld + t);
      s_keys[s] = is_nan(v) ? KEY_MAX : Key<T>::key(v);  // NaN sorts to the end
    }
    block_sort(s_keys, kDraw);
    if (tid < 2) {
      int lo = 0, hi = kDraw;  // valid sample size m: the NaN sentinels sit at the end
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (s_keys[mid] == KEY_MAX)
          hi = mid;
        else
          lo = mid + 1;
      }
      const int m = lo;
      const double q = (tid == 0 ? rg.p_lo : rg.p_hi) / 100.0;
      U plo = KEY_MIN, phi = KEY_MAX;
      if (m >= 16 && q >= 0.0 && q <= 1.0) {
        const double centre = q * (m - 1);
        // 3.5 sigma + 2: a rank escapes its bracket about once in a thousand regions (-> exact fallback)
        const double delta = 3.5 * sqrt(m * q * (1.0 - q)) + 2.0;
        const int a = (int)floor(centre - delta), b = (int)ceil(centre + delta);
        if (a >= 0) plo = s_keys[a];
        if (b < m) phi = s_keys[b];
      }
      s_pivot[tid][0] = plo, s_pivot[tid][1] = phi;
    }
    __syncthreads();  // pivots are out: the buffer becomes the candidate lists
  } else if (tid < 2) {
    s_pivot[tid][0] = KEY_MIN, s_pivot[tid][1] = KEY_MAX;
  }
  __syncthreads();
  const U plo0 = s_pivot[0][0], phi0 = s_pivot[0][1], plo1 = s_pivot[1][0], phi1 = s_pivot[1][1];
  const bool share = small;  // both percentiles read list 0

  // ---- the pass: everything a cell needs is per-thread work -- no votes, no shuffles.  A finite cell
  // (nansum leaves nothing else but the rare inf / inf-inf) costs two min/max, the positive count and
  // minimum, and two compares that tell whether it reaches into a bracket at all; only those cells touch
  // the bracket counters (float compares against the pivot VALUES: for finite cells they order exactly
  // like the keys, with -0.0 == +0.0) and the few strictly inside a bracket are appended to its list with
  // one shared-memory atomic each.  Non-finite cells take the key-based branch.
  unsigned n_pos = 0;
  unsigned lt0 = 0, eqlo0 = 0, eqhi0 = 0, eqlo1 = 0, eqhi1 = 0, ge1 = 0;
  const T kInf = (T)CUDART_INF;
  T min_pos = kInf, fin_min = kInf, fin_max = -kInf;
  auto pivot_value = [&](U key, T sentinel) { return (key == KEY_MIN || key == KEY_MAX) ? sentinel : Key<T>::val(key); };
  const T lo0 = pivot_value(plo0, -kInf), hi0 = pivot_value(phi0, kInf);
  const T lo1 = pivot_value(plo1, -kInf), hi1 = pivot_value(phi1, kInf);
  auto append = [&](int b, U key) {
    const int pos = atomicAdd(&s_ncand[b], 1);
    if (pos < kCand) s_cand[b][pos] = key;
  };
  auto by_key = [&](T v) {  // NaN / +-inf: rare, so its counters live in shared memory and the key pivots are re-read
    if (is_nan(v)) {
      atomicAdd(&s_rare[0], 1u);
      return;
    }
    atomicAdd(&s_rare[v > T(0) ? 1 : 2], 1u);
    if (!pct) return;
    const U k = Key<T>::key(v);
    const U a0 = s_pivot[0][0], b0 = s_pivot[0][1], a1 = s_pivot[1][0], b1 = s_pivot[1][1];
    lt0 += k < a0 ? 1u : 0u;
    eqlo0 += k == a0 ? 1u : 0u;
    eqhi0 += k == b0 ? 1u : 0u;
    if (k > a0 && k < b0) append(0, k);
    if (!share) {
      ge1 += k >= a1 ? 1u : 0u;
      eqlo1 += k == a1 ? 1u : 0u;
      eqhi1 += k == b1 ? 1u : 0u;
      if (k > a1 && k < b1) append(1, k);
    }
  };
  auto cell = [&](T v, auto pct_c, auto share_c) {
    if (!is_finite(v)) {
      by_key(v);
      return;
    }
    fin_min = min_of(fin_min, v);  // v is finite here: the hardware min / max (NaN-agnostic) are exact
    fin_max = max_of(fin_max, v);
    count_if_gt(n_pos, v, T(0));
    min_pos = min_of(min_pos, v > T(0) ? v : kInf);
    if constexpr (decltype(pct_c)::value) {
      constexpr bool SHARE = decltype(share_c)::value;
      if (v <= hi0) {
        count_if_lt(lt0, v, lo0);
        count_if_eq(eqlo0, v, lo0);
        count_if_eq(eqhi0, v, hi0);
        if (v > lo0 && v < hi0) append(0, Key<T>::key(v));
      }
      if constexpr (!SHARE) {
        if (v >= lo1) {
          ++ge1;
          count_if_eq(eqlo1, v, lo1);
          count_if_eq(eqhi1, v, hi1);
          if (v > lo1 && v < hi1) append(1, Key<T>::key(v));
        }
      }
    }
  };
  if (!pct)
    for_each_real_cell<T>(mats, rg, pool, s_cols, cols_in_smem, [&](T v) { cell(v, Flag<false>{}, Flag<true>{}); });
  else if (share)
    for_each_real_cell<T>(mats, rg, pool, s_cols, cols_in_smem, [&](T v) { cell(v, Flag<true>{}, Flag<true>{}); });
  else
    for_each_real_cell<T>(mats, rg, pool, s_cols, cols_in_smem, [&](T v) { cell(v, Flag<true>{}, Flag<false>{}); });

  auto addu = [](unsigned a, unsigned b) { return a + b; };
  auto mint = [](T a, T b) { return a < b ? a : b; };
  auto maxt = [](T a, T b) { return a > b ? a : b; };
  n_pos = block_reduce(n_pos, addu, 0u, s_u);  // (its barriers also publish s_rare)
  const unsigned n_nan = s_rare[0], n_pinf = s_rare[1], n_ninf = s_rare[2];
  const unsigned n_valid = (unsigned)n_cells - n_nan;
  min_pos = block_reduce(min_pos, mint, kInf, s_t);
  fin_min = block_reduce(fin_min, mint, kInf, s_t);
  fin_max = block_reduce(fin_max, maxt, (T)(-kInf), s_t);

  csg_region_stats st;
  st.p_lo = st.p_hi = CUDART_NAN;
  st.min_pos = (double)min_pos;
  st.fin_min = (double)fin_min;
  st.fin_max = (double)fin_max;
  st.n_valid = (long long)n_valid;
  st.n_nan = (int)n_nan, st.n_neginf = (int)n_ninf, st.n_posinf = (int)n_pinf, st.n_pos = (int)n_pos;
  if (!pct || n_valid == 0) {
    if (tid == 0) out[blockIdx.x] = st;
    return;
  }
  lt0 = block_reduce(lt0, addu, 0u, s_u);
  eqlo0 = block_reduce(eqlo0, addu, 0u, s_u);
  eqhi0 = block_reduce(eqhi0, addu, 0u, s_u);
  // a bracket collapsed onto one (tied) value -- equal keys, or -0.0 / +0.0 -- counted it under both names
  if (plo0 == phi0 || lo0 == hi0) eqhi0 = 0;
  unsigned lt1 = 0;
  if (!share) {
    ge1 = block_reduce(ge1, addu, 0u, s_u);
    lt1 = n_valid - ge1;
    eqlo1 = block_reduce(eqlo1, addu, 0u, s_u);
    eqhi1 = block_reduce(eqhi1, addu, 0u, s_u);
    if (plo1 == phi1 || lo1 == hi1) eqhi1 = 0;
  }
  __syncthreads();
  const int n0 = s_ncand[0], n1 = share ? n0 : s_ncand[1];
  // force_exact (csg_region_stats_force_exact): every percentile region is handed to the exact radix
  // select below -- the path a rank that escapes its bracket takes about once in a thousand regions
  bool fail = force_exact != 0 || n0 > kCand || n1 > kCand;
  // every thread holds the same reduced counters: the rank arithmetic below is block-uniform
  long long rank[kTargets];
  T gamma[2];
  percentile_ranks<T>((long long)n_valid, rg.p_lo, rank[0], rank[1], gamma[0]);
  percentile_ranks<T>((long long)n_valid, rg.p_hi, rank[2], rank[3], gamma[1]);
  T val[kTargets];
  // the candidate lists are NOT sorted: the (at most four, pairwise adjacent) ranks that fall inside
  // a list are answered by an 8-bit MSD radix select over the list in shared memory; a rank inside
  // the run of equal keys found last re-uses it, the rank just after it is that key's successor
  int prev_b = -1;
  long long prev_first = 0, prev_last = -1;  // in-list rank range of prev_key
  U prev_key = 0;
  for (int j = 0; j < kTargets && !fail; ++j) {
    const int b = share ? 0 : (j >> 1);
    const U plo = b == 0 ? plo0 : plo1, phi = b == 0 ? phi0 : phi1;
    const long long below = b == 0 ? lt0 : lt1, at_lo = b == 0 ? eqlo0 : eqlo1, at_hi = b == 0 ? eqhi0 : eqhi1;
    const long long inside = b == 0 ? n0 : n1;
    long long r = rank[j];
    U key = 0;
    if (r < below) {
      fail = true;
    } else if ((r -= below) < at_lo) {
      key = plo;
    } else if ((r -= at_lo) < inside) {
      const U* list = s_cand[b];
      const int n = (int)inside;
      if (b == prev_b && r >= prev_first && r <= prev_last) {
        key = prev_key;
      } else if (b == prev_b && r == prev_last + 1) {
        // successor: the smallest key above prev_key (exists: rank r is inside the list)
        U m = KEY_MAX;
        for (int i = tid; i < n; i += kThreads) {
          const U k = list[i];
          if (k > prev_key && k < m) m = k;
        }
        auto minu = [](U a, U c) { return a < c ? a : c; };
        __syncthreads();
        m = block_reduce(m, minu, KEY_MAX, reinterpret_cast<U*>(s_sel_scratch));
        key = m;
        prev_key = m, prev_first = r, prev_last = r;  // its run length is unknown: claim one rank only
      } else {
        U prefix = 0, mask = 0;
        long long want = r;
        unsigned eq = 0;
        for (int shift = Key<T>::BITS - 8; shift >= 0; shift -= 8) {
          __syncthreads();
          s_hist8[tid] = 0;  // kThreads == 256 bins
          __syncthreads();
          for (int i = tid; i < n; i += kThreads) {
            const U k = list[i];
            if ((k & mask) == prefix) atomicAdd(&s_hist8[(unsigned)(k >> shift) & 255u], 1u);
          }
          __syncthreads();
          if (tid < 32) {  // warp 0: lane l owns bins 8l..8l+7
            unsigned c[8], mine = 0;
#pragma unroll
            for (int q = 0; q < 8; ++q) c[q] = s_hist8[8 * tid + q], mine += c[q];
            unsigned inc = mine;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
              const unsigned t = __shfl_up_sync(0xffffffffu, inc, o);
              if (tid >= o) inc += t;
            }
            long long run = (long long)(inc - mine);
            if (want >= run && want < run + (long long)mine) {
#pragma unroll
              for (int q = 0; q < 8; ++q) {
                if (want >= run && want < run + (long long)c[q]) {
                  s_sel_digit = 8 * tid + q;
                  s_sel_rank = want - run;
                  s_sel_eq = c[q];
                }
                run += c[q];
              }
            }
          }
          __syncthreads();
          prefix |= (U)s_sel_digit << shift;
          mask |= (U)255 << shift;
          want = s_sel_rank;
          eq = s_sel_eq;
        }
        key = prefix;
        prev_key = key, prev_first = r - want, prev_last = r - want + (long long)eq - 1;
      }
      prev_b = b;
    } else if ((r -= inside) < at_hi) {
      key = phi;
    } else {
      fail = true;
    }
    val[j] = Key<T>::val(key);
  }
  if (tid == 0) {
    if (!fail) {
      st.p_lo = (double)numpy_lerp<T>(val[0], val[1], gamma[0]);
      st.p_hi = (double)numpy_lerp<T>(val[2], val[3], gamma[1]);
    } else {
      todo[blockIdx.x] = 1;  // redone exactly by region_select_kernel
    }
    out[blockIdx.x] = st;
  }
}

// ---------------------------------------------------------------------------------------------
// Exact fallback: MSD radix select on the order-preserving key (a histogram pass per 11-bit digit,
// the bucket holding each wanted rank is followed into the next digit).  Runs only for regions
// flagged by region_stats_kernel (n_valid is already in `out`).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kThreads)
    region_select_kernel(const T* __restrict__ mats, const csg_region* __restrict__ regions,
                         const int32_t* __restrict__ pool, csg_region_stats* __restrict__ out,
                         const uint8_t* __restrict__ todo) {
  typedef typename Key<T>::U U;
  if (!todo[blockIdx.x]) return;
  __shared__ unsigned s_hist[kTargets][kBins];
  __shared__ int s_cols[kMaxCols];
  __shared__ long long s_ll[32];
  __shared__ U s_prefix[kTargets];
  __shared__ long long s_rank[kTargets];
  __shared__ int s_hidx[kTargets];

  const csg_region rg = regions[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool cols_in_smem = rg.ne <= kMaxCols;
  if (cols_in_smem)
    for (int i = tid; i < rg.ne; i += kThreads) s_cols[i] = __ldg(pool + rg.cols_off + i);
  const long long n_valid = out[blockIdx.x].n_valid;
  long long rank[kTargets];
  T gamma[2];
  percentile_ranks<T>(n_valid, rg.p_lo, rank[0], rank[1], gamma[0]);
  percentile_ranks<T>(n_valid, rg.p_hi, rank[2], rank[3], gamma[1]);
  if (tid < kTargets) {
    s_prefix[tid] = 0;
    s_rank[tid] = rank[tid];
    s_hidx[tid] = 0;
  }
  int shift = Key<T>::BITS;
  while (shift > 0) {
    const int bits = shift < kDigitBits ? shift : kDigitBits;
    const int prev_shift = shift;
    shift -= bits;
    const unsigned mask = (1u << bits) - 1u;
    const int nb = 1 << bits;
    __syncthreads();
    if (tid == 0) {  // one histogram per DISTINCT prefix (lo/hi neighbours usually share theirs)
      for (int j = 0; j < kTargets; ++j) {
        int h = j;
        for (int i = 0; i < j; ++i)
          if (s_prefix[i] == s_prefix[j]) {
            h = s_hidx[i];
            break;
          }
        s_hidx[j] = h;
      }
    }
    for (int i = tid; i < kTargets * kBins; i += kThreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const U p0 = s_prefix[0], p1 = s_prefix[1], p2 = s_prefix[2], p3 = s_prefix[3];
    const bool u1 = s_hidx[1] == 1, u2 = s_hidx[2] == 2, u3 = s_hidx[3] == 3;
    const bool first = prev_shift == Key<T>::BITS;
    for_each_cell<T>(mats, rg, pool, s_cols, cols_in_smem, [&](T v, bool in) {
      if (!in || is_nan(v)) return;
      const U k = Key<T>::key(v);
      const U hi = first ? (U)0 : (U)(k >> (prev_shift & (Key<T>::BITS - 1)));
      const unsigned b = (unsigned)(k >> shift) & mask;
      if (hi == p0) atomicAdd(&s_hist[0][b], 1u);
      if (u1 && hi == p1) atomicAdd(&s_hist[1][b], 1u);
      if (u2 && hi == p2) atomicAdd(&s_hist[2][b], 1u);
      if (u3 && hi == p3) atomicAdd(&s_hist[3][b], 1u);
    });
    __syncthreads();
    // locate each target's bucket in the histogram of its prefix
    for (int j = 0; j < kTargets; ++j) {
      const unsigned* h = s_hist[s_hidx[j]];
      const long long want_rank = s_rank[j];
      const int per = (nb + kThreads - 1) / kThreads;
      const int b0 = tid * per;
      long long mine = 0;
      for (int b = b0; b < b0 + per && b < nb; ++b) mine += h[b];
      long long inc = mine;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        long long n = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += n;
      }
      __syncthreads();
      if (lane == 31) s_ll[warp] = inc;
      __syncthreads();
      long long warp_off = 0;
      for (int w = 0; w < warp; ++w) warp_off += s_ll[w];
      const long long excl = warp_off + inc - mine;
      if (want_rank >= excl && want_rank < excl + mine) {
        long long run = excl;
        for (int b = b0; b < b0 + per && b < nb; ++b) {
          const long long cnt = h[b];
          if (want_rank < run + cnt) {
            s_prefix[j] = (s_prefix[j] << bits) | (U)b;
            s_rank[j] = want_rank - run;
            break;
          }
          run += cnt;
        }
      }
      __syncthreads();
    }
  }
  if (tid == 0) {
    const T a0 = Key<T>::val(s_prefix[0]), b0 = Key<T>::val(s_prefix[1]);
    const T a1 = Key<T>::val(s_prefix[2]), b1 = Key<T>::val(s_prefix[3]);
    out[blockIdx.x].p_lo = (double)numpy_lerp<T>(a0, b0, gamma[0]);
    out[blockIdx.x].p_hi = (double)numpy_lerp<T>(a1, b1, gamma[1]);
  }
}

}  // namespace

int csg_scratch(csg_ctx* ctx, size_t bytes, void** out) {
  if (ctx->scratch_bytes < bytes) {
    if (ctx->scratch) {
      CSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
      CSG_CUDA(ctx, cudaFree(ctx->scratch));
      ctx->scratch = nullptr, ctx->scratch_bytes = 0;
    }
    const size_t want = bytes + bytes / 2 + 4096;
    cudaError_t e = cudaMalloc(&ctx->scratch, want);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return csg_fail(ctx, CSG_ERR_NOMEM, "cudaMalloc(%zu) for scratch failed: %s", want, cudaGetErrorString(e));
    }
    ctx->scratch_bytes = want;
  }
  *out = ctx->scratch;
  return CSG_OK;
}

extern "C" int csg_region_stats_run(csg_ctx* ctx, const void* d_mats, int dtype, const csg_region* d_regions,
                                    int n_regions, const int32_t* d_index_pool, csg_region_stats* d_out) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_regions <= 0) return CSG_OK;
  if (!d_mats || !d_regions || !d_index_pool || !d_out) return csg_fail(ctx, CSG_ERR_ARG, "NULL argument");
  if (dtype != CSG_F32 && dtype != CSG_F64) return csg_fail(ctx, CSG_ERR_ARG, "bad dtype %d", dtype);
  void* todo = nullptr;
  int st = csg_scratch(ctx, (size_t)n_regions, &todo);
  if (st != CSG_OK) return st;
  if (dtype == CSG_F32) {
    region_stats_kernel<float><<<n_regions, kThreads, 0, ctx->stream>>>((const float*)d_mats, d_regions, d_index_pool, d_out,
                                                                        (uint8_t*)todo, ctx->stats_force_exact);
    CSG_LAUNCH_CHECK(ctx, "region_stats_kernel");
    region_select_kernel<float><<<n_regions, kThreads, 0, ctx->stream>>>((const float*)d_mats, d_regions, d_index_pool, d_out,
                                                                         (const uint8_t*)todo);
  } else {
    region_stats_kernel<double><<<n_regions, kThreads, 0, ctx->stream>>>((const double*)d_mats, d_regions, d_index_pool,
                                                                         d_out, (uint8_t*)todo, ctx->stats_force_exact);
    CSG_LAUNCH_CHECK(ctx, "region_stats_kernel");
    region_select_kernel<double><<<n_regions, kThreads, 0, ctx->stream>>>((const double*)d_mats, d_regions, d_index_pool,
                                                                          d_out, (const uint8_t*)todo);
  }
  CSG_LAUNCH_CHECK(ctx, "region_select_kernel");
  return CSG_OK;
}

extern "C" int csg_region_stats_force_exact(csg_ctx* ctx, int on) {
  if (!ctx) return CSG_ERR_ARG;
  ctx->stats_force_exact = on ? 1 : 0;
  return CSG_OK;
}

// how many regions of the last csg_region_stats_run() needed the radix-select fallback (synchronises)
extern "C" int csg_region_stats_fallbacks(csg_ctx* ctx, int n_regions, int* count) {
  if (!ctx || !count) return CSG_ERR_ARG;
  *count = 0;
  if (n_regions <= 0 || !ctx->scratch || ctx->scratch_bytes < (size_t)n_regions) return CSG_OK;
  unsigned char* h = (unsigned char*)malloc((size_t)n_regions);
  if (!h) return csg_fail(ctx, CSG_ERR_NOMEM, "malloc(%d) failed", n_regions);
  cudaError_t e = cudaMemcpyAsync(h, ctx->scratch, (size_t)n_regions, cudaMemcpyDeviceToHost, ctx->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(ctx->stream);
  if (e != cudaSuccess) {
    free(h);
    return csg_fail(ctx, CSG_ERR_CUDA, "reading fallback flags failed: %s", cudaGetErrorString(e));
  }
  int n = 0;
  for (int i = 0; i < n_regions; ++i) n += h[i] ? 1 : 0;
  free(h);
  *count = n;
  return CSG_OK;
}
