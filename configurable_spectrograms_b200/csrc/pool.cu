// K2b -- global extrema over the pooled finite-positive samples of every orbit.
//
// Replaces the numeric core of compute_global_extrema (CS/fast/extrema.py:259-285): the
// finite-positive mask (:260), per-energy positive counts (:261-264) and
// nanpercentile(concatenate(blocks so far), max_percentile) (:280-285), which the reference
// recomputes for EVERY prefix of the ascending-orbit sequence (quadratic on the CPU).
//
// The pool is never materialised.  Per file the collapsed total matrix is histogrammed by
// radix digit of the positive-float key; histograms are prefix-scanned along each
// instrument's file sequence (csg_pool_scan, optionally on top of lower ranks' totals), so
// row k holds the digit histogram of the pool after k+1 files; csg_pool_locate walks one
// rank per query into its bucket.  Repeating this digit by digit (host-driven, with pruning
// of prefixes that cannot hold the running maximum) yields exact order statistics of every
// prefix pool.  Finite positive floats order like their raw bit patterns (31 / 63 key bits).
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

template <typename T>
__global__ void __launch_bounds__(kThreads)
    pool_hist_first_kernel(const T* __restrict__ mats, const csg_pool_item* __restrict__ items, int max_pos,
                           int bits, int max_E, uint32_t* __restrict__ hist, int32_t* __restrict__ counts,
                           int32_t* __restrict__ npos) {
  extern __shared__ unsigned s_mem[];
  const int nb = 1 << bits;
  unsigned* s_hist = s_mem;        // [nb]
  unsigned* s_cnt = s_mem + nb;    // [E]
  __shared__ unsigned s_red[32];
  const csg_pool_item it = items[blockIdx.x];
  for (int i = threadIdx.x; i < nb + it.E; i += kThreads) s_mem[i] = 0;
  __syncthreads();
  const T* m = mats + it.mat_off;
  constexpr int kShiftBase = Key<T>::POS_BITS;
  const int shift = kShiftBase - bits;
  unsigned mine = 0;
  for (int i = threadIdx.x; i < it.n_cells; i += kThreads) {
    const T v = __ldg(m + i);
    if (is_finite(v) && v > T(0)) {
      ++mine;
      atomicAdd(&s_hist[(unsigned)(Key<T>::bits(v) >> shift)], 1u);
      atomicAdd(&s_cnt[i % it.E], 1u);
    }
  }
  auto addu = [](unsigned a, unsigned b) { return a + b; };
  mine = block_reduce(mine, addu, 0u, s_red);
  __syncthreads();
  uint32_t* h = hist + ((size_t)it.inst * max_pos + it.pos) * nb;  // slot stride = 1 slot at level 0
  for (int i = threadIdx.x; i < nb; i += kThreads) h[i] = s_hist[i];
  int32_t* c = counts + (size_t)blockIdx.x * max_E;
  for (int i = threadIdx.x; i < it.E; i += kThreads) c[i] = (int32_t)s_cnt[i];
  if (threadIdx.x == 0) npos[blockIdx.x] = (int32_t)mine;
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    pool_hist_refine_kernel(const T* __restrict__ mats, const csg_pool_item* __restrict__ items, int max_pos,
                            int n_slots, const uint64_t* __restrict__ slot_prefix, int prefix_shift, int shift,
                            int bits, uint32_t* __restrict__ hist) {
  typedef typename Key<T>::U U;
  __shared__ uint64_t s_pref[64];
  const csg_pool_item it = items[blockIdx.x];
  for (int i = threadIdx.x; i < n_slots; i += kThreads) s_pref[i] = slot_prefix[(size_t)it.inst * n_slots + i];
  __syncthreads();
  const T* m = mats + it.mat_off;
  const int nb = 1 << bits;
  uint32_t* h = hist + ((size_t)it.inst * max_pos + it.pos) * (size_t)n_slots * nb;
  for (int i = threadIdx.x; i < it.n_cells; i += kThreads) {
    const T v = __ldg(m + i);
    if (is_finite(v) && v > T(0)) {
      const U k = Key<T>::bits(v);
      const uint64_t hi = (uint64_t)(k >> prefix_shift);
      // sorted ascending, padded with UINT64_MAX: binary search
      int lo = 0, up = n_slots - 1;
      while (lo < up) {
        const int mid = (lo + up) >> 1;
        if (s_pref[mid] < hi)
          lo = mid + 1;
        else
          up = mid;
      }
      if (s_pref[lo] == hi) atomicAdd(&h[(size_t)lo * nb + ((unsigned)(k >> shift) & (unsigned)(nb - 1))], 1u);
    }
  }
}

// thread = one (inst, slot, bin) column; walk the instrument's files in orbit order
__global__ void pool_scan_kernel(uint32_t* __restrict__ hist, int n_inst, int max_pos,
                                 const int32_t* __restrict__ inst_len, int n_slots, int nb,
                                 uint32_t* __restrict__ totals) {
  const size_t cols = (size_t)n_slots * nb;
  const size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (gid >= (size_t)n_inst * cols) return;
  const int inst = (int)(gid / cols);
  const size_t col = gid - (size_t)inst * cols;
  const int len = inst_len[inst];
  uint32_t* p = hist + (size_t)inst * max_pos * cols + col;
  uint32_t run = 0;
  for (int k = 0; k < len; ++k) {
    run += p[(size_t)k * cols];
    p[(size_t)k * cols] = run;
  }
  if (totals) totals[gid] = run;
}

// block = one query: find the bin where the cumulative count of row (inst,pos,slot) crosses rank
__global__ void __launch_bounds__(kThreads)
    pool_locate_kernel(const uint32_t* __restrict__ hist, int max_pos, int n_slots, int nb,
                       const uint32_t* __restrict__ base, csg_pool_query* __restrict__ queries) {
  __shared__ long long s_warp[32];
  csg_pool_query q = queries[blockIdx.x];
  const uint32_t* row = hist + (((size_t)q.inst * max_pos + q.pos) * n_slots + q.slot) * (size_t)nb;
  const uint32_t* brow = base ? base + ((size_t)q.inst * n_slots + q.slot) * (size_t)nb : nullptr;
  auto cell = [&](int b) -> long long { return (long long)row[b] + (brow ? (long long)brow[b] : 0ll); };
  const int per = (nb + kThreads - 1) / kThreads;
  const int b0 = threadIdx.x * per;
  long long mine = 0;
  for (int b = b0; b < b0 + per && b < nb; ++b) mine += cell(b);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  long long inc = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const long long n = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += n;
  }
  if (lane == 31) s_warp[warp] = inc;
  __syncthreads();
  long long off = 0, total = 0;
  for (int w = 0; w < kThreads / 32; ++w) {
    if (w < warp) off += s_warp[w];
    total += s_warp[w];
  }
  const long long excl = off + inc - mine;
  if (threadIdx.x == 0) {
    queries[blockIdx.x].row_total = total;
    if (q.rank < 0 || q.rank >= total) queries[blockIdx.x].bin = -1;  // empty pool / out of range
  }
  if (q.rank >= excl && q.rank < excl + mine) {
    long long run = excl;
    for (int b = b0; b < b0 + per && b < nb; ++b) {
      const long long c = cell(b);
      if (q.rank < run + c) {
        queries[blockIdx.x].bin = b;
        queries[blockIdx.x].rank = q.rank - run;
        break;
      }
      run += c;
    }
  }
}

}  // namespace

extern "C" {

int csg_pool_hist_first(csg_ctx* ctx, const void* d_mats, int dtype, const csg_pool_item* d_items, int n_items,
                        int max_pos, int bits, int max_E, uint32_t* d_hist, int32_t* d_counts, int32_t* d_npos) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_items <= 0) return CSG_OK;
  if (bits < 1 || bits > 12) return csg_fail(ctx, CSG_ERR_ARG, "bits %d out of range (1..12)", bits);
  if (max_E <= 0 || max_E > 8192) return csg_fail(ctx, CSG_ERR_ARG, "max_E %d out of range", max_E);
  const size_t smem = ((size_t)(1 << bits) + max_E) * sizeof(unsigned);
  if (dtype == CSG_F32)
    pool_hist_first_kernel<float><<<n_items, kThreads, smem, ctx->stream>>>((const float*)d_mats, d_items, max_pos, bits,
                                                                            max_E, d_hist, d_counts, d_npos);
  else if (dtype == CSG_F64)
    pool_hist_first_kernel<double><<<n_items, kThreads, smem, ctx->stream>>>((const double*)d_mats, d_items, max_pos,
                                                                             bits, max_E, d_hist, d_counts, d_npos);
  else
    return csg_fail(ctx, CSG_ERR_ARG, "bad dtype %d", dtype);
  CSG_LAUNCH_CHECK(ctx, "pool_hist_first_kernel");
  return CSG_OK;
}

int csg_pool_hist_refine(csg_ctx* ctx, const void* d_mats, int dtype, const csg_pool_item* d_items, int n_items,
                         int max_pos, int n_slots, const uint64_t* d_slot_prefix, int prefix_shift, int shift, int bits,
                         uint32_t* d_hist) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_items <= 0) return CSG_OK;
  if (n_slots < 1 || n_slots > 64) return csg_fail(ctx, CSG_ERR_ARG, "n_slots %d out of range (1..64)", n_slots);
  if (bits < 1 || bits > 12) return csg_fail(ctx, CSG_ERR_ARG, "bits %d out of range (1..12)", bits);
  if (dtype == CSG_F32)
    pool_hist_refine_kernel<float><<<n_items, kThreads, 0, ctx->stream>>>((const float*)d_mats, d_items, max_pos, n_slots,
                                                                          d_slot_prefix, prefix_shift, shift, bits, d_hist);
  else if (dtype == CSG_F64)
    pool_hist_refine_kernel<double><<<n_items, kThreads, 0, ctx->stream>>>((const double*)d_mats, d_items, max_pos,
                                                                           n_slots, d_slot_prefix, prefix_shift, shift,
                                                                           bits, d_hist);
  else
    return csg_fail(ctx, CSG_ERR_ARG, "bad dtype %d", dtype);
  CSG_LAUNCH_CHECK(ctx, "pool_hist_refine_kernel");
  return CSG_OK;
}

int csg_pool_scan(csg_ctx* ctx, uint32_t* d_hist, int n_inst, int max_pos, const int32_t* d_inst_len, int n_slots,
                  int bits, uint32_t* d_totals) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_inst <= 0 || max_pos <= 0) return CSG_OK;
  const size_t n = (size_t)n_inst * n_slots * ((size_t)1 << bits);
  const int blocks = (int)((n + 255) / 256);
  pool_scan_kernel<<<blocks, 256, 0, ctx->stream>>>(d_hist, n_inst, max_pos, d_inst_len, n_slots, 1 << bits, d_totals);
  CSG_LAUNCH_CHECK(ctx, "pool_scan_kernel");
  return CSG_OK;
}

int csg_pool_locate(csg_ctx* ctx, const uint32_t* d_hist, int max_pos, int n_slots, int bits, const uint32_t* d_base,
                    csg_pool_query* d_queries, int n_queries) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_queries <= 0) return CSG_OK;
  pool_locate_kernel<<<n_queries, kThreads, 0, ctx->stream>>>(d_hist, max_pos, n_slots, 1 << bits, d_base, d_queries);
  CSG_LAUNCH_CHECK(ctx, "pool_locate_kernel");
  return CSG_OK;
}

}  // extern "C"
