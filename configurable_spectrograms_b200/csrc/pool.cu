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

constexpr int kSplit = 4;  // blocks per file: each takes a contiguous band of energy rows

// Walk rows [e0, e1) of an energy-major [E][ld] matrix, warp per row, 128-bit loads (two in
// flight); fn(v, in, e) is called with warp-uniform control flow, row_done(e) once per row.
template <typename T, typename Fn, typename RowDone>
__device__ __forceinline__ void walk_rows(const T* __restrict__ m, int ld, int nt, int e0, int e1, Fn&& fn,
                                          RowDone&& row_done) {
  constexpr int V = 16 / sizeof(T);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int e = e0 + warp; e < e1; e += nw) {
    const T* p = m + (long long)e * ld;  // 16-byte aligned: mat_off and ld are multiples of 4 elements
    const int nvec = nt / V;
    for (int i0 = 0; i0 < nvec; i0 += 64) {
      const int ia = i0 + lane, ib = i0 + 32 + lane;
      const bool ina = ia < nvec, inb = ib < nvec;
      T a[V], b[V];
      if constexpr (sizeof(T) == 4) {
        const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 ra = ina ? __ldg(reinterpret_cast<const float4*>(p) + ia) : z;
        const float4 rb = inb ? __ldg(reinterpret_cast<const float4*>(p) + ib) : z;
        a[0] = ra.x, a[1] = ra.y, a[2] = ra.z, a[3] = ra.w;
        b[0] = rb.x, b[1] = rb.y, b[2] = rb.z, b[3] = rb.w;
      } else {
        const double2 z = make_double2(0.0, 0.0);
        const double2 ra = ina ? __ldg(reinterpret_cast<const double2*>(p) + ia) : z;
        const double2 rb = inb ? __ldg(reinterpret_cast<const double2*>(p) + ib) : z;
        a[0] = ra.x, a[1] = ra.y;
        b[0] = rb.x, b[1] = rb.y;
      }
#pragma unroll
      for (int v = 0; v < V; ++v) fn(a[v], ina, e);
      if (i0 + 32 < nvec) {
#pragma unroll
        for (int v = 0; v < V; ++v) fn(b[v], inb, e);
      }
    }
    const int done = nvec * V;
    if (done < nt) {
      const bool in = done + lane < nt;
      fn(in ? __ldg(p + done + lane) : T(0), in, e);
    }
    row_done(e);
  }
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    pool_hist_first_kernel(const T* __restrict__ mats, const csg_pool_item* __restrict__ items, int max_pos,
                           int bits, int max_E, uint32_t* __restrict__ hist, int32_t* __restrict__ counts,
                           int32_t* __restrict__ npos, uint32_t* __restrict__ ehist) {
  extern __shared__ unsigned s_hist[];  // [1 << bits]
  __shared__ unsigned s_red[32];
  const int nb = 1 << bits;
  const csg_pool_item it = items[blockIdx.x / kSplit];
  const int part = blockIdx.x % kSplit;
  const int per = (it.E + kSplit - 1) / kSplit;
  const int e0 = part * per, e1 = min(it.E, e0 + per);
  for (int i = threadIdx.x; i < nb; i += kThreads) s_hist[i] = 0;
  __syncthreads();
  const int shift = Key<T>::POS_BITS - bits;
  const int lane = threadIdx.x & 31;
  unsigned mine = 0, row_count = 0;
  int32_t* c = counts + (size_t)(blockIdx.x / kSplit) * max_E;
  // the same counts keyed by (instrument, position): scanned along the sequence they give the
  // per-energy totals of every prefix pool (csg_pool_energy_candidates)
  uint32_t* ec = ehist ? ehist + ((size_t)it.inst * max_pos + it.pos) * max_E : nullptr;
  walk_rows<T>(
      mats + it.mat_off, (it.T + 3) & ~3, it.T, e0, e1,
      [&](T v, bool in, int) {
        if (in && is_finite(v) && v > T(0)) {
          ++row_count;
          atomicAdd(&s_hist[(unsigned)(Key<T>::bits(v) >> shift)], 1u);
        }
      },
      [&](int e) {  // per-energy positive count (CS/fast/extrema.py:260-264): the row belongs to this warp
        const unsigned total = __reduce_add_sync(0xffffffffu, row_count);
        if (lane == 0) {
          c[e] = (int32_t)total;
          if (ec) ec[e] = total;
        }
        mine += row_count;
        row_count = 0;
      });
  auto addu = [](unsigned a, unsigned b) { return a + b; };
  mine = block_reduce(mine, addu, 0u, s_red);
  __syncthreads();
  uint32_t* h = hist + ((size_t)it.inst * max_pos + it.pos) * nb;  // slot stride = 1 slot at level 0
  for (int i = threadIdx.x; i < nb; i += kThreads) {
    const unsigned n = s_hist[i];
    if (n) atomicAdd(&h[i], n);
  }
  if (threadIdx.x == 0 && mine) atomicAdd(&npos[blockIdx.x / kSplit], (int32_t)mine);
}

template <typename T>
__global__ void __launch_bounds__(kThreads)
    pool_hist_refine_kernel(const T* __restrict__ mats, const csg_pool_item* __restrict__ items, int max_pos,
                            int n_slots, const uint64_t* __restrict__ slot_prefix, int prefix_shift, int shift,
                            int bits, uint32_t* __restrict__ hist) {
  typedef typename Key<T>::U U;
  __shared__ uint64_t s_pref[64];
  const csg_pool_item it = items[blockIdx.x / kSplit];
  const int part = blockIdx.x % kSplit;
  const int per = (it.E + kSplit - 1) / kSplit;
  const int e0 = part * per, e1 = min(it.E, e0 + per);
  for (int i = threadIdx.x; i < n_slots; i += kThreads) s_pref[i] = slot_prefix[(size_t)it.inst * n_slots + i];
  __syncthreads();
  // sorted ascending, padded with UINT64_MAX: the live prefixes span [first, last]
  int live = 0;
  while (live < n_slots && s_pref[live] != ~0ull) ++live;
  if (live == 0) return;
  const uint64_t first = s_pref[0], last = s_pref[live - 1];
  const int nb = 1 << bits;
  uint32_t* h = hist + ((size_t)it.inst * max_pos + it.pos) * (size_t)n_slots * nb;
  walk_rows<T>(
      mats + it.mat_off, (it.T + 3) & ~3, it.T, e0, e1,
      [&](T v, bool in, int) {
        if (!(in && is_finite(v) && v > T(0))) return;
        const U k = Key<T>::bits(v);
        const uint64_t hi = (uint64_t)(k >> prefix_shift);
        if (hi < first || hi > last) return;  // almost every cell
        for (int s = 0; s < live; ++s)
          if (s_pref[s] == hi) {
            atomicAdd(&h[(size_t)s * nb + ((unsigned)(k >> shift) & (unsigned)(nb - 1))], 1u);
            break;
          }
      },
      [](int) {});
}

// Inclusive scan along the file sequence.  Block = 32 consecutive (slot, bin) columns of one
// instrument x 8 segments of the sequence: every thread sums its segment, the segment totals
// are scanned in shared memory, a second walk writes the running sums (coalesced 128-byte rows).
// Slots whose table entry is the UINT64_MAX padding hold no data and are skipped.
constexpr int kScanCols = 32;
constexpr int kScanSegs = 8;

__global__ void __launch_bounds__(kScanCols* kScanSegs)
    pool_scan_kernel(uint32_t* __restrict__ hist, int n_inst, int max_pos, const int32_t* __restrict__ inst_len,
                     int n_slots, int nb, const uint64_t* __restrict__ slot_table, uint32_t* __restrict__ totals) {
  __shared__ uint32_t s_seg[kScanSegs][kScanCols + 1];
  const size_t cols = (size_t)n_slots * nb;
  const int blocks_per_inst = (int)((cols + kScanCols - 1) / kScanCols);
  const int inst = blockIdx.x / blocks_per_inst;
  const size_t col = (size_t)(blockIdx.x - inst * blocks_per_inst) * kScanCols + (threadIdx.x % kScanCols);
  const int seg = threadIdx.x / kScanCols;
  const bool in = col < cols;
  const int slot = in ? (int)(col / nb) : 0;
  const bool live = in && (slot_table == nullptr || slot_table[(size_t)inst * n_slots + slot] != ~0ull);
  const int len = inst_len[inst];
  const int per = (len + kScanSegs - 1) / kScanSegs;
  const int k0 = seg * per, k1 = min(len, k0 + per);
  uint32_t* p = hist + (size_t)inst * max_pos * cols + col;
  uint32_t sum = 0;
  if (live)
    for (int k = k0; k < k1; ++k) sum += p[(size_t)k * cols];
  s_seg[seg][threadIdx.x % kScanCols] = sum;
  __syncthreads();
  uint32_t run = 0;
  for (int g = 0; g < seg; ++g) run += s_seg[g][threadIdx.x % kScanCols];
  if (live)
    for (int k = k0; k < k1; ++k) {
      run += p[(size_t)k * cols];
      p[(size_t)k * cols] = run;
    }
  if (totals && in && seg == kScanSegs - 1) {
    uint32_t total = 0;
    for (int g = 0; g < kScanSegs; ++g) total += s_seg[g][threadIdx.x % kScanCols];
    totals[(size_t)inst * cols + col] = live ? total : 0u;
  }
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
                        int max_pos, int bits, int max_E, uint32_t* d_hist, int32_t* d_counts, int32_t* d_npos,
                        uint32_t* d_ehist) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_items <= 0) return CSG_OK;
  if (bits < 1 || bits > 12) return csg_fail(ctx, CSG_ERR_ARG, "bits %d out of range (1..12)", bits);
  if (max_E <= 0 || max_E > 8192) return csg_fail(ctx, CSG_ERR_ARG, "max_E %d out of range", max_E);
  const size_t smem = (size_t)(1 << bits) * sizeof(unsigned);
  if (int rc = csg_fill(ctx, d_npos, 0, (size_t)n_items * sizeof(int32_t))) return rc;
  if (int rc = csg_fill(ctx, d_counts, 0, (size_t)n_items * max_E * sizeof(int32_t))) return rc;
  if (dtype == CSG_F32)
    pool_hist_first_kernel<float><<<n_items * kSplit, kThreads, smem, ctx->stream>>>((const float*)d_mats, d_items, max_pos, bits,
                                                                            max_E, d_hist, d_counts, d_npos, d_ehist);
  else if (dtype == CSG_F64)
    pool_hist_first_kernel<double><<<n_items * kSplit, kThreads, smem, ctx->stream>>>((const double*)d_mats, d_items, max_pos,
                                                                             bits, max_E, d_hist, d_counts, d_npos, d_ehist);
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
    pool_hist_refine_kernel<float><<<n_items * kSplit, kThreads, 0, ctx->stream>>>((const float*)d_mats, d_items, max_pos, n_slots,
                                                                          d_slot_prefix, prefix_shift, shift, bits, d_hist);
  else if (dtype == CSG_F64)
    pool_hist_refine_kernel<double><<<n_items * kSplit, kThreads, 0, ctx->stream>>>((const double*)d_mats, d_items, max_pos,
                                                                           n_slots, d_slot_prefix, prefix_shift, shift,
                                                                           bits, d_hist);
  else
    return csg_fail(ctx, CSG_ERR_ARG, "bad dtype %d", dtype);
  CSG_LAUNCH_CHECK(ctx, "pool_hist_refine_kernel");
  return CSG_OK;
}

int csg_pool_scan(csg_ctx* ctx, uint32_t* d_hist, int n_inst, int max_pos, const int32_t* d_inst_len, int n_slots,
                  int bits, const uint64_t* d_slot_table, uint32_t* d_totals) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_inst <= 0 || max_pos <= 0) return CSG_OK;
  const size_t cols = (size_t)n_slots * ((size_t)1 << bits);
  const int blocks = n_inst * (int)((cols + kScanCols - 1) / kScanCols);
  pool_scan_kernel<<<blocks, kScanCols * kScanSegs, 0, ctx->stream>>>(d_hist, n_inst, max_pos, d_inst_len, n_slots, 1 << bits,
                                                                       d_slot_table, d_totals);
  CSG_LAUNCH_CHECK(ctx, "pool_scan_kernel");
  return CSG_OK;
}

int csg_pool_scan_cols(csg_ctx* ctx, uint32_t* d_hist, int n_inst, int max_pos, const int32_t* d_inst_len, int cols,
                       uint32_t* d_totals) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_inst <= 0 || max_pos <= 0 || cols <= 0) return CSG_OK;
  const int blocks = n_inst * ((cols + kScanCols - 1) / kScanCols);
  pool_scan_kernel<<<blocks, kScanCols * kScanSegs, 0, ctx->stream>>>(d_hist, n_inst, max_pos, d_inst_len, 1, cols, nullptr,
                                                                       d_totals);
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
