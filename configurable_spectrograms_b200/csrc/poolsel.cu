// K2b, device-resident selection -- the digit-by-digit descent of pool.cu driven entirely on
// the GPU, so one batch step needs no host round trip between radix digits.
//
// Replaces the per-step np.nanpercentile(np.concatenate(blocks so far), p) and the running
// max-merge of CS/fast/extrema.py:280-300 for EVERY prefix of an instrument's ascending-orbit
// file sequence at once.  State lives in a dense table csg_pool_sel[request][pos]: the two
// neighbour ranks numpy's linear interpolation needs (lo, hi), the key bits resolved so far
// and an `active` flag.  Per digit:
//   locate   both ranks are walked into their bucket of the scanned histogram row
//   bounds   best[request] = max over active prefixes of the smallest value lo can still take
//   slots    prefixes whose largest possible value is below best are dropped (they cannot hold
//            the running maximum); the distinct surviving key prefixes (a handful) become the
//            sorted slot table of the next refinement histogram
//   assign   (after the cross-rank all-gather of slot tables) merge, give every target its slot
// After the last digit the prefix IS the order statistic; finish interpolates with numpy's
// arithmetic in D and reduces per request.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr unsigned long long kNone = ~0ull;

enum { FLAG_LOCATE_MISS = 0, FLAG_SLOT_OVERFLOW = 1, FLAG_ASSIGN_MISS = 2 };
constexpr int kBestSlots = 64;  // int64 words reserved for the per-request bounds ahead of the slot lists

__global__ void __launch_bounds__(kThreads)
    pool_row_totals_kernel(const uint32_t* __restrict__ hist, int max_pos, int nb, const uint32_t* __restrict__ base,
                           int64_t* __restrict__ n_after, int64_t* __restrict__ below) {
  __shared__ long long s_a[32], s_b[32];
  const int pos = blockIdx.x, inst = blockIdx.y;
  const uint32_t* row = hist + ((size_t)inst * max_pos + pos) * nb;
  const uint32_t* brow = base ? base + (size_t)inst * nb : nullptr;
  long long s = 0, b = 0;
  for (int i = threadIdx.x; i < nb; i += kThreads) {
    s += row[i];
    if (brow) b += brow[i];
  }
  auto add = [](long long x, long long y) { return x + y; };
  s = block_reduce(s, add, 0ll, s_a);
  b = block_reduce(b, add, 0ll, s_b);
  if (threadIdx.x == 0) {
    n_after[(size_t)inst * max_pos + pos] = s + b;
    if (pos == 0) below[inst] = b;
  }
}

template <typename T>
__global__ void pool_sel_init_kernel(const csg_pool_request* __restrict__ reqs, int n_req,
                                     const int32_t* __restrict__ inst_len, int max_pos,
                                     const int64_t* __restrict__ n_after, const int64_t* __restrict__ below,
                                     const int64_t* __restrict__ above, csg_pool_sel* __restrict__ sel) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_req * max_pos) return;
  const int r = idx / max_pos, pos = idx - r * max_pos;
  const csg_pool_request rq = reqs[r];
  const int L = inst_len[rq.inst];
  csg_pool_sel s;
  s.inst = rq.inst, s.pos = pos, s.req = r, s.active = 0;
  s.slot[0] = s.slot[1] = 0;
  s.rank[0] = s.rank[1] = 0;
  s.prefix[0] = s.prefix[1] = 0;
  s.gamma = 0.0;
  if (pos < L) {
    const long long n = n_after[(size_t)rq.inst * max_pos + pos];
    const long long lower = below[rq.inst];
    const long long prev = pos ? n_after[(size_t)rq.inst * max_pos + pos - 1] : lower;
    bool active;
    if (rq.mode == 0)  // running maximum: a file without positives repeats the previous candidate
      active = n > 0 && n != prev;
    else  // whole pool: asked once, by the last rank that holds positives
      active = pos == L - 1 && (n - lower) > 0 && (above == nullptr || above[rq.inst] == 0);
    if (active) {
      long long lo, hi;
      T g;
      percentile_ranks<T>(n, rq.p, lo, hi, g);
      s.rank[0] = lo, s.rank[1] = hi, s.gamma = (double)g, s.active = 1;
    }
  }
  sel[idx] = s;
}

// block = one neighbour (j) of one table entry: walk its rank into the bucket of its scanned row
__global__ void __launch_bounds__(kThreads)
    pool_sel_locate_kernel(const uint32_t* __restrict__ hist, int max_pos, int n_slots, int nb, int bits,
                           const uint32_t* __restrict__ base, csg_pool_sel* __restrict__ sel, int32_t* __restrict__ flags) {
  __shared__ long long s_warp[32];
  csg_pool_sel* s = sel + (blockIdx.x >> 1);
  const int j = blockIdx.x & 1;
  if (!s->active) return;
  const int inst = s->inst, slot = s->slot[j];
  const long long want = s->rank[j];
  const uint32_t* row = hist + (((size_t)inst * max_pos + s->pos) * n_slots + slot) * (size_t)nb;
  const uint32_t* brow = base ? base + ((size_t)inst * n_slots + slot) * (size_t)nb : nullptr;
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
  if (threadIdx.x == 0 && (want < 0 || want >= total)) atomicOr(&flags[FLAG_LOCATE_MISS], 1);
  if (want >= excl && want < excl + mine) {
    long long run = excl;
    for (int b = b0; b < b0 + per && b < nb; ++b) {
      const long long c = cell(b);
      if (want < run + c) {
        s->prefix[j] = (s->prefix[j] << bits) | (uint64_t)b;
        s->rank[j] = want - run;
        break;
      }
      run += c;
    }
  }
}

// best[r] = max over active running-max entries of (prefix_lo << shift): the smallest key the
// lower neighbour can still resolve to (positive floats order like their bit patterns)
__global__ void __launch_bounds__(1024)
    pool_sel_bounds_kernel(const csg_pool_sel* __restrict__ sel, const csg_pool_request* __restrict__ reqs, int n_req,
                           int max_pos, int shift, int64_t* __restrict__ best) {
  __shared__ long long s_best[64];
  for (int i = threadIdx.x; i < n_req; i += blockDim.x) s_best[i] = -1;
  __syncthreads();
  const int n = n_req * max_pos;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (!sel[i].active) continue;
    const int r = sel[i].req;
    if (reqs[r].mode != 0) continue;
    atomicMax(&s_best[r], (long long)(sel[i].prefix[0] << shift));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_req; i += blockDim.x) best[i] = s_best[i];
}

// ascending distinct values of a multiset, by repeated block-wide minimum (the result is tiny)
template <typename Visit>
__device__ int distinct_ascending(Visit visit, unsigned long long* s_min, uint64_t* out, int cap, bool* overflow) {
  int count = 0;
  unsigned long long last = 0;
  *overflow = false;
  while (true) {
    __syncthreads();
    if (threadIdx.x == 0) *s_min = kNone;
    __syncthreads();
    visit([&](unsigned long long v) {
      if (v != kNone && (count == 0 || v > last)) atomicMin(s_min, v);
    });
    __syncthreads();
    const unsigned long long m = *s_min;
    if (m == kNone) break;
    if (count == cap) {
      *overflow = true;
      break;
    }
    if (threadIdx.x == 0) out[count] = m;
    ++count;
    last = m;
  }
  __syncthreads();
  for (int i = count + threadIdx.x; i < cap; i += blockDim.x) out[i] = kNone;
  return count;
}

// block = one instrument: prune, then list the distinct surviving prefixes (sorted, padded)
__global__ void __launch_bounds__(kThreads)
    pool_sel_slots_kernel(csg_pool_sel* __restrict__ sel, const csg_pool_request* __restrict__ reqs, int n_req, int max_pos,
                          int shift, const int64_t* __restrict__ best, int n_slots, uint64_t* __restrict__ local_slots,
                          int32_t* __restrict__ flags) {
  __shared__ unsigned long long s_min;
  const int inst = blockIdx.x;
  const int n = n_req * max_pos;
  for (int i = threadIdx.x; i < n; i += kThreads) {
    if (!sel[i].active || sel[i].inst != inst) continue;
    const int r = sel[i].req;
    if (reqs[r].mode != 0) continue;
    const long long hi_key = (long long)(((sel[i].prefix[1] + 1ull) << shift) - 1ull);
    if (hi_key < best[r]) sel[i].active = 0;  // cannot reach the running maximum any more
  }
  __syncthreads();
  bool overflow;
  distinct_ascending(
      [&](auto&& emit) {
        for (int i = threadIdx.x; i < n; i += kThreads) {
          if (!sel[i].active || sel[i].inst != inst) continue;
          emit(sel[i].prefix[0]);
          emit(sel[i].prefix[1]);
        }
      },
      &s_min, local_slots + (size_t)inst * n_slots, n_slots, &overflow);
  if (overflow && threadIdx.x == 0) atomicOr(&flags[FLAG_SLOT_OVERFLOW], 1);
}

// block = one instrument.  gathered: n_ranks payloads (byte stride `stride`) of
// [best: kBestSlots int64][local slot lists: n_inst x n_slots uint64].  The global best lower
// bound per request is the maximum over the ranks; prefixes below it are dropped, the ranks'
// slot lists are merged into the table and every surviving target gets its slot.
__global__ void __launch_bounds__(kThreads)
    pool_sel_assign_kernel(csg_pool_sel* __restrict__ sel, const csg_pool_request* __restrict__ reqs, int n_req, int max_pos,
                           int shift, const unsigned char* __restrict__ gathered, size_t stride, int n_ranks, int n_inst,
                           int n_slots, uint64_t* __restrict__ table, int32_t* __restrict__ flags) {
  __shared__ unsigned long long s_min;
  __shared__ uint64_t s_table[64];
  __shared__ long long s_best[64];
  const int inst = blockIdx.x;
  const int n = n_req * max_pos;
  for (int r = threadIdx.x; r < n_req; r += kThreads) {
    long long b = -1;
    for (int rk = 0; rk < n_ranks; ++rk)
      b = max(b, reinterpret_cast<const long long*>(gathered + (size_t)rk * stride)[r]);
    s_best[r] = b;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kThreads) {
    if (!sel[i].active || sel[i].inst != inst) continue;
    const int r = sel[i].req;
    if (reqs[r].mode != 0) continue;
    const long long hi_key = (long long)(((sel[i].prefix[1] + 1ull) << shift) - 1ull);
    if (hi_key < s_best[r]) sel[i].active = 0;  // cannot reach the (global) running maximum any more
  }
  __syncthreads();
  bool overflow;
  uint64_t* out = table + (size_t)inst * n_slots;
  const int count = distinct_ascending(
      [&](auto&& emit) {
        for (int i = threadIdx.x; i < n_ranks * n_slots; i += kThreads) {
          const int rk = i / n_slots, k = i - rk * n_slots;
          const uint64_t* lists = reinterpret_cast<const uint64_t*>(gathered + (size_t)rk * stride + kBestSlots * 8);
          emit(lists[(size_t)inst * n_slots + k]);
        }
      },
      &s_min, out, n_slots, &overflow);
  if (overflow && threadIdx.x == 0) atomicOr(&flags[FLAG_SLOT_OVERFLOW], 1);
  __syncthreads();
  for (int i = threadIdx.x; i < n_slots; i += kThreads) s_table[i] = out[i];
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += kThreads) {
    if (!sel[i].active || sel[i].inst != inst) continue;
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      const uint64_t p = sel[i].prefix[j];
      int slot = -1;
      for (int k = 0; k < count; ++k)
        if (s_table[k] == p) slot = k;
      if (slot < 0) {
        atomicOr(&flags[FLAG_ASSIGN_MISS], 1);
        slot = 0;
      }
      sel[i].slot[j] = slot;
    }
  }
}

// interpolate the exact neighbours with numpy's arithmetic in T; per request the maximum over
// the surviving prefixes ("last" requests hold a single entry)
template <typename T>
__global__ void __launch_bounds__(1024)
    pool_sel_finish_kernel(const csg_pool_sel* __restrict__ sel, int n_req, int max_pos, double* __restrict__ values,
                           int32_t* __restrict__ has) {
  typedef typename Key<T>::U U;
  __shared__ long long s_val[64];
  for (int i = threadIdx.x; i < n_req; i += blockDim.x) s_val[i] = -1;
  __syncthreads();
  const int n = n_req * max_pos;
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (!sel[i].active) continue;
    T a, b;
    if (sizeof(T) == 4) {
      a = (T)__uint_as_float((unsigned)sel[i].prefix[0]);
      b = (T)__uint_as_float((unsigned)sel[i].prefix[1]);
    } else {
      a = (T)__longlong_as_double((long long)sel[i].prefix[0]);
      b = (T)__longlong_as_double((long long)sel[i].prefix[1]);
    }
    const T v = numpy_lerp<T>(a, b, (T)sel[i].gamma);
    const double d = (double)v;  // finite and positive: the bit pattern orders like the value
    atomicMax(&s_val[sel[i].req], __double_as_longlong(d));
  }
  __syncthreads();
  for (int i = threadIdx.x; i < n_req; i += blockDim.x) {
    const bool ok = s_val[i] >= 0;
    values[i] = ok ? __longlong_as_double(s_val[i]) : -CUDART_INF;
    has[i] = ok ? 1 : 0;
  }
}

// base[c] = sum over lower ranks of gathered[rank][c]; above[inst] = cells held by higher ranks.
// `stride`: uint32 elements between two ranks' payloads (>= n when the payload carries more).
__global__ void pool_base_kernel(const uint32_t* __restrict__ gathered, size_t stride, int rank, size_t n,
                                 uint32_t* __restrict__ base) {
  const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t s = 0;
  for (int r = 0; r < rank; ++r) s += gathered[(size_t)r * stride + i];
  base[i] = s;
}
__global__ void __launch_bounds__(kThreads)
    pool_above_kernel(const uint32_t* __restrict__ gathered, size_t stride, int n_ranks, int rank, size_t cols,
                      int64_t* __restrict__ above) {
  __shared__ long long s_a[32];
  const int inst = blockIdx.x;
  long long s = 0;
  for (int r = rank + 1; r < n_ranks; ++r)
    for (size_t c = threadIdx.x; c < cols; c += kThreads) s += gathered[(size_t)r * stride + (size_t)inst * cols + c];
  auto add = [](long long x, long long y) { return x + y; };
  s = block_reduce(s, add, 0ll, s_a);
  if (threadIdx.x == 0) above[inst] = s;
}

// ---- y extrema: the 99 %-coverage energy of every prefix pool (CS/fast/extrema.py:270-278)
// block = (position, instrument).  v_j = positive cells of energy key j in the pool after this
// file (scanned per-energy counts + what lower ranks hold); numpy: cum = cumsum(v) over the
// ascending keys, target = 0.99 * total (float64), idx = searchsorted(cum, target, 'right') =
// first j with cum_j > target.  The candidate (0.0 for an empty pool) is max-merged per
// instrument as an order-preserving int64 key.
constexpr long long kNoCandidate = (long long)0x8080808080808080ull;  // cudaMemset(0x80) pattern

__device__ __forceinline__ long long ordered_key(double v) {
  const long long k = __double_as_longlong(v);
  return k < 0 ? (k ^ 0x7fffffffffffffffll) : k;
}
__device__ __forceinline__ double ordered_value(long long k) {
  return __longlong_as_double(k < 0 ? (k ^ 0x7fffffffffffffffll) : k);
}

__global__ void __launch_bounds__(128)
    pool_ecand_kernel(const uint32_t* __restrict__ ehist, int max_pos, int max_E, const int32_t* __restrict__ order,
                      const double* __restrict__ keys, const int32_t* __restrict__ n_keys,
                      const int32_t* __restrict__ limit, const uint32_t* __restrict__ g_etot, size_t g_stride, int rank,
                      long long* __restrict__ ycand) {
  __shared__ long long s_red[32];
  __shared__ long long s_carry;
  __shared__ int s_first;
  const int pos = blockIdx.x, inst = blockIdx.y;
  if (pos >= limit[inst]) return;
  const int nk = n_keys[inst];
  const uint32_t* row = ehist + ((size_t)inst * max_pos + pos) * max_E;
  const int32_t* ord = order + (size_t)inst * max_E;
  auto cell = [&](int j) -> long long {
    const int e = ord[j];
    long long v = row[e];
    for (int r = 0; r < rank; ++r) v += g_etot[(size_t)r * g_stride + (size_t)inst * max_E + e];
    return v;
  };
  long long mine = 0;
  for (int j = threadIdx.x; j < nk; j += blockDim.x) mine += cell(j);
  auto add = [](long long a, long long b) { return a + b; };
  const long long total = block_reduce(mine, add, 0ll, s_red);
  double cand = 0.0;
  if (total > 0) {
    const double target = 0.99 * (double)total;
    if (threadIdx.x == 0) s_carry = 0, s_first = nk;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    for (int j0 = 0; j0 < nk; j0 += blockDim.x) {
      const int j = j0 + threadIdx.x;
      const long long v = j < nk ? cell(j) : 0;
      long long inc = v;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const long long t = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += t;
      }
      __syncthreads();  // s_red free (block_reduce / previous chunk)
      if (lane == 31) s_red[warp] = inc;
      __syncthreads();
      long long off = s_carry;
      for (int w = 0; w < warp; ++w) off += s_red[w];
      const long long cum = off + inc;
      if (j < nk && (double)cum > target) atomicMin(&s_first, j);
      __syncthreads();
      if (threadIdx.x == 0) {
        long long c = s_carry;
        for (int w = 0; w < nw; ++w) c += s_red[w];
        s_carry = c;
      }
      __syncthreads();
      if (s_first < nk) break;  // uniform: every later cumulative sum is larger still
    }
    cand = keys[(size_t)inst * max_E + s_first];
  }
  if (threadIdx.x == 0) atomicMax(&ycand[inst], ordered_key(cand));
}

// out = [values (n_req) | y candidates (n_inst, -inf = none) | selection flags (4) | peer error | pad]
__global__ void pool_pack_kernel(const double* __restrict__ values, int n_req, const long long* __restrict__ ycand, int n_inst,
                                 const int32_t* __restrict__ flags, const int* __restrict__ peer_error,
                                 double* __restrict__ out, int n_out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_out) return;
  double v = 0.0;
  if (i < n_req)
    v = values[i];
  else if (i < n_req + n_inst) {
    const long long k = ycand ? ycand[i - n_req] : kNoCandidate;
    v = k == kNoCandidate ? -CUDART_INF : ordered_value(k);
  } else if (i < n_req + n_inst + 4)
    v = (double)flags[i - n_req - n_inst];
  else if (i == n_req + n_inst + 4)
    v = peer_error ? (double)*peer_error : 0.0;
  out[i] = v;
}

__global__ void pool_max_kernel(const double* __restrict__ gathered, int n_ranks, int n, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double v = gathered[i];
  for (int r = 1; r < n_ranks; ++r) v = fmax(v, gathered[(size_t)r * n + i]);
  out[i] = v;
}

}  // namespace

extern "C" {

int csg_pool_row_totals(csg_ctx* ctx, const uint32_t* d_hist, int n_inst, int max_pos, int bits, const uint32_t* d_base,
                        int64_t* d_n_after, int64_t* d_below) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_inst <= 0 || max_pos <= 0) return CSG_OK;
  if (!d_hist || !d_n_after || !d_below) return csg_fail(ctx, CSG_ERR_ARG, "NULL argument");
  pool_row_totals_kernel<<<dim3(max_pos, n_inst), kThreads, 0, ctx->stream>>>(d_hist, max_pos, 1 << bits, d_base, d_n_after,
                                                                             d_below);
  CSG_LAUNCH_CHECK(ctx, "pool_row_totals_kernel");
  return CSG_OK;
}

int csg_pool_sel_init(csg_ctx* ctx, int dtype, const csg_pool_request* d_requests, int n_req, const int32_t* d_inst_len,
                      int max_pos, const int64_t* d_n_after, const int64_t* d_below, const int64_t* d_above,
                      csg_pool_sel* d_sel) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_req <= 0 || max_pos <= 0) return CSG_OK;
  if (n_req > 64) return csg_fail(ctx, CSG_ERR_ARG, "at most 64 percentile requests per selection (got %d)", n_req);
  if (!d_requests || !d_inst_len || !d_n_after || !d_below || !d_sel) return csg_fail(ctx, CSG_ERR_ARG, "NULL argument");
  const int n = n_req * max_pos, blocks = (n + 255) / 256;
  if (dtype == CSG_F32)
    pool_sel_init_kernel<float><<<blocks, 256, 0, ctx->stream>>>(d_requests, n_req, d_inst_len, max_pos, d_n_after, d_below,
                                                                 d_above, d_sel);
  else if (dtype == CSG_F64)
    pool_sel_init_kernel<double><<<blocks, 256, 0, ctx->stream>>>(d_requests, n_req, d_inst_len, max_pos, d_n_after, d_below,
                                                                  d_above, d_sel);
  else
    return csg_fail(ctx, CSG_ERR_ARG, "bad dtype %d", dtype);
  CSG_LAUNCH_CHECK(ctx, "pool_sel_init_kernel");
  return CSG_OK;
}

int csg_pool_sel_locate(csg_ctx* ctx, const uint32_t* d_hist, int max_pos, int n_slots, int bits, const uint32_t* d_base,
                        csg_pool_sel* d_sel, int n_req, int32_t* d_flags) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_req <= 0 || max_pos <= 0) return CSG_OK;
  if (!d_hist || !d_sel || !d_flags) return csg_fail(ctx, CSG_ERR_ARG, "NULL argument");
  pool_sel_locate_kernel<<<2 * n_req * max_pos, kThreads, 0, ctx->stream>>>(d_hist, max_pos, n_slots, 1 << bits, bits, d_base,
                                                                            d_sel, d_flags);
  CSG_LAUNCH_CHECK(ctx, "pool_sel_locate_kernel");
  return CSG_OK;
}

int csg_pool_sel_bounds(csg_ctx* ctx, const csg_pool_sel* d_sel, const csg_pool_request* d_requests, int n_req, int max_pos,
                        int shift, int64_t* d_best) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_req <= 0 || max_pos <= 0) return CSG_OK;
  if (n_req > 64) return csg_fail(ctx, CSG_ERR_ARG, "at most 64 percentile requests per selection (got %d)", n_req);
  if (!d_sel || !d_requests || !d_best) return csg_fail(ctx, CSG_ERR_ARG, "NULL argument");
  pool_sel_bounds_kernel<<<1, 1024, 0, ctx->stream>>>(d_sel, d_requests, n_req, max_pos, shift, d_best);
  CSG_LAUNCH_CHECK(ctx, "pool_sel_bounds_kernel");
  return CSG_OK;
}

int csg_pool_sel_slots(csg_ctx* ctx, csg_pool_sel* d_sel, const csg_pool_request* d_requests, int n_req, int max_pos,
                       int shift, const int64_t* d_best, int n_inst, int n_slots, uint64_t* d_local_slots,
                       int32_t* d_flags) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_inst <= 0) return CSG_OK;
  if (n_slots < 1 || n_slots > 64) return csg_fail(ctx, CSG_ERR_ARG, "n_slots %d out of range (1..64)", n_slots);
  if (!d_sel || !d_requests || !d_best || !d_local_slots || !d_flags) return csg_fail(ctx, CSG_ERR_ARG, "NULL argument");
  pool_sel_slots_kernel<<<n_inst, kThreads, 0, ctx->stream>>>(d_sel, d_requests, n_req, max_pos > 0 ? max_pos : 0, shift, d_best,
                                                              n_slots, d_local_slots, d_flags);
  CSG_LAUNCH_CHECK(ctx, "pool_sel_slots_kernel");
  return CSG_OK;
}

int csg_pool_sel_assign(csg_ctx* ctx, csg_pool_sel* d_sel, const csg_pool_request* d_requests, int n_req, int max_pos,
                        int shift, const void* d_gathered, size_t rank_stride_bytes, int n_ranks, int n_inst, int n_slots,
                        uint64_t* d_table, int32_t* d_flags) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_inst <= 0) return CSG_OK;
  if (n_slots < 1 || n_slots > 64) return csg_fail(ctx, CSG_ERR_ARG, "n_slots %d out of range (1..64)", n_slots);
  if (n_req > 64) return csg_fail(ctx, CSG_ERR_ARG, "at most 64 percentile requests per selection (got %d)", n_req);
  if (n_ranks < 1) return csg_fail(ctx, CSG_ERR_ARG, "n_ranks %d < 1", n_ranks);
  if (!d_sel || !d_requests || !d_gathered || !d_table || !d_flags) return csg_fail(ctx, CSG_ERR_ARG, "NULL argument");
  pool_sel_assign_kernel<<<n_inst, kThreads, 0, ctx->stream>>>(d_sel, d_requests, n_req, max_pos > 0 ? max_pos : 0, shift,
                                                               (const unsigned char*)d_gathered, rank_stride_bytes, n_ranks,
                                                               n_inst, n_slots, d_table, d_flags);
  CSG_LAUNCH_CHECK(ctx, "pool_sel_assign_kernel");
  return CSG_OK;
}

size_t csg_pool_slot_payload_bytes(int n_inst, int n_slots) { return (size_t)kBestSlots * 8 + (size_t)n_inst * n_slots * 8; }

int csg_pool_sel_finish(csg_ctx* ctx, int dtype, const csg_pool_sel* d_sel, int n_req, int max_pos, double* d_values,
                        int32_t* d_has) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_req <= 0) return CSG_OK;
  if (n_req > 64) return csg_fail(ctx, CSG_ERR_ARG, "at most 64 percentile requests per selection (got %d)", n_req);
  if (!d_sel || !d_values || !d_has) return csg_fail(ctx, CSG_ERR_ARG, "NULL argument");
  if (dtype == CSG_F32)
    pool_sel_finish_kernel<float><<<1, 1024, 0, ctx->stream>>>(d_sel, n_req, max_pos > 0 ? max_pos : 0, d_values, d_has);
  else if (dtype == CSG_F64)
    pool_sel_finish_kernel<double><<<1, 1024, 0, ctx->stream>>>(d_sel, n_req, max_pos > 0 ? max_pos : 0, d_values, d_has);
  else
    return csg_fail(ctx, CSG_ERR_ARG, "bad dtype %d", dtype);
  CSG_LAUNCH_CHECK(ctx, "pool_sel_finish_kernel");
  return CSG_OK;
}

int csg_pool_base(csg_ctx* ctx, const uint32_t* d_gathered, size_t rank_stride, int n_ranks, int rank, int n_inst,
                  size_t cols_per_inst, uint32_t* d_base, int64_t* d_above) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_inst <= 0 || cols_per_inst == 0) return CSG_OK;
  if (!d_gathered || !d_base || rank < 0 || rank >= n_ranks) return csg_fail(ctx, CSG_ERR_ARG, "bad argument");
  const size_t n = (size_t)n_inst * cols_per_inst;
  if (rank_stride < n) return csg_fail(ctx, CSG_ERR_ARG, "rank stride smaller than the payload");
  pool_base_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(d_gathered, rank_stride, rank, n, d_base);
  CSG_LAUNCH_CHECK(ctx, "pool_base_kernel");
  if (d_above) {
    pool_above_kernel<<<n_inst, kThreads, 0, ctx->stream>>>(d_gathered, rank_stride, n_ranks, rank, cols_per_inst, d_above);
    CSG_LAUNCH_CHECK(ctx, "pool_above_kernel");
  }
  return CSG_OK;
}

int csg_pool_energy_candidates(csg_ctx* ctx, const uint32_t* d_ehist, int n_inst, int max_pos, int max_E,
                               const int32_t* d_order, const double* d_keys, const int32_t* d_n_keys,
                               const int32_t* d_limit, const uint32_t* d_gathered_etot, size_t rank_stride, int rank,
                               int64_t* d_ycand) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_inst <= 0) return CSG_OK;
  if (!d_ehist || !d_order || !d_keys || !d_n_keys || !d_limit || !d_ycand) return csg_fail(ctx, CSG_ERR_ARG, "NULL argument");
  if (rank > 0 && !d_gathered_etot) return csg_fail(ctx, CSG_ERR_ARG, "rank %d needs the lower ranks' energy totals", rank);
  if (int rc = csg_fill(ctx, d_ycand, 0x80, (size_t)n_inst * sizeof(int64_t))) return rc;
  if (max_pos <= 0) return CSG_OK;
  pool_ecand_kernel<<<dim3(max_pos, n_inst), 128, 0, ctx->stream>>>(d_ehist, max_pos, max_E, d_order, d_keys, d_n_keys, d_limit,
                                                                    d_gathered_etot, rank_stride, rank, (long long*)d_ycand);
  CSG_LAUNCH_CHECK(ctx, "pool_ecand_kernel");
  return CSG_OK;
}

int csg_pool_pack_results(csg_ctx* ctx, const double* d_values, int n_req, const int64_t* d_ycand, int n_inst,
                          const int32_t* d_flags, const void* d_peer_error, double* d_out, int n_out) {
  if (!ctx) return CSG_ERR_ARG;
  if (!d_values || !d_flags || !d_out || n_out < n_req + n_inst + 5) return csg_fail(ctx, CSG_ERR_ARG, "bad argument");
  pool_pack_kernel<<<(n_out + 127) / 128, 128, 0, ctx->stream>>>(d_values, n_req, (const long long*)d_ycand, n_inst, d_flags,
                                                                 (const int*)d_peer_error, d_out, n_out);
  CSG_LAUNCH_CHECK(ctx, "pool_pack_kernel");
  return CSG_OK;
}

int csg_pool_reduce_max(csg_ctx* ctx, const double* d_gathered, int n_ranks, int n, double* d_out) {
  if (!ctx) return CSG_ERR_ARG;
  if (n <= 0) return CSG_OK;
  if (!d_gathered || !d_out || n_ranks < 1) return csg_fail(ctx, CSG_ERR_ARG, "bad argument");
  pool_max_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(d_gathered, n_ranks, n, d_out);
  CSG_LAUNCH_CHECK(ctx, "pool_max_kernel");
  return CSG_OK;
}

}  // extern "C"
