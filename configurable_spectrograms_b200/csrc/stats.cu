// K2a -- per-region exact percentiles + min/max reductions.
//
// Replaces compute_percentile_bounds -> np.nanpercentile(matrix, p) (CS/percentile_utils.py:87-88;
// called at CS/fast/plotting.py:134,286 and CS/plotting.py:259) and the reductions
// safe_vmin = nanmin(matrix[isfinite & > 0]) (CS/plotting.py:261-262), nanmin/nanmax (:314-315).
//
// One thread block per region.  Selection is an MSD radix select on the order-preserving key of
// the dtype: a histogram pass per digit (11 bits), the bucket holding each wanted rank is
// followed into the next digit; after the last digit the key IS the order statistic, so the
// result is exact.  The two neighbours are then interpolated with numpy's float arithmetic
// (q = D(p)/D(100); v = D(n-1)*q; lerp rounded after every operation -- SURVEY.md Appendix B).
//
// The first digit sees every cell: each warp owns a private histogram copy and the elected
// lane of every distinct bin updates it with a plain read-modify-write (spectrogram counts
// repeat heavily; shared-memory atomics on a handful of hot bins serialise).  Later digits
// only touch the few cells inside the followed buckets and skip whole warps otherwise.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kDigitBits = 11;
constexpr int kBins = 1 << kDigitBits;
constexpr int kTargets = 4;     // (lo, hi) neighbours of two percentiles
constexpr int kMaxCols = 1024;  // column lists up to this length are staged in shared memory
static_assert(kWarps >= kTargets, "private copies are reused as target histograms");

__device__ __forceinline__ void hist_add_private(unsigned* wh, unsigned bin, bool valid) {
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  if (valid) {
    const unsigned peers = __match_any_sync(act, bin);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) wh[bin] += (unsigned)__popc(peers);
  }
  __syncwarp();
}

template <typename T, bool HEAVY>
__global__ void __launch_bounds__(kThreads)
    region_stats_kernel(const T* __restrict__ mats, const csg_region* __restrict__ regions,
                        const int32_t* __restrict__ pool, csg_region_stats* __restrict__ out) {
  typedef typename Key<T>::U U;
  extern __shared__ unsigned s_dyn[];  // HEAVY: [kWarps][kBins] private copies, later [kTargets][kBins]
  __shared__ int s_cols[kMaxCols];
  __shared__ long long s_ll[32];
  __shared__ T s_t[32];
  __shared__ unsigned s_u[32];
  __shared__ U s_prefix[kTargets];
  __shared__ long long s_rank[kTargets];
  __shared__ int s_hidx[kTargets];

  const csg_region rg = regions[blockIdx.x];
  if (rg.want_pct != (HEAVY ? 1 : 0)) return;  // the other specialisation handles it / geometry only
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const bool cols_in_smem = rg.ne <= kMaxCols;
  const int ne = rg.ne, nt = rg.nt;

  if (HEAVY)
    for (int i = tid; i < kWarps * kBins; i += kThreads) s_dyn[i] = 0;
  if (cols_in_smem)
    for (int i = tid; i < ne; i += kThreads) s_cols[i] = __ldg(pool + rg.cols_off + i);
  __syncthreads();

  // Walk the region warp-per-time-row (two rows in flight): lanes stride over the energy
  // columns, which are (nearly) contiguous in the collapsed (T,E) matrix -> coalesced and
  // division-free.  fn(v, in) is called with warp-uniform control flow.
  auto for_each_cell = [&](auto&& fn) {
    for (int r = warp; r < nt; r += 2 * kWarps) {
      const int r2 = r + kWarps;
      const bool two = r2 < nt;
      const int rowa = rg.rows_off < 0 ? rg.t0 + r : __ldg(pool + rg.rows_off + r);
      const int rowb = two ? (rg.rows_off < 0 ? rg.t0 + r2 : __ldg(pool + rg.rows_off + r2)) : rowa;
      const T* pa = mats + rg.mat_off + (long long)rowa * rg.ld;
      const T* pb = mats + rg.mat_off + (long long)rowb * rg.ld;
      for (int c0 = 0; c0 < ne; c0 += 96) {
        T va[3], vb[3];
        bool in[3];
#pragma unroll
        for (int u = 0; u < 3; ++u) {
          const int c = c0 + u * 32 + lane;
          in[u] = c < ne;
          va[u] = vb[u] = T(0);
          if (in[u]) {
            const int col = cols_in_smem ? s_cols[c] : __ldg(pool + rg.cols_off + c);
            va[u] = __ldg(pa + col);
            vb[u] = __ldg(pb + col);
          }
        }
#pragma unroll
        for (int u = 0; u < 3; ++u)
          if (c0 + u * 32 < ne) {
            fn(va[u], in[u]);
            if (two) fn(vb[u], in[u]);
          }
      }
    }
  };

  // ---- pass 0: classification (+ first digit into the warp-private copy)
  unsigned n_valid_w = 0, n_nan_w = 0, n_pos_w = 0, n_pinf_w = 0, n_ninf_w = 0;  // warp-uniform counters
  const T kInf = (T)CUDART_INF;
  T min_pos = kInf, fin_min = kInf, fin_max = -kInf;
  constexpr int kTopShift = Key<T>::BITS - kDigitBits;
  unsigned* my_hist = HEAVY ? s_dyn + warp * kBins : nullptr;
  for_each_cell([&](T v, bool in) {
    const bool valid = in && !is_nan(v);
    const bool fin = valid && is_finite(v);
    const bool pos = v > T(0);
    n_valid_w += __popc(__ballot_sync(0xffffffffu, valid));
    n_nan_w += __popc(__ballot_sync(0xffffffffu, in && !valid));
    n_pos_w += __popc(__ballot_sync(0xffffffffu, fin && pos));
    const unsigned infs = __ballot_sync(0xffffffffu, valid && !fin);
    if (infs) {  // rare
      n_pinf_w += __popc(__ballot_sync(0xffffffffu, valid && !fin && pos));
      n_ninf_w += __popc(__ballot_sync(0xffffffffu, valid && !fin && !pos));
    }
    const T f = fin ? v : kInf;
    fin_min = f < fin_min ? f : fin_min;
    const T g = fin ? v : -kInf;
    fin_max = g > fin_max ? g : fin_max;
    const T h = (fin && pos) ? v : kInf;
    min_pos = h < min_pos ? h : min_pos;
    if (HEAVY) hist_add_private(my_hist, valid ? (unsigned)(Key<T>::key(v) >> kTopShift) : 0u, valid);
  });
  auto addll = [](long long a, long long b) { return a + b; };
  auto addu = [](unsigned a, unsigned b) { return a + b; };
  auto mint = [](T a, T b) { return a < b ? a : b; };
  auto maxt = [](T a, T b) { return a > b ? a : b; };
  const bool lead = lane == 0;
  const long long n_valid = block_reduce((long long)(lead ? n_valid_w : 0u), addll, 0ll, s_ll);
  const unsigned n_nan = block_reduce(lead ? n_nan_w : 0u, addu, 0u, s_u);
  const unsigned n_pos = block_reduce(lead ? n_pos_w : 0u, addu, 0u, s_u);
  const unsigned n_pinf = block_reduce(lead ? n_pinf_w : 0u, addu, 0u, s_u);
  const unsigned n_ninf = block_reduce(lead ? n_ninf_w : 0u, addu, 0u, s_u);
  min_pos = block_reduce(min_pos, mint, kInf, s_t);
  fin_min = block_reduce(fin_min, mint, kInf, s_t);
  fin_max = block_reduce(fin_max, maxt, (T)(-kInf), s_t);

  csg_region_stats st;
  st.p_lo = st.p_hi = CUDART_NAN;
  st.min_pos = (double)min_pos;
  st.fin_min = (double)fin_min;
  st.fin_max = (double)fin_max;
  st.n_valid = n_valid;
  st.n_nan = (int)n_nan, st.n_neginf = (int)n_ninf, st.n_posinf = (int)n_pinf, st.n_pos = (int)n_pos;

  if (!HEAVY || n_valid == 0) {
    if (tid == 0) out[blockIdx.x] = st;
    return;
  }
  unsigned(*s_hist)[kBins] = reinterpret_cast<unsigned(*)[kBins]>(s_dyn);
  // fold the warp-private copies into histogram 0
  __syncthreads();
  for (int b = tid; b < kBins; b += kThreads) {
    unsigned sum = 0;
#pragma unroll
    for (int w = 0; w < kWarps; ++w) sum += s_dyn[w * kBins + b];
    s_dyn[b] = sum;  // thread b only ever touches column b of every copy
  }

  // ---- wanted ranks
  long long rank[kTargets];
  T gamma[2];
  percentile_ranks<T>(n_valid, rg.p_lo, rank[0], rank[1], gamma[0]);
  percentile_ranks<T>(n_valid, rg.p_hi, rank[2], rank[3], gamma[1]);
  if (tid < kTargets) {
    s_prefix[tid] = 0;
    s_rank[tid] = rank[tid];
    s_hidx[tid] = 0;
  }
  __syncthreads();

  // ---- digit loop
  int shift = kTopShift;
  int bits = kDigitBits;
  while (true) {
    const int nb = 1 << bits;
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
          const long long c = h[b];
          if (want_rank < run + c) {
            s_prefix[j] = (s_prefix[j] << bits) | (U)b;
            s_rank[j] = want_rank - run;
            break;
          }
          run += c;
        }
      }
      __syncthreads();
    }
    if (shift == 0) break;
    // next digit: one histogram per DISTINCT prefix (lo/hi neighbours usually share theirs)
    const int prev_shift = shift;
    bits = shift < kDigitBits ? shift : kDigitBits;
    shift -= bits;
    const unsigned mask = (1u << bits) - 1u;
    if (tid == 0) {
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
    for (int i = tid; i < kTargets * kBins; i += kThreads) s_dyn[i] = 0;
    __syncthreads();
    const U p0 = s_prefix[0], p1 = s_prefix[1], p2 = s_prefix[2], p3 = s_prefix[3];
    const bool u1 = s_hidx[1] == 1, u2 = s_hidx[2] == 2, u3 = s_hidx[3] == 3;
    for_each_cell([&](T v, bool in) {
      const U k = Key<T>::key(v);
      const U hi = k >> prev_shift;
      const bool valid = in && !is_nan(v);
      const bool m0 = valid && hi == p0, m1 = u1 && valid && hi == p1, m2 = u2 && valid && hi == p2,
                 m3 = u3 && valid && hi == p3;
      if (__ballot_sync(0xffffffffu, m0 | m1 | m2 | m3) == 0u) return;  // nothing of this warp in a followed bucket
      const unsigned b = (unsigned)(k >> shift) & mask;
      if (m0) atomicAdd(&s_hist[0][b], 1u);
      if (m1) atomicAdd(&s_hist[1][b], 1u);
      if (m2) atomicAdd(&s_hist[2][b], 1u);
      if (m3) atomicAdd(&s_hist[3][b], 1u);
    });
    __syncthreads();
  }

  if (tid == 0) {
    const T a0 = Key<T>::val(s_prefix[0]), b0 = Key<T>::val(s_prefix[1]);
    const T a1 = Key<T>::val(s_prefix[2]), b1 = Key<T>::val(s_prefix[3]);
    st.p_lo = (double)numpy_lerp<T>(a0, b0, gamma[0]);
    st.p_hi = (double)numpy_lerp<T>(a1, b1, gamma[1]);
    out[blockIdx.x] = st;
  }
}

}  // namespace

extern "C" int csg_region_stats_run(csg_ctx* ctx, const void* d_mats, int dtype, const csg_region* d_regions,
                                    int n_regions, const int32_t* d_index_pool, csg_region_stats* d_out) {
  if (!ctx) return CSG_ERR_ARG;
  if (n_regions <= 0) return CSG_OK;
  if (!d_mats || !d_regions || !d_index_pool || !d_out) return csg_fail(ctx, CSG_ERR_ARG, "NULL argument");
  if (dtype != CSG_F32 && dtype != CSG_F64) return csg_fail(ctx, CSG_ERR_ARG, "bad dtype %d", dtype);
  // two specialisations over the same table (a block whose region belongs to the other one
  // exits at once): percentile regions need the histogram machinery, the rest only reductions
  const size_t heavy_smem = (size_t)kWarps * kBins * sizeof(unsigned);
  if (dtype == CSG_F32) {
    auto heavy = region_stats_kernel<float, true>;
    cudaFuncSetAttribute(heavy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heavy_smem);
    heavy<<<n_regions, kThreads, heavy_smem, ctx->stream>>>((const float*)d_mats, d_regions, d_index_pool, d_out);
    CSG_LAUNCH_CHECK(ctx, "region_stats_kernel<heavy>");
    region_stats_kernel<float, false><<<n_regions, kThreads, 0, ctx->stream>>>((const float*)d_mats, d_regions, d_index_pool, d_out);
  } else {
    auto heavy = region_stats_kernel<double, true>;
    cudaFuncSetAttribute(heavy, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)heavy_smem);
    heavy<<<n_regions, kThreads, heavy_smem, ctx->stream>>>((const double*)d_mats, d_regions, d_index_pool, d_out);
    CSG_LAUNCH_CHECK(ctx, "region_stats_kernel<heavy>");
    region_stats_kernel<double, false><<<n_regions, kThreads, 0, ctx->stream>>>((const double*)d_mats, d_regions, d_index_pool, d_out);
  }
  CSG_LAUNCH_CHECK(ctx, "region_stats_kernel<light>");
  return CSG_OK;
}
