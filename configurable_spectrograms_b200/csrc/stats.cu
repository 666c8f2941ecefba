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
// Traffic: the region's cells are re-read once per digit (3 passes f32, 6 passes f64) from L2.
#include "common.cuh"

namespace {

constexpr int kThreads = 512;
constexpr int kDigitBits = 11;
constexpr int kBins = 1 << kDigitBits;
constexpr int kTargets = 4;  // (lo, hi) neighbours of two percentiles

constexpr int kMaxCols = 1024;  // column lists up to this length are staged in shared memory

// one atomic per distinct bin per warp: spectrogram counts repeat heavily, so a plain
// shared-memory atomic per lane serialises on a handful of addresses
__device__ __forceinline__ void hist_add(unsigned* h, unsigned bin, bool valid) {
  const unsigned act = __ballot_sync(0xffffffffu, valid);
  if (valid) {
    const unsigned peers = __match_any_sync(act, bin);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&h[bin], (unsigned)__popc(peers));
  }
}

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

template <typename T>
__global__ void __launch_bounds__(kThreads)
    region_stats_kernel(const T* __restrict__ mats, const csg_region* __restrict__ regions,
                        const int32_t* __restrict__ pool, csg_region_stats* __restrict__ out) {
  typedef typename Key<T>::U U;
  __shared__ unsigned s_hist[kTargets][kBins];
  __shared__ int s_cols[kMaxCols];
  __shared__ long long s_ll[32];
  __shared__ double s_d[32];
  __shared__ unsigned s_u[32];
  __shared__ U s_prefix[kTargets];
  __shared__ long long s_rank[kTargets];
  __shared__ int s_hidx[kTargets];

  const csg_region rg = regions[blockIdx.x];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int kWarps = kThreads / 32;
  const bool cols_in_smem = rg.ne <= kMaxCols;

  for (int i = tid; i < kBins; i += kThreads) s_hist[0][i] = 0;
  if (cols_in_smem)
    for (int i = tid; i < rg.ne; i += kThreads) s_cols[i] = __ldg(pool + rg.cols_off + i);
  __syncthreads();

  // Walk the region warp-per-time-row: lanes stride over the energy columns, which are
  // (nearly) contiguous in the collapsed (T,E) matrix -> coalesced, division-free.
  auto for_each_cell = [&](auto&& fn) {
    for (int r = warp; r < rg.nt; r += kWarps) {
      const int row = rg.rows_off < 0 ? rg.t0 + r : __ldg(pool + rg.rows_off + r);
      const T* rp = mats + rg.mat_off + (long long)row * rg.ld;
      for (int c0 = 0; c0 < rg.ne; c0 += 32) {
        const int c = c0 + lane;
        const bool in = c < rg.ne;
        T v = T(0);
        if (in) v = __ldg(rp + (cols_in_smem ? s_cols[c] : __ldg(pool + rg.cols_off + c)));
        fn(v, in);
      }
    }
  };

  // ---- pass 0: classification + first digit
  long long n_valid = 0;
  unsigned n_nan = 0, n_ninf = 0, n_pinf = 0, n_pos = 0;
  double min_pos = CUDART_INF, fin_min = CUDART_INF, fin_max = -CUDART_INF;
  constexpr int kTopShift = Key<T>::BITS - kDigitBits;
  const bool want = rg.want_pct != 0;
  for_each_cell([&](T v, bool in) {
    const bool valid = in && !is_nan(v);
    if (in) {
      if (!valid) {
        ++n_nan;
      } else {
        ++n_valid;
        if (is_finite(v)) {
          const double dv = (double)v;
          fin_min = fmin(fin_min, dv);
          fin_max = fmax(fin_max, dv);
          if (v > T(0)) {
            ++n_pos;
            min_pos = fmin(min_pos, dv);
          }
        } else if (v > T(0)) {
          ++n_pinf;
        } else {
          ++n_ninf;
        }
      }
    }
    if (want) hist_add(s_hist[0], valid ? (unsigned)(Key<T>::key(v) >> kTopShift) : 0u, valid);
  });
  auto addll = [](long long a, long long b) { return a + b; };
  auto addu = [](unsigned a, unsigned b) { return a + b; };
  auto mind = [](double a, double b) { return fmin(a, b); };
  auto maxd = [](double a, double b) { return fmax(a, b); };
  n_valid = block_reduce(n_valid, addll, 0ll, s_ll);
  n_nan = block_reduce(n_nan, addu, 0u, s_u);
  n_ninf = block_reduce(n_ninf, addu, 0u, s_u);
  n_pinf = block_reduce(n_pinf, addu, 0u, s_u);
  n_pos = block_reduce(n_pos, addu, 0u, s_u);
  min_pos = block_reduce(min_pos, mind, (double)CUDART_INF, s_d);
  fin_min = block_reduce(fin_min, mind, (double)CUDART_INF, s_d);
  fin_max = block_reduce(fin_max, maxd, -(double)CUDART_INF, s_d);

  csg_region_stats st;
  st.p_lo = st.p_hi = CUDART_NAN;
  st.min_pos = min_pos;
  st.fin_min = fin_min;
  st.fin_max = fin_max;
  st.n_valid = n_valid;
  st.n_nan = (int)n_nan, st.n_neginf = (int)n_ninf, st.n_posinf = (int)n_pinf, st.n_pos = (int)n_pos;

  if (!want || n_valid == 0) {
    if (tid == 0) out[blockIdx.x] = st;
    return;
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
    const int nb2 = 1 << bits;
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
    for (int i = tid; i < kTargets * kBins; i += kThreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const U p0 = s_prefix[0], p1 = s_prefix[1], p2 = s_prefix[2], p3 = s_prefix[3];
    const bool u1 = s_hidx[1] == 1, u2 = s_hidx[2] == 2, u3 = s_hidx[3] == 3;
    for_each_cell([&](T v, bool in) {
      const bool valid = in && !is_nan(v);
      const U k = Key<T>::key(v);
      const U hi = k >> prev_shift;
      const unsigned b = (unsigned)(k >> shift) & (unsigned)(nb2 - 1);
      hist_add(s_hist[0], b, valid && hi == p0);
      if (u1) hist_add(s_hist[1], b, valid && hi == p1);
      if (u2) hist_add(s_hist[2], b, valid && hi == p2);
      if (u3) hist_add(s_hist[3], b, valid && hi == p3);
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
  if (dtype == CSG_F32)
    region_stats_kernel<float><<<n_regions, kThreads, 0, ctx->stream>>>((const float*)d_mats, d_regions, d_index_pool, d_out);
  else if (dtype == CSG_F64)
    region_stats_kernel<double><<<n_regions, kThreads, 0, ctx->stream>>>((const double*)d_mats, d_regions, d_index_pool, d_out);
  else
    return csg_fail(ctx, CSG_ERR_ARG, "bad dtype %d", dtype);
  CSG_LAUNCH_CHECK(ctx, "region_stats_kernel");
  return CSG_OK;
}
