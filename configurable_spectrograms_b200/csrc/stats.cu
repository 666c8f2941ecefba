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

template <typename T>
__device__ __forceinline__ T region_cell(const T* __restrict__ mats, const csg_region& rg,
                                         const int32_t* __restrict__ pool, long long i) {
  const int r = (int)(i / rg.ne);
  const int c = (int)(i - (long long)r * rg.ne);
  const int row = rg.rows_off < 0 ? rg.t0 + r : __ldg(pool + rg.rows_off + r);
  const int col = __ldg(pool + rg.cols_off + c);
  return __ldg(mats + rg.mat_off + (long long)row * rg.ld + col);
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
  __shared__ long long s_ll[32];
  __shared__ double s_d[32];
  __shared__ unsigned s_u[32];
  __shared__ U s_prefix[kTargets];
  __shared__ long long s_rank[kTargets];

  const csg_region rg = regions[blockIdx.x];
  const long long n_cells = (long long)rg.nt * rg.ne;
  const int tid = threadIdx.x;

  for (int i = tid; i < kBins; i += kThreads) s_hist[0][i] = 0;
  __syncthreads();

  // ---- pass 0: classification + first digit
  long long n_valid = 0;
  unsigned n_nan = 0, n_ninf = 0, n_pinf = 0, n_pos = 0;
  double min_pos = CUDART_INF, fin_min = CUDART_INF, fin_max = -CUDART_INF;
  constexpr int kTopShift = Key<T>::BITS - kDigitBits;
  for (long long i = tid; i < n_cells; i += kThreads) {
    const T v = region_cell(mats, rg, pool, i);
    if (is_nan(v)) {
      ++n_nan;
      continue;
    }
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
    if (rg.want_pct) atomicAdd(&s_hist[0][(unsigned)(Key<T>::key(v) >> kTopShift)], 1u);
  }
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

  if (!rg.want_pct || n_valid == 0) {
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
  }
  __syncthreads();

  // ---- digit loop
  int shift = kTopShift;
  int bits = kDigitBits;
  bool first = true;
  while (true) {
    const int nb = 1 << bits;
    // locate each target's bucket in its histogram (level 0: all share histogram 0)
    for (int j = 0; j < kTargets; ++j) {
      const unsigned* h = s_hist[first ? 0 : j];
      const long long want = s_rank[j];
      // each thread owns nb/kThreads consecutive bins (nb >= kThreads is not required)
      const int per = (nb + kThreads - 1) / kThreads;
      const int b0 = tid * per;
      long long mine = 0;
      for (int b = b0; b < b0 + per && b < nb; ++b) mine += h[b];
      // exclusive scan of `mine` over threads
      const int lane = tid & 31, warp = tid >> 5;
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
      if (want >= excl && want < excl + mine) {
        long long run = excl;
        for (int b = b0; b < b0 + per && b < nb; ++b) {
          const long long c = h[b];
          if (want < run + c) {
            s_prefix[j] = (s_prefix[j] << bits) | (U)b;
            s_rank[j] = want - run;
            break;
          }
          run += c;
        }
      }
      __syncthreads();
    }
    if (shift == 0) break;
    // next digit
    const int prev_shift = shift;
    bits = shift < kDigitBits ? shift : kDigitBits;
    shift -= bits;
    const int nb2 = 1 << bits;
    for (int i = tid; i < kTargets * kBins; i += kThreads) (&s_hist[0][0])[i] = 0;
    __syncthreads();
    const U p0 = s_prefix[0], p1 = s_prefix[1], p2 = s_prefix[2], p3 = s_prefix[3];
    for (long long i = tid; i < n_cells; i += kThreads) {
      const T v = region_cell(mats, rg, pool, i);
      if (is_nan(v)) continue;
      const U k = Key<T>::key(v);
      const U hi = k >> prev_shift;
      const unsigned b = (unsigned)(k >> shift) & (unsigned)(nb2 - 1);
      if (hi == p0) atomicAdd(&s_hist[0][b], 1u);
      if (hi == p1) atomicAdd(&s_hist[1][b], 1u);
      if (hi == p2) atomicAdd(&s_hist[2][b], 1u);
      if (hi == p3) atomicAdd(&s_hist[3][b], 1u);
    }
    __syncthreads();
    first = false;
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
