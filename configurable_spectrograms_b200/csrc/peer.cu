// Peer exchange -- the one cross-GPU step of the path (the histogram merge of the global
// extrema, CS/fast/extrema.py:259-300 spread over ranks) done with stores into the other
// ranks' HBM over NVLink / NVSwitch instead of library collectives.
//
// Every rank owns a mailbox in its own HBM:  flags[region][src_rank] (uint64 epochs) and
// data[region][src_rank][...].  An all-gather is ONE kernel: each rank copies its (tiny: a few
// hundred KB of bucket totals at most) payload into every peer's mailbox with 128-bit stores,
// fences at system scope, publishes the epoch in every peer's flag word with a release store,
// and its last block then spins (acquire loads) on the local flag words until every rank's
// epoch has arrived.  Consumers follow on the stream and read their own HBM.
// Payloads travel once, nothing leaves the stream, the host never blocks, and the latency
// of one exchange is a few microseconds.
//
// Regions are used round-robin.  A rank can run at most one exchange ahead of the slowest
// reader (it needs that reader's next flag to get past its own next wait), so two regions
// would do; four are allocated.
//
// The mailboxes of the other ranks are mapped either through CUDA IPC handles (one process
// per GPU, the production layout) or handed over as raw device pointers (several ranks
// simulated inside one process: the single-GPU tests).
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

#define CSG_PEER_MAX 16

struct csg_peer {
  int rank, n_ranks, n_regions;
  size_t slot_bytes;    // capacity of one rank's payload in one region
  size_t flags_bytes;   // offset of the data area inside the mailbox
  size_t mailbox_bytes;
  unsigned char* local;
  unsigned char* peers[CSG_PEER_MAX];
  bool opened[CSG_PEER_MAX];  // mapped through cudaIpcOpenMemHandle (to be closed)
  unsigned long long epoch;
  unsigned* d_counter;
  int* d_error;  // 0, or 1 + the rank whose epoch did not arrive in time
  long long* d_wait_log;  // cycles the last 64 exchanges spent waiting for the other ranks' epochs
  unsigned long long* d_trace;  // globaltimer ns of the last 64 exchanges: [epoch % 64][kernel start, published, wait over, -]
  bool connected;
};

namespace {

struct PeerTable {
  unsigned char* p[CSG_PEER_MAX];
};

__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// One kernel per all-gather.  grid (chunks, n_ranks): blockIdx.y = destination rank.  Every block stores its
// share of the payload into that rank's mailbox and fences; the block that finishes last publishes the
// epoch in every peer's flag word and then -- still inside the kernel, so the consumer simply follows on
// the stream -- waits until every rank's epoch has arrived in the local flags (lane r watches rank r).
// wait_log[epoch % 64] receives the cycles spent in that wait (rank skew + link latency).
__global__ void __launch_bounds__(256)
    peer_exchange_kernel(PeerTable tbl, const uint4* __restrict__ src, size_t n16, size_t data_off, size_t flag_off,
                         unsigned long long epoch, unsigned* __restrict__ counter, int n_ranks,
                         const unsigned long long* __restrict__ local_flags, int* __restrict__ error, long long timeout_cycles,
                         long long* __restrict__ wait_log, unsigned long long* __restrict__ trace) {
  if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) trace[(epoch & 63ull) * 4 + 0] = global_ns();
  uint4* dst = reinterpret_cast<uint4*>(tbl.p[blockIdx.y] + data_off);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = src[i];
  __threadfence_system();
  __syncthreads();
  __shared__ bool s_last;
  if (threadIdx.x == 0) {
    const unsigned done = atomicAdd(counter, 1u) + 1u;
    s_last = done == gridDim.x * gridDim.y;
    if (s_last) {  // every block's stores are fenced
      *counter = 0;
      __threadfence_system();
    }
  }
  __syncthreads();
  if (!s_last || threadIdx.x >= 32) return;
  // publish: lane r tells rank r (one NVLink round trip for the whole warp instead of one per rank)
  if ((int)threadIdx.x < n_ranks)
    st_release_sys(reinterpret_cast<unsigned long long*>(tbl.p[threadIdx.x] + flag_off), epoch);
  __syncwarp();
  if (threadIdx.x == 0) trace[(epoch & 63ull) * 4 + 1] = global_ns();
  const long long t0 = clock64();
  if ((int)threadIdx.x < n_ranks) {
    while (ld_acquire_sys(local_flags + threadIdx.x) < epoch) {
      if (clock64() - t0 > timeout_cycles) {  // never hang the GPU: flag the step, let it finish
        atomicCAS(error, 0, 1 + (int)threadIdx.x);
        break;
      }
      __nanosleep(32);
    }
  }
  __syncwarp();
  if (threadIdx.x == 0) {
    wait_log[epoch & 63ull] = clock64() - t0;
    trace[(epoch & 63ull) * 4 + 2] = global_ns();
  }
}

size_t round_up(size_t n, size_t a) { return (n + a - 1) / a * a; }

}  // namespace

extern "C" {

int csg_peer_create(csg_ctx* ctx, int rank, int n_ranks, size_t slot_bytes, csg_peer** out, void* ipc_handle_64) {
  if (!ctx || !out) return CSG_ERR_ARG;
  if (n_ranks < 1 || n_ranks > CSG_PEER_MAX || rank < 0 || rank >= n_ranks)
    return csg_fail(ctx, CSG_ERR_ARG, "peer group: rank %d of %d (at most %d ranks)", rank, n_ranks, CSG_PEER_MAX);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
  csg_peer* p = (csg_peer*)calloc(1, sizeof(csg_peer));
  if (!p) return csg_fail(ctx, CSG_ERR_ARG, "out of host memory");
  p->rank = rank, p->n_ranks = n_ranks, p->n_regions = 4;
  p->slot_bytes = round_up(slot_bytes ? slot_bytes : 16, 256);
  p->flags_bytes = round_up((size_t)p->n_regions * CSG_PEER_MAX * sizeof(unsigned long long), 4096);
  p->mailbox_bytes = p->flags_bytes + (size_t)p->n_regions * n_ranks * p->slot_bytes;
  cudaError_t e = cudaMalloc((void**)&p->local, p->mailbox_bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**)&p->d_counter, 4096);
  if (e == cudaSuccess) e = cudaMemset(p->local, 0, p->mailbox_bytes);  // epochs start at 0
  if (e == cudaSuccess) e = cudaMemset(p->d_counter, 0, 4096);
  if (e != cudaSuccess) {
    if (p->local) cudaFree(p->local);
    if (p->d_counter) cudaFree(p->d_counter);
    free(p);
    return csg_fail(ctx, CSG_ERR_CUDA, "peer mailbox allocation failed: %s", cudaGetErrorString(e));
  }
  p->d_error = reinterpret_cast<int*>(p->d_counter) + 32;
  p->d_wait_log = reinterpret_cast<long long*>(p->d_counter) + 32;  // bytes 256 .. 767
  p->d_trace = reinterpret_cast<unsigned long long*>(p->d_counter) + 128;  // bytes 1024 .. 3071
  p->peers[rank] = p->local;
  if (ipc_handle_64) {
    cudaIpcMemHandle_t h;
    e = cudaIpcGetMemHandle(&h, p->local);
    if (e != cudaSuccess) {
      cudaGetLastError();
      memset(ipc_handle_64, 0, 64);  // no IPC on this platform: raw-pointer groups still work
    } else {
      memcpy(ipc_handle_64, &h, 64);
    }
  }
  cudaDeviceSynchronize();  // the zeroed flags must be in place before any peer writes
  *out = p;
  return CSG_OK;
}

void* csg_peer_mailbox(csg_peer* p) { return p ? (void*)p->local : nullptr; }

int csg_peer_connect_ipc(csg_ctx* ctx, csg_peer* p, const void* all_handles) {
  if (!ctx || !p || !all_handles) return CSG_ERR_ARG;
  for (int r = 0; r < p->n_ranks; ++r) {
    if (r == p->rank) continue;
    cudaIpcMemHandle_t h;
    memcpy(&h, (const unsigned char*)all_handles + (size_t)r * 64, 64);
    void* ptr = nullptr;
    CSG_CUDA(ctx, cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess));
    p->peers[r] = (unsigned char*)ptr;
    p->opened[r] = true;
  }
  p->connected = true;
  return CSG_OK;
}

int csg_peer_connect_ptrs(csg_ctx* ctx, csg_peer* p, void* const* mailboxes) {
  if (!ctx || !p || !mailboxes) return CSG_ERR_ARG;
  for (int r = 0; r < p->n_ranks; ++r)
    if (r != p->rank) p->peers[r] = (unsigned char*)mailboxes[r];
  p->connected = true;
  return CSG_OK;
}

int csg_peer_allgather(csg_ctx* ctx, csg_peer* p, const void* d_src, size_t nbytes, void** d_gathered) {
  if (!ctx || !p || !d_src || !d_gathered) return CSG_ERR_ARG;
  if (!p->connected) return csg_fail(ctx, CSG_ERR_ARG, "peer group is not connected");
  if (nbytes == 0 || nbytes % 16 != 0 || ((uintptr_t)d_src & 15) != 0)
    return csg_fail(ctx, CSG_ERR_ARG, "peer all-gather payload must be a non-empty multiple of 16 bytes, 16-byte aligned");
  if (nbytes > p->slot_bytes)
    return csg_fail(ctx, CSG_ERR_ARG, "peer all-gather payload %zu exceeds the mailbox slot (%zu bytes)", nbytes, p->slot_bytes);
  const unsigned long long epoch = ++p->epoch;
  const int region = (int)(epoch % (unsigned long long)p->n_regions);
  const size_t region_off = p->flags_bytes + (size_t)region * p->n_ranks * p->slot_bytes;
  const size_t data_off = region_off + (size_t)p->rank * nbytes;  // packed: rank stride = nbytes
  const size_t flag_off = ((size_t)region * CSG_PEER_MAX + p->rank) * sizeof(unsigned long long);
  PeerTable tbl;
  for (int r = 0; r < CSG_PEER_MAX; ++r) tbl.p[r] = r < p->n_ranks ? p->peers[r] : nullptr;
  const size_t n16 = nbytes / 16;
  int chunks = (int)((n16 + 1023) / 1024);
  if (chunks < 1) chunks = 1;
  if (chunks > 32) chunks = 32;
  const unsigned long long* flags =
      reinterpret_cast<const unsigned long long*>(p->local + (size_t)region * CSG_PEER_MAX * sizeof(unsigned long long));
  // a rank may arrive seconds late (first-step planning, lazy module loads): wait long, never forever
  static long long timeout_cycles = 0;
  if (timeout_cycles == 0) {
    const char* env = getenv("CSG_PEER_TIMEOUT_S");
    double seconds = env ? atof(env) : 20.0;
    if (!(seconds > 0.0)) seconds = 20.0;
    timeout_cycles = (long long)(seconds * 2.0e9);
  }
  peer_exchange_kernel<<<dim3(chunks, p->n_ranks), 256, 0, ctx->stream>>>(tbl, (const uint4*)d_src, n16, data_off, flag_off,
                                                                          epoch, p->d_counter, p->n_ranks, flags, p->d_error,
                                                                          timeout_cycles, p->d_wait_log, p->d_trace);
  CSG_LAUNCH_CHECK(ctx, "peer_exchange_kernel");
  *d_gathered = p->local + region_off;
  return CSG_OK;
}

void* csg_peer_error_word(csg_peer* p) { return p ? (void*)p->d_error : nullptr; }

int csg_peer_wait_stats(csg_ctx* ctx, csg_peer* p, int last_n, double* mean_us, double* max_us) {
  if (!ctx || !p || !mean_us || !max_us) return CSG_ERR_ARG;
  *mean_us = *max_us = 0.0;
  long long log[64];
  CSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CSG_CUDA(ctx, cudaMemcpy(log, p->d_wait_log, sizeof log, cudaMemcpyDeviceToHost));
  int khz = 0;
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, ctx->device);  // clock64 ticks at the SM clock
  const double us_per_cycle = khz > 0 ? 1000.0 / (double)khz : 1.0 / 1965.0;
  if (last_n > 64) last_n = 64;
  if ((unsigned long long)last_n > p->epoch) last_n = (int)p->epoch;
  if (last_n <= 0) return CSG_OK;
  double sum = 0.0, top = 0.0;
  for (int k = 0; k < last_n; ++k) {
    const double us = (double)log[(p->epoch - (unsigned long long)k) & 63ull] * us_per_cycle;
    sum += us;
    if (us > top) top = us;
  }
  *mean_us = sum / last_n, *max_us = top;
  return CSG_OK;
}

int csg_peer_trace(csg_ctx* ctx, csg_peer* p, int last_n, uint64_t* ns3, int* n_written) {
  if (!ctx || !p || !ns3 || !n_written) return CSG_ERR_ARG;
  *n_written = 0;
  unsigned long long log[256];
  CSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  CSG_CUDA(ctx, cudaMemcpy(log, p->d_trace, sizeof log, cudaMemcpyDeviceToHost));
  if (last_n > 64) last_n = 64;
  if ((unsigned long long)last_n > p->epoch) last_n = (int)p->epoch;
  for (int k = 0; k < last_n; ++k) {  // oldest first
    const unsigned long long e = p->epoch - (unsigned long long)(last_n - 1 - k);
    for (int j = 0; j < 3; ++j) ns3[3 * k + j] = log[(e & 63ull) * 4 + j];
  }
  *n_written = last_n > 0 ? last_n : 0;
  return CSG_OK;
}

int csg_peer_clear_error(csg_ctx* ctx, csg_peer* p) {
  if (!ctx || !p) return CSG_ERR_ARG;
  return csg_fill(ctx, p->d_error, 0, sizeof(int));
}

int csg_peer_disconnect(csg_ctx* ctx, csg_peer* p) {
  if (!p) return CSG_OK;
  if (ctx) cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < p->n_ranks; ++r)
    if (p->opened[r] && p->peers[r]) {
      cudaIpcCloseMemHandle(p->peers[r]);
      p->peers[r] = nullptr;
      p->opened[r] = false;
    }
  p->connected = false;
  return CSG_OK;
}

int csg_peer_destroy(csg_ctx* ctx, csg_peer* p) {
  if (!p) return CSG_OK;
  if (ctx) cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < p->n_ranks; ++r)
    if (p->opened[r] && p->peers[r]) cudaIpcCloseMemHandle(p->peers[r]);
  cudaFree(p->local);
  cudaFree(p->d_counter);
  free(p);
  return CSG_OK;
}

}  // extern "C"
