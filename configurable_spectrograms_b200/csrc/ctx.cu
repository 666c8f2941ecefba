// Context, memory, timing: the plumbing half of the C ABI (include/csgpu.h).
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <unordered_map>
#include <utility>
#include <vector>

#include "common.cuh"

char g_csg_err[512] = "";

int csg_fail(csg_ctx* ctx, int status, const char* fmt, ...) {
  char* dst = ctx ? ctx->err : g_csg_err;
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(dst, 512, fmt, ap);
  va_end(ap);
  if (ctx) memcpy(g_csg_err, ctx->err, 512);
  return status;
}

namespace {
__global__ void __launch_bounds__(256) fill16_kernel(uint4* __restrict__ dst, size_t n16, unsigned word) {
  const uint4 v = make_uint4(word, word, word, word);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) dst[i] = v;
}
__global__ void __launch_bounds__(256) fill1_kernel(unsigned char* __restrict__ dst, size_t n, unsigned char b) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] = b;
}

// ------------------------------------------------------------------ device block cache
// cudaMalloc / cudaFree are device-wide synchronisation points that cost milliseconds for the table- and
// raster-sized buffers a directory run allocates per chunk (measured: 17 ms per cudaFree, 4.3 s of a 9 s
// run).  csg_dev_free therefore parks a block instead of returning it to the driver, and csg_dev_alloc
// hands a parked block of the same size class to the next caller.  What cudaFree guaranteed implicitly --
// nothing is still using the memory -- is kept explicit: a parked block carries one event per stream of
// every context alive on the device, recorded at release, and is only handed out once they are all done.
struct Block {
  void* ptr = nullptr;
  size_t bytes = 0;
  std::vector<cudaEvent_t> fences;
};
struct BlockCache {
  std::mutex lock;
  std::multimap<size_t, Block> idle;        // size class -> parked blocks
  std::unordered_map<void*, size_t> live;   // blocks handed out -> their size class
  std::vector<csg_ctx*> contexts;           // contexts alive on this device
  std::vector<cudaEvent_t> spare_events;
  size_t idle_bytes = 0;
  size_t limit_bytes = (size_t)32 << 30;    // CSG_POOL_MAX_MB
  bool enabled = true;                      // CSG_POOL=0: plain cudaMalloc / cudaFree
  BlockCache() {
    if (const char* v = getenv("CSG_POOL")) enabled = atoi(v) != 0;
    if (const char* v = getenv("CSG_POOL_MAX_MB")) limit_bytes = (size_t)strtoull(v, nullptr, 10) << 20;
  }
};
BlockCache& cache_of(int device) {
  static BlockCache caches[64];
  return caches[(device >= 0 && device < 64) ? device : 0];
}
// Four size classes per power of two (at most 25 % over), 512 bytes at least: per-chunk tables whose sizes
// wobble with the number of figures still find their predecessor's block.
size_t size_class(size_t bytes) {
  if (bytes <= 512) return 512;
  size_t top = (size_t)1 << 9;
  while ((top << 1) < bytes && (top << 1) != 0) top <<= 1;  // top < bytes <= 2 * top
  const size_t step = top >> 2;
  return top + (bytes - top + step - 1) / step * step;
}
size_t trim(BlockCache& cache) {
  std::multimap<size_t, Block> idle;
  {
    std::lock_guard<std::mutex> hold(cache.lock);
    idle.swap(cache.idle);
    cache.idle_bytes = 0;
  }
  size_t n = 0;
  std::vector<cudaEvent_t> events;
  for (auto& kv : idle) {
    cudaFree(kv.second.ptr);
    n += kv.second.bytes;
    events.insert(events.end(), kv.second.fences.begin(), kv.second.fences.end());
  }
  std::lock_guard<std::mutex> hold(cache.lock);
  cache.spare_events.insert(cache.spare_events.end(), events.begin(), events.end());
  return n;
}
void enroll(csg_ctx* ctx, bool on) {
  BlockCache& cache = cache_of(ctx->device);
  std::lock_guard<std::mutex> hold(cache.lock);
  for (size_t i = 0; i < cache.contexts.size(); ++i)
    if (cache.contexts[i] == ctx) {
      cache.contexts.erase(cache.contexts.begin() + i);
      break;
    }
  if (on) cache.contexts.push_back(ctx);
}
}  // namespace

int csg_fill(csg_ctx* ctx, void* d_dst, int byte_value, size_t bytes) {
  if (!ctx) return CSG_ERR_ARG;
  if (bytes == 0) return CSG_OK;
  if (!d_dst) return csg_fail(ctx, CSG_ERR_ARG, "csg_fill: NULL destination");
  const unsigned b = (unsigned)byte_value & 0xffu;
  unsigned char* p = (unsigned char*)d_dst;
  const size_t head = (16 - ((uintptr_t)p & 15)) & 15;
  const size_t h = head < bytes ? head : bytes;
  if (h) {
    fill1_kernel<<<1, 32, 0, ctx->stream>>>(p, h, (unsigned char)b);
    CSG_LAUNCH_CHECK(ctx, "fill1_kernel");
  }
  const size_t n16 = (bytes - h) / 16;
  if (n16) {
    size_t blocks = (n16 + 255) / 256;
    const size_t cap = (size_t)ctx->sm_count * 16;
    if (blocks > cap) blocks = cap;
    fill16_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>((uint4*)(p + h), n16, b * 0x01010101u);
    CSG_LAUNCH_CHECK(ctx, "fill16_kernel");
  }
  const size_t tail = bytes - h - n16 * 16;
  if (tail) {
    fill1_kernel<<<1, 32, 0, ctx->stream>>>(p + h + n16 * 16, tail, (unsigned char)b);
    CSG_LAUNCH_CHECK(ctx, "fill1_kernel");
  }
  return CSG_OK;
}

extern "C" {

int csg_abi_version(void) { return CSG_ABI_VERSION; }

int csg_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

csg_ctx* csg_create(int device, void* external_stream) {
  int n = csg_device_count();
  if (n <= 0) {
    csg_fail(nullptr, CSG_ERR_NODEV, "no CUDA device available: libcsgpu has no CPU fallback");
    return nullptr;
  }
  if (device < 0 || device >= n) {
    csg_fail(nullptr, CSG_ERR_ARG, "device %d out of range (0..%d)", device, n - 1);
    return nullptr;
  }
  if (cudaSetDevice(device) != cudaSuccess) {
    csg_fail(nullptr, CSG_ERR_CUDA, "cudaSetDevice(%d) failed: %s", device,
             cudaGetErrorString(cudaGetLastError()));
    return nullptr;
  }
  csg_ctx* ctx = (csg_ctx*)calloc(1, sizeof(csg_ctx));
  if (!ctx) return nullptr;
  ctx->device = device;
  if (external_stream) {
    ctx->stream = (cudaStream_t)external_stream;
    ctx->own_stream = false;
  } else {
    if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) {
      csg_fail(nullptr, CSG_ERR_CUDA, "cudaStreamCreate failed: %s",
               cudaGetErrorString(cudaGetLastError()));
      free(ctx);
      return nullptr;
    }
    ctx->own_stream = true;
  }
  cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device);
  for (int i = 0; i < 32; ++i) {
    cudaEventCreate(&ctx->ev_start[i]);
    cudaEventCreate(&ctx->ev_stop[i]);
    cudaEventCreateWithFlags(&ctx->ev_user[i], cudaEventDisableTiming);
  }
  ctx->side = nullptr;  // created on first use (csg_d2h_side)
  cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->ev_side, cudaEventDisableTiming);
  enroll(ctx, true);
  return ctx;
}

csg_ctx* csg_create_with_priority(int device, int level) {
  csg_ctx* ctx = csg_create(device, nullptr);
  if (!ctx || level <= 0) return ctx;
  // swap the default-priority stream for a more urgent one: the block scheduler serves it first, so its
  // kernels slip in between the blocks of a wide kernel on a less urgent stream
  int least = 0, greatest = 0;
  cudaStream_t s = nullptr;
  if (cudaDeviceGetStreamPriorityRange(&least, &greatest) == cudaSuccess) {
    int prio = least - level;  // numerically lower = more urgent
    if (prio < greatest) prio = greatest;
    if (cudaStreamCreateWithPriority(&s, cudaStreamNonBlocking, prio) == cudaSuccess) {
      std::lock_guard<std::mutex> hold(cache_of(device).lock);  // csg_dev_free records fences on enrolled streams
      cudaStreamDestroy(ctx->stream);
      ctx->stream = s;
      return ctx;
    }
  }
  cudaGetLastError();
  return ctx;
}

csg_ctx* csg_create_side(int device, int high_priority) {
  return csg_create_with_priority(device, high_priority ? 1 << 20 : 0);  // the most urgent level there is
}

int csg_wait_for(csg_ctx* waiter, csg_ctx* signal) {
  if (!waiter || !signal) return CSG_ERR_ARG;
  CSG_CUDA(waiter, cudaEventRecord(signal->ev_fork, signal->stream));
  CSG_CUDA(waiter, cudaStreamWaitEvent(waiter->stream, signal->ev_fork, 0));
  return CSG_OK;
}

void csg_destroy(csg_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  if (ctx->side) cudaStreamSynchronize(ctx->side);
  enroll(ctx, false);  // (synchronised above: blocks parked from now on need no fence on these streams)
  for (int i = 0; i < 32; ++i) {
    cudaEventDestroy(ctx->ev_start[i]);
    cudaEventDestroy(ctx->ev_stop[i]);
    cudaEventDestroy(ctx->ev_user[i]);
  }
  if (ctx->side) {
    cudaStreamSynchronize(ctx->side);
    cudaStreamDestroy(ctx->side);
  }
  cudaEventDestroy(ctx->ev_fork);
  cudaEventDestroy(ctx->ev_side);
  if (ctx->scratch) cudaFree(ctx->scratch);
  if (ctx->own_stream) cudaStreamDestroy(ctx->stream);
  free(ctx);
}

const char* csg_last_error(csg_ctx* ctx) { return ctx ? ctx->err : g_csg_err; }

void* csg_stream_handle(csg_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

int csg_sync(csg_ctx* ctx) {
  CSG_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return CSG_OK;
}

int csg_device_info(csg_ctx* ctx, char* name, int name_len, int* sm_count, size_t* total_mem) {
  cudaDeviceProp prop;
  CSG_CUDA(ctx, cudaGetDeviceProperties(&prop, ctx->device));
  if (name && name_len > 0) {
    strncpy(name, prop.name, name_len - 1);
    name[name_len - 1] = 0;
  }
  if (sm_count) *sm_count = prop.multiProcessorCount;
  if (total_mem) *total_mem = prop.totalGlobalMem;
  return CSG_OK;
}

int csg_dev_alloc(csg_ctx* ctx, size_t bytes, void** d_ptr) {
  if (!ctx || !d_ptr) return CSG_ERR_ARG;
  CSG_CUDA(ctx, cudaSetDevice(ctx->device));
  BlockCache& cache = cache_of(ctx->device);
  const size_t want = cache.enabled ? size_class(bytes) : (bytes ? bytes : 1);
  if (cache.enabled) {
    Block hit;
    bool found = false;
    {
      // a parked block whose fences have all passed; one that is still fenced (another context's kernel
      // is running: the K4 encoder beside the planner) is left alone rather than waited for
      std::lock_guard<std::mutex> hold(cache.lock);
      auto range = cache.idle.equal_range(want);
      for (auto it = range.first; it != range.second && !found; ++it) {
        bool ready = true;
        for (cudaEvent_t ev : it->second.fences)
          if (cudaEventQuery(ev) != cudaSuccess) {
            cudaGetLastError();
            ready = false;
            break;
          }
        if (!ready) continue;
        hit = std::move(it->second);
        cache.idle.erase(it);
        cache.idle_bytes -= want;
        cache.live[hit.ptr] = want;
        cache.spare_events.insert(cache.spare_events.end(), hit.fences.begin(), hit.fences.end());
        found = true;
        break;
      }
    }
    if (found) {
      *d_ptr = hit.ptr;
      return CSG_OK;
    }
  }
  cudaError_t e = cudaMalloc(d_ptr, want);
  if (e != cudaSuccess && cache.enabled) {  // give the driver back what is parked here, then once more
    cudaGetLastError();
    trim(cache);  // (cudaFree synchronises: whatever the fences guarded is over afterwards)
    e = cudaMalloc(d_ptr, want);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    return csg_fail(ctx, CSG_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
  }
  if (cache.enabled) {
    std::lock_guard<std::mutex> hold(cache.lock);
    cache.live[*d_ptr] = want;
  }
  return CSG_OK;
}
int csg_dev_free(csg_ctx* ctx, void* d_ptr) {
  if (!ctx) return CSG_ERR_ARG;
  if (!d_ptr) return CSG_OK;
  CSG_CUDA(ctx, cudaSetDevice(ctx->device));
  BlockCache& cache = cache_of(ctx->device);
  Block block;
  bool park = false;
  if (cache.enabled) {
    std::lock_guard<std::mutex> hold(cache.lock);
    auto it = cache.live.find(d_ptr);
    if (it != cache.live.end()) {
      block.ptr = d_ptr;
      block.bytes = it->second;
      cache.live.erase(it);
      park = cache.idle_bytes + block.bytes <= cache.limit_bytes;
      if (park) {
        for (csg_ctx* c : cache.contexts) {  // one fence per stream that may still be touching the block
          cudaStream_t streams[2] = {c->stream, c->side};
          for (int k = 0; k < (c->side ? 2 : 1); ++k) {
            cudaStream_t st = streams[k];
            cudaEvent_t ev;
            if (!cache.spare_events.empty()) {
              ev = cache.spare_events.back();
              cache.spare_events.pop_back();
            } else if (cudaEventCreateWithFlags(&ev, cudaEventDisableTiming) != cudaSuccess) {
              cudaGetLastError();
              park = false;
              break;
            }
            if (cudaEventRecord(ev, st) != cudaSuccess) {
              cudaGetLastError();
              cache.spare_events.push_back(ev);
              park = false;
              break;
            }
            block.fences.push_back(ev);
          }
          if (!park) break;
        }
        if (park) {
          const size_t n = block.bytes;
          cache.idle.emplace(n, std::move(block));
          cache.idle_bytes += n;
          return CSG_OK;
        }
        cache.spare_events.insert(cache.spare_events.end(), block.fences.begin(), block.fences.end());
      }
    }
  }
  CSG_CUDA(ctx, cudaFree(d_ptr));  // (synchronises the device: nothing can still be using the block afterwards)
  return CSG_OK;
}
int csg_dev_trim(csg_ctx* ctx, size_t* released_bytes) {
  if (!ctx) return CSG_ERR_ARG;
  CSG_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t n = trim(cache_of(ctx->device));
  if (released_bytes) *released_bytes = n;
  return CSG_OK;
}
int csg_dev_cached(csg_ctx* ctx, size_t* idle_bytes, size_t* live_bytes) {
  if (!ctx) return CSG_ERR_ARG;
  BlockCache& cache = cache_of(ctx->device);
  std::lock_guard<std::mutex> hold(cache.lock);
  if (idle_bytes) *idle_bytes = cache.idle_bytes;
  if (live_bytes) {
    size_t n = 0;
    for (auto& kv : cache.live) n += kv.second;
    *live_bytes = n;
  }
  return CSG_OK;
}
int csg_host_alloc(csg_ctx* ctx, size_t bytes, void** h_ptr) {
  cudaError_t e = cudaHostAlloc(h_ptr, bytes ? bytes : 1, cudaHostAllocDefault);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return csg_fail(ctx, CSG_ERR_NOMEM, "cudaHostAlloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
  }
  return CSG_OK;
}
int csg_host_free(csg_ctx* ctx, void* h_ptr) {
  CSG_CUDA(ctx, cudaFreeHost(h_ptr));
  return CSG_OK;
}
int csg_host_register(csg_ctx* ctx, void* h_ptr, size_t bytes) {
  CSG_CUDA(ctx, cudaHostRegister(h_ptr, bytes, cudaHostRegisterDefault));
  return CSG_OK;
}
int csg_host_unregister(csg_ctx* ctx, void* h_ptr) {
  CSG_CUDA(ctx, cudaHostUnregister(h_ptr));
  return CSG_OK;
}
int csg_h2d(csg_ctx* ctx, void* d_dst, const void* h_src, size_t bytes) {
  if (bytes == 0) return CSG_OK;
  CSG_CUDA(ctx, cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  return CSG_OK;
}
int csg_d2h(csg_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
  if (bytes == 0) return CSG_OK;
  CSG_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->stream));
  return CSG_OK;
}
int csg_d2d(csg_ctx* ctx, void* d_dst, const void* d_src, size_t bytes) {
  if (bytes == 0) return CSG_OK;
  CSG_CUDA(ctx, cudaMemcpyAsync(d_dst, d_src, bytes, cudaMemcpyDeviceToDevice, ctx->stream));
  return CSG_OK;
}
int csg_d2h_side(csg_ctx* ctx, void* h_dst, const void* d_src, size_t bytes) {
  if (!ctx) return CSG_ERR_ARG;
  if (bytes == 0) return CSG_OK;
  if (!ctx->side) CSG_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
  CSG_CUDA(ctx, cudaEventRecord(ctx->ev_fork, ctx->stream));
  CSG_CUDA(ctx, cudaStreamWaitEvent(ctx->side, ctx->ev_fork, 0));
  CSG_CUDA(ctx, cudaMemcpyAsync(h_dst, d_src, bytes, cudaMemcpyDeviceToHost, ctx->side));
  CSG_CUDA(ctx, cudaEventRecord(ctx->ev_side, ctx->side));
  return CSG_OK;
}
int csg_side_join(csg_ctx* ctx) {
  if (!ctx) return CSG_ERR_ARG;
  if (!ctx->side) return CSG_OK;  // nothing was ever copied out
  CSG_CUDA(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_side, 0));
  return CSG_OK;
}
int csg_side_sync(csg_ctx* ctx) {
  if (!ctx) return CSG_ERR_ARG;
  if (!ctx->side) return CSG_OK;
  CSG_CUDA(ctx, cudaStreamSynchronize(ctx->side));
  return CSG_OK;
}
int csg_memset(csg_ctx* ctx, void* d_dst, int byte_value, size_t bytes) { return csg_fill(ctx, d_dst, byte_value, bytes); }

int csg_timer_start(csg_ctx* ctx, int slot) {
  if (slot < 0 || slot >= 32) return csg_fail(ctx, CSG_ERR_ARG, "timer slot %d out of range", slot);
  CSG_CUDA(ctx, cudaEventRecord(ctx->ev_start[slot], ctx->stream));
  return CSG_OK;
}
int csg_timer_stop(csg_ctx* ctx, int slot) {
  if (slot < 0 || slot >= 32) return csg_fail(ctx, CSG_ERR_ARG, "timer slot %d out of range", slot);
  CSG_CUDA(ctx, cudaEventRecord(ctx->ev_stop[slot], ctx->stream));
  return CSG_OK;
}
int csg_timer_ms(csg_ctx* ctx, int slot, float* ms) {
  if (slot < 0 || slot >= 32) return csg_fail(ctx, CSG_ERR_ARG, "timer slot %d out of range", slot);
  CSG_CUDA(ctx, cudaEventSynchronize(ctx->ev_stop[slot]));
  CSG_CUDA(ctx, cudaEventElapsedTime(ms, ctx->ev_start[slot], ctx->ev_stop[slot]));
  return CSG_OK;
}
int64_t csg_launch_count(csg_ctx* ctx) { return ctx ? ctx->launches : 0; }

int csg_event_record(csg_ctx* ctx, int slot) {
  if (slot < 0 || slot >= 32) return csg_fail(ctx, CSG_ERR_ARG, "event slot %d out of range", slot);
  CSG_CUDA(ctx, cudaEventRecord(ctx->ev_user[slot], ctx->stream));
  return CSG_OK;
}
int csg_event_sync(csg_ctx* ctx, int slot) {
  if (slot < 0 || slot >= 32) return csg_fail(ctx, CSG_ERR_ARG, "event slot %d out of range", slot);
  CSG_CUDA(ctx, cudaEventSynchronize(ctx->ev_user[slot]));
  return CSG_OK;
}

}  // extern "C"
