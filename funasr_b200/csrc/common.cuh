// Shared helpers for the funasr_b200 sm_100a kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <atomic>
#include "../../include/funasr_b200.h"

namespace fa {

extern std::atomic<unsigned long long> g_launch_count;

inline void count_launch(unsigned n = 1) { g_launch_count.fetch_add(n, std::memory_order_relaxed); }

// Returns FA_ERR_CUDA from the enclosing function when the launch that just happened failed.
#define FA_CHECK_LAUNCH()                                        \
  do {                                                           \
    ::fa::count_launch();                                        \
    cudaError_t _e = cudaGetLastError();                         \
    if (_e != cudaSuccess) return FA_ERR_CUDA;                   \
  } while (0)

#define FA_CUDA_OK(expr)                                         \
  do {                                                           \
    cudaError_t _e = (expr);                                     \
    if (_e != cudaSuccess) return FA_ERR_CUDA;                   \
  } while (0)

#define FA_RETURN_IF_ERR(expr)                                   \
  do {                                                           \
    int _s = (expr);                                             \
    if (_s != FA_OK) return _s;                                  \
  } while (0)

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Programmatic dependent launch (PDL): every hot kernel is launched with programmaticStreamSerialization, runs its
// global-memory-free prologue (barrier init, TMEM alloc, descriptor prefetch) while the previous kernel in the stream drains,
// then pdl_wait() — which returns once the predecessor grid has completed and its writes are visible — before the first
// global access, and immediately lets its own successor start its prologue (pdl_trigger).  FA_PDL=0 disables the attribute.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

bool pdl_enabled();

template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, int cluster_x, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute attr[2];
  int n = 0;
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  if (cluster_x > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster_x; attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
    ++n;
  }
  cfg.attrs = attr; cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

// cudaFuncAttributeMaxDynamicSharedMemorySize and the SM count are PER-DEVICE: a process that drives two GPUs (two handles of
// the offline API, a server with one worker thread per device) must set / query them once per (kernel instantiation, device).
struct PerDeviceOnce { std::atomic<uint32_t> mask{0}; };
template <typename K>
inline int ensure_dyn_smem(K kern, size_t bytes, PerDeviceOnce& once) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return FA_ERR_CUDA;
  const uint32_t bit = 1u << (dev & 31);
  if (!(once.mask.load(std::memory_order_acquire) & bit)) {
    if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) return FA_ERR_CUDA;
    once.mask.fetch_or(bit, std::memory_order_release);
  }
  return FA_OK;
}
// SM count of the current device (cached per device)
inline int sm_count() {
  static std::atomic<int> cache[32];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return 148;
  int n = cache[dev & 31].load(std::memory_order_relaxed);
  if (n <= 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev & 31].store(n, std::memory_order_relaxed);
  }
  return n;
}

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over the caller's workspace.
struct Arena {
  char* base;
  size_t cap, off;
  Arena(void* p, size_t bytes) : base(static_cast<char*>(p)), cap(bytes), off(0) {}
  template <typename T>
  T* take(size_t n) {
    size_t o = align_up(off, 256);
    size_t need = n * sizeof(T);
    if (base == nullptr || o + need > cap) { off = cap + 1; return nullptr; }
    off = o + need;
    return reinterpret_cast<T*>(base + o);
  }
  bool ok() const { return off <= cap; }
};
// Same arithmetic as Arena::take, for the *_workspace_bytes queries.
struct ArenaSizer {
  size_t off = 0;
  void take(size_t bytes) { off = align_up(off, 256) + bytes; }
};

}  // namespace fa
