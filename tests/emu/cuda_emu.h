// cuda_emu.h -- TEST-ONLY shim that lets g++ compile the qpwc kernels (qpwcnet_b200/csrc/*.cu,
// built with -DQPWC_EMU) and run them thread-by-thread on the CPU.
//
// Purpose: the authoring container has no GPU; this harness checks the kernels' index arithmetic,
// tiling, barrier placement and epilogues against the oracle before GPU minutes are spent.  It is
// NOT a backend: the product never builds, ships or loads it (qpwcnet_b200/_cabi.py loads only
// qpwcnet_b200/lib/libqpwc.so and raises if that is absent).  Only tests/test_emu_kernels.py uses it.
//
// Model: one OS thread per CUDA thread of a block (blocks run one after another), __syncthreads =
// a counting barrier, warp shuffles = a 32-thread exchange buffer, "device memory" = host memory.
#pragma once
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __launch_bounds__(...)
#define __align__(n) alignas(n)

struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};
struct float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
inline float2 make_float2(float x, float y) { return {x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return {x, y, z, w}; }
inline int4 make_int4(int x, int y, int z, int w) { return {x, y, z, w}; }

typedef void* cudaStream_t;
typedef int cudaError_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToDevice = 1, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3 };
enum { cudaStreamNonBlocking = 1 };
inline const char* cudaGetErrorString(cudaError_t) { return "emu"; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline cudaError_t cudaMemsetAsync(void* p, int v, size_t n, cudaStream_t) { memset(p, v, n); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { *p = (T*)aligned_alloc(256, (n + 255) / 256 * 256); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = (void*)1; return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }

namespace qpwc_emu {

[[noreturn]] inline void watchdog_abort(const char* what, int a, int b);
inline void mark(const char* what, int id);

class Barrier {
 public:
  void arrive_and_wait(int n, const char* what = "barrier", int id = -1) {
    mark(what, id);
    std::unique_lock<std::mutex> lk(mu_);
    const unsigned gen = gen_;
    if (++count_ == n) { count_ = 0; ++gen_; cv_.notify_all(); }
    else if (!cv_.wait_for(lk, std::chrono::seconds(20), [&] { return gen != gen_; })) watchdog_abort(what, id, count_);
  }
 private:
  std::mutex mu_;
  std::condition_variable cv_;
  int count_ = 0;
  unsigned gen_ = 0;
};

struct WarpState { Barrier bar; uint32_t buf[32]; int nthreads = 32; };

struct BlockState {
  std::vector<const char*> where;   // last wait entered by each thread (debug)
  std::vector<int> where_id;
  unsigned char* smem = nullptr;
  int nthreads = 0;
  Barrier bar;
  Barrier named[16];
  std::vector<WarpState> warps;
};

inline thread_local BlockState* tl_block = nullptr;
inline thread_local int tl_tid = 0;

inline unsigned char* dyn_smem() { return tl_block->smem; }

}  // namespace qpwc_emu

inline thread_local dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace qpwc_emu {
inline void mark(const char* what, int id) {
  if (tl_block) { tl_block->where[tl_tid] = what; tl_block->where_id[tl_tid] = id; }
}
// a barrier nobody completes within 20 s is a protocol bug: report who is stuck instead of hanging
[[noreturn]] inline void watchdog_abort(const char* what, int a, int b) {
  fprintf(stderr, "[qpwc_emu] DEADLOCK: thread %d of block %u stuck in %s (id/parity=%d, state=%d)\n",
          tl_tid, blockIdx.x, what, a, b);
  for (int t = 0; t < tl_block->nthreads; ++t)
    fprintf(stderr, "  t%d: %s %d\n", t, tl_block->where[t], tl_block->where_id[t]);
  fflush(stderr);
  abort();
}
}  // namespace qpwc_emu

namespace qpwc_emu {

template <class F>
inline void launch(dim3 grid, dim3 block, size_t smem_bytes, F body) {
  const int nt = (int)(block.x * block.y * block.z);
  BlockState bs;
  bs.nthreads = nt;
  bs.smem = (unsigned char*)aligned_alloc(1024, ((smem_bytes + 1023) / 1024 + 1) * 1024);
  bs.warps = std::vector<WarpState>((nt + 31) / 32);
  bs.where.assign(nt, "running"); bs.where_id.assign(nt, 0);
  for (size_t w = 0; w < bs.warps.size(); ++w) bs.warps[w].nthreads = std::min(32, nt - (int)w * 32);
  const long long nblocks = (long long)grid.x * grid.y * grid.z;
  std::vector<std::thread> ths;
  ths.reserve(nt);
  for (int t = 0; t < nt; ++t)
    ths.emplace_back([&, t] {
      tl_block = &bs;
      tl_tid = t;
      blockDim = block; gridDim = grid;
      threadIdx = dim3(t % block.x, (t / block.x) % block.y, t / (block.x * block.y));
      for (long long b = 0; b < nblocks; ++b) {
        blockIdx = dim3((unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((long long)grid.x * grid.y)));
        body();
        bs.bar.arrive_and_wait(nt, "block-end");  // block boundary: shared memory is reused by the next block
      }
    });
  for (auto& th : ths) th.join();
  free(bs.smem);
}

inline uint32_t shfl_exchange(uint32_t v, int src_lane_of_me /* computed by caller */, bool) { (void)v; (void)src_lane_of_me; return 0; }

}  // namespace qpwc_emu

inline void __syncthreads() { qpwc_emu::tl_block->bar.arrive_and_wait(qpwc_emu::tl_block->nthreads, "__syncthreads"); }
inline void __syncwarp(unsigned = 0xffffffffu) {
  auto& w = qpwc_emu::tl_block->warps[qpwc_emu::tl_tid / 32];
  w.bar.arrive_and_wait(w.nthreads, "__syncwarp");
}
// bar.sync id, nthreads
inline void qpwc_emu_named_barrier(int id, int nthreads) { qpwc_emu::tl_block->named[id].arrive_and_wait(nthreads, "bar.sync(named)", id); }

template <class T, class SrcFn>
inline T qpwc_emu_shfl(T v, SrcFn src_of) {
  static_assert(sizeof(T) == 4, "32-bit shuffles only");
  auto& w = qpwc_emu::tl_block->warps[qpwc_emu::tl_tid / 32];
  const int lane = qpwc_emu::tl_tid % 32;
  uint32_t bits; memcpy(&bits, &v, 4);
  w.buf[lane] = bits;
  w.bar.arrive_and_wait(w.nthreads);
  int src = src_of(lane);
  if (src < 0 || src >= w.nthreads) src = lane;
  const uint32_t r = w.buf[src];
  w.bar.arrive_and_wait(w.nthreads);
  T out; memcpy(&out, &r, 4);
  return out;
}
template <class T> inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return qpwc_emu_shfl(v, [m](int l) { return l ^ m; }); }
template <class T> inline T __shfl_down_sync(unsigned, T v, int d, int = 32) { return qpwc_emu_shfl(v, [d](int l) { return l + d > 31 ? l : l + d; }); }
template <class T> inline T __shfl_up_sync(unsigned, T v, int d, int = 32) { return qpwc_emu_shfl(v, [d](int l) { return l - d < 0 ? l : l - d; }); }
template <class T> inline T __shfl_sync(unsigned, T v, int s, int = 32) { return qpwc_emu_shfl(v, [s](int) { return s & 31; }); }

template <class T> inline T __ldg(const T* p) { return *p; }

inline float atomicAdd(float* p, float v) {
  uint32_t* u = reinterpret_cast<uint32_t*>(p);
  uint32_t old = __atomic_load_n(u, __ATOMIC_RELAXED), neu;
  float f;
  do {
    memcpy(&f, &old, 4);
    const float s = f + v;
    memcpy(&neu, &s, 4);
  } while (!__atomic_compare_exchange_n(u, &old, neu, false, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return f;
}
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_RELAXED); }

inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return {std::fmaf(a.x, b.x, c.x), std::fmaf(a.y, b.y, c.y)}; }
inline float __fadd_rn(float a, float b) { return a + b; }
inline float __fsub_rn(float a, float b) { return a - b; }
inline float __fmul_rn(float a, float b) { return a * b; }
inline int __float2int_rz(float v) {
  if (!(v == v)) return 0;
  if (v >= 2147483647.0f) return 2147483647;
  if (v <= -2147483648.0f) return (-2147483647 - 1);
  return (int)v;
}
using std::max;
using std::min;
