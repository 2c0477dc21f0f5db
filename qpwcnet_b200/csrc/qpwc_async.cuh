// qpwc_async.cuh -- thin wrappers over the sm_90+/sm_100a asynchronous machinery used by the tiled
// kernels: mbarrier, TMA tensor loads (cp.async.bulk.tensor), named barriers, setmaxnreg.
// Under -DQPWC_EMU (tests/emu, CPU harness) each wrapper has a functional stand-in so that the same
// kernel body -- pipeline phases, swizzled addressing, barrier protocol -- runs on the CPU.
#pragma once
#include "qpwc_common.cuh"

#ifndef QPWC_EMU
#include <cuda.h>  // CUtensorMap, CU_TENSOR_MAP_*
#endif

namespace qpwc {

// TMA SWIZZLE_32B: byte-address bit 4 ^= bit 7 (pattern repeats every 256 B; buffers 256-B aligned)
__host__ __device__ __forceinline__ uint32_t swz32(uint32_t byte_off) {
  return byte_off ^ (((byte_off >> 7) & 1u) << 4);
}
// TMA SWIZZLE_64B: bits [4:5] ^= bits [7:8] (period 512 B)
__host__ __device__ __forceinline__ uint32_t swz64(uint32_t byte_off) {
  return byte_off ^ (((byte_off >> 7) & 3u) << 4);
}
// TMA SWIZZLE_128B: bits [4:6] ^= bits [7:9] (period 1024 B)
__host__ __device__ __forceinline__ uint32_t swz128(uint32_t byte_off) {
  return byte_off ^ (((byte_off >> 7) & 7u) << 4);
}
// swizzle matching a pixel stride of PXB bytes (32 -> SWIZZLE_32B, 64 -> SWIZZLE_64B, 128 -> SWIZZLE_128B)
template <int PXB> __host__ __device__ __forceinline__ uint32_t swz(uint32_t byte_off) {
  return PXB == 32 ? swz32(byte_off) : (PXB == 64 ? swz64(byte_off) : swz128(byte_off));
}

#ifndef QPWC_EMU
// =============================================================================== real hardware
typedef CUtensorMap TensorMap;
#define QPWC_GRID_CONSTANT __grid_constant__

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.expect_tx.shared::cta.b64 st, [%0], %1;\n\t}" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// Same wait for the helper warps (TMA issue, store agents, repack): with a suspend-time hint the
// hardware parks the thread until the phase completes or ~the hint (ns) elapses, instead of
// re-issuing try_wait every ~20 cycles.  The tiled correlation is issue-bound: the un-hinted spin
// loops of its 4 helper warps were 26 % of all issued instructions (ncu, profiles/README.md).
__device__ __forceinline__ void mbar_wait_parked(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred P1;\n\t"
      "LAB_WAIT:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t"
      "@P1 bra DONE;\n\t"
      "bra LAB_WAIT;\n\t"
      "DONE:\n\t"
      "}" ::"r"(smem_u32(bar)), "r"(parity), "r"(20000u) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const TensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
// 4-D tiled load global -> shared, completion signalled on `bar` as transaction bytes
__device__ __forceinline__ void tma_load_4d(void* smem_dst, const TensorMap* tm, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(smem_u32(smem_dst)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
// L2 prefetch of a 4-D tile (no shared-memory destination, no completion signal)
__device__ __forceinline__ void tma_prefetch_l2_4d(const TensorMap* tm, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// generic-proxy writes (st.shared) -> async-proxy reads (bulk store): order them
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
// asynchronous 1-D bulk store shared -> global (TMA engine); bytes % 16 == 0, both 16-byte aligned
__device__ __forceinline__ void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
               ::"l"(reinterpret_cast<uint64_t>(gdst)), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
// asynchronous 4-D tensor store shared -> global (TMA engine): the whole box in one instruction, clipped
// at the tensor's extent; completion is tracked by the same bulk groups as bulk_store
__device__ __forceinline__ void tma_store_4d(const TensorMap* tm, const void* smem_src, int c0, int c1, int c2, int c3) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
// L2 eviction-priority policies (createpolicy) and the hinted forms of the stores
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ void tma_store_4d_hint(const TensorMap* tm, const void* smem_src, int c0, int c1, int c2, int c3, uint64_t policy) {
  asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group.L2::cache_hint [%0, {%2, %3, %4, %5}], [%1], %6;"
               ::"l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(smem_src)), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(policy) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// wait until at most N of this thread's bulk groups still have to READ their shared-memory source
template <int N> __device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// opaque 8-byte shared-memory load (address in the shared window): ptxas may not merge or sink it
__device__ __forceinline__ float2 lds_f2(uint32_t saddr) {
  float2 v;
  asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(saddr));
  return v;
}
#define QPWC_SMEM_ADDR(p) smem_u32(p)
template <int N> __device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// Host: encode a 4-D tiled tensor map over a dense NHWC fp32 tensor, dims (C, W, H, B), box
// (boxC, boxW, boxH, 1), SWIZZLE_32B, zero fill out of bounds, 128-B L2 promotion.
// Returns false (and sets the error) on failure.
bool make_tmap_nhwc(TensorMap* tm, const float* base, int B, int H, int W, int C, int boxC, int boxW, int boxH);  // swizzle = boxC*4 bytes
// dense (B,H,W,81) cost volume seen as (216, W*81/216, H, B): a box (216, 6, th, 1) is the 16-pixel x th-row
// output tile of the tensor-core kernels (one tensor store per tile).  Needs W % 8 == 0.
bool make_tmap_cv_tiles(TensorMap* tm, float* out, int B, int H, int W, int th);
// dense NCHW fp32 tensor, dims (W, H, C, B), box (boxW, boxH, boxC, 1), no swizzle (channel-planar tiles)
bool make_tmap_nchw(TensorMap* tm, const float* base, int B, int C, int H, int W, int boxW, int boxH, int boxC);

#else
// ============================================================================ CPU emulation
#define QPWC_GRID_CONSTANT
struct TensorMap {
  const float* base;
  long long dim[4];      // C, W, H, B  (NCHW maps: W, H, C, B)
  long long stride[4];   // element strides
  int box[4];
  int swizzle;           // 0 = none, 1 = SWIZZLE_32B, 3 = SWIZZLE_64B (mask of address bits [7:8] folded into [4:5])
};
struct EmuMbar { uint16_t init; uint16_t pending; int32_t tx; };  // init bit 15 = phase parity
static_assert(sizeof(EmuMbar) == 8, "mbarrier is 8 bytes");

namespace emu_detail {
inline std::mutex& mu() { static std::mutex m; return m; }
inline std::condition_variable& cv() { static std::condition_variable c; return c; }
inline void settle(EmuMbar* b) {
  if (b->pending == 0 && b->tx == 0) {
    b->init ^= 0x8000u;
    b->pending = b->init & 0x7fffu;
    cv().notify_all();
  }
}
}  // namespace emu_detail

inline void mbar_init(uint64_t* bar, uint32_t count) {
  EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
  b->init = (uint16_t)count; b->pending = (uint16_t)count; b->tx = 0;
}
inline void fence_mbar_init() {}
inline void mbar_arrive(uint64_t* bar) {
  std::lock_guard<std::mutex> lk(emu_detail::mu());
  EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
  b->pending--; emu_detail::settle(b);
}
inline void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  std::lock_guard<std::mutex> lk(emu_detail::mu());
  EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
  b->tx += (int32_t)bytes; b->pending--; emu_detail::settle(b);
}
inline void mbar_wait(uint64_t* bar, uint32_t parity) {
  qpwc_emu::mark("mbar_wait", (int)(reinterpret_cast<uintptr_t>(bar) & 0xff) * 10 + (int)parity);
  std::unique_lock<std::mutex> lk(emu_detail::mu());
  EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
  if (!emu_detail::cv().wait_for(lk, std::chrono::seconds(20), [&] { return (uint32_t)((b->init >> 15) & 1u) != parity; }))
    qpwc_emu::watchdog_abort("mbar_wait", (int)parity, (int)b->pending * 1000000 + b->tx);
}
inline void mbar_wait_parked(uint64_t* bar, uint32_t parity) { mbar_wait(bar, parity); }
inline void tma_prefetch_desc(const TensorMap*) {}
inline void tma_prefetch_l2_4d(const TensorMap*, int, int, int, int) {}
inline void tma_load_4d(void* smem_dst, const TensorMap* tm, uint64_t* bar, int c0, int c1, int c2, int c3) {
  const int co[4] = {c0, c1, c2, c3};
  unsigned char* dst = static_cast<unsigned char*>(smem_dst);
  uint32_t lin = 0, bytes = 0;
  for (int e3 = 0; e3 < tm->box[3]; ++e3)
    for (int e2 = 0; e2 < tm->box[2]; ++e2)
      for (int e1 = 0; e1 < tm->box[1]; ++e1)
        for (int e0 = 0; e0 < tm->box[0]; ++e0, lin += 4, bytes += 4) {
          const long long x[4] = {co[0] + e0, co[1] + e1, co[2] + e2, co[3] + e3};
          float v = 0.f;
          bool in = true;
          for (int k = 0; k < 4; ++k) in = in && x[k] >= 0 && x[k] < tm->dim[k];
          if (in) v = tm->base[x[0] * tm->stride[0] + x[1] * tm->stride[1] + x[2] * tm->stride[2] + x[3] * tm->stride[3]];
          // hardware swizzles on absolute shared-memory address bits
          const uintptr_t abs = reinterpret_cast<uintptr_t>(dst) + lin;
          const uintptr_t phys = abs ^ (((abs >> 7) & (uintptr_t)tm->swizzle) << 4);
          *reinterpret_cast<float*>(phys) = v;
        }
  std::lock_guard<std::mutex> lk(emu_detail::mu());
  EmuMbar* b = reinterpret_cast<EmuMbar*>(bar);
  b->tx -= (int32_t)bytes; emu_detail::settle(b);
}
inline void fence_proxy_async() {}
inline void bulk_store(void* gdst, const void* smem_src, uint32_t bytes) { memcpy(gdst, smem_src, bytes); }
inline void bulk_commit() {}
template <int N> inline void bulk_wait_read() {}
inline void named_bar_sync(int id, int nthreads) { qpwc_emu_named_barrier(id, nthreads); }
// emulation: "shared window addresses" are offsets from the CTA's dynamic shared memory base
inline float2 lds_f2(uint32_t saddr) { return *reinterpret_cast<const float2*>(qpwc_emu::dyn_smem() + saddr); }
#define QPWC_SMEM_ADDR(p) ((uint32_t)(reinterpret_cast<const unsigned char*>(p) - qpwc_emu::dyn_smem()))
template <int N> inline void setmaxnreg_inc() {}
template <int N> inline void setmaxnreg_dec() {}

inline bool make_tmap_nhwc(TensorMap* tm, const float* base, int B, int H, int W, int C, int boxC, int boxW, int boxH) {
  tm->base = base;
  tm->dim[0] = C; tm->dim[1] = W; tm->dim[2] = H; tm->dim[3] = B;
  tm->stride[0] = 1; tm->stride[1] = C; tm->stride[2] = (long long)W * C; tm->stride[3] = (long long)H * W * C;
  tm->box[0] = boxC; tm->box[1] = boxW; tm->box[2] = boxH; tm->box[3] = 1;
  tm->swizzle = boxC * 4 == 32 ? 1 : (boxC * 4 == 64 ? 3 : 7);
  return true;
}
inline bool make_tmap_nchw(TensorMap* tm, const float* base, int B, int C, int H, int W, int boxW, int boxH, int boxC) {
  tm->base = base;
  tm->dim[0] = W; tm->dim[1] = H; tm->dim[2] = C; tm->dim[3] = B;
  tm->stride[0] = 1; tm->stride[1] = W; tm->stride[2] = (long long)H * W; tm->stride[3] = (long long)C * H * W;
  tm->box[0] = boxW; tm->box[1] = boxH; tm->box[2] = boxC; tm->box[3] = 1;
  tm->swizzle = 0;
  return true;
}
#endif

}  // namespace qpwc
