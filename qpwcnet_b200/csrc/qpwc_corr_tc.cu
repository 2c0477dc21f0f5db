// qpwc_corr_tc.cu -- local-correlation cost volume on the 5th-generation tensor cores (tcgen05 + TMEM),
// fp32 in / fp32 out through a 3xTF32 operand split, for sm_100a.
//
//   out[b,i,j,(di+4)*9+(dj+4)] = lrelu( (1/C) sum_c P[b,i,j,c] * N[b,i+di,j+dj,c] ),  N == 0 outside
//   (CostVolume / CostVolumeV2, qpwcnet/core/layers.py:72-100,117-132)
//
// Formulation.  An 8x16 tile of first-frame pixels (M = 128 rows of A) against its 16x24 halo tile of
// second-frame pixels (N = 384 rows of B) is the dense product D[128 x 384] = A[128 x C] . B[384 x C]^T;
// the 81 wanted displacements of pixel (r, c) are the 9x9 window D[(r,c), (r+m, c+k)], m,k in [0,9) --
// 81/384 = 21 % of the dense tile.  fp32 operands are split x = hi + lo with hi = the fp32 word itself
// (the tensor core ignores the 13 low mantissa bits of a tf32 operand: tools/ubench/tf32x3_tile.cu,
// raw and masked words give bit-identical products) and lo = rna_tf32(x - trunc_tf32(x)); three
// tcgen05.mma kind::tf32 passes hi.hi + lo.hi + hi.lo accumulate in fp32 in TMEM.  Measured error
// against fp64: <= 8e-7 * mean|p.n| (fp32 FFMA in sequence: 2.4e-7; the contract is 1e-5), and the
// tile sustains 104 clk per 128x192x8 MMA = 2.9x the useful FMA rate of the FFMA kernel
// (profiles/r02_tf32x3_ubench.txt).
//
// Structure: one persistent CTA per SM, warp-specialised.
//   warp 0      TMA producer: boxes of 8 x 8 first-frame pixels (so that TMEM lane quadrant q holds a 4x8
//               pixel block) and 24 x 8 second-frame pixels (half tiles; zero fill == ZeroPadding2D), with
//               64- or 128-byte channel rows (SWIZZLE_64B / SWIZZLE_128B = the canonical K-major UMMA
//               layouts): the TMA engine retires about one row per 2 clk whatever its width, and with
//               32-byte rows it, not the tensor core, set the pace (profiles/README.md).
//   warp 1      MMA issuer (one thread): per 8-channel K step and half of N the three passes of the
//               split product; tcgen05.commit releases operands / publishes the accumulator.
//   warps 4-7   operand split: the first-frame operand goes through TMEM (warp = lane quadrant, lane =
//               pixel: hi and lo written with tcgen05.st into spare columns), the second-frame lo
//               is computed elementwise in shared memory.
//   warps 8-19  epilogue: warp = TMEM lane quadrant = 4x8 pixel block; the three warps of a quadrant
//               read four rows each of its 12 x 16 accumulator window (tcgen05.ld 32x32b.x16), shift the
//               lane's 9 columns into place, scale, leaky-relu and store them into the tile's NHWC
//               staging image; the 8 staged rows leave as TMA bulk stores (16 x 324 contiguous bytes).
#include <stdlib.h>

#include "qpwc_async.cuh"

namespace qpwc {

#ifndef QPWC_EMU

// dev instrumentation (tools/tc_trace.py): when set, CTA 0 of the resident kernel stamps clock64() at the
// hand-over points of its first QPWC_TRACE_TILES tiles: g_tc_trace[tile * 24 + event]
#define QPWC_TRACE_TILES 40
__device__ long long* g_tc_trace = nullptr;
// (`t` = g_tc_trace read ONCE per thread at kernel start: a load per stamp stalled the issuing threads ~40 clk each)
__device__ __forceinline__ void tc_stamp(long long* t, uint32_t tile, int ev) {
  if (t != nullptr && tile < QPWC_TRACE_TILES) t[tile * 24 + ev] = clock64();
}

// dev (QPWC_ABLATE bit 9): spin instead of parking on the role hand-over barriers
__device__ __forceinline__ void tc_wait(uint64_t* bar, uint32_t parity, int spin) {
  if (spin) mbar_wait(bar, parity); else mbar_wait_parked(bar, parity);
}

struct TcCfg {
  static constexpr int NDISP = 81;
  static constexpr int TH = 8, TW = 16, NROW = 16, NCOL = 24, NHALF = 192;
  static constexpr int ROW_FLOATS = TW * NDISP;           // staging: the tile's NHWC output image, 8 rows of 16 x 81
  static constexpr int STAGING_BYTES = TH * ROW_FLOATS * 4;
  static constexpr int NEPI = 12;                         // epilogue warps: 3 per TMEM lane quadrant
  // warpgroup 0: TMA warp, MMA warp (+2 idle); warpgroup 1: 4 split warps; warpgroups 2-4: epilogue.
  // Roles are warpgroup-aligned so that setmaxnreg can move registers to the epilogue warps, which
  // hold four accumulator rows (64 registers) at a time.
  static constexpr int W_SPLIT = 4, W_EPI = 8;
  static constexpr int NTHREADS = (W_EPI + NEPI) * 32;
  static constexpr int REG_CTRL = 40, REG_SPLIT = 56, REG_EPI = 128;  // launch: 96 each; the decs free exactly what the incs take
  static_assert(128 * REG_CTRL + 128 * REG_SPLIT + NEPI * 32 * REG_EPI <= 96 * NTHREADS, "register budget (the pool is the launch allocation)");
  // TMEM columns: accumulator 0..383 (two halves of N); first-frame operand (hi, lo) from column 384:
  // 16 columns per 8-channel K step (8 hi + 8 lo)
  static constexpr int TM_A = 384;
  // ---- resident kernel (C <= 32): 128-byte channel rows.  One landing buffer for A (128 px; it is moved to
  // TMEM at once), a ring of three raw half-tile B blocks (192 px; a block is loaded two tile periods
  // before it is needed), a ring of TWO lo blocks (the lo of a block is written after the block two
  // allocations earlier has died -- one tile period of slack, no HBM latency on that path), and TWO staging
  // images: the TMA store of tile T drains (~2000 clk for 41 KB) while tile T+1 is being staged.
  static constexpr int R_PXB = 128;
  static constexpr int RA_BYTES = 128 * R_PXB, RB_BYTES = NHALF * R_PXB;          // 16 KB, 24 KB
  static constexpr int NBLK = 3, NLO = 2, NSTG = 2;
  static constexpr int R_OFF_A = 0, R_OFF_B = RA_BYTES, R_OFF_BLO = R_OFF_B + NBLK * RB_BYTES;
  static constexpr int R_OFF_STAGING = R_OFF_BLO + NLO * RB_BYTES;
  static constexpr int R_OFF_BARS = R_OFF_STAGING + NSTG * STAGING_BYTES;
  static constexpr int R_SMEM_BYTES = R_OFF_BARS + 25 * 8 + 16;
  // ---- streaming kernel (any C % 8 == 0): stages of 16 channels, 64-byte rows; per stage both B half
  // tiles (raw + lo) and an A landing buffer
  static constexpr int S_PXB = 64, S_KC = 16, NST = 3;
  static constexpr int SB_BYTES = 2 * NHALF * S_PXB, SA_BYTES = 128 * S_PXB;      // 24 KB (raw; lo follows), 8 KB
  static constexpr int S_OFF_A = NST * 2 * SB_BYTES;
  static constexpr int S_OFF_STAGING = S_OFF_A + NST * SA_BYTES;
  static constexpr int S_OFF_BARS = S_OFF_STAGING + STAGING_BYTES;
  static constexpr int S_SMEM_BYTES = S_OFF_BARS + (4 * NST + 5) * 8 + 16;
  static_assert(S_SMEM_BYTES <= 232448 && R_SMEM_BYTES <= 232448, "shared memory budget");
  static_assert(TM_A + 2 * 64 <= 512 && TM_A + NST * 32 <= 512, "TMEM columns");
};

// cute::UMMA::SmemDescriptor, K-major, swizzle = PXB bytes (pixel rows PXB bytes apart, 8-row core
// groups contiguous): start address, LBO (unused for swizzled K-major), SBO = 8 rows, version 1,
// layout type (SWIZZLE_32B 6, SWIZZLE_64B 4, SWIZZLE_128B 2).  A K step of 8 tf32 inside a wider row is
// addressed by advancing the start address by 32 bytes.
template <int PXB> __device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  constexpr uint64_t layout = PXB == 32 ? 6 : (PXB == 64 ? 4 : 2);
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)1 << 16) | ((uint64_t)((8 * PXB) >> 4) << 32) |
         ((uint64_t)1 << 46) | (layout << 61);
}
// kind::tf32, fp32 accumulator, A and B K-major, M = 128, N = 192
static constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TcCfg::NHALF >> 3) << 17) | ((128u >> 4) << 24);

// A operand from tensor memory (lane = pixel, one column per channel of the K step), B from shared memory
__device__ __forceinline__ void umma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(kIdesc), "r"(acc) : "memory");
}
// the three passes of the split product for one K step and one half of N: hi.hi + lo.hi + hi.lo.
// `d_raw` / `d_lo` are the shared-memory descriptors of the raw / lo B operand; the issuing thread's own
// instruction stream is on the critical path (tools/ubench/tf32x3_tile.cu: 96 clk per MMA with
// descriptors prepared, ~130 with the address arithmetic in the loop), so descriptors are built once
// per block and advanced by adding to the start-address field (16-byte units, no carry: < 2^14).
__device__ __forceinline__ void umma_x3_ts(uint32_t d, uint32_t a_tm, uint64_t d_raw, uint64_t d_lo, uint32_t acc) {
  umma_tf32_ts(d, a_tm, d_raw, acc);
  umma_tf32_ts(d, a_tm + 8, d_raw, 1u);
  umma_tf32_ts(d, a_tm, d_lo, 1u);
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ float tf32_lo(float x) {
  // lo = x - trunc_tf32(x) is exact (<= 13 significant bits).  Adding half a tf32 ulp to its bit
  // pattern makes the tensor core's own truncation of the operand a round-to-nearest (cvt.rna.tf32
  // is a ~8-instruction sequence on sm_100; this is 3 instructions per element).
  const float r = x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  return __uint_as_float(__float_as_uint(r) + 0x1000u);
}
// lo for `bytes` of a raw operand block, elementwise (any swizzle: same offsets)
__device__ __forceinline__ void split_block(const unsigned char* src, unsigned char* dst, int bytes, int st) {
  for (int off = st * 16; off < bytes; off += 128 * 16) {
    const float4 v = *reinterpret_cast<const float4*>(src + off);
    *reinterpret_cast<float4*>(dst + off) = make_float4(tf32_lo(v.x), tf32_lo(v.y), tf32_lo(v.z), tf32_lo(v.w));
  }
}
// First-frame operand -> TMEM: this thread's pixel (row m of the landing buffer, PXB-byte rows), NKS K
// steps of 8 channels: hi (the fp32 words) to columns col0 + 16*ks, lo to col0 + 16*ks + 8 of the
// thread's TMEM lane.  `taddr` = TMEM address of (this warp's lane quadrant, col0).
template <int PXB, int NKS>
__device__ __forceinline__ void a_to_tmem(const unsigned char* landing, int m, uint32_t taddr, int nks, float a_scale) {
#pragma unroll
  for (int ks = 0; ks < NKS; ++ks) {
    if (ks >= nks) break;
    const float4 v0 = *reinterpret_cast<const float4*>(landing + swz<PXB>((uint32_t)(m * PXB + ks * 32)));
    const float4 v1 = *reinterpret_cast<const float4*>(landing + swz<PXB>((uint32_t)(m * PXB + ks * 32 + 16)));
    // a_scale = 1/C when C is a power of two (exact: the mean's division folded into the operand), else 1
    const float x[8] = {v0.x * a_scale, v0.y * a_scale, v0.z * a_scale, v0.w * a_scale,
                        v1.x * a_scale, v1.y * a_scale, v1.z * a_scale, v1.w * a_scale};
    uint32_t hi[8], lo[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) { hi[e] = __float_as_uint(x[e]); lo[e] = __float_as_uint(tf32_lo(x[e])); }
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr + (uint32_t)(ks * 16)), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]), "r"(hi[4]), "r"(hi[5]), "r"(hi[6]), "r"(hi[7]) : "memory");
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 ::"r"(taddr + (uint32_t)(ks * 16 + 8)), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]), "r"(lo[4]), "r"(lo[5]), "r"(lo[6]), "r"(lo[7]) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
#define QPWC_TMEM_LD16(v, taddr)                                                                          \
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];" \
               : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),   \
                 "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]) \
               : "r"(taddr))

// Epilogue of one tile, executed by the 12 epilogue warps (ew = 0..11).  TMEM lane quadrant q = warp
// id % 4 holds the 4x8 pixel block (rows 4*(q&1).., cols 8*(q>>1)..) of the tile; the three warps of a
// quadrant take four rows each of its 12 x 16 accumulator window, every four-row group lies in one
// half of N (tfull/tempty[h]).  Per window row a lane needs the 9 columns c..c+8 of the 16 it
// loaded: a 3-level conditional shift on the bits of c brings them to registers 0..8, they are scaled,
// leaky-relu'd and stored as row (y - r) of the lane's 9x9 band in the NHWC staging image of the tile
// (compact 81-float pixels: at most 2-way bank conflicts).  After a barrier the 8 staged rows leave
// as TMA bulk stores (one contiguous 16 x 324-byte run each).
__device__ __forceinline__ void tc_epilogue_tile(float* staging, uint32_t tmem, uint64_t* tfull, uint64_t* tempty, uint64_t* sfree,
                                                 uint32_t tcount, int q, int part, int ew, int lane,
                                                 float* __restrict__ out, int b, int i0, int j0, int H, int W,
                                                 long long ops, float inv_c, float slope, int ablate, int chb, int qo,
                                                 const TensorMap* tmO, int nstg, long long* trace) {
  // chb / qo: first output channel and channel pitch of a displacement row.  d = 4: (0, 9), the 81
  // results of a pixel are contiguous.  d = 8 runs as four 9x9 windows of the 17x17 range: qo = 17,
  // chb = (oi + 4) * 17 + (oj + 4) for the window offset (oi, oj) in {-4, +4}^2.
  using Cfg = TcCfg;
  // nstg staging images (1 or 2): tile T uses image T % nstg, free once the store of tile T - nstg has read it
  staging += (tcount & (uint32_t)(nstg - 1)) * (Cfg::STAGING_BYTES / 4);
  const int rb = q & 1, cb = q >> 1, r = lane >> 3, c = lane & 7;
  const int h = (rb + part) >= 2 ? 1 : 0;
  const int y0 = 4 * part;
  const bool c4 = c & 4, c2 = c & 2, c1 = c & 1;
  const bool fast = inv_c == 1.f && slope >= 0.f && slope <= 1.f;
  const int spin = (ablate >> 9) & 1;
  // row (y - r) of the band of pixel (rb*4 + r, cb*8 + c) starts at lane_base[9*y]
  float* lane_base = staging + (rb * 4 + r) * Cfg::ROW_FLOATS + (cb * 8 + c) * Cfg::NDISP - 9 * r;
  const uint32_t tq = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)((rb * 4 + y0) * Cfg::NCOL + cb * 8);
  tc_wait(&tfull[h], tcount & 1u, spin);
  tc_fence_after();
  if (ew == 0) tc_stamp(trace, tcount, 10); else if (ew == 8) tc_stamp(trace, tcount, 15);
  // all four rows of this warp's share into registers first: the accumulator half is released as soon
  // as the loads have landed, before the shifting and staging work
  uint32_t u[4][16];
  if (!(ablate & 1)) {
#pragma unroll
    for (int yy = 0; yy < 4; ++yy) QPWC_TMEM_LD16(u[yy], tq + (uint32_t)(yy * Cfg::NCOL));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncwarp();
  if (lane == 0) mbar_arrive(&tempty[h]);        // this half of the accumulator may be overwritten
  if (ew == 0) tc_stamp(trace, tcount, 11); else if (ew == 8) tc_stamp(trace, tcount, 16);
  // the bulk stores of the previous tile must have finished reading the staging image (waited for
  // here, after the accumulator loads, so that the engine's reads overlap the wait for the MMAs)
  // (one lane issues and tracks the stores; the others learn through `sfree`, so a warp that has its
  // rows goes straight on to the shifting/staging work without waiting for the other half's loads)
  if (ew == 0 && lane == 0) {
    if (nstg == 2) bulk_wait_read<1>(); else bulk_wait_read<0>();
    mbar_arrive(sfree);
  }
  mbar_wait(sfree, tcount & 1u);
  if (ew == 0) tc_stamp(trace, tcount, 12);
  if (!(ablate & 1)) {
#pragma unroll
    for (int yy = 0; yy < 4; ++yy) {
      float v[16];
#pragma unroll
      for (int x = 0; x < 16; ++x) v[x] = __uint_as_float(u[yy][x]);
#pragma unroll
      for (int x = 0; x < 12; ++x) v[x] = c4 ? v[x + 4] : v[x];
#pragma unroll
      for (int x = 0; x < 10; ++x) v[x] = c2 ? v[x + 2] : v[x];
#pragma unroll
      for (int x = 0; x < 9; ++x) v[x] = c1 ? v[x + 1] : v[x];
      const int y = y0 + yy;
      if ((unsigned)(y - r) <= 8u) {
        float* sp = lane_base + 9 * y;
        if (fast) {      // scale folded into A, 0 <= slope <= 1: leaky relu = max(v, slope * v)
#pragma unroll
          for (int k = 0; k < 9; ++k) sp[k] = fmaxf(v[k], v[k] * slope);
        } else {
#pragma unroll
          for (int k = 0; k < 9; ++k) sp[k] = lrelu(v[k] * inv_c, slope);
        }
      }
    }
  }
  fence_proxy_async();                           // staging stores -> bulk-store (async proxy) reads
  if (ew == 0) tc_stamp(trace, tcount, 13); else if (ew == 8) tc_stamp(trace, tcount, 17);
  named_bar_sync(2, Cfg::NEPI * 32);             // all pixel blocks staged
  if (ew == 0) tc_stamp(trace, tcount, 14);
  if (ablate & 16) return;
  const int wv = min(Cfg::TW, W - j0);
  const bool bulk = ops == Cfg::NDISP && qo == 9 && (W & 3) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0;
  if (tmO != nullptr) {
    // W % 8 == 0: the tile is one box of the (216, W*81/216, H, B) view of the cost volume -- a single
    // tensor store, clipped at the image edge (eight per-row bulk copies cost the issuing lane ~1100 clk
    // per tile, on the critical chain stores -> staging free -> next tile staged: profiles/r02_tc_trace.txt)
    if (ew == 0 && lane == 0) {
      // the cost volume is written once and not read again by this kernel: evict-first, so that the 297 MB output
      // stream does not push the operands (the warped frame the warp kernel has just left in L2) out
      if (ablate & 1024) tma_store_4d(tmO, staging, 0, (j0 >> 4) * 6, i0, b);
      else tma_store_4d_hint(tmO, staging, 0, (j0 >> 4) * 6, i0, b, l2_policy_evict_first());
      bulk_commit();
    }
  } else if (bulk) {
    if (ew == 0 && lane == 0) {
#pragma unroll
      for (int row = 0; row < Cfg::TH; ++row)
        if (i0 + row < H)
          bulk_store(out + ((size_t)((size_t)b * H + i0 + row) * W + j0) * Cfg::NDISP, staging + row * Cfg::ROW_FLOATS,
                     (uint32_t)(wv * Cfg::NDISP * 4));
      bulk_commit();
    }
  } else {
    // strided / unaligned output: 8 rows x 3 thirds = 24 items over the 12 warps, coalesced scalar copies
    // (measured alternative: a thread per displacement walking the pixels with immediate offsets has a third
    // of the instructions and is 20 % SLOWER at search range 8 -- a warp that writes a run of pixels one after
    // the other lets their 36-byte runs combine on the way to L2)
    const int n = wv * Cfg::NDISP;
#pragma unroll
    for (int it = 0; it < 2; ++it) {
      const int item = ew + it * Cfg::NEPI, row = item / 3, third = item - row * 3, i = i0 + row;
      if (i >= H) continue;
      const float* src = staging + row * Cfg::ROW_FLOATS;
      float* dst = out + ((size_t)((size_t)b * H + i) * W + j0) * (size_t)ops;
      const int g0 = third * 432 + lane, gend = min(n, (third + 1) * 432);
      int p = g0 / Cfg::NDISP, e = g0 - p * Cfg::NDISP;  // element e of pixel p
      for (int g = g0; g < gend; g += 32) {
        const int m = e / 9;
        dst[(size_t)p * ops + chb + m * qo + (e - m * 9)] = src[g];
        e += 32;
        if (e >= Cfg::NDISP) { e -= Cfg::NDISP; ++p; }
      }
    }
    named_bar_sync(1, Cfg::NEPI * 32);           // every warp has read its share before the image is reused
  }
}

__device__ __forceinline__ uint32_t tc_prologue(uint64_t* bars, int nbars_1, int nbars_4, int nbars_e, uint32_t* tmem_slot,
                                                const unsigned char* smem, int tid, int warp) {
  // bars: [0, nbars_1) count 1, then nbars_4 with count 4 (split warps), then nbars_e with count 6
  // (epilogue warps of one half).  Warp 1 allocates all 512 TMEM columns (384 used; one CTA per SM).
  if (tid == 0) {
    if (smem_u32(smem) & 1023u) __trap();
    for (int k = 0; k < nbars_1 + nbars_4 + nbars_e; ++k)
      mbar_init(&bars[k], k < nbars_1 ? 1 : (k < nbars_1 + nbars_4 ? 4 : TcCfg::NEPI / 2));
    fence_mbar_init();
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  return *tmem_slot;
}
__device__ __forceinline__ void tc_teardown(uint32_t tmem, int warp) {
  bulk_wait_read<0>();  // (epilogue issuer lanes) shared memory must outlive the engine's reads
  tc_fence_before();
  __syncthreads();
  if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

// ---------------------------------------------------------------------------------------------
// Streaming kernel: any C % 8 == 0.  Stages of 16 channels flow through NST buffers; the MMAs of a
// stage cover both halves of N, so the accumulator is published once per tile.
__global__ void __launch_bounds__(TcCfg::NTHREADS, 1)
corr_fwd_tc_stream_kernel(const __grid_constant__ TensorMap tmP, const __grid_constant__ TensorMap tmN,
                          const __grid_constant__ TensorMap tmO, int tstore, float* __restrict__ out, int B, int H, int W, int C, float slope, long long ops,
                          int tiles_x, int tiles_y, int ntiles, int ablate, int nwin, int qo) {
  // nwin = 1: the 9x9 window is the whole search range (d = 4).  nwin = 4 (d = 8): a tile index also selects
  // one of the four 9x9 windows of the 17x17 range -- window offset (oi, oj) = (+-4, +-4) of the second-frame
  // tile, first output channel chb.  The window index runs FASTEST, so the four windows of a tile are
  // written by neighbouring CTAs at about the same time and their 36-byte runs merge into whole sectors
  // in L2 (one launch per window cost a DRAM read-modify-write per run: the image is as large as L2).
  // ablate (dev, QPWC_ABLATE): bit0 no accumulator drain, bit1 no operand split, bit2 no MMAs, bit3 no loads,
  // bit4 no copy-out, bit5 force the streaming kernel, bit6 per-row bulk copies instead of the tensor store, bit7 no L2 prefetch
  using Cfg = TcCfg;
  constexpr int PXB = Cfg::S_PXB;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::S_OFF_BARS);
  uint64_t* raw_full = bars;                      // count 1 (+tx)
  uint64_t* stage_free = raw_full + Cfg::NST;     // count 1 (commit)
  uint64_t* tfull = stage_free + Cfg::NST;        // [2] count 1 (commit)
  uint64_t* sfree = tfull + 2;                    // count 1: staging image read by the previous tile's bulk stores
  uint64_t* lo_full = sfree + 1;                  // count 4: second-frame lo of the stage written
  uint64_t* a_full = lo_full + Cfg::NST;          // count 4: first-frame operand of the stage in TMEM
  uint64_t* tempty = a_full + Cfg::NST;           // [2] count 6
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* staging = reinterpret_cast<float*>(smem + Cfg::S_OFF_STAGING);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, spin = (ablate >> 9) & 1;
  const int nstages = (C + Cfg::S_KC - 1) / Cfg::S_KC;  // per tile; the last one may hold a single K step
  const uint32_t tmem = tc_prologue(bars, 2 * Cfg::NST + 3, 2 * Cfg::NST, 2, tmem_slot, smem, tid, warp);
#define QPWC_SB(s) (smem + (s) * 2 * Cfg::SB_BYTES)
#define QPWC_SA(s) (smem + Cfg::S_OFF_A + (s) * Cfg::SA_BYTES)

  if (warp < Cfg::W_SPLIT) {
   setmaxnreg_dec<Cfg::REG_CTRL>();
   if (warp == 0) {
    if (lane == 0) {  // ------------------------------------------------------------- TMA producer
      tma_prefetch_desc(&tmP);
      tma_prefetch_desc(&tmN);
      // linear walk over (tile, stage) items; the item PF stages ahead is prefetched into L2 so that the
      // load proper finds it there (the NST stages in flight cover an L2 hit, not an HBM round trip)
      constexpr int PF = 2 * Cfg::NST;
      const int my_tiles = ntiles > (int)blockIdx.x ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
      const int nitems = my_tiles * nstages;
      auto coords = [&](int n, int& c, int& i0, int& j0, int& b, int& oi, int& oj) {   // second-frame tile origin: (i0 - 4 + oi, j0 - 4 + oj)
        const int tl = n / nstages;
        c = n - tl * nstages;
        const int tw = (int)blockIdx.x + tl * (int)gridDim.x, win = tw % nwin, tile = tw / nwin;
        const int tx = tile % tiles_x, rest = tile / tiles_x, ty = rest % tiles_y;
        b = rest / tiles_y; i0 = ty * Cfg::TH; j0 = tx * Cfg::TW;
        oi = nwin == 1 ? 0 : ((win >> 1) * 8 - 4); oj = nwin == 1 ? 0 : ((win & 1) * 8 - 4);
      };
      auto prefetch = [&](int n) {
        if (n >= nitems || (ablate & (8 | 128))) return;
        int c, i0, j0, b, oi, oj;
        coords(n, c, i0, j0, b, oi, oj);
        tma_prefetch_l2_4d(&tmP, c * Cfg::S_KC, j0, i0, b);
        tma_prefetch_l2_4d(&tmP, c * Cfg::S_KC, j0 + 8, i0, b);
        tma_prefetch_l2_4d(&tmN, c * Cfg::S_KC, j0 - 4 + oj, i0 - 4 + oi, b);
        tma_prefetch_l2_4d(&tmN, c * Cfg::S_KC, j0 - 4 + oj, i0 + 4 + oi, b);
      };
      for (int n = 0; n < PF; ++n) prefetch(n);
      for (int n = 0; n < nitems; ++n) {
        const uint32_t g = (uint32_t)n;
        int c, i0, j0, b, oi, oj;
        coords(n, c, i0, j0, b, oi, oj);
        {
          const int s = (int)(g % Cfg::NST);
          tc_wait(&stage_free[s], ((g / Cfg::NST) & 1u) ^ 1u, spin);
          prefetch(n + PF);
          if (ablate & 8) { mbar_arrive(&raw_full[s]); continue; }
          mbar_arrive_expect_tx(&raw_full[s], Cfg::SA_BYTES + Cfg::SB_BYTES);
          // A: two boxes of 8 cols x 8 rows; TMEM lane quadrant q = 2*cb + rb is the block rows 4*rb.., cols 8*cb..
          tma_load_4d(QPWC_SA(s), &tmP, &raw_full[s], c * Cfg::S_KC, j0, i0, b);
          tma_load_4d(QPWC_SA(s) + Cfg::SA_BYTES / 2, &tmP, &raw_full[s], c * Cfg::S_KC, j0 + 8, i0, b);
          tma_load_4d(QPWC_SB(s), &tmN, &raw_full[s], c * Cfg::S_KC, j0 - 4 + oj, i0 - 4 + oi, b);
          tma_load_4d(QPWC_SB(s) + Cfg::SB_BYTES / 2, &tmN, &raw_full[s], c * Cfg::S_KC, j0 - 4 + oj, i0 + 4 + oi, b);
        }
      }
    }
   } else if (warp == 1) {
    if (lane == 0) {  // --------------------------------------------------------------- MMA issuer
      uint32_t g = 0, tcount = 0;
      for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++tcount) {
        mbar_wait(&tempty[0], (tcount & 1u) ^ 1u);  // the epilogue has drained both halves
        mbar_wait(&tempty[1], (tcount & 1u) ^ 1u);
        for (int c = 0; c < nstages; ++c, ++g) {
          const int s = (int)(g % Cfg::NST);
          // the passes that read only the raw second-frame words (hi.hi, lo.hi) go out as soon as the
          // first-frame operand is in TMEM; the elementwise lo of the 24 KB second-frame block is
          // computed by the split warps meanwhile, and the hi.lo pass follows
          mbar_wait(&raw_full[s], (g / Cfg::NST) & 1u);
          mbar_wait(&a_full[s], (g / Cfg::NST) & 1u);
          tc_fence_after();
          const uint64_t d_raw = umma_desc<PXB>(smem_u32(QPWC_SB(s))), d_lo = d_raw + (uint64_t)(Cfg::SB_BYTES >> 4);
          const uint32_t a_tm = tmem + (uint32_t)(Cfg::TM_A + s * 32);
          const int nks = min(2, (C - c * Cfg::S_KC) / 8);
          // consecutive MMAs alternate between the two accumulator halves (back-to-back MMAs into the same
          // accumulator serialise on it: half-major order inside a stage measured 3-5 us slower per level)
          if (!(ablate & 4))
            for (int ks = 0; ks < nks; ++ks)
#pragma unroll
              for (int ps = 0; ps < 2; ++ps)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                  const uint64_t off = (uint64_t)((h * (Cfg::SB_BYTES / 2) + ks * 32) >> 4);
                  umma_tf32_ts(tmem + (uint32_t)(h * Cfg::NHALF), a_tm + ks * 16 + ps * 8, d_raw + off, (c | ks | ps) ? 1u : 0u);
                }
          mbar_wait(&lo_full[s], (g / Cfg::NST) & 1u);
          tc_fence_after();
          if (!(ablate & 4))
            for (int ks = 0; ks < nks; ++ks)
#pragma unroll
              for (int h = 0; h < 2; ++h) {
                const uint64_t off = (uint64_t)((h * (Cfg::SB_BYTES / 2) + ks * 32) >> 4);
                umma_tf32_ts(tmem + (uint32_t)(h * Cfg::NHALF), a_tm + ks * 16, d_lo + off, 1u);
              }
          umma_commit(&stage_free[s]);  // stage reusable once these MMAs have read it
        }
        umma_commit(&tfull[0]);
        umma_commit(&tfull[1]);
      }
    }
   }
  } else if (warp < Cfg::W_EPI) {  // ------------------------------------------------- operand split
    setmaxnreg_dec<Cfg::REG_SPLIT>();
    const int st = tid - Cfg::W_SPLIT * 32, qd = warp & 3, m = qd * 32 + lane;
    const float a_scale = (C & (C - 1)) == 0 ? 1.f / (float)C : 1.f;
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      for (int c = 0; c < nstages; ++c, ++g) {
        const int s = (int)(g % Cfg::NST);
        tc_wait(&raw_full[s], (g / Cfg::NST) & 1u, spin);
        tc_fence_after();  // (the MMAs that read TMEM buffer s completed before the stage was reloaded)
        if (!(ablate & 2))
          a_to_tmem<PXB, 2>(QPWC_SA(s), m, tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)(Cfg::TM_A + s * 32), min(2, (C - c * Cfg::S_KC) / 8), a_scale);
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&a_full[s]);
        if (!(ablate & 2)) split_block(QPWC_SB(s), QPWC_SB(s) + Cfg::SB_BYTES, Cfg::SB_BYTES, st);
        fence_proxy_async();  // generic-proxy stores -> tensor-core (async proxy) reads
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&lo_full[s]);
      }
    }
  } else {  // --------------------------------------------------------------------------- epilogue
    setmaxnreg_inc<Cfg::REG_EPI>();
    const int ew = warp - Cfg::W_EPI, q = warp & 3, part = ew >> 2;
    const float inv_c = (C & (C - 1)) == 0 ? 1.f : 1.f / (float)C;  // power-of-two C: folded into the first-frame operand
    uint32_t tcount = 0;
    if (tstore && ew == 0 && lane == 0) tma_prefetch_desc(&tmO);
    for (int tw = blockIdx.x; tw < ntiles; tw += gridDim.x, ++tcount) {
      const int win = tw % nwin, tile = tw / nwin;
      const int tx = tile % tiles_x, rest = tile / tiles_x, ty = rest % tiles_y, b = rest / tiles_y;
      const int chb = nwin == 1 ? 0 : ((win >> 1) * 8) * qo + (win & 1) * 8;
      tc_epilogue_tile(staging, tmem, tfull, tempty, sfree, tcount, q, part, ew, lane, out, b, ty * Cfg::TH, tx * Cfg::TW,
                       H, W, ops, inv_c, slope, ablate, chb, qo, tstore ? &tmO : nullptr, 1, nullptr);
    }
  }
  tc_teardown(tmem, warp);
#undef QPWC_SA
#undef QPWC_SB
}

// ---------------------------------------------------------------------------------------------
// Resident kernel: C <= 32.  All channels of a tile fit on chip, so (1) the MMAs run half-major -- the
// first half of N is published (and drained by the epilogue) while the second is being computed, and
// the next tile's first half while this tile's second drains -- and (2) a CTA walks vertically
// consecutive tiles of one strip: the lower half-tile of second-frame rows of tile k is the upper
// half-tile of tile k+1 and stays where it is ("rolling rows"): only 8 new second-frame rows are
// loaded and split per tile instead of 16.  The raw half-tile blocks form a ring of three, so the block a
// tile frees after its first pass is refilled two tile periods before it is needed; their lo parts
// live in a ring of two (block n reuses the lo slot of block n-2, which dies one tile period before
// block n's lo is read).  The first-frame operand goes through TMEM (double-buffered there, one landing
// buffer in shared memory).  The shared memory this frees holds a second staging image, so the TMA
// store of a tile drains while the next tile is staged (profiles/r02_tc_trace.txt: with one image the
// epilogue warps waited ~1300 of 3900 clk per tile for the previous store to finish reading it).
__global__ void __launch_bounds__(TcCfg::NTHREADS, 1)
corr_fwd_tc_res_kernel(const __grid_constant__ TensorMap tmP, const __grid_constant__ TensorMap tmN,
                       const __grid_constant__ TensorMap tmO, int tstore, float* __restrict__ out, int B, int H, int W, int C, float slope, long long ops,
                       int tiles_x, int tiles_y, int seg, int nseg, int nunits, int ablate, int nwin, int qo) {
  using Cfg = TcCfg;
  constexpr int PXB = Cfg::R_PXB;
  extern __shared__ __align__(1024) unsigned char smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::R_OFF_BARS);
  uint64_t* afull = bars;          // count 1 (+tx): A landed in the landing buffer
  uint64_t* bfull = afull + 1;     // [3] count 1 (+tx): raw B block p landed
  uint64_t* afree = bfull + 3;     // [2] count 1 (commit): TMEM A buffer a no longer read
  uint64_t* bfree = afree + 2;     // [3] count 1 (commit): raw B block p no longer read
  uint64_t* lofree = bfree + 3;    // [2] count 1 (commit): lo block l no longer read
  uint64_t* tfull = lofree + 2;    // [2] count 1 (commit)
  uint64_t* sfree = tfull + 2;     // count 1: staging image of this tile no longer read by the store of two tiles ago
  uint64_t* arawfree = sfree + 1;  // count 4: landing buffer consumed by the split warps
  uint64_t* alo = arawfree + 1;    // [2] count 4: TMEM A buffer a written
  uint64_t* blo = alo + 2;         // [2] count 4: lo block l written
  uint64_t* tempty = blo + 2;      // [2] count 6
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);
  float* staging = reinterpret_cast<float*>(smem + Cfg::R_OFF_STAGING);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, spin = (ablate >> 9) & 1;
  long long* const trace = blockIdx.x == 0 ? g_tc_trace : nullptr;
  const int nks = C / 8;           // K steps (<= 4); channels C..31 of the 128-byte rows are TMA zero fill
  const uint32_t tmem = tc_prologue(bars, 14, 5, 2, tmem_slot, smem, tid, warp);
#define QPWC_ABUF (smem + Cfg::R_OFF_A)
#define QPWC_BRAW(p) (smem + Cfg::R_OFF_B + (p) * Cfg::RB_BYTES)
#define QPWC_BLO(l) (smem + Cfg::R_OFF_BLO + (l) * Cfg::RB_BYTES)
  // every role walks the same tile sequence; `ring` counts the B blocks allocated so far (raw slot = ring % 3,
  // lo slot = ring % 2): the first tile of a segment takes two fresh blocks (top, bot), every other tile
  // inherits its top from the previous tile's bot and takes one fresh block
#define QPWC_FOR_UNITS                                                                           \
  for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {                                \
    const int win = unit % nwin, uu_ = unit / nwin;   /* window fastest (see the streaming kernel) */ \
    const int oi = nwin == 1 ? 0 : ((win >> 1) * 8 - 4), oj = nwin == 1 ? 0 : ((win & 1) * 8 - 4); \
    const int chb = nwin == 1 ? 0 : (oi + 4) * qo + (oj + 4);                                    \
    (void)oi; (void)oj; (void)chb;                                                               \
    const int tx = uu_ % tiles_x, rest_ = uu_ / tiles_x, sg = rest_ % nseg, b = rest_ / nseg;    \
    const int ty0 = sg * seg, nt = min(seg, tiles_y - ty0), j0 = tx * Cfg::TW;                   \
    (void)b; (void)j0;                                                                           \
    for (int k = 0; k < nt; ++k, ++T) {                                                          \
      const int i0 = (ty0 + k) * Cfg::TH;                                                        \
      if (k == 0) { top = (int)(ring % 3u); ltop = (int)(ring & 1u); ++ring; } else { top = bot; ltop = lbot; } \
      bot = (int)(ring % 3u); lbot = (int)(ring & 1u); ++ring;                                   \
      (void)i0; (void)ltop; (void)lbot;
#define QPWC_PAR(mask, idx) (((mask) >> (idx)) & 1u)

  if (warp < Cfg::W_SPLIT) {
   setmaxnreg_dec<Cfg::REG_CTRL>();
   if (warp == 0) {
    if (lane == 0) {  // ------------------------------------------------------------- TMA producer
      tma_prefetch_desc(&tmP);
      tma_prefetch_desc(&tmN);
      uint32_t T = 0, ring = 0, ub = 0;
      int top = 0, bot = 0, ltop = 0, lbot = 0;
      QPWC_FOR_UNITS
        tc_wait(arawfree, (T & 1u) ^ 1u, spin);  // the landing buffer has been moved to TMEM
        tc_stamp(trace, T, 0);
        if (ablate & 8) mbar_arrive(afull);
        else {
          mbar_arrive_expect_tx(afull, Cfg::RA_BYTES);
          tma_load_4d(QPWC_ABUF, &tmP, afull, 0, j0, i0, b);
          tma_load_4d(QPWC_ABUF + Cfg::RA_BYTES / 2, &tmP, afull, 0, j0 + 8, i0, b);
        }
        // second-frame half-tile blocks: rows i0-4..i0+3 (top) and i0+4..i0+11 (bot); only the first
        // tile of a segment loads its top
        for (int hb = (k == 0 ? 0 : 1); hb < 2; ++hb) {
          const int p = hb ? bot : top;
          tc_wait(&bfree[p], QPWC_PAR(ub, p) ^ 1u, spin);
          ub ^= 1u << p;
          if (hb) tc_stamp(trace, T, 1);
          if (ablate & 8) { mbar_arrive(&bfull[p]); continue; }
          mbar_arrive_expect_tx(&bfull[p], Cfg::RB_BYTES);
          tma_load_4d(QPWC_BRAW(p), &tmN, &bfull[p], 0, j0 - 4 + oj, i0 - 4 + oi + hb * 8, b);
        }
      }}
    }
   } else if (warp == 1) {
    if (lane == 0) {  // --------------------------------------------------------------- MMA issuer
      uint32_t T = 0, ring = 0, ma = 0, mb = 0;
      int top = 0, bot = 0, ltop = 0, lbot = 0;
      QPWC_FOR_UNITS
        const int a = (int)(T & 1u);
        const uint32_t a_tm = tmem + (uint32_t)(Cfg::TM_A + a * 64);
        mbar_wait(&alo[a], QPWC_PAR(ma, a));
        ma ^= 1u << a;
        if (k == 0) { mbar_wait(&blo[ltop], QPWC_PAR(mb, ltop)); mb ^= 1u << ltop; }  // k > 0: waited for as `bot` of tile k-1
        mbar_wait(&tempty[0], (T & 1u) ^ 1u);
        tc_fence_after();
        tc_stamp(trace, T, 6);
        {
          const uint64_t d_raw = umma_desc<PXB>(smem_u32(QPWC_BRAW(top))), d_lo = umma_desc<PXB>(smem_u32(QPWC_BLO(ltop)));
          if (!(ablate & 4))
            for (int ks = 0; ks < nks; ++ks)
              umma_x3_ts(tmem, a_tm + ks * 16, d_raw + (uint64_t)(2 * ks), d_lo + (uint64_t)(2 * ks), ks > 0 ? 1u : 0u);
          umma_commit(&tfull[0]);
          umma_commit(&bfree[top]);     // the upper block is dead once these MMAs have read it
          umma_commit(&lofree[ltop]);
          tc_stamp(trace, T, 7);
        }
        mbar_wait(&blo[lbot], QPWC_PAR(mb, lbot));
        mb ^= 1u << lbot;
        mbar_wait(&tempty[1], (T & 1u) ^ 1u);
        tc_fence_after();
        tc_stamp(trace, T, 8);
        {
          const uint64_t d_raw = umma_desc<PXB>(smem_u32(QPWC_BRAW(bot))), d_lo = umma_desc<PXB>(smem_u32(QPWC_BLO(lbot)));
          if (!(ablate & 4))
            for (int ks = 0; ks < nks; ++ks)
              umma_x3_ts(tmem + Cfg::NHALF, a_tm + ks * 16, d_raw + (uint64_t)(2 * ks), d_lo + (uint64_t)(2 * ks), ks > 0 ? 1u : 0u);
          umma_commit(&tfull[1]);
          umma_commit(&afree[a]);
          if (k == nt - 1) { umma_commit(&bfree[bot]); umma_commit(&lofree[lbot]); }  // end of the segment: nobody inherits the lower block
          tc_stamp(trace, T, 9);
        }
      }}
    }
   }
  } else if (warp < Cfg::W_EPI) {  // ------------------------------------------------- operand split
    setmaxnreg_dec<Cfg::REG_SPLIT>();
    const int st = tid - Cfg::W_SPLIT * 32, qd = warp & 3;  // qd: the TMEM lane quadrant this warp may write
    const int m = qd * 32 + lane;                           // its pixel = row of A = TMEM lane
    const float a_scale = (C & (C - 1)) == 0 ? 1.f / (float)C : 1.f;
    uint32_t T = 0, ring = 0, jf = 0, jb = 0, jl = 0;
    int top = 0, bot = 0, ltop = 0, lbot = 0;
    QPWC_FOR_UNITS
      const int a = (int)(T & 1u);
      tc_wait(afull, T & 1u, spin);
      if (warp == Cfg::W_SPLIT) tc_stamp(trace, T, 2);
      tc_wait(&afree[a], QPWC_PAR(jf, a) ^ 1u, spin);  // the MMAs of two tiles ago have finished reading TMEM buffer a
      jf ^= 1u << a;
      tc_fence_after();
      if (!(ablate & 2))
        a_to_tmem<PXB, 4>(QPWC_ABUF, m, tmem + ((uint32_t)(qd * 32) << 16) + (uint32_t)(Cfg::TM_A + a * 64), nks, a_scale);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(arawfree); mbar_arrive(&alo[a]); }
      if (warp == Cfg::W_SPLIT) tc_stamp(trace, T, 3);
      for (int hb = (k == 0 ? 0 : 1); hb < 2; ++hb) {
        const int p = hb ? bot : top, l = hb ? lbot : ltop;
        tc_wait(&bfull[p], QPWC_PAR(jb, p), spin);
        jb ^= 1u << p;
        tc_wait(&lofree[l], QPWC_PAR(jl, l) ^ 1u, spin);  // the block that held this lo slot (two allocations ago) is dead
        jl ^= 1u << l;
        if (warp == Cfg::W_SPLIT && hb) tc_stamp(trace, T, 4);
        if (!(ablate & 2)) split_block(QPWC_BRAW(p), QPWC_BLO(l), Cfg::RB_BYTES, st);
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&blo[l]);
        if (warp == Cfg::W_SPLIT && hb) tc_stamp(trace, T, 5);
      }
    }}
  } else {  // --------------------------------------------------------------------------- epilogue
    setmaxnreg_inc<Cfg::REG_EPI>();
    const int ew = warp - Cfg::W_EPI, q = warp & 3, part = ew >> 2;
    const float inv_c = (C & (C - 1)) == 0 ? 1.f : 1.f / (float)C;  // power-of-two C: folded into the first-frame operand
    uint32_t T = 0, ring = 0;
    int top = 0, bot = 0, ltop = 0, lbot = 0;
    if (tstore && ew == 0 && lane == 0) tma_prefetch_desc(&tmO);
    QPWC_FOR_UNITS
      tc_epilogue_tile(staging, tmem, tfull, tempty, sfree, T, q, part, ew, lane, out, b, i0, j0, H, W, ops, inv_c, slope, ablate, chb, qo, tstore ? &tmO : nullptr, Cfg::NSTG, trace);
    }}
    (void)top; (void)bot;
  }
  tc_teardown(tmem, warp);
#undef QPWC_PAR
#undef QPWC_FOR_UNITS
#undef QPWC_ABUF
#undef QPWC_BRAW
#undef QPWC_BLO
}

int sm_count_cached();

extern "C" int qpwc_debug_tc_trace(long long* device_buffer) {   // dev tool, not part of include/qpwc.h
  return cudaMemcpyToSymbol(g_tc_trace, &device_buffer, sizeof(device_buffer)) == cudaSuccess ? 0 : 3;
}

int launch_corr_fwd_tc(const float* prv, const float* nxt, float* out, int B, int H, int W, int C, int d,
                       float slope, long long ops, cudaStream_t stream) {
  // domain: d == 4 (one 9x9 window) or d == 8 (four 9x9 windows of the 17x17 range, one launch each; the
  // shared lines di = 0 / dj = 0 are written twice, bit-identically), C a multiple of 8, 16-byte aligned inputs
  if ((d != 4 && d != 8) || (C & 7) || C < 8) return QPWC_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(prv) & 15) || (reinterpret_cast<uintptr_t>(nxt) & 15)) return QPWC_ERR_UNSUPPORTED;
  using Cfg = TcCfg;
  const int tiles_x = cdiv(W, Cfg::TW), tiles_y = cdiv(H, Cfg::TH);
  const long long nt = (long long)tiles_x * tiles_y * B;
  if (nt >= (1LL << 31)) return QPWC_ERR_UNSUPPORTED;
  static const int ablate = [] { const char* ev = getenv("QPWC_ABLATE"); return ev ? atoi(ev) : 0; }();  // dev switches, read once (thread-safe)
  const int sms = sm_count_cached();
  const bool resident = C <= 32 && !(ablate & 32);
  TensorMap tmP, tmN;
  if (resident) {
    if (!make_tmap_nhwc(&tmP, prv, B, H, W, C, 32, 8, 8)) return QPWC_ERR_CUDA;           // A: 8 cols x 8 rows x 128 B
    if (!make_tmap_nhwc(&tmN, nxt, B, H, W, C, 32, Cfg::NCOL, 8)) return QPWC_ERR_CUDA;   // B: 24 cols x 8 rows (half tile)
  } else {
    if (!make_tmap_nhwc(&tmP, prv, B, H, W, C, Cfg::S_KC, 8, 8)) return QPWC_ERR_CUDA;
    if (!make_tmap_nhwc(&tmN, nxt, B, H, W, C, Cfg::S_KC, Cfg::NCOL, 8)) return QPWC_ERR_CUDA;
  }
  // output tiles as single tensor stores when the cost volume is dense, d = 4 and W % 8 == 0
  const int tstore = (d == 4 && ops == Cfg::NDISP && (W & 7) == 0 && (reinterpret_cast<uintptr_t>(out) & 15) == 0 && !(ablate & 64)) ? 1 : 0;
  TensorMap tmO = tmP;
  if (tstore && !make_tmap_cv_tiles(&tmO, out, B, H, W, Cfg::TH)) return QPWC_ERR_CUDA;
  const int smem = resident ? Cfg::R_SMEM_BYTES : Cfg::S_SMEM_BYTES;
  const cudaError_t e = resident
      ? cudaFuncSetAttribute(corr_fwd_tc_res_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem)
      : cudaFuncSetAttribute(corr_fwd_tc_stream_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "corr_fwd_tc: smem attribute (%d B): %s", smem, cudaGetErrorString(e));
  // segments of vertically consecutive tiles (resident kernel); short enough that every SM gets several
  int seg = tiles_y;
  while (seg > 2 && (long long)tiles_x * B * cdiv(tiles_y, seg) < 4LL * sms) seg = cdiv(seg, 2);
  const int nseg = cdiv(tiles_y, seg);
  const int nunits = tiles_x * B * nseg, ntiles = (int)nt;
  const int nwin = d == 8 ? 4 : 1, qo = 2 * d + 1;
  if ((long long)nunits * nwin >= (1LL << 31) || nt * nwin >= (1LL << 31)) return QPWC_ERR_UNSUPPORTED;
  const int nu = nunits * nwin, ntw = ntiles * nwin;
  if (resident)
    corr_fwd_tc_res_kernel<<<nu < sms ? nu : sms, Cfg::NTHREADS, smem, stream>>>(
        tmP, tmN, tmO, tstore, out, B, H, W, C, slope, ops, tiles_x, tiles_y, seg, nseg, nu, ablate, nwin, qo);
  else
    corr_fwd_tc_stream_kernel<<<ntw < sms ? ntw : sms, Cfg::NTHREADS, smem, stream>>>(
        tmP, tmN, tmO, tstore, out, B, H, W, C, slope, ops, tiles_x, tiles_y, ntw, ablate, nwin, qo);
  const int rc = check_launch(resident ? "corr_fwd_tc_res" : "corr_fwd_tc_stream");
  if (rc != QPWC_OK) return rc;
  return QPWC_OK;
}

#else  // CPU emulation build: the tensor-core kernel has no stand-in; callers fall back to the FFMA kernels
int launch_corr_fwd_tc(const float*, const float*, float*, int, int, int, int, int, float, long long, cudaStream_t) {
  return QPWC_ERR_UNSUPPORTED;
}

#endif

}  // namespace qpwc
