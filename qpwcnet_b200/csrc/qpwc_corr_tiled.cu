// qpwc_corr_tiled.cu -- register-tiled FFMA local correlation for sm_100a, plain and fused with
// the bilinear warp of the second frame (the warped tensor lives only in shared memory).
//
//   out[b,i,j,(di+d)*(2d+1)+(dj+d)] = lrelu( (1/C) sum_c P[b,i,j,c] * N[b,i+di,j+dj,c] ),  N == 0 outside
//   P = prv;  N = nxt (CostVolume, qpwcnet/core/layers.py:72-100)  or  N = warp(nxt, flow)
//   (UpFlow, qpwcnet/core/non_layers.py:377-380).
//
// Why FFMA and not tcgen05: the contraction is banded -- every output pixel contracts against its
// own (2d+1)^2 neighbourhood -- and fp32 parity (<= 1e-5 rel) rules out tf32/bf16 operands.
//
// Decomposition (d = 4): one persistent CTA per SM walks tiles of TH x 56 first-frame pixels.
//   * consumer thread (ti, tc), tc in [0,64): second-frame column s = j0-4+tc of tile row ti.  Per
//     channel it forms the 9x9 outer product  acc[m][k] += P[i, s-(k-4)] * N[i+(m-4), s]
//     (a = 9 first-frame pixels left/right of s in row i, b = 9 second-frame pixels above/below
//     in column s): 18 shared-memory operands feed 81 FFMAs, every product is a valid
//     displacement pair.  acc[m][k] is output channel m*9+k of pixel column s-(k-4).
//   * operands are streamed through shared memory in chunks of 8 channels, 3 stages, TMA tensor
//     loads (zero fill outside the image == ZeroPadding2D) with SWIZZLE_32B so that the 16-byte
//     operand loads of 8 neighbouring pixels hit 8 distinct bank groups.
//   * a producer warpgroup drives the pipeline over mbarriers (full/empty per stage), across tile
//     boundaries.  Plain: one thread issues the two TMA loads per chunk.  Fused: the second-frame
//     stage is produced by the warpgroup itself -- 4 gathers + blend per halo pixel, bit-identical
//     to the stand-alone warp kernel -- so warped features never reach HBM.
//   * epilogue: accumulators -> (x 1/C, leaky relu) -> private per-row staging slots in shared
//     memory, half a row (28 px) at a time (transposes the per-thread 9x9 blocks into the NHWC
//     81-vector) -> one asynchronous TMA bulk store per contiguous 28 x 324-byte run.
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "qpwc_async.cuh"
#include "qpwc_upsample.cuh"

namespace qpwc {

template <int TH_, int WARP_, int MODE_>
struct TiledCfg {
  // (round 2: a 2-CTA-per-SM variant -- two 2-row CTAs per SM so that one's tile prologue/epilogue
  // overlaps the other's FFMA loop -- measured identical to this one at every level and was dropped:
  // the kernel is throughput-, not lock-step-bound.)
  static constexpr int D = 4, Q = 2 * D + 1, NDISP = Q * Q;
  static constexpr int TH = TH_, TWT = 64, TW = TWT - 2 * D;  // 56 pixel columns per tile
  static constexpr int WARP = WARP_, MODE = MODE_;
  // channels per pipeline stage: 8 (32 B/pixel, SWIZZLE_32B); the fused variant uses 16 (64 B,
  // SWIZZLE_64B) -- its gathers cost one L1 wavefront per distinct 128-byte line *per instruction*
  // (measured ~2 clk each), so wider per-pixel reads halve the producer's L1 time
  static constexpr int KC = WARP_ ? 16 : 8, PXB = KC * 4, NQ = KC / 4;
  static constexpr int NROW = TH + 2 * D, NCOL = TWT, PCOL = TW;         // P tile: the 56 valid columns only
  static constexpr int NST = WARP_ ? 2 : (TH_ <= 4 ? 4 : 3);
  static constexpr int P_BYTES = TH * PCOL * PXB, N_BYTES = NROW * NCOL * PXB;
  static constexpr int STAGE_BYTES = P_BYTES + N_BYTES;
  // producers (128-thread warpgroups, the unit of setmaxnreg): plain = 1 TMA thread + 3 store-agent
  // warps; fused = 8 gather warps -- the gathers are latency-bound, they need the parallelism
  // (measured: 4 gather warps with 200-register consumers are slower at every level)
  static constexpr int NCONS = TH * TWT, NPROD = WARP_ ? 256 : 128, NTHREADS = NCONS + NPROD;
  // Epilogue staging.  AGENT (plain variant): one full-row slot per tile row; consumer warps only
  // deposit their accumulators and move on, a store-agent warp hands the row to the TMA engine.
  // Fused variant: the taps table needs the room => half-row slots, stores issued by the row itself.
  static constexpr int AGENT = WARP ? 0 : 1;
  static constexpr int HALF = TW / 2;
  static constexpr int SLOT_BYTES = (((AGENT ? TW : HALF) * NDISP * 4 + 127) / 128) * 128;
  static constexpr int NSLOT = 1;
  static constexpr int TAPS_BYTES = WARP ? NROW * NCOL * 32 : 0;
  static constexpr int OFF_STAGING = NST * STAGE_BYTES;
  static constexpr int OFF_TAPS = OFF_STAGING + TH * NSLOT * SLOT_BYTES;
  static constexpr int OFF_BARS = OFF_TAPS + TAPS_BYTES;
  static constexpr int SMEM_BYTES = OFF_BARS + 2 * NST * 8 + 2 * TH * 8;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget (227 KB)");
  static constexpr int FULL_COUNT = 1 + (WARP ? NPROD / 32 : 0);
  static constexpr int EMPTY_COUNT = NCONS / 32;
  // register split (setmaxnreg), sum <= 65536:  plain TH=4: 256*200 + 128*56;  fused TH=4: 256*176 + 256*80
  static constexpr int REG_CONS = WARP_ ? 176 : 200;
  static constexpr int REG_PROD = WARP_ ? 80 : 56;
  static constexpr int UB = 2;  // fused producer: units in flight per thread (8 independent 16-byte gathers)
  static_assert(TH_ <= 4, "4-row tiles at most (register budget of the consumers)");
  static_assert(TW % 8 == 0 && P_BYTES % 512 == 0 && N_BYTES % 512 == 0 && TH <= 14, "tile shape");
  static_assert(NCONS * REG_CONS + NPROD * REG_PROD <= 65536, "register budget");
};

struct TapsEntry { int o00, o01, o10, o11; float w00, w01, w10, w11; };  // 32 bytes

// Work enumeration shared by the three roles (producer, consumers, store agent): a CTA takes
// segments segi = blockIdx.x, blockIdx.x + gridDim.x, ...; a segment is `seg` vertically consecutive
// tiles of one (batch, window, tile column) strip.  seg == 1 is the plain tile list.  seg > 1 is
// used by the fused variant when all channels fit in the pipeline stages (C <= KC*NST): consecutive
// tiles then share TH+2D-TH = 2D halo rows, which stay in shared memory ("rolling" rows) -- only TH
// new warped rows are produced per tile instead of TH+2D.
#define QPWC_FOR_TILES_BEGIN                                                                     \
  for (int segi = blockIdx.x; segi < nsegs; segi += gridDim.x) {                                 \
    /* the window index runs fastest: the four windows of a tile (search range 8) are written by    */ \
    /* neighbouring CTAs at the same time, so their 36-byte runs merge into whole sectors in L2      */ \
    /* instead of each costing a DRAM read-modify-write one image (~L2 size) later                  */ \
    const int win = segi % nwin, segw_ = segi / nwin;                                            \
    const int tx = segw_ % tiles_x, rest_ = segw_ / tiles_x; /* then tile column: concurrent     */ \
    const int sy = rest_ % segs_per_strip, b = rest_ / segs_per_strip; /* CTAs share image rows  */ \
    const int oi = nwin == 1 ? 0 : ((win >> 1) * 8 - 4), oj = nwin == 1 ? 0 : ((win & 1) * 8 - 4); \
    const int seg_ = Cfg::WARP ? seg : 1; /* plain variants: compile-time single-tile segments */ \
    for (int tseg = 0; tseg < seg_; ++tseg) {                                                    \
      const int ty = sy * seg_ + tseg;                                                           \
      if (ty >= tiles_y) break;                                                                  \
      const int i0 = ty * TH, j0 = tx * TW;                                                      \
      (void)oi; (void)oj; (void)b;
#define QPWC_FOR_TILES_END }}

// ---------------------------------------------------------------------------------------------
template <class Cfg>
__global__ void __launch_bounds__(Cfg::NTHREADS, 1)
corr_fwd_tiled_kernel(const QPWC_GRID_CONSTANT TensorMap tmP, const QPWC_GRID_CONSTANT TensorMap tmN,
                      const float* __restrict__ nxt, const float* __restrict__ flow,
                      float* __restrict__ out, int B, int H, int W, int C, float slope, long long ops,
                      int tiles_x, int tiles_y, int nsegs, int segs_per_strip, int seg, int ablate, int nwin,
                      int dsearch, float up_scale) {
  // up_scale != 0 (fused variant): `flow` is the coarse flow (B, H/2, W/2, 2) and the taps are set up
  // from up_scale * bilinear_x2(flow) (qpwc_upsample.cuh) -- the upsampled flow is never read back
  // nwin = 1: the 9x9 displacement window is the whole search range (d = 4).  nwin = 4 (d = 8): the
  // 17x17 range is covered by four 9x9 windows centred at (+-4, +-4); a tile index then also
  // selects the window, whose offset (oi, oj) shifts the second-frame tile and the output channels
  // (the shared lines di = 0 / dj = 0 are produced twice, bit-identically).
  constexpr int D = Cfg::D, Q = Cfg::Q, NDISP = Cfg::NDISP, TH = Cfg::TH, TW = Cfg::TW;
  constexpr int NCOL = Cfg::NCOL, PCOL = Cfg::PCOL, NROW = Cfg::NROW, NST = Cfg::NST, KC = Cfg::KC;
  constexpr int NCONS = Cfg::NCONS, NPROD = Cfg::NPROD, HALF = Cfg::HALF;

  // dynamic shared memory starts at the CTA's window base (no static __shared__ in this kernel):
  // 1024-byte aligned, which the swizzle arithmetic relies on (checked below, once per CTA)
  QPWC_DYN_SMEM(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
  uint64_t* empty = full + NST;
  uint64_t* sfull = empty + NST;   // per tile row: both consumer warps deposited their outputs
  uint64_t* sfree = sfull + TH;    // per tile row: the store of the previous tile has drained the slot

  const int tid = threadIdx.x;
  const int nchunks = (C + KC - 1) / KC;
  const bool fixed_stage = Cfg::WARP && nchunks <= NST;  // fused variant with all channels resident

  if (tid == 0) {
#ifndef QPWC_EMU
    if (smem_u32(smem) & 511u) __trap();
#endif
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], Cfg::FULL_COUNT); mbar_init(&empty[s], Cfg::EMPTY_COUNT); }
    for (int r = 0; r < TH; ++r) { mbar_init(&sfull[r], NCOL / 32); mbar_init(&sfree[r], 1); }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= NCONS) {
    // ========================================================================== producers
    setmaxnreg_dec<Cfg::REG_PROD>();
    const int ptid = tid - NCONS;
    if (ptid == 0) { tma_prefetch_desc(&tmP); if (!Cfg::WARP) tma_prefetch_desc(&tmN); }
    // Plain variant: the pipeline is driven by one thread.  The other producer threads must NOT
    // idle along on the empty barriers: nothing would gate on them, so a slow one could fall two
    // phases behind and alias the parity wait (found by the CPU emulation harness).  In the fused
    // variant every producer warp arrives on `full`, which keeps all of them within one phase.
    if (Cfg::AGENT && ptid >= 32) {
      // ---- store agent (plain variant): waits for each row's slot, issues its bulk store, and
      // releases the slots once the engine has read them.  Consumers never wait for stores.
      if (ablate & 2) return;
      const int lane = ptid & 31;
      const int aw = (ptid >> 5) - 1;  // agent warp 0..2: rows r with r % 3 == aw
      uint32_t n = 0;
      QPWC_FOR_TILES_BEGIN
        const int twv = min(TW, W - j0);
        for (int r = aw; r < TH; r += 3) {
          mbar_wait_parked(&sfull[r], n & 1u);
          const int i = i0 + r;
          if (i < H) {
            const float* slot = reinterpret_cast<const float*>(smem + Cfg::OFF_STAGING + r * Cfg::SLOT_BYTES);
            float* dst = out + ((size_t)((size_t)b * H + i) * W + j0) * (size_t)ops;
            const int cnt = twv * NDISP;
            if (nwin == 1 && ops == NDISP && (cnt & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
              if (lane == 0) bulk_store(dst, slot, (uint32_t)cnt * 4u);
            } else if (nwin == 1 && ops == NDISP) {  // unaligned row: plain coalesced copy
              for (int e = lane; e < cnt; e += 32) dst[e] = slot[e];
            } else {  // strided output (concat buffer) and/or a window of a wider search range
              const int qo = 2 * dsearch + 1;
              const int chb = (oi - D + dsearch) * qo + (oj - D + dsearch);
              // lanes own channels (their (m,k) -> output channel map is loop invariant), pixels are walked
              // in order: coalesced runs per pixel, no per-element divisions
              for (int ch = lane; ch < NDISP; ch += 32) {
                const int m = ch / Q, k = ch - m * Q, och = chb + m * qo + k;
                for (int px = 0; px < cnt / NDISP; ++px) dst[(size_t)px * ops + och] = slot[px * NDISP + ch];
              }
            }
          }
        }
        if (lane == 0) { bulk_commit(); bulk_wait_read<0>(); }
        __syncwarp();
        if (lane == 0) for (int r = aw; r < TH; r += 3) mbar_arrive(&sfree[r]);
        ++n;
      QPWC_FOR_TILES_END
      return;
    }
    if (!Cfg::WARP && ptid != 0) return;
    if (ablate & 4) return;  // dev ablation: no loads at all
    TapsEntry* taps = reinterpret_cast<TapsEntry*>(smem + Cfg::OFF_TAPS);
    uint32_t g = 0, ptile = 0;
    QPWC_FOR_TILES_BEGIN
      // rolling rows: the first tile of a segment produces the whole (TH+2D)-row halo window, the
      // following ones only its last TH rows; row rr of the window lives in ring slot (rot+rr)%NROW
      const int rr0 = (tseg == 0) ? 0 : (NROW - TH);
      const int rot = (tseg * TH) % NROW;
      const int nprod_px = (NROW - rr0) * NCOL;
      if (Cfg::WARP) {
        // per-tile table of sampling taps for every halo pixel of the warped second frame
        named_bar_sync(15, NPROD);  // previous tile's last chunk no longer reads the table
        for (int p = ptid; p < nprod_px; p += NPROD) {
          const int r = i0 - D + oi + rr0 + p / NCOL, s = j0 - D + oj + p % NCOL;
          TapsEntry e;
          e.o00 = -1; e.o01 = e.o10 = e.o11 = 0; e.w00 = e.w01 = e.w10 = e.w11 = 0.f;
          if (r >= 0 && r < H && s >= 0 && s < W) {  // outside: zero padding of the warped frame
            const float2 f = up_scale != 0.f
                ? up2_flow(flow + (size_t)b * (H / 2) * (W / 2) * 2, r, s, H / 2, W / 2, up_scale)
                : __ldg(reinterpret_cast<const float2*>(flow) + ((size_t)b * H * W + (size_t)r * W + s));
            const Taps t = make_taps<Cfg::MODE>(r, s, f.x, f.y, H, W);
            e.o00 = t.o00; e.o01 = t.o01; e.o10 = t.o10; e.o11 = t.o11;
            e.w00 = t.w00; e.w01 = t.w01; e.w10 = t.w10; e.w11 = t.w11;
          }
          taps[p] = e;
        }
        named_bar_sync(15, NPROD);
      }
      for (int c = 0; c < nchunks; ++c, ++g) {
        // rolling rows need chunk c of every tile in the same stage: stage = c, one use per tile
        const int stage = fixed_stage ? c : (int)(g % NST);
        const uint32_t ph = fixed_stage ? (ptile & 1u) : ((g / NST) & 1u);
        mbar_wait_parked(&empty[stage], ph ^ 1u);  // consumers released this stage
        unsigned char* sb = smem + stage * Cfg::STAGE_BYTES;
        if (ptid == 0) {
          mbar_arrive_expect_tx(&full[stage], Cfg::P_BYTES + (Cfg::WARP ? 0 : Cfg::N_BYTES));
          tma_load_4d(sb, &tmP, &full[stage], c * KC, j0, i0, b);
          if (!Cfg::WARP) tma_load_4d(sb + Cfg::P_BYTES, &tmN, &full[stage], c * KC, j0 - D + oj, i0 - D + oi, b);
        }
        if (Cfg::WARP) {
          // units: (halo pixel, 16-byte channel quad); NQ adjacent lanes share a pixel => 32/64-byte reads
          const float* nb = nxt + (size_t)b * H * W * C + (size_t)c * KC;
          unsigned char* ns = sb + Cfg::P_BYTES;
          const int NU = nprod_px * Cfg::NQ;
          constexpr int UB = Cfg::UB;
          for (int u0 = ptid; u0 < NU; u0 += NPROD * UB) {
            TapsEntry e[UB];
            float4 v00[UB], v01[UB], v10[UB], v11[UB];
            bool live[UB];
#pragma unroll
            for (int x = 0; x < UB; ++x) {
              const int u = u0 + x * NPROD;
              live[x] = false;
              if (u < NU) {
                const int pix = u / Cfg::NQ, qd = u % Cfg::NQ;
                e[x] = taps[pix];
                live[x] = e[x].o00 >= 0 && c * KC + qd * 4 < C;
                if (live[x]) {
                  const float* base = nb + qd * 4;
                  v00[x] = __ldg(reinterpret_cast<const float4*>(base + (size_t)e[x].o00 * C));
                  v01[x] = __ldg(reinterpret_cast<const float4*>(base + (size_t)e[x].o01 * C));
                  v10[x] = __ldg(reinterpret_cast<const float4*>(base + (size_t)e[x].o10 * C));
                  v11[x] = __ldg(reinterpret_cast<const float4*>(base + (size_t)e[x].o11 * C));
                }
              }
            }
#pragma unroll
            for (int x = 0; x < UB; ++x) {
              const int u = u0 + x * NPROD;
              if (u < NU) {
                float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
                if (live[x]) {
                  Taps t;
                  t.o00 = t.o01 = t.o10 = t.o11 = 0;
                  t.w00 = e[x].w00; t.w01 = e[x].w01; t.w10 = e[x].w10; t.w11 = e[x].w11;
                  v.x = blend<Cfg::MODE>(t, v00[x].x, v01[x].x, v10[x].x, v11[x].x);
                  v.y = blend<Cfg::MODE>(t, v00[x].y, v01[x].y, v10[x].y, v11[x].y);
                  v.z = blend<Cfg::MODE>(t, v00[x].z, v01[x].z, v10[x].z, v11[x].z);
                  v.w = blend<Cfg::MODE>(t, v00[x].w, v01[x].w, v10[x].w, v11[x].w);
                }
                const int pp = u / Cfg::NQ;                                   // produced pixel (taps index)
                const int srow = (rot + rr0 + pp / NCOL) % NROW;               // its ring slot row
                *reinterpret_cast<float4*>(ns + swz<Cfg::PXB>((uint32_t)((srow * NCOL + pp % NCOL) * Cfg::PXB + (u % Cfg::NQ) * 16))) = v;
              }
            }
          }
          __syncwarp();
          if ((tid & 31) == 0) mbar_arrive(&full[stage]);
        }
      }
      ++ptile;
    QPWC_FOR_TILES_END
  } else {
    // ========================================================================== consumers
    setmaxnreg_inc<Cfg::REG_CONS>();
    // warp w owns columns (w / TH) * 32 .. +31 of tile row w % TH: the two half-row warps of a row
    // sit on the same SM sub-partition (w and w + TH, TH even), so in a ragged tile whose right half
    // is dead (see dead_warp below) every sub-partition keeps exactly one live warp
    const int lane = tid & 31, ti = (tid >> 5) % TH, tc = ((tid >> 5) / TH) * 32 + lane;
    // byte offsets inside a stage (channel quad 0; quad q = offset ^ (q << 4), see swz<>)
    // column part of the second-frame operand offset (the swizzle depends on the column only: the
    // row pitch is a multiple of the swizzle period); the row part is brow[m], per tile
    const uint32_t nb_off = Cfg::P_BYTES + swz<Cfg::PXB>((uint32_t)(((Cfg::WARP ? 0 : ti) * NCOL + tc) * Cfg::PXB));
    uint32_t a_off[Q];
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      // first-frame pixel column of acc[.][k] is lp = tc - k; columns outside the tile belong to
      // accumulators that are never stored, so they may read any resident pixel: clamp
      const int lp = min(max(tc - k, 0), TW - 1);
      a_off[k] = swz<Cfg::PXB>((uint32_t)((ti * PCOL + lp) * Cfg::PXB));
    }
    const float inv_c = 1.f / (float)C;
    float* slot0 = reinterpret_cast<float*>(smem + Cfg::OFF_STAGING + ti * Cfg::NSLOT * Cfg::SLOT_BYTES);  // private to this row
    const bool leader = (tc == 0);

    uint32_t g = 0, tcount = 0, ctile = 0;
    QPWC_FOR_TILES_BEGIN
      // second-frame rows of this thread, as byte offsets inside the stage's N area (fused variant:
      // ring slots, see the producer; plain variant: compile-time multiples of the row pitch)
      uint32_t brow[Q];
#pragma unroll
      for (int m = 0; m < Q; ++m)
        brow[m] = Cfg::WARP ? (uint32_t)(((tseg * TH) % NROW + ti + m) % NROW) * (NCOL * Cfg::PXB)
                            : (uint32_t)(m * (NCOL * Cfg::PXB));   // compile-time: folds into the load

      float acc[Q][Q];
#pragma unroll
      for (int m = 0; m < Q; ++m)
#pragma unroll
        for (int k = 0; k < Q; ++k) acc[m][k] = 0.f;

      // pixel columns of this warp's accumulators: lp = tc - k in [(tc & ~31) - 8, (tc | 31)]
      const bool dead_warp = (tc & ~31) - (Q - 1) >= min(TW, W - j0);
      for (int c = 0; c < nchunks; ++c, ++g) {
        const int stage = fixed_stage ? c : (int)(g % NST);
        if (!(ablate & 4)) mbar_wait(&full[stage], fixed_stage ? (ctile & 1u) : ((g / NST) & 1u));
        const unsigned char* sb = smem + stage * Cfg::STAGE_BYTES;
        // a warp none of whose columns reaches a valid pixel of a ragged tile (e.g. W = 512: the last
        // tile column is 8 pixels wide, so columns 32..63 feed nothing) only keeps the pipeline going
        if (!(ablate & 1) && !dead_warp)
#pragma unroll
        for (int qd = 0; qd < Cfg::NQ; ++qd) {
          const unsigned char* nbp = sb + (nb_off ^ (uint32_t)(qd << 4));
          float4 bv[Q];
#pragma unroll
          for (int m = 0; m < Q; ++m) bv[m] = *reinterpret_cast<const float4*>(nbp + brow[m]);
#pragma unroll
          for (int k = 0; k < Q; ++k) {
            const float4 a = *reinterpret_cast<const float4*>(sb + (a_off[k] ^ (uint32_t)(qd << 4)));
#pragma unroll
            for (int m = 0; m < Q; ++m) {
              acc[m][k] = fmaf(a.x, bv[m].x, acc[m][k]);
              acc[m][k] = fmaf(a.y, bv[m].y, acc[m][k]);
              acc[m][k] = fmaf(a.z, bv[m].z, acc[m][k]);
              acc[m][k] = fmaf(a.w, bv[m].w, acc[m][k]);
            }
          }
        }
        __syncwarp();
        if (lane == 0 && !(ablate & 4)) mbar_arrive(&empty[stage]);
      }
      ++ctile;
      if (ablate & 2) continue;  // dev ablation: no epilogue

      // -------------------------------------------------------------------------- epilogue
      // Each tile row (2 warps, named barrier 1+ti) stages its 56 px x 81 outputs in two halves of
      // 28 px through private slots and hands each contiguous 28 x 324-byte run to the TMA engine
      // (cp.async.bulk shared->global): no copy loop, rows never wait for each other, and the
      // store drains while the next tile is being computed.
#define QPWC_ACC(m, k) acc[m][k]
      const int twv = min(TW, W - j0);  // valid pixel columns of this tile
      if (Cfg::AGENT) {
        // scale + leaky relu of all 81 outputs as straight-line code (the per-k range checks below
        // would otherwise split it into nine short dependent chains), then deposit the 9x9 block
        // (transposed into the NHWC 81-vector order) and carry on
        float res[Q][Q];
#pragma unroll
        for (int k = 0; k < Q; ++k)
#pragma unroll
          for (int m = 0; m < Q; ++m) res[m][k] = lrelu(QPWC_ACC(m, k) * inv_c, slope);
        mbar_wait(&sfree[ti], (tcount & 1u) ^ 1u);  // previous tile's store has drained this slot
#pragma unroll
        for (int k = 0; k < Q; ++k) {
          const int lp = tc - k;  // local pixel column of acc[.][k]
          if (lp >= 0 && lp < twv) {
            float* dstp = slot0 + lp * NDISP + k;
#pragma unroll
            for (int m = 0; m < Q; ++m) dstp[m * Q] = res[m][k];
          }
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(&sfull[ti]);
        ++tcount;
        continue;
      }
      const int i = i0 + ti;
#pragma unroll
      for (int half = 0; half < 2; ++half) {        const int lp0 = half * HALF;
        const int lp1 = min(lp0 + HALF, twv);
        float* slot = slot0 + (Cfg::NSLOT == 2 ? half : 0) * (Cfg::SLOT_BYTES / 4);
        // the bulk store that last read this slot (NSLOT==2: a tile ago; else the previous half)
        if (leader) { if (Cfg::NSLOT == 2) bulk_wait_read<1>(); else bulk_wait_read<0>(); }
        named_bar_sync(1 + ti, NCOL);
#pragma unroll
        for (int k = 0; k < Q; ++k) {
          const int lp = tc - k;  // local pixel column of acc[.][k]
          if (lp >= lp0 && lp < lp1) {
            float* dstp = slot + (lp - lp0) * NDISP + k;
#pragma unroll
            for (int m = 0; m < Q; ++m) dstp[m * Q] = lrelu(QPWC_ACC(m, k) * inv_c, slope);
          }
        }
        fence_proxy_async();
        named_bar_sync(1 + ti, NCOL);  // half-slot complete
        if (i < H && lp1 > lp0) {
          float* dst = out + ((size_t)((size_t)b * H + i) * W + j0 + lp0) * (size_t)ops;
          const int n = (lp1 - lp0) * NDISP;
          if (nwin == 1 && ops == NDISP && (n & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
            if (leader) bulk_store(dst, slot, (uint32_t)n * 4u);
          } else if (nwin == 1 && ops == NDISP) {  // unaligned: plain coalesced copy by the row's 64 threads
            for (int e = tc; e < n; e += NCOL) dst[e] = slot[e];
          } else {  // strided output and/or a window of a wider search range
            const int qo = 2 * dsearch + 1;
            const int chb = (oi - D + dsearch) * qo + (oj - D + dsearch);
            // lanes own channels (their (m,k) -> output channel map is loop invariant), pixels are walked
            // in order: coalesced runs per pixel, no per-element divisions
            for (int ch = tc; ch < NDISP; ch += NCOL) {
              const int m = ch / Q, k = ch - m * Q, och = chb + m * qo + k;
              for (int px = 0; px < n / NDISP; ++px) dst[(size_t)px * ops + och] = slot[px * NDISP + ch];
            }
          }
        }
        if (leader) bulk_commit();  // one group per half, also when empty: keeps wait_group counts uniform
      }
    QPWC_FOR_TILES_END
    if (!Cfg::AGENT && leader) bulk_wait_read<0>();  // shared memory must outlive the engine's reads
  }
}

// -------------------------------------------------------------------------------------- host
#ifndef QPWC_EMU
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

// cuTensorMapEncodeTiled is a driver call and needs a current context; a thread that has only had
// cudaSetDevice applied (e.g. an autograd worker running a backward pass) has none until its first
// runtime call that touches the device.  cudaFree(nullptr) is that call.
// Only when no context is current: cudaFree is not allowed while a stream is being captured, and a
// capturing thread always has its context bound already.
static void bind_primary_context() {
  typedef CUresult (*CtxGetCurrentFn)(CUcontext*);
  static CtxGetCurrentFn get_current = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuCtxGetCurrent", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      get_current = reinterpret_cast<CtxGetCurrentFn>(p);
  }
  CUcontext ctx = nullptr;
  if (get_current && get_current(&ctx) == CUDA_SUCCESS && ctx != nullptr) return;
  (void)cudaFree(nullptr);
}

// Encoded descriptors are cached per calling thread (keyed by base pointer, shape and box): a layer
// called in a loop with the same tensors pays the two driver calls once, and there is no shared state.
struct TmapKey { const void* base; int d[7]; };
struct TmapEntry { TmapKey key; TensorMap tm; bool valid; };
static thread_local TmapEntry g_tmap_cache[32];   // a pyramid step encodes 15 distinct maps (3 per cost volume)
static thread_local unsigned g_tmap_next = 0;
static bool tmap_lookup(const TmapKey& k, TensorMap* tm, int) {
  for (int e = 0; e < 32; ++e)
    if (g_tmap_cache[e].valid && memcmp(&g_tmap_cache[e].key, &k, sizeof(k)) == 0) { *tm = g_tmap_cache[e].tm; return true; }
  return false;
}
static void tmap_store(const TmapKey& k, const TensorMap* tm, int) {
  TmapEntry& slot = g_tmap_cache[g_tmap_next++ & 31u];
  slot.key = k; slot.tm = *tm; slot.valid = true;
}

bool make_tmap_nhwc(TensorMap* tm, const float* base, int B, int H, int W, int C, int boxC, int boxW, int boxH) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = base; key.d[0] = B; key.d[1] = H; key.d[2] = W; key.d[3] = C; key.d[4] = boxC; key.d[5] = boxW; key.d[6] = boxH;
  if (tmap_lookup(key, tm, 0)) return true;
  EncodeTiledFn enc = get_encode_fn();
  bind_primary_context();
  if (!enc) { set_error(QPWC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available"); return false; }
  const cuuint64_t gdim[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstr[3] = {(cuuint64_t)C * 4, (cuuint64_t)W * C * 4, (cuuint64_t)H * W * C * 4};
  const cuuint32_t box[4] = {(cuuint32_t)boxC, (cuuint32_t)boxW, (cuuint32_t)boxH, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, boxC * 4 == 32 ? CU_TENSOR_MAP_SWIZZLE_32B : (boxC * 4 == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_128B),
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error(QPWC_ERR_CUDA, "cuTensorMapEncodeTiled failed (CUresult %d)", (int)r); return false; }
  tmap_store(key, tm, 0);
  return true;
}
bool make_tmap_cv_tiles(TensorMap* tm, float* out, int B, int H, int W, int th) {
  TmapKey key;
  memset(&key, 0, sizeof(key));
  key.base = out; key.d[0] = B; key.d[1] = H; key.d[2] = W; key.d[3] = -81; key.d[4] = 216; key.d[5] = 6; key.d[6] = th;
  if (tmap_lookup(key, tm, 0)) return true;
  EncodeTiledFn enc = get_encode_fn();
  bind_primary_context();
  if (!enc) { set_error(QPWC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available"); return false; }
  const cuuint64_t gdim[4] = {216, (cuuint64_t)W * 81 / 216, (cuuint64_t)H, (cuuint64_t)B};
  const cuuint64_t gstr[3] = {216 * 4, (cuuint64_t)W * 81 * 4, (cuuint64_t)H * W * 81 * 4};
  const cuuint32_t box[4] = {216, 6, (cuuint32_t)th, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, out, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error(QPWC_ERR_CUDA, "cuTensorMapEncodeTiled (cost-volume tiles) failed (CUresult %d)", (int)r); return false; }
  tmap_store(key, tm, 0);
  return true;
}
bool make_tmap_nchw(TensorMap* tm, const float* base, int B, int C, int H, int W, int boxW, int boxH, int boxC) {
  EncodeTiledFn enc = get_encode_fn();
  bind_primary_context();
  if (!enc) { set_error(QPWC_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available"); return false; }
  const cuuint64_t gdim[4] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)C, (cuuint64_t)B};
  const cuuint64_t gstr[3] = {(cuuint64_t)W * 4, (cuuint64_t)H * W * 4, (cuuint64_t)C * H * W * 4};
  const cuuint32_t box[4] = {(cuuint32_t)boxW, (cuuint32_t)boxH, (cuuint32_t)boxC, 1};
  const cuuint32_t estr[4] = {1, 1, 1, 1};
  const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, const_cast<float*>(base), gdim, gstr, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error(QPWC_ERR_CUDA, "cuTensorMapEncodeTiled (NCHW) failed (CUresult %d)", (int)r); return false; }
  return true;
}
static int sm_count() {
  static std::atomic<int> n[64];  // per device ordinal: a process may drive differently sized parts
  int dev = 0;
  cudaGetDevice(&dev);
  int v = n[dev & 63].load(std::memory_order_relaxed);
  if (!v) {
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev & 63].store(v, std::memory_order_relaxed);
  }
  return v;
}
#else
static int sm_count() { return 3; }  // small persistent grid: exercises the multi-tile loop
#endif

int sm_count_cached() { return sm_count(); }

// QPWC_ABLATE (dev only): bit0 skip FFMA loop, bit1 skip epilogue, bit2 skip loads+pipeline
static int ablate_flags() {
  static const int v = [] { const char* e = getenv("QPWC_ABLATE"); return e ? atoi(e) : 0; }();  // read once, thread-safe
  return v;
}

template <class Cfg>
static int run_tiled(const float* prv, const float* nxt, const float* flow, float* out, int B, int H,
                     int W, int C, float slope, long long ops, cudaStream_t stream, int dsearch = 4,
                     float up_scale = 0.f) {
  const int nwin = dsearch == 8 ? 4 : 1;
  TensorMap tmP, tmN;
  if (!make_tmap_nhwc(&tmP, prv, B, H, W, C, Cfg::KC, Cfg::PCOL, Cfg::TH)) return QPWC_ERR_CUDA;  // box 8 x 56 x TH
  if (!make_tmap_nhwc(&tmN, nxt, B, H, W, C, Cfg::KC, Cfg::NCOL, Cfg::NROW)) return QPWC_ERR_CUDA;
  const int tiles_x = cdiv(W, Cfg::TW), tiles_y = cdiv(H, Cfg::TH);
  // rolling rows (fused variant, all channels resident in the stages, single window): segments of
  // vertically consecutive tiles; short enough that there are several segments per SM
  int seg = 1;
  if (Cfg::WARP && nwin == 1 && cdiv(C, Cfg::KC) <= Cfg::NST) {
    seg = 8;
    while (seg > 2 && (long long)tiles_x * B * cdiv(tiles_y, seg) < 3LL * sm_count()) seg >>= 1;
    static const int forced = [] { const char* e = getenv("QPWC_SEG"); return e ? atoi(e) : 0; }();  // dev/tests: force the segment length
    if (forced > 0) seg = forced;
  }
  const int segs_per_strip = cdiv(tiles_y, seg);
  const long long ns = (long long)tiles_x * B * nwin * segs_per_strip;
  if (ns >= (1LL << 31)) return QPWC_ERR_UNSUPPORTED;
  const int nsegs = (int)ns;
  const int grid = nsegs < sm_count() ? nsegs : sm_count();
  auto k = corr_fwd_tiled_kernel<Cfg>;
#ifndef QPWC_EMU
  static std::atomic<unsigned> attr_done{0};  // per instantiation, one bit per device (the attribute is per device)
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_done.load(std::memory_order_acquire) >> (dev & 31) & 1u)) {
    const cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "corr_fwd_tiled: smem attribute (%d B): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
    attr_done.fetch_or(1u << (dev & 31), std::memory_order_release);
  }
#endif
  QPWC_LAUNCH(k, grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, stream, tmP, tmN, nxt, flow, out, B, H, W, C, slope, ops,
              tiles_x, tiles_y, nsegs, segs_per_strip, seg, ablate_flags(), nwin, dsearch, up_scale);
  return check_launch("corr_fwd_tiled");
}

int launch_corr_fwd_tiled(const float* prv, const float* nxt, const float* flow, int mode, float* out,
                          int B, int H, int W, int C, int d, float slope, long long ops,
                          cudaStream_t stream, float up_scale) {
  // domain: d == 4 (one 9x9 window) or d == 8 (four windows), C a multiple of 4, 16-byte aligned
  // inputs (TMA), maps at least one tile wide
  if ((d != 4 && d != 8) || (C & 3) || C < 4) return QPWC_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(prv) & 15) || (reinterpret_cast<uintptr_t>(nxt) & 15)) return QPWC_ERR_UNSUPPORTED;
  if ((long long)H * W < 64) return QPWC_ERR_UNSUPPORTED;
  if (!flow) {
    // few tiles (coarse pyramid levels): 2-row tiles double the number of busy SMs
    const long long tiles4 = (long long)cdiv(W, 56) * cdiv(H, 4) * B * (d == 8 ? 4 : 1);
    if (tiles4 * 2 <= sm_count())
      return run_tiled<TiledCfg<2, 0, QPWC_MODE_TF>>(prv, nxt, flow, out, B, H, W, C, slope, ops, stream, d);
    // scalar FFMA, 4-row tiles, 200 registers per consumer thread (measured against a packed
    // channel-parity FFMA2 consumer and a row-pair FFMA2 kernel in round 1: 178 vs 192 / 196 us at
    // 224x512x32 B=8; both variants were removed, profiles/r01_corr_variants.txt keeps the numbers)
    return run_tiled<TiledCfg<4, 0, QPWC_MODE_TF>>(prv, nxt, flow, out, B, H, W, C, slope, ops, stream, d);
  }
  if (mode == QPWC_MODE_TF) return run_tiled<TiledCfg<4, 1, QPWC_MODE_TF>>(prv, nxt, flow, out, B, H, W, C, slope, ops, stream, d, up_scale);
  return run_tiled<TiledCfg<4, 1, QPWC_MODE_TFA>>(prv, nxt, flow, out, B, H, W, C, slope, ops, stream, d, up_scale);
}

#ifdef QPWC_EMU
// the tiled backward kernel (qpwc_corr_bwd_tiled.cu) uses a 3-D grid and is not part of the CPU
// emulation build; the emulated library always takes the direct backward kernels
int launch_corr_bwd_tiled(const float*, const float*, const float*, const float*, float*, float*,
                          int, int, int, int, int, float, long long, cudaStream_t) {
  return QPWC_ERR_UNSUPPORTED;
}
#endif

}  // namespace qpwc
