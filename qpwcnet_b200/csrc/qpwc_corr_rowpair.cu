// qpwc_corr_rowpair.cu -- "row-pair" register-tiled FFMA2 cost volume for sm_100a (d = 4 window).
//
//   out[b,i,j,(di+4)*9+(dj+4)] = lrelu( (1/C) sum_c P[b,i,j,c] * N[b,i+di,j+dj,c] ),  N == 0 outside
//   (CostVolume / CostVolumeV2, qpwcnet/core/layers.py:72-100,117-132).
//
// Same decomposition idea as qpwc_corr_tiled.cu (a thread owns one second-frame column and forms an
// outer product against the first-frame pixels left/right of it), with a different packing of the
// fp32x2 FMAs.  Measured on B200 (tools/ubench/ffma2_patterns.cu, corr_loop_bench.cu):
//   * FFMA2 accepts a SCALAR multiplicand broadcast to both halves (`FFMA2 Rd, Ra.F32, Rb, Rc`),
//     so two different OUTPUTS can share one instruction:  (acc_lo, acc_hi) += n * (p_lo, p_hi).
//   * pairing the channel parity instead (round-1 kernel) doubles the accumulator registers (162
//     for 81 outputs) and its inner loop tops out at 0.60 FMA/lane/clk; the row-pair loop below
//     reaches 0.73 with 28 instead of 36 shared-memory operand words per 162 FMAs.
// A consumer thread (tp, tc) owns second-frame column s = j0-4+tc and the TWO first-frame rows
// i0+2tp, i0+2tp+1.  For N row offset mm = 0..9 (rows i0+2tp-4+mm) and first-frame pixel column
// s-(k-4):  acc2[mm-1][k] += N[mm][s] * (P[row0][s-k+4], P[row1][s-k+4])   -- .x is displacement row
// m = mm of row0, .y is displacement row m = mm-1 of row1; mm = 0 / mm = 9 only exist for row0 /
// row1 and use scalar FFMAs.  The (row0,row1) pairs must sit in adjacent registers, i.e. adjacent
// in shared memory: a repack warp turns each TMA-loaded first-frame chunk into "paired planes"
// [channel pair][row pair][pixel] x (r0c, r1c, r0c', r1c') so that one 16-byte load yields two
// ready-made register pairs (building them with MOVs makes ptxas rematerialise every pair per use).
//
// Pipeline per 8-channel chunk: TMA (N tile 16x64 px, P tile 8x56 px, SWIZZLE_32B, zero fill ==
// ZeroPadding2D) -> tfull -> repack warp -> full -> 8 consumer warps -> empty.  Epilogue: x 1/C,
// leaky relu, deposited (transposed into NHWC 81-vectors) in a full-row staging slot per row pair;
// two store-agent warps issue one asynchronous bulk store per row and release the slot.
#include <stdlib.h>

#include "qpwc_async.cuh"

namespace qpwc {

struct RowPairCfg {
  static constexpr int D = 4, Q = 9, NDISP = 81;
  static constexpr int TH = 8, NRP = TH / 2, TWT = 64, TW = TWT - 2 * D;
  static constexpr int KC = 8, PXB = KC * 4;
  static constexpr int NROW = TH + 2 * D, NCOL = TWT, PCOL = TW;
  static constexpr int NST = 2;
  static constexpr int N_BYTES = NROW * NCOL * PXB;            // 32768
  static constexpr int RAW_BYTES = TH * PCOL * PXB;            // 14336
  static constexpr int GUARD = 128;                            // 8 pixels of slack either side of a plane
  static constexpr int PLANE_BYTES = GUARD + NRP * PCOL * 16 + GUARD;
  static constexpr int NPLANE = KC / 2;
  static constexpr int PP_BYTES = NPLANE * PLANE_BYTES;        // 15360
  static constexpr int OFF_RAW = N_BYTES, OFF_PP = N_BYTES + RAW_BYTES;
  static constexpr int STAGE_BYTES = N_BYTES + RAW_BYTES + PP_BYTES;
  // epilogue staging: one full-row slot per row pair (its two rows go through it one after the
  // other); consumers only deposit, store-agent warps hand the rows to the TMA engine
  static constexpr int SLOT_BYTES = ((TW * NDISP * 4 + 127) / 128) * 128;
  static constexpr int OFF_STAGING = NST * STAGE_BYTES;
  static constexpr int OFF_BARS = OFF_STAGING + NRP * SLOT_BYTES;
  static constexpr int SMEM_BYTES = OFF_BARS + 3 * NST * 8 + 2 * NRP * 8;
  static constexpr int NCONS = NRP * TWT, NPROD = 128, NTHREADS = NCONS + NPROD;
  static constexpr int REG_CONS = 224, REG_PROD = 56;
  static constexpr int NREPACK = 2;
  static_assert(NRP == 4, "the store agent releases four slots per row");  // repack warps (producer warps 1, 2); warp 3 is the store agent
  static_assert(SMEM_BYTES <= 232448, "shared memory budget (227 KB)");
  static_assert(STAGE_BYTES % 512 == 0 && OFF_RAW % 512 == 0 && OFF_PP % 128 == 0, "stage alignment");
  static_assert(NCONS * REG_CONS + NPROD * REG_PROD <= 65536, "register budget");
  static_assert((NRP * PCOL * (KC / 4)) % 64 == 0, "repack items per warp pass");
};

#define QPWC_RP_FOR_TILES_BEGIN                                                                   \
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {                                 \
    const int tx = tile % tiles_x, rest_ = tile / tiles_x;                                        \
    const int ty = rest_ % tiles_y, bw = rest_ / tiles_y;                                         \
    const int b = bw / nwin, win = bw - b * nwin;                                                 \
    const int oi = nwin == 1 ? 0 : ((win >> 1) * 8 - 4), oj = nwin == 1 ? 0 : ((win & 1) * 8 - 4); \
    const int i0 = ty * Cfg::TH, j0 = tx * Cfg::TW;                                               \
    (void)oi; (void)oj; (void)b; (void)i0; (void)j0;
#define QPWC_RP_FOR_TILES_END }

__global__ void __launch_bounds__(RowPairCfg::NTHREADS, 1)
corr_fwd_rowpair_kernel(const QPWC_GRID_CONSTANT TensorMap tmP, const QPWC_GRID_CONSTANT TensorMap tmN,
                        float* __restrict__ out, int B, int H, int W, int C, float slope, long long ops,
                        int tiles_x, int tiles_y, int ntiles, int nwin, int dsearch, int ablate) {
  using Cfg = RowPairCfg;
  constexpr int D = Cfg::D, Q = Cfg::Q, NDISP = Cfg::NDISP, TW = Cfg::TW;
  constexpr int NCOL = Cfg::NCOL, PCOL = Cfg::PCOL, NST = Cfg::NST, KC = Cfg::KC, PXB = Cfg::PXB;
  constexpr int NCONS = Cfg::NCONS;

  QPWC_DYN_SMEM(smem);
  uint64_t* tfull = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);  // TMA landed (raw P + N)
  uint64_t* full = tfull + NST;                                         // paired planes written
  uint64_t* empty = full + NST;                                         // consumers done with the stage
  uint64_t* sfull = empty + NST;        // per row pair: both consumer warps deposited a row
  uint64_t* sfree = sfull + Cfg::NRP;   // per row pair: the row's store has drained the slot

  const int tid = threadIdx.x;
  const int nchunks = (C + KC - 1) / KC;

  if (tid == 0) {
#ifndef QPWC_EMU
    if (smem_u32(smem) & 511u) __trap();
#endif
    for (int s = 0; s < NST; ++s) { mbar_init(&tfull[s], 1); mbar_init(&full[s], Cfg::NREPACK); mbar_init(&empty[s], NCONS / 32); }
    for (int r = 0; r < Cfg::NRP; ++r) { mbar_init(&sfull[r], NCOL / 32); mbar_init(&sfree[r], 1); }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= NCONS) {
    // ========================================================================== producers
    setmaxnreg_dec<Cfg::REG_PROD>();
    const int ptid = tid - NCONS, pw = ptid >> 5, lane = ptid & 31;
    if (pw == 0) {
      // ---- TMA issue (one thread)
      if (lane != 0) return;
      tma_prefetch_desc(&tmP); tma_prefetch_desc(&tmN);
      uint32_t g = 0;
      QPWC_RP_FOR_TILES_BEGIN
        for (int c = 0; c < nchunks; ++c, ++g) {
          const int stage = (int)(g % NST);
          mbar_wait_parked(&empty[stage], ((g / NST) & 1u) ^ 1u);
          unsigned char* sb = smem + stage * Cfg::STAGE_BYTES;
          if (ablate & 4) { mbar_arrive(&tfull[stage]); continue; }  // dev ablation: no loads
          mbar_arrive_expect_tx(&tfull[stage], Cfg::RAW_BYTES + Cfg::N_BYTES);
          tma_load_4d(sb + Cfg::OFF_RAW, &tmP, &tfull[stage], c * KC, j0, i0, b);
          tma_load_4d(sb, &tmN, &tfull[stage], c * KC, j0 - D + oj, i0 - D + oi, b);
        }
      QPWC_RP_FOR_TILES_END
    } else if (pw <= Cfg::NREPACK) {
      // ---- repack: raw first-frame chunk [row][px][8 ch] -> paired planes
      //      plane pl (channels 2pl, 2pl+1): [row pair][px] x (r0 c, r1 c, r0 c+1, r1 c+1)
      // items (row pair, pixel, channel quad), PER per lane; their offsets are loop invariant
      constexpr int NITEM = Cfg::NRP * PCOL * (KC / 4), PER = NITEM / (32 * Cfg::NREPACK);
      static_assert(NITEM % (32 * Cfg::NREPACK) == 0 && (PCOL * PXB) % 256 == 0, "repack items per lane");
      uint32_t src_off[PER], dst_off[PER];
#pragma unroll
      for (int j = 0; j < PER; ++j) {
        const int it = (j * Cfg::NREPACK + (pw - 1)) * 32 + lane;
        const int grp = it / PCOL, px = it - grp * PCOL, rp = grp >> 1, qd = grp & 1;
        src_off[j] = swz32((uint32_t)(((2 * rp) * PCOL + px) * PXB)) ^ (uint32_t)(qd << 4);
        dst_off[j] = (uint32_t)((2 * qd) * Cfg::PLANE_BYTES + (rp * PCOL + px) * 16);
      }
      uint32_t g = 0;
      QPWC_RP_FOR_TILES_BEGIN
        for (int c = 0; c < nchunks; ++c, ++g) {
          const int stage = (int)(g % NST);
          mbar_wait_parked(&tfull[stage], (g / NST) & 1u);  // implies the stage was released by the consumers
          unsigned char* sb = smem + stage * Cfg::STAGE_BYTES;
          const unsigned char* raw = sb + Cfg::OFF_RAW;
          unsigned char* pp = sb + Cfg::OFF_PP + Cfg::GUARD;
          if (!(ablate & 2)) {  // (dev ablation bit 1: no repack)
            // loads batched ahead of the stores: the shared-memory latency is paid once per batch
            constexpr int BATCH = 4;
#pragma unroll
            for (int j0_ = 0; j0_ < PER; j0_ += BATCH) {
              float4 r0[BATCH], r1[BATCH];
#pragma unroll
              for (int x = 0; x < BATCH; ++x) if (j0_ + x < PER) {
                r0[x] = *reinterpret_cast<const float4*>(raw + src_off[j0_ + x]);
                r1[x] = *reinterpret_cast<const float4*>(raw + src_off[j0_ + x] + PCOL * PXB);  // next row: pitch % 256 == 0
              }
#pragma unroll
              for (int x = 0; x < BATCH; ++x) if (j0_ + x < PER) {
                unsigned char* dst = pp + dst_off[j0_ + x];
                *reinterpret_cast<float4*>(dst) = make_float4(r0[x].x, r1[x].x, r0[x].y, r1[x].y);
                *reinterpret_cast<float4*>(dst + Cfg::PLANE_BYTES) = make_float4(r0[x].z, r1[x].z, r0[x].w, r1[x].w);
              }
            }
          }
          __syncwarp();
          if (lane == 0) mbar_arrive(&full[stage]);
        }
      QPWC_RP_FOR_TILES_END
    } else {
      // ---- store agent (warp 3).  Per tile and per row of each pair: wait for the deposit, issue
      //      the bulk store of the row, release the slot once the engine has read it.  Consumers
      //      never wait for a store to reach memory.
      const int aw = 0;
      uint32_t n = 0;
      if (ablate & 8) return;
      QPWC_RP_FOR_TILES_BEGIN
        const int twv = min(TW, W - j0);
        for (int rr = 0; rr < 2; ++rr, ++n) {
          for (int rp = aw; rp < Cfg::NRP; ++rp) {
            mbar_wait_parked(&sfull[rp], n & 1u);
            const int i = i0 + 2 * rp + rr;
            if (i < H && !(ablate & 1)) {  // (dev ablation bit 0: no stores)
              const float* slot = reinterpret_cast<const float*>(smem + Cfg::OFF_STAGING + rp * Cfg::SLOT_BYTES);
              float* dst = out + ((size_t)((size_t)b * H + i) * W + j0) * (size_t)ops;
              const int cnt = twv * NDISP;
              if (nwin == 1 && ops == NDISP && (cnt & 3) == 0 && (reinterpret_cast<uintptr_t>(dst) & 15) == 0) {
                if (lane == 0) bulk_store(dst, slot, (uint32_t)cnt * 4u);
              } else if (nwin == 1 && ops == NDISP) {  // unaligned row: plain coalesced copy
                for (int e = lane; e < cnt; e += 32) dst[e] = slot[e];
              } else {  // strided output (concat buffer) and/or one window of a wider search range
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
            if (lane == 0) bulk_commit();  // one group per row pair (also when nothing was issued)
          }
          // groups complete in order: release every slot as soon as the engine has read it
          __syncwarp();
          if (lane == 0) {
            bulk_wait_read<3>(); mbar_arrive(&sfree[0]);
            bulk_wait_read<2>(); mbar_arrive(&sfree[1]);
            bulk_wait_read<1>(); mbar_arrive(&sfree[2]);
            bulk_wait_read<0>(); mbar_arrive(&sfree[3]);
          }
        }
      QPWC_RP_FOR_TILES_END
    }
    return;
  }

  // ============================================================================ consumers
  setmaxnreg_inc<Cfg::REG_CONS>();
  const int tp = tid / NCOL, tc = tid % NCOL, lane = tid & 31;
  // operand byte offsets inside a stage.  P: plane 2qd(+1), pixel column tc-k of row pair tp -- the
  // columns outside [0,56) belong to accumulators that are never stored and read the guard / the
  // neighbouring row pair.  N: row 2tp+mm, column tc (row pitch is a multiple of the swizzle period)
  const uint32_t p_off = Cfg::OFF_PP + Cfg::GUARD + (uint32_t)((tp * PCOL + tc) * 16);
  const uint32_t nb_off = swz32((uint32_t)((2 * tp * NCOL + tc) * PXB));
  const float inv_c = 1.f / (float)C;
  float* slot = reinterpret_cast<float*>(smem + Cfg::OFF_STAGING + tp * Cfg::SLOT_BYTES);  // private to this row pair

  uint32_t g = 0, dep = 0;
  QPWC_RP_FOR_TILES_BEGIN
    float2 acc2[8][Q];   // [mm-1][k]: .x = row0, displacement row mm; .y = row1, displacement row mm-1
    float accT[Q], accB[Q];  // row0 / mm = 0   and   row1 / mm = 9
#pragma unroll
    for (int k = 0; k < Q; ++k) {
      accT[k] = accB[k] = 0.f;
#pragma unroll
      for (int m = 0; m < 8; ++m) acc2[m][k] = make_float2(0.f, 0.f);
    }

    for (int c = 0; c < nchunks; ++c, ++g) {
      const int stage = (int)(g % NST);
      mbar_wait(&full[stage], (g / NST) & 1u);
      const unsigned char* sb = smem + stage * Cfg::STAGE_BYTES;
#define QPWC_RP_STEP(NX, PX)                                                                  \
      accT[k] = fmaf(n[0].NX, PX.x, accT[k]);                                                 \
      _Pragma("unroll") for (int m = 0; m < 8; ++m)                                           \
        acc2[m][k] = __ffma2_rn(make_float2(n[m + 1].NX, n[m + 1].NX), PX, acc2[m][k]);       \
      accB[k] = fmaf(n[9].NX, PX.y, accB[k]);
      // one channel quad per iteration; kept rolled: with both quads in one body ptxas renames the
      // accumulators and undoes the permutation with ~300 MOVs per chunk
      // One channel PAIR (plane hq) per block of 9 x (8 FFMA2 + 2 FFMA) x 2.  N is read per channel
      // pair as 8-byte loads (2-way bank conflict on the 32-byte-swizzled layout == the wavefronts of
      // the 16-byte load, but 20 instead of 40 live operand registers: at 40 ptxas spills inside the
      // loop and shuffles accumulators with MOVs).  The loads are opaque (asm volatile) so that ptxas
      // neither merges them back into 16-byte loads nor sinks them: the N operands of the next
      // channel pair are fetched in the middle of the current block (ping-pong nA / nB).
      const uint32_t nbase = QPWC_SMEM_ADDR(sb) + nb_off;
#define QPWC_RP_LOADN(dst, hq)                                                                  \
      _Pragma("unroll") for (int mm = 0; mm < 10; ++mm)                                         \
        dst[mm] = lds_f2((nbase ^ (uint32_t)(((hq) >> 1) << 4)) + (uint32_t)(((hq) & 1) * 8 + mm * (NCOL * PXB)));
#define QPWC_RP_BLOCK(n, hq, PREFETCH)                                                          \
      {                                                                                         \
        const unsigned char* pbp = sb + p_off + (hq) * Cfg::PLANE_BYTES;                        \
        _Pragma("unroll") for (int k = 0; k < Q; ++k) {                                         \
          const float4 pa = *reinterpret_cast<const float4*>(pbp - k * 16);                     \
          const float2 p0 = make_float2(pa.x, pa.y), p1 = make_float2(pa.z, pa.w);              \
          if (k == 3) { PREFETCH }                                                              \
          QPWC_RP_STEP(x, p0) QPWC_RP_STEP(y, p1)                                               \
        }                                                                                       \
      }
      float2 nA[10], nB[10];
      QPWC_RP_LOADN(nA, 0)
#pragma unroll 1
      for (int hq = 0; hq < KC / 2; hq += 2) {
#define n nA
        QPWC_RP_BLOCK(nA, hq, QPWC_RP_LOADN(nB, hq + 1))
#undef n
#define n nB
        QPWC_RP_BLOCK(nB, hq + 1, if (hq + 2 < KC / 2) { QPWC_RP_LOADN(nA, hq + 2) })
#undef n
      }
#undef QPWC_RP_BLOCK
#undef QPWC_RP_LOADN
#undef QPWC_RP_STEP
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }

    // ---------------------------------------------------------------------------- epilogue
    if (ablate & 8) continue;  // dev ablation: no epilogue at all
    // deposit row0 and row1 of the pair through the pair's slot (9x9 blocks transposed into the
    // NHWC 81-vector order); the store agent does the rest
    const int twv = min(TW, W - j0);  // valid pixel columns of this tile
#pragma unroll
    for (int rr = 0; rr < 2; ++rr, ++dep) {
      // scale + leaky relu of this row's 81 outputs (row1's run while row0's store drains)
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        if (rr == 0) accT[k] = lrelu(accT[k] * inv_c, slope); else accB[k] = lrelu(accB[k] * inv_c, slope);
#pragma unroll
        for (int m = 0; m < 8; ++m) {
          if (rr == 0) acc2[m][k].x = lrelu(acc2[m][k].x * inv_c, slope);
          else acc2[m][k].y = lrelu(acc2[m][k].y * inv_c, slope);
        }
      }
      mbar_wait(&sfree[tp], (dep & 1u) ^ 1u);  // the previous row's store has drained the slot
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        const int lp = tc - k;  // local pixel column of the accumulators [.][k]
        if (lp >= 0 && lp < twv) {
          float* dstp = slot + (tc - k) * NDISP + k;
          if (rr == 0) {
            dstp[0] = accT[k];
#pragma unroll
            for (int m = 1; m < Q; ++m) dstp[m * Q] = acc2[m - 1][k].x;
          } else {
#pragma unroll
            for (int m = 0; m < Q - 1; ++m) dstp[m * Q] = acc2[m][k].y;
            dstp[(Q - 1) * Q] = accB[k];
          }
        }
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(&sfull[tp]);
    }
  QPWC_RP_FOR_TILES_END
}

// -------------------------------------------------------------------------------------- host
int sm_count_cached();  // qpwc_corr_tiled.cu

// QPWC_ABLATE_RP (dev only): bit0 no stores, bit1 no repack, bit2 no loads, bit3 no epilogue
static int ablate_flags_rp() {
  const char* e = getenv("QPWC_ABLATE_RP");
  return e ? atoi(e) : 0;
}

int launch_corr_fwd_rowpair(const float* prv, const float* nxt, float* out, int B, int H, int W, int C,
                            int d, float slope, long long ops, cudaStream_t stream) {
  using Cfg = RowPairCfg;
  const int nwin = d == 8 ? 4 : 1;
  TensorMap tmP, tmN;
  if (!make_tmap_nhwc(&tmP, prv, B, H, W, C, Cfg::KC, Cfg::PCOL, Cfg::TH)) return QPWC_ERR_CUDA;
  if (!make_tmap_nhwc(&tmN, nxt, B, H, W, C, Cfg::KC, Cfg::NCOL, Cfg::NROW)) return QPWC_ERR_CUDA;
  const int tiles_x = cdiv(W, Cfg::TW), tiles_y = cdiv(H, Cfg::TH);
  const long long nt = (long long)tiles_x * tiles_y * B * nwin;
  if (nt >= (1LL << 31)) return QPWC_ERR_UNSUPPORTED;
  const int ntiles = (int)nt;
  const int grid = ntiles < sm_count_cached() ? ntiles : sm_count_cached();
  auto k = corr_fwd_rowpair_kernel;
#ifndef QPWC_EMU
  static unsigned attr_done = 0;  // one bit per device (the attribute is per device)
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_done >> (dev & 31) & 1u)) {
    const cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "corr_fwd_rowpair: smem attribute (%d B): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
    attr_done |= 1u << (dev & 31);
  }
#endif
  QPWC_LAUNCH(k, grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, stream, tmP, tmN, out, B, H, W, C, slope, ops,
              tiles_x, tiles_y, ntiles, nwin, d, ablate_flags_rp());
  return check_launch("corr_fwd_rowpair");
}

}  // namespace qpwc
