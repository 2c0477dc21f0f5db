// qpwc_corr_nchw.cu -- native channels_first (NCHW) cost volume for sm_100a, d = 4.
//
//   out[b,(di+4)*9+(dj+4),i,j] = lrelu( (1/C) sum_c P[b,c,i,j] * N[b,c,i+di,j+dj] ),  N == 0 outside
//   (CostVolume / CostVolumeV2 with data_format='channels_first', qpwcnet/core/layers.py:72-100 with
//   the axis handling of layers.py:83-85 -- the reference's training layout, pre_train.py:34.)
//
// In NCHW a TMA tile is channel-planar ([channel][row][column], columns contiguous), so ADJACENT
// PIXELS are adjacent in shared memory.  That is exactly what the scalar-broadcast FFMA2 wants
// (tools/ubench/ffma2_patterns.cu): a thread owns the aligned second-frame column pair
// (s, s+1) = (S, S+1), S = j0-4+2t, and the ten first-frame pixels p = S-4 .. S+5 of its row;
//   acc2[q][m] += P[c][i][p_q] * (N[c][i+m-4][S], N[c][i+m-4][S+1])        q = 0..9, m = 0..8
// is one FFMA2 whose pair operand is a natural 8-byte load and whose scalar is a component of
// another: .x is displacement column dj = 4-q of pixel p_q, .y is dj = 5-q.  q = 0 / q = 9 only have
// a valid .x / .y half and use scalar FFMAs, so 162 accumulator registers hold 162 outputs, with 14
// 8-byte loads and 72 FFMA2 + 18 FFMA per channel (the NHWC scalar kernel: 81 FFMA and the
// equivalent of 18 loads for half as many outputs).  No repack (cf. qpwc_corr_rowpair.cu), no
// swizzle (a warp's 8-byte loads are 256 contiguous bytes), 120 of 128 columns valid.
// The output is NCHW too: plane (m,k) receives two adjacent pixels from every thread, i.e. a warp
// writes 256 contiguous bytes per store -- straight from registers, no staging, no store agents.
#include <atomic>
#include <stdlib.h>

#include "qpwc_async.cuh"

namespace qpwc {

template <int TH_>
struct NchwCfg {
  static constexpr int D = 4, Q = 9, NDISP = 81;
  static constexpr int TH = TH_, NCOLS = 128, TW = NCOLS - 2 * D;   // 120 valid pixel columns per tile
  static constexpr int PCOLS = NCOLS + 2 * D;                     // first-frame columns j0-8 .. j0+127
  static constexpr int KC = 8, NROW = TH + 2 * D;
  static constexpr int NST = TH_ <= 2 ? 4 : 3;
  static constexpr int N_BYTES = KC * NROW * NCOLS * 4;           // TH = 4: 49152
  static constexpr int P_BYTES = KC * TH * PCOLS * 4;             //         17408
  static constexpr int STAGE_BYTES = N_BYTES + P_BYTES;           //         66560
  static constexpr int OFF_BARS = NST * STAGE_BYTES;
  static constexpr int SMEM_BYTES = OFF_BARS + 2 * NST * 8;
  static constexpr int NCONS = TH * (NCOLS / 2), NPROD = 128, NTHREADS = NCONS + NPROD;
  static constexpr int REG_CONS = 232, REG_PROD = 32;
  static_assert(SMEM_BYTES <= 232448, "shared memory budget (227 KB)");
  static_assert(N_BYTES % 128 == 0 && STAGE_BYTES % 128 == 0, "TMA destination alignment");
  static_assert(NCONS * REG_CONS + NPROD * REG_PROD <= 65536, "register budget");
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::NTHREADS, 1)
corr_fwd_nchw_kernel(const QPWC_GRID_CONSTANT TensorMap tmP, const QPWC_GRID_CONSTANT TensorMap tmN,
                     float* __restrict__ out, int B, int H, int W, int C, float slope,
                     int tiles_x, int tiles_y, int ntiles) {
  constexpr int D = Cfg::D, Q = Cfg::Q, TH = Cfg::TH, TW = Cfg::TW, NST = Cfg::NST, KC = Cfg::KC;
  constexpr int NCOLS = Cfg::NCOLS, PCOLS = Cfg::PCOLS, NROW = Cfg::NROW, NCONS = Cfg::NCONS;

  QPWC_DYN_SMEM(smem);
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + Cfg::OFF_BARS);
  uint64_t* empty = full + NST;
  const int tid = threadIdx.x;
  const int nchunks = (C + KC - 1) / KC;

  if (tid == 0) {
    for (int s = 0; s < NST; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], NCONS / 32); }
    fence_mbar_init();
  }
  __syncthreads();

  if (tid >= NCONS) {
    // ============================================================================ TMA producer
    setmaxnreg_dec<Cfg::REG_PROD>();
    if (tid != NCONS) return;
    tma_prefetch_desc(&tmP); tma_prefetch_desc(&tmN);
    uint32_t g = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int tx = tile % tiles_x, rest = tile / tiles_x, ty = rest % tiles_y, b = rest / tiles_y;
      const int i0 = ty * TH, j0 = tx * TW;
      for (int c = 0; c < nchunks; ++c, ++g) {
        const int stage = (int)(g % NST);
        mbar_wait_parked(&empty[stage], ((g / NST) & 1u) ^ 1u);
        unsigned char* sb = smem + stage * Cfg::STAGE_BYTES;
        mbar_arrive_expect_tx(&full[stage], Cfg::STAGE_BYTES);
        tma_load_4d(sb, &tmN, &full[stage], j0 - D, i0 - D, c * KC, b);                  // [ch][12 rows][128 cols]
        tma_load_4d(sb + Cfg::N_BYTES, &tmP, &full[stage], j0 - 2 * D, i0, c * KC, b);   // [ch][4 rows][136 cols]
      }
    }
    return;
  }

  // ================================================================================== consumers
  setmaxnreg_inc<Cfg::REG_CONS>();
  const int ti = tid / (NCOLS / 2), t = tid % (NCOLS / 2), lane = tid & 31;   // column pair t of row ti
  const int tw0 = t & ~31;                                                      // first column pair of this warp
  // operand byte offsets inside one channel plane of a stage
  const uint32_t n_off = (uint32_t)((ti * NCOLS + 2 * t) * 4);                  // row ti+m: + m * NCOLS * 4
  const uint32_t p_off = (uint32_t)(Cfg::N_BYTES + (ti * PCOLS + 2 * t) * 4);   // pixel q: + q * 4
  const float inv_c = 1.f / (float)C;
  const size_t plane = (size_t)H * W;

  uint32_t g = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, rest = tile / tiles_x, ty = rest % tiles_y, b = rest / tiles_y;
    const int i0 = ty * TH, j0 = tx * TW;

    float2 acc2[8][Q];        // q = 1..8: .x = (p_q, dj = 4-q), .y = (p_q, dj = 5-q), rows m
    float accL[Q], accR[Q];   // q = 0: dj = 4 only (.x);  q = 9: dj = -4 only (.y)
#pragma unroll
    for (int m = 0; m < Q; ++m) {
      accL[m] = accR[m] = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) acc2[q][m] = make_float2(0.f, 0.f);
    }

    for (int c = 0; c < nchunks; ++c, ++g) {
      const int stage = (int)(g % NST);
      mbar_wait(&full[stage], (g / NST) & 1u);
      const unsigned char* sb = smem + stage * Cfg::STAGE_BYTES;
#pragma unroll 2
      for (int ch = 0; ch < KC; ++ch) {
        const unsigned char* np = sb + ch * (NROW * NCOLS * 4) + n_off;
        const unsigned char* pp = sb + ch * (TH * PCOLS * 4) + p_off;
        float2 n[Q], pv[5];
#pragma unroll
        for (int m = 0; m < Q; ++m) n[m] = *reinterpret_cast<const float2*>(np + m * (NCOLS * 4));
#pragma unroll
        for (int u = 0; u < 5; ++u) pv[u] = *reinterpret_cast<const float2*>(pp + u * 8);
#pragma unroll
        for (int m = 0; m < Q; ++m) {
          accL[m] = fmaf(pv[0].x, n[m].x, accL[m]);
#pragma unroll
          for (int q = 1; q <= 8; ++q) {
            const float ps = (q & 1) ? pv[q >> 1].y : pv[q >> 1].x;
            acc2[q - 1][m] = __ffma2_rn(make_float2(ps, ps), n[m], acc2[q - 1][m]);
          }
          accR[m] = fmaf(pv[4].y, n[m].y, accR[m]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }

    // ------------------------------------------------------------------------------ epilogue
    // plane (m,k), k = dj+4: this thread holds its pixels p = S+4-k (from .x of q = 8-k, or accL for
    // k = 8) and p+1 (from .y of q = 9-k, or accR for k = 0).  S is even, so p is even iff k is.
    const int i = i0 + ti;
    if (i < H) {
      const int S = j0 - D + 2 * t;
      const size_t plane9 = (size_t)Q * plane;   // m -> m+1 at fixed k
      float* base = out + (size_t)b * Cfg::NDISP * plane + (size_t)i * W;
      // A thread's outputs for pixels just outside [j0, j0+120) are COMPLETE and bit-identical to what
      // the neighbouring tile computes (same operands, same channel order), so they are stored as
      // well: only the image bounds need checking, and a warp whose pixel range [pmin, pmax] lies
      // inside the row stores without any predicate.
      const int pmin = j0 - D + 2 * tw0 - 4, pmax = j0 - D + 2 * tw0 + 67;
      if (pmin >= 0 && pmax < W) {
#pragma unroll
        for (int k = 0; k < Q; ++k) {
          float* dst = base + (size_t)k * plane + (S + 4 - k);
#pragma unroll
          for (int m = 0; m < Q; ++m) {
            const float v0 = lrelu((k == 8 ? accL[m] : acc2[k == 8 ? 0 : 7 - k][m].x) * inv_c, slope);
            const float v1 = lrelu((k == 0 ? accR[m] : acc2[k == 0 ? 0 : 8 - k][m].y) * inv_c, slope);
            if ((k & 1) == 0) *reinterpret_cast<float2*>(dst) = make_float2(v0, v1);   // p even: aligned pair
            else { dst[0] = v0; dst[1] = v1; }
            dst += plane9;
          }
        }
      } else {
        // warps that reach over the left / right image border: per-pixel bounds
#pragma unroll
        for (int k = 0; k < Q; ++k) {
          const int p = S + 4 - k;
          const bool ok0 = p >= 0 && p < W, ok1 = p + 1 >= 0 && p + 1 < W;
          float* dst = base + (size_t)k * plane + p;
          if (ok0 || ok1) {
#pragma unroll
            for (int m = 0; m < Q; ++m) {
              const float v0 = lrelu((k == 8 ? accL[m] : acc2[k == 8 ? 0 : 7 - k][m].x) * inv_c, slope);
              const float v1 = lrelu((k == 0 ? accR[m] : acc2[k == 0 ? 0 : 8 - k][m].y) * inv_c, slope);
              if (ok0) dst[0] = v0;
              if (ok1) dst[1] = v1;
              dst += plane9;
            }
          }
        }
      }
    }
  }
}

// -------------------------------------------------------------------------------------- host
int sm_count_cached();  // qpwc_corr_tiled.cu

template <class Cfg>
static int run_nchw(const float* prv, const float* nxt, float* out, int B, int C, int H, int W, float slope,
                    cudaStream_t stream) {
  TensorMap tmP, tmN;
  if (!make_tmap_nchw(&tmP, prv, B, C, H, W, Cfg::PCOLS, Cfg::TH, Cfg::KC)) return QPWC_ERR_CUDA;
  if (!make_tmap_nchw(&tmN, nxt, B, C, H, W, Cfg::NCOLS, Cfg::NROW, Cfg::KC)) return QPWC_ERR_CUDA;
  const int tiles_x = cdiv(W, Cfg::TW), tiles_y = cdiv(H, Cfg::TH);
  const long long nt = (long long)tiles_x * tiles_y * B;
  if (nt >= (1LL << 31)) return QPWC_ERR_UNSUPPORTED;
  const int ntiles = (int)nt;
  const int grid = ntiles < sm_count_cached() ? ntiles : sm_count_cached();
  auto k = corr_fwd_nchw_kernel<Cfg>;
#ifndef QPWC_EMU
  static std::atomic<unsigned> attr_done{0};  // per instantiation, one bit per device (the attribute is per device)
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_done.load(std::memory_order_acquire) >> (dev & 31) & 1u)) {
    const cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "corr_fwd_nchw: smem attribute (%d B): %s", Cfg::SMEM_BYTES, cudaGetErrorString(e));
    attr_done.fetch_or(1u << (dev & 31), std::memory_order_release);
  }
#endif
  QPWC_LAUNCH(k, grid, Cfg::NTHREADS, Cfg::SMEM_BYTES, stream, tmP, tmN, out, B, H, W, C, slope, tiles_x, tiles_y, ntiles);
  return check_launch("corr_fwd_nchw");
}

// Shape-generic channels_first kernel (any search range, any W, any alignment): one thread per output element
// (b, displacement, i, j), consecutive lanes = consecutive columns, so every plane access is a coalesced row
// segment.  It serves the shapes outside the tiled kernel's domain -- W % 4 != 0 (TMA strides are multiples of
// 16 bytes: e.g. the 8x14 level of a 256x448 training crop), search ranges other than 4 -- so that no
// channels_first shape falls back to transposing the tensors.
__global__ void __launch_bounds__(256) corr_fwd_nchw_generic_kernel(const float* __restrict__ prv, const float* __restrict__ nxt,
                                                                    float* __restrict__ out, int C, int H, int W, int d,
                                                                    float slope, long long total) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int q = 2 * d + 1, D = q * q;
  const int j = (int)(idx % W);
  long long r = idx / W;
  const int i = (int)(r % H); r /= H;
  const int e = (int)(r % D);
  const long long b = r / D;
  const int ii = i + e / q - d, jj = j + e % q - d;
  float acc = 0.f;
  if (ii >= 0 && ii < H && jj >= 0 && jj < W) {   // ZeroPadding2D: products with the padding are zero
    const size_t plane = (size_t)H * W;
    const float* p = prv + (size_t)b * C * plane + (size_t)i * W + j;
    const float* n = nxt + (size_t)b * C * plane + (size_t)ii * W + jj;
    for (int c = 0; c < C; ++c, p += plane, n += plane) acc = fmaf(__ldg(p), __ldg(n), acc);
  }
  out[idx] = lrelu(acc * (1.f / (float)C), slope);
}

int launch_corr_fwd_nchw(const float* prv, const float* nxt, float* out, int B, int C, int H, int W,
                         int d, float slope, cudaStream_t stream) {
  // tiled kernel: d == 4, W a multiple of 4 (TMA strides are multiples of 16 bytes; 8-byte output stores),
  // 16-byte aligned inputs, 8-byte aligned output; everything else takes the generic kernel
  if (d < 1 || C < 1) return QPWC_ERR_UNSUPPORTED;
  if (d != 4 || (W & 3) || (reinterpret_cast<uintptr_t>(prv) & 15) || (reinterpret_cast<uintptr_t>(nxt) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 7)) {
    const long long total = (long long)B * (2 * d + 1) * (2 * d + 1) * H * W;
    if (total == 0) return QPWC_OK;
    if (cdivll(total, 256) >= (1LL << 31)) return QPWC_ERR_UNSUPPORTED;
    QPWC_LAUNCH(corr_fwd_nchw_generic_kernel, (unsigned)cdivll(total, 256), 256, 0, stream, prv, nxt, out, C, H, W, d, slope, total);
    return check_launch("corr_fwd_nchw_generic");
  }
  // few tiles (coarse pyramid levels): 2-row tiles double the number of busy SMs
  const long long tiles4 = (long long)cdiv(W, 120) * cdiv(H, 4) * B;
  if (tiles4 * 2 <= sm_count_cached()) return run_nchw<NchwCfg<2>>(prv, nxt, out, B, C, H, W, slope, stream);
  return run_nchw<NchwCfg<4>>(prv, nxt, out, B, C, H, W, slope, stream);
}

}  // namespace qpwc
