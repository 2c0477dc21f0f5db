// qpwc_corr_bwd_nchw.cu -- gradients of the channels_first (NCHW) cost volume for sm_100a, d = 4.
//
//   forward  out[b,(di+4)*9+(dj+4),i,j] = lrelu( (1/C) sum_c P[b,c,i,j] * N[b,c,i+di,j+dj] )
//   G'[b,k,i,j] = g_out * (out > 0 ? 1 : slope) / C                        (LeakyReluGrad, mean)
//   g_prv[b,c,i,j] = sum_{di,dj} G'[b,(di,dj),i,j]       * N[b,c,i+di,j+dj]
//   g_nxt[b,c,y,x] = sum_{di,dj} G'[b,(di,dj),y-di,x-dj] * P[b,c,y-di,x-dj]
//   (autodiff of CostVolume / CostVolumeV2 under data_format='channels_first',
//   qpwcnet/core/layers.py:72-100 with the axis handling of layers.py:83-85.)
//
// Both gradients are the same 81-tap stencil over one channel plane with per-pixel coefficients:
// with m' = 4-di, k' = 4-dj the second line reads sum G'[.., y+m'-4, x+k'-4] * P[.., y+m'-4, x+k'-4].
// In NCHW the coefficients do not depend on the channel, so a thread that owns a pixel pair keeps its
// 2 x 81 coefficients in REGISTERS for the whole tile (read once from the G' planes: consecutive
// lanes are consecutive columns, every load of a plane is a coalesced row segment) and then walks
// the channels: per channel 45 8-byte shared-memory loads (9 rows x 10 columns of the haloed plane
// tile, delivered by TMA, zero-filled outside the image), 162 FFMAs, and one 8-byte store of the two
// finished gradients -- a warp writes 256 contiguous bytes of one NCHW row.  No atomics, no
// zero-initialisation, no transposes; each output element is written exactly once.
#include <atomic>
#include "qpwc_async.cuh"

namespace qpwc {

template <int TH_>
struct BwdNchwCfg {
  static constexpr int D = 4, Q = 9, NDISP = 81;
  static constexpr int TH = TH_, TW = 128;                         // pixels per tile: TH rows x 64 column pairs
  static constexpr int XCOLS = TW + 2 * D, XROW = TH + 2 * D;      // haloed operand tile (TH+8) x 136 per channel
  static constexpr int KC = 8, NST = TH_ <= 2 ? 4 : 3;
  static constexpr int STAGE_BYTES = KC * XROW * XCOLS * 4;        // TH = 4: 52224
  static constexpr int OFF_BARS = NST * STAGE_BYTES;
  static constexpr int SMEM_BYTES = OFF_BARS + 2 * NST * 8;
  static constexpr int NCONS = TH * (TW / 2), NPROD = 128, NTHREADS = NCONS + NPROD;
  static constexpr int REG_CONS = 232, REG_PROD = 32;
  static_assert(STAGE_BYTES % 128 == 0, "TMA destination alignment");
  static_assert(SMEM_BYTES <= 232448, "shared memory budget (227 KB)");
  static_assert(NCONS * REG_CONS + NPROD * REG_PROD <= 65536, "register budget");
};

// WHICH = 0: g = g_prv, X = nxt.   WHICH = 1: g = g_nxt, X = prv.
template <int WHICH, class Cfg>
__global__ void __launch_bounds__(Cfg::NTHREADS, 1)
corr_bwd_nchw_kernel(const QPWC_GRID_CONSTANT TensorMap tmX, const float* __restrict__ out,
                     const float* __restrict__ g_out, float* __restrict__ g, int B, int H, int W, int C,
                     float slope, int tiles_x, int tiles_y, int ntiles) {
  constexpr int D = Cfg::D, Q = Cfg::Q, NDISP = Cfg::NDISP, TH = Cfg::TH, TW = Cfg::TW, XCOLS = Cfg::XCOLS;
  constexpr int XROW = Cfg::XROW, KC = Cfg::KC, NST = Cfg::NST, STAGE_BYTES = Cfg::STAGE_BYTES, NCONS = Cfg::NCONS;
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
    tma_prefetch_desc(&tmX);
    uint32_t gi = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int tx = tile % tiles_x, rest = tile / tiles_x, ty = rest % tiles_y, b = rest / tiles_y;
      for (int c = 0; c < nchunks; ++c, ++gi) {
        const int stage = (int)(gi % NST);
        mbar_wait_parked(&empty[stage], ((gi / NST) & 1u) ^ 1u);
        mbar_arrive_expect_tx(&full[stage], STAGE_BYTES);
        tma_load_4d(smem + stage * STAGE_BYTES, &tmX, &full[stage], tx * TW - D, ty * TH - D, c * KC, b);
      }
    }
    return;
  }

  // ================================================================================== consumers
  setmaxnreg_inc<Cfg::REG_CONS>();
  const int ti = tid / (TW / 2), t = tid % (TW / 2), lane = tid & 31;
  const uint32_t x_off = (uint32_t)((ti * XCOLS + 2 * t) * 4);   // stencil row m: + m * XCOLS * 4
  const float sc_pos = 1.f / (float)C, sc_neg = slope * sc_pos;
  const size_t plane = (size_t)H * W;

  uint32_t gi = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int tx = tile % tiles_x, rest = tile / tiles_x, ty = rest % tiles_y, b = rest / tiles_y;
    const int i = ty * TH + ti, p0 = tx * TW + 2 * t;           // this thread's pixels (i, p0), (i, p0+1)
    const bool live = i < H && p0 < W;                          // W is even: pairs are whole
    // whole warps outside the image only keep the pipeline moving
    const bool warp_live = i < H && (tx * TW + 2 * (t & ~31)) < W;

    // ------------------------------------------------------------------ coefficients -> registers
    // One stencil row (9 taps) at a time: all of its loads are issued before the first use, so a
    // thread has 18-36 requests in flight instead of one round trip per tap.
    float G0[Q][Q], G1[Q][Q];
    {
      const float* ob = out + (size_t)b * NDISP * plane;
      const float* gb = g_out + (size_t)b * NDISP * plane;
      if (WHICH == 0) {
        const size_t o0 = (size_t)(live ? i : 0) * W + (live ? p0 : 0);
        const float2* op = reinterpret_cast<const float2*>(ob + o0);
        const float2* gp = reinterpret_cast<const float2*>(gb + o0);
        const size_t pstep = plane >> 1;                         // plane is even (W % 4 == 0)
#pragma unroll
        for (int m = 0; m < Q; ++m) {
          float2 ov[Q], gv[Q];
#pragma unroll
          for (int k = 0; k < Q; ++k) { ov[k] = __ldg(op + (size_t)k * pstep); gv[k] = __ldg(gp + (size_t)k * pstep); }
#pragma unroll
          for (int k = 0; k < Q; ++k) {
            G0[m][k] = live ? gv[k].x * (ov[k].x > 0.f ? sc_pos : sc_neg) : 0.f;
            G1[m][k] = live ? gv[k].y * (ov[k].y > 0.f ? sc_pos : sc_neg) : 0.f;
          }
          op += (size_t)Q * pstep; gp += (size_t)Q * pstep;
        }
      } else {
        // stencil tap (m,k) of output pixel (i, x) is the forward pixel (i+m-4, x+k-4) under
        // displacement (di,dj) = (4-m, 4-k), i.e. plane (8-m)*9 + (8-k).  Columns are clamped for the
        // loads and the out-of-image taps zeroed afterwards, so the loads carry no predicates.
        int xc[Q + 1];
#pragma unroll
        for (int u = 0; u <= Q; ++u) xc[u] = min(max(p0 + u - D, 0), W - 1);
#pragma unroll
        for (int m = 0; m < Q; ++m) {
          const int y = i + m - D;
          const bool yok = live && y >= 0 && y < H;              // warp-uniform up to the right edge
          const size_t rowo = (size_t)((Q - 1 - m) * Q + (Q - 1)) * plane + (size_t)(yok ? y : 0) * W;
          float o0[Q], o1[Q], g0[Q], g1[Q];
          if (yok) {
#pragma unroll
            for (int k = 0; k < Q; ++k) {
              const float* po = ob + rowo - (size_t)k * plane;   // plane (8-m)*9 + (8-k)
              const float* pg = gb + rowo - (size_t)k * plane;
              o0[k] = __ldg(po + xc[k]); o1[k] = __ldg(po + xc[k + 1]);
              g0[k] = __ldg(pg + xc[k]); g1[k] = __ldg(pg + xc[k + 1]);
            }
          }
#pragma unroll
          for (int k = 0; k < Q; ++k) {
            const int x0 = p0 + k - D;
            const bool ok0 = yok && x0 >= 0 && x0 < W, ok1 = yok && x0 + 1 >= 0 && x0 + 1 < W;
            G0[m][k] = ok0 ? g0[k] * (o0[k] > 0.f ? sc_pos : sc_neg) : 0.f;
            G1[m][k] = ok1 ? g1[k] * (o1[k] > 0.f ? sc_pos : sc_neg) : 0.f;
          }
        }
      }
    }

    float* dst = g + (size_t)b * C * plane + (size_t)i * W + p0;
    for (int c = 0; c < nchunks; ++c, ++gi) {
      const int stage = (int)(gi % NST);
      mbar_wait(&full[stage], (gi / NST) & 1u);
      if (warp_live) {
        const unsigned char* sb = smem + stage * STAGE_BYTES + x_off;
        const int nch = min(KC, C - c * KC);
        for (int ch = 0; ch < nch; ++ch) {
          const unsigned char* xp = sb + ch * (XROW * XCOLS * 4);
          float a0[3] = {0.f, 0.f, 0.f}, a1[3] = {0.f, 0.f, 0.f};   // three independent chains per pixel
#pragma unroll
          for (int m = 0; m < Q; ++m) {
            float xr[10];
#pragma unroll
            for (int u = 0; u < 5; ++u) {
              const float2 v = *reinterpret_cast<const float2*>(xp + m * (XCOLS * 4) + u * 8);
              xr[2 * u] = v.x; xr[2 * u + 1] = v.y;
            }
#pragma unroll
            for (int k = 0; k < Q; ++k) {
              a0[m % 3] = fmaf(G0[m][k], xr[k], a0[m % 3]);
              a1[m % 3] = fmaf(G1[m][k], xr[k + 1], a1[m % 3]);
            }
          }
          if (live)
            *reinterpret_cast<float2*>(dst + (size_t)(c * KC + ch) * plane) =
                make_float2((a0[0] + a0[1]) + a0[2], (a1[0] + a1[1]) + a1[2]);
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[stage]);
    }
  }
}

// -------------------------------------------------------------------------------------- host
int sm_count_cached();  // qpwc_corr_tiled.cu

template <int WHICH, class Cfg>
static int run_bwd_nchw(const float* x, const float* out, const float* g_out, float* g, int B, int C, int H, int W,
                        float slope, cudaStream_t stream) {
  constexpr int SMEM_BYTES = Cfg::SMEM_BYTES;
  TensorMap tmX;
  if (!make_tmap_nchw(&tmX, x, B, C, H, W, Cfg::XCOLS, Cfg::XROW, Cfg::KC)) return QPWC_ERR_CUDA;
  const int tiles_x = cdiv(W, Cfg::TW), tiles_y = cdiv(H, Cfg::TH);
  const long long nt = (long long)tiles_x * tiles_y * B;
  if (nt >= (1LL << 31)) return QPWC_ERR_UNSUPPORTED;
  const int ntiles = (int)nt;
  const int grid = ntiles < sm_count_cached() ? ntiles : sm_count_cached();
  auto k = corr_bwd_nchw_kernel<WHICH, Cfg>;
#ifndef QPWC_EMU
  static std::atomic<unsigned> attr_done{0};  // per instantiation, one bit per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_done.load(std::memory_order_acquire) >> (dev & 31) & 1u)) {
    const cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "corr_bwd_nchw: smem attribute (%d B): %s", SMEM_BYTES, cudaGetErrorString(e));
    attr_done.fetch_or(1u << (dev & 31), std::memory_order_release);
  }
#endif
  QPWC_LAUNCH(k, grid, Cfg::NTHREADS, SMEM_BYTES, stream, tmX, out, g_out, g, B, H, W, C, slope, tiles_x, tiles_y, ntiles);
  return QPWC_OK;
}

// Shape-generic channels_first gradients (any search range / W / alignment; see corr_fwd_nchw_generic_kernel):
// one thread per gradient element (b, c, i, j) gathers its (2d+1)^2 terms of both gradients, every access a
// coalesced row segment:  g_prv += G'[(i,j),e] * nxt[i+di,j+dj],   g_nxt += G'[(i-di,j-dj),e] * prv[i-di,j-dj]
// with G' = g_out * (out > 0 ? 1 : slope) / C.
__global__ void __launch_bounds__(256) corr_bwd_nchw_generic_kernel(const float* __restrict__ prv, const float* __restrict__ nxt,
                                                                    const float* __restrict__ out, const float* __restrict__ g_out,
                                                                    float* __restrict__ g_prv, float* __restrict__ g_nxt,
                                                                    int C, int H, int W, int d, float slope, long long total) {
  const long long idx = (long long)blockIdx.x * 256 + threadIdx.x;
  if (idx >= total) return;
  const int q = 2 * d + 1, D = q * q;
  const int j = (int)(idx % W);
  long long r = idx / W;
  const int i = (int)(r % H); r /= H;
  const int c = (int)(r % C);
  const long long b = r / C;
  const size_t plane = (size_t)H * W;
  const float* gb = g_out + (size_t)b * D * plane;
  const float* ob = out + (size_t)b * D * plane;
  const float* pc = prv + ((size_t)b * C + c) * plane;
  const float* nc = nxt + ((size_t)b * C + c) * plane;
  const float pos = 1.f / (float)C, neg = slope / (float)C;
  float gp = 0.f, gn = 0.f;
  for (int e = 0; e < D; ++e) {
    const int di = e / q - d, dj = e % q - d;
    const int ia = i + di, ja = j + dj;          // second-frame pixel this first-frame pixel was matched with
    if (ia >= 0 && ia < H && ja >= 0 && ja < W) {
      const size_t o = (size_t)e * plane + (size_t)i * W + j;
      gp = fmaf(__ldg(gb + o) * (__ldg(ob + o) > 0.f ? pos : neg), __ldg(nc + (size_t)ia * W + ja), gp);
    }
    const int is = i - di, js = j - dj;          // first-frame pixel that matched this second-frame pixel at e
    if (is >= 0 && is < H && js >= 0 && js < W) {
      const size_t o = (size_t)e * plane + (size_t)is * W + js;
      gn = fmaf(__ldg(gb + o) * (__ldg(ob + o) > 0.f ? pos : neg), __ldg(pc + (size_t)is * W + js), gn);
    }
  }
  g_prv[idx] = gp;
  g_nxt[idx] = gn;
}

int launch_corr_bwd_nchw(const float* prv, const float* nxt, const float* out, const float* g_out,
                         float* g_prv, float* g_nxt, int B, int C, int H, int W, int d, float slope,
                         cudaStream_t stream) {
  // tiled kernels: d == 4, W a multiple of 4 (TMA strides are multiples of 16 bytes; 8-byte loads/stores of
  // pixel pairs), 16-byte aligned inputs, 8-byte aligned out / g_out / gradients; everything else is generic
  if (d < 1 || C < 1) return QPWC_ERR_UNSUPPORTED;
  if (d != 4 || (W & 3) || (reinterpret_cast<uintptr_t>(prv) & 15) || (reinterpret_cast<uintptr_t>(nxt) & 15) ||
      (reinterpret_cast<uintptr_t>(out) & 7) || (reinterpret_cast<uintptr_t>(g_out) & 7) ||
      (reinterpret_cast<uintptr_t>(g_prv) & 7) || (reinterpret_cast<uintptr_t>(g_nxt) & 7)) {
    const long long total = (long long)B * C * H * W;
    if (total == 0) return QPWC_OK;
    if (cdivll(total, 256) >= (1LL << 31)) return QPWC_ERR_UNSUPPORTED;
    QPWC_LAUNCH(corr_bwd_nchw_generic_kernel, (unsigned)cdivll(total, 256), 256, 0, stream, prv, nxt, out, g_out, g_prv, g_nxt, C, H, W, d, slope, total);
    return check_launch("corr_bwd_nchw_generic");
  }
  // few tiles (coarse pyramid levels): 2-row tiles double the number of busy SMs
  const long long tiles4 = (long long)cdiv(W, 128) * cdiv(H, 4) * B;
  int rc;
  if (tiles4 * 2 <= sm_count_cached()) {
    rc = run_bwd_nchw<0, BwdNchwCfg<2>>(nxt, out, g_out, g_prv, B, C, H, W, slope, stream);
    if (rc == QPWC_OK) rc = run_bwd_nchw<1, BwdNchwCfg<2>>(prv, out, g_out, g_nxt, B, C, H, W, slope, stream);
  } else {
    rc = run_bwd_nchw<0, BwdNchwCfg<4>>(nxt, out, g_out, g_prv, B, C, H, W, slope, stream);
    if (rc == QPWC_OK) rc = run_bwd_nchw<1, BwdNchwCfg<4>>(prv, out, g_out, g_nxt, B, C, H, W, slope, stream);
  }
  if (rc != QPWC_OK) return rc;
  return check_launch("corr_bwd_nchw");
}

}  // namespace qpwc
