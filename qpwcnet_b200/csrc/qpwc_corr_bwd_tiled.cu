// qpwc_corr_bwd_tiled.cu -- register-tiled gradient of the local-correlation cost volume (d = 4).
//
// Gradient of CostVolume / CostVolumeV2 (TF autodiff of qpwcnet/core/layers.py:77-99; tfa
// CorrelationCostGrad + LeakyReluGrad), with  G'[p,(di,dj)] = g_out * (out > 0 ? 1 : slope) / C :
//   g_prv[i,j,c] = sum_{di,dj} G'[(i,j),(di,dj)]       * nxt[i+di, j+dj, c]
//   g_nxt[i,j,c] = sum_{di,dj} G'[(i-di,j-dj),(di,dj)] * prv[i-di, j-dj, c]
//
// Both are the same banded product  R[p,c] = sum_e Gs[x,e] * X[x,c],  x = p + e (per displacement
// row), once the gradient slice is laid out per SOURCE pixel x ("skewed" for g_prv, as stored for
// g_nxt), so one kernel template serves both (instantiated per gradient).
//
// Decomposition: a CTA owns 4 rows x 64 pixels x 32 channels of one gradient; thread (row, g, h)
// owns 8 consecutive pixels x 8 channels = 32 fp32x2 accumulators.  It walks the nine displacement
// rows t; per t it sweeps the 16 source pixels x of its span: two 16-byte loads give X[x, 8 ch] as
// four natural channel pairs, three give the nine Gs[x, e], and every (p = x - e) in range takes
// four scalar-broadcast FFMA2s  acc2[p][cp] += Gs * (X[c], X[c+1])  (see qpwc_corr_rowpair.cu for
// the FFMA2 form).  The X rows of successive t are the same image rows shifted by one, so they
// live in a 5-slot ring in shared memory: each row is loaded once per tile (cp.async, in flight
// under the FMAs of the previous step).  The G' slice of a step is 9 of the 81 gradient channels
// of 4 rows: its elements are fetched into registers before the previous step's FMAs (element-
// linear, so a warp touches 4-5 cache lines per load; a per-thread table held in 21 registers maps
// element -> global offset / skewed destination) and masked, scaled and stored after them.
// Two 128-thread CTAs share an SM (three fit with a 4-byte table and 168 registers but run slower, 1130 vs
// 700 us at the finest level: only ~14 KB of L1 would remain for the gradient sectors shared by consecutive steps).  No atomics: every gradient element is written once, in a fixed
// summation order.
#include <atomic>
#include <stdlib.h>

#include "qpwc_common.cuh"

namespace qpwc {

namespace bwdcfg {
constexpr int D = 4, Q = 9, NDISP = 81;
constexpr int TH = 4, TW = 64, XW = TW + 2 * D;  // 72 source pixels per row
constexpr int CB = 32;                           // channels per CTA
constexpr int PXT = 8, CHT = 8;                  // per-thread tile
constexpr int NG = TW / PXT, NH = CB / CHT;      // 8 pixel groups x 4 channel groups (one row per warp)
constexpr int NTHREADS = TH * NG * NH;           // 128
constexpr int RING = TH + 1;
constexpr int XROW_BYTES = XW * CB * 4;          // 9216
constexpr int GS_XSTRIDE = 48;                   // 9 floats padded to 12
constexpr int GS_ROW_BYTES = XW * GS_XSTRIDE + (XW / PXT + 1) * 16;  // + 16 bytes of skew per 8 pixels
constexpr int NEL = TH * XW * Q;                 // gradient elements per step (g_nxt; g_prv uses TH*TW*Q)
constexpr int NK = (NEL + NTHREADS - 1) / NTHREADS;  // 21 per thread
constexpr int OFF_GS = RING * XROW_BYTES;
constexpr int OFF_DUMMY = OFF_GS + TH * GS_ROW_BYTES;  // sink for the padding elements of the table
constexpr int SMEM_BYTES = OFF_DUMMY + 16;
constexpr int GOFF_BITS = 20;  // per-thread element table entry: global element offset | Gs word offset << 20
static_assert(SMEM_BYTES <= 113 * 1024, "two CTAs per SM");
static_assert(GS_ROW_BYTES % 16 == 0, "alignment");
static_assert((OFF_DUMMY - OFF_GS) / 4 + 4 < (1 << (32 - GOFF_BITS)), "Gs word offset fits the table entry");
}  // namespace bwdcfg

// X ring: [slot][x][8 units of 16 B], unit index XORed with bit 3 of x: the 32 lanes (g, h) of a
// row then cover every bank group exactly four times per 16-byte load (the minimum).
__device__ __forceinline__ uint32_t xr_off(int slot, int x, int unit) {
  return (uint32_t)(slot * bwdcfg::XROW_BYTES + x * (bwdcfg::CB * 4) + ((unit ^ ((x >> 3) & 1)) << 4));
}
// Gs: [row][x][12 floats] + 16 bytes of skew per 8 pixels: the 8 pixel groups of a warp read 8 bank groups
__device__ __forceinline__ uint32_t gs_off(int r, int x) {
  return (uint32_t)(bwdcfg::OFF_GS + r * bwdcfg::GS_ROW_BYTES + x * bwdcfg::GS_XSTRIDE + ((x >> 3) << 4));
}
#ifndef QPWC_EMU
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst));
  const int n = valid ? 16 : 0;  // src-size 0: the 16 bytes are zero-filled
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_group 0;" ::: "memory"); }

template <int WHICH>  // 0 = g_prv, 1 = g_nxt
__global__ void __launch_bounds__(bwdcfg::NTHREADS, 2)
corr_bwd_tiled_kernel(const float* __restrict__ prv, const float* __restrict__ nxt,
                      const float* __restrict__ out, const float* __restrict__ g_out,
                      float* __restrict__ g_prv, float* __restrict__ g_nxt, int H, int W, int C,
                      float slope, long long ops, int tiles_x, int ncb) {
  using namespace bwdcfg;
  QPWC_DYN_SMEM(smem);
  const int tid = threadIdx.x;
  const int h = tid % NH, g = (tid / NH) % NG, r = tid / (NH * NG);  // warp = row r, lane = (g, h)
  const int tx = blockIdx.x % tiles_x, ty = blockIdx.x / tiles_x;
  const int cb = blockIdx.y;
  constexpr int which = WHICH;
  const int b = blockIdx.z;
  const int i0 = ty * TH, j0 = tx * TW, c0 = cb * CB;
  const float* X = which == 0 ? nxt : prv;
  float* R = which == 0 ? g_prv : g_nxt;
  const size_t bpix = (size_t)b * H * W;
  const float inv_c = 1.f / (float)C;
  const int C4 = C >> 2;

  // ---- element table.  Element n of a step = (rr, px, e): tile row, pixel of the slice, channel
  //      of the 9-channel slice.   g_prv: pixel (i0+rr, j0+px), px < 64;  source x = px+e, slot e.
  //      g_nxt: pixel (X row of the step, j0-4+px), px < 72;  source x = px, slot 8-e.
  //      entry (one register per element, loop invariant) = global element offset relative to
  //      (row0 of the step, column 0 of the slice, channel 0 of the slice) in the low 20 bits |
  //      Gs word offset (relative to OFF_GS) << 20.  Padding elements (n >= nel) re-read element 0
  //      and store into a dummy word.
  constexpr int npx = which == 0 ? TW : XW;
  constexpr int nel = TH * npx * Q;
  uint32_t tab[NK];
#pragma unroll
  for (int k = 0; k < NK; ++k) {
    const int n = tid + k * NTHREADS;
    tab[k] = (uint32_t)((OFF_DUMMY - OFF_GS) / 4) << GOFF_BITS;
    if (n < nel) {
      const int e = n % Q, pr = n / Q, px = pr % npx, rr = pr / npx;   // compile-time divisors
      const int xs = which == 0 ? px + e : px, es = which == 0 ? e : Q - 1 - e;
      tab[k] = (uint32_t)((rr * W + px) * (int)ops + e) | ((gs_off(rr, xs) - OFF_GS + es * 4) / 4) << GOFF_BITS;
    }
  }
  // tile classification: can every slice / X pixel of every step be read without bounds checks?
  const int colbase = which == 0 ? j0 : j0 - D;
  const bool interior = i0 - D >= 0 && i0 + TH + D <= H && j0 - D >= 0 && j0 + TW + D <= W;

  // X rows: piece = (source pixel xs, 16-byte unit); 128 threads = 16 pixels x 8 units per pass, so
  // a thread keeps its unit and walks xs = xs0, xs0+16, ... (the swizzle bit (xs>>3)&1 is invariant)
  const int xunit = tid & 7, xs0 = tid >> 3;
  const uint32_t xdst0 = xr_off(0, xs0, xunit);
  const bool xch_ok = c0 + xunit * 4 < C;
  auto issue_xrow = [&](int n) {  // ring row n = image row i0 - D + n  ->  slot n % RING (cp.async, zero fill)
    const int row = i0 - D + n;
    unsigned char* dst = smem + (n % RING) * XROW_BYTES + xdst0;
    const bool row_ok = row >= 0 && row < H && xch_ok;
    const float4* src = reinterpret_cast<const float4*>(X) + (bpix + (size_t)(row_ok ? row : 0) * W) * C4 + ((c0 >> 2) + xunit);
#pragma unroll
    for (int x = 0; x < (XW + 15) / 16; ++x) {
      const int xs = xs0 + 16 * x, col = j0 - D + xs;
      if (xs < XW) {
        const bool ok = row_ok && col >= 0 && col < W;
        cp_async16(dst + x * 16 * (CB * 4), src + (ok ? (size_t)col * C4 : 0), ok);
      }
    }
  };
  // gradient slice of step t: global base (element units) of its (row0, column 0, channel 0)
  auto slice_base = [&](int t, int& row0) -> long long {
    const int dr = which == 0 ? t : (Q - 1 - t);  // displacement row index di + D of this step
    row0 = which == 0 ? i0 : i0 - D + t;
    return ((long long)bpix + (long long)row0 * W + colbase) * ops + dr * Q;
  };
  float gq[NK], oq[NK];
  auto fetch_gs = [&](int t) {
    int row0;
    const long long base = slice_base(t, row0);
    const float* gb_ = g_out + base;
    const float* ob_ = out + base;
    if (interior) {  // CTA-uniform: no bounds checks
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const uint32_t off = tab[k] & ((1u << GOFF_BITS) - 1);
        gq[k] = __ldg(gb_ + off); oq[k] = __ldg(ob_ + off);
      }
    } else {
#pragma unroll
      for (int k = 0; k < NK; ++k) {
        const int n = tid + k * NTHREADS;
        const int pr = n / Q, px = pr % npx, rr = pr / npx;
        const int row = row0 + rr, col = colbase + px;
        const bool ok = n < nel && row >= 0 && row < H && col >= 0 && col < W;
        const uint32_t off = tab[k] & ((1u << GOFF_BITS) - 1);
        gq[k] = 0.f; oq[k] = 1.f;
        if (ok) { gq[k] = __ldg(gb_ + off); oq[k] = __ldg(ob_ + off); }
      }
    }
  };
  const float sc_pos = inv_c, sc_neg = slope * inv_c;
  auto store_gs = [&]() {
#pragma unroll
    for (int k = 0; k < NK; ++k)
      *reinterpret_cast<float*>(smem + OFF_GS + (tab[k] >> GOFF_BITS) * 4u) = gq[k] * (oq[k] > 0.f ? sc_pos : sc_neg);
  };

  float2 acc[PXT][CHT / 2];
#pragma unroll
  for (int p = 0; p < PXT; ++p)
#pragma unroll
    for (int c = 0; c < CHT / 2; ++c) acc[p][c] = make_float2(0.f, 0.f);

  // prologue: ring rows 0 .. TH-1 and the slice of step 0 (each thread reads only its own table entries)
  for (int n = 0; n < TH; ++n) issue_xrow(n);
  cp_async_commit();
  fetch_gs(0);
  store_gs();
  cp_async_wait_all();
  __syncthreads();

  for (int t = 0; t < Q; ++t) {
    const bool more = t + 1 < Q;
    if (more) {
      // ring row TH+t (needed from step t+1) overwrites the slot of row t-1, last read in step t-1:
      // every thread is past the barrier that closed that step
      issue_xrow(TH + t);
      cp_async_commit();
      fetch_gs(t + 1);  // in flight during this step's FMAs
    }
    const int slot = (r + t) % RING;
    // per-thread operand bases; everything else is an immediate.  Source pixel x = 8g + xl: the
    // swizzle bit and the Gs skew of x are those of group g for xl < 8 and of group g+1 above.
    const unsigned char* xb = smem + (uint32_t)(slot * XROW_BYTES + g * PXT * (CB * 4));
    const unsigned char* gb = smem + OFF_GS + r * GS_ROW_BYTES + g * (PXT * GS_XSTRIDE + 16);
    const uint32_t ua0 = (uint32_t)(((2 * h) ^ (g & 1)) << 4), ua1 = (uint32_t)(((2 * h + 1) ^ (g & 1)) << 4);
    const uint32_t ub0 = ua0 ^ 16u, ub1 = ua1 ^ 16u;
#pragma unroll
    for (int xl = 0; xl < PXT + Q - 1; ++xl) {
      const float4 xa = *reinterpret_cast<const float4*>(xb + xl * (CB * 4) + (xl < PXT ? ua0 : ub0));
      const float4 xc = *reinterpret_cast<const float4*>(xb + xl * (CB * 4) + (xl < PXT ? ua1 : ub1));
      const unsigned char* gp = gb + xl * GS_XSTRIDE + (xl < PXT ? 0 : 16);
      const float4 ga = *reinterpret_cast<const float4*>(gp);
      const float4 gc = *reinterpret_cast<const float4*>(gp + 16);
      const float g8 = *reinterpret_cast<const float*>(gp + 32);
      const float gv[Q] = {ga.x, ga.y, ga.z, ga.w, gc.x, gc.y, gc.z, gc.w, g8};
      const float2 x0 = make_float2(xa.x, xa.y), x1 = make_float2(xa.z, xa.w);
      const float2 x2 = make_float2(xc.x, xc.y), x3 = make_float2(xc.z, xc.w);
#pragma unroll
      for (int e = 0; e < Q; ++e) {
        const int p = xl - e;
        if (p >= 0 && p < PXT) {
          const float2 gg = make_float2(gv[e], gv[e]);
          acc[p][0] = __ffma2_rn(gg, x0, acc[p][0]);
          acc[p][1] = __ffma2_rn(gg, x1, acc[p][1]);
          acc[p][2] = __ffma2_rn(gg, x2, acc[p][2]);
          acc[p][3] = __ffma2_rn(gg, x3, acc[p][3]);
        }
      }
    }
    if (more) {
      __syncthreads();  // all readers of Gs(t) are done
      store_gs();
      cp_async_wait_all();
      __syncthreads();  // Gs(t+1) and ring row TH+t visible
    }
  }

  // ---- write the 8 x 8 block
  const int i = i0 + r;
  if (i < H) {
    const int ch = c0 + h * CHT;
#pragma unroll
    for (int p = 0; p < PXT; ++p) {
      const int j = j0 + g * PXT + p;
      if (j < W) {
        float4* dst = reinterpret_cast<float4*>(R) + (bpix + (size_t)i * W + j) * C4 + (ch >> 2);
        if (ch < C) dst[0] = make_float4(acc[p][0].x, acc[p][0].y, acc[p][1].x, acc[p][1].y);
        if (ch + 4 < C) dst[1] = make_float4(acc[p][2].x, acc[p][2].y, acc[p][3].x, acc[p][3].y);
      }
    }
  }
}
#endif  // !QPWC_EMU

int get_corr_bwd_variant();   // qpwc_api.cu (QPWC_OPT_CORR_BWD: 0 auto, 1 untiled kernels everywhere)

int launch_corr_bwd_tiled(const float* prv, const float* nxt, const float* out, const float* g_out,
                          float* g_prv, float* g_nxt, int B, int H, int W, int C, int d, float slope,
                          long long ops, cudaStream_t stream) {
  using namespace bwdcfg;
  // domain: d == 4, C a multiple of 4, 16-byte aligned feature tensors, grid dimensions in range
  if (d != 4 || (C & 3) || C < 4) return QPWC_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(prv) | reinterpret_cast<uintptr_t>(nxt) |
       reinterpret_cast<uintptr_t>(g_prv) | reinterpret_cast<uintptr_t>(g_nxt)) & 15) return QPWC_ERR_UNSUPPORTED;
  if (get_corr_bwd_variant() == 1) return QPWC_ERR_UNSUPPORTED;  // QPWC_OPT_CORR_BWD: untiled kernels forced
  const int tiles_x = cdiv(W, TW), tiles_y = cdiv(H, TH), ncb = cdiv(C, CB);
  if (ncb > 65535 || B > 65535) return QPWC_ERR_UNSUPPORTED;
  if (((long long)(TH - 1) * W + XW) * ops + Q >= (1LL << GOFF_BITS)) return QPWC_ERR_UNSUPPORTED;  // table entry range
#ifndef QPWC_EMU
  static std::atomic<unsigned> attr_done{0};  // one bit per device (the attribute is per device)
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(attr_done.load(std::memory_order_acquire) >> (dev & 31) & 1u)) {
    cudaError_t e = cudaFuncSetAttribute(corr_bwd_tiled_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(corr_bwd_tiled_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES);
    if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "corr_bwd_tiled: smem attribute (%d B): %s", SMEM_BYTES, cudaGetErrorString(e));
    attr_done.fetch_or(1u << (dev & 31), std::memory_order_release);
  }
  const dim3 grid((unsigned)(tiles_x * tiles_y), (unsigned)ncb, (unsigned)B);
  corr_bwd_tiled_kernel<0><<<grid, NTHREADS, SMEM_BYTES, stream>>>(prv, nxt, out, g_out, g_prv, g_nxt, H, W, C, slope, ops, tiles_x, ncb);
  corr_bwd_tiled_kernel<1><<<grid, NTHREADS, SMEM_BYTES, stream>>>(prv, nxt, out, g_out, g_prv, g_nxt, H, W, C, slope, ops, tiles_x, ncb);
  return check_launch("corr_bwd_tiled");
#else
  return QPWC_ERR_UNSUPPORTED;
#endif
}

}  // namespace qpwc
