// qpwc_warp.cu -- bilinear backward warp, forward and backward, NHWC fp32.
//
// Replaces  tf_warp / Warp   (qpwcnet/core/warp.py:63-153, layers.py:144-168)   -> MODE_TF
//           WarpV2 / tfa dense_image_warp (qpwcnet/core/layers.py:171-186)        -> MODE_TFA
//
// Gather-bound (HBM/L2): algorithmic bytes per output pixel 4*(2C+2) forward, 4*(3C+4) backward.
// Forward : one thread per (pixel, V-channel vector), V = 4/2/1 by C and pointer alignment; the
//           V-lanes of one pixel share the flow load (warp broadcast) and the tap set-up.
// Backward: sub-warp groups of G lanes own one pixel at a time (G = pow2 >= C/V, <= 32); each lane
//           scatter-adds its channel vector into the four taps (vector red.global.add) and the
//           flow gradient is reduced across the group with xor-shuffles in a fixed order.
#include <atomic>
#include <stdlib.h>

#include "qpwc_upsample.cuh"

namespace qpwc {

template <int V> struct Vec;
template <> struct Vec<4> { typedef float4 type; };
template <> struct Vec<2> { typedef float2 type; };
template <> struct Vec<1> { typedef float type; };

template <int V> __device__ __forceinline__ void vload(const float* p, float (&v)[V]);
template <> __device__ __forceinline__ void vload<4>(const float* p, float (&v)[4]) {
  const float4 t = __ldg(reinterpret_cast<const float4*>(p));
  v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
}
template <> __device__ __forceinline__ void vload<2>(const float* p, float (&v)[2]) {
  const float2 t = __ldg(reinterpret_cast<const float2*>(p));
  v[0] = t.x; v[1] = t.y;
}
template <> __device__ __forceinline__ void vload<1>(const float* p, float (&v)[1]) { v[0] = __ldg(p); }

template <int V> __device__ __forceinline__ void vstore(float* p, const float (&v)[V]);
template <> __device__ __forceinline__ void vstore<4>(float* p, const float (&v)[4]) {
  *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> __device__ __forceinline__ void vstore<2>(float* p, const float (&v)[2]) {
  *reinterpret_cast<float2*>(p) = make_float2(v[0], v[1]);
}
template <> __device__ __forceinline__ void vstore<1>(float* p, const float (&v)[1]) { *p = v[0]; }

template <int V> __device__ __forceinline__ void vatomic_add(float* p, const float (&v)[V]);
template <> __device__ __forceinline__ void vatomic_add<4>(float* p, const float (&v)[4]) {
#ifdef QPWC_EMU
  for (int k = 0; k < 4; ++k) atomicAdd(p + k, v[k]);
#else
  atomicAdd(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));  // red.global.add.v4.f32
#endif
}
template <> __device__ __forceinline__ void vatomic_add<2>(float* p, const float (&v)[2]) {
#ifdef QPWC_EMU
  for (int k = 0; k < 2; ++k) atomicAdd(p + k, v[k]);
#else
  atomicAdd(reinterpret_cast<float2*>(p), make_float2(v[0], v[1]));
#endif
}
template <> __device__ __forceinline__ void vatomic_add<1>(float* p, const float (&v)[1]) { atomicAdd(p, v[0]); }

// ------------------------------------------------------------------------------------------ fwd
// grid = (ceil(W*CV / 256), H, B): row and batch come from the block index, so the only index
// arithmetic per thread is one 32-bit divide by CV (the first version decoded a 64-bit linear index
// with three 64-bit divisions per thread and was instruction-bound: ncu 207 instr/thread).
//
// Extended form (FrameInterpolate, qpwcnet/core/non_layers.py:303-311): the flow is multiplied by
// `scale` first (0.5 * flo; one rounded multiply, exactly the reference's op), the output pixel
// stride `ops` may exceed C (channel slice of a concat buffer), and batch entries z >= bsplit take
// a second (image, flow) pair and write C channels further -- two warps in one launch.
#define WARP_FWD_ROWS 8
template <int MODE, int V, int NV>  // NV channel vectors per thread (2: taps amortised over 32 bytes per tap)
__global__ void __launch_bounds__(256) warp_fwd_kernel(const float* __restrict__ img,
                                                       const float* __restrict__ flow,
                                                       const float* __restrict__ img2,
                                                       const float* __restrict__ flow2,
                                                       float* __restrict__ out, int H, int W, int C,
                                                       int bsplit, float scale, long long ops,
                                                       float up_scale, int row_off, int Hfull, int keep_l2) {
  // keep_l2: the output is the scratch of the UpFlow pair, read back by the cost-volume kernel that follows on
  // the stream -- stored with an evict-last L2 policy
  // row_off / Hfull (row-sharded frames, qpwcnet_b200/sharded.py): `img` is rows [row_off, row_off + H) of
  // an image Hfull rows tall.  The sampling coordinate, truncation and clamping use the ABSOLUTE row --
  // the reference adds the flow to the absolute pixel index in fp32, so its rounding depends on it --
  // and the taps are translated back into the view (the caller keeps only rows whose taps lie inside).
  // up_scale != 0: `flow` is the COARSE flow (B, H/2, W/2, 2); the sampling flow is
  // up_scale * bilinear_x2(flow), interpolated here instead of being read back from HBM
  // (Upsample(scale=2.0) feeding UpFlow's warp, non_layers.py:183-193, pwcnet.py:49-56)
  const int CV = C / (V * NV);
  // a block covers WARP_FWD_ROWS image rows x 64 (pixel, vector) slots: the bottom taps of one row are
  // the top taps of the next, so vertically adjacent pixels share their source lines in L1
  const int idx = blockIdx.x * (256 / WARP_FWD_ROWS) + (threadIdx.x % (256 / WARP_FWD_ROWS));  // j * CV + cv
  const int i = blockIdx.y * WARP_FWD_ROWS + threadIdx.x / (256 / WARP_FWD_ROWS);
  if (idx >= W * CV || i >= H) return;
  const int j = idx / CV, cv = idx - j * CV;
  const bool second = (int)blockIdx.z >= bsplit;
  const int b = second ? (int)blockIdx.z - bsplit : (int)blockIdx.z;
  if (second) { img = img2; flow = flow2; }
  const size_t row = (size_t)b * H + i;                    // b*H + i
  const size_t pix = row * W + j;
  float2 f;
  if (up_scale != 0.f) f = up2_flow(flow + (size_t)b * (H / 2) * (W / 2) * 2, i, j, H / 2, W / 2, up_scale);
  else f = __ldg(reinterpret_cast<const float2*>(flow) + pix);    // ch0 = x, ch1 = y
  f.x = __fmul_rn(scale, f.x); f.y = __fmul_rn(scale, f.y);       // exact for scale == 1
  Taps t = make_taps<MODE>(i + row_off, j, f.x, f.y, Hfull, W);
  if (row_off != 0 || Hfull != H) {
    const int lim = (H - 1) * W, sh = row_off * W;
    t.o00 = min(max(t.o00 - sh, 0), lim + W - 1); t.o01 = min(max(t.o01 - sh, 0), lim + W - 1);
    t.o10 = min(max(t.o10 - sh, 0), lim + W - 1); t.o11 = min(max(t.o11 - sh, 0), lim + W - 1);
  }
  // vector n of lane cv is channel vector n*CV + cv: per load/store instruction the CV lanes of a
  // pixel cover CV*V contiguous floats (whole 32-byte sectors), not every other 16 bytes
  const float* base = img + (size_t)b * H * W * C + (size_t)cv * V;
  float v00[NV][V], v01[NV][V], v10[NV][V], v11[NV][V], o[NV][V];
#pragma unroll
  for (int n = 0; n < NV; ++n) {
    vload<V>(base + (size_t)t.o00 * C + n * CV * V, v00[n]);
    vload<V>(base + (size_t)t.o01 * C + n * CV * V, v01[n]);
    vload<V>(base + (size_t)t.o10 * C + n * CV * V, v10[n]);
    vload<V>(base + (size_t)t.o11 * C + n * CV * V, v11[n]);
  }
  float* dst = out + pix * (size_t)ops + (second ? C : 0) + (size_t)cv * V;
#pragma unroll
  for (int n = 0; n < NV; ++n) {
#pragma unroll
    for (int k = 0; k < V; ++k) o[n][k] = blend<MODE>(t, v00[n][k], v01[n][k], v10[n][k], v11[n][k]);
#ifndef QPWC_EMU
    if (V == 4 && keep_l2) {
      uint64_t pol;
      asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
      asm volatile("st.global.L2::cache_hint.v4.f32 [%0], {%1, %2, %3, %4}, %5;"
                   ::"l"(dst + n * CV * V), "f"(o[n][0]), "f"(o[n][V > 1 ? 1 : 0]), "f"(o[n][V > 2 ? 2 : 0]), "f"(o[n][V > 3 ? 3 : 0]),
                     "l"(pol) : "memory");
      continue;
    }
#endif
    vstore<V>(dst + n * CV * V, o[n]);
  }
}

// ------------------------------------------------------------------------- channels_first fwd
// img (B,C,H,W), flow (B,2,H,W) (plane 0 = x, plane 1 = y), out (B,C,H,W).  One thread per pixel:
// the taps are set up once and every channel plane is sampled with them -- consecutive lanes are
// consecutive columns, so each of the four gathers of a plane is (flow noise aside) one coalesced
// row segment.  Same per-element arithmetic as the NHWC kernel (bit-identical results).
template <int MODE>
__global__ void __launch_bounds__(256) warp_fwd_nchw_kernel(const float* __restrict__ img,
                                                            const float* __restrict__ flow,
                                                            float* __restrict__ out, int C, int H, int W,
                                                            float scale, int nsplit, int Ctot) {
  // blockIdx.z = (batch item, channel split): this thread samples planes [cs*C, cs*C + C) of the Ctot planes
  // of its batch item (the chain over the planes is serial per thread, so levels with few pixels and many
  // planes are split over more threads).
  // block = 32 columns x 8 rows: a warp is one 128-byte row segment, and vertically adjacent pixels
  // (whose taps share source rows) sit in the same block, i.e. the same L1
  const int j = blockIdx.x * 32 + (threadIdx.x & 31);
  const int i = blockIdx.y * 8 + (threadIdx.x >> 5);
  if (j >= W || i >= H) return;
  const size_t plane = (size_t)H * W;
  const size_t pix = (size_t)i * W + j;
  const int bz = (int)blockIdx.z / nsplit, cs = (int)blockIdx.z - bz * nsplit;
  const float* fb = flow + (size_t)bz * 2 * plane;
  const float fx = __fmul_rn(scale, __ldg(fb + pix)), fy = __fmul_rn(scale, __ldg(fb + plane + pix));
  const Taps t = make_taps<MODE>(i, j, fx, fy, H, W);
  const float* src = img + ((size_t)bz * Ctot + (size_t)cs * C) * plane;
  float* dst = out + ((size_t)bz * Ctot + (size_t)cs * C) * plane + pix;
  // eight planes at a time, all 32 gathers issued before the first blend: left to the unroller the loads of
  // a plane were issued only after the store of the plane before (~6 k clk per four planes; ncu: issue
  // slots 22 % busy, nothing saturated): the per-thread chain over the planes is what bounds this kernel
  constexpr int G = 8;
  int c = 0;
  for (; c + G <= C; c += G, src += G * plane, dst += G * plane) {
    float v[G][4];
#pragma unroll
    for (int k = 0; k < G; ++k) {
      const float* sp = src + (size_t)k * plane;
      v[k][0] = __ldg(sp + t.o00); v[k][1] = __ldg(sp + t.o01); v[k][2] = __ldg(sp + t.o10); v[k][3] = __ldg(sp + t.o11);
    }
#pragma unroll
    for (int k = 0; k < G; ++k) dst[(size_t)k * plane] = blend<MODE>(t, v[k][0], v[k][1], v[k][2], v[k][3]);
  }
  for (; c < C; ++c, src += plane, dst += plane)
    *dst = blend<MODE>(t, __ldg(src + t.o00), __ldg(src + t.o01), __ldg(src + t.o10), __ldg(src + t.o11));
}

int launch_warp_fwd_nchw(const float* img, const float* flow, float* out, int B, int C, int H, int W,
                         int mode, float scale, cudaStream_t stream) {
  if ((long long)B * C * H * W == 0) return QPWC_OK;
  if (H > 8 * 65535 || B > 65535) return set_error(QPWC_ERR_UNSUPPORTED, "warp_fwd_nchw: H > 524280 or B > 65535");
  // planes per thread: all of them while the pixels alone fill the machine a few times over, else a divisor of C
  int cpt = C;
  const long long px = (long long)B * H * W;
  while (cpt >= 16 && cpt % 2 == 0 && px * (C / cpt) < 600000 && (long long)B * (C / cpt) * 2 <= 65535) cpt /= 2;
  const int nsplit = C / cpt;
  const dim3 grid((unsigned)cdiv(W, 32), (unsigned)cdiv(H, 8), (unsigned)(B * nsplit));
  if (mode == QPWC_MODE_TF) {
    auto k = warp_fwd_nchw_kernel<QPWC_MODE_TF>;
    QPWC_LAUNCH(k, grid, 256, 0, stream, img, flow, out, cpt, H, W, scale, nsplit, C);
  } else {
    auto k = warp_fwd_nchw_kernel<QPWC_MODE_TFA>;
    QPWC_LAUNCH(k, grid, 256, 0, stream, img, flow, out, cpt, H, W, scale, nsplit, C);
  }
  return check_launch("warp_fwd_nchw");
}

// ------------------------------------------------------------------------- channels_first bwd
// One thread per pixel, looping over the channel planes: consecutive lanes are consecutive columns,
// so the g_out reads, the four tap gathers and the four atomics of a plane are (flow noise aside)
// coalesced row segments, and the flow gradient is a per-thread sum over the channels in order -- no
// cross-lane reduction.  Per-element arithmetic of warp_bwd_kernel (V = 1).  g_img pre-zeroed.
template <int MODE>
__global__ void __launch_bounds__(256) warp_bwd_nchw_kernel(const float* __restrict__ img,
                                                            const float* __restrict__ flow,
                                                            const float* __restrict__ g_out,
                                                            float* __restrict__ g_img,
                                                            float* __restrict__ g_flow, int C, int H, int W,
                                                            float scale, int nsplit, int Ctot) {
  // blockIdx.z = (batch item, channel split), as in the forward kernel; with nsplit > 1 the flow gradient
  // (a sum over the planes) is accumulated with atomics on a pre-zeroed g_flow
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= W) return;
  const int i = blockIdx.y;
  const size_t plane = (size_t)H * W;
  const size_t pix = (size_t)i * W + j;
  const int bz = (int)blockIdx.z / nsplit, cs = (int)blockIdx.z - bz * nsplit;
  const float* fb = flow + (size_t)bz * 2 * plane;
  const float fx = __fmul_rn(scale, __ldg(fb + pix)), fy = __fmul_rn(scale, __ldg(fb + plane + pix));
  bool px = true, py = true;
  Taps t;
  if (MODE == QPWC_MODE_TF) t = taps_tf(i, j, fx, fy, H, W);
  else t = taps_tfa(i, j, fx, fy, H, W, &px, &py);
  float ax1 = 0.f, ax0 = 0.f, ay1 = 0.f, ay0 = 0.f;
  if (MODE == QPWC_MODE_TF) {
    const float x = __fadd_rn((float)j, fx), y = __fadd_rn((float)i, fy);
    const int y0 = t.o00 / W, x0 = t.o00 - y0 * W, y1 = t.o11 / W, x1 = t.o11 - y1 * W;
    ax1 = __fsub_rn((float)x1, x); ax0 = __fsub_rn(x, (float)x0);
    ay1 = __fsub_rn((float)y1, y); ay0 = __fsub_rn(y, (float)y0);
  }
  const bool dupx = (MODE == QPWC_MODE_TF) && (t.o00 == t.o01);
  const bool dupy = (MODE == QPWC_MODE_TF) && (t.o00 == t.o10);
  const size_t cbase = ((size_t)bz * Ctot + (size_t)cs * C) * plane;
  const float* src = img + cbase;
  const float* gsrc = g_out + cbase + pix;
  float* gi = g_img + cbase;
  float gx = 0.f, gy = 0.f;
#pragma unroll 2
  for (int c = 0; c < C; ++c, src += plane, gsrc += plane, gi += plane) {
    const float v00 = __ldg(src + t.o00), v01 = __ldg(src + t.o01), v10 = __ldg(src + t.o10), v11 = __ldg(src + t.o11);
    const float g = __ldg(gsrc);
    float a00, a01, a10, a11;
    if (MODE == QPWC_MODE_TF) {
      a00 = __fmul_rn(t.w00, g); a10 = __fmul_rn(t.w10, g);
      a01 = __fmul_rn(t.w01, g); a11 = __fmul_rn(t.w11, g);
      const float sx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(-ay1, v00), __fmul_rn(-ay0, v10)),
                                           __fmul_rn(ay1, v01)), __fmul_rn(ay0, v11));
      const float sy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(-ax1, v00), __fmul_rn(ax1, v10)),
                                           __fmul_rn(-ax0, v01)), __fmul_rn(ax0, v11));
      gx = __fadd_rn(gx, __fmul_rn(g, sx));
      gy = __fadd_rn(gy, __fmul_rn(g, sy));
    } else {
      const float ax = t.w00, ay = t.w01;
      const float top = __fadd_rn(__fmul_rn(ax, __fsub_rn(v01, v00)), v00);
      const float bot = __fadd_rn(__fmul_rn(ax, __fsub_rn(v11, v10)), v10);
      gy = __fadd_rn(gy, __fmul_rn(g, __fsub_rn(bot, top)));
      const float g_bot = __fmul_rn(ay, g);
      const float g_top = __fsub_rn(g, g_bot);
      gx = __fadd_rn(gx, __fadd_rn(__fmul_rn(g_top, __fsub_rn(v01, v00)), __fmul_rn(g_bot, __fsub_rn(v11, v10))));
      const float g_tr = __fmul_rn(ax, g_top), g_br = __fmul_rn(ax, g_bot);
      a01 = g_tr; a00 = __fsub_rn(g_top, g_tr); a11 = g_br; a10 = __fsub_rn(g_bot, g_br);
    }
    // clipped taps coincide (mode TF): fold in registers so that they cancel exactly (see below)
    if (dupx) { a00 = __fadd_rn(a00, a01); a10 = __fadd_rn(a10, a11); }
    if (dupy) { a00 = __fadd_rn(a00, a10); a01 = __fadd_rn(a01, a11); }
    atomicAdd(gi + t.o00, a00);
    if (!dupx) atomicAdd(gi + t.o01, a01);
    if (!dupy) atomicAdd(gi + t.o10, a10);
    if (!dupx && !dupy) atomicAdd(gi + t.o11, a11);
  }
  float* gf = g_flow + (size_t)bz * 2 * plane + pix;
  if (nsplit == 1) {
    gf[0] = px ? __fmul_rn(scale, gx) : 0.f;
    gf[plane] = py ? __fmul_rn(scale, gy) : 0.f;
  } else {
    if (px) atomicAdd(gf, __fmul_rn(scale, gx));
    if (py) atomicAdd(gf + plane, __fmul_rn(scale, gy));
  }
}

int launch_warp_bwd_nchw(const float* img, const float* flow, const float* g_out, float* g_img,
                         float* g_flow, int B, int C, int H, int W, int mode, float scale, cudaStream_t stream) {
  const long long n = (long long)B * C * H * W;
  if (n == 0) return QPWC_OK;
  if (H > 65535 || B > 65535) return set_error(QPWC_ERR_UNSUPPORTED, "warp_bwd_nchw: H or B > 65535");
  cudaError_t e = cudaMemsetAsync(g_img, 0, sizeof(float) * (size_t)n, stream);
  if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "warp_bwd_nchw: memset g_img: %s", cudaGetErrorString(e));
  // planes per thread as in the forward launcher; a split needs g_flow zeroed (partial sums are added atomically)
  int cpt = C;
  const long long px = (long long)B * H * W;
  while (cpt >= 16 && cpt % 2 == 0 && px * (C / cpt) < 600000 && (long long)B * (C / cpt) * 2 <= 65535) cpt /= 2;
  const int nsplit = C / cpt;
  if (nsplit > 1) {
    e = cudaMemsetAsync(g_flow, 0, sizeof(float) * 2 * (size_t)px, stream);
    if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "warp_bwd_nchw: memset g_flow: %s", cudaGetErrorString(e));
  }
  const dim3 grid((unsigned)cdiv(W, 128), (unsigned)H, (unsigned)(B * nsplit));
  if (mode == QPWC_MODE_TF) {
    auto k = warp_bwd_nchw_kernel<QPWC_MODE_TF>;
    QPWC_LAUNCH(k, grid, 128, 0, stream, img, flow, g_out, g_img, g_flow, cpt, H, W, scale, nsplit, C);
  } else {
    auto k = warp_bwd_nchw_kernel<QPWC_MODE_TFA>;
    QPWC_LAUNCH(k, grid, 128, 0, stream, img, flow, g_out, g_img, g_flow, cpt, H, W, scale, nsplit, C);
  }
  return check_launch("warp_bwd_nchw");
}

// ------------------------------------------------------------------------------------------ bwd
// G lanes per pixel (power of two, <= 32).  Lane l of a group handles channel vectors l, l+G, ...
// grid = (ceil(W / groups_per_block), H, B): row and batch come from the block index (32-bit maths).
template <int MODE, int V>
__global__ void __launch_bounds__(256) warp_bwd_kernel(const float* __restrict__ img,
                                                       const float* __restrict__ flow,
                                                       const float* __restrict__ g_out,
                                                       float* __restrict__ g_img,
                                                       float* __restrict__ g_flow, int H, int W, int C,
                                                       int G, float scale, long long gops) {
  const int CV = C / V;
  const int lane = threadIdx.x & 31;
  const int gl = lane & (G - 1);           // lane inside the group
  const int groups_per_block = blockDim.x / G;
  const int j = blockIdx.x * groups_per_block + threadIdx.x / G;
  const int i = blockIdx.y;
  const bool live = j < W;                 // whole groups are live or not; shuffles below need the full warp
  const size_t boff = (size_t)blockIdx.z * H * W * C;
  const size_t pix = ((size_t)blockIdx.z * H + i) * W + (live ? j : 0);
  float gx = 0.f, gy = 0.f;
  bool px = true, py = true;
  if (live) {
    float2 f = __ldg(reinterpret_cast<const float2*>(flow) + pix);
    f.x = __fmul_rn(scale, f.x); f.y = __fmul_rn(scale, f.y);  // the flow the forward pass sampled with
    Taps t;
    if (MODE == QPWC_MODE_TF) t = taps_tf(i, j, f.x, f.y, H, W);
    else t = taps_tfa(i, j, f.x, f.y, H, W, &px, &py);
    // mode TF: the 1-D factors of the weights, recomputed exactly as taps_tf does
    float ax1 = 0.f, ax0 = 0.f, ay1 = 0.f, ay0 = 0.f;
    if (MODE == QPWC_MODE_TF) {
      const float x = __fadd_rn((float)j, f.x), y = __fadd_rn((float)i, f.y);
      const int y0 = t.o00 / W, x0 = t.o00 - y0 * W, y1 = t.o11 / W, x1 = t.o11 - y1 * W;
      ax1 = __fsub_rn((float)x1, x); ax0 = __fsub_rn(x, (float)x0);
      ay1 = __fsub_rn((float)y1, y); ay0 = __fsub_rn(y, (float)y0);
    }
    const bool dupx = (MODE == QPWC_MODE_TF) && (t.o00 == t.o01);
    const bool dupy = (MODE == QPWC_MODE_TF) && (t.o00 == t.o10);
    for (int cv = gl; cv < CV; cv += G) {
      const size_t co = (size_t)cv * V;
      float v00[V], v01[V], v10[V], v11[V], g[V], a00[V], a01[V], a10[V], a11[V];
      vload<V>(img + boff + (size_t)t.o00 * C + co, v00);
      vload<V>(img + boff + (size_t)t.o01 * C + co, v01);
      vload<V>(img + boff + (size_t)t.o10 * C + co, v10);
      vload<V>(img + boff + (size_t)t.o11 * C + co, v11);
      vload<V>(g_out + pix * (size_t)gops + co, g);
#pragma unroll
      for (int k = 0; k < V; ++k) {
        if (MODE == QPWC_MODE_TF) {
          // gather_nd grad: scatter w*g into the four clipped taps; d/dx = g*[-(y1-y)Ia - (y-y0)Ib +
          // (y1-y)Ic + (y-y0)Id], d/dy analogous (Ia=v00 Ib=v10 Ic=v01 Id=v11).  Un-contracted on
          // purpose: clipped taps (Ia == Ic, ...) must cancel exactly as in TF.
          a00[k] = __fmul_rn(t.w00, g[k]); a10[k] = __fmul_rn(t.w10, g[k]);
          a01[k] = __fmul_rn(t.w01, g[k]); a11[k] = __fmul_rn(t.w11, g[k]);
          const float sx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(-ay1, v00[k]), __fmul_rn(-ay0, v10[k])),
                                               __fmul_rn(ay1, v01[k])), __fmul_rn(ay0, v11[k]));
          const float sy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(-ax1, v00[k]), __fmul_rn(ax1, v10[k])),
                                               __fmul_rn(-ax0, v01[k])), __fmul_rn(ax0, v11[k]));
          gx = __fadd_rn(gx, __fmul_rn(g[k], sx));
          gy = __fadd_rn(gy, __fmul_rn(g[k], sy));
        } else {
          // every product/sum rounded on its own, like the TF gradient graph (and the oracle)
          const float ax = t.w00, ay = t.w01;
          const float top = __fadd_rn(__fmul_rn(ax, __fsub_rn(v01[k], v00[k])), v00[k]);
          const float bot = __fadd_rn(__fmul_rn(ax, __fsub_rn(v11[k], v10[k])), v10[k]);
          gy = __fadd_rn(gy, __fmul_rn(g[k], __fsub_rn(bot, top)));
          const float g_bot = __fmul_rn(ay, g[k]);
          const float g_top = __fsub_rn(g[k], g_bot);
          gx = __fadd_rn(gx, __fadd_rn(__fmul_rn(g_top, __fsub_rn(v01[k], v00[k])),
                                       __fmul_rn(g_bot, __fsub_rn(v11[k], v10[k]))));
          const float g_tr = __fmul_rn(ax, g_top), g_br = __fmul_rn(ax, g_bot);
          a01[k] = g_tr; a00[k] = __fsub_rn(g_top, g_tr); a11[k] = g_br; a10[k] = __fsub_rn(g_bot, g_br);
        }
      }
      // Clipped taps coincide (mode TF, sample outside the image): fold them in registers first.
      // Their weights are exact negatives of each other, so the fold cancels exactly -- as the
      // reference's sequential scatter does -- instead of leaving +-|w*g| rounding residue in an
      // order-dependent atomic sum; it also saves the redundant atomics.
      if (dupx) {
#pragma unroll
        for (int k = 0; k < V; ++k) { a00[k] = __fadd_rn(a00[k], a01[k]); a10[k] = __fadd_rn(a10[k], a11[k]); }
      }
      if (dupy) {
#pragma unroll
        for (int k = 0; k < V; ++k) { a00[k] = __fadd_rn(a00[k], a10[k]); a01[k] = __fadd_rn(a01[k], a11[k]); }
      }
      vatomic_add<V>(g_img + boff + (size_t)t.o00 * C + co, a00);
      if (!dupx) vatomic_add<V>(g_img + boff + (size_t)t.o01 * C + co, a01);
      if (!dupy) vatomic_add<V>(g_img + boff + (size_t)t.o10 * C + co, a10);
      if (!dupx && !dupy) vatomic_add<V>(g_img + boff + (size_t)t.o11 * C + co, a11);
    }
  }
  // fixed-order butterfly over the G lanes of the group
  for (int o = G >> 1; o > 0; o >>= 1) {
    gx += __shfl_xor_sync(0xffffffffu, gx, o);
    gy += __shfl_xor_sync(0xffffffffu, gy, o);
  }
  // chain rule through scale * flow: one more rounded multiply, as TF's gradient of the Mul
  if (live && gl == 0)
    reinterpret_cast<float2*>(g_flow)[pix] = make_float2(px ? __fmul_rn(scale, gx) : 0.f, py ? __fmul_rn(scale, gy) : 0.f);
}


// ----------------------------------------------------------------- bwd, shared-memory pre-aggregation
// Scatter-add with per-tile accumulation in shared memory before the global atomics (north_star:
// "shared-memory pre-aggregation before global atomics").  A CTA owns a tile of TBH x TBW output pixels
// and one chunk of 32 channels; the image gradient of the tile and a margin of TBM pixels around it is
// accumulated with shared-memory atomics (flows of a few pixels land inside the window; anything further
// goes straight to global memory), then the window is flushed with one vector `red.global.add.v4.f32`
// per 16 bytes that received anything -- about (TBH+2M)(TBW+2M)/(4*TBH*TBW) of the direct kernel's
// global atomics.  Per-element arithmetic is that of warp_bwd_kernel; the flow gradient is reduced over
// the 8 lanes of a pixel in a fixed order and (C > 32) across channel chunks by atomics on a pre-zeroed
// g_flow.
// MEASURED (tools/ab_warp_bwd.py, B200, B=8): 400 us vs 158 us for the direct kernel at 224x512x32, 2.5-3x
// slower at every level, smooth and noisy flows alike: shared-memory float atomics retire about one lane
// per clock per SM (117 M lane-atomics / 148 SMs = 0.79 M clk = 400 us), while `red.global.add.v4.f32`
// is executed 16 bytes at a time by the L2 slices.  Pre-aggregation in shared memory is therefore NOT
// the default on this part; the kernel stays selectable (QPWC_OPT_WARP_BWD = 2) as the measured
// alternative and is parity-tested.
#ifndef QPWC_EMU   // (the CPU emulation build keeps the direct kernel)
#define TBH 16
#define TBW 32
#define TBM 3
#define TBCK 32
#define TB_WH (TBH + 2 * TBM)
#define TB_WW (TBW + 2 * TBM)
template <int MODE>
__global__ void __launch_bounds__(256) warp_bwd_tile_kernel(const float* __restrict__ img, const float* __restrict__ flow,
                                                           const float* __restrict__ g_out, float* __restrict__ g_img,
                                                           float* __restrict__ g_flow, int H, int W, int C, int nchunks,
                                                           float scale, long long gops) {
  extern __shared__ __align__(16) float acc[];   // [TB_WH][TB_WW][TBCK]
  const int tid = threadIdx.x;
  const int b = (int)blockIdx.z / nchunks, ck = (int)blockIdx.z - b * nchunks;
  const int i0 = (int)blockIdx.y * TBH, j0 = (int)blockIdx.x * TBW;
  const int wi0 = i0 - TBM, wj0 = j0 - TBM;                 // window origin (may be negative)
  for (int e = tid; e < TB_WH * TB_WW * TBCK / 4; e += 256) reinterpret_cast<float4*>(acc)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  __syncthreads();
  const int cq = tid & 7;                                    // channel quad inside the chunk
  const size_t boff = (size_t)b * H * W * C;
  const int c0 = ck * TBCK + cq * 4;
  for (int pp = tid >> 3; pp < TBH * TBW; pp += 32) {
    const int i = i0 + pp / TBW, j = j0 + pp % TBW;
    const bool live = i < H && j < W;                        // the 8 lanes of a pixel agree
    float gx = 0.f, gy = 0.f;
    bool px = true, py = true;
    size_t pix = 0;
    if (live) {
      pix = ((size_t)b * H + i) * W + j;
      float2 f = __ldg(reinterpret_cast<const float2*>(flow) + pix);
      f.x = __fmul_rn(scale, f.x); f.y = __fmul_rn(scale, f.y);
      Taps t;
      if (MODE == QPWC_MODE_TF) t = taps_tf(i, j, f.x, f.y, H, W);
      else t = taps_tfa(i, j, f.x, f.y, H, W, &px, &py);
      float ax1 = 0.f, ax0 = 0.f, ay1 = 0.f, ay0 = 0.f;
      if (MODE == QPWC_MODE_TF) {
        const float x = __fadd_rn((float)j, f.x), y = __fadd_rn((float)i, f.y);
        const int y0 = t.o00 / W, x0 = t.o00 - y0 * W, y1 = t.o11 / W, x1 = t.o11 - y1 * W;
        ax1 = __fsub_rn((float)x1, x); ax0 = __fsub_rn(x, (float)x0);
        ay1 = __fsub_rn((float)y1, y); ay0 = __fsub_rn(y, (float)y0);
      }
      const bool dupx = (MODE == QPWC_MODE_TF) && (t.o00 == t.o01);
      const bool dupy = (MODE == QPWC_MODE_TF) && (t.o00 == t.o10);
      float v00[4], v01[4], v10[4], v11[4], g[4], a00[4], a01[4], a10[4], a11[4];
      vload<4>(img + boff + (size_t)t.o00 * C + c0, v00);
      vload<4>(img + boff + (size_t)t.o01 * C + c0, v01);
      vload<4>(img + boff + (size_t)t.o10 * C + c0, v10);
      vload<4>(img + boff + (size_t)t.o11 * C + c0, v11);
      vload<4>(g_out + pix * (size_t)gops + c0, g);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (MODE == QPWC_MODE_TF) {
          a00[k] = __fmul_rn(t.w00, g[k]); a10[k] = __fmul_rn(t.w10, g[k]);
          a01[k] = __fmul_rn(t.w01, g[k]); a11[k] = __fmul_rn(t.w11, g[k]);
          const float sx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(-ay1, v00[k]), __fmul_rn(-ay0, v10[k])),
                                               __fmul_rn(ay1, v01[k])), __fmul_rn(ay0, v11[k]));
          const float sy = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(-ax1, v00[k]), __fmul_rn(ax1, v10[k])),
                                               __fmul_rn(-ax0, v01[k])), __fmul_rn(ax0, v11[k]));
          gx = __fadd_rn(gx, __fmul_rn(g[k], sx));
          gy = __fadd_rn(gy, __fmul_rn(g[k], sy));
        } else {
          const float ax = t.w00, ay = t.w01;
          const float top = __fadd_rn(__fmul_rn(ax, __fsub_rn(v01[k], v00[k])), v00[k]);
          const float bot = __fadd_rn(__fmul_rn(ax, __fsub_rn(v11[k], v10[k])), v10[k]);
          gy = __fadd_rn(gy, __fmul_rn(g[k], __fsub_rn(bot, top)));
          const float g_bot = __fmul_rn(ay, g[k]);
          const float g_top = __fsub_rn(g[k], g_bot);
          gx = __fadd_rn(gx, __fadd_rn(__fmul_rn(g_top, __fsub_rn(v01[k], v00[k])),
                                       __fmul_rn(g_bot, __fsub_rn(v11[k], v10[k]))));
          const float g_tr = __fmul_rn(ax, g_top), g_br = __fmul_rn(ax, g_bot);
          a01[k] = g_tr; a00[k] = __fsub_rn(g_top, g_tr); a11[k] = g_br; a10[k] = __fsub_rn(g_bot, g_br);
        }
      }
      if (dupx) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { a00[k] = __fadd_rn(a00[k], a01[k]); a10[k] = __fadd_rn(a10[k], a11[k]); }
      }
      if (dupy) {
#pragma unroll
        for (int k = 0; k < 4; ++k) { a00[k] = __fadd_rn(a00[k], a10[k]); a01[k] = __fadd_rn(a01[k], a11[k]); }
      }
      // scatter: into the shared window when the tap lies inside it, else straight to global memory
      auto scatter = [&](int o, const float (&a)[4]) {
        const int ty = o / W, tx = o - ty * W, wy = ty - wi0, wx = tx - wj0;
        if ((unsigned)wy < (unsigned)TB_WH && (unsigned)wx < (unsigned)TB_WW) {
          float* d = acc + ((wy * TB_WW + wx) * TBCK + cq * 4);
#pragma unroll
          for (int k = 0; k < 4; ++k) atomicAdd(d + k, a[k]);
        } else {
          vatomic_add<4>(g_img + boff + (size_t)o * C + c0, a);
        }
      };
      scatter(t.o00, a00);
      if (!dupx) scatter(t.o01, a01);
      if (!dupy) scatter(t.o10, a10);
      if (!dupx && !dupy) scatter(t.o11, a11);
    }
    // flow gradient: fixed-order butterfly over the 8 lanes of the pixel (lanes 8k..8k+7 of a warp)
#pragma unroll
    for (int o = 4; o > 0; o >>= 1) {
      gx += __shfl_xor_sync(0xffffffffu, gx, o);
      gy += __shfl_xor_sync(0xffffffffu, gy, o);
    }
    if (live && cq == 0) {
      const float fx = px ? __fmul_rn(scale, gx) : 0.f, fy = py ? __fmul_rn(scale, gy) : 0.f;
      if (nchunks == 1) reinterpret_cast<float2*>(g_flow)[pix] = make_float2(fx, fy);
      else { atomicAdd(g_flow + 2 * pix, fx); atomicAdd(g_flow + 2 * pix + 1, fy); }
    }
  }
  __syncthreads();
  // flush the window: one vector atomic per 16 bytes that received anything
  for (int e = tid; e < TB_WH * TB_WW * TBCK / 4; e += 256) {
    const int q4 = e & 7, cell = e >> 3, wy = cell / TB_WW, wx = cell - wy * TB_WW;
    const int ty = wi0 + wy, tx = wj0 + wx;
    if (ty < 0 || ty >= H || tx < 0 || tx >= W) continue;
    const float4 v = reinterpret_cast<const float4*>(acc)[e];
    if (v.x == 0.f && v.y == 0.f && v.z == 0.f && v.w == 0.f) continue;
    const float a[4] = {v.x, v.y, v.z, v.w};
    vatomic_add<4>(g_img + boff + ((size_t)ty * W + tx) * C + ck * TBCK + q4 * 4, a);
  }
}

#endif  // !QPWC_EMU

static std::atomic<int> g_warp_bwd_variant{0};   // 0 auto (= direct), 1 direct (register-folded global atomics), 2 shared-memory tiles
void set_warp_bwd_variant(int v) { g_warp_bwd_variant.store(v); }
int get_warp_bwd_variant() { return g_warp_bwd_variant.load(); }

// ------------------------------------------------------------------------------------ launchers
static inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

static int pick_vec(int C, const void* a, const void* b, const void* c = nullptr, const void* d = nullptr) {
  const void* ps[4] = {a, b, c, d};
  int v = (C % 4 == 0) ? 4 : (C % 2 == 0 ? 2 : 1);
  for (int k = 0; k < 4; ++k)
    if (ps[k]) while (v > 1 && !aligned(ps[k], sizeof(float) * v)) v >>= 1;
  return v;
}

template <int MODE, int V, int NV>
static void run_warp_fwd_nv(const float* img, const float* flow, const float* img2, const float* flow2,
                            float* out, int B, int H, int W, int C, float scale, long long ops,
                            float up_scale, cudaStream_t stream, int row_off, int Hfull, int keep);

template <int MODE, int V>
static void run_warp_fwd(const float* img, const float* flow, const float* img2, const float* flow2,
                         float* out, int B, int H, int W, int C, float scale, long long ops,
                         float up_scale, cudaStream_t stream, int row_off, int Hfull, int keep) {
  const int block = 256;
  if (V == 4 && C % 8 == 0) {  // two 16-byte vectors per thread
    run_warp_fwd_nv<MODE, V, 2>(img, flow, img2, flow2, out, B, H, W, C, scale, ops, up_scale, stream, row_off, Hfull, keep);
    return;
  }
  run_warp_fwd_nv<MODE, V, 1>(img, flow, img2, flow2, out, B, H, W, C, scale, ops, up_scale, stream, row_off, Hfull, keep);
}

template <int MODE, int V, int NV>
static void run_warp_fwd_nv(const float* img, const float* flow, const float* img2, const float* flow2,
                            float* out, int B, int H, int W, int C, float scale, long long ops,
                            float up_scale, cudaStream_t stream, int row_off, int Hfull, int keep) {
  const int block = 256;
  const int CV = C / (V * NV);
  auto k = warp_fwd_kernel<MODE, V, NV>;
  if (img2) {  // pair: grid.z = 2B (B <= 32767 checked by the caller)
    const dim3 grid((unsigned)cdiv(W * CV, block / WARP_FWD_ROWS), (unsigned)cdiv(H, WARP_FWD_ROWS), (unsigned)(2 * B));
    QPWC_LAUNCH(k, grid, block, 0, stream, img, flow, img2, flow2, out, H, W, C, B, scale, ops, up_scale, row_off, Hfull, keep);
    return;
  }
  // gridDim.y/z are limited to 65535: chunk the batch (and refuse absurd heights upstream)
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = (B - b0 < 65535) ? (B - b0) : 65535;
    const dim3 grid((unsigned)cdiv(W * CV, block / WARP_FWD_ROWS), (unsigned)cdiv(H, WARP_FWD_ROWS), (unsigned)nb);
    const size_t off = (size_t)b0 * H * W;
    const size_t foff = up_scale != 0.f ? (size_t)b0 * (H / 2) * (W / 2) * 2 : off * 2;
    QPWC_LAUNCH(k, grid, block, 0, stream, img + off * C, flow + foff, img2, flow2, out + off * ops, H, W, C,
                nb, scale, ops, up_scale, row_off, Hfull, keep);
  }
}

// img2/flow2 != nullptr: two warps in one launch, the second writing channels [C, 2C) of each pixel
int launch_warp_fwd_ex(const float* img, const float* flow, const float* img2, const float* flow2,
                       float* out, int B, int H, int W, int C, int mode, float scale, long long ops,
                       cudaStream_t stream, float up_scale, int row_off, int Hfull, int keep_l2) {
  if (Hfull <= 0) Hfull = H;
  int V = pick_vec(C, img, out, img2);
  while (V > 1 && ops % V) V >>= 1;
  if ((long long)B * H * W * C == 0) return QPWC_OK;
  if (H > 65535 || (long long)W * (C / V) >= (1LL << 31)) return set_error(QPWC_ERR_UNSUPPORTED, "warp_fwd: H > 65535 or W*C too large");
  if (img2 && B > 32767) return set_error(QPWC_ERR_UNSUPPORTED, "warp_pair_fwd: B > 32767");
  if (mode == QPWC_MODE_TF) {
    if (V == 4) run_warp_fwd<QPWC_MODE_TF, 4>(img, flow, img2, flow2, out, B, H, W, C, scale, ops, up_scale, stream, row_off, Hfull, keep_l2);
    else if (V == 2) run_warp_fwd<QPWC_MODE_TF, 2>(img, flow, img2, flow2, out, B, H, W, C, scale, ops, up_scale, stream, row_off, Hfull, keep_l2);
    else run_warp_fwd<QPWC_MODE_TF, 1>(img, flow, img2, flow2, out, B, H, W, C, scale, ops, up_scale, stream, row_off, Hfull, keep_l2);
  } else {
    if (V == 4) run_warp_fwd<QPWC_MODE_TFA, 4>(img, flow, img2, flow2, out, B, H, W, C, scale, ops, up_scale, stream, row_off, Hfull, keep_l2);
    else if (V == 2) run_warp_fwd<QPWC_MODE_TFA, 2>(img, flow, img2, flow2, out, B, H, W, C, scale, ops, up_scale, stream, row_off, Hfull, keep_l2);
    else run_warp_fwd<QPWC_MODE_TFA, 1>(img, flow, img2, flow2, out, B, H, W, C, scale, ops, up_scale, stream, row_off, Hfull, keep_l2);
  }
  return check_launch("warp_fwd");
}

int launch_warp_fwd(const float* img, const float* flow, float* out, int B, int H, int W, int C,
                    int mode, cudaStream_t stream) {
  return launch_warp_fwd_ex(img, flow, nullptr, nullptr, out, B, H, W, C, mode, 1.f, C, stream, 0.f, 0, 0, 0);
}

template <int MODE, int V>
static void run_warp_bwd(const float* img, const float* flow, const float* g_out, float* g_img,
                         float* g_flow, int B, int H, int W, int C, float scale, long long gops,
                         cudaStream_t stream) {
  const int CV = C / V;
  int G = 1;
  while (G < CV && G < 32) G <<= 1;
  const int block = 256;
  const int groups_per_block = block / G;
  auto k = warp_bwd_kernel<MODE, V>;
  for (int b0 = 0; b0 < B; b0 += 65535) {
    const int nb = (B - b0 < 65535) ? (B - b0) : 65535;
    const dim3 grid((unsigned)cdiv(W, groups_per_block), (unsigned)H, (unsigned)nb);
    const size_t off = (size_t)b0 * H * W;
    QPWC_LAUNCH(k, grid, block, 0, stream, img + off * C, flow + off * 2, g_out + off * gops, g_img + off * C,
                g_flow + off * 2, H, W, C, G, scale, gops);
  }
}

int launch_warp_bwd_ex(const float* img, const float* flow, const float* g_out, float* g_img,
                       float* g_flow, int B, int H, int W, int C, int mode, float scale,
                       long long gops, cudaStream_t stream) {
  const long long npix = (long long)B * H * W;
  if (npix == 0) return QPWC_OK;
  cudaError_t e = cudaMemsetAsync(g_img, 0, sizeof(float) * (size_t)npix * C, stream);
  if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "warp_bwd: memset g_img: %s", cudaGetErrorString(e));
  int V = pick_vec(C, img, g_out, g_img);
  while (V > 1 && gops % V) V >>= 1;
  if (H > 65535) return set_error(QPWC_ERR_UNSUPPORTED, "warp_bwd: H > 65535");
#ifndef QPWC_EMU
  const int variant = g_warp_bwd_variant.load(std::memory_order_relaxed);
  if (variant == 2 && V == 4 && C % TBCK == 0 && (long long)B * (C / TBCK) <= 65535 && cdiv(H, TBH) <= 65535) {
    const int nchunks = C / TBCK;
    if (nchunks > 1) {
      e = cudaMemsetAsync(g_flow, 0, sizeof(float) * 2 * (size_t)npix, stream);
      if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "warp_bwd: memset g_flow: %s", cudaGetErrorString(e));
    }
    const int smem = TB_WH * TB_WW * TBCK * 4;
    static std::atomic<unsigned> attr_done{0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (!(attr_done.load(std::memory_order_acquire) >> (dev & 31) & 1u)) {
      cudaFuncSetAttribute(warp_bwd_tile_kernel<QPWC_MODE_TF>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      cudaFuncSetAttribute(warp_bwd_tile_kernel<QPWC_MODE_TFA>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
      attr_done.fetch_or(1u << (dev & 31), std::memory_order_release);
    }
    const dim3 grid((unsigned)cdiv(W, TBW), (unsigned)cdiv(H, TBH), (unsigned)(B * nchunks));
    if (mode == QPWC_MODE_TF) warp_bwd_tile_kernel<QPWC_MODE_TF><<<grid, 256, smem, stream>>>(img, flow, g_out, g_img, g_flow, H, W, C, nchunks, scale, gops);
    else warp_bwd_tile_kernel<QPWC_MODE_TFA><<<grid, 256, smem, stream>>>(img, flow, g_out, g_img, g_flow, H, W, C, nchunks, scale, gops);
    return check_launch("warp_bwd_tile");
  }
#endif
  if (mode == QPWC_MODE_TF) {
    if (V == 4) run_warp_bwd<QPWC_MODE_TF, 4>(img, flow, g_out, g_img, g_flow, B, H, W, C, scale, gops, stream);
    else if (V == 2) run_warp_bwd<QPWC_MODE_TF, 2>(img, flow, g_out, g_img, g_flow, B, H, W, C, scale, gops, stream);
    else run_warp_bwd<QPWC_MODE_TF, 1>(img, flow, g_out, g_img, g_flow, B, H, W, C, scale, gops, stream);
  } else {
    if (V == 4) run_warp_bwd<QPWC_MODE_TFA, 4>(img, flow, g_out, g_img, g_flow, B, H, W, C, scale, gops, stream);
    else if (V == 2) run_warp_bwd<QPWC_MODE_TFA, 2>(img, flow, g_out, g_img, g_flow, B, H, W, C, scale, gops, stream);
    else run_warp_bwd<QPWC_MODE_TFA, 1>(img, flow, g_out, g_img, g_flow, B, H, W, C, scale, gops, stream);
  }
  return check_launch("warp_bwd");
}

int launch_warp_bwd(const float* img, const float* flow, const float* g_out, float* g_img,
                    float* g_flow, int B, int H, int W, int C, int mode, cudaStream_t stream) {
  return launch_warp_bwd_ex(img, flow, g_out, g_img, g_flow, B, H, W, C, mode, 1.f, C, stream);
}

}  // namespace qpwc
