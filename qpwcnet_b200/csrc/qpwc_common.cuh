// qpwc_common.cuh -- shared device helpers for the qpwc CUDA library (sm_100a).
//
// Build modes
//   default      : nvcc, real CUDA (the product).
//   -DQPWC_EMU   : g++ with tests/emu/cuda_emu.h force-included.  TEST-ONLY harness that runs the
//                  very same kernel bodies thread-by-thread on the CPU so that index arithmetic /
//                  tiling logic can be checked against the oracle in the no-GPU container.  The
//                  product never builds or loads that variant (see tests/emu/README.md).
#pragma once

#ifndef QPWC_EMU
#include <cuda_runtime.h>
#endif
#include <stdint.h>

#define QPWC_OK 0
#define QPWC_ERR_INVALID 1
#define QPWC_ERR_UNSUPPORTED 2
#define QPWC_ERR_CUDA 3

#define QPWC_MODE_TF 0   // Warp / tf_warp      (qpwcnet/core/warp.py:63-153)
#define QPWC_MODE_TFA 1  // WarpV2 / tfa bilinear (qpwcnet/core/layers.py:171-186)

#ifdef QPWC_EMU
#define QPWC_LAUNCH(kernel, grid, block, smem, stream, ...) \
  qpwc_emu::launch((grid), (block), (smem), [=]() { kernel(__VA_ARGS__); })
#define QPWC_DYN_SMEM(name) unsigned char* name = qpwc_emu::dyn_smem()
#else
#define QPWC_LAUNCH(kernel, grid, block, smem, stream, ...) \
  kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define QPWC_DYN_SMEM(name) extern __shared__ __align__(1024) unsigned char name[]
#endif

namespace qpwc {

// error plumbing (qpwc_api.cu)
int set_error(int code, const char* fmt, ...);
int check_launch(const char* what);

__host__ __device__ __forceinline__ int cdiv(int a, int b) { return (a + b - 1) / b; }
__host__ __device__ __forceinline__ long long cdivll(long long a, long long b) { return (a + b - 1) / b; }

__device__ __forceinline__ float lrelu(float v, float slope) { return v > 0.f ? v : slope * v; }

// ---------------------------------------------------------------------------------------------
// Bilinear sampling set-up shared by the stand-alone warp kernels and the fused warp->correlation
// kernels, so that both produce bit-identical warped values.
//   o??  : element offsets (in pixels, i.e. to be multiplied by C) of the four taps inside one
//          batch item;  w?? : tap weights (mode TF)  or  ax/ay lerp factors (mode TFA).
// ---------------------------------------------------------------------------------------------
struct Taps {
  int o00, o01, o10, o11;  // (y0,x0) (y0,x1) (y1,x0) (y1,x1) pixel indices  y*W + x
  float w00, w01, w10, w11;  // TF : wa, wc, wb, wd   (weights of the taps above)
                             // TFA: w00 = ax, w01 = ay (others unused)
};

// mode TF -- qpwcnet/core/warp.py:100-142.  No fused multiply-add anywhere: TF rounds every op.
__device__ __forceinline__ Taps taps_tf(int i, int j, float fx, float fy, int H, int W) {
  const float x = __fadd_rn((float)j, fx);
  const float y = __fadd_rn((float)i, fy);
  int x0 = __float2int_rz(x);  // tf.cast -> truncation toward zero (saturating, NaN -> 0)
  int y0 = __float2int_rz(y);
  int x1 = (int)((unsigned)x0 + 1u);
  int y1 = (int)((unsigned)y0 + 1u);
  x0 = min(max(x0, 0), W - 1); x1 = min(max(x1, 0), W - 1);
  y0 = min(max(y0, 0), H - 1); y1 = min(max(y1, 0), H - 1);
  const float ax1 = __fsub_rn((float)x1, x), ax0 = __fsub_rn(x, (float)x0);
  const float ay1 = __fsub_rn((float)y1, y), ay0 = __fsub_rn(y, (float)y0);
  Taps t;
  t.o00 = y0 * W + x0; t.o01 = y0 * W + x1; t.o10 = y1 * W + x0; t.o11 = y1 * W + x1;
  t.w00 = __fmul_rn(ax1, ay1);  // wa on I[y0,x0]
  t.w10 = __fmul_rn(ax1, ay0);  // wb on I[y1,x0]
  t.w01 = __fmul_rn(ax0, ay1);  // wc on I[y0,x1]
  t.w11 = __fmul_rn(ax0, ay0);  // wd on I[y1,x1]
  return t;
}
// out = add_n([wa*Ia, wb*Ib, wc*Ic, wd*Id])  (warp.py:151), left to right
__device__ __forceinline__ float blend_tf(const Taps& t, float v00, float v01, float v10, float v11) {
  return __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(t.w00, v00), __fmul_rn(t.w10, v10)),
                             __fmul_rn(t.w01, v01)),
                   __fmul_rn(t.w11, v11));
}

// mode TFA -- tfa interpolate_bilinear semantics (see oracle/qpwc_oracle_body.inc, W2).
// pass_x / pass_y: does the gradient reach the flow (TF max/min tie rule: iff 0 < q-floor <= 1).
__device__ __forceinline__ Taps taps_tfa(int i, int j, float fx, float fy, int H, int W,
                                         bool* pass_x = nullptr, bool* pass_y = nullptr) {
  const float qy = __fsub_rn((float)i, -fy);
  const float qx = __fsub_rn((float)j, -fx);
  float fly = floorf(qy), flx = floorf(qx);
  fly = fly > 0.f ? fly : 0.f; fly = fly < (float)(H - 2) ? fly : (float)(H - 2);
  flx = flx > 0.f ? flx : 0.f; flx = flx < (float)(W - 2) ? flx : (float)(W - 2);
  const float ry = __fsub_rn(qy, fly), rx = __fsub_rn(qx, flx);
  float ay = ry > 0.f ? ry : 0.f; ay = ay < 1.f ? ay : 1.f;
  float ax = rx > 0.f ? rx : 0.f; ax = ax < 1.f ? ax : 1.f;
  const int iy = (int)fly, ix = (int)flx;
  Taps t;
  t.o00 = iy * W + ix; t.o01 = t.o00 + 1; t.o10 = t.o00 + W; t.o11 = t.o10 + 1;
  t.w00 = ax; t.w01 = ay; t.w10 = 0.f; t.w11 = 0.f;
  if (pass_x) *pass_x = (rx > 0.f) && (rx <= 1.f);
  if (pass_y) *pass_y = (ry > 0.f) && (ry <= 1.f);
  return t;
}
// top = ax*(TR-TL)+TL; bot = ax*(BR-BL)+BL; out = ay*(bot-top)+top
__device__ __forceinline__ float blend_tfa(const Taps& t, float tl, float tr, float bl, float br) {
  const float top = __fadd_rn(__fmul_rn(t.w00, __fsub_rn(tr, tl)), tl);
  const float bot = __fadd_rn(__fmul_rn(t.w00, __fsub_rn(br, bl)), bl);
  return __fadd_rn(__fmul_rn(t.w01, __fsub_rn(bot, top)), top);
}

template <int MODE>
__device__ __forceinline__ Taps make_taps(int i, int j, float fx, float fy, int H, int W) {
  if (MODE == QPWC_MODE_TF) return taps_tf(i, j, fx, fy, H, W);
  return taps_tfa(i, j, fx, fy, H, W);
}
template <int MODE>
__device__ __forceinline__ float blend(const Taps& t, float v00, float v01, float v10, float v11) {
  if (MODE == QPWC_MODE_TF) return blend_tf(t, v00, v01, v10, v11);
  return blend_tfa(t, v00, v01, v10, v11);
}

}  // namespace qpwc
