// qpwc_upsample.cuh -- the x2 bilinear flow upsampling that feeds the warp in the reference:
//   Upsample(scale=2.0):  tf.constant(scale) * UpSampling2D(interpolation='bilinear')(x)
//   (qpwcnet/core/non_layers.py:183-193; used on the flow before every UpFlow, core/pwcnet.py:49-56)
// UpSampling2D(bilinear) is tf.image.resize(..., 'bilinear') with half-pixel centres:
//   in = (o + 0.5) * 0.5 - 0.5;  lo = max(floor(in), 0);  hi = min(ceil(in), n-1);  lerp = in - floor(in)
//   top = tl + (tr - tl) * xl;  bot = bl + (br - bl) * xl;  out = top + (bot - top) * yl
// (TF ResizeBilinear kernel, restated from its published algorithm -- TF is not installable here, so
// parity for this op is UNPINNED; tests cross-check against torch.nn.functional.interpolate, which
// implements the same half-pixel rule.)  Every product/sum is rounded on its own.
#pragma once
#include "qpwc_common.cuh"

namespace qpwc {

struct Up2 { int lo, hi; float lerp; };

__device__ __forceinline__ Up2 up2_coord(int o, int n_in) {
  const float in = __fsub_rn(__fmul_rn(__fadd_rn((float)o, 0.5f), 0.5f), 0.5f);
  const float fl = floorf(in);
  Up2 u;
  u.lo = max((int)fl, 0);
  u.hi = min((int)ceilf(in), n_in - 1);
  u.lerp = __fsub_rn(in, fl);
  return u;
}
__device__ __forceinline__ float up2_blend(float tl, float tr, float bl, float br, float xl, float yl) {
  const float top = __fadd_rn(tl, __fmul_rn(__fsub_rn(tr, tl), xl));
  const float bot = __fadd_rn(bl, __fmul_rn(__fsub_rn(br, bl), xl));
  return __fadd_rn(top, __fmul_rn(__fsub_rn(bot, top), yl));
}
// scale * upsampled coarse flow at output pixel (i, j); coarse flow of one batch item: (Hc, Wc, 2)
__device__ __forceinline__ float2 up2_flow(const float* __restrict__ flow_c, int i, int j, int Hc, int Wc, float scale) {
  const Up2 y = up2_coord(i, Hc), x = up2_coord(j, Wc);
  const float2* f = reinterpret_cast<const float2*>(flow_c);
  const float2 tl = __ldg(f + (size_t)y.lo * Wc + x.lo), tr = __ldg(f + (size_t)y.lo * Wc + x.hi);
  const float2 bl = __ldg(f + (size_t)y.hi * Wc + x.lo), br = __ldg(f + (size_t)y.hi * Wc + x.hi);
  return make_float2(__fmul_rn(scale, up2_blend(tl.x, tr.x, bl.x, br.x, x.lerp, y.lerp)),
                     __fmul_rn(scale, up2_blend(tl.y, tr.y, bl.y, br.y, x.lerp, y.lerp)));
}

}  // namespace qpwc
