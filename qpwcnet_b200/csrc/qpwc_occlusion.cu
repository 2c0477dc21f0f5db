// qpwc_occlusion.cu -- estimate_occlusion_map (qpwcnet/core/occlusion.py:27-118) as two small
// launches on one stream.  flow is (B,H,W,2) or (B,2,H,W) fp32 with channel 0 = dx, 1 = dy
// (occlusion.py:59); the map is (B,H,W) fp32, 1 = "no value in the next frame".
//
//   map3 = 1                                                        (occlusion.py:93, ones_like)
//   inv  = -tf_warp(flow, flow);  map3[b, clip(int(i + inv_y)), clip(int(j + inv_x))] = 0
//                                                                    (occlusion.py:83-93; every
//        update of tensor_scatter_nd_min is 0, so the scatter is order-independent: a plain store)
//   out  = max(oob, map3),  oob = i+dy < 0 | i+dy >= H | j+dx < 0 | j+dx >= W   (occlusion.py:74,96)
//
// The warp of the 2-channel flow by itself is evaluated in registers (taps_tf/blend_tf: the same
// arithmetic as the stand-alone warp kernel), so `inv` never reaches HBM: 8 B read + 4 gathered
// 8-byte taps + the landing site's flow (L2 hits) + 4 B written per pixel.
#include "qpwc_common.cuh"

namespace qpwc {

__global__ void __launch_bounds__(256) occl_fill_kernel(float* __restrict__ out, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if ((reinterpret_cast<uintptr_t>(out) & 15) == 0) {
    const long long n4 = total >> 2;
    for (long long q = t0; q < n4; q += stride) reinterpret_cast<float4*>(out)[q] = make_float4(1.f, 1.f, 1.f, 1.f);
    for (long long idx = (n4 << 2) + t0; idx < total; idx += stride) out[idx] = 1.f;
  } else {
    for (long long idx = t0; idx < total; idx += stride) out[idx] = 1.f;
  }
}

// One block = 256 consecutive pixels of one image row: (b, i, segment) from one 32-bit division.
// NHWC: a pixel's (dx, dy) is one float2;  NCHW: two planes H*W apart.
// A landing site is cleared only if its own flow stays inside the image -- out = max(oob, map3) is 1
// there whatever map3 says (occlusion.py:96) -- so no third pass over the map is needed.
template <bool NHWC>
__global__ void __launch_bounds__(256) occl_mark_kernel(const float* __restrict__ flow, float* __restrict__ out,
                                                        int H, int W, unsigned segs) {
  const unsigned row = blockIdx.x / segs;                       // b * H + i
  const int j = (int)((blockIdx.x - row * segs) * 256u + threadIdx.x);
  if (j >= W) return;
  const unsigned b = row / (unsigned)H;
  const int i = (int)(row - b * (unsigned)H);
  const size_t plane = (size_t)H * W;
  const float* f = flow + (size_t)b * plane * 2;
  auto px = [&](int o, float& vx, float& vy) {
    if (NHWC) {
      const float2 v = __ldg(reinterpret_cast<const float2*>(f) + o);
      vx = v.x; vy = v.y;
    } else {
      vx = __ldg(f + o); vy = __ldg(f + plane + o);
    }
  };
  float fx, fy;
  px(i * W + j, fx, fy);
  const Taps t = taps_tf(i, j, fx, fy, H, W);
  float ax, ay, bx, by, cx, cy, dx, dy;
  px(t.o00, ax, ay); px(t.o01, bx, by); px(t.o10, cx, cy); px(t.o11, dx, dy);
  const float ix = -blend_tf(t, ax, bx, cx, dx);
  const float iy = -blend_tf(t, ay, by, cy, dy);
  int i3 = __float2int_rz(__fadd_rn((float)i, iy));
  int j3 = __float2int_rz(__fadd_rn((float)j, ix));
  i3 = min(max(i3, 0), H - 1);
  j3 = min(max(j3, 0), W - 1);
  // oob of the landing site (occlusion.py:74)
  float gx, gy;
  px(i3 * W + j3, gx, gy);
  const float i2 = __fadd_rn((float)i3, gy), j2 = __fadd_rn((float)j3, gx);
  const bool oob = i2 < 0.f || i2 >= (float)H || j2 < 0.f || j2 >= (float)W;
  if (!oob) out[(size_t)b * plane + (size_t)i3 * W + j3] = 0.f;
}

int launch_occlusion_map(const float* flow, float* out, int B, int H, int W, int channels_first, cudaStream_t stream) {
  const long long total = (long long)B * H * W;
  if (total == 0) return QPWC_OK;
  const unsigned segs = (unsigned)cdiv(W, 256);
  const long long blocks = (long long)B * H * segs;
  if (blocks > 0x7fffffffLL) return set_error(QPWC_ERR_UNSUPPORTED, "occlusion_map: B*H*ceil(W/256) exceeds 2^31-1");
  {
    const long long want = cdivll(cdivll(total, 4), 256);
    auto k = occl_fill_kernel;
    QPWC_LAUNCH(k, (int)(want < 148LL * 16 ? want : 148LL * 16), 256, 0, stream, out, total);
  }
  if (channels_first) {
    auto k1 = occl_mark_kernel<false>;
    QPWC_LAUNCH(k1, (unsigned)blocks, 256, 0, stream, flow, out, H, W, segs);
  } else {
    auto k1 = occl_mark_kernel<true>;
    QPWC_LAUNCH(k1, (unsigned)blocks, 256, 0, stream, flow, out, H, W, segs);
  }
  return check_launch("occlusion_map");
}

}  // namespace qpwc
