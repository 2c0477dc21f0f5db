// qpwc_upsample.cu -- stand-alone x2 bilinear upsampling (forward and its adjoint), NHWC fp32.
// See qpwc_upsample.cuh for the reference lines and the arithmetic.
#include "qpwc_upsample.cuh"

namespace qpwc {

// dst (B, 2H, 2W, C) = scale * resize_bilinear(src (B, H, W, C)); one thread per output element
__global__ void __launch_bounds__(256) upsample2x_fwd_kernel(const float* __restrict__ src, float* __restrict__ dst,
                                                             int H, int W, int C, float scale, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int c = (int)(idx % C);
    const long long p = idx / C;
    const int oj = (int)(p % (2 * W));
    const long long q = p / (2 * W);
    const int oi = (int)(q % (2 * H));
    const long long b = q / (2 * H);
    const Up2 y = up2_coord(oi, H), x = up2_coord(oj, W);
    const float* s = src + (size_t)b * H * W * C + c;
    const float tl = __ldg(s + ((size_t)y.lo * W + x.lo) * C), tr = __ldg(s + ((size_t)y.lo * W + x.hi) * C);
    const float bl = __ldg(s + ((size_t)y.hi * W + x.lo) * C), br = __ldg(s + ((size_t)y.hi * W + x.hi) * C);
    dst[idx] = __fmul_rn(scale, up2_blend(tl, tr, bl, br, x.lerp, y.lerp));
  }
}

// adjoint: g_src[b,i,j,c] = scale * sum over the output pixels that sampled (i,j) of weight * g_dst.
// Gather form (deterministic): the outputs that can touch input row i are oi in [2i-2, 2i+3].
__global__ void __launch_bounds__(256) upsample2x_bwd_kernel(const float* __restrict__ g_dst, float* __restrict__ g_src,
                                                             int H, int W, int C, float scale, long long total) {
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int c = (int)(idx % C);
    const long long p = idx / C;
    const int j = (int)(p % W);
    const long long q = p / W;
    const int i = (int)(q % H);
    const long long b = q / H;
    const float* g = g_dst + (size_t)b * 4 * H * W * C + c;
    float acc = 0.f;
    for (int oi = max(2 * i - 2, 0); oi <= min(2 * i + 3, 2 * H - 1); ++oi) {
      const Up2 y = up2_coord(oi, H);
      float wy = 0.f;
      if (y.lo == i) wy = __fadd_rn(wy, __fsub_rn(1.f, y.lerp));
      if (y.hi == i) wy = __fadd_rn(wy, y.lerp);
      if (wy == 0.f) continue;
      for (int oj = max(2 * j - 2, 0); oj <= min(2 * j + 3, 2 * W - 1); ++oj) {
        const Up2 x = up2_coord(oj, W);
        float wx = 0.f;
        if (x.lo == j) wx = __fadd_rn(wx, __fsub_rn(1.f, x.lerp));
        if (x.hi == j) wx = __fadd_rn(wx, x.lerp);
        if (wx == 0.f) continue;
        acc = __fadd_rn(acc, __fmul_rn(__fmul_rn(wy, wx), __ldg(g + ((size_t)oi * 2 * W + oj) * C)));
      }
    }
    g_src[idx] = __fmul_rn(scale, acc);
  }
}

static int grid_for_n(long long total) {
  const long long want = cdivll(total, 256);
  return (int)(want < 148LL * 32 ? want : 148LL * 32);
}

int launch_upsample2x_fwd(const float* src, float* dst, int B, int H, int W, int C, float scale, cudaStream_t stream) {
  const long long total = (long long)B * 4 * H * W * C;
  if (total == 0) return QPWC_OK;
  auto k = upsample2x_fwd_kernel;
  QPWC_LAUNCH(k, grid_for_n(total), 256, 0, stream, src, dst, H, W, C, scale, total);
  return check_launch("upsample2x_fwd");
}

int launch_upsample2x_bwd(const float* g_dst, float* g_src, int B, int H, int W, int C, float scale, cudaStream_t stream) {
  const long long total = (long long)B * H * W * C;
  if (total == 0) return QPWC_OK;
  auto k = upsample2x_bwd_kernel;
  QPWC_LAUNCH(k, grid_for_n(total), 256, 0, stream, g_dst, g_src, H, W, C, scale, total);
  return check_launch("upsample2x_bwd");
}

}  // namespace qpwc
