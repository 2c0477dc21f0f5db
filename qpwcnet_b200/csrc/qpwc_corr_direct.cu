// qpwc_corr_direct.cu -- direct (untiled) local-correlation kernels: any C, any search range d.
//
// Replaces CostVolume.call / CostVolumeV2.call (qpwcnet/core/layers.py:72-100, 117-132):
//   out[b,i,j,(di+d)*(2d+1)+(dj+d)] = leaky_relu_slope( (1/C) sum_c prv[b,i,j,c] * nxt[b,i+di,j+dj,c] )
// with nxt == 0 outside the image (ZeroPadding2D), optionally on a second frame that is first
// warped by a flow field (UpFlow, non_layers.py:377-380) without materialising the warped tensor.
//
// These are the shape-generic kernels (C not a multiple of 4, d other than 4, tiny maps).  The
// register-tiled kernels in qpwc_corr_tiled.cu take over for the pyramid shapes.  Forward: one
// thread per output element, sequential channel sum (same order as the fp32 oracle).  Backward:
// one thread per (pixel, channel) gathering over the (2d+1)^2 displacements for both gradients.
#include "qpwc_common.cuh"

namespace qpwc {

// WARP: 0 = plain second frame, 1 = sample the second frame through flow (mode MODE)
template <int WARP, int MODE>
__global__ void __launch_bounds__(256) corr_fwd_direct_kernel(
    const float* __restrict__ prv, const float* __restrict__ nxt, const float* __restrict__ flow,
    float* __restrict__ out, int H, int W, int C, int d, float slope, long long ops,
    long long total /* B*H*W*D */) {
  const int q = 2 * d + 1;
  const int D = q * q;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int k = (int)(idx % D);
    const long long pix = idx / D;
    const int j = (int)(pix % W);
    const long long bi = pix / W;
    const int i = (int)(bi % H);
    const long long b = bi / H;
    const int r = i + k / q - d, s = j + k % q - d;
    float acc = 0.f;
    if (r >= 0 && r < H && s >= 0 && s < W) {
      const float* p = prv + (size_t)pix * C;
      const float* nb = nxt + (size_t)b * H * W * C;
      if (WARP) {
        const float2 f = __ldg(reinterpret_cast<const float2*>(flow) + ((size_t)b * H * W + (size_t)r * W + s));
        const Taps t = make_taps<MODE>(r, s, f.x, f.y, H, W);
        const float* n00 = nb + (size_t)t.o00 * C; const float* n01 = nb + (size_t)t.o01 * C;
        const float* n10 = nb + (size_t)t.o10 * C; const float* n11 = nb + (size_t)t.o11 * C;
        for (int c = 0; c < C; ++c)
          acc = fmaf(__ldg(p + c), blend<MODE>(t, __ldg(n00 + c), __ldg(n01 + c), __ldg(n10 + c), __ldg(n11 + c)), acc);
      } else {
        const float* n = nb + ((size_t)r * W + s) * C;
        for (int c = 0; c < C; ++c) acc = fmaf(__ldg(p + c), __ldg(n + c), acc);
      }
    }
    out[(size_t)pix * ops + k] = lrelu(acc / (float)C, slope);
  }
}

// g_pre[p,k] = g_out[p,k] * (out[p,k] > 0 ? 1 : slope) / C
//   g_prv[p,c] = sum_k g_pre[p,k] * nxt[p+delta_k, c]
//   g_nxt[p,c] = sum_k g_pre[p-delta_k, k] * prv[p-delta_k, c]
__global__ void __launch_bounds__(256) corr_bwd_direct_kernel(
    const float* __restrict__ prv, const float* __restrict__ nxt, const float* __restrict__ out,
    const float* __restrict__ g_out, float* __restrict__ g_prv, float* __restrict__ g_nxt, int H,
    int W, int C, int d, float slope, long long ops, long long total /* B*H*W*C */) {
  const int q = 2 * d + 1;
  const float inv_c = 1.f / (float)C;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += stride) {
    const int c = (int)(idx % C);
    const long long pix = idx / C;
    const int j = (int)(pix % W);
    const long long bi = pix / W;
    const int i = (int)(bi % H);
    const long long b = bi / H;
    const size_t bpix = (size_t)b * H * W;
    float ap = 0.f, an = 0.f;
    for (int i0 = 0; i0 < q; ++i0)
      for (int j0 = 0; j0 < q; ++j0) {
        const int k = i0 * q + j0;
        const int di = i0 - d, dj = j0 - d;
        {  // g_prv: this pixel's own gradient row against the displaced second frame
          const int r = i + di, s = j + dj;
          if (r >= 0 && r < H && s >= 0 && s < W) {
            const size_t o = (size_t)pix * ops + k;
            float g = __ldg(g_out + o);
            g = __ldg(out + o) > 0.f ? g : slope * g;
            ap = fmaf(g * inv_c, __ldg(nxt + (bpix + (size_t)r * W + s) * C + c), ap);
          }
        }
        {  // g_nxt: first-frame pixels that looked at this second-frame pixel
          const int r = i - di, s = j - dj;
          if (r >= 0 && r < H && s >= 0 && s < W) {
            const size_t pp = bpix + (size_t)r * W + s;
            const size_t o = pp * ops + k;
            float g = __ldg(g_out + o);
            g = __ldg(out + o) > 0.f ? g : slope * g;
            an = fmaf(g * inv_c, __ldg(prv + pp * C + c), an);
          }
        }
      }
    g_prv[idx] = ap;
    g_nxt[idx] = an;
  }
}

static int grid_for(long long total, int block, int waves) {
  const long long want = cdivll(total, block);
  const long long cap = 148LL * waves;
  return (int)(want < cap ? want : cap);
}

int launch_corr_fwd_direct(const float* prv, const float* nxt, const float* flow, int mode,
                           float* out, int B, int H, int W, int C, int d, float slope,
                           long long ops, cudaStream_t stream) {
  const int D = (2 * d + 1) * (2 * d + 1);
  const long long total = (long long)B * H * W * D;
  if (total == 0) return QPWC_OK;
  const int block = 256, grid = grid_for(total, block, 64);
  if (!flow) {
    auto k = corr_fwd_direct_kernel<0, QPWC_MODE_TF>;
    QPWC_LAUNCH(k, grid, block, 0, stream, prv, nxt, flow, out, H, W, C, d, slope, ops, total);
  } else if (mode == QPWC_MODE_TF) {
    auto k = corr_fwd_direct_kernel<1, QPWC_MODE_TF>;
    QPWC_LAUNCH(k, grid, block, 0, stream, prv, nxt, flow, out, H, W, C, d, slope, ops, total);
  } else {
    auto k = corr_fwd_direct_kernel<1, QPWC_MODE_TFA>;
    QPWC_LAUNCH(k, grid, block, 0, stream, prv, nxt, flow, out, H, W, C, d, slope, ops, total);
  }
  return check_launch("corr_fwd_direct");
}

int launch_corr_bwd_direct(const float* prv, const float* nxt, const float* out, const float* g_out,
                           float* g_prv, float* g_nxt, int B, int H, int W, int C, int d,
                           float slope, long long ops, cudaStream_t stream) {
  const long long total = (long long)B * H * W * C;
  if (total == 0) return QPWC_OK;
  const int block = 256, grid = grid_for(total, block, 64);
  auto k = corr_bwd_direct_kernel;
  QPWC_LAUNCH(k, grid, block, 0, stream, prv, nxt, out, g_out, g_prv, g_nxt, H, W, C, d, slope, ops, total);
  return check_launch("corr_bwd_direct");
}

}  // namespace qpwc
