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
#include "qpwc_upsample.cuh"

namespace qpwc {

// WARP: 0 = plain second frame, 1 = sample the second frame through flow (mode MODE)
template <int WARP, int MODE>
__global__ void __launch_bounds__(256) corr_fwd_direct_kernel(
    const float* __restrict__ prv, const float* __restrict__ nxt, const float* __restrict__ flow,
    float* __restrict__ out, int H, int W, int C, int d, float slope, long long ops,
    long long total /* B*H*W*D */, float up_scale) {
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
        const float2 f = up_scale != 0.f
            ? up2_flow(flow + (size_t)b * (H / 2) * (W / 2) * 2, r, s, H / 2, W / 2, up_scale)
            : __ldg(reinterpret_cast<const float2*>(flow) + ((size_t)b * H * W + (size_t)r * W + s));
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

// Vectorised variant for C % 4 == 0 and 16-byte aligned tensors: one thread per (pixel, channel
// quad); grid = (ceil(W*C4/256), H, B) so all index arithmetic is 32-bit.  The 8..64 quad-lanes of
// a pixel share the g_out/out loads (warp broadcast), the neighbour vectors are 16-byte loads.
__global__ void __launch_bounds__(256) corr_bwd_quad_kernel(
    const float* __restrict__ prv, const float* __restrict__ nxt, const float* __restrict__ out,
    const float* __restrict__ g_out, float* __restrict__ g_prv, float* __restrict__ g_nxt, int H,
    int W, int C, int d, float slope, long long ops) {
  const int C4 = C >> 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;  // j * C4 + c4
  if (idx >= W * C4) return;
  const int j = idx / C4, c4 = idx - j * C4;
  const int i = blockIdx.y;
  const size_t bpix = (size_t)blockIdx.z * H * W;
  const size_t pix = bpix + (size_t)i * W + j;
  const int q = 2 * d + 1;
  const float inv_c = 1.f / (float)C;
  const float4* prv4 = reinterpret_cast<const float4*>(prv);
  const float4* nxt4 = reinterpret_cast<const float4*>(nxt);
  const float* go_own = g_out + pix * (size_t)ops;
  const float* o_own = out + pix * (size_t)ops;
  float4 ap = make_float4(0.f, 0.f, 0.f, 0.f), an = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int i0 = 0; i0 < q; ++i0) {
    const int r = i + i0 - d, rn = i - (i0 - d);
    const bool r_ok = r >= 0 && r < H, rn_ok = rn >= 0 && rn < H;
    for (int j0 = 0; j0 < q; ++j0) {
      const int k = i0 * q + j0;
      const int s = j + j0 - d, sn = j - (j0 - d);
      if (r_ok && s >= 0 && s < W) {  // g_prv: own gradient row against the displaced second frame
        float g = __ldg(go_own + k);
        g = (__ldg(o_own + k) > 0.f ? g : slope * g) * inv_c;
        const float4 x = __ldg(nxt4 + (bpix + (size_t)r * W + s) * C4 + c4);
        ap.x = fmaf(g, x.x, ap.x); ap.y = fmaf(g, x.y, ap.y); ap.z = fmaf(g, x.z, ap.z); ap.w = fmaf(g, x.w, ap.w);
      }
      if (rn_ok && sn >= 0 && sn < W) {  // g_nxt: first-frame pixels that looked at this pixel
        const size_t pp = bpix + (size_t)rn * W + sn;
        float g = __ldg(g_out + pp * (size_t)ops + k);
        g = (__ldg(out + pp * (size_t)ops + k) > 0.f ? g : slope * g) * inv_c;
        const float4 x = __ldg(prv4 + pp * C4 + c4);
        an.x = fmaf(g, x.x, an.x); an.y = fmaf(g, x.y, an.y); an.z = fmaf(g, x.z, an.z); an.w = fmaf(g, x.w, an.w);
      }
    }
  }
  reinterpret_cast<float4*>(g_prv)[pix * C4 + c4] = ap;
  reinterpret_cast<float4*>(g_nxt)[pix * C4 + c4] = an;
}

static int grid_for(long long total, int block, int waves) {
  const long long want = cdivll(total, block);
  const long long cap = 148LL * waves;
  return (int)(want < cap ? want : cap);
}

int launch_corr_fwd_direct(const float* prv, const float* nxt, const float* flow, int mode,
                           float* out, int B, int H, int W, int C, int d, float slope,
                           long long ops, cudaStream_t stream, float up_scale) {
  const int D = (2 * d + 1) * (2 * d + 1);
  const long long total = (long long)B * H * W * D;
  if (total == 0) return QPWC_OK;
  const int block = 256, grid = grid_for(total, block, 64);
  if (!flow) {
    auto k = corr_fwd_direct_kernel<0, QPWC_MODE_TF>;
    QPWC_LAUNCH(k, grid, block, 0, stream, prv, nxt, flow, out, H, W, C, d, slope, ops, total, up_scale);
  } else if (mode == QPWC_MODE_TF) {
    auto k = corr_fwd_direct_kernel<1, QPWC_MODE_TF>;
    QPWC_LAUNCH(k, grid, block, 0, stream, prv, nxt, flow, out, H, W, C, d, slope, ops, total, up_scale);
  } else {
    auto k = corr_fwd_direct_kernel<1, QPWC_MODE_TFA>;
    QPWC_LAUNCH(k, grid, block, 0, stream, prv, nxt, flow, out, H, W, C, d, slope, ops, total, up_scale);
  }
  return check_launch("corr_fwd_direct");
}

int launch_corr_bwd_direct(const float* prv, const float* nxt, const float* out, const float* g_out,
                           float* g_prv, float* g_nxt, int B, int H, int W, int C, int d,
                           float slope, long long ops, cudaStream_t stream) {
  const long long total = (long long)B * H * W * C;
  if (total == 0) return QPWC_OK;
  const bool al16 = !((reinterpret_cast<uintptr_t>(prv) | reinterpret_cast<uintptr_t>(nxt) |
                       reinterpret_cast<uintptr_t>(g_prv) | reinterpret_cast<uintptr_t>(g_nxt)) & 15);
  if ((C & 3) == 0 && al16 && H <= 65535 && (long long)W * (C >> 2) < (1LL << 31)) {
    auto k = corr_bwd_quad_kernel;
    for (int b0 = 0; b0 < B; b0 += 65535) {
      const int nb = (B - b0 < 65535) ? (B - b0) : 65535;
      const dim3 grid((unsigned)cdiv(W * (C >> 2), 256), (unsigned)H, (unsigned)nb);
      const size_t off = (size_t)b0 * H * W;
      QPWC_LAUNCH(k, grid, 256, 0, stream, prv + off * C, nxt + off * C, out + off * ops, g_out + off * ops,
                  g_prv + off * C, g_nxt + off * C, H, W, C, d, slope, ops);
    }
    return check_launch("corr_bwd_quad");
  }
  const int block = 256, grid = grid_for(total, block, 64);
  auto k = corr_bwd_direct_kernel;
  QPWC_LAUNCH(k, grid, block, 0, stream, prv, nxt, out, g_out, g_prv, g_nxt, H, W, C, d, slope, ops, total);
  return check_launch("corr_bwd_direct");
}

}  // namespace qpwc
