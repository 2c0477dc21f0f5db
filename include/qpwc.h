/* qpwc.h -- C ABI of libqpwc.so: B200 (sm_100a) kernels for the QPWCNet per-pyramid-level hot path.
 *
 * The reference (yycho0108/qpwcnet) has no FFI boundary of its own: the hot path is the Python call
 * surface of four Keras layers / functors and one function.  Each entry point below is what a
 * binding for that surface would call; the interface it replaces is cited as file:line relative to
 * the reference checkout.  INTEGRATION.md shows the TF2-side stub (tf.custom_gradient + DLPack)
 * and the in-repo PyTorch/ctypes host layer (qpwcnet_b200/) that already binds them.
 *
 * Conventions
 *  - All tensors: dense fp32, NHWC ("channels_last"): index ((b*H + i)*W + j)*C + c.  Flow is
 *    (B,H,W,2) with channel 0 = x (width) and channel 1 = y (height) displacement, in pixels
 *    (qpwcnet/core/warp.py:102,109-111).
 *  - Device entry points take DEVICE pointers of the current CUDA device and enqueue on `stream`
 *    (a cudaStream_t passed as void*; NULL = legacy default stream).  No implicit synchronisation.
 *    The `_host` entry points take HOST pointers (pageable or pinned), stage through an internal,
 *    lazily grown per-device workspace and return after the results are in the host buffers.
 *  - Ownership: the caller allocates every buffer, outputs and workspaces included; the library
 *    keeps no pointer past the call.  Scatter-add targets (g_img, g_nxt) are zero-filled by the
 *    library with a stream-ordered memset.
 *  - Errors: no exception crosses the ABI.  Every function returns QPWC_OK (0) or a non-zero code
 *    and records a message retrievable (per calling thread) with qpwc_last_error().
 *  - Threading: device-pointer entry points keep no state between calls except caches that are
 *    per device and published with atomics (SM count, kernel attributes) or per calling thread (TMA
 *    descriptors, the last error); concurrent calls from several host threads, streams and devices
 *    are legal (tests/test_gpu_parity.py::test_two_threads_two_streams).  The _host entry points
 *    serialise per device on an internal mutex (they share that device's staging slots); their
 *    deferred-completion switch is per calling thread.  ctypes releases the GIL during a call.
 *  - `search_range` d >= 1: D = (2d+1)^2 output channels, channel = (di+d)*(2d+1) + (dj+d), row
 *    displacement outer (qpwcnet/core/layers.py:80-81).  `out_pixel_stride` (in floats, >= D) lets
 *    the cost volume be written straight into its slice of a wider concat buffer
 *    (qpwcnet/core/non_layers.py:381-382); pass D for the dense reference layout.
 *  - warp `mode`: QPWC_WARP_TF  = Warp / tf_warp (truncate + clip + weights from clipped corners),
 *                 QPWC_WARP_TFA = WarpV2 / tfa.image.dense_image_warp (floor, edge clamp).
 */
#ifndef QPWC_H_
#define QPWC_H_

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QPWC_OK 0
#define QPWC_ERR_INVALID 1     /* null / misaligned pointer, bad shape, bad mode            */
#define QPWC_ERR_UNSUPPORTED 2 /* valid request this build cannot serve                      */
#define QPWC_ERR_CUDA 3        /* CUDA runtime / launch error (message holds the CUDA string) */

#define QPWC_WARP_TF 0
#define QPWC_WARP_TFA 1

/* Library version (major*10000 + minor*100 + patch). */
int qpwc_version(void);
/* Message of the last failing call on this thread ("" if none). */
const char* qpwc_last_error(void);

/* Process-wide options.  QPWC_OPT_CORR_ENGINE selects the arithmetic of the cost-volume forward
 * kernels: QPWC_ENGINE_AUTO (default) = tensor cores where they win (search_range 4 with
 * C % 8 == 0; search_range 8 with C % 8 == 0 and C >= 32): fp32 operands split into tf32 hi + lo, three tcgen05.mma passes, fp32 accumulation,
 * measured error <= 8e-7 * mean|prv*nxt| (the reference contract is 1e-5); QPWC_ENGINE_FFMA = plain
 * fp32 FFMA kernels only (the UpFlow pair is then the in-kernel fusion; under AUTO it is the warp
 * kernel followed by the cost-volume kernel through a scratch owned by the call); QPWC_ENGINE_TC = tensor cores (search_range 4, or 8 as
 * four 9x9 windows; C % 8 == 0) or QPWC_ERR_UNSUPPORTED. */
#define QPWC_OPT_CORR_ENGINE 0
#define QPWC_ENGINE_AUTO 0
#define QPWC_ENGINE_FFMA 1
#define QPWC_ENGINE_TC 2
/* QPWC_OPT_WARP_BWD: warp backward kernel -- 0 auto / 1: vector global atomics with coincident taps folded in
 * registers (the faster one on B200); 2: per-tile pre-aggregation in shared memory before the global
 * atomics (C % 32 == 0; measured 2.5x slower: shared-memory float atomics retire ~1 lane/clk/SM). */
#define QPWC_OPT_WARP_BWD 1
/* QPWC_OPT_CORR_BWD: cost-volume gradient kernels -- 0 auto (register-tiled kernel where its domain allows:
   d == 4, C % 4 == 0, 16-byte aligned tensors) / 1: the shape-generic untiled kernels everywhere (a second
   summation order of the same sums; used by the tests). */
#define QPWC_OPT_CORR_BWD 2
int qpwc_set_option(int key, int value);
int qpwc_get_option(int key); /* -1 for an unknown key */

/* CostVolume.call / CostVolumeV2.call -- qpwcnet/core/layers.py:72-100, 117-132
 * (functors: qpwcnet/core/non_layers.py:72-104, 112-123), leaky_relu(slope) included. */
int qpwc_corr_fwd(const float* prv, const float* nxt, float* out, int B, int H, int W, int C,
                  int search_range, float leaky_slope, long long out_pixel_stride, void* stream);

/* The same layer with data_format='channels_first' (layers.py:83-85; the reference's training
 * layout, pre_train.py:34), natively: prv, nxt (B,C,H,W) -> out (B,(2d+1)^2,H,W), all dense.
 * search_range == 4, W % 4 == 0 and 16-byte aligned tensors take the tiled TMA kernel; every other
 * shape (any search range, any W, any alignment) a shape-generic NCHW kernel -- no caller ever
 * needs to transpose. */
int qpwc_corr_fwd_nchw(const float* prv, const float* nxt, float* out, int B, int C, int H, int W,
                       int search_range, float leaky_slope, void* stream);

/* Gradient of the above (TF autodiff of layers.py:77-99 / tfa CorrelationCostGrad + LeakyReluGrad).
 * `out` is the forward result (sign gives the leaky mask); g_out shares its pixel stride. */
int qpwc_corr_bwd(const float* prv, const float* nxt, const float* out, const float* g_out,
                  float* g_prv, float* g_nxt, int B, int H, int W, int C, int search_range,
                  float leaky_slope, long long out_pixel_stride, void* stream);

/* Warp.call -> tf_warp (qpwcnet/core/warp.py:63-153; layers.py:166-168) and
 * WarpV2.call -> tfa.image.dense_image_warp(img, -flo[..., ::-1]) (layers.py:177-186). */
int qpwc_warp_fwd(const float* img, const float* flow, float* out, int B, int H, int W, int C,
                  int mode, void* stream);

/* The same layers under data_format='channels_first' (warp.py:36-40 transposes around gather_nd;
 * layers.py:179-183), natively: img (B,C,H,W), flow (B,2,H,W) with plane 0 = x, out (B,C,H,W);
 * the flow is multiplied by flow_scale first (1.0 for Warp/WarpV2). */
int qpwc_warp_fwd_nchw(const float* img, const float* flow, float* out, int B, int C, int H, int W,
                       int mode, float flow_scale, void* stream);

/* Gradient of the above: g_img (zero-filled here, then scatter-added) and g_flow. */
int qpwc_warp_bwd(const float* img, const float* flow, const float* g_out, float* g_img,
                  float* g_flow, int B, int H, int W, int C, int mode, void* stream);

/* FrameInterpolate's two half-flow warps -- qpwcnet/core/non_layers.py:303-311 (layers.py:384-385):
 *   nxt_w = warp((nxt, 0.5 * flo_01));  prv_w = warp((prv, 0.5 * flo_10));  concat([prv_w, nxt_w, ...])
 * `_ex`: the flow is multiplied by flow_scale inside the kernel (one rounded multiply, the
 * reference's own op) and the output may be a channel slice of a wider buffer (pixel stride >= C).
 * `_pair_fwd`: both warps in ONE launch; out[..., 0:C] = warp(img_a, s*flow_a),
 * out[..., C:2C] = warp(img_b, s*flow_b), pixel stride >= 2C.  `_bwd_ex`: g_out may be such a
 * slice; g_flow is the gradient with respect to the UNSCALED flow. */
int qpwc_warp_fwd_ex(const float* img, const float* flow, float* out, int B, int H, int W, int C,
                     int mode, float flow_scale, long long out_pixel_stride, void* stream);
/* Row-sharded frames (BASELINE config 5, qpwcnet_b200/sharded.py): `img`, `flow`, `out` are rows
 * [row_offset, row_offset + H) of tensors `full_height` rows tall.  Sampling coordinates, truncation and
 * border clamping use the absolute row index -- the reference adds the flow to the absolute pixel index
 * in fp32 (warp.py:100-111), so the result is bit-identical to the same rows of the unsharded warp for
 * every output row whose four taps lie inside the view (rows further than max|flow_y| + 1 from a cut). */
int qpwc_warp_fwd_rows(const float* img, const float* flow, float* out, int B, int H, int W, int C,
                       int mode, int row_offset, int full_height, void* stream);

int qpwc_warp_pair_fwd(const float* img_a, const float* flow_a, const float* img_b,
                       const float* flow_b, float* out, int B, int H, int W, int C, int mode,
                       float flow_scale, long long out_pixel_stride, void* stream);
int qpwc_warp_bwd_ex(const float* img, const float* flow, const float* g_out, float* g_img,
                     float* g_flow, int B, int H, int W, int C, int mode, float flow_scale,
                     long long g_out_pixel_stride, void* stream);

/* Upsample(scale) -- qpwcnet/core/non_layers.py:183-193: scale * UpSampling2D(interpolation='bilinear')
 * (tf.image.resize bilinear, half-pixel centres), x2.  src (B,H,W,C) -> dst (B,2H,2W,C); `_bwd` is
 * its adjoint (g_dst (B,2H,2W,C) -> g_src (B,H,W,C)). */
int qpwc_upsample2x_fwd(const float* src, float* dst, int B, int H, int W, int C, float scale, void* stream);
int qpwc_upsample2x_bwd(const float* g_dst, float* g_src, int B, int H, int W, int C, float scale, void* stream);

/* Gradients of qpwc_corr_fwd_nchw (autodiff of the channels_first layer): out / g_out (B,(2d+1)^2,H,W),
 * g_prv / g_nxt (B,C,H,W), every element written exactly once (no atomics, no zero-init).  Tiled
 * kernels inside the forward kernel's tiled domain, shape-generic NCHW kernel elsewhere. */
int qpwc_corr_bwd_nchw(const float* prv, const float* nxt, const float* out, const float* g_out,
                       float* g_prv, float* g_nxt, int B, int C, int H, int W, int search_range,
                       float leaky_slope, void* stream);

/* Gradient of qpwc_warp_fwd_nchw: g_out (B,C,H,W) -> g_img (B,C,H,W) (zeroed here, accumulated with
 * atomics like the NHWC kernel) and g_flow (B,2,H,W) (per-pixel sum over the channels in order). */
int qpwc_warp_bwd_nchw(const float* img, const float* flow, const float* g_out, float* g_img,
                       float* g_flow, int B, int C, int H, int W, int mode, float flow_scale, void* stream);

/* estimate_occlusion_map(flow, data_format) -- qpwcnet/core/occlusion.py:27-118.  flow is
 * (B,H,W,2) (channels_first = 0) or (B,2,H,W) (channels_first = 1), channel 0 = dx, 1 = dy;
 * out (B,H,W) = max(oob, map3): oob = the flow target leaves the image (occlusion.py:74), map3 = 0
 * where some pixel's naive inverse flow -tf_warp(flow, flow) lands (int cast, clipped), else 1
 * (occlusion.py:83-96).  The self-warp is evaluated in registers; two launches on `stream`. */
int qpwc_occlusion_map(const float* flow, float* out, int B, int H, int W, int channels_first, void* stream);

/* The flow upsampling fused into its consumers (pwcnet.py:49-56: flo = Upsample(2.0)(flo) feeds
 * UpFlow): flow_coarse is (B,H/2,W/2,2); the kernels sample with up_scale * bilinear_x2(flow_coarse)
 * interpolated in registers, so the upsampled flow is not read back from HBM.  H, W (output size) even. */
int qpwc_warp_fwd_up(const float* img, const float* flow_coarse, float* out, int B, int H, int W, int C,
                     int mode, float up_scale, void* stream);
int qpwc_warp_corr_fwd_up(const float* prv, const float* nxt, const float* flow_coarse, float* out, int B,
                          int H, int W, int C, int search_range, float leaky_slope, int mode,
                          long long out_pixel_stride, float up_scale, void* stream);

/* UpFlow's  CostVolumeV2((prv, WarpV2((nxt, flo))))  -- qpwcnet/core/non_layers.py:377-380
 * (layers.py:478-481) as ONE call without caller workspace.  AUTO / tensor-core engine: the warp kernel and
 * the cost-volume kernel run back to back on chunks of frame pairs through a stream-ordered scratch buffer
 * owned by the call and reused chunk after chunk (what stays in L2 between the two kernels depends on the
 * chunk size against the 126 MB L2: the coarse levels fit, the finest config-2 level largely does not --
 * 1.27x the algorithmic DRAM bytes there).  FFMA engine: one kernel, the warped tile lives in shared
 * memory only (fewer DRAM bytes, slower: DESIGN.md 4.3). */
int qpwc_warp_corr_fwd(const float* prv, const float* nxt, const float* flow, float* out, int B,
                       int H, int W, int C, int search_range, float leaky_slope, int mode,
                       long long out_pixel_stride, void* stream);

/* Gradient of the fused op: g_prv, g_nxt (zero-filled here), g_flow.  No caller workspace (since
 * 0.2): the warped frame and its gradient are rebuilt chunk by chunk in an L2-resident, stream-ordered
 * scratch buffer owned by the call; `workspace` / `workspace_bytes` are accepted and ignored, and
 * qpwc_warp_corr_bwd_workspace() returns 0. */
size_t qpwc_warp_corr_bwd_workspace(int B, int H, int W, int C);
int qpwc_warp_corr_bwd(const float* prv, const float* nxt, const float* flow, const float* out,
                       const float* g_out, float* g_prv, float* g_nxt, float* g_flow,
                       void* workspace, size_t workspace_bytes, int B, int H, int W, int C,
                       int search_range, float leaky_slope, int mode, long long out_pixel_stride,
                       void* stream);

/* Host-buffer variants of the forward ops (what a TF-CPU caller of the layers would use): copy in,
 * run the device kernels, copy out, batch-sliced so that H2D, kernels and D2H overlap on three
 * streams.  `device` = CUDA device ordinal. */
int qpwc_corr_fwd_host(const float* prv, const float* nxt, float* out, int B, int H, int W, int C,
                       int search_range, float leaky_slope, int device);
int qpwc_warp_fwd_host(const float* img, const float* flow, float* out, int B, int H, int W, int C,
                       int mode, int device);
int qpwc_warp_corr_fwd_host(const float* prv, const float* nxt, const float* flow, float* out,
                            int B, int H, int W, int C, int search_range, float leaky_slope,
                            int mode, int device);

/* Deferred completion for the _host entry points: after qpwc_host_set_deferred(1) a _host call
 * returns as soon as its copies and kernels are enqueued (on the library's internal streams), so
 * that consecutive calls -- e.g. the five levels of one pyramid pass -- overlap their H2D, compute
 * and D2H phases; qpwc_host_sync(device) then waits for everything enqueued so far on that device
 * (device = -1: on every device).  The caller must keep input and output host buffers alive and
 * unmodified until the sync.  The switch is per calling thread: other threads' _host calls keep
 * returning completed outputs. */
int qpwc_host_set_deferred(int on);
int qpwc_host_sync(int device);

#ifdef __cplusplus
}
#endif
#endif /* QPWC_H_ */
