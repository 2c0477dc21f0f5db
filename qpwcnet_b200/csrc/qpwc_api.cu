// qpwc_api.cu -- extern "C" entry points of libqpwc.so (declared in include/qpwc.h):
// argument validation, kernel selection, error plumbing, and the host-buffer (staged) variants.
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>

#include "qpwc_common.cuh"
#include "../../include/qpwc.h"

namespace qpwc {

// kernels (qpwc_warp.cu, qpwc_corr_direct.cu, qpwc_corr_tiled.cu)
int launch_warp_fwd(const float*, const float*, float*, int, int, int, int, int, cudaStream_t);
int launch_warp_bwd(const float*, const float*, const float*, float*, float*, int, int, int, int, int, cudaStream_t);
int launch_warp_fwd_ex(const float*, const float*, const float*, const float*, float*, int, int, int, int, int, float, long long, cudaStream_t, float up_scale = 0.f, int row_off = 0, int Hfull = 0, int keep_l2 = 0);
int launch_upsample2x_fwd(const float*, float*, int, int, int, int, float, cudaStream_t);
int launch_warp_fwd_nchw(const float*, const float*, float*, int, int, int, int, int, float, cudaStream_t);
int launch_corr_fwd_nchw(const float*, const float*, float*, int, int, int, int, int, float, cudaStream_t);
int launch_upsample2x_bwd(const float*, float*, int, int, int, int, float, cudaStream_t);
int launch_warp_bwd_nchw(const float*, const float*, const float*, float*, float*, int, int, int, int, int, float, cudaStream_t);
int launch_corr_bwd_nchw(const float*, const float*, const float*, const float*, float*, float*, int, int, int, int, int, float, cudaStream_t);
int launch_occlusion_map(const float*, float*, int, int, int, int, cudaStream_t);
int launch_warp_bwd_ex(const float*, const float*, const float*, float*, float*, int, int, int, int, int, float, long long, cudaStream_t);
int launch_corr_fwd_direct(const float*, const float*, const float*, int, float*, int, int, int, int, int, float, long long, cudaStream_t, float up_scale = 0.f);
int launch_corr_bwd_direct(const float*, const float*, const float*, const float*, float*, float*, int, int, int, int, int, float, long long, cudaStream_t);
// returns QPWC_ERR_UNSUPPORTED (without setting an error) when the shape is outside its domain
int launch_corr_fwd_tiled(const float*, const float*, const float*, int, float*, int, int, int, int, int, float, long long, cudaStream_t, float up_scale = 0.f);
int launch_corr_bwd_tiled(const float*, const float*, const float*, const float*, float*, float*, int, int, int, int, int, float, long long, cudaStream_t);
// tensor-core (3xTF32) cost volume, qpwc_corr_tc.cu; QPWC_ERR_UNSUPPORTED outside its domain
int launch_corr_fwd_tc(const float*, const float*, float*, int, int, int, int, int, float, long long, cudaStream_t);

static thread_local char g_err[512] = "";

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_launch(const char* what) {
  const cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e));
  return QPWC_OK;
}

static int check_shape(const char* fn, int B, int H, int W, int C) {
  if (B < 0 || H < 0 || W < 0 || C < 0) return set_error(QPWC_ERR_INVALID, "%s: negative dimension (B=%d H=%d W=%d C=%d)", fn, B, H, W, C);
  if ((long long)H * W >= (1LL << 31) || (long long)H * W * (long long)(C > 0 ? C : 1) >= (1LL << 40))
    return set_error(QPWC_ERR_INVALID, "%s: image too large (H=%d W=%d C=%d)", fn, H, W, C);
  return QPWC_OK;
}
static bool empty(int B, int H, int W, int C) { return B == 0 || H == 0 || W == 0 || C == 0; }

static int check_ptr(const char* fn, const char* name, const void* p) {
  if (!p) return set_error(QPWC_ERR_INVALID, "%s: %s is NULL", fn, name);
  if (reinterpret_cast<uintptr_t>(p) % sizeof(float)) return set_error(QPWC_ERR_INVALID, "%s: %s is not 4-byte aligned", fn, name);
  return QPWC_OK;
}
#define QPWC_TRY(expr) do { const int rc_ = (expr); if (rc_ != QPWC_OK) return rc_; } while (0)

static int check_corr_args(const char* fn, int d, long long ops) {
  if (d < 1 || d > 32) return set_error(QPWC_ERR_INVALID, "%s: search_range %d outside [1,32]", fn, d);
  const long long D = (long long)(2 * d + 1) * (2 * d + 1);
  if (ops < D) return set_error(QPWC_ERR_INVALID, "%s: out_pixel_stride %lld < (2d+1)^2 = %lld", fn, ops, D);
  return QPWC_OK;
}
static int check_mode(const char* fn, int mode, int H, int W) {
  if (mode != QPWC_WARP_TF && mode != QPWC_WARP_TFA) return set_error(QPWC_ERR_INVALID, "%s: unknown warp mode %d", fn, mode);
  if (mode == QPWC_WARP_TFA && (H < 2 || W < 2))
    return set_error(QPWC_ERR_INVALID, "%s: Grid must be at least 2x2 for the tfa bilinear mode (H=%d W=%d)", fn, H, W);
  return QPWC_OK;
}

// qpwc_set_option(QPWC_OPT_CORR_ENGINE, ...): 0 = auto (tensor cores where the shape allows), 1 = FFMA
// kernels only (plain fp32 arithmetic), 2 = tensor cores (3xTF32 split) or fail
static std::atomic<int> g_corr_engine{0};
void set_warp_bwd_variant(int);   // qpwc_warp.cu
static std::atomic<int> g_corr_bwd_variant{0};   // QPWC_OPT_CORR_BWD
void set_corr_bwd_variant(int v) { g_corr_bwd_variant.store(v); }
int get_corr_bwd_variant() { return g_corr_bwd_variant.load(std::memory_order_relaxed); }
int get_warp_bwd_variant();

// Engine policy.  AUTO takes the tensor-core kernels for search range 4, and for search range 8 from 32
// channels up.  Range 8 runs on either engine as four 9x9 windows whose results leave as 36-byte runs at a
// 1156-byte pixel pitch (no bulk / tensor-map store: the 68-byte displacement-row pitch is not a multiple of
// 16); with the window index running fastest in the tile order the runs of a pixel meet in L2
// (profiles/r02c_tc_d8.txt, B = 8: 218x512x16 ffma 730 / tc 834 us, 109x256x32 248 / 240, 224x512x32 952 / 856,
// 56x128x64 110 / 78).  Forcing an engine selects it at every shape of its domain.
#ifdef QPWC_EMU
static bool tc_wanted(int, int, int = 0) { return false; }   // the emulation harness has no tensor-core stand-in
#else
static bool tc_wanted(int engine, int d, int C = 1 << 30) { return engine == 2 || (engine == 0 && (d == 4 || (d == 8 && C >= 32))); }
#endif

static int corr_fwd_any(const float* prv, const float* nxt, const float* flow, int mode, float* out,
                        int B, int H, int W, int C, int d, float slope, long long ops, cudaStream_t st,
                        float up_scale = 0.f) {
  const int engine = g_corr_engine.load(std::memory_order_relaxed);
  if (!flow && tc_wanted(engine, d, C)) {
    const int rt = launch_corr_fwd_tc(prv, nxt, out, B, H, W, C, d, slope, ops, st);
    if (rt != QPWC_ERR_UNSUPPORTED) return rt;
    if (engine == 2) return set_error(QPWC_ERR_UNSUPPORTED, "corr_fwd: tensor-core engine needs search_range 4 or 8, C %% 8 == 0 and 16-byte aligned inputs");
  }
  const int rc = launch_corr_fwd_tiled(prv, nxt, flow, mode, out, B, H, W, C, d, slope, ops, st, up_scale);
  if (rc != QPWC_ERR_UNSUPPORTED) return rc;
  return launch_corr_fwd_direct(prv, nxt, flow, mode, out, B, H, W, C, d, slope, ops, st, up_scale);
}
static int corr_bwd_any(const float* prv, const float* nxt, const float* out, const float* g_out,
                        float* g_prv, float* g_nxt, int B, int H, int W, int C, int d, float slope,
                        long long ops, cudaStream_t st) {
  const int rc = launch_corr_bwd_tiled(prv, nxt, out, g_out, g_prv, g_nxt, B, H, W, C, d, slope, ops, st);
  if (rc != QPWC_ERR_UNSUPPORTED) return rc;
  return launch_corr_bwd_direct(prv, nxt, out, g_out, g_prv, g_nxt, B, H, W, C, d, slope, ops, st);
}


// ------------------------------------------------------------------ warp -> corr through L2
// The UpFlow pair with the tensor-core cost volume: the stand-alone warp kernel writes the warped second
// frame of a chunk of frame pairs into a stream-ordered scratch buffer owned by the call and the
// cost-volume kernel reads it back; the scratch is reused chunk after chunk.  Every warped pixel is
// computed once -- the in-kernel fusion (FFMA engine) recomputes the 8-pixel halo of every tile with
// latency-bound gathers, which is why it loses to this composition at every level although it moves
// fewer DRAM bytes (profiles/README.md has both byte counts).  What stays in L2 between the two
// kernels depends on the chunk size against the 126 MB L2: the coarse levels fit, the finest config-2
// level (117 MB) largely does not.
// Scratch per chunk.  Measured (profiles/README.md): splitting B = 8 at the finest level into two
// L2-sized chunks costs more in tile-count imbalance of the persistent kernel (24.2 tiles per SM run as 28)
// than the L2 hits save, so a chunk is as large as a whole config-2 level; bigger batches are cut here.
static const size_t kL2ChunkBytes = (size_t)160 << 20;

// Stream-ordered scratch: cudaMallocAsync from the device's default pool.  By default that pool hands
// its memory back to the OS at every synchronisation (release threshold 0), which turns each call into
// a driver allocation of ~100 MB; raise the threshold once per device so that the block is recycled.
#ifdef QPWC_EMU
static int scratch_alloc(float** p, size_t bytes, cudaStream_t, const char*) { *p = static_cast<float*>(malloc(bytes)); return QPWC_OK; }
static cudaError_t scratch_free(float* p, cudaStream_t) { free(p); return cudaSuccess; }
#else
static int scratch_alloc(float** p, size_t bytes, cudaStream_t st, const char* fn) {
  static std::atomic<unsigned> pool_ready{0};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!(pool_ready.load(std::memory_order_acquire) >> (dev & 31) & 1u)) {
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
      unsigned long long keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
    pool_ready.fetch_or(1u << (dev & 31), std::memory_order_release);
  }
  const cudaError_t e = cudaMallocAsync(p, bytes, st);
  if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "%s: scratch (%zu B): %s", fn, bytes, cudaGetErrorString(e));
  return QPWC_OK;
}
static cudaError_t scratch_free(float* p, cudaStream_t st) { return cudaFreeAsync(p, st); }
#endif

static int chunk_batch(int B, size_t per_item_bytes) {
  size_t n = per_item_bytes ? kL2ChunkBytes / per_item_bytes : (size_t)B;
  if (n < 1) n = 1;
  if (n > (size_t)B) n = (size_t)B;
  // even chunks: B = 8 with room for 5 items runs as 4 + 4, not 5 + 3
  const int nchunks = (B + (int)n - 1) / (int)n;
  return (B + nchunks - 1) / nchunks;
}

static bool tc_domain(const float* prv, const float* nxt, int C, int d) {
  return (d == 4 || d == 8) && C >= 8 && (C & 7) == 0 && !(reinterpret_cast<uintptr_t>(prv) & 15) && !(reinterpret_cast<uintptr_t>(nxt) & 15);
}

static int warp_corr_fwd_l2(const float* prv, const float* nxt, const float* flow, int mode, float* out, int B,
                            int H, int W, int C, int d, float slope, long long ops, cudaStream_t st, float up_scale) {
  const size_t item = (size_t)H * W * C, fitem = up_scale != 0.f ? (size_t)(H / 2) * (W / 2) * 2 : (size_t)H * W * 2;
  const int per = chunk_batch(B, item * sizeof(float));
  float* scratch = nullptr;
  int rc = scratch_alloc(&scratch, item * per * sizeof(float), st, "warp_corr_fwd");
  if (rc != QPWC_OK) return rc;
  cudaError_t e;
  for (int b0 = 0; b0 < B && rc == QPWC_OK; b0 += per) {
    const int nb = B - b0 < per ? B - b0 : per;
    static const int keep = [] { const char* ev = getenv("QPWC_NO_L2_KEEP"); return ev && ev[0] == '1' ? 0 : 1; }();   // dev switch
    rc = launch_warp_fwd_ex(nxt + item * b0, flow + fitem * b0, nullptr, nullptr, scratch, nb, H, W, C, mode, 1.f, C, st, up_scale, 0, 0, keep);
    // (cost volume of the scratch: tensor cores inside their domain, the FFMA kernels elsewhere)
    if (rc == QPWC_OK) rc = corr_fwd_any(prv + item * b0, scratch, nullptr, 0, out + (size_t)H * W * ops * b0, nb, H, W, C, d, slope, ops, st);
  }
  e = scratch_free(scratch, st);
  if (rc == QPWC_OK && e != cudaSuccess) rc = set_error(QPWC_ERR_CUDA, "warp_corr_fwd: scratch free: %s", cudaGetErrorString(e));
  return rc;
}

static int warp_corr_fwd_any(const float* prv, const float* nxt, const float* flow, int mode, float* out, int B,
                             int H, int W, int C, int d, float slope, long long ops, cudaStream_t st, float up_scale = 0.f) {
  const int engine = g_corr_engine.load(std::memory_order_relaxed);
  if (engine == 2 && !(tc_wanted(engine, d) && tc_domain(prv, nxt, C, d)))
    return set_error(QPWC_ERR_UNSUPPORTED, "warp_corr_fwd: tensor-core engine needs search_range 4 or 8, C %% 8 == 0 and 16-byte aligned inputs");
  // AUTO: warp kernel + cost-volume kernel through the call's scratch at every shape -- the in-kernel FFMA
  // fusion is slower than the pair of kernels at every level measured (profiles/r02b_levels.txt; 48 vs 22 us
  // at 8x7x16x196, profiles/r02c_cfg4_sweep.json) and stays the pair of the explicitly selected FFMA engine
  if (engine != 1) return warp_corr_fwd_l2(prv, nxt, flow, mode, out, B, H, W, C, d, slope, ops, st, up_scale);
  return corr_fwd_any(prv, nxt, flow, mode, out, B, H, W, C, d, slope, ops, st, up_scale);  // in-kernel fusion (FFMA)
}

// ------------------------------------------------------------------------------- host staging
// One workspace per device: NSLOT independent (stream, device buffer) slots.  A slot's stream runs
// H2D -> kernel -> D2H for one batch slice; different slots overlap (both copy engines + SMs).
struct HostStage {
  static const int NSLOT = 6;  // deep enough for the H2D engine to run ahead of the (slower, output-heavy) D2H side
  cudaStream_t stream[NSLOT] = {};
  float* buf[NSLOT] = {};
  size_t cap[NSLOT] = {};
  size_t want = 0;    // largest slot size requested so far: every (re)allocation is made at this size, so that the
                      // slots stop growing after the first pass over a workload (cudaFree/cudaMalloc synchronise the device)
  int next_slot = 0;  // keeps rotating across calls: consecutive deferred calls overlap (guarded by mu)
  std::mutex mu;
};
static HostStage g_stage[16];
// qpwc_host_set_deferred(): _host calls of THIS thread return after enqueueing (per-thread state: a
// thread that did not ask for deferral always gets completed outputs)
static thread_local bool g_host_deferred = false;

static int stage_reserve(HostStage& hs, int slot, size_t bytes) {
  if (!hs.stream[slot]) {
    const cudaError_t e = cudaStreamCreateWithFlags(&hs.stream[slot], cudaStreamNonBlocking);
    if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "host stage: stream create: %s", cudaGetErrorString(e));
  }
  if (bytes > hs.want) hs.want = bytes;
  if (hs.cap[slot] < bytes) {
    if (hs.buf[slot]) cudaFree(hs.buf[slot]);
    hs.buf[slot] = nullptr; hs.cap[slot] = 0;
    const cudaError_t e = cudaMalloc(&hs.buf[slot], hs.want);
    if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "host stage: cudaMalloc(%zu): %s", hs.want, cudaGetErrorString(e));
    hs.cap[slot] = hs.want;
  }
  return QPWC_OK;
}

// kind: 0 = corr(prv,nxt)  1 = warp(img=a, flow=f)  2 = warp_corr(prv=a, nxt=b, flow=f)
static int run_host(int kind, const float* a, const float* b, const float* f, float* out, int B,
                    int H, int W, int C, int d, float slope, int mode, int device) {
  if (device < 0 || device >= 16) return set_error(QPWC_ERR_INVALID, "host call: device ordinal %d outside [0,16)", device);
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return set_error(QPWC_ERR_CUDA, "host call: cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
  const size_t D = (size_t)(2 * d + 1) * (2 * d + 1);
  const size_t px = (size_t)H * W;
  const size_t n_a = px * C, n_b = (kind == 1) ? 0 : px * C, n_f = (kind == 0) ? 0 : px * 2;
  const size_t n_o = (kind == 1) ? px * C : px * D;
  // sub-buffers of a slot are packed back to back; each starts on a 16-byte boundary (float4 / float2
  // accesses of the kernels, TMA) whatever the tensor sizes are
  auto pad4 = [](size_t n) { return (n + 3) & ~(size_t)3; };
  const size_t item = (n_a + n_b + n_f + n_o) * sizeof(float);
  // slice the batch so that one slice moves >= ~128 MiB: every copy costs the DMA engines ~10 us of dead time,
  // and with 8 MiB slices (round 1) the 100 copies of a config-2 step added 1 ms to an 11.8 ms step
  // (tools/e2e_probe.py, one box: 8 MiB 11.76 ms, 32 MiB 10.84, 70 MiB 10.75, 600 MiB 12.26; another, with the half-sized
  // first slice: 32 MiB 11.31, 64 MiB 10.88, 96 MiB 10.58, 140 MiB 10.53).  The FIRST slice of a
  // call is half as large: the D2H engine idles until it has been copied in and computed.
  static const size_t slice_mb = [] { const char* ev = getenv("QPWC_HOST_SLICE_MB"); const long v = ev ? atol(ev) : 0; return (size_t)(v > 0 ? v : 128); }();
  int per = (int)(((slice_mb << 20) + item - 1) / item);
  if (per < 1) per = 1;
  if (per > B) per = B;
  HostStage& hs = g_stage[device];
  std::lock_guard<std::mutex> lock(hs.mu);
  int rc = QPWC_OK, slot = hs.next_slot;
  const bool l2pair = kind == 2 && g_corr_engine.load(std::memory_order_relaxed) != 1;  // warp kernel + cost-volume kernel (see warp_corr_fwd_any)
  const size_t n_w = l2pair ? n_a : 0;  // warped second frame of a slice (tensor-core engine: warp + cost volume)
  const size_t slot_floats = pad4(n_a * per) + pad4(n_b * per) + pad4(n_f * per) + pad4(n_o * per) + pad4(n_w * per);
  for (int b0 = 0, nb = 0; b0 < B && rc == QPWC_OK; b0 += nb, slot = (slot + 1) % HostStage::NSLOT) {
    nb = b0 == 0 ? (per + 1) / 2 : per;
    if (nb > B - b0) nb = B - b0;
    rc = stage_reserve(hs, slot, slot_floats * sizeof(float));
    if (rc != QPWC_OK) break;
    cudaStream_t st = hs.stream[slot];
    float* da = hs.buf[slot];
    float* db = da + pad4(n_a * per);
    float* df = db + pad4(n_b * per);
    float* dout = df + pad4(n_f * per);
    float* dw = dout + pad4(n_o * per);
    e = cudaMemcpyAsync(da, a + n_a * b0, n_a * nb * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_b) e = cudaMemcpyAsync(db, b + n_b * b0, n_b * nb * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess && n_f) e = cudaMemcpyAsync(df, f + n_f * b0, n_f * nb * sizeof(float), cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) { rc = set_error(QPWC_ERR_CUDA, "host call: H2D: %s", cudaGetErrorString(e)); break; }
    if (kind == 0) rc = corr_fwd_any(da, db, nullptr, 0, dout, nb, H, W, C, d, slope, (long long)D, st);
    else if (kind == 1) rc = launch_warp_fwd(da, df, dout, nb, H, W, C, mode, st);
    else if (l2pair) {
      rc = launch_warp_fwd(db, df, dw, nb, H, W, C, mode, st);
      if (rc == QPWC_OK) rc = corr_fwd_any(da, dw, nullptr, 0, dout, nb, H, W, C, d, slope, (long long)D, st);
    } else rc = corr_fwd_any(da, db, df, mode, dout, nb, H, W, C, d, slope, (long long)D, st);
    if (rc != QPWC_OK) break;
    e = cudaMemcpyAsync(out + n_o * b0, dout, n_o * nb * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) { rc = set_error(QPWC_ERR_CUDA, "host call: D2H: %s", cudaGetErrorString(e)); break; }
    hs.next_slot = (slot + 1) % HostStage::NSLOT;
  }
  if (g_host_deferred && rc == QPWC_OK) return rc;  // completion is collected by qpwc_host_sync()
  for (int s = 0; s < HostStage::NSLOT; ++s)
    if (hs.stream[s]) {
      e = cudaStreamSynchronize(hs.stream[s]);
      if (e != cudaSuccess && rc == QPWC_OK) rc = set_error(QPWC_ERR_CUDA, "host call: sync: %s", cudaGetErrorString(e));
    }
  return rc;
}

static int host_sync(int device) {
  // device == -1: every device that has received host-buffer work from this process
  if (device < -1 || device >= 16) return set_error(QPWC_ERR_INVALID, "qpwc_host_sync: device ordinal %d outside [-1,16)", device);
  int rc = QPWC_OK;
  for (int dv = (device < 0 ? 0 : device); dv <= (device < 0 ? 15 : device); ++dv) {
    HostStage& hs = g_stage[dv];
    std::lock_guard<std::mutex> lock(hs.mu);
    for (int s = 0; s < HostStage::NSLOT; ++s)
      if (hs.stream[s]) {
        const cudaError_t e = cudaStreamSynchronize(hs.stream[s]);
        if (e != cudaSuccess && rc == QPWC_OK) rc = set_error(QPWC_ERR_CUDA, "qpwc_host_sync(device %d): %s", dv, cudaGetErrorString(e));
      }
  }
  return rc;
}

}  // namespace qpwc

using namespace qpwc;

extern "C" {

int qpwc_version(void) { return 200; /* 0.2.0 */ }
int qpwc_set_option(int key, int value) {
  if (key == QPWC_OPT_CORR_ENGINE && value >= 0 && value <= 2) { g_corr_engine.store(value); return QPWC_OK; }
  if (key == QPWC_OPT_WARP_BWD && value >= 0 && value <= 2) { set_warp_bwd_variant(value); return QPWC_OK; }
  if (key == QPWC_OPT_CORR_BWD && value >= 0 && value <= 1) { set_corr_bwd_variant(value); return QPWC_OK; }
  return set_error(QPWC_ERR_INVALID, "qpwc_set_option: unknown key %d or value %d", key, value);
}
int qpwc_get_option(int key) {
  return key == QPWC_OPT_CORR_ENGINE ? g_corr_engine.load() : (key == QPWC_OPT_WARP_BWD ? get_warp_bwd_variant() : (key == QPWC_OPT_CORR_BWD ? get_corr_bwd_variant() : -1));
}
int qpwc_host_set_deferred(int on) { g_host_deferred = on != 0; return QPWC_OK; }
int qpwc_host_sync(int device) { return host_sync(device); }
const char* qpwc_last_error(void) { return g_err; }

int qpwc_corr_fwd(const float* prv, const float* nxt, float* out, int B, int H, int W, int C,
                  int search_range, float leaky_slope, long long out_pixel_stride, void* stream) {
  const char* fn = "qpwc_corr_fwd";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  QPWC_TRY(check_corr_args(fn, search_range, out_pixel_stride));
  if (B == 0 || H == 0 || W == 0) return QPWC_OK;
  if (C == 0) return set_error(QPWC_ERR_INVALID, "%s: C == 0 (mean over an empty channel axis)", fn);
  QPWC_TRY(check_ptr(fn, "prv", prv)); QPWC_TRY(check_ptr(fn, "nxt", nxt)); QPWC_TRY(check_ptr(fn, "out", out));
  return corr_fwd_any(prv, nxt, nullptr, 0, out, B, H, W, C, search_range, leaky_slope, out_pixel_stride, (cudaStream_t)stream);
}

int qpwc_corr_bwd(const float* prv, const float* nxt, const float* out, const float* g_out,
                  float* g_prv, float* g_nxt, int B, int H, int W, int C, int search_range,
                  float leaky_slope, long long out_pixel_stride, void* stream) {
  const char* fn = "qpwc_corr_bwd";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  QPWC_TRY(check_corr_args(fn, search_range, out_pixel_stride));
  if (empty(B, H, W, C)) return QPWC_OK;
  QPWC_TRY(check_ptr(fn, "prv", prv)); QPWC_TRY(check_ptr(fn, "nxt", nxt)); QPWC_TRY(check_ptr(fn, "out", out));
  QPWC_TRY(check_ptr(fn, "g_out", g_out)); QPWC_TRY(check_ptr(fn, "g_prv", g_prv)); QPWC_TRY(check_ptr(fn, "g_nxt", g_nxt));
  return corr_bwd_any(prv, nxt, out, g_out, g_prv, g_nxt, B, H, W, C, search_range, leaky_slope, out_pixel_stride, (cudaStream_t)stream);
}

int qpwc_warp_fwd(const float* img, const float* flow, float* out, int B, int H, int W, int C,
                  int mode, void* stream) {
  const char* fn = "qpwc_warp_fwd";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  if (empty(B, H, W, C)) return QPWC_OK;
  QPWC_TRY(check_mode(fn, mode, H, W));
  QPWC_TRY(check_ptr(fn, "img", img)); QPWC_TRY(check_ptr(fn, "flow", flow)); QPWC_TRY(check_ptr(fn, "out", out));
  if (reinterpret_cast<uintptr_t>(flow) % 8) return set_error(QPWC_ERR_INVALID, "%s: flow must be 8-byte aligned", fn);
  return launch_warp_fwd(img, flow, out, B, H, W, C, mode, (cudaStream_t)stream);
}

int qpwc_warp_bwd(const float* img, const float* flow, const float* g_out, float* g_img,
                  float* g_flow, int B, int H, int W, int C, int mode, void* stream) {
  const char* fn = "qpwc_warp_bwd";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  if (B == 0 || H == 0 || W == 0) return QPWC_OK;
  QPWC_TRY(check_mode(fn, mode, H, W));
  QPWC_TRY(check_ptr(fn, "flow", flow)); QPWC_TRY(check_ptr(fn, "g_flow", g_flow));
  if (reinterpret_cast<uintptr_t>(flow) % 8 || reinterpret_cast<uintptr_t>(g_flow) % 8)
    return set_error(QPWC_ERR_INVALID, "%s: flow and g_flow must be 8-byte aligned", fn);
  if (C == 0) {
    const cudaError_t e = cudaMemsetAsync(g_flow, 0, sizeof(float) * 2 * (size_t)B * H * W, (cudaStream_t)stream);
    return e == cudaSuccess ? QPWC_OK : set_error(QPWC_ERR_CUDA, "%s: memset: %s", fn, cudaGetErrorString(e));
  }
  QPWC_TRY(check_ptr(fn, "img", img)); QPWC_TRY(check_ptr(fn, "g_out", g_out)); QPWC_TRY(check_ptr(fn, "g_img", g_img));
  return launch_warp_bwd(img, flow, g_out, g_img, g_flow, B, H, W, C, mode, (cudaStream_t)stream);
}

static int check_warp_stride(const char* fn, const char* what, long long stride, long long need) {
  if (stride < need) return set_error(QPWC_ERR_INVALID, "%s: %s %lld < %lld", fn, what, stride, need);
  return QPWC_OK;
}

int qpwc_warp_fwd_ex(const float* img, const float* flow, float* out, int B, int H, int W, int C,
                     int mode, float flow_scale, long long out_pixel_stride, void* stream) {
  const char* fn = "qpwc_warp_fwd_ex";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  QPWC_TRY(check_warp_stride(fn, "out_pixel_stride", out_pixel_stride, C));
  if (empty(B, H, W, C)) return QPWC_OK;
  QPWC_TRY(check_mode(fn, mode, H, W));
  QPWC_TRY(check_ptr(fn, "img", img)); QPWC_TRY(check_ptr(fn, "flow", flow)); QPWC_TRY(check_ptr(fn, "out", out));
  if (reinterpret_cast<uintptr_t>(flow) % 8) return set_error(QPWC_ERR_INVALID, "%s: flow must be 8-byte aligned", fn);
  return launch_warp_fwd_ex(img, flow, nullptr, nullptr, out, B, H, W, C, mode, flow_scale, out_pixel_stride, (cudaStream_t)stream);
}

int qpwc_warp_fwd_rows(const float* img, const float* flow, float* out, int B, int H, int W, int C,
                       int mode, int row_offset, int full_height, void* stream) {
  const char* fn = "qpwc_warp_fwd_rows";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  if (row_offset < 0 || full_height < row_offset + H)
    return set_error(QPWC_ERR_INVALID, "%s: rows [%d, %d) are not inside an image of %d rows", fn, row_offset, row_offset + H, full_height);
  if (empty(B, H, W, C)) return QPWC_OK;
  QPWC_TRY(check_mode(fn, mode, full_height, W));
  QPWC_TRY(check_ptr(fn, "img", img)); QPWC_TRY(check_ptr(fn, "flow", flow)); QPWC_TRY(check_ptr(fn, "out", out));
  if (reinterpret_cast<uintptr_t>(flow) % 8) return set_error(QPWC_ERR_INVALID, "%s: flow must be 8-byte aligned", fn);
  return launch_warp_fwd_ex(img, flow, nullptr, nullptr, out, B, H, W, C, mode, 1.f, C, (cudaStream_t)stream, 0.f, row_offset, full_height);
}

int qpwc_warp_pair_fwd(const float* img_a, const float* flow_a, const float* img_b, const float* flow_b,
                       float* out, int B, int H, int W, int C, int mode, float flow_scale,
                       long long out_pixel_stride, void* stream) {
  const char* fn = "qpwc_warp_pair_fwd";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  QPWC_TRY(check_warp_stride(fn, "out_pixel_stride", out_pixel_stride, 2LL * C));
  if (empty(B, H, W, C)) return QPWC_OK;
  QPWC_TRY(check_mode(fn, mode, H, W));
  QPWC_TRY(check_ptr(fn, "img_a", img_a)); QPWC_TRY(check_ptr(fn, "flow_a", flow_a));
  QPWC_TRY(check_ptr(fn, "img_b", img_b)); QPWC_TRY(check_ptr(fn, "flow_b", flow_b)); QPWC_TRY(check_ptr(fn, "out", out));
  if (reinterpret_cast<uintptr_t>(flow_a) % 8 || reinterpret_cast<uintptr_t>(flow_b) % 8)
    return set_error(QPWC_ERR_INVALID, "%s: flows must be 8-byte aligned", fn);
  return launch_warp_fwd_ex(img_a, flow_a, img_b, flow_b, out, B, H, W, C, mode, flow_scale, out_pixel_stride, (cudaStream_t)stream);
}

int qpwc_warp_bwd_ex(const float* img, const float* flow, const float* g_out, float* g_img,
                     float* g_flow, int B, int H, int W, int C, int mode, float flow_scale,
                     long long g_out_pixel_stride, void* stream) {
  const char* fn = "qpwc_warp_bwd_ex";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  QPWC_TRY(check_warp_stride(fn, "g_out_pixel_stride", g_out_pixel_stride, C));
  if (B == 0 || H == 0 || W == 0) return QPWC_OK;
  QPWC_TRY(check_mode(fn, mode, H, W));
  QPWC_TRY(check_ptr(fn, "flow", flow)); QPWC_TRY(check_ptr(fn, "g_flow", g_flow));
  if (reinterpret_cast<uintptr_t>(flow) % 8 || reinterpret_cast<uintptr_t>(g_flow) % 8)
    return set_error(QPWC_ERR_INVALID, "%s: flow and g_flow must be 8-byte aligned", fn);
  if (C == 0) {
    const cudaError_t e = cudaMemsetAsync(g_flow, 0, sizeof(float) * 2 * (size_t)B * H * W, (cudaStream_t)stream);
    return e == cudaSuccess ? QPWC_OK : set_error(QPWC_ERR_CUDA, "%s: memset: %s", fn, cudaGetErrorString(e));
  }
  QPWC_TRY(check_ptr(fn, "img", img)); QPWC_TRY(check_ptr(fn, "g_out", g_out)); QPWC_TRY(check_ptr(fn, "g_img", g_img));
  return launch_warp_bwd_ex(img, flow, g_out, g_img, g_flow, B, H, W, C, mode, flow_scale, g_out_pixel_stride, (cudaStream_t)stream);
}

int qpwc_corr_fwd_nchw(const float* prv, const float* nxt, float* out, int B, int C, int H, int W,
                       int search_range, float leaky_slope, void* stream) {
  const char* fn = "qpwc_corr_fwd_nchw";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  QPWC_TRY(check_corr_args(fn, search_range, (long long)(2 * search_range + 1) * (2 * search_range + 1)));
  if (B == 0 || H == 0 || W == 0) return QPWC_OK;
  if (C == 0) return set_error(QPWC_ERR_INVALID, "%s: C == 0 (mean over an empty channel axis)", fn);
  QPWC_TRY(check_ptr(fn, "prv", prv)); QPWC_TRY(check_ptr(fn, "nxt", nxt)); QPWC_TRY(check_ptr(fn, "out", out));
  const int rc = launch_corr_fwd_nchw(prv, nxt, out, B, C, H, W, search_range, leaky_slope, (cudaStream_t)stream);
  if (rc == QPWC_ERR_UNSUPPORTED)
    return set_error(QPWC_ERR_UNSUPPORTED, "%s: native channels_first kernel needs search_range 4, W %% 4 == 0 and 16-byte aligned tensors "
                                           "(transpose to NHWC and call qpwc_corr_fwd instead)", fn);
  return rc;
}

int qpwc_corr_bwd_nchw(const float* prv, const float* nxt, const float* out, const float* g_out,
                       float* g_prv, float* g_nxt, int B, int C, int H, int W, int search_range,
                       float leaky_slope, void* stream) {
  const char* fn = "qpwc_corr_bwd_nchw";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  QPWC_TRY(check_corr_args(fn, search_range, (long long)(2 * search_range + 1) * (2 * search_range + 1)));
  if (B == 0 || H == 0 || W == 0 || C == 0) return QPWC_OK;
  QPWC_TRY(check_ptr(fn, "prv", prv)); QPWC_TRY(check_ptr(fn, "nxt", nxt)); QPWC_TRY(check_ptr(fn, "out", out));
  QPWC_TRY(check_ptr(fn, "g_out", g_out)); QPWC_TRY(check_ptr(fn, "g_prv", g_prv)); QPWC_TRY(check_ptr(fn, "g_nxt", g_nxt));
  const int rc = launch_corr_bwd_nchw(prv, nxt, out, g_out, g_prv, g_nxt, B, C, H, W, search_range, leaky_slope, (cudaStream_t)stream);
  if (rc == QPWC_ERR_UNSUPPORTED)
    return set_error(QPWC_ERR_UNSUPPORTED, "%s: native channels_first kernel needs search_range 4, W %% 4 == 0 and 16-byte aligned tensors "
                                           "(transpose to NHWC and call qpwc_corr_bwd instead)", fn);
  return rc;
}

int qpwc_warp_fwd_nchw(const float* img, const float* flow, float* out, int B, int C, int H, int W,
                       int mode, float flow_scale, void* stream) {
  const char* fn = "qpwc_warp_fwd_nchw";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  if (empty(B, H, W, C)) return QPWC_OK;
  QPWC_TRY(check_mode(fn, mode, H, W));
  QPWC_TRY(check_ptr(fn, "img", img)); QPWC_TRY(check_ptr(fn, "flow", flow)); QPWC_TRY(check_ptr(fn, "out", out));
  return launch_warp_fwd_nchw(img, flow, out, B, C, H, W, mode, flow_scale, (cudaStream_t)stream);
}

int qpwc_warp_bwd_nchw(const float* img, const float* flow, const float* g_out, float* g_img,
                       float* g_flow, int B, int C, int H, int W, int mode, float flow_scale, void* stream) {
  const char* fn = "qpwc_warp_bwd_nchw";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  if (empty(B, H, W, C)) return QPWC_OK;
  QPWC_TRY(check_mode(fn, mode, H, W));
  QPWC_TRY(check_ptr(fn, "img", img)); QPWC_TRY(check_ptr(fn, "flow", flow)); QPWC_TRY(check_ptr(fn, "g_out", g_out));
  QPWC_TRY(check_ptr(fn, "g_img", g_img)); QPWC_TRY(check_ptr(fn, "g_flow", g_flow));
  return launch_warp_bwd_nchw(img, flow, g_out, g_img, g_flow, B, C, H, W, mode, flow_scale, (cudaStream_t)stream);
}

static int check_up(const char* fn, int H, int W, float up_scale) {
  if ((H & 1) || (W & 1)) return set_error(QPWC_ERR_INVALID, "%s: H and W must be even (x2 upsampled flow), got %dx%d", fn, H, W);
  if (up_scale == 0.f) return set_error(QPWC_ERR_INVALID, "%s: up_scale must be non-zero", fn);
  return QPWC_OK;
}

int qpwc_upsample2x_fwd(const float* src, float* dst, int B, int H, int W, int C, float scale, void* stream) {
  const char* fn = "qpwc_upsample2x_fwd";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  if (empty(B, H, W, C)) return QPWC_OK;
  QPWC_TRY(check_ptr(fn, "src", src)); QPWC_TRY(check_ptr(fn, "dst", dst));
  return launch_upsample2x_fwd(src, dst, B, H, W, C, scale, (cudaStream_t)stream);
}

int qpwc_upsample2x_bwd(const float* g_dst, float* g_src, int B, int H, int W, int C, float scale, void* stream) {
  const char* fn = "qpwc_upsample2x_bwd";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  if (empty(B, H, W, C)) return QPWC_OK;
  QPWC_TRY(check_ptr(fn, "g_dst", g_dst)); QPWC_TRY(check_ptr(fn, "g_src", g_src));
  return launch_upsample2x_bwd(g_dst, g_src, B, H, W, C, scale, (cudaStream_t)stream);
}

int qpwc_occlusion_map(const float* flow, float* out, int B, int H, int W, int channels_first, void* stream) {
  const char* fn = "qpwc_occlusion_map";
  QPWC_TRY(check_shape(fn, B, H, W, 2));
  if (empty(B, H, W, 2)) return QPWC_OK;
  QPWC_TRY(check_ptr(fn, "flow", flow)); QPWC_TRY(check_ptr(fn, "out", out));
  if (!channels_first && reinterpret_cast<uintptr_t>(flow) % 8) return set_error(QPWC_ERR_INVALID, "%s: flow must be 8-byte aligned", fn);
  if ((long long)H * W > 0x7fffffffLL) return set_error(QPWC_ERR_UNSUPPORTED, "%s: H*W exceeds 2^31-1", fn);
  return launch_occlusion_map(flow, out, B, H, W, channels_first ? 1 : 0, (cudaStream_t)stream);
}

int qpwc_warp_fwd_up(const float* img, const float* flow_coarse, float* out, int B, int H, int W, int C,
                     int mode, float up_scale, void* stream) {
  const char* fn = "qpwc_warp_fwd_up";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  QPWC_TRY(check_up(fn, H, W, up_scale));
  if (empty(B, H, W, C)) return QPWC_OK;
  QPWC_TRY(check_mode(fn, mode, H, W));
  QPWC_TRY(check_ptr(fn, "img", img)); QPWC_TRY(check_ptr(fn, "flow_coarse", flow_coarse)); QPWC_TRY(check_ptr(fn, "out", out));
  if (reinterpret_cast<uintptr_t>(flow_coarse) % 8) return set_error(QPWC_ERR_INVALID, "%s: flow must be 8-byte aligned", fn);
  return launch_warp_fwd_ex(img, flow_coarse, nullptr, nullptr, out, B, H, W, C, mode, 1.f, C, (cudaStream_t)stream, up_scale);
}

int qpwc_warp_corr_fwd_up(const float* prv, const float* nxt, const float* flow_coarse, float* out, int B,
                          int H, int W, int C, int search_range, float leaky_slope, int mode,
                          long long out_pixel_stride, float up_scale, void* stream) {
  const char* fn = "qpwc_warp_corr_fwd_up";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  QPWC_TRY(check_corr_args(fn, search_range, out_pixel_stride));
  QPWC_TRY(check_up(fn, H, W, up_scale));
  if (B == 0 || H == 0 || W == 0) return QPWC_OK;
  if (C == 0) return set_error(QPWC_ERR_INVALID, "%s: C == 0", fn);
  QPWC_TRY(check_mode(fn, mode, H, W));
  QPWC_TRY(check_ptr(fn, "prv", prv)); QPWC_TRY(check_ptr(fn, "nxt", nxt)); QPWC_TRY(check_ptr(fn, "flow_coarse", flow_coarse));
  QPWC_TRY(check_ptr(fn, "out", out));
  if (reinterpret_cast<uintptr_t>(flow_coarse) % 8) return set_error(QPWC_ERR_INVALID, "%s: flow must be 8-byte aligned", fn);
  return warp_corr_fwd_any(prv, nxt, flow_coarse, mode, out, B, H, W, C, search_range, leaky_slope, out_pixel_stride,
                           (cudaStream_t)stream, up_scale);
}

int qpwc_warp_corr_fwd(const float* prv, const float* nxt, const float* flow, float* out, int B,
                       int H, int W, int C, int search_range, float leaky_slope, int mode,
                       long long out_pixel_stride, void* stream) {
  const char* fn = "qpwc_warp_corr_fwd";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  QPWC_TRY(check_corr_args(fn, search_range, out_pixel_stride));
  if (B == 0 || H == 0 || W == 0) return QPWC_OK;
  if (C == 0) return set_error(QPWC_ERR_INVALID, "%s: C == 0", fn);
  QPWC_TRY(check_mode(fn, mode, H, W));
  QPWC_TRY(check_ptr(fn, "prv", prv)); QPWC_TRY(check_ptr(fn, "nxt", nxt));
  QPWC_TRY(check_ptr(fn, "flow", flow)); QPWC_TRY(check_ptr(fn, "out", out));
  if (reinterpret_cast<uintptr_t>(flow) % 8) return set_error(QPWC_ERR_INVALID, "%s: flow must be 8-byte aligned", fn);
  return warp_corr_fwd_any(prv, nxt, flow, mode, out, B, H, W, C, search_range, leaky_slope, out_pixel_stride, (cudaStream_t)stream);
}

size_t qpwc_warp_corr_bwd_workspace(int B, int H, int W, int C) {
  (void)B; (void)H; (void)W; (void)C;
  return 0;  // since 0.2: the intermediate lives in a stream-ordered, L2-sized scratch owned by the call
}

int qpwc_warp_corr_bwd(const float* prv, const float* nxt, const float* flow, const float* out,
                       const float* g_out, float* g_prv, float* g_nxt, float* g_flow,
                       void* workspace, size_t workspace_bytes, int B, int H, int W, int C,
                       int search_range, float leaky_slope, int mode, long long out_pixel_stride,
                       void* stream) {
  const char* fn = "qpwc_warp_corr_bwd";
  (void)workspace; (void)workspace_bytes;  // accepted for source compatibility with 0.1, unused
  QPWC_TRY(check_shape(fn, B, H, W, C));
  QPWC_TRY(check_corr_args(fn, search_range, out_pixel_stride));
  if (empty(B, H, W, C)) return QPWC_OK;
  QPWC_TRY(check_mode(fn, mode, H, W));
  QPWC_TRY(check_ptr(fn, "prv", prv)); QPWC_TRY(check_ptr(fn, "nxt", nxt)); QPWC_TRY(check_ptr(fn, "flow", flow));
  QPWC_TRY(check_ptr(fn, "out", out)); QPWC_TRY(check_ptr(fn, "g_out", g_out)); QPWC_TRY(check_ptr(fn, "g_prv", g_prv));
  QPWC_TRY(check_ptr(fn, "g_nxt", g_nxt)); QPWC_TRY(check_ptr(fn, "g_flow", g_flow));
  if (reinterpret_cast<uintptr_t>(flow) % 8 || reinterpret_cast<uintptr_t>(g_flow) % 8)
    return set_error(QPWC_ERR_INVALID, "%s: flow and g_flow must be 8-byte aligned", fn);
  cudaStream_t st = (cudaStream_t)stream;
  // Chain rule over the two stages, a chunk of frame pairs at a time: the warped frame and its
  // gradient (2 x chunk x H x W x C floats) are rebuilt in a scratch buffer that is reused chunk after
  // chunk and sized to stay resident in L2 -- neither tensor is ever materialised in HBM for the
  // whole batch, and the caller provides no workspace.
  const size_t item = (size_t)H * W * C;
  const int per = chunk_batch(B, 2 * item * sizeof(float));
  float* scratch = nullptr;
  int rc = scratch_alloc(&scratch, 2 * item * per * sizeof(float), st, fn);
  if (rc != QPWC_OK) return rc;
  cudaError_t e;
  float* nxt_w = scratch;
  float* g_nxt_w = scratch + item * per;
  for (int b0 = 0; b0 < B && rc == QPWC_OK; b0 += per) {
    const int nb = B - b0 < per ? B - b0 : per;
    const size_t o = item * b0, of = (size_t)H * W * 2 * b0, oo = (size_t)H * W * out_pixel_stride * b0;
    rc = launch_warp_fwd(nxt + o, flow + of, nxt_w, nb, H, W, C, mode, st);
    if (rc == QPWC_OK) rc = corr_bwd_any(prv + o, nxt_w, out + oo, g_out + oo, g_prv + o, g_nxt_w, nb, H, W, C, search_range, leaky_slope, out_pixel_stride, st);
    if (rc == QPWC_OK) rc = launch_warp_bwd(nxt + o, flow + of, g_nxt_w, g_nxt + o, g_flow + of, nb, H, W, C, mode, st);
  }
  e = scratch_free(scratch, st);
  if (rc == QPWC_OK && e != cudaSuccess) rc = set_error(QPWC_ERR_CUDA, "%s: scratch free: %s", fn, cudaGetErrorString(e));
  return rc;
}

int qpwc_corr_fwd_host(const float* prv, const float* nxt, float* out, int B, int H, int W, int C,
                       int search_range, float leaky_slope, int device) {
  const char* fn = "qpwc_corr_fwd_host";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  QPWC_TRY(check_corr_args(fn, search_range, (long long)(2 * search_range + 1) * (2 * search_range + 1)));
  if (B == 0 || H == 0 || W == 0) return QPWC_OK;
  if (C == 0) return set_error(QPWC_ERR_INVALID, "%s: C == 0", fn);
  QPWC_TRY(check_ptr(fn, "prv", prv)); QPWC_TRY(check_ptr(fn, "nxt", nxt)); QPWC_TRY(check_ptr(fn, "out", out));
  return run_host(0, prv, nxt, nullptr, out, B, H, W, C, search_range, leaky_slope, 0, device);
}

int qpwc_warp_fwd_host(const float* img, const float* flow, float* out, int B, int H, int W, int C,
                       int mode, int device) {
  const char* fn = "qpwc_warp_fwd_host";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  if (empty(B, H, W, C)) return QPWC_OK;
  QPWC_TRY(check_mode(fn, mode, H, W));
  QPWC_TRY(check_ptr(fn, "img", img)); QPWC_TRY(check_ptr(fn, "flow", flow)); QPWC_TRY(check_ptr(fn, "out", out));
  return run_host(1, img, nullptr, flow, out, B, H, W, C, 0, 0.f, mode, device);
}

int qpwc_warp_corr_fwd_host(const float* prv, const float* nxt, const float* flow, float* out,
                            int B, int H, int W, int C, int search_range, float leaky_slope,
                            int mode, int device) {
  const char* fn = "qpwc_warp_corr_fwd_host";
  QPWC_TRY(check_shape(fn, B, H, W, C));
  QPWC_TRY(check_corr_args(fn, search_range, (long long)(2 * search_range + 1) * (2 * search_range + 1)));
  if (B == 0 || H == 0 || W == 0) return QPWC_OK;
  if (C == 0) return set_error(QPWC_ERR_INVALID, "%s: C == 0", fn);
  QPWC_TRY(check_mode(fn, mode, H, W));
  QPWC_TRY(check_ptr(fn, "prv", prv)); QPWC_TRY(check_ptr(fn, "nxt", nxt));
  QPWC_TRY(check_ptr(fn, "flow", flow)); QPWC_TRY(check_ptr(fn, "out", out));
  return run_host(2, prv, nxt, flow, out, B, H, W, C, search_range, leaky_slope, mode, device);
}

}  // extern "C"
