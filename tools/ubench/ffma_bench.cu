// Micro-benchmark (dev tool, first half of round 1): issue cost of the 9x9 outer-product step in three
// formulations.  NOTE: K1 issues its four FMAs per accumulator back to back (dependent), so its "scalar"
// figure is latency-, not issue-bound; ffma2_patterns.cu / corr_loop_bench.cu supersede these numbers.
//   K1 scalar FFMA, float4 operands               (current kernel's inner step)
//   K2 FFMA2, accumulator pairs over m (rows), duplicated a                ("m-pairs")
//   K3 FFMA2, even/odd-channel accumulator pairs, 5x9 outputs per thread   ("even/odd")
// Prints cycles per FMA per warp (1.0 = one FMA per lane per cycle per SMSP = FP32 peak).
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2000

__global__ void k1(const float4* __restrict__ in, float* out, long long* cyc) {
  float4 a[9], b[9];
  for (int i = 0; i < 9; ++i) { a[i] = in[threadIdx.x * 18 + i]; b[i] = in[threadIdx.x * 18 + 9 + i]; }
  float acc[9][9];
#pragma unroll
  for (int m = 0; m < 9; ++m)
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[m][k] = 0.f;
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int m = 0; m < 9; ++m) {
        acc[m][k] = fmaf(a[k].x, b[m].x, acc[m][k]);
        acc[m][k] = fmaf(a[k].y, b[m].y, acc[m][k]);
        acc[m][k] = fmaf(a[k].z, b[m].z, acc[m][k]);
        acc[m][k] = fmaf(a[k].w, b[m].w, acc[m][k]);
      }
    // perturb operands a little so nothing is hoisted
    a[it % 9].x += 1e-9f;
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int m = 0; m < 9; ++m)
#pragma unroll
    for (int k = 0; k < 9; ++k) s += acc[m][k];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// K2: pairs over m: accp[p][k] = (acc[2p][k], acc[2p+1][k]), p = 0..3, plus row 8 scalar.
// b operands arrive as pairs (b_{2p}.c, b_{2p+1}.c); a is duplicated (a.c, a.c).
__global__ void k2(const float4* __restrict__ in, float* out, long long* cyc) {
  float4 a[9];
  float2 bp[4][4];   // [pair][channel]
  float4 b8;
  for (int i = 0; i < 9; ++i) a[i] = in[threadIdx.x * 18 + i];
  for (int p = 0; p < 4; ++p) { float4 u = in[threadIdx.x * 18 + 9 + 2 * p], v = in[threadIdx.x * 18 + 10 + 2 * p];
    bp[p][0] = make_float2(u.x, v.x); bp[p][1] = make_float2(u.y, v.y); bp[p][2] = make_float2(u.z, v.z); bp[p][3] = make_float2(u.w, v.w); }
  b8 = in[threadIdx.x * 18 + 17];
  float2 accp[4][9];
  float acc8[9];
#pragma unroll
  for (int k = 0; k < 9; ++k) { acc8[k] = 0.f;
#pragma unroll
    for (int p = 0; p < 4; ++p) accp[p][k] = make_float2(0.f, 0.f); }
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      const float ac[4] = {a[k].x, a[k].y, a[k].z, a[k].w};
      const float bc[4] = {b8.x, b8.y, b8.z, b8.w};
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const float2 ad = make_float2(ac[c], ac[c]);
#pragma unroll
        for (int p = 0; p < 4; ++p) accp[p][k] = __ffma2_rn(ad, bp[p][c], accp[p][k]);
        acc8[k] = fmaf(ac[c], bc[c], acc8[k]);
      }
    }
    a[it % 9].x += 1e-9f;
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < 9; ++k) { s += acc8[k];
#pragma unroll
    for (int p = 0; p < 4; ++p) s += accp[p][k].x + accp[p][k].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

// K3: even/odd channel accumulators, 5 rows x 9 columns per thread (in-warp m-split)
__global__ void k3(const float4* __restrict__ in, float* out, long long* cyc) {
  float4 a[9], b[5];
  for (int i = 0; i < 9; ++i) a[i] = in[threadIdx.x * 18 + i];
  for (int i = 0; i < 5; ++i) b[i] = in[threadIdx.x * 18 + 9 + i];
  float2 acc[5][9];
#pragma unroll
  for (int m = 0; m < 5; ++m)
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[m][k] = make_float2(0.f, 0.f);
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int k = 0; k < 9; ++k)
#pragma unroll
      for (int m = 0; m < 5; ++m) {
        acc[m][k] = __ffma2_rn(make_float2(a[k].x, a[k].y), make_float2(b[m].x, b[m].y), acc[m][k]);
        acc[m][k] = __ffma2_rn(make_float2(a[k].z, a[k].w), make_float2(b[m].z, b[m].w), acc[m][k]);
      }
    a[it % 9].x += 1e-9f;
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int m = 0; m < 5; ++m)
#pragma unroll
    for (int k = 0; k < 9; ++k) s += acc[m][k].x + acc[m][k].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
  float4* in; float* out; long long* cyc;
  const int blocks = 148, threads = 128;   // 1 warp per SMSP: pure issue cost, no co-scheduling
  cudaMalloc(&in, sizeof(float4) * 18 * threads); cudaMemset(in, 0, sizeof(float4) * 18 * threads);
  cudaMalloc(&out, sizeof(float) * blocks * threads); cudaMalloc(&cyc, sizeof(long long) * blocks);
  long long h[148];
  for (int w = 0; w < 2; ++w) {
    for (int nthr = 128; nthr <= 384; nthr += 128) {
      k1<<<blocks, nthr>>>(in, out, cyc); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c1 = (double)h[0] / ITERS / 324.0;
      k2<<<blocks, nthr>>>(in, out, cyc); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c2 = (double)h[0] / ITERS / 324.0;
      k3<<<blocks, nthr>>>(in, out, cyc); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
      double c3 = (double)h[0] / ITERS / 180.0;
      if (w) printf("warps/SMSP %d: cycles per FMA (per warp): scalar %.3f  ffma2 m-pairs %.3f  ffma2 even/odd(5x9) %.3f   [x warps/SMSP = SMSP cycles per warp-FMA]\n",
                    nthr / 128, c1, c2, c3);
    }
  }
  printf("%s %s\n", cudaGetErrorString(cudaGetLastError()), cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
