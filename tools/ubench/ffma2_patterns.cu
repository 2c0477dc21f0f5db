// Micro-benchmark (dev tool): issue rate of FFMA2 (fma.rn.f32x2) operand patterns, registers only.
// Reports cycles per FFMA2 per SMSP (2.0 = FP32 peak) at 1 and 2 warps per SMSP.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4000
#define F2(a, b, c) __ffma2_rn((a), (b), (c))
#define BEGIN(NACC)                                                          \
  float2 acc[NACC];                                                          \
  _Pragma("unroll") for (int i = 0; i < NACC; ++i) acc[i] = make_float2(0.f, 0.f); \
  float2 v[20];                                                              \
  _Pragma("unroll") for (int i = 0; i < 20; ++i) v[i] = in[threadIdx.x * 20 + i]; \
  long long t0 = clock64();                                                  \
  for (int it = 0; it < ITERS; ++it) {
#define END(NACC)                                                            \
    v[0].x += 1e-9f; v[9].y += 1e-9f;                                                   \
  }                                                                          \
  long long t1 = clock64();                                                  \
  float s = 0.f;                                                             \
  _Pragma("unroll") for (int i = 0; i < NACC; ++i) s += acc[i].x + 2.f * acc[i].y; \
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;                            \
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;

// P1: both multiplicands fixed
__global__ void p1(const float2* __restrict__ in, float* out, long long* cyc) {
  BEGIN(72)
#pragma unroll
  for (int j = 0; j < 72; ++j) acc[j] = F2(v[0], v[1], acc[j]);
  END(72)
}
// P2: outer product 9 x 8 with pair operands, inner loop over the 8 (a fixed per group)
__global__ void p2(const float2* __restrict__ in, float* out, long long* cyc) {
  BEGIN(72)
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int m = 0; m < 8; ++m) acc[k * 8 + m] = F2(v[k], v[9 + m], acc[k * 8 + m]);
  END(72)
}
// P4: scalar-broadcast x pair, inner loop over the scalars (pair fixed per group of 8)
__global__ void p4(const float2* __restrict__ in, float* out, long long* cyc) {
  BEGIN(72)
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int m = 0; m < 8; ++m) {
      const float sc = (m & 1) ? v[9 + (m >> 1)].y : v[9 + (m >> 1)].x;
      acc[k * 8 + m] = F2(make_float2(sc, sc), v[k], acc[k * 8 + m]);
    }
  END(72)
}
// P5: scalar-broadcast x pair, inner loop over the pairs (scalar fixed per group of 9)
__global__ void p5(const float2* __restrict__ in, float* out, long long* cyc) {
  BEGIN(72)
#pragma unroll
  for (int m = 0; m < 8; ++m) {
    const float sc = (m & 1) ? v[9 + (m >> 1)].y : v[9 + (m >> 1)].x;
#pragma unroll
    for (int k = 0; k < 9; ++k) acc[k * 8 + m] = F2(make_float2(sc, sc), v[k], acc[k * 8 + m]);
  }
  END(72)
}
// P6: scalar FFMA 3-register, 72 accumulators (outer product), for reference
__global__ void p6(const float2* __restrict__ in, float* out, long long* cyc) {
  BEGIN(36)
#pragma unroll
  for (int k = 0; k < 9; ++k)
#pragma unroll
    for (int m = 0; m < 4; ++m) {
      acc[k * 4 + m].x = fmaf(v[k].x, v[9 + m].x, acc[k * 4 + m].x);
      acc[k * 4 + m].y = fmaf(v[k].x, v[9 + m].y, acc[k * 4 + m].y);
    }
  END(36)
}
int main() {
  float2* in; float* out; long long* cyc;
  cudaMalloc(&in, sizeof(float2) * 20 * 256); cudaMemset(in, 0, sizeof(float2) * 20 * 256);
  cudaMalloc(&out, sizeof(float) * 148 * 256); cudaMalloc(&cyc, sizeof(long long) * 148);
  long long h[148];
  for (int nthr = 128; nthr <= 256; nthr += 128) {
    const int wps = nthr / 128;
    double r[5];
    for (int w = 0; w < 2; ++w) {
      p1<<<148, nthr>>>(in, out, cyc); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); r[0] = (double)h[0] / ITERS / 72 / wps;
      p2<<<148, nthr>>>(in, out, cyc); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); r[1] = (double)h[0] / ITERS / 72 / wps;
      p4<<<148, nthr>>>(in, out, cyc); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); r[2] = (double)h[0] / ITERS / 72 / wps;
      p5<<<148, nthr>>>(in, out, cyc); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); r[3] = (double)h[0] / ITERS / 72 / wps;
      p6<<<148, nthr>>>(in, out, cyc); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); r[4] = (double)h[0] / ITERS / 72 / wps;
    }
    printf("warps/SMSP %d: SMSP cycles per instr: P1 fixed a,b %.2f | P2 pairs outer %.2f | P4 bcast (pair reused) %.2f | P5 bcast (scalar reused) %.2f | P6 scalar FFMA %.2f\n",
           wps, r[0], r[1], r[2], r[3], r[4]);
  }
  printf("%s %s\n", cudaGetErrorString(cudaGetLastError()), cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
