// Micro-benchmark (dev tool): inner loop of the tiled correlation consumer with its shared-memory
// operand loads, 8 consumer warps per SM, no pipeline / epilogue -- the ceiling of a formulation.
//   KA  channel-parity FFMA2 (9x9 outputs, 162 accumulator registers)          [round-1 kernel]
//   KB  row-pair FFMA2: thread = N column x 2 P rows; scalar-broadcast N operand times natural
//       (row i, row i+1) P pairs read from a paired plane layout; 162 accumulators
// Prints SM cycles per 8-channel chunk and the FMA-lane utilisation (valid FMAs / 128 / cycles).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t swz32(uint32_t o) { return o ^ (((o >> 7) & 1u) << 4); }
constexpr int Q = 9, NCOL = 64, PCOL = 56, PXB = 32;
extern __shared__ __align__(1024) unsigned char smem[];

// ---- KA: TH = 4, one row per 64 threads
namespace ka {
constexpr int TH = 4, NROW = TH + 8;
constexpr int P_BYTES = TH * PCOL * PXB, N_BYTES = NROW * NCOL * PXB, STAGE = P_BYTES + N_BYTES;
}
__global__ void __maxnreg__(232) kA(float* out, long long* cyc, int nchunks) {
  using namespace ka;
  const int tid = threadIdx.x, ti = tid / NCOL, tc = tid % NCOL;
  for (int e = tid; e < 2 * STAGE / 4; e += 256) reinterpret_cast<float*>(smem)[e] = (float)(e % 13) * 0.01f;
  __syncthreads();
  uint32_t a_off[Q];
#pragma unroll
  for (int k = 0; k < Q; ++k) { int lp = min(max(tc - k, 0), PCOL - 1); a_off[k] = swz32((uint32_t)((ti * PCOL + lp) * PXB)); }
  const uint32_t nb_off = P_BYTES + swz32((uint32_t)((ti * NCOL + tc) * PXB));
  float2 acc2[Q][Q];
#pragma unroll
  for (int m = 0; m < Q; ++m)
#pragma unroll
    for (int k = 0; k < Q; ++k) acc2[m][k] = make_float2(0.f, 0.f);
  long long t0 = clock64();
  for (int c = 0; c < nchunks; ++c) {
    const unsigned char* sb = smem + (c & 1) * STAGE;
#pragma unroll
    for (int qd = 0; qd < 2; ++qd) {
      const unsigned char* nbp = sb + (nb_off ^ (uint32_t)(qd << 4));
      float4 bq[Q];
#pragma unroll
      for (int m = 0; m < Q; ++m) bq[m] = *reinterpret_cast<const float4*>(nbp + m * NCOL * PXB);
      float2 ahi_prev = make_float2(0.f, 0.f);
#pragma unroll
      for (int k = 0; k <= Q; ++k) {
        float2 alo = make_float2(0.f, 0.f), ahi = make_float2(0.f, 0.f);
        if (k < Q) {
          const float4 a = *reinterpret_cast<const float4*>(sb + (a_off[k] ^ (uint32_t)(qd << 4)));
          alo = make_float2(a.x, a.y); ahi = make_float2(a.z, a.w);
#pragma unroll
          for (int m = 0; m < Q; ++m) acc2[m][k] = __ffma2_rn(alo, make_float2(bq[m].x, bq[m].y), acc2[m][k]);
        }
        if (k > 0) {
#pragma unroll
          for (int m = 0; m < Q; ++m) acc2[m][k - 1] = __ffma2_rn(ahi_prev, make_float2(bq[m].z, bq[m].w), acc2[m][k - 1]);
        }
        ahi_prev = ahi;
      }
    }
    __syncwarp();
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int m = 0; m < Q; ++m)
#pragma unroll
    for (int k = 0; k < Q; ++k) s += acc2[m][k].x + 2.f * acc2[m][k].y;
  out[blockIdx.x * 256 + tid] = s;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

// ---- KB: TH = 8, one row pair per 64 threads
namespace kb {
constexpr int TH = 8, NROW = TH + 8;
constexpr int PP_PLANE = (TH / 2) * PCOL * 16 + 256;  // [rowpair][px] x (r0c,r1c,r0c',r1c'), 128 B guards
constexpr int PP_BYTES = 4 * PP_PLANE, N_BYTES = NROW * NCOL * PXB, STAGE = PP_BYTES + N_BYTES;
}
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
template <int HINT>
__device__ __forceinline__ void spin_wait(uint64_t* bar, uint32_t parity) {
  if (HINT)
    asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1, %2;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity), "r"(0x989680) : "memory");
  else
    asm volatile("{\n\t.reg .pred P1;\n\tLAB_WAIT:\n\tmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n\t@P1 bra DONE;\n\tbra LAB_WAIT;\n\tDONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
template <int NLD, int NREG, int SPIN = 0>
__global__ void __maxnreg__(NREG) kB(float* out, long long* cyc, int nchunks) {
  using namespace kb;
  __shared__ uint64_t spinbar;
  if (threadIdx.x == 0) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&spinbar)), "r"(8) : "memory"); }
  __syncthreads();
  if (threadIdx.x >= 256) {   // SPIN variants: 4 extra warps waiting on an mbarrier, as the kernel's producers do
    if (SPIN == 1) spin_wait<0>(&spinbar, 0); else spin_wait<1>(&spinbar, 0);
    return;
  }
  const int tid = threadIdx.x, tp = tid / NCOL, tc = tid % NCOL;
  for (int e = tid; e < 2 * STAGE / 4; e += 256) reinterpret_cast<float*>(smem)[e] = (float)(e % 13) * 0.01f;
  asm volatile("bar.sync 1, 256;" ::: "memory");
  const uint32_t p_off = 128 + (uint32_t)((tp * PCOL + tc) * 16);
  const uint32_t nb_off = PP_BYTES + swz32((uint32_t)((2 * tp * NCOL + tc) * PXB));
  float2 acc2[8][Q]; float accT[Q], accB[Q];
#pragma unroll
  for (int k = 0; k < Q; ++k) { accT[k] = accB[k] = 0.f;
#pragma unroll
    for (int m = 0; m < 8; ++m) acc2[m][k] = make_float2(0.f, 0.f); }
  long long t0 = clock64();
  for (int c = 0; c < nchunks; ++c) {
    const unsigned char* sb = smem + (c & 1) * STAGE;
#define KB_STEP(NX, PX)                                                                    \
    accT[k] = fmaf(n[0].NX, PX.x, accT[k]);                                                \
    _Pragma("unroll") for (int m = 0; m < 8; ++m)                                          \
      acc2[m][k] = __ffma2_rn(make_float2(n[m + 1].NX, n[m + 1].NX), PX, acc2[m][k]);      \
    accB[k] = fmaf(n[9].NX, PX.y, accB[k]);
#pragma unroll 1
    for (int qd = 0; qd < 2; ++qd) {
      if (NLD == 0) {
        float4 n[10];
#pragma unroll
        for (int mm = 0; mm < 10; ++mm) n[mm] = *reinterpret_cast<const float4*>(sb + (nb_off ^ (uint32_t)(qd << 4)) + mm * NCOL * PXB);
#pragma unroll
        for (int k = 0; k < Q; ++k) {
          const float4 pa = *reinterpret_cast<const float4*>(sb + p_off + (2 * qd) * PP_PLANE - k * 16);
          const float4 pb = *reinterpret_cast<const float4*>(sb + p_off + (2 * qd + 1) * PP_PLANE - k * 16);
          const float2 p0 = make_float2(pa.x, pa.y), p1 = make_float2(pa.z, pa.w), p2 = make_float2(pb.x, pb.y), p3 = make_float2(pb.z, pb.w);
          KB_STEP(x, p0) KB_STEP(y, p1) KB_STEP(z, p2) KB_STEP(w, p3)
        }
      } else {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh) {   // 2 channels at a time: N re-read (only half of each 16-byte load is kept)
          float2 n[10];
#pragma unroll
          for (int mm = 0; mm < 10; ++mm) {
            const float4 t = *reinterpret_cast<const float4*>(sb + (nb_off ^ (uint32_t)(qd << 4)) + mm * NCOL * PXB);
            n[mm] = hh ? make_float2(t.z, t.w) : make_float2(t.x, t.y);
          }
#pragma unroll
          for (int k = 0; k < Q; ++k) {
            const float4 pa = *reinterpret_cast<const float4*>(sb + p_off + (2 * qd + hh) * PP_PLANE - k * 16);
            const float2 p0 = make_float2(pa.x, pa.y), p1 = make_float2(pa.z, pa.w);
            KB_STEP(x, p0) KB_STEP(y, p1)
          }
        }
      }
    }
    __syncwarp();
  }
  long long t1 = clock64();
  __syncwarp();
  if ((tid & 31) == 0) asm volatile("{\n\t.reg .b64 st;\n\tmbarrier.arrive.shared::cta.b64 st, [%0];\n\t}" ::"r"(smem_u32(&spinbar)) : "memory");
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < Q; ++k) { s += accT[k] + accB[k];
#pragma unroll
    for (int m = 0; m < 8; ++m) s += acc2[m][k].x * 3.f + acc2[m][k].y; }
  out[blockIdx.x * 256 + tid] = s;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}


// ---- KS: scalar FFMA, 81 accumulators, loop order (k, channel, m): the first-frame operand stays in
//      the register-reuse cache across the 9 m's.  TH rows x 64 threads (TH = 4 or 6).
template <int TH, int ORDER>
__global__ void __launch_bounds__(TH * 64, 1) kS(float* out, long long* cyc, int nchunks) {
  constexpr int NROW = TH + 8, P_BYTES = TH * PCOL * PXB, N_BYTES = NROW * NCOL * PXB, STAGE = P_BYTES + N_BYTES;
  const int tid = threadIdx.x, ti = tid / NCOL, tc = tid % NCOL;
  for (int e = tid; e < 2 * STAGE / 4; e += TH * 64) reinterpret_cast<float*>(smem)[e] = (float)(e % 13) * 0.01f;
  __syncthreads();
  uint32_t a_off[Q];
#pragma unroll
  for (int k = 0; k < Q; ++k) { int lp = min(max(tc - k, 0), PCOL - 1); a_off[k] = swz32((uint32_t)((ti * PCOL + lp) * PXB)); }
  const uint32_t nb_off = P_BYTES + swz32((uint32_t)((ti * NCOL + tc) * PXB));
  float acc[Q][Q];
#pragma unroll
  for (int m = 0; m < Q; ++m)
#pragma unroll
    for (int k = 0; k < Q; ++k) acc[m][k] = 0.f;
  long long t0 = clock64();
  for (int c = 0; c < nchunks; ++c) {
    const unsigned char* sb = smem + (c & 1) * STAGE;
#pragma unroll
    for (int qd = 0; qd < 2; ++qd) {
      const unsigned char* nbp = sb + (nb_off ^ (uint32_t)(qd << 4));
      float4 bv[Q];
#pragma unroll
      for (int m = 0; m < Q; ++m) bv[m] = *reinterpret_cast<const float4*>(nbp + m * NCOL * PXB);
#pragma unroll
      for (int k = 0; k < Q; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(sb + (a_off[k] ^ (uint32_t)(qd << 4)));
        if (ORDER == 0) {
#pragma unroll
          for (int m = 0; m < Q; ++m) {
            acc[m][k] = fmaf(a.x, bv[m].x, acc[m][k]); acc[m][k] = fmaf(a.y, bv[m].y, acc[m][k]);
            acc[m][k] = fmaf(a.z, bv[m].z, acc[m][k]); acc[m][k] = fmaf(a.w, bv[m].w, acc[m][k]);
          }
        } else {
#pragma unroll
          for (int m = 0; m < Q; ++m) acc[m][k] = fmaf(a.x, bv[m].x, acc[m][k]);
#pragma unroll
          for (int m = 0; m < Q; ++m) acc[m][k] = fmaf(a.y, bv[m].y, acc[m][k]);
#pragma unroll
          for (int m = 0; m < Q; ++m) acc[m][k] = fmaf(a.z, bv[m].z, acc[m][k]);
#pragma unroll
          for (int m = 0; m < Q; ++m) acc[m][k] = fmaf(a.w, bv[m].w, acc[m][k]);
        }
      }
    }
    __syncwarp();
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int m = 0; m < Q; ++m)
#pragma unroll
    for (int k = 0; k < Q; ++k) s += acc[m][k] * (1.f + m);
  out[blockIdx.x * (TH * 64) + tid] = s;
  if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}

template <class K> static double run(K kern, int smem_bytes, float* out, long long* cyc, int nchunks, int nthr = 256) {
  long long h[148];
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
  for (int w = 0; w < 2; ++w) { kern<<<148, nthr, smem_bytes>>>(out, cyc, nchunks); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost); }
  return (double)h[0] / nchunks;
}
int main() {
  float* out; long long* cyc;
  const int nchunks = 4000;
  cudaMalloc(&out, sizeof(float) * 148 * 256); cudaMalloc(&cyc, sizeof(long long) * 148);
  const double ca = run(kA, 2 * ka::STAGE, out, cyc, nchunks);
  printf("KA channel-parity      : %5.0f cycles/chunk -> %.3f valid FMA/lane/clk\n", ca, 256.0 * 81 * 8 / 128 / ca);
  const double b0 = run(kB<0, 232>, 2 * kb::STAGE, out, cyc, nchunks);
  printf("KB row-pair N x4ch  232: %5.0f cycles/chunk -> %.3f valid FMA/lane/clk (pipe ideal 0.9)\n", b0, 256.0 * 162 * 8 / 128 / b0);
  const double b1 = run(kB<0, 240>, 2 * kb::STAGE, out, cyc, nchunks);
  printf("KB row-pair N x4ch  240: %5.0f cycles/chunk -> %.3f\n", b1, 256.0 * 162 * 8 / 128 / b1);
  const double s1 = run(kB<0, 168, 1>, 2 * kb::STAGE, out, cyc, nchunks, 384);
  printf("KB 168 regs + 4 warps spinning (try_wait)       : %5.0f cycles/chunk\n", s1);
  const double s2 = run(kB<0, 168, 2>, 2 * kb::STAGE, out, cyc, nchunks, 384);
  printf("KB 168 regs + 4 warps spinning (try_wait + hint): %5.0f cycles/chunk\n", s2);
  const double s0 = run(kB<0, 168, 0>, 2 * kb::STAGE, out, cyc, nchunks, 256);
  printf("KB 168 regs, no extra warps                      : %5.0f cycles/chunk\n", s0);
  const double b2 = run(kB<1, 232>, 2 * kb::STAGE, out, cyc, nchunks);
  printf("KB row-pair N x2ch  232: %5.0f cycles/chunk -> %.3f\n", b2, 256.0 * 162 * 8 / 128 / b2);
  {
    cudaFree(out); cudaMalloc(&out, sizeof(float) * 148 * 512);
    const int st4 = 2 * (4 * PCOL * PXB + 12 * NCOL * PXB), st6 = 2 * (6 * PCOL * PXB + 14 * NCOL * PXB);
    const double a4 = run(kS<4, 0>, st4, out, cyc, nchunks, 256), b4 = run(kS<4, 1>, st4, out, cyc, nchunks, 256);
    const double a6 = run(kS<6, 0>, st6, out, cyc, nchunks, 384), b6 = run(kS<6, 1>, st6, out, cyc, nchunks, 384);
    const double a8 = run(kS<8, 0>, 2 * (8 * PCOL * PXB + 16 * NCOL * PXB), out, cyc, nchunks, 512), b8 = run(kS<8, 1>, 2 * (8 * PCOL * PXB + 16 * NCOL * PXB), out, cyc, nchunks, 512);
    printf("KS scalar TH=4 (8 warps): order m-inner/4ch %5.0f -> %.3f | order (k,c,m) %5.0f -> %.3f valid FMA/lane/clk\n", a4, 256.0 * 81 * 8 / 128 / a4, b4, 256.0 * 81 * 8 / 128 / b4);
    printf("KS scalar TH=6 (12 warps): %5.0f -> %.3f | %5.0f -> %.3f\n", a6, 384.0 * 81 * 8 / 128 / a6, b6, 384.0 * 81 * 8 / 128 / b6);
    printf("KS scalar TH=8 (16 warps): %5.0f -> %.3f | %5.0f -> %.3f\n", a8, 512.0 * 81 * 8 / 128 / a8, b8, 512.0 * 81 * 8 / 128 / b8);
  }
  printf("%s %s\n", cudaGetErrorString(cudaGetLastError()), cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
