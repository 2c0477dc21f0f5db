// tf32x3_tile.cu -- micro-benchmark that settles the tensor-core question for the cost volume
// (VERDICT r01 "next" #4): an 8x16-pixel first-frame tile against its 16x24-pixel second-frame halo is
// a dense [128 x C] . [C x 384] GEMM of which 81/384 = 21 % is the wanted band.  With fp32 operands
// split into tf32 hi + lo and three tcgen05.mma kind::tf32 passes (hi.hi + lo.hi + hi.lo, fp32
// accumulation in TMEM) the tile costs 3 x (C/8) x 2 MMAs of 128 x 192 x 8.
// Measures (1) what the tensor core does with the low 13 mantissa bits of an fp32 word (truncate?),
// (2) the error of 1x / 3x TF32 against fp64 relative to mean|a.b| (the 1e-5 condition-aware bound),
// (3) the sustained MMA rate of the tile in SMSP clocks.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tf32x3_tile tf32x3_tile.cu && ./tf32x3_tile
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <vector>

#define M 128
#define N 384
#define NH 192
#define KC 8           // channels per K chunk = one MMA K step (8 tf32 = 32 bytes per pixel row)
#define NCHUNK 4       // K = 32 resident
#define A_BYTES (M * 32)
#define B_BYTES (N * 32)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t swz32(uint32_t off) { return off ^ (((off >> 7) & 1u) << 4); }

// K-major, SWIZZLE_32B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows (pixels)
// 32 bytes apart, 8-row core groups 256 bytes apart
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address
  d |= (uint64_t)1 << 16;                           // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(256 >> 4) << 32;                  // stride byte offset: 8 rows x 32 B
  d |= (uint64_t)1 << 46;                           // descriptor version (Blackwell)
  d |= (uint64_t)6 << 61;                           // SWIZZLE_32B
  return d;
}
// kind::tf32, fp32 accumulate, A and B K-major, M x Nn
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t da, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(da), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
// A operand from tensor memory (lane = row, one 32-bit column per K element), B from shared memory
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t db, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(db), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ float rna_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// mode 0: one pass, raw fp32 words      mode 1: one pass, low 13 bits masked off by the kernel
// mode 2: three passes, hi = raw word (relies on the hardware ignoring the low bits), lo = rna(x - trunc(x))
// mode 3: three passes, hi = rna(x) written explicitly, lo = rna(x - hi)
// mode 4: as mode 2 with the A operand (hi and lo) held in TMEM columns 384.. (tcgen05.st), B in shared memory
__global__ void __launch_bounds__(128, 1)
tf32x3_kernel(const float* __restrict__ A, const float* __restrict__ B, float* __restrict__ D, int mode, int reps,
              long long* cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + NCHUNK * A_BYTES;
  unsigned char* b_hi = a_lo + NCHUNK * A_BYTES;
  unsigned char* b_lo = b_hi + NCHUNK * B_BYTES;
  __shared__ uint32_t tmem_base_s;
  __shared__ __align__(8) uint64_t bar;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

  // operands -> shared memory in the canonical K-major SWIZZLE_32B layout, one buffer per K chunk
  for (int e = tid; e < (M + N) * NCHUNK * KC; e += blockDim.x) {
    const bool isb = e >= M * NCHUNK * KC;
    const int ee = isb ? e - M * NCHUNK * KC : e;
    const int row = ee / (NCHUNK * KC), k = ee % (NCHUNK * KC), chunk = k / KC, kk = k % KC;
    const float x = isb ? B[row * NCHUNK * KC + k] : A[row * NCHUNK * KC + k];
    float hi, lo;
    if (mode == 0) { hi = x; lo = 0.f; }
    else if (mode == 1) { hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); lo = 0.f; }
    else if (mode == 2 || mode == 4) { hi = x; lo = rna_tf32(x - __uint_as_float(__float_as_uint(x) & 0xFFFFE000u)); }
    else { hi = rna_tf32(x); lo = rna_tf32(x - hi); }
    const uint32_t off = swz32((uint32_t)(row * 32 + kk * 4));
    unsigned char* ph = (isb ? b_hi + chunk * B_BYTES : a_hi + chunk * A_BYTES) + off;
    unsigned char* pl = (isb ? b_lo + chunk * B_BYTES : a_lo + chunk * A_BYTES) + off;
    *reinterpret_cast<float*>(ph) = hi;
    *reinterpret_cast<float*>(pl) = lo;
  }
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(smem_u32(&bar)) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base_s)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores -> tensor-core (async proxy) reads
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem = tmem_base_s;
  if (mode == 4) {
    // thread = TMEM lane = row of A: chunk c -> columns 384 + 16c (hi) and 384 + 16c + 8 (lo)
    const int row = warp * 32 + lane;
    for (int c = 0; c < NCHUNK; ++c)
      for (int part = 0; part < 2; ++part) {
        uint32_t v[8];
        const unsigned char* src = (part ? a_lo : a_hi) + c * A_BYTES;
        for (int k = 0; k < 8; ++k) v[k] = *reinterpret_cast<const uint32_t*>(src + swz32((uint32_t)(row * 32 + k * 4)));
        const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)(384 + c * 16 + part * 8);
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                     ::"r"(taddr), "r"(v[0]), "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7]) : "memory");
      }
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  }

  long long t0 = 0, t1 = 0;
  if (tid == 0) {
    const uint32_t idesc = make_idesc(M, NH);
    const int npass = mode >= 2 ? 3 : 1;
    // descriptors are built before the timed loop: the issuing thread's own instruction stream must not
    // be what is measured (with the address arithmetic inside the loop the rate drops to ~130 clk/MMA)
    uint64_t da[NCHUNK][3], db[NCHUNK][2][3];
    uint32_t ta[NCHUNK][3];
    for (int c = 0; c < NCHUNK; ++c)
      for (int p = 0; p < 3; ++p) {
        da[c][p] = make_desc(smem_u32((p == 1 ? a_lo : a_hi) + c * A_BYTES));
        ta[c][p] = tmem + 384 + c * 16 + (p == 1 ? 8 : 0);
        for (int h = 0; h < 2; ++h) db[c][h][p] = make_desc(smem_u32((p == 2 ? b_lo : b_hi) + c * B_BYTES + h * NH * 32));
      }
    t0 = clock64();
    if (mode == 4) {
      for (int r = 0; r < reps; ++r)
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
          for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int p = 0; p < 3; ++p) mma_tf32_ts(tmem + h * NH, ta[c][p], db[c][h][p], idesc, (r | c | p) != 0);
    } else {
      for (int r = 0; r < reps; ++r)
#pragma unroll
        for (int c = 0; c < NCHUNK; ++c)
#pragma unroll
          for (int h = 0; h < 2; ++h)
            for (int p = 0; p < npass; ++p) mma_tf32(tmem + h * NH, da[c][p], db[c][h][p], idesc, (r | c | p) != 0);
    }
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
  }
  // everybody waits for the MMAs
  asm volatile(
      "{\n\t.reg .pred P1;\n\tW: mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], 0;\n\t@P1 bra DN;\n\tbra W;\n\tDN:\n\t}"
      ::"r"(smem_u32(&bar)) : "memory");
  if (tid == 0) { t1 = clock64(); cycles[0] = t1 - t0; }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");

  // epilogue: warp w reads TMEM lanes 32w..32w+31 (row = lane), 16 columns at a time
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < N; c0 += 16) {
    uint32_t v[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
          "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int j = 0; j < 16; ++j) D[(size_t)row * N + c0 + j] = __uint_as_float(v[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem) : "memory");
}

int main() {
  const int K = NCHUNK * KC;
  std::vector<float> A(M * K), B(N * K), D(M * N);
  srand(1);
  auto rnd = []() {  // ~N(0,1)
    float s = 0.f;
    for (int i = 0; i < 12; ++i) s += (float)rand() / RAND_MAX;
    return s - 6.f;
  };
  for (auto& x : A) x = rnd();
  for (auto& x : B) x = rnd();
  std::vector<double> ref(M * N), cond(M * N);
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      double s = 0, c = 0;
      for (int k = 0; k < K; ++k) { s += (double)A[m * K + k] * B[n * K + k]; c += fabs((double)A[m * K + k] * B[n * K + k]); }
      ref[m * N + n] = s / K; cond[m * N + n] = c / K;
    }
  // the same sums in plain fp32 FFMA order (what the FFMA kernel computes)
  double ffma_worst = 0;
  for (int m = 0; m < M; ++m)
    for (int n = 0; n < N; ++n) {
      float s = 0.f;
      for (int k = 0; k < K; ++k) s = fmaf(A[m * K + k], B[n * K + k], s);
      const double e = fabs((double)(s / K) - ref[m * N + n]) / cond[m * N + n];
      if (e > ffma_worst) ffma_worst = e;
    }
  float *dA, *dB, *dD; long long* dC;
  cudaMalloc(&dA, A.size() * 4); cudaMalloc(&dB, B.size() * 4); cudaMalloc(&dD, D.size() * 4); cudaMalloc(&dC, 8);
  cudaMemcpy(dA, A.data(), A.size() * 4, cudaMemcpyHostToDevice);
  cudaMemcpy(dB, B.data(), B.size() * 4, cudaMemcpyHostToDevice);
  const int smem = NCHUNK * (2 * A_BYTES + 2 * B_BYTES);
  cudaFuncSetAttribute(tf32x3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  std::vector<float> D0;
  printf("fp32 FFMA (sequential fmaf)        : worst |err| / mean|a.b| = %.3e\n", ffma_worst);
  for (int mode = 0; mode < 5; ++mode) {
    cudaMemset(dD, 0, D.size() * 4);
    tf32x3_kernel<<<1, 128, smem>>>(dA, dB, dD, mode, 1, dC);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("mode %d: %s\n", mode, cudaGetErrorString(e)); return 1; }
    cudaMemcpy(D.data(), dD, D.size() * 4, cudaMemcpyDeviceToHost);
    double worst = 0, rms = 0;
    for (int i = 0; i < M * N; ++i) {
      const double err = fabs((double)D[i] / K - ref[i]) / cond[i];
      worst = fmax(worst, err); rms += err * err;
    }
    const char* names[5] = {"1xTF32 raw fp32 words             ", "1xTF32 low 13 bits masked         ",
                            "3xTF32 hi=raw word, lo=rna(x-trunc)", "3xTF32 hi=rna(x), lo=rna(x-hi)    ", "3xTF32 as [2], A operand in TMEM  "};
    printf("%s: worst |err| / mean|a.b| = %.3e  rms %.3e", names[mode], worst, sqrt(rms / (M * N)));
    if (mode == 0) D0 = D;
    if (mode == 1) {
      int same = 1;
      for (int i = 0; i < M * N; ++i) same &= (D[i] == D0[i]);
      printf("   [raw == masked bit for bit: %s => the tensor core %s the low 13 bits]", same ? "yes" : "NO",
             same ? "ignores (truncates)" : "does NOT simply ignore");
    }
    printf("\n");
  }
  // rate: many repetitions of the tile's MMA sequence (K = 32)
  for (int mode = 0; mode <= 4; mode += 2) {
    const int reps = 400;
    tf32x3_kernel<<<1, 128, smem>>>(dA, dB, dD, mode, reps, dC);
    cudaDeviceSynchronize();
    long long cyc; cudaMemcpy(&cyc, dC, 8, cudaMemcpyDeviceToHost);
    const int npass = mode >= 2 ? 3 : 1;
    const double nmma = (double)reps * NCHUNK * 2 * npass;
    printf("mode %d, %d-pass tile: %.1f clk per 128x192x8 MMA (floor 96), tile (C=32) = %.0f clk, %.1f dense TF32 TFLOP/s/SM-equivalent x148 = %.0f TFLOP/s at 1.965 GHz\n",
           mode, npass, cyc / nmma, cyc / (double)reps, 2.0 * 128 * 192 * 8 / (cyc / nmma) * 1.965e9 / 1e12,
           2.0 * 128 * 192 * 8 / (cyc / nmma) * 1.965e9 / 1e12 * 148);
  }
  printf("band utilisation 81/384 = %.3f; fp32-equivalent useful rate of the 3-pass tile = dense rate x 0.211 / 3\n", 81.0 / 384);
  return 0;
}
