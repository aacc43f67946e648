// Standalone ceiling of the Newton-Schulz tile product (bxg_core.cuh tile_matmul, the code the step
// kernel runs): one CTA per SM, `warps` warps each multiplying its own pair of 24x24 matrices in
// shared memory `reps` times (store + __syncwarp between products, as in the kernel).  Reports the
// achieved fraction of the FP32 FMA peak, i.e. what this tiling can reach when nothing else runs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I brax_b200/csrc -o tools/microbench/ns_tile_bench tools/microbench/ns_tile_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#include "bxg_core.cuh"

using namespace bxg;

template <int G, int W, int MAXT>
__global__ void __launch_bounds__(MAXT) k_tiles(float* out, int reps, int ld) {
  extern __shared__ __align__(16) float sm[];
  using T = Tile<G, W>;
  const int groups = blockDim.x / G, group = threadIdx.x / G, lane = (threadIdx.x & 31) % G;
  const unsigned mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << (((threadIdx.x & 31) / G) * G));
  float* A = sm + group * 3 * W * ld; float* B = A + W * ld; float* C = B + W * ld;
  for (int i = lane; i < W * ld; i += G) { A[i] = 1e-3f * (i % 7); B[i] = 1e-3f * (i % 5); C[i] = 0.f; }
  __syncwarp(mask);
  float acc[T::TM * T::TN];
  for (int r = 0; r < reps; ++r) {
    tile_matmul<T, W>(lane, A, B, ld, acc);
    __syncwarp(mask);
    store_tile<T>(lane, C, ld, acc);
    __syncwarp(mask);
    float* t = B; B = C; C = t;
  }
  if (lane == 0) out[blockIdx.x * groups + group] = B[1];
}

// Variants of the 3x6 loop that isolate one resource each (MODE 1: loads only, accumulate a checksum with one
// FADD per load; MODE 2: FFMA2 only, operands loaded once; MODE 3: scalar FFMA instead of FFMA2, with loads)
template <int MODE>
__global__ void __launch_bounds__(608) k_probe(float* out, int reps, int ld) {
  extern __shared__ __align__(16) float sm[];
  constexpr int W = 24, TM = 3, TN = 6;
  const int group = threadIdx.x / 32, lane = threadIdx.x & 31;
  float* A = sm + group * 3 * W * ld; float* B = A + W * ld;
  for (int i = lane; i < 2 * W * ld; i += 32) A[i] = 1e-3f * (i % 7);
  __syncwarp();
  const int rg = lane / 4, cg = lane % 4;
  const float* a0 = A + rg * TM * ld;
  float2 acc2[TM][TN / 2]; float accs[TM][TN]; float chk = 0.f;
  for (int r = 0; r < TM; ++r) for (int c = 0; c < TN / 2; ++c) { acc2[r][c] = make_float2(0.f, 0.f); accs[r][2 * c] = 0.f; accs[r][2 * c + 1] = 0.f; }
  F4 a[TM]; float bv[TN];
  for (int r = 0; r < TM; ++r) a[r] = ldv4(a0 + r * ld);
  load_cols<TN>(B + cg * TN, bv);
  for (int rep = 0; rep < reps; ++rep) {
#pragma unroll 2
    for (int k0 = 0; k0 < W; k0 += 4) {
      if (MODE != 2 && MODE != 5) {
#pragma unroll
        for (int r = 0; r < TM; ++r) a[r] = ldv4(a0 + r * ld + k0);
      }
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) {
        if (MODE != 2 && MODE != 5) load_cols<TN>(B + (k0 + kk) * ld + cg * TN, bv);
        if (MODE == 1) {
          chk += bv[0] + bv[2] + bv[4] + (kk == 0 ? a[0].x + a[1].x + a[2].x : 0.f);
        } else {
#pragma unroll
          for (int r = 0; r < TM; ++r) {
            const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
            if (MODE == 4 || MODE == 5) {
              // handled below (column-outer order)
            } else if (MODE == 3) {
#pragma unroll
              for (int c = 0; c < TN; ++c) accs[r][c] = fmaf(av, bv[c], accs[r][c]);
            } else {
              const float2 av2 = make_float2(av, av);
#pragma unroll
              for (int c = 0; c < TN / 2; ++c) acc2[r][c] = __ffma2_rn(av2, make_float2(bv[2 * c], bv[2 * c + 1]), acc2[r][c]);
            }
          }
          if (MODE == 4 || MODE == 5) {   // B pair outer, rows inner: consecutive FFMA2 share the 64-bit operand
#pragma unroll
            for (int c = 0; c < TN / 2; ++c) {
              const float2 b2 = make_float2(bv[2 * c], bv[2 * c + 1]);
#pragma unroll
              for (int r = 0; r < TM; ++r) {
                const float av = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
                acc2[r][c] = __ffma2_rn(make_float2(av, av), b2, acc2[r][c]);
              }
            }
          }
        }
      }
    }
  }
  for (int r = 0; r < TM; ++r) for (int c = 0; c < TN / 2; ++c) chk += acc2[r][c].x + acc2[r][c].y + accs[r][2 * c] + accs[r][2 * c + 1];
  if (lane == 0) out[blockIdx.x * (blockDim.x / 32) + group] = chk;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  const int sms = p.multiProcessorCount; const double ghz = p.clockRate * 1e-6;
  float* out; cudaMalloc(&out, sizeof(float) * sms * 64);
  const int reps = 4000;
  auto run = [&](auto kern, int G, int W, int ld, int groups, const char* name) {
    size_t smem = sizeof(float) * groups * 3 * W * ld;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kern<<<sms, groups * G, smem>>>(out, 100, ld); cudaDeviceSynchronize();
    cudaEventRecord(a); kern<<<sms, groups * G, smem>>>(out, reps, ld); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double fma = (double)reps * groups * W * W * W;          // per SM
    double per_clk = fma / (ms * 1e-3 * ghz * 1e9);
    printf("%-34s groups/SM %2d  %.3f ms  %.1f FMA/clk/SM = %.1f%% of the FP32 peak (128)  err=%s\n", name, groups, ms, per_clk,
           100.0 * per_clk / 128.0, cudaGetErrorString(cudaGetLastError()));
  };
  for (int g : {8, 14, 16, 19}) run(k_tiles<32, 24, 608>, 32, 24, 24, g, "G32 W24 3x6 tiles (Humanoid)");
  for (int g : {12, 20, 26}) run(k_tiles<16, 16, 416>, 16, 16, 20, g, "G16 W16 4x4 tiles (Ant, pipelined)");
  for (int g : {14, 18, 20}) run(k_tiles<16, 24, 320>, 16, 24, 24, g, "G16 W24 6x6 tiles, stride 24");
  for (int g : {14, 18, 20}) run(k_tiles<16, 24, 320>, 16, 24, 28, g, "G16 W24 6x6 tiles, stride 28");
  auto probe = [&](auto kern, int groups, const char* name) {
    size_t smem = sizeof(float) * groups * 3 * 24 * 24;
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    kern<<<sms, groups * 32, smem>>>(out, 100, 24); cudaDeviceSynchronize();
    cudaEventRecord(a); kern<<<sms, groups * 32, smem>>>(out, reps, 24); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double cyc = ms * 1e-3 * ghz * 1e9 / reps;     // cycles per 24^3 product round of `groups` warps on one SM
    printf("%-44s warps/SM %2d  %.0f cycles per product round (FMA bound %d, 9-wavefront/k bound %d)  err=%s\n", name, groups, cyc,
           groups * 432 / 4, groups * 216, cudaGetErrorString(cudaGetLastError()));
  };
  for (int g : {8, 19}) {
    probe(k_probe<0>, g, "3x6 FFMA2 + loads (no store/sync)");
    probe(k_probe<1>, g, "3x6 loads only");
    probe(k_probe<2>, g, "3x6 FFMA2 only (operands in registers)");
    probe(k_probe<3>, g, "3x6 scalar FFMA + loads");
    probe(k_probe<4>, g, "3x6 FFMA2 + loads, B-pair-outer order");
    probe(k_probe<5>, g, "3x6 FFMA2 only, B-pair-outer order");
  }
  return 0;
}
