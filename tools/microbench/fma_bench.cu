// Microbenchmarks that steer the Newton-Schulz tile kernel (run under gpurun):
//   1. FFMA (3 register operands) vs packed fma.rn.f32x2 issue throughput per SM
//   2. shared-memory wavefront cost of broadcast LDS.128 / LDS.64 / LDS.32
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_bench fma_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void k_ffma(float* out, int iters, float a, float b) {
  float acc[ILP];
#pragma unroll
  for (int i = 0; i < ILP; ++i) acc[i] = threadIdx.x + i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) acc[i] = fmaf(acc[i], a, b);
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int ILP>
__global__ void k_ffma2(float* out, int iters, float a, float b) {
  unsigned long long acc[ILP], av, bv;
  asm("mov.b64 %0, {%1, %1};" : "=l"(av) : "f"(a));
  asm("mov.b64 %0, {%1, %1};" : "=l"(bv) : "f"(b));
#pragma unroll
  for (int i = 0; i < ILP; ++i) { float x = threadIdx.x + i; asm("mov.b64 %0, {%1, %1};" : "=l"(acc[i]) : "f"(x)); }
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(acc[i]) : "l"(av), "l"(bv));
  }
  float s = 0;
#pragma unroll
  for (int i = 0; i < ILP; ++i) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i])); s += lo + hi; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mode 0: LDS.32 broadcast, 1: LDS.64 broadcast, 2: LDS.128 broadcast, 3: LDS.128 own-row (stride 28 floats),
// 4: LDS.64 4 distinct addresses (tile B pattern), 5: LDS.128 8 distinct addresses (tile A pattern)
template <int MODE>
__global__ void k_lds(float* out, int iters) {
  __shared__ __align__(16) float sm[32 * 28 + 64];
  for (int i = threadIdx.x; i < 32 * 28 + 64; i += blockDim.x) sm[i] = i;
  __syncthreads();
  const int lane = threadIdx.x & 31;
  float s = 0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int u = 0; u < 16; ++u) {
      int off = (it + u) & 7;
      unsigned addr;
      float x, y, z, w;
      if (MODE == 0) { addr = (unsigned)__cvta_generic_to_shared(&sm[off * 4]); asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(addr)); s += x; }
      if (MODE == 1) { addr = (unsigned)__cvta_generic_to_shared(&sm[off * 4]); asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr)); s += x + y; }
      if (MODE == 2) { addr = (unsigned)__cvta_generic_to_shared(&sm[off * 4]); asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(addr)); s += x + y + z + w; }
      if (MODE == 3) { addr = (unsigned)__cvta_generic_to_shared(&sm[lane * 28 + (off & 3) * 4]); asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(addr)); s += x + y + z + w; }
      if (MODE == 4) { addr = (unsigned)__cvta_generic_to_shared(&sm[off * 28 + (lane & 3) * 6]); asm volatile("ld.shared.v2.f32 {%0,%1}, [%2];" : "=f"(x), "=f"(y) : "r"(addr)); s += x + y; }
      if (MODE == 5) { addr = (unsigned)__cvta_generic_to_shared(&sm[(lane >> 2) * 3 * 28 + (off & 3) * 4]); asm volatile("ld.shared.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(x), "=f"(y), "=f"(z), "=f"(w) : "r"(addr)); s += x + y + z + w; }
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_ms(F f) {
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(a); f(); cudaEventRecord(b); cudaEventSynchronize(b);
  float ms; cudaEventElapsedTime(&ms, a, b); return ms;
}

int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  int sms = p.multiProcessorCount; double ghz = p.clockRate * 1e-6;
  float* out; cudaMalloc(&out, sizeof(float) * sms * 1024 * 8);
  const int iters = 4096;
  printf("SMs %d clock %.3f GHz (nominal)\n", sms, ghz);
  for (int warps = 4; warps <= 16; warps *= 2) {
    int threads = warps * 32;
    float t1 = time_ms([&] { k_ffma<16><<<sms, threads>>>(out, iters, 1.0001f, 0.5f); });
    float t2 = time_ms([&] { k_ffma2<16><<<sms, threads>>>(out, iters, 1.0001f, 0.5f); });
    double inst = (double)iters * 16 * warps;  // warp instructions per SM
    printf("warps/SM %2d: FFMA  %.3f ms -> %.2f warp-inst/clk/SM (%.1f FMA/clk/SM)\n", warps, t1, inst / (t1 * 1e-3 * ghz * 1e9), 32 * inst / (t1 * 1e-3 * ghz * 1e9));
    printf("warps/SM %2d: FFMA2 %.3f ms -> %.2f warp-inst/clk/SM (%.1f FMA/clk/SM)\n", warps, t2, inst / (t2 * 1e-3 * ghz * 1e9), 64 * inst / (t2 * 1e-3 * ghz * 1e9));
  }
  const char* names[] = {"LDS.32 bcast", "LDS.64 bcast", "LDS.128 bcast", "LDS.128 own-row s28", "LDS.64 4-addr (tile B)", "LDS.128 8-addr (tile A)"};
  for (int mode = 0; mode < 6; ++mode) {
    int threads = 256; float t = 0;
    auto run = [&](auto kern) { t = time_ms([&] { kern<<<sms, threads>>>(out, 2048); }); };
    switch (mode) { case 0: run(k_lds<0>); break; case 1: run(k_lds<1>); break; case 2: run(k_lds<2>); break;
                    case 3: run(k_lds<3>); break; case 4: run(k_lds<4>); break; default: run(k_lds<5>); }
    double inst = 2048.0 * 16 * (threads / 32);
    printf("%-26s %.3f ms -> %.2f clk per warp-LDS per SM\n", names[mode], t, (t * 1e-3 * ghz * 1e9) / inst);
  }
  return 0;
}
