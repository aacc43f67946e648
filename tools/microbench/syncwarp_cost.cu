// Cost of a group barrier for lane groups narrower than a warp (the Ant-class kernels run two 16-lane envs per warp).
// `__syncwarp(mask)` with a run-time half-warp mask is lowered by ptxas to MATCH.ANY + REDUX.OR + VOTEU.ANY + a divergent-path
// WARPSYNC; this measures that sequence against WARPSYNC.ALL and against alternatives, in cycles per barrier,
// with 1 / 4 warps per scheduler running the same loop.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/syncwarp_cost tools/microbench/syncwarp_cost.cu && /tmp/syncwarp_cost
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
  __shared__ float sm[1024];
  const int wl = threadIdx.x & 31, half = wl >> 4;
  const unsigned mask = half ? 0xffff0000u : 0x0000ffffu;
  float* s = sm + (threadIdx.x >> 5) * 32;
  float v = threadIdx.x;
  s[wl] = v;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    s[wl] = v;
    if (MODE == 0) __syncwarp(mask);                       // what DevExec<16>::lanes does
    else if (MODE == 1) __syncwarp();                      // full warp
    else if (MODE == 2) asm volatile("" ::: "memory");     // compiler fence only
    else if (MODE == 3) { if (half) __syncwarp(0xffff0000u); else __syncwarp(0x0000ffffu); }   // constant masks behind a branch
    else if (MODE == 4) asm volatile("bar.sync %0, 32;" :: "r"(1 + (int)(threadIdx.x >> 5)) : "memory");   // a named barrier per warp
    v = s[(wl & 16) | ((wl + 1) & 15)] + 1.0f;             // read the neighbour's value within the half
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = v;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int MODE>
void run(const char* name, int threads) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 20000;
  k<MODE><<<148, threads>>>(out, cyc, iters);
  k<MODE><<<148, threads>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-44s %4d threads/CTA: %7.1f cycles per iteration (%s)\n", name, threads, (double)h / iters, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  for (int threads : {128, 512}) {
    run<2>("no barrier (compiler fence)", threads);
    run<1>("__syncwarp() full mask", threads);
    run<0>("__syncwarp(run-time half-warp mask)", threads);
    run<3>("constant half masks behind a branch", threads);
    if (threads <= 480) run<4>("named barrier per warp (bar.sync id, 32)", threads);
  }
  return 0;
}
