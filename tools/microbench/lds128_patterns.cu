// Shared-memory cost of LDS.128 for the operand-address patterns of the 16-lane (Ant-class) tile product, two envs per
// warp: how many wavefronts (LSU cycles per instruction) each lane -> (row group, column group) mapping needs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/lds128 tools/microbench/lds128_patterns.cu && /tmp/lds128
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ int addr_of(int pattern, int l) {
  // floats; env slabs 1776 floats apart (= 16 banks mod 32), matrix row stride 20 floats, tile 4 x 4
  const int env = l >> 4, j = l & 15, slab = env * 1776;
  const int q = (j >> 3), k = j & 7;
  switch (pattern) {
    case 0: return 0;                                              // one address: broadcast
    case 1: return slab + (j >> 2) * 4 * 20;                       // A operand now: row group = j / 4 (rows 4 rg .. ), first of its 4 rows
    case 2: return slab + (j & 3) * 4;                             // B operand now: column group = j % 4
    case 3: return slab + (k >> 1) * 4 * 20;                       // A operand, interleaved map: rg = (j % 8) / 2
    case 4: return slab + ((((k >> 1) + 2 * q + (k & 1)) & 3)) * 4;  // B operand, interleaved map: cg = (rg + 2 q + k % 2) % 4
    case 5: return l * 4;                                          // 32 distinct consecutive 16-byte chunks (512 B)
    case 6: return slab + (j >> 1 & 3) * 4 * 20;                   // A operand, map rg = (j / 2) % 4 (pairs), quarters equal
    case 7: return (l >> 2) * 3 * 28;                              // Humanoid A operand: warp per env, row group = l / 4 (3 rows each, stride 28)
    case 8: return (l & 3) * 6;                                    // Humanoid B operand: column group = l % 4 (6 columns each)
    case 9: return l * 2;                                          // 32 distinct consecutive 8-byte chunks
    // which duplicates does the hardware merge?  8 distinct, bank-disjoint 16-byte chunks per warp, chunk index c(l):
    case 10: return (l >> 2) * 4;                                  // adjacent quads share
    case 11: return (l & 7) * 4;                                   // lanes l, l + 8, l + 16, l + 24 share
    case 12: return ((l >> 1) & 7) * 4;                            // adjacent pairs share, the pattern repeats every 16 lanes
    case 13: return ((l & 3) + 4 * (l >> 4)) * 4;                  // per half-warp: lanes j, j + 4, j + 8, j + 12 share (the B operand, bank-disjoint)
    case 14: return ((l >> 2 & 3) + 4 * (l >> 4)) * 4;             // per half-warp: adjacent quads share (the A operand with interleaved rows)
    case 15: return ((l >> 1 & 3) + 4 * (l >> 4)) * 4;             // per half-warp: adjacent pairs share, repeats every 8 lanes
    case 16: return (l >> 1) * 4;                                  // 16 distinct chunks, adjacent pairs share
    case 17: return ((l >> 1) & 3) * 6;                            // LDS.64: 4 column groups (6 floats apart), adjacent pairs share
    case 18: return (l >> 1) * 2;                                  // LDS.64: 16 distinct 8-byte chunks, adjacent pairs share
    case 19: return (l >> 2) * 2;                                  // LDS.64: 8 distinct, adjacent quads share
    case 20: return ((l >> 1) & 3) * 6 + 2;                        // LDS.64 second pair of the 6 columns
    default: return 0;
  }
}

template <int P, int W>
__global__ void k(float* out, long long* cyc, int iters) {
  extern __shared__ float4 sm4[];
  float* sm = reinterpret_cast<float*>(sm4);
  for (int i = threadIdx.x; i < 2 * 1776 + 1024; i += blockDim.x) sm[i] = 1.0f;
  __syncthreads();
  const int l = threadIdx.x & 31;
  int a = addr_of(P, l);
  float4 acc = make_float4(0, 0, 0, 0);
  const unsigned sa = (unsigned)__cvta_generic_to_shared(sm + a);
  long long t0 = clock64();
  for (int i = 0; i < iters; i += 8) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {              // eight independent loads in flight per warp: the LSU is the limit
      float4 v = make_float4(0, 0, 0, 0);
      if (W == 4) asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(sa) : "memory");
      else if (W == 2) asm volatile("ld.volatile.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(sa) : "memory");
      else asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v.x) : "r"(sa) : "memory");
      acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w;
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc.x + acc.y + acc.z + acc.w;
  if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int P, int W = 4>
void run(const char* name) {
  float* out; long long* cyc; long long h;
  cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
  const int iters = 20000, threads = 512;
  size_t smem = (2 * 1776 + 1024) * 4;
  k<P, W><<<148, threads, smem>>>(out, cyc, iters);
  k<P, W><<<148, threads, smem>>>(out, cyc, iters);
  cudaDeviceSynchronize();
  cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
  // 16 warps issue `iters` loads each: LSU cycles per load instruction = elapsed / (iters * 16)
  printf("%-66s %6.2f LSU cycles per load instruction (%s)\n", name, (double)h / iters / 16.0, cudaGetErrorString(cudaGetLastError()));
  cudaFree(out); cudaFree(cyc);
}

int main() {
  run<0>("LDS.128 one address (broadcast)");
  run<5>("LDS.128 32 distinct 16-byte chunks");
  run<1>("A operand, map rg = j / 4 (now)");
  run<2>("B operand, map cg = j % 4 (now)");
  run<3>("A operand, interleaved map rg = (j % 8) / 2");
  run<4>("B operand, interleaved map cg = (rg + 2 q + j % 2) % 4");
  run<6>("A operand, map rg = (j / 2) % 4");
  run<7>("LDS.128 Humanoid A operand (8 distinct rows per warp)");
  run<10>("LDS.128 8 chunks: adjacent quads share");
  run<11>("LDS.128 8 chunks: lanes l, l+8, l+16, l+24 share");
  run<12>("LDS.128 8 chunks: adjacent pairs share, repeats every 16 lanes");
  run<13>("LDS.128 4 chunks per half-warp: lanes j, j+4, j+8, j+12 share");
  run<14>("LDS.128 4 chunks per half-warp: adjacent quads share");
  run<15>("LDS.128 4 chunks per half-warp: adjacent pairs, repeats every 8");
  run<16>("LDS.128 16 chunks: adjacent pairs share");
  run<8, 2>("LDS.64  Humanoid B operand (4 distinct column groups per warp)");
  run<17, 2>("LDS.64  4 column groups, adjacent pairs share");
  run<20, 2>("LDS.64  4 column groups (+8 B), adjacent pairs share");
  run<18, 2>("LDS.64  16 distinct chunks, adjacent pairs share");
  run<19, 2>("LDS.64  8 distinct chunks, adjacent quads share");
  run<0, 2>("LDS.64  one address");
  run<9, 2>("LDS.64  32 distinct 8-byte chunks");
  run<0, 1>("LDS.32  one address");
  run<5, 1>("LDS.32  32 distinct words (stride 4)");
  run<8, 1>("LDS.32  4 distinct addresses");
  return 0;
}
