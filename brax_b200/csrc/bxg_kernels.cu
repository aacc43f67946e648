// bxg_kernels.cu -- sm_100a kernels + the C ABI declared in include/bxg.h.
//
// One lane-group (G = 16 or 32 lanes) per environment, whole working state in
// shared memory, all n_frames substeps inside one launch; HBM is touched only
// at kernel entry (load_env) and exit (store_env).  See bxg_core.cuh for the
// algorithm and DESIGN.md for the layout and roofline accounting.
#include <cuda_runtime.h>
#include <stdio.h>

#include <atomic>
#include <mutex>
#include <string>

#include "bxg_core.cuh"

namespace bxg {

// ------------------------------------------------------- device executor
template <int G_>
struct DevExec {
  static constexpr int G = G_;
  int lane;
  unsigned mask;
  struct LaneF {
    float v;
    __device__ __forceinline__ float& operator()(int) { return v; }
  };
  __device__ __forceinline__ void sync() { __syncwarp(mask); }
  template <class F>
  __device__ __forceinline__ void lanes(F&& f) {
    __syncwarp(mask);
    f(lane);
    __syncwarp(mask);
  }
  __device__ __forceinline__ float sum(LaneF& p) {
    float v = p.v;
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(mask, v, o, G);
    return v;
  }
  __device__ __forceinline__ float max(LaneF& p) {
    float v = p.v;
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, o, G));
    return v;
  }
};

constexpr int kThreads = 128;

template <int G>
__device__ __forceinline__ DevExec<G> make_exec() {
  DevExec<G> ex;
  int wl = threadIdx.x & 31;
  ex.lane = wl % G;
  ex.mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((wl / G) * G));
  return ex;
}

__device__ __forceinline__ void stage_model(const Dims& D, const uint32_t* __restrict__ model, uint32_t* smem) {
  for (int i = threadIdx.x; i < D.model_words; i += blockDim.x) smem[i] = model[i];
  __syncthreads();
}

template <int G>
__global__ void __launch_bounds__(kThreads)
step_kernel(const Dims D, const uint32_t* __restrict__ model, const BxgState in, const float* __restrict__ act,
            const BxgState out, int64_t n_env, int n_frames, int flags, const BxgDiag diag) {
  extern __shared__ __align__(16) uint32_t smem_u[];
  stage_model(D, model, smem_u);
  const int groups = kThreads / G, group = threadIdx.x / G;
  Ctx c;
  c.D = &D;
  c.mf = reinterpret_cast<const float*>(smem_u);
  c.mi = reinterpret_cast<const int*>(smem_u);
  c.s = reinterpret_cast<float*>(smem_u) + D.model_words + group * D.env_words;
  DevExec<G> ex = make_exec<G>();
  for (int64_t e = (int64_t)blockIdx.x * groups + group; e < n_env; e += (int64_t)gridDim.x * groups) {
    Stats st{0, 0, 0, 0};
    load_env(ex, c, in, act, e);
    for (int f = 0; f < n_frames; ++f) substep(ex, c, &st);
    store_env(ex, c, out, e, (flags & BXG_STEP_DIAGNOSTICS) ? &diag : nullptr, st);
  }
}

template <int G>
__global__ void __launch_bounds__(kThreads)
init_kernel(const Dims D, const uint32_t* __restrict__ model, const float* __restrict__ q, const float* __restrict__ qd,
            const BxgState out, int64_t n_env) {
  extern __shared__ __align__(16) uint32_t smem_u[];
  stage_model(D, model, smem_u);
  const int groups = kThreads / G, group = threadIdx.x / G;
  Ctx c;
  c.D = &D;
  c.mf = reinterpret_cast<const float*>(smem_u);
  c.mi = reinterpret_cast<const int*>(smem_u);
  c.s = reinterpret_cast<float*>(smem_u) + D.model_words + group * D.env_words;
  DevExec<G> ex = make_exec<G>();
  for (int64_t e = (int64_t)blockIdx.x * groups + group; e < n_env; e += (int64_t)gridDim.x * groups) {
    Stats st{0, 0, 0, 0};
    load_env_qqd(ex, c, q, qd, e);
    init_env(ex, c, &st);
    store_env(ex, c, out, e, nullptr, st);
  }
}

}  // namespace bxg

// ============================================================== C ABI
struct BxgModel {
  bxg::PackedModel pm;
  int device = 0;
  int lanes = 32;          // G
  int sm_count = 0;
  uint32_t* d_blob = nullptr;
  size_t smem_bytes = 0;
  int blocks_per_sm_step = 1, blocks_per_sm_init = 1;
};

namespace {
thread_local std::string g_err;
std::atomic<int64_t> g_launches{0};

int fail(int code, const std::string& msg) { g_err = msg; return code; }
int cuda_fail(cudaError_t e, const char* what) {
  g_err = std::string(what) + ": " + cudaGetErrorString(e);
  return BXG_E_CUDA;
}
#define BXG_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return cuda_fail(e__, #x); } while (0)

// every leaf must be non-null, except the constraint leaves of a model with no
// constraint rows (their arrays have zero elements)
bool state_ok(const BxgState* s, int nc) {
  if (!s) return false;
  const float* const* p = reinterpret_cast<const float* const*>(s);
  for (size_t i = 0; i < sizeof(BxgState) / sizeof(float*); ++i) {
    const float* const* f = p + i;
    bool con_leaf = f == (const float* const*)&s->con_jac || f == (const float* const*)&s->con_diag || f == (const float* const*)&s->con_aref;
    if (!p[i] && !(con_leaf && nc == 0)) return false;
  }
  return true;
}
}  // namespace

extern "C" {

int bxg_abi_version(void) { return BXG_ABI_VERSION; }
const char* bxg_last_error(void) { return g_err.c_str(); }
int64_t bxg_launch_count(void) { return g_launches.load(); }

int bxg_model_create(const BxgModelDesc* desc, int device, BxgModel** out) {
  if (!desc || !out) return fail(BXG_E_INVALID, "null argument");
  *out = nullptr;
  BxgModel* m = new BxgModel();
  std::string err = bxg::pack_model(*desc, &m->pm);
  if (!err.empty()) { delete m; return fail(BXG_E_UNSUPPORTED, err); }
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) { delete m; return fail(BXG_E_CUDA, "no CUDA device: this library has no CPU fallback"); }
  if (device < 0 || device >= ndev) { delete m; return fail(BXG_E_INVALID, "bad device ordinal"); }
  m->device = device;
  const bxg::Dims& D = m->pm.d;
  // half-warp groups when the tree fits 16 lanes (Ant), else full warps
  m->lanes = (D.L <= 16 && D.nv <= 16) ? 16 : 32;
  int prev = 0;
  cudaGetDevice(&prev);
  auto cleanup = [&](int code) { cudaSetDevice(prev); if (m->d_blob) cudaFree(m->d_blob); delete m; return code; };
  if (cudaSetDevice(device) != cudaSuccess) return cleanup(fail(BXG_E_CUDA, "cudaSetDevice failed"));
  cudaDeviceProp prop;
  if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cleanup(cuda_fail(ce, "cudaGetDeviceProperties"));
  m->sm_count = prop.multiProcessorCount;
  const int groups = bxg::kThreads / m->lanes;
  m->smem_bytes = sizeof(uint32_t) * ((size_t)D.model_words + (size_t)groups * D.env_words);
  if (m->smem_bytes > (size_t)prop.sharedMemPerBlockOptin)
    return cleanup(fail(BXG_E_UNSUPPORTED, "model needs more shared memory per CTA than the device offers"));
  if ((ce = cudaMalloc(&m->d_blob, m->pm.blob.size() * sizeof(uint32_t))) != cudaSuccess) return cleanup(cuda_fail(ce, "cudaMalloc(model)"));
  if ((ce = cudaMemcpy(m->d_blob, m->pm.blob.data(), m->pm.blob.size() * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess)
    return cleanup(cuda_fail(ce, "cudaMemcpy(model)"));
  const void* ks = m->lanes == 16 ? (const void*)bxg::step_kernel<16> : (const void*)bxg::step_kernel<32>;
  const void* ki = m->lanes == 16 ? (const void*)bxg::init_kernel<16> : (const void*)bxg::init_kernel<32>;
  if ((ce = cudaFuncSetAttribute(ks, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem_bytes)) != cudaSuccess) return cleanup(cuda_fail(ce, "cudaFuncSetAttribute(step)"));
  if ((ce = cudaFuncSetAttribute(ki, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)m->smem_bytes)) != cudaSuccess) return cleanup(cuda_fail(ce, "cudaFuncSetAttribute(init)"));
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m->blocks_per_sm_step, ks, bxg::kThreads, m->smem_bytes);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m->blocks_per_sm_init, ki, bxg::kThreads, m->smem_bytes);
  if (m->blocks_per_sm_step < 1 || m->blocks_per_sm_init < 1) return cleanup(fail(BXG_E_UNSUPPORTED, "kernel does not fit on an SM"));
  cudaSetDevice(prev);
  *out = m;
  return BXG_OK;
}

void bxg_model_destroy(BxgModel* m) {
  if (!m) return;
  if (m->d_blob) {
    int prev = 0;
    cudaGetDevice(&prev);
    cudaSetDevice(m->device);
    cudaFree(m->d_blob);
    cudaSetDevice(prev);
  }
  delete m;
}

int bxg_model_num_constraints(const BxgModel* m) { return m ? m->pm.d.nc : -1; }

static int grid_for(const BxgModel* m, int64_t n_env, int blocks_per_sm) {
  const int groups = bxg::kThreads / m->lanes;
  int64_t need = (n_env + groups - 1) / groups;
  int64_t cap = (int64_t)m->sm_count * blocks_per_sm;
  return (int)(need < cap ? need : cap);
}

int bxg_init(const BxgModel* m, int64_t n_env, const float* q, const float* qd, const BxgState* out, void* stream) {
  if (!m) return fail(BXG_E_INVALID, "null model");
  if (n_env <= 0) return n_env == 0 ? BXG_OK : fail(BXG_E_INVALID, "n_env < 0");
  if (!q || !qd || !state_ok(out, m->pm.d.nc)) return fail(BXG_E_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(m, n_env, m->blocks_per_sm_init);
  if (m->lanes == 16) bxg::init_kernel<16><<<grid, bxg::kThreads, m->smem_bytes, st>>>(m->pm.d, m->d_blob, q, qd, *out, n_env);
  else bxg::init_kernel<32><<<grid, bxg::kThreads, m->smem_bytes, st>>>(m->pm.d, m->d_blob, q, qd, *out, n_env);
  g_launches.fetch_add(1);
  BXG_CUDA(cudaGetLastError());
  return BXG_OK;
}

int bxg_step(const BxgModel* m, int64_t n_env, int32_t n_frames, const BxgState* in, const float* act, const BxgState* out,
             int32_t flags, const BxgDiag* diag, void* stream) {
  if (!m) return fail(BXG_E_INVALID, "null model");
  if (n_frames < 0) return fail(BXG_E_INVALID, "n_frames < 0");
  if (n_env <= 0) return n_env == 0 ? BXG_OK : fail(BXG_E_INVALID, "n_env < 0");
  if (!state_ok(in, m->pm.d.nc) || !state_ok(out, m->pm.d.nc)) return fail(BXG_E_INVALID, "null argument");
  if (m->pm.d.nu > 0 && !act) return fail(BXG_E_INVALID, "act is NULL but the model has actuators");
  BxgDiag dg{nullptr, nullptr};
  if ((flags & BXG_STEP_DIAGNOSTICS) && diag) dg = *diag;
  cudaStream_t st = (cudaStream_t)stream;
  int grid = grid_for(m, n_env, m->blocks_per_sm_step);
  if (m->lanes == 16) bxg::step_kernel<16><<<grid, bxg::kThreads, m->smem_bytes, st>>>(m->pm.d, m->d_blob, *in, act, *out, n_env, n_frames, flags, dg);
  else bxg::step_kernel<32><<<grid, bxg::kThreads, m->smem_bytes, st>>>(m->pm.d, m->d_blob, *in, act, *out, n_env, n_frames, flags, dg);
  g_launches.fetch_add(1);
  BXG_CUDA(cudaGetLastError());
  return BXG_OK;
}

}  // extern "C"
