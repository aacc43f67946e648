// bxg_kernels.cuh -- sm_100a kernel templates (instantiated per variant by
// bxg_inst.cu; the C ABI that launches them is bxg_api.cu).
//
// One lane-group (G = 16 or 32 lanes) per environment, whole working state in
// shared memory, all n_frames substeps inside one launch; HBM is touched only
// at kernel entry (load_env) and exit (store_env).  See bxg_core.cuh for the
// algorithm and DESIGN.md for the layout and roofline accounting.
#pragma once
#include <cuda_runtime.h>

#include "bxg_core.cuh"

namespace bxg {

// ------------------------------------------------------- device executor
template <int G_>
struct DevExec {
  static constexpr int G = G_;
  int lane;
  unsigned mask;
  int bar_id, bar_threads;   // named barrier of this warp's phase group
  struct LaneF {
    float v;
    __device__ __forceinline__ float& operator()(int) { return v; }
  };
  template <int N>
  struct LaneVec {   // per-lane register array that persists across lanes() calls
    float v[N];
    __device__ __forceinline__ float* operator()(int) { return v; }
  };
  __device__ __forceinline__ void sync() { __syncwarp(mask); }
  // CTA-wide phase alignment: every warp of the CTA streams the same straight-line
  // code at the same time, so instruction-cache lines are fetched once per CTA
  __device__ __forceinline__ void cta_sync() {
    asm volatile("bar.sync %0, %1;" ::"r"(bar_id), "r"(bar_threads) : "memory");
  }
  template <class F>
  __device__ __forceinline__ void lanes(F&& f) {
    f(lane);
    __syncwarp(mask);   // group barrier + memory ordering for the next phase
  }
  __device__ __forceinline__ float sum(LaneF& p) {
    float v = p.v;
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) v += __shfl_xor_sync(mask, v, o, G);
    return v;
  }
  // sum and max in one butterfly (two independent shuffle chains interleave)
  __device__ __forceinline__ void sum_max(LaneF& ps, LaneF& pm, float* s_out, float* m_out) {
    float a = ps.v, b = pm.v;
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) {
      float a2 = __shfl_xor_sync(mask, a, o, G), b2 = __shfl_xor_sync(mask, b, o, G);
      a += a2; b = fmaxf(b, b2);
    }
    *s_out = a; *m_out = b;
  }
  // three sums in one butterfly
  __device__ __forceinline__ void sum3(LaneF& p0, LaneF& p1, LaneF& p2, float* o0, float* o1, float* o2) {
    float a = p0.v, b = p1.v, c = p2.v;
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) {
      float a2 = __shfl_xor_sync(mask, a, o, G), b2 = __shfl_xor_sync(mask, b, o, G), c2 = __shfl_xor_sync(mask, c, o, G);
      a += a2; b += b2; c += c2;
    }
    *o0 = a; *o1 = b; *o2 = c;
  }
  // bit l of the result = (lane l of this group holds a non-zero value)
  __device__ __forceinline__ uint32_t ballot(LaneF& p) {
    uint32_t b = __ballot_sync(mask, p.v != 0.f);
    if (G == 32) return b;
    return (b >> ((threadIdx.x & 31) / G * G)) & ((1u << G) - 1u);
  }
  __device__ __forceinline__ float max(LaneF& p) {
    float v = p.v;
#pragma unroll
    for (int o = G / 2; o >= 1; o >>= 1) v = fmaxf(v, __shfl_xor_sync(mask, v, o, G));
    return v;
  }
};


template <int G>
__device__ __forceinline__ DevExec<G> make_exec(int phase_groups) {
  DevExec<G> ex;
  // the warps of a CTA are split into `phase_groups` groups that align their
  // phases independently (named barriers 1..): within a group the warps share
  // instruction-cache lines, across groups FMA-bound and latency-bound phases overlap
  {
    const int nw = blockDim.x >> 5, w = threadIdx.x >> 5;
    int pg = phase_groups < 1 ? 1 : (phase_groups > nw ? nw : phase_groups);
    const int base = nw / pg, rem = nw % pg;      // first `rem` groups have base + 1 warps
    int g = 0, start = 0;
    for (; g < pg; ++g) { int sz = base + (g < rem ? 1 : 0); if (w < start + sz) { ex.bar_threads = sz * 32; break; } start += sz; }
    ex.bar_id = 1 + g;
  }
  int wl = threadIdx.x & 31;
  ex.lane = wl % G;
  ex.mask = G == 32 ? 0xffffffffu : (((1u << G) - 1u) << ((wl / G) * G));
  return ex;
}

__device__ __forceinline__ void stage_model(const Dims& D, const uint32_t* __restrict__ model, uint32_t* smem) {
  for (int i = threadIdx.x; i < D.model_words; i += blockDim.x) smem[i] = model[i];
  __syncthreads();
}

// LEAN (BXG_STEP_LEAN, include/bxg.h) is a compile-time mode: with the lean entry code in the same kernel the
// default path loses 4 % (Ant) to the larger register / code footprint (profiles/r02_ab_lean_code.txt).
// ID: the kernel id of the translation unit (bxg_inst.cu).  It makes every instantiation a distinct symbol: ids 10 / 11
// use the same Cfg as variants 0 / 1 and would otherwise be ONE weak symbol for two different kernels.
// PERENV: one model per env (bxg_model_create_batched: the reference's DomainRandomizationVmapWrapper, a System with
// batched leaves, envs/wrappers/training.py:223-260).  `model` is then [n_env][model_words] in global memory and env e
// reads its constants from its own row (through L1 / L2) instead of the CTA's shared-memory copy; the slab layout is
// unchanged.  A compile-time mode: the default kernels keep their shared-memory addressing.
template <class Cfg, int INV, bool LEAN, int ID, bool PERENV = false>
__global__ void __launch_bounds__(Cfg::MAX_THREADS)
step_kernel(const Dims Dparam, const uint32_t* __restrict__ model, const BxgState in, const float* __restrict__ act,
            const BxgState out, int64_t n_env, int n_frames, int flags, const BxgDiag diag,
            const BxgEnvSpec env, const BxgEnvIO eio, const BxgState first) {
  extern __shared__ __align__(16) uint32_t smem_u[];
  constexpr int G = Cfg::G;
#if defined(BXG_CONST_DIMS)
  constexpr Dims D = BXG_CONST_DIMS_FN();   // (the launch passes the same values in Dparam: bxg_api.cu checked them at model creation)
#else
  const Dims& D = Dparam;
#endif
  if constexpr (!PERENV) stage_model(D, model, smem_u);
  const int groups = blockDim.x / G, group = threadIdx.x / G;
  Ctx c;
  c.D = &Dparam;
  c.mf = reinterpret_cast<const float*>(smem_u);
  c.mi = reinterpret_cast<const int*>(smem_u);
  c.s = reinterpret_cast<float*>(smem_u) + D.model_words + group * D.env_words;
  DevExec<G> ex = make_exec<G>(D.phase_groups);
  // uniform trip count per CTA (phases are CTA-synchronous): groups past the end
  // of the batch redo the last env and skip the store
  const int64_t per_pass = (int64_t)gridDim.x * groups;
  const int64_t passes = (n_env + per_pass - 1) / per_pass;
  for (int64_t p = 0; p < passes; ++p) {
    int64_t e = p * per_pass + (int64_t)blockIdx.x * groups + group;
    const bool valid = e < n_env;
    if (!valid) e = n_env - 1;
    if constexpr (PERENV) {
      const uint32_t* mb = model + e * (int64_t)D.model_words;
      c.mf = reinterpret_cast<const float*>(mb); c.mi = reinterpret_cast<const int*>(mb);
    }
    Stats st{};
#if defined(BXG_PHASE_TIMERS)
    st.phase_cycles = valid ? diag.phase_cycles : nullptr;
#endif
    constexpr bool lean = LEAN;
    BXG_PHASE_BEGIN(&st);
    prepare_env(ex, c);
    if constexpr (lean) load_env_lean(ex, c, in, act, e);
    else load_env(ex, c, in, act, e);
    BXG_PHASE_END(&st, 0);
    if constexpr (lean) { lean_entry<DevExec<G>, Cfg>(ex, c); BXG_PHASE_END(&st, 11); }
    if (env.kind) { env_prologue(ex, c, env, in, e); BXG_PHASE_END(&st, 9); }
    for (int f = 0; f < n_frames; ++f) substep<DevExec<G>, Cfg, INV>(ex, c, &st);
    bool done = false;
    BXG_PHASE_BEGIN(&st);
    finish_env(ex, c);      // link velocities xd (once, not per substep)
    BXG_PHASE_END(&st, 4);
    if (env.kind) { env_epilogue(ex, c, env, eio, e, valid, &done); BXG_PHASE_END(&st, 9); }
    if (valid) {
      const BxgDiag* dg = (flags & BXG_STEP_DIAGNOSTICS) ? &diag : nullptr;
      if (done && eio.first_state) {   // AutoResetWrapper
        if constexpr (lean) store_first_state_lean(ex, c, out, first, e); else store_first_state(ex, c, out, first, e);
      } else {
        if constexpr (lean) store_env_lean(ex, c, out, e, dg, st); else store_env(ex, c, out, e, dg, st);
      }
    }
    BXG_PHASE_END(&st, 10);
  }
}

template <class Cfg, int ID, bool PERENV = false>
__global__ void __launch_bounds__(Cfg::MAX_THREADS)
init_kernel(const Dims Dparam, const uint32_t* __restrict__ model, const float* __restrict__ q, const float* __restrict__ qd,
            const BxgState out, int64_t n_env, const BxgEnvSpec env, float* __restrict__ obs) {
  extern __shared__ __align__(16) uint32_t smem_u[];
  constexpr int G = Cfg::G;
#if defined(BXG_CONST_DIMS)
  constexpr Dims D = BXG_CONST_DIMS_FN();   // (the launch passes the same values in Dparam: bxg_api.cu checked them at model creation)
#else
  const Dims& D = Dparam;
#endif
  if constexpr (!PERENV) stage_model(D, model, smem_u);
  const int groups = blockDim.x / G, group = threadIdx.x / G;
  Ctx c;
  c.D = &Dparam;
  c.mf = reinterpret_cast<const float*>(smem_u);
  c.mi = reinterpret_cast<const int*>(smem_u);
  c.s = reinterpret_cast<float*>(smem_u) + D.model_words + group * D.env_words;
  DevExec<G> ex = make_exec<G>(D.phase_groups);
  const int64_t per_pass = (int64_t)gridDim.x * groups;
  const int64_t passes = (n_env + per_pass - 1) / per_pass;
  for (int64_t p = 0; p < passes; ++p) {
    int64_t e = p * per_pass + (int64_t)blockIdx.x * groups + group;
    const bool valid = e < n_env;
    if (!valid) e = n_env - 1;
    if constexpr (PERENV) {
      const uint32_t* mb = model + e * (int64_t)D.model_words;
      c.mf = reinterpret_cast<const float*>(mb); c.mi = reinterpret_cast<const int*>(mb);
    }
    Stats st{};
    prepare_env(ex, c);
    load_env_qqd(ex, c, q, qd, e);
    init_env<DevExec<G>, Cfg>(ex, c, &st);
    finish_env(ex, c);
    if (env.kind && valid) env_reset_obs(ex, c, env, obs + e * env_obs_size(D, env));
    if (valid) store_env(ex, c, out, e, nullptr, st);
  }
}

}  // namespace bxg
