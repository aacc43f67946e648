// bxg_sim.cpp -- TEST-ONLY host emulation of the kernel's lane-group execution.
//
// Compiles brax_b200/csrc/bxg_core.cuh (the exact algorithm source the CUDA
// kernel is built from) with g++ and runs the G lanes of a group as a loop, so
// the CPU test-suite can check kernel LOGIC (indexing, phase structure, the
// solver state machine) against the oracle without a GPU.  `reverse` runs the
// lanes of every phase in descending order: any result difference between the
// two orders exposes an intra-phase cross-lane dependency (a race on device).
// mode bit 0 = reverse lane order, bit 1 = force the generic (non register-row) kernels.
//
// This is not a CPU fallback: it lives under tests/, is never built by the
// package and nothing in brax_b200/ can load it.
#include <string.h>

#include <string>
#include <vector>

#include "../../brax_b200/csrc/bxg_core.cuh"

namespace {

template <int G_>
struct HostExec {
  static constexpr int G = G_;
  bool reverse = false;
  struct LaneF {
    float v[G_];
    float& operator()(int l) { return v[l]; }
  };
  template <int N>
  struct LaneVec {
    float v[G_][N];
    float* operator()(int l) { return v[l]; }
  };
  void sync() {}
  void cta_sync() {}
  template <class F>
  void lanes(F&& f) {
    if (!reverse) for (int l = 0; l < G; ++l) f(l);
    else for (int l = G - 1; l >= 0; --l) f(l);
  }
  float sum(LaneF& p) {  // same pairing as the xor-shuffle butterfly
    float a[G_], b[G_];
    memcpy(a, p.v, sizeof a);
    for (int o = G / 2; o >= 1; o >>= 1) { for (int l = 0; l < G; ++l) b[l] = a[l] + a[l ^ o]; memcpy(a, b, sizeof a); }
    return a[0];
  }
  void sum_max(LaneF& ps, LaneF& pm, float* s_out, float* m_out) { *s_out = sum(ps); *m_out = max(pm); }
  void sum3(LaneF& p0, LaneF& p1, LaneF& p2, float* o0, float* o1, float* o2) { *o0 = sum(p0); *o1 = sum(p1); *o2 = sum(p2); }
  uint32_t ballot(LaneF& p) {
    uint32_t b = 0;
    for (int l = 0; l < G; ++l) if (p.v[l] != 0.f) b |= 1u << l;
    return b;
  }
  float max(LaneF& p) {
    float m = p.v[0];
    for (int l = 1; l < G; ++l) m = fmaxf(m, p.v[l]);
    return m;
  }
};

template <class Cfg>
int run(const BxgModelDesc* desc, int vid, int mode, bool init, int64_t n_env, int n_frames, const float* q, const float* qd,
        const BxgState* in, const float* act, const BxgState* out, int flags, const BxgDiag* diag,
        const BxgEnvSpec* env = nullptr, const BxgEnvIO* eio = nullptr) {
  constexpr int G = Cfg::G;
  bxg::PackedModel pm;
  std::string err = bxg::pack_model(*desc, &pm, vid);
  if (!err.empty()) return 3;
  const bool reverse = mode & 1;
  pm.d.force_generic = (mode & 2) ? 1 : 0;
  std::vector<float> slab_store(pm.d.env_words + 4);
  float* slab_base = slab_store.data();
  while (reinterpret_cast<uintptr_t>(slab_base) % 16) ++slab_base;
  bxg::Ctx c;
  c.D = &pm.d;
  c.mf = reinterpret_cast<const float*>(pm.blob.data());
  c.mi = reinterpret_cast<const int*>(pm.blob.data());
  c.s = slab_base;
  HostExec<G> ex;
  ex.reverse = reverse;
  for (int64_t e = 0; e < n_env; ++e) {
    // poison the slab so stale-data bugs show up as NaN
    for (int i = 0; i < pm.d.env_words; ++i) slab_base[i] = NAN;
    bxg::Stats st{0, 0, 0, 0};
    bxg::prepare_env(ex, c);
    if (init) {
      bxg::load_env_qqd(ex, c, q, qd, e);
      bxg::init_env<HostExec<G>, Cfg>(ex, c, &st);
      if (env) bxg::env_reset_obs(ex, c, *env, eio->obs + e * bxg::env_obs_size(pm.d, *env));
      bxg::store_env(ex, c, *out, e, nullptr, st);
    } else {
      bxg::load_env(ex, c, *in, act, e);
      if (env) bxg::env_prologue(ex, c, *env, *in, e);
      for (int f = 0; f < n_frames; ++f) {
        if (pm.d.minv_mode == BXG_MINV_CHOLESKY) bxg::substep<HostExec<G>, Cfg, 1>(ex, c, &st);
        else bxg::substep<HostExec<G>, Cfg, 0>(ex, c, &st);
      }
      bool done = false;
      if (env) bxg::env_epilogue(ex, c, *env, *eio, e, true, &done);
      if (done && eio && eio->first_state) bxg::store_first_state(ex, c, *out, *eio->first_state, e);
      else bxg::store_env(ex, c, *out, e, (flags & BXG_STEP_DIAGNOSTICS) ? diag : nullptr, st);
    }
  }
  return 0;
}

}  // namespace

template <class... A>
int dispatch(const BxgModelDesc* desc, int vid, A... a) {
  if (vid < 0) {
    bxg::PackedModel pm;
    if (!bxg::pack_model(*desc, &pm).empty()) return 3;
    vid = pm.variant_id;
  }
  switch (vid) {
    case 0: return run<bxg::KernelCfg<16, 4, 6, true>>(desc, vid, a...);
    case 1: return run<bxg::KernelCfg<32, 6, 7, true>>(desc, vid, a...);
    case 2: return run<bxg::KernelCfg<32, 8, 8, true>>(desc, vid, a...);
    case 3: return run<bxg::KernelCfg<32, 0, 0>>(desc, vid, a...);
    case 4: return run<bxg::KernelCfg<16, 6, 7, true>>(desc, vid, a...);
    case 5: return run<bxg::KernelCfg<32, 4, 16, true>>(desc, vid, a...);
    case 6: return run<bxg::KernelCfg<32, 6, 20, true>>(desc, vid, a...);
    case 7: return run<bxg::KernelCfg<4, 1, 1, true>>(desc, vid, a...);
    case 8: return run<bxg::KernelCfg<4, 2, 2, true>>(desc, vid, a...);
    case 9: return run<bxg::KernelCfg<4, 2, 4, true>>(desc, vid, a...);
  }
  return 3;
}

extern "C" {
// variant: -1 = the one the library would pick, else a forced kernel variant id
int sim_init(const BxgModelDesc* desc, int variant, int mode, int64_t n_env, const float* q, const float* qd, const BxgState* out) {
  return dispatch(desc, variant, mode, true, n_env, 0, q, qd, (const BxgState*)nullptr, (const float*)nullptr, out, 0, (const BxgDiag*)nullptr,
                  (const BxgEnvSpec*)nullptr, (const BxgEnvIO*)nullptr);
}
int sim_step(const BxgModelDesc* desc, int variant, int mode, int64_t n_env, int n_frames, const BxgState* in, const float* act,
             const BxgState* out, int flags, const BxgDiag* diag) {
  return dispatch(desc, variant, mode, false, n_env, n_frames, (const float*)nullptr, (const float*)nullptr, in, act, out, flags, diag,
                  (const BxgEnvSpec*)nullptr, (const BxgEnvIO*)nullptr);
}
int sim_env_reset(const BxgModelDesc* desc, int variant, int mode, const BxgEnvSpec* spec, int64_t n_env, const float* q,
                  const float* qd, const BxgState* out, float* obs) {
  BxgEnvIO io{}; io.obs = obs;
  return dispatch(desc, variant, mode, true, n_env, 0, q, qd, (const BxgState*)nullptr, (const float*)nullptr, out, 0,
                  (const BxgDiag*)nullptr, spec, (const BxgEnvIO*)&io);
}
int sim_env_step(const BxgModelDesc* desc, int variant, int mode, const BxgEnvSpec* spec, int64_t n_env, int n_frames,
                 const BxgState* in, const float* action, const BxgState* out, const BxgEnvIO* io) {
  return dispatch(desc, variant, mode, false, n_env, n_frames, (const float*)nullptr, (const float*)nullptr, in, action, out, 0,
                  (const BxgDiag*)nullptr, spec, io);
}
}
