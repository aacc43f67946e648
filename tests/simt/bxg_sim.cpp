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
// Built twice (tests/simt/sim.py): libbxg_sim.so with the product's scalar type (float), and
// libbxg_sim_f64.so with -DBXG_REAL=double: the same algorithm source in double precision,
// State leaves as double arrays, so that every variant's LOGIC can be held against the
// reference-source goldens (float64) to 1e-9 on every leaf (tests/test_kernel_logic_f64.py).
// The double build carries the physics path only (init / step), not the env epilogue, whose
// reference constants are Python floats (float64) that BxgEnvSpec holds as float.
//
// This is not a CPU fallback: it lives under tests/, is never built by the
// package and nothing in brax_b200/ can load it.
#include <stddef.h>
#include <stdio.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../brax_b200/csrc/bxg_core.cuh"

namespace {

using bxg::real;
constexpr bool kF64 = sizeof(real) == 8;

// BxgState with leaves of the scalar type (same member names and order: the Python side passes
// one struct of 25 pointers either way)
struct SimState {
  real *q, *qd, *x_pos, *x_rot, *xd_ang, *xd_vel, *root_com, *cinr_pos, *cinr_rot, *cinr_i, *cinr_mass, *cd_ang, *cd_vel,
       *cdof_ang, *cdof_vel, *cdofd_ang, *cdofd_vel, *mass_mx, *mass_mx_inv, *con_jac, *con_diag, *con_aref,
       *qf_smooth, *qf_constraint, *qdd;
};
static_assert(sizeof(SimState) == sizeof(BxgState), "SimState mirrors BxgState");

template <int G_>
struct HostExec {
  static constexpr int G = G_;
  bool reverse = false;
  struct LaneF {
    real v[G_];
    real& operator()(int l) { return v[l]; }
  };
  template <int N>
  struct LaneVec {
    real v[G_][N];
    real* operator()(int l) { return v[l]; }
  };
  void sync() {}
  void cta_sync() {}
  template <class F>
  void lanes(F&& f) {
    if (!reverse) for (int l = 0; l < G; ++l) f(l);
    else for (int l = G - 1; l >= 0; --l) f(l);
  }
  real sum(LaneF& p) {  // same pairing as the xor-shuffle butterfly
    real a[G_], b[G_];
    memcpy(a, p.v, sizeof a);
    for (int o = G / 2; o >= 1; o >>= 1) { for (int l = 0; l < G; ++l) b[l] = a[l] + a[l ^ o]; memcpy(a, b, sizeof a); }
    return a[0];
  }
  void sum_max(LaneF& ps, LaneF& pm, real* s_out, real* m_out) { *s_out = sum(ps); *m_out = max(pm); }
  void sum3(LaneF& p0, LaneF& p1, LaneF& p2, real* o0, real* o1, real* o2) { *o0 = sum(p0); *o1 = sum(p1); *o2 = sum(p2); }
  uint32_t ballot(LaneF& p) {
    uint32_t b = 0;
    for (int l = 0; l < G; ++l) if (p.v[l] != real(0)) b |= 1u << l;
    return b;
  }
  real max(LaneF& p) {
    real m = p.v[0];
    for (int l = 1; l < G; ++l) m = bxg::r_max(m, p.v[l]);
    return m;
  }
};

template <class Cfg>
int run(const BxgModelDesc* desc, int vid, int mode, bool init, int64_t n_env, int n_frames, const real* q, const real* qd,
        const SimState* in, const real* act, const SimState* out, int flags, const BxgDiag* diag,
        const BxgEnvSpec* env = nullptr, const BxgEnvIO* eio = nullptr) {
  constexpr int G = Cfg::G;
  bxg::PackedModelT<real> pm;
  std::string err = bxg::pack_model_t<real>(*desc, &pm, vid);
  if (!err.empty()) return 3;
  const bool reverse = mode & 1;
  pm.d.force_generic = (mode & 2) ? 1 : 0;
  std::vector<real> slab_store(pm.d.env_words + 4);
  real* slab_base = slab_store.data();
  while (reinterpret_cast<uintptr_t>(slab_base) % 16) ++slab_base;
  bxg::Ctx c;
  c.D = &pm.d;
  if constexpr (kF64) c.mf = reinterpret_cast<const real*>(pm.blob_r.data());
  else c.mf = reinterpret_cast<const real*>(pm.blob.data());
  c.mi = reinterpret_cast<const int*>(pm.blob.data());
  c.s = slab_base;
  HostExec<G> ex;
  ex.reverse = reverse;
  for (int64_t e = 0; e < n_env; ++e) {
    // poison the slab so stale-data bugs show up as NaN
    for (int i = 0; i < pm.d.env_words; ++i) slab_base[i] = NAN;
    bxg::Stats st{};
    bxg::prepare_env(ex, c);
    if (init) {
      bxg::load_env_qqd(ex, c, q, qd, e);
      bxg::init_env<HostExec<G>, Cfg>(ex, c, &st);
      bxg::finish_env(ex, c);
      if constexpr (!kF64) { if (env) bxg::env_reset_obs(ex, c, *env, eio->obs + e * bxg::env_obs_size(pm.d, *env)); }
      bxg::store_env(ex, c, *out, e, nullptr, st);
    } else {
      const bool lean = (flags & BXG_STEP_LEAN) || (eio && (eio->flags & BXG_STEP_LEAN));
      if (lean) { bxg::load_env_lean(ex, c, *in, act, e); bxg::lean_entry<HostExec<G>, Cfg>(ex, c); }
      else bxg::load_env(ex, c, *in, act, e);
      if constexpr (!kF64) { if (env) bxg::env_prologue(ex, c, *env, *in, e); }
      for (int f = 0; f < n_frames; ++f) {
        if (pm.d.minv_mode == BXG_MINV_CHOLESKY) bxg::substep<HostExec<G>, Cfg, 1>(ex, c, &st);
        else bxg::substep<HostExec<G>, Cfg, 0>(ex, c, &st);
      }
      bool done = false;
      bxg::finish_env(ex, c);
      if constexpr (!kF64) {
        if (env) bxg::env_epilogue(ex, c, *env, *eio, e, true, &done);
        if (done && eio && eio->first_state) {
          if (lean) bxg::store_first_state_lean(ex, c, *out, *reinterpret_cast<const SimState*>(eio->first_state), e);
          else bxg::store_first_state(ex, c, *out, *reinterpret_cast<const SimState*>(eio->first_state), e);
          continue;
        }
      }
      if (lean) bxg::store_env_lean(ex, c, *out, e, (flags & BXG_STEP_DIAGNOSTICS) ? diag : nullptr, st);
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
int sim_sizeof_real() { return (int)sizeof(real); }
// The packed Dims of a model as a C++ aggregate initializer (bxg_model.h field order): input of
// tools/gen_const_dims.py, which writes the headers the model-specialised kernel variants are compiled with.
int sim_dims_initializer(const BxgModelDesc* desc, int variant, int phase_groups, char* buf, int cap) {
  bxg::PackedModel pm;
  if (!bxg::pack_model(*desc, &pm, variant).empty()) return -1;
  if (phase_groups > 0) pm.d.phase_groups = phase_groups;   // a tuning value of the specialised build, not part of the model's layout
  static_assert(sizeof(bxg::Dims) % 4 == 0, "Dims is a sequence of 4-byte fields");
  const size_t f0 = offsetof(bxg::Dims, dt) / 4, f1 = offsetof(bxg::Dims, gz) / 4;
  const uint32_t* w = reinterpret_cast<const uint32_t*>(&pm.d);
  std::string out = "{";
  for (size_t i = 0; i < sizeof(bxg::Dims) / 4; ++i) {
    char tmp[64];
    if (i >= f0 && i <= f1) { float f; memcpy(&f, w + i, 4); snprintf(tmp, sizeof tmp, "%af", (double)f); }
    else snprintf(tmp, sizeof tmp, "%d", (int)w[i]);
    out += tmp; out += i + 1 < sizeof(bxg::Dims) / 4 ? ", " : "}";
  }
  if ((int)out.size() + 1 > cap) return -2;
  memcpy(buf, out.c_str(), out.size() + 1);
  return pm.variant_id;
}
// the kernel's sine / cosine (bxg_core.cuh r_sincos) on n values: tests/test_kernel_math.py
void sim_sincos(const real* x, int n, real* s, real* c) { for (int i = 0; i < n; ++i) bxg::r_sincos(x[i], s + i, c + i); }
// variant: -1 = the one the library would pick, else a forced kernel variant id
int sim_init(const BxgModelDesc* desc, int variant, int mode, int64_t n_env, const real* q, const real* qd, const SimState* out) {
  return dispatch(desc, variant, mode, true, n_env, 0, q, qd, (const SimState*)nullptr, (const real*)nullptr, out, 0, (const BxgDiag*)nullptr,
                  (const BxgEnvSpec*)nullptr, (const BxgEnvIO*)nullptr);
}
int sim_step(const BxgModelDesc* desc, int variant, int mode, int64_t n_env, int n_frames, const SimState* in, const real* act,
             const SimState* out, int flags, const BxgDiag* diag) {
  return dispatch(desc, variant, mode, false, n_env, n_frames, (const real*)nullptr, (const real*)nullptr, in, act, out, flags, diag,
                  (const BxgEnvSpec*)nullptr, (const BxgEnvIO*)nullptr);
}
#if !defined(BXG_SIM_F64)
int sim_env_reset(const BxgModelDesc* desc, int variant, int mode, const BxgEnvSpec* spec, int64_t n_env, const real* q,
                  const real* qd, const SimState* out, real* obs) {
  BxgEnvIO io{}; io.obs = obs;
  return dispatch(desc, variant, mode, true, n_env, 0, q, qd, (const SimState*)nullptr, (const real*)nullptr, out, 0,
                  (const BxgDiag*)nullptr, spec, (const BxgEnvIO*)&io);
}
int sim_env_step(const BxgModelDesc* desc, int variant, int mode, const BxgEnvSpec* spec, int64_t n_env, int n_frames,
                 const SimState* in, const real* action, const SimState* out, const BxgEnvIO* io) {
  return dispatch(desc, variant, mode, false, n_env, n_frames, (const real*)nullptr, (const real*)nullptr, in, action, out, 0,
                  (const BxgDiag*)nullptr, spec, io);
}
#endif
}
