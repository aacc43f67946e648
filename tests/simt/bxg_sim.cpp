// bxg_sim.cpp -- TEST-ONLY host emulation of the kernel's lane-group execution.
//
// Compiles brax_b200/csrc/bxg_core.cuh (the exact algorithm source the CUDA
// kernel is built from) with g++ and runs the G lanes of a group as a loop, so
// the CPU test-suite can check kernel LOGIC (indexing, phase structure, the
// solver state machine) against the oracle without a GPU.  `reverse` runs the
// lanes of every phase in descending order: any result difference between the
// two orders exposes an intra-phase cross-lane dependency (a race on device).
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
  void sync() {}
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
  float max(LaneF& p) {
    float m = p.v[0];
    for (int l = 1; l < G; ++l) m = fmaxf(m, p.v[l]);
    return m;
  }
};

template <int G>
int run(const BxgModelDesc* desc, bool reverse, bool init, int64_t n_env, int n_frames, const float* q, const float* qd,
        const BxgState* in, const float* act, const BxgState* out, int flags, const BxgDiag* diag) {
  bxg::PackedModel pm;
  std::string err = bxg::pack_model(*desc, &pm);
  if (!err.empty()) return 3;
  if (pm.d.L > G) return 3;
  std::vector<float> slab(pm.d.env_words);
  bxg::Ctx c;
  c.D = &pm.d;
  c.mf = reinterpret_cast<const float*>(pm.blob.data());
  c.mi = reinterpret_cast<const int*>(pm.blob.data());
  c.s = slab.data();
  HostExec<G> ex;
  ex.reverse = reverse;
  for (int64_t e = 0; e < n_env; ++e) {
    // poison the slab so stale-data bugs show up as NaN
    for (auto& v : slab) v = NAN;
    bxg::Stats st{0, 0, 0, 0};
    if (init) {
      bxg::load_env_qqd(ex, c, q, qd, e);
      bxg::init_env(ex, c, &st);
      bxg::store_env(ex, c, *out, e, nullptr, st);
    } else {
      bxg::load_env(ex, c, *in, act, e);
      for (int f = 0; f < n_frames; ++f) bxg::substep(ex, c, &st);
      bxg::store_env(ex, c, *out, e, (flags & BXG_STEP_DIAGNOSTICS) ? diag : nullptr, st);
    }
  }
  return 0;
}

}  // namespace

extern "C" {
int sim_init(const BxgModelDesc* desc, int G, int reverse, int64_t n_env, const float* q, const float* qd, const BxgState* out) {
  if (G == 16) return run<16>(desc, reverse, true, n_env, 0, q, qd, nullptr, nullptr, out, 0, nullptr);
  return run<32>(desc, reverse, true, n_env, 0, q, qd, nullptr, nullptr, out, 0, nullptr);
}
int sim_step(const BxgModelDesc* desc, int G, int reverse, int64_t n_env, int n_frames, const BxgState* in, const float* act,
             const BxgState* out, int flags, const BxgDiag* diag) {
  if (G == 16) return run<16>(desc, reverse, false, n_env, n_frames, nullptr, nullptr, in, act, out, flags, diag);
  return run<32>(desc, reverse, false, n_env, n_frames, nullptr, nullptr, in, act, out, flags, diag);
}
}
