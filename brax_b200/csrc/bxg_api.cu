// bxg_api.cu -- the C ABI declared in include/bxg.h: model upload, launch
// configuration and the two launches (init, step).  Kernels live in
// bxg_kernels.cuh and are instantiated per variant in bxg_inst.cu.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <string>

#include "bxg_model.h"
#include "gen/bxg_dims_ant.h"
#include "gen/bxg_dims_humanoid.h"

namespace bxg {
constexpr size_t kSmemBudget = 227 * 1024;  // usable shared memory per SM on sm_100
// envs per CTA: as many as fit the shared-memory budget of one SM (one CTA per
// SM, all of its warps phase-aligned), whole warps only
inline int envs_per_cta(const Dims& d, const Variant& var) {
  const int G = var.G;
  size_t avail = kSmemBudget - sizeof(uint32_t) * (size_t)d.model_words;
  int n = (int)(avail / (sizeof(uint32_t) * (size_t)d.env_words));
  int per_warp = 32 / G;
  n -= n % per_warp;
  int cap = variant_max_threads(var.G, var.VC4, var.NC4) / G;
  if (const char* e = getenv("BXG_MAX_ENVS_PER_CTA")) { int v = atoi(e); if (v >= per_warp && v < cap) cap = v - v % per_warp; }   // tuning knob
  return n > cap ? cap : n;
}
}

extern "C" {
const void* bxg_step_kernel_v0(); const void* bxg_step_kernel_v1(); const void* bxg_step_kernel_v2(); const void* bxg_step_kernel_v3(); const void* bxg_step_kernel_v4(); const void* bxg_step_kernel_v5(); const void* bxg_step_kernel_v6(); const void* bxg_step_kernel_v7(); const void* bxg_step_kernel_v8(); const void* bxg_step_kernel_v9(); const void* bxg_step_kernel_v10(); const void* bxg_step_kernel_v11();
const void* bxg_init_kernel_v0(); const void* bxg_init_kernel_v1(); const void* bxg_init_kernel_v2(); const void* bxg_init_kernel_v3(); const void* bxg_init_kernel_v4(); const void* bxg_init_kernel_v5(); const void* bxg_init_kernel_v6(); const void* bxg_init_kernel_v7(); const void* bxg_init_kernel_v8(); const void* bxg_init_kernel_v9(); const void* bxg_init_kernel_v10(); const void* bxg_init_kernel_v11();
}
extern "C" {
const void* bxg_step_chol_kernel_v0(); const void* bxg_step_chol_kernel_v1(); const void* bxg_step_chol_kernel_v2();
const void* bxg_step_chol_kernel_v3(); const void* bxg_step_chol_kernel_v4(); const void* bxg_step_chol_kernel_v5(); const void* bxg_step_chol_kernel_v6(); const void* bxg_step_chol_kernel_v7(); const void* bxg_step_chol_kernel_v8(); const void* bxg_step_chol_kernel_v9(); const void* bxg_step_chol_kernel_v10(); const void* bxg_step_chol_kernel_v11();
}
extern "C" {
const void* bxg_step_lean_kernel_v0(); const void* bxg_step_lean_kernel_v1(); const void* bxg_step_lean_kernel_v2(); const void* bxg_step_lean_kernel_v3(); const void* bxg_step_lean_kernel_v4(); const void* bxg_step_lean_kernel_v5(); const void* bxg_step_lean_kernel_v6(); const void* bxg_step_lean_kernel_v7(); const void* bxg_step_lean_kernel_v8(); const void* bxg_step_lean_kernel_v9(); const void* bxg_step_lean_kernel_v10(); const void* bxg_step_lean_kernel_v11();
const void* bxg_step_chol_lean_kernel_v0(); const void* bxg_step_chol_lean_kernel_v1(); const void* bxg_step_chol_lean_kernel_v2(); const void* bxg_step_chol_lean_kernel_v3(); const void* bxg_step_chol_lean_kernel_v4(); const void* bxg_step_chol_lean_kernel_v5(); const void* bxg_step_chol_lean_kernel_v6(); const void* bxg_step_chol_lean_kernel_v7(); const void* bxg_step_chol_lean_kernel_v8(); const void* bxg_step_chol_lean_kernel_v9(); const void* bxg_step_chol_lean_kernel_v10(); const void* bxg_step_chol_lean_kernel_v11();
}
extern "C" {
const void* bxg_step_perenv_kernel_v0(); const void* bxg_step_perenv_kernel_v1(); const void* bxg_step_perenv_kernel_v3();
const void* bxg_step_perenv_lean_kernel_v0(); const void* bxg_step_perenv_lean_kernel_v1(); const void* bxg_step_perenv_lean_kernel_v3();
const void* bxg_init_perenv_kernel_v0(); const void* bxg_init_perenv_kernel_v1(); const void* bxg_init_perenv_kernel_v3();
}
// per-env models run on variants 0, 1 and 3 (bxg_inst.cu)
static const void* step_perenv_kernel_of(int v, bool lean) {
  switch (v) {
    case 0: return lean ? bxg_step_perenv_lean_kernel_v0() : bxg_step_perenv_kernel_v0();
    case 1: return lean ? bxg_step_perenv_lean_kernel_v1() : bxg_step_perenv_kernel_v1();
    default: return lean ? bxg_step_perenv_lean_kernel_v3() : bxg_step_perenv_kernel_v3();
  }
}
static const void* init_perenv_kernel_of(int v) {
  switch (v) { case 0: return bxg_init_perenv_kernel_v0(); case 1: return bxg_init_perenv_kernel_v1(); default: return bxg_init_perenv_kernel_v3(); }
}
static const void* step_lean_kernel_of(int v, bool chol) {
  switch (v) {
    case 0: return chol ? bxg_step_chol_lean_kernel_v0() : bxg_step_lean_kernel_v0();
    case 1: return chol ? bxg_step_chol_lean_kernel_v1() : bxg_step_lean_kernel_v1();
    case 2: return chol ? bxg_step_chol_lean_kernel_v2() : bxg_step_lean_kernel_v2();
    case 4: return chol ? bxg_step_chol_lean_kernel_v4() : bxg_step_lean_kernel_v4();
    case 5: return chol ? bxg_step_chol_lean_kernel_v5() : bxg_step_lean_kernel_v5();
    case 6: return chol ? bxg_step_chol_lean_kernel_v6() : bxg_step_lean_kernel_v6();
    case 7: return chol ? bxg_step_chol_lean_kernel_v7() : bxg_step_lean_kernel_v7();
    case 8: return chol ? bxg_step_chol_lean_kernel_v8() : bxg_step_lean_kernel_v8();
    case 9: return chol ? bxg_step_chol_lean_kernel_v9() : bxg_step_lean_kernel_v9();
    case 10: return chol ? bxg_step_chol_lean_kernel_v10() : bxg_step_lean_kernel_v10();
    case 11: return chol ? bxg_step_chol_lean_kernel_v11() : bxg_step_lean_kernel_v11();
    default: return chol ? bxg_step_chol_lean_kernel_v3() : bxg_step_lean_kernel_v3();
  }
}
static const void* step_chol_kernel_of(int v) {
  switch (v) { case 0: return bxg_step_chol_kernel_v0(); case 1: return bxg_step_chol_kernel_v1(); case 2: return bxg_step_chol_kernel_v2();
               case 4: return bxg_step_chol_kernel_v4(); case 5: return bxg_step_chol_kernel_v5(); case 6: return bxg_step_chol_kernel_v6(); case 7: return bxg_step_chol_kernel_v7(); case 8: return bxg_step_chol_kernel_v8(); case 9: return bxg_step_chol_kernel_v9(); case 10: return bxg_step_chol_kernel_v10(); case 11: return bxg_step_chol_kernel_v11(); default: return bxg_step_chol_kernel_v3(); }
}
static const void* step_kernel_of(int v) {
  switch (v) { case 0: return bxg_step_kernel_v0(); case 1: return bxg_step_kernel_v1(); case 2: return bxg_step_kernel_v2(); case 4: return bxg_step_kernel_v4(); case 5: return bxg_step_kernel_v5(); case 6: return bxg_step_kernel_v6(); case 7: return bxg_step_kernel_v7(); case 8: return bxg_step_kernel_v8(); case 9: return bxg_step_kernel_v9(); case 10: return bxg_step_kernel_v10(); case 11: return bxg_step_kernel_v11(); default: return bxg_step_kernel_v3(); }
}
static const void* init_kernel_of(int v) {
  switch (v) { case 0: return bxg_init_kernel_v0(); case 1: return bxg_init_kernel_v1(); case 2: return bxg_init_kernel_v2(); case 4: return bxg_init_kernel_v4(); case 5: return bxg_init_kernel_v5(); case 6: return bxg_init_kernel_v6(); case 7: return bxg_init_kernel_v7(); case 8: return bxg_init_kernel_v8(); case 9: return bxg_init_kernel_v9(); case 10: return bxg_init_kernel_v10(); case 11: return bxg_init_kernel_v11(); default: return bxg_init_kernel_v3(); }
}

struct BxgModel {
  bxg::PackedModel pm;
  int kernel_id = 0;       // pm.variant_id, or 10 / 11: that variant compiled for exactly this model's packed layout
  int device = 0;
  int lanes = 32;          // G
  int sm_count = 0;
  int groups = 0;          // envs per CTA
  int threads = 0;         // CTA size
  uint32_t* d_blob = nullptr;
  int64_t n_models = 0;    // 0: one model for every env; n: env e of an n-env batch reads row e of d_blob ([n][model_words])
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
// launches must go to the model's device whatever the caller's current device is
struct DeviceGuard {
  int prev = -1; bool switched = false;
  explicit DeviceGuard(int dev) {
    if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
  }
  ~DeviceGuard() { if (switched) cudaSetDevice(prev); }
};

// BXG_STEP_LEAN: only the leaves that mode touches
bool state_ok_lean(const BxgState* s) {
  return s && s->q && s->qd && s->x_pos && s->x_rot && s->xd_ang && s->xd_vel && s->mass_mx_inv;
}
// Kernel id for a packed model: its variant, or the model-specialised build of that variant (constexpr Dims:
// kernel ids 10 = Ant layout on variant 0, 11 = Humanoid layout on variant 1) when the packed layout, sizes and
// solver settings are IDENTICAL to the ones that build was compiled for.  Any other model keeps the generic kernel.
int specialised_kernel_id(const bxg::PackedModel& pm) {
  if (getenv("BXG_NO_SPECIALISE")) return pm.variant_id;
  auto same = [&](bxg::Dims c) { c.minv_mode = pm.d.minv_mode; c.phase_groups = pm.d.phase_groups; return memcmp(&c, &pm.d, sizeof c) == 0; };   // (phase_groups: a tuning value of the build)
  if (pm.variant_id == 0 && same(bxg::const_dims_ant())) return 10;
  if (pm.variant_id == 1 && same(bxg::const_dims_humanoid())) return 11;
  return pm.variant_id;
}
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

// device side of model creation: upload `blob` (one model, or n_models of them back to back) and configure the kernels
static int model_upload(BxgModel* m, int device, const std::vector<uint32_t>& blob) {
  int ndev = 0;
  cudaError_t ce = cudaGetDeviceCount(&ndev);
  if (ce != cudaSuccess || ndev == 0) { delete m; return fail(BXG_E_CUDA, "no CUDA device: this library has no CPU fallback"); }
  if (device < 0 || device >= ndev) { delete m; return fail(BXG_E_INVALID, "bad device ordinal"); }
  m->device = device;
  const bxg::Dims& D = m->pm.d;
  // half-warp groups when the model fits the 16-lane variant (Ant), else full warps
  m->lanes = bxg::variant(m->pm.variant_id).G;
  int prev = 0;
  cudaGetDevice(&prev);
  auto cleanup = [&](int code) { cudaSetDevice(prev); if (m->d_blob) cudaFree(m->d_blob); delete m; return code; };
  if (cudaSetDevice(device) != cudaSuccess) return cleanup(fail(BXG_E_CUDA, "cudaSetDevice failed"));
  cudaDeviceProp prop;
  if ((ce = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cleanup(cuda_fail(ce, "cudaGetDeviceProperties"));
  m->sm_count = prop.multiProcessorCount;
  const int groups = bxg::envs_per_cta(D, bxg::variant(m->pm.variant_id));
  if (groups < 1) return cleanup(fail(BXG_E_UNSUPPORTED, "one env does not fit the shared memory of an SM"));
  m->groups = groups; m->threads = groups * m->lanes;
  m->smem_bytes = sizeof(uint32_t) * ((size_t)D.model_words + (size_t)groups * D.env_words);
  if (m->smem_bytes > (size_t)prop.sharedMemPerBlockOptin)
    return cleanup(fail(BXG_E_UNSUPPORTED, "model needs more shared memory per CTA than the device offers"));
  if ((ce = cudaMalloc(&m->d_blob, blob.size() * sizeof(uint32_t))) != cudaSuccess) return cleanup(cuda_fail(ce, "cudaMalloc(model)"));
  if ((ce = cudaMemcpy(m->d_blob, blob.data(), blob.size() * sizeof(uint32_t), cudaMemcpyHostToDevice)) != cudaSuccess)
    return cleanup(cuda_fail(ce, "cudaMemcpy(model)"));
  // the attribute is per function (shared by every model of this variant): always
  // raise it to the device maximum, never to this model's own size
  const bool chol = m->pm.d.minv_mode == BXG_MINV_CHOLESKY;
  const void* ks = m->n_models ? step_perenv_kernel_of(m->kernel_id, false) : (chol ? step_chol_kernel_of(m->kernel_id) : step_kernel_of(m->kernel_id));
  const void* ki = m->n_models ? init_perenv_kernel_of(m->kernel_id) : init_kernel_of(m->kernel_id);
  const void* kl = m->n_models ? step_perenv_kernel_of(m->kernel_id, true) : step_lean_kernel_of(m->kernel_id, chol);
  // (dynamic + the kernel's few bytes of static shared memory must fit the opt-in limit)
  for (const void* k : {ks, ki, kl}) {
    cudaFuncAttributes fa;
    if ((ce = cudaFuncGetAttributes(&fa, k)) != cudaSuccess) return cleanup(cuda_fail(ce, "cudaFuncGetAttributes"));
    const int dyn_max = (int)prop.sharedMemPerBlockOptin - (int)fa.sharedSizeBytes;
    if ((int)m->smem_bytes > dyn_max) return cleanup(fail(BXG_E_UNSUPPORTED, "model needs more shared memory per CTA than the device offers"));
    if ((ce = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, dyn_max)) != cudaSuccess) return cleanup(cuda_fail(ce, "cudaFuncSetAttribute(max dynamic shared memory)"));
  }
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m->blocks_per_sm_step, ks, m->threads, m->smem_bytes);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m->blocks_per_sm_init, ki, m->threads, m->smem_bytes);
  if (m->blocks_per_sm_step < 1 || m->blocks_per_sm_init < 1) return cleanup(fail(BXG_E_UNSUPPORTED, "kernel does not fit on an SM"));
  cudaSetDevice(prev);
  return BXG_OK;
}

int bxg_model_create(const BxgModelDesc* desc, int device, BxgModel** out) {
  if (!desc || !out) return fail(BXG_E_INVALID, "null argument");
  *out = nullptr;
  BxgModel* m = new BxgModel();
  int force_variant = -1;
  if (const char* fv = getenv("BXG_FORCE_VARIANT")) force_variant = atoi(fv);   // tuning knob (e.g. 4: Humanoid class on half-warps)
  std::string err = bxg::pack_model(*desc, &m->pm, force_variant);
  if (!err.empty()) { delete m; return fail(BXG_E_UNSUPPORTED, err); }
  m->kernel_id = force_variant >= 0 ? m->pm.variant_id : specialised_kernel_id(m->pm);   // (before the tuning knobs below change Dims)
  if (getenv("BXG_SYNC_LEVEL") || getenv("BXG_PHASE_GROUPS")) m->kernel_id = m->pm.variant_id;
  if (const char* sl = getenv("BXG_SYNC_LEVEL")) m->pm.d.sync_level = atoi(sl);   // tuning knobs
  if (const char* pg = getenv("BXG_PHASE_GROUPS")) m->pm.d.phase_groups = atoi(pg);
  int rc = model_upload(m, device, m->pm.blob);
  if (rc != BXG_OK) return rc;
  *out = m;
  return BXG_OK;
}

// n models of ONE topology (same links, dofs, actuators, contact pairs, solver settings), different constants: env e of
// every n-env batch uses descs[e].  The packed layout (Dims) depends on the topology only and must come out identical.
int bxg_model_create_batched(const BxgModelDesc* descs, int64_t n, int device, BxgModel** out) {
  if (!descs || !out || n < 1) return fail(BXG_E_INVALID, "null argument or n < 1");
  *out = nullptr;
  if (descs[0].minv_mode == BXG_MINV_CHOLESKY) return fail(BXG_E_UNSUPPORTED, "per-env models run the reference's Newton-Schulz mode only");
  BxgModel* m = new BxgModel();
  // the variant the nominal model gets, if it carries the per-env kernels (0, 1, 3), else the generic one
  std::string err = bxg::pack_model(descs[0], &m->pm, -1);
  if (err.empty() && m->pm.variant_id != 0 && m->pm.variant_id != 1 && m->pm.variant_id != 3) err = bxg::pack_model(descs[0], &m->pm, 3);
  if (!err.empty()) { delete m; return fail(BXG_E_UNSUPPORTED, err); }
  const int vid = m->pm.variant_id;
  m->kernel_id = vid; m->n_models = n;
  const size_t words = m->pm.blob.size();
  std::vector<uint32_t> all(words * (size_t)n);
  memcpy(all.data(), m->pm.blob.data(), words * sizeof(uint32_t));
  bxg::PackedModel pe{};
  for (int64_t e = 1; e < n; ++e) {
    err = bxg::pack_model(descs[e], &pe, vid);
    if (!err.empty()) { delete m; return fail(BXG_E_UNSUPPORTED, "model " + std::to_string(e) + ": " + err); }
    if (memcmp(&pe.d, &m->pm.d, sizeof(bxg::Dims)) != 0 || pe.blob.size() != words) {
      delete m;
      return fail(BXG_E_UNSUPPORTED, "model " + std::to_string(e) + " differs from model 0 in topology, sizes, time step, gravity or solver settings: only constants stored in the model blob may vary per env");
    }
    memcpy(all.data() + words * (size_t)e, pe.blob.data(), words * sizeof(uint32_t));
  }
  int rc = model_upload(m, device, all);
  if (rc != BXG_OK) return rc;
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
int bxg_model_kernel_id(const BxgModel* m) { return m ? m->kernel_id : -1; }
int64_t bxg_model_num_models(const BxgModel* m) { return m ? m->n_models : -1; }

int bxg_plan(const BxgModelDesc* desc, int32_t info[8]) {
  if (!desc || !info) return fail(BXG_E_INVALID, "null argument");
  bxg::PackedModel pm;
  std::string err = bxg::pack_model(*desc, &pm);
  if (!err.empty()) return fail(BXG_E_UNSUPPORTED, err);
  const int G = bxg::variant(pm.variant_id).G, groups = bxg::envs_per_cta(pm.d, bxg::variant(pm.variant_id));
  info[0] = pm.variant_id; info[1] = G; info[2] = pm.d.model_words; info[3] = pm.d.env_words; info[4] = groups;
  info[5] = (int32_t)(sizeof(uint32_t) * ((size_t)pm.d.model_words + (size_t)groups * pm.d.env_words));
  info[6] = pm.d.nc; info[7] = specialised_kernel_id(pm);
  return BXG_OK;
}

// Launch shape for a batch.  The kernel is persistent (one CTA per SM, `passes`
// sweeps over the batch).  With the largest CTA that fits an SM, small or awkward
// batch sizes leave SMs idle or pay for a nearly empty last pass, so the number of
// envs per CTA is chosen per call: minimise passes * (cost of one pass), where a
// pass with e of e_max envs resident costs about (e_max + e) -- at least half of a
// full pass is latency that more resident warps hide.
struct LaunchShape { int grid, threads; size_t smem; };
static LaunchShape launch_shape(const BxgModel* m, int64_t n_env) {
  const int per_warp = 32 / m->lanes;
  int best = m->groups; double best_cost = 1e300;
  for (int e = per_warp; e <= m->groups; e += per_warp) {
    int64_t ctas = (n_env + e - 1) / e;
    int64_t grid = ctas < m->sm_count ? ctas : m->sm_count;
    int64_t passes = (n_env + grid * e - 1) / (grid * e);
    double cost = (double)passes * (1.0 * m->groups + e);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = e; }
  }
  int64_t ctas = (n_env + best - 1) / best;
  LaunchShape ls;
  ls.grid = (int)(ctas < m->sm_count ? ctas : m->sm_count);
  ls.threads = best * m->lanes;
  ls.smem = sizeof(uint32_t) * ((size_t)m->pm.d.model_words + (size_t)best * m->pm.d.env_words);
  return ls;
}

int bxg_launch_shape(const BxgModel* m, int64_t n_env, int32_t info[4]) {
  if (!m || !info || n_env <= 0) return fail(BXG_E_INVALID, "null model or empty batch");
  LaunchShape ls = launch_shape(m, n_env);
  info[0] = ls.grid; info[1] = ls.threads; info[2] = (int32_t)ls.smem; info[3] = ls.threads / m->lanes;
  return BXG_OK;
}

int bxg_init(const BxgModel* m, int64_t n_env, const float* q, const float* qd, const BxgState* out, void* stream) {
  if (!m) return fail(BXG_E_INVALID, "null model");
  if (n_env <= 0) return n_env == 0 ? BXG_OK : fail(BXG_E_INVALID, "n_env < 0");
  if (m->n_models && n_env != m->n_models) return fail(BXG_E_INVALID, "a batched model takes exactly one env per model: n_env != number of models");
  if (!q || !qd || !state_ok(out, m->pm.d.nc)) return fail(BXG_E_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  DeviceGuard guard(m->device);
  LaunchShape ls = launch_shape(m, n_env);
  BxgEnvSpec env{}; float* obs = nullptr;
  void* args[] = {(void*)&m->pm.d, (void*)&m->d_blob, (void*)&q, (void*)&qd, (void*)out, (void*)&n_env, (void*)&env, (void*)&obs};
  BXG_CUDA(cudaLaunchKernel(m->n_models ? init_perenv_kernel_of(m->kernel_id) : init_kernel_of(m->kernel_id), dim3(ls.grid), dim3(ls.threads), args, ls.smem, st));
  g_launches.fetch_add(1);
  BXG_CUDA(cudaGetLastError());
  return BXG_OK;
}

int bxg_step(const BxgModel* m, int64_t n_env, int32_t n_frames, const BxgState* in, const float* act, const BxgState* out,
             int32_t flags, const BxgDiag* diag, void* stream) {
  if (!m) return fail(BXG_E_INVALID, "null model");
  if (n_frames < 0) return fail(BXG_E_INVALID, "n_frames < 0");
  if (n_env <= 0) return n_env == 0 ? BXG_OK : fail(BXG_E_INVALID, "n_env < 0");
  if (m->n_models && n_env != m->n_models) return fail(BXG_E_INVALID, "a batched model takes exactly one env per model: n_env != number of models");
  const bool lean = flags & BXG_STEP_LEAN;
  if (lean ? (!state_ok_lean(in) || !state_ok_lean(out)) : (!state_ok(in, m->pm.d.nc) || !state_ok(out, m->pm.d.nc))) return fail(BXG_E_INVALID, "null argument");
  if (m->pm.d.nu > 0 && !act) return fail(BXG_E_INVALID, "act is NULL but the model has actuators");
  BxgDiag dg{nullptr, nullptr, nullptr};
  if ((flags & BXG_STEP_DIAGNOSTICS) && diag) dg = *diag;
  cudaStream_t st = (cudaStream_t)stream;
  DeviceGuard guard(m->device);
  LaunchShape ls = launch_shape(m, n_env);
  int nf = n_frames, fl = flags;
  BxgEnvSpec env{}; BxgEnvIO eio{}; BxgState first{};
  void* args[] = {(void*)&m->pm.d, (void*)&m->d_blob, (void*)in, (void*)&act, (void*)out, (void*)&n_env, (void*)&nf, (void*)&fl, (void*)&dg,
                  (void*)&env, (void*)&eio, (void*)&first};
  const bool chol = m->pm.d.minv_mode == BXG_MINV_CHOLESKY;
  const void* kern = m->n_models ? step_perenv_kernel_of(m->kernel_id, lean) : lean ? step_lean_kernel_of(m->kernel_id, chol) : (chol ? step_chol_kernel_of(m->kernel_id) : step_kernel_of(m->kernel_id));
  BXG_CUDA(cudaLaunchKernel(kern, dim3(ls.grid), dim3(ls.threads), args, ls.smem, st));
  g_launches.fetch_add(1);
  BXG_CUDA(cudaGetLastError());
  return BXG_OK;
}

// what the kernel's env code assumes about the model for each env kind
static const char* env_spec_error(const BxgModel* m, const BxgEnvSpec* spec) {
  const bxg::Dims& d = m->pm.d;
  if (spec->kind < BXG_ENV_ROOT_VELOCITY || spec->kind > BXG_ENV_PUSHER) return "unknown env kind";
  if (spec->kind == BXG_ENV_PUSHER && (d.nu > d.nq || d.nu > d.nv || spec->tip_link < 0 || spec->tip_link >= d.L || spec->target_link < 0 || spec->target_link >= d.L || spec->object_link < 0 || spec->object_link >= d.L)) return "pusher env kind needs tip, object and target links";
  if (spec->kind == BXG_ENV_CARTPOLE && d.nq < 2) return "cartpole env kind needs q = [x, angle, ...]";
  if (spec->kind == BXG_ENV_DOUBLE_CARTPOLE && (d.nv < 3 || spec->tip_link < 0 || spec->tip_link >= d.L)) return "double cartpole env kind needs nv >= 3 and a tip link";
  if (spec->kind == BXG_ENV_REACHER && (d.nq < 2 || spec->tip_link < 0 || spec->tip_link >= d.L || spec->target_link < 0 || spec->target_link >= d.L)) return "reacher env kind needs a tip link and a target link";
  if (spec->kind == BXG_ENV_SWIMMER && d.nq < 2) return "swimmer env kind needs q = [x, y, ...]";
  return nullptr;
}

int bxg_env_obs_size(const BxgModel* m, const BxgEnvSpec* spec) {
  if (!m || !spec) return -1;
  return bxg::env_obs_size(m->pm.d, *spec);
}

int bxg_env_reset(const BxgModel* m, const BxgEnvSpec* spec, int64_t n_env, const float* q, const float* qd,
                  const BxgState* out, float* obs, void* stream) {
  if (!m || !spec) return fail(BXG_E_INVALID, "null argument");
  if (const char* why = env_spec_error(m, spec)) return fail(BXG_E_INVALID, why);
  if (n_env <= 0) return n_env == 0 ? BXG_OK : fail(BXG_E_INVALID, "n_env < 0");
  if (m->n_models && n_env != m->n_models) return fail(BXG_E_INVALID, "a batched model takes exactly one env per model: n_env != number of models");
  if (!q || !qd || !obs || !state_ok(out, m->pm.d.nc)) return fail(BXG_E_INVALID, "null argument");
  cudaStream_t st = (cudaStream_t)stream;
  DeviceGuard guard(m->device);
  LaunchShape ls = launch_shape(m, n_env);
  BxgEnvSpec env = *spec;
  void* args[] = {(void*)&m->pm.d, (void*)&m->d_blob, (void*)&q, (void*)&qd, (void*)out, (void*)&n_env, (void*)&env, (void*)&obs};
  BXG_CUDA(cudaLaunchKernel(m->n_models ? init_perenv_kernel_of(m->kernel_id) : init_kernel_of(m->kernel_id), dim3(ls.grid), dim3(ls.threads), args, ls.smem, st));
  g_launches.fetch_add(1);
  BXG_CUDA(cudaGetLastError());
  return BXG_OK;
}

int bxg_env_step(const BxgModel* m, const BxgEnvSpec* spec, int64_t n_env, int32_t n_frames, const BxgState* in,
                 const float* action, const BxgState* out, const BxgEnvIO* io, void* stream) {
  if (!m || !spec || !io) return fail(BXG_E_INVALID, "null argument");
  if (const char* why = env_spec_error(m, spec)) return fail(BXG_E_INVALID, why);
  if (spec->obs_skip < 0 || spec->obs_skip > m->pm.d.nq || !(spec->env_dt > 0.f)) return fail(BXG_E_INVALID, "bad env spec");
  if (spec->kind == BXG_ENV_PLANAR && m->pm.d.nq < 3) return fail(BXG_E_INVALID, "planar env kind needs q = [x, z, angle, ...]");
  if (n_frames < 1) return fail(BXG_E_INVALID, "n_frames < 1");
  if (n_env <= 0) return n_env == 0 ? BXG_OK : fail(BXG_E_INVALID, "n_env < 0");
  if (m->n_models && n_env != m->n_models) return fail(BXG_E_INVALID, "a batched model takes exactly one env per model: n_env != number of models");
  const bool lean = io->flags & BXG_STEP_LEAN;
  if (lean ? (!state_ok_lean(in) || !state_ok_lean(out)) : (!state_ok(in, m->pm.d.nc) || !state_ok(out, m->pm.d.nc))) return fail(BXG_E_INVALID, "null state leaf");
  if (!io->obs || !io->reward || !io->done || !io->metrics) return fail(BXG_E_INVALID, "null env output");
  if (io->first_state && (!(lean ? state_ok_lean(io->first_state) : state_ok(io->first_state, m->pm.d.nc)) || !io->first_obs)) return fail(BXG_E_INVALID, "null first_state leaf");
  if (m->pm.d.nu > 0 && !action) return fail(BXG_E_INVALID, "action is NULL but the model has actuators");
  cudaStream_t st = (cudaStream_t)stream;
  DeviceGuard guard(m->device);
  LaunchShape ls = launch_shape(m, n_env);
  int nf = n_frames, fl = io->flags & BXG_STEP_LEAN;
  BxgDiag dg{nullptr, nullptr, nullptr};
  BxgEnvSpec env = *spec; BxgEnvIO eio = *io; BxgState first{};
  if (io->first_state) first = *io->first_state;
  void* args[] = {(void*)&m->pm.d, (void*)&m->d_blob, (void*)in, (void*)&action, (void*)out, (void*)&n_env, (void*)&nf, (void*)&fl, (void*)&dg,
                  (void*)&env, (void*)&eio, (void*)&first};
  const bool chol = m->pm.d.minv_mode == BXG_MINV_CHOLESKY;
  const void* kern = m->n_models ? step_perenv_kernel_of(m->kernel_id, lean) : lean ? step_lean_kernel_of(m->kernel_id, chol) : (chol ? step_chol_kernel_of(m->kernel_id) : step_kernel_of(m->kernel_id));
  BXG_CUDA(cudaLaunchKernel(kern, dim3(ls.grid), dim3(ls.threads), args, ls.smem, st));
  g_launches.fetch_add(1);
  BXG_CUDA(cudaGetLastError());
  return BXG_OK;
}

}  // extern "C"
