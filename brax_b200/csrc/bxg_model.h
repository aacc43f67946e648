// bxg_model.h -- host-side packing of BxgModelDesc into the flat "model blob"
// the kernel stages into shared memory, plus the per-env shared-memory layout.
//
// This is the compile-time half of the reference's `scan.tree` /
// `scan.link_types` (brax/scan.py:53-193): tree levels, child lists, ancestor
// masks and dof->link maps are derived once per model here, so the kernel never
// walks the tree by pointer chasing.
#pragma once
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include "../../include/bxg.h"

namespace bxg {

// All offsets are in 4-byte words.  m_* index the model blob, s_* index the
// per-env shared-memory slab.
struct Dims {
  int L, nq, nv, nu, ncon, nlim, nc;
  int nvw, ncw;   // nv / nc rounded up to the register-row widths the kernels are built for
  int nvp;        // row stride (floats) of M, Minv and the Newton-Schulz buffer (see mat_stride)
  int jld;        // row stride (floats) of J: nvw (rows 16-byte aligned)
  int ncp;        // row stride (floats) of A: ncw
  int max_depth;  // deepest tree level
  int solver_iterations, solver_maxls, ns_iters, minv_mode;
  int force_generic;  // tests only: bypass the register-row kernels
  int phase_groups;   // warps of a CTA align their phases in this many independent groups
  int sync_level;     // bit mask of CTA-wide phase alignments per substep: 1 after constraint.force, 2 after
                      // Newton-Schulz, 4 before dynamics, 8 before mass.matrix, 16 before N-S, 32 before constraint.force
  float dt, gx, gy, gz;
  int fluid;          // sys.enable_fluid: fluid forces in dynamics._passive (generic variant only)
  // ---- model blob ----
  int m_link_parent, m_link_ndof, m_link_qadr, m_link_dadr, m_link_depth, m_link_root;
  int m_child_start, m_child_list;       // CSR, children in DEscending index order
  int m_dof_link, m_dof_qidx;            // dof -> link, dof -> q index
  int m_mm_pairs, n_mm_pairs;            // (i | j << 8) for dofs j <= i whose link is ancestor-or-self of dof i's link: the non-zeros of the lower triangle of M
  int m_dof_act_start, m_dof_act_list;   // CSR actuators per dof
  int m_lim_dof;                         // [nlim] dof of each limit row
  int m_tf_pos, m_tf_rot, m_joint_pos, m_in_pos, m_in_rot, m_in_i, m_in_mass, m_link_invw;
  int m_dof_ang, m_dof_vel, m_arm, m_stiff, m_damp, m_lim_lo, m_lim_hi, m_dof_invw, m_dof_sp;
  int m_act_qid, m_act_did, m_act_gain, m_act_gear, m_act_clo, m_act_chi, m_act_flo, m_act_fhi, m_act_bq, m_act_bqd;
  int m_con_lb, m_con_ppos, m_con_frame, m_con_spos, m_con_rad, m_con_mu, m_con_sp;  // sp: [ncon, kImpStride]
  int m_con_kind, m_con_gquat, m_con_half;   // plane-capsule end points: kind 1, geom quaternion [ncon,4], signed half length
  int m_con_anc_lo, m_con_anc_hi;        // [ncon] bitmask of dofs that move link_b
  int m_fluid;                           // [L, 8] per-link fluid constants (pack_fluid), only when `fluid`
  // 80-row variant only: impedance rows stored once per distinct parameter set (most models have one or two),
  // m_dof_sp_idx [nv] / m_con_sp_idx [ncon] pick the row: a smaller blob means one more env per SM there
  int sp_dedup, m_dof_sp_idx, m_con_sp_idx;
  int two_body;                          // some contact has kind 2 (capsule-capsule, link_a may move)
  int m_con_la, m_con_apos, m_con_aquat, m_con_ahalf, m_con_arad;   // kind 2: link and shape of the first capsule
  int m_con_anca_lo, m_con_anca_hi;      // [ncon] bitmask of dofs that move link_a
  int model_words;
  // ---- per-env slab ----
  int s_q, s_qd, s_act, s_tau, s_qfs, s_qfc, s_qdd;
  int s_x_pos, s_x_rot, s_xd_ang, s_xd_vel, s_root_com;
  int s_cinr_pos, s_cinr_rot, s_cinr_i, s_cinr_mass;
  int s_cd_ang, s_cd_vel, s_cdof_ang, s_cdof_vel, s_cdofd_ang, s_cdofd_vel;
  int s_t_ang, s_t_vel;        // [L,3] x2 temps: cdd (RNE) / xi_pos (com)
  int s_f_ang, s_f_vel;        // [L,3] x2 temps: cfrc (RNE) / joint frame pos
  int s_j_rot;                 // [L,4] joint frame rot
  int s_crb_pos, s_crb_i, s_crb_mass;   // alias the t/f temporaries (different phases)
  // Matrices.  In the specialised variants Minv has its own slot and two more are
  // time-shared:   slot0 {M | A}   slot1 {Newton-Schulz ping-pong buffer | J}
  // (M lives from mass.matrix to the end of Newton-Schulz, and after the last
  // substep until store_env; the Newton-Schulz buffer holds I + r and the candidate
  // in turn and trades places with Minv's slot when a candidate is accepted; J lives
  // from constraint.jacobian to the end of constraint.force of the NEXT substep; A
  // inside constraint.force).  The generic variant keeps them apart.
  int s_M, s_Minv, s_Xn, s_B, s_J, s_A, s_JM;
  int s_scr;                   // generic path / Cholesky scratch (2 matrices)
  int s_diag, s_aref, s_b, s_px, s_py, s_pg, s_pres, s_pxn;  // solver vectors (b..pxn) alias the t/f temporaries
  int s_dist;                  // [ncon]
  int s_rowact;                // [nc] BYTES: 1 where the constraint row is active (else J, diag, aref of the row are all zero)
  int s_red;                   // [4] scalars (the env's pre-step reference point)
  int env_words;
};

// Tuning switches (A/B builds, tools/build_alt.py).  XD_TAIL: link velocities once per launch (finish_env), their
// slab words aliased to dead temporaries.  KIN_REGS: the joint transform of kinematics.forward stays in registers.
#ifndef BXG_XD_TAIL
#define BXG_XD_TAIL 1
#endif
#ifndef BXG_DEFAULT_SYNC_LEVEL
#define BXG_DEFAULT_SYNC_LEVEL 1   // one CTA-wide phase alignment per substep, behind constraint.force (Dims::sync_level)
#endif
#ifndef BXG_DEFAULT_PHASE_GROUPS
#define BXG_DEFAULT_PHASE_GROUPS 1 // the warps of a CTA align their phases in this many independent groups (Dims::phase_groups)
#endif
#ifndef BXG_KIN_REGS
#define BXG_KIN_REGS 0
#endif

// Row widths the register-row kernels are instantiated for (bxg_core.cuh).
// Row stride: a multiple of 4 floats (16-byte rows for 128-bit shared loads) whose
// chunk count is odd, so that lanes reading "their own row" with 128-bit
// accesses fall into distinct bank groups.
inline int row_stride(int w) { int c = w / 4; return 4 * (c | 1); }

// Kernel variants that are compiled (bxg_kernels.cu).  A model takes the first
// variant it fits; the last one is the generic any-size kernel.
struct Variant { int G, VC4, NC4, max_links, max_nv, max_nc; };
constexpr int kNumVariantsAll = 10;
// order in which a model is offered to the variants (first fit); 3 is the generic kernel, 4 is forced only
constexpr int kAutoOrder[] = {7, 8, 9, 0, 1, 2, 5, 6, 3};
inline Variant variant(int id) {
  switch (id) {
    case 0: return {16, 4, 6, 16, 16, 24};   // Ant class: half-warp per env
    case 1: return {32, 6, 7, 32, 24, 28};   // Humanoid class: warp per env
    case 2: return {32, 8, 8, 32, 32, 32};
    case 4: return {16, 6, 7, 16, 24, 28};   // Humanoid class on a half-warp (6x6 tiles); forced only
    case 5: return {32, 4, 16, 32, 16, 64};  // few dofs, many constraint rows (Walker2d, HalfCheetah): rows of A stay in shared memory
    case 9: return {4, 2, 4, 4, 8, 16};      // four links with contacts (Hopper: 6 dofs, 14 rows)
    case 8: return {4, 2, 2, 4, 8, 8};       // as 7 with 8-wide matrices; carries the fluid forces (Swimmer: 3 links, 5 dofs)
    case 7: return {4, 1, 1, 4, 4, 4};       // classic-control size (pendulums, Reacher): 4 lanes per env, eight envs per warp
    case 6: return {32, 6, 20, 32, 24, 80};  // Humanoid-size tree with up to 80 constraint rows (HumanoidStandup: 15 contacts); 128-bit active mask
    default: return {32, 0, 0, 32, 64, 128}; // generic
  }
}
// Row stride of the nv-column matrices.  The register-tile products read TM rows per
// lane row-group: the stride must keep the row-groups of one 128-bit load phase (8
// lanes) in distinct banks.  24 floats does for the 24-wide variants (rows 3 or 6
// apart: 72 / 144 floats); the 16- and 32-wide ones need the odd chunk count.
inline int mat_stride(const Variant& v, int w) { return v.VC4 == 6 ? w : row_stride(w); }
// Largest CTA a variant is launched with (its __launch_bounds__): half-warp variants
// fill an SM's shared memory with fewer threads and can keep more registers each.
#ifndef BXG_G32_MAXT
#define BXG_G32_MAXT 640
#endif
#ifndef BXG_G16_MAXT
#define BXG_G16_MAXT 512   // 32 Ant envs per SM (16 warps: four per scheduler) at 128 registers per thread
#endif
// (the 80-row variant holds few envs per SM: small CTAs, so each thread may keep up to 255 registers)
constexpr int variant_max_threads(int G, int VC4, int NC4 = 0) { return G == 16 ? (VC4 == 6 ? 320 : BXG_G16_MAXT) : (NC4 >= 20 ? 256 : BXG_G32_MAXT); }
inline bool variant_fits(const Variant& v, int L, int nv, int nc) { return L <= v.max_links && nv <= v.max_nv && nc <= v.max_nc; }

// blob: the 32-bit words the kernel stages into shared memory (integers and float bit patterns).
// blob_r: only for R != float (the host emulator's double instantiation, tests/simt/): the same
// words as values of type R, index for index; the integer view is then `blob`.
template <class R>
struct PackedModelT {
  Dims d;
  int variant_id = -1;
  std::vector<uint32_t> blob;
  std::vector<R> blob_r;
};
using PackedModel = PackedModelT<float>;

// Impedance parameters of one constraint row as the kernel reads them: from
// [timeconst, dampratio, dmin, dmax, width, mid, power] (solref ++ solimp) to
// [dmin, dmax, width, mid, power, 1/mid^(p-1), 1/(1-mid)^(p-1), b, k].  The last four are the
// row-constant subexpressions of constraint._imp_aref (constraint.py:40-61), evaluated here in
// float with the same operations the oracle applies, instead of four divisions per call.
constexpr int kImpStride = 9;
inline float m_pow(float a, float b) { return powf(a, b); }
inline double m_pow(double a, double b) { return pow(a, b); }
inline float m_sqrt(float a) { return sqrtf(a); }
inline double m_sqrt(double a) { return sqrt(a); }
template <class R>
inline void pack_impedance(const float* p7, R* o9) {
  const R tc = p7[0], dr = p7[1], dmin = p7[2], dmax = p7[3], width = p7[4], mid = p7[5], power = p7[6];
  o9[0] = dmin; o9[1] = dmax; o9[2] = width; o9[3] = mid; o9[4] = power;
  o9[5] = R(1) / m_pow(mid, power - R(1));
  o9[6] = R(1) / m_pow(R(1) - mid, power - R(1));
  R b = R(2) / (dmax * tc);
  R k = R(1) / (dmax * dmax * tc * tc * dr * dr);
  if (dr <= R(0)) b = -dr / dmax;
  if (tc <= R(0)) k = -tc / (dmax * dmax);
  o9[7] = b; o9[8] = k;
}

// Fluid constants of one link (brax/fluid.py:24-53,73-77), evaluated in float with the same
// operations the oracle applies: [ang_scale, vel_scale, cv[3], ca[3]] with
// frc.ang = ang_scale * w + ca * |w| * w / 64,  frc.vel = vel_scale * v + cv * |v| * v
constexpr int kFluidStride = 8;
template <class R>
inline void pack_fluid(const float* inertia_i9, R mass, R viscosity, R density, R* o8) {
  const R dg[3] = {inertia_i9[0], inertia_i9[4], inertia_i9[8]};
  R box[3];
  for (int i = 0; i < 3; ++i) {
    R sum = R(0);
    for (int j = 0; j < 3; ++j) sum += dg[j] * (i == j ? R(-1) : R(1));
    sum = R(6) * (sum > R(1e-12) ? sum : R(1e-12));
    box[i] = m_sqrt(sum / mass);
  }
  const R pi = R(3.14159265358979323846);
  const R diam = (box[0] + box[1] + box[2]) / R(3);
  o8[0] = -pi * (diam * diam * diam) * viscosity;
  o8[1] = (R)(-3.0 * 3.14159265358979323846) * diam * viscosity;
  const R bmv[3] = {box[1] * box[2], box[0] * box[2], box[0] * box[1]};
  const R p2[3] = {box[0] * box[0], box[1] * box[1], box[2] * box[2]};
  const R p4[3] = {p2[0] * p2[0], p2[1] * p2[1], p2[2] * p2[2]};
  const R bma[3] = {box[0] * (p4[1] + p4[2]), box[1] * (p4[0] + p4[2]), box[2] * (p4[0] + p4[1])};
  for (int i = 0; i < 3; ++i) { o8[2 + i] = R(-0.5) * density * bmv[i]; o8[5 + i] = R(-1.0) * density * bma[i]; }
}

// distinct rows (bitwise) of an [n, kImpStride] table and the row each entry maps to
template <class R>
inline void dedup_rows(const std::vector<R>& rows, int n, std::vector<R>* uniq, std::vector<int>* idx) {
  uniq->clear(); idx->assign(n, 0);
  for (int i = 0; i < n; ++i) {
    int found = -1;
    for (int u = 0; u < (int)uniq->size() / kImpStride && found < 0; ++u)
      if (memcmp(uniq->data() + u * kImpStride, rows.data() + i * kImpStride, sizeof(R) * kImpStride) == 0) found = u;
    if (found < 0) { found = (int)uniq->size() / kImpStride; uniq->insert(uniq->end(), rows.begin() + i * kImpStride, rows.begin() + (i + 1) * kImpStride); }
    (*idx)[i] = found;
  }
}

// Returns empty string on success, else an error message.
template <class R>
inline std::string pack_model_t(const BxgModelDesc& m, PackedModelT<R>* out, int force_variant = -1) {
  Dims& d = out->d;
  std::vector<uint32_t>& b = out->blob;
  std::vector<R>& br = out->blob_r;
  constexpr bool kWide = sizeof(R) != sizeof(float);
  b.clear(); br.clear();
  if (m.abi_version != BXG_ABI_VERSION) return "abi_version mismatch";
  if (m.num_links < 1 || m.nv < 1 || m.nq < 1) return "empty model";
  if (m.num_links > 32) return "num_links > 32 not supported";
  if (m.nv > 64) return "nv > 64 not supported";
  const int L = m.num_links;
  d.L = L; d.nq = m.nq; d.nv = m.nv; d.nu = m.nu; d.ncon = m.ncon;
  int nonfree = 0, nq = 0, nv = 0;
  std::vector<int> qadr(L), dadr(L), depth(L), root(L);
  for (int l = 0; l < L; ++l) {
    int nd = m.link_ndof[l];
    if (nd < 0 || nd > 3) return "link_ndof must be 0 (free) or 1..3";
    int p = m.link_parent[l];
    if (p >= l || p < -1) return "link_parents must be depth-first ordered";
    if (nd == 0 && p != -1) return "free joints must be roots";
    qadr[l] = nq; dadr[l] = nv;
    nq += nd == 0 ? 7 : nd; nv += nd == 0 ? 6 : nd;
    nonfree += nd;
    depth[l] = p < 0 ? 0 : depth[p] + 1;
    root[l] = p < 0 ? l : root[p];
  }
  if (nq != m.nq || nv != m.nv) return "nq/nv inconsistent with link_ndof";
  d.nlim = m.has_limit ? nonfree : 0;
  d.nc = 4 * m.ncon + d.nlim;
  if (d.nc > 128) return "more than 128 constraint rows not supported";
  int vid = force_variant;
  d.fluid = m.enable_fluid ? 1 : 0;
  // fluid forces are compiled into the small 8-wide variants (8, 9) and the generic one; the other specialised variants stay exactly as profiled
  auto fluid_ok = [](int k) { return variant(k).VC4 == 0 || (variant(k).G == 4 && variant(k).VC4 == 2); };
  d.two_body = 0;
  for (int c = 0; c < m.ncon; ++c) if (m.con_kind && m.con_kind[c] == BXG_CON_CAPSULE_CAPSULE) d.two_body = 1;
  // two-body contacts are compiled into variant 5 and the generic one only
  auto two_body_ok = [](int k) { return variant(k).VC4 == 0 || variant(k).NC4 == 16; };
  // a variant is allowed when it carries every optional code path the model needs; a model is offered to the
  // allowed variants in kAutoOrder (first fit).  Only an explicitly forced variant can be rejected here.
  auto allowed = [&](int k) { return (!d.fluid || fluid_ok(k)) && (!d.two_body || two_body_ok(k)); };
  if (vid >= 0 && vid < kNumVariantsAll && d.fluid && !fluid_ok(vid)) return "fluid forces are compiled into kernel variants 3, 8 and 9 only";
  if (vid >= 0 && vid < kNumVariantsAll && d.two_body && !two_body_ok(vid)) return "capsule-capsule contacts are compiled into kernel variants 3 and 5 only";
  if (vid < 0) {
    for (int k : kAutoOrder) { if (!allowed(k)) continue; vid = k; if (variant_fits(variant(k), L, m.nv, d.nc)) break; }
  }
  if (vid >= kNumVariantsAll || !variant_fits(variant(vid), L, m.nv, d.nc)) return "model does not fit the requested kernel variant";
  out->variant_id = vid;
  const Variant var = variant(vid);
  d.nvw = var.VC4 ? 4 * var.VC4 : (m.nv + 3) & ~3;
  d.ncw = var.NC4 ? 4 * var.NC4 : ((d.nc > 0 ? d.nc : 1) + 3) & ~3;
  d.nvp = mat_stride(var, d.nvw); d.jld = d.nvw; d.ncp = d.ncw;
  // wide variants keep the rows of A in shared memory only (bxg_core.cuh con_force_rows AREG): their stride
  // follows the model, not the compiled width
  if (var.NC4 > 0 && ((4 * var.NC4 + var.G - 1) / var.G) * 4 * var.NC4 > 56) d.ncp = ((d.nc > 0 ? d.nc : 1) + 3) & ~3;
  d.max_depth = 0;
  for (int l = 0; l < L; ++l) d.max_depth = depth[l] > d.max_depth ? depth[l] : d.max_depth;
  d.solver_iterations = m.solver_iterations; d.solver_maxls = m.solver_maxls;
  d.ns_iters = m.matrix_inv_iterations; d.minv_mode = m.minv_mode; d.force_generic = 0; d.sync_level = BXG_DEFAULT_SYNC_LEVEL; d.phase_groups = BXG_DEFAULT_PHASE_GROUPS;
  d.dt = m.dt; d.gx = m.gravity[0]; d.gy = m.gravity[1]; d.gz = m.gravity[2];

  auto put_i = [&](const std::vector<int>& v) { int o = (int)b.size(); for (int x : v) { b.push_back((uint32_t)x); if (kWide) br.push_back(R(0)); } return o; };
  auto put_f = [&](const float* p, int n) {   // model constants: float32 values (exact in R)
    int o = (int)b.size();
    for (int i = 0; i < n; ++i) { union { float f; uint32_t u; } c; c.f = p ? p[i] : 0.f; b.push_back(c.u); if (kWide) br.push_back((R)c.f); }
    return o;
  };
  auto put_r = [&](const R* p, int n) {       // constants derived on the host in the scalar type R
    int o = (int)b.size();
    for (int i = 0; i < n; ++i) { union { float f; uint32_t u; } c; c.f = (float)p[i]; b.push_back(c.u); if (kWide) br.push_back(p[i]); }
    return o;
  };
  auto put_ip = [&](const int32_t* p, int n) { int o = (int)b.size(); for (int i = 0; i < n; ++i) { b.push_back((uint32_t)p[i]); if (kWide) br.push_back(R(0)); } return o; };

  d.m_link_parent = put_ip(m.link_parent, L);
  d.m_link_ndof = put_ip(m.link_ndof, L);
  d.m_link_qadr = put_i(qadr); d.m_link_dadr = put_i(dadr);
  d.m_link_depth = put_i(depth); d.m_link_root = put_i(root);
  std::vector<int> cstart(L + 1, 0), clist;
  for (int l = 0; l < L; ++l) {
    cstart[l] = (int)clist.size();
    for (int c = L - 1; c > l; --c) if (m.link_parent[c] == l) clist.push_back(c);
  }
  cstart[L] = (int)clist.size();
  if (clist.empty()) clist.push_back(0);
  d.m_child_start = put_i(cstart); d.m_child_list = put_i(clist);

  std::vector<int> dof_link(m.nv), dof_q(m.nv);
  for (int l = 0; l < L; ++l) {
    int w = m.link_ndof[l] == 0 ? 6 : m.link_ndof[l];
    for (int k = 0; k < w; ++k) { dof_link[dadr[l] + k] = l; dof_q[dadr[l] + k] = m.link_ndof[l] == 0 ? -1 : qadr[l] + k; }
  }
  d.m_dof_link = put_i(dof_link); d.m_dof_qidx = put_i(dof_q);
  auto is_anc = [&](int anc, int l) { while (l >= 0) { if (l == anc) return true; l = m.link_parent[l]; } return false; };
  std::vector<int> pairs;
  for (int i = 0; i < m.nv; ++i)
    for (int j = 0; j <= i; ++j) if (is_anc(dof_link[j], dof_link[i])) pairs.push_back(i | (j << 8));
  d.n_mm_pairs = (int)pairs.size();
  d.m_mm_pairs = put_i(pairs);
  std::vector<int> astart(m.nv + 1, 0), alist;
  for (int dd = 0; dd < m.nv; ++dd) {
    astart[dd] = (int)alist.size();
    for (int a = 0; a < m.nu; ++a) {
      if (m.act_qd_id[a] < 0 || m.act_qd_id[a] >= m.nv || m.act_q_id[a] < 0 || m.act_q_id[a] >= m.nq) return "actuator index out of range";
      if (m.act_qd_id[a] == dd) alist.push_back(a);
    }
  }
  astart[m.nv] = (int)alist.size();
  if (alist.empty()) alist.push_back(0);
  d.m_dof_act_start = put_i(astart); d.m_dof_act_list = put_i(alist);
  std::vector<int> lim_dof;
  for (int i = 0; i < m.nv; ++i) if (dof_q[i] >= 0) lim_dof.push_back(i);
  if (lim_dof.empty()) lim_dof.push_back(0);
  d.m_lim_dof = put_i(lim_dof);

  d.m_tf_pos = put_f(m.link_tf_pos, L * 3); d.m_tf_rot = put_f(m.link_tf_rot, L * 4);
  d.m_joint_pos = put_f(m.link_joint_pos, L * 3);
  d.m_in_pos = put_f(m.inertia_pos, L * 3); d.m_in_rot = put_f(m.inertia_rot, L * 4);
  d.m_in_i = put_f(m.inertia_i, L * 9); d.m_in_mass = put_f(m.inertia_mass, L);
  d.m_link_invw = put_f(m.link_invweight, L);
  d.m_dof_ang = put_f(m.dof_ang, m.nv * 3); d.m_dof_vel = put_f(m.dof_vel, m.nv * 3);
  d.m_arm = put_f(m.dof_armature, m.nv); d.m_stiff = put_f(m.dof_stiffness, m.nv);
  d.m_damp = put_f(m.dof_damping, m.nv);
  d.m_lim_lo = put_f(m.has_limit ? m.dof_limit_lo : nullptr, m.nv);
  d.m_lim_hi = put_f(m.has_limit ? m.dof_limit_hi : nullptr, m.nv);
  d.m_dof_invw = put_f(m.dof_invweight, m.nv); {
    std::vector<R> dsp(m.nv * kImpStride, R(0));
    for (int i = 0; i < m.nv; ++i) pack_impedance<R>(m.dof_solver_params + 7 * i, dsp.data() + kImpStride * i);
    d.sp_dedup = var.NC4 >= 20 ? 1 : 0;
    d.m_dof_sp_idx = d.m_con_sp_idx = (int)b.size();
    if (d.sp_dedup) {
      std::vector<int> idx; std::vector<R> uniq;
      dedup_rows<R>(dsp, m.nv, &uniq, &idx);
      d.m_dof_sp_idx = put_i(idx);
      d.m_dof_sp = put_r(uniq.data(), (int)uniq.size());
    } else {
      d.m_dof_sp = put_r(dsp.data(), m.nv * kImpStride);
    }
  }
  int nu1 = m.nu > 0 ? m.nu : 0;
  d.m_act_qid = put_ip(m.act_q_id, nu1); d.m_act_did = put_ip(m.act_qd_id, nu1);
  d.m_act_gain = put_f(m.act_gain, nu1); d.m_act_gear = put_f(m.act_gear, nu1);
  d.m_act_clo = put_f(m.act_ctrl_lo, nu1); d.m_act_chi = put_f(m.act_ctrl_hi, nu1);
  d.m_act_flo = put_f(m.act_force_lo, nu1); d.m_act_fhi = put_f(m.act_force_hi, nu1);
  d.m_act_bq = put_f(m.act_bias_q, nu1); d.m_act_bqd = put_f(m.act_bias_qd, nu1);
  for (int c = 0; c < m.ncon; ++c) {
    const bool cc2 = m.con_kind && m.con_kind[c] == BXG_CON_CAPSULE_CAPSULE;
    if (m.con_link_b[c] < 0 || m.con_link_b[c] >= L) return "con_link_b out of range";
    if (!cc2 && m.con_link_a[c] != -1) return "plane must be attached to the world (link_a == -1)";
    if (cc2 && (m.con_link_a[c] < -1 || m.con_link_a[c] >= L)) return "con_link_a out of range";
    if (cc2 && (!m.con_a_pos || !m.con_a_quat || !m.con_a_half || !m.con_a_radius || !m.con_geom_quat || !m.con_half_len)) return "capsule-capsule contact without its capsule shapes";
  }
  d.m_con_lb = put_ip(m.con_link_b, m.ncon);
  d.m_con_ppos = put_f(m.con_plane_pos, m.ncon * 3); d.m_con_frame = put_f(m.con_frame, m.ncon * 9);
  d.m_con_spos = put_f(m.con_sphere_pos, m.ncon * 3); d.m_con_rad = put_f(m.con_radius, m.ncon);
  d.m_con_mu = put_f(m.con_friction, m.ncon);
  std::vector<float> sp(m.ncon * 7 + 1, 0.f);
  for (int c = 0; c < m.ncon; ++c) {
    sp[c * 7 + 0] = m.con_solref[c * 2]; sp[c * 7 + 1] = m.con_solref[c * 2 + 1];
    for (int k = 0; k < 5; ++k) sp[c * 7 + 2 + k] = m.con_solimp[c * 5 + k];
  }
  {
    std::vector<R> csp(m.ncon * kImpStride + 1, R(0));
    for (int c = 0; c < m.ncon; ++c) pack_impedance<R>(sp.data() + 7 * c, csp.data() + kImpStride * c);
    if (d.sp_dedup && m.ncon > 0) {
      std::vector<int> idx; std::vector<R> uniq;
      csp.resize(m.ncon * kImpStride);
      dedup_rows<R>(csp, m.ncon, &uniq, &idx);
      d.m_con_sp_idx = put_i(idx);
      d.m_con_sp = put_r(uniq.data(), (int)uniq.size());
    } else {
      d.m_con_sp = put_r(csp.data(), m.ncon * kImpStride);
    }
  }
  {
    std::vector<int> kind(m.ncon > 0 ? m.ncon : 1, 0);
    std::vector<float> gq((m.ncon > 0 ? m.ncon : 1) * 4, 0.f), half(m.ncon > 0 ? m.ncon : 1, 0.f);
    for (int c = 0; c < m.ncon; ++c) {
      kind[c] = m.con_kind ? m.con_kind[c] : BXG_CON_PLANE_SPHERE;
      if (kind[c] != BXG_CON_PLANE_SPHERE && kind[c] != BXG_CON_PLANE_CAPSULE_END && kind[c] != BXG_CON_CAPSULE_CAPSULE) return "unknown contact kind";
      if (kind[c] == BXG_CON_PLANE_CAPSULE_END && (!m.con_geom_quat || !m.con_half_len)) return "capsule contact without con_geom_quat / con_half_len";
      gq[4 * c] = 1.f;
      if (kind[c] != BXG_CON_PLANE_SPHERE) { for (int k = 0; k < 4; ++k) gq[4 * c + k] = m.con_geom_quat[4 * c + k]; half[c] = m.con_half_len[c]; }
    }
    d.m_con_kind = put_i(kind); d.m_con_gquat = put_f(gq.data(), (int)gq.size()); d.m_con_half = put_f(half.data(), (int)half.size());
  }
  std::vector<int> clo(m.ncon > 0 ? m.ncon : 1, 0), chi(m.ncon > 0 ? m.ncon : 1, 0);
  for (int c = 0; c < m.ncon; ++c) {
    uint64_t mask = 0;
    for (int j = 0; j < m.nv; ++j) if (is_anc(dof_link[j], m.con_link_b[c])) mask |= (uint64_t)1 << j;
    clo[c] = (int)(uint32_t)(mask & 0xffffffffu); chi[c] = (int)(uint32_t)(mask >> 32);
  }
  d.m_con_anc_lo = put_i(clo); d.m_con_anc_hi = put_i(chi);
  d.m_con_la = d.m_con_apos = d.m_con_aquat = d.m_con_ahalf = d.m_con_arad = d.m_con_anca_lo = d.m_con_anca_hi = (int)b.size();
  if (d.two_body) {
    std::vector<int> la(m.ncon), alo(m.ncon, 0), ahi(m.ncon, 0);
    for (int c = 0; c < m.ncon; ++c) {
      la[c] = m.con_link_a[c];
      uint64_t mask = 0;
      for (int j = 0; j < m.nv; ++j) if (la[c] >= 0 && is_anc(dof_link[j], la[c])) mask |= (uint64_t)1 << j;
      alo[c] = (int)(uint32_t)(mask & 0xffffffffu); ahi[c] = (int)(uint32_t)(mask >> 32);
    }
    d.m_con_la = put_i(la); d.m_con_anca_lo = put_i(alo); d.m_con_anca_hi = put_i(ahi);
    d.m_con_apos = put_f(m.con_a_pos, m.ncon * 3); d.m_con_aquat = put_f(m.con_a_quat, m.ncon * 4);
    d.m_con_ahalf = put_f(m.con_a_half, m.ncon); d.m_con_arad = put_f(m.con_a_radius, m.ncon);
  }
  d.m_fluid = (int)b.size();
  if (d.fluid) {
    std::vector<R> fl(L * kFluidStride, R(0));
    for (int l = 0; l < L; ++l) pack_fluid<R>(m.inertia_i + 9 * l, (R)m.inertia_mass[l], (R)m.viscosity, (R)m.density, fl.data() + kFluidStride * l);
    d.m_fluid = put_r(fl.data(), L * kFluidStride);
  }
  while (b.size() % 4) { b.push_back(0); if (kWide) br.push_back(R(0)); }
  d.model_words = (int)b.size();

  // ---- per-env slab ----
  // Only what is read with 128-bit loads is 16-byte aligned and padded (matrices, qf_smooth,
  // the solver vectors); everything else is packed to its exact size: shared memory per env
  // decides how many envs an SM holds, and throughput follows that number.
  int o = 0;
  auto take = [&](int n) { o = (o + 3) & ~3; int r = o; o += n; return r; };   // aligned start
  auto take1 = [&](int n) { int r = o; o += n; return r; };                     // packed
  const int nvv = m.nv, nc = d.nc, ncz = nc > 0 ? nc : 1;
  d.s_q = take1(m.nq); d.s_qd = take1(nvv); d.s_act = take1(m.nu > 0 ? m.nu : 1);
  d.s_qfc = take1(nvv); d.s_qdd = take1(nvv);
  // (the Cholesky mode shares them too: its factor lives in the Newton-Schulz buffer's slot)
  const bool shared_slots = var.VC4 > 0 && m.matrix_inv_iterations > 0;
  // Everything kinematics.forward and dynamics.transform_com produce (x, xd, root_com, cinr, cd, cdof, cdofd).
  // It is dead from the end of dynamics.forward until kinematics recomputes it after integrate, i.e. for the
  // whole of constraint.force: in the specialised variants the part of A that does not fit M's slot spills
  // over this block, which therefore sits directly behind that slot (see below).
  auto take_kin_block = [&]() {
    d.s_x_pos = take1(L * 3); d.s_x_rot = take1(L * 4);   // (xd: see s_xd_* below)
    if (!BXG_XD_TAIL) { d.s_xd_ang = take1(L * 3); d.s_xd_vel = take1(L * 3); }
    d.s_root_com = take1(L * 3);
    d.s_cinr_pos = take1(L * 3); d.s_cinr_rot = take1(L * 4); d.s_cinr_i = take1(L * 9); d.s_cinr_mass = -1;   // cinr.mass is the model's link mass: never stored per env
    d.s_cd_ang = take1(L * 3); d.s_cd_vel = take1(L * 3);
    d.s_cdof_ang = take1(nvv * 3); d.s_cdof_vel = take1(nvv * 3);
    d.s_cdofd_ang = take1(nvv * 3); d.s_cdofd_vel = take1(nvv * 3);
  };
  if (!shared_slots) take_kin_block();
  d.s_diag = take1(ncz); d.s_aref = take1(ncz);
  d.s_qfs = take(d.nvw);
  // union of phase-local temporaries: kinematics / RNE / com temps, composite
  // inertias (CRBA) and the constraint-solver vectors never live at the same time.
  // tau (actuator.to_tau) lives from the start of dynamics to qf_smooth and, for the
  // COM-kind env, from the epilogue's actuator pass to the observation: it takes the
  // place of the joint-frame rotations, which only kinematics uses.
  {
    o = (o + 3) & ~3;
    int base = o;
    d.s_t_ang = take1(L * 3); d.s_t_vel = take1(L * 3); d.s_f_ang = take1(L * 3); d.s_f_vel = take1(L * 3);
    // The link velocities xd are written once per launch, by finish_env after the last substep, and read only by
    // the env epilogue and store_env: they live in the (then dead) t_ang / t_vel temporaries.
    if (BXG_XD_TAIL) { d.s_xd_ang = d.s_t_ang; d.s_xd_vel = d.s_t_vel; }
    d.s_j_rot = take1(L * 4 > nvv ? L * 4 : nvv);
    d.s_tau = d.s_j_rot;
    int end_tf = o;
    o = base;
    d.s_crb_pos = take1(L * 3); d.s_crb_i = take1(L * 9); d.s_crb_mass = take1(L);
    int end_crb = o;
    o = base;
    d.s_b = take(d.ncw); d.s_px = take(d.ncw); d.s_py = take(d.ncw); d.s_pg = take(d.ncw); d.s_pres = take(d.ncw); d.s_pxn = take(d.ncw);
    int end_pg = o;
    o = end_tf > end_crb ? end_tf : end_crb;
    o = o > end_pg ? o : end_pg;
  }
  // matrices with nv columns keep nvw rows: rows/columns past nv stay zero so the
  // register-tile kernels can run their loops to the compile-time width
  const int mat_v = d.nvw * d.nvp;          // [nvw][nvp]
  const int mat_j = ncz * d.jld;            // J   [nc][jld]
  const int mat_a = ncz * d.ncp;            // A   [nc][ncp]
  auto mx = [](int a, int b) { return a > b ? a : b; };
  d.s_Minv = take(mat_v);
  if (shared_slots) {
    // slot1 {Newton-Schulz buffer | J}, then slot0 {M | A}: M takes mat_v words, A runs on into the kinematics
    // block behind it (dead while A lives), plus padding only if even that is too small
    int s1 = take(mx(mx(mat_v, mat_j), 6 * nvv));   // (mass.matrix parks crb * cdof, 6 floats per dof, in the Newton-Schulz buffer)
    int s0 = take(mat_v);
    take_kin_block();
    if (o - s0 < mat_a) o = s0 + mat_a;
    d.s_M = s0; d.s_A = s0; d.s_Xn = s1; d.s_B = s1; d.s_J = s1;
    d.s_JM = s0; d.s_scr = s1;  // unused by the specialised kernels
  } else {
    d.s_M = take(mat_v); d.s_J = take(mat_j);
    d.s_scr = take(mat_v + mx(mat_v, 6 * nvv));   // Cholesky: dst/Lm; generic Newton-Schulz: candidate, I + r (also mass.matrix's 6 floats per dof)
    d.s_Xn = d.s_scr; d.s_B = d.s_scr + mat_v;
    d.s_A = take(mat_a); d.s_JM = take(mat_j);
  }
  d.s_dist = take1(m.ncon > 0 ? m.ncon : 1);
  d.s_rowact = take1((ncz + 3) / 4);   // one byte per row
  d.s_red = take1(4);
  o = (o + 3) & ~3;
  // half-warp variants put two envs in one warp: offset their slabs by 16 banks so that the
  // two envs' broadcast row loads (64 B each) never share a bank
  if (var.G == 16) while (o % 32 != 16) o += 4;
  d.env_words = o;
  if (getenv("BXG_DEBUG_LAYOUT"))   // tuning aid: where the slab's words go
    fprintf(stderr, "slab: q %d qd %d act %d qfc %d qdd %d diag %d aref %d qfs %d t_ang %d j_rot %d b %d Minv %d slot1 %d slot0 %d x_pos %d cdofd_vel %d dist %d rowact %d red %d end %d (mat_v %d mat_a %d mat_j %d)\n",
            d.s_q, d.s_qd, d.s_act, d.s_qfc, d.s_qdd, d.s_diag, d.s_aref, d.s_qfs, d.s_t_ang, d.s_j_rot, d.s_b, d.s_Minv, d.s_Xn, d.s_M, d.s_x_pos,
            d.s_cdofd_vel, d.s_dist, d.s_rowact, d.s_red, o, mat_v, mat_a, mat_j);
  return "";
}
inline std::string pack_model(const BxgModelDesc& m, PackedModel* out, int force_variant = -1) { return pack_model_t<float>(m, out, force_variant); }

// Length of one observation vector (shared by the kernels and the C ABI).
#if defined(__CUDACC__)
#define BXG_MODEL_HD __host__ __device__ inline
#else
#define BXG_MODEL_HD inline
#endif
BXG_MODEL_HD int env_obs_size(const Dims& D, const BxgEnvSpec& sp) {
  int base = (D.nq - sp.obs_skip) + D.nv;
  if (sp.kind == BXG_ENV_DOUBLE_CARTPOLE) return 1 + 2 * (D.nq - 1) + D.nv;   // q[0], sin, cos of q[1:], qd
  if (sp.kind == BXG_ENV_PUSHER) return 2 * D.nu + 9;                           // q[:nu], qd[:nu], three centres of mass
  if (sp.kind == BXG_ENV_REACHER) return 4 + (D.nq - 2) + 2 + 3;               // cos, sin of q[:2], q[2:], tip_vel[:2], tip - target
  return (sp.kind == BXG_ENV_COM_VELOCITY || sp.kind == BXG_ENV_STANDUP) ? base + 10 * D.L + 6 * D.L + D.nv : base;
}

}  // namespace bxg
