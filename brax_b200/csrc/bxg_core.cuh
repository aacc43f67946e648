// bxg_core.cuh -- the fused generalized physics step, one lane-group per env.
//
// Execution model: G lanes (16 or 32) cooperate on one environment whose whole
// working state lives in a shared-memory slab (layout: bxg_model.h Dims::s_*).
// The algorithm is written ONCE against a tiny executor interface X:
//
//   ex.lanes(f)    run f(lane) on every lane of the group, then group barrier
//   ex.sum/max(p)  group all-reduce over a per-lane value
//
// DevExec (bxg_kernels.cu) maps it to __syncwarp / shuffles on sm_100a.
// HostExec (tests/simt/) runs the lanes as a loop so that kernel LOGIC can be
// checked against the oracle on a machine without a GPU; it is test-only code
// and is never reachable from the product library.
//
// Reference statements restated here (paths relative to the brax tree):
//   pipeline.step        brax/generalized/pipeline.py:64-94      substep()
//   actuator.to_tau      brax/actuator.py:23-57                  dyn_forces()
//   dynamics.forward     brax/generalized/dynamics.py:137-236    dyn_forces()
//   constraint.force     brax/generalized/constraint.py:194-240  con_force()
//   jaxopt PG            (third party; see oracle/bxg_oracle.c)   con_force()
//   integrator.integrate brax/generalized/integrator.py:46-84    integrate()
//   kinematics.forward   brax/kinematics.py:31-108               kinematics()
//   dynamics.transform_com  generalized/dynamics.py:27-134       transform_com()
//   mass.matrix          brax/generalized/mass.py:27-83          mass_matrix()
//   math.inv_approximate brax/math.py:278-305                    minv_newton_schulz()
//   constraint.jacobian  brax/generalized/constraint.py:101-191  con_jacobian()
#pragma once
#include <math.h>
#include <stdint.h>
#include <type_traits>

#include "bxg_model.h"

// Model-specialised translation units (bxg_inst.cu, kernel ids 10 and 11): BXG_CONST_DIMS names a generated header
// (tools/gen_const_dims.py) with the packed Dims of ONE model as a constexpr function BXG_CONST_DIMS_FN, so every
// size, slab offset, blob offset and loop bound of the step is a compile-time constant -- what jit gives the
// reference, whose tree scans unroll at trace time (scan.py:87-134).  The generic kernels read the same fields from
// the kernel parameter; bxg_model_create picks the specialised kernel only when the model's packed Dims are identical.
#if defined(BXG_CONST_DIMS)
#include BXG_CONST_DIMS
#define BXG_GET_DIMS(c) constexpr ::bxg::Dims D = ::bxg::BXG_CONST_DIMS_FN()
#define BXG_DIMS_OF(c) (::bxg::BXG_CONST_DIMS_FN())
#else
#define BXG_GET_DIMS(c) const ::bxg::Dims& D = *(c).D
#define BXG_DIMS_OF(c) (*(c).D)
#endif

#if defined(__CUDACC__)
#define BXG_HD __host__ __device__ __forceinline__
#define BXG_HD_NOINLINE static __host__ __device__ __noinline__   // one copy: keeps the per-substep code footprint small
#else
#define BXG_HD inline
#define BXG_HD_NOINLINE inline
#endif

// k-loop unroll of the register-tile product (in blocks of 4 k-steps).  8 = the whole loop for every tile width.  While the
// products were LSU-bound the unroll factor did not matter (1 / 2 / 3 / 5 within 1 %); with the pair-shared B operand
// (TileBase::PAIR_COLS) the loop overhead and the accumulator hand-over MOVs of the peeled first block show:
// Humanoid 8192 4.45 -> 4.59 / 4.67 / 4.68 M env-steps/s at unroll 3 / 4 / 5 (= full), 512 k 4.86 -> 5.10 M
#ifndef BXG_TILE_UNROLL
#define BXG_TILE_UNROLL 8
#endif
// explicit software pipeline in the tile product of the 16-wide variants: unroll of its k-block loop.  4 = the whole loop
// (W = 16: four blocks; Ant's specialised build: three and the tail).  Rolled, the hand-over of the prefetched operands is
// 41 register MOVs per 48 FFMA2 and the loop control on top: Ant 14.91 -> 15.8 M env-steps/s unrolled (pipelined or not)
#ifndef BXG_PIPE16_UNROLL
#define BXG_PIPE16_UNROLL 4
#endif
// explicit software pipeline in the 6x6 tile product of the half-warp 24-wide variant (320-thread CTAs: registers to spare)
#ifndef BXG_PIPE_3X6
#define BXG_PIPE_3X6 false    // (with the 3x6 tiles at 96 registers the explicit pipeline costs 5 %: profiles/r01_sweep_r1i.json)
#endif
#ifndef BXG_PIPE_4X4
#define BXG_PIPE_4X4 true
#endif
#ifndef BXG_PIPE_6X6
#define BXG_PIPE_6X6 true
#endif
#define BXG_PRAGMA_(x) _Pragma(#x)
#define BXG_PRAGMA_UNROLL(n) BXG_PRAGMA_(unroll n)

// Scalar type of the algorithm.  The product (device) build is float, the only type the
// reference computes in (fp32).  The host emulator (tests/simt/) also
// instantiates the SAME source with BXG_REAL = double, so that the kernel logic of every
// variant can be compared leaf by leaf, to 1e-9, with the reference-source goldens, which
// are float64 (tools/gen_reference_golden.py): no percentile, no float32 noise floor.
#ifndef BXG_REAL
#define BXG_REAL float
#endif
#define R(x) ((::bxg::real)(x))   // a literal of the scalar type: R(1e-8) is exactly 1e-8f in the float build

namespace bxg {

using real = BXG_REAL;
#if defined(__CUDA_ARCH__)
static_assert(sizeof(real) == 4, "device code is float only");
#endif

BXG_HD float r_abs(float x) { return fabsf(x); }
BXG_HD float r_sqrt(float x) { return sqrtf(x); }
BXG_HD float r_max(float a, float b) { return fmaxf(a, b); }
BXG_HD float r_min(float a, float b) { return fminf(a, b); }
BXG_HD float r_pow(float a, float b) { return powf(a, b); }
BXG_HD float r_fma(float a, float b, float c) { return fmaf(a, b, c); }
// sin and cos in float from ONE source for host and device (libm's and CUDA's sincosf differ in the last bit):
// Cody-Waite reduction by pi/2 in three parts with fused multiply-adds, then the Cephes minimax polynomials on
// [-pi/4, pi/4].  Measured against float64: <= 1.5 ulp for |x| <= 100, <= 2.4 ulp (sin) up to 1e4
// (tests/test_kernel_math.py); the angles here are joint angles and dt * |omega| / 2.
BXG_HD void r_sincos(float x, float* s, float* c) {
  const float k = rintf(x * 0.636619747f);                       // x * 2/pi, to nearest even
  float r = fmaf(k, -1.57079637050628662109e+00f, x);
  r = fmaf(k, 4.37113882867379288655e-08f, r);
  r = fmaf(k, 1.77635683940025046468e-15f, r);
  r = r < -1.0f ? -1.0f : (r > 1.0f ? 1.0f : r);                  // no-op in range (|r| <= pi/4 + rounding); keeps the result bounded for |x| > ~1e6, NaN stays NaN
  const float z = r * r;
  float sp = fmaf(z, -1.9515295891e-4f, 8.3321608736e-3f);
  sp = fmaf(sp, z, -1.6666654611e-1f);
  const float sn = fmaf(sp * z, r, r);
  float cp = fmaf(z, 2.443315711809948e-5f, -1.388731625493765e-3f);
  cp = fmaf(cp, z, 4.166664568298827e-2f);
  const float cs = fmaf(cp * z, z, fmaf(z, -0.5f, 1.0f));
  const int q = (int)fminf(fmaxf(k, -1.0e9f), 1.0e9f) & 3;       // (far outside the reduction's range the result is only bounded)
  *s = q == 0 ? sn : q == 1 ? cs : q == 2 ? -sn : -cs;
  *c = q == 0 ? cs : q == 1 ? -sn : q == 2 ? -cs : sn;
}
#if !defined(__CUDA_ARCH__)
inline double r_abs(double x) { return fabs(x); }
inline double r_sqrt(double x) { return sqrt(x); }
inline double r_max(double a, double b) { return fmax(a, b); }
inline double r_min(double a, double b) { return fmin(a, b); }
inline double r_pow(double a, double b) { return pow(a, b); }
inline double r_fma(double a, double b, double c) { return fma(a, b, c); }
inline void r_sincos(double x, double* s, double* c) { *s = sin(x); *c = cos(x); }
#endif
// machine epsilon of the scalar type (jaxopt's line search adds jnp.finfo(dtype).eps)
BXG_HD real r_eps() { return sizeof(real) == 4 ? (real)1.1920929e-07 : (real)2.220446049250313e-16; }

// ------------------------------------------------------------------ algebra
struct V3 { real x, y, z; };
struct Q4 { real w, x, y, z; };

struct alignas(16) F4 { real x, y, z, w; };
// 128-bit shared-memory access (LDS.128 / STS.128 on device); p must be 16-byte aligned
BXG_HD F4 ldv4(const real* p) { return *reinterpret_cast<const F4*>(p); }
BXG_HD void stv4(real* p, F4 v) { *reinterpret_cast<F4*>(p) = v; }

BXG_HD V3 ld3(const real* p) { return V3{p[0], p[1], p[2]}; }
BXG_HD void st3(real* p, V3 v) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }
BXG_HD Q4 ld4(const real* p) { return Q4{p[0], p[1], p[2], p[3]}; }
BXG_HD void st4(real* p, Q4 q) { p[0] = q.w; p[1] = q.x; p[2] = q.y; p[3] = q.z; }
BXG_HD V3 operator+(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
BXG_HD V3 operator-(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
BXG_HD V3 operator*(V3 a, real s) { return V3{a.x * s, a.y * s, a.z * s}; }
// Fused multiply-adds are written out (r_fma): the device code is compiled with -fmad=false and the host
// emulator with -ffp-contract=off, so the SOURCE fixes every rounding and both produce the same bits
// (tests/test_gpu_bitexact.py).  s * a + b on vectors:
BXG_HD V3 fmav(V3 a, real s, V3 b) { return V3{r_fma(a.x, s, b.x), r_fma(a.y, s, b.y), r_fma(a.z, s, b.z)}; }
BXG_HD real dot(V3 a, V3 b) { return r_fma(a.z, b.z, r_fma(a.y, b.y, a.x * b.x)); }
BXG_HD V3 cross(V3 a, V3 b) { return V3{r_fma(a.y, b.z, -(a.z * b.y)), r_fma(a.z, b.x, -(a.x * b.z)), r_fma(a.x, b.y, -(a.y * b.x))}; }

// math.rotate (brax/math.py:25-41)
BXG_HD V3 rotate(V3 v, Q4 q) {
  V3 u{q.x, q.y, q.z};
  real s = q.w, d = dot(u, v), k = r_fma(s, s, -dot(u, u));
  V3 c = cross(u, v);
  const real d2 = R(2.) * d;
  V3 r{r_fma(k, v.x, d2 * u.x), r_fma(k, v.y, d2 * u.y), r_fma(k, v.z, d2 * u.z)};
  real s2 = R(2.) * s;
  return V3{r_fma(s2, c.x, r.x), r_fma(s2, c.y, r.y), r_fma(s2, c.z, r.z)};
}
// math.quat_mul (brax/math.py:86-101)
BXG_HD Q4 qmul(Q4 u, Q4 v) {
  return Q4{r_fma(-u.z, v.z, r_fma(-u.y, v.y, r_fma(-u.x, v.x, u.w * v.w))),
            r_fma(-u.z, v.y, r_fma(u.y, v.z, r_fma(u.x, v.w, u.w * v.x))),
            r_fma(u.z, v.x, r_fma(u.y, v.w, r_fma(-u.x, v.z, u.w * v.y))),
            r_fma(u.z, v.w, r_fma(-u.y, v.x, r_fma(u.x, v.y, u.w * v.z)))};
}
// math.normalize on a quaternion incl. safe_norm's all-close-to-zero rule
// (brax/math.py:308-345)
BXG_HD Q4 qnormalize(Q4 q) {
  bool zero = r_abs(q.w) <= R(1e-8) && r_abs(q.x) <= R(1e-8) && r_abs(q.y) <= R(1e-8) && r_abs(q.z) <= R(1e-8);
  real n = zero ? R(0.) : r_sqrt(r_fma(q.z, q.z, r_fma(q.y, q.y, r_fma(q.x, q.x, q.w * q.w))));
  real d = n + R(1e-6) * (n == R(0.) ? R(1.) : R(0.));
  return Q4{q.w / d, q.x / d, q.y / d, q.z / d};
}
// normalize(quat_rot_axis(axis, angle)) (brax/math.py:133-147)
BXG_HD Q4 axis_quat(V3 axis, real angle) {
  real sn, cs;
  r_sincos(angle * R(0.5), &sn, &cs);
  return qnormalize(Q4{cs, axis.x * sn, axis.y * sn, axis.z * sn});
}
// Transform.do(Transform) (brax/base.py:557-562)
BXG_HD void tf_do(V3 ap, Q4 ar, V3 bp, Q4 br, V3* op, Q4* orr) {
  *op = ap + rotate(bp, ar);
  *orr = qmul(ar, br);
}
// Inertia.mul(Motion) -> Force (brax/base.py:297-302); im row-major 3x3
BXG_HD void inertia_mul(V3 ipos, const real* im, real mass, V3 mang, V3 mvel, V3* fang, V3* fvel) {
  V3 c1 = cross(ipos, mvel), c2 = cross(ipos, mang);
  fang->x = r_fma(im[2], mang.z, r_fma(im[1], mang.y, im[0] * mang.x)) + c1.x;
  fang->y = r_fma(im[5], mang.z, r_fma(im[4], mang.y, im[3] * mang.x)) + c1.y;
  fang->z = r_fma(im[8], mang.z, r_fma(im[7], mang.y, im[6] * mang.x)) + c1.z;
  *fvel = V3{r_fma(mvel.x, mass, -c2.x), r_fma(mvel.y, mass, -c2.y), r_fma(mvel.z, mass, -c2.z)};
}

struct Ctx {
  const Dims* D;
  const real* mf;  // model blob viewed as float
  const int* mi;    // model blob viewed as int32
  real* s;         // this env's slab
};

struct Stats {
  int pg_iters, pg_trials, ns_accepts, ns_cold;
#if defined(BXG_PHASE_TIMERS)
  long long t_last;                  // tuning builds: start of the current phase (clock64)
  unsigned long long* phase_cycles;  // BxgDiag.phase_cycles: lane 0 of every warp adds its cycles at each phase end
#endif
};
// phases: 0 load, 1 dynamics, 2 constraint.force, 12 the CTA barrier behind it, 3 integrate, 4 kinematics, 5 transform_com,
// 6 mass.matrix, 7 matrix_inv, 8 constraint.jacobian, 9 env prologue / epilogue, 10 store, 11 lean entry,
// 13 / 14 / 15 inside constraint.force: active set + A + b, projected-gradient iterations without their line searches, line-search trials
#if defined(BXG_PHASE_TIMERS) && defined(__CUDA_ARCH__)
#define BXG_PHASE_BEGIN(st) ((st)->t_last = clock64())
#define BXG_PHASE_END(st, k) do { long long t__ = clock64(); if ((st)->phase_cycles && (threadIdx.x & 31) == 0) atomicAdd((st)->phase_cycles + (k), (unsigned long long)(t__ - (st)->t_last)); (st)->t_last = clock64(); } while (0)
#else
#define BXG_PHASE_BEGIN(st) ((void)0)
#define BXG_PHASE_END(st, k) ((void)0)
#endif

// Compile-time kernel configuration: lanes per env and register-row widths.
// GENERIC_TOO keeps the any-size code reachable next to the specialised kernels
// (host emulator only: the CPU tests compare the two); device variants drop it.
template <int G_, int VC4_, int NC4_, bool GENERIC_TOO_ = false>
struct KernelCfg {
  static constexpr int G = G_, VC4 = VC4_, NC4 = NC4_;
  static constexpr int MAX_THREADS = variant_max_threads(G_, VC4_, NC4_);
  static constexpr bool GENERIC_TOO = GENERIC_TOO_ || VC4_ == 0;
};

// one byte per constraint row: 1 where the row is active (else J, diag and aref of the row are all zero)
BXG_HD uint8_t* row_active(real* s, const Dims& D) { return reinterpret_cast<uint8_t*>(s + D.s_rowact); }

// constraint._imp_aref (brax/generalized/constraint.py:29-65); prm as packed by pack_impedance
// (bxg_model.h): the row-constant quotients are precomputed on the host
BXG_HD_NOINLINE void imp_aref(const real* prm, real pos, real vel, real* imp_out, real* aref_out) {
  const real dmin = prm[0], dmax = prm[1], width = prm[2], mid = prm[3], power = prm[4];
  const real inv_a = prm[5], inv_b = prm[6], b = prm[7], k = prm[8];
  real imp_x = r_abs(pos) / width;
  real imp_a, imp_b;
  if (power == R(2.)) {
    // x^2 is an exact operation: no transcendental needed (MuJoCo's default solimp power;
    // XLA's simplifier lowers pow(x, 2) to x * x as well)
    imp_a = inv_a * (imp_x * imp_x);
    imp_b = R(1.) - inv_b * ((R(1.) - imp_x) * (R(1.) - imp_x));
  } else {
    imp_a = inv_a * r_pow(imp_x, power);
    imp_b = R(1.) - inv_b * r_pow(R(1.) - imp_x, power);
  }
  real imp_y = imp_x < mid ? imp_a : imp_b;
  real imp = dmin + imp_y * (dmax - dmin);
  imp = r_max(dmin, r_min(imp, dmax));
  if (imp_x > R(1.0)) imp = dmax;
  *imp_out = imp;
  *aref_out = -b * vel - k * imp * pos;
}

// ------------------------------------------------ forces: tau, passive, RNE
// actuator.to_tau (brax/actuator.py:23-57), lanes <-> dofs, into s_tau
template <class X>
BXG_HD void actuator_tau(X& ex, const Ctx& c) {
  BXG_GET_DIMS(c); const real* mf = c.mf; const int* mi = c.mi; real* s = c.s;
  const int nv = D.nv;
  ex.lanes([&](int lane) {
    for (int d = lane; d < nv; d += X::G) {
      real tau = R(0.);
      for (int k = mi[D.m_dof_act_start + d]; k < mi[D.m_dof_act_start + d + 1]; ++k) {
        int a = mi[D.m_dof_act_list + k];
        real qv = s[D.s_q + mi[D.m_act_qid + a]], qdv = s[D.s_qd + mi[D.m_act_did + a]];
        real ctrl = r_max(mf[D.m_act_clo + a], r_min(s[D.s_act + a], mf[D.m_act_chi + a]));
        real gear = mf[D.m_act_gear + a];
        real bias = gear * (qv * mf[D.m_act_bq + a] + qdv * mf[D.m_act_bqd + a]);
        real f = mf[D.m_act_gain + a] * ctrl + bias;
        f = r_max(mf[D.m_act_flo + a], r_min(f, mf[D.m_act_fhi + a]));
        tau = r_fma(f, gear, tau);
      }
      s[D.s_tau + d] = tau;
    }
  });
}

// fluid.force (brax/fluid.py:24-91) projected on the dofs as dynamics._passive does
// (dynamics.py:198-211).  Lanes <-> links: force of each link's inertia box in the world
// orientation into s_t_ang / s_t_vel, its lever arm x_i.pos - root_com into s_f_ang; then
// lanes <-> dofs: sum over the links the dof moves of jac . frc, left in s_qfs for the
// assembly of qf_smooth.  Runs before the RNE passes, which reuse the same temporaries.
template <class X>
BXG_HD void fluid_passive(X& ex, const Ctx& c) {
  BXG_GET_DIMS(c); const real* mf = c.mf; const int* mi = c.mi; real* s = c.s;
  const int L = D.L, nv = D.nv;
  ex.lanes([&](int lane) {
    for (int l = lane; l < L; l += X::G) {
      V3 xi_pos; Q4 xi_rot;
      tf_do(ld3(s + D.s_x_pos + 3 * l), ld4(s + D.s_x_rot + 4 * l), ld3(mf + D.m_in_pos + 3 * l), ld4(mf + D.m_in_rot + 4 * l), &xi_pos, &xi_rot);
      V3 off = xi_pos - ld3(s + D.s_root_com + 3 * l);
      Q4 rinv{xi_rot.w, -xi_rot.x, -xi_rot.y, -xi_rot.z};
      V3 cda = ld3(s + D.s_cd_ang + 3 * l);
      V3 ang = rotate(cda, rinv);
      V3 vel = rotate(ld3(s + D.s_cd_vel + 3 * l) - cross(off, cda), rinv);
      const real* k = mf + D.m_fluid + kFluidStride * l;
      V3 fa{k[0] * ang.x + k[5] * r_abs(ang.x) * ang.x / R(64.0), k[0] * ang.y + k[6] * r_abs(ang.y) * ang.y / R(64.0),
            k[0] * ang.z + k[7] * r_abs(ang.z) * ang.z / R(64.0)};
      V3 fv{k[1] * vel.x + k[2] * r_abs(vel.x) * vel.x, k[1] * vel.y + k[3] * r_abs(vel.y) * vel.y,
            k[1] * vel.z + k[4] * r_abs(vel.z) * vel.z};
      st3(s + D.s_t_vel + 3 * l, rotate(fv, xi_rot));
      st3(s + D.s_t_ang + 3 * l, rotate(fa, xi_rot));
      st3(s + D.s_f_ang + 3 * l, off);
    }
  });
  ex.lanes([&](int lane) {
    for (int d = lane; d < nv; d += X::G) {
      const int dl = mi[D.m_dof_link + d];
      V3 ca = ld3(s + D.s_cdof_ang + 3 * d), cv = ld3(s + D.s_cdof_vel + 3 * d);
      real acc = R(0.);
      for (int l = dl; l < L; ++l) {
        int p = l;
        while (p > dl) p = mi[D.m_link_parent + p];   // parents precede their children
        if (p != dl) continue;
        V3 jv = cv - cross(ld3(s + D.s_f_ang + 3 * l), ca);
        acc += dot(jv, ld3(s + D.s_t_vel + 3 * l)) + dot(ca, ld3(s + D.s_t_ang + 3 * l));
      }
      s[D.s_qfs + d] = acc;
    }
  });
}

template <class X, class Cfg>
BXG_HD void dyn_forces(X& ex, const Ctx& c) {
  BXG_GET_DIMS(c); const real* mf = c.mf; const int* mi = c.mi; real* s = c.s;
  const int L = D.L, nv = D.nv;
  // fluid models run on the generic variant or on the small 8-wide one (bxg_model.h): the others carry no fluid code
  constexpr bool kFluidCode = Cfg::VC4 == 0 || (Cfg::G == 4 && Cfg::VC4 == 2);
  if constexpr (kFluidCode) { if (D.fluid) fluid_passive(ex, c); }
  actuator_tau(ex, c);
  // RNE forward scan: cdd, then cfrc_flat (lanes <-> links, one tree level at a time)
  for (int lvl = 0; lvl <= D.max_depth; ++lvl) {
    ex.lanes([&](int l) {
      if (l >= L || mi[D.m_link_depth + l] != lvl) return;
      int p = mi[D.m_link_parent + l], da = mi[D.m_link_dadr + l], nd = mi[D.m_link_ndof + l];
      nd = nd == 0 ? 6 : nd;
      V3 ca = p >= 0 ? ld3(s + D.s_t_ang + 3 * p) : V3{R(0.), R(0.), R(0.)};
      V3 cv = p >= 0 ? ld3(s + D.s_t_vel + 3 * p) : V3{-D.gx, -D.gy, -D.gz};
      for (int k = 0; k < nd; ++k) {
        real qd = s[D.s_qd + da + k];
        ca = fmav(ld3(s + D.s_cdofd_ang + 3 * (da + k)), qd, ca);
        cv = fmav(ld3(s + D.s_cdofd_vel + 3 * (da + k)), qd, cv);
      }
      st3(s + D.s_t_ang + 3 * l, ca); st3(s + D.s_t_vel + 3 * l, cv);
    });
  }
  // cfrc = cinr.mul(cdd) + cd.cross(cinr.mul(cd)): no dependency between links, one pass
  {
    ex.lanes([&](int l) {
      if (l >= L) return;
      V3 ca = ld3(s + D.s_t_ang + 3 * l), cv = ld3(s + D.s_t_vel + 3 * l);
      V3 ip = ld3(s + D.s_cinr_pos + 3 * l); const real* im = s + D.s_cinr_i + 9 * l; real mass = mf[D.m_in_mass + l];
      V3 cda = ld3(s + D.s_cd_ang + 3 * l), cdv = ld3(s + D.s_cd_vel + 3 * l);
      V3 fa, fv, ga, gv;
      inertia_mul(ip, im, mass, ca, cv, &fa, &fv);
      inertia_mul(ip, im, mass, cda, cdv, &ga, &gv);
      V3 c1 = cross(cda, gv), c2 = cross(cda, ga), c3 = cross(cdv, gv);
      st3(s + D.s_f_vel + 3 * l, fv + c1);
      st3(s + D.s_f_ang + 3 * l, fa + (c2 + c3));
    });
  }
  // RNE backward scan: accumulate children (descending index order)
  for (int lvl = D.max_depth - 1; lvl >= 0; --lvl) {
    ex.lanes([&](int l) {
      if (l >= L || mi[D.m_link_depth + l] != lvl) return;
      V3 fa = ld3(s + D.s_f_ang + 3 * l), fv = ld3(s + D.s_f_vel + 3 * l);
      for (int k = mi[D.m_child_start + l]; k < mi[D.m_child_start + l + 1]; ++k) {
        int ch = mi[D.m_child_list + k];
        fa = fa + ld3(s + D.s_f_ang + 3 * ch); fv = fv + ld3(s + D.s_f_vel + 3 * ch);
      }
      st3(s + D.s_f_ang + 3 * l, fa); st3(s + D.s_f_vel + 3 * l, fv);
    });
  }
  // qf_smooth = passive - bias + tau (lanes <-> dofs)
  ex.lanes([&](int lane) {
    for (int d = lane; d < nv; d += X::G) {
      int l = mi[D.m_dof_link + d], qi = mi[D.m_dof_qidx + d];
      real bias = dot(ld3(s + D.s_cdof_vel + 3 * d), ld3(s + D.s_f_vel + 3 * l)) +
                   dot(ld3(s + D.s_cdof_ang + 3 * d), ld3(s + D.s_f_ang + 3 * l));
      real qd = s[D.s_qd + d];
      real passive = qi < 0 ? R(0.) : -s[D.s_q + qi] * mf[D.m_stiff + d];
      passive = r_fma(-mf[D.m_damp + d], qd, passive);
      if constexpr (kFluidCode) { if (D.fluid) passive = passive + s[D.s_qfs + d]; }
      s[D.s_qfs + d] = (passive - bias) + s[D.s_tau + d];
    }
  });
}

// --------------------------------------- constraint.force + projected gradient
template <class X>
BXG_HD void con_force_generic(X& ex, const Ctx& c, Stats* st) {
  BXG_GET_DIMS(c); real* s = c.s;
  const int nv = D.nv, nc = D.nc, nvp = D.nvp, ncp = D.ncp;
  if (nc == 0) {
    ex.lanes([&](int lane) { for (int d = lane; d < nv; d += X::G) s[D.s_qfc + d] = R(0.); });
    return;
  }
  real* J = s + D.s_J; real* Mi = s + D.s_Minv; const int jld = D.jld;
  real* JM = s + D.s_JM; real* A = s + D.s_A;
  real* b = s + D.s_b; real* x = s + D.s_px; real* y = s + D.s_py; real* g = s + D.s_pg;
  real* res = s + D.s_pres; real* xn = s + D.s_pxn;
  // A = J Minv J^T + diag, b = J Minv qf_smooth - aref   (lanes <-> rows)
  ex.lanes([&](int lane) {
    for (int i = lane; i < nc; i += X::G) {
      for (int j = 0; j < nv; ++j) {
        real acc = R(0.);
        for (int k = 0; k < nv; ++k) acc = r_fma(J[i * jld + k], Mi[k * nvp + j], acc);
        JM[i * jld + j] = acc;
      }
      for (int j = 0; j < nc; ++j) {
        real acc = R(0.);
        for (int k = 0; k < nv; ++k) acc = r_fma(JM[i * jld + k], J[j * jld + k], acc);
        A[i * ncp + j] = acc + (i == j ? s[D.s_diag + i] : R(0.));
      }
      real acc = R(0.);
      for (int k = 0; k < nv; ++k) acc = r_fma(JM[i * jld + k], s[D.s_qfs + k], acc);
      b[i] = acc - s[D.s_aref + i];
      x[i] = R(0.); y[i] = R(0.);
    }
  });
  real t = R(1.), stepsize = R(1.), error = INFINITY;
  const real tol = R(1e-3), eps = r_eps();
  int it = 0;
  while (it < D.solver_iterations && (it == 0 || error > tol)) {
    ex.lanes([&](int lane) {
      for (int i = lane; i < nc; i += X::G) {
        real acc = R(0.);
        for (int j = 0; j < nc; ++j) acc = r_fma(A[i * ncp + j], y[j], acc);
        res[i] = acc + b[i];
      }
    });
    real fy = R(0.);
    for (int i = 0; i < nc; ++i) fy = r_fma(R(0.5), (res[i] * res[i]), fy);
    ex.lanes([&](int lane) {
      for (int j = lane; j < nc; j += X::G) {
        real acc = R(0.);
        for (int i = 0; i < nc; ++i) acc = r_fma(A[i * ncp + j], res[i], acc);
        g[j] = acc;
      }
    });   // (writes g only: no hazard with the redundant reads of res above)
    real sz = stepsize;
    for (int ls = 0;; ++ls) {
      ex.lanes([&](int lane) { for (int i = lane; i < nc; i += X::G) xn[i] = r_max(y[i] - sz * g[i], R(0.)); });
      ex.lanes([&](int lane) {
        for (int i = lane; i < nc; i += X::G) {
          real acc = R(0.);
          for (int j = 0; j < nc; ++j) acc = r_fma(A[i * ncp + j], xn[j], acc);
          res[i] = acc + b[i];
        }
      });
      real sqdist = R(0.), vd = R(0.), fn = R(0.);
      for (int i = 0; i < nc; ++i) { real dlt = xn[i] - y[i]; sqdist = r_fma(dlt, dlt, sqdist); }
      for (int i = 0; i < nc; ++i) { real dlt = xn[i] - y[i]; vd = r_fma(dlt, g[i], vd); }
      for (int i = 0; i < nc; ++i) fn = r_fma(R(0.5), (res[i] * res[i]), fn);
      st->pg_trials++;
      ex.sync();   // every lane finished its redundant reads of xn / res before they are rewritten
      real fun_decrease = sz * (fn - fy);
      real condition = sz * vd + R(0.5) * sqdist;
      if (!(fun_decrease > condition + eps) || ls >= D.solver_maxls) break;
      sz = sz * R(0.5);
    }
    stepsize = sz <= R(1e-6) ? R(1.) : sz / R(0.5);
    real tn = R(0.5) * (R(1.) + r_sqrt(R(1.) + R(4.) * (t * t)));
    real mom = (t - R(1.)) / tn;
    // res now holds A x+ + b: reuse it for the error gradient
    ex.lanes([&](int lane) {
      for (int j = lane; j < nc; j += X::G) {
        real acc = R(0.);
        for (int i = 0; i < nc; ++i) acc = r_fma(A[i * ncp + j], res[i], acc);
        g[j] = acc;
        real dlt = xn[j] - x[j];
        y[j] = xn[j] + mom * dlt;
        x[j] = xn[j];
      }
    });
    real err2 = R(0.);
    for (int i = 0; i < nc; ++i) { real dlt = r_max(xn[i] - g[i], R(0.)) - xn[i]; err2 = r_fma(dlt, dlt, err2); }
    error = r_sqrt(err2);
    t = tn;
    ++it;
    st->pg_iters++;
    ex.sync();
  }
  // qf_constraint = J^T x
  ex.lanes([&](int lane) {
    for (int j = lane; j < nv; j += X::G) {
      real acc = R(0.);
      for (int i = 0; i < nc; ++i) acc = r_fma(J[i * jld + j], x[i], acc);
      s[D.s_qfc + j] = acc;
    }
  });
}

// SPD inverse by Cholesky: dst = src^-1, Lm is an nv x nvp scratch.
// Used by init (mass.py:103-104), integrate's implicit-damping branch
// (integrator.py:58-60) and BXG_MINV_CHOLESKY.
template <class X>
BXG_HD void spd_inverse(X& ex, const Ctx& c, const real* src, real* dst, real* Lm, const real* add_diag, real diag_scale) {
  BXG_GET_DIMS(c);
  const int n = D.nv, nvp = D.nvp;
  for (int j = 0; j < n; ++j) {
    ex.lanes([&](int lane) {
      real sd = src[j * nvp + j] + (add_diag ? add_diag[j] * diag_scale : R(0.));
      for (int k = 0; k < j; ++k) sd -= Lm[j * nvp + k] * Lm[j * nvp + k];
      real dg = r_sqrt(sd);
      for (int i = j + lane; i < n; i += X::G) {
        if (i == j) { Lm[j * nvp + j] = dg; continue; }
        real tt = src[i * nvp + j];
        for (int k = 0; k < j; ++k) tt -= Lm[i * nvp + k] * Lm[j * nvp + k];
        Lm[i * nvp + j] = tt / dg;
      }
    });
  }
  ex.lanes([&](int lane) {
    for (int col = lane; col < n; col += X::G) {
      for (int i = 0; i < n; ++i) {
        real tt = i == col ? R(1.) : R(0.);
        for (int k = 0; k < i; ++k) tt -= Lm[i * nvp + k] * dst[k * nvp + col];
        dst[i * nvp + col] = tt / Lm[i * nvp + i];
      }
      for (int i = n - 1; i >= 0; --i) {
        real tt = dst[i * nvp + col];
        for (int k = i + 1; k < n; ++k) tt -= Lm[k * nvp + i] * dst[k * nvp + col];
        dst[i * nvp + col] = tt / Lm[i * nvp + i];
      }
    }
  });
}

// SPD inverse for the specialised variants: same arithmetic, in the same order, as
// spd_inverse above, restructured for the lane group.  Factorisation: lane i keeps row i of
// L in registers; per column j every lane reads row j of L (complete by then) with
// broadcast 128-bit loads.  L is stored together with its transpose (Lm[j][i] = Lm[i][j]),
// so that the back substitution reads rows too.  Solve: lane c owns column c of the
// inverse in registers.  W = nvw.
template <class X, int W>
BXG_HD void spd_inverse_rows(X& ex, const Ctx& c, const real* src, real* dst, real* Lm) {
  BXG_GET_DIMS(c);
  const int n = D.nv, ld = D.nvp;
  typename X::template LaneVec<W> lrow;
#pragma unroll
  for (int j = 0; j < W; ++j) {
    if (j < n) {
      ex.lanes([&](int i) {
        real lj[W > 1 ? W : 1];
#pragma unroll
        for (int k4 = 0; k4 < j; k4 += 4) { F4 t = ldv4(Lm + j * ld + k4); lj[k4] = t.x; if (k4 + 1 < W) lj[k4 + 1] = t.y; if (k4 + 2 < W) lj[k4 + 2] = t.z; if (k4 + 3 < W) lj[k4 + 3] = t.w; }
        real sd = src[j * ld + j];
#pragma unroll
        for (int k = 0; k < j; ++k) sd -= lj[k] * lj[k];
        const real dg = r_sqrt(sd);
        if (i == j) {
          lrow(i)[j] = dg; Lm[j * ld + j] = dg;
        } else if (i > j && i < n) {
          real tt = src[i * ld + j];
#pragma unroll
          for (int k = 0; k < j; ++k) tt -= lrow(i)[k] * lj[k];
          const real v = tt / dg;
          lrow(i)[j] = v; Lm[i * ld + j] = v; Lm[j * ld + i] = v;
        }
      });
    }
  }
  ex.lanes([&](int col) {
    if (col >= n) return;
    real y[W];
#pragma unroll
    for (int i = 0; i < W; ++i) {
      if (i < n) {
        real li[W];
#pragma unroll
        for (int k4 = 0; k4 <= i; k4 += 4) { F4 t = ldv4(Lm + i * ld + k4); li[k4] = t.x; if (k4 + 1 < W) li[k4 + 1] = t.y; if (k4 + 2 < W) li[k4 + 2] = t.z; if (k4 + 3 < W) li[k4 + 3] = t.w; }
        real tt = i == col ? R(1.) : R(0.);
#pragma unroll
        for (int k = 0; k < i; ++k) tt -= li[k] * y[k];
        y[i] = tt / li[i];
      }
    }
#pragma unroll
    for (int i = W - 1; i >= 0; --i) {
      if (i < n) {
        real li[W];
#pragma unroll
        for (int k4 = (i / 4) * 4; k4 < W; k4 += 4) { F4 t = ldv4(Lm + i * ld + k4); li[k4] = t.x; if (k4 + 1 < W) li[k4 + 1] = t.y; if (k4 + 2 < W) li[k4 + 2] = t.z; if (k4 + 3 < W) li[k4 + 3] = t.w; }
        real tt = y[i];
#pragma unroll
        for (int k = i + 1; k < W; ++k) if (k < n) tt -= li[k] * y[k];
        y[i] = tt / li[i];
      }
    }
#pragma unroll
    for (int i = 0; i < W; ++i) if (i < n) dst[i * ld + col] = y[i];
  });
}

// ------------------------------------------------------ integrator.integrate
template <class X>
BXG_HD void integrate(X& ex, const Ctx& c) {
  BXG_GET_DIMS(c); const real* mf = c.mf; const int* mi = c.mi; real* s = c.s;
  const int nv = D.nv, nvp = D.nvp, L = D.L;
  const real dt = D.dt;
  const real* Mi = s + D.s_Minv;
  if (D.ns_iters == 0) {
    real* tmp = s + D.s_scr;  // dst
    spd_inverse(ex, c, s + D.s_M, tmp, s + D.s_scr + D.nvw * nvp, mf + D.m_damp, dt);
    Mi = tmp;
  }
  ex.lanes([&](int lane) {
    for (int i = lane; i < nv; i += X::G) {
      real acc = R(0.);
      for (int j = 0; j < nv; ++j) acc = r_fma(Mi[i * nvp + j], (s[D.s_qfs + j] + s[D.s_qfc + j]), acc);
      s[D.s_qdd + i] = acc;
    }
  });
  ex.lanes([&](int lane) {
    for (int i = lane; i < nv; i += X::G) s[D.s_qd + i] = r_fma(s[D.s_qdd + i], dt, s[D.s_qd + i]);
  });
  ex.lanes([&](int l) {
    if (l >= L) return;
    int qa = mi[D.m_link_qadr + l], da = mi[D.m_link_dadr + l], nd = mi[D.m_link_ndof + l];
    if (nd == 0) {
      Q4 rot = ld4(s + D.s_q + qa + 3);
      V3 ang = ld3(s + D.s_qd + da + 3);
      real ang_norm = r_sqrt(ang.x * ang.x + ang.y * ang.y + ang.z * ang.z) + R(1e-8);
      V3 axis{ang.x / ang_norm, ang.y / ang_norm, ang.z / ang_norm};
      real sn, cs;
      r_sincos((dt * ang_norm) * R(0.5), &sn, &cs);
      Q4 nr = qmul(rot, Q4{cs, axis.x * sn, axis.y * sn, axis.z * sn});
      real n = r_sqrt(nr.w * nr.w + nr.x * nr.x + nr.y * nr.y + nr.z * nr.z);
      st4(s + D.s_q + qa + 3, Q4{nr.w / n, nr.x / n, nr.y / n, nr.z / n});
      for (int i = 0; i < 3; ++i) s[D.s_q + qa + i] = r_fma(s[D.s_qd + da + i], dt, s[D.s_q + qa + i]);
    } else {
      for (int k = 0; k < nd; ++k) s[D.s_q + qa + k] = r_fma(s[D.s_qd + da + k], dt, s[D.s_q + qa + k]);
    }
  });
}

// -------------------------------------------------------- kinematics.forward
// VEL = false: link poses x only.  Nothing in the physics reads the link velocities xd (they are an output leaf and
// an input of some env observations), so the substeps skip them; finish_env() runs the full function once, after
// the last substep, and xd then lives in the t_ang / t_vel temporaries (dead by then) instead of its own slab words.
// The joint transform / motion of a link is produced and consumed by the same lane: it stays in registers.
template <class X, bool VEL>
BXG_HD void kinematics(X& ex, const Ctx& c) {
  BXG_GET_DIMS(c); const real* mf = c.mf; const int* mi = c.mi; real* s = c.s;
  const int L = D.L;
#if BXG_KIN_REGS
  typename X::template LaneVec<VEL ? 13 : 7> jt;   // [pos 3, rot 4, (ang 3, vel 3)] of this lane's link
#else
  typename X::template LaneVec<VEL ? 6 : 1> jdv;   // joint motion (ang 3, vel 3) of this lane's link; the joint transform goes through shared memory
  real* jpos = s + D.s_f_ang; real* jrot = s + D.s_j_rot;
#endif
  ex.lanes([&](int l) {
    if (l >= L) return;
    int qa = mi[D.m_link_qadr + l], da = mi[D.m_link_dadr + l], nd = mi[D.m_link_ndof + l];
    V3 p, a{0, 0, 0}, v{0, 0, 0}; Q4 r;
    if (nd == 0) {
      p = ld3(s + D.s_q + qa); r = ld4(s + D.s_q + qa + 3);
      if constexpr (VEL) { v = ld3(s + D.s_qd + da); a = ld3(s + D.s_qd + da + 3); }
    } else {
      p = V3{0, 0, 0}; r = Q4{1, 0, 0, 0};
      for (int k = 0; k < nd; ++k) {
        int d = da + k; real qk = s[D.s_q + qa + k];
        V3 mang = ld3(mf + D.m_dof_ang + 3 * d), mvel = ld3(mf + D.m_dof_vel + 3 * d);
        Q4 sr = axis_quat(mang, qk);
        V3 sp = mvel * qk;
        V3 sa{0, 0, 0}, sv{0, 0, 0};
        if constexpr (VEL) { real qdk = s[D.s_qd + d]; sa = mang * qdk; sv = mvel * qdk; }
        if (k == 0) { p = sp; r = sr; a = sa; v = sv; }
        else {
          V3 np_; Q4 nr;
          tf_do(p, r, sp, sr, &np_, &nr);
          if constexpr (VEL) {
            a = a + rotate(sa, sr);
            v = v + rotate(sv + cross(sp, sa), sr);
          }
          p = np_; r = nr;
        }
      }
    }
    V3 jp = ld3(mf + D.m_joint_pos + 3 * l);
    V3 anc = rotate(jp, r);
    p = (p + jp) - anc;
    V3 tp; Q4 tr;
    tf_do(ld3(mf + D.m_tf_pos + 3 * l), ld4(mf + D.m_tf_rot + 4 * l), p, r, &tp, &tr);
#if BXG_KIN_REGS
    real* j = jt(l);
    st3(j, tp); st4(j + 3, tr);
    if constexpr (VEL) { st3(j + 7, a); st3(j + 10, v); }
#else
    st3(jpos + 3 * l, tp); st4(jrot + 4 * l, tr);
    if constexpr (VEL) { st3(jdv(l), a); st3(jdv(l) + 3, v); }
#endif
  });
  for (int lvl = 0; lvl <= D.max_depth; ++lvl) {
    ex.lanes([&](int l) {
      if (l >= L || mi[D.m_link_depth + l] != lvl) return;
      int p = mi[D.m_link_parent + l];
      V3 ja{0, 0, 0}, jv{0, 0, 0};
#if BXG_KIN_REGS
      const real* j = jt(l);
      V3 jp = ld3(j); Q4 jr = ld4(j + 3);
      if constexpr (VEL) { ja = ld3(j + 7); jv = ld3(j + 10); }
#else
      V3 jp = ld3(jpos + 3 * l); Q4 jr = ld4(jrot + 4 * l);
      if constexpr (VEL) { ja = ld3(jdv(l)); jv = ld3(jdv(l) + 3); }
#endif
      if (p < 0) {
        st3(s + D.s_x_pos + 3 * l, jp); st4(s + D.s_x_rot + 4 * l, jr);
        if constexpr (VEL) { st3(s + D.s_xd_ang + 3 * l, rotate(ja, jr)); st3(s + D.s_xd_vel + 3 * l, jv); }
      } else {
        V3 pp = ld3(s + D.s_x_pos + 3 * p); Q4 pr = ld4(s + D.s_x_rot + 4 * p);
        V3 xp; Q4 xr;
        tf_do(pp, pr, jp, jr, &xp, &xr);
        st3(s + D.s_x_pos + 3 * l, xp); st4(s + D.s_x_rot + 4 * l, xr);
        if constexpr (VEL) {
          V3 pa = ld3(s + D.s_xd_ang + 3 * p), pv = ld3(s + D.s_xd_vel + 3 * p);
          V3 vel = (pv + cross(pa, xp - pp)) + rotate(jv, pr);
          V3 ang = pa + rotate(ja, xr);
          st3(s + D.s_xd_ang + 3 * l, ang); st3(s + D.s_xd_vel + 3 * l, vel);
        }
      }
    });
  }
  ex.lanes([&](int l) {
    if (l >= L) return;
    st4(s + D.s_x_rot + 4 * l, qnormalize(ld4(s + D.s_x_rot + 4 * l)));
  });
}
// After the last substep (and after init): kinematics.forward in full.  x comes out as it already is (same inputs,
// same code), xd is produced here, into the t_ang / t_vel temporaries (bxg_model.h: s_xd_* alias them).
template <class X>
BXG_HD void finish_env(X& ex, const Ctx& c) {
#if BXG_XD_TAIL
  kinematics<X, true>(ex, c);
#endif
}

// ---------------------------------------------------- dynamics.transform_com
template <class X>
BXG_HD void transform_com(X& ex, const Ctx& c) {
  BXG_GET_DIMS(c); const real* mf = c.mf; const int* mi = c.mi; real* s = c.s;
  const int L = D.L;
  real* xi_pos = s + D.s_t_ang;
  ex.lanes([&](int l) {
    if (l >= L) return;
    V3 xp; Q4 xr;
    tf_do(ld3(s + D.s_x_pos + 3 * l), ld4(s + D.s_x_rot + 4 * l), ld3(mf + D.m_in_pos + 3 * l), ld4(mf + D.m_in_rot + 4 * l), &xp, &xr);
    st3(xi_pos + 3 * l, xp); st4(s + D.s_cinr_rot + 4 * l, xr);
  });
  ex.lanes([&](int l) {
    if (l >= L) return;
    // root_com: mass-weighted mean over the links of this tree (index order)
    int root = mi[D.m_link_root + l];
    V3 msum{0, 0, 0}; real mtot = R(0.);
    for (int k = 0; k < L; ++k) {
      if (mi[D.m_link_root + k] != root) continue;
      real mk = mf[D.m_in_mass + k];
      msum = fmav(ld3(xi_pos + 3 * k), mk, msum); mtot += mk;
    }
    V3 com{msum.x / mtot, msum.y / mtot, msum.z / mtot};
    st3(s + D.s_root_com + 3 * l, com);
    // cinr = Transform(x_i.pos - com, x_i.rot).do(inertia)  (base.py:588-594)
    real mass = mf[D.m_in_mass + l];
    V3 p = ld3(xi_pos + 3 * l) - com;
    Q4 q = ld4(s + D.s_cinr_rot + 4 * l);
    real R[9];
    {
      real dq = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z, sc = R(2.) / dq;
      real xs = q.x * sc, ys = q.y * sc, zs = q.z * sc;
      real wx = q.w * xs, wy = q.w * ys, wz = q.w * zs, xx = q.x * xs, xy = q.x * ys, xz = q.x * zs;
      real yy = q.y * ys, yz = q.y * zs, zz = q.z * zs;
      R[0] = R(1.) - (yy + zz); R[1] = xy - wz; R[2] = xz + wy;
      R[3] = xy + wz; R[4] = R(1.) - (xx + zz); R[5] = yz - wx;
      R[6] = xz - wy; R[7] = yz + wx; R[8] = R(1.) - (xx + yy);
    }
    const real* I0 = mf + D.m_in_i + 9 * l;
    real T[9];
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
      real acc = R(0.);
      for (int k = 0; k < 3; ++k) acc = r_fma(R[3 * a + k], I0[3 * k + b], acc);
      T[3 * a + b] = acc;
    }
    real h[9] = {R(0.), -p.z, p.y, p.z, R(0.), -p.x, -p.y, p.x, R(0.)};
    for (int a = 0; a < 3; ++a) for (int b = 0; b < 3; ++b) {
      real acc = R(0.), hh = R(0.);
      for (int k = 0; k < 3; ++k) acc = r_fma(T[3 * a + k], R[3 * b + k], acc);
      for (int k = 0; k < 3; ++k) hh = r_fma(h[3 * a + k], h[3 * b + k], hh);
      s[D.s_cinr_i + 9 * l + 3 * a + b] = r_fma(hh, mass, acc);
    }
    st3(s + D.s_cinr_pos + 3 * l, p * mass);
    // joint frame j = parent.do(link.transform).do(link.joint)  (dynamics.py:47-52)
    int nd = mi[D.m_link_ndof + l], qa = mi[D.m_link_qadr + l], da = mi[D.m_link_dadr + l];
    int pi = nd == 0 ? l : mi[D.m_link_parent + l];
    V3 pp = pi >= 0 ? ld3(s + D.s_x_pos + 3 * pi) : V3{0, 0, 0};
    Q4 pr = pi >= 0 ? ld4(s + D.s_x_rot + 4 * pi) : Q4{1, 0, 0, 0};
    V3 tp, jpv; Q4 tr, jrq;
    tf_do(pp, pr, ld3(mf + D.m_tf_pos + 3 * l), ld4(mf + D.m_tf_rot + 4 * l), &tp, &tr);
    tf_do(tp, tr, ld3(mf + D.m_joint_pos + 3 * l), Q4{1, 0, 0, 0}, &jpv, &jrq);
    V3 off = com - jpv;
    // cdof (dynamics.py:55-89)
    if (nd == 0) {
      for (int k = 0; k < 6; ++k) {
        V3 a = rotate(ld3(mf + D.m_dof_ang + 3 * (da + k)), jrq);
        V3 v = ld3(mf + D.m_dof_vel + 3 * (da + k)) - cross(off, a);
        st3(s + D.s_cdof_ang + 3 * (da + k), a); st3(s + D.s_cdof_vel + 3 * (da + k), v);
      }
    } else {
      V3 lp{0, 0, 0}; Q4 lr{1, 0, 0, 0};
      for (int k = 0; k < nd; ++k) {
        int d = da + k;
        V3 mang = ld3(mf + D.m_dof_ang + 3 * d), mvel = ld3(mf + D.m_dof_vel + 3 * d);
        V3 a = rotate(mang, lr);
        V3 v = rotate(mvel, lr) + cross(lp, a);
        real qk = s[D.s_q + qa + k];
        V3 np_; Q4 nr;
        tf_do(lp, lr, mvel * qk, axis_quat(mang, qk), &np_, &nr);
        lp = np_; lr = nr;
        V3 aw = rotate(a, jrq);
        st3(s + D.s_cdof_ang + 3 * d, aw); st3(s + D.s_cdof_vel + 3 * d, v - cross(off, aw));
      }
    }
  });
  // cd forward scan (dynamics.py:92-103)
  for (int lvl = 0; lvl <= D.max_depth; ++lvl) {
    ex.lanes([&](int l) {
      if (l >= L || mi[D.m_link_depth + l] != lvl) return;
      int p = mi[D.m_link_parent + l], da = mi[D.m_link_dadr + l], nd = mi[D.m_link_ndof + l];
      nd = nd == 0 ? 6 : nd;
      V3 ca = p >= 0 ? ld3(s + D.s_cd_ang + 3 * p) : V3{0, 0, 0};
      V3 cv = p >= 0 ? ld3(s + D.s_cd_vel + 3 * p) : V3{0, 0, 0};
      for (int k = 0; k < nd; ++k) {
        real qd = s[D.s_qd + da + k];
        ca = fmav(ld3(s + D.s_cdof_ang + 3 * (da + k)), qd, ca);
        cv = fmav(ld3(s + D.s_cdof_vel + 3 * (da + k)), qd, cv);
      }
      st3(s + D.s_cd_ang + 3 * l, ca); st3(s + D.s_cd_vel + 3 * l, cv);
    });
  }
  // cdofd (dynamics.py:106-130)
  ex.lanes([&](int l) {
    if (l >= L) return;
    int p = mi[D.m_link_parent + l], da = mi[D.m_link_dadr + l], nd = mi[D.m_link_ndof + l];
    if (nd == 0) {
      V3 ca{0, 0, 0}, cv{0, 0, 0};
      for (int k = 0; k < 3; ++k) {
        real qd = s[D.s_qd + da + k];
        ca = fmav(ld3(s + D.s_cdof_ang + 3 * (da + k)), qd, ca);
        cv = fmav(ld3(s + D.s_cdof_vel + 3 * (da + k)), qd, cv);
      }
      for (int k = 0; k < 6; ++k) {
        V3 da_ = ld3(s + D.s_cdof_ang + 3 * (da + k)), dv_ = ld3(s + D.s_cdof_vel + 3 * (da + k));
        V3 vel = cross(ca, dv_) + cross(cv, da_), ang = cross(ca, da_);
        if (k < 3) { vel = V3{0, 0, 0}; ang = vel; }
        st3(s + D.s_cdofd_ang + 3 * (da + k), ang); st3(s + D.s_cdofd_vel + 3 * (da + k), vel);
      }
    } else {
      V3 ca = p >= 0 ? ld3(s + D.s_cd_ang + 3 * p) : V3{0, 0, 0};
      V3 cv = p >= 0 ? ld3(s + D.s_cd_vel + 3 * p) : V3{0, 0, 0};
      for (int k = 0; k < nd; ++k) {
        int d = da + k;
        V3 da_ = ld3(s + D.s_cdof_ang + 3 * d), dv_ = ld3(s + D.s_cdof_vel + 3 * d);
        st3(s + D.s_cdofd_vel + 3 * d, cross(ca, dv_) + cross(cv, da_));
        st3(s + D.s_cdofd_ang + 3 * d, cross(ca, da_));
        real qd = s[D.s_qd + d];
        ca = fmav(da_, qd, ca); cv = fmav(dv_, qd, cv);
      }
    }
  });
}

// --------------------------------------------------------------- mass.matrix
template <class X>
BXG_HD void mass_matrix(X& ex, const Ctx& c) {
  BXG_GET_DIMS(c); const real* mf = c.mf; const int* mi = c.mi; real* s = c.s;
  const int L = D.L, nv = D.nv, nvp = D.nvp;
  real* M = s + D.s_M;
  ex.lanes([&](int lane) {
    if (lane < L) {
      for (int i = 0; i < 3; ++i) s[D.s_crb_pos + 3 * lane + i] = s[D.s_cinr_pos + 3 * lane + i];
      for (int i = 0; i < 9; ++i) s[D.s_crb_i + 9 * lane + i] = s[D.s_cinr_i + 9 * lane + i];
      s[D.s_crb_mass + lane] = mf[D.m_in_mass + lane];
    }
    for (int i = 4 * lane; i < D.nvw * nvp; i += 4 * X::G) stv4(M + i, F4{R(0.), R(0.), R(0.), R(0.)});   // incl. padding rows (slot shared with A)
  });
  for (int lvl = D.max_depth - 1; lvl >= 0; --lvl) {
    ex.lanes([&](int l) {
      if (l >= L || mi[D.m_link_depth + l] != lvl) return;
      for (int k = mi[D.m_child_start + l]; k < mi[D.m_child_start + l + 1]; ++k) {
        int ch = mi[D.m_child_list + k];
        for (int i = 0; i < 3; ++i) s[D.s_crb_pos + 3 * l + i] += s[D.s_crb_pos + 3 * ch + i];
        for (int i = 0; i < 9; ++i) s[D.s_crb_i + 9 * l + i] += s[D.s_crb_i + 9 * ch + i];
        s[D.s_crb_mass + l] += s[D.s_crb_mass + ch];
      }
    });
  }
  // f[i] = crb[link(i)] * cdof[i] per dof, parked in the Newton-Schulz buffer (J is dead
  // between constraint.force and constraint.jacobian); then one lane per non-zero (i, j <= i)
  real* fbuf = s + D.s_B;
  ex.lanes([&](int lane) {
    for (int i = lane; i < nv; i += X::G) {
      int li = mi[D.m_dof_link + i];
      V3 fa, fv;
      inertia_mul(ld3(s + D.s_crb_pos + 3 * li), s + D.s_crb_i + 9 * li, s[D.s_crb_mass + li],
                  ld3(s + D.s_cdof_ang + 3 * i), ld3(s + D.s_cdof_vel + 3 * i), &fa, &fv);
      st3(fbuf + 6 * i, fv); st3(fbuf + 6 * i + 3, fa);
    }
  });
  ex.lanes([&](int lane) {
    for (int p = lane; p < D.n_mm_pairs; p += X::G) {
      const int ij = mi[D.m_mm_pairs + p], i = ij & 255, j = ij >> 8;
      real v = dot(ld3(s + D.s_cdof_vel + 3 * j), ld3(fbuf + 6 * i)) + dot(ld3(s + D.s_cdof_ang + 3 * j), ld3(fbuf + 6 * i + 3));
      if (i == j) v += mf[D.m_arm + i];
      M[i * nvp + j] = v;
      M[j * nvp + i] = v;
    }
  });
}

// ------------------------------------------------------ math.inv_approximate
template <class X>
BXG_HD void minv_newton_schulz_generic(X& ex, const Ctx& c, Stats* st) {
  BXG_GET_DIMS(c); real* s = c.s;
  const int n = D.nv, nvp = D.nvp;
  const real* M = s + D.s_M;
  real* Xc = s + D.s_Minv;   // current estimate
  real* Xn = s + D.s_Xn;     // candidate
  real* RB = s + D.s_B;      // residual r, then I + r in place
  typename X::LaneF p_sum, p_max;
  // r0 = I - M X
  ex.lanes([&](int lane) {
    real ss = R(0.), mx = R(0.);
    for (int i = lane; i < n; i += X::G) {
      for (int j = 0; j < n; ++j) {
        real acc = R(0.);
        for (int k = 0; k < n; ++k) acc = r_fma(M[i * nvp + k], Xc[k * nvp + j], acc);
        real r = (i == j ? R(1.) : R(0.)) - acc;
        RB[i * nvp + j] = r;
        ss = r_fma(r, r, ss); mx = r_max(mx, r_abs(r));
      }
    }
    p_sum(lane) = ss; p_max(lane) = mx;
  });
  real ss = ex.sum(p_sum), mx = ex.max(p_max);
  real nrm0 = mx <= R(1e-8) ? R(0.) : r_sqrt(ss);
  if (nrm0 > R(1.)) {
    ex.lanes([&](int lane) {
      real tr = R(0.);
      for (int i = lane; i < n; i += X::G) for (int k = 0; k < n; ++k) tr = r_fma(M[i * nvp + k], M[i * nvp + k], tr);
      p_sum(lane) = tr;
    });
    real tr = ex.sum(p_sum);
    ex.lanes([&](int lane) {
      for (int i = lane; i < n; i += X::G) for (int j = 0; j < n; ++j) Xc[i * nvp + j] = R(0.5) * M[j * nvp + i] / tr;
    });
    st->ns_cold++;
  }
  real err = R(1.);
  for (int it = 0; it < D.ns_iters; ++it) {
    // RB <- I + r ; Xn = Xc (I + r)
    ex.lanes([&](int lane) { for (int i = lane; i < n; i += X::G) RB[i * nvp + i] = R(1.) + RB[i * nvp + i]; });
    ex.lanes([&](int lane) {
      for (int i = lane; i < n; i += X::G) {
        for (int j = 0; j < n; ++j) {
          real acc = R(0.);
          for (int k = 0; k < n; ++k) acc = r_fma(Xc[i * nvp + k], RB[k * nvp + j], acc);
          Xn[i * nvp + j] = acc;
        }
      }
    });
    // r' = I - M Xn
    ex.lanes([&](int lane) {
      real s2 = R(0.), m2 = R(0.);
      for (int i = lane; i < n; i += X::G) {
        for (int j = 0; j < n; ++j) {
          real acc = R(0.);
          for (int k = 0; k < n; ++k) acc = r_fma(M[i * nvp + k], Xn[k * nvp + j], acc);
          real r = (i == j ? R(1.) : R(0.)) - acc;
          RB[i * nvp + j] = r;
          s2 = r_fma(r, r, s2); m2 = r_max(m2, r_abs(r));
        }
      }
      p_sum(lane) = s2; p_max(lane) = m2;
    });
    real s2 = ex.sum(p_sum), m2 = ex.max(p_max);
    real err_next = m2 <= R(1e-8) ? R(0.) : r_sqrt(s2);
    if (err_next < err) {
      ex.lanes([&](int lane) { for (int i = lane; i < n * nvp; i += X::G) Xc[i] = Xn[i]; });
      st->ns_accepts++;
    }
    err = err_next;
  }
}


// ===================================================== register-row kernels
// Lane i owns row i of the left operand in registers; the right operand's rows
// are broadcast from shared memory with 128-bit loads (all lanes of the group
// read the same address: one wavefront), so one LDS.128 feeds four FFMA per lane.
// acc[0..4*C4) = sum_k a[k] * B[k][0..4*C4), k ascending (same order as the oracle).
template <int K, int C4>
BXG_HD void row_times_mat(const real* a, const real* B, int ldb, real* acc) {
#pragma unroll
  for (int j = 0; j < 4 * C4; ++j) acc[j] = R(0.);
#pragma unroll
  for (int k = 0; k < K; ++k) {
    const real ak = a[k];
#pragma unroll
    for (int cc = 0; cc < C4; ++cc) {
      F4 b = ldv4(B + k * ldb + 4 * cc);
      acc[4 * cc + 0] = r_fma(ak, b.x, acc[4 * cc + 0]); acc[4 * cc + 1] = r_fma(ak, b.y, acc[4 * cc + 1]);
      acc[4 * cc + 2] = r_fma(ak, b.z, acc[4 * cc + 2]); acc[4 * cc + 3] = r_fma(ak, b.w, acc[4 * cc + 3]);
    }
  }
}
// Same product with the left row read from shared memory (own row, 128-bit
// chunks) and the k loop rolled: small code, no register-resident left operand.
template <int K, int C4>
BXG_HD void smem_row_times_mat(const real* arow_sm, const real* B, int ldb, real* acc) {
#pragma unroll
  for (int j = 0; j < 4 * C4; ++j) acc[j] = R(0.);
#pragma unroll 1
  for (int k0 = 0; k0 < K; k0 += 4) {
    const F4 a4 = ldv4(arow_sm + k0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      const real ak = kk == 0 ? a4.x : kk == 1 ? a4.y : kk == 2 ? a4.z : a4.w;
#pragma unroll
      for (int cc = 0; cc < C4; ++cc) {
        F4 b = ldv4(B + (k0 + kk) * ldb + 4 * cc);
        acc[4 * cc + 0] = r_fma(ak, b.x, acc[4 * cc + 0]); acc[4 * cc + 1] = r_fma(ak, b.y, acc[4 * cc + 1]);
        acc[4 * cc + 2] = r_fma(ak, b.z, acc[4 * cc + 2]); acc[4 * cc + 3] = r_fma(ak, b.w, acc[4 * cc + 3]);
      }
    }
  }
}
template <int C4>
BXG_HD void load_row(const real* p, real* r) {
#pragma unroll
  for (int cc = 0; cc < C4; ++cc) { F4 v = ldv4(p + 4 * cc); r[4 * cc] = v.x; r[4 * cc + 1] = v.y; r[4 * cc + 2] = v.z; r[4 * cc + 3] = v.w; }
}
template <int C4>
BXG_HD void store_row(real* p, const real* r) {
#pragma unroll
  for (int cc = 0; cc < C4; ++cc) stv4(p + 4 * cc, F4{r[4 * cc], r[4 * cc + 1], r[4 * cc + 2], r[4 * cc + 3]});
}
// dot of a register row with a shared-memory vector (broadcast 128-bit loads)
template <int C4>
BXG_HD real row_dot(const real* a, const real* v) {
  real acc = R(0.);
#pragma unroll
  for (int cc = 0; cc < C4; ++cc) {
    F4 b = ldv4(v + 4 * cc);
    acc = r_fma(a[4 * cc], b.x, acc); acc = r_fma(a[4 * cc + 1], b.y, acc); acc = r_fma(a[4 * cc + 2], b.z, acc); acc = r_fma(a[4 * cc + 3], b.w, acc);
  }
  return acc;
}

// ---- 2-D register tiles for the Newton-Schulz products -----------------------
// A G-lane group computes C[W x W] = A * B with both operands in shared memory
// (row-major, stride ld) and a TM x TN tile of C per lane.  Per 4 k-steps a lane
// issues TM 128-bit loads of A and 4*TN/2 64-bit (or TN/4 128-bit) loads of B for
// 4*TM*TN FFMA, which balances the shared-memory pipe against the FMA pipe; the
// k loop stays rolled so the body lives in the instruction cache.
template <int N> struct IntC { static constexpr int value = N; };
template <int G, int W> struct Tile;
// PIPE: request the operands of the next k-block before the FMAs of the current one, explicitly.  Pays where
// registers allow (4x4 tiles: Ant +2.4 %); with the 3x6 tiles at 96 registers it costs 5 % (profiles/r01_sweep_r1i.json)
// Lane -> tile maps (MAP) and row ownership.  Measured on B200 (tools/microbench/lds128_patterns.cu,
// profiles/r02_lds128_patterns.txt): a shared-memory load costs the LSU 4 bytes per lane and cycle (LDS.64: 2 cycles,
// LDS.128: 4) UNLESS the two lanes of every adjacent pair (2i, 2i + 1) read the same address: then half (1.05 / 2.09).
// Duplicate addresses further apart (lane l and l + 4) buy nothing, bank conflicts multiply.  So:
//   MAP_PAIR_COLS (3x6 tiles, a warp per matrix): the lanes of a pair share the COLUMN group, i.e. the B operand, whose
//     72 LDS.64 per product drop to 1 cycle each; the 18 LDS.128 of the A operand go from 2 to 4.  181 -> 148 LSU cycles
//     per product and warp (the products are LSU-bound: 20 warps x 181 = 3630 of the 3900 cycles a product round took).
//   ROWS_INTERLEAVED (4x4 tiles, two matrices per warp): row group rg owns rows rg, rg + RG, ... instead of TM consecutive
//     ones.  With consecutive rows the four row groups' A loads sit 4 rows = 80 floats = 16 banks apart: groups 0 / 2 and
//     1 / 3 and the other env's collide (4 cycles although adjacent quads share the address); interleaved they are 20 banks
//     apart: conflict-free, 2 cycles.
template <int RG_, int CG_, int TM_, int TN_, bool PIPE_, bool PAIR_COLS_ = false, bool ROWS_ILV_ = false>
struct TileBase {
  static constexpr int RG = RG_, CG = CG_, TM = TM_, TN = TN_;
  static constexpr bool PIPE = PIPE_, PAIR_COLS = PAIR_COLS_, ROWS_ILV = ROWS_ILV_;
  static_assert(!PAIR_COLS_ || (CG_ == 4 && RG_ == 8), "MAP_PAIR_COLS is written for an 8 x 4 lane grid");
  BXG_HD static int rg(int lane) { return PAIR_COLS ? ((lane & 1) | ((lane >> 3) << 1)) : lane / CG; }
  BXG_HD static int cg(int lane) { return PAIR_COLS ? ((lane >> 1) & 3) : lane - (lane / CG) * CG; }
  BXG_HD static int row(int rg_, int r) { return ROWS_ILV ? rg_ + RG * r : rg_ * TM + r; }   // r-th row of row group rg_
};
#ifndef BXG_MAP_3X6
#define BXG_MAP_3X6 true
#endif
#ifndef BXG_ILV_4X4
#define BXG_ILV_4X4 true
#endif
template <> struct Tile<32, 24> : TileBase<8, 4, 3, 6, BXG_PIPE_3X6, BXG_MAP_3X6> {};
template <> struct Tile<16, 24> : TileBase<4, 4, 6, 6, BXG_PIPE_6X6> {};
template <> struct Tile<16, 16> : TileBase<4, 4, 4, 4, BXG_PIPE_4X4, false, BXG_ILV_4X4> {};
template <> struct Tile<32, 32> : TileBase<8, 4, 4, 8, false, BXG_MAP_3X6> {};
template <> struct Tile<32, 16> : TileBase<8, 4, 2, 4, true, BXG_MAP_3X6> {};
template <> struct Tile<4, 4>   : TileBase<2, 2, 2, 2, false> {};
template <> struct Tile<4, 8>   : TileBase<2, 2, 4, 4, false> {};
struct alignas(8) F2 { real x, y; };

template <int TN>
BXG_HD void load_cols(const real* p, real* v) {
  if constexpr (TN % 4 == 0) {
#pragma unroll
    for (int c = 0; c < TN / 4; ++c) { F4 t = ldv4(p + 4 * c); v[4 * c] = t.x; v[4 * c + 1] = t.y; v[4 * c + 2] = t.z; v[4 * c + 3] = t.w; }
  } else {
#pragma unroll
    for (int c = 0; c < TN / 2; ++c) { F2 t = *reinterpret_cast<const F2*>(p + 2 * c); v[2 * c] = t.x; v[2 * c + 1] = t.y; }
  }
}
template <int TN>
BXG_HD void store_cols(real* p, const real* v) {
  if constexpr (TN % 4 == 0) {
#pragma unroll
    for (int c = 0; c < TN / 4; ++c) stv4(p + 4 * c, F4{v[4 * c], v[4 * c + 1], v[4 * c + 2], v[4 * c + 3]});
  } else {
#pragma unroll
    for (int c = 0; c < TN / 2; ++c) *reinterpret_cast<F2*>(p + 2 * c) = F2{v[2 * c], v[2 * c + 1]};
  }
}

// acc[r * TN + c] = sum_k A[row0 + r][k] * B[k][col0 + c], k ascending; NEG accumulates
// the negated products instead (rounding is symmetric: exactly -(A B), no extra operation)
// K <= W: the k range that can hold non-zeros (columns of A / rows of B past it are zero padding).  Only the
// model-specialised builds know it at compile time (K = nv): their products skip the padded k steps, exact zeros
// whose omission leaves every sum unchanged (Ant: 14 of 16, one eighth of the multiply-adds; Humanoid: 23 of 24).
template <class T, int W, bool NEG = false, int K = W>
BXG_HD void tile_matmul(int lane, const real* A, const real* B, int ld, real* acc) {
  const int rg = T::rg(lane), cg = T::cg(lane);
  const real* a0 = A + T::row(rg, 0) * ld;
  const int rs = (T::row(rg, 1) - T::row(rg, 0)) * ld;      // distance between the lane's consecutive tile rows
  constexpr int KB = (K / 4) * 4, KT = K - KB;   // k steps in full blocks of four, and in the tail block
#if defined(__CUDA_ARCH__)
  // sm_100a packed FP32: one FFMA2 = two fused multiply-adds per lane (same
  // rounding as two scalar FFMA), halving the issue slots of the inner product
  float2 acc2[T::TM][T::TN / 2];
#pragma unroll
  for (int r = 0; r < T::TM; ++r)
#pragma unroll
    for (int cc = 0; cc < T::TN / 2; ++cc) acc2[r][cc] = make_float2(R(0.), R(0.));
  if constexpr (T::PIPE) {
  // explicit software pipeline: the operands of k-block k0 + 4 are requested before the FMAs of block k0
  F4 a_nxt[T::TM]; real b_nxt[4][T::TN];
#pragma unroll
  for (int r = 0; r < T::TM; ++r) a_nxt[r] = ldv4(a0 + r * rs);
#pragma unroll
  for (int kk = 0; kk < 4; ++kk) load_cols<T::TN>(B + kk * ld + cg * T::TN, b_nxt[kk]);
  auto block = [&](const F4* a, const real (*bv)[T::TN], auto nk) {
#pragma unroll
    for (int kk = 0; kk < decltype(nk)::value; ++kk) {
#pragma unroll
      for (int r = 0; r < T::TM; ++r) {
        const real ap = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
        const real av = NEG ? -ap : ap;
        const float2 av2 = make_float2(av, av);
#pragma unroll
        for (int cc = 0; cc < T::TN / 2; ++cc)
          acc2[r][cc] = __ffma2_rn(av2, make_float2(bv[kk][2 * cc], bv[kk][2 * cc + 1]), acc2[r][cc]);
      }
    }
  };
  BXG_PRAGMA_UNROLL(BXG_PIPE16_UNROLL)
  for (int k0 = 0; k0 < KB; k0 += 4) {
    F4 a[T::TM]; real bv[4][T::TN];
#pragma unroll
    for (int r = 0; r < T::TM; ++r) a[r] = a_nxt[r];
#pragma unroll
    for (int kk = 0; kk < 4; ++kk)
#pragma unroll
      for (int cc = 0; cc < T::TN; ++cc) bv[kk][cc] = b_nxt[kk][cc];
    if (k0 + 4 < K) {
#pragma unroll
      for (int r = 0; r < T::TM; ++r) a_nxt[r] = ldv4(a0 + r * rs + k0 + 4);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) if (k0 + 4 + kk < K) load_cols<T::TN>(B + (k0 + 4 + kk) * ld + cg * T::TN, b_nxt[kk]);
    }
    block(a, bv, IntC<4>{});
  }
  if constexpr (KT > 0) block(a_nxt, b_nxt, IntC<KT>{});   // tail block: the real k steps only
  } else {
  auto block = [&](int k0, auto nk) {
    F4 a[T::TM];
#pragma unroll
    for (int r = 0; r < T::TM; ++r) a[r] = ldv4(a0 + r * rs + k0);
#pragma unroll
    for (int kk = 0; kk < decltype(nk)::value; ++kk) {
      real bv[T::TN];
      load_cols<T::TN>(B + (k0 + kk) * ld + cg * T::TN, bv);
#pragma unroll
      for (int r = 0; r < T::TM; ++r) {
        const real ap = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
        const real av = NEG ? -ap : ap;
        const float2 av2 = make_float2(av, av);
#pragma unroll
        for (int cc = 0; cc < T::TN / 2; ++cc)
          acc2[r][cc] = __ffma2_rn(av2, make_float2(bv[2 * cc], bv[2 * cc + 1]), acc2[r][cc]);
      }
    }
  };
  BXG_PRAGMA_UNROLL(BXG_TILE_UNROLL)
  for (int k0 = 0; k0 < KB; k0 += 4) block(k0, IntC<4>{});
  if constexpr (KT > 0) block(KB, IntC<KT>{});   // tail block: the real k steps only
  }
#pragma unroll
  for (int r = 0; r < T::TM; ++r)
#pragma unroll
    for (int cc = 0; cc < T::TN / 2; ++cc) { acc[r * T::TN + 2 * cc] = acc2[r][cc].x; acc[r * T::TN + 2 * cc + 1] = acc2[r][cc].y; }
#else
#pragma unroll
  for (int r = 0; r < T::TM; ++r)
#pragma unroll
    for (int cc = 0; cc < T::TN; ++cc) acc[r * T::TN + cc] = R(0.);
#pragma unroll 1
  for (int k0 = 0; k0 < K; k0 += 4) {
    F4 a[T::TM];
#pragma unroll
    for (int r = 0; r < T::TM; ++r) a[r] = ldv4(a0 + r * rs + k0);
#pragma unroll
    for (int kk = 0; kk < 4; ++kk) {
      if (k0 + kk >= K) break;
      real bv[T::TN];
      load_cols<T::TN>(B + (k0 + kk) * ld + cg * T::TN, bv);
#pragma unroll
      for (int r = 0; r < T::TM; ++r) {
        const real ap = kk == 0 ? a[r].x : kk == 1 ? a[r].y : kk == 2 ? a[r].z : a[r].w;
        const real av = NEG ? -ap : ap;
#pragma unroll
        for (int cc = 0; cc < T::TN; ++cc) acc[r * T::TN + cc] = r_fma(av, bv[cc], acc[r * T::TN + cc]);
      }
    }
  }
#endif
}

// Residual tile.  In: acc = -(M X) (tile_matmul<NEG>), i.e. already r = I - M X away
// from the diagonal.  The few lanes whose tile crosses the diagonal add the identity
// (real rows only); every lane accumulates the Frobenius partials; out: acc = I + r
// for the next product.
// calls f(element) for the tile elements of this lane that lie on the diagonal of real rows.
// The tile's diagonal offset delta = row0 - col0 is a multiple of gcd(TM, TN): only those
// offsets are compiled (two for the 3x6 tile, one for the square ones).
constexpr int tile_gcd(int a, int b) { return b == 0 ? a : tile_gcd(b, a % b); }
template <class T, class F>
BXG_HD void tile_diagonal(int lane, int n, F f) {
  const int rg = T::rg(lane), cg = T::cg(lane);
  if constexpr (T::ROWS_ILV) {
    // rows rg + RG r, columns TN cg + c with TN == RG: the lane's one diagonal element is (r, c) = (cg, rg)
    static_assert(T::TN == T::RG && T::TM == T::CG, "interleaved rows: square lane grid and tile");
    const int row = rg + T::RG * cg;
#pragma unroll
    for (int r = 0; r < T::TM; ++r)
#pragma unroll
      for (int cc = 0; cc < T::TN; ++cc) if (r == cg && cc == rg && row < n) f(r * T::TN + cc);
  } else {
  constexpr int S = tile_gcd(T::TM, T::TN);
  const int row0 = rg * T::TM, delta = row0 - cg * T::TN;     // diagonal elements: cc = r + delta
  if (delta > -T::TM && delta < T::TN) {
#pragma unroll
    for (int dlt = -((T::TM - 1) / S) * S; dlt < T::TN; dlt += S) {
      if (delta == dlt) {
#pragma unroll
        for (int r = 0; r < T::TM; ++r) {
          if (r + dlt >= 0 && r + dlt < T::TN) { if (row0 + r < n) f(r * T::TN + r + dlt); }
        }
      }
    }
  }
  }
}
template <class T>
BXG_HD void residual_tile(int lane, int n, real* acc, real* ss, real* mx) {
  tile_diagonal<T>(lane, n, [&](int e) { acc[e] = R(1.) + acc[e]; });     // r = I - M X
#pragma unroll
  for (int e = 0; e < T::TM * T::TN; ++e) { *ss = r_fma(acc[e], acc[e], *ss); *mx = r_max(*mx, r_abs(acc[e])); }
  tile_diagonal<T>(lane, n, [&](int e) { acc[e] = R(1.) + acc[e]; });     // I + r
}
template <class T>
BXG_HD void store_tile(int lane, real* C, int ld, const real* acc) {
  const int rg = T::rg(lane), cg = T::cg(lane);
#pragma unroll
  for (int r = 0; r < T::TM; ++r) store_cols<T::TN>(C + T::row(rg, r) * ld + cg * T::TN, acc + r * T::TN);
}

// math.inv_approximate (brax/math.py:278-305) on W x W zero-padded matrices.
// Shared memory holds M, the current estimate X and ONE more buffer Q.  A product's
// result tile waits in registers until every lane has finished reading the operand
// it replaces: the candidate X (I + r) overwrites I + r (dead once the product is
// formed: the next residual is computed from M and the candidate), and the next
// I + r' overwrites whichever of X / candidate lost the comparison.
template <class X, int W>
BXG_HD void minv_newton_schulz_tiles(X& ex, const Ctx& c, Stats* st) {
  using T = Tile<X::G, W>;
  BXG_GET_DIMS(c); real* s = c.s;
  const int n = D.nv, ld = D.nvp;
#if defined(BXG_CONST_DIMS)
  constexpr int K = D.nv;      // model-specialised build: the padded k steps are skipped (tile_matmul)
#else
  constexpr int K = W;
#endif
  const real* M = s + D.s_M;
  real* Xa = s + D.s_Minv;
  real* Xc = Xa;                      // current estimate
  real* Q = s + D.s_B;                // I + r, then the candidate
  typename X::LaneF p_sum, p_max;
  typename X::template LaneVec<T::TM * T::TN> tile;
  // r0 = I - M X
  ex.lanes([&](int lane) {
    real* acc = tile(lane); real ss = R(0.), mx = R(0.);
    tile_matmul<T, W, true, K>(lane, M, Xc, ld, acc);
    residual_tile<T>(lane, n, acc, &ss, &mx);
    store_tile<T>(lane, Q, ld, acc);
    p_sum(lane) = ss; p_max(lane) = mx;
  });
  real ss0, mx0;
  ex.sum_max(p_sum, p_max, &ss0, &mx0);
  real nrm0 = mx0 <= R(1e-8) ? R(0.) : r_sqrt(ss0);
  if (nrm0 > R(1.)) {
    // cold start 0.5 M^T / tr(M M^T); M is exactly symmetric
    ex.lanes([&](int lane) {
      real tr = R(0.);
      for (int i = lane; i < W * ld; i += X::G) tr += M[i] * M[i];
      p_sum(lane) = tr;
    });
    const real tr = ex.sum(p_sum);
    // x / tr with one shared reciprocal: q = x * (1/tr) corrected by its own remainder
    // (the correctly rounded quotient whenever 1/tr is; three operations per element
    // instead of a division sequence each)
    const real rinv = R(1.) / tr;
    ex.lanes([&](int lane) {
      for (int i = 4 * lane; i < W * ld; i += 4 * X::G) {
        F4 m = ldv4(M + i);
        real x[4] = {R(0.5) * m.x, R(0.5) * m.y, R(0.5) * m.z, R(0.5) * m.w}, q[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { q[u] = x[u] * rinv; q[u] = r_fma(r_fma(-tr, q[u], x[u]), rinv, q[u]); }
        stv4(Xc + i, F4{q[0], q[1], q[2], q[3]});
      }
    });
    st->ns_cold++;
  }
  real err = R(1.);
  for (int it = 0; it < D.ns_iters; ++it) {
    ex.lanes([&](int lane) { tile_matmul<T, W, false, K>(lane, Xc, Q, ld, tile(lane)); });   // candidate = X (I + r)
    ex.lanes([&](int lane) { store_tile<T>(lane, Q, ld, tile(lane)); });           // ... replaces I + r
    ex.lanes([&](int lane) {       // r' = I - M candidate
      real* acc = tile(lane); real ss = R(0.), mx = R(0.);
      tile_matmul<T, W, true, K>(lane, M, Q, ld, acc);
      residual_tile<T>(lane, n, acc, &ss, &mx);
      p_sum(lane) = ss; p_max(lane) = mx;
    });
    real s2, m2;
    ex.sum_max(p_sum, p_max, &s2, &m2);
    real err_next = m2 <= R(1e-8) ? R(0.) : r_sqrt(s2);
    const bool accept = err_next < err;
    real* dst = accept ? Xc : Q;     // I + r' replaces the loser
    ex.lanes([&](int lane) { store_tile<T>(lane, dst, ld, tile(lane)); });
    if (accept) { real* t = Xc; Xc = Q; Q = t; st->ns_accepts++; }
    err = err_next;
  }
  if (Xc != Xa) ex.lanes([&](int lane) { for (int i = lane; i < W * ld; i += X::G) Xa[i] = Xc[i]; });
}

template <class X, class Cfg>
BXG_HD void minv_newton_schulz(X& ex, const Ctx& c, Stats* st) {
  if constexpr (Cfg::VC4 > 0) {
    if (!Cfg::GENERIC_TOO || !BXG_DIMS_OF(c).force_generic) { minv_newton_schulz_tiles<X, 4 * Cfg::VC4>(ex, c, st); return; }
  }
  if constexpr (Cfg::GENERIC_TOO) minv_newton_schulz_generic(ex, c, st);
}

// dot of a register row with a shared-memory vector, first `chunks` float4 chunks
template <int C4>
BXG_HD real row_dot_n(const real* a, const real* v, int chunks) {
  real acc = R(0.);
#pragma unroll
  for (int cc = 0; cc < C4; ++cc) {
    if (cc < chunks) {
      F4 b = ldv4(v + 4 * cc);
      acc = r_fma(a[4 * cc], b.x, acc); acc = r_fma(a[4 * cc + 1], b.y, acc); acc = r_fma(a[4 * cc + 2], b.z, acc); acc = r_fma(a[4 * cc + 3], b.w, acc);
    }
  }
  return acc;
}
// index of the n-th (0-based) set bit of m
BXG_HD int nth_set_bit(uint64_t m, int n) {
  for (int k = 0; k < n; ++k) m &= m - 1ull;
#if defined(__CUDA_ARCH__)
  return __ffsll((long long)m) - 1;
#else
  return __builtin_ffsll((long long)m) - 1;
#endif
}
// dot of a row of A kept in shared memory (128-bit loads of the lane's own row) with a vector
BXG_HD real smem_row_dot(const real* arow_sm, const real* v, int chunks) {
  real acc = R(0.);
  for (int cc = 0; cc < chunks; ++cc) {
    F4 a = ldv4(arow_sm + 4 * cc), b = ldv4(v + 4 * cc);
    acc = r_fma(a.x, b.x, acc); acc = r_fma(a.y, b.y, acc); acc = r_fma(a.z, b.z, acc); acc = r_fma(a.w, b.w, acc);
  }
  return acc;
}

// constraint.force on the ACTIVE rows only.  Rows the jacobian masked out
// (inactive contact pyramids, joints inside their limits: constraint.py:124-131,174)
// are identically zero in J, diag and aref, so they contribute exact zeros to A,
// b, the objective, the gradient and J^T x and their x stays 0: dropping them
// leaves every sum unchanged.  The na active rows are compacted to rows
// 0..na-1 (lane p owns compact rows p and p+G): A = (J Minv) J^T + diag is built
// row by row against the active rows of J (no transposed copy), the row of A then
// stays in registers for FISTA + backtracking line search as in
// jaxopt.ProjectedGradient (see oracle/bxg_oracle.c).  VC4 = nvw/4, NC4 = ncw/4.
// Wide variants (more than 56 row registers per lane) read their row of A from shared
// memory instead (AREG false).
template <class X, int VC4, int NC4, int R>
BXG_HD void con_force_rows(X& ex, const Ctx& c, Stats* st) {
  constexpr int VW = 4 * VC4, CW = 4 * NC4, G = X::G;
  constexpr bool AREG = R * CW <= 56;
  BXG_GET_DIMS(c); real* s = c.s;
  const int nv = D.nv, nc = D.nc, ldv = D.nvp, ldj = D.jld, ldc = D.ncp;
  const real* J = s + D.s_J; const real* Mi = s + D.s_Minv;
  real* A = s + D.s_A;
  real* xs = s + D.s_px; real* ys = s + D.s_py; real* ress = s + D.s_pres; real* xns = s + D.s_pxn;
  int* orig = reinterpret_cast<int*>(s + D.s_pg);   // compact row -> row of J
  typename X::template LaneVec<AREG ? R * CW : 1> arow;
  typename X::template LaneVec<R> bi, xi, yi, gi, xni, resi;
  typename X::LaneF p0, p1, p2;
  // ---- active set (rows 0..63 in am, rows 64.. in am_hi for the widest variant) -------------
  uint64_t am = 0, am_hi = 0;
#pragma unroll
  for (int r = 0; r < R; ++r) {
    ex.lanes([&](int lane) {
      int i = lane + r * G;
      p0(lane) = i < nc && row_active(s, D)[i] ? R(1.) : R(0.);   // set by constraint.jacobian (load_env for the incoming state)
    });
    if (r * G < 64) am |= (uint64_t)ex.ballot(p0) << (r * G);
    else am_hi |= (uint64_t)ex.ballot(p0) << (r * G - 64);
  }
#if defined(__CUDA_ARCH__)
  const int na_lo = __popcll(am), na = na_lo + (CW > 64 ? __popcll(am_hi) : 0);
#else
  const int na_lo = __builtin_popcountll(am), na = na_lo + (CW > 64 ? __builtin_popcountll(am_hi) : 0);
#endif
  if (na == 0) {   // nothing active: the solver would return x = 0 after one trivial iteration
    ex.lanes([&](int lane) { for (int d = lane; d < nv; d += G) s[D.s_qfc + d] = R(0.); });
    st->pg_iters++; st->pg_trials++;
    return;
  }
  const int nch = (na + 3) >> 2;           // float4 chunks that hold active columns
  ex.lanes([&](int lane) {
    for (int i = lane; i < CW; i += G) { xs[i] = R(0.); ys[i] = R(0.); xns[i] = R(0.); ress[i] = R(0.); }
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int p = lane + r * G;
      if (p < CW) orig[p] = p < na ? (CW > 64 && p >= na_lo ? 64 + nth_set_bit(am_hi, p - na_lo) : nth_set_bit(am, p)) : -1;
    }
  });
  // b = (J Minv) qf_smooth - aref;  A[p, q] = (J Minv)[orig p, :] . J[orig q, :] + diag,
  // four active columns at a time (rows of J are broadcast 128-bit loads feeding four
  // independent k-ascending chains).  The row goes to shared memory for the column
  // reads of the gradient and stays in registers for the row products.
  ex.lanes([&](int lane) {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      int p = lane + r * G;
      real* ar = arow(lane) + (AREG ? r * CW : 0);
      if constexpr (AREG) {
#pragma unroll
        for (int j = 0; j < CW; ++j) ar[j] = R(0.);
      }
      bi(lane)[r] = R(0.);
      if (p < na) {
        const int i = orig[p];
        real jm[VW];
        smem_row_times_mat<VW, VC4>(J + i * ldj, Mi, ldv, jm);        // (J Minv)[i, :]
        real bacc = R(0.);
#pragma unroll
        for (int cc = 0; cc < VC4; ++cc) {
          F4 f = ldv4(s + D.s_qfs + 4 * cc);
          bacc = r_fma(jm[4 * cc], f.x, bacc); bacc = r_fma(jm[4 * cc + 1], f.y, bacc); bacc = r_fma(jm[4 * cc + 2], f.z, bacc); bacc = r_fma(jm[4 * cc + 3], f.w, bacc);
        }
        bi(lane)[r] = bacc - s[D.s_aref + i];
        const real dg = s[D.s_diag + i];
        real* arow_sm = A + p * ldc;
#pragma unroll 1
        for (int qc = 0; qc < nch; ++qc) {
          real a4[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const int q = 4 * qc + u, iq = orig[q];
            const real* jq = J + (iq < 0 ? i : iq) * ldj;
            real acc = R(0.);
#pragma unroll
            for (int cc = 0; cc < VC4; ++cc) {
              F4 b = ldv4(jq + 4 * cc);
              acc = r_fma(jm[4 * cc], b.x, acc); acc = r_fma(jm[4 * cc + 1], b.y, acc); acc = r_fma(jm[4 * cc + 2], b.z, acc); acc = r_fma(jm[4 * cc + 3], b.w, acc);
            }
            if (q == p) acc += dg;
            a4[u] = iq < 0 ? R(0.) : acc;
          }
          stv4(arow_sm + 4 * qc, F4{a4[0], a4[1], a4[2], a4[3]});
        }
        if constexpr (AREG) {
#pragma unroll
          for (int cc = 0; cc < NC4; ++cc) {
            if (cc < nch) { F4 v = ldv4(arow_sm + 4 * cc); ar[4 * cc] = v.x; ar[4 * cc + 1] = v.y; ar[4 * cc + 2] = v.z; ar[4 * cc + 3] = v.w; }
          }
        }
      }
      xi(lane)[r] = R(0.); yi(lane)[r] = R(0.); gi(lane)[r] = R(0.); xni(lane)[r] = R(0.); resi(lane)[r] = R(0.);
    }
  });
  BXG_PHASE_END(st, 13);   // (tuning builds) active set + A, b built
  real t = R(1.), stepsize = R(1.), error = INFINITY;
  const real tol = R(1e-3), eps = r_eps();
  int it = 0;
  while (it < D.solver_iterations && (it == 0 || error > tol)) {
    // value and gradient at y
    ex.lanes([&](int lane) {
      real f = R(0.);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        int p = lane + r * G;
        if (p < na) {
          real rv = (AREG ? row_dot_n<NC4>(arow(lane) + (AREG ? r * CW : 0), ys, nch) : smem_row_dot(A + p * ldc, ys, nch)) + bi(lane)[r];
          ress[p] = rv;
          f = r_fma(R(0.5), (rv * rv), f);
        }
      }
      p0(lane) = f;
    });
    real fy = ex.sum(p0);
    ex.lanes([&](int lane) {
#pragma unroll
      for (int r = 0; r < R; ++r) {
        int j = lane + r * G;
        if (j < na) {
          real acc = R(0.);
#pragma unroll 4
          for (int i = 0; i < na; ++i) acc = r_fma(A[i * ldc + j], ress[i], acc);
          gi(lane)[r] = acc;
        }
      }
    });
    BXG_PHASE_END(st, 14);
    real sz = stepsize;
    for (int ls = 0;; ++ls) {
      real sqdist, vd, fn;
#if defined(BXG_LS_REPEAT)
#pragma unroll 1
      for (int rep = 0; rep < BXG_LS_REPEAT; ++rep) {     // tuning builds: every trial evaluated BXG_LS_REPEAT times (same result): what a trial costs
#endif
      ex.lanes([&](int lane) {
#pragma unroll
        for (int r = 0; r < R; ++r) {
          int p = lane + r * G;
          if (p < na) { real v = r_max(yi(lane)[r] - sz * gi(lane)[r], R(0.)); xni(lane)[r] = v; xns[p] = v; }
        }
      });
      ex.lanes([&](int lane) {
        real a0 = R(0.), a1 = R(0.), a2 = R(0.);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          int p = lane + r * G;
          if (p < na) {
            real rv = (AREG ? row_dot_n<NC4>(arow(lane) + (AREG ? r * CW : 0), xns, nch) : smem_row_dot(A + p * ldc, xns, nch)) + bi(lane)[r];
            resi(lane)[r] = rv;
            real dlt = xni(lane)[r] - yi(lane)[r];
            a0 = r_fma(dlt, dlt, a0); a1 = r_fma(dlt, gi(lane)[r], a1); a2 = r_fma(R(0.5), (rv * rv), a2);
          }
        }
        p0(lane) = a0; p1(lane) = a1; p2(lane) = a2;
      });
      ex.sum3(p0, p1, p2, &sqdist, &vd, &fn);
#if defined(BXG_LS_REPEAT)
      }
#endif
      st->pg_trials++;
      real fun_decrease = sz * (fn - fy);
      real condition = sz * vd + R(0.5) * sqdist;
      if (!(fun_decrease > condition + eps) || ls >= D.solver_maxls) break;
      sz = sz * R(0.5);
    }
    BXG_PHASE_END(st, 15);   // (tuning builds) the line-search trials
    stepsize = sz <= R(1e-6) ? R(1.) : sz / R(0.5);
    real tn = R(0.5) * (R(1.) + r_sqrt(R(1.) + R(4.) * (t * t)));
    real mom = (t - R(1.)) / tn;
    ex.lanes([&](int lane) {
#pragma unroll
      for (int r = 0; r < R; ++r) { int p = lane + r * G; if (p < na) ress[p] = resi(lane)[r]; }
    });
    ex.lanes([&](int lane) {
      real e2 = R(0.);
#pragma unroll
      for (int r = 0; r < R; ++r) {
        int j = lane + r * G;
        if (j < na) {
          real acc = R(0.);
#pragma unroll 4
          for (int i = 0; i < na; ++i) acc = r_fma(A[i * ldc + j], ress[i], acc);
          real xn = xni(lane)[r];
          real dlt = r_max(xn - acc, R(0.)) - xn;
          e2 = r_fma(dlt, dlt, e2);
          real yv = xn + mom * (xn - xi(lane)[r]);
          yi(lane)[r] = yv; ys[j] = yv;
          xi(lane)[r] = xn; xs[j] = xn;
        }
      }
      p0(lane) = e2;
    });
    error = r_sqrt(ex.sum(p0));
    t = tn;
    ++it;
    st->pg_iters++;
  }
  BXG_PHASE_END(st, 14);   // (tuning builds) FISTA iterations + line searches
  // qf_constraint = J^T x over the active rows (lanes read consecutive columns of J)
  ex.lanes([&](int lane) {
    for (int j = lane; j < nv; j += G) {
      real acc = R(0.);
#pragma unroll 4
      for (int p = 0; p < na; ++p) acc = r_fma(J[orig[p] * ldj + j], xs[p], acc);
      s[D.s_qfc + j] = acc;
    }
  });
}

template <class X, class Cfg>
BXG_HD void con_force(X& ex, const Ctx& c, Stats* st) {
  BXG_GET_DIMS(c);
  if (D.nc == 0) {
    ex.lanes([&](int lane) { for (int d = lane; d < D.nv; d += X::G) c.s[D.s_qfc + d] = R(0.); });
    return;
  }
  if constexpr (Cfg::VC4 > 0 && Cfg::NC4 > 0) {
    if (!Cfg::GENERIC_TOO || !D.force_generic) {
      con_force_rows<X, Cfg::VC4, Cfg::NC4, (4 * Cfg::NC4 + X::G - 1) / X::G>(ex, c, st);
      return;
    }
  }
  if constexpr (Cfg::GENERIC_TOO) con_force_generic(ex, c, st);
}

// One end of a capsule against a plane (mjx plane_capsule, restated in
// oracle/bxg_oracle.c): the end sphere's centre is the geom centre moved along the
// capsule axis (third column of the geom's world rotation, brax/contact.py:48-53 with
// math.quat_to_3x3), and the contact frame's first tangent follows the axis projected
// into the plane (falls back to y or z when the capsule stands upright).
BXG_HD void capsule_end(Q4 link_rot, Q4 geom_quat, real half_len, V3 n, V3* centre, V3* t1, V3* t2) {
  Q4 q = qmul(link_rot, geom_quat);
  real dq = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z, sc = R(2.) / dq;
  real xs = q.x * sc, ys = q.y * sc, zs = q.z * sc;
  V3 axis{q.x * zs + q.w * ys, q.y * zs - q.w * xs, R(1.) - (q.x * xs + q.y * ys)};
  *centre = *centre + axis * half_len;
  real na = dot(n, axis);
  V3 b{axis.x - n.x * na, axis.y - n.y * na, axis.z - n.z * na};
  bool zero = r_abs(b.x) <= R(1e-8) && r_abs(b.y) <= R(1e-8) && r_abs(b.z) <= R(1e-8);     // math.safe_norm
  real bn = zero ? R(0.) : r_sqrt(b.x * b.x + b.y * b.y + b.z * b.z);
  real d = bn + R(1e-6) * (bn == R(0.) ? R(1.) : R(0.));
  b = V3{b.x / d, b.y / d, b.z / d};
  if (bn < R(0.5)) b = (-R(0.5) < n.y && n.y < R(0.5)) ? V3{R(0.), R(1.), R(0.)} : V3{R(0.), R(0.), R(1.)};
  *t1 = b;
  *t2 = cross(n, b);
}

// z column of math.quat_to_3x3(q): the axis of a capsule whose world orientation is q
BXG_HD V3 capsule_axis(Q4 q) {
  real dq = q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z, sc = R(2.) / dq;
  real xs = q.x * sc, ys = q.y * sc, zs = q.z * sc;
  return V3{q.x * zs + q.w * ys, q.y * zs - q.w * xs, R(1.) - (q.x * xs + q.y * ys)};
}
// mjx math.normalize_with_norm: x / (n + 1e-6 * (n == 0)), n the safe norm
BXG_HD V3 normalize_with_norm(V3 v, real* norm) {
  bool zero = r_abs(v.x) <= R(1e-8) && r_abs(v.y) <= R(1e-8) && r_abs(v.z) <= R(1e-8);
  real n = zero ? R(0.) : r_sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
  real d = n + R(1e-6) * (n == R(0.) ? R(1.) : R(0.));
  *norm = n;
  return V3{v.x / d, v.y / d, v.z / d};
}
// mjx math.closest_segment_point
BXG_HD V3 closest_segment_point(V3 a, V3 b, V3 pt) {
  V3 ab = b - a;
  real t = dot(pt - a, ab) / (dot(ab, ab) + R(1e-6));
  t = r_max(R(0.), r_min(t, R(1.)));
  return a + ab * t;
}
// mjx collision_primitive.capsule_capsule: closest points of the two segments
// (math.closest_segment_to_segment_points), then _sphere_sphere; frame = math.make_frame(n).
// No reference test pins capsule-capsule numbers (parity unpinned for this pair type).
BXG_HD_NOINLINE void capsule_capsule(V3 ca, V3 xa, real half_a, real rad_a, V3 cb, V3 xb, real half_b, real rad_b,
                                     real* dist, V3* pos, V3* n_out, V3* t1, V3* t2) {
  V3 a0 = ca - xa * half_a, a1 = ca + xa * half_a, b0 = cb - xb * half_b, b1 = cb + xb * half_b;
  real len_a, len_b;
  V3 dir_a = normalize_with_norm(a1 - a0, &len_a), dir_b = normalize_with_norm(b1 - b0, &len_b);
  real ha = len_a * R(0.5), hb = len_b * R(0.5);
  V3 a_mid = a0 + dir_a * ha, b_mid = b0 + dir_b * hb;
  V3 trans = a_mid - b_mid;
  real dab = dot(dir_a, dir_b), dat = dot(dir_a, trans), dbt = dot(dir_b, trans);
  real denom = R(1.) - dab * dab;
  real orig_t_a = (-dat + dab * dbt) / (denom + R(1e-6));
  real orig_t_b = dbt + orig_t_a * dab;
  real t_a = r_max(-ha, r_min(orig_t_a, ha)), t_b = r_max(-hb, r_min(orig_t_b, hb));
  V3 best_a = a_mid + dir_a * t_a, best_b = b_mid + dir_b * t_b;
  V3 new_a = closest_segment_point(a0, a1, best_b), new_b = closest_segment_point(b0, b1, best_a);
  V3 u = new_a - best_b, v = best_a - new_b;
  if (dot(u, u) < dot(v, v)) best_a = new_a; else best_b = new_b;
  real d;
  V3 n = normalize_with_norm(best_b - best_a, &d);
  if (d == R(0.)) n = V3{R(1.), R(0.), R(0.)};
  d = d - (rad_a + rad_b);
  *dist = d;
  *pos = best_a + n * (rad_a + d * R(0.5));
  real nn;
  V3 a = normalize_with_norm(n, &nn);
  V3 b = (-R(0.5) < a.y && a.y < R(0.5)) ? V3{R(0.), R(1.), R(0.)} : V3{R(0.), R(0.), R(1.)};
  real ab = dot(a, b);
  b = normalize_with_norm(V3{b.x - a.x * ab, b.y - a.y * ab, b.z - a.z * ab}, &nn);
  *n_out = a; *t1 = b; *t2 = cross(a, b);
}

// ------------------------------------------------------- constraint.jacobian
template <class X, class Cfg>
BXG_HD void con_jacobian(X& ex, const Ctx& c) {
  // contacts between two moving links (capsule-capsule) are compiled into variant 5 and the
  // generic variant only (bxg_model.h picks one of them for such models)
  constexpr bool kTwoBody = Cfg::VC4 == 0 || Cfg::NC4 == 16;
  BXG_GET_DIMS(c); const real* mf = c.mf; const int* mi = c.mi; real* s = c.s;
  const int nv = D.nv, nvp = D.jld;
  real* J = s + D.s_J;
  // J shares its slot with Newton-Schulz scratch: clear it (limit rows and the
  // padding columns rely on zeros)
  ex.lanes([&](int lane) { for (int i = lane; i < D.nc * nvp; i += X::G) J[i] = R(0.); });
  for (int cc = 0; cc < D.ncon; ++cc) {
    int lb = mi[D.m_con_lb + cc];
    // contact.get, plane-sphere / plane-capsule (contact.py:28-67 + mjx): every lane redundantly
    V3 n = ld3(mf + D.m_con_frame + 9 * cc), t1 = ld3(mf + D.m_con_frame + 9 * cc + 3), t2 = ld3(mf + D.m_con_frame + 9 * cc + 6);
    real rad = mf[D.m_con_rad + cc], mu = mf[D.m_con_mu + cc];
    V3 sp = ld3(s + D.s_x_pos + 3 * lb) + rotate(ld3(mf + D.m_con_spos + 3 * cc), ld4(s + D.s_x_rot + 4 * lb));
    if (mi[D.m_con_kind + cc] == BXG_CON_PLANE_CAPSULE_END) capsule_end(ld4(s + D.s_x_rot + 4 * lb), ld4(mf + D.m_con_gquat + 4 * cc), mf[D.m_con_half + cc], n, &sp, &t1, &t2);
    real dist = dot(sp - ld3(mf + D.m_con_ppos + 3 * cc), n) - rad;
    V3 pos = sp - n * (rad + R(0.5) * dist);
    int la = -1; uint32_t alo = 0u, ahi = 0u;
    if constexpr (kTwoBody) {
      if (mi[D.m_con_kind + cc] == BXG_CON_CAPSULE_CAPSULE) {
        la = mi[D.m_con_la + cc]; alo = (uint32_t)mi[D.m_con_anca_lo + cc]; ahi = (uint32_t)mi[D.m_con_anca_hi + cc];
        V3 ca = ld3(mf + D.m_con_apos + 3 * cc); Q4 qa = ld4(mf + D.m_con_aquat + 4 * cc), qb = ld4(mf + D.m_con_gquat + 4 * cc);
        if (la >= 0) { ca = ld3(s + D.s_x_pos + 3 * la) + rotate(ca, ld4(s + D.s_x_rot + 4 * la)); qa = qmul(ld4(s + D.s_x_rot + 4 * la), qa); }
        qb = qmul(ld4(s + D.s_x_rot + 4 * lb), qb);
        capsule_capsule(ca, capsule_axis(qa), mf[D.m_con_ahalf + cc], mf[D.m_con_arad + cc], sp, capsule_axis(qb), mf[D.m_con_half + cc], rad,
                        &dist, &pos, &n, &t1, &t2);
      }
    }
    bool active = dist < R(0.);
    V3 off = pos - ld3(s + D.s_root_com + 3 * lb);
    V3 off_a = la >= 0 ? pos - ld3(s + D.s_root_com + 3 * la) : V3{R(0.), R(0.), R(0.)};
    uint32_t lo = (uint32_t)mi[D.m_con_anc_lo + cc], hi = (uint32_t)mi[D.m_con_anc_hi + cc];
    V3 dir[4];
    for (int k = 0; k < 4; ++k) {
      V3 tt = k < 2 ? t1 : t2; real f = (k & 1) ? mu : -mu;
      dir[k] = V3{(-tt.x) * f + n.x, (-tt.y) * f + n.y, (-tt.z) * f + n.z};
    }
    ex.lanes([&](int lane) {
      if (lane == 0) s[D.s_dist + cc] = dist;
      for (int d = lane; d < nv; d += X::G) {
        uint32_t bit = d < 32 ? (lo >> d) & 1u : (hi >> (d - 32)) & 1u;
        real r0 = R(0.), r1 = R(0.), r2 = R(0.), r3 = R(0.);
        if (active && bit) {
          V3 a = ld3(s + D.s_cdof_ang + 3 * d), v = ld3(s + D.s_cdof_vel + 3 * d);
          V3 df = v - cross(off, a);
          r0 = dot(df, dir[0]); r1 = dot(df, dir[1]); r2 = dot(df, dir[2]); r3 = dot(df, dir[3]);
        }
        if constexpr (kTwoBody) {   // diff = J_b.vel - J_a.vel (constraint.py:158)
          uint32_t abit = d < 32 ? (alo >> d) & 1u : (ahi >> (d - 32)) & 1u;
          if (active && abit) {
            V3 a = ld3(s + D.s_cdof_ang + 3 * d), v = ld3(s + D.s_cdof_vel + 3 * d);
            V3 ja = v - cross(off_a, a);
            V3 jb = bit ? v - cross(off, a) : V3{R(0.), R(0.), R(0.)};
            V3 df = jb - ja;
            r0 = dot(df, dir[0]); r1 = dot(df, dir[1]); r2 = dot(df, dir[2]); r3 = dot(df, dir[3]);
          }
        }
        J[(4 * cc + 0) * nvp + d] = r0; J[(4 * cc + 1) * nvp + d] = r1;
        J[(4 * cc + 2) * nvp + d] = r2; J[(4 * cc + 3) * nvp + d] = r3;
      }
    });
    ex.lanes([&](int lane) {
      if (lane >= 4) return;
      int row = 4 * cc + lane;
      real diag = R(0.), aref = R(0.);
      if (active) {
        real vel = R(0.);
        for (int d = 0; d < nv; ++d) vel = r_fma(J[row * nvp + d], s[D.s_qd + d], vel);
        real imp;
        int spi = cc;
        if constexpr (Cfg::NC4 >= 20) spi = mi[D.m_con_sp_idx + cc];   // 80-row variant: distinct parameter sets stored once
        imp_aref(mf + D.m_con_sp + kImpStride * spi, dist, vel, &imp, &aref);
        real tw = mf[D.m_link_invw + lb];  // link_a is the world: contributes 0
        if constexpr (kTwoBody) { if (la >= 0) tw = mf[D.m_link_invw + la] + mf[D.m_link_invw + lb]; }   // invweight[a] * (a > -1) + invweight[b]
        diag = (tw + mu * mu * tw) * (R(2.) * mu * mu * (R(1.) - imp) / (imp + R(1e-8)));
      }
      s[D.s_diag + row] = diag; s[D.s_aref + row] = aref; row_active(s, D)[row] = active ? 1 : 0;
    });
  }
  if (D.nlim > 0) {
    ex.lanes([&](int lane) {
      for (int r = lane; r < D.nlim; r += X::G) {
        int d = mi[D.m_lim_dof + r], row = 4 * D.ncon + r;
        real q = s[D.s_q + mi[D.m_dof_qidx + d]];
        real pos_min = q - mf[D.m_lim_lo + d], pos_max = mf[D.m_lim_hi + d] - q;
        real pos = r_min(r_min(pos_min, pos_max), R(0.));
        bool active = pos < R(0.);
        real side = active ? (pos_min < pos_max ? R(1.) : -R(1.)) : R(0.);
        real diag = R(0.), aref = R(0.);
        if (active) {
          real imp;
          int spi = d;
          if constexpr (Cfg::NC4 >= 20) spi = mi[D.m_dof_sp_idx + d];
          imp_aref(mf + D.m_dof_sp + kImpStride * spi, pos, side * s[D.s_qd + d], &imp, &aref);
          diag = mf[D.m_dof_invw + d] * (R(1.) - imp) / (imp + R(1e-8));
        }
        J[row * nvp + d] = side;
        s[D.s_diag + row] = diag; s[D.s_aref + row] = aref; row_active(s, D)[row] = active ? 1 : 0;
      }
    });
  }
}

// ------------------------------------------------------------------ pipeline
// INV selects, at compile time, how mass_mx_inv is produced, so that each kernel carries
// only one of the two code paths:
//   0  the reference's Newton-Schulz iteration (or its exact solve when matrix_inv_iterations == 0)
//   1  exact SPD inverse by Cholesky (pipeline.init; BXG_MINV_CHOLESKY steps)
template <class X, class Cfg, int INV>
BXG_HD void update_position_terms(X& ex, const Ctx& c, Stats* st, bool in_step) {
  const int sl = in_step ? BXG_DIMS_OF(c).sync_level : 0;
  kinematics<X, !BXG_XD_TAIL>(ex, c);
  BXG_PHASE_END(st, 4);
  transform_com(ex, c);
  BXG_PHASE_END(st, 5);
  if (sl & 8) ex.cta_sync();
  mass_matrix(ex, c);
  BXG_PHASE_END(st, 6);
  if (sl & 16) ex.cta_sync();
  if constexpr (INV == 1) {
    bool done = false;
    if constexpr (Cfg::VC4 > 0 && 4 * Cfg::VC4 <= X::G) {   // one lane per row / column
      if (!Cfg::GENERIC_TOO || !BXG_DIMS_OF(c).force_generic) { spd_inverse_rows<X, 4 * Cfg::VC4>(ex, c, c.s + BXG_DIMS_OF(c).s_M, c.s + BXG_DIMS_OF(c).s_Minv, c.s + BXG_DIMS_OF(c).s_scr); done = true; }
    }
    if (!done) spd_inverse(ex, c, c.s + BXG_DIMS_OF(c).s_M, c.s + BXG_DIMS_OF(c).s_Minv, c.s + BXG_DIMS_OF(c).s_scr, nullptr, R(0.));
  } else {
    if (BXG_DIMS_OF(c).ns_iters == 0) spd_inverse(ex, c, c.s + BXG_DIMS_OF(c).s_M, c.s + BXG_DIMS_OF(c).s_Minv, c.s + BXG_DIMS_OF(c).s_scr, nullptr, R(0.));
    else minv_newton_schulz<X, Cfg>(ex, c, st);
  }
  BXG_PHASE_END(st, 7);
  if (sl & 2) ex.cta_sync();
  con_jacobian<X, Cfg>(ex, c);
  BXG_PHASE_END(st, 8);
}

// pipeline.step (pipeline.py:78-94)
template <class X, class Cfg, int INV>
BXG_HD void substep(X& ex, const Ctx& c, Stats* st) {
  // CTA-wide phase alignment keeps the warps of a CTA on the same straight-line
  // code (instruction-cache locality); sync_level trades that against barrier waits
  const int sl = BXG_DIMS_OF(c).sync_level;
  BXG_PHASE_BEGIN(st);
  if (sl & 4) ex.cta_sync();
  dyn_forces<X, Cfg>(ex, c);
  BXG_PHASE_END(st, 1);
  if (sl & 32) ex.cta_sync();
  con_force<X, Cfg>(ex, c, st);
  BXG_PHASE_END(st, 2);
  if (sl & 1) ex.cta_sync();
  BXG_PHASE_END(st, 12);
  integrate(ex, c);
  BXG_PHASE_END(st, 3);
  update_position_terms<X, Cfg, INV>(ex, c, st, true);
}

// pipeline.init (pipeline.py:51-61); q, qd already in the slab
template <class X, class Cfg>
BXG_HD void init_env(X& ex, const Ctx& c, Stats* st) {
  BXG_GET_DIMS(c); real* s = c.s;
  ex.lanes([&](int lane) {
    for (int i = lane; i < D.nv; i += X::G) { s[D.s_qfs + i] = R(0.); s[D.s_qfc + i] = R(0.); s[D.s_qdd + i] = R(0.); }
    for (int i = lane; i < (D.nc > 0 ? D.nc : 1) * D.jld; i += X::G) s[D.s_J + i] = R(0.);
  });
  update_position_terms<X, Cfg, 1>(ex, c, st, false);
}

// Zeroes every region whose padding the register-row kernels rely on (rows and
// columns past nv / nc are read by compile-time-width loops and must contribute 0).
template <class X>
BXG_HD void prepare_env(X& ex, const Ctx& c) {
  BXG_GET_DIMS(c); real* s = c.s;
  const int ncz = D.nc > 0 ? D.nc : 1;
  ex.lanes([&](int lane) {
    const int G = X::G;
    for (int i = lane; i < D.nvw * D.nvp; i += G) { s[D.s_M + i] = R(0.); s[D.s_Minv + i] = R(0.); }
    for (int i = lane; i < ncz * D.jld; i += G) s[D.s_J + i] = R(0.);
    for (int i = lane; i < D.nvw; i += G) s[D.s_qfs + i] = R(0.);   // read 128 bits at a time: the padding must be zero
    for (int i = lane; i < D.nv; i += G) { s[D.s_qfc + i] = R(0.); s[D.s_qdd + i] = R(0.); }
    for (int i = lane; i < D.ncw; i += G) { s[D.s_b + i] = R(0.); s[D.s_px + i] = R(0.); s[D.s_py + i] = R(0.); s[D.s_pg + i] = R(0.); s[D.s_pres + i] = R(0.); s[D.s_pxn + i] = R(0.); }
  });
}

// ============================================================ env epilogue
// What the reference envs compute around pipeline_step, evaluated while the
// post-step state is still in the slab.  `before` holds what must be captured
// from the pre-step state: position of link 0 (Ant) or the centre of mass
// (Humanoid), stored in s_red[0..2] by env_prologue.
// a point fixed in the frame of a link: x.take(link).do(Transform.create(pos=p)).pos
BXG_HD V3 env_tip(const Ctx& c, const BxgEnvSpec& sp) {
  BXG_GET_DIMS(c); const real* s = c.s;
  return ld3(s + D.s_x_pos + 3 * sp.tip_link) + rotate(V3{sp.tip_pos[0], sp.tip_pos[1], sp.tip_pos[2]}, ld4(s + D.s_x_rot + 4 * sp.tip_link));
}
// x.take(link).do(Transform.create(pos=inertia.transform.pos[link])).pos: the link's centre of mass
BXG_HD V3 env_link_com(const Ctx& c, int l) {
  BXG_GET_DIMS(c); const real* s = c.s;
  return ld3(s + D.s_x_pos + 3 * l) + rotate(ld3(c.mf + D.m_in_pos + 3 * l), ld4(s + D.s_x_rot + 4 * l));
}
// math.safe_norm (brax/math.py:308-328)
BXG_HD real env_safe_norm(V3 v) {
  bool zero = r_abs(v.x) <= R(1e-8) && r_abs(v.y) <= R(1e-8) && r_abs(v.z) <= R(1e-8);
  return zero ? R(0.) : r_sqrt(v.x * v.x + v.y * v.y + v.z * v.z);
}

// whole-model centre of mass from link poses in s_x_pos / s_x_rot
// (envs/humanoid.py:339-354 `_com`); every lane computes it redundantly
BXG_HD V3 env_com(const Ctx& c) {
  BXG_GET_DIMS(c); const real* mf = c.mf; const real* s = c.s;
  V3 msum{0, 0, 0}; real mtot = R(0.);
  for (int l = 0; l < D.L; ++l) {
    real m = mf[D.m_in_mass + l];
    V3 xi = ld3(s + D.s_x_pos + 3 * l) + rotate(ld3(mf + D.m_in_pos + 3 * l), ld4(s + D.s_x_rot + 4 * l));
    msum = fmav(xi, m, msum); mtot += m;
  }
  return V3{msum.x / mtot, msum.y / mtot, msum.z / mtot};
}

// captures the pre-step reference point and rescales the action (Humanoid)
template <class X, class ST>
BXG_HD void env_prologue(X& ex, const Ctx& c, const BxgEnvSpec& sp, const ST& g, int64_t e) {
  BXG_GET_DIMS(c); const real* mf = c.mf; real* s = c.s;
  const int L = D.L;
  ex.lanes([&](int lane) {
    for (int i = lane; i < L * 3; i += X::G) s[D.s_x_pos + i] = g.x_pos[e * L * 3 + i];
    for (int i = lane; i < L * 4; i += X::G) s[D.s_x_rot + i] = g.x_rot[e * L * 4 + i];
    if (sp.kind == BXG_ENV_COM_VELOCITY || sp.kind == BXG_ENV_CARTPOLE || sp.kind == BXG_ENV_STANDUP || sp.kind == BXG_ENV_PUSHER) {
      // action = (a + 1) * (hi - lo) * 0.5 + lo   (envs/humanoid.py:260-262, inverted_pendulum.py:134-137)
      for (int a = lane; a < D.nu; a += X::G) {
        real lo = mf[D.m_act_clo + a], hi = mf[D.m_act_chi + a];
        s[D.s_act + a] = (s[D.s_act + a] + R(1.)) * (hi - lo) * R(0.5) + lo;
      }
    }
  });
  V3 ref = sp.kind == BXG_ENV_COM_VELOCITY ? env_com(c) : ld3(s + D.s_x_pos);
  if (sp.kind == BXG_ENV_SWIMMER) ref = V3{s[D.s_q], s[D.s_q + 1], R(0.)};   // pipeline_state0.q[:2] (envs/swimmer.py:165-167)
  if (sp.kind == BXG_ENV_PUSHER) {
    // reward_near / reward_dist come from the PRE-step state (envs/pusher.py:204-211: state.pipeline_state.x)
    V3 obj = env_link_com(c, sp.object_link);
    ref = V3{-env_safe_norm(obj - env_link_com(c, sp.tip_link)), -env_safe_norm(obj - env_link_com(c, sp.target_link)), R(0.)};
  }
  ex.lanes([&](int lane) { if (lane == 0) st3(s + D.s_red, ref); });
}

// _get_obs (envs/ant.py:271-279, envs/humanoid.py:301-337) from the slab; for the
// COM kind s_tau must already hold actuator.to_tau at the current q, qd
template <class X>
BXG_HD void env_write_obs(X& ex, const Ctx& c, const BxgEnvSpec& sp, real* o) {
  BXG_GET_DIMS(c); const real* mf = c.mf; real* s = c.s;
  const int L = D.L, nv = D.nv, nq = D.nq;
  ex.lanes([&](int lane) {
    const int G = X::G;
    const int np = nq - sp.obs_skip;
    if (sp.kind == BXG_ENV_DOUBLE_CARTPOLE) {
      // [q[:1], sin(q[1:]), cos(q[1:]), clip(qd, -10, 10)]  (envs/inverted_double_pendulum.py:187-195)
      const int na = nq - 1;
      for (int i = lane; i < nq; i += G) {
        if (i == 0) { o[0] = s[D.s_q]; continue; }
        real sn, cs; r_sincos(s[D.s_q + i], &sn, &cs);
        o[i] = sn; o[na + i] = cs;
      }
      for (int i = lane; i < nv; i += G) o[1 + 2 * na + i] = r_max(-R(10.), r_min(s[D.s_qd + i], R(10.)));
    } else if (sp.kind == BXG_ENV_PUSHER) {
      // [q[:7], qd[:7], x_i.pos[tips_arm], x_i.pos[object], x_i.pos[goal]]  (envs/pusher.py:224-237)
      const int na = D.nu;
      for (int i = lane; i < na; i += G) { o[i] = s[D.s_q + i]; o[na + i] = s[D.s_qd + i]; }
      if (lane < 3) st3(o + 2 * na + 3 * lane, env_link_com(c, lane == 0 ? sp.tip_link : (lane == 1 ? sp.object_link : sp.target_link)));
    } else if (sp.kind == BXG_ENV_REACHER) {
      // [cos(theta), sin(theta), q[2:], tip_vel[:2], tip_pos - target_pos]  (envs/reacher.py:215-239)
      if (lane == 0) {
        for (int i = 0; i < 2; ++i) { real sn, cs; r_sincos(s[D.s_q + i], &sn, &cs); o[i] = cs; o[2 + i] = sn; }
        for (int i = 2; i < nq; ++i) o[2 + i] = s[D.s_q + i];
        // Transform.create(pos=tip).do(xd.take(tip_link)).vel = vel - tip x ang  (identity rotation, base.py:565-570)
        V3 tp{sp.tip_pos[0], sp.tip_pos[1], sp.tip_pos[2]};
        V3 tv = ld3(s + D.s_xd_vel + 3 * sp.tip_link) - cross(tp, ld3(s + D.s_xd_ang + 3 * sp.tip_link));
        V3 tt = env_tip(c, sp) - ld3(s + D.s_x_pos + 3 * sp.target_link);
        real* t = o + 2 + nq;
        t[0] = tv.x; t[1] = tv.y; t[2] = tt.x; t[3] = tt.y; t[4] = tt.z;
      }
    } else if (sp.kind == BXG_ENV_PLANAR) {
      // position = q.at[1].set(x.pos[0, 2]); velocity = clip(qd, -10, 10)  (envs/hopper.py:266-276, walker2d.py:263-273)
      for (int i = lane; i < np; i += G) o[i] = sp.obs_skip + i == 1 ? s[D.s_x_pos + 2] : s[D.s_q + sp.obs_skip + i];
      for (int i = lane; i < nv; i += G) o[np + i] = r_max(-R(10.), r_min(s[D.s_qd + i], R(10.)));
    } else {
      for (int i = lane; i < np; i += G) o[i] = s[D.s_q + sp.obs_skip + i];
      for (int i = lane; i < nv; i += G) o[np + i] = s[D.s_qd + i];
    }
    if (sp.kind == BXG_ENV_COM_VELOCITY || sp.kind == BXG_ENV_STANDUP) {
      real mass_sum = R(0.);
      for (int l = 0; l < L; ++l) mass_sum += mf[D.m_in_mass + l];
      real* oi = o + np + nv; real* ov = oi + 10 * L; real* of = ov + 6 * L;
      // com_inertia = [cinr.i (9), mass]: x_i - com is what transform_com used (one tree)
      for (int i = lane; i < L * 10; i += G) { int l = i / 10, k = i - 10 * l; oi[i] = k < 9 ? s[D.s_cinr_i + 9 * l + k] : mf[D.m_in_mass + l]; }
      // com_velocity = [m * (xd.vel - (x_i - x) x xd.ang) / mass_sum, xd.ang]  (humanoid.py:316-323)
      for (int l = lane; l < L; l += G) {
        V3 xp = ld3(s + D.s_x_pos + 3 * l);
        V3 off = (xp + rotate(ld3(mf + D.m_in_pos + 3 * l), ld4(s + D.s_x_rot + 4 * l))) - xp;
        V3 ang = ld3(s + D.s_xd_ang + 3 * l);
        V3 v = ld3(s + D.s_xd_vel + 3 * l) - cross(off, ang);
        real m = mf[D.m_in_mass + l];
        ov[6 * l + 0] = m * v.x / mass_sum; ov[6 * l + 1] = m * v.y / mass_sum; ov[6 * l + 2] = m * v.z / mass_sum;
        ov[6 * l + 3] = ang.x; ov[6 * l + 4] = ang.y; ov[6 * l + 5] = ang.z;
      }
      for (int i = lane; i < nv; i += G) of[i] = s[D.s_tau + i];
    }
  });
}

// observation of a freshly initialised state: action = zeros (humanoid.py:241)
template <class X>
BXG_HD void env_reset_obs(X& ex, const Ctx& c, const BxgEnvSpec& sp, real* o) {
  BXG_GET_DIMS(c); real* s = c.s;
  ex.lanes([&](int lane) { for (int a = lane; a < D.nu; a += X::G) s[D.s_act + a] = R(0.); });
  if (sp.kind == BXG_ENV_COM_VELOCITY || sp.kind == BXG_ENV_STANDUP) actuator_tau(ex, c);
  env_write_obs(ex, c, sp, o);
}

template <class X, class IO>
BXG_HD void env_epilogue(X& ex, const Ctx& c, const BxgEnvSpec& sp, const IO& io, int64_t e, bool valid, bool* done_out) {
  BXG_GET_DIMS(c); const real* mf = c.mf; real* s = c.s;
  const int L = D.L, nv = D.nv, nq = D.nq;
  const bool com_kind = sp.kind == BXG_ENV_COM_VELOCITY;
  if (com_kind || sp.kind == BXG_ENV_STANDUP) actuator_tau(ex, c);   // qfrc_actuator at the post-step q, qd (humanoid.py:325-327)
  // ---- reward / done / metrics: every lane redundantly (uniform scalars) ----
  V3 before = ld3(s + D.s_red);
  V3 after = com_kind ? env_com(c) : ld3(s + D.s_x_pos);
  if (sp.kind == BXG_ENV_SWIMMER) after = V3{s[D.s_q], s[D.s_q + 1], R(0.)};   // xy_position = q[:2] (envs/swimmer.py:164)
  V3 vel{(after.x - before.x) / sp.env_dt, (after.y - before.y) / sp.env_dt, (after.z - before.z) / sp.env_dt};
  real forward_reward = sp.forward_reward_weight * vel.x;   // Ant has no weight (envs/ant.py:240): its spec carries 1.0, exact
  real z = s[D.s_x_pos + 2];
  real is_healthy = z < sp.healthy_z_min ? R(0.) : R(1.);
  if (z > sp.healthy_z_max) is_healthy = R(0.);
  if (sp.kind == BXG_ENV_PLANAR) {
    // strict ranges on z, the root angle q[2] and every entry of [q[2:], qd]
    // (envs/hopper.py:233-244; walker2d.py:215-219 is the same test without the state range)
    const real angle = s[D.s_q + 2];
    bool ok = sp.healthy_z_min < z && z < sp.healthy_z_max && sp.healthy_angle_min < angle && angle < sp.healthy_angle_max;
    for (int i = 2; i < nq; ++i) ok = ok && sp.healthy_state_min < s[D.s_q + i] && s[D.s_q + i] < sp.healthy_state_max;
    for (int i = 0; i < nv; ++i) ok = ok && sp.healthy_state_min < s[D.s_qd + i] && s[D.s_qd + i] < sp.healthy_state_max;
    is_healthy = ok ? R(1.) : R(0.);
  }
  real healthy_reward = sp.terminate_when_unhealthy ? sp.healthy_reward : sp.healthy_reward * is_healthy;
  real sq = R(0.);
  for (int a = 0; a < D.nu; ++a) sq = r_fma(s[D.s_act + a], s[D.s_act + a], sq);
  real ctrl_cost = sp.ctrl_cost_weight * sq;
  real reward = com_kind ? (forward_reward + healthy_reward) - ctrl_cost : ((forward_reward + healthy_reward) - ctrl_cost) - R(0.);
  real done = sp.terminate_when_unhealthy ? R(1.) - is_healthy : R(0.);
  // the small classic-control envs: their own reward / done; metrics (if any) in ms[]
  const bool simple_kind = sp.kind >= BXG_ENV_CARTPOLE;
  real ms[BXG_ENV_NUM_METRICS] = {R(0.), R(0.), R(0.), R(0.), R(0.), R(0.), R(0.), R(0.), R(0.), R(0.)};
  if (sp.kind == BXG_ENV_CARTPOLE) {
    // reward = 1.0; done = |obs[1]| > 0.2  (envs/inverted_pendulum.py:141-142)
    reward = R(1.0); done = r_abs(s[D.s_q + 1]) > sp.healthy_angle_max ? R(1.) : R(0.);
  } else if (sp.kind == BXG_ENV_DOUBLE_CARTPOLE) {
    // envs/inverted_double_pendulum.py:164-177
    V3 tip = env_tip(c, sp);
    const real x = tip.x, y = tip.z, v1 = s[D.s_qd + 1], v2 = s[D.s_qd + 2];
    real dist_penalty = R(0.01) * (x * x) + (y - R(2.)) * (y - R(2.));
    real vel_penalty = R(1e-3) * (v1 * v1) + R(5e-3) * (v2 * v2);
    done = y <= sp.healthy_z_min ? R(1.) : R(0.);
    reward = ((R(1.) - done) * sp.healthy_reward - dist_penalty) - vel_penalty;
  } else if (sp.kind == BXG_ENV_REACHER) {
    // reward_dist = -safe_norm(obs[-3:]); reward_ctrl = -sum(action^2)  (envs/reacher.py:203-207)
    V3 tt = env_tip(c, sp) - ld3(s + D.s_x_pos + 3 * sp.target_link);
    ms[0] = -env_safe_norm(tt); ms[1] = -sq;
    reward = ms[0] + ms[1]; done = R(0.);
  } else if (sp.kind == BXG_ENV_PUSHER) {
    // reward = reward_dist + 0.1 * reward_ctrl + 0.5 * reward_near  (envs/pusher.py:212-215); before = (near, dist) of the old state
    ms[0] = before.y; ms[1] = -sq; ms[2] = before.x;
    reward = (ms[0] + R(0.1) * ms[1]) + R(0.5) * ms[2]; done = R(0.);
  } else if (sp.kind == BXG_ENV_STANDUP) {
    // uph_cost = (z - 0) / dt; reward = uph_cost + 1 - 0.01 * sum(action^2)  (envs/humanoidstandup.py:227-236)
    real uph_cost = (z - R(0.)) / sp.env_dt;
    reward = (uph_cost + sp.healthy_reward) - ctrl_cost; done = R(0.);
    ms[0] = uph_cost; ms[1] = -ctrl_cost;
  } else if (sp.kind == BXG_ENV_SWIMMER) {
    // envs/swimmer.py:168-183 (jp.linalg.norm, not safe_norm; 'forward_reward' is never updated)
    reward = forward_reward - ctrl_cost; done = R(0.);
    ms[0] = forward_reward; ms[2] = -ctrl_cost; ms[4] = after.x; ms[5] = after.y;
    ms[6] = r_sqrt(after.x * after.x + after.y * after.y); ms[7] = vel.x; ms[8] = vel.y;
  }
  // EpisodeWrapper (wrappers/training.py:98-135) after AutoResetWrapper's step reset (:141-146)
  real trunc = R(0.);
  if (io.steps && valid) {
    real steps = io.done[e] != R(0.) ? R(0.) : io.steps[e];
    steps += R(1.);
    if (sp.episode_length > 0 && steps >= (real)sp.episode_length) { trunc = R(1.) - done; done = R(1.); }
    ex.lanes([&](int lane) { if (lane == 0) { io.steps[e] = steps; if (io.truncation) io.truncation[e] = trunc; } });
  }
  *done_out = done != R(0.);
  if (!valid) return;
  const int osz = env_obs_size(D, sp);
  const bool use_first = *done_out && io.first_state != nullptr;
  ex.lanes([&](int lane) {
    const int G = X::G;
    if (lane == 0) {
      io.reward[e] = reward; io.done[e] = done;
      real* m = io.metrics + e * BXG_ENV_NUM_METRICS;
      real dist;
      if (simple_kind) {
        for (int i = 0; i < BXG_ENV_NUM_METRICS; ++i) m[i] = ms[i];
      } else if (com_kind) {
        dist = r_sqrt(after.x * after.x + after.y * after.y + after.z * after.z);
        m[0] = forward_reward; m[1] = forward_reward; m[2] = -ctrl_cost; m[3] = healthy_reward;
        m[4] = after.x; m[5] = after.y; m[6] = dist; m[7] = vel.x; m[8] = vel.y; m[9] = R(0.);
      } else {
        bool zero = r_abs(after.x) <= R(1e-8) && r_abs(after.y) <= R(1e-8) && r_abs(after.z) <= R(1e-8);  // math.safe_norm
        dist = zero ? R(0.) : r_sqrt(after.x * after.x + after.y * after.y + after.z * after.z);
        m[0] = forward_reward; m[1] = healthy_reward; m[2] = -ctrl_cost; m[3] = -R(0.);
        m[4] = after.x; m[5] = after.y; m[6] = dist; m[7] = vel.x; m[8] = vel.y; m[9] = forward_reward;
      }
    }
    if (use_first) {   // AutoResetWrapper: obs = where(done, first_obs, obs)
      real* o = io.obs + e * osz;
      for (int i = lane; i < osz; i += G) o[i] = io.first_obs[e * osz + i];
    }
  });
  if (!use_first) env_write_obs(ex, c, sp, io.obs + e * osz);
}

// AutoResetWrapper for the pipeline state: copy first_state's leaves for a done env
template <class X, class ST>
BXG_HD void store_first_state(X& ex, const Ctx& c, const ST& g, const ST& f, int64_t e) {
  BXG_GET_DIMS(c);
  const int L = D.L, nv = D.nv, nq = D.nq, nc = D.nc;
  ex.lanes([&](int lane) {
    const int G = X::G;
    auto cp = [&](real* dst, const real* src, int n) { for (int i = lane; i < n; i += G) dst[e * n + i] = src[e * n + i]; };
    cp(g.q, f.q, nq); cp(g.qd, f.qd, nv); cp(g.x_pos, f.x_pos, L * 3); cp(g.x_rot, f.x_rot, L * 4);
    cp(g.xd_ang, f.xd_ang, L * 3); cp(g.xd_vel, f.xd_vel, L * 3); cp(g.root_com, f.root_com, L * 3);
    cp(g.cinr_pos, f.cinr_pos, L * 3); cp(g.cinr_rot, f.cinr_rot, L * 4); cp(g.cinr_i, f.cinr_i, L * 9); cp(g.cinr_mass, f.cinr_mass, L);
    cp(g.cd_ang, f.cd_ang, L * 3); cp(g.cd_vel, f.cd_vel, L * 3); cp(g.cdof_ang, f.cdof_ang, nv * 3); cp(g.cdof_vel, f.cdof_vel, nv * 3);
    cp(g.cdofd_ang, f.cdofd_ang, nv * 3); cp(g.cdofd_vel, f.cdofd_vel, nv * 3);
    cp(g.mass_mx, f.mass_mx, nv * nv); cp(g.mass_mx_inv, f.mass_mx_inv, nv * nv);
    cp(g.con_jac, f.con_jac, nc * nv); cp(g.con_diag, f.con_diag, nc); cp(g.con_aref, f.con_aref, nc);
    cp(g.qf_smooth, f.qf_smooth, nv); cp(g.qf_constraint, f.qf_constraint, nv); cp(g.qdd, f.qdd, nv);
  });
}

// ------------------------------------------------------- global <-> slab I/O
template <class X, class F>
BXG_HD void io_vec(X& ex, int n, F f) {
  ex.lanes([&](int lane) { for (int i = lane; i < n; i += X::G) f(i); });
}

// ---- global -> shared copies with the loads in flight together ---------------------------------------------
// A plain `for (i = lane; i < n; i += G) dst[i] = src[i]` compiles to load, wait, store, next: one DRAM / L2 latency
// per element and lane (the load phase of a full State was ~48 such round trips: 26 k cycles per env pass, SM idle).
// These helpers issue the loads of several leaves and several elements per lane first and store afterwards.
#ifndef BXG_IO_BATCH
#define BXG_IO_BATCH 4
#endif
template <class T>
struct IoLeaf { real* dst; const T* src; int n; };
// K leaves, each n <= any size: per round every lane fetches up to U elements of every leaf, then stores them
template <int G, int U, int K, class T>
BXG_HD void copy_in_many(const IoLeaf<T> (&lv)[K], int lane) {
  int nmax = 0;
#pragma unroll
  for (int k = 0; k < K; ++k) nmax = lv[k].n > nmax ? lv[k].n : nmax;
  for (int i0 = lane; i0 < nmax; i0 += G * U) {
    real v[K][U];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int u = 0; u < U; ++u) { const int i = i0 + u * G; v[k][u] = i < lv[k].n ? (real)lv[k].src[i] : R(0.); }
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int u = 0; u < U; ++u) { const int i = i0 + u * G; if (i < lv[k].n) lv[k].dst[i] = v[k][u]; }
  }
}
// K row-major [rows][nv] matrices of one shape into shared-memory copies of row stride ld; the (row, column) of
// element i advances incrementally (no integer division per element)
template <int G, int U, int K, class T>
BXG_HD void copy_in_matrices(const IoLeaf<T> (&lv)[K], int nv, int ld, int lane) {
  const int n = lv[0].n;
  const int dq = G / nv, dr = G - dq * nv;    // i += G  <=>  (r, cc) += (dq, dr) with one carry
  int r = lane / nv, cc = lane - r * nv;
  for (int i0 = lane; i0 < n; i0 += G * U) {
    real v[K][U];
#pragma unroll
    for (int k = 0; k < K; ++k)
#pragma unroll
      for (int u = 0; u < U; ++u) { const int i = i0 + u * G; v[k][u] = i < n ? (real)lv[k].src[i] : R(0.); }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = i0 + u * G;
      if (i < n) {
#pragma unroll
        for (int k = 0; k < K; ++k) lv[k].dst[r * ld + cc] = v[k][u];
      }
      r += dq; cc += dr; if (cc >= nv) { cc -= nv; ++r; }
    }
  }
}

// Loads the State leaves pipeline.step reads.
// ST: BxgState (include/bxg.h), or the emulator's double-precision twin with the same member names
template <class X, class ST>
BXG_HD void load_env(X& ex, const Ctx& c, const ST& g, const real* act, int64_t e) {
  BXG_GET_DIMS(c); real* s = c.s;
  const int L = D.L, nv = D.nv, nq = D.nq, nc = D.nc, nvp = D.nvp;
  ex.lanes([&](int lane) {
    constexpr int G = X::G, U = BXG_IO_BATCH;
    using T = typename std::remove_cv<typename std::remove_pointer<decltype(g.q)>::type>::type;
    using Lf = IoLeaf<T>;
    {
      const Lf lv[8] = {{s + D.s_q, g.q + e * nq, nq}, {s + D.s_qd, g.qd + e * nv, nv}, {s + D.s_act, (const T*)act + e * D.nu, D.nu},
                        {s + D.s_cinr_pos, g.cinr_pos + e * L * 3, L * 3}, {s + D.s_cd_ang, g.cd_ang + e * L * 3, L * 3},
                        {s + D.s_cd_vel, g.cd_vel + e * L * 3, L * 3}, {s + D.s_diag, g.con_diag + e * nc, nc}, {s + D.s_aref, g.con_aref + e * nc, nc}};
      copy_in_many<G, U, 8, T>(lv, lane);
    }
    {
      const Lf lv[5] = {{s + D.s_cinr_i, g.cinr_i + e * L * 9, L * 9}, {s + D.s_cdof_ang, g.cdof_ang + e * nv * 3, nv * 3}, {s + D.s_cdof_vel, g.cdof_vel + e * nv * 3, nv * 3},
                        {s + D.s_cdofd_ang, g.cdofd_ang + e * nv * 3, nv * 3}, {s + D.s_cdofd_vel, g.cdofd_vel + e * nv * 3, nv * 3}};
      copy_in_many<G, U, 5, T>(lv, lane);
    }
    if (D.fluid) {   // fluid.force reads x and root_com of the incoming state (dynamics.py:199-206)
      const Lf lv[3] = {{s + D.s_x_pos, g.x_pos + e * L * 3, L * 3}, {s + D.s_root_com, g.root_com + e * L * 3, L * 3}, {s + D.s_x_rot, g.x_rot + e * L * 4, L * 4}};
      copy_in_many<G, U, 3, T>(lv, lane);
    }
    // matrices: flat, fully coalesced reads into the re-strided shared-memory copies
    {
      const Lf lv[2] = {{s + D.s_Minv, g.mass_mx_inv + e * nv * nv, nv * nv}, {s + D.s_M, g.mass_mx + e * nv * nv, nv * nv}};
      copy_in_matrices<G, U, 2, T>(lv, nv, nvp, lane);
    }
    {
      const Lf lv[1] = {{s + D.s_J, g.con_jac + e * nc * nv, nc * nv}};
      copy_in_matrices<G, 2 * U, 1, T>(lv, nv, D.jld, lane);
    }
  });
  // active rows of the incoming jacobian: anything non-zero in the row (lane i starts its scan
  // at column i so that the lanes of a group read different banks)
  ex.lanes([&](int lane) {
    const int G = X::G, jld = D.jld;
    for (int i = lane; i < nc; i += G) {
      bool nz = s[D.s_diag + i] != R(0.) || s[D.s_aref + i] != R(0.);
      const int i0 = i % jld;   // (models with more than 2 * jld rows: the start column must wrap as often as needed)
      for (int k = 0; k < jld && !nz; ++k) { int kk = k + i0; kk = kk >= jld ? kk - jld : kk; nz = s[D.s_J + i * jld + kk] != R(0.); }
      row_active(s, D)[i] = nz ? 1 : 0;
    }
  });
}

// ---- lean state I/O (BXG_STEP_LEAN) ------------------------------------------------
// q, qd, act and mass_mx_inv are the step's only independent inputs: every other leaf it reads is a function of
// (q, qd) that the previous step (or init) computed with the code below, so recomputing it at entry gives the
// same bits.  x of the incoming state is read by the env prologue (pre-step reference point).
template <class X, class ST>
BXG_HD void load_env_lean(X& ex, const Ctx& c, const ST& g, const real* act, int64_t e) {
  BXG_GET_DIMS(c); real* s = c.s;
  const int nv = D.nv, nq = D.nq, nvp = D.nvp;
  ex.lanes([&](int lane) {
    constexpr int G = X::G, U = BXG_IO_BATCH;
    using T = typename std::remove_cv<typename std::remove_pointer<decltype(g.q)>::type>::type;
    using Lf = IoLeaf<T>;
    const Lf lv[3] = {{s + D.s_q, g.q + e * nq, nq}, {s + D.s_qd, g.qd + e * nv, nv}, {s + D.s_act, (const T*)act + e * D.nu, D.nu}};
    copy_in_many<G, U, 3, T>(lv, lane);
    const Lf mv[1] = {{s + D.s_Minv, g.mass_mx_inv + e * nv * nv, nv * nv}};
    copy_in_matrices<G, 2 * U, 1, T>(mv, nv, nvp, lane);
  });
}
// what update_position_terms leaves behind, minus the Newton-Schulz run that produced the incoming mass_mx_inv
template <class X, class Cfg>
BXG_HD void lean_entry(X& ex, const Ctx& c) {
  kinematics<X, !BXG_XD_TAIL>(ex, c);
  transform_com(ex, c);
  if (BXG_DIMS_OF(c).ns_iters == 0) mass_matrix(ex, c);   // integrate's implicit-damping solve reads M (integrator.py:58-60)
  con_jacobian<X, Cfg>(ex, c);
}
template <class X, class ST>
BXG_HD void store_env_lean(X& ex, const Ctx& c, const ST& g, int64_t e, const BxgDiag* dg, const Stats& st) {
  BXG_GET_DIMS(c); const real* s = c.s;
  const int L = D.L, nv = D.nv, nq = D.nq, nvp = D.nvp;
  ex.lanes([&](int lane) {
    const int G = X::G;
    for (int i = lane; i < nq; i += G) g.q[e * nq + i] = s[D.s_q + i];
    for (int i = lane; i < nv; i += G) g.qd[e * nv + i] = s[D.s_qd + i];
    for (int i = lane; i < L * 3; i += G) {
      g.x_pos[e * L * 3 + i] = s[D.s_x_pos + i]; g.xd_ang[e * L * 3 + i] = s[D.s_xd_ang + i]; g.xd_vel[e * L * 3 + i] = s[D.s_xd_vel + i];
    }
    for (int i = lane; i < L * 4; i += G) g.x_rot[e * L * 4 + i] = s[D.s_x_rot + i];
    const int dq = G / nv, dr = G - dq * nv;
    int r = lane / nv, cc = lane - r * nv;
    for (int i = lane; i < nv * nv; i += G) {
      g.mass_mx_inv[e * nv * nv + i] = s[D.s_Minv + r * nvp + cc];
      r += dq; cc += dr; if (cc >= nv) { cc -= nv; ++r; }
    }
    if (dg) {
      if (dg->con_dist) for (int i = lane; i < D.ncon; i += G) dg->con_dist[e * D.ncon + i] = s[D.s_dist + i];
      if (dg->stats && lane == 0) {
        int32_t* o = dg->stats + e * 4;
        o[0] += st.pg_iters; o[1] += st.pg_trials; o[2] += st.ns_accepts; o[3] += st.ns_cold;
      }
    }
  });
}
template <class X, class ST>
BXG_HD void store_first_state_lean(X& ex, const Ctx& c, const ST& g, const ST& f, int64_t e) {
  BXG_GET_DIMS(c);
  const int L = D.L, nv = D.nv, nq = D.nq;
  ex.lanes([&](int lane) {
    const int G = X::G;
    auto cp = [&](real* dst, const real* src, int n) { for (int i = lane; i < n; i += G) dst[e * n + i] = src[e * n + i]; };
    cp(g.q, f.q, nq); cp(g.qd, f.qd, nv); cp(g.x_pos, f.x_pos, L * 3); cp(g.x_rot, f.x_rot, L * 4);
    cp(g.xd_ang, f.xd_ang, L * 3); cp(g.xd_vel, f.xd_vel, L * 3); cp(g.mass_mx_inv, f.mass_mx_inv, nv * nv);
  });
}

template <class X>
BXG_HD void load_env_qqd(X& ex, const Ctx& c, const real* q, const real* qd, int64_t e) {
  BXG_GET_DIMS(c); real* s = c.s;
  ex.lanes([&](int lane) {
    for (int i = lane; i < D.nq; i += X::G) s[D.s_q + i] = q[e * D.nq + i];
    for (int i = lane; i < D.nv; i += X::G) s[D.s_qd + i] = qd[e * D.nv + i];
  });
}

template <class X, class ST>
BXG_HD void store_env(X& ex, const Ctx& c, const ST& g, int64_t e, const BxgDiag* dg, const Stats& st) {
  BXG_GET_DIMS(c); const real* s = c.s;
  const int L = D.L, nv = D.nv, nq = D.nq, nc = D.nc, nvp = D.nvp;
  ex.lanes([&](int lane) {
    const int G = X::G;
    for (int i = lane; i < nq; i += G) g.q[e * nq + i] = s[D.s_q + i];
    for (int i = lane; i < nv; i += G) {
      g.qd[e * nv + i] = s[D.s_qd + i];
      g.qf_smooth[e * nv + i] = s[D.s_qfs + i]; g.qf_constraint[e * nv + i] = s[D.s_qfc + i]; g.qdd[e * nv + i] = s[D.s_qdd + i];
    }
    for (int i = lane; i < L * 3; i += G) {
      g.x_pos[e * L * 3 + i] = s[D.s_x_pos + i]; g.xd_ang[e * L * 3 + i] = s[D.s_xd_ang + i]; g.xd_vel[e * L * 3 + i] = s[D.s_xd_vel + i];
      g.root_com[e * L * 3 + i] = s[D.s_root_com + i]; g.cinr_pos[e * L * 3 + i] = s[D.s_cinr_pos + i];
      g.cd_ang[e * L * 3 + i] = s[D.s_cd_ang + i]; g.cd_vel[e * L * 3 + i] = s[D.s_cd_vel + i];
    }
    for (int i = lane; i < L * 4; i += G) { g.x_rot[e * L * 4 + i] = s[D.s_x_rot + i]; g.cinr_rot[e * L * 4 + i] = s[D.s_cinr_rot + i]; }
    for (int i = lane; i < L * 9; i += G) g.cinr_i[e * L * 9 + i] = s[D.s_cinr_i + i];
    for (int i = lane; i < L; i += G) g.cinr_mass[e * L + i] = c.mf[D.m_in_mass + i];   // cinr.mass is the link's mass (base.py:588-594)
    for (int i = lane; i < nv * 3; i += G) {
      g.cdof_ang[e * nv * 3 + i] = s[D.s_cdof_ang + i]; g.cdof_vel[e * nv * 3 + i] = s[D.s_cdof_vel + i];
      g.cdofd_ang[e * nv * 3 + i] = s[D.s_cdofd_ang + i]; g.cdofd_vel[e * nv * 3 + i] = s[D.s_cdofd_vel + i];
    }
    {
      const int dq = G / nv, dr = G - dq * nv;
      int r = lane / nv, cc = lane - r * nv;
      for (int i = lane; i < nv * nv; i += G) {
        g.mass_mx[e * nv * nv + i] = s[D.s_M + r * nvp + cc];
        g.mass_mx_inv[e * nv * nv + i] = s[D.s_Minv + r * nvp + cc];
        r += dq; cc += dr; if (cc >= nv) { cc -= nv; ++r; }
      }
      r = lane / nv; cc = lane - r * nv;
      for (int i = lane; i < nc * nv; i += G) {
        g.con_jac[e * nc * nv + i] = s[D.s_J + r * D.jld + cc];
        r += dq; cc += dr; if (cc >= nv) { cc -= nv; ++r; }
      }
    }
    for (int i = lane; i < nc; i += G) { g.con_diag[e * nc + i] = s[D.s_diag + i]; g.con_aref[e * nc + i] = s[D.s_aref + i]; }
    if (dg) {
      if (dg->con_dist) for (int i = lane; i < D.ncon; i += G) dg->con_dist[e * D.ncon + i] = s[D.s_dist + i];
      if (dg->stats && lane == 0) {
        int32_t* o = dg->stats + e * 4;
        o[0] += st.pg_iters; o[1] += st.pg_trials; o[2] += st.ns_accepts; o[3] += st.ns_cold;
      }
    }
  });
}

}  // namespace bxg
