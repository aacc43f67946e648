/*
 * ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product.
 *
 * CPU restatement of the reference's `generalized` physics step
 * (google/brax 0.14.2), one environment at a time, scalar C, written to follow
 * the reference's arithmetic statement by statement.  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg
 * may build, load or call this file.  The product (brax_b200/) never does.
 *
 * PARITY STATUS: pinned against outputs of the reference's own source for
 * everything that lives in /root/reference; "parity unpinned" for the three
 * third-party pieces that do not.
 *   PINNED  tests/golden/ref_*.npz are written by tools/gen_reference_golden.py,
 *           which imports brax.generalized.pipeline (and brax.envs + the training
 *           wrappers) UNMODIFIED from /root/reference and runs them on NumPy
 *           float64 through small stand-ins for jax / flax (tools/refshim/).
 *           tests/test_reference_golden.py: this oracle reproduces every State
 *           leaf after init and after every step to 1e-9 (Ant, Humanoid,
 *           HalfCheetah, Hopper, a motorised pendulum; contacts and joint limits
 *           active).  Also pinned: the known answers embedded in the reference's
 *           tests (tests/test_oracle_known_answers.py) and physics invariants
 *           (tests/test_oracle_invariants.py).
 *   UNPINNED (restated from the published algorithms of packages that are not
 *           installable here, identically in the stand-ins and in this file, so
 *           the golden files cannot tell a shared misreading): jaxopt.
 *           ProjectedGradient (unpinned version; FISTA + backtracking line
 *           search), the mujoco.mjx plane-sphere / plane-capsule colliders, and
 *           MuJoCo's model compiler (brax_b200/io/mjcf.py produces the System
 *           constants both sides consume).  The real JAX float32 execution
 *           (XLA op ordering) is not reproducible here either.
 *
 * Build twice: -DORC_REAL=float (parity) and -DORC_REAL=double (sanity).
 *
 * Reference files restated here (paths relative to /root/reference):
 *   brax/generalized/pipeline.py:32-94     orc_init / orc_step
 *   brax/kinematics.py:31-108              kin_forward
 *   brax/generalized/dynamics.py:27-134    transform_com
 *   brax/generalized/dynamics.py:137-236   rne_inverse / dyn_forward
 *   brax/generalized/mass.py:27-106        mass_matrix / matrix_inv
 *   brax/math.py:278-345                   inv_approximate / safe_norm
 *   brax/generalized/constraint.py:29-240  imp_aref / jacobian / force
 *   brax/contact.py:28-67 (+ mjx)          contact_get
 *   brax/actuator.py:23-57                 to_tau
 *   brax/generalized/integrator.py:46-84   integrate
 *   brax/base.py:297-302,558-614           spatial algebra
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifndef ORC_REAL
#define ORC_REAL float
#endif
typedef ORC_REAL real;

#define MAXL 32
#define MAXQ 96
#define MAXV 64
#define MAXU 64
#define MAXCON 16
#define MAXC (4 * MAXCON + MAXV)

static inline real r_sin(real x) { return sizeof(real) == 4 ? (real)sinf((float)x) : (real)sin((double)x); }
static inline real r_cos(real x) { return sizeof(real) == 4 ? (real)cosf((float)x) : (real)cos((double)x); }
static inline real r_sqrt(real x) { return sizeof(real) == 4 ? (real)sqrtf((float)x) : (real)sqrt((double)x); }
static inline real r_pow(real x, real y) { return sizeof(real) == 4 ? (real)powf((float)x, (float)y) : (real)pow((double)x, (double)y); }
static inline real r_abs(real x) { return x < 0 ? -x : x; }
static inline real r_min(real a, real b) { return a < b ? a : b; }
static inline real r_max(real a, real b) { return a > b ? a : b; }
#define EPS_REAL (sizeof(real) == 4 ? (real)1.1920929e-07f : (real)2.220446049250313e-16)

/* ------------------------------------------------------------------ model */
typedef struct {
  int32_t num_links, nq, nv, nu, ncon, nlim;
  int32_t solver_iterations, solver_maxls, matrix_inv_iterations, pad;
  real dt, gravity[3];
  int32_t link_parent[MAXL], link_ndof[MAXL]; /* ndof==0 => free joint */
  int32_t link_q_adr[MAXL], link_qd_adr[MAXL];
  real link_tf_pos[MAXL][3], link_tf_rot[MAXL][4], link_joint_pos[MAXL][3];
  real inertia_pos[MAXL][3], inertia_rot[MAXL][4], inertia_i[MAXL][9], inertia_mass[MAXL];
  real link_invweight[MAXL];
  int32_t dof_link[MAXV];
  real dof_ang[MAXV][3], dof_vel[MAXV][3];
  real dof_armature[MAXV], dof_stiffness[MAXV], dof_damping[MAXV];
  real dof_limit_lo[MAXV], dof_limit_hi[MAXV], dof_invweight[MAXV], dof_solver_params[MAXV][7];
  int32_t act_q_id[MAXU], act_qd_id[MAXU];
  real act_gain[MAXU], act_gear[MAXU], act_ctrl_lo[MAXU], act_ctrl_hi[MAXU];
  real act_force_lo[MAXU], act_force_hi[MAXU], act_bias_q[MAXU], act_bias_qd[MAXU];
  int32_t con_link_a[MAXCON], con_link_b[MAXCON];
  real con_plane_pos[MAXCON][3], con_frame[MAXCON][9], con_sphere_pos[MAXCON][3];
  real con_radius[MAXCON], con_friction[MAXCON], con_solref[MAXCON][2], con_solimp[MAXCON][5];
  int32_t con_kind[MAXCON];          /* 0 plane-sphere, 1 one end of a plane-capsule pair */
  real con_geom_quat[MAXCON][4];     /* capsule orientation in link_b frame */
  real con_half_len[MAXCON];         /* signed: end point = centre + axis * half_len */
  int32_t enable_fluid, pad2;        /* sys.enable_fluid (io/mjcf.py:467) */
  real viscosity, density;           /* sys.viscosity, sys.density */
  /* capsule-capsule pairs (con_kind 2): shape of geom1 in the frame of link_a */
  real con_a_pos[MAXCON][3], con_a_quat[MAXCON][4], con_a_half[MAXCON], con_a_radius[MAXCON];
} OrcModel;

/* per-environment working state (reference generalized/base.py:25-92) */
typedef struct {
  real q[MAXQ], qd[MAXV];
  real x_pos[MAXL][3], x_rot[MAXL][4], xd_ang[MAXL][3], xd_vel[MAXL][3];
  real root_com[MAXL][3];
  real cinr_pos[MAXL][3], cinr_rot[MAXL][4], cinr_i[MAXL][9], cinr_mass[MAXL];
  real cd_ang[MAXL][3], cd_vel[MAXL][3];
  real cdof_ang[MAXV][3], cdof_vel[MAXV][3], cdofd_ang[MAXV][3], cdofd_vel[MAXV][3];
  real mass_mx[MAXV * MAXV], mass_mx_inv[MAXV * MAXV];
  real con_jac[MAXC * MAXV], con_diag[MAXC], con_aref[MAXC];
  real qf_smooth[MAXV], qf_constraint[MAXV], qdd[MAXV];
  /* diagnostics (not reference State fields) */
  real con_dist[MAXCON];
  int32_t pg_iters, pg_ls_trials, ns_accepts, ns_cold;
} Env;

/* batched state: env-major arrays, same field set as the product boundary */
typedef struct {
  real *q, *qd, *x_pos, *x_rot, *xd_ang, *xd_vel, *root_com;
  real *cinr_pos, *cinr_rot, *cinr_i, *cinr_mass, *cd_ang, *cd_vel;
  real *cdof_ang, *cdof_vel, *cdofd_ang, *cdofd_vel;
  real *mass_mx, *mass_mx_inv, *con_jac, *con_diag, *con_aref;
  real *qf_smooth, *qf_constraint, *qdd;
  real *con_dist;      /* [n_env, ncon]   diagnostics, may be NULL */
  int32_t *stats;      /* [n_env, 4]      diagnostics, may be NULL */
} OrcState;

/* ------------------------------------------------------------ vec / quat */
static inline real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(const real* a, const real* b, real* o) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  o[0] = x; o[1] = y; o[2] = z;
}
/* math.rotate, brax/math.py:25-41 */
static inline void rotate(const real* v, const real* q, real* o) {
  real s = q[0]; const real* u = q + 1;
  real d = dot3(u, v), uu = dot3(u, u), c[3], r[3];
  cross3(u, v, c);
  for (int i = 0; i < 3; i++) r[i] = 2 * (d * u[i]) + (s * s - uu) * v[i];
  for (int i = 0; i < 3; i++) o[i] = r[i] + 2 * s * c[i];
}
/* math.quat_mul, brax/math.py:86-101 */
static inline void quat_mul(const real* u, const real* v, real* o) {
  real w = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
  real x = u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2];
  real y = u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1];
  real z = u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0];
  o[0] = w; o[1] = x; o[2] = y; o[3] = z;
}
/* math.quat_rot_axis, brax/math.py:133-147 */
static inline void quat_rot_axis(const real* axis, real angle, real* o) {
  real s = r_sin(angle / 2), c = r_cos(angle / 2);
  o[0] = c; o[1] = axis[0] * s; o[2] = axis[1] * s; o[3] = axis[2] * s;
}
/* math.safe_norm, brax/math.py:308-328 (allclose(x,0): |x| <= 1e-8) */
static real safe_norm(const real* x, int n) {
  int is_zero = 1;
  for (int i = 0; i < n; i++) if (!(r_abs(x[i]) <= (real)1e-8)) { is_zero = 0; break; }
  if (is_zero) return 0;
  real s = 0;
  for (int i = 0; i < n; i++) s += x[i] * x[i];
  return r_sqrt(s);
}
/* math.normalize, brax/math.py:331-345 */
static inline void normalize(real* x, int n) {
  real nrm = safe_norm(x, n);
  real d = nrm + (real)1e-6 * (nrm == 0 ? (real)1 : (real)0);
  for (int i = 0; i < n; i++) x[i] = x[i] / d;
}
/* math.quat_to_3x3, brax/math.py:150-165 */
static void quat_to_3x3(const real* q, real* m) {
  real d = q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3];
  real w = q[0], x = q[1], y = q[2], z = q[3], s = 2 / d;
  real xs = x * s, ys = y * s, zs = z * s;
  real wx = w * xs, wy = w * ys, wz = w * zs, xx = x * xs, xy = x * ys, xz = x * zs;
  real yy = y * ys, yz = y * zs, zz = z * zs;
  m[0] = 1 - (yy + zz); m[1] = xy - wz; m[2] = xz + wy;
  m[3] = xy + wz; m[4] = 1 - (xx + zz); m[5] = yz - wx;
  m[6] = xz - wy; m[7] = yz + wx; m[8] = 1 - (xx + yy);
}
/* Transform.do(Transform), brax/base.py:557-562 */
static inline void tf_do_tf(const real* ap, const real* ar, const real* bp, const real* br, real* op, real* orr) {
  real t[3], q[4];
  rotate(bp, ar, t);
  quat_mul(ar, br, q);
  for (int i = 0; i < 3; i++) op[i] = ap[i] + t[i];
  for (int i = 0; i < 4; i++) orr[i] = q[i];
}
/* Inertia.mul(Motion) -> Force, brax/base.py:297-302 */
static inline void inertia_mul(const real* ipos, const real* im, real mass, const real* mang, const real* mvel, real* fang, real* fvel) {
  real c1[3], c2[3];
  cross3(ipos, mvel, c1);
  cross3(ipos, mang, c2);
  for (int i = 0; i < 3; i++) {
    fang[i] = (im[3 * i] * mang[0] + im[3 * i + 1] * mang[1] + im[3 * i + 2] * mang[2]) + c1[i];
    fvel[i] = mass * mvel[i] - c2[i];
  }
}

static int is_ancestor_or_self(const OrcModel* m, int anc, int l) {
  while (l >= 0) { if (l == anc) return 1; l = m->link_parent[l]; }
  return 0;
}

/* ---------------------------------------------------- kinematics.forward */
/* brax/kinematics.py:31-108 */
static void kin_forward(const OrcModel* m, Env* e) {
  int L = m->num_links;
  real jpos[MAXL][3], jrot[MAXL][4], jdang[MAXL][3], jdvel[MAXL][3];
  for (int l = 0; l < L; l++) {
    int qa = m->link_q_adr[l], da = m->link_qd_adr[l], nd = m->link_ndof[l];
    if (nd == 0) { /* free: kinematics.py:48-51 */
      for (int i = 0; i < 3; i++) { jpos[l][i] = e->q[qa + i]; jdvel[l][i] = e->qd[da + i]; jdang[l][i] = e->qd[da + 3 + i]; }
      for (int i = 0; i < 4; i++) jrot[l][i] = e->q[qa + 3 + i];
    } else { /* stacked 1..3 dof: kinematics.py:52-79 */
      real sp[3][3], sr[3][4], sa[3][3], sv[3][3];
      for (int k = 0; k < nd; k++) {
        int d = da + k; real qk = e->q[qa + k], qdk = e->qd[d];
        quat_rot_axis(m->dof_ang[d], qk, sr[k]);
        normalize(sr[k], 4);
        for (int i = 0; i < 3; i++) { sp[k][i] = m->dof_vel[d][i] * qk; sa[k][i] = m->dof_ang[d][i] * qdk; sv[k][i] = m->dof_vel[d][i] * qdk; }
      }
      real p[3], r[4], a[3], v[3];
      memcpy(p, sp[0], sizeof p); memcpy(r, sr[0], sizeof r); memcpy(a, sa[0], sizeof a); memcpy(v, sv[0], sizeof v);
      for (int k = 1; k < nd; k++) {
        real np_[3], nr[4], ra[3], c[3], t[3], rv[3];
        tf_do_tf(p, r, sp[k], sr[k], np_, nr);
        rotate(sa[k], sr[k], ra);
        cross3(sp[k], sa[k], c);
        for (int i = 0; i < 3; i++) t[i] = sv[k][i] + c[i];
        rotate(t, sr[k], rv);
        for (int i = 0; i < 3; i++) { a[i] = a[i] + ra[i]; v[i] = v[i] + rv[i]; p[i] = np_[i]; }
        for (int i = 0; i < 4; i++) r[i] = nr[i];
      }
      memcpy(jpos[l], p, sizeof p); memcpy(jrot[l], r, sizeof r); memcpy(jdang[l], a, sizeof a); memcpy(jdvel[l], v, sizeof v);
    }
    /* joint anchor offset, kinematics.py:85-86 */
    real anc[3];
    rotate(m->link_joint_pos[l], jrot[l], anc);
    for (int i = 0; i < 3; i++) jpos[l][i] = jpos[l][i] + m->link_joint_pos[l][i] - anc[i];
    /* link transform, kinematics.py:87 */
    real tp[3], tr[4];
    tf_do_tf(m->link_tf_pos[l], m->link_tf_rot[l], jpos[l], jrot[l], tp, tr);
    memcpy(jpos[l], tp, sizeof tp); memcpy(jrot[l], tr, sizeof tr);
  }
  /* world(), kinematics.py:89-104; parents precede children in link order */
  for (int l = 0; l < L; l++) {
    int p = m->link_parent[l];
    if (p < 0) {
      memcpy(e->x_pos[l], jpos[l], sizeof jpos[l]); memcpy(e->x_rot[l], jrot[l], sizeof jrot[l]);
      rotate(jdang[l], jrot[l], e->xd_ang[l]);
      memcpy(e->xd_vel[l], jdvel[l], sizeof jdvel[l]);
    } else {
      tf_do_tf(e->x_pos[p], e->x_rot[p], jpos[l], jrot[l], e->x_pos[l], e->x_rot[l]);
      real dlt[3], c[3], rv[3], ra[3];
      for (int i = 0; i < 3; i++) dlt[i] = e->x_pos[l][i] - e->x_pos[p][i];
      cross3(e->xd_ang[p], dlt, c);
      rotate(jdvel[l], e->x_rot[p], rv);
      rotate(jdang[l], e->x_rot[l], ra);
      for (int i = 0; i < 3; i++) {
        e->xd_vel[l][i] = (e->xd_vel[p][i] + c[i]) + rv[i];
        e->xd_ang[l][i] = e->xd_ang[p][i] + ra[i];
      }
    }
  }
  /* kinematics.py:106 -- normalised only after the whole tree is composed */
  for (int l = 0; l < L; l++) normalize(e->x_rot[l], 4);
}

/* -------------------------------------------------- dynamics.transform_com */
/* brax/generalized/dynamics.py:27-134 */
static void transform_com(const OrcModel* m, Env* e) {
  int L = m->num_links, nv = m->nv;
  real xi_pos[MAXL][3], xi_rot[MAXL][4];
  for (int l = 0; l < L; l++)
    tf_do_tf(e->x_pos[l], e->x_rot[l], m->inertia_pos[l], m->inertia_rot[l], xi_pos[l], xi_rot[l]);
  /* root_com, dynamics.py:37-43 */
  int root[MAXL]; real msum[MAXL][3], mtot[MAXL];
  for (int l = 0; l < L; l++) { int r = l; while (m->link_parent[r] >= 0) r = m->link_parent[r]; root[l] = r; mtot[l] = 0; msum[l][0] = msum[l][1] = msum[l][2] = 0; }
  for (int l = 0; l < L; l++) {
    for (int i = 0; i < 3; i++) msum[root[l]][i] += m->inertia_mass[l] * xi_pos[l][i];
    mtot[root[l]] += m->inertia_mass[l];
  }
  for (int l = 0; l < L; l++) for (int i = 0; i < 3; i++) e->root_com[l][i] = msum[root[l]][i] / mtot[root[l]];
  /* cinr = Transform(x_i.pos - root_com, x_i.rot).do(inertia), base.py:588-594 */
  for (int l = 0; l < L; l++) {
    real p[3], R[9], t[9], mass = m->inertia_mass[l];
    for (int i = 0; i < 3; i++) p[i] = xi_pos[l][i] - e->root_com[l][i];
    quat_to_3x3(xi_rot[l], R);
    /* h rows: cross(p, -e_k); (h h^T)[a][b] = sum_c h[a][c] h[b][c] */
    real h[3][3] = {{0, -p[2], p[1]}, {p[2], 0, -p[0]}, {-p[1], p[0], 0}};
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
      real s = 0;
      for (int c = 0; c < 3; c++) s += R[3 * a + c] * m->inertia_i[l][3 * c + b];
      t[3 * a + b] = s;
    }
    for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
      real s = 0, hh = 0;
      for (int c = 0; c < 3; c++) s += t[3 * a + c] * R[3 * b + c];
      for (int c = 0; c < 3; c++) hh += h[a][c] * h[b][c];
      e->cinr_i[l][3 * a + b] = s + hh * mass;
    }
    for (int i = 0; i < 3; i++) e->cinr_pos[l][i] = p[i] * mass;
    for (int i = 0; i < 4; i++) e->cinr_rot[l][i] = xi_rot[l][i];
    e->cinr_mass[l] = mass;
  }
  /* joint frames j = parent.do(link.transform).do(link.joint), dynamics.py:47-52 */
  real j_pos[MAXL][3], j_rot[MAXL][4];
  static const real ident_q[4] = {1, 0, 0, 0}, zero3[3] = {0, 0, 0};
  for (int l = 0; l < L; l++) {
    int pi = m->link_ndof[l] == 0 ? l : m->link_parent[l];
    const real* pp = pi >= 0 ? e->x_pos[pi] : zero3; const real* pr = pi >= 0 ? e->x_rot[pi] : ident_q;
    real tp[3], tr[4];
    tf_do_tf(pp, pr, m->link_tf_pos[l], m->link_tf_rot[l], tp, tr);
    tf_do_tf(tp, tr, m->link_joint_pos[l], ident_q, j_pos[l], j_rot[l]);
  }
  /* cdof_fn, dynamics.py:55-85 */
  for (int l = 0; l < L; l++) {
    int qa = m->link_q_adr[l], da = m->link_qd_adr[l], nd = m->link_ndof[l];
    if (nd == 0) {
      for (int k = 0; k < 6; k++) for (int i = 0; i < 3; i++) { e->cdof_ang[da + k][i] = m->dof_ang[da + k][i]; e->cdof_vel[da + k][i] = m->dof_vel[da + k][i]; }
    } else {
      real jp_[3] = {0, 0, 0}, jr[4] = {1, 0, 0, 0};
      for (int k = 0; k < nd; k++) {
        int d = da + k;
        /* jds[k] = j.inv_do(jd_k), base.py:573-578 */
        real a[3], v[3], c[3];
        rotate(m->dof_ang[d], jr, a);
        rotate(m->dof_vel[d], jr, v);
        cross3(jp_, a, c);
        for (int i = 0; i < 3; i++) { e->cdof_ang[d][i] = a[i]; e->cdof_vel[d][i] = v[i] + c[i]; }
        /* j = j.do(j_k) */
        real sr[4], sp[3], np_[3], nr[4];
        quat_rot_axis(m->dof_ang[d], e->q[qa + k], sr);
        normalize(sr, 4);
        for (int i = 0; i < 3; i++) sp[i] = m->dof_vel[d][i] * e->q[qa + k];
        tf_do_tf(jp_, jr, sp, sr, np_, nr);
        memcpy(jp_, np_, sizeof np_); memcpy(jr, nr, sizeof nr);
      }
    }
  }
  /* dynamics.py:86-89: rotate ang into world, shift to com */
  for (int d = 0; d < nv; d++) {
    int l = m->dof_link[d];
    real a[3], off[3], c[3];
    rotate(e->cdof_ang[d], j_rot[l], a);
    for (int i = 0; i < 3; i++) { e->cdof_ang[d][i] = a[i]; off[i] = e->root_com[l][i] - j_pos[l][i]; }
    /* Transform(pos=off, rot=identity).do(Motion): rotating by the conjugate of
       the identity quaternion is exact, so only the cross term remains */
    cross3(off, a, c);
    for (int i = 0; i < 3; i++) e->cdof_vel[d][i] = e->cdof_vel[d][i] - c[i];
  }
  /* cd forward scan, dynamics.py:92-103 */
  for (int l = 0; l < L; l++) {
    int p = m->link_parent[l], da = m->link_qd_adr[l], nd = m->link_ndof[l] == 0 ? 6 : m->link_ndof[l];
    for (int i = 0; i < 3; i++) { e->cd_ang[l][i] = p >= 0 ? e->cd_ang[p][i] : 0; e->cd_vel[l][i] = p >= 0 ? e->cd_vel[p][i] : 0; }
    for (int k = 0; k < nd; k++) for (int i = 0; i < 3; i++) {
      e->cd_ang[l][i] += e->cdof_ang[da + k][i] * e->qd[da + k];
      e->cd_vel[l][i] += e->cdof_vel[da + k][i] * e->qd[da + k];
    }
  }
  /* cdofd, dynamics.py:106-130 */
  for (int l = 0; l < L; l++) {
    int da = m->link_qd_adr[l], nd = m->link_ndof[l];
    if (nd == 0) {
      real ca[3] = {0, 0, 0}, cv[3] = {0, 0, 0};
      for (int k = 0; k < 3; k++) for (int i = 0; i < 3; i++) { ca[i] += e->cdof_ang[da + k][i] * e->qd[da + k]; cv[i] += e->cdof_vel[da + k][i] * e->qd[da + k]; }
      for (int k = 0; k < 6; k++) {
        real c1[3], c2[3], c3[3];
        cross3(ca, e->cdof_vel[da + k], c1); cross3(cv, e->cdof_ang[da + k], c2); cross3(ca, e->cdof_ang[da + k], c3);
        for (int i = 0; i < 3; i++) { e->cdofd_vel[da + k][i] = k < 3 ? 0 : c1[i] + c2[i]; e->cdofd_ang[da + k][i] = k < 3 ? 0 : c3[i]; }
      }
    } else {
      int p = m->link_parent[l];
      real ca[3], cv[3];
      for (int i = 0; i < 3; i++) { ca[i] = p >= 0 ? e->cd_ang[p][i] : 0; cv[i] = p >= 0 ? e->cd_vel[p][i] : 0; }
      for (int k = 0; k < nd; k++) {
        int d = da + k; real c1[3], c2[3], c3[3];
        /* Motion.cross(Motion), base.py:603-607 */
        cross3(ca, e->cdof_vel[d], c1); cross3(cv, e->cdof_ang[d], c2); cross3(ca, e->cdof_ang[d], c3);
        for (int i = 0; i < 3; i++) { e->cdofd_vel[d][i] = c1[i] + c2[i]; e->cdofd_ang[d][i] = c3[i]; }
        for (int i = 0; i < 3; i++) { ca[i] = ca[i] + e->cdof_ang[d][i] * e->qd[d]; cv[i] = cv[i] + e->cdof_vel[d][i] * e->qd[d]; }
      }
    }
  }
}

/* ----------------------------------------------------------- mass.matrix */
/* brax/generalized/mass.py:27-83 */
static void mass_matrix(const OrcModel* m, Env* e) {
  int L = m->num_links, nv = m->nv;
  real cp[MAXL][3], ci[MAXL][9], cm[MAXL];
  for (int l = 0; l < L; l++) { memcpy(cp[l], e->cinr_pos[l], sizeof cp[l]); memcpy(ci[l], e->cinr_i[l], sizeof ci[l]); cm[l] = e->cinr_mass[l]; }
  for (int l = L - 1; l >= 0; l--) {
    int p = m->link_parent[l];
    if (p < 0) continue;
    for (int i = 0; i < 3; i++) cp[p][i] += cp[l][i];
    for (int i = 0; i < 9; i++) ci[p][i] += ci[l][i];
    cm[p] += cm[l];
  }
  for (int i = 0; i < nv * nv; i++) e->mass_mx[i] = 0;
  for (int i = 0; i < nv; i++) {
    int li = m->dof_link[i]; real fa[3], fv[3];
    inertia_mul(cp[li], ci[li], cm[li], e->cdof_ang[i], e->cdof_vel[i], fa, fv);
    for (int j = 0; j <= i; j++) {
      if (!is_ancestor_or_self(m, m->dof_link[j], li)) continue;
      /* Motion.dot: vel.vel + ang.ang, base.py:240-241 */
      real v = dot3(e->cdof_vel[j], fv) + dot3(e->cdof_ang[j], fa);
      e->mass_mx[i * nv + j] = v;
      e->mass_mx[j * nv + i] = v;
    }
  }
  for (int i = 0; i < nv; i++) e->mass_mx[i * nv + i] += m->dof_armature[i];
}

/* SPD solve used by init (mass.py:103-104) and by integrate when
   matrix_inv_iterations == 0 (integrator.py:58-60): Cholesky, then inverse */
static void spd_inverse(const real* a, int n, real* out) {
  real Lm[MAXV * MAXV];
  memset(Lm, 0, sizeof(real) * n * n);
  for (int j = 0; j < n; j++) {
    real s = a[j * n + j];
    for (int k = 0; k < j; k++) s -= Lm[j * n + k] * Lm[j * n + k];
    real d = r_sqrt(s);
    Lm[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      real t = a[i * n + j];
      for (int k = 0; k < j; k++) t -= Lm[i * n + k] * Lm[j * n + k];
      Lm[i * n + j] = t / d;
    }
  }
  for (int c = 0; c < n; c++) {
    real y[MAXV];
    for (int i = 0; i < n; i++) {
      real t = (i == c) ? (real)1 : (real)0;
      for (int k = 0; k < i; k++) t -= Lm[i * n + k] * y[k];
      y[i] = t / Lm[i * n + i];
    }
    for (int i = n - 1; i >= 0; i--) {
      real t = y[i];
      for (int k = i + 1; k < n; k++) t -= Lm[k * n + i] * out[k * n + c];
      out[i * n + c] = t / Lm[i * n + i];
    }
  }
}

/* math.inv_approximate, brax/math.py:278-305 */
static void inv_approximate(const real* a, real* a_inv, int n, int num_iter, Env* e) {
  real r[MAXV * MAXV], b[MAXV * MAXV], xn[MAXV * MAXV], rn[MAXV * MAXV];
  /* r0 = I - a @ a_inv */
  for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
    real s = 0;
    for (int k = 0; k < n; k++) s += a[i * n + k] * a_inv[k * n + j];
    r[i * n + j] = (i == j ? (real)1 : (real)0) - s;
  }
  if (safe_norm(r, n * n) > 1) {
    real tr = 0; /* trace(a @ a.T) */
    for (int i = 0; i < n; i++) { real s = 0; for (int k = 0; k < n; k++) s += a[i * n + k] * a[i * n + k]; tr += s; }
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) a_inv[i * n + j] = (real)0.5 * a[j * n + i] / tr;
    e->ns_cold++;
  }
  real err = 1;
  for (int it = 0; it < num_iter; it++) {
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) b[i * n + j] = (i == j ? (real)1 : (real)0) + r[i * n + j];
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
      real s = 0;
      for (int k = 0; k < n; k++) s += a_inv[i * n + k] * b[k * n + j];
      xn[i * n + j] = s;
    }
    for (int i = 0; i < n; i++) for (int j = 0; j < n; j++) {
      real s = 0;
      for (int k = 0; k < n; k++) s += a[i * n + k] * xn[k * n + j];
      rn[i * n + j] = (i == j ? (real)1 : (real)0) - s;
    }
    real err_next = safe_norm(rn, n * n);
    if (err_next < err) { memcpy(a_inv, xn, sizeof(real) * n * n); e->ns_accepts++; }
    memcpy(r, rn, sizeof(real) * n * n);
    err = err_next;
  }
}

/* mass.matrix_inv, brax/generalized/mass.py:86-106 */
static void matrix_inv(const OrcModel* m, Env* e, int num_iter) {
  mass_matrix(m, e);
  if (num_iter > 0) inv_approximate(e->mass_mx, e->mass_mx_inv, m->nv, num_iter, e);
  else spd_inverse(e->mass_mx, m->nv, e->mass_mx_inv);
}

/* ------------------------------------------------------------ constraints */
/* constraint._imp_aref, brax/generalized/constraint.py:29-65 */
static void imp_aref(const real* prm, real pos, real vel, real* imp_out, real* aref_out) {
  real timeconst = prm[0], dampratio = prm[1], dmin = prm[2], dmax = prm[3], width = prm[4], mid = prm[5], power = prm[6];
  real imp_x = r_abs(pos) / width;
  real imp_a = ((real)1 / r_pow(mid, power - 1)) * r_pow(imp_x, power);
  real imp_b = 1 - ((real)1 / r_pow(1 - mid, power - 1)) * r_pow(1 - imp_x, power);
  real imp_y = imp_x < mid ? imp_a : imp_b;
  real imp = dmin + imp_y * (dmax - dmin);
  imp = r_max(dmin, r_min(imp, dmax));     /* jp.clip */
  if (imp_x > 1) imp = dmax;
  real b = 2 / (dmax * timeconst);
  real k = 1 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  real stiffness = prm[0], damping = prm[1];
  if (damping <= 0) b = -damping / dmax;
  if (stiffness <= 0) k = -stiffness / (dmax * dmax);
  *imp_out = imp;
  *aref_out = -b * vel - k * imp * pos;
}

/* contact.get for plane-sphere and plane-capsule pairs, brax/contact.py:28-67 + mjx.
 * mujoco-mjx is not under /root/reference (unpinned dependency, pyproject.toml:45);
 * restated from its published source, mjx/_src/collision_primitive.py:
 *   _plane_sphere: dist = (c - p).n - r;  pos = c - n (r + dist/2)
 *   plane_sphere:  frame = make_frame(n)  (constant per pair: precomputed on the host)
 *   plane_capsule: axis = cap.mat[:, 2]; b = normalize(axis - n (n.axis)), with the
 *                  fallback b = y if |n_y| < 0.5 else z when |b| < 0.5;
 *                  frame = [n, b, n x b]; two contacts, the end spheres at
 *                  cap.pos + axis * half and cap.pos - axis * half (in that order).
 * geom pose as brax/contact.py:48-53: pos = x.pos + rotate(geom_pos, x.rot),
 * mat = quat_to_3x3(x.rot * geom_quat).  No reference test pins plane-capsule numbers
 * (parity unpinned for this pair type). */
/* mjx math.normalize_with_norm: x / (n + 1e-6 * (n == 0)), n the safe norm */
static real normalize_with_norm(real* x) {
  real n = safe_norm(x, 3);
  real d = n + (real)1e-6 * (n == 0 ? (real)1 : (real)0);
  for (int i = 0; i < 3; i++) x[i] = x[i] / d;
  return n;
}
/* mjx math.closest_segment_point */
static void closest_segment_point(const real* a, const real* b, const real* pt, real* o) {
  real ab[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]}, pa[3] = {pt[0] - a[0], pt[1] - a[1], pt[2] - a[2]};
  real t = dot3(pa, ab) / (dot3(ab, ab) + (real)1e-6);
  t = r_max((real)0, r_min(t, (real)1));
  for (int i = 0; i < 3; i++) o[i] = a[i] + t * ab[i];
}
/* mjx math.closest_segment_to_segment_points */
static void closest_segment_to_segment_points(const real* a0, const real* a1, const real* b0, const real* b1, real* best_a, real* best_b) {
  real dir_a[3], dir_b[3], a_mid[3], b_mid[3], trans[3];
  for (int i = 0; i < 3; i++) { dir_a[i] = a1[i] - a0[i]; dir_b[i] = b1[i] - b0[i]; }
  real len_a = normalize_with_norm(dir_a), len_b = normalize_with_norm(dir_b);
  real half_a = len_a * (real)0.5, half_b = len_b * (real)0.5;
  for (int i = 0; i < 3; i++) { a_mid[i] = a0[i] + dir_a[i] * half_a; b_mid[i] = b0[i] + dir_b[i] * half_b; trans[i] = a_mid[i] - b_mid[i]; }
  real dab = dot3(dir_a, dir_b), dat = dot3(dir_a, trans), dbt = dot3(dir_b, trans);
  real denom = 1 - dab * dab;
  real orig_t_a = (-dat + dab * dbt) / (denom + (real)1e-6);
  real orig_t_b = dbt + orig_t_a * dab;
  real t_a = r_max(-half_a, r_min(orig_t_a, half_a)), t_b = r_max(-half_b, r_min(orig_t_b, half_b));
  real new_a[3], new_b[3], u[3], v[3];
  for (int i = 0; i < 3; i++) { best_a[i] = a_mid[i] + dir_a[i] * t_a; best_b[i] = b_mid[i] + dir_b[i] * t_b; }
  closest_segment_point(a0, a1, best_b, new_a);
  closest_segment_point(b0, b1, best_a, new_b);
  for (int i = 0; i < 3; i++) { u[i] = new_a[i] - best_b[i]; v[i] = best_a[i] - new_b[i]; }
  if (dot3(u, u) < dot3(v, v)) { for (int i = 0; i < 3; i++) best_a[i] = new_a[i]; }
  else { for (int i = 0; i < 3; i++) best_b[i] = new_b[i]; }
}
/* centre and axis (z column of the geom's world matrix) of a capsule attached to `link` */
static void capsule_world(const Env* e, int link, const real* gpos, const real* gquat, real* centre, real* axis) {
  real t[3], q[4], mat[9];
  static const real world_pos[3] = {0, 0, 0}, world_rot[4] = {1, 0, 0, 0};   /* contact.py:48-53 appends the world at index -1 */
  const real* xp = link >= 0 ? e->x_pos[link] : world_pos;
  const real* xr = link >= 0 ? e->x_rot[link] : world_rot;
  rotate(gpos, xr, t);
  for (int i = 0; i < 3; i++) centre[i] = xp[i] + t[i];
  quat_mul(xr, gquat, q);
  quat_to_3x3(q, mat);
  axis[0] = mat[2]; axis[1] = mat[5]; axis[2] = mat[8];
}
/* mjx collision_primitive.capsule_capsule -> _sphere_sphere on the closest points; frame = math.make_frame(n) */
static void contact_capsule_capsule(const OrcModel* m, const Env* e, int c, real* dist, real* pos, real* frame) {
  real ca[3], xa[3], cb[3], xb[3], a0[3], a1[3], b0[3], b1[3], pa[3], pb[3], n[3], b[3];
  capsule_world(e, m->con_link_a[c], m->con_a_pos[c], m->con_a_quat[c], ca, xa);
  capsule_world(e, m->con_link_b[c], m->con_sphere_pos[c], m->con_geom_quat[c], cb, xb);
  for (int i = 0; i < 3; i++) {
    real sa = xa[i] * m->con_a_half[c], sb = xb[i] * m->con_half_len[c];
    a0[i] = ca[i] - sa; a1[i] = ca[i] + sa; b0[i] = cb[i] - sb; b1[i] = cb[i] + sb;
  }
  closest_segment_to_segment_points(a0, a1, b0, b1, pa, pb);
  for (int i = 0; i < 3; i++) n[i] = pb[i] - pa[i];
  real d = normalize_with_norm(n);
  if (d == 0) { n[0] = 1; n[1] = 0; n[2] = 0; }
  real r1 = m->con_a_radius[c], r2 = m->con_radius[c];
  d = d - (r1 + r2);
  *dist = d;
  for (int i = 0; i < 3; i++) pos[i] = pa[i] + n[i] * (r1 + d * (real)0.5);
  /* make_frame: a = normalize(n); b = y (or z when a is near +-y), made orthogonal to a and normalised */
  real a[3] = {n[0], n[1], n[2]};
  normalize(a, 3);
  int use_y = (real)-0.5 < a[1] && a[1] < (real)0.5;
  b[0] = 0; b[1] = use_y ? 1 : 0; b[2] = use_y ? 0 : 1;
  real ab = dot3(a, b);
  for (int i = 0; i < 3; i++) b[i] = b[i] - a[i] * ab;
  normalize(b, 3);
  for (int i = 0; i < 3; i++) { frame[i] = a[i]; frame[3 + i] = b[i]; }
  cross3(a, b, frame + 6);
}

static void contact_get(const OrcModel* m, const Env* e, int c, real* dist, real* pos, real* frame) {
  if (m->con_kind[c] == 2) { contact_capsule_capsule(m, e, c, dist, pos, frame); return; }
  int lb = m->con_link_b[c];
  real t[3], sp[3], d[3];
  rotate(m->con_sphere_pos[c], e->x_rot[lb], t);
  for (int i = 0; i < 3; i++) sp[i] = e->x_pos[lb][i] + t[i];
  const real* n = m->con_frame[c];
  for (int i = 0; i < 9; i++) frame[i] = m->con_frame[c][i];
  if (m->con_kind[c] == 1) {
    real q[4], mat[9], axis[3], b[3];
    quat_mul(e->x_rot[lb], m->con_geom_quat[c], q);
    quat_to_3x3(q, mat);
    axis[0] = mat[2]; axis[1] = mat[5]; axis[2] = mat[8];
    for (int i = 0; i < 3; i++) sp[i] = sp[i] + axis[i] * m->con_half_len[c];
    real na = dot3(n, axis);
    for (int i = 0; i < 3; i++) b[i] = axis[i] - n[i] * na;
    real bn = safe_norm(b, 3);
    normalize(b, 3);
    if (bn < (real)0.5) {
      int use_y = (real)-0.5 < n[1] && n[1] < (real)0.5;
      b[0] = 0; b[1] = use_y ? 1 : 0; b[2] = use_y ? 0 : 1;
    }
    for (int i = 0; i < 3; i++) frame[3 + i] = b[i];
    cross3(n, b, frame + 6);
  }
  for (int i = 0; i < 3; i++) d[i] = sp[i] - m->con_plane_pos[c][i];
  *dist = dot3(d, n) - m->con_radius[c];
  for (int i = 0; i < 3; i++) pos[i] = sp[i] - n[i] * (m->con_radius[c] + (real)0.5 * *dist);
}

/* constraint.point_jacobian (vel part), constraint.py:68-98 */
static void point_jac_vel(const OrcModel* m, const Env* e, const real* pos, int link, real out[][3]) {
  int nv = m->nv;
  for (int d = 0; d < nv; d++) {
    real mask = (link >= 0 && is_ancestor_or_self(m, m->dof_link[d], link)) ? (real)1 : (real)0;
    int lw = link >= 0 ? link : m->num_links - 1; /* com[-1] wraps; masked anyway */
    real a[3], v[3], off[3], c[3];
    for (int i = 0; i < 3; i++) { a[i] = e->cdof_ang[d][i] * mask; v[i] = e->cdof_vel[d][i] * mask; off[i] = pos[i] - e->root_com[lw][i]; }
    cross3(off, a, c);
    for (int i = 0; i < 3; i++) out[d][i] = v[i] - c[i];
  }
}

/* constraint.jacobian = jac_contact ++ jac_limit, constraint.py:101-191 */
static void con_jacobian(const OrcModel* m, Env* e) {
  int nv = m->nv, row = 0;
  for (int c = 0; c < m->ncon; c++) {
    real dist, pos[3], fr[9], ja[MAXV][3], jb[MAXV][3];
    contact_get(m, e, c, &dist, pos, fr);
    e->con_dist[c] = dist;
    point_jac_vel(m, e, pos, m->con_link_a[c], ja);
    point_jac_vel(m, e, pos, m->con_link_b[c], jb);
    real mu = m->con_friction[c];
    int la = m->con_link_a[c], lb = m->con_link_b[c];
    real active = dist < 0 ? (real)1 : (real)0;
    real prm[7] = {m->con_solref[c][0], m->con_solref[c][1], m->con_solimp[c][0], m->con_solimp[c][1], m->con_solimp[c][2], m->con_solimp[c][3], m->con_solimp[c][4]};
    real t = (la > -1 ? m->link_invweight[la] : (real)0) + m->link_invweight[lb];
    int k = 0;
    for (int di = 1; di <= 2; di++) for (int s = 0; s < 2; s++, k++) {
      real f = s == 0 ? -mu : mu, dir[3];
      for (int i = 0; i < 3; i++) dir[i] = (-fr[3 * di + i]) * f + fr[i];
      real vel = 0;
      real* jr = &e->con_jac[(row + k) * nv];
      for (int d = 0; d < nv; d++) {
        real df[3] = {jb[d][0] - ja[d][0], jb[d][1] - ja[d][1], jb[d][2] - ja[d][2]};
        jr[d] = dot3(df, dir);
      }
      for (int d = 0; d < nv; d++) vel += jr[d] * e->qd[d];
      real imp, aref;
      imp_aref(prm, dist, vel, &imp, &aref);
      real diag = (t + mu * mu * t) * (2 * mu * mu * (1 - imp) / (imp + (real)1e-8));
      for (int d = 0; d < nv; d++) jr[d] = jr[d] * active;
      e->con_diag[row + k] = diag * active;
      e->con_aref[row + k] = aref * active;
    }
    row += 4;
  }
  if (m->nlim > 0) {
    for (int l = 0; l < m->num_links; l++) {
      int nd = m->link_ndof[l];
      for (int k = 0; k < nd; k++) {
        int d = m->link_qd_adr[l] + k, qi = m->link_q_adr[l] + k;
        real pos_min = e->q[qi] - m->dof_limit_lo[d], pos_max = m->dof_limit_hi[d] - e->q[qi];
        real pos = r_min(r_min(pos_min, pos_max), 0);
        real active = pos < 0 ? (real)1 : (real)0;
        real side = (real)((pos_min < pos_max) * 2 - 1) * active;
        real* jr = &e->con_jac[row * nv];
        for (int j = 0; j < nv; j++) jr[j] = (j == d ? (real)1 : (real)0) * side;
        real vel = 0;
        for (int j = 0; j < nv; j++) vel += jr[j] * e->qd[j];
        real imp, aref;
        imp_aref(m->dof_solver_params[d], pos, vel, &imp, &aref);
        e->con_diag[row] = m->dof_invweight[d] * active * (1 - imp) / (imp + (real)1e-8);
        e->con_aref[row] = aref * active;
        row++;
      }
    }
  }
}

/* constraint.force + jaxopt.ProjectedGradient, constraint.py:194-240.
 * jaxopt (not in /root/reference) restated from its published source,
 * jaxopt/_src/proximal_gradient.py: FISTA acceleration, backtracking line search
 * (stepsize carried across iterations, decrease_factor 0.5, at most maxls
 * halvings), stop when ||prox(x - grad(x)) - x|| <= tol = 1e-3 after the first
 * unconditional iteration, at most maxiter iterations.  prox = max(., 0). */
static real pg_fun(const real* A, const real* b, const real* x, int n, real* res) {
  real f = 0;
  for (int i = 0; i < n; i++) {
    real s = 0;
    for (int j = 0; j < n; j++) s += A[i * n + j] * x[j];
    res[i] = s + b[i];
  }
  for (int i = 0; i < n; i++) f += (real)0.5 * (res[i] * res[i]);
  return f;
}
static void pg_grad_from_res(const real* A, const real* res, int n, real* g) { /* A^T res */
  for (int j = 0; j < n; j++) {
    real s = 0;
    for (int i = 0; i < n; i++) s += A[i * n + j] * res[i];
    g[j] = s;
  }
}
/* jaxopt.ProjectedGradient(...).run(zeros) with projection_non_negative */
static void pg_solve(const real* A, const real* b, int nc, int maxiter, int maxls, real* x, int32_t* iters, int32_t* trials) {
  real y[MAXC], g[MAXC], res[MAXC], xn[MAXC], gn[MAXC];
  for (int i = 0; i < nc; i++) { x[i] = 0; y[i] = 0; }
  real t = 1, stepsize = 1, error = INFINITY;
  const real tol = (real)1e-3;
  int it = 0;
  while (it < maxiter && (it == 0 || error > tol)) {
    /* _update_accel: value and grad at the velocity point y */
    real fy = pg_fun(A, b, y, nc, res);
    pg_grad_from_res(A, res, nc, g);
    /* _ls: backtracking from the carried stepsize */
    real s = stepsize;
    for (int i = 0; i < nc; i++) xn[i] = r_max(y[i] - s * g[i], 0);
    for (int ls = 0; ; ls++) {
      real sqdist = 0, vd = 0, resn[MAXC];
      for (int i = 0; i < nc; i++) { real d = xn[i] - y[i]; sqdist += d * d; }
      for (int i = 0; i < nc; i++) { real d = xn[i] - y[i]; vd += d * g[i]; }
      real fn = pg_fun(A, b, xn, nc, resn);
      if (trials) (*trials)++;
      real fun_decrease = s * (fn - fy);
      real condition = s * vd + (real)0.5 * sqdist;
      if (!(fun_decrease > condition + EPS_REAL) || ls >= maxls) break;
      s = s * (real)0.5;
      for (int i = 0; i < nc; i++) xn[i] = r_max(y[i] - s * g[i], 0);
    }
    stepsize = s <= (real)1e-6 ? (real)1 : s / (real)0.5;
    real tn = (real)0.5 * (1 + r_sqrt(1 + 4 * (t * t)));
    real mom = (t - 1) / tn;
    for (int i = 0; i < nc; i++) { real d = xn[i] - x[i]; y[i] = xn[i] + mom * d; }
    /* error = ||prox(x+ - grad f(x+)) - x+|| */
    pg_fun(A, b, xn, nc, res);
    pg_grad_from_res(A, res, nc, gn);
    real err2 = 0;
    for (int i = 0; i < nc; i++) { real d = r_max(xn[i] - gn[i], 0) - xn[i]; err2 += d * d; }
    error = r_sqrt(err2);
    for (int i = 0; i < nc; i++) x[i] = xn[i];
    t = tn;
    it++;
    if (iters) (*iters)++;
  }
}

static void con_force(const OrcModel* m, Env* e) {
  int nv = m->nv, nc = 4 * m->ncon + m->nlim;
  if (nc == 0) { for (int i = 0; i < nv; i++) e->qf_constraint[i] = 0; return; }
  static __thread real A[MAXC * MAXC], JM[MAXC * MAXV];
  real b[MAXC], x[MAXC];
  const real* J = e->con_jac; const real* Mi = e->mass_mx_inv;
  for (int i = 0; i < nc; i++) for (int j = 0; j < nv; j++) {
    real s = 0;
    for (int k = 0; k < nv; k++) s += J[i * nv + k] * Mi[k * nv + j];
    JM[i * nv + j] = s;
  }
  for (int i = 0; i < nc; i++) {
    for (int j = 0; j < nc; j++) {
      real s = 0;
      for (int k = 0; k < nv; k++) s += JM[i * nv + k] * J[j * nv + k];
      A[i * nc + j] = s + (i == j ? e->con_diag[i] : (real)0);
    }
    real s = 0;
    for (int k = 0; k < nv; k++) s += JM[i * nv + k] * e->qf_smooth[k];
    b[i] = s - e->con_aref[i];
  }
  pg_solve(A, b, nc, m->solver_iterations, m->solver_maxls, x, &e->pg_iters, &e->pg_ls_trials);
  for (int j = 0; j < nv; j++) {
    real s = 0;
    for (int i = 0; i < nc; i++) s += J[i * nv + j] * x[i];
    e->qf_constraint[j] = s;
  }
}

/* ---------------------------------------------------------------- dynamics */
/* actuator.to_tau, brax/actuator.py:23-57 */
static void to_tau(const OrcModel* m, const Env* e, const real* act, real* tau) {
  for (int i = 0; i < m->nv; i++) tau[i] = 0;
  for (int a = 0; a < m->nu; a++) {
    real qv = e->q[m->act_q_id[a]], qdv = e->qd[m->act_qd_id[a]];
    real c = r_max(m->act_ctrl_lo[a], r_min(act[a], m->act_ctrl_hi[a]));
    real bias = m->act_gear[a] * (qv * m->act_bias_q[a] + qdv * m->act_bias_qd[a]);
    real f = m->act_gain[a] * c + bias;
    f = r_max(m->act_force_lo[a], r_min(f, m->act_force_hi[a]));
    f = f * m->act_gear[a];
    tau[m->act_qd_id[a]] += f;
  }
}

/* fluid.force (brax/fluid.py:24-91) as dynamics._passive uses it (dynamics.py:198-211): the
 * viscous + inertial drag of each link's inertia box, projected on the dofs through the point
 * jacobian of the link's centre of mass.  out[d] = sum over links of jac[l][d] . frc[l] */
static void fluid_passive(const OrcModel* m, const Env* e, real* out) {
  int L = m->num_links, nv = m->nv;
  const real pi = (real)3.14159265358979323846;
  for (int d = 0; d < nv; d++) out[d] = 0;
  for (int l = 0; l < L; l++) {
    /* x_i = x.do(inertia.transform); offset = x_i.pos - root_com; xd_i = Transform(offset, x_i.rot).do(cd) */
    real xi_pos[3], xi_rot[4], off[3], rinv[4], c[3], t[3], ang[3], vel[3];
    tf_do_tf(e->x_pos[l], e->x_rot[l], m->inertia_pos[l], m->inertia_rot[l], xi_pos, xi_rot);
    for (int i = 0; i < 3; i++) off[i] = xi_pos[i] - e->root_com[l][i];
    rinv[0] = xi_rot[0]; rinv[1] = -xi_rot[1]; rinv[2] = -xi_rot[2]; rinv[3] = -xi_rot[3];
    rotate(e->cd_ang[l], rinv, ang);
    cross3(off, e->cd_ang[l], c);
    for (int i = 0; i < 3; i++) t[i] = e->cd_vel[l][i] - c[i];
    rotate(t, rinv, vel);
    /* box from the diagonal inertia (fluid.py:73-77) */
    const real* I = m->inertia_i[l];
    real dg[3] = {I[0], I[4], I[8]}, box[3];
    for (int i = 0; i < 3; i++) {
      real sum = 0;
      for (int j = 0; j < 3; j++) sum += dg[j] * (i == j ? (real)-1 : (real)1);
      sum = (real)6 * r_max(sum, (real)1e-12);
      box[i] = r_sqrt(sum / m->inertia_mass[l]);
    }
    /* _box_viscosity + _box_density (fluid.py:24-53) */
    real diam = (box[0] + box[1] + box[2]) / (real)3;
    real ang_scale = -pi * (diam * diam * diam) * m->viscosity, vel_scale = (real)(-3.0 * 3.14159265358979323846) * diam * m->viscosity;
    real bmv[3] = {box[1] * box[2], box[0] * box[2], box[0] * box[1]};
    real p2[3] = {box[0] * box[0], box[1] * box[1], box[2] * box[2]};
    real p4[3] = {p2[0] * p2[0], p2[1] * p2[1], p2[2] * p2[2]};   /* x ** 4 by repeated squaring */
    real bma[3] = {box[0] * (p4[1] + p4[2]), box[1] * (p4[0] + p4[2]), box[2] * (p4[0] + p4[1])};
    real fa[3], fv[3], wa[3], wv[3];
    for (int i = 0; i < 3; i++) {
      real dv = (real)-0.5 * m->density * bmv[i] * r_abs(vel[i]) * vel[i];
      real da = (real)-1.0 * m->density * bma[i] * r_abs(ang[i]) * ang[i] / (real)64.0;
      fa[i] = ang_scale * ang[i] + da; fv[i] = vel_scale * vel[i] + dv;
    }
    /* back to the world orientation: Transform.create(rot=x_i.rot).do(Force), base.py:581-585 */
    rotate(fv, xi_rot, wv); rotate(fa, xi_rot, wa);
    /* point jacobian of x_i.pos on link l (constraint.py:68-98), dotted with the force */
    for (int d = 0; d < nv; d++) {
      if (!is_ancestor_or_self(m, m->dof_link[d], l)) continue;
      real jv[3];
      cross3(off, e->cdof_ang[d], c);
      for (int i = 0; i < 3; i++) jv[i] = e->cdof_vel[d][i] - c[i];
      out[d] += dot3(jv, wv) + dot3(e->cdof_ang[d], wa);
    }
  }
}

/* dynamics.forward = passive - inverse(RNE) + tau, dynamics.py:137-236 */
static void dyn_forward(const OrcModel* m, Env* e, const real* tau) {
  int L = m->num_links, nv = m->nv;
  real cdd_a[MAXL][3], cdd_v[MAXL][3], cf_a[MAXL][3], cf_v[MAXL][3], fluid[MAXV];
  if (m->enable_fluid) fluid_passive(m, e, fluid);
  for (int l = 0; l < L; l++) {
    int p = m->link_parent[l], da = m->link_qd_adr[l], nd = m->link_ndof[l] == 0 ? 6 : m->link_ndof[l];
    for (int i = 0; i < 3; i++) { cdd_a[l][i] = p >= 0 ? cdd_a[p][i] : 0; cdd_v[l][i] = p >= 0 ? cdd_v[p][i] : -m->gravity[i]; }
    for (int k = 0; k < nd; k++) for (int i = 0; i < 3; i++) {
      cdd_a[l][i] += e->cdofd_ang[da + k][i] * e->qd[da + k];
      cdd_v[l][i] += e->cdofd_vel[da + k][i] * e->qd[da + k];
    }
  }
  for (int l = 0; l < L; l++) {
    real fa[3], fv[3], ga[3], gv[3], c1[3], c2[3], c3[3];
    inertia_mul(e->cinr_pos[l], e->cinr_i[l], e->cinr_mass[l], cdd_a[l], cdd_v[l], fa, fv);
    inertia_mul(e->cinr_pos[l], e->cinr_i[l], e->cinr_mass[l], e->cd_ang[l], e->cd_vel[l], ga, gv);
    /* Motion.cross(Force), base.py:610-614 */
    cross3(e->cd_ang[l], gv, c1); cross3(e->cd_ang[l], ga, c2); cross3(e->cd_vel[l], gv, c3);
    for (int i = 0; i < 3; i++) { cf_v[l][i] = fv[i] + c1[i]; cf_a[l][i] = fa[i] + (c2[i] + c3[i]); }
  }
  for (int l = L - 1; l >= 0; l--) {
    int p = m->link_parent[l];
    if (p < 0) continue;
    for (int i = 0; i < 3; i++) { cf_a[p][i] += cf_a[l][i]; cf_v[p][i] += cf_v[l][i]; }
  }
  for (int d = 0; d < nv; d++) {
    int l = m->dof_link[d];
    real bias = dot3(e->cdof_vel[d], cf_v[l]) + dot3(e->cdof_ang[d], cf_a[l]);
    int is_free = m->link_ndof[l] == 0;
    int qi = m->link_q_adr[l] + (d - m->link_qd_adr[l]);
    real passive = is_free ? (real)0 : -e->q[qi] * m->dof_stiffness[d];
    passive = passive - m->dof_damping[d] * e->qd[d];
    if (m->enable_fluid) passive = passive + fluid[d];
    e->qf_smooth[d] = (passive - bias) + tau[d];
  }
}

/* integrator.integrate, brax/generalized/integrator.py:46-84 */
static void integrate(const OrcModel* m, Env* e) {
  int nv = m->nv; real dt = m->dt;
  real mi_local[MAXV * MAXV]; const real* Mi = e->mass_mx_inv;
  if (m->matrix_inv_iterations == 0) {
    real mx[MAXV * MAXV];
    for (int i = 0; i < nv * nv; i++) mx[i] = e->mass_mx[i];
    for (int i = 0; i < nv; i++) mx[i * nv + i] += m->dof_damping[i] * dt;
    spd_inverse(mx, nv, mi_local);
    Mi = mi_local;
  }
  for (int i = 0; i < nv; i++) {
    real s = 0;
    for (int j = 0; j < nv; j++) s += Mi[i * nv + j] * (e->qf_smooth[j] + e->qf_constraint[j]);
    e->qdd[i] = s;
  }
  for (int i = 0; i < nv; i++) e->qd[i] = e->qd[i] + e->qdd[i] * dt;
  for (int l = 0; l < m->num_links; l++) {
    int qa = m->link_q_adr[l], da = m->link_qd_adr[l], nd = m->link_ndof[l];
    if (nd == 0) {
      real* rot = &e->q[qa + 3]; const real* ang = &e->qd[da + 3];
      real ang_norm = r_sqrt(ang[0] * ang[0] + ang[1] * ang[1] + ang[2] * ang[2]) + (real)1e-8;
      real axis[3] = {ang[0] / ang_norm, ang[1] / ang_norm, ang[2] / ang_norm};
      real qrot[4], nr[4];
      quat_rot_axis(axis, dt * ang_norm, qrot);
      quat_mul(rot, qrot, nr);
      real n = r_sqrt(nr[0] * nr[0] + nr[1] * nr[1] + nr[2] * nr[2] + nr[3] * nr[3]);
      for (int i = 0; i < 4; i++) rot[i] = nr[i] / n;
      for (int i = 0; i < 3; i++) e->q[qa + i] = e->q[qa + i] + e->qd[da + i] * dt;
    } else {
      for (int k = 0; k < nd; k++) e->q[qa + k] = e->q[qa + k] + e->qd[da + k] * dt;
    }
  }
}

/* ---------------------------------------------------------------- pipeline */
/* pipeline.init, brax/generalized/pipeline.py:32-61 */
static void env_init(const OrcModel* m, Env* e) {
  kin_forward(m, e);
  transform_com(m, e);
  matrix_inv(m, e, 0);
  con_jacobian(m, e);
  for (int i = 0; i < m->nv; i++) { e->qf_smooth[i] = 0; e->qf_constraint[i] = 0; e->qdd[i] = 0; }
}
/* pipeline.step, brax/generalized/pipeline.py:64-94 */
static void env_step(const OrcModel* m, Env* e, const real* act) {
  real tau[MAXV];
  to_tau(m, e, act, tau);
  dyn_forward(m, e, tau);
  con_force(m, e);
  integrate(m, e);
  kin_forward(m, e);
  transform_com(m, e);
  matrix_inv(m, e, m->matrix_inv_iterations);
  con_jacobian(m, e);
}

/* ------------------------------------------------------ batch load / store */
#define CP(dst, src, n) memcpy((dst), (src), sizeof(real) * (size_t)(n))
static void env_load(const OrcModel* m, const OrcState* s, long i, Env* e) {
  int L = m->num_links, nq = m->nq, nv = m->nv, nc = 4 * m->ncon + m->nlim;
  CP(e->q, s->q + i * nq, nq); CP(e->qd, s->qd + i * nv, nv);
  CP(e->x_pos, s->x_pos + i * L * 3, L * 3); CP(e->x_rot, s->x_rot + i * L * 4, L * 4);
  CP(e->xd_ang, s->xd_ang + i * L * 3, L * 3); CP(e->xd_vel, s->xd_vel + i * L * 3, L * 3);
  CP(e->root_com, s->root_com + i * L * 3, L * 3);
  CP(e->cinr_pos, s->cinr_pos + i * L * 3, L * 3); CP(e->cinr_rot, s->cinr_rot + i * L * 4, L * 4);
  CP(e->cinr_i, s->cinr_i + i * L * 9, L * 9); CP(e->cinr_mass, s->cinr_mass + i * L, L);
  CP(e->cd_ang, s->cd_ang + i * L * 3, L * 3); CP(e->cd_vel, s->cd_vel + i * L * 3, L * 3);
  CP(e->cdof_ang, s->cdof_ang + i * nv * 3, nv * 3); CP(e->cdof_vel, s->cdof_vel + i * nv * 3, nv * 3);
  CP(e->cdofd_ang, s->cdofd_ang + i * nv * 3, nv * 3); CP(e->cdofd_vel, s->cdofd_vel + i * nv * 3, nv * 3);
  CP(e->mass_mx, s->mass_mx + i * nv * nv, nv * nv); CP(e->mass_mx_inv, s->mass_mx_inv + i * nv * nv, nv * nv);
  CP(e->con_jac, s->con_jac + i * nc * nv, nc * nv); CP(e->con_diag, s->con_diag + i * nc, nc); CP(e->con_aref, s->con_aref + i * nc, nc);
  CP(e->qf_smooth, s->qf_smooth + i * nv, nv); CP(e->qf_constraint, s->qf_constraint + i * nv, nv); CP(e->qdd, s->qdd + i * nv, nv);
  e->pg_iters = e->pg_ls_trials = e->ns_accepts = e->ns_cold = 0;
}
static void env_store(const OrcModel* m, OrcState* s, long i, const Env* e) {
  int L = m->num_links, nq = m->nq, nv = m->nv, nc = 4 * m->ncon + m->nlim;
  CP(s->q + i * nq, e->q, nq); CP(s->qd + i * nv, e->qd, nv);
  CP(s->x_pos + i * L * 3, e->x_pos, L * 3); CP(s->x_rot + i * L * 4, e->x_rot, L * 4);
  CP(s->xd_ang + i * L * 3, e->xd_ang, L * 3); CP(s->xd_vel + i * L * 3, e->xd_vel, L * 3);
  CP(s->root_com + i * L * 3, e->root_com, L * 3);
  CP(s->cinr_pos + i * L * 3, e->cinr_pos, L * 3); CP(s->cinr_rot + i * L * 4, e->cinr_rot, L * 4);
  CP(s->cinr_i + i * L * 9, e->cinr_i, L * 9); CP(s->cinr_mass + i * L, e->cinr_mass, L);
  CP(s->cd_ang + i * L * 3, e->cd_ang, L * 3); CP(s->cd_vel + i * L * 3, e->cd_vel, L * 3);
  CP(s->cdof_ang + i * nv * 3, e->cdof_ang, nv * 3); CP(s->cdof_vel + i * nv * 3, e->cdof_vel, nv * 3);
  CP(s->cdofd_ang + i * nv * 3, e->cdofd_ang, nv * 3); CP(s->cdofd_vel + i * nv * 3, e->cdofd_vel, nv * 3);
  CP(s->mass_mx + i * nv * nv, e->mass_mx, nv * nv); CP(s->mass_mx_inv + i * nv * nv, e->mass_mx_inv, nv * nv);
  CP(s->con_jac + i * nc * nv, e->con_jac, nc * nv); CP(s->con_diag + i * nc, e->con_diag, nc); CP(s->con_aref + i * nc, e->con_aref, nc);
  CP(s->qf_smooth + i * nv, e->qf_smooth, nv); CP(s->qf_constraint + i * nv, e->qf_constraint, nv); CP(s->qdd + i * nv, e->qdd, nv);
  if (s->con_dist) CP(s->con_dist + i * m->ncon, e->con_dist, m->ncon);
  if (s->stats) { int32_t* st = s->stats + i * 4; st[0] += e->pg_iters; st[1] += e->pg_ls_trials; st[2] += e->ns_accepts; st[3] += e->ns_cold; }
}

/* ------------------------------------------------------------ entry points */
int orc_sizeof_real(void) { return (int)sizeof(real); }
int orc_sizeof_model(void) { return (int)sizeof(OrcModel); }
int orc_max_links(void) { return MAXL; }
/* Threads of the batch loops below, set explicitly: launchers such as torch.distributed.run
 * export OMP_NUM_THREADS=1, which would silently serialise the CPU baseline. Returns the count in effect. */
#ifdef _OPENMP
#include <omp.h>
int orc_set_threads(int n) { if (n > 0) omp_set_num_threads(n); return omp_get_max_threads(); }
#else
int orc_set_threads(int n) { (void)n; return 1; }
#endif

/* pipeline.init over a batch: q [n,nq], qd [n,nv] -> full state */
int orc_init(const OrcModel* m, long n_env, const real* q, const real* qd, OrcState* out) {
  if (m->num_links > MAXL || m->nv > MAXV || m->nq > MAXQ || m->nu > MAXU || m->ncon > MAXCON) return 1;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n_env; i++) {
    Env* e = (Env*)calloc(1, sizeof(Env));
    CP(e->q, q + i * m->nq, m->nq); CP(e->qd, qd + i * m->nv, m->nv);
    env_init(m, e);
    env_store(m, out, i, e);
    free(e);
  }
  return 0;
}

/* n_frames x pipeline.step with the same action (envs/base.py:128-137) */
int orc_step(const OrcModel* m, long n_env, int n_frames, OrcState* s, const real* act) {
  if (m->num_links > MAXL || m->nv > MAXV || m->nq > MAXQ || m->nu > MAXU || m->ncon > MAXCON) return 1;
#pragma omp parallel for schedule(static)
  for (long i = 0; i < n_env; i++) {
    Env* e = (Env*)calloc(1, sizeof(Env));
    env_load(m, s, i, e);
    for (int f = 0; f < n_frames; f++) env_step(m, e, act + i * m->nu);
    env_store(m, s, i, e);
    free(e);
  }
  return 0;
}

/* component entry points for the known-answer tests (single env) */
int orc_to_tau(const OrcModel* m, const real* q, const real* qd, const real* act, real* tau) {
  Env* e = (Env*)calloc(1, sizeof(Env));
  CP(e->q, q, m->nq); CP(e->qd, qd, m->nv);
  to_tau(m, e, act, tau);
  free(e);
  return 0;
}
int orc_inv_approximate(const real* a, real* a_inv, int n, int num_iter) {
  Env* e = (Env*)calloc(1, sizeof(Env));
  inv_approximate(a, a_inv, n, num_iter, e);
  free(e);
  return 0;
}
/* contact.get for one env: x from (q, qd=0) via kinematics.forward */
int orc_contact(const OrcModel* m, const real* q, real* dist, real* pos) {
  Env* e = (Env*)calloc(1, sizeof(Env));
  CP(e->q, q, m->nq);
  kin_forward(m, e);
  for (int c = 0; c < m->ncon; c++) { real fr[9]; contact_get(m, e, c, dist + c, pos + 3 * c, fr); }
  free(e);
  return 0;
}
/* the projected-gradient solver alone (used by the solver property tests) */
int orc_pg_solve(const real* A, const real* b, int nc, int maxiter, int maxls, real* x_out, int32_t* stats2) {
  if (nc > MAXC) return 1;
  stats2[0] = stats2[1] = 0;
  pg_solve(A, b, nc, maxiter, maxls, x_out, &stats2[0], &stats2[1]);
  return 0;
}
