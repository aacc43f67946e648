/*
 * bxg.h -- C ABI of the B200-native `generalized` physics step.
 *
 * Drop-in boundary for google/brax 0.14.2's generalized pipeline.  The
 * reference has no FFI for this path: the seam is a Python module with two
 * functions selected by a backend string (brax/envs/base.py:104-114),
 *
 *   generalized.pipeline.init(sys, q, qd, ...)  -> State   pipeline.py:32-61
 *   generalized.pipeline.step(sys, state, act)  -> State   pipeline.py:64-94
 *
 * called under vmap over envs (envs/wrappers/training.py:66-72) and looped
 * n_frames times with one action (envs/base.py:128-137).  The entry points
 * below are what a JAX-FFI / ctypes binding for that seam binds: the batch
 * axis and the n_frames loop are inside the call.
 *
 * Conventions
 *   - every `float*` / `int32_t*` in BxgState and in the call arguments is a
 *     DEVICE pointer to a dense env-major array [n_env, ...] (fp32), caller
 *     allocated; BxgModelDesc pointers are HOST pointers, copied at create time
 *   - calls are asynchronous on `stream` (a cudaStream_t passed as void*),
 *     never allocate, never synchronise, are CUDA-graph capturable and
 *     re-entrant across models/devices
 *   - return 0 on success, non-zero BXG_E_* otherwise; bxg_last_error() gives
 *     a thread-local message.  There is no CPU fallback: without a CUDA device
 *     every compute entry point returns BXG_E_CUDA.
 */
#ifndef BXG_H_
#define BXG_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BXG_ABI_VERSION 4

enum {
  BXG_OK = 0,
  BXG_E_INVALID = 1,      /* bad argument / unsupported model            */
  BXG_E_CUDA = 2,         /* CUDA runtime error (incl. no device)        */
  BXG_E_UNSUPPORTED = 3   /* model outside the kernel's compiled limits  */
};

/* bxg_step flags */
enum {
  BXG_STEP_DEFAULT = 0,
  /* also write the per-contact penetration distance and solver statistics */
  BXG_STEP_DIAGNOSTICS = 1,
  /* Lean state I/O.  Of `in` only q, qd, x_pos, x_rot and mass_mx_inv are read: every other leaf is a pure
   * function of (q, qd) (pipeline.py:51-61,86-93) and is recomputed in shared memory at entry by the code that
   * would have produced it; of `out` only q, qd, x_pos, x_rot, xd_ang, xd_vel and mass_mx_inv are written (the
   * other leaves may be NULL).  Those leaves are bit-identical to a default step's: a rollout needs about a
   * quarter of the State in HBM (Ant 1.4 KB instead of 5.5 KB per env) and half of the DRAM traffic. */
  BXG_STEP_LEAN = 2
};

/* How mass_mx_inv is produced inside step (reference mass.py:86-106).
 * NEWTON_SCHULZ is the reference algorithm (math.py:278-305): warm-started
 * from the incoming mass_mx_inv, `matrix_inv_iterations` iterations.       */
enum {
  BXG_MINV_NEWTON_SCHULZ = 0,
  BXG_MINV_CHOLESKY = 1   /* exact SPD inverse every substep (north_star) */
};

#define BXG_CON_PLANE_SPHERE 0
#define BXG_CON_PLANE_CAPSULE_END 1
#define BXG_CON_CAPSULE_CAPSULE 2   /* one contact between the closest points of two capsules; link_a may move */

/* Model constants = the fields of brax.base.System the path reads
 * (brax/base.py:415-540).  Host pointers. */
typedef struct BxgModelDesc {
  int32_t abi_version;           /* BXG_ABI_VERSION */
  int32_t num_links, nq, nv, nu; /* len(link_types), q_size, qd_size, act_size */
  int32_t ncon;                  /* static contact count (mjx.make_data(sys).ncon) */
  int32_t has_limit;             /* sys.dof.limit is not None */
  int32_t solver_iterations, solver_maxls, matrix_inv_iterations;
  int32_t minv_mode;             /* BXG_MINV_* */
  float dt;                      /* sys.opt.timestep */
  float gravity[3];              /* sys.gravity */
  /* links [num_links] */
  const int32_t* link_parent;    /* sys.link_parents (-1 = root) */
  const int32_t* link_ndof;      /* 0 = free joint ('f'), else 1..3 */
  const float* link_tf_pos;      /* [L,3] sys.link.transform.pos */
  const float* link_tf_rot;      /* [L,4] sys.link.transform.rot */
  const float* link_joint_pos;   /* [L,3] sys.link.joint.pos */
  const float* inertia_pos;      /* [L,3] sys.link.inertia.transform.pos */
  const float* inertia_rot;      /* [L,4] sys.link.inertia.transform.rot */
  const float* inertia_i;        /* [L,3,3] sys.link.inertia.i */
  const float* inertia_mass;     /* [L] */
  const float* link_invweight;   /* [L] */
  /* dofs [nv] */
  const float* dof_ang;          /* [nv,3] sys.dof.motion.ang */
  const float* dof_vel;          /* [nv,3] sys.dof.motion.vel */
  const float* dof_armature;
  const float* dof_stiffness;
  const float* dof_damping;
  const float* dof_limit_lo;     /* may be NULL when !has_limit */
  const float* dof_limit_hi;
  const float* dof_invweight;
  const float* dof_solver_params; /* [nv,7] */
  /* actuators [nu] */
  const int32_t* act_q_id;
  const int32_t* act_qd_id;
  const float* act_gain;
  const float* act_gear;
  const float* act_ctrl_lo;
  const float* act_ctrl_hi;
  const float* act_force_lo;
  const float* act_force_hi;
  const float* act_bias_q;
  const float* act_bias_qd;
  /* plane-sphere / plane-capsule / capsule-capsule contacts [ncon] (brax/contact.py:28-67 + mjx collision).
   * A capsule contributes two contacts (its end spheres, +axis first). */
  const int32_t* con_link_a;     /* plane link (-1 = world); kind 2: link of the first capsule */
  const int32_t* con_link_b;     /* sphere / capsule link */
  const float* con_plane_pos;    /* [ncon,3] world */
  const float* con_frame;        /* [ncon,3,3] rows normal,t1,t2 (capsules: t1, t2 follow the axis at run time) */
  const float* con_sphere_pos;   /* [ncon,3] geom centre in link_b frame */
  const float* con_radius;
  const float* con_friction;     /* sliding friction mu */
  const float* con_solref;       /* [ncon,2] */
  const float* con_solimp;       /* [ncon,5] */
  const int32_t* con_kind;       /* [ncon] BXG_CON_*; NULL = all plane-sphere */
  const float* con_geom_quat;    /* [ncon,4] capsule orientation in link_b frame (kind 1) */
  const float* con_half_len;     /* [ncon] signed half length: end point = centre + axis * half_len (kind 1) */
  /* fluid forces on the links' inertia boxes (brax/fluid.py:24-91, dynamics.py:198-211) */
  int32_t enable_fluid;          /* sys.enable_fluid = viscosity > 0 or density > 0 (io/mjcf.py:467) */
  float viscosity, density;      /* sys.viscosity, sys.density */
  /* capsule-capsule pairs (kind 2): geom1's shape in the frame of con_link_a (which may be a moving link);
   * geom2's is con_sphere_pos / con_geom_quat / con_half_len / con_radius.  NULL when no pair has kind 2. */
  const float* con_a_pos;        /* [ncon,3] */
  const float* con_a_quat;       /* [ncon,4] */
  const float* con_a_half;       /* [ncon] half length */
  const float* con_a_radius;     /* [ncon] */
} BxgModelDesc;

/* The generalized State (brax/generalized/base.py:25-92 + brax/base.py:396-412)
 * as one dense fp32 array per leaf, env-major.  nc = 4*ncon + nlim where nlim
 * is the number of non-free dofs when has_limit, else 0. */
typedef struct BxgState {
  float* q;             /* [n, nq] */
  float* qd;            /* [n, nv] */
  float* x_pos;         /* [n, L, 3] */
  float* x_rot;         /* [n, L, 4] */
  float* xd_ang;        /* [n, L, 3] */
  float* xd_vel;        /* [n, L, 3] */
  float* root_com;      /* [n, L, 3] */
  float* cinr_pos;      /* [n, L, 3]   cinr.transform.pos */
  float* cinr_rot;      /* [n, L, 4]   cinr.transform.rot */
  float* cinr_i;        /* [n, L, 3, 3] */
  float* cinr_mass;     /* [n, L] */
  float* cd_ang;        /* [n, L, 3] */
  float* cd_vel;        /* [n, L, 3] */
  float* cdof_ang;      /* [n, nv, 3] */
  float* cdof_vel;      /* [n, nv, 3] */
  float* cdofd_ang;     /* [n, nv, 3] */
  float* cdofd_vel;     /* [n, nv, 3] */
  float* mass_mx;       /* [n, nv, nv] */
  float* mass_mx_inv;   /* [n, nv, nv] */
  float* con_jac;       /* [n, nc, nv] */
  float* con_diag;      /* [n, nc] */
  float* con_aref;      /* [n, nc] */
  float* qf_smooth;     /* [n, nv] */
  float* qf_constraint; /* [n, nv] */
  float* qdd;           /* [n, nv] */
} BxgState;

/* Optional outputs written when BXG_STEP_DIAGNOSTICS is set (may be NULL). */
typedef struct BxgDiag {
  float* con_dist;      /* [n, ncon]  contact.dist after the last substep   */
  int32_t* stats;       /* [n, 4]  accumulated: pg iterations, line-search
                           evaluations, Newton-Schulz accepts, cold starts  */
  unsigned long long* phase_cycles;  /* [BXG_NUM_PHASES] per-warp cycles summed per phase of the step; written
                           only by tuning builds (-DBXG_PHASE_TIMERS, tools/build_alt.py); may be NULL */
} BxgDiag;
#define BXG_NUM_PHASES 16

/* ---- environment step ---------------------------------
 * The env arithmetic the reference does around pipeline_step, fused after the
 * last substep while the state is still in shared memory:
 *   envs/ant.py:233-279, envs/humanoid.py:256-354   obs / reward / done / metrics
 *   envs/wrappers/training.py:75-158                 EpisodeWrapper + AutoResetWrapper
 */
enum {
  BXG_ENV_ROOT_VELOCITY = 1,  /* Ant:      velocity of link 0, obs = [q[skip:], qd]          */
  BXG_ENV_PLANAR = 3,         /* Hopper / Walker2d: velocity of link 0; healthy = z, angle q[2] and state ranges
                                 (strict); obs = [q[skip:] with q[1] := z of link 0, clip(qd, -10, 10)]          */
  BXG_ENV_COM_VELOCITY = 2,   /* Humanoid: velocity of the centre of mass, obs = [q[skip:],
                                 qd, com_inertia, com_velocity, qfrc_actuator]; the action is
                                 rescaled from [-1,1] to the ctrl range first                */
  BXG_ENV_CARTPOLE = 4,       /* InvertedPendulum (envs/inverted_pendulum.py:131-154): action rescaled to the ctrl
                                 range; obs = [q, qd]; reward = 1; done = |q[1]| > healthy_angle_max */
  BXG_ENV_DOUBLE_CARTPOLE = 5,/* InvertedDoublePendulum (envs/inverted_double_pendulum.py:161-195): tip = tip_pos in
                                 the frame of link tip_link; obs = [q[0], sin(q[1:]), cos(q[1:]), clip(qd, -10, 10)];
                                 done = tip z <= healthy_z_min; reward = (1 - done) * healthy_reward - penalties */
  BXG_ENV_REACHER = 6,        /* Reacher (envs/reacher.py:199-239): obs = [cos(q[:2]), sin(q[:2]), q[2:], tip_vel[:2],
                                 tip - target]; tip on link tip_link, target = position of link target_link;
                                 reward = -|tip - target| - sum(action^2) */
  BXG_ENV_SWIMMER = 7,        /* Swimmer (envs/swimmer.py:157-194): velocity of q[:2]; obs = [q[skip:], qd];
                                 reward = forward_reward_weight * vx - ctrl_cost_weight * sum(action^2) */
  BXG_ENV_PUSHER = 9,         /* Pusher (envs/pusher.py:195-237): action rescaled to the ctrl range; reward from the
                                 PRE-step centres of mass of tip_link, object_link, target_link:
                                 -|obj - goal| - 0.1 * sum(action^2) - 0.5 * |obj - tip|;
                                 obs = [q[:nu], qd[:nu], com(tip), com(object), com(goal)] of the post-step state */
  BXG_ENV_STANDUP = 8         /* HumanoidStandup (envs/humanoidstandup.py:220-274): action rescaled to the ctrl range;
                                 obs as COM_VELOCITY; reward = z of link 0 / env_dt + healthy_reward
                                 - ctrl_cost_weight * sum(action^2); never done */
};
#define BXG_ENV_NUM_METRICS 10

typedef struct BxgEnvSpec {
  int32_t kind;                      /* BXG_ENV_* */
  int32_t obs_skip;                  /* 2 when exclude_current_positions_from_observation */
  int32_t terminate_when_unhealthy;
  int32_t episode_length;            /* EpisodeWrapper; <= 0 disables truncation */
  float forward_reward_weight;       /* 1.0 for Ant */
  float ctrl_cost_weight;
  float healthy_reward;
  float healthy_z_min, healthy_z_max;
  float env_dt;                      /* sys.opt.timestep * n_frames */
  float healthy_angle_min, healthy_angle_max;   /* BXG_ENV_PLANAR: range of q[2] */
  float healthy_state_min, healthy_state_max;   /* BXG_ENV_PLANAR: range of every entry of [q[2:], qd] */
  int32_t tip_link, target_link;     /* BXG_ENV_DOUBLE_CARTPOLE / BXG_ENV_REACHER */
  float tip_pos[3];                  /* tip in the frame of tip_link */
  int32_t object_link;               /* BXG_ENV_PUSHER */
} BxgEnvSpec;

/* Per-env arrays, device pointers.  metrics order:
 *  ROOT_VELOCITY: reward_forward, reward_survive, reward_ctrl, reward_contact, x_position,
 *                 y_position, distance_from_origin, x_velocity, y_velocity, forward_reward
 *  PLANAR:        reward_forward, reward_healthy, reward_ctrl, -, x_position, -, -, x_velocity, -, -
 *  COM_VELOCITY:  forward_reward, reward_linvel, reward_quadctrl, reward_alive, x_position,
 *                 y_position, distance_from_origin, x_velocity, y_velocity, (unused)
 *  CARTPOLE, DOUBLE_CARTPOLE: none
 *  REACHER:       reward_dist, reward_ctrl
 *  STANDUP:       reward_linup, reward_quadctrl
 *  PUSHER:        reward_dist, reward_ctrl, reward_near
 *  SWIMMER:       reward_fwd, -, reward_ctrl, -, x_position, y_position, distance_from_origin,
 *                 x_velocity, y_velocity, forward_reward (always 0: swimmer.py never updates it)  */
typedef struct BxgEnvIO {
  float* obs;                 /* [n, obs_size] out */
  float* reward;              /* [n] out */
  float* done;                /* [n] in (previous done, resets `steps`) / out */
  float* metrics;             /* [n, BXG_ENV_NUM_METRICS] out */
  float* steps;               /* [n] in/out  info['steps'];      may be NULL */
  float* truncation;          /* [n] out     info['truncation']; may be NULL */
  const BxgState* first_state; /* AutoResetWrapper source; NULL disables auto-reset */
  const float* first_obs;      /* [n, obs_size] */
  int32_t flags;               /* BXG_STEP_LEAN or 0 */
} BxgEnvIO;

typedef struct BxgModel BxgModel;

int bxg_abi_version(void);
const char* bxg_last_error(void);

/* Uploads model constants to `device` (cuda ordinal).  Replaces the implicit
 * capture of `sys` by jit in the reference (envs/base.py:126,133). */
int bxg_model_create(const BxgModelDesc* desc, int device, BxgModel** out);
void bxg_model_destroy(BxgModel* model);
/* nc for this model (rows of con_jac). */
int bxg_model_num_constraints(const BxgModel* model);
/* The kernel id this model's launches use (bxg_plan info[7]). */
int bxg_model_kernel_id(const BxgModel* model);
/* A batched System: n models of ONE topology whose constants differ, env e of every launch uses descs[e].  Replaces
 * `jax.vmap(step, in_axes=[sys_in_axes, 0, 0])(sys_v, state, action)` of the reference's DomainRandomizationVmapWrapper
 * (envs/wrappers/training.py:223-260).  What may differ: every float array of BxgModelDesc (link / inertia / dof /
 * actuator / contact constants, solver parameters), viscosity and density.  What must be equal (BXG_E_UNSUPPORTED
 * otherwise): sizes, link_parent / link_ndof, actuator and contact index arrays, dt, gravity, iteration counts.
 * Launches on the returned model require n_env == n (BXG_E_INVALID otherwise), run the reference's Newton-Schulz mode
 * only and read each env's constants from global memory (kernel variants 0, 1 and 3). */
int bxg_model_create_batched(const BxgModelDesc* descs, int64_t n, int device, BxgModel** out);
/* n of bxg_model_create_batched, 0 for a model every env shares. */
int64_t bxg_model_num_models(const BxgModel* model);
/* Host-only planning query (no CUDA needed): which kernel variant a model maps
 * to and its shared-memory footprint.  info[0] = variant id, [1] = lanes per env,
 * [2] = model words, [3] = per-env slab words, [4] = envs per CTA,
 * [5] = dynamic shared memory bytes per CTA, [6] = nc, [7] = kernel id: the variant id, or 10 / 11 when the
 * model's packed layout is exactly the one a model-specialised build of that variant was compiled for
 * (Ant on variant 0, Humanoid on variant 1: sizes and offsets are compile-time constants there). */
int bxg_plan(const BxgModelDesc* desc, int32_t info[8]);
/* The launch a batch of n_env envs gets on this model's device (the persistent kernel picks
 * the number of envs per CTA per call: small or awkward batches use smaller CTAs).
 * info[0] = grid (CTAs), [1] = threads per CTA, [2] = dynamic shared memory bytes per CTA,
 * [3] = envs per CTA. */
int bxg_launch_shape(const BxgModel* model, int64_t n_env, int32_t info[4]);

/* generalized.pipeline.init over a batch (pipeline.py:32-61):
 * q [n,nq], qd [n,nv] -> every leaf of `out`. */
int bxg_init(const BxgModel* model, int64_t n_env, const float* q,
             const float* qd, const BxgState* out, void* stream);

/* n_frames x generalized.pipeline.step with one action per env
 * (pipeline.py:64-94 looped as envs/base.py:128-137).  `in` and `out` may
 * alias field by field (in-place update).  act [n,nu] (NULL iff nu == 0). */
int bxg_step(const BxgModel* model, int64_t n_env, int32_t n_frames,
             const BxgState* in, const float* act, const BxgState* out,
             int32_t flags, const BxgDiag* diag, void* stream);

/* One wrapped env.step for every env: AutoReset(Episode(env)).step(state, action)
 * with action_repeat = 1 (envs/wrappers/training.py:98-158 around envs/ant.py:233 /
 * envs/humanoid.py:256).  `action` is the raw policy action [n, nu]. */
int bxg_env_step(const BxgModel* model, const BxgEnvSpec* spec, int64_t n_env,
                 int32_t n_frames, const BxgState* in, const float* action,
                 const BxgState* out, const BxgEnvIO* io, void* stream);
/* Length of one observation vector for this model / env kind. */
int bxg_env_obs_size(const BxgModel* model, const BxgEnvSpec* spec);
/* Env.reset (envs/ant.py:205-231, envs/humanoid.py:227-254) without the RNG:
 * pipeline.init(q, qd) plus the observation of the fresh state (zero action). */
int bxg_env_reset(const BxgModel* model, const BxgEnvSpec* spec, int64_t n_env,
                  const float* q, const float* qd, const BxgState* out, float* obs,
                  void* stream);

/* ---- the two launch-bound pieces of the PPO loop -------------------------------
 * compute_gae (agents/ppo/losses.py:38-101) on time-major [T, B] arrays; bootstrap [B]; writes vs and advantages. */
int bxg_gae(const float* truncation, const float* termination, const float* reward, const float* values,
            const float* bootstrap, int32_t T, int64_t B, float lambda, float discount, float* vs, float* advantages,
            void* stream);
/* The policy's inference step of the rollout (training/acting.py:33-53 with agents/ppo/networks.py:66-88):
 * normalise obs with (mean, std) (acme/running_statistics.py:303-328; clip <= 0: no clipping), MLP
 * [obs, h1, h2, 2 act] with swish (torch Linear layout: W [out, in]), loc / scale = softplus + min_std, raw action
 * = loc + scale * noise, action = tanh(raw).  Compiled for h1 = h2 = 64; other widths return BXG_E_UNSUPPORTED. */
int bxg_policy_act(const float* obs, const float* mean, const float* std, float clip, const float* W1, const float* b1,
                   const float* W2, const float* b2, const float* W3, const float* b3, const float* noise, int64_t n,
                   int32_t obs_size, int32_t h1, int32_t h2, int32_t act_size, float min_std, float* logits, float* pre,
                   float* action, void* stream);

/* compute_ppo_loss from the network outputs on (agents/ppo/losses.py:186-303): termination, reward scaling,
 * compute_gae, advantage normalisation, tanh-normal log-probs, clipped surrogate, value loss (x 0.5 x vf 0.5), sampled
 * entropy -- and the gradients of the total loss wrt the policy logits [T, B, 2 act] and the values [T + 1, B]
 * (row T is the bootstrap value: zero gradient), in two launches.  loss4 = (total, policy, value, entropy). */
int bxg_ppo_head(const float* logits, const float* values, const float* behaviour_logits, const float* raw_action,
                 const float* reward, const float* done, const float* truncation, const float* entropy_noise, int32_t T, int64_t B,
                 int32_t act_size, float reward_scaling, float lambda, float discount, float clip_epsilon, float entropy_cost,
                 int32_t normalize_advantage, float min_std, float* vs_scratch, float* adv_scratch, double* stats_scratch,
                 float* loss4, float* dlogits, float* dvalues, void* stream);

/* Number of kernel launches issued by this library so far in this process
 * (bench.py reports the delta over the timed region as gpu_launches). */
int64_t bxg_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* BXG_H_ */
