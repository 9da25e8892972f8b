/* mmz_model.h - binary layout of the maze model blob.
 *
 * GENERATED from mujoco-maze_b200/mujoco_maze/model_layout.py - do not edit.
 * The blob is produced on the host by the model compiler (the stand-in for
 * MuJoCo's MJCF compiler + MazeEnv.__init__ geometry injection,
 * reference maze_env.py:97-218) and consumed by mmz_create() (include/mmz.h)
 * and by the CPU oracle. Define MMZ_REAL_IS_DOUBLE for the oracle's layout.
 */
#ifndef MMZ_MODEL_H
#define MMZ_MODEL_H
#include <stdint.h>

#ifdef MMZ_REAL_IS_DOUBLE
typedef double mmz_real;
#else
typedef float mmz_real;
#endif

#define MMZ_MAGIC 0x4D4D5A31
#define MMZ_VERSION 5
#define MMZ_MAXBODY 16
#define MMZ_MAXJNT 20
#define MMZ_MAXDOF 20
#define MMZ_MAXQ 24
#define MMZ_MAXGEOM 20
#define MMZ_MAXACT 8
#define MMZ_MAXGOAL 4
#define MMZ_MAXSEG 64
#define MMZ_MAXCELL 144
#define MMZ_MAXOBJ 8

#define MMZ_JNT_FREE 0
#define MMZ_JNT_BALL 1
#define MMZ_JNT_SLIDE 2
#define MMZ_JNT_HINGE 3
#define MMZ_GEOM_PLANE 0
#define MMZ_GEOM_SPHERE 2
#define MMZ_GEOM_CAPSULE 3
#define MMZ_GEOM_BOX 6
#define MMZ_STEP_TORQUE 0
#define MMZ_STEP_TELEPORT 1
#define MMZ_RESET_POINT 0
#define MMZ_RESET_ANT 1
#define MMZ_RESET_SWIMMER 2
#define MMZ_CELL_WALL 1
#define MMZ_CELL_PLATFORM 2
#define MMZ_CELL_CHASM 4
#define MMZ_VIEW_DIM 75
/* resolved reward / termination rules */
#define MMZ_REWARD_REACH 0         /* 1.0 if terminated else penalty        maze_task.py:110-111 */
#define MMZ_REWARD_SCALED 1        /* first reached goal's reward_scale     maze_task.py:356-360 */
#define MMZ_REWARD_SCALED_OBJECT 2 /* same on obs[3:6]                      maze_task.py:592-597 */
#define MMZ_REWARD_DIST_OBJECT 3   /* -|obs[3:6]-goal0|/scale               maze_task.py:619-621 */
#define MMZ_REWARD_ZERO 4          /* NoReward*                                                  */
#define MMZ_REWARD_DIST 5          /* -|obs[:dim]-goal0|/scale              maze_task.py:98-99   */
#define MMZ_REWARD_HOST 6          /* user-defined: outer reward left to the host wrapper       */
#define MMZ_TERM_AGENT 0           /* any goal within threshold of obs[:dim] maze_task.py:77-81  */
#define MMZ_TERM_OBJECT 1          /* ... of obs[3:6]                       maze_task.py:599-604 */
#define MMZ_TERM_HOST 2
/* forward_reward_fn of AntEnv / SwimmerEnv (reference ant.py:18-23) on the xy velocity of the step */
#define MMZ_FWD_VNORM 0            /* forward_reward_vnorm: |v|, the default                     */
#define MMZ_FWD_VABS 1             /* forward_reward_vabs: |vx| + |vy|                           */
#define MMZ_FWD_HOST 2             /* any other callable: the host wrapper adds weight * fn(v)   */

typedef struct mmz_model {
  int32_t magic;
  int32_t version;
  int32_t real_bytes; /* 4 or 8 */
  int32_t total_bytes; /* sizeof(mmz_model) */
  int32_t nbody; /* moving bodies */
  int32_t njnt;
  int32_t nv;
  int32_t nq;
  int32_t ngeom; /* geoms on moving bodies */
  int32_t nu;
  int32_t ngoal;
  int32_t nseg;
  int32_t grid_h;
  int32_t grid_w;
  int32_t step_kind; /* MMZ_STEP_* */
  int32_t frame_skip;
  int32_t manual_collision; /* segment clamp on the agent xy (point.py:30) */
  int32_t collision_on; /* 0: option collision=predefined with no pairs (swimmer.xml:3) */
  int32_t has_floor;
  int32_t elevated;
  int32_t reward_rule; /* MMZ_REWARD_* (resolved, survey A9) */
  int32_t term_rule; /* MMZ_TERM_* */
  int32_t max_episode_steps; /* TimeLimit (__init__.py:31) */
  int32_t obs_dim;
  int32_t n_agent_q; /* agent qpos entries copied to obs */
  int32_t n_agent_v; /* agent qvel entries copied to obs */
  int32_t nobj; /* observed bodies spliced after obs[:3] */
  int32_t reset_kind; /* MMZ_RESET_* */
  int32_t forward_reward_kind; /* MMZ_FWD_*: forward_reward_fn of the torque agents (ant.py:18-23,44-53) */
  int32_t nviewb; /* bodies latched for the top-down view after the observed ones: torso, then movable blocks */
  int32_t view_dim; /* 0, or 75 = 5x5x3 top-down view between the state part of obs and t (maze_env.py:353-369) */
  int32_t obj_body[8]; /* nobj observed bodies, then nviewb view bodies */
  int32_t body_parent[16]; /* -1 = world */
  int32_t body_jntadr[16];
  int32_t body_jntnum[16];
  int32_t body_dofadr[16];
  int32_t body_dofnum[16];
  int32_t body_level[16]; /* depth in its tree, roots = 0 */
  int32_t body_root[16]; /* root body of its tree */
  int32_t body_dofmask[16]; /* bit d set: dof d moves this body */
  int32_t jnt_type[20];
  int32_t jnt_body[20];
  int32_t jnt_qadr[20];
  int32_t jnt_dadr[20];
  int32_t jnt_limited[20];
  int32_t dof_body[20];
  int32_t dof_jnt[20];
  int32_t dof_parent[20]; /* -1 = none */
  int32_t geom_type[20];
  int32_t geom_body[20];
  int32_t geom_contype[20];
  int32_t geom_conaffinity[20];
  int32_t geom_condim[20];
  int32_t act_dof[8];
  int32_t act_limited[8];
  int32_t goal_dim[4];
  int32_t grid[144]; /* row-major, bit0 wall box (BLOCK cell), bit1 platform box, bit2 CHASM cell */
  int32_t pad_;
  mmz_real timestep;
  mmz_real gravity[3];
  mmz_real density; /* fluid */
  mmz_real viscosity; /* fluid */
  mmz_real inner_reward_scale; /* maze_env.py:477 */
  mmz_real forward_reward_weight; /* ant.py:47 */
  mmz_real ctrl_cost_weight; /* ant.py:48 */
  mmz_real restitution; /* maze_env.py:36 */
  mmz_real penalty;
  mmz_real task_scale; /* MazeTask.scale */
  mmz_real vel_limit; /* point.py:33 */
  mmz_real reset_noise; /* 0.1 */
  mmz_real cell_size;
  mmz_real origin[2]; /* robot cell centre (torso_x, torso_y) */
  mmz_real wall_half[3]; /* half extents of a wall box */
  mmz_real wall_z; /* centre z of wall boxes */
  mmz_real plat_z; /* centre z of platform boxes */
  mmz_real wall_margin;
  mmz_real wall_friction[3];
  mmz_real wall_solref[2];
  mmz_real wall_solimp[5];
  mmz_real floor_z;
  mmz_real floor_margin;
  mmz_real floor_friction[3];
  mmz_real floor_solref[2];
  mmz_real floor_solimp[5];
  mmz_real body_pos[16][3];
  mmz_real body_quat[16][4];
  mmz_real body_ipos[16][3];
  mmz_real body_iquat[16][4];
  mmz_real body_mass[16];
  mmz_real body_inertia[16][3];
  mmz_real jnt_pos[20][3];
  mmz_real jnt_axis[20][3];
  mmz_real jnt_range[20][2];
  mmz_real jnt_margin[20];
  mmz_real jnt_solref[20][2];
  mmz_real jnt_solimp[20][5];
  mmz_real qpos0[24];
  mmz_real dof_armature[20];
  mmz_real dof_damping[20];
  mmz_real dof_invweight0[20];
  mmz_real geom_size[20][3];
  mmz_real geom_pos[20][3];
  mmz_real geom_quat[20][4];
  mmz_real geom_margin[20];
  mmz_real geom_friction[20][3];
  mmz_real geom_solref[20][2];
  mmz_real geom_solimp[20][5];
  mmz_real geom_invweight[20]; /* translational body_invweight0 of the geom's (unmerged) body */
  mmz_real act_gear[8];
  mmz_real act_ctrlrange[8][2];
  mmz_real goal_pos[4][3];
  mmz_real goal_thr[4];
  mmz_real goal_scale[4];
  mmz_real seg[64][4]; /* x1 y1 x2 y2 */
} mmz_model;

#define MMZ_MODEL_NINT 592
#define MMZ_MODEL_NREAL 1477

#endif /* MMZ_MODEL_H */
