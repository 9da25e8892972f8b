/* mmz.h - C ABI of libmmz.so, the B200-native batched maze step engine.
 *
 * This is the drop-in boundary for the reference's hot path
 *   MazeEnv.step           mujoco_maze/maze_env.py:448-481
 *   MazeEnv.reset          mujoco_maze/maze_env.py:371-382
 *   MazeEnv._get_obs       mujoco_maze/maze_env.py:351-369
 * and everything those call per step (AgentModel.step -> MuJoCo mj_step,
 * CollisionDetector.detect, MazeTask.reward / termination): one call advances
 * N lock-step environments. The reference reaches its physics through the
 * Cython FFI of mujoco-py (`MujocoEnv.do_simulation / set_state`, call sites
 * ant.py:63,95,108, point.py:57-59,80,89, swimmer.py:39,67,73); this header is
 * what a binding for the batched path binds instead (INTEGRATION.md shows the
 * ctypes stub).
 *
 * Conventions
 *  - plain C, no exceptions cross the boundary: every function returns 0
 *    (MMZ_OK) or a negative MMZ_ERR_* code; mmz_last_error() gives the text
 *    (thread-local).
 *  - every `d_*` pointer is CALLER-OWNED DEVICE memory on the handle's device
 *    and must stay valid until the work enqueued on `stream` has completed.
 *    `h_*` pointers are host memory (pinned for async overlap).
 *  - calls only ENQUEUE on `stream` (a cudaStream_t passed as void*, NULL =
 *    default stream); nothing synchronises except mmz_step_host and
 *    mmz_destroy.
 *  - one handle is driven by one host thread at a time; handles on different
 *    devices are independent (multi-GPU = one handle per rank, no collective).
 *  - arrays are env-major: action [N][nu], obs [N][obs_dim], reward [N],
 *    done [N], info [N][4]. State transfers take a layout flag.
 */
#ifndef MMZ_H
#define MMZ_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mmz_env* mmz_handle;

#define MMZ_OK 0
#define MMZ_ERR_INVALID (-1)  /* bad argument                                 */
#define MMZ_ERR_MODEL (-2)    /* blob failed validation (magic/version/size)  */
#define MMZ_ERR_CUDA (-3)     /* CUDA runtime error, see mmz_last_error()     */
#define MMZ_ERR_CAPACITY (-4) /* model exceeds compiled kernel capacities     */

/* bits of done[i] */
#define MMZ_DONE 1u      /* terminated (MazeTask.termination) or truncated      */
#define MMZ_TRUNCATED 2u /* t >= max_episode_steps (gym TimeLimit, __init__.py:31) */
#define MMZ_UNSTABLE 4u  /* non-finite state detected; env was re-initialised
                            (MuJoCo's mj_checkPos/Vel/Acc auto-reset [EXT])     */

/* create flags */
#define MMZ_AUTO_RESET 1u /* re-initialise an env inside the step that ends it */

/* state layouts for mmz_get_state / mmz_set_state */
#define MMZ_LAYOUT_ENV_MAJOR 0 /* qpos [N][nq], qvel [N][nv] */
#define MMZ_LAYOUT_SOA 1       /* qpos [nq][N], qvel [nv][N] */

/* Build N environments from a model blob (include/mmz_model.h, float layout)
 * on CUDA device `device`. Replaces MazeEnv.__init__ -> model_cls(file_path)
 * (maze_env.py:218) -> MujocoEnv.__init__ [EXT]. */
int mmz_create(const void* model_blob, size_t bytes, int num_envs, int device, uint32_t flags, mmz_handle* out);

/* Sizes implied by the model. Any out pointer may be NULL. */
int mmz_dims(mmz_handle h, int* num_envs, int* nq, int* nv, int* nu, int* obs_dim);

/* How the step kernel was configured for this model on this device: lanes of a warp that
 * cooperate on one environment (8, 16 or 32 = one warp per environment), threads per block,
 * dynamic shared memory per block, resident environments per SM and the per-environment
 * shared-memory workspace in floats. Any out pointer may be NULL. */
int mmz_kernel_config(mmz_handle h, int* lanes_per_env, int* threads_per_block, int* smem_bytes, int* envs_per_sm,
                      int* floats_per_env);

/* Name of the step kernel this handle launches: "maze_hkernel<14>" (hybrid: tree phases lane = environment,
 * solver 16 lanes per environment) or "maze_kernel<G,NVP,FEAT>" (G lanes per environment). Static string. */
const char* mmz_kernel_name(mmz_handle h);

/* Multi-GPU sharding: this handle holds environments [first_global_env, first_global_env + N)
 * of a larger batch. Only the reset noise depends on it (Philox streams are keyed by the GLOBAL
 * environment index), so 1 GPU x N and G GPUs x N/G give identical results. There is no
 * reference counterpart: the reference holds one mjData per Python object (maze_env.py:218). */
int mmz_set_env_offset(mmz_handle h, int first_global_env);

/* Fused observation gather - BASELINE configs[3], "NCCL over NVLink used only to all-gather observations when the
 * caller wants a single tensor": every following mmz_step ALSO stores the observation row of environment e at row
 * `row_offset + e` of each of the `npeers` buffers d_peer_obs[k] (row-major [total_envs][obs_dim] float32: this rank's
 * gathered tensor and the peer-mapped gathered tensors of the other ranks, e.g. CUDA IPC / symmetric memory), from
 * inside the step kernel, so no collective runs after it. With `multicast` non-zero d_peer_obs[0] is ONE multicast
 * (NVLS) address that fans out to every rank (multimem.st). npeers = 0 turns the gather off. The caller orders the
 * ranks (a barrier after the step, before anyone reads the gathered tensors or starts the next step). No reference
 * counterpart: the reference holds one mjData per Python object (maze_env.py:218). Not available for tasks with
 * TOP_DOWN_VIEW (MMZ_ERR_INVALID). */
int mmz_set_obs_peers(mmz_handle h, float* const* d_peer_obs, int npeers, int64_t row_offset, int multicast);

/* MazeEnv.reset (maze_env.py:371-382) + reset_model (point.py:71-81,
 * ant.py:84-96, swimmer.py:55-68) for the envs whose d_mask byte is non-zero
 * (NULL = all). Noise is Philox4x32-10 keyed by (seed, env index): the
 * distributions match the reference, the bit stream does not. Writes the
 * first observation of the reset envs into d_obs if non-NULL. */
int mmz_reset(mmz_handle h, const uint8_t* d_mask, uint64_t seed, float* d_obs, void* stream);

/* MazeEnv.step (maze_env.py:448-481) for all N envs: one fused kernel launch.
 * d_info (optional) receives [x, y, reward_forward, reward_ctrl] per env:
 * info["position"] (maze_env.py:480) and ant.py:72 / swimmer.py:46. */
int mmz_step(mmz_handle h, const float* d_action, float* d_obs, float* d_reward, uint8_t* d_done, float* d_info,
             void* stream);

/* K consecutive MazeEnv.step calls in ONE host call (SURVEY section 7 "steps per launch"): d_actions is [K][N][nu], the
 * outputs are [K][N][obs_dim] / [K][N] / [K][N] / [K][N][4] (d_info optional), step k reading / writing slice k. Same
 * results, bit for bit, as K mmz_step calls, including TimeLimit truncation and in-kernel auto-reset in the middle of
 * the block. The K launches are recorded once into a CUDA graph per (K, buffers) and replayed: one driver call per
 * rollout segment - what the small robots need, whose step is ~0.2 ms (reference loop: point.py:44-61 called once per
 * Python iteration). */
int mmz_step_k(mmz_handle h, int K, const float* d_actions, float* d_obs, float* d_reward, uint8_t* d_done, float* d_info,
               void* stream);

/* Same step through HOST buffers, then a stream synchronise. This is the
 * end-to-end call a host-resident caller (the reference's numpy world) makes.
 * PINNED buffers (cudaHostAlloc / cudaHostRegister / torch pin_memory) are
 * mapped into the device's address space: the step kernel reads the actions
 * and writes obs/reward/done/info straight over PCIe, block by block, under
 * the physics of the blocks still running - one launch, no staging copies.
 * Pageable buffers (or MMZ_HOST_ZERO_COPY=0) take the staged path: H2D of the
 * actions, the kernel, D2H of the results, batches of 64 blocks or more as 4
 * block ranges on internal streams so that the copies of one range overlap the
 * kernel of another. Results are identical to mmz_step either way. */
int mmz_step_host(mmz_handle h, const float* h_action, float* h_obs, float* h_reward, uint8_t* h_done, float* h_info,
                  void* stream);

/* MazeEnv._get_obs (maze_env.py:351-369) of the current state. */
int mmz_observe(mmz_handle h, float* d_obs, void* stream);

/* sim.get_state / MujocoEnv.set_state [EXT] (call sites point.py:57,80,89;
 * ant.py:95,108). d_t is the per-env step counter MazeEnv.t (maze_env.py:45). */
int mmz_get_state(mmz_handle h, int layout, float* d_qpos, float* d_qvel, int32_t* d_t, void* stream);
int mmz_set_state(mmz_handle h, int layout, const float* d_qpos, const float* d_qvel, const int32_t* d_t,
                  void* stream);

/* One forward-dynamics evaluation (mj_forward [EXT]) at the current state
 * under d_action, without advancing: qacc [N][nv] and per-env diagnostics
 * diag [N][4] = {contacts, constraint rows, solver iterations, overflow}.
 * Test / debugging aid for parity against the oracle. */
int mmz_forward(mmz_handle h, const float* d_action, float* d_qacc, int32_t* d_diag, void* stream);

/* Optional solver diagnostics of mmz_step: when d_diag is non-NULL every following step writes
 * diag [N][4] = {Newton iterations, line-search iterations, max simultaneous contacts, solves
 * that hit the iteration cap}, summed over the forward evaluations of that step. NULL disables. */
int mmz_set_step_diag(mmz_handle h, int32_t* d_diag);

/* MazeEnv.render(mode="rgb_array") (maze_env.py:389-420) for environments [first_env, first_env + count): an
 * orthographic top-down RGB image of each (floor / platforms / chasms, maze boxes, goal sites, agent geoms, movable
 * blocks, object balls), d_rgb [count][height][width][3] uint8, row 0 = largest y. The window is the maze's bounding
 * box plus half a cell. One launch; the reference reads one OpenGL frame per call. No pixel parity is claimed. */
int mmz_render(mmz_handle h, int first_env, int count, int width, int height, uint8_t* d_rgb, void* stream);

/* Number of CUDA kernels this handle has launched so far. */
uint64_t mmz_launch_count(mmz_handle h);

/* Text of the last error on the calling thread ("" if none). */
const char* mmz_last_error(void);

/* Waits for the handle's outstanding work and frees it. */
void mmz_destroy(mmz_handle h);

/* ABI version of this library (bumped on any signature change). */
int mmz_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* MMZ_H */
