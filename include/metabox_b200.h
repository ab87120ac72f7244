/* metabox_b200.h -- C ABI of the B200-native MetaBBO rollout engine (libmbx.so).
 *
 * Drop-in boundary for the per-generation hot path of GMC-DRL/MetaBox (reference paths are relative
 * to /root/reference/src).  The reference has no FFI: its boundary is the duck-typed Python plug-in
 * API (docs/Customization.md:9-210).  Every entry point below names the reference method it
 * replaces; the Python shims in metabox_b200/ keep those methods' names and meaning and bind this
 * library through ctypes.
 *
 * Conventions
 *   - plain C: integers, doubles, raw pointers.  No torch / C++ types.
 *   - every `d_*` pointer is a DEVICE pointer owned by the caller; `h_*` pointers are HOST pointers.
 *   - every launch goes to the caller's `stream` (a cudaStream_t passed as void*); entry points never
 *     allocate device memory, never synchronise, and return 0 or a negative MBX_E* code;
 *     mbx_last_error() gives the message of the last failure on the calling thread.
 *   - "lane" = one (instance, run) population advanced in lockstep with the others; B = #lanes.
 *   - population state is float32 in HBM, arithmetic is float64 in registers.
 */
#ifndef METABOX_B200_H
#define METABOX_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MBX_OK 0
#define MBX_EINVAL -1   /* bad argument (shape, id, null pointer)            */
#define MBX_ECUDA -2    /* CUDA runtime error (launch / attribute)           */
#define MBX_EUNSUP -3   /* (kind, dim) not compiled: the library carries BBOB dims 10, 20, 30 (+ one -DMBX_DIM_EXTRA=<even d <= 32> at build
                           time, csrc/mbx_launch.h) and protein dim 12; the reference's `--dim` is free (src/config.py:10) */

#define MBX_MAX_DIM 64
#define MBX_MAX_NP 128
#define MBX_CURVE_LEN 51          /* n_logpoint + 1, tester.py:204-205 pads to 51 */

/* problem kinds */
#define MBX_KIND_BBOB 0
#define MBX_KIND_PROTEIN 1
/* noise kinds (problem/bbob.py:108-146) */
#define MBX_NOISE_NONE 0
#define MBX_NOISE_GAUSS 1
#define MBX_NOISE_UNIFORM 2
#define MBX_NOISE_CAUCHY 3

/* One problem instance.  Built on the host by metabox_b200/problem.py (`pack_host`, `BBOB_Problem.header` / `const_block`) from the instance constants that
 * problem/bbob.py:31-45 + the per-class constructors (bbob.py:149-890) or
 * problem/protein_docking.py:163-181 produce; `off` is the instance's offset (in doubles) into the
 * constants buffer.  BBOB constants layout at `off`:
 *     shift[D] | R[D*D] | M2[D*D] | cv0[D] | cv1[D] | cv2[D] | (Gallagher) y[P*D] | C[P*D] | w[P] | Ry[P*D] (= R y_p)
 * protein layout at `off`:
 *     coor_init[300] | basis[12*300] | inv_sqrt_eig[12] | q[100] | e4[100] (= e_i^(1/4)) | r[100] |
 *     candidate pair list: n_peaks x uint16 (i << 8 | j), the atom pairs that can come within the 9 A cut-off   */
typedef struct MbxInstance {
    int32_t kind;          /* MBX_KIND_*                                         */
    int32_t base;          /* BBOB base shape 1..24 (noisy ids map to their base) */
    int32_t dim;
    int32_t noise;         /* MBX_NOISE_*                                        */
    int32_t n_peaks;       /* Gallagher: 101 / 21; protein: # candidate atom pairs; else 0 */
    int32_t has_optimum;   /* 0 for protein (`optimum is None`)                  */
    int32_t fid;           /* reference function id (1..24, 101..130), 0 protein */
    int32_t pad_;
    double lb, ub, bias, optimum;
    double bh_coef;        /* boundaryHandling(x) = bh_coef * pen(x)             */
    double n_alpha, n_beta, n_p;   /* noise parameters                           */
    double sc[8];          /* function-specific scalars (problem.py header())    */
    int64_t off;
} MbxInstance;

/* Scalars of one optimizer configuration (config.py:74-102 + optimizer constructors). */
typedef struct MbxRunConfig {
    int32_t kind;          /* MBX_KIND_* of every lane in the launch             */
    int32_t NP;            /* population size                                    */
    int32_t dim;
    int32_t maxFEs;
    int32_t n_logpoint;    /* 50 (bbob) / 5 (protein)                            */
    int32_t log_interval;  /* maxFEs // n_logpoint                               */
    int32_t use_tape;      /* 1: random numbers are read from MbxTape (parity tests); 0: Philox */
    int32_t lane_base;     /* Philox lane id of lane 0 (global lane index when a sweep is sharded) */
    uint64_t seed;         /* Philox key                                         */
} MbxRunConfig;

const char* mbx_last_error(void);
int mbx_version(void);
/* size in bytes of MbxInstance / MbxRunConfig as compiled (ABI check for the ctypes mirror) */
int mbx_sizeof_instance(void);
int mbx_sizeof_runconfig(void);

/* ---------------------------------------------------------------------------------------------
 * Problem.eval  -- replaces Basic_Problem.eval / func (problem/basic_problem.py:12-37) for the
 * BBOB classes (problem/bbob.py:149-890, noise mixins :96-146) and Protein_Docking.func
 * (problem/protein_docking.py:28-48).
 *   d_inst      [n_inst] MbxInstance,  d_consts  f64 constants buffer
 *   d_lane_inst [B] int32: instance index of each lane
 *   d_x         [B, NP, D] candidates, float32 (x_is_f64 = 0) or float64 (x_is_f64 = 1)
 *   d_out       [B, NP]  f(x) as float32 / float64 (same width as x); raw f(x) incl. bias, i.e. what
 *               `problem.eval(x)` returns;  if subtract_optimum != 0 returns eval(x) - optimum
 *   noise: when the instance is noisy, draws come from Philox(seed, lane, gen) unless d_noise_tape
 *          is non-null: [B, 3, NP] float64 in the reference draw order (uniforms, then normals).
 *   kind:  MBX_KIND_*, optionally OR-ed with ONE promise about the lanes of this launch.  BBOB lanes go through two kernels: a streaming
 *          kernel (one thread per candidate, no shared memory) for the functions without a rotation (base shapes 5, 8, 20 -- Linear slope,
 *          Rosenbrock, Schwefel, and the noisy ids built on Rosenbrock), a CTA-cooperative kernel for the rest; each skips the other's
 *          lanes.  MBX_EVAL_ALL_STREAMING / MBX_EVAL_NONE_STREAMING: every / no lane is of the first sort, the idle kernel is not
 *          launched.  A wrong promise leaves the mis-promised lanes' outputs unwritten.                                            */
#define MBX_EVAL_ALL_STREAMING 0x100
#define MBX_EVAL_NONE_STREAMING 0x200
int mbx_eval(int kind, const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst,
             const void* d_x, int x_is_f64, void* d_out, int B, int NP, int D, int subtract_optimum,
             uint64_t seed, uint32_t gen, const double* d_noise_tape, void* stream);

/* ---------------------------------------------------------------------------------------------
 * RLEPSO backbone (PSO family) -- replaces RLEPSO_Optimizer.init_population / update
 * (optimizer/rlepso_optimizer.py:39-65, 173-263) for B lanes at once.
 * State (all device, updated in place):
 *   x, v, pbest_x [B,NP,D] f32 ; c_cost, pbest [B,NP] f32 ; pni [B,NP] i32 (per_no_improve)
 *   gbest_x [B,D] f32 ; gbest_val [B] f32 ; gbest_idx, no_improve, fes, log_index, curve_len, gen [B] i32
 *   curve [B, MBX_CURVE_LEN] f32 ; done [B] u8                                                    */
typedef struct MbxRlepsoState {
    float *x, *v, *pbest_x, *c_cost, *pbest;
    int32_t* pni;
    float *gbest_x, *gbest_val;
    int32_t *gbest_idx, *no_improve, *fes, *log_index, *curve_len, *gen;
    float* curve;
    uint8_t* done;
} MbxRlepsoState;

/* Recorded random draws for parity tests (all device, float64 / int32, may be null when use_tape=0):
 *   init_pos_u, init_vel_u, clpso_u, fdr_u, reinit_pos_u, reinit_vel_u [B,NP,D] ; tour [B,NP,D,2] i32
 *   rand1, rand2, reinit_u [B,NP] ; noise [B,2,3,NP] (pass 0 = main evaluation, pass 1 = re-init)   */
typedef struct MbxRlepsoTape {
    const double *init_pos_u, *init_vel_u, *clpso_u, *fdr_u, *reinit_pos_u, *reinit_vel_u;
    const int32_t* tour;
    const double *rand1, *rand2, *reinit_u, *noise;
} MbxRlepsoTape;

int mbx_rlepso_init(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst,
                    const MbxRunConfig* cfg, const MbxRlepsoState* st, const MbxRlepsoTape* tape,
                    float* d_feature /* [B,1] fes/maxFEs */, int B, void* stream);

/* d_action [B,35] f32 in [0,1] ; outputs d_feature [B,1] f32, d_reward [B] f32 (+1/-1), st->done.
 * d_pci [NP] f64 = the CLPSO learning-probability table (rlepso_optimizer.py:24-25).
 * Lanes whose done flag is already set are skipped: state and feature untouched, reward = 0.
 * d_dbg (optional, may be null): [B, NP, D] int32 FDR target indices for index-parity tests.       */
int mbx_rlepso_step(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst,
                    const MbxRunConfig* cfg, const MbxRlepsoState* st, const MbxRlepsoTape* tape,
                    const double* d_pci, const float* d_action, float* d_feature, float* d_reward,
                    int32_t* d_dbg, int B, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LDE backbone (DE family, whole population per step) -- replaces LDE_Optimizer.init_population /
 * update / __get_feature (optimizer/lde_optimizer.py:132-140, 159-198, 145-157) for B lanes at once.
 * State (all device, updated in place); the population is kept sorted by cost (ascending, stable):
 *   pop [B,NP,D] f32 ; fit [B,NP] f32 ; hist_sum [B,5] f64 (column sums of past_histo) ; hist_cnt [B] i32
 *   gbest [B] f32 ; fes, log_index, curve_len, gen [B] i32 ; curve [B, MBX_CURVE_LEN] f32 ; done [B] u8   */
typedef struct MbxLdeState {
    float *pop, *fit;
    double* hist_sum;
    int32_t* hist_cnt;
    float* gbest;
    int32_t *fes, *log_index, *curve_len, *gen;
    float* curve;
    uint8_t* done;
} MbxLdeState;

/* Recorded draws for parity tests (device; null when use_tape=0):
 *   init_u, cross_u [B,NP,D] f64 ; pbest_idx, r1, r2, jrand [B,NP] i32 (r1/r2 = final donor indices after the
 *   reference's rejection loop, lde_optimizer.py:109-123) ; noise [B,3,NP] f64                            */
typedef struct MbxLdeTape {
    const double* init_u;
    const int32_t *pbest_idx, *r1, *r2;
    const double* cross_u;
    const int32_t* jrand;
    const double* noise;
} MbxLdeTape;

/* d_feature [B, NP+10] f32 = sorted min-max-normalised fitness | 5-bin histogram | mean past histogram */
int mbx_lde_init(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst,
                 const MbxRunConfig* cfg, const MbxLdeState* st, const MbxLdeTape* tape, float* d_feature, int B,
                 void* stream);

/* d_action [B, 2*NP] f32 (scale factors | crossover rates) ; outputs d_feature, d_reward [B] f32
 * ((bsf - bsf_next)/bsf), st->done.  Finished lanes are skipped (reward = 0).
 * d_dbg (optional): [B, NP*(D+1)] f32 = trial vectors [NP,D] then trial costs [NP] (index-parity tests). */
int mbx_lde_step(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst,
                 const MbxRunConfig* cfg, const MbxLdeState* st, const MbxLdeTape* tape, const float* d_action,
                 float* d_feature, float* d_reward, float* d_dbg, int B, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DE-DDQN backbone (DE family, ONE individual per step) -- replaces DE_DDQN_Optimizer.init_population /
 * update / __get_state (optimizer/de_ddqn_optimizer.py:46-74, 131-220, 76-129) for B lanes at once.
 * State (all device, updated in place):
 *   X [B,NP,D] f32 ; cost [B,NP] f32 ; X_gbest, X_prebest [B,D] f32 ; c3 [B,3] f32 = c_gbest, c_gworst, c_prebest
 *   ints [B,8] i32 = pointer, gen, stagcount, prebest_row (>=0 while X_prebest is the reference's *view* of that
 *                    row of X, else -1), gbest_rebound, window length, next window sequence number, error flag
 *   r [B,5] i32 (indices for the next mutation) ; n_tot [B,4,10] i32 ; n_succ [B,4,4,10] i32 ;
 *   om_sum, om_max [B,4,4,10] f64 (slot 0 = current generation; the reference's maxlen-10 deques)
 *   win [B,50,8] f64 rows = (action, om0..om3, trial_cost, sequence number, 0): the OM_W window, unordered slots
 *   fes, log_index, curve_len [B] i32 ; curve [B, MBX_CURVE_LEN] f32 ; done [B] u8                         */
typedef struct MbxDdqnState {
    float *X, *cost, *X_gbest, *X_prebest, *c3;
    int32_t *ints, *r, *n_tot, *n_succ;
    double *om_sum, *om_max, *win;
    int32_t *fes, *log_index, *curve_len;
    float* curve;
    uint8_t* done;
} MbxDdqnState;

/* Recorded draws for parity tests: init_u [B,NP,D] f64 ; r5 [B,5] i32 (this call's `randint(0,NP,5)`) ;
 * noise [B,3,n] f64 with n = NP for init and n = 1 for a step.                                           */
typedef struct MbxDdqnTape {
    const double* init_u;
    const int32_t* r5;
    const double* noise;
} MbxDdqnTape;

/* d_feature [B,99] f32 */
int mbx_deddqn_init(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst,
                    const MbxRunConfig* cfg, const MbxDdqnState* st, const MbxDdqnTape* tape, float* d_feature, int B,
                    void* stream);

/* d_action [B] i32 in 0..3 (rand/1, rand/2, rand-to-best/2, cur-to-rand/1) ; outputs d_feature [B,99] f32,
 * d_reward [B] f32 (max(parent - trial, 0)), st->done.  A lane given an action outside 0..3 is left untouched
 * and its ints[7] error flag is set (the reference raises ValueError('Action error')).
 * d_dbg (optional): [B, D+1] f32 = trial vector then trial cost.                                           */
int mbx_deddqn_step(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst,
                    const MbxRunConfig* cfg, const MbxDdqnState* st, const MbxDdqnTape* tape, const int32_t* d_action,
                    float* d_feature, float* d_reward, float* d_dbg, int B, void* stream);

/* Persistent DE-DDQN rollout segment: `n_steps` consecutive function evaluations of every unfinished lane in ONE launch --
 * the agent's greedy action (agent/de_ddqn_agent.py:59-68: argmax of the Q-network on the NaN/inf-sanitised 99 features, weights packed as
 * for mbx_deddqn_policy) followed by DE_DDQN_Optimizer.update (optimizer/de_ddqn_optimizer.py:131-220), i.e. the body of
 * DE_DDQN_Agent.rollout_episode (agent/de_ddqn_agent.py:108-117) -- with the lane's population, costs, credit history and window resident in
 * shared memory for the whole launch.  Bit-identical to n_steps x (mbx_deddqn_policy ; mbx_deddqn_step): same state `st`, same `d_feature`
 * [B,99] (in: the features mbx_deddqn_init / a previous step or run wrote; out: the features after the last step).
 *   d_return [B] f32  in/out: the episode return, += max(parent - trial, 0) per step (a NaN / inf reward counts as 0)
 * Lanes with st->done set are left untouched; a lane that finishes inside the launch stops there.  Philox only (cfg->use_tape must be 0). */
int mbx_deddqn_run(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                   const MbxDdqnState* st, float* d_feature, const float* d_weights, int n_weights, float* d_return, int n_steps, int B,
                   void* stream);
/* lanes that share one CTA of the persistent kernel (a build-time constant, for grid sizing by callers that care) */
int mbx_deddqn_run_lanes_per_cta(void);

/* ---------------------------------------------------------------------------------------------
 * GLEET backbone (PSO family, whole swarm per step, per-particle action) -- replaces GLEET_Optimizer.init_population /
 * update / observe (optimizer/gleet_optimizer.py:76-112, 193-314, 128-158) for B lanes at once; the reference's fixed
 * configuration: ps = 100, c = 4.1, w = 0.9 decaying by 0.5 / (maxFEs / ps) per step, clipping border, 'direct' reward
 * x 100 (:10-32).  State (all device, updated in place):
 *   x, v, pbest_x [B,NP,D] f32 ; c_cost, pbest [B,NP] f32 ; pni [B,NP] i32 (per_no_improve) ; gbest_x [B,D] f32 ;
 *   gbest_val [B] f32 ; gbest_idx, no_improve, fes, log_index, curve_len, gen [B] i32 ; w [B] f64 (inertia) ;
 *   max_cost [B] f32 (= min of the initial costs, :49) ; pbest_feature [B,NP,9] f32 ; gbest_feature [B,9] f32 ;
 *   curve [B, MBX_CURVE_LEN] f32 ; done [B] u8                                                                 */
typedef struct MbxGleetState {
    float *x, *v, *pbest_x, *c_cost, *pbest;
    int32_t* pni;
    float *gbest_x, *gbest_val;
    int32_t *gbest_idx, *no_improve, *fes, *log_index, *curve_len, *gen;
    double* w;
    float *max_cost, *pbest_feature, *gbest_feature;
    float* curve;
    uint8_t* done;
} MbxGleetState;

/* Recorded draws for parity tests (device, float64; null when use_tape=0):
 *   init_pos_u, init_vel_u [B,NP,D] ; rand1, rand2 [B,NP] ; noise [B,3,NP]                                   */
typedef struct MbxGleetTape {
    const double *init_pos_u, *init_vel_u, *rand1, *rand2, *noise;
} MbxGleetTape;

/* d_state [B,NP,27] f32 = observe (9) | pbest_feature (9) | gbest_feature (9) per particle (:110-112, :297-300) */
int mbx_gleet_init(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst,
                   const MbxRunConfig* cfg, const MbxGleetState* st, const MbxGleetTape* tape, float* d_state, int B,
                   void* stream);

/* d_action [B,NP] f32 in [0,1] (share of the pbest attraction per particle) ; outputs d_state [B,NP,27] f32,
 * d_reward [B] f32 (100 (gbest_before - gbest_after) / max_cost), st->done.  Finished lanes are skipped (reward = 0). */
int mbx_gleet_step(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst,
                   const MbxRunConfig* cfg, const MbxGleetState* st, const MbxGleetTape* tape, const float* d_action,
                   float* d_state, float* d_reward, int B, void* stream);

/* ---------------------------------------------------------------------------------------------
 * RL_PSO backbone (PSO family, ONE particle per step, scalar action) -- replaces RL_PSO_Optimizer.init_population /
 * update / __get_state (optimizer/rl_pso_optimizer.py:32-62, 79-148, 64-65) for B lanes at once; fixed configuration
 * NP = 100, c = 2.05, w = 0.9 decaying by 0.5 / (maxFEs / NP) per step (:10-24).  State (device, updated in place):
 *   x, v, pbest_x [B,NP,D] f32 ; c_cost, pbest [B,NP] f32 ; gbest_x [B,D] f32 ; gbest_val, max_cost [B] f32 (max_cost =
 *   max of the initial costs, :44) ; w [B] f64 ; gbest_idx, cur_index, fes, log_index, curve_len [B] i32 ;
 *   curve [B, MBX_CURVE_LEN] f32 ; done [B] u8                                                                   */
typedef struct MbxRlPsoState {
    float *x, *v, *pbest_x, *c_cost, *pbest;
    float *gbest_x, *gbest_val, *max_cost;
    double* w;
    int32_t *gbest_idx, *cur_index, *fes, *log_index, *curve_len;
    float* curve;
    uint8_t* done;
} MbxRlPsoState;

/* Recorded draws (device, float64; null when use_tape=0): init_pos_u, init_vel_u [B,NP,D] ; rand1 [B] (this step's
 * np.random.rand()) ; noise [B,3,n] with n = NP for init and n = 1 for a step.                                       */
typedef struct MbxRlPsoTape {
    const double *init_pos_u, *init_vel_u, *rand1, *noise;
} MbxRlPsoTape;

/* d_state [B, 2D] f32 = gbest_position | current_position[cur_index] (:64-65) */
int mbx_rl_pso_init(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst,
                    const MbxRunConfig* cfg, const MbxRlPsoState* st, const MbxRlPsoTape* tape, float* d_state, int B,
                    void* stream);

/* d_action [B] f32 (the reference feeds it unclamped as rand2) ; outputs d_state [B,2D] f32 (of the NEXT particle),
 * d_reward [B] f32 ((cost_before - cost_after) / (max_cost - gbest)), st->done.  Finished lanes: reward = 0.         */
int mbx_rl_pso_step(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst,
                    const MbxRunConfig* cfg, const MbxRlPsoState* st, const MbxRlPsoTape* tape, const float* d_action,
                    float* d_state, float* d_reward, int B, void* stream);

/* ---------------------------------------------------------------------------------------------
 * QLPSO backbone (PSO family, ONE particle per step; the integer action 0..3 selects a ring neighbourhood of 4 / 8 / 16 /
 * 30 particles) -- replaces QLPSO_Optimizer.init_population / update (optimizer/qlpso_optimizer.py:74-89, 91-126) for B
 * lanes at once; fixed configuration NP = 30, C = 1.49618, W = 0.729844 (:22-25); NP <= 128.  State (device, in place):
 *   x, v, pbest_x [B,NP,D] f32 ; cost [B,NP] f32 (current costs) ; gbest_cost [B] f32 ; diversity [B] f64 (mean distance
 *   to the centroid, numpy reduction order, :45-46) ; q_state [B,NP] i32 (last action of every particle, random 0..3 at
 *   start) ; pointer, fes, log_index, curve_len [B] i32 ; curve [B, MBX_CURVE_LEN] f32 ; done [B] u8                  */
typedef struct MbxQlpsoState {
    float *x, *v, *pbest_x, *cost, *gbest_cost;
    double* diversity;
    int32_t *q_state, *pointer, *fes, *log_index, *curve_len;
    float* curve;
    uint8_t* done;
} MbxQlpsoState;

/* Recorded draws (device; null when use_tape=0): init_pos_u [B,NP,D] f64 ; init_state [B,NP] i32 ; rand2 [B,2] f64 (this
 * step's two np.random.rand() scalars, neighbourhood term first) ; noise [B,3,n] f64, n = NP for init, 1 for a step.   */
typedef struct MbxQlpsoTape {
    const double* init_pos_u;
    const int32_t* init_state;
    const double *rand2, *noise;
} MbxQlpsoTape;

/* d_state [B] i32 = q_state of the particle the next action applies to (:89) */
int mbx_qlpso_init(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                   const MbxQlpsoState* st, const MbxQlpsoTape* tape, int32_t* d_state, int B, void* stream);

/* d_action [B] i32 in 0..3 (anything else: the particle is its own neighbourhood) ; outputs d_state [B] i32 (of the NEXT
 * particle), d_reward [B] f32 in {2, 1, 0, -2} (cal_reward, :7-16), st->done.  Finished lanes: reward = 0.           */
int mbx_qlpso_step(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                   const MbxQlpsoState* st, const MbxQlpsoTape* tape, const int32_t* d_action, int32_t* d_state,
                   float* d_reward, int B, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Landscape features of sampled walks -- replaces cal_fdc / cal_rie / cal_acf / cal_nop
 * (optimizer/dedqn_optimizer.py:8-76) and the DFDC / DRIE statistics of RL_HPSDE
 * (optimizer/rl_hpsde_optimizer.py:169-216) for B walks at once.
 *   d_sample [B,S,D] f64, d_cost [B,S] f64, d_ref [B,D] f64 or null (null: the best sample of each walk is the reference)
 *   d_out [B,6] f64 = fdc, rie, acf, nop, dfdc_r, drie_r   (the reference thresholds the last two at 0.15 / 0.5)   */
int mbx_landscape_features(const double* d_sample, const double* d_cost, const double* d_ref, int B, int S, int D,
                           double* d_out, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Progressive random walk of RL_HPSDE's state computation -- replaces `__progressive_random_walk`
 * (optimizer/rl_hpsde_optimizer.py:150-167) for B lanes at once; the walk's costs then go through mbx_eval and its statistics through
 * mbx_landscape_features (columns dfdc_r, drie_r).
 *   d_zone_u, d_r_u [B,D] f64 uniforms (start half / start offset) ; d_rD [B] i32 (the dimension that starts on a box face) ;
 *   d_step_u [B,n_steps,D] f64 uniforms ; lb, ub: the box ; step_size: 10 in the reference
 *   outputs: d_out [B,n_steps+1,D] f64 ; optional d_out32 (same shape, f32: the points as they are evaluated)                       */
int mbx_progressive_walk(const double* d_zone_u, const double* d_r_u, const int32_t* d_rD, const double* d_step_u, double lb, double ub,
                         double step_size, int B, int D, int n_steps, double* d_out, float* d_out32, void* stream);

/* ---------------------------------------------------------------------------------------------
 * GLEET attention actor, rollout forward -- replaces `Actor.forward` (agent/gleet_agent.py:386-440, called once per
 * generation from `rollout_episode`, :293-303) on agent/networks.py:47-356 for B swarms at once: embedder, encoder block
 * (4-head self-attention over the swarm, whole-swarm normalisation, feed-forward), decoder block (the exploration /
 * exploitation feature groups query the encoder output), mu / sigma heads, action = clamp(mu + sigma * noise, 0, 1).
 *   d_state   [B,NP,27] f32   the state mbx_gleet_init / mbx_gleet_step write
 *   d_weights [mbx_gleet_policy_weight_count()] f32, 16-byte aligned, packed as documented in
 *             metabox_b200/csrc/mbx_gleet_policy.cuh (metabox_b200/agent.py GLEET_Actor.pack() produces it from the reference's
 *             state_dict); read on every launch, so an optimiser may update it in place between launches
 *   d_noise   [B,NP] f32 standard-normal draws, or null (action = clamp(mu, 0, 1))
 *   outputs:  d_action [B,NP] f32 ; optional (null to skip) d_mu [B,NP], d_sigma [B,NP], d_logits [B,NP,16] (the decoder output
 *             the reference hands to its critic, `to_critic=True`)
 *   NP <= 128.                                                                                                     */
int mbx_gleet_policy_weight_count(void);
int mbx_gleet_policy(const float* d_state, const float* d_weights, int n_weights, const float* d_noise, float* d_action,
                     float* d_mu, float* d_sigma, float* d_logits, int B, int NP, void* stream);

/* ---------------------------------------------------------------------------------------------
 * RLEPSO actor, rollout forward -- replaces `Actor.forward` (agent/rlepso_agent.py:22-47; two MLPs 1 -> 64 -> 32 -> 35), called once per
 * generation from `rollout_episode` (:294-303), for B lanes at once.
 *   d_state   [B] (or [B,1]) f32 = fes / maxFEs, the feature mbx_rlepso_init / mbx_rlepso_step write
 *   d_weights [mbx_rlepso_policy_weight_count()] f32, packed as documented in metabox_b200/csrc/mbx_rlepso_policy.cuh
 *             (agent.py RLEPSO_Actor.pack()); read on every launch
 *   d_noise   [B,35] f32 standard-normal draws, or null (action = clamp(mu, 0, 1))
 *   outputs:  d_action [B,35] f32 = clamp(mu + sigma * noise, 0, 1) ; optional d_mu, d_sigma [B,35]                              */
int mbx_rlepso_policy_weight_count(void);
int mbx_rlepso_policy(const float* d_state, const float* d_weights, int n_weights, const float* d_noise, float* d_action,
                      float* d_mu, float* d_sigma, int B, void* stream);

/* ---------------------------------------------------------------------------------------------
 * RL_PSO policy network, rollout forward -- replaces `PolicyNetwork.forward` (agent/rl_pso_agent.py:22-45; two MLPs 2*dim -> 32 -> 8 -> 1),
 * called once per function evaluation from `rollout_episode` (:116-124), for B lanes at once.
 *   d_state   [B, n_in] f32, n_in = 2*dim <= 64: the state mbx_rl_pso_init / mbx_rl_pso_step write
 *   d_weights [mbx_rl_pso_policy_weight_count(n_in)] f32, packed as documented in metabox_b200/csrc/mbx_pso_policies.cuh
 *   d_noise   [B] f32 standard-normal draws, or null (eps = 0)
 *   outputs:  d_action [B] f32 (incl. the reference's transformation of samples outside (0, 1), :33-34) ; optional d_mu, d_sigma [B] */
int mbx_rl_pso_policy_weight_count(int n_in);
int mbx_rl_pso_policy(const float* d_state, const float* d_weights, int n_weights, const float* d_noise, float* d_action,
                      float* d_mu, float* d_sigma, int B, int n_in, void* stream);

/* QLPSO tabular policy -- replaces `__get_action` (agent/qlpso_agent.py:33-36): action ~ softmax(Q[state]).
 *   d_q_state [B] i32 in 0..3 (what mbx_qlpso_init / mbx_qlpso_step write) ; d_q_table [4,4] f64 ; d_u [B] f32 uniforms in [0, 1)
 *   output:   d_action [B] i32 in 0..3 = the first k with u < cumulative softmax probability (inverse-CDF sampling)               */
int mbx_qlpso_policy(const int32_t* d_q_state, const double* d_q_table, const float* d_u, int32_t* d_action, int B, void* stream);

/* ---------------------------------------------------------------------------------------------
 * DE-DDQN Q-network, greedy rollout action -- replaces `__get_action(state, epsilon_greedy=False)` (agent/de_ddqn_agent.py:59-68) on
 * the MLP 99 -> 100 -> 100 -> 100 -> 100 -> 4 (:21-25) for B lanes at once; called once per function evaluation from
 * `rollout_episode` (:108-117).
 *   d_state   [B,99] f32   the features mbx_deddqn_init / mbx_deddqn_step write (NaN / inf entries count as 0)
 *   d_weights [mbx_deddqn_policy_weight_count()] f32, 16-byte aligned, packed as documented in
 *             metabox_b200/csrc/mbx_deddqn_policy.cuh (agent.py DE_DDQN_Net.pack()); read on every launch
 *   outputs:  d_action [B] i32 = argmax_a Q (first maximum) ; optional d_q [B,4] f32                                  */
int mbx_deddqn_policy_weight_count(void);
int mbx_deddqn_policy(const float* d_state, const float* d_weights, int n_weights, int32_t* d_action, float* d_q, int B,
                      void* stream);

/* ---------------------------------------------------------------------------------------------
 * Random_search -- replaces Random_search.run_episode (optimizer/random_search.py:17-59): the whole
 * episode (maxFEs/NP iterations) in one launch per lane.  d_tape_u (optional) [B, n_iter, NP, D] f64.
 *   outputs: curve [B, MBX_CURVE_LEN] f32, curve_len [B] i32, fes [B] i32                           */
int mbx_random_search(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst,
                      const MbxRunConfig* cfg, const double* d_tape_u, int n_tape_iter,
                      float* d_curve, int32_t* d_curve_len, int32_t* d_fes, int B, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* METABOX_B200_H */
