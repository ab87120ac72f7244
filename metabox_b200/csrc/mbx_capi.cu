// mbx_capi.cu -- the extern "C" boundary declared in include/metabox_b200.h.
#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include "mbx_deddqn_policy.cuh"
#include "mbx_gleet_policy.cuh"
#include "mbx_launch.h"
#include "mbx_pso_policies.cuh"
#include "mbx_rlepso_policy.cuh"

static thread_local char g_err[512] = "";

void mbx_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

#define MBX_REQUIRE(cond, ...)          \
    do {                                \
        if (!(cond)) {                  \
            mbx_set_error(__VA_ARGS__); \
            return MBX_EINVAL;          \
        }                               \
    } while (0)

extern "C" {

const char* mbx_last_error(void) { return g_err; }
int mbx_version(void) { return 100; }
int mbx_sizeof_instance(void) { return (int)sizeof(MbxInstance); }
int mbx_sizeof_runconfig(void) { return (int)sizeof(MbxRunConfig); }

int mbx_eval(int kind, const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const void* d_x,
             int x_is_f64, void* d_out, int B, int NP, int D, int subtract_optimum, uint64_t seed, uint32_t gen,
             const double* d_noise_tape, void* stream) {
    MBX_REQUIRE(B >= 0 && NP >= 0 && D > 0 && D <= MBX_MAX_DIM, "mbx_eval: bad shape B=%d NP=%d D=%d", B, NP, D);
    if (B == 0 || NP == 0) return MBX_OK;                       // empty input: nothing to do
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && d_x && d_out, "mbx_eval: null pointer");
    MBX_REQUIRE(B <= 65535, "mbx_eval: B=%d exceeds gridDim.y; split the call", B);
    return mbx_launch_eval(kind, D, x_is_f64, d_inst, d_consts, d_lane_inst, d_x, d_out, B, NP, subtract_optimum, seed, gen,
                           d_noise_tape, (cudaStream_t)stream);
}

static int check_cfg(const char* who, const MbxRunConfig* cfg, int B) {
    MBX_REQUIRE(cfg, "%s: null config", who);
    MBX_REQUIRE(cfg->NP >= 5 && cfg->NP <= MBX_MAX_NP, "%s: NP=%d outside [5,%d]", who, cfg->NP, MBX_MAX_NP);
    MBX_REQUIRE(cfg->dim > 0 && cfg->dim <= MBX_MAX_DIM, "%s: dim=%d", who, cfg->dim);
    MBX_REQUIRE(cfg->maxFEs > 0 && cfg->n_logpoint > 0 && cfg->n_logpoint + 1 <= MBX_CURVE_LEN && cfg->log_interval > 0,
                "%s: bad maxFEs/n_logpoint/log_interval", who);
    MBX_REQUIRE(B >= 0, "%s: B=%d", who, B);
    return MBX_OK;
}

int mbx_rlepso_init(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                    const MbxRlepsoState* st, const MbxRlepsoTape* tape, float* d_feature, int B, void* stream) {
    int rc = check_cfg("mbx_rlepso_init", cfg, B);
    if (rc) return rc;
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && st && d_feature, "mbx_rlepso_init: null pointer");
    MBX_REQUIRE(!cfg->use_tape || (tape && tape->init_pos_u && tape->init_vel_u), "mbx_rlepso_init: use_tape without tape");
    if (B == 0) return MBX_OK;
    return mbx_launch_rlepso_init(cfg->kind, d_inst, d_consts, d_lane_inst, cfg, st, tape, d_feature, B, (cudaStream_t)stream);
}

int mbx_rlepso_step(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                    const MbxRlepsoState* st, const MbxRlepsoTape* tape, const double* d_pci, const float* d_action,
                    float* d_feature, float* d_reward, int32_t* d_dbg, int B, void* stream) {
    int rc = check_cfg("mbx_rlepso_step", cfg, B);
    if (rc) return rc;
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && st && d_pci && d_action && d_feature && d_reward,
                "mbx_rlepso_step: null pointer");
    MBX_REQUIRE(cfg->NP % 5 == 0, "mbx_rlepso_step: NP=%d must be a multiple of the 5 coefficient groups", cfg->NP);
    MBX_REQUIRE(!cfg->use_tape || (tape && tape->clpso_u && tape->fdr_u && tape->tour && tape->rand1 && tape->rand2 &&
                                   tape->reinit_u && tape->reinit_pos_u && tape->reinit_vel_u),
                "mbx_rlepso_step: use_tape without a complete tape");
    if (B == 0) return MBX_OK;
    return mbx_launch_rlepso_step(cfg->kind, d_inst, d_consts, d_lane_inst, cfg, st, tape, d_pci, d_action, d_feature,
                                  d_reward, d_dbg, B, (cudaStream_t)stream);
}

int mbx_lde_init(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                 const MbxLdeState* st, const MbxLdeTape* tape, float* d_feature, int B, void* stream) {
    int rc = check_cfg("mbx_lde_init", cfg, B);
    if (rc) return rc;
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && st && d_feature, "mbx_lde_init: null pointer");
    MBX_REQUIRE(!cfg->use_tape || (tape && tape->init_u), "mbx_lde_init: use_tape without tape");
    return mbx_launch_lde_init(cfg->kind, d_inst, d_consts, d_lane_inst, cfg, st, tape, d_feature, B, (cudaStream_t)stream);
}

int mbx_lde_step(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                 const MbxLdeState* st, const MbxLdeTape* tape, const float* d_action, float* d_feature, float* d_reward,
                 float* d_dbg, int B, void* stream) {
    int rc = check_cfg("mbx_lde_step", cfg, B);
    if (rc) return rc;
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && st && d_action && d_feature && d_reward, "mbx_lde_step: null pointer");
    MBX_REQUIRE(!cfg->use_tape || (tape && tape->pbest_idx && tape->r1 && tape->r2 && tape->cross_u && tape->jrand),
                "mbx_lde_step: use_tape without a complete tape");
    return mbx_launch_lde_step(cfg->kind, d_inst, d_consts, d_lane_inst, cfg, st, tape, d_action, d_feature, d_reward, d_dbg, B,
                               (cudaStream_t)stream);
}

int mbx_deddqn_init(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                    const MbxDdqnState* st, const MbxDdqnTape* tape, float* d_feature, int B, void* stream) {
    int rc = check_cfg("mbx_deddqn_init", cfg, B);
    if (rc) return rc;
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && st && d_feature, "mbx_deddqn_init: null pointer");
    MBX_REQUIRE(!cfg->use_tape || (tape && tape->init_u && tape->r5), "mbx_deddqn_init: use_tape without tape");
    return mbx_launch_deddqn_init(cfg->kind, d_inst, d_consts, d_lane_inst, cfg, st, tape, d_feature, B, (cudaStream_t)stream);
}

int mbx_deddqn_step(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                    const MbxDdqnState* st, const MbxDdqnTape* tape, const int32_t* d_action, float* d_feature,
                    float* d_reward, float* d_dbg, int B, void* stream) {
    int rc = check_cfg("mbx_deddqn_step", cfg, B);
    if (rc) return rc;
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && st && d_action && d_feature && d_reward, "mbx_deddqn_step: null pointer");
    MBX_REQUIRE(!cfg->use_tape || (tape && tape->r5), "mbx_deddqn_step: use_tape without tape");
    return mbx_launch_deddqn_step(cfg->kind, d_inst, d_consts, d_lane_inst, cfg, st, tape, d_action, d_feature, d_reward, d_dbg,
                                  B, (cudaStream_t)stream);
}

int mbx_deddqn_run(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                   const MbxDdqnState* st, float* d_feature, const float* d_weights, int n_weights, float* d_return, int n_steps, int B,
                   void* stream) {
    int rc = check_cfg("mbx_deddqn_run", cfg, B);
    if (rc) return rc;
    MBX_REQUIRE(n_steps >= 0, "mbx_deddqn_run: n_steps=%d", n_steps);
    MBX_REQUIRE(n_weights == ddq::N_WEIGHTS, "mbx_deddqn_run: %d packed weights, expected %d", n_weights, ddq::N_WEIGHTS);
    MBX_REQUIRE(!cfg->use_tape, "mbx_deddqn_run: the persistent kernel draws from Philox only (tape replay goes through mbx_deddqn_step)");
    if (B == 0 || n_steps == 0) return MBX_OK;
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && st && d_feature && d_weights && d_return, "mbx_deddqn_run: null pointer");
    return mbx_launch_deddqn_run(cfg->kind, d_inst, d_consts, d_lane_inst, cfg, st, d_feature, d_weights, d_return, n_steps, B,
                                 (cudaStream_t)stream);
}

int mbx_random_search(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                      const double* d_tape_u, int n_tape_iter, float* d_curve, int32_t* d_curve_len, int32_t* d_fes, int B,
                      void* stream) {
    int rc = check_cfg("mbx_random_search", cfg, B);
    if (rc) return rc;
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && d_curve && d_curve_len && d_fes, "mbx_random_search: null pointer");
    MBX_REQUIRE(!cfg->use_tape || (d_tape_u && n_tape_iter * cfg->NP >= cfg->maxFEs),
                "mbx_random_search: tape shorter than the episode");
    if (B == 0) return MBX_OK;
    return mbx_launch_rs(cfg->kind, d_inst, d_consts, d_lane_inst, cfg, d_tape_u, n_tape_iter, d_curve, d_curve_len, d_fes, B,
                         (cudaStream_t)stream);
}


int mbx_gleet_init(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                   const MbxGleetState* st, const MbxGleetTape* tape, float* d_state, int B, void* stream) {
    int rc = check_cfg("mbx_gleet_init", cfg, B);
    if (rc) return rc;
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && st && d_state, "mbx_gleet_init: null pointer");
    MBX_REQUIRE(!cfg->use_tape || (tape && tape->init_pos_u && tape->init_vel_u), "mbx_gleet_init: use_tape without tape");
    return mbx_launch_gleet_init(cfg->kind, d_inst, d_consts, d_lane_inst, cfg, st, tape, d_state, B, (cudaStream_t)stream);
}

int mbx_gleet_step(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                   const MbxGleetState* st, const MbxGleetTape* tape, const float* d_action, float* d_state, float* d_reward,
                   int B, void* stream) {
    int rc = check_cfg("mbx_gleet_step", cfg, B);
    if (rc) return rc;
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && st && d_action && d_state && d_reward, "mbx_gleet_step: null pointer");
    MBX_REQUIRE(!cfg->use_tape || (tape && tape->rand1 && tape->rand2), "mbx_gleet_step: use_tape without tape");
    return mbx_launch_gleet_step(cfg->kind, d_inst, d_consts, d_lane_inst, cfg, st, tape, d_action, d_state, d_reward, B,
                                 (cudaStream_t)stream);
}


int mbx_landscape_features(const double* d_sample, const double* d_cost, const double* d_ref, int B, int S, int D,
                           double* d_out, void* stream) {
    MBX_REQUIRE(B >= 0 && S >= 3 && S <= 512 && D > 0, "mbx_landscape_features: bad shape B=%d S=%d D=%d (3 <= S <= 512)", B, S, D);
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_sample && d_cost && d_out, "mbx_landscape_features: null pointer");
    return mbx_launch_landscape(d_sample, d_cost, d_ref, B, S, D, d_out, (cudaStream_t)stream);
}


int mbx_rlepso_policy_weight_count(void) { return rlp::N_WEIGHTS; }

int mbx_rlepso_policy(const float* d_state, const float* d_weights, int n_weights, const float* d_noise, float* d_action,
                      float* d_mu, float* d_sigma, int B, void* stream) {
    MBX_REQUIRE(B >= 0, "mbx_rlepso_policy: bad shape B=%d", B);
    MBX_REQUIRE(n_weights == rlp::N_WEIGHTS, "mbx_rlepso_policy: %d packed weights, expected %d", n_weights, rlp::N_WEIGHTS);
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_state && d_weights && d_action, "mbx_rlepso_policy: null pointer");
    return mbx_launch_rlepso_policy(d_state, d_weights, d_noise, d_action, d_mu, d_sigma, B, (cudaStream_t)stream);
}

int mbx_progressive_walk(const double* d_zone_u, const double* d_r_u, const int32_t* d_rD, const double* d_step_u, double lb, double ub,
                         double step_size, int B, int D, int n_steps, double* d_out, float* d_out32, void* stream) {
    MBX_REQUIRE(B >= 0 && D >= 1 && n_steps >= 0 && ub > lb, "mbx_progressive_walk: bad arguments B=%d D=%d n_steps=%d lb=%g ub=%g", B, D,
                n_steps, lb, ub);
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_zone_u && d_r_u && d_rD && d_out && (d_step_u || n_steps == 0), "mbx_progressive_walk: null pointer");
    return mbx_launch_progressive_walk(d_zone_u, d_r_u, d_rD, d_step_u, lb, ub, step_size, B, D, n_steps, d_out, d_out32,
                                       (cudaStream_t)stream);
}

int mbx_rl_pso_policy_weight_count(int n_in) { return rpp::n_weights(n_in); }

int mbx_rl_pso_policy(const float* d_state, const float* d_weights, int n_weights, const float* d_noise, float* d_action,
                      float* d_mu, float* d_sigma, int B, int n_in, void* stream) {
    MBX_REQUIRE(B >= 0 && n_in >= 1 && n_in <= 64, "mbx_rl_pso_policy: bad shape B=%d n_in=%d (1 <= n_in = 2*dim <= 64)", B, n_in);
    MBX_REQUIRE(n_weights == rpp::n_weights(n_in), "mbx_rl_pso_policy: %d packed weights, expected %d for n_in=%d", n_weights,
                rpp::n_weights(n_in), n_in);
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_state && d_weights && d_action, "mbx_rl_pso_policy: null pointer");
    return mbx_launch_rl_pso_policy(d_state, d_weights, d_noise, d_action, d_mu, d_sigma, B, n_in, (cudaStream_t)stream);
}

int mbx_qlpso_policy(const int32_t* d_q_state, const double* d_q_table, const float* d_u, int32_t* d_action, int B, void* stream) {
    MBX_REQUIRE(B >= 0, "mbx_qlpso_policy: bad shape B=%d", B);
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_q_state && d_q_table && d_u && d_action, "mbx_qlpso_policy: null pointer");
    return mbx_launch_qlpso_policy(d_q_state, d_q_table, d_u, d_action, B, (cudaStream_t)stream);
}

int mbx_deddqn_policy_weight_count(void) { return ddq::N_WEIGHTS; }

int mbx_deddqn_policy(const float* d_state, const float* d_weights, int n_weights, int32_t* d_action, float* d_q, int B,
                      void* stream) {
    MBX_REQUIRE(B >= 0, "mbx_deddqn_policy: bad shape B=%d", B);
    MBX_REQUIRE(n_weights == ddq::N_WEIGHTS, "mbx_deddqn_policy: %d packed weights, expected %d", n_weights, ddq::N_WEIGHTS);
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_state && d_weights && d_action, "mbx_deddqn_policy: null pointer");
    return mbx_launch_deddqn_policy(d_state, d_weights, d_action, d_q, B, (cudaStream_t)stream);
}

int mbx_gleet_policy_weight_count(void) { return glp::N_WEIGHTS; }

int mbx_gleet_policy(const float* d_state, const float* d_weights, int n_weights, const float* d_noise, float* d_action,
                     float* d_mu, float* d_sigma, float* d_logits, int B, int NP, void* stream) {
    MBX_REQUIRE(B >= 0 && NP >= 1 && NP <= 128, "mbx_gleet_policy: bad shape B=%d NP=%d (1 <= NP <= 128)", B, NP);
    MBX_REQUIRE(n_weights == glp::N_WEIGHTS, "mbx_gleet_policy: %d packed weights, expected %d", n_weights, glp::N_WEIGHTS);
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_state && d_weights && d_action, "mbx_gleet_policy: null pointer");
    MBX_REQUIRE(((uintptr_t)d_weights & 15) == 0 && (!d_logits || ((uintptr_t)d_logits & 15) == 0),
                "mbx_gleet_policy: d_weights / d_logits must be 16-byte aligned");
    return mbx_launch_gleet_policy(d_state, d_weights, d_noise, d_action, d_mu, d_sigma, d_logits, B, NP, (cudaStream_t)stream);
}


int mbx_rl_pso_init(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                    const MbxRlPsoState* st, const MbxRlPsoTape* tape, float* d_state, int B, void* stream) {
    int rc = check_cfg("mbx_rl_pso_init", cfg, B);
    if (rc) return rc;
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && st && d_state, "mbx_rl_pso_init: null pointer");
    MBX_REQUIRE(!cfg->use_tape || (tape && tape->init_pos_u && tape->init_vel_u), "mbx_rl_pso_init: use_tape without tape");
    return mbx_launch_rl_pso_init(cfg->kind, d_inst, d_consts, d_lane_inst, cfg, st, tape, d_state, B, (cudaStream_t)stream);
}

int mbx_rl_pso_step(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                    const MbxRlPsoState* st, const MbxRlPsoTape* tape, const float* d_action, float* d_state, float* d_reward,
                    int B, void* stream) {
    int rc = check_cfg("mbx_rl_pso_step", cfg, B);
    if (rc) return rc;
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && st && d_action && d_state && d_reward, "mbx_rl_pso_step: null pointer");
    MBX_REQUIRE(!cfg->use_tape || (tape && tape->rand1), "mbx_rl_pso_step: use_tape without tape");
    return mbx_launch_rl_pso_step(cfg->kind, d_inst, d_consts, d_lane_inst, cfg, st, tape, d_action, d_state, d_reward, B,
                                  (cudaStream_t)stream);
}

int mbx_qlpso_init(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                   const MbxQlpsoState* st, const MbxQlpsoTape* tape, int32_t* d_state, int B, void* stream) {
    int rc = check_cfg("mbx_qlpso_init", cfg, B);
    if (rc) return rc;
    MBX_REQUIRE(cfg->NP <= 128, "mbx_qlpso_init: NP=%d > 128", cfg->NP);
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && st && d_state, "mbx_qlpso_init: null pointer");
    MBX_REQUIRE(!cfg->use_tape || (tape && tape->init_pos_u && tape->init_state), "mbx_qlpso_init: use_tape without tape");
    return mbx_launch_qlpso_init(cfg->kind, d_inst, d_consts, d_lane_inst, cfg, st, tape, d_state, B, (cudaStream_t)stream);
}

int mbx_qlpso_step(const MbxInstance* d_inst, const double* d_consts, const int32_t* d_lane_inst, const MbxRunConfig* cfg,
                   const MbxQlpsoState* st, const MbxQlpsoTape* tape, const int32_t* d_action, int32_t* d_state, float* d_reward,
                   int B, void* stream) {
    int rc = check_cfg("mbx_qlpso_step", cfg, B);
    if (rc) return rc;
    MBX_REQUIRE(cfg->NP <= 128, "mbx_qlpso_step: NP=%d > 128", cfg->NP);
    if (B == 0) return MBX_OK;
    MBX_REQUIRE(d_inst && d_consts && d_lane_inst && st && d_action && d_state && d_reward, "mbx_qlpso_step: null pointer");
    MBX_REQUIRE(!cfg->use_tape || (tape && tape->rand2), "mbx_qlpso_step: use_tape without tape");
    return mbx_launch_qlpso_step(cfg->kind, d_inst, d_consts, d_lane_inst, cfg, st, tape, d_action, d_state, d_reward, B,
                                 (cudaStream_t)stream);
}

}  // extern "C"
