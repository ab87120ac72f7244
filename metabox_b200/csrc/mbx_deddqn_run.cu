// mbx_deddqn_run.cu -- DE-DDQN as a PERSISTENT kernel: ONE WARP owns one lane for `n_steps` function evaluations --
// costs (+ an ascending copy), credit history and window resident in shared memory, the population rows in HBM / L2 (a step touches seven of
// them), the agent's Q-network (MLP 99 -> 100 -> 100 -> 100 -> 100 -> 4, greedy argmax; agent/de_ddqn_agent.py:21-25, 59-68) evaluated
// inside the kernel -- instead of one policy launch + one step launch per evaluation (59 900 of each per run at dim 30).  Reference loop:
// agent/de_ddqn_agent.py:108-117 around optimizer/de_ddqn_optimizer.py:131-220.
//
// Why a warp per lane: a DE-DDQN step is a chain of short DEPENDENT phases (mutation -> one fitness -> credit / window bookkeeping -> 99
// features -> 5 network layers), most of them a few threads wide.  Measured with 128 threads per lane (profiles/r2d_ddrun_phases.txt): 65 k
// cycles of latency per step, issue slots 38 % busy, 8 lanes per SM (register file).  The chain cannot be shortened much, so the lever is
// the number of chains in flight: 32 threads and ~8.6 KB of shared memory per lane put 24 lanes on an SM, barriers become __syncwarp, and
// the instruction count per lane-step falls with the number of idle warps.
//
// A CTA carries MBX_LANE_GROUPS = 8 lanes.  The Q-network runs CTA-wide on the 8 lanes at once (activations transposed [input][lane],
// thread o = neuron o: one weight load serves 8 lanes, so the 161 KB of packed weights stream from L2 / L1 once per CTA-step); everything
// else is per warp, so lanes of different functions and finished lanes share a CTA without constraints.
//
// Arithmetic: the step is `dd_step_core` (mbx_deddqn.cuh), the very code the per-step kernel runs -- its reductions are written over 128
// "virtual threads" so that their summation trees do not depend on the group size; the network uses the per-neuron fma chain of
// mbx_deddqn_policy.cuh (bias, then inputs in ascending order).  A persistent launch therefore leaves bit-identical state, features,
// curves and returns to n_steps x (mbx_deddqn_policy ; mbx_deddqn_step) -- tests/test_gpu_deddqn_run.py.
#define MBX_NT 32
#define MBX_LANE_GROUPS 8
#include "mbx_deddqn.cuh"
#include "mbx_deddqn_policy.cuh"

#ifndef MBX_DDRUN_MIN_CTAS
#define MBX_DDRUN_MIN_CTAS 3
#endif
static_assert(MBX_LANE_GROUPS == ddq::TL, "the in-kernel Q-network is tiled for ddq::TL lanes per CTA");

#ifdef MBX_PROFILE
// slots 0..6: phases inside dd_step_core; 8: loop-top CTA barrier; 9: Q-network; 10: whole environment step (thread 0 of every CTA)
extern "C" int mbx_prof_read_deddqn_run(unsigned long long* h16, int reset) {
    if (cudaMemcpyFromSymbol(h16, g_prof, sizeof(unsigned long long) * 16) != cudaSuccess) return MBX_ECUDA;
    if (reset) {
        unsigned long long z[16] = {0};
        if (cudaMemcpyToSymbol(g_prof, z, sizeof(z)) != cudaSuccess) return MBX_ECUDA;
    }
    return MBX_OK;
}
#endif

// shared-memory plan of ONE lane, sized by the run-time population size (dim 30, pop 100: 8.6 KB)
template <int D, int KIND>
struct DdRunLayout {
    size_t oHdr, oConst, oCost, oSorted, oXg, oXp, oNtot, oNsucc, oOmSum, oOmMax, oWin, oRed, oScr, oTrial, oFeat, oMisc, bytes;
    __host__ __device__ DdRunLayout(int NP) {
        const int NPp = (NP + 3) & ~3;
        constexpr int kConst = 0;      // instance constants are read from global memory / L1 (shared by the lanes of an instance), BBOB and protein alike
        constexpr int kScratch = DdqnSmem<D, KIND>::kScratch;
        constexpr int Dp = (D + 3) & ~3;
        oHdr = 0;
        oConst = (sizeof(MbxInstance) + 15) & ~(size_t)15;
        oOmSum = oConst + sizeof(double) * kConst;
        oOmMax = oOmSum + sizeof(double) * 16 * DD_GEN_MAX;
        oWin = oOmMax + sizeof(double) * 16 * DD_GEN_MAX;
        oRed = oWin + sizeof(double) * DD_W * DD_WS;
        oScr = oRed + sizeof(double) * 4;
        oCost = oScr + sizeof(double) * kScratch;
        oSorted = oCost + sizeof(float) * NPp;
        oXg = oSorted + sizeof(float) * NPp;
        oXp = oXg + sizeof(float) * Dp;
        oTrial = oXp + sizeof(float) * Dp;
        oFeat = oTrial + sizeof(float) * Dp;
        oNtot = oFeat + sizeof(float) * 100;
        oNsucc = oNtot + sizeof(int32_t) * 4 * DD_GEN_MAX;
        oMisc = oNsucc + sizeof(int32_t) * 16 * DD_GEN_MAX;                  // mf f32 [8] | mi i32 [24]
        bytes = (oMisc + 32 + 96 + 15) & ~(size_t)15;
    }
};

template <int D, int KIND>
__global__ void __launch_bounds__(MBX_NT * MBX_LANE_GROUPS, MBX_DDRUN_MIN_CTAS)
deddqn_run_kernel(const MbxInstance* __restrict__ inst, const double* __restrict__ consts, const int32_t* __restrict__ lane_inst,
                  MbxRunConfig cfg, MbxDdqnState S, float* __restrict__ feature, const float* __restrict__ weights,
                  float* __restrict__ ret, const int n_steps, const int B) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int grp = threadIdx.x / MBX_NT, tid = MBX_TID, NP = cfg.NP;
    const int b = blockIdx.x * MBX_LANE_GROUPS + grp;
    const DdRunLayout<D, KIND> Lay(NP);
    unsigned char* sm = smem_raw + (size_t)grp * Lay.bytes;
    const bool mine = b < B && S.done[b] == 0;                             // lanes this launch advances (finished lanes stay frozen)
    MbxInstance* sh = reinterpret_cast<MbxInstance*>(sm + Lay.oHdr);
    double* sC = reinterpret_cast<double*>(sm + Lay.oConst);
    DdLane Ln;
    static_assert(D <= MBX_NT, "one thread per dimension of the trial");
    Ln.X = (b < B) ? S.X + (size_t)b * NP * D : nullptr;                  // population rows stay in HBM / L2
    Ln.Xg = reinterpret_cast<float*>(sm + Lay.oXg);
    Ln.Xp = reinterpret_cast<float*>(sm + Lay.oXp);
    Ln.n_tot = reinterpret_cast<int32_t*>(sm + Lay.oNtot);
    Ln.n_succ = reinterpret_cast<int32_t*>(sm + Lay.oNsucc);
    Ln.om_sum = reinterpret_cast<double*>(sm + Lay.oOmSum);
    Ln.om_max = reinterpret_cast<double*>(sm + Lay.oOmMax);
    Ln.curve = mine ? S.curve + (size_t)b * MBX_CURVE_LEN : nullptr;
    Ln.cost_wt = nullptr; Ln.win_wt = nullptr; Ln.r_wt = nullptr;
    Ln.sc = reinterpret_cast<float*>(sm + Lay.oCost);
    Ln.sorted = reinterpret_cast<float*>(sm + Lay.oSorted);
    Ln.swin = reinterpret_cast<double*>(sm + Lay.oWin);
    Ln.red = reinterpret_cast<double*>(sm + Lay.oRed);
    Ln.scratch = reinterpret_cast<double*>(sm + Lay.oScr);
    Ln.trial = reinterpret_cast<float*>(sm + Lay.oTrial);
    Ln.mf = reinterpret_cast<float*>(sm + Lay.oMisc);
    Ln.mi = reinterpret_cast<int*>(sm + Lay.oMisc + 32);
    Ln.feat = reinterpret_cast<float*>(sm + Lay.oFeat);
    float* sXa = reinterpret_cast<float*>(smem_raw + (size_t)MBX_LANE_GROUPS * Lay.bytes);       // CTA-shared: activations [100][8] x 2, Q [8][4]
    float* sXb = sXa + ddq::NHID * ddq::TL;
    float* sQ = sXb + ddq::NHID * ddq::TL;
    DdRegs R{};
    R.done = true;
    const double* cC = sC;
    const double* gC = consts;
    float racc = 0.f;
    if (mine) {
        // ---- the lane's state: HBM -> shared memory / registers, once per launch ------------------------------------------------
        const MbxInstance* gi = inst + lane_inst[b];
        gC = consts + gi->off;
        {
            const int* hs = reinterpret_cast<const int*>(gi);
            int* hd = reinterpret_cast<int*>(sh);
            for (int k = tid; k < (int)(sizeof(MbxInstance) / 4); k += MBX_NT) hd[k] = hs[k];
            cC = gC;
        }
        const size_t pb = (size_t)b * NP;
        for (int i = tid; i < NP; i += MBX_NT) Ln.sc[i] = S.cost[pb + i];
        for (int d = tid; d < D; d += MBX_NT) { Ln.Xg[d] = S.X_gbest[(size_t)b * D + d]; Ln.Xp[d] = S.X_prebest[(size_t)b * D + d]; }
        for (int k = tid; k < 4 * DD_GEN_MAX; k += MBX_NT) Ln.n_tot[k] = S.n_tot[(size_t)b * 4 * DD_GEN_MAX + k];
        for (int k = tid; k < 16 * DD_GEN_MAX; k += MBX_NT) {
            Ln.n_succ[k] = S.n_succ[(size_t)b * 16 * DD_GEN_MAX + k];
            Ln.om_sum[k] = S.om_sum[(size_t)b * 16 * DD_GEN_MAX + k];
            Ln.om_max[k] = S.om_max[(size_t)b * 16 * DD_GEN_MAX + k];
        }
        for (int k = tid; k < DD_W * DD_WS; k += MBX_NT) Ln.swin[k] = S.win[(size_t)b * DD_W * DD_WS + k];
        if (tid < 5) Ln.mi[5 + tid] = S.r[(size_t)b * 5 + tid];            // the indices the NEXT step mutates with
        for (int k = tid; k < DD_NFEAT; k += MBX_NT) Ln.feat[k] = feature[(size_t)b * DD_NFEAT + k];
        const int32_t* ints = S.ints + (size_t)b * 8;
        R.p = ints[DD_I_POINTER]; R.gen = ints[DD_I_GEN]; R.stag = ints[DD_I_STAG]; R.prow = ints[DD_I_PROW]; R.rebound = ints[DD_I_REBOUND];
        R.wlen = ints[DD_I_WLEN]; R.wseq = ints[DD_I_WSEQ];
        R.fes = S.fes[b]; R.log_index = S.log_index[b]; R.curve_len = S.curve_len[b];
        R.cgb = S.c3[(size_t)b * 3]; R.cgw = S.c3[(size_t)b * 3 + 1]; R.cpb = S.c3[(size_t)b * 3 + 2];
        R.done = false;
        racc = ret[b];
    }
    MBX_SYNC();
    if (mine)                                                              // the ascending copy of the costs (stable ranks: a permutation)
        for (int i = tid; i < NP; i += MBX_NT) Ln.sorted[dd_rank(Ln.sc, i, NP)] = Ln.sc[i];
    const uint32_t lane = (uint32_t)(cfg.lane_base + b);
    Philox ph(cfg.seed);
    bool live = mine;
    PROF_INIT
    for (int step = 0; step < n_steps; ++step) {
        // CTA-wide: (i) publishes the previous step's features / the staged state, (ii) keeps the groups' weight streams together so
        // that a weight row fetched by one group is an L1 hit for the others, (iii) ends the launch early when every lane of the CTA is done
        if (__syncthreads_and(live ? 0 : 1)) break;
        PROF_MARK(8);
        // ---- the agent: greedy action of the Q-network on the sanitised features (de_ddqn_agent.py:59-68), for the CTA's 8 lanes at
        //      once: activations transposed [input][lane] in CTA-shared memory, thread o = neuron o for the 8 lanes; lanes that are
        //      finished ride along with zeros -------------------------------------------------------------------------------------
        for (int k = tid; k < ddq::NIN; k += MBX_NT) sXa[k * ddq::TL + grp] = live ? ddq::sanitise(Ln.feat[k]) : 0.f;
        if (live && tid < 5) Ln.mi[tid] = Ln.mi[5 + tid];                  // last step's fresh indices become this step's donors (:85)
        __syncthreads();
        const int t = threadIdx.x;
        if (t < 2 * ddq::NHID) ddq::hidden_neuron_x8<ddq::NIN, 20, 4>(weights + ddq::O_L0, sXa, t >> 1, t & 1, sXb);
        __syncthreads();
        if (t < 2 * ddq::NHID) ddq::hidden_neuron_x8<ddq::NHID, 20, 4>(weights + ddq::O_L1, sXb, t >> 1, t & 1, sXa);
        __syncthreads();
        if (t < 2 * ddq::NHID) ddq::hidden_neuron_x8<ddq::NHID, 20, 4>(weights + ddq::O_L2, sXa, t >> 1, t & 1, sXb);
        __syncthreads();
        if (t < 2 * ddq::NHID) ddq::hidden_neuron_x8<ddq::NHID, 20, 4>(weights + ddq::O_L3, sXb, t >> 1, t & 1, sXa);
        __syncthreads();
        if (t < ddq::TL * ddq::NACT) sQ[t] = ddq::q_value_x8(weights + ddq::O_OUT, sXa, t & 3, t >> 2);     // sQ[lane][action]
        __syncthreads();
        PROF_MARK(9);
        if (!live) continue;
        const int a = ddq::greedy(sQ + grp * ddq::NACT);
        // ---- the environment step (shared with the per-step kernel) ---------------------------------------------------------------
        const float rew = dd_step_core<D, KIND>(cfg, *sh, cC, gC, Ln, R, a, ph, lane, nullptr, nullptr, nullptr);
        racc = racc + ddq::sanitise(rew);                                  // the sweep's `ret += nan_to_num(reward)`, float32, in order
        live = !R.done;
        PROF_MARK(10);
    }
    MBX_SYNC();
    if (mine) {
        // ---- shared memory / registers -> HBM ------------------------------------------------------------------------------------
        const size_t pb = (size_t)b * NP;
        for (int i = tid; i < NP; i += MBX_NT) S.cost[pb + i] = Ln.sc[i];
        for (int d = tid; d < D; d += MBX_NT) { S.X_gbest[(size_t)b * D + d] = Ln.Xg[d]; S.X_prebest[(size_t)b * D + d] = Ln.Xp[d]; }
        for (int k = tid; k < 4 * DD_GEN_MAX; k += MBX_NT) S.n_tot[(size_t)b * 4 * DD_GEN_MAX + k] = Ln.n_tot[k];
        for (int k = tid; k < 16 * DD_GEN_MAX; k += MBX_NT) {
            S.n_succ[(size_t)b * 16 * DD_GEN_MAX + k] = Ln.n_succ[k];
            S.om_sum[(size_t)b * 16 * DD_GEN_MAX + k] = Ln.om_sum[k];
            S.om_max[(size_t)b * 16 * DD_GEN_MAX + k] = Ln.om_max[k];
        }
        for (int k = tid; k < DD_W * DD_WS; k += MBX_NT) S.win[(size_t)b * DD_W * DD_WS + k] = Ln.swin[k];
        if (tid < 5) S.r[(size_t)b * 5 + tid] = Ln.mi[5 + tid];
        for (int k = tid; k < DD_NFEAT; k += MBX_NT) feature[(size_t)b * DD_NFEAT + k] = Ln.feat[k];
        if (tid == 0) {
            int32_t* ints = S.ints + (size_t)b * 8;
            ints[DD_I_POINTER] = R.p; ints[DD_I_GEN] = R.gen; ints[DD_I_STAG] = R.stag; ints[DD_I_PROW] = R.prow;
            ints[DD_I_REBOUND] = R.rebound; ints[DD_I_WLEN] = R.wlen; ints[DD_I_WSEQ] = R.wseq;
            S.c3[(size_t)b * 3] = R.cgb; S.c3[(size_t)b * 3 + 1] = R.cgw;
            S.log_index[b] = R.log_index; S.curve_len[b] = R.curve_len; S.fes[b] = R.fes; S.done[b] = R.done ? 1 : 0;
            ret[b] = racc;
        }
    }
}

template <int D, int KIND>
static int dd_launch_run_t(const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                           const MbxDdqnState* s, float* feature, const float* weights, float* ret, int n_steps, int B, cudaStream_t st) {
    auto kern = deddqn_run_kernel<D, KIND>;
    const size_t bytes = DdRunLayout<D, KIND>(cfg->NP).bytes * MBX_LANE_GROUPS + sizeof(float) * (2 * ddq::NHID * ddq::TL + ddq::TL * ddq::NACT);
    // per launch, not once per process: the attribute belongs to the CURRENT device, and a process may drive several
    MBX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    kern<<<(B + MBX_LANE_GROUPS - 1) / MBX_LANE_GROUPS, MBX_NT * MBX_LANE_GROUPS, bytes, st>>>(inst, consts, lane_inst, *cfg, *s, feature,
                                                                                             weights, ret, n_steps, B);
    MBX_CUDA(cudaGetLastError());
    return MBX_OK;
}

int mbx_launch_deddqn_run(int kind, const MbxInstance* inst, const double* consts, const int32_t* lane_inst, const MbxRunConfig* cfg,
                          const MbxDdqnState* s, float* feature, const float* weights, float* ret, int n_steps, int B, cudaStream_t st) {
    MBX_FOR_BBOB_DIMS(MBX_CASE_BBOB_, dd_launch_run_t, inst, consts, lane_inst, cfg, s, feature, weights, ret, n_steps, B, st)
    if (cfg->dim == 12 && kind == MBX_KIND_PROTEIN) return dd_launch_run_t<12, MBX_KIND_PROTEIN>(inst, consts, lane_inst, cfg, s, feature, weights, ret, n_steps, B, st);
    mbx_set_error("mbx_deddqn_run: (kind=%d, dim=%d) is not compiled", kind, cfg->dim);
    return MBX_EUNSUP;
}
int mbx_deddqn_run_lanes_per_cta(void) { return MBX_LANE_GROUPS; }
