// deepterrainrl_b200 -- the env-step kernel (sm_100a), warp-per-environment formulation.
//
// One environment per WARP: lane j owns link j of the kinematic tree (21 links) and, for the dense solve, row j of
// the 23x23 system; every tree recursion is level-synchronous over the tree depth (9 for the dog) with the
// parent <-> child hand-off done by warp shuffles, so per-link quantities live in registers and shared memory only
// holds the mass matrix.  All spatial quantities are planar 3-vectors expressed in WORLD axes about the root joint's
// current position O (an inertial frame that instantaneously coincides with the root), which removes every
// per-link coordinate transform from the recursions: motion (w, vOx, vOy), force (nO, fx, fy), joint axis of a
// revolute joint at r (relative to O): S = (1, r_y, -r_x).
//
// One launch = the *controller half* of env-step k followed by the *physics half* of env-step k+1:
//
//   ctrl half  (reference: cSimCharacter::Update -> cDogController::Update, sim/DogController.cpp:229-268)
//     CRBA mass matrix + RNEA bias force with the reference's Cj (restates sim/RBDUtil.cpp:4-176 in planar algebra)
//     ApplyFeedback, stable-PD solve (register-resident LDL^T, one row per lane), gravity compensation,
//     virtual forces, torque clamp (sim/Joint.cpp:257-264), fall counters (sim/SimCharSoftFall.cpp:74-125),
//     cycle bookkeeping: reward, tuple record, episode statistics (scenarios/ScenarioExp.cpp:209-243)
//   [end of outer update: fall -> episode reset, scenarios/ScenarioPoliEval.cpp:110-125, ScenarioExp.cpp:83-98]
//   phys half  (reference: cWorld::Update + cGroundVar2D::Update + the head of the controller update)
//     num_sim_substeps x articulated-body forward dynamics with implicit contact / joint-limit terms (this project's
//     own physics model -- Bullet is not restatable); box corners are evaluated one per lane
//     contact bits, streaming terrain update, cycle timers, gait FSM (sim/DogController.cpp:805-845)
//     warps whose FSM reaches a cycle boundary build the 283-float policy state cooperatively and enqueue a decision
//
// The decision itself (MACE forward pass + action decode) runs in trl_decide.cuh between two launches of this kernel.
#include <cuda_runtime.h>

#include <algorithm>

#include "trl_terrain.cuh"
#include "trl_types.h"

// The file is compiled twice: as is (namespace trl), and through trl_step_cg.cu with -Xptxas -dlcm=cg (namespace trl_cg), where
// every global load bypasses L1.  The cg build serves the catch-up launches of the overlapped schedule: they read env state the
// decision kernel has just written, while step CTAs running on the same SM may already have pulled the neighbouring envs'
// words of the same 32-byte sectors into that SM's L1.
#if defined(TRL_CG_VARIANT)
#define TRL_IMPL_NS trl_cg
#else
#define TRL_IMPL_NS trl
#endif

namespace TRL_IMPL_NS {
using namespace trl;

__constant__ ModelConst c_model;
// character-type branches: a run-time constant of the scene (the raptor's controller differs from the dog's / goat's); an experiment
// build can fix it at compile time (-DTRL_FIXED_CHAR=0 dog / goat, 2 raptor) to see what the unused branches cost
#ifdef TRL_FIXED_CHAR
#define TRL_IS_RAPTOR(mc) (TRL_FIXED_CHAR == 2)
#else
#define TRL_IS_RAPTOR(mc) ((mc).char_type == 2)
#endif
// Lane-indexed tables (one entry per link / corner) are read from a global-memory mirror of the same struct: the constant cache
// serves one address per access, so `c_model.kp[lane]` costs one replay per distinct lane, while the mirror is one coalesced,
// L1-cached read-only load.  Warp-uniform reads stay in constant memory.
#ifndef TRL_TABLE_MIRROR
#define TRL_TABLE_MIRROR 1      // + 4.5 % on a B200 (profiles/step_kernel_r02_session2_ab.txt); 0 = the round-1 reads from constant memory
#endif
#if TRL_TABLE_MIRROR
__device__ ModelConst g_model;
#define LT(field, idx) __ldg(&g_model.field[(idx)])
#define LT2(field, i, j) __ldg(&g_model.field[(i)][(j)])
#else
#define LT(field, idx) c_model.field[(idx)]
#define LT2(field, i, j) c_model.field[(i)][(j)]
#endif

#ifndef TRL_WARPS_PER_BLOCK
#define TRL_WARPS_PER_BLOCK 4
#endif
constexpr int kWarpsPerBlock = TRL_WARPS_PER_BLOCK;   // one env per warp; 16 resident warps per SM at 128 registers whatever the CTA size
constexpr int kBlockThreads = kWarpsPerBlock * kWarp;
constexpr unsigned kFull = 0xffffffffu;
#ifndef TRL_STEP_MIN_BLOCKS
#define TRL_STEP_MIN_BLOCKS 4   // CTAs of 4 warps per SM the register budget is sized for
#endif
constexpr int kStepSkipPending = 8, kStepCatchUp = 16;   // flag bits of trl_step_kernel beyond ctrl (1) / phys (2) / end (4)
// `lists` packs: bits 0-2 the list new boundary envs are appended to, bits 3-5 the list a catch-up launch serves, bits 6-9 the number
// of env-steps a catch-up launch advances its envs by
constexpr double kClearMargin = 0.05;   // >= contact_tol * sqrt(1 + slope^2) for any slope the generators produce
constexpr int kZeroLane = 31;   // always idle (nj <= 23): its per-link registers are zero, used as the "no source" lane
constexpr int kTri = kMaxDof * (kMaxDof + 1) / 2;   // 276
// per-warp shared scratch (doubles)
constexpr int X_M = 0;
// Experiments (profiles/step_kernel_r01_source_phases.md: shuffles are 27 % of the kernel's instructions).  Each knob moves one
// family of lane-to-lane exchanges from warp shuffles (2 SHFL + register moves per double) to per-warp shared memory (128-bit
// stores / loads); values and operation order are untouched, so every variant is bit-identical to the default build
// (tests/test_simt_cpu.py checks that on the CPU emulator).  Default build: the LDL^T exchange on, the others off (measured).  -DTRL_SMEM_XCHG=1: all on.
#ifndef TRL_SMEM_XCHG
#define TRL_SMEM_XCHG 0
#endif
#ifndef TRL_ACCUM_SMEM
#define TRL_ACCUM_SMEM TRL_SMEM_XCHG     // child -> parent hand-off of the inward rounds (9 doubles x 9 rounds x 5 sub-steps)
#endif
#ifndef TRL_LDLT_SMEM
#define TRL_LDLT_SMEM 1                  // ON by default since round 2 (B200 A/B, profiles/step_kernel_variants_r02_ab2.txt: +2.8 % env-steps/s;
                                         // spill stores 404 -> 228 B).  pivot column of the register LDL^T + fused forward substitution (253 + 23 doubles)
#endif
#ifndef TRL_KIN_SMEM
#define TRL_KIN_SMEM TRL_SMEM_XCHG       // pointer-jumping prefix sums of the kinematics (26 doubles x 6 per env-step)
#endif
#ifndef TRL_OUTWARD_SMEM
#define TRL_OUTWARD_SMEM TRL_SMEM_XCHG   // parent -> child accelerations of the outward pass + floating-base broadcast (36 doubles x 5)
#endif
#ifndef TRL_CONTACT_SMEM
#define TRL_CONTACT_SMEM TRL_SMEM_XCHG   // corner lanes read their body's kinematics, owners collect contact forces (7-21 + 9 per contact, x 5)
#endif
#define TRL_SMEM_ALIGNED (TRL_ACCUM_SMEM || TRL_LDLT_SMEM || TRL_KIN_SMEM || TRL_OUTWARD_SMEM || TRL_CONTACT_SMEM)
constexpr int X_M_END = X_M + kTri + 4;
#if TRL_ACCUM_SMEM || TRL_CONTACT_SMEM
constexpr int kAccStride = 10;                              // doubles per lane slot: 9 values, padded to 80 B (128-bit accesses stay conflict-free)
constexpr int X_ACC = (X_M_END + 1) & ~1;                   // 16-byte aligned
constexpr int X_ACC_END = X_ACC + kWarp * kAccStride;
#else
constexpr int X_ACC_END = X_M_END;
#endif
#if TRL_LDLT_SMEM
constexpr int kColStride = kWarp + 2;                       // 32 column entries + the solved right-hand-side entry, even (16-byte rows)
constexpr int X_COL = (X_ACC_END + 1) & ~1;                 // two buffers: step j writes buffer j & 1
constexpr int X_COL_END = X_COL + 2 * kColStride;
#else
constexpr int X_COL_END = X_ACC_END;
#endif
#if TRL_KIN_SMEM
constexpr int X_PFX = (X_COL_END + 1) & ~1;                 // three buffers of 32 double2: prefix rounds alternate 0 / 1, buffer 2 = (cos, sin) exchange
constexpr int X_PFX_END = X_PFX + 3 * 2 * kWarp;
#else
constexpr int X_PFX_END = X_COL_END;
#endif
#if TRL_OUTWARD_SMEM
constexpr int kOutBuf = 3 * kWarp;                          // (aw, ax) as double2 [32] + ay [32]
constexpr int X_OUT = (X_PFX_END + 1) & ~1;                 // two buffers (level parity) + the 9 floating-base values of lane 0
constexpr int X_BASE = X_OUT + 2 * kOutBuf;
constexpr int X_OUT_END = X_BASE + 10;
#else
constexpr int X_OUT_END = X_PFX_END;
#endif
#if TRL_CONTACT_SMEM
constexpr int X_KIN = (X_OUT_END + 1) & ~1;                 // per-link (cw, sw) | (rx, ry) | (w, vx) as double2 [32] each, vy [32]
constexpr int X_KIN_END = X_KIN + 7 * kWarp;
#else
constexpr int X_KIN_END = X_OUT_END;
#endif
constexpr int X_END = TRL_SMEM_ALIGNED ? ((X_KIN_END + 1) & ~1) : X_KIN_END;
// the scratch pointer is only threaded through the helpers that need it, so the default build's code is untouched
#if TRL_KIN_SMEM
#define TRL_KIN_XS_DECL , double* xs, int lane
#define TRL_KIN_XS_ARG , xs, lane
#define TRL_RESET_XS_DECL , double* xs
#define TRL_RESET_XS_ARG , xs
#else
#define TRL_KIN_XS_DECL
#define TRL_KIN_XS_ARG
#define TRL_RESET_XS_DECL
#define TRL_RESET_XS_ARG
#endif
#ifndef TRL_REUSE_KIN
#define TRL_REUSE_KIN 0    // 1 (experiment): the first physics sub-step of a launch reuses the kinematics the controller half of the same
#endif                     // launch computed for the same state (one of six kinematics() per env-step); same inputs, bit-identical
#if TRL_REUSE_KIN
#define TRL_REUSE_KIN_DECL , const Kin* k0
#else
#define TRL_REUSE_KIN_DECL
#endif
#if TRL_SMEM_ALIGNED
#define TRL_PHYS_XS_DECL , double* xs
#define TRL_PHYS_XS_ARG , xs
#else
#define TRL_PHYS_XS_DECL
#define TRL_PHYS_XS_ARG
#endif

// 128-bit shared-memory accesses of the experiment builds need 16-byte aligned addresses (a misaligned one faults on the GPU);
// the emulator build checks every such cast, the CUDA build compiles to the plain cast
#ifdef TRL_SIMT_EMU
template <typename T2, typename T>
inline T2* trl_as2(T* p) {
    if (reinterpret_cast<uintptr_t>(p) & 15u) { std::fprintf(stderr, "simt: misaligned 128-bit shared-memory access\n"); std::abort(); }
    return reinterpret_cast<T2*>(p);
}
#else
template <typename T2, typename T>
__device__ __forceinline__ T2* trl_as2(T* p) { return reinterpret_cast<T2*>(p); }
#endif

__device__ __forceinline__ int tri(int a, int b) { return a * (a + 1) / 2 + b; }  // a >= b
__device__ __forceinline__ double shf(double v, int src) { return __shfl_sync(kFull, v, src); }
__device__ __forceinline__ int shfi(int v, int src) { return __shfl_sync(kFull, v, src); }
__device__ __forceinline__ double warp_sum_all(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

__device__ __forceinline__ double wrap_pi(double a) {
    double s, c;
    sincos(a, &s, &c);
    double th = acos(fmin(1.0, fmax(-1.0, c)));
    return (s >= 0) ? th : -th;
}

// dog / goat joint indices (sim/SimDog.h:11-36)
enum DogJoint {
    jRoot, jSpine0, jSpine1, jSpine2, jSpine3, jTorso, jNeck0, jNeck1, jHead, jTail0, jTail1, jTail2, jTail3,
    jShoulder, jElbow, jWrist, jFinger, jHip, jKnee, jAnkle, jToe
};
// raptor joint indices (sim/SimRaptor.h:11-33) and controller layout (sim/RaptorController.h:14-47)
enum RaptorJoint {
    rRoot, rSpine0, rSpine1, rSpine2, rSpine3, rHead, rTail0, rTail1, rTail2, rTail3, rTail4,
    rRightHip, rRightKnee, rRightAnkle, rRightToe, rLeftHip, rLeftKnee, rLeftAnkle, rLeftToe
};
enum { rsContact, rsDown, rsPassing, rsUp };
enum { rmTransTime, rmCv, rmCd, rmForceX, rmForceY };
enum { rpRootPitch, rpSpineCurve, rpStanceHip, rpStanceKnee, rpStanceAnkle, rpSwingHip, rpSwingKnee, rpSwingAnkle };
enum { sBackStance, sExtend, sFrontStance, sGather };
enum { mTransTime, mCv, mBackForceX, mBackForceY, mFrontForceX, mFrontForceY, mMiscMax };
enum { spSpineCurve, spShoulder, spElbow, spHip, spKnee, spAnkle, spMax };

// ------------------------------------------------------------------------------------------------ counter RNG
struct CounterRng {
    uint64_t key, ctr;
    __device__ static uint64_t mix(uint64_t z) {
        z += 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    __device__ void init(uint64_t seed, uint64_t stream, uint64_t c) { key = mix(seed ^ mix(stream)); ctr = c; }
    __device__ uint64_t next() { return mix(key + (ctr++) * 0xD1342543DE82EF95ull); }
    __device__ double uniform() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    __device__ int rand_int(int mn, int mx) {
        if (mn == mx) return mn;
        int r = (int)(next() >> 33);
        return mn + r % (mx - mn);
    }
    __device__ bool flip_coin() { return uniform() < 0.5; }
    __device__ double normal() {
        double u1 = 1.0 - uniform(), u2 = uniform();
        return sqrt(-2.0 * log(u1)) * cos(6.283185307179586476925 * u2);
    }
};

// ------------------------------------------------------------------------------------------------ global accessors
#ifndef TRL_FIELD_SMEM
#define TRL_FIELD_SMEM 0    // 1: the step kernel stages every per-env field of its env in shared memory (one batch of independent loads at
#endif                      //    the start, one batch of stores at the end) instead of ~50 dependent L2 round trips spread over the launch
struct Lane {
    double* sm;          // unused by the warp kernel (kept for the decision kernel's scalar helpers)
    int env, n;
    double* D;
    int* I;
#if TRL_FIELD_SMEM
    double* sd;          // staged copy of the env's f64 fields [D_NUM_FIELDS] (null: read / write the planes directly)
    int* si;             // staged copy of the env's i32 fields [I_NUM_FIELDS]
    __device__ __forceinline__ double& d(int f) { return sd ? sd[f] : D[(size_t)f * n + env]; }
    __device__ __forceinline__ int& i(int f) { return si ? si[f] : I[(size_t)f * n + env]; }
#else
    __device__ __forceinline__ double& d(int f) { return D[(size_t)f * n + env]; }
    __device__ __forceinline__ int& i(int f) { return I[(size_t)f * n + env]; }
#endif
};

__device__ __forceinline__ bool has_fallen(Lane& L, double root_theta) {
    return L.d(D_SUM_FALL) > 0.25 || L.i(I_FAIL_FALL_DIST) != 0 || fabs(wrap_pi(root_theta)) > 3.14159265358979323846 * 0.8;
}

// cDogController::SetStateParams (sim/DogController.cpp:1042-1054) / cRaptorController::SetStateParams
// (sim/RaptorController.cpp:1108-1125) for the given state
__device__ void set_state_params(Lane& L, int state) {
    const ModelConst& m = c_model;
    int base = D_PARAMS + m.misc_max + state * m.sp_max;
    if (TRL_IS_RAPTOR(m)) {
        const int st = L.i(I_STANCE) == 0 ? rRightHip : rLeftHip, sw = L.i(I_STANCE) == 0 ? rLeftHip : rRightHip;
        L.d(D_PD_TARGET + st) = L.d(base + rpStanceHip); L.d(D_PD_TARGET + st + 1) = L.d(base + rpStanceKnee);
        L.d(D_PD_TARGET + st + 2) = L.d(base + rpStanceAnkle);
        L.d(D_PD_TARGET + sw) = L.d(base + rpSwingHip); L.d(D_PD_TARGET + sw + 1) = L.d(base + rpSwingKnee);
        L.d(D_PD_TARGET + sw + 2) = L.d(base + rpSwingAnkle);
        return;
    }
    double sc = L.d(base + spSpineCurve);
    L.d(D_PD_TARGET + jSpine0) = sc; L.d(D_PD_TARGET + jSpine1) = sc; L.d(D_PD_TARGET + jSpine2) = sc;
    L.d(D_PD_TARGET + jSpine3) = sc; L.d(D_PD_TARGET + jTorso) = sc;
    L.d(D_PD_TARGET + jShoulder) = L.d(base + spShoulder);
    L.d(D_PD_TARGET + jElbow) = L.d(base + spElbow);
    L.d(D_PD_TARGET + jHip) = L.d(base + spHip);
    L.d(D_PD_TARGET + jKnee) = L.d(base + spKnee);
    L.d(D_PD_TARGET + jAnkle) = L.d(base + spAnkle);
}

// cDogController::CalcReward (sim/DogController.cpp:594-623)
__device__ double calc_reward(Lane& L, bool fallen) {
    double vel_r = 0.0, stum_r = 0.0;
    if (!fallen) {
        double ct = L.d(D_PREV_CYCLE_T);
        double avg_vel = L.d(D_PREV_DIST_X) / ct;
        double err = c_model.target_vel_x - avg_vel;
        vel_r = exp(-0.5 * err * err);
        double avg_st = L.d(D_PREV_STUMBLE) / ct;
        stum_r = 1.0 / (1.0 + 10.0 * avg_st);
        if (TRL_IS_RAPTOR(c_model) && avg_vel < 0.0) { vel_r = 0.0; stum_r = 0.0; }   // sim/RaptorController.cpp:583-587
    }
    return 0.8 * vel_r + 0.2 * stum_r;
}

// cScenarioExp::NewCycleUpdate (scenarios/ScenarioExp.cpp:209-243): finish the previous tuple, start the next.
// Cooperative: the row copies are spread over the warp, the scalar bookkeeping is done by lane 0.
__device__ void exp_new_cycle_update(Lane& L, const Buffers& B, bool fallen, int lane) {
    const int S = B.S, A = B.A;
    const double* s_end = B.poli_state + (size_t)L.env * S;
    double* s_beg = B.tuple_sbeg + (size_t)L.env * S;
    double* act = B.tuple_action + (size_t)L.env * kNumParams;
    int slot = -1;
    if (lane == 0 && L.i(I_CYCLE_COUNT) > 1) slot = atomicAdd(B.tuple_count, 1);
    slot = shfi(slot, 0);
    if (slot >= 0 && slot < B.tuple_cap) {
        double* row = B.tuples + (size_t)slot * (1 + S + A + S);
        for (int k = lane; k < S; k += kWarp) { row[1 + k] = s_beg[k]; row[1 + S + A + k] = s_end[k]; }
        for (int k = lane; k < A; k += kWarp) row[1 + S + k] = act[k];
        if (lane == 0) {
            unsigned flags = (unsigned)L.i(I_TUPLE_FLAGS);
            flags = fallen ? (flags | 1u) : (flags & ~1u);
            row[0] = calc_reward(L, fallen);
            B.tuple_flags[slot] = flags;
            B.tuple_env[slot] = L.env;
        }
    }
    __syncwarp();
    for (int k = lane; k < S; k += kWarp) s_beg[k] = s_end[k];
    for (int k = lane; k < A; k += kWarp) act[k] = (k == 0) ? (double)L.i(I_ACTION_ID) : L.d(D_PARAMS + c_model.opt_idx[k - 1]);
    if (lane == 0) {
        int ef = L.i(I_EXP_FLAGS);
        unsigned nf = 0;
        if (ef & 1) nf |= 2u;   // exp critic -> eFlagExpCritic (bit 1)
        if (ef & 2) nf |= 4u;   // exp actor  -> eFlagExpActor  (bit 2)
        L.i(I_TUPLE_FLAGS) = (int)nf;
        L.i(I_CYCLE_COUNT) += 1;
    }
    __syncwarp();
}

// cDogController::BlendCtrlParams / BuildBaseAction (+ cDogControllerMACE::AssignFragID); writes params, returns id
__device__ int build_base_action(Lane& L, CounterRng& rng, int a, double* params /*local[30]*/) {
    const ModelConst& m = c_model;
    int i0 = m.act_idx0[a], i1 = m.act_idx1[a];
    double blend = m.act_blend[a];
    for (int k = 0; k < m.n_params; ++k) {
        double p0 = m.ctrl_params[i0][k], p1 = m.ctrl_params[i1][k];
        if (k == mTransTime || k == mCv || (TRL_IS_RAPTOR(m) && k == rmCd)) { p0 = fabs(p0); p1 = fabs(p1); }
        params[k] = (1.0 - blend) * p0 + blend * p1;
    }
    int id = a;
    if (m.is_mace) {
        int nf = m.has_net ? m.n_frags : 0, frag = 0;
        if (nf > 0) {
            if (i0 >= nf && i1 >= nf) frag = rng.rand_int(0, nf);
            else if (i0 >= nf) frag = i1;
            else if (i1 >= nf) frag = i0;
            else {
                frag = rng.flip_coin() ? i0 : i1;
                int ncp = m.n_ctrl, copies = nf / ncp, rem = nf % ncp;
                if (frag < rem) ++copies;
                frag += rng.rand_int(0, copies) * ncp;
            }
        }
        id = frag;
    }
    return id;
}

// cTerrainRLCharController::ApplyAction + cDogController::{NewCycleUpdate, ApplyAction}: commit an action
__device__ void apply_action(Lane& L, int id, const double* params, double comx, double comy) {
    for (int k = 0; k < c_model.n_params; ++k) L.d(D_PARAMS + k) = params[k];
    L.d(D_PARAMS + mTransTime) = fabs(params[mTransTime]);
    L.d(D_PARAMS + mCv) = fabs(params[mCv]);
    if (TRL_IS_RAPTOR(c_model)) L.d(D_PARAMS + rmCd) = fabs(params[rmCd]);
    L.i(I_ACTION_ID) = id;
    L.d(D_PREV_CYCLE_T) = L.d(D_CUR_CYCLE_T); L.d(D_CUR_CYCLE_T) = 0.0;
    L.d(D_PREV_STUMBLE) = L.d(D_CUR_STUMBLE); L.d(D_CUR_STUMBLE) = 0.0;
    L.d(D_PREV_DIST_X) = comx - L.d(D_PREV_COM_X); L.d(D_PREV_DIST_Y) = comy - L.d(D_PREV_COM_Y);
    L.d(D_PREV_COM_X) = comx; L.d(D_PREV_COM_Y) = comy;
    L.i(I_STATE) = sBackStance;
    L.d(D_PHASE) = 0.0;
    set_state_params(L, sBackStance);
}

__device__ GroundView load_ground(Lane& L, const Buffers& B) {
    GroundView g;
    g.data = B.terrain + (size_t)L.env * 2 * kTerrainCap;
    g.n[0] = L.i(I_SEG_N0); g.n[1] = L.i(I_SEG_N1);
    g.min_x[0] = L.d(D_SEG_MINX0); g.min_x[1] = L.d(D_SEG_MINX1);
    g.flip = L.i(I_SEG_FLIP);
    g.rng_state = (uint32_t)L.i(I_TERRAIN_RNG);
    return g;
}
__device__ void store_ground(Lane& L, const GroundView& g) {
    L.i(I_SEG_N0) = g.n[0]; L.i(I_SEG_N1) = g.n[1];
    L.d(D_SEG_MINX0) = g.min_x[0]; L.d(D_SEG_MINX1) = g.min_x[1];
    L.i(I_SEG_FLIP) = g.flip;
    L.i(I_TERRAIN_RNG) = (int)g.rng_state;
}

__device__ CounterRng load_rng(Lane& L) {
    CounterRng r;
    uint64_t c = ((uint64_t)(uint32_t)L.i(I_RNG_CTR_HI) << 32) | (uint32_t)L.i(I_RNG_CTR_LO);
    r.init(c_model.rng_seed, (uint64_t)L.env, c);
    return r;
}
__device__ void store_rng(Lane& L, const CounterRng& r) {
    L.i(I_RNG_CTR_LO) = (int)(uint32_t)(r.ctr & 0xffffffffull);
    L.i(I_RNG_CTR_HI) = (int)(uint32_t)(r.ctr >> 32);
}

// ================================================================================================ warp context
// Per-lane constants of link `lane` and the env state held in registers.
#ifndef TRL_LINK_SMEM
#define TRL_LINK_SMEM 1     // 1: the per-lane link constants live in shared memory (one table per CTA) instead of ~28 registers per thread
                            //    (the register build spills them and reloads them inside the ABA rounds; + 1.5 % measured, profiles/step_kernel_r02_session2_ab.txt)
#endif
#if TRL_LINK_SMEM
enum { LF_AX, LF_AY, LF_MASS, LF_BAX, LF_BAY, LF_IZZ, LF_LIM_LO, LF_LIM_HI, LF_NUM };
enum { LI_PARENT, LI_DEPTH, LI_ACC_ROUND, LI_ANC1, LI_ANC2, LI_ANC4, LI_ANC8, LI_NUM };
// a field of the table as a value: every read is one conflict-free LDS (tables are [field][lane])
template <int F> struct LkD { const double* p; __device__ __forceinline__ operator double() const { return p[F * kWarp]; } };
template <int F> struct LkI { const int* p; __device__ __forceinline__ operator int() const { return p[F * kWarp]; } };
struct LkU { const unsigned long long* p; __device__ __forceinline__ operator unsigned long long() const { return *p; } };
struct LinkTables {
    double d[LF_NUM * kWarp];
    unsigned long long u[kWarp];
    int i[LI_NUM * kWarp];
};
struct LinkC {
    int act;            // lane < nj
    LkI<LI_PARENT> parent; LkI<LI_DEPTH> depth; LkI<LI_ACC_ROUND> acc_round;
    LkU acc_src;
    LkI<LI_ANC1> anc1; LkI<LI_ANC2> anc2; LkI<LI_ANC4> anc4; LkI<LI_ANC8> anc8;
    LkD<LF_AX> ax; LkD<LF_AY> ay;
    LkD<LF_MASS> mass; LkD<LF_BAX> bax; LkD<LF_BAY> bay; LkD<LF_IZZ> izz_c;
    LkD<LF_LIM_LO> lim_lo; LkD<LF_LIM_HI> lim_hi;
};
#define TRL_LINK_TABLES_DECL __shared__ LinkTables s_link; LinkTables* const link_tabs = &s_link;
#else
struct LinkTables;
struct LinkC {
    int act;            // lane < nj
    int parent, depth;
    int acc_round;                // inward-pass round in which this link is eliminated (-1 idle lane, acc_rounds root)
    unsigned long long acc_src;   // 5 bits per inward round: lane to receive from (31 = none)
    int anc1, anc2, anc4, anc8;   // 2^k-th ancestors (31 = none: the idle zero lane) for the pointer-jumping prefix sums
    double ax, ay;      // attach point in the parent's joint frame
    double mass, bax, bay, izz_c;
    double lim_lo, lim_hi;
};
#define TRL_LINK_TABLES_DECL LinkTables* const link_tabs = nullptr;
#endif
struct Kin {            // world-axes kinematics of link `lane` about O (the root joint position)
    double phi, cw, sw, rx, ry, w, vx, vy;   // joint frame rotation, joint origin, spatial velocity (w, vO)
    double cx, cy;                           // body COM
};
struct EnvRegs {
    double q, qd;            // lane j: joint angle / rate of link j (lane 0: root angle)
    double ox, oy, oxd, oyd; // root translation and its rate (same value in every lane)
    double tau;              // held joint torque of link j
};

// the link constants of lane `lane` as plain values (constant memory -> registers)
struct LinkVals {
    int act, parent, depth, acc_round, anc1, anc2, anc4, anc8;
    unsigned long long acc_src;
    double ax, ay, mass, bax, bay, izz_c, lim_lo, lim_hi;
};
__device__ __forceinline__ LinkVals link_values(int lane) {
    const ModelConst& m = c_model;
    LinkVals c;
    c.act = lane < m.nj;
    int j = c.act ? lane : 0;
    c.parent = (c.act && j > 0) ? LT(parent, j) : 0;
    c.depth = c.act ? LT(depth, j) : -1;
    c.acc_round = c.act ? LT(acc_round, j) : -1;
    c.acc_src = c.act ? LT(acc_src, j) : ~0ull;
    c.anc1 = c.act ? LT2(anc_pow, j, 0) : -1; c.anc2 = c.act ? LT2(anc_pow, j, 1) : -1;
    c.anc4 = c.act ? LT2(anc_pow, j, 2) : -1; c.anc8 = c.act ? LT2(anc_pow, j, 3) : -1;
    if (c.anc1 < 0) c.anc1 = kZeroLane;
    if (c.anc2 < 0) c.anc2 = kZeroLane;
    if (c.anc4 < 0) c.anc4 = kZeroLane;
    if (c.anc8 < 0) c.anc8 = kZeroLane;
    c.ax = LT(attach_x, j); c.ay = LT(attach_y, j);
    c.mass = c.act ? LT(mass, j) : 0.0;
    c.bax = LT(body_ax, j); c.bay = LT(body_ay, j);
    c.izz_c = c.act ? LT(izz_c, j) : 0.0;
    const int has_lim = c.act ? LT(has_limit, j) : 0;
    c.lim_lo = has_lim ? LT(lim_lo, j) : -INFINITY;    // no limit: never violated, so the limit force needs no branch
    c.lim_hi = has_lim ? LT(lim_hi, j) : INFINITY;
    return c;
}
// TRL_LINK_SMEM: the first warp of the CTA fills the table (the caller's __syncthreads() publishes it)
__device__ __forceinline__ void stage_link_tables(LinkTables* t) {
#if TRL_LINK_SMEM
    if (threadIdx.x < kWarp) {
        const int lane = threadIdx.x;
        const LinkVals v = link_values(lane);
        t->d[LF_AX * kWarp + lane] = v.ax; t->d[LF_AY * kWarp + lane] = v.ay; t->d[LF_MASS * kWarp + lane] = v.mass;
        t->d[LF_BAX * kWarp + lane] = v.bax; t->d[LF_BAY * kWarp + lane] = v.bay; t->d[LF_IZZ * kWarp + lane] = v.izz_c;
        t->d[LF_LIM_LO * kWarp + lane] = v.lim_lo; t->d[LF_LIM_HI * kWarp + lane] = v.lim_hi;
        t->u[lane] = v.acc_src;
        t->i[LI_PARENT * kWarp + lane] = v.parent; t->i[LI_DEPTH * kWarp + lane] = v.depth; t->i[LI_ACC_ROUND * kWarp + lane] = v.acc_round;
        t->i[LI_ANC1 * kWarp + lane] = v.anc1; t->i[LI_ANC2 * kWarp + lane] = v.anc2; t->i[LI_ANC4 * kWarp + lane] = v.anc4;
        t->i[LI_ANC8 * kWarp + lane] = v.anc8;
    }
#else
    (void)t;
#endif
}
__device__ __forceinline__ LinkC load_link(int lane, const LinkTables* t) {
    LinkC c;
#if TRL_LINK_SMEM
    c.act = lane < c_model.nj;
    const double* d = t->d + lane;
    const int* i = t->i + lane;
    c.parent.p = i; c.depth.p = i; c.acc_round.p = i; c.anc1.p = i; c.anc2.p = i; c.anc4.p = i; c.anc8.p = i;
    c.acc_src.p = t->u + lane;
    c.ax.p = d; c.ay.p = d; c.mass.p = d; c.bax.p = d; c.bay.p = d; c.izz_c.p = d; c.lim_lo.p = d; c.lim_hi.p = d;
#else
    (void)t;
    const LinkVals v = link_values(lane);
    c.act = v.act; c.parent = v.parent; c.depth = v.depth; c.acc_round = v.acc_round; c.acc_src = v.acc_src;
    c.anc1 = v.anc1; c.anc2 = v.anc2; c.anc4 = v.anc4; c.anc8 = v.anc8;
    c.ax = v.ax; c.ay = v.ay; c.mass = v.mass; c.bax = v.bax; c.bay = v.bay; c.izz_c = v.izz_c; c.lim_lo = v.lim_lo; c.lim_hi = v.lim_hi;
#endif
    return c;
}

// root-ward prefix sums over the kinematic tree by pointer jumping: after 4 rounds every link holds the sum of its own
// value and those of all its ancestors (depth <= 15), using the static 2^k-th ancestor table.  A missing ancestor points
// at lane 31, which is idle and holds zeros, so the adds need no predicate.
#if TRL_KIN_SMEM
// the same rounds through shared memory: every lane stores its pair, the receiver loads the pair of its 2^r-th ancestor.
// Rounds alternate between two buffers, so one __syncwarp per round suffices (a buffer is rewritten two rounds later, and
// the barrier of the round in between orders that write after every read).  `xs`, `lane` must be in scope.
#define TRL_TREE_PREFIX2(a, b)                                                                  \
    do {                                                                                        \
        _Pragma("unroll") for (int r_ = 0; r_ < 4; ++r_) {                                      \
            const int an_ = (r_ == 0) ? kin_anc1 : ((r_ == 1) ? kin_anc2 : ((r_ == 2) ? kin_anc4 : kin_anc8));                 \
            double2* buf_ = trl_as2<double2>(xs + X_PFX + (r_ & 1) * 2 * kWarp);      \
            buf_[lane] = make_double2((a), (b));                                                \
            __syncwarp();                                                                       \
            const double2 t_ = buf_[an_];                                                       \
            (a) += t_.x; (b) += t_.y;                                                           \
        }                                                                                       \
    } while (0)
#else
#define TRL_TREE_PREFIX2(a, b)                                                                  \
    do {                                                                                        \
        _Pragma("unroll") for (int r_ = 0; r_ < 4; ++r_) {                                      \
            const int an_ = (r_ == 0) ? kin_anc1 : ((r_ == 1) ? kin_anc2 : ((r_ == 2) ? kin_anc4 : kin_anc8));                 \
            double ta_ = shf((a), an_), tb_ = shf((b), an_);                                    \
            (a) += ta_; (b) += tb_;                                                             \
        }                                                                                       \
    } while (0)

#endif

// outward kinematics: world rotation, joint origin (rel. O) and spatial velocity of every link
__device__ __forceinline__ Kin kinematics(const LinkC& c, const EnvRegs& e TRL_KIN_XS_DECL) {
    Kin k;
    // the link's table entries are read once per call (they live in shared memory; a read cannot move across the shuffles below)
    const int kin_anc1 = c.anc1, kin_anc2 = c.anc2, kin_anc4 = c.anc4, kin_anc8 = c.anc8, kin_parent = c.parent, kin_depth = c.depth;
    double phi = e.q, w = e.qd;
    TRL_TREE_PREFIX2(phi, w);
    k.phi = phi; k.w = w;
    sincos(phi, &k.sw, &k.cw);
    // offset of this joint from its parent's joint, in world axes
#if TRL_KIN_SMEM
    double pcw, psw;
    {
        double2* buf = trl_as2<double2>(xs + X_PFX + 2 * 2 * kWarp);
        buf[lane] = make_double2(k.cw, k.sw);
        __syncwarp();
        const double2 t = buf[kin_parent];
        pcw = t.x; psw = t.y;
    }
#else
    double pcw = shf(k.cw, kin_parent), psw = shf(k.sw, kin_parent);
#endif
    double rx = 0.0, ry = 0.0;
    if (kin_depth > 0) { const double ax = c.ax, ay = c.ay; rx = pcw * ax - psw * ay; ry = psw * ax + pcw * ay; }
    TRL_TREE_PREFIX2(rx, ry);
    // v_j = v_root + sum over the chain of S_a qd_a,  S_a = (1, r_ay, -r_ax)
    double vx = (kin_depth > 0) ? e.qd * ry : e.oxd, vy = (kin_depth > 0) ? -e.qd * rx : e.oyd;
    if (kin_depth < 0) { vx = 0.0; vy = 0.0; }
    TRL_TREE_PREFIX2(vx, vy);
    k.rx = rx; k.ry = ry; k.vx = vx; k.vy = vy;
    const double bax = c.bax, bay = c.bay;
    k.cx = rx + k.cw * bax - k.sw * bay;
    k.cy = ry + k.sw * bax + k.cw * bay;
    return k;
}

// child -> parent hand-off at the end of inward round r: every lane adds the NV register values of the one lane its
// schedule names for this round (lane 31 = none; it is idle and its values are zero, so the add is unconditional)
#if TRL_ACCUM_SMEM
// Experiment (profiles/step_kernel_r01_source_phases.md): the same hand-off staged through shared memory -- every lane stores its
// NV values as 128-bit words into its slot, the receiver loads the slot of its source lane.  Same values, same additions, so
// the results are bit-identical to the shuffle form; ~10 memory instructions per round instead of ~4 per value.
// `xs` (the warp's shared-memory block) must be in scope.
#define TRL_ACCUM_ROUND(r, NV, vals)                                                        \
    do {                                                                                    \
        const int src_ = (int)(acc_src_v >> (5 * (r))) & 31;                                \
        double2* mine_ = trl_as2<double2>(xs + X_ACC + lane * kAccStride);        \
        _Pragma("unroll") for (int v_ = 0; v_ + 1 < (NV); v_ += 2) mine_[v_ / 2] = make_double2((vals)[v_], (vals)[v_ + 1]); \
        if ((NV) & 1) xs[X_ACC + lane * kAccStride + (NV) - 1] = (vals)[(NV) - 1];         \
        __syncwarp();                                                                       \
        const double2* from_ = trl_as2<const double2>(xs + X_ACC + src_ * kAccStride); \
        _Pragma("unroll") for (int v_ = 0; v_ + 1 < (NV); v_ += 2) {                        \
            const double2 t_ = from_[v_ / 2];                                               \
            (vals)[v_] += t_.x; (vals)[v_ + 1] += t_.y;                                     \
        }                                                                                   \
        if ((NV) & 1) (vals)[(NV) - 1] += xs[X_ACC + src_ * kAccStride + (NV) - 1];         \
        __syncwarp();                                                                       \
    } while (0)
#else
#define TRL_ACCUM_ROUND(r, NV, vals)                                                        \
    do {                                                                                    \
        const int src_ = (int)(acc_src_v >> (5 * (r))) & 31;                                \
        _Pragma("unroll") for (int v_ = 0; v_ < (NV); ++v_) (vals)[v_] += shf((vals)[v_], src_); \
    } while (0)
#endif

// ================================================================================================ controller half
// Returns the clamped joint torque of link `lane` (0 for the root and idle lanes).
__device__ double controller_torque(Lane& L, const LinkC& lc, const EnvRegs& e, const Kin& k, double* xs, int lane,
                                    double h, int contact) {
    const double lk_mass = lc.mass;      // read once (shared-memory table; a read cannot move across the shuffles below)
    const ModelConst& m = c_model;
    const int nj = m.nj, nd = m.ndof, md = m.max_depth;
    double* M = xs + X_M;

    // ---- body inertia about O in world axes, RNEA accelerations with qdd = 0 and the reference's Cj
    const double hx = lk_mass * k.cx, hy = lk_mass * k.cy, Io = lc.izz_c + lk_mass * (k.cx * k.cx + k.cy * k.cy);
    double alx, aly;
    {
        // root: a0 = -g + R0 * cj, cj from cRBDUtil::BuildCjPlanar with c = s = cos(theta_dot) (sim/RBDUtil.cpp:821-822)
        double thd = shf(e.qd, 0), c0 = shf(k.cw, 0), s0 = shf(k.sw, 0);
        double cb = cos(thd);
        double cjx = (-cb * e.oxd + cb * e.oyd) * thd, cjy = (-cb * e.oxd - cb * e.oyd) * thd;
        alx = -m.gx + c0 * cjx - s0 * cjy;
        aly = -m.gy + s0 * cjx + c0 * cjy;
    }
    const int ctl_parent = lc.parent, ctl_depth = lc.depth;
    for (int l = 1; l <= md; ++l) {
        double pax = shf(alx, ctl_parent), pay = shf(aly, ctl_parent);
        if (ctl_depth == l) {
            alx = pax + e.qd * (k.w * k.rx + k.vy);   // a_j = a_parent + v_j x (S_j qd_j)
            aly = pay + e.qd * (k.w * k.ry - k.vx);
        }
    }
    // vals: composite inertia (Io, hx, hy, m) and subtree force (n, fx, fy)
    double vals[7];
    {
        double hlx = -hy * k.w + lk_mass * k.vx, hly = hx * k.w + lk_mass * k.vy;
        vals[0] = Io; vals[1] = hx; vals[2] = hy; vals[3] = lk_mass;
        vals[4] = (-hy * alx + hx * aly) + (k.vx * hly - k.vy * hlx);
        vals[5] = lk_mass * alx - k.w * hly;
        vals[6] = lk_mass * aly + k.w * hlx;
        if (!lc.act) { vals[4] = vals[5] = vals[6] = 0.0; }
    }
    {
        const unsigned long long acc_src_v = lc.acc_src;       // one read of the schedule word, not one per round
        for (int r = 0; r < m.acc_rounds; ++r) TRL_ACCUM_ROUND(r, 7, vals);
    }
    const double s1 = k.ry, s2 = -k.rx;                       // S_j = (1, s1, s2)
    // bias force C (per link lane; the root's three entries live in lane 0)
    const double Cj = vals[4] + s1 * vals[5] + s2 * vals[6];
    const double C0 = shf(vals[5], 0), C1 = shf(vals[6], 0), C2 = shf(vals[4], 0);

    // ---- CRBA: M[i][a] = S_a . (Ic_i S_i) for every ancestor a of i; lower triangle in shared memory
    for (int t = lane; t < kTri; t += kWarp) M[t] = 0.0;
    __syncwarp();
    {
        const double Fn = vals[0] - vals[2] * s1 + vals[1] * s2;      // Ic = [[I, -hy, hx], [-hy, m, 0], [hx, 0, m]]
        const double Fx = -vals[2] + vals[3] * s1, Fy = vals[1] + vals[3] * s2;
        const int dj = lane + 2;
        bool walking = lc.act && lane > 0;
        if (walking) M[tri(dj, dj)] = Fn + s1 * Fx + s2 * Fy;
        int cur = lc.parent;
        for (int it = 0; it < md; ++it) {
            double cry = shf(k.ry, cur), crx = shf(k.rx, cur);
            int cpar = shfi(lc.parent, cur);
            if (walking) {
                if (cur > 0) {
                    M[tri(dj, cur + 2)] = Fn + cry * Fx - crx * Fy;
                    cur = cpar;
                } else {
                    M[tri(dj, 0)] = Fx; M[tri(dj, 1)] = Fy; M[tri(dj, 2)] = Fn;
                    walking = false;
                }
            }
        }
        if (lane == 0) {
            M[tri(0, 0)] = vals[3]; M[tri(1, 1)] = vals[3];
            M[tri(2, 0)] = -vals[2]; M[tri(2, 1)] = vals[1]; M[tri(2, 2)] = vals[0];
        }
    }
    __syncwarp();

    // ---- ApplyFeedback (sim/DogController.cpp:903-945) / ApplySwingFeedback (sim/RaptorController.cpp:907-931)
    const int state = L.i(I_STATE);
    const bool raptor = TRL_IS_RAPTOR(m);
    const int stance = raptor ? L.i(I_STANCE) : 0;
    const int st_hip = stance == 0 ? rRightHip : rLeftHip, sw_hip = stance == 0 ? rLeftHip : rRightHip, st_toe = st_hip + 3;
    // cRaptorController::IsActiveVFEffector(stance toe): stance foot on the ground during Contact / Down
    const bool active_vf = raptor && (state == rsContact || state == rsDown) && ((contact >> st_toe) & 1);
    {
        double bvx = lk_mass * (k.vx - k.w * k.cy);
        double comvx = warp_sum_all(bvx) / m.total_mass;
        if (raptor) {
            double comx = warp_sum_all(lk_mass * k.cx) / m.total_mass;
            double toe_x = shf(k.cx, st_toe);
            if (lane == 0) {
                bool first_half = state == rsContact || state == rsDown;
                double cd = first_half ? 0.0 : L.d(D_PARAMS + rmCd), cv = first_half ? L.d(D_PARAMS + rmCv) : 0.0;
                int base = D_PARAMS + m.misc_max + state * m.sp_max;
                L.d(D_PD_TARGET + sw_hip) = L.d(base + rpSwingHip) + (cd * (comx - toe_x) + cv * comvx);
            }
        } else if (lane == 0) {
            double cv = L.d(D_PARAMS + mCv);
            int base = D_PARAMS + mMiscMax + state * spMax;
            if (!((contact >> jToe) & 1)) L.d(D_PD_TARGET + jHip) = L.d(base + spHip) + comvx * cv;
            if (!((contact >> jFinger) & 1)) L.d(D_PD_TARGET + jShoulder) = L.d(base + spShoulder) + comvx * cv;
        }
        __syncwarp();
    }

    // ---- cImpPDController::CalcControlForces (sim/ImpPDController.cpp:234-278); the raptor's stance hip PD is
    // switched off while its foot is an active virtual-force effector (UpdateStanceHip, sim/RaptorController.cpp:899-905):
    // zero gains in the torque law, but its Kd stays on the diagonal of the solved system
    double tau0 = 0.0, rhs_link = 0.0, kd_link = 0.0, kd_eff = 0.0;
    if (lc.act && lane > 0) {
        double theta = e.q;
        if (LT(world_pd, lane)) {
            // child body's world rotation, wrapped (cPDController::CalcTheta, sim/PDController.cpp:181-198)
            double c = k.cw * LT(body_cos, lane) - k.sw * LT(body_sin, lane);
            double s = k.sw * LT(body_cos, lane) + k.cw * LT(body_sin, lane);
            double a = acos(fmin(1.0, fmax(-1.0, c)));
            theta = (s >= 0) ? a : -a;
        }
        const bool pd_on = !(active_vf && lane == st_hip);
        kd_link = LT(kd, lane);
        kd_eff = pd_on ? kd_link : 0.0;
        double perr = L.d(D_PD_TARGET + lane) - theta, verr = LT(target_vel, lane) - e.qd;
        tau0 = pd_on ? (LT(kp, lane) * (perr - h * e.qd) + kd_link * verr) : 0.0;
        rhs_link = tau0 - Cj;
    }
    // dof-lane view: lane d holds row d of (M + h Kd) and rhs_d
    double rhs = shf(rhs_link, lane >= 2 ? lane - 2 : 0);
    double kdd = shf(kd_link, lane >= 2 ? lane - 2 : 0);
    if (lane == 0) { rhs = -C0; kdd = 0.0; }
    else if (lane == 1) { rhs = -C1; kdd = 0.0; }
    else if (lane == 2) { rhs = -C2; kdd = 0.0; }
    else if (lane >= nd) { rhs = 0.0; kdd = 0.0; }
    double row[kMaxDof];
#pragma unroll
    for (int c = 0; c < kMaxDof; ++c) row[c] = (lane < nd && c <= lane) ? M[tri(lane < nd ? lane : 0, c <= lane ? c : 0)] : 0.0;
    {
        double dadd = h * kdd;
#pragma unroll
        for (int c = 0; c < kMaxDof; ++c) if (c == lane) row[c] += dadd;
    }
    // register-resident LDL^T: pivot broadcast by shuffle, trailing update predicated on lane >= column
    double diag = 1.0;
#if TRL_LDLT_SMEM
    // Experiment: the pivot column travels through shared memory -- every lane stores its A_ij once, all lanes read the column
    // entries as broadcast 128-bit loads -- and the forward substitution rides along (y_j is final at step j).  Same
    // operations in the same order per lane as the shuffle form below, so the results are bit-identical.
    double y = rhs;
#pragma unroll
    for (int j = 0; j < kMaxDof; ++j) {
        double* col = xs + X_COL + (j & 1) * kColStride;
        col[lane] = row[j];
        if (lane == j) col[kWarp] = y;
        __syncwarp();
        double dj = col[j];
        if (j >= nd) dj = 1.0;
        const double inv = 1.0 / dj;
        if (lane == j) diag = dj;
        const double lij = row[j] * inv;
#pragma unroll
        for (int c0 = (j + 1) & ~1; c0 < kMaxDof; c0 += 2) {
            const double2 t = *trl_as2<const double2>(col + c0);
            if (c0 > j && lane >= c0) row[c0] -= lij * t.x;
            if (c0 + 1 < kMaxDof && lane >= c0 + 1) row[c0 + 1] -= lij * t.y;
        }
        const double yj = col[kWarp];
        if (lane > j) { y -= lij * yj; row[j] = lij; }
    }
    y /= diag;
#else
#pragma unroll
    for (int j = 0; j < kMaxDof; ++j) {
        double dj = shf(row[j], j);
        if (j >= nd) dj = 1.0;            // padding rows of a smaller character (raptor: 21 dof)
        double inv = 1.0 / dj;
        if (lane == j) diag = dj;
        double aij = row[j];              // A_ij before scaling (valid for lanes i > j)
        double lij = aij * inv;
#pragma unroll
        for (int c = j + 1; c < kMaxDof; ++c) {
            double akj = shf(aij, c);     // A_cj held by lane c
            if (lane >= c) row[c] -= lij * akj;
        }
        if (lane > j) row[j] = lij;
    }
    // forward substitution L y = rhs, then D, then L^T x = y (L^T read back from shared memory)
    double y = rhs;
#pragma unroll
    for (int c = 0; c < kMaxDof; ++c) {
        double yc = shf(y, c);
        if (lane > c) y -= row[c] * yc;
    }
    y /= diag;
#endif
#pragma unroll
    for (int c = 0; c < kMaxDof; ++c) if (lane < nd && c < lane) M[tri(lane, c)] = row[c];
    __syncwarp();
    double acc = y;
    for (int c = nd - 1; c >= 1; --c) {
        double xc = shf(acc, c);
        if (lane < c) acc -= M[tri(c, lane)] * xc;
    }
    double tau = 0.0;
    {
        double acc_link = shf(acc, lane + 2 < kWarp ? lane + 2 : 0);
        if (lc.act && lane > 0) tau = tau0 - kd_eff * h * acc_link;
    }

    // ---- ApplyGravityCompensation (sim/DogController.cpp:947-995, 1120-1175; sim/RaptorController.cpp:983-1026)
    // effector 0 / 1: dog back foot (toe) / front foot (finger); raptor right / left toe.  A dog foot supports when
    // it touches the ground, a raptor foot only while it is the active virtual-force effector.
    const int eff0 = raptor ? (int)rRightToe : (int)jToe, eff1 = raptor ? (int)rLeftToe : (int)jFinger;
    const bool toe_c = raptor ? (active_vf && stance == 0) : (((contact >> jToe) & 1) != 0);
    const bool fin_c = raptor ? (active_vf && stance == 1) : (((contact >> jFinger) & 1) != 0);
    // foot bottom-centre positions (GetEndEffectorContactPos), relative to O
    double ex[2], ey[2];
    {
        double bc = k.cw * LT(body_cos, lc.act ? lane : 0) - k.sw * LT(body_sin, lc.act ? lane : 0);
        double bs = k.sw * LT(body_cos, lc.act ? lane : 0) + k.cw * LT(body_sin, lc.act ? lane : 0);
        double ly = -LT(half_y, lc.act ? lane : 0);
        double px = k.cx - bs * ly, py = k.cy + bc * ly;
        ex[0] = shf(px, eff0); ey[0] = shf(py, eff0);
        ex[1] = shf(px, eff1); ey[1] = shf(py, eff1);
    }
    if (m.grav_comp && (toe_c || fin_c)) {
        // tau_g = -G,  G_k = (sum_subtree m (c - r_k)) x g   (generalised gravity force, sim/RBDUtil.cpp:850-895)
        const double ms = vals[3], mhx = vals[1], mhy = vals[2];
        double tg;
        if (lane == 0) tg = 0.0;
        else tg = -((mhx - ms * k.rx) * m.gy - (mhy - ms * k.ry) * m.gx);
        const double ms0 = shf(ms, 0), mhx0 = shf(mhx, 0), mhy0 = shf(mhy, 0);
        const double b0 = -(ms0 * m.gx), b1 = -(ms0 * m.gy), b2 = -(mhx0 * m.gy - mhy0 * m.gx);
        // root-row weights of the ridge least squares: identity for the dog, (1e-4, 1e-4, 1) for the raptor
        const double W0 = raptor ? 0.0001 : 1.0, W1 = raptor ? 0.0001 : 1.0, W2 = 1.0;
        // basis columns [eff0 +y, eff0 +x, eff1 +y, eff1 +x]; root rows = (Fx, Fy, r x F)
        double Ar[3][4];
#pragma unroll
        for (int e2 = 0; e2 < 2; ++e2) {
            bool on = e2 == 0 ? toe_c : fin_c;
            Ar[0][2 * e2] = 0.0;            Ar[0][2 * e2 + 1] = on ? 1.0 : 0.0;
            Ar[1][2 * e2] = on ? 1.0 : 0.0; Ar[1][2 * e2 + 1] = 0.0;
            Ar[2][2 * e2] = on ? ex[e2] : 0.0; Ar[2][2 * e2 + 1] = on ? -ey[e2] : 0.0;
        }
        double A[4][5];
#pragma unroll
        for (int a = 0; a < 4; ++a) {
#pragma unroll
            for (int b = 0; b < 4; ++b) A[a][b] = Ar[0][a] * W0 * Ar[0][b] + Ar[1][a] * W1 * Ar[1][b] + Ar[2][a] * W2 * Ar[2][b];
            A[a][a] += 0.0001;
            A[a][4] = Ar[0][a] * W0 * b0 + Ar[1][a] * W1 * b1 + Ar[2][a] * W2 * b2;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            int piv = c;
#pragma unroll
            for (int r = c + 1; r < 4; ++r) if (fabs(A[r][c]) > fabs(A[piv][c])) piv = r;
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if (r == piv && piv != c) {
#pragma unroll
                    for (int kk = 0; kk < 5; ++kk) { double t = A[c][kk]; A[c][kk] = A[r][kk]; A[r][kk] = t; }
                }
            }
#pragma unroll
            for (int r = c + 1; r < 4; ++r) {
                double f = A[r][c] / A[c][c];
#pragma unroll
                for (int kk = c; kk < 5; ++kk) A[r][kk] -= f * A[c][kk];
            }
        }
        double x[4];
#pragma unroll
        for (int r = 3; r >= 0; --r) {
            double v = A[r][4];
#pragma unroll
            for (int kk = r + 1; kk < 4; ++kk) v -= A[r][kk] * x[kk];
            x[r] = v / A[r][r];
        }
        if (lc.act && lane > 0) {
            double tc = 0.0;
            if (toe_c && ((m.anc_mask_toe >> lane) & 1)) { double rx = ex[0] - k.rx, ry = ey[0] - k.ry; tc += rx * x[0] - ry * x[1]; }
            if (fin_c && ((m.anc_mask_finger >> lane) & 1)) { double rx = ex[1] - k.rx, ry = ey[1] - k.ry; tc += rx * x[2] - ry * x[3]; }
            tau += tg - tc;
        }
    }

    // ---- ApplyStanceFeedback (sim/RaptorController.cpp:933-981): the stance hip balances the swing hip's torque and
    // a PD on the root pitch while its foot pushes
    if (raptor) {
        const double sw_tau = shf(tau, sw_hip), th0 = shf(e.q, 0), thd0 = shf(e.qd, 0);
        if (active_vf && lane == st_hip) {
            int base = D_PARAMS + m.misc_max + state * m.sp_max;
            double root_tau = m.kp[st_hip] * (L.d(base + rpRootPitch) - wrap_pi(th0)) + m.kd[st_hip] * (-thd0);
            tau += -sw_tau - root_tau;
        }
    }

    // ---- ApplyVirtualForces (sim/DogController.cpp:997-1029; sim/RaptorController.cpp:1028-1075)
    if (m.virt_forces && raptor) {
        // only the stance toe can be active; the chain runs up to (excluding) the root and the stance hip's share is
        // mirrored onto the swing hip
        const bool on = active_vf;
        const int e2 = stance;
        double fx = -L.d(D_PARAMS + rmForceX), fy = -L.d(D_PARAMS + rmForceY);
        unsigned mask = e2 == 0 ? m.vf_mask_toe : m.vf_mask_finger;
        double t = 0.0;
        if (on && lc.act && ((mask >> lane) & 1)) t = (ex[e2] - k.rx) * fy - (ey[e2] - k.ry) * fx;
        double t_hip = shf(t, st_hip);
        tau += t;
        if (on && lane == sw_hip) tau -= t_hip;
    } else if (m.virt_forces && lc.act && lane > 0) {
        if ((state == sBackStance || state == sExtend) && toe_c && ((m.vf_mask_toe >> lane) & 1)) {
            double fx = -L.d(D_PARAMS + mBackForceX), fy = -L.d(D_PARAMS + mBackForceY);
            tau += (ex[0] - k.rx) * fy - (ey[0] - k.ry) * fx;
        }
        if ((state == sFrontStance || state == sGather) && fin_c && ((m.vf_mask_finger >> lane) & 1)) {
            double fx = -L.d(D_PARAMS + mFrontForceX), fy = -L.d(D_PARAMS + mFrontForceY);
            tau += (ex[1] - k.rx) * fy - (ey[1] - k.ry) * fx;
        }
    }

    // ---- cJoint::ApplyTorque clamp (sim/Joint.cpp:171-201,257-264)
    if (lc.act && lane > 0) {
        double lim = LT(torque_lim, lane);
        if (fabs(tau) > lim) tau *= lim / fabs(tau);
    } else tau = 0.0;
    (void)nj;
    return tau;
}

// ================================================================================================ physics half
// One sub-step of articulated-body forward dynamics with linearly-implicit contact / joint-limit terms.
// Updates e (q, qd, root translation) in place and returns the contact bitmask (same value in every lane).
__device__ int physics_substep(const LinkC& lc, EnvRegs& e, const GroundView& g, const double* s_clx, const double* s_cly,
                               const int* s_cbody, int lane, double dt, double clear_y TRL_PHYS_XS_DECL TRL_REUSE_KIN_DECL) {
    const double lk_mass = lc.mass;      // read once (shared-memory table; a read cannot move across the shuffles below)
    // per-body box tables (link frame): centre offset, axis rotation, half sizes -- staged by the kernel next to the corner tables
    const double *s_bax = s_clx + 4 * kMaxJoints, *s_bay = s_bax + kMaxJoints, *s_bc = s_bay + kMaxJoints, *s_bs = s_bc + kMaxJoints,
                 *s_hx = s_bs + kMaxJoints, *s_hy = s_hx + kMaxJoints;
    const ModelConst& m = c_model;
    const PhysParams& pp = m.phys;
    const int md = m.max_depth;
#if TRL_REUSE_KIN
    Kin k;
    if (k0) k = *k0; else k = kinematics(lc, e TRL_KIN_XS_ARG);
#else
    Kin k = kinematics(lc, e TRL_KIN_XS_ARG);
#endif

    // rigid inertia about O, bias force incl. gravity as an external force
    const double hx = lk_mass * k.cx, hy = lk_mass * k.cy, Io = lc.izz_c + lk_mass * (k.cx * k.cx + k.cy * k.cy);
    double ia[9];   // articulated inertia (a, bx, by, cxx, cxy, cyy) and bias force (n, fx, fy)
    {
        double hlx = -hy * k.w + lk_mass * k.vx, hly = hx * k.w + lk_mass * k.vy;
        ia[0] = Io; ia[1] = -hy; ia[2] = hx; ia[3] = lk_mass; ia[4] = 0.0; ia[5] = lk_mass;
        ia[6] = (k.vx * hly - k.vy * hlx) - (hx * m.gy - hy * m.gx);
        ia[7] = -k.w * hly - lk_mass * m.gx;
        ia[8] = k.w * hlx - lk_mass * m.gy;
        if (!lc.act) { ia[6] = ia[7] = ia[8] = 0.0; }
    }
    const double cvx = e.qd * (k.w * k.rx + k.vy), cvy = e.qd * (k.w * k.ry - k.vx);   // c_j = v_j x (S_j qd_j)

    // ---- contacts: one box corner per lane per round against the height field
    int contact = 0;
    const int nc = m.n_corners;
#if TRL_CONTACT_SMEM
    // per-link kinematics staged once per sub-step; a corner lane reads the entries of the body it belongs to
    double2* kin0 = trl_as2<double2>(xs + X_KIN);                  // (cw, sw)
    double2* kin1 = trl_as2<double2>(xs + X_KIN + 2 * kWarp);      // (rx, ry)
    double2* kin2 = trl_as2<double2>(xs + X_KIN + 4 * kWarp);      // (w, vx)
    double* kin3 = xs + X_KIN + 6 * kWarp;                                   // vy
    kin0[lane] = make_double2(k.cw, k.sw); kin1[lane] = make_double2(k.rx, k.ry);
    kin2[lane] = make_double2(k.w, k.vx); kin3[lane] = k.vy;
    __syncwarp();
#endif
    // one compliant contact at the point (rpx, rpy) rel. O of a body moving with (bw, bvx, bvy): penetration `pen` along the unit
    // direction (nx, ny) the force pushes the body in; implicit in the point velocity (adds dt X^T D X to the body's articulated
    // inertia and X^T w to its bias force, D = cnn n n^T + ctt t t^T)
    auto contact_force = [&](double rpx, double rpy, double pen, double nx, double ny, double bw, double bvx, double bvy, double* add) {
        const double tx = ny, ty = -nx;
        const double pvx = bvx - bw * rpy, pvy = bvy + bw * rpx;
        const double vn = pvx * nx + pvy * ny, vt = pvx * tx + pvy * ty;
        const double fn0 = pp.kn * pen - pp.dn * vn;
        if (fn0 > 0.0) {
            const double cnn = pp.dn + dt * pp.kn;
            const double ctt = pp.mu * fn0 / fmax(fabs(vt), pp.v_eps);
            const double fwx = pp.kn * pen * nx - cnn * vn * nx - ctt * vt * tx;
            const double fwy = pp.kn * pen * ny - cnn * vn * ny - ctt * vt * ty;
            const double dxx = dt * (cnn * nx * nx + ctt * tx * tx), dxy = dt * (cnn * nx * ny + ctt * tx * ty),
                         dyy = dt * (cnn * ny * ny + ctt * ty * ty);
            const double kx = -rpy, ky = rpx;
            const double dkx = dxx * kx + dxy * ky, dky = dxy * kx + dyy * ky;
            add[0] = kx * dkx + ky * dky; add[1] = dkx; add[2] = dky; add[3] = dxx; add[4] = dxy; add[5] = dyy;
            add[6] = -(rpx * fwy - rpy * fwx); add[7] = -fwx; add[8] = -fwy;
        }
    };
    // rounds 0 .. nc/32: box corners against the height field.  rounds after that (pp.vertex_contacts): the terrain vertices inside a
    // box -- lane (body, slot) looks at the slot-th vertex under the body's x extent (a box spans at most 4 vertices); where the surface
    // is convex and the vertex lies inside, it is pushed out through the nearest face (sim/GroundVar2D.cpp:440-448 hands Bullet a
    // height field: an edge of the terrain can enter a box between two of its corners)
    const int nrounds = (nc + kWarp - 1) / kWarp;
    for (int round = 0; round < (pp.vertex_contacts ? 2 * nrounds : nrounds); ++round) {
        const bool vtx = round >= nrounds;
        const int base = (vtx ? round - nrounds : round) * kWarp;
        const int ci = base + lane;
        const bool valid = ci < nc;
        const int b = valid ? s_cbody[ci] : 0;
        if (vtx) {
            const double bcw = shf(k.cw, b), bsw = shf(k.sw, b), brx = shf(k.rx, b), bry = shf(k.ry, b);
            const double bax = s_bax[b], bay = s_bay[b], bc = s_bc[b], bs = s_bs[b], hxb = s_hx[b], hyb = s_hy[b];
            const double C = bcw * bc - bsw * bs, S = bsw * bc + bcw * bs;                      // box axes in the world
            const double cxr = brx + bcw * bax - bsw * bay, cyr = bry + bsw * bax + bcw * bay;   // box centre rel. O
            const bool cand = valid && (e.oy + cyr - (fabs(S) * hxb + fabs(C) * hyb) <= clear_y);
            if (!__any_sync(kFull, cand)) continue;
            const double bw = shf(k.w, b), bvx = shf(k.vx, b), bvy = shf(k.vy, b);
            double add[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
            bool touching = false;
            if (cand) {
                const double extx = fabs(C) * hxb + fabs(S) * hyb, cxw = e.ox + cxr;
                double xv, hv, hp, hn;
                if (g.vertex_slot(cxw - extx - pp.contact_tol, cxw + extx + pp.contact_tol, ci & 3, &xv, &hv, &hp, &hn) && hv - 0.5 * (hp + hn) > 1e-9) {
                    const double rx = xv - cxw, ry = hv - (e.oy + cyr);
                    const double lx = C * rx + S * ry, ly = -S * rx + C * ry;
                    const double dx = hxb - fabs(lx), dy = hyb - fabs(ly);
                    if (dx > -pp.contact_tol && dy > -pp.contact_tol) {
                        touching = true;
                        if (dx > 0.0 && dy > 0.0) {
                            double nlx = 0.0, nly = 0.0, pen;
                            if (dx < dy) { nlx = lx > 0.0 ? -1.0 : 1.0; pen = dx; } else { nly = ly > 0.0 ? -1.0 : 1.0; pen = dy; }
                            contact_force(xv - e.ox, hv - e.oy, pen, C * nlx - S * nly, S * nlx + C * nly, bw, bvx, bvy, add);
                        }
                    }
                }
            }
            const unsigned tmask = __ballot_sync(kFull, touching);
            unsigned fmask = __ballot_sync(kFull, add[3] != 0.0 || add[5] != 0.0);
            const int cb = lc.act ? LT(corner_base, lane) : -1;
            const bool mine = cb >= base && cb < base + kWarp;
            while (fmask) {
                const int src = __ffs(fmask) - 1;
                fmask &= fmask - 1;
                const bool to_me = mine && (src >= cb - base) && (src < cb - base + 4);
#pragma unroll
                for (int v = 0; v < 9; ++v) {
                    double t = shf(add[v], src);
                    if (to_me) ia[v] += t;
                }
            }
            if (mine && ((tmask >> (cb - base)) & 0xfu)) contact |= 1 << lane;
            continue;
        }
        const double lx = valid ? s_clx[ci] : 0.0, ly = valid ? s_cly[ci] : 0.0;
#if TRL_CONTACT_SMEM
        const double2 k0_ = kin0[b], k1_ = kin1[b];
        const double bcw = k0_.x, bsw = k0_.y, brx = k1_.x, bry = k1_.y;
#else
        const double bcw = shf(k.cw, b), bsw = shf(k.sw, b), brx = shf(k.rx, b), bry = shf(k.ry, b);
#endif
        const double rpx = brx + bcw * lx - bsw * ly, rpy = bry + bsw * lx + bcw * ly;   // corner rel. O
        // broad phase: a corner above clear_y (terrain maximum under the character + margin) cannot be within the
        // contact tolerance; a round with no candidate corner is skipped by the whole warp
        const bool cand = valid && (e.oy + rpy <= clear_y);
        if (!__any_sync(kFull, cand)) continue;
#if TRL_CONTACT_SMEM
        const double2 k2_ = kin2[b];
        const double bw = k2_.x, bvx = k2_.y, bvy = kin3[b];
#else
        const double bw = shf(k.w, b), bvx = shf(k.vx, b), bvy = shf(k.vy, b);
#endif
        double add[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        bool touching = false;
        if (cand) {
            double slope;
            const double hgt = g.sample_fast(e.ox + rpx, &slope);
            const double inv = rsqrt(1.0 + slope * slope);
            const double pen = (hgt - (e.oy + rpy)) * inv;
            if (pen > -pp.contact_tol) {
                touching = true;
                if (pen > 0.0) contact_force(rpx, rpy, pen, -slope * inv, inv, bw, bvx, bvy, add);
            }
        }
        // hand the (few) force-producing corners to the lanes that own their bodies, in corner order (deterministic)
        const unsigned tmask = __ballot_sync(kFull, touching);
        unsigned fmask = __ballot_sync(kFull, add[3] != 0.0 || add[5] != 0.0);
        const int cb = lc.act ? LT(corner_base, lane) : -1;
        const bool mine = cb >= base && cb < base + kWarp;
#if TRL_CONTACT_SMEM
        if (fmask) {
            // force-producing corner lanes publish their 9 values; the lane that owns the body adds its (up to 4) corners in
            // corner order -- the order the shuffle loop below visits them in
            double* slot = xs + X_ACC + lane * kAccStride;
            if ((fmask >> lane) & 1u) {
                double2* s2 = trl_as2<double2>(slot);
                s2[0] = make_double2(add[0], add[1]); s2[1] = make_double2(add[2], add[3]); s2[2] = make_double2(add[4], add[5]);
                s2[3] = make_double2(add[6], add[7]); slot[8] = add[8];
            }
            __syncwarp();
            const unsigned my = mine ? ((fmask >> (cb - base)) & 0xfu) : 0u;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                if ((my >> q) & 1u) {
                    const double* from = xs + X_ACC + (cb - base + q) * kAccStride;
                    const double2* f2 = trl_as2<const double2>(from);
                    const double2 t0 = f2[0], t1 = f2[1], t2 = f2[2], t3 = f2[3];
                    ia[0] += t0.x; ia[1] += t0.y; ia[2] += t1.x; ia[3] += t1.y; ia[4] += t2.x; ia[5] += t2.y;
                    ia[6] += t3.x; ia[7] += t3.y; ia[8] += from[8];
                }
            }
            __syncwarp();
        }
#else
        while (fmask) {
            const int src = __ffs(fmask) - 1;
            fmask &= fmask - 1;
            const bool to_me = mine && (src >= cb - base) && (src < cb - base + 4);
#pragma unroll
            for (int v = 0; v < 9; ++v) {
                double t = shf(add[v], src);
                if (to_me) ia[v] += t;
            }
        }
#endif
        if (mine && ((tmask >> (cb - base)) & 0xfu)) contact |= 1 << lane;
    }
    contact = (int)__ballot_sync(kFull, contact != 0);   // lane index == link index

    // ---- pass 2: articulated inertias / bias forces inward (level-synchronous)
    const double s1 = k.ry, s2 = -k.rx;
    double U0 = 0.0, U1 = 0.0, U2 = 0.0, dinv = 0.0, uu = 0.0;
    // one-sided implicit spring-damper at the joint limits (branch-free: the range of a free joint is infinite).  It depends on
    // the joint's own state only, so it is evaluated once per sub-step, not once per elimination round
#ifndef TRL_HOIST_LIMITS
#define TRL_HOIST_LIMITS 1
#endif
    const unsigned long long acc_src_v = lc.acc_src;
#if TRL_HOIST_LIMITS
    const double viol = fmax(e.q - lc.lim_hi, 0.0) + fmin(e.q - lc.lim_lo, 0.0);
    const double cl = (viol != 0.0) ? pp.d_lim + dt * pp.k_lim : 0.0;
    const double lim_f = pp.k_lim * viol + cl * e.qd;
    const int my_round = lc.acc_round;
#else
#define my_round lc.acc_round
#endif
    for (int r = 0; r < m.acc_rounds; ++r) {
        if (my_round == r) {
            U0 = ia[0] + ia[1] * s1 + ia[2] * s2;
            U1 = ia[1] + ia[3] * s1 + ia[4] * s2;
            U2 = ia[2] + ia[4] * s1 + ia[5] * s2;
            double Dj = U0 + s1 * U1 + s2 * U2;
            uu = e.tau - (ia[6] + s1 * ia[7] + s2 * ia[8]);
#if !TRL_HOIST_LIMITS
            const double viol = fmax(e.q - lc.lim_hi, 0.0) + fmin(e.q - lc.lim_lo, 0.0);
            const double cl = (viol != 0.0) ? pp.d_lim + dt * pp.k_lim : 0.0;
            const double lim_f = pp.k_lim * viol + cl * e.qd;
#endif
            uu -= lim_f;
            Dj += dt * cl;
            dinv = 1.0 / Dj;
            // Ia = IA - U U^T / D ;  pa = pA + Ia c + U u / D
            double n0 = ia[0] - U0 * U0 * dinv, n1 = ia[1] - U0 * U1 * dinv, n2 = ia[2] - U0 * U2 * dinv;
            double n3 = ia[3] - U1 * U1 * dinv, n4 = ia[4] - U1 * U2 * dinv, n5 = ia[5] - U2 * U2 * dinv;
            double ud = uu * dinv;
            ia[6] += n1 * cvx + n2 * cvy + U0 * ud;
            ia[7] += n3 * cvx + n4 * cvy + U1 * ud;
            ia[8] += n4 * cvx + n5 * cvy + U2 * ud;
            ia[0] = n0; ia[1] = n1; ia[2] = n2; ia[3] = n3; ia[4] = n4; ia[5] = n5;
        }
        TRL_ACCUM_ROUND(r, 9, ia);
    }
    // floating base: solve IA_0 a_0 = -pA_0 (symmetric 3x3 LDL^T); every lane computes it from lane 0's values
    double a0, a1, a2;
    {
#if TRL_OUTWARD_SMEM
        double* bs = xs + X_BASE;
        if (lane == 0) {
            double2* b2 = trl_as2<double2>(bs);
            b2[0] = make_double2(ia[0], ia[1]); b2[1] = make_double2(ia[2], ia[3]); b2[2] = make_double2(ia[4], ia[5]);
            b2[3] = make_double2(ia[6], ia[7]); bs[8] = ia[8];
        }
        __syncwarp();
        const double2 q0 = trl_as2<const double2>(bs)[0], q1 = trl_as2<const double2>(bs)[1],
                      q2 = trl_as2<const double2>(bs)[2], q3 = trl_as2<const double2>(bs)[3];
        double a = q0.x, bx = q0.y, by = q1.x, cxx = q1.y, cxy = q2.x, cyy = q2.y;
        double r0 = -q3.x, r1 = -q3.y, r2 = -bs[8];
#else
        double a = shf(ia[0], 0), bx = shf(ia[1], 0), by = shf(ia[2], 0), cxx = shf(ia[3], 0), cxy = shf(ia[4], 0), cyy = shf(ia[5], 0);
        double r0 = -shf(ia[6], 0), r1 = -shf(ia[7], 0), r2 = -shf(ia[8], 0);
#endif
        const double i0 = 1.0 / a, l10 = bx * i0, l20 = by * i0;
        const double d1 = cxx - l10 * bx, i1 = 1.0 / d1, t21 = cxy - l20 * bx, l21 = t21 * i1;
        const double d2 = cyy - l20 * by - l21 * t21, i2 = 1.0 / d2;
        const double y0 = r0, y1 = r1 - l10 * y0, y2 = r2 - l20 * y0 - l21 * y1;
        a2 = y2 * i2; a1 = y1 * i1 - l21 * a2; a0 = y0 * i0 - l10 * a1 - l20 * a2;
    }
    // pass 3: accelerations outward; each link lane integrates its own joint (semi-implicit Euler)
    double aw = a0, alx = a1, aly = a2;
    const double thd0 = shf(e.qd, 0);
#if TRL_HOIST_LIMITS
    const int my_depth = lc.depth, my_parent = lc.parent;
#else
#undef my_round
#define my_depth lc.depth
#define my_parent lc.parent
#endif
    for (int l = 1; l <= md; ++l) {
#if TRL_OUTWARD_SMEM
        // every lane publishes its current (aw, ax, ay); a link of depth l reads its parent's, which is final since level l-1.
        // Levels alternate between two buffers: one barrier per level (see TRL_TREE_PREFIX2)
        double2* o2 = trl_as2<double2>(xs + X_OUT + (l & 1) * kOutBuf);
        double* o1 = xs + X_OUT + (l & 1) * kOutBuf + 2 * kWarp;
        o2[lane] = make_double2(aw, alx); o1[lane] = aly;
        __syncwarp();
        const double2 pp_ = o2[my_parent];
        const double paw = pp_.x, pax = pp_.y, pay = o1[my_parent];
#else
        double paw = shf(aw, my_parent), pax = shf(alx, my_parent), pay = shf(aly, my_parent);
#endif
        if (my_depth == l) {
            double bx = pax + cvx, by = pay + cvy;
            double qdd = (uu - (U0 * paw + U1 * bx + U2 * by)) * dinv;
            aw = paw + qdd; alx = bx + s1 * qdd; aly = by + s2 * qdd;
            e.qd += dt * qdd;
            e.q += dt * e.qd;
        }
    }
    {
        // root: classical acceleration of the origin = spatial acceleration + w x v
        double xdd = a1 - thd0 * e.oyd, ydd = a2 + thd0 * e.oxd;
        e.oxd += dt * xdd; e.oyd += dt * ydd;
        e.ox += dt * e.oxd; e.oy += dt * e.oyd;
        if (lane == 0) { e.qd += dt * a0; e.q += dt * e.qd; }
    }
#if !TRL_HOIST_LIMITS
#undef my_depth
#undef my_parent
#endif
    return contact;
}

// cTerrainRLCharController::ParseGround + BuildPoliState (sim/TerrainRLCharController.cpp:168-285), cooperative
__device__ void build_poli_state(const LinkC& lc, const EnvRegs& e, const Kin& k, const Buffers& B, const GroundView& g, int env,
                                 int lane, int stance) {
    const ModelConst& m = c_model;
    double* out = B.poli_state + (size_t)env * B.S;
    const double oy = g.sample(e.ox);
    for (int i = lane; i < kNumGroundSamples; i += kWarp) {
        double dist = ((10.0 - (-0.5)) * i) / (kNumGroundSamples - 1) + (-0.5);
        out[i] = g.sample(dist + e.ox) - oy;
    }
    if (lane == 0) out[kNumGroundSamples] = e.oy - oy;
    if (lc.act) {
        // raptor: when the left leg is the stance leg the two legs' entries are swapped
        // (cRaptorController::FlipPoliPoseStance, sim/RaptorController.cpp:1414-1432,1469-1487)
        int slot = lane;
        if (TRL_IS_RAPTOR(m) && stance != 0 && lane >= rRightHip) slot = lane < rLeftHip ? lane + 4 : lane - 4;
        if (lane > 0) {
            out[kNumGroundSamples + 1 + 2 * (slot - 1)] = k.cx;       // body COM relative to the root joint position
            out[kNumGroundSamples + 1 + 2 * (slot - 1) + 1] = k.cy;
        }
        int vo = kNumGroundSamples + 2 * m.nj - 1;
        out[vo + 2 * slot] = k.vx - k.w * k.cy;                       // body COM velocity
        out[vo + 2 * slot + 1] = k.vy + k.w * k.cx;
    }
}

// cScenarioSimChar::Reset (+ PoliEval / Exp specifics): scenarios/ScenarioSimChar.cpp:121-132.  Cooperative kinematics,
// scalar bookkeeping + terrain generation on lane 0.
__device__ void reset_env(Lane& L, const LinkC& lc, EnvRegs& e, const Buffers& B, int lane TRL_RESET_XS_DECL) {
    const ModelConst& m = c_model;
    e.q = lc.act ? m.pose0[lane == 0 ? 2 : lane + 2] : 0.0;
    e.qd = lc.act ? m.vel0[lane == 0 ? 2 : lane + 2] : 0.0;
    e.ox = m.pose0[0]; e.oy = m.pose0[1]; e.oxd = m.vel0[0]; e.oyd = m.vel0[1];
    e.tau = 0.0;
    Kin k = kinematics(lc, e TRL_KIN_XS_ARG);
    double comx = e.ox + warp_sum_all(lc.mass * k.cx) / m.total_mass;
    double comy = e.oy + warp_sum_all(lc.mass * k.cy) / m.total_mass;
    double newx = e.ox, newy = e.oy;
    if (lane == 0) {
        L.i(I_CONTACT) = 0;
        CounterRng rng = load_rng(L);
        // controller reset: default action, FSM state 0, counters zeroed (sim/TerrainRLCharController.cpp:47-58,
        // sim/DogController.cpp:210-216,640-650)
        double params[kNumParams];
        L.i(I_STANCE) = 0;   // cRaptorController::Reset -> SetStance(gDefaultStance)
        int id = build_base_action(L, rng, m.default_action, params);
        apply_action(L, id, params, comx, comy);
        L.i(I_EXP_FLAGS) = 0;
        L.i(I_FIRST_CYCLE) = 1;
        L.d(D_PREV_CYCLE_T) = 0.0; L.d(D_CUR_CYCLE_T) = 0.0; L.d(D_PREV_STUMBLE) = 0.0; L.d(D_CUR_STUMBLE) = 0.0;
        L.d(D_PREV_DIST_X) = 0.0; L.d(D_PREV_DIST_Y) = 0.0;
        L.d(D_PREV_COM_X) = comx; L.d(D_PREV_COM_Y) = comy;
        L.i(I_CMD) = -1;
        L.i(I_PENDING) = 0;
        // cSimCharSoftFall::Reset
        L.d(D_FALL_DIST_CNT) = 5.0; L.d(D_PREV_CHECK_X) = e.ox; L.d(D_PREV_CHECK_Y) = e.oy;
        L.i(I_FAIL_FALL_DIST) = 0; L.d(D_FALL_CONTACT_CNT) = 0.1; L.d(D_SUM_FALL) = 0.0;
        // ResetGround + InitCharacterPos
        GroundView g = load_ground(L, B);
        g.n[0] = g.n[1] = 0; g.flip = 0;
        g.update(-11.0, 9.0, m.terrain_type, m.terrain_params, 20.0);
        if (m.has_init_x) newx = m.init_x;
        newy = e.oy + g.sample(newx);
        store_ground(L, g);
        if (m.exp_mode) {
            L.i(I_CYCLE_COUNT) = 0;
            L.i(I_CMD) = rng.rand_int(0, m.n_actions);   // cScenarioExp::CommandRandAction
        } else {
            L.d(D_POS_START_X) = newx;
        }
        store_rng(L, rng);
    }
    __syncwarp();
    e.ox = shf(newx, 0); e.oy = shf(newy, 0);
}

__device__ __forceinline__ void load_env(Lane& L, const LinkC& lc, EnvRegs& e, int lane) {
    const int dq = (lane == 0) ? 2 : lane + 2;
    e.q = lc.act ? L.d(D_Q + dq) : 0.0;
    e.qd = lc.act ? L.d(D_QD + dq) : 0.0;
    e.tau = (lc.act && lane > 0) ? L.d(D_TAU + dq) : 0.0;
    e.ox = L.d(D_Q); e.oy = L.d(D_Q + 1); e.oxd = L.d(D_QD); e.oyd = L.d(D_QD + 1);
}
__device__ __forceinline__ void store_env(Lane& L, const LinkC& lc, const EnvRegs& e, int lane) {
    const int dq = (lane == 0) ? 2 : lane + 2;
    if (lc.act) { L.d(D_Q + dq) = e.q; L.d(D_QD + dq) = e.qd; L.d(D_TAU + dq) = (lane > 0) ? e.tau : 0.0; }
    if (lane == 0) {
        L.d(D_Q) = e.ox; L.d(D_Q + 1) = e.oy; L.d(D_QD) = e.oxd; L.d(D_QD + 1) = e.oyd;
        L.d(D_TAU) = 0.0; L.d(D_TAU + 1) = 0.0;
    }
}

// ================================================================================================ the kernel
// flags: bit0 do_ctrl (finish env-step k), bit1 do_phys (start env-step k+1), bit2 end of outer update
// Catch-up launch bookkeeping: every warp calls this once when it is done; the last one of the grid re-arms the consumed list.
__device__ __forceinline__ void catchup_leave(const Buffers& B, int prev) {
    __syncwarp();
    if ((threadIdx.x & 31) == 0) {
        __threadfence();
        const int total = (int)gridDim.x * kWarpsPerBlock;
        // one counter per list: the catch-up launches of two consecutive env-steps can be in flight at the same time
        if (atomicAdd(&B.catchup_done[prev], 1) == total - 1) { B.pending_count[prev] = 0; B.catchup_done[prev] = 0; __threadfence(); }
    }
}

__global__ void __launch_bounds__(kBlockThreads, TRL_STEP_MIN_BLOCKS)
trl_step_kernel(Buffers B, double h, int flags, int lists) {
    __shared__ double s_clx[4 * kMaxJoints + 6 * kMaxJoints], s_cly[4 * kMaxJoints];     // s_clx also carries the six per-body box tables
    __shared__ int s_cbody[4 * kMaxJoints];
#if TRL_SMEM_ALIGNED
    __shared__ __align__(16) double s_x[kWarpsPerBlock * X_END];
#else
    __shared__ double s_x[kWarpsPerBlock * X_END];
#endif
    const ModelConst& m = c_model;
    for (int t = threadIdx.x; t < m.n_corners; t += kBlockThreads) {
        s_clx[t] = LT(corner_lx, t); s_cly[t] = LT(corner_ly, t); s_cbody[t] = LT(corner_body, t);
    }
    for (int t = threadIdx.x; t < m.nj; t += kBlockThreads) {
        double* tb = s_clx + 4 * kMaxJoints;
        tb[t] = LT(body_ax, t); tb[kMaxJoints + t] = LT(body_ay, t); tb[2 * kMaxJoints + t] = LT(body_cos, t); tb[3 * kMaxJoints + t] = LT(body_sin, t);
        tb[4 * kMaxJoints + t] = LT(half_x, t); tb[5 * kMaxJoints + t] = LT(half_y, t);
    }
    TRL_LINK_TABLES_DECL
#if TRL_FIELD_SMEM
    static_assert(I_NUM_FIELDS <= kWarp, "one lane per i32 field");
    __shared__ double s_fd[kWarpsPerBlock * D_NUM_FIELDS];
    __shared__ int s_fi[kWarpsPerBlock * I_NUM_FIELDS];
#endif
    stage_link_tables(link_tabs);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // pending-decision lists (lag + 1 of them, used round robin by successive env-steps): envs that reach a cycle boundary in this
    // launch are appended to list `app`; an env's I_PENDING tag is 1 + the list it was last appended to.  The decision of a list is
    // made on a side stream and its envs are then caught up over `lag` env-steps by a side launch (trl_host.cu: enqueue_update), so
    // the main launch of step j skips every tagged env except those of list j % (lag + 1), whose tag is lag + 1 steps old.
    const int app = lists & 7, prev = (lists >> 3) & 7;

    int env = blockIdx.x * kWarpsPerBlock + warp;
    int env_end = B.n;
    {
        // env groups (trl_host.cu: enqueue_update): a main launch covers group g of G contiguous env ranges, each on its own stream,
        // so that the tail of one group's launch is filled by the next group's CTAs instead of leaving SM slots idle
        const int G = (lists >> 14) & 15;
        if (G > 1) {
            const int chunk = group_chunk(B.n, G);
            env += ((lists >> 10) & 15) * chunk;
            env_end = min(B.n, (((lists >> 10) & 15) + 1) * chunk);
        }
    }
    if (flags & kStepCatchUp) {
        // catch-up launch: one warp per entry of list `prev`, after the decision kernel has served it.
        // The last CTA to leave re-arms that list (the main launch three steps later appends to it again).
        const int count = B.pending_count[prev];
        if (env >= count) { catchup_leave(B, prev); return; }
        env = B.pending_list[prev * B.n + env];
    } else if (env >= env_end) {
        return;
    }
    Lane L{nullptr, env, B.n, B.d, B.i};
    int tag = 0;
    if (flags & kStepSkipPending) {
        // overlapped main launch: envs waiting for a decision or being caught up are not touched
        tag = L.i(I_PENDING);
        if (tag != 0 && tag != 1 + app) return;
    }
#if TRL_FIELD_SMEM
    // every field of this env in one batch of independent loads (the planes are B.n apart: a field costs an L2 round trip wherever it
    // is first touched); the env is private to this warp for the whole launch, so the copy is written back in one batch at the end
    {
        double* sd = s_fd + warp * D_NUM_FIELDS;
        int* si = s_fi + warp * I_NUM_FIELDS;
        for (int f = lane; f < D_NUM_FIELDS; f += kWarp) sd[f] = L.D[(size_t)f * L.n + env];
        if (lane < I_NUM_FIELDS) si[lane] = L.I[(size_t)lane * L.n + env];
        __syncwarp();
        L.sd = sd; L.si = si;
    }
#endif
    if (flags & kStepSkipPending) {
        if (tag == 1 + app && lane == 0) L.i(I_PENDING) = 0;   // stale tag of lag + 1 steps ago (already caught up)
    } else if (!(flags & kStepCatchUp) && (flags & 1) && lane == 0) {
        L.i(I_PENDING) = 0;    // serial schedule / end of the update: every decision has been served before this launch
    }
    double* xs = s_x + warp * X_END;
    const LinkC lc = load_link(lane, link_tabs);
    EnvRegs e;
    load_env(L, lc, e, lane);
    int contact = L.i(I_CONTACT);

#if TRL_REUSE_KIN
    Kin k_ctrl;
    bool have_k = false;       // warp-uniform: the controller half ran and nothing has moved the state since
#endif
#ifdef TRL_CG_VARIANT
    const int reps = max(1, (lists >> 6) & 15);        // a catch-up launch advances its envs by several env-steps (registers stay live)
#else
    constexpr int reps = 1;                            // catch-up launches are the other build of this file (trl_step_cg.cu)
#endif
    for (int rep = 0; rep < reps; ++rep) {
    if (flags & 1) {
        // ---------------- controller half of env-step k
        Kin k = kinematics(lc, e TRL_KIN_XS_ARG);
#if TRL_REUSE_KIN
        k_ctrl = k; have_k = true;
#endif
        e.tau = controller_torque(L, lc, e, k, xs, lane, h, contact);
        if (lane == 0) {
            // fall checks (sim/SimCharSoftFall.cpp:74-125)
            double cnt = L.d(D_FALL_DIST_CNT) - h;
            if (cnt <= 0.0) {
                double dx = e.ox - L.d(D_PREV_CHECK_X), dy = e.oy - L.d(D_PREV_CHECK_Y);
                if (dx * dx + dy * dy < 0.25) L.i(I_FAIL_FALL_DIST) = 1;
                L.d(D_PREV_CHECK_X) = e.ox; L.d(D_PREV_CHECK_Y) = e.oy;
                cnt = 5.0;
            }
            L.d(D_FALL_DIST_CNT) = cnt;
            double cc = L.d(D_FALL_CONTACT_CNT) - h;
            if (cc <= 0.0) {
                double val = ((unsigned)contact & m.fall_mask) ? 1.0 : 0.0;
                const double norm = (1.0 + 1.0 / (1.0 - 0.9));
                L.d(D_SUM_FALL) = val / norm + 0.9 * L.d(D_SUM_FALL);
                cc = 0.1;
            }
            L.d(D_FALL_CONTACT_CNT) = cc;
            unsigned lo = (unsigned)L.i(I_STEPS_LO) + 1u;
            L.i(I_STEPS_LO) = (int)lo;
            if (lo == 0u) L.i(I_STEPS_HI) += 1;
        }
        __syncwarp();
        // PostSubstepUpdate: cycle boundary bookkeeping
        int nc = 0, fl = 0;
        if (lane == 0) { nc = (L.i(I_STATE) == 0 && L.d(D_PHASE) == 0.0) ? 1 : 0; fl = has_fallen(L, e.q) ? 1 : 0; }
        nc = shfi(nc, 0); fl = shfi(fl, 0);
        if (nc) {
            if (m.exp_mode) exp_new_cycle_update(L, B, fl != 0, lane);
            else if (lane == 0) L.i(I_CYCLE_COUNT) += 1;
        }
    }

    if (flags & 4) {
        // ---------------- end of the outer update: fall -> episode bookkeeping + reset
        int fallen = 0, new_cycle = 0;
        if (lane == 0) { fallen = has_fallen(L, e.q) ? 1 : 0; new_cycle = (L.i(I_STATE) == 0 && L.d(D_PHASE) == 0.0) ? 1 : 0; }
        fallen = shfi(fallen, 0); new_cycle = shfi(new_cycle, 0);
        bool do_reset = false;
        if (m.exp_mode) {
            if (!new_cycle && fallen) { exp_new_cycle_update(L, B, true, lane); do_reset = true; }
        } else if (fallen) {
            if (lane == 0 && L.i(I_CYCLE_COUNT) >= 1) {
                double dist = e.ox - L.d(D_POS_START_X);
                int ec = L.i(I_EPISODE_COUNT);
                L.d(D_AVG_DIST) = (ec * L.d(D_AVG_DIST) + dist) / (ec + 1.0);
                L.i(I_EPISODE_COUNT) = ec + 1;
                int slot = atomicAdd(B.dist_count, 1);
                if (slot < B.dist_cap) { B.dist_log[slot] = dist; B.dist_env[slot] = env; }
            }
            do_reset = true;
        }
        if (do_reset) {
            reset_env(L, lc, e, B, lane TRL_RESET_XS_ARG); contact = 0;
#if TRL_REUSE_KIN
            have_k = false;
#endif
        }
    }

    if (flags & 2) {
        // ---------------- physics half of env-step k+1
        GroundView g = load_ground(L, B);
        const int ns = m.num_sim_substeps;
        const double dt = h / ns;
        // terrain maximum over the window the character's corners can reach during this env-step; kClearMargin covers the
        // contact tolerance measured along the surface normal on (near-)vertical cliff faces and the root's travel
        const double clear_y = g.window_max(e.ox - m.reach - 0.25, e.ox + m.reach + 0.25, lane) + kClearMargin;
#if TRL_REUSE_KIN
        for (int s = 0; s < ns; ++s)
            contact = physics_substep(lc, e, g, s_clx, s_cly, s_cbody, lane, dt, clear_y TRL_PHYS_XS_ARG, (s == 0 && have_k) ? &k_ctrl : nullptr);
#else
        for (int s = 0; s < ns; ++s) contact = physics_substep(lc, e, g, s_clx, s_cly, s_cbody, lane, dt, clear_y TRL_PHYS_XS_ARG);
#endif
        // UpdateGround (scenarios/ScenarioSimChar.cpp:564-572): regenerate a segment when the view window crosses it
        {
            int smin = g.seg_id(0), smax = g.seg_id(1);
            double bmin = e.ox - 2.0, bmax = e.ox + 10.0 + 1.0;
            bool need = !(bmax < g.seg_max_x(smax) && bmin > g.seg_min_x(smin));
            if (need) {
                if (lane == 0) { g.update(bmin, bmax, m.terrain_type, m.terrain_params, 20.0); store_ground(L, g); }
                __syncwarp();
                g = load_ground(L, B);
            }
        }
        // head of cDogController::Update: cycle timers, stumble counter, gait FSM (lane 0)
        int end_step = 0;
        if (lane == 0) {
            L.i(I_CONTACT) = contact;
            L.d(D_CUR_CYCLE_T) += h;
            if ((unsigned)contact & m.stumble_mask) L.d(D_CUR_STUMBLE) += h;
            int state = L.i(I_STATE);
            int first = L.i(I_FIRST_CYCLE);
            double phase = L.d(D_PHASE) + h / L.d(D_PARAMS + mTransTime);
            bool advance = first != 0;
            int ns2;
            if (TRL_IS_RAPTOR(m)) {
                // cRaptorController::UpdateState (sim/RaptorController.cpp:804-849): Contact, Down, Passing are timed,
                // Up ends when the swing toe touches down; the stance flips at the end of every cycle but the first
                const int sw_toe = (L.i(I_STANCE) == 0 ? rLeftHip : rRightHip) + 3;
                if (state != rsUp && phase >= 1.0) advance = true;
                if (state == rsUp && ((contact >> sw_toe) & 1)) advance = true;
                ns2 = first ? rsContact : (state == rsUp ? -1 : state + 1);
            } else {
                if ((state == sBackStance || state == sFrontStance) && phase >= 1.0) advance = true;
                if (state == sExtend && ((contact >> jFinger) & 1)) advance = true;
                if (state == sGather && ((contact >> jToe) & 1)) advance = true;
                ns2 = first ? sBackStance : (state == sGather ? -1 : state + 1);
            }
            L.d(D_PHASE) = phase;
            if (advance) {
                if ((ns2 < 0) || first) {
                    end_step = 1;
                    if (TRL_IS_RAPTOR(m) && !first) L.i(I_STANCE) = 1 - L.i(I_STANCE);   // FlipStance
                } else { L.i(I_STATE) = ns2; L.d(D_PHASE) = 0.0; set_state_params(L, ns2); }
            }
        }
        end_step = shfi(end_step, 0);
        __syncwarp();
        if (end_step) {
            // cycle boundary: build the policy state and hand the env to the decision kernel
            Kin k = kinematics(lc, e TRL_KIN_XS_ARG);
            build_poli_state(lc, e, k, B, g, env, lane, TRL_IS_RAPTOR(m) ? L.i(I_STANCE) : 0);
            double comx = e.ox + warp_sum_all(lc.mass * k.cx) / m.total_mass;
            double comy = e.oy + warp_sum_all(lc.mass * k.cy) / m.total_mass;
            if (lane == 0) {
                B.com_stash[env] = comx; B.com_stash[B.n + env] = comy;
                L.i(I_FIRST_CYCLE) = 0;
                if (flags & kStepCatchUp) {
                    // a gait cycle shorter than the catch-up depth (the shipped controllers' timed states alone last 150 env-steps): the side
                    // launch cannot hand the env to a decision in time -- flag it (trl_sync reports; TRL_SERIAL_SCHEDULE=1 runs it)
                    B.catchup_done[kMaxLists] = 1;
                }
                L.i(I_PENDING) = 1 + app;
                int slot = atomicAdd(&B.pending_count[app], 1);
                B.pending_list[app * B.n + slot] = env;
            }
        }
    }
    }   // rep
    store_env(L, lc, e, lane);
#if TRL_FIELD_SMEM
    __syncwarp();
    for (int f = lane; f < D_NUM_FIELDS; f += kWarp) L.D[(size_t)f * L.n + env] = L.sd[f];
    if (lane < I_NUM_FIELDS) L.I[(size_t)lane * L.n + env] = L.si[lane];
#endif
    if (flags & kStepCatchUp) catchup_leave(B, prev);
}

// Initial reset of every env (trl_create / trl_reset): seeds the terrain RNG and runs the episode reset.
__global__ void __launch_bounds__(kBlockThreads)
trl_reset_kernel(Buffers B, const uint64_t* terrain_seeds, const int* env_ids, int count, int reseed) {
    TRL_LINK_TABLES_DECL
#if TRL_LINK_SMEM
    stage_link_tables(link_tabs);
    __syncthreads();
#endif
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int idx = blockIdx.x * kWarpsPerBlock + warp;
    if (idx >= count) return;
    const int env = env_ids ? env_ids[idx] : idx;
    Lane L{nullptr, env, B.n, B.d, B.i};
    const ModelConst& m = c_model;
    if (reseed && lane == 0) {
        uint64_t seed = terrain_seeds ? terrain_seeds[idx] : (uint64_t)(1 + env);
        L.i(I_TERRAIN_RNG) = (int)TerrainRng::seed_state(seed);
        L.i(I_RNG_CTR_LO) = 0; L.i(I_RNG_CTR_HI) = 0;
        L.i(I_CYCLE_COUNT) = 0; L.i(I_EPISODE_COUNT) = 0; L.d(D_AVG_DIST) = 0.0; L.d(D_POS_START_X) = 0.0;
        L.i(I_STEPS_LO) = 0; L.i(I_STEPS_HI) = 0; L.i(I_TUPLE_FLAGS) = 0;
        L.i(I_SEG_N0) = 0; L.i(I_SEG_N1) = 0; L.i(I_SEG_FLIP) = 0; L.d(D_SEG_MINX0) = 0.0; L.d(D_SEG_MINX1) = 0.0;
        for (int j = 0; j < m.nj; ++j) L.d(D_PD_TARGET + j) = m.target_theta0[j];
        for (int k = 0; k < kNumParams; ++k) L.d(D_PARAMS + k) = 0.0;
        L.d(D_CUR_CYCLE_T) = 0.0; L.d(D_CUR_STUMBLE) = 0.0; L.d(D_PREV_COM_X) = 0.0; L.d(D_PREV_COM_Y) = 0.0;
    }
    __syncwarp();
    const LinkC lc = load_link(lane, link_tabs);
    EnvRegs e;
#if TRL_KIN_SMEM
    __shared__ __align__(16) double s_pfx[kWarpsPerBlock * 3 * 2 * kWarp];
    double* xs = s_pfx + warp * 3 * 2 * kWarp - X_PFX;    // kinematics() only touches [X_PFX, X_PFX_END)
#endif
    reset_env(L, lc, e, B, lane TRL_RESET_XS_ARG);
    store_env(L, lc, e, lane);
}

// Terrain look-ahead: regenerates, ahead of the step kernels of the coming outer update, the segment of every env
// whose view window [x-2, x+11] will cross the end of its terrain during that update (look-ahead 0.5 m >> the distance
// covered in 1/30 s).  cGroundVar2D::Update's result does not depend on *when* it runs (segment bounds, joint height
// and the generator stream are functions of the previous segment only, sim/GroundVar2D.cpp:43-91), so the terrain is
// bit-identical to generating it at the exact step; doing it here keeps the serial generator off the critical path
// of the 21 step launches (it used to stall one warp for ~90 us per regeneration).
__global__ void trl_terrain_kernel(Buffers B, double lookahead) {
    const int env = blockIdx.x * blockDim.x + threadIdx.x;
    if (env >= B.n) return;
    const ModelConst& m = c_model;
    Lane L{nullptr, env, B.n, B.d, B.i};
    const double x = L.d(D_Q);
    GroundView g = load_ground(L, B);
    const int smin = g.seg_id(0), smax = g.seg_id(1);
    const double bmin = x - 2.0, bmax = x + 10.0 + 1.0 + lookahead;
    if (bmax < g.seg_max_x(smax) && bmin > g.seg_min_x(smin)) return;
    g.update(bmin, bmax, m.terrain_type, m.terrain_params, 20.0);
    store_ground(L, g);
}
void launch_terrain(const Buffers& B, double lookahead, cudaStream_t st) {
    TRL_LAUNCH(trl_terrain_kernel, (B.n + 127) / 128, 128, 0, st, B, lookahead);
}

// Batch statistics (cOptScenarioPoliEval::OutputResults merges the same counters under a mutex): one block reduces
// cycles / episodes / env-steps / episode-weighted distance into out[4] (doubles).
__global__ void trl_stats_kernel(Buffers B, double* out) {
    __shared__ double red[4][32];
    double c = 0, e = 0, st = 0, ds = 0;
    for (int env = threadIdx.x; env < B.n; env += blockDim.x) {
        int ec = B.i[(size_t)I_EPISODE_COUNT * B.n + env];
        c += B.i[(size_t)I_CYCLE_COUNT * B.n + env];
        e += ec;
        st += (double)(((uint64_t)(uint32_t)B.i[(size_t)I_STEPS_HI * B.n + env] << 32) | (uint32_t)B.i[(size_t)I_STEPS_LO * B.n + env]);
        ds += B.d[(size_t)D_AVG_DIST * B.n + env] * ec;
    }
    c = warp_sum_all(c); e = warp_sum_all(e); st = warp_sum_all(st); ds = warp_sum_all(ds);
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = c; red[1][w] = e; red[2][w] = st; red[3][w] = ds; }
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        for (int k = 0; k < 4; ++k) {
            double v = l < nw ? red[k][l] : 0.0;
            v = warp_sum_all(v);
            if (l == 0) out[k] = v;
        }
    }
}
void launch_stats(const Buffers& B, double* out, cudaStream_t st) { TRL_LAUNCH(trl_stats_kernel, 1, 1024, 0, st, B, out); }

// ---- host-side launch helpers (called from trl_host.cu)
cudaError_t upload_model(const ModelConst& mc) {
    cudaError_t e = cudaMemcpyToSymbol(c_model, &mc, sizeof(ModelConst));
#if TRL_TABLE_MIRROR
    if (e == cudaSuccess) e = cudaMemcpyToSymbol(g_model, &mc, sizeof(ModelConst));
#endif
    return e;
}
size_t step_smem_bytes() { return 0; }
cudaError_t configure_step_kernels() {
#if TRL_SMEM_ALIGNED && !defined(TRL_SIMT_EMU)
    // experiment builds stage lane-to-lane exchanges in shared memory (up to 42 KB per CTA): ask for the largest carve-out so
    // that the register-limited 4 CTAs per SM still fit
    return cudaFuncSetAttribute(trl_step_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared);
#else
    return cudaSuccess;
#endif
}
void launch_step(const Buffers& B, double h, int flags, int lists, cudaStream_t st, int group, int n_groups) {
    int envs = B.n;
    if (n_groups > 1) {
        const int chunk = group_chunk(B.n, n_groups);
        envs = std::max(0, std::min(B.n, (group + 1) * chunk) - group * chunk);
        lists |= (group << 10) | (n_groups << 14);
        if (envs == 0) return;
    }
    int blocks = (envs + kWarpsPerBlock - 1) / kWarpsPerBlock;
    TRL_LAUNCH(trl_step_kernel, blocks, kBlockThreads, 0, st, B, h, flags, lists);
}
void launch_reset(const Buffers& B, const uint64_t* seeds, const int* env_ids, int count, int reseed, cudaStream_t st) {
    int blocks = (count + kWarpsPerBlock - 1) / kWarpsPerBlock;
    TRL_LAUNCH(trl_reset_kernel, blocks, kBlockThreads, 0, st, B, seeds, env_ids, count, reseed);
}

}  // namespace TRL_IMPL_NS

#if !defined(TRL_CG_VARIANT)
#include "trl_decide.cuh"
#include "trl_decide2.cuh"
#endif
