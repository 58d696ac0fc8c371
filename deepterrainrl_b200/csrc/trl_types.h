// deepterrainrl_b200 -- batched locomotion rollout engine for B200 (sm_100a).
// Shared host/device types: model constants (constant memory), SoA state layout in HBM, launch parameters.
//
// HBM layout: every per-env scalar lives in a struct-of-arrays plane `plane[field][env]`, so a warp
// (32 consecutive envs, one env per lane) reads/writes each field as one coalesced 256-byte (f64) or 128-byte (i32)
// transaction.  Bulk per-env records that are touched by one env at a time (terrain strips, policy state, tuple
// staging) are env-major rows instead.
#pragma once
#include <cstdint>

// Kernel launches and dynamic shared memory are spelled through two macros so that the same sources also compile for the
// test-only SIMT emulator (tests/simt/, -DTRL_SIMT_EMU: g++, one fiber per CUDA thread).  In the product build they expand to
// exactly the CUDA syntax they name.
#ifdef TRL_SIMT_EMU
#define TRL_LAUNCH(kern, grid, block, smem, st, ...) SIMT_LAUNCH(1, kern, grid, block, smem, st, __VA_ARGS__)
#define TRL_LAUNCH_CLUSTER(csize, kern, grid, block, smem, st, ...) SIMT_LAUNCH(csize, kern, grid, block, smem, st, __VA_ARGS__)
#define TRL_DYN_SHARED(type, name) type* name = reinterpret_cast<type*>(simt::dyn_smem())
#else
#define TRL_LAUNCH(kern, grid, block, smem, st, ...) kern<<<grid, block, smem, st>>>(__VA_ARGS__)
#define TRL_LAUNCH_CLUSTER(csize, kern, grid, block, smem, st, ...) kern<<<grid, block, smem, st>>>(__VA_ARGS__)   // cluster size: __cluster_dims__ of the kernel
#define TRL_DYN_SHARED(type, name) extern __shared__ type name[]
#endif

#if defined(__CUDACC__) && !defined(TRL_SIMT_EMU)
#define TRL_HD __host__ __device__
#else
#define TRL_HD
#endif

namespace trl {

constexpr int kMaxJoints = 21;      // dog / goat: 21 joints, raptor: 19
constexpr int kMaxDof = 23;
constexpr int kNumParams = 37;      // max gait controller parameter vector: dog 30 (sim/DogController.h:14-46), raptor 37 (sim/RaptorController.h:14-47)
constexpr int kNumGroundSamples = 200;
constexpr int kTerrainCap = 512;    // floats per terrain segment (20 m nominal + overshoot + 2 m pad at 0.1 m)
constexpr int kTerrainParams = 40;
constexpr int kMaxActions = 16;
constexpr int kMaxCtrlSets = 8;
constexpr int kMaxNetOut = 96;
constexpr int kWarp = 32;
constexpr int kMaxGroups = 8;      // env groups a main step launch can be split into (one stream each)
// envs per group when a main launch is split into n_groups launches: a multiple of 16 envs (= one 128-byte line of an f64 plane)
TRL_HD inline int group_chunk(int n, int n_groups) { return ((n + 16 * n_groups - 1) / (16 * n_groups)) * 16; }
constexpr int kMaxLists = 8;       // pending-decision lists (overlap depth + 1 <= 8)

// ---- f64 SoA planes -------------------------------------------------------------------------------------------
enum DField : int {
    D_Q = 0,                               // 23 generalised positions [x, y, th_root, th_1..]
    D_QD = D_Q + kMaxDof,                  // 23 generalised velocities
    D_TAU = D_QD + kMaxDof,                // 23 held (clamped) joint torques, applied by the next env-step
    D_PARAMS = D_TAU + kMaxDof,            // 30 current action parameters
    D_PD_TARGET = D_PARAMS + kNumParams,   // 21 PD target angles
    D_PHASE = D_PD_TARGET + kMaxJoints,
    D_CUR_CYCLE_T, D_PREV_CYCLE_T, D_CUR_STUMBLE, D_PREV_STUMBLE,
    D_PREV_COM_X, D_PREV_COM_Y, D_PREV_DIST_X, D_PREV_DIST_Y,
    D_FALL_DIST_CNT, D_FALL_CONTACT_CNT, D_SUM_FALL, D_PREV_CHECK_X, D_PREV_CHECK_Y,
    D_AVG_DIST, D_POS_START_X,
    D_SEG_MINX0, D_SEG_MINX1,              // terrain segment origins
    D_NUM_FIELDS
};

// ---- i32 SoA planes -------------------------------------------------------------------------------------------
enum IField : int {
    I_STATE = 0, I_FIRST_CYCLE, I_ACTION_ID, I_CONTACT, I_FAIL_FALL_DIST,
    I_EXP_FLAGS,        // bit0 exp_critic, bit1 exp_actor, bit2 off_policy
    I_CYCLE_COUNT, I_EPISODE_COUNT, I_PENDING, I_CMD,
    I_SEG_N0, I_SEG_N1, I_SEG_FLIP, I_TERRAIN_RNG,   // minstd_rand0 state of the env's terrain generator
    I_TUPLE_FLAGS, I_RNG_CTR_LO, I_RNG_CTR_HI, I_STEPS_LO, I_STEPS_HI,
    I_STANCE,           // raptor: 0 = right leg is the stance leg, 1 = left
    I_NUM_FIELDS
};

struct PhysParams {
    double kn, dn, mu, v_eps, contact_tol, k_lim, d_lim;
    int vertex_contacts;      // terrain vertices inside body boxes produce contacts too (opt-in: TRL_VERTEX_CONTACTS=1)
};

// Model / scene constants, uploaded once into __constant__ memory (uniform across lanes -> constant-cache broadcast).
struct ModelConst {
    int nj, ndof;
    int parent[kMaxJoints];
    int dof[kMaxJoints];                   // parameter offset of each joint
    double attach_x[kMaxJoints], attach_y[kMaxJoints];
    double lim_lo[kMaxJoints], lim_hi[kMaxJoints];
    int has_limit[kMaxJoints];
    // bodies
    double mass[kMaxJoints], body_ax[kMaxJoints], body_ay[kMaxJoints], body_theta[kMaxJoints];
    double body_cos[kMaxJoints], body_sin[kMaxJoints];
    double half_x[kMaxJoints], half_y[kMaxJoints];
    double izz_o[kMaxJoints];              // planar rotational inertia about the joint origin
    double izz_c[kMaxJoints];              // planar rotational inertia about the body COM
    int collidable[kMaxJoints];
    double total_mass;
    // tree topology helpers for the level-synchronous warp passes
    int depth[kMaxJoints], max_depth;
    int child[kMaxJoints][4];              // up to 4 children per link (-1 = none)
    int anc_pow[kMaxJoints][4];            // 2^k-th ancestor of each link (k = 0..3), -1 if it does not exist
    // inward (leaf -> root) pass schedule: link j is eliminated in round acc_round[j] and handed to its parent at the end
    // of that round; siblings get distinct rounds so that every lane receives from at most ONE lane per round.
    // acc_src[j] packs, 5 bits per round, the lane that lane j receives from (31 = none: lane 31 is idle and holds zeros)
    int acc_round[kMaxJoints], acc_rounds;
    unsigned long long acc_src[kMaxJoints];
    int n_corners;                         // 4 * number of collidable bodies
    int corner_body[4 * kMaxJoints];
    double corner_lx[4 * kMaxJoints], corner_ly[4 * kMaxJoints];   // corner position in the link (joint) frame
    int corner_base[kMaxJoints];           // first corner index of a body (-1 if not collidable)
    double reach;                          // upper bound of |box corner - root joint| over all poses
    unsigned anc_mask_toe, anc_mask_finger;   // links on the chain effector -> root (inclusive)
    unsigned vf_mask_toe, vf_mask_finger;     // links that receive the virtual force (chain up to root / torso, exclusive)
    // PD
    double kp[kMaxJoints], kd[kMaxJoints], torque_lim[kMaxJoints], target_theta0[kMaxJoints], target_vel[kMaxJoints];
    int world_pd[kMaxJoints];
    // gait controller
    int char_type;                         // 1 dog / goat, 2 raptor
    int n_params, n_opt, misc_max, sp_max;  // parameter vector layout (30/29/6/6 dog, 37/28/5/8 raptor)
    int opt_idx[kNumParams];               // indices of the optimised parameters (= actor outputs)
    unsigned stumble_mask, fall_mask;      // parts whose ground contact counts as a stumble / towards a fall
    double exp_noise;                      // exploration noise std (sim/DogControllerMACE.cpp:3-9, RaptorControllerMACE.cpp:7)
    int n_ctrl, n_actions, default_action, grav_comp, virt_forces, is_mace;
    double ctrl_params[kMaxCtrlSets][kNumParams];
    int act_idx0[kMaxActions], act_idx1[kMaxActions], act_cyclic[kMaxActions];
    double act_blend[kMaxActions];
    double target_vel_x;
    // initial state
    double pose0[kMaxDof], vel0[kMaxDof];
    int has_init_x;
    double init_x;
    // terrain
    int terrain_type;
    double terrain_params[kTerrainParams];
    // scenario
    int num_sim_substeps, exp_mode, has_net;
    double gx, gy;
    // net dims
    int n_in, n_char, n_out, n_frags, frag;
    double out_scale_actor0[32];           // OutputScale of actor 0 (exploration noise scale)
    PhysParams phys;
    uint64_t rng_seed;
};

// Exploration settings can change between updates (cScenarioExp::SetExpRate/Temp/BaseActionRate).
struct ExpSettings {
    int enable;
    double rate, temp, base_rate, noise;
};

// Policy network weights (device pointers, f64, Caffe blob order).
struct NetWeights {
    const double *conv0_w, *conv0_b, *conv1_w, *conv1_b, *conv2_w, *conv2_b, *tip0_w, *tip0_b, *ip0_w, *ip0_b;
    const double *h0_w[4], *h0_b[4], *h1_w[4], *h1_b[4];
    const double *in_off, *in_scale, *out_off, *out_scale;
};

// Device buffers of one batch of environments.
struct Buffers {
    int n;                 // number of envs
    double* d;             // [D_NUM_FIELDS][n]
    int* i;                // [I_NUM_FIELDS][n]
    float* terrain;        // [n][2][kTerrainCap]
    double* poli_state;    // [n][S]   policy state built at the last decision
    double* net_out;       // [n][kMaxNetOut]
    double* tuple_sbeg;    // [n][S]
    double* tuple_action;  // [n][kNumParams]
    double* com_stash;     // [2][n]   COM at decision time
    int* pending_list;     // [lists][n]  used round robin by successive env-steps (see trl_host.cu: enqueue_update)
    int* pending_count;    // [kMaxLists]
    int* catchup_done;     // [kMaxLists + 1]  [l] CTA completion counter of the catch-up launch of list l (re-arms the list it consumed), [kMaxLists] fault flag
    // outputs
    double* tuples;        // [tuple_cap][1 + S + A + S]
    uint32_t* tuple_flags; // [tuple_cap]
    int* tuple_env;        // [tuple_cap]
    int* tuple_count;      // [1]
    int tuple_cap;
    double* dist_log;      // [dist_cap] episode distances
    int* dist_env;         // [dist_cap]
    int* dist_count;       // [1]
    int dist_cap;
    int S;                 // policy state size
    int A;                 // action record size (1 + number of optimised parameters)
};

}  // namespace trl
