// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/README.md). Never linked into the product path.
//
// One environment of the ScenarioExp / ScenarioPoliEval step loop for the dog / goat characters:
// physics sub-steps -> streaming ground -> gait controller (RBD model, FSM, feedback, implicit PD, gravity
// compensation, virtual forces) -> torque clamp -> fall logic -> cycle bookkeeping (policy decision, reward,
// tuple record, episode statistics).
//
// Controller / scenario arithmetic restates the reference (file:line under /root/reference):
//   scenarios/ScenarioSimChar.cpp:121-182,539-583   Reset, Update loop, InitCharacterPos, UpdateGround, ResetGround
//   scenarios/ScenarioExp.cpp:63-98,209-243,296-318 Exp Reset/Update, NewCycleUpdate, RecordTuple, IsValidTuple
//   scenarios/ScenarioExpMACE.cpp:16-28             tuple flags
//   scenarios/ScenarioPoliEval.cpp:88-125,202-259,406-410   eval Reset/Update, RecordDistTraveled, cycle counting
//   sim/SimCharacter.cpp:75-107,166-225,396-434,637-654      Reset, Update, pose/vel, COM, ApplyControlForces
//   sim/SimCharSoftFall.cpp:36-125, sim/SimDog.cpp:83-162    fall / stumble logic
//   sim/Joint.cpp:171-201,257-264                   torque clamp
//   sim/DogController.cpp:12-38,229-268,552-650,805-1175,1318-1426   the gait controller
//   sim/TerrainRLCharController.cpp:47-58,130-146,168-285            Reset, ApplyAction, terrain scan, policy state
//   sim/BaseControllerMACE.cpp:58-68,254-318,339-396,437-518         MACE decode / exploration
//   sim/DogControllerMACE.cpp:16-91, sim/GoatControllerMACE.cpp:11-14
//   sim/ImpPDController.cpp:234-310, sim/PDController.cpp:181-208    stable-PD torques
//
// The controller part is PINNED against the reference's own compiled controller sources (oracle/_ref/libref_ctrl.so,
// tests/test_ref_pinning_cpu.py): same joint torques, gait-machine state and policy state vectors over 1200-1800 env-steps for the
// dog (fixed gait and MACE), the goat and the raptor.  The scenario bookkeeping stays an unpinned restatement.
//
// PHYSICS IS NOT A RESTATEMENT.  The reference steps Bullet (btDiscreteDynamicsWorld, maximal coordinates, an
// un-vendored dependency of unpinned version: premake4.lua:142-195, sim/World.cpp:61-105), which cannot be built
// or restated here.  `physics_substep` below is this project's own reduced-coordinate planar model:
// forward dynamics M(q) qdd = tau - C + contact/limit forces with linearly-implicit spring-damper contacts, solved
// densely with the reference's own CRBA + RNEA (cRBDUtil::SolveForDyna, sim/RBDUtil.cpp:86-99) + LDL^T.  The CUDA
// path computes the same accelerations with a planar articulated-body (ABA) recursion instead.  PARITY UNPINNED
// against Bullet; pinned only GPU <-> this oracle.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

#include "net.h"
#include "pack_reader.h"
#include <cstdlib>
#include "rbd.h"
#include "terrain.h"

namespace orc {

// ---- engine constants shared (by value, restated) with the CUDA path; DESIGN.md §3 ----
struct PhysParams {
    double kn = 2.0e5;      // contact normal stiffness [N/m]
    double dn = 2.0e3;      // contact normal damping [N s/m]
    double mu = 0.81;       // Bullet combines friction by product: 0.9 * 0.9 (sim/SimCharacter.cpp:17, sim/GroundVar2D.cpp:10)
    double v_eps = 0.01;    // friction regularisation speed [m/s]
    double contact_tol = 0.00025;  // 0.001 Bullet units / world_scale 4 (sim/ContactManager.cpp:74)
    double k_lim = 2.0e4;   // joint-limit stiffness [N m/rad]
    double d_lim = 20.0;    // joint-limit damping [N m s/rad]
    int vertex_contacts = std::getenv("ORC_VTX") ? std::atoi(std::getenv("ORC_VTX")) : 0;   // terrain vertices inside body boxes produce contacts too (opt-in: ORC_VTX=1)
};

// dog / goat joint indices (sim/SimDog.h:11-36)
enum DogJoint {
    jRoot, jSpine0, jSpine1, jSpine2, jSpine3, jTorso, jNeck0, jNeck1, jHead, jTail0, jTail1, jTail2, jTail3,
    jShoulder, jElbow, jWrist, jFinger, jHip, jKnee, jAnkle, jToe, jDogMax
};
enum DogState { sBackStance, sExtend, sFrontStance, sGather, sStateMax };
enum DogMisc { mTransTime, mCv, mBackForceX, mBackForceY, mFrontForceX, mFrontForceY, mMiscMax };
enum DogStateParam { spSpineCurve, spShoulder, spElbow, spHip, spKnee, spAnkle, spMax };
constexpr int kDogParams = mMiscMax + sStateMax * spMax;  // 30
constexpr int kDogOptParams = kDogParams - 1;             // all but TransTime (sim/DogController.cpp:81-121)
// raptor (sim/SimRaptor.h:11-33, sim/RaptorController.h:14-47, sim/RaptorController.cpp:11-122)
enum RaptorJoint {
    rRoot, rSpine0, rSpine1, rSpine2, rSpine3, rHead, rTail0, rTail1, rTail2, rTail3, rTail4,
    rRightHip, rRightKnee, rRightAnkle, rRightToe, rLeftHip, rLeftKnee, rLeftAnkle, rLeftToe, rRaptorMax
};
enum RaptorState { rsContact, rsDown, rsPassing, rsUp };
enum RaptorMisc { rmTransTime, rmCv, rmCd, rmForceX, rmForceY, rmMiscMax };
enum RaptorStateParam { rpRootPitch, rpSpineCurve, rpStanceHip, rpStanceKnee, rpStanceAnkle, rpSwingHip, rpSwingKnee, rpSwingAnkle, rpMax };
constexpr int kRaptorParams = rmMiscMax + 4 * rpMax;      // 37
constexpr int kMaxParams = 37;
static const bool kRaptorOptMask[kRaptorParams] = {
    false, true, true, false, false,
    true, false, true, true, true, true, true, true,
    true, false, true, true, true, true, true, true,
    false, false, true, true, true, true, true, true,
    false, false, true, true, true, true, true, true};
constexpr int kNumGroundSamples = 200;
enum TupleFlag { fFail = 0, fExpCritic = 1, fExpActor = 2 };  // learning/MACETrainer.h:11-17

inline double wrap_pi(double a) {
    // axis-angle extraction of a z-rotation: acos(cos a) signed by sin a (util/MathUtil.cpp:129-149)
    double c = std::cos(a), s = std::sin(a);
    double th = std::acos(std::min(1.0, std::max(-1.0, c)));
    return (s >= 0) ? th : -th;
}

// counter-based RNG for the exploration draws (the reference uses one racy global std engine shared by all
// threads, util/MathUtil.cpp:4 -- only distributional parity is meaningful; ours is deterministic per env)
struct CounterRng {
    uint64_t key = 0, ctr = 0;
    // pinning mode (tests/test_ref_pinning_cpu.py): draw from the restated cRand instead (terrain.h's Rand, the engine behind the
    // reference's process-global cMathUtil::Rand*), so that exploration can be compared draw for draw with the compiled reference
    Rand* ref = nullptr;
    static uint64_t mix(uint64_t z) {
        z += 0x9E3779B97F4A7C15ull;
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
    void seed(uint64_t s, uint64_t stream) { key = mix(s ^ mix(stream)); ctr = 0; }
    uint64_t next() { return mix(key + (ctr++) * 0xD1342543DE82EF95ull); }
    double uniform() { return ref ? ref->rand_double() : (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
    int rand_int(int mn, int mx) {  // cRand::RandInt(min, max)
        if (ref) return ref->rand_int(mn, mx);
        if (mn == mx) return mn;
        int r = (int)(next() >> 33);
        return mn + r % (mx - mn);
    }
    bool flip_coin() { return ref ? ref->flip_coin() : uniform() < 0.5; }
    double normal() {
        if (ref) return ref->nd(ref->gen);          // cRand::RandDoubleNorm(0, 1)
        double u1 = 1.0 - uniform(), u2 = uniform();
        return std::sqrt(-2.0 * std::log(u1)) * std::cos(6.283185307179586476925 * u2);
    }
};

struct Scene {
    int char_type = 0, ctrl = 0, num_update_steps = 20, num_sim_substeps = 1, has_init_x = 0, terrain_type = 0;
    int n_terrain_sets = 0, has_net = 0, nj = 0, ndof = 0, n_ctrl = 0, n_actions = 0, default_action = 0;
    int grav_comp = 1, virt_forces = 1, tuple_buffer_size = 16;
    double gx = 0, gy = -9.8, init_x = 0, terrain_blend = 0, exp_rate = 0.1, exp_temp = 1, exp_base_rate = 0.01;
    std::vector<double> joints, bodies, pd, ctrl_params, actions, pose0, vel0, terrain_params, terrain_default;
    Skeleton sk;
    Net net;
    double target_vel_x = 4.0;
    PhysParams phys;
    bool is_raptor = false;
    int n_params = kDogParams, n_opt = kDogOptParams, misc_max = mMiscMax, sp_max = spMax;
    int opt_idx[kMaxParams];
    double exp_noise = 0.2;
    unsigned stumble_mask = 0, fall_mask = 0;   // parts whose contact counts as stumble / fall contact

    void load(const std::string& path) {
        Pack p = Pack::load(path);
        const auto& mi = p.i32("meta_i32");
        char_type = mi[0]; ctrl = mi[1]; num_update_steps = mi[2]; num_sim_substeps = mi[3]; has_init_x = mi[4];
        terrain_type = mi[5]; n_terrain_sets = mi[6]; has_net = mi[7]; nj = mi[8]; ndof = mi[9]; n_ctrl = mi[10];
        n_actions = mi[11]; default_action = mi[12]; grav_comp = mi[13]; virt_forces = mi[14]; tuple_buffer_size = mi[15];
        const auto& mf = p.f64("meta_f64");
        gx = mf[0]; gy = mf[1]; init_x = mf[2]; terrain_blend = mf[3]; exp_rate = mf[4]; exp_temp = mf[5]; exp_base_rate = mf[6];
        joints = p.f64("joints"); bodies = p.f64("bodies"); pd = p.f64("pd"); ctrl_params = p.f64("ctrl_params");
        actions = p.f64("actions"); pose0 = p.f64("pose0"); vel0 = p.f64("vel0");
        terrain_params = p.f64("terrain_params"); terrain_default = p.f64("terrain_default_params");
        sk.init(nj, joints.data(), bodies.data());
        if (sk.ndof != ndof) throw std::runtime_error("scene: dof mismatch");
        is_raptor = (char_type == 2);
        if (!((char_type == 1 && nj == jDogMax) || (is_raptor && nj == rRaptorMax)))
            throw std::runtime_error("oracle: unsupported character");
        if (has_net) net.load(p);
        target_vel_x = (ctrl == 4) ? 2.0 : 4.0;  // goat_mace (sim/GoatControllerMACE.cpp:11-14); raptor 4 (sim/RaptorController.cpp:596-600)
        if (is_raptor) {
            n_params = kRaptorParams; misc_max = rmMiscMax; sp_max = rpMax; n_opt = 0;
            for (int i = 0; i < n_params; ++i) if (kRaptorOptMask[i]) opt_idx[n_opt++] = i;
            exp_noise = 0.15;   // sim/RaptorControllerMACE.cpp:7
            for (int j = 0; j < nj; ++j) if (j != rRightToe && j != rLeftToe && j != rRightAnkle && j != rLeftAnkle) stumble_mask |= 1u << j;
            const int fp[6] = {rRoot, rSpine0, rSpine1, rSpine2, rSpine3, rHead};
            for (int j : fp) fall_mask |= 1u << j;
        } else {
            n_opt = 0;
            for (int i = 1; i < n_params; ++i) opt_idx[n_opt++] = i;
            for (int j = 0; j < nj; ++j) if (j != jToe && j != jFinger && j != jAnkle && j != jWrist) stumble_mask |= 1u << j;
            const int fp[9] = {jRoot, jSpine0, jSpine1, jSpine2, jSpine3, jTorso, jNeck0, jNeck1, jHead};
            for (int j : fp) fall_mask |= 1u << j;
        }
        if ((int)ctrl_params.size() != n_ctrl * n_params) throw std::runtime_error("scene: controller parameter size mismatch");
    }
    bool is_mace() const { return ctrl == 3 || ctrl == 4 || ctrl == 7; }
    // cScenarioSimChar::SetTerrainParamsLerp (scenarios/ScenarioSimChar.cpp:255-272)
    void terrain_params_lerp(double lerp, double* out) const {
        if (n_terrain_sets == 0) { for (int i = 0; i < pTerrainParamMax; ++i) out[i] = terrain_default[i]; return; }
        lerp = std::min(std::max(lerp, 0.0), n_terrain_sets - 1.0);
        int i0 = (int)lerp, i1 = std::min(i0 + 1, n_terrain_sets - 1);
        lerp -= i0;
        for (int i = 0; i < pTerrainParamMax; ++i)
            out[i] = (1 - lerp) * terrain_params[i0 * pTerrainParamMax + i] + lerp * terrain_params[i1 * pTerrainParamMax + i];
    }
};

struct Tuple {
    double reward = 0;
    unsigned flags = 0;
    std::vector<double> s_beg, action, s_end;
};

struct Action { int id = -1; double params[kMaxParams] = {0}; };

struct Env {
    const Scene* sc = nullptr;
    int env_id = 0;
    bool exp_mode = false;     // false: cScenarioPoliEval, true: cScenarioExpMACE
    bool enable_exp = false;   // controller mEnableExp
    double exp_rate = 0.2, exp_temp = 1, exp_base_rate = 0, exp_noise = 0.2;

    // simulation state
    double q[kMaxDof] = {0}, qd[kMaxDof] = {0}, tau_held[kMaxDof] = {0};
    bool contact[kMaxJoints] = {false};
    Ground ground;
    RBDModel ctrl_model, phys_model;
    CounterRng rng;

    // controller state
    int state = 0;
    int stance = 0;            // raptor: 0 = right leg is the stance leg (gDefaultStance), 1 = left
    double phase = 0;
    bool first_cycle = true, off_policy = false, exp_critic = false, exp_actor = false;
    Action cur;
    double pd_target[kMaxJoints] = {0};
    double cur_cycle_time = 0, prev_cycle_time = 0, cur_stumble = 0, prev_stumble = 0;
    double prev_com[2] = {0, 0}, prev_dist[2] = {0, 0};
    double ground_samples[kNumGroundSamples] = {0};
    double origin[2] = {0, 0};
    std::vector<double> poli_state;
    std::vector<int> commands;
    double last_tau[kMaxDof] = {0};   // controller output before the clamp (debug / parity probe)
    double last_net_out[128] = {0};

    // fall logic
    double fall_dist_counter = 5, fall_contact_counter = 0.1, sum_fall_contact = 0;
    double prev_check_pos[2] = {0, 0};
    bool fail_fall_dist = false;

    // scenario state
    double time = 0;
    int cycle_count = 0, episode_count = 0;
    double avg_dist = 0, pos_start_x = 0;
    std::vector<double> dist_log;
    Tuple cur_tuple;
    std::vector<Tuple> tuples;   // appended in record order (the batched boundary hands them over in bulk)
    int64_t total_steps = 0;

    // ------------------------------------------------------------------ kinematics helpers
    struct BodyKin { double px, py, ang, vx, vy, w; };
    BodyKin body[kMaxJoints];
    double jpx[kMaxJoints], jpy[kMaxJoints], jang[kMaxJoints], jvx[kMaxJoints], jvy[kMaxJoints], jw[kMaxJoints];

    // forward kinematics of joint frames and body COM frames (cKinTree::JointWorldTrans / BodyWorldTrans,
    // anim/KinTree.cpp:1050-1098) plus their velocities
    void update_kin() {
        const Skeleton& sk = sc->sk;
        for (int j = 0; j < sk.nj; ++j) {
            if (sk.parent[j] < 0) {
                jpx[j] = q[0]; jpy[j] = q[1]; jang[j] = q[2];
                jvx[j] = qd[0]; jvy[j] = qd[1]; jw[j] = qd[2];
            } else {
                int p = sk.parent[j];
                double c = std::cos(jang[p]), s = std::sin(jang[p]);
                double ax = c * sk.attach[j].x - s * sk.attach[j].y, ay = s * sk.attach[j].x + c * sk.attach[j].y;
                jpx[j] = jpx[p] + ax; jpy[j] = jpy[p] + ay;
                jang[j] = jang[p] + q[sk.offset[j]];
                jvx[j] = jvx[p] - jw[p] * ay; jvy[j] = jvy[p] + jw[p] * ax;
                jw[j] = jw[p] + qd[sk.offset[j]];
            }
            double c = std::cos(jang[j]), s = std::sin(jang[j]);
            double bx = c * sk.body_attach[j].x - s * sk.body_attach[j].y, by = s * sk.body_attach[j].x + c * sk.body_attach[j].y;
            body[j] = {jpx[j] + bx, jpy[j] + by, jang[j] + sk.body_theta[j], jvx[j] - jw[j] * by, jvy[j] + jw[j] * bx, jw[j]};
        }
    }
    void calc_com(double* com, double* com_vel) const {
        const Skeleton& sk = sc->sk;
        double m = 0, cx = 0, cy = 0, vx = 0, vy = 0;
        for (int j = 0; j < sk.nj; ++j) {
            if (!sk.valid_body(j)) continue;
            cx += sk.mass[j] * body[j].px; cy += sk.mass[j] * body[j].py;
            vx += sk.mass[j] * body[j].vx; vy += sk.mass[j] * body[j].vy;
            m += sk.mass[j];
        }
        if (com) { com[0] = cx / m; com[1] = cy / m; }
        if (com_vel) { com_vel[0] = vx / m; com_vel[1] = vy / m; }
    }
    // reduced-coordinate pose as cSimCharacter::BuildPose reports it: the root angle comes out of an axis-angle
    // extraction, i.e. wrapped to (-pi, pi].  Hinge angles are NOT wrapped: Bullet's hard limits keep them inside
    // (-pi, pi) in the reference, whereas this engine's compliant limits may overshoot 3.14 by ~1e-2 rad, and a
    // wrap there would flip the PD error sign (limit-less tail joints never approach pi in practice).
    void build_pose(double* pose) const {
        for (int k = 0; k < sc->ndof; ++k) pose[k] = q[k];
        pose[2] = wrap_pi(q[2]);
    }

    // ------------------------------------------------------------------ controller
    const double* cur_state_params() const { return cur.params + sc->misc_max + state * sc->sp_max; }
    // raptor leg joints for the current stance (sim/RaptorController.cpp:1489-1536)
    int r_stance(int k) const { return (stance == 0 ? rRightHip : rLeftHip) + k; }   // k: 0 hip, 1 knee, 2 ankle, 3 toe
    int r_swing(int k) const { return (stance == 0 ? rLeftHip : rRightHip) + k; }
    bool r_active_vf(int toe) const {   // cRaptorController::IsActiveVFEffector
        return toe == r_stance(3) && (state == rsContact || state == rsDown) && contact[toe];
    }
    // cDogController::SetStateParams (sim/DogController.cpp:1042-1054); cRaptorController::SetStateParams (:1108-1125)
    void set_state_params() {
        const double* p = cur_state_params();
        if (sc->is_raptor) {
            pd_target[r_stance(0)] = p[rpStanceHip]; pd_target[r_stance(1)] = p[rpStanceKnee]; pd_target[r_stance(2)] = p[rpStanceAnkle];
            pd_target[r_swing(0)] = p[rpSwingHip]; pd_target[r_swing(1)] = p[rpSwingKnee]; pd_target[r_swing(2)] = p[rpSwingAnkle];
            return;
        }
        const int spine[5] = {jSpine0, jSpine1, jSpine2, jSpine3, jTorso};
        for (int i = 0; i < 5; ++i) pd_target[spine[i]] = p[spSpineCurve];
        pd_target[jShoulder] = p[spShoulder]; pd_target[jElbow] = p[spElbow]; pd_target[jHip] = p[spHip];
        pd_target[jKnee] = p[spKnee]; pd_target[jAnkle] = p[spAnkle];
    }
    void transition_state(int s, double ph = 0) { state = s; phase = ph; set_state_params(); }
    void post_process(double* p) const {
        p[mTransTime] = std::abs(p[mTransTime]); p[mCv] = std::abs(p[mCv]);
        if (sc->is_raptor) p[rmCd] = std::abs(p[rmCd]);
    }
    // cDogController::BlendCtrlParams / BuildBaseAction; cDogControllerMACE::AssignFragID
    void build_base_action(int a, Action& out) {
        const double* act = &sc->actions[4 * a];
        int i0 = (int)act[0], i1 = (int)act[1];
        double blend = act[2];
        const int np = sc->n_params;
        for (int k = 0; k < np; ++k) {
            double p0 = sc->ctrl_params[i0 * np + k], p1 = sc->ctrl_params[i1 * np + k];
            // ReadParams stores post-processed parameter sets (sim/DogController.cpp:519-520)
            if (k == mTransTime || k == mCv || (sc->is_raptor && k == rmCd)) { p0 = std::abs(p0); p1 = std::abs(p1); }
            out.params[k] = (1 - blend) * p0 + blend * p1;
        }
        out.id = a;
        if (sc->is_mace()) {
            int nf = sc->has_net ? sc->net.n_frags : 0, frag = 0;
            if (nf > 0) {
                if (i0 >= nf && i1 >= nf) frag = rng.rand_int(0, nf);
                else if (i0 >= nf) frag = i1;
                else if (i1 >= nf) frag = i0;
                else {
                    frag = rng.flip_coin() ? i0 : i1;
                    int ncp = sc->n_ctrl, copies = nf / ncp, rem = nf % ncp;
                    if (frag < rem) ++copies;
                    frag += rng.rand_int(0, copies) * ncp;
                }
            }
            out.id = frag;
        }
    }
    // cDogController::NewCycleUpdate (sim/DogController.cpp:1329-1338)
    void ctrl_new_cycle_update() {
        prev_cycle_time = cur_cycle_time; cur_cycle_time = 0;
        prev_stumble = cur_stumble; cur_stumble = 0;
        double com[2];
        calc_com(com, nullptr);
        prev_dist[0] = com[0] - prev_com[0]; prev_dist[1] = com[1] - prev_com[1];
        prev_com[0] = com[0]; prev_com[1] = com[1];
    }
    // cTerrainRLCharController::ApplyAction + cDogController::ApplyAction
    void apply_action(const Action& a) {
        cur = a;
        post_process(cur.params);
        ctrl_new_cycle_update();
        transition_state(sBackStance);
    }
    bool has_fallen() const {
        return sum_fall_contact > 0.25 || fail_fall_dist || std::abs(wrap_pi(q[2])) > M_PI * 0.8;
    }
    bool check_contact(int j) const { return contact[j]; }
    unsigned contact_mask() const { unsigned m = 0; for (int j = 0; j < sc->nj; ++j) if (contact[j]) m |= 1u << j; return m; }
    bool has_stumbled() const { return (contact_mask() & sc->stumble_mask) != 0; }

    // cTerrainRLCharController::ParseGround + BuildPoliState (sim/TerrainRLCharController.cpp:168-285)
    void parse_ground_and_build_state() {
        origin[0] = q[0];
        origin[1] = ground.sample(q[0]);
        for (int i = 0; i < kNumGroundSamples; ++i) {
            double dist = ((10.0 - (-0.5)) * i) / (kNumGroundSamples - 1) + (-0.5);
            ground_samples[i] = ground.sample(dist + origin[0]) - origin[1];
        }
        int nb = sc->nj;
        poli_state.assign(kNumGroundSamples + (2 * nb - 1) + 2 * nb, 0.0);
        for (int i = 0; i < kNumGroundSamples; ++i) poli_state[i] = ground_samples[i];
        int idx = kNumGroundSamples;
        poli_state[idx++] = q[1] - ground.sample(q[0]);
        for (int i = 1; i < nb; ++i) { poli_state[idx++] = body[i].px - q[0]; poli_state[idx++] = body[i].py - q[1]; }
        for (int i = 0; i < nb; ++i) { poli_state[idx++] = body[i].vx; poli_state[idx++] = body[i].vy; }
        if (sc->is_raptor && stance != 0) {
            // cRaptorController::FlipPoliPoseStance: swap the two legs' entries (packed at the end of each block)
            const int nleg = 4 * 2;
            int pose_end = kNumGroundSamples + 2 * nb - 1, vel_end = pose_end + 2 * nb;
            for (int i = 0; i < nleg; ++i) {
                std::swap(poli_state[pose_end - 1 - i], poli_state[pose_end - nleg - 1 - i]);
                std::swap(poli_state[vel_end - 1 - i], poli_state[vel_end - nleg - 1 - i]);
            }
        }
    }

    // cBaseControllerMACE::BuildActorAction (+ cDogController::SetOptParams)
    void build_actor_action(const double* y, int a, Action& out) {
        out.id = a;
        for (int k = 0; k < sc->n_params; ++k) out.params[k] = cur.params[k];
        int nf = sc->net.n_frags, fs = sc->net.frag;
        for (int k = 0; k < fs; ++k) out.params[sc->opt_idx[k]] = y[nf + a * fs + k];
        post_process(out.params);
    }
    // cBaseControllerMACE::DecideActionBoltzmann (sim/BaseControllerMACE.cpp:254-318)
    void decide_action(Action& out) {
        off_policy = false;
        double base_rand = rng.uniform();
        if (enable_exp && base_rand < exp_base_rate) {
            int a = rng.rand_int(0, sc->n_actions);
            build_base_action(a, out);
            off_policy = true; exp_actor = true; exp_critic = true;
            return;
        }
        const Net& net = sc->net;
        double* y = last_net_out;
        net.eval(poli_state.data(), y);
        int nf = net.n_frags;
        int a_max = 0;
        for (int i = 1; i < nf; ++i) if (y[i] > y[a_max]) a_max = i;
        int a = a_max;
        if (enable_exp && exp_temp != 0) {  // BoltzmannSelectActor
            double vals[8], sum = 0;
            for (int i = 0; i < nf; ++i) { vals[i] = std::exp((y[i] - y[a_max]) / exp_temp); sum += vals[i]; }
            double r = rng.uniform() * sum;
            for (int i = 0; i < nf; ++i) { r -= vals[i]; if (r <= 0) { a = i; break; } }
        }
        build_actor_action(y, a, out);
        if (enable_exp) {
            double rn = rng.uniform();
            if (rn < exp_rate) {  // ApplyExpNoiseAction: N(0, mExpNoise) / OutputScale of actor 0
                for (int k = 0; k < net.frag; ++k) {
                    double noise = exp_noise * rng.normal();
                    out.params[sc->opt_idx[k]] += noise * (1.0 / net.out_scale[nf + k]);
                }
                exp_actor = true;
            }
            exp_critic = (a != a_max);
            off_policy = exp_actor || exp_critic;
        }
    }
    // cDogControllerMACE::UpdateAction / cDogController::UpdateAction (sim/DogController.cpp:847-868)
    void update_action() {
        exp_actor = false; exp_critic = false;
        parse_ground_and_build_state();
        off_policy = true;
        Action next = cur;
        if (!commands.empty()) {
            if (sc->is_mace()) { exp_actor = true; exp_critic = true; }
            int a = commands.back(); commands.pop_back();
            build_base_action(a, next);
        } else if (sc->has_net) {
            decide_action(next);
        } else {
            bool cyclic = sc->is_mace() ? false : (sc->actions[4 * cur.id + 3] != 0);
            if (!cyclic) build_base_action(sc->default_action, next);
        }
        apply_action(next);
    }
    // cDogController::UpdateState (sim/DogController.cpp:805-845) with the state table (:12-38)
    void update_state(double h) {
        if (sc->is_raptor) {   // cRaptorController::UpdateState (sim/RaptorController.cpp:804-849), state table :11-37
            bool advance = first_cycle;
            phase += h / cur.params[rmTransTime];
            if (state != rsUp && phase >= 1) advance = true;
            if (state == rsUp && contact[r_swing(3)]) advance = true;
            if (advance) {
                int ns = first_cycle ? rsContact : (state == rsUp ? -1 : state + 1);
                bool end_step = (ns < 0) || first_cycle;
                if (end_step) {
                    if (!first_cycle) { stance = 1 - stance; set_state_params(); }   // FlipStance -> SetStance
                    update_action();
                    first_cycle = false;
                } else transition_state(ns);
            }
            return;
        }
        static const bool trans_time[4] = {true, false, true, false};
        static const int trans_contact[4] = {-1, jFinger, -1, jToe};
        static const int next_state[4] = {sExtend, sFrontStance, sGather, -1};
        bool advance = first_cycle;
        phase += h / cur.params[mTransTime];
        if (trans_time[state] && phase >= 1) advance = true;
        if (trans_contact[state] >= 0 && check_contact(trans_contact[state])) advance = true;
        if (advance) {
            int ns = first_cycle ? sBackStance : next_state[state];
            bool end_step = (ns < 0) || first_cycle;
            if (end_step) { update_action(); first_cycle = false; }
            else transition_state(ns);
        }
    }
    bool is_new_cycle() const { return state == 0 && phase == 0; }

    // world position of the bottom-centre of a foot box (cDogController::GetEndEffectorContactPos)
    void effector_pos(int j, double* p) const {
        double c = std::cos(body[j].ang), s = std::sin(body[j].ang), ly = -0.5 * sc->sk.body_size[j].y;
        p[0] = body[j].px - s * ly; p[1] = body[j].py + c * ly;
    }

    // cDogController::Update (active mode) -> generalised forces tau[ndof]
    void controller_update(double h, double* tau) {
        const Skeleton& sk = sc->sk;
        int nd = sk.ndof;
        for (int k = 0; k < nd; ++k) tau[k] = 0;
        cur_cycle_time += h;
        if (has_stumbled()) cur_stumble += h;

        // UpdateRBDModel
        double pose[kMaxDof], vel[kMaxDof];
        build_pose(pose);
        for (int k = 0; k < nd; ++k) vel[k] = qd[k];
        ctrl_model.update(pose, vel);

        update_state(h);
        if (sc->is_raptor) { raptor_controller_tail(h, pose, vel, tau); return; }

        // ApplyFeedback (sim/DogController.cpp:903-945)
        {
            double cv[2];
            calc_com(nullptr, cv);
            const int jt[2] = {jHip, jShoulder}, ef[2] = {jToe, jFinger}, pr[2] = {spHip, spShoulder};
            for (int i = 0; i < 2; ++i)
                if (!contact[ef[i]]) pd_target[jt[i]] = cur_state_params()[pr[i]] + cv[0] * cur.params[mCv];
        }

        // cImpPDController::CalcControlForces (sim/ImpPDController.cpp:234-278)
        {
            double kp[kMaxDof] = {0}, kd[kMaxDof] = {0}, perr[kMaxDof] = {0}, verr[kMaxDof] = {0};
            for (int j = 1; j < sk.nj; ++j) {
                int o = sk.offset[j];
                const double* pdj = &sc->pd[6 * j];
                kp[o] = pdj[0]; kd[o] = pdj[1];
                bool world = pdj[5] != 0;
                // cPDController::CalcTheta: world-coordinate joints measure the child body's world rotation
                double theta = world ? wrap_pi(body[j].ang) : pose[o];
                perr[o] = pd_target[j] - theta;
                verr[o] = pdj[4] - qd[o];
            }
            static thread_local double A[kMaxDof][kMaxDof];
            double rhs[kMaxDof], acc[kMaxDof];
            for (int a = 0; a < nd; ++a) {
                for (int b = 0; b < nd; ++b) A[a][b] = ctrl_model.M[a][b];
                A[a][a] += h * kd[a];
                rhs[a] = kp[a] * (perr[a] - h * vel[a]) + kd[a] * verr[a] - ctrl_model.C[a];
            }
            ldlt_solve(nd, A, rhs, acc);
            for (int a = 0; a < nd; ++a) tau[a] += kp[a] * (perr[a] - h * vel[a]) + kd[a] * (verr[a] - h * acc[a]);
        }

        // ApplyGravityCompensation (sim/DogController.cpp:947-995) + BuildContactBasis (:1120-1175)
        if (sc->grav_comp) {
            const int eff[2] = {jToe, jFinger};
            double basis[kMaxDof][4];
            for (int a = 0; a < nd; ++a) for (int b = 0; b < 4; ++b) basis[a][b] = 0;
            bool support = false;
            for (int e = 0; e < 2; ++e) {
                if (!contact[eff[e]]) continue;
                support = true;
                double p[2];
                effector_pos(eff[e], p);
                SpTrans X; X.r = {-p[0], -p[1], 0};
                SV fb[2] = {apply_F(X, SV{{0, 0, 0}, {0, 1, 0}}), apply_F(X, SV{{0, 0, 0}, {1, 0, 0}})};
                for (int cur_j = eff[e]; cur_j >= 0; cur_j = sk.parent[cur_j]) {
                    int o = sk.offset[cur_j];
                    for (int k = 0; k < sk.size[cur_j]; ++k)
                        for (int b = 0; b < 2; ++b) basis[o + k][2 * e + b] = dot(ctrl_model.J[o + k], fb[b]);
                }
            }
            if (support) {
                double tg[kMaxDof];
                ctrl_model.gravity_force(tg);
                for (int a = 0; a < nd; ++a) tg[a] = -tg[a];
                double AtA[4][4], Atb[4], x[4];
                for (int a = 0; a < 4; ++a) {
                    Atb[a] = 0;
                    for (int r = 0; r < 3; ++r) Atb[a] += basis[r][a] * tg[r];
                    for (int b = 0; b < 4; ++b) {
                        AtA[a][b] = 0;
                        for (int r = 0; r < 3; ++r) AtA[a][b] += basis[r][a] * basis[r][b];
                    }
                    AtA[a][a] += 0.0001;
                }
                solve4(AtA, Atb, x);
                for (int a = 0; a < nd; ++a) {
                    double tc = 0;
                    for (int b = 0; b < 4; ++b) tc += basis[a][b] * x[b];
                    tg[a] -= tc;
                }
                tg[0] = tg[1] = tg[2] = 0;
                for (int a = 0; a < nd; ++a) tau[a] += tg[a];
            }
        }

        // ApplyVirtualForces (sim/DogController.cpp:997-1029)
        if (sc->virt_forces) {
            const int eff[2] = {jToe, jFinger};
            for (int e = 0; e < 2; ++e) {
                int j = eff[e];
                bool valid = ((state == sBackStance || state == sExtend) && j == jToe) ||
                             ((state == sFrontStance || state == sGather) && j == jFinger);
                if (!(valid && contact[j])) continue;
                double fx = (j == jToe) ? cur.params[mBackForceX] : cur.params[mFrontForceX];
                double fy = (j == jToe) ? cur.params[mBackForceY] : cur.params[mFrontForceY];
                double p[2];
                effector_pos(j, p);
                SpTrans X; X.r = {-p[0], -p[1], 0};
                SV f = apply_F(X, SV{{0, 0, 0}, {-fx, -fy, 0}});
                for (int cur_j = j; cur_j != jRoot && cur_j != jTorso; cur_j = sk.parent[cur_j]) {
                    int o = sk.offset[cur_j];
                    tau[o] += dot(ctrl_model.J[o], f);
                }
            }
        }
    }
    // cRaptorController::Update after UpdateState (sim/RaptorController.cpp:195-233): UpdateStanceHip,
    // ApplySwingFeedback, stable-PD with the stance hip switched off while its foot pushes, gravity compensation
    // (weighted ridge LS, root rows kept), ApplyStanceFeedback, ApplyVirtualForces
    void raptor_controller_tail(double h, const double* pose, const double* vel, double* tau) {
        const Skeleton& sk = sc->sk;
        const int nd = sk.ndof;
        const int st_hip = r_stance(0), sw_hip = r_swing(0), st_toe = r_stance(3);
        const bool active_vf = r_active_vf(st_toe);
        // ApplySwingFeedback (:907-931)
        {
            bool first_half = state == rsContact || state == rsDown;
            double cd = first_half ? 0 : cur.params[rmCd], cv = first_half ? cur.params[rmCv] : 0;
            double com[2], cvel[2];
            calc_com(com, cvel);
            double dth = cd * (com[0] - body[st_toe].px) + cv * cvel[0];
            pd_target[sw_hip] = cur_state_params()[rpSwingHip] + dth;
        }
        // stable PD (stance hip inactive while it is an active virtual-force effector, :899-905)
        {
            double kp[kMaxDof] = {0}, kd[kMaxDof] = {0}, kdm[kMaxDof] = {0}, perr[kMaxDof] = {0}, verr[kMaxDof] = {0};
            for (int j = 1; j < sk.nj; ++j) {
                int o = sk.offset[j];
                const double* pdj = &sc->pd[6 * j];
                bool active = !(j == st_hip && active_vf);
                kdm[o] = pdj[1];
                kp[o] = active ? pdj[0] : 0; kd[o] = active ? pdj[1] : 0;
                bool world = pdj[5] != 0;
                double theta = world ? wrap_pi(body[j].ang) : pose[o];
                perr[o] = pd_target[j] - theta;
                verr[o] = pdj[4] - qd[o];
            }
            static thread_local double A[kMaxDof][kMaxDof];
            double rhs[kMaxDof], acc[kMaxDof];
            for (int a = 0; a < nd; ++a) {
                for (int b = 0; b < nd; ++b) A[a][b] = ctrl_model.M[a][b];
                A[a][a] += h * kdm[a];
                rhs[a] = kp[a] * (perr[a] - h * vel[a]) + kd[a] * verr[a] - ctrl_model.C[a];
            }
            ldlt_solve(nd, A, rhs, acc);
            for (int a = 0; a < nd; ++a) tau[a] += kp[a] * (perr[a] - h * vel[a]) + kd[a] * (verr[a] - h * acc[a]);
        }
        const int eff[2] = {rRightToe, rLeftToe};
        // ApplyGravityCompensation (:983-1026) + BuildContactBasis (:1168-1230)
        if (sc->grav_comp) {
            double basis[kMaxDof][4];
            for (int a = 0; a < nd; ++a) for (int b = 0; b < 4; ++b) basis[a][b] = 0;
            bool support = false;
            for (int e = 0; e < 2; ++e) {
                if (!r_active_vf(eff[e])) continue;
                support = true;
                double p[2];
                effector_pos(eff[e], p);
                SpTrans X; X.r = {-p[0], -p[1], 0};
                SV fb[2] = {apply_F(X, SV{{0, 0, 0}, {0, 1, 0}}), apply_F(X, SV{{0, 0, 0}, {1, 0, 0}})};
                for (int cur_j = eff[e]; cur_j >= 0; cur_j = sk.parent[cur_j]) {
                    int o = sk.offset[cur_j];
                    for (int k = 0; k < sk.size[cur_j]; ++k)
                        for (int b = 0; b < 2; ++b) basis[o + k][2 * e + b] = dot(ctrl_model.J[o + k], fb[b]);
                }
            }
            if (support) {
                double tg[kMaxDof];
                ctrl_model.gravity_force(tg);
                for (int a = 0; a < nd; ++a) tg[a] = -tg[a];
                const double W[3] = {0.0001, 0.0001, 1};
                double AtA[4][4], Atb[4], x[4];
                for (int a = 0; a < 4; ++a) {
                    Atb[a] = 0;
                    for (int r = 0; r < 3; ++r) Atb[a] += basis[r][a] * W[r] * tg[r];
                    for (int b = 0; b < 4; ++b) {
                        AtA[a][b] = 0;
                        for (int r = 0; r < 3; ++r) AtA[a][b] += basis[r][a] * W[r] * basis[r][b];
                    }
                    AtA[a][a] += 0.0001;
                }
                solve4(AtA, Atb, x);
                for (int a = 0; a < nd; ++a) {
                    double tc = 0;
                    for (int b = 0; b < 4; ++b) tc += basis[a][b] * x[b];
                    tau[a] += tg[a] - tc;
                }
            }
        }
        // ApplyStanceFeedback (:933-981)
        if (active_vf) {
            double hip_tau = -tau[sk.offset[sw_hip]];
            const double* pdh = &sc->pd[6 * st_hip];
            double root_tau = pdh[0] * (cur_state_params()[rpRootPitch] - wrap_pi(q[2])) + pdh[1] * (-qd[2]);
            hip_tau += -root_tau;
            tau[sk.offset[st_hip]] += hip_tau;
        }
        // ApplyVirtualForces (:1028-1075)
        if (sc->virt_forces) {
            for (int e = 0; e < 2; ++e) {
                int j = eff[e];
                if (!r_active_vf(j)) continue;
                double p[2];
                effector_pos(j, p);
                SpTrans X; X.r = {-p[0], -p[1], 0};
                SV f = apply_F(X, SV{{0, 0, 0}, {-cur.params[rmForceX], -cur.params[rmForceY], 0}});
                for (int cur_j = j; cur_j != rRoot; cur_j = sk.parent[cur_j]) {
                    int o = sk.offset[cur_j];
                    double t = dot(ctrl_model.J[o], f);
                    tau[o] += t;
                    if (cur_j == st_hip) tau[sk.offset[sw_hip]] += -t;
                }
            }
        }
    }
    // 4x4 linear solve with partial pivoting (stands in for Eigen householderQr().solve on the SPD A^T A + lambda I)
    static void solve4(double A[4][4], double* b, double* x) {
        double M[4][5];
        for (int i = 0; i < 4; ++i) { for (int j = 0; j < 4; ++j) M[i][j] = A[i][j]; M[i][4] = b[i]; }
        for (int c = 0; c < 4; ++c) {
            int piv = c;
            for (int r = c + 1; r < 4; ++r) if (std::abs(M[r][c]) > std::abs(M[piv][c])) piv = r;
            if (piv != c) for (int k = 0; k < 5; ++k) std::swap(M[c][k], M[piv][k]);
            for (int r = c + 1; r < 4; ++r) {
                double f = M[r][c] / M[c][c];
                for (int k = c; k < 5; ++k) M[r][k] -= f * M[c][k];
            }
        }
        for (int r = 3; r >= 0; --r) {
            double v = M[r][4];
            for (int k = r + 1; k < 4; ++k) v -= M[r][k] * x[k];
            x[r] = v / M[r][r];
        }
    }

    // cDogController::CalcReward (sim/DogController.cpp:594-623)
    // cSimCharSoftFall::UpdateFallDistCheck / UpdateFallContactCheck (sim/SimCharSoftFall.cpp:74-125)
    void update_fall_checks(double h) {
        fall_dist_counter -= h;
        if (fall_dist_counter <= 0) {
            double dx = q[0] - prev_check_pos[0], dy = q[1] - prev_check_pos[1];
            if (dx * dx + dy * dy < 0.5 * 0.5) fail_fall_dist = true;
            prev_check_pos[0] = q[0]; prev_check_pos[1] = q[1];
            fall_dist_counter = 5;
        }
        fall_contact_counter -= h;
        if (fall_contact_counter <= 0) {
            bool hc = (contact_mask() & sc->fall_mask) != 0;
            const double norm = (1 + 1 / (1 - 0.9));
            sum_fall_contact = (hc ? 1.0 : 0.0) / norm + 0.9 * sum_fall_contact;
            fall_contact_counter = 0.1;
        }
    }

    double calc_reward() const {
        double vel_r = 0, stum_r = 0;
        if (!has_fallen()) {
            double avg_vel = prev_dist[0] / prev_cycle_time;
            double err = sc->target_vel_x - avg_vel;
            vel_r = std::exp(-0.5 * err * err);
            double avg_st = prev_stumble / prev_cycle_time;
            stum_r = 1.0 / (1 + 10 * avg_st);
            if (sc->is_raptor && avg_vel < 0) { vel_r = 0; stum_r = 0; }   // sim/RaptorController.cpp:583-587
        }
        return 0.8 * vel_r + 0.2 * stum_r;
    }

    // ------------------------------------------------------------------ physics (this project's model)
    struct ContactPoint { int body; double px, py; };
    double last_qdd[kMaxDof] = {0};

    // one compliant contact on body i at world point (px, py): penetration `pen` along the unit direction (nx, ny) the force pushes
    // the body in.  Implicit in the point velocity: adds dt J^T D J to A and J^T w to rhs, D = cnn n n^T + ctt t t^T.
    void add_contact(int i, double px, double py, double pen, double nx, double ny, double dt, const SV* avp, double (*A)[kMaxDof], double* rhs) {
        const Skeleton& sk = sc->sk;
        const PhysParams& pp = sc->phys;
        const int nd = sk.ndof;
        const double tx = ny, ty = -nx;
        // point Jacobian rows along the chain, point velocity, velocity-product acceleration at the point
        double Jx[kMaxDof] = {0}, Jy[kMaxDof] = {0};
        double vx = 0, vy = 0;
        for (int cj = i; cj >= 0; cj = sk.parent[cj]) {
            int o = sk.offset[cj];
            for (int k = 0; k < sk.size[cj]; ++k) {
                const SV& col = phys_model.J[o + k];
                Jx[o + k] = col.v.x - col.o.z * py;
                Jy[o + k] = col.v.y + col.o.z * px;
                vx += Jx[o + k] * qd[o + k];
                vy += Jy[o + k] * qd[o + k];
            }
        }
        double ax = avp[i].v.x - avp[i].o.z * py, ay = avp[i].v.y + avp[i].o.z * px;
        double vn = vx * nx + vy * ny, vt = vx * tx + vy * ty;
        double fn0 = pp.kn * pen - pp.dn * vn;
        if (fn0 <= 0) return;
        double cnn = pp.dn + dt * pp.kn;
        double ctt = pp.mu * fn0 / std::max(std::abs(vt), pp.v_eps);
        // D = cnn n n^T + ctt t t^T ; w = F0 - D (v + dt a_vp)
        double ux = vx + dt * ax, uy = vy + dt * ay;
        double un = ux * nx + uy * ny, ut = ux * tx + uy * ty;
        double wx = pp.kn * pen * nx - cnn * un * nx - ctt * ut * tx;
        double wy = pp.kn * pen * ny - cnn * un * ny - ctt * ut * ty;
        for (int a = 0; a < nd; ++a) {
            if (Jx[a] == 0 && Jy[a] == 0) continue;
            rhs[a] += Jx[a] * wx + Jy[a] * wy;
            double jan = Jx[a] * nx + Jy[a] * ny, jat = Jx[a] * tx + Jy[a] * ty;
            for (int b = 0; b < nd; ++b) {
                double jbn = Jx[b] * nx + Jy[b] * ny, jbt = Jx[b] * tx + Jy[b] * ty;
                A[a][b] += dt * (cnn * jan * jbn + ctt * jat * jbt);
            }
        }
    }

    // forward dynamics with implicit contact / joint-limit terms; returns qdd, optionally refreshes contact bits
    void forward_dynamics(const double* tau, double dt, double* qdd, bool write_contacts) {
        const Skeleton& sk = sc->sk;
        const PhysParams& pp = sc->phys;
        int nd = sk.ndof;
        phys_model.update_kinematics(q, qd);
        phys_model.build_mass_mat();
        phys_model.build_jacobian();
        double zero[kMaxDof] = {0}, Cb[kMaxDof], dummy[kMaxDof];
        // velocity-product accelerations per link (no gravity), link coordinates
        phys_model.inv_dyna(zero, V3{0, 0, 0}, false, dummy);
        SV avp[kMaxJoints];
        for (int j = 0; j < sk.nj; ++j) avp[j] = apply_inv_M(phys_model.sp_world_joint[j], phys_model.link_acc[j]);
        phys_model.inv_dyna(zero, -phys_model.gravity, false, Cb);

        static thread_local double A[kMaxDof][kMaxDof];
        double rhs[kMaxDof];
        for (int a = 0; a < nd; ++a) {
            for (int b = 0; b < nd; ++b) A[a][b] = phys_model.M[a][b];
            rhs[a] = tau[a] - Cb[a];
        }
        update_kin();
        if (write_contacts) for (int j = 0; j < sk.nj; ++j) contact[j] = false;
        for (int i = 0; i < sk.nj; ++i) {
            if (!collidable(i)) continue;
            double hx = 0.5 * sk.body_size[i].x, hy = 0.5 * sk.body_size[i].y;
            double c = std::cos(body[i].ang), s = std::sin(body[i].ang);
            for (int cn = 0; cn < 4; ++cn) {
                double lx = (cn & 1) ? hx : -hx, ly = (cn & 2) ? hy : -hy;
                double px = body[i].px + c * lx - s * ly, py = body[i].py + s * lx + c * ly;
                double slope = 0;
                double hgt = ground.sample(px, &slope);
                double inv = 1.0 / std::sqrt(1.0 + slope * slope);
                double pen = (hgt - py) * inv;
                if (pen <= -pp.contact_tol) continue;
                if (write_contacts) contact[i] = true;
                if (pen <= 0) continue;
                add_contact(i, px, py, pen, -slope * inv, inv, dt, avp, A, rhs);
            }
            if (!pp.vertex_contacts) continue;
            // terrain vertices inside the box (sim/GroundVar2D.cpp:440-448 hands Bullet a height field: an edge of the terrain can
            // enter a box between two of its corners).  Only vertices where the surface is convex can do that.  The vertex is
            // pushed out through the nearest face: depth = distance to that face, force on the box along the opposite direction.
            const double ext = std::abs(c) * hx + std::abs(s) * hy;
            ground.for_vertices(body[i].px - ext - pp.contact_tol, body[i].px + ext + pp.contact_tol, [&](double vx_, double vh, double hp, double hn) {
                if (!(vh - 0.5 * (hp + hn) > 1e-9)) return;
                const double rx = vx_ - body[i].px, ry = vh - body[i].py;
                const double lx = c * rx + s * ry, ly = -s * rx + c * ry;
                const double dx = hx - std::abs(lx), dy = hy - std::abs(ly);
                if (dx <= -pp.contact_tol || dy <= -pp.contact_tol) return;
                if (write_contacts) contact[i] = true;
                if (dx <= 0 || dy <= 0) return;
                double nlx = 0, nly = 0, pen;
                if (dx < dy) { nlx = lx > 0 ? -1.0 : 1.0; pen = dx; } else { nly = ly > 0 ? -1.0 : 1.0; pen = dy; }
                add_contact(i, vx_, vh, pen, c * nlx - s * nly, s * nlx + c * nly, dt, avp, A, rhs);
            });
        }
        // joint limits: one-sided implicit spring-damper (limits [1,0] mean "none", sim/World.cpp:28-29)
        for (int j = 1; j < sk.nj; ++j) {
            if (!(sk.lim_lo[j] <= sk.lim_hi[j])) continue;
            int o = sk.offset[j];
            double viol = 0;
            if (q[o] > sk.lim_hi[j]) viol = q[o] - sk.lim_hi[j];
            else if (q[o] < sk.lim_lo[j]) viol = q[o] - sk.lim_lo[j];
            if (viol == 0) continue;
            double cl = pp.d_lim + dt * pp.k_lim;
            rhs[o] += -pp.k_lim * viol - cl * qd[o];
            A[o][o] += dt * cl;
        }
        ldlt_solve(nd, A, rhs, qdd);
        for (int a = 0; a < nd; ++a) last_qdd[a] = qdd[a];
    }
    bool collidable(int j) const {
        // tail parts carry collision group "none" (sim/SimDog.cpp:7,21-24)
        if (sc->is_raptor) return sc->sk.valid_body(j);   // every raptor part collides (sim/SimRaptor.cpp:4-27)
        return sc->sk.valid_body(j) && !(j >= jTail0 && j <= jTail3);
    }
    void physics_substep(double dt, bool last) {
        double qdd[kMaxDof];
        forward_dynamics(tau_held, dt, qdd, last);
        for (int k = 0; k < sc->ndof; ++k) { qd[k] += dt * qdd[k]; q[k] += dt * qd[k]; }
    }

    // ------------------------------------------------------------------ scenario
    void init(const Scene* s, int id, bool exp, uint64_t terrain_seed, uint64_t rng_seed) {
        sc = s; env_id = id; exp_mode = exp;
        enable_exp = exp;
        exp_rate = s->exp_rate; exp_temp = s->exp_temp; exp_base_rate = s->exp_base_rate;
        V3 g{s->gx, s->gy, 0};
        ctrl_model.init(&s->sk, g);
        phys_model.init(&s->sk, g);
        ground.type = s->terrain_type;
        s->terrain_params_lerp(s->terrain_blend, ground.params);
        ground.rand.seed((unsigned long)terrain_seed);
        rng.seed(rng_seed, (uint64_t)id);
        for (int j = 0; j < s->nj; ++j) pd_target[j] = s->pd[6 * j + 3];
        const int S = kNumGroundSamples + 4 * s->nj - 1;
        cur_tuple.s_beg.assign(S, 0.0); cur_tuple.s_end.assign(S, 0.0); cur_tuple.action.assign(1 + s->n_opt, 0.0);
        exp_noise = s->exp_noise;
        poli_state.assign(kNumGroundSamples + 4 * s->nj - 1, 0.0);
        reset();
    }

    // cScenarioSimChar::Reset (+ cScenarioPoliEval::Reset / cScenarioExp::Reset)
    void reset() {
        const Skeleton& sk = sc->sk;
        time = 0;
        for (int k = 0; k < sk.ndof; ++k) { q[k] = sc->pose0[k]; qd[k] = sc->vel0[k]; tau_held[k] = 0; }
        for (int j = 0; j < sk.nj; ++j) contact[j] = false;
        update_kin();
        // controller Reset: cBaseControllerMACE::Reset, cTerrainRLCharController::Reset, cDogController::Reset
        // (cRaptorController::Reset additionally restores the default stance, sim/RaptorController.cpp:175-182,663-675)
        exp_critic = false; exp_actor = false;
        stance = 0;
        Action a;
        build_base_action(sc->default_action, a);
        apply_action(a);
        transition_state(0, 0);
        // cDogController::ResetParams
        phase = 0; first_cycle = true; off_policy = false; origin[0] = origin[1] = 0;
        state = sBackStance; prev_cycle_time = 0; prev_dist[0] = prev_dist[1] = 0; cur_cycle_time = 0;
        prev_stumble = 0; cur_stumble = 0;
        for (int i = 0; i < kNumGroundSamples; ++i) ground_samples[i] = 0;
        commands.clear();
        calc_com(prev_com, nullptr);
        // cSimCharSoftFall::Reset
        fall_dist_counter = 5; prev_check_pos[0] = q[0]; prev_check_pos[1] = q[1]; fail_fall_dist = false;
        fall_contact_counter = 0.1; sum_fall_contact = 0;
        // ResetGround, InitCharacterPos
        ground.clear();
        ground.update(-10.0 + -1.0, 10.0 + -1.0);
        if (sc->has_init_x) q[0] = sc->init_x;
        q[1] += ground.sample(q[0]);
        update_kin();
        if (exp_mode) {
            cycle_count = 0;
            commands.push_back(rng.rand_int(0, sc->n_actions));  // CommandRandAction
        } else {
            pos_start_x = q[0];
        }
    }

    // cScenarioExp::NewCycleUpdate (scenarios/ScenarioExp.cpp:209-243)
    void exp_new_cycle_update() {
        cur_tuple.s_end = poli_state;
        bool fail = has_fallen();
        cur_tuple.flags = fail ? (cur_tuple.flags | (1u << fFail)) : (cur_tuple.flags & ~(1u << fFail));
        cur_tuple.reward = calc_reward();
        if (cycle_count > 1) tuples.push_back(cur_tuple);
        cur_tuple.s_beg = cur_tuple.s_end;
        cur_tuple.action.assign(1 + sc->n_opt, 0.0);   // RecordPoliAction
        cur_tuple.action[0] = cur.id;
        for (int k = 0; k < sc->n_opt; ++k) cur_tuple.action[1 + k] = cur.params[sc->opt_idx[k]];
        cur_tuple.flags = 0;
        if (exp_critic) cur_tuple.flags |= (1u << fExpCritic);
        if (exp_actor) cur_tuple.flags |= (1u << fExpActor);
        ++cycle_count;
    }

    // one iteration of the loop at scenarios/ScenarioSimChar.cpp:162-173 (= 1 env-step)
    void env_step(double h) {
        const Skeleton& sk = sc->sk;
        int ns = sc->num_sim_substeps;
        double dt = h / ns;
        for (int s = 0; s < ns; ++s) physics_substep(dt, s == ns - 1);       // UpdateWorld
        ground.update(q[0] - 2.0, q[0] + 10.0 + 1.0);                         // UpdateGround
        update_kin();
        double tau[kMaxDof];
        controller_update(h, tau);                                            // UpdateCharacter
        for (int k = 0; k < sk.ndof; ++k) last_tau[k] = tau[k];
        tau_held[0] = tau_held[1] = tau_held[2] = 0;
        for (int j = 1; j < sk.nj; ++j) {                                     // cJoint::ApplyTorque clamp
            int o = sk.offset[j];
            double t = tau[o], lim = sc->pd[6 * j + 2];
            if (std::abs(t) > lim) t *= lim / std::abs(t);
            tau_held[o] = t;
        }
        update_fall_checks(h);
        if (is_new_cycle()) {                                                 // PostSubstepUpdate
            if (exp_mode) exp_new_cycle_update();
            else ++cycle_count;
        }
        ++total_steps;
    }

    // cScenarioPoliEval::Update / cScenarioExp::Update for one outer step of `dt` seconds
    void update(double dt) {
        time += dt;
        double h = dt / sc->num_update_steps;
        for (int i = 0; i < sc->num_update_steps; ++i) env_step(h);
        end_update();
    }

    // the tail of cScenarioPoliEval::Update / cScenarioExp::Update, after the env-steps (scenarios/ScenarioPoliEval.cpp:112-124,
    // scenarios/ScenarioExp.cpp:84-97); separate so that a test can interleave the env-steps with the compiled reference scenario
    void end_update() {
        if (exp_mode) {
            if (!is_new_cycle() && has_fallen()) { exp_new_cycle_update(); reset(); }
        } else if (has_fallen()) {
            if (cycle_count >= 1) {   // IsValidCycle; RecordDistTraveled
                double dist = q[0] - pos_start_x;
                avg_dist = (episode_count * avg_dist + dist) / (episode_count + 1.0);
                ++episode_count;
                dist_log.push_back(dist);
            }
            reset();
        }
    }
};

}  // namespace orc
