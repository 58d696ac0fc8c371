// TEST INFRASTRUCTURE ONLY -- C entry points of the CPU oracle for ctypes (tests/, smoke(), bench.py cpu_baseline).
#include <cstring>
#include <memory>
#include <thread>
#include <pthread.h>
#include <sched.h>
#include <vector>

#include "env.h"
#include "trainer.h"

using namespace orc;

struct OrcBatch {
    Scene scene;
    std::vector<std::unique_ptr<Env>> envs;
    std::string err;
    Rand ref_rand;      // see orc_use_ref_rand
};

static thread_local std::string g_err;

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

// mode: 0 = policy evaluation (cScenarioPoliEval), 1 = exploration (cScenarioExpMACE)
OrcBatch* orc_create(const char* pack_path, int num_envs, int mode, const uint64_t* terrain_seeds, uint64_t rng_seed) {
    try {
        auto* b = new OrcBatch();
        b->scene.load(pack_path);
        for (int i = 0; i < num_envs; ++i) {
            b->envs.emplace_back(new Env());
            b->envs.back()->init(&b->scene, i, mode != 0, terrain_seeds ? terrain_seeds[i] : (uint64_t)(1 + i), rng_seed);
        }
        return b;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void orc_destroy(OrcBatch* b) { delete b; }

int orc_num_dof(OrcBatch* b) { return b->scene.ndof; }
int orc_num_joints(OrcBatch* b) { return b->scene.nj; }
int orc_state_size(OrcBatch* b) { return kNumGroundSamples + 4 * b->scene.nj - 1; }
int orc_action_size(OrcBatch* b) { return 1 + b->scene.n_opt; }
int orc_num_params(OrcBatch* b) { return b->scene.n_params; }

void orc_set_phys(OrcBatch* b, const double* p7) {
    PhysParams& pp = b->scene.phys;
    pp.kn = p7[0]; pp.dn = p7[1]; pp.mu = p7[2]; pp.v_eps = p7[3]; pp.contact_tol = p7[4]; pp.k_lim = p7[5]; pp.d_lim = p7[6];
}
void orc_set_explore(OrcBatch* b, int enable, double rate, double temp, double base_rate) {
    for (auto& e : b->envs) { e->enable_exp = enable != 0; e->exp_rate = rate; e->exp_temp = temp; e->exp_base_rate = base_rate; }
}

// outer update for all envs, optionally multi-threaded (thread-per-env-slice like cOptScenarioPoliEval::Run)
void orc_update(OrcBatch* b, double dt, int num_threads) {
    int n = (int)b->envs.size();
    if (num_threads <= 1) { for (auto& e : b->envs) e->update(dt); return; }
    // each worker is pinned to one of the CPUs this process may run on (the timed baseline should not migrate)
    std::vector<int> cpus;
    cpu_set_t allowed;
    CPU_ZERO(&allowed);
    if (sched_getaffinity(0, sizeof(allowed), &allowed) == 0)
        for (int c = 0; c < CPU_SETSIZE; ++c) if (CPU_ISSET(c, &allowed)) cpus.push_back(c);
    std::vector<std::thread> th;
    for (int t = 0; t < num_threads; ++t)
        th.emplace_back([=]() {
            if (!cpus.empty()) {
                cpu_set_t one;
                CPU_ZERO(&one);
                CPU_SET(cpus[t % cpus.size()], &one);
                pthread_setaffinity_np(pthread_self(), sizeof(one), &one);
            }
            for (int i = t; i < n; i += num_threads) b->envs[i]->update(dt);
        });
    for (auto& t : th) t.join();
}
void orc_env_step(OrcBatch* b, int env, double h) { b->envs[env]->env_step(h); }
void orc_reset(OrcBatch* b, int env) { b->envs[env]->reset(); }
// scenario-level pinning (tests/test_ref_pinning_cpu.py): the end-of-update tail on its own, and the action id CommandRandAction
// queued at the last reset (-1: none pending)
void orc_end_update(OrcBatch* b, int env, double dt) { b->envs[env]->time += dt; b->envs[env]->end_update(); }
// every exploration draw of every env from one restated cRand (the reference has one process-global engine), then a fresh
// terrain stream + reset for env `env` (what cOptScenarioPoliEval::BuildScenePool / the pinning harness do after Init)
void orc_use_ref_rand(OrcBatch* b, unsigned long seed) {
    b->ref_rand.seed(seed);
    for (auto& e : b->envs) e->rng.ref = &b->ref_rand;
}
int orc_rand_peek(OrcBatch* b) { Rand c = b->ref_rand; return c.rand_int(); }
void orc_reseed_reset(OrcBatch* b, int env, unsigned long terrain_seed) {
    b->envs[env]->ground.rand.seed(terrain_seed);
    b->envs[env]->reset();
}
int orc_pending_command(OrcBatch* b, int env) { auto& c = b->envs[env]->commands; return c.empty() ? -1 : c.back(); }

void orc_get_state(OrcBatch* b, int env, double* q, double* qd, double* tau_held, uint8_t* contact) {
    Env& e = *b->envs[env];
    int nd = b->scene.ndof, nj = b->scene.nj;
    if (q) std::memcpy(q, e.q, 8 * nd);
    if (qd) std::memcpy(qd, e.qd, 8 * nd);
    if (tau_held) std::memcpy(tau_held, e.tau_held, 8 * nd);
    if (contact) for (int j = 0; j < nj; ++j) contact[j] = e.contact[j] ? 1 : 0;
}
void orc_set_state(OrcBatch* b, int env, const double* q, const double* qd, const double* tau_held, const uint8_t* contact) {
    Env& e = *b->envs[env];
    int nd = b->scene.ndof, nj = b->scene.nj;
    if (q) std::memcpy(e.q, q, 8 * nd);
    if (qd) std::memcpy(e.qd, qd, 8 * nd);
    if (tau_held) std::memcpy(e.tau_held, tau_held, 8 * nd);
    if (contact) for (int j = 0; j < nj; ++j) e.contact[j] = contact[j] != 0;
    e.update_kin();
}
// controller block: [state, phase, first_cycle, cur_cycle_time, prev_cycle_time, cur_stumble, prev_stumble,
//  prev_com(2), prev_dist(2), action_id, params[30], pd_target[nj], fall(5: dist_counter, contact_counter, sum,
//  prev_check(2)), fail_fall_dist, exp_critic, exp_actor, cycle_count]
int orc_get_ctrl(OrcBatch* b, int env, double* out) {
    Env& e = *b->envs[env];
    int k = 0;
    out[k++] = e.state; out[k++] = e.phase; out[k++] = e.first_cycle; out[k++] = e.cur_cycle_time; out[k++] = e.prev_cycle_time;
    out[k++] = e.cur_stumble; out[k++] = e.prev_stumble; out[k++] = e.prev_com[0]; out[k++] = e.prev_com[1];
    out[k++] = e.prev_dist[0]; out[k++] = e.prev_dist[1]; out[k++] = e.cur.id;
    for (int i = 0; i < b->scene.n_params; ++i) out[k++] = e.cur.params[i];
    for (int j = 0; j < b->scene.nj; ++j) out[k++] = e.pd_target[j];
    out[k++] = e.fall_dist_counter; out[k++] = e.fall_contact_counter; out[k++] = e.sum_fall_contact;
    out[k++] = e.prev_check_pos[0]; out[k++] = e.prev_check_pos[1]; out[k++] = e.fail_fall_dist;
    out[k++] = e.exp_critic; out[k++] = e.exp_actor; out[k++] = e.cycle_count; out[k++] = e.stance;
    return k;
}
// cSimCharacter::HasFallen / HasStumbled as the controller sees them, and the number of policy decisions taken so far
void orc_flags(OrcBatch* b, int env, int* out3) {
    Env& e = *b->envs[env];
    out3[0] = e.has_fallen() ? 1 : 0; out3[1] = e.has_stumbled() ? 1 : 0; out3[2] = (int)e.cycle_count;
}
// the soft-fall bookkeeping alone (no physics, no controller): reset as at an episode start, then one check per call
void orc_fall_reset(OrcBatch* b, int env) {
    Env& e = *b->envs[env];
    e.fall_dist_counter = 5; e.prev_check_pos[0] = e.q[0]; e.prev_check_pos[1] = e.q[1]; e.fail_fall_dist = false;
    e.fall_contact_counter = 0.1; e.sum_fall_contact = 0;
}
void orc_fall_update(OrcBatch* b, int env, double h) { b->envs[env]->update_fall_checks(h); }
double orc_calc_reward(OrcBatch* b, int env) { return b->envs[env]->calc_reward(); }
void orc_get_last_tau(OrcBatch* b, int env, double* tau) { std::memcpy(tau, b->envs[env]->last_tau, 8 * b->scene.ndof); }
void orc_get_poli_state(OrcBatch* b, int env, double* s) {
    Env& e = *b->envs[env];
    std::memcpy(s, e.poli_state.data(), 8 * e.poli_state.size());
}
void orc_get_net_out(OrcBatch* b, int env, double* y) { std::memcpy(y, b->envs[env]->last_net_out, 8 * b->scene.net.n_out); }
int orc_get_terrain(OrcBatch* b, int env, int seg, float* data, int cap, double* min_x, int* flip) {
    Env& e = *b->envs[env];
    const auto& s = e.ground.seg[seg];
    int n = (int)s.data.size();
    for (int i = 0; i < n && i < cap; ++i) data[i] = s.data[i];
    *min_x = s.min_x;
    *flip = e.ground.flip ? 1 : 0;
    return n;
}
double orc_sample_height(OrcBatch* b, int env, double x) { return b->envs[env]->ground.sample(x); }

// component probes
void orc_rbd(OrcBatch* b, int env, double* M, double* C) {
    Env& e = *b->envs[env];
    int nd = b->scene.ndof;
    double pose[kMaxDof];
    e.build_pose(pose);
    e.ctrl_model.update(pose, e.qd);
    for (int a = 0; a < nd; ++a) { for (int c = 0; c < nd; ++c) M[a * nd + c] = e.ctrl_model.M[a][c]; C[a] = e.ctrl_model.C[a]; }
}
// what: 0 gravity force [ndof] (cRBDUtil::CalcGravityForce), 1 Jacobian [6][ndof] (cRBDUtil::BuildJacobian; rows omega xyz, v xyz),
//       2 world position of every joint origin [nj][3]
void orc_rbd_extra(OrcBatch* b, int env, int what, double* out) {
    Env& e = *b->envs[env];
    const int nd = b->scene.ndof;
    double pose[kMaxDof];
    e.build_pose(pose);
    e.ctrl_model.update(pose, e.qd);
    if (what == 0) e.ctrl_model.gravity_force(out);
    else if (what == 1) {
        for (int k = 0; k < nd; ++k) {
            const SV& c = e.ctrl_model.J[k];
            out[0 * nd + k] = c.o.x; out[1 * nd + k] = c.o.y; out[2 * nd + k] = c.o.z;
            out[3 * nd + k] = c.v.x; out[4 * nd + k] = c.v.y; out[5 * nd + k] = c.v.z;
        }
    } else {
        for (int j = 0; j < b->scene.nj; ++j) { V3 p = e.ctrl_model.joint_world_pos(j); out[3 * j] = p.x; out[3 * j + 1] = p.y; out[3 * j + 2] = p.z; }
    }
}
void orc_forward_dynamics(OrcBatch* b, int env, const double* tau, double dt, double* qdd) {
    b->envs[env]->forward_dynamics(tau, dt, qdd, false);
}
void orc_net_eval(OrcBatch* b, const double* x, double* y) { b->scene.net.eval(x, y); }
// cNeuralNet::GetLayerState for the blob `name` after a forward pass on x; returns its size (0: unknown name), copies up to cap values
int orc_net_layer(OrcBatch* b, const double* x, const char* name, double* out, int cap) {
    std::vector<double> v;
    if (!b->scene.net.layer_state(x, name, v)) return 0;
    for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}
void orc_com(OrcBatch* b, int env, double* com, double* com_vel) {
    Env& e = *b->envs[env];
    e.update_kin();
    e.calc_com(com, com_vel);
}

// tuples: rows of [reward | s(S) | a(A) | s'(S)] doubles, flags, env ids; drained by orc_reset_tuples
int orc_num_tuples(OrcBatch* b) { int n = 0; for (auto& e : b->envs) n += (int)e->tuples.size(); return n; }
int orc_get_tuples(OrcBatch* b, double* rows, uint32_t* flags, int32_t* env_id, int cap) {
    int S = orc_state_size(b), A = 1 + b->scene.n_opt, n = 0;
    for (auto& e : b->envs)
        for (auto& t : e->tuples) {
            if (n >= cap) return n;
            double* r = rows + (size_t)n * (1 + S + A + S);
            r[0] = t.reward;
            std::memcpy(r + 1, t.s_beg.data(), 8 * S);
            std::memcpy(r + 1 + S, t.action.data(), 8 * A);
            std::memcpy(r + 1 + S + A, t.s_end.data(), 8 * S);
            flags[n] = t.flags;
            env_id[n] = e->env_id;
            ++n;
        }
    return n;
}
void orc_reset_tuples(OrcBatch* b) { for (auto& e : b->envs) e->tuples.clear(); }

void orc_eval_stats(OrcBatch* b, int64_t* cycles, int64_t* episodes, double* avg_dist, int64_t* steps) {
    int64_t c = 0, ep = 0, st = 0;
    double sum = 0;
    for (auto& e : b->envs) { c += e->cycle_count; ep += e->episode_count; sum += e->avg_dist * e->episode_count; st += e->total_steps; }
    *cycles = c; *episodes = ep; *avg_dist = ep ? sum / ep : 0; *steps = st;
}
int orc_dist_log(OrcBatch* b, int env, double* out, int cap) {
    auto& l = b->envs[env]->dist_log;
    int n = (int)l.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = l[i];
    return n;
}


// ---------------------------------------------------------------------------------------- MACE trainer (oracle/trainer.h)
struct OrcTrainer {
    Scene scene;
    MaceTrainer tr;
};
// p[10]: replay_cap, num_init_samples, num_steps_per_iter, freeze_target_iters, init_input_offset_scale, discount, base_lr,
//        momentum, weight_decay, seed
OrcTrainer* orc_trainer_create(const char* pack_path, const double* p) {
    try {
        auto* t = new OrcTrainer();
        t->scene.load(pack_path);
        if (!t->scene.net.valid) throw std::runtime_error("scene has no policy net");
        TrainerParams P;
        P.replay_cap = (int)p[0]; P.num_init_samples = (int)p[1]; P.num_steps_per_iter = (int)p[2]; P.freeze_target_iters = (int)p[3];
        P.init_input_offset_scale = (int)p[4]; P.discount = p[5]; P.base_lr = p[6]; P.momentum = p[7]; P.weight_decay = p[8];
        P.seed = (uint64_t)p[9];
        t->tr.init(t->scene.net, P);
        return t;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void orc_trainer_destroy(OrcTrainer* t) { delete t; }
int orc_trainer_num_params(OrcTrainer* t) { return (int)t->tr.net.theta.size(); }
int orc_trainer_tuple_width(OrcTrainer* t) { return t->tr.Wd; }
void orc_trainer_add_tuples(OrcTrainer* t, const double* rows, const uint32_t* flags, int n) {
    for (int i = 0; i < n; ++i) t->tr.add_tuple(rows + (size_t)i * t->tr.Wd, flags[i]);
}
void orc_trainer_train(OrcTrainer* t) { t->tr.train(); }
void orc_trainer_get_rows(OrcTrainer* t, const int32_t* ids, int n, float* rows, int32_t* flags) {
    for (int i = 0; i < n; ++i) {
        std::memcpy(rows + (size_t)i * t->tr.Wd, t->tr.row(ids[i]), sizeof(float) * t->tr.Wd);
        flags[i] = t->tr.flags[ids[i]];
    }
}
// what: 0 theta, 1 target theta, 2 history, 3 in_off, 4 in_scale, 5 out_off, 6 out_scale
void orc_trainer_get(OrcTrainer* t, int what, double* out) {
    const std::vector<double>* v = nullptr;
    switch (what) {
        case 0: v = &t->tr.net.theta; break;
        case 1: v = &t->tr.target.theta; break;
        case 2: v = &t->tr.history; break;
        case 3: v = &t->tr.net.in_off; break;
        case 4: v = &t->tr.net.in_scale; break;
        case 5: v = &t->tr.net.out_off; break;
        default: v = &t->tr.net.out_scale; break;
    }
    std::memcpy(out, v->data(), v->size() * 8);
}
void orc_trainer_set_theta(OrcTrainer* t, const double* theta) { std::memcpy(t->tr.net.theta.data(), theta, t->tr.net.theta.size() * 8); }
// counters: iter, actor_iter, stage, num, head, total, critic_count, actor_count, actor_batch_count
void orc_trainer_counters(OrcTrainer* t, int64_t* c) {
    c[0] = t->tr.iter; c[1] = t->tr.actor_iter; c[2] = t->tr.stage; c[3] = t->tr.num; c[4] = t->tr.head; c[5] = t->tr.total;
    c[6] = (int64_t)t->tr.critic_buf.size(); c[7] = (int64_t)t->tr.actor_buf.size(); c[8] = (int64_t)t->tr.actor_batch.size();
}
void orc_trainer_losses(OrcTrainer* t, double* l) { l[0] = t->tr.last_critic_loss; l[1] = t->tr.last_actor_loss; }
int orc_trainer_lists(OrcTrainer* t, int which, int32_t* out, int cap) {
    const std::vector<int>& v = which == 0 ? t->tr.critic_buf : (which == 1 ? t->tr.actor_buf : (which == 2 ? t->tr.actor_batch :
                                (which == 3 ? t->tr.last_critic_ids : t->tr.last_actor_ids)));
    int n = std::min((int)v.size(), cap);
    for (int i = 0; i < n; ++i) out[i] = v[i];
    return (int)v.size();
}
// Euclidean loss and its gradient for an un-normalised problem (X [B][S], Y [B][n_out]) at the current weights; no update.
double orc_trainer_loss_grad(OrcTrainer* t, const double* X, const double* Y, int B, double* grad_out) {
    MaceNet& net = t->tr.net;
    const int S = net.tp.n_in, no = net.tp.n_out;
    std::vector<double> xn((size_t)B * S), dy((size_t)B * no), grad;
    for (int n = 0; n < B; ++n)
        for (int i = 0; i < S; ++i) xn[(size_t)n * S + i] = (X[(size_t)n * S + i] + net.in_off[i]) * net.in_scale[i];
    BatchActs acts;
    net.forward(B, xn.data(), acts);
    double loss = 0;
    for (int n = 0; n < B; ++n)
        for (int i = 0; i < no; ++i) {
            double lab = (Y[(size_t)n * no + i] + net.out_off[i]) * net.out_scale[i];
            double d = acts.y[(size_t)n * no + i] - lab;
            loss += d * d;
            dy[(size_t)n * no + i] = d / B;
        }
    loss /= 2.0 * B;
    if (grad_out) {
        net.backward(B, acts, dy.data(), grad);
        std::memcpy(grad_out, grad.data(), grad.size() * 8);
    }
    return loss;
}
void orc_trainer_eval_batch(OrcTrainer* t, int target, const double* X, int B, double* Y) {
    std::vector<double> y;
    (target ? t->tr.target : t->tr.net).eval_batch(B, X, y);
    std::memcpy(Y, y.data(), y.size() * 8);
}

// ---- network-level operations of the trainer, for the pinning test that runs the reference's compiled cMACETrainer on top of them
// (its cNeuralNet calls land here), and the switch that makes the restated trainer draw like the reference
void orc_trainer_use_ref_rand(OrcTrainer* t, unsigned long seed) { t->tr.use_ref_rand = true; t->tr.ref_rand.seed(seed); }
double orc_trainer_solver_step(OrcTrainer* t, const double* X, const double* Y) {
    const int B = t->tr.P.batch;
    std::vector<double> x(X, X + (size_t)B * t->tr.S), y(Y, Y + (size_t)B * t->tr.net.tp.n_out);
    return t->tr.solver_step(x, y);
}
void orc_trainer_copy_to_target(OrcTrainer* t) {
    auto& tr = t->tr;
    tr.target.theta = tr.net.theta;
    tr.target.in_off = tr.net.in_off; tr.target.in_scale = tr.net.in_scale;
    tr.target.out_off = tr.net.out_off; tr.target.out_scale = tr.net.out_scale;
}
void orc_trainer_set_input_offset_scale(OrcTrainer* t, int target, const double* off, const double* scale) {
    auto& n = target ? t->tr.target : t->tr.net;
    for (int i = 0; i < t->tr.S; ++i) { n.in_off[i] = off[i]; n.in_scale[i] = scale[i]; }
}
// training-loop pin: the trainer draws from the environments' restated cRand (orc_use_ref_rand), the environments evaluate the
// trainer's current weights, and the curriculum sets the terrain parameters (cScenarioSimChar::SetTerrainParamsLerp)
void orc_trainer_set_batch(OrcTrainer* t, int batch) { t->tr.P.batch = batch; }      // Caffe's batch size comes from the net file (32); smaller for quick tests
void orc_trainer_share_rand(OrcTrainer* t, OrcBatch* b) { t->tr.shared_rand = &b->ref_rand; }
void orc_set_net_from_trainer(OrcBatch* b, OrcTrainer* t) { t->tr.net.store_to(b->scene.net); }
void orc_set_terrain_lerp(OrcBatch* b, double lerp) { for (auto& e : b->envs) b->scene.terrain_params_lerp(lerp, e->ground.params); }
void orc_calc_offset_scale(const double* X, int n, int S, double* off, double* scale) { MaceTrainer::calc_offset_scale(X, n, S, off, scale); }

// ---------------------------------------------------------------------------------------- generator-level terrain / RNG probes
// (compared bit for bit with the reference's own cTerrainGen2D / cRand compiled into oracle/_ref, tests/test_ref_pinning_cpu.py)
int orc_terrain_build(int type, const double* params40, unsigned long seed, double width, float* out, int cap, double* total_w) {
    Rand r;
    r.seed(seed);
    std::vector<float> data;
    double w = TerrainGen::build(type, width, params40, r, data);
    if (total_w) *total_w = w;
    int n = (int)data.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = data[i];
    return n;
}
int orc_terrain_build_after_flat(int type, const double* params40, unsigned long seed, double flat_w, double width, float* out, int cap) {
    Rand r;
    r.seed(seed);
    std::vector<float> data;
    TerrainGen::add_flat(flat_w, data);
    TerrainGen::build(type, width, params40, r, data);
    int n = (int)data.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = data[i];
    return n;
}
void orc_rand_stream(unsigned long seed, int kind, double a, double b, int n, double* out) {
    Rand r;
    r.seed(seed);
    for (int i = 0; i < n; ++i) {
        switch (kind) {
            case 0: out[i] = r.rand_double(); break;
            case 1: out[i] = r.rand_double(a, b); break;
            case 2: out[i] = r.rand_int(); break;
            case 3: out[i] = r.rand_int((int)a, (int)b); break;
            case 4: out[i] = r.flip_coin() ? 1 : 0; break;
            default: out[i] = r.rand_sign(); break;
        }
    }
}

// ---------------------------------------------------------------------------------------- streaming ground alone (terrain.h: Ground)
struct OrcGround { Ground g; };
OrcGround* orc_ground_create(int type, const double* params40, unsigned long seed, double bmin, double bmax) {
    OrcGround* o = new OrcGround();
    o->g.type = type;
    for (int i = 0; i < pTerrainParamMax; ++i) o->g.params[i] = params40[i];
    o->g.rand.seed(seed);
    o->g.init_segments(bmin, bmax);
    return o;
}
void orc_ground_destroy(OrcGround* o) { delete o; }
void orc_ground_update(OrcGround* o, double bmin, double bmax) { o->g.update(bmin, bmax); }
int orc_ground_segment(OrcGround* o, int s, float* out, int cap, double* min_x) {
    const Ground::Seg& sg = o->g.seg[o->g.seg_id(s)];
    int n = (int)sg.data.size();
    for (int i = 0; i < n && i < cap; ++i) out[i] = sg.data[i];
    *min_x = sg.min_x;
    return n;
}
int orc_ground_flipped(OrcGround* o) { return o->g.flip ? 1 : 0; }
double orc_ground_sample(OrcGround* o, double x) { return o->g.sample(x); }
}  // extern "C"
