// deepterrainrl_b200 -- host runtime behind the C ABI (include/terrainrl_b200.h).
// Owns device memory (plain cudaMalloc; no torch in the boundary), the CUDA stream, the captured CUDA graph of one
// outer update (21 step launches + 20 decision launches) and the pinned staging buffers for tuple / statistics reads.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstring>
#include <map>
#include <string>
#include <algorithm>
#include <vector>

#include "../../include/terrainrl_b200.h"
#include "ref_loader.h"
#include "scene_pack.h"
#include "model_io.h"
#include "trl_types.h"
#include "trl_fcmaps.h"

namespace trl {
cudaError_t upload_model(const ModelConst& mc);
size_t step_smem_bytes();
cudaError_t configure_step_kernels();
void launch_step(const Buffers& B, double h, int flags, int lists, cudaStream_t st, int group = 0, int n_groups = 1);
void launch_reset(const Buffers& B, const uint64_t* seeds, const int* env_ids, int count, int reseed, cudaStream_t st);
size_t decide_smem_bytes();
cudaError_t configure_decide_kernel();
void launch_decide(const Buffers& B, const NetWeights& W, const ExpSettings* ex, int* done_count, int grid, int list, int rearm,
                   cudaStream_t st);
void launch_stats(const Buffers& B, double* out, cudaStream_t st);
size_t decide_fc_smem_bytes();
cudaError_t configure_decide2_kernels();
void launch_decide2(const Buffers& B, const NetWeights& W, const ExpSettings* ex, const FcMaps& maps, double* act2, int* done_count, int grid,
                    int fc_clusters, int list, int rearm, cudaStream_t st, int part);
void launch_terrain(const Buffers& B, double lookahead, cudaStream_t st);
}  // namespace trl
namespace trl_cg {
cudaError_t upload_model(const trl::ModelConst& mc);
void launch_step(const trl::Buffers& B, double h, int flags, int lists, cudaStream_t st, int group = 0, int n_groups = 1);
}  // namespace trl_cg


using namespace trl;

// both builds of the step kernels keep their own copy of the model constants
struct trl_handle;
static const trl_handle* g_model_owner = nullptr;    // whose ModelConst currently sits in the __constant__ copies
static int g_device = -1;                             // the one device this process drives (set by the first handle)
static cudaError_t upload_model_all(const ModelConst& mc) {
    cudaError_t e = trl::upload_model(mc);
    return e != cudaSuccess ? e : trl_cg::upload_model(mc);
}

static thread_local std::string g_err;
static int fail(const std::string& msg) { g_err = msg; return 1; }
int trl_fail(const std::string& msg) { return fail(msg); }
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) return fail(std::string(#call) + ": " + cudaGetErrorString(e__)); \
    } while (0)

static const char* kNetLayers[13] = {"terr_conv0", "terr_conv1", "terr_conv2", "terr_ip0", "ip0", "val_ip0", "val_ip1",
                                     "a0_ip0", "a0_ip1", "a1_ip0", "a1_ip1", "a2_ip0", "a2_ip1"};

#include "trl_handle.h"

// The kernels read the scene from __constant__ memory, of which there is one copy per process: a handle that is not the
// current owner re-uploads its model (after draining the owner's work) before it launches anything.
static int ensure_fc_maps(trl_handle* h);
static int ensure_model_only(trl_handle* h);
static int ensure_model(trl_handle* h) {
    if (ensure_model_only(h)) return 1;
    return ensure_fc_maps(h);
}
static int ensure_model_only(trl_handle* h) {
    // one process drives one GPU: __constant__ scene memory and the streams belong to the device of the handle
    // that was created first; a handle on another device is refused at creation (create_common)
    if (g_model_owner == h) return 0;
    if (cudaDeviceSynchronize() != cudaSuccess) return 1;
    if (upload_model_all(h->mc) != cudaSuccess) return 1;
    g_model_owner = h;
    return 0;
}

int trl_reupload_model(trl_handle* h) {
    if (!h) return fail("trl_reupload_model: null handle");
    g_model_owner = nullptr;
    return ensure_model(h) ? fail("model upload failed") : 0;
}

template <typename T>
static cudaError_t dalloc(trl_handle* h, T** p, size_t count) {
    cudaError_t e = cudaMalloc((void**)p, count * sizeof(T));
    if (e == cudaSuccess) { h->allocs.push_back(*p); e = cudaMemset(*p, 0, count * sizeof(T)); }
    return e;
}

static int fill_model(trl_handle* h, uint64_t rng_seed) {
    const ScenePack& s = h->scene;
    ModelConst& m = h->mc;
    std::memset(&m, 0, sizeof(m));
    const auto& mi = s.i32("meta_i32");
    const auto& mf = s.f64("meta_f64");
    if (mi.size() < 15 || mf.size() < 7) return fail("scene pack: meta records missing or too short");
    int char_type = mi[0], ctrl = mi[1];
    m.nj = mi[8]; m.ndof = mi[9];
    const bool raptor = char_type == 2;
    if (!((char_type == 1 && m.nj == 21 && m.ndof == 23) || (raptor && m.nj == 19 && m.ndof == 21)))
        return fail("unsupported character: expected the dog / goat (21 joints) or raptor (19 joints) skeleton");
    m.char_type = char_type;
    {
        // parameter layout and optimised-parameter mask (sim/DogController.cpp:81-121, sim/RaptorController.cpp:78-122)
        static const bool raptor_mask[37] = {false, true, true, false, false,
                                             true, false, true, true, true, true, true, true,
                                             true, false, true, true, true, true, true, true,
                                             false, false, true, true, true, true, true, true,
                                             false, false, true, true, true, true, true, true};
        m.n_params = raptor ? 37 : 30; m.misc_max = raptor ? 5 : 6; m.sp_max = raptor ? 8 : 6;
        m.n_opt = 0;
        for (int i = 0; i < m.n_params; ++i)
            if (raptor ? raptor_mask[i] : (i != 0)) m.opt_idx[m.n_opt++] = i;
        m.exp_noise = raptor ? 0.15 : 0.2;
        m.stumble_mask = 0; m.fall_mask = 0;
        if (raptor) {
            for (int j = 0; j < m.nj; ++j) if (j != 14 && j != 18 && j != 13 && j != 17) m.stumble_mask |= 1u << j;
            for (int j : {0, 1, 2, 3, 4, 5}) m.fall_mask |= 1u << j;
        } else {
            for (int j = 0; j < m.nj; ++j) if (j != 20 && j != 16 && j != 19 && j != 15) m.stumble_mask |= 1u << j;
            for (int j : {0, 1, 2, 3, 4, 5, 6, 7, 8}) m.fall_mask |= 1u << j;
        }
    }
    h->num_update_steps = mi[2];
    m.num_sim_substeps = mi[3];
    m.has_init_x = mi[4]; m.init_x = mf[2];
    m.terrain_type = mi[5];
    int n_sets = mi[6];
    m.has_net = mi[7];
    m.n_ctrl = mi[10]; m.n_actions = mi[11]; m.default_action = mi[12]; m.grav_comp = mi[13]; m.virt_forces = mi[14];
    if (m.n_ctrl > kMaxCtrlSets || m.n_actions > kMaxActions) return fail("too many controller sets / actions");
    m.is_mace = (ctrl == 3 || ctrl == 4 || ctrl == 7) ? 1 : 0;
    m.target_vel_x = (ctrl == 4) ? 2.0 : 4.0;   // sim/GoatControllerMACE.cpp:11-14, sim/DogController.cpp:625-628
    m.gx = mf[0]; m.gy = mf[1];
    m.exp_mode = h->mode == TRL_MODE_EXPLORE;
    m.rng_seed = rng_seed;
    const auto& J = s.f64("joints");
    const auto& Bd = s.f64("bodies");
    const auto& P = s.f64("pd");
    {
        // every table the loops below index, against the sizes they assume
        const size_t nj = (size_t)m.nj, np = (size_t)m.n_params;
        const int n_ctrl = mi[10], n_act = mi[11];
        if (n_ctrl < 1 || n_ctrl > kMaxCtrlSets || n_act < 0 || n_act > kMaxActions) return fail("too many controller sets / actions");
        if (J.size() < 7 * nj || Bd.size() < 9 * nj || P.size() < 6 * nj) return fail("scene pack: joints / bodies / pd tables too short");
        if (s.f64("ctrl_params").size() < (size_t)n_ctrl * np || s.f64("actions").size() < (size_t)4 * n_act)
            return fail("scene pack: controller parameter / action tables too short");
        if (s.f64("pose0").size() < (size_t)m.ndof || s.f64("vel0").size() < (size_t)m.ndof) return fail("scene pack: initial state too short");
        if (s.f64("terrain_default_params").size() < (size_t)kTerrainParams || mi[6] < 0 ||
            s.f64("terrain_params").size() < (size_t)mi[6] * kTerrainParams)
            return fail("scene pack: terrain parameter sets too short");
        for (size_t j = 0; j < nj; ++j) {
            const int par = (int)J[7 * j + 1];
            if (par >= (int)j || par < -1) return fail("scene pack: a joint's parent must precede it");
        }
    }
    int off = 0;
    m.total_mass = 0;
    for (int j = 0; j < m.nj; ++j) {
        const double* r = &J[7 * j];
        int type = (int)r[0];
        m.parent[j] = (int)r[1];
        if ((j == 0) != (m.parent[j] < 0) || (j == 0 && type != 1) || (j > 0 && type != 0)) return fail("unsupported joint layout");
        m.attach_x[j] = r[2]; m.attach_y[j] = r[3];
        m.lim_lo[j] = r[5]; m.lim_hi[j] = r[6];
        m.has_limit[j] = (j > 0 && r[5] <= r[6]) ? 1 : 0;
        m.dof[j] = off;
        off += (j == 0) ? 3 : 1;
        const double* b = &Bd[9 * j];
        if ((int)b[0] != 0) return fail("only box bodies are supported");
        m.mass[j] = b[1]; m.body_ax[j] = b[2]; m.body_ay[j] = b[3]; m.body_theta[j] = b[5];
        m.body_cos[j] = std::cos(b[5]); m.body_sin[j] = std::sin(b[5]);
        m.half_x[j] = 0.5 * b[6]; m.half_y[j] = 0.5 * b[7];
        // planar inertia about the joint origin: box about its COM + parallel-axis shift (sim/RBDUtil.cpp:562-583,614-623)
        m.izz_o[j] = b[1] / 12.0 * (b[6] * b[6] + b[7] * b[7]) + b[1] * (b[2] * b[2] + b[3] * b[3]);
        m.izz_c[j] = b[1] / 12.0 * (b[6] * b[6] + b[7] * b[7]);
        m.total_mass += b[1];
        // tail parts carry collision group "none" (sim/SimDog.cpp:7,21-24)
        m.collidable[j] = (!raptor && j >= 9 && j <= 12) ? 0 : 1;   // every raptor part collides (sim/SimRaptor.cpp:4-27)
        const double* p = &P[6 * j];
        m.kp[j] = (j == 0) ? 0.0 : p[0]; m.kd[j] = (j == 0) ? 0.0 : p[1];
        m.torque_lim[j] = p[2]; m.target_theta0[j] = p[3]; m.target_vel[j] = p[4]; m.world_pd[j] = p[5] != 0;
    }
    // topology helpers
    m.max_depth = 0;
    for (int j = 0; j < m.nj; ++j) {
        m.depth[j] = (j == 0) ? 0 : m.depth[m.parent[j]] + 1;
        m.max_depth = std::max(m.max_depth, m.depth[j]);
        for (int c = 0; c < 4; ++c) m.child[j][c] = -1;
    }
    if (m.max_depth >= 12) return fail("kinematic tree too deep");
    for (int j = 1; j < m.nj; ++j) {
        int p = m.parent[j], slot = 0;
        while (slot < 4 && m.child[p][slot] >= 0) ++slot;
        if (slot == 4) return fail("a link has more than 4 children");
        m.child[p][slot] = j;
    }
    if (m.max_depth >= 16) return fail("kinematic tree too deep for the 4-round ancestor jumps");
    if (m.nj > 31) return fail("lane 31 must stay idle (zero source of the warp passes)");
    {
        // inward-pass schedule (parents precede children in index order, so walk the links backwards)
        std::vector<int> earliest(m.nj, 0);
        for (int j = 0; j < m.nj; ++j) { m.acc_round[j] = -1; m.acc_src[j] = ~0ull; }
        for (int p = m.nj - 1; p >= 0; --p) {
            // children of p, by the earliest round they can be eliminated in
            std::vector<std::pair<int, int>> ch;
            for (int c = 0; c < 4; ++c) if (m.child[p][c] >= 0) ch.push_back({earliest[m.child[p][c]], m.child[p][c]});
            std::sort(ch.begin(), ch.end());
            int last = -1;
            for (auto& ec : ch) {
                int t = std::max(ec.first, last + 1);
                m.acc_round[ec.second] = t;
                last = t;
            }
            earliest[p] = last + 1;
        }
        m.acc_rounds = earliest[0];
        m.acc_round[0] = m.acc_rounds;     // the root is never eliminated
        if (m.acc_rounds > 12) return fail("inward-pass schedule longer than 12 rounds");
        for (int j = 1; j < m.nj; ++j) {
            const int p = m.parent[j], r = m.acc_round[j];
            m.acc_src[p] = (m.acc_src[p] & ~(31ull << (5 * r))) | ((unsigned long long)j << (5 * r));
        }
    }
    for (int j = 0; j < m.nj; ++j) {
        m.anc_pow[j][0] = m.parent[j];
        for (int k = 1; k < 4; ++k) m.anc_pow[j][k] = m.anc_pow[j][k - 1] >= 0 ? m.anc_pow[m.anc_pow[j][k - 1]][k - 1] : -1;
    }
    {
        // box corners, ordered so that the bodies most likely to touch the ground (lowest in the zero pose: feet, shanks)
        // share the first 32-corner round; rounds whose corners are all clear of the terrain are skipped by the kernel.
        // reach = upper bound of |corner - root joint| over all poses (bounds the terrain window the contacts can see)
        std::vector<double> jy(m.nj, 0.0), jr(m.nj, 0.0), low(m.nj, 0.0);
        std::vector<int> order;
        m.reach = 0.0;
        for (int j = 0; j < m.nj; ++j) {
            if (j > 0) {
                jy[j] = jy[m.parent[j]] + m.attach_y[j];
                jr[j] = jr[m.parent[j]] + std::hypot(m.attach_x[j], m.attach_y[j]);
            }
            m.corner_base[j] = -1;
            if (!m.collidable[j]) continue;
            order.push_back(j);
            low[j] = 1e30;
            for (int cn = 0; cn < 4; ++cn) {
                double bx = (cn & 1) ? m.half_x[j] : -m.half_x[j], by = (cn & 2) ? m.half_y[j] : -m.half_y[j];
                double lx = m.body_ax[j] + m.body_cos[j] * bx - m.body_sin[j] * by;
                double ly = m.body_ay[j] + m.body_sin[j] * bx + m.body_cos[j] * by;
                low[j] = std::min(low[j], jy[j] + ly);
                m.reach = std::max(m.reach, jr[j] + std::hypot(lx, ly));
            }
        }
        std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return low[a] < low[b]; });
        m.n_corners = 0;
        for (int j : order) {
            m.corner_base[j] = m.n_corners;
            for (int cn = 0; cn < 4; ++cn) {
                double bx = (cn & 1) ? m.half_x[j] : -m.half_x[j], by = (cn & 2) ? m.half_y[j] : -m.half_y[j];
                m.corner_body[m.n_corners] = j;
                m.corner_lx[m.n_corners] = m.body_ax[j] + m.body_cos[j] * bx - m.body_sin[j] * by;
                m.corner_ly[m.n_corners] = m.body_ay[j] + m.body_sin[j] * bx + m.body_cos[j] * by;
                ++m.n_corners;
            }
        }
    }
    {
        // effector 0 / 1: dog toe (back foot) / finger (front foot); raptor right toe / left toe.
        // dog virtual forces stop at the root / torso (sim/DogController.cpp:1014-1016), raptor's at the root (:1053)
        const int toe = raptor ? 14 : 20, finger = raptor ? 18 : 16, torso = raptor ? 0 : 5;
        m.anc_mask_toe = m.anc_mask_finger = m.vf_mask_toe = m.vf_mask_finger = 0;
        for (int c = toe; c >= 0; c = m.parent[c]) m.anc_mask_toe |= 1u << c;
        for (int c = finger; c >= 0; c = m.parent[c]) m.anc_mask_finger |= 1u << c;
        for (int c = toe; c != 0 && c != torso; c = m.parent[c]) m.vf_mask_toe |= 1u << c;
        for (int c = finger; c != 0 && c != torso; c = m.parent[c]) m.vf_mask_finger |= 1u << c;
    }
    const auto& C = s.f64("ctrl_params");
    for (int c = 0; c < m.n_ctrl; ++c)
        for (int k = 0; k < m.n_params; ++k) m.ctrl_params[c][k] = C[c * m.n_params + k];
    const auto& A = s.f64("actions");
    for (int a = 0; a < m.n_actions; ++a) {
        m.act_idx0[a] = (int)A[4 * a]; m.act_idx1[a] = (int)A[4 * a + 1]; m.act_blend[a] = A[4 * a + 2]; m.act_cyclic[a] = A[4 * a + 3] != 0;
    }
    const auto& p0 = s.f64("pose0");
    const auto& v0 = s.f64("vel0");
    for (int k = 0; k < m.ndof; ++k) { m.pose0[k] = p0[k]; m.vel0[k] = v0[k]; }
    // cScenarioSimChar::SetTerrainParamsLerp (scenarios/ScenarioSimChar.cpp:255-272)
    const auto& tp = s.f64("terrain_params");
    const auto& td = s.f64("terrain_default_params");
    if (n_sets == 0) for (int i = 0; i < kTerrainParams; ++i) m.terrain_params[i] = td[i];
    else {
        double lerp = std::min(std::max(mf[3], 0.0), n_sets - 1.0);
        int i0 = (int)lerp, i1 = std::min(i0 + 1, n_sets - 1);
        lerp -= i0;
        for (int i = 0; i < kTerrainParams; ++i) m.terrain_params[i] = (1 - lerp) * tp[i0 * kTerrainParams + i] + lerp * tp[i1 * kTerrainParams + i];
    }
    {
        const char* vc = std::getenv("TRL_VERTEX_CONTACTS");
        m.phys = PhysParams{2.0e5, 2.0e3, 0.81, 0.01, 0.00025, 2.0e4, 20.0, (vc && vc[0] == '1') ? 1 : 0};
    }
    h->ex = ExpSettings{h->mode == TRL_MODE_EXPLORE ? 1 : 0, mf[4], mf[5], mf[6], m.exp_noise};   // uploaded by create_common
    if (m.has_net) {
        const auto& nd = s.i32("net_dims");
        if (nd.size() < 5) return fail("scene pack: net_dims missing");
        m.n_in = nd[0]; m.n_char = nd[1]; m.n_out = nd[2]; m.n_frags = nd[3]; m.frag = nd[4];
        if (m.n_out > kMaxNetOut || m.frag > 32 || m.n_frags > 8 || m.n_char > 96) return fail("unsupported net dimensions");
        const auto& os = s.f64("net_out_scale");
        if (m.n_frags < 1 || m.frag < 1 || os.size() < (size_t)(m.n_frags + m.frag)) return fail("scene pack: net_out_scale too short");
        for (int k = 0; k < m.frag; ++k) m.out_scale_actor0[k] = os[m.n_frags + k];
    }
    return 0;
}

static int upload_net(trl_handle* h, const double* const* blobs, const int64_t* counts, const double* const* vecs4, const int64_t* vcounts) {
    // blobs: 26 (w, b per layer), vecs4: in_off, in_scale, out_off, out_scale
    if (h->net_blobs.empty()) {
        h->net_blobs.assign(30, nullptr);
        h->net_counts.assign(30, 0);
        for (int i = 0; i < 30; ++i) {
            int64_t c = i < 26 ? counts[i] : vcounts[i - 26];
            h->net_counts[i] = c;
            CK(dalloc(h, &h->net_blobs[i], (size_t)c));
        }
    }
    for (int i = 0; i < 30; ++i) {
        int64_t c = i < 26 ? counts[i] : vcounts[i - 26];
        if (c != h->net_counts[i]) return fail("trl_set_weights: blob size mismatch");
        const double* src = i < 26 ? blobs[i] : vecs4[i - 26];
        CK(cudaMemcpyAsync(h->net_blobs[i], src, (size_t)c * 8, cudaMemcpyHostToDevice, h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    NetWeights& W = h->W;
    double** b = h->net_blobs.data();
    W.conv0_w = b[0]; W.conv0_b = b[1]; W.conv1_w = b[2]; W.conv1_b = b[3]; W.conv2_w = b[4]; W.conv2_b = b[5];
    W.tip0_w = b[6]; W.tip0_b = b[7]; W.ip0_w = b[8]; W.ip0_b = b[9];
    for (int k = 0; k < 4; ++k) { W.h0_w[k] = b[10 + 4 * k]; W.h0_b[k] = b[11 + 4 * k]; W.h1_w[k] = b[12 + 4 * k]; W.h1_b[k] = b[13 + 4 * k]; }
    W.in_off = b[26]; W.in_scale = b[27]; W.out_off = b[28]; W.out_scale = b[29];
    return 0;
}

// the decision kernel indexes the 26 blobs with the MACE topology's strides (data/policies/dog/nets/dog_mace3_deploy.prototxt):
// anything else must be refused before it reaches the device
static int check_net_counts(const ModelConst& m, const int64_t* counts, const int64_t* vcounts) {
    const int64_t want[10] = {16 * 8, 16, 32 * 16 * 4, 32, 32 * 32 * 4, 32, 64 * 32 * 187, 64, 256 * (int64_t)(64 + m.n_char), 256};
    for (int b = 0; b < 26; ++b) {
        int64_t w;
        if (b < 10) w = want[b];
        else {
            const int hd = (b - 10) / 4, r = (b - 10) % 4, nout = hd == 0 ? m.n_frags : m.frag;
            w = r == 0 ? 128 * 256 : (r == 1 ? 128 : (r == 2 ? (int64_t)nout * 128 : nout));
        }
        if (counts[b] != w) return fail("policy net: blob " + std::to_string(b) + " has " + std::to_string(counts[b]) + " values, the MACE topology needs " + std::to_string(w));
    }
    if (vcounts) {
        const int64_t vw[4] = {m.n_in, m.n_in, m.n_out, m.n_out};
        for (int k = 0; k < 4; ++k)
            if (vcounts[k] != vw[k]) return fail("policy net: offset / scale vector " + std::to_string(k) + " has the wrong length");
    }
    if (m.n_in != 200 + m.n_char || m.n_out != m.n_frags * (1 + m.frag)) return fail("policy net: input / output sizes do not match the MACE topology");
    return 0;
}

// ---- batched decision path: TMA descriptors over the terr_ip0 weights and the conv2-output scratch (trl_decide2.cuh)
#ifndef TRL_SIMT_EMU
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn tensor_map_encoder() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}
// [rows][5984] f64 row-major, box [box_rows][16 doubles = 128 B], 128-byte swizzle, rows past the end read as zero
static int encode_rows_map(CUtensorMap* map, const double* base, uint64_t rows, uint32_t box_rows) {
    EncodeTiledFn enc = tensor_map_encoder();
    if (!enc) return fail("cuTensorMapEncodeTiled is not available from this driver");
    if (((uintptr_t)base & 15u) != 0) return fail("TMA needs 16-byte aligned weights");
    const cuuint64_t dims[2] = {5984, rows};
    const cuuint64_t strides[1] = {5984 * 8};
    const cuuint32_t box[2] = {16, box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, (void*)base, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                           CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")");
    return 0;
}
#endif
// TMA descriptors for a forward pass through the batched kernels outside the decision path (the trainer's minibatch)
int trl_make_fc_maps(trl::FcMaps* out, const double* tip0_w, const double* act2, int rows) {
    std::memset(out, 0, sizeof(*out));
    out->w_ptr = tip0_w; out->a_ptr = act2; out->a_rows = rows; out->prof = nullptr;
#ifndef TRL_SIMT_EMU
    if (encode_rows_map(&out->w, tip0_w, 64, 64)) return 1;
    if (encode_rows_map(&out->a, act2, (uint64_t)rows, 32)) return 1;
#endif
    return 0;
}
static int ensure_fc_maps(trl_handle* h) {
    if (!h->decide_v2 || !h->mc.has_net || !h->act2[0] || !h->W.tip0_w) return 0;
    if (h->fc_maps_w == h->W.tip0_w) return 0;
    for (int k = 0; k < h->lag; ++k) {
        h->fc_maps[k].w_ptr = h->W.tip0_w;
        h->fc_maps[k].a_ptr = h->act2[k];
        h->fc_maps[k].a_rows = h->n;
        h->fc_maps[k].prof = nullptr;
#ifndef TRL_SIMT_EMU
        if (encode_rows_map(&h->fc_maps[k].w, h->W.tip0_w, 64, 64)) return 1;
        if (encode_rows_map(&h->fc_maps[k].a, h->act2[k], (uint64_t)h->n, 32)) return 1;
#endif
    }
    h->fc_maps_w = h->W.tip0_w;
    return 0;
}
// the decision launch(es) of one env-step
static void enqueue_decide(trl_handle* h, int list, int rearm, cudaStream_t st, int part = 3, int slot = 0) {
    if (h->decide_v2)
        launch_decide2(h->B, h->W, h->d_ex, h->fc_maps[slot], h->act2[slot], h->done_count, h->decide_grid, h->fc_clusters, list, rearm, st, part);
    else if (part & 1)
        launch_decide(h->B, h->W, h->d_ex, h->done_count, h->decide_grid, list, rearm, st);
}
static int num_decide_launches(const trl_handle* h) { return h->decide_v2 ? 2 : 1; }

static void destroy_graphs(trl_handle* h);
void trl_drop_graphs(trl_handle* h) { destroy_graphs(h); }
static void destroy_graphs(trl_handle* h) {
    for (auto& kv : h->graphs) cudaGraphExecDestroy(kv.second);
    h->graphs.clear();
}

// One outer update = num_update_steps env-steps.  An env-step is split at the policy decision: launch S_i runs the
// controller half of env-step i-1 and the physics half of env-step i for every env, then the decision kernel D_i serves the
// (few) envs that reached a cycle boundary in S_i.
//
//   serial schedule      T  S_0  D_0  S_1  D_1 ... S_{ns-1}  D_{ns-1}  S_end
//   overlapped schedule  main stream    T  S_0  S_1  S_2  S_3  S_4 ...  S_{ns-1}  S_end          (back to back)
//   (lag = 2 shown)      side stream 0        D_0 -> C_0 [steps 1, 2]     D_2 -> C_2 [steps 3, 4] ...
//                        side stream 1             D_1 -> C_1 [steps 2, 3]     D_3 -> C_3 ...
//
// In the overlapped schedule the envs that reach a cycle boundary in S_l wait for D_l and are then advanced by the catch-up
// launch C_l (same kernel, one warp per list entry, registers live across its `lag` env-steps) while S_{l+1} .. S_{l+lag} skip
// them; S_{l+lag+1} needs C_l.  lag + 1 pending lists rotate (S_l appends to list l % (lag + 1), which C_{l-lag-1} re-armed); side
// stream l % lag carries D_l and C_l.  Every env still advances by exactly one env-step per launch equivalent, so results are
// the same as in the serial schedule up to the rounding of the catch-up build (a second instantiation of the kernel source with the
// env-step loop: ~1e-12 after 45 updates; tests/test_gpu_scenarios.py::test_overlap_matches_serial).  Why several steps of slack
// (profiles/timeline_r02_*.txt): a lone warp needs ~110 us per env-step whatever else runs, and the decision kernels wait for SM
// resources the step launch holds, so the chain D_l -> C_l is ~2.5 step launches long; with one step of slack it bounded the update.
// optional recorder of a timeline (trl_update_timeline): an event pair around every launch, on the stream it is launched on
struct Timeline {
    std::vector<cudaEvent_t> beg, end, fork;
    std::vector<int> kind, idx;      // kind: 0 terrain, 1 step S_i, 2 decision D_i (conv stage of the batched path), 3 catch-up C_i, 4 FC stage of D_i
    std::string err;                 // first CUDA error seen while enqueueing, with the launch it followed
    void check(const char* where) {
        const cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && err.empty())
            err = std::string(where) + " of launch kind " + std::to_string(kind.empty() ? -1 : kind.back()) + " #" + std::to_string(idx.empty() ? -1 : idx.back()) + ": " + cudaGetErrorString(e);
    }
    void open(cudaStream_t st, int k, int i) {
        check("before");
        cudaEvent_t b, e;
        cudaEventCreate(&b); cudaEventCreate(&e);
        beg.push_back(b); end.push_back(e); kind.push_back(k); idx.push_back(i);
        cudaEventRecord(b, st);
        check("open");
    }
    void close(cudaStream_t st) { check("launch"); cudaEventRecord(end.back(), st); check("close"); }
};
static void enqueue_update(trl_handle* h, double dt, bool overlap, Timeline* tl = nullptr) {
    const int ns = h->num_update_steps;
    const double step = dt / ns;
    cudaStream_t A = h->stream;
#define TL_OPEN(st, k, i) do { if (tl) tl->open(st, k, i); } while (0)
#define TL_CLOSE(st) do { if (tl) tl->close(st); } while (0)
    TL_OPEN(A, 0, 0); launch_terrain(h->B, 0.5, A); TL_CLOSE(A);
    if (!overlap) {
        for (int i = 0; i < ns; ++i) {
            TL_OPEN(A, 1, i); launch_step(h->B, step, i == 0 ? 2 : 3, 0, A); TL_CLOSE(A);
            TL_OPEN(A, 2, i); enqueue_decide(h, 0, 1, A, 1); TL_CLOSE(A); TL_OPEN(A, 4, i); enqueue_decide(h, 0, 1, A, 2); TL_CLOSE(A);
        }
        TL_OPEN(A, 1, ns); launch_step(h->B, step, 1 | 4, 0, A); TL_CLOSE(A);
        return;
    }
    // Overlapped schedule, `lag` env-steps deep.  Main stream: T S_0 S_1 S_2 ... S_{ns-1} S_end back to back.  The envs that reach a
    // cycle boundary in S_l (list l % (lag + 1)) are served on side stream l % lag: D_l decides them, the catch-up launch C_l then
    // advances exactly those envs by `lag` env-steps (fewer if the update ends first) while S_{l+1} .. S_{l+lag} skip them.
    // S_{l+lag+1} needs C_l.  The chain D_l -> C_l therefore has `lag` step launches of slack.
    const int lag = h->lag, nl = lag + 1;
    // Env groups: the main launch of an env-step is split into G launches over contiguous env ranges, group g on its own stream
    // (group 0 on the engine stream).  The groups only meet where the side work does (D_l needs S_l of every group, S_{l+lag+1} of every
    // group needs C_l), so group g's S_{i+1} starts as soon as ITS S_i has drained and its CTAs fill the SM slots the tail of the other
    // groups' launches would leave idle (4096 warps over 2368 resident-warp slots = 1.73 waves per monolithic launch).
    const int G = std::max(1, std::min(h->groups, kMaxGroups));
    std::vector<cudaEvent_t>& fe = tl ? tl->fork : h->fork_events;
    const size_t need = (size_t)(ns + 1) * G + ns + 2;
    if (fe.size() < need) {
        size_t old = fe.size();
        fe.resize(need);
        for (size_t i = old; i < fe.size(); ++i) cudaEventCreateWithFlags(&fe[i], cudaEventDisableTiming);
    }
    cudaEvent_t* ev_s = fe.data();                    // ev_s[i * G + g]: S_i of group g done
    cudaEvent_t* ev_c = fe.data() + (ns + 1) * G;     // ev_c[l]: C_l (l < ns - 1) / D_{ns-1} done
    cudaEvent_t ev_t = fe[need - 1];                  // terrain look-ahead done
    auto lists_of = [&](int app, int prev, int reps) { return (app % nl) | ((prev % nl) << 3) | (reps << 6); };
    auto stream_of = [&](int g) { return g == 0 ? A : h->group_stream[g - 1]; };
    if (G > 1) { cudaEventRecord(ev_t, A); for (int g = 1; g < G; ++g) cudaStreamWaitEvent(stream_of(g), ev_t, 0); }
    for (int g = 0; g < G; ++g) {
        cudaStream_t M = stream_of(g);
        TL_OPEN(M, 1, 0); launch_step(h->B, step, 2, lists_of(0, 0, 0), M, g, G); TL_CLOSE(M);
        cudaEventRecord(ev_s[g], M);
    }
    for (int i = 1; i < ns; ++i) {
        // side work for the envs that became due in S_{i-1}
        const int l = i - 1, slot = l % lag;
        cudaStream_t X = h->side[slot];
        for (int g = 0; g < G; ++g) cudaStreamWaitEvent(X, ev_s[l * G + g], 0);
        TL_OPEN(X, 2, l); enqueue_decide(h, l % nl, 0, X, 1, slot); TL_CLOSE(X); TL_OPEN(X, 4, l); enqueue_decide(h, l % nl, 0, X, 2, slot); TL_CLOSE(X);
        const int reps = std::min(lag, ns - 1 - l);     // equivalents l + 1 .. l + reps of the main launches; S_end is never caught up
        TL_OPEN(X, 3, l); trl_cg::launch_step(h->B, step, 1 | 2 | 16, lists_of(l + 1, l, reps), X); TL_CLOSE(X);
        cudaEventRecord(ev_c[l], X);
        for (int g = 0; g < G; ++g) {
            cudaStream_t M = stream_of(g);
            if (i >= nl) cudaStreamWaitEvent(M, ev_c[i - nl], 0);
            TL_OPEN(M, 1, i); launch_step(h->B, step, 1 | 2 | 8, lists_of(i, i, 0), M, g, G); TL_CLOSE(M);
            cudaEventRecord(ev_s[i * G + g], M);
        }
    }
    {
        const int l = ns - 1, slot = l % lag;
        cudaStream_t X = h->side[slot];
        for (int g = 0; g < G; ++g) cudaStreamWaitEvent(X, ev_s[l * G + g], 0);
        TL_OPEN(X, 2, l); enqueue_decide(h, l % nl, 1, X, 1, slot); TL_CLOSE(X); TL_OPEN(X, 4, l); enqueue_decide(h, l % nl, 1, X, 2, slot); TL_CLOSE(X);
        cudaEventRecord(ev_c[l], X);
    }
    for (int l = std::max(0, ns - nl); l < ns; ++l) cudaStreamWaitEvent(A, ev_c[l], 0);     // every side stream joins here
    for (int g = 1; g < G; ++g) cudaStreamWaitEvent(A, ev_s[(ns - 1) * G + g], 0);          // and every group stream
    TL_OPEN(A, 1, ns); launch_step(h->B, step, 1 | 4, 0, A); TL_CLOSE(A);
#undef TL_OPEN
#undef TL_CLOSE
}
static int update_launches(const trl_handle* h, bool overlap) {
    const int ns = h->num_update_steps, d = num_decide_launches(h);
    const int G = std::max(1, std::min(h->groups, kMaxGroups));
    return overlap ? (1 + G + d) * ns : (1 + d) * ns + 2;    // G x S_0..S_{ns-1} + D_0..D_{ns-1} + S_end + C_0..C_{ns-2} (the terrain look-ahead is not counted)
}

extern "C" {

const char* trl_last_error(void) { return g_err.c_str(); }

static trl_handle* create_common(trl_handle* h, int num_envs, int device, int mode, const uint64_t* terrain_seeds, uint64_t rng_seed);

trl_handle* trl_create_from_pack(const char* pack_path, int num_envs, int device, int mode, const uint64_t* terrain_seeds,
                                 uint64_t rng_seed) {
    auto* h = new trl_handle();
    try {
        std::string err;
        if (!h->scene.load(pack_path, &err)) { g_err = err; delete h; return nullptr; }
        return create_common(h, num_envs, device, mode, terrain_seeds, rng_seed);
    } catch (const std::exception& e) {       // bad_alloc / out_of_range from a damaged pack must not cross the C ABI
        g_err = std::string("trl_create_from_pack: ") + e.what();
        delete h;
        return nullptr;
    }
}

trl_handle* trl_create(int argc, const char* const* argv, const char* data_root, int num_envs, int device, int mode,
                       const uint64_t* terrain_seeds, uint64_t rng_seed) {
    auto* h = new trl_handle();
    try {
        build_scene_from_args(argc, argv, data_root ? data_root : "", &h->scene);
    } catch (const std::exception& e) {
        g_err = e.what();
        delete h;
        return nullptr;
    }
    return create_common(h, num_envs, device, mode, terrain_seeds, rng_seed);
}

int trl_pack_from_args(int argc, const char* const* argv, const char* data_root, const char* out_path) {
    ScenePack pack;
    try {
        build_scene_from_args(argc, argv, data_root ? data_root : "", &pack);
    } catch (const std::exception& e) {
        return fail(e.what());
    }
    std::string err;
    if (!pack.save(out_path, &err)) return fail(err);
    return 0;
}

static trl_handle* create_common(trl_handle* h, int num_envs, int device, int mode, const uint64_t* terrain_seeds, uint64_t rng_seed) {
    auto bail = [&](const std::string& why) -> trl_handle* {
        if (!why.empty()) g_err = why;
        trl_destroy(h);
        return nullptr;
    };
    h->device = device; h->n = num_envs; h->mode = mode;
    if (num_envs <= 0) return bail("num_envs must be positive");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return bail("terrainrl_b200 needs a CUDA device: no GPU visible (the product path has no CPU fallback)");
    if (g_device >= 0 && g_device != device)
        return bail("terrainrl_b200 drives one GPU per process (device " + std::to_string(g_device) + " is in use): start one process per GPU");
    if (cudaSetDevice(device) != cudaSuccess) return bail("cudaSetDevice failed");
    g_device = device;
    if (fill_model(h, rng_seed)) return bail("");
    if (cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking) != cudaSuccess) return bail("cudaStreamCreate failed");
    {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);   // hi = numerically lowest = greatest priority
        const char* lg = std::getenv("TRL_LAG");
        if (lg && lg[0]) h->lag = std::min(kMaxLists - 1, std::max(1, std::atoi(lg)));
        for (int k = 0; k < h->lag; ++k)
            if (cudaStreamCreateWithPriority(&h->side[k], cudaStreamNonBlocking, hi) != cudaSuccess) return bail("cudaStreamCreate (side) failed");
        h->aux_stream = h->side[0];
        const char* gr = std::getenv("TRL_GROUPS");
        if (gr && gr[0]) h->groups = std::min(kMaxGroups, std::max(1, std::atoi(gr)));     // default 2: profiles/step_kernel_r02_session2_ab.txt
        while (h->groups > 1 && group_chunk(num_envs, h->groups) * (h->groups - 1) >= num_envs) --h->groups;   // no empty group
        for (int g = 1; g < h->groups; ++g)
            if (cudaStreamCreateWithFlags(&h->group_stream[g - 1], cudaStreamNonBlocking) != cudaSuccess) return bail("cudaStreamCreate (group) failed");
        const char* serial = std::getenv("TRL_SERIAL_SCHEDULE");
        h->overlap = !(serial && serial[0] == '1');
    }
    auto ck = [&](cudaError_t e, const char* what) -> bool {
        if (e != cudaSuccess) { g_err = std::string(what) + ": " + cudaGetErrorString(e); return false; }
        return true;
    };
    g_model_owner = nullptr;
    if (ensure_model(h)) return bail("upload_model failed");
    if (!ck(configure_step_kernels(), "configure_step_kernels")) return bail("");
    if (!ck(configure_decide_kernel(), "configure_decide_kernel")) return bail("");
    if (!ck(configure_decide2_kernels(), "configure_decide2_kernels")) return bail("");
    {
        const char* v1 = std::getenv("TRL_DECIDE_V1");
        h->decide_v2 = !(v1 && v1[0] == '1');
    }
    Buffers& B = h->B;
    std::memset(&B, 0, sizeof(B));
    const size_t n = (size_t)num_envs;
    B.n = num_envs;
    B.S = kNumGroundSamples + 4 * h->mc.nj - 1;
    B.A = 1 + h->mc.n_opt;
    const int A = 1 + h->mc.n_opt;
    // an env finishes at most one cycle per outer update; twice that leaves room for the bursts of a synchronised start while a
    // fixed-size exchange block (trl_gather_tuples) drains the queue
    B.tuple_cap = std::max(4096, 2 * num_envs);
    B.dist_cap = std::max(65536, 16 * num_envs);
    bool ok = ck(dalloc(h, &B.d, (size_t)D_NUM_FIELDS * n), "alloc d") && ck(dalloc(h, &B.i, (size_t)I_NUM_FIELDS * n), "alloc i") &&
              ck(dalloc(h, &B.terrain, n * 2 * kTerrainCap), "alloc terrain") && ck(dalloc(h, &B.poli_state, n * B.S), "alloc poli") &&
              ck(dalloc(h, &B.net_out, n * kMaxNetOut), "alloc net_out") && ck(dalloc(h, &B.tuple_sbeg, n * B.S), "alloc sbeg") &&
              ck(dalloc(h, &B.tuple_action, n * kNumParams), "alloc action") && ck(dalloc(h, &B.com_stash, 2 * n), "alloc com") &&
              ck(dalloc(h, &B.pending_list, (size_t)(h->lag + 1) * n), "alloc pending") && ck(dalloc(h, &B.pending_count, kMaxLists), "alloc pc") &&
              ck(dalloc(h, &B.catchup_done, kMaxLists + 1), "alloc cd") &&
              ck(dalloc(h, &B.tuples, (size_t)B.tuple_cap * (1 + B.S + A + B.S)), "alloc tuples") &&
              ck(dalloc(h, &B.tuple_flags, (size_t)B.tuple_cap), "alloc tf") && ck(dalloc(h, &B.tuple_env, (size_t)B.tuple_cap), "alloc te") &&
              ck(dalloc(h, &B.tuple_count, 1), "alloc tc") && ck(dalloc(h, &B.dist_log, (size_t)B.dist_cap), "alloc dl") &&
              ck(dalloc(h, &B.dist_env, (size_t)B.dist_cap), "alloc de") && ck(dalloc(h, &B.dist_count, 1), "alloc dc") &&
              ck(dalloc(h, &h->done_count, 1), "alloc done") && ck(dalloc(h, &h->d_ex, 1), "alloc ex");
    for (int k = 0; ok && h->mc.has_net && k < h->lag; ++k) ok = ck(dalloc(h, &h->act2[k], n * (size_t)5984), "alloc act2");
    if (ok) ok = ck(cudaMemcpy(h->d_ex, &h->ex, sizeof(ExpSettings), cudaMemcpyHostToDevice), "upload ex");
    if (!ok) return bail("");
    if (h->mc.has_net) {
        const double* blobs[26];
        int64_t counts[26];
        for (int l = 0; l < 13; ++l) {
            const auto& w = h->scene.f64(std::string("net_") + kNetLayers[l] + "_w");
            const auto& b = h->scene.f64(std::string("net_") + kNetLayers[l] + "_b");
            blobs[2 * l] = w.data(); counts[2 * l] = (int64_t)w.size();
            blobs[2 * l + 1] = b.data(); counts[2 * l + 1] = (int64_t)b.size();
        }
        const char* vn[4] = {"net_in_offset", "net_in_scale", "net_out_offset", "net_out_scale"};
        const double* vecs[4];
        int64_t vcounts[4];
        for (int k = 0; k < 4; ++k) { const auto& v = h->scene.f64(vn[k]); vecs[k] = v.data(); vcounts[k] = (int64_t)v.size(); }
        if (check_net_counts(h->mc, counts, vcounts)) return bail("");
        if (upload_net(h, blobs, counts, vecs, vcounts)) return bail("");
    } else {
        std::memset(&h->W, 0, sizeof(h->W));
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) h->decide_grid = 8 * std::max(1, (2 * prop.multiProcessorCount) / 8);
    if (trl_seed_terrain(h, terrain_seeds, terrain_seeds ? num_envs : 0)) return bail("");
    return h;
}

int trl_destroy(trl_handle* h) {
    if (!h) return 0;
    if (h->comm) trl_comm_destroy(h);
    if (h->trainer) trl_trainer_orphan(h->trainer);   // the trainer object outlives its scenario as an inert shell
    if (g_model_owner == h) g_model_owner = nullptr;
    if (h->stream) cudaStreamSynchronize(h->stream);
    destroy_graphs(h);
    if (h->copy_stream) { cudaStreamSynchronize(h->copy_stream); cudaStreamDestroy(h->copy_stream); }
    if (h->snap_ready) cudaEventDestroy(h->snap_ready);
    if (h->snap_copied) cudaEventDestroy(h->snap_copied);
    if (h->snap_host) cudaFreeHost(h->snap_host);
    for (void* p : h->allocs) cudaFree(p);
    for (int k = 0; k < 8; ++k)
        if (h->side[k]) { cudaStreamSynchronize(h->side[k]); cudaStreamDestroy(h->side[k]); }
    for (int g = 0; g < kMaxGroups - 1; ++g)
        if (h->group_stream[g]) { cudaStreamSynchronize(h->group_stream[g]); cudaStreamDestroy(h->group_stream[g]); }
    for (auto e : h->fork_events) cudaEventDestroy(e);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
    return 0;
}

int trl_seed_terrain(trl_handle* h, const uint64_t* seeds, int n) {
    if (!h) return fail("trl_seed_terrain: null handle");
    if (ensure_model(h)) return fail("model upload failed");
    uint64_t* d_seeds = nullptr;
    if (seeds) {
        if (n != h->n) return fail("trl_seed_terrain: need one seed per env");
        CK(cudaMalloc((void**)&d_seeds, (size_t)n * 8));
        cudaError_t e = cudaMemcpyAsync(d_seeds, seeds, (size_t)n * 8, cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) { cudaFree(d_seeds); return fail(std::string("trl_seed_terrain: ") + cudaGetErrorString(e)); }
    }
    cudaError_t e = cudaMemsetAsync(h->B.pending_count, 0, kMaxLists * 4, h->stream);
    if (e == cudaSuccess) {
        launch_reset(h->B, d_seeds, nullptr, h->n, 1, h->stream);
        h->launches += 1;
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (d_seeds) cudaFree(d_seeds);
    if (e != cudaSuccess) return fail(std::string("trl_seed_terrain: ") + cudaGetErrorString(e));
    return 0;
}

int trl_reset(trl_handle* h, const int32_t* env_ids, int n) {
    if (!h) return fail("trl_reset: null handle");
    if (ensure_model(h)) return fail("model upload failed");
    int* d_ids = nullptr;
    int count = h->n;
    if (env_ids) {
        if (n < 0 || n > h->n) return fail("trl_reset: env count out of range");
        for (int k = 0; k < n; ++k)
            if (env_ids[k] < 0 || env_ids[k] >= h->n) return fail("trl_reset: env id out of range");
        if (n == 0) return 0;
        count = n;
        CK(cudaMalloc((void**)&d_ids, (size_t)n * 4));
        cudaError_t e = cudaMemcpyAsync(d_ids, env_ids, (size_t)n * 4, cudaMemcpyHostToDevice, h->stream);
        if (e != cudaSuccess) { cudaFree(d_ids); return fail(std::string("trl_reset: ") + cudaGetErrorString(e)); }
    }
    launch_reset(h->B, nullptr, d_ids, count, 0, h->stream);
    h->launches += 1;
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaStreamSynchronize(h->stream);
    if (d_ids) cudaFree(d_ids);
    if (e != cudaSuccess) return fail(std::string("trl_reset: ") + cudaGetErrorString(e));
    return 0;
}

int trl_update(trl_handle* h, double dt) {
    if (!h) return fail("trl_update: null handle");
    if (ensure_model(h)) return fail("model upload failed");
    if (!(dt > 0)) return 0;
    const int nlaunch = update_launches(h, h->overlap);
    if (h->use_graph) {
        long long key;
        std::memcpy(&key, &dt, 8);
        auto it = h->graphs.find(key);
        if (it == h->graphs.end()) {
            cudaGraph_t graph;
            CK(cudaStreamBeginCapture(h->stream, cudaStreamCaptureModeThreadLocal));
            enqueue_update(h, dt, h->overlap);
            CK(cudaStreamEndCapture(h->stream, &graph));
            cudaGraphExec_t exec;
            CK(cudaGraphInstantiate(&exec, graph, 0));
            cudaGraphDestroy(graph);
            it = h->graphs.emplace(key, exec).first;
        }
        CK(cudaGraphLaunch(it->second, h->stream));
    } else {
        enqueue_update(h, dt, h->overlap);
        CK(cudaGetLastError());
    }
    h->launches += nlaunch;
    return 0;
}

int trl_env_step(trl_handle* h, double step) {
    if (!h) return fail("trl_env_step: null handle");
    if (ensure_model(h)) return fail("model upload failed");
    launch_step(h->B, step, 2, 0, h->stream);
    enqueue_decide(h, 0, 1, h->stream);
    launch_step(h->B, step, 1, 0, h->stream);
    h->launches += 2 + num_decide_launches(h);
    CK(cudaGetLastError());
    return 0;
}

int trl_sync(trl_handle* h) {
    if (!h) return fail("trl_sync: null handle");
    CK(cudaStreamSynchronize(h->stream));
    int fault = 0;
    CK(cudaMemcpy(&fault, h->B.catchup_done + kMaxLists, 4, cudaMemcpyDeviceToHost));
    if (fault)
        return fail("an env finished a gait cycle inside a catch-up launch (cycle shorter than three env-steps): the overlapped schedule "
                    "cannot serve it; run with TRL_SERIAL_SCHEDULE=1");
    return 0;
}

int trl_set_explore(trl_handle* h, int enable, double rate, double temp, double base_rate) {
    if (!h) return fail("trl_set_explore: null handle");
    // the decision kernel reads the settings from device memory, so annealing them every update (cScenarioTrain::CalcExpRate
    // ...) neither re-captures the update graph nor drains the stream: the copy is ordered behind the work already queued
    h->ex.enable = enable; h->ex.rate = rate; h->ex.temp = temp; h->ex.base_rate = base_rate;
    CK(cudaMemcpyAsync(h->d_ex, &h->ex, sizeof(ExpSettings), cudaMemcpyHostToDevice, h->stream));
    return 0;
}

// ---------------------------------------------------------------------------------------------- model files
static const char* kDeployLayers[27] = {"slice0", "terr_conv0", "terr_relu0", "terr_conv1", "terr_relu1", "terr_conv2", "terr_relu2",
                                        "terr_ip0", "terr_relu3", "char_flatten0", "concat0", "ip0", "relu0", "relu0_relu0_0_split",
                                        "val_ip0", "val_relu0", "val_ip1", "a0_ip0", "a0_relu0", "a0_ip1", "a1_ip0", "a1_relu0", "a1_ip1",
                                        "a2_ip0", "a2_relu0", "a2_ip1", "output"};
// Caffe blob shapes of the 13 parameter layers (weights; the bias of each is [num_output])
static std::vector<std::vector<uint64_t>> mace_weight_dims(int n_char, int n_frags, int frag) {
    std::vector<std::vector<uint64_t>> d = {{16, 1, 1, 8}, {32, 16, 1, 4}, {32, 32, 1, 4}, {64, 32 * 187}, {256, (uint64_t)(64 + n_char)}};
    for (int hd = 0; hd < 4; ++hd) { d.push_back({128, 256}); d.push_back({(uint64_t)(hd == 0 ? n_frags : frag), 128}); }
    return d;
}
// cNeuralNet::OutputModel for host-side weights: blobs[26] in layer order (w, b per layer)
int trl_write_model(const char* path, const double* const* blobs, int n_char, int n_frags, int frag, const double* in_off,
                    const double* in_scale, const double* out_off, const double* out_scale, uint32_t mtime) {
    try {
        auto dims = mace_weight_dims(n_char, n_frags, frag);
        std::map<std::string, std::vector<H5Writer::Blob>> named;
        for (int l = 0; l < 13; ++l)
            named[kNetLayers[l]] = {H5Writer::Blob{dims[l], blobs[2 * l]}, H5Writer::Blob{{dims[l][0]}, blobs[2 * l + 1]}};
        std::vector<std::string> layers(kDeployLayers, kDeployLayers + 27);
        std::vector<uint8_t> img = H5Writer::build(layers, named, mtime);
        const std::string p = path;
        write_file(p, img.data(), img.size());
        const size_t dot = p.find_last_of('.');
        const int n_in = 200 + n_char, n_out = n_frags * (1 + frag);
        write_scale_file((dot == std::string::npos ? p : p.substr(0, dot)) + "_scale.txt", in_off, in_scale, n_in, out_off, out_scale, n_out);
    } catch (const std::exception& e) {
        return fail(e.what());
    }
    return 0;
}
// the policy the scenario currently evaluates (the attached trainer's net if there is one)
int trl_output_model(trl_handle* h, const char* path, uint32_t mtime) {
    if (!h) return fail("trl_output_model: null handle");
    if (!h->mc.has_net) return fail("trl_output_model: scene has no policy net");
    CK(cudaStreamSynchronize(h->stream));
    const NetWeights& W = h->W;
    const double* dev[30] = {W.conv0_w, W.conv0_b, W.conv1_w, W.conv1_b, W.conv2_w, W.conv2_b, W.tip0_w, W.tip0_b, W.ip0_w, W.ip0_b};
    for (int k = 0; k < 4; ++k) { dev[10 + 4 * k] = W.h0_w[k]; dev[11 + 4 * k] = W.h0_b[k]; dev[12 + 4 * k] = W.h1_w[k]; dev[13 + 4 * k] = W.h1_b[k]; }
    dev[26] = W.in_off; dev[27] = W.in_scale; dev[28] = W.out_off; dev[29] = W.out_scale;
    std::vector<std::vector<double>> host(30);
    const double* ptr[30];
    for (int i = 0; i < 30; ++i) {
        host[i].resize((size_t)h->net_counts[i]);
        CK(cudaMemcpy(host[i].data(), dev[i], host[i].size() * 8, cudaMemcpyDeviceToHost));
        ptr[i] = host[i].data();
    }
    return trl_write_model(path, ptr, h->mc.n_char, h->mc.n_frags, h->mc.frag, ptr[26], ptr[27], ptr[28], ptr[29], mtime);
}
int trl_trainer_set_theta(trl_trainer* t, const double* theta);
// cNeuralNet::LoadModel + LoadScale (learning/NeuralNet.cpp:157-186): Caffe HDF5 weights + `_scale.txt`
int trl_load_model(trl_handle* h, const char* h5_path, const char* scale_path) {
    if (!h) return fail("trl_load_model: null handle");
    if (!h->mc.has_net) return fail("trl_load_model: scene has no policy net");
    if (h->trainer) return fail("trl_load_model: a trainer owns the policy weights (load before trl_trainer_create)");
    try {
        H5Reader h5(h5_path);
        const auto& ds = h5.datasets();
        const double* blobs[26];
        int64_t counts[26];
        for (int l = 0; l < 13; ++l)
            for (int k = 0; k < 2; ++k) {
                auto it = ds.find(std::string("/data/") + kNetLayers[l] + "/" + std::to_string(k));
                if (it == ds.end()) return fail(std::string("trl_load_model: missing dataset for layer ") + kNetLayers[l]);
                blobs[2 * l + k] = it->second.data(); counts[2 * l + k] = (int64_t)it->second.size();
            }
        JValue sc = load_json(scale_path);
        const char* keys[4] = {"InputOffset", "InputScale", "OutputOffset", "OutputScale"};
        std::vector<double> vec[4];
        for (int k = 0; k < 4; ++k) {
            const JValue* v = sc.get(keys[k]);
            if (!v || v->type != JValue::Arr) return fail(std::string("trl_load_model: scale file lacks ") + keys[k]);
            for (auto& e : v->arr) vec[k].push_back(e.num);
            const size_t want = (size_t)(k < 2 ? h->mc.n_in : h->mc.n_out);
            if (vec[k].size() != want)
                return fail(std::string("trl_load_model: ") + keys[k] + " has " + std::to_string(vec[k].size()) + " entries, the net needs " + std::to_string(want));
        }
        return trl_set_weights(h, blobs, counts, 26, vec[0].data(), vec[1].data(), vec[2].data(), vec[3].data());
    } catch (const std::exception& e) {
        return fail(e.what());
    }
}
int trl_get_output_offset_scale(trl_handle* h, double* off, double* scale, int n);
// the same for a scene pack, without a device (host-only: parses the pack and the controller tables); used to check the
// restatement against the reference's own compiled BuildNNOutputOffsetScale (tests/test_ref_pinning_cpu.py)
int trl_pack_output_offset_scale(const char* pack_path, double* off, double* scale, int n) {
    trl_handle h;
    std::string err;
    if (!h.scene.load(pack_path, &err)) return fail(err);
    if (fill_model(&h, 0)) return 1;
    return trl_get_output_offset_scale(&h, off, scale, n);
}

// cBaseControllerMACE::BuildNNOutputOffsetScale (sim/BaseControllerMACE.cpp:75-113,131-167): critic outputs offset -0.5 scale 2;
// actor f centred on the optimised parameters of control set f % n_ctrl (cDogControllerMACE::BuildActorBias) and scaled by
// 1 / max_a |opt(a) - opt(default action)| over the action library
int trl_get_output_offset_scale(trl_handle* h, double* off, double* scale, int n) {
    if (!h) return fail("trl_get_output_offset_scale: null handle");
    const ModelConst& m = h->mc;
    const int fs = m.n_opt, nf = m.has_net ? m.n_frags : 0;
    if (n != nf * (1 + fs)) return fail("trl_get_output_offset_scale: size mismatch");
    auto ctrl_opt = [&](int set, int k) {
        const int idx = m.opt_idx[k];
        double v = m.ctrl_params[set][idx];
        if (idx == 0 || idx == 1 || (m.char_type == 2 && idx == 2)) v = std::fabs(v);     // PostProcessParams (TransTime, Cv, raptor Cd)
        return v;
    };
    auto action_opt = [&](int a, int k) {
        const double b = m.act_blend[a];
        return (1.0 - b) * ctrl_opt(m.act_idx0[a], k) + b * ctrl_opt(m.act_idx1[a], k);   // BlendCtrlParams + GetOptParams
    };
    std::vector<double> frag_scale(fs, 1.0);
    if (m.n_actions > 1) {
        const int d0 = m.default_action >= 0 ? m.default_action : 0;
        for (int k = 0; k < fs; ++k) {
            double mx = 0;
            for (int a = 0; a < m.n_actions; ++a)
                if (a != d0) mx = std::max(mx, std::fabs(action_opt(a, k) - action_opt(d0, k)));
            frag_scale[k] = mx > 0 ? 1.0 / mx : 1.0;
        }
    }
    for (int f = 0; f < nf; ++f) {
        off[f] = -0.5; scale[f] = 2.0;
        for (int k = 0; k < fs; ++k) { off[nf + f * fs + k] = -ctrl_opt(f % m.n_ctrl, k); scale[nf + f * fs + k] = frag_scale[k]; }
    }
    return 0;
}

// cScenarioSimChar::SetTerrainParamsLerp (scenarios/ScenarioSimChar.cpp:255-272), as cScenarioTrain::UpdateSceneCurriculum calls it
// (scenarios/ScenarioTrain.cpp:412-416): blends the scene's terrain parameter sets; segments generated from now on use the
// blend (cGroundVar2D::SetTerrainParams), existing terrain stays.
int trl_set_terrain_lerp(trl_handle* h, double lerp) {
    if (!h) return fail("trl_set_terrain_lerp: null handle");
    const auto& tp = h->scene.f64("terrain_params");
    const int n_sets = (int)(tp.size() / kTerrainParams);
    if (n_sets <= 0) return 0;
    lerp = std::min(std::max(lerp, 0.0), n_sets - 1.0);
    const int i0 = (int)lerp, i1 = std::min(i0 + 1, n_sets - 1);
    lerp -= i0;
    double blended[kTerrainParams];
    bool same = true;
    for (int i = 0; i < kTerrainParams; ++i) {
        blended[i] = (1 - lerp) * tp[i0 * kTerrainParams + i] + lerp * tp[i1 * kTerrainParams + i];
        same = same && blended[i] == h->mc.terrain_params[i];
    }
    if (same) return 0;      // cScenarioTrain calls this after every trainer step; an unchanged blend must not drain the device
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < kTerrainParams; ++i) h->mc.terrain_params[i] = blended[i];
    g_model_owner = nullptr;
    if (ensure_model(h)) return fail("model upload failed");
    return 0;
}

// cScenarioTrain::CalcExpRate / CalcExpTemp / CalcExpBaseRate / CalcCurriculumPhase (scenarios/ScenarioTrain.cpp:418-460).
// sp[9] = {init_exp_rate, exp_rate, init_exp_temp, exp_temp, init_exp_base_rate, exp_base_rate, trainer_num_anneal_iters,
// exp_base_anneal_iters, trainer_curriculum_iters}; out[4] = {exp_rate, exp_temp, exp_base_rate, curriculum_phase}
int trl_train_schedule(const double* sp, int iters, double* out) {
    auto clamp01 = [](double v) { return std::min(std::max(v, 0.0), 1.0); };
    double lerp = clamp01((double)iters / sp[6]);
    out[0] = (1 - lerp) * sp[0] + lerp * sp[1];
    out[1] = (1 - lerp) * sp[2] + lerp * sp[3];
    lerp = clamp01((double)iters / sp[7]);
    out[2] = (1 - lerp) * sp[4] + lerp * sp[5];
    const bool enable = sp[8] >= 1;
    out[3] = clamp01(enable ? (double)iters / sp[8] : 1.0);
    if (iters == 0) out[3] = 1.0;      // gInitCurriculumPhase
    return 0;
}

int trl_set_phys_params(trl_handle* h, const double* p) {
    if (!h) return fail("trl_set_phys_params: null handle");
    CK(cudaStreamSynchronize(h->stream));
    h->mc.phys = PhysParams{p[0], p[1], p[2], p[3], p[4], p[5], p[6], h->mc.phys.vertex_contacts};
    g_model_owner = nullptr;
    if (ensure_model(h)) return fail("model upload failed");
    return 0;
}

int trl_set_weights(trl_handle* h, const double* const* blobs, const int64_t* counts, int nblobs, const double* in_off,
                    const double* in_scale, const double* out_off, const double* out_scale) {
    if (!h) return fail("trl_set_weights: null handle");
    if (nblobs != 26) return fail("trl_set_weights: expected 26 blobs");
    if (!h->mc.has_net) return fail("trl_set_weights: scene has no policy net");
    if (h->trainer) return fail("trl_set_weights: a trainer owns the policy weights (use trl_trainer_set_theta)");
    if (!blobs || !counts || !in_off || !in_scale || !out_off || !out_scale) return fail("trl_set_weights: null argument");
    if (check_net_counts(h->mc, counts, nullptr)) return 1;
    CK(cudaStreamSynchronize(h->stream));
    const double* vecs[4] = {in_off, in_scale, out_off, out_scale};
    int64_t vcounts[4] = {h->mc.n_in, h->mc.n_in, h->mc.n_out, h->mc.n_out};
    if (upload_net(h, blobs, counts, vecs, vcounts)) return 1;
    for (int k = 0; k < h->mc.frag; ++k) h->mc.out_scale_actor0[k] = out_scale[h->mc.n_frags + k];
    g_model_owner = nullptr;
    if (ensure_model(h)) return fail("model upload failed");
    return 0;
}

int trl_sizes(trl_handle* h, int* num_envs, int* state, int* action, int* num_frags, int* frag_size, int* num_dof, int* num_joints) {
    if (!h) return fail("trl_sizes: null handle");
    if (num_envs) *num_envs = h->n;
    if (state) *state = h->B.S;
    if (action) *action = h->B.A;
    if (num_frags) *num_frags = h->mc.n_frags;
    if (frag_size) *frag_size = h->mc.n_opt;
    if (num_dof) *num_dof = h->mc.ndof;
    if (num_joints) *num_joints = h->mc.nj;
    return 0;
}

// The step kernel takes tuple slots from an atomic cursor and refuses rows past tuple_cap (trl_step.cu: exp_new_cycle_update), so the
// cursor itself says how many were refused.  The readers deliver what fits, count the rest in h->tuples_dropped when the block is
// reset, and answer TRL_E_TUPLE_OVERFLOW so that the loss cannot go unnoticed.
static int overflow_error(trl_handle* h, int overflow) {
    g_err = "tuple block overflow: " + std::to_string(overflow) + " tuples were refused (capacity " + std::to_string(h->B.tuple_cap) +
            "); hand the tuples over more often";
    return TRL_E_TUPLE_OVERFLOW;
}
static int fetch_tuples(trl_handle* h, int* n_out, int* overflow) {
    int n = 0;
    CK(cudaMemcpyAsync(&n, h->B.tuple_count, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *overflow = n > h->B.tuple_cap ? n - h->B.tuple_cap : 0;
    if (n > h->B.tuple_cap) n = h->B.tuple_cap;
    const size_t W = 1 + h->B.S + h->B.A + h->B.S;
    h->h_tuples.resize((size_t)std::max(n, 1) * W);
    h->h_tuple_flags.resize(std::max(n, 1));
    h->h_tuple_env.resize(std::max(n, 1));
    if (n > 0) {
        CK(cudaMemcpyAsync(h->h_tuples.data(), h->B.tuples, (size_t)n * W * 8, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(h->h_tuple_flags.data(), h->B.tuple_flags, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaMemcpyAsync(h->h_tuple_env.data(), h->B.tuple_env, (size_t)n * 4, cudaMemcpyDeviceToHost, h->stream));
        CK(cudaStreamSynchronize(h->stream));
    }
    *n_out = n;
    return 0;
}

int trl_num_tuples(trl_handle* h, int* out) {
    if (!h) return fail("trl_num_tuples: null handle");
    int n = 0;
    CK(cudaMemcpyAsync(&n, h->B.tuple_count, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    *out = std::min(n, h->B.tuple_cap);
    return n > h->B.tuple_cap ? overflow_error(h, n - h->B.tuple_cap) : 0;
}

int trl_get_tuples_f64(trl_handle* h, const double** rows, const uint32_t** flags, const int32_t** env_id, int* n) {
    if (!h) return fail("trl_get_tuples_f64: null handle");
    int overflow = 0;
    if (fetch_tuples(h, n, &overflow)) return 1;
    *rows = h->h_tuples.data(); *flags = h->h_tuple_flags.data(); *env_id = h->h_tuple_env.data();
    return overflow ? overflow_error(h, overflow) : 0;
}

int trl_get_tuples(trl_handle* h, const float** rows, const uint32_t** flags, const int32_t** env_id, int* n) {
    if (!h) return fail("trl_get_tuples: null handle");
    int overflow = 0;
    if (fetch_tuples(h, n, &overflow)) return 1;
    const size_t W = 1 + h->B.S + h->B.A + h->B.S;
    h->h_tuples_f32.resize((size_t)std::max(*n, 1) * W);
    for (size_t k = 0; k < (size_t)*n * W; ++k) h->h_tuples_f32[k] = (float)h->h_tuples[k];
    *rows = h->h_tuples_f32.data(); *flags = h->h_tuple_flags.data(); *env_id = h->h_tuple_env.data();
    return overflow ? overflow_error(h, overflow) : 0;
}

int trl_reset_tuples(trl_handle* h) {
    if (!h) return fail("trl_reset_tuples: null handle");
    int n = 0;
    CK(cudaMemcpyAsync(&n, h->B.tuple_count, 4, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    if (n > h->B.tuple_cap) h->tuples_dropped += n - h->B.tuple_cap;
    CK(cudaMemsetAsync(h->B.tuple_count, 0, 4, h->stream));
    return 0;
}

// cScenarioPoliEval::ResetAvgDist (scenarios/ScenarioPoliEval.cpp:132-136) for every env: mean distance and episode count back to
// zero, cycle count and distance log kept -- cOptScenarioPoliEval::EvalHelper calls it after each UpdateRecord
// (optimizer/scenarios/OptScenarioPoliEval.cpp:189-195)
int trl_reset_avg_dist(trl_handle* h) {
    if (!h) return fail("trl_reset_avg_dist: null handle");
    const size_t n = (size_t)h->n;
    CK(cudaMemsetAsync(h->B.d + (size_t)D_AVG_DIST * n, 0, n * 8, h->stream));
    CK(cudaMemsetAsync(h->B.i + (size_t)I_EPISODE_COUNT * n, 0, n * 4, h->stream));
    return 0;
}

int trl_eval_stats(trl_handle* h, int64_t* cycles, int64_t* episodes, double* avg_dist, int64_t* env_steps) {
    if (!h) return fail("trl_eval_stats: null handle");
    const size_t n = (size_t)h->n;
    std::vector<int> cyc(n), eps(n), lo(n), hi(n);
    std::vector<double> avg(n);
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(cyc.data(), h->B.i + (size_t)I_CYCLE_COUNT * n, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(eps.data(), h->B.i + (size_t)I_EPISODE_COUNT * n, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(lo.data(), h->B.i + (size_t)I_STEPS_LO * n, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hi.data(), h->B.i + (size_t)I_STEPS_HI * n, n * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(avg.data(), h->B.d + (size_t)D_AVG_DIST * n, n * 8, cudaMemcpyDeviceToHost));
    int64_t c = 0, e = 0, st = 0;
    double sum = 0;
    for (size_t k = 0; k < n; ++k) {
        c += cyc[k]; e += eps[k]; sum += avg[k] * eps[k];
        st += ((int64_t)hi[k] << 32) | (uint32_t)lo[k];
    }
    if (cycles) *cycles = c;
    if (episodes) *episodes = e;
    if (avg_dist) *avg_dist = e ? sum / e : 0.0;
    if (env_steps) *env_steps = st;
    return 0;
}

int trl_dist_log(trl_handle* h, const double** dist, const int32_t** env_id, int* n) {
    if (!h) return fail("trl_dist_log: null handle");
    int cnt = 0;
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(&cnt, h->B.dist_count, 4, cudaMemcpyDeviceToHost));
    cnt = std::min(cnt, h->B.dist_cap);
    h->h_dist.resize(std::max(cnt, 1));
    h->h_dist_env.resize(std::max(cnt, 1));
    if (cnt > 0) {
        CK(cudaMemcpy(h->h_dist.data(), h->B.dist_log, (size_t)cnt * 8, cudaMemcpyDeviceToHost));
        CK(cudaMemcpy(h->h_dist_env.data(), h->B.dist_env, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
    }
    *dist = h->h_dist.data(); *env_id = h->h_dist_env.data(); *n = cnt;
    return 0;
}

static int copy_plane_d(trl_handle* h, int field, int count, int env, double* out) {
    // strided gather of `count` consecutive planes for one env
    return cudaMemcpy2D(out, 8, h->B.d + (size_t)field * h->n + env, (size_t)h->n * 8, 8, count, cudaMemcpyDeviceToHost) != cudaSuccess;
}
static int copy_plane_i(trl_handle* h, int field, int count, int env, int* out) {
    return cudaMemcpy2D(out, 4, h->B.i + (size_t)field * h->n + env, (size_t)h->n * 4, 4, count, cudaMemcpyDeviceToHost) != cudaSuccess;
}

int trl_get_state(trl_handle* h, int env, double* pose, double* vel, double* held_torque, uint8_t* contact) {
    if (!h) return fail("trl_get_state: null handle");
    if (env < 0 || env >= h->n) return fail("env out of range");
    CK(cudaStreamSynchronize(h->stream));
    int nd = h->mc.ndof;
    if (pose && copy_plane_d(h, D_Q, nd, env, pose)) return fail("copy failed");
    if (vel && copy_plane_d(h, D_QD, nd, env, vel)) return fail("copy failed");
    if (held_torque && copy_plane_d(h, D_TAU, nd, env, held_torque)) return fail("copy failed");
    if (contact) {
        int mask = 0;
        if (copy_plane_i(h, I_CONTACT, 1, env, &mask)) return fail("copy failed");
        for (int j = 0; j < h->mc.nj; ++j) contact[j] = (mask >> j) & 1;
    }
    return 0;
}

int trl_set_state(trl_handle* h, int env, const double* pose, const double* vel, const double* held_torque, const uint8_t* contact) {
    if (!h) return fail("trl_set_state: null handle");
    if (env < 0 || env >= h->n) return fail("env out of range");
    CK(cudaStreamSynchronize(h->stream));
    int nd = h->mc.ndof;
    auto put = [&](int field, const double* src) {
        return cudaMemcpy2D(h->B.d + (size_t)field * h->n + env, (size_t)h->n * 8, src, 8, 8, nd, cudaMemcpyHostToDevice);
    };
    if (pose) CK(put(D_Q, pose));
    if (vel) CK(put(D_QD, vel));
    if (held_torque) CK(put(D_TAU, held_torque));
    if (contact) {
        int mask = 0;
        for (int j = 0; j < h->mc.nj; ++j) if (contact[j]) mask |= 1 << j;
        CK(cudaMemcpy(h->B.i + (size_t)I_CONTACT * h->n + env, &mask, 4, cudaMemcpyHostToDevice));
    }
    return 0;
}

int trl_get_state_all(trl_handle* h, double* pose, double* vel) {
    if (!h) return fail("trl_get_state_all: null handle");
    CK(cudaStreamSynchronize(h->stream));
    size_t bytes = (size_t)h->mc.ndof * h->n * 8;
    if (pose) CK(cudaMemcpy(pose, h->B.d + (size_t)D_Q * h->n, bytes, cudaMemcpyDeviceToHost));
    if (vel) CK(cudaMemcpy(vel, h->B.d + (size_t)D_QD * h->n, bytes, cudaMemcpyDeviceToHost));
    return 0;
}

int trl_get_ctrl(trl_handle* h, int env, double* out, int cap, int* n_out) {
    if (!h) return fail("trl_get_ctrl: null handle");
    if (env < 0 || env >= h->n) return fail("env out of range");
    CK(cudaStreamSynchronize(h->stream));
    std::vector<double> d(D_NUM_FIELDS);
    std::vector<int> iv(I_NUM_FIELDS);
    if (copy_plane_d(h, 0, D_NUM_FIELDS, env, d.data()) || copy_plane_i(h, 0, I_NUM_FIELDS, env, iv.data())) return fail("copy failed");
    std::vector<double> o;
    o.push_back(iv[I_STATE]); o.push_back(d[D_PHASE]); o.push_back(iv[I_FIRST_CYCLE]); o.push_back(d[D_CUR_CYCLE_T]);
    o.push_back(d[D_PREV_CYCLE_T]); o.push_back(d[D_CUR_STUMBLE]); o.push_back(d[D_PREV_STUMBLE]); o.push_back(d[D_PREV_COM_X]);
    o.push_back(d[D_PREV_COM_Y]); o.push_back(d[D_PREV_DIST_X]); o.push_back(d[D_PREV_DIST_Y]); o.push_back(iv[I_ACTION_ID]);
    for (int k = 0; k < h->mc.n_params; ++k) o.push_back(d[D_PARAMS + k]);
    for (int j = 0; j < h->mc.nj; ++j) o.push_back(d[D_PD_TARGET + j]);
    o.push_back(d[D_FALL_DIST_CNT]); o.push_back(d[D_FALL_CONTACT_CNT]); o.push_back(d[D_SUM_FALL]); o.push_back(d[D_PREV_CHECK_X]);
    o.push_back(d[D_PREV_CHECK_Y]); o.push_back(iv[I_FAIL_FALL_DIST]); o.push_back(iv[I_EXP_FLAGS] & 1); o.push_back((iv[I_EXP_FLAGS] >> 1) & 1);
    o.push_back(iv[I_CYCLE_COUNT]); o.push_back(iv[I_STANCE]);
    int n = (int)std::min<size_t>(o.size(), (size_t)cap);
    std::memcpy(out, o.data(), (size_t)n * 8);
    if (n_out) *n_out = n;
    return 0;
}

int trl_get_poli_state(trl_handle* h, int env, double* out) {
    if (!h) return fail("trl_get_poli_state: null handle");
    if (env < 0 || env >= h->n) return fail("env out of range");
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(out, h->B.poli_state + (size_t)env * h->B.S, (size_t)h->B.S * 8, cudaMemcpyDeviceToHost));
    return 0;
}
int trl_get_net_out(trl_handle* h, int env, double* out) {
    if (!h) return fail("trl_get_net_out: null handle");
    if (env < 0 || env >= h->n) return fail("env out of range");
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaMemcpy(out, h->B.net_out + (size_t)env * kMaxNetOut, (size_t)h->mc.n_out * 8, cudaMemcpyDeviceToHost));
    return 0;
}
int trl_get_terrain(trl_handle* h, int env, int seg, float* data, int cap, int* n, double* min_x, int* flip) {
    if (!h) return fail("trl_get_terrain: null handle");
    if (env < 0 || env >= h->n) return fail("env out of range");
    if (seg < 0 || seg > 1) return fail("terrain segment out of range (0 or 1)");
    CK(cudaStreamSynchronize(h->stream));
    int sn = 0, fl = 0;
    double mx = 0;
    if (copy_plane_i(h, seg == 0 ? I_SEG_N0 : I_SEG_N1, 1, env, &sn) || copy_plane_i(h, I_SEG_FLIP, 1, env, &fl) ||
        copy_plane_d(h, seg == 0 ? D_SEG_MINX0 : D_SEG_MINX1, 1, env, &mx))
        return fail("copy failed");
    int cnt = std::min(sn, cap);
    if (cnt > 0) CK(cudaMemcpy(data, h->B.terrain + ((size_t)env * 2 + seg) * kTerrainCap, (size_t)cnt * 4, cudaMemcpyDeviceToHost));
    *n = sn; *min_x = mx; *flip = fl;
    return 0;
}

int64_t trl_kernel_launches(trl_handle* h) { return h ? h->launches : -1; }

// Micro-benchmark of the decision kernel: marks the first `n_pending` envs as pending (their policy state is whatever the
// last real decision left) and times `iters` launches.  Perturbs those envs' actions: measurement use only.
int trl_debug_time_decide(trl_handle* h, int n_pending, int iters, double* ms_avg) {
    if (!h) return fail("trl_debug_time_decide: null handle");
    if (ensure_model(h)) return fail("model upload failed");
    if (n_pending > h->n) n_pending = h->n;
    std::vector<int> ids(n_pending);
    for (int i = 0; i < n_pending; ++i) ids[i] = i;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    double total = 0;
    for (int it = 0; it < iters; ++it) {
        CK(cudaMemcpyAsync(h->B.pending_list, ids.data(), (size_t)n_pending * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->B.pending_count, &n_pending, 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaEventRecord(e0, h->stream));
        enqueue_decide(h, 0, 1, h->stream);
        CK(cudaEventRecord(e1, h->stream));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        total += ms;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    if (ms_avg) *ms_avg = total / iters;
    return 0;
}

// Phase stamps of the batched decision path's FC kernel for `n_pending` decisions (measurement only): out[k] = nanoseconds from
// kernel entry to stamp k (see trl_decide2.cuh: stamp()); conv_us / fc_us = event-timed durations of the two launches
int trl_debug_fc_phases(trl_handle* h, int n_pending, double* out_ns16, double* conv_us, double* fc_us) {
    if (!h) return fail("trl_debug_fc_phases: null handle");
    if (!h->decide_v2 || !h->mc.has_net) return fail("trl_debug_fc_phases: the batched decision path is not active");
    if (ensure_model(h)) return fail("model upload failed");
    n_pending = std::max(1, std::min(n_pending, h->n));
    std::vector<int> ids(n_pending);
    for (int i = 0; i < n_pending; ++i) ids[i] = i;
    unsigned long long* prof = nullptr;
    CK(cudaMalloc((void**)&prof, 16 * 8));
    CK(cudaMemset(prof, 0, 16 * 8));
    FcMaps maps = h->fc_maps[0];
    maps.prof = prof;
    cudaEvent_t e0, e1, e2;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
    for (int it = 0; it < 3; ++it) {      // the last of three runs is reported (weights L2-resident, clocks up)
        CK(cudaMemcpyAsync(h->B.pending_list, ids.data(), (size_t)n_pending * 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaMemcpyAsync(h->B.pending_count, &n_pending, 4, cudaMemcpyHostToDevice, h->stream));
        CK(cudaEventRecord(e0, h->stream));
        launch_decide2(h->B, h->W, h->d_ex, maps, h->act2[0], h->done_count, h->decide_grid, h->fc_clusters, 0, 1, h->stream, 1);
        CK(cudaEventRecord(e1, h->stream));
        launch_decide2(h->B, h->W, h->d_ex, maps, h->act2[0], h->done_count, h->decide_grid, h->fc_clusters, 0, 1, h->stream, 2);
        CK(cudaEventRecord(e2, h->stream));
        CK(cudaEventSynchronize(e2));
    }
    h->launches += 6;
    float a = 0, b = 0;
    CK(cudaEventElapsedTime(&a, e0, e1)); CK(cudaEventElapsedTime(&b, e1, e2));
    if (conv_us) *conv_us = a * 1e3;
    if (fc_us) *fc_us = b * 1e3;
    unsigned long long st[16];
    CK(cudaMemcpy(st, prof, sizeof(st), cudaMemcpyDeviceToHost));
    for (int k = 0; k < 16; ++k) out_ns16[k] = st[k] ? (double)(st[k] - st[0]) : -1.0;
    cudaFree(prof);
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaEventDestroy(e2);
    return 0;
}

// Pipelined read-back: trl_snapshot() enqueues, behind the work already on the handle's stream, a device-side copy
// of all pose / velocity planes plus the reduced batch statistics, then moves them to pinned host memory on a second
// stream; trl_snapshot_wait() blocks on that copy only, so the caller can enqueue the next trl_update() first and
// read update k's results while update k+1 runs.
int trl_snapshot(trl_handle* h) {
    if (!h) return fail("trl_snapshot: null handle");
    const size_t plane = (size_t)h->mc.ndof * h->n;
    const size_t total = 2 * plane + 4;
    if (!h->snap_dev) {
        CK(cudaMalloc((void**)&h->snap_dev, total * 8)); h->allocs.push_back(h->snap_dev);
        CK(cudaMallocHost((void**)&h->snap_host, total * 8));
        CK(cudaStreamCreateWithFlags(&h->copy_stream, cudaStreamNonBlocking));
        CK(cudaEventCreateWithFlags(&h->snap_ready, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&h->snap_copied, cudaEventDisableTiming));
    }
    if (h->snap_pending) CK(cudaEventSynchronize(h->snap_copied));   // previous snapshot must have left the staging area
    CK(cudaMemcpyAsync(h->snap_dev, h->B.d + (size_t)D_Q * h->n, plane * 8, cudaMemcpyDeviceToDevice, h->stream));
    CK(cudaMemcpyAsync(h->snap_dev + plane, h->B.d + (size_t)D_QD * h->n, plane * 8, cudaMemcpyDeviceToDevice, h->stream));
    launch_stats(h->B, h->snap_dev + 2 * plane, h->stream);
    h->launches += 1;
    CK(cudaEventRecord(h->snap_ready, h->stream));
    CK(cudaStreamWaitEvent(h->copy_stream, h->snap_ready, 0));
    CK(cudaMemcpyAsync(h->snap_host, h->snap_dev, total * 8, cudaMemcpyDeviceToHost, h->copy_stream));
    CK(cudaEventRecord(h->snap_copied, h->copy_stream));
    h->snap_pending = true;
    return 0;
}

int trl_snapshot_wait(trl_handle* h, double* pose, double* vel, int64_t* cycles, int64_t* episodes, double* avg_dist, int64_t* env_steps) {
    if (!h) return fail("trl_snapshot_wait: null handle");
    if (!h->snap_pending) return fail("trl_snapshot_wait: no snapshot in flight");
    CK(cudaEventSynchronize(h->snap_copied));
    h->snap_pending = false;
    const size_t plane = (size_t)h->mc.ndof * h->n;
    if (pose) std::memcpy(pose, h->snap_host, plane * 8);
    if (vel) std::memcpy(vel, h->snap_host + plane, plane * 8);
    const double* st = h->snap_host + 2 * plane;
    if (cycles) *cycles = (int64_t)st[0];
    if (episodes) *episodes = (int64_t)st[1];
    if (env_steps) *env_steps = (int64_t)st[2];
    if (avg_dist) *avg_dist = st[1] > 0 ? st[3] / st[1] : 0.0;
    return 0;
}

// Raw device views of the tuple block for zero-copy hand-off to a collective (NCCL all-gather of ExpTuple blocks,
// SURVEY §8e).  Pointers are device addresses on the handle's GPU; the caller must trl_sync() before using them on
// another stream.
int trl_device_tuple_block(trl_handle* h, void** rows_f64, void** flags_u32, void** env_i32, void** count_i32, int* cap, int* width) {
    if (!h) return fail("trl_device_tuple_block: null handle");
    if (rows_f64) *rows_f64 = h->B.tuples;
    if (flags_u32) *flags_u32 = h->B.tuple_flags;
    if (env_i32) *env_i32 = h->B.tuple_env;
    if (count_i32) *count_i32 = h->B.tuple_count;
    if (cap) *cap = h->B.tuple_cap;
    if (width) *width = 1 + h->B.S + h->B.A + h->B.S;
    return 0;
}

// K outer updates timed with CUDA events on the handle's own stream (the stream the kernels are launched on);
// optionally evicts L2 between updates by writing a 256 MiB scratch buffer.
int trl_bench_updates(trl_handle* h, double dt, int k, int flush_l2, double* ms_total) {
    if (!h) return fail("trl_bench_updates: null handle");
    if (ensure_model(h)) return fail("model upload failed");
    const size_t flush_bytes = (size_t)256 << 20;
    if (flush_l2 && !h->flush_buf) { CK(cudaMalloc(&h->flush_buf, flush_bytes)); h->allocs.push_back(h->flush_buf); }
    // One event pair per update; the L2 flush (a measurement device, not part of the path) sits BETWEEN the pairs, so the reported
    // time is the sum of the K updates' own device time, every one of them starting from a cold L2.
    std::vector<cudaEvent_t> ev(2 * (size_t)std::max(k, 0));
    for (auto& e : ev) CK(cudaEventCreate(&e));
    CK(cudaStreamSynchronize(h->stream));
    for (int i = 0; i < k; ++i) {
        if (flush_l2) CK(cudaMemsetAsync(h->flush_buf, i & 0xff, flush_bytes, h->stream));
        CK(cudaEventRecord(ev[2 * i], h->stream));
        if (trl_update(h, dt)) return 1;
        CK(cudaEventRecord(ev[2 * i + 1], h->stream));
    }
    CK(cudaStreamSynchronize(h->stream));
    double ms = 0.0;
    for (int i = 0; i < k; ++i) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, ev[2 * i], ev[2 * i + 1]));
        ms += t;
    }
    if (k > 0) {
        float t = 0.f;
        CK(cudaEventElapsedTime(&t, ev[0], ev[2 * k - 1]));      // first update's start to last update's end, flushes included
        h->bench_span_ms = t;
    }
    for (auto& e : ev) cudaEventDestroy(e);
    if (ms_total) *ms_total = ms;
    return 0;
}

// the span of the last trl_bench_updates call from the first update's start to the last update's end, L2 flushes included
int trl_bench_last_span(trl_handle* h, double* ms) {
    if (!h || !ms) return fail("trl_bench_last_span: null argument");
    *ms = h->bench_span_ms;
    return 0;
}

// One outer update launched kernel by kernel with an event pair around every launch: returns the summed device time
// of the step kernel launches and of the decision kernel launches (roofline numerator's denominator).
// One outer update in the schedule trl_update uses (overlapped unless TRL_SERIAL_SCHEDULE=1), launched eagerly with an event pair
// around every launch on its own stream: out[4 * k + {0,1,2,3}] = {kind (0 terrain, 1 step, 2 decision, 3 catch-up), index, start
// ms, end ms} relative to the first launch.  Shows what bounds the update: the step launches or the decision -> catch-up chain.
int trl_update_timeline(trl_handle* h, double dt, double* out, int cap, int* n_out) {
    if (!h) return fail("trl_update_timeline: null handle");
    if (ensure_model(h)) return fail("model upload failed");
    Timeline tl;
    CK(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < h->lag; ++k) CK(cudaStreamSynchronize(h->side[k]));
    (void)cudaGetLastError();        // a stale, already reported status of an earlier call must not be charged to this run
    enqueue_update(h, dt, h->overlap, &tl);
    {
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess && tl.err.empty()) tl.err = std::string("after enqueue: ") + cudaGetErrorString(e);
        if (!tl.err.empty()) return fail("trl_update_timeline: " + tl.err);
    }
    CK(cudaStreamSynchronize(h->stream));
    for (int k = 0; k < h->lag; ++k) CK(cudaStreamSynchronize(h->side[k]));
    h->launches += update_launches(h, h->overlap);
    const int n = (int)tl.beg.size();
    for (int k = 0; k < n; ++k) {
        float a = 0, b = 0;
        cudaEventElapsedTime(&a, tl.beg[0], tl.beg[k]);
        cudaEventElapsedTime(&b, tl.beg[0], tl.end[k]);
        if (k < cap) { out[4 * k] = tl.kind[k]; out[4 * k + 1] = tl.idx[k]; out[4 * k + 2] = a; out[4 * k + 3] = b; }
    }
    for (int k = 0; k < n; ++k) { cudaEventDestroy(tl.beg[k]); cudaEventDestroy(tl.end[k]); }
    for (auto e : tl.fork) cudaEventDestroy(e);
    if (n_out) *n_out = std::min(n, cap);
    return 0;
}

static int update_timed_impl(trl_handle* h, double dt, double* step_ms, int* step_launches, double* decide_ms, int* decide_launches,
                             double* per_step, double* per_decide);
int trl_update_timed(trl_handle* h, double dt, double* step_ms, int* step_launches, double* decide_ms, int* decide_launches) {
    if (!h) return fail("trl_update_timed: null handle");
    return update_timed_impl(h, dt, step_ms, step_launches, decide_ms, decide_launches, nullptr, nullptr);
}
// same, also returning every launch's duration: per_step[num_update_steps + 1], per_decide[num_update_steps]
int trl_update_timed_detail(trl_handle* h, double dt, double* per_step, double* per_decide) {
    if (!h) return fail("trl_update_timed_detail: null handle");
    return update_timed_impl(h, dt, nullptr, nullptr, nullptr, nullptr, per_step, per_decide);
}
static int update_timed_impl(trl_handle* h, double dt, double* step_ms, int* step_launches, double* decide_ms, int* decide_launches,
                             double* per_step, double* per_decide) {
    if (ensure_model(h)) return fail("model upload failed");
    const int ns = h->num_update_steps;
    const double step = dt / ns;
    std::vector<cudaEvent_t> ev(2 * (2 * ns + 1));
    for (auto& e : ev) CK(cudaEventCreate(&e));
    int k = 0;
    launch_terrain(h->B, 0.5, h->stream);
    for (int i = 0; i < ns; ++i) {
        CK(cudaEventRecord(ev[k++], h->stream));
        launch_step(h->B, step, i == 0 ? 2 : 3, 0, h->stream);
        CK(cudaEventRecord(ev[k++], h->stream));
        CK(cudaEventRecord(ev[k++], h->stream));
        enqueue_decide(h, 0, 1, h->stream);
        CK(cudaEventRecord(ev[k++], h->stream));
    }
    CK(cudaEventRecord(ev[k++], h->stream));
    launch_step(h->B, step, 1 | 4, 0, h->stream);
    CK(cudaEventRecord(ev[k++], h->stream));
    CK(cudaStreamSynchronize(h->stream));
    CK(cudaGetLastError());
    double sm = 0, dm = 0;
    int idx = 0;
    for (int i = 0; i < ns; ++i) {
        float a = 0, b = 0;
        CK(cudaEventElapsedTime(&a, ev[idx], ev[idx + 1])); idx += 2;
        CK(cudaEventElapsedTime(&b, ev[idx], ev[idx + 1])); idx += 2;
        sm += a; dm += b;
        if (per_step) per_step[i] = a;
        if (per_decide) per_decide[i] = b;
    }
    float a = 0;
    CK(cudaEventElapsedTime(&a, ev[idx], ev[idx + 1]));
    sm += a;
    if (per_step) per_step[ns] = a;
    for (auto& e : ev) cudaEventDestroy(e);
    h->launches += (1 + num_decide_launches(h)) * ns + 2;
    if (step_ms) *step_ms = sm;
    if (step_launches) *step_launches = ns + 1;
    if (decide_ms) *decide_ms = dm;
    if (decide_launches) *decide_launches = ns;
    return 0;
}

}  // extern "C"
