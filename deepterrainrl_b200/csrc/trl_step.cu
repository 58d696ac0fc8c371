// deepterrainrl_b200 -- the env-step kernel (sm_100a).
//
// One environment per warp lane, one warp per CTA, SoA state in HBM (coalesced 256-B plane reads), per-link
// quantities staged in shared memory transposed [slot][lane] so that a dynamic link index never causes a bank
// conflict (lane is the fastest-varying index; f64 => two conflict-free wavefronts per access).
//
// One launch = the *controller half* of env-step k followed by the *physics half* of env-step k+1:
//
//   ctrl half  (reference: cSimCharacter::Update -> cDogController::Update, sim/DogController.cpp:229-268)
//     planar CRBA mass matrix + RNEA bias force        (restates sim/RBDUtil.cpp:4-176 in 3-D planar algebra)
//     ApplyFeedback, stable-PD solve (LDL^T 23x23), gravity compensation, virtual forces
//     torque clamp (sim/Joint.cpp:257-264), fall counters (sim/SimCharSoftFall.cpp:74-125)
//     cycle bookkeeping: reward, tuple record, episode statistics (scenarios/ScenarioExp.cpp:209-243)
//   [end of outer update: fall -> episode reset, scenarios/ScenarioPoliEval.cpp:110-125, ScenarioExp.cpp:83-98]
//   phys half  (reference: cWorld::Update + cGroundVar2D::Update + the head of the controller update)
//     num_sim_substeps x planar articulated-body forward dynamics with implicit contact / joint-limit terms
//     (this project's own physics model -- Bullet is not restatable), semi-implicit Euler
//     contact bits, streaming terrain update, cycle timers, gait FSM (sim/DogController.cpp:805-845)
//     lanes whose FSM reaches a cycle boundary build the 283-float policy state and enqueue a decision
//
// The decision itself (MACE forward pass + action decode) runs in trl_decide.cu between two launches of this kernel;
// because the controller half consumes the decided action at the start of the *next* launch, all lanes stay
// convergent and no tail kernel is needed.
#include <cuda_runtime.h>

#include "trl_terrain.cuh"
#include "trl_types.h"

namespace trl {

__constant__ ModelConst c_model;

// ------------------------------------------------------------------------------------------------ scratch layout
// per-lane shared-memory slots (each slot = one f64 per lane)
enum Slot : int {
    S_Q = 0,
    S_QD = S_Q + kMaxDof,
    S_TAU = S_QD + kMaxDof,
    S_WC = S_TAU + kMaxDof,          // world cos of joint frame
    S_WS = S_WC + kMaxJoints,        // world sin
    S_WX = S_WS + kMaxJoints,        // world origin x
    S_WY = S_WX + kMaxJoints,
    S_JVX = S_WY + kMaxJoints,       // world velocity of joint origin
    S_JVY = S_JVX + kMaxJoints,
    S_JW = S_JVY + kMaxJoints,       // world angular velocity of link
    S_JC = S_JW + kMaxJoints,        // cos / sin of each joint angle (child -> parent rotation)
    S_JS = S_JC + kMaxJoints,
    S_COMMON_END = S_JS + kMaxJoints,
    // ---- controller phase
    C_M = S_COMMON_END,              // lower-triangular mass matrix, 276
    C_C = C_M + kMaxDof * (kMaxDof + 1) / 2,
    C_RHS = C_C + kMaxDof,
    C_ACC = C_RHS + kMaxDof,
    C_TAUC = C_ACC + kMaxDof,
    C_UNION = C_TAUC + kMaxDof,      // 189-slot union: {Ic[4][21]} | {LV,LA,LF [3][21] each} | {BASIS[23][4], TG[23]}
    C_END = C_UNION + 9 * kMaxJoints,
    // ---- physics phase (aliases the controller region)
    P_V = S_COMMON_END,              // link velocity (w, vx, vy) in link coords
    P_CV = P_V + 3 * kMaxJoints,     // velocity-product term (cx, cy)
    P_IA = P_CV + 2 * kMaxJoints,    // articulated inertia (a, bx, by, cxx, cxy, cyy)
    P_PA = P_IA + 6 * kMaxJoints,    // articulated bias force (n, fx, fy)
    P_U = P_PA + 3 * kMaxJoints,
    P_DINV = P_U + 3 * kMaxJoints,
    P_UU = P_DINV + kMaxJoints,
    P_A = P_UU + kMaxJoints,         // link acceleration
    P_END = P_A + 3 * kMaxJoints,
    S_NUM = (C_END > P_END ? C_END : P_END)
};
static_assert(S_NUM * kWarp * 8 <= 227 * 1024, "per-warp scratch exceeds shared memory");

#define SL(slot) sm[(slot) * kWarp]

__device__ __forceinline__ int tri(int a, int b) { return a * (a + 1) / 2 + b; }  // a >= b

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

// ------------------------------------------------------------------------------------------------ lane context
struct Lane {
    double* sm;          // shared scratch base for this lane
    int env, n;          // env index, number of envs
    double* D;           // f64 planes
    int* I;              // i32 planes
    __device__ __forceinline__ double& d(int f) { return D[(size_t)f * n + env]; }
    __device__ __forceinline__ int& i(int f) { return I[(size_t)f * n + env]; }
};

// forward kinematics of the joint frames: world rotation, origin, origin velocity, angular velocity
__device__ void fk_world(Lane& L) {
    double* sm = L.sm;
    const ModelConst& m = c_model;
    {
        double s, c;
        sincos(SL(S_Q + 2), &s, &c);
        SL(S_WC) = c; SL(S_WS) = s; SL(S_WX) = SL(S_Q); SL(S_WY) = SL(S_Q + 1);
        SL(S_JVX) = SL(S_QD); SL(S_JVY) = SL(S_QD + 1); SL(S_JW) = SL(S_QD + 2);
    }
    for (int j = 1; j < m.nj; ++j) {
        int p = m.parent[j], o = m.dof[j];
        double pc = SL(S_WC + p), ps = SL(S_WS + p);
        double ax = pc * m.attach_x[j] - ps * m.attach_y[j], ay = ps * m.attach_x[j] + pc * m.attach_y[j];
        double s, c;
        sincos(SL(S_Q + o), &s, &c);
        SL(S_JC + j) = c; SL(S_JS + j) = s;
        SL(S_WC + j) = pc * c - ps * s;
        SL(S_WS + j) = ps * c + pc * s;
        SL(S_WX + j) = SL(S_WX + p) + ax;
        SL(S_WY + j) = SL(S_WY + p) + ay;
        double pw = SL(S_JW + p);
        SL(S_JVX + j) = SL(S_JVX + p) - pw * ay;
        SL(S_JVY + j) = SL(S_JVY + p) + pw * ax;
        SL(S_JW + j) = pw + SL(S_QD + o);
    }
}
// body COM world position / velocity of link j
__device__ __forceinline__ void body_kin(double* sm, int j, double& px, double& py, double& vx, double& vy) {
    const ModelConst& m = c_model;
    double c = SL(S_WC + j), s = SL(S_WS + j);
    double bx = c * m.body_ax[j] - s * m.body_ay[j], by = s * m.body_ax[j] + c * m.body_ay[j];
    px = SL(S_WX + j) + bx; py = SL(S_WY + j) + by;
    double w = SL(S_JW + j);
    vx = SL(S_JVX + j) - w * by; vy = SL(S_JVY + j) + w * bx;
}
__device__ void calc_com(double* sm, double& cx, double& cy, double& vx, double& vy) {
    const ModelConst& m = c_model;
    cx = cy = vx = vy = 0.0;
    for (int j = 0; j < m.nj; ++j) {
        double px, py, bvx, bvy;
        body_kin(sm, j, px, py, bvx, bvy);
        cx += m.mass[j] * px; cy += m.mass[j] * py; vx += m.mass[j] * bvx; vy += m.mass[j] * bvy;
    }
    double inv = 1.0 / m.total_mass;
    cx *= inv; cy *= inv; vx *= inv; vy *= inv;
}
// bottom-centre of a foot box (cDogController::GetEndEffectorContactPos)
__device__ __forceinline__ void effector_pos(double* sm, int j, double& ex, double& ey) {
    const ModelConst& m = c_model;
    double px, py, vx, vy;
    body_kin(sm, j, px, py, vx, vy);
    double c = SL(S_WC + j) * m.body_cos[j] - SL(S_WS + j) * m.body_sin[j];
    double s = SL(S_WS + j) * m.body_cos[j] + SL(S_WC + j) * m.body_sin[j];
    double ly = -m.half_y[j];
    ex = px - s * ly; ey = py + c * ly;
}

__device__ __forceinline__ bool has_fallen(Lane& L, double root_theta) {
    return L.d(D_SUM_FALL) > 0.25 || L.i(I_FAIL_FALL_DIST) != 0 || fabs(wrap_pi(root_theta)) > 3.14159265358979323846 * 0.8;
}

// cDogController::SetStateParams for the current state (sim/DogController.cpp:1042-1054)
__device__ void set_state_params(Lane& L, int state) {
    int base = D_PARAMS + mMiscMax + state * spMax;
    double sc = L.d(base + spSpineCurve);
    L.d(D_PD_TARGET + jSpine0) = sc; L.d(D_PD_TARGET + jSpine1) = sc; L.d(D_PD_TARGET + jSpine2) = sc;
    L.d(D_PD_TARGET + jSpine3) = sc; L.d(D_PD_TARGET + jTorso) = sc;
    L.d(D_PD_TARGET + jShoulder) = L.d(base + spShoulder);
    L.d(D_PD_TARGET + jElbow) = L.d(base + spElbow);
    L.d(D_PD_TARGET + jHip) = L.d(base + spHip);
    L.d(D_PD_TARGET + jKnee) = L.d(base + spKnee);
    L.d(D_PD_TARGET + jAnkle) = L.d(base + spAnkle);
}

// ================================================================================================ controller half
// Planar restatement of cRBDModel::Update (CRBA + RNEA with the reference's Cj), the stable-PD solve, gravity
// compensation and virtual forces.  Writes the clamped torques to S_TAU.
__device__ void controller_torque(Lane& L, double h, int contact) {
    double* sm = L.sm;
    const ModelConst& m = c_model;
    const int nj = m.nj, nd = m.ndof;

    // pose as the controller sees it: root angle wrapped (axis-angle extraction), hinges as is
    const double th0 = wrap_pi(SL(S_Q + 2));
    double s0, c0;
    sincos(th0, &s0, &c0);

    // ---- CRBA: composite rigid-body inertias (m, hx, hy, I) in link coordinates, accumulated leaf -> root
    const int IC = C_UNION;
    for (int j = 0; j < nj; ++j) {
        SL(IC + 4 * j + 0) = m.mass[j];
        SL(IC + 4 * j + 1) = m.mass[j] * m.body_ax[j];
        SL(IC + 4 * j + 2) = m.mass[j] * m.body_ay[j];
        SL(IC + 4 * j + 3) = m.izz_o[j];
    }
    for (int k = 0; k < nd * (nd + 1) / 2; ++k) SL(C_M + k) = 0.0;
    for (int j = nj - 1; j >= 0; --j) {
        double mj = SL(IC + 4 * j), hx = SL(IC + 4 * j + 1), hy = SL(IC + 4 * j + 2), Iz = SL(IC + 4 * j + 3);
        int p = m.parent[j];
        if (p >= 0) {
            // shift the composite inertia into the parent's frame: rotate first moment, translate by attach
            double c = SL(S_JC + j), s = SL(S_JS + j);
            double hxp = c * hx - s * hy, hyp = s * hx + c * hy;
            double ax = m.attach_x[j], ay = m.attach_y[j];
            SL(IC + 4 * p + 0) += mj;
            SL(IC + 4 * p + 1) += hxp + mj * ax;
            SL(IC + 4 * p + 2) += hyp + mj * ay;
            SL(IC + 4 * p + 3) += Iz + 2.0 * (ax * hxp + ay * hyp) + mj * (ax * ax + ay * ay);
        }
        if (j > 0) {
            // F = Ic * S (S = unit rotation): force (n, fx, fy) in link j coords
            double fn = Iz, fx = -hy, fy = hx;
            int oj = m.dof[j];
            SL(C_M + tri(oj, oj)) = fn;
            int cur = j;
            while (cur > 0) {
                double c = SL(S_JC + cur), s = SL(S_JS + cur);
                double gx = c * fx - s * fy, gy = s * fx + c * fy;
                fn = fn + m.attach_x[cur] * gy - m.attach_y[cur] * gx;
                fx = gx; fy = gy;
                cur = m.parent[cur];
                if (cur > 0) SL(C_M + tri(oj, m.dof[cur])) = fn;
                else {
                    // root columns: S = [Rz(-th)^T e_x, Rz(-th)^T e_y, e_w] in root coords
                    SL(C_M + tri(oj, 0)) = c0 * fx - s0 * fy;
                    SL(C_M + tri(oj, 1)) = s0 * fx + c0 * fy;
                    SL(C_M + tri(oj, 2)) = fn;
                }
            }
        } else {
            // root block S^T Ic S with S = [(0; c, -s), (0; s, c), (1; 0, 0)]
            // Ic = [[Iz, -hy, hx], [-hy, m, 0], [hx, 0, m]]
            double ex0 = c0, ey0 = -s0, ex1 = s0, ey1 = c0;   // linear parts of columns 0, 1 (root coords)
            SL(C_M + tri(0, 0)) = mj;
            SL(C_M + tri(1, 0)) = 0.0;
            SL(C_M + tri(1, 1)) = mj;
            SL(C_M + tri(2, 0)) = -hy * ex0 + hx * ey0;
            SL(C_M + tri(2, 1)) = -hy * ex1 + hx * ey1;
            SL(C_M + tri(2, 2)) = Iz;
        }
    }

    // ---- RNEA bias force with qdd = 0, base acceleration -g, and the reference's BuildCjPlanar (cos/cos)
    const int LV = C_UNION, LA = C_UNION + 3 * kMaxJoints, LF = C_UNION + 6 * kMaxJoints;
    {
        double xd = SL(S_QD), yd = SL(S_QD + 1), thd = SL(S_QD + 2);
        double cb = cos(thd), sb = cb;   // sim/RBDUtil.cpp:821-822
        double cjx = (-sb * xd + cb * yd) * thd, cjy = (-cb * xd - sb * yd) * thd;
        // v = S qd (root coords); a = X(-g) + cj
        SL(LV + 0) = thd; SL(LV + 1) = c0 * xd + s0 * yd; SL(LV + 2) = -s0 * xd + c0 * yd;
        double gx = -m.gx, gy = -m.gy;
        SL(LA + 0) = 0.0; SL(LA + 1) = c0 * gx + s0 * gy + cjx; SL(LA + 2) = -s0 * gx + c0 * gy + cjy;
    }
    for (int j = 1; j < nj; ++j) {
        int p = m.parent[j], o = m.dof[j];
        double c = SL(S_JC + j), s = SL(S_JS + j);
        double ax = m.attach_x[j], ay = m.attach_y[j], qd = SL(S_QD + o);
        double pw = SL(LV + 3 * p), pvx = SL(LV + 3 * p + 1) - pw * ay, pvy = SL(LV + 3 * p + 2) + pw * ax;
        double w = pw + qd, vx = c * pvx + s * pvy, vy = -s * pvx + c * pvy;
        double pa = SL(LA + 3 * p), pax = SL(LA + 3 * p + 1) - pa * ay, pay = SL(LA + 3 * p + 2) + pa * ax;
        // crossM(v, vj), vj = (qd, 0, 0): linear part v_lin x (qd z)
        SL(LV + 3 * j) = w; SL(LV + 3 * j + 1) = vx; SL(LV + 3 * j + 2) = vy;
        SL(LA + 3 * j) = pa;
        SL(LA + 3 * j + 1) = c * pax + s * pay + vy * qd;
        SL(LA + 3 * j + 2) = -s * pax + c * pay - vx * qd;
    }
    for (int j = 0; j < nj; ++j) {
        double mj = m.mass[j], hx = mj * m.body_ax[j], hy = mj * m.body_ay[j], Iz = m.izz_o[j];
        double w = SL(LV + 3 * j), vx = SL(LV + 3 * j + 1), vy = SL(LV + 3 * j + 2);
        double aw = SL(LA + 3 * j), ax = SL(LA + 3 * j + 1), ay = SL(LA + 3 * j + 2);
        double hn = Iz * w - hy * vx + hx * vy, hpx = mj * vx - hy * w, hpy = mj * vy + hx * w;
        SL(LF + 3 * j) = Iz * aw - hy * ax + hx * ay + (vx * hpy - vy * hpx);
        SL(LF + 3 * j + 1) = mj * ax - hy * aw - w * hpy;
        SL(LF + 3 * j + 2) = mj * ay + hx * aw + w * hpx;
        (void)hn;
    }
    for (int j = nj - 1; j >= 1; --j) {
        int p = m.parent[j], o = m.dof[j];
        double fn = SL(LF + 3 * j), fx = SL(LF + 3 * j + 1), fy = SL(LF + 3 * j + 2);
        SL(C_C + o) = fn;
        double c = SL(S_JC + j), s = SL(S_JS + j);
        double gx = c * fx - s * fy, gy = s * fx + c * fy;
        SL(LF + 3 * p) += fn + m.attach_x[j] * gy - m.attach_y[j] * gx;
        SL(LF + 3 * p + 1) += gx;
        SL(LF + 3 * p + 2) += gy;
    }
    {
        double fn = SL(LF), fx = SL(LF + 1), fy = SL(LF + 2);
        SL(C_C + 0) = c0 * fx - s0 * fy;
        SL(C_C + 1) = s0 * fx + c0 * fy;
        SL(C_C + 2) = fn;
    }

    // ---- ApplyFeedback (sim/DogController.cpp:903-945)
    const int state = L.i(I_STATE);
    double comx, comy, comvx, comvy;
    calc_com(sm, comx, comy, comvx, comvy);
    {
        double cv = L.d(D_PARAMS + mCv);
        int base = D_PARAMS + mMiscMax + state * spMax;
        if (!((contact >> jToe) & 1)) L.d(D_PD_TARGET + jHip) = L.d(base + spHip) + comvx * cv;
        if (!((contact >> jFinger) & 1)) L.d(D_PD_TARGET + jShoulder) = L.d(base + spShoulder) + comvx * cv;
    }

    // ---- cImpPDController::CalcControlForces (sim/ImpPDController.cpp:234-278)
    SL(C_RHS) = -SL(C_C); SL(C_RHS + 1) = -SL(C_C + 1); SL(C_RHS + 2) = -SL(C_C + 2);
    SL(C_TAUC) = 0.0; SL(C_TAUC + 1) = 0.0; SL(C_TAUC + 2) = 0.0;
    for (int j = 1; j < nj; ++j) {
        int o = m.dof[j];
        double theta = SL(S_Q + o);
        if (m.world_pd[j]) {
            // child body's world rotation, wrapped (cPDController::CalcTheta, sim/PDController.cpp:181-198)
            double c = SL(S_WC + j) * m.body_cos[j] - SL(S_WS + j) * m.body_sin[j];
            double s = SL(S_WS + j) * m.body_cos[j] + SL(S_WC + j) * m.body_sin[j];
            double a = acos(fmin(1.0, fmax(-1.0, c)));
            theta = (s >= 0) ? a : -a;
        }
        double qd = SL(S_QD + o);
        double perr = L.d(D_PD_TARGET + j) - theta, verr = m.target_vel[j] - qd;
        double t0 = m.kp[j] * (perr - h * qd);
        SL(C_TAUC + o) = t0 + m.kd[j] * verr;          // tau = Kp(e - h qd) + Kd(ev - h acc): acc term added below
        SL(C_RHS + o) = t0 + m.kd[j] * verr - SL(C_C + o);
        SL(C_M + tri(o, o)) += h * m.kd[j];
    }
    // in-place LDL^T of the lower triangle, then solve
    for (int j = 0; j < nd; ++j) {
        double dj = SL(C_M + tri(j, j));
        for (int k = 0; k < j; ++k) { double l = SL(C_M + tri(j, k)); dj -= l * l * SL(C_M + tri(k, k)); }
        SL(C_M + tri(j, j)) = dj;
        double inv = 1.0 / dj;
        for (int i = j + 1; i < nd; ++i) {
            double v = SL(C_M + tri(i, j));
            for (int k = 0; k < j; ++k) v -= SL(C_M + tri(i, k)) * SL(C_M + tri(j, k)) * SL(C_M + tri(k, k));
            SL(C_M + tri(i, j)) = v * inv;
        }
    }
    for (int i = 0; i < nd; ++i) {
        double v = SL(C_RHS + i);
        for (int k = 0; k < i; ++k) v -= SL(C_M + tri(i, k)) * SL(C_ACC + k);
        SL(C_ACC + i) = v;
    }
    for (int i = 0; i < nd; ++i) SL(C_ACC + i) /= SL(C_M + tri(i, i));
    for (int i = nd - 1; i >= 0; --i) {
        double v = SL(C_ACC + i);
        for (int k = i + 1; k < nd; ++k) v -= SL(C_M + tri(k, i)) * SL(C_ACC + k);
        SL(C_ACC + i) = v;
    }
    for (int j = 1; j < nj; ++j) {
        int o = m.dof[j];
        SL(C_TAUC + o) -= m.kd[j] * h * SL(C_ACC + o);
    }

    // ---- ApplyGravityCompensation (sim/DogController.cpp:947-995, 1120-1175)
    const bool toe_c = (contact >> jToe) & 1, fin_c = (contact >> jFinger) & 1;
    if (m.grav_comp && (toe_c || fin_c)) {
        const int BAS = C_UNION, TG = C_UNION + 4 * kMaxDof;
        for (int k = 0; k < 4 * nd; ++k) SL(BAS + k) = 0.0;
        for (int e = 0; e < 2; ++e) {
            int ej = e == 0 ? jToe : jFinger;
            if (!((contact >> ej) & 1)) continue;
            double ex, ey;
            effector_pos(sm, ej, ex, ey);
            // column 2e: unit +y force, column 2e+1: unit +x force; entry = J_k^T f = torque of f about joint k
            for (int cur = ej; cur >= 0; cur = m.parent[cur]) {
                if (cur > 0) {
                    int o = m.dof[cur];
                    double rx = ex - SL(S_WX + cur), ry = ey - SL(S_WY + cur);
                    SL(BAS + 4 * o + 2 * e) = rx;        // (r x (0,1))
                    SL(BAS + 4 * o + 2 * e + 1) = -ry;   // (r x (1,0))
                } else {
                    double rx = ex - SL(S_WX), ry = ey - SL(S_WY);
                    SL(BAS + 4 * 0 + 2 * e) = 0.0; SL(BAS + 4 * 0 + 2 * e + 1) = 1.0;
                    SL(BAS + 4 * 1 + 2 * e) = 1.0; SL(BAS + 4 * 1 + 2 * e + 1) = 0.0;
                    SL(BAS + 4 * 2 + 2 * e) = rx;  SL(BAS + 4 * 2 + 2 * e + 1) = -ry;
                }
            }
        }
        // tau_g = -CalcGravityForce: per-link gravity wrench, accumulated towards the root
        // (generalised force of gravity acting as an acceleration field +g; sim/RBDUtil.cpp:850-895)
        for (int k = 0; k < nd; ++k) SL(TG + k) = 0.0;
        {
            // world-frame accumulation: torque about joint k of the weights of all bodies in its subtree
            // subtree mass moments via one leaf -> root sweep in world coordinates
            // reuse RHS slots as temporaries: (msum, mx, my) per link packed in C_RHS is too small -> use LA region
            const int SUB = C_UNION + 5 * kMaxDof;   // 3 * nj slots (fits: 5*23 + 63 = 178 <= 189)
            for (int j = 0; j < nj; ++j) {
                double px, py, vx, vy;
                body_kin(sm, j, px, py, vx, vy);
                SL(SUB + 3 * j) = m.mass[j]; SL(SUB + 3 * j + 1) = m.mass[j] * px; SL(SUB + 3 * j + 2) = m.mass[j] * py;
            }
            for (int j = nj - 1; j >= 1; --j) {
                int p = m.parent[j];
                SL(SUB + 3 * p) += SL(SUB + 3 * j); SL(SUB + 3 * p + 1) += SL(SUB + 3 * j + 1); SL(SUB + 3 * p + 2) += SL(SUB + 3 * j + 2);
            }
            // generalised gravity force G_k = sum_subtree (r_com - p_k) x (m g); tau_g = -G
            for (int j = 1; j < nj; ++j) {
                double ms = SL(SUB + 3 * j), mx = SL(SUB + 3 * j + 1), my = SL(SUB + 3 * j + 2);
                double rx = mx - ms * SL(S_WX + j), ry = my - ms * SL(S_WY + j);
                SL(TG + m.dof[j]) = -(rx * m.gy - ry * m.gx);
            }
            double ms = SL(SUB), mx = SL(SUB + 1), my = SL(SUB + 2);
            double rx = mx - ms * SL(S_WX), ry = my - ms * SL(S_WY);
            SL(TG + 0) = -(ms * m.gx); SL(TG + 1) = -(ms * m.gy); SL(TG + 2) = -(rx * m.gy - ry * m.gx);
        }
        // ridge least squares on the root rows: (A^T A + 1e-4 I) x = A^T b, A = basis[0:3, :], b = tau_g[0:3]
        double A[4][5];
        for (int a = 0; a < 4; ++a) {
            double atb = 0.0;
            for (int r = 0; r < 3; ++r) atb += SL(BAS + 4 * r + a) * SL(TG + r);
            for (int b = 0; b < 4; ++b) {
                double v = 0.0;
                for (int r = 0; r < 3; ++r) v += SL(BAS + 4 * r + a) * SL(BAS + 4 * r + b);
                A[a][b] = v;
            }
            A[a][a] += 0.0001;
            A[a][4] = atb;
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
                    for (int k = 0; k < 5; ++k) { double t = A[c][k]; A[c][k] = A[r][k]; A[r][k] = t; }
                }
            }
#pragma unroll
            for (int r = c + 1; r < 4; ++r) {
                double f = A[r][c] / A[c][c];
#pragma unroll
                for (int k = c; k < 5; ++k) A[r][k] -= f * A[c][k];
            }
        }
        double x[4];
#pragma unroll
        for (int r = 3; r >= 0; --r) {
            double v = A[r][4];
#pragma unroll
            for (int k = r + 1; k < 4; ++k) v -= A[r][k] * x[k];
            x[r] = v / A[r][r];
        }
        for (int a = 3; a < nd; ++a) {
            double tc = SL(BAS + 4 * a) * x[0] + SL(BAS + 4 * a + 1) * x[1] + SL(BAS + 4 * a + 2) * x[2] + SL(BAS + 4 * a + 3) * x[3];
            SL(C_TAUC + a) += SL(TG + a) - tc;
        }
    }

    // ---- ApplyVirtualForces (sim/DogController.cpp:997-1029)
    if (m.virt_forces) {
        for (int e = 0; e < 2; ++e) {
            int ej = e == 0 ? jToe : jFinger;
            bool valid = (e == 0) ? (state == sBackStance || state == sExtend) : (state == sFrontStance || state == sGather);
            if (!(valid && ((contact >> ej) & 1))) continue;
            double fx = -L.d(D_PARAMS + (e == 0 ? mBackForceX : mFrontForceX));
            double fy = -L.d(D_PARAMS + (e == 0 ? mBackForceY : mFrontForceY));
            double ex, ey;
            effector_pos(sm, ej, ex, ey);
            for (int cur = ej; cur != jRoot && cur != jTorso; cur = m.parent[cur]) {
                double rx = ex - SL(S_WX + cur), ry = ey - SL(S_WY + cur);
                SL(C_TAUC + m.dof[cur]) += rx * fy - ry * fx;
            }
        }
    }

    // ---- cJoint::ApplyTorque clamp (sim/Joint.cpp:171-201,257-264)
    SL(S_TAU) = 0.0; SL(S_TAU + 1) = 0.0; SL(S_TAU + 2) = 0.0;
    for (int j = 1; j < nj; ++j) {
        int o = m.dof[j];
        double t = SL(C_TAUC + o), lim = m.torque_lim[j];
        if (fabs(t) > lim) t *= lim / fabs(t);
        SL(S_TAU + o) = t;
    }
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
    }
    return 0.8 * vel_r + 0.2 * stum_r;
}

// cScenarioExp::NewCycleUpdate (scenarios/ScenarioExp.cpp:209-243): finish the previous tuple, start the next
__device__ void exp_new_cycle_update(Lane& L, const Buffers& B, bool fallen) {
    const int S = B.S, A = 1 + (kNumParams - 1);
    const double* s_end = B.poli_state + (size_t)L.env * S;
    double* s_beg = B.tuple_sbeg + (size_t)L.env * S;
    double* act = B.tuple_action + (size_t)L.env * kNumParams;
    unsigned flags = (unsigned)L.i(I_TUPLE_FLAGS);
    flags = fallen ? (flags | 1u) : (flags & ~1u);
    double reward = calc_reward(L, fallen);
    if (L.i(I_CYCLE_COUNT) > 1) {
        int slot = atomicAdd(B.tuple_count, 1);
        if (slot < B.tuple_cap) {
            double* row = B.tuples + (size_t)slot * (1 + S + A + S);
            row[0] = reward;
            for (int k = 0; k < S; ++k) row[1 + k] = s_beg[k];
            for (int k = 0; k < A; ++k) row[1 + S + k] = act[k];
            for (int k = 0; k < S; ++k) row[1 + S + A + k] = s_end[k];
            B.tuple_flags[slot] = flags;
            B.tuple_env[slot] = L.env;
        }
    }
    for (int k = 0; k < S; ++k) s_beg[k] = s_end[k];
    act[0] = (double)L.i(I_ACTION_ID);
    for (int k = 1; k < kNumParams; ++k) act[k] = L.d(D_PARAMS + k);
    int ef = L.i(I_EXP_FLAGS);
    unsigned nf = 0;
    if (ef & 1) nf |= 2u;   // exp critic -> eFlagExpCritic (bit 1)
    if (ef & 2) nf |= 4u;   // exp actor  -> eFlagExpActor  (bit 2)
    L.i(I_TUPLE_FLAGS) = (int)nf;
    L.i(I_CYCLE_COUNT) += 1;
}

// cDogController::BlendCtrlParams / BuildBaseAction (+ cDogControllerMACE::AssignFragID); writes params, returns id
__device__ int build_base_action(Lane& L, CounterRng& rng, int a, double* params /*local[30]*/) {
    const ModelConst& m = c_model;
    int i0 = m.act_idx0[a], i1 = m.act_idx1[a];
    double blend = m.act_blend[a];
    for (int k = 0; k < kNumParams; ++k) {
        double p0 = m.ctrl_params[i0][k], p1 = m.ctrl_params[i1][k];
        if (k == mTransTime || k == mCv) { p0 = fabs(p0); p1 = fabs(p1); }
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
    for (int k = 0; k < kNumParams; ++k) L.d(D_PARAMS + k) = params[k];
    L.d(D_PARAMS + mTransTime) = fabs(params[mTransTime]);
    L.d(D_PARAMS + mCv) = fabs(params[mCv]);
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

// cScenarioSimChar::Reset (+ PoliEval / Exp specifics): scenarios/ScenarioSimChar.cpp:121-132
__device__ void reset_env(Lane& L, const Buffers& B) {
    double* sm = L.sm;
    const ModelConst& m = c_model;
    for (int k = 0; k < m.ndof; ++k) { SL(S_Q + k) = m.pose0[k]; SL(S_QD + k) = m.vel0[k]; SL(S_TAU + k) = 0.0; }
    L.i(I_CONTACT) = 0;
    fk_world(L);
    double comx, comy, cvx, cvy;
    calc_com(sm, comx, comy, cvx, cvy);
    CounterRng rng = load_rng(L);
    // controller reset: default action, FSM state 0, counters zeroed (sim/TerrainRLCharController.cpp:47-58,
    // sim/DogController.cpp:210-216,640-650)
    double params[kNumParams];
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
    L.d(D_FALL_DIST_CNT) = 5.0; L.d(D_PREV_CHECK_X) = SL(S_Q); L.d(D_PREV_CHECK_Y) = SL(S_Q + 1);
    L.i(I_FAIL_FALL_DIST) = 0; L.d(D_FALL_CONTACT_CNT) = 0.1; L.d(D_SUM_FALL) = 0.0;
    // ResetGround + InitCharacterPos
    GroundView g = load_ground(L, B);
    g.n[0] = g.n[1] = 0; g.flip = 0;
    g.update(-11.0, 9.0, m.terrain_type, m.terrain_params, 20.0);
    if (m.has_init_x) SL(S_Q) = m.init_x;
    SL(S_Q + 1) += g.sample(SL(S_Q));
    store_ground(L, g);
    if (m.exp_mode) {
        L.i(I_CYCLE_COUNT) = 0;
        L.i(I_CMD) = rng.rand_int(0, m.n_actions);   // cScenarioExp::CommandRandAction
    } else {
        L.d(D_POS_START_X) = SL(S_Q);
    }
    store_rng(L, rng);
}

// ================================================================================================ physics half
// One sub-step of planar articulated-body forward dynamics with linearly-implicit contact / joint-limit terms.
__device__ int physics_substep(Lane& L, const GroundView& g, double dt, bool write_contacts) {
    double* sm = L.sm;
    const ModelConst& m = c_model;
    const PhysParams& pp = m.phys;
    const int nj = m.nj;
    int contact = 0;

    // pass 1: kinematics outward (link velocities in link coords, world frames, velocity-product terms, rigid
    // inertias, bias forces incl. gravity as an external force)
    {
        double s, c;
        sincos(SL(S_Q + 2), &s, &c);
        SL(S_WC) = c; SL(S_WS) = s; SL(S_WX) = SL(S_Q); SL(S_WY) = SL(S_Q + 1);
        double xd = SL(S_QD), yd = SL(S_QD + 1);
        SL(P_V) = SL(S_QD + 2); SL(P_V + 1) = c * xd + s * yd; SL(P_V + 2) = -s * xd + c * yd;
    }
    for (int j = 1; j < nj; ++j) {
        int p = m.parent[j], o = m.dof[j];
        double s, c;
        sincos(SL(S_Q + o), &s, &c);
        SL(S_JC + j) = c; SL(S_JS + j) = s;
        double pc = SL(S_WC + p), ps = SL(S_WS + p);
        double ax = m.attach_x[j], ay = m.attach_y[j];
        SL(S_WC + j) = pc * c - ps * s;
        SL(S_WS + j) = ps * c + pc * s;
        SL(S_WX + j) = SL(S_WX + p) + pc * ax - ps * ay;
        SL(S_WY + j) = SL(S_WY + p) + ps * ax + pc * ay;
        double qd = SL(S_QD + o);
        double pw = SL(P_V + 3 * p), pvx = SL(P_V + 3 * p + 1) - pw * ay, pvy = SL(P_V + 3 * p + 2) + pw * ax;
        double vx = c * pvx + s * pvy, vy = -s * pvx + c * pvy;
        SL(P_V + 3 * j) = pw + qd; SL(P_V + 3 * j + 1) = vx; SL(P_V + 3 * j + 2) = vy;
        SL(P_CV + 2 * j) = vy * qd; SL(P_CV + 2 * j + 1) = -vx * qd;
    }
    for (int j = 0; j < nj; ++j) {
        double mj = m.mass[j], hx = mj * m.body_ax[j], hy = mj * m.body_ay[j], Iz = m.izz_o[j];
        SL(P_IA + 6 * j) = Iz; SL(P_IA + 6 * j + 1) = -hy; SL(P_IA + 6 * j + 2) = hx;
        SL(P_IA + 6 * j + 3) = mj; SL(P_IA + 6 * j + 4) = 0.0; SL(P_IA + 6 * j + 5) = mj;
        double w = SL(P_V + 3 * j), vx = SL(P_V + 3 * j + 1), vy = SL(P_V + 3 * j + 2);
        double hpx = mj * vx - hy * w, hpy = mj * vy + hx * w;
        // gravity in link coords
        double wc = SL(S_WC + j), ws = SL(S_WS + j);
        double glx = wc * m.gx + ws * m.gy, gly = -ws * m.gx + wc * m.gy;
        SL(P_PA + 3 * j) = (vx * hpy - vy * hpx) - (hx * gly - hy * glx);
        SL(P_PA + 3 * j + 1) = -w * hpy - mj * glx;
        SL(P_PA + 3 * j + 2) = w * hpx - mj * gly;
    }

    // contacts: box corners against the height field
    for (int j = 0; j < nj; ++j) {
        if (!m.collidable[j]) continue;
        double wc = SL(S_WC + j), ws = SL(S_WS + j), wx = SL(S_WX + j), wy = SL(S_WY + j);
        double w = SL(P_V + 3 * j), vx = SL(P_V + 3 * j + 1), vy = SL(P_V + 3 * j + 2);
        double bc = m.body_cos[j], bs = m.body_sin[j], hx = m.half_x[j], hy = m.half_y[j];
#pragma unroll 1
        for (int cn = 0; cn < 4; ++cn) {
            double bx = (cn & 1) ? hx : -hx, by = (cn & 2) ? hy : -hy;
            double lx = m.body_ax[j] + bc * bx - bs * by, ly = m.body_ay[j] + bs * bx + bc * by;   // link coords
            double px = wx + wc * lx - ws * ly, py = wy + ws * lx + wc * ly;
            double slope;
            double hgt = g.sample(px, &slope);
            double inv = rsqrt(1.0 + slope * slope);
            double pen = (hgt - py) * inv;
            if (pen <= -pp.contact_tol) continue;
            contact |= (1 << j);
            if (pen <= 0.0) continue;
            // world normal / tangent -> link coords
            double nxw = -slope * inv, nyw = inv;
            double nx = wc * nxw + ws * nyw, ny = -ws * nxw + wc * nyw;   // n_l = R_w^T n_w
            double tx = ny, ty = -nx;                                       // t_w = (inv, slope inv) -> R^T t_w = (ny, -nx)
            double pvx = vx - w * ly, pvy = vy + w * lx;                    // point velocity, link coords
            double vn = pvx * nx + pvy * ny, vt = pvx * tx + pvy * ty;
            double fn0 = pp.kn * pen - pp.dn * vn;
            if (fn0 <= 0.0) continue;
            double cnn = pp.dn + dt * pp.kn;
            double ctt = pp.mu * fn0 / fmax(fabs(vt), pp.v_eps);
            double fwx = pp.kn * pen * nx - cnn * vn * nx - ctt * vt * tx;
            double fwy = pp.kn * pen * ny - cnn * vn * ny - ctt * vt * ty;
            SL(P_PA + 3 * j) -= lx * fwy - ly * fwx;
            SL(P_PA + 3 * j + 1) -= fwx;
            SL(P_PA + 3 * j + 2) -= fwy;
            double dxx = dt * (cnn * nx * nx + ctt * tx * tx), dxy = dt * (cnn * nx * ny + ctt * tx * ty),
                   dyy = dt * (cnn * ny * ny + ctt * ty * ty);
            double kx = -ly, ky = lx;
            double dkx = dxx * kx + dxy * ky, dky = dxy * kx + dyy * ky;
            SL(P_IA + 6 * j) += kx * dkx + ky * dky;
            SL(P_IA + 6 * j + 1) += dkx; SL(P_IA + 6 * j + 2) += dky;
            SL(P_IA + 6 * j + 3) += dxx; SL(P_IA + 6 * j + 4) += dxy; SL(P_IA + 6 * j + 5) += dyy;
        }
    }

    // pass 2: articulated inertias / bias forces inward
    for (int j = nj - 1; j >= 1; --j) {
        int p = m.parent[j], o = m.dof[j];
        double a = SL(P_IA + 6 * j), bx = SL(P_IA + 6 * j + 1), by = SL(P_IA + 6 * j + 2);
        double cxx = SL(P_IA + 6 * j + 3), cxy = SL(P_IA + 6 * j + 4), cyy = SL(P_IA + 6 * j + 5);
        double pn = SL(P_PA + 3 * j), pfx = SL(P_PA + 3 * j + 1), pfy = SL(P_PA + 3 * j + 2);
        double Dj = a, u = SL(S_TAU + o) - pn;
        if (m.has_limit[j]) {
            double q = SL(S_Q + o), viol = 0.0;
            if (q > m.lim_hi[j]) viol = q - m.lim_hi[j];
            else if (q < m.lim_lo[j]) viol = q - m.lim_lo[j];
            if (viol != 0.0) {
                double cl = pp.d_lim + dt * pp.k_lim;
                u += -pp.k_lim * viol - cl * SL(S_QD + o);
                Dj += dt * cl;
            }
        }
        double dinv = 1.0 / Dj;
        SL(P_U + 3 * j) = a; SL(P_U + 3 * j + 1) = bx; SL(P_U + 3 * j + 2) = by;
        SL(P_DINV + j) = dinv; SL(P_UU + j) = u;
        // Ia = IA - U U^T / D
        double ia = a - a * a * dinv, ibx = bx - a * bx * dinv, iby = by - a * by * dinv;
        double icxx = cxx - bx * bx * dinv, icxy = cxy - bx * by * dinv, icyy = cyy - by * by * dinv;
        // pa = pA + Ia c + U u / D,  c = (0, cvx, cvy)
        double cvx = SL(P_CV + 2 * j), cvy = SL(P_CV + 2 * j + 1), ud = u * dinv;
        double qn = pn + ibx * cvx + iby * cvy + a * ud;
        double qx = pfx + icxx * cvx + icxy * cvy + bx * ud;
        double qy = pfy + icxy * cvx + icyy * cvy + by * ud;
        // rotate into parent axes
        double c = SL(S_JC + j), s = SL(S_JS + j);
        double rbx = c * ibx - s * iby, rby = s * ibx + c * iby;
        double t1 = c * icxx - s * icxy, t2 = c * icxy - s * icyy, t3 = s * icxx + c * icxy, t4 = s * icxy + c * icyy;
        double rxx = t1 * c - t2 * s, rxy = t1 * s + t2 * c, ryy = t3 * s + t4 * c;
        double rfx = c * qx - s * qy, rfy = s * qx + c * qy;
        // shift by k = (-ay, ax)
        double kx = -m.attach_y[j], ky = m.attach_x[j];
        double ckx = rxx * kx + rxy * ky, cky = rxy * kx + ryy * ky;
        SL(P_IA + 6 * p) += ia + 2.0 * (kx * rbx + ky * rby) + kx * ckx + ky * cky;
        SL(P_IA + 6 * p + 1) += rbx + ckx; SL(P_IA + 6 * p + 2) += rby + cky;
        SL(P_IA + 6 * p + 3) += rxx; SL(P_IA + 6 * p + 4) += rxy; SL(P_IA + 6 * p + 5) += ryy;
        SL(P_PA + 3 * p) += qn + kx * rfx + ky * rfy;
        SL(P_PA + 3 * p + 1) += rfx; SL(P_PA + 3 * p + 2) += rfy;
    }

    // floating base: solve IA_0 a_0 = -pA_0 (symmetric 3x3)
    {
        double a = SL(P_IA), bx = SL(P_IA + 1), by = SL(P_IA + 2), cxx = SL(P_IA + 3), cxy = SL(P_IA + 4), cyy = SL(P_IA + 5);
        double r0 = -SL(P_PA), r1 = -SL(P_PA + 1), r2 = -SL(P_PA + 2);
        // LDL^T
        double d0 = a, l10 = bx / d0, l20 = by / d0;
        double d1 = cxx - l10 * l10 * d0, l21 = (cxy - l20 * l10 * d0) / d1;
        double d2 = cyy - l20 * l20 * d0 - l21 * l21 * d1;
        double y0 = r0, y1 = r1 - l10 * y0, y2 = r2 - l20 * y0 - l21 * y1;
        double z2 = y2 / d2, z1 = y1 / d1 - l21 * z2, z0 = y0 / d0 - l10 * z1 - l20 * z2;
        SL(P_A) = z0; SL(P_A + 1) = z1; SL(P_A + 2) = z2;
    }
    // pass 3: accelerations outward, integrate (semi-implicit Euler)
    for (int j = 1; j < nj; ++j) {
        int p = m.parent[j], o = m.dof[j];
        double c = SL(S_JC + j), s = SL(S_JS + j), ax = m.attach_x[j], ay = m.attach_y[j];
        double pa = SL(P_A + 3 * p), pax = SL(P_A + 3 * p + 1) - pa * ay, pay = SL(P_A + 3 * p + 2) + pa * ax;
        double aw = pa, alx = c * pax + s * pay + SL(P_CV + 2 * j), aly = -s * pax + c * pay + SL(P_CV + 2 * j + 1);
        double qdd = (SL(P_UU + j) - (SL(P_U + 3 * j) * aw + SL(P_U + 3 * j + 1) * alx + SL(P_U + 3 * j + 2) * aly)) * SL(P_DINV + j);
        SL(P_A + 3 * j) = aw + qdd; SL(P_A + 3 * j + 1) = alx; SL(P_A + 3 * j + 2) = aly;
        double qd = SL(S_QD + o) + dt * qdd;
        SL(S_QD + o) = qd;
        SL(S_Q + o) += dt * qd;
    }
    {
        // root: classical acceleration of the origin in world axes = R (a_lin + w x v_lin)
        double c = SL(S_WC), s = SL(S_WS);
        double w = SL(P_V), vx = SL(P_V + 1), vy = SL(P_V + 2);
        double alx = SL(P_A + 1) - w * vy, aly = SL(P_A + 2) + w * vx;
        double xdd = c * alx - s * aly, ydd = s * alx + c * aly, thdd = SL(P_A);
        double xd = SL(S_QD) + dt * xdd, yd = SL(S_QD + 1) + dt * ydd, thd = SL(S_QD + 2) + dt * thdd;
        SL(S_QD) = xd; SL(S_QD + 1) = yd; SL(S_QD + 2) = thd;
        SL(S_Q) += dt * xd; SL(S_Q + 1) += dt * yd; SL(S_Q + 2) += dt * thd;
    }
    return contact;
}

// cTerrainRLCharController::ParseGround + BuildPoliState (sim/TerrainRLCharController.cpp:168-285)
__device__ void build_poli_state(Lane& L, const Buffers& B, const GroundView& g) {
    double* sm = L.sm;
    const ModelConst& m = c_model;
    double* out = B.poli_state + (size_t)L.env * B.S;
    double ox = SL(S_Q), oy = g.sample(SL(S_Q));
    for (int i = 0; i < kNumGroundSamples; ++i) {
        double dist = ((10.0 - (-0.5)) * i) / (kNumGroundSamples - 1) + (-0.5);
        out[i] = g.sample(dist + ox) - oy;
    }
    int idx = kNumGroundSamples;
    out[idx++] = SL(S_Q + 1) - oy;
    for (int j = 1; j < m.nj; ++j) {
        double px, py, vx, vy;
        body_kin(sm, j, px, py, vx, vy);
        out[idx++] = px - SL(S_Q); out[idx++] = py - SL(S_Q + 1);
    }
    for (int j = 0; j < m.nj; ++j) {
        double px, py, vx, vy;
        body_kin(sm, j, px, py, vx, vy);
        out[idx++] = vx; out[idx++] = vy;
    }
}

// ================================================================================================ the kernel
// flags: bit0 do_ctrl (finish env-step k), bit1 do_phys (start env-step k+1), bit2 end of outer update
__global__ void __launch_bounds__(kWarp, 1)
trl_step_kernel(Buffers B, double h, int flags) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x;
    const int env = blockIdx.x * kWarp + lane;
    if (env >= B.n) return;
    Lane L{smem + lane, env, B.n, B.d, B.i};
    double* sm = L.sm;
    const ModelConst& m = c_model;
    const int nd = m.ndof;

    for (int k = 0; k < nd; ++k) { SL(S_Q + k) = L.d(D_Q + k); SL(S_QD + k) = L.d(D_QD + k); SL(S_TAU + k) = L.d(D_TAU + k); }
    int contact = L.i(I_CONTACT);

    if (flags & 1) {
        // ---------------- controller half of env-step k
        fk_world(L);
        controller_torque(L, h, contact);
        // fall checks (sim/SimCharSoftFall.cpp:74-125)
        double cnt = L.d(D_FALL_DIST_CNT) - h;
        if (cnt <= 0.0) {
            double dx = SL(S_Q) - L.d(D_PREV_CHECK_X), dy = SL(S_Q + 1) - L.d(D_PREV_CHECK_Y);
            if (dx * dx + dy * dy < 0.25) L.i(I_FAIL_FALL_DIST) = 1;
            L.d(D_PREV_CHECK_X) = SL(S_Q); L.d(D_PREV_CHECK_Y) = SL(S_Q + 1);
            cnt = 5.0;
        }
        L.d(D_FALL_DIST_CNT) = cnt;
        double cc = L.d(D_FALL_CONTACT_CNT) - h;
        if (cc <= 0.0) {
            const int fall_mask = (1 << jRoot) | (1 << jSpine0) | (1 << jSpine1) | (1 << jSpine2) | (1 << jSpine3) |
                                  (1 << jTorso) | (1 << jNeck0) | (1 << jNeck1) | (1 << jHead);
            double val = (contact & fall_mask) ? 1.0 : 0.0;
            const double norm = (1.0 + 1.0 / (1.0 - 0.9));
            L.d(D_SUM_FALL) = val / norm + 0.9 * L.d(D_SUM_FALL);
            cc = 0.1;
        }
        L.d(D_FALL_CONTACT_CNT) = cc;
        // PostSubstepUpdate: cycle boundary bookkeeping
        if (L.i(I_STATE) == 0 && L.d(D_PHASE) == 0.0) {
            if (m.exp_mode) exp_new_cycle_update(L, B, has_fallen(L, SL(S_Q + 2)));
            else L.i(I_CYCLE_COUNT) += 1;
        }
        unsigned lo = (unsigned)L.i(I_STEPS_LO) + 1u;
        L.i(I_STEPS_LO) = (int)lo;
        if (lo == 0u) L.i(I_STEPS_HI) += 1;
    }

    if (flags & 4) {
        // ---------------- end of the outer update: fall -> episode bookkeeping + reset
        bool fallen = has_fallen(L, SL(S_Q + 2));
        bool new_cycle = (L.i(I_STATE) == 0 && L.d(D_PHASE) == 0.0);
        if (m.exp_mode) {
            if (!new_cycle && fallen) { exp_new_cycle_update(L, B, true); reset_env(L, B); }
        } else if (fallen) {
            if (L.i(I_CYCLE_COUNT) >= 1) {
                double dist = SL(S_Q) - L.d(D_POS_START_X);
                int ec = L.i(I_EPISODE_COUNT);
                L.d(D_AVG_DIST) = (ec * L.d(D_AVG_DIST) + dist) / (ec + 1.0);
                L.i(I_EPISODE_COUNT) = ec + 1;
                int slot = atomicAdd(B.dist_count, 1);
                if (slot < B.dist_cap) { B.dist_log[slot] = dist; B.dist_env[slot] = env; }
            }
            reset_env(L, B);
        }
        contact = L.i(I_CONTACT);
    }

    if (flags & 2) {
        // ---------------- physics half of env-step k+1
        GroundView g = load_ground(L, B);
        const int ns = m.num_sim_substeps;
        const double dt = h / ns;
        for (int s = 0; s < ns; ++s) {
            int cmask = physics_substep(L, g, dt, s == ns - 1);
            if (s == ns - 1) contact = cmask;
        }
        L.i(I_CONTACT) = contact;
        // UpdateGround (scenarios/ScenarioSimChar.cpp:564-572)
        if (g.update(SL(S_Q) - 2.0, SL(S_Q) + 10.0 + 1.0, m.terrain_type, m.terrain_params, 20.0)) store_ground(L, g);
        // head of cDogController::Update: cycle timers, stumble counter, gait FSM
        L.d(D_CUR_CYCLE_T) += h;
        const int stumble_mask = ~((1 << jToe) | (1 << jFinger) | (1 << jAnkle) | (1 << jWrist));
        if (contact & stumble_mask) L.d(D_CUR_STUMBLE) += h;
        int state = L.i(I_STATE);
        int first = L.i(I_FIRST_CYCLE);
        double phase = L.d(D_PHASE) + h / L.d(D_PARAMS + mTransTime);
        bool advance = first != 0;
        if ((state == sBackStance || state == sFrontStance) && phase >= 1.0) advance = true;
        if (state == sExtend && ((contact >> jFinger) & 1)) advance = true;
        if (state == sGather && ((contact >> jToe) & 1)) advance = true;
        L.d(D_PHASE) = phase;
        if (advance) {
            int ns2 = first ? sBackStance : (state == sGather ? -1 : state + 1);
            bool end_step = (ns2 < 0) || first;
            if (end_step) {
                // cycle boundary: build the policy state and hand the env to the decision kernel
                fk_world(L);
                build_poli_state(L, B, g);
                double comx, comy, cvx, cvy;
                calc_com(sm, comx, comy, cvx, cvy);
                B.com_stash[env] = comx; B.com_stash[B.n + env] = comy;
                L.i(I_FIRST_CYCLE) = 0;
                L.i(I_PENDING) = 1;
                int slot = atomicAdd(B.pending_count, 1);
                B.pending_list[slot] = env;
            } else {
                L.i(I_STATE) = ns2;
                L.d(D_PHASE) = 0.0;
                set_state_params(L, ns2);
            }
        }
    }

    for (int k = 0; k < nd; ++k) { L.d(D_Q + k) = SL(S_Q + k); L.d(D_QD + k) = SL(S_QD + k); L.d(D_TAU + k) = SL(S_TAU + k); }
}

// Initial reset of every env (trl_create / trl_reset): seeds the terrain RNG and runs the episode reset.
__global__ void __launch_bounds__(kWarp, 1)
trl_reset_kernel(Buffers B, const uint64_t* terrain_seeds, const int* env_ids, int count, int reseed) {
    extern __shared__ double smem[];
    const int lane = threadIdx.x;
    const int idx = blockIdx.x * kWarp + lane;
    if (idx >= count) return;
    const int env = env_ids ? env_ids[idx] : idx;
    Lane L{smem + lane, env, B.n, B.d, B.i};
    double* sm = L.sm;
    const ModelConst& m = c_model;
    if (reseed) {
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
    reset_env(L, B);
    for (int k = 0; k < m.ndof; ++k) { L.d(D_Q + k) = SL(S_Q + k); L.d(D_QD + k) = SL(S_QD + k); L.d(D_TAU + k) = SL(S_TAU + k); }
}

// ---- host-side launch helpers (called from trl_host.cu)
cudaError_t upload_model(const ModelConst& mc) { return cudaMemcpyToSymbol(c_model, &mc, sizeof(ModelConst)); }
size_t step_smem_bytes() { return (size_t)S_NUM * kWarp * sizeof(double); }
cudaError_t configure_step_kernels() {
    cudaError_t e = cudaFuncSetAttribute(trl_step_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem_bytes());
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(trl_reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)step_smem_bytes());
}
void launch_step(const Buffers& B, double h, int flags, cudaStream_t st) {
    int blocks = (B.n + kWarp - 1) / kWarp;
    trl_step_kernel<<<blocks, kWarp, step_smem_bytes(), st>>>(B, h, flags);
}
void launch_reset(const Buffers& B, const uint64_t* seeds, const int* env_ids, int count, int reseed, cudaStream_t st) {
    int blocks = (count + kWarp - 1) / kWarp;
    trl_reset_kernel<<<blocks, kWarp, step_smem_bytes(), st>>>(B, seeds, env_ids, count, reseed);
}

}  // namespace trl

#include "trl_decide.cuh"
