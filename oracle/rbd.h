// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/README.md). Never linked into the product path.
//
// Double-precision restatement of the reference's articulated-body math, in the reference's own generic 6-D
// spatial-vector formulation ([omega; v], 3x4 Plucker transforms [E r]).  The CUDA product path uses a different
// (planar 3-D, single-pass) formulation, so agreement between the two is a real check.
//
// Pinned against the reference's own compiled KinTree / SpAlg / RBDModel / RBDUtil (oracle/_ref/libref_rbd.so,
// tests/test_ref_pinning_cpu.py): mass matrix, bias force, gravity force, Jacobian, COM, joint positions agree to 1e-14.
//
// Follows (file:line under /root/reference):
//   sim/SpAlg.cpp:46-345            spatial cross products, transforms, compositions
//   anim/KinTree.cpp:726-757,1025-1185   param offsets, child->parent / body->joint 4x4 transforms
//   sim/RBDUtil.cpp:4-84            RNEA (SolveInvDyna)
//   sim/RBDUtil.cpp:110-176         CRBA (BuildMassMat)
//   sim/RBDUtil.cpp:250-269         world-frame Jacobian
//   sim/RBDUtil.cpp:539-649         box inertia, spatial inertia about the joint, world transforms
//   sim/RBDUtil.cpp:740-848         joint subspaces, Cj (incl. the reference's cos/cos bug), bias force
//   sim/RBDUtil.cpp:850-895         gravity force
//   sim/RBDModel.cpp:39-55          per-step cache update order
#pragma once
#include <cmath>
#include <cstring>
#include <vector>

namespace orc {

constexpr int kMaxJoints = 24;
constexpr int kMaxDof = 26;

struct V3 { double x = 0, y = 0, z = 0; };
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }
inline V3 operator*(double s, V3 a) { return {s * a.x, s * a.y, s * a.z}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }

struct M3 {
    double m[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
    static M3 ident() { M3 r; r.m[0][0] = r.m[1][1] = r.m[2][2] = 1; return r; }
};
inline V3 operator*(const M3& A, V3 v) {
    return {A.m[0][0] * v.x + A.m[0][1] * v.y + A.m[0][2] * v.z, A.m[1][0] * v.x + A.m[1][1] * v.y + A.m[1][2] * v.z,
            A.m[2][0] * v.x + A.m[2][1] * v.y + A.m[2][2] * v.z};
}
inline M3 operator*(const M3& A, const M3& B) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = A.m[i][0] * B.m[0][j] + A.m[i][1] * B.m[1][j] + A.m[i][2] * B.m[2][j];
    return r;
}
inline M3 transpose(const M3& A) {
    M3 r;
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) r.m[i][j] = A.m[j][i];
    return r;
}
// cMathUtil::RotateMat about +z (util/MathUtil.cpp:90-107)
inline M3 rot_z(double th) {
    M3 r = M3::ident();
    double c = std::cos(th), s = std::sin(th);
    r.m[0][0] = c; r.m[0][1] = -s; r.m[1][0] = s; r.m[1][1] = c;
    return r;
}
// cMathUtil::CrossMat
inline M3 cross_mat(V3 a) {
    M3 r;
    r.m[0][1] = -a.z; r.m[0][2] = a.y; r.m[1][0] = a.z; r.m[1][2] = -a.x; r.m[2][0] = -a.y; r.m[2][1] = a.x;
    return r;
}

// rigid 4x4 as (R, t)
struct Rigid { M3 R = M3::ident(); V3 t; };
inline Rigid operator*(const Rigid& a, const Rigid& b) { return {a.R * b.R, a.R * b.t + a.t}; }
inline Rigid inv_rigid(const Rigid& a) { M3 Rt = transpose(a.R); return {Rt, -(Rt * a.t)}; }
inline V3 apply(const Rigid& a, V3 p) { return a.R * p + a.t; }

struct SV { V3 o, v; };  // spatial vector [omega; v]
inline SV operator+(SV a, SV b) { return {a.o + b.o, a.v + b.v}; }
inline SV operator*(double s, SV a) { return {s * a.o, s * a.v}; }
inline double dot(SV a, SV b) { return a.o.x * b.o.x + a.o.y * b.o.y + a.o.z * b.o.z + a.v.x * b.v.x + a.v.y * b.v.y + a.v.z * b.v.z; }

struct SpTrans { M3 E = M3::ident(); V3 r; };  // cSpAlg::tSpTrans

struct M6 {
    double m[6][6];
    M6() { std::memset(m, 0, sizeof(m)); }
};
inline M6 operator*(const M6& A, const M6& B) {
    M6 r;
    for (int i = 0; i < 6; ++i)
        for (int k = 0; k < 6; ++k) {
            double a = A.m[i][k];
            if (a == 0) continue;
            for (int j = 0; j < 6; ++j) r.m[i][j] += a * B.m[k][j];
        }
    return r;
}
inline M6 operator+(const M6& A, const M6& B) {
    M6 r;
    for (int i = 0; i < 6; ++i)
        for (int j = 0; j < 6; ++j) r.m[i][j] = A.m[i][j] + B.m[i][j];
    return r;
}
inline SV operator*(const M6& A, SV s) {
    double in[6] = {s.o.x, s.o.y, s.o.z, s.v.x, s.v.y, s.v.z}, out[6];
    for (int i = 0; i < 6; ++i) {
        double acc = 0;
        for (int j = 0; j < 6; ++j) acc += A.m[i][j] * in[j];
        out[i] = acc;
    }
    return {{out[0], out[1], out[2]}, {out[3], out[4], out[5]}};
}

// ---- cSpAlg ----
inline SV crossM(SV sv, SV m) { return {cross(sv.o, m.o), cross(sv.v, m.o) + cross(sv.o, m.v)}; }      // SpAlg.cpp:46-58
inline SV crossF(SV sv, SV f) { return {cross(sv.o, f.o) + cross(sv.v, f.v), cross(sv.o, f.v)}; }      // SpAlg.cpp:73-85
inline SpTrans mat_to_trans(const Rigid& g) { return {g.R, -(transpose(g.R) * g.t)}; }                  // SpAlg.cpp:161-168
inline Rigid trans_to_mat(const SpTrans& X) { return {X.E, -(X.E * X.r)}; }                             // SpAlg.cpp:170-178
inline SpTrans inv_trans(const SpTrans& X) { return {transpose(X.E), -(X.E * X.r)}; }                   // SpAlg.cpp:209-215
inline SV apply_M(const SpTrans& X, SV s) { return {X.E * s.o, X.E * (s.v - cross(X.r, s.o))}; }        // SpAlg.cpp:245-258
inline SV apply_F(const SpTrans& X, SV s) { return {X.E * (s.o - cross(X.r, s.v)), X.E * s.v}; }        // SpAlg.cpp:260-273
inline SV apply_inv_M(const SpTrans& X, SV s) {                                                         // SpAlg.cpp:299-312
    M3 Et = transpose(X.E);
    V3 o = Et * s.o;
    return {o, Et * s.v + cross(X.r, o)};
}
inline SV apply_inv_F(const SpTrans& X, SV s) {                                                         // SpAlg.cpp:313-326
    M3 Et = transpose(X.E);
    V3 v = Et * s.v;
    return {Et * s.o + cross(X.r, v), v};
}
inline SpTrans comp_trans(const SpTrans& X0, const SpTrans& X1) {                                       // SpAlg.cpp:354-364
    return {X0.E * X1.E, X1.r + transpose(X1.E) * X0.r};
}
inline M6 spatial_mat_M(const SpTrans& X) {                                                             // SpAlg.cpp:180-192
    M6 m;
    M3 Er = X.E * cross_mat(X.r);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            m.m[i][j] = X.E.m[i][j];
            m.m[3 + i][3 + j] = X.E.m[i][j];
            m.m[3 + i][j] = -Er.m[i][j];
        }
    return m;
}
inline M6 spatial_mat_F(const SpTrans& X) {                                                             // SpAlg.cpp:194-206
    M6 m;
    M3 Er = X.E * cross_mat(X.r);
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            m.m[i][j] = X.E.m[i][j];
            m.m[3 + i][3 + j] = X.E.m[i][j];
            m.m[i][3 + j] = -Er.m[i][j];
        }
    return m;
}

// ---- skeleton tables (anim/KinTree.h:23-64) ----
enum JointType { kRevolute = 0, kPlanar = 1, kPrismatic = 2, kFixed = 3 };

struct Skeleton {
    int nj = 0, ndof = 0;
    int type[kMaxJoints], parent[kMaxJoints], offset[kMaxJoints], size[kMaxJoints];
    V3 attach[kMaxJoints];
    double lim_lo[kMaxJoints], lim_hi[kMaxJoints];
    // bodies
    int shape[kMaxJoints];
    double mass[kMaxJoints], body_theta[kMaxJoints];
    V3 body_attach[kMaxJoints], body_size[kMaxJoints];

    void init(int nj_, const double* joints /*nj x 7*/, const double* bodies /*nj x 9*/) {
        nj = nj_;
        int off = 0;
        for (int j = 0; j < nj; ++j) {
            const double* r = joints + 7 * j;
            type[j] = (int)r[0];
            parent[j] = (int)r[1];
            attach[j] = {r[2], r[3], r[4]};
            lim_lo[j] = r[5];
            lim_hi[j] = r[6];
            bool is_root = parent[j] < 0;
            // cKinTree::GetParamSize (anim/KinTree.cpp:731-757)
            int sz = 0;
            if (type[j] == kRevolute || type[j] == kPrismatic) sz = 1;
            else if (type[j] == kPlanar) sz = 3;
            else if (type[j] == kFixed) sz = is_root ? 3 : 0;
            offset[j] = off;
            size[j] = sz;
            off += sz;
            const double* b = bodies + 9 * j;
            shape[j] = (int)b[0];
            mass[j] = b[1];
            body_attach[j] = {b[2], b[3], b[4]};
            body_theta[j] = b[5];
            body_size[j] = {b[6], b[7], b[8]};
        }
        ndof = off;
    }
    bool valid_body(int j) const { return shape[j] >= 0; }
    double total_mass() const { double m = 0; for (int j = 0; j < nj; ++j) if (valid_body(j)) m += mass[j]; return m; }
};

// cKinTree::ChildParentTrans (anim/KinTree.cpp:1025-1048,1117-1149)
inline Rigid child_parent_mat(const Skeleton& sk, const double* pose, int j) {
    Rigid T0; T0.t = sk.attach[j];
    if (sk.type[j] == kRevolute) {
        Rigid R; R.R = rot_z(pose[sk.offset[j]]);
        return T0 * R;
    }
    if (sk.type[j] == kPlanar) {
        int o = sk.offset[j];
        Rigid R; R.R = rot_z(pose[o + 2]);
        Rigid T1; T1.t = {pose[o], pose[o + 1], 0};
        return (sk.parent[j] < 0) ? (T0 * T1 * R) : (T0 * R * T1);
    }
    return T0;  // fixed (non-root); prismatic is unused by any shipped character
}

// cKinTree::BodyJointTrans (anim/KinTree.cpp:1086-1098); GetBodyLocalCoM is zero for boxes
inline Rigid body_joint_mat(const Skeleton& sk, int j) {
    Rigid T; T.t = sk.body_attach[j];
    Rigid R; R.R = rot_z(sk.body_theta[j]);
    return T * R;
}

// cRBDUtil::BuildMomentInertiaBox + BuildInertiaSpatialMat (sim/RBDUtil.cpp:562-583,614-623).
// NB: as in the reference, the body's own Theta rotation is NOT applied to the inertia tensor.
inline M6 inertia_spatial_mat(const Skeleton& sk, int j) {
    double m = sk.mass[j], sx = sk.body_size[j].x, sy = sk.body_size[j].y, sz = sk.body_size[j].z;
    M6 Ic;
    Ic.m[0][0] = m / 12.0 * (sy * sy + sz * sz);
    Ic.m[1][1] = m / 12.0 * (sx * sx + sz * sz);
    Ic.m[2][2] = m / 12.0 * (sx * sx + sy * sy);
    Ic.m[3][3] = Ic.m[4][4] = Ic.m[5][5] = m;
    SpTrans X; X.r = -sk.body_attach[j];
    return spatial_mat_F(X) * Ic * spatial_mat_M(inv_trans(X));
}

// cRBDModel (sim/RBDModel.cpp:39-55) -- per-step cache
struct RBDModel {
    const Skeleton* sk = nullptr;
    V3 gravity{0, -9.8, 0};
    double pose[kMaxDof], vel[kMaxDof];
    SV S[kMaxDof];                       // joint subspace columns, indexed by dof
    Rigid child_parent[kMaxJoints];      // 4x4 child->parent
    SpTrans sp_child_parent[kMaxJoints]; // GetSpChildParentTrans
    SpTrans sp_parent_child[kMaxJoints]; // GetSpParentChildTrans
    SpTrans sp_world_joint[kMaxJoints];  // GetSpWorldJointTrans
    M6 Ibody[kMaxJoints];                // constant spatial inertias
    double M[kMaxDof][kMaxDof];
    double C[kMaxDof];
    SV J[kMaxDof];                       // world-frame Jacobian columns (cRBDUtil::BuildJacobian)
    SV link_vel[kMaxJoints], link_acc[kMaxJoints];  // RNEA intermediates of the last inv-dyn call (link frame)

    void init(const Skeleton* s, V3 g) {
        sk = s;
        gravity = g;
        for (int j = 0; j < sk->nj; ++j)
            if (sk->valid_body(j)) Ibody[j] = inertia_spatial_mat(*sk, j);
    }

    // cRBDUtil::BuildJointSubspace* (sim/RBDUtil.cpp:721-764)
    void update_subspace() {
        for (int j = 0; j < sk->nj; ++j) {
            int o = sk->offset[j];
            if (sk->type[j] == kRevolute) { S[o] = {{0, 0, 1}, {0, 0, 0}}; }
            else if (sk->type[j] == kPlanar) {
                if (sk->parent[j] < 0) {
                    M3 E = rot_z(-pose[o + 2]);
                    S[o] = {{0, 0, 0}, {E.m[0][0], E.m[1][0], 0}};
                    S[o + 1] = {{0, 0, 0}, {E.m[0][1], E.m[1][1], 0}};
                    S[o + 2] = {{0, 0, 1}, {0, 0, 0}};
                } else {
                    S[o] = {{0, 0, 0}, {1, 0, 0}};
                    S[o + 1] = {{0, 0, 0}, {0, 1, 0}};
                    S[o + 2] = {{0, 0, 1}, {0, 0, 0}};
                }
            }
        }
    }

    // cRBDUtil::BuildCjPlanar (sim/RBDUtil.cpp:808-836).  ref_bug=true reproduces the reference literally
    // (c = s = cos(theta_dot)); ref_bug=false is the correct apparent derivative dS/dt * qdot used by the physics.
    SV build_cj(int j, const double* qd, bool ref_bug) const {
        SV cj;
        if (sk->type[j] == kPlanar && sk->parent[j] < 0) {
            int o = sk->offset[j];
            double x = qd[o], y = qd[o + 1], th = qd[o + 2];
            double c, s;
            if (ref_bug) { c = std::cos(th); s = std::cos(th); }
            else { c = std::cos(pose[o + 2]); s = std::sin(pose[o + 2]); }
            cj.v = {(-s * x + c * y) * th, (-c * x - s * y) * th, 0};
        }
        return cj;
    }

    void update_kinematics(const double* q, const double* qd) {
        std::memcpy(pose, q, sizeof(double) * sk->ndof);
        std::memcpy(vel, qd, sizeof(double) * sk->ndof);
        update_subspace();
        for (int j = 0; j < sk->nj; ++j) {
            child_parent[j] = child_parent_mat(*sk, pose, j);
            sp_child_parent[j] = mat_to_trans(child_parent[j]);
            sp_parent_child[j] = mat_to_trans(inv_rigid(child_parent[j]));
        }
        // cRBDUtil::CalcWorldJointTransforms (sim/RBDUtil.cpp:625-649)
        for (int j = 0; j < sk->nj; ++j) {
            SpTrans world_parent;
            if (sk->parent[j] >= 0) world_parent = sp_world_joint[sk->parent[j]];
            sp_world_joint[j] = comp_trans(sp_parent_child[j], world_parent);
        }
    }

    // cRBDUtil::SolveInvDyna (sim/RBDUtil.cpp:4-84); a0 is the base acceleration (reference: -gravity)
    void inv_dyna(const double* acc, V3 a0, bool ref_bug_cj, double* out_tau) {
        SV fs[kMaxJoints];
        SV vel0, acc0{{0, 0, 0}, a0};
        for (int j = 0; j < sk->nj; ++j) {
            if (!sk->valid_body(j)) continue;
            int o = sk->offset[j], sz = sk->size[j];
            SV vj, sdd;
            for (int k = 0; k < sz; ++k) { vj = vj + vel[o + k] * S[o + k]; sdd = sdd + acc[o + k] * S[o + k]; }
            SV cj = build_cj(j, vel, ref_bug_cj);
            SV vp = vel0, ap = acc0;
            if (sk->parent[j] >= 0) { vp = link_vel[sk->parent[j]]; ap = link_acc[sk->parent[j]]; }
            SV cv = apply_M(sp_parent_child[j], vp) + vj;
            SV ca = apply_M(sp_parent_child[j], ap) + sdd + cj + crossM(cv, vj);
            fs[j] = Ibody[j] * ca + crossF(cv, Ibody[j] * cv);
            link_vel[j] = cv;
            link_acc[j] = ca;
        }
        for (int k = 0; k < sk->ndof; ++k) out_tau[k] = 0;
        for (int j = sk->nj - 1; j >= 0; --j) {
            if (!sk->valid_body(j)) continue;
            int o = sk->offset[j], sz = sk->size[j];
            for (int k = 0; k < sz; ++k) out_tau[o + k] = dot(S[o + k], fs[j]);
            if (sk->parent[j] >= 0) fs[sk->parent[j]] = fs[sk->parent[j]] + apply_F(sp_child_parent[j], fs[j]);
        }
    }

    // cRBDUtil::BuildMassMat, composite-rigid-body algorithm (sim/RBDUtil.cpp:110-176)
    void build_mass_mat() {
        int nj = sk->nj, nd = sk->ndof;
        M6 Is[kMaxJoints], cpF[kMaxJoints], pcM[kMaxJoints];
        for (int j = 0; j < nj; ++j) {
            if (sk->valid_body(j)) Is[j] = Ibody[j];
            cpF[j] = spatial_mat_F(sp_child_parent[j]);
            pcM[j] = spatial_mat_M(inv_trans(sp_child_parent[j]));
        }
        for (int a = 0; a < nd; ++a)
            for (int b = 0; b < nd; ++b) M[a][b] = 0;
        for (int j = nj - 1; j >= 0; --j) {
            if (!sk->valid_body(j)) continue;
            if (sk->parent[j] >= 0) Is[sk->parent[j]] = Is[sk->parent[j]] + cpF[j] * Is[j] * pcM[j];
            int o = sk->offset[j], sz = sk->size[j];
            SV F[3];
            for (int k = 0; k < sz; ++k) F[k] = Is[j] * S[o + k];
            for (int a = 0; a < sz; ++a)
                for (int b = 0; b < sz; ++b) M[o + a][o + b] = dot(S[o + a], F[b]);
            int cur = j;
            while (sk->parent[cur] >= 0) {
                for (int k = 0; k < sz; ++k) F[k] = cpF[cur] * F[k];
                cur = sk->parent[cur];
                int co = sk->offset[cur], csz = sk->size[cur];
                for (int a = 0; a < sz; ++a)
                    for (int b = 0; b < csz; ++b) {
                        double v = dot(F[a], S[co + b]);
                        M[o + a][co + b] = v;
                        M[co + b][o + a] = v;
                    }
            }
        }
    }

    // cRBDUtil::BuildJacobian (sim/RBDUtil.cpp:250-269)
    void build_jacobian() {
        for (int j = 0; j < sk->nj; ++j) {
            int o = sk->offset[j];
            for (int k = 0; k < sk->size[j]; ++k) J[o + k] = apply_inv_M(sp_world_joint[j], S[o + k]);
        }
    }

    // cRBDModel::Update + cRBDUtil::BuildJacobian as called from cDogController::UpdateRBDModel
    void update(const double* q, const double* qd) {
        update_kinematics(q, qd);
        build_mass_mat();
        double zero[kMaxDof] = {0};
        inv_dyna(zero, -gravity, /*ref_bug_cj=*/true, C);  // BuildBiasForce (sim/RBDUtil.cpp:844-848)
        build_jacobian();
    }

    // cRBDUtil::CalcGravityForce (sim/RBDUtil.cpp:850-895)
    void gravity_force(double* out) const {
        SV fs[kMaxJoints];
        SV acc0{{0, 0, 0}, gravity};
        for (int j = 0; j < sk->nj; ++j)
            if (sk->valid_body(j)) fs[j] = Ibody[j] * apply_M(sp_world_joint[j], acc0);
        for (int k = 0; k < sk->ndof; ++k) out[k] = 0;
        for (int j = sk->nj - 1; j >= 0; --j) {
            if (!sk->valid_body(j)) continue;
            int o = sk->offset[j];
            for (int k = 0; k < sk->size[j]; ++k) out[o + k] = dot(S[o + k], fs[j]);
            if (sk->parent[j] >= 0) fs[sk->parent[j]] = fs[sk->parent[j]] + apply_F(sp_child_parent[j], fs[j]);
        }
    }

    // world position of joint j's origin: cRBDModel::CalcJointWorldPos = GetRad(world_joint_trans)
    V3 joint_world_pos(int j) const { return sp_world_joint[j].r; }
    // joint->world rigid transform
    Rigid joint_world_mat(int j) const { return trans_to_mat(inv_trans(sp_world_joint[j])); }
};

// Dense LDL^T solve of a symmetric positive-definite system A x = b (stands in for Eigen's ldlt().solve();
// sim/ImpPDController.cpp:271, sim/RBDUtil.cpp:98).  A is overwritten.
inline void ldlt_solve(int n, double A[][kMaxDof], const double* b, double* x) {
    double d[kMaxDof];
    for (int j = 0; j < n; ++j) {
        double dj = A[j][j];
        for (int k = 0; k < j; ++k) dj -= A[j][k] * A[j][k] * d[k];
        d[j] = dj;
        for (int i = j + 1; i < n; ++i) {
            double v = A[i][j];
            for (int k = 0; k < j; ++k) v -= A[i][k] * A[j][k] * d[k];
            A[i][j] = v / dj;
        }
    }
    for (int i = 0; i < n; ++i) {
        double v = b[i];
        for (int k = 0; k < i; ++k) v -= A[i][k] * x[k];
        x[i] = v;
    }
    for (int i = 0; i < n; ++i) x[i] /= d[i];
    for (int i = n - 1; i >= 0; --i) {
        double v = x[i];
        for (int k = i + 1; k < n; ++k) v -= A[k][i] * x[k];
        x[i] = v;
    }
}

}  // namespace orc
