// TEST INFRASTRUCTURE ONLY -- C entry points around the REFERENCE's own rigid-body code (anim/KinTree.cpp, sim/SpAlg.cpp,
// sim/RBDModel.cpp, sim/RBDUtil.cpp, util/MathUtil.cpp, util/JsonUtil.cpp), compiled where it lies under /root/reference into
// oracle/_ref/libref_rbd.so against the Eigen / jsoncpp stand-ins in oracle/ref_shim.  The character file is read by the
// reference's own loaders (cKinTree::Load / LoadBodyDefs); tests/test_ref_pinning_cpu.py compares mass matrix, bias force (with
// the reference's BuildCjPlanar), gravity force, Jacobian and centre of mass with oracle/rbd.h at random poses.
#include <fstream>
#include <string>

#include "anim/KinTree.h"
#include "sim/RBDModel.h"
#include "sim/RBDUtil.h"

struct RefRbd {
    Eigen::MatrixXd joint_mat, body_defs;
    cRBDModel model;
};

extern "C" {

RefRbd* ref_rbd_create(const char* char_file, double gx, double gy) {
    std::ifstream f(char_file);
    if (!f) return nullptr;
    Json::Value root;
    Json::Reader reader;
    if (!reader.parse(f, root)) return nullptr;
    RefRbd* r = new RefRbd();
    if (!cKinTree::Load(root["Skeleton"], r->joint_mat) || !cKinTree::LoadBodyDefs(char_file, r->body_defs)) { delete r; return nullptr; }
    r->model.Init(r->joint_mat, r->body_defs, tVector(gx, gy, 0, 0));
    return r;
}
void ref_rbd_destroy(RefRbd* r) { delete r; }
int ref_rbd_num_dof(RefRbd* r) { return r->model.GetNumDof(); }
int ref_rbd_num_joints(RefRbd* r) { return r->model.GetNumJoints(); }
// joint description columns: parent, attach x/y, limits ... exactly as the reference's loader filled them
int ref_rbd_joint_mat(RefRbd* r, double* out, int cap) {
    int n = r->joint_mat.rows() * r->joint_mat.cols();
    for (int i = 0; i < r->joint_mat.rows(); ++i)
        for (int j = 0; j < r->joint_mat.cols(); ++j)
            if (i * r->joint_mat.cols() + j < cap) out[i * r->joint_mat.cols() + j] = r->joint_mat(i, j);
    return n;
}
void ref_rbd_update(RefRbd* r, const double* pose, const double* vel) {
    const int nd = r->model.GetNumDof();
    Eigen::VectorXd p(nd), v(nd);
    for (int i = 0; i < nd; ++i) { p[i] = pose[i]; v[i] = vel[i]; }
    r->model.Update(p, v);
}
// cRBDModel::Update already ran BuildMassMat / BuildBiasForce (sim/RBDModel.cpp:39-55,259-276)
void ref_rbd_mass_bias(RefRbd* r, double* M, double* C) {
    const int nd = r->model.GetNumDof();
    const Eigen::MatrixXd& mm = r->model.GetMassMat();
    const Eigen::VectorXd& bf = r->model.GetBiasForce();
    for (int a = 0; a < nd; ++a) { for (int b = 0; b < nd; ++b) M[a * nd + b] = mm(a, b); C[a] = bf[a]; }
}
void ref_rbd_gravity_force(RefRbd* r, double* out) {
    Eigen::VectorXd g;
    cRBDUtil::CalcGravityForce(r->model, g);
    for (int i = 0; i < g.size(); ++i) out[i] = g[i];
}
// cRBDUtil::BuildJacobian: 6 x ndof, column k = spatial motion of dof k in world coordinates
void ref_rbd_jacobian(RefRbd* r, double* out) {
    Eigen::MatrixXd J;
    cRBDUtil::BuildJacobian(r->model, J);
    for (int i = 0; i < J.rows(); ++i) for (int j = 0; j < J.cols(); ++j) out[i * J.cols() + j] = J(i, j);
}
void ref_rbd_com(RefRbd* r, double* com, double* com_vel) {
    tVector c, v;
    cRBDUtil::CalcCoM(r->model, c, v);
    for (int i = 0; i < 3; ++i) { com[i] = c[i]; com_vel[i] = v[i]; }
}
void ref_rbd_joint_world_pos(RefRbd* r, int j, double* out) {
    tVector p = r->model.CalcJointWorldPos(j);
    for (int i = 0; i < 3; ++i) out[i] = p[i];
}

}  // extern "C"
