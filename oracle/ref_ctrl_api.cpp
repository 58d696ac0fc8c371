// TEST INFRASTRUCTURE ONLY -- the REFERENCE's own sources for the hot path, compiled where they lie under /root/reference into
// oracle/_ref/libref_ctrl.so (recipe: oracle/Makefile, REF_CTRL_SRCS) against header stand-ins for Eigen / jsoncpp / Bullet / Caffe
// (oracle/ref_shim):
//   * the controller stack (sim/Controller, CharController, NNController, TerrainRLCharController, Dog / Goat / Raptor controllers
//     incl. the Q / Cacla / MACE layers, ImpPDController, PDController, Joint), the characters (anim/Character, sim/SimCharSoftFall,
//     SimDog, SimRaptor), the kinematics and rigid-body sources of libref_rbd, the ground (sim/Ground, GroundVar2D, TerrainGen2D);
//   * the scenarios (scenarios/Scenario, ScenarioSimChar, ScenarioPoliEval, ScenarioExp, ScenarioExpMACE, ScenarioTrain,
//     ScenarioTrainMACE), util/ArgParser, learning/ExpTuple;
//   * the trainer (learning/TrainerInterface, NeuralNetTrainer, MACETrainer, NeuralNetLearner);
//   * the evaluation driver (optimizer/scenarios/OptScenarioPoliEval) and, compiled against the reference's headers, the product's
//     C++ adapter (include/terrainrl_b200_adapter.h): the drop-in seam as a maintainer would build it (end of this file,
//     tests/test_ref_adapter_cpu.py).
// What those sources need from the simulation -- pose, velocity, contacts, body-part positions and rotations, the world step, the
// network -- they obtain through virtual calls on cSimCharacter / cSimObj / cJoint / cWorld and through cNeuralNet; this file supplies
// that back end from a state the test installs (the CPU oracle's state), using the reference's own cKinTree kinematics for positions
// and velocities, and hands the world step and the network operations back to the test (ref_net_standin.h).
// tests/test_ref_pinning_cpu.py compares what the compiled reference code then computes -- torques, gait machine, policy states,
// rewards, fall verdicts, terrain, tuples, statistics, resets, exploration draws, trainer buffers and weights, annealing schedule --
// with oracle/env.h, oracle/trainer.h and the product's host functions.
//
// Every other virtual function of the Bullet / Caffe-backed classes is resolved to ref_abort_stub by the link recipe
// (link_with_stubs.sh): if the compiled reference code ever reached one, the test would abort instead of silently using made-up
// behaviour.
#include <execinfo.h>
#include <signal.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "learning/MACETrainer.h"
#include "learning/NeuralNet.h"
#include "sim/DogController.h"
#include "sim/DogControllerQ.h"
#include "sim/DogControllerMACE.h"
#include "sim/GoatControllerMACE.h"
#include "sim/RaptorControllerMACE.h"
#include "sim/RaptorControllerQ.h"
#include "sim/Ground.h"
#include "sim/GroundVar2D.h"
#include "sim/SimCharacter.h"
#include "sim/SimDog.h"
#include "sim/SimRaptor.h"
#include "scenarios/ScenarioExpMACE.h"
#include "scenarios/ScenarioPoliEval.h"
#include "scenarios/ScenarioTrainMACE.h"
#include "util/ArgParser.h"

extern "C" void ref_abort_stub() {
    std::fprintf(stderr, "oracle/_ref: the compiled reference code called a Bullet/Caffe-backed function that has no stand-in\n");
    void* frames[16];
    backtrace_symbols_fd(frames, backtrace(frames, 16), 2);
    std::abort();
}

static void ref_segv_handler(int) {
    void* frames[24];
    backtrace_symbols_fd(frames, backtrace(frames, 24), 2);
    _exit(139);
}
#include <unistd.h>

// ------------------------------------------------------------------------------------------------ state installed by the test
struct FakeChar;
typedef double (*height_fn)(double x, void* user);

struct FakePart : public cSimObj {
    FakeChar* owner = nullptr;
    int id = 0;
    bool contact = false;
    tVector GetPos() const override;
    tVector GetLinearVelocity() const override;
    tVector LocalToWorldPos(const tVector& local_pos) const override;
    bool IsInContact() const override { return contact; }
};

struct FakeJoint : public cJoint {
    FakeChar* owner = nullptr;
    int id = 0;
    bool valid = true;
    tVector axis_rel = tVector(0, 0, 1, 0);
    bool IsValid() const override { return valid; }
    const tVector& GetAxisRel() const override { return axis_rel; }
    tVector CalcAxisWorld() const override { return axis_rel; }                       // planar character: every hinge is about z
    void CalcRotation(tVector& out_axis, double& out_theta) const override;           // joint angle
    void GetChildRotation(tVector& out_axis, double& out_theta) const override;       // world rotation of the child link
    tVector CalcJointVelRel() const override;                                         // relative angular velocity in the joint frame
    // AddTorque / GetTorque / ClearTorque / SetTorqueLimit / ClampTotalTorque are the reference's own (sim/Joint.cpp, compiled)
};

// the state the test installs + the answers to the character's virtual calls; Base is the reference's own character class
// (cSimDog / cSimRaptor), so fall and stumble detection (cSimCharSoftFall, cSimDog::HasStumbled, ...) is the reference's code
struct FakeCharData {
    Eigen::VectorXd pose, vel, last_tau;
    std::vector<std::shared_ptr<cSimObj>> parts;
    std::vector<FakeJoint> joints;
    tVector com = tVector(0, 0, 0, 0), com_vel = tVector(0, 0, 0, 0);
    int n_apply = 0;
    virtual ~FakeCharData() {}
    virtual cSimCharacter* sim() = 0;
    virtual void soft_fall_update(double h) = 0;
    virtual void soft_fall_reset() = 0;
    virtual bool load(const std::string& char_file, const std::string& state_file) = 0;
};
struct FakeChar : public FakeCharData {};          // name used by the part / joint stand-ins

template <typename Base>
struct FakeCharT : public Base, public FakeChar {
    cSimCharacter* sim() override { return this; }
    void soft_fall_update(double h) override { Base::cSimCharSoftFall::Update(h); }
    void soft_fall_reset() override { Base::cSimCharSoftFall::Reset(); }
    bool Load(const std::string& char_file) {
        if (!cCharacter::Init(char_file)) return false;                // skeleton via the reference's own loader
        if (!cKinTree::LoadBodyDefs(char_file, this->mBodyDefs)) return false;
        const int nj = this->GetNumJoints();
        joints.resize(nj);
        for (int j = 0; j < nj; ++j) {
            auto p = std::make_shared<FakePart>();
            p->owner = this; p->id = j;
            parts.push_back(p);
            joints[j].owner = this; joints[j].id = j;
            joints[j].valid = cKinTree::HasParent(this->mJointMat, j);  // the root has no actuated joint
        }
        pose = Eigen::VectorXd::Zero(this->GetNumDof());
        vel = Eigen::VectorXd::Zero(this->GetNumDof());
        last_tau = Eigen::VectorXd::Zero(this->GetNumDof());
        return true;
    }
    bool load(const std::string& char_file, const std::string& state_file) override {
        if (!Load(char_file)) return false;
        if (state_file != "" && !this->ReadState(state_file)) return false;      // cCharacter::ReadState -> SetPose / SetVel below
        this->RecordDefaultState();                                               // pose0 / vel0 for Reset
        return true;
    }
    void Clear() override { cCharacter::Clear(); }
    void SetRootPos(const tVector& pos) override { cKinTree::SetRootPos(this->mJointMat, pos, pose); cCharacter::SetPose(pose); }
    void SetPose(const Eigen::VectorXd& p) override { pose = p; cCharacter::SetPose(p); }
    void SetVel(const Eigen::VectorXd& v) override { vel = v; cCharacter::SetVel(v); }
    void BuildPose(Eigen::VectorXd& out) const override { out = pose; }
    void BuildVel(Eigen::VectorXd& out) const override { out = vel; }
    tVector GetRootPos() const override { return cKinTree::GetRootPos(this->mJointMat, pose); }
    // root orientation / angular rate as the reference extracts them from the root body (sim/SimCharacter.cpp:118-128,181-225):
    // the planar root joint's angle and rate
    void GetRootRotation(tVector& out_axis, double& out_theta) const override {
        out_axis = tVector(0, 0, 1, 0);
        out_theta = cKinTree::GetRootTheta(this->mJointMat, pose);
    }
    tVector GetRootAngVel() const override { return tVector(0, 0, vel[2], 0); }
    tVector GetRootVel() const override { return tVector(vel[0], vel[1], 0, 0); }
    const Eigen::MatrixXd& GetBodyDefs() const override { return this->mBodyDefs; }
    int GetNumBodyParts() const override { return (int)parts.size(); }
    // cSimCharacter::CalcCOM / CalcCOMVel (sim/SimCharacter.cpp:396-434): mass-weighted body-part positions / velocities
    tVector CalcCOM() const override {
        tVector c = tVector::Zero();
        double total = 0;
        for (int i = 0; i < (int)parts.size(); ++i) {
            if (!IsValidBodyPart(i)) continue;
            const double m = cKinTree::GetBodyMass(this->mBodyDefs, i);
            c += m * parts[i]->GetPos();
            total += m;
        }
        c /= total;
        return c;
    }
    tVector CalcCOMVel() const override {
        tVector c = tVector::Zero();
        double total = 0;
        for (int i = 0; i < (int)parts.size(); ++i) {
            if (!IsValidBodyPart(i)) continue;
            const double m = cKinTree::GetBodyMass(this->mBodyDefs, i);
            c += m * parts[i]->GetLinearVelocity();
            total += m;
        }
        c /= total;
        return c;
    }
    const cJoint& GetJoint(int j) const override { return joints[j]; }
    cJoint& GetJoint(int j) override { return joints[j]; }
    const std::shared_ptr<cSimObj>& GetBodyPart(int i) const override { return parts[i]; }
    std::shared_ptr<cSimObj>& GetBodyPart(int i) override { return parts[i]; }
    bool IsValidBodyPart(int idx) const override { return cKinTree::IsValidBody(this->mBodyDefs, idx); }
    // cSimCharacter::ApplyControlForces (sim/SimCharacter.cpp:637-654): one z-torque per valid joint, accumulated in the joint
    // (cSimCharacter::ClearJointTorques empties the accumulators after the world step, here before the next accumulation)
    void ApplyControlForces(const Eigen::VectorXd& tau) override {
        last_tau = tau;
        ++n_apply;
        for (int j = 0; j < this->GetNumJoints(); ++j) {
            if (!joints[j].IsValid()) continue;
            joints[j].ClearTorque();
            joints[j].AddTorque(tVector(0, 0, tau[this->GetParamOffset(j)], 0));
        }
    }
};

static const Eigen::MatrixXd& jm(const FakeChar* c) { return const_cast<FakeChar*>(c)->sim()->GetJointMat(); }
static const Eigen::MatrixXd& bd(const FakeChar* c) { return const_cast<FakeChar*>(c)->sim()->GetBodyDefs(); }

tVector FakePart::GetPos() const { return cKinTree::CalcBodyPartPos(jm(owner), owner->pose, bd(owner), id); }
tVector FakePart::GetLinearVelocity() const {
    const tVector attach = cKinTree::GetBodyAttachPt(bd(owner), id);
    return cKinTree::CalcWorldVel(jm(owner), owner->pose, owner->vel, id, attach);
}
tVector FakePart::LocalToWorldPos(const tVector& local_pos) const {
    tMatrix m = cKinTree::BodyWorldTrans(jm(owner), owner->pose, bd(owner), id);
    tVector p = local_pos;
    p[3] = 1;
    tVector w = m * p;
    w[3] = 0;
    return w;
}
void FakeJoint::CalcRotation(tVector& out_axis, double& out_theta) const {
    out_axis = axis_rel;
    out_theta = cKinTree::GetJointTheta(jm(owner), owner->pose, id);
}
// cSimObj::GetRotation of the child body: Bullet hands back an axis-angle extracted from the body's orientation, not an accumulated
// angle, so a link that has turned past half a revolution reads wrapped.  The accumulated angle (cKinTree::CalcJointWorldTheta)
// goes through the reference's own rotation-matrix -> axis-angle extraction here (angle in [0, pi], axis +-z).
void FakeJoint::GetChildRotation(tVector& out_axis, double& out_theta) const {
    tVector z;
    double acc = 0;
    cKinTree::CalcJointWorldTheta(jm(owner), owner->pose, id, z, acc);
    cMathUtil::RotMatToAxisAngle(cMathUtil::RotateMat(z, acc), out_axis, out_theta);
}
tVector FakeJoint::CalcJointVelRel() const {
    const int o = cKinTree::GetParamOffset(jm(owner), id);
    return axis_rel * owner->vel[o];
}

struct FakeGround : public cGround {
    height_fn fn = nullptr;
    void* user = nullptr;
    double SampleHeight(const tVector& pos) const override { return fn ? fn(pos[0], user) : 0.0; }
    double SampleHeight(const tVector& pos, bool& out_valid) const override { out_valid = true; return SampleHeight(pos); }
};

// ---- trivial value types of the Bullet-backed classes (members of cSimObj / cJoint, so their constructors run)
cContactManager::tContactHandle::tContactHandle() : mID(-1), mFlags(-1), mFilterFlags(-1) {}
cWorld::tConstraintHandle::tConstraintHandle() : mCons(nullptr) {}
cWorld::tJointParams::tJointParams() {}

// ---- the Bullet-backed base classes: constructors / destructors only (their vtables land here; see the header comment)
cSimObj::cSimObj() {}
cSimObj::~cSimObj() {}
bool cWorld::tConstraintHandle::IsValid() const { return mCons != nullptr; }
void cWorld::tConstraintHandle::Clear() { mCons = nullptr; }
cSimCharacter::cSimCharacter() {}
cSimCharacter::~cSimCharacter() {}
// called as base-class functions by cSimCharSoftFall (the real ones drive Bullet)
// cSimCharacter::Update / Reset / Init without their Bullet halves (sim/SimCharacter.cpp:25-107): run the controller; restore
// pose0 / vel0 and reset the controller; load skeleton + state file.  g_reset_loads_pose0 is off for the tests that install
// every state themselves.
static bool g_reset_loads_pose0 = false;
void cSimCharacter::Update(double time_step) { if (mController) mController->Update(time_step); }
void cSimCharacter::Reset() {
    if (g_reset_loads_pose0) cCharacter::Reset();
    if (mController) mController->Reset();
}
bool cSimCharacter::Init(std::shared_ptr<cWorld> world, const tParams& params) {
    mWorld = world;
    FakeChar* f = dynamic_cast<FakeChar*>(this);
    return f && f->load(params.mCharFile, params.mStateFile);
}
cSimCharacter::tParams::tParams() : mCharFile(""), mStateFile(""), mPos(0, 0, 0, 0), mPlaneCons(cWorld::ePlaneConsNone) {}
void cSimCharacter::SetController(std::shared_ptr<cCharController> ctrl) { mController = ctrl; }
void cSimCharacter::RemoveController() { mController.reset(); }
bool cSimCharacter::HasController() const { return mController != nullptr; }
const std::shared_ptr<cCharController>& cSimCharacter::GetController() { return mController; }
const std::shared_ptr<cCharController>& cSimCharacter::GetController() const { return mController; }
void cSimCharacter::RegisterContacts(int, int) {}

// ---- the world, as far as sim/GroundVar2D.cpp needs it: the length scale, and the position of a terrain segment's (Bullet) body,
// which the reference stores in single precision at world scale (cWorld::SetPos / GetPos, sim/World.cpp:288-309)
cWorld::tParams::tParams() : mNumSubsteps(1), mScale(1), mGravity(gGravity) {}
cContactManager::cContactManager(cWorld& world) : mWorld(world) {}
cContactManager::~cContactManager() {}
cPerturbManager::cPerturbManager() {}
cPerturbManager::~cPerturbManager() {}
cWorld::cWorld() : mContactManager(*this) {}
cWorld::~cWorld() {}
typedef void (*world_fn)(double h, void* user);
// Bullet keeps a body's position in single precision, so the reference's terrain segments sit at float origins (at world scale)
// and its height samples carry that rounding (~1e-7 m).  With this flag the stand-in world keeps the origin in double instead:
// the scenario pin runs once each way -- exact origin to show that everything else agrees to rounding, float origin (the
// reference as it is) to bound the effect of the single-precision origin.
static bool g_world_exact_origin = false;
extern "C" void ref_world_exact_origin(int on) { g_world_exact_origin = on != 0; }
struct FakeWorld : public cWorld {
    double scale = 4.0;
    world_fn cb = nullptr;          // cWorld::Update: the test advances the physics (the oracle's) and installs the new state
    void* user = nullptr;
    void Update(double time_elapsed) override { if (cb) cb(time_elapsed, user); }
    void Reset() override {}
    mutable std::map<const cSimObj*, btVector3> origin;
    mutable std::map<const cSimObj*, tVector> origin_exact;     // see g_world_exact_origin
    double GetScale() const override { return scale; }
    void SetPos(const tVector& pos, cSimObj* obj) const override {
        const btScalar s = static_cast<btScalar>(GetScale());
        origin[obj] = btVector3(s * static_cast<btScalar>(pos[0]), s * static_cast<btScalar>(pos[1]), s * static_cast<btScalar>(pos[2]));
        origin_exact[obj] = pos * GetScale();
    }
    // Bullet's body->getAabb for a height field, WITHOUT its collision margin and in double precision from the stored (float)
    // origin: the ideal box of the grid.  (What real Bullet returns -- float arithmetic plus the shape's collision margin -- is
    // the part of the reference's ground that cannot be reproduced without Bullet; see tests/test_ref_pinning_cpu.py.)
    void CalcAABB(const cSimObj* obj, tVector& out_min, tVector& out_max) const override {
        const auto* hf = dynamic_cast<const btHeightfieldTerrainShape*>(obj->GetCollisionShape().get());
        const tVector o = org(obj);
        const btVector3& sc = hf->getLocalScaling();
        const double hx = 0.5 * (hf->width_ - 1) * sc[0], hz = 0.5 * (hf->length_ - 1) * sc[2];
        const double ymid = 0.5 * ((double)hf->min_h_ + hf->max_h_), hy = 0.5 * ((double)hf->max_h_ - hf->min_h_);
        (void)ymid;
        out_min = tVector(o[0] - hx, o[1] - hy, o[2] - hz, 0) / GetScale();
        out_max = tVector(o[0] + hx, o[1] + hy, o[2] + hz, 0) / GetScale();
    }
    tVector org(const cSimObj* obj) const {
        if (g_world_exact_origin) return origin_exact[obj];
        const btVector3& o = origin[obj];
        return tVector(o[0], o[1], o[2], 0);
    }
    tVector GetPos(const cSimObj* obj) const override {
        const tVector o = org(obj);
        tVector p(o[0], o[1], o[2], 0);
        p /= GetScale();
        return p;
    }
};
void cSimObj::Init(std::shared_ptr<cWorld> world) { mWorld = world; }
tVector cSimObj::GetPos() const { return mWorld->GetPos(this); }                 // sim/SimObj.cpp:16-24
void cSimObj::SetPos(const tVector& pos) { mWorld->SetPos(pos, this); }
void cSimObj::CalcAABB(tVector& out_min, tVector& out_max) const { mWorld->CalcAABB(this, out_min, out_max); }   // sim/SimObj.cpp:202-205
const std::unique_ptr<btCollisionShape>& cSimObj::GetCollisionShape() const { return mShape; }
void cSimObj::UpdateContact(int, int) {}
void cSimObj::RemoveFromWorld() {}

struct FakeGroundVar : public cGroundVar2D {
    int num_verts(int s) const { return GetSegment(s)->GetGridWidth(); }      // first row of the (duplicated) height grid
    const float* verts(int s) const { return GetSegment(s)->mData.data(); }
    double min_x(int s) const { return GetSegment(s)->GetMinX(); }
    bool flipped() const { return mFlipSeg; }
};

// ---- cNeuralNet: what the controllers, scenarios and trainers ask of it
#include "ref_net_standin.h"

// ------------------------------------------------------------------------------------------------ C entry points
struct RefCtrl {
    std::unique_ptr<FakeChar> chp;
    std::shared_ptr<FakeGround> ground;
    std::shared_ptr<cTerrainRLCharController> ctrl;
};

extern "C" {

// kind 0: cDogControllerQ (what -char_ctrl dog builds: fixed gait / commanded actions), 1: cDogControllerMACE, 2: cGoatControllerMACE,
//      3: cRaptorControllerQ, 4: cRaptorControllerMACE
RefCtrl* ref_ctrl_create(const char* char_file, int kind, double gx, double gy, height_fn fn, void* user) {
    if (std::getenv("REF_CTRL_DEBUG")) signal(SIGSEGV, ref_segv_handler);
    RefCtrl* r = new RefCtrl();
    bool ok;
    if (kind <= 2) { auto* c = new FakeCharT<cSimDog>(); r->chp.reset(c); ok = c->Load(char_file); }       // the goat is a cSimDog too
    else { auto* c = new FakeCharT<cSimRaptor>(); r->chp.reset(c); ok = c->Load(char_file); }
    if (!ok) { delete r; return nullptr; }
    cSimCharacter* sim = r->chp->sim();
    r->ground = std::make_shared<FakeGround>();
    r->ground->fn = fn; r->ground->user = user;
    const tVector g(gx, gy, 0, 0);
    if (kind <= 2) {
        std::shared_ptr<cDogController> c;
        if (kind == 0) c = std::make_shared<cDogControllerQ>();
        else if (kind == 1) c = std::make_shared<cDogControllerMACE>();
        else c = std::make_shared<cGoatControllerMACE>();
        c->SetGround(r->ground);
        c->Init(sim, g, char_file);
        r->ctrl = c;
    } else {
        std::shared_ptr<cRaptorController> c;
        if (kind == 3) c = std::make_shared<cRaptorControllerQ>();
        else c = std::make_shared<cRaptorControllerMACE>();
        c->SetGround(r->ground);
        c->Init(sim, g, char_file);
        r->ctrl = c;
    }
    return r;
}
void ref_ctrl_destroy(RefCtrl* r) { delete r; }
int ref_ctrl_valid(RefCtrl* r) { return r->ctrl->IsValid() ? 1 : 0; }
int ref_ctrl_num_dof(RefCtrl* r) { return r->chp->sim()->GetNumDof(); }
int ref_ctrl_num_joints(RefCtrl* r) { return r->chp->sim()->GetNumJoints(); }
// installs the simulation state the controller and the fall logic will see (pose, velocity, per-part contact bits, COM)
void ref_ctrl_set_state(RefCtrl* r, const double* pose, const double* vel, const unsigned char* contact, const double* com,
                        const double* com_vel) {
    FakeChar& ch = *r->chp;
    const int nd = ch.sim()->GetNumDof(), nj = ch.sim()->GetNumJoints();
    for (int i = 0; i < nd; ++i) { ch.pose[i] = pose[i]; ch.vel[i] = vel[i]; }
    for (int j = 0; j < nj; ++j) static_cast<FakePart*>(ch.parts[j].get())->contact = contact[j] != 0;
    ch.com = tVector(com[0], com[1], 0, 0);
    ch.com_vel = tVector(com_vel[0], com_vel[1], 0, 0);
}
// cSimCharSoftFall::Reset / Update (fall-distance and fall-contact counters) and the reference's own verdicts
void ref_char_reset(RefCtrl* r) { r->chp->soft_fall_reset(); }
void ref_char_update(RefCtrl* r, double h) { r->chp->soft_fall_update(h); }
int ref_char_has_fallen(RefCtrl* r) { return r->chp->sim()->HasFallen() ? 1 : 0; }
int ref_char_has_stumbled(RefCtrl* r) { return r->chp->sim()->HasStumbled() ? 1 : 0; }
void ref_ctrl_reset(RefCtrl* r) { r->ctrl->Reset(); }
void ref_ctrl_update(RefCtrl* r, double h) { r->ctrl->Update(h); }
// what cJoint::ApplyTorque would hand to the physics: the accumulated joint torque after cJoint::ClampTotalTorque with the limit
// cPDController installed (sim/Joint.cpp:171-190,257-264, sim/PDController.cpp:99-100)
static void applied_tau(FakeChar& ch, double* out);
void ref_ctrl_get_applied_tau(RefCtrl* r, double* out) { applied_tau(*r->chp, out); }
}  // extern "C"
static void applied_tau(FakeChar& ch, double* out) {
    cSimCharacter* sim = ch.sim();
    for (int i = 0; i < sim->GetNumDof(); ++i) out[i] = 0;
    for (int j = 0; j < sim->GetNumJoints(); ++j) {
        if (!ch.joints[j].IsValid()) continue;
        tVector t = ch.joints[j].GetTorque();
        ch.joints[j].ClampTotalTorque(t);
        out[sim->GetParamOffset(j)] = t[2];
    }
}
extern "C" {
void ref_ctrl_get_tau(RefCtrl* r, double* out) { for (int i = 0; i < r->chp->sim()->GetNumDof(); ++i) out[i] = r->chp->last_tau[i]; }
// state, phase, action id, then the full parameter vector of the current action
int ref_ctrl_get_fsm(RefCtrl* r, double* out, int cap) {
    int k = 0;
    out[k++] = r->ctrl->GetState();
    out[k++] = r->ctrl->GetPhase();
    out[k++] = r->ctrl->GetCurrActionID();
    Eigen::VectorXd p;
    r->ctrl->BuildOptParams(p);
    for (int i = 0; i < p.size() && k < cap; ++i) out[k++] = p[i];
    return k;
}
int ref_ctrl_num_actions(RefCtrl* r) { return r->ctrl->GetNumActions(); }
// cBaseControllerMACE::BuildNNOutputOffsetScale of the compiled reference controller
int ref_ctrl_output_offset_scale(RefCtrl* r, double* off, double* scale, int cap) {
    Eigen::VectorXd o, s;
    r->ctrl->BuildNNOutputOffsetScale(o, s);
    for (int i = 0; i < o.size() && i < cap; ++i) { off[i] = o[i]; scale[i] = s[i]; }
    return o.size();
}
double ref_ctrl_calc_reward(RefCtrl* r) { return r->ctrl->CalcReward(); }     // c{Dog,Raptor}Controller::CalcReward of the last cycle
int ref_ctrl_poli_state(RefCtrl* r, double* out, int cap) {
    Eigen::VectorXd s;
    r->ctrl->RecordPoliState(s);
    for (int i = 0; i < s.size() && i < cap; ++i) out[i] = s[i];
    return s.size();
}
// cNNController::LoadNet on the stand-in net (sizes installed with ref_ctrl_set_net_output first): runs the reference's own size
// checks and cBaseControllerMACE::UpdateFragParams (number of actor-critic pairs, fragment size)
int ref_ctrl_load_net(RefCtrl* r) { return r->ctrl->LoadNet("stand-in") ? 1 : 0; }
// the blob source behind cNeuralNet::GetLayerState of the stand-in (cScenarioPoliEval::RecordNNActivation)
void ref_ctrl_set_layer_cb(layer_fn cb) { g_layer_cb = cb; }
void ref_ctrl_set_net_output(int n_in, const double* y, const double* out_scale, int n_out) {
    g_net_in = n_in; g_net_out = n_out;
    g_net_output.resize(n_out); g_out_scale.resize(n_out);
    for (int i = 0; i < n_out; ++i) { g_net_output[i] = y[i]; g_out_scale[i] = out_scale[i]; }
}


// ---------------------------------------------------------------------------------------- streaming ground (sim/GroundVar2D.cpp)
struct RefGround {
    std::shared_ptr<FakeWorld> world;
    FakeGroundVar ground;
};
// cScenarioSimChar::BuildGround / ResetGround: terrain function + parameters, seed, Init over the initial view window
RefGround* ref_ground_create(int type, const double* params40, unsigned long seed, double bmin, double bmax) {
    RefGround* g = new RefGround();
    g->world = std::make_shared<FakeWorld>();
    Eigen::VectorXd p(cTerrainGen2D::eParamsMax);
    for (int i = 0; i < cTerrainGen2D::eParamsMax; ++i) p[i] = params40[i];
    g->ground.SetTerrainFunc(cTerrainGen2D::GetTerrainFunc(static_cast<cTerrainGen2D::eType>(type)));
    g->ground.SetTerrainParams(p);
    g->ground.SeedRand(seed);
    cGroundVar2D::tParams gp;
    g->ground.Init(g->world, gp, tVector(bmin, 0, 0, 0), tVector(bmax, 0, 0, 0));
    return g;
}
void ref_ground_destroy(RefGround* g) { delete g; }
void ref_ground_update(RefGround* g, double bmin, double bmax) { g->ground.Update(tVector(bmin, 0, 0, 0), tVector(bmax, 0, 0, 0)); }
// logical segment s (0 = min, 1 = max side as the reference orders them): vertex data (at world scale, as stored), min x
int ref_ground_segment(RefGround* g, int s, float* out, int cap, double* min_x) {
    const int n = g->ground.num_verts(s);
    for (int i = 0; i < n && i < cap; ++i) out[i] = g->ground.verts(s)[i];
    *min_x = g->ground.min_x(s);
    return n;
}
int ref_ground_flipped(RefGround* g) { return g->ground.flipped() ? 1 : 0; }
double ref_ground_sample(RefGround* g, double x) { return g->ground.SampleHeight(tVector(x, 0, 0, 0)); }

}  // extern "C"

// ------------------------------------------------------------------------------------------------ whole scenarios
// cScenarioPoliEval / cScenarioExpMACE (+ cScenarioExp, cScenarioSimChar, cScenario) compiled as they are.  Two of their factory
// functions are overridden: BuildWorld makes the FakeWorld above (whose Update hands the env-step to the test, which advances the
// oracle's physics and installs the new state) and CreateCharacter makes the fake-backed cSimDog / cSimRaptor.  Everything else --
// the step loop, ground streaming on the reference's own cGroundVar2D, the controller, the cycle / tuple / episode bookkeeping,
// fall handling and Reset -- runs as compiled.  cScenarioExp::CommandRandAction draws from the reference's process-global random
// engine; the override commands the action the test names (the oracle's own draw) instead.
typedef int (*cmd_fn)(void* user);
extern cRand g_math_util_rand asm("_ZN9cMathUtil5gRandE");      // cMathUtil::gRand (a private static member of the reference class)
template <typename Base>
struct FakeScn : public Base {
    world_fn wcb = nullptr;
    cmd_fn ccb = nullptr;
    void* user = nullptr;
    void BuildWorld() override {
        auto w = std::make_shared<FakeWorld>();
        w->scale = this->mWorldScale;
        w->cb = wcb; w->user = user;
        this->mWorld = w;
    }
    void CreateCharacter(std::shared_ptr<cSimCharacter>& out_char) const override {
        if (this->mCharType == cScenarioSimChar::eCharDog) out_char = std::shared_ptr<cSimCharacter>(new FakeCharT<cSimDog>());
        else out_char = std::shared_ptr<cSimCharacter>(new FakeCharT<cSimRaptor>());
    }
    FakeChar* fake() { return dynamic_cast<FakeChar*>(this->mChar.get()); }
    cGroundVar2D* ground() { return static_cast<cGroundVar2D*>(this->mGround.get()); }
    cTerrainRLCharController* ctrl() { return dynamic_cast<cTerrainRLCharController*>(this->mChar->GetController().get()); }
    double time() const { return this->mTime; }
};
struct FakeScnEval : public FakeScn<cScenarioPoliEval> {};
struct FakeScnExp : public FakeScn<cScenarioExpMACE> {
    // cScenarioTrain::BuildScenePool calls Init, sets the initial exploration rates, then Reset ("rebuild ground"): the seeds go in
    // at that Reset, as ref_scn_create does for a stand-alone scenario
    bool seed_pending = false;
    unsigned long ground_seed = 0, rand_seed = 0;
    void Reset() override {
        if (seed_pending) {
            seed_pending = false;
            ground()->SeedRand(ground_seed);
            g_math_util_rand = cRand();
            cMathUtil::SeedRand(rand_seed);
        }
        cScenarioExpMACE::Reset();
    }
    void CommandRandAction() override {
        if (ccb) this->mChar->GetController()->CommandAction(ccb(user));
        else cScenarioExpMACE::CommandRandAction();          // draws from cMathUtil's engine (ref_scn_create seeds it)
    }
    int tuple_count() const { return mTupleCount; }
    int cycle_count() const { return mCycleCount; }
};
struct RefScn {
    int mode = 0;
    std::unique_ptr<FakeScnEval> ev;
    std::unique_ptr<FakeScnExp> ex;
    cScenarioSimChar* scn() { return mode == 0 ? static_cast<cScenarioSimChar*>(ev.get()) : static_cast<cScenarioSimChar*>(ex.get()); }
    FakeChar* fake() { return mode == 0 ? ev->fake() : ex->fake(); }
    cGroundVar2D* ground() { return mode == 0 ? ev->ground() : ex->ground(); }
    cTerrainRLCharController* ctrl() { return mode == 0 ? ev->ctrl() : ex->ctrl(); }
};

extern "C" {

// mode 0: cScenarioPoliEval, 1: cScenarioExpMACE; arg_file as the reference's Main reads it (-arg_file=); extra: further "-key=
// value" tokens (tuple buffer size, exploration rates); rand_seed != 0: seed cMathUtil's process-global engine after Init.  Seeds the ground and rebuilds it with Reset, as
// cOptScenarioPoliEval::BuildScenePool does (optimizer/scenarios/OptScenarioPoliEval.cpp:150-160).
RefScn* ref_scn_create(const char* arg_file, int mode, char** extra, int n_extra, unsigned long seed, world_fn wcb, net_fn ncb, cmd_fn ccb,
                       void* user, unsigned long rand_seed) {
    if (std::getenv("REF_CTRL_DEBUG")) signal(SIGSEGV, ref_segv_handler);
    g_net_cb = ncb; g_net_user = user;
    g_reset_loads_pose0 = true;
    cArgParser parser;
    if (n_extra > 0) parser.AppendArgs(extra, n_extra);       // cArgParser returns the FIRST occurrence of a key: overrides go first
    parser.AppendArgs(std::string(arg_file));
    RefScn* r = new RefScn();
    r->mode = mode;
    if (mode == 0) { r->ev.reset(new FakeScnEval()); r->ev->wcb = wcb; r->ev->ccb = ccb; r->ev->user = user; }
    else { r->ex.reset(new FakeScnExp()); r->ex->wcb = wcb; r->ex->ccb = ccb; r->ex->user = user; }
    cScenarioSimChar* s = r->scn();
    s->ParseArgs(parser);
    s->Init();
    r->ground()->SeedRand(seed);
    if (rand_seed) {                                     // the engine behind every cMathUtil::Rand* call (exploration, random commands)
        g_math_util_rand = cRand();                      // a fresh object: cRand::Seed re-seeds the engine but leaves the normal distribution's cached value
        cMathUtil::SeedRand(rand_seed);
    }
    s->Reset();
    return r;
}
void ref_scn_destroy(RefScn* r) { g_net_cb = nullptr; g_reset_loads_pose0 = false; delete r; }
void ref_scn_update(RefScn* r, double dt) { r->scn()->Update(dt); }
// what the test's world callback installs after advancing the physics: pose, velocity, contact bits
void ref_scn_set_state(RefScn* r, const double* pose, const double* vel, const unsigned char* contact) {
    FakeChar& ch = *r->fake();
    const int nd = ch.sim()->GetNumDof(), nj = ch.sim()->GetNumJoints();
    for (int i = 0; i < nd; ++i) { ch.pose[i] = pose[i]; ch.vel[i] = vel[i]; }
    for (int j = 0; j < nj; ++j) static_cast<FakePart*>(ch.parts[j].get())->contact = contact[j] != 0;
}
void ref_scn_get_state(RefScn* r, double* pose, double* vel, double* tau) {
    FakeChar& ch = *r->fake();
    const int nd = ch.sim()->GetNumDof();
    for (int i = 0; i < nd; ++i) { pose[i] = ch.pose[i]; vel[i] = ch.vel[i]; tau[i] = ch.last_tau[i]; }
}
void ref_scn_get_applied_tau(RefScn* r, double* out) { applied_tau(*r->fake(), out); }      // as ref_ctrl_get_applied_tau
int ref_scn_num_dof(RefScn* r) { return r->fake()->sim()->GetNumDof(); }
int ref_scn_get_fsm(RefScn* r, double* out) {
    out[0] = r->ctrl()->GetState(); out[1] = r->ctrl()->GetPhase(); out[2] = r->ctrl()->GetCurrActionID();
    return 3;
}
int ref_scn_has_fallen(RefScn* r) { return r->scn()->HasFallen() ? 1 : 0; }
double ref_scn_time(RefScn* r) { return r->mode == 0 ? r->ev->time() : r->ex->time(); }
double ref_scn_sample_height(RefScn* r, double x) { return r->ground()->SampleHeight(tVector(x, 0, 0, 0)); }
// cScenarioPoliEval: cycles, episodes, average distance, distance log
void ref_scn_eval_stats(RefScn* r, long* cycles, long* episodes, double* avg_dist) {
    *cycles = r->ev->GetNumCycles(); *episodes = r->ev->GetNumEpisodes(); *avg_dist = r->ev->GetAvgDist();
}
int ref_scn_dist_log(RefScn* r, double* out, int cap) {
    const auto& log = r->ev->GetDistLog();
    for (int i = 0; i < (int)log.size() && i < cap; ++i) out[i] = log[i];
    return (int)log.size();
}
// cScenarioExp: tuples recorded so far (total count; the buffer keeps the last `tuple_buffer_size`), cycle count of this episode
void ref_scn_exp_counts(RefScn* r, long* tuples, long* cycles) { *tuples = r->ex->tuple_count(); *cycles = r->ex->cycle_count(); }
// tuple `idx` of the ring buffer: reward, flags, state_beg | action | state_end
int ref_scn_get_tuple(RefScn* r, int idx, double* reward, unsigned* flags, double* row, int cap) {
    const tExpTuple& t = r->ex->GetTuples()[idx];
    *reward = t.mReward; *flags = t.mFlags;
    int k = 0;
    for (int i = 0; i < (int)t.mStateBeg.size() && k < cap; ++i) row[k++] = t.mStateBeg[i];
    for (int i = 0; i < (int)t.mAction.size() && k < cap; ++i) row[k++] = t.mAction[i];
    for (int i = 0; i < (int)t.mStateEnd.size() && k < cap; ++i) row[k++] = t.mStateEnd[i];
    return k;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------ the training scenario
// cScenarioTrain + cScenarioTrainMACE compiled as they are, with the compiled cMACETrainer / cNeuralNetTrainer / cNeuralNetLearner
// behind them and one exploration scenario (the fake-backed cScenarioExpMACE above) in the pool: tuple hand-over when the buffer is
// full, cNeuralNetLearner::Train, the annealed exploration rates and temperature, the curriculum phase -> SetTerrainParamsLerp.
// Only BuildExpScene is overridden (to make the exploration scenario the fake-backed one).
struct FakeScnTrain : public cScenarioTrainMACE {
    world_fn wcb = nullptr;
    void* user = nullptr;
    unsigned long ground_seed = 0, rand_seed = 0;
    void BuildExpScene(std::shared_ptr<cScenarioExp>& out_exp) const override {
        auto e = std::make_shared<FakeScnExp>();
        e->wcb = wcb; e->user = user; e->ccb = nullptr;
        e->seed_pending = true; e->ground_seed = ground_seed; e->rand_seed = rand_seed;
        out_exp = e;
    }
    FakeScnExp* exp0() { return static_cast<FakeScnExp*>(mExpPool[0].get()); }
    void sched(int iters, double* out) const {
        out[0] = CalcExpRate(iters); out[1] = CalcExpTemp(iters); out[2] = CalcExpBaseRate(iters); out[3] = CalcCurriculumPhase(iters);
    }
    long trainer_tuples() const { return std::static_pointer_cast<cNeuralNetTrainer>(mTrainer)->GetNumTuples(); }
};
struct RefTrainScn {
    std::unique_ptr<FakeScnTrain> scn;
};

extern "C" {

// net sizes / batch and the network callbacks, then the scenario from the reference's training arg file (+ extra tokens)
RefTrainScn* ref_strain_create(const char* arg_file, char** extra, int n_extra, unsigned long ground_seed, unsigned long rand_seed,
                               const int* net_dims, world_fn wcb, eval_fn ev, train_fn tr, copy_fn cp, calc_os_fn cos, set_os_fn sos,
                               void* user) {
    if (std::getenv("REF_CTRL_DEBUG")) signal(SIGSEGV, ref_segv_handler);
    g_hooks.n_in = net_dims[0]; g_hooks.n_out = net_dims[1]; g_hooks.batch = net_dims[2];
    g_hooks.eval = ev; g_hooks.train = tr; g_hooks.copy = cp; g_hooks.calc_os = cos; g_hooks.set_os = sos; g_hooks.user = user;
    g_next_net = 0;
    g_net_id.clear();
    g_net_cb = nullptr;
    g_reset_loads_pose0 = true;
    cArgParser parser;
    if (n_extra > 0) parser.AppendArgs(extra, n_extra);       // cArgParser returns the FIRST occurrence of a key: overrides go first
    parser.AppendArgs(std::string(arg_file));
    RefTrainScn* r = new RefTrainScn();
    r->scn.reset(new FakeScnTrain());
    r->scn->wcb = wcb; r->scn->user = user; r->scn->ground_seed = ground_seed; r->scn->rand_seed = rand_seed;
    r->scn->ParseArgs(parser);
    r->scn->SetExpPoolSize(1);
    r->scn->Init();
    return r;
}
// the next value of cMathUtil's engine without advancing it (a copy draws): lets the test locate a divergence of the draw sequences
int ref_rand_peek() { cRand c = g_math_util_rand; return c.RandInt(); }
void ref_strain_destroy(RefTrainScn* r) {
    delete r;
    g_hooks = NetHooks();
    g_reset_loads_pose0 = false;
}
void ref_strain_update(RefTrainScn* r, double dt) { r->scn->Update(dt); }        // cScenarioTrain::Update -> UpdateExpScene
void ref_strain_set_state(RefTrainScn* r, const double* pose, const double* vel, const unsigned char* contact) {
    FakeChar& ch = *r->scn->exp0()->fake();
    const int nd = ch.sim()->GetNumDof(), nj = ch.sim()->GetNumJoints();
    for (int i = 0; i < nd; ++i) { ch.pose[i] = pose[i]; ch.vel[i] = vel[i]; }
    for (int j = 0; j < nj; ++j) static_cast<FakePart*>(ch.parts[j].get())->contact = contact[j] != 0;
}
void ref_strain_get_state(RefTrainScn* r, double* pose, double* vel, double* tau) {
    FakeChar& ch = *r->scn->exp0()->fake();
    const int nd = ch.sim()->GetNumDof();
    for (int i = 0; i < nd; ++i) { pose[i] = ch.pose[i]; vel[i] = ch.vel[i]; tau[i] = ch.last_tau[i]; }
}
// state, phase, action id, then the current action's parameter vector; returns the number of values
int ref_strain_get_fsm(RefTrainScn* r, double* out, int cap) {
    auto* c = r->scn->exp0()->ctrl();
    int k = 0;
    out[k++] = c->GetState(); out[k++] = c->GetPhase(); out[k++] = c->GetCurrActionID();
    Eigen::VectorXd p;
    c->BuildOptParams(p);
    for (int i = 0; i < (int)p.size() && k < cap; ++i) out[k++] = p[i];
    return k;
}
// iter, tuples seen by the trainer, tuples in the scenario's buffer, then the scenario's current exploration rate / temperature /
// base-action rate
void ref_strain_status(RefTrainScn* r, long* counts, double* rates) {
    counts[0] = r->scn->GetIter(); counts[1] = r->scn->trainer_tuples(); counts[2] = r->scn->exp0()->tuple_count();
    rates[0] = r->scn->exp0()->GetExpRate(); rates[1] = r->scn->exp0()->GetExpTemp(); rates[2] = r->scn->exp0()->GetExpBaseActionRate();
}
double ref_strain_sample_height(RefTrainScn* r, double x) { return r->scn->exp0()->ground()->SampleHeight(tVector(x, 0, 0, 0)); }
// cScenarioTrain::CalcExpRate / CalcExpTemp / CalcExpBaseRate / CalcCurriculumPhase at `iters`
void ref_strain_schedule(RefTrainScn* r, int iters, double* out4) { r->scn->sched(iters, out4); }

}  // extern "C"

// ------------------------------------------------------------------------------------------------ the drop-in seam, compiled
// include/terrainrl_b200_adapter.h (the binding INTEGRATION.md describes) compiled against the reference's headers and put under
// the reference's own compiled cScenarioTrainMACE: BuildExpScene returns the batched adapter, everything else -- BuildScenePool,
// SetupLearner, UpdateExpScene, the annealing schedule, cNeuralNetLearner / cMACETrainer -- runs as compiled and drives the C ABI.
// The trl_* entry points are weak references: the test loads a library that exports them (RTLD_GLOBAL) before this one; every
// other test of this library never reaches them.
#pragma weak trl_create_from_pack
#pragma weak trl_destroy
#pragma weak trl_reset
#pragma weak trl_update
#pragma weak trl_sizes
#pragma weak trl_num_tuples
#pragma weak trl_get_tuples_f64
#pragma weak trl_reset_tuples
#pragma weak trl_set_explore
#pragma weak trl_set_terrain_lerp
#pragma weak trl_set_weights
#pragma weak trl_last_error
#pragma weak trl_seed_terrain
#pragma weak trl_eval_stats
#pragma weak trl_dist_log
#pragma weak trl_reset_avg_dist
#define TRL_ADAPTER_WITH_REFERENCE_SCENARIO 1
#include "../include/terrainrl_b200_adapter.h"
// the deployment classes (Base = the reference's own cScenarioExpMACE / cScenarioPoliEval), every member instantiated: compile check
template class cScenarioExpBatchedT<cScenarioExpMACE>;
template class cScenarioPoliEvalBatchedT<cScenarioPoliEval>;

typedef cScenarioExpBatchedT<FakeScnExp> BatchedExp;
struct BatchedScnTrain : public cScenarioTrainMACE {
    std::string pack;
    int num_envs = 1;
    unsigned long long rng_seed = 1234;
    void BuildExpScene(std::shared_ptr<cScenarioExp>& out_exp) const override {
        auto e = std::make_shared<BatchedExp>();
        e->SetBatch(pack, num_envs, 0, rng_seed, nullptr);
        out_exp = e;
    }
    BatchedExp* exp0() { return static_cast<BatchedExp*>(mExpPool[0].get()); }
    long trainer_tuples() const { return std::static_pointer_cast<cNeuralNetTrainer>(mTrainer)->GetNumTuples(); }
};

extern "C" {
// as ref_strain_create, with the exploration scene replaced by the adapter over a batch of `num_envs` environments of `pack`
BatchedScnTrain* ref_btrain_create(const char* arg_file, char** extra, int n_extra, const char* pack, int num_envs, unsigned long long rng_seed,
                                   unsigned long rand_seed, const int* net_dims, eval_fn ev, train_fn tr, copy_fn cp, calc_os_fn cos,
                                   set_os_fn sos, void* user) {
    if (std::getenv("REF_CTRL_DEBUG")) signal(SIGSEGV, ref_segv_handler);
    if (!trl_create_from_pack) { std::fprintf(stderr, "ref_btrain_create: no library exporting the C ABI is loaded\n"); return nullptr; }
    g_hooks.n_in = net_dims[0]; g_hooks.n_out = net_dims[1]; g_hooks.batch = net_dims[2];
    g_hooks.eval = ev; g_hooks.train = tr; g_hooks.copy = cp; g_hooks.calc_os = cos; g_hooks.set_os = sos; g_hooks.user = user;
    g_next_net = 0;
    g_net_id.clear();
    g_net_cb = nullptr;
    g_reset_loads_pose0 = true;
    cArgParser parser;
    if (n_extra > 0) parser.AppendArgs(extra, n_extra);
    parser.AppendArgs(std::string(arg_file));
    auto* r = new BatchedScnTrain();
    r->pack = pack; r->num_envs = num_envs; r->rng_seed = rng_seed;
    g_math_util_rand = cRand();                  // the trainer's minibatch sampling draws from cMathUtil's engine
    cMathUtil::SeedRand(rand_seed);
    r->ParseArgs(parser);
    r->SetExpPoolSize(1);
    r->Init();
    return r;
}
void ref_btrain_destroy(BatchedScnTrain* r) {
    delete r;
    g_hooks = NetHooks();
    g_reset_loads_pose0 = false;
}
void ref_btrain_reseed(unsigned long rand_seed) { g_math_util_rand = cRand(); cMathUtil::SeedRand(rand_seed); }
void ref_btrain_update(BatchedScnTrain* r, double dt) { r->Update(dt); }          // cScenarioTrain::Update -> UpdateExpScene
void* ref_btrain_handle(BatchedScnTrain* r) { return r->exp0()->GetHandle(); }    // the trl_handle the adapter owns
// what a deployment does after cNeuralNetLearner::SyncNet with the blobs of the (real) cNeuralNet: cScenarioExpBatched::PushWeights
void ref_btrain_push_weights(BatchedScnTrain* r, const double* const* blobs, const int64_t* counts, int nblobs, const double* in_off,
                             const double* in_scale, const double* out_off, const double* out_scale) {
    r->exp0()->PushWeights(blobs, counts, nblobs, in_off, in_scale, out_off, out_scale);
}
// iter, tuples seen by the trainer; the exploration rate / temperature / base-action rate the compiled scenario holds
void ref_btrain_status(BatchedScnTrain* r, long* counts, double* rates) {
    counts[0] = r->GetIter(); counts[1] = r->trainer_tuples();
    rates[0] = r->exp0()->GetExpRate(); rates[1] = r->exp0()->GetExpTemp(); rates[2] = r->exp0()->GetExpBaseActionRate();
}
}  // extern "C"

// ---- policy evaluation through the adapter: the calls cOptScenarioPoliEval makes on a pooled scene, on a cScenarioPoliEval pointer
typedef cScenarioPoliEvalBatchedT<FakeScnEval> BatchedEval;
struct RefBEval {
    std::unique_ptr<BatchedEval> scn;
    cScenarioPoliEval* base() { return scn.get(); }       // every call below goes through the reference's own virtual interface
};
extern "C" {
RefBEval* ref_beval_create(const char* arg_file, const char* pack, int num_envs, unsigned long long rng_seed, unsigned long seed) {
    if (!trl_create_from_pack) { std::fprintf(stderr, "ref_beval_create: no library exporting the C ABI is loaded\n"); return nullptr; }
    g_net_cb = nullptr;
    g_reset_loads_pose0 = true;
    cArgParser parser;
    parser.AppendArgs(std::string(arg_file));
    auto* r = new RefBEval();
    r->scn.reset(new BatchedEval());
    r->scn->SetBatch(pack, num_envs, 0, rng_seed);
    cScenarioPoliEval* s = r->base();
    s->ParseArgs(parser);                                  // cOptScenarioPoliEval::BuildScenePool (OptScenarioPoliEval.cpp:135-163)
    s->Init();
    s->SetRandSeed(seed);
    s->Reset();
    return r;
}
void ref_beval_destroy(RefBEval* r) { g_reset_loads_pose0 = false; delete r; }
void* ref_beval_handle(RefBEval* r) { return r->scn->GetHandle(); }
void ref_beval_update(RefBEval* r, double dt) { r->base()->Update(dt); }
void ref_beval_stats(RefBEval* r, long* cycles, long* episodes, double* avg_dist) {
    *cycles = r->base()->GetNumCycles(); *episodes = r->base()->GetNumEpisodes(); *avg_dist = r->base()->GetAvgDist();
}
void ref_beval_reset_avg_dist(RefBEval* r) { r->base()->ResetAvgDist(); }
int ref_beval_dist_log(RefBEval* r, double* out, int cap) {
    const std::vector<double>& log = r->base()->GetDistLog();
    for (int i = 0; i < (int)log.size() && i < cap; ++i) out[i] = log[i];
    return (int)log.size();
}
}  // extern "C"

// ---- the reference's own evaluation driver (optimizer/scenarios/OptScenarioPoliEval.cpp, compiled as it is): Run -> one thread per
// pooled scene -> EvalHelper (Update until the episode / cycle budget is spent, UpdateRecord + ResetAvgDist every 10 episodes) ->
// OutputResults.  BuildScenePool names the scene class in a `new` expression, so the one edit a maintainer makes is restated
// here as an override: the reference's own lines (OptScenarioPoliEval.cpp:135-163) with `new cScenarioPoliEval()` replaced by the
// batched adapter.
#include "optimizer/scenarios/OptScenarioPoliEval.h"
struct BatchedOptEval : public cOptScenarioPoliEval {
    std::string pack;
    int num_envs = 1;
    unsigned long long rng_seed = 1234;
    unsigned long first_seed = 0;
    void BuildScenePool() override {
        mEvalPool.resize(mPoolSize);
        cRand rand;
        bool valid_seed = mRandSeed != 0;
        if (valid_seed) rand.Seed(mRandSeed);
        unsigned long curr_seed = static_cast<unsigned long>(std::abs(rand.RandInt()));
        first_seed = curr_seed;
        for (int i = 0; i < GetPoolSize(); ++i) {
            auto e = std::make_shared<BatchedEval>();
            e->SetBatch(pack, num_envs, 0, rng_seed);
            mEvalPool[i] = e;
            e->ParseArgs(mArgParser);
            e->Init();
            if (valid_seed) {
                e->SetRandSeed(curr_seed);
                e->Reset();
                curr_seed = static_cast<unsigned long>(std::abs(rand.RandInt()));
            }
        }
    }
    void results(long* counts, double* avg) const { counts[0] = mEpisodeCount; counts[1] = mCycleCount; *avg = mAvgDist; }
    BatchedEval* eval0() { return static_cast<BatchedEval*>(mEvalPool[0].get()); }
};
extern "C" {
BatchedOptEval* ref_opteval_create(const char* arg_file, char** extra, int n_extra, const char* pack, int num_envs, unsigned long long rng_seed) {
    if (!trl_create_from_pack) { std::fprintf(stderr, "ref_opteval_create: no library exporting the C ABI is loaded\n"); return nullptr; }
    g_net_cb = nullptr;
    g_reset_loads_pose0 = true;
    cArgParser parser;
    if (n_extra > 0) parser.AppendArgs(extra, n_extra);
    parser.AppendArgs(std::string(arg_file));
    auto* r = new BatchedOptEval();
    r->pack = pack; r->num_envs = num_envs; r->rng_seed = rng_seed;
    r->ParseArgs(parser);
    r->SetPoolSize(1);
    r->Init();
    return r;
}
void ref_opteval_run(BatchedOptEval* r) { r->Run(); }
void ref_opteval_results(BatchedOptEval* r, long* counts, double* avg, unsigned long* first_seed) { r->results(counts, avg); *first_seed = r->first_seed; }
void* ref_opteval_handle(BatchedOptEval* r) { return r->eval0()->GetHandle(); }
void ref_opteval_destroy(BatchedOptEval* r) { g_reset_loads_pose0 = false; delete r; }
}  // extern "C"
