// TEST INFRASTRUCTURE ONLY -- Bullet is not in this image.  Empty stand-ins for the Bullet types that the reference's *headers*
// mention, so that the controller sources (which never call Bullet themselves) can be compiled into oracle/_ref against the
// character back end of oracle/ref_fake_sim.cpp.  No Bullet behaviour is restated here.
#pragma once

typedef float btScalar;
struct btVector3 {
    btScalar m[4];
    btVector3() : m{0, 0, 0, 0} {}
    btVector3(btScalar x, btScalar y, btScalar z) : m{x, y, z, 0} {}
    btScalar operator[](int i) const { return m[i]; }
    btScalar& operator[](int i) { return m[i]; }
};
struct btQuaternion { btScalar m[4]; btQuaternion() : m{0, 0, 0, 1} {} };
struct btTransform { btQuaternion q; btVector3 o; };
class btMotionState { public: virtual ~btMotionState() {} virtual void getWorldTransform(btTransform&) const {} virtual void setWorldTransform(const btTransform&) {} };
class btDefaultMotionState : public btMotionState { public: btDefaultMotionState() {} };
class btCollisionShape {
public:
    virtual ~btCollisionShape() {}
    virtual const btVector3& getLocalScaling() const { static btVector3 one(1, 1, 1); return one; }
};
class btBoxShape : public btCollisionShape {};
class btCapsuleShape : public btCollisionShape {};
class btStaticPlaneShape : public btCollisionShape {};
class btCollisionObject { public: virtual ~btCollisionObject() {} };
class btRigidBody : public btCollisionObject {
public:
    struct btRigidBodyConstructionInfo {
        btRigidBodyConstructionInfo(btScalar, btMotionState*, btCollisionShape*, const btVector3&) {}
    };
    btRigidBody() {}
    explicit btRigidBody(const btRigidBodyConstructionInfo&) {}
    void setFriction(btScalar) {}
};
class btTypedConstraint { public: virtual ~btTypedConstraint() {} };
class btHingeConstraint : public btTypedConstraint {};
class btManifoldPoint {};
class btBroadphaseInterface { public: virtual ~btBroadphaseInterface() {} };
class btDefaultCollisionConfiguration { public: virtual ~btDefaultCollisionConfiguration() {} };
class btCollisionDispatcher { public: virtual ~btCollisionDispatcher() {} };
class btConstraintSolver { public: virtual ~btConstraintSolver() {} };
class btDiscreteDynamicsWorld { public: virtual ~btDiscreteDynamicsWorld() {} };

// ---- what sim/GroundVar2D.cpp constructs for a terrain segment (values are stored, nothing is simulated)
enum PHY_ScalarType { PHY_FLOAT, PHY_DOUBLE, PHY_INTEGER, PHY_SHORT, PHY_FIXEDPOINT88, PHY_UCHAR };
inline btScalar btVecGet(const btVector3& v, int i) { return v.m[i]; }
class btHeightfieldTerrainShape : public btCollisionShape {
public:
    btHeightfieldTerrainShape(int width, int length, const void*, btScalar, btScalar min_h, btScalar max_h, int, PHY_ScalarType, bool)
        : width_(width), length_(length), min_h_(min_h), max_h_(max_h) {}
    int width_, length_;
    btScalar min_h_, max_h_;
    void setLocalScaling(const btVector3& s) { scale_ = s; }
    const btVector3& getLocalScaling() const { return scale_; }
private:
    btVector3 scale_;
};
