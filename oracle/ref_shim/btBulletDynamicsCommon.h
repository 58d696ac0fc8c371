// TEST INFRASTRUCTURE ONLY -- Bullet is not in this image.  Empty stand-ins for the Bullet types that the reference's *headers*
// mention, so that the controller sources (which never call Bullet themselves) can be compiled into oracle/_ref against the
// character back end of oracle/ref_fake_sim.cpp.  No Bullet behaviour is restated here.
#pragma once

typedef float btScalar;
struct btVector3 { btScalar m[4]; btVector3() : m{0, 0, 0, 0} {} btVector3(btScalar x, btScalar y, btScalar z) : m{x, y, z, 0} {} };
struct btQuaternion { btScalar m[4]; btQuaternion() : m{0, 0, 0, 1} {} };
struct btTransform { btQuaternion q; btVector3 o; };
class btMotionState { public: virtual ~btMotionState() {} virtual void getWorldTransform(btTransform&) const {} virtual void setWorldTransform(const btTransform&) {} };
class btDefaultMotionState : public btMotionState { public: btDefaultMotionState() {} };
class btCollisionShape { public: virtual ~btCollisionShape() {} };
class btBoxShape : public btCollisionShape {};
class btCapsuleShape : public btCollisionShape {};
class btStaticPlaneShape : public btCollisionShape {};
class btCollisionObject { public: virtual ~btCollisionObject() {} };
class btRigidBody : public btCollisionObject {};
class btTypedConstraint { public: virtual ~btTypedConstraint() {} };
class btHingeConstraint : public btTypedConstraint {};
class btManifoldPoint {};
class btBroadphaseInterface { public: virtual ~btBroadphaseInterface() {} };
class btDefaultCollisionConfiguration { public: virtual ~btDefaultCollisionConfiguration() {} };
class btCollisionDispatcher { public: virtual ~btCollisionDispatcher() {} };
class btConstraintSolver { public: virtual ~btConstraintSolver() {} };
class btDiscreteDynamicsWorld { public: virtual ~btDiscreteDynamicsWorld() {} };
