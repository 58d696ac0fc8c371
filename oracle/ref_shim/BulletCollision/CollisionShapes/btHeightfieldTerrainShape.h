#pragma once
#include "btBulletDynamicsCommon.h"
