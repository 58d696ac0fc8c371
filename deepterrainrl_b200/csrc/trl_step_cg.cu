// deepterrainrl_b200 -- L1-bypassing build of the env-step kernel (compile with -Xptxas -dlcm=cg); see trl_step.cu.
#define TRL_CG_VARIANT 1
#include "trl_step.cu"
