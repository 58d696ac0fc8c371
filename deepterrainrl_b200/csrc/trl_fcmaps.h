// deepterrainrl_b200 -- what the host hands the batched decision kernel (trl_decide2.cuh) besides the buffers: the two TMA
// descriptors (CUDA build) and the plain pointers they describe (the test-only emulator build reads through those)
#pragma once
#ifndef TRL_SIMT_EMU
#include <cuda.h>      // CUtensorMap (type only; the encoder is resolved at run time, nothing links against libcuda)
#define TRL_GRID_CONSTANT __grid_constant__
#else
#define TRL_GRID_CONSTANT
#endif

namespace trl {
struct FcMaps {
#ifndef TRL_SIMT_EMU
    CUtensorMap w;      // terr_ip0 weights  [64 rows][5984] f64, box [64][16], 128-byte swizzle
    CUtensorMap a;      // conv2 outputs     [rows][5984]   f64, box [32][16], 128-byte swizzle
#endif
    const double* w_ptr;
    const double* a_ptr;
    int a_rows;
    unsigned long long* prof;    // measurement only (trl_debug_fc_phases): rank 0 / thread 0 of cluster 0 stamps %globaltimer at phase ends
};
}  // namespace trl
