// tests only: the slice of cooperative_groups the decision kernel uses (thread-block clusters, distributed shared memory)
#pragma once
#include "../simt_runtime.h"
namespace cooperative_groups {
struct cluster_group {
    unsigned block_rank() const { return (unsigned)simt::cur->block->rank_in_cluster; }
    unsigned num_blocks() const { return (unsigned)simt::cur->cluster->nblocks; }
    void sync() const { simt::cluster_barrier(); }
    // address of the same dynamic-shared-memory object in CTA `rank` of the cluster
    template <typename T>
    T* map_shared_rank(T* p, int rank) const {
        simt::Thread* t = simt::cur;
        return (T*)((char*)p - t->block->dyn_smem + t->cluster->blocks[rank].dyn_smem);
    }
};
inline cluster_group this_cluster() { return cluster_group(); }
}  // namespace cooperative_groups
