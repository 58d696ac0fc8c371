// deepterrainrl_b200 -- the batch handle behind the C ABI (private to csrc/: trl_host.cu, trl_train.cu)
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <map>
#include <string>
#include <vector>

#include "scene_pack.h"
#include "trl_types.h"
#include "trl_fcmaps.h"

struct trl_trainer;
struct trl_comm;

struct trl_handle {
    trl_trainer* trainer = nullptr;      // attached MACE trainer (trl_train.cu), owns the policy weights when present
    trl_comm* comm = nullptr;            // multi-GPU exchange (trl_comm.cu), present after trl_comm_init*
    int64_t tuples_dropped = 0;          // tuples the full tuple block refused, as observed by the host-side readers
    int device = 0, n = 0, mode = 0;
    trl::ScenePack scene;
    trl::ModelConst mc;
    trl::ExpSettings ex;                      // host copy; the kernels read the device copy d_ex
    trl::ExpSettings* d_ex = nullptr;
    trl::Buffers B;
    trl::NetWeights W{};
    std::vector<double*> net_blobs;      // 26 + 4 device arrays
    std::vector<int64_t> net_counts;
    int* done_count = nullptr;
    cudaStream_t stream = nullptr;
    cudaStream_t aux_stream = nullptr;           // = side[0]
    cudaStream_t side[8] = {nullptr};            // high-priority side streams: decisions + catch-up launches of env-step l run on side[l % lag]
    int groups = 2;                              // env groups of the main step launches (TRL_GROUPS, 1..kMaxGroups; reduced when a group would be empty): see enqueue_update
    cudaStream_t group_stream[7] = {nullptr};    // streams of groups 1 .. groups-1 (group 0 runs on `stream`)
    int lag = 6;                                 // overlap depth: env-steps a pending env may trail the main launches (TRL_LAG, 1..7)
    std::vector<cudaEvent_t> fork_events;        // dependencies between the two streams inside one update
    bool overlap = true;                         // TRL_SERIAL_SCHEDULE=1 turns the overlapped schedule off
    int decide_grid = 288;   // CTAs of the decision launch: a multiple of its cluster size (create_common sizes it from the SM count)
    int num_update_steps = 20;
    // batched decision path (trl_decide2.cuh): conv2 outputs of the pending decisions, the TMA descriptors over them and over the
    // terr_ip0 weights (re-encoded whenever the weight pointer changes), number of FC-stage clusters
    bool decide_v2 = true;                       // TRL_DECIDE_V1=1 selects the one-cluster-per-decision kernel (trl_decide.cuh)
    double* act2[8] = {nullptr};                 // one scratch matrix per side stream (the decisions of several env-steps can overlap)
    trl::FcMaps fc_maps[8];
    const double* fc_maps_w = nullptr;
    int fc_clusters = 2;
    int64_t launches = 0;
    std::map<long long, cudaGraphExec_t> graphs;   // keyed by the bit pattern of dt
    bool use_graph = true;
    // host staging
    std::vector<double> h_tuples;
    std::vector<float> h_tuples_f32;
    std::vector<uint32_t> h_tuple_flags;
    std::vector<int32_t> h_tuple_env;
    std::vector<double> h_dist;
    std::vector<int32_t> h_dist_env;
    std::vector<void*> allocs;
    void* flush_buf = nullptr;
    double bench_span_ms = 0.0;          // trl_bench_updates: first update's start to last update's end, flushes included
    double* probe_dump = nullptr;        // every blob of one forward pass (trl_probe.cu: trl_get_layer_state)
    // pipelined read-back (trl_snapshot / trl_snapshot_wait)
    cudaStream_t copy_stream = nullptr;
    cudaEvent_t snap_ready = nullptr, snap_copied = nullptr;
    double* snap_dev = nullptr;      // [2 * ndof * n + 4] device staging (pose planes, vel planes, stats)
    double* snap_host = nullptr;     // pinned mirror
    bool snap_pending = false;
};

// shared by the translation units behind the C ABI
int trl_fail(const std::string& msg);                 // records the message for trl_last_error(), returns 1
void trl_drop_graphs(trl_handle* h);
extern "C" void trl_trainer_orphan(trl_trainer* t);              // trl_destroy with a trainer still attached (trl_train.cu)
int trl_make_fc_maps(trl::FcMaps* out, const double* tip0_w, const double* act2, int rows);   // trl_host.cu
int trl_reupload_model(trl_handle* h);                // after editing h->mc: refresh the __constant__ copies                  // captured graphs bake kernel arguments (weight pointers) in
