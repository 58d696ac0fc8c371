// deepterrainrl_b200 -- policy decision kernel (included by trl_step.cu; same translation unit so the action
// bookkeeping device functions are shared).
//
// One thread-block CLUSTER (kClusterSize CTAs, 4 by default) per environment that reached a gait-cycle boundary in the preceding
// step kernel (a persistent grid of clusters walks the pending list); layers are split over the SMs of the cluster and activations
// are exchanged through distributed shared memory.  Thread 0 of rank 0 runs the scalar decision logic of cDogControllerMACE::UpdateAction /
// cBaseControllerMACE::DecideActionBoltzmann (sim/DogController.cpp:847-868, sim/BaseControllerMACE.cpp:254-318,
// 339-396, 437-518); all 8 x 512 threads evaluate the MACE network (data/policies/dog/nets/dog_mace3_deploy.prototxt:
// conv 1x8x16, 1x4x32, 1x4x32, FC 5984->64, FC 147->256, four 256->128->{3,29,29,29} heads) in f64 like the
// reference's Caffe Net<double> (learning/NeuralNet.h:13), with cNeuralNet::Eval's offset/scale normalisation
// (learning/NeuralNet.cpp:352-375, 977-986, 1027-1036).  Activations live in shared memory, weights stream from
// L2 (4.5 MB, resident).
#pragma once
#include <cooperative_groups.h>

#include "trl_types.h"

namespace trl {
namespace cg = cooperative_groups;

constexpr int kDecideThreads = 512;
#ifndef TRL_CLUSTER
#define TRL_CLUSTER 4    // measured: 4-CTA clusters 21.5 M env-steps/s, 8-CTA 20.6 M, 2-CTA 21.4 M (smaller footprint beside the step launch)
#endif
#ifndef TRL_DECIDE_TILE
#define TRL_DECIDE_TILE 0   // 1: register-tiled conv1 / conv2 (experiment, bit-identical; see net_forward_cluster)
#endif
#ifndef TRL_CONV_TILE
#define TRL_CONV_TILE 4
#endif
constexpr int kConvTile = TRL_CONV_TILE;                                         // adjacent output positions per thread in the tiled conv loops
constexpr int kClusterSize = TRL_CLUSTER;                            // CTAs (SMs) cooperating on one decision
constexpr int kConv0Out = 16, kConv0K = 8, kW0 = 193;
constexpr int kConv1Out = 32, kConv1K = 4, kW1 = 190;
constexpr int kConv2Out = 32, kConv2K = 4, kW2 = 187;
constexpr int kTip0Out = 64, kIp0Out = 256, kHeadHidden = 128;
constexpr int kC1Slice = kConv1Out / kClusterSize;         // conv1 / conv2 output channels per CTA
constexpr int kC2Slice = kConv2Out / kClusterSize;
// shared layout per CTA (doubles)
constexpr int kShX = 0;                                    // 283 (+pad) normalised input
constexpr int kShA0 = 288;                                 // conv0 output, full 16 x 193 (recomputed by every CTA)
constexpr int kShA1 = kShA0 + kConv0Out * kW0;             // conv1 output, full 32 x 190 (own slice computed, rest gathered)
constexpr int kShA2 = kShA1 + kConv1Out * kW1;             // conv2 output, own 4-channel slice 4 x 187
constexpr int kShT = kShA2 + kC2Slice * kW2;               // terr_ip0 partial sums of this CTA (64)
constexpr int kShCat = kShT + kTip0Out;                    // 64 + n_char (<= 160), reduced
constexpr int kShH = kShCat + 160;                         // ip0 output, full 256 (own 32 computed, rest gathered)
constexpr int kShHH = kShH + kIp0Out;                      // head hidden, full 4 x 128 (own 64 computed, rest gathered)
constexpr int kShY = kShHH + 4 * kHeadHidden;              // 96 (rank 0)
constexpr int kShW0 = kShY + kMaxNetOut;                   // conv0 weights 16 x 8 + 16 biases
constexpr int kShW1 = kShW0 + kConv0Out * kConv0K + kConv0Out;   // conv1 weight slice 4 x 16 x 4
constexpr int kShW2 = kShW1 + kC1Slice * kConv0Out * kConv1K;   // conv2 weight slice 4 x 32 x 4
constexpr int kShCtl = kShW2 + kC2Slice * kConv1Out * kConv2K;  // control words (rank 0)
constexpr int kDecideSmemDoubles = kShCtl + 8;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// MACE forward pass of ONE decision spread over a thread-block cluster: every CTA owns 1/8 of the output channels /
// rows of each layer and exchanges activations through distributed shared memory.  The pass is latency-bound (one
// decision, ~3.5 MFLOP over the cluster's SMs), so every weight read is either staged in shared memory up front or issued as a
// batch of independent loads before the first use.  Result (n_out values) lands in rank 0's Y.
__device__ void net_forward_cluster(cg::cluster_group& cluster, const NetWeights& W, const double* __restrict__ x_in, double* sh,
                                    int n_char, int n_frags, int frag) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarp = kDecideThreads / 32;
    const int rank = (int)cluster.block_rank();
    const int n_in = 200 + n_char;
    double* X = sh + kShX;
    double* A0 = sh + kShA0;
    double* A1 = sh + kShA1;
    double* A2 = sh + kShA2;
    double* T = sh + kShT;
    double* CAT = sh + kShCat;
    double* H = sh + kShH;
    double* HH = sh + kShHH;
    double* Y = sh + kShY;
    double* W0s = sh + kShW0;
    double* W1s = sh + kShW1;
    double* W2s = sh + kShW2;
    // ---- stage the input and every small weight block this CTA needs (independent coalesced loads)
    for (int i = tid; i < n_in; i += kDecideThreads) X[i] = (x_in[i] + W.in_off[i]) * W.in_scale[i];
    for (int i = tid; i < kConv0Out * kConv0K; i += kDecideThreads) W0s[i] = W.conv0_w[i];
    for (int i = tid; i < kConv0Out; i += kDecideThreads) W0s[kConv0Out * kConv0K + i] = W.conv0_b[i];
    for (int i = tid; i < kC1Slice * kConv0Out * kConv1K; i += kDecideThreads) W1s[i] = W.conv1_w[rank * kC1Slice * kConv0Out * kConv1K + i];
    for (int i = tid; i < kC2Slice * kConv1Out * kConv2K; i += kDecideThreads) W2s[i] = W.conv2_w[rank * kC2Slice * kConv1Out * kConv2K + i];
    __syncthreads();
    for (int idx = tid; idx < kConv0Out * kW0; idx += kDecideThreads) {
        int o = idx / kW0, t = idx - o * kW0;
        double acc = W0s[kConv0Out * kConv0K + o];
#pragma unroll
        for (int k = 0; k < kConv0K; ++k) acc += W0s[o * kConv0K + k] * X[t + k];
        A0[idx] = acc > 0.0 ? acc : 0.0;
    }
    __syncthreads();
#if TRL_DECIDE_TILE
    // Experiment: kConvTile adjacent output positions per thread -- the activations a[t .. t + tile + K - 2] and the K weights of
    // an input channel are loaded once and used for tile x K multiply-adds (shared-memory loads per multiply-add 2.0 -> 0.7; the
    // untiled loop is shared-memory-bandwidth bound).  Every output still sums in the same order (4 accumulators by c & 3, k
    // inner), so the results are bit-identical to the loop below.
    {
        constexpr int nt = (kW1 + kConvTile - 1) / kConvTile;
        for (int item = tid; item < kC1Slice * nt; item += kDecideThreads) {
            const int ol = item / nt, t0 = (item - ol * nt) * kConvTile, o = rank * kC1Slice + ol;
            const double* w = W1s + ol * kConv0Out * kConv1K;
            double acc[kConvTile][4];
#pragma unroll
            for (int j = 0; j < kConvTile; ++j) { acc[j][0] = W.conv1_b[o]; acc[j][1] = 0.0; acc[j][2] = 0.0; acc[j][3] = 0.0; }
#pragma unroll
            for (int c = 0; c < kConv0Out; ++c) {
                const double* a = A0 + c * kW0;
                double av[kConvTile + kConv1K - 1];
#pragma unroll
                for (int i = 0; i < kConvTile + kConv1K - 1; ++i) av[i] = a[min(t0 + i, kW0 - 1)];
#pragma unroll
                for (int k = 0; k < kConv1K; ++k) {
                    const double wk = w[c * kConv1K + k];
#pragma unroll
                    for (int j = 0; j < kConvTile; ++j) acc[j][c & 3] += wk * av[j + k];
                }
            }
#pragma unroll
            for (int j = 0; j < kConvTile; ++j) {
                const double v = (acc[j][0] + acc[j][1]) + (acc[j][2] + acc[j][3]);
                if (t0 + j < kW1) A1[o * kW1 + t0 + j] = v > 0.0 ? v : 0.0;
            }
        }
    }
#else
    // conv1: this CTA's 4 output channels; 4 independent accumulators per output
    for (int idx = tid; idx < kC1Slice * kW1; idx += kDecideThreads) {
        int ol = idx / kW1, t = idx - ol * kW1, o = rank * kC1Slice + ol;
        const double* w = W1s + ol * kConv0Out * kConv1K;
        double acc[4] = {W.conv1_b[o], 0.0, 0.0, 0.0};
#pragma unroll
        for (int c = 0; c < kConv0Out; ++c) {
            const double* a = A0 + c * kW0 + t;
#pragma unroll
            for (int k = 0; k < kConv1K; ++k) acc[c & 3] += w[c * kConv1K + k] * a[k];
        }
        double v = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        A1[o * kW1 + t] = v > 0.0 ? v : 0.0;
    }
#endif
    cluster.sync();
    // gather the other CTAs' conv1 slices through DSMEM
    for (int r = 1; r < kClusterSize; ++r) {
        int src = (rank + r) % kClusterSize;
        const double* remote = cluster.map_shared_rank(A1, src);
        for (int i = tid; i < kC1Slice * kW1; i += kDecideThreads) {
            int off = src * kC1Slice * kW1 + i;
            A1[off] = remote[off];
        }
    }
    __syncthreads();
#if TRL_DECIDE_TILE
    {
        constexpr int nt = (kW2 + kConvTile - 1) / kConvTile;
        for (int item = tid; item < kC2Slice * nt; item += kDecideThreads) {
            const int ol = item / nt, t0 = (item - ol * nt) * kConvTile, o = rank * kC2Slice + ol;
            const double* w = W2s + ol * kConv1Out * kConv2K;
            double acc[kConvTile][4];
#pragma unroll
            for (int j = 0; j < kConvTile; ++j) { acc[j][0] = W.conv2_b[o]; acc[j][1] = 0.0; acc[j][2] = 0.0; acc[j][3] = 0.0; }
#pragma unroll 8
            for (int c = 0; c < kConv1Out; ++c) {
                const double* a = A1 + c * kW1;
                double av[kConvTile + kConv2K - 1];
#pragma unroll
                for (int i = 0; i < kConvTile + kConv2K - 1; ++i) av[i] = a[min(t0 + i, kW1 - 1)];
#pragma unroll
                for (int k = 0; k < kConv2K; ++k) {
                    const double wk = w[c * kConv2K + k];
#pragma unroll
                    for (int j = 0; j < kConvTile; ++j) acc[j][c & 3] += wk * av[j + k];
                }
            }
#pragma unroll
            for (int j = 0; j < kConvTile; ++j) {
                const double v = (acc[j][0] + acc[j][1]) + (acc[j][2] + acc[j][3]);
                if (t0 + j < kW2) A2[ol * kW2 + t0 + j] = v > 0.0 ? v : 0.0;
            }
        }
    }
#else
    // conv2: this CTA's 4 output channels; 4 independent accumulators per output
    for (int idx = tid; idx < kC2Slice * kW2; idx += kDecideThreads) {
        int ol = idx / kW2, t = idx - ol * kW2, o = rank * kC2Slice + ol;
        const double* w = W2s + ol * kConv1Out * kConv2K;
        double acc[4] = {W.conv2_b[o], 0.0, 0.0, 0.0};
#pragma unroll 8
        for (int c = 0; c < kConv1Out; ++c) {
            const double* a = A1 + c * kW1 + t;
#pragma unroll
            for (int k = 0; k < kConv2K; ++k) acc[c & 3] += w[c * kConv2K + k] * a[k];
        }
        double v = (acc[0] + acc[1]) + (acc[2] + acc[3]);
        A2[idx] = v > 0.0 ? v : 0.0;
    }
#endif
    __syncthreads();
    // terr_ip0 (64 x 5984): K-split -- this CTA multiplies its own 4 x 187 slice of the flattened conv2 output.
    // 4 rows per warp, 4 columns per trip: 16 independent weight loads in flight per lane.
    const int nflat = kConv2Out * kW2, nslice = kC2Slice * kW2;
    {
        constexpr int kRows = kTip0Out / nwarp;   // 4
        const double* w0 = W.tip0_w + (size_t)(warp * kRows) * nflat + (size_t)rank * nslice;
        double acc[kRows];
#pragma unroll
        for (int r = 0; r < kRows; ++r) acc[r] = 0.0;
        for (int i0 = 0; i0 < nslice; i0 += 4 * 32) {
            double wv[kRows][4], av[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                int i = i0 + u * 32 + lane;
                bool ok = i < nslice;
                av[u] = ok ? A2[i] : 0.0;
#pragma unroll
                for (int r = 0; r < kRows; ++r) wv[r][u] = ok ? w0[(size_t)r * nflat + i] : 0.0;
            }
#pragma unroll
            for (int u = 0; u < 4; ++u)
#pragma unroll
                for (int r = 0; r < kRows; ++r) acc[r] += wv[r][u] * av[u];
        }
#pragma unroll
        for (int r = 0; r < kRows; ++r) {
            double v = warp_sum(acc[r]);
            if (lane == 0) T[warp * kRows + r] = v;
        }
    }
    // prefetch this warp's ip0 rows while the cluster synchronises (2 rows x <= 5 columns per lane)
    const int ncat = kTip0Out + n_char;
    constexpr int kIpRows = (kIp0Out / kClusterSize) / nwarp;   // 2
    const int ip_o0 = rank * (kIp0Out / kClusterSize) + warp * kIpRows;
    double ipw[kIpRows][5], ipb[kIpRows];
#pragma unroll
    for (int r = 0; r < kIpRows; ++r) {
        ipb[r] = W.ip0_b[ip_o0 + r];
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            int i = u * 32 + lane;
            ipw[r][u] = i < ncat ? W.ip0_w[(size_t)(ip_o0 + r) * ncat + i] : 0.0;
        }
    }
    const double tb = tid < kTip0Out ? W.tip0_b[tid] : 0.0;
    cluster.sync();
    // every CTA reduces the 8 partial vectors in rank order (deterministic) and builds concat0 = [terr_relu3 | char]
    if (tid < kTip0Out) {
        double part[kClusterSize];
#pragma unroll
        for (int r = 0; r < kClusterSize; ++r) part[r] = cluster.map_shared_rank(T, r)[tid];
        double acc = tb;
#pragma unroll
        for (int r = 0; r < kClusterSize; ++r) acc += part[r];
        CAT[tid] = acc > 0.0 ? acc : 0.0;
    }
    for (int i = tid; i < n_char; i += kDecideThreads) CAT[kTip0Out + i] = X[200 + i];
    __syncthreads();
    // ip0: 256 rows, 32 per CTA, 2 per warp (weights already in registers)
    {
        double acc[kIpRows];
#pragma unroll
        for (int r = 0; r < kIpRows; ++r) acc[r] = 0.0;
#pragma unroll
        for (int u = 0; u < 5; ++u) {
            int i = u * 32 + lane;
            double a = i < ncat ? CAT[i] : 0.0;
#pragma unroll
            for (int r = 0; r < kIpRows; ++r) acc[r] += ipw[r][u] * a;
        }
#pragma unroll
        for (int r = 0; r < kIpRows; ++r) {
            double v = warp_sum(acc[r]) + ipb[r];
            if (lane == 0) H[ip_o0 + r] = v > 0.0 ? v : 0.0;
        }
    }
    // prefetch this warp's head-hidden rows (4 rows x 8 columns per lane) before the exchange of H
    constexpr int kHdRows = (4 * kHeadHidden / kClusterSize) / nwarp;   // 4 (never straddles two heads)
    const int hd_oo0 = rank * (4 * kHeadHidden / kClusterSize) + warp * kHdRows;
    const int hd_h = hd_oo0 / kHeadHidden, hd_o0 = hd_oo0 - hd_h * kHeadHidden;
    double hw[kHdRows][8], hb[kHdRows];
    {
        const double* w0 = W.h0_w[hd_h] + (size_t)hd_o0 * kIp0Out;
#pragma unroll
        for (int r = 0; r < kHdRows; ++r) {
            hb[r] = W.h0_b[hd_h][hd_o0 + r];
#pragma unroll
            for (int u = 0; u < 8; ++u) hw[r][u] = w0[(size_t)r * kIp0Out + u * 32 + lane];
        }
    }
    cluster.sync();
    for (int r = 1; r < kClusterSize; ++r) {
        int src = (rank + r) % kClusterSize;
        const double* remote = cluster.map_shared_rank(H, src);
        for (int i = tid; i < kIp0Out / kClusterSize; i += kDecideThreads) {
            int off = src * (kIp0Out / kClusterSize) + i;
            H[off] = remote[off];
        }
    }
    __syncthreads();
    {
        double acc[kHdRows];
#pragma unroll
        for (int r = 0; r < kHdRows; ++r) acc[r] = 0.0;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            double a = H[u * 32 + lane];
#pragma unroll
            for (int r = 0; r < kHdRows; ++r) acc[r] += hw[r][u] * a;
        }
#pragma unroll
        for (int r = 0; r < kHdRows; ++r) {
            double v = warp_sum(acc[r]) + hb[r];
            if (lane == 0) HH[hd_oo0 + r] = v > 0.0 ? v : 0.0;
        }
    }
    // output layer: one output row per warp across the whole cluster (8 x 16 warps >= 90 rows with the default cluster size;
    // smaller clusters take several rows per warp); the first row's weights are prefetched before the cluster synchronises
    const int n_out = n_frags + n_frags * frag;
    const int oo0 = rank * nwarp + warp;
    int f_hd = 0, f_o = 0;
    double fw[4] = {0, 0, 0, 0}, fb = 0.0, fsc = 1.0, fof = 0.0;
    auto prefetch_row = [&](int oo) {
        if (oo < n_frags) { f_hd = 0; f_o = oo; }
        else { f_hd = 1 + (oo - n_frags) / frag; f_o = (oo - n_frags) - (f_hd - 1) * frag; }
        const double* w = W.h1_w[f_hd] + (size_t)f_o * kHeadHidden;
#pragma unroll
        for (int u = 0; u < 4; ++u) fw[u] = w[u * 32 + lane];
        fb = W.h1_b[f_hd][f_o]; fsc = W.out_scale[oo]; fof = W.out_off[oo];
    };
    if (oo0 < n_out) prefetch_row(oo0);
    cluster.sync();
    for (int oo = oo0; oo < n_out; oo += kClusterSize * nwarp) {
        if (oo != oo0) prefetch_row(oo);
        // the 128 hidden activations of head f_hd live in the CTAs that computed them
        double acc = 0.0;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            int j = f_hd * kHeadHidden + u * 32 + lane;
            int owner = j / (4 * kHeadHidden / kClusterSize);
            acc += fw[u] * cluster.map_shared_rank(HH, owner)[j];
        }
        acc = warp_sum(acc);
        if (lane == 0) cluster.map_shared_rank(Y, 0)[oo] = (acc + fb) / fsc - fof;
    }
    cluster.sync();   // Y complete in rank 0; peers may reuse T / H / HH afterwards
}

#ifndef TRL_DECIDE_MIN_BLOCKS
#define TRL_DECIDE_MIN_BLOCKS 2
#endif
__global__ void __cluster_dims__(kClusterSize, 1, 1) __launch_bounds__(kDecideThreads, TRL_DECIDE_MIN_BLOCKS)
trl_decide_kernel(Buffers B, NetWeights W, const ExpSettings* __restrict__ ex_dev, int* done_count, int list, int rearm) {
    TRL_DYN_SHARED(double, sh);
    cg::cluster_group cluster = cg::this_cluster();
    const ModelConst& m = c_model;
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / kClusterSize, ncl = gridDim.x / kClusterSize;
    const int count = B.pending_count[list];
    const ExpSettings ex = *ex_dev;
    for (int idx = cid; idx < count; idx += ncl) {
        const int env = B.pending_list[list * B.n + idx];
        Lane L{nullptr, env, B.n, B.d, B.i};
        CounterRng rng;
        double params[kNumParams];
        int id = 0, eflags = 4, need_net = 0;
        const bool boss = (rank == 0 && threadIdx.x == 0);
        if (boss) {
            // cDogControllerMACE::UpdateAction: exploration flags cleared, off-policy until decided otherwise
            rng = load_rng(L);
            for (int k = 0; k < m.n_params; ++k) params[k] = L.d(D_PARAMS + k);
            id = L.i(I_ACTION_ID);
            int cmd = L.i(I_CMD);
            if (cmd >= 0) {
                if (m.is_mace) eflags |= 3;
                id = build_base_action(L, rng, cmd, params);
                L.i(I_CMD) = -1;
            } else if (m.has_net) {
                double base_rand = rng.uniform();
                if (ex.enable && base_rand < ex.base_rate) {
                    int a = rng.rand_int(0, m.n_actions);
                    id = build_base_action(L, rng, a, params);
                    eflags = 4 | 3;
                } else {
                    need_net = 1;
                    eflags = 0;
                }
            } else {
                bool cyclic = m.is_mace ? false : (m.act_cyclic[id] != 0);
                if (!cyclic) id = build_base_action(L, rng, m.default_action, params);
            }
        }
        // the forward pass does not depend on the scalar branch above: all CTAs run it whenever the scene has a net
        // (its result is simply unused for the rare command / base-action decisions)
        if (m.has_net) {
            net_forward_cluster(cluster, W, B.poli_state + (size_t)env * B.S, sh, m.n_char, m.n_frags, m.frag);
            double* Y = sh + kShY;
            if (boss && need_net) {
                for (int i = 0; i < m.n_out; ++i) B.net_out[(size_t)env * kMaxNetOut + i] = Y[i];
                eflags = 0;
                const int nf = m.n_frags, fs = m.frag;
                int a_max = 0;
                for (int i = 1; i < nf; ++i) if (Y[i] > Y[a_max]) a_max = i;
                int a = a_max;
                if (ex.enable && ex.temp != 0.0) {   // BoltzmannSelectActor
                    double vals[8], sum = 0.0;
                    for (int i = 0; i < nf; ++i) { vals[i] = exp((Y[i] - Y[a_max]) / ex.temp); sum += vals[i]; }
                    double r = rng.uniform() * sum;
                    for (int i = 0; i < nf; ++i) { r -= vals[i]; if (r <= 0.0) { a = i; break; } }
                }
                // BuildActorAction: actor `a`'s 29 outputs overwrite params[1:30] of the current action
                id = a;
                for (int k = 0; k < fs; ++k) params[m.opt_idx[k]] = Y[nf + a * fs + k];
                params[mTransTime] = fabs(params[mTransTime]); params[mCv] = fabs(params[mCv]);
                if (m.char_type == 2) params[rmCd] = fabs(params[rmCd]);
                if (ex.enable) {
                    double rn = rng.uniform();
                    if (rn < ex.rate) {              // ApplyExpNoiseAction
                        for (int k = 0; k < fs; ++k) params[m.opt_idx[k]] += (ex.noise * rng.normal()) * (1.0 / m.out_scale_actor0[k]);
                        eflags |= 2;
                    }
                    if (a != a_max) eflags |= 1;
                    if (eflags & 3) eflags |= 4;
                }
            }
        }
        if (boss) {
            L.i(I_EXP_FLAGS) = eflags;
            apply_action(L, id, params, B.com_stash[env], B.com_stash[B.n + env]);
            store_rng(L, rng);
        }
        cluster.sync();   // rank 0's Y / peers' activations are reused by the next decision of this cluster
    }
    // serial schedule: the last CTA to finish re-arms the list (in the overlapped schedule the catch-up launch, which
    // still needs the count, does it)
    if (rearm && threadIdx.x == 0) {
        __threadfence();
        int done = atomicAdd(done_count, 1);
        if (done == (int)gridDim.x - 1) { B.pending_count[list] = 0; *done_count = 0; __threadfence(); }
    }
}

size_t decide_smem_bytes() { return (size_t)kDecideSmemDoubles * sizeof(double); }
cudaError_t configure_decide_kernel() {
    return cudaFuncSetAttribute(trl_decide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decide_smem_bytes());
}
void launch_decide(const Buffers& B, const NetWeights& W, const ExpSettings* ex, int* done_count, int grid, int list, int rearm,
                   cudaStream_t st) {
    TRL_LAUNCH_CLUSTER(kClusterSize, trl_decide_kernel, grid, kDecideThreads, decide_smem_bytes(), st, B, W, ex, done_count, list, rearm);
}

}  // namespace trl
