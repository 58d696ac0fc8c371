// deepterrainrl_b200 -- policy decision kernel (included by trl_step.cu; same translation unit so the action
// bookkeeping device functions are shared).
//
// One CTA per environment that reached a gait-cycle boundary in the preceding step kernel (a persistent grid walks
// the pending list).  Thread 0 runs the scalar decision logic of cDogControllerMACE::UpdateAction /
// cBaseControllerMACE::DecideActionBoltzmann (sim/DogController.cpp:847-868, sim/BaseControllerMACE.cpp:254-318,
// 339-396, 437-518); all 256 threads evaluate the MACE network (data/policies/dog/nets/dog_mace3_deploy.prototxt:
// conv 1x8x16, 1x4x32, 1x4x32, FC 5984->64, FC 147->256, four 256->128->{3,29,29,29} heads) in f64 like the
// reference's Caffe Net<double> (learning/NeuralNet.h:13), with cNeuralNet::Eval's offset/scale normalisation
// (learning/NeuralNet.cpp:352-375, 977-986, 1027-1036).  Activations live in shared memory, weights stream from
// L2 (4.5 MB, resident).
#pragma once
#include "trl_types.h"

namespace trl {

constexpr int kDecideThreads = 1024;
constexpr int kConv0Out = 16, kConv0K = 8, kW0 = 193;
constexpr int kConv1Out = 32, kConv1K = 4, kW1 = 190;
constexpr int kConv2Out = 32, kConv2K = 4, kW2 = 187;
constexpr int kTip0Out = 64, kIp0Out = 256, kHeadHidden = 128;
// shared layout (doubles)
constexpr int kShX = 0;                                    // 283 (+pad)
constexpr int kShA = 288;                                  // A0 (16x193 = 3088) aliased with A2 (32x187 = 5984)
constexpr int kShB = kShA + kConv2Out * kW2;               // A1 (32x190 = 6080)
constexpr int kShCat = kShB + kConv1Out * kW1;             // 64 + n_char (<= 160)
constexpr int kShH = kShCat + 160;                         // 256
constexpr int kShHH = kShH + kIp0Out;                      // 4 x 128
constexpr int kShY = kShHH + 4 * kHeadHidden;              // 96
constexpr int kShCtl = kShY + kMaxNetOut;                  // control words
constexpr int kDecideSmemDoubles = kShCtl + 8;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ void net_forward(const NetWeights& W, const double* __restrict__ x_in, double* sh, int n_char, int n_frags, int frag) {
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarp = kDecideThreads / 32;
    const int n_in = 200 + n_char;
    double* X = sh + kShX;
    double* A0 = sh + kShA;
    double* A1 = sh + kShB;
    double* A2 = sh + kShA;
    double* CAT = sh + kShCat;
    double* H = sh + kShH;
    double* HH = sh + kShHH;
    double* Y = sh + kShY;
    for (int i = tid; i < n_in; i += kDecideThreads) X[i] = (x_in[i] + W.in_off[i]) * W.in_scale[i];
    __syncthreads();
    for (int idx = tid; idx < kConv0Out * kW0; idx += kDecideThreads) {
        int o = idx / kW0, t = idx - o * kW0;
        double acc = W.conv0_b[o];
#pragma unroll
        for (int k = 0; k < kConv0K; ++k) acc += W.conv0_w[o * kConv0K + k] * X[t + k];
        A0[idx] = acc > 0.0 ? acc : 0.0;
    }
    __syncthreads();
    for (int idx = tid; idx < kConv1Out * kW1; idx += kDecideThreads) {
        int o = idx / kW1, t = idx - o * kW1;
        double acc = W.conv1_b[o];
        const double* w = W.conv1_w + o * kConv0Out * kConv1K;
        for (int c = 0; c < kConv0Out; ++c) {
            const double* a = A0 + c * kW0 + t;
#pragma unroll
            for (int k = 0; k < kConv1K; ++k) acc += w[c * kConv1K + k] * a[k];
        }
        A1[idx] = acc > 0.0 ? acc : 0.0;
    }
    __syncthreads();
    for (int idx = tid; idx < kConv2Out * kW2; idx += kDecideThreads) {
        int o = idx / kW2, t = idx - o * kW2;
        double acc = W.conv2_b[o];
        const double* w = W.conv2_w + o * kConv1Out * kConv2K;
        for (int c = 0; c < kConv1Out; ++c) {
            const double* a = A1 + c * kW1 + t;
#pragma unroll
            for (int k = 0; k < kConv2K; ++k) acc += w[c * kConv2K + k] * a[k];
        }
        A2[idx] = acc > 0.0 ? acc : 0.0;
    }
    __syncthreads();
    // terr_ip0: 64 x 5984, one warp per output row (coalesced weight reads), warp-shuffle reduction
    const int nflat = kConv2Out * kW2;
    for (int o = warp; o < kTip0Out; o += nwarp) {
        const double* w = W.tip0_w + (size_t)o * nflat;
        double acc = 0.0;
        for (int i = lane; i < nflat; i += 32) acc += w[i] * A2[i];
        acc = warp_sum(acc);
        if (lane == 0) { acc += W.tip0_b[o]; CAT[o] = acc > 0.0 ? acc : 0.0; }
    }
    for (int i = tid; i < n_char; i += kDecideThreads) CAT[kTip0Out + i] = X[200 + i];
    __syncthreads();
    const int ncat = kTip0Out + n_char;
    for (int o = warp; o < kIp0Out; o += nwarp) {
        const double* w = W.ip0_w + (size_t)o * ncat;
        double acc = 0.0;
        for (int i = lane; i < ncat; i += 32) acc += w[i] * CAT[i];
        acc = warp_sum(acc);
        if (lane == 0) { acc += W.ip0_b[o]; H[o] = acc > 0.0 ? acc : 0.0; }
    }
    __syncthreads();
    for (int oo = warp; oo < 4 * kHeadHidden; oo += nwarp) {
        int hd = oo / kHeadHidden, o = oo - hd * kHeadHidden;
        const double* w = W.h0_w[hd] + (size_t)o * kIp0Out;
        double acc = 0.0;
        for (int i = lane; i < kIp0Out; i += 32) acc += w[i] * H[i];
        acc = warp_sum(acc);
        if (lane == 0) { acc += W.h0_b[hd][o]; HH[oo] = acc > 0.0 ? acc : 0.0; }
    }
    __syncthreads();
    const int n_out = n_frags + n_frags * frag;
    for (int oo = warp; oo < n_out; oo += nwarp) {
        int hd, o;
        if (oo < n_frags) { hd = 0; o = oo; }
        else { hd = 1 + (oo - n_frags) / frag; o = (oo - n_frags) - (hd - 1) * frag; }
        const double* w = W.h1_w[hd] + (size_t)o * kHeadHidden;
        const double* hh = HH + hd * kHeadHidden;
        double acc = 0.0;
        for (int i = lane; i < kHeadHidden; i += 32) acc += w[i] * hh[i];
        acc = warp_sum(acc);
        if (lane == 0) Y[oo] = (acc + W.h1_b[hd][o]) / W.out_scale[oo] - W.out_off[oo];
    }
    __syncthreads();
}

__global__ void __launch_bounds__(kDecideThreads)
trl_decide_kernel(Buffers B, NetWeights W, ExpSettings ex, int* done_count) {
    extern __shared__ double sh[];
    const ModelConst& m = c_model;
    const int count = *B.pending_count;
    int* ctl = reinterpret_cast<int*>(sh + kShCtl);
    for (int idx = blockIdx.x; idx < count; idx += gridDim.x) {
        const int env = B.pending_list[idx];
        Lane L{nullptr, env, B.n, B.d, B.i};
        CounterRng rng;
        double params[kNumParams];
        int id = 0;
        if (threadIdx.x == 0) {
            // cDogControllerMACE::UpdateAction: exploration flags cleared, off-policy until decided otherwise
            rng = load_rng(L);
            int eflags = 4;
            int need_net = 0;
            for (int k = 0; k < kNumParams; ++k) params[k] = L.d(D_PARAMS + k);
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
            ctl[0] = need_net;
            ctl[1] = eflags;
        }
        __syncthreads();
        const int need_net = ctl[0];
        if (need_net) {
            net_forward(W, B.poli_state + (size_t)env * B.S, sh, m.n_char, m.n_frags, m.frag);
            double* Y = sh + kShY;
            for (int i = threadIdx.x; i < m.n_out; i += kDecideThreads) B.net_out[(size_t)env * kMaxNetOut + i] = Y[i];
            if (threadIdx.x == 0) {
                int eflags = 0;
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
                for (int k = 0; k < fs; ++k) params[1 + k] = Y[nf + a * fs + k];
                params[mTransTime] = fabs(params[mTransTime]); params[mCv] = fabs(params[mCv]);
                if (ex.enable) {
                    double rn = rng.uniform();
                    if (rn < ex.rate) {              // ApplyExpNoiseAction
                        for (int k = 0; k < fs; ++k) params[1 + k] += (ex.noise * rng.normal()) * (1.0 / m.out_scale_actor0[k]);
                        eflags |= 2;
                    }
                    if (a != a_max) eflags |= 1;
                    if (eflags & 3) eflags |= 4;
                }
                ctl[1] = eflags;
            }
        }
        if (threadIdx.x == 0) {
            L.i(I_EXP_FLAGS) = ctl[1];
            apply_action(L, id, params, B.com_stash[env], B.com_stash[B.n + env]);
            L.i(I_PENDING) = 0;
            store_rng(L, rng);
        }
        __syncthreads();
    }
    // last CTA to finish re-arms the pending list for the next step
    if (threadIdx.x == 0) {
        __threadfence();
        int done = atomicAdd(done_count, 1);
        if (done == (int)gridDim.x - 1) { *B.pending_count = 0; *done_count = 0; __threadfence(); }
    }
}

size_t decide_smem_bytes() { return (size_t)kDecideSmemDoubles * sizeof(double); }
cudaError_t configure_decide_kernel() {
    return cudaFuncSetAttribute(trl_decide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)decide_smem_bytes());
}
void launch_decide(const Buffers& B, const NetWeights& W, const ExpSettings& ex, int* done_count, int grid, cudaStream_t st) {
    trl_decide_kernel<<<grid, kDecideThreads, decide_smem_bytes(), st>>>(B, W, ex, done_count);
}

}  // namespace trl
