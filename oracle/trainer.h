// TEST INFRASTRUCTURE ONLY -- CPU oracle (see oracle/README.md). Never linked into the product path.
//
// cMACETrainer restated (learning/MACETrainer.cpp, learning/NeuralNetTrainer.cpp) together with the pieces of Caffe it
// drives: batch forward / backward of the MACE topology (data/policies/dog/nets/dog_mace3_train.prototxt: MemoryData batch 32,
// EuclideanLoss over the 90 normalised outputs) and the SGD solver step (dog_mace3_solver.prototxt: base_lr 1e-3 fixed,
// momentum 0.9, weight_decay 5e-4, L2; per-blob lr_mult 1 / 2 and decay_mult 1 / 0 from the train prototxt -- the three
// convolution layers give no decay_mult, so both their blobs decay).  Caffe (niuzhiheng/caffe @ 7b3e6f2) is an absent third
// party dependency: its published SGDSolver algorithm is restated (Regularize: diff += decay * data; ComputeUpdateValue:
// history = rate * diff + momentum * history; Update: data -= history; EuclideanLoss: sum (a-b)^2 / 2N, d/da = (a-b)/N).
// "parity unpinned" against Caffe itself; pinned by finite-difference gradient checks (tests/test_trainer_cpu.py).
//
// Sampling uses the engine's counter RNG instead of the reference's process-global std engine (util/MathUtil.cpp:4).
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "env.h"

namespace orc {

struct MaceTopo {
    int n_in = 283, n_char = 83, n_out = 90, n_frags = 3, frag = 29;
    static constexpr int C0 = 16, K0 = 8, W0 = 193, C1 = 32, K1 = 4, W1 = 190, C2 = 32, K2 = 4, W2 = 187;
    static constexpr int T = 64, H = 256, HH = 128;
    int blob_size(int b) const {
        const int cat = T + n_char;
        switch (b) {
            case 0: return C0 * 1 * K0;        case 1: return C0;
            case 2: return C1 * C0 * K1;       case 3: return C1;
            case 4: return C2 * C1 * K2;       case 5: return C2;
            case 6: return T * C2 * W2;        case 7: return T;
            case 8: return H * cat;            case 9: return H;
            default: {
                int h = (b - 10) / 4, r = (b - 10) % 4;      // head h: ip0_w, ip0_b, ip1_w, ip1_b
                int nout = (h == 0) ? n_frags : frag;
                if (r == 0) return HH * H;
                if (r == 1) return HH;
                if (r == 2) return nout * HH;
                return nout;
            }
        }
    }
    std::vector<size_t> offsets() const {
        std::vector<size_t> o(27, 0);
        for (int b = 0; b < 26; ++b) o[b + 1] = o[b] + blob_size(b);
        return o;
    }
    // Caffe param multipliers of blob b (dog_mace3_train.prototxt:36-39,132-137)
    double lr_mult(int b) const { return (b & 1) ? 2.0 : 1.0; }
    double decay_mult(int b) const { return (b & 1) ? (b < 6 ? 1.0 : 0.0) : 1.0; }
};

struct BatchActs {
    std::vector<double> xn, a0, a1, a2, t, cat, h, hh, y;     // y: raw (normalised) net output [B][n_out]
};

struct MaceNet {
    MaceTopo tp;
    std::vector<size_t> off;
    std::vector<double> theta;                        // 26 blobs, NET_LAYERS order (w, b per layer)
    std::vector<double> in_off, in_scale, out_off, out_scale;

    void init_from(const Net& n) {
        tp.n_in = n.n_in; tp.n_char = n.n_char; tp.n_out = n.n_out; tp.n_frags = n.n_frags; tp.frag = n.frag;
        off = tp.offsets();
        theta.assign(off[26], 0.0);
        const std::vector<double>* blobs[26] = {&n.conv0_w, &n.conv0_b, &n.conv1_w, &n.conv1_b, &n.conv2_w, &n.conv2_b, &n.tip0_w,
                                                &n.tip0_b, &n.ip0_w, &n.ip0_b};
        for (int h = 0; h < 4; ++h) {
            blobs[10 + 4 * h] = &n.head0_w[h]; blobs[11 + 4 * h] = &n.head0_b[h];
            blobs[12 + 4 * h] = &n.head1_w[h]; blobs[13 + 4 * h] = &n.head1_b[h];
        }
        for (int b = 0; b < 26; ++b) std::copy(blobs[b]->begin(), blobs[b]->end(), theta.begin() + off[b]);
        in_off = n.in_off; in_scale = n.in_scale; out_off = n.out_off; out_scale = n.out_scale;
    }
    // the inverse of init_from: the trained weights back into the policy network the environments evaluate (what
    // cNeuralNetLearner::SyncNet does with the controller's network after every cNeuralNetLearner::Train)
    void store_to(Net& n) const {
        std::vector<double>* blobs[26] = {&n.conv0_w, &n.conv0_b, &n.conv1_w, &n.conv1_b, &n.conv2_w, &n.conv2_b, &n.tip0_w,
                                          &n.tip0_b, &n.ip0_w, &n.ip0_b};
        for (int h = 0; h < 4; ++h) {
            blobs[10 + 4 * h] = &n.head0_w[h]; blobs[11 + 4 * h] = &n.head0_b[h];
            blobs[12 + 4 * h] = &n.head1_w[h]; blobs[13 + 4 * h] = &n.head1_b[h];
        }
        for (int b = 0; b < 26; ++b) std::copy(theta.begin() + off[b], theta.begin() + off[b + 1], blobs[b]->begin());
        n.in_off = in_off; n.in_scale = in_scale; n.out_off = out_off; n.out_scale = out_scale;
    }
    const double* blob(int b) const { return theta.data() + off[b]; }

    static void conv_fwd(int B, const double* x, int cin, int win, const double* w, const double* b, int cout, int k, double* y) {
        const int wout = win - k + 1;
        for (int n = 0; n < B; ++n)
            for (int o = 0; o < cout; ++o)
                for (int t = 0; t < wout; ++t) {
                    double acc = b[o];
                    for (int c = 0; c < cin; ++c)
                        for (int kk = 0; kk < k; ++kk)
                            acc += w[((size_t)o * cin + c) * k + kk] * x[((size_t)n * cin + c) * win + t + kk];
                    y[((size_t)n * cout + o) * wout + t] = acc > 0 ? acc : 0;
                }
    }
    static void fc_fwd(int B, const double* x, int nin, const double* w, const double* b, int nout, bool relu, double* y, int ldy) {
        for (int n = 0; n < B; ++n)
            for (int o = 0; o < nout; ++o) {
                double acc = b[o];
                const double* wr = w + (size_t)o * nin;
                const double* xr = x + (size_t)n * nin;
                for (int i = 0; i < nin; ++i) acc += wr[i] * xr[i];
                y[(size_t)n * ldy + o] = (relu && acc < 0) ? 0 : acc;
            }
    }
    // xn: normalised input [B][n_in]; fills the activations and the raw output
    void forward(int B, const double* xn, BatchActs& A) const {
        const int cat = MaceTopo::T + tp.n_char;
        A.xn.assign(xn, xn + (size_t)B * tp.n_in);
        std::vector<double> terr((size_t)B * 200);
        for (int n = 0; n < B; ++n) std::copy(xn + (size_t)n * tp.n_in, xn + (size_t)n * tp.n_in + 200, terr.begin() + (size_t)n * 200);
        A.a0.assign((size_t)B * MaceTopo::C0 * MaceTopo::W0, 0); A.a1.assign((size_t)B * MaceTopo::C1 * MaceTopo::W1, 0);
        A.a2.assign((size_t)B * MaceTopo::C2 * MaceTopo::W2, 0); A.t.assign((size_t)B * MaceTopo::T, 0);
        A.cat.assign((size_t)B * cat, 0); A.h.assign((size_t)B * MaceTopo::H, 0); A.hh.assign((size_t)B * 4 * MaceTopo::HH, 0);
        A.y.assign((size_t)B * tp.n_out, 0);
        conv_fwd(B, terr.data(), 1, 200, blob(0), blob(1), MaceTopo::C0, MaceTopo::K0, A.a0.data());
        conv_fwd(B, A.a0.data(), MaceTopo::C0, MaceTopo::W0, blob(2), blob(3), MaceTopo::C1, MaceTopo::K1, A.a1.data());
        conv_fwd(B, A.a1.data(), MaceTopo::C1, MaceTopo::W1, blob(4), blob(5), MaceTopo::C2, MaceTopo::K2, A.a2.data());
        fc_fwd(B, A.a2.data(), MaceTopo::C2 * MaceTopo::W2, blob(6), blob(7), MaceTopo::T, true, A.t.data(), MaceTopo::T);
        for (int n = 0; n < B; ++n) {
            std::copy(A.t.begin() + (size_t)n * MaceTopo::T, A.t.begin() + (size_t)(n + 1) * MaceTopo::T, A.cat.begin() + (size_t)n * cat);
            std::copy(xn + (size_t)n * tp.n_in + 200, xn + (size_t)(n + 1) * tp.n_in, A.cat.begin() + (size_t)n * cat + MaceTopo::T);
        }
        fc_fwd(B, A.cat.data(), cat, blob(8), blob(9), MaceTopo::H, true, A.h.data(), MaceTopo::H);
        int col = 0;
        for (int hd = 0; hd < 4; ++hd) {
            const int nout = hd == 0 ? tp.n_frags : tp.frag;
            std::vector<double> h0((size_t)B * MaceTopo::HH);
            fc_fwd(B, A.h.data(), MaceTopo::H, blob(10 + 4 * hd), blob(11 + 4 * hd), MaceTopo::HH, true, h0.data(), MaceTopo::HH);
            for (int n = 0; n < B; ++n)
                std::copy(h0.begin() + (size_t)n * MaceTopo::HH, h0.begin() + (size_t)(n + 1) * MaceTopo::HH,
                          A.hh.begin() + ((size_t)n * 4 + hd) * MaceTopo::HH);
            fc_fwd(B, h0.data(), MaceTopo::HH, blob(12 + 4 * hd), blob(13 + 4 * hd), nout, false, A.y.data() + col, tp.n_out);
            col += nout;
        }
    }

    // fc backward: dy [B][nout] (leading dim ldy), x [B][nin] -> dw += dy^T x, db += sum dy, dx = dy W (if dx)
    static void fc_bwd(int B, const double* dy, int ldy, int nout, const double* x, int nin, const double* w, double* dw, double* db,
                       double* dx) {
        for (int o = 0; o < nout; ++o) {
            double sb = 0;
            for (int n = 0; n < B; ++n) sb += dy[(size_t)n * ldy + o];
            db[o] += sb;
            for (int i = 0; i < nin; ++i) {
                double s = 0;
                for (int n = 0; n < B; ++n) s += dy[(size_t)n * ldy + o] * x[(size_t)n * nin + i];
                dw[(size_t)o * nin + i] += s;
            }
        }
        if (dx)
            for (int n = 0; n < B; ++n)
                for (int i = 0; i < nin; ++i) {
                    double s = 0;
                    for (int o = 0; o < nout; ++o) s += dy[(size_t)n * ldy + o] * w[(size_t)o * nin + i];
                    dx[(size_t)n * nin + i] = s;
                }
    }
    static void conv_bwd(int B, const double* dy, int cout, int k, const double* x, int cin, int win, const double* w, double* dw,
                         double* db, double* dx) {
        const int wout = win - k + 1;
        for (int o = 0; o < cout; ++o) {
            double sb = 0;
            for (int n = 0; n < B; ++n)
                for (int t = 0; t < wout; ++t) sb += dy[((size_t)n * cout + o) * wout + t];
            db[o] += sb;
            for (int c = 0; c < cin; ++c)
                for (int kk = 0; kk < k; ++kk) {
                    double s = 0;
                    for (int n = 0; n < B; ++n)
                        for (int t = 0; t < wout; ++t) s += dy[((size_t)n * cout + o) * wout + t] * x[((size_t)n * cin + c) * win + t + kk];
                    dw[((size_t)o * cin + c) * k + kk] += s;
                }
        }
        if (dx)
            for (int n = 0; n < B; ++n)
                for (int c = 0; c < cin; ++c)
                    for (int s = 0; s < win; ++s) {
                        double acc = 0;
                        for (int o = 0; o < cout; ++o)
                            for (int kk = 0; kk < k; ++kk) {
                                int t = s - kk;
                                if (t >= 0 && t < wout) acc += dy[((size_t)n * cout + o) * wout + t] * w[((size_t)o * cin + c) * k + kk];
                            }
                        dx[((size_t)n * cin + c) * win + s] = acc;
                    }
    }
    static void relu_mask(std::vector<double>& d, const std::vector<double>& act) {
        for (size_t i = 0; i < d.size(); ++i) if (!(act[i] > 0)) d[i] = 0;
    }
    // dy: gradient w.r.t. the raw output [B][n_out]; grad (same layout as theta) is overwritten
    void backward(int B, const BatchActs& A, const double* dy, std::vector<double>& grad) const {
        const int cat = MaceTopo::T + tp.n_char;
        grad.assign(theta.size(), 0.0);
        auto G = [&](int b) { return grad.data() + off[b]; };
        std::vector<double> dh((size_t)B * MaceTopo::H, 0.0);
        int col = 0;
        for (int hd = 0; hd < 4; ++hd) {
            const int nout = hd == 0 ? tp.n_frags : tp.frag;
            std::vector<double> h0((size_t)B * MaceTopo::HH), dh0((size_t)B * MaceTopo::HH), dhp((size_t)B * MaceTopo::H);
            for (int n = 0; n < B; ++n)
                std::copy(A.hh.begin() + ((size_t)n * 4 + hd) * MaceTopo::HH, A.hh.begin() + ((size_t)n * 4 + hd + 1) * MaceTopo::HH,
                          h0.begin() + (size_t)n * MaceTopo::HH);
            fc_bwd(B, dy + col, tp.n_out, nout, h0.data(), MaceTopo::HH, blob(12 + 4 * hd), G(12 + 4 * hd), G(13 + 4 * hd), dh0.data());
            relu_mask(dh0, h0);
            fc_bwd(B, dh0.data(), MaceTopo::HH, MaceTopo::HH, A.h.data(), MaceTopo::H, blob(10 + 4 * hd), G(10 + 4 * hd), G(11 + 4 * hd),
                   dhp.data());
            for (size_t i = 0; i < dh.size(); ++i) dh[i] += dhp[i];
            col += nout;
        }
        relu_mask(dh, A.h);
        std::vector<double> dcat((size_t)B * cat);
        fc_bwd(B, dh.data(), MaceTopo::H, MaceTopo::H, A.cat.data(), cat, blob(8), G(8), G(9), dcat.data());
        std::vector<double> dt((size_t)B * MaceTopo::T);
        for (int n = 0; n < B; ++n)
            for (int i = 0; i < MaceTopo::T; ++i) dt[(size_t)n * MaceTopo::T + i] = dcat[(size_t)n * cat + i];
        relu_mask(dt, A.t);
        std::vector<double> da2(A.a2.size()), da1(A.a1.size()), da0(A.a0.size());
        fc_bwd(B, dt.data(), MaceTopo::T, MaceTopo::T, A.a2.data(), MaceTopo::C2 * MaceTopo::W2, blob(6), G(6), G(7), da2.data());
        relu_mask(da2, A.a2);
        conv_bwd(B, da2.data(), MaceTopo::C2, MaceTopo::K2, A.a1.data(), MaceTopo::C1, MaceTopo::W1, blob(4), G(4), G(5), da1.data());
        relu_mask(da1, A.a1);
        conv_bwd(B, da1.data(), MaceTopo::C1, MaceTopo::K1, A.a0.data(), MaceTopo::C0, MaceTopo::W0, blob(2), G(2), G(3), da0.data());
        relu_mask(da0, A.a0);
        std::vector<double> terr((size_t)B * 200);
        for (int n = 0; n < B; ++n)
            std::copy(A.xn.begin() + (size_t)n * tp.n_in, A.xn.begin() + (size_t)n * tp.n_in + 200, terr.begin() + (size_t)n * 200);
        conv_bwd(B, da0.data(), MaceTopo::C0, MaceTopo::K0, terr.data(), 1, 200, blob(0), G(0), G(1), nullptr);
    }

    // cNeuralNet::EvalBatch: Y = net((X + off) * scale) / scale_out - off_out
    void eval_batch(int B, const double* X, std::vector<double>& Y, BatchActs* keep = nullptr) const {
        std::vector<double> xn((size_t)B * tp.n_in);
        for (int n = 0; n < B; ++n)
            for (int i = 0; i < tp.n_in; ++i) xn[(size_t)n * tp.n_in + i] = (X[(size_t)n * tp.n_in + i] + in_off[i]) * in_scale[i];
        BatchActs local;
        BatchActs& A = keep ? *keep : local;
        forward(B, xn.data(), A);
        Y.resize((size_t)B * tp.n_out);
        for (int n = 0; n < B; ++n)
            for (int i = 0; i < tp.n_out; ++i) Y[(size_t)n * tp.n_out + i] = A.y[(size_t)n * tp.n_out + i] / out_scale[i] - out_off[i];
    }
};

struct TrainerParams {
    int replay_cap = 500000, num_init_samples = 200, num_steps_per_iter = 1, freeze_target_iters = 0, batch = 32;
    int init_input_offset_scale = 1;
    double discount = 0.9, base_lr = 1e-3, momentum = 0.9, weight_decay = 5e-4;
    uint64_t seed = 1;
};

// cMACETrainer (pool size 1, synchronous mode, ENABLE_ACTOR_MULTI_SAMPLE_UPDATE)
struct MaceTrainer {
    enum { kFail = 1, kExpCritic = 2, kExpActor = 4 };
    TrainerParams P;
    MaceNet net, target;
    std::vector<double> history;
    int S = 0, A = 0, Wd = 0;
    std::vector<float> mem;                 // [cap][1 + S + A + S], float like the reference's Eigen::MatrixXf
    std::vector<int> flags;
    int head = 0, num = 0, iter = 0, actor_iter = 0, stage = 0;
    long long total = 0;
    std::vector<int> critic_buf, actor_buf, actor_batch;
    CounterRng rng;
    // pinning mode (tests/test_ref_pinning_cpu.py): draw the sample indices like the reference does -- cMathUtil::RandInt on its
    // process-global cRand (util/MathUtil.cpp, util/Rand.cpp:48-61), restated in terrain.h's Rand -- so that the compiled
    // cMACETrainer and this restatement can be compared draw for draw
    bool use_ref_rand = false;
    Rand ref_rand;
    Rand* shared_rand = nullptr;        // the reference has ONE engine for exploration and sampling: share the environments' one
    int draw(int n) { return shared_rand ? shared_rand->rand_int(0, n) : use_ref_rand ? ref_rand.rand_int(0, n) : rng.rand_int(0, n); }
    double last_critic_loss = 0, last_actor_loss = 0;
    std::vector<int> last_critic_ids, last_actor_ids;       // tuples of the most recent critic / actor solver step (for tests)

    void init(const Net& n, const TrainerParams& p) {
        P = p;
        net.init_from(n);
        target = net;
        history.assign(net.theta.size(), 0.0);
        S = n.n_in; A = 1 + n.frag; Wd = 1 + S + A + S;
        mem.assign((size_t)P.replay_cap * Wd, 0.f);
        flags.assign(P.replay_cap, 0);
        head = num = iter = actor_iter = stage = 0; total = 0;
        critic_buf.clear(); actor_buf.clear(); actor_batch.clear();
        rng.seed(P.seed, 0x7472616eull);
    }
    const float* row(int t) const { return mem.data() + (size_t)t * Wd; }
    bool exp_actor(int t) const { return (flags[t] & kExpActor) != 0; }

    static void remove_swap(std::vector<int>& v, int t) {
        auto it = std::find(v.begin(), v.end(), t);
        if (it != v.end()) { *it = v.back(); v.pop_back(); }
    }
    // cNeuralNetTrainer::AddTuple + cMACETrainer::SetTuple / UpdateBuffers (learning/MACETrainer.cpp:105-113,515-539,730-800)
    int add_tuple(const double* r, unsigned fl) {
        for (int i = 0; i < Wd; ++i) if (!std::isfinite(r[i])) return -1;     // CheckTuple
        const int t = head;
        float* dst = mem.data() + (size_t)t * Wd;
        for (int i = 0; i < Wd; ++i) dst[i] = (float)r[i];
        flags[t] = (int)fl;
        head = (head + 1) % P.replay_cap;
        num = std::min(P.replay_cap, num + 1);
        ++total;
        const bool ea = exp_actor(t);
        const bool in_actor = std::find(actor_buf.begin(), actor_buf.end(), t) != actor_buf.end();
        if (ea) { if (!in_actor) actor_buf.push_back(t); } else if (in_actor) remove_swap(actor_buf, t);
        const bool in_critic = std::find(critic_buf.begin(), critic_buf.end(), t) != critic_buf.end();
        if (!ea) { if (!in_critic) critic_buf.push_back(t); } else if (in_critic) remove_swap(critic_buf, t);
        remove_swap(actor_batch, t);
        return t;
    }

    double max_frag_val(const double* y) const {
        double m = y[0];
        for (int i = 1; i < net.tp.n_frags; ++i) m = std::max(m, y[i]);
        return m;
    }
    // CalcNewCumulativeRewardBatch (learning/MACETrainer.cpp:472-513); rows shorter than B are padded with row ids[0]
    void new_vals(const std::vector<int>& ids, std::vector<double>& out) {
        const int B = (int)ids.size();
        std::vector<double> X((size_t)B * S), Y;
        for (int i = 0; i < B; ++i)
            for (int j = 0; j < S; ++j) X[(size_t)i * S + j] = row(ids[i])[1 + S + A + j];
        target.eval_batch(B, X.data(), Y);
        const double norm = 1.0 - P.discount;
        out.resize(B);
        for (int i = 0; i < B; ++i) {
            double r = (double)row(ids[i])[0] * norm;
            out[i] = (flags[ids[i]] & kFail) ? r : r + P.discount * max_frag_val(&Y[(size_t)i * net.tp.n_out]);
        }
    }
    void curr_vals(const std::vector<int>& ids, std::vector<double>& out) {
        const int B = (int)ids.size();
        std::vector<double> X((size_t)B * S), Y;
        for (int i = 0; i < B; ++i)
            for (int j = 0; j < S; ++j) X[(size_t)i * S + j] = row(ids[i])[1 + j];
        target.eval_batch(B, X.data(), Y);
        out.resize(B);
        for (int i = 0; i < B; ++i) out[i] = max_frag_val(&Y[(size_t)i * net.tp.n_out]);
    }

    // cNeuralNet::Train on one batch: labels / data normalised as in LoadTrainData (learning/NeuralNet.cpp:1077-1117), one
    // SGD step.  X, Y are the un-normalised problem matrices.  Returns the Euclidean loss.
    double solver_step(const std::vector<double>& X, const std::vector<double>& Y) {
        const int B = P.batch, no = net.tp.n_out;
        std::vector<double> xn((size_t)B * S), lab((size_t)B * no), dy((size_t)B * no), grad;
        for (int n = 0; n < B; ++n) {
            for (int i = 0; i < S; ++i) xn[(size_t)n * S + i] = (X[(size_t)n * S + i] + net.in_off[i]) * net.in_scale[i];
            for (int i = 0; i < no; ++i) lab[(size_t)n * no + i] = (Y[(size_t)n * no + i] + net.out_off[i]) * net.out_scale[i];
        }
        BatchActs acts;
        net.forward(B, xn.data(), acts);
        double loss = 0;
        for (size_t i = 0; i < dy.size(); ++i) {
            double d = acts.y[i] - lab[i];
            loss += d * d;
            dy[i] = d / B;
        }
        loss /= 2.0 * B;
        net.backward(B, acts, dy.data(), grad);
        for (int b = 0; b < 26; ++b) {
            const double rate = P.base_lr * net.tp.lr_mult(b), decay = P.weight_decay * net.tp.decay_mult(b);
            for (size_t i = net.off[b]; i < net.off[b + 1]; ++i) {
                double g = grad[i] + decay * net.theta[i];
                history[i] = rate * g + P.momentum * history[i];
                net.theta[i] -= history[i];
            }
        }
        return loss;
    }

    bool critic_step() {
        const int B = P.batch, no = net.tp.n_out;
        if ((int)critic_buf.size() < B) return false;
        std::vector<int> ids(B);
        for (int i = 0; i < B; ++i) ids[i] = critic_buf[draw((int)critic_buf.size())];
        std::vector<double> X((size_t)B * S), Y, q;
        for (int i = 0; i < B; ++i)
            for (int j = 0; j < S; ++j) X[(size_t)i * S + j] = row(ids[i])[1 + j];
        new_vals(ids, q);
        net.eval_batch(B, X.data(), Y);
        for (int i = 0; i < B; ++i) {
            int a = (int)row(ids[i])[1 + S];
            Y[(size_t)i * no + a] = q[i];
        }
        last_critic_ids = ids;
        last_critic_loss = solver_step(X, Y);
        return true;
    }
    // UpdateActorBatchBuffer + UpdateActor (learning/MACETrainer.cpp:575-626)
    void actor_update() {
        const int B = P.batch, no = net.tp.n_out;
        {
            const int n_exp = (int)actor_buf.size();
            const int ns = std::min(B, n_exp);
            std::vector<int> cand;
            for (int i = 0; i < ns; ++i) {
                int t = actor_buf[draw(n_exp)];
                bool contains = std::find(actor_batch.begin(), actor_batch.end(), t) != actor_batch.end() ||
                                std::find(cand.begin(), cand.end(), t) != cand.end();
                if (!contains) cand.push_back(t);
            }
            if (!cand.empty()) {
                std::vector<double> v0, v1;
                curr_vals(cand, v0);
                new_vals(cand, v1);
                for (size_t i = 0; i < cand.size(); ++i) if (v1[i] > v0[i]) actor_batch.push_back(cand[i]);
            }
        }
        while ((int)actor_batch.size() >= B) {
            std::vector<double> X((size_t)B * S), Y;
            for (int i = 0; i < B; ++i)
                for (int j = 0; j < S; ++j) X[(size_t)i * S + j] = row(actor_batch[i])[1 + j];
            net.eval_batch(B, X.data(), Y);
            for (int i = 0; i < B; ++i) {
                const float* r = row(actor_batch[i]);
                int a = (int)r[1 + S];
                for (int k = 0; k < net.tp.frag; ++k) Y[(size_t)i * no + net.tp.n_frags + a * net.tp.frag + k] = r[1 + S + 1 + k];
            }
            last_actor_ids.assign(actor_batch.begin(), actor_batch.begin() + B);
            last_actor_loss = solver_step(X, Y);
            ++actor_iter;
            actor_batch.erase(actor_batch.begin(), actor_batch.begin() + B);
        }
    }
    // cNeuralNet::CalcOffsetScale over the replay memory (learning/NeuralNet.cpp:280-313, NeuralNetTrainer.cpp:696-719)
    static void calc_offset_scale(const double* X, int n, int S, double* off, double* scale) {
        std::vector<double> mean(S, 0.0), var(S, 0.0);
        const double norm = 1.0 / n;
        for (int t = 0; t < n; ++t)
            for (int j = 0; j < S; ++j) mean[j] += norm * X[(size_t)t * S + j];
        for (int t = 0; t < n; ++t)
            for (int j = 0; j < S; ++j) { double d = X[(size_t)t * S + j] - mean[j]; var[j] += norm * d * d; }
        for (int j = 0; j < S; ++j) {
            double sd = std::sqrt(var[j]);
            off[j] = -mean[j];
            scale[j] = sd == 0 ? 0 : 1.0 / sd;
        }
    }
    void update_offset_scale() {
        std::vector<double> X((size_t)num * S);
        for (int t = 0; t < num; ++t)
            for (int j = 0; j < S; ++j) X[(size_t)t * S + j] = (double)row(t)[1 + j];
        calc_offset_scale(X.data(), num, S, net.in_off.data(), net.in_scale.data());
        target.in_off = net.in_off; target.in_scale = net.in_scale;
    }
    // cNeuralNetTrainer::Train -> UpdateStage / ApplySteps / cMACETrainer::Step
    void train() {
        if (stage == 0) {
            int nis = std::min(P.num_init_samples, P.replay_cap);
            if (num >= nis && num > 0) {
                if (nis > 1 && P.init_input_offset_scale) update_offset_scale();
                stage = 1;
            }
        }
        if (stage != 1) return;
        bool succ = false;
        for (int s = 0; s < P.num_steps_per_iter; ++s) {
            succ = critic_step();
            actor_update();
            if (P.freeze_target_iters > 0 && iter > 0 && iter % P.freeze_target_iters == 0) {
                target.theta = net.theta;
                target.in_off = net.in_off; target.in_scale = net.in_scale; target.out_off = net.out_off; target.out_scale = net.out_scale;
            }
        }
        if (succ) ++iter;
    }
};

}  // namespace orc
