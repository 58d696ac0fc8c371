// TEST INFRASTRUCTURE ONLY -- the REFERENCE's own MACE trainer (learning/TrainerInterface.cpp, NeuralNetTrainer.cpp,
// MACETrainer.cpp, NeuralNetLearner.cpp, ExpTuple.cpp + util/MathUtil.cpp, Rand.cpp) compiled where it lies under /root/reference
// into oracle/_ref/libref_train.so, against the header stand-ins of oracle/ref_shim.  Everything the trainer does itself -- replay
// memory (float rows, ring buffer, flag buffer), critic / actor index buffers, minibatch sampling through cMathUtil::RandInt,
// target values (CalcNewCumulativeRewardBatch), the positive-temporal-difference filter of the actor batch buffer, stage
// handling, iteration counters, target-network refresh -- runs as compiled.  What it asks of cNeuralNet (Caffe in the
// reference: EvalBatch, Train on a problem, CopyModel, CalcOffsetScale, SetInputOffsetScale) is handed to callbacks; the test
// answers them with the network-level operations of oracle/trainer.h.  tests/test_ref_pinning_cpu.py then compares the compiled
// trainer with the oracle's restatement of it (MaceTrainer) run independently on the same tuples.
#include <cstdio>
#include <cstdlib>
#include <map>
#include <memory>
#include <vector>

#include "learning/MACETrainer.h"
#include "learning/NeuralNet.h"
#include "learning/NeuralNetLearner.h"

extern "C" void ref_abort_stub() {
    std::fprintf(stderr, "oracle/_ref: the compiled reference trainer called a Caffe-backed function that has no stand-in\n");
    std::abort();
}

// ------------------------------------------------------------------------------------------------ cNeuralNet over callbacks
typedef void (*eval_fn)(int net, const double* X, int B, double* Y, void* user);
typedef void (*train_fn)(int net, const double* X, const double* Y, int B, void* user);
typedef void (*copy_fn)(int dst, int src, void* user);
typedef void (*calc_os_fn)(const double* X, int n, double* off, double* scale, void* user);
typedef void (*set_os_fn)(int net, const double* off, const double* scale, void* user);
struct NetHooks {
    int n_in = 0, n_out = 0, batch = 32;
    eval_fn eval = nullptr;
    train_fn train = nullptr;
    copy_fn copy = nullptr;
    calc_os_fn calc_os = nullptr;
    set_os_fn set_os = nullptr;
    void* user = nullptr;
};
static NetHooks g_hooks;
static std::map<const cNeuralNet*, int> g_net_id;       // construction order: 0 trainer net, 1 its target net, 2 the learner's net
static int g_next_net = 0;
static int id_of(const cNeuralNet* n) { return g_net_id.at(n); }

static std::vector<double> flat(const Eigen::MatrixXd& M) {
    std::vector<double> v((size_t)M.rows() * M.cols());
    for (int i = 0; i < (int)M.rows(); ++i)
        for (int j = 0; j < (int)M.cols(); ++j) v[(size_t)i * M.cols() + j] = M(i, j);
    return v;
}

std::mutex cNeuralNet::gOutputLock;
cNeuralNet::tProblem::tProblem() : mPassesPerStep(1) {}
bool cNeuralNet::tProblem::HasData() const { return mX.size() > 0; }
cNeuralNet::cNeuralNet() : mValidModel(true), mAsync(false) { g_net_id[this] = g_next_net++; }
cNeuralNet::~cNeuralNet() { g_net_id.erase(this); }
void cNeuralNet::LoadNet(const std::string&) {}
void cNeuralNet::LoadModel(const std::string&) {}
void cNeuralNet::LoadSolver(const std::string&, bool) {}
void cNeuralNet::LoadScale(const std::string&) {}
void cNeuralNet::Clear() {}
void cNeuralNet::ResetSolver() {}
void cNeuralNet::OutputModel(const std::string&) const {}
bool cNeuralNet::HasNet() const { return true; }
bool cNeuralNet::HasSolver() const { return true; }
bool cNeuralNet::HasValidModel() const { return true; }
int cNeuralNet::GetInputSize() const { return g_hooks.n_in; }
int cNeuralNet::GetOutputSize() const { return g_hooks.n_out; }
int cNeuralNet::GetBatchSize() const { return g_hooks.batch; }
void cNeuralNet::Train(const tProblem& prob) {
    const std::vector<double> x = flat(prob.mX), y = flat(prob.mY);
    g_hooks.train(id_of(this), x.data(), y.data(), (int)prob.mX.rows(), g_hooks.user);
}
void cNeuralNet::EvalBatch(const Eigen::MatrixXd& X, Eigen::MatrixXd& out_Y) const {
    const int B = (int)X.rows();
    const std::vector<double> x = flat(X);
    std::vector<double> y((size_t)B * g_hooks.n_out);
    g_hooks.eval(id_of(this), x.data(), B, y.data(), g_hooks.user);
    out_Y.resize(B, g_hooks.n_out);
    for (int i = 0; i < B; ++i)
        for (int j = 0; j < g_hooks.n_out; ++j) out_Y(i, j) = y[(size_t)i * g_hooks.n_out + j];
}
void cNeuralNet::Eval(const Eigen::VectorXd& x, Eigen::VectorXd& out_y) const {
    std::vector<double> xi(x.size()), y(g_hooks.n_out);
    for (int i = 0; i < (int)x.size(); ++i) xi[i] = x[i];
    g_hooks.eval(id_of(this), xi.data(), 1, y.data(), g_hooks.user);
    out_y.resize(g_hooks.n_out);
    for (int j = 0; j < g_hooks.n_out; ++j) out_y[j] = y[j];
}
void cNeuralNet::CopyModel(const cNeuralNet& other) { g_hooks.copy(id_of(this), id_of(&other), g_hooks.user); }
void cNeuralNet::CalcOffsetScale(const Eigen::MatrixXd& X, Eigen::VectorXd& out_offset, Eigen::VectorXd& out_scale) const {
    const std::vector<double> x = flat(X);
    std::vector<double> off(X.cols()), sc(X.cols());
    g_hooks.calc_os(x.data(), (int)X.rows(), off.data(), sc.data(), g_hooks.user);
    out_offset.resize(X.cols()); out_scale.resize(X.cols());
    for (int j = 0; j < (int)X.cols(); ++j) { out_offset[j] = off[j]; out_scale[j] = sc[j]; }
}
void cNeuralNet::SetInputOffsetScale(const Eigen::VectorXd& offset, const Eigen::VectorXd& scale) {
    std::vector<double> off(offset.size()), sc(scale.size());
    for (int j = 0; j < (int)offset.size(); ++j) { off[j] = offset[j]; sc[j] = scale[j]; }
    g_hooks.set_os(id_of(this), off.data(), sc.data(), g_hooks.user);
}

// ------------------------------------------------------------------------------------------------ the compiled trainer, opened up
struct PinTrainer : public cMACETrainer {
    const std::vector<int>& critic() const { return mCriticBuffer; }
    const std::vector<int>& actor() const { return mActorBuffer; }
    const std::vector<int>& actor_batch() const { return mActorBatchBuffer; }
    int actor_iter() const { return mActorIter; }
    int stage() const { return mStage; }
    int num() const { return mNumTuples; }
    int head() const { return mBufferHead; }
    int total() const { return mTotalTuples; }
    int width() const { return (int)mPlaybackMem.cols(); }
    float mem(int t, int j) const { return mPlaybackMem(t, j); }
    unsigned flag(int t) const { return mFlagBuffer[t]; }
};
struct RefTrainer {
    std::shared_ptr<PinTrainer> tr;
    std::shared_ptr<cNeuralNetLearner> learner;
    std::unique_ptr<cNeuralNet> learner_net;
    int S = 0, A = 0;
};

extern "C" {

// p: n_in, n_out, batch, num_frags, frag_size, replay_cap, num_init_samples, num_steps_per_iter, freeze_target_iters, discount,
//    init_input_offset_scale, seed (cMathUtil::SeedRand: the engine behind cMathUtil::RandInt)
RefTrainer* ref_trainer_create(const double* p, eval_fn ev, train_fn tr, copy_fn cp, calc_os_fn cos, set_os_fn sos, void* user) {
    g_hooks.n_in = (int)p[0]; g_hooks.n_out = (int)p[1]; g_hooks.batch = (int)p[2];
    g_hooks.eval = ev; g_hooks.train = tr; g_hooks.copy = cp; g_hooks.calc_os = cos; g_hooks.set_os = sos; g_hooks.user = user;
    g_next_net = 0;
    g_net_id.clear();
    cMathUtil::SeedRand((unsigned long)p[11]);
    RefTrainer* r = new RefTrainer();
    r->tr = std::make_shared<PinTrainer>();
    r->tr->SetNumActionFrags((int)p[3]);
    r->tr->SetActionFragSize((int)p[4]);
    cTrainerInterface::tParams tp;                    // cScenarioTrain::InitTrainer fills these from the arg file
    tp.mNetFile = "stand-in"; tp.mSolverFile = "stand-in";
    tp.mPlaybackMemSize = (int)p[5];
    tp.mPoolSize = 1;
    tp.mNumInitSamples = (int)p[6];
    tp.mNumStepsPerIter = (int)p[7];
    tp.mFreezeTargetIters = (int)p[8];
    tp.mDiscount = p[9];
    tp.mInitInputOffsetScale = p[10] != 0;
    r->tr->Init(tp);
    r->S = (int)p[0]; r->A = 1 + (int)p[4];
    r->tr->RequestLearner(r->learner);                // cScenarioTrain::SetupLearner: one learner per exploration scenario
    r->learner_net.reset(new cNeuralNet());
    r->learner->SetNet(r->learner_net.get());
    return r;
}
void ref_trainer_destroy(RefTrainer* r) { delete r; }
// cNeuralNetLearner::Train(tuples): AddTuples + Train + SyncNet, what cScenarioTrain::UpdateExpScene does with a full tuple buffer
// rows: n x (reward | state_beg | action | state_end)
void ref_trainer_learn(RefTrainer* r, const double* rows, const unsigned* flags, int n) {
    const int W = 1 + r->S + r->A + r->S;
    std::vector<tExpTuple> tuples;
    for (int i = 0; i < n; ++i) {
        const double* row = rows + (size_t)i * W;
        tExpTuple t(r->S, r->A);
        t.mReward = row[0];
        for (int j = 0; j < r->S; ++j) { t.mStateBeg[j] = row[1 + j]; t.mStateEnd[j] = row[1 + r->S + r->A + j]; }
        for (int j = 0; j < r->A; ++j) t.mAction[j] = row[1 + r->S + j];
        t.mFlags = flags[i];
        tuples.push_back(t);
    }
    r->learner->Train(tuples);
}
// iter, actor_iter, stage, num, head, total, learner iter, learner tuples
void ref_trainer_counters(RefTrainer* r, long* c) {
    c[0] = r->tr->GetIter(); c[1] = r->tr->actor_iter(); c[2] = r->tr->stage(); c[3] = r->tr->num(); c[4] = r->tr->head();
    c[5] = r->tr->total(); c[6] = r->learner->GetIter(); c[7] = r->learner->GetNumTuples();
}
// which 0: critic buffer, 1: actor buffer, 2: actor batch buffer
int ref_trainer_list(RefTrainer* r, int which, int* out, int cap) {
    const std::vector<int>& v = which == 0 ? r->tr->critic() : which == 1 ? r->tr->actor() : r->tr->actor_batch();
    for (int i = 0; i < (int)v.size() && i < cap; ++i) out[i] = v[i];
    return (int)v.size();
}
int ref_trainer_width(RefTrainer* r) { return r->tr->width(); }
void ref_trainer_rows(RefTrainer* r, const int* ids, int n, float* rows, int* flags) {
    const int W = r->tr->width();
    for (int i = 0; i < n; ++i) {
        for (int j = 0; j < W; ++j) rows[(size_t)i * W + j] = r->tr->mem(ids[i], j);
        flags[i] = (int)r->tr->flag(ids[i]);
    }
}

}  // extern "C"
