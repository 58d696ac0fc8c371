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

#include "ref_net_standin.h"

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
