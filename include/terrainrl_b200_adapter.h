/* terrainrl_b200_adapter.h -- C++ adapter between the reference's scenario classes and the C ABI (terrainrl_b200.h).
 *
 * Compiles against the reference's own headers (scenarios/ScenarioExpMACE.h and what it includes); it is the binding a
 * maintainer adds to TerrainRL_Optimizer so that cScenarioTrain / cScenarioTrainMACE and the trainers behind them link
 * unchanged: cScenarioTrain::BuildExpScene (scenarios/ScenarioTrain.cpp:224-237) returns a cScenarioExpBatched instead of a
 * cScenarioExpMACE, everything else -- BuildScenePool, SetupLearner, UpdateExpScene, the annealing schedule, the learner
 * (scenarios/ScenarioTrain.cpp:197-222,277-282,376-460) -- runs as compiled and sees ONE pooled scene that happens to hold
 * num_envs environments stepped in lock-step on a GPU.
 *
 * What the caller of an exploration scene touches and where it goes:
 *   ParseArgs, Init                         Base (the reference's own code builds the character + controller it later asks for
 *                                           GetNet / BuildNNOutputOffsetScale / GetNumActionFrags; that world is never stepped)
 *                                           + trl_create_from_pack
 *   Reset                                   trl_reset (all environments)
 *   Update(dt)                              trl_update
 *   IsTupleBufferFull / GetTuples / ResetTupleBuffer
 *                                           trl_num_tuples / trl_get_tuples_f64 (materialised as tExpTuple, learning/ExpTuple.h:5-22)
 *                                           / trl_reset_tuples
 *   EnableExplore, SetExpRate/Temp/BaseActionRate
 *                                           Base (keeps the getters cScenarioTrain reads back consistent) + trl_set_explore
 *   SetTerrainParamsLerp                    trl_set_terrain_lerp
 *   weights after cNeuralNetLearner::SyncNet (learning/NeuralNetLearner.cpp:48-60)
 *                                           PushWeights -> trl_set_weights
 *
 * Base is cScenarioExpMACE in a deployment; the test harness of this repository (oracle/ref_ctrl_api.cpp, which compiles the
 * reference's scenario classes against header stand-ins because Bullet / Caffe are absent) passes its own subclass.
 * Errors follow the reference's convention for this layer: printf + assert (scenarios/ScenarioTrain.cpp:282).
 */
#pragma once
#include <cassert>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "learning/ExpTuple.h"
extern "C" {
#include "terrainrl_b200.h"
}

template <class Base>
class cScenarioExpBatchedT : public Base {
public:
    cScenarioExpBatchedT() {}
    virtual ~cScenarioExpBatchedT() {
        if (mHandle) trl_destroy(mHandle);
    }

    // before Init: the scene pack made from the same arg file (trl_pack_from_args), the batch size and the GPU
    void SetBatch(const std::string& pack, int num_envs, int device, uint64_t rng_seed = 1234, const uint64_t* terrain_seeds = nullptr) {
        mPack = pack; mNumEnvs = num_envs; mDevice = device; mRngSeed = rng_seed;
        mSeeds.clear();
        if (terrain_seeds) mSeeds.assign(terrain_seeds, terrain_seeds + num_envs);
    }
    trl_handle* GetHandle() const { return mHandle; }
    int GetNumEnvs() const { return mNumEnvs; }

    virtual void Init() {
        Base::Init();
        mHandle = trl_create_from_pack(mPack.c_str(), mNumEnvs, mDevice, TRL_MODE_EXPLORE, mSeeds.empty() ? nullptr : mSeeds.data(), mRngSeed);
        if (!mHandle) { printf("cScenarioExpBatched: %s\n", trl_last_error()); assert(false); }
        int n = 0, nd = 0, nj = 0, nf = 0, fs = 0;
        Check(trl_sizes(mHandle, &n, &mStateSize, &mActionSize, &nf, &fs, &nd, &nj));
        PushExplore();
    }
    // Base::Init reaches several of these virtuals before the batch exists (scenarios/ScenarioExp.cpp:42-53): until then they
    // are the reference's own
    virtual void Reset() {
        if (!mHandle) { Base::Reset(); return; }
        Check(trl_reset(mHandle, nullptr, 0));
    }
    virtual void Clear() {
        if (mHandle) { trl_destroy(mHandle); mHandle = nullptr; }
        Base::Clear();
    }
    virtual void Update(double time_elapsed) {
        if (!mHandle) { Base::Update(time_elapsed); return; }
        Check(trl_update(mHandle, time_elapsed));
    }

    virtual bool IsTupleBufferFull() const {
        if (!mHandle) return Base::IsTupleBufferFull();
        int n = 0;
        Check(trl_num_tuples(mHandle, &n));
        return n >= this->mTupleBufferSize;
    }
    virtual const std::vector<tExpTuple>& GetTuples() const {
        if (!mHandle) return Base::GetTuples();
        const double* rows = nullptr; const uint32_t* flags = nullptr; const int32_t* env = nullptr;
        int n = 0;
        Check(trl_get_tuples_f64(mHandle, &rows, &flags, &env, &n));
        const int W = 1 + mStateSize + mActionSize + mStateSize;
        mTuples.assign(n, tExpTuple(mStateSize, mActionSize));
        for (int i = 0; i < n; ++i) {
            const double* r = rows + (size_t)i * W;
            tExpTuple& t = mTuples[i];
            t.mID = env[i];
            t.mReward = r[0];
            t.mFlags = flags[i];
            for (int k = 0; k < mStateSize; ++k) { t.mStateBeg[k] = r[1 + k]; t.mStateEnd[k] = r[1 + mStateSize + mActionSize + k]; }
            for (int k = 0; k < mActionSize; ++k) t.mAction[k] = r[1 + mStateSize + k];
        }
        return mTuples;
    }
    virtual void ResetTupleBuffer() {
        if (!mHandle) { Base::ResetTupleBuffer(); return; }
        Check(trl_reset_tuples(mHandle));
    }

    virtual void EnableExplore(bool enable) { Base::EnableExplore(enable); PushExplore(); }
    virtual void SetExpRate(double rate) { Base::SetExpRate(rate); PushExplore(); }
    virtual void SetExpTemp(double temp) { Base::SetExpTemp(temp); PushExplore(); }
    virtual void SetExpBaseActionRate(double rate) { Base::SetExpBaseActionRate(rate); PushExplore(); }
    virtual void SetTerrainParamsLerp(double lerp) { if (mHandle) Check(trl_set_terrain_lerp(mHandle, lerp)); }

    // cNeuralNetLearner::SyncNet has copied the trainer's net into the controller's: hand the same parameters to the batch.
    // blobs / counts: the 26 parameter blobs in net order (weights, bias per layer), then the four offset / scale vectors.
    void PushWeights(const double* const* blobs, const int64_t* counts, int nblobs, const double* in_off, const double* in_scale,
                     const double* out_off, const double* out_scale) {
        Check(trl_set_weights(mHandle, blobs, counts, nblobs, in_off, in_scale, out_off, out_scale));
    }

protected:
    void Check(int rc) const {
        if (rc != 0) { printf("cScenarioExpBatched: %s\n", trl_last_error()); assert(false); }
    }
    void PushExplore() {
        if (mHandle) Check(trl_set_explore(mHandle, this->mEnableExplore ? 1 : 0, this->mExpRate, this->mExpTemp, this->mExpBaseActionRate));
    }

    trl_handle* mHandle = nullptr;
    std::string mPack;
    int mNumEnvs = 1, mDevice = 0;
    uint64_t mRngSeed = 1234;
    std::vector<uint64_t> mSeeds;
    int mStateSize = 0, mActionSize = 0;
    mutable std::vector<tExpTuple> mTuples;
};

// Policy evaluation: what cOptScenarioPoliEval::{BuildScenePool, EvalHelper, OutputResults} call on a pooled evaluation scene
// (optimizer/scenarios/OptScenarioPoliEval.cpp:135-163,170-198,213-237): ParseArgs, Init, SetRandSeed, Reset, Update,
// GetNumCycles, GetNumEpisodes, GetAvgDist, ResetAvgDist, GetDistLog.  Base = cScenarioPoliEval in a deployment.  The counters
// are those of the whole batch (sums; the mean distance is episode-weighted, as OutputResults merges its pool).
template <class Base>
class cScenarioPoliEvalBatchedT : public Base {
public:
    cScenarioPoliEvalBatchedT() {}
    virtual ~cScenarioPoliEvalBatchedT() {
        if (mHandle) trl_destroy(mHandle);
    }
    void SetBatch(const std::string& pack, int num_envs, int device, uint64_t rng_seed = 1234) {
        mPack = pack; mNumEnvs = num_envs; mDevice = device; mRngSeed = rng_seed;
    }
    trl_handle* GetHandle() const { return mHandle; }

    virtual void Init() {
        Base::Init();
        mHandle = trl_create_from_pack(mPack.c_str(), mNumEnvs, mDevice, TRL_MODE_POLI_EVAL, nullptr, mRngSeed);
        if (!mHandle) { printf("cScenarioPoliEvalBatched: %s\n", trl_last_error()); assert(false); }
    }
    virtual void Clear() {
        if (mHandle) { trl_destroy(mHandle); mHandle = nullptr; }
        Base::Clear();
    }
    // one seed per pooled scene in the reference (OptScenarioPoliEval.cpp:150-160); a batch derives its environments' terrain
    // seeds from it: seed, seed + 1, ...  (followed by the Reset the reference issues right after)
    virtual void SetRandSeed(unsigned long seed) {
        if (!mHandle) { Base::SetRandSeed(seed); return; }
        std::vector<uint64_t> seeds(mNumEnvs);
        for (int i = 0; i < mNumEnvs; ++i) seeds[i] = (uint64_t)seed + (uint64_t)i;
        Check(trl_seed_terrain(mHandle, seeds.data(), mNumEnvs));
    }
    virtual void Reset() {
        if (!mHandle) { Base::Reset(); return; }
        Check(trl_reset(mHandle, nullptr, 0));
    }
    virtual void Update(double time_elapsed) {
        if (!mHandle) { Base::Update(time_elapsed); return; }
        Check(trl_update(mHandle, time_elapsed));
    }
    virtual double GetAvgDist() const {
        if (!mHandle) return Base::GetAvgDist();
        double avg = 0;
        Check(trl_eval_stats(mHandle, nullptr, nullptr, &avg, nullptr));
        return avg;
    }
    virtual void ResetAvgDist() {
        if (!mHandle) { Base::ResetAvgDist(); return; }
        Check(trl_reset_avg_dist(mHandle));
    }
    virtual int GetNumEpisodes() const {
        if (!mHandle) return Base::GetNumEpisodes();
        int64_t e = 0;
        Check(trl_eval_stats(mHandle, nullptr, &e, nullptr, nullptr));
        return (int)e;
    }
    virtual int GetNumCycles() const {
        if (!mHandle) return Base::GetNumCycles();
        int64_t c = 0;
        Check(trl_eval_stats(mHandle, &c, nullptr, nullptr, nullptr));
        return (int)c;
    }
    virtual const std::vector<double>& GetDistLog() const {
        if (!mHandle) return Base::GetDistLog();
        const double* d = nullptr; const int32_t* env = nullptr;
        int n = 0;
        Check(trl_dist_log(mHandle, &d, &env, &n));
        mLog.assign(d, d + n);
        return mLog;
    }

protected:
    void Check(int rc) const {
        if (rc != 0) { printf("cScenarioPoliEvalBatched: %s\n", trl_last_error()); assert(false); }
    }
    trl_handle* mHandle = nullptr;
    std::string mPack;
    int mNumEnvs = 1, mDevice = 0;
    uint64_t mRngSeed = 1234;
    mutable std::vector<double> mLog;
};

#ifdef TRL_ADAPTER_WITH_REFERENCE_SCENARIO
#include "scenarios/ScenarioExpMACE.h"
#include "scenarios/ScenarioPoliEval.h"
typedef cScenarioExpBatchedT<cScenarioExpMACE> cScenarioExpBatched;
typedef cScenarioPoliEvalBatchedT<cScenarioPoliEval> cScenarioPoliEvalBatched;
#endif
