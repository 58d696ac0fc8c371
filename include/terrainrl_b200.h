/* terrainrl_b200 -- C ABI of the B200-native batched rollout engine.
 *
 * One handle = one batch of N independent environments stepped in lock-step on one GPU.  Each entry point names
 * the reference interface it replaces (file:line under the DeepTerrainRL tree); the reference has no FFI layer,
 * its seam is the C++ virtual scenario API that cScenarioTrain / cOptScenarioPoliEval call on every pooled env
 * (scenarios/ScenarioTrain.cpp:197-222,376-410; optimizer/scenarios/OptScenarioPoliEval.cpp:135-237).
 *
 * Conventions: all functions return 0 on success, non-zero on error (trl_last_error() gives the message; the
 * reference's printf+assert convention cannot cross an ABI).  A handle is single-writer.  Pointers returned by
 * the library stay valid until the next call that mutates the same data (documented per function).
 * Plain C types only: no torch / CUDA types in any signature.
 */
#ifndef TERRAINRL_B200_H
#define TERRAINRL_B200_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct trl_handle trl_handle;

enum { TRL_MODE_POLI_EVAL = 0, TRL_MODE_EXPLORE = 1 };

/* cScenario{PoliEval,ExpMACE}::ParseArgs + Init + SetRandSeed + Reset for N envs
 * (scenarios/ScenarioSimChar.cpp:80-119, ScenarioPoliEval.cpp:60-108,153-160, ScenarioExp.cpp:29-73).
 * `pack_path` is a scene pack produced by tools/pack_scene.py from a reference arg file.  terrain_seeds may be NULL
 * (env i gets seed 1+i).  Returns NULL on failure. */
trl_handle* trl_create_from_pack(const char* pack_path, int num_envs, int device, int mode,
                                 const uint64_t* terrain_seeds, uint64_t rng_seed);
/* Same, straight from the reference's own inputs: argv holds cArgParser-style tokens (e.g. "-arg_file=",
 * "args/dog_slopes_mixed_args.txt", plus overrides, CLI first = CLI wins, optimizer/Main.cpp:19-32); relative paths
 * inside are resolved against data_root (a DeepTerrainRL checkout).  Native readers for the arg file, the JSON assets
 * and the Caffe HDF5 weights (util/ArgParser.cpp:42-140, anim/KinTree.cpp:9-60, learning/NeuralNet.cpp:81-215). */
trl_handle* trl_create(int argc, const char* const* argv, const char* data_root, int num_envs, int device, int mode,
                       const uint64_t* terrain_seeds, uint64_t rng_seed);
/* Scene files -> .trlpack without touching the GPU (what tools/pack_scene.py does in Python) */
int trl_pack_from_args(int argc, const char* const* argv, const char* data_root, const char* out_path);
int trl_destroy(trl_handle* h);

/* cScenario::Reset for the listed envs (NULL = all)            scenarios/ScenarioSimChar.cpp:121-132 */
int trl_reset(trl_handle* h, const int32_t* env_ids, int n);
/* cScenarioPoliEval::SetRandSeed + Reset for every env         scenarios/ScenarioPoliEval.cpp:153-160 */
int trl_seed_terrain(trl_handle* h, const uint64_t* seeds, int n);

/* cScenario::Update(dt) for all envs in lock-step: num_update_steps env-steps, then the fall -> reset handling
 * (scenarios/ScenarioSimChar.cpp:141-182, ScenarioPoliEval.cpp:110-125, ScenarioExp.cpp:83-98).
 * Asynchronous: enqueued on the handle's stream; trl_sync() or any getter waits. */
int trl_update(trl_handle* h, double dt);
/* one iteration of the loop at scenarios/ScenarioSimChar.cpp:162-173 (parity probe; no fall handling) */
int trl_env_step(trl_handle* h, double step);
int trl_sync(trl_handle* h);

/* cScenarioExp::{EnableExplore,SetExpRate,SetExpTemp,SetExpBaseActionRate}   scenarios/ScenarioExp.cpp:156-185 */
int trl_set_explore(trl_handle* h, int enable, double rate, double temp, double base_rate);
/* engine contact / joint-limit parameters (7 doubles: kn, dn, mu, v_eps, contact_tol, k_lim, d_lim) */
int trl_set_phys_params(trl_handle* h, const double* p7);

/* cNeuralNet::CopyModel into every env's controller net (learning/NeuralNetLearner.cpp:85-89): 26 blobs in Caffe
 * order (w, b of terr_conv0..2, terr_ip0, ip0, val_ip0, val_ip1, a{0,1,2}_ip{0,1}) + the four offset/scale vectors. */
int trl_set_weights(trl_handle* h, const double* const* blobs, const int64_t* counts, int nblobs,
                    const double* in_off, const double* in_scale, const double* out_off, const double* out_scale);

/* sizes: policy state, action record, MACE fragments           sim/NNController.cpp:49-78, BaseControllerMACE.cpp:36-56 */
int trl_sizes(trl_handle* h, int* num_envs, int* state, int* action, int* num_frags, int* frag_size, int* num_dof,
              int* num_joints);

/* cScenarioExp::{IsTupleBufferFull,GetTuples,ResetTupleBuffer} for the whole batch: rows are
 * [reward | s(S) | a(A) | s'(S)] exactly like cMACETrainer's replay rows (learning/MACETrainer.cpp:515-539);
 * flags bit0 fail, bit1 expCritic, bit2 expActor (learning/MACETrainer.h:11-17).  The f32 view is what the
 * trainer stores; the f64 view is what tExpTuple holds.  Pointers valid until trl_reset_tuples / trl_update. */
int trl_num_tuples(trl_handle* h, int* out);
int trl_get_tuples(trl_handle* h, const float** rows, const uint32_t** flags, const int32_t** env_id, int* n);
int trl_get_tuples_f64(trl_handle* h, const double** rows, const uint32_t** flags, const int32_t** env_id, int* n);
int trl_reset_tuples(trl_handle* h);

/* cScenarioPoliEval::{GetNumCycles,GetNumEpisodes,GetAvgDist,GetDistLog}   scenarios/ScenarioPoliEval.cpp:127-151 */
int trl_eval_stats(trl_handle* h, int64_t* cycles, int64_t* episodes, double* avg_dist, int64_t* env_steps);
int trl_dist_log(trl_handle* h, const double** dist, const int32_t** env_id, int* n);
/* cScenarioPoliEval::ResetAvgDist for every env                scenarios/ScenarioPoliEval.cpp:132-136 */
int trl_reset_avg_dist(trl_handle* h);

/* cSimCharacter::{BuildPose,BuildVel} / SetPose+SetVel and contact bits for one env (sim/SimCharacter.cpp:166-315) */
int trl_get_state(trl_handle* h, int env, double* pose, double* vel, double* held_torque, uint8_t* contact);
int trl_set_state(trl_handle* h, int env, const double* pose, const double* vel, const double* held_torque,
                  const uint8_t* contact);
/* bulk SoA views: pose/vel as [num_dof][num_envs] doubles */
int trl_get_state_all(trl_handle* h, double* pose, double* vel);

/* parity probes (same layouts as the oracle's probes) */
int trl_get_ctrl(trl_handle* h, int env, double* out, int cap, int* n);
int trl_get_poli_state(trl_handle* h, int env, double* out);
int trl_get_net_out(trl_handle* h, int env, double* out);
/* cNeuralNet::GetLayerState (learning/NeuralNet.cpp:814-833) for the policy state of env's last decision: the blob `layer_name` of the
 * deploy net (any top name of data/policies/<char>/nets/<char>_mace3_deploy.prototxt, e.g. "terr_conv1", "relu0", "a1_ip0", "output");
 * what cScenarioPoliEval::RecordNNActivation writes per cycle (scenarios/ScenarioPoliEval.cpp:271-296).  *n = the blob's size. */
int trl_get_layer_state(trl_handle* h, int env, const char* layer_name, double* out, int cap, int* n);
int trl_get_terrain(trl_handle* h, int env, int seg, float* data, int cap, int* n, double* min_x, int* flip);

/* number of engine kernels launched since creation (bench.py's gpu_launches claim) */
int64_t trl_kernel_launches(trl_handle* h);

/* pipelined read-back of what cOptScenarioPoliEval::EvalHelper reads after every Update (counters; optionally every
 * env's pose/vel as [num_dof][num_envs] planes): trl_snapshot() enqueues the capture behind the updates already
 * submitted, trl_snapshot_wait() returns it from pinned host memory -- call trl_update() for the next step in between. */
int trl_snapshot(trl_handle* h);
int trl_snapshot_wait(trl_handle* h, double* pose, double* vel, int64_t* cycles, int64_t* episodes, double* avg_dist,
                      int64_t* env_steps);

/* device-resident view of the tuple block (rows f64 [cap][width], flags, env ids, count) for zero-copy hand-off to
 * NCCL; replaces the per-thread `learner->Train(exp->GetTuples())` hand-off (scenarios/ScenarioTrain.cpp:388-395) */
int trl_device_tuple_block(trl_handle* h, void** rows_f64, void** flags_u32, void** env_i32, void** count_i32,
                           int* cap, int* width);

/* measurement helpers: K outer updates timed with CUDA events on the handle's stream, one event pair per update, *ms_total = the sum
 * (flush_l2: a 256 MiB memset between the pairs evicts L2 before every update; it is a measurement device and is not timed --
 * trl_bench_last_span returns the span from the first update's start to the last update's end with the flushes in it);
 * one update with an event pair around every kernel launch (per-kernel device time for the roofline). */
int trl_bench_updates(trl_handle* h, double dt, int k, int flush_l2, double* ms_total);
int trl_bench_last_span(trl_handle* h, double* ms);
int trl_update_timed(trl_handle* h, double dt, double* step_ms, int* step_launches, double* decide_ms, int* decide_launches);
int trl_update_timed_detail(trl_handle* h, double dt, double* per_step_ms, double* per_decide_ms);
/* one update in trl_update's own (overlapped) schedule with an event pair around every launch: out[4k..4k+3] = {kind (0 terrain, 1 step,
 * 2 decision, 3 catch-up), index, start ms, end ms} */
int trl_update_timeline(trl_handle* h, double dt, double* out4, int cap, int* n);

int trl_debug_time_decide(trl_handle* h, int n_pending, int iters, double* ms_avg);
/* measurement only: in-kernel phase stamps (ns from entry, 16 slots) of the batched decision path's FC kernel + event-timed conv / FC launches */
int trl_debug_fc_phases(trl_handle* h, int n_pending, double* out_ns16, double* conv_us, double* fc_us);

/* ---------------------------------------------------------------------------------------------------------------------
 * MACE trainer on the GPU -- replaces cMACETrainer / cNeuralNetLearner behind cScenarioTrainMACE:
 *   trl_trainer_create        cTrainerInterface::Init + LoadModel (learning/NeuralNetTrainer.cpp:26-49, MACETrainer.cpp:78-90);
 *                             p[10] = {replay_mem_size, num_init_samples, num_steps_per_iter, freeze_target_iters,
 *                             init_input_offset_scale, discount, base_lr, momentum, weight_decay, seed}.  From this call on the
 *                             scenario's decision kernel evaluates the trainer's current net (cNeuralNetLearner::SyncNet,
 *                             learning/NeuralNetLearner.cpp:83-87, becomes a pointer binding).
 *   trl_trainer_add_from_scene  AddTuples(exp.GetTuples()) + ResetTupleBuffer (learning/NeuralNetLearner.cpp:33-45,
 *                             scenarios/ScenarioTrain.cpp:388-408), device to device
 *   trl_trainer_add_tuples    cNeuralNetTrainer::AddTuples from host rows [n][1 + S + A + S] (learning/NeuralNetTrainer.cpp:145-173)
 *   trl_trainer_train         iters x cNeuralNetTrainer::Train (learning/NeuralNetTrainer.cpp:175-183 -> cMACETrainer::Step,
 *                             learning/MACETrainer.cpp:335-361), asynchronous on the scenario's stream
 *   trl_trainer_counters      GetIter / GetNumTuples / buffer sizes / last losses
 *   trl_trainer_get, _set_theta, _list   model read-back (OutputModel), LoadModel, buffer inspection for the tests */
/* model files in the reference's formats, natively (no Python, no libhdf5):
 *   trl_load_model     cNeuralNet::LoadModel + LoadScale (learning/NeuralNet.cpp:157-186): Caffe ToHDF5 weights + `_scale.txt`
 *   trl_output_model   cNeuralNet::OutputModel + WriteOffsetScale (learning/NeuralNet.cpp:571-587,1182-1205) of the policy the
 *                      scenario currently evaluates (the attached trainer's net if there is one); mtime = HDF5 modification stamp
 *   trl_write_model    the same for host-side weights: blobs[26] = (weights, bias) of the 13 parameter layers in net order
 *   trl_get_output_offset_scale   cBaseControllerMACE::BuildNNOutputOffsetScale (sim/BaseControllerMACE.cpp:75-113) */
int trl_load_model(trl_handle* h, const char* h5_path, const char* scale_json_path);
int trl_output_model(trl_handle* h, const char* path, uint32_t mtime);
int trl_write_model(const char* path, const double* const* blobs26, int n_char, int n_frags, int frag_size, const double* in_off,
                    const double* in_scale, const double* out_off, const double* out_scale, uint32_t mtime);
int trl_get_output_offset_scale(trl_handle* h, double* offset, double* scale, int n_out);
int trl_pack_output_offset_scale(const char* pack_path, double* offset, double* scale, int n_out);   /* host only, no device */

/* cScenarioSimChar::SetTerrainParamsLerp (scenarios/ScenarioSimChar.cpp:255-272) and cScenarioTrain's annealing schedule
 * (scenarios/ScenarioTrain.cpp:412-460): sp[9] = {init_exp_rate, exp_rate, init_exp_temp, exp_temp, init_exp_base_rate,
 * exp_base_rate, trainer_num_anneal_iters, exp_base_anneal_iters, trainer_curriculum_iters};
 * out[4] = {exp_rate, exp_temp, exp_base_rate, curriculum_phase} */
int trl_set_terrain_lerp(trl_handle* h, double lerp);
int trl_train_schedule(const double* sp9, int iters, double* out4);

typedef struct trl_trainer trl_trainer;
trl_trainer* trl_trainer_create(trl_handle* h, const double* params10);
int trl_trainer_destroy(trl_trainer* t);
/* training from scratch: xavier-filled weights + cBaseControllerMACE::BuildNNOutputOffsetScale (sim/BaseControllerMACE.cpp:75-113),
 * what cScenarioTrain::InitTrainer leaves when no -policy_model is given (scenarios/ScenarioTrain.cpp:253-259,322-336) */
int trl_trainer_init_fresh(trl_trainer* t, uint64_t seed);
int trl_trainer_add_from_scene(trl_trainer* t);
int trl_trainer_add_tuples(trl_trainer* t, const double* rows, const uint32_t* flags, int n);
/* same from device memory (the all-gathered tuple blocks of every rank); ordered on the scenario's stream */
int trl_trainer_add_device(trl_trainer* t, const double* rows_dev, const uint32_t* flags_dev, int n);
int trl_trainer_train(trl_trainer* t, int iters);
/* cScenarioTrain::Run for one batch (scenarios/ScenarioTrain.cpp:100-115,376-410): num_updates x {update, tuple hand-over,
 * trainer iterations (iters_per_update, or 0 = one per tuple_buffer_size new tuples), annealed exploration + curriculum} */
int trl_train_run(trl_trainer* t, const double* sp9, int num_updates, int iters_per_update, int tuple_buffer_size, double time_step);
/* the same loop with the tuple exchange of a multi-GPU run (when the scenario has a communicator: trl_gather_tuples(block_rows) +
 * trl_trainer_add_gathered instead of the local hand-over), timed with CUDA events on the scenario's stream (flush_l2: a 256 MiB
 * memset before every update, as trl_bench_updates does); *iters_state carries the annealing position across calls (start at 0) */
int trl_train_run_timed(trl_trainer* t, const double* sp9, int num_updates, int iters_per_update, int block_rows, double time_step,
                        int flush_l2, int64_t* iters_state, double* ms);
/* asynchronous training (the reference's asynchronous trainer: exploration keeps running on the net it has while the trainer works):
 * hand-over + trainer iterations run on their own stream beside the NEXT update inside trl_train_run / trl_train_run_timed; the
 * decision kernels read a snapshot of the net refreshed between updates (tuples of update u shape the policy from update u + 2 on).
 * Needs a communicator (a single rank will do) and a fixed iters_per_update. */
int trl_trainer_set_async(trl_trainer* t, int enable);
int trl_trainer_counters(trl_trainer* t, int64_t* counters9, double* losses2);
int trl_trainer_num_params(trl_trainer* t);
int64_t trl_trainer_launches(trl_trainer* t);
int trl_trainer_get(trl_trainer* t, int what, double* out);
int trl_trainer_set_theta(trl_trainer* t, const double* theta);
int trl_trainer_list(trl_trainer* t, int which, int32_t* out, int cap, int* len);
int trl_trainer_rows(trl_trainer* t, const int32_t* slots, int n, float* rows, int32_t* flags);

/* number of tuples the engine could not record because the tuple block was full (cScenarioExp's ring overwrites, scenarios/
 * ScenarioExp.cpp:296-301; a batch block refuses instead and counts).  trl_num_tuples / trl_get_tuples* return
 * TRL_E_TUPLE_OVERFLOW (rows up to the capacity are still delivered) once this is non-zero since the last trl_reset_tuples. */
enum { TRL_E_TUPLE_OVERFLOW = 2 };
int trl_tuples_dropped(trl_handle* h, int64_t* out);

/* ---------------------------------------------------------------------------------------------------------------------
 * Multi-GPU exchange: one process per GPU, environments sharded by rank, no collective in the rollout.
 * The two couplings of the reference's training loop become two collectives on a side stream of the handle:
 *   trl_gather_tuples      every rank's finished tuples of this update -> every rank; replaces `learner->Train(exp->GetTuples())`
 *                          under the trainer mutex (scenarios/ScenarioTrain.cpp:388-395, learning/NeuralNetLearner.cpp:33-46):
 *                          a device pack kernel (f32 rows like cMACETrainer's replay memory, learning/MACETrainer.cpp:515-539, flags,
 *                          global env ids, count header) + ONE all-gather of a fixed-capacity block per rank.  No host
 *                          synchronisation; rows beyond block_rows stay queued in the tuple block for the next call.
 *   trl_trainer_add_gathered   AddTuples of all gathered blocks (rank order, env order inside a rank) into the attached trainer
 *   trl_trainer_broadcast  cNeuralNetLearner::SyncNet across ranks (learning/NeuralNetLearner.cpp:85-89): weights + offset/scale
 *                          (and the trainer's target net / momentum history) from `root` to every rank
 *   trl_comm_eval_stats    cOptScenarioPoliEval::OutputResults' merge of the per-thread results
 *                          (optimizer/scenarios/OptScenarioPoliEval.cpp:213-237) as one all-reduce
 * Backends: NCCL (libnccl.so.2 is opened at run time by trl_comm_init; the library does not link against it), or caller-supplied
 * collectives (trl_comm_init_external: MPI, gloo, a test double) that receive device pointers and the CUDA stream to enqueue on. */
enum { TRL_COMM_ID_BYTES = 128 };
int trl_comm_unique_id(void* id128);                                           /* ncclGetUniqueId; call on one rank, ship to all */
int trl_comm_init(trl_handle* h, const void* id128, int rank, int world);     /* ncclCommInitRank on the handle's device */
typedef struct trl_collectives {
    void* ctx;
    int (*all_gather)(void* ctx, const void* send, void* recv, size_t bytes_per_rank, void* cuda_stream);
    int (*broadcast)(void* ctx, void* buf, size_t bytes, int root, void* cuda_stream);
    int (*all_reduce_sum_f64)(void* ctx, void* buf, size_t count, void* cuda_stream);
} trl_collectives;
int trl_comm_init_external(trl_handle* h, const trl_collectives* coll, int rank, int world);
int trl_comm_destroy(trl_handle* h);
int trl_comm_info(trl_handle* h, int* rank, int* world, int* block_rows);
/* env i of this rank is reported as env_offset + i in gathered blocks (default rank * num_envs) */
int trl_comm_set_env_offset(trl_handle* h, int64_t env_offset);
int trl_gather_tuples(trl_handle* h, int block_rows);
/* gathered blocks as device memory: world blocks of block_bytes; block = {i32 count, i32 queued, i32 rank, i32 block_rows,
 * u32 flags[block_rows], i32 env[block_rows], f32 rows[block_rows][1+S+A+S]}.  Ordered on the handle's stream. */
int trl_gathered_blocks(trl_handle* h, const void** dev_blocks, size_t* block_bytes, int* block_rows, int* row_width);
/* host copy for tests / a CPU-side trainer: counts[world]; rows/flags/env may be NULL; cap in rows; synchronises */
int trl_gathered_fetch(trl_handle* h, int32_t* counts, float* rows, uint32_t* flags, int32_t* env, int cap, int* n_total);
/* exposed device time of the last gather (pack end -> all-gather end on the comm stream), ms; synchronises */
int trl_gather_last_ms(trl_handle* h, double* ms);
int trl_trainer_add_gathered(trl_trainer* t);
int trl_trainer_broadcast(trl_trainer* t, int root);
int trl_comm_broadcast_weights(trl_handle* h, int root);                       /* same for a handle without a trainer (eval ranks) */
int trl_comm_eval_stats(trl_handle* h, int64_t* cycles, int64_t* episodes, double* avg_dist, int64_t* env_steps);
/* max - min over ranks of every trainer parameter (replica drift check), via two all-reduces of +/-theta maxima; synchronises */
int trl_trainer_replica_spread(trl_trainer* t, double* max_abs_diff);

const char* trl_last_error(void);

/* ---- EXPERIMENT (not on the decision path, which stays f64): the policy's wide inner product terr_ip0 (5984 -> 64; cNeuralNet::Eval,
 * learning/NeuralNet.cpp:352-375) as a TMA-fed tcgen05.mma GEMM over a batch of decisions in split precision (csrc/trl_tc_policy.cu);
 * tools/tc_policy_probe.py measures what the narrow arithmetic does to real decisions.  kind: 0 = bf16 parts, 1 = tf32 parts.
 * trl_tc_split: x[rows][K] f64 -> planes[parts][rows][K]; trl_tc_fc: out[M][64] (FP32) = sum over the part pairs in `pairs` (bit 3 i + j)
 * of A_i[M][K] B_j[64][K]^T; all pointers are device pointers, stream is a cudaStream_t, *err_flag is set if a barrier wait gave up. */
int trl_tc_split(const double* x, long long rows, int K, int kind, int parts, void* planes, void* stream);
int trl_tc_fc(const void* a_planes, const void* b_planes, int M, int K, int kind, int parts, int pairs, int ksplit, float* out, int* err_flag,
              void* stream);
const char* trl_tc_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
