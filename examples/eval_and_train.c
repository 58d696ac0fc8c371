/* Plain-C use of the C ABI (include/terrainrl_b200.h): what a maintainer's adapter does, without Python.
 *   gcc -std=c99 -I include examples/eval_and_train.c -L deepterrainrl_b200/lib -lterrainrl_b200 -Wl,-rpath,$PWD/deepterrainrl_b200/lib -o /tmp/eval_and_train
 *   /tmp/eval_and_train assets/dog_slopes_mixed.trlpack 4096
 * Without a GPU the create call fails with a message (there is no CPU fallback); tests/test_abi_cpu.py builds and runs it for that. */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "terrainrl_b200.h"

static int die(const char* what) {
    fprintf(stderr, "%s: %s\n", what, trl_last_error());
    return 2;
}

int main(int argc, char** argv) {
    const char* pack = argc > 1 ? argv[1] : "assets/dog_slopes_mixed.trlpack";
    const int n = argc > 2 ? atoi(argv[2]) : 1024;

    /* cOptScenarioPoliEval: evaluate the shipped policy for 10 s of simulated time */
    trl_handle* ev = trl_create_from_pack(pack, n, 0, TRL_MODE_POLI_EVAL, NULL, 1234);
    if (!ev) return die("trl_create_from_pack");
    for (int u = 0; u < 300; ++u)
        if (trl_update(ev, 1.0 / 30.0)) return die("trl_update");
    int64_t cycles = 0, episodes = 0, steps = 0;
    double avg_dist = 0;
    if (trl_eval_stats(ev, &cycles, &episodes, &avg_dist, &steps)) return die("trl_eval_stats");
    printf("eval: %lld env-steps, %lld cycles, %lld episodes, avg distance %.2f m\n", (long long)steps, (long long)cycles,
           (long long)episodes, avg_dist);
    trl_destroy(ev);

    /* cScenarioTrainMACE: rollout + on-device trainer, annealed exploration (args/opt_args_train_mace.txt) */
    trl_handle* ex = trl_create_from_pack(pack, n, 0, TRL_MODE_EXPLORE, NULL, 99);
    if (!ex) return die("trl_create_from_pack (explore)");
    const double tp[10] = {500000, 2000, 1, 500, 1, 0.9, 1e-3, 0.9, 5e-4, 7};
    trl_trainer* tr = trl_trainer_create(ex, tp);
    if (!tr) return die("trl_trainer_create");
    const double sched[9] = {0.9, 0.2, 20, 0.025, 0.9, 0.002, 50000, 50000, 0};
    if (trl_train_run(tr, sched, 200, 2, 32, 1.0 / 30.0)) return die("trl_train_run");
    int64_t c[9];
    double loss[2];
    if (trl_trainer_counters(tr, c, loss)) return die("trl_trainer_counters");
    printf("train: iter %lld, %lld tuples in the replay memory, critic loss %.5f\n", (long long)c[0], (long long)c[3], loss[0]);
    if (trl_output_model(ex, "/tmp/terrainrl_b200_model.h5", 0)) return die("trl_output_model");
    trl_trainer_destroy(tr);
    trl_destroy(ex);
    return 0;
}
