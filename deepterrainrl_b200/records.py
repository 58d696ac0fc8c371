"""cScenarioPoliEval's per-cycle analysis dumps (scenarios/ScenarioPoliEval.cpp:262-404) for one env of a batch, derived from the
tuple stream (one tuple per gait cycle: state at the decision, action taken, reward): `RecordAction` (action id + optimised
parameters), `RecordActionIDState` (action id + policy state) and the reward per cycle, in the reference's text formats
(std::to_string -> 6 decimals, ",\\t" separated).  The record lines are byte-identical to the files the reference's compiled
cScenarioPoliEval writes for the same cycles (tests/test_ref_pinning_cpu.py); the table of base actions that
cScenarioPoliEval::InitActionRecord puts in front of the action records (:297-317) is not emitted."""
import numpy as np


def _line(head, vec):
    return str(int(head)) + "".join(",\t%f" % v for v in vec) + "\n"


class CycleRecorder:
    def __init__(self, env, state_size, action_size, action_file=None, action_id_state_file=None, reward_file=None):
        self.env, self.S, self.A = env, state_size, action_size
        self.files = dict(action=action_file, ids=action_id_state_file, reward=reward_file)
        for f in self.files.values():
            if f:
                open(f, "w").close()                     # InitActionIDState / InitVelRecord: cFileUtil::ClearFile
        self.cycles = 0

    def consume(self, rows, flags, env_ids):
        """rows [n][1 + S + A + S] as returned by GetTuples(); appends the cycles of the recorded env in arrival order."""
        sel = np.nonzero(np.asarray(env_ids) == self.env)[0]
        for i in sel:
            r = rows[i]
            s_beg = r[1:1 + self.S]
            act = r[1 + self.S:1 + self.S + self.A]
            if self.files["action"]:
                with open(self.files["action"], "a") as f:
                    f.write(_line(act[0], act[1:]))
            if self.files["ids"]:
                with open(self.files["ids"], "a") as f:
                    f.write(_line(act[0], s_beg))
            if self.files["reward"]:
                with open(self.files["reward"], "a") as f:
                    f.write("%f\n" % r[0])
            self.cycles += 1
        return len(sel)
