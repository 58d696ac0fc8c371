"""cScenarioPoliEval's per-cycle analysis dumps (scenarios/ScenarioPoliEval.cpp:262-404) for one env of a batch, derived from the
tuple stream (one tuple per gait cycle: state at the decision, action taken, reward): `RecordAction` (action id + optimised
parameters), `RecordActionIDState` (action id + policy state) and the reward per cycle, in the reference's text formats
(std::to_string -> 6 decimals, ",\\t" separated).  The record lines are byte-identical to the files the reference's compiled
cScenarioPoliEval writes for the same cycles (tests/test_ref_pinning_cpu.py); `action_table_lines` gives the table of base actions
that cScenarioPoliEval::InitActionRecord puts in front of the action records (:298-319).  `RecordNNActivation` (a named layer's blob of the deploy net after the
decision, `record_nn_activation` fed from trl_get_layer_state) is written per cycle as well.  `RecordVel` (:353-367: the
COM's mean forward velocity over the cycle that just ended) does not travel with the tuples; `poll` derives it from the controller
block after an outer update (a gait cycle is much longer than one update, so at most one cycle ends per call)."""
import numpy as np


_RAPTOR_OPT = [False, True, True, False, False] + [True, False, True, True, True, True, True, True,
                                                  True, False, True, True, True, True, True, True,
                                                  False, False, True, True, True, True, True, True,
                                                  False, False, True, True, True, True, True, True]


def _read_pack(path):
    import struct
    out = {}
    with open(path, "rb") as f:
        assert f.read(8) == b"TRLPACK1"
        n, = struct.unpack("<I", f.read(4))
        for _ in range(n):
            ln, = struct.unpack("<I", f.read(4))
            name = f.read(ln).decode()
            dt, cnt = struct.unpack("<IQ", f.read(12))
            out[name] = np.frombuffer(f.read((4 if dt == 1 else 8) * cnt), "<i4" if dt == 1 else "<f8").copy()
    return out


def action_table_lines(pack_path):
    """cScenarioPoliEval::InitActionRecord: one line per base action, "%i" then ", %.5f" per optimised parameter of the blended
    controller parameter sets (c{Dog,Raptor}Controller::BuildActionOptParams -> BlendCtrlParams + GetOptParams,
    sim/DogController.cpp:381-388,1340-1346; optimised-parameter masks sim/DogController.cpp:81-121, sim/RaptorController.cpp:78-122)."""
    p = _read_pack(pack_path)
    raptor = int(p["meta_i32"][0]) == 2
    n_params = 37 if raptor else 30
    C = p["ctrl_params"].reshape(-1, n_params)
    A = p["actions"].reshape(-1, 4)
    opt = [i for i in range(n_params) if (_RAPTOR_OPT[i] if raptor else i != 0)]
    lines = []
    for a, (i0, i1, blend, _) in enumerate(A):
        params = (1 - blend) * C[int(i0)] + blend * C[int(i1)]
        lines.append("%i" % a + "".join(", %.5f" % params[k] for k in opt) + "\n")
    return lines


def _line(head, vec):
    return str(int(head)) + "".join(",\t%f" % v for v in vec) + "\n"


class CycleRecorder:
    def __init__(self, env, state_size, action_size, action_file=None, action_id_state_file=None, reward_file=None, vel_file=None,
                 pack=None, nn_activation_file=None, nn_activation_layer=None):
        self.env, self.S, self.A = env, state_size, action_size
        self.nn_layer = nn_activation_layer
        self.files = dict(action=action_file, ids=action_id_state_file, reward=reward_file, vel=vel_file,
                          nn=nn_activation_file if nn_activation_layer else None)
        self._last_cycle = 0
        self._acc_dx = self._time = self._prev_time = 0.0
        for f in self.files.values():
            if f:
                open(f, "w").close()                     # InitActionIDState / InitVelRecord: cFileUtil::ClearFile
        if action_file and pack:
            with open(action_file, "w") as f:            # InitActionRecord: the base-action table heads the action records
                f.writelines(action_table_lines(pack))
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

    def record_nn_activation(self, action_id, blob):
        """cScenarioPoliEval::RecordNNActivation (:271-296): action id + the named layer's blob after the decision that opened the
        cycle (cNeuralNet::GetLayerState; product side: BatchedScenario.GetLayerState(layer, env) = trl_get_layer_state).  Call once per
        valid cycle (the warm-up cycle is not recorded), before the cycle's action records; an empty blob writes nothing, as there."""
        if self.files["nn"] and len(blob) > 0:
            with open(self.files["nn"], "a") as f:
                f.write(_line(action_id, blob))

    def poll(self, ctrl, dt=1.0 / 30.0, episode_reset=False):
        """cScenarioPoliEval::RecordVel (:353-367).  Call ONCE after every Update(dt) with the env's controller block
        (scenario.GetCtrl(env), layout of trl_get_ctrl: [4] = duration of the cycle that just ended, [9] = the COM's x displacement
        over it, [-2] = cycle count) and whether that update ended the env's episode.

        The reference divides the COM displacement between two records by the difference of the scenario clock, which advances
        once per outer update (scenarios/ScenarioSimChar.cpp:153-154) -- i.e. by a multiple of dt, not by the cycle's duration --
        and it skips the record of the warm-up cycle (IsValidCycle, :406-410) without moving its reference point, so the first
        record spans the warm-up cycle too.  Both are reproduced.  A cycle that ends in the very update that also ends the
        episode is not recorded (its displacement is gone from the controller block after the reset)."""
        self._time += dt
        cyc = int(ctrl[-2])
        new = cyc - self._last_cycle
        if new > 0 and ctrl[4] > 0.0 and not episode_reset:
            self._acc_dx += ctrl[9]
            if cyc - 1 >= 1:                                 # gNumWarmupCycles = 1
                if self.files["vel"]:
                    with open(self.files["vel"], "a") as f:
                        f.write("%f\n" % (self._acc_dx / (self._time - self._prev_time)))
                self._acc_dx = 0.0
                self._prev_time = self._time
        if episode_reset:                                    # cScenarioPoliEval::Reset: clock, reference point and time restart
            self._time = self._prev_time = 0.0
            self._acc_dx = 0.0
        self._last_cycle = cyc
        return max(new, 0)
