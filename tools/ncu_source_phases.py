"""Per-phase instruction / stall-sample breakdown of the step kernel from an `ncu --set full --import-source on` report.

Usage (here, no GPU needed):  ncu -i <report.ncu-rep> --page source --print-source cuda,sass --csv > /tmp/src.csv
                              python tools/ncu_source_phases.py /tmp/src.csv
Sums "Instructions Executed" (warp-level) and "# Samples" (warp-state samples) of the per-source-line rows over line ranges of
deepterrainrl_b200/csrc/trl_step.cu that correspond to the phases of one env-step.  The ranges below are those of the kernel the
round-1 report was taken from (commit 9fae429); re-derive them from the `// ----` markers when the file moves."""
import collections
import csv
import sys

PHASES = [(60, 93, "helpers (shuffle wrappers, wrap_pi, warp sums)"),
          (94, 299, "rng / per-env scalar helpers (reward, tuples, actions)"),
          (300, 337, "load_link (constant-memory model -> registers)"),
          (338, 368, "kinematics (pointer-jumping prefix sums)"),
          (369, 408, "ctrl: body inertia, RNEA bias force, composite inertia"),
          (409, 438, "ctrl: CRBA mass matrix -> shared memory"),
          (439, 466, "ctrl: swing / stance feedback"),
          (467, 501, "ctrl: implicit-PD system assembly"),
          (502, 540, "ctrl: register LDL^T + substitutions"),
          (541, 618, "ctrl: gravity compensation"),
          (619, 665, "ctrl: stance feedback, virtual forces, clamp"),
          (666, 685, "phys: rigid inertia / bias force"),
          (686, 748, "phys: contacts (corner rounds vs height field)"),
          (749, 777, "phys: ABA inward pass"),
          (778, 788, "phys: floating-base 3x3 solve"),
          (789, 812, "phys: ABA outward pass + integration"),
          (813, 915, "policy state, reset, load / store env"),
          (916, 1087, "kernel body (gait FSM, fall checks, scheduling, sub-step loop)")]


def main(path):
    cur, hdr = None, None
    data = collections.defaultdict(dict)
    for r in csv.reader(open(path)):
        if not r:
            continue
        if r[0] == "File Path":
            cur = r[1]
        elif r[0] == "Line No":
            hdr = r
        elif r[0] not in ("", "Function Name") and hdr:
            try:
                ln = int(r[0])
            except ValueError:
                continue
            gi = lambda k: int(r[hdr.index(k)]) if r[hdr.index(k)] not in ("-", "") else 0
            data[cur][ln] = (gi("Instructions Executed"), gi("# Samples"))
    tot = sum(v[0] for d in data.values() for v in d.values())
    tots = sum(v[1] for d in data.values() for v in d.values())
    step = next(f for f in data if f.endswith("trl_step.cu"))
    print("| phase | warp-instructions | share | warp-state samples | share |\n|---|---|---|---|---|")
    for a, b, name in PHASES:
        i = sum(v[0] for l, v in data[step].items() if a <= l <= b)
        s = sum(v[1] for l, v in data[step].items() if a <= l <= b)
        print(f"| {name} (trl_step.cu:{a}-{b}) | {i / 1e6:.2f} M | {100 * i / tot:.1f} % | {s} | {100 * s / tots:.1f} % |")
    for f in data:
        if f == step:
            continue
        i = sum(v[0] for v in data[f].values())
        s = sum(v[1] for v in data[f].values())
        print(f"| {f.split('/')[-1]} | {i / 1e6:.2f} M | {100 * i / tot:.1f} % | {s} | {100 * s / tots:.1f} % |")
    print(f"| total | {tot / 1e6:.2f} M | | {tots} | |")


if __name__ == "__main__":
    main(sys.argv[1])
