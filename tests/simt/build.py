"""TEST INFRASTRUCTURE ONLY: compile the kernel + host sources of deepterrainrl_b200/csrc for the SIMT emulator (g++, no nvcc,
no GPU) into tests/simt/_build/libterrainrl_simt[_<tag>].so.  `defines` selects an experimental kernel variant exactly as
TRL_NVCC_EXTRA does for the nvcc build (e.g. ["-DTRL_ACCUM_SMEM=1"])."""
import fcntl
import hashlib
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "deepterrainrl_b200", "csrc")
UNITS = ["trl_step.cu", "trl_step_cg.cu", "trl_host.cu", "trl_train.cu", "trl_comm.cu", "trl_probe.cu", "ref_loader.cpp"]
LOCAL_UNITS = ["simt_runtime.cpp", "selftest.cu"]
VARIANT_UNITS = ["trl_step.cu", "trl_step_cg.cu"]     # the only units the TRL_* experiment knobs reach
BUILD_ROOT = os.path.join(os.environ.get("TRL_BUILD_DIR", "/tmp/terrainrl_b200_build"), "simt")     # outside the tree: nothing here travels
CXXFLAGS = ["-std=c++17", "-O2", "-g1", "-fPIC", "-ffp-contract=off", "-fno-strict-aliasing", "-DTRL_SIMT_EMU=1", "-w",
            "-I", os.path.join(HERE, "include"), "-I", CSRC, "-I", os.path.join(ROOT, "include")]


def build(defines=(), force=False):
    defines = list(defines)
    tag = hashlib.md5(" ".join(defines).encode()).hexdigest()[:8] if defines else "default"
    bdir = os.path.join(BUILD_ROOT, tag)
    os.makedirs(bdir, exist_ok=True)
    out = os.path.join(bdir, "libterrainrl_simt.so")
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(HERE, f) for f in ["simt_runtime.h"] + LOCAL_UNITS]
    deps += [os.path.join(HERE, "include", f) for f in os.listdir(os.path.join(HERE, "include"))]
    lock = open(os.path.join(bdir, ".lock"), "w")
    fcntl.flock(lock, fcntl.LOCK_EX)      # pytest-xdist workers may ask for the same build at the same time
    try:
        return _build_locked(out, bdir, deps, defines, force)
    finally:
        fcntl.flock(lock, fcntl.LOCK_UN)
        lock.close()


def _compile(cxx, units, objdir, defines, deps, force):
    """compile `units` into objdir (skipping objects newer than every dependency); returns the object paths"""
    os.makedirs(objdir, exist_ok=True)
    procs, objs = [], []
    newest = max(os.path.getmtime(d) for d in deps)
    for u in units:
        src = os.path.join(HERE if u in LOCAL_UNITS else CSRC, u)
        obj = os.path.join(objdir, os.path.splitext(u)[0] + ".o")
        objs.append(obj)
        if not force and os.path.exists(obj) and os.path.getmtime(obj) >= newest:
            continue
        cmd = [cxx] + CXXFLAGS + defines + ["-x", "c++", "-c", src, "-o", obj]
        procs.append((u, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for u, p in procs:
        log = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError(f"simt build of {u} failed:\n{log[-6000:]}")
    return objs


def _build_locked(out, bdir, deps, defines, force):
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    cxx = os.environ.get("CXX", "g++")
    # the knobs only reach the env-step translation units; everything else is compiled once and shared by all variants
    common_dir = os.path.join(BUILD_ROOT, "common")
    os.makedirs(common_dir, exist_ok=True)
    with open(os.path.join(common_dir, ".lock"), "w") as clock:
        fcntl.flock(clock, fcntl.LOCK_EX)
        common = _compile(cxx, [u for u in UNITS + LOCAL_UNITS if u not in VARIANT_UNITS], common_dir, [], deps, force)
    objs = common + _compile(cxx, VARIANT_UNITS, bdir, defines, deps, force)
    # device functions defined in headers are not `inline` in CUDA sources: the same definition appears in several objects
    subprocess.run([cxx, "-shared", "-o", out + ".tmp"] + objs + ["-Wl,--allow-multiple-definition", "-lpthread", "-ldl"], check=True)
    os.replace(out + ".tmp", out)
    return out


if __name__ == "__main__":
    import sys
    print(build(sys.argv[1:], force=True))
