"""Build the TEST-ONLY host compilation of the kernel logic (see hostsim.cpp)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, "libhostsim.so")
CSRC = os.path.join(HERE, "..", "..", "mh-ppo_b200", "csrc")


SAN_LIB = os.path.join(HERE, "libhostsim_san.so")


def build_sanitized(force=False):
    """The same translation unit under AddressSanitizer + UndefinedBehaviorSanitizer (every UB report aborts).  The host build
    walks the kernel's own thread body (env_step.cuh / env_core.cuh / env_state.cuh), so out-of-range slots, lanes, shifts and
    arena indices of the step / reset / injection / snapshot logic are caught here; compute-sanitizer is closed on the GPU pool
    (profiles/round2_sanitizer_closed_on_pool.txt)."""
    srcs = [os.path.join(HERE, "hostsim.cpp")] + [os.path.join(CSRC, f) for f in ("env_core.cuh", "env_state.cuh", "env_step.cuh", "philox.cuh")]
    if not force and os.path.exists(SAN_LIB) and os.path.getmtime(SAN_LIB) >= max(os.path.getmtime(s) for s in srcs):
        return SAN_LIB
    subprocess.run(["g++", "-O1", "-g", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-D_GNU_SOURCE",
                    "-fsanitize=address,undefined", "-fno-sanitize-recover=all", "-fno-omit-frame-pointer",
                    "-o", SAN_LIB, srcs[0], "-lm"], check=True, capture_output=True)
    return SAN_LIB


def build(force=False):
    srcs = [os.path.join(HERE, "hostsim.cpp")] + [os.path.join(CSRC, f) for f in ("env_core.cuh", "env_state.cuh", "env_step.cuh", "philox.cuh")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(s) for s in srcs):
        return LIB
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-D_GNU_SOURCE",
                    "-o", LIB, srcs[0], "-lm"], check=True, capture_output=True)
    return LIB


if __name__ == "__main__":
    print(build(force=True))
