#!/usr/bin/env python
"""bench.py -- headline benchmark of the MH-PPO hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

Metric (BASELINE.json): agent-steps/sec of env.step.  A "step" is one env.step over the whole batch
of synthetic envs held by a GPU.  Default workload = the configuration the north-star target is
quoted on: Env_hybrid_multi_coop_scalable, nb_car=4 / nb_ped=3 / nb_lines=2, 262144 envs per GPU
(BASELINE.json configs[3]); envs shard across ranks with no data-path collective ("weak" scaling).

`value`  : device-timed (CUDA events on the launching stream, one pair per step, max over ranks),
           inputs resident in HBM, L2 flushed (untimed 256 MiB memset) between timed steps.
`e2e`    : the same metric through the reference-facing call with HOST buffers
           (mhppo_env_step_host: numpy-in/numpy-out convention of env.step): H2D of the actions and
           D2H of obs/rewards/reward_light/done inside the timed region.
`roofline`: algorithmic bytes per env-step (B = 2*(32C+64P+16) + 8*C_act + 8*C +
           4*obs_floats + 4) x envs / mean kernel time, against MEASURED_PEAKS.json hbm_gbs.
`cpu_baseline`: the C oracle port (oracle/, the only CPU code; test infrastructure) timed on this
           box's host cores on a bounded sample of the same workload.
--impl reference: the reference's own CPU path.  The reference is pure Python and cannot travel to
the GPU box; the arm times the pinned C restatement (oracle port) with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (variant, nb_car, nb_ped, nb_lines, envs per GPU)
    "scalable_432": ("coop_scalable", 4, 3, 2, 262144),      # BASELINE configs[3] (north-star target)
    "coop_212": ("coop", 2, 1, 2, 4096),                     # configs[1]
    "4cars_222": ("coop_4cars", 2, 2, 2, 65536),             # configs[2]
    "stop_121": ("stop", 1, 2, 1, 4096),                     # configs[0] shape, batched
}


def algorithmic_bytes(variant, nb_car, nb_ped, nb_lines):
    """SURVEY.md 8(d) per-env-step figure."""
    if variant == "coop_scalable":
        C, C_act, carw, envw = 2 * nb_lines, 2 * nb_lines, 7, 4
    elif "4cars" in variant:
        C, C_act, carw, envw = 2 * nb_car, (2 * nb_car if variant.endswith("2") else nb_car), 6, 3
    else:
        C, C_act, carw, envw = nb_car, nb_car, 6, 3
    obs = carw * C + envw + 9 * nb_ped
    S = 32 * C + 64 * nb_ped + 16
    return 2 * S + 8 * C_act + 8 * C + 4 * obs + 4, C, obs


def recorded_traffic(n_envs):
    """dram__bytes_read.sum + dram__bytes_write.sum of k_env_step from the committed ncu --set full capture."""
    p = os.path.join(ROOT, "profiles", "round1_env_step_ncu_metrics.json")
    try:
        d = json.load(open(p))
        return d["traffic_bytes_per_launch"] if int(d["n_envs"]) == int(n_envs) else None
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy burst)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks / throttle reasons during the timed region."""
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._stop_ev = index, [], threading.Event()

    def run(self):
        while not self._stop_ev.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(",")]
                if len(f) >= 6:
                    self.rows.append(f)
            except Exception:
                pass
            self._stop_ev.wait(0.2)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=3)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def cpu_port_rate(variant, nb_car, nb_ped, nb_lines, n_envs, n_steps, threads=0):
    """env-steps/s of the C oracle port on the host cores (actions pre-generated, untimed)."""
    import numpy as np
    from oracle import oracle as O
    env = O.OracleVecEnv(variant, n_envs, nb_car, nb_ped, nb_lines, seed=1234, n_threads=threads, store_f32=False)
    env.reset()
    rng = np.random.default_rng(0)
    A = env.n_action
    pool = []
    for _ in range(4):
        a = np.empty((n_envs, A))
        a[:, :A // 2] = rng.uniform(-4, 2, (n_envs, A // 2))
        a[:, A // 2:] = rng.choice([-1.0, 1.0], (n_envs, A // 2))
        pool.append(a)
    for t in range(3):
        env.step(pool[t % 4], autoreset=True)
    t0 = time.perf_counter()
    for t in range(n_steps):
        env.step(pool[t % 4], autoreset=True)
    dt = time.perf_counter() - t0
    s = env.get_state()
    active = float((s["env_i"][:, 1] + s["env_i"][:, 2]).mean())
    return n_envs * n_steps / dt, dt, active


def run_ppo(mh, torch, dist, world, rank, dev, n_envs, iters, warm=1):
    """PPO samples/s (BASELINE.json metric, second half): one iteration = one 80-step episode in every env
    (Env_rollout.iterations_rand) + reward-to-go + 10 x (cross, wait) + 10 x choice update epochs (Algo_PPO.train)."""
    env = mh.VecCrosswalkEnv("coop_scalable", n_envs, nb_car=4, nb_ped=3, nb_lines=2, seed=1234, env_id0=rank * n_envs, device=dev)
    torch.manual_seed(0)
    algo = mh.Algo_PPO(mh.Model_PPO, env, num_states_c=13, num_states_d=30, num_actions=1, mean=-1.0, std=3.0, nb_cars=4, dt=0.3)
    r = algo.rollout
    ev = lambda: torch.cuda.Event(enable_timing=True)
    for _ in range(warm):
        r.iterations_rand(algo.actor_net_cross, algo.actor_net_wait, algo.actor_net_choice)
        algo.update()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    n0 = mh.launch_count()
    t_roll = t_upd = 0.0
    samples = 0
    for _ in range(iters):
        e0, e1, e2 = ev(), ev(), ev()
        e0.record()
        r.iterations_rand(algo.actor_net_cross, algo.actor_net_wait, algo.actor_net_choice)
        e1.record()
        algo.update()
        e2.record()
        torch.cuda.synchronize()
        t_roll += e0.elapsed_time(e1); t_upd += e1.elapsed_time(e2)
        samples += sum(r.counts())
    launches = mh.launch_count() - n0
    t = torch.tensor([t_roll + t_upd, t_roll, t_upd, float(samples)], dtype=torch.float64, device=dev)
    if world > 1:
        mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX); dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tot_ms, roll_ms, upd_ms, samples = mx[0].item(), mx[1].item(), mx[2].item(), t[3].item()
    else:
        tot_ms, roll_ms, upd_ms, samples = [x.item() for x in t]
    del algo, env
    torch.cuda.empty_cache()
    return {"samples_per_s": samples / (tot_ms * 1e-3), "unit": "PPO samples/s (cross + wait + choice samples consumed by the update)",
            "iteration_ms": tot_ms / iters, "rollout_ms": roll_ms / iters, "update_ms": upd_ms / iters, "iterations": iters,
            "samples_per_iteration": samples / iters, "n_envs_per_gpu": n_envs, "update_epochs": "10 x (cross, wait) + 10 x choice",
            "gpu_launches": int(launches),
            "config": "Coop-MH-PPO-scalable on Env_hybrid_multi_coop_scalable nb_car=4 nb_ped=3 nb_lines=2, %d envs/GPU x 80 steps" % n_envs}


def run_reference(args, wl):
    variant, nb_car, nb_ped, nb_lines, n_envs = wl
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    n_sample = 16384                      # bounded sample of the n_envs-per-GPU batch, per step
    rate, dt, active = cpu_port_rate(variant, nb_car, nb_ped, nb_lines, n_sample, max(args.steps, 1), threads=cores)
    B, C, nobs = algorithmic_bytes(variant, nb_car, nb_ped, nb_lines)
    sample = "%d envs x %d steps of the %d-env batch, C oracle port (oracle/mhppo_oracle.c), %d threads" % (
        n_sample, args.steps, n_envs, cores)
    val = rate * active
    line = {"metric": "agent-steps/sec (env.step)", "value": val, "unit": "agent-steps/s", "impl": "reference",
            "n_gpus": args.gpus, "steps": args.steps, "warmup": 3, "ms_per_step": 1e3 * dt / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Env_hybrid_multi_%s nb_car=%d nb_ped=%d nb_lines=%d, %d envs/GPU, env.step" % (
                variant, nb_car, nb_ped, nb_lines, n_envs)},
            "env_steps_per_s": rate, "active_agents_per_env": active,
            "cpu_baseline": {"value": val, "unit": "agent-steps/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "agent-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "note": "the reference is single-process Python (one env, global RNG, GIL) and is not shippable to the GPU box; "
                    "this arm is its pinned C restatement on all host threads, far faster than the Python original "
                    "(3.4k env-steps/s/core, BASELINE.md section 2)"}
    emit(line)


_JSON_FD = None


def claim_stdout():
    """Libraries (NCCL's version banner, torchrun) write to file descriptor 1; the contract is ONE JSON line on stdout.
    Everything else is sent to stderr and the line is written to the original descriptor at the end."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode()); sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=160)
    ap.add_argument("--warmup", type=int, default=8)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="scalable_432", choices=sorted(WORKLOADS))
    ap.add_argument("--envs", type=int, default=0, help="override envs per GPU")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-flush", action="store_true", help="do not flush L2 between timed steps")
    ap.add_argument("--ppo-envs", type=int, default=131072, help="envs per GPU of the PPO samples/s section (0 = skip)")
    ap.add_argument("--ppo-iters", type=int, default=2)
    ap.add_argument("--row-major-actions", action="store_true", help="device-resident actions as a contiguous [N, A] tensor (gym layout)")
    args = ap.parse_args()
    wl = list(WORKLOADS[args.workload])
    if args.envs:
        wl[4] = args.envs
    if args.impl == "reference":
        return run_reference(args, wl)
    variant, nb_car, nb_ped, nb_lines, n_envs = wl

    import torch
    import torch.distributed as dist
    import mhppo_b200

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    W = max(args.warmup, 3)
    K = args.steps

    env = mhppo_b200.VecCrosswalkEnv(variant, n_envs, nb_car=nb_car, nb_ped=nb_ped, nb_lines=nb_lines,
                                     seed=1234, env_id0=rank * n_envs, device=dev, autoreset=True)
    A = env.n_action
    state_bytes, n_obs_env, n_lead_env = env.state_bytes_per_env, env.n_obs, env.n_lead
    g = torch.Generator(device=dev).manual_seed(1000 + rank)
    pool = []
    for _ in range(4):   # synthetic actions resident in HBM: acc ~ U(-4,2), light in {-1,+1}
        # [N, A] view of a component-major buffer: the layout the rollout kernels hand to env.step (coalesced reads)
        a = torch.empty(n_envs, A, device=dev) if args.row_major_actions else torch.empty(A, n_envs, device=dev).t()
        a[:, :A // 2] = torch.rand(n_envs, A // 2, device=dev, generator=g) * 6 - 4
        a[:, A // 2:] = (torch.rand(n_envs, A // 2, device=dev, generator=g) < 0.5).float() * 2 - 1
        pool.append(a)
    flush = None if args.no_flush else torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    env.reset()
    for t in range(W):
        env.step(pool[t % 4])
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-timed region: K steps, one event pair per step, L2 flushed between steps -----------
    sampler = ClockSampler(local)
    sampler.start()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    barrier()
    n0 = mhppo_b200.launch_count()
    wall0 = time.perf_counter()
    for t in range(K):
        if flush is not None:
            flush.zero_()
        ev[t][0].record()
        env.step(pool[t % 4])
        ev[t][1].record()
    barrier()
    wall = time.perf_counter() - wall0
    launches = mhppo_b200.launch_count() - n0
    dev_ms = sum(a.elapsed_time(b) for a, b in ev)

    # ---- end-to-end region: host buffers through the reference-facing call -------------------------
    Ke = max(8, min(K, 40))
    act_h = [p.cpu().pin_memory() for p in pool]
    obs_h = torch.empty(n_envs, env.n_obs).pin_memory()
    rew_h = torch.empty(n_envs, env.n_lead).pin_memory()
    rl_h = torch.empty(n_envs, env.n_lead).pin_memory()
    done_h = torch.empty(n_envs, dtype=torch.uint8).pin_memory()
    for t in range(3):
        env.step_host(act_h[t % 4], obs_h, rew_h, rl_h, done_h)
    barrier()
    e0 = time.perf_counter()
    for t in range(Ke):
        env.step_host(act_h[t % 4], obs_h, rew_h, rl_h, done_h)
    barrier()
    e2e_s = time.perf_counter() - e0
    h2d = n_envs * A * 4
    d2h = n_envs * (n_obs_env + 2 * n_lead_env) * 4 + n_envs

    st = env.get_state()
    active = float((st["env_i"][:, 1] + st["env_i"][:, 2]).float().mean().item())
    flushed = flush is not None
    del st, env, flush, pool, act_h, obs_h
    torch.cuda.empty_cache()
    ppo = run_ppo(mhppo_b200, torch, dist, world, rank, dev, args.ppo_envs, args.ppo_iters) if args.ppo_envs > 0 and args.ppo_iters > 0 else None
    clocks = sampler.stop()        # sampled every 200 ms from the start of the device-timed region to the end of the PPO section
    t_dev = torch.tensor([dev_ms, e2e_s, active], dtype=torch.float64, device=dev)
    if world > 1:
        mx = t_dev.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        dist.all_reduce(t_dev, op=dist.ReduceOp.SUM)
        dev_ms, e2e_s, active = mx[0].item(), mx[1].item(), t_dev[2].item() / world

    if rank == 0:
        B, C, nobs = algorithmic_bytes(variant, nb_car, nb_ped, nb_lines)
        peak, peak_src = measured_peaks()
        env_steps = world * n_envs * K / (dev_ms * 1e-3)
        kern_ms = dev_ms / K
        achieved = B * n_envs / (kern_ms * 1e-3) / 1e9
        line = {
            "metric": "agent-steps/sec (env.step)", "value": env_steps * active, "unit": "agent-steps/s",
            "n_gpus": world, "steps": K, "warmup": W, "ms_per_step": kern_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "Env_hybrid_multi_%s nb_car=%d nb_ped=%d nb_lines=%d, %d envs/GPU, env.step with in-kernel auto-reset" % (
                variant, nb_car, nb_ped, nb_lines, n_envs),
                "l2": "flushed: an untimed 256 MiB memset (2x the 126 MB L2) runs between timed steps; the per-step working set is ~150 MB"
                if flushed else "not flushed", "state_bytes_per_env": state_bytes,
                "parallelism": "env shards per rank, no data-path collective"},
            "env_steps_per_s": env_steps, "slot_agent_steps_per_s": env_steps * (C + nb_ped),
            "active_agents_per_env": active, "wall_s_timed_region": wall, "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": recorded_traffic(n_envs) if args.workload == "scalable_432" else None,
                         "traffic_source": "profiles/round1_env_step_ncu_metrics.json (ncu --set full, bytes per launch)", "algorithmic_bytes_per_env_step": B, "peak_source": peak_src,
                         "kernel": "k_env_step<%s>" % variant},
            "e2e": {"value": world * n_envs * Ke / e2e_s * active, "unit": "agent-steps/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "steps": Ke, "call": "mhppo_env_step_host (pinned host buffers, H2D+kernel+D2H+sync)"},
        }
        if ppo is not None:
            line["ppo"] = ppo
        if not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            ns = 16384
            r0, dt0, _ = cpu_port_rate(variant, nb_car, nb_ped, nb_lines, ns, 20, threads=cores)      # calibration
            ks = int(max(40, min(20000, 12.0 * r0 / ns)))                                            # ~12 s of CPU work
            rate, dt, act_cpu = cpu_port_rate(variant, nb_car, nb_ped, nb_lines, ns, ks, threads=cores)
            line["cpu_baseline"] = {"value": rate * act_cpu, "unit": "agent-steps/s", "cores": cores, "kind": "port",
                                    "sample": "%d envs x %d steps of the same workload, C oracle port on all host threads (%.1f s)" % (ns, ks, dt),
                                    "env_steps_per_s": rate}
        emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
