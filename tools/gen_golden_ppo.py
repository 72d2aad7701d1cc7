"""Generate tests/golden/ppo_*.npz from the UNMODIFIED reference PPO classes (build container only).

  ppo_rollout_432.npz : whole-episode rollout buffers of Env_rollout.iterations_rand (sampling noise injected through
                        the policy-noise contract, tools/ref_rollout.py) for a few (seed, env_id) streams + the nets used
  ppo_eval_432.npz    : whole episodes of the deterministic evaluation rollout Env_rollout.iterations (env.step wrapped to
                        record actions / rewards / reward_light / waiting times) + the nets used
  ppo_train_{c,d}.npz : Algo_PPO.train_model_c / train_model_d on a synthetic batch: nets before, nets after 1..4 epochs
Re-run:  python tools/gen_golden_ppo.py
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, "oracle", "refshim"), os.path.join(ROOT, "tools")]
import ref_rollout  # noqa: E402
import refppo  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sd_np(prefix, net):
    return {prefix + "." + k: v.detach().numpy().copy() for k, v in net.state_dict().items()}


def main():
    ns = refppo.load_namespace()
    algo, env = refppo.make_algo(ns, "coop_scalable", 4, 3, 2, seed=1)
    out = {}
    for name, net in (("cross", algo.actor_net_cross), ("wait", algo.actor_net_wait), ("choice", algo.actor_net_choice)):
        out.update(sd_np(name, net))
    streams = [(777, 5), (778, 9), (901, 123), (5, 1)]
    out["streams"] = np.array(streams, np.int64)
    for e, (seed, env_id) in enumerate(streams):
        w = ref_rollout.reference_episode(ns, algo, env, seed, env_id)
        for k, v in w.items():
            out["ep%d.%s" % (e, k)] = v
    np.savez_compressed(os.path.join(OUT, "ppo_rollout_432.npz"), **out)

    # deterministic evaluation rollout (Env_rollout.iterations): both regimes (ped_traffic == nb_ped and < nb_ped)
    ev = {k: v for k, v in out.items() if k.split(".")[0] in ("cross", "wait", "choice")}
    streams = [(777, 5), (901, 123), (7, 3), (8, 4)]
    ev["streams"] = np.array(streams, np.int64)
    for e, (seed, env_id) in enumerate(streams):
        w = ref_rollout.reference_eval_episode(ns, algo, env, seed, env_id)
        for k, v in w.items():
            ev["ep%d.%s" % (e, k)] = v
    np.savez_compressed(os.path.join(OUT, "ppo_eval_432.npz"), **ev)

    for kind in ("c", "d"):
        g = torch.Generator().manual_seed(11)
        M, n_in = (700, 13) if kind == "c" else (90, 30)
        states = torch.randn(M, n_in, generator=g).numpy().astype(np.float32)
        rtgs = (torch.randn(M, generator=g) * 5 - 3).float()
        if kind == "c":
            actor, critic, oa, oc = algo.actor_net_cross, algo.critic_net_cross, algo.optimizer_actor_cross, algo.optimizer_critic_cross
            actions = (torch.randn(M, 1, generator=g) * 2 - 1).numpy().astype(np.float32)
            logp_old = (-torch.rand(M, generator=g) * 3).numpy().astype(np.float32)
            step, cov = algo.train_model_c, algo.cov_mat
        else:
            actor, critic, oa, oc = algo.actor_net_choice, algo.critic_net_choice, algo.optimizer_actor_choice, algo.optimizer_critic_choice
            actions = torch.randint(0, 2, (M, 1), generator=g).float().numpy()
            logp_old = (-torch.rand(M, generator=g) * 1.5).numpy().astype(np.float32)
            step, cov = algo.train_model_d, algo.cov_mat_d
        o = dict(states=states, rtgs=rtgs.numpy(), actions=actions, logp_old=logp_old)
        o.update(sd_np("actor0", actor)); o.update(sd_np("critic0", critic))
        for ep in range(1, 5):
            step(actor, critic, oa, oc, states, actions.astype(np.float64) if kind == "d" else actions, logp_old.astype(np.float64) if kind == "c" else logp_old, rtgs, cov)
            o.update(sd_np("actor%d" % ep, actor)); o.update(sd_np("critic%d" % ep, critic))
        np.savez_compressed(os.path.join(OUT, "ppo_train_%s.npz" % kind), **o)
    for f in sorted(os.listdir(OUT)):
        if f.startswith("ppo_"):
            print(f, os.path.getsize(os.path.join(OUT, f)) // 1024, "KiB")


if __name__ == "__main__":
    main()
