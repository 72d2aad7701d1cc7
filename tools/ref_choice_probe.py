import sys, os, numpy as np, torch
sys.path[:0] = ['/root/repo/oracle/refshim', '/root/repo/tools']
import refppo, refdriver as rd
ns = refppo.load_namespace()
for seed in range(4):
    algo, env = refppo.make_algo(ns, "coop_scalable", 1, 1, 1, seed=seed)
    r = algo.rollout
    r.reset()
    st = r.prev_state
    prev = np.concatenate([i.flatten() for i in list(st.values())])
    ps = []
    for i in range(2):
        f, ex = r.obs_car_ped_d(prev, i, 0)
        p = algo.actor_net_choice(torch.tensor(f).float().unsqueeze(0)).detach().numpy().ravel()
        ps.append((np.round(f, 2).tolist(), np.round(p, 4).tolist()))
    print(seed, ps)
# a few training iterations of the unmodified reference
algo, env = refppo.make_algo(ns, "coop_scalable", 1, 1, 1, seed=0)
import io, contextlib
for it in range(6):
    algo.rollout.reset()
    with contextlib.redirect_stdout(io.StringIO()):
        algo.rollout.iterations_rand(algo.actor_net_cross, algo.actor_net_wait, algo.actor_net_choice, algo.cov_mat, algo.cov_mat_d, 400, 0.0)
    print("iter", it, "cross", len(algo.rollout.batch_obs_cross), "wait", len(algo.rollout.batch_obs_wait), "choice acts", np.array(algo.rollout.batch_acts_choice).ravel()[:10])
    rt = algo.rollout.futur_rewards()
    for _ in range(10):
        if len(algo.rollout.batch_obs_cross): algo.train_model_c(algo.actor_net_cross, algo.critic_net_cross, algo.optimizer_actor_cross, algo.optimizer_critic_cross, np.array(algo.rollout.batch_obs_cross), np.array(algo.rollout.batch_acts_cross), np.array(algo.rollout.batch_log_probs_cross), rt[0], algo.cov_mat)
        if len(algo.rollout.batch_obs_wait): algo.train_model_c(algo.actor_net_wait, algo.critic_net_wait, algo.optimizer_actor_wait, algo.optimizer_critic_wait, np.array(algo.rollout.batch_obs_wait), np.array(algo.rollout.batch_acts_wait), np.array(algo.rollout.batch_log_probs_wait), rt[1], algo.cov_mat)
    for _ in range(10):
        algo.train_model_d(algo.actor_net_choice, algo.critic_net_choice, algo.optimizer_actor_choice, algo.optimizer_critic_choice, np.array(algo.rollout.batch_obs_choice), np.array(algo.rollout.batch_acts_choice), np.array(algo.rollout.batch_log_probs_choice), rt[2], algo.cov_mat_d)
