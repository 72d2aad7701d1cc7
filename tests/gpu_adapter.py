"""Adapter giving the CUDA VecCrosswalkEnv the same numpy-facing interface as the oracle wrapper,
so the same comparison driver (tests/common.compare_vec_envs) runs on both."""
import numpy as np
import torch

import mhppo_b200


class CudaEnv:
    def __init__(self, variant, n_envs, nb_car, nb_ped, nb_lines, seed=0, env_id0=0, **kw):
        self.env = mhppo_b200.VecCrosswalkEnv(variant, n_envs, nb_car=nb_car, nb_ped=nb_ped, nb_lines=nb_lines, seed=seed,
                                              env_id0=env_id0, **kw)
        self.N, self.n_action, self.n_obs, self.n_lead = n_envs, self.env.n_action, self.env.n_obs, self.env.n_lead

    def reset(self, mask=None):
        self.env.reset(mask=None if mask is None else torch.as_tensor(np.asarray(mask, np.uint8)))
        return self.env.flat_obs.cpu().numpy()

    def step(self, actions, autoreset=True):
        self.env.autoreset = bool(autoreset)
        a = torch.as_tensor(np.ascontiguousarray(actions, np.float32)).cuda()
        _, rew, done, _, _ = self.env.step(a)
        return (self.env.flat_obs.cpu().numpy(), rew.cpu().numpy(), self.env.reward_light.cpu().numpy(), done.cpu().numpy())

    def reset_pedestrian(self, num_ped, *vals, mask=None):
        self.env.reset_pedestrian(num_ped, *[torch.as_tensor(np.asarray(v, np.float32)).cuda() if np.ndim(v) else v for v in vals], mask=mask)

    def reset_cars(self, num_car, *vals, mask=None):
        self.env.reset_cars(num_car, *[torch.as_tensor(np.asarray(v, np.float32)).cuda() if np.ndim(v) else v for v in vals], mask=mask)

    def observe(self):
        self.env.observe()
        return self.env.flat_obs.cpu().numpy()

    def get_state(self):
        return {k: v.cpu().numpy() for k, v in self.env.get_state().items()}

    def set_state(self, s):
        self.env.set_state(s)
