"""Model_PPO -- device-resident mirror of the reference's policy / value network class
(Coop-MH-PPO-scalable.py:42-93): 4-layer MLP in->32->64->32->out with ReLU and one of three heads
(0 linear, 1 std*tanh+mean, 2 softmax over logit pairs).

Same constructor arguments and the same `state_dict()` key layout (`layer{1..4}.{weight,bias}`), so the
shipped checkpoints (`load_model/weights/*.pth`) load unchanged.  The parameters live in ONE flat fp32
CUDA tensor in the layout the kernels consume (include/mhppo.h "Flat parameter layout"), next to the
Adam moments; `forward` exists for API parity / evaluation and is plain torch plumbing -- the hot paths
(rollout inference, update) run in libmhppo_b200.so.
"""
import math

import torch

from . import _lib


def _pack(sd, kp, device):
    """state_dict (reference layout) -> flat kernel layout."""
    w1, w4, b4 = sd["layer1.weight"].float(), sd["layer4.weight"].float(), sd["layer4.bias"].float()
    W1t = torch.zeros(kp, 32); W1t[: w1.shape[1]] = w1.t()
    W4t = torch.zeros(32, 4); W4t[:, : w4.shape[0]] = w4.t()
    B4 = torch.zeros(4); B4[: b4.shape[0]] = b4
    parts = [W1t, sd["layer1.bias"].float(), sd["layer2.weight"].float().t(), sd["layer2.bias"].float(),
             sd["layer3.weight"].float().t(), sd["layer3.bias"].float(), W4t, B4]
    return torch.cat([p.contiguous().reshape(-1).cpu() for p in parts]).to(device)


def _unpack(flat, n_in, n_out, kp):
    f = flat.detach().cpu()
    o = 0
    def take(n):
        nonlocal o
        v = f[o:o + n]; o += n
        return v
    W1t = take(kp * 32).reshape(kp, 32); b1 = take(32)
    W2t = take(32 * 64).reshape(32, 64); b2 = take(64)
    W3t = take(64 * 32).reshape(64, 32); b3 = take(32)
    W4t = take(32 * 4).reshape(32, 4); b4 = take(4)
    return {"layer1.weight": W1t[:n_in].t().contiguous(), "layer1.bias": b1.clone(), "layer2.weight": W2t.t().contiguous(),
            "layer2.bias": b2.clone(), "layer3.weight": W3t.t().contiguous(), "layer3.bias": b3.clone(),
            "layer4.weight": W4t[:, :n_out].t().contiguous(), "layer4.bias": b4[:n_out].clone()}


class Model_PPO:
    def __init__(self, np_inputs, nb_outputs, model_type=0, nb_car=1, mean=0, std=1, device="cuda"):
        L = _lib.lib()
        self.n_in, self.n_out, self.model_type, self.nb_car, self.mean, self.std = np_inputs, nb_outputs, model_type, nb_car, mean, std
        self.kp = L.mhppo_net_padded_in(np_inputs)
        if self.kp < 0:
            raise ValueError("unsupported input width %d" % np_inputs)
        self.n_param = L.mhppo_net_param_count(np_inputs)
        self.device = torch.device(device)
        # same initialisation as the reference: torch default Linear init + orthogonal_ on layer4 (PY:54-68)
        layers = [torch.nn.Linear(np_inputs, 32), torch.nn.Linear(32, 64), torch.nn.Linear(64, 32), torch.nn.Linear(32, nb_outputs)]
        torch.nn.init.orthogonal_(layers[3].weight)
        sd = {}
        for i, l in enumerate(layers):
            sd["layer%d.weight" % (i + 1)] = l.weight.detach()
            sd["layer%d.bias" % (i + 1)] = l.bias.detach()
        self.flat = _pack(sd, self.kp, self.device)
        assert self.flat.numel() == self.n_param
        self.m = torch.zeros_like(self.flat)
        self.v = torch.zeros_like(self.flat)
        self.grad = torch.zeros_like(self.flat)
        self.step = 0

    def state_dict(self):
        return _unpack(self.flat, self.n_in, self.n_out, self.kp)

    def load_state_dict(self, sd):
        self.flat.copy_(_pack(sd, self.kp, self.device))

    def parameters(self):
        return [self.flat]

    def forward(self, x):
        sd = {k: v.to(self.device) for k, v in self.state_dict().items()}
        x = torch.as_tensor(x, dtype=torch.float32, device=self.device)
        h = torch.relu(torch.nn.functional.linear(x, sd["layer1.weight"], sd["layer1.bias"]))
        h = torch.relu(torch.nn.functional.linear(h, sd["layer2.weight"], sd["layer2.bias"]))
        h = torch.relu(torch.nn.functional.linear(h, sd["layer3.weight"], sd["layer3.bias"]))
        o = torch.nn.functional.linear(h, sd["layer4.weight"], sd["layer4.bias"])
        if self.model_type == 2:
            return torch.flatten(torch.softmax(o.reshape(-1, 2), dim=-1))
        if self.model_type == 1:
            return torch.tanh(o) * self.std + self.mean
        return o

    __call__ = forward


class Adam:
    """torch.optim.Adam(params, lr) defaults (betas .9/.999, eps 1e-8), PY:719-724, on a flat parameter tensor."""

    def __init__(self, net, lr):
        self.net, self.lr = net, lr

    def step(self, grad_scale=1.0):
        import ctypes as C
        n = self.net
        n.step += 1
        _lib.check(_lib.lib().mhppo_adam(n.flat.data_ptr(), n.grad.data_ptr(), n.m.data_ptr(), n.v.data_ptr(), n.n_param,
                                         self.lr, 0.9, 0.999, 1e-8, n.step, grad_scale,
                                         C.c_void_p(torch.cuda.current_stream(n.device).cuda_stream)))
