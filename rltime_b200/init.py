"""Parameter initialisation of the reference's modules, as flat {state_dict name: tensor}.

conv / linear / LSTM matrices: U(+-sqrt(1/fan_in)); conv / linear biases: zero
(rltime/models/torch/utils.py:6-25); the LSTM biases keep torch.nn.LSTMCell's default
U(+-1/sqrt(hidden)) (rltime/models/torch/modules/lstm.py:42-48 re-initialises only the two
weight matrices)."""
import math

import torch


def init_params(param_info, lstm_units, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = {}
    for name, shape in param_info:
        if name.endswith("bias") or "bias_" in name:
            if "lstm_cell" in name:
                b = 1.0 / math.sqrt(lstm_units)
                out[name] = (torch.rand(shape, generator=g) * 2 - 1) * b
            else:
                out[name] = torch.zeros(shape)
        else:
            fan_in = 1
            for d in shape[1:]:
                fan_in *= d
            b = (1.0 / fan_in) ** 0.5
            out[name] = (torch.rand(shape, generator=g) * 2 - 1) * b
    return out
