"""Pins oracle/learner_oracle.py against golden results of the unmodified reference trainer
(oracle/gen_golden_learner.py).  CPU only; fp32 tolerances stated inline."""
import numpy as np
import pytest
import torch

from oracle import learner_oracle as lo
from oracle.learner_cases import CASES, batch_of, params_of, spec_of, taus_of, update_kwargs
from tests.util import load_golden


@pytest.mark.parametrize("name", sorted(CASES))
def test_learner_oracle_matches_reference(name):
    torch.set_num_threads(1)
    c = CASES[name]
    g = load_golden("learner_%s.npz" % name)
    spec = spec_of(c)
    p_online, p_target = params_of(g, "online"), params_of(g, "target")
    assert set(p_online) == set(spec.param_shapes())
    # init_params is deterministic and is what the generator used
    for k, v in spec.init_params(seed=11).items():
        np.testing.assert_array_equal(v.numpy(), p_online[k].numpy())
    opt = lo.Adam(p_online, lr=1e-3, eps=c["adam_eps"])   # train_init ignores lr (torch_trainer.py:80-83)
    dyn = lo.DynamicClip(c["clip_grad"], c["clip_dyn_alpha"]) if c.get("clip_dyn_alpha") is not None else None
    for u in range(c["updates"]):
        batch, _ = batch_of(g, c, u)
        res = lo.learner_update(spec, p_online, p_target, opt, batch, taus_of(g, u), c["gamma"],
                                double_q=c["double_q"], rnn_bootstrap=c["rnn_bootstrap"],
                                vf_eps=c["vf_eps"], clip_grad=c["clip_grad"],
                                burn_in_timesteps=c["P"], dynamic_clip=dyn, **update_kwargs(c))
        # same library, same op order: targets / loss agree to fp32 round-off
        np.testing.assert_allclose(res["targets"].numpy(), g["u%d/targets" % u], rtol=1e-6, atol=1e-6)
        np.testing.assert_allclose(float(res["loss"]), float(g["u%d/qloss" % u]), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(res["report"].numpy(), g["u%d/report" % u], rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(float(res["td_mean"]), float(g["u%d/td_mean" % u]), rtol=1e-6)
        np.testing.assert_allclose(res["grad_norm"], float(g["u%d/grad_norm" % u]), rtol=1e-5)
        for k in p_online:
            np.testing.assert_allclose(res["grads"][k].numpy(), g["u%d/grad/%s" % (u, k)],
                                       rtol=1e-4, atol=1e-7, err_msg="grad " + k)
            np.testing.assert_allclose(p_online[k].numpy(), g["u%d/after/%s" % (u, k)],
                                       rtol=1e-5, atol=1e-6, err_msg="param " + k)
