"""CPU oracle for the learner half of the hot path (plain PyTorch fp32, functional).

TEST INFRASTRUCTURE — NOT PRODUCT CODE (same import rule as oracle/replay_oracle.py).

Restates, as pure functions over a {name: tensor} parameter dict that uses the reference's
state_dict names, the arithmetic of

    SequentialModel.forward / CNN / LSTM / FC     rltime/models/torch/sequential.py:167-210,
                                                  modules/cnn.py:43-50, lstm.py:50-122, fc.py:29-36
    IQNPolicy._apply_quantile_layer / predict     rltime/policies/torch/iqn.py:67-122
    DQNPolicy.predict / _process_dueling          rltime/policies/torch/dqn.py:74-112
    IQN._get_bootstrap_target_value               rltime/training/torch/iqn.py:15-52
    TorchTrainer.calc_target_values, _vf_(un)scale rltime/training/torch/torch_trainer.py:46-78,101-147
    IQN._compute_grads (+ DQN loss helpers)       rltime/training/torch/iqn.py:54-129, dqn.py:83-130
    MultiStepTrainer._burn_in                     rltime/training/multi_step_trainer.py:90-131
    TorchTrainer.train_batch (clip + Adam)        rltime/training/torch/torch_trainer.py:177-199

The quantile fractions tau are explicit inputs (the reference draws them with torch.rand
inside the forward, iqn.py:88; the golden generator injects the same values there).
Pinned by tests/test_learner_oracle_golden.py against tests/golden/learner_*.npz produced by
oracle/gen_golden_learner.py from the unmodified reference classes.
"""
import math

import torch
import torch.nn.functional as F


# ------------------------------------------------------------------------- model spec
class ModelSpec:
    """Topology family of the hot path: CNN -> [LSTM] -> FC -> (IQN, dueling) heads."""

    def __init__(self, in_shape, conv, lstm_units, fc_size, num_actions, num_quantiles=32,
                 embedding_dim=64, dueling=True, policy="iqn", pre_fc=(), extra_dim=0):
        """pre_fc: FC modules between the CNN (or the raw 1-D observation when conv is empty) and
        the LSTM / last FC module, one list of layer sizes per module (fc.py: fc_count layers of
        fc_size); extra_dim: width of the extra 1-D feature vector of a tuple observation,
        concatenated at the LSTM input (sequential.py:146-165, extra_input_layer = first recurrent
        layer)."""
        assert policy in ("iqn", "dqn")
        self.policy = policy                      # "dqn": DQNPolicy, no quantile layer
        self.in_shape = tuple(in_shape)           # (C, H, W), or (D,) without a CNN
        self.conv = [tuple(c) for c in conv]      # (filters, kernel, stride)
        self.pre_fc = [list(m) for m in pre_fc]
        self.extra_dim = int(extra_dim)
        assert not self.extra_dim or lstm_units, "extra features are fed to the LSTM layer"
        self.lstm_units = int(lstm_units)         # 0 = no recurrent layer
        self.fc_size = int(fc_size)
        self.num_actions = int(num_actions)
        self.num_quantiles = int(num_quantiles)
        self.embedding_dim = int(embedding_dim)
        self.dueling = bool(dueling)

    @property
    def conv_out(self):
        if not self.conv:
            return self.in_shape
        c, h, w = self.in_shape
        for f, k, s in self.conv:
            h = (h - k) // s + 1
            w = (w - k) // s + 1
            c = f
        return c, h, w

    @property
    def conv_feat(self):
        n = 1
        for d in self.conv_out:
            n *= d
        return n

    @property
    def feat(self):
        """Width of the trunk features that reach the LSTM / the last FC module."""
        return self.pre_fc[-1][-1] if self.pre_fc else self.conv_feat

    # module indices inside SequentialModel.layers
    @property
    def pre_fc_index(self):
        return 1 if self.conv else 0

    @property
    def lstm_index(self):
        return self.pre_fc_index + len(self.pre_fc)

    @property
    def fc_layer_index(self):
        return self.lstm_index + (1 if self.lstm_units else 0)

    @property
    def quantile_dim(self):
        # injection_layer=-1: the quantile layer multiplies the input of the last (FC) layer
        return self.lstm_units if self.lstm_units else self.feat

    def param_shapes(self):
        """Reference state_dict names -> shapes, in nn.Module registration order."""
        shapes = {}
        cin = self.in_shape[0]
        for i, (f, k, s) in enumerate(self.conv):
            shapes["model.layers.0.layers.%d.weight" % i] = (f, cin, k, k)
            shapes["model.layers.0.layers.%d.bias" % i] = (f,)
            cin = f
        d = self.conv_feat
        for m, sizes in enumerate(self.pre_fc):
            for j, sz in enumerate(sizes):
                shapes["model.layers.%d.layers.%d.0.weight" % (self.pre_fc_index + m, j)] = (sz, d)
                shapes["model.layers.%d.layers.%d.0.bias" % (self.pre_fc_index + m, j)] = (sz,)
                d = sz
        if self.lstm_units:
            u = self.lstm_units
            ln = "model.layers.%d.lstm_cell." % self.lstm_index
            shapes[ln + "weight_ih"] = (4 * u, self.feat + self.extra_dim)
            shapes[ln + "weight_hh"] = (4 * u, u)
            shapes[ln + "bias_ih"] = (4 * u,)
            shapes[ln + "bias_hh"] = (4 * u,)
        li = self.fc_layer_index
        shapes["model.layers.%d.layers.0.0.weight" % li] = (self.fc_size, self.quantile_dim)
        shapes["model.layers.%d.layers.0.0.bias" % li] = (self.fc_size,)
        shapes["out_layer.weight"] = (self.num_actions, self.fc_size)
        shapes["out_layer.bias"] = (self.num_actions,)
        if self.dueling:
            shapes["value_hidden_layer.weight"] = (self.fc_size, self.quantile_dim)
            shapes["value_hidden_layer.bias"] = (self.fc_size,)
            shapes["value_layer.weight"] = (1, self.fc_size)
            shapes["value_layer.bias"] = (1,)
        if self.policy == "iqn":
            shapes["quantile_layer.weight"] = (self.quantile_dim, self.embedding_dim)
            shapes["quantile_layer.bias"] = (self.quantile_dim,)
        return shapes

    def init_params(self, seed=0):
        """U(+-sqrt(1/fan_in)) weights, zero biases (models/torch/utils.py:6-25); the LSTM
        biases use torch's default U(+-1/sqrt(hidden)) (lstm.py:47-48 leaves them)."""
        g = torch.Generator().manual_seed(seed)
        params = {}
        for name, shp in self.param_shapes().items():
            if name.endswith("bias") or "bias_" in name:
                if "lstm_cell" in name:
                    b = 1.0 / math.sqrt(self.lstm_units)
                    params[name] = (torch.rand(shp, generator=g) * 2 - 1) * b
                else:
                    params[name] = torch.zeros(shp)
            else:
                fan_in = 1
                for d in shp[1:]:
                    fan_in *= d
                b = (1.0 / fan_in) ** 0.5
                params[name] = (torch.rand(shp, generator=g) * 2 - 1) * b
        return params


# ------------------------------------------------------------------------- forward
def cnn_forward(spec, p, x_u8):
    if not spec.conv:
        return x_u8.float().reshape(x_u8.shape[0], -1)     # fc.py:30 flattens whatever arrives
    x = x_u8.float() * (1.0 / 255.0)                       # cnn.py:44-45
    for i, (f, k, s) in enumerate(spec.conv):
        x = F.conv2d(x, p["model.layers.0.layers.%d.weight" % i],
                     p["model.layers.0.layers.%d.bias" % i], stride=s)
        x = F.relu(x)
    return x.reshape(x.shape[0], -1)


def pre_fc_forward(spec, p, x):
    """FC modules in front of the LSTM / last FC module (fc.py:29-36: linear + ReLU per layer)."""
    for m, sizes in enumerate(spec.pre_fc):
        for j in range(len(sizes)):
            n = "model.layers.%d.layers.%d.0." % (spec.pre_fc_index + m, j)
            x = F.relu(F.linear(x, p[n + "weight"], p[n + "bias"]))
    return x


def split_obs(x):
    """TorchModel._get_inputs (torch_model.py:58-65): (main observation, extra features or None)."""
    if isinstance(x, (tuple, list)):
        return x[0], torch.cat([e.float() for e in x[1:]], dim=-1)
    return x, None


def lstm_forward(spec, p, x, hx, cx, initials, timesteps):
    """lstm.py:50-122 for the single-sample case (IQN injected after the LSTM).
    x: (T*B, feat) time-major; hx/cx: (T*B, U) of which only t=0 is used; initials (T*B,)."""
    U = spec.lstm_units
    B = x.shape[0] // timesteps
    h = hx.view(timesteps, B, U)[0]
    c = cx.view(timesteps, B, U)[0]
    xs = x.view(timesteps, B, -1)
    ini = initials.view(timesteps, B)
    ln = "model.layers.%d.lstm_cell." % spec.lstm_index
    w_ih, w_hh = p[ln + "weight_ih"], p[ln + "weight_hh"]
    b_ih, b_hh = p[ln + "bias_ih"], p[ln + "bias_hh"]
    outs = []
    for t in range(timesteps):
        keep = (1 - ini[t]).unsqueeze(-1)
        h = h * keep
        c = c * keep
        gates = F.linear(xs[t], w_ih, b_ih) + F.linear(h, w_hh, b_hh)
        i, f, g, o = gates.chunk(4, dim=1)
        c = torch.sigmoid(f) * c + torch.sigmoid(i) * torch.tanh(g)
        h = torch.sigmoid(o) * torch.tanh(c)
        outs.append(h)
    return torch.cat(outs), (h, c)


def quantile_layer(spec, p, x, taus):
    """iqn.py:67-106.  x (M, D) -> (M*Nq, D); taus (M*Nq,)."""
    Nq = spec.num_quantiles
    xt = x.repeat_interleave(Nq, dim=0)
    rng = torch.arange(1, spec.embedding_dim + 1, dtype=torch.float32)
    qn = taus.unsqueeze(1).repeat([1, spec.embedding_dim])
    qn = torch.cos(rng * math.pi * qn)
    qn = F.relu(F.linear(qn, p["quantile_layer.weight"], p["quantile_layer.bias"]))
    return xt * qn


def trunk_forward(spec, p, states, timesteps):
    """CNN (+LSTM) part shared by predict and burn-in.  Returns (features (M, D), last (h, c))."""
    main, extra = split_obs(states["x"])
    x = pre_fc_forward(spec, p, cnn_forward(spec, p, main))
    last = None
    if spec.lstm_units:
        if extra is not None:                              # sequential.py:146-165,193-195
            x = torch.cat([x, extra.reshape(extra.shape[0], -1)], dim=-1)
        ls = states["layer%d_state" % spec.lstm_index]
        x, last = lstm_forward(spec, p, x, ls["hx"], ls["cx"], ls["initials"], timesteps)
    return x, last


def predict(spec, p, states, timesteps, taus):
    """IQNPolicy.predict: (values (M, Nq, A), last LSTM state)."""
    x, last = trunk_forward(spec, p, states, timesteps)
    xq = quantile_layer(spec, p, x, taus)                              # layer_inputs[-1]
    li = spec.fc_layer_index
    hdn = F.relu(F.linear(xq, p["model.layers.%d.layers.0.0.weight" % li],
                          p["model.layers.%d.layers.0.0.bias" % li]))
    adv = F.linear(hdn, p["out_layer.weight"], p["out_layer.bias"])
    adv = adv.view(-1, spec.num_quantiles, spec.num_actions)
    if spec.dueling:                                                   # dqn.py:74-87
        v = F.relu(F.linear(xq, p["value_hidden_layer.weight"], p["value_hidden_layer.bias"]))
        v = F.linear(v, p["value_layer.weight"], p["value_layer.bias"])
        v = v.view(-1, spec.num_quantiles, 1)
        out = v + adv - adv.mean(2, keepdim=True)
    else:
        out = adv
    return out, last


def predict_dqn(spec, p, states, timesteps):
    """DQNPolicy.predict (dqn.py:96-112): (q-values (M, A), last LSTM state)."""
    x, last = trunk_forward(spec, p, states, timesteps)               # layer_inputs[-1]
    li = spec.fc_layer_index
    hdn = F.relu(F.linear(x, p["model.layers.%d.layers.0.0.weight" % li],
                          p["model.layers.%d.layers.0.0.bias" % li]))
    adv = F.linear(hdn, p["out_layer.weight"], p["out_layer.bias"])
    if spec.dueling:                                                   # dqn.py:74-87
        v = F.relu(F.linear(x, p["value_hidden_layer.weight"], p["value_hidden_layer.bias"]))
        v = F.linear(v, p["value_layer.weight"], p["value_layer.bias"])
        return v + adv - adv.mean(1, keepdim=True), last
    return adv, last


# ------------------------------------------------------------------------- targets
def vf_scale(x, eps):
    if not eps:
        return x
    return torch.sign(x) * (torch.sqrt(torch.abs(x) + 1) - 1) + eps * x   # torch_trainer.py:46-52


def vf_unscale(sx, eps):
    if not eps:
        return sx
    sx = sx.double()                                                      # torch_trainer.py:61-78
    a = torch.abs(sx)
    x = a / eps - ((1 / (2. * (eps ** 2))) * torch.sqrt(4 * eps * a + (2. * eps + 1) ** 2)) + \
        (2. * eps + 1) / (2. * (eps ** 2))
    x = x * torch.sign(sx)
    return x.float()


def bootstrap_target(spec, p_online, p_target, target_states, timesteps, taus_target,
                     taus_select, double_q, margin_out=None):
    """iqn.py:15-52."""
    tq, _ = predict(spec, p_target, target_states, timesteps, taus_target)
    sel_p = p_online if double_q else p_target
    sq, _ = predict(spec, sel_p, target_states, timesteps, taus_select)
    act = sq.mean(1).argmax(dim=-1, keepdim=True)                      # (M, 1)
    act = act.unsqueeze(1).repeat([1, spec.num_quantiles, 1])
    if margin_out is not None and sq.shape[-1] > 1:
        # test aid: gap between the best and second-best selection value per row (a row whose gap is
        # below the arithmetic noise has no well-defined argmax to compare against)
        top2 = sq.mean(1).topk(2, dim=-1).values
        margin_out.append(top2[:, 0] - top2[:, 1])
    return torch.gather(tq, dim=-1, index=act).squeeze(-1)             # (M, Nq)


def bootstrap_target_dqn(spec, p_online, p_target, target_states, timesteps, double_q):
    """dqn.py:52-71."""
    tq, _ = predict_dqn(spec, p_target, target_states, timesteps)
    sq = predict_dqn(spec, p_online, target_states, timesteps)[0] if double_q else tq
    act = sq.argmax(dim=-1, keepdim=True)
    return tq.gather(dim=-1, index=act).squeeze(-1)                    # (M,)


def calc_targets(returns, boot, target_masks, nsteps, gamma, vf_eps):
    """torch_trainer.py:101-147 with make_tensor's float32 casts (models/torch/utils.py:95-121)."""
    returns, masks, nsteps = returns.float(), target_masks.float(), nsteps.float()
    if boot.dim() == 2:                                               # distributional: (M, Nq)
        returns, masks, nsteps = returns.unsqueeze(-1), masks.unsqueeze(-1), nsteps.unsqueeze(-1)
    boot = vf_unscale(boot, vf_eps)
    return vf_scale(returns + (gamma ** nsteps) * boot * masks, vf_eps)


# ------------------------------------------------------------------------- loss
def iqn_loss(spec, p, states, targets, actions, weights, timesteps, taus, kappa=1.0,
             aggregation="mean", timestep_aggregation=None):
    """iqn.py:54-129 + dqn.py:83-130.  Returns (loss scalar, reported per-row |td| mean,
    td_mean scalar)."""
    M = targets.shape[0]
    q, _ = predict(spec, p, states, timesteps, taus)                   # (M, Nq, A)
    Nq = q.shape[1]
    idx = actions.long().view(M, 1, 1).repeat([1, Nq, 1])
    theta = torch.gather(q, dim=-1, index=idx).squeeze(-1)             # (M, Nq)
    td = targets.unsqueeze(2) - theta.unsqueeze(1)                     # (M, Nq', Nq)
    a = torch.abs(td)
    huber = torch.where(a <= kappa, 0.5 * td.pow(2), kappa * (a - 0.5 * kappa))
    tau = taus.view(M, Nq).unsqueeze(1).repeat([1, targets.shape[1], 1])
    under = (td < 0).float().detach()
    loss = (torch.abs(tau - under) * huber / kappa).sum(2).mean(1)     # (M,)
    report = a.mean(1).mean(1)
    if weights is not None:
        loss = loss * weights.float()
    agg = {"mean": torch.mean, "sum": torch.sum}
    if timestep_aggregation:
        loss = agg[timestep_aggregation](loss.view(timesteps, -1), dim=0)
    loss = agg[aggregation](loss)
    return loss, report, a.mean()


def dqn_loss(spec, p, states, targets, actions, weights, timesteps, kappa=1.0, loss_mode="huber",
             aggregation="mean", timestep_aggregation=None):
    """dqn.py:126-160 (+ :83-124).  Returns (loss scalar, reported SIGNED td errors, mean chosen q)."""
    q, _ = predict_dqn(spec, p, states, timesteps)
    chosen = torch.gather(q, dim=-1, index=actions.long().unsqueeze(-1)).squeeze(-1)
    td = chosen - targets
    if loss_mode == "mse":
        loss = td.pow(2)
    else:
        a = torch.abs(td)
        loss = torch.where(a <= kappa, 0.5 * td.pow(2), kappa * (a - 0.5 * kappa))
    if weights is not None:
        loss = loss * weights.float()
    agg = {"mean": torch.mean, "sum": torch.sum}
    if timestep_aggregation:
        loss = agg[timestep_aggregation](loss.view(timesteps, -1), dim=0)
    loss = agg[aggregation](loss)
    return loss, td, chosen.mean()


# ------------------------------------------------------------------------- burn-in
def burn_in(spec, p, states, burn_in_timesteps):
    """multi_step_trainer.py:90-131: forward the first P time-steps and overwrite the stored
    LSTM state of step P with the produced one (masked by that step's initials, lstm.py:150-152).
    Mutates states['layer1_state']['hx'/'cx'][P] and returns nothing; callers slice [P:]."""
    P = burn_in_timesteps
    flat = lambda v: v[:P].reshape((-1,) + v.shape[2:])
    x = states["x"]
    sub = {"x": tuple(flat(v) for v in x) if isinstance(x, (tuple, list)) else flat(x)}
    key = "layer%d_state" % spec.lstm_index
    ls = states[key]
    sub[key] = {k: flat(v) for k, v in ls.items()}
    _, (h, c) = trunk_forward(spec, p, sub, P)
    keep = (1 - ls["initials"][P]).unsqueeze(-1)
    ls["hx"][P] = h * keep
    ls["cx"][P] = c * keep


# ------------------------------------------------------------------------- optimiser
def grad_norm(grads):
    """torch_policy.py:70-78: python sum of per-parameter squared L2 norms."""
    total = 0.0
    for g in grads.values():
        total += float(g.norm(2)) ** 2
    return total ** 0.5


def clip_grads(grads, max_norm):
    """torch.nn.utils.clip_grad_norm_ (norm of per-tensor norms; coef clamped to 1)."""
    norms = torch.stack([g.norm(2) for g in grads.values()])
    total = norms.norm(2)
    coef = torch.clamp(max_norm / (total + 1e-6), max=1.0)
    return {k: g * coef for k, g in grads.items()}


class Adam:
    """torch.optim.Adam defaults as used at torch_trainer.py:82-83 (lr 1e-3 unless set_lr,
    betas (0.9, 0.999), eps=adam_epsilon, no weight decay, no amsgrad)."""

    def __init__(self, params, lr=1e-3, eps=1e-8, betas=(0.9, 0.999)):
        self.lr, self.eps, self.b1, self.b2 = lr, eps, betas[0], betas[1]
        self.m = {k: torch.zeros_like(v) for k, v in params.items()}
        self.v = {k: torch.zeros_like(v) for k, v in params.items()}
        self.t = 0

    def step(self, params, grads):
        self.t += 1
        bc1 = 1 - self.b1 ** self.t
        bc2 = 1 - self.b2 ** self.t
        step_size = self.lr / bc1
        for k in params:
            g = grads[k]
            self.m[k].mul_(self.b1).add_(g, alpha=1 - self.b1)
            self.v[k].mul_(self.b2).addcmul_(g, g, value=1 - self.b2)
            denom = (self.v[k].sqrt() / math.sqrt(bc2)).add_(self.eps)
            params[k].addcdiv_(self.m[k], denom, value=-step_size)


class DynamicClip:
    """torch_trainer.py:153-175: clip value = clip_grad x EMA(alpha) of the gradient norm."""

    def __init__(self, clip_grad, alpha):
        self.clip_grad, self.alpha, self.ma = clip_grad, alpha, None

    def value(self, cur):
        self.ma = cur if self.ma is None else self.ma * self.alpha + cur * (1 - self.alpha)
        return self.ma * self.clip_grad


def learner_update(spec, p_online, p_target, opt, batch, taus, gamma, double_q=True,
                   rnn_bootstrap=True, vf_eps=None, kappa=1.0, clip_grad=None,
                   burn_in_timesteps=0, aggregation="mean", timestep_aggregation=None,
                   loss_mode="huber", dynamic_clip=None, rnn_steps_train=None):
    """One full learner update on a (S, B, ...) time-major batch (multi_step_trainer.py:
    278-340): burn-in -> targets -> loss/grads -> clip -> Adam.  `taus` = dict with
    'burn_online', 'burn_target' (ignored values, forwards still draw), 'target', 'select',
    'train' tensors.  Mutates p_online / opt.  Returns diagnostics."""
    import copy
    batch = copy.deepcopy(batch)
    P = burn_in_timesteps
    S, B = batch["returns"].shape
    T = S - P
    R = rnn_steps_train or T      # LSTM sequence length of the target / training passes
    if P:
        with torch.no_grad():
            burn_in(spec, p_online, batch["states"], P)
            if rnn_bootstrap:
                burn_in(spec, p_target, batch["target_states"], P)

    def cut(tree):
        if isinstance(tree, dict):
            return {k: cut(v) for k, v in tree.items()}
        if isinstance(tree, (tuple, list)):
            return tuple(cut(v) for v in tree)
        return tree[P:].reshape((-1,) + tree.shape[2:])
    states, tstates = cut(batch["states"]), cut(batch["target_states"])
    returns, masks, nsteps = cut(batch["returns"]), cut(batch["target_masks"]), cut(batch["nsteps"])
    actions = cut(batch["actions"])
    weights = cut(batch["importance_weights"]) if batch.get("importance_weights") is not None else None
    dqn = spec.policy == "dqn"
    margins = []
    with torch.no_grad():
        if dqn:
            boot = bootstrap_target_dqn(spec, p_online, p_target, tstates,
                                        R if rnn_bootstrap else 1, double_q)
        else:
            boot = bootstrap_target(spec, p_online, p_target, tstates, R if rnn_bootstrap else 1,
                                    taus["target"], taus["select"], double_q, margin_out=margins)
        targets = calc_targets(returns, boot, masks, nsteps, gamma, vf_eps)
    leaf = {k: v.detach().clone().requires_grad_(True) for k, v in p_online.items()}
    if dqn:
        loss, report, td_mean = dqn_loss(spec, leaf, states, targets, actions, weights, R, kappa,
                                         loss_mode, aggregation, timestep_aggregation)
    else:
        loss, report, td_mean = iqn_loss(spec, leaf, states, targets, actions, weights, R,
                                         taus["train"], kappa, aggregation, timestep_aggregation)
    loss.backward()
    grads = {k: v.grad for k, v in leaf.items()}
    gn = grad_norm(grads)
    if clip_grad is not None:
        grads = clip_grads(grads, dynamic_clip.value(gn) if dynamic_clip is not None else clip_grad)
    with torch.no_grad():
        opt.step(p_online, grads)
    return {"loss": loss.detach(), "report": report.detach(), "td_mean": td_mean.detach(),
            "grad_norm": gn, "targets": targets, "grads": grads,
            "select_margin": margins[0] if margins else None}
