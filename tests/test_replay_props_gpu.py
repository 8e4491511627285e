"""Size-independent properties of the CUDA replay at real frame size (4x84x84 uint8, LSTM 512):
byte-exact gather against the pooled source frames, n-step returns against a numpy recomputation,
stratified-draw ordering, weight normalisation, and the fp64 tree against the oracle tree."""
import ctypes as C
import random

import numpy as np
import pytest


@pytest.mark.gpu
def test_tree_kernels_match_oracle_tree_bit_exact():
    from oracle.replay_oracle import SumTree
    from rltime_b200 import _lib
    lib = _lib.load()
    cap = 1 << 14
    h = C.c_void_p()
    _lib.check(lib.rt_tree_create(cap, 0, C.byref(h)))
    try:
        rs = np.random.RandomState(0)
        ref = SumTree(cap)
        for rnd in range(6):
            m = int(rs.randint(1, 3000))
            idx = rs.randint(0, cap, m).astype(np.int32)
            val = np.abs(rs.randn(m)) ** rs.choice([0.6, 0.9, 3.0])
            if rnd == 3:
                val[::5] = 0.0
            for i, v in zip(idx, val):
                ref.set(int(i), float(v))
            _lib.check(lib.rt_tree_set(h, m, idx.ctypes.data, val.ctypes.data, None))
            s = C.c_double()
            _lib.check(lib.rt_tree_sum(h, C.byref(s), None))
            assert s.value == ref.root()                      # bit-exact fp64 root
            masses = rs.rand(512) * ref.root()
            out = np.empty(512, dtype=np.int32)
            _lib.check(lib.rt_tree_find(h, 512, masses.ctypes.data, out.ctypes.data, None))
            want = np.array([ref.find_prefixsum_idx(float(x)) for x in masses])
            np.testing.assert_array_equal(out, want)
    finally:
        lib.rt_tree_destroy(h)


@pytest.mark.gpu
def test_full_frame_replay_properties():
    import torch
    from rltime_b200.history import DevicePrioritizedReplayHistoryBuffer
    from rltime_b200.synthetic import SyntheticStream
    E, T, P, n, B, N = 8, 20, 4, 3, 16, 6000
    gamma = 0.99
    stream = SyntheticStream(num_envs=E, frame_shape=(4, 84, 84), num_actions=6, lstm_units=512, seed=5,
                             done_mode="bernoulli", done_p=0.03, pool=64)
    hist = DevicePrioritizedReplayHistoryBuffer(
        size=N, train_frequency=None, alpha=0.9, beta=0.6, nstep_target=n, nstep_train=T, prefix_steps=P,
        gamma=gamma, max_envs=E)
    rew, done, fidx = [], [], []          # per vector step, per env
    random.seed(3)
    rs = np.random.RandomState(9)
    steps = 2 * N // E                    # wraps the buffer twice
    try:
        for k in range(steps):
            a = stream.next_arrays()
            rew.append(a["reward"]); done.append(a["done"]); fidx.append(a["frame_idx"])
            hist.update(stream.samples_from_arrays(a))
            if k > (T + P + n) * 2 and k % 37 == 0:
                td = hist.get_train_data(B, 0.5)
                assert td is not None
                li = td["extra_data"]["loss_indices"].cpu().numpy()          # (S, B, 2)
                idx = hist.last_sampled_idxes
                w = td["extra_data"]["importance_weights"].cpu().numpy()
                assert w.max() == 1.0 and w.min() > 0 and (w == w[0:1]).all()   # one weight per sequence
                assert (li[:P] == -1).all()
                env, base = li[P, :, 0], li[P, :, 1]
                assert ((base % 10) == 0).all()                              # gap = T - T//2
                # gather: state of offset o is the next_state of offset o-1 -> pooled frame of step o-1
                x = td["states"]["x"].cpu().numpy()
                tx = td["target_states"]["x"].cpu().numpy()
                ret = td["returns"].cpu().numpy()
                msk = td["target_masks"].cpu().numpy()
                for b in range(B):
                    for t in (0, P, P + T - 1):
                        o = base[b] - P + t
                        src = stream.pool[fidx[o - 1][env[b]]] if o > 0 else stream.pool[fidx[0][env[b]]]
                        np.testing.assert_array_equal(x[t, b], src)
                        np.testing.assert_array_equal(tx[t, b], stream.pool[fidx[o + n - 1][env[b]]])
                        # n-step return / mask recomputed in the reference's arithmetic order
                        r, m = float(rew[o][env[b]]), 1 - int(done[o][env[b]])
                        for q in range(1, n):
                            if m:
                                r += (gamma ** q) * rew[o + q][env[b]]
                            if done[o + q][env[b]]:
                                m = 0
                        assert ret[t, b] == r and msk[t, b] == m
                losses = np.abs(rs.randn(T * B)).astype(np.float32)
                hist.update_losses(li[P:].reshape(-1, 2), losses.astype(np.float64))
        assert hist._lib.rt_replay_len(hist._h) == N
        # stratified draw: with uniforms all equal the drawn leaves have non-decreasing cumulative mass
        td = hist.get_train_data(B, 1.0)
        assert td is not None and len(set(hist.last_sampled_idxes)) >= B // 2
        torch.cuda.synchronize()
    finally:
        hist.close()


@pytest.mark.gpu
def test_non_integer_env_ids_round_trip():
    """Remote actor pools hand out env ids that are not integers (acting/actor_pool.py builds (actor, env)
    style ids): the buffer maps them to dense indices, echoes an integer code in loss_indices and accepts
    that code back in update_losses; every other field is identical to the same run with integer ids."""
    import random
    from oracle import scenario as sc
    from rltime_b200.history import DevicePrioritizedReplayHistoryBuffer
    from rltime_b200.synthetic import SyntheticStream
    kw = dict(size=300, train_frequency=None, alpha=0.8, beta=0.5, nstep_target=2, nstep_train=4, prefix_steps=1)
    outs = []
    for named in (False, True):
        h = DevicePrioritizedReplayHistoryBuffer(**kw, discount_function=sc.discount_function, max_envs=3)
        stream = SyntheticStream(num_envs=3, frame_shape=(1, 4, 4), num_actions=3, lstm_units=2, seed=2,
                                 done_mode="bernoulli", done_p=0.05, pool=64)
        random.seed(3)
        rs = np.random.RandomState(4)
        trace = []
        try:
            for it in range(60):
                samples = stream.next_samples()
                if named:
                    for s in samples:
                        s["env_id"] = ("actor%d" % (s["env_id"] % 2), s["env_id"])
                h.update(samples)
                td = h.get_train_data(4, 0.1)
                if td is None:
                    continue
                flat = sc.flatten_train_data(td)
                li = flat["extra_data/loss_indices"][1:].reshape(-1, 2)
                h.update_losses(li, np.abs(rs.randn(len(li))))
                flat["idx"] = np.asarray(h.last_sampled_idxes)
                trace.append(flat)
        finally:
            h.close()
        outs.append(trace)
    assert len(outs[0]) == len(outs[1]) > 20
    for a, b in zip(*outs):
        for k in a:
            if k == "extra_data/loss_indices":
                np.testing.assert_array_equal(a[k][..., 1], b[k][..., 1])      # env offsets
                assert ((b[k][1:, :, 0] >= (1 << 40)) | (b[k][1:, :, 0] == -1)).all()
            else:
                np.testing.assert_array_equal(a[k], b[k], err_msg=k)
