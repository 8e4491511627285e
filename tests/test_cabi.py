"""CPU-only: the C-ABI library builds, loads, and exports every symbol the header declares
(no compute calls without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from rltime_b200 import _lib
    return _lib


def declared_functions():
    src = open(os.path.join(ROOT, "include", "rltime_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(rt_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_functions():
    names = declared_functions()
    assert "rt_replay_create" in names and "rt_replay_sample_prioritized" in names


def test_every_declared_symbol_is_exported_and_bound(lib):
    cdll = ctypes.CDLL(lib.LIB_PATH)
    for name in declared_functions():
        assert hasattr(cdll, name), "library does not export %s" % name
        assert name in lib.SIGNATURES, "python binding lacks %s" % name
    for name in lib.SIGNATURES:
        assert name in declared_functions(), "%s bound but not declared in the header" % name


def test_struct_layouts_match_header(lib):
    # rt_replay_config: 8+4*6 -> pad to 8 -> 4 doubles ... sanity-check against a C compile
    import subprocess
    import tempfile
    prog = r'''
#include <stdio.h>
#include <stddef.h>
#include "rltime_b200.h"
int main(){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(rt_replay_config), offsetof(rt_replay_config, gamma),
  offsetof(rt_replay_config, state_field_bytes), sizeof(rt_batch), offsetof(rt_batch, returns),
  offsetof(rt_batch, target_states), sizeof(rt_model_desc), offsetof(rt_model_desc, extra_dim),
  offsetof(rt_model_desc, pre_fc_sub), sizeof(rt_train_desc), offsetof(rt_train_desc, rnn_steps_train),
  sizeof(rt_learner_io)); return 0; }
'''
    with tempfile.TemporaryDirectory() as d:
        c = os.path.join(d, "t.c")
        open(c, "w").write(prog)
        exe = os.path.join(d, "t")
        subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), c, "-o", exe])
        out = subprocess.check_output([exe]).decode().split()
    got = [ctypes.sizeof(lib.ReplayConfig), lib.ReplayConfig.gamma.offset,
           lib.ReplayConfig.state_field_bytes.offset, ctypes.sizeof(lib.Batch),
           lib.Batch.returns.offset, lib.Batch.target_states.offset, ctypes.sizeof(lib.ModelDesc),
           lib.ModelDesc.extra_dim.offset, lib.ModelDesc.pre_fc_sub.offset, ctypes.sizeof(lib.TrainDesc),
           lib.TrainDesc.rnn_steps_train.offset, ctypes.sizeof(lib.LearnerIO)]
    assert [int(x) for x in out] == got


def test_error_path_without_gpu(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from rltime_b200.history import DevicePrioritizedReplayHistoryBuffer
    with pytest.raises(lib.RtError):
        DevicePrioritizedReplayHistoryBuffer(size=100, train_frequency=None, nstep_target=1,
                                             nstep_train=1)
