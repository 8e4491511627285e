#!/bin/bash
mkdir -p gpurun_out
echo "=== trainer / learner suites"
timeout -k 10 500 python -m pytest tests/test_trainer_gpu.py tests/test_learner_gpu.py tests/test_dropin_reference_gpu.py -q -m gpu --timeout 200 2>&1 | tail -25 | cut -c1-300 | tee gpurun_out/pytest_default.log
echo "=== bench with side lines"
timeout -k 5 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_side.json 2> gpurun_out/bench_side.err
tail -2 gpurun_out/bench_side.err; python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_side.json").read().strip().splitlines()[-1])
for k in ("value", "ms_per_step", "value_long", "e2e", "config3_burnin40", "config2_cnn_iqn"):
    print(k, json.dumps(d.get(k))[:300])
PY
