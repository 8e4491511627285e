#!/bin/bash
# Installs the unmodified reference (opherlieber/rltime, /root/reference) under baseline/_ref so that
# the reference arm of bench.py can run it on the GPU box (baseline/_ref is git-ignored but travels
# with the gpurun snapshot).  --no-deps: gym / opencv-python are not in the offline wheelhouse; the
# learner / replay path imports gym at module import only (oracle/stubs/gym provides those names).
# The source tree is read-only, so the build runs from a copy under /tmp.
set -e
cd "$(dirname "$0")/.."
rm -rf /tmp/rltime_ref_src baseline/_ref
cp -r "${RLTIME_REFERENCE:-/root/reference}" /tmp/rltime_ref_src
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
  --target baseline/_ref /tmp/rltime_ref_src
rm -rf /tmp/rltime_ref_src
