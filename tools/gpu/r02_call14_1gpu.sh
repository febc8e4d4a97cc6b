#!/bin/bash
# Round 2, GPU call 14 (one GPU, the last seconds of the budget): default-path parity on the final tree.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
(timeout 50 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "strict_matches or golden or in_place_variant") > gpurun_out/r02_c14_gputests.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/r02_c14_gputests.log
