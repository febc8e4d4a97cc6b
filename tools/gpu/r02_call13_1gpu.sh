#!/bin/bash
# Round 2, GPU call 13 (one GPU, ~1 minute): the in-place variant with the tile order on its SHIFT launches only
# (the final rule): 1024^3 parity test and per-step times.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
(timeout 60 python -m pytest tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -x -q -k "1024_fp32_in_place or aa_variant_uses_half") > $O/r02_c13_gputests.log 2>&1; echo "pytest rc=$?"
tail -2 $O/r02_c13_gputests.log
timeout 40 python tools/step_times.py 8 1024 f32 50 > $O/r02_swz_aa_1024_shift_only.log 2>&1
tail -n 1 $O/r02_swz_aa_1024_shift_only.log
