#!/bin/bash
# Round 2, GPU call 11 (gpurun --gpus 2): the driver's scaling command at N = 2 on the final code, and the >= 2-GPU
# tests that a 1-GPU box skips.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 2"
S0=$(date +%s)
(timeout 300 $TR --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 5) > $O/r02_scale_f32_n2_s20.json 2> $O/r02_scale_f32_n2_s20.err; echo "bench n2 s20 rc=$? t=$(( $(date +%s) - S0 ))"
(timeout 300 python -m pytest tests/test_gpu_multiproc.py tests/test_gpu_parity.py tests/test_host_cli.py -m gpu -x -q -k "torchrun or across_devices or gpus_flag or transport") > $O/r02_c11_gputests_2gpu.log 2>&1; echo "pytest rc=$? t=$(( $(date +%s) - S0 ))"
tail -4 $O/r02_c11_gputests_2gpu.log
grep -h -o '"value": [0-9.]*, "unit": "MLUPS", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' $O/r02_scale_f32_n2_s20.json | head -1
tail -n 3 $O/r02_scale_f32_n2_s20.err
echo "total seconds: $(( $(date +%s) - S0 ))"
