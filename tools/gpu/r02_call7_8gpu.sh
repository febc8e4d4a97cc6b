#!/bin/bash
# Round 2, GPU call 7 (gpurun --gpus 8, charged 8x: kept short): the driver's scaling command at N = 8
# (--steps 20 --warmup 5) and at --steps 200 (they must agree), the NCCL-token transport for comparison,
# the 8-rank parity log of every transport, fp64 512^3, and the reference arm under torchrun.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1 --nproc-per-node 8"
S0=$(date +%s)
nvidia-smi -L > $O/r02_c7_gpus.log 2>&1
(timeout 300 $TR --master-port 29531 bench.py --gpus 8 --steps 20 --warmup 5) > $O/r02_scale_f32_n8_s20.json 2> $O/r02_scale_f32_n8_s20.err; echo "bench n8 s20 rc=$? t=$(( $(date +%s) - S0 ))"
(timeout 300 $TR --master-port 29532 bench.py --gpus 8 --steps 200 --warmup 5 --no-extra --no-parity-gate --no-e2e) > $O/r02_scale_f32_n8_s200.json 2> $O/r02_scale_f32_n8_s200.err; echo "bench n8 s200 rc=$? t=$(( $(date +%s) - S0 ))"
(timeout 300 $TR --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 --transport token --no-extra --no-e2e) > $O/r02_scale_f32_n8_token_s20.json 2> $O/r02_scale_f32_n8_token_s20.err; echo "bench n8 token s20 rc=$? t=$(( $(date +%s) - S0 ))"
(timeout 300 $TR --master-port 29534 tools/multi_gpu_check.py) > $O/r02_multi_gpu_check_n8.log 2>&1; echo "check n8 rc=$? t=$(( $(date +%s) - S0 ))"
grep -c "bit-identical" $O/r02_multi_gpu_check_n8.log; grep -h "MISMATCH\|Error\|error" $O/r02_multi_gpu_check_n8.log | head -5
(timeout 300 $TR --master-port 29535 bench.py --gpus 8 --steps 20 --warmup 5 --precision f64 --dim 512 --no-extra --no-e2e) > $O/r02_scale_f64_n8_s20.json 2> $O/r02_scale_f64_n8_s20.err; echo "bench f64 n8 s20 rc=$? t=$(( $(date +%s) - S0 ))"
(timeout 300 $TR --master-port 29536 bench.py --gpus 8 --steps 200 --warmup 5 --precision f64 --dim 512 --no-extra --no-parity-gate --no-e2e) > $O/r02_scale_f64_n8_s200.json 2> $O/r02_scale_f64_n8_s200.err; echo "bench f64 n8 s200 rc=$? t=$(( $(date +%s) - S0 ))"
(timeout 200 $TR --master-port 29537 bench.py --impl reference --gpus 8 --steps 20 --warmup 5) > $O/r02_scale_ref_n8.json 2> $O/r02_scale_ref_n8.err; echo "ref n8 rc=$? t=$(( $(date +%s) - S0 ))"
for f in $O/r02_scale_*.json; do echo "$f $(grep -h -o '"value": [0-9.]*, "unit": "MLUPS", "n_gpus": [0-9]*, "steps": [0-9]*, "warmup": [0-9]*, "ms_per_step": [0-9.]*' $f | head -1)"; done
tail -n 3 $O/r02_scale_*.err | tail -30
echo "total seconds: $(( $(date +%s) - S0 ))"
