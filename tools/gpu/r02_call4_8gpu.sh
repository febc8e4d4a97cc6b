#!/bin/bash
# Round 2, GPU call 4 (gpurun --gpus 8): the driver's scaling run reproduced -- bench.py under torchrun at
# N = 8 with the driver's --steps 20 --warmup 5 and with --steps 200 (they must agree), fp64 512^3, the 8-rank
# parity log, and the other transports for comparison.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > $O/r02_c4_gpus.log 2>&1
for K in 20 200; do
  (timeout 600 $TR --nproc-per-node 8 --master-port 29531 bench.py --gpus 8 --steps $K --warmup 5) > $O/r02_scale_f32_n8_s$K.json 2> $O/r02_scale_f32_n8_s$K.err; echo "bench n8 s$K rc=$?"
done
(timeout 900 $TR --nproc-per-node 8 --master-port 29533 tools/multi_gpu_check.py) > $O/r02_multi_gpu_check_n8.log 2>&1; echo "check n8 rc=$?"
grep -c "bit-identical" $O/r02_multi_gpu_check_n8.log; grep -h "MISMATCH\|Error\|error" $O/r02_multi_gpu_check_n8.log | head -5
for K in 20 200; do
  (timeout 600 $TR --nproc-per-node 8 --master-port 29534 bench.py --gpus 8 --steps $K --warmup 5 --precision f64 --dim 512 --no-extra --no-parity-gate) > $O/r02_scale_f64_n8_s$K.json 2> $O/r02_scale_f64_n8_s$K.err; echo "bench f64 n8 s$K rc=$?"
done
(timeout 600 $TR --nproc-per-node 8 --master-port 29535 bench.py --gpus 8 --steps 100 --warmup 5 --transport token --no-extra --no-parity-gate) > $O/r02_scale_f32_n8_token.json 2> $O/r02_scale_f32_n8_token.err
(timeout 600 $TR --nproc-per-node 8 --master-port 29536 bench.py --gpus 8 --steps 100 --warmup 5 --transport dense --no-extra --no-parity-gate) > $O/r02_scale_f32_n8_dense.json 2> $O/r02_scale_f32_n8_dense.err
(timeout 600 python bench.py --gpus 8 --steps 100 --warmup 5 --no-e2e) > $O/r02_scale_f32_n8_group.json 2> $O/r02_scale_f32_n8_group.err
(timeout 300 $TR --nproc-per-node 8 --master-port 29537 bench.py --impl reference --gpus 8 --steps 20 --warmup 5) > $O/r02_scale_ref_n8.json 2> $O/r02_scale_ref_n8.err
(timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "across_devices") > $O/r02_c4_gputests_8gpu.log 2>&1; echo "pytest rc=$?"
tail -3 $O/r02_c4_gputests_8gpu.log
D=/dev/shm/lbmcl_out_$$; mkdir -p $D
lbmcl_b200/host/lbmcl -D 0 -d 1024 -i 100 -e 0 -G 8 > $O/r02_cli_1024_G8.log 2>&1
rm -rf $D
grep -h "Kernels MLUPS" $O/r02_cli_1024_G8.log $O/r02_cli_512_f64_G8.log
grep -h -o '"value": [0-9.]*, "unit": "MLUPS", "n_gpus": [0-9]*, "steps": [0-9]*' $O/r02_scale_*.json
ls $O/r02_scale_*.json | tr '\n' ' '
tail -n 3 $O/r02_scale_*.err | tail -40
