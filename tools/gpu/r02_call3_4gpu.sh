#!/bin/bash
# Round 2, GPU call 3 (gpurun --gpus 4): interior ranks (two faces) of the flag transport on real NVLink peers,
# N = 4 and N = 2 bench lines at --steps 20 and 200, fp64 512^3.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > $O/r02_c3_gpus.log 2>&1
(timeout 900 $TR --nproc-per-node 4 --master-port 29521 tools/multi_gpu_check.py) > $O/r02_multi_gpu_check_n4.log 2>&1; echo "check n4 rc=$?"
grep -c "bit-identical" $O/r02_multi_gpu_check_n4.log; grep -h "MISMATCH\|Error\|error" $O/r02_multi_gpu_check_n4.log | head -5
(timeout 900 python -m pytest tests/test_gpu_multiproc.py tests/test_host_cli.py tests/test_gpu_fullsize.py -m gpu -x -q) > $O/r02_c3_gputests_4gpu.log 2>&1; echo "pytest rc=$?"
tail -4 $O/r02_c3_gputests_4gpu.log
for K in 20 200; do
  (timeout 600 $TR --nproc-per-node 4 --master-port 29522 bench.py --gpus 4 --steps $K --warmup 5) > $O/r02_c3_bench_n4_s$K.json 2> $O/r02_c3_bench_n4_s$K.err; echo "bench n4 s$K rc=$?"
done
(timeout 600 $TR --nproc-per-node 4 --master-port 29523 bench.py --gpus 4 --steps 100 --warmup 5 --transport token --no-extra --no-parity-gate) > $O/r02_c3_bench_n4_token.json 2> $O/r02_c3_bench_n4_token.err
(timeout 600 $TR --nproc-per-node 4 --master-port 29524 bench.py --gpus 4 --steps 100 --warmup 5 --precision f64 --dim 512 --no-extra) > $O/r02_c3_bench_n4_f64.json 2> $O/r02_c3_bench_n4_f64.err
(timeout 600 $TR --nproc-per-node 4 --master-port 29525 bench.py --gpus 4 --steps 20 --warmup 5 --precision f64 --dim 512 --no-extra) > $O/r02_c3_bench_n4_f64_s20.json 2> $O/r02_c3_bench_n4_f64_s20.err
for K in 20 200; do
  (timeout 600 $TR --nproc-per-node 2 --master-port 29526 bench.py --gpus 2 --steps $K --warmup 5) > $O/r02_c3_bench_n2_s$K.json 2> $O/r02_c3_bench_n2_s$K.err; echo "bench n2 s$K rc=$?"
done
(timeout 600 $TR --nproc-per-node 2 --master-port 29527 bench.py --gpus 2 --steps 100 --warmup 5 --precision f64 --dim 512 --no-extra) > $O/r02_c3_bench_n2_f64.json 2> $O/r02_c3_bench_n2_f64.err
(timeout 600 python bench.py --gpus 4 --steps 100 --warmup 5 --no-e2e) > $O/r02_c3_bench_group_n4.json 2> $O/r02_c3_bench_group_n4.err
(timeout 300 $TR --nproc-per-node 4 --master-port 29528 bench.py --impl reference --gpus 4 --steps 20 --warmup 5) > $O/r02_c3_bench_ref_n4.json 2> $O/r02_c3_bench_ref_n4.err
grep -h -o '"value": [0-9.]*, "unit": "MLUPS", "n_gpus": [0-9]*, "steps": [0-9]*' $O/r02_c3_bench_*.json
ls $O/r02_c3_bench_*.json | tr '\n' ' '
tail -n 3 $O/r02_c3_*.err | tail -40
