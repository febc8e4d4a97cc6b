#!/bin/bash
# Round 2, GPU call 1 (gpurun --gpus 2): first run of the one-launch flag transport on real NVLink peers.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
nvidia-smi -L > $O/r02_c1_gpus.log 2>&1
(timeout 300 python __graft_entry__.py smoke) > $O/r02_c1_smoke.log 2>&1; echo "smoke rc=$?"
(timeout 600 $TR --nproc-per-node 2 --master-port 29511 tools/multi_gpu_check.py) > $O/r02_c1_multi_gpu_check_n2.log 2>&1; echo "check rc=$?"
tail -20 $O/r02_c1_multi_gpu_check_n2.log
(timeout 1500 python -m pytest tests -m gpu -x -q) > $O/r02_c1_gputests_2gpu.log 2>&1; echo "pytest rc=$?"
tail -15 $O/r02_c1_gputests_2gpu.log
for K in 20 200; do
  (timeout 600 $TR --nproc-per-node 2 --master-port 29512 bench.py --gpus 2 --steps $K --warmup 5) > $O/r02_c1_bench_n2_s$K.json 2> $O/r02_c1_bench_n2_s$K.err; echo "bench n2 s$K rc=$?"
  tail -c 1500 $O/r02_c1_bench_n2_s$K.json
done
for T in token dense; do
  (timeout 600 $TR --nproc-per-node 2 --master-port 29513 bench.py --gpus 2 --steps 100 --warmup 5 --transport $T --no-extra --no-parity-gate) > $O/r02_c1_bench_n2_$T.json 2> $O/r02_c1_bench_n2_$T.err; echo "bench $T rc=$?"
done
(LBM_B200_LIB=$PWD/lbmcl_b200/csrc/liblbm_b200_tight.so timeout 600 $TR --nproc-per-node 2 --master-port 29514 bench.py --gpus 2 --steps 100 --warmup 5 --no-extra --no-parity-gate) > $O/r02_c1_bench_n2_tight.json 2> $O/r02_c1_bench_n2_tight.err; echo "bench tight rc=$?"
(timeout 600 $TR --nproc-per-node 2 --master-port 29515 bench.py --gpus 2 --steps 100 --warmup 5 --no-extra --no-parity-gate) > $O/r02_c1_bench_n2_s100.json 2> $O/r02_c1_bench_n2_s100.err; echo "bench s100 rc=$?"
(timeout 600 $TR --nproc-per-node 2 --master-port 29516 bench.py --gpus 2 --steps 100 --warmup 5 --precision f64 --dim 512 --no-extra) > $O/r02_c1_bench_n2_f64.json 2> $O/r02_c1_bench_n2_f64.err; echo "bench f64 rc=$?"
(timeout 900 python bench.py --steps 20 --warmup 5) > $O/r02_c1_bench_n1.json 2> $O/r02_c1_bench_n1.err; echo "bench n1 rc=$?"
tail -c 3000 $O/r02_c1_bench_n1.json
(timeout 300 python bench.py --steps 200 --warmup 5 --variant 32 --no-extra --no-cpu-baseline --no-e2e) > $O/r02_c1_bench_n1_nvrtc.json 2> $O/r02_c1_bench_n1_nvrtc.err; echo "bench nvrtc rc=$?"
(timeout 300 python bench.py --steps 200 --warmup 5 --no-extra --no-cpu-baseline --no-e2e) > $O/r02_c1_bench_n1_s200.json 2> $O/r02_c1_bench_n1_s200.err
(timeout 300 python bench.py --steps 100 --warmup 5 --dim 512 --stride 262144 --no-extra --no-cpu-baseline --no-e2e) > $O/r02_c1_bench_512_blockrows.json 2> $O/r02_c1_bench_512_blockrows.err
(timeout 300 python bench.py --steps 100 --warmup 5 --dim 512 --no-extra --no-cpu-baseline --no-e2e) > $O/r02_c1_bench_512_s32.json 2> $O/r02_c1_bench_512_s32.err
grep -h -o '"value": [0-9.]*, "unit": "MLUPS", "n_gpus": [0-9]*, "steps": [0-9]*' $O/r02_c1_bench_*.json
tail -3 $O/*.err | tail -40
