#!/bin/bash
# Round 2, GPU call 12 (one GPU, the last 4 GPU-minutes): the tile order of the blocks (block_yz) -- parity with it
# forced on small lattices and on by default at DIM >= 512, A/B timings against the grid's own order -- then
# smoke() and a short benchmark.sh sweep (reference-format CSV + roofline column).
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
B="timeout 120 python bench.py --no-extra --no-cpu-baseline --no-e2e"
S0=$(date +%s)
(LBM_BLOCK_SWIZZLE=1,1 timeout 200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "strict_matches or in_place_variant or row_base or graph_chunks or population_dump or slab_group_on_one or flag_transport") > $O/r02_c12_gputests_swz11.log 2>&1; echo "pytest swizzle 1,1 rc=$? t=$(( $(date +%s) - S0 ))"
tail -2 $O/r02_c12_gputests_swz11.log
(timeout 200 python -m pytest tests/test_gpu_fullsize.py -m gpu -x -q) > $O/r02_c12_gputests_fullsize.log 2>&1; echo "pytest fullsize rc=$? t=$(( $(date +%s) - S0 ))"
tail -2 $O/r02_c12_gputests_fullsize.log
for S in auto 0 3,3 5,4; do
  E=""; [ $S != auto ] && E="LBM_BLOCK_SWIZZLE=$S"
  env $E python tools/step_times.py 8 1024 f32 50 > $O/r02_swz_aa_1024_$S.log 2>&1
  env $E $B --steps 60 --warmup 5 --dim 512 > $O/r02_swz_pull_512_$S.json 2>> $O/r02_c12.err
  echo "swz $S: $(tail -n 1 $O/r02_swz_aa_1024_$S.log | cut -c1-150) | pull 512: $(grep -h -o '"value": [0-9.]*' $O/r02_swz_pull_512_$S.json | head -1)"
done
for S in auto 0; do
  E=""; [ $S != auto ] && E="LBM_BLOCK_SWIZZLE=$S"
  env $E $B --steps 40 --warmup 5 --dim 512 --precision f64 > $O/r02_swz_pull_512_f64_$S.json 2>> $O/r02_c12.err
  echo "swz $S f64 512: $(grep -h -o '"value": [0-9.]*' $O/r02_swz_pull_512_f64_$S.json | head -1)"
done
LBM_BLOCK_SWIZZLE=4,4 $B --steps 200 --warmup 5 > $O/r02_swz_pull_256_4,4.json 2>> $O/r02_c12.err
$B --steps 200 --warmup 5 > $O/r02_swz_pull_256_off.json 2>> $O/r02_c12.err
echo "256: swz 4,4 $(grep -h -o '"value": [0-9.]*' $O/r02_swz_pull_256_4,4.json | head -1) off $(grep -h -o '"value": [0-9.]*' $O/r02_swz_pull_256_off.json | head -1)"
echo "ab seconds: $(( $(date +%s) - S0 ))"
(timeout 60 python __graft_entry__.py smoke) > $O/r02_c12_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $O/r02_c12_smoke.log
(cd lbmcl_b200/host && DIMS="64 128 256" STRIDES="32 full" REPS=4 OPTIMIZE="" timeout 120 ./benchmark.sh > ../../$O/r02_benchmark_sh.log 2>&1; cp benchmarks/benchmark.csv ../../$O/r02_benchmark_sweep_single.csv; cp benchmarks/benchmark_reference.csv ../../$O/r02_benchmark_sweep_single_reference_format.csv)
tail -8 $O/r02_benchmark_sh.log
echo "total seconds: $(( $(date +%s) - S0 ))"
