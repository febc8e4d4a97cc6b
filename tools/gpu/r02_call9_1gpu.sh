#!/bin/bash
# Round 2, GPU call 9 (one GPU): the warp-specialised TMA kernel with direct stores (3 CTAs per SM) -- parity,
# then a sweep over tile width, ring depth and resident CTAs; a 4-CTA build (56 registers) beside it.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
B="timeout 120 python bench.py --no-extra --no-cpu-baseline --no-e2e --variant 16"
S0=$(date +%s)
(LBM_SYNC_TIMEOUT_S=5 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "tma or (strict_matches and -16-) or (fast_within and -16-)") > $O/r02_c9_gputests.log 2>&1; echo "pytest rc=$?"
tail -5 $O/r02_c9_gputests.log
echo "pytest seconds: $(( $(date +%s) - S0 ))"
run() {  # name, env...
  local name=$1; shift
  (env "$@" $B --steps 100 --warmup 5 > $O/r02_tma3_$name.json 2>> $O/r02_c9.err)
  echo "$name $(grep -h -o '"value": [0-9.]*' $O/r02_tma3_$name.json | head -1) $(grep -h -o '"frac": [0-9.]*' $O/r02_tma3_$name.json | head -1)"
}
run default A=1
for TX in 256 128; do
  for NS in 2 3 4; do
    run tx${TX}_ns${NS} LBM_TMA_TX=$TX LBM_TMA_NS=$NS
  done
done
run tx256_ns2_c2 LBM_TMA_TX=256 LBM_TMA_NS=2 LBM_TMA_CTAS=2
run tx256_ns4_c2 LBM_TMA_TX=256 LBM_TMA_NS=4 LBM_TMA_CTAS=2
L4=$PWD/lbmcl_b200/csrc/liblbm_b200_tma4.so
run minb4_default LBM_B200_LIB=$L4
run minb4_tx256_ns2 LBM_B200_LIB=$L4 LBM_TMA_TX=256 LBM_TMA_NS=2
run minb4_tx128_ns2 LBM_B200_LIB=$L4 LBM_TMA_TX=128 LBM_TMA_NS=2
run minb4_tx128_ns3 LBM_B200_LIB=$L4 LBM_TMA_TX=128 LBM_TMA_NS=3
(env $B --steps 100 --warmup 5 --fast-math 1 > $O/r02_tma3_fast.json 2>> $O/r02_c9.err); grep -h -o '"value": [0-9.]*' $O/r02_tma3_fast.json | head -1
(env $B --steps 40 --warmup 5 --dim 512 --precision f64 > $O/r02_tma3_f64_512.json 2>> $O/r02_c9.err); grep -h -o '"value": [0-9.]*' $O/r02_tma3_f64_512.json | head -1
(env LBM_TMA_TX=256 $B --steps 40 --warmup 5 --dim 512 --precision f64 > $O/r02_tma3_f64_512_tx256.json 2>> $O/r02_c9.err); grep -h -o '"value": [0-9.]*' $O/r02_tma3_f64_512_tx256.json | head -1
(env $B --steps 40 --warmup 5 --dim 512 > $O/r02_tma3_f32_512.json 2>> $O/r02_c9.err); grep -h -o '"value": [0-9.]*' $O/r02_tma3_f32_512.json | head -1
echo "sweep seconds: $(( $(date +%s) - S0 ))"
N="timeout 300 ncu --set full --clock-control none --import-source on -f"
$N -k regex:step_tma -s 5 -c 1 -o $O/r02_prof_tma3 python bench.py --steps 8 --warmup 3 --variant 16 --no-extra --no-e2e --no-cpu-baseline > $O/r02_prof_tma3.log 2>&1
tail -5 $O/r02_c9.err
echo "total seconds: $(( $(date +%s) - S0 ))"
