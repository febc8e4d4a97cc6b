#!/bin/bash
# Round 2, GPU call 6 (one GPU): the aligned SHIFT step of the in-place variant (parity + A/B against the
# unaligned kernel and other register budgets) and the aligned-load A/B build of the pull kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
L=$PWD/lbmcl_b200/csrc
B="timeout 600 python bench.py --no-extra --no-cpu-baseline --no-e2e"
S0=$(date +%s)
(timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -x -q) > $O/r02_c6_gputests.log 2>&1; echo "pytest rc=$?"
tail -5 $O/r02_c6_gputests.log
# the aligned-load build of the pull kernel must be bit-identical too (block x extent >= 32 only)
(LBM_B200_LIB=$L/liblbm_b200_pal.so timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q \
   -k "(strict_matches and (d32 or d64) and (0- or 1-)) or three_kernels or 512_fp64 or large_lattice or row_base") > $O/r02_c6_gputests_pal.log 2>&1; echo "pytest pal rc=$?"
tail -5 $O/r02_c6_gputests_pal.log
echo "pytest seconds: $(( $(date +%s) - S0 ))"
# in-place variant: per-step times (odd = LOCAL, even = SHIFT)
for D in 256 1024; do
  N=400; [ $D = 1024 ] && N=60
  python tools/step_times.py 8 $D f32 $N > $O/r02_aa_${D}_aligned40.log 2>&1
  LBM_B200_LIB=$L/liblbm_b200_aas5.so python tools/step_times.py 8 $D f32 $N > $O/r02_aa_${D}_aligned48.log 2>&1
  LBM_B200_LIB=$L/liblbm_b200_aas0.so python tools/step_times.py 8 $D f32 $N > $O/r02_aa_${D}_aligned56.log 2>&1
  LBM_AA_SHIFT=unaligned python tools/step_times.py 8 $D f32 $N > $O/r02_aa_${D}_unaligned.log 2>&1
done
python tools/step_times.py 8 1024 f32 60 256,1,1 1024 > $O/r02_aa_1024_s1024_aligned40.log 2>&1
LBM_AA_SHIFT=unaligned python tools/step_times.py 8 1024 f32 60 256,1,1 1024 > $O/r02_aa_1024_s1024_unaligned.log 2>&1
python tools/step_times.py 8 512 f64 100 > $O/r02_aa_512_f64_aligned.log 2>&1
LBM_B200_LIB=$L/liblbm_b200_aas0.so python tools/step_times.py 8 512 f64 100 > $O/r02_aa_512_f64_aligned_unbounded.log 2>&1
LBM_AA_SHIFT=unaligned python tools/step_times.py 8 512 f64 100 > $O/r02_aa_512_f64_unaligned.log 2>&1
tail -n 1 $O/r02_aa_*.log
echo "aa seconds: $(( $(date +%s) - S0 ))"
# pull kernel: unaligned x +- 1 gathers (default) against aligned loads + shuffle
for rep in 1 2; do
  $B --steps 200 --warmup 5 > $O/r02_pull_default_$rep.json 2>> $O/r02_c6.err
  LBM_B200_LIB=$L/liblbm_b200_pal.so $B --steps 200 --warmup 5 > $O/r02_pull_aligned_$rep.json 2>> $O/r02_c6.err
done
$B --steps 60 --warmup 5 --dim 512 --precision f64 > $O/r02_pull_f64_default.json 2>> $O/r02_c6.err
LBM_B200_LIB=$L/liblbm_b200_pal.so $B --steps 60 --warmup 5 --dim 512 --precision f64 > $O/r02_pull_f64_aligned.json 2>> $O/r02_c6.err
$B --steps 100 --warmup 5 --dim 512 > $O/r02_pull_512_default.json 2>> $O/r02_c6.err
LBM_B200_LIB=$L/liblbm_b200_pal.so $B --steps 100 --warmup 5 --dim 512 > $O/r02_pull_512_aligned.json 2>> $O/r02_c6.err
for f in $O/r02_pull_*.json; do echo "$f $(grep -h -o '"value": [0-9.]*, "unit": "MLUPS", "n_gpus": [0-9]*, "steps": [0-9]*' $f)"; done
echo "pull seconds: $(( $(date +%s) - S0 ))"
N="timeout 300 ncu --set full --clock-control none --import-source on -f"
$N -k regex:step_aa -s 4 -c 2 -o $O/r02_prof_aa_aligned_256 python tools/step_times.py 8 256 f32 30 > $O/r02_prof_aa_aligned.log 2>&1
LBM_B200_LIB=$L/liblbm_b200_pal.so $N -k regex:step_pull -s 5 -c 2 -o $O/r02_prof_pull_aligned python bench.py --steps 8 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > $O/r02_prof_pull_aligned.log 2>&1
ls -la $O/*.ncu-rep | tail -4
tail -5 $O/r02_c6.err
echo "total seconds: $(( $(date +%s) - S0 ))"
