#!/bin/bash
# Round 2, GPU call 5 (one GPU): the final-code -m gpu suite, the driver's bench command + reference arm,
# one ncu --set full capture of the default step kernel (roofline.traffic), the stride sweep and the CLI
# `-e 20` Total MLUPS.  (Calls 1-3a ran before the container was replaced; their raw outputs were lost.)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
B="timeout 600 python bench.py --no-extra --no-cpu-baseline --no-e2e"
S0=$(date +%s)
(timeout 1500 python -m pytest tests -m gpu -x -q --durations=15) > $O/r02_c5_gputests.log 2>&1; echo "pytest rc=$?"
tail -25 $O/r02_c5_gputests.log
echo "pytest seconds: $(( $(date +%s) - S0 ))"
(timeout 900 python bench.py --steps 20 --warmup 5) > $O/r02_bench_default.json 2> $O/r02_bench_default.err; echo "bench n1 rc=$?"
(timeout 900 python bench.py --impl reference --steps 20 --warmup 5) > $O/r02_bench_reference.json 2> $O/r02_bench_reference.err; echo "ref rc=$?"
echo "bench seconds: $(( $(date +%s) - S0 ))"
for S in 32 512 4096 262144; do $B --steps 100 --warmup 5 --dim 512 --stride $S > $O/r02_stride_512_$S.json 2>> $O/r02_c5.err; done
for S in 256 4096 65536 16777216; do $B --steps 200 --warmup 5 --stride $S > $O/r02_stride_256_$S.json 2>> $O/r02_c5.err; done
$B --steps 200 --warmup 5 --variant 16 > $O/r02_tma_256.json 2>> $O/r02_c5.err
$B --steps 200 --warmup 5 > $O/r02_default_256_s200.json 2>> $O/r02_c5.err
grep -h -o '"value": [0-9.]*, "unit": "MLUPS", "n_gpus": [0-9]*, "steps": [0-9]*' $O/r02_stride_*.json $O/r02_tma_256.json $O/r02_default_256_s200.json
ls $O/r02_stride_*.json | tr '\n' ' '
echo "sweep seconds: $(( $(date +%s) - S0 ))"
# CLI with output: Total MLUPS of `lbmcl -d 256 -i 100 -e 20` (6 VTI files, 9 GB of text), RAM disk and /tmp
for W in /dev/shm /tmp; do
  D=$W/lbmcl_out_$$; rm -rf $D; mkdir -p $D
  L=$O/r02_cli_e20_$(basename $W).log
  T0=$(date +%s.%N)
  timeout 600 lbmcl_b200/host/lbmcl -D 0 -d 256 -i 100 -e 20 -s 32 -v $D -p $D > $L 2>&1
  T1=$(date +%s.%N)
  echo "process wall seconds: $(echo "$T1 - $T0" | bc)" >> $L
  echo "bytes written: $(du -sb $D | cut -f1)" >> $L
  rm -rf $D
done
grep -H "Total MLUPS\|Total time\|process wall\|Kernels MLUPS" $O/r02_cli_e20_*.log
lbmcl_b200/host/lbmcl -D 0 -d 256 -i 1000 -e 0 > $O/r02_cli_256.log 2>&1
lbmcl_b200/host/lbmcl -D 0 -d 512 -i 500 -e 0 -F > $O/r02_cli_512_f64.log 2>&1
grep -h "Kernels MLUPS" $O/r02_cli_256.log $O/r02_cli_512_f64.log
echo "cli seconds: $(( $(date +%s) - S0 ))"
N="timeout 300 ncu --set full --clock-control none --import-source on -f"
$N -k regex:step_pull -s 5 -c 2 -o $O/r02_prof_default python bench.py --steps 8 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > $O/r02_prof_default.log 2>&1
$N -k regex:step_pull -s 5 -c 2 -o $O/r02_prof_blockrows_512 python bench.py --steps 8 --warmup 3 --dim 512 --stride 262144 --no-extra --no-e2e --no-cpu-baseline > $O/r02_prof_blockrows.log 2>&1
$N -k regex:step_aa -s 4 -c 2 -o $O/r02_prof_aa_256 python bench.py --steps 8 --warmup 3 --variant 8 --no-extra --no-e2e --no-cpu-baseline > $O/r02_prof_aa.log 2>&1
ls -la $O/*.ncu-rep | tail
tail -5 $O/r02_c5.err
echo "total seconds: $(( $(date +%s) - S0 ))"
