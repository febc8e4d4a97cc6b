#!/bin/bash
# Round 2, GPU call 10 (one GPU): final-code check -- the whole -m gpu suite, the driver's bench command, the ncu
# launch list of it, the host program with output before (round-1 binaries rebuilt from commit a490fe0 into
# .r1_build) and after, the in-place SHIFT step at 40 registers, the TMA kernel's final defaults.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
B="timeout 300 python bench.py --no-extra --no-cpu-baseline --no-e2e"
S0=$(date +%s)
(timeout 900 python -m pytest tests -m gpu -x -q) > $O/r02_c10_gputests.log 2>&1; echo "pytest rc=$?"
tail -4 $O/r02_c10_gputests.log
echo "pytest seconds: $(( $(date +%s) - S0 ))"
(timeout 600 python bench.py --steps 20 --warmup 5) > $O/r02_bench_default.json 2> $O/r02_bench_default.err; echo "bench n1 rc=$?"
echo "bench seconds: $(( $(date +%s) - S0 ))"
# the host program with output: Total MLUPS of `lbmcl -d 256 -i 100 -e 20` (6 VTI files, 9 GB of text)
for W in /dev/shm /tmp; do
  for V in r1 r2; do
    D=$W/lbmcl_out_$$; rm -rf $D; mkdir -p $D
    if [ $V = r1 ]; then EXE=.r1_build/host/lbmcl; else EXE=lbmcl_b200/host/lbmcl; fi
    L=$O/r02_cli_e20_${V}_$(basename $W).log
    T0=$(date +%s%N)
    timeout 300 $EXE -D 0 -d 256 -i 100 -e 20 -s 32 -v $D -p $D > $L 2>&1
    T1=$(date +%s%N)
    echo "process wall ms: $(( (T1 - T0) / 1000000 ))" >> $L
    echo "bytes written: $(du -sb $D | cut -f1)" >> $L
    if [ $W = /dev/shm ]; then (cd $D && md5sum *.vti) > $O/r02_cli_e20_md5_$V.log; fi
    rm -rf $D
  done
done
grep -H "Total MLUPS\|Total time\|process wall\|Kernels MLUPS" $O/r02_cli_e20_r*.log
diff $O/r02_cli_e20_md5_r1.log $O/r02_cli_e20_md5_r2.log && echo "VTI files byte-identical between the round-1 and the round-2 writer"
for V in r1 r2; do
  D=/dev/shm/lbmcl_out_$$; rm -rf $D; mkdir -p $D
  if [ $V = r1 ]; then EXE=.r1_build/host/lbmcl; else EXE=lbmcl_b200/host/lbmcl; fi
  timeout 300 $EXE -D 0 -d 128 -i 200 -e 10 -s 32 -F -v $D -p $D > $O/r02_cli_128_f64_e10_$V.log 2>&1
  rm -rf $D
done
grep -H "Total MLUPS" $O/r02_cli_128_f64_e10_*.log
echo "cli seconds: $(( $(date +%s) - S0 ))"
# in-place variant: SHIFT step at 40 registers (6 blocks per SM, a few bytes spilled) against ptxas' 48
for D in 256 1024; do
  N=400; [ $D = 1024 ] && N=60
  python tools/step_times.py 8 $D f32 $N > $O/r02_aa_${D}_shift48.log 2>&1
  LBM_B200_LIB=$PWD/lbmcl_b200/csrc/liblbm_b200_aas6.so python tools/step_times.py 8 $D f32 $N > $O/r02_aa_${D}_shift40.log 2>&1
done
tail -n 1 $O/r02_aa_*_shift4*.log
# TMA kernel, final defaults (TX 256, 2 stages, direct stores)
$B --steps 100 --warmup 5 --variant 16 > $O/r02_tma_final_256.json 2>> $O/r02_c10.err
$B --steps 40 --warmup 5 --variant 16 --dim 512 --precision f64 > $O/r02_tma_final_512_f64.json 2>> $O/r02_c10.err
$B --steps 40 --warmup 5 --variant 16 --dim 512 > $O/r02_tma_final_512.json 2>> $O/r02_c10.err
for f in $O/r02_tma_final_*.json; do echo "$f $(grep -h -o '"value": [0-9.]*' $f | head -1)"; done
(timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/r02_launches.csv python bench.py --steps 20 --warmup 5 --no-extra --no-cpu-baseline) > $O/r02_launches_bench.log 2>&1; echo "ncu list rc=$?"
tail -3 $O/r02_c10.err
echo "total seconds: $(( $(date +%s) - S0 ))"
