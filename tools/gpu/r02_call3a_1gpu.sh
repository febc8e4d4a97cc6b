#!/bin/bash
# Round 2, GPU call 3a (one GPU): re-run of the -m gpu suite on the new row-base addressing / TMA / writer code,
# CLI -e N Total MLUPS before (round-1 binaries in .r1_build) and after, stride sweep, ncu of the row-base kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
B="timeout 600 python bench.py --no-extra --no-cpu-baseline --no-e2e"
(timeout 1500 python -m pytest tests -m gpu -x -q) > $O/r02_c3a_gputests.log 2>&1; echo "pytest rc=$?"
tail -6 $O/r02_c3a_gputests.log
for S in 32 512 4096 262144; do $B --steps 100 --warmup 5 --dim 512 --stride $S > $O/r02_c3a_stride_512_$S.json 2>> $O/r02_c3a.err; done
for S in 32 256 1024 4096 65536 16777216; do $B --steps 200 --warmup 5 --stride $S > $O/r02_c3a_stride_256_$S.json 2>> $O/r02_c3a.err; done
$B --steps 60 --warmup 5 --dim 512 --precision f64 --stride 262144 > $O/r02_c3a_stride_512_f64_262144.json 2>> $O/r02_c3a.err
$B --steps 20 --warmup 5 --dim 1024 --variant 8 --stride 1048576 > $O/r02_c3a_aa_1024_s1M.json 2>> $O/r02_c3a.err
$B --steps 200 --warmup 5 --variant 16 > $O/r02_c3a_tma_256.json 2>> $O/r02_c3a.err
grep -h -o '"value": [0-9.]*, "unit": "MLUPS", "n_gpus": [0-9]*, "steps": [0-9]*' $O/r02_c3a_*.json | paste - - - - | head
ls $O/r02_c3a_*.json | tr '\n' ' '
# CLI with output: Total MLUPS of `lbmcl -d 256 -i 100 -e 20` (6 VTI files, 9 GB of text) before / after
for W in /dev/shm /tmp; do
  for V in r1 r2; do
    D=$W/lbmcl_out_$$; rm -rf $D; mkdir -p $D
    if [ $V = r1 ]; then EXE=.r1_build/host/lbmcl; else EXE=lbmcl_b200/host/lbmcl; fi
    L=$O/r02_cli_e20_${V}_$(basename $W).log
    S0=$(date +%s.%N)
    timeout 900 $EXE -D 0 -d 256 -i 100 -e 20 -s 32 -v $D -p $D > $L 2>&1
    S1=$(date +%s.%N)
    echo "process wall seconds: $(echo "$S1 - $S0" | bc)" >> $L
    echo "bytes written: $(du -sb $D | cut -f1)" >> $L
    if [ $W = /dev/shm ]; then (cd $D && md5sum *.vti) > $O/r02_cli_e20_md5_$V.log; fi
    rm -rf $D
  done
done
grep -H "Total MLUPS\|Total time\|process wall\|Kernels MLUPS" $O/r02_cli_e20_*.log
diff $O/r02_cli_e20_md5_r1.log $O/r02_cli_e20_md5_r2.log && echo "VTI files byte-identical between the round-1 and the round-2 writer"
# fp64 and a smaller cube with frequent output
for V in r1 r2; do
  D=/dev/shm/lbmcl_out_$$; rm -rf $D; mkdir -p $D
  if [ $V = r1 ]; then EXE=.r1_build/host/lbmcl; else EXE=lbmcl_b200/host/lbmcl; fi
  timeout 900 $EXE -D 0 -d 128 -i 200 -e 10 -s 32 -F -v $D -p $D > $O/r02_cli_128_f64_e10_$V.log 2>&1
  rm -rf $D
done
grep -H "Total MLUPS" $O/r02_cli_128_f64_e10_*.log
N="timeout 600 ncu --set full --clock-control none --import-source on -f"
$N -k regex:step_pull -s 5 -c 2 -o $O/r02_prof_blockrows2_512 python bench.py --steps 8 --warmup 3 --dim 512 --stride 262144 --no-extra --no-e2e --no-cpu-baseline > $O/r02_prof_blockrows2.log 2>&1
tail -5 $O/r02_c3a.err
