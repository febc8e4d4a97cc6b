#!/bin/bash
# Round 2, GPU call 2 (one GPU): full -m gpu suite, bench line, occupancy / NVRTC / addressing A-B runs,
# CLI -e N before/after (round-1 binaries in .r1_build), ncu launch list and full captures.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
O=gpurun_out
mkdir -p $O
B="timeout 600 python bench.py --no-extra --no-cpu-baseline --no-e2e"
L=$PWD/lbmcl_b200/csrc
df -h /tmp /dev/shm > $O/r02_c2_df.log 2>&1; nproc >> $O/r02_c2_df.log; free -g >> $O/r02_c2_df.log
(timeout 1500 python -m pytest tests -m gpu -x -q) > $O/r02_c2_gputests.log 2>&1; echo "pytest rc=$?"
tail -6 $O/r02_c2_gputests.log
(timeout 900 python bench.py --steps 20 --warmup 5) > $O/r02_c2_bench_n1.json 2> $O/r02_c2_bench_n1.err; echo "bench n1 rc=$?"
(timeout 900 python bench.py --impl reference --steps 20 --warmup 5) > $O/r02_c2_bench_ref.json 2> $O/r02_c2_bench_ref.err; echo "ref rc=$?"
# A/B: occupancy targets, no-allocate gathers, NVRTC specialisation (separate short runs, 200 steps each)
for rep in 1 2; do
$B --steps 200 --warmup 5 > $O/r02_c2_ab_default_$rep.json 2>> $O/r02_c2_ab.err
LBM_B200_LIB=$L/liblbm_b200_occ.so $B --steps 200 --warmup 5 > $O/r02_c2_ab_occ8_$rep.json 2>> $O/r02_c2_ab.err
LBM_B200_LIB=$L/liblbm_b200_noalloc.so $B --steps 200 --warmup 5 > $O/r02_c2_ab_noalloc_$rep.json 2>> $O/r02_c2_ab.err
$B --steps 200 --warmup 5 --variant 32 > $O/r02_c2_ab_nvrtc_$rep.json 2>> $O/r02_c2_ab.err
$B --steps 60 --warmup 5 --dim 512 --precision f64 > $O/r02_c2_ab_f64_default_$rep.json 2>> $O/r02_c2_ab.err
LBM_B200_LIB=$L/liblbm_b200_occ.so $B --steps 60 --warmup 5 --dim 512 --precision f64 > $O/r02_c2_ab_f64_occ4_$rep.json 2>> $O/r02_c2_ab.err
$B --steps 60 --warmup 5 --dim 512 --precision f64 --variant 32 > $O/r02_c2_ab_f64_nvrtc_$rep.json 2>> $O/r02_c2_ab.err
done
$B --steps 200 --warmup 5 --fast-math 1 > $O/r02_c2_ab_fast.json 2>> $O/r02_c2_ab.err
$B --steps 200 --warmup 5 --fast-math 1 --variant 32 > $O/r02_c2_ab_fast_nvrtc.json 2>> $O/r02_c2_ab.err
# addressing: DIM < stride < cells (row bases) against stride 32 / DIM at 512^3 fp32
for S in 32 512 4096 262144; do $B --steps 100 --warmup 5 --dim 512 --stride $S > $O/r02_c2_stride_512_$S.json 2>> $O/r02_c2_ab.err; done
$B --steps 200 --warmup 5 --stride 4096 > $O/r02_c2_stride_256_4096.json 2>> $O/r02_c2_ab.err
$B --steps 200 --warmup 5 --stride 65536 > $O/r02_c2_stride_256_65536.json 2>> $O/r02_c2_ab.err
grep -h -o '"value": [0-9.]*, "unit": "MLUPS", "n_gpus": [0-9]*, "steps": [0-9]*' $O/r02_c2_ab_*.json $O/r02_c2_stride_*.json | paste - - - - | head -40
ls $O/r02_c2_ab_*.json $O/r02_c2_stride_*.json | tr '\n' ' '
# CLI with output: Total MLUPS of `lbmcl -d 256 -i 100 -e 20` (6 VTI files, 9 GB of text) before / after, to RAM disk and to /tmp
for W in /dev/shm /tmp; do
  for V in r1 r2; do
    D=$W/lbmcl_out_$$; rm -rf $D; mkdir -p $D
    if [ $V = r1 ]; then EXE=.r1_build/host/lbmcl; else EXE=lbmcl_b200/host/lbmcl; fi
    /usr/bin/time -v timeout 900 $EXE -D 0 -d 256 -i 100 -e 20 -s 32 -v $D -p $D > $O/r02_c2_cli_${V}_$(basename $W).log 2>&1
    du -sb $D | cut -f1 >> $O/r02_c2_cli_${V}_$(basename $W).log
    if [ $V = r2 ] && [ $W = /dev/shm ]; then md5sum $D/*.vti > $O/r02_c2_cli_md5_r2.log; fi
    if [ $V = r1 ] && [ $W = /dev/shm ]; then md5sum $D/*.vti > $O/r02_c2_cli_md5_r1.log; fi
    rm -rf $D
  done
done
grep -h "Total MLUPS\|Total time\|Elapsed (wall" $O/r02_c2_cli_*.log
diff <(cut -d' ' -f1 $O/r02_c2_cli_md5_r1.log) <(cut -d' ' -f1 $O/r02_c2_cli_md5_r2.log) && echo "VTI files byte-identical between round-1 and round-2 writers"
# -e 0 CLI lines for config 3 / 4 (statistics line on stderr)
lbmcl_b200/host/lbmcl -D 0 -d 256 -i 1000 -e 0 > $O/r02_c2_cli_256.log 2>&1
lbmcl_b200/host/lbmcl -D 0 -d 512 -i 500 -e 0 -F > $O/r02_c2_cli_512_f64.log 2>&1
lbmcl_b200/host/lbmcl -D 0 -d 1024 -i 200 -e 0 -A > $O/r02_c2_cli_1024_aa.log 2>&1
grep -h "Kernels MLUPS" $O/r02_c2_cli_256.log $O/r02_c2_cli_512_f64.log $O/r02_c2_cli_1024_aa.log
# ncu: launch list of the bench command, then full captures
(timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/r02_launches.csv python bench.py --steps 20 --warmup 5) > $O/r02_launches_bench.log 2>&1; echo "ncu list rc=$?"
N="timeout 600 ncu --set full --clock-control none --import-source on -f"
$N -k regex:step_pull -s 5 -c 2 -o $O/r02_prof_default python bench.py --steps 8 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > $O/r02_prof_default.log 2>&1
$N -k regex:lbm_step_spec -s 5 -c 2 -o $O/r02_prof_nvrtc python bench.py --steps 8 --warmup 3 --variant 32 --no-extra --no-e2e --no-cpu-baseline > $O/r02_prof_nvrtc.log 2>&1
$N -k regex:step_pull -s 5 -c 2 -o $O/r02_prof_blockrows_512 python bench.py --steps 8 --warmup 3 --dim 512 --stride 262144 --no-extra --no-e2e --no-cpu-baseline > $O/r02_prof_blockrows.log 2>&1
LBM_B200_LIB=$L/liblbm_b200_occ.so $N -k regex:step_pull -s 5 -c 2 -o $O/r02_prof_occ8 python bench.py --steps 8 --warmup 3 --no-extra --no-e2e --no-cpu-baseline > $O/r02_prof_occ8.log 2>&1
$N -k regex:step_pull -s 5 -c 2 -o $O/r02_prof_f64_512 python bench.py --steps 8 --warmup 3 --dim 512 --precision f64 --no-extra --no-e2e --no-cpu-baseline > $O/r02_prof_f64.log 2>&1
ls -la $O/*.ncu-rep | tail
tail -5 $O/r02_c2_ab.err
