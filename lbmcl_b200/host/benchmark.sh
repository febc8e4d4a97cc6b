#!/bin/bash
# Parameter sweep of the host program, the B200 counterpart of the reference's benchmark.sh:
# same protocol (reference benchmark.sh:8-15, 60-107: -i 50 -e 0, REPS repetitions per configuration,
# the ';'-separated statistics line of stderr appended to benchmarks/stats.csv) and the same
# aggregation (benchmark.sh:109-176: per configuration drop the fastest and the slowest run by total
# time and average the rest) into benchmarks/benchmark.csv.
# Output: benchmarks/benchmark.csv in the extended format (one line per configuration incl. achieved GB/s and
# the fraction of the HBM roofline, PEAK_GBS) and benchmarks/benchmark_reference.csv in the reference's own
# 9-column format (see aggregate.awk).
# The work-group size is a hint on this path (the library uses x-major rows); LWS="8,8,8 32,8,1 256,1,1" sweeps
# CUDA block shapes anyway (the reference's lws sweep, benchmark.sh:60-107) by running with LBM_EXACT_BLOCK=1.
#
#   ./benchmark.sh                      # fp32, dims 8..256
#   PRECISION=double DIMS="64 128 256 512" GPUS=1 ./benchmark.sh
#   DIMS=256 STRIDES=32 LWS="256,1,1 128,2,1 64,4,1 32,8,1 8,8,4" ./benchmark.sh
set -e
cd "$(dirname "$0")"

BENCHMARK_DIR=./benchmarks
LOG=$BENCHMARK_DIR/stats.csv
OUT=$BENCHMARK_DIR/benchmark.csv
OUT_REF=$BENCHMARK_DIR/benchmark_reference.csv
PRECISION=${PRECISION:-single}
DEVICE=${DEVICE:-0}
GPUS=${GPUS:-1}
VISCOSITY=0.0089
VELOCITY=0.05
ITERATIONS=${ITERATIONS:-50}
EVERY=0
REPS=${REPS:-10}
DIMS=${DIMS:-"8 16 32 64 128 256"}
STRIDES=${STRIDES:-"1 8 16 32 64 128 full"}
OPTIMIZE=${OPTIMIZE:-"-o"}          # the reference sweep runs with -o; set OPTIMIZE="" for strict kernels
EXTRA=${EXTRA:-}                    # e.g. EXTRA=-A for the in-place AA kernels
LWS=${LWS:-}                        # block shapes to sweep ("x,y,z ..."); empty = the library's choice only
PEAK_GBS=${PEAK_GBS:-$(python3 -c "import json;print(json.load(open('../../MEASURED_PEAKS.json'))['hbm_gbs'])" 2>/dev/null || echo 6550.7)}

mkdir -p $BENCHMARK_DIR
rm -f $LOG $OUT $OUT_REF
make -s lbmcl

FLAGS="$OPTIMIZE $EXTRA"
[ "$PRECISION" = double ] && FLAGS="$FLAGS -F"
[ "$GPUS" != 1 ] && FLAGS="$FLAGS -G $GPUS"

for d in $DIMS; do
    for s in $STRIDES; do
        [ "$s" = full ] && s=$((d * d * d))
        [ "$s" -gt $((d * d * d)) ] && continue
        for w in ${LWS:-default}; do
            for r in $(seq 1 "$REPS"); do
                if [ "$w" = default ]; then
                    ./lbmcl -D "$DEVICE" -d "$d" -n $VISCOSITY -u $VELOCITY -i "$ITERATIONS" -e $EVERY -s "$s" $FLAGS \
                        2>> $LOG > /dev/null
                else
                    LBM_EXACT_BLOCK=1 ./lbmcl -D "$DEVICE" -d "$d" -n $VISCOSITY -u $VELOCITY -i "$ITERATIONS" -e $EVERY \
                        -w "$w" -s "$s" $FLAGS 2>> $LOG > /dev/null
                fi
            done
        done
    done
done

awk -F';' -v peak="$PEAK_GBS" -f aggregate.awk $LOG > $OUT
awk -F';' -v mode=reference -f aggregate.awk $LOG > $OUT_REF
echo "wrote $LOG, $OUT and $OUT_REF (roofline against $PEAK_GBS GB/s)"
if command -v column > /dev/null; then column -s';' -t $OUT | cut -c1-200; else cat $OUT; fi
