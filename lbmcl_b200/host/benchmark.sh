#!/bin/bash
# Parameter sweep of the host program, the B200 counterpart of the reference's benchmark.sh:
# same protocol (reference benchmark.sh:8-15, 60-107: -i 50 -e 0, REPS repetitions per configuration,
# the ';'-separated statistics line of stderr appended to benchmarks/stats.csv) and the same
# aggregation (benchmark.sh:109-176: per configuration drop the fastest and the slowest run by total
# time and average the rest) into benchmarks/benchmark.csv.
# The work-group size is not swept: on this path it is a hint that does not change the CUDA block.
#
#   ./benchmark.sh                      # fp32, dims 8..256
#   PRECISION=double DIMS="64 128 256 512" GPUS=1 ./benchmark.sh
set -e
cd "$(dirname "$0")"

BENCHMARK_DIR=./benchmarks
LOG=$BENCHMARK_DIR/stats.csv
OUT=$BENCHMARK_DIR/benchmark.csv
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

mkdir -p $BENCHMARK_DIR
rm -f $LOG $OUT
make -s lbmcl

FLAGS="$OPTIMIZE $EXTRA"
[ "$PRECISION" = double ] && FLAGS="$FLAGS -F"
[ "$GPUS" != 1 ] && FLAGS="$FLAGS -G $GPUS"

for d in $DIMS; do
    for s in $STRIDES; do
        [ "$s" = full ] && s=$((d * d * d))
        [ "$s" -gt $((d * d * d)) ] && continue
        for r in $(seq 1 "$REPS"); do
            ./lbmcl -D "$DEVICE" -d "$d" -n $VISCOSITY -u $VELOCITY -i "$ITERATIONS" -e $EVERY -s "$s" $FLAGS \
                2>> $LOG > /dev/null
        done
    done
done

awk -F';' -f aggregate.awk $LOG > $OUT
echo "wrote $LOG and $OUT"
if command -v column > /dev/null; then column -s';' -t $OUT | cut -c1-200; else cat $OUT; fi
