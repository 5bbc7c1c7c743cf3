#!/bin/bash
# One-GPU profiling pass (run under gpurun): each command first runs plain (must exit 0), then under ncu.
#   gpurun_out/prof_<what>_full.ncu-rep : ncu --set full captures of full-size launches (read with tools/ncu_summary.py)
#   gpurun_out/launches_bench.csv       : launch list of the bench command (gpu__time_duration only)
# usage: bash tools/profile_all.sh <tag> [what...]   what in {fhew, ntt, tfhe, ckks, launches}; default all
set -u
TAG=${1:-rXX}; shift || true
WHAT=${*:-fhew ntt tfhe ckks launches}
OUT=gpurun_out
mkdir -p $OUT
NCU="ncu --set full --clock-control none --import-source on -f"
run() { # name, kernel regex, launch count, command...
    local name=$1 re=$2 cnt=$3; shift 3
    "$@" > $OUT/plain_${name}.log 2>&1 || { echo "plain run of $name failed"; tail -5 $OUT/plain_${name}.log; return 1; }
    $NCU -k regex:$re -c $cnt -o $OUT/prof_${name}_${TAG} "$@" > $OUT/ncu_${name}.log 2>&1 || { echo "ncu run of $name failed"; tail -5 $OUT/ncu_${name}.log; return 1; }
    python tools/ncu_summary.py $OUT/prof_${name}_${TAG}.ncu-rep ${TRAFFIC:+--traffic $TRAFFIC --traffic-out $OUT/ncu_traffic.json} > $OUT/sum_${name}_${TAG}.csv
    # per-source-line attribution of the first captured launch (needs -lineinfo + --import-source)
    ncu -i $OUT/prof_${name}_${TAG}.ncu-rep --page source --csv --print-source cuda,sass --launch-count 1 > $OUT/src_${name}_${TAG}.csv 2>/dev/null
    gzip -f $OUT/src_${name}_${TAG}.csv
    [ "${KEEP:-0}" = 1 ] || rm -f $OUT/prof_${name}_${TAG}.ncu-rep   # gpurun copies back at most 64 MiB
    echo "$name: ok"
}
for w in $WHAT; do
    case $w in
        fhew) TRAFFIC=fhew_blind_rotate_fast_kernel=16384 KEEP=1 run fhew fhew_blind_rotate_fast 1 python tools/prof_cmd.py fhew --batch 16384 ;;
        ntt) TRAFFIC= run ntt ntt_fast 12 python tools/prof_cmd.py ntt ;;
        tfhe) TRAFFIC=tfhe_blind_rotate_fast_kernel=16384 KEEP=1 run tfhe tfhe_blind_rotate_fast 1 python tools/tfhe_bench.py tfhe --batch 16384 --modes 3 ;;
        tfhe_exact) TRAFFIC=tfhe_blind_rotate_kernel=16384 KEEP=1 run tfhe_exact 'tfhe_blind_rotate_kernel' 1 python tools/tfhe_bench.py tfhe --batch 16384 --modes 0 ;;
        fhew64) TRAFFIC= KEEP=1 run fhew64 'fhew_blind_rotate_kernel' 1 python tools/fhew_wide_bench.py --batch 296 ;;
        bf) TRAFFIC= KEEP=1 run bf bf_rate 8 python tools/tfhe_bench.py bf ;;
        ckks) TRAFFIC= run ckks 'rns_|ckks_' 8 python tools/tfhe_bench.py ckks --count 128 ;;
        traffic)  # DRAM bytes of whole multi-kernel operations (one metric pass): Ckks::mul on 512 pairs, NTT fwd 4096 x 2^16 u64
            M="--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none --csv"
            python tools/tfhe_bench.py ckks --count 512 > $OUT/plain_traffic_ckks.log 2>&1 &&
            ncu $M --log-file $OUT/traffic_ckks_${TAG}.csv python tools/tfhe_bench.py ckks --count 512 > $OUT/ncu_traffic_ckks.log 2>&1 &&
            python tools/ncu_traffic_sum.py $OUT/traffic_ckks_${TAG}.csv ckks_mul_whole_op 512 3 'ckks_|rns_|ntt_fast' && cp profiles/ncu_traffic.json $OUT/ncu_traffic_merged.json
            python tools/prof_cmd.py ntt16 > $OUT/plain_traffic_ntt.log 2>&1 &&
            ncu $M --log-file $OUT/traffic_ntt_${TAG}.csv python tools/prof_cmd.py ntt16 > $OUT/ncu_traffic_ntt.log 2>&1 &&
            python tools/ncu_traffic_sum.py $OUT/traffic_ntt_${TAG}.csv 'ntt_fwd_u64_2^16' 4096 3 'ntt_fast' && cp profiles/ncu_traffic.json $OUT/ncu_traffic_merged.json
            echo "traffic: ok" ;;
        launches)
            python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/plain_bench.log 2>&1 || { echo "plain bench failed"; continue; }
            ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file $OUT/launches_bench_${TAG}.csv \
                python bench.py --steps 2 --warmup 3 --no-cpu > $OUT/ncu_bench.log 2>&1 && echo "launches: ok" ;;
    esac
done
