#!/bin/bash
# A/B of library builds on the TFHE PBS leg: bash tools/tfhe_ab.sh [variant names under learn-fhe_b200/variants]
for v in default "$@"; do
    if [ $v = default ]; then unset FHE_B200_LIB; else export FHE_B200_LIB=/root/repo/learn-fhe_b200/variants/libfhe_b200_$v.so; fi
    echo -n "$v: "; python tools/tfhe_bench.py tfhe --batch ${TFHE_AB_BATCH:-16384} --modes ${TFHE_AB_MODES:-3} 2>&1 | tail -1 | cut -c1-200
done
