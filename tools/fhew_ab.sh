#!/bin/bash
# A/B of library builds on the FHEW headline: bash tools/fhew_ab.sh [variant names under learn-fhe_b200/variants]
for v in default "$@"; do
    if [ $v = default ]; then unset FHE_B200_LIB; else export FHE_B200_LIB=/root/repo/learn-fhe_b200/variants/libfhe_b200_$v.so; fi
    python bench.py --no-ntt --no-tfhe --no-ckks --no-cpu --steps 5 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$v', 'value %.0f e2e %.0f ms/launch %.3f frac %.3f'%(d['value'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac']))"
done
