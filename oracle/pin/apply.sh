#!/bin/bash
# TEST INFRASTRUCTURE: installs the known-answer dumpers into a checkout of han0110/learn-fhe and runs them.
#   usage: oracle/pin/apply.sh /path/to/learn-fhe [/path/to/learn-fhe_b200/tests/golden/ref]
# Needs cargo (absent from the image this repository was developed in, which is why parity is "pinned to the restatement"
# until somebody runs this once).  It adds four files and three `#[cfg(test)] mod pin_dump;` lines to the checkout - test
# code only, nothing the library builds - then `cargo test` writes ref_util.json, ref_fhew.json, ref_tfhe.json, ref_ckks.json.
# Afterwards:  python -m pytest tests/test_cpu_refpin.py -q        (oracle vs the reference's own outputs, CPU)
#              python -m pytest tests/test_gpu_refpin.py -q -m gpu (CUDA library vs the reference's own outputs)
set -euo pipefail
REF=${1:?path to a checkout of han0110/learn-fhe}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=${2:-$HERE/../../tests/golden/ref}
mkdir -p "$OUT" "$REF/util/tests" "$REF/scheme/fhew/src/bootstrapping" "$REF/scheme/tfhe/src/bootstrapping" "$REF/scheme/ckks/src/ckks"
cp "$HERE/util_pin_dump.rs" "$REF/util/tests/pin_dump.rs"
cp "$HERE/fhew_pin_dump.rs" "$REF/scheme/fhew/src/bootstrapping/pin_dump.rs"
cp "$HERE/tfhe_pin_dump.rs" "$REF/scheme/tfhe/src/bootstrapping/pin_dump.rs"
cp "$HERE/ckks_pin_dump.rs" "$REF/scheme/ckks/src/ckks/pin_dump.rs"
for f in scheme/fhew/src/bootstrapping.rs scheme/tfhe/src/bootstrapping.rs scheme/ckks/src/ckks.rs; do
    grep -q 'mod pin_dump;' "$REF/$f" || printf '\n#[cfg(test)]\nmod pin_dump;\n' >> "$REF/$f"
done
cd "$REF"
export FHE_PIN_OUT=$(cd "$OUT" && pwd)
cargo test --release -p util --test pin_dump -- --nocapture
cargo test --release -p fhew pin_dump -- --nocapture
cargo test --release -p tfhe pin_dump -- --nocapture
cargo test --release -p ckks pin_dump -- --nocapture
ls -l "$FHE_PIN_OUT"
