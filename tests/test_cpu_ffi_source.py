"""CPU tier: the Rust -sys crate source under ffi/ is generated from include/fhe_b200.h and must not drift from it."""
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_rust_sys_crate_matches_header(pkg):
    path = os.path.join(ROOT, "ffi", "fhe-b200-sys", "src", "lib.rs")
    committed = open(path).read()
    subprocess.check_call([sys.executable, os.path.join(ROOT, "tools", "gen_rust_sys.py")], stdout=subprocess.DEVNULL)
    assert open(path).read() == committed, "ffi/fhe-b200-sys/src/lib.rs is stale: run tools/gen_rust_sys.py"
    declared = set(re.findall(r"pub fn (fhe_\w+)\(", committed))
    assert declared == set(pkg.header_symbols())
