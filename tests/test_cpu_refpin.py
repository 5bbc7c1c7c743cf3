"""Known-answer replay of the reference's own outputs (oracle/pin/README.md).

test_reference_known_answers: consumes tests/golden/ref/ref_*.json, which only a machine with cargo can produce
(`oracle/pin/apply.sh /path/to/learn-fhe`); skipped - NOT passed - while those files are absent, and DESIGN.md keeps saying
"bitwise parity pinned to the restatement only" until they are committed.
test_replay_code_on_self_generated_files: the same replay code on files the ORACLE wrote in the same schema, so the consumer
cannot rot; it pins nothing."""
import importlib.util
import os

import pytest

import refpin


def _run_all(orc, d):
    counts = {}
    for name, fn in (("ref_util.json", refpin.check_util), ("ref_fhew.json", refpin.check_fhew), ("ref_tfhe.json", refpin.check_tfhe),
                     ("ref_ckks.json", refpin.check_ckks)):
        G = refpin.load(d, name)
        if G is not None:
            counts[name] = fn(orc, G)
    return counts


def test_reference_known_answers(orc):
    if not any(os.path.exists(os.path.join(refpin.REF_DIR, f)) for f in refpin.FILES):
        pytest.skip("no tests/golden/ref/ref_*.json: run oracle/pin/apply.sh on a machine with cargo to pin the oracle to the reference")
    counts = _run_all(orc, refpin.REF_DIR)
    assert counts and all(v > 0 for v in counts.values()), counts


def test_replay_code_on_self_generated_files(orc, tmp_path):
    spec = importlib.util.spec_from_file_location("selfcheck", os.path.join(os.path.dirname(refpin.HERE), "oracle", "pin", "selfcheck.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main(str(tmp_path))
    counts = _run_all(orc, str(tmp_path))
    assert set(counts) == set(refpin.FILES) and all(v > 0 for v in counts.values()), counts
