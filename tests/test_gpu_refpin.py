"""The CUDA library against the reference's own known answers (tests/golden/ref/, see oracle/pin/README.md) when they exist, and
against oracle-written files of the same schema otherwise (keeps the replay path exercised; pins nothing)."""
import importlib.util
import os

import pytest

import refpin

pytestmark = pytest.mark.gpu


def test_cuda_reproduces_reference_known_answers(pkg, ctx, orc):
    if not any(os.path.exists(os.path.join(refpin.REF_DIR, f)) for f in refpin.FILES):
        pytest.skip("no tests/golden/ref/ref_*.json: run oracle/pin/apply.sh on a machine with cargo")
    assert refpin.check_gpu(pkg, ctx, orc, refpin.REF_DIR) > 0


def test_cuda_replay_on_self_generated_files(pkg, ctx, orc, tmp_path):
    spec = importlib.util.spec_from_file_location("selfcheck", os.path.join(os.path.dirname(refpin.HERE), "oracle", "pin", "selfcheck.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.main(str(tmp_path))
    assert refpin.check_gpu(pkg, ctx, orc, str(tmp_path)) >= 10
