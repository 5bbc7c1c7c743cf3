"""CPU tier: the C-ABI library loads, exports every symbol include/fhe_b200.h declares, and fails loudly (no CPU
fallback) when no GPU is present.  No compute calls here."""
import ctypes as C
import os

import numpy as np
import pytest


def test_library_exports_every_declared_symbol(pkg):
    assert os.path.exists(pkg.LIB_PATH), "run __graft_entry__.build() first"
    L = pkg.lib()
    syms = pkg.header_symbols()
    assert len(syms) >= 50
    missing = [s for s in syms if not hasattr(L, s)]
    assert not missing, missing


def test_no_cpu_fallback(pkg):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    st = pkg.lib().fhe_ctx_create(0, C.byref(h))
    assert st in (pkg.FHE_ECUDA, pkg.FHE_EUNSUPPORTED) and not h.value
    with pytest.raises(pkg.FheError):
        pkg.Context(0)
    # every entry point refuses a null context instead of computing anything
    assert pkg.lib().fhe_ntt_fwd_u64(None, 97, 3, 1, None) == pkg.FHE_EINVAL
    assert pkg.lib().fhe_fhew_bootstrap_batch(None, None, None, 0, 1, None, None) == pkg.FHE_EINVAL


def test_product_package_does_not_import_oracle(pkg):
    import glob
    import re
    root = os.path.dirname(pkg.__file__)
    for path in glob.glob(os.path.join(root, "**", "*"), recursive=True):
        if path.endswith((".py", ".cu", ".cuh", ".hpp", ".h")):
            src = open(path).read()
            assert not re.search(r"^\s*(from|import)\s+oracle|#include\s+\"[^\"]*oracle/|liborc", src, re.M), path


def test_two_adic_primes_host_helper(pkg, orc):
    for bits, log_n, cnt in ((28, 10, 3), (55, 17, 4), (45, 5, 10), (61, 3, 2)):
        assert pkg.two_adic_primes(bits, log_n, cnt) == orc.two_adic_primes(bits, log_n, cnt)


def test_fhew_host_mirror(pkg, orc):
    from learn_fhe_b200 import fhew
    P = orc.fhew_testing_param()
    param = fhew.single_key_testing_param(P.big_q)
    for f in ("log_n", "big_q", "p", "rlwe_log_b", "rlwe_d", "rgsw_log_b", "rgsw_d", "n_s", "q_ks", "ks_log_b", "ks_d", "w"):
        assert getattr(param, f) == getattr(P, f), f
    for table in ([1, 1, 1, 0], [0, 0, 0, 1], [0, 1, 1, 1]):
        assert (fhew.gate_poly(param, table) == orc.fhew_gate_poly(P, table)).all()
    K = orc.FhewKey(P, 3)
    cts = K.encrypt(np.array([0, 1, 1], dtype=np.int32), 4)
    ct2n = K.prologue(cts)
    n_ext, n_auto = fhew.schedule_counts(param, ct2n)
    for i in range(3):
        st = orc.fhew_schedule(P, ct2n[i][:P.n_s])
        assert n_ext[i] == (st[:, 0] == 0).sum() and n_auto[i] == (st[:, 0] == 1).sum()
    assert (K.decrypt(fhew.Fhew.not_(param, cts)) == np.array([1, 0, 0])).all()
