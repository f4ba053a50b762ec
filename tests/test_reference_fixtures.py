"""Provenance of tests/golden/ref_*.npz: when the reference checkout is present (the build container), re-running
tools/make_reference_goldens.py — which EXECUTES the reference's own bflow_jax_maf.py / statutils.py — must reproduce the
committed fixtures bit for bit.  Skipped where /root/reference does not exist (the GPU box)."""
import importlib.util
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/src/naz/flows/bflow_jax_maf.py"


@pytest.mark.skipif(not os.path.exists(REF), reason="reference checkout not present")
def test_committed_reference_fixtures_are_reproducible(tmp_path):
    spec = importlib.util.spec_from_file_location("make_reference_goldens", os.path.join(ROOT, "tools", "make_reference_goldens.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gen.main(str(tmp_path))
    gen.stats_fixture(str(tmp_path))
    gen.truncnorm_fixture(str(tmp_path))
    gen.grad_fixture(str(tmp_path))     # torch-backed run of the same reference code + autograd
    gen.bounded_sampler_fixture(str(tmp_path))
    names = [f for f in os.listdir(tmp_path) if f.endswith(".npz")]
    assert len(names) >= 9
    for f in names:
        new = np.load(os.path.join(tmp_path, f))
        old = np.load(os.path.join(ROOT, "tests", "golden", f))
        assert sorted(new.files) == sorted(old.files), f
        for k in new.files:
            if f.startswith("ref_twin_grad_"):   # autograd through threaded fp64 matmuls: reproducible to round-off, not bitwise
                assert np.allclose(new[k], old[k], rtol=1e-10, atol=1e-12), (f, k)
            else:
                assert np.array_equal(new[k], old[k]), (f, k)
