"""The BASELINE-size fixtures (tests/golden/c{2,3,4,5}_oracle.npz) are what the oracle produces: re-derive their
first iterations here on the CPU (tools/make_parity_fixtures.py wrote them; the GPU tests and bench.py compare the
CUDA path with them)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, golden

sys.path.insert(0, os.path.join(ROOT, "tools"))
import make_parity_fixtures as mk  # noqa: E402


@pytest.mark.parametrize("name,iters", [("c3", 5), ("c2", 12), ("c4", 8)])
def test_fixture_is_reproduced_by_the_oracle(orc, tables, name, iters):
    c = dict(mk.CASES[name])
    c["nmax"] = iters            # the schedule of iteration k does not depend on nmax
    c["fixed"] = True
    out = mk.run_case(orc, tables, c, threads=min(8, os.cpu_count() or 1))
    ref = golden(name + "_oracle")
    m = len(out["k"])
    assert np.array_equal(out["k"], ref["k"][:m]) and np.array_equal(out["t"], ref["t"][:m])
    for key in ("hy2", "hyb", "bb", "boundary", "lambda_min"):
        np.testing.assert_allclose(out[key], ref[key][:m], rtol=1e-12)
    scale = np.abs(ref["hy2"][:m]) + 2 * np.abs(ref["hyb"][:m]) + np.abs(ref["bb"][:m])
    assert np.all(np.abs(out["r_comp"] - ref["r_comp"][:m]) <= 1e-12 * scale)
    # columns 1..iters of H (column iters+1 belongs to a step the short run never takes)
    np.testing.assert_allclose(out["H1"][:, :iters], ref["H1"][: iters + 1, :iters], rtol=0, atol=1e-12 * np.abs(ref["H1"]).max())


def test_fixtures_cover_the_benchmarked_configuration():
    for name, d in (("c3", 256), ("c5", 1024)):
        ref = golden(name + "_oracle")
        assert f"'d': {d}" in str(ref["meta"]) and "'n': 10000" in str(ref["meta"])
        assert ref["k"][0] == 2 and ref["k"][-1] == 16 and int(ref["fallbacks"]) == 0
