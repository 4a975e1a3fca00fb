"""CPU-only checks of the drop-in boundary: the C-ABI library loads, exports every
symbol include/sdpsr.h declares, and fails loudly (no CPU fallback) without a GPU."""
import os
import re

import numpy as np
import pytest

from sdpsr_b200 import binding as B
import sdpsr_b200 as S

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "sdpsr.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdpsr_[a-z0-9_]+)\s*\(", src)))


def test_header_and_binding_agree():
    assert header_symbols() == B.declared_symbols()


def test_library_loads_and_exports_every_symbol():
    lib = B.load_library()
    for name in header_symbols():
        assert hasattr(lib, name), name
    assert lib.sdpsr_version() == 100


def test_no_cpu_fallback():
    """Without a CUDA device the product must refuse, not fall back."""
    if B.device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(B.SdpsrError) as e:
        B.Context(8)
    assert e.value.code == B.E_NO_DEVICE
    with pytest.raises(B.SdpsrError):
        S.Partition(np.eye(3))
    with pytest.raises(B.SdpsrError):
        S.admissible_subspace(*S.problems.petersen())


def test_product_does_not_import_oracle():
    """The oracle is test infrastructure: nothing under the package may import it."""
    pkg = os.path.join(ROOT, "sdpsymmetryreduction.jl_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f


def test_host_scalar_logic_matches_oracle():
    """Otsu threshold / eigen clustering / union-find are host logic on both sides."""
    import oracle as O
    from sdpsr_b200 import api
    rng = np.random.default_rng(0)
    for _ in range(20):
        ne = int(rng.integers(2, 30))
        X = np.abs(rng.standard_normal((ne, ne))) * 10.0 ** rng.integers(-14, 2, size=(ne, ne))
        X = np.maximum(X, X.T)
        assert api.otsu_threshold(X, 1e-8) == O.otsu_threshold(X, 1e-8)
    v = np.sort(rng.random(50).round(2))
    assert np.array_equal(api.eigen_clusters(v, 1e-3), O.eigen_clusters(v, 1e-3))


def test_problem_generators():
    p = S.problems.lovasz_er(3)
    assert p.n == 13 and p.A.shape == (2, 169) and p.A[0].sum() == 13 * 4 - 4   # 4 absolute points
    p = S.problems.hamming(2, 8)
    assert p.n == 64 and p.A[0].sum() == 64 * 14
    p = S.problems.kneser(5, 2)
    assert p.n == 10 and p.A[0].sum() == 30
    p = S.problems.synthetic_product_scheme(3, 2, 8)
    assert p.n == 64 and p.A.shape == (8, 4096) and p.expected_dim == 27
    ps = S.problems.hamming(2, 8, sparse=True)
    assert np.array_equal(ps.A.toarray(), S.problems.hamming(2, 8, sparse=False).A)


def test_bench_reference_arm_runs_on_cpu():
    """bench.py --impl reference is CPU-only (oracle port) and must keep working without a GPU."""
    import subprocess
    import sys
    import json
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload",
                        "theta-H(3,4)-N64", "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-1000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "s" and line["higher_is_better"] is False
    assert line["cpu_baseline"]["kind"] == "port" and line["e2e"]["h2d_bytes_per_step"] == 0
