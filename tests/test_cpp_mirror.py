"""The C++ host-side mirror (include/hyrise_b200.hpp) must compile and link against the C-ABI it wraps, and the product
must fail loudly — not fall back to a CPU path — when no B200 is present."""
import os
import subprocess

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_example_compiles_links_and_has_no_cpu_fallback():
    build = subprocess.run(["make", "-C", REPO, "example"], capture_output=True, text=True)
    assert build.returncode == 0, build.stdout + build.stderr
    binary = os.path.join(REPO, "build", "tpch_operators")
    assert os.path.exists(binary)
    run = subprocess.run([binary, "0.01"], capture_output=True, text=True, timeout=300)
    if run.returncode != 0:  # no usable GPU here: the only acceptable outcome is a loud hyb_status error
        assert "hyb_status" in run.stderr, run.stderr
    else:  # on a B200 the plan runs: Q1 on TPC-H data has the four (returnflag, linestatus) groups
        assert "AggregateHash  4 groups" in run.stdout, run.stdout


def test_mirror_predicate_casts_match_reference_expectations():
    """lossless_predicate_variant_cast / flip_predicate_condition / ScanPredicate::normalized of the C++ mirror against the
    expectations of the reference's lossless_predicate_cast_test.cpp:49-91 (host logic, no GPU)."""
    build = subprocess.run(["make", "-C", REPO, "build/predicate_cast_check"], capture_output=True, text=True)
    assert build.returncode == 0, build.stdout + build.stderr
    environment = dict(os.environ, HYB_BINARY_FIXTURE=os.path.join(REPO, "tests", "golden", "bin", "AllTypesNullValues", "Dictionary.bin"))
    run = subprocess.run([os.path.join(REPO, "build", "predicate_cast_check")], capture_output=True, text=True, timeout=60,
                         env=environment)   # + the BinaryTable wrapper on one of the reference's binary fixtures
    assert run.returncode == 0, run.stdout + run.stderr


import pytest  # noqa: E402


@pytest.mark.gpu
def test_example_runs_on_the_gpu():
    """The C++ mirror (TableScan / JoinHash / AggregateHash classes of include/hyrise_b200.hpp) executing the bench plan on
    the B200: the example checks its own results (match counts, pair count = lineitem rows, the four Q1 groups)."""
    build = subprocess.run(["make", "-C", REPO, "example"], capture_output=True, text=True)
    assert build.returncode == 0, build.stdout + build.stderr
    run = subprocess.run([os.path.join(REPO, "build", "tpch_operators"), "0.1"], capture_output=True, text=True, timeout=600)
    assert run.returncode == 0, run.stdout[-2000:] + run.stderr[-2000:]
    assert "AggregateHash  4 groups" in run.stdout, run.stdout
    assert "JoinHash" in run.stdout and "TableScan" in run.stdout, run.stdout
