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
