"""include/b9_spec_chain.hpp: batched evaluation under an unmodified sequential MCMC step gives,
for the same seed, the sequential chain bit for bit.  Builds and runs tests/cpp/spec_chain_test.cpp
(g++, CPU only); the samplers and target in it are made up for the test."""
import json
import subprocess
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def cases(tmp_path_factory):
    exe = tmp_path_factory.mktemp("spec") / "spec_chain_test"
    subprocess.run(["g++", "-std=c++17", "-O2", "-Wall", "-Wextra", "-Werror", f"-I{ROOT / 'include'}",
                    str(ROOT / "tests" / "cpp" / "spec_chain_test.cpp"), "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    lines = [json.loads(l) for l in r.stdout.splitlines()]
    assert r.returncode == 0, lines
    return lines


def test_every_speculative_chain_is_the_sequential_chain_bit_for_bit(cases):
    runs = [c for c in cases if "mode" in c]
    assert len(runs) == 19 and {c["mode"] for c in runs} == {0, 1, 2}
    for c in runs:
        assert c["identical"] is True, c
        assert c["used"] == c["sequential_evals"], c      # replay consumed exactly the evaluations the sequential run made


def test_depth_one_is_the_sequential_driver(cases):
    for c in cases:
        if c.get("depth") == 1 and c["mode"] in (0, 1):
            assert c["launches"] == c["steps"] and c["undone"] == 0 and c["evaluated"] == c["sequential_evals"]


def test_steps_per_launch_follow_the_acceptance_rate(cases):
    """A round of depth K completes (1 - (1-a)^K) / a steps on average when each step asks for one value."""
    for c in cases:
        if c.get("chains") == 1 and c["mode"] in (0, 1) and c["depth"] > 1:
            a, k = c["acceptance"], c["depth"]
            want = (1 - (1 - a) ** k) / a
            assert c["chain_steps_per_launch"] == pytest.approx(want, rel=0.06), c


def test_independent_chains_share_each_launch(cases):
    many = [c for c in cases if c.get("chains") == 64]
    assert len(many) == 3
    for c in many:
        assert c["largest_batch"] > 64 * c["depth"] // 2 and c["launches"] < c["steps"], c


def test_a_step_that_needs_no_evaluation_still_completes(cases):
    assert {"idle_steps": 10} in cases
