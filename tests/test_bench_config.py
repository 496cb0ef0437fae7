"""bench.py's host-side bookkeeping (no GPU): how functions are sharded over ranks, what a line's `config` says, and the
algorithmic bytes per function behind `roofline.achieved`."""
import argparse
import importlib.util
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench_module", os.path.join(ROOT, "bench.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def _args(bench, workload, scaling=None, n=None):
    w = dict(bench.WORKLOADS[workload])
    if n:
        w["n"] = n
    if scaling and w["scaling"] != "replicas":
        w["scaling"] = scaling
    return argparse.Namespace(workload=workload, w=w)


@pytest.mark.parametrize("world", [1, 2, 3, 4, 8])
def test_strong_scaling_shards_partition_the_functions(bench, world):
    a = _args(bench, "nstar", "strong", n=1_000_003)
    sizes = bench.shard_sizes(a, world)
    assert len(sizes) == world and sum(sizes) == 1_000_003 and max(sizes) - min(sizes) <= 1
    cfg = bench.config_dict(a, world)
    assert cfg["n_total"] == 1_000_003 and cfg["scaling_mode"] == "strong" and cfg["name"] == "nstar"


@pytest.mark.parametrize("world", [1, 2, 8])
def test_weak_scaling_and_replicas_keep_the_shard_size(bench, world):
    a = _args(bench, "nstar")
    assert bench.shard_sizes(a, world) == [1_000_000] * world
    assert bench.config_dict(a, world)["n_total"] == 1_000_000 * world
    c5 = _args(bench, "c5", "strong")                     # replicas: one independent chain per GPU whatever is asked
    assert c5.w["scaling"] == "replicas" and bench.shard_sizes(c5, world) == [200_000] * world
    assert bench.config_dict(c5, world)["n_total"] == 200_000


def test_workloads_are_the_baseline_configurations(bench):
    w = bench.WORKLOADS
    assert (w["nstar"]["K"], w["nstar"]["P"], w["nstar"]["M"], w["nstar"]["n"], w["nstar"]["T"]) == (3, 20, 3, 1_000_000, 200)
    assert (w["c2"]["n"], w["c3"]["P"], w["c3"]["M"], w["c3"]["n"]) == (100_000, 64, 4, 1_000_000)
    assert (w["c4"]["D"], w["c4"]["kind"], w["c4"]["n"]) == (2, "ragged", 1_000_000)
    assert (w["c5"]["K"], w["c5"]["P"], w["c5"]["n"]) == (4, 400, 200_000)
    assert "n_funct=1000000 per GPU" in bench.unit_name(_args(bench, "nstar"), 1)


def test_algorithmic_bytes_per_function(bench):
    """8(P+2K+M) for the Z pass (+ the proposal cache in the two-kernel form), 8(P+1+K+2M) chi, 8(P+1+K+M) SSR,
    8(P+K+M) statistics: the figures of DESIGN.md section 3 at K=3, P=20, M=3."""
    ab = bench.algorithmic_bytes(_args(bench, "nstar"), 1, None)
    assert ab == {"z_kernel": 232, "chi_kernel": 240, "ssr_kernel": 216, "stats_kernels": 208}
    K, P, M = 3, 20, 3
    assert ab["z_kernel"] + 8 * (K + 2) == 272 and 8 * (2 * K + 2) == 64            # accept half, proposal kernel
    assert 8 * (P + 1 + K + M + (M + 1)) == 248 and 8 * (K + M + (M + 1) + M) == 104  # moments pass, chi draw
