"""Host-side logic of bench.py that the driver depends on (no GPU): defaults, launch-group choice, identical workload strings in
both arms, and the memory cap of the CPU arms (an uncapped thread sweep took the GPU box down twice in round 2)."""
import importlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def _parse(argv):
    bench = importlib.import_module("bench")
    old = sys.argv
    sys.argv = ["bench.py"] + argv
    try:
        return bench, bench.parse()
    finally:
        sys.argv = old


def test_defaults_are_the_north_star_headline():
    bench, a = _parse([])
    assert (a.gpus, a.impl, a.config) == (1, "ours", "c4")
    assert (a.nodes, a.edges, a.batch, a.fanout, a.dim) == (100_000_000, 1_000_000_000, 8192, "15,10", 256)
    assert a.warmup >= 3 and a.steps >= 1 and a.label != "custom"


def test_driver_run_keeps_every_lane_busy():
    bench, a = _parse(["--steps", "20", "--warmup", "5"])
    counts = [int(x) for x in a.fanout.split(",")]
    G = bench.auto_group(a, counts)
    assert G == 3                                      # ceil(20 / 8 lanes), under the 5M-row budget of a launch group
    assert G * a.batch * counts[0] * counts[1] <= 5_000_000
    assert -(-a.steps // G) >= a.lanes - 1             # 6 full groups + a tail: 7 of 8 lanes in flight
    bench, c2 = _parse(["--config", "c2", "--steps", "64", "--warmup", "16"])
    assert bench.auto_group(c2, [25, 10]) == 16


def test_both_arms_print_the_same_config():
    bench, ours = _parse(["--steps", "20", "--warmup", "5"])
    _, ref = _parse(["--impl", "reference", "--steps", "20", "--warmup", "5"])
    counts = [15, 10]
    for n in (1, 2, 8):
        assert bench.workload_config(ours, counts, n) == bench.workload_config(ref, counts, n)
    assert (ours.steps, ours.warmup) == (ref.steps, ref.warmup)


def test_cpu_arm_memory_cap():
    bench, a = _parse([])
    assert 0 < bench.host_mem_budget() <= 48 << 30
    cap = bench.cpu_threads_cap(a, [15, 10])
    rows = a.batch * 15 * 10
    assert 1 <= cap <= bench.host_cores()
    assert cap * rows * a.dim * 4 * 2.6 <= 48 << 30    # the threads' minibatch buffers fit the budget
