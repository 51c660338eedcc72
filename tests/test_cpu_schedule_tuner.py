"""The schedule tuner's state machine (csrc/schedule_tuner.h, behind include/rtc_core.h rtc_trace_schedule_get) on the CPU: the
header is compiled unchanged by g++ with its six CUDA event calls replaced by a scripted clock (tests/native/tuner_host.cpp), and
batches are "run" by advancing that clock by a duration per schedule.  Checked: the sequence of schedules the batches run with
(warm-up, group, one triangle, two triangles, group, then the decision), the decision rule (a capped schedule must beat the
FASTER group batch by 3 %; the faster capped one wins), what small / ineligible batches and batches of another size do, that a
fixed schedule is left alone, and that a CUDA error at ANY call inside the tuner ends in the measured schedule, never in a
failed launch."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import helpers as H

GROUP, ONE, TWO = 0, 1, 2
WARMUP, TIMING, PENDING, DONE = 0, 1, 2, 3
BIG = 1 << 21


@pytest.fixture(scope="module")
def tuner():
    src = os.path.join(H.ROOT, "tests", "native", "tuner_host.cpp")
    out_dir = os.path.join(H.ROOT, "oracle", "_build")
    os.makedirs(out_dir, exist_ok=True)
    so = os.path.join(out_dir, "libtuner_host.so")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-I" + os.path.join(H.ROOT, "include"),
                           "-I" + os.path.join(H.ROOT, "tweeker_raytracer_b200", "csrc"), "-I/usr/local/cuda/include", "-o", so, src])
    lib = C.CDLL(so)
    lib.tt_play.argtypes = [C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_long, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]

    def play(paths, durations, eligible=None, drift=0.0, fail_at=-1, forced=-1):
        n = len(paths)
        p = np.ascontiguousarray(paths, dtype=np.uint64)
        e = np.ascontiguousarray(eligible if eligible is not None else [1] * n, dtype=np.int32)
        d = np.ascontiguousarray(durations, dtype=np.float64)
        sched, slots, out = np.zeros(n, dtype=np.int32), np.zeros(n, dtype=np.int32), np.zeros(9, dtype=np.int64)
        lib.tt_play(n, p.ctypes.data, e.ctypes.data, d.ctypes.data, drift, fail_at, forced, sched.ctypes.data, slots.ctypes.data, out.ctypes.data)
        return {"schedules": sched.tolist(), "slots": slots.tolist(), "state": int(out[0]), "schedule": int(out[1]), "restarts": int(out[2]),
                "preloads": int(out[3]), "destroyed": int(out[4]), "ms": [v / 1000.0 for v in out[5:9]]}
    return play


def test_sequence_and_decision(tuner):
    # one triangle 10 % faster: warm-up, group, one, two, group, then the winner
    r = tuner([BIG] * 8, [26.0, 23.4, 24.0])
    assert r["schedules"] == [GROUP, GROUP, ONE, TWO, GROUP, ONE, ONE, ONE] and r["slots"] == [-1, 0, 1, 2, 3, -1, -1, -1]
    assert r["state"] == DONE and r["schedule"] == ONE and r["preloads"] == 1 and r["destroyed"] == 8
    assert r["ms"] == pytest.approx([26.25, 23.65, 24.25, 26.25], abs=2e-3)
    # two triangles faster than one
    assert tuner([BIG] * 7, [26.0, 24.5, 23.0])["schedule"] == TWO
    # nobody beats group by 3 %: 2 % is not enough
    r = tuner([BIG] * 7, [26.0, 25.5, 25.6])
    assert r["schedule"] == GROUP and r["schedules"][5:] == [GROUP, GROUP]
    # slower capped schedules
    assert tuner([BIG] * 7, [26.0, 27.0, 30.0])["schedule"] == GROUP


def test_the_faster_group_batch_is_the_reference(tuner):
    # the clocks drift: every batch is 0.3 ms faster than the one before; group runs 26.0 (batch 1) and 25.1 (batch 4), one triangle
    # 24.7 -- 5 % better than the first group batch but only 1.6 % better than the second: not enough
    r = tuner([BIG] * 7, [26.3, 25.3, 25.6], drift=-0.3)
    assert r["ms"][0] > r["ms"][3] and r["schedule"] == GROUP
    # and the other way round (getting slower): the FIRST group batch is the faster one
    r = tuner([BIG] * 7, [26.0, 25.0, 25.3], drift=+0.3)
    assert r["ms"][0] < r["ms"][3] and r["schedule"] == GROUP


def test_small_and_ineligible_batches_do_not_tune(tuner):
    r = tuner([1 << 19] * 10, [1.0, 0.5, 0.5])
    assert r["state"] == WARMUP and r["schedules"] == [GROUP] * 10 and r["preloads"] == 0 and r["destroyed"] == 0
    r = tuner([BIG] * 10, [26.0, 20.0, 20.0], eligible=[0] * 10)
    assert r["state"] == WARMUP and r["schedules"] == [GROUP] * 10
    # ineligible batches in between are passed over with the measured schedule; the measurement resumes where it was
    r = tuner([BIG] * 10, [26.0, 23.0, 24.0], eligible=[1, 1, 0, 1, 0, 1, 1, 1, 1, 1])
    assert r["schedules"] == [GROUP, GROUP, GROUP, ONE, GROUP, TWO, GROUP, ONE, ONE, ONE]
    assert r["slots"] == [-1, 0, -1, 1, -1, 2, 3, -1, -1, -1]


def test_batches_of_another_size_restart_the_measurement(tuner):
    r = tuner([BIG, BIG, BIG, 2 * BIG, 2 * BIG, 2 * BIG, 2 * BIG, 2 * BIG, 2 * BIG], [26.0, 23.0, 24.0])
    assert r["slots"] == [-1, 0, 1, 0, 1, 2, 3, -1, -1] and r["restarts"] == 1 and r["schedule"] == ONE
    # sizes that never settle: the tuner gives up after eight restarts with the measured schedule
    sizes = [BIG + (i % 2) * 4096 for i in range(40)]
    r = tuner(sizes, [26.0, 20.0, 20.0])
    assert r["state"] == DONE and r["schedule"] == GROUP and r["restarts"] == 9


def test_a_fixed_schedule_is_left_alone(tuner):
    for forced in (GROUP, ONE, TWO):
        r = tuner([BIG] * 6, [26.0, 20.0, 20.0], forced=forced)
        assert r["schedules"] == [forced] * 6 and r["slots"] == [-1] * 6 and r["preloads"] == 0 and r["schedule"] == forced


def test_the_decision_waits_until_it_is_asked_for(tuner):
    # five batches: warm-up + the four timed ones; no sixth batch takes the decision, rtc_trace_schedule_get (finish) does
    r = tuner([BIG] * 5, [26.0, 23.0, 24.0])
    assert r["schedules"] == [GROUP, GROUP, ONE, TWO, GROUP] and r["state"] == DONE and r["schedule"] == ONE
    # four batches: still timing, the measured schedule is in force
    r = tuner([BIG] * 4, [26.0, 23.0, 24.0])
    assert r["state"] == TIMING and r["schedule"] == GROUP


def test_a_cuda_error_anywhere_retires_the_tuner(tuner):
    # the full run makes 8 creates + 8 records + 1 synchronise + 4 elapsed-time calls; fail each of them in turn
    clean = tuner([BIG] * 8, [26.0, 23.0, 24.0])
    assert clean["schedule"] == ONE
    for k in range(1, 22):
        r = tuner([BIG] * 8, [26.0, 23.0, 24.0], fail_at=k)
        assert r["state"] == DONE and r["schedule"] == GROUP, k
        first_bad = next((i for i, s in enumerate(r["schedules"]) if s != clean["schedules"][i]), None)
        # every batch after the failure runs the measured schedule
        assert all(s == GROUP for s in r["schedules"][(first_bad or 0):]) or first_bad is None, k
    assert tuner([BIG] * 8, [26.0, 23.0, 24.0], fail_at=22)["schedule"] == ONE       # there is no 22nd call
