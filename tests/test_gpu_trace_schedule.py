"""GPU parity, part 11: the schedules of the triangle tests (csrc/trace.cuh Traversal::step, include/rtc_core.h
rtc_trace_schedule_get).  The capped ones (RTC_SCHEDULE_ONE_TRI, RTC_SCHEDULE_TWO_TRI) were written after the last GPU session of
round 2, so by default the library only uses one where its own measurement says it is faster; whatever it picks, frames, hits
and ray counts must not change:

  * forced (RTC_TRACE_SCHEDULE=onetri | twotri): rendered frames, ray statistics, closest hits and occlusion of random rays equal
    the oracle bit for bit -- the same assertions the default schedule passes in parts 1 and 2;
  * measured (the default): batches of >= 1 Mi paths run warm-up / group / one triangle / two triangles / group / decided; the
    tuner reports four positive batch times and a decision, and the frame, whose iterations ran under ALL schedules, equals the
    oracle's rows bit for bit;
  * RTC_TRACE_SCHEDULE=group and rtc_trace_schedule_set pin the choice without measuring.
"""
import numpy as np
import pytest

import helpers as H
from oracle import orc
from tweeker_raytracer_b200 import host

pytestmark = pytest.mark.gpu


def frames(tmp, name, iterations, batch, **overrides):
    app = host.App(H.write_system(tmp, name, **overrides), H.scene_path(name))
    try:
        ref = H.oracle_scene(app)
        done = 0
        while done < iterations:
            done = app.render(min(batch, iterations - done))
        got = app.frame()
        w, h = app.resolution
        st = orc.Stats()
        want = ref.render(H.oracle_sys(app), app.info.miss, w, h, iter_count=iterations, stats=st).reshape(h, w, 4)
        stats = app.stats()
        ctx = app.context(0)
        rays = H.random_rays(40000, seed=17)
        top = app.system_data(0).topObject
        hits, occluded = ctx.trace_closest_host(top, rays), ctx.trace_any_host(top, rays)
        return got, want, stats, st, hits, occluded, ref.trace_closest(rays), ref.trace_any(rays), ctx.trace_schedule()
    finally:
        app.close()


@pytest.mark.parametrize("name,iterations,batch,overrides", [
    ("rtigo3_cornell_box", 16, 16, dict(resolution="128 128")),
    ("rtigo3_cornell_box", 9, 4, dict(resolution="96 64", samplesSqrt=3)),
    ("rtigo3_geometry", 8, 8, dict(resolution="240 136", samplesSqrt=3)),
    ("rtigo3_geometry", 4, 2, dict(resolution="200 112", samplesSqrt=2, miss=2, envMap="procedural 256 128", envRotation=0.15)),
])
@pytest.mark.parametrize("forced", ["onetri", "twotri"])
def test_forced_capped_schedules_are_bit_exact(cuda_device, tmp_path, monkeypatch, forced, name, iterations, batch, overrides):
    monkeypatch.setenv("RTC_TRACE_SCHEDULE", forced)
    got, want, stats, st, hits, occluded, want_hits, want_occluded, schedule = frames(tmp_path, name, iterations, batch, **overrides)
    assert schedule["schedule"] == {"onetri": "one_tri", "twotri": "two_tri"}[forced] and schedule["decided"] and not schedule["measured"]
    assert got.tobytes() == want.tobytes()
    assert stats.radianceRays == st.radianceRays and stats.shadowRays == st.shadowRays and stats.stackOverflows == 0
    assert H.hits_equal(hits, want_hits)
    assert np.array_equal(occluded.astype(bool), want_occluded.astype(bool))


def test_fixed_group_schedule_and_the_setter(cuda_device, tmp_path, monkeypatch):
    monkeypatch.setenv("RTC_TRACE_SCHEDULE", "group")
    app = host.App(H.write_system(tmp_path, "rtigo3_cornell_box", resolution="64 64", samplesSqrt=2), H.scene_path("rtigo3_cornell_box"))
    try:
        ctx = app.context(0)
        s = ctx.trace_schedule()
        assert s["schedule"] == "group" and s["decided"] and not s["measured"]
        ctx.set_trace_schedule("two_tri")
        assert ctx.trace_schedule()["schedule"] == "two_tri"
        app.render(4)
        got = app.frame()
        ref = H.oracle_scene(app)
        assert got.tobytes() == ref.render(H.oracle_sys(app), app.info.miss, 64, 64, iter_count=4).reshape(64, 64, 4).tobytes()
        ctx.set_trace_schedule("auto")
        s = ctx.trace_schedule()
        assert s["schedule"] == "group" and not s["decided"]
    finally:
        app.close()


def test_schedule_tuner_measures_and_frames_stay_bit_exact(cuda_device, tmp_path, monkeypatch):
    monkeypatch.delenv("RTC_TRACE_SCHEDULE", raising=False)
    w, h = 1280, 832                      # 1 064 960 paths per one-iteration batch: just above the tuner's 1 Mi threshold
    app = host.App(H.write_system(tmp_path, "rtigo3_geometry", resolution="%d %d" % (w, h), samplesSqrt=3), H.scene_path("rtigo3_geometry"))
    try:
        ctx = app.context(0)
        assert not ctx.trace_schedule()["decided"]
        for i in range(7):                # warm-up, group, one triangle, two triangles, group, then two batches with the decision taken
            assert app.render(1) == i + 1
            app.synchronize()
        s = ctx.trace_schedule()
        assert s["decided"] and s["measured"] and s["paths_per_batch"] == w * h
        times = {"group": min(s["group_ms"]), "one_tri": s["one_tri_ms"], "two_tri": s["two_tri_ms"]}
        assert min(s["group_ms"]) > 0.0 and times["one_tri"] > 0.0 and times["two_tri"] > 0.0
        if s["schedule"] == "group":      # nobody beat the faster group batch by 3 %
            assert min(times["one_tri"], times["two_tri"]) >= 0.9699 * times["group"]
        else:
            assert times[s["schedule"]] == min(times["one_tri"], times["two_tri"]) and times[s["schedule"]] < 0.9701 * times["group"]
        got = app.frame()
        ref = H.oracle_scene(app)
        want = ref.render(H.oracle_sys(app), app.info.miss, w, h, iter_count=7, row_step=64).reshape(h, w, 4)
        assert got[::64].tobytes() == want[::64].tobytes()
        stats = app.stats()
        assert stats.pathSamples == w * h * 7 and stats.stackOverflows == 0
    finally:
        app.close()
