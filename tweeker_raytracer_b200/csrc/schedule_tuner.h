// schedule_tuner.h -- the state machine behind rtc_context::tuner (rtc_internal.h ScheduleTuner, include/rtc_core.h
// rtc_trace_schedule_get): which schedule of the triangle tests the launches of a context use, decided by timing one batch
// with each.  Host code; included by kernels_shade.cu, and by tests/native/tuner_host.cpp with the six CUDA calls below replaced
// by a scripted clock, so that the sequence of schedules, the decision rule and the error paths are tested without a GPU.
// Nothing in here may fail a launch: on any CUDA error the tuner retires with the schedule rounds 1 and 2 measured.
#pragma once

#include "rtc_internal.h"

namespace rtc_tuner {

constexpr uint64_t kMinPaths = 1ull << 20;     // smaller batches are launch- and tail-bound: nothing to learn from timing them
constexpr float    kMargin = 0.97f;            // a capped schedule must beat the faster group batch by 3 %
// the schedule each timed batch runs: group first and last, so that a drift of the clocks cannot favour a capped schedule
constexpr int      kSchedule[ScheduleTuner::kSlots] = { RTC_SCHEDULE_GROUP, RTC_SCHEDULE_ONE_TRI, RTC_SCHEDULE_TWO_TRI, RTC_SCHEDULE_GROUP };

inline void give_up(rtc_context* ctx)
{
  cudaGetLastError();
  ctx->tuner.state = ScheduleTuner::DONE;
  ctx->traceSchedule = RTC_SCHEDULE_GROUP;
}

// Decides once the timed batches have finished (waits for the last one).
inline void finish(rtc_context* ctx)
{
  ScheduleTuner& t = ctx->tuner;
  if (t.state != ScheduleTuner::PENDING) return;
  bool ok = cudaEventSynchronize(t.ev[2 * ScheduleTuner::kSlots - 1]) == cudaSuccess;
  for (int slot = 0; ok && slot < ScheduleTuner::kSlots; ++slot) ok = cudaEventElapsedTime(&t.ms[slot], t.ev[2 * slot], t.ev[2 * slot + 1]) == cudaSuccess;
  if (!ok) { give_up(ctx); return; }
  const float group = t.ms[0] < t.ms[3] ? t.ms[0] : t.ms[3];
  int best = RTC_SCHEDULE_GROUP; float bestMs = kMargin * group;
  for (int slot = 1; slot <= 2; ++slot) if (t.ms[slot] > 0.0f && t.ms[slot] < bestMs) { bestMs = t.ms[slot]; best = kSchedule[slot]; }
  ctx->traceSchedule = best;
  t.state = ScheduleTuner::DONE;
}

// Called before the launches of a batch.  Returns the timed slot the batch fills (0 .. kSlots - 1) or -1, and sets
// ctx->traceSchedule for the batch.  preload: loads the capped kernels (called once, during the warm-up batch).
inline int begin(rtc_context* ctx, uint64_t paths, bool eligible, void (*preload)())
{
  ScheduleTuner& t = ctx->tuner;
  if (t.state == ScheduleTuner::DONE) return -1;
  if (t.state == ScheduleTuner::PENDING) { finish(ctx); return -1; }
  ctx->traceSchedule = RTC_SCHEDULE_GROUP;
  if (!eligible || paths < kMinPaths) return -1;
  if (t.state == ScheduleTuner::WARMUP)
  {
    // the first batch pays for allocations, lazily loaded kernels and the clock ramp: not timed
    for (int k = 0; k < 2 * ScheduleTuner::kSlots; ++k) if (!t.ev[k] && cudaEventCreate(&t.ev[k]) != cudaSuccess) { give_up(ctx); return -1; }
    if (preload) preload();
    t.state = ScheduleTuner::TIMING;
    t.slot = 0;
    return -1;
  }
  if (t.slot > 0 && paths != t.paths)
  {
    if (++t.restarts > 8) { t.state = ScheduleTuner::DONE; return -1; }
    t.slot = 0;
  }
  if (t.slot == 0) t.paths = paths;
  ctx->traceSchedule = kSchedule[t.slot];
  if (cudaEventRecord(t.ev[2 * t.slot], ctx->stream) != cudaSuccess) { give_up(ctx); return -1; }
  return t.slot;
}

// Called behind the launches of a batch begin() gave a slot.
inline void end(rtc_context* ctx, int slot)
{
  if (slot < 0) return;
  ScheduleTuner& t = ctx->tuner;
  ctx->traceSchedule = RTC_SCHEDULE_GROUP;
  if (t.state != ScheduleTuner::TIMING) return;
  if (cudaEventRecord(t.ev[2 * slot + 1], ctx->stream) != cudaSuccess) { give_up(ctx); return; }
  t.slot = slot + 1;
  if (t.slot == ScheduleTuner::kSlots) t.state = ScheduleTuner::PENDING;
}

inline void release(rtc_context* ctx)
{
  for (int k = 0; k < 2 * ScheduleTuner::kSlots; ++k) if (ctx->tuner.ev[k]) { cudaEventDestroy(ctx->tuner.ev[k]); ctx->tuner.ev[k] = nullptr; }
}

} // namespace rtc_tuner
