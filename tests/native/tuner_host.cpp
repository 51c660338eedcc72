// tests/native/tuner_host.cpp -- TEST INFRASTRUCTURE: the schedule tuner's state machine (csrc/schedule_tuner.h) on the CPU.
//
// The header is included unchanged; the six CUDA calls it makes are replaced by a scripted clock: an "event" is an index into a
// table of time stamps, cudaEventRecord stamps the current simulated time, and the test advances the time by the duration it
// assigns to the schedule the tuner picked for the batch.  A call counter injects a CUDA error at a chosen call.
// tests/test_cpu_schedule_tuner.py drives it.
#include "rtc_internal.h"

#include <vector>

namespace mock {
double now = 0.0;
std::vector<double> stamps;
long calls = 0, failAt = -1;       // failAt: the 1-based CUDA call that returns an error
int destroyed = 0, preloads = 0, lastErrorCleared = 0;
inline bool fail() { ++calls; return calls == failAt; }
inline cudaError_t eventCreate(cudaEvent_t* e) { if (fail()) return cudaErrorMemoryAllocation; stamps.push_back(-1.0); *e = reinterpret_cast<cudaEvent_t>(stamps.size()); return cudaSuccess; }
inline cudaError_t eventRecord(cudaEvent_t e, cudaStream_t) { if (fail()) return cudaErrorInvalidResourceHandle; stamps[reinterpret_cast<size_t>(e) - 1] = now; return cudaSuccess; }
inline cudaError_t eventSynchronize(cudaEvent_t) { return fail() ? cudaErrorLaunchFailure : cudaSuccess; }
inline cudaError_t eventElapsedTime(float* ms, cudaEvent_t a, cudaEvent_t b)
{
  if (fail()) return cudaErrorNotReady;
  const double ta = stamps[reinterpret_cast<size_t>(a) - 1], tb = stamps[reinterpret_cast<size_t>(b) - 1];
  if (ta < 0.0 || tb < 0.0) return cudaErrorInvalidResourceHandle;
  *ms = (float)(tb - ta);
  return cudaSuccess;
}
inline cudaError_t eventDestroy(cudaEvent_t) { ++destroyed; return cudaSuccess; }
inline cudaError_t getLastError() { ++lastErrorCleared; return cudaSuccess; }
void preload() { ++preloads; }
}

#define cudaEventCreate      mock::eventCreate
#define cudaEventRecord      mock::eventRecord
#define cudaEventSynchronize mock::eventSynchronize
#define cudaEventElapsedTime mock::eventElapsedTime
#define cudaEventDestroy     mock::eventDestroy
#define cudaGetLastError     mock::getLastError
#include "schedule_tuner.h"

extern "C" {

// Plays `n` batches: paths[i] path samples, eligible[i] != 0, and a batch under schedule s takes duration[s] (+ drift * i) ms.
// Writes the schedule each batch ran with and the slot it filled; returns the final state.  out: {state, schedule, restarts,
// preloads, events destroyed at release, ms[0..3] * 1000 rounded}.
int tt_play(int n, const uint64_t* paths, const int* eligible, const double duration[3], double drift, long failAt, int forced,
            int* schedules, int* slots, long out[9])
{
  mock::now = 0.0; mock::stamps.clear(); mock::calls = 0; mock::failAt = failAt; mock::destroyed = 0; mock::preloads = 0;
  rtc_context ctx;
  if (forced >= 0) { ctx.traceSchedule = forced; ctx.tuner.state = ScheduleTuner::DONE; }
  for (int i = 0; i < n; ++i)
  {
    const int slot = rtc_tuner::begin(&ctx, paths[i], eligible[i] != 0, mock::preload);
    schedules[i] = ctx.traceSchedule; slots[i] = slot;
    mock::now += 0.25;                                   // the host gets round to recording the end a little later
    mock::now += duration[ctx.traceSchedule] + drift * i;
    rtc_tuner::end(&ctx, slot);
    mock::now += 0.125;
  }
  rtc_tuner::finish(&ctx);                               // what rtc_trace_schedule_get does
  out[0] = ctx.tuner.state; out[1] = ctx.traceSchedule; out[2] = ctx.tuner.restarts; out[3] = mock::preloads;
  for (int k = 0; k < 4; ++k) out[5 + k] = (long)(ctx.tuner.ms[k] * 1000.0f + 0.5f);
  rtc_tuner::release(&ctx);
  out[4] = mock::destroyed;
  return ctx.tuner.state;
}

} // extern "C"
