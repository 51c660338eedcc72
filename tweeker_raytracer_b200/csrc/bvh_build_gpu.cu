// bvh_build_gpu.cu -- placeholder until the device LBVH builder lands (next milestone).
#include "rtc_internal.h"
int build_gas_gpu(rtc_context*, GasRecord&) { RTC_FAIL("GPU LBVH builder not available in this build"); }
