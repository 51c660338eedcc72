// CheckMacros.h -- error convention of the host classes (apps/rtigo3/inc/CheckMacros.h:39-89): a failing
// core call throws std::runtime_error carrying "ERROR: file(line): call (code) text"; the _NO_THROW variant
// (used in destructors) only prints.
#pragma once
#include <iostream>
#include <sstream>
#include <stdexcept>

#include "rtc_core.h"

#define RTC_CHECK(call) \
  do { \
    const int rtc_result_ = (call); \
    if (rtc_result_ != 0) { \
      std::ostringstream message_; \
      message_ << "ERROR: " << __FILE__ << "(" << __LINE__ << "): " << #call << " (" << rtc_result_ << ") " << rtc_last_error(); \
      throw std::runtime_error(message_.str()); \
    } \
  } while (0)

#define RTC_CHECK_NO_THROW(call) \
  do { \
    const int rtc_result_ = (call); \
    if (rtc_result_ != 0) { \
      std::cerr << "ERROR: " << __FILE__ << "(" << __LINE__ << "): " << #call << " (" << rtc_result_ << ") " << rtc_last_error() << '\n'; \
    } \
  } while (0)

#define MY_ASSERT(expr) do { if (!(expr)) { std::cerr << "ASSERT: " << __FILE__ << "(" << __LINE__ << "): " #expr << '\n'; } } while (0)
