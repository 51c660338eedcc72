#include "Devices.h"

#include <cstring>

static size_t framePixels(SystemData const& s) { return (size_t)s.resolution.x * (size_t)s.resolution.y; }

// ------------------------------------------------------------------ single GPU
DeviceSingleGPU::~DeviceSingleGPU()
{
  if (m_context && m_systemData.outputBuffer) { RTC_CHECK_NO_THROW(rtc_free(m_context, m_systemData.outputBuffer)); m_systemData.outputBuffer = 0; }
}

void DeviceSingleGPU::renderIterations(const unsigned int iterationFirst, const unsigned int count, void** buffer)
{
  (void)buffer;
  if (m_isDirtyOutputBuffer)
  {
    synchronizeStream();
    hostBuffer(framePixels(m_systemData));
    if (m_systemData.outputBuffer) RTC_CHECK(rtc_free(m_context, m_systemData.outputBuffer));
    RTC_CHECK(rtc_malloc(m_context, sizeof(float4) * framePixels(m_systemData), &m_systemData.outputBuffer));
    RTC_CHECK(rtc_memset(m_context, m_systemData.outputBuffer, 0, sizeof(float4) * framePixels(m_systemData)));
    m_isDirtyOutputBuffer = false;
    m_isDirtySystemData = true;
  }
  launch((unsigned int)m_systemData.resolution.x, RTC_RAYGEN_FULL_FRAME, iterationFirst, count);
}

const void* DeviceSingleGPU::getOutputBufferHost()
{
  float4* host = hostBuffer(framePixels(m_systemData));
  if (m_systemData.outputBuffer)
    RTC_CHECK(rtc_download(m_context, host, m_systemData.outputBuffer, sizeof(float4) * framePixels(m_systemData)));
  synchronizeStream();
  return host;
}

// ------------------------------------------------------------------ zero copy: every GPU accumulates into one pinned host buffer
DeviceMultiGPUZeroCopy::~DeviceMultiGPUZeroCopy()
{
  if (m_context && m_ownsSharedBuffer && m_pinned) { RTC_CHECK_NO_THROW(rtc_synchronize(m_context)); RTC_CHECK_NO_THROW(rtc_host_free(m_context, m_pinned)); }
  m_systemData.outputBuffer = 0;
}

void DeviceMultiGPUZeroCopy::setState(DeviceState const& state)
{
  if (m_systemData.resolution != state.resolution || m_systemData.tileSize != state.tileSize) m_launchWidth = tiledLaunchWidth(state, m_count);
  Device::setState(state);
}

void DeviceMultiGPUZeroCopy::renderIterations(const unsigned int iterationFirst, const unsigned int count, void** buffer)
{
  if (m_isDirtyOutputBuffer)
  {
    synchronizeStream();
    MY_ASSERT(buffer != nullptr);
    if (*buffer == nullptr)
    {
      if (m_ownsSharedBuffer && m_pinned) RTC_CHECK(rtc_host_free(m_context, m_pinned));
      RTC_CHECK(rtc_host_alloc(m_context, sizeof(float4) * framePixels(m_systemData), &m_pinned));
      std::memset(m_pinned, 0, sizeof(float4) * framePixels(m_systemData));
      *buffer = m_pinned;
      m_ownsSharedBuffer = true;
    }
    m_systemData.outputBuffer = (uint64_t)(uintptr_t)*buffer;   // UVA: the mapped host pointer is valid on every device
    m_pinned = *buffer;
    m_isDirtyOutputBuffer = false;
    m_isDirtySystemData = true;
  }
  launch((unsigned int)m_launchWidth, RTC_RAYGEN_FULL_FRAME, iterationFirst, count);
}

const void* DeviceMultiGPUZeroCopy::getOutputBufferHost()
{
  synchronizeStream();
  return m_pinned;   // the accumulation buffer already lives in host memory
}

// ------------------------------------------------------------------ peer access: one device owns the frame, the others store through NVLink
DeviceMultiGPUPeerAccess::~DeviceMultiGPUPeerAccess()
{
  if (m_context && m_ownsSharedBuffer && m_systemData.outputBuffer) RTC_CHECK_NO_THROW(rtc_free(m_context, m_systemData.outputBuffer));
  m_systemData.outputBuffer = 0;
}

void DeviceMultiGPUPeerAccess::setState(DeviceState const& state)
{
  if (m_systemData.resolution != state.resolution || m_systemData.tileSize != state.tileSize) m_launchWidth = tiledLaunchWidth(state, m_count);
  Device::setState(state);
}

void DeviceMultiGPUPeerAccess::renderIterations(const unsigned int iterationFirst, const unsigned int count, void** buffer)
{
  if (m_isDirtyOutputBuffer)
  {
    synchronizeStream();
    MY_ASSERT(buffer != nullptr);
    if (*buffer == nullptr)
    {
      hostBuffer(framePixels(m_systemData));
      if (m_ownsSharedBuffer && m_systemData.outputBuffer) RTC_CHECK(rtc_free(m_context, m_systemData.outputBuffer));
      RTC_CHECK(rtc_malloc(m_context, sizeof(float4) * framePixels(m_systemData), &m_systemData.outputBuffer));
      RTC_CHECK(rtc_memset(m_context, m_systemData.outputBuffer, 0, sizeof(float4) * framePixels(m_systemData)));
      synchronizeStream();
      *buffer = (void*)(uintptr_t)m_systemData.outputBuffer;
      m_ownsSharedBuffer = true;
    }
    else
    {
      m_systemData.outputBuffer = (uint64_t)(uintptr_t)*buffer;
    }
    m_isDirtyOutputBuffer = false;
    m_isDirtySystemData = true;
  }
  launch((unsigned int)m_launchWidth, RTC_RAYGEN_FULL_FRAME, iterationFirst, count);
}

const void* DeviceMultiGPUPeerAccess::getOutputBufferHost()
{
  // only called on the owner, after Raytracer::synchronize() of all devices
  float4* host = hostBuffer(framePixels(m_systemData));
  RTC_CHECK(rtc_download(m_context, host, m_systemData.outputBuffer, sizeof(float4) * framePixels(m_systemData)));
  synchronizeStream();
  return host;
}

// ------------------------------------------------------------------ local copy: per-device texel slab, composited on the first device
DeviceMultiGPULocalCopy::~DeviceMultiGPULocalCopy()
{
  if (!m_context) return;
  if (m_ownsSharedBuffer)
  {
    if (m_systemData.outputBuffer) RTC_CHECK_NO_THROW(rtc_free(m_context, m_systemData.outputBuffer));
    if (m_systemData.tileBuffer) RTC_CHECK_NO_THROW(rtc_free(m_context, m_systemData.tileBuffer));
  }
  if (m_systemData.texelBuffer) RTC_CHECK_NO_THROW(rtc_free(m_context, m_systemData.texelBuffer));
  m_systemData.outputBuffer = 0; m_systemData.tileBuffer = 0; m_systemData.texelBuffer = 0;
}

void DeviceMultiGPULocalCopy::setState(DeviceState const& state)
{
  if (m_systemData.resolution != state.resolution || m_systemData.tileSize != state.tileSize) m_launchWidth = tiledLaunchWidth(state, m_count);
  Device::setState(state);
}

void DeviceMultiGPULocalCopy::renderIterations(const unsigned int iterationFirst, const unsigned int count, void** buffer)
{
  if (m_isDirtyOutputBuffer)
  {
    synchronizeStream();
    MY_ASSERT(buffer != nullptr);
    const size_t slab = sizeof(float4) * (size_t)m_launchWidth * (size_t)m_systemData.resolution.y;
    if (*buffer == nullptr)   // the device called first holds the full-resolution frame and the staging slab of the compositor
    {
      hostBuffer(framePixels(m_systemData));
      if (m_ownsSharedBuffer && m_systemData.outputBuffer) RTC_CHECK(rtc_free(m_context, m_systemData.outputBuffer));
      if (m_ownsSharedBuffer && m_systemData.tileBuffer) RTC_CHECK(rtc_free(m_context, m_systemData.tileBuffer));
      RTC_CHECK(rtc_malloc(m_context, sizeof(float4) * framePixels(m_systemData), &m_systemData.outputBuffer));
      RTC_CHECK(rtc_memset(m_context, m_systemData.outputBuffer, 0, sizeof(float4) * framePixels(m_systemData)));
      RTC_CHECK(rtc_malloc(m_context, slab, &m_systemData.tileBuffer));
      *buffer = (void*)(uintptr_t)m_systemData.outputBuffer;
      m_ownsSharedBuffer = true;
    }
    if (m_systemData.texelBuffer) RTC_CHECK(rtc_free(m_context, m_systemData.texelBuffer));
    RTC_CHECK(rtc_malloc(m_context, slab, &m_systemData.texelBuffer));
    RTC_CHECK(rtc_memset(m_context, m_systemData.texelBuffer, 0, slab));
    m_isDirtyOutputBuffer = false;
    m_isDirtySystemData = true;
  }
  launch((unsigned int)m_launchWidth, RTC_RAYGEN_LOCAL_COPY, iterationFirst, count);
}

// `this` is the destination; `other` may be this device itself (DeviceMultiGPULocalCopy.cpp:279-337).
void DeviceMultiGPULocalCopy::compositor(Device* other)
{
  DeviceMultiGPULocalCopy* src = static_cast<DeviceMultiGPULocalCopy*>(other);
  const size_t slab = sizeof(float4) * (size_t)m_launchWidth * (size_t)m_systemData.resolution.y;
  RTC_CHECK(rtc_memcpy_peer(m_context, m_systemData.tileBuffer, src->getContext(), src->getTexelBuffer(), slab));
  CompositorData args;
  args.outputBuffer = m_systemData.outputBuffer;
  args.tileBuffer = m_systemData.tileBuffer;
  args.resolution = m_systemData.resolution;
  args.tileSize = m_systemData.tileSize;
  args.tileShift = m_systemData.tileShift;
  args.launchWidth = m_launchWidth;
  args.deviceCount = m_systemData.deviceCount;
  args.deviceIndex = src->m_index;
  RTC_CHECK(rtc_composite(m_context, &args));
  synchronizeStream();   // the staging slab is reused by the next call
}

const void* DeviceMultiGPULocalCopy::getOutputBufferHost()
{
  float4* host = hostBuffer(framePixels(m_systemData));
  RTC_CHECK(rtc_download(m_context, host, m_systemData.outputBuffer, sizeof(float4) * framePixels(m_systemData)));
  synchronizeStream();
  return host;
}
