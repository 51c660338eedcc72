// Devices.h -- the four per-strategy device classes of rtigo3:
//   DeviceSingleGPU           apps/rtigo3/src/DeviceSingleGPU.cpp:104-270
//   DeviceMultiGPUZeroCopy    apps/rtigo3/src/DeviceMultiGPUZeroCopy.cpp:69-142   (shared pinned host buffer)
//   DeviceMultiGPUPeerAccess  apps/rtigo3/src/DeviceMultiGPUPeerAccess.cpp:74-181 (shared buffer on the first device, P2P stores)
//   DeviceMultiGPULocalCopy   apps/rtigo3/src/DeviceMultiGPULocalCopy.cpp:84-337  (per-device texel buffer + compositor)
// Each render() (re)allocates its buffers when the resolution changed and enqueues one launch.
#pragma once
#include "Device.h"

class DeviceSingleGPU : public Device
{
public:
  using Device::Device;
  ~DeviceSingleGPU() override;
  void activateContext() override {}
  void synchronizeStream() override { RTC_CHECK(rtc_synchronize(m_context)); }
  void render(const unsigned int iterationIndex, void** buffer) override { renderIterations(iterationIndex, 1, buffer); }
  void renderIterations(const unsigned int iterationFirst, const unsigned int count, void** buffer) override;
  void updateDisplayTexture() override {}
  const void* getOutputBufferHost() override;
};

class DeviceMultiGPUZeroCopy : public Device
{
public:
  using Device::Device;
  ~DeviceMultiGPUZeroCopy() override;
  void setState(DeviceState const& state) override;
  void activateContext() override {}
  void synchronizeStream() override { RTC_CHECK(rtc_synchronize(m_context)); }
  void render(const unsigned int iterationIndex, void** buffer) override { renderIterations(iterationIndex, 1, buffer); }
  void renderIterations(const unsigned int iterationFirst, const unsigned int count, void** buffer) override;
  void updateDisplayTexture() override {}
  const void* getOutputBufferHost() override;
private:
  void* m_pinned = nullptr;
};

class DeviceMultiGPUPeerAccess : public Device
{
public:
  using Device::Device;
  ~DeviceMultiGPUPeerAccess() override;
  void setState(DeviceState const& state) override;
  void activateContext() override {}
  void synchronizeStream() override { RTC_CHECK(rtc_synchronize(m_context)); }
  void render(const unsigned int iterationIndex, void** buffer) override { renderIterations(iterationIndex, 1, buffer); }
  void renderIterations(const unsigned int iterationFirst, const unsigned int count, void** buffer) override;
  void updateDisplayTexture() override {}
  const void* getOutputBufferHost() override;
};

class DeviceMultiGPULocalCopy : public Device
{
public:
  using Device::Device;
  ~DeviceMultiGPULocalCopy() override;
  void setState(DeviceState const& state) override;
  void compositor(Device* other) override;
  void activateContext() override {}
  void synchronizeStream() override { RTC_CHECK(rtc_synchronize(m_context)); }
  void render(const unsigned int iterationIndex, void** buffer) override { renderIterations(iterationIndex, 1, buffer); }
  void renderIterations(const unsigned int iterationFirst, const unsigned int count, void** buffer) override;
  void updateDisplayTexture() override {}
  const void* getOutputBufferHost() override;
  int getLaunchWidth() const { return m_launchWidth; }
  uint64_t getTexelBuffer() const { return m_systemData.texelBuffer; }
  uint64_t getOutputBuffer() const { return m_systemData.outputBuffer; }
};

// launch width of the tiled strategies: ceil(res.x / count) rounded up to whole tiles (DeviceMultiGPULocalCopy.cpp:91-93)
inline int tiledLaunchWidth(DeviceState const& state, int count)
{
  const int width = (state.resolution.x + count - 1) / count;
  const int mask = state.tileSize.x - 1;
  return (width + mask) & ~mask;
}
