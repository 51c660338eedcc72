// main.cpp -- headless rtigo3: the reference's main (apps/rtigo3/src/main.cpp:47-202) without the GLFW window.
//   rtigo3_b200 -s system_rtigo3_cornell_box.txt -d scene_rtigo3_cornell_box.txt [-m 1]
// mode 1 = benchmark(): render samplesSqrt^2 iterations, print "<spp> / <seconds> = <fps> fps", write the tonemapped PNG;
// mode 0 = the same accumulation without the timing print, writes the tonemapped PNG and the linear .hdr.
#include <iostream>

#include "Application.h"

int main(int argc, char* argv[])
{
  Options options;
  if (!options.parseCommandLine(argc, argv)) return 1;
  Application app(options);
  if (!app.isValid()) { std::cerr << "ERROR: Application failed to initialize successfully." << std::endl; return 2; }
  if (options.getMode() == 1)
  {
    app.benchmark();
  }
  else
  {
    // samplesSqrt^2 iterations, or this rank's share of them in a process group (render() stops at the local budget)
    const unsigned int spp = app.getRaytracer()->getSamplesPerPixelLocal();
    while (app.render(1) < spp && app.isValid()) {}
    app.screenshot(true);
    app.screenshot(false);
  }
  rtc_stats stats;
  app.getRaytracer()->getStats(stats);
  std::cout << "paths " << stats.pathSamples << ", radiance rays " << stats.radianceRays << ", shadow rays " << stats.shadowRays
            << ", kernel launches " << stats.kernelLaunches << std::endl;
  return app.isValid() ? 0 : 3;
}
