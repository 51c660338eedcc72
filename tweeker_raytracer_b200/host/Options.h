// Options.h -- the five command line options of rtigo3 (apps/rtigo3/src/Options.cpp:44-113):
//   -w/--width <int>  -h/--height <int>  -m/--mode <0|1>  -s/--system <file>  -d/--desc <file>
#pragma once
#include <string>

class Options
{
public:
  bool parseCommandLine(int argc, char* argv[]);
  int getWidth() const { return m_width; }
  int getHeight() const { return m_height; }
  int getMode() const { return m_mode; }
  std::string getSystem() const { return m_filenameSystem; }
  std::string getScene() const { return m_filenameScene; }
  void set(int w, int h, int mode, std::string const& system, std::string const& scene)
  { m_width = w; m_height = h; m_mode = mode; m_filenameSystem = system; m_filenameScene = scene; }

private:
  void printUsage(std::string const& argv0);
  int m_width = 512, m_height = 512, m_mode = 0;
  std::string m_filenameSystem, m_filenameScene;
};
