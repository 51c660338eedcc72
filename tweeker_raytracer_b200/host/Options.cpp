#include "Options.h"

#include <cstdlib>
#include <iostream>

bool Options::parseCommandLine(int argc, char* argv[])
{
  for (int i = 1; i < argc; ++i)
  {
    const std::string arg(argv[i]);
    const bool hasValue = i + 1 < argc;
    if (arg == "?" || arg == "help" || arg == "--help") { printUsage(argv[0]); return false; }
    else if (arg == "-w" || arg == "--width")  { if (!hasValue) { std::cerr << "Option '" << arg << "' is missing argument.\n"; printUsage(argv[0]); return false; } m_width = std::atoi(argv[++i]); }
    else if (arg == "-h" || arg == "--height") { if (!hasValue) { std::cerr << "Option '" << arg << "' is missing argument.\n"; printUsage(argv[0]); return false; } m_height = std::atoi(argv[++i]); }
    else if (arg == "-m" || arg == "--mode")   { if (!hasValue) { std::cerr << "Option '" << arg << "' is missing argument.\n"; printUsage(argv[0]); return false; } m_mode = std::atoi(argv[++i]); }
    else if (arg == "-s" || arg == "--system") { if (!hasValue) { std::cerr << "Option '" << arg << "' is missing argument.\n"; printUsage(argv[0]); return false; } m_filenameSystem = argv[++i]; }
    else if (arg == "-d" || arg == "--desc")   { if (!hasValue) { std::cerr << "Option '" << arg << "' is missing argument.\n"; printUsage(argv[0]); return false; } m_filenameScene = argv[++i]; }
    else { std::cerr << "Unknown option '" << arg << "'\n"; printUsage(argv[0]); return false; }
  }
  if (m_filenameSystem.empty()) { std::cerr << "ERROR: Options::parseCommandLine() System description filename is empty.\n"; printUsage(argv[0]); return false; }
  if (m_filenameScene.empty())  { std::cerr << "ERROR: Options::parseCommandLine() Scene description filename is empty.\n"; printUsage(argv[0]); return false; }
  return true;
}

void Options::printUsage(std::string const& argv0)
{
  std::cerr << "\nUsage: " << argv0 << " [options]\n"
    "App Options:\n"
    "   ? | help | --help       Print this usage message and exit.\n"
    "  -w | --width <int>       Width of the client window  (512) [unused: headless]\n"
    "  -h | --height <int>      Height of the client window (512) [unused: headless]\n"
    "  -m | --mode <int>        0 = interactive (renders samplesSqrt^2 iterations, no window), 1 = benchmark (default 0)\n"
    "  -s | --system <filename> Filename for system options (empty).\n"
    "  -d | --desc <filename>   Filename for scene description (empty).\n";
}
