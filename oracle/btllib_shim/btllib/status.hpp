// TEST INFRASTRUCTURE ONLY (oracle). Minimal stand-in for btllib/status.hpp so that the
// reference's own sources (/root/reference/src/*.cpp, subprojects/ntedit/ntedit.cpp)
// compile unmodified.  btllib itself is NOT in /root/reference (meson.build:37 finds it as
// an external library), so this is a restatement of its published behaviour: parity for
// anything defined only here is "unpinned" (see DESIGN.md).
#ifndef GP_SHIM_BTLLIB_STATUS_HPP
#define GP_SHIM_BTLLIB_STATUS_HPP

#include <cerrno>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

namespace btllib {

inline bool&
shim_quiet()
{
  static bool q = (std::getenv("GP_ORACLE_QUIET") != nullptr);
  return q;
}

inline void
log_info(const std::string& msg)
{
  if (!shim_quiet()) {
    std::cerr << "[btllib-shim] [INFO] " << msg << std::endl;
  }
}

inline void
log_warning(const std::string& msg)
{
  std::cerr << "[btllib-shim] [WARNING] " << msg << std::endl;
}

inline void
log_error(const std::string& msg)
{
  std::cerr << "[btllib-shim] [ERROR] " << msg << std::endl;
}

inline void
check_error(bool condition, const std::string& msg)
{
  if (condition) {
    log_error(msg);
    std::exit(EXIT_FAILURE);
  }
}

inline void
check_warning(bool condition, const std::string& msg)
{
  if (condition) {
    log_warning(msg);
  }
}

inline std::string
get_strerror()
{
  return std::strerror(errno);
}

inline void
check_stream(const std::ios& stream, const std::string& name)
{
  check_error(!stream.good(), "'" + name + "' stream error: " + get_strerror());
}

} // namespace btllib

#endif
