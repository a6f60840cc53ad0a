// TEST INFRASTRUCTURE ONLY (oracle). Stand-in for btllib/util.hpp: split, endswith,
// calc_phred_avg (call sites: /root/reference/src/seqindex.cpp:32-33,45,51;
// /root/reference/src/mappings.cpp:21-24).
#ifndef GP_SHIM_BTLLIB_UTIL_HPP
#define GP_SHIM_BTLLIB_UTIL_HPP

#include "status.hpp"

#include <string>
#include <vector>

namespace btllib {

inline std::vector<std::string>
split(const std::string& s, const std::string& delim)
{
  std::vector<std::string> tokens;
  size_t pos1 = 0, pos2 = 0;
  while ((pos2 = s.find(delim, pos1)) != std::string::npos) {
    tokens.push_back(s.substr(pos1, pos2 - pos1));
    pos1 = pos2 + delim.size();
  }
  tokens.push_back(s.substr(pos1));
  return tokens;
}

inline bool
endswith(const std::string& s, const std::string& suffix)
{
  return s.size() >= suffix.size() &&
         s.compare(s.size() - suffix.size(), suffix.size(), suffix) == 0;
}

// Mean of the quality characters minus the Sanger offset (33).  len == 0 means
// "to the end of the string" in btllib; the reference always passes len explicitly.
inline double
calc_phred_avg(const std::string& qual, const size_t start_pos = 0, size_t len = 0)
{
  if (len == 0) {
    len = qual.size() - start_pos;
  }
  check_error(len == 0, "calc_phred_avg: empty quality string.");
  check_error(start_pos + len > qual.size(), "calc_phred_avg: range exceeds string.");
  size_t phred_sum = 0;
  for (size_t i = start_pos; i < start_pos + len; ++i) {
    phred_sum += size_t((unsigned char)qual[i]);
  }
  return (double(phred_sum) / double(len)) - 33.0;
}

} // namespace btllib

#endif
