// TEST INFRASTRUCTURE ONLY (oracle). Stand-in for btllib/data_stream.hpp: the reference only
// needs a DataSource that converts to FILE* for getline() (mappings.cpp:134-139,187-192).
// Plain files only (btllib would also pipe .gz/.bam through external tools).
#ifndef GP_SHIM_BTLLIB_DATA_STREAM_HPP
#define GP_SHIM_BTLLIB_DATA_STREAM_HPP

#include "status.hpp"

#include <cstdio>
#include <string>

namespace btllib {

class DataSource
{
public:
  explicit DataSource(const std::string& path)
    : file(std::fopen(path.c_str(), "r"))
  {
    check_error(file == nullptr, "DataSource: cannot open " + path + ": " + get_strerror());
  }
  DataSource(const DataSource&) = delete;
  DataSource& operator=(const DataSource&) = delete;
  ~DataSource()
  {
    if (file != nullptr) {
      std::fclose(file);
    }
  }
  operator FILE*() const { return file; }

private:
  FILE* file;
};

} // namespace btllib

#endif
