// TEST INFRASTRUCTURE ONLY (oracle). Stand-in for btllib/bloom_filter.hpp (KmerBloomFilter;
// call sites /root/reference/src/goldpolish_targeted_bfs.cpp:75-76,139,
// /root/reference/src/utils.cpp:118, ntedit.cpp:2012-2022,1470).
//
// UNPINNED (btllib is absent from /root/reference): bit n = h % (bytes*8) lives in byte n/8
// under mask 1<<(n%8); the file is a TOML-style text header, "[HeaderEnd]", 50 placeholder
// newlines (the second carrying "  <binary data>"), then the raw payload.  Everything that
// depends on these choices sits in this header and in goldpolish_b200/csrc/bf_format.hpp.
#ifndef GP_SHIM_BTLLIB_BLOOM_FILTER_HPP
#define GP_SHIM_BTLLIB_BLOOM_FILTER_HPP

#include "status.hpp"

#include <climits>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

namespace btllib {

static const char* const KMER_BLOOM_FILTER_SIGNATURE = "[BTLKmerBloomFilter_v6]";
static const char* const SHIM_HASH_FN = "ntHash_v2";
static const unsigned PLACEHOLDER_NEWLINES = 50;

class KmerBloomFilter
{
public:
  KmerBloomFilter(size_t bytes, unsigned hash_num, unsigned k)
    : bytes(((bytes + 7) / 8) * 8)
    , array_bits(this->bytes * CHAR_BIT)
    , hash_num(hash_num)
    , k(k)
    , array(this->bytes, 0)
  {
    check_error(hash_num == 0, "KmerBloomFilter: hash_num must be > 0.");
  }

  explicit KmerBloomFilter(const std::string& path)
  {
    std::ifstream ifs(path, std::ios::in | std::ios::binary);
    check_error(!ifs.good(), "KmerBloomFilter: cannot open " + path);
    std::string line;
    bool got_sig = false, got_end = false;
    size_t f_bytes = 0;
    unsigned f_hash_num = 0, f_k = 0;
    while (bool(std::getline(ifs, line))) {
      if (line == "[HeaderEnd]") {
        got_end = true;
        break;
      }
      if (!got_sig) {
        check_error(line.rfind("[BTLKmerBloomFilter_v", 0) != 0,
                    "KmerBloomFilter: bad signature in " + path);
        got_sig = true;
        continue;
      }
      const auto eq = line.find('=');
      if (eq == std::string::npos) {
        continue;
      }
      auto trim = [](std::string s) {
        const auto b = s.find_first_not_of(" \t\"");
        const auto e = s.find_last_not_of(" \t\"\r");
        return b == std::string::npos ? std::string() : s.substr(b, e - b + 1);
      };
      const auto key = trim(line.substr(0, eq));
      const auto val = trim(line.substr(eq + 1));
      if (key == "bytes") {
        f_bytes = std::stoull(val);
      } else if (key == "hash_num") {
        f_hash_num = unsigned(std::stoul(val));
      } else if (key == "k") {
        f_k = unsigned(std::stoul(val));
      }
    }
    check_error(!got_end, "KmerBloomFilter: no [HeaderEnd] in " + path);
    for (unsigned i = 0; i < PLACEHOLDER_NEWLINES; i++) {
      std::getline(ifs, line);
    }
    bytes = f_bytes;
    array_bits = bytes * CHAR_BIT;
    hash_num = f_hash_num;
    k = f_k;
    array.assign(bytes, 0);
    ifs.read(reinterpret_cast<char*>(array.data()), std::streamsize(bytes));
    check_error(size_t(ifs.gcount()) != bytes, "KmerBloomFilter: truncated payload in " + path);
  }

  void insert(const uint64_t* hashes)
  {
    for (unsigned i = 0; i < hash_num; ++i) {
      const uint64_t n = hashes[i] % array_bits;
      array[n / CHAR_BIT] |= uint8_t(1u << (n % CHAR_BIT));
    }
  }

  bool contains(const uint64_t* hashes) const
  {
    for (unsigned i = 0; i < hash_num; ++i) {
      const uint64_t n = hashes[i] % array_bits;
      if ((array[n / CHAR_BIT] & uint8_t(1u << (n % CHAR_BIT))) == 0) {
        return false;
      }
    }
    return true;
  }

  void save(const std::string& path)
  {
    std::ofstream ofs(path, std::ios::out | std::ios::binary);
    check_error(!ofs.good(), "KmerBloomFilter: cannot write " + path);
    ofs << KMER_BLOOM_FILTER_SIGNATURE << '\n'
        << "bytes = " << bytes << '\n'
        << "hash_fn = \"" << SHIM_HASH_FN << "\"\n"
        << "hash_num = " << hash_num << '\n'
        << "k = " << k << '\n'
        << "[HeaderEnd]\n";
    for (unsigned i = 0; i < PLACEHOLDER_NEWLINES; i++) {
      if (i == 1) {
        ofs << "  <binary data>";
      }
      ofs << '\n';
    }
    ofs.write(reinterpret_cast<const char*>(array.data()), std::streamsize(bytes));
  }

  size_t get_bytes() const { return bytes; }
  unsigned get_hash_num() const { return hash_num; }
  unsigned get_k() const { return k; }
  const uint8_t* data() const { return array.data(); }
  uint8_t* data() { return array.data(); }

private:
  size_t bytes = 0;
  size_t array_bits = 0;
  unsigned hash_num = 0;
  unsigned k = 0;
  std::vector<uint8_t> array;
};

} // namespace btllib

#endif
