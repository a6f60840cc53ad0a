// TEST INFRASTRUCTURE ONLY (oracle). Stand-in for btllib/nthash.hpp (class NtHash; call site
// /root/reference/src/utils.cpp:113-118).  The arithmetic is NOT restated here: it is the
// reference's own vendored statement of ntHash, lib/nthash.hpp (N-aware NTMC64 seed at
// :412-437, rolling NTMC64 at :304-314), driven with the skip rule of
// lib/ntHashIterator.hpp:45-72.  ntEdit queries btllib-built filters with those in-tree
// functions (ntedit.cpp:1442,1470), which pins btllib::NtHash == lib/nthash.hpp for ACGT.
#ifndef GP_SHIM_BTLLIB_NTHASH_HPP
#define GP_SHIM_BTLLIB_NTHASH_HPP

#include "lib/nthash.hpp" // found via -I/root/reference/subprojects/ntedit

#include <cstddef>
#include <cstdint>
#include <limits>
#include <memory>
#include <string>

namespace btllib {

class NtHash
{
public:
  NtHash(const char* seq, size_t seq_len, unsigned hash_num, unsigned k, size_t pos = 0)
    : seq(seq)
    , seq_len(seq_len)
    , hash_num(hash_num)
    , k(k)
    , pos(pos)
    , hashes_array(new uint64_t[hash_num])
  {
  }

  NtHash(const std::string& seq, unsigned hash_num, unsigned k, size_t pos = 0)
    : NtHash(seq.data(), seq.size(), hash_num, k, pos)
  {
  }

  bool roll()
  {
    if (!initialized) {
      return init();
    }
    if (pos >= seq_len - k) {
      return false;
    }
    if (seedTab[(unsigned char)seq[pos + k]] == seedN) {
      pos += k;
      return init();
    }
    NTMC64((unsigned char)seq[pos],
           (unsigned char)seq[pos + k],
           k,
           hash_num,
           fh,
           rh,
           hashes_array.get());
    ++pos;
    return true;
  }

  const uint64_t* hashes() const { return hashes_array.get(); }
  size_t get_pos() const { return pos; }
  unsigned get_hash_num() const { return hash_num; }
  unsigned get_k() const { return k; }
  uint64_t get_forward_hash() const { return fh; }
  uint64_t get_reverse_hash() const { return rh; }

private:
  bool init()
  {
    if (k > seq_len) {
      pos = std::numeric_limits<size_t>::max();
      return false;
    }
    unsigned loc_n = 0;
    while (pos <= seq_len - k &&
           !NTMC64(seq + pos, k, hash_num, fh, rh, loc_n, hashes_array.get())) {
      pos += loc_n + 1;
    }
    if (pos > seq_len - k) {
      pos = std::numeric_limits<size_t>::max();
      return false;
    }
    initialized = true;
    return true;
  }

  const char* seq;
  const size_t seq_len;
  const unsigned hash_num;
  const unsigned k;
  size_t pos;
  std::unique_ptr<uint64_t[]> hashes_array;
  uint64_t fh = 0, rh = 0;
  bool initialized = false;
};

} // namespace btllib

#endif
