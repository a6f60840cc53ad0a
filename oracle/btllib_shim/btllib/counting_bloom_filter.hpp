// TEST INFRASTRUCTURE ONLY (oracle). Stand-in for btllib/counting_bloom_filter.hpp
// (KmerCountingBloomFilter8; call sites /root/reference/src/goldpolish_targeted_bfs.cpp:73-74,
// /root/reference/src/utils.cpp:115-117).
//
// UNPINNED (btllib is absent from /root/reference): counters = bytes / sizeof(uint8_t);
// insert_thresh_contains reads min over the hash_num counters; if min < threshold every
// counter still equal to min is raised to min+1 (a compare-exchange per hash, so a counter
// hit by two of the four hashes moves once) and min+1 is returned, otherwise min is
// returned.  Define GP_SHIM_CBF_RETURN_BEFORE to model the alternative "return the count
// before the insert" convention (shifts every threshold by one).
#ifndef GP_SHIM_BTLLIB_COUNTING_BLOOM_FILTER_HPP
#define GP_SHIM_BTLLIB_COUNTING_BLOOM_FILTER_HPP

#include "status.hpp"

#include <cstdint>
#include <limits>
#include <vector>

namespace btllib {

template<typename T>
class KmerCountingBloomFilter
{
public:
  KmerCountingBloomFilter(size_t bytes, unsigned hash_num, unsigned k)
    : bytes(((bytes + 7) / 8) * 8)
    , array_size(this->bytes / sizeof(T))
    , hash_num(hash_num)
    , k(k)
    , array(array_size, 0)
  {
    check_error(hash_num == 0, "KmerCountingBloomFilter: hash_num must be > 0.");
  }

  T contains(const uint64_t* hashes) const
  {
    T min = array[hashes[0] % array_size];
    for (unsigned i = 1; i < hash_num; ++i) {
      const T v = array[hashes[i] % array_size];
      if (v < min) {
        min = v;
      }
    }
    return min;
  }

  T insert_thresh_contains(const uint64_t* hashes, const T threshold)
  {
    const T min_val = contains(hashes);
    if (min_val < threshold) {
      for (unsigned i = 0; i < hash_num; ++i) {
        T& ctr = array[hashes[i] % array_size];
        if (ctr == min_val) { // compare_exchange_strong(min_val, min_val + 1)
          ctr = T(min_val + 1);
        }
      }
#ifdef GP_SHIM_CBF_RETURN_BEFORE
      return min_val;
#else
      return T(min_val + 1);
#endif
    }
    return min_val;
  }

  size_t get_bytes() const { return bytes; }
  unsigned get_hash_num() const { return hash_num; }
  unsigned get_k() const { return k; }
  const T* data() const { return array.data(); }
  size_t size() const { return array_size; }

private:
  size_t bytes;
  size_t array_size;
  unsigned hash_num;
  unsigned k;
  std::vector<T> array;
};

using KmerCountingBloomFilter8 = KmerCountingBloomFilter<uint8_t>;

} // namespace btllib

#endif
