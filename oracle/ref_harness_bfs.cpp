// TEST INFRASTRUCTURE ONLY (oracle/_ref).  In-process access to the reference's OWN
// Bloom-filter build code.  The reference translation units are #included from where they lie
// under /root/reference (found through -I, never copied); only main() is renamed.
//   fill_bfs                         /root/reference/src/utils.cpp:96-123
//   serve_batch                      /root/reference/src/goldpolish_targeted_bfs.cpp:55-146
//   mappings_bases_to_kmer_threshold /root/reference/src/goldpolish_targeted_bfs.cpp:45-53
//   SeqIndex / AllMappings           /root/reference/src/seqindex.cpp, mappings.cpp
#define main gp_ref_targeted_bfs_main
#include "goldpolish_targeted_bfs.cpp" // NOLINT  (reference source, via -I$(REF)/src)
#undef main
#include "mappings.cpp" // NOLINT
#include "seqindex.cpp" // NOLINT
#include "utils.cpp"    // NOLINT

#include "btllib/nthash.hpp"

#include <chrono>
#include <cstring>

namespace {

struct RefBuild
{
  std::vector<unsigned> k_values;
  std::vector<std::unique_ptr<btllib::KmerCountingBloomFilter8>> cbfs;
  std::vector<std::unique_ptr<btllib::KmerBloomFilter>> bfs;
};

} // namespace

extern "C" {

int
ref_kmer_threshold(unsigned long mappings_bases)
{
  return mappings_bases_to_kmer_threshold(mappings_bases);
}

void*
ref_build_open(const unsigned* k_values, int nk, size_t cbf_bytes, size_t bf_bytes, unsigned hash_num)
{
  auto* b = new RefBuild;
  for (int i = 0; i < nk; i++) {
    b->k_values.push_back(k_values[i]);
    b->cbfs.emplace_back(new btllib::KmerCountingBloomFilter8(cbf_bytes, hash_num, k_values[i]));
    b->bfs.emplace_back(new btllib::KmerBloomFilter(bf_bytes, hash_num, k_values[i]));
  }
  return b;
}

// One call == one iteration of the loop at goldpolish_targeted_bfs.cpp:130-133.
void
ref_build_add_read(void* h, const char* seq, size_t len, unsigned kmer_threshold)
{
  auto* b = static_cast<RefBuild*>(h);
  fill_bfs(seq, len, 4, b->k_values, kmer_threshold, b->cbfs, b->bfs);
}

void
ref_build_get_bf(void* h, int i, unsigned char* out)
{
  auto* b = static_cast<RefBuild*>(h);
  std::memcpy(out, b->bfs[i]->data(), b->bfs[i]->get_bytes());
}

void
ref_build_get_cbf(void* h, int i, unsigned char* out)
{
  auto* b = static_cast<RefBuild*>(h);
  std::memcpy(out, b->cbfs[i]->data(), b->cbfs[i]->size());
}

void
ref_build_close(void* h)
{
  delete static_cast<RefBuild*>(h);
}

// All valid k-mer hashes of a sequence exactly as the loop at utils.cpp:113-114 sees them.
// Returns the number of k-mers; fills pos[i] and hashes[4*i..4*i+3] up to cap entries.
size_t
ref_nthash_all(const char* seq, size_t len, unsigned k, size_t cap, uint64_t* pos, uint64_t* hashes)
{
  btllib::NtHash nthash(seq, len, 4, k);
  size_t n = 0;
  while (nthash.roll()) {
    if (n < cap) {
      pos[n] = nthash.get_pos();
      std::memcpy(hashes + 4 * n, nthash.hashes(), 4 * sizeof(uint64_t));
    }
    n++;
  }
  return n;
}

// CPU baseline / end-to-end oracle for the filter build: loads both indexes and the
// mappings once (as main() does, goldpolish_targeted_bfs.cpp:274-281), then runs the
// reference's serve_batch for every batch with one OpenMP thread per batch (the reference's
// own schedule, :177-192).  ids_files[i] is a regular file holding the batch's target ids
// (stands in for the FIFO; serve_batch reads it through an ifstream either way); the
// "<batch>-k<K>.bf" files land in the current directory.  Returns seconds spent in the
// batch loop only (index/mapping load excluded), or < 0 on error.
double
ref_serve_batches(const char* target_fa,
                  const char* target_index,
                  const char* mappings_path,
                  const char* reads_path,
                  const char* reads_index,
                  double mx_max,
                  double subsample_max,
                  int threads,
                  const unsigned* k_values,
                  int nk,
                  const char* const* batch_names,
                  const char* const* ids_files,
                  int n_batches)
{
  std::vector<unsigned> ks(k_values, k_values + nk);
  std::vector<std::string> bf_names;
  for (const auto k : ks) {
    bf_names.push_back("k" + std::to_string(k) + BF_EXTENSION);
  }
  SeqIndex target_seqs_index(target_index, target_fa);
  SeqIndex mapped_seqs_index(reads_index, reads_path);
  AllMappings all_mappings(
    mappings_path, target_seqs_index, MX_THRESHOLD_MIN, MX_THRESHOLD_MAX, mx_max);
  const size_t cbf_bytes = 10ULL * 1024ULL * 1024ULL;
  const size_t bf_bytes = 512ULL * 1024ULL;
  const auto t0 = std::chrono::steady_clock::now();
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads)
  for (int i = 0; i < n_batches; i++) {
    const std::string ready = std::string(batch_names[i]) + "-bfs_ready.tmp";
    // serve_batch std::remove()s its "pipes"; hand it a scratch copy of the ids file.
    const std::string ids_copy = std::string(batch_names[i]) + "-target_ids_input.tmp";
    {
      std::ifstream in(ids_files[i]);
      std::ofstream out(ids_copy);
      out << in.rdbuf();
    }
    serve_batch(target_seqs_index,
                mapped_seqs_index,
                all_mappings,
                cbf_bytes,
                bf_bytes,
                batch_names[i],
                ids_copy,
                ready,
                bf_names,
                4,
                ks,
                subsample_max);
  }
  const auto t1 = std::chrono::steady_clock::now();
  return std::chrono::duration<double>(t1 - t0).count();
}

// What the reference's own AllMappings (src/mappings.cpp) holds after loading `mappings_path`: one line per target of
// the index that has mappings, sorted by id, "<target>\t<read> <read> ...\n" (gp-host-check prints the same from
// goldpolish_b200's loader).  Returns 0, or < 0 when the output cannot be written.
int
ref_mappings_dump(const char* target_fa, const char* target_index, const char* mappings_path, double mx_max, const char* out_path)
{
  SeqIndex target_seqs_index(target_index, target_fa);
  AllMappings all_mappings(mappings_path, target_seqs_index, MX_THRESHOLD_MIN, MX_THRESHOLD_MAX, mx_max);
  std::vector<std::string> ids;
  {
    std::ifstream ifs(target_index);
    std::string tok;
    unsigned long i = 0;
    while (bool(ifs >> tok)) {
      if (i++ % 4 == 0) ids.push_back(tok);
    }
  }
  std::sort(ids.begin(), ids.end());
  ids.erase(std::unique(ids.begin(), ids.end()), ids.end());
  std::ofstream out(out_path);
  if (!out.good()) return -1;
  for (const auto& id : ids) {
    const auto& v = all_mappings.get_mappings(id);
    if (v.empty()) continue;
    out << id << '\t';
    for (size_t j = 0; j < v.size(); j++) out << (j ? " " : "") << v[j];
    out << '\n';
  }
  return 0;
}

} // extern "C"
