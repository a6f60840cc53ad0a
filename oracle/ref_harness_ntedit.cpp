// TEST INFRASTRUCTURE ONLY (oracle/_ref).  In-process access to the reference's OWN ntEdit
// (subprojects/ntedit/ntedit.cpp, #included from where it lies, main() renamed) plus the
// k-chain and 0.75 size guard of scripts/goldpolish-ntedit:20-40 (that script needs `bc`,
// absent here, so its arithmetic is restated: scale=4 truncation, compare < 0.75).
#define main gp_ref_ntedit_main
#include "ntedit.cpp" // NOLINT  (reference source, via -I$(REF)/subprojects/ntedit)
#undef main

#include <sys/stat.h>
#include <sys/wait.h>

#include <chrono>

namespace {

long
file_size(const std::string& p)
{
  struct stat st;
  if (stat(p.c_str(), &st) != 0) {
    return -1;
  }
  return long(st.st_size);
}

int
run_ntedit_once(const std::string& in_fa, const std::string& bf, const std::string& prefix)
{
  // scripts/goldpolish-ntedit:27
  std::vector<std::string> args = { "ntedit-gr", "-f", in_fa, "-r", bf,     "-d5",  "-i5",
                                    "-m1",       "-X0.5", "-Y0.5", "-b",  prefix, "-t1", "-a1" };
  std::vector<char*> argv;
  for (auto& a : args) {
    argv.push_back(const_cast<char*>(a.c_str()));
  }
  argv.push_back(nullptr);
  optind = 0; // re-initialise glibc getopt
  return gp_ref_ntedit_main(int(args.size()), argv.data());
}

} // namespace

extern "C" {

// Equivalent of `goldpolish-ntedit <base> "<bfs>" "<ks>" 0.5 0.5 1 <out>` for one batch.
// Writes the chain files next to <base>.fa exactly as the script names them and copies the
// chosen result to out_path (the script symlinks).  Returns 1 if the guard rejected the
// edit (original kept), 0 otherwise, <0 on error.
int
ref_ntedit_chain(const char* seqs_basename, const char* const* bfs, const unsigned* ks, int nk, const char* out_path)
{
  const std::string base(seqs_basename);
  const long input_size = file_size(base + ".fa");
  std::string prev;
  for (int i = 0; i < nk; i++) {
    const std::string input = prev.empty() ? base : prev;
    const std::string prefix = input + ".k" + std::to_string(ks[i]) + ".X0.5.Y0.5";
    if (run_ntedit_once(input + ".fa", bfs[i], prefix) != 0) {
      return -1;
    }
    prev = prefix + "_edited";
  }
  const long output_size = file_size(prev + ".fa");
  if (input_size <= 0 || output_size < 0) {
    return -2;
  }
  // bc: scale=4; output/input  (truncated to 4 decimals), then  < 0.75
  const long ratio_1e4 = (output_size * 10000L) / input_size;
  const bool skip = ratio_1e4 < 7500L;
  const std::string chosen = skip ? base + ".fa" : prev + ".fa";
  std::ifstream in(chosen, std::ios::binary);
  std::ofstream out(out_path, std::ios::binary);
  out << in.rdbuf();
  return skip ? 1 : 0;
}

// CPU baseline for the edit stage: n batches over `procs` forked workers (ntEdit keeps its
// options in globals, so workers are processes, as in the reference where every batch is
// its own process with -t1).  Returns wall seconds, or < 0 on error.
double
ref_ntedit_chain_many(const char* const* seqs_basenames,
                      const char* const* bfs_flat, // n * nk paths
                      const unsigned* ks,
                      int nk,
                      const char* const* out_paths,
                      int n,
                      int procs)
{
  const auto t0 = std::chrono::steady_clock::now();
  std::vector<pid_t> pids;
  for (int p = 0; p < procs; p++) {
    const pid_t pid = fork();
    if (pid < 0) {
      return -1.0;
    }
    if (pid == 0) {
      int rc = 0;
      for (int i = p; i < n; i += procs) {
        if (ref_ntedit_chain(seqs_basenames[i], bfs_flat + size_t(i) * nk, ks, nk, out_paths[i]) < 0) {
          rc = 1;
        }
      }
      _exit(rc);
    }
    pids.push_back(pid);
  }
  bool ok = true;
  for (const auto pid : pids) {
    int status = 0;
    waitpid(pid, &status, 0);
    ok = ok && WIFEXITED(status) && WEXITSTATUS(status) == 0;
  }
  const auto t1 = std::chrono::steady_clock::now();
  return ok ? std::chrono::duration<double>(t1 - t0).count() : -1.0;
}

// One contig through the reference's kmerizeAndCorrect (ntedit.cpp:1414-1771) against an
// in-memory filter payload, with GoldPolish's options (-d5 -i5 -m1 -X0.5 -Y0.5 -a1 unless
// overridden).  Output record ("\n"-stripped sequence) is returned in out (cap bytes);
// returns its length, -1 if the contig is shorter than min_contig_len (reference drops it,
// :1850), -2 if cap is too small.
long
ref_ntedit_contig(const char* seq,
                  size_t len,
                  const unsigned char* bf_payload,
                  size_t bf_bytes,
                  unsigned k,
                  unsigned hash_num,
                  unsigned max_ins,
                  unsigned max_del,
                  int mode,
                  int mask,
                  float X,
                  float Y,
                  const char* scratch_path,
                  char* out,
                  size_t cap)
{
  opt::k = k;
  opt::h = hash_num;
  opt::max_insertions = max_ins;
  opt::max_deletions = max_del;
  opt::mode = mode;
  opt::mask = mask;
  opt::missing_ratio = X;
  opt::edit_ratio = Y;
  opt::use_ratio = true;
  opt::snv = 0;
  opt::secbf = 0;
  opt::verbose = 0;
  opt::jump = 3;
  opt::insertion_cap =
    static_cast<unsigned>(static_cast<float>(opt::k) * opt::default_insertion_cap_ratio); // :2024
  current_bases_array = polish_bases_array;                                                 // :1995
  if (len < opt::min_contig_len) {
    return -1;
  }
  btllib::KmerBloomFilter bloom(bf_bytes, hash_num, k);
  std::memcpy(bloom.data(), bf_payload, bf_bytes);
  btllib::KmerBloomFilter bloomrep(125, 1, 1); // :2118
  std::string hdr = "c";
  std::string s(seq, len);
  {
    std::ofstream dfout(scratch_path);
    kmerizeAndCorrect(hdr, s, unsigned(len), bloom, bloomrep, dfout);
  }
  std::ifstream in(scratch_path);
  std::string line;
  std::getline(in, line); // header
  std::getline(in, line);
  if (line.size() > cap) {
    return -2;
  }
  std::memcpy(out, line.data(), line.size());
  return long(line.size());
}

} // extern "C"
