// goldpolish-mask / goldpolish-to-upper (B200): drop-ins for bcgsc/goldpolish scripts/goldpolish-mask and
// scripts/goldpolish-to-upper, over gp_prep (csrc/gp_prep.cu).  The binary acts as goldpolish-to-upper when its
// name ends in "to-upper".
//   goldpolish-mask (-s | -n) -k K <seqs | ->          masked records on stdout   (goldpolish-mask:14-41,75-83)
//   goldpolish-to-upper <seqs> <output>                 upper-cased records        (goldpolish-to-upper:5-21)
// Records are written the way btllib::SeqWriter writes FASTA: ">id[ comment]\n" + sequence + "\n".
#include "gp_host.hpp"

using namespace gph;

static int run(const std::string& in_path, const std::string& out_path, int mode, unsigned k, int to_upper)
{
  const auto recs = read_fasta(in_path);
  std::string all;
  std::vector<uint64_t> off(1, 0);
  for (const auto& r : recs) { all += r.seq; off.push_back(all.size()); }
  gp_config cfg;
  gp_default_config(&cfg);
  if (const char* d = std::getenv("GP_DEVICE")) cfg.device = std::atoi(d);
  gp_ctx* ctx = nullptr;
  if (gp_ctx_create(&cfg, &ctx) != GP_OK) die(std::string("gp_ctx_create: ") + gp_last_error(nullptr));
  std::vector<char> out(all.size() + recs.size() + 16);
  std::vector<uint64_t> ooff(recs.size() + 1, 0);
  check_gp(ctx, gp_prep(ctx, uint32_t(recs.size()), all.data(), off.data(), mode, k, to_upper, out.data(), out.size(), ooff.data()),
           "gp_prep");
  gp_ctx_destroy(ctx);
  FILE* f = out_path == "-" ? stdout : std::fopen(out_path.c_str(), "w");
  if (!f) die("cannot open " + out_path);
  for (size_t i = 0; i < recs.size(); i++) {
    std::fputc('>', f);
    std::fwrite(recs[i].name.data(), 1, recs[i].name.size(), f);
    if (!recs[i].comment.empty()) { std::fputc(' ', f); std::fwrite(recs[i].comment.data(), 1, recs[i].comment.size(), f); }
    std::fputc('\n', f);
    std::fwrite(out.data() + ooff[i], 1, size_t(ooff[i + 1] - ooff[i]), f);
    std::fputc('\n', f);
  }
  if (f != stdout) std::fclose(f);
  return 0;
}

int main(int argc, char** argv)
{
  const std::string self = argv[0];
  if (endswith(self, "to-upper")) {
    if (argc != 3) { std::cerr << "usage: goldpolish-to-upper seqs_path output_path\n"; return 2; }
    return run(argv[1], argv[2], 0, 0, 1);
  }
  bool soft = false, hard = false;
  long k = -1;
  std::string path;
  for (int i = 1; i < argc; i++) {
    const std::string a = argv[i];
    if (a == "-s") soft = true;
    else if (a == "-n") hard = true;
    else if (a == "-k" && i + 1 < argc) k = std::atol(argv[++i]);
    else if (a.rfind("-k", 0) == 0 && a.size() > 2) k = std::atol(a.c_str() + 2);
    else if (path.empty()) path = a;
    else { std::cerr << "goldpolish-mask: error: unrecognized arguments: " << a << "\n"; return 2; }
  }
  // argparse's errors (goldpolish-mask:30-36) exit with status 2
  if (k < 0 || path.empty()) { std::cerr << "goldpolish-mask: error: the following arguments are required: -k, seqspath\n"; return 2; }
  if (!soft && !hard) { std::cerr << "goldpolish-mask: error: Either -n or -s must be set\n"; return 2; }
  if (soft && hard) { std::cerr << "goldpolish-mask: error: Both -n and -s cannot be set -- choose one to hard mask OR soft mask.\n"; return 2; }
  if (k < 1 || k > 64) die("goldpolish-mask: -k must be within 1..64 on this path");
  if (path == "-") path = "/dev/stdin";
  return run(path, "-", hard ? 2 : 1, unsigned(k), 0);
}
