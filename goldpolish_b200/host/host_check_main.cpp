// gp-host-check: prints what the host side of the drop-ins decided, without touching a GPU -- for parity tests of the
// feeder logic against the reference's own classes (the test harness dumps the same text from its AllMappings).
//   gp-host-check mappings <targets.fa> <targets.index> <mappings> <mx_max_per_10kbp> [threads]
//       one line per target of the index that has mappings, sorted by id: "<target>\t<read> <read> ...\n"
//       (src/mappings.cpp:15-330 with MX_THRESHOLD_MIN/MAX of src/goldpolish_targeted_bfs.cpp:34-35)
#include "gp_host.hpp"

#include <chrono>

using namespace gph;

// gp-host-check reads <seqs> <seqs.index> <ids file> <slab bytes> [threads]: the sequences named in the ids file, one per
// line, fetched slab by slab the way the server feeds the device (SeqIndex::read_many)
static int reads_mode(int argc, char** argv)
{
  if (argc < 6) return 2;
  const SeqIndex ix = SeqIndex::load(argv[3], argv[2]);
  std::vector<std::string> ids;
  {
    std::ifstream f(argv[4]);
    std::string id;
    while (bool(f >> id)) ids.push_back(id);
  }
  std::vector<const std::string*> order;
  for (const auto& id : ids) order.push_back(&id);
  const uint64_t slab_bytes = std::stoull(argv[5]);
  const unsigned threads = argc > 6 ? unsigned(std::atoi(argv[6])) : 0;
  std::string slab;
  for (size_t first = 0; first < order.size();) {
    size_t n = 0;
    uint64_t bytes = 0;
    while (first + n < order.size() && (n == 0 || bytes + ix.at(*order[first + n]).len <= slab_bytes)) bytes += ix.at(*order[first + n++]).len;
    ix.read_many(order, first, n, slab, threads);
    size_t pos = 0;
    for (size_t i = 0; i < n; i++) {
      const size_t len = ix.at(*order[first + i]).len;
      std::fwrite(slab.data() + pos, 1, len, stdout);
      std::fputc('\n', stdout);
      pos += len;
    }
    first += n;
  }
  return 0;
}

int main(int argc, char** argv)
{
  if (argc > 1 && std::string(argv[1]) == "reads") return reads_mode(argc, argv);
  if (argc < 6 || std::string(argv[1]) != "mappings") {
    std::cerr << "usage: gp-host-check mappings <targets.fa> <targets.index> <mappings> <mx_max_per_10kbp> [threads]\n"
                 "       gp-host-check reads <seqs> <seqs.index> <ids file> <slab bytes> [threads]\n";
    return 2;
  }
  const SeqIndex targets = SeqIndex::load(argv[3], argv[2]);
  const unsigned threads = argc > 6 ? unsigned(std::atoi(argv[6])) : 0;
  const auto t0 = std::chrono::steady_clock::now();
  const Mappings maps(argv[4], targets, 1, 30, std::stod(argv[5]), threads);
  const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  info("mappings loaded in " + std::to_string(s) + " s, " + std::to_string(maps.n_targets()) + " targets");
  std::vector<std::string> ids = targets.order;
  std::sort(ids.begin(), ids.end());
  std::string out;
  for (const auto& id : ids) {
    const auto& v = maps.get(id);
    if (v.empty()) continue;
    out += id;
    out += '\t';
    for (size_t i = 0; i < v.size(); i++) { if (i) out += ' '; out += v[i]; }
    out += '\n';
  }
  std::fwrite(out.data(), 1, out.size(), stdout);
  return 0;
}
