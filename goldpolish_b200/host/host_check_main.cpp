// gp-host-check: prints what the host side of the drop-ins decided, without touching a GPU -- for parity tests of the
// feeder logic against the reference's own classes (the test harness dumps the same text from its AllMappings).
//   gp-host-check mappings <targets.fa> <targets.index> <mappings> <mx_max_per_10kbp> [threads]
//       one line per target of the index that has mappings, sorted by id: "<target>\t<read> <read> ...\n"
//       (src/mappings.cpp:15-330 with MX_THRESHOLD_MIN/MAX of src/goldpolish_targeted_bfs.cpp:34-35)
#include "gp_host.hpp"

#include <chrono>

using namespace gph;

int main(int argc, char** argv)
{
  if (argc < 6 || std::string(argv[1]) != "mappings") {
    std::cerr << "usage: gp-host-check mappings <targets.fa> <targets.index> <mappings> <mx_max_per_10kbp> [threads]\n";
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
