// goldpolish-targeted-bfs (B200): drop-in for bcgsc/goldpolish src/goldpolish_targeted_bfs.cpp.
//
// Same argv (:250-268), same named-pipe protocol in the current directory (:148-244), same
// "<batch>-k<K>.bf" outputs (:79-84,138-140).  Differences are internal only: the mapped reads
// are uploaded to the GPU once, 2-bit packed; batch requests that are pending at the same time
// are built by ONE gp_build_filters call (one warp per (batch, k) stream).
#include "gp_host.hpp"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <memory>
#include <mutex>
#include <thread>

#include <sys/stat.h>
#include <sys/types.h>

using namespace gph;

namespace {

const unsigned MX_THRESHOLD_MIN = 1, MX_THRESHOLD_MAX = 30; // :34-35
const std::string BATCH_NAME_INPUT_PIPE = "batch_name_input";
const std::string BATCH_TARGET_IDS_INPUT_READY_PIPE = "batch_target_ids_input_ready";
const std::string END_SYMBOL = "x";

void make_pipe(const std::string& p)
{ // src/utils.cpp:72-78
  if (mkfifo(p.c_str(), S_IRUSR | S_IWUSR) != 0) die("mkfifo failed: " + std::string(std::strerror(errno)));
}
std::string read_pipe(const std::string& p)
{ // :80-87
  std::string s;
  std::ifstream in(p);
  in >> s;
  return s;
}
void confirm_pipe(const std::string& p)
{ // :89-94
  std::ofstream o(p);
  o << "1" << std::endl;
}
void bind_to_parent()
{ // :55-70
  std::thread([] {
    while (getppid() != 1) std::this_thread::sleep_for(std::chrono::seconds(1));
    std::exit(EXIT_FAILURE);
  }).detach();
}

struct Request {
  std::string batch;
  std::vector<gp_read_entry> entries;
  std::vector<uint8_t> payload; // nk * GP_BF_BYTES once built
  bool done = false;
};

struct Server {
  gp_ctx* ctx = nullptr;
  std::vector<unsigned> ks;
  std::mutex mu;
  std::condition_variable cv_work, cv_done;
  std::deque<std::shared_ptr<Request>> queue;
  bool stopping = false;

  // GPU worker: builds every request that is pending in one call.  The payloads land in page-locked memory
  // that the build kernel fills filter by filter as they become final (gp_build_output_host): no bulk copy
  // after the build, and what bfs[i]->save() writes (:138-140) is ready when the call returns.
  void gpu_loop()
  {
    uint8_t* pinned = nullptr;
    size_t pinned_cap = 0;
    for (;;) {
      std::vector<std::shared_ptr<Request>> todo;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_work.wait(lk, [&] { return stopping || !queue.empty(); });
        if (queue.empty() && stopping) { gp_build_output_host(ctx, nullptr); gp_host_free(pinned); return; }
        todo.assign(queue.begin(), queue.end());
        queue.clear();
      }
      std::vector<uint64_t> off(1, 0);
      std::vector<gp_read_entry> ents;
      for (const auto& r : todo) {
        ents.insert(ents.end(), r->entries.begin(), r->entries.end());
        off.push_back(ents.size());
      }
      const size_t need = todo.size() * ks.size() * GP_BF_BYTES;
      std::vector<uint8_t> pageable;
      uint8_t* out = nullptr;
      if (need > pinned_cap) {
        gp_build_output_host(ctx, nullptr);
        gp_host_free(pinned);
        pinned_cap = need + need / 2;
        pinned = static_cast<uint8_t*>(gp_host_alloc(pinned_cap));
        if (!pinned) pinned_cap = 0;
        else check_gp(ctx, gp_build_output_host(ctx, pinned), "gp_build_output_host");
      }
      if (pinned) out = pinned;
      else { pageable.resize(need); out = pageable.data(); } // no page-locked memory to be had: plain copy
      check_gp(ctx, gp_build_filters(ctx, uint32_t(todo.size()), off.data(), ents.data(), out), "gp_build_filters");
      {
        std::lock_guard<std::mutex> lk(mu);
        for (size_t i = 0; i < todo.size(); i++) {
          todo[i]->payload.assign(out + i * ks.size() * GP_BF_BYTES, out + (i + 1) * ks.size() * GP_BF_BYTES);
          todo[i]->done = true;
        }
      }
      cv_done.notify_all();
    }
  }
};

} // namespace

int main(int argc, char** argv)
{
  if (argc < 8) die("Wrong args."); // :253
  bind_to_parent();
  int a = 1;
  const std::string target_seqs = argv[a++], target_index_path = argv[a++], mappings_path = argv[a++];
  const std::string reads_path = argv[a++], reads_index_path = argv[a++];
  const double mx_max_per_10kbp = std::stod(argv[a++]);
  const double subsample_max_per_10kbp = std::stod(argv[a++]);
  const int threads = std::stoi(argv[a++]);
  (void)threads; // batches are concurrent on the GPU, not on host threads
  std::vector<unsigned> ks;
  while (a < argc) ks.push_back(unsigned(std::stoi(argv[a++])));
  if (ks.empty() || ks.size() > GP_MAX_K_VALUES) die("need 1..8 k values");
  for (const unsigned k : ks) // the device hashes 4 packed bases per table lookup and keeps a k-mer in one 64-bit word
    if (k < 4 || k > 32 || k % 4 != 0) die("k = " + std::to_string(k) + " is not supported: k must be a multiple of 4 within 4..32");

  info("Loading index from " + target_index_path);
  const SeqIndex targets = SeqIndex::load(target_index_path, target_seqs);
  info("Loading index from " + reads_index_path);
  const SeqIndex reads = SeqIndex::load(reads_index_path, reads_path);
  info("Loading mappings from " + mappings_path);
  const Mappings maps(mappings_path, targets, MX_THRESHOLD_MIN, MX_THRESHOLD_MAX, mx_max_per_10kbp);

  gp_config cfg;
  gp_default_config(&cfg);
  if (const char* d = std::getenv("GP_DEVICE")) cfg.device = std::atoi(d);
  cfg.nk = uint32_t(ks.size());
  for (size_t i = 0; i < ks.size(); i++) cfg.k[i] = ks[i];
  Server srv;
  srv.ks = ks;
  if (gp_ctx_create(&cfg, &srv.ctx) != GP_OK) die(std::string("gp_ctx_create: ") + gp_last_error(nullptr));

  // every read that some target maps: uploaded once (SeqIndex::get_seq<1> re-reads per use)
  std::unordered_map<std::string, uint32_t> read_slot;
  {
    std::string all, one;
    std::vector<uint64_t> off(1, 0);
    for (const auto& kv : maps.all())
      for (const auto& id : kv.second)
        if (read_slot.find(id) == read_slot.end()) {
          read_slot.emplace(id, uint32_t(off.size() - 1));
          reads.read_seq(id, one);
          all += one;
          off.push_back(all.size());
        }
    info("Uploading " + std::to_string(off.size() - 1) + " mapped reads (" + std::to_string(all.size()) + " bases)");
    check_gp(srv.ctx, gp_reads_upload(srv.ctx, all.data(), off.data(), off.size() - 1), "gp_reads_upload");
  }

  make_pipe(BATCH_NAME_INPUT_PIPE);
  make_pipe(BATCH_TARGET_IDS_INPUT_READY_PIPE);
  info("Accepting batch names at " + BATCH_NAME_INPUT_PIPE);
  std::thread gpu(&Server::gpu_loop, &srv);
  std::vector<std::thread> workers;
  for (;;) { // process_batch_name, :148-195
    const std::string batch = read_pipe(BATCH_NAME_INPUT_PIPE);
    if (batch.empty() || batch == END_SYMBOL) break;
    const std::string ids_pipe = batch + "-target_ids_input", ready_pipe = batch + "-bfs_ready";
    make_pipe(ids_pipe);
    make_pipe(ready_pipe);
    confirm_pipe(BATCH_TARGET_IDS_INPUT_READY_PIPE);
    workers.emplace_back([&, batch, ids_pipe, ready_pipe] { // serve_batch, :55-146
      auto req = std::make_shared<Request>();
      req->batch = batch;
      {
        std::ifstream in(ids_pipe);
        std::string id;
        while (bool(in >> id) && id != END_SYMBOL) {
          const uint64_t tlen = targets.at(id).len;
          const auto& mapped = maps.get(id);
          if (mapped.empty()) continue;
          const Selection sel = select_reads(mapped, reads, tlen, subsample_max_per_10kbp);
          for (const auto& r : sel.reads) req->entries.push_back(gp_read_entry{ read_slot.at(r), uint32_t(sel.kmer_threshold) });
        }
      }
      {
        std::unique_lock<std::mutex> lk(srv.mu);
        srv.queue.push_back(req);
        srv.cv_work.notify_one();
        srv.cv_done.wait(lk, [&] { return req->done; });
      }
      for (size_t i = 0; i < ks.size(); i++)
        bf_format::save(batch + "-k" + std::to_string(ks[i]) + ".bf", req->payload.data() + i * GP_BF_BYTES, GP_BF_BYTES,
                        GP_HASH_NUM, ks[i]);
      confirm_pipe(ready_pipe);
      std::remove(ids_pipe.c_str());
      std::remove(ready_pipe.c_str());
    });
  }
  for (auto& w : workers) w.join();
  {
    std::lock_guard<std::mutex> lk(srv.mu);
    srv.stopping = true;
  }
  srv.cv_work.notify_all();
  gpu.join();
  std::remove(BATCH_NAME_INPUT_PIPE.c_str());
  std::remove(BATCH_TARGET_IDS_INPUT_READY_PIPE.c_str());
  gp_ctx_destroy(srv.ctx);
  info("Targeted BF builder done!");
  return 0;
}
