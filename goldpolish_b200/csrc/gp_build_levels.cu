// Counting-Bloom-gated filter build, level-synchronous form (sm_100a).
//
// Same result as fill_bfs (bcgsc/goldpolish src/utils.cpp:96-123) in the reference's order,
// computed WITHOUT walking the stream in order.  The reference's update is
//     c = min over the k-mer's 4 counters;  if (c < thr) every counter equal to c becomes c+1;
//     the k-mer enters the Bloom filter iff the count after the update is >= thr.
// Give every k-mer occurrence of a stream its position t in the reference's order and let
// T_L(x) be the time at which counter x goes from L-1 to L.  Then
//     T_{L+1}(x) = min { t : occurrence t touches x,  thr(t) > L,  t > max_j T_L(x_j(t)) }
// (all four counters had reached L before t, so the minimum is >= L; those still at L move, and
// the first such occurrence is the one that moves x), and occurrence t enters the filter iff
//     t > max_j T_{thr(t)-1}(x_j(t)).
// An occurrence that fails the test at level L fails it at every higher level.  So a stream is
// Lmax = max thr rounds of: "survivors test against T_L" then "survivors atomicMin their time
// into T_{L+1}" -- order-free inside a round, which lets ALL SMs work on ONE stream, whose single
// 40 MiB timestamp array then lives in L2 instead of HBM.  One array serves every level: entries
// carry a 6-bit epoch tag (newer epochs compare smaller, so atomicMin overwrites stale entries,
// and readers treat a stale tag as "never"), so nothing is cleared between levels or streams.
#include "gp_hashing.cuh"

#include <cooperative_groups.h>

namespace cg = cooperative_groups;

namespace gp {

constexpr uint32_t kTimeBits = 26;
constexpr uint32_t kTimeMask = (1u << kTimeBits) - 1u;
constexpr uint32_t kMaxEpoch = 62; // tags 63 - epoch; tag 63 (epoch 0) is the cleared state
constexpr int kLevelWarps = 8;

struct LevelCtx {
  const uint64_t* tf;
  const uint64_t* tr;
  uint32_t lane, gwarp, nwarps;
};

// Walk the steps [s_begin, s_end) of stream (batch, ki) in order and call f(step, entry thr, valid,
// base hash, ci, bi) for each.  A step is 32 consecutive k-mer starts of one read; steps of a stream are
// numbered in the reference's order, which makes (step * 32 + lane) the occurrence time.
template<typename P, typename F>
__device__ __forceinline__ void for_steps(const LevelParams& p, const LevelCtx& c, uint32_t batch, uint32_t ki,
                                          const StreamConsts& sc, uint32_t s_begin, uint32_t s_end, P&& wanted, F&& f)
{
  if (s_begin >= s_end) return;
  const uint32_t* pre = p.step_pre + uint64_t(ki) * (p.n_entries + 1);
  const uint64_t e0 = p.batch_entry_off[batch], e1 = p.batch_entry_off[batch + 1];
  const uint32_t base = pre[e0];
  // entry that holds step s_begin: last e with pre[e] - base <= s_begin
  uint64_t lo = e0, hi = e1;
  while (hi - lo > 1) {
    const uint64_t mid = (lo + hi) >> 1;
    if (pre[mid] - base <= s_begin) lo = mid; else hi = mid;
  }
  uint32_t s = s_begin;
  for (uint64_t e = lo; e < e1 && s < s_end; e++) {
    const uint32_t first = pre[e] - base, last = pre[e + 1] - base;
    if (last <= s) continue;
    const gp_read_entry ent = p.entries[e];
    const uint32_t thr = ent.kmer_threshold - 2u + ki; // utils.cpp:108,121
    const uint32_t len = p.read_len[ent.read_id];
    const uint64_t wbase = p.read_boff[ent.read_id] >> 5;
    const uint32_t npos = len - sc.k + 1;
    // the words of step s+1 are requested before step s is processed (one read per warp is a
    // dependent chain of HBM latencies otherwise)
    uint32_t rs = s - first;
    uint64_t w0 = __ldg(p.pk + wbase + rs), w1 = __ldg(p.pk + wbase + rs + 1);
    uint32_t m0 = __ldg(p.nm + wbase + rs), m1 = __ldg(p.nm + wbase + rs + 1);
    for (; s < last && s < s_end; s++) {
      rs = s - first; // step within the read
      const uint64_t cw0 = w0, cw1 = w1;
      const uint32_t cm0 = m0, cm1 = m1;
      if (s + 1 < last && s + 1 < s_end) {
        w0 = cw1; m0 = cm1; // consecutive steps share a word
        w1 = __ldg(p.pk + wbase + rs + 2);
        m1 = __ldg(p.nm + wbase + rs + 2);
      }
      if (!wanted(s)) continue;
      uint32_t ci[4], bi[4];
      uint64_t h0 = 0;
      const bool valid = hash_from_words(c.tf, c.tr, cw0, cw1, cm0, cm1, rs * 32u, npos, c.lane, sc, ci, bi, &h0);
      f(s, thr, valid, h0, ci, bi);
    }
  }
}

// Survivor lists: occurrences that passed level L.  An entry is five words: the four counter
// indices (24 bits; bit 24 carries bit 21 of the hash, so that the 22-bit filter index is
// recoverable without re-hashing) and time | thr << 26.  Warps reserve list space in chunks of
// kChunk entries (one atomicAdd per chunk instead of one per step); unused tail slots of a
// chunk are marked invalid.
constexpr uint32_t kChunk = 64;
constexpr uint32_t kInvalidMeta = 0xFFFFFFFFu;
struct SurvList {
  uint32_t* w[5];   // w[0..3] packed indices, w[4] meta
  uint32_t* count;  // entries reserved so far (multiple of kChunk)
};
struct SurvCursor {
  uint32_t next, end; // this warp's current chunk [next, end)
};
__device__ __forceinline__ uint32_t pack_index(uint32_t ci, uint32_t bi) { return ci | ((bi >> 21) << 24); }
__device__ __forceinline__ void unpack_index(uint32_t w, uint32_t& ci, uint32_t& bi)
{
  ci = w & 0xFFFFFFu;
  bi = (w & 0x1FFFFFu) | ((w >> 24) << 21);
}
__device__ __forceinline__ void surv_close(const SurvList& l, SurvCursor& cur, uint32_t lane)
{ // invalidate what is left of the warp's chunk
  for (uint32_t i = cur.next + lane; i < cur.end; i += 32) l.w[4][i] = kInvalidMeta;
  cur.next = cur.end = 0;
}
__device__ __forceinline__ void surv_append(const SurvList& l, SurvCursor& cur, bool q, const uint32_t (&ci)[4],
                                            const uint32_t (&bi)[4], uint32_t meta, uint32_t lane)
{
  const uint32_t m = __ballot_sync(0xffffffffu, q);
  if (m == 0u) return;
  const uint32_t n = __popc(m);
  if (cur.end - cur.next < n) { // not enough room: retire the chunk, reserve a new one
    surv_close(l, cur, lane);
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(l.count, kChunk);
    cur.next = __shfl_sync(0xffffffffu, base, 0);
    cur.end = cur.next + kChunk;
  }
  if (q) {
    const uint32_t i = cur.next + __popc(m & ((1u << lane) - 1u));
#pragma unroll
    for (int j = 0; j < 4; j++) l.w[j][i] = pack_index(ci[j], bi[j]);
    l.w[4][i] = meta;
  }
  cur.next += n;
}
// the level test: all four counters carry the current tag and a time before t
__device__ __forceinline__ bool level_test(const uint32_t* __restrict__ V, uint8_t* __restrict__ cbf, uint32_t tag, uint32_t t,
                                           uint32_t L, const uint32_t (&ci)[4])
{
  uint32_t v[4];
#pragma unroll
  for (int j = 0; j < 4; j++) v[j] = __ldcg(V + ci[j]);
  bool reached = true;
  uint32_t mx = 0;
#pragma unroll
  for (int j = 0; j < 4; j++) {
    reached &= (v[j] & ~kTimeMask) == tag;
    mx = max(mx, v[j] & kTimeMask);
    if (cbf && v[j] == (tag | t)) cbf[ci[j]] = (uint8_t)L; // this occurrence moved counter j to level L
  }
  return reached && t > mx;
}

__global__ void __launch_bounds__(kLevelWarps * 32, 3) build_filters_levels_kernel(LevelParams p)
{
  __shared__ uint64_t tf[8 * 256];
  __shared__ uint64_t tr[8 * 256];
  cg::grid_group grid = cg::this_grid();
  fill_hash_tables(tf, tr);
  __syncthreads();
  LevelCtx c;
  c.tf = tf; c.tr = tr;
  c.lane = threadIdx.x & 31u;
  c.gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  c.nwarps = (gridDim.x * blockDim.x) >> 5;
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  uint32_t* __restrict__ V = p.V;
  unsigned long long ops = 0;
  uint32_t epoch = 0; // V arrives cleared (all 0xFFFFFFFF = tag 63)
  SurvList cur, nxt;
  for (int j = 0; j < 5; j++) { cur.w[j] = p.surv + size_t(j) * p.surv_cap; nxt.w[j] = p.surv + size_t(5 + j) * p.surv_cap; }
  cur.count = p.surv_count; nxt.count = p.surv_count + 1;

  for (uint32_t sid = 0; sid < p.n_streams; sid++) {
    const uint32_t lb = sid / p.nk, ki = sid - lb * p.nk;
    const uint32_t batch = p.first_batch + lb;
    const StreamConsts sc = stream_consts(p.k[ki]);
    const uint32_t* pre = p.step_pre + uint64_t(ki) * (p.n_entries + 1);
    const uint32_t n_steps = pre[p.batch_entry_off[batch + 1]] - pre[p.batch_entry_off[batch]];
    if (n_steps == 0) continue;
    const uint32_t lmax = p.batch_max_thr[batch] - 2u + ki; // largest thr of the stream
    uint32_t* __restrict__ bf = p.bf_pool + (uint64_t(batch) * p.nk + ki) * kBfWords;
    uint8_t* __restrict__ cbf = p.cbf_pool ? p.cbf_pool + uint64_t(sid) * kCbfCounters : nullptr;
    // contiguous share of the steps for this warp
    const uint32_t per = (n_steps + c.nwarps - 1) / c.nwarps;
    const uint32_t s_begin = min(n_steps, c.gwarp * per), s_end = min(n_steps, s_begin + per);

    // make sure the epochs of this stream fit below the tag wrap
    if (epoch + lmax + 1 > kMaxEpoch) {
      for (uint64_t i = gtid; i < kCbfCounters; i += gthreads) V[i] = 0xFFFFFFFFu;
      epoch = 0;
      grid.sync();
    }

    // ---- level 0 -> 1: every occurrence writes its time (the first toucher of a counter wins) ----
    epoch++;
    uint32_t tag = (63u - epoch) << kTimeBits;
    if (gtid == 0) *cur.count = 0u;
    for_steps(p, c, batch, ki, sc, s_begin, s_end, [](uint32_t) { return true; },
              [&](uint32_t s, uint32_t thr, bool valid, uint64_t h0, const uint32_t (&ci)[4], const uint32_t (&bi)[4]) {
                const uint32_t t = s * 32u + c.lane;
                if (valid && thr > 0u) {
#pragma unroll
                  for (int j = 0; j < 4; j++) atomicMin(V + ci[j], tag | t);
                  if (thr == 1u) {
#pragma unroll
                    for (int j = 0; j < 4; j++) atomicOr(bf + (bi[j] >> 5), 1u << (bi[j] & 31u));
                  }
                }
                if (valid) ops++;
                (void)h0;
              });
    grid.sync();

    // levels that need a read round: up to lmax-1 for the filter bits (an insert happens at
    // L = thr-1); one more when the counter bytes themselves are wanted (who reached lmax)
    const uint32_t lread = cbf ? lmax : lmax - 1u;
    if (lread >= 1u) {
      // ---- level 1 read, from the sequence: survivors go to a compact list ----
      SurvCursor sc1 = { 0, 0 };
      for_steps(p, c, batch, ki, sc, s_begin, s_end, [](uint32_t) { return true; },
                [&](uint32_t s, uint32_t thr, bool valid, uint64_t h0, const uint32_t (&ci)[4], const uint32_t (&bi)[4]) {
                  const uint32_t t = s * 32u + c.lane;
                  bool q = false;
                  if (valid && thr > 0u) {
                    q = level_test(V, cbf, tag, t, 1u, ci) && thr > 1u;
                    if (q && thr == 2u) { // count after the update reaches thr: Bloom filter insert
#pragma unroll
                      for (int j = 0; j < 4; j++) atomicOr(bf + (bi[j] >> 5), 1u << (bi[j] & 31u));
                    }
                  }
                  surv_append(cur, sc1, q, ci, bi, t | (thr << kTimeBits), c.lane);
                  (void)h0;
                });
      surv_close(cur, sc1, c.lane);
      grid.sync();
    }
    for (uint32_t L = 2; L <= lread; L++) {
      // ---- write: survivors of level L-1 race for T_L of their counters ----
      const uint32_t n_cur = *((volatile uint32_t*)cur.count);
      epoch++;
      tag = (63u - epoch) << kTimeBits;
      if (gtid == 0) *nxt.count = 0u;
      for (uint32_t i = gtid; i < n_cur; i += gthreads) {
        const uint32_t meta = cur.w[4][i];
        if (meta == kInvalidMeta) continue;
        const uint32_t t = meta & kTimeMask;
#pragma unroll
        for (int j = 0; j < 4; j++) atomicMin(V + (cur.w[j][i] & 0xFFFFFFu), tag | t);
      }
      grid.sync();
      // ---- read: who sees all four counters at >= L before its own time? ----
      SurvCursor scn = { 0, 0 };
      for (uint32_t i0 = gtid - c.lane; i0 < n_cur; i0 += gthreads) { // warp-uniform trip count
        const uint32_t i = i0 + c.lane;
        bool q = false;
        uint32_t meta = kInvalidMeta;
        uint32_t ci[4] = { 0, 0, 0, 0 }, bi[4] = { 0, 0, 0, 0 };
        if (i < n_cur) meta = cur.w[4][i];
        if (meta != kInvalidMeta) {
#pragma unroll
          for (int j = 0; j < 4; j++) unpack_index(cur.w[j][i], ci[j], bi[j]);
          const uint32_t t = meta & kTimeMask, thr = meta >> kTimeBits;
          q = level_test(V, cbf, tag, t, L, ci) && thr > L;
          if (q && thr == L + 1u) {
#pragma unroll
            for (int j = 0; j < 4; j++) atomicOr(bf + (bi[j] >> 5), 1u << (bi[j] & 31u));
          }
        }
        surv_append(nxt, scn, q, ci, bi, meta, c.lane);
      }
      surv_close(nxt, scn, c.lane);
      grid.sync();
      SurvList tmp = cur; cur = nxt; nxt = tmp;
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ops += __shfl_xor_sync(0xffffffffu, ops, o);
  if (c.lane == 0 && ops) atomicAdd(p.counters + 0, ops);
}

int levels_max_grid(int sm_count)
{
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, build_filters_levels_kernel, kLevelWarps * 32, 0);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  return sm_count * per_sm;
}

cudaError_t launch_build_filters_levels(const LevelParams& p, int sm_count, cudaStream_t s)
{
  if (p.n_streams == 0) return cudaSuccess;
  LevelParams lp = p;
  void* args[] = { &lp };
  return cudaLaunchCooperativeKernel((const void*)build_filters_levels_kernel, dim3(levels_max_grid(sm_count)),
                                     dim3(kLevelWarps * 32), args, 0, s);
}

} // namespace gp
