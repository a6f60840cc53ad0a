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
//
// "Read level L" and "write level L+1" are ONE round over two alternating timestamp arrays (T_L lives in array
// (L + 1) & 1), and round 0 already puts every occurrence on its warp's list: max thr rounds per stream.
//
// Rounds of one stream are separated by grid-wide barriers.  To keep the SMs busy while a barrier
// drains, the grid works on `n_slots` streams at a time (each with its own timestamp array and
// survivor lists; 2 x 40 MiB still sits in the 126 MB L2): a CTA runs one round of slot 0,
// ARRIVES at slot 0's barrier, runs one round of slot 1, arrives, and only then WAITS for slot
// 0's barrier -- by which time the other CTAs have normally arrived (split-phase barrier).  Every
// CTA takes the same deterministic sequence of (slot, stream, round) decisions from uniform data.
#include "gp_hashing.cuh"

#include <algorithm>
#include <cstdlib>

namespace gp {

// list entries: time in the low 26 bits of the meta word, thr above
constexpr uint32_t kTimeBits = 26;
constexpr uint32_t kTimeMask = (1u << kTimeBits) - 1u;
// timestamp entries: tag << time_bits | time, where time_bits (LevelParams) is just wide enough for the longest
// stream of the launch: 26 bits leave 62 epochs between two clear rounds, 22 bits (3 M k-mers) a thousand.
// Tags count down (newer epochs compare smaller); the all-ones tag (epoch 0) is the cleared state.
constexpr int kLevelWarps = 8;

struct LevelCtx {
  const uint64_t* tf;
  const uint64_t* tr;
  uint32_t lane, gwarp, nwarps;
  // this warp's share of every stream: the fraction [cum[0], cum[1]) of 2^32 (all warps together tile [0, 2^32]);
  // kept in shared memory, read where needed
  const uint64_t* cum;
};

// A step is 32 consecutive k-mer starts of one read; steps of a stream are numbered in the
// reference's order, which makes (step * 32 + lane) the occurrence time.  A warp's share of a stream is a
// fraction [cum[0], cum[1]) of it -- equal fractions to begin with, then proportional to the speed measured for
// the warp's SM (the same work takes 88..106 us depending on where the SM sits) -- cut into kRuns runs that are
// spread over the whole time axis: the stream is kRuns "rows" (time slabs) and the warp takes its fraction of
// every row, the row boundaries dithered by r / kRuns so that the roundings of the rows do not add up (at
// 5 steps per warp every row gives a warp 0 or 1 step).  Survivors crowd towards late times; contiguous shares
// left the early warps idle in every list round.  anchor[] maps a step to its read entry; lane r fetches the
// head of run r, so that the dependent loads of all runs overlap.
constexpr uint32_t kRuns = 8;
// steps before the warp's share (its survivor list starts there, + one slack slot per row and warp)
__device__ __forceinline__ uint64_t warp_list_base(uint32_t n_steps, const LevelCtx& c)
{
  return ((uint64_t(n_steps) * c.cum[0]) >> 32) + uint64_t(c.gwarp) * (kRuns + 1u);
}
// run of row r (0..kRuns-1) of this warp: first step and length (0: none)
__device__ __forceinline__ void run_of_row(uint32_t r, uint32_t n_steps, const LevelCtx& c, uint32_t& s0, uint32_t& len)
{
  const uint32_t r0 = uint32_t(uint64_t(n_steps) * r / kRuns), r1 = uint32_t(uint64_t(n_steps) * (r + 1u) / kRuns);
  const uint64_t rowlen = r1 - r0, phase = (uint64_t(r) << 32) / kRuns;
  // boundaries floor(rowlen * cum + phase): 0 for cum = 0, rowlen for cum = 2^32, the same on both sides of a warp border
  const uint32_t a = uint32_t((rowlen * c.cum[0] + phase) >> 32), b = uint32_t((rowlen * c.cum[1] + phase) >> 32);
  s0 = r0 + a;
  len = r < kRuns ? b - a : 0u;
}

// Calls f(step, entry thr, valid, ci, bi) for every step of this warp's runs.
template<typename F>
__device__ __forceinline__ void for_runs(const LevelParams& p, const LevelCtx& c, uint32_t batch, uint32_t ki,
                                         const StreamConsts& sc, uint32_t n_steps, F&& f)
{
  const uint32_t* pre = p.step_pre + uint64_t(ki) * (p.n_entries + 1);
  const uint64_t e0 = p.batch_entry_off[batch], e1 = p.batch_entry_off[batch + 1];
  const uint32_t base = __ldg(pre + e0);
  const uint16_t* anchor = p.anchor + uint64_t(ki) * p.anchor_stride + base;
  uint32_t h_e = 0, h_first = 0, h_last = 0, h_thr = 0, h_len = 0, h_wlo = 0, h_whi = 0;
  {
    uint32_t s0 = 0, len = 0;
    if (c.lane < kRuns) run_of_row(c.lane, n_steps, c, s0, len);
    if (len) {
      h_e = __ldg(anchor + s0);
      const uint64_t e = e0 + h_e;
      h_first = __ldg(pre + e) - base; h_last = __ldg(pre + e + 1) - base;
      const gp_read_entry ent = p.entries[e];
      h_thr = ent.kmer_threshold;
      h_len = p.read_len[ent.read_id];
      const uint64_t wb = p.read_boff[ent.read_id] >> 5;
      h_wlo = uint32_t(wb); h_whi = uint32_t(wb >> 32);
    }
  }
#pragma unroll 1
  for (uint32_t r = 0; r < kRuns; r++) {
    uint32_t s0, len;
    run_of_row(r, n_steps, c, s0, len);
    if (len == 0) continue;
    const uint32_t s_end = s0 + len;
    const uint64_t e_head = e0 + __shfl_sync(0xffffffffu, h_e, r);
    uint32_t first = __shfl_sync(0xffffffffu, h_first, r), last = __shfl_sync(0xffffffffu, h_last, r);
    uint32_t kthr = __shfl_sync(0xffffffffu, h_thr, r), len_r = __shfl_sync(0xffffffffu, h_len, r);
    uint64_t wbase = uint64_t(__shfl_sync(0xffffffffu, h_wlo, r)) | (uint64_t(__shfl_sync(0xffffffffu, h_whi, r)) << 32);
    uint32_t s = s0;
    for (uint64_t e = e_head; e < e1 && s < s_end; e++) {
      if (e != e_head) { // the run crosses into the next read
        first = __ldg(pre + e) - base; last = __ldg(pre + e + 1) - base;
        if (last <= s) continue;
        const gp_read_entry ent = p.entries[e];
        kthr = ent.kmer_threshold;
        len_r = p.read_len[ent.read_id];
        wbase = p.read_boff[ent.read_id] >> 5;
      }
      const uint32_t thr = kthr - 2u + ki; // utils.cpp:108,121
      const uint32_t npos = len_r - sc.k + 1;
      // the words of step s+1 are requested before step s is processed
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
        uint32_t ci[4], bi[4];
        const bool valid = hash_from_words(c.tf, c.tr, cw0, cw1, cm0, cm1, rs * 32u, npos, c.lane, sc, ci, bi);
        f(s, thr, valid, ci, bi);
      }
    }
  }
}

// anchor[ki][s] = entry (relative to its batch's first entry) that holds global step s of k index ki
__global__ void fill_anchor_kernel(const uint32_t* __restrict__ step_pre, const uint16_t* __restrict__ entry_rel,
                                   uint16_t* __restrict__ anchor, uint32_t n_entries, uint32_t nk, uint64_t anchor_stride)
{
  const uint64_t w = (uint64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31u;
  if (w >= uint64_t(n_entries) * nk) return;
  const uint32_t ki = uint32_t(w / n_entries), e = uint32_t(w - uint64_t(ki) * n_entries);
  const uint32_t* pre = step_pre + uint64_t(ki) * (n_entries + 1);
  const uint32_t a = pre[e], b = pre[e + 1];
  const uint16_t v = entry_rel[e];
  for (uint32_t s = a + lane; s < b; s += 32) anchor[uint64_t(ki) * anchor_stride + s] = v;
}

void launch_fill_anchor(const uint32_t* step_pre, const uint16_t* entry_rel, uint16_t* anchor, uint32_t n_entries,
                        uint32_t nk, uint64_t anchor_stride, cudaStream_t s)
{
  const uint64_t warps = uint64_t(n_entries) * nk;
  if (warps == 0) return;
  fill_anchor_kernel<<<uint32_t((warps + 7) / 8), 256, 0, s>>>(step_pre, entry_rel, anchor, n_entries, nk, anchor_stride);
}

// Survivor lists: occurrences that passed level L.  An entry is five words: the four counter
// indices (24 bits; bit 24 carries bit 21 of the hash, so that the 22-bit filter index is
// recoverable without re-hashing) and time | thr << 26.  Lists are WARP-PRIVATE: a warp keeps the
// survivors of its own share of the stream in its own region (share * 32 entries) and compacts
// them in place level after level -- no list counters, no reservation atomics, no second buffer.
struct SurvList {
  uint32_t* w[5];   // w[0..3] packed indices, w[4] meta; already offset to the warp's region
};
__device__ __forceinline__ uint32_t pack_index(uint32_t ci, uint32_t bi) { return ci | ((bi >> 21) << 24); }
__device__ __forceinline__ void unpack_index(uint32_t w, uint32_t& ci, uint32_t& bi)
{
  ci = w & 0xFFFFFFu;
  bi = (w & 0x1FFFFFu) | ((w >> 24) << 21);
}
// append the lanes with q set at position cnt of the warp's list (pw = the four packed indices)
__device__ __forceinline__ void surv_append(const SurvList& l, uint32_t& cnt, bool q, const uint32_t (&pw)[4], uint32_t meta,
                                            uint32_t lane)
{
  const uint32_t m = __ballot_sync(0xffffffffu, q);
  if (q) {
    const uint32_t i = cnt + __popc(m & ((1u << lane) - 1u));
#pragma unroll
    for (int j = 0; j < 4; j++) l.w[j][i] = pw[j];
    l.w[4][i] = meta;
  }
  cnt += __popc(m);
}
__device__ __forceinline__ void surv_append(const SurvList& l, uint32_t& cnt, bool q, const uint32_t (&ci)[4],
                                            const uint32_t (&bi)[4], uint32_t meta, uint32_t lane)
{
  uint32_t pw[4];
#pragma unroll
  for (int j = 0; j < 4; j++) pw[j] = pack_index(ci[j], bi[j]);
  surv_append(l, cnt, q, pw, meta, lane);
}
// fire-and-forget atomics (REDG): spelled in PTX because ptxas keeps the returning form (ATOMG
// with a dead destination) for atomicMin/atomicOr in this kernel
__device__ __forceinline__ void red_min(uint32_t* p, uint32_t v)
{
  asm volatile("red.relaxed.gpu.global.min.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void red_or(uint32_t* p, uint32_t v)
{
  asm volatile("red.relaxed.gpu.global.or.b32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

__device__ __forceinline__ void bf_insert(uint32_t* __restrict__ bf, const uint32_t (&bi)[4])
{
  uint32_t w[4]; // the filter is 512 KiB and mostly hit by repeats of the same k-mers: look first
#pragma unroll
  for (int j = 0; j < 4; j++) w[j] = __ldcg(bf + (bi[j] >> 5));
#pragma unroll
  for (int j = 0; j < 4; j++)
    if (!(w[j] & (1u << (bi[j] & 31u)))) red_or(bf + (bi[j] >> 5), 1u << (bi[j] & 31u));
}

// ---- split-phase grid barrier: one monotone counter per slot ----
// polling uses a relaxed load (an acquire load costs an L1 invalidate -- CCTL.IVALL -- per poll);
// one fence after the loop orders the data reads of the next round behind it
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p)
{
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long globaltimer_ns()
{
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

// A stream's filter is final (its last barrier has drained on this CTA): stream it to the caller's pinned
// host buffer if there is one -- every CTA copies its slice straight over PCIe, under the following rounds, so
// that no bulk D2H is left at the end -- and tell the edit kernel (gp_pipeline_run).
__device__ __forceinline__ void filter_final(const LevelParams& p, uint32_t batch, uint32_t ki, uint32_t gtid, uint32_t gthreads)
{
  if (p.bf_host) {
    const uint64_t o = (uint64_t(batch) * p.nk + ki) * kBfWords;
    const uint4* __restrict__ src = reinterpret_cast<const uint4*>(p.bf_pool + o);
    uint4* __restrict__ dst = reinterpret_cast<uint4*>(p.bf_host + o);
    for (uint32_t i = gtid; i < kBfWords / 4u; i += gthreads) dst[i] = __ldcg(src + i);
  }
  if (p.batch_done && gtid == 0) {
    atomicAdd(p.batch_done + batch, 1u);
    atomicAdd(p.batch_done + p.n_batches_total, 1u); // progress beacon for the edit kernel's watchdog
  }
}

// sum over the CTA, the same value in every thread (scratch: one slot per warp)
__device__ __forceinline__ unsigned long long block_sum(unsigned long long v, unsigned long long* scratch)
{
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31u) == 0u) scratch[threadIdx.x >> 5] = v;
  __syncthreads();
  unsigned long long t = 0;
  for (uint32_t w = 0; w < (blockDim.x >> 5); w++) t += scratch[w];
  return t;
}

enum : uint32_t { PH_CLEAR = 0, PH_L0 = 1, PH_READ = 2 };

struct SlotState {      // uniform over the grid; written by thread 0 of each CTA between rounds
  uint32_t sid;         // wave-local stream, >= n_streams when the slot has run dry
  uint32_t phase, L, lread, epoch, tag, tag_next, n_steps, batch;
  uint32_t ord, publish, recal; // streams begun in this slot; this stream's round 0 publishes / re-reads the SM speeds
  uint32_t done_b1, done_ki; // batch + 1 and k index of the slot's previous stream while its last barrier drains (0: none)
  unsigned long long target; // barrier count that must be reached before the slot's next round
};

constexpr int kMaxSlots = 3;

// 3 CTAs of 8 warps per SM (80 registers); a 64-register build with 4 CTAs spills and measured 8 % slower
__global__ void __launch_bounds__(kLevelWarps * 32, 3) build_filters_levels_kernel(LevelParams p)
{
  __shared__ uint64_t tf[8 * 256];
  __shared__ uint64_t tr[8 * 256];
  __shared__ SlotState slots[kMaxSlots];
  __shared__ uint32_t next_sid;
  __shared__ uint32_t warp_cnt[kMaxSlots][kLevelWarps]; // survivors in each warp's private list
  __shared__ uint64_t cum_sh[kLevelWarps + 1];          // share boundaries of this CTA's warps (fractions of 2^32)
  __shared__ unsigned long long cal_ns, cal_steps;      // round-0 work of this CTA since its last speed publication
  __shared__ unsigned long long red_sh[3][kLevelWarps];
  __shared__ unsigned long long diag[15];               // round-time diagnostics of this CTA (thread 0)
  if (p.batch_done) asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); // the edit kernel may join us now
  if (blockIdx.x == 0 && threadIdx.x == 0) p.counters[20] = globaltimer_ns();
  fill_hash_tables(tf, tr);
  LevelCtx c;
  c.tf = tf; c.tr = tr;
  c.lane = threadIdx.x & 31u;
  c.gwarp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  c.nwarps = (gridDim.x * blockDim.x) >> 5;
  c.cum = cum_sh + (threadIdx.x >> 5);                   // equal shares (below) until speeds have been measured
  const uint32_t gtid = blockIdx.x * blockDim.x + threadIdx.x, gthreads = gridDim.x * blockDim.x;
  const uint32_t tb = p.time_bits, vmask = (1u << tb) - 1u, maxtag = (1u << (32u - tb)) - 1u;
  unsigned long long ops = 0, list_seen = 0;
  if (threadIdx.x == 0) { cal_ns = 0; cal_steps = 0; }
  if (threadIdx.x < 15) diag[threadIdx.x] = 0;
  if (threadIdx.x <= uint32_t(kLevelWarps)) cum_sh[threadIdx.x] = (uint64_t(blockIdx.x * kLevelWarps + threadIdx.x) << 32) / c.nwarps;

  // stream -> (n_steps, lread); called by thread 0 only
  // stream_tab[sid] = {steps, largest thr, batch} in launch order, fetched one stream ahead (thread 0's registers)
  uint4 nx_info = make_uint4(0u, 0u, 0u, 0u);
  if (threadIdx.x == 0 && p.n_streams) nx_info = __ldg(p.stream_tab);
  auto begin_stream = [&](SlotState& S) {
    for (;;) {
      S.sid = next_sid++;
      if (S.sid >= p.n_streams) return;
      const uint4 info = nx_info;
      if (S.sid + 1u < p.n_streams) nx_info = __ldg(p.stream_tab + S.sid + 1u);
      const uint32_t ki = S.sid % p.nk, batch = info.z;
      S.n_steps = info.x;
      S.batch = batch;
      if (S.n_steps == 0) { // nothing to insert: the (zeroed) filter is final
        if (p.batch_done && blockIdx.x == 0) { atomicAdd(p.batch_done + batch, 1u); atomicAdd(p.batch_done + p.n_batches_total, 1u); }
        continue;
      }
      const uint32_t lmax = info.y; // largest thr of the stream (kmer_threshold - 2 + k index, utils.cpp:108,121)
      (void)ki;
      // levels that need a read round: up to lmax-1 for the filter bits (an insert happens at
      // L = thr-1); one more when the counter bytes themselves are wanted (who reached lmax)
      S.lread = p.cbf_pool ? lmax : lmax - 1u;
      S.ord++;
      S.publish = p.weighted && (S.ord & 63u) == 4u; // after 4 streams, then every 64: publish the CTA's speed ...
      S.recal = p.weighted && (S.ord & 63u) == 5u;   // ... and the next stream starts with re-weighted shares
      if (S.epoch + lmax + 1 > maxtag - 1u) { S.phase = PH_CLEAR; return; } // the tags of this stream would wrap
      S.phase = PH_L0;
      S.epoch++;
      S.tag = (maxtag - S.epoch) << tb;
      return;
    }
  };
  if (threadIdx.x == 0) {
    next_sid = 0;
    for (uint32_t s = 0; s < p.n_slots; s++) {
      slots[s].epoch = 0; slots[s].target = 0; // V arrives cleared (all 0xFFFFFFFF = tag 63)
      slots[s].done_b1 = 0; slots[s].done_ki = 0; slots[s].ord = 0; slots[s].publish = 0; slots[s].recal = 0;
      begin_stream(slots[s]);
    }
  }

  for (;;) {
    bool any = false;
    for (uint32_t sl = 0; sl < p.n_slots; sl++) {
      __syncthreads(); // slot states are stable from here to the next __syncthreads
      SlotState& S = slots[sl];
      if (S.sid >= p.n_streams) continue;
      any = true;
      unsigned long long* bar = p.bars + sl;
      unsigned long long t_a = 0, t_b = 0;
      // the warp's survivor list (room for every occurrence of its runs) is its own: the first entries of a list
      // round are fetched BEFORE the barrier wait, they do not depend on the other CTAs
      const uint32_t wib = threadIdx.x >> 5;
      SurvList lst;
      {
        uint32_t* base = p.surv + uint64_t(sl) * 5u * p.surv_cap + warp_list_base(S.n_steps, c) * 32u;
#pragma unroll
        for (int j = 0; j < 5; j++) lst.w[j] = base + size_t(j) * p.surv_cap;
      }
      const uint32_t list_cnt = S.phase == PH_READ ? warp_cnt[sl][wib] : 0u;
      uint32_t nx_pw[2][4], nx_meta[2]; // entries of the next iteration of the list round
      auto fetch_entries = [&](uint32_t r) {
#pragma unroll
        for (int e = 0; e < 2; e++) {
          const uint32_t i = r + 32u * e + c.lane;
          nx_meta[e] = 0;
#pragma unroll
          for (int j = 0; j < 4; j++) nx_pw[e][j] = 0;
          if (i < list_cnt) {
            nx_meta[e] = __ldcg(lst.w[4] + i);
#pragma unroll
            for (int j = 0; j < 4; j++) nx_pw[e][j] = __ldcg(lst.w[j] + i);
          }
        }
      };
      fetch_entries(0);
      if (threadIdx.x == 0) {
        t_a = globaltimer_ns();
        const unsigned long long target = S.target;
        while (ld_relaxed_u64(bar) < target) { }
        __threadfence();
        t_b = globaltimer_ns();
      }
      __syncthreads(); // the slot's previous round is complete everywhere

      const uint32_t sid = S.sid, phase = S.phase, tag = S.tag, L = S.L, n_steps = S.n_steps, lread = S.lread;
      if (phase == PH_L0 && S.recal) {
        // every CTA published the rate of its round-0 passes (steps per time) a few rounds ago: shares become
        // proportional to them.  Integer sums, so that every CTA derives the very same boundaries.
        unsigned long long tot = 0;
        for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) tot += __ldcg(p.speed + i);
        tot = block_sum(tot, red_sh[0]);
        const unsigned long long mean = max(1ull, tot / gridDim.x), lo = max(1ull, mean * 7ull / 10ull), hi = mean * 14ull / 10ull + 1ull;
        unsigned long long all = 0, before = 0, mine = 0;
        for (uint32_t i = threadIdx.x; i < gridDim.x; i += blockDim.x) {
          const unsigned long long v = min(max((unsigned long long)__ldcg(p.speed + i), lo), hi);
          all += v;
          if (i < blockIdx.x) before += v;
          if (i == blockIdx.x) mine = v;
        }
        all = block_sum(all, red_sh[0]); before = block_sum(before, red_sh[1]); mine = block_sum(mine, red_sh[2]);
        if (threadIdx.x <= uint32_t(kLevelWarps))
          cum_sh[threadIdx.x] = ((before * kLevelWarps + mine * threadIdx.x) << 32) / (all * kLevelWarps);
        __syncthreads();
        uint32_t* base = p.surv + uint64_t(sl) * 5u * p.surv_cap + warp_list_base(n_steps, c) * 32u;
#pragma unroll
        for (int j = 0; j < 5; j++) lst.w[j] = base + size_t(j) * p.surv_cap;
      }
      if (S.done_b1) filter_final(p, S.done_b1 - 1u, S.done_ki, gtid, gthreads); // the slot's previous stream is complete everywhere
      const uint32_t ki = sid % p.nk, batch = S.batch;
      // two timestamp arrays per slot: T_L lives in array (L + 1) & 1
      uint32_t* __restrict__ V0 = p.V + uint64_t(sl) * 2u * kCbfCounters;
      uint32_t* __restrict__ V = V0 + (phase == PH_READ ? ((L + 1u) & 1u) * kCbfCounters : 0u);
      uint32_t* __restrict__ Vn = V0 + (L & 1u) * kCbfCounters; // where T_{L+1} goes
      const uint32_t tag_next = S.tag_next;
      uint32_t* __restrict__ bf = p.bf_pool + (uint64_t(batch) * p.nk + ki) * kBfWords;
      uint8_t* __restrict__ cbf = p.cbf_pool ? p.cbf_pool + uint64_t(sid) * kCbfCounters : nullptr;
      if (phase == PH_CLEAR) {
        for (uint64_t i = gtid; i < kCbfCounters * 2u; i += gthreads) V0[i] = 0xFFFFFFFFu;
      } else if (phase == PH_L0) {
        // ---- round 0: every occurrence writes its time into T_1 (the first toucher of a counter wins) and goes
        // to the warp's list, so that level 1 is an ordinary list round (no second hashing pass) ----
        const StreamConsts sc = stream_consts(p.k[ki]);
        uint32_t cnt = 0;
        for_runs(p, c, batch, ki, sc, n_steps,
                 [&](uint32_t s, uint32_t thr, bool valid, const uint32_t (&ci)[4], const uint32_t (&bi)[4]) {
                   const uint32_t t = s * 32u + c.lane;
                   const bool q = valid && thr > 0u;
                   if (q) {
#pragma unroll
                     for (int j = 0; j < 4; j++) red_min(V + ci[j], tag | t);
                     if (thr == 1u) bf_insert(bf, bi);
                   }
                   if (valid) ops++;
                   surv_append(lst, cnt, q && thr > 1u, ci, bi, t | (thr << kTimeBits), c.lane);
                 });
        if (c.lane == 0) warp_cnt[sl][wib] = cnt;
      } else {
        // ---- list round: who sees all four counters at >= L before its own time?  Survivors with thr = L + 1 enter
        // the filter and, unless it is the stream's last round, race for T_{L+1} right away in the other array; the
        // list is compacted in place ----
        if (c.lane == 0) list_seen += list_cnt; // (diagnostic, flushed once at the end)
        // two entries per lane and iteration: their list words, then their timestamps, are all in flight together
        const uint32_t cnt = list_cnt;
        uint32_t kept = 0;
        for (uint32_t r = 0; r < cnt; r += 64) {
          uint32_t pw[2][4], meta[2];
          bool live[2], q[2];
#pragma unroll
          for (int e = 0; e < 2; e++) {
            live[e] = r + 32u * e + c.lane < cnt;
            meta[e] = nx_meta[e];
#pragma unroll
            for (int j = 0; j < 4; j++) pw[e][j] = nx_pw[e][j];
          }
          // the entries of the next iteration are requested now: this iteration's appends stay below r + 64
          if (r + 64u < cnt) fetch_entries(r + 64u);
          if (cbf || L > 1u) {
            uint32_t v[2][4];
#pragma unroll
            for (int e = 0; e < 2; e++)
#pragma unroll
              for (int j = 0; j < 4; j++) v[e][j] = live[e] ? __ldcg(V + (pw[e][j] & 0xFFFFFFu)) : 0u;
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const uint32_t t = meta[e] & kTimeMask;
              bool reached = live[e];
              uint32_t mx = 0;
#pragma unroll
              for (int j = 0; j < 4; j++) {
                reached &= (v[e][j] & ~vmask) == tag;
                mx = max(mx, v[e][j] & vmask);
                if (cbf && live[e] && v[e][j] == (tag | t)) cbf[pw[e][j] & 0xFFFFFFu] = (uint8_t)L; // moved counter j to level L
              }
              q[e] = reached && t > mx;
            }
          } else { // level 1: most entries fail on their first counter (they are its first toucher)
            uint32_t v0[2];
#pragma unroll
            for (int e = 0; e < 2; e++) v0[e] = live[e] ? __ldcg(V + (pw[e][0] & 0xFFFFFFu)) : 0u;
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const uint32_t t = meta[e] & kTimeMask;
              q[e] = live[e] && (v0[e] & ~vmask) == tag && (v0[e] & vmask) < t;
            }
            uint32_t v[2][3];
#pragma unroll
            for (int e = 0; e < 2; e++)
#pragma unroll
              for (int j = 0; j < 3; j++) v[e][j] = q[e] ? __ldcg(V + (pw[e][j + 1] & 0xFFFFFFu)) : 0u;
#pragma unroll
            for (int e = 0; e < 2; e++) {
              const uint32_t t = meta[e] & kTimeMask;
#pragma unroll
              for (int j = 0; j < 3; j++) q[e] = q[e] && (v[e][j] & ~vmask) == tag && (v[e][j] & vmask) < t;
            }
          }
#pragma unroll
          for (int e = 0; e < 2; e++) {
            const uint32_t t = meta[e] & kTimeMask, thr = meta[e] >> kTimeBits;
            q[e] = q[e] && thr > L;
            if (q[e] && thr == L + 1u) {
              uint32_t ci, bi[4];
#pragma unroll
              for (int j = 0; j < 4; j++) unpack_index(pw[e][j], ci, bi[j]);
              bf_insert(bf, bi);
            }
            if (q[e] && L < lread) {
#pragma unroll
              for (int j = 0; j < 4; j++) red_min(Vn + (pw[e][j] & 0xFFFFFFu), tag_next | t);
            }
          }
          // every lane has both entries in registers before the ballots inside return; kept <= r
          surv_append(lst, kept, q[0], pw[0], meta[0], c.lane);
          surv_append(lst, kept, q[1], pw[1], meta[1], c.lane);
        }
        __syncwarp();
        if (c.lane == 0) warp_cnt[sl][wib] = kept;
      }

      __syncthreads(); // every thread of the CTA has issued its part of the round
      if (threadIdx.x == 0) {
        const unsigned long long t_c = globaltimer_ns();
        if (phase == PH_L0 && p.weighted) { // this CTA's round-0 rate, for the weighted shares
          cal_ns += t_c - t_b;
          cal_steps += ((uint64_t(n_steps) * cum_sh[kLevelWarps]) >> 32) - ((uint64_t(n_steps) * cum_sh[0]) >> 32);
          if (S.publish) {
            p.speed[blockIdx.x] = uint32_t(min(max(cal_steps * 4000000ull / max(cal_ns, 1ull), 1ull), 262143ull));
            cal_ns = 0; cal_steps = 0;
          }
        }
        __threadfence();
        atomicAdd(bar, 1ull);
        // where the time of a CTA goes, per kind of round: barrier wait, work, rounds (kept in shared memory: global
        // read-modify-writes here would make the reporting CTA late for every barrier)
        diag[phase * 3 + 0] += t_b - t_a;
        diag[phase * 3 + 1] += t_c - t_b;
        diag[phase * 3 + 2] += 1;
        S.target += gridDim.x;
        S.done_b1 = 0;
        // next round of this slot (same decision in every CTA)
        switch (phase) {
        case PH_CLEAR:
          S.epoch = 1; S.tag = (maxtag - 1u) << tb; S.phase = PH_L0;
          break;
        case PH_L0: // T_1 carries S.tag; the round that reads it writes T_2 under the next tag
          if (S.lread >= 1u) { S.phase = PH_READ; S.L = 1; S.epoch++; S.tag_next = (maxtag - S.epoch) << tb; }
          else { S.done_b1 = batch + 1u; S.done_ki = ki; begin_stream(S); }
          break;
        default: // PH_READ
          if (L < S.lread) { S.L = L + 1u; S.tag = S.tag_next; S.epoch++; S.tag_next = (maxtag - S.epoch) << tb; }
          else { S.done_b1 = batch + 1u; S.done_ki = ki; begin_stream(S); }
          break;
        }
      }
    }
    if (!any) break;
  }
  for (uint32_t sl = 0; sl < p.n_slots; sl++) { // last streams of the slots (states are final and uniform here)
    __syncthreads();
    const SlotState& S = slots[sl];
    if (!S.done_b1) continue;
    if (threadIdx.x == 0) {
      while (ld_relaxed_u64(p.bars + sl) < S.target) { }
      __threadfence();
    }
    __syncthreads();
    filter_final(p, S.done_b1 - 1u, S.done_ki, gtid, gthreads);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ops += __shfl_xor_sync(0xffffffffu, ops, o);
  if (c.lane == 0 && ops) atomicAdd(p.counters + 0, ops);
  if (c.lane == 0 && list_seen) atomicAdd(p.counters + 17, list_seen);
  if (blockIdx.x == 0 && threadIdx.x == 0) p.counters[21] = globaltimer_ns();
  if (blockIdx.x == p.report_cta && threadIdx.x == 0)
    for (int i = 0; i < 15; i++) p.counters[2 + i] += diag[i]; // several waves add up
}

int levels_max_grid(int sm_count, int ctas_per_sm)
{
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, build_filters_levels_kernel, kLevelWarps * 32, 0);
  if (per_sm < 1) per_sm = 1;
  if (per_sm > 4) per_sm = 4;
  if (ctas_per_sm > 0) per_sm = std::min(per_sm, ctas_per_sm);
  if (const char* e = std::getenv("GP_LEVEL_CTAS")) per_sm = std::max(1, std::min(per_sm, std::atoi(e))); // experiments
  return sm_count * per_sm;
}

int levels_max_slots() { return kMaxSlots; }

void preload_levels()
{
  cudaFuncAttributes a;
  if (cudaFuncGetAttributes(&a, (const void*)build_filters_levels_kernel) != cudaSuccess) cudaGetLastError();
  if (cudaFuncGetAttributes(&a, (const void*)fill_anchor_kernel) != cudaSuccess) cudaGetLastError();
}

cudaError_t launch_build_filters_levels(const LevelParams& p, int sm_count, cudaStream_t s, int ctas_per_sm)
{
  if (p.n_streams == 0) return cudaSuccess;
  // 132 KB of shared memory (3 CTAs x 33 KB), the rest L1; the edit kernel asks for the same so that the two can
  // share an SM.  Function attributes are per device: set on every launch (a cheap host-side call).
  cudaFuncSetAttribute((const void*)build_filters_levels_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 58);
  LevelParams lp = p;
  void* args[] = { &lp };
  // cooperative launch: the barriers need every CTA resident
  return cudaLaunchCooperativeKernel((const void*)build_filters_levels_kernel, dim3(levels_max_grid(sm_count, ctas_per_sm)),
                                     dim3(kLevelWarps * 32), args, 0, s);
}

} // namespace gp
