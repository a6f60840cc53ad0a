// Read packing and the counting-Bloom-gated Bloom filter build (sm_100a).
//
// Replaces fill_bfs (bcgsc/goldpolish src/utils.cpp:96-123) and the per-batch loop around it
// (src/goldpolish_targeted_bfs.cpp:124-140).  One warp owns one (batch, k) stream and walks it
// in the reference's order -- reads as listed, positions ascending -- 32 k-mers per step.
// Inside a step the warp keeps the reference's sequential semantics exactly:
//   * lanes whose four counters are touched by no lower lane of the step ("independent")
//     read-modify-write in parallel: their result cannot depend on any other lane;
//   * the remaining lanes (same k-mer twice in 32 positions, or a chance counter collision)
//     are replayed one at a time in lane order against the updated counters.
// Bloom-filter bit sets commute, so they are plain 32-bit atomicOr.
#include "gp_common.cuh"
#include "gp_kernels.cuh"

namespace gp {

// ------------------------------------------------------------------------------------
// pack: ASCII -> 2 bits/base (32 bases per u64) + 1 bit/base "no seed" mask (32 per u32)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) pack_reads_kernel(const char* __restrict__ ascii,
                                                         const uint64_t* __restrict__ ascii_off,
                                                         const uint64_t* __restrict__ base_off, // multiple of 32
                                                         uint64_t* __restrict__ pk, uint32_t* __restrict__ nm,
                                                         uint32_t n_reads)
{
  for (uint32_t r = blockIdx.x; r < n_reads; r += gridDim.x) {
    const uint64_t a0 = ascii_off[r];
    const uint32_t len = uint32_t(ascii_off[r + 1] - a0);
    const uint64_t g0 = base_off[r] >> 5;
    const uint32_t groups = (len + 31u) >> 5;
    const unsigned char* src = reinterpret_cast<const unsigned char*>(ascii) + a0;
    for (uint32_t gi = threadIdx.x; gi < groups; gi += blockDim.x) {
      uint64_t w = 0;
      uint32_t m = 0;
      const uint32_t base = gi << 5;
#pragma unroll 8
      for (uint32_t i = 0; i < 32; i++) {
        const uint32_t q = base + i;
        uint32_t c = q < len ? src[q] : 0u;
        const uint32_t lc = c | 0x20u;
        const bool ok = (lc == 'a') | (lc == 'c') | (lc == 'g') | (lc == 't');
        uint32_t x = (lc >> 1) & 3u; // a:0 c:1 g:3 t:2
        x ^= x >> 1;                 // a:0 c:1 g:2 t:3
        w |= uint64_t(ok ? x : 0u) << (2 * i);
        m |= (ok ? 0u : 1u) << i;
      }
      pk[g0 + gi] = w;
      nm[g0 + gi] = m;
    }
  }
}

void launch_pack_reads(const char* ascii, const uint64_t* ascii_off, const uint64_t* base_off, uint64_t* pk,
                       uint32_t* nm, uint32_t n_reads, cudaStream_t s)
{
  if (n_reads == 0) return;
  const uint32_t grid = n_reads < 148u * 16u ? n_reads : 148u * 16u;
  pack_reads_kernel<<<grid, 128, 0, s>>>(ascii, ascii_off, base_off, pk, nm, n_reads);
}

// ------------------------------------------------------------------------------------
// build
// ------------------------------------------------------------------------------------
constexpr int kBuildWarps = 8;
constexpr int kTabBits = 9;
constexpr int kTabSlots = 1 << kTabBits;

// byte tables for hashing 4 packed bases at a time; k-independent because groups are counted
// from the k-mer end for the forward strand and from its start for the reverse strand.
//   tf[g][b] = XOR_j srol^{4g+3-j}(seed[code_j(b)])      (bases 4 apart from the end)
//   tr[m][b] = XOR_j srol^{4m+j}(seed[3-code_j(b)])      (complement strand)
__device__ __forceinline__ void fill_hash_tables(uint64_t* tf, uint64_t* tr)
{
  for (uint32_t e = threadIdx.x; e < 8u * 256u; e += blockDim.x) {
    const uint32_t g = e >> 8, b = e & 255u;
    uint64_t f = 0, r = 0;
#pragma unroll
    for (uint32_t j = 0; j < 4; j++) {
      const uint32_t code = (b >> (2 * j)) & 3u;
      f ^= srol(seed_of_code(code), 4 * g + 3 - j);
      r ^= srol(seed_of_code(3u - code), 4 * g + j);
    }
    tf[e] = f;
    tr[e] = r;
  }
}

__device__ __forceinline__ void cbf_bf_apply(uint8_t* __restrict__ cbf, uint32_t* __restrict__ bf,
                                             const uint32_t (&ci)[4], const uint32_t (&bi)[4],
                                             const uint32_t (&c)[4], uint32_t thr)
{
  const uint32_t mn = min(min(c[0], c[1]), min(c[2], c[3]));
  uint32_t cnt = mn;
  if (mn < thr) {
#pragma unroll
    for (int j = 0; j < 4; j++)
      if (c[j] == mn) __stcg(cbf + ci[j], (uint8_t)(mn + 1));
    cnt = mn + 1;
  }
  if (cnt >= thr) {
#pragma unroll
    for (int j = 0; j < 4; j++) atomicOr(bf + (bi[j] >> 5), 1u << (bi[j] & 31u));
  }
}

__global__ void __launch_bounds__(kBuildWarps * 32) build_filters_kernel(BuildParams p)
{
  __shared__ uint64_t tf[8 * 256];
  __shared__ uint64_t tr[8 * 256];
  __shared__ uint32_t tabs[kBuildWarps][kTabSlots];
  fill_hash_tables(tf, tr);
  for (uint32_t i = threadIdx.x; i < kBuildWarps * kTabSlots; i += blockDim.x) (&tabs[0][0])[i] = 0u;
  __syncthreads();

  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t warp = threadIdx.x >> 5;
  uint32_t* tab = tabs[warp];
  unsigned long long ops = 0, serial = 0;

  for (;;) {
    uint32_t slot = 0;
    if (lane == 0) slot = atomicAdd(p.next_stream, 1u);
    slot = __shfl_sync(0xffffffffu, slot, 0);
    if (slot >= p.n_streams) break;
    const uint32_t sid = p.stream_order[slot];     // local stream id within this wave
    const uint32_t lb = sid / p.nk, ki = sid - lb * p.nk;
    const uint32_t batch = p.first_batch + lb;
    const uint32_t k = p.k[ki];
    const uint32_t kq = k >> 2; // bytes per k-mer
    const uint32_t kmask = k >= 32 ? 0xFFFFFFFFu : ((1u << k) - 1u);
    const uint64_t mul1 = 1ull ^ (uint64_t(k) * kMultiSeed);
    const uint64_t mul2 = 2ull ^ (uint64_t(k) * kMultiSeed);
    const uint64_t mul3 = 3ull ^ (uint64_t(k) * kMultiSeed);
    uint8_t* __restrict__ cbf = p.cbf_pool + uint64_t(sid) * kCbfCounters;
    uint32_t* __restrict__ bf = p.bf_pool + (uint64_t(batch) * p.nk + ki) * kBfWords;

    const uint64_t e0 = p.batch_entry_off[batch], e1 = p.batch_entry_off[batch + 1];
    for (uint64_t e = e0; e < e1; e++) {
      const gp_read_entry ent = p.entries[e];
      const uint32_t thr = ent.kmer_threshold - 2u + ki; // utils.cpp:108,121
      const uint32_t len = p.read_len[ent.read_id];
      if (len < k) continue;
      const uint64_t boff = p.read_boff[ent.read_id];
      const uint32_t npos = len - k + 1;
      // One step = 32 consecutive positions.  Steps are software-pipelined: while step i waits
      // for its counters, step i+1 is already hashed and its counter sectors are prefetched
      // into L2, so that a lone stream (the tail of a launch) pays L2 latency per step instead
      // of HBM latency.  Loads of step i+1 are only ISSUED after step i has stored.
      auto prepare = [&](uint32_t p0, uint32_t (&ci)[4], uint32_t (&bi)[4]) -> bool {
        const uint32_t q = p0 + lane;
        const uint64_t g = boff + q;
        const uint64_t wi = g >> 5;
        const uint32_t sh = uint32_t(g & 31u);
        // 1 bit/base validity window
        const uint32_t m0 = __ldg(p.nm + wi), m1 = __ldg(p.nm + wi + 1);
        const uint32_t mw = __funnelshift_r(m0, m1, sh);
        const bool valid = (q < npos) && ((mw & kmask) == 0u);
        // 2 bit/base window
        const uint64_t w0 = __ldg(p.pk + wi), w1 = __ldg(p.pk + wi + 1);
        const uint64_t w = sh ? ((w0 >> (2 * sh)) | (w1 << (64 - 2 * sh))) : w0;
        ci[0] = 0xFFFFFFF0u; ci[1] = 0xFFFFFFF1u; ci[2] = 0xFFFFFFF2u; ci[3] = 0xFFFFFFF3u;
        if (valid) {
          uint64_t fh = 0, rh = 0;
          for (uint32_t m = 0; m < kq; m++) {
            const uint32_t b = uint32_t(w >> (8 * m)) & 255u;
            fh ^= tf[((kq - 1 - m) << 8) | b];
            rh ^= tr[(m << 8) | b];
          }
          const uint64_t h0 = fh + rh;
          uint64_t h1 = h0 * mul1, h2 = h0 * mul2, h3 = h0 * mul3;
          h1 ^= h1 >> kMultiShift; h2 ^= h2 >> kMultiShift; h3 ^= h3 >> kMultiShift;
          ci[0] = cbf_index(h0); ci[1] = cbf_index(h1); ci[2] = cbf_index(h2); ci[3] = cbf_index(h3);
          bi[0] = bf_index(h0); bi[1] = bf_index(h1); bi[2] = bf_index(h2); bi[3] = bf_index(h3);
        }
        return valid;
      };
      uint32_t ci[4], bi[4], c[4], nci[4], nbi[4];
      bool valid = prepare(0, ci, bi), nvalid = false;
      if (valid) {
#pragma unroll
        for (int j = 0; j < 4; j++) c[j] = __ldcg(cbf + ci[j]);
      }
      for (uint32_t p0 = 0; p0 < npos; p0 += 32) {
        const bool has_next = p0 + 32 < npos;
        if (has_next) {
          nvalid = prepare(p0 + 32, nci, nbi);
          if (nvalid) {
#pragma unroll
            for (int j = 0; j < 4; j++) asm volatile("prefetch.global.L2 [%0];" ::"l"(cbf + nci[j]));
          }
        }

        // Which lanes share a counter with a LOWER lane of this step?  Lanes publish themselves
        // in a small direct-mapped table of lane masks (slot = hash of the counter index);
        // slot sharing only nominates candidates, the counter indices themselves are then
        // compared through shuffles, so the answer is exact whatever the table size.
        uint32_t sl[4];
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; j++) {
            sl[j] = (ci[j] * 2654435761u) >> (32 - kTabBits);
            atomicOr(&tab[sl[j]], 1u << lane);
          }
        }
        __syncwarp();
        uint32_t lower = 0;
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; j++) lower |= tab[sl[j]];
          lower &= (1u << lane) - 1u;
        }
        __syncwarp();
        if (valid) {
#pragma unroll
          for (int j = 0; j < 4; j++) tab[sl[j]] = 0u; // leave the table clean for the next step
        }
        bool indep = true;
        const uint32_t rounds = __reduce_max_sync(0xffffffffu, (uint32_t)__popc(lower));
        for (uint32_t it = 0; it < rounds; it++) {
          const uint32_t l = lower ? (uint32_t)__ffs(lower) - 1u : lane;
          lower &= lower - 1u;
          const uint32_t o0 = __shfl_sync(0xffffffffu, ci[0], l), o1 = __shfl_sync(0xffffffffu, ci[1], l);
          const uint32_t o2 = __shfl_sync(0xffffffffu, ci[2], l), o3 = __shfl_sync(0xffffffffu, ci[3], l);
          if (l != lane) {
#pragma unroll
            for (int j = 0; j < 4; j++) indep &= (ci[j] != o0) & (ci[j] != o1) & (ci[j] != o2) & (ci[j] != o3);
          }
        }
        if (valid) {
          if (indep) cbf_bf_apply(cbf, bf, ci, bi, c, thr);
          ops++;
        }
        uint32_t dep = __ballot_sync(0xffffffffu, valid && !indep);
        while (dep) {
          __syncwarp();
          const uint32_t l = __ffs(dep) - 1;
          if (lane == l) {
#pragma unroll
            for (int j = 0; j < 4; j++) c[j] = __ldcg(cbf + ci[j]);
            cbf_bf_apply(cbf, bf, ci, bi, c, thr);
            serial++;
          }
          dep &= dep - 1;
        }
        __syncwarp(); // orders this step's counter stores before the next step's loads
        if (has_next) {
          valid = nvalid;
#pragma unroll
          for (int j = 0; j < 4; j++) { ci[j] = nci[j]; bi[j] = nbi[j]; }
          if (valid) {
#pragma unroll
            for (int j = 0; j < 4; j++) c[j] = __ldcg(cbf + ci[j]);
          }
        }
      }
    }
  }
  // per-warp totals
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ops += __shfl_xor_sync(0xffffffffu, ops, o);
    serial += __shfl_xor_sync(0xffffffffu, serial, o);
  }
  if (lane == 0) {
    if (ops) atomicAdd(p.counters + 0, ops);
    if (serial) atomicAdd(p.counters + 1, serial);
  }
}

void launch_build_filters(const BuildParams& p, int sm_count, cudaStream_t s)
{
  if (p.n_streams == 0) return;
  // persistent grid: as many 8-warp CTAs as can be resident, never more warps than streams
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, build_filters_kernel, kBuildWarps * 32, 0);
  if (per_sm < 1) per_sm = 1;
  uint32_t grid = uint32_t(sm_count) * uint32_t(per_sm);
  const uint32_t need = (p.n_streams + kBuildWarps - 1) / kBuildWarps;
  if (grid > need) grid = need;
  build_filters_kernel<<<grid, kBuildWarps * 32, 0, s>>>(p);
}

// ------------------------------------------------------------------------------------
// random-access roof: the build kernel's memory shape with the hashing taken out
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) roof_kernel(uint8_t* cbf_pool, uint32_t* bf_pool, uint64_t region,
                                                   uint32_t iters, uint32_t total_warps)
{
  const uint32_t lane = threadIdx.x & 31u;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (gw >= total_warps) return;
  uint8_t* cbf = cbf_pool + uint64_t(gw) * region;
  uint32_t* bf = bf_pool + uint64_t(gw) * kBfWords;
  uint64_t x = 0x9e3779b97f4a7c15ULL * (uint64_t(gw) * 32 + lane + 1);
  for (uint32_t it = 0; it < iters; it++) {
    uint32_t ci[4], bi[4], c[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
      x ^= x << 13; x ^= x >> 7; x ^= x << 17;
      ci[j] = uint32_t((x >> 11) % region);
      bi[j] = uint32_t(x) & uint32_t(kBfBits - 1);
    }
#pragma unroll
    for (int j = 0; j < 4; j++) c[j] = __ldcg(cbf + ci[j]);
    const uint32_t mn = min(min(c[0], c[1]), min(c[2], c[3]));
    // alternate between the two halves of the algorithmic traffic: counter write-back below
    // the threshold, filter bit sets at it
    if ((it & 1u) == 0u) {
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (c[j] == mn) __stcg(cbf + ci[j], (uint8_t)((mn + 1) & 15u));
    } else {
#pragma unroll
      for (int j = 0; j < 4; j++) atomicOr(bf + (bi[j] >> 5), 1u << (bi[j] & 31u));
    }
    __syncwarp();
  }
}

void launch_roof(uint8_t* cbf_pool, uint32_t* bf_pool, uint64_t region, uint32_t iters, uint32_t warps, cudaStream_t s)
{
  const uint32_t threads = 256;
  const uint32_t grid = (warps * 32 + threads - 1) / threads;
  roof_kernel<<<grid, threads, 0, s>>>(cbf_pool, bf_pool, region, iters, warps);
}

} // namespace gp
