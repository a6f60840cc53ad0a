// goldpolish-mask (+ goldpolish-to-upper) on the device, run right after the edit kernel on the
// polished contigs while they are still resident.
//
// Semantics of scripts/goldpolish-mask:44-72 (`extend_gaps`), per record:
//   1. the first and the last k characters are upper-cased (the whole record if it is shorter than 2k);
//   2. the record is cut into the successive matches of ([ACTG]+|[Nn]+|[actgUNMRWSYKVHDBunmrwsykvhdb]+)
//      -- characters outside all three classes are dropped; a run of the third class swallows N/n that
//      follow it, a run that STARTS with N/n is a run of N/n only;
//   3. a run whose first character is not 'N' and that is shorter than k is lower-cased (-s) or replaced by
//      N's (-n); every other run is copied;
//   4. leading and trailing N/n are stripped; an empty result becomes "N".
// goldpolish-to-upper (scripts/goldpolish-to-upper:15-21) upper-cases the record.
//
// One warp per contig: the warp stages 1 KiB chunks through shared memory (coalesced both ways), lane 0 runs
// the run automaton over a chunk.  The records are a few kbp: the whole pass is a few ms for the longest contig
// and is not on the critical path of the filter build.
#include "gp_kernels.cuh"

#include <algorithm>

namespace gp {

namespace {

constexpr int kPrepWarps = 4;
constexpr uint32_t kPrepChunk = 1024;
constexpr uint32_t kPrepMaxK = 64;

enum : unsigned char { T_X = 0, T_A = 1, T_N = 2, T_L = 3 };

__device__ __forceinline__ char up(char c) { return (c >= 'a' && c <= 'z') ? char(c - 32) : c; }
__device__ __forceinline__ char low(char c) { return (c >= 'A' && c <= 'Z') ? char(c + 32) : c; }

__global__ void __launch_bounds__(kPrepWarps * 32) prep_kernel(PrepParams p)
{
  __shared__ unsigned char type_sh[256];
  __shared__ char in_sh[kPrepWarps][kPrepChunk];
  __shared__ char out_sh[kPrepWarps][kPrepChunk + kPrepMaxK];
  __shared__ char pend_sh[kPrepWarps][kPrepMaxK];
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
    const char c = char(i);
    unsigned char t = T_X;
    if (c == 'A' || c == 'C' || c == 'T' || c == 'G') t = T_A;
    else if (c == 'N' || c == 'n') t = T_N;
    else {
      const char u = up(c);
      const bool iupac = u == 'U' || u == 'M' || u == 'R' || u == 'W' || u == 'S' || u == 'Y' || u == 'K' || u == 'V' ||
                         u == 'H' || u == 'D' || u == 'B';
      if (c == 'a' || c == 'c' || c == 't' || c == 'g' || iupac) t = T_L;
    }
    type_sh[i] = t;
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, wib = threadIdx.x >> 5;
  const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, nw = (gridDim.x * blockDim.x) >> 5;
  char* in = in_sh[wib];
  char* out = out_sh[wib];
  char* pend = pend_sh[wib];
  const uint32_t k = p.k;
  for (uint32_t c = gw; c < p.n_contigs; c += nw) {
    if (p.dropped[c]) continue;
    const uint32_t w = p.which[c];
    const char* __restrict__ src = p.buf[w] + p.cap_off[c];
    char* __restrict__ dst = p.buf[w ^ 1u] + p.cap_off[c];
    const uint32_t len = p.cur_len[c];
    if (p.mode == 0) { // to-upper only
      for (uint32_t i = lane; i < len; i += 32) dst[i] = up(src[i]);
      if (lane == 0) p.which[c] = uint8_t(w ^ 1u);
      continue;
    }
    // automaton state (meaningful in lane 0)
    uint32_t run_type = T_X, pend_n = 0, o_total = 0, last_non_n = 0;
    bool run_keep = false, started = false;
    const bool all_upper = len < 2u * k;
    for (uint32_t base = 0; base < len; base += kPrepChunk) {
      const uint32_t n = min(kPrepChunk, len - base);
      for (uint32_t i = lane; i < n; i += 32) in[i] = src[base + i];
      __syncwarp();
      uint32_t o = 0;
      if (lane == 0) {
        auto emit = [&](char ch) {
          if (p.to_upper) ch = up(ch);
          const bool isn = ch == 'N' || ch == 'n';
          if (!started) { if (isn) return; started = true; }
          out[o++] = ch;
          if (!isn) last_non_n = o_total + o;
        };
        for (uint32_t i = 0; i < n; i++) {
          const uint32_t pos = base + i;
          char ch = in[i];
          if (all_upper || pos < k || pos >= len - k) ch = up(ch);
          const uint32_t t = type_sh[(unsigned char)ch];
          const bool cont = (run_type == T_A && t == T_A) || (run_type == T_N && t == T_N) ||
                            (run_type == T_L && (t == T_L || t == T_N));
          if (!cont) {
            if (!run_keep) // the run that ends here was shorter than k
              for (uint32_t j = 0; j < pend_n; j++) emit(p.mode == 1 ? low(pend[j]) : 'N');
            pend_n = 0;
            run_type = t;
            run_keep = t == T_N && ch == 'N';
          }
          if (t == T_X) { run_type = T_X; continue; } // outside every class: dropped
          if (run_keep) emit(ch);
          else {
            pend[pend_n++] = ch;
            if (pend_n == k) { // the run reached k: copied as it is
              for (uint32_t j = 0; j < pend_n; j++) emit(pend[j]);
              pend_n = 0;
              run_keep = true;
            }
          }
        }
        if (base + n == len && !run_keep) { // last run of the record
          for (uint32_t j = 0; j < pend_n; j++) emit(p.mode == 1 ? low(pend[j]) : 'N');
          pend_n = 0;
        }
      }
      o = __shfl_sync(0xffffffffu, o, 0);
      const uint32_t ot = __shfl_sync(0xffffffffu, o_total, 0);
      __syncwarp();
      for (uint32_t i = lane; i < o; i += 32) dst[ot + i] = out[i];
      if (lane == 0) o_total += o;
      __syncwarp();
    }
    if (lane == 0) {
      uint32_t out_len = last_non_n; // trailing N/n stripped
      if (out_len == 0) { dst[0] = 'N'; out_len = 1; }
      p.cur_len[c] = out_len;
      p.which[c] = uint8_t(w ^ 1u);
    }
  }
}

} // namespace

cudaError_t launch_prep(const PrepParams& p, int sm_count, cudaStream_t s)
{
  if (p.n_contigs == 0) return cudaSuccess;
  if (p.mode != 0 && (p.k == 0 || p.k > kPrepMaxK)) return cudaErrorInvalidValue;
  uint32_t grid = (p.n_contigs + kPrepWarps - 1) / kPrepWarps;
  grid = std::min<uint32_t>(grid, uint32_t(sm_count) * 8u);
  prep_kernel<<<grid, kPrepWarps * 32, 0, s>>>(p);
  return cudaGetLastError();
}

} // namespace gp
