// ntEdit scan / edit loop on the GPU (sm_100a): one warp per contig, all k rounds chained.
//
// Replaces kmerizeAndCorrect and its helpers (bcgsc/goldpolish subprojects/ntedit/ntedit.cpp
// :1414-1771, :392-413, :441-451, :480-777, :939-1411) plus the FASTA body of writeEditsToFile
// (:797-935) and the k-chain of scripts/goldpolish-ntedit:20-29.
//
// The contig is a sequential state machine (every edit changes what the next position sees),
// so parallelism lives *inside* a decision:
//   * scan: 32 consecutive k-mer windows are hashed and looked up at once; the first absent
//     window / non-IUPAC base / end of contig is found with a ballot;
//   * look-ahead and substitution checks: lane j holds window j, the every-3rd sample set is
//     a lane mask, the counts are ballot + popc;
//   * insertion / deletion search: the 341 insertion strings and <=10 deletions are spread
//     over the lanes, each lane rolling its own candidate; the winner (support, then the
//     reference's evaluation order for ties) is a warp max-reduction.
// The reference's rope (vector<seqNode>) is kept literally in global memory and mutated by
// lane 0 with the reference's own slot arithmetic, because its low-complexity rollback
// (:1038-1100) leaves holes whose position decides what the writer emits.  Hashing never walks
// the rope: the k window characters live in a shared-memory ring and the characters ahead of
// the tail cursor in a shared-memory look-ahead buffer refilled from the rope.
#include "gp_common.cuh"
#include "gp_kernels.cuh"

namespace gp {

constexpr int kEditWarps = 4;
constexpr uint32_t kABuf = 128; // look-ahead ring (power of two)
constexpr uint32_t kRing = 64;  // window ring (power of two, >= 32)
constexpr uint32_t kFull = 0xffffffffu;

struct Cur {
  uint32_t pos, idx; // (seq_i, node_index) of the reference
  EdNode n;          // cached nodes[idx]; type -2 when idx is past the vector
};

struct WS {
  char* seq;
  uint32_t len;
  EdNode* nd;
  uint32_t nn, ncap;
  const uint32_t* bf;
  uint32_t k, jump, icap;
  uint64_t mul1, mul2, mul3;
  float thrM, thrE, thrD;
  uint32_t max_ins, max_del;
  int mode, mask;
  Cur h, t, m;
  HashState hs;
  uint32_t rh;     // ring head
  uint32_t a0, an; // look-ahead window [a0, a0+an)
  bool exhausted;  // refilling hit the end of the stream
  unsigned char* ring;
  unsigned char* abuf;
  int err;
  uint32_t lane;
  uint32_t n_trig, n_edit, n_mask, n_roll;
};

__device__ __forceinline__ bool is_accepted(uint32_t c)
{ // isAcceptedBase(toupper(c)), ntedit.cpp:363-367
  c &= ~0x20u; // toupper for letters; non-letters never match below either way
  return c == 'A' || c == 'T' || c == 'G' || c == 'C' || c == 'R' || c == 'Y' || c == 'S' || c == 'W' ||
         c == 'K' || c == 'M' || c == 'B' || c == 'D' || c == 'H' || c == 'V';
}
__device__ __forceinline__ uint32_t to_upper(uint32_t c) { return (c >= 'a' && c <= 'z') ? c - 32u : c; }
__device__ __forceinline__ uint32_t to_lower(uint32_t c) { return (c >= 'A' && c <= 'Z') ? c + 32u : c; }
__device__ __forceinline__ unsigned char rc_char(uint32_t c)
{ // RC, ntedit.cpp:369-388
  switch (c) {
  case 'A': case 'a': return 'T';
  case 'T': case 't': return 'A';
  case 'G': case 'g': return 'C';
  case 'C': case 'c': return 'G';
  default: return 'N';
  }
}

__device__ __forceinline__ EdNode ld_node(const EdNode* p)
{
  const int4 v = *reinterpret_cast<const int4*>(p);
  EdNode n;
  n.type = v.x; n.s = uint32_t(v.y); n.e = uint32_t(v.z); n.c = uint32_t(v.w);
  return n;
}
__device__ __forceinline__ void st_node(EdNode* p, const EdNode& n)
{
  *reinterpret_cast<int4*>(p) = make_int4(n.type, int(n.s), int(n.e), int(n.c));
}

__device__ __forceinline__ bool bf_contains(const WS& w, const HashState& h)
{ // btllib KmerBloomFilter::contains with the four ntHash values (ntedit.cpp:1470)
  const uint64_t b = h.fh + h.rh;
  uint64_t h1 = b * w.mul1, h2 = b * w.mul2, h3 = b * w.mul3;
  h1 ^= h1 >> kMultiShift; h2 ^= h2 >> kMultiShift; h3 ^= h3 >> kMultiShift;
  const uint32_t n0 = bf_index(b), n1 = bf_index(h1), n2 = bf_index(h2), n3 = bf_index(h3);
  const uint32_t w0 = __ldg(w.bf + (n0 >> 5)), w1 = __ldg(w.bf + (n1 >> 5));
  const uint32_t w2 = __ldg(w.bf + (n2 >> 5)), w3 = __ldg(w.bf + (n3 >> 5));
  return ((w0 >> (n0 & 31u)) & (w1 >> (n1 & 31u)) & (w2 >> (n2 & 31u)) & (w3 >> (n3 & 31u)) & 1u) != 0u;
}

// ---- rope cursors (getCharacter :667-678, increment :681-699) -------------------------
__device__ __forceinline__ void cur_load(const WS& w, Cur& c)
{
  if (c.idx < w.nn) c.n = ld_node(w.nd + c.idx);
  else c.n.type = -2;
}
__device__ __forceinline__ uint32_t cur_char(const WS& w, const Cur& c)
{
  if (c.n.type == 0) return c.pos < w.len ? (unsigned char)w.seq[c.pos] : 0u;
  if (c.n.type == 1) return c.n.c;
  return 0u;
}
__device__ __forceinline__ void cur_increment(const WS& w, Cur& c)
{
  if (c.n.type == 0) {
    c.pos++;
    if (c.pos > c.n.e) {
      c.idx++;
      cur_load(w, c);
      if (c.n.type == 0) c.pos = c.n.s;
    }
  } else if (c.n.type == 1) {
    c.idx++;
    cur_load(w, c);
    if (c.n.type == 0) c.pos = c.n.s;
  }
}
__device__ __forceinline__ bool cur_dead(const WS& w, const Cur& c) { return c.pos >= w.len || c.idx >= w.nn; }

__device__ __forceinline__ uint32_t ring_at(const WS& w, uint32_t j) { return w.ring[(w.rh + j) & (kRing - 1)]; }
__device__ __forceinline__ uint32_t ahead_at(const WS& w, uint32_t j) { return w.abuf[(w.a0 + j) & (kABuf - 1)]; }

// Refill the look-ahead buffer by walking the rope from the materialisation cursor exactly as
// successive roll() calls would move the tail cursor (:958-966).
__device__ void ahead_fill(WS& w, uint32_t want)
{
  while (w.an < want && !w.exhausted) {
    if (cur_dead(w, w.m)) { w.exhausted = true; break; }
    if (w.m.n.type == 0 && w.m.pos < w.m.n.e && w.m.pos + 1 < w.len) {
      uint32_t run = min(min(w.m.n.e, w.len - 1) - w.m.pos, min(want - w.an, 32u));
      if (w.lane < run) w.abuf[(w.a0 + w.an + w.lane) & (kABuf - 1)] = (unsigned char)w.seq[w.m.pos + 1 + w.lane];
      w.m.pos += run;
      w.an += run;
      continue;
    }
    cur_increment(w, w.m);
    if (cur_dead(w, w.m)) { w.exhausted = true; break; }
    if (w.lane == 0) w.abuf[(w.a0 + w.an) & (kABuf - 1)] = (unsigned char)cur_char(w, w.m);
    w.an++;
  }
  __syncwarp();
}
__device__ __forceinline__ void ahead_reset(WS& w)
{
  w.a0 = 0; w.an = 0; w.exhausted = false;
  w.m = w.t;
}

// One main-loop roll (:939-969 + NTMC64 :304-314).  Returns false when the reference's roll()
// would; on success `in` is the incoming character.
__device__ bool roll_main(WS& w, uint32_t& in)
{
  if (cur_dead(w, w.h)) return false;
  if (w.an == 0) ahead_fill(w, 64);
  const uint32_t out = ring_at(w, 0);
  cur_increment(w, w.h);
  if (cur_dead(w, w.t)) return false;
  if (w.an == 0) return false; // tail cursor cannot advance
  in = ahead_at(w, 0);
  cur_increment(w, w.t);
  hs_roll(w.hs, w.k, out, in);
  __syncwarp();
  if (w.lane == 0) w.ring[(w.rh + w.k) & (kRing - 1)] = (unsigned char)in;
  w.rh = (w.rh + 1) & (kRing - 1);
  w.a0 = (w.a0 + 1) & (kABuf - 1);
  w.an--;
  __syncwarp();
  return true;
}

// ---- lane-0 rope surgery ---------------------------------------------------------------
struct Rope { // lane-0 view
  EdNode* nd;
  uint32_t nn, ncap;
  char* seq;
  uint32_t len;
  int err;
};
__device__ void rope_set(Rope& r, uint32_t i, const EdNode& v)
{ // "if (i < size) a[i] = v; else push_back(v)"
  if (i < r.nn) { st_node(r.nd + i, v); return; }
  if (i != r.nn || r.nn >= r.ncap) { r.err = 1; return; }
  st_node(r.nd + r.nn, v);
  r.nn++;
}
__device__ uint32_t rope_char(const Rope& r, uint32_t pos, uint32_t idx)
{
  if (idx >= r.nn) return 0;
  const EdNode n = ld_node(r.nd + idx);
  if (n.type == 0) return pos < r.len ? (unsigned char)r.seq[pos] : 0u;
  if (n.type == 1) return n.c;
  return 0;
}
__device__ void rope_increment(const Rope& r, uint32_t& pos, uint32_t& idx)
{
  if (idx >= r.nn) return;
  const EdNode n = ld_node(r.nd + idx);
  if (n.type == 0) {
    pos++;
    if (pos > n.e) {
      idx++;
      if (idx < r.nn) { const EdNode x = ld_node(r.nd + idx); if (x.type == 0) pos = x.s; }
    }
  } else if (n.type == 1) {
    idx++;
    if (idx < r.nn) { const EdNode x = ld_node(r.nd + idx); if (x.type == 0) pos = x.s; }
  }
}

// makeInsertion, :480-569
__device__ void rope_insert(Rope& r, uint32_t& t_idx, uint32_t insert_pos, const unsigned char* ins, uint32_t L)
{
  const EdNode orig = ld_node(r.nd + t_idx);
  if (orig.type == 0 && !(int(insert_pos) <= int(orig.s))) {
    EdNode head = orig;
    head.e = insert_pos - 1;
    st_node(r.nd + t_idx, head);
    for (uint32_t i = 0; i < L; i++) { EdNode c = { 1, 0, 0, ins[i] }; rope_set(r, t_idx + i + 1, c); }
    EdNode after = { 0, insert_pos, orig.e, 0 };
    rope_set(r, t_idx + L + 1, after);
    t_idx++;
    return;
  }
  if (orig.type == 0 || orig.type == 1) {
    // lift nodes [t_idx, t_idx+nre) up by L, then write the inserted characters
    uint32_t nre = 0;
    while (t_idx + nre < r.nn && ld_node(r.nd + t_idx + nre).type != -1) nre++;
    const uint32_t new_n = max(r.nn, t_idx + L + nre);
    if (new_n > r.ncap) { r.err = 1; return; }
    for (uint32_t q = nre; q > 0; q--) st_node(r.nd + t_idx + L + q - 1, ld_node(r.nd + t_idx + q - 1));
    r.nn = new_n;
    for (uint32_t q = 0; q < L; q++) { EdNode c = { 1, 0, 0, ins[q] }; st_node(r.nd + t_idx + q, c); }
  }
}

// makeDeletion, :574-664 (tail recursion unrolled)
__device__ void rope_delete(Rope& r, uint32_t& t_idx, uint32_t& pos, uint32_t num_del)
{
  for (;;) {
    if (t_idx >= r.nn) { r.err = 1; return; }
    const EdNode orig = ld_node(r.nd + t_idx);
    uint32_t leftover = 0;
    if (orig.type == 0) {
      if (pos <= orig.s) {
        if (pos + num_del <= orig.e) {
          EdNode n = orig; n.s = pos + num_del; st_node(r.nd + t_idx, n);
          pos = n.s;
          return;
        }
        leftover = pos + num_del - orig.e; // sic, :594
        pos = orig.e + 1;
        uint32_t i = t_idx + 1;
        while (i < r.nn) {
          EdNode x = ld_node(r.nd + i);
          if (x.type == -1) break;
          st_node(r.nd + i - 1, x);
          x.type = -1; st_node(r.nd + i, x);
          i++;
        }
      } else {
        if (pos + num_del <= orig.e) {
          EdNode split = { 0, pos + num_del, orig.e, 0 };
          EdNode n = orig; n.e = pos - 1; st_node(r.nd + t_idx, n);
          pos = split.s;
          t_idx++;
          rope_set(r, t_idx, split);
          return;
        }
        leftover = pos + num_del - orig.e; // sic, :622
        EdNode n = orig; n.e = pos - 1; st_node(r.nd + t_idx, n);
        pos = orig.e + 1;
        t_idx++;
      }
    } else if (orig.type == 1) {
      uint32_t i = t_idx;
      leftover = num_del;
      while (i < r.nn && leftover > 0) {
        EdNode x = ld_node(r.nd + i);
        if (x.type != 1) break;
        x.type = -1; st_node(r.nd + i, x);
        leftover--; i++;
      }
      uint32_t j = t_idx;
      while (i < r.nn) {
        EdNode x = ld_node(r.nd + i);
        if (x.type == -1) break;
        st_node(r.nd + j, x);
        x.type = -1; st_node(r.nd + i, x);
        i++; j++;
      }
    } else return;
    if (leftover == 0) return;
    if (!(t_idx < r.nn)) return;
    const EdNode nx = ld_node(r.nd + t_idx);
    if (nx.type == -1) return;
    if (nx.type == 0) pos = nx.s;
    num_del = leftover;
  }
}

// getPrevInsertion, :762-777
__device__ int rope_prev_insertion(const Rope& r, uint32_t t_seq, uint32_t t_idx, unsigned char* buf, int cap)
{
  int n = 0;
  uint32_t idx = t_idx;
  if (idx < r.nn) {
    const EdNode tn = ld_node(r.nd + idx);
    if ((tn.type == 0 && t_seq == tn.s) || tn.type == 1) idx--;
  }
  while (idx < r.nn) {
    const EdNode x = ld_node(r.nd + idx);
    if (x.type != 1) break;
    if (n < cap) buf[n] = rc_char(x.c);
    n++;
    idx--;
  }
  return n;
}

// isRepeatInsertion, :416-451
__device__ bool is_repeat(const unsigned char* s, int n)
{
  if (n <= 0 || n > 160) return false;
  int lps[160];
  lps[0] = 0;
  int len = 0, i = 1;
  while (i < n) {
    if (s[i] == s[len]) { len++; lps[i] = len; i++; }
    else if (len != 0) len = lps[len - 1];
    else { lps[i] = 0; i++; }
  }
  const int l = lps[n - 1];
  return l > 0 && n % (n - l) == 0;
}

// node shuffle of :1043-1056 / :1074-1088
__device__ void rope_remove_prev(Rope& r, uint32_t t_seq, uint32_t t_idx, uint32_t count)
{
  uint32_t j = 1;
  if (t_idx < r.nn) { const EdNode tn = ld_node(r.nd + t_idx); if (tn.type == 0 && t_seq == tn.s) j = 0; }
  for (uint32_t i = count; i > 0; i--) {
    if (i > t_idx) { r.err = 1; continue; }
    bool moved = false;
    if (t_idx + j < r.nn) {
      EdNode src = ld_node(r.nd + t_idx + j);
      if (src.type != -1) {
        st_node(r.nd + t_idx - i, src);
        src.type = -1; st_node(r.nd + t_idx + j, src);
        j++;
        moved = true;
      }
    }
    if (!moved) { EdNode x = ld_node(r.nd + t_idx - i); x.type = -1; st_node(r.nd + t_idx - i, x); }
  }
}

// findAcceptedKmer, :703-758.  kmer must hold 32 bytes.
__device__ bool rope_find_kmer(const Rope& r, uint32_t k, uint32_t& h_seq, uint32_t& t_seq, uint32_t& h_idx,
                               uint32_t& t_idx, unsigned char* kmer)
{
  uint32_t tt = t_idx, th = 0, i = t_seq, curr = t_idx;
  while (i < r.len && tt < r.nn && ld_node(r.nd + tt).type != -1) {
    uint32_t ch = rope_char(r, i, curr);
    if (is_accepted(ch)) {
      uint32_t kl = 0;
      kmer[kl++] = (unsigned char)ch;
      th = tt;
      uint32_t j = i;
      rope_increment(r, j, tt);
      while (j < r.len && tt < r.nn && ld_node(r.nd + tt).type != -1) {
        curr = tt;
        ch = rope_char(r, j, curr);
        if (!is_accepted(ch)) { i = j; break; }
        kmer[kl++] = (unsigned char)ch;
        if (kl == k) break;
        rope_increment(r, j, tt);
      }
      if (kl == k) { h_seq = i; t_seq = j; h_idx = th; t_idx = tt; return true; }
    }
    if (tt < r.nn) rope_increment(r, i, tt); else break;
  }
  h_seq = r.len; t_seq = r.len;
  return false;
}

// i-th string of multi_possible_bases[first] (:198-343): length-major, A<C<G<T
__device__ __forceinline__ uint32_t insertion_string(uint32_t first, uint32_t i, unsigned char* out)
{
  const uint32_t L = i < 1 ? 1 : i < 5 ? 2 : i < 21 ? 3 : i < 85 ? 4 : 5;
  const uint32_t start = L == 1 ? 0 : L == 2 ? 1 : L == 3 ? 5 : L == 4 ? 21 : 85;
  uint32_t r = i - start;
  out[0] = (unsigned char)first;
  for (uint32_t p = L - 1; p >= 1; p--) {
    const uint32_t d = r & 3u;
    out[p] = d == 0 ? 'A' : d == 1 ? 'C' : d == 2 ? 'G' : 'T';
    r >>= 2;
  }
  return L;
}
__device__ __forceinline__ uint32_t num_tries(uint32_t max_ins)
{ // :150
  return max_ins == 0 ? 0 : max_ins == 1 ? 1 : max_ins == 2 ? 5 : max_ins == 3 ? 21 : max_ins == 4 ? 85 : 341;
}
// polish_bases_array, :158-174, packed little-endian into one word; returns the count
__device__ __forceinline__ uint32_t polish_bases(uint32_t draft, uint32_t& packed)
{
#define GP_P3(a, b, c) (uint32_t(a) | (uint32_t(b) << 8) | (uint32_t(c) << 16))
  switch (draft) {
  case 'A': packed = GP_P3('T', 'C', 'G'); return 3;
  case 'T': packed = GP_P3('A', 'C', 'G'); return 3;
  case 'C': packed = GP_P3('A', 'T', 'G'); return 3;
  case 'G': packed = GP_P3('A', 'T', 'C'); return 3;
  case 'R': packed = GP_P3('T', 'C', 0); return 2;
  case 'Y': packed = GP_P3('A', 'G', 0); return 2;
  case 'S': packed = GP_P3('A', 'T', 0); return 2;
  case 'W': packed = GP_P3('C', 'G', 0); return 2;
  case 'K': packed = GP_P3('A', 'C', 0); return 2;
  case 'M': packed = GP_P3('T', 'G', 0); return 2;
  case 'B': packed = 'A'; return 1;
  case 'D': packed = 'C'; return 1;
  case 'H': packed = 'G'; return 1;
  case 'V': packed = 'T'; return 1;
  case 'N': packed = GP_P3('A', 'T', 'C') | (uint32_t('G') << 24); return 4;
  default: packed = 0; return 0;
  }
#undef GP_P3
}

struct Best {
  uint32_t type;    // 0 none, 1 substitution, 2 insertion, 3 deletion
  uint32_t support;
  uint32_t sub_base;
  uint32_t ins_first, ins_index; // insertion string = insertion_string(ins_first, ins_index)
  uint32_t del_len;
};

// tryIndels + tryDeletion, :1157-1411.  Lanes share the candidates; the reference's visiting
// order (ins_0, del_0, ins_1, del_1, ..., ins_340) breaks support ties, later wins (:1347,:1384).
__device__ bool try_indels(WS& w, uint32_t draft_char, uint32_t index_char, uint32_t& num_deletions, Best& best)
{
  const uint32_t ntry = num_tries(w.max_ins);
  const uint32_t k = w.k;
  uint32_t best_key = 0; // (support << 12) | order+1 for modes 1/2; mode 0 keeps the smallest order
  uint32_t first_key = 0xffffffffu;
  for (uint32_t i = w.lane; i < ntry; i += 32) {
    unsigned char s[8];
    const uint32_t L = insertion_string(index_char, i, s);
    // the characters that follow the candidate's first base, then the draft base (:1279),
    // packed so that indexing stays in registers
    uint64_t tailp = 0;
    for (uint32_t q = 1; q < L; q++) tailp |= uint64_t(s[q]) << (8 * (q - 1));
    tailp |= uint64_t(draft_char) << (8 * (L - 1));
    HashState t = w.hs;
    hs_changelast(t, k, draft_char, index_char); // :1290
    uint32_t present = 0;
    for (uint32_t kk = 0; kk + 1 < k; kk++) { // :1294-1326
      const uint32_t out = ring_at(w, kk);
      const uint32_t in = kk < L ? uint32_t(tailp >> (8 * kk)) & 255u : ahead_at(w, kk - L);
      hs_roll(t, k, out, in);
      if (kk % w.jump == 0 && bf_contains(w, t)) present++;
    }
    if (float(present) >= w.thrE && (w.mode == 0 || present > 0)) { // :1333-1337, :1400
      const uint32_t order = 2 * i;
      const uint32_t key = (present << 12) | (order + 1);
      best_key = max(best_key, key);
      first_key = min(first_key, order);
    }
  }
  // deletions num_deletions .. max_del, one per visited insertion index (:1359-1396)
  uint32_t ndel = 0;
  if (num_deletions <= w.max_del) ndel = min(ntry, w.max_del - num_deletions + 1);
  if (w.lane < ndel) {
    const uint32_t nd = num_deletions + w.lane;
    HashState t = w.hs;
    uint32_t present = 0;
    if (nd - 1 < w.an) { // the character that follows the deleted run must exist
      hs_changelast(t, k, draft_char, ahead_at(w, nd - 1)); // :1190-1197
      if (bf_contains(w, t)) present++;                     // :1201-1203
      for (uint32_t kk = 1; kk + 2 <= k; kk++) {            // :1204-1220
        const uint32_t ai = nd - 1 + kk;
        if (ai >= w.an) break; // roll() fails: end of contig
        hs_roll(t, k, ring_at(w, kk - 1), ahead_at(w, ai));
        if (kk % w.jump == 0 && bf_contains(w, t)) present++;
      }
    }
    if (float(present) >= w.thrD && present > 0) { // :1226-1233 (returns 0 when rejected)
      const uint32_t order = 2 * w.lane + 1;
      const uint32_t key = (present << 12) | (order + 1);
      best_key = max(best_key, key);
      first_key = min(first_key, order);
    }
  }
  num_deletions += ndel;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    best_key = max(best_key, __shfl_xor_sync(kFull, best_key, o));
    first_key = min(first_key, __shfl_xor_sync(kFull, first_key, o));
  }
  if (best_key == 0 && first_key == 0xffffffffu) return false;
  uint32_t order, support;
  if (w.mode == 0) { // first good indel wins (:1338-1344, :1377-1382)
    order = first_key;
    // its support: recompute from the lane that owns it is unnecessary -- support is only
    // recorded, never compared again in mode 0
    support = 1;
  } else {
    order = (best_key & 0xfffu) - 1;
    support = best_key >> 12;
  }
  if (w.mode == 2 && !(support > best.support)) return true; // :1401
  if (order & 1u) { best.type = 3; best.del_len = (num_deletions - ndel) + (order >> 1); }
  else { best.type = 2; best.ins_first = index_char; best.ins_index = order >> 1; }
  best.support = support;
  return true;
}

// makeEdit, :972-1154.  Returns false when the contig cannot continue (buffer overflow).
__device__ void make_edit(WS& w, uint32_t draft_char, const Best& best)
{
  const uint32_t k = w.k;
  uint32_t new_last = 0;   // character now sitting at the tail of the window
  bool stream_changed = false, reseeded = false, found = false;
  __shared__ unsigned char kmer_sh[kEditWarps][32];
  unsigned char* kmer = kmer_sh[(threadIdx.x >> 5)];
  if (w.lane == 0) {
    Rope r = { w.nd, w.nn, w.ncap, w.seq, w.len, 0 };
    uint32_t t_idx = w.t.idx, t_seq = w.t.pos, h_idx = w.h.idx, h_seq = w.h.pos;
    const EdNode tn = ld_node(r.nd + t_idx);
    switch (best.type) {
    case 1: // :1002-1033
      if (tn.type == 0) r.seq[t_seq] = (char)best.sub_base;
      else if (tn.type == 1) { EdNode x = tn; x.c = best.sub_base; st_node(r.nd + t_idx, x); }
      new_last = best.sub_base;
      break;
    case 2: { // :1034-1115
      unsigned char ins[8];
      const uint32_t L = insertion_string(best.ins_first, best.ins_index, ins);
      unsigned char prev[176];
      int np = rope_prev_insertion(r, t_seq, t_idx, prev, 160);
      if (np > 160) { r.err = 1; np = 160; }
      bool skipped = false;
      if (uint32_t(np) + L >= k) {
        if (is_repeat(prev, np) || uint32_t(np) + L >= w.icap) {
          rope_remove_prev(r, t_seq, t_idx, uint32_t(np));
          found = rope_find_kmer(r, k, h_seq, t_seq, h_idx, t_idx, kmer);
          skipped = true;
        } else {
          for (uint32_t x = 0; x < L; x++) { // :1070-1100
            for (int q = np; q > 0; q--) prev[q] = prev[q - 1];
            prev[0] = rc_char(ins[x]);
            np++;
            if (is_repeat(prev, np)) {
              rope_remove_prev(r, t_seq, t_idx, uint32_t(np) - x);
              found = rope_find_kmer(r, k, h_seq, t_seq, h_idx, t_idx, kmer);
              skipped = true;
            }
          }
        }
      }
      if (skipped) { reseeded = true; break; }
      rope_insert(r, t_idx, t_seq, ins, L);
      new_last = ins[0];
      stream_changed = true;
      break;
    }
    case 3: // :1116-1130
      rope_delete(r, t_idx, t_seq, best.del_len);
      new_last = rope_char(r, t_seq, t_idx);
      stream_changed = true;
      break;
    default: // :1131-1146
      if (w.mask) {
        const uint32_t lc = to_lower(draft_char);
        if (tn.type == 0) r.seq[t_seq] = (char)lc;
        else if (tn.type == 1) { EdNode x = tn; x.c = lc; st_node(r.nd + t_idx, x); }
        new_last = lc;
      } else new_last = draft_char;
      break;
    }
    w.nn = r.nn;
    w.err |= r.err;
    w.t.idx = t_idx; w.t.pos = t_seq; w.h.idx = h_idx; w.h.pos = h_seq;
  }
  __syncwarp();
  w.nn = __shfl_sync(kFull, w.nn, 0);
  w.err = __shfl_sync(kFull, w.err, 0);
  w.t.idx = __shfl_sync(kFull, w.t.idx, 0);
  w.t.pos = __shfl_sync(kFull, w.t.pos, 0);
  w.h.idx = __shfl_sync(kFull, w.h.idx, 0);
  w.h.pos = __shfl_sync(kFull, w.h.pos, 0);
  new_last = __shfl_sync(kFull, new_last, 0);
  stream_changed = __shfl_sync(kFull, int(stream_changed), 0) != 0;
  reseeded = __shfl_sync(kFull, int(reseeded), 0) != 0;
  found = __shfl_sync(kFull, int(found), 0) != 0;
  cur_load(w, w.t);
  cur_load(w, w.h);
  if (reseeded) {
    w.n_roll++;
    if (found) { // re-seed window, hash and look-ahead from the k-mer the reference found
      w.hs.fh = 0; w.hs.rh = 0;
      for (uint32_t i = 0; i < k; i++) w.hs.fh = srol1(w.hs.fh) ^ seed_of_char(kmer[i]);
      for (uint32_t i = 0; i < k; i++) w.hs.rh = srol1(w.hs.rh) ^ cseed_of_char(kmer[k - 1 - i]);
      if (w.lane < k) w.ring[(w.rh + w.lane) & (kRing - 1)] = kmer[w.lane];
      __syncwarp();
    }
    ahead_reset(w);
    return;
  }
  if (best.type != 0 || w.mask) {
    hs_changelast(w.hs, k, draft_char, new_last); // :1028, :1106, :1122-1129, :1145
    if (w.lane == 0) w.ring[(w.rh + k - 1) & (kRing - 1)] = (unsigned char)new_last;
    __syncwarp();
  }
  if (best.type == 0) w.n_mask += w.mask ? 1u : 0u; else w.n_edit++;
  if (stream_changed) ahead_reset(w);
}

// One contig through one k: kmerizeAndCorrect, :1414-1771.  Result is left in the rope.
__device__ void edit_round(WS& w)
{
  const uint32_t k = w.k, lane = w.lane, len = w.len;
  // findFirstAcceptedKmer(0), :392-413: smallest i with [i, i+k) accepted and i + k < len
  uint32_t h0 = len - 1;
  {
    uint32_t run = 0;
    bool found = false;
    for (uint32_t base = 0; base + 1 < len && !found; base += 32) {
      const uint32_t e = base + lane;
      const bool acc = (e + 1 < len) && is_accepted((unsigned char)w.seq[e]);
      const uint32_t bits = __ballot_sync(kFull, acc);
      for (uint32_t i = 0; i < 32; i++) {
        if ((bits >> i) & 1u) { run++; if (run >= k) { h0 = base + i + 1 - k; found = true; break; } }
        else run = 0;
      }
    }
  }
  if (len == 0) { w.nn = 0; return; }
  // root node (:1451-1456)
  if (lane == 0) { EdNode root = { 0, 0, len - 1, 0 }; st_node(w.nd, root); }
  w.nn = 1;
  __syncwarp();
  if (uint64_t(h0) + k - 1 >= len) return; // no seed: the do-loop breaks at once (:1463)
  // seed k-mer (:1441-1444)
  w.hs.fh = 0; w.hs.rh = 0;
  for (uint32_t i = 0; i < k; i++) w.hs.fh = srol1(w.hs.fh) ^ seed_of_char((unsigned char)w.seq[h0 + i]);
  for (uint32_t i = 0; i < k; i++) w.hs.rh = srol1(w.hs.rh) ^ cseed_of_char((unsigned char)w.seq[h0 + k - 1 - i]);
  w.rh = 0;
  if (lane < k) w.ring[lane] = (unsigned char)w.seq[h0 + lane];
  w.h.pos = h0; w.h.idx = 0; cur_load(w, w.h);
  w.t.pos = h0 + k - 1; w.t.idx = 0; cur_load(w, w.t);
  ahead_reset(w);
  __syncwarp();

  for (;;) {
    if (w.err) return;
    if (uint64_t(w.h.pos) + k - 1 >= len) break; // :1463
    ahead_fill(w, 32 + k + 12);
    // ---- scan: windows 0..31 from the current state ----
    // stage the next 32 incoming characters behind the window so that ring_at(j) is the
    // outgoing character of roll j even when j >= k (k < 32)
    if (lane < w.an) w.ring[(w.rh + k + lane) & (kRing - 1)] = (unsigned char)ahead_at(w, lane);
    __syncwarp();
    HashState mine = w.hs, run = w.hs;
    for (uint32_t j = 0; j < 32; j++) {
      if (j < w.an) hs_roll(run, k, ring_at(w, j), ahead_at(w, j));
      if (lane == j + 1) mine = run;
    }
    // lane j: window j exists when j rolls were possible
    const bool exists = lane <= w.an;
    const bool absent = exists && !bf_contains(w, mine);
    const bool bad = (lane < w.an) && !is_accepted(ahead_at(w, lane));
    const bool endw = w.exhausted && lane == w.an; // window checked, then roll() fails
    const uint32_t ev = __ballot_sync(kFull, absent | bad | endw);
    uint32_t adv;
    HashState at;
    if (ev == 0) {
      adv = 32;
      at = run; // state after 32 rolls
    } else {
      adv = __ffs(ev) - 1;
      at.fh = __shfl_sync(kFull, mine.fh, adv);
      at.rh = __shfl_sync(kFull, mine.rh, adv);
    }
    const bool trig = __shfl_sync(kFull, int(absent), adv & 31u) != 0 && ev != 0;
    // advance the real state by `adv` rolls
    if (adv > 0) {
      for (uint32_t i = 0; i < adv; i++) { cur_increment(w, w.h); cur_increment(w, w.t); }
      w.rh = (w.rh + adv) & (kRing - 1);
      w.a0 = (w.a0 + adv) & (kABuf - 1);
      w.an -= adv;
      w.hs = at;
      __syncwarp();
    }
    if (ev == 0) continue;
    if (uint64_t(w.h.pos) + k - 1 >= len) break; // loop-top check of the iteration we landed on

    if (trig) {
      // ---- the window is absent: look-ahead confirmation (:1470-1523) ----
      w.n_trig++;
      ahead_fill(w, k + 12);
      const uint32_t draft_char = to_upper(ring_at(w, k - 1)); // :1480
      const bool have_k = w.an >= k;
      const bool lane_ok = lane >= k || (lane < w.an && is_accepted(ahead_at(w, lane)));
      const bool all_ok = __all_sync(kFull, lane_ok) && have_k;
      if (all_ok) {
        HashState mine2 = w.hs, r2 = w.hs;
        for (uint32_t kk = 0; kk < k; kk++) {
          hs_roll(r2, k, ring_at(w, kk), ahead_at(w, kk));
          if (lane == kk) mine2 = r2;
        }
        const bool miss = lane < k && (lane % w.jump == 0) && !bf_contains(w, mine2);
        const uint32_t check_missing = __popc(__ballot_sync(kFull, miss));
        if (float(check_missing) >= w.thrM) { // :1517-1523
          uint32_t num_deletions = 1;       // :1526
          Best best = { 0, 0, 0, 0, 0, 0 };
          uint32_t packed;
          const uint32_t nb = polish_bases(draft_char, packed);
          // gate: is the k-mer ending in the candidate base present? (:1565-1570)
          HashState g = w.hs;
          const uint32_t my_base = (packed >> (8 * (lane & 3u))) & 255u;
          hs_changelast(g, k, draft_char, my_base);
          const bool gate = lane < nb && (bf_contains(w, g) || w.mode == 2);
          const uint32_t gates = __ballot_sync(kFull, gate);
          for (uint32_t b = 0; b < nb; b++) {
            if (!((gates >> b) & 1u)) continue;
            const uint32_t sub_base = (packed >> (8 * b)) & 255u;
            HashState m3 = w.hs, r3 = w.hs;
            hs_changelast(r3, k, draft_char, sub_base);
            for (uint32_t kk = 0; kk < k; kk++) { // :1585-1606
              const uint32_t out = kk + 1 < k ? ring_at(w, kk) : sub_base;
              hs_roll(r3, k, out, ahead_at(w, kk));
              if (lane == kk) m3 = r3;
            }
            const bool hit = lane < k && (lane % w.jump == 0) && bf_contains(w, m3);
            const uint32_t present = __popc(__ballot_sync(kFull, hit));
            if (float(present) >= w.thrE) { // :1621-1626
              if (present >= best.support) { best.type = 1; best.sub_base = sub_base; best.support = present; }
              if (w.mode == 0 || w.mode == 1) continue; // :1680-1682
            }
            if (w.mode == 2 || best.type != 1) { // :1686
              if (try_indels(w, draft_char, sub_base, num_deletions, best)) {
                if (w.mode == 0 || w.mode == 1) break; // :1707-1709
              }
            }
          }
          if (gates != 0u && lane == 0) {
            // a substitution trial was made and reverted with the UPPER-cased base (:1609-1615)
            if (w.t.n.type == 0) w.seq[w.t.pos] = (char)draft_char;
            else if (w.t.n.type == 1) { EdNode x = w.t.n; x.c = draft_char; st_node(w.nd + w.t.idx, x); }
          }
          __syncwarp();
          cur_load(w, w.t);
          make_edit(w, draft_char, best); // :1715-1736
          if (w.err) return;
        }
      }
    }
    // ---- roll forward, skipping k past any non-accepted incoming character (:1740-1759) ----
    long long target = -1;
    bool alive = true;
    do {
      uint32_t in;
      if (roll_main(w, in)) {
        if (!is_accepted(in)) target = (long long)w.t.pos + (long long)k;
      } else { alive = false; break; }
    } while (target >= 0 && (long long)w.t.pos != target);
    if (!alive) break;
  }
}

// writeEditsToFile body (:797-935): walk the rope until the first unset node
__device__ uint32_t emit_rope(const WS& w, char* dst, uint32_t cap, int& err)
{
  uint32_t o = 0;
  for (uint32_t i = 0; i < w.nn; i++) {
    const EdNode n = ld_node(w.nd + i);
    if (n.type == -1) break;
    if (n.type == 0) {
      if (n.s > w.len) continue; // substr would throw; never reached in valid runs
      uint32_t cnt = (n.e + 1u >= n.s) ? (n.e + 1u - n.s) : (w.len - n.s);
      if (n.s + cnt > w.len) cnt = w.len - n.s;
      if (o + cnt > cap) { err = 1; return o; }
      for (uint32_t q = w.lane; q < cnt; q += 32) dst[o + q] = w.seq[n.s + q];
      o += cnt;
    } else {
      if (o + 1 > cap) { err = 1; return o; }
      if (w.lane == 0) dst[o] = (char)n.c;
      o++;
    }
  }
  __syncwarp();
  return o;
}

__global__ void __launch_bounds__(kEditWarps * 32) edit_kernel(EditParams p)
{
  __shared__ unsigned char ring_sh[kEditWarps][kRing];
  __shared__ unsigned char ahead_sh[kEditWarps][kABuf];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  unsigned long long n_trig = 0, n_edit = 0, n_mask = 0, n_roll = 0;
  for (;;) {
    uint32_t slot = 0;
    if (lane == 0) slot = atomicAdd(p.next_contig, 1u);
    slot = __shfl_sync(kFull, slot, 0);
    if (slot >= p.n_contigs) break;
    const uint32_t ci = p.order[slot];
    const uint64_t off = p.cap_off[ci];
    const uint32_t cap = uint32_t(p.cap_off[ci + 1] - off);
    uint32_t len = p.cur_len[ci];
    uint32_t cur = 0;
    bool dropped = false;
    WS w;
    w.lane = lane;
    w.ring = ring_sh[warp];
    w.abuf = ahead_sh[warp];
    w.nd = p.nodes + p.node_off[ci];
    w.ncap = uint32_t(p.node_off[ci + 1] - p.node_off[ci]);
    w.max_ins = p.max_insertions; w.max_del = p.max_deletions; w.jump = p.jump;
    w.mode = p.mode; w.mask = p.mask;
    w.err = 0;
    w.n_trig = w.n_edit = w.n_mask = w.n_roll = 0;
    for (uint32_t ki = 0; ki < p.nk; ki++) {
      if (len < p.min_contig_len) { dropped = true; break; } // readAndCorrect, :1850
      w.seq = p.buf[cur] + off;
      w.len = len;
      w.k = p.k[ki];
      w.icap = p.insertion_cap[ki];
      w.thrM = p.thr_missing[ki]; w.thrE = p.thr_edit[ki]; w.thrD = p.thr_del[ki];
      w.mul1 = 1ull ^ (uint64_t(w.k) * kMultiSeed);
      w.mul2 = 2ull ^ (uint64_t(w.k) * kMultiSeed);
      w.mul3 = 3ull ^ (uint64_t(w.k) * kMultiSeed);
      w.bf = p.bf_pool + (uint64_t(p.contig_batch[ci]) * p.nk + ki) * kBfWords;
      w.nn = 0;
      edit_round(w);
      __syncwarp();
      int err = w.err;
      const uint32_t nl = err ? 0u : emit_rope(w, p.buf[cur ^ 1u] + off, cap, err);
      if (err) { if (lane == 0) atomicExch(p.error, 1); break; }
      __syncwarp();
      len = nl;
      cur ^= 1u;
    }
    if (lane == 0) {
      p.cur_len[ci] = len;
      p.which[ci] = (uint8_t)cur;
      p.dropped[ci] = dropped ? 1 : 0;
    }
    n_trig += w.n_trig; n_edit += w.n_edit; n_mask += w.n_mask; n_roll += w.n_roll;
  }
  if (lane == 0) {
    if (n_trig) atomicAdd(p.counters + 0, n_trig);
    if (n_edit) atomicAdd(p.counters + 1, n_edit);
    if (n_mask) atomicAdd(p.counters + 2, n_mask);
    if (n_roll) atomicAdd(p.counters + 3, n_roll);
  }
}

void launch_edit(const EditParams& p, int sm_count, cudaStream_t s)
{
  if (p.n_contigs == 0) return;
  int per_sm = 0;
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, edit_kernel, kEditWarps * 32, 0);
  if (per_sm < 1) per_sm = 1;
  uint32_t grid = uint32_t(sm_count) * uint32_t(per_sm);
  const uint32_t need = (p.n_contigs + kEditWarps - 1) / kEditWarps;
  if (grid > need) grid = need;
  edit_kernel<<<grid, kEditWarps * 32, 0, s>>>(p);
}

} // namespace gp
