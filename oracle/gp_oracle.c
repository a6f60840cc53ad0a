/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the GoldPolish hot path.
 *
 * Plain, sequential C that follows the reference's algorithm step by step; every function
 * cites the reference lines it restates (paths relative to /root/reference).  It is pinned by
 * tests/ against (a) the reference's own sources compiled under oracle/_ref/ and (b) the
 * golden vectors under tests/golden/ minted from those binaries.
 *
 * PARITY NOTE: the hash, the edit state machine, read selection and thresholds are pinned
 * against in-tree reference code.  The Bloom / counting-Bloom bit layout, the counter update
 * convention and the .bf container are btllib's (bcgsc/btllib >= 1.6.2, README.md:13,19), a
 * dependency that is absent from /root/reference: for those facts parity is UNPINNED and this
 * file follows btllib's published behaviour as restated in oracle/btllib_shim/.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use this file.
 */
#include "gp_oracle.h"

#include <ctype.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------ */
/* ntHash: subprojects/ntedit/lib/nthash.hpp                                             */
/* ------------------------------------------------------------------------------------ */

#define SEED_A 0x3c8bfbb395c60474ULL /* nthash.hpp:24 */
#define SEED_C 0x3193c18562a02b4cULL /* nthash.hpp:25 */
#define SEED_G 0x20323ed082572324ULL /* nthash.hpp:26 */
#define SEED_T 0x295549f54be24456ULL /* nthash.hpp:27 */
#define MULTI_SEED 0x90b45d39fb6da1faULL /* nthash.hpp:21 */
#define MULTI_SHIFT 27                   /* nthash.hpp:18 */

/* seedTab[c] (nthash.hpp:30-63): non-zero only at 0..7 (complement slots) and ACGTacgt. */
static uint64_t seed_of(unsigned c)
{
  switch (c) {
  case 1: return SEED_T;
  case 3: return SEED_G;
  case 4: return SEED_A;
  case 7: return SEED_C;
  case 'A': case 'a': return SEED_A;
  case 'C': case 'c': return SEED_C;
  case 'G': case 'g': return SEED_G;
  case 'T': case 't': return SEED_T;
  default: return 0;
  }
}
/* complement lookup seedTab[c & cpOff] (nthash.hpp:15,116) -- on the RAW byte */
static uint64_t cseed_of(unsigned c) { return seed_of(c & 7u); }

/* one step of the split rotation: rol1 then swapbits033 (nthash.hpp:66-91,102-104) ==
 * rotate the high 31 bits and the low 33 bits left by one, independently */
static uint64_t srol1(uint64_t v)
{
  uint64_t m = ((v & 0x8000000000000000ULL) >> 30) | ((v & 0x100000000ULL) >> 32);
  return ((v << 1) & 0xFFFFFFFDFFFFFFFFULL) | m;
}
/* inverse step: ror1 then swapbits3263 (nthash.hpp:71-73,94-97,149-150) */
static uint64_t sror1(uint64_t v)
{
  uint64_t m = ((v & 0x200000000ULL) << 30) | ((v & 1ULL) << 32);
  return ((v >> 1) & 0x7FFFFFFEFFFFFFFFULL) | m;
}
/* rotate each half by s (nthash.hpp:76-85,126-128) */
static uint64_t srol(uint64_t v, unsigned s)
{
  unsigned a = s % 31u, b = s % 33u;
  uint64_t hi = v >> 33, lo = v & 0x1FFFFFFFFULL;
  hi = ((hi << a) | (hi >> (31u - a))) & 0x7FFFFFFFULL;
  lo = ((lo << b) | (lo >> (33u - b))) & 0x1FFFFFFFFULL;
  return (hi << 33) | lo;
}

uint64_t gpo_ntf64(const char* kmer, unsigned k) /* nthash.hpp:100-108 */
{
  uint64_t h = 0;
  for (unsigned i = 0; i < k; i++) h = srol1(h) ^ seed_of((unsigned char)kmer[i]);
  return h;
}
uint64_t gpo_ntr64(const char* kmer, unsigned k) /* nthash.hpp:111-119 */
{
  uint64_t h = 0;
  for (unsigned i = 0; i < k; i++) h = srol1(h) ^ cseed_of((unsigned char)kmer[k - 1 - i]);
  return h;
}
static uint64_t ntf64_roll(uint64_t fh, unsigned k, unsigned out, unsigned in) /* :122-131 */
{
  return srol1(fh) ^ seed_of(in) ^ srol(seed_of(out), k);
}
static uint64_t ntr64_roll(uint64_t rh, unsigned k, unsigned out, unsigned in) /* :143-152 */
{
  return sror1(rh ^ srol(cseed_of(in), k) ^ cseed_of(out));
}
static uint64_t ntf64_changelast(uint64_t fh, unsigned out, unsigned in) /* :134-140 */
{
  return fh ^ seed_of(out) ^ seed_of(in);
}
static uint64_t ntr64_changelast(uint64_t rh, unsigned k, unsigned out, unsigned in) /* :154-169 */
{
  return sror1(srol1(rh) ^ srol(cseed_of(out), k) ^ srol(cseed_of(in), k));
}
void gpo_extend_hashes(uint64_t base, unsigned k, unsigned m, uint64_t* h) /* :297-301 */
{
  h[0] = base;
  for (unsigned i = 1; i < m; i++) {
    uint64_t t = base * ((uint64_t)i ^ ((uint64_t)k * MULTI_SEED));
    t ^= t >> MULTI_SHIFT;
    h[i] = t;
  }
}

/* btllib::NtHash as used at src/utils.cpp:113-114: seed on the first window free of
 * non-ACGT bytes (nthash.hpp:412-437), roll (:304-314), and when the incoming byte has no seed
 * jump past it and re-seed (lib/ntHashIterator.hpp:45-72). */
typedef struct {
  const char* seq; size_t len; unsigned k; size_t pos; uint64_t fh, rh; int init;
} nthash_it;

static int nth_init(nthash_it* it)
{
  if (it->k > it->len) return 0;
  while (it->pos <= it->len - it->k) {
    int bad = -1;
    for (int i = (int)it->k - 1; i >= 0; i--) /* scans from the right: locN = rightmost N */
      if (seed_of((unsigned char)it->seq[it->pos + (size_t)i]) == 0) { bad = i; break; }
    if (bad < 0) {
      it->fh = gpo_ntf64(it->seq + it->pos, it->k);
      it->rh = gpo_ntr64(it->seq + it->pos, it->k);
      it->init = 1;
      return 1;
    }
    it->pos += (size_t)bad + 1;
  }
  return 0;
}
static int nth_roll(nthash_it* it)
{
  if (!it->init) return nth_init(it);
  if (it->pos >= it->len - it->k) return 0;
  unsigned in = (unsigned char)it->seq[it->pos + it->k];
  if (seed_of(in) == 0) {
    it->pos += it->k;
    it->init = 0;
    return nth_init(it);
  }
  unsigned out = (unsigned char)it->seq[it->pos];
  it->fh = ntf64_roll(it->fh, it->k, out, in);
  it->rh = ntr64_roll(it->rh, it->k, out, in);
  it->pos++;
  return 1;
}

size_t gpo_nthash_all(const char* seq, size_t len, unsigned k, size_t cap, uint64_t* pos, uint64_t* hashes)
{
  nthash_it it = { seq, len, k, 0, 0, 0, 0 };
  size_t n = 0;
  while (nth_roll(&it)) {
    if (n < cap) {
      pos[n] = it.pos;
      gpo_extend_hashes(it.fh + it.rh, k, GPO_HASH_NUM, hashes + 4 * n);
    }
    n++;
  }
  return n;
}

/* ------------------------------------------------------------------------------------ */
/* Filter build: src/goldpolish_targeted_bfs.cpp, src/utils.cpp                          */
/* ------------------------------------------------------------------------------------ */

int gpo_kmer_threshold(uint64_t mappings_bases) /* goldpolish_targeted_bfs.cpp:45-53 */
{
  int t = (int)round(4.66943 + (double)mappings_bases * 2.11391e-07);
  return t < 13 ? t : 13;
}

uint64_t gpo_mappings_cap(uint64_t target_len, double s) /* goldpolish_targeted_bfs.cpp:96-99 */
{
  return (uint64_t)((double)target_len * s / 10000.0);
}

/* btllib KmerCountingBloomFilter8::insert_thresh_contains followed by the gated
 * KmerBloomFilter::insert (src/utils.cpp:115-119; conventions: oracle/btllib_shim). */
static void cbf_bf_update(uint8_t* cbf, uint8_t* bf, const uint64_t* h, unsigned thr)
{
  size_t idx[GPO_HASH_NUM];
  unsigned mn = 255;
  for (int i = 0; i < GPO_HASH_NUM; i++) {
    idx[i] = (size_t)(h[i] % GPO_CBF_COUNTERS);
    if (cbf[idx[i]] < mn) mn = cbf[idx[i]];
  }
  unsigned count = mn;
  if (mn < thr) {
    for (int i = 0; i < GPO_HASH_NUM; i++)
      if (cbf[idx[i]] == mn) cbf[idx[i]] = (uint8_t)(mn + 1);
    count = mn + 1;
  }
  if (count >= thr) {
    for (int i = 0; i < GPO_HASH_NUM; i++) {
      uint64_t n = h[i] % (GPO_BF_BYTES * 8ULL);
      bf[n >> 3] |= (uint8_t)(1u << (n & 7u));
    }
  }
}

long gpo_fill_bfs(const char* seq, size_t len, const unsigned* ks, int nk, unsigned T,
                  uint8_t* const* cbfs, uint8_t* const* bfs) /* src/utils.cpp:96-123 */
{
  if (T < 4) return -1; /* utils.cpp:105-107 */
  unsigned thr = T - 2; /* utils.cpp:108 */
  long ops = 0;
  for (int i = 0; i < nk; i++, thr++) { /* k outermost within a read, utils.cpp:109,121 */
    nthash_it it = { seq, len, ks[i], 0, 0, 0, 0 };
    uint64_t h[GPO_HASH_NUM];
    while (nth_roll(&it)) {
      gpo_extend_hashes(it.fh + it.rh, ks[i], GPO_HASH_NUM, h);
      cbf_bf_update(cbfs[i], bfs[i], h, thr);
      ops++;
    }
  }
  return ops;
}

/* ------------------------------------------------------------------------------------ */
/* ntEdit: subprojects/ntedit/ntedit.cpp                                                 */
/* ------------------------------------------------------------------------------------ */

typedef struct { int type; uint32_t s, e; unsigned char c; } node_t; /* seqNode, :468-475 */

typedef struct { uint32_t h_seq, t_seq, h_node, t_node; } cursor_t;
typedef struct { uint64_t fh, rh; } hstate_t;

typedef struct {
  char* seq;      /* contigSeq (substitutions / masks are written into it, :1005,1135) */
  uint32_t len;   /* contigSeq.size() */
  node_t* nodes;  /* newSeq */
  size_t n, cap;  /* newSeq.size() */
  const uint8_t* bf;
  uint64_t bf_bits;
  gpo_ntedit_opts o;
  unsigned insertion_cap;
  float thr_missing, thr_edit, thr_del;
  gpo_ntedit_stats st;
} ed_t;

static int is_accepted(int c) /* isAcceptedBase(toupper(c)), :363-367 */
{
  c = toupper(c);
  return c == 'A' || c == 'T' || c == 'G' || c == 'C' || c == 'R' || c == 'Y' || c == 'S' || c == 'W' ||
         c == 'K' || c == 'M' || c == 'B' || c == 'D' || c == 'H' || c == 'V';
}
static char rc_char(unsigned char c) /* RC, :369-388 */
{
  switch (c) {
  case 'A': case 'a': return 'T';
  case 'T': case 't': return 'A';
  case 'G': case 'g': return 'C';
  case 'C': case 'c': return 'G';
  default: return 'N';
  }
}

static int bf_contains(const ed_t* e, const hstate_t* hs) /* btllib contains(), :1470 */
{
  uint64_t h[8];
  gpo_extend_hashes(hs->fh + hs->rh, e->o.k, e->o.hash_num, h);
  for (unsigned i = 0; i < e->o.hash_num; i++) {
    uint64_t n = h[i] % e->bf_bits;
    if (!((e->bf[n >> 3] >> (n & 7u)) & 1u)) return 0;
  }
  return 1;
}

static void nodes_set(ed_t* e, size_t i, node_t v) /* "if (i < size) a[i]=v else push_back(v)" */
{
  if (i < e->n) { e->nodes[i] = v; return; }
  if (i != e->n) { e->st.ref_ub++; return; }
  if (e->n == e->cap) {
    e->cap = e->cap ? e->cap * 2 : 64;
    e->nodes = (node_t*)realloc(e->nodes, e->cap * sizeof(node_t));
  }
  e->nodes[e->n++] = v;
}

static unsigned char get_char(ed_t* e, uint32_t pos, size_t node_index) /* getCharacter, :667-678 */
{
  if (node_index >= e->n) { e->st.ref_ub++; return 0; }
  const node_t* nd = &e->nodes[node_index];
  if (nd->type == 0) {
    if (pos >= e->len) { e->st.ref_ub++; return 0; } /* std::string::at would throw */
    return (unsigned char)e->seq[pos];
  }
  if (nd->type == 1) return nd->c;
  return 0;
}

static void increment(ed_t* e, uint32_t* pos, uint32_t* node_index) /* :681-699 */
{
  if (*node_index >= e->n) { e->st.ref_ub++; return; }
  const node_t nd = e->nodes[*node_index];
  if (nd.type == 0) {
    (*pos)++;
    if (*pos > nd.e) {
      (*node_index)++;
      if (*node_index < e->n && e->nodes[*node_index].type == 0) *pos = e->nodes[*node_index].s;
    }
  } else if (nd.type == 1) {
    (*node_index)++;
    if (*node_index < e->n && e->nodes[*node_index].type == 0) *pos = e->nodes[*node_index].s;
  }
}

static int roll(ed_t* e, cursor_t* c, unsigned char* out, unsigned char* in) /* :939-969 */
{
  if (c->h_seq >= e->len || c->h_node >= e->n) return 0;
  *out = get_char(e, c->h_seq, c->h_node);
  increment(e, &c->h_seq, &c->h_node);
  if (c->t_seq >= e->len || c->t_node >= e->n) return 0;
  increment(e, &c->t_seq, &c->t_node);
  if (c->t_seq >= e->len || c->t_node >= e->n) return 0;
  *in = get_char(e, c->t_seq, c->t_node);
  return 1;
}

static void hs_roll(const ed_t* e, hstate_t* hs, unsigned out, unsigned in) /* NTMC64 roll, nthash.hpp:304-314 */
{
  hs->fh = ntf64_roll(hs->fh, e->o.k, out, in);
  hs->rh = ntr64_roll(hs->rh, e->o.k, out, in);
}
static void hs_changelast(const ed_t* e, hstate_t* hs, unsigned out, unsigned in) /* nthash.hpp:316-325 */
{
  hs->fh = ntf64_changelast(hs->fh, out, in);
  hs->rh = ntr64_changelast(hs->rh, e->o.k, out, in);
}

static uint32_t find_first_accepted_kmer(const ed_t* e, uint32_t b_i) /* :392-413 */
{
  const uint32_t k = e->o.k;
  for (uint32_t i = b_i; (uint64_t)i + k < e->len;) {
    if (is_accepted((unsigned char)e->seq[i])) {
      int good = 1;
      for (uint32_t j = i + 1; j < i + k; j++) {
        if (!is_accepted((unsigned char)e->seq[j])) { good = 0; i = j + 1; break; }
      }
      if (good) return i;
    } else i++;
  }
  return e->len - 1;
}

/* i-th entry of multi_possible_bases[first] (:198-343): strings of length 1..5 that start with
 * `first`, ordered by length then lexicographically over A<C<G<T (num_tries, :150). */
int gpo_insertion_string(unsigned char first, int i, char out[6])
{
  static const int start[6] = { 0, 0, 1, 5, 21, 85 };
  static const char B[4] = { 'A', 'C', 'G', 'T' };
  int L = 1;
  while (L < 5 && i >= start[L + 1]) L++;
  int r = i - start[L];
  out[0] = (char)first;
  for (int p = L - 1; p >= 1; p--) { out[p] = B[r & 3]; r >>= 2; }
  out[L] = 0;
  return L;
}
static const int NUM_TRIES[6] = { 0, 1, 5, 21, 85, 341 }; /* :150 */

/* polish_bases_array (:158-174); returns count */
static int polish_bases(unsigned char draft, unsigned char out[4])
{
  const char* s;
  switch (draft) {
  case 'A': s = "TCG"; break;
  case 'T': s = "ACG"; break;
  case 'C': s = "ATG"; break;
  case 'G': s = "ATC"; break;
  case 'R': s = "TC"; break;
  case 'Y': s = "AG"; break;
  case 'S': s = "AT"; break;
  case 'W': s = "CG"; break;
  case 'K': s = "AC"; break;
  case 'M': s = "TG"; break;
  case 'B': s = "A"; break;
  case 'D': s = "C"; break;
  case 'H': s = "G"; break;
  case 'V': s = "T"; break;
  case 'N': s = "ATCG"; break;
  default: s = ""; break; /* operator[] would default-construct an empty vector (:1559) */
  }
  int n = 0;
  while (s[n]) { out[n] = (unsigned char)s[n]; n++; }
  return n;
}

static int is_repeat_insertion(const char* s, int n) /* computeLPSArray + isRepeatInsertion, :416-451 */
{
  if (n <= 0) return 0; /* n == 0: the reference indexes lps[-1]; treated as not-a-repeat */
  int* lps = (int*)calloc((size_t)n, sizeof(int));
  int len = 0, i = 1;
  while (i < n) {
    if (s[i] == s[len]) { len++; lps[i] = len; i++; }
    else if (len != 0) len = lps[len - 1];
    else { lps[i] = 0; i++; }
  }
  int l = lps[n - 1];
  free(lps);
  return l > 0 && n % (n - l) == 0;
}

/* getPrevInsertion, :762-777.  Returns length, fills buf (capacity cap). */
static int get_prev_insertion(ed_t* e, uint32_t t_seq, uint32_t t_node, char* buf, int cap)
{
  int n = 0;
  size_t idx = t_node;
  if ((idx < e->n && e->nodes[idx].type == 0 && t_seq == e->nodes[idx].s) ||
      (idx < e->n && e->nodes[idx].type == 1))
    idx--; /* unsigned wrap when 0: the loop below then does not run */
  while (idx < e->n && e->nodes[idx].type == 1) {
    if (n < cap) buf[n] = rc_char(e->nodes[idx].c);
    n++;
    idx--;
  }
  return n;
}

/* findAcceptedKmer, :703-758.  Returns 1 and the k-mer in kmer[] if found. */
static int find_accepted_kmer(ed_t* e, cursor_t* c, char* kmer)
{
  const uint32_t k = e->o.k;
  uint32_t temp_t_node = c->t_node, temp_h_node = 0;
  uint32_t i = c->t_seq;
  size_t curr = c->t_node; /* curr_node = newSeq[t_node_index] (:714) */
  if (curr >= e->n) { e->st.ref_ub++; }
  while (i < e->len && temp_t_node < e->n && e->nodes[temp_t_node].type != -1) {
    unsigned char ch = (curr < e->n) ? get_char(e, i, curr) : 0;
    if (is_accepted(ch)) {
      uint32_t kl = 0;
      kmer[kl++] = (char)ch;
      temp_h_node = temp_t_node;
      uint32_t j = i;
      increment(e, &j, &temp_t_node);
      while (j < e->len && temp_t_node < e->n && e->nodes[temp_t_node].type != -1) {
        curr = temp_t_node;
        ch = get_char(e, j, curr);
        if (!is_accepted(ch)) { i = j; break; }
        kmer[kl++] = (char)ch;
        if (kl == k) break;
        increment(e, &j, &temp_t_node);
      }
      if (kl == k) {
        c->h_seq = i; c->t_seq = j; c->h_node = temp_h_node; c->t_node = temp_t_node;
        return 1;
      }
    }
    if (temp_t_node < e->n) increment(e, &i, &temp_t_node);
    else break;
  }
  c->h_seq = e->len;
  c->t_seq = e->len;
  return 0;
}

/* makeInsertion, :480-569 */
static void make_insertion(ed_t* e, uint32_t* t_node, uint32_t insert_pos, const char* ins, int L)
{
  const node_t orig = e->nodes[*t_node];
  if (orig.type == 0 && !((int)insert_pos <= (int)orig.s)) {
    /* split the position node: [s..ip-1] ins... [ip..e] */
    node_t after = { 0, insert_pos, orig.e, 0 };
    e->nodes[*t_node].e = insert_pos - 1;
    for (int i = 0; i < L; i++) {
      node_t nd = { 1, 0, 0, (unsigned char)ins[i] };
      nodes_set(e, (size_t)*t_node + (size_t)i + 1, nd);
    }
    nodes_set(e, (size_t)*t_node + (size_t)L + 1, after);
    (*t_node)++;
    return;
  }
  if (orig.type == 0 || orig.type == 1) {
    /* insert in front of this node: lift the tail of the list, write, re-append */
    size_t i = *t_node, nre = 0;
    while (i < e->n && e->nodes[i].type != -1) { nre++; i++; }
    node_t* re = (node_t*)malloc((nre ? nre : 1) * sizeof(node_t));
    for (size_t q = 0; q < nre; q++) { re[q] = e->nodes[*t_node + q]; e->nodes[*t_node + q].type = -1; }
    for (int q = 0; q < L; q++) {
      node_t nd = { 1, 0, 0, (unsigned char)ins[q] };
      nodes_set(e, (size_t)*t_node + (size_t)q, nd);
    }
    for (size_t q = 0; q < nre; q++) nodes_set(e, (size_t)*t_node + (size_t)L + q, re[q]);
    free(re);
  }
}

/* makeDeletion, :574-664 */
static void make_deletion(ed_t* e, uint32_t* t_node, uint32_t* pos, uint32_t num_del)
{
  if (*t_node >= e->n) { e->st.ref_ub++; return; }
  const node_t orig = e->nodes[*t_node];
  if (orig.type == 0) {
    uint32_t leftover = 0;
    if (*pos <= orig.s) {
      if (*pos + num_del <= orig.e) { /* off the front of a position node */
        e->nodes[*t_node].s = *pos + num_del;
        *pos = e->nodes[*t_node].s;
        return;
      }
      leftover = *pos + num_del - orig.e; /* sic (:594) */
      *pos = orig.e + 1;
      size_t i = (size_t)*t_node + 1;
      while (i < e->n && e->nodes[i].type != -1) {
        e->nodes[i - 1] = e->nodes[i];
        e->nodes[i].type = -1;
        i++;
      }
    } else {
      if (*pos + num_del <= orig.e) { /* from the middle of a position node */
        node_t split = { 0, *pos + num_del, orig.e, 0 };
        e->nodes[*t_node].e = *pos - 1;
        *pos = split.s;
        (*t_node)++;
        nodes_set(e, *t_node, split);
        return;
      }
      leftover = *pos + num_del - orig.e; /* sic (:622) */
      e->nodes[*t_node].e = *pos - 1;
      *pos = orig.e + 1;
      (*t_node)++;
    }
    if (leftover > 0 && *t_node < e->n && e->nodes[*t_node].type != -1) {
      if (e->nodes[*t_node].type == 0) *pos = e->nodes[*t_node].s;
      make_deletion(e, t_node, pos, leftover);
    }
  } else if (orig.type == 1) {
    size_t i = *t_node;
    uint32_t leftover = num_del;
    while (i < e->n && e->nodes[i].type == 1 && leftover > 0) { e->nodes[i].type = -1; leftover--; i++; }
    size_t j = *t_node;
    while (i < e->n && e->nodes[i].type != -1) {
      e->nodes[j] = e->nodes[i];
      e->nodes[i].type = -1;
      i++; j++;
    }
    if (leftover > 0 && *t_node < e->n && e->nodes[*t_node].type != -1) {
      if (e->nodes[*t_node].type == 0) *pos = e->nodes[*t_node].s;
      make_deletion(e, t_node, pos, leftover);
    }
  }
}

/* tryDeletion, :1157-1234.  Returns the support (>0) when accepted, else 0. */
static unsigned try_deletion(ed_t* e, unsigned char draft_char, unsigned num_del, const cursor_t* cur,
                             const hstate_t* hs)
{
  cursor_t c = *cur;
  hstate_t t = *hs;
  unsigned char out = 0, in = 0;
  for (unsigned i = 0; i < num_del; i++) {
    (void)get_char(e, c.t_seq, c.t_node); /* deleted_bases += ... */
    increment(e, &c.t_seq, &c.t_node);
  }
  hs_changelast(e, &t, draft_char, get_char(e, c.t_seq, c.t_node));
  unsigned present = 0;
  if (bf_contains(e, &t)) present++;
  for (unsigned kk = 1; kk <= e->o.k - 2 && c.h_seq < e->len; kk++) {
    if (roll(e, &c, &out, &in)) {
      hs_roll(e, &t, out, in);
      if (kk % e->o.jump == 0 && bf_contains(e, &t)) present++;
    }
  }
  return ((float)present >= e->thr_del) ? present : 0;
}

typedef struct { unsigned type; char indel[16]; int indel_len; unsigned support; unsigned char sub_base; } best_t;

/* tryIndels, :1237-1411 */
static int try_indels(ed_t* e, unsigned char draft_char, unsigned char index_char, unsigned* num_deletions,
                      const cursor_t* cur, const hstate_t* hs, best_t* best)
{
  unsigned tb_support = 0, tb_type = 0;
  char tb_indel[16];
  int tb_len = 0;
  unsigned char out = 0, in = 0;
  e->st.indel_calls++;
  for (int i = 0; i < NUM_TRIES[e->o.max_insertions]; i++) {
    char ins[8];
    int L = gpo_insertion_string(index_char, i, ins);
    ins[L] = (char)draft_char; /* insertion_bases += draft_char (:1279) */
    cursor_t c = *cur;
    hstate_t t = *hs;
    hs_changelast(e, &t, draft_char, index_char);
    unsigned present = 0, kk = 0;
    for (; kk < (unsigned)L && c.h_seq < e->len; kk++) { /* :1294-1308 */
      hs_roll(e, &t, get_char(e, c.h_seq, c.h_node), (unsigned char)ins[kk + 1]);
      increment(e, &c.h_seq, &c.h_node);
      if (kk % e->o.jump == 0 && bf_contains(e, &t)) present++;
    }
    for (; kk < e->o.k - 1 && c.h_seq < e->len; kk++) { /* :1310-1326 */
      if (roll(e, &c, &out, &in)) {
        hs_roll(e, &t, out, in);
        if (kk % e->o.jump == 0 && bf_contains(e, &t)) present++;
      }
    }
    if ((float)present >= e->thr_edit) { /* :1333-1337 (use_ratio) */
      if (e->o.mode == 0) {
        best->type = 2; memcpy(best->indel, ins, (size_t)L); best->indel_len = L; best->support = present;
        return 1;
      }
      if (present >= tb_support) { /* :1347 */
        tb_type = 2; memcpy(tb_indel, ins, (size_t)L); tb_len = L; tb_support = present;
      }
    }
    if (*num_deletions <= e->o.max_deletions) { /* :1359-1396 */
      unsigned ds = try_deletion(e, draft_char, *num_deletions, cur, hs);
      if (ds > 0) {
        if (e->o.mode == 0) {
          best->type = 3; best->indel_len = (int)*num_deletions; best->support = ds;
          return 1;
        }
        if (ds >= tb_support) { tb_type = 3; tb_len = (int)*num_deletions; tb_support = ds; }
      }
      (*num_deletions)++;
    }
  }
  if (tb_support > 0) { /* :1400-1409 */
    if ((e->o.mode == 2 && tb_support > best->support) || e->o.mode == 1) {
      best->type = tb_type;
      best->indel_len = tb_len;
      if (tb_type == 2) memcpy(best->indel, tb_indel, (size_t)tb_len);
      best->support = tb_support;
    }
    return 1;
  }
  return 0;
}

/* the node shuffle shared by both low-complexity branches of makeEdit (:1043-1056, :1074-1088) */
static void remove_prev_insertion(ed_t* e, const cursor_t* c, size_t count)
{
  size_t j = 1;
  if (c->t_node < e->n && e->nodes[c->t_node].type == 0 && c->t_seq == e->nodes[c->t_node].s) j = 0;
  for (size_t i = count; i > 0; i--) {
    if (i > c->t_node) { e->st.ref_ub++; continue; } /* would index before the vector */
    if ((size_t)c->t_node + j < e->n && e->nodes[c->t_node + j].type != -1) {
      e->nodes[c->t_node - i] = e->nodes[c->t_node + j];
      e->nodes[c->t_node + j].type = -1;
      j++;
    } else {
      e->nodes[c->t_node - i].type = -1;
    }
  }
}

static void reseed_after_rollback(ed_t* e, cursor_t* c, hstate_t* hs) /* :1057-1065, :1089-1097 */
{
  char kmer[64];
  if (find_accepted_kmer(e, c, kmer)) {
    hs->fh = gpo_ntf64(kmer, e->o.k);
    hs->rh = gpo_ntr64(kmer, e->o.k);
  } else {
    /* the reference hashes k bytes starting at an empty string's terminator (out-of-bounds
       read); the value is never observed because h_seq_i == size ends the contig (:952). */
    hs->fh = hs->rh = 0;
  }
}

/* makeEdit, :972-1154 */
static void make_edit(ed_t* e, unsigned char draft_char, const best_t* best, cursor_t* c, hstate_t* hs)
{
  if (c->t_node >= e->n) { e->st.ref_ub++; return; }
  const node_t tnode = e->nodes[c->t_node];
  switch (best->type) {
  case 1: /* substitution, :1002-1033 */
    if (tnode.type == 0) e->seq[c->t_seq] = (char)best->sub_base;
    else if (tnode.type == 1) e->nodes[c->t_node].c = best->sub_base;
    hs_changelast(e, hs, draft_char, best->sub_base);
    e->st.subs++;
    break;
  case 2: { /* insertion, :1034-1115 */
    char prev[4096];
    int np = get_prev_insertion(e, c->t_seq, c->t_node, prev, (int)sizeof prev - 8);
    if (np > (int)sizeof prev - 8) { e->st.ref_ub++; np = (int)sizeof prev - 8; }
    const int L = best->indel_len;
    int skipped = 0;
    if ((unsigned)(np + L) >= e->o.k) {
      if (is_repeat_insertion(prev, np) || (unsigned)(np + L) >= e->insertion_cap) {
        remove_prev_insertion(e, c, (size_t)np);
        reseed_after_rollback(e, c, hs);
        skipped = 1;
      } else {
        for (int w = 0; w < L; w++) { /* :1070-1100 */
          memmove(prev + 1, prev, (size_t)np);
          prev[0] = rc_char((unsigned char)best->indel[w]);
          np++;
          if (is_repeat_insertion(prev, np)) {
            remove_prev_insertion(e, c, (size_t)(np - w));
            reseed_after_rollback(e, c, hs);
            skipped = 1;
          }
        }
      }
    }
    if (skipped) { e->st.rollbacks++; break; }
    make_insertion(e, &c->t_node, c->t_seq, best->indel, L);
    hs_changelast(e, hs, draft_char, (unsigned char)best->indel[0]);
    e->st.inss++;
    break;
  }
  case 3: /* deletion, :1116-1130 */
    make_deletion(e, &c->t_node, &c->t_seq, (uint32_t)best->indel_len);
    hs_changelast(e, hs, draft_char, get_char(e, c->t_seq, c->t_node));
    e->st.dels++;
    break;
  case 0: /* no fix: soft-mask, :1131-1146 */
    if (e->o.mask) {
      unsigned char lc = (unsigned char)tolower(draft_char);
      if (tnode.type == 0) e->seq[c->t_seq] = (char)lc;
      else if (tnode.type == 1) e->nodes[c->t_node].c = lc;
      hs_changelast(e, hs, draft_char, lc);
      e->st.masks++;
    }
    break;
  default: break;
  }
}

void gpo_ntedit_default_opts(gpo_ntedit_opts* o, unsigned k) /* scripts/goldpolish-ntedit:27 */
{
  o->k = k; o->hash_num = GPO_HASH_NUM; o->max_insertions = 5; o->max_deletions = 5; o->mode = 1;
  o->mask = 1; o->missing_ratio = 0.5f; o->edit_ratio = 0.5f; o->jump = 3; o->min_contig_len = 100;
}

int gpo_guard_rejects(uint64_t input_size, uint64_t output_size) /* scripts/goldpolish-ntedit:31-34 */
{
  if (input_size == 0) return 0;
  return (output_size * 10000ULL) / input_size < 7500ULL; /* bc scale=4 truncates */
}

/* kmerizeAndCorrect, :1414-1771, followed by the FASTA body of writeEditsToFile, :780-936 */
long gpo_ntedit_contig(const char* seq_in, size_t len, const uint8_t* bf, size_t bf_bytes,
                       const gpo_ntedit_opts* opts, char* outbuf, size_t cap, gpo_ntedit_stats* stats)
{
  if (len < opts->min_contig_len) return -1; /* readAndCorrect, :1850 */
  ed_t E;
  memset(&E, 0, sizeof E);
  ed_t* e = &E;
  e->o = *opts;
  e->seq = (char*)malloc(len + 1);
  memcpy(e->seq, seq_in, len);
  e->seq[len] = 0;
  e->len = (uint32_t)len;
  e->bf = bf;
  e->bf_bits = (uint64_t)bf_bytes * 8ULL;
  e->insertion_cap = (unsigned)((float)opts->k * 1.5f); /* :2024-2025 */
  /* float thresholds exactly as written at :1521-1523, :1624-1626 / :1335-1337, :1228-1230 */
  e->thr_missing = ((float)opts->k / (float)opts->jump) * opts->missing_ratio;
  e->thr_edit = ((float)opts->k / (float)opts->jump) * opts->edit_ratio;
  e->thr_del = (1 + ((float)opts->k / (float)opts->jump)) * opts->edit_ratio;
  const uint32_t k = opts->k;

  hstate_t hs = { 0, 0 };
  unsigned char char_in = 0, char_out = 0;
  cursor_t c;
  c.h_seq = find_first_accepted_kmer(e, 0);
  c.t_seq = c.h_seq + k - 1;
  if ((uint64_t)c.h_seq + k - 1 < len) { /* :1441-1444 */
    hs.fh = gpo_ntf64(e->seq + c.h_seq, k);
    hs.rh = gpo_ntr64(e->seq + c.h_seq, k);
    char_in = (unsigned char)e->seq[c.t_seq];
  }
  node_t root = { 0, 0, (uint32_t)len - 1, 0 };
  nodes_set(e, 0, root);
  c.h_node = c.t_node = 0;

  int continue_edit = 1;
  do {
    if ((uint64_t)c.h_seq + k - 1 >= len) break; /* :1463 */
    if (!bf_contains(e, &hs)) {                /* :1470 */
      e->st.triggers++;
      cursor_t tc = c;
      hstate_t th = hs;
      const unsigned char draft_char = (unsigned char)toupper(char_in); /* :1480 */
      unsigned check_missing = 0;
      int do_not_fix = 0;
      for (unsigned kk = 0; kk < k && tc.h_seq < len; kk++) { /* :1487-1512 */
        if (roll(e, &tc, &char_out, &char_in)) {
          hs_roll(e, &th, char_out, char_in);
          if (!is_accepted(char_in)) { do_not_fix = 1; break; }
          if (kk % opts->jump == 0 && !bf_contains(e, &th)) check_missing++;
        } else { do_not_fix = 1; break; }
      }
      if (!do_not_fix && (float)check_missing >= e->thr_missing) { /* :1517-1523 */
        e->st.attempts++;
        unsigned num_deletions = 1; /* :1526 */
        best_t best;
        memset(&best, 0, sizeof best);
        unsigned char bases[4];
        const int nb = polish_bases(draft_char, bases);
        for (int b = 0; b < nb; b++) { /* :1559-1713 */
          const unsigned char sub_base = bases[b];
          th = hs;
          hs_changelast(e, &th, draft_char, sub_base);
          if (!bf_contains(e, &th) && opts->mode != 2) continue; /* :1569-1570 */
          tc = c;
          const node_t tn = e->nodes[c.t_node];
          if (tn.type == 0) e->seq[tc.t_seq] = (char)sub_base; /* :1578-1582 */
          else if (tn.type == 1) e->nodes[c.t_node].c = sub_base;
          unsigned present = 0;
          for (unsigned kk = 0; kk < k && tc.h_seq < len && tc.t_seq < len; kk++) { /* :1585-1606 */
            if (roll(e, &tc, &char_out, &char_in)) {
              hs_roll(e, &th, char_out, char_in);
              if (kk % opts->jump == 0 && bf_contains(e, &th)) present++;
            } else break;
          }
          if (tn.type == 0) e->seq[c.t_seq] = (char)draft_char; /* revert with the UPPER-cased char, :1609-1615 */
          else if (tn.type == 1) e->nodes[c.t_node].c = draft_char;
          if ((float)present >= e->thr_edit) { /* :1621-1626 */
            if (present >= best.support) {   /* :1629 (>= : later base wins ties) */
              best.type = 1; best.sub_base = sub_base; best.support = present;
            }
            if (opts->mode == 0 || opts->mode == 1) continue; /* :1680-1682 */
          }
          if (opts->mode == 2 || best.type != 1) { /* :1686 */
            if (try_indels(e, draft_char, sub_base, &num_deletions, &c, &hs, &best)) {
              if (opts->mode == 0 || opts->mode == 1) break; /* :1707-1709 */
            }
          }
        }
        make_edit(e, draft_char, &best, &c, &hs); /* :1715-1736 */
      }
    }
    /* roll forward, skipping k past any non-accepted incoming char, :1740-1759 */
    long target = -1;
    do {
      if (roll(e, &c, &char_out, &char_in)) {
        if (!is_accepted(char_in)) target = (long)c.t_seq + (long)k;
        hs_roll(e, &hs, char_out, char_in);
      } else { continue_edit = 0; break; }
    } while (target >= 0 && (long)c.t_seq != target);
  } while (continue_edit);

  /* writeEditsToFile body, :797-935 */
  size_t o = 0;
  long rc = 0;
  for (size_t i = 0; i < e->n && e->nodes[i].type != -1; i++) {
    const node_t* nd = &e->nodes[i];
    if (nd->type == 0) {
      /* substr(s_pos, e_pos - s_pos + 1) with size_t arithmetic: e_pos < s_pos wraps to a
         huge count == "to the end of the string"; s_pos > size throws */
      size_t s = nd->s, cnt;
      if (s > len) { e->st.ref_ub++; continue; }
      if (nd->e + 1u >= nd->s) cnt = (size_t)nd->e + 1 - s; else cnt = len - s;
      if (s + cnt > len) cnt = len - s;
      if (o + cnt > cap) { rc = -2; break; }
      memcpy(outbuf + o, e->seq + s, cnt);
      o += cnt;
    } else {
      if (o + 1 > cap) { rc = -2; break; }
      outbuf[o++] = (char)nd->c;
    }
  }
  if (stats) *stats = e->st;
  free(e->seq);
  free(e->nodes);
  return rc < 0 ? rc : (long)o;
}
