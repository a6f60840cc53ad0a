/* TEST INFRASTRUCTURE ONLY -- CPU restatement (oracle) of the GoldPolish hot path.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this library;
 * the product (goldpolish_b200/) never does.  See gp_oracle.c for the reference citations.
 */
#ifndef GP_ORACLE_H
#define GP_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GPO_CBF_COUNTERS 10485760ULL /* 10 MiB of 8-bit counters, goldpolish_targeted_bfs.cpp:270 */
#define GPO_BF_BYTES 524288ULL       /* 512 KiB, goldpolish_targeted_bfs.cpp:271 */
#define GPO_HASH_NUM 4               /* goldpolish_targeted_bfs.cpp:272 */

/* ---- ntHash (lib/nthash.hpp) ---- */
uint64_t gpo_ntf64(const char* kmer, unsigned k);
uint64_t gpo_ntr64(const char* kmer, unsigned k);
void gpo_extend_hashes(uint64_t base, unsigned k, unsigned m, uint64_t* h);
/* btllib::NtHash iteration (valid ACGT k-mers only, positions ascending).  Returns the number
 * of k-mers; fills pos[i], hashes[4i..4i+3] for i < cap. */
size_t gpo_nthash_all(const char* seq, size_t len, unsigned k, size_t cap, uint64_t* pos, uint64_t* hashes);

/* ---- filter build ---- */
int gpo_kmer_threshold(uint64_t mappings_bases);
/* One read through fill_bfs.  cbfs[i]: GPO_CBF_COUNTERS bytes, bfs[i]: GPO_BF_BYTES bytes.
 * Returns the number of k-mer ops (valid k-mers summed over k), or -1 if T < 4. */
long gpo_fill_bfs(const char* seq, size_t len, const unsigned* ks, int nk, unsigned kmer_threshold,
                  uint8_t* const* cbfs, uint8_t* const* bfs);
/* Reads-per-target cap: size_t(double(len) * s / 10000.0). */
uint64_t gpo_mappings_cap(uint64_t target_len, double subsample_max_per_10kbp);

/* ---- ntEdit ---- */
typedef struct {
  unsigned k;              /* from the .bf header */
  unsigned hash_num;       /* from the .bf header */
  unsigned max_insertions; /* -i (0..5) */
  unsigned max_deletions;  /* -d (0..10) */
  int mode;                /* -m (0,1,2) */
  int mask;                /* -a */
  float missing_ratio;     /* -X */
  float edit_ratio;        /* -Y */
  unsigned jump;           /* -j */
  unsigned min_contig_len; /* -z */
} gpo_ntedit_opts;

typedef struct {
  uint32_t triggers, attempts, subs, inss, dels, masks, rollbacks, indel_calls;
  uint32_t ref_ub; /* the reference would have read out of bounds / thrown on this input */
} gpo_ntedit_stats;

void gpo_ntedit_default_opts(gpo_ntedit_opts* o, unsigned k);
/* One contig through kmerizeAndCorrect.  Returns the edited length (written to out, no
 * terminator), -1 if len < min_contig_len (record dropped), -2 if cap is too small. */
long gpo_ntedit_contig(const char* seq, size_t len, const uint8_t* bf, size_t bf_bytes,
                       const gpo_ntedit_opts* opts, char* out, size_t cap, gpo_ntedit_stats* stats);
/* i-th insertion string (length-major, A<C<G<T) starting with `first`; returns its length. */
int gpo_insertion_string(unsigned char first, int i, char out[6]);
/* scripts/goldpolish-ntedit:31-34: 1 when output_size/input_size (bc scale=4) < 0.75 */
int gpo_guard_rejects(uint64_t input_size, uint64_t output_size);

#ifdef __cplusplus
}
#endif
#endif
