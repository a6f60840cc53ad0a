/* Seeded synthetic-data generator for the GoldPolish hot path.
 *
 * Tooling for tests and bench.py, not part of the product path: a truth genome, a draft cut
 * into golden-path-sized contigs with planted errors, simulated long reads with errors and
 * per-read quality, and the mappings (true origins) as PAF and ntLink-style triples.
 * Deterministic for a given (params, seed) on every platform: integer RNG, no libm in the
 * base-level decisions (libm only shapes lengths).
 */
#include "gpsim.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- rng: xoshiro256** seeded by splitmix64 ---- */
typedef struct { uint64_t s[4]; } rng_t;
static uint64_t splitmix(uint64_t* x) {
  uint64_t z = (*x += 0x9e3779b97f4a7c15ULL);
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
static void rng_seed(rng_t* r, uint64_t seed) { for (int i = 0; i < 4; i++) r->s[i] = splitmix(&seed); }
static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static inline uint64_t rng_next(rng_t* r) {
  uint64_t* s = r->s;
  const uint64_t result = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
  s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
  return result;
}
static inline double rng_unit(rng_t* r) { return (double)(rng_next(r) >> 11) * (1.0 / 9007199254740992.0); }
static inline uint64_t rng_below(rng_t* r, uint64_t n) { return (uint64_t)(rng_unit(r) * (double)n); }
static double rng_normal(rng_t* r) {
  double u1 = rng_unit(r), u2 = rng_unit(r);
  if (u1 < 1e-300) u1 = 1e-300;
  return sqrt(-2.0 * log(u1)) * cos(6.283185307179586 * u2);
}
static const char ACGT[4] = { 'A', 'C', 'G', 'T' };
static inline char rng_base(rng_t* r) { return ACGT[rng_next(r) >> 62]; }
static inline char other_base(rng_t* r, char c) {
  char b;
  do { b = rng_base(r); } while (b == c);
  return b;
}
static inline char comp(char c) {
  switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
               case 'a': return 't'; case 'c': return 'g'; case 'g': return 'c'; case 't': return 'a'; default: return c; }
}

/* growable byte buffer */
typedef struct { char* p; size_t n, cap; } buf_t;
static void buf_reserve(buf_t* b, size_t extra) {
  if (b->n + extra <= b->cap) return;
  size_t nc = b->cap ? b->cap : 1 << 20;
  while (nc < b->n + extra) nc *= 2;
  b->p = (char*)realloc(b->p, nc);
  b->cap = nc;
}
static inline void buf_push(buf_t* b, char c) { buf_reserve(b, 1); b->p[b->n++] = c; }

void gpsim_default_params(gpsim_params* p) {
  memset(p, 0, sizeof(*p));
  p->seed = 20250607ULL;
  p->genome_len = 200000;
  p->repeat_frac = 0.05;
  p->contig_median = 7000; p->contig_sigma = 0.8; p->contig_min = 2000; p->contig_max = 50000;
  p->tiny_contig_frac = 0.001;
  p->draft_err = 0.01;
  p->draft_n_run_rate = 0.05;       /* per contig */
  p->draft_lower_rate = 0.05;       /* per contig */
  p->draft_iupac_rate = 0.02;       /* per contig */
  p->coverage = 30.0;
  p->read_mean = 15000; p->read_sigma = 0.6; p->read_min = 1000; p->read_max = 100000;
  p->read_err = 0.05; p->read_sub = 0.3; p->read_ins = 0.3;
  p->read_n_rate = 0.0005;          /* per read base: N */
  p->phred_mean = 12.0; p->phred_sd = 3.0;
  p->min_overlap = 200;
  p->fastq = 1;
}

static uint32_t lognormal_len(rng_t* r, double median, double sigma, uint32_t lo, uint32_t hi) {
  double v = median * exp(sigma * rng_normal(r));
  if (v < lo) v = lo;
  if (v > hi) v = hi;
  return (uint32_t)v;
}

gpsim_t* gpsim_generate(const gpsim_params* p) {
  gpsim_t* g = (gpsim_t*)calloc(1, sizeof(gpsim_t));
  rng_t rg; rng_seed(&rg, p->seed);
  const uint64_t G = p->genome_len;

  /* truth genome with tandem / homopolymer repeats */
  char* truth = (char*)malloc(G + 1);
  for (uint64_t i = 0; i < G; i++) truth[i] = rng_base(&rg);
  {
    uint64_t budget = (uint64_t)(p->repeat_frac * (double)G), used = 0;
    while (used < budget) {
      uint32_t unit = 1 + (uint32_t)rng_below(&rg, 6);           /* 1..6 bp unit */
      uint32_t len = 12 + (uint32_t)rng_below(&rg, 60);          /* 12..71 bp tract */
      if (G <= len + 1) break;
      uint64_t at = rng_below(&rg, G - len);
      for (uint32_t j = unit; j < len; j++) truth[at + j] = truth[at + (j % unit)];
      used += len;
    }
  }
  truth[G] = 0;

  /* cut into contigs, plant draft errors */
  buf_t cs = { 0 };
  size_t ccap = 1024;
  g->contig_off = (uint64_t*)malloc((ccap + 1) * sizeof(uint64_t));
  g->contig_tstart = (uint64_t*)malloc(ccap * sizeof(uint64_t));
  g->contig_tend = (uint64_t*)malloc(ccap * sizeof(uint64_t));
  uint64_t pos = 0; size_t nc = 0;
  while (pos < G) {
    uint32_t L;
    if (rng_unit(&rg) < p->tiny_contig_frac) L = 40 + (uint32_t)rng_below(&rg, 59); /* < 100 bp: dropped by ntEdit */
    else L = lognormal_len(&rg, p->contig_median, p->contig_sigma, p->contig_min, p->contig_max);
    if (pos + L > G) L = (uint32_t)(G - pos);
    if (nc + 1 >= ccap) {
      ccap *= 2;
      g->contig_off = (uint64_t*)realloc(g->contig_off, (ccap + 1) * sizeof(uint64_t));
      g->contig_tstart = (uint64_t*)realloc(g->contig_tstart, ccap * sizeof(uint64_t));
      g->contig_tend = (uint64_t*)realloc(g->contig_tend, ccap * sizeof(uint64_t));
    }
    g->contig_off[nc] = cs.n; g->contig_tstart[nc] = pos; g->contig_tend[nc] = pos + L;
    size_t start_n = cs.n;
    for (uint64_t i = pos; i < pos + L; i++) {
      char c = truth[i];
      if (rng_unit(&rg) < p->draft_err) {
        uint32_t kind = (uint32_t)rng_below(&rg, 3);
        uint32_t n = 1 + (uint32_t)rng_below(&rg, 3);
        if (kind == 0) { buf_push(&cs, other_base(&rg, c)); }
        else if (kind == 1) { buf_push(&cs, c); for (uint32_t j = 0; j < n; j++) buf_push(&cs, rng_base(&rg)); }
        else { i += n - 1; } /* deletion of n truth bases (including this one) */
      } else buf_push(&cs, c);
    }
    size_t clen = cs.n - start_n;
    if (clen >= 200) {
      if (rng_unit(&rg) < p->draft_n_run_rate) {
        uint32_t n = 1 + (uint32_t)rng_below(&rg, 50);
        size_t at = start_n + rng_below(&rg, clen - n);
        for (uint32_t j = 0; j < n; j++) cs.p[at + j] = 'N';
      }
      if (rng_unit(&rg) < p->draft_lower_rate) {
        uint32_t n = 1 + (uint32_t)rng_below(&rg, 150);
        size_t at = start_n + rng_below(&rg, clen - n);
        for (uint32_t j = 0; j < n; j++) { char c = cs.p[at + j]; if (c >= 'A' && c <= 'Z') cs.p[at + j] = (char)(c + 32); }
      }
      if (rng_unit(&rg) < p->draft_iupac_rate) {
        static const char IU[] = "RYSWKMBDHV";
        size_t at = start_n + rng_below(&rg, clen);
        cs.p[at] = IU[rng_below(&rg, 10)];
      }
    }
    pos += L; nc++;
  }
  g->contig_off[nc] = cs.n;
  g->n_contigs = nc; g->contig_seq = cs.p; g->contig_bases = cs.n;

  /* reads */
  buf_t rs = { 0 };
  size_t rcap = 1024, mcap = 4096;
  g->read_off = (uint64_t*)malloc((rcap + 1) * sizeof(uint64_t));
  g->read_phred = (double*)malloc(rcap * sizeof(double));
  g->read_qchar = (uint8_t*)malloc(rcap);
  g->read_qlast = (uint8_t*)malloc(rcap);
  g->map_read = (uint32_t*)malloc(mcap * sizeof(uint32_t));
  g->map_contig = (uint32_t*)malloc(mcap * sizeof(uint32_t));
  g->map_overlap = (uint32_t*)malloc(mcap * sizeof(uint32_t));
  g->map_strand = (uint8_t*)malloc(mcap);
  g->map_tstart = (uint32_t*)malloc(mcap * sizeof(uint32_t));
  g->map_tend = (uint32_t*)malloc(mcap * sizeof(uint32_t));
  size_t nr = 0, nm = 0;
  const double target_bases = p->coverage * (double)G;
  /* "mean" of the lognormal: median = mean / exp(sigma^2/2) */
  const double rmedian = p->read_mean / exp(0.5 * p->read_sigma * p->read_sigma);
  double bases = 0;
  char* tmp = NULL; size_t tmpcap = 0;
  size_t first_contig_hint = 0;
  (void)first_contig_hint;
  while (bases < target_bases) {
    uint32_t L = lognormal_len(&rg, rmedian, p->read_sigma, p->read_min, p->read_max);
    if (L > G) L = (uint32_t)G;
    uint64_t a = rng_below(&rg, G - L + 1), b = a + L;
    int rev = (int)(rng_next(&rg) >> 63);
    if (nr + 1 >= rcap) {
      rcap *= 2;
      g->read_off = (uint64_t*)realloc(g->read_off, (rcap + 1) * sizeof(uint64_t));
      g->read_phred = (double*)realloc(g->read_phred, rcap * sizeof(double));
      g->read_qchar = (uint8_t*)realloc(g->read_qchar, rcap);
      g->read_qlast = (uint8_t*)realloc(g->read_qlast, rcap);
    }
    /* mutate forward-strand copy into tmp */
    if (tmpcap < (size_t)L * 2 + 64) { tmpcap = (size_t)L * 2 + 64; tmp = (char*)realloc(tmp, tmpcap); }
    size_t tn = 0;
    for (uint64_t i = a; i < b; i++) {
      char c = truth[i];
      if (tn + 8 >= tmpcap) { tmpcap *= 2; tmp = (char*)realloc(tmp, tmpcap); }
      double u = rng_unit(&rg);
      if (u < p->read_err) {
        double v = u / p->read_err;
        if (v < p->read_sub) tmp[tn++] = other_base(&rg, c);
        else if (v < p->read_sub + p->read_ins) { tmp[tn++] = c; tmp[tn++] = rng_base(&rg); }
        else { /* deletion */ }
      } else if (u > 1.0 - p->read_n_rate) tmp[tn++] = 'N';
      else tmp[tn++] = c;
    }
    g->read_off[nr] = rs.n;
    buf_reserve(&rs, tn);
    if (rev) for (size_t i = 0; i < tn; i++) rs.p[rs.n + i] = comp(tmp[tn - 1 - i]);
    else memcpy(rs.p + rs.n, tmp, tn);
    rs.n += tn;
    /* quality: one repeated character per read plus a different last character, so that
       calc_phred_avg(qual, 0, len-1) (seqindex.cpp:45 ignores the last char) is exercised */
    double q = p->phred_mean + p->phred_sd * rng_normal(&rg);
    if (q < 2) q = 2;
    if (q > 40) q = 40;
    uint8_t qc = (uint8_t)(33 + (int)(q + 0.5));
    g->read_qchar[nr] = qc;
    g->read_qlast[nr] = (uint8_t)(33 + (int)rng_below(&rg, 41));
    g->read_phred[nr] = (double)(qc - 33);
    /* mappings from the true origin: binary search first overlapping contig */
    size_t lo = 0, hi = nc;
    while (lo < hi) { size_t mid = (lo + hi) / 2; if (g->contig_tend[mid] <= a) lo = mid + 1; else hi = mid; }
    for (size_t c = lo; c < nc && g->contig_tstart[c] < b; c++) {
      uint64_t os = a > g->contig_tstart[c] ? a : g->contig_tstart[c];
      uint64_t oe = b < g->contig_tend[c] ? b : g->contig_tend[c];
      if (oe <= os || oe - os < p->min_overlap) continue;
      if (nm + 1 >= mcap) {
        mcap *= 2;
        g->map_read = (uint32_t*)realloc(g->map_read, mcap * sizeof(uint32_t));
        g->map_contig = (uint32_t*)realloc(g->map_contig, mcap * sizeof(uint32_t));
        g->map_overlap = (uint32_t*)realloc(g->map_overlap, mcap * sizeof(uint32_t));
        g->map_strand = (uint8_t*)realloc(g->map_strand, mcap);
        g->map_tstart = (uint32_t*)realloc(g->map_tstart, mcap * sizeof(uint32_t));
        g->map_tend = (uint32_t*)realloc(g->map_tend, mcap * sizeof(uint32_t));
      }
      g->map_read[nm] = (uint32_t)nr; g->map_contig[nm] = (uint32_t)c;
      g->map_overlap[nm] = (uint32_t)(oe - os); g->map_strand[nm] = (uint8_t)rev;
      g->map_tstart[nm] = (uint32_t)(os - g->contig_tstart[c]);
      g->map_tend[nm] = (uint32_t)(oe - g->contig_tstart[c]);
      nm++;
    }
    bases += (double)tn; nr++;
  }
  free(tmp);
  g->read_off[nr] = rs.n;
  g->n_reads = nr; g->read_seq = rs.p; g->read_bases = rs.n; g->n_maps = nm;
  g->truth = truth; g->truth_len = G;
  g->fastq = p->fastq;
  /* minimizer counts for the ntLink-style triples: ~ overlap/1000 + 1 with jitter */
  g->map_mx = (uint32_t*)malloc((nm ? nm : 1) * sizeof(uint32_t));
  for (size_t i = 0; i < nm; i++) g->map_mx[i] = 1 + g->map_overlap[i] / 1000 + (uint32_t)rng_below(&rg, 3);
  return g;
}

void gpsim_free(gpsim_t* g) {
  if (!g) return;
  free(g->truth); free(g->contig_seq); free(g->contig_off); free(g->contig_tstart); free(g->contig_tend);
  free(g->read_seq); free(g->read_off); free(g->read_phred); free(g->read_qchar); free(g->read_qlast);
  free(g->map_read); free(g->map_contig); free(g->map_overlap); free(g->map_strand);
  free(g->map_tstart); free(g->map_tend); free(g->map_mx);
  free(g);
}

/* names: deliberately not zero padded so that lexicographic order != numeric order */
void gpsim_contig_name(size_t i, char* out, size_t cap) { snprintf(out, cap, "ctg%zu", i); }
void gpsim_read_name(size_t i, char* out, size_t cap) { snprintf(out, cap, "read%zu", i); }

int gpsim_write_files(const gpsim_t* g, const char* draft_fa, const char* reads_path, const char* paf_path, const char* ntlink_path) {
  char name[64], cname[64];
  if (draft_fa) {
    FILE* f = fopen(draft_fa, "w"); if (!f) return -1;
    for (size_t i = 0; i < g->n_contigs; i++) {
      gpsim_contig_name(i, name, sizeof name);
      fprintf(f, ">%s\n", name);
      fwrite(g->contig_seq + g->contig_off[i], 1, g->contig_off[i + 1] - g->contig_off[i], f);
      fputc('\n', f);
    }
    fclose(f);
  }
  if (reads_path) {
    FILE* f = fopen(reads_path, "w"); if (!f) return -1;
    char* q = NULL; size_t qcap = 0;
    for (size_t i = 0; i < g->n_reads; i++) {
      size_t len = g->read_off[i + 1] - g->read_off[i];
      gpsim_read_name(i, name, sizeof name);
      if (g->fastq) {
        fprintf(f, "@%s len=%zu\n", name, len);
        fwrite(g->read_seq + g->read_off[i], 1, len, f);
        fputs("\n+\n", f);
        if (qcap < len + 1) { qcap = len + 1; q = (char*)realloc(q, qcap); }
        memset(q, g->read_qchar[i], len);
        if (len) q[len - 1] = (char)g->read_qlast[i];
        fwrite(q, 1, len, f);
        fputc('\n', f);
      } else {
        fprintf(f, ">%s\n", name);
        fwrite(g->read_seq + g->read_off[i], 1, len, f);
        fputc('\n', f);
      }
    }
    free(q);
    fclose(f);
  }
  if (paf_path) {
    FILE* f = fopen(paf_path, "w"); if (!f) return -1;
    for (size_t m = 0; m < g->n_maps; m++) {
      size_t r = g->map_read[m], c = g->map_contig[m];
      size_t rlen = g->read_off[r + 1] - g->read_off[r], clen = g->contig_off[c + 1] - g->contig_off[c];
      gpsim_read_name(r, name, sizeof name); gpsim_contig_name(c, cname, sizeof cname);
      fprintf(f, "%s\t%zu\t%u\t%zu\t%c\t%s\t%zu\t%u\t%u\t%u\t%u\t60\n", name, rlen, 0u, rlen,
              g->map_strand[m] ? '-' : '+', cname, clen, g->map_tstart[m], g->map_tend[m],
              (unsigned)(g->map_overlap[m] * 0.9), g->map_overlap[m]);
    }
    fclose(f);
  }
  if (ntlink_path) {
    FILE* f = fopen(ntlink_path, "w"); if (!f) return -1;
    for (size_t m = 0; m < g->n_maps; m++) {
      gpsim_read_name(g->map_read[m], name, sizeof name); gpsim_contig_name(g->map_contig[m], cname, sizeof cname);
      fprintf(f, "%s %s %u\n", name, cname, g->map_mx[m]);
    }
    fclose(f);
  }
  return 0;
}
