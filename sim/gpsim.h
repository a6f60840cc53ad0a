/* Seeded synthetic-data generator for the GoldPolish hot path (tests / bench tooling). */
#ifndef GPSIM_H
#define GPSIM_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
  uint64_t seed;
  uint64_t genome_len;
  double repeat_frac;
  double contig_median, contig_sigma;
  uint32_t contig_min, contig_max;
  double tiny_contig_frac;
  double draft_err, draft_n_run_rate, draft_lower_rate, draft_iupac_rate;
  double coverage;
  double read_mean, read_sigma;
  uint32_t read_min, read_max;
  double read_err, read_sub, read_ins, read_n_rate;
  double phred_mean, phred_sd;
  uint32_t min_overlap;
  int32_t fastq;
} gpsim_params;

typedef struct {
  char* truth; uint64_t truth_len;
  size_t n_contigs; char* contig_seq; uint64_t contig_bases; uint64_t* contig_off;
  uint64_t* contig_tstart; uint64_t* contig_tend;
  size_t n_reads; char* read_seq; uint64_t read_bases; uint64_t* read_off;
  double* read_phred; uint8_t* read_qchar; uint8_t* read_qlast;
  size_t n_maps; uint32_t* map_read; uint32_t* map_contig; uint32_t* map_overlap;
  uint8_t* map_strand; uint32_t* map_tstart; uint32_t* map_tend; uint32_t* map_mx;
  int32_t fastq;
} gpsim_t;

void gpsim_default_params(gpsim_params* p);
gpsim_t* gpsim_generate(const gpsim_params* p);
void gpsim_free(gpsim_t* g);
void gpsim_contig_name(size_t i, char* out, size_t cap);
void gpsim_read_name(size_t i, char* out, size_t cap);
int gpsim_write_files(const gpsim_t* g, const char* draft_fa, const char* reads_path,
                      const char* paf_path, const char* ntlink_path);

#ifdef __cplusplus
}
#endif
#endif
