// synth_core.h — deterministic, integer-only synthetic all-vs-all alignment
// generator (SURVEY.md §8d `gen_synth`). The same inline code runs on the host
// (OpenMP) and, being free of floating point, can run unchanged on the device.
//
// G genomes x C contigs of length L, named g{i}#1#c{j} (PanSN). For every
// ordered genome pair i != j and contig c: A collinear alignments tiling the
// contig. CIGAR = alternating '=' runs (long-tailed, mean ~eq_mean) and events
// {X len 1: 80 %, I: 10 %, D: 10 %, indel length 1..50, mean ~3.5}; first and
// last run are '='. Strand '-' with probability rev_permille / 1000.
#pragma once
#include <stdint.h>

#include "../../include/impgx.h"

#if defined(__CUDACC__)
#define SYNTH_HD __host__ __device__ __forceinline__
#else
#define SYNTH_HD inline
#endif

typedef struct impgx_synth_cfg {
  uint32_t genomes;      /* G */
  uint32_t contigs;      /* C */
  uint32_t contig_len;   /* L */
  uint32_t tiles;        /* A alignments per (pair, contig) */
  uint32_t eq_mean;      /* mean '=' run length = 1 / divergence */
  uint32_t rev_permille; /* probability of '-' strand, in 1/1000 */
  uint64_t seed;
  uint32_t partners;     /* 0: all-vs-all; k > 0: genome i is aligned (as query) against k others only,
                            j = (i + 1 + m * ((G - 1) / k)) mod G for m < k (sparsified pairs) */
  uint32_t reserved;
} impgx_synth_cfg;

struct SynthRng {
  uint64_t s;
  SYNTH_HD uint64_t next() {
    uint64_t z = (s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
  }
};

SYNTH_HD uint32_t synth_clz32(uint32_t v) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__clz((int)v);
#else
  return v ? (uint32_t)__builtin_clz(v) : 32u;
#endif
}

SYNTH_HD uint32_t synth_partners(const impgx_synth_cfg &c) {
  return c.partners && c.partners < c.genomes - 1 ? c.partners : c.genomes - 1;
}
SYNTH_HD uint64_t synth_num_alignments(const impgx_synth_cfg &c) {
  return (uint64_t)c.genomes * synth_partners(c) * c.contigs * c.tiles;
}

// Walks alignment n. With runs == nullptr only counts. Returns the run count
// and fills rec (coordinates derive from the CIGAR, so they are consistent).
SYNTH_HD uint32_t synth_alignment(const impgx_synth_cfg &c, uint64_t n, impgx_record *rec, uint32_t *runs) {
  const uint32_t A = c.tiles, C = c.contigs, G = c.genomes;
  const uint32_t k = (uint32_t)(n % A);
  const uint32_t ctg = (uint32_t)((n / A) % C);
  const uint64_t p = n / ((uint64_t)A * C);
  const uint32_t K = synth_partners(c);
  const uint32_t i = (uint32_t)(p / K);
  const uint32_t jj = (uint32_t)(p % K);
  // all-vs-all: every other genome in ascending order; sparsified: K partners spread evenly around the circle
  const uint32_t j = K == G - 1 ? (jj < i ? jj : jj + 1) : (uint32_t)(((uint64_t)i + 1 + (uint64_t)jj * ((G - 1) / K)) % G);
  const uint32_t T = c.contig_len / A;
  uint32_t margin = T / 32;
  if (margin < 64) margin = 64;
  SynthRng rng{c.seed ^ (n * 0xD1B54A32D192ED03ull + 0x2545F4914F6CDD1Dull)};
  const uint64_t h = rng.next();
  const uint32_t t_start = k * T + (uint32_t)(h % (margin / 2));
  const uint32_t q_start = k * T + (uint32_t)((h >> 32) % (margin / 4));
  const uint32_t strand = (rng.next() % 1000u) < c.rev_permille ? 1u : 0u;
  const uint32_t tspan = T - margin;
  uint32_t ins_budget = margin / 4;
  uint32_t m = (c.eq_mean * 2u) / 3u;
  if (m < 1) m = 1;

  uint32_t t_rem = tspan, q_len = 0, nr = 0;
  for (;;) {
    uint64_t r = rng.next();
    uint32_t lz = synth_clz32((uint32_t)r);
    if (lz > 31) lz = 31;
    uint32_t len = 1 + lz * m + (uint32_t)((r >> 32) % m);
    uint64_t e = rng.next();
    uint32_t kind = (uint32_t)(e % 10u);  // 0..7 X, 8 I, 9 D
    uint32_t lz2 = synth_clz32((uint32_t)(e >> 32));
    uint32_t il = 1 + lz2 * 2 + (uint32_t)((e >> 8) & 1u);
    if (il > 50) il = 50;
    uint32_t op, ol, ev_t, ev_q;
    if (kind < 8) {
      op = IMPGX_OP_X; ol = 1; ev_t = 1; ev_q = 1;
    } else if (kind == 8 && il <= ins_budget) {
      op = IMPGX_OP_I; ol = il; ev_t = 0; ev_q = il;
    } else if (kind == 9) {
      op = IMPGX_OP_D; ol = il; ev_t = il; ev_q = 0;
    } else {
      op = IMPGX_OP_X; ol = 1; ev_t = 1; ev_q = 1;
    }
    if ((uint64_t)len + ev_t + 1 > t_rem) {
      if (runs) runs[nr] = IMPGX_RUN(IMPGX_OP_EQ, t_rem);
      nr++;
      q_len += t_rem;
      break;
    }
    if (runs) {
      runs[nr] = IMPGX_RUN(IMPGX_OP_EQ, len);
      runs[nr + 1] = IMPGX_RUN(op, ol);
    }
    nr += 2;
    t_rem -= len + ev_t;
    q_len += len + ev_q;
    if (op == IMPGX_OP_I) ins_budget -= ol;
  }
  if (rec) {
    rec->query_id = i * C + ctg;
    rec->target_id = j * C + ctg;
    rec->query_start = (int32_t)q_start;
    rec->query_end = (int32_t)(q_start + q_len);
    rec->target_start = (int32_t)t_start;
    rec->target_end = (int32_t)(t_start + tspan);
    rec->strand = strand;
    rec->reserved = 0;
  }
  return nr;
}

// BED row k: uniform sequence, length U[min_len, max_len], uniform start.
SYNTH_HD void synth_bed_row(const impgx_synth_cfg &c, uint64_t seed, uint64_t k, uint32_t min_len, uint32_t max_len,
                            impgx_range *out) {
  SynthRng rng{seed ^ (k * 0xA24BAED4963EE407ull + 0x9FB21C651E98DF25ull)};
  uint32_t n_seqs = c.genomes * c.contigs;
  uint32_t seq = (uint32_t)(rng.next() % n_seqs);
  uint32_t len = min_len + (uint32_t)(rng.next() % (uint64_t)(max_len - min_len + 1));
  if (len > c.contig_len) len = c.contig_len;
  uint32_t start = (uint32_t)(rng.next() % (uint64_t)(c.contig_len - len + 1));
  out->target_id = seq;
  out->start = (int32_t)start;
  out->end = (int32_t)(start + len);
}
