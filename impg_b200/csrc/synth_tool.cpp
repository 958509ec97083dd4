// synth_tool.cpp — the synthetic workload generator of bench.py and the tests (SURVEY.md 8d `gen_synth`) and the
// CIGAR-text writer of the CPU reference arm. TEST / BENCH TOOLING: it is built into its own shared library
// (libimpgx_synth.so) so that the product library holds no generator code and the reference arm of bench.py never
// loads the product. Plain C++ (OpenMP), no CUDA.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <new>
#include <stdexcept>
#include <string>
#include <vector>

#include "synth_core.h"

namespace {
thread_local std::string g_err;
struct Fail : std::runtime_error {
  int code;
  Fail(int c, const std::string &m) : std::runtime_error(m), code(c) {}
};
}  // namespace
#define REQUIRE(cond, code, msg) \
  do {                           \
    if (!(cond)) throw Fail((code), (msg)); \
  } while (0)
#define API_BEGIN try {
#define API_END                      \
  }                                  \
  catch (const Fail &e) {            \
    g_err = e.what();                \
    return e.code;                   \
  }                                  \
  catch (const std::bad_alloc &) {   \
    g_err = "host allocation failed"; \
    return IMPGX_E_NOMEM;            \
  }                                  \
  return IMPGX_OK;

extern "C" {

const char *impgx_synth_last_error(void) { return g_err.c_str(); }

uint64_t impgx_synth_num_alignments(const impgx_synth_cfg *c) { return synth_num_alignments(*c); }

// records + per-alignment run counts for alignments [first, first+count)
int impgx_synth_records(const impgx_synth_cfg *c, uint64_t first, uint64_t count, impgx_record *recs, uint32_t *n_runs) {
  API_BEGIN
  REQUIRE(c && recs && n_runs, IMPGX_E_INVALID, "NULL argument");
  REQUIRE(c->genomes >= 2 && c->contigs >= 1 && c->tiles >= 1 && c->contig_len / c->tiles >= 256, IMPGX_E_INVALID,
          "bad synthetic config");
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)count; i++) n_runs[i] = synth_alignment(*c, first + (uint64_t)i, &recs[i], nullptr);
  API_END
}
int impgx_synth_runs(const impgx_synth_cfg *c, uint64_t first, uint64_t count, const uint64_t *run_offsets, uint32_t *runs) {
  API_BEGIN
  REQUIRE(c && run_offsets && runs, IMPGX_E_INVALID, "NULL argument");
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)count; i++)
    synth_alignment(*c, first + (uint64_t)i, nullptr, runs + (run_offsets[i] - run_offsets[0]));
  API_END
}
// runs of the alignments ids[0..count) only (a shard generates just what it walks)
int impgx_synth_runs_subset(const impgx_synth_cfg *c, const uint64_t *ids, uint64_t count, const uint64_t *run_offsets,
                            uint32_t *runs) {
  API_BEGIN
  REQUIRE(c && ids && run_offsets && runs, IMPGX_E_INVALID, "NULL argument");
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)count; i++) synth_alignment(*c, ids[i], nullptr, runs + run_offsets[i]);
  API_END
}
int impgx_synth_bed(const impgx_synth_cfg *c, uint64_t seed, uint64_t n_rows, uint32_t min_len, uint32_t max_len,
                    impgx_range *out) {
  API_BEGIN
  REQUIRE(c && out && min_len >= 1 && max_len >= min_len, IMPGX_E_INVALID, "bad argument");
  for (uint64_t k = 0; k < n_rows; k++) synth_bed_row(*c, seed, k, min_len, max_len, &out[k]);
  API_END
}

// CIGAR text of a run stream, for the reference-cost CPU baseline (it preads
// and parses text per hit): writes the concatenated CIGAR strings to `path`
// and returns per-alignment byte offsets and lengths.
int impgx_write_cigar_text(const uint32_t *runs, const uint64_t *run_offsets, uint64_t n, const char *path,
                           uint64_t *offsets, uint64_t *lens) {
  API_BEGIN
  REQUIRE(runs && run_offsets && path && offsets && lens, IMPGX_E_INVALID, "NULL argument");
  static const char OPS[] = "=XIDM";
  auto digits = [](uint32_t v) {
    int d = 1;
    while (v >= 10) {
      v /= 10;
      d++;
    }
    return d;
  };
#pragma omp parallel for schedule(static)
  for (long long i = 0; i < (long long)n; i++) {
    uint64_t b = 0;
    for (uint64_t k = run_offsets[i]; k < run_offsets[i + 1]; k++) b += (uint64_t)digits(IMPGX_RUN_LEN(runs[k])) + 1;
    lens[i] = b;
  }
  uint64_t total = 0;
  for (uint64_t i = 0; i < n; i++) {
    offsets[i] = total;
    total += lens[i];
  }
  FILE *f = fopen(path, "wb");
  REQUIRE(f, IMPGX_E_IO, std::string("cannot create '") + path + "'");
  const uint64_t CH = 1ull << 16;  // alignments per chunk
  std::vector<char> buf;
  for (uint64_t a = 0; a < n; a += CH) {
    uint64_t b = std::min(n, a + CH);
    uint64_t bytes = (b < n ? offsets[b] : total) - offsets[a];
    buf.resize(bytes);
#pragma omp parallel for schedule(static)
    for (long long i = (long long)a; i < (long long)b; i++) {
      char *p = buf.data() + (offsets[i] - offsets[a]);
      for (uint64_t k = run_offsets[i]; k < run_offsets[i + 1]; k++) {
        uint32_t len = IMPGX_RUN_LEN(runs[k]);
        int dg = digits(len);
        for (int j = dg - 1; j >= 0; j--) {
          p[j] = (char)('0' + len % 10);
          len /= 10;
        }
        p += dg;
        *p++ = OPS[IMPGX_RUN_OP(runs[k])];
      }
    }
    if (bytes && fwrite(buf.data(), 1, bytes, f) != bytes) {
      fclose(f);
      throw Fail(IMPGX_E_IO, std::string("short write to '") + path + "'");
    }
  }
  fclose(f);
  API_END
}

}  // extern "C"
