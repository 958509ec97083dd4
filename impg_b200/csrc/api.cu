// api.cu — the extern "C" boundary (include/impgx.h). No exception crosses it.
#include <exception>
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include <cmath>
#include <cstring>
#include <fstream>
#include <set>
#include <sstream>

#include "engine.cuh"
#include "comm.cuh"

using namespace impgx;

namespace impgx {
void set_last_error(const std::string &m);
const char *last_error();
std::string format_rows(const impgx_index *idx, const impgx_results *res, size_t row, const char *name, int32_t d,
                        int format);

// ---- host text helpers (reference parsers / writers for this path)

// parse_cigar_to_delta (src/impg.rs:2935-2950): digits accumulate, any other
// byte ends a run; an unknown op is an error instead of the reference's panic.
long parse_cigar(const char *s, size_t n, std::vector<uint32_t> &out) {
  int64_t len = 0;
  long cnt = 0;
  for (size_t i = 0; i < n; i++) {
    unsigned char c = (unsigned char)s[i];
    if (c >= '0' && c <= '9') {
      len = len * 10 + (c - '0');
      if (len >= (1ll << 29)) return -2;
    } else {
      uint32_t op;
      switch (c) {
        case '=': op = IMPGX_OP_EQ; break;
        case 'X': op = IMPGX_OP_X; break;
        case 'I': op = IMPGX_OP_I; break;
        case 'D': op = IMPGX_OP_D; break;
        case 'M': op = IMPGX_OP_M; break;
        default: return -1;
      }
      out.push_back(IMPGX_RUN(op, (uint32_t)len));
      len = 0;
      cnt++;
    }
  }
  return cnt;
}

static bool parse_usize(const char *b, const char *e, uint64_t &v) {
  if (b < e && *b == '+') b++;
  if (b >= e) return false;
  v = 0;
  for (; b < e; b++) {
    if (*b < '0' || *b > '9') return false;
    const uint64_t dgt = (uint64_t)(*b - '0');
    if (v > (UINT64_MAX - dgt) / 10) return false;  // Rust's parse::<usize> fails on overflow, it never wraps
    v = v * 10 + dgt;
  }
  return true;
}

// The bytes of a PAF file: a read-only mapping of a plain file, or the inflated text of a gzip / BGZF one.
struct PafText {
  const char *p = nullptr;
  size_t n = 0;
  void *map = nullptr;
  size_t map_len = 0;
  std::string owned;
  PafText() {}
  PafText(const PafText &) = delete;
  PafText &operator=(const PafText &) = delete;
  ~PafText() {
    if (map) munmap(map, map_len);
  }
  const char *data() const { return p; }
  size_t size() const { return n; }
  char operator[](size_t i) const { return p[i]; }
  size_t find(char c, size_t pos) const {
    if (pos >= n) return std::string::npos;
    const void *q = memchr(p + pos, c, n - pos);
    return q ? (size_t)((const char *)q - p) : std::string::npos;
  }
};

// The same parse with the lines spread over the host cores. Per line: fields, coordinates, the CIGAR decoded into
// the thread's own run buffer (threads own contiguous blocks of lines, so the buffers concatenate in file order).
// What depends on file order is done in two cheap sequential passes: the reference's byte offsets (a prefix sum over
// line lengths) and the sequence ids by first appearance (query column first). The first bad line in file order
// reports its error, exactly as the serial loop would.
static void parse_paf_parallel(const PafText &data, const std::string &path, PafData &out) {
  std::vector<size_t> starts;
  for (size_t pos = 0; pos < data.size();) {
    starts.push_back(pos);
    const void *nl = memchr(data.data() + pos, '\n', data.size() - pos);
    pos = nl ? (size_t)((const char *)nl - data.data()) + 1 : data.size();
  }
  const size_t n_lines = starts.size();
  struct Line {
    impgx_record r;
    const char *qn_b, *qn_e, *tn_b, *tn_e;
    uint64_t ql, tl, cg_rel, cg_len, ref_len;
    uint32_t n_runs;
    const char *err;  // nullptr = fine
  };
  std::vector<Line> lines(n_lines);
  int n_threads = 1;
#ifdef _OPENMP
  n_threads = std::max(1, omp_get_max_threads());
#endif
  n_threads = (int)std::min<size_t>((size_t)n_threads, std::max<size_t>(n_lines, 1));
  std::vector<std::vector<uint32_t>> runs_of((size_t)n_threads);
  std::vector<size_t> first_line((size_t)n_threads + 1);
  for (int t = 0; t <= n_threads; t++) first_line[t] = n_lines * (size_t)t / (size_t)n_threads;
  // an exception (std::bad_alloc of a run buffer) must not leave the OpenMP region: it is kept and rethrown after it
  std::vector<std::exception_ptr> thread_err((size_t)n_threads);
#pragma omp parallel for schedule(static, 1) num_threads(n_threads)
  for (int t = 0; t < n_threads; t++) try {
    std::vector<uint32_t> &my_runs = runs_of[t];
    {
      // a run takes at least two bytes of text: reserving once spares the growth copies (untouched pages cost nothing)
      const size_t b0 = first_line[t] < n_lines ? starts[first_line[t]] : data.size();
      const size_t b1 = first_line[t + 1] < n_lines ? starts[first_line[t + 1]] : data.size();
      my_runs.reserve((b1 - b0) / 2 + 16);
    }
    std::vector<std::pair<const char *, const char *>> fld;
    for (size_t i = first_line[t]; i < first_line[t + 1]; i++) {
      Line &ln = lines[i];
      ln.err = nullptr;
      ln.n_runs = 0;
      const size_t pos = starts[i];
      size_t eol = (i + 1 < n_lines) ? starts[i + 1] - 1 : data.size();
      if (i + 1 == n_lines && eol > pos && data[eol - 1] == '\n') eol--;  // last line with its terminator
      size_t end = eol;
      if (end > pos && data[end - 1] == '\r') end--;
      ln.ref_len = (uint64_t)(end - pos) + 1;
      fld.clear();
      {
        const char *b = data.data() + pos, *e = data.data() + end, *p = b;
        for (;;) {
          const char *tb = (const char *)memchr(p, '\t', (size_t)(e - p));
          if (!tb) {
            fld.emplace_back(p, e);
            break;
          }
          fld.emplace_back(p, tb);
          p = tb + 1;
        }
      }
      if (fld.size() < 12) {
        ln.err = "Not enough fields in PAF record";
        continue;
      }
      uint64_t qs, qe, ts, te;
      if (!(parse_usize(fld[1].first, fld[1].second, ln.ql) && parse_usize(fld[2].first, fld[2].second, qs) &&
            parse_usize(fld[3].first, fld[3].second, qe) && parse_usize(fld[6].first, fld[6].second, ln.tl) &&
            parse_usize(fld[7].first, fld[7].second, ts) && parse_usize(fld[8].first, fld[8].second, te))) {
        ln.err = "Invalid field";
        continue;
      }
      if (!(fld[4].first < fld[4].second)) {
        ln.err = "Expected '+' or '-' for strand";
        continue;
      }
      const char sc = *fld[4].first;
      if (sc != '+' && sc != '-') {
        ln.err = "Invalid strand";
        continue;
      }
      if (!(ln.ql <= INT32_MAX && ln.tl <= INT32_MAX && qe <= INT32_MAX && te <= INT32_MAX && qs <= INT32_MAX &&
            ts <= INT32_MAX)) {
        ln.err = "coordinate beyond 2^31-1";
        continue;
      }
      ln.qn_b = fld[0].first; ln.qn_e = fld[0].second;
      ln.tn_b = fld[5].first; ln.tn_e = fld[5].second;
      ln.r.query_start = (int32_t)qs; ln.r.query_end = (int32_t)qe;
      ln.r.target_start = (int32_t)ts; ln.r.target_end = (int32_t)te;
      ln.r.strand = sc == '-' ? 1 : 0;
      ln.r.reserved = 0;
      bool have = false;
      uint64_t rel = 0;
      ln.cg_len = 0;
      for (auto &tg : fld) {
        if (tg.second - tg.first >= 5 && memcmp(tg.first, "cg:Z:", 5) == 0) {
          const size_t before = my_runs.size();
          const long k = parse_cigar(tg.first + 5, (size_t)(tg.second - tg.first - 5), my_runs);
          if (k < 0) {
            my_runs.resize(before);
            ln.err = "Invalid CIGAR operation";
            break;
          }
          have = k > 0;
          ln.n_runs = (uint32_t)k;
          rel += 5;
          ln.cg_len = (uint64_t)(tg.second - tg.first - 5);
          break;
        }
        rel += (uint64_t)(tg.second - tg.first) + 1;
      }
      ln.cg_rel = rel;
      if (!ln.err && !have) ln.err = "The alignment file does not contain CIGAR strings ('cg:Z' tag)";
    }
  } catch (...) {
    thread_err[t] = std::current_exception();
  }
  for (auto &ep : thread_err)
    if (ep) std::rethrow_exception(ep);
  for (size_t i = 0; i < n_lines; i++)
    if (lines[i].err)
      throw Error(IMPGX_E_PARSE, std::string(lines[i].err) + " (line " + std::to_string(i + 1) + " of '" + path + "')");
  // sequential passes: ids by first appearance, the reference's byte offsets
  auto get_id = [&](const char *b, const char *e, uint64_t len) {
    std::string name(b, e);
    auto it = out.ids.find(name);
    if (it != out.ids.end()) return it->second;
    uint32_t id = (uint32_t)out.names.size();
    out.ids.emplace(name, id);
    out.names.push_back(name);
    out.lens.push_back(len);
    return id;
  };
  out.recs.reserve(out.recs.size() + n_lines);
  uint64_t ref_pos = 0, run_total = out.run_off.back();
  for (size_t i = 0; i < n_lines; i++) {
    Line &ln = lines[i];
    ln.r.query_id = get_id(ln.qn_b, ln.qn_e, ln.ql);
    ln.r.target_id = get_id(ln.tn_b, ln.tn_e, ln.tl);
    out.recs.push_back(ln.r);
    run_total += ln.n_runs;
    out.run_off.push_back(run_total);
    out.cg_off.push_back(ref_pos + ln.cg_rel);
    out.cg_len.push_back(ln.cg_len);
    out.file_idx.push_back(out.n_files);
    ref_pos += ln.ref_len;
  }
  size_t add = 0;
  for (auto &v : runs_of) add += v.size();
  out.runs.reserve(out.runs.size() + add);
  for (auto &v : runs_of) {  // thread blocks are in file order
    out.runs.insert(out.runs.end(), v.begin(), v.end());
    std::vector<uint32_t>().swap(v);
  }
  out.n_files++;
}

// parse_paf_line / parse_paf (src/paf.rs:118-194) with SequenceIndex ids by
// first appearance (src/seqidx.rs:22-35). The CIGAR is decoded here, once.
void parse_paf(const std::string &path, PafData &out) {
  // gzopen reads plain text transparently and inflates gzip / BGZF (a BGZF file
  // is a series of gzip members, src/paf.rs:199-302); the CIGARs are decoded
  // here once, so no virtual offsets need to be kept.
  PafText data;
  {
    // a plain file is mapped; only gzip / BGZF input goes through zlib
    bool mapped = false;
    const int fd = open(path.c_str(), O_RDONLY);
    REQUIRE(fd >= 0, IMPGX_E_IO, "cannot open PAF file '" + path + "'");
    unsigned char magic[2] = {0, 0};
    const ssize_t got = pread(fd, magic, 2, 0);
    struct stat st;
    if (!(got == 2 && magic[0] == 0x1f && magic[1] == 0x8b) && fstat(fd, &st) == 0 && S_ISREG(st.st_mode)) {
      if (st.st_size == 0) {
        mapped = true;  // an empty file: no records
      } else {
        void *m = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
        if (m != MAP_FAILED) {
          data.map = m;
          data.map_len = (size_t)st.st_size;
          data.p = (const char *)m;
          data.n = (size_t)st.st_size;
          mapped = true;
        }
      }
    }
    close(fd);
    if (!mapped) {
      gzFile gz = gzopen(path.c_str(), "rb");
      REQUIRE(gz != nullptr, IMPGX_E_IO, "cannot open PAF file '" + path + "'");
      std::vector<char> buf(1 << 20);
      int k;
      while ((k = gzread(gz, buf.data(), (unsigned)buf.size())) > 0) data.owned.append(buf.data(), (size_t)k);
      const bool bad = k < 0;
      gzclose(gz);
      REQUIRE(!bad, IMPGX_E_IO, "error while reading / inflating '" + path + "'");
      data.p = data.owned.data();
      data.n = data.owned.size();
    }
  }
  // large files: lines parsed on every host core (below); small ones keep the plain loop
  {
    static const size_t min_parallel = [] {
      const char *v = getenv("IMPGX_PAF_PARALLEL_MIN_BYTES");
      return (v && *v) ? (size_t)strtoull(v, nullptr, 10) : (size_t)(8u << 20);
    }();
    if (data.size() >= min_parallel) {
      parse_paf_parallel(data, path, out);
      return;
    }
  }
  size_t pos = 0, line_no = 0;
  uint64_t ref_pos = 0;
  auto get_id = [&](const char *b, const char *e, uint64_t len) {
    std::string name(b, e);
    auto it = out.ids.find(name);
    if (it != out.ids.end()) return it->second;
    uint32_t id = (uint32_t)out.names.size();
    out.ids.emplace(name, id);
    out.names.push_back(name);
    out.lens.push_back(len);
    return id;
  };
  while (pos < data.size()) {
    size_t eol = data.find('\n', pos);
    if (eol == std::string::npos) eol = data.size();
    size_t end = eol;
    if (end > pos && data[end - 1] == '\r') end--;
    line_no++;
    std::vector<std::pair<const char *, const char *>> fld;
    {
      const char *b = data.data() + pos, *e = data.data() + end, *p = b;
      for (;;) {
        const char *t = (const char *)memchr(p, '\t', (size_t)(e - p));
        if (!t) {
          fld.emplace_back(p, e);
          break;
        }
        fld.emplace_back(p, t);
        p = t + 1;
      }
    }
    const std::string where = " (line " + std::to_string(line_no) + " of '" + path + "')";
    REQUIRE(fld.size() >= 12, IMPGX_E_PARSE, "Not enough fields in PAF record" + where);
    uint64_t ql, qs, qe, tl, ts, te;
    REQUIRE(parse_usize(fld[1].first, fld[1].second, ql) && parse_usize(fld[2].first, fld[2].second, qs) &&
                parse_usize(fld[3].first, fld[3].second, qe) && parse_usize(fld[6].first, fld[6].second, tl) &&
                parse_usize(fld[7].first, fld[7].second, ts) && parse_usize(fld[8].first, fld[8].second, te),
            IMPGX_E_PARSE, "Invalid field" + where);
    REQUIRE(fld[4].first < fld[4].second, IMPGX_E_PARSE, "Expected '+' or '-' for strand" + where);
    const char sc = *fld[4].first;
    REQUIRE(sc == '+' || sc == '-', IMPGX_E_PARSE, "Invalid strand" + where);
    REQUIRE(ql <= INT32_MAX && tl <= INT32_MAX && qe <= INT32_MAX && te <= INT32_MAX && qs <= INT32_MAX && ts <= INT32_MAX,
            IMPGX_E_PARSE,
            "coordinate beyond 2^31-1" + where);
    impgx_record r;
    r.query_id = get_id(fld[0].first, fld[0].second, ql);
    r.target_id = get_id(fld[5].first, fld[5].second, tl);
    r.query_start = (int32_t)qs; r.query_end = (int32_t)qe;
    r.target_start = (int32_t)ts; r.target_end = (int32_t)te;
    r.strand = sc == '-' ? 1 : 0;
    r.reserved = 0;
    bool have = false;
    // byte offset / length of the CIGAR text as the reference records them (src/paf.rs:150-162, :182-191:
    // line lengths are counted without the line terminator, plus one)
    uint64_t cg_off = ref_pos, cg_len = 0;
    for (auto &t : fld) {
      if (t.second - t.first >= 5 && memcmp(t.first, "cg:Z:", 5) == 0) {
        long k = parse_cigar(t.first + 5, (size_t)(t.second - t.first - 5), out.runs);
        REQUIRE(k >= 0, IMPGX_E_PARSE, "Invalid CIGAR operation" + where);
        have = k > 0;
        cg_off += 5;
        cg_len = (uint64_t)(t.second - t.first - 5);
        break;
      }
      cg_off += (uint64_t)(t.second - t.first) + 1;
    }
    ref_pos += (uint64_t)(end - pos) + 1;
    // the reference panics at query time when an alignment has no cg:Z tag
    // (src/impg.rs:506-511); here it is a build-time error
    REQUIRE(have, IMPGX_E_PARSE, "The alignment file does not contain CIGAR strings ('cg:Z' tag)" + where);
    out.recs.push_back(r);
    out.run_off.push_back(out.runs.size());
    out.cg_off.push_back(cg_off);
    out.cg_len.push_back(cg_len);
    out.file_idx.push_back(out.n_files);
    pos = eol + 1;
  }
  out.n_files++;
}

}  // namespace impgx

#define API_BEGIN try {
#define API_END                                  \
  }                                              \
  catch (const impgx::Error &e) {                \
    impgx::set_last_error(e.what());             \
    return e.code;                               \
  }                                              \
  catch (const std::bad_alloc &) {               \
    impgx::set_last_error("host allocation failed"); \
    return IMPGX_E_NOMEM;                        \
  }                                              \
  catch (const std::exception &e) {              \
    impgx::set_last_error(e.what());             \
    return IMPGX_E_INVALID;                      \
  }                                              \
  return IMPGX_OK;

extern "C" {

int impgx_abi_version(void) { return IMPGX_ABI_VERSION; }
const char *impgx_last_error(void) { return impgx::last_error(); }
int impgx_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

int impgx_index_build(const impgx_record *records, size_t n_records, const uint32_t *runs, const uint64_t *run_offsets,
                      const uint64_t *seq_lens, uint32_t n_seqs, int bidirectional, int device, impgx_index **out) {
  API_BEGIN
  REQUIRE(out, IMPGX_E_INVALID, "out is NULL");
  *out = nullptr;
  REQUIRE(runs || n_records == 0 || run_offsets[n_records] == 0, IMPGX_E_INVALID, "runs is NULL");
  *out = impgx::index_build(records, n_records, runs, run_offsets, seq_lens, n_seqs, bidirectional != 0, device);
  API_END
}

int impgx_index_from_paf(const char *paf_path, int bidirectional, int device, impgx_index **out) {
  API_BEGIN
  REQUIRE(out && paf_path, IMPGX_E_INVALID, "NULL argument");
  *out = nullptr;
  impgx::check_device(device);  // fail before parsing: there is no CPU fallback
  impgx::PafData pd;
  impgx::parse_paf(paf_path, pd);
  impgx_index *idx = impgx::index_build(pd.recs.data(), pd.recs.size(), pd.runs.data(), pd.run_off.data(), pd.lens.data(),
                                        (uint32_t)pd.lens.size(), bidirectional != 0, device);
  idx->names = pd.names;
  idx->name_to_id = pd.ids;
  *out = idx;
  API_END
}

int impgx_index_from_pafs(const char *const *paf_paths, size_t n_paths, int bidirectional, int device, impgx_index **out) {
  API_BEGIN
  REQUIRE(out && paf_paths && n_paths >= 1, IMPGX_E_INVALID, "NULL argument");
  *out = nullptr;
  impgx::check_device(device);
  impgx::PafData pd;  // shared SequenceIndex: unified ids by first appearance over the files (src/multi_impg.rs:159-176)
  for (size_t f = 0; f < n_paths; f++) {
    REQUIRE(paf_paths[f], IMPGX_E_INVALID, "NULL path");
    impgx::parse_paf(paf_paths[f], pd);
  }
  impgx_index *idx = impgx::index_build(pd.recs.data(), pd.recs.size(), pd.runs.data(), pd.run_off.data(), pd.lens.data(),
                                        (uint32_t)pd.lens.size(), bidirectional != 0, device);
  idx->names = pd.names;
  idx->name_to_id = pd.ids;
  *out = idx;
  API_END
}

void impgx_index_free(impgx_index *idx) { delete idx; }
uint32_t impgx_index_num_seqs(const impgx_index *idx) { return idx ? idx->n_seqs : 0; }
uint64_t impgx_index_num_entries(const impgx_index *idx) { return idx ? idx->n_entries : 0; }
uint64_t impgx_index_device_bytes(const impgx_index *idx) { return idx ? idx->device_bytes : 0; }
const char *impgx_index_seq_name(const impgx_index *idx, uint32_t id) {
  if (!idx || id >= idx->names.size() || idx->names[id].empty()) return nullptr;
  return idx->names[id].c_str();
}
uint64_t impgx_index_seq_len(const impgx_index *idx, uint32_t id) {
  return (idx && id < idx->seq_lens.size()) ? idx->seq_lens[id] : 0;
}
int impgx_index_seq_id(const impgx_index *idx, const char *name, uint32_t *id_out) {
  API_BEGIN
  REQUIRE(idx && name && id_out, IMPGX_E_INVALID, "NULL argument");
  auto it = idx->name_to_id.find(name);
  REQUIRE(it != idx->name_to_id.end(), IMPGX_E_INVALID, std::string("Target sequence '") + name + "' not found in index");
  *id_out = it->second;
  API_END
}
int impgx_index_set_names(impgx_index *idx, const char *const *names, uint32_t n) {
  API_BEGIN
  REQUIRE(idx && names && n == idx->n_seqs, IMPGX_E_INVALID, "names must cover every sequence");
  idx->names.assign(names, names + n);
  idx->name_to_id.clear();
  for (uint32_t i = 0; i < n; i++) idx->name_to_id[idx->names[i]] = i;
  API_END
}

static int do_query(impgx_index *idx, const impgx_range *ranges, size_t n, const impgx_params *params, bool bed,
                    bool dev_in, bool host_out, void *stream, impgx_results **out) {
  API_BEGIN
  REQUIRE(out && params, IMPGX_E_INVALID, "NULL argument");
  *out = nullptr;
  *out = impgx::query_batch(idx, ranges, n, *params, bed, dev_in, host_out, stream);
  API_END
}

int impgx_query_batch(impgx_index *idx, const impgx_range *ranges, size_t n, const impgx_params *params,
                      impgx_results **out) {
  return do_query(idx, ranges, n, params, false, false, true, nullptr, out);
}
int impgx_query_batch_bed(impgx_index *idx, const impgx_range *ranges, size_t n, const impgx_params *params,
                          impgx_results **out) {
  return do_query(idx, ranges, n, params, true, false, true, nullptr, out);
}
int impgx_query_batch_bed_device(impgx_index *idx, const impgx_range *d_ranges, size_t n, const impgx_params *params,
                                 void *stream, impgx_results **out) {
  return do_query(idx, d_ranges, n, params, true, true, false, stream, out);
}

// ---- target-sharded index
int impgx_comm_unique_id(uint8_t id[IMPGX_COMM_ID_BYTES]) {
  API_BEGIN
  REQUIRE(id, IMPGX_E_INVALID, "NULL argument");
  impgx::nccl_unique_id(id);
  API_END
}
int impgx_comm_init_nccl(const uint8_t id[IMPGX_COMM_ID_BYTES], int rank, int n_ranks, int device, impgx_comm **out) {
  API_BEGIN
  REQUIRE(id && out, IMPGX_E_INVALID, "NULL argument");
  *out = nullptr;
  impgx::check_device(device);
  std::unique_ptr<impgx_comm> c(new impgx_comm());
  c->c.reset(impgx::nccl_comm_create(id, rank, n_ranks, device));
  *out = c.release();
  API_END
}
int impgx_comm_init_local(int n_ranks, impgx_comm **out) {
  API_BEGIN
  REQUIRE(out, IMPGX_E_INVALID, "NULL argument");
  std::vector<impgx::Comm *> g = impgx::local_comm_group(n_ranks);
  for (int r = 0; r < n_ranks; r++) {
    out[r] = new impgx_comm();
    out[r]->c.reset(g[r]);
  }
  API_END
}
int impgx_comm_rank(const impgx_comm *c) { return c ? c->c->rank() : -1; }
int impgx_comm_size(const impgx_comm *c) { return c ? c->c->size() : 0; }
int impgx_comm_traffic(const impgx_comm *c, uint64_t *sent, uint64_t *received, uint64_t *exchanges) {
  API_BEGIN
  REQUIRE(c, IMPGX_E_INVALID, "NULL argument");
  if (sent) *sent = c->c->bytes_sent;
  if (received) *received = c->c->bytes_received;
  if (exchanges) *exchanges = c->c->exchanges;
  API_END
}
void impgx_comm_free(impgx_comm *c) { delete c; }

int impgx_assign_owners(const impgx_record *records, size_t n_records, const uint64_t *run_offsets, uint32_t n_seqs,
                        int bidirectional, uint32_t n_ranks, uint32_t *owner_out) {
  API_BEGIN
  REQUIRE((records || n_records == 0) && run_offsets && owner_out, IMPGX_E_INVALID, "NULL argument");
  impgx::assign_owners(records, n_records, run_offsets, n_seqs, bidirectional != 0, n_ranks, owner_out);
  API_END
}
int impgx_index_build_shard(const impgx_record *records, size_t n_records, const uint32_t *runs,
                            const uint64_t *run_offsets, const uint64_t *seq_lens, uint32_t n_seqs, int bidirectional,
                            int device, const uint32_t *owner, uint32_t rank, uint32_t n_ranks, impgx_index **out) {
  API_BEGIN
  REQUIRE(out && owner, IMPGX_E_INVALID, "NULL argument");
  *out = nullptr;
  REQUIRE(runs || n_records == 0 || run_offsets[n_records] == 0, IMPGX_E_INVALID, "runs is NULL");
  *out = impgx::index_build(records, n_records, runs, run_offsets, seq_lens, n_seqs, bidirectional != 0, device, owner,
                            rank, n_ranks);
  API_END
}
int impgx_query_batch_bed_sharded(impgx_index *shard, impgx_comm *comm, const impgx_range *ranges, size_t n,
                                  const impgx_params *params, impgx_results **out) {
  API_BEGIN
  REQUIRE(out && params && comm, IMPGX_E_INVALID, "NULL argument");
  *out = nullptr;
  *out = impgx::query_batch(shard, ranges, n, *params, true, false, true, nullptr, comm->c.get());
  API_END
}
int impgx_query_batch_bed_sharded_device(impgx_index *shard, impgx_comm *comm, const impgx_range *d_ranges, size_t n,
                                         const impgx_params *params, void *stream, impgx_results **out) {
  API_BEGIN
  REQUIRE(out && params && comm, IMPGX_E_INVALID, "NULL argument");
  *out = nullptr;
  *out = impgx::query_batch(shard, d_ranges, n, *params, true, true, false, stream, comm->c.get());
  API_END
}
// Rows of one input row are sorted by sequence id in the reference's BED output
// (stage B sorts by (q_id, start, strand)); each part holds whole (row, q_id)
// groups, so the union is a k-way merge by q_id per row.
int impgx_results_merge_shards(const impgx_results *const *parts, int n_parts, impgx_results **out) {
  API_BEGIN
  REQUIRE(parts && out && n_parts >= 1, IMPGX_E_INVALID, "NULL argument");
  *out = nullptr;
  const size_t n_rows = parts[0]->n_rows;
  uint64_t total = 0;
  for (int p = 0; p < n_parts; p++) {
    REQUIRE(parts[p] && !parts[p]->on_device && parts[p]->n_rows == n_rows && !parts[p]->has_cigar, IMPGX_E_INVALID,
            "parts must be host-resident BED results of the same batch");
    total += parts[p]->n_results;
  }
  std::unique_ptr<impgx_results> res(new impgx_results());
  res->device = parts[0]->device;
  res->n_rows = n_rows;
  res->n_results = total;
  res->row_off.assign(n_rows + 1, 0);
  res->qid.resize(total); res->tid.resize(total);
  res->qf.resize(total); res->ql.resize(total); res->tf.resize(total); res->tl.resize(total);
  uint64_t w = 0;
  std::vector<uint64_t> cur((size_t)n_parts);
  for (size_t r = 0; r < n_rows; r++) {
    for (int p = 0; p < n_parts; p++) cur[p] = parts[p]->row_off[r];
    for (;;) {
      int best = -1;
      uint32_t bq = 0;
      for (int p = 0; p < n_parts; p++) {
        if (cur[p] >= parts[p]->row_off[r + 1]) continue;
        const uint32_t q = parts[p]->qid[cur[p]];
        if (best < 0 || q < bq) {
          best = p;
          bq = q;
        }
      }
      if (best < 0) break;
      const impgx_results *P = parts[best];
      uint64_t &c = cur[best];
      while (c < P->row_off[r + 1] && P->qid[c] == bq) {
        res->qid[w] = P->qid[c]; res->tid[w] = P->tid[c];
        res->qf[w] = P->qf[c]; res->ql[w] = P->ql[c]; res->tf[w] = P->tf[c]; res->tl[w] = P->tl[c];
        w++;
        c++;
      }
    }
    res->row_off[r + 1] = w;
  }
  *out = res.release();
  API_END
}

int impgx_results_view(const impgx_results *res, impgx_view *v) {
  API_BEGIN
  REQUIRE(res && v, IMPGX_E_INVALID, "NULL argument");
  REQUIRE(!res->on_device, IMPGX_E_INVALID, "results are device resident; use impgx_results_device_view");
  v->n_rows = res->n_rows;
  v->n_results = res->n_results;
  v->row_offsets = res->row_off.data();
  v->q_id = res->qid.data(); v->q_first = res->qf.data(); v->q_last = res->ql.data();
  v->t_id = res->tid.data(); v->t_first = res->tf.data(); v->t_last = res->tl.data();
  v->cigar_offsets = res->has_cigar ? res->cig_off.data() : nullptr;
  v->cigar_runs = res->has_cigar ? res->cig.data() : nullptr;
  API_END
}
int impgx_results_device_view(const impgx_results *res, impgx_view *v) {
  API_BEGIN
  REQUIRE(res && v, IMPGX_E_INVALID, "NULL argument");
  REQUIRE(res->on_device, IMPGX_E_INVALID, "results are host resident; use impgx_results_view");
  v->n_rows = res->n_rows;
  v->n_results = res->n_results;
  v->row_offsets = res->d_row_off;
  v->q_id = res->d_qid; v->q_first = res->d_qf; v->q_last = res->d_ql;
  v->t_id = res->d_tid; v->t_first = res->d_tf; v->t_last = res->d_tl;
  v->cigar_offsets = nullptr;
  v->cigar_runs = nullptr;
  API_END
}
void impgx_results_free(impgx_results *res) { delete res; }

int impgx_index_stats(const impgx_index *idx, impgx_stats *out) {
  API_BEGIN
  REQUIRE(idx && out, IMPGX_E_INVALID, "NULL argument");
  *out = idx->last;
  API_END
}

int impgx_project_batch(int device, size_t n, const int32_t *req_start, const int32_t *req_end,
                        const impgx_record *records, const uint32_t *runs, const uint64_t *run_offsets, int32_t *out4,
                        uint8_t *ok, uint64_t *out_run_offsets, uint32_t *out_runs, size_t out_runs_cap) {
  API_BEGIN
  impgx::project_batch(device, n, req_start, req_end, records, runs, run_offsets, out4, ok, out_run_offsets, out_runs,
                       out_runs_cap);
  API_END
}

long impgx_parse_cigar(const char *text, size_t len, uint32_t *out, size_t cap) {
  std::vector<uint32_t> v;
  long k = impgx::parse_cigar(text, len, v);
  if (k < 0) {
    impgx::set_last_error("Invalid CIGAR operation");
    return IMPGX_E_PARSE;
  }
  if ((size_t)k > cap) {
    impgx::set_last_error("output capacity too small");
    return IMPGX_E_INVALID;
  }
  if (k) memcpy(out, v.data(), (size_t)k * 4);
  return k;
}

char *impgx_format_bed(const impgx_index *idx, const impgx_results *res, size_t row, const char *name) {
  if (!idx || !res || res->on_device || row >= res->n_rows || !name) return nullptr;
  std::string s;
  for (uint64_t i = res->row_off[row]; i < res->row_off[row + 1]; i++) {
    int32_t f = res->qf[i], l = res->ql[i];
    char strand = '+';
    if (f > l) {
      std::swap(f, l);
      strand = '-';
    }
    uint32_t id = res->qid[i];
    std::string qn = (id < idx->names.size() && !idx->names[id].empty()) ? idx->names[id] : "seq" + std::to_string(id);
    uint32_t uf = (uint32_t)f, ul = (uint32_t)l, off = 0;
    std::string base;
    if (idx->original_coordinates && impgx::to_original_coordinates(qn, base, off)) {
      qn = base;
      uf += off;
      ul += off;
    }
    s += qn;
    s += '\t';
    s += std::to_string(uf);
    s += '\t';
    s += std::to_string(ul);
    s += '\t';
    s += name;
    s += "\t.\t";
    s += strand;
    s += '\n';
  }
  char *p = (char *)malloc(s.size() + 1);
  if (p) memcpy(p, s.c_str(), s.size() + 1);
  return p;
}

// All rows of a merged result set at once, formatted by every host core: the text `impg query -b ... -o bed`
// prints for the whole BED file (rows in input order). Two passes: line lengths, then the bytes.
namespace {
inline int dec_len(uint32_t v) {
  int n = 1;
  while (v >= 10) {
    v /= 10;
    n++;
  }
  return n;
}
inline char *put_dec(char *p, uint32_t v) {
  const int n = dec_len(v);
  for (int i = n - 1; i >= 0; i--) {
    p[i] = (char)('0' + v % 10);
    v /= 10;
  }
  return p + n;
}
}  // namespace

char *impgx_format_bed_batch(const impgx_index *idx, const impgx_results *res, const char *const *names,
                             size_t *len_out) {
  if (!idx || !res || res->on_device || !names) return nullptr;
  const size_t R = res->n_rows;
  std::vector<std::string> fallback(idx->n_seqs);
  std::vector<const std::string *> seq(idx->n_seqs);
  std::vector<uint32_t> shift(idx->n_seqs, 0);  // --original-sequence-coordinates: subsequence start per sequence
  for (uint32_t q = 0; q < idx->n_seqs; q++) {
    if (q < idx->names.size() && !idx->names[q].empty()) seq[q] = &idx->names[q];
    else {
      fallback[q] = "seq" + std::to_string(q);
      seq[q] = &fallback[q];
    }
    std::string base;
    uint32_t off = 0;
    if (idx->original_coordinates && impgx::to_original_coordinates(*seq[q], base, off)) {
      fallback[q] = base;
      seq[q] = &fallback[q];
      shift[q] = off;
    }
  }
  std::vector<size_t> name_len(R), off(R + 1, 0);
#pragma omp parallel for schedule(static)
  for (long r = 0; r < (long)R; r++) {
    name_len[r] = strlen(names[r]);
    size_t b = 0;
    for (uint64_t i = res->row_off[r]; i < res->row_off[r + 1]; i++) {
      const int32_t f = res->qf[i], l = res->ql[i];
      const uint32_t sh = shift[res->qid[i]];
      b += seq[res->qid[i]]->size() + dec_len((uint32_t)std::min(f, l) + sh) + dec_len((uint32_t)std::max(f, l) + sh) + name_len[r] + 8;
    }
    off[r + 1] = b;
  }
  for (size_t r = 0; r < R; r++) off[r + 1] += off[r];
  char *out = (char *)malloc(off[R] + 1);
  if (!out) return nullptr;
#pragma omp parallel for schedule(static)
  for (long r = 0; r < (long)R; r++) {
    char *p = out + off[r];
    for (uint64_t i = res->row_off[r]; i < res->row_off[r + 1]; i++) {
      const int32_t f = res->qf[i], l = res->ql[i];
      const std::string &nm = *seq[res->qid[i]];
      memcpy(p, nm.data(), nm.size());
      p += nm.size();
      *p++ = '\t';
      const uint32_t sh = shift[res->qid[i]];
      p = put_dec(p, (uint32_t)std::min(f, l) + sh);
      *p++ = '\t';
      p = put_dec(p, (uint32_t)std::max(f, l) + sh);
      *p++ = '\t';
      memcpy(p, names[r], name_len[r]);
      p += name_len[r];
      *p++ = '\t';
      *p++ = '.';
      *p++ = '\t';
      *p++ = f > l ? '-' : '+';
      *p++ = '\n';
    }
  }
  out[off[R]] = 0;
  if (len_out) *len_out = off[R];
  return out;
}

static char *format_with(const impgx_index *idx, const impgx_results *res, size_t row, const char *name, int32_t d,
                         int fmt) {
  if (!idx || !res || res->on_device || row >= res->n_rows || !name) return nullptr;
  try {
    // PAF in original coordinates needs the lengths of the original sequences from the FASTA / AGC index
    // (get_original_sequence_length, src/main.rs:4680-4705): sequence access is outside the path
    REQUIRE(!(fmt == 2 && idx->original_coordinates), IMPGX_E_UNSUPPORTED,
            "--original-sequence-coordinates with PAF output needs the sequence files (outside the accelerated path)");
    std::string s = impgx::format_rows(idx, res, row, name, d, fmt);
    char *p = (char *)malloc(s.size() + 1);
    if (p) memcpy(p, s.c_str(), s.size() + 1);
    return p;
  } catch (const std::exception &e) {
    impgx::set_last_error(e.what());
    return nullptr;
  }
}
char *impgx_format_bedpe(const impgx_index *idx, const impgx_results *res, size_t row, const char *name,
                         int32_t merge_distance) {
  return format_with(idx, res, row, name, merge_distance, 1);
}
char *impgx_format_paf(const impgx_index *idx, const impgx_results *res, size_t row, const char *name,
                       int32_t merge_distance) {
  return format_with(idx, res, row, name, merge_distance, 2);
}

void impgx_free(void *p) { free(p); }

// ---- BED / range parsing (src/commands/partition.rs:1719-1789)
}  // extern "C"
struct impgx_bed {
  std::vector<std::string> seq, name;
  std::vector<int32_t> start, end;
};
namespace {
bool parse_i32(const std::string &t, int32_t &v) {
  // Rust i32::from_str: optional sign, digits only, no whitespace
  if (t.empty()) return false;
  size_t i = 0;
  bool neg = false;
  if (t[0] == '+' || t[0] == '-') {
    neg = t[0] == '-';
    i = 1;
  }
  if (i >= t.size()) return false;
  int64_t x = 0;
  for (; i < t.size(); i++) {
    if (t[i] < '0' || t[i] > '9') return false;
    x = x * 10 + (t[i] - '0');
    if (x > (int64_t)INT32_MAX + 1) return false;
  }
  x = neg ? -x : x;
  if (x > INT32_MAX || x < INT32_MIN) return false;
  v = (int32_t)x;
  return true;
}
void parse_range(const std::string &a, const std::string &b, int32_t &s, int32_t &e) {
  REQUIRE(parse_i32(a, s), IMPGX_E_PARSE, "Invalid start value");
  REQUIRE(parse_i32(b, e), IMPGX_E_PARSE, "Invalid end value");
  REQUIRE(s < e, IMPGX_E_PARSE, "Start value must be less than end value");
}
std::string trim(const std::string &t) {
  size_t a = 0, b = t.size();
  while (a < b && isspace((unsigned char)t[a])) a++;
  while (b > a && isspace((unsigned char)t[b - 1])) b--;
  return t.substr(a, b - a);
}
}  // namespace
extern "C" {

int impgx_bed_parse(const char *path, impgx_bed **out) {
  API_BEGIN
  REQUIRE(path && out, IMPGX_E_INVALID, "NULL argument");
  *out = nullptr;
  std::ifstream f(path, std::ios::binary);
  REQUIRE(f.good(), IMPGX_E_IO, std::string("cannot open BED file '") + path + "'");
  std::unique_ptr<impgx_bed> bed(new impgx_bed());
  std::string line;
  while (std::getline(f, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    std::vector<std::string> parts;
    size_t a = 0;
    for (;;) {
      size_t b = line.find('\t', a);
      if (b == std::string::npos) {
        parts.push_back(line.substr(a));
        break;
      }
      parts.push_back(line.substr(a, b - a));
      a = b + 1;
    }
    REQUIRE(parts.size() >= 3, IMPGX_E_PARSE, "Invalid BED file format");
    int32_t s, e;
    parse_range(parts[1], parts[2], s, e);
    std::string name;
    if (parts.size() > 3) {
      name = trim(parts[3]);
      if (name == ".") name.clear();
    }
    if (name.empty()) name = parts[0] + ":" + std::to_string(s) + "-" + std::to_string(e);
    bed->seq.push_back(parts[0]);
    bed->name.push_back(name);
    bed->start.push_back(s);
    bed->end.push_back(e);
  }
  *out = bed.release();
  API_END
}
size_t impgx_bed_len(const impgx_bed *b) { return b ? b->seq.size() : 0; }
const char *impgx_bed_seq(const impgx_bed *b, size_t i) { return (b && i < b->seq.size()) ? b->seq[i].c_str() : nullptr; }
const char *impgx_bed_name(const impgx_bed *b, size_t i) { return (b && i < b->name.size()) ? b->name[i].c_str() : nullptr; }
int32_t impgx_bed_start(const impgx_bed *b, size_t i) { return (b && i < b->start.size()) ? b->start[i] : 0; }
int32_t impgx_bed_end(const impgx_bed *b, size_t i) { return (b && i < b->end.size()) ? b->end[i] : 0; }
void impgx_bed_free(impgx_bed *b) { delete b; }

int impgx_parse_target_range(const char *text, char *seq_out, size_t seq_cap, int32_t *start, int32_t *end,
                             char *name_out, size_t name_cap) {
  API_BEGIN
  REQUIRE(text && seq_out && start && end, IMPGX_E_INVALID, "NULL argument");
  const std::string t(text);
  const size_t c = t.rfind(':');
  REQUIRE(c != std::string::npos, IMPGX_E_PARSE, "Target range format should be `seq_name:start-end`");
  const std::string seq = t.substr(0, c), rng = t.substr(c + 1);
  std::vector<std::string> p;
  size_t a = 0;
  for (;;) {
    size_t b = rng.find('-', a);
    if (b == std::string::npos) {
      p.push_back(rng.substr(a));
      break;
    }
    p.push_back(rng.substr(a, b - a));
    a = b + 1;
  }
  REQUIRE(p.size() == 2, IMPGX_E_PARSE, "Range format should be `start-end`");
  parse_range(p[0], p[1], *start, *end);
  REQUIRE(seq.size() + 1 <= seq_cap, IMPGX_E_INVALID, "seq_out too small");
  memcpy(seq_out, seq.c_str(), seq.size() + 1);
  if (name_out) {
    const std::string name = seq + ":" + std::to_string(*start) + "-" + std::to_string(*end);
    REQUIRE(name.size() + 1 <= name_cap, IMPGX_E_INVALID, "name_out too small");
    memcpy(name_out, name.c_str(), name.size() + 1);
  }
  API_END
}

// ---- debug / test hooks (host only, usable without a GPU)
// visit ranks of a target with n entries
void impgx_debug_visit_ranks(size_t n, uint32_t *rank) { impgx::visit_ranks(n, rank); }

// host columns of an index build (no device needed): returns entry count, fills
// the arrays when non-NULL (sized by a first call with NULLs)
long impgx_debug_host_columns_shard(const impgx_record *records, size_t n, const uint64_t *run_offsets, uint32_t n_seqs,
                                    int bidirectional, int32_t *e_start, int32_t *e_end, int32_t *e_pmax,
                                    uint32_t *e_vrank, uint32_t *e_query_id, uint32_t *e_flags, uint32_t *e_aln,
                                    uint64_t *tgt_off, const uint32_t *owner, uint32_t rank);
long impgx_debug_host_columns(const impgx_record *records, size_t n, const uint64_t *run_offsets, uint32_t n_seqs,
                              int bidirectional, int32_t *e_start, int32_t *e_end, int32_t *e_pmax, uint32_t *e_vrank,
                              uint32_t *e_query_id, uint32_t *e_flags, uint32_t *e_aln, uint64_t *tgt_off) {
  return impgx_debug_host_columns_shard(records, n, run_offsets, n_seqs, bidirectional, e_start, e_end, e_pmax, e_vrank,
                                        e_query_id, e_flags, e_aln, tgt_off, nullptr, 0);
}
long impgx_debug_host_columns_shard(const impgx_record *records, size_t n, const uint64_t *run_offsets, uint32_t n_seqs,
                                    int bidirectional, int32_t *e_start, int32_t *e_end, int32_t *e_pmax,
                                    uint32_t *e_vrank, uint32_t *e_query_id, uint32_t *e_flags, uint32_t *e_aln,
                                    uint64_t *tgt_off, const uint32_t *owner, uint32_t rank) {
  try {
    impgx::HostColumns hc;
    impgx::build_host_columns(records, n, run_offsets, n_seqs, bidirectional != 0, hc, owner, rank);
    size_t E = hc.e_start.size();
    if (e_start) {
      memcpy(e_start, hc.e_start.data(), E * 4);
      memcpy(e_end, hc.e_end.data(), E * 4);
      memcpy(e_pmax, hc.e_pmax.data(), E * 4);
      memcpy(e_vrank, hc.e_vrank.data(), E * 4);
      for (size_t i = 0; i < E; i++) {
        e_query_id[i] = hc.e_rec[i].query_id;
        e_flags[i] = hc.e_rec[i].nruns_flags;
        e_aln[i] = hc.e_aln[i];
      }
      memcpy(tgt_off, hc.tgt_off.data(), (n_seqs + 1) * 8);
    }
    return (long)E;
  } catch (const std::exception &e) {
    impgx::set_last_error(e.what());
    return -1;
  }
}

int impgx_index_set_original_coordinates(impgx_index *idx, int on) {
  if (!idx) return IMPGX_E_INVALID;
  idx->original_coordinates = on != 0;
  return IMPGX_OK;
}
int impgx_parse_subsequence_coordinates(const char *seq_name, char *base_out, size_t base_cap, int32_t *start_out) {
  if (!seq_name) return IMPGX_E_INVALID;
  std::string base;
  uint32_t off = 0;
  if (!impgx::to_original_coordinates(seq_name, base, off)) return 0;
  if (base_out && base_cap) {
    const size_t n = std::min(base.size(), base_cap - 1);
    memcpy(base_out, base.data(), n);
    base_out[n] = 0;
  }
  if (start_out) *start_out = (int32_t)off;
  return 1;
}

// parse_merge_distance (src/main.rs:47-55 over sweepga::parse_metric_number, un-vendored): a non-negative
// decimal number with an optional k / m / g suffix (either case), at most i32::MAX after scaling; pinned by the
// reference's tests (src/main.rs:13702-13715: 50000, 50k, 1m, 1M, 1.5k accepted; 10kb, 3g rejected)
int impgx_parse_merge_distance(const char *text, int32_t *out) {
  if (!text || !out) return IMPGX_E_INVALID;
  std::string t(text);
  while (!t.empty() && isspace((unsigned char)t.back())) t.pop_back();
  size_t b = 0;
  while (b < t.size() && isspace((unsigned char)t[b])) b++;
  t = t.substr(b);
  double mul = 1;
  if (!t.empty()) {
    const char c = (char)tolower((unsigned char)t.back());
    if (c == 'k' || c == 'm' || c == 'g') {
      mul = c == 'k' ? 1e3 : (c == 'm' ? 1e6 : 1e9);
      t.pop_back();
    }
  }
  bool digits = false, ok = !t.empty();
  int dots = 0;
  for (char c : t) {
    if (c >= '0' && c <= '9') digits = true;
    else if (c == '.') dots++;
    else ok = false;
  }
  if (!ok || !digits || dots > 1) {
    impgx::set_last_error(std::string("invalid merge distance '") + text + "'");
    return IMPGX_E_PARSE;
  }
  const double v = strtod(t.c_str(), nullptr) * mul;
  if (v > 2147483647.0) {
    impgx::set_last_error(std::string("merge distance ") + text + " exceeds maximum supported value 2147483647");
    return IMPGX_E_PARSE;
  }
  *out = (int32_t)llround(v);
  return IMPGX_OK;
}

// ---- --subset-sequence-list (src/subset_filter.rs): which sequences a list file keeps
namespace {
struct SubsetList {
  std::set<std::string> exact, normalized, sample_ids;
  std::set<std::pair<std::string, std::string>> sample_haps;
};
std::string trim_ws(const std::string &t) {
  size_t b = 0, e = t.size();
  while (b < e && isspace((unsigned char)t[b])) b++;
  while (e > b && isspace((unsigned char)t[e - 1])) e--;
  return t.substr(b, e - b);
}
std::string leading_digits(const std::string &t) {
  size_t k = 0;
  while (k < t.size() && t[k] >= '0' && t[k] <= '9') k++;
  return t.substr(0, k);
}
// extract_sample_and_hap (:147-178): "<sample>_hap<digits>…", PanSN "<sample>#<hap>#…", or a bare name
bool sample_and_hap(const std::string &name, std::string &sample, std::string &hap, bool &has_hap) {
  size_t k = name.find("_hap");
  if (k != std::string::npos) {
    sample = name.substr(0, k);
    hap = leading_digits(name.substr(k + 4));
    has_hap = !hap.empty();
    return true;
  }
  k = name.find('#');
  if (k != std::string::npos) {
    sample = name.substr(0, k);
    const std::string rest = name.substr(k + 1);
    hap = leading_digits(rest.substr(0, rest.find('#')));
    has_hap = !hap.empty();
    return true;
  }
  if (name.find(':') == std::string::npos && !trim_ws(name).empty()) {
    sample = name;
    hap.clear();
    has_hap = false;
    return true;
  }
  return false;
}
SubsetList parse_subset_list(const std::string &text) {  // parse_subset_filter (:117-145)
  SubsetList f;
  size_t pos = 0;
  while (pos < text.size()) {
    size_t eol = text.find('\n', pos);
    if (eol == std::string::npos) eol = text.size();
    const std::string t = trim_ws(text.substr(pos, eol - pos));
    pos = eol + 1;
    if (t.empty() || t[0] == '#') continue;
    f.exact.insert(t);
    const std::string no_coords = t.substr(0, t.find(':'));
    f.normalized.insert(no_coords);
    std::string sample, hap;
    bool has_hap = false;
    if (sample_and_hap(no_coords, sample, hap, has_hap)) {
      if (has_hap) f.sample_haps.insert({sample, hap});
      else f.sample_ids.insert(sample);
    }
  }
  return f;
}
bool sample_keys_match(const SubsetList &f, const std::string &name) {  // :44-58
  std::string sample, hap;
  bool has_hap = false;
  if (!sample_and_hap(name, sample, hap, has_hap)) return false;
  if (has_hap && f.sample_haps.count({sample, hap})) return true;
  return f.sample_ids.count(sample) != 0;
}
bool subset_matches(const SubsetList &f, const std::string &name) {  // SubsetFilter::matches (:23-42)
  if (f.exact.count(name)) return true;
  const std::string no_coords = name.substr(0, name.find(':'));
  if (name != no_coords && f.exact.count(no_coords)) return true;
  if (f.normalized.count(no_coords)) return true;
  if (sample_keys_match(f, no_coords)) return true;
  return sample_keys_match(f, name);
}
}  // namespace

long impgx_subset_mask(const impgx_index *idx, const char *list_text, uint8_t *mask_out) {
  try {
    REQUIRE(idx && list_text && mask_out, IMPGX_E_INVALID, "NULL argument");
    const SubsetList f = parse_subset_list(list_text);
    REQUIRE(!f.exact.empty(), IMPGX_E_PARSE, "Subset sequence list did not contain any sequence names");
    for (uint32_t s = 0; s < idx->n_seqs; s++)
      mask_out[s] = (s < idx->names.size() && subset_matches(f, idx->names[s])) ? 1 : 0;
    return (long)f.exact.size();
  } catch (const impgx::Error &e) {
    impgx::set_last_error(e.what());
    return e.code;
  } catch (const std::exception &e) {
    impgx::set_last_error(e.what());
    return IMPGX_E_INVALID;
  }
}
int impgx_subset_matches(const char *list_text, const char *seq_name) {
  if (!list_text || !seq_name) return IMPGX_E_INVALID;
  try {
    return subset_matches(parse_subset_list(list_text), seq_name) ? 1 : 0;
  } catch (const std::exception &e) {
    impgx::set_last_error(e.what());
    return IMPGX_E_INVALID;
  }
}

// test hook: what parse_paf extracts from a PAF file (records, decoded runs, the reference's CIGAR offsets), without a
// device. Arrays may be NULL (sizes only). Returns the number of records or a negative status.
long impgx_debug_parse_paf(const char *path, impgx_record *recs, uint64_t *run_off, uint32_t *runs, uint64_t *cg_off,
                           uint64_t *cg_len, uint64_t *n_runs_out, uint32_t *n_seqs_out) {
  try {
    REQUIRE(path, IMPGX_E_INVALID, "NULL path");
    impgx::PafData pd;
    impgx::parse_paf(path, pd);
    const size_t n = pd.recs.size();
    if (recs) memcpy(recs, pd.recs.data(), n * sizeof(impgx_record));
    if (run_off) memcpy(run_off, pd.run_off.data(), (n + 1) * 8);
    if (runs && !pd.runs.empty()) memcpy(runs, pd.runs.data(), pd.runs.size() * 4);
    if (cg_off && n) memcpy(cg_off, pd.cg_off.data(), n * 8);
    if (cg_len && n) memcpy(cg_len, pd.cg_len.data(), n * 8);
    if (n_runs_out) *n_runs_out = pd.runs.size();
    if (n_seqs_out) *n_seqs_out = (uint32_t)pd.names.size();
    return (long)n;
  } catch (const impgx::Error &e) {
    impgx::set_last_error(e.what());
    return e.code;
  } catch (const std::exception &e) {
    impgx::set_last_error(e.what());
    return IMPGX_E_INVALID;
  }
}
// "name\tlength\n" per sequence id of the same parse; malloc'ed, impgx_free
char *impgx_debug_parse_paf_seqs(const char *path) {
  try {
    impgx::PafData pd;
    impgx::parse_paf(path, pd);
    std::string s;
    for (size_t i = 0; i < pd.names.size(); i++) s += pd.names[i] + "\t" + std::to_string(pd.lens[i]) + "\n";
    char *p = (char *)malloc(s.size() + 1);
    if (p) memcpy(p, s.c_str(), s.size() + 1);
    return p;
  } catch (const std::exception &e) {
    impgx::set_last_error(e.what());
    return nullptr;
  }
}

// test hook: output_results_bedpe / _paf (merge_adjusted_intervals + writer, host code) on rows given as arrays, so the
// CIGAR surgery can be compared with the oracle on adversarial rows without a device. format 1 = bedpe, 2 = paf.
char *impgx_debug_format_rows(const char *const *names, const uint64_t *lens, uint32_t n_seqs, size_t n,
                              const uint32_t *q_id, const int32_t *q_first, const int32_t *q_last, const uint32_t *t_id,
                              const int32_t *t_first, const int32_t *t_last, const uint64_t *cig_off, const uint32_t *cig,
                              const char *name, int32_t d, int format, int original_coordinates) {
  try {
    REQUIRE(name && (n == 0 || (q_id && q_first && q_last && t_id && t_first && t_last)) && (format == 1 || format == 2),
            IMPGX_E_INVALID, "bad argument");
    const std::string s = impgx::debug_format_rows(names, lens, n_seqs, n, q_id, q_first, q_last, t_id, t_first, t_last, cig_off,
                                                   cig, name, d, format, original_coordinates != 0);
    char *p = (char *)malloc(s.size() + 1);
    if (p) memcpy(p, s.c_str(), s.size() + 1);
    return p;
  } catch (const std::exception &e) {
    impgx::set_last_error(e.what());
    return nullptr;
  }
}

// test hooks: the device sort / scan primitives on host arrays (both the single-CTA kernels for small inputs
// and the CUB pipelines sit behind the same calls; the size decides)
int impgx_debug_sort_pairs(int device, uint64_t *keys, uint32_t *vals, uint64_t n, int begin_bit, int end_bit, int key_bytes) {
  API_BEGIN
  REQUIRE(keys && vals && (key_bytes == 4 || key_bytes == 8) && begin_bit >= 0 && end_bit > begin_bit && end_bit <= 8 * key_bytes,
          IMPGX_E_INVALID, "bad argument");
  impgx::debug_sort_pairs(device, keys, vals, n, begin_bit, end_bit, key_bytes);
  API_END
}
int impgx_debug_exclusive_scan(int device, uint64_t *a, uint64_t n_plus_1) {
  API_BEGIN
  REQUIRE(a && n_plus_1 >= 1, IMPGX_E_INVALID, "bad argument");
  impgx::debug_exclusive_scan(device, a, n_plus_1);
  API_END
}


}  // extern "C"
