// impg_file.cu — the reference's `.impg` index files: read an index written by
// stock impg into the HBM layout, and write one stock impg can open. Host code only.
//
// Layout (Impg::serialize_with_forest_map / load_from_file, src/impg.rs:1655-1850):
//   "IMPGIDX2" (what the writer always emits, :1659) | "IMPGIDX1" (legacy files, accepted by the loader :1793-1812)
//   u64 LE   offset of the forest map
//   SequenceIndex { name_to_id: map<String,u32>, id_to_name: map<u32,String>,
//                   id_to_len: map<u32,usize>, next_id: u32 }            (src/seqidx.rs:4-10)
//   per target: (u32 target_id, Vec<SerializableInterval{first: i32, last: i32,
//                metadata: QueryMetadata}>) in the tree's in-order (sorted) sequence (:235-240, :1686-1693)
//   ForestMap { entries: map<u32 target_id, u64 offset> }                  (src/forest_map.rs)
// every blob encoded by serde + bincode 2 `config::standard()`: little endian, VARIABLE-LENGTH
// integers (u < 251: one byte; 251 + u16; 252 + u32; 253 + u64; signed: zigzag first; usize as
// u64), sequences / maps / strings with a varint length prefix, struct fields in declaration
// order without names. bincode 2.0.1 is not vendored (Cargo.lock:186-189) and the reference has
// no .impg fixture: BYTE PARITY IS UNPINNED; tests/_impg_format.py restates the same published
// encoding independently in Python and both sides must agree byte for byte.
//
// An index file carries no CIGARs, only (alignment_file_index, byte offset, byte length) of the
// cg:Z text in the alignment files (QueryMetadata, src/impg.rs:164-173), so opening one needs the
// same alignment files; their CIGARs are decoded once into the HBM run stream as for any build.
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include <zlib.h>

#include "engine.cuh"

namespace impgx {

namespace {

constexpr uint64_t STRAND_BIT = 0x8000000000000000ull;    // src/impg.rs:177
constexpr uint64_t REVERSED_BIT = 0x4000000000000000ull;  // src/impg.rs:178

struct Entry {  // SerializableInterval + the target it is filed under
  uint32_t target_id;
  int32_t first, last;
  uint32_t query_id;
  int32_t target_start, target_end, query_start, query_end;
  uint32_t file_index;
  uint64_t strand_and_data_offset;
  uint64_t data_bytes;
};

struct Reader {
  const uint8_t *p, *e;
  uint8_t byte() {
    REQUIRE(p < e, IMPGX_E_PARSE, "truncated .impg file");
    return *p++;
  }
  uint64_t fixed(int n) {
    REQUIRE(e - p >= n, IMPGX_E_PARSE, "truncated .impg file");
    uint64_t v = 0;
    for (int i = 0; i < n; i++) v |= (uint64_t)p[i] << (8 * i);
    p += n;
    return v;
  }
  uint64_t varint() {
    const uint8_t b = byte();
    if (b < 251) return b;
    if (b == 251) return fixed(2);
    if (b == 252) return fixed(4);
    if (b == 253) return fixed(8);
    throw Error(IMPGX_E_PARSE, "128-bit integer in a .impg file");
  }
  uint32_t u32() {
    const uint64_t v = varint();
    REQUIRE(v <= 0xffffffffull, IMPGX_E_PARSE, "u32 out of range in .impg file");
    return (uint32_t)v;
  }
  int32_t i32() {
    const uint64_t z = varint();
    const int64_t v = (int64_t)(z >> 1) ^ -(int64_t)(z & 1);
    REQUIRE(v >= INT32_MIN && v <= INT32_MAX, IMPGX_E_PARSE, "i32 out of range in .impg file");
    return (int32_t)v;
  }
  std::string str() {
    const uint64_t n = varint();
    REQUIRE((uint64_t)(e - p) >= n, IMPGX_E_PARSE, "truncated string in .impg file");
    std::string s((const char *)p, (size_t)n);
    p += n;
    return s;
  }
};

struct Writer {
  std::vector<uint8_t> b;
  void fixed(uint64_t v, int n) {
    for (int i = 0; i < n; i++) b.push_back((uint8_t)(v >> (8 * i)));
  }
  void varint(uint64_t v) {
    if (v < 251) b.push_back((uint8_t)v);
    else if (v < (1ull << 16)) {
      b.push_back(251);
      fixed(v, 2);
    } else if (v < (1ull << 32)) {
      b.push_back(252);
      fixed(v, 4);
    } else {
      b.push_back(253);
      fixed(v, 8);
    }
  }
  void i32(int32_t v) { varint((uint64_t)(((uint32_t)v << 1) ^ (uint32_t)(v >> 31))); }
  void str(const std::string &s) {
    varint(s.size());
    b.insert(b.end(), s.begin(), s.end());
  }
};

}  // namespace

}  // namespace impgx

struct impgx_impg {
  int version = 2;
  bool bidirectional = false;  // some entry carries the REVERSED bit
  std::vector<std::string> names;  // by id
  std::vector<uint64_t> lens;
  std::vector<impgx::Entry> entries;  // file order (per target, in-order)
  std::vector<size_t> forward;        // indices of the non-reversed entries in PAF order (file index, offset)
};

namespace impgx {

static std::vector<uint8_t> slurp(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  REQUIRE(f != nullptr, IMPGX_E_IO, "cannot open '" + path + "'");
  std::vector<uint8_t> d;
  uint8_t buf[1 << 16];
  size_t k;
  while ((k = fread(buf, 1, sizeof buf, f)) > 0) d.insert(d.end(), buf, buf + k);
  const bool bad = ferror(f) != 0;
  fclose(f);
  REQUIRE(!bad, IMPGX_E_IO, "error while reading '" + path + "'");
  return d;
}

static impgx_impg *open_impg(const std::string &path) {
  const std::vector<uint8_t> d = slurp(path);
  REQUIRE(d.size() >= 16 && (!memcmp(d.data(), "IMPGIDX2", 8) || !memcmp(d.data(), "IMPGIDX1", 8)), IMPGX_E_PARSE,
          "Invalid magic bytes - not a valid IMPG index file");
  std::unique_ptr<impgx_impg> f(new impgx_impg());
  f->version = d[7] == '1' ? 1 : 2;
  Reader hd{d.data() + 8, d.data() + 16};
  const uint64_t forest_off = hd.fixed(8);
  REQUIRE(forest_off >= 16 && forest_off <= d.size(), IMPGX_E_PARSE, "forest map offset outside the .impg file");
  Reader r{d.data() + 16, d.data() + d.size()};
  // SequenceIndex: the three maps are redundant; ids need not be dense in principle, names / lengths are kept by id
  std::map<uint32_t, std::string> id_name;
  std::map<uint32_t, uint64_t> id_len;
  for (uint64_t n = r.varint(); n > 0; n--) {
    std::string name = r.str();
    id_name[r.u32()] = std::move(name);
  }
  for (uint64_t n = r.varint(); n > 0; n--) {
    const uint32_t id = r.u32();
    std::string name = r.str();
    auto it = id_name.find(id);
    REQUIRE(it != id_name.end() && it->second == name, IMPGX_E_PARSE, "name_to_id and id_to_name disagree in .impg file");
  }
  for (uint64_t n = r.varint(); n > 0; n--) {
    const uint32_t id = r.u32();
    id_len[id] = r.varint();
  }
  const uint32_t next_id = r.u32();
  // a crafted next_id must not size the tables: every id below it that is in use costs bytes of this file
  REQUIRE((uint64_t)next_id <= d.size(), IMPGX_E_PARSE, "next_id larger than the .impg file itself");
  f->names.assign(next_id, std::string());
  f->lens.assign(next_id, 0);
  for (auto &kv : id_name) {
    REQUIRE(kv.first < next_id, IMPGX_E_PARSE, "sequence id beyond next_id in .impg file");
    f->names[kv.first] = kv.second;
  }
  for (auto &kv : id_len) {
    REQUIRE(kv.first < next_id, IMPGX_E_PARSE, "sequence id beyond next_id in .impg file");
    f->lens[kv.first] = kv.second;
  }
  // forest map, then every tree at its offset (load_tree_from_disk, :1725-1775)
  Reader fm{d.data() + forest_off, d.data() + d.size()};
  std::map<uint32_t, uint64_t> forest;
  for (uint64_t n = fm.varint(); n > 0; n--) {
    const uint32_t t = fm.u32();
    forest[t] = fm.varint();
  }
  for (auto &kv : forest) {
    REQUIRE(kv.second >= 16 && kv.second < forest_off, IMPGX_E_PARSE, "tree offset outside the .impg file");
    Reader t{d.data() + kv.second, d.data() + forest_off};
    const uint32_t target_id = t.u32();
    REQUIRE(target_id == kv.first, IMPGX_E_PARSE, "Tree mismatch in .impg file");
    REQUIRE(target_id < next_id, IMPGX_E_PARSE, "target id beyond next_id in .impg file");
    for (uint64_t n = t.varint(); n > 0; n--) {
      Entry en;
      en.target_id = target_id;
      en.first = t.i32();
      en.last = t.i32();
      en.query_id = t.u32();
      en.target_start = t.i32();
      en.target_end = t.i32();
      en.query_start = t.i32();
      en.query_end = t.i32();
      en.file_index = t.u32();
      en.strand_and_data_offset = t.varint();
      en.data_bytes = t.varint();
      REQUIRE(en.query_id < next_id, IMPGX_E_PARSE, "query id beyond next_id in .impg file");
      f->entries.push_back(en);
    }
  }
  // the alignments = the entries that are not the reversed copy (:1562-1605), back in the order the
  // alignment files list them: ties on `first` inside a tree are resolved by that order (coitrees'
  // stable sort over the insertion order, which a reloaded tree inherits from the in-order dump)
  for (size_t i = 0; i < f->entries.size(); i++) {
    if (!(f->entries[i].strand_and_data_offset & REVERSED_BIT)) f->forward.push_back(i);
    else f->bidirectional = true;
  }
  std::stable_sort(f->forward.begin(), f->forward.end(), [&](size_t a, size_t b) {
    const Entry &x = f->entries[a], &y = f->entries[b];
    if (x.file_index != y.file_index) return x.file_index < y.file_index;
    return (x.strand_and_data_offset & ~(STRAND_BIT | REVERSED_BIT)) < (y.strand_and_data_offset & ~(STRAND_BIT | REVERSED_BIT));
  });
  return f.release();
}

static impgx_record record_of(const Entry &e) {
  impgx_record r;
  r.query_id = e.query_id;
  r.target_id = e.target_id;
  r.query_start = e.query_start;
  r.query_end = e.query_end;
  r.target_start = e.target_start;
  r.target_end = e.target_end;
  r.strand = (e.strand_and_data_offset & STRAND_BIT) ? 1 : 0;
  r.reserved = 0;
  return r;
}

// ---- BGZF alignment files behind an index (src/paf.rs:47-114, :199-302). The reference treats a path that ends
// in .gz / .bgz as compressed, insists on BGZF (a plain gzip file is an error) and addresses a CIGAR by the BGZF
// VIRTUAL POSITION of its first byte: (offset of the compressed block in the file) << 16 | (offset inside the
// inflated block). A position at the very end of a block is reported as offset 0 of the next block (noodles'
// Block::virtual_position; the GZI lookup of parse_paf_bgzf_with_gzi picks the same block).
static bool has_compressed_suffix(const std::string &p) {
  auto ends = [&](const char *sfx) {
    const size_t k = strlen(sfx);
    return p.size() >= k && p.compare(p.size() - k, k, sfx) == 0;
  };
  return ends(".gz") || ends(".bgz");
}
struct BgzfFile {
  std::vector<uint64_t> coff, ustart;  // per block, ascending; a final sentinel holds (file size, inflated size)
  std::vector<uint8_t> text;           // the inflated file
  // virtual position -> offset in `text`
  uint64_t to_offset(uint64_t vpos, const std::string &path) const {
    const uint64_t c = vpos >> 16, w = vpos & 0xffffu;
    auto it = std::lower_bound(coff.begin(), coff.end() - 1, c);
    REQUIRE(it != coff.end() - 1 && *it == c, IMPGX_E_PARSE,
            "virtual position " + std::to_string(vpos) + " does not start at a BGZF block of '" + path + "'");
    const size_t b = (size_t)(it - coff.begin());
    REQUIRE(w < ustart[b + 1] - ustart[b], IMPGX_E_PARSE,
            "virtual position " + std::to_string(vpos) + " lies beyond its BGZF block of '" + path + "'");
    return ustart[b] + w;
  }
  // offset in `text` -> virtual position: the block that holds the byte (the later one at a block border)
  uint64_t to_vpos(uint64_t off) const {
    size_t b = (size_t)(std::upper_bound(ustart.begin(), ustart.end() - 1, off) - ustart.begin());
    b = b ? b - 1 : 0;
    return (coff[b] << 16) | (off - ustart[b]);
  }
};
static void load_bgzf(const std::string &path, BgzfFile &out) {
  const std::vector<uint8_t> d = slurp(path);
  size_t p = 0;
  while (p < d.size()) {
    // is_bgzf (src/paf.rs:47-66): gzip, DEFLATE, FEXTRA, XLEN = 6, subfield 'B' 'C' of length 2
    const bool hdr = d.size() - p >= 18 && d[p] == 0x1f && d[p + 1] == 0x8b && d[p + 2] == 0x08 && (d[p + 3] & 0x04) &&
                     d[p + 10] == 6 && d[p + 11] == 0 && d[p + 12] == 'B' && d[p + 13] == 'C' && d[p + 14] == 2 && d[p + 15] == 0;
    REQUIRE(hdr, IMPGX_E_PARSE,
            p == 0 ? "'" + path + "' is regular gzip, not BGZF. Convert with: zcat '" + path + "' | bgzip > output.paf.gz"
                   : "damaged BGZF block header in '" + path + "'");
    const size_t bsize = (size_t)(d[p + 16] | (d[p + 17] << 8)) + 1;
    REQUIRE(bsize >= 26 && bsize <= d.size() - p, IMPGX_E_PARSE, "truncated BGZF block in '" + path + "'");
    const uint32_t isize = (uint32_t)d[p + bsize - 4] | ((uint32_t)d[p + bsize - 3] << 8) | ((uint32_t)d[p + bsize - 2] << 16) |
                           ((uint32_t)d[p + bsize - 1] << 24);
    REQUIRE(isize <= 65536, IMPGX_E_PARSE, "BGZF block inflates to more than 64 KiB in '" + path + "'");
    out.coff.push_back(p);
    out.ustart.push_back(out.text.size());
    if (isize) {
      const size_t at = out.text.size();
      out.text.resize(at + isize);
      z_stream zs;
      memset(&zs, 0, sizeof(zs));
      REQUIRE(inflateInit2(&zs, -15) == Z_OK, IMPGX_E_IO, "zlib initialisation failed");
      zs.next_in = const_cast<Bytef *>(d.data() + p + 18);
      zs.avail_in = (uInt)(bsize - 18 - 8);
      zs.next_out = out.text.data() + at;
      zs.avail_out = isize;
      const int rc = inflate(&zs, Z_FINISH);
      const bool ok = rc == Z_STREAM_END && zs.avail_out == 0;
      inflateEnd(&zs);
      REQUIRE(ok, IMPGX_E_PARSE, "BGZF block of '" + path + "' does not inflate");
    }
    p += bsize;
  }
  out.coff.push_back(d.size());
  out.ustart.push_back(out.text.size());
}

static bool looks_compressed(const std::string &path) {
  FILE *f = fopen(path.c_str(), "rb");
  REQUIRE(f != nullptr, IMPGX_E_IO, "cannot open alignment file '" + path + "'");
  unsigned char m[2] = {0, 0};
  const size_t k = fread(m, 1, 2, f);
  fclose(f);
  return k == 2 && m[0] == 0x1f && m[1] == 0x8b;
}

// Impg::serialize_with_forest_map over the entries from_multi_alignment_records would build
// (forward under the target, reversed copy under the query unless self or unidirectional, :1562-1605)
static void write_impg(const PafData &pd, bool bidirectional, const std::string &out_path) {
  const uint32_t n_seqs = (uint32_t)pd.names.size();
  std::vector<std::vector<Entry>> trees(n_seqs);
  for (size_t i = 0; i < pd.recs.size(); i++) {
    const impgx_record &r = pd.recs[i];
    Entry f;
    f.target_id = r.target_id;
    f.first = r.target_start;
    f.last = r.target_end;
    f.query_id = r.query_id;
    f.target_start = r.target_start;
    f.target_end = r.target_end;
    f.query_start = r.query_start;
    f.query_end = r.query_end;
    f.file_index = pd.file_idx[i];
    f.strand_and_data_offset = pd.cg_off[i] | (r.strand ? STRAND_BIT : 0);
    f.data_bytes = pd.cg_len[i];
    trees[r.target_id].push_back(f);
    if (bidirectional && r.query_id != r.target_id) {
      Entry v = f;  // coordinates swapped, REVERSED bit set
      v.target_id = r.query_id;
      v.first = r.query_start;
      v.last = r.query_end;
      v.query_id = r.target_id;
      v.target_start = r.query_start;
      v.target_end = r.query_end;
      v.query_start = r.target_start;
      v.query_end = r.target_end;
      v.strand_and_data_offset |= REVERSED_BIT;
      trees[r.query_id].push_back(v);
    }
  }
  Writer w;
  // the reference always writes the V2 magic (:1659); a unidirectional index simply holds no reversed entries
  w.b.insert(w.b.end(), {'I', 'M', 'P', 'G', 'I', 'D', 'X', '2'});
  w.fixed(0, 8);
  w.varint(n_seqs);  // name_to_id (a hash map in the reference: any order decodes; here by id)
  for (uint32_t s = 0; s < n_seqs; s++) {
    w.str(pd.names[s]);
    w.varint(s);
  }
  w.varint(n_seqs);  // id_to_name
  for (uint32_t s = 0; s < n_seqs; s++) {
    w.varint(s);
    w.str(pd.names[s]);
  }
  w.varint(n_seqs);  // id_to_len
  for (uint32_t s = 0; s < n_seqs; s++) {
    w.varint(s);
    w.varint(pd.lens[s]);
  }
  w.varint(n_seqs);  // next_id
  std::vector<std::pair<uint32_t, uint64_t>> forest;
  for (uint32_t t = 0; t < n_seqs; t++) {
    if (trees[t].empty()) continue;
    // tree.iter() is in-order = sorted by `first`, equal keys in insertion order
    std::stable_sort(trees[t].begin(), trees[t].end(), [](const Entry &a, const Entry &b) { return a.first < b.first; });
    forest.push_back({t, (uint64_t)w.b.size()});
    w.varint(t);
    w.varint(trees[t].size());
    for (auto &e : trees[t]) {
      w.i32(e.first);
      w.i32(e.last);
      w.varint(e.query_id);
      w.i32(e.target_start);
      w.i32(e.target_end);
      w.i32(e.query_start);
      w.i32(e.query_end);
      w.varint(e.file_index);
      w.varint(e.strand_and_data_offset);
      w.varint(e.data_bytes);
    }
  }
  const uint64_t forest_off = w.b.size();
  w.varint(forest.size());
  for (auto &kv : forest) {
    w.varint(kv.first);
    w.varint(kv.second);
  }
  for (int i = 0; i < 8; i++) w.b[8 + i] = (uint8_t)(forest_off >> (8 * i));
  FILE *f = fopen(out_path.c_str(), "wb");
  REQUIRE(f != nullptr, IMPGX_E_IO, "cannot write '" + out_path + "'");
  const size_t k = fwrite(w.b.data(), 1, w.b.size(), f);
  const bool bad = k != w.b.size() || fclose(f) != 0;
  REQUIRE(!bad, IMPGX_E_IO, "error while writing '" + out_path + "'");
}

}  // namespace impgx

#define API_BEGIN try {
#define API_END                                      \
  }                                                  \
  catch (const impgx::Error &e) {                    \
    impgx::set_last_error(e.what());                 \
    return e.code;                                   \
  }                                                  \
  catch (const std::bad_alloc &) {                   \
    impgx::set_last_error("host allocation failed"); \
    return IMPGX_E_NOMEM;                            \
  }                                                  \
  catch (const std::exception &e) {                  \
    impgx::set_last_error(e.what());                 \
    return IMPGX_E_INVALID;                          \
  }                                                  \
  return IMPGX_OK;

extern "C" {

int impgx_impg_open(const char *path, impgx_impg **out) {
  API_BEGIN
  REQUIRE(path && out, IMPGX_E_INVALID, "NULL argument");
  *out = nullptr;
  *out = impgx::open_impg(path);
  API_END
}
void impgx_impg_close(impgx_impg *f) { delete f; }
int impgx_impg_version(const impgx_impg *f) { return f ? f->version : 0; }
int impgx_impg_bidirectional(const impgx_impg *f) { return (f && f->bidirectional) ? 1 : 0; }
uint32_t impgx_impg_num_seqs(const impgx_impg *f) { return f ? (uint32_t)f->names.size() : 0; }
const char *impgx_impg_seq_name(const impgx_impg *f, uint32_t id) {
  return (f && id < f->names.size()) ? f->names[id].c_str() : nullptr;
}
uint64_t impgx_impg_seq_len(const impgx_impg *f, uint32_t id) { return (f && id < f->lens.size()) ? f->lens[id] : 0; }
uint64_t impgx_impg_num_entries(const impgx_impg *f) { return f ? f->entries.size() : 0; }
uint64_t impgx_impg_num_records(const impgx_impg *f) { return f ? f->forward.size() : 0; }

int impgx_impg_records(const impgx_impg *f, impgx_record *records, uint32_t *file_index, uint64_t *data_offset,
                       uint64_t *data_bytes) {
  if (!f) return IMPGX_E_INVALID;
  for (size_t k = 0; k < f->forward.size(); k++) {
    const impgx::Entry &e = f->entries[f->forward[k]];
    if (records) records[k] = impgx::record_of(e);
    if (file_index) file_index[k] = e.file_index;
    if (data_offset) data_offset[k] = e.strand_and_data_offset & ~(impgx::STRAND_BIT | impgx::REVERSED_BIT);
    if (data_bytes) data_bytes[k] = e.data_bytes;
  }
  return IMPGX_OK;
}

int impgx_impg_write(const char *const *paf_paths, size_t n_paths, int bidirectional, const char *out_path) {
  API_BEGIN
  REQUIRE(paf_paths && n_paths >= 1 && out_path, IMPGX_E_INVALID, "NULL argument");
  impgx::PafData pd;
  for (size_t k = 0; k < n_paths; k++) {
    REQUIRE(paf_paths[k], IMPGX_E_INVALID, "NULL path");
    const std::string path = paf_paths[k];
    const bool bgzf = impgx::has_compressed_suffix(path);  // parse_paf_file (src/paf.rs:306-362) goes by the suffix
    REQUIRE(bgzf || !impgx::looks_compressed(path), IMPGX_E_UNSUPPORTED,
            "'" + path + "' is compressed but not named .gz / .bgz: the reference would read it as plain text");
    const size_t first = pd.recs.size();
    impgx::parse_paf(path, pd);
    if (bgzf) {
      // the parser counted offsets in the inflated text; the index stores the virtual position of every CIGAR
      impgx::BgzfFile bz;
      impgx::load_bgzf(path, bz);
      for (size_t i = first; i < pd.recs.size(); i++) {
        REQUIRE(pd.cg_off[i] < bz.text.size(), IMPGX_E_PARSE, "CIGAR offset beyond the inflated size of '" + path + "'");
        pd.cg_off[i] = bz.to_vpos(pd.cg_off[i]);
      }
    }
  }
  impgx::write_impg(pd, bidirectional != 0, out_path);
  API_END
}

int impgx_index_from_impg(const char *impg_path, const char *const *alignment_files, size_t n_files, int device,
                          impgx_index **out) {
  API_BEGIN
  REQUIRE(impg_path && alignment_files && n_files >= 1 && out, IMPGX_E_INVALID, "NULL argument");
  *out = nullptr;
  impgx::check_device(device);
  std::unique_ptr<impgx_impg> f(impgx::open_impg(impg_path));
  std::vector<std::vector<uint8_t>> text(n_files);
  std::vector<std::unique_ptr<impgx::BgzfFile>> bgzf(n_files);
  std::vector<char> loaded(n_files, 0);
  std::vector<impgx_record> recs;
  std::vector<uint32_t> runs;
  std::vector<uint64_t> run_off{0};
  recs.reserve(f->forward.size());
  for (size_t k : f->forward) {
    const impgx::Entry &e = f->entries[k];
    REQUIRE(e.file_index < n_files, IMPGX_E_INVALID, "the index refers to alignment file " + std::to_string(e.file_index) +
                                                          " but only " + std::to_string(n_files) + " were given");
    if (!loaded[e.file_index]) {
      const std::string p = alignment_files[e.file_index] ? alignment_files[e.file_index] : "";
      const bool tp = p.size() > 5 && (p.rfind(".1aln") == p.size() - 5 || p.rfind(".tpa") == p.size() - 4);
      REQUIRE(!tp, IMPGX_E_UNSUPPORTED, "tracepoint alignment files (.1aln / .tpa) are outside the path");
      if (impgx::has_compressed_suffix(p)) {  // read_cigar_data (src/paf.rs:68-114): BGZF, offsets are virtual positions
        bgzf[e.file_index].reset(new impgx::BgzfFile());
        impgx::load_bgzf(p, *bgzf[e.file_index]);
      } else {
        REQUIRE(!impgx::looks_compressed(p), IMPGX_E_UNSUPPORTED,
                "'" + p + "' is compressed but not named .gz / .bgz: the reference would read it as plain text");
        text[e.file_index] = impgx::slurp(p);
      }
      loaded[e.file_index] = 1;
    }
    const bool bz = bgzf[e.file_index] != nullptr;
    const std::vector<uint8_t> &t = bz ? bgzf[e.file_index]->text : text[e.file_index];
    uint64_t off = e.strand_and_data_offset & ~(impgx::STRAND_BIT | impgx::REVERSED_BIT);
    if (bz) off = bgzf[e.file_index]->to_offset(off, alignment_files[e.file_index]);
    REQUIRE(off <= t.size() && e.data_bytes <= t.size() - off, IMPGX_E_PARSE,
            "CIGAR offset beyond the end of '" + std::string(alignment_files[e.file_index]) + "' (is it the file the index was built from?)");
    const long n = impgx::parse_cigar((const char *)t.data() + off, (size_t)e.data_bytes, runs);
    REQUIRE(n > 0, IMPGX_E_PARSE, "no valid CIGAR at the recorded offset of '" + std::string(alignment_files[e.file_index]) + "'");
    recs.push_back(impgx::record_of(e));
    run_off.push_back(runs.size());
  }
  impgx_index *idx = impgx::index_build(recs.data(), recs.size(), runs.data(), run_off.data(), f->lens.data(),
                                        (uint32_t)f->lens.size(), f->bidirectional, device);
  idx->names = f->names;
  for (uint32_t s = 0; s < f->names.size(); s++) idx->name_to_id.emplace(f->names[s], s);
  *out = idx;
  API_END
}

}  // extern "C"
