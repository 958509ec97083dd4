// host_output.cu — BEDPE / PAF output of a raw result set: the CIGAR-faithful
// merge and the two writers of the reference's query driver, on the host.
//
//   merge_adjusted_intervals          src/main.rs:12563-12845
//   merge_adjusted_intervals_gap_2d   src/main.rs:12858-13011 (used when a CIGAR is empty)
//   CIGAR helpers                     src/main.rs:13014-13180 (f32 scaling + truncation kept in IEEE f32)
//   output_results_bedpe / _paf       src/main.rs:11894-12103
//
// The device produces the raw AdjustedIntervals with their clipped CIGARs
// (impgx_query_batch with store_cigar); per-row text work stays on the host
// (SURVEY.md H7). No device code in this file.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <string>

#include "engine.cuh"

namespace impgx {
namespace {

struct Row {
  uint32_t q_id, t_id;
  int32_t q_first, q_last, t_first, t_last;
  std::vector<uint32_t> cg;
};

inline uint32_t op_of(uint32_t v) { return v >> 29; }
inline int32_t len_of(uint32_t v) { return (int32_t)(v & 0x1fffffffu); }
inline uint32_t mk(int32_t len, uint32_t op) { return (op << 29) | (uint32_t)len; }
inline int32_t tdelta(uint32_t v) { return op_of(v) == IMPGX_OP_I ? 0 : len_of(v); }
inline int32_t qdelta_abs(uint32_t v) { return op_of(v) == IMPGX_OP_D ? 0 : len_of(v); }

void merge_consecutive(std::vector<uint32_t> &c) {
  if (c.size() <= 1) return;
  size_t w = 0;
  for (size_t r = 1; r < c.size(); r++) {
    // a combined length of 2^29 or more would spill into the op bits: such runs stay apart
    if (op_of(c[w]) == op_of(c[r]) && (int64_t)len_of(c[w]) + len_of(c[r]) < (1ll << 29))
      c[w] = mk(len_of(c[w]) + len_of(c[r]), op_of(c[w]));
    else c[++w] = c[r];
  }
  c.resize(w + 1);
}

std::vector<uint32_t> cigar_suffix(const std::vector<uint32_t> &c, int32_t qlen) {
  std::vector<uint32_t> out;
  int32_t rem = qlen;
  for (size_t i = c.size(); i-- > 0;) {
    if (rem <= 0) break;
    int32_t qd = qdelta_abs(c[i]);
    if (qd <= rem) {
      out.push_back(c[i]);
      rem -= qd;
    } else if (qd > 0) {
      float scale = (float)rem / (float)qd;
      out.push_back(mk((int32_t)((float)len_of(c[i]) * scale), op_of(c[i])));
      rem = 0;
    }
  }
  std::reverse(out.begin(), out.end());
  return out;
}

std::vector<uint32_t> cigar_prefix(const std::vector<uint32_t> &c, int32_t qlen) {
  std::vector<uint32_t> out;
  int32_t rem = qlen;
  for (uint32_t v : c) {
    if (rem <= 0) break;
    int32_t qd = qdelta_abs(v);
    if (qd <= rem) {
      out.push_back(v);
      rem -= qd;
    } else if (qd > 0) {
      float scale = (float)rem / (float)qd;
      out.push_back(mk((int32_t)((float)len_of(v) * scale), op_of(v)));
      rem = 0;
    }
  }
  return out;
}

std::vector<uint32_t> trim_prefix(const std::vector<uint32_t> &c, int32_t qlen, int32_t tlen) {
  std::vector<uint32_t> out;
  int32_t qc = 0, tc = 0;
  size_t start = 0;
  for (size_t i = 0; i < c.size(); i++) {
    int32_t qd = qdelta_abs(c[i]), td = tdelta(c[i]);
    if (qc + qd > qlen || tc + td > tlen) {
      int32_t qr = qlen - qc, tr = tlen - tc;
      float ratio;
      if (qd > 0 && td > 0) ratio = std::min((float)qr / (float)qd, (float)tr / (float)td);
      else if (qd > 0) ratio = (float)qr / (float)qd;
      else if (td > 0) ratio = (float)tr / (float)td;
      else ratio = 0.0f;
      int32_t skip = (int32_t)((float)len_of(c[i]) * ratio);
      if (skip < len_of(c[i])) out.push_back(mk(len_of(c[i]) - skip, op_of(c[i])));
      start = i + 1;
      break;
    }
    qc += qd;
    tc += td;
    if (qc >= qlen && tc >= tlen) {
      start = i + 1;
      break;
    }
  }
  out.insert(out.end(), c.begin() + start, c.end());
  return out;
}

void prepend(std::vector<uint32_t> &cur, const std::vector<uint32_t> &front, const std::vector<uint32_t> &mid) {
  std::vector<uint32_t> n;
  n.reserve(front.size() + mid.size() + cur.size());
  n.insert(n.end(), front.begin(), front.end());
  n.insert(n.end(), mid.begin(), mid.end());
  n.insert(n.end(), cur.begin(), cur.end());
  cur.swap(n);
}

// src/main.rs:12563-12845
void merge_with_cigars(std::vector<Row> &rows, int32_t d) {
  if (!(rows.size() > 1 && d >= 0)) return;
  std::stable_sort(rows.begin(), rows.end(), [](const Row &a, const Row &b) {
    bool af = a.q_first < a.q_last, bf = b.q_first < b.q_last;
    int32_t ap = af ? a.q_first : a.q_last, bp = bf ? b.q_first : b.q_last;
    if (a.q_id != b.q_id) return a.q_id < b.q_id;
    if (af != bf) return !af;  // false sorts before true
    if (ap != bp) return ap < bp;
    if (a.t_id != b.t_id) return a.t_id < b.t_id;
    return a.t_first < b.t_first;
  });
  std::vector<Row> out;
  out.reserve(rows.size());
  Row cur = std::move(rows[0]);
  for (size_t k = 1; k < rows.size(); k++) {
    Row nx = std::move(rows[k]);
    const bool qf = cur.q_first <= cur.q_last, nqf = nx.q_first <= nx.q_last;
    if (cur.q_id != nx.q_id || cur.t_id != nx.t_id || qf != nqf) {
      out.push_back(std::move(cur));
      cur = std::move(nx);
      continue;
    }
    bool qc, tc, qo, to;
    if (qf) {
      qc = cur.q_last == nx.q_first; tc = cur.t_last == nx.t_first;
      qo = cur.q_last > nx.q_first; to = cur.t_last > nx.t_first;
    } else {
      qc = cur.q_first == nx.q_last; tc = cur.t_first == nx.t_last;
      qo = cur.q_first > nx.q_last; to = cur.t_first < nx.t_last;
    }
    if (qc && tc) {
      if (qf) {
        cur.q_last = nx.q_last; cur.t_last = nx.t_last;
        cur.cg.insert(cur.cg.end(), nx.cg.begin(), nx.cg.end());
      } else {
        cur.q_first = nx.q_first; cur.t_first = nx.t_first;
        prepend(cur.cg, nx.cg, {});
      }
      merge_consecutive(cur.cg);
      continue;
    }
    if (qo && to) {
      int32_t qol = qf ? nx.q_first - cur.q_last : nx.q_last - cur.q_first;
      int32_t tol = qf ? nx.t_first - cur.t_last : cur.t_first - nx.t_last;
      if (qol > 0 && tol > 0 && cigar_suffix(cur.cg, qol) == cigar_prefix(nx.cg, qol)) {
        std::vector<uint32_t> trimmed = trim_prefix(nx.cg, qol, tol);
        if (qf) {
          cur.q_last = nx.q_last; cur.t_last = nx.t_last;
          cur.cg.insert(cur.cg.end(), trimmed.begin(), trimmed.end());
        } else {
          cur.q_first = nx.q_first; cur.t_first = nx.t_first;
          prepend(cur.cg, trimmed, {});
        }
        continue;
      }
    }
    if (!qo && !to) {
      int32_t qg = qf ? nx.q_first - cur.q_last : cur.q_first - nx.q_last;
      int32_t tg = qf ? nx.t_first - cur.t_last : cur.t_first - nx.t_last;
      if (qg >= 0 && tg >= 0 && (qg > 0 || tg > 0) && qg <= d && tg <= d) {
        std::vector<uint32_t> gap;
        if (qg > 0) gap.push_back(mk(qg, IMPGX_OP_I));
        if (tg > 0) gap.push_back(mk(tg, IMPGX_OP_D));
        if (qf) {
          cur.q_last = nx.q_last; cur.t_last = nx.t_last;
          cur.cg.insert(cur.cg.end(), gap.begin(), gap.end());
          cur.cg.insert(cur.cg.end(), nx.cg.begin(), nx.cg.end());
        } else {
          cur.q_first = nx.q_first; cur.t_first = nx.t_first;
          prepend(cur.cg, nx.cg, gap);
        }
        merge_consecutive(cur.cg);
        continue;
      }
    }
    out.push_back(std::move(cur));
    cur = std::move(nx);
  }
  out.push_back(std::move(cur));
  rows.swap(out);
}

// src/main.rs:12858-13011 (host form; the device has its own in merge_kernels.cuh)
void merge_gap_2d(std::vector<Row> &rows, int32_t dist) {
  if (rows.size() <= 1 || dist < 0) return;
  const int64_t d = dist;
  const size_t n = rows.size();
  std::map<std::tuple<uint32_t, uint32_t, bool>, std::vector<size_t>> groups;
  for (size_t i = 0; i < n; i++) groups[{rows[i].q_id, rows[i].t_id, rows[i].q_first <= rows[i].q_last}].push_back(i);
  std::vector<size_t> parent(n);
  for (size_t i = 0; i < n; i++) parent[i] = i;
  auto find = [&](size_t x) {
    while (parent[x] != x) {
      parent[x] = parent[parent[x]];
      x = parent[x];
    }
    return x;
  };
  for (auto &kv : groups) {
    const bool fwd = std::get<2>(kv.first);
    auto &ix = kv.second;
    std::stable_sort(ix.begin(), ix.end(), [&](size_t a, size_t b) {
      int64_t ka = fwd ? rows[a].q_first : -(int64_t)rows[a].q_first, kb = fwd ? rows[b].q_first : -(int64_t)rows[b].q_first;
      return ka < kb;
    });
    for (size_t ap = 0; ap < ix.size(); ap++) {
      const Row &A = rows[ix[ap]];
      const int64_t qa_s = fwd ? A.q_first : A.q_last, qa_e = fwd ? A.q_last : A.q_first;
      for (size_t bp = ap + 1; bp < ix.size(); bp++) {
        const Row &B = rows[ix[bp]];
        const int64_t qb_s = fwd ? B.q_first : B.q_last;
        if (qb_s < qa_s) continue;
        if (qb_s - qa_e > d) break;
        int64_t tg;
        bool tf;
        if (fwd) {
          tg = (int64_t)B.t_first - A.t_last;
          tf = B.t_first > A.t_first;
        } else {
          tg = (int64_t)A.t_first - B.t_last;
          tf = B.t_last < A.t_last;
        }
        if (!tf || tg > d) continue;
        size_t ra = find(ix[ap]), rb = find(ix[bp]);
        if (ra != rb) parent[ra] = rb;
      }
    }
  }
  std::map<size_t, std::vector<size_t>> buckets;
  for (size_t i = 0; i < n; i++) buckets[find(i)].push_back(i);
  std::vector<Row> out;
  std::vector<char> taken(n, 0);
  for (size_t i = 0; i < n; i++) {
    if (taken[i]) continue;
    auto it = buckets.find(find(i));
    if (it == buckets.end()) continue;
    std::vector<size_t> members = std::move(it->second);
    buckets.erase(it);
    for (size_t m : members) taken[m] = 1;
    const bool fwd = rows[members[0]].q_first <= rows[members[0]].q_last;
    std::stable_sort(members.begin(), members.end(), [&](size_t a, size_t b) {
      int64_t ka = fwd ? rows[a].q_first : -(int64_t)rows[a].q_first, kb = fwd ? rows[b].q_first : -(int64_t)rows[b].q_first;
      return ka < kb;
    });
    Row r = rows[members[0]];
    r.cg.clear();
    for (size_t m : members) {
      const Row &x = rows[m];
      if (fwd) {
        r.q_first = std::min(r.q_first, x.q_first);
        r.q_last = std::max(r.q_last, x.q_last);
      } else {
        r.q_first = std::max(r.q_first, x.q_first);
        r.q_last = std::min(r.q_last, x.q_last);
      }
      r.t_first = std::min(r.t_first, x.t_first);
      r.t_last = std::max(r.t_last, x.t_last);
      r.cg.insert(r.cg.end(), x.cg.begin(), x.cg.end());
    }
    merge_consecutive(r.cg);
    out.push_back(std::move(r));
  }
  rows.swap(out);
}

// Rust `{:.6}` of an f32 followed by trim_end_matches('0').trim_end_matches('.')
std::string f32_trim(float v) {
  if (std::isnan(v)) return "NaN";
  if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
  char buf[64];
  snprintf(buf, sizeof buf, "%.6f", (double)v);
  std::string s(buf);
  while (!s.empty() && s.back() == '0') s.pop_back();
  while (!s.empty() && s.back() == '.') s.pop_back();
  return s;
}

struct Counts {
  int32_t m = 0, x = 0, ni = 0, ibp = 0, nd = 0, dbp = 0, bl = 0;
};
Counts count(const std::vector<uint32_t> &c) {
  Counts k;
  for (uint32_t v : c) {
    int32_t l = len_of(v);
    switch (op_of(v)) {
      case IMPGX_OP_M:
      case IMPGX_OP_EQ: k.m += l; k.bl += l; break;
      case IMPGX_OP_X: k.x += l; k.bl += l; break;
      case IMPGX_OP_I: k.ni += 1; k.ibp += l; k.bl += l; break;
      case IMPGX_OP_D: k.nd += 1; k.dbp += l; k.bl += l; break;
    }
  }
  return k;
}

std::string name_of(const impgx_index *idx, uint32_t id) {
  if (id < idx->names.size() && !idx->names[id].empty()) return idx->names[id];
  return "seq" + std::to_string(id);
}

std::vector<Row> rows_of(const impgx_results *res, size_t row) {
  std::vector<Row> v;
  // the driver drops result[0] (the input range) before BEDPE / PAF output, src/main.rs:7474,7486
  for (uint64_t i = res->row_off[row] + 1; i < res->row_off[row + 1]; i++) {
    Row r{res->qid[i], res->tid[i], res->qf[i], res->ql[i], res->tf[i], res->tl[i], {}};
    if (res->has_cigar) r.cg.assign(res->cig.begin() + res->cig_off[i], res->cig.begin() + res->cig_off[i + 1]);
    v.push_back(std::move(r));
  }
  return v;
}

}  // namespace

// format: 1 = bedpe (src/main.rs:11894-11987), 2 = paf (src/main.rs:11989-12103)
// parse_subsequence_coordinates (src/main.rs:4642-4659): split at the LAST ':', the range part must hold a '-'
// and an i32 before it; pinned by the reference's test (src/main.rs:13330-13346)
bool to_original_coordinates(const std::string &seq_name, std::string &base, uint32_t &offset) {
  const size_t colon = seq_name.rfind(':');
  if (colon == std::string::npos) return false;
  const std::string range = seq_name.substr(colon + 1);
  const size_t dash = range.find('-');
  if (dash == std::string::npos) return false;
  const std::string st = range.substr(0, dash);  // str::parse::<i32>: optional sign, digits only, in range
  size_t k = (!st.empty() && (st[0] == '+' || st[0] == '-')) ? 1 : 0;
  if (k >= st.size()) return false;
  int64_t v = 0;
  for (size_t i = k; i < st.size(); i++) {
    if (st[i] < '0' || st[i] > '9') return false;
    v = v * 10 + (st[i] - '0');
    if (v > 2147483648ll) return false;
  }
  if (st[0] == '-') v = -v;
  if (v > 2147483647ll || v < -2147483648ll) return false;
  base = seq_name.substr(0, colon);
  offset = (uint32_t)(int32_t)v;  // `offset as u32`
  return true;
}

static std::string format_row_vector(const impgx_index *idx, std::vector<Row> rows, const char *name, int32_t d, int format);

std::string format_rows(const impgx_index *idx, const impgx_results *res, size_t row, const char *name, int32_t d,
                        int format) {
  return format_row_vector(idx, rows_of(res, row), name, d, format);
}

// Test hook: the same merge + writer on rows given as host arrays (no result object, no device): row i =
// (q_id, q_first, q_last, t_id, t_first, t_last, cigar runs [cig_off[i], cig_off[i+1])); names / lengths by id.
std::string debug_format_rows(const char *const *names, const uint64_t *lens, uint32_t n_seqs, size_t n,
                              const uint32_t *q_id, const int32_t *q_first, const int32_t *q_last, const uint32_t *t_id,
                              const int32_t *t_first, const int32_t *t_last, const uint64_t *cig_off, const uint32_t *cig,
                              const char *name, int32_t d, int format, bool original_coordinates) {
  impgx_index idx;  // host fields only; nothing of it lives on a device
  idx.n_seqs = n_seqs;
  idx.original_coordinates = original_coordinates;
  for (uint32_t s = 0; s < n_seqs; s++) {
    idx.names.push_back(names && names[s] ? names[s] : "");
    idx.seq_lens.push_back(lens ? lens[s] : 0);
  }
  std::vector<Row> rows;
  for (size_t i = 0; i < n; i++) {
    REQUIRE(q_id[i] < n_seqs && t_id[i] < n_seqs, IMPGX_E_INVALID, "sequence id out of range");
    Row r{q_id[i], t_id[i], q_first[i], q_last[i], t_first[i], t_last[i], {}};
    if (cig_off) r.cg.assign(cig + cig_off[i], cig + cig_off[i + 1]);
    rows.push_back(std::move(r));
  }
  return format_row_vector(&idx, std::move(rows), name, d, format);
}

static std::string format_row_vector(const impgx_index *idx, std::vector<Row> rows, const char *name, int32_t d, int format) {
  if (format == 1) {
    bool any_empty = false;
    for (auto &r : rows) any_empty |= r.cg.empty();
    if (any_empty) merge_gap_2d(rows, d);
    else merge_with_cigars(rows, d);
  } else {
    merge_with_cigars(rows, d);
  }
  static const char OPS[] = "=XIDM";
  std::string out;
  for (auto &r : rows) {
    int32_t f = r.q_first, l = r.q_last;
    char strand = '+';
    if (f > l) {
      std::swap(f, l);
      strand = '-';
    }
    Counts k = count(r.cg);
    float gi = (float)k.m / (float)(k.m + k.x + k.ni + k.nd);
    float bi = (float)k.m / (float)(k.m + (k.x + k.ibp + k.dbp));
    if (format == 1) {
      // --original-sequence-coordinates (src/main.rs:11920-11934): both sides, each by its own subsequence start
      std::string qn = name_of(idx, r.q_id), tn = name_of(idx, r.t_id), base;
      uint32_t qf = (uint32_t)f, ql = (uint32_t)l, tf = (uint32_t)r.t_first, tl = (uint32_t)r.t_last, off = 0;
      if (idx->original_coordinates && to_original_coordinates(qn, base, off)) {
        qn = base;
        qf += off;
        ql += off;
      }
      if (idx->original_coordinates && to_original_coordinates(tn, base, off)) {
        tn = base;
        tf += off;
        tl += off;
      }
      out += qn + "\t" + std::to_string(qf) + "\t" + std::to_string(ql) + "\t" + tn + "\t" + std::to_string(tf) + "\t" +
             std::to_string(tl) + "\t" + name + "\t0\t" + strand + "\t+\tgi:f:" + f32_trim(gi) + "\tbi:f:" + f32_trim(bi) + "\n";
    } else {
      std::string cg;
      for (uint32_t v : r.cg) {
        cg += std::to_string(len_of(v));
        cg += OPS[op_of(v)];
      }
      out += name_of(idx, r.q_id) + "\t" + std::to_string(idx->seq_lens[r.q_id]) + "\t" + std::to_string((uint32_t)f) +
             "\t" + std::to_string((uint32_t)l) + "\t" + strand + "\t" + name_of(idx, r.t_id) + "\t" +
             std::to_string(idx->seq_lens[r.t_id]) + "\t" + std::to_string((uint32_t)r.t_first) + "\t" +
             std::to_string((uint32_t)r.t_last) + "\t" + std::to_string(k.m) + "\t" + std::to_string(k.bl) +
             "\t255\tgi:f:" + f32_trim(gi) + "\tbi:f:" + f32_trim(bi) + "\tcg:Z:" + cg + "\tan:Z:" + name + "\n";
    }
  }
  return out;
}

}  // namespace impgx
