// refine.cu — SURVEY.md §8f-3: the callers that drive Impg::query_with_cache / populate_cigar_cache
// (reference src/impg.rs:1930-2035), i.e. refine's flank search (src/commands/refine.rs:144-545).
//
// The reference evaluates one candidate interval at a time (rayon over the flanks of one sweep) and keeps a
// per-locus CIGAR cache so that the overlapping candidate queries do not decode the same alignments again.
// With the run stream resident in HBM the cache has no work left; what the device wants instead is a BATCH:
// the search of a locus is a baseline plus three sweeps (left, right with the left flank fixed, left again
// with the right flank fixed), the candidates of one sweep are independent, and so are the loci. Every
// sweep of every locus of the call is therefore answered by ONE impgx::query_batch; the support statistics
// of a candidate (which sequences span both boundaries of the candidate interval) are host code on the
// rows that come back.
#include <cmath>
#include <map>
#include <set>

#include "engine.cuh"

struct impgx_refine_results {
  std::vector<int32_t> refined_start, refined_end, original_start, original_end, left, right;
  std::vector<uint64_t> support, original_support, ent_off{0};
  std::vector<uint32_t> ent_seq;
  std::vector<int32_t> ent_start, ent_end;
  uint64_t candidates = 0, batches = 0;
};

namespace impgx {

// sequences that have an entry in the tree of `target` (compute_max_entities walks tree.iter(), refine.rs:591-632)
__global__ void k_mark_query_ids(const uint32_t *__restrict__ e_qid, uint64_t lo, uint64_t hi, uint8_t *__restrict__ mark) {
  for (uint64_t i = lo + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += (uint64_t)gridDim.x * blockDim.x)
    mark[e_qid[i]] = 1;
}

namespace {

// sweepga::pansn::extract_pansn_key (un-vendored, restated: unpinned). PanSN = sample#haplotype#contig.
std::string pansn_prefix(const std::string &name, uint32_t level) {
  size_t cut = std::string::npos;
  if (level >= 1) {
    cut = name.find('#');
    if (level >= 2 && cut != std::string::npos) cut = name.find('#', cut + 1);
  }
  return level == 0 || cut == std::string::npos ? name : name.substr(0, cut);
}

struct Cand {
  int32_t start = 0, end = 0, left = 0, right = 0;
  uint64_t support = 0;
  std::vector<uint32_t> e_seq;
  std::vector<int32_t> e_start, e_end;
};

// compare_candidates (refine.rs:564-582): more support, then less total extension, then the smaller larger side,
// then the shorter interval
bool better(const Cand &a, const Cand &b) {
  if (a.support != b.support) return a.support > b.support;
  const int64_t ta = (int64_t)a.left + a.right, tb = (int64_t)b.left + b.right;
  if (ta != tb) return ta < tb;
  const int32_t ma = std::max(a.left, a.right), mb = std::max(b.left, b.right);
  if (ma != mb) return ma < mb;
  return (int64_t)a.end - a.start < (int64_t)b.end - b.start;
}

struct Locus {
  uint32_t target = 0;
  int32_t s = 0, e = 0, seq_len = 0;
  std::vector<int32_t> flanks;
  bool capped = false;  // sample / haplotype level: the search stops at max_entities
  uint64_t max_entities = 0;
  bool have_best = false, done = false;
  Cand best;
  uint64_t original_support = 0;
};

struct Piece {
  uint32_t q;
  int32_t qs, qe, ts, te;
};

uint32_t udist(int32_t a, int32_t b) { return a > b ? (uint32_t)a - (uint32_t)b : (uint32_t)b - (uint32_t)a; }

// compute_support_sets (refine.rs:665-783) on the rows of one candidate. The per-sequence pieces are brought
// together by ONE stable sort on (sequence, query start, query end) instead of a hash map of vectors.
void support_of(const impgx_index *idx, const impgx_refine_params &p, const Locus &L, const impgx_results &res, size_t row,
                Cand &c, std::vector<Piece> &buf) {
  c.support = 0;
  c.e_seq.clear(); c.e_start.clear(); c.e_end.clear();
  const uint64_t a = res.row_off[row], b = res.row_off[row + 1];
  if (b - a <= 1) return;
  buf.clear();
  for (uint64_t i = a; i < b; i++) {
    if (res.qid[i] == L.target) continue;
    buf.push_back(Piece{res.qid[i], std::min(res.qf[i], res.ql[i]), std::max(res.qf[i], res.ql[i]),
                        std::min(res.tf[i], res.tl[i]), std::max(res.tf[i], res.tl[i])});
  }
  if (p.merge_distance >= 0)
    std::stable_sort(buf.begin(), buf.end(), [](const Piece &x, const Piece &y) {
      if (x.q != y.q) return x.q < y.q;
      if (x.qs != y.qs) return x.qs < y.qs;
      return x.qe < y.qe;
    });
  else  // --no-merge: the pieces of a sequence keep the result order
    std::stable_sort(buf.begin(), buf.end(), [](const Piece &x, const Piece &y) { return x.q < y.q; });
  const int32_t span = std::min(std::max(c.end - c.start, 0), std::max(p.span_bp, 0));
  const int32_t left_thr = c.start + span, right_thr = c.end - span;
  std::set<std::string> keys;
  struct Ent {
    const std::string *name;
    uint32_t seq;
    int32_t s, e;
  };
  std::vector<Ent> ents;
  size_t i = 0;
  while (i < buf.size()) {
    const uint32_t q = buf[i].q;
    bool have = false;
    int32_t qs = 0, qe = 0;
    Piece cur = buf[i];
    auto close_piece = [&](const Piece &m) {  // covers_boundaries (refine.rs:785-797)
      if (m.ts <= c.start && m.te >= c.end && m.te >= left_thr && m.ts <= right_thr) {
        qs = have ? std::min(qs, m.qs) : m.qs;
        qe = have ? std::max(qe, m.qe) : m.qe;
        have = true;
      }
    };
    for (i++; i < buf.size() && buf[i].q == q; i++) {
      const Piece &nx = buf[i];
      bool join = false;
      if (p.merge_distance >= 0) {  // should_merge (refine.rs:834-850)
        const uint32_t d = (uint32_t)p.merge_distance;
        join = std::min(udist(cur.qe, nx.qs), udist(cur.qs, nx.qe)) <= d ||
               std::min(udist(cur.te, nx.ts), udist(cur.ts, nx.te)) <= d;
      }
      if (join) {
        cur.qs = std::min(cur.qs, nx.qs); cur.qe = std::max(cur.qe, nx.qe);
        cur.ts = std::min(cur.ts, nx.ts); cur.te = std::max(cur.te, nx.te);
      } else {
        close_piece(cur);
        cur = nx;
      }
    }
    close_piece(cur);
    if (!have) continue;
    if (p.blacklist_offsets) {  // closed overlap with any blacklisted (start, end) of the sequence
      bool hit = false;
      for (uint64_t k = p.blacklist_offsets[q]; k < p.blacklist_offsets[q + 1] && !hit; k++)
        hit = p.blacklist_ranges[2 * k] <= qe && qs <= p.blacklist_ranges[2 * k + 1];
      if (hit) continue;
    }
    ents.push_back(Ent{&idx->names[q], q, qs, qe});
    keys.insert(pansn_prefix(idx->names[q], p.support_level));
    // the reference stops as soon as every possible entity is in (its hash map order decides which sequences
    // were listed by then; ascending id here: unpinned)
    if (L.capped && keys.size() >= L.max_entities) break;
  }
  std::sort(ents.begin(), ents.end(), [](const Ent &x, const Ent &y) {
    const int k = x.name->compare(*y.name);
    return k != 0 ? k < 0 : x.s < y.s;
  });
  c.support = keys.size();
  for (auto &en : ents) {
    c.e_seq.push_back(en.seq);
    c.e_start.push_back(en.s);
    c.e_end.push_back(en.e);
  }
}

void check_refine_params(const impgx_index *idx, const impgx_refine_params *p) {
  REQUIRE(idx && p, IMPGX_E_INVALID, "NULL argument");
  REQUIRE(p->span_bp >= 0, IMPGX_E_INVALID, "--span-bp must be >= 0");
  REQUIRE(p->max_extension >= 0.0, IMPGX_E_INVALID, "--max-extension must be >= 0");
  REQUIRE(p->extension_step > 0, IMPGX_E_INVALID, "--extension-step must be > 0");
  REQUIRE(p->support_level <= 2 && p->transitive <= 2, IMPGX_E_INVALID, "unknown support level / traversal");
  REQUIRE(!p->blacklist_offsets || p->blacklist_ranges || p->blacklist_offsets[idx->n_seqs] == 0, IMPGX_E_INVALID,
          "blacklist_ranges is NULL");
  bool named = idx->names.size() == idx->n_seqs;
  for (size_t i = 0; named && i < idx->names.size(); i++) named = !idx->names[i].empty();
  REQUIRE(named, IMPGX_E_INVALID, "refine counts and orders its support by sequence name (impgx_index_set_names / a PAF)");
}

impgx_params query_params(const impgx_refine_params &p) {
  impgx_params q;
  memset(&q, 0, sizeof(q));
  q.mode = p.transitive == 1 ? IMPGX_MODE_BFS : (p.transitive == 2 ? IMPGX_MODE_DFS : IMPGX_MODE_QUERY);
  q.max_depth = p.max_depth;
  q.min_transitive_len = p.min_transitive_len;
  q.min_distance_between_ranges = p.min_distance_between_ranges;
  q.min_output_length = -1;  // "No min_output_length for refine" (refine.rs:502,517)
  q.store_cigar = 0;
  q.min_identity = p.min_identity;
  q.subset_mask = p.subset_mask;
  q.merge_distance = -1;
  q.merge_strands = 0;
  return q;
}

}  // namespace

impgx_refine_results *refine(impgx_index *idx, const impgx_range *loci, size_t n, const impgx_refine_params &p) {
  std::unique_ptr<impgx_refine_results> out(new impgx_refine_results());
  std::vector<Locus> L(n);
  for (size_t i = 0; i < n; i++) {
    Locus &l = L[i];
    REQUIRE(loci[i].target_id < idx->n_seqs, IMPGX_E_INVALID, "locus " + std::to_string(i) + ": target sequence not in the index");
    REQUIRE(loci[i].end > loci[i].start, IMPGX_E_INVALID,
            "locus " + std::to_string(i) + ": invalid range (end must be greater than start)");
    l.target = loci[i].target_id;
    l.s = loci[i].start;
    l.e = loci[i].end;
    l.seq_len = (int32_t)idx->seq_lens[l.target];
    // max_extension_bp (refine.rs:171-179) and the flank grid (build_flanks, :852-877)
    const double raw = p.max_extension <= 1.0 ? std::ceil((double)(l.e - l.s) * p.max_extension) : std::ceil(p.max_extension);
    const int32_t max_ext = (int32_t)std::min(std::max(raw, 0.0), (double)INT32_MAX);
    for (int64_t f = 0; f <= max_ext; f += p.extension_step) {
      l.flanks.push_back((int32_t)f);
      if (max_ext - f < p.extension_step) break;
    }
    if (l.flanks.empty() || l.flanks.back() != max_ext) l.flanks.push_back(max_ext);
    l.capped = p.support_level != 0;
  }
  // max_entities per distinct target (refine.rs:591-632): one device pass marks the sequences of the target's tree
  if (p.support_level != 0) {
    check_device(idx->device);
    std::map<uint32_t, uint64_t> by_target;
    uint8_t *d_mark = nullptr;
    CUDA_CHECK(cudaMalloc((void **)&d_mark, std::max<uint32_t>(idx->n_seqs, 1)));
    std::vector<uint8_t> mark(idx->n_seqs);
    std::vector<uint64_t> off(2);
    for (auto &l : L) {
      auto it = by_target.find(l.target);
      if (it == by_target.end()) {
        CUDA_CHECK(cudaMemcpy(off.data(), idx->d_tgt_off + l.target, 16, cudaMemcpyDeviceToHost));
        CUDA_CHECK(cudaMemset(d_mark, 0, idx->n_seqs));
        if (off[1] > off[0]) {
          k_mark_query_ids<<<(unsigned)std::min<uint64_t>((off[1] - off[0] + 255) / 256, 1184), 256>>>(idx->d_qid, off[0], off[1], d_mark);
          CUDA_CHECK(cudaGetLastError());
        }
        CUDA_CHECK(cudaMemcpy(mark.data(), d_mark, idx->n_seqs, cudaMemcpyDeviceToHost));
        std::set<std::string> keys;
        const std::string own = pansn_prefix(idx->names[l.target], p.support_level);
        for (uint32_t q = 0; q < idx->n_seqs; q++) {
          if (!mark[q] || q == l.target || (p.subset_mask && !p.subset_mask[q])) continue;
          std::string k = pansn_prefix(idx->names[q], p.support_level);
          if (k != own) keys.insert(std::move(k));
        }
        it = by_target.emplace(l.target, keys.size()).first;
      }
      l.max_entities = it->second;
    }
    cudaFree(d_mark);
  }

  const impgx_params q = query_params(p);
  struct Job {
    size_t locus;
    int32_t left, right;
  };
  // phase 0: the baseline (0, 0); 1: left > 0 with right 0; 2: right sweep, left fixed; 3: left sweep, right fixed
  for (int phase = 0; phase < 4; phase++) {
    std::vector<Job> jobs;
    std::vector<impgx_range> rows;
    for (size_t i = 0; i < n; i++) {
      Locus &l = L[i];
      if (l.done) continue;
      auto add = [&](int32_t left, int32_t right) {
        const int32_t s = (int32_t)std::max<int64_t>((int64_t)l.s - left, 0);
        const int32_t e = (int32_t)std::min<int64_t>((int64_t)l.e + right, l.seq_len);
        if (e <= s) return;  // "Skipping non-positive range" (refine.rs:417-423)
        jobs.push_back(Job{i, left, right});
        rows.push_back(impgx_range{l.target, s, e});
      };
      if (phase == 0) add(0, 0);
      else if (phase == 1) {
        for (int32_t f : l.flanks)
          if (f > 0) add(f, 0);
      } else if (phase == 2) {
        for (int32_t f : l.flanks) add(l.have_best ? l.best.left : 0, f);
      } else {
        for (int32_t f : l.flanks) add(f, l.have_best ? l.best.right : 0);
      }
    }
    if (!rows.empty()) {
      std::unique_ptr<impgx_results> res(query_batch(idx, rows.data(), rows.size(), q, /*bed=*/false,
                                                     /*ranges_on_device=*/false, /*results_to_host=*/true, nullptr));
      out->candidates += rows.size();
      out->batches++;
      // the candidates of one locus are consecutive and in flank order: the reduce of a sweep keeps the first of
      // equal candidates, then the sweep's winner meets the best so far (update_best_candidate, refine.rs:548-562).
      // Loci are independent (the reference runs them on rayon workers, refine.rs:116-132): one OpenMP task each.
      std::vector<std::pair<size_t, size_t>> groups;
      for (size_t k = 0; k < jobs.size();) {
        size_t e = k + 1;
        while (e < jobs.size() && jobs[e].locus == jobs[k].locus) e++;
        groups.emplace_back(k, e);
        k = e;
      }
      std::string failure;
      int fail_code = 0;
#pragma omp parallel
      {
        std::vector<Piece> buf;
#pragma omp for schedule(dynamic, 4)
        for (long long gi = 0; gi < (long long)groups.size(); gi++) {
          try {
            const size_t li = jobs[groups[gi].first].locus;
            Locus &l = L[li];
            bool have = false;
            Cand win;
            for (size_t k = groups[gi].first; k < groups[gi].second; k++) {
              Cand c;
              c.start = rows[k].start; c.end = rows[k].end;
              c.left = l.s - c.start; c.right = c.end - l.e;
              support_of(idx, p, l, *res, k, c, buf);
              if (phase == 0) l.original_support = c.support;
              if (!have || better(c, win)) win = std::move(c);
              have = true;
            }
            if (have && (!l.have_best || better(win, l.best))) l.best = std::move(win);
            l.have_best = l.have_best || have;
          } catch (const Error &e) {
#pragma omp critical(impgx_refine_fail)
            if (!fail_code) {
              fail_code = e.code;
              failure = e.what();
            }
          } catch (const std::exception &e) {
#pragma omp critical(impgx_refine_fail)
            if (!fail_code) {
              fail_code = IMPGX_E_INVALID;
              failure = e.what();
            }
          }
        }
      }
      if (fail_code) throw Error(fail_code, failure);
    }
    for (auto &l : L)
      if (!l.done && l.capped && l.have_best && l.best.support >= l.max_entities) l.done = true;  // check_max
  }
  for (size_t i = 0; i < n; i++) {
    const Locus &l = L[i];
    REQUIRE(l.have_best, IMPGX_E_INVALID, "locus " + std::to_string(i) + ": no valid flank sizes evaluated");
    out->refined_start.push_back(l.best.start); out->refined_end.push_back(l.best.end);
    out->original_start.push_back(l.s); out->original_end.push_back(l.e);
    out->left.push_back(l.best.left); out->right.push_back(l.best.right);
    out->support.push_back(l.best.support); out->original_support.push_back(l.original_support);
    out->ent_seq.insert(out->ent_seq.end(), l.best.e_seq.begin(), l.best.e_seq.end());
    out->ent_start.insert(out->ent_start.end(), l.best.e_start.begin(), l.best.e_start.end());
    out->ent_end.insert(out->ent_end.end(), l.best.e_end.begin(), l.best.e_end.end());
    out->ent_off.push_back(out->ent_seq.size());
  }
  return out.release();
}

}  // namespace impgx

#define RF_BEGIN try {
#define RF_END                                       \
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

int impgx_refine(impgx_index *idx, const impgx_range *loci, size_t n, const impgx_refine_params *params,
                 impgx_refine_results **out) {
  RF_BEGIN
  REQUIRE(out && (loci || n == 0), IMPGX_E_INVALID, "NULL argument");
  impgx::check_refine_params(idx, params);
  *out = impgx::refine(idx, loci, n, *params);
  RF_END
}

int impgx_refine_view_get(const impgx_refine_results *r, impgx_refine_view *v) {
  if (!r || !v) return IMPGX_E_INVALID;
  v->n = r->refined_start.size();
  v->refined_start = r->refined_start.data(); v->refined_end = r->refined_end.data();
  v->original_start = r->original_start.data(); v->original_end = r->original_end.data();
  v->applied_left_extension = r->left.data(); v->applied_right_extension = r->right.data();
  v->support_count = r->support.data(); v->original_support_count = r->original_support.data();
  v->entity_offsets = r->ent_off.data();
  v->entity_seq = r->ent_seq.data(); v->entity_start = r->ent_start.data(); v->entity_end = r->ent_end.data();
  v->candidates_evaluated = r->candidates;
  v->batches = r->batches;
  return IMPGX_OK;
}
void impgx_refine_results_free(impgx_refine_results *r) { delete r; }

int impgx_populate_cigar_cache(impgx_index *idx, uint32_t target_id, int32_t start, int32_t end, uint64_t *n_keys) {
  RF_BEGIN
  REQUIRE(idx && n_keys, IMPGX_E_INVALID, "NULL argument");
  REQUIRE(target_id < idx->n_seqs, IMPGX_E_INVALID, "unknown target id");
  // a tree holds at most one entry per alignment (the reversed entry of a record lives in the OTHER sequence's
  // tree, self alignments have none), so the cache keys of a stab are its visited entries
  *n_keys = impgx::stab_count_closed(idx, target_id, start, end);
  RF_END
}

int impgx_query_with_cache_batch(impgx_index *idx, uint32_t target_id, int32_t orig_start, int32_t orig_end,
                                 const int32_t *left, const int32_t *right, size_t n, const impgx_params *params,
                                 impgx_results **out) {
  RF_BEGIN
  REQUIRE(idx && params && out && ((left && right) || n == 0), IMPGX_E_INVALID, "NULL argument");
  REQUIRE(target_id < idx->n_seqs, IMPGX_E_INVALID, "unknown target id");
  REQUIRE(params->mode == IMPGX_MODE_QUERY, IMPGX_E_INVALID, "query_with_cache is Impg::query: mode must be IMPGX_MODE_QUERY");
  const int32_t seq_len = (int32_t)idx->seq_lens[target_id];
  std::vector<impgx_range> rows;
  std::vector<size_t> src;
  for (size_t k = 0; k < n; k++) {
    REQUIRE(left[k] >= 0 && right[k] >= 0, IMPGX_E_INVALID, "negative flank");
    const int32_t s = (int32_t)std::max<int64_t>((int64_t)orig_start - left[k], 0);
    const int32_t e = (int32_t)std::min<int64_t>((int64_t)orig_end + right[k], seq_len);
    if (e <= s) continue;
    rows.push_back(impgx_range{target_id, s, e});
    src.push_back(k);
  }
  std::unique_ptr<impgx_results> res(impgx::query_batch(idx, rows.data(), rows.size(), *params, /*bed=*/false,
                                                        /*ranges_on_device=*/false, /*results_to_host=*/true, nullptr));
  if (rows.size() != n) {  // empty candidates: rows without results
    std::vector<uint64_t> ro(n + 1, 0);
    for (size_t j = 0; j < rows.size(); j++) ro[src[j] + 1] = res->row_off[j + 1] - res->row_off[j];
    for (size_t k = 0; k < n; k++) ro[k + 1] += ro[k];
    res->row_off.resize(n + 1);
    for (size_t k = 0; k <= n; k++) res->row_off[k] = ro[k];
    res->n_rows = n;
  }
  *out = res.release();
  RF_END
}

}  // extern "C"
