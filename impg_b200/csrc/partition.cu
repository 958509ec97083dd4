// partition.cu — `impg partition -o bed` over the HBM index.
//
// partition_alignments (src/commands/partition.rs:158-712) is a loop over windows in which
// every window depends on the regions the previous windows claimed, so the windows cannot be
// batched: the device runs ONE masked transitive query per window (stab, liftover, fold with
// the mask as the initial visited set, and — for -d >= 0 — the on-chip BED merge, which on the
// query axis computes exactly merge_overlaps' hulls, see Partitioner::feed) and the
// bookkeeping between windows (what is masked, what is still missing, which window is next)
// is this host code. Only (sequence, min, max) of an interval reaches partitions.bed
// (:1509-1542, :1682-1717), and every step between the query and the writer treats the two
// orientations symmetrically, so intervals are kept as start <= end.
#include <algorithm>
#include <cmath>
#include <map>
#include <set>
#include <tuple>
#include <string>
#include <vector>

#include "engine.cuh"

namespace impgx {

struct Iv {
  uint32_t seq;
  int32_t s, e;
};

// SortedRanges with min_distance = 0 (src/impg.rs:242-369), the form partition uses for its
// masked and missing maps (:250-267)
struct Ranges {
  std::vector<int2> r;  // x = start, y = end; sorted by start, disjoint
  int32_t seq_len = 0;

  size_t lower(int32_t key) const {
    size_t lo = 0, hi = r.size();
    while (lo < hi) {
      size_t mid = (lo + hi) / 2;
      if (r[mid].x < key) lo = mid + 1;
      else hi = mid;
    }
    return lo;
  }
  // first range that can reach `key`: the one starting at key, else the previous one if it
  // ends after key (the "previous range might overlap" idiom, :1010-1028, :1147-1162, :1269-1284)
  size_t reach(int32_t key) const {
    size_t pos = lower(key);
    if (pos < r.size() && r[pos].x == key) return pos;
    return (pos > 0 && r[pos - 1].y > key) ? pos - 1 : pos;
  }
  void fuse_from(size_t w) {  // merge_forward_from :352-368
    size_t rd = w + 1;
    for (; rd < r.size(); rd++) {
      if (r[w].y >= r[rd].x) r[w].y = std::max(r[w].y, r[rd].y);
      else r[++w] = r[rd];
    }
    r.resize(w + 1);
  }
  void insert(int32_t a, int32_t b) {  // :273-349 with min_distance 0; the returned pieces are unused here
    int32_t start = std::min(a, b), end = std::max(a, b);
    if (start < 0) start = 0;            // `start < min_distance`
    if (end > seq_len) end = seq_len;    // `end > sequence_length - min_distance`
    size_t pos = lower(start);
    if (pos > 0 && r[pos - 1].y >= start) {
      r[pos - 1].y = std::max(r[pos - 1].y, end);
      fuse_from(pos - 1);
    } else if (pos < r.size() && end >= r[pos].x) {
      r[pos].x = std::min(start, r[pos].x);
      r[pos].y = std::max(end, r[pos].y);
      fuse_from(pos);
    } else {
      r.insert(r.begin() + pos, make_int2(start, end));
    }
  }
  int64_t total() const {
    int64_t t = 0;
    for (auto &x : r) t += (int64_t)(x.y - x.x);
    return t;
  }
};

}  // namespace impgx

struct impgx_partitions {
  std::vector<uint32_t> pnum, seq;
  std::vector<int32_t> start, end;
  size_t n_partitions = 0;
  uint64_t n_windows = 0, partitioned_bp = 0, total_bp = 0;
};

struct impgx_partitioner {
  // parameters
  int64_t window_size = 0;
  int32_t merge_distance = 0, min_missing = 0, min_boundary = 0;
  bool rehome = false;
  enum Select { LONGEST, TOTAL, GROUP } select = LONGEST;
  int group_fields = 1;
  std::string sep = "#";
  // state
  uint32_t n_seqs = 0;
  std::vector<std::string> names;
  std::vector<impgx::Ranges> masked, missing;  // missing[s].r empty = sequence done (removed from the map, :1313-1316)
  std::vector<impgx_range> queue;              // windows of the current round
  size_t head = 0;
  bool awaiting_feed = false;
  std::vector<uint64_t> mask_off;
  std::vector<int32_t> mask_rng;
  std::vector<uint32_t> touched;  // sequences whose mask changed since the CSR above was brought up to date
  int32_t no_ranges[2] = {0, 0};
  const int32_t *mask_ranges_ptr() const { return mask_rng.empty() ? no_ranges : mask_rng.data(); }  // never NULL
  // what select_round asks for, kept up to date as regions go missing -> masked (only the active mode's index)
  std::set<std::tuple<int32_t, uint32_t, int32_t, int32_t>> by_len;  // LONGEST: (length, sequence, start, end)
  std::set<std::pair<int64_t, uint32_t>> by_total;                   // TOTAL: (missing bases, sequence), non-empty only
  std::vector<uint32_t> group_of;                                    // GROUP: sequence -> group (ascending prefix)
  std::vector<int64_t> group_missing;
  std::vector<uint32_t> group_live;                                  // sequences of the group that still miss something
  std::vector<std::vector<uint32_t>> group_members;                  // longest first, then by id
  void index_missing(uint32_t seq, bool add);
  // collected partitions: intervals of partition k are ivs[part_off[k] .. part_off[k+1])
  std::vector<impgx::Iv> ivs;
  std::vector<size_t> part_off{0};
  uint64_t n_windows = 0, partitioned_bp = 0, total_bp = 0;

  void add_windows(uint32_t seq, int32_t a, int32_t b, bool tail_rule_same_seq_only);
  void select_round();
  void merge_within(std::vector<impgx::Iv> &v, int32_t d) const;
  void cut_by_mask(uint32_t seq, const impgx::Iv *first, const impgx::Iv *last, std::vector<impgx::Iv> &out);
  void feed(size_t n, const uint32_t *q_id, const int32_t *q_first, const int32_t *q_last);
  impgx_partitions *finish() const;
};

using impgx::Iv;
using impgx::Ranges;

// windows of one range (:911-932): a tail shorter than the window joins the previous window
// OF THIS RANGE. For the starting-sequences file (:229-245) the rule looks at the last window
// of the whole list instead, provided it is on the same sequence.
void impgx_partitioner::add_windows(uint32_t seq, int32_t a, int32_t b, bool starting_file_rule) {
  const size_t base = queue.size();
  for (int32_t pos = a; pos < b;) {
    const int32_t we = (int32_t)std::min<int64_t>((int64_t)pos + window_size, b);
    const bool is_short = (int64_t)(we - pos) < window_size;
    const bool have_prev = starting_file_rule ? (!queue.empty() && queue.back().target_id == seq) : queue.size() > base;
    if (is_short && have_prev) queue.back().end = b;
    else queue.push_back({seq, pos, we});
    pos = we;
  }
}

// The missing regions of `seq` enter (add) or leave the selection index of the active mode; called around
// every change of missing[seq], so a round costs O(log) instead of a scan over all sequences.
void impgx_partitioner::index_missing(uint32_t seq, bool add) {
  const Ranges &ms = missing[seq];
  if (select == LONGEST) {
    for (auto &x : ms.r) {
      const auto key = std::make_tuple(x.y - x.x, seq, x.x, x.y);
      if (add) by_len.insert(key);
      else by_len.erase(key);
    }
  } else if (ms.r.empty()) {
    return;  // a sequence without missing regions is not in the map (:1313-1316)
  } else if (select == TOTAL) {
    const auto key = std::make_pair(ms.total(), seq);
    if (add) by_total.insert(key);
    else by_total.erase(key);
  } else {
    const uint32_t g = group_of[seq];
    group_missing[g] += add ? ms.total() : -ms.total();
    group_live[g] += add ? 1u : (uint32_t)-1;
  }
}

// select_and_window_sequences (:715-937)
void impgx_partitioner::select_round() {
  queue.clear();
  head = 0;
  if (select == LONGEST) {
    // max_by (length, id); of equal keys (two regions of one sequence) the later one wins
    if (!by_len.empty()) {
      const auto &k = *by_len.rbegin();
      add_windows(std::get<1>(k), std::get<2>(k), std::get<3>(k), false);
    }
  } else if (select == TOTAL) {
    // max_by (total missing, id)
    if (!by_total.empty()) {
      const uint32_t s = by_total.rbegin()->second;
      add_windows(s, 0, missing[s].seq_len, false);
    }
  } else {
    // sample / haplotype (:807-895): the group (name prefix) with the most missing bases, ties to the greater
    // prefix; its sequences that still miss something, longest first (equal lengths: the reference's order is
    // hash order, here by id)
    long best = -1;
    for (size_t g = 0; g < group_missing.size(); g++)  // groups are numbered by ascending prefix
      if (group_live[g] && (best < 0 || group_missing[g] >= group_missing[(size_t)best])) best = (long)g;
    if (best >= 0)
      for (uint32_t s : group_members[(size_t)best])
        if (!missing[s].r.empty()) add_windows(s, 0, missing[s].seq_len, false);
  }
}

// merge_overlaps (:939-976) on normalised intervals: hulls of the runs whose starts stay within
// `d` of the running end, per sequence, in (sequence, start) order
void impgx_partitioner::merge_within(std::vector<Iv> &v, int32_t d) const {
  if (v.size() <= 1 || d < 0) return;
  std::stable_sort(v.begin(), v.end(), [](const Iv &a, const Iv &b) { return a.seq != b.seq ? a.seq < b.seq : a.s < b.s; });
  size_t w = 0;
  for (size_t rd = 1; rd < v.size(); rd++) {
    // the reference computes curr_max + merge_distance in i32 (debug builds would panic on overflow)
    if (v[w].seq != v[rd].seq || (int64_t)v[rd].s > (int64_t)v[w].e + d) v[++w] = v[rd];
    else {
      v[w].s = std::min(v[w].s, v[rd].s);
      v[w].e = std::max(v[w].e, v[rd].e);
    }
  }
  v.resize(w + 1);
}

// process_sequence_overlaps (:984-1319) for one run of intervals on `seq`
void impgx_partitioner::cut_by_mask(uint32_t seq, const Iv *first, const Iv *last, std::vector<Iv> &out) {
  Ranges &mk = masked[seq];
  Ranges &ms = missing[seq];
  // 1. slivers of missing sequence (< min_missing) left next to an interval end are absorbed
  std::vector<int2> ext;
  for (const Iv *q = first; q != last; q++)
    for (size_t i = ms.reach(q->s); i < ms.r.size() && ms.r[i].x <= q->e; i++) {
      const int2 m = ms.r[i];
      if (q->s > m.x && q->s < m.y && q->s - m.x < min_missing) ext.push_back(make_int2(m.x, q->s));
      if (q->e > m.x && q->e < m.y && m.y - q->e < min_missing) ext.push_back(make_int2(q->e, m.y));
    }
  if (!ext.empty()) {
    std::stable_sort(ext.begin(), ext.end(), [](const int2 &a, const int2 &b) { return a.x < b.x; });
    size_t w = 0;
    for (size_t rd = 1; rd < ext.size(); rd++) {
      if (ext[rd].x <= ext[w].y) ext[w].y = std::max(ext[w].y, ext[rd].y);
      else ext[++w] = ext[rd];
    }
    ext.resize(w + 1);
  }
  // 2. every interval: grow by the extensions touching an end (in order, the grown interval is
  //    what the next extension is tested against), then keep what the mask does not cover yet
  std::vector<int2> claimed;
  for (const Iv *q = first; q != last; q++) {
    int32_t s = q->s, e = q->e;
    for (auto &x : ext)
      if ((x.y >= s && x.x <= s) || (x.x <= e && x.y >= e)) {
        s = std::min(s, x.x);
        e = std::max(e, x.y);
      }
    claimed.push_back(make_int2(s, e));
    int32_t cur = s;
    for (size_t i = mk.reach(cur); i < mk.r.size(); i++) {
      const int2 m = mk.r[i];
      if (m.x > e) break;
      if (m.y <= cur) continue;
      if (cur < m.x) out.push_back({seq, cur, m.x});
      cur = std::max(cur, m.y);
      if (cur >= e) break;
    }
    if (cur < e) out.push_back({seq, cur, e});
  }
  // 3. the claimed intervals join the mask; missing = missing minus mask
  for (auto &c : claimed) mk.insert(c.x, c.y);
  touched.push_back(seq);
  index_missing(seq, false);
  if (!ms.r.empty()) {
    std::vector<int2> old;
    old.swap(ms.r);
    for (auto &m : old) {
      int32_t cur = m.x;
      for (size_t i = mk.reach(m.x); i < mk.r.size() && cur < m.y; i++) {
        const int2 k = mk.r[i];
        if (k.x > m.y) break;
        if (k.y <= cur) continue;
        if (cur < k.x) ms.insert(cur, k.x);
        cur = std::max(cur, k.y);
      }
      if (cur < m.y) ms.insert(cur, m.y);
    }
  }
  index_missing(seq, true);
}

void impgx_partitioner::feed(size_t n, const uint32_t *q_id, const int32_t *q_first, const int32_t *q_last) {
  std::vector<Iv> ov(n);
  for (size_t i = 0; i < n; i++) {
    REQUIRE(q_id[i] < n_seqs, IMPGX_E_INVALID, "overlap on an unknown sequence");
    ov[i] = {q_id[i], std::min(q_first[i], q_last[i]), std::max(q_first[i], q_last[i])};
  }
  // merge_overlaps(merge_distance). Fed with the device's BED rows (impgx_partition, -d >= 0) this
  // pass finds nothing left to merge: merge_adjusted_intervals_gap_2d only unions boxes whose query
  // gap is <= d and merge_query_adjusted_intervals with merge_strands sweeps the query axis with
  // the same `next_start <= curr_end + d` rule, so its rows are these hulls already.
  merge_within(ov, merge_distance);
  if (min_boundary > 0)  // extend_to_close_boundaries (:1369-1408)
    for (auto &q : ov) {
      const int32_t len = missing[q.seq].seq_len;
      if (q.s < min_boundary) q.s = 0;
      if (len - q.e < min_boundary) q.e = len;
    }
  // mask_and_update_regions (:978-1366): contiguous runs of one sequence
  std::vector<Iv> kept;
  for (size_t a = 0; a < ov.size();) {
    size_t b = a + 1;
    while (b < ov.size() && ov[b].seq == ov[a].seq) b++;
    cut_by_mask(ov[a].seq, ov.data() + a, ov.data() + b, kept);
    a = b;
  }
  if (!kept.empty()) {
    merge_within(kept, 0);
    for (auto &q : kept) partitioned_bp += (uint64_t)(q.e - q.s);
    ivs.insert(ivs.end(), kept.begin(), kept.end());
    part_off.push_back(ivs.size());
  }
}

// rehome_singleton_slivers (:45-156) + the flattening of write_single_partition_file
impgx_partitions *impgx_partitioner::finish() const {
  const size_t P = part_off.size() - 1;
  std::vector<size_t> owner(ivs.size());
  std::vector<size_t> order(ivs.size());
  for (size_t k = 0; k < P; k++)
    for (size_t i = part_off[k]; i < part_off[k + 1]; i++) owner[i] = k;
  for (size_t i = 0; i < order.size(); i++) order[i] = i;
  std::vector<size_t> count(P, 0);
  for (size_t k = 0; k < P; k++) count[k] = part_off[k + 1] - part_off[k];
  bool resorted = false;
  if (rehome && P > 0 && std::count(count.begin(), count.end(), (size_t)1) > 0) {
    resorted = true;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) {
      const Iv &x = ivs[a], &y = ivs[b];
      if (x.seq != y.seq) return x.seq < y.seq;
      if (x.s != y.s) return x.s < y.s;
      return x.e < y.e;
    });
    std::vector<size_t> own(order.size());  // partition of the k-th interval in sorted order
    for (size_t k = 0; k < order.size(); k++) own[k] = owner[order[k]];
    for (int pass = 1;; pass++) {
      std::vector<char> single(P);
      for (size_t k = 0; k < P; k++) single[k] = count[k] == 1;
      std::vector<std::pair<size_t, size_t>> moves;
      for (size_t k = 0; k < order.size(); k++) {
        if (!single[own[k]]) continue;
        const Iv &me = ivs[order[k]];
        // flanks: the neighbours in sorted order that abut this interval on the same sequence
        long lp = -1, rp = -1;
        if (k > 0) {
          const Iv &l = ivs[order[k - 1]];
          if (l.seq == me.seq && l.e == me.s) lp = (long)own[k - 1];
        }
        if (k + 1 < order.size()) {
          const Iv &r = ivs[order[k + 1]];
          if (r.seq == me.seq && r.s == me.e) rp = (long)own[k + 1];
        }
        const bool lok = lp >= 0 && !single[lp], rok = rp >= 0 && !single[rp];
        long to;
        if (lok && rok) to = count[lp] >= count[rp] ? lp : rp;
        else if (lok) to = lp;
        else if (rok) to = rp;
        else continue;
        if ((size_t)to != own[k]) moves.push_back({k, (size_t)to});
      }
      if (moves.empty() || pass > 100) break;
      for (auto &m : moves) {
        count[own[m.first]]--;
        count[m.second]++;
        own[m.first] = m.second;
      }
    }
    for (size_t k = 0; k < order.size(); k++) owner[order[k]] = own[k];
  }
  impgx_partitions *out = new impgx_partitions();
  out->n_partitions = P;
  out->n_windows = n_windows;
  out->partitioned_bp = partitioned_bp;
  out->total_bp = total_bp;
  // partitions in creation order; inside a partition the intervals keep their order, or — once the
  // rehoming pass ran — the (sequence, start, end) order it rebuilds every partition in (:137-142)
  std::vector<std::vector<size_t>> members(P);
  if (resorted) for (size_t k = 0; k < order.size(); k++) members[owner[order[k]]].push_back(order[k]);
  else for (size_t i = 0; i < ivs.size(); i++) members[owner[i]].push_back(i);
  for (size_t k = 0; k < P; k++)
    for (size_t i : members[k]) {
      out->pnum.push_back((uint32_t)k);
      out->seq.push_back(ivs[i].seq);
      out->start.push_back(ivs[i].s);
      out->end.push_back(ivs[i].e);
    }
  return out;
}

static impgx_partitioner *make_partitioner(const uint64_t *seq_lens, const char *const *names, uint32_t n_seqs,
                                           const impgx_partition_params &pp) {
  REQUIRE(seq_lens || n_seqs == 0, IMPGX_E_INVALID, "seq_lens is NULL");
  REQUIRE(pp.window_size > 0 && pp.window_size <= 0x7fffffffull, IMPGX_E_INVALID, "window_size must be in 1..2^31-1");
  std::unique_ptr<impgx_partitioner> p(new impgx_partitioner());
  p->window_size = (int64_t)pp.window_size;
  p->merge_distance = pp.merge_distance;
  p->min_missing = pp.min_missing_size;
  p->min_boundary = pp.min_boundary_distance;
  p->rehome = pp.rehome_singletons != 0;
  const std::string mode = pp.selection_mode ? pp.selection_mode : "longest";
  auto starts = [&](const char *pre) { return mode.rfind(pre, 0) == 0; };
  if (mode == "longest") p->select = impgx_partitioner::LONGEST;
  else if (mode == "total") p->select = impgx_partitioner::TOTAL;
  else if (mode == "sample" || mode == "haplotype" || starts("sample,") || starts("haplotype,")) {
    p->select = impgx_partitioner::GROUP;
    p->group_fields = starts("haplotype") ? 2 : 1;
    const size_t comma = mode.find(',');
    if (comma != std::string::npos) p->sep = mode.substr(comma + 1);
    REQUIRE(names, IMPGX_E_INVALID, "the sample / haplotype selection modes need the sequence names");
  } else {
    throw impgx::Error(IMPGX_E_INVALID,
                       "Invalid selection mode. Must be 'longest', 'total', 'sample[,sep]', or 'haplotype[,sep]'.");
  }
  p->n_seqs = n_seqs;
  p->masked.resize(n_seqs);
  p->missing.resize(n_seqs);
  if (names && p->select == impgx_partitioner::GROUP) {
    p->names.resize(n_seqs);
    for (uint32_t s = 0; s < n_seqs; s++) p->names[s] = names[s] ? names[s] : "";
  }
  for (uint32_t s = 0; s < n_seqs; s++) {
    REQUIRE(seq_lens[s] <= 0x7fffffffull, IMPGX_E_INVALID, "sequence longer than 2^31-1");
    const int32_t len = (int32_t)seq_lens[s];
    p->masked[s].seq_len = p->missing[s].seq_len = len;
    p->missing[s].insert(0, len);  // :259-267 (an empty sequence leaves the degenerate range (0, 0) behind)
    p->total_bp += seq_lens[s];
  }
  if (p->select == impgx_partitioner::GROUP) {
    // sequences grouped by name prefix (:807-839): the first field, or the first two joined by the separator
    std::map<std::string, std::vector<uint32_t>> by_prefix;
    for (uint32_t s = 0; s < n_seqs; s++) {
      const std::string &nm = p->names[s];
      const std::string &sep = p->sep;
      std::string key;
      if (sep.empty()) {
        // str::split("") yields "", c1, c2, …, "": the first field is empty, the second the first char
        key = p->group_fields == 1 ? std::string() : nm.substr(0, nm.empty() ? 0 : 1);
      } else {
        const size_t f1 = nm.find(sep);
        if (p->group_fields == 1) key = nm.substr(0, f1);
        else if (f1 == std::string::npos) key = nm + sep;
        else key = nm.substr(0, nm.find(sep, f1 + sep.size()));
      }
      by_prefix[key].push_back(s);
    }
    p->group_of.assign(n_seqs, 0);
    for (auto &kv : by_prefix) {  // ascending prefix = ascending group number
      const uint32_t g = (uint32_t)p->group_members.size();
      std::vector<uint32_t> members = kv.second;
      std::stable_sort(members.begin(), members.end(),
                       [&](uint32_t a, uint32_t b) { return p->missing[a].seq_len > p->missing[b].seq_len; });
      for (uint32_t m : members) p->group_of[m] = g;
      p->group_members.push_back(std::move(members));
    }
    p->group_missing.assign(p->group_members.size(), 0);
    p->group_live.assign(p->group_members.size(), 0);
  }
  for (uint32_t s = 0; s < n_seqs; s++) p->index_missing(s, true);
  for (size_t k = 0; k < pp.n_starting_seqs; k++) {
    REQUIRE(pp.starting_seqs && pp.starting_seqs[k] < n_seqs, IMPGX_E_INVALID, "starting sequence id out of range");
    p->add_windows(pp.starting_seqs[k], 0, (int32_t)seq_lens[pp.starting_seqs[k]], true);
  }
  if (p->queue.empty()) p->select_round();
  p->mask_off.assign((size_t)n_seqs + 1, 0);
  return p.release();
}

static bool next_window(impgx_partitioner *p, impgx_range *w) {
  REQUIRE(!p->awaiting_feed, IMPGX_E_INVALID, "the previous window was not fed back");
  if (p->head == p->queue.size()) {
    // the reference's loop ends when a round begins with no window (:295); the very first
    // round may come from the starting-sequences file
    if (p->queue.empty()) return false;
    p->select_round();
    if (p->queue.empty()) return false;
  }
  *w = p->queue[p->head++];
  p->n_windows++;
  p->awaiting_feed = true;
  // masked_regions as CSR over all sequences, spliced: the slices of the sequences the last feed touched are
  // rebuilt from their range lists, everything else is copied in bulk from the previous arrays (no walk over
  // every sequence's list; the offsets are one sequential pass)
  if (!p->touched.empty()) {
    std::sort(p->touched.begin(), p->touched.end());
    p->touched.erase(std::unique(p->touched.begin(), p->touched.end()), p->touched.end());
    const std::vector<uint64_t> &old_off = p->mask_off;
    const std::vector<int32_t> &old_rng = p->mask_rng;
    std::vector<uint64_t> off((size_t)p->n_seqs + 1);
    std::vector<int32_t> rng;
    rng.reserve(old_rng.size() + 4 * p->touched.size());
    uint32_t from = 0;  // first sequence not yet emitted
    int64_t delta = 0;  // new offset - old offset from `from` on
    for (uint32_t t : p->touched) {
      for (uint32_t s = from; s <= t; s++) off[s] = (uint64_t)((int64_t)old_off[s] + delta);
      rng.insert(rng.end(), old_rng.begin() + 2 * old_off[from], old_rng.begin() + 2 * old_off[t]);
      for (auto &x : p->masked[t].r) {
        rng.push_back(x.x);
        rng.push_back(x.y);
      }
      delta += (int64_t)p->masked[t].r.size() - (int64_t)(old_off[t + 1] - old_off[t]);
      from = t + 1;
    }
    for (uint32_t s = from; s <= p->n_seqs; s++) off[s] = (uint64_t)((int64_t)old_off[s] + delta);
    rng.insert(rng.end(), old_rng.begin() + 2 * old_off[from], old_rng.begin() + 2 * old_off[p->n_seqs]);
    p->mask_off.swap(off);
    p->mask_rng.swap(rng);
    p->touched.clear();
  }
  return true;
}

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
  }

extern "C" {

int impgx_partitioner_new(const uint64_t *seq_lens, const char *const *names, uint32_t n_seqs,
                          const impgx_partition_params *params, impgx_partitioner **out) {
  API_BEGIN
  REQUIRE(params && out, IMPGX_E_INVALID, "NULL argument");
  *out = nullptr;
  *out = make_partitioner(seq_lens, names, n_seqs, *params);
  API_END
  return IMPGX_OK;
}

int impgx_partitioner_next(impgx_partitioner *p, impgx_range *window, const uint64_t **mask_offsets,
                           const int32_t **mask_ranges) {
  API_BEGIN
  REQUIRE(p && window, IMPGX_E_INVALID, "NULL argument");
  if (!next_window(p, window)) return 0;
  if (mask_offsets) *mask_offsets = p->mask_off.data();
  if (mask_ranges) *mask_ranges = p->mask_ranges_ptr();
  return 1;
  API_END
  return IMPGX_E_INVALID;
}

int impgx_partitioner_feed(impgx_partitioner *p, size_t n, const uint32_t *q_id, const int32_t *q_first,
                           const int32_t *q_last) {
  API_BEGIN
  REQUIRE(p && (n == 0 || (q_id && q_first && q_last)), IMPGX_E_INVALID, "NULL argument");
  REQUIRE(p->awaiting_feed, IMPGX_E_INVALID, "no window is outstanding");
  p->awaiting_feed = false;
  p->feed(n, q_id, q_first, q_last);
  API_END
  return IMPGX_OK;
}

int impgx_partitioner_finish(impgx_partitioner *p, impgx_partitions **out) {
  API_BEGIN
  REQUIRE(p && out, IMPGX_E_INVALID, "NULL argument");
  *out = p->finish();
  API_END
  return IMPGX_OK;
}

void impgx_partitioner_free(impgx_partitioner *p) { delete p; }

int impgx_partition(impgx_index *idx, const impgx_partition_params *params, impgx_partitions **out) {
  API_BEGIN
  REQUIRE(idx && params && out, IMPGX_E_INVALID, "NULL argument");
  *out = nullptr;
  impgx::check_device(idx->device);  // no CPU fallback: fails with IMPGX_E_NO_DEVICE before any window is taken
  std::vector<const char *> nm;
  if (idx->names.size() == idx->n_seqs)
    for (auto &s : idx->names) nm.push_back(s.c_str());
  std::unique_ptr<impgx_partitioner> p(
      make_partitioner(idx->seq_lens.data(), nm.empty() ? nullptr : nm.data(), idx->n_seqs, *params));
  impgx_params q{};
  q.mode = params->multi_impg ? (params->transitive_dfs ? IMPGX_MODE_MULTI_DFS : IMPGX_MODE_MULTI_BFS)
                              : (params->transitive_dfs ? IMPGX_MODE_DFS : IMPGX_MODE_BFS);
  q.max_depth = params->max_depth;
  q.min_transitive_len = params->min_transitive_len;
  q.min_distance_between_ranges = params->min_distance_between_ranges;
  q.min_output_length = -1;  // None for partition (:368, :384)
  q.store_cigar = 0;         // :369, :385
  q.min_identity = params->min_identity;
  q.subset_mask = nullptr;   // :373, :389
  q.merge_distance = params->merge_distance;
  q.merge_strands = 1;
  // -d >= 0: the device merges (BED path) and returns the hulls; --no-merge: the raw result list
  const bool device_merge = params->merge_distance >= 0;
  impgx_range w;
  while (next_window(p.get(), &w)) {
    q.mask_offsets = p->mask_off.data();
    q.mask_ranges = p->mask_ranges_ptr();
    std::unique_ptr<impgx_results> res(
        impgx::query_batch(idx, &w, 1, q, device_merge, /*ranges_on_device=*/false, /*results_to_host=*/true, nullptr));
    p->awaiting_feed = false;
    p->feed(res->n_results, res->qid.data(), res->qf.data(), res->ql.data());
  }
  *out = p->finish();
  API_END
  return IMPGX_OK;
}

int impgx_partitions_view(const impgx_partitions *parts, impgx_partition_view *v) {
  if (!parts || !v) return IMPGX_E_INVALID;
  v->n_intervals = parts->pnum.size();
  v->n_partitions = parts->n_partitions;
  v->n_windows = parts->n_windows;
  v->partitioned_bp = parts->partitioned_bp;
  v->total_bp = parts->total_bp;
  v->partition_num = parts->pnum.data();
  v->seq_id = parts->seq.data();
  v->start = parts->start.data();
  v->end = parts->end.data();
  return IMPGX_OK;
}

char *impgx_partitions_format_bed(const impgx_index *idx, const impgx_partitions *parts, int64_t partition) {
  if (!idx || !parts) return nullptr;
  std::string s;
  for (size_t i = 0; i < parts->pnum.size(); i++) {
    if (partition >= 0 && (int64_t)parts->pnum[i] != partition) continue;
    const uint32_t q = parts->seq[i];
    s += (q < idx->names.size() && !idx->names[q].empty()) ? idx->names[q] : ("seq" + std::to_string(q));
    s += '\t';
    s += std::to_string(parts->start[i]);
    s += '\t';
    s += std::to_string(parts->end[i]);
    if (partition < 0) {
      s += '\t';
      s += std::to_string(parts->pnum[i]);
    }
    s += '\n';
  }
  char *o = (char *)malloc(s.size() + 1);
  if (o) memcpy(o, s.c_str(), s.size() + 1);
  return o;
}

void impgx_partitions_free(impgx_partitions *parts) { delete parts; }

}  // extern "C"
