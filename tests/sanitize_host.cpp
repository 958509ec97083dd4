// TEST INFRASTRUCTURE: AddressSanitizer / UBSan run of the host-only parts of libimpgx (the partition
// bookkeeping of csrc/partition.cu and the .impg reader / writer of csrc/impg_file.cu), compiled as plain C++
// and driven with random inputs: random partition runs must terminate and tile every sequence, damaged index
// files must be rejected without touching memory out of bounds. Built and run by tests/test_host_sanitizers.py.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "impgx.h"
// test hooks of libimpgx (csrc/api.cu), not part of the public header
extern "C" long impgx_debug_host_columns_shard(const impgx_record *records, size_t n, const uint64_t *run_offsets, uint32_t n_seqs,
                                               int bidirectional, int32_t *e_start, int32_t *e_end, int32_t *e_pmax,
                                               uint32_t *e_vrank, uint32_t *e_query_id, uint32_t *e_flags, uint32_t *e_aln,
                                               uint64_t *tgt_off, const uint32_t *owner, uint32_t rank);
extern "C" void impgx_debug_visit_ranks(size_t n, uint32_t *rank);
extern "C" char *impgx_debug_format_rows(const char *const *names, const uint64_t *lens, uint32_t n_seqs, size_t n,
                                         const uint32_t *q_id, const int32_t *q_first, const int32_t *q_last, const uint32_t *t_id,
                                         const int32_t *t_first, const int32_t *t_last, const uint64_t *cig_off, const uint32_t *cig,
                                         const char *name, int32_t d, int format, int original_coordinates);
int main(int argc, char **argv) {
  std::mt19937_64 rng(7);
  const char *modes[] = {"longest", "total", "sample", "haplotype,#", "sample,"};
  for (int trial = 0; trial < 300; trial++) {
    uint32_t n = 1 + rng() % 40;
    std::vector<uint64_t> lens(n);
    std::vector<std::string> nm(n);
    std::vector<const char *> names(n);
    for (uint32_t i = 0; i < n; i++) {
      lens[i] = (rng() % 7 == 0) ? 0 : 50 + rng() % 5000;
      nm[i] = "g" + std::to_string(rng() % 5) + "#" + std::to_string(rng() % 2) + "#c" + std::to_string(i);
      names[i] = nm[i].c_str();
    }
    impgx_partition_params pp;
    memset(&pp, 0, sizeof pp);
    pp.window_size = 20 + rng() % 3000;
    pp.selection_mode = modes[rng() % 5];
    pp.merge_distance = (int32_t)(rng() % 4 == 0 ? -1 : rng() % 500);
    pp.min_missing_size = (int32_t)(rng() % 400);
    pp.min_boundary_distance = (int32_t)(rng() % 400);
    pp.rehome_singletons = rng() % 2;
    pp.max_depth = 2;
    std::vector<uint32_t> st;
    if (rng() % 3 == 0) for (int k = 0; k < 3; k++) st.push_back((uint32_t)(rng() % n));
    pp.starting_seqs = st.data();
    pp.n_starting_seqs = st.size();
    impgx_partitioner *p = nullptr;
    if (impgx_partitioner_new(lens.data(), names.data(), n, &pp, &p) != 0) { printf("new failed: %s\n", impgx_last_error()); return 1; }
    impgx_range w;
    const uint64_t *mo; const int32_t *mr;
    int guard = 0;
    while (impgx_partitioner_next(p, &w, &mo, &mr) == 1 && guard++ < 100000) {
      // touch the CSR the way the engine would
      uint64_t tot = mo[n];
      long long chk = 0;
      for (uint64_t k = 0; k < 2 * tot; k++) chk += mr[k];
      (void)chk;
      // random "query result": the window itself (cut to unmasked is the caller's business) plus random intervals
      std::vector<uint32_t> q; std::vector<int32_t> a, b;
      q.push_back(w.target_id); a.push_back(w.start); b.push_back(w.end);
      int extra = (int)(rng() % 6);
      for (int k = 0; k < extra; k++) {
        uint32_t s = (uint32_t)(rng() % n);
        if (!lens[s]) continue;
        int32_t x = (int32_t)(rng() % lens[s]), y = (int32_t)(rng() % (lens[s] + 1));
        if (x == y) continue;
        q.push_back(s); a.push_back(x); b.push_back(y);
      }
      if (impgx_partitioner_feed(p, q.size(), q.data(), a.data(), b.data()) != 0) { printf("feed failed: %s\n", impgx_last_error()); return 1; }
    }
    if (guard >= 100000) { printf("no termination in trial %d\n", trial); return 1; }
    impgx_partitions *parts = nullptr;
    if (impgx_partitioner_finish(p, &parts) != 0) return 1;
    impgx_partition_view v;
    impgx_partitions_view(parts, &v);
    if (v.partitioned_bp != v.total_bp) { printf("trial %d: %llu of %llu bp\n", trial, (unsigned long long)v.partitioned_bp, (unsigned long long)v.total_bp); return 1; }
    impgx_partitions_free(parts);
    impgx_partitioner_free(p);
  }
  for (int i = 1; i < argc; i++) {
    std::string out = std::string(argv[0]) + ".x" + std::to_string(i) + ".impg";
    const char *paths[1] = {argv[i]};
    if (impgx_impg_write(paths, 1, 1, out.c_str()) != 0) { printf("write failed %s\n", impgx_last_error()); return 1; }
    impgx_impg *f = nullptr;
    if (impgx_impg_open(out.c_str(), &f) != 0) { printf("open failed %s\n", impgx_last_error()); return 1; }
    std::vector<impgx_record> r(impgx_impg_num_records(f));
    impgx_impg_records(f, r.data(), nullptr, nullptr, nullptr);
    impgx_impg_close(f);
    // truncated / corrupted copies must fail cleanly
    FILE *fp = fopen(out.c_str(), "rb"); std::vector<unsigned char> d; int c; while ((c = fgetc(fp)) != EOF) d.push_back((unsigned char)c); fclose(fp);
    for (int k = 0; k < 200; k++) {
      std::vector<unsigned char> e = d;
      if (k % 2) e.resize(rng() % e.size()); else e[rng() % e.size()] ^= (unsigned char)(1 + rng() % 255);
      fp = fopen((std::string(argv[0]) + ".bad.impg").c_str(), "wb"); fwrite(e.data(), 1, e.size(), fp); fclose(fp);
      impgx_impg *g = nullptr;
      if (impgx_impg_open((std::string(argv[0]) + ".bad.impg").c_str(), &g) == 0) impgx_impg_close(g);
    }
  }
  // BEDPE / PAF merge with CIGAR surgery (csrc/host_output.cu) on random chains of alignments
  {
    const char *nm[4] = {"s0#1#a", "s1#1#b:5-900", "s2", "s3#2#c"};
    const uint64_t ln[4] = {100000, 100001, 100002, 100003};
    const char opc[] = {0, 0, 1, 2, 3, 4};  // = = X I D M
    for (int trial = 0; trial < 3000; trial++) {
      std::vector<uint32_t> q, t, cig;
      std::vector<int32_t> qf, ql, tf, tl;
      std::vector<uint64_t> off{0};
      int chains = 1 + (int)(rng() % 4);
      for (int c = 0; c < chains; c++) {
        uint32_t qi = (uint32_t)(rng() % 4), ti = (uint32_t)(rng() % 4);
        bool rev = rng() % 5 < 2;
        long qs = (long)(rng() % 2000), ts = (long)(rng() % 2000);
        int m = 1 + (int)(rng() % 5);
        for (int k = 0; k < m; k++) {
          long qlen = 0, tlen = 0;
          int nops = 1 + (int)(rng() % 6), last = -1;
          std::vector<uint32_t> ops;
          for (int o = 0; o < nops; o++) {
            int op = opc[rng() % 6];
            if (op == last) continue;
            last = op;
            uint32_t len = 1 + (uint32_t)(rng() % 40);
            ops.push_back(IMPGX_RUN(op, len));
            if (op != IMPGX_OP_D) qlen += len;
            if (op != IMPGX_OP_I) tlen += len;
          }
          if (!qlen || !tlen) continue;
          static const long steps[] = {0, 0, 0, 3, 60, 2000, -5, -15};
          long step = steps[rng() % 8];
          if (rev) {
            if (qs - qlen < 0) break;
            q.push_back(qi); qf.push_back((int32_t)qs); ql.push_back((int32_t)(qs - qlen));
            qs = qs - qlen - step;
          } else {
            q.push_back(qi); qf.push_back((int32_t)qs); ql.push_back((int32_t)(qs + qlen));
            qs = qs + qlen + step;
          }
          t.push_back(ti); tf.push_back((int32_t)ts); tl.push_back((int32_t)(ts + tlen));
          cig.insert(cig.end(), ops.begin(), ops.end());
          off.push_back(cig.size());
          ts = ts + tlen + (rng() % 2 ? step : (long)(rng() % 9) - 3);
          if (qs < 0 || ts < 0) break;
        }
      }
      if (q.empty()) continue;
      static const int32_t ds[] = {0, 50, 1000, -1};
      for (int fmt = 1; fmt <= 2; fmt++) {
        char *text = impgx_debug_format_rows(nm, ln, 4, q.size(), q.data(), qf.data(), ql.data(), t.data(), tf.data(), tl.data(),
                                             off.data(), cig.data(), "reg", ds[rng() % 4], fmt, fmt == 1 ? (int)(rng() % 2) : 0);
        if (!text) { printf("format failed: %s\n", impgx_last_error()); return 1; }
        impgx_free(text);
      }
    }
  }
  // index build on the host (csrc/index_host.cu): entry columns, visit ranks, owner map, for random record sets
  // (self alignments, equal starts, empty targets, one shard of several)
  for (int trial = 0; trial < 400; trial++) {
    const uint32_t n_seqs = 1 + (uint32_t)(rng() % 12);
    const size_t n = (size_t)(rng() % 200);
    std::vector<impgx_record> recs(n);
    std::vector<uint64_t> ro(n + 1, 0);
    for (size_t i = 0; i < n; i++) {
      impgx_record &r = recs[i];
      r.query_id = (uint32_t)(rng() % n_seqs);
      r.target_id = rng() % 6 ? (uint32_t)(rng() % n_seqs) : r.query_id;
      r.query_start = (int32_t)(rng() % 50) * 10;
      r.query_end = r.query_start + 1 + (int32_t)(rng() % 500);
      r.target_start = (int32_t)(rng() % 50) * 10;
      r.target_end = r.target_start + 1 + (int32_t)(rng() % 500);
      r.strand = (uint32_t)(rng() % 2);
      r.reserved = 0;
      ro[i + 1] = ro[i] + 1 + rng() % 40;
    }
    const int bidir = (int)(rng() % 2);
    const uint32_t n_ranks = 1 + (uint32_t)(rng() % 4);
    std::vector<uint32_t> owner(n_seqs);
    if (impgx_assign_owners(recs.data(), n, ro.data(), n_seqs, bidir, n_ranks, owner.data()) != 0) { printf("owners: %s\n", impgx_last_error()); return 1; }
    for (uint32_t rank = 0; rank < n_ranks; rank++) {
      long E = impgx_debug_host_columns_shard(recs.data(), n, ro.data(), n_seqs, bidir, nullptr, nullptr, nullptr, nullptr, nullptr,
                                              nullptr, nullptr, nullptr, n_ranks > 1 ? owner.data() : nullptr, rank);
      if (E < 0) { printf("columns: %s\n", impgx_last_error()); return 1; }
      std::vector<int32_t> a(E), b(E), c(E);
      std::vector<uint32_t> d(E), e(E), f(E), g(E);
      std::vector<uint64_t> to(n_seqs + 1);
      impgx_debug_host_columns_shard(recs.data(), n, ro.data(), n_seqs, bidir, a.data(), b.data(), c.data(), d.data(), e.data(),
                                     f.data(), g.data(), to.data(), n_ranks > 1 ? owner.data() : nullptr, rank);
      for (uint32_t s2 = 0; s2 < n_seqs; s2++)
        for (uint64_t k = to[s2] + 1; k < to[s2 + 1]; k++)
          if (a[k - 1] > a[k]) { printf("entries of a target not sorted by start\n"); return 1; }
    }
    std::vector<uint32_t> vr(1 + rng() % 300);
    impgx_debug_visit_ranks(vr.size(), vr.data());
  }
  printf("ok\n");
  return 0;
}
