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
  printf("ok\n");
  return 0;
}
