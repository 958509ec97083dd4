// TEST INFRASTRUCTURE: the text parsers of csrc/api.cu (PAF through impgx_impg_write, BED, target range, CIGAR,
// subset list, merge distance, subsequence names) fed with mutated inputs under ASan / UBSan; see test_host_sanitizers.py.
// fuzz the text parsers of api.cu under ASan/UBSan: PAF (through impgx_impg_write), BED, target range, CIGAR, subset list
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>
#include <vector>
#include "impgx.h"
static std::vector<unsigned char> slurp(const char *p) {
  FILE *f = fopen(p, "rb"); std::vector<unsigned char> d; int c; while ((c = fgetc(f)) != EOF) d.push_back((unsigned char)c); fclose(f); return d;
}
static void spit(const std::string &p, const std::vector<unsigned char> &d) { FILE *f = fopen(p.c_str(), "wb"); fwrite(d.data(), 1, d.size(), f); fclose(f); }
int main(int argc, char **argv) {
  std::mt19937_64 rng(11);
  const std::string tmp = std::string(argv[0]) + ".fuzz";
  const char alphabet[] = "0123456789=XIDM\t\n:-+#cg:Z:chr \r";
  auto mutate = [&](std::vector<unsigned char> d) {
    int k = 1 + (int)(rng() % 8);
    while (k-- && !d.empty()) {
      size_t pos = rng() % d.size();
      switch (rng() % 4) {
        case 0: d[pos] = (unsigned char)alphabet[rng() % (sizeof alphabet - 1)]; break;
        case 1: d.erase(d.begin() + pos); break;
        case 2: d.insert(d.begin() + pos, (unsigned char)alphabet[rng() % (sizeof alphabet - 1)]); break;
        default: d.resize(pos); break;
      }
    }
    return d;
  };
  long accepted = 0;
  for (int i = 1; i < argc; i++) {
    const std::vector<unsigned char> paf = slurp(argv[i]);
    for (int t = 0; t < 300; t++) {
      spit(tmp + ".paf", mutate(paf));
      const std::string in = tmp + ".paf", out = tmp + ".impg";
      const char *paths[1] = {in.c_str()};
      if (impgx_impg_write(paths, 1, (int)(rng() % 2), out.c_str()) == 0) accepted++;
    }
  }
  // BED files
  const std::string bed0 = "chr1\t10\t200\tname\nchr2\t5\t6\n#c\nchr3\t1\t2\t.\n";
  for (int t = 0; t < 2000; t++) {
    std::vector<unsigned char> d(bed0.begin(), bed0.end());
    spit(tmp + ".bed", mutate(d));
    impgx_bed *b = nullptr;
    if (impgx_bed_parse((tmp + ".bed").c_str(), &b) == 0) {
      for (size_t k = 0; k < impgx_bed_len(b); k++) { (void)impgx_bed_seq(b, k); (void)impgx_bed_name(b, k); (void)impgx_bed_start(b, k); }
      impgx_bed_free(b);
    }
  }
  // target ranges, CIGARs, subset lists, merge distances, subsequence names
  const std::string seeds[] = {"chr1:10-200", "a#1#b:0-250:5-9", "12=3X4I5D6M", "50k", "HG1#1#chr2\nHG2_hap2_x\n", "x:1-2"};
  for (int t = 0; t < 20000; t++) {
    const std::string &sd = seeds[rng() % 6];
    std::vector<unsigned char> d = mutate(std::vector<unsigned char>(sd.begin(), sd.end()));
    d.erase(std::remove(d.begin(), d.end(), (unsigned char)0), d.end());
    std::string s(d.begin(), d.end());
    char seq[64], name[64];
    int32_t a, b;
    (void)impgx_parse_target_range(s.c_str(), seq, sizeof seq, &a, &b, name, sizeof name);
    std::vector<uint32_t> runs(s.size() + 1);
    (void)impgx_parse_cigar(s.c_str(), s.size(), runs.data(), runs.size());
    (void)impgx_parse_cigar(s.c_str(), s.size(), runs.data(), 1);
    (void)impgx_subset_matches(s.c_str(), "HG1#1#chr2:5-6");
    (void)impgx_subset_matches("HG1#1\nchr2\n", s.c_str());
    (void)impgx_parse_merge_distance(s.c_str(), &a);
    char base[8];
    (void)impgx_parse_subsequence_coordinates(s.c_str(), base, sizeof base, &a);
  }
  printf("ok (%ld mutated PAFs still parsed)\n", accepted);
  return 0;
}
