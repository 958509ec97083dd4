// impgx-query — command-line driver over libimpgx's C ABI that mirrors the
// flags of `impg query` for the PAF -> BED / BEDPE / PAF path (reference
// src/main.rs:6513-7496, option structs :4259-4410). It exists so the drop-in
// can be exercised end to end exactly like the reference binary:
//
//   impgx-query -a X.paf -b regions.bed -x -m 2 -d 1000 -o bed
//
// and `impgx-query partition -a X.paf -w W -d D` mirrors `impg partition -o bed`.
//
// Only the path of SURVEY.md §8 is supported; everything else the reference's
// `query` offers (gfa/maf/fasta outputs, tracepoint inputs, --approximate, …)
// is rejected with an error.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <filesystem>
#include <string>
#include <vector>

#include "../../include/impgx.h"

static void die(const std::string &m) {
  fprintf(stderr, "Error: %s\n", m.c_str());
  exit(1);
}
static void check(int code) {
  if (code != 0) die(impgx_last_error());
}

// parse_merge_distance (src/main.rs:47-55) lives in the library
static int32_t parse_distance(const std::string &s) {
  int32_t v = 0;
  if (impgx_parse_merge_distance(s.c_str(), &v) != 0) die(impgx_last_error());
  return v;
}

// impgx-query partition — `impg partition -o bed` (src/main.rs:4765-4880, :6286-6420;
// src/commands/partition.rs:158-712): partitions.bed, or partition<N>.bed with --separate-files.
static int partition_main(int argc, char **argv) {
  std::vector<std::string> pafs;
  std::string start_file, mode = "longest", folder, out_format = "bed", index_path, index_mode = "auto";
  bool have_d = false, no_merge = false, separate = false, unidirectional = false;
  impgx_partition_params pp;
  memset(&pp, 0, sizeof pp);
  pp.min_missing_size = 3000;
  pp.min_boundary_distance = 3000;
  pp.max_depth = 2;
  pp.min_transitive_len = 101;
  pp.min_distance_between_ranges = 10;
  pp.rehome_singletons = 1;
  pp.min_identity = NAN;
  int device = 0;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&]() -> std::string {
      if (i + 1 >= argc) die("missing value for " + a);
      return argv[++i];
    };
    if (a == "-a" || a == "--alignment-files") {
      pafs.push_back(val());
      while (i + 1 < argc && argv[i + 1][0] != '-') pafs.push_back(argv[++i]);
    } else if (a == "-i" || a == "--index") index_path = val();
    else if (a == "--index-mode") index_mode = val();
    else if (a == "-w" || a == "--window-size") pp.window_size = strtoull(val().c_str(), nullptr, 10);
    else if (a == "--starting-sequences-file") start_file = val();
    else if (a == "--selection-mode") mode = val();
    else if (a == "--min-missing-size") pp.min_missing_size = atoi(val().c_str());
    else if (a == "--min-boundary-distance") pp.min_boundary_distance = atoi(val().c_str());
    else if (a == "--separate-files") separate = true;
    else if (a == "--no-rehome-singletons") pp.rehome_singletons = 0;
    else if (a == "-d" || a == "--merge-distance") { pp.merge_distance = parse_distance(val()); have_d = true; }
    else if (a == "--no-merge") no_merge = true;
    else if (a == "--min-result-identity") pp.min_identity = atof(val().c_str());
    else if (a == "--transitive-dfs") pp.transitive_dfs = 1;
    else if (a == "-m" || a == "--max-depth") pp.max_depth = (uint32_t)atoi(val().c_str());
    else if (a == "--min-transitive-len") pp.min_transitive_len = atoi(val().c_str());
    else if (a == "--min-distance-between-ranges") pp.min_distance_between_ranges = atoi(val().c_str());
    else if (a == "-o" || a == "--output-format") out_format = val();
    else if (a == "--output-folder") folder = val();
    else if (a == "--unidirectional") unidirectional = true;
    else if (a == "--device") device = atoi(val().c_str());
    else if (a == "-h" || a == "--help") {
      printf("usage: impgx-query partition -a X.paf [Y.paf ...] -w WINDOW (-d D | --no-merge)\n"
             "       [--starting-sequences-file FILE] [--selection-mode longest|total|sample[,sep]|haplotype[,sep]]\n"
             "       [--min-missing-size N] [--min-boundary-distance N] [--separate-files] [--no-rehome-singletons]\n"
             "       [-m N] [--min-transitive-len N] [--min-distance-between-ranges N] [--transitive-dfs]\n"
             "       [--min-result-identity F] [--output-folder DIR] [--device N]\n");
      return 0;
    } else die("unsupported option '" + a + "' (only `partition -o bed` over PAF input is implemented)");
  }
  if (pafs.empty()) die("-a/--alignment-files is required");
  if (pp.window_size == 0) die("-w/--window-size is required");
  if (have_d && no_merge) die("-d and --no-merge are mutually exclusive");
  if (!have_d && !no_merge) die("-d/--merge-distance is required. Use `--no-merge` to explicitly disable merging.");
  if (no_merge) pp.merge_distance = -1;
  if (out_format != "bed") die("output format '" + out_format + "' needs sequence files and is outside the accelerated path (bed)");
  if (index_mode != "auto" && index_mode != "single" && index_mode != "per-file")
    die("invalid --index-mode '" + index_mode + "' (auto, single, per-file)");
  // with one sub-index per file the reference's windows are answered by MultiImpg's walk (src/multi_impg.rs:687-755)
  pp.multi_impg = (index_mode == "per-file" || (index_mode == "auto" && pafs.size() >= 100)) ? 1 : 0;
  pp.selection_mode = mode.c_str();

  impgx_index *idx = nullptr;
  {
    std::vector<const char *> paths;
    for (auto &f : pafs) paths.push_back(f.c_str());
    if (!index_path.empty()) check(impgx_index_from_impg(index_path.c_str(), paths.data(), paths.size(), device, &idx));
    else check(impgx_index_from_pafs(paths.data(), paths.size(), unidirectional ? 0 : 1, device, &idx));
  }
  std::vector<uint32_t> starting;
  if (!start_file.empty()) {  // first tab field of every line that is not blank or a comment (:184-212)
    std::ifstream f(start_file);
    if (!f.good()) die("Could not open starting sequences file " + start_file);
    std::string line;
    while (std::getline(f, line)) {
      std::string nm = line.substr(0, line.find('\t'));
      size_t b = nm.find_first_not_of(" \t\r\n"), e = nm.find_last_not_of(" \t\r\n");
      if (b == std::string::npos) continue;
      nm = nm.substr(b, e - b + 1);
      if (nm[0] == '#') continue;
      uint32_t id = 0;
      if (impgx_index_seq_id(idx, nm.c_str(), &id) == 0) starting.push_back(id);  // unknown names are skipped
    }
  }
  pp.starting_seqs = starting.data();
  pp.n_starting_seqs = starting.size();
  impgx_partitions *parts = nullptr;
  check(impgx_partition(idx, &pp, &parts));
  impgx_partition_view v;
  check(impgx_partitions_view(parts, &v));
  if (!folder.empty()) {  // --output-folder: created once, without a shell (reference: std::fs::create_dir_all)
    std::error_code ec;
    std::filesystem::create_directories(folder, ec);
    if (ec) die("cannot create " + folder + ": " + ec.message());
  }
  auto write_file = [&](const std::string &name, int64_t which) {
    std::string path = folder.empty() ? name : folder + "/" + name;
    char *text = impgx_partitions_format_bed(idx, parts, which);
    if (!text) die("formatting failed");
    FILE *f = fopen(path.c_str(), "w");
    if (!f) die("cannot write " + path);
    fputs(text, f);
    fclose(f);
    impgx_free(text);
  };
  if (separate) {
    int64_t last = -1;
    for (size_t i = 0; i < v.n_intervals; i++)
      if ((int64_t)v.partition_num[i] != last) {
        last = v.partition_num[i];
        write_file("partition" + std::to_string(last) + ".bed", last);
      }
  } else if (v.n_intervals) {
    write_file("partitions.bed", -1);
  }
  double pct = v.total_bp ? 100.0 * (double)v.partitioned_bp / (double)v.total_bp : 0.0;
  fprintf(stderr, "Partitioned into %zu regions: %llu bp total written / %llu bp total sequence (%.4g%%), %llu windows\n",
          v.n_partitions, (unsigned long long)v.partitioned_bp, (unsigned long long)v.total_bp, pct,
          (unsigned long long)v.n_windows);
  impgx_partitions_free(parts);
  impgx_index_free(idx);
  return 0;
}

// impgx-query index -a X.paf [Y.paf ...] -i out.impg [--unidirectional] — `impg index` (src/impg.rs:1655-1720)
static int index_main(int argc, char **argv) {
  std::vector<std::string> pafs;
  std::string out;
  bool unidirectional = false;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    if (a == "-a" || a == "--alignment-files") {
      if (i + 1 >= argc) die("missing value for " + a);
      pafs.push_back(argv[++i]);
      while (i + 1 < argc && argv[i + 1][0] != '-') pafs.push_back(argv[++i]);
    } else if (a == "-i" || a == "--index") {
      if (i + 1 >= argc) die("missing value for " + a);
      out = argv[++i];
    } else if (a == "--unidirectional") unidirectional = true;
    else die("unsupported option '" + a + "'");
  }
  if (pafs.empty() || out.empty()) die("usage: impgx-query index -a X.paf [Y.paf ...] -i out.impg [--unidirectional]");
  std::vector<const char *> paths;
  for (auto &f : pafs) paths.push_back(f.c_str());
  check(impgx_impg_write(paths.data(), paths.size(), unidirectional ? 0 : 1, out.c_str()));
  return 0;
}

int main(int argc, char **argv) {
  if (argc > 1 && !strcmp(argv[1], "partition")) return partition_main(argc - 1, argv + 1);
  if (argc > 1 && !strcmp(argv[1], "index")) return index_main(argc - 1, argv + 1);
  std::vector<std::string> pafs;
  std::string bed_path, range_text, out_format = "auto", subset_path, index_mode = "auto", index_path;
  bool transitive = false, dfs = false, unidirectional = false, consider_strand = false, no_merge = false, have_d = false;
  bool original_coords = false;
  int32_t d = 0, min_transitive_len = 101, min_dist = 10, min_out = -1;
  uint32_t max_depth = 2;
  double min_identity = NAN;
  int device = 0;
  for (int i = 1; i < argc; i++) {
    std::string a = argv[i];
    auto val = [&]() -> std::string {
      if (i + 1 >= argc) die("missing value for " + a);
      return argv[++i];
    };
    if (a == "-a" || a == "--alignment-files") {
      pafs.push_back(val());
      while (i + 1 < argc && argv[i + 1][0] != '-') pafs.push_back(argv[++i]);  // -a takes one or more files
    } else if (a == "--alignment-list") {  // one path per line (src/main.rs: resolve_alignment_files)
      const std::string lp = val();
      std::ifstream f(lp);
      if (!f.good()) die("Failed to read alignment list '" + lp + "'");
      std::string line;
      while (std::getline(f, line)) {
        while (!line.empty() && isspace((unsigned char)line.back())) line.pop_back();
        if (!line.empty() && line[0] != '#') pafs.push_back(line);
      }
    } else if (a == "--index-mode") index_mode = val();
    else if (a == "-i" || a == "--index") index_path = val();
    else if (a == "-b" || a == "--target-bed") bed_path = val();
    else if (a == "-r" || a == "--target-range") range_text = val();
    else if (a == "-x" || a == "--transitive") transitive = true;
    else if (a == "--transitive-dfs") dfs = true;
    else if (a == "-m" || a == "--max-depth") max_depth = (uint32_t)atoi(val().c_str());
    else if (a == "-d" || a == "--merge-distance") { d = parse_distance(val()); have_d = true; }
    else if (a == "--no-merge") no_merge = true;
    else if (a == "-l" || a == "--min-output-length") min_out = atoi(val().c_str());
    else if (a == "-o" || a == "--output-format") out_format = val();
    else if (a == "--min-transitive-len") min_transitive_len = atoi(val().c_str());
    else if (a == "--min-distance-between-ranges") min_dist = atoi(val().c_str());
    else if (a == "--unidirectional") unidirectional = true;
    else if (a == "--consider-strandness") consider_strand = true;
    else if (a == "--original-sequence-coordinates") original_coords = true;
    else if (a == "--min-result-identity") min_identity = atof(val().c_str());
    else if (a == "--subset-sequence-list") subset_path = val();
    else if (a == "--device") device = atoi(val().c_str());
    else if (a == "-h" || a == "--help") {
      printf("usage: impgx-query (-a X.paf [Y.paf ...] | --alignment-list FILE) [-i X.impg] [--index-mode auto|single|per-file]\n"
             "       (-b BED | -r seq:start-end) [-x] [-m N] (-d D | --no-merge) [-l L]\n"
             "       [-o auto|bed|bedpe|paf] [--min-transitive-len N] [--min-distance-between-ranges N]\n"
             "       [--transitive-dfs] [--unidirectional] [--consider-strandness] [--min-result-identity F]\n"
             "       [--original-sequence-coordinates]\n"
             "       [--subset-sequence-list FILE] [--device N]\n");
      return 0;
    } else die("unsupported option '" + a + "' (only the PAF -> BED/BEDPE/PAF query path is implemented)");
  }
  if (pafs.empty()) die("-a/--alignment-files or --alignment-list is required");
  if (index_mode != "auto" && index_mode != "single" && index_mode != "per-file")
    die("invalid --index-mode '" + index_mode + "' (auto, single, per-file)");
  // the reference switches to MultiImpg (one sub-index per file, src/multi_impg.rs) with
  // --index-mode per-file or, in auto mode, from 100 alignment files on; its result order and
  // transitive walk differ from Impg's, which the IMPGX_MODE_MULTI_* modes reproduce
  const bool multi = index_mode == "per-file" || (index_mode == "auto" && pafs.size() >= 100);
  if (bed_path.empty() == range_text.empty()) die("exactly one of -r/--target-range and -b/--target-bed is required");
  if (have_d && no_merge) die("-d and --no-merge are mutually exclusive");
  if (!have_d && !no_merge)
    die("-d/--merge-distance is required. Use `--no-merge` to explicitly disable merging.");  // src/main.rs:4288-4315
  if (no_merge) d = -1;
  if (out_format == "auto") out_format = range_text.empty() ? "bedpe" : "bed";  // src/main.rs:7365-7373
  if (out_format != "bed" && out_format != "bedpe" && out_format != "paf")
    die("output format '" + out_format + "' is outside the accelerated path (bed, bedpe, paf)");

  impgx_index *idx = nullptr;
  {
    std::vector<const char *> pp;
    for (auto &f : pafs) pp.push_back(f.c_str());
    // -i: an index file written by `impg index` (or `impgx-query index`) over the same alignment files
    if (!index_path.empty()) check(impgx_index_from_impg(index_path.c_str(), pp.data(), pp.size(), device, &idx));
    else check(impgx_index_from_pafs(pp.data(), pp.size(), unidirectional ? 0 : 1, device, &idx));
  }

  if (original_coords) {
    if (out_format == "paf") die("--original-sequence-coordinates with PAF output needs the sequence files (outside the accelerated path)");
    check(impgx_index_set_original_coordinates(idx, 1));
  }

  // rows
  std::vector<impgx_range> rows;
  std::vector<std::string> names;
  auto add_row = [&](const std::string &seq, int32_t s, int32_t e, const std::string &name) {
    uint32_t id = 0;
    if (impgx_index_seq_id(idx, seq.c_str(), &id) != 0) die("Sequence '" + seq + "' not found in index");
    // validate_sequence_range / validate_range_min_length (src/main.rs:10387-10512)
    if (s < 0) die("Start position " + std::to_string(s) + " cannot be negative");
    if (s >= e) die("Start position " + std::to_string(s) + " must be less than end position " + std::to_string(e));
    if ((uint64_t)e > impgx_index_seq_len(idx, id))
      die("End position " + std::to_string(e) + " exceeds sequence length " + std::to_string(impgx_index_seq_len(idx, id)) +
          " for sequence '" + seq + "'");
    if (e - s < min_transitive_len)
      die("Range '" + name + "' (" + std::to_string(e - s) + " bp) is below minimum of " + std::to_string(min_transitive_len) +
          " bp. Lower --min-transitive-len or use a longer range");
    rows.push_back(impgx_range{id, s, e});
    names.push_back(name);
  };
  if (!range_text.empty()) {
    char seq[4096], name[4200];
    int32_t s, e;
    check(impgx_parse_target_range(range_text.c_str(), seq, sizeof seq, &s, &e, name, sizeof name));
    add_row(seq, s, e, name);
  } else {
    impgx_bed *bed = nullptr;
    check(impgx_bed_parse(bed_path.c_str(), &bed));
    for (size_t i = 0; i < impgx_bed_len(bed); i++)
      add_row(impgx_bed_seq(bed, i), impgx_bed_start(bed, i), impgx_bed_end(bed, i), impgx_bed_name(bed, i));
    impgx_bed_free(bed);
  }

  // --subset-sequence-list: SubsetFilter's matching rules live in the library (src/subset_filter.rs)
  std::vector<uint8_t> mask;
  if (!subset_path.empty()) {
    std::ifstream f(subset_path);
    if (!f.good()) die("Failed to read subset sequence list '" + subset_path + "'");
    std::string text((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    mask.assign(impgx_index_num_seqs(idx), 0);
    if (impgx_subset_mask(idx, text.c_str(), mask.data()) <= 0)
      die("Subset sequence list '" + subset_path + "' did not contain any sequence names");
  }

  impgx_params p;
  memset(&p, 0, sizeof p);
  p.mode = transitive ? (dfs ? IMPGX_MODE_DFS : IMPGX_MODE_BFS) : IMPGX_MODE_QUERY;
  if (multi) p.mode = transitive ? (dfs ? IMPGX_MODE_MULTI_DFS : IMPGX_MODE_MULTI_BFS) : IMPGX_MODE_MULTI_QUERY;
  p.max_depth = max_depth;
  p.min_transitive_len = min_transitive_len;
  p.min_distance_between_ranges = min_dist;
  p.min_output_length = min_out;
  p.store_cigar = out_format == "bed" ? 0 : 1;  // src/main.rs:7447
  p.min_identity = min_identity;
  p.subset_mask = mask.empty() ? nullptr : mask.data();
  p.merge_distance = d;
  p.merge_strands = consider_strand ? 0 : 1;  // src/main.rs:4395-4409

  impgx_results *res = nullptr;
  if (out_format == "bed") check(impgx_query_batch_bed(idx, rows.data(), rows.size(), &p, &res));
  else check(impgx_query_batch(idx, rows.data(), rows.size(), &p, &res));
  if (out_format == "bed") {  // the whole file at once, formatted on every host core
    std::vector<const char *> nm;
    for (auto &s : names) nm.push_back(s.c_str());
    size_t len = 0;
    char *text = impgx_format_bed_batch(idx, res, nm.data(), &len);
    if (!text) die("formatting failed");
    fwrite(text, 1, len, stdout);
    impgx_free(text);
  } else {
    for (size_t r = 0; r < rows.size(); r++) {
      char *text = out_format == "bedpe" ? impgx_format_bedpe(idx, res, r, names[r].c_str(), d)
                                         : impgx_format_paf(idx, res, r, names[r].c_str(), d);
      if (!text) die(impgx_last_error());
      fputs(text, stdout);
      impgx_free(text);
    }
  }
  impgx_results_free(res);
  impgx_index_free(idx);
  return 0;
}
