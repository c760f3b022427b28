// Host-side map containers and the CMAP / alignment-file I/O of the drop-in SegAligner executable.
// Mirrors SAHelper.h (/root/reference/SegAligner/SAHelper.h:8-41) and the map-preparation half of
// SegAligner.h, re-designed around flat arrays that go straight to the GPU (no map-of-vectors copies).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace sah {

// A set of optical maps in ascending-id order (the iteration order of the reference's std::map<int,...>).
// Map m occupies pos[off[m] .. off[m+1]) = [pos_0 .. pos_{n-1}, length]  (SAHelper.cpp:44 pushes every row).
struct MapSet {
    std::vector<int> ids;
    std::vector<int64_t> off{0};
    std::vector<float> pos;
    std::vector<float> prob;  // non-collapse probabilities, labels(m) per map at off[m]-m

    int n_maps() const { return (int)ids.size(); }
    int size(int m) const { return (int)(off[m + 1] - off[m]); }  // vector size in the reference (= labels + 1)
    int labels(int m) const { return size(m) - 1; }
    const float *map(int m) const { return pos.data() + off[m]; }
    const float *probs(int m) const { return prob.data() + (off[m] - m); }
    int index_of(int id) const;  // -1 when absent
    int64_t total_size() const { return off.back(); }
};

// parse_cmap (SAHelper.cpp:17-48) with all of its quirks: '#' lines skipped; column 0 = id (stoi), column 5 =
// position (stof) and only when FOLLOWED by a tab; a row whose position cannot be reached pushes the previous
// position again; a change of id re-creates that map empty; blank lines push the stale position too.
// Returns false when the file cannot be opened (the reference then silently continues with no maps).
bool parse_cmap(const std::string &fname, MapSet &out);

// make_reverse_cmap (SAHelper.cpp:61-81): keeps maps with size() > min_map_len and adds the reversed copy under
// id -id.  Output in ascending id order (negatives first).
MapSet make_reverse_cmap(const MapSet &in, int min_map_len);

// non_collapse_probs (SegAligner.cpp:40-68) over the whole set in id order (the `non_collapse` carry crosses maps).
void non_collapse_probs(MapSet &maps, int inst_gen);

// parse_contig_list (SAHelper.cpp:151-159): whitespace-separated ints, stops at the first non-int.
std::vector<int> parse_contig_list(const std::string &fname);

// ostream-default float formatting ("%g", 6 significant digits) -- what `outfile << float` prints.
int format_float(char *buf, size_t cap, float v);

struct AlnEntry {
    int j, q;
    float s;
};
// print_alignment (SAHelper.cpp:97-135).  aln is end->start like the reference's aln_list; seg_vec_size is
// cmaps_ref[seg_id].size() (labels + 1), used to renumber reverse-orientation labels (:119-123).
bool print_alignment(const std::string &outname, const std::vector<AlnEntry> &aln, int contig_id, int seg_id,
                     int seg_vec_size, float curr_best_score);

struct ScoreRow {
    int seg, contig;
    float score;
};
bool write_all_scores(const std::vector<ScoreRow> &rows, const std::string &prefix);           // SAHelper.cpp:209-218
bool write_score_thresholds(const std::vector<std::pair<int, float>> &thr, const std::string &full_prefix);  // :187-198

// compute_score_thresholds (main.cpp:43-105) for every segment of `segs` (ascending id).  rows = the full
// scoring vector.  Defined behaviour where the reference has UB (window outside the score vector): -NaN.
std::vector<std::pair<int, float>> compute_score_thresholds(const std::vector<ScoreRow> &rows, const MapSet &segs,
                                                            const MapSet &contigs, double pvalue, int n_detect_scores);

// the same for several p-values at once (the grouping and ordering of the scores is shared): out[p][seg index]
void compute_score_thresholds_multi(const std::vector<ScoreRow> &rows, const MapSet &segs, const MapSet &contigs,
                                    const std::vector<double> &pvalues, int n_detect_scores,
                                    std::vector<std::vector<std::pair<int, float>>> &out);

// compute_partial_score_threshold (SegAligner.cpp:23-36)
void partial_score_threshold(float full_threshold, const std::vector<AlnEntry> &aln, const float *contig, int contig_size,
                             float &by_label, float &by_length);
float aln_median(const std::vector<AlnEntry> &aln);  // compute_aln_median (SegAligner.cpp:321-343)
float aln_min(const std::vector<AlnEntry> &aln);     // compute_aln_min (SegAligner.cpp:345-354)

}  // namespace sah
