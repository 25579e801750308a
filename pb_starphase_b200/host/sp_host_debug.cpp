// sp_host_debug.cpp -- hla_debug.json (src/hla/debug.rs), the consensus preparation in front of score_read
// (splice_read, strand handling) and the hemizygous test.  Host-only logic: no alignment is computed here.
#include <algorithm>
#include <cmath>
#include <unordered_map>

#include "starphase_host.hpp"

namespace starphase {

// ------------------------------------------------------------------------------------------
// DetailedMappingStats: the minimap2::Mapping fields src/hla/debug.rs:161-181 reads
// ------------------------------------------------------------------------------------------
std::string cigar_string(const std::vector<std::pair<uint32_t, uint8_t>> &cigar) {
    static const char ops[] = "MIDNSHP=XB";
    std::string out;
    for (const auto &op : cigar) {
        if (op.second > 9) throw HostError("Unexpected cigar type: " + std::to_string(op.second));
        out += std::to_string(op.first);
        out.push_back(ops[op.second]);
    }
    return out;
}

// nt4 code of minimap2's seq_nt4_table: A/C/G/T (either case) 0..3, everything else 4; MD prints "ACGTN"[code]
static int nt4(char c) {
    switch (c) {
        case 'A': case 'a': return 0;
        case 'C': case 'c': return 1;
        case 'G': case 'g': return 2;
        case 'T': case 't': return 3;
        default: return 4;
    }
}

std::string md_string(const std::vector<std::pair<uint32_t, uint8_t>> &cigar, const std::string &target, size_t t_off,
                      const std::string &query, size_t q_off) {
    // minimap2 2.28 format.c write_MD_core: runs of equal bases as counts, a mismatch as the target base, a deletion
    // as ^ + the deleted target bases; a trailing zero count is not written
    std::string out;
    size_t run = 0;
    for (const auto &op : cigar) {
        const size_t len = op.first;
        switch (op.second) {
            case 0: case 7: case 8:
                if (t_off + len > target.size() || q_off + len > query.size()) throw HostError("md_string: cigar runs past a sequence");
                for (size_t j = 0; j < len; ++j) {
                    const int tq = nt4(target[t_off + j]);
                    if (nt4(query[q_off + j]) != tq) {
                        out += std::to_string(run);
                        out.push_back("ACGTN"[tq]);
                        run = 0;
                    } else {
                        ++run;
                    }
                }
                t_off += len; q_off += len;
                break;
            case 1: q_off += len; break;
            case 2:
                if (t_off + len > target.size()) throw HostError("md_string: cigar runs past the target");
                out += std::to_string(run);
                out.push_back('^');
                for (size_t j = 0; j < len; ++j) out.push_back("ACGTN"[nt4(target[t_off + j])]);
                run = 0;
                t_off += len;
                break;
            case 3: t_off += len; break;
            default: throw HostError("Unexpected cigar type: " + std::to_string(op.second));
        }
    }
    if (run > 0) out += std::to_string(run);
    return out;
}

DetailedMappingStats detailed_mapping_stats(const Mapping &m, const std::string &target, const std::string &query) {
    DetailedMappingStats d;
    d.query_len = m.query_len;
    d.target_len = m.target_len;
    for (const auto &op : m.cigar)
        if (op.second == 7) d.match_len += op.first;  // mm_reg1_t::mlen: matching bases of the alignment
    d.nm = m.nm;
    d.query_unmapped = m.query_len - (m.query_end - m.query_start);
    d.target_unmapped = m.target_len - (m.target_end - m.target_start);
    d.cigar = cigar_string(m.cigar);
    d.md = md_string(m.cigar, target, m.target_start, query, m.query_start);
    return d;
}

Json DetailedMappingStats::to_json() const {
    Json j = Json::object();
    j.set("query_len", query_len).set("target_len", target_len).set("match_len", match_len).set("nm", nm);
    j.set("query_unmapped", query_unmapped).set("target_unmapped", target_unmapped).set("cigar", cigar).set("md", md);
    return j;
}

Json PairedMappingStats::to_json() const {
    Json j = Json::object();
    j.set("cdna_mapping", cdna_mapping ? cdna_mapping->to_json() : Json()).set("dna_mapping", dna_mapping ? dna_mapping->to_json() : Json());
    return j;
}

void ReadMappingStats::add_mapping(const std::string &hla_id, std::optional<DetailedMappingStats> cdna, std::optional<DetailedMappingStats> dna) {
    if (mapping_stats_.count(hla_id)) throw HostError("Entry " + hla_id + " is already occupied!");
    PairedMappingStats p;
    p.cdna_mapping = std::move(cdna);
    p.dna_mapping = std::move(dna);
    mapping_stats_.emplace(hla_id, std::move(p));
}

static Json opt_string(const std::optional<std::string> &s) { return s ? Json(*s) : Json(); }

Json ReadMappingStats::to_json() const {
    Json maps = Json::object();
    for (const auto &kv : mapping_stats_) maps.set(kv.first, kv.second.to_json());
    Json j = Json::object();
    j.set("best_match_id", opt_string(best_match_id_)).set("best_match_star", opt_string(best_match_star_)).set("mapping_stats", maps);
    return j;
}

DualPassingStats DualPassingStats::new_dual(bool is_passing, size_t c1, size_t c2, double maf, double cdf) {
    DualPassingStats s;
    s.is_passing = is_passing; s.is_dual = true;
    s.counts1 = c1; s.counts2 = c2; s.maf = maf; s.cdf = cdf;
    return s;
}

Json DualPassingStats::to_json() const {
    Json j = Json::object();
    j.set("is_passing", is_passing).set("is_dual", is_dual);
    j.set("counts1", counts1 ? Json(*counts1) : Json()).set("counts2", counts2 ? Json(*counts2) : Json());
    j.set("maf", maf ? Json::number(*maf) : Json()).set("cdf", cdf ? Json::number(*cdf) : Json());
    return j;
}

DualPassingStats dual_passing_stats(bool is_dual, size_t counts1, size_t counts2, double min_consensus_fraction, double min_cdf,
                                    double expected_maf) {
    if (!is_dual) return DualPassingStats::new_non_dual();
    const size_t total = counts1 + counts2, minor = std::min(counts1, counts2);
    const double maf = static_cast<double>(minor) / static_cast<double>(total);
    const double cdf = binomial_cdf(total, expected_maf, minor);
    return DualPassingStats::new_dual(maf >= min_consensus_fraction && cdf >= min_cdf, counts1, counts2, maf, cdf);
}

void HlaDebug::add_read(const std::string &gene, const std::string &qname, ReadMappingStats stats) {
    auto &g = read_mapping_stats_[gene];
    if (g.count(qname)) throw HostError("Entry " + qname + " is already occupied");
    g.emplace(qname, std::move(stats));
}

void HlaDebug::add_dual_passing_stats(const std::string &gene, DualPassingStats stats) {
    if (!dual_passing_stats_) dual_passing_stats_.emplace();
    if (dual_passing_stats_->count(gene)) throw HostError("Entry " + gene + " is already occupied");
    dual_passing_stats_->emplace(gene, std::move(stats));
}

Json HlaDebug::to_json() const {
    Json reads = Json::object();
    for (const auto &g : read_mapping_stats_) {
        Json per = Json::object();
        for (const auto &r : g.second) per.set(r.first, r.second.to_json());
        reads.set(g.first, per);
    }
    Json dual;
    if (dual_passing_stats_) {
        dual = Json::object();
        for (const auto &g : *dual_passing_stats_) dual.set(g.first, g.second.to_json());
    }
    Json j = Json::object();
    j.set("read_mapping_stats", reads).set("dual_passing_stats", dual);
    return j;
}

// ------------------------------------------------------------------------------------------
// homopolymer compression (src/util/homopolymers.rs:18-42)
// ------------------------------------------------------------------------------------------
std::string hpc(const std::string &sequence) {
    std::string out;
    for (char c : sequence)
        if (out.empty() || out.back() != c) out.push_back(c);
    return out;
}

size_t hpc_pos(const std::string &sequence, size_t position) {
    size_t total = 0, offset = 0, i = 0;
    while (i < sequence.size()) {
        size_t j = i;
        while (j < sequence.size() && sequence[j] == sequence[i]) ++j;
        total += j - i;
        if (position < total) break;
        ++offset;
        i = j;
    }
    return offset;
}

// ------------------------------------------------------------------------------------------
// consensus preparation
// ------------------------------------------------------------------------------------------
std::string reverse_complement(const std::string &s) {
    std::string out(s.size(), 'N');
    for (size_t i = 0; i < s.size(); ++i) {
        const char c = s[s.size() - 1 - i];
        switch (c) {
            case 'A': out[i] = 'T'; break;
            case 'C': out[i] = 'G'; break;
            case 'G': out[i] = 'C'; break;
            case 'T': out[i] = 'A'; break;
            case 'N': out[i] = 'N'; break;
            default: throw HostError("Unexpected character for reverse-complement: " + std::to_string(static_cast<unsigned char>(c)));
        }
    }
    return out;
}

std::pair<std::vector<std::pair<size_t, size_t>>, size_t> splice_segments(size_t sequence_len, int64_t pos,
                                                                          const std::vector<std::pair<uint32_t, uint8_t>> &cigar,
                                                                          const std::vector<std::pair<uint64_t, uint64_t>> &exons) {
    // reference coordinate -> read coordinate for aligned (M / = / X) columns: rust-htslib's aligned_pairs()
    std::unordered_map<uint64_t, size_t> lookup;
    size_t q = 0;
    int64_t r = pos;
    for (const auto &op : cigar) {
        const size_t len = op.first;
        switch (op.second) {
            case 0: case 7: case 8:
                for (size_t k = 0; k < len; ++k) lookup[static_cast<uint64_t>(r + static_cast<int64_t>(k))] = q + k;  // later pairs overwrite
                q += len; r += static_cast<int64_t>(len);
                break;
            case 1: case 4: q += len; break;
            case 2: case 3: r += static_cast<int64_t>(len); break;
            case 5: case 6: break;
            default: throw HostError("Unexpected cigar type: " + std::to_string(op.second));
        }
    }
    if (q > sequence_len) throw HostError("splice_read: cigar is longer than the sequence");
    size_t offset = 0;
    std::vector<std::pair<size_t, size_t>> segments;
    for (const auto &ex : exons) {
        // usize arithmetic of :1541-1552 (`last -= 1` below zero would panic in the reference; exons start above 0)
        uint64_t first = ex.first, last = ex.second - 1;
        while (!lookup.count(first) && first <= last) ++first;
        while (!lookup.count(last) && first <= last) --last;
        if (segments.empty()) offset += static_cast<size_t>(first - ex.first);
        if (first <= last) segments.emplace_back(lookup.at(first), lookup.at(last) + 1);
    }
    for (const auto &s : segments)
        if (s.first > s.second || s.second > sequence_len) throw HostError("splice_read: slice index out of range");
    return {segments, offset};
}

std::pair<std::string, size_t> splice_read(const std::string &sequence, int64_t pos, const std::vector<std::pair<uint32_t, uint8_t>> &cigar,
                                           const std::vector<std::pair<uint64_t, uint64_t>> &exons) {
    const auto seg = splice_segments(sequence.size(), pos, cigar, exons);
    std::string spliced;
    for (const auto &s : seg.first) spliced.append(sequence, s.first, s.second - s.first);
    return {spliced, seg.second};
}

ScoreReadTargets prepare_score_read_targets(const std::string &read_sequence, int64_t pos, const std::vector<std::pair<uint32_t, uint8_t>> &cigar,
                                            const std::vector<std::pair<uint64_t, uint64_t>> &exons, bool is_forward_strand,
                                            const DiplotypeSettings &settings) {
    ScoreReadTargets t;
    t.dna_target = is_forward_strand ? read_sequence : reverse_complement(read_sequence);
    if (settings.disable_cdna_scoring) {
        t.cdna_target = "N";
    } else {
        const std::string fw = splice_read(read_sequence, pos, cigar, exons).first;
        if (fw.empty()) t.cdna_target = "N";
        else t.cdna_target = is_forward_strand ? fw : reverse_complement(fw);
    }
    return t;
}

// ------------------------------------------------------------------------------------------
// is_hemizygous_better
// ------------------------------------------------------------------------------------------
double binomial_ln_pmf(uint64_t n, double p, uint64_t x) {  // statrs 0.16 distribution/binomial.rs ln_pmf
    const double ninf = -std::numeric_limits<double>::infinity();
    if (x > n) return ninf;
    if (p == 0.0) return x == 0 ? 0.0 : ninf;
    if (p == 1.0) return x == n ? 0.0 : ninf;
    const double ln_binom = ln_factorial(n) - ln_factorial(x) - ln_factorial(n - x);  // factorial::ln_binomial
    return ln_binom + static_cast<double>(x) * std::log(p) + static_cast<double>(n - x) * std::log(1.0 - p);
}

double normal_ln_pdf(double mean, double std_dev, double x) {  // statrs 0.16 distribution/normal.rs ln_pdf
    static const double kLnSqrt2Pi = 0.91893853320467274178032973640561763986139747363778341281715;
    const double d = (x - mean) / std_dev;
    return (-0.5 * d * d) - kLnSqrt2Pi - std::log(std_dev);
}

bool is_hemizygous_better(const std::vector<std::optional<size_t>> &scores1, const std::vector<std::optional<size_t>> &scores2,
                          const std::vector<bool> &is_consensus1, bool is_dual, size_t dual_max_ed_delta,
                          std::optional<double> normalized_coverage) {
    const size_t read_count = is_consensus1.size();
    size_t min_ed = 0;
    if (is_dual) {
        if (scores1.size() != read_count || scores2.size() != read_count) throw HostError("is_hemizygous_better: score vectors differ in length");
        size_t c1 = 0, c2 = 0;
        for (size_t r = 0; r < read_count; ++r) {
            if (!scores1[r] && !scores2[r]) throw HostError("assertion failed: o1.is_some() || o2.is_some()");
            const size_t s1 = scores1[r] ? *scores1[r] : scores2[r].value_or(0) + dual_max_ed_delta;
            const size_t s2 = scores2[r] ? *scores2[r] : scores1[r].value_or(0) + dual_max_ed_delta;
            const size_t mn = std::min(s1, s2);
            c1 += s1 - mn;
            c2 += s2 - mn;
        }
        min_ed = std::min(c1, c2);
    }
    const double haploid_ed_cost = 2.0 * static_cast<double>(min_ed);
    const double rc = static_cast<double>(read_count);
    const double nc_hap = normalized_coverage.value_or(rc);
    const double nc_dev = nc_hap * 0.1;
    // Normal::new rejects a non-positive or NaN deviation ("Bad distribution parameters")
    if (!(nc_dev > 0.0) || std::isnan(nc_hap)) throw HostError("Bad distribution parameters");
    const double haploid_cost = haploid_ed_cost + std::fabs(normal_ln_pdf(nc_hap, nc_dev, rc));
    size_t obs1 = 0;
    for (bool b : is_consensus1) obs1 += b ? 1 : 0;
    const double balance = is_dual ? 2.0 * std::fabs(binomial_ln_pmf(read_count, 0.5, obs1)) : 0.0;
    const double nc_dip = 2.0 * normalized_coverage.value_or(rc);
    const double diploid_cost = balance + std::fabs(normal_ln_pdf(nc_dip, nc_dev, rc));
    return haploid_cost < diploid_cost;
}

}  // namespace starphase
