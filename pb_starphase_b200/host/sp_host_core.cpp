// sp_host_core.cpp -- scores, mapping selection, processed CIGARs, statistics, JSON writer.
// Each function follows the reference file:line named in starphase_host.hpp.
#include <algorithm>
#include <charconv>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <limits>

#include "starphase_host.hpp"

namespace starphase {

AlignerStandIns &aligner_stand_ins() {
    static AlignerStandIns s;
    return s;
}


// ------------------------------------------------------------------------------------------
// Json: serde_json::ser::PrettyFormatter (empty containers print as [] / {})
// ------------------------------------------------------------------------------------------
static void json_escape(const std::string &s, std::string &out) {
    out.push_back('"');
    for (unsigned char ch : s) {
        switch (ch) {
            case '"': out += "\\\""; break;
            case '\\': out += "\\\\"; break;
            case '\n': out += "\\n"; break;
            case '\r': out += "\\r"; break;
            case '\t': out += "\\t"; break;
            case 8: out += "\\b"; break;
            case 12: out += "\\f"; break;
            default:
                if (ch < 0x20) {
                    char buf[8];
                    std::snprintf(buf, sizeof buf, "\\u%04x", ch);
                    out += buf;
                } else {
                    out.push_back(static_cast<char>(ch));
                }
        }
    }
    out.push_back('"');
}

// f64 the way serde_json prints it: ryu's shortest round-trip digits (std::to_chars gives the same digit string) laid
// out by ryu::pretty::format64 -- plain decimals for exponents in (-5, 16], scientific otherwise, always with ".0" on
// integral values; NaN / infinity become null (serde_json::ser::Formatter::write_f64 via Value)
static std::string format_f64(double v) {
    if (!std::isfinite(v)) return "null";
    if (v == 0.0) return std::signbit(v) ? "-0.0" : "0.0";
    char buf[64];
    const auto res = std::to_chars(buf, buf + sizeof buf, std::fabs(v), std::chars_format::scientific);
    const std::string sci(buf, res.ptr);  // d[.ddd]e[+-]xx
    const size_t epos = sci.find('e');
    std::string digits;
    for (size_t i = 0; i < epos; ++i)
        if (sci[i] != '.') digits.push_back(sci[i]);
    const int exp10 = std::atoi(sci.c_str() + epos + 1);
    const int length = static_cast<int>(digits.size());
    const int kk = exp10 + 1;       // position of the decimal point relative to the digit string
    const int k = kk - length;      // value = digits * 10^k
    std::string out = std::signbit(v) ? "-" : "";
    if (0 <= k && kk <= 16) {
        out += digits + std::string(static_cast<size_t>(k), '0') + ".0";
    } else if (0 < kk && kk <= 16) {
        out += digits.substr(0, static_cast<size_t>(kk)) + "." + digits.substr(static_cast<size_t>(kk));
    } else if (-5 < kk && kk <= 0) {
        out += "0." + std::string(static_cast<size_t>(-kk), '0') + digits;
    } else if (length == 1) {
        out += digits + "e" + std::to_string(kk - 1);
    } else {
        out += digits.substr(0, 1) + "." + digits.substr(1) + "e" + std::to_string(kk - 1);
    }
    return out;
}

std::string Json::pretty(int indent) const {
    const std::string pad(static_cast<size_t>(2 * (indent + 1)), ' '), end(static_cast<size_t>(2 * indent), ' ');
    std::string out;
    switch (kind_) {
        case Null: return "null";
        case Bool: return i_ ? "true" : "false";
        case Int: return std::to_string(i_);
        case Float: return format_f64(f_);
        case Str: json_escape(s_, out); return out;
        case Arr:
            if (a_.empty()) return "[]";
            out = "[\n";
            for (size_t k = 0; k < a_.size(); ++k) {
                out += pad + a_[k].pretty(indent + 1);
                out += k + 1 < a_.size() ? ",\n" : "\n";
            }
            return out + end + "]";
        case Obj:
            if (o_.empty()) return "{}";
            out = "{\n";
            for (size_t k = 0; k < o_.size(); ++k) {
                out += pad;
                json_escape(o_[k].first, out);
                out += ": " + o_[k].second.pretty(indent + 1);
                out += k + 1 < o_.size() ? ",\n" : "\n";
            }
            return out + end + "}";
    }
    return out;
}

static Json opt_usize(const std::optional<size_t> &v) { return v ? Json(static_cast<long long>(*v)) : Json(); }

// ------------------------------------------------------------------------------------------
// scores
// ------------------------------------------------------------------------------------------
static double score_value(size_t mapping_len, size_t nm, size_t unmapped) {  // src/data_types/mapping.rs:191-195
    return std::max(static_cast<double>(nm + unmapped), 0.1) / static_cast<double>(mapping_len);
}

double harmonic_mean(const std::vector<double> &scores) {
    double sum = 0.0;
    for (double v : scores) {
        if (!(v > 0.0)) throw HostError("dna_score must be > 0.0");
        sum += 1.0 / v;
    }
    return sum > 0.0 ? static_cast<double>(scores.size()) / sum : 0.0;
}

double MappingStats::custom_score(bool penalize_unmapped) const {
    if (penalize_unmapped) return score_value(seq_len, nm, unmapped);
    return score_value(seq_len - unmapped, nm, 0);
}

Json MappingStats::to_json() const {
    Json j = Json::object();
    j.set("seq_len", static_cast<long long>(seq_len)).set("nm", static_cast<long long>(nm)).set("unmapped", static_cast<long long>(unmapped));
    j.set("clipped_start", opt_usize(clipped_start)).set("clipped_end", opt_usize(clipped_end));
    return j;
}

std::pair<double, double> HlaMappingStats::mapping_score() const {
    return {cdna_stats ? cdna_stats->mapping_score() : 1.0, dna_stats ? dna_stats->mapping_score() : 1.0};
}

Json HlaMappingStats::to_json() const {
    Json j = Json::object();
    j.set("cdna_stats", cdna_stats ? cdna_stats->to_json() : Json()).set("dna_stats", dna_stats ? dna_stats->to_json() : Json());
    return j;
}

std::pair<std::optional<size_t>, MappingStats> select_best_mapping(const std::vector<Mapping> &mappings, bool unmapped_from_target,
                                                                   bool penalize_unmapped, std::optional<size_t> base_length_override) {
    const size_t o = base_length_override.value_or(1);
    MappingStats best(o, o, 0);
    std::optional<size_t> best_idx;
    for (size_t idx = 0; idx < mappings.size(); ++idx) {
        const Mapping &m = mappings[idx];
        size_t bl, um;
        if (unmapped_from_target) {
            bl = base_length_override.value_or(m.target_len);
            um = bl - (m.target_end - m.target_start);
        } else {
            bl = base_length_override.value_or(m.query_len);
            um = bl - (m.query_end - m.query_start);
        }
        const MappingStats stats(bl, m.nm, um);
        if (stats.custom_score(penalize_unmapped) < best.custom_score(penalize_unmapped)) {
            best = stats;
            best_idx = idx;
        }
    }
    return {best_idx, best};
}

// ------------------------------------------------------------------------------------------
// processed matches
// ------------------------------------------------------------------------------------------
std::vector<size_t> process_mm_cigar(const std::vector<std::pair<uint32_t, uint8_t>> &cigar, size_t target_offset, size_t target_len,
                                     size_t clip_start, size_t clip_end) {
    const size_t zero_padding = target_offset > clip_start ? target_offset - clip_start : 0;
    const size_t nm_padding = target_offset - zero_padding;
    std::vector<size_t> ret(zero_padding + 1, 0);
    ret.reserve(target_len + 1);
    size_t cur = 0;
    for (size_t i = 0; i < nm_padding; ++i) ret.push_back(++cur);
    for (const auto &op : cigar) {
        switch (op.second) {
            case 1: cur += op.first; break;                                               // I
            case 2: case 8: for (uint32_t i = 0; i < op.first; ++i) ret.push_back(++cur); break;  // D | X
            case 7: ret.insert(ret.end(), op.first, cur); break;                          // =
            default: throw HostError("Unexpected cigar type: " + std::to_string(op.second));
        }
    }
    if (ret.size() > target_len + 1) throw HostError("process_mm_cigar: cigar longer than the target");
    const size_t missing = target_len + 1 - ret.size();
    const size_t ext = std::min(clip_end, missing);
    for (size_t i = 0; i < ext; ++i) ret.push_back(++cur);
    ret.insert(ret.end(), missing - ext, cur);
    return ret;
}

HlaProcessedMatch HlaProcessedMatch::worst_match(size_t num_sequences) {
    HlaProcessedMatch m("");
    m.full_mapping_stats_.assign(num_sequences, std::nullopt);
    m.processed_cigars_.assign(num_sequences, std::nullopt);
    m.processed_ranges_.assign(num_sequences, {0, 0});
    return m;
}

void HlaProcessedMatch::add_mapping(const std::optional<Mapping> &mapping) {
    if (!mapping) {
        full_mapping_stats_.push_back(std::nullopt);
        processed_cigars_.push_back(std::nullopt);
        processed_ranges_.push_back({0, 0});
        return;
    }
    const Mapping &m = *mapping;
    if (!m.forward) throw HostError("Reverse strand mappings are not supported by HlaProcessedMatch");
    const size_t clip_start = m.query_start, clip_end = m.query_len - m.query_end;
    std::vector<size_t> pc = process_mm_cigar(m.cigar, m.target_start, m.target_len, clip_start, clip_end);
    const size_t pc_start = m.target_start > clip_start ? m.target_start - clip_start : 0;
    const size_t pc_end = m.target_end + std::min(clip_end, m.target_len - m.target_end);
    const size_t unmapped = m.query_len - (m.query_end - m.query_start);
    if (pc.size() != m.target_len + 1) throw HostError("processed cigar has the wrong length");
    full_mapping_stats_.push_back(MappingStats(m.query_len, m.nm, unmapped));
    processed_cigars_.push_back(std::move(pc));
    processed_ranges_.push_back({pc_start, pc_end});
}

bool HlaProcessedMatch::is_better_match(const HlaProcessedMatch &rhs) const {
    if (processed_cigars_.size() != rhs.processed_cigars_.size()) throw HostError("RHS has different processed cigar length");
    for (size_t i = 0; i < processed_cigars_.size(); ++i) {
        const auto &l = processed_cigars_[i], &r = rhs.processed_cigars_[i];
        if (l && r) {
            const size_t os = std::max(processed_ranges_[i].first, rhs.processed_ranges_[i].first);
            const size_t oe = std::min(processed_ranges_[i].second, rhs.processed_ranges_[i].second);
            size_t lnm = 0, rnm = 0;
            if (os < oe) { lnm = (*l)[oe] - (*l)[os]; rnm = (*r)[oe] - (*r)[os]; }
            if (lnm < rnm) return true;
            if (lnm > rnm) return false;
        } else if (!l && !r) {
            continue;
        } else {
            return static_cast<bool>(l);
        }
    }
    if (full_mapping_stats_.size() != 2 || rhs.full_mapping_stats_.size() != 2) throw HostError("expected cDNA and DNA entries");
    HlaMappingStats a, b;
    a.cdna_stats = full_mapping_stats_[0]; a.dna_stats = full_mapping_stats_[1];
    b.cdna_stats = rhs.full_mapping_stats_[0]; b.dna_stats = rhs.full_mapping_stats_[1];
    return a.mapping_score() < b.mapping_score();
}

// ------------------------------------------------------------------------------------------
// statistics: statrs 0.16 ln_gamma (Lanczos, Math.NET coefficients), ln_factorial (171-entry table)
// ------------------------------------------------------------------------------------------
static const double kGammaR = 10.900511;
static const double kGammaDk[11] = {
    2.48574089138753565546e-5, 1.05142378581721974210, -3.45687097222016235469, 4.51227709466894823700,
    -2.98285225323576655721, 1.05639711577126713077, -1.95428773191645869583e-1, 1.70970543404441224307e-2,
    -5.71926117404305781283e-4, 4.63399473359905636708e-6, -2.71994908488607703910e-9,
};
static const double kLn2SqrtEOverPi = 0.6207822376352452223455184457816472122518527279025978;
static const double kLnPi = 1.1447298858494001741434273513530587116472948129153;

double ln_gamma(double x) {
    if (x < 0.5) {
        double s = kGammaDk[0];
        for (int i = 1; i < 11; ++i) s += kGammaDk[i] / (static_cast<double>(i) - x);
        return kLnPi - std::log(std::sin(M_PI * x)) - std::log(s) - kLn2SqrtEOverPi - (0.5 - x) * std::log((0.5 - x + kGammaR) / M_E);
    }
    double s = kGammaDk[0];
    for (int i = 1; i < 11; ++i) s += kGammaDk[i] / (x + static_cast<double>(i) - 1.0);
    return std::log(s) + kLn2SqrtEOverPi + (x - 0.5) * std::log((x - 0.5 + kGammaR) / M_E);
}

double ln_factorial(uint64_t x) {
    static const std::vector<double> cache = [] {
        std::vector<double> c(171, 1.0);
        for (int i = 1; i < 171; ++i) c[static_cast<size_t>(i)] = c[static_cast<size_t>(i) - 1] * static_cast<double>(i);
        return c;
    }();
    if (x < cache.size()) return std::log(cache[x]);
    return ln_gamma(static_cast<double>(x) + 1.0);
}

double multinomial_ln_pmf(const std::vector<double> &probs, const std::vector<uint64_t> &obs) {
    if (probs.size() != obs.size()) throw HostError("multinomial: probs and obs differ in length");
    uint64_t total = 0;
    for (uint64_t o : obs) total += o;
    double coeff = ln_factorial(total);
    for (uint64_t o : obs) coeff -= ln_factorial(o);
    double acc = 0.0;
    for (size_t i = 0; i < probs.size(); ++i) acc = acc + static_cast<double>(obs[i]) * std::log(probs[i]);  // ln(0) = -inf as in Rust
    return coeff + acc;
}

// statrs 0.16 function::beta::beta_reg: regularized incomplete beta by the modified Lentz continued fraction
// (Math.NET's BetaRegularized), 140 iterations at most, symmetry transform above (a + 1) / (a + b + 2)
double beta_reg(double a, double b, double x) {
    if (!(a > 0.0) || !(b > 0.0) || !(x >= 0.0 && x <= 1.0)) throw HostError("beta_reg: arguments out of range");
    const double bt = (x == 0.0 || x == 1.0)
                          ? 0.0
                          : std::exp(ln_gamma(a + b) - ln_gamma(a) - ln_gamma(b) + a * std::log(x) + b * std::log(1.0 - x));
    const bool symm = x >= (a + 1.0) / (a + b + 2.0);
    const double eps = 0.00000000000000011102230246251565;  // prec::F64_PREC
    const double fpmin = std::numeric_limits<double>::min() / eps;
    if (symm) { const double swap = a; x = 1.0 - x; a = b; b = swap; }
    const double qab = a + b, qap = a + 1.0, qam = a - 1.0;
    double c = 1.0, d = 1.0 - qab * x / qap;
    if (std::fabs(d) < fpmin) d = fpmin;
    d = 1.0 / d;
    double h = d;
    for (int mi = 1; mi < 141; ++mi) {
        const double m = static_cast<double>(mi), m2 = m * 2.0;
        double aa = m * (b - m) * x / ((qam + m2) * (a + m2));
        d = 1.0 + aa * d;
        if (std::fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c;
        if (std::fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        h = h * d * c;
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2));
        d = 1.0 + aa * d;
        if (std::fabs(d) < fpmin) d = fpmin;
        c = 1.0 + aa / c;
        if (std::fabs(c) < fpmin) c = fpmin;
        d = 1.0 / d;
        const double del = d * c;
        h *= del;
        if (std::fabs(del - 1.0) <= eps) break;
    }
    return symm ? 1.0 - bt * h / a : bt * h / a;
}

double binomial_cdf(uint64_t n, double p, uint64_t k) {  // statrs 0.16 Binomial::cdf
    if (k >= n) return 1.0;
    return beta_reg(static_cast<double>(n) - static_cast<double>(k), static_cast<double>(k) + 1.0, 1.0 - p);
}

bool is_passing_dual(size_t counts1, size_t counts2, double min_consensus_fraction, double min_cdf, double expected_maf) {
    const size_t total = counts1 + counts2, minor = std::min(counts1, counts2);
    const double maf = static_cast<double>(minor) / static_cast<double>(total);
    const double cdf = binomial_cdf(total, expected_maf, minor);
    return maf >= min_consensus_fraction && cdf >= min_cdf;
}

std::string starphase_json(const std::string &pbstarphase_version, const std::map<std::string, std::string> &database_metadata,
                           const std::map<std::string, Json> &gene_details) {
    Json md = Json::object();
    for (const char *k : {"pbstarphase_version", "cpic_version", "hla_version", "pharmvar_version", "build_time"}) {
        auto it = database_metadata.find(k);
        if (it == database_metadata.end()) throw HostError(std::string("database_metadata lacks ") + k);
        md.set(k, it->second);
    }
    Json genes = Json::object();
    for (const auto &kv : gene_details) genes.set(kv.first, kv.second);  // BTreeMap: sorted keys
    Json doc = Json::object();
    doc.set("pbstarphase_version", pbstarphase_version).set("database_metadata", md).set("gene_details", genes);
    return doc.pretty();
}

}  // namespace starphase
