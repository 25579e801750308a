// sp_host_cyp2d6.cpp -- CYP2D6 labels, weights, chains and the chain-pair search above the GPU path.
#include <algorithm>
#include <cmath>
#include <limits>
#include <numeric>
#include <tuple>

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include "starphase_host.hpp"

namespace starphase {

using RT = Cyp2d6RegionType;

// ------------------------------------------------------------------------------------------
// labels -- src/cyp2d6/region_label.rs
// ------------------------------------------------------------------------------------------
const char *region_type_name(RT t) {  // Display of Cyp2d6RegionType (strum serialize names)
    switch (t) {
        case RT::Unknown: return "UNKNOWN";
        case RT::Rep6: return "REP6";
        case RT::Cyp2d6: return "CYP2D6";
        case RT::LinkRegion: return "link_region";
        case RT::Rep7: return "REP7";
        case RT::Spacer: return "spacer";
        case RT::Cyp2d7: return "CYP2D7";
        case RT::Cyp2d6Deletion: return "CYP2D6*5";
        case RT::Hybrid: return "Hybrid";
        case RT::FalseAllele: return "FalseAllele";
    }
    return "UNKNOWN";
}

RT region_type_from_name(const std::string &s) {
    for (RT t : {RT::Unknown, RT::Rep6, RT::Cyp2d6, RT::LinkRegion, RT::Rep7, RT::Spacer, RT::Cyp2d7, RT::Cyp2d6Deletion, RT::Hybrid,
                 RT::FalseAllele})
        if (s == region_type_name(t)) return t;
    throw HostError("unknown CYP2D6 region type: " + s);
}

bool Cyp2d6RegionLabel::is_cyp2d() const {  // :39-55
    return region_type == RT::Cyp2d6 || region_type == RT::Cyp2d7 || region_type == RT::Cyp2d6Deletion || region_type == RT::Hybrid;
}
bool Cyp2d6RegionLabel::is_rep() const { return region_type == RT::Rep6 || region_type == RT::Rep7; }  // :58-60
bool Cyp2d6RegionLabel::is_reported_allele() const {                                                    // :63-69
    return region_type == RT::Cyp2d6 || region_type == RT::Cyp2d6Deletion || region_type == RT::Hybrid;
}

std::string Cyp2d6RegionLabel::full_allele() const {  // :139-168
    const std::string t = region_type_name(region_type);
    switch (region_type) {
        case RT::Cyp2d6: return subtype_label ? t + "*" + *subtype_label : t;
        case RT::Hybrid: return subtype_label ? *subtype_label : t;
        case RT::FalseAllele: return subtype_label ? t + "_" + *subtype_label : t;
        default: return t;
    }
}

// Rust: `if let Ok(float_value) = subtype.parse::<f64>() { format!("*{}", float_value.floor() as i64) }`
static bool parse_rust_f64(const std::string &s, double &out) {
    if (s.empty()) return false;
    size_t i = 0;
    if (s[i] == '+' || s[i] == '-') ++i;
    std::string rest = s.substr(i);
    std::string low;
    for (char c : rest) low.push_back(static_cast<char>(std::tolower(static_cast<unsigned char>(c))));
    const double sign = s[0] == '-' ? -1.0 : 1.0;
    if (low == "inf" || low == "infinity") { out = sign * std::numeric_limits<double>::infinity(); return true; }
    if (low == "nan") { out = std::numeric_limits<double>::quiet_NaN(); return true; }
    size_t k = 0, digits = 0;
    while (k < rest.size() && std::isdigit(static_cast<unsigned char>(rest[k]))) { ++k; ++digits; }
    if (k < rest.size() && rest[k] == '.') {
        ++k;
        while (k < rest.size() && std::isdigit(static_cast<unsigned char>(rest[k]))) { ++k; ++digits; }
    }
    if (digits == 0) return false;
    if (k < rest.size() && (rest[k] == 'e' || rest[k] == 'E')) {
        ++k;
        if (k < rest.size() && (rest[k] == '+' || rest[k] == '-')) ++k;
        size_t ed = 0;
        while (k < rest.size() && std::isdigit(static_cast<unsigned char>(rest[k]))) { ++k; ++ed; }
        if (ed == 0) return false;
    }
    if (k != rest.size()) return false;
    out = std::strtod(s.c_str(), nullptr);
    return true;
}

std::string Cyp2d6RegionLabel::simplify_allele(bool detailed, const std::map<std::string, std::string> &cyp_translate) const {  // :101-136
    if (region_type == RT::Cyp2d6 || region_type == RT::Hybrid) {
        if (!subtype_label) return full_allele();
        const std::string &s = *subtype_label;
        auto it = cyp_translate.find(s);
        if (it != cyp_translate.end()) return "*" + it->second;
        if (detailed) return "*" + s;
        double v;
        if (parse_rust_f64(s, v)) {  // `as i64` saturates, NaN -> 0
            long long iv;
            if (std::isnan(v)) iv = 0;
            else if (v >= 9223372036854775807.0) iv = std::numeric_limits<long long>::max();
            else if (v <= -9223372036854775808.0) iv = std::numeric_limits<long long>::min();
            else iv = static_cast<long long>(std::floor(v));
            return "*" + std::to_string(iv);
        }
        return "*" + s;
    }
    if (region_type == RT::Cyp2d6Deletion) return "*5";
    return full_allele();
}

bool Cyp2d6RegionLabel::is_allowed_label() const { return region_type != RT::Unknown && region_type != RT::FalseAllele; }  // :171-173

bool Cyp2d6RegionLabel::is_allowed_label_pair(const Cyp2d6RegionLabel &nxt) const {  // :178-222
    const RT t1 = region_type, t2 = nxt.region_type;
    const bool c1 = is_cyp2d(), c2 = nxt.is_cyp2d();
    const bool double_star5 = t1 == RT::Cyp2d6Deletion && t2 == RT::Cyp2d6Deletion;
    const bool unexpected = t2 == RT::Rep6 || (c1 && t1 != RT::Cyp2d6Deletion && t2 != RT::LinkRegion) || (t2 == RT::LinkRegion && !c1) ||
                            (t1 == RT::LinkRegion && !nxt.is_rep()) || (nxt.is_rep() && t1 != RT::LinkRegion) ||
                            (is_rep() && !(t2 == RT::Spacer || c2)) || (t2 == RT::Spacer && !(is_rep() || t1 == RT::Cyp2d6Deletion)) ||
                            (t1 == RT::Spacer && !c2) || (t2 == RT::Cyp2d7 && t1 != RT::Spacer) || t1 == RT::Cyp2d7;
    return !double_star5 && !unexpected;
}

bool Cyp2d6RegionLabel::is_normalizing_allele(bool normalize_all) const {  // :254-262
    return normalize_all ? is_cyp2d() : region_type == RT::Cyp2d6;
}

bool Cyp2d6RegionLabel::is_candidate_chain_head(bool normalize_all) const {  // :227-246
    if (region_type == RT::Rep6 || region_type == RT::Cyp2d6Deletion) return true;
    if (region_type == RT::Cyp2d6 || region_type == RT::Hybrid) return is_normalizing_allele(normalize_all);
    return false;
}

std::string Cyp2d6Region::index_label() const {
    return (unique_id ? std::to_string(*unique_id) : std::string("X")) + "_" + label.full_allele();
}

Cyp2d6Config Cyp2d6Config::default_config() {  // src/cyp2d6/definitions.rs:238-296
    Cyp2d6Config c;
    for (const char *part : {"intron1", "exon2", "intron2", "exon3", "intron3", "exon4", "intron4", "exon5", "intron5", "exon6", "intron6",
                             "exon7", "intron7", "exon8", "intron8", "exon9"})
        c.cyp_translate[std::string("CYP2D7::CYP2D6::") + part] = "13";
    c.cyp_translate["CYP2D6::CYP2D7::intron1"] = "68";
    c.cyp_translate["CYP2D6::CYP2D7::exon2"] = "68";
    c.cyp_translate["CYP2D6::CYP2D7::exon8"] = "61";
    c.cyp_translate["CYP2D6::CYP2D7::intron8"] = "63";
    for (const char *d : {"1", "2", "3", "4", "6", "9", "10", "17", "28", "29", "35", "41", "43", "45", "146"})
        c.inferred_connections.insert({std::string("*") + d, std::string("*") + d});
    c.inferred_connections.insert({"*4", "*68"});
    c.inferred_connections.insert({"*10", "*36"});
    c.unexpected_singletons = {"*36", "*68"};
    return c;
}

std::string convert_chain_to_hap(const std::vector<size_t> &chain, const std::vector<Cyp2d6Region> &hap_regions, Cyp2d6DetailLevel level,
                                 const std::map<std::string, std::string> &cyp_translate) {
    size_t num_non_deletion = 0;
    std::vector<size_t> reportable;
    for (auto it = chain.rbegin(); it != chain.rend(); ++it) {
        const Cyp2d6RegionLabel &lab = hap_regions[*it].label;
        const bool keep = lab.is_cyp2d() && lab.region_type != RT::Cyp2d7;
        if (keep && lab.region_type != RT::Cyp2d6Deletion) ++num_non_deletion;
        if (keep) reportable.push_back(*it);
    }
    std::vector<std::string> names;
    for (size_t c : reportable) {
        const Cyp2d6RegionLabel &lab = hap_regions[c].label;
        if (lab.region_type == RT::Cyp2d6Deletion && num_non_deletion > 0) continue;
        switch (level) {
            case Cyp2d6DetailLevel::CoreAlleles: names.push_back(lab.simplify_allele(false, cyp_translate)); break;
            case Cyp2d6DetailLevel::SubAlleles: names.push_back(lab.simplify_allele(true, cyp_translate)); break;
            case Cyp2d6DetailLevel::DeepAlleles: names.push_back("(" + hap_regions[c].deep_label() + ")"); break;
        }
    }
    std::string out;
    for (size_t i = 0; i < names.size();) {
        size_t j = i;
        while (j < names.size() && names[j] == names[i]) ++j;
        if (!out.empty()) out += " + ";
        out += j - i > 1 ? names[i] + "x" + std::to_string(j - i) : names[i];
        i = j;
    }
    return out;
}

// ------------------------------------------------------------------------------------------
// template search -- src/cyp2d6/haplotyper.rs:142-315
// ------------------------------------------------------------------------------------------
Cyp2d6Extractor::Cyp2d6Extractor(GpuAligner &gpu, std::vector<std::pair<Cyp2d6RegionLabel, std::string>> hybrid_sequences)
    : gpu_(gpu), templates_(std::move(hybrid_sequences)) {
    std::stable_sort(templates_.begin(), templates_.end(),
                     [](const auto &a, const auto &b) { return a.first.full_allele() < b.first.full_allele(); });  // key_order, :175-183
}

double overlap_score(size_t s1, size_t e1, size_t s2, size_t e2) {  // :877-893
    const size_t min_end = std::min(e1, e2), max_start = std::max(s1, s2);
    if (max_start >= min_end) return 0.0;
    return static_cast<double>(min_end - max_start) / std::min(static_cast<double>(e1 - s1), static_cast<double>(e2 - s2));
}
static size_t get_allele_priority(const Cyp2d6RegionLabel &l) { return l.region_type == RT::Cyp2d6Deletion ? 1 : 0; }  // :898-903
static bool is_penalized_type(const Cyp2d6RegionLabel &l) {  // :185-191
    return l.region_type == RT::Cyp2d6Deletion || l.region_type == RT::Rep6 || l.region_type == RT::Rep7;
}

std::vector<std::vector<AlleleMapping>> Cyp2d6Extractor::find_base_type_in_sequences(const SeqList &seqs, bool penalize_unmapped,
                                                                                     double max_missing_frac) {
    (void)penalize_unmapped;  // only changes the reference's debug strings
    const double max_ed_frac = 0.05;  // :160
    const size_t nt = templates_.size();
    struct Hit { size_t start, end; MappingStats stats; size_t tmpl; };
    std::vector<std::vector<Hit>> uncollapsed(seqs.size());
    std::vector<std::vector<size_t>> n_hits(seqs.size(), std::vector<size_t>(nt, 0));
    struct Item { size_t seq, lo, hi, tmpl; };  // search template tmpl in seqs[seq][lo, hi)
    std::vector<Item> items;
    if (nt == 0) return std::vector<std::vector<AlleleMapping>>(seqs.size());
    if (!tmpl_patterns_) {  // the 39 templates stay on the device for the life of the extractor: packed for K1, ASCII for K4
        SeqList ts;
        for (const auto &t : templates_) ts.push_back(t.second);
        tmpl_patterns_ = gpu_.prepare_patterns(ts);
    }
    const SeqList &tmpl_seqs = tmpl_patterns_->sequences();
    const std::shared_ptr<ResidentSeqs> resident_seqs = gpu_.upload(seqs);  // the reads: K1 texts of round 0, K4 texts of every round

    // every round: K1 scores each open (segment, template) item (segment x all templates in one launch; round 0 = the whole
    // sequences); only placements that could pass go to the K4 traceback, on the placement window
    for (size_t s = 0; s < seqs.size(); ++s)
        if (!seqs[s].empty())  // :148-151
            for (size_t t = 0; t < nt; ++t) items.push_back({s, 0, seqs[s].size(), t});
    const bool timing = getenv("SP_TIMING") != nullptr;  // diagnostic only: wall-clock phases to stderr
    auto t_prev = std::chrono::steady_clock::now();
    auto mark = [&](const char *what, int round, size_t n) {
        if (!timing) return;
        const auto t1 = std::chrono::steady_clock::now();
        fprintf(stderr, "[sp_timing] template search round %d %-18s %8.2f ms (%zu)\n", round, what,
                std::chrono::duration<double, std::milli>(t1 - t_prev).count(), n);
        t_prev = t1;
    };
    for (int round = 0; round < 5 && !items.empty(); ++round) {
        std::map<std::tuple<size_t, size_t, size_t>, size_t> seg_index;  // (seq, lo, hi) -> row of the K1 matrix
        SeqList seg_texts;
        for (const Item &it : items) {
            const auto key = std::make_tuple(it.seq, it.lo, it.hi);
            if (seg_index.emplace(key, seg_texts.size()).second) seg_texts.push_back(seqs[it.seq].substr(it.lo, it.hi - it.lo));
        }
        mark("segments", round, seg_texts.size());
        std::vector<int32_t> D, E;
        {
            // round 0 scores the whole sequences, which are resident already; later rounds upload their sub-segments
            const bool whole = round == 0 && seg_texts.size() == seqs.size();
            const std::shared_ptr<ResidentSeqs> segs = whole ? resident_seqs : gpu_.upload(seg_texts);
            const std::unique_ptr<DeviceMatrix> M = gpu_.score_device(*segs, *tmpl_patterns_, true);
            gpu_.matrix_to_host(*M, D, &E);
        }
        mark("K1 score_device", round, items.size());
        std::vector<std::pair<int32_t, int32_t>> pairs, windows, bounds;  // (sequence, template), [lo + w0, lo + e) inside the sequence, [lo, hi)
        std::vector<Item> pair_item;
        for (const Item &it : items) {
            const size_t row = seg_index.at(std::make_tuple(it.seq, it.lo, it.hi));
            const size_t m = tmpl_seqs[it.tmpl].size();
            const size_t d = static_cast<size_t>(D[row * nt + it.tmpl]), e = static_cast<size_t>(E[row * nt + it.tmpl]);
            if (m == 0 || (aligner_stand_ins().template_half_prefilter && 2 * d > m)) continue;  // more than half of the template unexplained: nothing minimap2 would report
            const size_t w0 = e > m + d ? e - (m + d) : 0;
            pairs.emplace_back(static_cast<int32_t>(it.seq), static_cast<int32_t>(it.tmpl));
            windows.emplace_back(static_cast<int32_t>(it.lo + w0), static_cast<int32_t>(it.lo + e));
            bounds.emplace_back(static_cast<int32_t>(it.lo), static_cast<int32_t>(it.hi));
            pair_item.push_back(it);
        }
        mark("windows", round, pairs.size());
        const std::vector<Alignment> alns = gpu_.align_pairs(*resident_seqs, tmpl_patterns_->resident(), pairs, &windows, 1, &bounds, aligner_stand_ins().min_dp_score);  // windows of resident reads
        mark("K4 align_pairs", round, pairs.size());
        std::vector<Item> next;
        for (size_t q = 0; q < pairs.size(); ++q) {
            const Alignment &a = alns[q];
            const Item &it = pair_item[q];
            const size_t m = tmpl_seqs[it.tmpl].size();
            if (a.cigar.empty() || a.score < aligner_stand_ins().min_dp_score) continue;  // no mapping reported
            const size_t clipped_start = static_cast<size_t>(a.p_start), clipped_end = m - static_cast<size_t>(a.p_end);
            MappingStats st(m, static_cast<size_t>(a.nm), m - static_cast<size_t>(a.p_end - a.p_start));
            st.clipped_start = clipped_start; st.clipped_end = clipped_end;  // new_with_clippings, :214-217
            if (st.custom_score(is_penalized_type(templates_[it.tmpl].first)) > max_ed_frac) continue;  // :221-226
            const size_t hs = static_cast<size_t>(a.t_base + a.t_start), he = static_cast<size_t>(a.t_base + a.t_end);  // t_base: K4's window or K9's
            uncollapsed[it.seq].push_back({hs, he, st, it.tmpl});
            if (++n_hits[it.seq][it.tmpl] >= 5) continue;  // best_n = 5 (src/util/mapping.rs:8-14)
            // the same template may occur again left or right of this hit (duplications)
            const size_t min_len = static_cast<size_t>(0.9 * (1.0 - std::min(max_missing_frac, 1.0)) * static_cast<double>(m));
            if (hs > it.lo && hs - it.lo >= std::max<size_t>(min_len, 200)) next.push_back({it.seq, it.lo, hs, it.tmpl});
            if (it.hi > he && it.hi - he >= std::max<size_t>(min_len, 200)) next.push_back({it.seq, he, it.hi, it.tmpl});
        }
        items = std::move(next);
        mark("accept", round, items.size());
    }

    std::vector<std::vector<AlleleMapping>> out(seqs.size());
    for (size_t s = 0; s < seqs.size(); ++s) {
        std::vector<Hit> &u = uncollapsed[s];
        // the reference pushes template by template in key order, mappings in the aligner's order (best first)
        std::stable_sort(u.begin(), u.end(), [](const Hit &a, const Hit &b) {
            if (a.tmpl != b.tmpl) return a.tmpl < b.tmpl;
            return a.stats.custom_score(true) < b.stats.custom_score(true);
        });
        std::stable_sort(u.begin(), u.end(), [](const Hit &a, const Hit &b) {  // :245-247
            return std::make_pair(a.start, a.end) < std::make_pair(b.start, b.end);
        });
        std::vector<Hit> region_mappings;
        std::optional<Hit> cur;
        for (const Hit &h : u) {  // :252-290
            if (!cur) { cur = h; continue; }
            if (overlap_score(h.start, h.end, cur->start, cur->end) > 0.9) {
                const auto &hl = templates_[h.tmpl].first, &cl = templates_[cur->tmpl].first;
                const bool penalized = is_penalized_type(hl) || is_penalized_type(cl);
                const size_t hp = get_allele_priority(hl), cp = get_allele_priority(cl);
                if ((h.stats.custom_score(penalized) < cur->stats.custom_score(penalized) && hp >= cp) || hp > cp) cur = h;
            } else {
                region_mappings.push_back(*cur);
                cur = h;
            }
        }
        if (cur) region_mappings.push_back(*cur);
        for (const Hit &h : region_mappings)  // :297-312
            if (!(h.stats.custom_score(true) > max_missing_frac))
                out[s].push_back({templates_[h.tmpl].first, h.start, h.end, h.stats});
    }
    return out;
}

// src/cyp2d6/haplotyper.rs:326-361 + :371-468, batched
std::vector<std::optional<Cyp2d6Region>> Cyp2d6Extractor::find_full_type_in_sequences(const SeqList &seqs, double max_missing_frac,
                                                                                      bool force_assignment, const Cyp2d6TypingDb &db,
                                                                                      size_t graph_band) {
    const bool penalize_unmapped = true;  // :332
    const std::vector<std::vector<AlleleMapping>> matches = find_base_type_in_sequences(seqs, penalize_unmapped, max_missing_frac);
    std::vector<std::optional<Cyp2d6Region>> out(seqs.size());
    std::vector<size_t> deep;  // sequences that go through assign_haplotype
    for (size_t s = 0; s < seqs.size(); ++s) {
        if (matches[s].empty()) continue;  // :340-342 "no matches found"
        const AlleleMapping *best = &matches[s][0];  // :345-350: min_by keeps the first of equal minima
        for (const AlleleMapping &m : matches[s])
            if (m.mapping_stats.custom_score(penalize_unmapped) < best->mapping_stats.custom_score(penalize_unmapped)) best = &m;
        bool is_mapped = false;
        for (const Cyp2d6RegionLabel &l : db.mapped_hybrids)
            is_mapped = is_mapped || (l.region_type == best->allele_label.region_type && l.subtype_label == best->allele_label.subtype_label);
        if (is_mapped) {
            deep.push_back(s);
        } else {
            Cyp2d6Region r;
            r.label = best->allele_label;
            out[s] = r;
        }
    }
    if (deep.empty()) return out;
    // :383-420: the consensus (query) on the backbone (target); one co-linear mapping per consensus here, where the reference
    // takes the longest of minimap2's mappings
    SeqList deep_seqs;
    std::vector<std::pair<int32_t, int32_t>> pairs;
    for (size_t k = 0; k < deep.size(); ++k) {
        deep_seqs.push_back(seqs[deep[k]]);
        pairs.emplace_back(0, static_cast<int32_t>(k));
    }
    const std::vector<Alignment> alns = gpu_.align_pairs(SeqList{db.backbone}, deep_seqs, pairs, nullptr, 1);
    std::vector<VariantGraph> graphs(deep.size());
    std::vector<const VariantGraph *> graph_ptrs;
    SeqList sub_sequences;
    for (size_t k = 0; k < deep.size(); ++k) {
        const Alignment &a = alns[k];
        if (a.cigar.empty() || a.score < aligner_stand_ins().min_dp_score)  // :398 assert!(!mappings.is_empty())
            throw HostError("find_full_type_in_sequences: a consensus does not map onto the CYP2D6 backbone");
        const size_t ts = static_cast<size_t>(a.t_start), te = static_cast<size_t>(a.t_end);
        graphs[k] = VariantGraph::from_reference_variants(db.backbone.substr(ts, te - ts), db.backbone_start + ts, db.variants);  // :432-441
        graph_ptrs.push_back(&graphs[k]);
        sub_sequences.push_back(deep_seqs[k].substr(static_cast<size_t>(a.p_start), static_cast<size_t>(a.p_end - a.p_start)));  // :422-425
    }
    const std::vector<GraphAlignment> walks = graph_edit_distance(gpu_, graph_ptrs, sub_sequences, graph_band);  // :445
    std::vector<std::vector<uint8_t>> vectors;
    for (size_t k = 0; k < deep.size(); ++k) {
        if (!walks[k].found) throw HostError("find_full_type_in_sequences: no graph alignment inside the band");
        vectors.push_back(graph_alleles(graphs[k], walks[k], db.variants.size()));  // :452-468
    }
    const std::vector<HaplotypeAssignment> assigned = assign_haplotypes_from_alleles(gpu_, vectors, db.haplotype_lookup, db.metadata, force_assignment);
    for (size_t k = 0; k < deep.size(); ++k) {
        Cyp2d6Region r;
        r.label = assigned[k].label;
        r.variants = assigned[k].variants;
        out[deep[k]] = r;
    }
    return out;
}

// ------------------------------------------------------------------------------------------
// allele-vector typing (src/cyp2d6/haplotyper.rs:452-601)
// ------------------------------------------------------------------------------------------
const char *variant_state_name(VariantAlleleRelationship v) {
    switch (v) {
        case VariantAlleleRelationship::Unknown: return "Unknown";
        case VariantAlleleRelationship::Match: return "Match";
        case VariantAlleleRelationship::Unexpected: return "Unexpected";
        case VariantAlleleRelationship::Missing: return "Missing";
        case VariantAlleleRelationship::AmbiguousUnexpected: return "AmbiguousUnexpected";
        case VariantAlleleRelationship::AmbiguousMissing: return "AmbiguousMissing";
        case VariantAlleleRelationship::UnknownUnexpected: return "UnknownUnexpected";
        case VariantAlleleRelationship::UnknownMissing: return "UnknownMissing";
    }
    return "Unknown";
}

std::string RegionVariant::to_string() const {  // Display, src/data_types/region_variants.rs:56-74
    const char *pre = variant_state == VariantAlleleRelationship::Match ? "=" : variant_state == VariantAlleleRelationship::Unexpected ? "+"
                    : variant_state == VariantAlleleRelationship::Missing ? "-" : "?";
    return pre + label;
}

std::string Cyp2d6Region::deep_label() const {  // src/cyp2d6/region.rs:60-95
    std::string out = index_label();
    if (variants)
        for (const RegionVariant &v : *variants) {
            switch (v.variant_state) {
                case VariantAlleleRelationship::Match: case VariantAlleleRelationship::UnknownUnexpected: break;
                case VariantAlleleRelationship::Unexpected: out += " +" + v.label; break;
                case VariantAlleleRelationship::Missing: out += " -" + v.label; break;
                default: out += " ?" + v.label; break;
            }
        }
    return out;
}

std::string cyp2d6_alleles_json(const std::vector<std::vector<size_t>> &best, const std::vector<Cyp2d6Region> &hap_regions,
                                const std::map<std::string, std::string> &cyp_translate) {
    if (best.size() != 2) throw HostError("assertion failed: best_diplotype_indices.len() == 2");
    auto hap = [&](const std::vector<size_t> &chain) {
        Json j = Json::object();
        j.set("deep_form", convert_chain_to_hap(chain, hap_regions, Cyp2d6DetailLevel::DeepAlleles, cyp_translate));
        j.set("suballele_form", convert_chain_to_hap(chain, hap_regions, Cyp2d6DetailLevel::SubAlleles, cyp_translate));
        j.set("core_form", convert_chain_to_hap(chain, hap_regions, Cyp2d6DetailLevel::CoreAlleles, cyp_translate));
        return j;
    };
    std::map<std::string, Json> alleles;  // BTreeMap<index_label, Vec<RegionVariant>>
    for (const Cyp2d6Region &r : hap_regions) {
        if (!r.variants) continue;
        Json arr = Json::array();
        for (const RegionVariant &v : *r.variants) arr.push(v.to_json());
        if (!alleles.emplace(r.index_label(), arr).second) throw HostError("assertion failed: alleles.insert(core_label, vec_variants).is_none()");
    }
    Json al = Json::object();
    for (const auto &kv : alleles) al.set(kv.first, kv.second);
    Json doc = Json::object();
    doc.set("hap1", hap(best[0])).set("hap2", hap(best[1])).set("alleles", al);
    return doc.pretty();
}

Json RegionVariant::to_json() const {
    Json j = Json::object();
    j.set("label", label).set("is_vi", is_vi).set("variant_state", variant_state_name(variant_state));
    return j;
}

std::vector<uint8_t> alleles_from_traversal(size_t num_variants, const std::vector<size_t> &traversed_nodes,
                                            const std::map<size_t, std::vector<std::pair<size_t, uint8_t>>> &node_to_alleles) {
    std::vector<uint8_t> alleles(num_variants, 3);
    for (size_t node : traversed_nodes) {
        const auto it = node_to_alleles.find(node);
        if (it == node_to_alleles.end()) continue;
        for (const auto &va : it->second) {
            if (va.first >= num_variants) throw HostError("index out of bounds: variant index " + std::to_string(va.first));
            if (alleles[va.first] == 3) alleles[va.first] = va.second;
            else if (alleles[va.first] != va.second) alleles[va.first] = 2;
        }
    }
    return alleles;
}

std::vector<HaplotypeAssignment> assign_haplotypes_from_alleles(GpuAligner &gpu, const std::vector<std::vector<uint8_t>> &alleles,
                                                                const std::map<std::string, std::vector<uint8_t>> &haplotype_lookup,
                                                                const std::vector<VariantMetadata> &variants, bool force_assignment) {
    const size_t nv = variants.size(), nh = haplotype_lookup.size();
    std::vector<std::vector<uint8_t>> haps;
    std::vector<const std::string *> stars;  // BTreeMap<Cyp2d6RegionLabel, _> order: all keys are (Cyp2d6, Some(star)) -> bytewise by star
    for (const auto &kv : haplotype_lookup) { stars.push_back(&kv.first); haps.push_back(kv.second); }
    std::vector<uint8_t> is_vi(nv);
    for (size_t v = 0; v < nv; ++v) is_vi[v] = variants[v].is_vi ? 1 : 0;
    for (const auto &a : alleles)
        if (a.size() != nv) throw HostError("assertion `left == right` failed: alleles.len() vs haplotype_vec.len()");
    std::vector<uint32_t> vi_match, all_match;
    gpu.variant_match(alleles, haps, is_vi, vi_match, all_match);  // K6

    std::vector<HaplotypeAssignment> out(alleles.size());
    for (size_t s = 0; s < alleles.size(); ++s) {
        // :471-517: start from {Unknown} at (0, 0); strictly greater replaces, equal joins the set
        std::vector<int> best_set = {-1};  // -1 = the Unknown label
        std::pair<size_t, size_t> best_score(0, 0);
        for (size_t h = 0; h < nh; ++h) {
            const std::pair<size_t, size_t> score(vi_match[s * nh + h], all_match[s * nh + h]);
            if (score > best_score) { best_set.assign(1, static_cast<int>(h)); best_score = score; }
            else if (score == best_score) best_set.push_back(static_cast<int>(h));
        }
        auto label_of = [&](int h) {
            Cyp2d6RegionLabel l;
            if (h >= 0) { l.region_type = Cyp2d6RegionType::Cyp2d6; l.subtype_label = *stars[static_cast<size_t>(h)]; }
            return l;
        };
        int best = -1;
        if (best_set.size() == 1) {
            best = best_set[0];
        } else if (force_assignment) {  // :523-534: sorted by full_allele, first candidate
            std::stable_sort(best_set.begin(), best_set.end(), [&](int a, int b) { return label_of(a).full_allele() < label_of(b).full_allele(); });
            best = best_set[0];
        }
        HaplotypeAssignment &r = out[s];
        r.label = label_of(best);
        r.vi_match = best_score.first; r.all_match = best_score.second;
        if (best < 0) continue;  // Unknown: no variant list (:545, :596-599)
        std::vector<RegionVariant> rv;
        const std::vector<uint8_t> &hv = haps[static_cast<size_t>(best)];
        for (size_t i = 0; i < nv; ++i) {
            using V = VariantAlleleRelationship;
            static const V ref_states[4] = {V::Match, V::Unexpected, V::AmbiguousUnexpected, V::UnknownUnexpected};
            static const V alt_states[4] = {V::Missing, V::Match, V::AmbiguousMissing, V::UnknownMissing};
            const V state = hv[i] == 0 ? ref_states[alleles[s][i]] : alt_states[alleles[s][i]];
            if (state == V::Match && hv[i] == 0) continue;  // REF matching REF is not reported (:581-583)
            rv.push_back({variants[i].label, variants[i].is_vi, state});
        }
        r.variants = std::move(rv);
    }
    return out;
}

// ------------------------------------------------------------------------------------------
// weights and chains
// ------------------------------------------------------------------------------------------
std::vector<SequenceWeights> weight_sequences(GpuAligner &gpu, const SeqList &segments, const SeqList &consensuses,
                                              const std::vector<Cyp2d6Region> &con_regions) {
    if (consensuses.size() != con_regions.size()) throw HostError("weight_sequences: consensuses and regions differ in length");
    std::vector<SequenceWeights> out(segments.size());
    if (segments.empty()) return out;
    // pattern = read segment (must be explained completely, chaining.rs:66), text = consensus (free ends, its clips form the overlap)
    // minimap2 reports nothing for sequences that far apart (unrelated DNA sits near 50 % unit-cost distance); the exhaustive
    // aligner always finds some placement, so pairs above 35 % count as "no mapping" and keep the default (|S|, 0.0) of :41
    const int kNoMappingPermille = aligner_stand_ins().no_mapping_permille;
    std::vector<int32_t> D, S, E;
    gpu.score_spans(consensuses, segments, D, S, E, kNoMappingPermille);
    const size_t ns = segments.size(), nc = consensuses.size();
    for (size_t s = 0; s < ns; ++s) {
        const size_t seq_len = segments[s].size();
        SequenceWeights ret(nc, {seq_len, 0.0});  // :41
        double min_ed_frac = 1.0;
        for (size_t k = 0; k < nc; ++k) {
            if (!con_regions[k].label.is_allowed_label()) continue;  // :52-55
            if (consensuses[k].empty() || seq_len == 0) continue;     // no mapping can exist
            const size_t o = k * ns + s;                              // [target = consensus][pattern = segment]
            const size_t match_score = static_cast<size_t>(D[o]);
            if (match_score >= seq_len) continue;                     // nothing aligned: the aligner reports no hit
            if (S[o] < 0 || match_score * 1000 > seq_len * static_cast<size_t>(kNoMappingPermille)) continue;  // too far apart: no hit
            const size_t con_len = consensuses[k].size();
            const size_t clipped = static_cast<size_t>(S[o]) + (con_len - static_cast<size_t>(E[o]));
            const double overlap = 1.0 - static_cast<double>(clipped) / static_cast<double>(con_len);  // :81
            if (match_score < ret[k].first || (match_score == ret[k].first && overlap > ret[k].second)) {  // :87-92
                ret[k] = {match_score, overlap};
                min_ed_frac = std::min(min_ed_frac, std::max(static_cast<double>(match_score), 0.1) / static_cast<double>(seq_len));
            }
        }
        if (min_ed_frac <= 0.05) out[s] = std::move(ret);  // :96-102
    }
    return out;
}

ChainBuild build_chains(const std::map<std::string, std::vector<SequenceWeights>> &read_weights, size_t n_haps) {
    ChainBuild b;
    b.best_allele_mapping_counts.assign(n_haps, 0);
    for (const auto &kv : read_weights) {  // BTreeMap qname order
        if (kv.second.empty()) continue;
        std::vector<std::vector<size_t>> putative(1);
        std::vector<SequenceWeights> weighted;
        for (const SequenceWeights &ws : kv.second) {
            if (ws.empty()) continue;
            size_t min_ed = std::numeric_limits<size_t>::max(), n_min = 0;
            for (const auto &w : ws) min_ed = std::min(min_ed, w.first);
            for (const auto &w : ws) n_min += w.first == min_ed;
            std::vector<std::vector<size_t>> next;
            for (const auto &pc : putative)
                for (size_t ci = 0; ci < ws.size(); ++ci)
                    if (ws[ci].first == min_ed) {
                        std::vector<size_t> e = pc;
                        e.push_back(ci);
                        next.push_back(std::move(e));
                        if (n_min == 1) ++b.best_allele_mapping_counts[ci];  // inside the per-chain loop, caller.rs:471-483
                    }
            putative = std::move(next);
            weighted.push_back(ws);
        }
        if (putative.empty() || (putative.size() == 1 && putative[0].empty())) continue;
        b.qname_chains[kv.first] = std::move(putative);
        b.qname_chain_scores[kv.first] = std::move(weighted);
    }
    for (auto &kv : b.qname_chains) {  // caller.rs:520-537
        std::vector<std::vector<size_t>> kept;
        for (const auto &c : kv.second)
            if (std::all_of(c.begin(), c.end(), [&](size_t x) { return b.best_allele_mapping_counts[x] > 0; })) kept.push_back(c);
        if (kept.empty()) throw HostError("chain collapse for read " + kv.first);
        kv.second = std::move(kept);
    }
    return b;
}

// ------------------------------------------------------------------------------------------
// chain-pair search
// ------------------------------------------------------------------------------------------
namespace {

struct ChainCtx {
    const Cyp2d6Config &cfg;
    const std::vector<Cyp2d6Region> &regions;
    std::vector<std::vector<bool>> down, inferred;
};

double rust_round(double x) {  // f64::round: half away from zero
    if (std::isnan(x) || std::isinf(x)) return x;
    return x >= 0 ? std::floor(x + 0.5) : -std::floor(-x + 0.5);
}

bool is_sub(const std::vector<size_t> &hay, const std::vector<size_t> &needle) {  // chaining.rs:782-784
    if (needle.size() > hay.size()) return false;
    for (size_t s = 0; s + needle.size() <= hay.size(); ++s)
        if (std::equal(needle.begin(), needle.end(), hay.begin() + static_cast<long>(s))) return true;
    return false;
}

size_t unexpected_count(const std::vector<size_t> &chain, const ChainCtx &cx) {  // chaining.rs:739-775
    std::vector<std::string> reduced;
    for (size_t c : chain) {
        const auto &lab = cx.regions[c].label;
        if (lab.is_cyp2d() && lab.region_type != RT::Cyp2d7) reduced.push_back(lab.simplify_allele(false, cx.cfg.cyp_translate));
    }
    size_t errors = 0;
    if (reduced.empty() || reduced[0].rfind("*", 0) != 0) ++errors;
    if (reduced.size() == 1 && cx.cfg.unexpected_singletons.count(reduced[0])) ++errors;
    for (size_t i = 0; i + 1 < reduced.size(); ++i)
        if (!cx.cfg.inferred_connections.count({reduced[i], reduced[i + 1]})) ++errors;
    return errors;
}

size_t count_inferred_edges(const std::vector<size_t> &chain, const ChainCtx &cx) {  // chaining.rs:828-840 (one chain)
    size_t n = 0;
    for (size_t i = 0; i + 1 < chain.size(); ++i) n += cx.inferred[chain[i]][chain[i + 1]];
    return n;
}

// chaining.rs:603-674: (keep extending?, may be a candidate?)
std::pair<bool, bool> check_chain_inferrences(const std::vector<size_t> &chain, const ChainCtx &cx) {
    const size_t last = chain.back();
    const bool last_is_cyp2d = cx.regions[last].label.is_cyp2d();
    std::optional<size_t> opt_index;
    for (size_t ci = chain.size() - 1; ci-- > 0;)
        if (cx.regions[chain[ci]].label.is_cyp2d()) { opt_index = ci; break; }
    const size_t start = opt_index.value_or(0);
    bool detected = false;
    for (size_t i = start; i + 1 < chain.size(); ++i) detected = detected || cx.inferred[chain[i]][chain[i + 1]];
    if (!detected) return {true, true};
    if (!last_is_cyp2d) return {true, false};
    if (!opt_index) return {true, true};
    const size_t prev = chain[*opt_index];
    const auto &h1 = cx.regions[prev].label, &h2 = cx.regions[last].label;
    const std::string h1m = h1.simplify_allele(false, cx.cfg.cyp_translate), h2m = h2.simplify_allele(false, cx.cfg.cyp_translate);
    const bool connected = prev != last && cx.cfg.inferred_connections.count({h1m, h2m}) > 0;
    const bool d7_tail = h2.region_type == RT::Cyp2d7 && h1.region_type != RT::Cyp2d7 && h1.is_cyp2d();
    const bool allowed = connected || d7_tail;
    return {allowed, allowed};
}

}  // namespace

ChainPairResult find_best_chain_pair(GpuAligner &gpu, const Cyp2d6Config &cfg,
                                     const std::map<std::string, std::vector<std::vector<size_t>>> &obs_chains,
                                     const std::map<std::string, std::vector<SequenceWeights>> &chain_scores,
                                     const std::vector<Cyp2d6Region> &hap_regions, bool infer, bool normalize_all, const ChainPenalties &pen,
                                     bool ignore_limits) {
    if (pen.lasso_penalty < 0.0) throw HostError("Lasso penalty must be >= 0.0");
    const size_t n = hap_regions.size();
    ChainCtx cx{cfg, hap_regions, std::vector<std::vector<bool>>(n, std::vector<bool>(n, false)),
                std::vector<std::vector<bool>>(n, std::vector<bool>(n, false))};
    auto lab = [&](size_t i) -> const Cyp2d6RegionLabel & { return hap_regions[i].label; };
    // (i) observed edges, chaining.rs:243-266
    for (const auto &kv : obs_chains)
        for (const auto &chain : kv.second)
            for (size_t i = 0; i + 1 < chain.size(); ++i) {
                const size_t a = chain[i], b = chain[i + 1];
                if (lab(a).is_allowed_label() && lab(b).is_allowed_label() && (ignore_limits || lab(a).is_allowed_label_pair(lab(b))))
                    cx.down[a][b] = true;
            }
    // (ii) inferred edges, :269-305
    if (infer)
        for (size_t i = 0; i < n; ++i) {
            const bool down_no_link = std::none_of(cx.down[i].begin(), cx.down[i].end(), [](bool b) { return b; });
            for (size_t j = 0; j < n; ++j) {
                bool up_no_link = true;
                for (size_t r = 0; r < n; ++r) up_no_link = up_no_link && !cx.down[r][j];
                if ((down_no_link || up_no_link) && !cx.down[i][j] && lab(i).is_allowed_label() && lab(j).is_allowed_label() &&
                    lab(i).is_allowed_label_pair(lab(j)))
                    cx.inferred[i][j] = true;
            }
        }
    // (iii) heads, (iv) enumeration: stack, pop from the back, :308-396
    std::vector<std::vector<size_t>> remaining, possible;
    for (size_t i = 0; i < n; ++i)
        if (ignore_limits || lab(i).is_candidate_chain_head(normalize_all)) remaining.push_back({i});
    if (remaining.empty()) throw NoChainingHead();
    while (!remaining.empty()) {
        std::vector<size_t> cur = std::move(remaining.back());
        remaining.pop_back();
        const auto ok = check_chain_inferrences(cur, cx);
        if (!ok.first) continue;
        const std::string simplified = convert_chain_to_hap(cur, hap_regions, Cyp2d6DetailLevel::SubAlleles, cfg.cyp_translate);
        if (ignore_limits || (!simplified.empty() && ok.second)) possible.push_back(cur);
        const size_t last = cur.back();
        auto extend = [&](const std::vector<std::vector<bool>> &edges) {
            for (size_t ext = 0; ext < n; ++ext)
                if (edges[last][ext] && std::count(cur.begin(), cur.end(), ext) < 3) {
                    std::vector<size_t> e = cur;
                    e.push_back(ext);
                    remaining.push_back(std::move(e));
                }
        };
        extend(cx.down);
        if (infer) extend(cx.inferred);
    }
    if (possible.empty()) throw NoChainsFound();
    const size_t P = possible.size();

    // integer ED of every pair from the GPU: S[i][j] = sum_r min(B[i][r], B[j][r]); ED = S - sum_r optimum_r (:683-731, :470-485)
    std::vector<std::vector<int32_t>> chains32(P);
    for (size_t c = 0; c < P; ++c) chains32[c].assign(possible[c].begin(), possible[c].end());
    std::vector<std::vector<std::vector<uint32_t>>> W;
    uint64_t optimum_total = 0;
    for (const auto &kv : chain_scores) {
        std::vector<std::vector<uint32_t>> rw;
        for (const SequenceWeights &seg : kv.second) {
            if (seg.size() != n) throw HostError("find_best_chain_pair: weight row has the wrong width");
            std::vector<uint32_t> row(n);
            uint32_t mn = std::numeric_limits<uint32_t>::max();
            for (size_t k = 0; k < n; ++k) { row[k] = static_cast<uint32_t>(seg[k].first); mn = std::min(mn, row[k]); }
            optimum_total += mn;
            rw.push_back(std::move(row));
        }
        W.push_back(std::move(rw));
    }
    std::vector<uint64_t> S(P * P, 0);
    if (!W.empty()) S = gpu.chain_pair_sums(chains32, W, static_cast<int64_t>(n));

    // per-chain integer terms
    std::vector<std::vector<uint32_t>> counts(P, std::vector<uint32_t>(n, 0));
    std::vector<size_t> unexp(P, 0), n_inf(P, 0);
    for (size_t c = 0; c < P; ++c) {
        for (size_t h : possible[c]) ++counts[c][h];
        unexp[c] = ignore_limits ? 0 : unexpected_count(possible[c], cx);
        n_inf[c] = infer ? count_inferred_edges(possible[c], cx) : 0;
    }
    std::vector<bool> lasso_hap(n);
    for (size_t h = 0; h < n; ++h)
        lasso_hap[h] = lab(h).is_allowed_label() && (ignore_limits || lab(h).is_normalizing_allele(normalize_all) || lab(h).is_reported_allele());

    struct Cand { double lb; uint32_t i, j; uint64_t ed; double lasso, unexpected, inferred; };
    std::vector<Cand> cands;
    cands.reserve(P * (P + 1) / 2);
    for (size_t i = 0; i < P; ++i)
        for (size_t j = i; j < P; ++j) {
            size_t extra = 0;  // count_unexpected_alleles, :794-819
            for (size_t h = 0; h < n; ++h) {
                const uint32_t hc = counts[i][h] + counts[j][h];
                if (lasso_hap[h] && hc > 0) extra += hc - 1;
            }
            Cand c;
            c.i = static_cast<uint32_t>(i); c.j = static_cast<uint32_t>(j);
            c.ed = W.empty() ? 0 : S[i * P + j] - optimum_total;
            c.lasso = pen.lasso_penalty * static_cast<double>(extra);
            c.unexpected = static_cast<double>(unexp[i] + unexp[j]) * pen.unexpected_chain_penalty;
            c.inferred = static_cast<double>(n_inf[i] + n_inf[j]) * pen.inferred_edge_penalty;
            // primary_score (:172-174) without the multinomial term, which is >= 0
            c.lb = static_cast<double>(c.ed) * pen.ln_ed_penalty + 0.0 + c.lasso + c.unexpected + c.inferred;
            cands.push_back(c);
        }
    std::sort(cands.begin(), cands.end(), [](const Cand &a, const Cand &b) {
        if (a.lb != b.lb) return a.lb < b.lb;
        if (a.i != b.i) return a.i < b.i;
        return a.j < b.j;
    });

    // exact evaluation in bound order until the bound passes the best score
    ChainPairResult res;
    res.n_possible_chains = P;
    bool have = false;
    for (const Cand &c : cands) {
        if (have && c.lb > res.score) break;
        ++res.n_full_evaluations;
        const std::vector<size_t> &ci = possible[c.i], &cj = possible[c.j];
        std::vector<double> hap_weights(n, 0.0);
        uint64_t ed = 0;
        for (const auto &kv : chain_scores) {  // containment_score per read, :683-731, BTreeMap order
            const std::vector<SequenceWeights> &cw = kv.second;
            const size_t wl = cw.size();
            uint64_t optimum = 0, worst = 0;
            for (const auto &seg : cw) {
                size_t mn = std::numeric_limits<size_t>::max(), mx = 0;
                for (const auto &w : seg) { mn = std::min(mn, w.first); mx = std::max(mx, w.first); }
                optimum += mn; worst += mx;
            }
            uint64_t best = 2 * worst;
            std::vector<std::pair<const std::vector<size_t> *, size_t>> best_windows;
            for (const std::vector<size_t> *other : {&ci, &cj}) {
                if (other->size() < wl) continue;
                for (size_t s = 0; s + wl <= other->size(); ++s) {
                    uint64_t total = 0;
                    for (size_t t = 0; t < wl; ++t) total += cw[t][(*other)[s + t]].first;
                    if (total < best) { best = total; best_windows.clear(); }
                    if (total == best) best_windows.emplace_back(other, s);
                }
            }
            const uint64_t score = best - optimum;
            ed = ed + score < ed ? std::numeric_limits<uint64_t>::max() : ed + score;  // saturating_add
            if (!best_windows.empty()) {
                const double split = 1.0 / static_cast<double>(best_windows.size());
                for (const auto &bw : best_windows)
                    for (size_t off = 0; off < wl; ++off) {
                        const size_t con = (*bw.first)[bw.second + off];
                        hap_weights[con] += split * cw[off][con].second;
                    }
            }
        }
        if (ed != c.ed) throw HostError("find_best_chain_pair: GPU edit distance disagrees with the window scan");
        // multinomial, :854-903
        std::vector<uint32_t> hc(n);
        for (size_t h = 0; h < n; ++h) hc[h] = counts[c.i][h] + counts[c.j][h];
        std::vector<double> cnt;
        std::vector<uint64_t> coverage;
        for (size_t h = 0; h < n; ++h)
            if (hc[h] > 0 && (ignore_limits || lab(h).is_normalizing_allele(normalize_all))) {
                cnt.push_back(static_cast<double>(hc[h]));
                const double r = rust_round(hap_weights[h]);
                coverage.push_back(r <= 0.0 ? 0 : static_cast<uint64_t>(r));
            }
        const double total = std::accumulate(cnt.begin(), cnt.end(), 0.0);
        std::vector<double> probs;
        for (double x : cnt) probs.push_back(x / total);
        const uint64_t cov_sum = std::accumulate(coverage.begin(), coverage.end(), uint64_t{0});
        double mn_pen;
        if (probs.empty() || cov_sum == 0) {
            auto has_del = [&](const std::vector<size_t> &ch) {
                return std::any_of(ch.begin(), ch.end(), [&](size_t h) { return lab(h).region_type == RT::Cyp2d6Deletion; });
            };
            if (!normalize_all && has_del(ci) && has_del(cj)) mn_pen = 0.0;
            else continue;  // invalid pair
        } else {
            mn_pen = std::fabs(multinomial_ln_pmf(probs, coverage));
        }
        const double ln_ed = static_cast<double>(ed) * pen.ln_ed_penalty;
        const double score = ln_ed + mn_pen + c.lasso + c.unexpected + c.inferred;  // primary_score, :172-174
        const bool better = !have || score < res.score || (score == res.score && (c.i < res.index1 || (c.i == res.index1 && c.j < res.index2)));
        if (better) {
            have = true;
            res.score = score; res.index1 = c.i; res.index2 = c.j; res.edit_distance = ed;
        }
    }
    if (!have) throw NoScorePairs();
    res.best_chains = {possible[res.index1], possible[res.index2]};
    std::sort(res.best_chains.begin(), res.best_chains.end());  // :568-573
    std::set<size_t> used;
    for (const auto &ch : res.best_chains) used.insert(ch.begin(), ch.end());
    for (size_t i = 0; i < n; ++i)
        if (!used.count(i)) res.dangling_alleles.push_back(std::to_string(i) + "_" + lab(i).full_allele());  // :576-589
    (void)is_sub;  // unmet_observations is debug output only (:427-440)
    return res;
}

// ------------------------------------------------------------------------------------------
// the chaining half of diplotype_cyp2d6
// ------------------------------------------------------------------------------------------
Json Cyp2d6Call::gene_details() const {
    Json dips = Json::array(), simple = Json::array(), inexact = Json::array();
    dips.push(diplotype.to_json());
    simple.push(simple_diplotype.to_json());
    Json deep = Json::object();
    deep.set("basic_diplotype", deep_diplotype.to_json()).set("haplotype_1", Json()).set("haplotype_2", Json());
    inexact.push(deep);
    Json j = Json::object();
    j.set("diplotypes", dips).set("simple_diplotypes", simple).set("inexact_diplotypes", inexact).set("variant_details", Json());
    j.set("mapping_details", Json()).set("multi_mapping_details", multi_mapping_details);
    return j;
}

Cyp2d6Call call_cyp2d6_chains(GpuAligner &gpu, const Cyp2d6Config &cfg, const SeqList &consensuses, std::vector<Cyp2d6Region> hap_regions,
                              const std::map<std::string, std::vector<Cyp2d6ReadRegion>> &regions_of_interest, bool infer_connections,
                              bool normalize_all_alleles) {
    // every read segment against every consensus in one batch (weight_sequence is called once per segment at caller.rs:450)
    SeqList segments;
    for (const auto &kv : regions_of_interest)
        for (const auto &r : kv.second) segments.push_back(r.sequence);
    const std::vector<SequenceWeights> ws = weight_sequences(gpu, segments, consensuses, hap_regions);
    std::map<std::string, std::vector<SequenceWeights>> read_weights;
    size_t s = 0;
    for (const auto &kv : regions_of_interest) {
        auto &v = read_weights[kv.first];
        for (size_t k = 0; k < kv.second.size(); ++k) v.push_back(ws[s++]);
    }
    ChainBuild chains = build_chains(read_weights, hap_regions.size());

    Cyp2d6Call call;
    for (const auto &kv : chains.qname_chains) {  // caller.rs:556-567: zip with ALL regions of the read
        if (kv.second.size() != 1) continue;
        const auto &regs = regions_of_interest.at(kv.first);
        for (size_t k = 0; k < std::min(kv.second[0].size(), regs.size()); ++k) {
            const size_t con = kv.second[0][k];
            Json range = Json::object();
            range.set("start", static_cast<long long>(regs[k].start)).set("end", static_cast<long long>(regs[k].end));
            Json d = Json::object();
            d.set("read_qname", kv.first).set("read_position", range).set("consensus_id", static_cast<long long>(con));
            d.set("consensus_star_allele", hap_regions[con].index_label());
            call.multi_mapping_details.push(d);
        }
    }
    for (size_t c = 0; c < hap_regions.size(); ++c)  // caller.rs:572-583
        if (chains.best_allele_mapping_counts[c] == 0 && hap_regions[c].label.region_type != RT::Unknown &&
            hap_regions[c].label.region_type != RT::FalseAllele)
            hap_regions[c].label.mark_false_allele();

    call.chain_pair = find_best_chain_pair(gpu, cfg, chains.qname_chains, chains.qname_chain_scores, hap_regions, infer_connections,
                                           normalize_all_alleles, ChainPenalties(), false);
    const auto &best = call.chain_pair.best_chains;
    auto hap = [&](size_t k, Cyp2d6DetailLevel lvl) { return convert_chain_to_hap(best[k], hap_regions, lvl, cfg.cyp_translate); };
    call.deep_diplotype = {hap(0, Cyp2d6DetailLevel::DeepAlleles), hap(1, Cyp2d6DetailLevel::DeepAlleles)};
    call.diplotype = {hap(0, Cyp2d6DetailLevel::SubAlleles), hap(1, Cyp2d6DetailLevel::SubAlleles)};
    call.simple_diplotype = {hap(0, Cyp2d6DetailLevel::CoreAlleles), hap(1, Cyp2d6DetailLevel::CoreAlleles)};
    call.hap_regions = std::move(hap_regions);
    return call;
}

}  // namespace starphase
