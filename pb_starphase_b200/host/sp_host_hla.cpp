// sp_host_hla.cpp -- HLA callers above the GPU scoring path.
#include <algorithm>
#include <numeric>

#include "starphase_host.hpp"

namespace starphase {

long dp_score(const std::vector<std::pair<uint32_t, uint8_t>> &cigar, long match_score) {
    long s = 0;
    for (const auto &op : cigar) {
        const long l = static_cast<long>(op.first);
        switch (op.second) {
            case 7: s += match_score * l; break;
            case 8: s -= 4 * l; break;
            case 1: case 2: s -= std::min(6 + 2 * l, 26 + l); break;
            default: throw HostError("Unexpected cigar type: " + std::to_string(op.second));
        }
    }
    return s;
}

// sp_align_rec -> the Mapping of (query = pattern, target = text); nullopt when minimap2 would not have reported one
static std::optional<Mapping> mapping_from_alignment(const Alignment &a, size_t pattern_len, size_t text_len, int min_dp_score) {
    if (a.cigar.empty() || dp_score(a.cigar) < min_dp_score) return std::nullopt;
    Mapping m;
    m.query_start = static_cast<size_t>(a.p_start); m.query_end = static_cast<size_t>(a.p_end); m.query_len = pattern_len;
    m.target_start = static_cast<size_t>(a.t_start); m.target_end = static_cast<size_t>(a.t_end); m.target_len = text_len;
    m.nm = static_cast<size_t>(a.nm);
    m.forward = true;
    m.cigar = a.cigar;
    return m;
}

static bool is_allowed_allele_def(const HlaAlleleDefinition &def, const std::string &gene_name, const DiplotypeSettings &s) {
    return def.gene_name == gene_name && (def.dna_sequence.has_value() || !s.hla_require_dna);  // src/hla/caller.rs:1090-1095
}

static std::string join_star(const std::vector<std::string> &fields) {
    std::string out;
    for (size_t i = 0; i < fields.size(); ++i) out += (i ? ":" : "") + fields[i];
    return out;
}

ScoreReadResult score_read(GpuAligner &gpu, const std::string &dna_target, const std::string &cdna_target, const HlaDatabase &database,
                           const std::string &gene_name, const DiplotypeSettings &settings) {
    // one batch instead of the per-allele aligner.map calls of src/hla/caller.rs:1433-1462: target 0 = cDNA, 1 = DNA
    std::vector<const HlaAlleleDefinition *> allowed;
    for (const auto &kv : database)
        if (is_allowed_allele_def(kv.second, gene_name, settings)) allowed.push_back(&kv.second);
    SeqList targets = {cdna_target, dna_target}, patterns;
    std::vector<std::pair<int32_t, int32_t>> pairs;
    std::vector<int> cdna_pair(allowed.size(), -1), dna_pair(allowed.size(), -1);
    for (size_t a = 0; a < allowed.size(); ++a) {
        if (!settings.disable_cdna_scoring) {
            cdna_pair[a] = static_cast<int>(pairs.size());
            pairs.emplace_back(0, static_cast<int32_t>(patterns.size()));
            patterns.push_back(allowed[a]->cdna_sequence);
        }
        if (allowed[a]->dna_sequence) {
            dna_pair[a] = static_cast<int>(pairs.size());
            pairs.emplace_back(1, static_cast<int32_t>(patterns.size()));
            patterns.push_back(*allowed[a]->dna_sequence);
        }
    }
    const std::vector<Alignment> alns = gpu.align_pairs(targets, patterns, pairs);

    ScoreReadResult ret;
    HlaProcessedMatch best_match = HlaProcessedMatch::worst_match(2);
    for (size_t a = 0; a < allowed.size(); ++a) {  // BTreeMap order: the scan order is part of the semantics
        HlaProcessedMatch current(allowed[a]->hla_id);
        for (int which = 0; which < 2; ++which) {
            const int q = which == 0 ? cdna_pair[a] : dna_pair[a];
            std::optional<Mapping> best_mapping;
            if (q >= 0) {
                const size_t plen = patterns[static_cast<size_t>(pairs[static_cast<size_t>(q)].second)].size();
                std::vector<Mapping> mappings;
                if (auto m = mapping_from_alignment(alns[static_cast<size_t>(q)], plen, targets[static_cast<size_t>(which)].size(), settings.min_dp_score))
                    mappings.push_back(*m);
                const auto sel = select_best_mapping(mappings, false, true, std::nullopt);  // :1448-1452
                if (sel.first) best_mapping = mappings[*sel.first];
            }
            current.add_mapping(best_mapping);
        }
        HlaMappingStats stats;
        stats.cdna_stats = current.full_mapping_stats()[0];
        stats.dna_stats = current.full_mapping_stats()[1];
        if (current.is_better_match(best_match)) best_match = current;
        ret.stats.emplace(allowed[a]->hla_id, stats);
    }
    ret.best_hla_id = best_match.haplotype();
    if (!ret.best_hla_id.empty()) ret.best_star_allele = join_star(database.at(ret.best_hla_id).star_allele);
    return ret;
}

Json PgxMappingDetails::to_json() const {
    Json j = Json::object();
    j.set("read_qname", read_qname).set("best_hla_id", best_hla_id).set("best_star_allele", best_star_allele);
    j.set("best_mapping_stats", best_mapping_stats.to_json()).set("is_ignored", is_ignored);
    return j;
}

Json Diplotype::to_json() const {
    Json j = Json::object();
    j.set("hap1", hap1).set("hap2", hap2).set("diplotype", hap1 + "/" + hap2);
    return j;
}

// ------------------------------------------------------------------------------------------
// realigner
// ------------------------------------------------------------------------------------------
HlaRealigner::HlaRealigner(GpuAligner &gpu, const std::vector<std::string> &gene_list, const HlaDatabase &database)
    : gpu_(gpu), database_(database) {
    // create_hla_fasta (src/hla/realigner.rs:497-526): every allele of the listed genes that has a DNA sequence
    for (const auto &kv : database) {
        if (std::find(gene_list.begin(), gene_list.end(), kv.second.gene_name) == gene_list.end()) continue;
        if (!kv.second.dna_sequence) continue;
        alleles_.push_back(&kv.second);
        allele_seqs_.push_back(*kv.second.dna_sequence);
    }
}

// the best_n hits of one read: the alleles with the smallest distance, ties by database order
static std::vector<int32_t> top_candidates(const int32_t *row, size_t n_alleles, int n_candidates) {
    std::vector<int32_t> idx(n_alleles);
    std::iota(idx.begin(), idx.end(), 0);
    const size_t k = std::min<size_t>(static_cast<size_t>(std::max(n_candidates, 1)), n_alleles);
    std::partial_sort(idx.begin(), idx.begin() + static_cast<long>(k), idx.end(),
                      [&](int32_t a, int32_t b) { return row[a] != row[b] ? row[a] < row[b] : a < b; });
    idx.resize(k);
    return idx;
}

std::vector<PgxMappingDetails> HlaRealigner::realign_records(const std::vector<std::pair<std::string, std::string>> &reads, int n_candidates) {
    SeqList targets;
    for (const auto &r : reads) targets.push_back(r.second);
    const std::vector<int32_t> D = alleles_.empty() ? std::vector<int32_t>() : gpu_.score_batch(targets, allele_seqs_);
    return realign_records_scored(reads, D, n_candidates);
}

std::vector<PgxMappingDetails> HlaRealigner::realign_records_scored(const std::vector<std::pair<std::string, std::string>> &reads,
                                                                    const std::vector<int32_t> &D, int n_candidates) {
    SeqList targets;
    for (const auto &r : reads) targets.push_back(r.second);
    const size_t A = alleles_.size();
    if (D.size() != reads.size() * A) throw HostError("realign_records_scored: distance matrix has the wrong shape");
    std::vector<std::pair<int32_t, int32_t>> pairs;
    std::vector<size_t> first_pair(reads.size() + 1, 0);
    for (size_t r = 0; r < reads.size(); ++r) {
        first_pair[r] = pairs.size();
        if (A && !reads[r].second.empty())
            for (int32_t a : top_candidates(D.data() + r * A, A, n_candidates)) pairs.emplace_back(static_cast<int32_t>(r), a);
    }
    first_pair[reads.size()] = pairs.size();
    const std::vector<Alignment> alns = gpu_.align_pairs(targets, allele_seqs_, pairs);

    std::vector<PgxMappingDetails> out;
    for (size_t r = 0; r < reads.size(); ++r) {
        const size_t read_len = reads[r].second.size();
        MappingStats best_stats(read_len, read_len, 0);  // src/hla/realigner.rs:124
        int best_allele = -1;
        for (size_t q = first_pair[r]; q < first_pair[r + 1]; ++q) {
            const Alignment &a = alns[q];
            if (a.cigar.empty() || dp_score(a.cigar) < 200) continue;  // no hit reported
            const size_t target_len = allele_seqs_[static_cast<size_t>(pairs[q].second)].size();  // minimap2's target is the allele here
            const size_t unmapped = target_len - static_cast<size_t>(a.p_end - a.p_start);
            const MappingStats stats(target_len, static_cast<size_t>(a.nm), unmapped);
            if (stats.mapping_score() <= 0.5 && stats.custom_score(false) <= 0.03 &&
                stats.custom_score(false) < best_stats.custom_score(false)) {  // :137-146
                best_stats = stats;
                best_allele = pairs[q].second;
            }
        }
        PgxMappingDetails d;
        d.read_qname = reads[r].first;
        d.best_mapping_stats.dna_stats = best_stats;
        if (best_allele < 0) {  // :149-162
            d.best_hla_id = "REFERENCE"; d.best_star_allele = "REFERENCE"; d.is_ignored = true;
        } else {
            const HlaAlleleDefinition *def = alleles_[static_cast<size_t>(best_allele)];
            d.best_hla_id = def->hla_id;
            d.best_star_allele = def->gene_name + "*" + join_star(def->star_allele);  // :204-211
            d.is_ignored = false;
        }
        out.push_back(std::move(d));
    }
    return out;
}

// ------------------------------------------------------------------------------------------
// north_star diplotype: exhaustive scoring + allele-pair ranking
// ------------------------------------------------------------------------------------------
Json HlaGeneCall::gene_details() const {
    Json dips = Json::array();
    dips.push(diplotype.to_json());
    Json maps = Json::array();
    for (const auto &m : mapping_details) maps.push(m.to_json());
    Json j = Json::object();
    j.set("diplotypes", dips).set("simple_diplotypes", Json()).set("inexact_diplotypes", Json()).set("variant_details", Json());
    j.set("mapping_details", maps).set("multi_mapping_details", Json());
    return j;
}

HlaGeneCall diplotype_hla_gene(GpuAligner &gpu, const HlaDatabase &database, const std::string &gene_name, const std::vector<HlaRead> &reads,
                               const DiplotypeSettings &settings) {
    std::vector<const HlaAlleleDefinition *> allowed;
    for (const auto &kv : database)
        if (is_allowed_allele_def(kv.second, gene_name, settings)) {
            if (!kv.second.dna_sequence) throw HostError("diplotype_hla_gene: allele " + kv.first + " has no DNA sequence (pair ranking needs hla_require_dna)");
            allowed.push_back(&kv.second);
        }
    HlaGeneCall call;
    if (reads.empty() || allowed.empty()) {  // sentinels of src/hla/caller.rs:32-37
        call.hla_id1 = call.hla_id2 = "NO_READS";
        call.diplotype = {"NO_READS", "NO_READS"};
        return call;
    }
    SeqList dna_targets, cdna_targets, dna, cdna;
    for (const auto &r : reads) { dna_targets.push_back(r.dna_target); cdna_targets.push_back(r.cdna_target); }
    for (const auto *a : allowed) { dna.push_back(*a->dna_sequence); cdna.push_back(a->cdna_sequence); }
    const int64_t R = static_cast<int64_t>(reads.size()), A = static_cast<int64_t>(allowed.size());
    const std::vector<int32_t> Dd = gpu.score_batch(dna_targets, dna);
    std::vector<sp_pair_rec> top;
    if (settings.disable_cdna_scoring) {
        top = gpu.pair_minsum_topk(Dd, nullptr, R, A, 10);
    } else {
        const std::vector<int32_t> Dc = gpu.score_batch(cdna_targets, cdna);
        top = gpu.pair_minsum_topk(Dc, &Dd, R, A, 10);  // (cDNA, DNA) lexicographic: src/hla/mapping.rs:111-117
    }
    if (top.empty()) throw HostError("diplotype_hla_gene: pair ranking returned nothing");
    const sp_pair_rec &b = top[0];
    call.pair_score_cdna = settings.disable_cdna_scoring ? 0 : b.score;
    call.pair_score_dna = settings.disable_cdna_scoring ? b.score : b.score2;
    call.counts1 = b.c1;
    call.counts2 = static_cast<size_t>(R) - b.c1;
    const std::string id1 = allowed[b.i]->hla_id, id2 = allowed[b.j]->hla_id;
    if (b.i == b.j) {  // one allele explains every read best: homozygous (src/hla/caller.rs:905-912)
        call.hla_id1 = call.hla_id2 = id1;
    } else if (is_passing_dual(call.counts1, call.counts2, settings.min_consensus_fraction, settings.min_cdf, settings.expected_maf)) {
        call.hla_id1 = id1; call.hla_id2 = id2;
    } else if (call.counts1 > call.counts2) {  // :893-899
        call.hla_id1 = call.hla_id2 = id1;
    } else {
        call.hla_id1 = call.hla_id2 = id2;
    }
    auto star = [&](const std::string &id) { return "*" + join_star(database.at(id).star_allele); };  // :1046-1065
    call.diplotype = {star(call.hla_id1), star(call.hla_id2)};

    // per-read database assignment, as HlaRealigner::realign_record reports it (restricted to this gene)
    HlaDatabase gene_db;
    for (const auto *a : allowed) gene_db.emplace(a->hla_id, *a);
    HlaRealigner realigner(gpu, {gene_name}, gene_db);
    std::vector<std::pair<std::string, std::string>> qs;
    for (const auto &r : reads) qs.emplace_back(r.qname, r.dna_target);
    call.mapping_details = realigner.realign_records_scored(qs, Dd);  // same allele order: reuse the DNA distances
    return call;
}

}  // namespace starphase
