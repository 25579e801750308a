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
    if (a.cigar.empty() || a.score < min_dp_score) return std::nullopt;
    Mapping m;
    m.query_start = static_cast<size_t>(a.p_start); m.query_end = static_cast<size_t>(a.p_end); m.query_len = pattern_len;
    m.target_start = static_cast<size_t>(a.t_start); m.target_end = static_cast<size_t>(a.t_end); m.target_len = text_len;
    m.nm = static_cast<size_t>(a.nm);
    m.forward = true;
    m.cigar = a.cigar;
    return m;
}

bool is_allowed_allele_def(const HlaAlleleDefinition &def, const std::string &gene_name, const DiplotypeSettings &s) {
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
    const std::vector<Alignment> alns = gpu.align_pairs(targets, patterns, pairs, nullptr, 5);  // a = 5: src/hla/caller.rs:1370-1379

    ScoreReadResult ret;
    HlaProcessedMatch best_match = HlaProcessedMatch::worst_match(2);
    for (size_t a = 0; a < allowed.size(); ++a) {  // BTreeMap order: the scan order is part of the semantics
        HlaProcessedMatch current(allowed[a]->hla_id);
        std::optional<DetailedMappingStats> detail[2];  // full_mappings()[0..1] as src/hla/debug.rs records them
        for (int which = 0; which < 2; ++which) {
            const int q = which == 0 ? cdna_pair[a] : dna_pair[a];
            std::optional<Mapping> best_mapping;
            if (q >= 0) {
                const std::string &pattern = patterns[static_cast<size_t>(pairs[static_cast<size_t>(q)].second)];
                std::vector<Mapping> mappings;
                if (auto m = mapping_from_alignment(alns[static_cast<size_t>(q)], pattern.size(), targets[static_cast<size_t>(which)].size(), settings.min_dp_score))
                    mappings.push_back(*m);
                const auto sel = select_best_mapping(mappings, false, true, std::nullopt);  // :1448-1452
                if (sel.first) {
                    best_mapping = mappings[*sel.first];
                    detail[which] = detailed_mapping_stats(*best_mapping, targets[static_cast<size_t>(which)], pattern);
                }
            }
            current.add_mapping(best_mapping);
        }
        ret.read_mapping_stats.add_mapping(join_star(allowed[a]->star_allele), detail[0], detail[1]);  // :1471-1477
        HlaMappingStats stats;
        stats.cdna_stats = current.full_mapping_stats()[0];
        stats.dna_stats = current.full_mapping_stats()[1];
        if (current.is_better_match(best_match)) best_match = current;
        ret.stats.emplace(allowed[a]->hla_id, stats);
    }
    ret.best_hla_id = best_match.haplotype();
    if (!ret.best_hla_id.empty()) {
        ret.best_star_allele = join_star(database.at(ret.best_hla_id).star_allele);
        ret.read_mapping_stats.set_best_match(ret.best_hla_id, ret.best_star_allele);  // :1502-1509
    }
    return ret;
}

ScoreReadResult score_consensus(GpuAligner &gpu, const std::string &reference_sequence, int64_t ref_start, const std::string &consensus,
                                const HlaDatabase &database, const std::string &gene_name,
                                const std::vector<std::pair<uint64_t, uint64_t>> &exons, bool is_forward_strand,
                                const DiplotypeSettings &settings) {
    if (consensus.empty()) return ScoreReadResult();  // :1264-1268
    // target = reference, query = consensus (:1277-1280); the ref_aligner is the plain map-hifi preset (a = 1)
    const std::vector<Alignment> alns = gpu.align_pairs({reference_sequence}, {consensus}, {{0, 0}}, nullptr, 1);
    std::vector<Mapping> mappings;
    const Alignment &a = alns.at(0);
    if (!a.cigar.empty() && a.score >= aligner_stand_ins().min_dp_score) {
        Mapping m;
        m.query_start = static_cast<size_t>(a.p_start); m.query_end = static_cast<size_t>(a.p_end); m.query_len = consensus.size();
        m.target_start = static_cast<size_t>(a.t_start); m.target_end = static_cast<size_t>(a.t_end); m.target_len = reference_sequence.size();
        m.nm = static_cast<size_t>(a.nm); m.forward = true; m.cigar = a.cigar;
        mappings.push_back(std::move(m));
    }
    if (mappings.empty()) return ScoreReadResult();  // "Failed to align consensus to reference genome" (:1283-1288)
    const auto sel = select_best_mapping(mappings, true, true, std::nullopt);  // :1291-1295
    if (!sel.first) throw HostError("called `Option::unwrap()` on a `None` value");  // :1298: best_mapping.unwrap()
    const Mapping &d_map = mappings[*sel.first];
    // convert_mapping_to_cigar(d_map, None, None): alignment ops + soft clips for the unaligned consensus ends
    std::vector<std::pair<uint32_t, uint8_t>> cigar;
    if (d_map.query_start > 0) cigar.emplace_back(static_cast<uint32_t>(d_map.query_start), uint8_t(4));
    cigar.insert(cigar.end(), d_map.cigar.begin(), d_map.cigar.end());
    if (d_map.query_len > d_map.query_end) cigar.emplace_back(static_cast<uint32_t>(d_map.query_len - d_map.query_end), uint8_t(4));
    const int64_t start_align = ref_start + static_cast<int64_t>(d_map.target_start);  // :1315-1316
    const ScoreReadTargets t = prepare_score_read_targets(consensus, start_align, cigar, exons, is_forward_strand, settings);
    return score_read(gpu, t.dna_target, t.cdna_target, database, gene_name, settings);
}

Json PgxMappingDetails::to_json() const {
    Json j = Json::object();
    j.set("read_qname", read_qname).set("best_hla_id", best_hla_id).set("best_star_allele", best_star_allele);
    j.set("best_mapping_stats", best_mapping_stats.to_json()).set("is_ignored", is_ignored);
    return j;
}

std::string Diplotype::pharmcat_diplotype() const {
    auto wrap = [](const std::string &h) { return h.find('+') != std::string::npos ? "[" + h + "]" : h; };
    return wrap(hap1) + "/" + wrap(hap2);
}

Json Diplotype::to_json() const {
    Json j = Json::object();
    j.set("hap1", hap1).set("hap2", hap2).set("diplotype", hap1 + "/" + hap2);
    return j;
}

// ------------------------------------------------------------------------------------------
// realigner
// ------------------------------------------------------------------------------------------
HlaRealigner::HlaRealigner(GpuAligner &gpu, const std::vector<std::string> &gene_list, const HlaDatabase &database) : gpu_(gpu) {
    // create_hla_fasta (src/hla/realigner.rs:497-526): every allele of the listed genes that has a DNA sequence
    SeqList seqs;
    for (const auto &kv : database) {
        if (std::find(gene_list.begin(), gene_list.end(), kv.second.gene_name) == gene_list.end()) continue;
        if (!kv.second.dna_sequence) continue;
        if (kv.second.dna_sequence->size() > SP_MAX_PATTERN_LEN) { skipped_.push_back(kv.first); continue; }  // see HlaGeneIndex
        alleles_.push_back(&kv.second);
        seqs.push_back(*kv.second.dna_sequence);
    }
    index_ = gpu.prepare_patterns(seqs);
}

std::vector<PgxMappingDetails> HlaRealigner::realign_records(const std::vector<std::pair<std::string, std::string>> &reads, int n_candidates) {
    SeqList targets;
    for (const auto &r : reads) targets.push_back(r.second);
    const std::shared_ptr<ResidentSeqs> T = gpu_.upload(targets);  // one upload for K1 and the K4 tracebacks
    const std::unique_ptr<DeviceMatrix> D = gpu_.score_device(*T, *index_);
    return realign_records_scored(reads, *D, n_candidates, T.get());
}

HlaRealigner::HlaRealigner(GpuAligner &gpu, const std::vector<std::string> &gene_list, const HlaDatabase &database,
                           const std::map<std::string, HlaGeneDefinition> &gene_definitions)
    : gpu_(gpu), gene_definitions_(gene_definitions) {
    for (const std::string &g : gene_list)
        if (!gene_definitions_.count(g)) throw HostError("Gene definition for " + g + " not found.");  // :64-67
    SeqList seqs;
    for (const auto &kv : database) {
        if (std::find(gene_list.begin(), gene_list.end(), kv.second.gene_name) == gene_list.end()) continue;
        if (!kv.second.dna_sequence) continue;
        if (kv.second.dna_sequence->size() > SP_MAX_PATTERN_LEN) { skipped_.push_back(kv.first); continue; }
        alleles_.push_back(&kv.second);
        const bool fwd = gene_definitions_.at(kv.second.gene_name).is_forward_strand;
        seqs.push_back(fwd ? *kv.second.dna_sequence : reverse_complement(*kv.second.dna_sequence));  // :511-515
    }
    index_ = gpu.prepare_patterns(seqs);
}


// The selection loop of realign_record (src/hla/realigner.rs:116-146) for every read: the n best alleles by distance (K5)
// get a traceback (K4); minimap2's target is the allele here, so `unmapped` is allele-side.  The db_aligner is the plain
// map-hifi preset (a = 1, src/util/mapping.rs:8-14), which is what decides whether a hit would have been reported at all.
std::vector<HlaRealigner::BestHit> HlaRealigner::best_hits(const std::vector<std::pair<std::string, std::string>> &reads,
                                                           const DeviceMatrix &D, int n_candidates, const ResidentSeqs *resident_reads) {
    std::shared_ptr<ResidentSeqs> own;
    if (!resident_reads) {
        SeqList targets;
        for (const auto &r : reads) targets.push_back(r.second);
        own = gpu_.upload(targets);
        resident_reads = own.get();
    }
    const size_t A = alleles_.size();
    if (D.n_targets() != static_cast<int64_t>(reads.size()) || D.n_patterns() != static_cast<int64_t>(A))
        throw HostError("realign_records_scored: distance matrix has the wrong shape");
    // the best_n hits of every read (K5): minimap2 ranks its hits by alignment score, which under map-hifi (a = 1, b = 4, gaps
    // dearer) is about (allele bases aligned) - 5 x edits; the candidates are therefore the alleles with the smallest
    // candidate_edit_weight (5 = b + a) * (nm + unmapped) - |allele|, ties by database order.  A read covering part of a long allele keeps that
    // allele ahead of short alleles lying wholly inside the read, and an allele with fewer edits ahead of a longer one with more.
    // Measured against the affine cost model: weight 5 -> 91 of 96 read assignments agree, weight 1 -> 83 of 96.
    const int k = std::max(1, std::min(n_candidates, 16));
    const SeqList &allele_seqs = index_->sequences();
    size_t max_len = 0;
    for (const std::string &a : allele_seqs) max_len = std::max(max_len, a.size());
    std::vector<int32_t> bias(A);
    for (size_t a = 0; a < A; ++a) bias[a] = static_cast<int32_t>(max_len - allele_seqs[a].size());
    std::vector<int32_t> cand, cand_dist;
    if (A && !reads.empty()) gpu_.row_topk(D, k, cand, cand_dist, &bias, aligner_stand_ins().candidate_edit_weight);
    std::vector<std::pair<int32_t, int32_t>> pairs;
    std::vector<size_t> first_pair(reads.size() + 1, 0);
    for (size_t r = 0; r < reads.size(); ++r) {
        first_pair[r] = pairs.size();
        if (A && !reads[r].second.empty())
            for (int q = 0; q < k; ++q)
                if (cand[r * static_cast<size_t>(k) + static_cast<size_t>(q)] >= 0)
                    pairs.emplace_back(static_cast<int32_t>(r), cand[r * static_cast<size_t>(k) + static_cast<size_t>(q)]);
    }
    first_pair[reads.size()] = pairs.size();
    std::vector<Alignment> alns = gpu_.align_pairs(*resident_reads, index_->resident(), pairs, nullptr, 1, nullptr, aligner_stand_ins().min_dp_score);  // both sides already on the device

    std::vector<BestHit> out(reads.size());
    for (size_t r = 0; r < reads.size(); ++r) {
        const size_t read_len = reads[r].second.size();
        BestHit &b = out[r];
        b.stats = MappingStats(read_len, read_len, 0);  // src/hla/realigner.rs:124
        for (size_t q = first_pair[r]; q < first_pair[r + 1]; ++q) {
            Alignment &a = alns[q];
            if (a.cigar.empty() || a.score < aligner_stand_ins().min_dp_score) continue;  // no hit reported
            const size_t target_len = allele_seqs[static_cast<size_t>(pairs[q].second)].size();
            const size_t unmapped = target_len - static_cast<size_t>(a.p_end - a.p_start);
            const MappingStats stats(target_len, static_cast<size_t>(a.nm), unmapped);
            if (stats.mapping_score() <= 0.5 && stats.custom_score(false) <= 0.03 &&
                stats.custom_score(false) < b.stats.custom_score(false)) {  // :137-146
                b.stats = stats;
                b.allele = pairs[q].second;
                b.aln = std::move(a);
            }
        }
    }
    return out;
}

std::vector<PgxMappingDetails> HlaRealigner::realign_records_scored(const std::vector<std::pair<std::string, std::string>> &reads,
                                                                    const DeviceMatrix &D, int n_candidates, const ResidentSeqs *resident_reads) {
    const std::vector<BestHit> hits = best_hits(reads, D, n_candidates, resident_reads);
    std::vector<PgxMappingDetails> out;
    for (size_t r = 0; r < reads.size(); ++r) {
        PgxMappingDetails d;
        d.read_qname = reads[r].first;
        d.best_mapping_stats.dna_stats = hits[r].stats;
        if (hits[r].allele < 0) {  // :149-162
            d.best_hla_id = "REFERENCE"; d.best_star_allele = "REFERENCE"; d.is_ignored = true;
        } else {
            const HlaAlleleDefinition *def = alleles_[static_cast<size_t>(hits[r].allele)];
            d.best_hla_id = def->hla_id;
            d.best_star_allele = def->gene_name + "*" + join_star(def->star_allele);  // :204-211
            d.is_ignored = false;
        }
        out.push_back(std::move(d));
    }
    return out;
}

// K4 reports (pattern, text) = (allele, read); minimap2's view at this call site is (query, target) = (read, allele), so
// the roles and the insertion / deletion letters swap
static Mapping read_vs_allele_mapping(const Alignment &a, size_t read_len, size_t allele_len) {
    Mapping m;
    m.query_start = static_cast<size_t>(a.t_start); m.query_end = static_cast<size_t>(a.t_end); m.query_len = read_len;
    m.target_start = static_cast<size_t>(a.p_start); m.target_end = static_cast<size_t>(a.p_end); m.target_len = allele_len;
    m.nm = static_cast<size_t>(a.nm);
    m.forward = true;
    for (const auto &op : a.cigar) m.cigar.emplace_back(op.first, op.second == 1 ? uint8_t(2) : op.second == 2 ? uint8_t(1) : op.second);
    return m;
}

std::vector<RealignmentResult> HlaRealigner::realign_records_full(const std::vector<std::pair<std::string, std::string>> &reads,
                                                                  int n_candidates) {
    if (gene_definitions_.empty()) throw HostError("realign_records_full: the realigner was built without gene definitions");
    SeqList targets;
    for (const auto &r : reads) targets.push_back(r.second);
    const std::shared_ptr<ResidentSeqs> T = gpu_.upload(targets);
    const std::unique_ptr<DeviceMatrix> D = gpu_.score_device(*T, *index_);
    const std::vector<BestHit> hits = best_hits(reads, *D, n_candidates, T.get());
    const SeqList &allele_seqs = index_->sequences();  // hg38 orientation

    std::vector<RealignmentResult> out(reads.size());
    // second batch: the buffered read segment of every accepted read against its gene's hg38 sequence (:216-231)
    SeqList ref_texts, seg_patterns;
    std::map<std::string, int32_t> ref_index;
    std::vector<std::pair<int32_t, int32_t>> seg_pairs;
    std::vector<size_t> seg_read, seg_buffered_start;
    for (size_t r = 0; r < reads.size(); ++r) {
        RealignmentResult &res = out[r];
        PgxMappingDetails &d = res.mapping_details;
        d.read_qname = reads[r].first;
        d.best_mapping_stats.dna_stats = hits[r].stats;
        if (hits[r].allele < 0) {  // :149-162: RealignmentResult::failure
            d.best_hla_id = "REFERENCE"; d.best_star_allele = "REFERENCE"; d.is_ignored = true;
            continue;
        }
        const HlaAlleleDefinition *def = alleles_[static_cast<size_t>(hits[r].allele)];
        const std::string best_star = join_star(def->star_allele);
        d.best_hla_id = def->hla_id;
        d.best_star_allele = def->gene_name + "*" + best_star;
        d.is_ignored = false;
        res.gene_name = def->gene_name;
        const std::string &allele_seq = allele_seqs[static_cast<size_t>(hits[r].allele)];
        const Mapping bm = read_vs_allele_mapping(hits[r].aln, reads[r].second.size(), allele_seq.size());
        res.read_mapping_stats.add_mapping(def->hla_id, std::nullopt, detailed_mapping_stats(bm, allele_seq, reads[r].second));  // :199-201
        res.read_mapping_stats.set_best_match(def->hla_id, best_star);
        const size_t buffer = 1000;  // :221-224
        const size_t bs = bm.query_start > buffer ? bm.query_start - buffer : 0;
        const size_t be = std::min(bm.query_end + buffer, reads[r].second.size());
        auto it = ref_index.find(def->gene_name);
        if (it == ref_index.end()) {
            it = ref_index.emplace(def->gene_name, static_cast<int32_t>(ref_texts.size())).first;
            ref_texts.push_back(gene_definitions_.at(def->gene_name).reference_sequence);
        }
        seg_pairs.emplace_back(it->second, static_cast<int32_t>(seg_patterns.size()));
        seg_patterns.push_back(reads[r].second.substr(bs, be - bs));
        seg_read.push_back(r);
        seg_buffered_start.push_back(bs);
    }
    const std::vector<Alignment> seg_alns = gpu_.align_pairs(ref_texts, seg_patterns, seg_pairs, nullptr, 1);

    // third batch: the best allele against hg38 for the reads whose hg38 mapping starts no earlier than the allele mapping (:268-290)
    struct Pending { size_t r; Mapping ref_mapping; size_t hg38_start, hg38_end; int32_t allele_pair = -1; };
    std::vector<Pending> pending;
    SeqList allele_patterns;
    std::vector<std::pair<int32_t, int32_t>> allele_pairs;
    for (size_t q = 0; q < seg_pairs.size(); ++q) {
        const size_t r = seg_read[q];
        const std::string &ref = ref_texts[static_cast<size_t>(seg_pairs[q].first)];
        std::vector<Mapping> mappings;
        const Alignment &a = seg_alns[q];
        if (!a.cigar.empty() && a.score >= aligner_stand_ins().min_dp_score) {  // (query, target) = (read segment, hg38) = (pattern, text)
            Mapping m;
            m.query_start = static_cast<size_t>(a.p_start); m.query_end = static_cast<size_t>(a.p_end); m.query_len = seg_patterns[q].size();
            m.target_start = static_cast<size_t>(a.t_start); m.target_end = static_cast<size_t>(a.t_end); m.target_len = ref.size();
            m.nm = static_cast<size_t>(a.nm); m.forward = true; m.cigar = a.cigar;
            mappings.push_back(std::move(m));
        }
        const auto sel = select_best_mapping(mappings, true, true, std::nullopt);  // :234-238
        if (!sel.first) continue;  // "Remapping of {qname} to reference failed, ignoring." (:337-339)
        Pending p;
        p.r = r;
        p.ref_mapping = mappings[*sel.first];
        p.hg38_start = seg_buffered_start[q] + p.ref_mapping.query_start;  // :257-258
        p.hg38_end = seg_buffered_start[q] + p.ref_mapping.query_end;
        const size_t db_start = static_cast<size_t>(hits[r].aln.t_start);
        if (!(p.hg38_start < db_start)) {
            p.allele_pair = static_cast<int32_t>(allele_pairs.size());
            allele_pairs.emplace_back(seg_pairs[q].first, static_cast<int32_t>(allele_patterns.size()));
            allele_patterns.push_back(allele_seqs[static_cast<size_t>(hits[r].allele)]);  // hg38 orientation (:276-282)
        }
        pending.push_back(std::move(p));
    }
    const std::vector<Alignment> allele_alns = gpu_.align_pairs(ref_texts, allele_patterns, allele_pairs, nullptr, 1);

    for (const Pending &p : pending) {
        const size_t r = p.r;
        const HlaAlleleDefinition *def = alleles_[static_cast<size_t>(hits[r].allele)];
        const std::string &ref = gene_definitions_.at(def->gene_name).reference_sequence;
        const size_t db_start = static_cast<size_t>(hits[r].aln.t_start), db_end = static_cast<size_t>(hits[r].aln.t_end);
        const size_t bm_target_start = static_cast<size_t>(hits[r].aln.p_start);
        RealignedHlaRecord rec;
        rec.segment_start = std::min(db_start, p.hg38_start);  // :263-264
        rec.segment_end = std::max(db_end, p.hg38_end);
        size_t d = p.ref_mapping.target_start, h = 0;
        bool from_ref = true;
        if (p.allele_pair >= 0) {
            const Alignment &a = allele_alns[static_cast<size_t>(p.allele_pair)];
            std::vector<Mapping> mappings;
            if (!a.cigar.empty() && a.score >= aligner_stand_ins().min_dp_score) {
                Mapping m;
                m.query_start = static_cast<size_t>(a.p_start); m.query_end = static_cast<size_t>(a.p_end);
                m.query_len = allele_patterns[static_cast<size_t>(allele_pairs[static_cast<size_t>(p.allele_pair)].second)].size();
                m.target_start = static_cast<size_t>(a.t_start); m.target_end = static_cast<size_t>(a.t_end); m.target_len = ref.size();
                m.nm = static_cast<size_t>(a.nm); m.forward = true; m.cigar = a.cigar;
                mappings.push_back(std::move(m));
            }
            const auto sel = select_best_mapping(mappings, false, true, std::nullopt);  // :296-300
            if (sel.first) {
                const Mapping &am = mappings[*sel.first];
                const long added = std::max(0l, static_cast<long>(am.target_start) - static_cast<long>(am.query_start));  // :304
                d = static_cast<size_t>(added) + bm_target_start;                                                       // :310
                h = hpc_pos(ref, static_cast<size_t>(added)) + hpc_pos(*def->dna_sequence, bm_target_start);          // :312
                from_ref = false;
            }  // else: "Failed to map allele ... ignoring offset adjustment" (:315-319)
        }
        if (from_ref) h = hpc_pos(ref, d);  // :269-272, :317-318
        rec.dna_offset = d; rec.hpc_offset = h;
        rec.dna_sequence = reads[r].second.substr(rec.segment_start, rec.segment_end - rec.segment_start);  // RealignedHlaRecord::new
        rec.hpc_sequence = hpc(rec.dna_sequence);
        out[r].realigned_record = std::move(rec);
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

HlaGeneIndex::HlaGeneIndex(GpuAligner &gpu, const HlaDatabase &database, const std::string &gene_name, const DiplotypeSettings &settings)
    : gene_name_(gene_name) {
    for (const auto &kv : database)
        if (is_allowed_allele_def(kv.second, gene_name, settings)) {
            if (!kv.second.dna_sequence)
                throw HostError("diplotype_hla_gene: allele " + kv.first + " has no DNA sequence (pair ranking needs hla_require_dna)");
            if (kv.second.dna_sequence->size() > SP_MAX_PATTERN_LEN || kv.second.cdna_sequence.size() > SP_MAX_PATTERN_LEN) {
                // the device formats hold patterns up to SP_MAX_PATTERN_LEN rows; a longer allele is left out of the index (it
                // cannot be called) instead of failing the whole gene -- the reference's aligner has no such limit
                skipped_.push_back(kv.first);
                continue;
            }
            gene_db_.emplace(kv.first, kv.second);
        }
    SeqList cdna;
    for (const auto &kv : gene_db_) { allowed_.push_back(&kv.second); cdna.push_back(kv.second.cdna_sequence); }
    realigner_.reset(new HlaRealigner(gpu, {gene_name}, gene_db_));  // DNA index, same allele order
    if (!settings.disable_cdna_scoring) cdna_ = gpu.prepare_patterns(cdna);
}

// the call from resident targets: Td = DNA targets (K1 now, K4 of the per-read assignment below), Tc = cDNA targets or null
HlaGeneCall diplotype_from_targets(GpuAligner &gpu, HlaGeneIndex &index, const std::vector<std::string> &qnames,
                                          const std::shared_ptr<ResidentSeqs> &Td, const std::shared_ptr<ResidentSeqs> &Tc,
                                          const DiplotypeSettings &settings);

HlaGeneCall diplotype_hla_gene(GpuAligner &gpu, HlaGeneIndex &index, const std::vector<HlaRead> &reads, const DiplotypeSettings &settings) {
    if (reads.empty() || index.allowed_.empty()) return diplotype_from_targets(gpu, index, {}, nullptr, nullptr, settings);
    SeqList dna_targets, cdna_targets;
    std::vector<std::string> qnames;
    for (const auto &r : reads) { dna_targets.push_back(r.dna_target); cdna_targets.push_back(r.cdna_target); qnames.push_back(r.qname); }
    const bool dual = !(settings.disable_cdna_scoring || !index.cdna_);
    return diplotype_from_targets(gpu, index, qnames, gpu.upload(dna_targets), dual ? gpu.upload(cdna_targets) : nullptr, settings);
}

HlaGeneCall diplotype_hla_gene_records(GpuAligner &gpu, HlaGeneIndex &index, const std::vector<HlaRecord> &records,
                                       const std::vector<std::pair<uint64_t, uint64_t>> &exons, bool is_forward_strand,
                                       const DiplotypeSettings &settings) {
    if (records.empty() || index.allowed_.empty()) return diplotype_from_targets(gpu, index, {}, nullptr, nullptr, settings);
    // the reads go up once, with one extra sequence "N": the cDNA target of a read whose splice is empty (src/hla/caller.rs:1355-1366)
    SeqList seqs;
    std::vector<std::string> qnames;
    for (const auto &r : records) { seqs.push_back(r.sequence); qnames.push_back(r.qname); }
    seqs.push_back("N");
    const std::shared_ptr<ResidentSeqs> reads = gpu.upload(seqs);
    const bool dual = !(settings.disable_cdna_scoring || !index.cdna_);
    std::vector<std::pair<size_t, std::vector<std::pair<size_t, size_t>>>> dna_pieces, cdna_pieces;
    std::vector<bool> dna_rc, cdna_rc;
    for (size_t q = 0; q < records.size(); ++q) {
        dna_pieces.push_back({q, {{0, records[q].sequence.size()}}});
        dna_rc.push_back(!is_forward_strand);
        if (!dual) continue;
        const auto seg = splice_segments(records[q].sequence.size(), records[q].pos, records[q].cigar, exons);
        size_t total = 0;
        for (const auto &sg : seg.first) total += sg.second - sg.first;
        if (total == 0) { cdna_pieces.push_back({records.size(), {{0, 1}}}); cdna_rc.push_back(false); }
        else { cdna_pieces.push_back({q, seg.first}); cdna_rc.push_back(!is_forward_strand); }
    }
    return diplotype_from_targets(gpu, index, qnames, gpu.derive(*reads, dna_pieces, dna_rc), dual ? gpu.derive(*reads, cdna_pieces, cdna_rc) : nullptr,
                                  settings);
}

HlaGeneCall diplotype_from_targets(GpuAligner &gpu, HlaGeneIndex &index, const std::vector<std::string> &qnames,
                                          const std::shared_ptr<ResidentSeqs> &Td, const std::shared_ptr<ResidentSeqs> &Tc,
                                          const DiplotypeSettings &settings) {
    const auto &allowed = index.allowed_;
    HlaGeneCall call;
    if (qnames.empty() || allowed.empty()) {  // sentinels of src/hla/caller.rs:32-37
        call.hla_id1 = call.hla_id2 = "NO_READS";
        call.diplotype = {"NO_READS", "NO_READS"};
        return call;
    }
    const int64_t R = static_cast<int64_t>(qnames.size());
    const std::unique_ptr<DeviceMatrix> Dd = gpu.score_device(*Td, index.realigner_->index());
    std::vector<sp_pair_rec> top;
    if (settings.disable_cdna_scoring || !index.cdna_ || !Tc) {
        top = gpu.pair_minsum_topk(*Dd, nullptr, 10);
    } else {
        const std::unique_ptr<DeviceMatrix> Dc = gpu.score_device(*Tc, *index.cdna_);
        top = gpu.pair_minsum_topk(*Dc, Dd.get(), 10);  // (cDNA, DNA) lexicographic: src/hla/mapping.rs:111-117
    }
    if (top.empty()) throw HostError("diplotype_hla_gene: pair ranking returned nothing");
    const bool dual = !(settings.disable_cdna_scoring || !index.cdna_);
    const sp_pair_rec &b = top[0];
    call.pair_score_cdna = dual ? b.score : 0;
    call.pair_score_dna = dual ? b.score2 : b.score;
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
    auto star = [&](const std::string &id) { return "*" + join_star(index.gene_db_.at(id).star_allele); };  // :1046-1065
    call.diplotype = {star(call.hla_id1), star(call.hla_id2)};
    // per-read database assignment, as HlaRealigner::realign_record reports it (restricted to this gene): the DNA
    // distances are already on the device
    std::vector<std::pair<std::string, std::string>> qs;
    for (size_t r = 0; r < qnames.size(); ++r) qs.emplace_back(qnames[r], Td->sequences()[r]);
    call.mapping_details = index.realigner_->realign_records_scored(qs, *Dd, 5, Td.get());
    return call;
}

HlaGeneCall diplotype_hla_gene(GpuAligner &gpu, const HlaDatabase &database, const std::string &gene_name, const std::vector<HlaRead> &reads,
                               const DiplotypeSettings &settings) {
    HlaGeneIndex index(gpu, database, gene_name, settings);
    return diplotype_hla_gene(gpu, index, reads, settings);
}

}  // namespace starphase
