// sp_host_gpu.cpp -- GpuAligner: the only place the host touches libstarphase_gpu.so.
#include <algorithm>

#include "starphase_host.hpp"

namespace starphase {

namespace {
struct Packed {
    std::string bases;
    std::vector<int64_t> offs;
    sp_seqset set;
    explicit Packed(const SeqList &seqs) {
        size_t total = 0;
        for (const auto &s : seqs) total += s.size();
        bases.reserve(total + 1);
        offs.reserve(seqs.size() + 1);
        offs.assign(1, 0);
        for (const auto &s : seqs) {
            bases += s;
            offs.push_back(static_cast<int64_t>(bases.size()));
        }
        if (bases.empty()) bases.push_back('N');  // never read: keeps the pointer non-null
        set.bases = reinterpret_cast<const uint8_t *>(bases.data());
        set.offsets = offs.data();
        set.n = static_cast<int64_t>(seqs.size());
    }
};
}  // namespace

GpuAligner::GpuAligner(int device) {
    const sp_status st = sp_ctx_create(device, nullptr, &ctx_);
    if (st != SP_OK) throw HostError(std::string("sp_ctx_create: ") + sp_last_error(nullptr));
}

GpuAligner::~GpuAligner() {
    for (uint32_t *b : cig_buf_) sp_pinned_free(ctx_, b);
    sp_ctx_destroy(ctx_);
}

void GpuAligner::share_device(bool on) { check(sp_ctx_share_device(ctx_, on ? 1 : 0), "sp_ctx_share_device"); }

uint32_t *GpuAligner::cigar_buffer(int slot, int64_t entries) {
    if (entries > cig_cap_[slot]) {
        sp_pinned_free(ctx_, cig_buf_[slot]);
        cig_buf_[slot] = nullptr; cig_cap_[slot] = 0;
        const int64_t want = entries + entries / 4;
        void *p = nullptr;
        check(sp_pinned_alloc(ctx_, static_cast<size_t>(want) * 4, &p), "sp_pinned_alloc");
        cig_buf_[slot] = static_cast<uint32_t *>(p); cig_cap_[slot] = want;
    }
    return cig_buf_[slot];
}

void GpuAligner::check(sp_status st, const char *what) {
    if (st != SP_OK) throw HostError(std::string(what) + ": " + sp_last_error(ctx_));
}

uint64_t GpuAligner::launch_count() const { return sp_launch_count(ctx_); }

std::vector<int32_t> GpuAligner::score_batch(const SeqList &targets, const SeqList &patterns, std::vector<int32_t> *end_col) {
    Packed t(targets), p(patterns);
    std::vector<int32_t> D(std::max<size_t>(targets.size() * patterns.size(), 1));
    if (end_col) end_col->assign(D.size(), 0);
    check(sp_score_batch(ctx_, &t.set, &p.set, SP_INFIX, D.data(), end_col ? end_col->data() : nullptr), "sp_score_batch");
    D.resize(targets.size() * patterns.size());
    if (end_col) end_col->resize(D.size());
    return D;
}

void GpuAligner::score_spans(const SeqList &targets, const SeqList &patterns, std::vector<int32_t> &D, std::vector<int32_t> &start,
                             std::vector<int32_t> &end, int max_dist_permille) {
    Packed t(targets), p(patterns);
    const size_t n = std::max<size_t>(targets.size() * patterns.size(), 1);
    D.assign(n, 0); start.assign(n, 0); end.assign(n, 0);
    check(sp_score_spans_filtered(ctx_, &t.set, &p.set, max_dist_permille, D.data(), start.data(), end.data()), "sp_score_spans");
}

std::shared_ptr<ResidentSeqs> GpuAligner::upload(const SeqList &seqs) {
    std::shared_ptr<ResidentSeqs> r(new ResidentSeqs());
    r->seqs_ = seqs;
    Packed p(seqs);
    check(sp_targets_create(ctx_, &p.set, &r->t_), "sp_targets_create");
    return r;
}

std::shared_ptr<ResidentSeqs> GpuAligner::derive(const ResidentSeqs &src, const std::vector<std::pair<size_t, std::vector<std::pair<size_t, size_t>>>> &pieces,
                                                 const std::vector<bool> &revcomp) {
    if (!revcomp.empty() && revcomp.size() != pieces.size()) throw HostError("derive: one revcomp flag per output expected");
    std::shared_ptr<ResidentSeqs> r(new ResidentSeqs());
    std::vector<int32_t> src_index, iv_begin, iv_end;
    std::vector<int64_t> iv_off(1, 0);
    std::vector<uint8_t> rc;
    for (size_t q = 0; q < pieces.size(); ++q) {
        if (pieces[q].first >= src.seqs_.size()) throw HostError("derive: source index outside the set");
        const std::string &s = src.seqs_[pieces[q].first];
        std::string out;
        for (const auto &iv : pieces[q].second) {
            if (iv.first > iv.second || iv.second > s.size()) throw HostError("derive: interval outside its source sequence");
            out.append(s, iv.first, iv.second - iv.first);
            iv_begin.push_back(static_cast<int32_t>(iv.first)); iv_end.push_back(static_cast<int32_t>(iv.second));
        }
        const bool flip = !revcomp.empty() && revcomp[q];
        r->seqs_.push_back(flip ? reverse_complement(out) : out);  // the host copy (lengths, CIGAR / MD strings)
        src_index.push_back(static_cast<int32_t>(pieces[q].first));
        iv_off.push_back(static_cast<int64_t>(iv_begin.size()));
        rc.push_back(flip ? 1 : 0);
    }
    if (iv_begin.empty()) { iv_begin.push_back(0); iv_end.push_back(0); }
    if (src_index.empty()) { src_index.push_back(0); rc.push_back(0); }
    check(sp_targets_derive(ctx_, src.t_, static_cast<int64_t>(pieces.size()), src_index.data(), iv_off.data(), iv_begin.data(), iv_end.data(), rc.data(),
                            &r->t_),
          "sp_targets_derive");
    return r;
}

ResidentSeqs::~ResidentSeqs() { sp_targets_destroy(t_); }

// host sequence lists: uploaded for this call only
std::vector<Alignment> GpuAligner::align_pairs(const SeqList &targets, const SeqList &patterns,
                                               const std::vector<std::pair<int32_t, int32_t>> &pairs,
                                               const std::vector<std::pair<int32_t, int32_t>> *windows, int match_score) {
    if (pairs.empty()) return {};
    const std::shared_ptr<ResidentSeqs> t = upload(targets), p = upload(patterns);
    return align_pairs(*t, *p, pairs, windows, match_score);
}

// DP score of a run-length CIGAR under (a, b = 4, q = 6, e = 2, q2 = 26, e2 = 1): dp_score (sp_host_core.cpp) on raw pool entries
static long raw_dp_score(const uint32_t *cig, int32_t n, long a) {
    long s = 0;
    for (int32_t k = 0; k < n; ++k) {
        const long len = cig[k] >> 4;
        const uint32_t op = cig[k] & 15u;
        if (op == 7) s += a * len;
        else if (op == 8) s -= 4 * len;
        else s -= std::min(6 + 2 * len, 26 + len);
    }
    return s;
}

std::vector<Alignment> GpuAligner::align_pairs(const ResidentSeqs &texts, const ResidentSeqs &pats,
                                               const std::vector<std::pair<int32_t, int32_t>> &pairs,
                                               const std::vector<std::pair<int32_t, int32_t>> *windows, int match_score,
                                               const std::vector<std::pair<int32_t, int32_t>> *bounds, long report_floor) {
    if (pairs.empty()) return {};
    if (bounds && bounds->size() != pairs.size()) throw HostError("align_pairs: one bounds entry per pair expected");
    const UnitResult unit = align_pairs_unit(texts, pats, pairs, windows);
    const SeqList &targets = texts.sequences(), &patterns = pats.sequences();
    const bool refine = match_score > 0 && aligner_stand_ins().affine_refine;
    // K9 for every placement whose band fits: half width = half the diagonal hull of the unit-cost path (text column - pattern
    // row along the CIGAR, absolute text coordinates) + 24, centre = the middle of the hull
    std::vector<size_t> sel;
    std::vector<int32_t> pt, pp, wb, we, centre, bandw;
    if (refine) {
        for (size_t q = 0; q < pairs.size(); ++q) {
            const sp_align_rec &r = unit.recs[q];
            if (r.n_cigar == 0) continue;
            const int64_t base = windows ? (*windows)[q].first : 0;
            int64_t d = base + r.t_start - r.p_start, lo = d, hi = d;
            const uint32_t *cg = unit.cigar + r.cigar_off;
            for (int32_t k = 0; k < r.n_cigar; ++k) {
                const uint32_t op = cg[k] & 15u;
                if (op == 1) d -= cg[k] >> 4;       // I: pattern bases without text
                else if (op == 2) d += cg[k] >> 4;  // D: text bases without pattern
                else continue;
                lo = std::min(lo, d); hi = std::max(hi, d);
            }
            const int64_t w = (hi - lo + 1) / 2 + 24;
            if (w > 255) continue;
            const int64_t c_abs = (lo + hi >= 0) ? (lo + hi) / 2 : -((-(lo + hi) + 1) / 2);  // floor: independent of the coordinate origin
            const int64_t W = w <= 31 ? 31 : w <= 63 ? 63 : w <= 127 ? 127 : 255;            // the width classes of the library
            const int64_t m = static_cast<int64_t>(patterns[static_cast<size_t>(pairs[q].second)].size());
            const int64_t n = static_cast<int64_t>(targets[static_cast<size_t>(pairs[q].first)].size());
            const int64_t blo = bounds ? (*bounds)[q].first : 0, bhi = bounds ? (*bounds)[q].second : n;
            // the text window that holds every band cell: columns (1-based) i + c - W .. i + c + W for i = 1 .. m
            const int64_t b = std::max<int64_t>(blo, std::min<int64_t>(bhi, c_abs - W));
            const int64_t e = std::min<int64_t>(bhi, std::max<int64_t>(b, m + c_abs + W + 1));
            sel.push_back(q);
            pt.push_back(pairs[q].first); pp.push_back(pairs[q].second);
            wb.push_back(static_cast<int32_t>(b)); we.push_back(static_cast<int32_t>(e));
            centre.push_back(static_cast<int32_t>(c_abs - b));
            bandw.push_back(static_cast<int32_t>(W));
        }
    }
    std::vector<sp_align_rec> arecs(std::max<size_t>(sel.size(), 1));
    std::vector<int32_t> ascores(std::max<size_t>(sel.size(), 1));
    const uint32_t *acig = nullptr;
    if (!sel.empty()) {
        const sp_affine_costs costs = {match_score, 4, 6, 2, 26, 1};
        int64_t cap = std::max<int64_t>(1 << 16, static_cast<int64_t>(sel.size()) * affine_entries_per_pair_), used = 0;
        for (int attempt = 0;; ++attempt) {
            uint32_t *buf = cigar_buffer(1, cap);
            const sp_status st = sp_align_affine_resident(ctx_, texts.t_, pats.t_, static_cast<int64_t>(sel.size()), pt.data(), pp.data(), wb.data(),
                                                          we.data(), centre.data(), 0, bandw.data(), &costs, arecs.data(), ascores.data(), buf, cap,
                                                          &used);
            if (st == SP_ERR_RANGE && attempt == 0 && used > cap) { cap = used; continue; }
            check(st, "sp_align_affine_resident");
            acig = buf;
            break;
        }
        affine_entries_per_pair_ = std::max<int64_t>(affine_entries_per_pair_, 2 * used / static_cast<int64_t>(sel.size()) + 64);
    }
    std::vector<Alignment> out(pairs.size());
    auto fill = [](Alignment &a, const sp_align_rec &r, const uint32_t *pool, bool with_cigar) {
        a.dist = r.dist; a.nm = r.nm; a.p_start = r.p_start; a.p_end = r.p_end; a.t_start = r.t_start; a.t_end = r.t_end;
        if (!with_cigar) return;
        a.cigar.reserve(static_cast<size_t>(r.n_cigar));
        const uint32_t *cg = pool + r.cigar_off;
        for (int32_t k = 0; k < r.n_cigar; ++k) a.cigar.emplace_back(cg[k] >> 4, static_cast<uint8_t>(cg[k] & 15u));
    };
    size_t k = 0;
    for (size_t q = 0; q < pairs.size(); ++q) {
        Alignment &a = out[q];
        if (k < sel.size() && sel[k] == q) {  // the mapping fields come from the affine alignment
            fill(a, arecs[k], acig, ascores[k] >= report_floor);
            a.refined = true;
            a.score = ascores[k];
            a.t_base = wb[k];
            if (!windows) {  // callers without windows read text coordinates from the start of the text
                a.t_start += static_cast<int32_t>(a.t_base); a.t_end += static_cast<int32_t>(a.t_base);
                a.t_base = 0;
            }
            ++k;
        } else {
            const sp_align_rec &r = unit.recs[q];
            if (match_score > 0) {
                a.score = r.n_cigar ? raw_dp_score(unit.cigar + r.cigar_off, r.n_cigar, match_score) : 0;
                a.t_base = windows ? (*windows)[q].first : 0;
            }
            fill(a, r, unit.cigar, match_score <= 0 || a.score >= report_floor);
        }
    }
    return out;
}

GpuAligner::UnitResult GpuAligner::align_pairs_unit(const ResidentSeqs &texts, const ResidentSeqs &pats,
                                                    const std::vector<std::pair<int32_t, int32_t>> &pairs,
                                                    const std::vector<std::pair<int32_t, int32_t>> *windows) {
    if (windows && windows->size() != pairs.size()) throw HostError("align_pairs: one window per pair expected");
    const SeqList &targets = texts.sequences(), &patterns = pats.sequences();
    std::vector<int32_t> pt(pairs.size()), pp(pairs.size()), wb, we;
    if (windows) {
        wb.resize(pairs.size()); we.resize(pairs.size());
        for (size_t q = 0; q < pairs.size(); ++q) { wb[q] = (*windows)[q].first; we[q] = (*windows)[q].second; }
    }
    // capacity of the run-length CIGAR pool: HiFi-like pairs need a few dozen entries each; start from a generous guess and
    // retry once with the exact need the library reports (SP_ERR_RANGE leaves *cigar_used = entries needed)
    int64_t worst = 0;
    for (size_t q = 0; q < pairs.size(); ++q) {
        pt[q] = pairs[q].first; pp[q] = pairs[q].second;
        if (pt[q] < 0 || pp[q] < 0 || static_cast<size_t>(pt[q]) >= targets.size() || static_cast<size_t>(pp[q]) >= patterns.size())
            throw HostError("align_pairs: pair index outside the sequence sets");
        const int64_t m = static_cast<int64_t>(patterns[static_cast<size_t>(pp[q])].size());
        int64_t n = static_cast<int64_t>(targets[static_cast<size_t>(pt[q])].size());
        if (windows) {
            if (wb[q] < 0 || we[q] < wb[q] || we[q] > n) throw HostError("align_pairs: window outside its text");
            n = we[q] - wb[q];
        }
        worst += m + std::min(n, 2 * m) + 1;
    }
    UnitResult res;
    res.recs.resize(std::max<size_t>(pairs.size(), 1));
    int64_t cap = std::min<int64_t>(worst, std::max<int64_t>(1 << 16, static_cast<int64_t>(pairs.size()) * cigar_entries_per_pair_));
    int64_t used = 0;
    for (int attempt = 0;; ++attempt) {
        uint32_t *buf = cigar_buffer(0, std::max<int64_t>(cap, 1));
        const sp_status st = sp_align_resident(ctx_, texts.t_, pats.t_, static_cast<int64_t>(pairs.size()), pt.data(), pp.data(),
                                               windows ? wb.data() : nullptr, windows ? we.data() : nullptr, res.recs.data(), buf, cap, &used);
        if (st == SP_ERR_RANGE && attempt == 0 && used > cap) { cap = used; continue; }
        check(st, "sp_align_resident");
        res.cigar = buf;
        break;
    }
    // remember how long the CIGARs of this workload are (divergent CYP2D6 templates need ~400 entries, HLA alleles a few dozen)
    cigar_entries_per_pair_ = std::max<int64_t>(cigar_entries_per_pair_, 2 * used / static_cast<int64_t>(pairs.size()) + 64);
    return res;
}

std::vector<sp_pair_rec> GpuAligner::pair_minsum_topk(const std::vector<int32_t> &D, const std::vector<int32_t> *D2, int64_t R, int64_t A,
                                                      int k) {
    std::vector<sp_pair_rec> out(static_cast<size_t>(std::max(k, 1)));
    int n = 0;
    check(sp_pair_minsum_topk_host(ctx_, D.data(), D2 ? D2->data() : nullptr, R, A, k, out.data(), &n), "sp_pair_minsum_topk_host");
    out.resize(static_cast<size_t>(n));
    return out;
}

void GpuAligner::variant_match(const std::vector<std::vector<uint8_t>> &seq_alleles, const std::vector<std::vector<uint8_t>> &hap_alleles,
                               const std::vector<uint8_t> &is_vi, std::vector<uint32_t> &vi_match, std::vector<uint32_t> &all_match) {
    const size_t nv = is_vi.size();
    auto flat = [&](const std::vector<std::vector<uint8_t>> &rows, const char *what) {
        std::vector<uint8_t> out;
        out.reserve(rows.size() * nv + 1);
        for (const auto &r : rows) {
            if (r.size() != nv) throw HostError(std::string("variant_match: a ") + what + " row differs in length from is_vi");
            out.insert(out.end(), r.begin(), r.end());
        }
        if (out.empty()) out.push_back(0);
        return out;
    };
    const std::vector<uint8_t> s = flat(seq_alleles, "sequence"), h = flat(hap_alleles, "haplotype");
    std::vector<uint8_t> vi = is_vi;
    if (vi.empty()) vi.push_back(0);
    const size_t cells = seq_alleles.size() * hap_alleles.size();
    vi_match.assign(std::max<size_t>(cells, 1), 0); all_match.assign(std::max<size_t>(cells, 1), 0);
    check(sp_variant_match(ctx_, static_cast<int64_t>(seq_alleles.size()), static_cast<int64_t>(hap_alleles.size()), static_cast<int64_t>(nv),
                           s.data(), h.data(), vi.data(), vi_match.data(), all_match.data()),
          "sp_variant_match");
    vi_match.resize(cells); all_match.resize(cells);
}

PatternSet::~PatternSet() { sp_patterns_destroy(p_); }
DeviceMatrix::~DeviceMatrix() { sp_dmatrix_destroy(d_); }

std::shared_ptr<PatternSet> GpuAligner::prepare_patterns(const SeqList &patterns) {
    std::shared_ptr<PatternSet> ps(new PatternSet());
    ps->ascii_ = upload(patterns);
    Packed p(patterns);
    check(sp_patterns_create(ctx_, &p.set, SP_INFIX, &ps->p_), "sp_patterns_create");
    return ps;
}

std::unique_ptr<DeviceMatrix> GpuAligner::score_device(const SeqList &targets, const PatternSet &patterns) {
    const std::shared_ptr<ResidentSeqs> t = upload(targets);
    return score_device(*t, patterns);
}

std::unique_ptr<DeviceMatrix> GpuAligner::score_device(const ResidentSeqs &targets, const PatternSet &patterns, bool want_end_col) {
    std::unique_ptr<DeviceMatrix> m(new DeviceMatrix());
    check(sp_score_device(ctx_, targets.t_, patterns.p_, want_end_col ? 32 : 16, want_end_col ? 1 : 0, &m->d_), "sp_score_device");
    m->nt_ = static_cast<int64_t>(targets.size());
    m->np_ = static_cast<int64_t>(patterns.size());
    return m;
}

void GpuAligner::matrix_to_host(const DeviceMatrix &d, std::vector<int32_t> &D, std::vector<int32_t> *end_col) {
    const size_t n = std::max<size_t>(static_cast<size_t>(d.nt_) * static_cast<size_t>(d.np_), 1);
    D.assign(n, 0);
    if (end_col) end_col->assign(n, 0);
    check(sp_dmatrix_to_host(ctx_, d.d_, D.data(), end_col ? end_col->data() : nullptr), "sp_dmatrix_to_host");
    D.resize(static_cast<size_t>(d.nt_) * static_cast<size_t>(d.np_));
    if (end_col) end_col->resize(D.size());
}

std::vector<sp_pair_rec> GpuAligner::pair_minsum_topk(const DeviceMatrix &d, const DeviceMatrix *d2, int k) {
    std::vector<sp_pair_rec> out(static_cast<size_t>(std::max(k, 1)));
    int n = 0;
    check(sp_pair_minsum_topk(ctx_, d.d_, d2 ? d2->d_ : nullptr, 0, d.np_, k, out.data(), &n), "sp_pair_minsum_topk");
    out.resize(static_cast<size_t>(n));
    return out;
}

void GpuAligner::row_topk(const DeviceMatrix &d, int k, std::vector<int32_t> &idx, std::vector<int32_t> &dist,
                          const std::vector<int32_t> *pattern_bias, int dist_weight) {
    const size_t n = std::max<size_t>(static_cast<size_t>(d.nt_) * static_cast<size_t>(k), 1);
    idx.assign(n, -1); dist.assign(n, -1);
    if (pattern_bias && pattern_bias->size() != static_cast<size_t>(d.np_)) throw HostError("row_topk: one bias per pattern expected");
    check(sp_row_topk_weighted(ctx_, d.d_, dist_weight, pattern_bias ? pattern_bias->data() : nullptr, k, idx.data(), dist.data()), "sp_row_topk");
}

std::vector<uint64_t> GpuAligner::chain_pair_sums(const std::vector<std::vector<int32_t>> &chains,
                                                  const std::vector<std::vector<std::vector<uint32_t>>> &read_weights, int64_t n_haps) {
    std::vector<int32_t> coff(1, 0), items, soff(1, 0);
    for (const auto &c : chains) {
        items.insert(items.end(), c.begin(), c.end());
        coff.push_back(static_cast<int32_t>(items.size()));
    }
    std::vector<uint32_t> W;
    for (const auto &rw : read_weights) {
        for (const auto &seg : rw) {
            if (static_cast<int64_t>(seg.size()) != n_haps) throw HostError("chain_pair_sums: weight row has the wrong width");
            W.insert(W.end(), seg.begin(), seg.end());
        }
        soff.push_back(soff.back() + static_cast<int32_t>(rw.size()));
    }
    if (items.empty()) items.push_back(0);
    if (W.empty()) W.push_back(0);
    sp_dmatrix *B = nullptr;
    check(sp_chain_window_scores(ctx_, static_cast<int64_t>(chains.size()), coff.data(), items.data(), static_cast<int64_t>(read_weights.size()),
                                 soff.data(), W.data(), n_haps, &B),
          "sp_chain_window_scores");
    std::vector<uint64_t> S(std::max<size_t>(chains.size() * chains.size(), 1), 0);
    const sp_status st = sp_pair_minsum_full(ctx_, B, S.data());
    sp_dmatrix_destroy(B);
    check(st, "sp_pair_minsum_full");
    return S;
}

}  // namespace starphase
