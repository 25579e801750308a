// sp_host_consensus.cpp -- consensus search above K7 (sp_consensus_extend): row N1 of SURVEY.md 8f.
//
// Interface: waffle_con's ConsensusDWFA / DualConsensusDWFA as the reference drives them (src/hla/caller.rs:1124-1219:
// with_config, add_sequence, add_sequence_offset, consensus() -> list of equally good solutions, first one taken).
// Policy: a best-first search over consensus prefixes.  A node is one consensus (or two: a dual node) with, for every read,
// its edit distance to the prefix (end-free in the read) -- kept on the device, one track per consensus; reads vote for the
// next symbol with the bases that follow their best prefixes, one vote per read split evenly; a symbol is extended when it
// has min_count votes and a min_af share (the best-voted symbol when none qualifies); when two symbols qualify on a single
// node, a dual node with one consensus per symbol is opened as well; in a dual node a read counts towards, and votes for,
// the consensus it is closer to; at most max_capacity_per_size nodes are expanded per total length and the queue keeps the
// max_queue_size best nodes; a node nobody votes on is complete, and the cheapest complete nodes are the answer (a dual
// answer needs min_count reads and a min_af share on its minor side).  Every expansion is ONE device call covering all
// (consensus, symbol) extensions of the node.  This restates the published outline of waffle_con's search, not its code.
#include <algorithm>
#include <array>
#include <cmath>
#include <memory>
#include <set>

#include "starphase_host.hpp"

namespace starphase {

namespace {
constexpr int32_t kInf = 0x3FFFFFFF;
constexpr uint8_t kFinished = 1u << 5, kInactive = 1u << 6;
constexpr long kVoteUnit = 12;  // one read's vote, split evenly over its 1..4 candidate symbols

struct TrackPool {
    std::vector<int> free_list;
    explicit TrackPool(int n) { for (int t = n - 1; t >= 0; --t) free_list.push_back(t); }
    int take() {
        if (free_list.empty()) throw HostError("consensus: out of device tracks");
        const int t = free_list.back();
        free_list.pop_back();
        return t;
    }
};
struct Side {  // one consensus of a node: its device track and the per-read outputs of the step that made it
    int track = -1;
    TrackPool *pool = nullptr;
    std::vector<int32_t> ed, full;
    std::vector<uint8_t> votes;
    ~Side() { if (pool && track >= 0) pool->free_list.push_back(track); }
    std::vector<long> costs() const {
        std::vector<long> c(ed.size());
        for (size_t r = 0; r < ed.size(); ++r) c[r] = (votes[r] & kInactive) ? 0 : std::min(ed[r], full[r]);
        return c;
    }
};
using SidePtr = std::shared_ptr<Side>;
struct Node {
    long cost = 0;
    std::string c1, c2;
    bool dual = false;
    SidePtr s1, s2;
    size_t size() const { return c1.size() + (dual ? c2.size() : 0); }
};
struct NodeLess {  // (cost, longer first, consensus1, consensus2)
    bool operator()(const std::shared_ptr<Node> &a, const std::shared_ptr<Node> &b) const {
        if (a->cost != b->cost) return a->cost < b->cost;
        if (a->size() != b->size()) return a->size() > b->size();
        if (a->c1 != b->c1) return a->c1 < b->c1;
        return a->c2 < b->c2;
    }
};

std::array<long, 4> tally(const Side &s, const std::vector<char> *voters) {
    std::array<long, 4> t{0, 0, 0, 0};
    for (size_t r = 0; r < s.ed.size(); ++r) {
        if (voters && !(*voters)[r]) continue;
        if ((s.votes[r] & kInactive) || s.full[r] < s.ed[r]) continue;  // inactive, or already consumed at a better column
        const int n = __builtin_popcount(s.votes[r] & 15u);
        for (int k = 0; k < 4 && n; ++k)
            if (s.votes[r] >> k & 1u) t[static_cast<size_t>(k)] += kVoteUnit / n;
    }
    return t;
}
std::vector<int> strictly_passing(const std::array<long, 4> &t, const CdwfaConfig &cfg) {
    const long total = t[0] + t[1] + t[2] + t[3], permille = std::lround(cfg.min_af * 1000.0);
    std::vector<int> ok;
    for (int k = 0; k < 4; ++k)
        if (t[static_cast<size_t>(k)] >= kVoteUnit * static_cast<long>(cfg.min_count) && t[static_cast<size_t>(k)] * 1000 >= total * permille && total > 0)
            ok.push_back(k);
    return ok;
}
std::vector<int> passing(const std::array<long, 4> &t, const CdwfaConfig &cfg) {
    const long total = t[0] + t[1] + t[2] + t[3];
    if (total == 0) return {};
    std::vector<int> ok = strictly_passing(t, cfg);
    if (!ok.empty()) return ok;
    int best = 0;
    for (int k = 1; k < 4; ++k)
        if (t[static_cast<size_t>(k)] > t[static_cast<size_t>(best)]) best = k;
    return {best};
}

struct Search {
    GpuAligner &gpu;
    const CdwfaConfig &cfg;
    const SeqList &reads;
    sp_consensus *h = nullptr;
    TrackPool pool;
    size_t n_calls = 0;
    Search(GpuAligner &g, const CdwfaConfig &c, const SeqList &r, const std::vector<int32_t> &offsets)
        : gpu(g), cfg(c), reads(r), pool(256) {
        std::string bases;
        std::vector<int64_t> offs(1, 0);
        for (const auto &s : reads) { bases += s; offs.push_back(static_cast<int64_t>(bases.size())); }
        if (bases.empty()) bases.push_back('N');
        sp_seqset set{reinterpret_cast<const uint8_t *>(bases.data()), offs.data(), static_cast<int64_t>(reads.size())};
        const sp_status st = sp_consensus_create(gpu.raw(), &set, offsets.data(), static_cast<int32_t>(cfg.offset_window), static_cast<int32_t>(cfg.band),
                                                 256, &h);
        if (st != SP_OK) throw HostError(std::string("sp_consensus_create: ") + sp_last_error(gpu.raw()));
    }
    ~Search() { sp_consensus_destroy(h); }

    // one device call: every (parent side, symbol) of the list; symbol 0 = report the fresh track
    std::vector<SidePtr> extend(const std::vector<std::pair<const Side *, char>> &tasks) {
        const size_t n = tasks.size(), R = reads.size();
        std::vector<int32_t> src(n), dst(n), ed(n * R), full(n * R);
        std::vector<uint8_t> sym(n), votes(n * R);
        std::vector<SidePtr> out(n);
        for (size_t q = 0; q < n; ++q) {
            out[q] = std::make_shared<Side>();
            out[q]->pool = &pool;
            out[q]->track = pool.take();
            if (tasks[q].first) {
                src[q] = tasks[q].first->track;
            } else {  // root: a fresh track, reported
                src[q] = out[q]->track;
                if (sp_consensus_reset(h, out[q]->track) != SP_OK) throw HostError(std::string("sp_consensus_reset: ") + sp_last_error(gpu.raw()));
            }
            dst[q] = out[q]->track;
            sym[q] = static_cast<uint8_t>(tasks[q].second);
        }
        if (n && R) {
            if (sp_consensus_extend(h, static_cast<int32_t>(n), src.data(), sym.data(), dst.data(), ed.data(), votes.data(), full.data()) != SP_OK)
                throw HostError(std::string("sp_consensus_extend: ") + sp_last_error(gpu.raw()));
            ++n_calls;
        }
        for (size_t q = 0; q < n; ++q) {
            out[q]->ed.assign(ed.begin() + static_cast<std::ptrdiff_t>(q * R), ed.begin() + static_cast<std::ptrdiff_t>((q + 1) * R));
            out[q]->full.assign(full.begin() + static_cast<std::ptrdiff_t>(q * R), full.begin() + static_cast<std::ptrdiff_t>((q + 1) * R));
            out[q]->votes.assign(votes.begin() + static_cast<std::ptrdiff_t>(q * R), votes.begin() + static_cast<std::ptrdiff_t>((q + 1) * R));
        }
        return out;
    }

    std::shared_ptr<Node> make(std::string c1, SidePtr s1, bool dual, std::string c2, SidePtr s2) {
        auto n = std::make_shared<Node>();
        n->c1 = std::move(c1); n->s1 = std::move(s1); n->dual = dual; n->c2 = std::move(c2); n->s2 = std::move(s2);
        const std::vector<long> a = n->s1->costs();
        if (!dual) {
            for (long x : a) n->cost += x;
        } else {
            const std::vector<long> b = n->s2->costs();
            for (size_t r = 0; r < a.size(); ++r) n->cost += std::min(a[r], b[r]);
        }
        return n;
    }

    std::vector<std::shared_ptr<Node>> run(bool allow_dual) {
        static const char kSym[4] = {'A', 'C', 'G', 'T'};
        std::multiset<std::shared_ptr<Node>, NodeLess> queue;
        queue.insert(make("", extend({{nullptr, 0}})[0], false, "", nullptr));
        std::vector<std::shared_ptr<Node>> best;
        bool have_best = false;
        long best_cost = 0;
        std::map<size_t, size_t> expanded;
        const long permille = std::lround(cfg.min_af * 1000.0);
        while (!queue.empty()) {
            const std::shared_ptr<Node> node = *queue.begin();
            queue.erase(queue.begin());
            if (have_best && node->cost > best_cost) break;
            std::vector<int> p1, p2, strict;
            std::vector<long> c1, c2;
            if (!node->dual) {
                const auto t1 = tally(*node->s1, nullptr);
                p1 = passing(t1, cfg);
                strict = strictly_passing(t1, cfg);
            } else {
                c1 = node->s1->costs(); c2 = node->s2->costs();
                std::vector<char> v1(c1.size()), v2(c1.size());
                for (size_t r = 0; r < c1.size(); ++r) { v1[r] = c1[r] <= c2[r]; v2[r] = c2[r] <= c1[r]; }
                p1 = passing(tally(*node->s1, &v1), cfg);
                p2 = passing(tally(*node->s2, &v2), cfg);
            }
            if (p1.empty() && p2.empty()) {  // nobody wants to go on: complete
                if (node->dual) {
                    size_t n1 = 0;
                    for (size_t r = 0; r < c1.size(); ++r) n1 += c1[r] <= c2[r];
                    const size_t n2 = c1.size() - n1, mn = std::min(n1, n2);
                    if (mn < cfg.min_count || static_cast<long>(mn) * 1000 < static_cast<long>(n1 + n2) * permille) continue;
                }
                if (!have_best || node->cost < best_cost) { best.assign(1, node); best_cost = node->cost; have_best = true; }
                else if (node->cost == best_cost) best.push_back(node);
                continue;
            }
            size_t &cnt = expanded[node->size()];
            if (cnt >= cfg.max_capacity_per_size) continue;
            ++cnt;
            std::vector<std::pair<const Side *, char>> tasks;
            for (int k : p1) tasks.emplace_back(node->s1.get(), kSym[k]);
            for (int k : p2) tasks.emplace_back(node->s2.get(), kSym[k]);
            const std::vector<SidePtr> ext = extend(tasks);
            if (!node->dual) {
                for (size_t a = 0; a < p1.size(); ++a) queue.insert(make(node->c1 + kSym[p1[a]], ext[a], false, "", nullptr));
                if (allow_dual && strict.size() >= 2)
                    for (size_t ia = 0; ia < strict.size(); ++ia)
                        for (size_t ib = ia + 1; ib < strict.size(); ++ib) {
                            const size_t a = static_cast<size_t>(std::find(p1.begin(), p1.end(), strict[ia]) - p1.begin());
                            const size_t b = static_cast<size_t>(std::find(p1.begin(), p1.end(), strict[ib]) - p1.begin());
                            queue.insert(make(node->c1 + kSym[strict[ia]], ext[a], true, node->c1 + kSym[strict[ib]], ext[b]));
                        }
            } else {
                // a finished side stays as it is
                const size_t n1 = std::max<size_t>(p1.size(), 1), n2 = std::max<size_t>(p2.size(), 1);
                for (size_t a = 0; a < n1; ++a)
                    for (size_t b = 0; b < n2; ++b) {
                        const bool e1 = !p1.empty(), e2 = !p2.empty();
                        queue.insert(make(e1 ? node->c1 + kSym[p1[a]] : node->c1, e1 ? ext[a] : node->s1, true,
                                          e2 ? node->c2 + kSym[p2[b]] : node->c2, e2 ? ext[p1.size() + b] : node->s2));
                    }
            }
            while (queue.size() > cfg.max_queue_size) queue.erase(std::prev(queue.end()));
        }
        return best;
    }
};
}  // namespace

void ConsensusDWFA::add_sequence_offset(const std::string &sequence, std::optional<size_t> offset) {
    reads_.push_back(sequence);
    offsets_.push_back(offset ? static_cast<int32_t>(*offset) : -1);  // -1: anchored at the consensus start
}

std::vector<Consensus> ConsensusDWFA::consensus() {
    if (reads_.empty()) throw HostError("consensus: no sequences were added");
    Search s(gpu_, config_, reads_, offsets_);
    std::vector<Consensus> out;
    for (const auto &n : s.run(false)) {
        Consensus c;
        c.sequence = n->c1;
        for (long x : n->s1->costs()) c.scores.push_back(static_cast<size_t>(x));
        out.push_back(std::move(c));
    }
    n_calls_ = s.n_calls;
    return out;
}

std::vector<DualConsensus> DualConsensusDWFA::consensus() {
    if (inner_.reads_.empty()) throw HostError("consensus: no sequences were added");
    Search s(inner_.gpu_, inner_.config_, inner_.reads_, inner_.offsets_);
    std::vector<DualConsensus> out;
    for (const auto &n : s.run(true)) {
        DualConsensus d;
        d.consensus1 = n->c1;
        const std::vector<long> a = n->s1->costs();
        if (!n->dual) {
            d.is_consensus1.assign(a.size(), true);
            for (long x : a) d.scores1.emplace_back(static_cast<size_t>(x));
            d.scores2.assign(a.size(), std::nullopt);
        } else {
            d.consensus2 = n->c2;
            const std::vector<long> b = n->s2->costs();
            for (size_t r = 0; r < a.size(); ++r) {
                d.is_consensus1.push_back(a[r] <= b[r]);
                d.scores1.emplace_back(static_cast<size_t>(a[r]));
                d.scores2.emplace_back(static_cast<size_t>(b[r]));
            }
        }
        out.push_back(std::move(d));
    }
    inner_.n_calls_ = s.n_calls;
    return out;
}

}  // namespace starphase
