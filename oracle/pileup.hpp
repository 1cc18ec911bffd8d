// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called from the product path.
//
// CPU restatement of the per-interval loop of `modkit pileup` (reference v0.4.4):
//   process_region          src/pileup/mod.rs:718-1020
//   ReadCache::add_record   src/read_cache.rs:111-211
//   get_mod_call            src/read_cache.rs:232-297
//   add_feature / decode    src/pileup/mod.rs:238-445
//   combine_strand_features src/pileup/mod.rs:469-561
//   aligned pairs           src/util.rs:122-145 (+ rust-htslib aligned_pairs: M/=/X pair, I/S query, D/N ref)
//   htslib pileup admission flag mask UNMAP|SECONDARY|QCFAIL|DUP (htslib bam_plp_init default) plus
//   supplementary/seq_len==0 (src/pileup/mod.rs:783-791). --max-depth (src/pileup/mod.rs:755-759 ->
//   htslib bam_plp_set_maxcnt) is restated from htslib's bam_plp_push in plp_maxcnt_filter below; no reference test
//   exercises it (parity unpinned).
// The reference iterates htslib pileup columns; this restatement walks each read's CIGAR once and
// scatters into dense per-interval arrays. The per-(read,position) feature rules are identical.
#pragma once
#include <functional>
#include <mutex>
#include <set>

#include "bam.hpp"
#include "fasta.hpp"
#include "modbam.hpp"

namespace orc {

// (primary base, mod code) -> small id, shared across intervals/threads
struct StateTable {
    std::mutex mu;
    std::vector<std::pair<int, ModCode>> states;
    int id_of(int pb, ModCode c) {
        // thread-local front cache: the shared table is append-only, so a cached id never goes stale
        static thread_local std::vector<std::pair<std::pair<int, ModCode>, int>> cache;
        static thread_local const StateTable* owner = nullptr;
        if (owner != this) { cache.clear(); owner = this; }
        for (auto& e : cache) if (e.first.first == pb && e.first.second == c) return e.second;
        int id = id_of_locked(pb, c);
        cache.push_back({{pb, c}, id});
        return id;
    }
    int id_of_locked(int pb, ModCode c) {
        std::lock_guard<std::mutex> g(mu);
        for (size_t i = 0; i < states.size(); i++) if (states[i].first == pb && states[i].second == c) return (int)i;
        states.push_back({pb, c});
        return (int)states.size() - 1;
    }
    size_t size() { std::lock_guard<std::mutex> g(mu); return states.size(); }
    std::pair<int, ModCode> get(int id) { std::lock_guard<std::mutex> g(mu); return states[id]; }
};

enum NumericMode { PASSTHROUGH = 0, COMBINE = 1, COLLAPSE = 2 };

struct PileupParams {
    Caller caller;
    int numeric = PASSTHROUGH;
    ModCode collapse_code = 0;
    bool force_allow_implicit = false;
    bool combine_strands = false;
    EdgeFilter edge;
    uint32_t max_depth = 8000;     // 0 = no limit
};

struct Row {  // PileupFeatureCounts
    uint32_t pos;
    char strand;
    ModCode code;
    int motif_idx;  // -1 none
    uint32_t cov, n_mod, n_canon, n_other, n_delete, n_filtered, n_diff, n_nocall;
};

// FocusPositions for one interval (src/interval_chunks.rs:32-59)
struct Focus {
    bool all = true;
    std::map<uint32_t, uint8_t> rule;                      // position -> StrandRule bits
    std::map<uint32_t, std::vector<int>> pos_ids, neg_ids;  // motif ids per strand
    // combine-strands: position -> [(motif index)] for + motifs (BTreeMap order)
    bool combine = false;
};

struct Interval {
    uint32_t tid, start, end;
    Focus focus;
};

// Per-read decode result used by the interval loop
struct ReadCalls {
    bool skipped = true;
    // per query position (BAM order): 0 none, 1 filtered, 2 canonical, 3+state id
    std::vector<uint16_t> plus, minus;
    std::vector<int> pos_codes, neg_codes;  // state ids observed, by reference strand
};

inline std::string forward_sequence(const BamRecord& r) {
    int L = r.l_seq();
    std::string s((size_t)L, 'N');
    if (!r.is_reverse()) { for (int i = 0; i < L; i++) s[i] = r.base_at(i); return s; }
    // bio::alphabets::dna::revcomp (IUPAC aware)
    auto comp = [](char c) -> char {
        switch (c) { case 'A': return 'T'; case 'C': return 'G'; case 'G': return 'C'; case 'T': return 'A';
                     case 'M': return 'K'; case 'K': return 'M'; case 'R': return 'Y'; case 'Y': return 'R';
                     case 'V': return 'B'; case 'B': return 'V'; case 'H': return 'D'; case 'D': return 'H';
                     default: return c; }
    };
    for (int i = 0; i < L; i++) s[L - 1 - i] = comp(r.base_at(i));
    return s;
}

// MM/Mm + ML/Ml + MN lookup (src/mod_bam.rs:1388-1470, src/util.rs:174-188)
inline bool raw_mod_tags(const BamRecord& r, AuxField* mm, AuxField* ml) {
    if (!find_aux(r, "MM", mm) && !find_aux(r, "Mm", mm)) return false;
    if (mm->type != 'Z') return false;
    if (!find_aux(r, "ML", ml) && !find_aux(r, "Ml", ml)) return false;
    if (!(ml->type == 'B' && ml->subtype == 'C')) return false;
    AuxField mn;
    if (find_aux(r, "MN", &mn)) {
        int64_t v;
        switch (mn.type) {
            case 'c': v = (int8_t)mn.p[0]; break; case 'C': v = mn.p[0]; break;
            case 's': v = (int16_t)rd16(mn.p); break; case 'S': v = rd16(mn.p); break;
            case 'i': v = rdi32(mn.p); break; case 'I': v = rd32(mn.p); break;
            default: return false;
        }
        if ((size_t)v != (size_t)r.l_seq()) return false;
    }
    return true;
}

inline bool decode_mod_base_info(const BamRecord& r, ModBaseInfo* info, std::string* fwd) {
    AuxField mm, ml;
    if (!raw_mod_tags(r, &mm, &ml)) return false;
    *fwd = forward_sequence(r);
    std::vector<MmList> lists;
    if (!parse_mm((const char*)mm.p, mm.n, &lists)) return false;
    return build_mod_base_info(lists, ml.p, ml.n, *fwd, info);
}

// query position (BAM order) -> reference position, -1 when unaligned
inline void aligned_ref_positions(const BamRecord& r, std::vector<int64_t>* q2r) {
    q2r->assign((size_t)r.l_seq(), -1);
    int64_t rp = r.pos();
    int q = 0;
    for (int i = 0; i < r.n_cigar(); i++) {
        uint32_t c = r.cigar_op(i);
        int op = c & 0xf, len = (int)(c >> 4);
        switch (op) {
            case 0: case 7: case 8: for (int k = 0; k < len && q < r.l_seq(); k++) (*q2r)[q++] = rp++; rp += 0; break;
            case 1: case 4: q += len; break;
            case 2: case 3: rp += len; break;
            default: break;
        }
    }
}

// ReadCache::add_record for the pileup path
inline void decode_read_for_pileup(const BamRecord& r, const PileupParams& P, StateTable& st, ReadCalls* out) {
    out->skipped = true;
    out->plus.clear(); out->minus.clear(); out->pos_codes.clear(); out->neg_codes.clear();
    ModBaseInfo info;
    std::string fwd;
    if (!decode_mod_base_info(r, &info, &fwd)) return;
    if (info.is_empty()) return;
    for (int s = 0; s < 2; s++) for (int b = 0; b < 4; b++)
        if (info.present[s][b] && info.tab[s][b].mode == DEFAULT_IMPLICIT && !P.force_allow_implicit) return;
    int L = r.l_seq();
    std::vector<int64_t> q2r;
    aligned_ref_positions(r, &q2r);
    out->plus.assign((size_t)L, 0);
    out->minus.assign((size_t)L, 0);
    bool added = false;
    bool rev = r.is_reverse();
    for (int s = 0; s < 2; s++) for (int b = 0; b < 4; b++) {
        if (!info.present[s][b]) continue;
        SeqPosTable& T = info.tab[s][b];
        int tb = s == 0 ? b : comp_idx(b);
        if (P.edge.on) {
            if (!P.edge.read_can_be_trimmed((size_t)L)) continue;
            for (auto it = T.pos.begin(); it != T.pos.end();) { if (!P.edge.keep(it->first, (size_t)L)) it = T.pos.erase(it); else ++it; }
            if (T.pos.empty()) continue;
        }
        std::set<int> codes;
        std::vector<uint16_t>& dst = s == 0 ? out->plus : out->minus;
        for (auto& kv : T.pos) {
            BaseModProbs bmp = P.numeric == COLLAPSE ? redistribute(kv.second, P.collapse_code) : kv.second;
            bmp.probs.for_each([&](ModCode c, float) { codes.insert(st.id_of(tb, c)); });
            uint32_t f = kv.first;
            int q = rev ? L - 1 - (int)f : (int)f;
            if (q < 0 || q >= L || q2r[q] < 0) continue;
            Call call = make_call(P.caller, tb, bmp);
            uint16_t v = call.kind == CALL_FILTERED ? 1 : call.kind == CALL_CANONICAL ? 2 : (uint16_t)(3 + st.id_of(tb, call.code));
            dst[q] = v;
        }
        // (mod strand, read orientation) -> reference strand (read_cache.rs:181-188)
        bool to_pos = (s == 0) != rev;
        std::vector<int>& cs = to_pos ? out->pos_codes : out->neg_codes;
        for (int id : codes) if (std::find(cs.begin(), cs.end(), id) == cs.end()) cs.push_back(id);
        added = true;
    }
    out->skipped = !added;
    if (out->skipped) { out->plus.clear(); out->minus.clear(); out->pos_codes.clear(); out->neg_codes.clear(); }
}

struct StrandTally {
    uint32_t n_delete = 0, n_filtered = 0;
    uint32_t basecall[4] = {0, 0, 0, 0};
    uint32_t canon[4] = {0, 0, 0, 0};
};

inline bool admitted_for_pileup(const BamRecord& r) {
    uint16_t f = r.flag();
    if (f & (0x4 | 0x100 | 0x200 | 0x400 | 0x800)) return false;
    if (r.l_seq() == 0) return false;
    return true;
}

// bam_plp_push (htslib sam.c) with bam_plp_set_maxcnt: `iter->tid == b->core.tid && iter->pos == b->core.pos &&
// iter->mp->cnt > iter->maxcnt` => the read is not buffered. iter->pos only equals a read's start for the second and later
// reads of one start position (the first one arrives while the engine still stands on an earlier column); mp->cnt is the
// number of buffered reads plus the list's sentinel node; a buffered read is released when the column at its end position
// is built, i.e. every read whose end lies before the current start is gone.
inline void plp_maxcnt_filter(std::vector<const BamRecord*>* recs, uint32_t maxcnt) {
    if (!maxcnt || recs->size() <= maxcnt) return;
    std::multiset<int64_t> ends;
    int64_t cur = INT64_MIN;
    std::vector<const BamRecord*> out;
    for (const BamRecord* r : *recs) {
        const int64_t s = r->pos();
        int64_t rlen = 0;
        for (int ci = 0; ci < r->n_cigar(); ci++) { const uint32_t c = r->cigar_op(ci); const int op = c & 0xf; if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rlen += c >> 4; }
        const int64_t e = s + (rlen ? rlen : 1);
        if (s != cur) { cur = s; ends.erase(ends.begin(), ends.lower_bound(s)); }
        else if (ends.size() + 1 > (size_t)maxcnt) continue;
        ends.insert(e);
        out.push_back(r);
    }
    recs->swap(out);
}

// process_region: rows for one interval, sorted by position then (strand, code)
inline void process_interval(const BamFile& bam, const Interval& iv, const PileupParams& P, StateTable& st,
                             const std::vector<Motif>* motifs, std::vector<Row>* rows_out,
                             size_t* n_processed = nullptr, size_t* n_skipped = nullptr,
                             const std::function<bool(const BamRecord&)>* keep = nullptr) {
    const uint32_t start = iv.start, end = iv.end;
    if (end <= start) return;
    const size_t W = end - start;
    std::vector<const BamRecord*> recs;
    // `keep`: the reads of one partition (--partition-tag): every partition key is an independent pileup
    // (tallies and observed-code sets are keyed by PartitionKey, src/pileup/mod.rs:767-830, 942-965)
    // htslib's pileup engine sees every fetched record (one engine per interval, src/pileup/mod.rs:732-759): bam_plp_push drops
    // the reads of its flag mask, then - maxcnt - a read whose start equals the engine's current column while the buffer is
    // full; what is left reaches the column loop, where modkit's own filter and the partition key apply.
    {
        std::vector<const BamRecord*> pushed;
        bam.fetch(iv.tid, start, end, [&](const BamRecord& r) { if (!(r.flag() & (0x4 | 0x100 | 0x200 | 0x400))) pushed.push_back(&r); });
        plp_maxcnt_filter(&pushed, P.max_depth);
        for (const BamRecord* r : pushed) if (admitted_for_pileup(*r) && (!keep || (*keep)(*r))) recs.push_back(r);
    }
    if (recs.empty()) return;
    static thread_local std::vector<ReadCalls> calls;
    if (calls.size() < recs.size()) calls.resize(recs.size());
    for (size_t i = 0; i < recs.size(); i++) decode_read_for_pileup(*recs[i], P, st, &calls[i]);
    if (n_processed) for (size_t i = 0; i < recs.size(); i++) { if (calls[i].skipped) { if (n_skipped) (*n_skipped)++; } else (*n_processed)++; }
    const size_t NS = st.size();
    std::vector<std::pair<int, ModCode>> states(NS);
    for (size_t i = 0; i < NS; i++) states[i] = st.get((int)i);

    // focus rule per position
    std::vector<uint8_t> rule(W, iv.focus.all ? 3 : 0);
    if (!iv.focus.all)
        for (auto& kv : iv.focus.rule) if (kv.first >= start && kv.first < end) rule[kv.first - start] = kv.second;

    // per-thread scratch, reused across intervals (allocation + page-fault cost would otherwise dominate)
    static thread_local std::vector<StrandTally> tally;
    static thread_local std::vector<uint32_t> mods;
    static thread_local std::vector<int32_t> obs;     // difference arrays [strand][state][pos]
    tally.assign(2 * W, StrandTally());
    mods.assign(2 * W * std::max<size_t>(NS, 1), 0);
    obs.assign(2 * (W + 1) * std::max<size_t>(NS, 1), 0);
    auto OBS = [&](int s, int id, size_t x) -> int32_t& { return obs[((size_t)s * NS + id) * (W + 1) + x]; };
    auto MODS = [&](int s, size_t x, int id) -> uint32_t& { return mods[((size_t)s * W + x) * NS + id]; };

    for (size_t i = 0; i < recs.size(); i++) {
        const BamRecord& r = *recs[i];
        const ReadCalls& rc = calls[i];
        const int a = r.is_reverse() ? 1 : 0;
        int64_t rp = r.pos();
        int q = 0;
        const int L = r.l_seq();
        auto cover = [&](int64_t b, int64_t e) {  // observed-code coverage over [b,e)
            if (rc.skipped) return;
            int64_t lo = std::max<int64_t>(b, start), hi = std::min<int64_t>(e, end);
            if (lo >= hi) return;
            for (int id : rc.pos_codes) { OBS(0, id, lo - start)++; OBS(0, id, hi - start)--; }
            for (int id : rc.neg_codes) { OBS(1, id, lo - start)++; OBS(1, id, hi - start)--; }
        };
        for (int ci = 0; ci < r.n_cigar(); ci++) {
            uint32_t c = r.cigar_op(ci);
            int op = c & 0xf, len = (int)(c >> 4);
            if (op == 0 || op == 7 || op == 8) {
                cover(rp, rp + len);
                for (int k = 0; k < len; k++, q++, rp++) {
                    if (rp < start || rp >= end || q >= L) continue;
                    size_t x = (size_t)(rp - start);
                    uint8_t ru = rule[x];
                    if (!ru) continue;
                    int b = base_idx(r.base_at(q));
                    if (b < 0) continue;
                    if (a) b = comp_idx(b);
                    uint16_t pc = rc.skipped ? 0 : rc.plus[q], nc = rc.skipped ? 0 : rc.minus[q];
                    auto add = [&](int ts, uint16_t v, int pb) {
                        if (!(ru & (ts == 0 ? 1 : 2))) return;
                        StrandTally& T = tally[(size_t)ts * W + x];
                        if (v == 1) T.n_filtered++;
                        else if (v == 2) T.canon[pb]++;
                        else MODS(ts, x, v - 3)++;
                    };
                    if (!pc && !nc) { if (ru & (a == 0 ? 1 : 2)) tally[(size_t)a * W + x].basecall[b]++; }
                    if (pc) add(a, pc, b);
                    if (nc) add(1 - a, nc, comp_idx(b));
                }
            } else if (op == 1 || op == 4) {
                q += len;
            } else if (op == 2) {
                cover(rp, rp + len);
                for (int k = 0; k < len; k++, rp++) {
                    if (rp < start || rp >= end) continue;
                    size_t x = (size_t)(rp - start);
                    uint8_t ru = rule[x];
                    if (ru & (a == 0 ? 1 : 2)) tally[(size_t)a * W + x].n_delete++;
                }
            } else if (op == 3) {
                rp += len;
            }
        }
    }
    // prefix sums of the coverage difference arrays
    for (int s = 0; s < 2; s++) for (size_t id = 0; id < NS; id++) {
        int32_t run = 0;
        for (size_t x = 0; x < W; x++) { run += OBS(s, (int)id, x); OBS(s, (int)id, x) = run; }
    }

    // decode (pileup/mod.rs:283-445)
    std::vector<Row> rows;
    for (size_t x = 0; x < W; x++) {
        if (!rule[x]) continue;
        uint32_t pos = start + (uint32_t)x;
        size_t row0 = rows.size();
        for (int s = 0; s < 2; s++) {
            const StrandTally& T = tally[(size_t)s * W + x];
            const std::vector<int>* mids = nullptr;
            if (!iv.focus.all) {
                const auto& m = s == 0 ? iv.focus.pos_ids : iv.focus.neg_ids;
                auto it = m.find(pos);
                if (it != m.end()) mids = &it->second;
            }
            uint32_t mod_by_base[4] = {0, 0, 0, 0};
            for (size_t id = 0; id < NS; id++) mod_by_base[states[id].first] += MODS(s, x, (int)id);
            for (int pb = 0; pb < 4; pb++) {
                uint32_t n_can = T.canon[pb], total_mod = mod_by_base[pb];
                if (n_can + total_mod == 0) continue;  // primary base absent from modcall_counts
                uint32_t cov = n_can + total_mod;
                uint32_t n_diff = 0;
                for (int b = 0; b < 4; b++) if (b != pb) n_diff += T.basecall[b] + T.canon[b] + mod_by_base[b];
                uint32_t n_nocall = T.basecall[pb];
                auto push = [&](ModCode code, uint32_t n_mod, uint32_t n_other) {
                    Row R{pos, s == 0 ? '+' : '-', code, -1, cov, n_mod, n_can, n_other, T.n_delete, T.n_filtered, n_diff, n_nocall};
                    if (mids) for (int idx : *mids) { R.motif_idx = idx; rows.push_back(R); }
                    else rows.push_back(R);
                };
                if (P.numeric == COMBINE) {
                    push((ModCode)(uint8_t)BASES[pb], total_mod, 0);
                } else {
                    for (size_t id = 0; id < NS; id++) {
                        if (states[id].first != pb || OBS(s, (int)id, x) <= 0) continue;
                        uint32_t n_mod = MODS(s, x, (int)id);
                        push(states[id].second, n_mod, total_mod - n_mod);
                    }
                }
            }
        }
        std::stable_sort(rows.begin() + row0, rows.end(), [](const Row& A, const Row& B) {
            if (A.strand != B.strand) return A.strand < B.strand;  // '+' (43) < '-' (45)
            return A.code < B.code;
        });
    }

    if (P.combine_strands && !iv.focus.all && motifs) {
        // combine_strand_features (pileup/mod.rs:469-561)
        std::multimap<uint32_t, size_t> by_pos;
        for (size_t i = 0; i < rows.size(); i++) by_pos.emplace(rows[i].pos, i);
        std::vector<Row> combined;
        for (auto& kv : iv.focus.pos_ids) {  // BTreeMap order over + motif positions
            uint32_t p = kv.first;
            for (int idx : kv.second) {
                const Motif& m = (*motifs)[idx];
                if (!m.palindrome) continue;
                int64_t partner = (int64_t)p + (m.rev_off - m.fwd_off);
                if (partner < 0) continue;
                std::map<ModCode, Row> grouped;
                auto take = [&](uint32_t at, char strand) {
                    auto rg = by_pos.equal_range(at);
                    for (auto it = rg.first; it != rg.second; ++it) {
                        const Row& R = rows[it->second];
                        if (R.strand != strand || R.motif_idx != idx) continue;
                        auto g = grouped.find(R.code);
                        if (g == grouped.end()) { Row Z{p, '.', R.code, idx, 0, 0, 0, 0, 0, 0, 0, 0}; g = grouped.emplace(R.code, Z).first; }
                        Row& A = g->second;
                        A.cov += R.cov; A.n_mod += R.n_mod; A.n_canon += R.n_canon; A.n_other += R.n_other;
                        A.n_delete += R.n_delete; A.n_filtered += R.n_filtered; A.n_diff += R.n_diff; A.n_nocall += R.n_nocall;
                    }
                };
                take(p, '+');
                take((uint32_t)partner, '-');
                for (auto& g : grouped) combined.push_back(g.second);
            }
        }
        rows.swap(combined);
    }
    rows_out->insert(rows_out->end(), rows.begin(), rows.end());
}

}  // namespace orc
