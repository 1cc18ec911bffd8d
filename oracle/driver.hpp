// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called from the product path.
//
// CPU restatement of the orchestration around the hot loop of `modkit pileup` (reference v0.4.4):
//   interval feeder        src/interval_chunks.rs:563-643
//   focus positions        src/interval_chunks.rs:61-297, src/fasta.rs:92-226
//   threshold estimation   src/command_utils.rs:74-134, src/thresholds.rs:17-156,
//                          src/reads_sampler/mod.rs:30-376, src/reads_sampler/sampling_schedule.rs:171-615,
//                          src/reads_sampler/record_sampler.rs, src/read_ids_to_base_mod_probs.rs:223-362,965-1069
//   bedMethyl writer       src/writers.rs:43-183
#pragma once
#include <set>
#include <charconv>
#include <cmath>
#include <charconv>
#include <cmath>
#include <functional>
#include <unordered_set>

#include "pileup.hpp"

namespace orc {

struct RefRecord { uint32_t tid, start, length; std::string name; uint32_t end() const { return start + length; } };

struct Region { std::string name; uint32_t start, end; };

// StrandedPositionFilter (src/position_filter.rs:20-365): per contig, merged [start,stop) intervals per strand.
// rust-lapper merge_overlaps joins intervals that overlap or touch.
struct PositionFilter {
    typedef std::vector<std::pair<uint64_t, uint64_t>> Ivs;
    std::map<uint32_t, Ivs> pos, neg;
    static void merge(Ivs* v) {
        std::sort(v->begin(), v->end());
        Ivs out;
        for (auto& iv : *v) { if (!out.empty() && !(out.back().second < iv.first)) out.back().second = std::max(out.back().second, iv.second); else out.push_back(iv); }
        v->swap(out);
    }
    static bool hit(const Ivs& v, uint64_t a, uint64_t b) {   // any interval with start < b && stop > a
        auto it = std::lower_bound(v.begin(), v.end(), std::make_pair(b, (uint64_t)0));
        while (it != v.begin()) { --it; if (it->second > a) return it->first < b; if (it->second <= a) break; }
        return false;
    }
    void load(const std::string& path, const std::map<std::string, uint32_t>& chrom_to_tid) {   // from_bed_file :244-345
        std::ifstream f(path);
        if (!f) throw std::runtime_error("failed to open BED " + path);
        std::string line;
        std::set<std::string> warned;
        while (std::getline(f, line)) {
            if (line.empty()) continue;
            std::istringstream ss(line);
            std::vector<std::string> parts;
            std::string tok;
            while (ss >> tok) parts.push_back(tok);
            if (parts.size() < 3 || warned.count(parts[0])) continue;
            uint64_t a, b;
            try { size_t k; a = std::stoull(parts[1], &k); if (k != parts[1].size()) continue; b = std::stoull(parts[2], &k); if (k != parts[2].size()) continue; } catch (...) { continue; }
            bool p, n;
            if (parts.size() == 3) p = n = true;
            else if (parts.size() >= 6) { if (parts[5] == "+") { p = true; n = false; } else if (parts[5] == "-") { p = false; n = true; } else if (parts[5] == ".") p = n = true; else continue; }
            else continue;
            auto it = chrom_to_tid.find(parts[0]);
            if (it == chrom_to_tid.end()) { warned.insert(parts[0]); continue; }
            if (p) pos[it->second].push_back({a, b});
            if (n) neg[it->second].push_back({a, b});
        }
        if (pos.empty() && neg.empty()) throw std::runtime_error("zero valid positions parsed from BED file");
        for (auto& kv : pos) merge(&kv.second);
        for (auto& kv : neg) merge(&kv.second);
    }
    bool contains(uint32_t tid, uint64_t p, bool negative) const {
        const auto& m = negative ? neg : pos;
        auto it = m.find(tid);
        return it != m.end() && hit(it->second, p, p + 1);
    }
    bool overlaps(uint32_t tid, uint64_t a, uint64_t b) const {
        auto it = pos.find(tid);
        if (it != pos.end() && hit(it->second, a, b)) return true;
        it = neg.find(tid);
        return it != neg.end() && hit(it->second, a, b);
    }
    bool has_contig(uint32_t tid) const { return pos.count(tid) || neg.count(tid); }
};

inline Region parse_region(const std::string& raw, const BamFile& bam) {  // util.rs:475-523
    Region r;
    auto c = raw.find(':');
    if (c == std::string::npos) {
        int tid = bam.tid_of(raw);
        if (tid < 0) throw std::runtime_error("contig missing from header: " + raw);
        return Region{raw, 0, bam.ref_lens[tid]};
    }
    r.name = raw.substr(0, c);
    std::string se = raw.substr(c + 1);
    if (se.find(':') != std::string::npos) throw std::runtime_error("invalid region " + raw);
    auto d = se.find('-');
    if (d == std::string::npos || se.find('-', d + 1) != std::string::npos) throw std::runtime_error("invalid region " + raw);
    auto num = [&](std::string s) -> uint32_t {
        s.erase(std::remove(s.begin(), s.end(), ','), s.end());
        if (s.empty()) throw std::runtime_error("invalid region " + raw);
        uint64_t v = 0;
        for (char ch : s) { if (ch < '0' || ch > '9') throw std::runtime_error("invalid region " + raw); v = v * 10 + (ch - '0'); }
        return (uint32_t)v;
    };
    r.start = num(se.substr(0, d));
    r.end = num(se.substr(d + 1));
    if (r.end <= r.start) throw std::runtime_error("invalid region " + raw);
    return r;
}

inline std::vector<RefRecord> get_targets(const BamFile& bam, const Region* region) {  // util.rs:409-446
    std::vector<RefRecord> out;
    for (uint32_t tid = 0; tid < bam.ref_names.size(); tid++) {
        if (region) { if (bam.ref_names[tid] == region->name) out.push_back({tid, region->start, region->end - region->start, region->name}); }
        else out.push_back({tid, 0, bam.ref_lens[tid], bam.ref_names[tid]});
    }
    return out;
}

// optimize_reference_records / group_genome_intervals (src/position_filter.rs:106-212)
inline std::vector<RefRecord> optimize_reference_records(const PositionFilter& pf, const std::vector<RefRecord>& recs, uint32_t interval_size) {
    std::map<uint32_t, RefRecord> lut;
    for (auto& r : recs) lut[r.tid] = r;
    std::set<uint32_t> tids;
    for (auto& kv : pf.pos) tids.insert(kv.first);
    for (auto& kv : pf.neg) tids.insert(kv.first);
    std::vector<RefRecord> out;
    for (uint32_t tid : tids) {
        auto lr = lut.find(tid);
        if (lr == lut.end()) continue;
        PositionFilter::Ivs all;
        auto a = pf.pos.find(tid); if (a != pf.pos.end()) all.insert(all.end(), a->second.begin(), a->second.end());
        auto b = pf.neg.find(tid); if (b != pf.neg.end()) all.insert(all.end(), b->second.begin(), b->second.end());
        PositionFilter::merge(&all);
        if (all.empty()) continue;
        std::vector<std::pair<uint64_t, uint64_t>> agg;
        auto cur = all[0];
        for (size_t i = 1; i < all.size(); i++) {
            if (cur.second - cur.first > interval_size) { agg.push_back(cur); cur = all[i]; continue; }
            cur.second = all[i].second;
        }
        agg.push_back(cur);
        for (auto& g : agg) out.push_back({tid, (uint32_t)g.first, (uint32_t)(g.second - g.first), lr->second.name});
    }
    return out;
}

// FocusPositions::new_regions + check_position (src/interval_chunks.rs:299-371)
inline void focus_from_regions(const PositionFilter& pf, uint32_t tid, uint32_t start, uint32_t end, Focus* f) {
    f->all = false;
    auto add = [&](const std::map<uint32_t, PositionFilter::Ivs>& m, uint8_t bit) {
        auto it = m.find(tid);
        if (it == m.end()) return;
        for (auto& iv : it->second) {
            uint64_t a = std::max<uint64_t>(iv.first, start), b = std::min<uint64_t>(iv.second, end);
            for (uint64_t p = a; p < b; p++) f->rule[(uint32_t)p] |= bit;
        }
    };
    add(pf.pos, 1);
    add(pf.neg, 2);
}

// --- focus positions ------------------------------------------------------------------------
inline void focus_from_locs(const std::vector<MotifLocs>& locs, const std::vector<Motif>& motifs,
                            uint32_t start, uint32_t end, bool combine, Focus* f) {
    f->all = false;
    f->combine = combine;
    auto in = [&](uint32_t p) { return p >= start && p < end; };
    auto merge_rule = [&](uint32_t p, uint8_t r) { auto it = f->rule.find(p); if (it == f->rule.end()) f->rule[p] = r; else if (it->second != r) it->second = 3; };
    if (combine) {  // new_motif_combine_strands
        for (size_t id = 0; id < locs.size(); id++)
            for (auto& kv : locs[id]) {
                if (!in(kv.first)) continue;
                merge_rule(kv.first, kv.second);
                if (kv.second == 1 || kv.second == 3) f->pos_ids[kv.first].push_back((int)id);
                else f->neg_ids[kv.first].push_back((int)id);
            }
        return;
    }
    bool all_single = true;
    for (auto& m : motifs) if (m.length != 1) all_single = false;
    if (locs.size() == 1) {
        for (auto& kv : locs[0]) {
            if (!in(kv.first)) continue;
            merge_rule(kv.first, kv.second);
            if (kv.second & 1) f->pos_ids[kv.first] = {0};
            if (kv.second & 2) f->neg_ids[kv.first] = {0};
        }
    } else if (all_single) {
        auto find_id = [&](const char* b) -> int { for (size_t i = 0; i < motifs.size(); i++) if (motifs[i].raw == b) return (int)i; return -1; };
        auto add_pair = [&](const char* top, const char* bot) {
            int a = find_id(top);
            if (a < 0) return;
            int t = find_id(bot);
            for (auto& kv : locs[a]) {
                if (!in(kv.first)) continue;
                if (t >= 0) { f->rule[kv.first] = 3; f->pos_ids[kv.first] = {a, t}; f->neg_ids[kv.first] = {a, t}; }
                else { f->rule[kv.first] = kv.second; if (kv.second == 1) f->pos_ids[kv.first] = {a}; else if (kv.second == 2) f->neg_ids[kv.first] = {a}; }
            }
        };
        add_pair("A", "T");
        add_pair("C", "G");
    } else {
        for (size_t id = 0; id < locs.size(); id++)
            for (auto& kv : locs[id]) {
                if (!in(kv.first)) continue;
                merge_rule(kv.first, kv.second);
                if (kv.second & 1) f->pos_ids[kv.first].push_back((int)id);
                if (kv.second & 2) f->neg_ids[kv.first].push_back((int)id);
            }
    }
}

struct MotifLookup {
    Fasta fa;
    std::vector<Motif> motifs;
    bool mask = false;
    uint64_t longest = 0;
    const PositionFilter* pf = nullptr;
    uint32_t cur_tid = 0;
    // motif hits are restricted to the include-bed positions of the matching strand (src/fasta.rs:55-69)
    void apply_filter(std::vector<MotifLocs>* locs) const {
        if (!pf) return;
        for (auto& l : *locs) for (auto it = l.begin(); it != l.end();) {
            uint8_t r = it->second, keep = 0;
            if ((r & 1) && pf->contains(cur_tid, it->first, false)) keep |= 1;
            if ((r & 2) && pf->contains(cur_tid, it->first, true)) keep |= 2;
            if (!keep) it = l.erase(it); else { it->second = keep; ++it; }
        }
    }
    std::string prep(std::string s) const { if (!mask) for (char& c : s) c = (char)toupper((unsigned char)c); return s; }
    // returns interval end (fasta.rs:192-226 / 92-188)
    uint32_t positions(const std::string& contig, uint64_t ref_end, uint64_t start, uint64_t end, bool combine, std::vector<MotifLocs>* out) {
        if (!combine) {
            *out = motifs_on_seq(prep(fa.fetch(contig, start, end)), start, motifs);
            apply_filter(out);
            return (uint32_t)end;
        }
        uint64_t buffer = longest * 5;
        uint64_t e = end;
        uint64_t ewb = std::min(end + buffer, ref_end);
        uint64_t too_close = ewb >= longest ? ewb - longest : 0;
        while (true) {
            std::vector<MotifLocs> locs = motifs_on_seq(prep(fa.fetch(contig, start, ewb)), start, motifs);
            apply_filter(&locs);
            // merged motif intervals (rust-lapper merge_overlaps: touching intervals merge)
            std::vector<std::pair<uint64_t, uint64_t>> ivs;
            for (size_t id = 0; id < locs.size(); id++) {
                uint64_t adj = motifs[id].length >= motifs[id].fwd_off ? motifs[id].length - motifs[id].fwd_off : motifs[id].length;
                for (auto& kv : locs[id]) ivs.push_back({kv.first, kv.first + adj});
            }
            std::sort(ivs.begin(), ivs.end());
            std::vector<std::pair<uint64_t, uint64_t>> merged;
            for (auto& iv : ivs) {
                if (!merged.empty() && !(merged.back().second < iv.first)) merged.back().second = std::max(merged.back().second, iv.second);
                else merged.push_back(iv);
            }
            uint64_t search_end = e;
            uint64_t q0 = e ? e - 1 : 0, q1 = e;
            for (auto& iv : merged) if (iv.first < q1 && iv.second > q0) { search_end = iv.second; break; }
            if (search_end < too_close || ewb >= ref_end) {
                for (auto& l : locs) for (auto it = l.begin(); it != l.end();) { if (it->first > search_end) it = l.erase(it); else ++it; }
                *out = std::move(locs);
                return (uint32_t)search_end;
            }
            e = ewb;
            ewb += buffer;
            if (ewb > ref_end) ewb = ref_end;  // (reference does not clamp; a fetch past the contig end would error there)
            too_close = ewb >= longest ? ewb - longest : 0;
        }
    }
};

// ReferenceIntervalsFeeder, flattened: every ChromCoordinates in feeder order. `batches` receives the
// grouping into MultiChromCoordinates (only the sampler needs it).
inline std::vector<Interval> make_intervals(const std::vector<RefRecord>& contigs, uint32_t interval_size,
                                            bool combine_strands, MotifLookup* lookup,
                                            std::vector<std::vector<size_t>>* batches = nullptr, const PositionFilter* pf = nullptr) {
    std::vector<Interval> out;
    std::vector<size_t> batch;
    uint64_t batch_len = 0;
    for (const RefRecord& c : contigs) {
        if (c.length == 0) continue;
        uint32_t cur = c.start;
        while (true) {
            uint32_t start = cur;
            uint32_t end = (uint32_t)std::min<uint64_t>((uint64_t)start + interval_size, c.end());
            Interval iv;
            iv.tid = c.tid;
            iv.start = start;
            if (lookup) {
                std::vector<MotifLocs> locs;
                lookup->cur_tid = c.tid;
                end = lookup->positions(c.name, c.end(), start, end, combine_strands, &locs);
                end = std::min(end, c.end());
                iv.end = end;
                focus_from_locs(locs, lookup->motifs, start, end, combine_strands, &iv.focus);
            } else {
                iv.end = end;
                if (pf) focus_from_regions(*pf, c.tid, start, end, &iv.focus);
            }
            batch_len += iv.end > iv.start ? iv.end - iv.start : 0;
            batch.push_back(out.size());
            out.push_back(std::move(iv));
            if (batch_len >= interval_size) { if (batches) batches->push_back(batch); batch.clear(); batch_len = 0; }
            if (end >= c.end()) break;
            cur = end;
        }
    }
    if (!batch.empty() && batches) batches->push_back(batch);
    return out;
}

// --- threshold estimation ----------------------------------------------------------------------
inline float percentile_linear_interp(const std::vector<float>& xs, float q) {  // thresholds.rs:17-39
    if (xs.size() < 2) throw std::runtime_error("not enough datapoints for percentile: " + std::to_string(xs.size()));
    if (q > 1.0f) throw std::runtime_error("invalid quantile");
    if (q == 1.0f) return xs.back();
    float l = (float)(xs.size() - 1);
    float x = l * q;
    float left = std::floor(x);
    size_t right = (size_t)std::ceil(x);
    float g = x - std::trunc(x);
    float y0 = xs[(size_t)left], y1 = xs[right];
    volatile float a = y0 * (1.0f - g);
    volatile float b = y1 * g;
    return a + b;
}

struct SampleOptions {
    int threads = 4;
    uint32_t sampling_interval_size = 1000000;
    bool frac_all = false;        // -f 1.0
    size_t num_reads = 10042;
    const Region* region = nullptr;
    bool include_unmapped = false;
    bool collapse = false;
    ModCode collapse_code = 0;
    EdgeFilter edge;
    const PositionFilter* pf = nullptr;
};

// values (argmax probabilities) per canonical base from one record; false when the record contributes nothing
inline bool sample_record(const BamRecord& r, const SampleOptions& o, std::vector<float> vals[4], std::vector<BaseModProbs>* probs = nullptr) {
    ModBaseInfo info;
    std::string fwd;
    if (!decode_mod_base_info(r, &info, &fwd) || info.is_empty()) return false;
    const bool only_mapped = !o.include_unmapped;
    std::vector<int64_t> q2r;
    if (only_mapped) aligned_ref_positions(r, &q2r);
    const int L = r.l_seq();
    const bool rev = r.is_reverse();
    bool added = false;
    for (int s = 0; s < 2; s++) for (int b = 0; b < 4; b++) {
        if (!info.present[s][b]) continue;
        int cb = s == 0 ? b : comp_idx(b);
        if (o.edge.on && !o.edge.read_can_be_trimmed((size_t)L)) continue;
        size_t kept = 0;
        for (auto& kv : info.tab[s][b].pos) {
            uint32_t f = kv.first;
            if (o.edge.on && !o.edge.keep(f, (size_t)L)) continue;
            int64_t rpos = -1;
            if (only_mapped) { int q = rev ? L - 1 - (int)f : (int)f; if (q < 0 || q >= L || q2r[q] < 0) continue; rpos = q2r[q]; }
            if (o.pf) {   // filter_positions: aligned position on the matching reference strand (read_ids_to_base_mod_probs.rs:1020-1047)
                if (rpos < 0) continue;
                const bool neg_strand = (s == 0) == rev;
                if (!o.pf->contains((uint32_t)r.tid(), (uint64_t)rpos, neg_strand)) continue;
            }
            BaseModProbs bmp = o.collapse ? redistribute(kv.second, o.collapse_code) : kv.second;
            vals[cb].push_back(argmax_prob(bmp));
            if (probs) probs[cb].push_back(bmp);
            kept++;
        }
        if (kept) added = true;
    }
    return added;
}

inline bool sampler_admits(const BamRecord& r, bool only_mapped_or_edge) {
    uint16_t f = r.flag();
    if (f & (0x100 | 0x400 | 0x800)) return false;   // record_is_not_primary
    if (r.l_seq() == 0) return false;
    if (only_mapped_or_edge && (f & 0x4)) return false;
    return true;
}

// Returns per-base thresholds (base_set / base_thr of the Caller)
inline void estimate_thresholds(const BamFile& bam, const SampleOptions& o, float percentile, Caller* caller,
                                std::vector<float> all_vals_out[4] = nullptr, std::vector<const BamRecord*>* selected_out = nullptr) {
    // 1. index stats restricted to the region's contig
    int region_tid = o.region ? bam.tid_of(o.region->name) : -1;
    if (o.region && region_tid < 0) throw std::runtime_error("did not find target_id for region in header");
    std::map<uint32_t, uint64_t> mapped;
    uint64_t total_mapped = 0, total_unmapped = 0;
    for (uint32_t t = 0; t < bam.ref_names.size(); t++) {
        if (o.region && (int)t != region_tid) continue;
        if (!o.region && o.pf && !o.pf->has_contig(t)) continue;
        mapped[t] = bam.n_mapped[t];
        total_mapped += bam.n_mapped[t];
        total_unmapped += bam.n_unmapped[t];
    }
    if (!o.region) total_unmapped += bam.n_no_coor;
    uint64_t total = o.include_unmapped ? total_mapped + total_unmapped : total_mapped;
    if (total == 0) throw std::runtime_error("zero reads found in bam index");
    // 2. per contig quota: -1 == All
    std::map<uint32_t, int64_t> quota;
    for (auto& kv : mapped) {
        if (kv.second == 0) continue;
        if (o.frac_all) { quota[kv.first] = -1; continue; }
        float frac = (float)kv.second / (float)total;
        uint64_t n = (uint64_t)std::ceil((float)o.num_reads * frac);
        quota[kv.first] = (int64_t)std::min<uint64_t>(n, kv.second);
    }
    // (pruning loop of sampling_schedule.rs:218-250 not restated: unreachable for < ~5000 contigs)
    // 3. geometry
    std::vector<RefRecord> contigs;
    for (auto& c : get_targets(bam, o.region)) if (quota.count(c.tid)) contigs.push_back(c);
    std::map<uint32_t, uint32_t> contig_sizes;
    for (auto& c : contigs) contig_sizes[c.tid] = c.length;
    std::vector<std::vector<size_t>> batches;
    std::vector<Interval> ivs = contigs.empty() ? std::vector<Interval>() : make_intervals(contigs, o.sampling_interval_size, false, nullptr, &batches);
    const size_t B = (size_t)std::floor((float)o.threads * 1.5f);
    std::map<uint32_t, size_t> sampled_so_far;
    std::unordered_set<const uint8_t*> seen;   // read identity across intervals (reference keys by read name)
    std::vector<float> vals[4];
    const bool only_mapped = !o.include_unmapped;

    for (size_t sb = 0; sb < batches.size(); sb += std::max<size_t>(B, 1)) {
        // 4. accumulate_sample_counts over one super batch
        std::vector<size_t> coords;
        for (size_t k = sb; k < std::min(batches.size(), sb + std::max<size_t>(B, 1)); k++) for (size_t i : batches[k]) coords.push_back(i);
        std::sort(coords.begin(), coords.end(), [&](size_t a, size_t b) { return ivs[a].tid != ivs[b].tid ? ivs[a].tid < ivs[b].tid : ivs[a].start < ivs[b].start; });
        std::map<uint32_t, uint32_t> len_c;
        for (size_t i : coords) len_c[ivs[i].tid] += ivs[i].end - ivs[i].start;
        std::map<uint32_t, int64_t> k_c;  // -1 all
        for (auto& kv : len_c) {
            auto q = quota.find(kv.first);
            if (q == quota.end()) continue;
            if (q->second < 0) { k_c[kv.first] = -1; continue; }
            size_t so_far = sampled_so_far.count(kv.first) ? sampled_so_far[kv.first] : 0;
            if ((size_t)q->second <= so_far) continue;
            size_t remaining = (size_t)q->second - so_far;
            float f = (float)kv.second / (float)contig_sizes[kv.first];
            k_c[kv.first] = (int64_t)std::ceil(f * (float)remaining);
        }
        struct Grp { uint32_t tid, start, end; int64_t n; };
        std::vector<Grp> grouped;
        bool have_slack = false;
        Grp slack{};
        for (size_t i : coords) {
            const Interval& iv = ivs[i];
            auto kc = k_c.find(iv.tid);
            if (kc == k_c.end()) continue;
            if (kc->second < 0) { grouped.push_back({iv.tid, iv.start, iv.end, -1}); continue; }
            float f = (float)(iv.end - iv.start) / (float)len_c[iv.tid];
            int64_t x = (int64_t)std::ceil((float)kc->second * f);
            Grp cur{iv.tid, iv.start, iv.end, x};
            if (x < 50) {
                if (have_slack) {
                    if (slack.tid == cur.tid) {
                        Grp m{cur.tid, std::min(slack.start, cur.start), std::max(slack.end, cur.end), slack.n + x};
                        if (m.n < 50) slack = m; else { grouped.push_back(m); have_slack = false; }
                    } else { grouped.push_back(slack); slack = cur; }
                } else { slack = cur; have_slack = true; }
            } else {
                if (have_slack) {
                    have_slack = false;
                    if (slack.tid == cur.tid) grouped.push_back({cur.tid, std::min(slack.start, cur.start), std::max(slack.end, cur.end), slack.n + x});
                    else { grouped.push_back(slack); grouped.push_back(cur); }
                } else grouped.push_back(cur);
            }
        }
        if (have_slack) grouped.push_back(slack);
        // 5. per interval: first n contributing admissible records in file order
        for (const Grp& g : grouped) {
            if (o.pf && !o.pf->overlaps(g.tid, g.start, g.end)) continue;
            size_t used = 0, returned = 0;
            std::unordered_set<const uint8_t*> seen_here;
            std::vector<const BamRecord*> cands;
            bam.fetch(g.tid, g.start, g.end, [&](const BamRecord& r) { cands.push_back(&r); });
            for (const BamRecord* rp : cands) {
                const BamRecord& r = *rp;
                if (!sampler_admits(r, only_mapped || o.edge.on)) continue;
                // with_mod_base_info(): records whose tags fail to decode / are empty never reach the sampler
                std::vector<float> v[4];
                ModBaseInfo probe;
                std::string fwd;
                if (!decode_mod_base_info(r, &probe, &fwd) || probe.is_empty()) continue;
                if (g.n >= 0 && used >= (size_t)g.n) break;
                if (seen_here.count(r.data)) continue;
                bool added = sample_record(r, o, v);
                if (added) {
                    seen_here.insert(r.data);
                    returned++;
                    used++;
                    if (seen.insert(r.data).second) { for (int b = 0; b < 4; b++) vals[b].insert(vals[b].end(), v[b].begin(), v[b].end()); if (selected_out) selected_out->push_back(&r); }
                }
            }
            sampled_so_far[g.tid] += returned;
        }
    }
    // unmapped / unplaced reads (reads_sampler/mod.rs:85-129): schedule.has_unmapped() holds whenever
    // --include-unmapped was given
    if (!only_mapped) {
        size_t limit = o.frac_all ? (size_t)-1 : (o.num_reads > seen.size() ? o.num_reads - seen.size() : 0);
        size_t used = 0;
        for (uint32_t idx : bam.unplaced) {
            const BamRecord& r = bam.records[idx];
            if (!sampler_admits(r, o.edge.on)) continue;
            ModBaseInfo probe;
            std::string fwd;
            if (!decode_mod_base_info(r, &probe, &fwd) || probe.is_empty()) continue;
            if (used >= limit) break;
            std::vector<float> v[4];
            if (sample_record(r, o, v)) {
                used++;
                if (seen.insert(r.data).second) { for (int b = 0; b < 4; b++) vals[b].insert(vals[b].end(), v[b].begin(), v[b].end()); if (selected_out) selected_out->push_back(&r); }
            }
        }
    }
    for (int b = 0; b < 4; b++) {
        if (vals[b].empty()) continue;
        std::sort(vals[b].begin(), vals[b].end());
        caller->base_set[b] = true;
        caller->base_thr[b] = percentile_linear_interp(vals[b], percentile);
        if (all_vals_out) all_vals_out[b] = vals[b];
    }
}

// --- `modkit summary` (src/summarize.rs:117-252) and the two report formats (src/writers.rs:394-684) ------------------------
// The reference prints hash maps in their iteration order; rows here are ordered: bases A C G T, canonical before modified
// states, codes in ModCodeRepr order (the same order the product prints).
struct ModSummaryOut {
    uint64_t reads_with[4] = {0, 0, 0, 0};
    std::map<uint64_t, uint64_t> pass[4], fail[4];      // key 0 = canonical, else 1 + code order key
    std::set<uint64_t> observed[4];
    uint64_t total_reads = 0;
};
inline uint64_t state_order_key(ModCode c) { return 1ull + ((c & 0x80000000u) ? (1ull << 32) + (c & 0x7fffffffu) : (uint64_t)c); }
inline std::string state_label_of_key(uint64_t k) { return k > (1ull << 32) ? std::to_string(k - 1 - (1ull << 32)) : std::string(1, (char)(k - 1)); }

inline void summarize_reads(const std::vector<const BamRecord*>& selected, const SampleOptions& o, const Caller& caller, ModSummaryOut* S) {
    for (const BamRecord* rp : selected) {
        std::vector<float> v[4];
        std::vector<BaseModProbs> probs[4];
        if (!sample_record(*rp, o, v, probs)) continue;
        S->total_reads++;
        for (int b = 0; b < 4; b++) {
            if (probs[b].empty()) continue;
            S->reads_with[b]++;
            for (const BaseModProbs& bmp : probs[b]) {
                bmp.probs.for_each([&](ModCode c, float) { S->observed[b].insert(state_order_key(c)); });
                const Call tc = make_call(caller, b, bmp);
                // arg-max call (mod_bam.rs:489-505), last maximum wins
                const float cp = bmp.canonical_prob();
                bool have = false; float mp = 0.f; ModCode mc = 0;
                bmp.probs.for_each([&](ModCode c, float p) { if (!have || p >= mp) { have = true; mp = p; mc = c; } });
                const bool arg_mod = have && mp > cp;
                if (tc.kind == CALL_CANONICAL) S->pass[b][0]++;
                else if (tc.kind == CALL_MODIFIED) S->pass[b][state_order_key(tc.code)]++;
                else if (arg_mod) S->fail[b][state_order_key(mc)]++;
                else S->fail[b][0]++;
            }
        }
    }
}

inline std::string f64_display(double v) {
    if (std::isnan(v)) return "NaN";
    if (std::isinf(v)) return v > 0 ? "inf" : "-inf";
    char buf[400];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}
inline std::string f32_display(float v);

inline std::string summary_text(const ModSummaryOut& S, const Caller& caller, bool tsv, const std::string& region_text) {
    std::string out, bases;
    for (int b = 0; b < 4; b++) if (!S.pass[b].empty() || S.reads_with[b]) { if (!bases.empty()) bases += ","; bases += BASES[b]; }
    if (tsv) {
        out += "mod_bases\t" + bases + "\n";
        for (int b = 0; b < 4; b++) if (S.reads_with[b]) out += std::string("count_reads_") + BASES[b] + "\t" + std::to_string(S.reads_with[b]) + "\n";
        for (int b = 0; b < 4; b++) {
            if (S.pass[b].empty() && !S.reads_with[b]) continue;
            uint64_t total = 0, total_f = 0;
            for (auto& kv : S.pass[b]) total += kv.second;
            for (auto& kv : S.fail[b]) total_f += kv.second;
            const std::string B(1, BASES[b]);
            for (auto& kv : S.pass[b]) {
                const std::string label = kv.first == 0 ? std::string("unmodified") : "modified_" + state_label_of_key(kv.first);
                auto fi = S.fail[b].find(kv.first);
                out += B + "_pass_calls_" + label + "\t" + std::to_string(kv.second) + "\n";
                out += B + "_pass_frac_" + label + "\t" + f64_display((double)kv.second / (double)total) + "\n";
                out += B + "_fail_calls_" + label + "\t" + std::to_string(fi == S.fail[b].end() ? 0 : fi->second) + "\n";
            }
            out += B + "_total_mod_calls\t" + std::to_string(total) + "\n";
            out += B + "_total_fail_mod_calls\t" + std::to_string(total_f) + "\n";
        }
        out += "total_reads_used\t" + std::to_string(S.total_reads) + "\n";
        return out;
    }
    out += "# bases             " + bases + "\n";
    out += "# total_reads_used  " + std::to_string(S.total_reads) + "\n";
    for (int b = 0; b < 4; b++) if (S.reads_with[b]) out += std::string("# count_reads_") + BASES[b] + "     " + std::to_string(S.reads_with[b]) + "\n";
    for (int b = 0; b < 4; b++) if (caller.base_set[b]) out += std::string("# pass_threshold_") + BASES[b] + "  " + f32_display(caller.base_thr[b]) + "\n";
    if (!region_text.empty()) out += "# region            " + region_text + "\n";
    std::vector<std::vector<std::string>> rows;
    rows.push_back({"base", "code", "pass_count", "pass_frac", "all_count", "all_frac"});
    for (int b = 0; b < 4; b++) {
        uint64_t total_p = 0, total_f = 0;
        for (auto& kv : S.pass[b]) total_p += kv.second;
        for (auto& kv : S.fail[b]) total_f += kv.second;
        const uint64_t total = total_p + total_f;
        std::set<uint64_t> keys;
        for (auto& kv : S.pass[b]) keys.insert(kv.first);
        for (uint64_t k : S.observed[b]) keys.insert(k);
        if (!S.pass[b].empty() || !S.fail[b].empty()) keys.insert(0);
        for (uint64_t k : keys) {
            auto pi = S.pass[b].find(k); auto fi = S.fail[b].find(k);
            const uint64_t pc = pi == S.pass[b].end() ? 0 : pi->second, fc = fi == S.fail[b].end() ? 0 : fi->second;
            rows.push_back({std::string(1, BASES[b]), k == 0 ? std::string("-") : state_label_of_key(k), std::to_string(pc),
                            f32_display((float)pc / (float)total_p), std::to_string(pc + fc), f32_display((float)(pc + fc) / (float)total)});
        }
    }
    std::vector<size_t> w(6, 0);
    for (auto& r : rows) for (size_t i = 0; i < 6; i++) w[i] = std::max(w[i], r[i].size());
    for (auto& r : rows) { for (size_t i = 0; i < 6; i++) { out += " " + r[i] + std::string(w[i] - r[i].size(), ' ') + " "; } out += "\n"; }
    return out;
}

// --- bedMethyl writer (writers.rs:87-156) ---------------------------------------------------------
inline void format_row(const Row& R, const std::string& chrom, const std::vector<std::string>& motif_labels,
                       bool mixed_delim, std::string* out) {
    char sp = mixed_delim ? ' ' : '\t';
    std::string name = code_to_string(R.code);
    if (motif_labels.size() >= 2 && R.motif_idx >= 0 && (size_t)R.motif_idx < motif_labels.size()) name += "," + motif_labels[R.motif_idx];
    float frac = (float)R.n_mod / (float)R.cov;
    volatile float pct = frac * 100.0f;
    char buf[512];
    int n = snprintf(buf, sizeof buf, "%s\t%u\t%u\t%s\t%u\t%c\t%u\t%u\t255,0,0\t%u%c%.2f%c%u%c%u%c%u%c%u%c%u%c%u%c%u\n",
                     chrom.c_str(), R.pos, R.pos + 1, name.c_str(), R.cov, R.strand, R.pos, R.pos + 1,
                     R.cov, sp, (double)pct, sp, R.n_mod, sp, R.n_canon, sp, R.n_other, sp, R.n_delete, sp,
                     R.n_filtered, sp, R.n_diff, sp, R.n_nocall);
    out->append(buf, (size_t)n);
}

// ---- --partition-tag (src/pileup/mod.rs:629-646, src/util.rs:670-688) and --bedgraph (src/writers.rs:264-381) ----
// Rust `{}` of an f32: the shortest decimal that round-trips, never in exponent form
inline std::string f32_display(float v) {
    char buf[64];
    auto r = std::to_chars(buf, buf + sizeof buf, v, std::chars_format::fixed);
    return std::string(buf, r.ptr);
}

// get_stringable_aux: text of one aux value, false for absent tags and array types
inline bool stringable_aux(const BamRecord& r, const std::string& tag, std::string* out) {
    AuxField f;
    const char t[2] = {tag[0], tag[1]};
    if (!find_aux(r, t, &f)) return false;
    switch (f.type) {
        case 'Z': case 'H': *out = std::string((const char*)f.p, f.n); return true;
        case 'A': *out = std::string(1, (char)f.p[0]); return true;
        case 'c': *out = std::to_string((int)(int8_t)f.p[0]); return true;
        case 'C': *out = std::to_string((unsigned)f.p[0]); return true;
        case 's': { int16_t v; memcpy(&v, f.p, 2); *out = std::to_string(v); return true; }
        case 'S': { uint16_t v; memcpy(&v, f.p, 2); *out = std::to_string(v); return true; }
        case 'i': { int32_t v; memcpy(&v, f.p, 4); *out = std::to_string(v); return true; }
        case 'I': { uint32_t v; memcpy(&v, f.p, 4); *out = std::to_string(v); return true; }
        case 'f': { float v; memcpy(&v, f.p, 4); *out = f32_display(v); return true; }
        default: return false;
    }
}

// parse_tags_from_record: values joined by '_', "missing" for absent ones; false when no tag is present (NoKey)
inline bool partition_key_of(const BamRecord& r, const std::vector<std::string>& tags, std::string* key) {
    bool any = false;
    std::string k;
    for (size_t i = 0; i < tags.size(); i++) {
        std::string v;
        const bool have = stringable_aux(r, tags[i], &v);
        any = any || have;
        if (i) k += "_";
        k += have ? v : std::string("missing");
    }
    if (any) *key = k;
    return any;
}

// bedGraph line + the label part of its file name (BedGraphWriter::write)
inline void format_bedgraph_row(const Row& R, const std::string& chrom, const std::vector<std::string>& motif_labels,
                                std::string* label, std::string* line) {
    *label = code_to_string(R.code);
    if (R.motif_idx >= 0 && (size_t)R.motif_idx < motif_labels.size()) {
        std::string l = motif_labels[R.motif_idx];
        l.erase(std::remove(l.begin(), l.end(), ','), l.end());
        *label += "_" + l;
    }
    const float frac = (float)R.n_mod / (float)R.cov;
    *line = chrom + "\t" + std::to_string(R.pos) + "\t" + std::to_string(R.pos + 1) + "\t" + f32_display(frac) + "\t" + std::to_string(R.cov) + "\n";
}

inline const char* strand_label(char s) { return s == '+' ? "positive" : s == '-' ? "negative" : s == '.' ? "combined" : "_unknown"; }

inline const char* bedmethyl_header() {
    return "chrom\tchromStart\tchromEnd\tname\tscore\tstrand\tthickStart\tthickEnd\tcolor\tvalid_coverage\t"
           "percent_modified\tcount_modified\tcount_canonical\tcount_other_mod\tcount_delete\tcount_fail\t"
           "count_diff\tcount_nocall\n";
}

}  // namespace orc
