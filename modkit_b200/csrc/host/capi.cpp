// Host C API used by the Python mirror (modkit_b200/__init__.py), tests and bench.py.
#include "pileup_run.hpp"

using namespace mkh;

struct mkh_bam { BamReader reader; };
struct mkh_packed { PackedChunk pc; };

extern "C" {

// In-process `modkit pileup`: returns the process exit code the reference would (0 ok, 1 error -> "> Error! ...").
int mkh_pileup_main(int argc, const char* const* argv) {
    PileupOptions o;
    std::string err;
    if (!parse_pileup_args(argc, argv, &o, &err)) { fprintf(stderr, "error: %s\n", err.c_str()); return 2; }
    RunSummary s;
    if (run_pileup(o, &s, &err)) { fprintf(stderr, "> Error! %s\n", err.c_str()); return 1; }
    return 0;
}

int mkh_bam_open(const char* path, int threads, mkh_bam** out) {
    try { mkh_bam* b = new mkh_bam(); b->reader.open(path, threads); *out = b; return 0; }
    catch (const std::exception& e) { fprintf(stderr, "mkh_bam_open: %s\n", e.what()); return -1; }
}
// device ingest: inflate + record discovery on the GPU of `ctx` (include/mkp.h, mkp_bam_load)
int mkh_bam_open_device(const char* path, mkp_ctx* ctx, mkh_bam** out) {
    try { mkh_bam* b = new mkh_bam(); b->reader.open_device(path, ctx); *out = b; return 0; }
    catch (const std::exception& e) { fprintf(stderr, "mkh_bam_open_device: %s\n", e.what()); return -1; }
}
// the reads overlapping [start,end) of tid become the resident chunk of the ingest context; returns the read count or -1
int64_t mkh_device_chunk(const mkh_bam* b, uint32_t tid, uint32_t start, uint32_t end, const uint32_t* focus_pos, const uint32_t* focus_neg) {
    try {
        std::vector<RecRef> recs;
        b->reader.ensure_tid(tid);
        b->reader.for_overlapping(tid, start, end, [&](const RecRef& r) { recs.push_back(r); });
        return (int64_t)device_chunk(b->reader, recs, start, end, focus_pos, focus_neg);
    } catch (const std::exception& e) { fprintf(stderr, "mkh_device_chunk: %s\n", e.what()); return -1; }
}
void mkh_bam_ingest_ms(const mkh_bam* b, float* ms) { for (int i = 0; i < 4; i++) ms[i] = b->reader.ingest_ms[i]; }
uint32_t mkh_bam_n_ranges(const mkh_bam* b) { return (uint32_t)b->reader.n_ranges(); }
uint64_t mkh_bam_total_records(const mkh_bam* b) { uint64_t n = b->reader.unplaced.size(); for (auto& v : b->reader.by_tid) n += v.size(); return n; }
// host-side helpers of --partition-tag / --bedgraph, exposed for the CPU tests
int mkh_f32_display(float v, char* out, int cap) { const std::string s = f32_display(v); if ((int)s.size() + 1 > cap) return -1; memcpy(out, s.c_str(), s.size() + 1); return (int)s.size(); }
// partition key of the i-th record of tid (file order) for the ':'-separated tag list; returns 1 key, 0 NoKey, -1 error
int mkh_bam_partition_key(const mkh_bam* b, uint32_t tid, uint64_t i, const char* tags, char* out, int cap) {
    try {
        std::vector<std::string> tv;
        for (const char* p = tags; *p;) { const char* q = strchr(p, ':'); if (!q) q = p + strlen(p); tv.emplace_back(p, q - p); p = *q ? q + 1 : q; }
        const RecRef& r = b->reader.by_tid.at(tid).at(i);
        std::string k;
        const bool have = partition_key_of(b->reader.rec(r), r.size, tv, &k);
        if ((int)k.size() + 1 > cap) return -1;
        memcpy(out, k.c_str(), k.size() + 1);
        return have ? 1 : 0;
    } catch (const std::exception&) { return -1; }
}
void mkh_bam_close(mkh_bam* b) { delete b; }
uint32_t mkh_bam_n_refs(const mkh_bam* b) { return (uint32_t)b->reader.ref_names.size(); }
const char* mkh_bam_ref_name(const mkh_bam* b, uint32_t tid) { return b->reader.ref_names[tid].c_str(); }
uint32_t mkh_bam_ref_len(const mkh_bam* b, uint32_t tid) { return b->reader.ref_lens[tid]; }
uint64_t mkh_bam_n_mapped(const mkh_bam* b, uint32_t tid) { return b->reader.stats.n_mapped[tid]; }
uint64_t mkh_bam_n_records(const mkh_bam* b, uint32_t tid) { return b->reader.by_tid[tid].size(); }

int mkh_pack_region(const mkh_bam* b, uint32_t tid, uint32_t start, uint32_t end, mkh_packed** out) {
    try {
        mkh_packed* p = new mkh_packed();
        pack_region_mt(b->reader, tid, start, end, (int)std::min(32u, std::max(1u, std::thread::hardware_concurrency())), &p->pc);
        *out = p;
        return 0;
    }
    catch (const std::exception& e) { fprintf(stderr, "mkh_pack_region: %s\n", e.what()); return -1; }
}
void mkh_packed_free(mkh_packed* p) { delete p; }
uint32_t mkh_packed_n_reads(const mkh_packed* p) { return (uint32_t)p->pc.hdrs.size(); }
const mkp_read_hdr* mkh_packed_hdrs(const mkh_packed* p) { return p->pc.hdrs.data(); }
const uint8_t* mkh_packed_heap(const mkh_packed* p) { return p->pc.heap.data(); }
uint64_t mkh_packed_heap_bytes(const mkh_packed* p) { return p->pc.heap.size(); }
uint64_t mkh_packed_algorithmic_bytes(const mkh_packed* p) { return p->pc.algorithmic_bytes(); }

// bedMethyl text of device rows for an all-positions run (bench: rows -> text, writers.rs:87-156)
uint64_t mkh_format_rows(const mkp_row* rows, uint64_t n, const char* chrom, int mixed_delim, char* out, uint64_t cap) {
    BedFormat fmt; fmt.mixed_delim = mixed_delim != 0;
    std::string text;
    const std::string c = chrom;
    for (uint64_t i = 0; i < n; i++) { OutRow o{rows[i], -1, (char)rows[i].strand}; format_bed_row(o, c, fmt, &text); }
    if (out && text.size() <= cap) memcpy(out, text.data(), text.size());
    return text.size();
}

// Focus bitmaps (FocusPositions::check_position rules) of [start,end) for a motif list given as "CG:0,GATC:1".
// Intervals follow the reference grid (interval_size from `start`), so motif hits straddling an interval boundary are
// dropped exactly like the reference does (src/fasta.rs:207-227).
int mkh_motif_focus(const char* fasta, const char* contig, uint32_t start, uint32_t end, uint32_t interval_size,
                    const char* motifs, int combine_strands, uint32_t* pos_bits, uint32_t* neg_bits) {
    try {
        MotifContext mc;
        mc.fasta.open(fasta);
        std::string ms = motifs;
        for (size_t i = 0; i < ms.size();) {
            size_t j = ms.find(',', i); if (j == std::string::npos) j = ms.size();
            std::string one = ms.substr(i, j - i); size_t c = one.find(':');
            mc.motifs.push_back(parse_motif(one.substr(0, c), std::stoi(one.substr(c + 1))));
            mc.longest = std::max<uint64_t>(mc.longest, mc.motifs.back().len);
            i = j + 1;
        }
        std::vector<RefTarget> t{{0, start, end - start, contig}};
        std::vector<RefInterval> ivs = reference_intervals(t, interval_size, combine_strands != 0, &mc);
        std::vector<uint32_t> fp, fn;
        focus_bitmaps(ivs, 0, ivs.size(), start, end, &fp, &fn);
        memcpy(pos_bits, fp.data(), fp.size() * 4);
        memcpy(neg_bits, fn.data(), fn.size() * 4);
        return 0;
    } catch (const std::exception& e) { fprintf(stderr, "mkh_motif_focus: %s\n", e.what()); return -1; }
}

}  // extern "C"
