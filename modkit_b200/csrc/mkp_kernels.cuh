// Device kernels of the B200-native `modkit pileup` hot path (sm_100a).
//
// Pipeline per chunk (all on one stream, intermediates stay in L2/HBM):
//   k_decode      warp per read: CIGAR prefix, MM text parse, delta -> forward position (rank/select on
//                 the 4-bit SEQ), ML -> probability, list merging, collapse, threshold call, projection
//                 to reference coordinates; emits compact call records + marks the hot-position bitmap.
//                 (reference: src/mod_bam.rs:900-1577, src/read_cache.rs:69-211,
//                  src/threshold_mod_caller.rs:28-63, src/util.rs:122-145)
//   k_rank_*      popcount prefix over the hot bitmap  -> counter slot index per hot position
//   k_count_calls one thread per call record: modcall / filtered counters (src/pileup/mod.rs:876-937)
//   k_count_bases warp per read: walks CIGAR + SEQ with bit-parallel window tests against the hot
//                 bitmap; base / delete counters and observed-code masks (src/pileup/mod.rs:831-874)
//   k_rows_*      per slot: FeatureVector::decode (src/pileup/mod.rs:283-445) -> mkp_row
// Integer/indexing work: no tensor cores; the bound is HBM bandwidth.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/mkp.h"

namespace mkp {

constexpr int MAX_LISTS = 16;
constexpr int MAX_LIST_CODES = 4;
constexpr int MAX_MAP = 7;          // codes at one read position
constexpr int MAX_STATES = 32;
constexpr int CUM_CAP = 512;        // 32-base blocks whose occurrence counts fit in shared memory (reads <= 16 kb)
#ifndef MKP_CQ_CAP
#define MKP_CQ_CAP 640
#endif
constexpr int CQ_CAP = MKP_CQ_CAP;  // CIGAR ops whose prefix sums fit in shared memory per warp
constexpr int QT_CAP = 512;         // buckets of the query -> op table
constexpr uint32_t FULL = 0xffffffffu;

// slot layout (u32 words)
constexpr int SL_BASE = 0;    // [a][b] 8 words, signed
constexpr int SL_DEL = 8;     // [a]
constexpr int SL_FILT = 10;   // [s]
constexpr int SL_CANON = 12;  // [s][pb]
constexpr int SL_OBS = 20;    // pos mask, neg mask
constexpr int SL_MOD = 22;    // [s][state]

struct DevParams {
    float default_thr;
    float base_thr[4];
    uint32_t base_set[4];
    uint32_t n_mod_thr;
    uint32_t mod_code[MKP_MAX_MOD_THRESHOLDS];
    float mod_thr[MKP_MAX_MOD_THRESHOLDS];
    uint32_t numeric_mode, collapse_code, force_allow_implicit;
    uint32_t edge_on, edge_inv, edge_start, edge_end;
};

struct ReadMeta {       // 40 bytes, written by k_decode
    int32_t ref_end;
    uint32_t flags;     // bit0 admitted, bit1 has mod info (not in skip_set), bit2 call records are one position-sorted run
    uint32_t pos_mask, neg_mask;
    uint32_t n_calls;
    uint32_t n_hist;    // sampling: values contributed
    uint64_t entry_off;
    // implicit ('.'/default mode) tables: imp[s] byte b = 0x80 | state of the inferred-canonical entry of the
    // (mod strand s, forward base b) table (src/mod_bam.rs:1265-1292); 0 = no implicit table
    uint32_t imp[2];
};

struct ReadLists;

struct ChunkDev {
    const mkp_read_hdr* hdrs;
    const uint8_t* heap;
    const uint64_t* entry_off;  // exclusive prefix of len_ml
    uint32_t n_reads;
    uint32_t cs, ce;            // chunk range
    const uint32_t* focus_pos;  // may be null
    const uint32_t* focus_neg;
    ReadMeta* meta;
    uint32_t* P;                // forward positions per raw entry
    uint2* calls;               // (ref_pos, info)
    uint32_t* hot;              // bitmap over [cs,ce)
    uint32_t* hot_prefix;
    unsigned long long* states; // registry: (pb<<32 | code), ~0 empty
    uint32_t* n_states;
    uint32_t* err;
    uint32_t* work;             // dynamic read counters: +0 parse, +1 resolve, +2/+3 counts, +4 queue length, +5 queue cursor
    uint32_t* slow_list;        // reads left to the generic k_resolve instantiation
    unsigned long long* total_calls;
    // per-warp scratch
    uint32_t* scr_cq; uint32_t* scr_cr;
    uint32_t max_ncigar, max_blocks;
    // sampling
    unsigned long long* hist;   // [4][1025]
    unsigned long long* hist_inexact;
    const uint8_t* take;
    uint32_t hist_include_unaligned;
    // `modkit summary` (src/summarize.rs:117-252): with the caller's thresholds in c_par, every value of MODE_HIST is also classed
    // by (thresholded call, arg-max call): summ[((base * 2 + fail) * 33) + s], s = 0 canonical, 1 + state id; then reads per base
    // [4] and, as u32, the states observed per base [4]. nullptr = off.
    unsigned long long* summ;
    uint32_t mode;              // MODE_PILEUP / MODE_HIST (admission rules of k_parse)
    struct ReadLists* rl;
    // chunks with focus bitmaps (the hot bitmap is the focus set, ranked at upload, and is not written):
    // 1 = list mode: the reads k_pileup_fused left to the generic kernels; k_parse walks slow_list[0 .. *(work + 4));
    // 2 = every read through the per-stage kernels
    uint32_t list_mode;
    // processing order of the dynamic read queue: longest reads first (a 200 kb read that starts last is the kernel's tail); null = index order
    const uint32_t* order;
};

__constant__ DevParams c_par;

__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { uint32_t t = __shfl_up_sync(FULL, v, d); if (lane_id() >= d) v += t; }
    return v;
}
__device__ __forceinline__ unsigned long long warp_incl_scan64(unsigned long long v) {
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) { unsigned long long t = __shfl_up_sync(FULL, v, d); if (lane_id() >= d) v += t; }
    return v;
}

// count of nibbles equal to X in a 32-bit word of BAM 4-bit sequence
__device__ __forceinline__ uint32_t nib_eq_flags(uint32_t w, uint32_t X) {
    uint32_t x = w ^ (X * 0x11111111u);
    uint32_t t = (x | (x >> 1) | (x >> 2) | (x >> 3)) & 0x11111111u;
    return t ^ 0x11111111u;  // bit 4n set where nibble n == X
}
// flags (bit 4n) -> 8-bit mask in base order (base j <-> nibble j^1 inside a little-endian word)
__device__ __forceinline__ uint32_t nib_flags_to_mask(uint32_t y) {
    y = (y | (y >> 3)) & 0x03030303u;
    y = (y | (y >> 6)) & 0x000F000Fu;
    y = (y | (y >> 12)) & 0xFFu;
    return ((y & 0x55u) << 1) | ((y & 0xAAu) >> 1);
}
__device__ __forceinline__ uint32_t seq_nibble(const uint8_t* seq, uint32_t q) {
    uint32_t b = seq[q >> 1];
    return (q & 1) ? (b & 0xf) : (b >> 4);
}
// BAM nibble -> 0..3 (A C G T) or 4
__device__ __forceinline__ int nib_to_base(uint32_t n) { return n == 1 ? 0 : n == 2 ? 1 : n == 4 ? 2 : n == 8 ? 3 : 4; }
__device__ __forceinline__ uint32_t base_to_nib(int b) { return 1u << b; }

__device__ __forceinline__ bool is_ws(uint8_t c) { return c == ' ' || c == '\t' || c == '\n' || c == '\r'; }
__device__ __forceinline__ bool is_digit(uint8_t c) { return c >= '0' && c <= '9'; }

// ---- FxHashMap<ModCodeRepr,f32> iteration order -------------------------
__device__ __forceinline__ unsigned long long fx_hash_code(uint32_t c) {
    const unsigned long long K = 0x517cc1b727220a95ull;
    unsigned long long disc = (c & 0x80000000u) ? 1ull : 0ull;
    unsigned long long h = disc * K;                       // (rotl(0,5) ^ disc) * K
    unsigned long long v = (c & 0x80000000u) ? (unsigned long long)(c & 0x7fffffffu) : (unsigned long long)c;
    h = (((h << 5) | (h >> 59)) ^ v) * K;
    return h;
}

struct ProbMap {
    uint32_t code[MAX_MAP];
    float p[MAX_MAP];
    int8_t slot[8];   // slot -> item, -1 empty
    int n, nb;
    __device__ void init() { n = 0; nb = 4; for (int i = 0; i < 8; i++) slot[i] = -1; }
    __device__ __noinline__ void place(int item) {
        int s = (int)(fx_hash_code(code[item]) & (unsigned long long)(nb - 1));
        while (slot[s] >= 0) s = (s + 1) & (nb - 1);
        slot[s] = (int8_t)item;
    }
    __device__ int find(uint32_t c) const { for (int i = 0; i < n; i++) if (code[i] == c) return i; return -1; }
    // returns item index, or -1 on overflow
    __device__ __noinline__ int insert(uint32_t c, float v) {
        if (n >= MAX_MAP) return -1;
        if (nb == 4 && n + 1 > 3) {
            int8_t old[4];
            for (int i = 0; i < 4; i++) old[i] = slot[i];
            nb = 8;
            for (int i = 0; i < 8; i++) slot[i] = -1;
            for (int i = 0; i < 4; i++) if (old[i] >= 0) place(old[i]);
        }
        code[n] = c; p[n] = v;
        place(n);
        return n++;
    }
    // item at iteration rank k (k < n)
    __device__ int item_at(int k) const {
        for (int s = 0; s < 8; s++) if (slot[s] >= 0) { if (k == 0) return slot[s]; k--; }
        return -1;
    }
    __device__ float sum() const {
        float s = 0.f;
        for (int k = 0; k < 8; k++) if (slot[k] >= 0) s = __fadd_rn(s, p[slot[k]]);
        return s;
    }
};

// CollapseMethod::ReDistribute (src/mod_bam.rs:558-600)
__device__ __noinline__ void redistribute(const ProbMap& in, uint32_t drop, ProbMap& out) {
    float marginal = 0.f;
    int n_other = 0;
    for (int s = 0; s < 8; s++) if (in.slot[s] >= 0) { int it = in.slot[s]; if (in.code[it] == drop) marginal = __fadd_rn(marginal, in.p[it]); else n_other++; }
    float share = __fdiv_rn(marginal, __fadd_rn((float)n_other, 1.0f));
    out.init();
    for (int s = 0; s < 8; s++) if (in.slot[s] >= 0) { int it = in.slot[s]; if (in.code[it] != drop) out.insert(in.code[it], __fadd_rn(in.p[it], share)); }
}

// MultipleThresholdModCaller::call (src/threshold_mod_caller.rs:28-63): 0 filtered, 1 canonical, 2 = modified (code in *code)
__device__ __noinline__ int make_call(const ProbMap& m, int tb, uint32_t* code) {
    bool have = false;
    float best = 0.f;
    int kind = 0;
    const uint32_t any_code = (uint32_t)("ACGT"[tb]);
    for (int s = 0; s < 8; s++) {
        if (m.slot[s] < 0) continue;
        int it = m.slot[s];
        uint32_t c = m.code[it];
        float thr;
        bool found = false;
        for (uint32_t k = 0; k < c_par.n_mod_thr && !found; k++) if (c_par.mod_code[k] == c) { thr = c_par.mod_thr[k]; found = true; }
        for (uint32_t k = 0; k < c_par.n_mod_thr && !found; k++) if (c_par.mod_code[k] == any_code) { thr = c_par.mod_thr[k]; found = true; }
        if (!found) thr = c_par.base_set[tb] ? c_par.base_thr[tb] : c_par.default_thr;
        float pm = m.p[it];
        if (pm >= thr) { if (!have || pm >= best) { have = true; best = pm; kind = 2; *code = c; } }
    }
    float cthr = c_par.base_set[tb] ? c_par.base_thr[tb] : c_par.default_thr;
    float cp = __fsub_rn(1.0f, m.sum());
    if (cp >= cthr) { if (!have || cp >= best) { have = true; best = cp; kind = 1; } }
    return have ? kind : 0;
}
// BaseModProbs::argmax_base_mod_call value (src/mod_bam.rs:489-505)
__device__ __noinline__ float argmax_prob(const ProbMap& m) {
    float cp = __fsub_rn(1.0f, m.sum());
    bool have = false;
    float mp = 0.f;
    for (int s = 0; s < 8; s++) if (m.slot[s] >= 0) { float v = m.p[m.slot[s]]; if (!have || v >= mp) { have = true; mp = v; } }
    return (have && mp > cp) ? mp : cp;
}

// ---- probabilities of one read position in insertion order (BaseModProbs) ---------------------------------------
struct Items {
    uint32_t code[MAX_MAP];
    float p[MAX_MAP];
    int n;
    __device__ __forceinline__ int find(uint32_t c) const { for (int i = 0; i < n; i++) if (code[i] == c) return i; return -1; }
    __device__ __forceinline__ bool push(uint32_t c, float v) { if (n >= MAX_MAP) return false; code[n] = c; p[n] = v; n++; return true; }
};
// low two bits of the Fx hash: K = 1 (mod 4), rotl(K,5) = 2 (mod 4)
__device__ __forceinline__ uint32_t bucket4(uint32_t c) { return (c & 0x80000000u) ? ((2u ^ c) & 3u) : (c & 3u); }
__device__ __noinline__ void iter_order_slow(const Items& it, int* ord) {
    ProbMap m;
    m.init();
    for (int i = 0; i < it.n; i++) m.insert(it.code[i], it.p[i]);
    int k = 0;
    for (int s = 0; s < 8; s++) if (m.slot[s] >= 0) ord[k++] = m.slot[s];
}
// FxHashMap iteration order of the items (1 and 2 entries in closed form: 4 buckets, linear probing)
__device__ __forceinline__ void iter_order(const Items& it, int* ord) {
    if (it.n <= 1) { ord[0] = 0; return; }
    if (it.n == 2) {
        const uint32_t b0 = bucket4(it.code[0]), b1 = bucket4(it.code[1]);
        const bool swap = b0 != b1 ? (b1 < b0) : (b0 == 3);
        ord[0] = swap ? 1 : 0; ord[1] = swap ? 0 : 1;
        return;
    }
    iter_order_slow(it, ord);
}
__device__ __forceinline__ float items_sum(const Items& it, const int* ord) {
    float s = 0.f;
    for (int k = 0; k < it.n; k++) s = __fadd_rn(s, it.p[ord[k]]);
    return s;
}
// CollapseMethod::ReDistribute (src/mod_bam.rs:558-600)
__device__ __forceinline__ void redistribute_items(const Items& in, const int* ord, uint32_t drop, Items& out) {
    float marginal = 0.f;
    int n_other = 0;
    for (int k = 0; k < in.n; k++) { const int i = ord[k]; if (in.code[i] == drop) marginal = __fadd_rn(marginal, in.p[i]); else n_other++; }
    const float share = __fdiv_rn(marginal, __fadd_rn((float)n_other, 1.0f));
    out.n = 0;
    for (int k = 0; k < in.n; k++) { const int i = ord[k]; if (in.code[i] != drop) out.push(in.code[i], __fadd_rn(in.p[i], share)); }
}
// MultipleThresholdModCaller::call (src/threshold_mod_caller.rs:28-63): 0 filtered, 1 canonical, 2 modified (*code)
__device__ __forceinline__ int make_call_items(const Items& it, const int* ord, int tb, uint32_t* code) {
    bool have = false;
    float best = 0.f;
    int kind = 0;
    const uint32_t any_code = (uint32_t)("ACGT"[tb]);
    const float base_thr = c_par.base_set[tb] ? c_par.base_thr[tb] : c_par.default_thr;
    for (int k = 0; k < it.n; k++) {
        const int i = ord[k];
        const uint32_t c = it.code[i];
        float thr = base_thr;
        if (c_par.n_mod_thr) {
            bool found = false;
            for (uint32_t t = 0; t < c_par.n_mod_thr && !found; t++) if (c_par.mod_code[t] == c) { thr = c_par.mod_thr[t]; found = true; }
            for (uint32_t t = 0; t < c_par.n_mod_thr && !found; t++) if (c_par.mod_code[t] == any_code) { thr = c_par.mod_thr[t]; found = true; }
        }
        const float pm = it.p[i];
        if (pm >= thr && (!have || pm >= best)) { have = true; best = pm; kind = 2; *code = c; }
    }
    const float cp = __fsub_rn(1.0f, items_sum(it, ord));
    if (cp >= base_thr && (!have || cp >= best)) { have = true; kind = 1; }
    return have ? kind : 0;
}
// BaseModProbs::argmax_base_mod_call value (src/mod_bam.rs:489-505)
__device__ __forceinline__ float argmax_items(const Items& it, const int* ord) {
    const float cp = __fsub_rn(1.0f, items_sum(it, ord));
    bool have = false;
    float mp = 0.f;
    for (int k = 0; k < it.n; k++) { const float v = it.p[ord[k]]; if (!have || v >= mp) { have = true; mp = v; } }
    return (have && mp > cp) ? mp : cp;
}

// Registry of (primary base, mod code) states. A per-thread 4-entry cache keeps the hot path off the (single,
// heavily shared) global table: every warp asking L2 for the same sector serialises at one slice.
struct StateCache {      // two most recent (key -> id) pairs, in registers
    unsigned long long k0, k1;
    int i0, i1;
    __device__ void init() { k0 = k1 = ~0ull; i0 = i1 = 0; }
};

__device__ __noinline__ int state_id_global(unsigned long long* states, uint32_t* n_states, uint32_t* err, unsigned long long key) {
    for (int i = 0; i < MAX_STATES; i++) {
        unsigned long long v = *((volatile unsigned long long*)&states[i]);
        if (v == key) return i;
        if (v == ~0ull) {
            unsigned long long old = atomicCAS(&states[i], ~0ull, key);
            if (old == ~0ull) { atomicMax(n_states, (uint32_t)i + 1); return i; }
            if (old == key) return i;
        }
    }
    atomicOr(err, MKP_DERR_TOO_MANY_STATES);
    return 0;
}

__device__ __forceinline__ int state_id(const ChunkDev& C, StateCache& sc, int pb, uint32_t code) {
    const unsigned long long key = ((unsigned long long)pb << 32) | code;
    if (sc.k0 == key) return sc.i0;
    if (sc.k1 == key) return sc.i1;
    const int id = state_id_global(C.states, C.n_states, C.err, key);
    sc.k1 = sc.k0; sc.i1 = sc.i0; sc.k0 = key; sc.i0 = id;
    return id;
}

// per-warp parsed MM list table
struct ListTab {
    uint32_t d_start[MAX_LISTS];   // offset of the first delta byte (after the header comma), == d_end when no deltas
    uint32_t d_end[MAX_LISTS];
    uint32_t code[MAX_LISTS][MAX_LIST_CODES];
    uint32_t n_delta[MAX_LISTS];
    uint32_t ent_off[MAX_LISTS];
    uint32_t ml_off[MAX_LISTS];
    uint8_t base[MAX_LISTS];       // fundamental base char
    uint8_t strand[MAX_LISTS];     // 0 '+', 1 '-'
    uint8_t mode[MAX_LISTS];       // 0 '?', 1 '.', 2 default
    uint8_t ncodes[MAX_LISTS];
    uint32_t n;
    uint32_t tot[4];               // occurrences of nibble-base x in SEQ (query order)
};

enum { MODE_PILEUP = 0, MODE_HIST = 1 };

// per-read hand-over between k_parse and k_resolve (global memory; only the used records are touched)
struct ListRec { uint32_t n_delta, ent_off, ml_off; uint32_t code[MAX_LIST_CODES]; uint8_t base, strand, mode, ncodes; };   // 32 bytes
struct ReadLists { uint32_t n, ent, flags, pad; uint16_t imp_lists[8]; ListRec rec[MAX_LISTS]; };

// binary search: first index in sorted P[0..n) with value >= f
__device__ __forceinline__ uint32_t lower_bound_u32(const uint32_t* P, uint32_t n, uint32_t f) {
    uint32_t lo = 0, hi = n;
    while (lo < hi) { uint32_t mid = (lo + hi) >> 1; if (P[mid] < f) lo = mid + 1; else hi = mid; }
    return lo;
}

// Kernel A: per read, everything that depends only on the read's own bytes: admission, reference span, MM list
// discovery, per-base occurrence counts, MM token parse + select -> forward positions P[] and the list table.
#ifndef MKP_MINB_PARSE
#define MKP_MINB_PARSE 10
#endif
#ifndef MKP_MINB_RESOLVE_FAST
#define MKP_MINB_RESOLVE_FAST 8
#endif
#ifndef MKP_MINB_RESOLVE
#define MKP_MINB_RESOLVE 7
#endif
#ifndef MKP_MINB_BASES
#define MKP_MINB_BASES 8
#endif
__global__ void __launch_bounds__(128, MKP_MINB_PARSE) k_parse(ChunkDev C) {
    __shared__ ListTab s_tab[4];
    __shared__ __align__(16) uint8_t s_txt[4][160];
    __shared__ uint8_t s_tok[4][132];
    __shared__ uint32_t s_cb[4][32], s_bm[4][32], s_tp[4][MAX_LISTS];
    __shared__ uint32_t s_val[4][96];
    const uint32_t lane = lane_id();
    const uint32_t wib = threadIdx.x >> 5;
    ListTab& T = s_tab[wib];

    for (;;) {
        // dynamic work distribution: reads differ in length by 2-3 orders of magnitude
        uint32_t ri = 0;
        if (lane == 0) {
            ri = atomicAdd(C.work, 1u);
            if (C.list_mode == 1) ri = ri < C.work[4] ? C.slow_list[ri] : 0xffffffffu;
            else if (C.order && ri < C.n_reads) ri = C.order[ri];
        }
        ri = __shfl_sync(FULL, ri, 0);
        if (ri >= C.n_reads) break;
        const mkp_read_hdr h = C.hdrs[ri];
        const uint32_t flag = h.flags & 0xffffu;
        const uint32_t L = h.l_seq;
        const bool rev = flag & 0x10;
        ReadMeta meta;
        meta.ref_end = h.ref_start;
        meta.flags = 0; meta.pos_mask = 0; meta.neg_mask = 0; meta.n_calls = 0; meta.n_hist = 0; meta.imp[0] = 0; meta.imp[1] = 0;
        meta.entry_off = C.entry_off[ri];
        bool admitted;
        if (C.mode == MODE_PILEUP) admitted = !(flag & (0x4 | 0x100 | 0x200 | 0x400 | 0x800)) && L > 0;
        else admitted = !(flag & (0x100 | 0x400 | 0x800)) && L > 0 && (C.take == nullptr || C.take[ri]) &&
                        !((flag & 0x4) && (!C.hist_include_unaligned || c_par.edge_on));
        if (!admitted) { if (lane == 0) { C.meta[ri] = meta; C.rl[ri].n = 0; C.rl[ri].flags = 1; } continue; }
        meta.flags = 1;
        const uint32_t* cig = (const uint32_t*)(C.heap + h.off);
        const uint8_t* seq = C.heap + h.off + 4ull * h.n_cigar;
        const uint8_t* ml = seq + ((L + 1) >> 1);
        const uint8_t* mm = ml + h.len_ml;
        // ---- phase 0: reference span (k_resolve builds the per-op prefix; here only the end is needed) ----
        {
            uint32_t rsum = 0;
#pragma unroll 4
            for (uint32_t i = lane; i < h.n_cigar; i += 32) {
                const uint32_t c = cig[i], op = c & 15;
                if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rsum += c >> 4;
            }
            meta.ref_end = (int32_t)((uint32_t)h.ref_start + __reduce_add_sync(FULL, rsum));
        }
        bool err = (h.flags & MKP_RF_TAGS_INVALID) != 0;
        // ---- phase 1: list discovery: warp-parallel scan for ';' and the first ',' of each part; the few
        //      header bytes of each part are parsed by lane 0 (src/mod_bam.rs:900-1000) ------------------
        {
            const uint32_t M = h.len_mm;
            uint32_t n = 0, seg = 0, hdr_end = 0xffffffffu;   // uniform across the warp
            bool e0 = false;                                   // lane 0 only
            auto close_part = [&](uint32_t j) {                // part = mm[seg..j)
                if (j > seg) {
                    if (n >= MAX_LISTS) { if (lane == 0) { atomicOr(C.err, MKP_DERR_TOO_MANY_LISTS); e0 = true; } }
                    else if (lane == 0 && !e0) {
                        const uint32_t hl = hdr_end == 0xffffffffu ? j : hdr_end;
                        uint32_t k = seg;
                        const uint8_t fb = mm[k];
                        bool e = !(fb == 'A' || fb == 'C' || fb == 'G' || fb == 'T' || fb == 'U' || fb == 'N');
                        k++;
                        uint8_t st = 0;
                        if (!e) { if (k >= hl) e = true; else { st = mm[k]; if (st != '+' && st != '-') e = true; k++; } }
                        uint32_t nc = 0;
                        bool seen_chebi = false;
                        int mode = 2;
                        if (!e && k < hl && is_digit(mm[k])) {
                            unsigned long long v = 0;
                            while (k < hl && is_digit(mm[k])) { v = v * 10 + (mm[k] - '0'); if (v > 0x7fffffffull) { e = true; break; } k++; }
                            T.code[n][nc++] = 0x80000000u | (uint32_t)v;
                            seen_chebi = true;
                        }
                        for (; !e && k < hl; k++) {
                            const uint8_t c = mm[k];
                            if (c == '?') mode = 0;
                            else if (c == '.') mode = 1;
                            else if (is_digit(c) || seen_chebi) e = true;
                            else if (nc >= MAX_LIST_CODES) { atomicOr(C.err, MKP_DERR_TOO_MANY_CODES); e = true; }
                            else T.code[n][nc++] = c;
                        }
                        if (nc == 0) e = true;
                        T.base[n] = fb; T.strand[n] = st == '-'; T.mode[n] = (uint8_t)mode; T.ncodes[n] = (uint8_t)nc;
                        // delta text mm[d_start..d_end) starts right after the header comma; n_delta = 0xffffffff marks
                        // "a comma follows the header, so at least one number must parse" (nom separated_list1)
                        if (hl < j) { T.d_start[n] = hl + 1; T.d_end[n] = j; T.n_delta[n] = 0xffffffffu; }
                        else { T.d_start[n] = j; T.d_end[n] = j; T.n_delta[n] = 0u; }
                        if (e) e0 = true;
                    }
                    if (n < MAX_LISTS) n++;
                }
                seg = j + 1;
                hdr_end = 0xffffffffu;
            };
            // four text bytes per lane (one aligned word), 128 per round; SWAR byte compare gives a 0x80 flag per
            // matching byte, the first flagged byte at or after a given offset is a warp min-reduction
            const int iM = (int)M;
            for (int c0 = -(int)((uintptr_t)mm & 3u); c0 < iM && !err; c0 += 128) {
                const int g0 = c0 + 4 * (int)lane;
                uint32_t zs = 0, zc = 0;
                if (g0 < iM && g0 + 4 > 0) {
                    const uint32_t w = *(const uint32_t*)(mm + g0);
                    const uint32_t xs = w ^ 0x3b3b3b3bu, xc = w ^ 0x2c2c2c2cu;
                    zs = ~(((xs & 0x7f7f7f7fu) + 0x7f7f7f7fu) | xs | 0x7f7f7f7fu);
                    zc = ~(((xc & 0x7f7f7f7fu) + 0x7f7f7f7fu) | xc | 0x7f7f7f7fu);
                    uint32_t keepm = 0xffffffffu;
                    if (g0 < 0) keepm &= 0xffffffffu << (8 * (-g0));
                    if (g0 + 4 > iM) keepm &= 0xffffffffu >> (8 * (g0 + 4 - iM));
                    zs &= keepm; zc &= keepm;
                }
                if (hdr_end != 0xffffffffu && !__any_sync(FULL, zs != 0)) continue;
                // offset (0..127 in this round, 128 = none) of the first flagged byte at or after offset t
                auto first_from = [&](uint32_t z, uint32_t t) -> uint32_t {
                    const uint32_t lb = 4 * lane;
                    if (t > lb) z = t - lb >= 4 ? 0u : (z & (0xffffffffu << (8 * (t - lb))));
                    const uint32_t cand = z ? lb + (((uint32_t)__ffs(z) - 1u) >> 3) : 128u;
                    return __reduce_min_sync(FULL, cand);
                };
                uint32_t lo = 0;
                while (lo < 128) {
                    const uint32_t sp = first_from(zs, lo);
                    if (hdr_end == 0xffffffffu) { const uint32_t cp = first_from(zc, lo); if (cp < sp) hdr_end = (uint32_t)(c0 + (int)cp); }
                    if (sp == 128) break;
                    close_part((uint32_t)(c0 + (int)sp));
                    lo = sp + 1;
                }
            }
            if (!err && seg < M) close_part(M);
            if (__shfl_sync(FULL, e0 ? 1 : 0, 0)) err = true;
            if (lane == 0) T.n = err ? 0xffffffffu : n;
        }
        __syncwarp();
        if (T.n == 0xffffffffu) err = true;
        const uint32_t nl = err ? 0 : T.n;
        // ---- which bases need rank/select over the SEQ (query-order nibble index x: 0..3 = A C G T as stored) ----
        uint32_t need = 0;
        for (uint32_t l = 0; l < nl; l++) {
            uint8_t fb = T.base[l];
            if (fb == 'N') continue;
            int b = fb == 'A' ? 0 : fb == 'C' ? 1 : fb == 'G' ? 2 : 3;
            int x = rev ? 3 - b : b;
            need |= 1u << x;
        }
        const uint32_t nblk = (L + 31) >> 5;
        if (lane < 4) T.tot[lane] = 0;
        __syncwarp();
        // ---- phase 3: tokens -> forward positions ------------------------------------------------------
        uint32_t* P = C.P + meta.entry_off;
        uint32_t ent = 0, mlp = 0;
        uint32_t alias_mask = 0;              // lists that share the entries of the list before them
        for (uint32_t l = 0; l < nl && !err; l++) {
            const uint32_t ds = T.d_start[l], de = T.d_end[l];
            const bool must = T.n_delta[l] == 0xffffffffu;
            unsigned long long carry = 0;     // sum of (d+1) so far
            uint32_t ntok = 0;
            // A list whose delta text repeats the previous list's byte for byte, on the same base (basecallers write one
            // delta list per code: C+h?,...;C+m?,...), has the same tokens, occurrence indices and positions: it shares the
            // previous list's entries instead of being tokenised and selected again.
            if (l > 0 && must && de > ds && T.base[l] == T.base[l - 1] && de - ds == T.d_end[l - 1] - T.d_start[l - 1] && T.n_delta[l - 1] > 0) {
                const uint32_t pds = T.d_start[l - 1], len = de - ds;
                bool same = true;
                for (uint32_t c0 = 0; c0 < len && same; c0 += 128) {
                    bool eq = true;
#pragma unroll
                    for (int t = 0; t < 4; t++) { const uint32_t i = c0 + 32 * t + lane; if (i < len && mm[ds + i] != mm[pds + i]) eq = false; }
                    same = __all_sync(FULL, eq);
                }
                if (same) {
                    ntok = T.n_delta[l - 1];
                    const uint32_t shared = T.ent_off[l - 1];
                    __syncwarp();
                    if (lane == 0) { T.n_delta[l] = ntok; T.ent_off[l] = shared; T.ml_off[l] = mlp; }
                    if ((unsigned long long)mlp + (unsigned long long)ntok * T.ncodes[l] > (unsigned long long)h.len_ml) err = true;
                    mlp += ntok * T.ncodes[l];
                    alias_mask |= 1u << l;
                    __syncwarp();
                    continue;
                }
            }
            // delta values of up to 32 tokens (one per lane, `mine` lanes in token order) -> forward positions P[]
            // delta values of up to 32 tokens (one per lane, `mine` lanes in token order) -> 0-based occurrence index of the
            // list's base in forward-read order (for 'N' lists that already is the forward position); stored in P[] and
            // turned into positions by the streaming pass below
            auto place = [&](unsigned long long val, bool mine, uint32_t keep_mask) {
                const unsigned long long inc = mine ? (val + 1ull) : 0ull;
                const unsigned long long pre = warp_incl_scan64(inc);
                const uint32_t idx = ntok + __popc(keep_mask & ((1u << lane) - 1u));
                bool bad = false;
                if (mine) {
                    const unsigned long long k = carry + pre - 1ull;
                    if (k >= (unsigned long long)L || ent + idx >= h.len_ml) bad = true;   // past the last base whatever the base is; more tokens than ML entries is an error anyway (and P has len_ml slots)
                    else P[ent + idx] = (uint32_t)k;
                }
                if (__any_sync(FULL, bad)) err = true;
                carry += __shfl_sync(FULL, pre, 31);
                ntok += __popc(keep_mask);
            };
            // ---- fast path: the text is exactly digits(,digits)* with <= 9 digits per number. Four text bytes per lane
            //      (one aligned word), 128 per round. The number in progress is an affine map v -> 10 v + d (reset at
            //      ','): every lane composes its four bytes, a TRUNCATED warp scan (distance 1 and 2: three lanes = 12
            //      bytes back) gives the value / digit count entering the lane, which is exact for numbers of <= 9
            //      digits, and any 10th digit is seen by the same window and sends the list to the exact path.
            //      Finished numbers (<= 2 per lane) are compacted into a queue and placed 32 at a time. Anything else
            //      (white space, empty or over-long tokens, other bytes) falls back to the exact path below.
            bool slow = false;
            if (must) {
                uint32_t* qv = s_val[wib];
                uint32_t qn = 0, v_carry = 0, nd_carry = 0;
                const int ids = (int)ds, ide = (int)de;
                for (int c0 = ids - (int)((uintptr_t)(mm + ds) & 3u); c0 < ide && !slow && !err; c0 += 128) {
                    const int g0 = c0 + 4 * (int)lane;
                    const uint32_t w = g0 < ide ? *(const uint32_t*)(mm + g0) : 0u;
                    uint32_t nb = __shfl_down_sync(FULL, w, 1) & 0xffu;
                    if (lane == 31) nb = g0 + 4 < ide ? mm[g0 + 4] : 0u;
                    const uint32_t wn = (w >> 8) | (nb << 24);        // byte j = the byte after byte j of w
                    bool odd = false;
                    uint32_t A = 1, B = 0, R = 1u << 16;              // v_out = A v_in + B; R = carries_in << 16 | digits
                    uint32_t dm = 0, cm = 0, em = 0;                  // per byte: digit, comma, last digit of a number
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int g = g0 + j;
                        const bool in = g >= ids && g < ide, nin = g + 1 < ide;
                        const uint32_t ch = (w >> (8 * j)) & 0xffu, nx = (wn >> (8 * j)) & 0xffu;
                        const bool dig = in && (ch - '0') < 10u, com = in && ch == ',';
                        // anomalies: a byte that is neither digit nor comma; an empty token (comma at the very start,
                        // before another comma, or at the very end)
                        if (in && !dig && !com) odd = true;
                        if (com && (!nin || nx == ',' || g == ids)) odd = true;
                        if (dig) { B = B * 10u + (ch - '0'); A *= 10u; R += 1u; dm |= 1u << j; if (!nin || nx == ',') em |= 1u << j; }
                        if (com) { A = 0; B = 0; R = 0; cm |= 1u << j; }
                    }
                    if (lane == 0) { B = A * v_carry + B; if (R >> 16) R += nd_carry; A = 0; R &= 0xffffu; }   // absorb the carry
#pragma unroll
                    for (int d = 1; d <= 2; d <<= 1) {
                        const uint32_t pA = __shfl_up_sync(FULL, A, d), pB = __shfl_up_sync(FULL, B, d), pR = __shfl_up_sync(FULL, R, d);
                        if (lane >= (uint32_t)d) { B = A * pB + B; A = A * pA; if (R >> 16) R = pR + (R & 0xffffu); }
                    }
                    uint32_t v = __shfl_up_sync(FULL, B, 1), n = __shfl_up_sync(FULL, R, 1) & 0xffffu;   // entering this lane
                    if (lane == 0) { v = v_carry; n = nd_carry; }
                    uint32_t e0 = 0, e1 = 0, ec = 0;
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        if ((dm >> j) & 1u) {
                            v = v * 10u + (((w >> (8 * j)) & 0xffu) - '0'); n++;
                            if (n > 9) odd = true;
                            if ((em >> j) & 1u) { if (ec == 0) e0 = v; else e1 = v; ec++; }
                        } else if ((cm >> j) & 1u) { v = 0; n = 0; }
                    }
                    if (__any_sync(FULL, odd)) { slow = true; break; }
                    const uint32_t m1 = __ballot_sync(FULL, ec >= 1), m2 = __ballot_sync(FULL, ec == 2);
                    const uint32_t lt = (1u << lane) - 1u;
                    const uint32_t at = qn + __popc(m1 & lt) + __popc(m2 & lt);
                    if (ec >= 1) qv[at] = e0;
                    if (ec == 2) qv[at + 1] = e1;
                    qn += __popc(m1) + __popc(m2);
                    v_carry = __shfl_sync(FULL, v, 31); nd_carry = __shfl_sync(FULL, n, 31);
                    __syncwarp();
                    uint32_t qh = 0;
                    while (qn - qh >= 32 && !err) { place(qv[qh + lane], true, FULL); qh += 32; }
                    if (qh) {
                        const uint32_t rest = qn - qh;
                        const uint32_t mv = lane < rest ? qv[qh + lane] : 0;
                        __syncwarp();
                        if (lane < rest) qv[lane] = mv;
                        qn = rest;
                        __syncwarp();
                    }
                }
                if (!slow && !err && qn) { const bool mine = lane < qn; place(mine ? qv[lane] : 0, mine, qn >= 32 ? FULL : ((1u << qn) - 1u)); }
                if (!slow && !err && ntok == 0) slow = true;         // (empty text: let the exact path raise the error)
            }
            if (slow) { carry = 0; ntok = 0; }
            bool stop = false;                // list truncated by a malformed token
            // ---- exact path. Tokens start after every ',' of mm[ds-1 .. de) (mm[ds-1] is the header comma); grammar per
            // token: ws* digit+ ws* (nom separated_list1). 128 text bytes per round are staged in shared memory, the
            // token starts are compacted, then every lane parses one token.
            uint8_t* txt = s_txt[wib];
            uint8_t* tok = s_tok[wib];
            for (uint32_t c0 = ds - 1; must && slow && c0 < de && !stop && !err; c0 += 128) {
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 5; k++) { const uint32_t o = lane + 32 * k, g = c0 + o; txt[o] = g < de ? mm[g] : 0; }
                __syncwarp();
                // commas among the first 128 staged bytes (only those before `de`)
                const uint32_t w4 = ((const uint32_t*)txt)[lane];
                const uint32_t xr = w4 ^ 0x2c2c2c2cu;
                uint32_t zf = ~(((xr & 0x7f7f7f7fu) + 0x7f7f7f7fu) | xr | 0x7f7f7f7fu);   // 0x80 in every byte that is ','
                const uint32_t nvalid = de - c0;                                           // staged bytes that are text
                if (4 * lane + 4 > nvalid) { const uint32_t keepb = nvalid > 4 * lane ? nvalid - 4 * lane : 0; zf &= keepb ? (0xffffffffu >> (8 * (4 - keepb))) : 0u; }
                const uint32_t ncom = __popc(zf);
                const uint32_t incl = warp_incl_scan(ncom);
                uint32_t at = incl - ncom;
                while (zf) { const uint32_t bi = (uint32_t)(__ffs(zf) - 1) >> 3; zf &= zf - 1; tok[at++] = (uint8_t)(4 * lane + bi + 1); }
                const uint32_t nt = __shfl_sync(FULL, incl, 31);
                __syncwarp();
                for (uint32_t g0 = 0; g0 < nt && !stop && !err; g0 += 32) {
                    const bool starts = g0 + lane < nt;
                    unsigned long long val = 0;
                    bool ok_start = false, clean = false;
                    if (starts) {
                        uint32_t k = tok[g0 + lane];                      // offset in the staged text; text ends at nvalid
                        auto chr = [&](uint32_t o) -> uint8_t { return o < 160 ? txt[o] : mm[c0 + o]; };
                        while (k < nvalid && is_ws(chr(k))) k++;
                        const uint32_t d0 = k;
                        while (k < nvalid && is_digit(chr(k))) { val = val * 10 + (chr(k) - '0'); if (val > 0xffffffffull) val = 0x1ffffffffull; k++; }
                        ok_start = k > d0 && val <= 0xffffffffull;
                        while (k < nvalid && is_ws(chr(k))) k++;
                        clean = (k >= nvalid) || (chr(k) == ',');
                    }
                    const uint32_t m_start = __ballot_sync(FULL, starts);
                    const uint32_t m_badstart = __ballot_sync(FULL, starts && !ok_start);
                    const uint32_t m_dirty = __ballot_sync(FULL, starts && ok_start && !clean);
                    // first terminating token: a bad start excludes itself, a dirty end includes itself
                    const uint32_t first_bad = m_badstart ? (uint32_t)__ffs(m_badstart) - 1 : 32;
                    const uint32_t first_dirty = m_dirty ? (uint32_t)__ffs(m_dirty) - 1 : 32;
                    uint32_t keep_mask = m_start;
                    if (first_bad < 32 || first_dirty < 32) {
                        stop = true;
                        const uint32_t cut = first_bad <= first_dirty ? first_bad : first_dirty + 1;   // lanes < cut are kept
                        keep_mask &= cut >= 32 ? FULL : ((1u << cut) - 1u);
                    }
                    if (ntok == 0 && g0 == 0 && !(keep_mask & 1u)) err = true;             // the first token must parse
                    place(val, (keep_mask >> lane) & 1u, keep_mask);
                }
            }
            if (must && ntok == 0) err = true;
            // N lists: the reference checks only deltas after the first against the sequence length; the first
            // then panics on indexing, which we turn into a read error (handled by `bad` above).
            __syncwarp();
            if (lane == 0) { T.n_delta[l] = ntok; T.ent_off[l] = ent; T.ml_off[l] = mlp; }
            if ((unsigned long long)mlp + (unsigned long long)ntok * T.ncodes[l] > (unsigned long long)h.len_ml) err = true;
            mlp += ntok * T.ncodes[l];
            ent += ntok;
            __syncwarp();
        }
        // ---- streaming select: for every needed base, one pass over the SEQ in forward-read order (32 bases per lane:
        //      occurrence mask by SWAR nibble compare, warp scan of the popcounts) during which the tokens of all lists of
        //      that base are placed: a token with occurrence index k lands in the block whose running count brackets k
        //      (5-step search over the 32 counts in shared memory), at the (k - count)-th set bit of its mask.
        if (!err && need) {
            const uint32_t* seqw = (const uint32_t*)seq;   // 4-byte aligned
            const uint32_t nbytes = (L + 1) >> 1;
            for (int x = 0; x < 4 && !err; x++) {
                if (!(need & (1u << x))) continue;
                const int bfw = rev ? 3 - x : x;                        // forward base of this nibble
                uint32_t lists = 0;                                      // lists whose base is bfw
                for (uint32_t l = 0; l < nl; l++) {
                    const uint8_t fb = T.base[l];
                    const int b = fb == 'A' ? 0 : fb == 'C' ? 1 : fb == 'G' ? 2 : (fb == 'T' || fb == 'U') ? 3 : 4;
                    if (b == bfw) lists |= 1u << l;
                }
                lists &= ~alias_mask;                                    // shared entries are placed once, by their first list
                if (lane < MAX_LISTS) s_tp[wib][lane] = 0;
                __syncwarp();
                uint32_t run = 0;                                        // occurrences in the blocks already streamed
                for (uint32_t i0 = 0; i0 < nblk; i0 += 32) {
                    const uint32_t bi = i0 + lane;                       // index in streaming (forward-read) order
                    const bool valid = bi < nblk;
                    const uint32_t blk = rev ? nblk - 1u - bi : bi;      // block in query order
                    uint32_t bm = 0;
                    if (valid) {
#pragma unroll
                        for (int w = 0; w < 4; w++) {
                            const uint32_t byte0 = blk * 16 + w * 4;
                            if (byte0 >= nbytes) continue;
                            uint32_t word = seqw[blk * 4 + w];
                            const uint32_t vb = nbytes - byte0;
                            if (vb < 4) word &= (1u << (8 * vb)) - 1u;
                            bm |= nib_flags_to_mask(nib_eq_flags(word, 1u << x)) << (8 * w);
                        }
                    }
                    const uint32_t cnt = __popc(bm);
                    const uint32_t incl = warp_incl_scan(cnt);
                    s_cb[wib][lane] = run + incl - cnt;
                    s_bm[wib][lane] = bm;
                    const uint32_t hi = run + __shfl_sync(FULL, incl, 31);      // tokens with index in [run, hi) live here
                    __syncwarp();
                    uint32_t ls = lists;
                    while (ls) {
                        const uint32_t l = __ffs(ls) - 1;
                        ls &= ls - 1;
                        const uint32_t n = T.n_delta[l];
                        uint32_t* Kp = P + T.ent_off[l];
                        uint32_t tp = s_tp[wib][l];
                        while (tp < n) {
                            const uint32_t t = tp + lane;
                            const uint32_t kv = t < n ? Kp[t] : 0xffffffffu;
                            const bool in = t < n && kv < hi;
                            if (in) {
                                // largest b with s_cb[b] <= kv
                                uint32_t lo = 0, hb = 32;
#pragma unroll
                                for (int st = 0; st < 5; st++) { const uint32_t mid = (lo + hb) >> 1; if (s_cb[wib][mid] <= kv) lo = mid; else hb = mid; }
                                uint32_t msk = s_bm[wib][lo];
                                uint32_t within = kv - s_cb[wib][lo];
                                if (rev) within = __popc(msk) - 1u - within;          // forward order runs down the query
                                uint32_t pos = 0, c;
                                c = __popc(msk & 0xffffu); if (within >= c) { within -= c; pos += 16; msk >>= 16; }
                                c = __popc(msk & 0xffu);   if (within >= c) { within -= c; pos += 8;  msk >>= 8; }
                                c = __popc(msk & 0xfu);    if (within >= c) { within -= c; pos += 4;  msk >>= 4; }
                                c = __popc(msk & 0x3u);    if (within >= c) { within -= c; pos += 2;  msk >>= 2; }
                                c = msk & 1u;              if (within >= c) { pos += 1; }
                                const uint32_t bq = rev ? nblk - 1u - (i0 + lo) : i0 + lo;
                                const uint32_t q = bq * 32 + pos;
                                Kp[t] = rev ? L - 1u - q : q;
                            }
                            const uint32_t c_in = __popc(__ballot_sync(FULL, in));
                            tp += c_in;
                            if (c_in < 32) break;
                        }
                        __syncwarp();
                        if (lane == 0) s_tp[wib][l] = tp;
                    }
                    run = hi;
                    __syncwarp();
                }
                if (lane == 0) T.tot[x] = run;
                // a token left unplaced refers to an occurrence past the end of the read (mod_bam.rs:705-727)
                uint32_t ls = lists;
                while (ls) { const uint32_t l = __ffs(ls) - 1; ls &= ls - 1; if (s_tp[wib][l] < T.n_delta[l]) err = true; }
                __syncwarp();
            }
        }
        __syncwarp();
        // '.'/default-mode lists (src/read_cache.rs:122-137, src/mod_bam.rs:1041-1043, 1265-1292)
        bool any_entries = ent > 0;
        uint32_t imp_lists[2][4] = {{0, 0, 0, 0}, {0, 0, 0, 0}};   // lists that infer canonical on (strand, base)
        if (!err) {
            for (uint32_t l = 0; l < nl; l++) {
                const uint8_t md = T.mode[l];
                if (md == 0) continue;
                const uint8_t bl = T.base[l] == 'U' ? 'T' : T.base[l];
                bool contributes;
                if (bl == 'N') contributes = T.n_delta[l] > 0;
                else { int b = bl == 'A' ? 0 : bl == 'C' ? 1 : bl == 'G' ? 2 : 3; contributes = T.tot[rev ? 3 - b : b] > 0; }
                if (!contributes) continue;
                // the (strand, base) table keeps the default mode only if every contributing list has it
                bool all_default = md == 2;
                for (uint32_t l2 = 0; l2 < nl && all_default; l2++) {
                    if (l2 == l || T.strand[l2] != T.strand[l]) continue;
                    const uint8_t b2 = T.base[l2] == 'U' ? 'T' : T.base[l2];
                    if (b2 != bl && !(b2 == 'N' && T.n_delta[l2] > 0) && bl != 'N') continue;
                    if (T.mode[l2] != 2 && (T.n_delta[l2] > 0 || T.mode[l2] == 1)) all_default = false;
                }
                if (all_default && !c_par.force_allow_implicit) { err = true; break; }   // InvalidImplicitMode: read skipped
                if (bl == 'N') continue;      // N lists get no implicit fill
                { int b = bl == 'A' ? 0 : bl == 'C' ? 1 : bl == 'G' ? 2 : 3; imp_lists[T.strand[l]][b] |= 1u << l; any_entries = true; }
            }
        }
        // ---- hand-over to k_resolve: list table (only the used records), implicit-table masks, flags -------
        __syncwarp();
        {
            ReadLists* R = C.rl + ri;
            if (lane == 0) {
                R->n = nl; R->ent = ent; R->flags = (err ? 1u : 0u) | (any_entries ? 2u : 0u);
                for (int k = 0; k < 8; k++) R->imp_lists[k] = (uint16_t)imp_lists[k >> 2][k & 3];
                C.meta[ri] = meta;
            }
            if (lane < nl) {
                ListRec rec;
                rec.n_delta = T.n_delta[lane]; rec.ent_off = T.ent_off[lane]; rec.ml_off = T.ml_off[lane];
                for (int c = 0; c < MAX_LIST_CODES; c++) rec.code[c] = T.code[lane][c];
                rec.base = T.base[lane]; rec.strand = T.strand[lane]; rec.mode = T.mode[lane]; rec.ncodes = T.ncodes[lane];
                R->rec[lane] = rec;
            }
        }
        __syncwarp();
    }
}

// Kernel B: per read with mod info: CIGAR prefix, then every list entry -> merged probabilities -> collapse ->
// threshold call -> reference position; call records, observed-code masks, implicit tables, hot-bitmap marks.
// FAST_ONLY = true: handles the reads whose lists fit the register-only path (no implicit tables, no 'N' list, at most
// two codes per position) and queues the others; FAST_ONLY = false: the generic path over the queued reads. Two
// instantiations keep the common case's register and instruction footprint small.
template <int MODE, bool FAST_ONLY>
__global__ void __launch_bounds__(128, FAST_ONLY ? MKP_MINB_RESOLVE_FAST : MKP_MINB_RESOLVE) k_resolve(ChunkDev C) {
    __shared__ ListTab s_tab[4];
    // CIGAR prefix of the current read (global scratch when longer): s_cq = 2 * query start + (op is M/=/X), one
    // sentinel entry 2 * query length at the end; s_cr = reference start; s_qt = op holding every (1 << shift)-th query base
    __shared__ __align__(16) uint32_t s_cq[4][CQ_CAP + 4], s_cr[4][CQ_CAP];
    __shared__ uint16_t s_qt[4][QT_CAP + 2];
    const uint32_t lane = lane_id();
    const uint32_t wib = threadIdx.x >> 5;
    const uint32_t gw = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    ListTab& T = s_tab[wib];
    StateCache scache;
    scache.init();
    uint32_t* const gcq = C.scr_cq + (size_t)gw * (C.max_ncigar + 4);
    uint32_t* const gcr = C.scr_cr + (size_t)gw * (C.max_ncigar + 4);

    for (;;) {
        uint32_t ri = 0;
        if (FAST_ONLY) {
            if (lane == 0) { ri = atomicAdd(C.work + 1, 1u); if (C.order && ri < C.n_reads) ri = C.order[ri]; }
            ri = __shfl_sync(FULL, ri, 0);
            if (ri >= C.n_reads) break;
        } else {
            if (lane == 0) { ri = atomicAdd(C.work + 5, 1u); ri = ri < C.work[4] ? C.slow_list[ri] : 0xffffffffu; }
            ri = __shfl_sync(FULL, ri, 0);
            if (ri == 0xffffffffu) break;
        }
        ReadMeta meta = C.meta[ri];
        if (!(meta.flags & 1)) continue;
        const ReadLists* R = C.rl + ri;
        const uint32_t nl = R->n;
        bool err = R->flags & 1u;
        const bool any_entries = R->flags & 2u;
        if (err || !any_entries) continue;          // skip_set: counts only as bases (meta already says so)
        const mkp_read_hdr h = C.hdrs[ri];
        const uint32_t L = h.l_seq;
        const bool rev = h.flags & 0x10;
        const uint32_t* cig = (const uint32_t*)(C.heap + h.off);
        uint32_t* const cq = h.n_cigar <= CQ_CAP ? s_cq[wib] : gcq;
        uint32_t* const cr = h.n_cigar <= CQ_CAP ? s_cr[wib] : gcr;
        uint16_t* const qt = s_qt[wib];
        const uint8_t* seq = C.heap + h.off + 4ull * h.n_cigar;
        const uint8_t* ml = seq + ((L + 1) >> 1);
        uint32_t imp_lists[2][4];
        for (int k = 0; k < 8; k++) imp_lists[k >> 2][k & 3] = R->imp_lists[k];
        __syncwarp();
        if (lane < nl) {
            const ListRec rec = R->rec[lane];
            T.n_delta[lane] = rec.n_delta; T.ent_off[lane] = rec.ent_off; T.ml_off[lane] = rec.ml_off;
            for (int c = 0; c < MAX_LIST_CODES; c++) T.code[lane][c] = rec.code[c];
            T.base[lane] = rec.base; T.strand[lane] = rec.strand; T.mode[lane] = rec.mode; T.ncodes[lane] = rec.ncodes;
        }
        __syncwarp();
        bool imp_any = false;
        for (int k = 0; k < 8; k++) imp_any = imp_any || imp_lists[k >> 2][k & 3];
        // register-only fast path: no implicit tables, no 'N' list, and per strand either one list with <= 2 codes or two
        // single-code lists
#ifdef MKP_NO_FAST_RESOLVE
        bool fast_read = false;
#else
        bool fast_read = !imp_any;
#endif
        {
            uint32_t cnt0 = 0, cnt1 = 0, max0 = 0, max1 = 0;
            for (uint32_t l2 = 0; l2 < nl; l2++) {
                if (T.base[l2] == 'N') fast_read = false;
                if (!T.n_delta[l2]) continue;
                const uint32_t ncd = T.ncodes[l2];
                if (T.strand[l2]) { cnt1++; max1 = max(max1, ncd); } else { cnt0++; max0 = max(max0, ncd); }
            }
            if (!(cnt0 <= 1 ? max0 <= 2 : (cnt0 == 2 && max0 == 1))) fast_read = false;
            if (!(cnt1 <= 1 ? max1 <= 2 : (cnt1 == 2 && max1 == 1))) fast_read = false;
        }
        if (FAST_ONLY && !fast_read) {                 // leave it to the generic instantiation
            if (lane == 0) C.slow_list[atomicAdd(C.work + 4, 1u)] = ri;
            continue;
        }
        // CIGAR prefix: query / reference start of every op; four ops per lane (one 16-byte load), 128 per round
        const uint32_t nc = h.n_cigar;
        uint32_t q_total = 0;
        {
            uint32_t qc = 0, rc = (uint32_t)h.ref_start;
            for (uint32_t b = 0; b < nc; b += 128) {
                const uint32_t i = b + 4 * lane;
                uint4 c4 = make_uint4(0, 0, 0, 0);
                if (i < nc) c4 = *(const uint4*)(cig + i);    // 16-byte aligned; words past nc belong to the read's own block
                const uint32_t cw[4] = {c4.x, c4.y, c4.z, c4.w};
                uint32_t ql[4], rl[4], qs = 0, rs = 0;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    const uint32_t c = i + k < nc ? cw[k] : 0u, op = c & 15, len = c >> 4;
                    ql[k] = (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) ? len : 0;
                    rl[k] = (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ? len : 0;
                    qs += ql[k]; rs += rl[k];
                }
                const uint32_t qi = warp_incl_scan(qs), rr = warp_incl_scan(rs);
                uint32_t q0 = qc + qi - qs, r0 = rc + rr - rs;
#pragma unroll
                for (int k = 0; k < 4; k++) {
                    if (i + k < nc) {
                        const uint32_t op = cw[k] & 15;
                        cq[i + k] = 2u * q0 + ((op == 0 || op == 7 || op == 8) ? 1u : 0u);
                        cr[i + k] = r0;
                    }
                    q0 += ql[k]; r0 += rl[k];
                }
                qc += __shfl_sync(FULL, qi, 31);
                rc += __shfl_sync(FULL, rr, 31);
            }
            if (lane == 0) cq[nc] = 2u * qc;
            q_total = qc;
        }
        __syncwarp();
        // query -> op table: qt[t] = the op that holds query base t << shift (the largest i with cq[i] <= that base)
        uint32_t qt_shift = 7;
        while (qt_shift < 31 && (q_total >> qt_shift) >= (uint32_t)QT_CAP) qt_shift++;
        const bool use_qt = nc > 0 && nc <= 65535u;
        if (use_qt) {
            const uint32_t B = 1u << qt_shift;
            for (uint32_t i = lane; i < nc; i += 32) {
                const uint32_t s0 = cq[i] >> 1, e0 = cq[i + 1] >> 1;
                for (uint32_t m = (s0 + B - 1) & ~(B - 1); m < e0; m += B) qt[m >> qt_shift] = (uint16_t)i;
            }
            if (lane == 0) qt[(q_total + B - 1) >> qt_shift] = (uint16_t)(nc - 1);
        }
        __syncwarp();
        uint32_t* P = C.P + meta.entry_off;
        // ---- phase 4: resolve entries -> calls ---------------------------------------------------
        uint32_t n_calls = 0, n_hist = 0;
        uint32_t n_runs = 0;          // lists that emitted call records (each list's records are sorted by position)
        uint32_t pos_mask = 0, neg_mask = 0;
        bool table_survived = false;
        uint32_t imp_explicit[8] = {0, 0, 0, 0, 0, 0, 0, 0};   // HIST: explicit values on implicit tables (per lane)
        uint32_t sum_obs[4] = {0, 0, 0, 0};                    // summary: states seen per canonical base (per lane)
        uint2* calls = C.calls + meta.entry_off;
        if (!err && any_entries) {
            const bool trim_ok = !c_par.edge_on || !(L <= c_par.edge_start || L <= c_par.edge_end);
            uint32_t skip_mask = 0;       // lists whose every entry sits at the same index of an earlier same-strand list
            for (uint32_t l = 0; l < nl && !err; l++) {
                if ((skip_mask >> l) & 1u) continue;    // all its entries are absorbed (basecallers repeat one delta list per code)
                const uint32_t n = T.n_delta[l];
                const uint32_t* Pl = P + T.ent_off[l];
                const uint32_t st = T.strand[l];
                const uint8_t lfb = T.base[l];
                const int lb = lfb == 'A' ? 0 : lfb == 'C' ? 1 : lfb == 'G' ? 2 : (lfb == 'T' || lfb == 'U') ? 3 : 4;
                uint32_t lnext = 0xffffffffu;           // nearest later list of the same strand with entries
                for (uint32_t l2 = nl; l2-- > l + 1;) if (T.strand[l2] == st && T.n_delta[l2]) lnext = l2;
                uint32_t cnt_next = 0;
                const uint32_t calls_before = n_calls;
                for (uint32_t j0 = 0; j0 < n; j0 += 32) {
                    uint32_t j = j0 + lane;
                    bool active = j < n;
                    bool matched_next = false;
                    bool emit = false;
                    uint32_t rpos = 0, info = 0;
                    bool e2 = false;
                    bool hist_ok = false;
                    float hist_v = 0.f;
                    int hist_base = 0;
                    if (active) {
                        uint32_t f = Pl[j];
                        uint32_t q = rev ? L - 1u - f : f;
                        // forward-read base at f: the list's own base (the position was selected as one of its occurrences);
                        // only 'N' lists have to look at the sequence (mod_bam.rs:1245)
                        int b = lb;
                        if (!FAST_ONLY && lb > 3) {
                            const int nb = nib_to_base(seq_nibble(seq, q));
                            b = nb > 3 ? 4 : (rev ? 3 - nb : nb);
                            if (b > 3) e2 = true;
                        }
                        // ExplicitConflictInferred (mod_bam.rs:629-634): an implicit list of this (strand, base) table
                        // that does not list f explicitly holds an inferred entry there
                        if (!FAST_ONLY && b <= 3) {
                            uint32_t others = imp_lists[st][b] & ~(1u << l);
                            while (others && !e2) {
                                const uint32_t l2 = __ffs(others) - 1;
                                others &= others - 1;
                                const uint32_t* P2 = P + T.ent_off[l2];
                                const uint32_t n2 = T.n_delta[l2];
                                if (j < n2 && P2[j] == f) continue;
                                const uint32_t k = lower_bound_u32(P2, n2, f);
                                if (!(k < n2 && P2[k] == f)) e2 = true;
                            }
                        }
                        // absorbed by an earlier list of the same strand?
                        bool owner = true;
                        for (uint32_t l2 = 0; l2 < l && owner; l2++) {
                            if (T.strand[l2] != st || T.n_delta[l2] == 0) continue;
                            const uint32_t* P2 = P + T.ent_off[l2];
                            uint32_t n2 = T.n_delta[l2];
                            if (j < n2 && P2[j] == f) { owner = false; break; }
                            uint32_t k = lower_bound_u32(P2, n2, f);
                            if (k < n2 && P2[k] == f) owner = false;
                        }
                        if (owner && !e2) {
                            const int tb = st == 0 ? b : 3 - b;
                            uint32_t mask = 0, ccode = 0;      // observed states, called code
                            int kind = 0;                      // 0 filtered, 1 canonical, 2 modified
                            float hv = 0.f;                    // arg-max value (histogram mode)
                            if (FAST_ONLY || fast_read) {
                                // ---- register-only path: at most two codes at the position (one list with <= 2 codes, or
                                //      two single-code lists of the strand). Same arithmetic as the generic path below.
                                uint32_t c0 = T.code[l][0], c1 = 0;
                                float p0, p1 = 0.f;
                                int n2c = 1;
                                {
                                    const uint8_t* mlq = ml + T.ml_off[l] + (size_t)j * T.ncodes[l];
                                    p0 = __fdiv_rn(__fadd_rn((float)mlq[0], 0.5f), 256.0f);
                                    if (T.ncodes[l] == 2) {
                                        const float pb = __fdiv_rn(__fadd_rn((float)mlq[1], 0.5f), 256.0f);
                                        const uint32_t cb = T.code[l][1];
                                        if (cb == c0) { if (__fadd_rn(p0, pb) > 1.01f) e2 = true; else p0 = __fadd_rn(p0, pb); }
                                        else { c1 = cb; p1 = pb; n2c = 2; }
                                    }
                                }
                                if (lnext != 0xffffffffu && !e2) {       // the strand's other list (single code by construction)
                                    const uint32_t* P2 = P + T.ent_off[lnext];
                                    const uint32_t nn = T.n_delta[lnext];
                                    uint32_t k = j;
                                    const uint8_t* mln = ml + T.ml_off[lnext];
                                    const uint32_t p2j = j < nn ? P2[j] : 0xffffffffu;
                                    uint8_t qn = j < nn ? mln[j] : 0;          // same index in the other list: the usual case
                                    bool found = j < nn && p2j == f;
                                    if (found) matched_next = true;
                                    else { k = lower_bound_u32(P2, nn, f); found = k < nn && P2[k] == f; if (found) qn = mln[k]; }
                                    if (found) {
                                        const float pn = __fdiv_rn(__fadd_rn((float)qn, 0.5f), 256.0f);
                                        const uint32_t cn = T.code[lnext][0];
                                        if (cn == c0) p0 = __fadd_rn(p0, pn);
                                        else { c1 = cn; p1 = pn; n2c = 2; }
                                        if ((n2c == 2 ? __fadd_rn(p0, p1) : p0) > 1.01f) e2 = true;   // combine_checked
                                    }
                                }
                                if (!e2) {
                                    if (c_par.numeric_mode == 2) {       // ReDistribute (mod_bam.rs:558-600)
                                        const uint32_t drop = c_par.collapse_code;
                                        if (n2c == 1) {
                                            if (c0 == drop) n2c = 0;
                                            else p0 = __fadd_rn(p0, __fdiv_rn(0.f, 2.0f));
                                        } else if (c0 == drop) { c0 = c1; p0 = __fadd_rn(p1, __fdiv_rn(p0, 2.0f)); n2c = 1; }
                                        else if (c1 == drop) { p0 = __fadd_rn(p0, __fdiv_rn(p1, 2.0f)); n2c = 1; }
                                        else { const float sh = __fdiv_rn(0.f, 3.0f); p0 = __fadd_rn(p0, sh); p1 = __fadd_rn(p1, sh); }
                                    }
                                    // FxHashMap iteration order of the (<= 2) codes
                                    if (n2c == 2) {
                                        const uint32_t b0 = bucket4(c0), b1 = bucket4(c1);
                                        if (b0 != b1 ? (b1 < b0) : (b0 == 3)) { const uint32_t tc = c0; c0 = c1; c1 = tc; const float tp = p0; p0 = p1; p1 = tp; }
                                    }
                                    const float sum = n2c == 0 ? 0.f : n2c == 1 ? __fadd_rn(0.f, p0) : __fadd_rn(__fadd_rn(0.f, p0), p1);
                                    const float cp = __fsub_rn(1.0f, sum);
                                    if (MODE == MODE_PILEUP) {
                                        if (n2c >= 1) mask |= 1u << state_id(C, scache, tb, c0);
                                        if (n2c == 2) mask |= 1u << state_id(C, scache, tb, c1);
                                        const float base_thr = c_par.base_set[tb] ? c_par.base_thr[tb] : c_par.default_thr;
                                        bool have = false;
                                        float best = 0.f;
                                        for (int k2 = 0; k2 < n2c; k2++) {
                                            const uint32_t cc = k2 == 0 ? c0 : c1;
                                            const float pp = k2 == 0 ? p0 : p1;
                                            float thr = base_thr;
                                            if (c_par.n_mod_thr) {
                                                const uint32_t any_code = (uint32_t)("ACGT"[tb]);
                                                bool fnd = false;
                                                for (uint32_t t = 0; t < c_par.n_mod_thr && !fnd; t++) if (c_par.mod_code[t] == cc) { thr = c_par.mod_thr[t]; fnd = true; }
                                                for (uint32_t t = 0; t < c_par.n_mod_thr && !fnd; t++) if (c_par.mod_code[t] == any_code) { thr = c_par.mod_thr[t]; fnd = true; }
                                            }
                                            if (pp >= thr && (!have || pp >= best)) { have = true; best = pp; kind = 2; ccode = cc; }
                                        }
                                        if (cp >= base_thr && (!have || cp >= best)) { have = true; kind = 1; }
                                        if (!have) kind = 0;
                                    } else {
                                        float mp = 0.f;
                                        bool hm2 = false;
                                        if (n2c >= 1) { mp = p0; hm2 = true; }
                                        if (n2c == 2 && p1 >= mp) mp = p1;
                                        hv = (hm2 && mp > cp) ? mp : cp;
                                        if (C.summ) {
                                            // thresholded call (as in the pileup) and arg-max call (last maximum wins, mod_bam.rs:489-505)
                                            uint32_t s0 = 0, s1 = 0;
                                            if (n2c >= 1) { s0 = (uint32_t)state_id(C, scache, tb, c0); mask |= 1u << s0; }
                                            if (n2c == 2) { s1 = (uint32_t)state_id(C, scache, tb, c1); mask |= 1u << s1; }
                                            const float base_thr = c_par.base_set[tb] ? c_par.base_thr[tb] : c_par.default_thr;
                                            bool have = false;
                                            float best = 0.f;
                                            uint32_t tk = 0, ts = 0;
                                            for (int k2 = 0; k2 < n2c; k2++) {
                                                const uint32_t cc = k2 == 0 ? c0 : c1;
                                                const float pp = k2 == 0 ? p0 : p1;
                                                float thr = base_thr;
                                                if (c_par.n_mod_thr) {
                                                    const uint32_t any_code = (uint32_t)("ACGT"[tb]);
                                                    bool fnd = false;
                                                    for (uint32_t t = 0; t < c_par.n_mod_thr && !fnd; t++) if (c_par.mod_code[t] == cc) { thr = c_par.mod_thr[t]; fnd = true; }
                                                    for (uint32_t t = 0; t < c_par.n_mod_thr && !fnd; t++) if (c_par.mod_code[t] == any_code) { thr = c_par.mod_thr[t]; fnd = true; }
                                                }
                                                if (pp >= thr && (!have || pp >= best)) { have = true; best = pp; tk = 2; ts = k2 == 0 ? s0 : s1; }
                                            }
                                            if (cp >= base_thr && (!have || cp >= best)) { have = true; tk = 1; }
                                            if (!have) tk = 0;
                                            const uint32_t as = (n2c == 2 && p1 >= p0) ? s1 : s0;
                                            const uint32_t ak = (hm2 && mp > cp) ? 2u : 1u;
                                            ccode = tk | (ts << 2) | (ak << 8) | (as << 10);     // summary record of this value
                                        }
                                    }
                                }
                            } else {
                            Items m;
                            m.n = 0;
                            int ord[MAX_MAP];
                            const uint8_t* mlq = ml + T.ml_off[l] + (size_t)j * T.ncodes[l];
                            for (uint32_t c = 0; c < T.ncodes[l] && !e2; c++) {
                                const float p = __fdiv_rn(__fadd_rn((float)mlq[c], 0.5f), 256.0f);
                                const int it = m.find(T.code[l][c]);
                                if (it < 0) { if (!m.push(T.code[l][c], p)) { atomicOr(C.err, MKP_DERR_TOO_MANY_CODES); e2 = true; } }
                                else { if (__fadd_rn(m.p[it], p) > 1.01f) e2 = true; else m.p[it] = __fadd_rn(m.p[it], p); }
                            }
                            for (uint32_t l2 = l + 1; l2 < nl && !e2; l2++) {
                                if (T.strand[l2] != st || T.n_delta[l2] == 0) continue;
                                const uint32_t* P2 = P + T.ent_off[l2];
                                const uint32_t n2 = T.n_delta[l2];
                                uint32_t k;
                                if (j < n2 && P2[j] == f) { k = j; if (l2 == lnext) matched_next = true; }
                                else { k = lower_bound_u32(P2, n2, f); if (!(k < n2 && P2[k] == f)) continue; }
                                // per-list table first (add_base_mod_prob), then combine_checked into the aggregate
                                Items t2;
                                t2.n = 0;
                                const uint8_t* ml2 = ml + T.ml_off[l2] + (size_t)k * T.ncodes[l2];
                                for (uint32_t c = 0; c < T.ncodes[l2] && !e2; c++) {
                                    const float p = __fdiv_rn(__fadd_rn((float)ml2[c], 0.5f), 256.0f);
                                    const int it = t2.find(T.code[l2][c]);
                                    if (it < 0) { if (!t2.push(T.code[l2][c], p)) { atomicOr(C.err, MKP_DERR_TOO_MANY_CODES); e2 = true; } }
                                    else { if (__fadd_rn(t2.p[it], p) > 1.01f) e2 = true; else t2.p[it] = __fadd_rn(t2.p[it], p); }
                                }
                                int o2[MAX_MAP];
                                iter_order(t2, o2);
                                for (int s2 = 0; s2 < t2.n && !e2; s2++) {
                                    const int i2 = o2[s2];
                                    const int it = m.find(t2.code[i2]);
                                    if (it < 0) { if (!m.push(t2.code[i2], t2.p[i2])) { atomicOr(C.err, MKP_DERR_TOO_MANY_CODES); e2 = true; } }
                                    else m.p[it] = __fadd_rn(m.p[it], t2.p[i2]);
                                }
                                if (!e2) { iter_order(m, ord); if (items_sum(m, ord) > 1.01f) e2 = true; }
                            }
                            if (!e2) {
                                iter_order(m, ord);
                                Items mc;
                                const Items* use = &m;
                                if (c_par.numeric_mode == 2) { redistribute_items(m, ord, c_par.collapse_code, mc); iter_order(mc, ord); use = &mc; }
                                if (MODE == MODE_PILEUP) {
                                    for (int k2 = 0; k2 < use->n; k2++) mask |= 1u << state_id(C, scache, tb, use->code[k2]);
                                    kind = make_call_items(*use, ord, tb, &ccode);
                                } else {
                                    hv = argmax_items(*use, ord);
                                    if (C.summ) {
                                        uint32_t tcode = 0;
                                        const int tkind = make_call_items(*use, ord, tb, &tcode);
                                        uint32_t ts = 0, as = 0;
                                        if (tkind == 2) ts = (uint32_t)state_id(C, scache, tb, tcode);
                                        const float cp2 = __fsub_rn(1.0f, items_sum(*use, ord));
                                        bool hm = false;
                                        float mp2 = 0.f;
                                        uint32_t acode = 0;
                                        for (int k2 = 0; k2 < use->n; k2++) {
                                            mask |= 1u << state_id(C, scache, tb, use->code[k2]);
                                            const float v = use->p[ord[k2]];
                                            if (!hm || v >= mp2) { hm = true; mp2 = v; acode = use->code[ord[k2]]; }
                                        }
                                        const uint32_t ak = (hm && mp2 > cp2) ? 2u : 1u;
                                        if (ak == 2) as = (uint32_t)state_id(C, scache, tb, acode);
                                        ccode = (uint32_t)tkind | (ts << 2) | (ak << 8) | (as << 10);
                                    }
                                }
                            }
                            }
                            if (!e2) {
                                bool keep = trim_ok;
                                if (keep && c_par.edge_on) {
                                    if (c_par.edge_inv) keep = f < c_par.edge_start || f >= L - c_par.edge_end;
                                    else keep = f >= c_par.edge_start && f < L - c_par.edge_end;
                                }
                                if (keep) {
                                    // aligned?  the op holding q = largest i with cq[i] >> 1 <= q, narrowed by the table
                                    bool aligned = false;
                                    if (nc > 0 && q < q_total) {
                                        uint32_t lo = 0, hi = nc;
                                        if (use_qt) { const uint32_t t = q >> qt_shift; lo = qt[t]; hi = (uint32_t)qt[t + 1] + 1u; }
                                        const uint32_t key = 2u * q + 1u;
                                        while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (cq[mid] <= key) lo = mid; else hi = mid; }
                                        const uint32_t e = cq[lo];
                                        if ((e & 1u) && q < (cq[lo + 1] >> 1)) { aligned = true; rpos = cr[lo] + (q - (e >> 1)); }
                                    }
                                    if (MODE == MODE_PILEUP) {
                                        // (mod strand, read orientation) -> reference strand (read_cache.rs:181-188)
                                        if ((st == 0) != rev) pos_mask |= mask; else neg_mask |= mask;
                                        table_survived = true;
                                        // positions outside the focus set (motif / include-bed) are dropped when the hot bitmap
                                        // is ranked (k_block_popc) and skipped by k_count_calls: no bitmap lookup here
                                        if (aligned && rpos >= C.cs && rpos < C.ce) {
                                            const uint32_t state = kind == 0 ? 0u : kind == 1 ? 1u : 2u + (uint32_t)state_id(C, scache, tb, ccode);
                                            // does a '+' list cover the same position (both pos_call and neg_call present)?
                                            uint32_t nosub = 0;
                                            if (st == 1) {
                                                for (uint32_t l2 = 0; l2 < nl && !nosub; l2++) {
                                                    if (T.strand[l2] != 0 || T.n_delta[l2] == 0) continue;
                                                    const uint32_t* P2 = P + T.ent_off[l2];
                                                    const uint32_t k = lower_bound_u32(P2, T.n_delta[l2], f);
                                                    if (k < T.n_delta[l2] && P2[k] == f) nosub = 1;
                                                }
                                            }
                                            emit = true;
                                            info = st | ((uint32_t)b << 1) | (state << 3) | (nosub << 11);
                                        }
                                    } else {
                                        bool pass = aligned || C.hist_include_unaligned;
                                        if (C.focus_pos) {
                                            // --include-bed: keep only aligned positions inside the BED on the matching reference
                                            // strand (read_ids_to_base_mod_probs.rs:1020-1047); chunk bitmaps hold the BED
                                            pass = false;
                                            if (aligned && rpos >= C.cs && rpos < C.ce) {
                                                const uint32_t x = rpos - C.cs;
                                                const uint32_t* bm = ((st == 0) == rev) ? C.focus_neg : C.focus_pos;
                                                pass = (bm[x >> 5] >> (x & 31)) & 1u;
                                            }
                                        }
                                        if (pass) {
                                            hist_ok = true; hist_v = hv; hist_base = tb;
                                            if (C.summ) { hist_v = __uint_as_float(ccode); sum_obs[tb] |= mask; }
                                            if (!FAST_ONLY && imp_lists[st][b]) imp_explicit[st * 4 + b]++;
                                        }
                                    }
                                }
                            }
                        }
                    }
                    err = __any_sync(FULL, e2) || err;
                    cnt_next += __popc(__ballot_sync(FULL, matched_next));
                    if (MODE == MODE_PILEUP) {
                        uint32_t em = __ballot_sync(FULL, emit);
                        if (emit) {
                            calls[n_calls + __popc(em & ((1u << lane) - 1u))] = make_uint2(rpos, info);
                            // marks of a read that fails later stay behind: a hot position without calls yields no rows
                            const uint32_t x = rpos - C.cs;
                            if (!C.list_mode) atomicOr(&C.hot[x >> 5], 1u << (x & 31));
                        }
                        n_calls += __popc(em);
                    } else {
                        uint32_t hm = __ballot_sync(FULL, hist_ok);
                        if (hist_ok) {
                            // stash (bin | base<<16 | inexact<<20) in the call buffer; committed after validation
                            float sc = __fmul_rn(C.summ ? 0.f : hist_v, 1024.0f);
                            int bin = (int)rintf(sc);
                            uint32_t inexact = ((float)bin != sc || bin < 0 || bin > 1024) ? 1u : 0u;
                            if (bin < 0) bin = 0;
                            if (bin > 1024) bin = 1024;
                            calls[n_hist + __popc(hm & ((1u << lane) - 1u))] = make_uint2((uint32_t)bin | ((uint32_t)hist_base << 16) | (inexact << 20), __float_as_uint(hist_v));
                        }
                        n_hist += __popc(hm);
                    }
                }
                if (n_calls != calls_before) n_runs++;
                if (!imp_any && lnext != 0xffffffffu && T.base[lnext] != 'N' && T.n_delta[lnext] == cnt_next) skip_mask |= 1u << lnext;
            }
        }
        // ---- phase 4b: implicit tables (every other occurrence of the base is an inferred-canonical entry) ----
        uint32_t imp_meta[2] = {0, 0};
        unsigned long long imp_hist[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        if (!FAST_ONLY && !err) {
            const bool trim_ok = !c_par.edge_on || !(L <= c_par.edge_start || L <= c_par.edge_end);
            for (uint32_t sb = 0; sb < 8; sb++) {
                const uint32_t s = sb >> 2, b = sb & 3;
                const uint32_t lists = imp_lists[s][b];
                if (!lists || !trim_ok) continue;
                const uint32_t xn = 1u << (rev ? 3 - b : b);          // SEQ nibble of the forward base b
                // survives the edge filter?  (any occurrence of the base with a kept forward position)
                bool any = false;
                for (uint32_t q = lane; q < L && !any; q += 32) {
                    if (seq_nibble(seq, q) != xn) continue;
                    const uint32_t f = rev ? L - 1u - q : q;
                    bool keep = true;
                    if (c_par.edge_on) keep = c_par.edge_inv ? (f < c_par.edge_start || f >= L - c_par.edge_end) : (f >= c_par.edge_start && f < L - c_par.edge_end);
                    any = keep;
                }
                if (!__any_sync(FULL, any)) continue;
                Items m, mc;
                m.n = 0;
                int ord[MAX_MAP];
                for (uint32_t l = 0; l < nl; l++) if ((lists >> l) & 1u) for (uint32_t c = 0; c < T.ncodes[l]; c++) if (m.find(T.code[l][c]) < 0) m.push(T.code[l][c], 0.f);
                iter_order(m, ord);
                const Items* use = &m;
                if (c_par.numeric_mode == 2) { redistribute_items(m, ord, c_par.collapse_code, mc); iter_order(mc, ord); use = &mc; }
                const int tb = s == 0 ? (int)b : 3 - (int)b;
                table_survived = true;
                if (MODE == MODE_PILEUP) {
                    uint32_t mask = 0;
                    for (int k = 0; k < use->n; k++) mask |= 1u << state_id(C, scache, tb, use->code[k]);
                    if ((s == 0) != rev) pos_mask |= mask; else neg_mask |= mask;
                    uint32_t code = 0;
                    const int kind = make_call_items(*use, ord, tb, &code);
                    const uint32_t state = kind == 0 ? 0u : kind == 1 ? 1u : 2u + (uint32_t)state_id(C, scache, tb, code);
                    imp_meta[s] |= (0x80u | state) << (8 * b);
                } else {
                    if (C.summ) for (int k = 0; k < use->n; k++) sum_obs[tb] |= 1u << state_id(C, scache, tb, use->code[k]);
                    // values: argmax of an all-zero map = canonical probability 1.0 for every passing inferred position
                    unsigned long long n_pass = 0;
                    if (C.hist_include_unaligned && !C.focus_pos) {
                        for (uint32_t q = lane; q < L; q += 32) {
                            if (seq_nibble(seq, q) != xn) continue;
                            const uint32_t f = rev ? L - 1u - q : q;
                            bool keep = true;
                            if (c_par.edge_on) keep = c_par.edge_inv ? (f < c_par.edge_start || f >= L - c_par.edge_end) : (f >= c_par.edge_start && f < L - c_par.edge_end);
                            if (keep) n_pass++;
                        }
                    } else {
                        uint32_t qc2 = 0, rc2 = (uint32_t)h.ref_start;
                        const uint32_t* bm = C.focus_pos ? (((s == 0) == rev) ? C.focus_neg : C.focus_pos) : nullptr;
                        for (uint32_t i = 0; i < h.n_cigar; i++) {
                            const uint32_t c = cig[i], op = c & 15, len = c >> 4;
                            if (op == 0 || op == 7 || op == 8) {
                                for (uint32_t k = lane; k < len; k += 32) {
                                    const uint32_t q = qc2 + k;
                                    if (q >= L || seq_nibble(seq, q) != xn) continue;
                                    const uint32_t f = rev ? L - 1u - q : q;
                                    bool keep = true;
                                    if (c_par.edge_on) keep = c_par.edge_inv ? (f < c_par.edge_start || f >= L - c_par.edge_end) : (f >= c_par.edge_start && f < L - c_par.edge_end);
                                    if (keep && bm) {
                                        const uint32_t r = rc2 + k;
                                        keep = r >= C.cs && r < C.ce && ((bm[(r - C.cs) >> 5] >> ((r - C.cs) & 31)) & 1u);
                                    }
                                    if (keep) n_pass++;
                                }
                            }
                            if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) qc2 += len;
                            if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rc2 += len;
                        }
                    }
                    n_pass = __reduce_add_sync(FULL, (uint32_t)n_pass);
                    const uint32_t n_exp = __reduce_add_sync(FULL, imp_explicit[sb]);
                    imp_hist[sb] = n_pass > n_exp ? n_pass - n_exp : 0;
                }
            }
        }
        pos_mask = __reduce_or_sync(FULL, pos_mask);
        neg_mask = __reduce_or_sync(FULL, neg_mask);
        table_survived = __any_sync(FULL, table_survived);
        __syncwarp();
        // ---- phase 5: commit -------------------------------------------------------------------------
        if (MODE == MODE_PILEUP) {
            if (!err && table_survived) {
                meta.flags |= 2;
                if (n_runs <= 1 && !(imp_meta[0] | imp_meta[1])) meta.flags |= 4;     // one sorted run: k_count_bases merges it
                meta.pos_mask = pos_mask; meta.neg_mask = neg_mask; meta.n_calls = n_calls;
                meta.imp[0] = imp_meta[0]; meta.imp[1] = imp_meta[1];
                if (lane == 0 && n_calls) atomicAdd(C.total_calls, (unsigned long long)n_calls);
                if (!FAST_ONLY && (imp_meta[0] | imp_meta[1]) && !C.list_mode) {
                    // every aligned occurrence of an implicit table's base is a call position
                    uint32_t qc2 = 0, rc2 = (uint32_t)h.ref_start;
                    for (uint32_t i = 0; i < h.n_cigar; i++) {
                        const uint32_t c = cig[i], op = c & 15, len = c >> 4;
                        if (op == 0 || op == 7 || op == 8) {
                            for (uint32_t k = lane; k < len; k += 32) {
                                const uint32_t q = qc2 + k, r = rc2 + k;
                                if (q >= L || r < C.cs || r >= C.ce) continue;
                                const int nb = nib_to_base(seq_nibble(seq, q));
                                if (nb > 3) continue;
                                const uint32_t b = rev ? 3 - nb : nb;
                                if (!(((imp_meta[0] | imp_meta[1]) >> (8 * b)) & 0x80u)) continue;
                                const uint32_t f = rev ? L - 1u - q : q;
                                if (c_par.edge_on && !(c_par.edge_inv ? (f < c_par.edge_start || f >= L - c_par.edge_end) : (f >= c_par.edge_start && f < L - c_par.edge_end))) continue;
                                const uint32_t x = r - C.cs;
                                if (C.focus_pos && !(((C.focus_pos[x >> 5] | C.focus_neg[x >> 5]) >> (x & 31)) & 1u)) continue;
                                atomicOr(&C.hot[x >> 5], 1u << (x & 31));
                            }
                        }
                        if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) qc2 += len;
                        if (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) rc2 += len;
                    }
                }
            }
        } else {
            if (!err && any_entries) {
                unsigned long long extra = 0;
                for (int sb = 0; sb < 8; sb++) extra += imp_hist[sb];
                meta.n_hist = n_hist + (uint32_t)(extra < 0x7fffffffull ? extra : 0x7fffffffull);
                if (C.summ) {
                    uint32_t seen = 0;
                    for (uint32_t k = lane; k < n_hist; k += 32) {
                        const uint2 v = calls[k];
                        const uint32_t tb = (v.x >> 16) & 3u, tk = v.y & 3u, ts = (v.y >> 2) & 63u, ak = (v.y >> 8) & 3u, as = (v.y >> 10) & 63u;
                        seen |= 1u << tb;
                        // pass: the thresholded call; fail: filed under the arg-max call (src/summarize.rs:186-214)
                        if (tk == 1) atomicAdd(&C.summ[(tb * 2 + 0) * 33 + 0], 1ull);
                        else if (tk == 2) atomicAdd(&C.summ[(tb * 2 + 0) * 33 + 1 + ts], 1ull);
                        else if (ak == 1) atomicAdd(&C.summ[(tb * 2 + 1) * 33 + 0], 1ull);
                        else atomicAdd(&C.summ[(tb * 2 + 1) * 33 + 1 + as], 1ull);
                    }
                    seen = __reduce_or_sync(FULL, seen);
                    uint32_t obs[4];
                    for (int b = 0; b < 4; b++) obs[b] = __reduce_or_sync(FULL, sum_obs[b]);
                    if (lane == 0) {
                        for (int sb = 0; sb < 8; sb++) if (imp_hist[sb]) {
                            const int tb = (sb >> 2) == 0 ? (sb & 3) : 3 - (sb & 3);
                            // an inferred entry is the all-zero map: canonical with probability 1.0
                            const float base_thr = c_par.base_set[tb] ? c_par.base_thr[tb] : c_par.default_thr;
                            atomicAdd(&C.summ[(tb * 2 + (1.0f >= base_thr ? 0 : 1)) * 33 + 0], imp_hist[sb]);
                            seen |= 1u << tb;
                        }
                        for (int b = 0; b < 4; b++) {
                            if ((seen >> b) & 1u) atomicAdd(&C.summ[4 * 2 * 33 + b], 1ull);
                            if (obs[b]) atomicOr((uint32_t*)(C.summ + 4 * 2 * 33 + 4) + b, obs[b]);
                        }
                    }
                }
                if (C.hist) {
                    for (uint32_t k = lane; k < n_hist; k += 32) {
                        uint32_t v = calls[k].x;
                        atomicAdd(&C.hist[((v >> 16) & 3u) * 1025u + (v & 0xffffu)], 1ull);
                        if ((v >> 20) & 1u) atomicAdd(C.hist_inexact, 1ull);
                    }
                    if (lane == 0) for (int sb = 0; sb < 8; sb++) if (imp_hist[sb]) {
                        const int tb = (sb >> 2) == 0 ? (sb & 3) : 3 - (sb & 3);
                        atomicAdd(&C.hist[tb * 1025u + 1024u], imp_hist[sb]);
                    }
                }
            }
        }
        if (lane == 0) C.meta[ri] = meta;
        __syncwarp();
    }
}

// ---- rank over the hot bitmap -----------------------------------------------------------------------
// (marks outside the focus set, which k_resolve does not look at, are cleared here)
__global__ void k_block_popc(uint32_t* __restrict__ bits, uint32_t n_words, uint32_t* __restrict__ block_sums,
                             const uint32_t* __restrict__ focus_pos, const uint32_t* __restrict__ focus_neg) {
    __shared__ uint32_t s[32];
    uint32_t i = blockIdx.x * 1024 + threadIdx.x;
    uint32_t v = 0;
    if (i < n_words) {
        uint32_t word = bits[i];
        if (focus_pos && word) { const uint32_t keep = word & (focus_pos[i] | focus_neg[i]); if (keep != word) bits[i] = keep; word = keep; }
        v = __popc(word);
    }
    uint32_t w = __reduce_add_sync(FULL, v);
    if (lane_id() == 0) s[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x < 32) { uint32_t t = __reduce_add_sync(FULL, s[threadIdx.x]); if (threadIdx.x == 0) block_sums[blockIdx.x] = t; }
}
// single block: exclusive scan of block_sums in place, total -> *total
__global__ void k_scan_blocks(uint32_t* block_sums, uint32_t n, uint32_t* total) {
    __shared__ uint32_t s_carry;
    __shared__ uint32_t s_w[32];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint32_t b = 0; b < n; b += 1024) {
        uint32_t i = b + threadIdx.x;
        uint32_t v = i < n ? block_sums[i] : 0;
        uint32_t inc = warp_incl_scan(v);
        if (lane_id() == 31) s_w[threadIdx.x >> 5] = inc;
        __syncthreads();
        if (threadIdx.x < 32) { uint32_t t = s_w[threadIdx.x]; uint32_t ti = warp_incl_scan(t); s_w[threadIdx.x] = ti - t; }
        __syncthreads();
        uint32_t ex = s_carry + s_w[threadIdx.x >> 5] + inc - v;
        if (i < n) block_sums[i] = ex;
        __syncthreads();
        if (threadIdx.x == 1023) s_carry = ex + v;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = s_carry;
}
__global__ void k_word_prefix(const uint32_t* __restrict__ bits, uint32_t n_words, const uint32_t* __restrict__ block_sums, uint32_t* __restrict__ prefix) {
    __shared__ uint32_t s_w[32];
    uint32_t i = blockIdx.x * 1024 + threadIdx.x;
    uint32_t v = i < n_words ? __popc(bits[i]) : 0;
    uint32_t inc = warp_incl_scan(v);
    if (lane_id() == 31) s_w[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) { uint32_t t = s_w[threadIdx.x]; uint32_t ti = warp_incl_scan(t); s_w[threadIdx.x] = ti - t; }
    __syncthreads();
    if (i < n_words) prefix[i] = block_sums[blockIdx.x] + s_w[threadIdx.x >> 5] + inc - v;
}

struct CountDev {
    const mkp_read_hdr* hdrs;
    const uint8_t* heap;
    const ReadMeta* meta;
    const uint2* calls;
    uint32_t n_reads, cs, ce;
    const uint32_t* focus_pos; const uint32_t* focus_neg;
    const uint32_t* hot; const uint32_t* hot_prefix;
    uint32_t* slots;
    uint32_t stride;    // words per slot
    uint32_t n_states;
    uint32_t n_words;
    uint32_t* obs_word;   // [2][n_words] observed-code masks of fully covered bitmap words
    uint32_t* work;       // dynamic read counter
    const uint32_t* list; // list mode: the reads to count (k_pileup_fused counted the others) and their number
    const uint32_t* list_count;
    const uint32_t* order; // else: processing order (longest reads first), or null
};

__device__ __forceinline__ uint32_t slot_of(const CountDev& D, uint32_t x) {
    uint32_t w = x >> 5;
    return D.hot_prefix[w] + __popc(D.hot[w] & ((1u << (x & 31)) - 1u));
}

// FeatureVector::add_feature + Tally::add_feature (src/pileup/mod.rs:176-193, 238-281) as counter updates
__device__ __forceinline__ void add_feature(uint32_t* S, uint32_t n_states, uint32_t t, uint32_t pb, uint32_t state, bool ok_pos, bool ok_neg, uint32_t delta) {
    if (!(t == 0 ? ok_pos : ok_neg)) return;
    if (state == 0) atomicAdd(&S[SL_FILT + t], delta);
    else if (state == 1) atomicAdd(&S[SL_CANON + t * 4 + pb], delta);
    else atomicAdd(&S[SL_MOD + t * n_states + (state - 2)], delta);
}

// warp per read, lanes over its call records
__global__ void __launch_bounds__(256) k_count_calls(CountDev D) {
    const uint32_t lane = lane_id();
    for (;;) {
        uint32_t ri = 0;
        if (lane == 0) { ri = atomicAdd(D.work + 1, 1u); if (D.list) ri = ri < *D.list_count ? D.list[ri] : 0xffffffffu; else if (D.order && ri < D.n_reads) ri = D.order[ri]; }
        ri = __shfl_sync(FULL, ri, 0);
        if (ri >= D.n_reads) break;
        const ReadMeta m = D.meta[ri];
        if (!(m.flags & 2) || m.n_calls == 0) continue;
#ifndef MKP_NO_FUSED_CALLS
        if (m.flags & 4) continue;                 // one sorted run of records: counted by k_count_bases while it streams them
#endif
        const uint32_t a = (D.hdrs[ri].flags & 0x10) ? 1u : 0u;
        const uint2* calls = D.calls + m.entry_off;
        for (uint32_t k = lane; k < m.n_calls; k += 32) {
            uint2 c = calls[k];
            uint32_t x = c.x - D.cs;
            uint32_t info = c.y;
            uint32_t st = info & 1u, b = (info >> 1) & 3u, state = (info >> 3) & 0xffu, nosub = (info >> 11) & 1u;
            uint32_t fp = FULL, fn = FULL;
            if (D.focus_pos) { fp = D.focus_pos[x >> 5]; fn = D.focus_neg[x >> 5]; }
            bool ok_pos = (fp >> (x & 31)) & 1u, ok_neg = (fn >> (x & 31)) & 1u;
            const uint32_t hw = D.hot[x >> 5];
            if (!((hw >> (x & 31)) & 1u)) continue;            // outside the focus set
            uint32_t* S = D.slots + (size_t)(D.hot_prefix[x >> 5] + __popc(hw & ((1u << (x & 31)) - 1u))) * D.stride;
            // cancel what k_count_bases adds for this (read, position): the inferred entry of this table if the
            // table is implicit, else the NoCall base on tally[a] (unless the other strand's table is implicit
            // there, in which case no NoCall is counted, or the '+' record of a +/- pair already cancels it)
            const uint32_t imp_own = (m.imp[st] >> (8 * b)) & 0xffu, imp_other = (m.imp[1 - st] >> (8 * b)) & 0xffu;
            if (imp_own & 0x80u) add_feature(S, D.n_states, st == 0 ? a : 1u - a, st == 0 ? b : 3u - b, imp_own & 0x7fu, ok_pos, ok_neg, 0xffffffffu);
            else if (!(m.flags & 4) && !(imp_other & 0x80u) && !nosub && (a == 0 ? ok_pos : ok_neg)) atomicAdd(&S[SL_BASE + a * 4 + b], 0xffffffffu);
            add_feature(S, D.n_states, st == 0 ? a : 1u - a, st == 0 ? b : 3u - b, state, ok_pos, ok_neg, 1u);
        }
    }
}

// warp per read: every aligned base / deleted position that lands on a hot position, plus the observed-code
// coverage of the read (src/pileup/mod.rs:831-835: unioned for every alignment at the position, deletions included,
// reference skips excluded).  Coverage is recorded per 32-position bitmap word when the read covers the whole word
// (one coalesced check-then-OR per word) and per hot position only in the partial words at the ends of a run.
__global__ void __launch_bounds__(256, MKP_MINB_BASES) k_count_bases(CountDev D) {
    __shared__ uint32_t s_op[8][32], s_q[8][32], s_r[8][33];   // the current batch of 32 ops: op|len<<4, query start, reference start
    __shared__ uint32_t s_mask[8][32];                          // positions of the current 32 words where this read has a call
    const uint32_t lane = lane_id();
    const uint32_t wib = threadIdx.x >> 5;
    for (;;) {
        uint32_t ri = 0;
        if (lane == 0) { ri = atomicAdd(D.work, 1u); if (D.list) ri = ri < *D.list_count ? D.list[ri] : 0xffffffffu; else if (D.order && ri < D.n_reads) ri = D.order[ri]; }
        ri = __shfl_sync(FULL, ri, 0);
        if (ri >= D.n_reads) break;
        const ReadMeta m = D.meta[ri];
        if (!(m.flags & 1)) continue;
        const mkp_read_hdr h = D.hdrs[ri];
        if ((uint32_t)m.ref_end <= D.cs || (uint32_t)h.ref_start >= D.ce) continue;
        const uint32_t a = (h.flags & 0x10) ? 1u : 0u;
        const uint32_t* cig = (const uint32_t*)(D.heap + h.off);
        const uint8_t* seq = D.heap + h.off + 4ull * h.n_cigar;
        const bool has_mods = m.flags & 2;
        const bool has_imp = (m.imp[0] | m.imp[1]) != 0;
        const uint32_t pm = has_mods ? m.pos_mask : 0u, nm = has_mods ? m.neg_mask : 0u;
        const bool merge = has_mods && (m.flags & 4) && m.n_calls > 0;
        const uint2* rcalls = D.calls + m.entry_off;
        uint32_t cp = 0;                                        // call records consumed so far (ascending position)
        // observed-code coverage of one run [ra, rb) of reference positions
        auto cover = [&](uint32_t ra, uint32_t rb) {
            if (!(pm | nm)) return;
            if (ra < D.cs) ra = D.cs;
            if (rb > D.ce) rb = D.ce;
            if (ra >= rb) return;
            const uint32_t wl = (ra - D.cs + 31) >> 5, wh = (rb - D.cs) >> 5;   // words [wl, wh) are fully covered
            for (uint32_t w = wl + lane; w < wh; w += 32) {
                if (pm && (D.obs_word[w] & pm) != pm) atomicOr(&D.obs_word[w], pm);
                if (nm && (D.obs_word[D.n_words + w] & nm) != nm) atomicOr(&D.obs_word[D.n_words + w], nm);
            }
            // partial words: lane = bit
            auto partial = [&](uint32_t pa, uint32_t pb) {     // [pa,pb) lies inside one bitmap word
                if (pa >= pb) return;
                const uint32_t w = (pa - D.cs) >> 5, wbase = D.cs + (w << 5);
                const uint32_t word = D.hot[w];
                uint32_t bits = word & (FULL << (pa - wbase));
                if (pb < wbase + 32) bits &= (1u << (pb - wbase)) - 1u;
                if ((bits >> lane) & 1u) {
                    uint32_t* S = D.slots + (size_t)(D.hot_prefix[w] + __popc(word & ((1u << lane) - 1u))) * D.stride;
                    if (pm && (S[SL_OBS] & pm) != pm) atomicOr(&S[SL_OBS], pm);
                    if (nm && (S[SL_OBS + 1] & nm) != nm) atomicOr(&S[SL_OBS + 1], nm);
                }
            };
            if (wl > wh) partial(ra, rb);                       // the run starts and ends inside one word
            else { partial(ra, D.cs + (wl << 5)); partial(D.cs + (wh << 5), rb); }
        };
        uint32_t qc = 0, rc = (uint32_t)h.ref_start;
        uint32_t run_start = rc;                                // start of the current run without reference skips
        for (uint32_t b0 = 0; b0 < h.n_cigar; b0 += 32) {
            uint32_t i = b0 + lane;
            uint32_t c = i < h.n_cigar ? cig[i] : 0;
            uint32_t op = c & 15, len = c >> 4;
            uint32_t ql = (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) ? len : 0;
            uint32_t rl = (op == 0 || op == 2 || op == 3 || op == 7 || op == 8) ? len : 0;
            uint32_t qi = warp_incl_scan(ql), rr = warp_incl_scan(rl);
            uint32_t q0 = qc + qi - ql, r0 = rc + rr - rl;
            qc += __shfl_sync(FULL, qi, 31);
            rc += __shfl_sync(FULL, rr, 31);
            // reference skips ('N') end a covered run
            uint32_t skips = __ballot_sync(FULL, op == 3 && len > 0 && i < h.n_cigar);
            while (skips) {
                const int sl = __ffs(skips) - 1;
                skips &= skips - 1;
                const uint32_t sr = __shfl_sync(FULL, r0, sl), slen = __shfl_sync(FULL, len, sl);
                cover(run_start, sr);
                run_start = sr + slen;
            }
            // hot-bitmap words under this batch of ops, one per lane: the word is tested first (most hold no position this
            // read can add to); the op under a hot position is found by one search per word over the batch's reference
            // starts, then by stepping forward. Only M,=,X (a base) and D (a deletion) count.
            const uint32_t R0 = __shfl_sync(FULL, r0, 0), R1 = rc;
            s_op[wib][lane] = op | (len << 4); s_q[wib][lane] = q0; s_r[wib][lane] = r0;
            if (lane == 0) s_r[wib][32] = R1;
            __syncwarp();
            const uint32_t blo = R0 > D.cs ? R0 : D.cs, bhi = R1 < D.ce ? R1 : D.ce;
            if (blo < bhi) {
                const uint32_t w_first = (blo - D.cs) >> 5, w_last = (bhi - 1 - D.cs) >> 5;
                for (uint32_t wt = w_first; wt <= w_last; wt += 32) {
                    const uint32_t w = wt + lane;
                    // reads whose call records form one run sorted by position (flag 4): a position with a call needs no
                    // NoCall (k_count_calls does not cancel one for these reads), so the calls under these 32 words are
                    // streamed in (one cursor per read, ascending positions) and masked out before the per-position work
                    uint32_t called = 0;
                    if (merge) {
                        const uint32_t t_end = D.cs + ((wt + 32) << 5);
                        const uint32_t lim = bhi < t_end ? bhi : t_end;
                        s_mask[wib][lane] = 0;
                        __syncwarp();
                        for (;;) {
                            const uint32_t k = cp + lane;
                            uint32_t x = 0xffffffffu, info = 0;
                            if (k < m.n_calls) { const uint2 c = rcalls[a ? m.n_calls - 1u - k : k]; x = c.x; info = c.y; }
                            const bool take = x < lim;
                            if (take) {
                                const uint32_t rel = x - D.cs;
                                if ((rel >> 5) >= wt) atomicOr(&s_mask[wib][(rel >> 5) - wt], 1u << (rel & 31));
#ifndef MKP_NO_FUSED_CALLS
                                // the record's own counter (k_count_calls leaves these reads alone): modcall / canonical / filtered
                                const uint32_t hw = D.hot[rel >> 5];
                                if ((hw >> (rel & 31)) & 1u) {                       // else: outside the focus set
                                    uint32_t fp = FULL, fn = FULL;
                                    if (D.focus_pos) { fp = D.focus_pos[rel >> 5]; fn = D.focus_neg[rel >> 5]; }
                                    const uint32_t st = info & 1u, cb = (info >> 1) & 3u, state = (info >> 3) & 0xffu;
                                    uint32_t* S = D.slots + (size_t)(D.hot_prefix[rel >> 5] + __popc(hw & ((1u << (rel & 31)) - 1u))) * D.stride;
                                    add_feature(S, D.n_states, st == 0 ? a : 1u - a, st == 0 ? cb : 3u - cb, state, (fp >> (rel & 31)) & 1u, (fn >> (rel & 31)) & 1u, 1u);
                                }
#endif
                            }
                            const uint32_t nt = __popc(__ballot_sync(FULL, take));
                            cp += nt;
                            if (nt < 32) break;
                        }
                        __syncwarp();
                        called = s_mask[wib][lane];
                    }
                    if (w > w_last) continue;
                    const uint32_t word = D.hot[w];
                    if (!word) continue;
                    const uint32_t wbase = D.cs + (w << 5);
                    uint32_t bits = word & ~called;
                    if (blo > wbase) bits &= FULL << (blo - wbase);
                    if (bhi < wbase + 32) bits &= (1u << (bhi - wbase)) - 1u;
                    if (!bits) continue;
                    // positions whose strand rule admits what this read can add there
                    uint32_t fp = FULL, fn = FULL;
                    if (D.focus_pos) { fp = D.focus_pos[w]; fn = D.focus_neg[w]; }
                    const uint32_t ok = a == 0 ? fp : fn;
                    bits &= has_imp ? (fp | fn) : ok;
                    if (!bits) continue;
                    const uint32_t pre = D.hot_prefix[w];
                    // op under the first candidate position: largest j with s_r[j] <= r (ops without reference length share
                    // their start with the next op, so the largest index is the op that holds r)
                    uint32_t j = 0;
                    {
                        const uint32_t r = wbase + (uint32_t)__ffs(bits) - 1u;
#pragma unroll
                        for (int stp = 16; stp >= 1; stp >>= 1) if (s_r[wib][j + stp] <= r) j += stp;
                    }
                    while (bits) {
                        const uint32_t bit = __ffs(bits) - 1;
                        bits &= bits - 1;
                        const uint32_t r = wbase + bit;
                        while (s_r[wib][j + 1] <= r) j++;
                        const uint32_t jc = s_op[wib][j], jop = jc & 15;
                        if (!(jop == 0 || jop == 7 || jop == 8 || jop == 2)) continue;       // reference skip
                        if (jop == 2 && !((ok >> bit) & 1u)) continue;
                        uint32_t* S = D.slots + (size_t)(pre + __popc(word & ((1u << bit) - 1u))) * D.stride;
                        if (jop == 2) { atomicAdd(&S[SL_DEL + a], 1u); continue; }
                        const uint32_t q = s_q[wib][j] + (r - s_r[wib][j]);
                        const int nb = nib_to_base(seq_nibble(seq, q));
                        if (nb > 3) continue;
                        const uint32_t b = a ? 3 - nb : nb;
                        if (!has_imp) { atomicAdd(&S[SL_BASE + a * 4 + b], 1u); continue; }
                        uint32_t ip = (m.imp[0] >> (8 * b)) & 0xffu, in = (m.imp[1] >> (8 * b)) & 0xffu;
                        if ((ip | in) & 0x80u) {
                            // inferred-canonical entries of implicit tables, subject to the edge filter
                            const uint32_t f = a ? h.l_seq - 1u - q : q;
                            if (c_par.edge_on && !(c_par.edge_inv ? (f < c_par.edge_start || f >= h.l_seq - c_par.edge_end) : (f >= c_par.edge_start && f < h.l_seq - c_par.edge_end))) ip = in = 0;
                        }
                        const bool okp = (fp >> bit) & 1u, okn = (fn >> bit) & 1u;
                        if (!((ip | in) & 0x80u)) { if ((ok >> bit) & 1u) atomicAdd(&S[SL_BASE + a * 4 + b], 1u); continue; }
                        if (ip & 0x80u) add_feature(S, D.n_states, a, b, ip & 0x7fu, okp, okn, 1u);
                        if (in & 0x80u) add_feature(S, D.n_states, 1u - a, 3u - b, in & 0x7fu, okp, okn, 1u);
                    }
                }
            }
            __syncwarp();
        }
        cover(run_start, rc);
    }
}

struct RowDev {
    const uint32_t* hot; const uint32_t* hot_prefix;
    uint32_t n_words, cs, ce;
    const uint32_t* slots;
    uint32_t stride, n_states;
    const unsigned long long* states;
    uint32_t numeric_mode;
    const uint32_t* obs_word;
    uint32_t* row_counts;     // per hot-bitmap word
    const uint32_t* row_prefix;
    mkp_row* rows;
    uint32_t rows_cap;        // emit pass: a word whose rows do not fit is skipped (the host grows the buffer and emits again)
};

// rows of one slot, in output order. emit == nullptr => count only
__device__ __forceinline__ uint32_t slot_rows(const RowDev& R, const uint32_t* S, uint32_t pos, uint32_t w, mkp_row* out) {
    uint32_t n = 0;
    for (uint32_t s = 0; s < 2; s++) {
        uint32_t mod_by_base[4] = {0, 0, 0, 0};
        for (uint32_t id = 0; id < R.n_states; id++) mod_by_base[(uint32_t)(R.states[id] >> 32) & 3u] += S[SL_MOD + s * R.n_states + id];
        const uint32_t obs = S[SL_OBS + s] | R.obs_word[s * R.n_words + w];
        const uint32_t row0 = n;
        for (uint32_t pb = 0; pb < 4; pb++) {
            uint32_t n_can = S[SL_CANON + s * 4 + pb], total_mod = mod_by_base[pb];
            if (n_can + total_mod == 0) continue;
            uint32_t n_diff = 0;
            for (uint32_t b = 0; b < 4; b++) if (b != pb) n_diff += S[SL_BASE + s * 4 + b] + S[SL_CANON + s * 4 + b] + mod_by_base[b];
            mkp_row r;
            r.pos = pos; r.strand = s == 0 ? '+' : '-'; r.primary_base = (uint8_t)pb; r.reserved = 0;
            r.n_canon = n_can; r.n_delete = S[SL_DEL + s]; r.n_filtered = S[SL_FILT + s]; r.n_diff = n_diff; r.n_nocall = S[SL_BASE + s * 4 + pb];
            if (R.numeric_mode == 1) {
                r.code = (uint32_t)("ACGT"[pb]); r.n_mod = total_mod; r.n_other = 0;
                if (out) out[n] = r;
                n++;
            } else {
                for (uint32_t id = 0; id < R.n_states; id++) {
                    unsigned long long key = R.states[id];
                    if (((uint32_t)(key >> 32) & 3u) != pb || !((obs >> id) & 1u)) continue;
                    uint32_t n_mod = S[SL_MOD + s * R.n_states + id];
                    r.code = (uint32_t)key; r.n_mod = n_mod; r.n_other = total_mod - n_mod;
                    if (out) out[n] = r;
                    n++;
                }
            }
        }
        // stable insertion sort of this strand's rows by code (derived Ord: Code(char) < ChEbi == u32 order)
        if (out) for (uint32_t i = row0 + 1; i < n; i++) {
            mkp_row key = out[i];
            uint32_t j = i;
            while (j > row0 && out[j - 1].code > key.code) { out[j] = out[j - 1]; j--; }
            out[j] = key;
        }
    }
    return n;
}

// thread per hot-bitmap word
template <bool EMIT>
__global__ void __launch_bounds__(256) k_rows(RowDev R) {
    uint32_t w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= R.n_words) return;
    uint32_t bits = R.hot[w];
    if (!bits) { if (!EMIT) R.row_counts[w] = 0; return; }
    uint32_t slot = R.hot_prefix[w];
    uint32_t n = 0;
    if (EMIT && (unsigned long long)R.row_prefix[w] + R.row_counts[w] > (unsigned long long)R.rows_cap) return;
    mkp_row* out = EMIT ? R.rows + R.row_prefix[w] : nullptr;
    while (bits) {
        uint32_t bit = __ffs(bits) - 1;
        bits &= bits - 1;
        n += slot_rows(R, R.slots + (size_t)slot * R.stride, R.cs + (w << 5) + bit, w, EMIT ? out + n : nullptr);
        slot++;
    }
    if (!EMIT) R.row_counts[w] = n;
}

// exclusive scan of arbitrary u32 array via the same 3-kernel scheme (counts -> prefix)
__global__ void k_block_sum(const uint32_t* __restrict__ v, uint32_t n, uint32_t* __restrict__ block_sums) {
    __shared__ uint32_t s[32];
    uint32_t i = blockIdx.x * 1024 + threadIdx.x;
    uint32_t x = i < n ? v[i] : 0;
    uint32_t w = __reduce_add_sync(FULL, x);
    if (lane_id() == 0) s[threadIdx.x >> 5] = w;
    __syncthreads();
    if (threadIdx.x < 32) { uint32_t t = __reduce_add_sync(FULL, s[threadIdx.x]); if (threadIdx.x == 0) block_sums[blockIdx.x] = t; }
}
__global__ void k_value_prefix(const uint32_t* __restrict__ v, uint32_t n, const uint32_t* __restrict__ block_sums, uint32_t* __restrict__ prefix) {
    __shared__ uint32_t s_w[32];
    uint32_t i = blockIdx.x * 1024 + threadIdx.x;
    uint32_t x = i < n ? v[i] : 0;
    uint32_t inc = warp_incl_scan(x);
    if (lane_id() == 31) s_w[threadIdx.x >> 5] = inc;
    __syncthreads();
    if (threadIdx.x < 32) { uint32_t t = s_w[threadIdx.x]; uint32_t ti = warp_incl_scan(t); s_w[threadIdx.x] = ti - t; }
    __syncthreads();
    if (i < n) prefix[i] = block_sums[blockIdx.x] + s_w[threadIdx.x >> 5] + inc - x;
}

// ---- processing order: reads by descending length class (32 classes by the position of the top bit of l_seq); a counting sort,
//      the order inside a class does not matter (all updates are commutative integer atomics)
__global__ void k_order_hist(const mkp_read_hdr* __restrict__ hdrs, uint32_t n, uint32_t* __restrict__ hist) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) atomicAdd(&hist[31 - (hdrs[i].l_seq ? 31 - __clz(hdrs[i].l_seq) : 0)], 1u);
}
__global__ void k_order_scan(uint32_t* hist) {
    uint32_t acc = 0;
    for (int b = 0; b < 32; b++) { const uint32_t v = hist[b]; hist[b] = acc; acc += v; }
}
__global__ void k_order_scatter(const mkp_read_hdr* __restrict__ hdrs, uint32_t n, uint32_t* __restrict__ hist, uint32_t* __restrict__ order) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) order[atomicAdd(&hist[31 - (hdrs[i].l_seq ? 31 - __clz(hdrs[i].l_seq) : 0)], 1u)] = i;
}

}  // namespace mkp
