// ORACLE — TEST INFRASTRUCTURE ONLY. Never linked into or called from the product path.
//
// Indexed FASTA slice reader (`.fai`) and motif search, restating
//   src/find_motifs/motif_bed.rs:21-330  (iupac_to_regex, motif_rev_comp, find_motif_hits)
//   src/fasta.rs:92-226                  (interval motif lookup, combine-strands interval extension)
// The reference uses the `regex` crate for fixed-length IUPAC patterns; a fixed-length class
// matcher with overlapping starts is equivalent (OverlappingPatternIterator, motif_bed.rs:68-84).
#pragma once
#include <cstdint>
#include <cstdio>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace orc {

struct FaiEntry { uint64_t length, offset, linebases, linewidth; };

struct Fasta {
    std::string path;
    std::map<std::string, FaiEntry> idx;
    FILE* fh = nullptr;
    void open(const std::string& p) {
        path = p;
        std::ifstream f(p + ".fai");
        if (!f) throw std::runtime_error("missing FASTA index " + p + ".fai");
        std::string line;
        while (std::getline(f, line)) {
            std::istringstream ss(line);
            std::string name;
            FaiEntry e;
            if (ss >> name >> e.length >> e.offset >> e.linebases >> e.linewidth) idx[name] = e;
        }
        fh = fopen(p.c_str(), "rb");
        if (!fh) throw std::runtime_error("cannot open " + p);
    }
    ~Fasta() { if (fh) fclose(fh); }
    // bases [start, stop) of contig
    std::string fetch(const std::string& contig, uint64_t start, uint64_t stop) {
        auto it = idx.find(contig);
        if (it == idx.end()) throw std::runtime_error("contig " + contig + " not in FASTA index");
        const FaiEntry& e = it->second;
        if (stop > e.length) stop = e.length;
        std::string out;
        if (start >= stop) return out;
        out.reserve(stop - start);
        uint64_t line = start / e.linebases, col = start % e.linebases;
        uint64_t fo = e.offset + line * e.linewidth + col;
        uint64_t nbytes = (stop - start) + ((stop - start) / e.linebases + 2) * (e.linewidth - e.linebases);
        std::vector<char> buf(nbytes);
        fseek(fh, (long)fo, SEEK_SET);
        size_t got = fread(buf.data(), 1, nbytes, fh);
        for (size_t i = 0; i < got && out.size() < stop - start; i++) {
            char c = buf[i];
            if (c == '\n' || c == '\r') continue;
            out.push_back(c);
        }
        return out;
    }
};

struct Motif {
    std::string raw;
    int fwd_off = 0, rev_off = 0, length = 0;
    bool palindrome = false;
    std::string fwd_re, rev_re;                 // regex strings as the reference builds them
    std::vector<uint8_t> fwd_cls, rev_cls;      // per-position 4-bit class masks A=1 C=2 G=4 T=8

    static bool iupac(char c, std::string* re) {
        switch (c) {
            case 'A': *re = "A"; break; case 'C': *re = "C"; break; case 'G': *re = "G"; break; case 'T': *re = "T"; break;
            case 'U': *re = "U"; break;
            case 'M': *re = "[AC]"; break; case 'R': *re = "[AG]"; break; case 'W': *re = "[AT]"; break;
            case 'S': *re = "[CG]"; break; case 'Y': *re = "[CT]"; break; case 'K': *re = "[GT]"; break;
            case 'V': *re = "[ACG]"; break; case 'H': *re = "[ACT]"; break; case 'D': *re = "[AGT]"; break;
            case 'B': *re = "[CGT]"; break; case 'X': case 'N': *re = "[ACGT]"; break;
            default: return false;
        }
        return true;
    }
    static std::vector<uint8_t> classes(const std::string& re) {
        std::vector<uint8_t> out;
        for (size_t i = 0; i < re.size(); i++) {
            uint8_t m = 0;
            auto bit = [](char c) -> uint8_t { return c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 4 : c == 'T' ? 8 : 0; };
            if (re[i] == '[') { i++; while (i < re.size() && re[i] != ']') m |= bit(re[i++]); }
            else m = bit(re[i]);  // 'U' never matches an upper-cased DNA reference
            out.push_back(m);
        }
        return out;
    }
    static Motif parse(const std::string& raw, int offset) {
        Motif m;
        m.raw = raw;
        m.length = (int)raw.size();
        if (m.length == 1 && !(raw == "A" || raw == "C" || raw == "G" || raw == "T"))
            throw std::runtime_error("degenerate bases are not supported as single base motifs");
        for (char c : raw) { std::string r; if (!iupac(c, &r)) throw std::runtime_error(std::string("Invalid IUPAC code: ") + c); m.fwd_re += r; }
        // motif_rev_comp on the regex string
        std::string rc(m.fwd_re.rbegin(), m.fwd_re.rend());
        for (char& c : rc) {
            switch (c) { case 'A': c = 'T'; break; case 'C': c = 'G'; break; case 'G': c = 'C'; break; case 'T': c = 'A'; break;
                         case 'U': c = 'A'; break; case '[': c = ']'; break; case ']': c = '['; break; default: break; }
        }
        m.rev_re = rc;
        if (offset + 1 > m.length) throw std::runtime_error("motif not long enough for offset");
        m.fwd_off = offset;
        m.rev_off = m.length - (offset + 1);
        m.palindrome = (m.fwd_re == m.rev_re);
        m.fwd_cls = classes(m.fwd_re);
        m.rev_cls = classes(m.rev_re);
        return m;
    }
    std::string label() const { return raw + "," + std::to_string(fwd_off); }
};

inline uint8_t nt_bit(char c) { return c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 4 : c == 'T' ? 8 : 0; }

inline bool match_at(const std::string& seq, size_t i, const std::vector<uint8_t>& cls) {
    if (i + cls.size() > seq.size()) return false;
    for (size_t k = 0; k < cls.size(); k++) if (!(nt_bit(seq[i + k]) & cls[k])) return false;
    return true;
}

// find_motif_hits (motif_bed.rs:260-330): (position-in-seq, strand '+'/'-')
inline std::vector<std::pair<uint32_t, char>> find_motif_hits(const std::string& seq, const Motif& m) {
    std::vector<std::pair<uint32_t, char>> hits;
    if (m.palindrome) {
        for (size_t i = 0; i < seq.size(); i++) {
            if (!match_at(seq, i, m.fwd_cls)) continue;
            if (m.fwd_off <= m.rev_off) { hits.push_back({(uint32_t)(i + m.fwd_off), '+'}); hits.push_back({(uint32_t)(i + m.rev_off), '-'}); }
            else { hits.push_back({(uint32_t)(i + m.rev_off), '-'}); hits.push_back({(uint32_t)(i + m.fwd_off), '+'}); }
        }
    } else if (m.length == 1) {
        char fw = m.raw[0];
        char rv = fw == 'A' ? 'T' : fw == 'C' ? 'G' : fw == 'G' ? 'C' : 'A';
        for (size_t i = 0; i < seq.size(); i++) {
            if (seq[i] == fw) hits.push_back({(uint32_t)i, '+'});
            else if (seq[i] == rv) hits.push_back({(uint32_t)i, '-'});
        }
    } else {
        for (size_t i = 0; i < seq.size(); i++) if (match_at(seq, i, m.fwd_cls)) hits.push_back({(uint32_t)(i + m.fwd_off), '+'});
        for (size_t i = 0; i < seq.size(); i++) if (match_at(seq, i, m.rev_cls)) hits.push_back({(uint32_t)(i + m.rev_off), '-'});
    }
    return hits;
}

// StrandRule: 1 = Positive, 2 = Negative, 3 = Both
typedef std::map<uint32_t, uint8_t> MotifLocs;

inline std::vector<MotifLocs> motifs_on_seq(const std::string& seq, uint64_t start, const std::vector<Motif>& motifs) {
    std::vector<MotifLocs> out;
    for (const Motif& m : motifs) {
        MotifLocs locs;
        for (auto& h : find_motif_hits(seq, m)) {
            uint32_t p = (uint32_t)(h.first + start);
            uint8_t r = h.second == '+' ? 1 : 2;
            auto it = locs.find(p);
            if (it == locs.end()) locs[p] = r; else it->second = (it->second == r) ? r : 3;  // StrandRule::absorb
        }
        out.push_back(std::move(locs));
    }
    return out;
}

}  // namespace orc
