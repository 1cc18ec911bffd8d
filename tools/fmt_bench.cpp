// Development: single-thread cost per output row of the host's strand combining + bedMethyl formatting
// (finish_interval_rows + format_bed_row of csrc/host/pileup_host.hpp) on rows shaped like a CpG `--preset traditional` run.
// build: g++ -O3 -std=c++17 -I include -I modkit_b200/csrc/host tools/fmt_bench.cpp -o tools/_build/fmt_bench -lz -pthread
#include <chrono>
#include <cstdio>
#include <random>

#include "pileup_host.hpp"

using namespace mkh;

int main() {
    std::mt19937 rng(7);
    const uint32_t L = 100000;
    const int n_iv = 200;
    std::vector<RefInterval> ivs(n_iv);
    std::vector<std::vector<mkp_row>> rows(n_iv);
    size_t n_rows = 0;
    for (int k = 0; k < n_iv; k++) {
        RefInterval& iv = ivs[k];
        iv.tid = 0; iv.start = k * L; iv.end = iv.start + L; iv.all_positions = false; iv.flat_valid = true;
        for (uint32_t p = iv.start + rng() % 60; p + 1 < iv.end; p += 20 + rng() % 80) { iv.flat.push_back({p, 1}); iv.flat.push_back({p + 1, 2}); }
        for (auto& s : iv.flat) {
            mkp_row r;
            memset(&r, 0, sizeof r);
            r.pos = s.first; r.strand = (s.second & 1) ? '+' : '-'; r.code = 'm'; r.primary_base = 'C';
            r.n_mod = rng() % 20; r.n_canon = 1 + rng() % 20; r.n_filtered = rng() % 3; r.n_nocall = rng() % 2; r.n_diff = rng() % 2;
            rows[k].push_back(r);
        }
        n_rows += rows[k].size();
    }
    std::vector<MotifSpec> motifs{parse_motif("CG", 0)};
    BedFormat fmt;
    fmt.motif_labels.push_back(motifs[0].label());
    const std::string chrom = "syn1";
    for (int rep = 0; rep < 3; rep++) {
        std::string text;
        std::vector<OutRow> local;
        size_t out_rows = 0;
        const auto t0 = std::chrono::steady_clock::now();
        double t_fin = 0;
        for (int k = 0; k < n_iv; k++) {
            local.clear();
            const auto a = std::chrono::steady_clock::now();
            finish_interval_rows(ivs[k], rows[k].data(), rows[k].size(), &motifs, true, &local);
            t_fin += std::chrono::duration<double>(std::chrono::steady_clock::now() - a).count();
            for (auto& o : local) format_bed_row(o, chrom, fmt, &text);
            out_rows += local.size();
        }
        const double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("in rows %zu out rows %zu: %.1f ns per out row (combine %.1f ns), %zu bytes\n", n_rows, out_rows, s * 1e9 / out_rows, t_fin * 1e9 / out_rows, text.size());
    }
    return 0;
}
