// replay.cpp -- replays processPileUp's call sequence (processPileUps/package.d:474-619) through the C++
// mirror of dazzler.d (include/dentist_b200.hpp) on a tiny synthetic pile and prints what came back.
// Usage: replay <seed>   (tests/test_cpp_mirror.py builds it, runs it and compares with the Python path)
#include "dentist_b200.hpp"

#include <cstdio>
#include <cstdlib>

using namespace dentist::dazzler;

static uint64_t lcg(uint64_t &s) { s = s * 6364136223846793005ull + 1442695040888963407ull; return s >> 33; }

int main(int argc, char **argv) {
    uint64_t s = argc > 1 ? strtoull(argv[1], nullptr, 10) : 1;
    try {
        const int L = 3000, N = 8;
        std::vector<uint8_t> truth(L);
        for (auto &b : truth) b = (uint8_t)(lcg(s) & 3);
        std::vector<uint8_t> bases; std::vector<int32_t> rlen; std::vector<int64_t> boff;
        for (int r = 0; r < N; r++) {
            boff.push_back((int64_t)bases.size());
            int n = 0;
            for (int i = 0; i < L; i++) {
                uint64_t u = lcg(s) % 100;
                if (u < 3) continue;                                          // deletion
                if (u < 6) { bases.push_back((uint8_t)(lcg(s) & 3)); n++; }   // insertion
                bases.push_back(u < 8 ? (uint8_t)((truth[i] + 1 + lcg(s) % 3) & 3) : truth[i]); n++;
            }
            rlen.push_back(n);
        }
        dn_block_desc d{}; d.nreads = N; d.format = DN_SEQ_BYTES; d.rlen = rlen.data(); d.boff = boff.data();
        d.data = bases.data(); d.data_bytes = (int64_t)bases.size();
        Block db(d);
        printf("dust %lld\n", (long long)db.maskDust());
        dn_align_params p; dn_align_params_default(&p);
        p.tspace = 126; p.minlen = 500; p.self_block = 1;                      // daligner -s126 -l500 -e0.7 X X
        Las las = align(db, db, p);
        printf("raw %lld\n", (long long)las.size());
        las.filterLocalAlignments(0.3);
        las.chainLocalAlignments();
        printf("chained %lld\n", (long long)las.size());
        auto qv = computeQVs(rlen, las, 4);
        las.filterPileUpAlignments(rlen, rlen, 126);
        printf("filtered %lld\n", (long long)las.size());
        long long qsum = 0; for (auto &q : qv) for (auto x : q) qsum += x;
        printf("qvsum %lld\n", qsum);
        auto cov = maskCoverage(las, rlen, rlen, 0, 5);
        long long covered = 0; for (auto &c : cov) for (auto &iv : c) covered += iv.second - iv.first;
        printf("overcovered %lld\n", covered);
        Mask in(N); in[0].emplace_back(500, 900);
        auto prop = propagateMask(las, in, rlen);
        long long pb = 0; for (auto &c : prop) for (auto &iv : c) pb += iv.second - iv.first;
        printf("propagated %lld\n", pb);
        auto cons = getConsensus(db, las, 1);
        unsigned long long h = 1469598103934665603ull; int same = 0;
        for (size_t i = 0; i < cons.size(); i++) { h = (h ^ cons[i]) * 1099511628211ull; }
        for (size_t i = 0; i < cons.size() && i < truth.size(); i++) same += cons[i] == truth[i];
        printf("consensus %zu %llu\n", cons.size(), h);
        return 0;
    } catch (const DazzlerCommandException &e) {
        printf("DazzlerCommandException: %s\n", e.what());
        return 3;
    }
}
