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
        // the same pile-up (and a copy with another reference-read restriction) through the ONE batch entry point: the first
        // and the last 1200 truth bases are its flanking contigs
        std::vector<uint8_t> fl(truth.begin(), truth.begin() + 1200), fr(truth.end() - 1200, truth.end());
        std::vector<uint8_t> fbases(fl); fbases.insert(fbases.end(), fr.begin(), fr.end());
        std::vector<int32_t> flen{1200, 1200}; std::vector<int64_t> foff{0, 1200};
        dn_block_desc fd{}; fd.nreads = 2; fd.format = DN_SEQ_BYTES; fd.rlen = flen.data(); fd.boff = foff.data(); fd.data = fbases.data(); fd.data_bytes = 2400;
        Block ref(fd);
        PileUp pu; pu.flankingContigs = {0, 1};
        for (int r = 0; r < N; r++) pu.croppedReads.emplace_back(bases.begin() + boff[r], bases.begin() + boff[r] + rlen[r]);
        PileUp pv = pu; pv.allowedReferenceRead.assign(N, 0); pv.allowedReferenceRead[3] = 1;
        auto res = processPileUps(ref, {pu, pv});
        for (auto &o : res) {
            unsigned long long hc = 1469598103934665603ull, hf = 1469598103934665603ull;
            for (auto b : o.consensus) hc = (hc ^ b) * 1099511628211ull;
            for (size_t x = 0; x < o.postConsensusAlignment.size(); x++) {
                const dn_las_record &q = o.postConsensusAlignment[x];
                const int32_t f[8] = {q.aread, q.abpos, q.aepos, q.bbpos, q.bepos, q.diffs, (int32_t)q.flags, q.tlen};
                for (int32_t v : f) hf = (hf ^ (unsigned long long)(uint32_t)v) * 1099511628211ull;
                for (auto t : o.traces[x]) hf = (hf ^ t) * 1099511628211ull;
            }
            printf("batch %d|%d|%zu|%llu|%zu|%llu\n", o.status, o.referenceReadIdx, o.consensus.size(), hc, o.postConsensusAlignment.size(), hf);
        }
        return 0;
    } catch (const DazzlerCommandException &e) {
        printf("DazzlerCommandException: %s\n", e.what());
        return 3;
    }
}
