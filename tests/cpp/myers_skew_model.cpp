// CPU model of the skewed Myers score pass (genomeworks_b200/csrc/myers_skew.cuh): runs the kernel's own lane-level functions
// (same header, compiled by g++) lane by lane in the kernel's step order, writes the records in the kernel's layout and
// compares score_at() with the oracle's get_myers_score() for EVERY cell of the band. The oracle follows the reference
// formulation word by word (oracle/myers_oracle.cpp), so equality here means the new formulation computes the reference's band.
// Build + run: tests/test_myers_skew_model.py
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <string>
#include <vector>

#include "../../genomeworks_b200/csrc/myers_skew.cuh"

extern "C" int32_t oracle_myers_band_scores(const char* query, int32_t query_length, const char* target, int32_t target_length, int32_t band_width,
                                            int32_t p, int32_t* out, int32_t* diagonal_begin, int32_t* diagonal_end);

using namespace gwb200::myers::skew;

namespace
{

struct Records
{
    const Geom* g;
    std::vector<uint32_t> w;
    void pvmv(int32_t B, int32_t j, uint64_t& pv, uint64_t& mv) const
    {
        const int64_t c = chunk_index(*g, B, j, j % kK) * 4;
        pv              = static_cast<uint64_t>(w[c]) | (static_cast<uint64_t>(w[c + 1]) << 32);
        mv              = static_cast<uint64_t>(w[c + 2]) | (static_cast<uint64_t>(w[c + 3]) << 32);
    }
    int32_t S(int32_t B, int32_t j) const
    {
        const int64_t c = chunk_index(*g, B, j, kK + (j % kK) / 4) * 4;
        return static_cast<int32_t>(w[c + (j % kK) % 4]);
    }
};

uint64_t peq64(const std::string& q, int32_t B, char x)
{
    static const char chars[4] = {'A', 'C', 'T', 'G'};
    const char c               = chars[(x >> 1) & 3]; // get_query_pattern's character index (myers_gpu.cu:215)
    uint64_t r                 = 0;
    for (int32_t b = 0; b < 64; b++)
    {
        const int64_t i = 64ll * B + b;
        if (i < static_cast<int64_t>(q.size()) && q[i] == c)
            r |= 1ull << b;
    }
    return r;
}

// the warp of the kernel, one lane after the other; the links are exchanged at the start of a step (shuffle of last step's value)
void run_model(const Geom& g, const std::string& q, const std::string& t, Records& R)
{
    R.g = &g;
    R.w.assign(words_needed(g), 0xdeadbeefu);
    std::vector<LaneState> L(g.nbl);
    std::vector<Link> link(g.nbl), next(g.nbl);
    for (int32_t l = 0; l < g.nbl; l++)
    {
        lane_init(L[l], l);
        link[l].hbits = 0;
        link[l].S0    = 0;
    }
    for (int32_t s = 0; s < g.n_steps; s++)
    {
        for (int32_t l = 0; l < g.nbl; l++)
        {
            const Link in = link[(l + g.nbl - 1) % g.nbl];
            next[l]       = link[l];
            int32_t cb    = s - L[l].B;
            if (L[l].B <= g.last_block && cb >= 0 && cb < g.n_batches && block_retired(g, L[l].B, cb))
            {
                lane_init(L[l], L[l].B + g.nbl);
                cb = s - L[l].B;
            }
            if (L[l].B > g.last_block || cb < 0 || cb >= g.n_batches)
                continue;
            uint64_t eqs[kK];
            for (int32_t k = 0; k < kK; k++)
            {
                const int32_t col = kK * cb + k;
                eqs[k]            = (col >= 1 && col <= g.tsize) ? peq64(q, L[l].B, t[col - 1]) : 0ull;
            }
            uint64_t rec[kK][2];
            int32_t recS[kK];
            next[l] = lane_step(g, L[l], cb, eqs, in, rec, recS);
            if (!block_in_band(g, L[l].B, cb))
                continue; // the kernel does not store these records (nobody may read them: the slots keep 0xdeadbeef)
            for (int32_t k = 0; k < kK; k++)
            {
                const int64_t c = chunk_index(g, L[l].B, kK * cb + k, k) * 4;
                R.w[c]          = static_cast<uint32_t>(rec[k][0]);
                R.w[c + 1]      = static_cast<uint32_t>(rec[k][0] >> 32);
                R.w[c + 2]      = static_cast<uint32_t>(rec[k][1]);
                R.w[c + 3]      = static_cast<uint32_t>(rec[k][1] >> 32);
                const int64_t cs = chunk_index(g, L[l].B, kK * cb + k, kK + k / 4) * 4;
                R.w[cs + k % 4]  = static_cast<uint32_t>(recS[k]);
            }
        }
        link = next;
    }
}

std::string random_seq(std::mt19937& rng, int32_t n)
{
    static const char a[4] = {'A', 'C', 'G', 'T'};
    std::string s(n, 'A');
    for (auto& c : s)
        c = a[rng() & 3];
    return s;
}

std::string mutate(std::mt19937& rng, const std::string& s, int32_t subs, int32_t ins, int32_t dels, int32_t out_len)
{
    static const char a[4] = {'A', 'C', 'G', 'T'};
    std::string r = s;
    for (int32_t k = 0; k < subs && !r.empty(); k++)
        r[rng() % r.size()] = a[rng() & 3];
    for (int32_t k = 0; k < ins; k++)
        r.insert(r.begin() + rng() % (r.size() + 1), a[rng() & 3]);
    for (int32_t k = 0; k < dels && r.size() > 1; k++)
        r.erase(r.begin() + rng() % r.size());
    if (out_len > 0)
    {
        while (static_cast<int32_t>(r.size()) < out_len)
            r.push_back(a[rng() & 3]);
        r.resize(out_len);
    }
    return r;
}

int64_t n_checked = 0, n_passes = 0;

// all passes of the Ukkonen loop (myers_gpu.cu:955-1002) for this pair, each checked cell by cell
bool check_pair(const std::string& q, const std::string& t, int32_t max_bandwidth, bool verbose)
{
    const int32_t qs = static_cast<int32_t>(q.size()), ts = static_cast<int32_t>(t.size());
    if (max_bandwidth > qs)
        max_bandwidth = (qs % 32 == 1 ? qs + 1 : qs);
    if (max_bandwidth - 1 < std::abs(ts - qs))
        return true;
    int32_t estimate = std::max(1, std::abs(ts - qs) + std::min(ts, qs) / 20);
    for (int32_t pass = 0; pass < 12; pass++, estimate *= 2)
    {
        int32_t p  = std::min(std::min(ts, qs), (estimate - std::abs(ts - qs)) / 2);
        int32_t bw = std::min(1 + 2 * p + std::abs(ts - qs), qs);
        if (bw % 32 == 1 && bw != qs)
        {
            p += 1;
            bw = std::min(1 + 2 * p + std::abs(ts - qs), qs);
        }
        if (bw > max_bandwidth)
        {
            bw = max_bandwidth;
            p  = (bw - 1 - std::abs(ts - qs)) / 2;
        }
        std::vector<int32_t> ref(static_cast<size_t>(bw) * (ts + 1));
        int32_t db = 0, de = 0;
        oracle_myers_band_scores(q.data(), qs, t.data(), ts, bw, p, ref.data(), &db, &de);
        const Geom g = make_geom(bw, qs, ts, db);
        if (usable(g))
        {
            Records R;
            run_model(g, q, t, R);
            n_passes++;
            for (int32_t j = 0; j <= ts; j++)
                for (int32_t i = 1; i <= bw; i++)
                {
                    const int32_t ours = score_at(g, i, j, R);
                    const int32_t want = ref[(i - 1) + static_cast<size_t>(bw) * j];
                    n_checked++;
                    if (ours != want)
                    {
                        std::printf("MISMATCH q=%d t=%d max_bw=%d pass=%d bw=%d p=%d db=%d de=%d nbl=%d: D(i=%d, j=%d) ours %d reference %d (top %d)\n", qs,
                                    ts, max_bandwidth, pass, bw, p, db, de, g.nbl, i, j, ours, want, g.top(j));
                        return false;
                    }
                }
            if (verbose)
                std::printf("ok q=%d t=%d max_bw=%d pass=%d bw=%d p=%d db=%d de=%d lanes=%d steps=%d words=%lld (reference layout %lld)\n", qs, ts,
                            max_bandwidth, pass, bw, p, db, de, g.nbl, g.n_steps, static_cast<long long>(words_needed(g)),
                            static_cast<long long>(3ll * ((bw + 31) / 32) * (ts + 1)));
        }
        const int32_t dist = ref[(bw - 1) + static_cast<size_t>(bw) * ts];
        if (dist <= estimate || bw == qs || bw == max_bandwidth)
            break;
    }
    return true;
}

// one band given directly (the unbanded case band_width == query_size never comes first in the Ukkonen loop of a similar pair)
bool check_band(const std::string& q, const std::string& t, int32_t bw, int32_t p)
{
    const int32_t qs = static_cast<int32_t>(q.size()), ts = static_cast<int32_t>(t.size());
    std::vector<int32_t> ref(static_cast<size_t>(bw) * (ts + 1));
    int32_t db = 0, de = 0;
    oracle_myers_band_scores(q.data(), qs, t.data(), ts, bw, p, ref.data(), &db, &de);
    const Geom g = make_geom(bw, qs, ts, db);
    if (!usable(g))
    {
        std::printf("not usable: q=%d t=%d bw=%d\n", qs, ts, bw);
        return false;
    }
    Records R;
    run_model(g, q, t, R);
    n_passes++;
    for (int32_t j = 0; j <= ts; j++)
        for (int32_t i = 1; i <= bw; i++)
        {
            n_checked++;
            if (score_at(g, i, j, R) != ref[(i - 1) + static_cast<size_t>(bw) * j])
            {
                std::printf("MISMATCH (direct) q=%d t=%d bw=%d p=%d db=%d: D(i=%d, j=%d) ours %d reference %d\n", qs, ts, bw, p, db, i, j,
                            score_at(g, i, j, R), ref[(i - 1) + static_cast<size_t>(bw) * j]);
                return false;
            }
        }
    std::printf("ok (direct) q=%d t=%d bw=%d p=%d db=%d de=%d lanes=%d\n", qs, ts, bw, p, db, de, g.nbl);
    return true;
}

} // namespace

int main(int argc, char** argv)
{
    const bool big = argc > 1 && std::string(argv[1]) == "--big";
    std::mt19937 rng(20260924);
    bool ok = true;
    // shapes: (query, target, substitutions, insertions, deletions, max_bandwidth)
    const int32_t shapes[][6] = {
        {1000, 1000, 33, 33, 33, 1024}, {1000, 1000, 100, 100, 100, 512},  {1000, 1000, 60, 60, 60, 256},   {1500, 1400, 40, 10, 110, 640},
        {1400, 1500, 40, 110, 10, 640}, {2000, 2000, 200, 200, 200, 1024}, {900, 1300, 20, 400, 0, 1024},   {1300, 900, 20, 0, 400, 1024},
        {640, 640, 5, 5, 5, 640},       {641, 700, 30, 80, 21, 1024},      {3000, 3000, 300, 300, 300, 800}, {2047, 2049, 100, 100, 98, 2000},
        {513, 513, 150, 150, 150, 1024}, {700, 700, 0, 0, 0, 1024},        {1984, 1984, 250, 250, 250, 1984}, {300, 5000, 10, 4700, 0, 8192},
    };
    for (const auto& sh : shapes)
        for (int32_t rep = 0; rep < 3 && ok; rep++)
        {
            const std::string q = random_seq(rng, sh[0]);
            const std::string t = mutate(rng, q, sh[2], sh[3], sh[4], sh[1]);
            ok                  = check_pair(q, t, sh[5], rep == 0);
        }
    // unrelated sequences (distance far beyond every estimate: all passes up to the largest band)
    for (int32_t rep = 0; rep < 2 && ok; rep++)
        ok = check_pair(random_seq(rng, 1200), random_seq(rng, 1150), 1024, rep == 0);
    // unbanded passes (band_width == query_size: no diagonal phase) and bands at the smallest supported width
    const int32_t direct[][4] = {{700, 720, 700, 40}, {1984, 1500, 1984, 300}, {130, 400, 130, 10}, {128, 100, 128, 20}, {1000, 1000, 128, 63},
                                 {1000, 1040, 170, 64}, {2048, 2048, 1024, 511}, {2048, 2048, 1022, 510}, {4000, 3900, 1024, 461}};
    for (const auto& d : direct)
    {
        if (!ok)
            break;
        const std::string q = random_seq(rng, d[0]);
        const std::string t = mutate(rng, q, d[0] / 30, d[0] / 30, d[0] / 30, d[1]);
        ok                  = check_band(q, t, d[2], d[3]);
    }
    if (big && ok)
    {
        // the C4 shape: 10 000 x 10 000, 333 edits of each kind, bands 501 and 1001
        const std::string q = random_seq(rng, 10000);
        const std::string t = mutate(rng, q, 333, 333, 333, 10000);
        ok                  = check_pair(q, t, 1024, true);
    }
    std::printf("%s: %lld passes, %lld cells compared\n", ok ? "PASS" : "FAIL", static_cast<long long>(n_passes), static_cast<long long>(n_checked));
    return ok ? 0 : 1;
}
