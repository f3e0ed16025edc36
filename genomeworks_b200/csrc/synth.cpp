// gw-b200: C-ABI wrappers around include/claraparabricks/genomeworks/utils/genomeutils.hpp so that Python callers
// (tests, bench.py) build exactly the synthetic workloads SURVEY.md 8d / BASELINE.md define.
#include "../../include/gwb200.h"
#include "../../include/claraparabricks/genomeworks/utils/genomeutils.hpp"

#include <cstring>

using namespace claraparabricks::genomeworks;

extern "C" {

int64_t gwb200_synth_poa_windows(int32_t n_windows, uint32_t seed0, int32_t backbone_len, int32_t n_reads, int32_t max_mut, int32_t max_ins,
                                 int32_t max_del, int32_t max_read_len, int32_t* seq_len, char* seq_data, int64_t capacity)
{
    int64_t off = 0;
    for (int32_t w = 0; w < n_windows; ++w)
    {
        std::minstd_rand rng(seed0 + static_cast<uint32_t>(w));
        std::string backbone        = genomeutils::generate_random_genome(backbone_len, rng);
        std::vector<std::string> rs = genomeutils::generate_random_sequences(backbone, n_reads, rng, max_mut, max_ins, max_del);
        for (int32_t r = 0; r < n_reads; ++r)
        {
            std::string& s = rs[r];
            if (max_read_len > 0 && static_cast<int32_t>(s.size()) > max_read_len)
                s.resize(max_read_len);
            if (off + static_cast<int64_t>(s.size()) > capacity)
                return -1;
            std::memcpy(seq_data + off, s.data(), s.size());
            seq_len[static_cast<int64_t>(w) * n_reads + r] = static_cast<int32_t>(s.size());
            off += static_cast<int64_t>(s.size());
        }
    }
    return off;
}

int64_t gwb200_synth_aligner_pairs(int32_t n_pairs, uint32_t seed, int32_t genome_size, int32_t* q_len, char* q_data, int64_t q_capacity,
                                   int32_t* t_len, char* t_data, int64_t t_capacity)
{
    // cudaaligner/benchmarks/main.cpp:116-129: one rng(seed) for the whole batch
    std::minstd_rand rng(seed);
    int64_t qo = 0, to = 0;
    for (int32_t i = 0; i < n_pairs; ++i)
    {
        std::string g1 = genomeutils::generate_random_genome(genome_size, rng);
        std::string g2 = genomeutils::generate_random_sequence(g1, rng, genome_size / 30, genome_size / 30, genome_size / 30);
        if (static_cast<int32_t>(g2.size()) > genome_size)
            g2.resize(genome_size);
        if (qo + static_cast<int64_t>(g1.size()) > q_capacity || to + static_cast<int64_t>(g2.size()) > t_capacity)
            return -1;
        std::memcpy(q_data + qo, g1.data(), g1.size());
        std::memcpy(t_data + to, g2.data(), g2.size());
        q_len[i] = static_cast<int32_t>(g1.size());
        t_len[i] = static_cast<int32_t>(g2.size());
        qo += static_cast<int64_t>(g1.size());
        to += static_cast<int64_t>(g2.size());
    }
    return qo + to;
}

} // extern "C"
