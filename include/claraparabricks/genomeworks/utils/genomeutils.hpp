// gw-b200: synthetic genome / read generators with the semantics (and, with the same libstdc++, the same random
// streams) of the reference's header-only helpers
// common/base/include/claraparabricks/genomeworks/utils/genomeutils.hpp:33-142 -- generate_random_genome,
// generate_random_sequence, generate_random_sequences, reverse_complement. Used by tests, samples and bench.py to
// build the BASELINE.json workloads.
#pragma once

#include <algorithm>
#include <cstdint>
#include <random>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace claraparabricks
{
namespace genomeworks
{
namespace genomeutils
{

/// Uniform random genome over {A,C,G,T}; one draw of uniform_int_distribution<int32_t>(0,3) per base.
inline std::string generate_random_genome(const int32_t length, std::minstd_rand& rng)
{
    static const char alphabet[4] = {'A', 'C', 'G', 'T'};
    std::uniform_int_distribution<int32_t> pick(0, 3);
    std::string genome;
    genome.reserve(length > 0 ? length : 0);
    for (int32_t i = 0; i < length; ++i)
        genome.push_back(alphabet[pick(rng)]);
    return genome;
}

/// Noisy copy of `backbone`: per range, up to max_deletions single-base deletions, then up to max_insertions
/// single-base insertions, then up to max_mutations substitutions, each applied with probability 1/2
/// (a draw of uniform_real_distribution<double>(0,1) > 0.5), in that order.
inline std::string generate_random_sequence(const std::string& backbone, std::minstd_rand& rng, int max_mutations, int max_insertions,
                                            int max_deletions, std::vector<std::pair<int, int>>* ranges = nullptr)
{
    if (max_mutations < 0)
        throw std::invalid_argument("max_mutations cannot be negative.");
    if (max_insertions < 0)
        throw std::invalid_argument("max_insertions cannot be negative.");
    if (max_deletions < 0)
        throw std::invalid_argument("max_deletions cannot be negative.");
    static const char alphabet[4] = {'A', 'C', 'G', 'T'};
    std::uniform_int_distribution<int> random_base(0, 3);
    std::string sequence = backbone;
    std::vector<std::pair<int, int>> whole(1, std::make_pair(0, static_cast<int>(backbone.size())));
    if (ranges == nullptr)
        ranges = &whole;
    for (const auto& range : *ranges)
    {
        const int start_index = range.first;
        const int end_index   = range.second;
        if (start_index < 0)
            throw std::invalid_argument("start_index of the range cannot be negative.");
        if (end_index - start_index < 0)
            throw std::invalid_argument("end_index of the range cannot be smaller than start_index.");
        if (static_cast<int>(backbone.size()) < end_index)
            throw std::invalid_argument("end_index should be smaller than backbone's length.");
        const int range_length = end_index - start_index;
        std::string piece      = backbone.substr(start_index, range_length);
        std::uniform_real_distribution<double> coin(0, 1);
        for (int j = 0; j < std::min(max_deletions, range_length); ++j)
        {
            if (coin(rng) > 0.5)
            {
                const int length = static_cast<int>(piece.length());
                std::uniform_int_distribution<int> where(0, length - 1);
                piece.erase(where(rng), 1);
            }
        }
        for (int j = 0; j < std::min(max_insertions, range_length); ++j)
        {
            if (coin(rng) > 0.5)
            {
                const int length = static_cast<int>(piece.length());
                std::uniform_int_distribution<int> where(0, length);
                const int pos  = where(rng);
                const int base = random_base(rng);
                piece.insert(pos, 1, alphabet[base]);
            }
        }
        const int length = static_cast<int>(piece.length());
        if (length > 0)
        {
            std::uniform_int_distribution<int> where(0, length - 1);
            for (int j = 0; j < std::min(max_mutations, range_length); ++j)
            {
                if (coin(rng) > 0.5)
                {
                    const int pos  = where(rng);
                    const int base = random_base(rng);
                    piece[pos]     = alphabet[base];
                }
            }
        }
        if (start_index < static_cast<int>(sequence.length()))
            sequence.replace(start_index, range_length, piece);
    }
    return sequence;
}

/// n sequences: the backbone itself followed by n-1 noisy copies.
inline std::vector<std::string> generate_random_sequences(std::string const& backbone, int n, std::minstd_rand& rng, int max_mutations = 1,
                                                          int max_insertion = 1, int max_deletions = 1)
{
    if (n < 0)
        throw std::invalid_argument("n cannot be negative!");
    std::vector<std::string> out;
    out.reserve(n);
    out.push_back(backbone);
    for (int i = 1; i < n; ++i)
        out.push_back(generate_random_sequence(backbone, rng, max_mutations, max_insertion, max_deletions));
    return out;
}

/// dest[pos] = complement(src[length-1-pos]); A<->T, C<->G (lookup on bits 1..2 of the ASCII code, like the reference).
inline void reverse_complement(const char* src, const int32_t length, char* dest)
{
    static const char lookup[4] = {'T', 'G', 'A', 'C'}; // index (c >> 1) & 3 : A=0, C=1, T=2, G=3
    for (int32_t pos = 0; pos < length; ++pos)
    {
        const unsigned char c = static_cast<unsigned char>(src[length - 1 - pos]);
        dest[pos]             = lookup[(c >> 1) & 0x3];
    }
}

} // namespace genomeutils
} // namespace genomeworks
} // namespace claraparabricks
