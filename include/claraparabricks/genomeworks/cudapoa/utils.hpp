// gw-b200: cudapoa host utilities with the reference's names and semantics
// (cudapoa/include/claraparabricks/genomeworks/cudapoa/utils.hpp:55-185, cudapoa/src/utils.cu:30-147):
// get_multi_batch_sizes, resize_windows, parse_cudapoa_file, parse_fasta_files, parse_golden_value_file.
// Header-only; the only engine call is gwb200_poa_estimate_max_poas (this engine's arena layout decides how many windows of
// a given size fit a batch, so bin boundaries can differ from the reference's while the binning rule is the same).
#pragma once

#include "batch.hpp"

#include <algorithm>
#include <cstdint>
#include <fstream>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace claraparabricks
{
namespace genomeworks
{
namespace cudapoa
{

/// Groups POA groups of similar size into batches (utils.hpp:55-68). A group's size class is the number of windows of its own
/// size that one batch could hold; classes are the bins `bins_capacity` (default 1, 2, 4, ... 2^19). Every non-empty bin becomes
/// one BatchConfig sized for the longest read / largest group in it, and it absorbs the following non-empty bins as long as their
/// group count does not exceed its capacity.
inline void get_multi_batch_sizes(std::vector<BatchConfig>& list_of_batch_sizes, std::vector<std::vector<int32_t>>& list_of_groups_per_batch,
                                  const std::vector<Group>& poa_groups, bool msa_flag = false, int32_t band_width = 256,
                                  BandMode band_mode = BandMode::adaptive_band, float adaptive_storage_factor = 2.0f,
                                  float graph_length_factor = 3.0f, int32_t max_pred_distance = 0, std::vector<int32_t>* bins_capacity = nullptr,
                                  float gpu_memory_usage_quota = 0.9, int32_t mismatch_score = -6, int32_t gap_score = -8,
                                  int32_t match_score = 8)
{
    const int32_t num_groups = static_cast<int32_t>(poa_groups.size());
    std::vector<int64_t> capacity_of(num_groups);
    std::vector<int32_t> longest_read(num_groups);
    for (int32_t i = 0; i < num_groups; i++)
    {
        int32_t longest = 0;
        for (const Entry& e : poa_groups[i])
            longest = std::max(longest, e.length);
        const BatchConfig cfg(longest, static_cast<int32_t>(poa_groups[i].size()), band_width, band_mode, adaptive_storage_factor,
                              graph_length_factor, max_pred_distance);
        const gwb200_poa_config c = cfg.to_c();
        const int64_t n           = gwb200_poa_estimate_max_poas(&c, msa_flag ? 1 : 0, gpu_memory_usage_quota, static_cast<int16_t>(mismatch_score),
                                                                 static_cast<int16_t>(gap_score), static_cast<int16_t>(match_score));
        if (n < 0)
            detail::check(static_cast<int>(n));
        capacity_of[i]  = n;
        longest_read[i] = longest;
    }

    std::vector<int32_t> powers_of_two;
    if (bins_capacity == nullptr)
    {
        powers_of_two.resize(20);
        for (int32_t j = 0; j < 20; j++)
            powers_of_two[j] = 1 << j;
        bins_capacity = &powers_of_two;
    }
    const int32_t num_bins = static_cast<int32_t>(bins_capacity->size());

    struct Bin
    {
        std::vector<int32_t> groups;
        int32_t longest = 0, most_reads = 0;
    };
    std::vector<Bin> bins(num_bins);
    for (int32_t i = 0; i < num_groups; i++)
    {
        int32_t j = 0;
        while (j < num_bins - 1 && capacity_of[i] > bins_capacity->at(j))
            j++;
        bins[j].groups.push_back(i);
        bins[j].longest    = std::max(bins[j].longest, longest_read[i]);
        bins[j].most_reads = std::max(bins[j].most_reads, static_cast<int32_t>(poa_groups[i].size()));
    }
    for (int32_t j = 0; j < num_bins; j++)
    {
        if (bins[j].groups.empty())
            continue;
        list_of_batch_sizes.emplace_back(bins[j].longest, bins[j].most_reads, band_width, band_mode, adaptive_storage_factor, graph_length_factor,
                                         max_pred_distance);
        list_of_groups_per_batch.push_back(bins[j].groups);
        // smaller groups (bins of higher capacity) ride along while they fit this bin's capacity
        for (int32_t k = j + 1; k < num_bins; k++)
        {
            if (bins[k].groups.empty())
                continue;
            if (bins_capacity->at(j) < static_cast<int32_t>(bins[k].groups.size()))
                break;
            std::vector<int32_t>& current = list_of_groups_per_batch.back();
            current.insert(current.end(), bins[k].groups.begin(), bins[k].groups.end());
            bins[k].groups.clear();
        }
    }
}

/// Truncates the window list to `total_windows`, or repeats the windows read so far (cyclically, in order) until there are
/// `total_windows` of them; a negative `total_windows` leaves the list alone (utils.hpp:77-95).
inline void resize_windows(std::vector<std::vector<std::string>>& windows, const int32_t total_windows)
{
    if (total_windows < 0)
        return;
    const size_t want = static_cast<size_t>(total_windows);
    if (windows.size() > want)
    {
        windows.resize(want);
        return;
    }
    const size_t have = windows.size();
    if (have == 0)
        return;
    for (size_t k = 0; windows.size() < want; k++)
        windows.push_back(windows[k % have]);
}

/// Reads a cudapoa window file: a line with the number of sequences of a window, then that many sequence lines, repeated
/// (utils.hpp:97-142).
inline void parse_cudapoa_file(std::vector<std::vector<std::string>>& windows, const std::string& filename, int32_t total_windows)
{
    std::ifstream in(filename);
    if (!in.good())
        throw std::runtime_error("Cannot read file " + filename);
    std::string line;
    int32_t remaining = 0;
    while (std::getline(in, line))
    {
        if (remaining == 0)
        {
            std::istringstream header(line);
            header >> remaining;
            windows.emplace_back();
        }
        else
        {
            windows.back().push_back(line);
            remaining--;
        }
    }
    resize_windows(windows, total_windows);
}

/// One window per FASTA file, one sequence per record (utils.hpp:144-166). Plain-text FASTA (multi-line records allowed) and
/// FASTQ; compressed input is not handled by this reader.
inline void parse_fasta_files(std::vector<std::vector<std::string>>& windows, const std::vector<std::string>& input_paths,
                              const int32_t total_windows)
{
    windows.resize(input_paths.size());
    for (size_t f = 0; f < input_paths.size(); f++)
    {
        std::ifstream in(input_paths[f]);
        if (!in.good())
            throw std::runtime_error("Cannot read file " + input_paths[f]);
        std::string line;
        bool fastq = false, in_quality = false, have_record = false;
        std::string seq;
        auto flush = [&]() {
            if (have_record)
                windows[f].push_back(seq);
            seq.clear();
            have_record = false;
        };
        while (std::getline(in, line))
        {
            if (!line.empty() && line.back() == '\r')
                line.pop_back();
            if (line.empty())
                continue;
            if (in_quality)
            {
                in_quality = false; // single-line quality strings
                continue;
            }
            if (line[0] == '>' || (line[0] == '@' && (!have_record || fastq)))
            {
                flush();
                fastq       = line[0] == '@';
                have_record = true;
            }
            else if (line[0] == '+' && fastq)
            {
                in_quality = true;
            }
            else if (have_record)
            {
                seq += line;
            }
        }
        flush();
    }
    resize_windows(windows, total_windows);
}

/// First line of a golden-value file (utils.hpp:168-185).
inline std::string parse_golden_value_file(const std::string& filename)
{
    std::ifstream in(filename);
    if (!in.good())
        throw std::runtime_error("Cannot read file " + filename);
    std::string line;
    std::getline(in, line);
    return line;
}

} // namespace cudapoa
} // namespace genomeworks
} // namespace claraparabricks
