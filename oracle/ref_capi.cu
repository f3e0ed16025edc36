// TEST INFRASTRUCTURE ONLY -- not part of the product path.
//
// Thin extern "C" driver around the UNMODIFIED reference libraries (cudapoa + cudaaligner), which
// oracle/Makefile compiles from the sources where they lie under /root/reference into
// oracle/_ref/libgwref.so. Only tests/, __graft_entry__.smoke() and bench.py's baseline legs may
// load it. It calls nothing but the reference's own public API:
//   cudapoa:     create_batch / add_poa_group / generate_poa / get_consensus / get_msa
//                (cudapoa/include/claraparabricks/genomeworks/cudapoa/batch.hpp:90-204)
//   cudaaligner: create_aligner / add_alignment / align_all / sync_alignments
//                (cudaaligner/include/claraparabricks/genomeworks/cudaaligner/aligner.hpp:76-219)
// so its outputs are the reference's outputs on the GPU it runs on (the primary parity oracle).

#include <claraparabricks/genomeworks/cudapoa/cudapoa.hpp>
#include <claraparabricks/genomeworks/cudapoa/batch.hpp>
#include <claraparabricks/genomeworks/cudaaligner/cudaaligner.hpp>
#include <claraparabricks/genomeworks/cudaaligner/aligner.hpp>
#include <claraparabricks/genomeworks/cudaaligner/alignment.hpp>
// the in-library aligners that no factory exposes (AlignerGlobalMyers, AlignerGlobalUkkonen): the reference's own class headers
#include <aligner_global_myers.hpp>
#include <aligner_global_ukkonen.hpp>
#include <aligner_global_hirschberg_myers.hpp>

#include <cuda_runtime_api.h>

#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

using namespace claraparabricks::genomeworks;

namespace
{
double now_ms()
{
    using clk = std::chrono::steady_clock;
    return std::chrono::duration<double, std::milli>(clk::now().time_since_epoch()).count();
}
} // namespace

extern "C" {

// Runs the reference cudapoa over a flat list of windows.
//   win_nseq[n_windows], seq_len[sum(win_nseq)], seq_data = all reads concatenated (no padding).
//   weights may be NULL (=> reference fills weight 1).
// Outputs (caller allocated):
//   consensus  [n_windows * max_consensus]  forward orientation, NUL terminated ("" on error)
//   coverage   [n_windows * max_consensus]
//   status     [n_windows]                  cudapoa::StatusType per window
//   msa        [n_windows * max_seq_per_poa * max_consensus]  (only when output_mask & msa)
//   timings_ms [3]: total host wall time, sum of (generate_poa .. results on host), number of batches
// Returns 0 on success, <0 on failure (exception text in errbuf).
int ref_poa_run(int32_t n_windows, const int32_t* win_nseq, const int32_t* seq_len, const char* seq_data, const int8_t* weights,
                int32_t max_seq_size, int32_t max_seq_per_poa, int32_t band_width, int32_t band_mode,
                float adaptive_storage_factor, float graph_length_factor, int32_t max_pred_dist,
                int32_t output_mask, int32_t gap, int32_t mismatch, int32_t match,
                double mem_fraction, int32_t max_windows_per_batch,
                char* consensus, uint16_t* coverage, int32_t* status, char* msa,
                int32_t* max_consensus_out, double* timings_ms, char* errbuf, int32_t errbuf_len)
{
    try
    {
        cudapoa::Init();
        cudapoa::BatchConfig cfg(max_seq_size, max_seq_per_poa, band_width, static_cast<cudapoa::BandMode>(band_mode),
                                 adaptive_storage_factor, graph_length_factor, max_pred_dist);
        if (max_consensus_out)
            *max_consensus_out = cfg.max_consensus_size;
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        int64_t mem = static_cast<int64_t>(mem_fraction * static_cast<double>(free_b));
        cudaStream_t stream;
        cudaStreamCreate(&stream);
        const double t_all0 = now_ms();
        double t_proc       = 0.;
        int n_batches       = 0;
        {
            std::unique_ptr<cudapoa::Batch> batch = cudapoa::create_batch(0, stream, mem, static_cast<int8_t>(output_mask), cfg,
                                                                          static_cast<int16_t>(gap), static_cast<int16_t>(mismatch), static_cast<int16_t>(match));
            // prefix offsets
            std::vector<int64_t> seq_off;
            std::vector<int32_t> win_first;
            {
                int64_t off = 0;
                int32_t si  = 0;
                for (int32_t w = 0; w < n_windows; ++w)
                {
                    win_first.push_back(si);
                    for (int32_t s = 0; s < win_nseq[w]; ++s)
                    {
                        seq_off.push_back(off);
                        off += seq_len[si++];
                    }
                }
            }
            const int32_t mc = cfg.max_consensus_size;
            int32_t w        = 0;
            while (w < n_windows)
            {
                const int32_t batch_begin = w;
                batch->reset();
                while (w < n_windows)
                {
                    if (max_windows_per_batch > 0 && (w - batch_begin) >= max_windows_per_batch)
                        break;
                    cudapoa::Group g;
                    for (int32_t s = 0; s < win_nseq[w]; ++s)
                    {
                        const int32_t si = win_first[w] + s;
                        cudapoa::Entry e{};
                        e.seq     = seq_data + seq_off[si];
                        e.weights = weights ? weights + seq_off[si] : nullptr;
                        e.length  = seq_len[si];
                        g.push_back(e);
                    }
                    std::vector<cudapoa::StatusType> per_seq;
                    cudapoa::StatusType st = batch->add_poa_group(per_seq, g);
                    if (st == cudapoa::StatusType::exceeded_maximum_poas)
                    {
                        if (w == batch_begin)
                            throw std::runtime_error("window does not fit an empty reference batch");
                        break;
                    }
                    if (st != cudapoa::StatusType::success)
                        throw std::runtime_error("reference add_poa_group returned status " + std::to_string(static_cast<int>(st)));
                    ++w;
                }
                const double t0 = now_ms();
                batch->generate_poa();
                if (output_mask & cudapoa::OutputType::msa)
                {
                    std::vector<std::vector<std::string>> m;
                    std::vector<cudapoa::StatusType> st;
                    batch->get_msa(m, st);
                    t_proc += now_ms() - t0;
                    for (int32_t i = 0; i < static_cast<int32_t>(m.size()); ++i)
                    {
                        const int32_t ww = batch_begin + i;
                        status[ww]       = static_cast<int32_t>(st[i]);
                        if (consensus)
                            consensus[static_cast<int64_t>(ww) * mc] = 0;
                        for (int32_t r = 0; r < static_cast<int32_t>(m[i].size()); ++r)
                        {
                            char* dst = msa + (static_cast<int64_t>(ww) * max_seq_per_poa + r) * mc;
                            std::strncpy(dst, m[i][r].c_str(), mc - 1);
                            dst[mc - 1] = 0;
                        }
                    }
                }
                else
                {
                    std::vector<std::string> c;
                    std::vector<std::vector<uint16_t>> cov;
                    std::vector<cudapoa::StatusType> st;
                    batch->get_consensus(c, cov, st);
                    t_proc += now_ms() - t0;
                    for (int32_t i = 0; i < static_cast<int32_t>(c.size()); ++i)
                    {
                        const int32_t ww = batch_begin + i;
                        status[ww]       = static_cast<int32_t>(st[i]);
                        char* dst        = consensus + static_cast<int64_t>(ww) * mc;
                        std::strncpy(dst, c[i].c_str(), mc - 1);
                        dst[mc - 1] = 0;
                        if (coverage)
                        {
                            uint16_t* cd = coverage + static_cast<int64_t>(ww) * mc;
                            for (size_t k = 0; k < cov[i].size() && k < static_cast<size_t>(mc); ++k)
                                cd[k] = cov[i][k];
                        }
                    }
                }
                ++n_batches;
            }
        }
        if (timings_ms)
        {
            timings_ms[0] = now_ms() - t_all0;
            timings_ms[1] = t_proc;
            timings_ms[2] = n_batches;
        }
        cudaStreamDestroy(stream);
        return 0;
    }
    catch (const std::exception& e)
    {
        if (errbuf && errbuf_len > 0)
        {
            std::strncpy(errbuf, e.what(), errbuf_len - 1);
            errbuf[errbuf_len - 1] = 0;
        }
        return -1;
    }
}

// Runs the reference AlignerGlobalMyersBanded (create_aligner(global_alignment, max_bandwidth, ...),
// cudaaligner/src/aligner.cpp:76-124) over n_pairs pairs.
//   q_len/t_len[n_pairs]; q_data/t_data concatenated.
// Outputs per pair: status (cudaaligner::StatusType of the Alignment), is_optimal, edit_distance,
//   cigar_basic / cigar_ext: NUL-terminated strings in slots of cigar_stride bytes.
// timings_ms[2]: total wall, align_all..sync_alignments wall.
int ref_aligner_run(int32_t n_pairs, const int32_t* q_len, const char* q_data, const int32_t* t_len, const char* t_data,
                    int32_t max_bandwidth, int64_t max_device_memory,
                    int32_t* status, int32_t* is_optimal, int32_t* edit_distance,
                    char* cigar_basic, char* cigar_ext, int32_t cigar_stride,
                    double* timings_ms, char* errbuf, int32_t errbuf_len)
{
    try
    {
        cudaaligner::Init();
        cudaStream_t stream;
        cudaStreamCreate(&stream);
        const double t_all0 = now_ms();
        double t_proc       = 0.;
        {
            std::unique_ptr<cudaaligner::FixedBandAligner> aligner =
                cudaaligner::create_aligner(cudaaligner::AlignmentType::global_alignment, max_bandwidth, stream, 0, max_device_memory);
            int64_t qo = 0, to = 0;
            int32_t done = 0;
            std::vector<int64_t> qoff(n_pairs), toff(n_pairs);
            for (int32_t i = 0; i < n_pairs; ++i)
            {
                qoff[i] = qo;
                toff[i] = to;
                qo += q_len[i];
                to += t_len[i];
            }
            int32_t i = 0;
            while (i < n_pairs)
            {
                const int32_t begin = i;
                while (i < n_pairs)
                {
                    cudaaligner::StatusType st = aligner->add_alignment(q_data + qoff[i], q_len[i], t_data + toff[i], t_len[i]);
                    if (st == cudaaligner::StatusType::exceeded_max_alignments)
                    {
                        if (i == begin)
                            throw std::runtime_error("pair does not fit an empty reference aligner");
                        break;
                    }
                    if (st != cudaaligner::StatusType::success)
                        throw std::runtime_error("reference add_alignment returned status " + std::to_string(static_cast<int>(st)));
                    ++i;
                }
                const double t0 = now_ms();
                aligner->align_all();
                aligner->sync_alignments();
                t_proc += now_ms() - t0;
                const auto& res = aligner->get_alignments();
                for (int32_t k = 0; k < static_cast<int32_t>(res.size()); ++k)
                {
                    const int32_t p  = begin + k;
                    status[p]        = static_cast<int32_t>(res[k]->get_status());
                    is_optimal[p]    = res[k]->is_optimal() ? 1 : 0;
                    edit_distance[p] = res[k]->get_edit_distance();
                    std::string cb   = res[k]->convert_to_cigar(cudaaligner::CigarFormat::basic);
                    std::string ce   = res[k]->convert_to_cigar(cudaaligner::CigarFormat::extended);
                    if (static_cast<int32_t>(cb.size()) >= cigar_stride || static_cast<int32_t>(ce.size()) >= cigar_stride)
                        throw std::runtime_error("cigar_stride too small");
                    std::strcpy(cigar_basic + static_cast<int64_t>(p) * cigar_stride, cb.c_str());
                    std::strcpy(cigar_ext + static_cast<int64_t>(p) * cigar_stride, ce.c_str());
                }
                done += static_cast<int32_t>(res.size());
                aligner->reset();
            }
            (void)done;
        }
        if (timings_ms)
        {
            timings_ms[0] = now_ms() - t_all0;
            timings_ms[1] = t_proc;
        }
        cudaStreamDestroy(stream);
        return 0;
    }
    catch (const std::exception& e)
    {
        if (errbuf && errbuf_len > 0)
        {
            std::strncpy(errbuf, e.what(), errbuf_len - 1);
            errbuf[errbuf_len - 1] = 0;
        }
        return -1;
    }
}

// Runs one of the reference's fixed-size global aligners over n_pairs pairs in one batch:
//   algorithm 0: create_aligner(max_query, max_target, n_pairs, global_alignment, stream, 0) -- the deprecated factory
//                (cudaaligner/src/aligner.cpp:31-74: AlignerGlobalHirschbergMyers)
//   algorithm 1: AlignerGlobalMyers, 2: AlignerGlobalUkkonen (cudaaligner/src/aligner_global_{myers,ukkonen}.hpp)
// Outputs as ref_aligner_run.
int ref_global_aligner_run(int32_t n_pairs, const int32_t* q_len, const char* q_data, const int32_t* t_len, const char* t_data,
                           int32_t algorithm, int32_t max_query_length, int32_t max_target_length, int32_t* status, int32_t* is_optimal,
                           int32_t* edit_distance, char* cigar_basic, char* cigar_ext, int32_t cigar_stride, double* timings_ms, char* errbuf,
                           int32_t errbuf_len)
{
    try
    {
        cudaaligner::Init();
        cudaStream_t stream;
        cudaStreamCreate(&stream);
        const double t_all0 = now_ms();
        double t_proc       = 0.;
        {
            std::unique_ptr<cudaaligner::Aligner> aligner;
            if (algorithm == 0)
            {
                aligner = cudaaligner::create_aligner(max_query_length, max_target_length, n_pairs, cudaaligner::AlignmentType::global_alignment,
                                                      stream, 0);
            }
            else
            {
                DefaultDeviceAllocator allocator = create_default_device_allocator(int64_t(8) << 30);
                if (algorithm == 1)
                    aligner.reset(new cudaaligner::AlignerGlobalMyers(max_query_length, max_target_length, n_pairs, allocator, stream, 0));
                else
                    aligner.reset(new cudaaligner::AlignerGlobalUkkonen(max_query_length, max_target_length, n_pairs, allocator, stream, 0));
            }
            int64_t qo = 0, to = 0;
            for (int32_t i = 0; i < n_pairs; ++i)
            {
                cudaaligner::StatusType st = aligner->add_alignment(q_data + qo, q_len[i], t_data + to, t_len[i]);
                if (st != cudaaligner::StatusType::success)
                    throw std::runtime_error("reference add_alignment returned status " + std::to_string(static_cast<int>(st)));
                qo += q_len[i];
                to += t_len[i];
            }
            const double t0 = now_ms();
            aligner->align_all();
            aligner->sync_alignments();
            t_proc += now_ms() - t0;
            const auto& res = aligner->get_alignments();
            for (int32_t p = 0; p < static_cast<int32_t>(res.size()); ++p)
            {
                status[p]        = static_cast<int32_t>(res[p]->get_status());
                is_optimal[p]    = res[p]->is_optimal() ? 1 : 0;
                edit_distance[p] = res[p]->get_edit_distance();
                std::string cb   = res[p]->convert_to_cigar(cudaaligner::CigarFormat::basic);
                std::string ce   = res[p]->convert_to_cigar(cudaaligner::CigarFormat::extended);
                if (static_cast<int32_t>(cb.size()) >= cigar_stride || static_cast<int32_t>(ce.size()) >= cigar_stride)
                    throw std::runtime_error("cigar_stride too small");
                std::strcpy(cigar_basic + static_cast<int64_t>(p) * cigar_stride, cb.c_str());
                std::strcpy(cigar_ext + static_cast<int64_t>(p) * cigar_stride, ce.c_str());
            }
        }
        if (timings_ms)
        {
            timings_ms[0] = now_ms() - t_all0;
            timings_ms[1] = t_proc;
        }
        cudaStreamDestroy(stream);
        return 0;
    }
    catch (const std::exception& e)
    {
        if (errbuf && errbuf_len > 0)
        {
            std::strncpy(errbuf, e.what(), errbuf_len - 1);
            errbuf[errbuf_len - 1] = 0;
        }
        return -1;
    }
}

} // extern "C"
