// TEST / BASELINE INFRASTRUCTURE ONLY -- not part of the product path.
//
// extern "C" driver around the UNMODIFIED 3rdparty/spoa (pinned 0b9e56da, .SUBMODULES.json) compiled by
// oracle/Makefile from /root/reference into oracle/_ref/libspoa_ref.so. This is the CPU baseline named by
// BASELINE.json's north_star ("next to 3rdparty/spoa timed on the box's own host cores"); usage pattern
// follows cudapoa/tests/Test_CudapoaGenerateMSA2.cu:60-79 (createAlignmentEngine(kNW, match, mismatch, gap),
// createGraph, align, add_alignment, generate_consensus). spoa is a timing/plumbing reference, not a
// bit-exact oracle for cudapoa.
#include <spoa/spoa.hpp>

#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

extern "C" {

// Consensus for n_windows windows over n_threads host threads (dynamic work queue).
//   consensus: [n_windows * consensus_stride] NUL-terminated (may be NULL to skip copy-out).
//   cells_out: spoa DP cells = sum over reads>=1 of (graph nodes at alignment time x read length).
// Returns wall seconds.
double spoa_consensus_run(int32_t n_windows, const int32_t* win_nseq, const int32_t* seq_len, const char* seq_data,
                          int32_t match, int32_t mismatch, int32_t gap, int32_t n_threads,
                          char* consensus, int32_t consensus_stride, double* cells_out)
{
    std::vector<int64_t> seq_off;
    std::vector<int32_t> win_first(n_windows);
    {
        int64_t off = 0;
        int32_t si  = 0;
        for (int32_t w = 0; w < n_windows; ++w)
        {
            win_first[w] = si;
            for (int32_t s = 0; s < win_nseq[w]; ++s)
            {
                seq_off.push_back(off);
                off += seq_len[si++];
            }
        }
    }
    if (n_threads <= 0)
        n_threads = static_cast<int32_t>(std::thread::hardware_concurrency());
    if (n_threads <= 0)
        n_threads = 1;
    std::atomic<int32_t> next(0);
    std::vector<double> cells(n_threads, 0.);
    auto t0     = std::chrono::steady_clock::now();
    auto worker = [&](int32_t tid) {
        auto engine = spoa::createAlignmentEngine(spoa::AlignmentType::kNW, static_cast<int8_t>(match), static_cast<int8_t>(mismatch), static_cast<int8_t>(gap));
        while (true)
        {
            const int32_t w = next.fetch_add(1);
            if (w >= n_windows)
                break;
            auto graph = spoa::createGraph();
            for (int32_t s = 0; s < win_nseq[w]; ++s)
            {
                const int32_t si = win_first[w] + s;
                std::string seq(seq_data + seq_off[si], static_cast<size_t>(seq_len[si]));
                if (s > 0)
                    cells[tid] += static_cast<double>(graph->nodes().size()) * static_cast<double>(seq_len[si]);
                auto alignment = engine->align(seq, graph);
                graph->add_alignment(alignment, seq);
            }
            std::string c = graph->generate_consensus();
            if (consensus)
            {
                char* dst = consensus + static_cast<int64_t>(w) * consensus_stride;
                std::strncpy(dst, c.c_str(), consensus_stride - 1);
                dst[consensus_stride - 1] = 0;
            }
        }
    };
    std::vector<std::thread> threads;
    for (int32_t t = 0; t < n_threads; ++t)
        threads.emplace_back(worker, t);
    for (auto& t : threads)
        t.join();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (cells_out)
    {
        double tot = 0.;
        for (double c : cells)
            tot += c;
        *cells_out = tot;
    }
    return secs;
}

int32_t spoa_hardware_threads()
{
    return static_cast<int32_t>(std::thread::hardware_concurrency());
}

} // extern "C"
