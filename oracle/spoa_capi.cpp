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
#include <algorithm>
#include <cstring>
#include <memory>
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

// Multiple sequence alignment of ONE window (the pattern of cudapoa/tests/Test_CudapoaGenerateMSA2.cu:60-79): rows are written
// NUL-terminated into msa[n_seqs * stride]. Returns the MSA width, -1 when stride is too small.
int32_t spoa_msa_run(int32_t n_seqs, const int32_t* seq_len, const char* seq_data, int32_t match, int32_t mismatch, int32_t gap, char* msa,
                     int32_t stride)
{
    auto engine = spoa::createAlignmentEngine(spoa::AlignmentType::kNW, static_cast<int8_t>(match), static_cast<int8_t>(mismatch), static_cast<int8_t>(gap));
    auto graph  = spoa::createGraph();
    int64_t off = 0;
    for (int32_t s = 0; s < n_seqs; ++s)
    {
        std::string seq(seq_data + off, static_cast<size_t>(seq_len[s]));
        off += seq_len[s];
        auto alignment = engine->align(seq, graph);
        graph->add_alignment(alignment, seq);
    }
    std::vector<std::string> rows;
    graph->generate_multiple_sequence_alignment(rows);
    int32_t width = 0;
    for (int32_t s = 0; s < static_cast<int32_t>(rows.size()); ++s)
    {
        if (static_cast<int32_t>(rows[s].size()) >= stride)
            return -1;
        std::strcpy(msa + static_cast<int64_t>(s) * stride, rows[s].c_str());
        width = static_cast<int32_t>(rows[s].size());
    }
    return width;
}

// ---- streaming interface: FULL windows at a bounded cost per step ----------------------------------------------------------
// One window of the long-read workload costs spoa minutes per core, far more than a benchmark step may take. A stream keeps one
// window in progress per host thread and every step() call fuses the next `reads_per_step` reads of every thread's window (a
// finished window produces its consensus and the thread moves on to its next window). Over 31 read positions every read of a
// full 32-read window is aligned against the true graph it meets, so per-position times add up to the real full-window cost.
struct SpoaStream
{
    int32_t n_windows = 0, n_threads = 0;
    std::vector<int32_t> win_nseq, seq_len, win_first;
    std::vector<int64_t> seq_off;
    std::string data;
    int8_t match = 8, mismatch = -6, gap = -8;
    struct Slot
    {
        std::unique_ptr<spoa::AlignmentEngine> engine;
        std::unique_ptr<spoa::Graph> graph;
        int32_t window = 0; // index into the thread's own window sequence: tid, tid + n_threads, ...
        int32_t next_read = 0;
        int64_t windows_done = 0;
    };
    std::vector<Slot> slots;
};

void* spoa_stream_create(int32_t n_windows, const int32_t* win_nseq, const int32_t* seq_len, const char* seq_data, int32_t match,
                         int32_t mismatch, int32_t gap, int32_t n_threads)
{
    SpoaStream* s = new SpoaStream;
    s->n_windows  = n_windows;
    s->win_nseq.assign(win_nseq, win_nseq + n_windows);
    s->win_first.resize(n_windows);
    int64_t off = 0;
    int32_t si  = 0;
    for (int32_t w = 0; w < n_windows; ++w)
    {
        s->win_first[w] = si;
        for (int32_t k = 0; k < win_nseq[w]; ++k)
        {
            s->seq_off.push_back(off);
            off += seq_len[si++];
        }
    }
    s->seq_len.assign(seq_len, seq_len + si);
    s->data.assign(seq_data, static_cast<size_t>(off));
    if (n_threads <= 0)
        n_threads = static_cast<int32_t>(std::thread::hardware_concurrency());
    s->n_threads = std::max(1, std::min(n_threads, n_windows));
    s->match     = static_cast<int8_t>(match);
    s->mismatch  = static_cast<int8_t>(mismatch);
    s->gap       = static_cast<int8_t>(gap);
    s->slots.resize(s->n_threads);
    for (int32_t t = 0; t < s->n_threads; ++t)
    {
        s->slots[t].engine = spoa::createAlignmentEngine(spoa::AlignmentType::kNW, s->match, s->mismatch, s->gap);
        s->slots[t].graph  = spoa::createGraph();
        s->slots[t].window = t;
    }
    return s;
}

int32_t spoa_stream_threads(void* h) { return static_cast<SpoaStream*>(h)->n_threads; }

// Fuses the next reads_per_step reads on every thread. Returns wall seconds of the step.
//   pos_seconds / pos_cells / pos_count: [max_pos] accumulated over all threads per read position (index of the read inside its
//   window; position 0 = backbone), so that the caller can add up one full window; windows_done: completed windows so far.
double spoa_stream_step(void* h, int32_t reads_per_step, int32_t max_pos, double* pos_seconds, double* pos_cells, int64_t* pos_count,
                        int64_t* windows_done)
{
    SpoaStream* s = static_cast<SpoaStream*>(h);
    std::vector<std::vector<double>> tsec(s->n_threads, std::vector<double>(max_pos, 0.)), tcel(s->n_threads, std::vector<double>(max_pos, 0.));
    std::vector<std::vector<int64_t>> tcnt(s->n_threads, std::vector<int64_t>(max_pos, 0));
    auto t0     = std::chrono::steady_clock::now();
    auto worker = [&](int32_t tid) {
        SpoaStream::Slot& sl = s->slots[tid];
        for (int32_t k = 0; k < reads_per_step; ++k)
        {
            const int32_t w  = sl.window % s->n_windows;
            const int32_t si = s->win_first[w] + sl.next_read;
            std::string seq(s->data.data() + s->seq_off[si], static_cast<size_t>(s->seq_len[si]));
            const auto a0      = std::chrono::steady_clock::now();
            const double cells = sl.next_read > 0 ? static_cast<double>(sl.graph->nodes().size()) * static_cast<double>(seq.size()) : 0.;
            auto alignment     = sl.engine->align(seq, sl.graph);
            sl.graph->add_alignment(alignment, seq);
            const int32_t pos = sl.next_read;
            sl.next_read++;
            if (sl.next_read == s->win_nseq[w])
            {
                std::string c = sl.graph->generate_consensus(); // part of the window's cost, charged to its last read
                (void)c;
                sl.graph     = spoa::createGraph();
                sl.next_read = 0;
                sl.window += s->n_threads;
                sl.windows_done++;
            }
            const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - a0).count();
            if (pos < max_pos)
            {
                tsec[tid][pos] += dt;
                tcel[tid][pos] += cells;
                tcnt[tid][pos] += 1;
            }
        }
    };
    std::vector<std::thread> threads;
    for (int32_t t = 0; t < s->n_threads; ++t)
        threads.emplace_back(worker, t);
    for (auto& t : threads)
        t.join();
    const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    int64_t done      = 0;
    for (int32_t t = 0; t < s->n_threads; ++t)
    {
        done += s->slots[t].windows_done;
        for (int32_t p = 0; p < max_pos; ++p)
        {
            pos_seconds[p] += tsec[t][p];
            pos_cells[p] += tcel[t][p];
            pos_count[p] += tcnt[t][p];
        }
    }
    if (windows_done)
        *windows_done = done;
    return secs;
}

void spoa_stream_destroy(void* h) { delete static_cast<SpoaStream*>(h); }

int32_t spoa_hardware_threads()
{
    return static_cast<int32_t>(std::thread::hardware_concurrency());
}

} // extern "C"
