// C++ API test: the reference's own usage patterns (cudapoa/samples/sample_cudapoa.cpp, cudaaligner/samples,
// Test_CudapoaBatch.cu:70-203, Test_AlignerGlobal.cpp:73-157) compiled against include/claraparabricks/... and run on the GPU.
// Prints "CPP_API_OK" on success.
#include <claraparabricks/genomeworks/cudaaligner/aligner.hpp>
#include <claraparabricks/genomeworks/cudaaligner/alignment.hpp>
#include <claraparabricks/genomeworks/cudapoa/batch.hpp>
#include <claraparabricks/genomeworks/cudapoa/cudapoa.hpp>
#include <claraparabricks/genomeworks/utils/allocator.hpp>
#include <claraparabricks/genomeworks/utils/genomeutils.hpp>

#include <cstdio>
#include <cstdlib>
#include <iostream>

using namespace claraparabricks::genomeworks;

#define REQUIRE(cond)                                                         \
    do                                                                        \
    {                                                                         \
        if (!(cond))                                                          \
        {                                                                     \
            std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); \
            return 1;                                                         \
        }                                                                     \
    } while (0)

int main()
{
    // ---- cudapoa
    cudapoa::Init();
    {
        cudapoa::BatchConfig cfg(1024, 10, 256, cudapoa::BandMode::static_band);
        REQUIRE(cfg.max_nodes_per_graph == 3072 && cfg.matrix_sequence_dimension == 264 && cfg.max_consensus_size == 2048);
        bool threw = false;
        try
        {
            cudapoa::create_batch(0, nullptr, 0, cudapoa::OutputType::consensus, cfg, -8, -6, 8);
        }
        catch (const std::runtime_error&)
        {
            threw = true;
        }
        REQUIRE(threw); // zero memory => std::runtime_error (Test_CudapoaBatch.cu)
        auto batch = cudapoa::create_batch(0, nullptr, 1ll << 30, cudapoa::OutputType::consensus, cfg, -8, -6, 8);
        std::string a(1023, 'A');
        cudapoa::Group g;
        for (int i = 0; i < 3; i++)
            g.push_back(cudapoa::Entry{a.c_str(), nullptr, static_cast<int32_t>(a.size())});
        std::vector<cudapoa::StatusType> per_seq;
        REQUIRE(batch->add_poa_group(per_seq, g) == cudapoa::StatusType::success);
        REQUIRE(per_seq.size() == 3);
        std::minstd_rand rng(1);
        std::string backbone          = genomeutils::generate_random_genome(300, rng);
        std::vector<std::string> seqs = genomeutils::generate_random_sequences(backbone, 8, rng, 6, 3, 3);
        cudapoa::Group g2;
        for (auto& s : seqs)
            g2.push_back(cudapoa::Entry{s.c_str(), nullptr, static_cast<int32_t>(s.size())});
        REQUIRE(batch->add_poa_group(per_seq, g2) == cudapoa::StatusType::success);
        REQUIRE(batch->get_total_poas() == 2);
        batch->generate_poa();
        std::vector<std::string> cons;
        std::vector<std::vector<uint16_t>> cov;
        std::vector<cudapoa::StatusType> st;
        REQUIRE(batch->get_consensus(cons, cov, st) == cudapoa::StatusType::success);
        REQUIRE(cons.size() == 2 && st[0] == cudapoa::StatusType::success && st[1] == cudapoa::StatusType::success);
        REQUIRE(cons[0] == a && cov[0].size() == a.size() && cov[0][10] == 3);
        REQUIRE(cons[1].size() > 250);
        std::vector<std::vector<std::string>> msa;
        REQUIRE(batch->get_msa(msa, st) == cudapoa::StatusType::output_type_unavailable);
        std::vector<DirectedGraph> graphs;
        st.clear();
        batch->get_graphs(graphs, st);
        REQUIRE(graphs.size() == 2 && graphs[0].get_node_ids().size() == 1023 && graphs[0].get_edges().size() == 1022);
        REQUIRE(!graphs[1].serialize_to_dot().empty());
        batch->reset();
        REQUIRE(batch->get_total_poas() == 0);
        std::string msg, hint;
        cudapoa::decode_error(cudapoa::StatusType::exceeded_maximum_poas, msg, hint);
        REQUIRE(!msg.empty());
    }
    {
        // MSA output
        cudapoa::BatchConfig cfg(1024, 10, 256, cudapoa::BandMode::adaptive_band);
        auto batch = cudapoa::create_batch(0, nullptr, 1ll << 30, cudapoa::OutputType::msa, cfg, -8, -6, 8);
        std::minstd_rand rng(3);
        std::string backbone          = genomeutils::generate_random_genome(120, rng);
        std::vector<std::string> seqs = genomeutils::generate_random_sequences(backbone, 6, rng, 4, 2, 2);
        cudapoa::Group g;
        for (auto& s : seqs)
            g.push_back(cudapoa::Entry{s.c_str(), nullptr, static_cast<int32_t>(s.size())});
        std::vector<cudapoa::StatusType> per_seq, st;
        REQUIRE(batch->add_poa_group(per_seq, g) == cudapoa::StatusType::success);
        batch->generate_poa();
        std::vector<std::vector<std::string>> msa;
        REQUIRE(batch->get_msa(msa, st) == cudapoa::StatusType::success);
        REQUIRE(msa.size() == 1 && msa[0].size() == 6);
        for (size_t r = 0; r < 6; r++)
        {
            std::string stripped;
            for (char c : msa[0][r])
                if (c != '-')
                    stripped += c;
            REQUIRE(stripped == seqs[r]);
            REQUIRE(msa[0][r].size() == msa[0][0].size());
        }
    }
    // ---- cudaaligner
    cudaaligner::Init();
    {
        auto aligner = cudaaligner::create_aligner(cudaaligner::AlignmentType::global_alignment, 1024, nullptr, 0, 1ll << 30);
        const char* q[] = {"AAAA", "ATAAAAAAAA", "AAAAAAAAA", "ACTGA"};
        const char* t[] = {"TTAT", "AAAAAAAAA", "ATAAAAAAAA", "GCTAG"};
        const char* c[] = {"4M", "1M1D8M", "1M1I8M", "3M1D1M1I"};
        const int e[]   = {3, 1, 1, 3};
        for (int i = 0; i < 4; i++)
            REQUIRE(aligner->add_alignment(q[i], static_cast<int32_t>(std::string(q[i]).size()), t[i], static_cast<int32_t>(std::string(t[i]).size())) ==
                    cudaaligner::StatusType::success);
        REQUIRE(aligner->num_alignments() == 4);
        aligner->align_all();
        aligner->sync_alignments();
        const auto& res = aligner->get_alignments();
        REQUIRE(res.size() == 4);
        for (int i = 0; i < 4; i++)
        {
            REQUIRE(res[i]->get_status() == cudaaligner::StatusType::success);
            REQUIRE(res[i]->convert_to_cigar() == c[i]);
            REQUIRE(res[i]->get_edit_distance() == e[i]);
            REQUIRE(res[i]->is_optimal());
        }
        REQUIRE(res[1]->convert_to_cigar(cudaaligner::CigarFormat::extended) == "1=1D8=");
        bool threw = false;
        try
        {
            aligner->reset_max_bandwidth(33);
        }
        catch (const std::invalid_argument&)
        {
            threw = true;
        }
        REQUIRE(threw);
        auto legacy = cudaaligner::create_aligner(10, 10, 2, cudaaligner::AlignmentType::global_alignment, nullptr, 0, 1ll << 30);
        REQUIRE(legacy->add_alignment("AAATC", 5, "TACGTTTT", 8) == cudaaligner::StatusType::success);
        REQUIRE(legacy->add_alignment("AAAAAAAAAAA", 11, "A", 1) == cudaaligner::StatusType::exceeded_max_length);
        // AlignerGlobal semantics (aligner_global.cpp:78-141,162-190), which the Cython shim relies on (cudaaligner.pyx:237-243):
        // the Alignment exists from add_alignment() on and sync_alignments() fills it in place
        REQUIRE(legacy->get_alignments().size() == 1 && legacy->num_alignments() == 1);
        std::shared_ptr<cudaaligner::Alignment> before = legacy->get_alignments()[0];
        REQUIRE(before->get_status() == cudaaligner::StatusType::uninitialized);
        legacy->align_all();
        legacy->sync_alignments();
        REQUIRE(legacy->get_alignments()[0] == before && before->get_status() == cudaaligner::StatusType::success);
        REQUIRE(before->convert_to_cigar() == "3M1I2M2I");
        REQUIRE(before->get_alignment().size() == 8); // one AlignmentState per step
        REQUIRE(legacy->add_alignment("TGCA", 4, "ATACGCT", 7) == cudaaligner::StatusType::success);
        REQUIRE(legacy->add_alignment("TGCA", 4, "ATACGCT", 7) == cudaaligner::StatusType::exceeded_max_alignments);
        legacy->align_all();
        legacy->sync_alignments();
        REQUIRE(legacy->get_alignments().size() == 2 && legacy->get_alignments()[1]->convert_to_cigar() == "1I1M2I3M");
        legacy->reset();
        REQUIRE(legacy->get_alignments().empty());
    }
    // ---- the allocator that is part of the API (utils/allocator.hpp:208-358): one pool, copies share it, Batch and Aligner carve from it
    {
        DefaultDeviceAllocator alloc = create_default_device_allocator(3ull << 30);
        const int64_t whole          = get_size_of_largest_free_memory_block(alloc);
        REQUIRE(whole >= (3ll << 30) - 256);
        DefaultDeviceAllocator copy = alloc;
        char* p                     = copy.allocate(1 << 20);
        REQUIRE(p != nullptr && alloc.get_size_of_largest_free_memory_block() == whole - (1 << 20));
        CachingDeviceAllocator<int32_t, details::DevicePreallocatedAllocator> rebound(alloc);
        int32_t* q = rebound.allocate(256);
        REQUIRE(reinterpret_cast<char*>(q) == p + (1 << 20));
        rebound.deallocate(q, 256);
        copy.deallocate(p, 1 << 20);
        REQUIRE(alloc.get_size_of_largest_free_memory_block() == whole);
        REQUIRE(alloc.memory_resource() == copy.memory_resource());
        {
            cudapoa::BatchConfig cfg(1024, 10, 256, cudapoa::BandMode::static_band);
            auto b1 = cudapoa::create_batch(0, nullptr, alloc, 1ll << 30, cudapoa::OutputType::consensus, cfg, -8, -6, 8);
            auto b2 = cudapoa::create_batch(0, nullptr, alloc, 1ll << 30, cudapoa::OutputType::consensus, cfg, -8, -6, 8);
            REQUIRE(alloc.get_size_of_largest_free_memory_block() == whole - (2ll << 30)); // two batches, two disjoint blocks
            bool threw = false;
            try
            {
                cudapoa::create_batch(0, nullptr, alloc, 2ll << 30, cudapoa::OutputType::consensus, cfg, -8, -6, 8);
            }
            catch (const device_memory_allocation_exception&)
            {
                threw = true;
            }
            REQUIRE(threw);
            std::string a(500, 'C');
            cudapoa::Group g;
            for (int i = 0; i < 4; i++)
                g.push_back(cudapoa::Entry{a.c_str(), nullptr, static_cast<int32_t>(a.size())});
            std::vector<cudapoa::StatusType> per_seq, st;
            REQUIRE(b1->add_poa_group(per_seq, g) == cudapoa::StatusType::success);
            REQUIRE(b2->add_poa_group(per_seq, g) == cudapoa::StatusType::success);
            b1->generate_poa();
            b2->generate_poa();
            std::vector<std::string> cons;
            std::vector<std::vector<uint16_t>> cov;
            REQUIRE(b2->get_consensus(cons, cov, st) == cudapoa::StatusType::success && cons[0] == a);
            REQUIRE(b1->get_consensus(cons, cov, st) == cudapoa::StatusType::success && cons[1] == a);
            auto al = cudaaligner::create_aligner(cudaaligner::AlignmentType::global_alignment, 64, nullptr, 0, alloc, -1);
            REQUIRE(al->add_alignment("AAATC", 5, "TACGTTTT", 8) == cudaaligner::StatusType::success);
            al->align_all();
            al->sync_alignments();
            REQUIRE(al->get_alignments()[0]->convert_to_cigar() == "3M1I2M2I");
            REQUIRE(al->get_device_allocator().memory_resource() == alloc.memory_resource());
            REQUIRE(alloc.get_size_of_largest_free_memory_block() < whole - (2ll << 30)); // the aligner's buffers are pool blocks too
        }
        REQUIRE(alloc.get_size_of_largest_free_memory_block() == whole); // everything went back to the pool
    }
    std::printf("CPP_API_OK\n");
    return 0;
}
