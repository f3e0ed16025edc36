// CPU-only checks of the cudapoa host utilities (include/claraparabricks/genomeworks/cudapoa/utils.hpp): window-file and FASTA
// parsing, resize_windows, golden-value parsing. Semantics as in the reference's utils.hpp:77-185. Prints "ok" on success.
#include <claraparabricks/genomeworks/cudapoa/utils.hpp>

#include <cstdio>
#include <fstream>
#include <iostream>

using namespace claraparabricks::genomeworks::cudapoa;

#define EXPECT(c)                                                        \
    do                                                                   \
    {                                                                    \
        if (!(c))                                                        \
        {                                                                \
            std::cerr << "FAILED: " #c " (line " << __LINE__ << ")\n";   \
            return 1;                                                    \
        }                                                                \
    } while (0)

int main(int argc, char** argv)
{
    const std::string dir = argc > 1 ? argv[1] : "/tmp";
    const std::string wf  = dir + "/gwb200_windows.txt";
    {
        std::ofstream o(wf);
        o << "2\nACGT\nACGA\n3\nTTTT\nTTTA\nTTAA\n1\nGG\n";
    }
    std::vector<std::vector<std::string>> w;
    parse_cudapoa_file(w, wf, -1);
    EXPECT(w.size() == 3 && w[0].size() == 2 && w[1].size() == 3 && w[2].size() == 1);
    EXPECT(w[1][2] == "TTAA" && w[2][0] == "GG");
    // truncate
    std::vector<std::vector<std::string>> t;
    parse_cudapoa_file(t, wf, 2);
    EXPECT(t.size() == 2 && t[1][0] == "TTTT");
    // repeat cyclically, in order
    std::vector<std::vector<std::string>> r;
    parse_cudapoa_file(r, wf, 8);
    EXPECT(r.size() == 8);
    EXPECT(r[3] == r[0] && r[4] == r[1] && r[5] == r[2] && r[6] == r[0] && r[7] == r[1]);
    // missing file
    bool threw = false;
    try
    {
        parse_cudapoa_file(w, dir + "/does_not_exist.txt", -1);
    }
    catch (const std::runtime_error&)
    {
        threw = true;
    }
    EXPECT(threw);
    // FASTA: one window per file, multi-line records
    const std::string f1 = dir + "/gwb200_a.fa", f2 = dir + "/gwb200_b.fa";
    {
        std::ofstream o(f1);
        o << ">r1 desc\nACGT\nAC\n>r2\nGGGG\n";
        std::ofstream p(f2);
        p << ">x\nTT\n";
    }
    std::vector<std::vector<std::string>> fw;
    parse_fasta_files(fw, {f1, f2}, -1);
    EXPECT(fw.size() == 2 && fw[0].size() == 2 && fw[0][0] == "ACGTAC" && fw[0][1] == "GGGG" && fw[1].size() == 1 && fw[1][0] == "TT");
    // golden value
    const std::string gf = dir + "/gwb200_golden.txt";
    {
        std::ofstream o(gf);
        o << "ACGTACGT\nsecond line\n";
    }
    EXPECT(parse_golden_value_file(gf) == "ACGTACGT");
    std::remove(wf.c_str());
    std::remove(f1.c_str());
    std::remove(f2.c_str());
    std::remove(gf.c_str());
    std::cout << "ok" << std::endl;
    return 0;
}
