// Plain C++14 client of the C ABI and the C++ adapter (no torch, no Python): built by tests/test_gpu_cpp.py with g++ and
// run on the GPU box. Reads one KAT and one tiny populate problem from stdin-free hardcoded data, prints results.
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>
#include "phmm_b200.hpp"

int main()
{
    using namespace octopus_b200;
    try {
        // reference KAT: sse2_band_size_8_check_alignments test 5 (test/unit/core/models/pair_hmm_tests.cpp:290-313)
        const std::string truth = "CCCCACGTATATATATATATATGGGGACGT", read = "CCCCACGTGGGACGT";
        std::vector<std::int8_t> quals(read.size(), 40), gap_open(truth.size(), 90);
        gap_open[8] = 70;
        GpuPairHMM<8> hmm;
        std::vector<char> a1(2 * truth.size() + 1, 0), a2(2 * truth.size() + 1, 0);
        int first_pos = -7;
        const int s1 = hmm.align(truth.data(), read.data(), quals.data(), (int)truth.size(), (int)read.size(), gap_open.data(), (short)1, (short)4);
        const int s2 = hmm.align(truth.data(), read.data(), quals.data(), (int)truth.size(), (int)read.size(), gap_open.data(), (short)1, (short)4,
                                 first_pos, a1.data(), a2.data());
        std::printf("KAT %d %d %d %s %s\n", s1, s2, first_pos, a1.data(), a2.data());
        // batch seam: 2 haplotypes x 2 reads
        HaplotypeBlock haps;
        ReadBlock reads;
        const std::string h0 = "ACGTTGCAAGCTTAGGCTAACGTTAGCATCGATCGGATCTAGCTAGGATCGATACGATCGATCGTAGCTAGCTAGTCGATCGATTTAGCGCGATATCGCGAT";
        std::string h1 = h0; h1[50] = h1[50] == 'A' ? 'C' : 'A';
        for (const auto& h : {h0, h1}) {
            std::vector<char> mf(h.begin(), h.end()), mr(h.begin(), h.end());
            std::vector<std::int8_t> p(h.size(), 50), go(h.size(), 40), ge(h.size(), 3);
            haps.add(h, mf, p, mr, p, go, ge, 0);
        }
        reads.add(h0.substr(30, 40), std::vector<std::uint8_t>(40, 30), 60, false, 30);
        reads.add(h1.substr(35, 40), std::vector<std::uint8_t>(40, 25), 60, true, 35);
        auto cfg = HaplotypeLikelihoodArray::default_config();
        cfg.max_indel_error = 16;
        cfg.map_positions = 0;   // the check below evaluates the original position only
        HaplotypeLikelihoodArray arr {cfg};
        arr.populate(reads, haps);
        for (std::size_t h = 0; h < 2; ++h) { const auto row = arr[h]; std::printf("ROW %zu %.17g %.17g\n", h, row[0], row[1]); }
        {   // multi-sample container semantics (haplotype_likelihood_array.cpp:200-409): samples are column ranges of one batch
            ReadBlock r0, r1;
            r0.add(h0.substr(30, 40), std::vector<std::uint8_t>(40, 30), 60, false, 30);
            r1.add(h1.substr(35, 40), std::vector<std::uint8_t>(40, 25), 60, true, 35);
            r1.add(h0.substr(30, 40), std::vector<std::uint8_t>(40, 30), 60, false, 30);
            HaplotypeLikelihoodArray multi {cfg};
            multi.populate(HaplotypeLikelihoodArray::ReadMap {{"S1", r0}, {"S2", r1}}, haps);
            bool ok = multi.samples().size() == 2 && multi.num_likelihoods("S1") == 1 && multi.num_likelihoods("S2") == 2 && !multi.is_primed();
            for (std::size_t h = 0; h < 2; ++h) {
                ok = ok && multi("S1", h)[0] == arr[h][0] && multi("S2", h)[0] == arr[h][1] && multi("S2", h)[1] == arr[h][0];
            }
            multi.prime("S2");
            ok = ok && multi.is_primed() && multi.num_likelihoods() == 2 && multi[1][0] == arr[1][1];
            const auto merged = multi.merge_samples();
            ok = ok && merged.samples().size() == 1 && merged.samples()[0] == "S1S2" && merged.is_primed() && merged.num_likelihoods() == 3 &&
                 merged[0][0] == arr[0][0] && merged[0][1] == arr[0][1] && merged[1][2] == arr[1][0];
            const auto only2 = multi.merge_samples({"S2"});
            ok = ok && only2.num_likelihoods() == 2 && only2[1][0] == arr[1][1];
            ok = ok && multi.extract_sample("S1").size() == 2 && multi.extract_sample("S1")[1][0] == arr[1][0];
            multi.reset({1});
            ok = ok && multi.num_haplotypes() == 1 && multi("S1", 0)[0] == arr[1][0];
            bool threw = false;
            try { multi("nope", 0); } catch (const std::out_of_range&) { threw = true; }
            // TemplateMap: a template's value is the sum of its reads' values (haplotype_likelihood_model.cpp:306-320)
            TemplateBlock tb;
            tb.add(reads);                  // one template holding both reads
            tb.add(r0);                     // and one holding the first read alone
            HaplotypeLikelihoodArray tarr {cfg};
            tarr.populate(HaplotypeLikelihoodArray::TemplateMap {{"S1", tb}}, haps);
            tarr.prime("S1");
            for (std::size_t h = 0; h < 2; ++h) ok = ok && tarr.num_likelihoods() == 2 && tarr[h][0] == arr[h][0] + arr[h][1] && tarr[h][1] == arr[h][0];
            std::printf("SAMPLES %s %s\n", ok ? "ok" : "MISMATCH", threw ? "throws" : "nothrow");
        }
        // ShortHaplotypeError must surface as the reference's exception type
        ReadBlock longread;
        longread.add(h0.substr(0, 95), std::vector<std::uint8_t>(95, 30), 60, false, 0);
        try { arr.populate(longread, haps); std::printf("SHORT none\n"); }
        catch (const ShortHaplotypeError& e) { std::printf("SHORT hap=%zu ext=%u\n", e.haplotype_index(), e.required_extension()); }
    } catch (const std::exception& e) {
        std::printf("EXCEPTION %s\n", e.what());
        return 1;
    }
    return 0;
}
