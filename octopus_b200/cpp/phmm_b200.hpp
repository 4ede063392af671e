// phmm_b200.hpp — C++ host-side adapter over the C ABI (include/phmm_b200.h), mirroring the reference's own
// interface for the pair-HMM path so that Octopus code compiles against it unchanged:
//
//   octopus_b200::GpuPairHMM<Band>   satisfies the duck-typed PairHMM concept that octopus::hmm::evaluate / align are
//                                    templated on (src/core/models/pairhmm/pair_hmm.hpp:373-380, 397-406, 441-451, 471-483,
//                                    542-551, 590-601; canonical signatures simd_pair_hmm.hpp:433-549):
//                                    band_size(), name(), 4 x align(), calculate_flank_score().
//   octopus_b200::HaplotypeLikelihoodArray  the batch seam: populate(reads, haplotypes, flank_state) fills the
//                                    [haplotype][read] matrix (haplotype_likelihood_array.hpp:65-72, .cpp:51-103) and throws
//                                    ShortHaplotypeError exactly where the reference does (haplotype_likelihood_model.cpp:238-256).
//
// Header-only; link with -lphmm_b200. All likelihood arithmetic runs on the GPU behind the C ABI; the only host
// arithmetic here is calculate_flank_score, which in the reference too is a scalar replay of two alignment strings
// (simd_pair_hmm.hpp:352-430), not part of the DP.
#ifndef PHMM_B200_HPP
#define PHMM_B200_HPP

#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "phmm_b200.h"

namespace octopus_b200 {

class Error : public std::runtime_error
{
public:
    Error(int code, const std::string& what) : std::runtime_error {what}, code_ {code} {}
    int code() const noexcept { return code_; }
private:
    int code_;
};

// HaplotypeLikelihoodModel::ShortHaplotypeError (haplotype_likelihood_model.hpp:123-139)
class ShortHaplotypeError : public std::runtime_error
{
public:
    ShortHaplotypeError(std::size_t haplotype_index, unsigned required_extension)
    : std::runtime_error {"Haplotype is too short for alignment"}, haplotype_index_ {haplotype_index}, required_extension_ {required_extension} {}
    std::size_t haplotype_index() const noexcept { return haplotype_index_; }
    unsigned required_extension() const noexcept { return required_extension_; }
private:
    std::size_t haplotype_index_;
    unsigned required_extension_;
};

// simd::PairHMMWrapper::TooLargeBandSizeError (simd_pair_hmm_wrapper.hpp:45-61)
class TooLargeBandSizeError : public std::runtime_error
{
public:
    explicit TooLargeBandSizeError(unsigned band) : std::runtime_error {"Band size too large"}, band_ {band} {}
    unsigned band() const noexcept { return band_; }
private:
    unsigned band_;
};

class Engine
{
public:
    explicit Engine(int device = -1)
    {
        phmm_engine* e = nullptr;
        const int rc = phmm_create(&e, device);
        if (rc != PHMM_OK) throw Error {rc, phmm_last_error(nullptr)};
        handle_.reset(e, [] (phmm_engine* p) { phmm_destroy(p); });
    }
    phmm_engine* get() const noexcept { return handle_.get(); }
    void check(int rc) const
    {
        if (rc == PHMM_OK) return;
        throw Error {rc, phmm_last_error(handle_.get())};
    }
private:
    std::shared_ptr<phmm_engine> handle_;
};

// ---------------------------------------------------------------------------------------------------------------------
// Per-call seam
// ---------------------------------------------------------------------------------------------------------------------
template <int BandSize>
class GpuPairHMM
{
public:
    using ScoreType = int;
    constexpr static char gap_label = '-';

    explicit GpuPairHMM(Engine engine = Engine {}) : engine_ {std::move(engine)} {}

    constexpr static const char* name() noexcept { return "B200"; }
    constexpr static int band_size() noexcept { return BandSize; }

    // score only, no SNV model (simd_pair_hmm.hpp:438-452)
    template <typename GapExtend>
    int align(const char* truth, const char* target, const std::int8_t* qualities, int truth_len, int target_len,
              const std::int8_t* gap_open, GapExtend gap_extend, short nuc_prior) const noexcept
    {
        int first_pos;
        return run(truth, target, qualities, truth_len, target_len, nullptr, nullptr, gap_open, gap_extend, nuc_prior, first_pos, nullptr, nullptr);
    }
    // score only, SNV model (:454-470)
    template <typename GapExtend>
    int align(const char* truth, const char* target, const std::int8_t* qualities, int truth_len, int target_len,
              const char* snv_mask, const std::int8_t* snv_prior, const std::int8_t* gap_open, GapExtend gap_extend, short nuc_prior) const noexcept
    {
        int first_pos;
        return run(truth, target, qualities, truth_len, target_len, snv_mask, snv_prior, gap_open, gap_extend, nuc_prior, first_pos, nullptr, nullptr);
    }
    // traceback, no SNV model (:472-489)
    template <typename GapExtend>
    int align(const char* truth, const char* target, const std::int8_t* qualities, int truth_len, int target_len,
              const std::int8_t* gap_open, GapExtend gap_extend, short nuc_prior, int& first_pos, char* align1, char* align2) const noexcept
    {
        return run(truth, target, qualities, truth_len, target_len, nullptr, nullptr, gap_open, gap_extend, nuc_prior, first_pos, align1, align2);
    }
    // traceback, SNV model (:491-509)
    template <typename GapExtend>
    int align(const char* truth, const char* target, const std::int8_t* qualities, int truth_len, int target_len,
              const char* snv_mask, const std::int8_t* snv_prior, const std::int8_t* gap_open, GapExtend gap_extend, short nuc_prior,
              int& first_pos, char* align1, char* align2) const noexcept
    {
        return run(truth, target, qualities, truth_len, target_len, snv_mask, snv_prior, gap_open, gap_extend, nuc_prior, first_pos, align1, align2);
    }

    // simd_pair_hmm.hpp:530-549 (SNV overload): replay of the alignment strings; scalar on the host in the reference too
    template <typename GapExtend>
    int calculate_flank_score(int truth_len, int lhs_flank_len, int rhs_flank_len, const char* target, const std::int8_t* quals,
                              const char* snv_mask, const std::int8_t* snv_prior, const std::int8_t* gap_open, GapExtend gap_extend,
                              short nuc_prior, int first_pos, const char* aln1, const char* aln2, int& target_mask_size) const noexcept
    {
        enum { M, I, D };
        int prev = M, truth_idx = first_pos, target_idx = 0, result = 0;
        const int rhs_begin = truth_len - rhs_flank_len;
        target_mask_size = 0;
        for (int a = 0; aln1[a]; ++a) {
            int st = M;
            if (aln1[a] == gap_label) st = I; else if (aln2[a] == gap_label) st = D;
            const bool in_flank = truth_idx < lhs_flank_len || truth_idx >= rhs_begin;
            if (st == M) {
                if (in_flank) {
                    if (aln1[a] != aln2[a]) {
                        if (aln1[a] != 'N') {
                            int q = quals[target_idx];
                            if (snv_mask[truth_idx] == target[target_idx] && snv_prior[truth_idx] < q) q = snv_prior[truth_idx];
                            result += q;
                        } else result += 2;
                    }
                    ++target_mask_size;
                }
                ++truth_idx; ++target_idx;
            } else if (st == I) {
                if (in_flank) { result += (prev == I ? get(gap_extend, truth_idx - 1) : gap_open[truth_idx - 1]) + nuc_prior; ++target_mask_size; }
                ++target_idx;
            } else {
                if (in_flank) result += prev == D ? get(gap_extend, truth_idx) : gap_open[truth_idx];
                ++truth_idx;
            }
            prev = st;
        }
        return result;
    }

private:
    Engine engine_;

    static int get(const std::int8_t* v, int i) noexcept { return v[i]; }
    template <typename T> static int get(T v, int) noexcept { return static_cast<int>(v); }
    static const std::int8_t* ext_ptr(const std::int8_t* p) noexcept { return p; }
    template <typename T> static const std::int8_t* ext_ptr(T) noexcept { return nullptr; }
    static int ext_scalar(const std::int8_t*) noexcept { return 0; }
    template <typename T> static int ext_scalar(T v) noexcept { return static_cast<int>(v); }

    template <typename GapExtend>
    int run(const char* truth, const char* target, const std::int8_t* quals, int truth_len, int target_len,
            const char* snv_mask, const std::int8_t* snv_prior, const std::int8_t* gap_open, GapExtend gap_extend, short nuc_prior,
            int& first_pos, char* align1, char* align2) const noexcept
    {
        int score = 0;
        std::vector<char> a1, a2;
        if (!align1) { a1.assign(2 * (target_len + BandSize) + 1, 0); a2 = a1; align1 = a1.data(); align2 = a2.data(); }
        const int rc = phmm_align_traceback(engine_.get(), BandSize, truth, target, quals, truth_len, target_len, snv_mask, snv_prior,
                                            gap_open, ext_ptr(gap_extend), ext_scalar(gap_extend), nuc_prior, &score, &first_pos, align1, align2);
        if (rc != PHMM_OK) first_pos = -1;   // the reference signals failure through first_pos (simd_pair_hmm.hpp:176-199)
        return score;
    }
};

// ---------------------------------------------------------------------------------------------------------------------
// Batch seam
// ---------------------------------------------------------------------------------------------------------------------
struct HaplotypeBlock   // what HaplotypeLikelihoodModel::reset derives per haplotype (haplotype_likelihood_model.cpp:60-78)
{
    std::vector<std::int64_t> off {0};
    std::string seq, snv_mask_fwd, snv_mask_rev;
    std::vector<std::int8_t> snv_prior_fwd, snv_prior_rev, gap_open, gap_extend;
    std::vector<std::int64_t> begin;

    void add(const std::string& sequence, const std::vector<char>& fwd_mask, const std::vector<std::int8_t>& fwd_priors,
             const std::vector<char>& rev_mask, const std::vector<std::int8_t>& rev_priors,
             const std::vector<std::int8_t>& open, const std::vector<std::int8_t>& extend, std::int64_t mapped_begin = 0)
    {
        seq += sequence;
        snv_mask_fwd.append(fwd_mask.begin(), fwd_mask.end());
        snv_mask_rev.append(rev_mask.begin(), rev_mask.end());
        snv_prior_fwd.insert(snv_prior_fwd.end(), fwd_priors.begin(), fwd_priors.end());
        snv_prior_rev.insert(snv_prior_rev.end(), rev_priors.begin(), rev_priors.end());
        gap_open.insert(gap_open.end(), open.begin(), open.end());
        gap_extend.insert(gap_extend.end(), extend.begin(), extend.end());
        begin.push_back(mapped_begin);
        off.push_back(static_cast<std::int64_t>(seq.size()));
    }
    std::size_t size() const noexcept { return off.size() - 1; }
    phmm_haplotypes view() const noexcept
    {
        return phmm_haplotypes {static_cast<std::int32_t>(size()), off.data(), seq.data(), snv_mask_fwd.data(), snv_prior_fwd.data(),
                                snv_mask_rev.data(), snv_prior_rev.data(), gap_open.data(), gap_extend.data(), begin.data()};
    }
};

struct ReadBlock   // AlignedRead fields on the path (basics/aligned_read.hpp:36-39,120-146)
{
    std::vector<std::int64_t> off {0};
    std::string bases;
    std::vector<std::uint8_t> quals, mapq, reverse;
    std::vector<std::int64_t> begin;

    void add(const std::string& sequence, const std::vector<std::uint8_t>& base_qualities, std::uint8_t mapping_quality,
             bool is_marked_reverse_mapped, std::int64_t mapped_begin)
    {
        bases += sequence;
        quals.insert(quals.end(), base_qualities.begin(), base_qualities.end());
        mapq.push_back(mapping_quality);
        reverse.push_back(is_marked_reverse_mapped ? 1 : 0);
        begin.push_back(mapped_begin);
        off.push_back(static_cast<std::int64_t>(bases.size()));
    }
    std::size_t size() const noexcept { return off.size() - 1; }
    phmm_reads view() const noexcept
    {
        return phmm_reads {static_cast<std::int32_t>(size()), off.data(), bases.data(), quals.data(), mapq.data(), reverse.data(), begin.data()};
    }
};

struct FlankState { std::int64_t lhs_flank, rhs_flank; };   // HaplotypeLikelihoodModel::FlankState

class HaplotypeLikelihoodArray
{
public:
    using LogProbability = double;
    using LikelihoodVector = std::vector<LogProbability>;

    explicit HaplotypeLikelihoodArray(phmm_config config, Engine engine = Engine {}) : engine_ {std::move(engine)}, config_ {config}
    {
        if (config.max_indel_error > 256) throw TooLargeBandSizeError {static_cast<unsigned>(config.max_indel_error)};
    }
    static phmm_config default_config() noexcept { phmm_config c; phmm_default_config(&c); return c; }

    // haplotype_likelihood_array.cpp:51-103. positions: optional candidate mapping positions (CSR over [H][R]); flank: optional.
    void populate(const ReadBlock& reads, const HaplotypeBlock& haplotypes, const FlankState* flank_state = nullptr,
                  const phmm_positions* positions = nullptr)
    {
        const auto hv = haplotypes.view();
        const auto rv = reads.view();
        num_reads_ = reads.size();
        likelihoods_.assign(haplotypes.size() * reads.size(), 0.0);
        std::vector<std::int32_t> status(likelihoods_.size(), 0);
        phmm_flank_state fs {flank_state ? 1 : 0, flank_state ? flank_state->lhs_flank : 0, flank_state ? flank_state->rhs_flank : 0};
        const int rc = phmm_populate(engine_.get(), &config_, &hv, &rv, positions, &fs, likelihoods_.data(), status.data(), PHMM_SPACE_HOST);
        if (rc == PHMM_ERR_SHORT_HAPLOTYPE) {
            for (std::size_t i = 0; i < status.size(); ++i) {
                if ((status[i] & 0xFFFF) == PHMM_STATUS_SHORT_HAP) throw ShortHaplotypeError {i / num_reads_, static_cast<unsigned>(status[i] >> 16)};
            }
        }
        engine_.check(rc);
    }
    // likelihoods_[haplotype][sample] for the single sample (haplotype_likelihood_array.cpp:212-236)
    LikelihoodVector operator[](std::size_t haplotype_index) const
    {
        const auto first = likelihoods_.begin() + static_cast<std::ptrdiff_t>(haplotype_index * num_reads_);
        return LikelihoodVector(first, first + static_cast<std::ptrdiff_t>(num_reads_));
    }
    const double* data() const noexcept { return likelihoods_.data(); }
    std::size_t num_likelihoods() const noexcept { return num_reads_; }
    bool is_empty() const noexcept { return likelihoods_.empty(); }
    void clear() noexcept { likelihoods_.clear(); num_reads_ = 0; }

private:
    Engine engine_;
    phmm_config config_;
    std::vector<double> likelihoods_;
    std::size_t num_reads_ = 0;
};

} // namespace octopus_b200

#endif
