// TEST INFRASTRUCTURE ONLY — never linked into or called from the product path.
//
// extern "C" driver around the UNMODIFIED reference layer ABOVE the SIMD kernel, compiled from where it lies:
//   src/core/models/pairhmm/pair_hmm.hpp           hmm::evaluate (:827-841: try_naive_evaluate :275-319, simd_evaluate :694-782),
//                                                  hmm::align (:858-874: try_naive_align :321-340, simd_align :784-823, make_cigar :152-188)
//   src/core/models/pairhmm/simd_pair_hmm_wrapper.hpp   PairHMMWrapper: runtime band / precision choice (:85-88, :209-241)
//   src/utils/kmer_mapper.hpp                      the candidate-position mapper populate() runs inline
//   src/core/models/haplotype_likelihood_model.cpp HaplotypeLikelihoodModel::{reset, evaluate, align} (compiled alongside, with
//                                                  error/{snv,indel}_error_model.cpp) over stand-in Haplotype / AlignedRead types
//   src/core/models/haplotype_likelihood_array.cpp HaplotypeLikelihoodArray::populate (ReadMap and TemplateMap) — the batch seam itself
// pair_hmm.hpp's own includes that need Boost or Octopus's config (basics/cigar_string.hpp, exceptions/*.hpp, utils/maths.hpp,
// <boost/variant.hpp>) resolve to the minimal stand-ins under oracle/ref_shim/ (each says what it replaces); the reference
// headers themselves are untouched. Needs -std=c++17 (the boost::variant stand-in is std::variant).
//
// Used by tests/ only: it pins oracle/phmm_oracle.c's restatement of hmm::evaluate / hmm::align / the band rounding to the
// reference's own code, with the MutationModel parameter set HaplotypeLikelihoodModel uses (haplotype_likelihood_model.hpp:106).

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "core/models/pairhmm/pair_hmm.hpp"
#include "utils/kmer_mapper.hpp"          // self-contained (std only): the K = 6 vote mapper, utils/kmer_mapper.hpp:43-159
#include "core/models/error/error_model_factory.hpp"   // stand-in: "error models" that hand back the arrays set below
#include "core/models/haplotype_likelihood_model.hpp"  // the reference's own; its .cpp is compiled alongside (oracle/Makefile)

#include "core/models/haplotype_likelihood_array.hpp"  // likewise (its .cpp and utils/thread_pool.cpp are compiled alongside)

#include <unordered_map>

namespace {
thread_local octopus::FixedPenalties default_penalties;
thread_local std::unordered_map<const octopus::Haplotype*, octopus::FixedPenalties> block_penalties;
} // namespace
namespace octopus {
const FixedPenalties& fixed_penalties(const Haplotype& haplotype) noexcept
{
    const auto itr = block_penalties.find(&haplotype);
    return itr != block_penalties.end() ? itr->second : default_penalties;
}
} // namespace octopus

namespace {

using namespace octopus::hmm;

struct Inputs
{
    std::string truth, target;
    std::vector<std::uint8_t> quals;
    PenaltyVector gap_open, gap_extend, snv_priors;
    NucleotideVector snv_mask;
    Inputs(const char* truth_, int truth_len, const char* target_, int target_len, const std::uint8_t* quals_,
           const std::int8_t* go, const std::int8_t* ge, const char* mask, const std::int8_t* prior)
    : truth(truth_, truth_ + truth_len), target(target_, target_ + target_len), quals(quals_, quals_ + target_len),
      gap_open(go, go + truth_len), gap_extend(ge, ge + truth_len), snv_priors(prior, prior + truth_len), snv_mask(mask, mask + truth_len) {}
};

simd::PairHMMWrapper make_hmm(int min_band, int use_int32)
{
    return simd::PairHMMWrapper {min_band, use_int32 ? simd::PairHMMWrapper::ScorePrecision::int32 : simd::PairHMMWrapper::ScorePrecision::int16};
}

} // namespace

extern "C" {

// band the wrapper picks for a request (simd_pair_hmm_wrapper.hpp:209-241), or -1 on TooLargeBandSizeError
int ref_hmm_band(int min_band, int use_int32)
{
    try { return make_hmm(min_band, use_int32).band_size(); }
    catch (const simd::PairHMMWrapper::TooLargeBandSizeError&) { return -1; }
}

// hmm::evaluate(truth, target, qualities, target_offset, hmm, MutationModel{...}) — pair_hmm.hpp:827-841
double ref_hmm_evaluate(int min_band, int use_int32, const char* truth, int truth_len, const char* target, int target_len,
                        const std::uint8_t* quals, long long target_offset,
                        const std::int8_t* gap_open, const std::int8_t* gap_extend, const char* snv_mask, const std::int8_t* snv_prior,
                        long long lhs_flank, long long rhs_flank, int nuc_prior)
{
    const Inputs in {truth, truth_len, target, target_len, quals, gap_open, gap_extend, snv_mask, snv_prior};
    const MutationModel params {in.gap_open, in.gap_extend, in.snv_mask, in.snv_priors, {}, static_cast<std::size_t>(lhs_flank),
                                static_cast<std::size_t>(rhs_flank), static_cast<short>(nuc_prior)};
    const auto hmm = make_hmm(min_band, use_int32);
    return evaluate(in.truth, in.target, in.quals, static_cast<std::size_t>(target_offset), hmm, params);
}

// hmm::align(truth, target, qualities, target_offset, hmm, MutationModel{...}, result) — pair_hmm.hpp:858-874.
// cigar: SAM text ("37=1X12=2I98="). Returns 0, or 1 if the text does not fit cigar_cap.
int ref_hmm_align(int min_band, int use_int32, const char* truth, int truth_len, const char* target, int target_len,
                  const std::uint8_t* quals, long long target_offset,
                  const std::int8_t* gap_open, const std::int8_t* gap_extend, const char* snv_mask, const std::int8_t* snv_prior,
                  long long lhs_flank, long long rhs_flank, int nuc_prior,
                  long long* out_target_offset, double* out_likelihood, char* cigar, int cigar_cap)
{
    const Inputs in {truth, truth_len, target, target_len, quals, gap_open, gap_extend, snv_mask, snv_prior};
    const MutationModel params {in.gap_open, in.gap_extend, in.snv_mask, in.snv_priors, {}, static_cast<std::size_t>(lhs_flank),
                                static_cast<std::size_t>(rhs_flank), static_cast<short>(nuc_prior)};
    const auto hmm = make_hmm(min_band, use_int32);
    Alignment result {};
    align(in.truth, in.target, in.quals, static_cast<std::size_t>(target_offset), hmm, params, result);
    *out_target_offset = static_cast<long long>(result.target_offset);
    *out_likelihood = result.likelihood;
    std::string text;
    for (const auto& op : result.cigar) { text += std::to_string(op.size()); text += static_cast<char>(op.flag()); }
    if (static_cast<int>(text.size()) + 1 > cigar_cap) return 1;
    std::memcpy(cigar, text.c_str(), text.size() + 1);
    return 0;
}

// The candidate-position mapping HaplotypeLikelihoodArray::populate makes per (read, haplotype)
// (haplotype_likelihood_array.cpp:76-93): hashes of the read, hash table + vote counts of the haplotype,
// map_query_to_target(..., maxMappingPositions). Returns the number of positions written.
// DeNovoModel's call (core/models/mutation/denovo_model.cpp:249-262): hmm::PairHMM<VariableGapExtendMutationModel, 32, int> — no SNV mask, a scalar
// mismatch penalty in place of base qualities, gap_open / gap_extend arrays — align(target, padded given) at offset = band (pair_hmm.hpp:876-890).
// Returns 0, 1 if the CIGAR does not fit, 2 on HMMOverflow.
int ref_hmm_align_mutation_model(const char* truth, int truth_len, const char* target, int target_len, int mismatch,
                                 const std::int8_t* gap_open, const std::int8_t* gap_extend,
                                 long long* out_target_offset, double* out_likelihood, char* cigar, int cigar_cap)
{
    const std::string tr(truth, truth + truth_len), tg(target, target + target_len);
    const PenaltyVector go(gap_open, gap_open + truth_len), ge(gap_extend, gap_extend + truth_len);
    const VariableGapExtendMutationModel params {go, ge, {}, {}, static_cast<Penalty>(mismatch)};
    octopus::hmm::PairHMM<VariableGapExtendMutationModel, 32, int> hmm {};
    hmm.set(params);
    Alignment result {};
    try { hmm.align(tg, tr, result); } catch (const HMMOverflow&) { return 2; }
    *out_target_offset = static_cast<long long>(result.target_offset);
    *out_likelihood = result.likelihood;
    std::string text;
    for (const auto& op : result.cigar) { text += std::to_string(op.size()); text += static_cast<char>(op.flag()); }
    if (static_cast<int>(text.size()) + 1 > cigar_cap) return 1;
    std::memcpy(cigar, text.c_str(), text.size() + 1);
    return 0;
}

int ref_kmer_map(const char* query, int query_len, const char* target, int target_len, int max_positions, long long* out_positions)
{
    constexpr unsigned char K = 6;                        // HaplotypeLikelihoodArray::mapperKmerSize (haplotype_likelihood_array.hpp:103)
    const std::string q(query, query + query_len), t(target, target + target_len);
    if (q.size() < K || t.size() < K) return 0;           // the reference never maps sequences shorter than a k-mer
    const auto read_hashes = octopus::compute_kmer_hashes<K>(q);
    const auto haplotype_hashes = octopus::make_kmer_hash_table<K>(t);
    auto counts = octopus::init_mapping_counts(haplotype_hashes);
    std::vector<std::size_t> positions(static_cast<std::size_t>(max_positions));
    const auto last = octopus::map_query_to_target(read_hashes, haplotype_hashes, counts, positions.begin(), static_cast<std::size_t>(max_positions));
    const int n = static_cast<int>(last - positions.begin());
    for (int i = 0; i < n; ++i) out_positions[i] = static_cast<long long>(positions[i]);
    return n;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// HaplotypeLikelihoodModel (src/core/models/haplotype_likelihood_model.cpp, compiled UNMODIFIED): reset(haplotype, flank_state)
// then evaluate(read, positions) (:187-304: in-range rule, max over mapping positions U original position, shifted fallback,
// ShortHaplotypeError, mapping-quality mixing, clamp) or align(read, positions) (:335-431). Haplotype, AlignedRead and the error
// models are the stand-ins of oracle/ref_shim/ (the path takes the error models' arrays as input).
// ---------------------------------------------------------------------------------------------------------------------
namespace {

struct ModelCall
{
    octopus::HaplotypeLikelihoodModel model;
    octopus::Haplotype haplotype;
    octopus::AlignedRead read;
    octopus::HaplotypeLikelihoodModel::MappingPositionVector positions;
    boost::optional<octopus::HaplotypeLikelihoodModel::FlankState> flank_state;
};

octopus::HaplotypeLikelihoodModel::Config make_config(int max_indel_error, int use_int_scores, int use_mapping_quality, int mapq_cap, int mapq_cap_trigger)
{
    octopus::HaplotypeLikelihoodModel::Config config {};
    config.use_mapping_quality = use_mapping_quality != 0;
    if (mapq_cap_trigger >= 0) config.mapping_quality_cap_trigger = static_cast<octopus::AlignedRead::MappingQuality>(mapq_cap_trigger);
    config.mapping_quality_cap = static_cast<octopus::AlignedRead::MappingQuality>(mapq_cap);
    config.max_indel_error = static_cast<unsigned>(max_indel_error);
    config.use_int_scores = use_int_scores != 0;
    return config;
}

struct ModelArgs
{
    int max_indel_error, use_int_scores, use_mapping_quality, mapq_cap, mapq_cap_trigger;
    const char* hap; int hap_len; long long hap_begin;
    const char* mask_f; const std::int8_t* prior_f; const char* mask_r; const std::int8_t* prior_r;
    const std::int8_t* gap_open; const std::int8_t* gap_extend;
    int has_flank; long long lhs_flank, rhs_flank;
    const char* read; const std::uint8_t* quals; int read_len, mapping_quality, reverse; long long read_begin;
    const long long* positions; int n_positions, map_positions;
};

ModelCall prepare(const ModelArgs& a)
{
    auto& p = default_penalties;
    p.forward_mask.assign(a.mask_f, a.mask_f + a.hap_len); p.reverse_mask.assign(a.mask_r, a.mask_r + a.hap_len);
    p.forward_priors.assign(a.prior_f, a.prior_f + a.hap_len); p.reverse_priors.assign(a.prior_r, a.prior_r + a.hap_len);
    p.gap_open.assign(a.gap_open, a.gap_open + a.hap_len); p.gap_extend.assign(a.gap_extend, a.gap_extend + a.hap_len);
    ModelCall c {octopus::HaplotypeLikelihoodModel {make_config(a.max_indel_error, a.use_int_scores, a.use_mapping_quality, a.mapq_cap, a.mapq_cap_trigger)},
                 octopus::Haplotype {std::string(a.hap, a.hap + a.hap_len), static_cast<octopus::ContigRegion::Position>(a.hap_begin)},
                 octopus::AlignedRead {std::string(a.read, a.read + a.read_len), std::vector<std::uint8_t>(a.quals, a.quals + a.read_len),
                                       static_cast<std::uint8_t>(a.mapping_quality), a.reverse != 0, static_cast<octopus::ContigRegion::Position>(a.read_begin)},
                 {}, {}};
    if (a.map_positions) {
        // as HaplotypeLikelihoodArray::populate does inline (haplotype_likelihood_array.cpp:76-93)
        constexpr unsigned char K = 6;
        if (c.read.sequence().size() >= K && c.haplotype.sequence().size() >= K) {
            const auto read_hashes = octopus::compute_kmer_hashes<K>(c.read.sequence());
            const auto haplotype_hashes = octopus::make_kmer_hash_table<K>(c.haplotype.sequence());
            auto counts = octopus::init_mapping_counts(haplotype_hashes);
            c.positions.resize(10);
            c.positions.erase(octopus::map_query_to_target(read_hashes, haplotype_hashes, counts, c.positions.begin(), 10), c.positions.end());
        }
    } else {
        for (int i = 0; i < a.n_positions; ++i) c.positions.push_back(static_cast<std::size_t>(a.positions[i]));
    }
    if (a.has_flank) c.flank_state = octopus::HaplotypeLikelihoodModel::FlankState {static_cast<octopus::ContigRegion::Position>(a.lhs_flank),
                                                                                   static_cast<octopus::ContigRegion::Position>(a.rhs_flank)};
    return c;
}

} // namespace

extern "C" {

// returns 0, or 1 on ShortHaplotypeError (*required_extension set)
int ref_model_evaluate(const ModelArgs* args, double* out, int* required_extension)
{
    auto c = prepare(*args);
    c.model.reset(c.haplotype, c.flank_state);
    try {
        *out = c.model.evaluate(c.read, c.positions);
    } catch (const octopus::HaplotypeLikelihoodModel::ShortHaplotypeError& e) {
        *required_extension = static_cast<int>(e.required_extension());
        return 1;
    }
    return 0;
}

// returns 0, 1 on ShortHaplotypeError, 2 if the CIGAR text does not fit
int ref_model_align(const ModelArgs* args, long long* mapping_position, double* likelihood, char* cigar, int cigar_cap, int* required_extension)
{
    auto c = prepare(*args);
    c.model.reset(c.haplotype, c.flank_state);
    try {
        const auto a = c.model.align(c.read, c.positions);
        *mapping_position = static_cast<long long>(a.mapping_position);
        *likelihood = a.likelihood;
        std::string text;
        for (const auto& op : a.cigar) { text += std::to_string(op.size()); text += static_cast<char>(op.flag()); }
        if (static_cast<int>(text.size()) + 1 > cigar_cap) return 2;
        std::memcpy(cigar, text.c_str(), text.size() + 1);
    } catch (const octopus::HaplotypeLikelihoodModel::ShortHaplotypeError& e) {
        *required_extension = static_cast<int>(e.required_extension());
        return 1;
    }
    return 0;
}

} // extern "C"

// ---------------------------------------------------------------------------------------------------------------------
// HaplotypeLikelihoodArray::populate (src/core/models/haplotype_likelihood_array.cpp, compiled UNMODIFIED): the H x S x R loop with
// the inline k-mer mapping (:51-103), and the TemplateMap overload (:105-199). Samples are named s000, s001, ... so that the
// stand-in ReadMap (an ordered map) iterates them in index order. out is [H][all samples' reads (or templates) back to back].
// ---------------------------------------------------------------------------------------------------------------------
extern "C" {

struct ArrayArgs
{
    int max_indel_error, use_int_scores, use_mapping_quality, mapq_cap, mapq_cap_trigger;
    int H; const long long* hap_off; const char* seq; const char* mask_f; const std::int8_t* prior_f; const char* mask_r;
    const std::int8_t* prior_r; const std::int8_t* gap_open; const std::int8_t* gap_extend; const long long* hap_begin;
    int S; const long long* sample_off;      // reads (or templates) of sample s: [sample_off[s], sample_off[s+1])
    int R; const long long* read_off; const char* bases; const std::uint8_t* quals; const std::uint8_t* mapq;
    const std::uint8_t* reverse; const long long* read_begin;
    int T; const long long* template_off;    // T == 0: ReadMap overload; else template t owns reads [template_off[t], template_off[t+1])
    int has_flank; long long lhs_flank, rhs_flank;
};

// returns 0, or 1 on ShortHaplotypeError (*required_extension set)
int ref_array_populate(const ArrayArgs* a, double* out, int* required_extension)
{
    using namespace octopus;
    MappableBlock<Haplotype> haplotypes;
    haplotypes.reserve(static_cast<std::size_t>(a->H));
    for (int h = 0; h < a->H; ++h) {
        haplotypes.emplace_back(std::string(a->seq + a->hap_off[h], a->seq + a->hap_off[h + 1]), static_cast<ContigRegion::Position>(a->hap_begin[h]));
    }
    block_penalties.clear();
    for (int h = 0; h < a->H; ++h) {
        const auto b = a->hap_off[h], e = a->hap_off[h + 1];
        auto& p = block_penalties[&haplotypes[static_cast<std::size_t>(h)]];
        p.forward_mask.assign(a->mask_f + b, a->mask_f + e); p.reverse_mask.assign(a->mask_r + b, a->mask_r + e);
        p.forward_priors.assign(a->prior_f + b, a->prior_f + e); p.reverse_priors.assign(a->prior_r + b, a->prior_r + e);
        p.gap_open.assign(a->gap_open + b, a->gap_open + e); p.gap_extend.assign(a->gap_extend + b, a->gap_extend + e);
    }
    const auto make_read = [a] (long long r) {
        return AlignedRead {std::string(a->bases + a->read_off[r], a->bases + a->read_off[r + 1]),
                            std::vector<std::uint8_t>(a->quals + a->read_off[r], a->quals + a->read_off[r + 1]),
                            a->mapq[r], a->reverse[r] != 0, static_cast<ContigRegion::Position>(a->read_begin[r])};
    };
    std::vector<SampleName> samples;
    for (int s = 0; s < a->S; ++s) { char name[16]; std::snprintf(name, sizeof name, "s%03d", s); samples.emplace_back(name); }
    boost::optional<HaplotypeLikelihoodArray::FlankState> flank_state {};
    if (a->has_flank) flank_state = HaplotypeLikelihoodArray::FlankState {static_cast<ContigRegion::Position>(a->lhs_flank), static_cast<ContigRegion::Position>(a->rhs_flank)};
    HaplotypeLikelihoodArray array {HaplotypeLikelihoodModel {make_config(a->max_indel_error, a->use_int_scores, a->use_mapping_quality, a->mapq_cap, a->mapq_cap_trigger)},
                                    static_cast<unsigned>(a->H), samples};
    const long long width = a->sample_off[a->S];
    try {
        if (a->T == 0) {
            ReadMap reads;
            for (int s = 0; s < a->S; ++s) {
                auto& v = reads[samples[static_cast<std::size_t>(s)]];
                for (long long r = a->sample_off[s]; r < a->sample_off[s + 1]; ++r) v.push_back(make_read(r));
            }
            array.populate(reads, haplotypes, flank_state);
        } else {
            TemplateMap templates;
            for (int s = 0; s < a->S; ++s) {
                auto& v = templates[samples[static_cast<std::size_t>(s)]];
                for (long long t = a->sample_off[s]; t < a->sample_off[s + 1]; ++t) {
                    std::vector<AlignedRead> members;
                    for (long long r = a->template_off[t]; r < a->template_off[t + 1]; ++r) members.push_back(make_read(r));
                    v.emplace_back(std::move(members));
                }
            }
            array.populate(templates, haplotypes, flank_state);
        }
    } catch (const HaplotypeLikelihoodModel::ShortHaplotypeError& e) {
        block_penalties.clear();
        *required_extension = static_cast<int>(e.required_extension());
        return 1;
    }
    for (int h = 0; h < a->H; ++h) {
        for (int s = 0; s < a->S; ++s) {
            // by index (the accessor the genotype models use): haplotypes with equal sequences share one key in haplotype_indices_
            const auto& v = array(samples[static_cast<std::size_t>(s)], IndexedHaplotype<> {haplotypes[static_cast<std::size_t>(h)], static_cast<std::size_t>(h)});
            std::copy(v.begin(), v.end(), out + h * width + a->sample_off[s]);
        }
    }
    block_penalties.clear();
    return 0;
}

} // extern "C"
