// TEST INFRASTRUCTURE ONLY — Haplotype as the path sees it (core/types/haplotype.hpp:54, 247): a nucleotide sequence with a
// mapped region. The real class is built from alleles over a ReferenceGenome (HTSlib).
#ifndef REF_SHIM_HAPLOTYPE_HPP
#define REF_SHIM_HAPLOTYPE_HPP
#include <cstddef>
#include <functional>
#include <string>
#include <utility>
#include <vector>
#include "basics/contig_region.hpp"
#include "basics/cigar_string.hpp"
namespace octopus {
class Haplotype
{
public:
    using NucleotideSequence = std::string;
    Haplotype(NucleotideSequence sequence, ContigRegion::Position begin)
    : sequence_ {std::move(sequence)}, region_ {begin, static_cast<ContigRegion::Position>(begin + sequence_.size())} {}
    const NucleotideSequence& sequence() const noexcept { return sequence_; }
    const ContigRegion& mapped_region() const noexcept { return region_; }
    // Haplotype::cigar() (core/types/haplotype.hpp): '=' / 'X' / 'I' / 'D' against the reference. The stand-in is told which
    // bases are substitutions (all the SNV error model reads from it, repeat_based_snv_error_model.cpp:128-140).
    void set_substitutions(std::vector<bool> is_substitution) { substitutions_ = std::move(is_substitution); }
    CigarString cigar() const
    {
        CigarString result {};
        for (std::size_t i = 0; i < sequence_.size();) {
            const bool sub = i < substitutions_.size() && substitutions_[i];
            std::size_t j = i + 1;
            while (j < sequence_.size() && (j < substitutions_.size() && substitutions_[j]) == sub) ++j;
            result.emplace_back(static_cast<CigarOperation::Size>(j - i), sub ? CigarOperation::Flag::substitution : CigarOperation::Flag::sequenceMatch);
            i = j;
        }
        return result;
    }
private:
    std::vector<bool> substitutions_;
    NucleotideSequence sequence_;
    ContigRegion region_;
};
inline Haplotype::NucleotideSequence::size_type sequence_size(const Haplotype& haplotype) noexcept { return haplotype.sequence().size(); }
// identity as the array's unordered_map<Haplotype, index> needs it (the real class compares region + sequence and caches a hash)
inline bool operator==(const Haplotype& lhs, const Haplotype& rhs) noexcept
{
    return lhs.mapped_region().begin() == rhs.mapped_region().begin() && lhs.sequence() == rhs.sequence();
}
struct HaplotypeHash
{
    std::size_t operator()(const Haplotype& haplotype) const noexcept { return std::hash<std::string> {}(haplotype.sequence()) ^ haplotype.mapped_region().begin(); }
};
namespace debug {
template <typename S> void print_variant_alleles(S&& stream, const Haplotype& haplotype) { stream << haplotype.sequence(); }
} // namespace debug
} // namespace octopus
namespace std {
template <> struct hash<octopus::Haplotype>
{
    size_t operator()(const octopus::Haplotype& haplotype) const noexcept { return octopus::HaplotypeHash {}(haplotype); }
};
template <> struct hash<reference_wrapper<const octopus::Haplotype>>
{
    size_t operator()(const reference_wrapper<const octopus::Haplotype> haplotype) const noexcept { return octopus::HaplotypeHash {}(haplotype.get()); }
};
} // namespace std
#endif
