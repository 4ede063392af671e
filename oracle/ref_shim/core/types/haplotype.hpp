// TEST INFRASTRUCTURE ONLY — Haplotype as the path sees it (core/types/haplotype.hpp:54, 247): a nucleotide sequence with a
// mapped region. The real class is built from alleles over a ReferenceGenome (HTSlib).
#ifndef REF_SHIM_HAPLOTYPE_HPP
#define REF_SHIM_HAPLOTYPE_HPP
#include <string>
#include <utility>
#include "basics/contig_region.hpp"
namespace octopus {
class Haplotype
{
public:
    using NucleotideSequence = std::string;
    Haplotype(NucleotideSequence sequence, ContigRegion::Position begin)
    : sequence_ {std::move(sequence)}, region_ {begin, static_cast<ContigRegion::Position>(begin + sequence_.size())} {}
    const NucleotideSequence& sequence() const noexcept { return sequence_; }
    const ContigRegion& mapped_region() const noexcept { return region_; }
private:
    NucleotideSequence sequence_;
    ContigRegion region_;
};
inline Haplotype::NucleotideSequence::size_type sequence_size(const Haplotype& haplotype) noexcept { return haplotype.sequence().size(); }
} // namespace octopus
#endif
