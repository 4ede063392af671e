// TEST INFRASTRUCTURE ONLY — the AlignedRead fields the path reads (basics/aligned_read.hpp:36-39, 120-146): sequence,
// base qualities, mapping quality, strand flag, mapped region. The real class needs HTSlib-facing types and Boost.
#ifndef REF_SHIM_ALIGNED_READ_HPP
#define REF_SHIM_ALIGNED_READ_HPP
#include <cstdint>
#include <string>
#include <utility>
#include <vector>
#include "basics/contig_region.hpp"
namespace octopus {
class AlignedRead
{
public:
    using NucleotideSequence = std::string;
    using MappingQuality     = std::uint8_t;
    using BaseQuality        = std::uint8_t;
    using BaseQualityVector  = std::vector<BaseQuality>;
    AlignedRead(NucleotideSequence sequence, BaseQualityVector qualities, MappingQuality mapping_quality, bool reverse, ContigRegion::Position begin)
    : sequence_ {std::move(sequence)}, qualities_ {std::move(qualities)}, mapping_quality_ {mapping_quality}, reverse_ {reverse},
      region_ {begin, static_cast<ContigRegion::Position>(begin + sequence_.size())} {}
    const NucleotideSequence& sequence() const noexcept { return sequence_; }
    const BaseQualityVector& base_qualities() const noexcept { return qualities_; }
    MappingQuality mapping_quality() const noexcept { return mapping_quality_; }
    bool is_marked_reverse_mapped() const noexcept { return reverse_; }
    const ContigRegion& mapped_region() const noexcept { return region_; }
    // named by the array header's debug printers only
    const std::string& name() const noexcept { return name_; }
    const std::string& cigar() const noexcept { return name_; }
private:
    std::string name_;
    NucleotideSequence sequence_;
    BaseQualityVector qualities_;
    MappingQuality mapping_quality_;
    bool reverse_;
    ContigRegion region_;
};
inline AlignedRead::NucleotideSequence::size_type sequence_size(const AlignedRead& read) noexcept { return read.sequence().size(); }
template <typename T> const ContigRegion& mapped_region(const T& mappable) noexcept { return mappable.mapped_region(); }
} // namespace octopus
#endif
