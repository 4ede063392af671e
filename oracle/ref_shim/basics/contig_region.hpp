// TEST INFRASTRUCTURE ONLY — the types of basics/contig_region.hpp (:28-30) that the path names: a half-open [begin, end).
#ifndef REF_SHIM_CONTIG_REGION_HPP
#define REF_SHIM_CONTIG_REGION_HPP
#include <cstdint>
namespace octopus {
class ContigRegion
{
public:
    using Position = std::uint_fast32_t;
    using Size     = Position;
    using Distance = std::int_fast64_t;
    ContigRegion() = default;
    ContigRegion(Position begin, Position end) noexcept : begin_ {begin}, end_ {end} {}
    Position begin() const noexcept { return begin_; }
    Position end() const noexcept { return end_; }
private:
    Position begin_ = 0, end_ = 0;
};
// basics/contig_region.hpp:346-349
inline ContigRegion::Distance begin_distance(const ContigRegion& first, const ContigRegion& second) noexcept
{
    return static_cast<ContigRegion::Distance>(second.begin()) - first.begin();
}
} // namespace octopus
#endif
