// TEST INFRASTRUCTURE ONLY — the two free functions of concepts/mappable.hpp the path calls on (haplotype, read):
// begin_distance (:1017-1022 → contig_region.hpp:346-349) and contains (inside an assert only).
#ifndef REF_SHIM_MAPPABLE_HPP
#define REF_SHIM_MAPPABLE_HPP
#include "basics/contig_region.hpp"
namespace octopus {
template <typename T1, typename T2>
auto begin_distance(const T1& first, const T2& second) noexcept { return begin_distance(first.mapped_region(), second.mapped_region()); }
template <typename T1, typename T2>
bool contains(const T1& lhs, const T2& rhs) noexcept
{
    return lhs.mapped_region().begin() <= rhs.mapped_region().begin() && rhs.mapped_region().end() <= lhs.mapped_region().end();
}
} // namespace octopus
#endif
