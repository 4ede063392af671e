// TEST INFRASTRUCTURE ONLY — IndexedHaplotype<> as the array's accessors see it (core/types/indexed_haplotype.hpp:25-55): a
// haplotype reference plus its position in the block.
#ifndef REF_SHIM_INDEXED_HAPLOTYPE_HPP
#define REF_SHIM_INDEXED_HAPLOTYPE_HPP
#include <cstddef>
#include "core/types/haplotype.hpp"
namespace octopus {
template <typename IndexTp = std::size_t>
class IndexedHaplotype
{
public:
    using IndexType = IndexTp;
    IndexedHaplotype(const Haplotype& haplotype, IndexType index) noexcept : haplotype_ {&haplotype}, index_ {index} {}
    const Haplotype& haplotype() const noexcept { return *haplotype_; }
    operator const Haplotype&() const noexcept { return *haplotype_; }
    IndexType index() const noexcept { return index_; }
private:
    const Haplotype* haplotype_;
    IndexType index_;
};
template <typename IndexTp>
IndexTp index_of(const IndexedHaplotype<IndexTp>& haplotype) noexcept { return haplotype.index(); }
} // namespace octopus
#endif
