// TEST INFRASTRUCTURE ONLY — MappableBlock<T> as haplotype_likelihood_array.{hpp,cpp} uses it: an ordered, indexable,
// copyable sequence of mappables. The real container (containers/mappable_block.hpp) adds region bookkeeping over the
// full Mappable concept.
#ifndef REF_SHIM_MAPPABLE_BLOCK_HPP
#define REF_SHIM_MAPPABLE_BLOCK_HPP
#include <initializer_list>
#include <vector>
namespace octopus {
template <typename T>
class MappableBlock : public std::vector<T>
{
public:
    using std::vector<T>::vector;
    MappableBlock() = default;
    MappableBlock(std::initializer_list<T> values) : std::vector<T>(values) {}
};
} // namespace octopus
#endif
