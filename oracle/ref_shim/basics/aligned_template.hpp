// TEST INFRASTRUCTURE ONLY — AlignedTemplate as the path sees it (basics/aligned_template.hpp:44-74): an iterable of reads.
#ifndef REF_SHIM_ALIGNED_TEMPLATE_HPP
#define REF_SHIM_ALIGNED_TEMPLATE_HPP
#include <cstddef>
#include <utility>
#include <vector>
#include "basics/aligned_read.hpp"
namespace octopus {
class AlignedTemplate
{
public:
    using const_iterator = std::vector<AlignedRead>::const_iterator;
    explicit AlignedTemplate(std::vector<AlignedRead> reads) : reads_ {std::move(reads)} {}
    std::size_t size() const noexcept { return reads_.size(); }
    const AlignedRead& operator[](std::size_t idx) const noexcept { return reads_[idx]; }
    const_iterator begin() const noexcept { return reads_.begin(); }
    const_iterator end() const noexcept { return reads_.end(); }
    const_iterator cbegin() const noexcept { return reads_.cbegin(); }
    const_iterator cend() const noexcept { return reads_.cend(); }
private:
    std::vector<AlignedRead> reads_;
};
} // namespace octopus
#endif
