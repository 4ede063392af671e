// TEST INFRASTRUCTURE ONLY — stand-in for <boost/optional.hpp> (Boost is not in this image) for the UNMODIFIED reference files
// src/core/models/haplotype_likelihood_{model,array}.{hpp,cpp} and utils/parallel_transform.hpp, which use optional<T>, none,
// operator bool, * and -> — and optional<T&> (an optional thread pool), which std::optional does not offer.
#ifndef REF_SHIM_BOOST_OPTIONAL_HPP
#define REF_SHIM_BOOST_OPTIONAL_HPP
#include <optional>
#include <type_traits>
#include <utility>
namespace boost {
using none_t = std::nullopt_t;
inline constexpr none_t none = std::nullopt;
template <typename T>
class optional : public std::optional<T>
{
public:
    using std::optional<T>::optional;
    optional() = default;
    optional(const optional&) = default;
    optional(optional&&) = default;
    optional& operator=(const optional&) = default;
    optional& operator=(optional&&) = default;
    optional& operator=(none_t) noexcept { this->reset(); return *this; }
    template <typename U, typename = std::enable_if_t<!std::is_same<std::decay_t<U>, optional>::value && !std::is_same<std::decay_t<U>, none_t>::value>>
    optional& operator=(U&& value) { std::optional<T>::operator=(std::forward<U>(value)); return *this; }
};
template <typename T>
class optional<T&>
{
public:
    optional() = default;
    optional(none_t) noexcept {}
    optional(T& value) noexcept : ptr_ {&value} {}
    explicit operator bool() const noexcept { return ptr_ != nullptr; }
    T& operator*() const noexcept { return *ptr_; }
    T* operator->() const noexcept { return ptr_; }
private:
    T* ptr_ = nullptr;
};
} // namespace boost
#endif
