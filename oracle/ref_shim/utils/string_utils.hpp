// TEST INFRASTRUCTURE ONLY — the one function of utils/string_utils.hpp the error-model factory uses (capitalise, for the
// library / sequencer names); the real header pulls in utils/maths.hpp → Boost.
#ifndef REF_SHIM_STRING_UTILS_HPP
#define REF_SHIM_STRING_UTILS_HPP
#include <algorithm>
#include <cctype>
#include <string>
namespace octopus { namespace utils {
inline std::string& capitalise(std::string& str) noexcept
{
    std::transform(str.begin(), str.end(), str.begin(), [] (const unsigned char c) { return static_cast<char>(std::toupper(c)); });
    return str;
}
} } // namespace octopus::utils
#endif
