// TEST INFRASTRUCTURE ONLY — stand-in for boost::lexical_cast<int>(std::string) (custom error-model parser): whole-token
// conversion, bad_lexical_cast otherwise.
#ifndef REF_SHIM_BOOST_LEXICAL_CAST_HPP
#define REF_SHIM_BOOST_LEXICAL_CAST_HPP
#include <cstddef>
#include <stdexcept>
#include <string>
namespace boost {
struct bad_lexical_cast : std::runtime_error { bad_lexical_cast() : std::runtime_error {"bad lexical cast"} {} };
template <typename Target>
inline Target lexical_cast(const std::string& token)
{
    std::size_t used = 0;
    long long value = 0;
    try { value = std::stoll(token, &used); } catch (const std::exception&) { throw bad_lexical_cast {}; }
    if (token.empty() || used != token.size()) throw bad_lexical_cast {};
    return static_cast<Target>(value);
}
} // namespace boost
#endif
