// TEST INFRASTRUCTURE ONLY — minimal octopus::Error hierarchy root (the real exceptions/error.hpp pulls in config/ and Boost).
#ifndef REF_SHIM_ERROR_HPP
#define REF_SHIM_ERROR_HPP
#include <exception>
#include <string>
namespace octopus {
class Error : public std::exception
{
public:
    virtual ~Error() = default;
    std::string type() const { return do_type(); }
    std::string where() const { return do_where(); }
    std::string why() const { return do_why(); }
    std::string help() const { return do_help(); }
private:
    virtual std::string do_type() const = 0;
    virtual std::string do_where() const = 0;
    virtual std::string do_why() const = 0;
    virtual std::string do_help() const = 0;
};
} // namespace octopus
#endif
