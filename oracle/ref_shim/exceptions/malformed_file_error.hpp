// TEST INFRASTRUCTURE ONLY — stand-in for exceptions/malformed_file_error.hpp (base of the factory's MalformedErrorModelFile).
#ifndef REF_SHIM_MALFORMED_FILE_ERROR_HPP
#define REF_SHIM_MALFORMED_FILE_ERROR_HPP
#include <string>
#include <utility>
#include <boost/filesystem/path.hpp>
#include "error.hpp"
namespace octopus {
class MalformedFileError : public UserError
{
public:
    MalformedFileError(boost::filesystem::path file, std::string type) : file_ {std::move(file)}, type_ {std::move(type)} {}
    virtual ~MalformedFileError() = default;
private:
    std::string do_why() const override { return "the " + type_ + " file " + file_.string() + " is malformed"; }
    boost::filesystem::path file_;
    std::string type_;
};
} // namespace octopus
#endif
