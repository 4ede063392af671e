// TEST INFRASTRUCTURE ONLY — stand-in for <boost/filesystem/path.hpp> (Boost is not in this image): the reference's
// core/models/error/error_model_factory.{hpp,cpp} only take a path and call .string() on it.
#ifndef REF_SHIM_BOOST_FILESYSTEM_PATH_HPP
#define REF_SHIM_BOOST_FILESYSTEM_PATH_HPP
#include <filesystem>
namespace boost { namespace filesystem { using path = std::filesystem::path; } }
#endif
