// pc_errors.h -- the engine's own exception types.
//
// polychord_c_interface follows the reference's error convention for the ENGINE's failures (banner + exit(1),
// abort.F90:19-29) and is exception-transparent for everything else: an exception a caller's C++ callback throws --
// whatever its type, std::runtime_error included -- unwinds through the engine's RAII frames back to the caller
// (the reference's CPython shim relies on that: _pypolychord.cpp:219-224).  So the engine throws only these two
// types, and the boundary catches only them.
#pragma once
#include <stdexcept>
#include <string>

namespace pc {
struct ArgError : std::invalid_argument {   // a setting or argument the engine cannot run with
    explicit ArgError(const std::string& m) : std::invalid_argument(m) {}
};
struct RunError : std::runtime_error {      // a failure while running (CUDA, files, live-point generation)
    explicit RunError(const std::string& m) : std::runtime_error(m) {}
};
}  // namespace pc
