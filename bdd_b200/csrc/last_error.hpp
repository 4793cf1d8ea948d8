// bdd_b200/csrc/last_error.hpp -- the per-thread message behind bddb200_last_error(), shared by the translation units of libbdd_b200.so.
#pragma once
#include <string>

namespace bddb200 { namespace detail {
__attribute__((visibility("hidden"))) void set_last_error(const std::string& message);      // defined in bdd_b200.cu
} }
