// bgx_internal.h — declarations shared by the translation units of libbgx (not installed)
#pragma once
#include <cstdarg>
#include <cstdio>

namespace bgx {
// thread-local message behind bgx_last_error()
void set_error(const char *fmt, ...);
} // namespace bgx
