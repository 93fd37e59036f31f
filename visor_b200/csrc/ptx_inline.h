// ptx_inline.h — textual inlining of PTX device functions at their call sites (ptx_inline.cpp).
#pragma once
#include <string>

namespace vb200
{
// Replaces every call of `name` inside `text` by the body of its definition found in `defs` (registers and
// labels renamed per call site, parameter loads and return-value stores turned into moves). `*serial` numbers
// the instances. Returns how many call sites were rewritten; a call whose shape is not understood, or a callee
// that needs a stack frame, is left alone.
int ptx_inline_calls(std::string &text, const std::string &defs, const std::string &name, int *serial);
bool ptx_has_call(const std::string &text, const std::string &name);
}    // namespace vb200
