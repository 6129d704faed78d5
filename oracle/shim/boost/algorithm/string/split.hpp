// stand-in (oracle/shim/README.md): boost::split with token_compress_off (the default): every separator ends a token,
// so consecutive separators and a trailing separator produce empty tokens; an empty input produces one empty token.
#pragma once
#include <string>
namespace boost {
template <class Container, class Pred>
Container& split(Container& out, const std::string& in, Pred is_sep)
{
    out.clear();
    std::string cur;
    for (char c : in) {
        if (is_sep(c)) {
            out.push_back(cur);
            cur.clear();
        } else {
            cur.push_back(c);
        }
    }
    out.push_back(cur);
    return out;
}
}  // namespace boost
