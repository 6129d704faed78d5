// stand-in (oracle/shim/README.md): the Boost.Range algorithms the reference calls, over std:: algorithms
#pragma once
#include <algorithm>
#include <iterator>
namespace boost {
template <class R, class Out>
Out copy(const R& r, Out out) { return std::copy(std::begin(r), std::end(r), out); }
template <class R, class V>
R& fill(R& r, const V& v) { std::fill(std::begin(r), std::end(r), v); return r; }
template <class R, class G>
R& generate(R& r, G g) { std::generate(std::begin(r), std::end(r), g); return r; }
template <class R, class Out, class F>
Out transform(const R& r, Out out, F f) { return std::transform(std::begin(r), std::end(r), out, f); }
}  // namespace boost
