// oracle/ref_alloc.cpp -- TEST INFRASTRUCTURE, linked into everything under oracle/_ref.
// Replacement global allocation functions: zero-filled blocks with 64 zero bytes of slack behind them.
// Why: the reference's analyze_dht (src/decoder/jpezy_decoder.hpp:229-231) evaluates `ht_.sizeTP[k] == si` with k == n,
// one element past the end of the vector, before it tests k >= n; if that stray word equals the current code length it
// goes on writing codeTP[k] past the end.  Which word it sees is allocator history.  With this allocator it is always 0,
// i.e. the reference behaves as it does on a pristine heap, deterministically, and its own sources stay untouched.
#include <cstdlib>
#include <new>

void* operator new(std::size_t n)
{
    if (void* p = std::calloc(1, n + 64)) return p;
    throw std::bad_alloc();
}
void* operator new[](std::size_t n) { return operator new(n); }
void operator delete(void* p) noexcept { std::free(p); }
void operator delete[](void* p) noexcept { std::free(p); }
void operator delete(void* p, std::size_t) noexcept { std::free(p); }
void operator delete[](void* p, std::size_t) noexcept { std::free(p); }
