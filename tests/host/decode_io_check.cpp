// Host-only check of jpezy::decode_io's P3 writer (include/jpezy/decode_io.hpp): small images take the single-thread path,
// large ones are formatted in ranges on the host's cores; both must give the reference's text (src/decoder/decode_io.hpp:37-54).
#include <fstream>
#include <iostream>
#include <sstream>
#include <vector>
#include <random>
#include "jpezy/decode_io.hpp"
int main() {
    for (std::size_t W : {3u, 700u}) for (std::size_t H : {2u, 600u}) {
        const std::size_t plane = (W + 16) * (H + 16);
        std::vector<jpezy::byte> r(plane), g(plane), b(plane);
        std::mt19937 rng(W * 31 + H);
        for (std::size_t i = 0; i < plane; ++i) r[i] = jpezy::byte(rng()), g[i] = jpezy::byte(rng()), b[i] = jpezy::byte(rng());
        std::ostringstream got, want;
        got << jpezy::decode_io(W, H, r, g, b);
        want << "P3\n# Decoded by jpezy\n" << W << " " << H << "\n255\n";
        for (std::size_t i = 0; i < W * H; ++i) want << unsigned(r[i]) << " " << unsigned(g[i]) << " " << unsigned(b[i]) << "\n";
        const bool ok = got.str() == want.str();
        std::cout << W << "x" << H << (ok ? " ok " : " MISMATCH ") << got.str().size() << "\n";
        if (!ok) return 1;
    }
}
