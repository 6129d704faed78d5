// include/jpezy/jpezy_encoder.hpp -- drop-in for jpezy::encoder<T> (src/encoder/jpezy_encoder.hpp:22-39,84,277-278).
// Same constructor, same encode<MODE_TAG>(output_file) -> bytes written, same console lines.  The MCU loop
// (src/encoder/jpezy_encoder.hpp:55-67: make_YCC, DCT, quantization, encode_huffman) runs on the B200 behind
// jpezyb200_encode; header, EOI and the file stay here.
#ifndef JPEZY_B200_JPEZY_ENCODER_HPP
#define JPEZY_B200_JPEZY_ENCODER_HPP

#include <stdexcept>
#include <type_traits>
#include <vector>

#include "jpezy.hpp"
#include "jpezy_writer.hpp"
#include "runtime.hpp"

namespace jpezy {

template <class T>
struct encoder {
    static_assert(sizeof(T) == 1, "8-bit samples");
    // copies the three planes, keeps a reference to the property (src/encoder/jpezy_encoder.hpp:24-36,265-266)
    encoder(const property& pr_, const std::vector<T>& r_, const std::vector<T>& g_, const std::vector<T>& b_) : pr(pr_), r(r_), g(g_), b(b_) {}

    template <class MODE_TAG = COLOR_MODE>
    std::size_t encode(const char* output_file)
    {
        const std::size_t W = pr.get<property::At::HSize>(), H = pr.get<property::At::VSize>();
        // src/encoder/jpezy_encoder.hpp:41-43 computes this in `int` (overflows above 715 Mpixel); size_t here (decision O5)
        std::size_t size = W * H * 3;
        if (size < 10240) size = 10240;
        jpezy_writer jpeg(size, pr, output_file);
        {
            raii_messenger mes("Write JPEG Header ...");
            jpeg.write_header();
        }
        {
            raii_messenger mes("Encoding ...");
            if (r.size() < W * H || g.size() < W * H || b.size() < W * H) throw std::runtime_error("encode: planes smaller than width*height");
            jpezy_writer::stream_type& s = jpeg.get_stream();
            const std::size_t head = s.size();
            const std::size_t room = jpeg.capacity() > head + 2 ? jpeg.capacity() - head - 2 : 0;
            s.resize(head + room);
            std::size_t nbytes = 0;
            const int rc = jpezyb200_encode(b200::runtime::ctx(), reinterpret_cast<const std::uint8_t*>(r.data()),
                                            reinterpret_cast<const std::uint8_t*>(g.data()), reinterpret_cast<const std::uint8_t*>(b.data()),
                                            static_cast<std::uint32_t>(W), static_cast<std::uint32_t>(H), std::is_same_v<MODE_TAG, GRAY_MODE> ? 1 : 0,
                                            s.data() + head, room, &nbytes, nullptr);
            if (rc != JPEZYB200_OK) {
                s.resize(head);
                throw b200::runtime::error("encode", rc);   // std::runtime_error, as the reference's encode_huffman / bofstream throw
            }
            s.resize(head + nbytes);
        }
        {
            raii_messenger mes("Write EOI ...");
            jpeg.write_eoi();
        }
        jpeg.output_file();
        return jpeg.wrote_size();
    }

    static constexpr int block_size = 8;

private:
    const property& pr;
    const std::vector<T> r, g, b;
};

template <class T>
encoder(const property&, const std::vector<T>&, const std::vector<T>&, const std::vector<T>&) -> encoder<T>;

}  // namespace jpezy
#endif
