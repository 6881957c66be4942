// Built-in texture decoders of the compiled host: JPEG (baseline and progressive Huffman, 8-bit) and PNG (non-interlaced).
// The reference decodes its textures with the `image` crate (src/obj.rs:16-24, `image::io::Reader::open(..).decode()`) and
// converts them to RGB8 (src/texture.rs:57-59, `to_rgb8`); its dependencies (jpeg-decoder, png) are not vendored, so the JPEG
// arithmetic here follows ITU T.81 with the integer inverse DCT, triangle chroma upsampling and fixed-point YCbCr conversion of
// the IJG implementation -- pinned bit for bit against libjpeg-turbo (PIL) in tests/test_image_decode.py; against jpeg-decoder a
// texel may differ by an LSB (unpinned).  PNG decoding is exact by definition.
#pragma once
#include <cstddef>
#include <cstdint>
#include <string>
#include <vector>

namespace cray {

// RGB8, row-major, top row first.  Return false with `err` set when the data is not a supported JPEG / PNG.
bool decode_jpeg(const uint8_t* data, size_t n, uint32_t& width, uint32_t& height, std::vector<uint8_t>& rgb, std::string& err);
bool decode_png(const uint8_t* data, size_t n, uint32_t& width, uint32_t& height, std::vector<uint8_t>& rgb, std::string& err);

// Dispatch on the file's magic bytes (not its name, like image::io::Reader::with_guessed_format is not used by the reference --
// it goes by extension -- but every file the reference ships has matching magic).
bool decode_image(const uint8_t* data, size_t n, uint32_t& width, uint32_t& height, std::vector<uint8_t>& rgb, std::string& err);

}  // namespace cray
