"""The compiled host's built-in texture decoders (csrc/image_decode.cpp) against libjpeg-turbo / libpng as shipped in PIL.

The reference decodes textures with the `image` crate (src/obj.rs:16-24) and converts to RGB8 (src/texture.rs:57-59).  Its JPEG
back end (jpeg-decoder) is not vendored; the decoder here is pinned bit for bit to the IJG arithmetic (integer slow IDCT, triangle
chroma upsampling, 16-bit fixed-point YCbCr), i.e. to what PIL returns.  PNG decoding is lossless, so any conforming decoder agrees.
"""
import ctypes as C
import io
import os

import numpy as np
import pytest

from craytracer_b200 import _abi

Image = pytest.importorskip("PIL.Image")

REF_TEXTURES = "/root/reference/objs/staircase/textures"


def decode(data: bytes) -> np.ndarray:
    L = _abi.lib()
    w, h, p = C.c_uint32(), C.c_uint32(), C.c_void_p()
    buf = (C.c_uint8 * len(data)).from_buffer_copy(data)
    rc = L.cray_debug_decode_image(buf, len(data), C.byref(w), C.byref(h), C.byref(p))
    if rc != 0:
        raise RuntimeError((L.cray_last_error() or b"").decode())
    out = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(h.value, w.value, 3)).copy()
    L.cray_free(p)
    return out


def pil_rgb(data: bytes) -> np.ndarray:
    return np.asarray(Image.open(io.BytesIO(data)).convert("RGB"))


def test_pattern(w, h, seed):
    rng = np.random.default_rng(seed)
    ys, xs = np.mgrid[0:h, 0:w]
    img = np.stack([(xs * 255 // max(w - 1, 1)), (ys * 255 // max(h - 1, 1)), ((xs * 7 + ys * 13) % 256)], axis=-1).astype(np.int32)
    img += rng.integers(-40, 40, size=img.shape)  # high-frequency content: every AC coefficient gets used
    return np.clip(img, 0, 255).astype(np.uint8)


test_pattern.__test__ = False


@pytest.mark.parametrize("size", [(1, 1), (7, 5), (8, 8), (17, 33), (64, 48), (131, 97)])
@pytest.mark.parametrize("subsampling", [0, 1, 2])  # 4:4:4, 4:2:2, 4:2:0
@pytest.mark.parametrize("progressive", [False, True])
def test_jpeg_matches_libjpeg_bit_for_bit(size, subsampling, progressive):
    w, h = size
    buf = io.BytesIO()
    Image.fromarray(test_pattern(w, h, w * 1000 + h)).save(buf, "JPEG", quality=87, subsampling=subsampling, progressive=progressive, optimize=progressive)
    data = buf.getvalue()
    assert np.array_equal(decode(data), pil_rgb(data))


@pytest.mark.parametrize("quality", [5, 50, 100])
def test_jpeg_greyscale_restart_markers_and_extreme_quality(quality):
    buf = io.BytesIO()
    Image.fromarray(test_pattern(75, 41, quality)[..., 0], "L").save(buf, "JPEG", quality=quality)
    data = buf.getvalue()
    assert np.array_equal(decode(data), pil_rgb(data))
    buf = io.BytesIO()
    Image.fromarray(test_pattern(90, 70, quality)).save(buf, "JPEG", quality=quality, restart_marker_blocks=3)
    data = buf.getvalue()
    assert b"\xff\xdd" in data  # DRI present
    assert np.array_equal(decode(data), pil_rgb(data))


@pytest.mark.parametrize("mode", ["RGB", "RGBA", "L", "LA", "P", "1", "I;16"])
def test_png_every_colour_type(mode):
    rgb = test_pattern(53, 37, 9)
    if mode == "I;16":
        img = Image.fromarray(rgb[..., 0].astype(np.uint16) * 257 + 77)
    elif mode == "RGBA":
        img = Image.fromarray(np.concatenate([rgb, rgb[..., :1]], axis=-1), "RGBA")
    else:
        img = Image.fromarray(rgb).convert(mode)
    buf = io.BytesIO()
    img.save(buf, "PNG")
    got = decode(buf.getvalue())
    if mode == "I;16":
        want16 = np.asarray(img).astype(np.uint32)
        want = np.repeat((((want16 + 128) // 257).astype(np.uint8))[..., None], 3, axis=-1)  # image crate's u16 -> u8 rounding
    elif mode in ("RGBA", "LA"):
        want = np.asarray(img.convert("RGBA"))[..., :3]  # DynamicImage::to_rgb8 drops alpha (PIL's convert("RGB") does too)
    else:
        want = np.asarray(img.convert("RGB"))
    assert np.array_equal(got, want)


def test_png_low_bit_depths_and_stored_blocks():
    # 2- and 4-bit greyscale, written with compress_level 0 (stored deflate blocks) and 9 (dynamic Huffman)
    base = test_pattern(40, 23, 3)[..., 0]
    for bits in (2, 4):
        levels = (base >> (8 - bits)).astype(np.uint8)
        img = Image.fromarray((levels * (255 // ((1 << bits) - 1))).astype(np.uint8), "L")
        for level in (0, 9):
            buf = io.BytesIO()
            img.save(buf, "PNG", bits=bits, compress_level=level)
            want = np.asarray(Image.open(io.BytesIO(buf.getvalue())).convert("RGB"))
            assert np.array_equal(decode(buf.getvalue()), want)


def test_rejects_what_it_cannot_decode():
    with pytest.raises(RuntimeError, match="unsupported image format"):
        decode(b"GIF89a" + b"\0" * 32)
    buf = io.BytesIO()
    Image.fromarray(test_pattern(32, 32, 1)).save(buf, "JPEG")
    with pytest.raises(RuntimeError):
        decode(buf.getvalue()[:60])  # cut inside the tables
    buf = io.BytesIO()
    Image.fromarray(test_pattern(16, 16, 1)).save(buf, "PNG")
    with pytest.raises(RuntimeError):
        decode(buf.getvalue()[:50])


@pytest.mark.skipif(not os.path.isdir(REF_TEXTURES), reason="the reference tree is only present in the build container")
def test_the_reference_staircase_textures_decode_like_libjpeg():
    names = sorted(n for n in os.listdir(REF_TEXTURES) if n.lower().endswith(".jpg"))
    assert len(names) == 10
    texels = 0
    for n in names:
        data = open(os.path.join(REF_TEXTURES, n), "rb").read()
        got = decode(data)
        assert np.array_equal(got, pil_rgb(data)), n
        texels += got.shape[0] * got.shape[1]
    assert texels == 27_642_338  # 83 MB of RGB8 texels
