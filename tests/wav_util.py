"""Builders of WAV byte blobs for the io tests (format per the RIFF specification; no reference code involved)."""
import struct

import numpy as np


def chunk(cid, payload, e="<"):
    return cid + struct.pack(e + "I", len(payload)) + payload + (b"\x00" if len(payload) & 1 else b"")


def make_wav(samples, rate=16000, width=2, channels=1, tag=1, big_endian=False, extra_before=(), extra_after=(),
             extensible=False, depth=None, riff_size=None, truncate=None, align=None):
    """samples: integer / float array of raw sample values (already in container units), interleaved for channels > 1."""
    e = ">" if big_endian else "<"
    depth = depth if depth is not None else 8 * width
    a = np.asarray(samples)
    if tag == 3:
        payload = a.astype(e + "f%d" % width).tobytes()
    elif width in (1, 2, 4, 8):
        payload = a.astype(("u1" if width == 1 else e + "i%d" % width)).tobytes()
    else:   # 3, 5, 6, 7 byte containers: the low `width` bytes of the two's-complement value
        v = a.astype(np.int64)
        cols = [((v >> (8 * k)) & 0xFF).astype(np.uint8) for k in range(width)]
        if big_endian:
            cols = cols[::-1]
        payload = np.stack(cols, axis=1).tobytes()
    align = align if align is not None else width * channels
    fmt_tag = 0xFFFE if extensible else tag
    fmt = struct.pack(e + "HHIIHH", fmt_tag, channels, rate, rate * align, align, depth)
    if extensible:
        guid_tail = (b"\x00\x00\x00\x10" if big_endian else b"\x00\x00\x10\x00") + b"\x80\x00\x00\xAA\x00\x38\x9B\x71"
        fmt += struct.pack(e + "H", 22) + struct.pack(e + "H", depth) + struct.pack(e + "I", 0) + struct.pack(e + "I", tag) + guid_tail
    body = b"WAVE" + chunk(b"fmt ", fmt, e)
    for cid, pl in extra_before:
        body += chunk(cid, pl, e)
    body += chunk(b"data", payload, e)
    for cid, pl in extra_after:
        body += chunk(cid, pl, e)
    size = len(body) if riff_size is None else riff_size
    blob = (b"RIFX" if big_endian else b"RIFF") + struct.pack(e + "I", size) + body
    return blob[:truncate] if truncate is not None else blob


def corpus(seed=0):
    """name -> (blob, offset, duration, filelike): the cases the io tests share."""
    r = np.random.default_rng(seed)
    i16 = r.integers(-32768, 32768, 1600)
    i24 = r.integers(-(1 << 23), 1 << 23, 999)
    i32 = r.integers(-(1 << 31), 1 << 31, 800)
    f = r.standard_normal(640)
    c = {
        "pcm16": (make_wav(i16), 0.0, None, False),
        "pcm16_filelike": (make_wav(i16), 0.0, None, True),
        "pcm16_extremes": (make_wav(np.array([-32768, 32767, 0, -1, 1])), 0.0, None, False),
        "pcm16_stereo": (make_wav(i16, channels=2), 0.0, None, False),
        "pcm16_be": (make_wav(i16, big_endian=True), 0.0, None, False),
        "pcm16_odd_payload": (make_wav(i16[:333], width=2, extra_after=[(b"LIST", b"abc")]), 0.0, None, False),
        "pcm16_chunks": (make_wav(i16, extra_before=[(b"LIST", b"INFOx"), (b"cue ", b"1234567"), (b"fact", b"\x00" * 4)],
                                  extra_after=[(b"JUNK", b"zz"), (b"smpl", b"q" * 9)]), 0.0, None, False),
        "pcm16_offset": (make_wav(i16), 0.01, None, False),
        "pcm16_duration": (make_wav(i16), 0.0, 0.05, False),
        "pcm16_offset_duration": (make_wav(i16), 0.02, 0.03, False),
        "pcm16_duration_filelike": (make_wav(i16), 0.0, 0.05, True),
        "pcm16_truncated": (make_wav(i16, truncate=44 + 1000), 0.0, None, False),
        "pcm16_riff_size_small": (make_wav(i16, extra_after=[(b"LIST", b"abcd")], riff_size=36 + 3200), 0.0, None, False),
        "pcm16_extensible": (make_wav(i16, extensible=True), 0.0, None, False),
        "pcm16_empty": (make_wav(np.zeros(0)), 0.0, None, False),
        "pcm8": (make_wav(r.integers(0, 256, 500), width=1), 0.0, None, False),
        "pcm24": (make_wav(i24, width=3), 0.0, None, False),
        "pcm24_be": (make_wav(i24, width=3, big_endian=True), 0.0, None, False),
        "pcm24_stereo": (make_wav(i24[:998], width=3, channels=2), 0.0, None, False),
        "pcm32": (make_wav(i32, width=4), 0.0, None, False),
        "pcm32_be": (make_wav(i32, width=4, big_endian=True), 0.0, None, False),
        "pcm40": (make_wav(r.integers(-(1 << 39), 1 << 39, 77), width=5), 0.0, None, False),
        "pcm64": (make_wav(r.integers(-(1 << 62), 1 << 62, 64), width=8), 0.0, None, False),
        "float32": (make_wav(f, width=4, tag=3), 0.0, None, False),
        "float32_be": (make_wav(f, width=4, tag=3, big_endian=True), 0.0, None, False),
        "float64": (make_wav(f, width=8, tag=3), 0.0, None, False),
        "float32_ext": (make_wav(f, width=4, tag=3, extensible=True), 0.0, None, False),
    }
    return c


def bad_corpus():
    """name -> (blob, exception class)."""
    good = make_wav(np.arange(10))
    return {
        "not_riff": (b"FORM" + good[4:], ValueError),
        "not_wave": (good[:8] + b"AVI " + good[12:], TypeError),
        "no_data": (good[:36], ValueError),                   # file ends after fmt: "Unexpected end of file."
        "no_data_short_size": (good[:4] + b"\x1c\x00\x00\x00" + good[8:36], UnboundLocalError),
        "short_fmt": (good[:16] + b"\x0e\x00\x00\x00" + good[20:], ValueError),
        "mulaw": (make_wav(np.arange(10), tag=7), ValueError),
        "bad_rate": (good[:28] + b"\x01\x00\x00\x00" + good[32:], ValueError),
        "data_before_fmt": (good[:12] + good[36:] + good[12:36], ValueError),
        "float16": (make_wav(np.arange(10), width=2, tag=3, depth=16), ValueError),
        "pcm_depth_72": (make_wav(np.arange(8), width=8, depth=72), ValueError),
        "incomplete_id": (good[:36] + b"da", ValueError),
    }
