"""Byte framing shared with the reference (cbench/utils/bytes_ops.py:19-69): native-endian u32 length
prefixes, last segment unprefixed when the segment count is known."""
import struct
from typing import List


def merge_bytes(data: List[bytes], num_bytes_length=4, num_segments=None) -> bytes:
    fmt = {1: "B", 2: "H", 4: "I", 8: "L"}[num_bytes_length]
    out = []
    for i, bs in enumerate(data):
        if num_segments is not None:
            assert i < num_segments, "Number of segments exceed predefined {}".format(num_segments)
        if num_segments is None or i < num_segments - 1:
            out.append(struct.pack(fmt, len(bs)))
        out.append(bs)
    return b"".join(out)


def split_merged_bytes(data: bytes, num_bytes_length=4, num_segments=None) -> List[bytes]:
    fmt = {1: "B", 2: "H", 4: "I", 8: "L"}[num_bytes_length]
    pos, parts = 0, []
    while pos < len(data):
        if num_segments is not None and len(parts) >= num_segments - 1:
            parts.append(data[pos:])
            pos = len(data)
        else:
            n = struct.unpack(fmt, data[pos:pos + num_bytes_length])[0]
            pos += num_bytes_length
            parts.append(data[pos:pos + n])
            pos += n
    if num_segments is not None:
        parts.extend([b""] * (num_segments - len(parts)))
    return parts
