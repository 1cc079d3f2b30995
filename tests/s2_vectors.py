"""Decode known-answer vectors of the reference: s2/s2_test.go:252-470 (TestDecode) and :214-250 (TestInvalidVarint)."""
_LIT40 = bytes(range(40))

# (input, want, ok)
DECODE_TABLE = [
    (b"\x00", b"", True),
    (b"\x03" + b"\x08\xff\xff\xff", b"\xff\xff\xff", True),
    (b"\x02" + b"\x08\xff\xff\xff", b"", False),
    (b"\x03" + b"\x08\xff\xff", b"", False),
    (b"\x28" + b"\x9c" + _LIT40, _LIT40, True),
    (b"\x01" + b"\xf0", b"", False),
    (b"\x03" + b"\xf0\x02\xff\xff\xff", b"\xff\xff\xff", True),
    (b"\x01" + b"\xf4\x00", b"", False),
    (b"\x03" + b"\xf4\x02\x00\xff\xff\xff", b"\xff\xff\xff", True),
    (b"\x01" + b"\xf8\x00\x00", b"", False),
    (b"\x03" + b"\xf8\x02\x00\x00\xff\xff\xff", b"\xff\xff\xff", True),
    (b"\x01" + b"\xfc\x00\x00\x00", b"", False),
    (b"\x01" + b"\xfc\x02\x00\x00\x00\xff\xff\xff", b"", False),
    (b"\x04" + b"\xfc\x02\x00\x00\x00\xff", b"", False),
    (b"\x03" + b"\xfc\x02\x00\x00\x00\xff\xff\xff", b"\xff\xff\xff", True),
    (b"\x04" + b"\x01", b"", False),
    (b"\x04" + b"\x02\x00", b"", False),
    (b"\x04" + b"\x03\x00\x00\x00", b"", False),
    (b"\x04" + b"\x0cabcd", b"abcd", True),
    (b"\x0d" + b"\x0cabcd" + b"\x15\x04", b"abcdabcdabcda", True),
    (b"\x08" + b"\x0cabcd" + b"\x01\x04", b"abcdabcd", True),
    (b"\x08" + b"\x0cabcd" + b"\x01\x02", b"abcdcdcd", True),
    (b"\x08" + b"\x0cabcd" + b"\x01\x01", b"abcddddd", True),
    (b"\x08" + b"\x0cabcd" + b"\x01\x00", b"", False),
    (b"\x0d" + b"\x0cabcd" + b"\x01\x01" + b"\x00z" + b"\x01\x00", b"abcdddddzzzzz", True),
    (b"\x09" + b"\x0cabcd" + b"\x01\x04", b"", False),
    (b"\x08" + b"\x0cabcd" + b"\x01\x05", b"", False),
    (b"\x07" + b"\x0cabcd" + b"\x01\x04", b"", False),
    (b"\x06" + b"\x0cabcd" + b"\x06\x03\x00", b"abcdbc", True),
    (b"\x06" + b"\x0cabcd" + b"\x07\x03\x00\x00\x00", b"abcdbc", True),
]

INVALID_VARINT = [b"\xff", b"\xff\xff\xff\xff\xff\xff\xff\xff\xff\xff\x00", b"\x80\x80\x80\x80\x10",
                  b"\x84\x80\x80\x80\x80\x80\x80\x00" + b"\x00" * 7 + b"\x30"]


def decode_copy4_case():
    """s2/s2_test.go:472-500 TestDecodeCopy4: (input, want)."""
    dots = b"." * 65536
    inp = b"\x89\x80\x04" + b"\x0cpqrs" + b"\xf4\xff\xff" + dots + b"\x13\x04\x00\x01\x00"
    return inp, b"pqrs" + dots + b"pqrs."


def decode_length_offset_cases():
    """s2/s2_test.go:502-597 TestDecodeLengthOffset: literal + copy(length, offset) + literal, 18 x 18 x 19 cases."""
    prefix, suffix = b"abcdefghijklmnopqr", b"ABCDEFGHIJKLMNOPQR"
    out = []
    for length in range(1, 19):
        for offset in range(1, 19):
            for suffix_len in range(0, 19):
                total = len(prefix) + length + suffix_len
                inp = bytes([total]) + bytes([4 * (len(prefix) - 1)]) + prefix + bytes([2 + 4 * (length - 1), offset, 0])
                if suffix_len:
                    inp += bytes([4 * (suffix_len - 1)]) + suffix[:suffix_len]
                want = bytearray(prefix)
                for _ in range(length):
                    want.append(want[-offset])
                want += suffix[:suffix_len]
                out.append((inp, bytes(want)))
    return out
