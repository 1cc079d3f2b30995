"""Reference model (pure Python / numpy) of the S2 / Snappy framing format as the reference writes and reads it
(s2/writer.go:395-470 block chunks, s2/s2.go:75-126 constants and the masked CRC32-C, s2/reader.go:249-420 chunk handling).
Test infrastructure: it builds streams with the oracle's block encoders and reads streams back with the oracle's s2Decode."""
import numpy as np

POLY = 0x82F63B78          # CRC-32C (Castagnoli), reflected
_T = np.zeros(256, dtype=np.uint32)
for _i in range(256):
    _c = _i
    for _ in range(8):
        _c = (_c >> 1) ^ POLY if _c & 1 else _c >> 1
    _T[_i] = _c
_TL = [int(x) for x in _T]

MAGIC_S2 = b"\xff\x06\x00\x00S2sTwO"
MAGIC_SNAPPY = b"\xff\x06\x00\x00sNaPpY"


def crc32c(data, c=0):
    c ^= 0xFFFFFFFF
    for b in bytes(data):
        c = _TL[(c ^ b) & 0xFF] ^ (c >> 8)
    return c ^ 0xFFFFFFFF


def masked_crc(data):
    """crc() of s2/s2.go:120-126: rotate right by 15, add the constant."""
    c = crc32c(data)
    return (((c >> 15) | (c << 17)) + 0xA282EAD8) & 0xFFFFFFFF


def multmodp(a, b):
    m, p = 1 << 31, 0
    while True:
        if a & m:
            p ^= b
            if (a & (m - 1)) == 0:
                break
        m >>= 1
        b = (b >> 1) ^ POLY if b & 1 else b >> 1
    return p


_X2N = [1 << 30]
for _ in range(31):
    _X2N.append(multmodp(_X2N[-1], _X2N[-1]))


def x2nmodp(n, k):
    p = 1 << 31
    while n:
        if n & 1:
            p = multmodp(_X2N[k & 31], p)
        n >>= 1
        k += 1
    return p


def crc_combine(c1, c2, len2):
    return multmodp(x2nmodp(len2, 3), c1) ^ c2


def write_stream(data, encode_block, block_size=65536, snappy=False, extra_chunks=False):
    """Writer.EncodeBuffer's stream for `data`: identifier, then per block a compressed chunk (uvarint + body from
    encode_block(block) -> bytes, b'' = not compressible) or an uncompressed chunk.  extra_chunks: also a padding chunk
    and a skippable chunk between blocks (legal, s2/reader.go:392-420)."""
    out = bytearray(MAGIC_SNAPPY if snappy else MAGIC_S2)
    for o in range(0, len(data), block_size):
        blk = data[o:o + block_size]
        body = encode_block(blk)
        crc = masked_crc(blk)
        if body:
            n, uv = len(blk), bytearray()
            while n >= 0x80:
                uv.append((n & 0x7F) | 0x80); n >>= 7
            uv.append(n)
            payload, typ = bytes(uv) + body, 0x00
        else:
            payload, typ = blk, 0x01
        ln = 4 + len(payload)
        out += bytes([typ, ln & 0xFF, (ln >> 8) & 0xFF, (ln >> 16) & 0xFF]) + crc.to_bytes(4, "little") + payload
        if extra_chunks:
            out += bytes([0xFE, 3, 0, 0]) + b"\0\0\0" + bytes([0x80, 2, 0, 0]) + b"zz"
    return bytes(out)


def read_stream(stream, decode_block, max_block=4 << 20):
    """Reader.Read's chunk walk.  decode_block(bytes, decoded_len) -> bytes or None.  Returns bytes, or raises ValueError with
    the reference's error class ('corrupt', 'crc', 'unsupported')."""
    out = bytearray()
    o, seen, snappy = 0, False, False
    while o < len(stream):
        if o + 4 > len(stream):
            raise ValueError("corrupt")           # io.ErrUnexpectedEOF
        typ = stream[o]
        ln = stream[o + 1] | (stream[o + 2] << 8) | (stream[o + 3] << 16)
        o += 4
        if not seen:
            if typ != 0xFF:
                raise ValueError("corrupt")
            seen = True
        if typ in (0x00, 0x01):
            if ln < 4 or o + ln > len(stream):
                raise ValueError("corrupt")
            want = int.from_bytes(stream[o:o + 4], "little")
            body = stream[o + 4:o + ln]
            if typ == 0x00:
                n, shift, k = 0, 0, 0
                while True:
                    if k >= len(body) or k >= 10:
                        raise ValueError("corrupt")
                    b = body[k]; k += 1
                    n |= (b & 0x7F) << shift
                    if b < 0x80:
                        break
                    shift += 7
                if n > max_block or (snappy and n > 65536):
                    raise ValueError("corrupt")
                dec = decode_block(body, n)
                if dec is None:
                    raise ValueError("corrupt")
            else:
                if len(body) > max_block or (snappy and len(body) > 65536):
                    raise ValueError("corrupt")
                dec = body
            if masked_crc(dec) != want:
                raise ValueError("crc")
            out += dec
        elif typ == 0xFF:
            if ln != 6 or o + 6 > len(stream):
                raise ValueError("corrupt")
            if stream[o:o + 6] == b"S2sTwO":
                snappy = False
            elif stream[o:o + 6] == b"sNaPpY":
                snappy = True
            else:
                raise ValueError("corrupt")
        elif typ <= 0x7F:
            raise ValueError("unsupported")
        else:
            if o + ln > len(stream):
                raise ValueError("corrupt")
        o += ln
    return bytes(out)
