"""Seek index of S2 / Snappy streams: the mirror of s2.Index, s2.IndexStream, RemoveIndexHeaders / RestoreIndexHeaders
(s2/index.go) and of the lookup a seeking reader does with it (s2/reader.go ReadSeeker, ExampleIndex_Load in
s2/index_test.go:16-104).

The index is a small host-side table -- (compressed offset, uncompressed offset) pairs at least 1 MiB of content apart, stored
as a skippable chunk (type 0x99) of zig-zag varint deltas against a running prediction -- so it is plain host code here as it
is in the reference; nothing in it touches block contents.  What it buys on the GPU path is `read_range`: a byte range of a
large stream is served by handing only the chunks that cover it to the device stream decoder.
"""
import json

S2IndexHeader = b"s2idx\x00"
S2IndexTrailer = b"\x00xdi2s"
ChunkTypeIndex = 0x99                     # s2/s2.go:107
maxIndexEntries = 1 << 16
minIndexDist = 1 << 20                    # entries closer than this (uncompressed) are not recorded
_maxChunkSize = (1 << 24) - 1             # s2/s2.go:93
_maxBlockSize = 4 << 20                   # s2/s2.go:87
_MAGIC_S2 = b"\xff\x06\x00\x00S2sTwO"
_MAGIC_SNAPPY = b"\xff\x06\x00\x00sNaPpY"


class ErrCorrupt(ValueError):
    """s2.ErrCorrupt"""


class ErrUnsupported(ValueError):
    """s2.ErrUnsupported"""


class ErrUnexpectedEOF(EOFError):
    """io.ErrUnexpectedEOF"""


def put_varint(x):
    """encoding/binary.PutVarint: zig-zag, then base-128 little endian."""
    ux = (x << 1) ^ (x >> 63) if x < 0 else x << 1
    ux &= (1 << 64) - 1
    out = bytearray()
    while ux >= 0x80:
        out.append((ux & 0x7F) | 0x80)
        ux >>= 7
    out.append(ux)
    return bytes(out)


def varint(b, pos=0):
    """encoding/binary.Varint on b[pos:]: (value, bytes read); read <= 0 on a short buffer (0) or an overflow (< 0)."""
    ux = shift = 0
    for i in range(pos, len(b)):
        c = b[i]
        k = i - pos
        if k == 10:
            return 0, -(k + 1)
        if c < 0x80:
            if k == 9 and c > 1:
                return 0, -(k + 1)
            ux |= c << shift
            x = ux >> 1
            if ux & 1:
                x = ~x
            return x, k + 1
        ux |= (c & 0x7F) << shift
        shift += 7
    return 0, 0


def _half(v):
    """Go's v / 2 on int64: truncation toward zero."""
    return v // 2 if v >= 0 else -((-v) // 2)


class Index:
    """s2.Index (s2/index.go:24-33): TotalUncompressed / TotalCompressed (-1 = unknown) and the offset pairs."""

    def __init__(self):
        self.TotalUncompressed = 0
        self.TotalCompressed = 0
        self.info = []                    # [(compressedOffset, uncompressedOffset)]
        self.estBlockUncomp = 0

    def reset(self, max_block):
        self.estBlockUncomp = int(max_block)
        self.TotalCompressed = -1
        self.TotalUncompressed = -1
        self.info = []

    def add(self, compressed_offset, uncompressed_offset):
        """Entries arrive in stream order (s2/index.go:57-88)."""
        if self.info:
            c, u = self.info[-1]
            if u == uncompressed_offset:
                self.info[-1] = (compressed_offset, u)     # nothing was output in between: move the start forward
                return
            if u > uncompressed_offset or c > compressed_offset:
                raise ValueError("internal error: earlier offset received")
            if u + minIndexDist > uncompressed_offset:
                return
        self.info.append((compressed_offset, uncompressed_offset))

    def Find(self, offset):
        """-> (compressedOff, uncompressedOff) of the entry at or before the uncompressed offset; a negative offset counts
        from the end (s2/index.go:90-128)."""
        if self.TotalUncompressed < 0:
            raise ErrCorrupt("index has no total size")
        if offset < 0:
            offset += self.TotalUncompressed
            if offset < 0:
                raise ErrUnexpectedEOF("offset before the start")
        if offset > self.TotalUncompressed:
            raise ErrUnexpectedEOF("offset past the end")
        lo, hi = 0, len(self.info)          # first entry whose uncompressed offset is > offset
        while lo < hi:
            mid = (lo + hi) // 2
            if self.info[mid][1] > offset:
                hi = mid
            else:
                lo = mid + 1
        if lo == 0:
            return (self.info[0] if len(self.info) > 200 else (0, 0))    # (the reference's two code paths differ here)
        return self.info[lo - 1]

    def reduce(self):
        """Stay below maxIndexEntries, and thin out entries of small blocks (s2/index.go:130-152)."""
        if len(self.info) < maxIndexEntries and self.estBlockUncomp >= minIndexDist:
            return
        remove_n = (len(self.info) + 1) // maxIndexEntries
        while self.estBlockUncomp * (remove_n + 1) < minIndexDist and len(self.info) // (remove_n + 1) > 1000:
            remove_n += 1
        self.info = self.info[::remove_n + 1]
        self.estBlockUncomp += self.estBlockUncomp * remove_n

    def appendTo(self, b, uncomp_total, comp_total):
        """Serialise behind b (s2/index.go:154-235): skippable chunk header, "s2idx\\0", totals, estimate, count, the
        uncompressed deltas when they are not implied, the compressed deltas against the running prediction, total size
        (so the index can be found from the end of a stream) and the trailer."""
        self.reduce()
        out = bytearray(b)
        init = len(out)
        out += bytes([ChunkTypeIndex, 0, 0, 0]) + S2IndexHeader
        out += put_varint(uncomp_total) + put_varint(comp_total) + put_varint(self.estBlockUncomp) + put_varint(len(self.info))
        has_uncompressed = 0
        for k, (c, u) in enumerate(self.info):
            if k == 0:
                if u != 0:
                    has_uncompressed = 1
                    break
                continue
            if u != self.info[k - 1][1] + self.estBlockUncomp:
                has_uncompressed = 1
                break
        out.append(has_uncompressed)
        if has_uncompressed:
            for k, (c, u) in enumerate(self.info):
                if k > 0:
                    u -= self.info[k - 1][1] + self.estBlockUncomp
                out += put_varint(u)
        predict = self.estBlockUncomp // 2
        for k, (c, u) in enumerate(self.info):
            if k > 0:
                c -= self.info[k - 1][0] + predict
                predict += _half(c)
            out += put_varint(c)
        out += (len(out) - init + 4 + len(S2IndexTrailer)).to_bytes(4, "little")
        out += S2IndexTrailer
        chunk_len = len(out) - init - 4
        out[init + 1:init + 4] = chunk_len.to_bytes(3, "little")
        return bytes(out)

    def Load(self, b):
        """Parse a serialised index; returns what follows it (s2/index.go:237-374)."""
        b = bytes(b)
        if len(b) <= 4 + len(S2IndexHeader) + len(S2IndexTrailer):
            raise ErrUnexpectedEOF("index too short")
        if b[0] != ChunkTypeIndex:
            raise ErrCorrupt("not an index chunk")
        chunk_len = int.from_bytes(b[1:4], "little")
        p = 4
        if len(b) - p < chunk_len:
            raise ErrUnexpectedEOF("index cut short")
        if b[p:p + 6] != S2IndexHeader:
            raise ErrUnsupported("unknown index header")
        p += 6

        def take(nonneg):
            nonlocal p
            v, n = varint(b, p)
            if n <= 0 or (nonneg and v < 0):
                raise ErrCorrupt("bad varint in index")
            p += n
            return v
        self.TotalUncompressed = take(True)
        self.TotalCompressed = take(False)
        self.estBlockUncomp = take(True)
        entries = take(True)
        if entries > maxIndexEntries:
            raise ErrCorrupt("too many index entries")
        if p >= len(b):
            raise ErrUnexpectedEOF("index cut short")
        has_uncompressed = b[p]
        p += 1
        if has_uncompressed & 1 != has_uncompressed:
            raise ErrCorrupt("bad flag in index")
        us = []
        for k in range(entries):
            u = take(False) if has_uncompressed else 0
            if k > 0:
                u += us[-1] + self.estBlockUncomp
                if u <= us[-1]:
                    raise ErrCorrupt("index offsets not increasing")
            if u < 0:
                raise ErrCorrupt("negative offset in index")
            us.append(u)
        predict = self.estBlockUncomp // 2
        cs = []
        for k in range(entries):
            c = take(False)
            if k > 0:
                new_predict = predict + _half(c)
                c += cs[-1] + predict
                if c <= cs[-1]:
                    raise ErrCorrupt("index offsets not increasing")
                predict = new_predict
            if c < 0:
                raise ErrCorrupt("negative offset in index")
            cs.append(c)
        self.info = list(zip(cs, us))
        if len(b) - p < 4 + len(S2IndexTrailer):
            raise ErrUnexpectedEOF("index cut short")
        p += 4
        if b[p:p + 6] != S2IndexTrailer:
            raise ErrCorrupt("bad index trailer")
        return b[p + 6:]

    def LoadStream(self, rs):
        """Load the index appended to a seekable stream (s2/index.go:376-414)."""
        rs.seek(-10, 2)
        tail = rs.read(10)
        if len(tail) != 10:
            raise ErrUnexpectedEOF("stream too short")
        if tail[4:] != S2IndexTrailer:
            raise ErrUnsupported("no index at the end of the stream")
        sz = int.from_bytes(tail[:4], "little")
        if sz > _maxChunkSize + 4:
            raise ErrCorrupt("index size out of range")
        rs.seek(-sz, 2)
        buf = rs.read(sz)
        if len(buf) != sz:
            raise ErrUnexpectedEOF("stream too short")
        self.Load(buf)

    def JSON(self):
        return json.dumps({"total_uncompressed": self.TotalUncompressed, "total_compressed": self.TotalCompressed,
                           "offsets": [{"compressed": c, "uncompressed": u} for c, u in self.info],
                           "est_block_uncompressed": self.estBlockUncomp}, indent=2).encode()


def _decoded_len(buf):
    """s2.DecodedLen on a block (uvarint, at most 32 bits)."""
    v = shift = 0
    for k, c in enumerate(buf[:10]):
        v |= (c & 0x7F) << shift
        if c < 0x80:
            if k > 4 or v > 0xFFFFFFFF:
                raise ErrCorrupt("bad block length")
            return v
        shift += 7
    raise ErrCorrupt("bad block length")


def IndexStream(r, est_block=0):
    """Index an existing stream by walking its chunk headers; block contents are not verified (s2/index.go:416-513).
    r: bytes-like or a file-like object.  est_block: the writer's block size when known (s2.Writer resets its index with it,
    s2/writer.go); 0 = take the first block's size, as IndexStream does."""
    data = r if isinstance(r, (bytes, bytearray, memoryview)) else r.read()
    data = bytes(data)
    idx = Index()
    idx.estBlockUncomp = int(est_block)
    p, n = 0, len(data)
    read_header = False
    while True:
        if p == n:
            return idx.appendTo(b"", idx.TotalUncompressed, idx.TotalCompressed)
        if n - p < 4:
            raise ErrUnexpectedEOF("chunk header cut short")
        start = idx.TotalCompressed
        idx.TotalCompressed += 4
        typ = data[p]
        if not read_header:
            if typ != 0xFF:
                raise ErrCorrupt("stream does not start with an identifier")
            read_header = True
        chunk_len = int.from_bytes(data[p + 1:p + 4], "little")
        if chunk_len < 4:
            raise ErrCorrupt("chunk too short")
        idx.TotalCompressed += chunk_len
        body = data[p + 4:p + 4 + chunk_len]
        if len(body) != chunk_len:
            raise ErrUnexpectedEOF("chunk cut short")
        p += 4 + chunk_len
        if typ in (0x00, 0x01):
            d = _decoded_len(body[4:]) if typ == 0x00 else chunk_len - 4
            if d > _maxBlockSize:
                raise ErrCorrupt("block too large")
            if idx.estBlockUncomp == 0:
                idx.estBlockUncomp = d
            idx.add(start, idx.TotalUncompressed)
            idx.TotalUncompressed += d
        elif typ == 0xFF:
            if chunk_len != 6 or body not in (b"S2sTwO", b"sNaPpY"):
                raise ErrCorrupt("bad stream identifier")
        elif typ <= 0x7F:
            raise ErrUnsupported("reserved unskippable chunk")
        # 0x80-0xfe: padding and skippable chunks


def RemoveIndexHeaders(b):
    """Strip the chunk header, "s2idx\\0", the size and the trailer (20 bytes); None when b is not an index
    (s2/index.go:541-577)."""
    b = bytes(b)
    if len(b) <= 4 + 6 + 6 + 4 or b[0] != ChunkTypeIndex:
        return None
    chunk_len = int.from_bytes(b[1:4], "little")
    b = b[4:]
    if len(b) < chunk_len:
        return None
    b = b[:chunk_len]
    if b[:6] != S2IndexHeader or not b.endswith(S2IndexTrailer):
        return None
    b = b[6:-6]
    if len(b) < 4:
        return None
    return b[:-4]


def RestoreIndexHeaders(body):
    """Inverse of RemoveIndexHeaders (s2/index.go:579-602)."""
    if len(body) == 0:
        return body
    out = bytearray([ChunkTypeIndex, 0, 0, 0]) + S2IndexHeader + bytes(body)
    out += (len(out) + 4 + len(S2IndexTrailer)).to_bytes(4, "little") + S2IndexTrailer
    out[1:4] = (len(out) - 4).to_bytes(3, "little")
    return bytes(out)


def read_range(stream, start, length, decode_stream, index=None):
    """Bytes [start, start+length) of the stream's content (length None = to the end; negative start = from the end), decoding
    only the chunks that cover them: Index.Find gives the chunk to start at, the stream identifier is put back in front of the
    slice (the reference starts a reader there with ReaderIgnoreStreamIdentifier and skips forward: ExampleIndex_Load),
    and `decode_stream(bytes) -> bytes` -- the device stream decoder -- does the rest.  index: serialised index bytes or an
    Index; default: the one appended to the stream."""
    stream = bytes(stream)
    if isinstance(index, Index):
        idx = index
    else:
        idx = Index()
        if index is None:
            import io
            idx.LoadStream(io.BytesIO(stream))
        else:
            idx.Load(index)
    total = idx.TotalUncompressed
    if start < 0:
        start += total
    if start < 0 or start > total:
        raise ErrUnexpectedEOF("offset outside the stream")
    end = total if length is None else min(total, start + length)
    if end <= start:
        return b""
    c0, u0 = idx.Find(start)
    c1 = len(stream) if idx.TotalCompressed < 0 else min(len(stream), idx.TotalCompressed)
    for c, u in idx.info:                     # first recorded chunk that starts at or after the end of the range
        if u >= end:
            c1 = c
            break
    if c0 == 0:
        part = stream[:c1]
    else:
        magic = _MAGIC_SNAPPY if stream[:10] == _MAGIC_SNAPPY else _MAGIC_S2
        part = magic + stream[c0:c1]
    out = decode_stream(part)
    return out[start - u0:end - u0]
