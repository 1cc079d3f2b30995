"""GPU tests of the zstd frame mode (b2c_zstd_encode_frames / _frames_device: one frame per input of any size, blocks
parsed with the history before them) through the C ABI.  SURVEY section 8 rows a-7, a-15, f-1.  The device's frames
must equal the emulator's byte for byte (tests/test_emu_frames.py checks those against the oracle block by block), decode
with the oracle, libzstd and the GPU decoder."""
import numpy as np
import pytest
import torch

import helpers as H
from emu_util import emu_encode_frames

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("level", [1, 2, 3])
def test_frames_equal_emulator_and_decode(level):
    from compress_b200 import zstd
    tw = H.golden("twain.txt")
    fblock = 49152 if level == 1 else 98304
    inputs = [b"", b"a", tw[:200], tw[:1025], tw[:fblock], tw[:fblock + 1], tw[:3 * fblock + 777], bytes(2 * fblock + 5),
              b"abcd" * (fblock // 2), H.golden("html.txt"), H.golden("e.txt")[:70000], tw]
    enc = zstd.Encoder(level=level, max_chunks=64)
    frames = enc.encode_frames(inputs)
    want = emu_encode_frames(H.emu(), inputs, level=level, dump=False)[0]
    for i, (a, b) in enumerate(zip(frames, want)):
        assert a == b, f"level {level} input {i}: device frame differs from the emulated kernels ({len(a)} vs {len(b)} B)"
    for data, fr in zip(inputs, frames):
        r, dec = H.oracle_decode(fr, len(data) + 64)
        assert r == len(data) and dec == data
        assert H.libzstd_decode(fr, len(data)) == data
    # the GPU decoder reads them too (multi-block frames go to the one-warp decoder)
    d = zstd.Decoder()
    back, codes = d.decode_chunks(frames, [len(x) + 16 for x in inputs])
    assert back == inputs, codes
    # EncodeAll is frame mode
    assert enc.EncodeAll(tw) == frames[-1]
    d.close()
    enc.close()


def test_frames_device_batch_one_mib_frames():
    """64 frames of 1 MiB of synthetic text, device-resident: sizes, offsets, decode of a sample, smaller than the same
    bytes as independent chunks."""
    from compress_b200 import zstd
    n, fs = 64, 1 << 20
    src = H.synth_text_torch(n * fs, "cuda", seed=5)
    enc = zstd.Encoder(level=1)
    dst, foff, fsz = enc.encode_frames_device(src, [i * fs for i in range(n)], [fs] * n)
    torch.cuda.synchronize()
    fo, fz = foff.cpu().numpy().astype(np.int64), fsz.cpu().numpy()
    assert (fz > 0).all()
    assert fo[0] == 0 and (fo[1:] == np.cumsum(fz)[:-1]).all(), "frames are written back to back"
    host = src.cpu().numpy()
    outh = dst[: int(fo[-1] + fz[-1])].cpu().numpy()
    for f in (0, 17, 63):
        fr = bytes(outh[fo[f]:fo[f] + fz[f]])
        assert H.libzstd_decode(fr, fs) == bytes(host[f * fs:(f + 1) * fs])
    _, outs = enc.encode_device(src)
    torch.cuda.synchronize()
    assert int(fz.sum()) < int(outs.sum())
    enc.close()


def test_writer_streams_frames():
    """zstd.Writer (Write / Flush / Close over frame mode): the concatenated frames decode to what was written."""
    import io
    from compress_b200 import zstd
    tw = H.golden("twain.txt")
    sink = io.BytesIO()
    w = zstd.Writer(sink, level=2, frame_bytes=150000)
    for o in range(0, len(tw), 70001):
        w.Write(tw[o:o + 70001])
    w.Close()
    out = sink.getvalue()
    r, dec = H.oracle_decode(out, len(tw) + 64)
    assert r == len(tw) and dec == tw
    d = zstd.Decoder()
    assert d.DecodeAll(out, size_hint=len(tw) + 64) == tw
    d.close()
    sink = io.BytesIO()
    w = zstd.Writer(sink)
    w.Close()
    assert H.oracle_decode(sink.getvalue(), 16)[0] == 0
