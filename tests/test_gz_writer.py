"""The CLI's gzip writer / reader (host only: svb_write_gz, svb_read_gz) against Python's gzip, which is zlib - the library
the reference's gzstream (gzstream.C:53-114) and bwa read these files with."""
import gzip
import os
import random

import pytest

import seeksv_b200.lib as L
from conftest import GOLDEN


def _cases():
    rnd = random.Random(7)
    text = open(os.path.join(GOLDEN, "example", "cancer.clip.txt"), "rb").read()
    return {
        "empty": b"",
        "one_byte": b"A",
        "one_symbol": b"I" * 300000,
        "clip_text": text * 6,                                    # many members
        "incompressible": rnd.randbytes(2_200_000),
        "skewed": bytes(rnd.choices(range(256), weights=[2 ** (-i / 8) for i in range(256)], k=1_500_000)),
        "needs_length_limit": b"".join(bytes([i]) * (2 ** min(i, 21)) for i in range(23)),  # Fibonacci-like: codes > 15 bits
        "piece_edges": b"ACGT" * (16384 * 3) + b"N",              # exact multiples of the 64 KiB deflate piece, odd tail
    }


@pytest.mark.parametrize("level", [None, "1", "6"])
def test_writer_output_is_plain_gzip_and_reader_round_trips(tmp_path, monkeypatch, level):
    if level is None:
        monkeypatch.delenv("SEEKSV_B200_GZ_LEVEL", raising=False)   # Huffman-only dynamic blocks
    else:
        monkeypatch.setenv("SEEKSV_B200_GZ_LEVEL", level)           # zlib at that level
    for name, data in _cases().items():
        p = str(tmp_path / (name + ".gz"))
        L.write_gz(p, data, threads=3)
        raw = open(p, "rb").read()
        assert raw[:4] == b"\x1f\x8b\x08\x04" and raw[12:14] == b"SV", name
        assert gzip.decompress(raw) == data, name
        for reader in ("", "zlib"):     # the two-literals-per-lookup decoder of our own members (falls back to zlib), and zlib alone
            monkeypatch.setenv("SEEKSV_B200_GZ_READ", reader)
            assert L.read_gz(p) == data, (name, reader)


def _members(raw):
    """(member size, ISIZE) of every member of a file of our writers, walked through the 'SV' size fields"""
    out, o = [], 0
    while o < len(raw):
        assert raw[o:o + 4] == b"\x1f\x8b\x08\x04" and raw[o + 12:o + 16] == b"SV\x04\x00", o
        size = int.from_bytes(raw[o + 16:o + 20], "little")
        assert 28 <= size <= len(raw) - o
        out.append((size, int.from_bytes(raw[o + size - 4:o + size], "little")))
        o += size
    return out


def test_default_members_hold_at_most_64_kib_of_text(tmp_path, monkeypatch):
    """what makes a file readable by the device inflate kernel (svb_read_gz_device; getsv reads P.clip.gz through it): the default
    writer's members are 64 KiB of text each; with an explicit zlib level they stay at 1 MiB"""
    data = _cases()["clip_text"]
    monkeypatch.delenv("SEEKSV_B200_GZ_LEVEL", raising=False)
    p = str(tmp_path / "d.gz")
    L.write_gz(p, data, threads=2)
    ms = _members(open(p, "rb").read())
    assert sum(u for _, u in ms) == len(data) and all(u == 65536 for _, u in ms[:-1]) and 0 < ms[-1][1] <= 65536
    monkeypatch.setenv("SEEKSV_B200_GZ_LEVEL", "1")
    L.write_gz(p, data, threads=2)
    ms = _members(open(p, "rb").read())
    assert sum(u for _, u in ms) == len(data) and all(u == 1 << 20 for _, u in ms[:-1])


def test_reader_accepts_foreign_gzip_and_plain_text(tmp_path):
    data = open(os.path.join(GOLDEN, "example", "cancer.clip.txt"), "rb").read()
    p = str(tmp_path / "foreign.gz")
    with gzip.open(p, "wb") as f:       # single member, no 'SV' field: the gzread path (what the reference's igzstream does)
        f.write(data)
    assert L.read_gz(p) == data
    q = str(tmp_path / "plain.txt")
    open(q, "wb").write(data)
    assert L.read_gz(q) == data
    with pytest.raises(L.SvbError):
        L.read_gz(str(tmp_path / "missing.gz"))


def test_fast_reader_rejects_damage_like_zlib_does(tmp_path, monkeypatch):
    """a damaged member must not come back as text from either reader (the fast decoder checks CRC-32 and length, then zlib decides)"""
    rnd = random.Random(9)
    data = bytes(rnd.choices(b"ACGTN\t\n0123456789", k=400000))
    p = str(tmp_path / "x.gz")
    L.write_gz(p, data, threads=2)
    raw = bytearray(open(p, "rb").read())
    for reader in ("", "zlib"):
        monkeypatch.setenv("SEEKSV_B200_GZ_READ", reader)
        for trial in range(20):
            bad = bytearray(raw)
            bad[40 + rnd.randrange(len(raw) - 60)] ^= 1 << rnd.randrange(8)
            q = str(tmp_path / "bad.gz")
            open(q, "wb").write(bad)
            try:
                assert L.read_gz(q) != data or bytes(bad) == bytes(raw)
            except L.SvbError:
                pass
