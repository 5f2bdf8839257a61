"""CPU tests of the dataset path: the native TFRecord / tf.train.Example reader and writer in
libflownet.so against the pure-Python restatement in oracle/tfrecord_oracle.py, CRC-32C against the RFC 3720
vectors, the shuffle queue's batching rules, and the MechSys crop rule."""
import os
import struct

import numpy as np
import pytest

from oracle import tfrecord_oracle as tfo


@pytest.fixture(scope="module")
def fi(built_lib):
    import input.flow_input as flow_input
    return flow_input


def test_crc32c_rfc3720_vectors(built_lib):
    vecs = [(b"123456789", 0xE3069283), (bytes(32), 0x8A9136AA), (b"\xff" * 32, 0x62A8AB43),
            (bytes(range(32)), 0x46DD794E), (bytes(range(31, -1, -1)), 0x113FDB5C), (b"", 0)]
    for data, want in vecs:
        assert tfo.crc32c(data) == want
        arr = np.frombuffer(data, np.uint8)
        assert built_lib.lib.fn_crc32c(arr.ctypes.data if len(data) else None, len(data)) == want
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 9, 63, 1000):
        d = rng.integers(0, 256, n, dtype=np.uint8)
        assert built_lib.lib.fn_crc32c(d.ctypes.data, n) == tfo.crc32c(d.tobytes())
        assert built_lib.lib.fn_crc32c_portable(d.ctypes.data, n) == tfo.crc32c(d.tobytes())


def _examples(n, H=6, W=10, seed=0):
    rng = np.random.default_rng(seed)
    b = (rng.random((n, H, W, 1)) > 0.6).astype(np.uint8)
    s = rng.standard_normal((n, H, W, 2)).astype(np.float32)
    return b, s


def test_native_writer_matches_oracle_bytes_and_reads_back(fi, tmp_path):
    b, s = _examples(5)
    path = str(tmp_path / "a.tfrecords")
    fi.write_tfrecords(path, b, s)
    raw = open(path, "rb").read()
    want = b"".join(tfo.frame(tfo.encode_example({"boundary": b[i].tobytes(), "sflow": s[i].tobytes()})) for i in range(5))
    assert raw == want                                       # byte-identical files
    recs = list(tfo.read_records(raw))
    assert len(recs) == 5
    for i, r in enumerate(recs):
        d = tfo.decode_example(r)
        assert d["boundary"] == b[i].tobytes() and d["sflow"] == s[i].tobytes()
    got = list(fi.read_data(path, shape=(6, 10)))
    assert len(got) == 5
    for i, (gb, gs) in enumerate(got):
        assert gb.dtype == np.uint8 and gs.dtype == np.float32
        assert np.array_equal(gb, b[i]) and np.array_equal(gs, s[i])


def test_native_reader_accepts_any_feature_order_and_extra_features(fi, tmp_path):
    b, s = _examples(2, seed=1)
    path = str(tmp_path / "b.tfrecords")
    with open(path, "wb") as f:
        f.write(tfo.frame(tfo.encode_example({"sflow": s[0].tobytes(), "extra": b"xyz", "boundary": b[0].tobytes()})))
        f.write(tfo.frame(tfo.encode_example({"boundary": b[1].tobytes(), "sflow": s[1].tobytes()})))
    got = list(fi.read_data(path, shape=(6, 10)))
    assert np.array_equal(got[0][0], b[0]) and np.array_equal(got[0][1], s[0])
    assert np.array_equal(got[1][0], b[1]) and np.array_equal(got[1][1], s[1])


def test_native_reader_rejects_corrupt_truncated_and_mis_sized(fi, tmp_path):
    b, s = _examples(2, seed=2)
    good = b"".join(tfo.frame(tfo.encode_example({"boundary": b[i].tobytes(), "sflow": s[i].tobytes()})) for i in range(2))
    cases = {"flip": good[:40] + bytes([good[40] ^ 1]) + good[41:], "trunc": good[:-3], "lencrc": good[:8] + b"\0\0\0\0" + good[12:]}
    for name, data in cases.items():
        p = str(tmp_path / (name + ".tfrecords"))
        open(p, "wb").write(data)
        with pytest.raises(ValueError):
            fi.TFRecordFile(p)
    p = str(tmp_path / "ok.tfrecords")
    open(p, "wb").write(good)
    with pytest.raises(ValueError):                          # wrong image size for these records
        list(fi.read_data(p, shape=(8, 8)))
    p2 = str(tmp_path / "missing.tfrecords")
    open(p2, "wb").write(tfo.frame(tfo.encode_example({"boundary": b[0].tobytes()})))
    with pytest.raises(ValueError):
        list(fi.read_data(p2, shape=(6, 10)))
    p3 = str(tmp_path / "empty.tfrecords")
    open(p3, "wb").write(b"")
    assert len(fi.TFRecordFile(p3)) == 0


def test_index_rejects_a_file_cut_at_every_offset_of_the_last_record(built_lib):
    """A dataset file cut off mid-write must be a clean FN_E_INVAL at every cut point -- in particular inside the 12-byte
    header and the 4-byte footer, where a 12..15-byte remainder used to pass the length check -- and must
    never read past the buffer: the image is placed flush against the end of its allocation, with and without CRC checks."""
    b, s = _examples(2, H=3, W=4, seed=3)
    recs = [tfo.frame(tfo.encode_example({"boundary": b[i].tobytes(), "sflow": s[i].tobytes()})) for i in range(2)]
    good = recs[0] + recs[1]
    lib = built_lib.lib
    for cut in range(len(recs[0]), len(good) + 1):
        data = np.frombuffer(good[:cut], np.uint8).copy()            # exactly `cut` bytes: no slack behind the image
        for crc in (0, 1):
            n = lib.fn_tfrecord_index(data.ctypes.data if cut else None, cut, None, None, 0, crc)
            if cut == len(recs[0]):
                assert n == 1
            elif cut == len(good):
                assert n == 2
            else:
                assert n == -1, (cut, crc, n)
    # a 14-byte image with a valid length CRC and a huge length: header complete, footer missing
    import struct as _st
    hdr = _st.pack("<Q", 1 << 20)
    img = np.frombuffer(hdr + _st.pack("<I", tfo.masked_crc(hdr)) + b"\0\0", np.uint8).copy()
    for crc in (0, 1):
        assert lib.fn_tfrecord_index(img.ctypes.data, img.size, None, None, 0, crc) == -1


def test_input_queue_fifo_and_shuffle(fi, tmp_path):
    b, s = _examples(10, seed=3)
    p1, p2 = str(tmp_path / "1.tfrecords"), str(tmp_path / "2.tfrecords")
    fi.write_tfrecords(p1, b[:6], s[:6])
    fi.write_tfrecords(p2, b[6:], s[6:])
    q = fi.FlowInputQueue([p1, p2], 4, shape=(6, 10), device="cpu", shuffle=False)
    seen = []
    for _ in range(5):                                       # 20 examples = two epochs, in file order
        gb, gs = q.next()
        assert tuple(gb.shape) == (4, 6, 10, 1) and tuple(gs.shape) == (4, 6, 10, 2)
        seen += [x.numpy().tobytes() for x in gb]
    q.close()
    assert seen == [b[i % 10].tobytes() for i in range(20)]
    q = fi.FlowInputQueue([p1, p2], 5, shape=(6, 10), min_after_dequeue=5, seed=1, device="cpu")
    assert q.capacity == 5 + 3 * 5                           # flow_input.py:38
    key = {s[i].tobytes(): i for i in range(10)}
    order = []
    for _ in range(8):
        gb, gs = q.next()
        for j in range(5):
            i = key[gs[j].numpy().tobytes()]
            assert gb[j].numpy().tobytes() == b[i].tobytes()  # boundary and target stay paired
            order.append(i)
    q.close()
    assert order != sorted(order) and set(order) == set(range(10))
    counts = np.bincount(order, minlength=10)
    assert counts.max() - counts.min() <= 3                  # a sliding shuffle window, not sampling with replacement


def test_input_queue_rank_shards_are_disjoint_and_cover_the_stream(fi, tmp_path):
    """data-parallel training: rank r of R reads every R-th record of the stream (train/flow_train.py passes shard=(rank, world))"""
    b, s = _examples(12, seed=4)
    p1, p2 = str(tmp_path / "1.tfrecords"), str(tmp_path / "2.tfrecords")
    fi.write_tfrecords(p1, b[:7], s[:7])
    fi.write_tfrecords(p2, b[7:], s[7:])
    key = {b[i].tobytes(): i for i in range(12)}
    seen = []
    for rank in range(3):
        q = fi.FlowInputQueue([p1, p2], 4, shape=(6, 10), device="cpu", shuffle=False, shard=(rank, 3))
        gb, _ = q.next()
        q.close()
        seen.append([key[x.numpy().tobytes()] for x in gb])
    assert seen == [[0, 3, 6, 9], [1, 4, 7, 10], [2, 5, 8, 11]]
    with pytest.raises(ValueError):
        fi.FlowInputQueue([p1], 1, shape=(6, 10), device="cpu", shard=(3, 3))


def test_mechsys_crop_rule():
    from utils import flow_reader
    H, W = 4, 6
    vel = np.arange(H * (W + 128) * 3, dtype=np.float32)
    gam = np.arange(H * (W + 128), dtype=np.float32)
    f = flow_reader.crop_flow(vel, (H, W))
    g = flow_reader.crop_boundary(gam, (H, W))
    assert f.shape == (H, W, 2) and g.shape == (H, W, 1)
    assert f[1, 2, 1] == vel[(1 * (W + 128) + 2) * 3 + 1] and g[3, 5, 0] == gam[3 * (W + 128) + 5]


def test_native_batch_parse_matches_single_record_parse(fi, tmp_path):
    b, s = _examples(9, seed=7)
    path = str(tmp_path / "c.tfrecords")
    fi.write_tfrecords(path, b, s)
    f = fi.TFRecordFile(path)
    idx = [8, 0, 3, 3, 7, 1]
    for threads in (1, 3, 16):
        nb = np.zeros((len(idx), 6, 10, 1), np.uint8)
        ns = np.zeros((len(idx), 6, 10, 2), np.float32)
        f.read_batch_into(idx, nb, ns, threads)
        assert np.array_equal(nb, b[idx]) and np.array_equal(ns, s[idx])
    with pytest.raises(ValueError):
        f.read_batch_into([0, 1], np.zeros((2, 8, 8, 1), np.uint8), np.zeros((2, 8, 8, 2), np.float32), 2)
