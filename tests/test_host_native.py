"""Host-side native helpers of libtt (csrc/tt_host.cu): CRC32C, TFRecord framing, tf.train.Example parsing, the
vocabulary hash map -- and the TF-free TFRecord writer / dataset classes built on them.  No GPU needed.

Known answers: CRC-32C check values from RFC 3720 (B.4); the serialized bytes of two one-feature tf.train.Example
messages as TensorFlow itself emits them (protobuf wire format, hand-assembled below); the TFRecord framing rule of
tensorflow/core/lib/io/record_writer.cc (u64 length, masked crc, payload, masked crc; mask = rotr15(crc) + 0xa282ead8).
"""
import ctypes
import os
import struct

import numpy as np
import pytest

from pkg import _native as N
from pkg.modelling import _device as D
from pkg.modelling.tfrecord_dataset import TFRecordDatasetFactory, read_tfrecord_file
from pkg.schema import dtypes as tt
from pkg.schema.features import Feature, FeatureFamily
from pkg.tfrecord_writer.tfrecord_writer import TFRecordWriter


def _crc32c_py(data: bytes) -> int:   # bitwise reference
    crc = 0xFFFFFFFF
    for b in data:
        crc ^= b
        for _ in range(8):
            crc = (crc >> 1) ^ 0x82F63B78 if crc & 1 else crc >> 1
    return crc ^ 0xFFFFFFFF


def _mask(c: int) -> int:
    return (((c >> 15) | (c << 17)) + 0xa282ead8) & 0xFFFFFFFF


def test_crc32c_known_answers(lib):
    assert lib.tt_crc32c(b"123456789", 9) == 0xE3069283
    assert lib.tt_crc32c(bytes(32), 32) == 0x8A9136AA
    assert lib.tt_crc32c(b"\xff" * 32, 32) == 0x62A8AB43
    assert lib.tt_crc32c(bytes(range(32)), 32) == 0x46DD794E
    assert lib.tt_crc32c(b"", 0) == 0
    rng = np.random.default_rng(0)
    for n in (1, 7, 8, 9, 63, 1000, 4097):
        buf = rng.integers(0, 256, size=n, dtype=np.uint8).tobytes()
        want = _crc32c_py(buf)
        assert lib.tt_crc32c(buf, n) == want and lib.tt_crc32c_portable(buf, n) == want
        assert lib.tt_crc32c_masked(buf, n) == _mask(want)


EXAMPLE_BYTES_A_X = b"\n\x0c\n\n\n\x01a\x12\x05\n\x03\n\x01x"                    # {"a": bytes_list [b"x"]}
EXAMPLE_FLOAT_F_1 = b"\n\x0f\n\r\n\x01f\x12\x08\x12\x06\n\x04\x00\x00\x80?"      # {"f": float_list [1.0]}


def test_example_encoding_known_answers():
    fa = Feature("a", tt.string, FeatureFamily.QUERY, embedding_size=2)
    ff = Feature("f", tt.float32, FeatureFamily.QUERY)
    assert TFRecordWriter([fa])._get_features_from_row({"a": "x"}) == EXAMPLE_BYTES_A_X
    assert TFRecordWriter([ff])._get_features_from_row({"f": 1.0}) == EXAMPLE_FLOAT_F_1
    with pytest.raises(TypeError):
        TFRecordWriter([fa])._parse_feature("x", "int64")


def test_tfrecord_framing_and_scan(lib, tmp_path):
    payload = EXAMPLE_BYTES_A_X
    out = ctypes.create_string_buffer(len(payload) + 16)
    assert lib.tt_tfrecord_frame(payload, len(payload), out) == 0
    raw = out.raw
    assert raw[:8] == struct.pack("<Q", len(payload))
    assert struct.unpack("<I", raw[8:12])[0] == _mask(_crc32c_py(raw[:8]))
    assert raw[12:12 + len(payload)] == payload
    assert struct.unpack("<I", raw[12 + len(payload):])[0] == _mask(_crc32c_py(payload))
    two = raw + raw
    off = (ctypes.c_int64 * 4)()
    ln = (ctypes.c_int64 * 4)()
    assert lib.tt_tfrecord_scan(two, len(two), 1, off, ln, 4) == 2
    assert list(off[:2]) == [12, 12 + len(raw)] and list(ln[:2]) == [len(payload)] * 2
    bad = bytearray(two)
    bad[20] ^= 1                                   # flip one payload bit: the data CRC must catch it
    assert lib.tt_tfrecord_scan(bytes(bad), len(bad), 1, off, ln, 4) == -5 and b"CRC" in lib.tt_last_error()
    assert lib.tt_tfrecord_scan(bytes(bad), len(bad), 0, off, ln, 4) == 2          # unchecked scan still walks the framing
    assert lib.tt_tfrecord_scan(two[:-3], len(two) - 3, 1, off, ln, 4) == -5       # truncated


def _features():
    return [Feature("customer_id", tt.string, FeatureFamily.QUERY, embedding_size=4),
            Feature("age", tt.float32, FeatureFamily.QUERY),
            Feature("article_id", tt.string, FeatureFamily.CANDIDATE, embedding_size=4)]


def test_writer_reader_round_trip(tmp_path):
    feats = _features()
    rng = np.random.default_rng(1)
    n = 257
    cols = {"customer_id": np.array([f"c{rng.integers(0, 50)}" * int(rng.integers(1, 4)) for _ in range(n)], dtype=object),
            "age": rng.random(n).astype(np.float32),
            "article_id": np.array([f"{rng.integers(0, 10 ** 9):010d}" for _ in range(n)], dtype=object)}
    path = str(tmp_path / "train" / "train.tfrecord")
    TFRecordWriter(feats).write_tfrecords(cols, path, max_file_size=100)
    files = sorted(os.listdir(tmp_path / "train"))
    assert files == ["train_0.tfrecord", "train_1.tfrecord", "train_2.tfrecord"]      # 100 + 100 + 57 rows
    first = read_tfrecord_file(str(tmp_path / "train" / files[0]), feats)
    assert len(first["age"]) == 100
    ds = TFRecordDatasetFactory(feats).create_tfrecord_dataset(str(tmp_path / "train"), batch_size=64)
    batches = list(ds)
    assert [b["age"].shape for b in batches] == [(64, 1)] * 4 + [(1, 1)]
    got = {k: np.concatenate([b[k] for b in batches]).reshape(-1) for k in cols}
    np.testing.assert_array_equal(got["age"], cols["age"])
    assert [v.decode() for v in got["customer_id"]] == list(cols["customer_id"])
    assert [v.decode() for v in got["article_id"]] == list(cols["article_id"])
    assert batches[0]["customer_id"].dtype.kind == "S"
    # a subset of the features parses the same files (candidate-only view, reference runner.py:37,55-58)
    cand = list(TFRecordDatasetFactory(feats[2:]).create_tfrecord_dataset(str(tmp_path / "train"), batch_size=300))
    assert list(cand[0].keys()) == ["article_id"] and cand[0]["article_id"].shape == (257, 1)
    # a feature the files do not hold is an error, as parse_single_example(FixedLenFeature) raises
    with pytest.raises(ValueError):
        list(TFRecordDatasetFactory([Feature("nope", tt.string, FeatureFamily.QUERY, embedding_size=2)]).create_tfrecord_dataset(str(tmp_path / "train")))


def test_dataset_shuffle_map_and_unbatched(tmp_path):
    feats = _features()
    n = 50
    cols = {"customer_id": np.array([f"c{i}" for i in range(n)], dtype=object), "age": np.arange(n, dtype=np.float32),
            "article_id": np.array([f"a{i}" for i in range(n)], dtype=object)}
    TFRecordWriter(feats).write_tfrecords(cols, str(tmp_path / "d" / "x"))
    fac = TFRecordDatasetFactory(feats)
    assert set(fac.feature_description) == {"customer_id", "age", "article_id"}
    plain = fac.create_tfrecord_dataset(str(tmp_path / "d"))
    items = list(plain)
    assert len(items) == n and items[3]["age"].shape == (1,) and items[3]["age"][0] == 3.0
    shuf = fac.create_tfrecord_dataset(str(tmp_path / "d"), batch_size=10, shuffle_size=7)
    shuf.seed = 3
    a = np.concatenate([b["age"] for b in shuf]).reshape(-1)
    assert sorted(a.tolist()) == list(range(n)) and a.tolist() != list(range(n))
    assert max(int(v) - i for i, v in enumerate(a)) < 7 + 1                      # a buffer of 7 cannot pull an element further forward
    b = np.concatenate([x["age"] for x in shuf]).reshape(-1)
    np.testing.assert_array_equal(a, b)                                            # re-iterable, same seed
    mapped = shuf.map(lambda x: ({"customer_id": x["customer_id"]}, x["article_id"]))
    q, t = next(iter(mapped))
    assert list(q) == ["customer_id"] and t.shape == (10, 1)


def test_vocab_native_lookup_matches_string_lookup():
    rng = np.random.default_rng(2)
    words = [f"id{rng.integers(0, 5000)}" for _ in range(3000)] + ["", "ünïcode", "a" * 70, "dup", "dup"]
    v = D.Vocab(words)
    table = {}
    for i, w in enumerate(words):
        table.setdefault(w, i + 1)                                                # first occurrence wins
    queries = [f"id{rng.integers(0, 6000)}" for _ in range(10000)] + ["", "ünïcode", "a" * 70, "a" * 71, "dup", "zzz"]
    want = np.array([table.get(q, 0) for q in queries], dtype=np.int32)
    np.testing.assert_array_equal(v.encode(np.array(queries, dtype=object)), want)
    np.testing.assert_array_equal(v.encode(np.array([q.encode() for q in queries], dtype=object)), want)
    np.testing.assert_array_equal(v.encode(np.array(queries)), want)              # '<U' array
    np.testing.assert_array_equal(v.encode(np.array(queries).reshape(-1, 1)), want)
    assert v.encode([]).shape == (0,) and v.rows == len(words) + 1 and v.token(0) == "[UNK]" and v.token(1) == words[0]
    import pickle

    v2 = pickle.loads(pickle.dumps(v))                                            # Schema objects are pickled (schema.py)
    np.testing.assert_array_equal(v2.encode(queries), want)


def _frame(lib, payload: bytes) -> bytes:
    out = ctypes.create_string_buffer(len(payload) + 16)
    assert lib.tt_tfrecord_frame(payload, len(payload), out) == 0
    return out.raw


def test_corrupted_files_are_rejected_not_crashed_on(lib, tmp_path):
    """The native readers see files from disk: truncated, bit-flipped, random or absurd-length framing must come back as an error
    (or as slices inside the file), and a well-framed record whose payload is not a valid Example as a ValueError -- never a crash."""
    rng = np.random.default_rng(5)
    good = b"".join(_frame(lib, bytes(rng.integers(0, 256, int(rng.integers(0, 40)), dtype=np.uint8))) for _ in range(6))
    for _ in range(3000):
        b = bytearray(good)
        mode = int(rng.integers(0, 4))
        if mode == 0:
            b = b[:int(rng.integers(0, len(b) + 1))]
        elif mode == 1:
            for _ in range(int(rng.integers(1, 6))):
                b[int(rng.integers(0, len(b)))] = int(rng.integers(0, 256))
        elif mode == 2:
            b = bytearray(rng.integers(0, 256, int(rng.integers(0, 100)), dtype=np.uint8).tobytes())
        else:
            b[0:8] = struct.pack("<Q", int(rng.integers(0, 2 ** 63 - 1)))          # a length field far beyond the file
        buf = bytes(b)
        for verify in (0, 1):
            off, ln = (ctypes.c_int64 * 8)(), (ctypes.c_int64 * 8)()
            n = lib.tt_tfrecord_scan(buf, len(buf), verify, off, ln, 8)
            assert n == -5 or 0 <= n <= 6
            for i in range(max(n, 0)):
                assert 0 <= off[i] and off[i] + ln[i] <= len(buf)
    feats = _features()
    w = TFRecordWriter(feats)
    base = [w._get_features_from_row({"customer_id": f"c{i}", "age": 0.25 * i, "article_id": f"{i * 7919:010d}"}) for i in range(8)]
    path = str(tmp_path / "f.tfrecord")
    parsed = rejected = 0
    for _ in range(1500):
        recs = []
        for _ in range(int(rng.integers(1, 5))):
            p = bytearray(base[int(rng.integers(0, 8))])
            mode = int(rng.integers(0, 5))
            if mode == 0:
                for _ in range(int(rng.integers(1, 4))):
                    p[int(rng.integers(0, len(p)))] = int(rng.integers(0, 256))
            elif mode == 1:
                p = p[:int(rng.integers(0, len(p)))]
            elif mode == 2:
                p = p + bytes(rng.integers(0, 256, int(rng.integers(1, 20)), dtype=np.uint8))
            elif mode == 3:
                p = bytearray(rng.integers(0, 256, int(rng.integers(0, 60)), dtype=np.uint8).tobytes())
            recs.append(_frame(lib, bytes(p)))
        with open(path, "wb") as fh:
            fh.write(b"".join(recs))
        try:
            cols = read_tfrecord_file(path, feats)
            assert all(len(v) == len(recs) for v in cols.values())
            parsed += 1
        except ValueError:
            rejected += 1
    assert parsed > 50 and rejected > 1000


def test_buffer_shuffle_order_properties():
    """tf.data buffer-shuffle semantics of TFRecordDataset._order (tfrecord_dataset.py:86-92 of the reference: ds.shuffle(size)): a
    permutation; an element is never emitted more than ``size - 1`` steps early; size 1 is the identity; the emission time of an
    element follows the distribution of the step-by-step simulation."""
    from pkg.modelling.tfrecord_dataset import TFRecordDataset

    def sequential(n, size, rng):
        buf, nxt, out = list(range(min(size, n))), min(size, n), []
        for _ in range(n):
            j = int(rng.integers(len(buf)))
            out.append(buf[j])
            if nxt < n:
                buf[j] = nxt
                nxt += 1
            else:
                buf[j] = buf[-1]
                buf.pop()
        return out

    for n in (0, 1, 2, 5, 50, 1000):
        for size in (1, 2, 7, 49, 50, 51, 5000):
            o = TFRecordDataset([], [], None, size, seed=n + size)._order(n)
            assert sorted(o.tolist()) == list(range(n))
            assert all(int(v) - i < min(size, n) for i, v in enumerate(o))
            if size == 1:
                assert o.tolist() == list(range(n))
    n, size, trials = 60, 8, 3000
    fast = np.array([np.argsort(TFRecordDataset([], [], None, size, seed=k)._order(n)) for k in range(trials)])
    slow = np.array([np.argsort(sequential(n, size, np.random.default_rng(10_000 + k))) for k in range(trials)])
    for element in (0, 10, 59):
        assert abs(fast[:, element].mean() - slow[:, element].mean()) < 0.6
        assert abs(fast[:, element].std() - slow[:, element].std()) < 0.5
