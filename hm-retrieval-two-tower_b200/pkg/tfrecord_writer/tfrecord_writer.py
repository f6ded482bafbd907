"""TFRecord writer without TensorFlow (reference pkg/tfrecord_writer/tfrecord_writer.py:11-126).

Same class and method names.  A row becomes a serialized ``tf.train.Example`` (hand-encoded protobuf wire format:
Example{1: Features{1: map<string, Feature>}}, Feature{1: BytesList{1: bytes} | 2: FloatList{1: packed float}}), framed
as a TFRecord by libtt (tt_tfrecord_frame: u64 length, masked CRC32C, payload, masked CRC32C).  Files written here are
read back by TensorFlow's own TFRecordDataset / parse_single_example and by pkg.modelling.tfrecord_dataset.

One deliberate difference: the reference's float32 branch uses ``val`` before assignment (tfrecord_writer.py:47-48), so it
cannot write numeric features at all; here the value is written as a one-element FloatList.
"""
from __future__ import annotations

import logging
import os
import struct
from typing import List, Optional, Union

import numpy as np

from pkg import _native as N
from pkg.schema import dtypes as tt
from pkg.schema.features import Feature

logger = logging.getLogger(__name__)


_ONE_BYTE = [bytes((i,)) for i in range(128)]


def _varint(n: int) -> bytes:
    if n < 128:                       # almost every length on this path: feature names, ids, one-float lists
        return _ONE_BYTE[n]
    out = bytearray()
    while True:
        b = n & 0x7F
        n >>= 7
        if n:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


_LD_TAG = [_varint((field << 3) | 2) for field in range(16)]


def _ld(field: int, payload: bytes) -> bytes:
    """length-delimited field"""
    return _LD_TAG[field] + _varint(len(payload)) + payload


class TFRecordWriter:
    def __init__(self, features: List[Feature]):
        self.features = features
        self._key_bytes = {}             # feature name -> serialized map key (constant per feature)

    def _parse_feature(self, feature_val: Union[str, float, int], dtype) -> bytes:
        """Serialized tf.train.Feature holding ``feature_val``; TypeError for an unsupported dtype (reference :49-53)."""
        if dtype == tt.string:
            val = feature_val if isinstance(feature_val, (bytes, np.bytes_)) else str(feature_val).encode()
            return _ld(1, _ld(1, bytes(val)))                                   # bytes_list { value: val }
        if dtype == tt.float32:
            return _ld(2, _ld(1, struct.pack("<f", float(feature_val))))        # float_list { value: [val] } (packed)
        raise TypeError(f"Invalid dtype {dtype} provided, must be one of {Feature.VALID_DTYPES}")

    def _get_features_from_row(self, row) -> bytes:
        """``row``: a NamedTuple from DataFrame.itertuples() or a mapping name -> value.  Returns Example bytes."""
        entries = []
        is_dict = isinstance(row, dict)
        for feature in self.features:
            v = row[feature.name] if is_dict else getattr(row, feature.name)
            key = self._key_bytes.get(feature.name)
            if key is None:
                key = self._key_bytes[feature.name] = _ld(1, feature.name.encode())
            entry = key + _ld(2, self._parse_feature(v, feature.dtype))                             # map entry {key, value}
            entries.append(_ld(1, entry))
        return _ld(1, b"".join(entries))                                        # Example { features { feature: ... } }

    @staticmethod
    def _rows(df, start: int, end: int):
        if hasattr(df, "iloc"):                      # pandas
            return df.iloc[start:end].itertuples()
        names = list(df.keys())                      # dict of equal-length arrays
        n = len(df[names[0]])
        return ({k: df[k][i] for k in names} for i in range(start, min(end, n)))

    def write_tfrecords(self, df, filepath: str, max_file_size: Optional[int] = None) -> None:
        """Write ``df`` (pandas DataFrame or dict of columns) as ``{filepath}_{i}.tfrecord`` partitions of at most
        ``max_file_size`` rows (reference :80-126)."""
        lib = N.load()
        filepath = filepath.replace(".tfrecord", "")
        os.makedirs(os.path.dirname(filepath) or ".", exist_ok=True)
        n = len(df) if hasattr(df, "iloc") else len(next(iter(df.values())))
        if max_file_size:
            num_files = n // max_file_size + (1 if n % max_file_size else 0)
            logger.info(f"Num files: {num_files} with max rows {max_file_size}")
        else:
            num_files, max_file_size = 1, n
            logger.info("Writing all rows to single file")
        for file_num in range(num_files):
            start, end = file_num * max_file_size, (file_num + 1) * max_file_size
            name = f"{filepath}_{file_num}.tfrecord"
            logger.info(f"Writing to TFRecord file: {name}")
            with open(name, "wb") as fh:
                for row in self._rows(df, start, end):
                    ex = self._get_features_from_row(row)
                    out = bytearray(len(ex) + 16)
                    N.check(lib.tt_tfrecord_frame(ex, len(ex), (N.ctypes.c_char * len(out)).from_buffer(out)), "tt_tfrecord_frame")
                    fh.write(out)
