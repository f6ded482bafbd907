"""TF-free stand-ins for the two dtypes the reference allows on a Feature (features.py:43)."""
from enum import Enum


class DType(Enum):
    string = "string"
    float32 = "float32"

    def __repr__(self) -> str:  # reads like tf.string / tf.float32 in messages
        return f"tt.{self.value}"


string = DType.string
float32 = DType.float32


def as_dtype(d) -> DType:
    """Accepts DType, the names "string"/"float32", numpy/python types, or a tf.DType-like object
    exposing ``.name`` (so a schema written against TensorFlow dtypes still loads)."""
    if isinstance(d, DType):
        return d
    name = getattr(d, "name", None) or (d if isinstance(d, str) else getattr(d, "__name__", None))
    if name in ("string", "str", "bytes", "object"):
        return DType.string
    if name in ("float32", "float"):
        return DType.float32
    raise TypeError(f"dtype must be one of {[DType.string, DType.float32]}, got {d}")
