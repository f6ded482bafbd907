"""B200-native drop-in for the reference's ``pkg`` namespace (hot path only).

Like the reference, importing the package makes its INFO messages visible; unlike it, an application that has already
configured logging is left alone, and only the ``pkg`` logger tree is touched."""
import logging as _logging


def _default_logging() -> None:
    tree = _logging.getLogger(__name__)
    if tree.handlers or _logging.getLogger().handlers:
        return
    handler = _logging.StreamHandler()
    handler.setFormatter(_logging.Formatter("%(asctime)s | %(name)s | %(levelname)s | %(message)s"))
    tree.addHandler(handler)
    tree.setLevel(_logging.INFO)


_default_logging()
