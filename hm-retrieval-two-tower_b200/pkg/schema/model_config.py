from pkg.schema.config import ModelConfig  # noqa: F401  (reference import path)
