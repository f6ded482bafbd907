from pkg.schema.config import TrainingConfig  # noqa: F401  (reference import path)
