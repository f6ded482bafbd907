"""
The one loss the reference trains with (pkg/modelling/runner.py:78-83):
``tf.keras.losses.CategoricalCrossentropy(from_logits=True, reduction=SUM)`` against ``eye(B)`` labels.
It is fused into the in-batch softmax kernel; this class only carries (and validates) the configuration.
"""


class Reduction:
    SUM = "sum"
    NONE = "none"
    SUM_OVER_BATCH_SIZE = "sum_over_batch_size"


class CategoricalCrossentropy:
    def __init__(self, from_logits: bool = False, reduction: str = Reduction.SUM_OVER_BATCH_SIZE, label_smoothing: float = 0.0):
        self.from_logits = bool(from_logits)
        self.reduction = getattr(reduction, "value", reduction)
        self.label_smoothing = float(label_smoothing)

    def validate(self) -> None:
        if not self.from_logits or str(self.reduction).lower() != Reduction.SUM or self.label_smoothing != 0.0:
            raise NotImplementedError(
                "only CategoricalCrossentropy(from_logits=True, reduction=SUM, label_smoothing=0) is implemented "
                "(the configuration of the reference runner)")
