from bitdance_b200.modeling.utils import MLPconnector  # noqa: F401
