from bitdance_b200.modeling.vision_head.flow_head_parallel_x import DiffHead  # noqa: F401
