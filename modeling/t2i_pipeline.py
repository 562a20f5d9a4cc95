from bitdance_b200.modeling.t2i_pipeline import IMAGE_SIZE_LIST, BitDanceT2IPipeline  # noqa: F401
