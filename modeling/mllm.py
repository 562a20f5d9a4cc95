"""Drop-in module: ``from modeling.mllm import MLLModel`` resolves to the B200-native mirror (inference surface of the
image path: gen_image / gen_image_block_causal / encode_image / decode_image / get_2d_embed, and the interleaved
text+image inference forward / forward_inference / forward_inference_block_causal)."""
from bitdance_b200.modeling.mllm import MLLModel  # noqa: F401
