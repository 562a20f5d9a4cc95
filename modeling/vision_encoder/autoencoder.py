from bitdance_b200.modeling.vision_encoder.autoencoder import VQModel  # noqa: F401
