from bitdance_b200.modeling.utils import (MLPconnector, remove_first_user_block, sample_codebook,  # noqa: F401
                                          top_k_top_p_filtering)
