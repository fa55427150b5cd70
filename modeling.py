"""Drop-in for the reference's modeling.py: `from modeling import MAG` (bert.py:24, xlnet.py:7,
README.md:76-88) resolves to the B200-native implementation."""
from global_configs import *  # noqa: F401,F403  (the reference star-imports these names, modeling.py:4)
from bert_multimodal_transformer_b200.modeling import MAG  # noqa: F401
