"""Drop-in for the reference's xlnet.py: `from xlnet import MAG_XLNetForSequenceClassification`
(multimodal_driver.py:30) resolves to the B200-native implementation."""
from global_configs import *  # noqa: F401,F403  (xlnet.py:6)
from bert_multimodal_transformer_b200.xlnet import MAG_XLNetForSequenceClassification, MAG_XLNetModel  # noqa: F401
from bert_multimodal_transformer_b200.modeling import MAG  # noqa: F401
