"""Drop-in for the reference's bert.py: `from bert import MAG_BertForSequenceClassification`
(multimodal_driver.py:29) resolves to the B200-native implementation."""
from global_configs import TEXT_DIM, ACOUSTIC_DIM, VISUAL_DIM, DEVICE  # noqa: F401  (bert.py:26)
from bert_multimodal_transformer_b200.bert import MAG_BertForSequenceClassification, MAG_BertModel  # noqa: F401
from bert_multimodal_transformer_b200.modeling import MAG  # noqa: F401
