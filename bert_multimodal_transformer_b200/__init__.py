"""bert_multimodal_transformer_b200 -- B200-native (sm_100a) MAG-BERT / MAG-XLNet hot path.

Host code is Python; all arithmetic runs in hand-written CUDA kernels behind the C ABI of
libmagb200.so (include/magb200.h).  There is no CPU / PyTorch fallback.
"""
__version__ = "0.1.0"
