"""Drop-in for the reference's global_configs.py (global_configs.py:1-19): same names, but it does
NOT pin CUDA_VISIBLE_DEVICES (the reference's line 4 would break one-process-per-GPU launches) and
DEVICE follows LOCAL_RANK."""
import os

import torch

os.environ.setdefault("WANDB_PROGRAM", "multimodal_driver.py")

_local_rank = int(os.environ.get("LOCAL_RANK", "0"))
DEVICE = torch.device(f"cuda:{_local_rank}" if torch.cuda.is_available() else "cpu")

# MOSI SETTING (global_configs.py:9-12); override with MAG_VISUAL_DIM / MAG_ACOUSTIC_DIM for MOSEI (35 / 74)
ACOUSTIC_DIM = int(os.environ.get("MAG_ACOUSTIC_DIM", "74"))
VISUAL_DIM = int(os.environ.get("MAG_VISUAL_DIM", "47"))
TEXT_DIM = 768

XLNET_INJECTION_INDEX = 1
