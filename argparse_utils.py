"""Drop-in for the reference's argparse_utils.py (argparse_utils.py:5-31): the two argparse `type=`
helpers multimodal_driver.py:35-57 imports, with the same accepted values and the same
ArgumentTypeError texts.  Quirk kept on purpose: `seed` accepts an *int instance* or the string
"random" only, so `--seed 42` (a str) is rejected exactly as in the reference (argparse_utils.py:27-31)."""
import argparse
import random

_TRUE = {"yes", "true", "t", "y", "1"}
_FALSE = {"no", "false", "f", "n", "0"}
_SEED_MAX = 9999


def str2bool(s):
    if isinstance(s, bool):
        return s
    word = s.lower()
    if word in _TRUE:
        return True
    if word in _FALSE:
        return False
    raise argparse.ArgumentTypeError("Boolean value expected. Recieved {0}".format(s))


def seed(s):
    if s == "random":
        return random.randint(0, _SEED_MAX)
    if isinstance(s, int):
        if 0 <= s <= _SEED_MAX:
            return s
        raise argparse.ArgumentTypeError("Seed must be between 0 and 2**32 - 1. Received {0}".format(s))
    raise argparse.ArgumentTypeError("Integer value is expected. Recieved {0}".format(s))
