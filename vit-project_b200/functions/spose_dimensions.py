"""The 66 SPoSE dimension prompts (reference: Training/functions/spose_dimensions.py, DIMS:1-68),
kept as a plain-text resource, one prompt per line, in the reference order.  They are tokenised
once (NEW:282), so the text tower's input is constant for the whole study."""
import os

with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "spose66_prompts.txt")) as _f:
    classnames66 = [line.rstrip("\n") for line in _f if line.strip()]

assert len(classnames66) == 66
