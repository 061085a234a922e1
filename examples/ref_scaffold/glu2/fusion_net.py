"""Stand-in for ``glu2.fusion_net`` (code/fusion_train.py:734): the same module as ``DR_2.fusion_net``."""
from DR_2.fusion_net import *  # noqa: F401,F403
from DR_2.fusion_net import MedFusion, STATS  # noqa: F401
