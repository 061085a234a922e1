"""Import-time name for ``import ot`` (POT; code/fusion_net.py:5,9,12 import it and never use it)."""
