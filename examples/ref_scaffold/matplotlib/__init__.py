"""No-op stand-in for matplotlib (absent from this image): the drivers only draw loss/accuracy curves with it
(code/fusion_train.py:65-76,120-135,771-772)."""
from . import pyplot  # noqa: F401
