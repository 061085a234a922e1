"""Import-time name for ``from metrics import cal_ece`` (code/fusion_train.py:21; unpublished, only used by the
ensemble test that the MedFusion path never reaches)."""


def cal_ece(*args, **kwargs):
    raise RuntimeError("metrics.cal_ece is unpublished upstream and out of scope here")
