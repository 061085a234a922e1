"""Import-time names for ``from metrics2 import calc_aurc_eaurc, calc_nll_brier`` (code/fusion_train.py:30)."""


def calc_aurc_eaurc(*args, **kwargs):
    raise RuntimeError("metrics2.calc_aurc_eaurc is unpublished upstream and out of scope here")


def calc_nll_brier(*args, **kwargs):
    raise RuntimeError("metrics2.calc_nll_brier is unpublished upstream and out of scope here")
