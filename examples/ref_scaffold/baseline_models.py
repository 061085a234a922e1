"""Import-time names for ``from baseline_models import ...`` (code/fusion_train.py:18).  The reference's comparison zoo
(root ``baseline_models.py``) needs unpublished ``Models.*`` and pretrained weights at absolute paths and is out of scope
(SURVEY.md section 2); the MedFusion path never constructs these."""


def _unavailable(name):
    class _Missing:
        def __init__(self, *a, **k):
            raise RuntimeError(f"{name}: the baseline zoo is out of scope of this scaffolding (SURVEY.md section 2)")
    _Missing.__name__ = name
    return _Missing


Res2Net2D = _unavailable("Res2Net2D")
ResNet3D = _unavailable("ResNet3D")
Multi_ResNet = _unavailable("Multi_ResNet")
Multi_EF_ResNet = _unavailable("Multi_EF_ResNet")
Multi_CBAM_ResNet = _unavailable("Multi_CBAM_ResNet")
Multi_dropout_ResNet = _unavailable("Multi_dropout_ResNet")
