"""Stand-in for the reference's unpublished ``glu2`` package (code/fusion_train.py:596,734)."""
