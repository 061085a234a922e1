"""Stand-in for the reference's unpublished ``DR_2`` package (code/fusion_train.py:551,731)."""
