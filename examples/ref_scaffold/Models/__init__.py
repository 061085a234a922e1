"""Stand-in for the reference's unpublished ``Models`` package: only the two encoder factories
``code/fusion_net.py:1-2`` imports, with the output contracts MedFusion relies on."""
