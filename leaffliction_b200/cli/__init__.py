"""Command-line entry points with the reference's flags (srcs/cli/Augmentation.py:32-78,
srcs/cli/Transformation.py:568-608, srcs/cli/Distribution.py)."""
