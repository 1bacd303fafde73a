"""Empty stand-in: only algos/emlp_torch/reps/utils.py imports matplotlib."""
