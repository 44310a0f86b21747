"""Stand-in for the picklable_itertools dependency of bin/run.py (only `extras.equizip` is used there)."""
