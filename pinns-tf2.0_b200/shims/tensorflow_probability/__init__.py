"""Imported (and unused) by 1d-burgers/inf_cont_burgers.py:6."""
