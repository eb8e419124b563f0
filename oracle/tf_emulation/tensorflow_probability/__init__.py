"""Placeholder: the reference imports tensorflow_probability but only mentions it in comments (utils/neuralnetwork.py:121-128).
TEST INFRASTRUCTURE ONLY (oracle/tf_emulation)."""
