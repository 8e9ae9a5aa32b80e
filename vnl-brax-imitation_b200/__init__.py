"""vnl-brax-imitation_b200: B200-native fused physics + imitation-reward step.

Import with `importlib.import_module("vnl-brax-imitation_b200")` or through the
`vnl_b200` alias module at the repo root.
"""
__version__ = "0.1.0"
