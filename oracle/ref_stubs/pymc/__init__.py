"""Import stub (oracle tooling only): posterior_model_priors.py imports pymc at module top for its M step; the fixtures
never run that step."""


def __getattr__(name):
    raise ImportError(f"pymc.{name} is not available in this environment (stub)")
