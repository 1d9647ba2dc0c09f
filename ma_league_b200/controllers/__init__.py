from .basic_controller import BasicMAC
from .ensemble_agent_controller import EnsembleMAC

# reference registry: marl/controllers/__init__.py:6-11 ("distinct" and "gpe" are outside the hot path)
REGISTRY = {"basic": BasicMAC, "ensemble": EnsembleMAC}
