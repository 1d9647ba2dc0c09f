from .basic_controller import BasicMAC

# reference registry: marl/controllers/__init__.py:6-11 (the other entries are outside the hot path)
REGISTRY = {"basic": BasicMAC}
