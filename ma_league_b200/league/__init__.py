from .agent_pool import DeviceAgentPool

__all__ = ["DeviceAgentPool"]
