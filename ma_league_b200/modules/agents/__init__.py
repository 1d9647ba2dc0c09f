from .agent_network import AgentNetwork
from .drqn_agent import DRQNAgentNetwork

# reference registry: marl/modules/agents/__init__.py:5-8 ("dqn" is outside the hot path)
REGISTRY = {"rnn": DRQNAgentNetwork}
