from .agent_network import AgentNetwork
from .dqn_agent import DQNAgentNetwork
from .drqn_agent import DRQNAgentNetwork

# reference registry: marl/modules/agents/__init__.py:5-8
REGISTRY = {"rnn": DRQNAgentNetwork, "dqn": DQNAgentNetwork}
