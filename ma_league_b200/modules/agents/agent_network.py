import torch.nn as nn


class AgentNetwork(nn.Module):
    """Agent interface (reference: marl/modules/agents/agent_network.py:5-28)."""

    def __init__(self, input_shape, args):
        super().__init__()
        self.args = args
        self.input_shape = input_shape
        self._trained_steps_host = 0
        self._trained_steps_dev = None    # int64 device counter fed by the learner without a host sync

    @property
    def trained_steps(self):
        extra = int(self._trained_steps_dev.item()) if self._trained_steps_dev is not None else 0
        return self._trained_steps_host + extra

    @trained_steps.setter
    def trained_steps(self, value):
        self._trained_steps_host = int(value)
        if self._trained_steps_dev is not None:
            self._trained_steps_dev.zero_()   # in place: a captured CUDA graph of the learner step keeps adding into it

    def init_hidden(self):
        raise NotImplementedError()

    def forward(self, inputs, hidden_state):
        raise NotImplementedError()

    def count_parameters(self):
        return sum(p.numel() for p in self.parameters() if p.requires_grad)

    def print_parameters(self):
        for name, param in self.state_dict().items():
            print(name, param)
