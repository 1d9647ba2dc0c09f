"""Controller base (reference: marl/controllers/multi_agent_controller.py:15-76); COMA's `_softmax` is out of scope."""
from ..components.action_selectors import REGISTRY as action_REGISTRY


class MultiAgentController:
    def __init__(self, scheme, groups, args):
        self.n_agents = args.n_agents
        self.n_actions = args.n_actions
        self.args = args
        self.input_shape = self._get_input_shape(scheme)
        self.agent = self._build_agent(self.input_shape)
        self.agent_output_type = args.agent_output_type
        self.action_selector = action_REGISTRY[args.action_selector](args)
        self.hidden_states = None
        self.agent.trained_steps = 0
        if args.freeze_native:
            self.freeze_agent_weights()

    def _build_agent(self, input_shape):
        raise NotImplementedError()

    def _get_input_shape(self, scheme):
        raise NotImplementedError()

    def select_actions(self, ep_batch, t_ep, t_env, bs=slice(None), test_mode=False):
        raise NotImplementedError()

    def forward(self, ep_batch, t, test_mode=False):
        raise NotImplementedError()

    def init_hidden(self, batch_size):
        raise NotImplementedError()

    def parameters(self):
        raise NotImplementedError()

    def load_state(self, other_mac):
        raise NotImplementedError()

    def load_state_dict(self, agent):
        raise NotImplementedError()

    def cuda(self):
        raise NotImplementedError()

    def save_models(self, path, name):
        raise NotImplementedError()

    def load_models(self, path, name):
        raise NotImplementedError()

    def update_trained_steps(self, trained_steps):
        raise NotImplementedError()

    def freeze_agent_weights(self):
        for p in self.agent.parameters():
            p.requires_grad = False
